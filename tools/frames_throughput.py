"""Throughput of the frame-level path (cvx_compute_frames, the nnet3-compute drop-in):
bottleneck features of the phonetic TDNN (v4 c-vector net, output tdnn5.batchnorm, 128-d) and
log-posteriors of a senone head, utterances of 10-300 s, pinned host buffers both ways.
Reports frames/s by wall clock around the C-ABI call and the device time of the call."""
import importlib
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
P = "speaker-embedding-with-phonetic-information_b200"
cvx = importlib.import_module(P)
kf = importlib.import_module(P + ".kaldi_format")
sy = importlib.import_module(P + ".synth")

n_utts = int(sys.argv[1]) if len(sys.argv) > 1 else 300
d = tempfile.mkdtemp(prefix="cvx_frames_")
lengths = sy.make_lengths(n_utts, 1000, 30000, 7)
frames = int(lengths.sum())
pin_in = cvx.PinnedArray((frames, 23))
pin_in.array[:] = np.random.default_rng(1).standard_normal((frames, 23), dtype=np.float32)
x = {"input": pin_in.array}
for label, topo, node, kw in (("AM bottleneck 128-d", "v4_cvector", "tdnn5.batchnorm", {}),
                              ("log-posteriors, 2000 senones", "v5_cvector", "output_am.log-softmax", {"num_senones": 2000})):
    path = os.path.join(d, topo + ".raw")
    kf.write_nnet3(path, sy.make_model(topo, random_outputs=True, **kw), binary=True)
    mod = cvx.Model(path, node, frames=True)
    ex = cvx.Extractor(mod)
    ex.compute_frames({"input": x["input"][:200000]}, [200000])  # warm-up
    pin_out = cvx.PinnedArray((frames, mod.info.embedding_dim))
    best = None
    for _ in range(3):
        t0 = time.perf_counter()
        out = ex.compute_frames(x, lengths, out=pin_out.array)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    print(f"{label}: {frames} frames, out {out.shape[1]}-d, MAC/frame {mod.info.mac_per_frame}: "
          f"{best * 1e3:.1f} ms wall ({frames / best / 1e6:.2f} M frames/s = {frames / best / 100:,.0f} audio-s/s), "
          f"device {ex.last_extract_ms():.1f} ms", flush=True)
