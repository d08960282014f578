"""Where do the ~15 ms per step outside the kernels go on the short-segment workload?
Times cvx_extract on 125 000 x 200-frame segments (x-vector net, features resident in HBM) with the
embeddings going to (a) pageable host memory, (b) pinned host memory, (c) device memory."""
import ctypes as C
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
P = "speaker-embedding-with-phonetic-information_b200"
cvx = importlib.import_module(P)
kf = importlib.import_module(P + ".kaldi_format")
sy = importlib.import_module(P + ".synth")
import torch

topology = sys.argv[1] if len(sys.argv) > 1 else "v2_xvector"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 125000
path = f"/tmp/cvx_probe_{os.getpid()}.raw"
kf.write_nnet3(path, sy.make_model(topology, seed=42), binary=True)
model = cvx.Model(path, sy.EMBEDDING_NODE[topology])
os.unlink(path)
rows = n * 200
x = torch.randn((rows, 23), device="cuda")
starts = np.arange(n, dtype=np.int64) * 200
lens = np.full(n, 200, np.int32)
utts = np.arange(n, dtype=np.int32)
for profile in (True, False):
    ex = cvx.Extractor(model, precision="fp16", profile=profile)
    pageable = np.zeros((n, 512), np.float32)
    pinned = cvx.PinnedArray((n, 512))
    dev = torch.zeros((n, 512), device="cuda")
    for label, out, flags in (("pageable host", pageable, cvx.FLAG_DEVICE_INPUT), ("pinned host", pinned.array, cvx.FLAG_DEVICE_INPUT),
                              ("device", dev.data_ptr(), cvx.FLAG_DEVICE_INPUT | cvx.FLAG_DEVICE_OUTPUT)):
        ms, wall = [], []
        for i in range(5):
            t0 = time.perf_counter()
            ex.extract_chunks([x.data_ptr()], rows, starts, lens, utts, None, n, out=out, flags=flags)
            wall.append((time.perf_counter() - t0) * 1e3)
            ms.append(ex.last_extract_ms())
        print(f"{topology} profile={profile} out={label}: device {np.mean(ms[2:]):.2f} ms, wall {np.mean(wall[2:]):.2f} ms "
              f"-> {rows / 100 / (np.mean(ms[2:]) / 1e3) / 1e6:.3f} M audio-s/s", flush=True)
    if profile:
        print({p["name"]: round(p["total_ms"] / 15, 3) for p in ex.profile() if not p["name"].startswith("layer:")})
    ex.close()
