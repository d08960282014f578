"""Profiling driver for ncu: ONE device batch per call, so kernel launches line up with the plan.

  python tools/profile_step.py [--topology v5_cvector] [--utts 50] [--frames 10000] [--calls 4] [--fused-pool 1]

Every call runs, in this order: pool_meta, stage_input (one per staging buffer), the frame layers in
plan order (tdnn_gemm_halo_kernel / tdnn_gemm_pair_kernel; the pooled layer is the kPool instance),
stats_pool_finalize_fused, split_f32 + the segment GEMM (tdnn_gemm_pair_kernel<.., kOutF32, ..>),
utt_average.  With the default 50 x 10 000-frame utterances a call is one 500 000-row batch of the
headline net, i.e. `-k regex:tdnn_gemm` sees 10 frame layers + 1 segment GEMM per call.
Writes gpurun_out/<tag>_plan.json (layer order, shapes, algorithmic bytes and FLOPs per launch) for
tools/summarize_ncu.py."""
import argparse
import importlib
import json
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
P = "speaker-embedding-with-phonetic-information_b200"
cvx = importlib.import_module(P)
kf = importlib.import_module(P + ".kaldi_format")
sy = importlib.import_module(P + ".synth")

ap = argparse.ArgumentParser()
ap.add_argument("--topology", default="v5_cvector")
ap.add_argument("--utts", type=int, default=50)
ap.add_argument("--frames", type=int, default=10000)
ap.add_argument("--chunk-size", type=int, default=10000)
ap.add_argument("--calls", type=int, default=4)
ap.add_argument("--tag", default="r2")
a = ap.parse_args()

tmp = os.environ.get("TMPDIR", "/tmp")
path = os.path.join(tmp, f"cvx_prof_{os.getpid()}.raw")
kf.write_nnet3(path, sy.make_model(a.topology, seed=42), binary=True)
model = cvx.Model(path, sy.EMBEDDING_NODE[a.topology])
os.unlink(path)
lengths = [a.frames] * a.utts
rows = a.frames * a.utts
feats = {n: sy.make_features(lengths, d, 5 + i) for i, (n, d) in enumerate(zip(model.input_names, model.input_dims))}

# plan -> per-launch algorithmic work (bytes at the stored pitch: every source buffer read once,
# the output written once, weights once; the pooled layer writes partials instead of activations)
ru = lambda v, m: (v + m - 1) // m * m
layers, out_ld = [], {}
for ln in model.describe().splitlines():
    m = re.match(r"frame\[(\d+)\] (\S+) (\d+)->(\d+) taps=\{([^}]*)\}", ln)
    if not m:
        continue
    i, name, K, N = int(m.group(1)), m.group(2), int(m.group(3)), int(m.group(4))
    taps = m.group(5).split(",")
    srcs = sorted({t.split("@")[0] for t in taps})
    all_inputs = all(s.startswith("in") for s in srcs)
    in_bytes = 0
    for s in srcs:
        if s.startswith("in"):
            d = model.input_dims[int(s[2:])]
            ld = ru(sum(model.input_dims[int(t.split("@")[0][2:])] for t in taps), 64) if all_inputs else ru(d, 64)
        else:
            ld = out_ld[int(s[1:])]
        in_bytes += rows * ld * 2
        if all_inputs:
            break
    out_ld[i] = ru(N, 64)
    layers.append({"index": i, "name": name, "K": K, "N": N, "taps": len(taps), "flops": 2.0 * rows * K * N,
                   "in_bytes": in_bytes, "out_bytes": rows * out_ld[i] * 2, "w_bytes": ru(N, 64) * K * 2})
pm = re.search(r"pool src=L(\d+)", model.describe())
if pm:
    pl = layers[int(pm.group(1))]
    pl["pooled"] = True
    pl["out_bytes"] = (rows // 32 + a.utts) * 2 * pl["N"] * 4      # per-slot partials instead of the activation
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump({"topology": a.topology, "rows": rows, "utts": a.utts, "frames": a.frames, "layers": layers,
           "describe": model.describe()}, open(os.path.join(ROOT, "gpurun_out", f"{a.tag}_plan.json"), "w"), indent=1)

ex = cvx.Extractor(model, precision="fp16")
for c in range(a.calls):
    emb, ok = ex.extract(feats, lengths, chunk_size=a.chunk_size, min_chunk_size=25)
assert ok.all() and np.isfinite(emb).all()
print("profile_step:", a.topology, rows, "rows,", a.calls, "calls,", ex.launch_count(), "launches")
