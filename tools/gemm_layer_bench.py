"""Per-layer micro-benchmark of the TDNN GEMM kernel (tuning aid, not the headline bench).
Runs each frame-layer shape of the v5 c-vector net on `rows` rows and prints ms,
algorithmic TFLOP/s and MMA-issue efficiency.  Kernel variants are selected with the
environment knobs the engine reads (CVX_GEMM_1CTA, CVX_GEMM_STAGES)."""
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
cvx = importlib.import_module("speaker-embedding-with-phonetic-information_b200")

LAYERS = [  # name, in_dim, offsets, out_dim
    ("tdnn2_xvec 1536->512", 512, [-2, 0, 2], 512),
    ("tdnn4_xvec 512->512", 512, [0], 512),
    ("am tdnn2 1950->650", 650, [-1, 0, 1], 650),
    ("am tdnn5 1950->128", 650, [-6, -3, 0], 128),
    ("tdnn5_xvec* 512->1500", 512, [0], 1500),
    ("tdnn1 (128)->512 tap path", 128, [0], 512),
]
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
tag = f"1cta={os.environ.get('CVX_GEMM_1CTA', '0')} stages={os.environ.get('CVX_GEMM_STAGES', 'max')}"
for name, d, offs, n in LAYERS:
    ms = cvx.bench_tdnn_layer(rows, d, offs, n, iters=20)
    fl = 2.0 * rows * d * len(offs) * n
    print(f"{tag:22s} {name:28s} {ms:8.3f} ms  {fl / ms / 1e9:8.1f} TFLOP/s", flush=True)
