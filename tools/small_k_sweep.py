"""Tuning aid: single-tap layers with short K (the first TDNN layers) -- time, TFLOP/s and the
output / total bytes per second they move, to tell an epilogue bound from a bandwidth bound."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
cvx = importlib.import_module("speaker-embedding-with-phonetic-information_b200")
rows = 524288
for d, n in [(64, 512), (128, 512), (256, 512), (512, 512), (128, 256), (128, 1024), (128, 672), (64, 256)]:
    ms = cvx.bench_tdnn_layer(rows, d, [0], n, iters=20)
    out_b = rows * n * 2; in_b = rows * d * 2
    print(f"K={d:4d} N={n:5d}: {ms:7.3f} ms  {2.0*rows*d*n/ms/1e9:7.1f} TF/s  out {out_b/ms/1e6:6.0f} GB/s  in+out {(in_b+out_b)/ms/1e6:6.0f} GB/s", flush=True)
