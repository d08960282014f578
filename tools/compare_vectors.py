"""Compares two Kaldi text archives of vectors (or of matrices): per key max|a-b|/max|b| and cosine.
  python tools/compare_vectors.py <got.txt> <expected.txt> [--tol-rel 1e-4] [--tol-cos 0.999999]
Exit status 0 iff the key lists agree and every vector is within tolerance."""
import argparse
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
kf = importlib.import_module("speaker-embedding-with-phonetic-information_b200.kaldi_format")

ap = argparse.ArgumentParser()
ap.add_argument("got")
ap.add_argument("expected")
ap.add_argument("--tol-rel", type=float, default=1e-4, help="fp32 Kaldi vs the fp64 oracle: ~1e-5 expected")
ap.add_argument("--tol-cos", type=float, default=0.999999)
a = ap.parse_args()
got, want = dict(kf.read_ark(a.got)), dict(kf.read_ark(a.expected))
ok = list(got) == list(want)
if not ok:
    print(f"KEYS DIFFER: got {list(got)} expected {list(want)}")
worst_rel, worst_cos = 0.0, 1.0
for k in want:
    if k not in got:
        continue
    x, y = np.asarray(got[k], np.float64).ravel(), np.asarray(want[k], np.float64).ravel()
    if x.shape != y.shape:
        print(f"{k}: shape {x.shape} vs {y.shape}")
        ok = False
        continue
    rel = float(np.max(np.abs(x - y)) / np.max(np.abs(y)))
    cos = float(x @ y / (np.linalg.norm(x) * np.linalg.norm(y)))
    worst_rel, worst_cos = max(worst_rel, rel), min(worst_cos, cos)
    print(f"{k}: max rel err {rel:.3e}  cosine {cos:.9f}")
    ok = ok and rel <= a.tol_rel and cos >= a.tol_cos
print(f"{'MATCH' if ok else 'MISMATCH'}: {len(want)} entries, worst max-rel {worst_rel:.3e}, worst cosine {worst_cos:.9f} "
      f"(tolerance {a.tol_rel:g} / {a.tol_cos})")
sys.exit(0 if ok else 1)
