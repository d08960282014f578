"""Parity margins of the CUDA path against the fp64 oracle, per topology (fp16 operands):
max|a-b|/max|b| and the minimum per-utterance cosine.  `python tools/parity_report.py [repo_root]`
runs the library built under another checkout (A/B of kernel changes)."""
import importlib
import os
import sys
import tempfile

import numpy as np

root = os.path.abspath(sys.argv[1]) if len(sys.argv) > 1 else os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root)
P = "speaker-embedding-with-phonetic-information_b200"
cvx = importlib.import_module(P)
kf = importlib.import_module(P + ".kaldi_format")
sy = importlib.import_module(P + ".synth")
from oracle.nnet3_oracle import Oracle, extract_all, parity_metrics  # noqa: E402

d = tempfile.mkdtemp(prefix="cvx_parity_")
lengths = [25, 300, 57, 1234, 100, 31, 640, 3000]
for topo in ("v2_xvector", "v3_multitask", "v4_cvector", "v4_two_input", "v5_cvector"):
    m = sy.make_model(topo)
    path = os.path.join(d, topo + ".raw")
    kf.write_nnet3(path, m, True)
    node = sy.EMBEDDING_NODE[topo]
    mod = cvx.Model(path, node)
    feats = {n: sy.make_features(lengths, dim, 11 + i) for i, (n, dim) in enumerate(zip(mod.input_names, mod.input_dims))}
    got, ok = cvx.Extractor(mod).extract(feats, lengths, chunk_size=1000, min_chunk_size=25)
    ref, _ = extract_all(Oracle(m, node, np.float64), feats, lengths, chunk_size=1000, min_chunk_size=25)
    r, c = parity_metrics(got[ok], ref[ok])
    print(f"{topo:14s} max rel err {r:.3e}  min cosine {c:.8f}", flush=True)
