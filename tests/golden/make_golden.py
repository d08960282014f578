"""Generates the committed golden fixtures under tests/golden/:

  v2_small_*   v2 x-vector topology (run_xvector_new.sh:94-114), widths shrunk to
               hidden=128 / stats=300 so the binary model stays < 1 MB
  v5_small_*   v5 c-vector topology (train_cvector_with_am.sh:65-89) with the
               phonetic branch, hidden=96 / stats=150 / am 80 -> 24 (dims that are
               not multiples of 64, to exercise every padding path)

For each: a binary nnet3 model in genuine Kaldi format (orphan layers and heads
kept, as in a real final.raw), a feature archive, extract.config, and the
ORACLE's fp64 embeddings; for v5_small also the phonetic bottleneck at every frame
(v5_small_frames.npz: what `nnet3-compute` on output tdnn5.batchnorm writes,
extract_am_embedding.sh:79-85).  The reference ships no golden vectors for this path
(parity unpinned); these pin the oracle and the CUDA path against
regressions and let a Kaldi owner replay e.g.

  nnet3-xvector-compute --chunk-size=300 --min-chunk-size=25 \
    "nnet3-copy --nnet-config=v2_small_extract.config v2_small_final.raw - |" \
    ark:v2_small_feats.ark ark,t:-

and compare with the `embeddings` array.  Run from the repo root:
  python tests/golden/make_golden.py
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
P = "speaker-embedding-with-phonetic-information_b200"
kf = importlib.import_module(P + ".kaldi_format")
sy = importlib.import_module(P + ".synth")
from oracle.nnet3_oracle import Oracle, extract_all  # noqa: E402

here = os.path.dirname(os.path.abspath(__file__))
SPECS = {
    "v2_small": dict(topology="v2_xvector", node="tdnn6.affine", kw=dict(hidden=128, stats_dim=300)),
    "v5_small": dict(topology="v5_cvector", node="tdnn6_xvec.affine",
                     kw=dict(hidden=96, stats_dim=150, am_hidden=80, am_bottleneck=24)),
}
LENGTHS = [300, 25, 24, 731, 57, 0, 45]
CHUNK, MIN_CHUNK = 300, 25

for name, spec in SPECS.items():
    model = sy.make_model(spec["topology"], seed=42, num_speakers=16, num_senones=24, **spec["kw"])
    kf.write_nnet3(os.path.join(here, f"{name}_final.raw"), model, binary=True)
    model = kf.read_nnet3(os.path.join(here, f"{name}_final.raw"))   # what every reader of the file sees (fp32 scalars)
    feats = sy.make_features(LENGTHS, 23, seed=1234)
    emb, ok = extract_all(Oracle(model, spec["node"], np.float64), {"input": feats}, LENGTHS,
                          chunk_size=CHUNK, min_chunk_size=MIN_CHUNK)
    emb_pad, ok_pad = extract_all(Oracle(model, spec["node"], np.float64), {"input": feats}, LENGTHS,
                                  chunk_size=CHUNK, min_chunk_size=MIN_CHUNK, pad_input=True)
    np.savez_compressed(os.path.join(here, f"{name}.npz"), feats=feats, lengths=np.array(LENGTHS), embeddings=emb,
                        ok=ok, embeddings_pad_input=emb_pad, ok_pad_input=ok_pad, chunk_size=CHUNK,
                        min_chunk_size=MIN_CHUNK, output_node=spec["node"])
    starts = np.cumsum([0] + LENGTHS[:-1])
    kf.write_ark(os.path.join(here, f"{name}_feats.ark"),
                 [(k, feats[s:s + n]) for k, s, n in zip(sy.utt_keys(len(LENGTHS)), starts, LENGTHS)])
    with open(os.path.join(here, f"{name}_extract.config"), "w") as f:
        f.write(f"output-node name=output input={spec['node']}\n")
    print(name, emb.shape, ok.tolist(), ok_pad.tolist())
    if name == "v5_small":
        node = "tdnn5.batchnorm"
        orc = Oracle(model, node, np.float64)
        rows = [orc.forward_frames({"input": feats[s:s + n]}) for s, n in zip(starts, LENGTHS) if n > 0]
        np.savez_compressed(os.path.join(here, "v5_small_frames.npz"), output_node=node,
                            frames=np.concatenate(rows).astype(np.float32), context=np.array(orc.frame_context()))
        print("v5_small_frames", np.concatenate(rows).shape, orc.frame_context())
