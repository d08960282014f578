#!/bin/bash
# Tuning aid: per-kernel profile of the drop-in program in raw-feature mode (device pre-step).
set -e
D=$(mktemp -d)
python - "$D" <<'PY'
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
P = "speaker-embedding-with-phonetic-information_b200"
kf = importlib.import_module(P + ".kaldi_format"); sy = importlib.import_module(P + ".synth")
d = sys.argv[1]
kf.write_nnet3(os.path.join(d, "final.raw"), sy.make_model("v5_cvector"), binary=True)
lengths = sy.make_lengths(600, 1000, 30000, 7)
rng = np.random.default_rng(3)
with open(os.path.join(d, "feats.ark"), "wb") as f:
    for k, T in zip(sy.utt_keys(600), lengths):
        kf.write_ark(f, [(k, rng.standard_normal((int(T), 23), dtype=np.float32))])
PY
B=speaker-embedding-with-phonetic-information_b200/bin
for extra in "" "--cmn-window=300"; do
  echo "== extra: $extra"
  $B/nnet3-xvector-compute --profile=true --min-chunk-size=25 --chunk-size=10000 --output-node=tdnn6_xvec.affine $extra \
     $D/final.raw ark:$D/feats.ark ark:$D/x.ark 2>&1 | grep -E "profile|Time taken" | sed 's/.*) //'
done
rm -rf $D
