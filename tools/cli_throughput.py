"""End-to-end throughput of the drop-in program (ark file / pipe in -> ark,scp out).
Writes a synthetic feature archive (v5 c-vector net, utterances 10-300 s), runs
bin/nnet3-xvector-compute on it the way extract_xvectors_new.sh does and reports
audio-seconds per second of wall clock (model load and GPU init included)."""
import importlib
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
P = "speaker-embedding-with-phonetic-information_b200"
cvx = importlib.import_module(P)
kf = importlib.import_module(P + ".kaldi_format")
sy = importlib.import_module(P + ".synth")

n_utts = int(sys.argv[1]) if len(sys.argv) > 1 else 400
num_gpus = int(sys.argv[2]) if len(sys.argv) > 2 else 1   # > 1: also run --num-gpus=N and compare the archives
d = tempfile.mkdtemp(prefix="cvx_cli_")
model = os.path.join(d, "final.raw")
kf.write_nnet3(model, sy.make_model("v5_cvector"), binary=True)
lengths = sy.make_lengths(n_utts, 1000, 30000, 7)
rng = np.random.default_rng(3)
ark = os.path.join(d, "feats.ark")
with open(ark, "wb") as f:
    for k, T in zip(sy.utt_keys(n_utts), lengths):
        kf.write_ark(f, [(k, rng.standard_normal((int(T), 23), dtype=np.float32))])
frames = int(lengths.sum())
exe = os.path.join(cvx.BIN_DIR, "nnet3-xvector-compute")
for label, feat, extra in (("ark file", f"ark:{ark}", []), ("ark through a pipe", f"ark:cat {ark} |", []),
                           ("ark file, sliding CMN on the GPU (raw mode)", f"ark:{ark}", ["--cmn-window=300"])):
    t0 = time.perf_counter()
    r = subprocess.run([exe, "--use-gpu=no", "--min-chunk-size=25", "--chunk-size=10000", "--output-node=tdnn6_xvec.affine",
                        *extra, model, feat, f"ark,scp:{d}/x{len(extra)}.ark,{d}/x{len(extra)}.scp"], capture_output=True, text=True)
    dt = time.perf_counter() - t0
    assert r.returncode == 0, r.stderr[-2000:]
    tail = [ln for ln in r.stderr.splitlines() if "Time taken" in ln or "Done " in ln]
    print(f"{label}: {frames} frames ({os.path.getsize(ark) / 1e9:.2f} GB) in {dt:.2f} s wall -> {frames / 100 / dt:,.0f} audio-s/s | "
          + " | ".join(t.split(") ", 1)[-1] for t in tail), flush=True)

if num_gpus > 1:
    t0 = time.perf_counter()
    r = subprocess.run([exe, "--use-gpu=no", "--min-chunk-size=25", "--chunk-size=10000", "--output-node=tdnn6_xvec.affine",
                        f"--num-gpus={num_gpus}", model, f"ark:{ark}", f"ark,scp:{d}/xn.ark,{d}/xn.scp"],
                       capture_output=True, text=True)
    dt = time.perf_counter() - t0
    assert r.returncode == 0, r.stderr[-2000:]
    tail = [ln for ln in r.stderr.splitlines() if "Time taken" in ln or "Done " in ln]
    same = open(f"{d}/x0.ark", "rb").read() == open(f"{d}/xn.ark", "rb").read()
    print(f"ark file, --num-gpus={num_gpus}: {dt:.2f} s wall -> {frames / 100 / dt:,.0f} audio-s/s | "
          + " | ".join(t.split(") ", 1)[-1] for t in tail) + f" | archive identical to the 1-GPU run: {same}", flush=True)
