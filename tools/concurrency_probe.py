"""Diagnostic: do concurrent drop-in jobs slow each other down, and where?  Runs the program on a
synthetic table (v5 net) alone, then G copies at once (one per GPU, --gpu-id=g), and prints each
job's lane-timing line (--verbose=1).
  python tools/concurrency_probe.py [n_utts=600] [gpus=2]"""
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
n_utts = int(sys.argv[1]) if len(sys.argv) > 1 else 600
gpus = int(sys.argv[2]) if len(sys.argv) > 2 else 2
d = tempfile.mkdtemp(prefix="cvx_probe_", dir="/dev/shm" if os.access("/dev/shm", os.W_OK) else None)
model = os.path.join(d, "final.raw")
kf.write_nnet3(model, sy.make_model("v5_cvector"), binary=True)
lengths = sy.make_lengths(n_utts, 1000, 30000, 7)
rng = np.random.default_rng(3)
arks = []
for g in range(gpus):
    ark, scp = os.path.join(d, f"feats{g}.ark"), os.path.join(d, f"feats{g}.scp")
    with open(ark, "wb") as f, open(scp, "w") as s:
        for k, T in zip(sy.utt_keys(n_utts), lengths):
            f.write(k.encode() + b" ")
            s.write(f"{k} {ark}:{f.tell()}\n")
            f.write(b"\0BFM \x04" + np.int32(T).tobytes() + b"\x04" + np.int32(23).tobytes())
            f.write(rng.standard_normal((int(T), 23), dtype=np.float32).tobytes())
    arks.append(scp)
frames = int(lengths.sum())
exe = os.path.join(cvx.BIN_DIR, "nnet3-xvector-compute")
base = [exe, "--verbose=1", "--min-chunk-size=25", "--chunk-size=10000", "--output-node=tdnn6_xvec.affine"]


def run_many(label, cmds):
    t0 = time.perf_counter()
    ps = [subprocess.Popen(c, stderr=subprocess.PIPE, text=True) for c in cmds]
    errs = [p.communicate()[1] for p in ps]
    dt = time.perf_counter() - t0
    print(f"== {label}: {len(cmds)} job(s) x {frames} frames, {dt:.2f} s wall", flush=True)
    for e in errs:
        for ln in e.splitlines():
            if "lane timing" in ln or "Time taken" in ln or "ERROR" in ln:
                print("   " + ln.split(") ", 1)[-1], flush=True)


for rep in range(2):
    run_many(f"alone on GPU 0 (run {rep})", [base + ["--gpu-id=0", model, f"scp:{arks[0]}", f"ark:{d}/o0.ark"]])
run_many(f"{gpus} jobs at once, one per GPU", [base + [f"--gpu-id={g}", model, f"scp:{arks[g]}", f"ark:{d}/o{g}.ark"] for g in range(gpus)])
run_many(f"{gpus} jobs at once, one per GPU, device-resident pre-step off, batch-rows 4M",
         [base + [f"--gpu-id={g}", "--batch-rows=4194304", model, f"scp:{arks[g]}", f"ark:{d}/o{g}.ark"] for g in range(gpus)])
import shutil
shutil.rmtree(d, ignore_errors=True)
