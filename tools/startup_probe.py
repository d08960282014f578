"""Diagnostic: start-up cost of G drop-in jobs launched together with NO device hint (the recipe's GPU branch,
extract_xvectors_new.sh:83-89), device picked (a) through CUDA enumeration (CVX_SELECT_WITH_CUDA=1: every job
initialises every visible GPU first) and (b) before CUDA starts (default: among the CUDA_VISIBLE_DEVICES entries /
/proc bus ids; the job then sees one device).  Prints every job's start-up and lane-timing lines.
  python tools/startup_probe.py [gpus=8] [n_utts=600] [all|short]
(c) the same work from ONE process with --num-gpus=G."""
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
gpus = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n_utts = int(sys.argv[2]) if len(sys.argv) > 2 else 600
d = tempfile.mkdtemp(prefix="cvx_startup_", dir="/dev/shm" if os.access("/dev/shm", os.W_OK) else None)
model = os.path.join(d, "final.raw")
kf.write_nnet3(model, sy.make_model("v5_cvector"), binary=True)
lengths = sy.make_lengths(n_utts, 1000, 30000, 7)
rng = np.random.default_rng(3)
ark, scp = os.path.join(d, "feats.ark"), os.path.join(d, "feats.scp")
with open(ark, "wb") as f, open(scp, "w") as s:
    for k, T in zip(sy.utt_keys(n_utts), lengths):
        f.write(k.encode() + b" ")
        s.write(f"{k} {ark}:{f.tell()}\n")
        f.write(b"\0BFM \x04" + np.int32(T).tobytes() + b"\x04" + np.int32(23).tobytes())
        f.write(rng.standard_normal((int(T), 23), dtype=np.float32).tobytes())
frames = int(lengths.sum())
exe = os.path.join(cvx.BIN_DIR, "nnet3-xvector-compute")
base = [exe, "--use-gpu=yes", "--verbose=1", "--min-chunk-size=25", "--chunk-size=10000", "--output-node=tdnn6_xvec.affine"]
print("CUDA_VISIBLE_DEVICES =", os.environ.get("CUDA_VISIBLE_DEVICES"), flush=True)


def run_many(label, extra_env):
    env = dict(os.environ, CVX_LOCK_DIR=d, **extra_env)
    env.pop("CVX_DEVICE", None)
    t0 = time.perf_counter()
    ps = [subprocess.Popen(base + [model, f"scp:{scp}", f"ark:{d}/o{g}.ark"], stderr=subprocess.PIPE, text=True, env=env) for g in range(gpus)]
    errs = [p.communicate()[1] for p in ps]
    dt = time.perf_counter() - t0
    assert all(p.returncode == 0 for p in ps), errs
    print(f"== {label}: {gpus} job(s) x {frames} frames at once, {dt:.2f} s wall", flush=True)
    for e in errs:
        keep = [ln.split(") ", 1)[-1] for ln in e.splitlines()
                if "Selected GPU" in ln or "sharing GPU" in ln or "start-up:" in ln or "lane timing" in ln or "since main" in ln]
        print("   " + " | ".join(k[:150] for k in keep), flush=True)


def run_one_process(label):
    """the same G shares from ONE process: --num-gpus=G (reader lanes), CUDA initialised once"""
    env = dict(os.environ, CVX_LOCK_DIR=d)
    env.pop("CVX_DEVICE", None)
    big = os.path.join(d, "all.scp")  # G copies of the table under distinct keys = the G jobs' total work
    with open(big, "w") as f:
        for g in range(gpus):
            for ln in open(scp):
                f.write(f"g{g}_{ln}")
    t0 = time.perf_counter()
    r = subprocess.run(base + [f"--num-gpus={gpus}", model, f"scp:{big}", f"ark:{d}/all.ark"], stderr=subprocess.PIPE, text=True, env=env)
    dt = time.perf_counter() - t0
    assert r.returncode == 0, r.stderr[-2000:]
    print(f"== {label}: 1 process, --num-gpus={gpus}, {gpus} x {frames} frames, {dt:.2f} s wall", flush=True)
    for ln in r.stderr.splitlines():
        if "start-up:" in ln or "since main" in ln or "Time taken" in ln:
            print("   " + ln.split(") ", 1)[-1][:200], flush=True)


mode = sys.argv[3] if len(sys.argv) > 3 else "all"
for rep in range(2):
    if mode == "all":
        run_many(f"selection through CUDA (run {rep})", {"CVX_SELECT_WITH_CUDA": "1"})
    run_many(f"selection before CUDA (run {rep})", {})
    run_one_process(f"one process (run {rep})")
import shutil
shutil.rmtree(d, ignore_errors=True)
