"""End-to-end throughput of the drop-in program (SURVEY 8d's third number: table in -> ark,scp out).
Writes a synthetic data dir (v5 c-vector net, utterances 10-300 s: feats.ark / feats.scp /
utt2num_frames), runs bin/nnet3-xvector-compute on it the way extract_xvectors_new.sh does and
reports audio-seconds per second of wall clock (model load and GPU init included) and of the
program's own processing loop.

  python tools/cli_throughput.py [n_utts=400] [num_gpus=1]

num_gpus > 1 adds: --num-gpus=G on the scp (one reader lane per GPU) and on the archive (one reader,
batches handed round), and G concurrent --shard=J/G jobs (the recipe's fan-out, devices picked
through the lock files); every variant's archive is compared with the one-GPU run."""
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
num_gpus = int(sys.argv[2]) if len(sys.argv) > 2 else 1
base = "/dev/shm" if os.access("/dev/shm", os.W_OK) else None
d = tempfile.mkdtemp(prefix="cvx_cli_", dir=base)
model = os.path.join(d, "final.raw")
kf.write_nnet3(model, sy.make_model("v5_cvector"), binary=True)
lengths = sy.make_lengths(n_utts, 1000, 30000, 7)
rng = np.random.default_rng(3)
ark, scp = os.path.join(d, "feats.ark"), os.path.join(d, "feats.scp")
keys = sy.utt_keys(n_utts)
with open(ark, "wb") as f, open(scp, "w") as s, open(os.path.join(d, "utt2num_frames"), "w") as u:
    for k, T in zip(keys, lengths):
        f.write(k.encode() + b" ")
        s.write(f"{k} {ark}:{f.tell()}\n")
        u.write(f"{k} {int(T)}\n")
        f.write(b"\0BFM \x04" + np.int32(T).tobytes() + b"\x04" + np.int32(23).tobytes())
        f.write(rng.standard_normal((int(T), 23), dtype=np.float32).tobytes())
frames = int(lengths.sum())
exe = os.path.join(cvx.BIN_DIR, "nnet3-xvector-compute")
common = ["--use-gpu=yes", "--verbose=1", "--min-chunk-size=25", "--chunk-size=10000", "--output-node=tdnn6_xvec.affine"]


def report(label, dt, stderr):
    tail = [ln.split(") ", 1)[-1] for ln in stderr.splitlines() if "Time taken" in ln or "Done " in ln]
    loop = [float(t.split("Time taken")[1].split("s:")[0]) for t in tail if "Time taken" in t]
    print(f"{label}: {frames} frames ({os.path.getsize(ark) / 1e9:.2f} GB) in {dt:.2f} s wall -> {frames / 100 / dt:,.0f} audio-s/s"
          + (f"; processing loop {max(loop):.2f} s -> {frames / 100 / max(loop):,.0f} audio-s/s" if loop else ""), flush=True)


def run(label, out, extra, feat):
    t0 = time.perf_counter()
    r = subprocess.run([exe, *common, *extra, model, feat, f"ark,scp:{d}/{out}.ark,{d}/{out}.scp"], capture_output=True, text=True)
    dt = time.perf_counter() - t0
    assert r.returncode == 0, r.stderr[-2000:]
    report(label, dt, r.stderr)
    for ln in r.stderr.splitlines():
        if "lane timing" in ln or "Selected GPU" in ln or "start-up:" in ln or "tear-down" in ln or "cvx trace" in ln:
            print("     " + ln.split(") ", 1)[-1], flush=True)


def same(a, b):
    x, y = dict(kf.read_ark(f"{d}/{a}.ark")), dict(kf.read_ark(f"{d}/{b}.ark"))
    return list(x) == list(y) and all(np.max(np.abs(x[k] - y[k])) <= 5e-4 * np.max(np.abs(y[k])) for k in x)


run("ark file", "x_ark", ["--gpu-id=0"], f"ark:{ark}")
run("scp (same file)", "x_scp", ["--gpu-id=0"], f"scp:{scp}")
run("ark through a pipe (cat)", "x_pipe", ["--gpu-id=0"], f"ark:cat {ark} |")
run("scp, automatic GPU selection", "x_auto", [], f"scp:{scp}")
run("scp, sliding CMN on the GPU (raw mode)", "x_raw", ["--gpu-id=0", "--cmn-window=300"], f"scp:{scp}")
print("scp == ark:", same("x_scp", "x_ark"), flush=True)
if num_gpus > 1:
    run(f"scp, --num-gpus={num_gpus} (one reader lane per GPU)", "x_lanes", [f"--num-gpus={num_gpus}"], f"scp:{scp}")
    run(f"ark file, --num-gpus={num_gpus} (one reader, batches handed round)", "x_rr", [f"--num-gpus={num_gpus}"], f"ark:{ark}")
    run(f"scp raw mode, --num-gpus={num_gpus}", "x_raw_lanes", [f"--num-gpus={num_gpus}", "--cmn-window=300"], f"scp:{scp}")
    print("lanes == 1 GPU:", same("x_lanes", "x_ark"), "| round-robin == 1 GPU:", same("x_rr", "x_ark"),
          "| raw lanes == raw 1 GPU:", same("x_raw_lanes", "x_raw"), flush=True)
    # the recipe's fan-out: G concurrent jobs, each --shard=J/G of the unsplit table, devices picked through lock files
    env = dict(os.environ, CVX_LOCK_DIR=d)
    env.pop("CVX_DEVICE", None)
    t0 = time.perf_counter()
    ps = [subprocess.Popen([exe, *common, f"--shard={j}/{num_gpus}", f"--utt2num-frames={d}/utt2num_frames", model, f"scp:{scp}",
                            f"ark,scp:{d}/job{j}.ark,{d}/job{j}.scp"], stderr=subprocess.PIPE, text=True, env=env)
          for j in range(1, num_gpus + 1)]
    errs = [p.communicate()[1] for p in ps]
    dt = time.perf_counter() - t0
    assert all(p.returncode == 0 for p in ps), errs
    picked = [ln.split(") ", 1)[-1] for e in errs for ln in e.splitlines() if "Selected GPU" in ln or "sharing GPU" in ln]
    report(f"{num_gpus} concurrent --shard=J/{num_gpus} jobs (slowest job)", dt, "\n".join(errs))
    print("   device selection:", picked, flush=True)
    merged = {}
    for j in range(1, num_gpus + 1):
        merged.update(dict(kf.read_ark(f"{d}/job{j}.ark")))
    one = dict(kf.read_ark(f"{d}/x_ark.ark"))
    print("   union of the jobs' outputs == 1-GPU run:", set(merged) == set(one) and
          all(np.max(np.abs(merged[k] - one[k])) <= 5e-4 * np.max(np.abs(one[k])) for k in one), flush=True)
import shutil
shutil.rmtree(d, ignore_errors=True)
