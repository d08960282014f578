"""CPU baseline runner -- TEST/BENCH INFRASTRUCTURE, NOT PRODUCT CODE.

Times the oracle (oracle/nnet3_oracle.py, float32 = Kaldi's BaseFloat) the way the
reference runs extraction on CPUs: `--use-gpu=no`, one single-threaded job per
shard, `nj` jobs side by side (egs/sre/v5/run_sre10.sh:25,82;
sid/nnet3/xvector/extract_xvectors_new.sh:91-93; "nnet3 does not yet support
multiple threads", sid/nnet3_cvector/cvector/extract_am_embedding.sh:79).
Kaldi itself cannot be built here (not shipped, no network), so this is the
*restated* CPU path (kind "port"): spliced-input copy + sgemm per layer + ReLU +
BatchNorm passes + pooling + tdnn6, numpy/BLAS pinned to one thread per process.

Only bench.py (cpu_baseline leg and --impl reference) and tests/ use this.

  python -m oracle.cpu_baseline --topology v5_cvector --procs 8 --target-seconds 15
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_THREAD_ENV = {"OMP_NUM_THREADS": "1", "MKL_NUM_THREADS": "1", "OPENBLAS_NUM_THREADS": "1",
               "NUMEXPR_NUM_THREADS": "1", "VECLIB_MAXIMUM_THREADS": "1"}


def _worker(args) -> None:
    import importlib
    import numpy as np
    sys.path.insert(0, ROOT)
    sy = importlib.import_module("speaker-embedding-with-phonetic-information_b200.synth")
    from oracle.nnet3_oracle import Oracle, extract_all
    model = sy.make_model(args.topology, seed=42)
    node = sy.EMBEDDING_NODE[args.topology]
    o = Oracle(model, node, np.float32)
    names = o.g.input_names
    dims = [o.g.nodes[n].dim for n in names]
    rng = np.random.default_rng(1000 + args.rank)
    lengths = rng.integers(args.min_frames, args.max_frames + 1, size=4096).tolist()
    # calibration: one short chunk, then size the sample for ~target seconds
    cal = {n: sy.make_features([1000], d, 7) for n, d in zip(names, dims)}
    t0 = time.perf_counter()
    extract_all(o, cal, [1000], chunk_size=args.chunk_size, min_chunk_size=args.min_chunk_size)
    per_frame = (time.perf_counter() - t0) / 1000.0
    budget = max(int(args.target_seconds / per_frame), 2000)
    sel, tot = [], 0
    for n in lengths:
        if tot + n > budget and sel:
            break
        n = min(n, budget)
        sel.append(n)
        tot += n
    feats = {n: sy.make_features(sel, d, 99 + args.rank + i) for i, (n, d) in enumerate(zip(names, dims))}
    print(json.dumps({"ready": True}), flush=True)
    sys.stdin.readline()                      # parent's "go": all workers start together
    t_start = time.time()
    emb, ok = extract_all(o, feats, sel, chunk_size=args.chunk_size, min_chunk_size=args.min_chunk_size)
    t_end = time.time()
    print(json.dumps({"frames": int(sum(sel)), "utts": len(sel), "t_start": t_start, "t_end": t_end,
                      "checksum": float(np.abs(emb).sum())}), flush=True)


def run(topology: str, procs: int, target_seconds: float, min_frames: int, max_frames: int,
        chunk_size: int = 10000, min_chunk_size: int = 25) -> dict:
    """Runs `procs` single-threaded workers concurrently; returns aggregate
    audio-seconds per second (frames / 100 / wall)."""
    env = dict(os.environ)
    env.update(_THREAD_ENV)
    cmd = [sys.executable, "-m", "oracle.cpu_baseline", "--worker", "--topology", topology,
           "--target-seconds", str(target_seconds), "--min-frames", str(min_frames), "--max-frames", str(max_frames),
           "--chunk-size", str(chunk_size), "--min-chunk-size", str(min_chunk_size)]
    ps = [subprocess.Popen(cmd + ["--rank", str(r)], cwd=ROOT, env=env, stdin=subprocess.PIPE, stdout=subprocess.PIPE,
                           text=True) for r in range(procs)]
    for p in ps:
        line = p.stdout.readline()
        if not line or not json.loads(line).get("ready"):
            raise RuntimeError("cpu baseline worker failed to start")
    for p in ps:
        p.stdin.write("go\n")
        p.stdin.flush()
    res = []
    for p in ps:
        line = p.stdout.readline()
        p.wait()
        if p.returncode != 0 or not line:
            raise RuntimeError("cpu baseline worker failed")
        res.append(json.loads(line))
    frames = sum(r["frames"] for r in res)
    wall = max(r["t_end"] for r in res) - min(r["t_start"] for r in res)
    return {"audio_sec_per_sec": frames / 100.0 / wall, "frames": frames, "utts": sum(r["utts"] for r in res),
            "wall_s": wall, "procs": procs}


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--worker", action="store_true")
    ap.add_argument("--rank", type=int, default=0)
    ap.add_argument("--topology", default="v5_cvector")
    ap.add_argument("--procs", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--target-seconds", type=float, default=15.0)
    ap.add_argument("--min-frames", type=int, default=1000)
    ap.add_argument("--max-frames", type=int, default=30000)
    ap.add_argument("--chunk-size", type=int, default=10000)
    ap.add_argument("--min-chunk-size", type=int, default=25)
    a = ap.parse_args()
    if a.worker:
        _worker(a)
    else:
        print(json.dumps(run(a.topology, a.procs, a.target_seconds, a.min_frames, a.max_frames, a.chunk_size,
                             a.min_chunk_size)))


if __name__ == "__main__":
    main()
