#!/usr/bin/env python
"""bench.py -- c-vector extraction throughput (BASELINE.json metric) on N B200s.

  python bench.py --gpus 1 --steps 5 --warmup 3
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port P bench.py --gpus N --steps K --warmup W
  python bench.py --impl reference ...      # the restated CPU path on the host cores

Workload (config.workload): BASELINE.json configs[3] -- the v5 final c-vector net
(phonetic bottleneck TDNN + speaker TDNN, random init), 10 000 synthetic 23-dim
utterances of 10-300 s, --chunk-size 10000, length-balanced utterance shards over 8
GPUs: 1 250 utterances (~19.4 M frames, 1.8 GB of fp32 features) per GPU.  Scaling is
weak: each rank owns one LPT shard of N x 1250 utterances; N=8 is the full config.
A step = one pass of the hot path (cvx_extract through the C ABI) over the rank's shard.

  value    audio-seconds (frames/100) processed per second by all ranks, features
           already resident in HBM, timed with CUDA events on the engine's stream
           (max over ranks of the summed step times)
  e2e      the same pass from pinned HOST buffers (H2D of the features and D2H of the
           embeddings inside the timed region), wall clock around the synchronous call
  e2e_cli  the same shard through the drop-in PROGRAM (bin/nnet3-xvector-compute, ark file in ->
           ark,scp out, one process per GPU like the recipe's nj jobs), wall clock
  parity   the timed output of k utterances of this shard against the fp64 oracle (outside the
           timed region); the run FAILS if max rel err > 1e-3 or cosine < 0.9999
  other_workloads   the other BASELINE configs, device-resident leg only, a few seconds each
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
PKG = "speaker-embedding-with-phonetic-information_b200"
METRIC = "c-vector extraction audio-sec/sec"
UNIT = "audio-sec/sec"
# name -> (topology, utterances per GPU, min frames, max frames, chunk-size, length seed, description)
WORKLOADS = {
    "cfg4": ("v5_cvector", 1250, 1000, 30000, 10000, 1234 + 4,
             "BASELINE configs[3]: egs/sre/v5 final c-vector at SRE16 scale, utterance-sharded "
             "(1250 utts of 10-300 s per GPU; 8 GPUs = the 10k-utt config)"),
    "cfg1": ("v2_xvector", 200, 300, 3000, 10000, 1234 + 1, "BASELINE configs[0]: v2 x-vector, 200 utts of 3-30 s"),
    "cfg2": ("v3_multitask", 1000, 300, 3000, 10000, 1234 + 2, "BASELINE configs[1]: v3 multitask c-vector, 1k utts of 3-30 s"),
    "cfg3": ("v4_two_input", 1000, 300, 3000, 10000, 1234 + 3,
             "BASELINE configs[2]: v4 phonetic-adaptation c-vector, 23-dim + 40-dim hires streams, 1k utts"),
    "cfg5x": ("v2_xvector", 125000, 200, 200, -1, 1234 + 5,
              "BASELINE configs[4]: short-segment sweep, 2 s segments, x-vector net (125k segments per GPU; 8 GPUs = 1M)"),
    "cfg5c": ("v5_cvector", 125000, 200, 200, -1, 1234 + 5,
              "BASELINE configs[4]: short-segment sweep, 2 s segments, c-vector net (125k segments per GPU; 8 GPUs = 1M)"),
}
MIN_CHUNK_SIZE = 25
TOL_REL, TOL_COS = 1e-3, 0.9999          # BASELINE.json north_star tolerance
CLI_REPLICAS = 4                         # e2e_cli: the shard's table is listed this many times


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"tflops": d["bf16_tflops_sustained"], "tflops_burst": d["bf16_tflops"], "hbm": d["hbm_gbs"],
                "source": "measured (MEASURED_PEAKS.json; sustained bf16 GEMM, STREAM copy)"}
    return {"tflops": 1400.0, "tflops_burst": 1590.0, "hbm": 6650.0,
            "source": "fallback (B200_PROFILING.md: 1.59 PF burst / ~1.4 PF sustained, 6.65 TB/s)"}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self._stop_evt.is_set():
            try:
                o = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                    "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in o.strip().split(",")]
                if len(f) >= 7:
                    self.rows.append(f)
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self) -> dict:
        self._stop_evt.set()
        self.join(timeout=6)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_min_mhz": sm[0], "sm_max_mhz": float(self.rows[0][1]),
                "power_w_max": max(float(r[2]) for r in self.rows), "samples": len(self.rows), "reasons": reasons}


def run_reference(args, rank: int) -> None:
    """--impl reference: the reference's CPU path (restated: Kaldi is not shipped and
    cannot be built offline) with one single-threaded job per host core."""
    if rank != 0:
        return
    from oracle import cpu_baseline
    topology, _, min_frames, max_frames, chunk_size, _, desc = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    secs = 8.0
    tot_frames, tot_wall, per_step = 0, 0.0, []
    for i in range(args.warmup + args.steps):
        r = cpu_baseline.run(topology, cores, secs, min_frames, max_frames, chunk_size, MIN_CHUNK_SIZE)
        if i >= args.warmup:
            tot_frames += r["frames"]
            tot_wall += r["wall_s"]
            per_step.append(r["wall_s"] * 1e3)
    value = tot_frames / 100.0 / tot_wall
    sample = (f"each step: {cores} single-threaded jobs x ~{secs:.0f} s of v5 c-vector extraction "
              f"(utterances U[{min_frames},{max_frames}] frames, chunk-size {chunk_size}); numpy/BLAS fp32 oracle")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": float(np.mean(per_step)), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc + " (bounded sample)", "workload_id": args.workload,
                   "topology": topology, "chunk_size": chunk_size, "min_chunk_size": MIN_CHUNK_SIZE},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}), flush=True)


class Workload:
    """One rank's LPT shard of a BASELINE config: model, utterance lengths, chunk tables."""

    def __init__(self, wid: str, world: int, rank: int, utts_per_gpu: int, mods):
        cvx, kf, sy, sh = mods
        (self.topology, default_utts, self.min_frames, self.max_frames, self.chunk_size, self.seed,
         self.desc) = WORKLOADS[wid]
        self.wid = wid
        self.utts_per_gpu = utts_per_gpu if utts_per_gpu > 0 else default_utts
        lengths_all = sy.make_lengths(world * self.utts_per_gpu, self.min_frames, self.max_frames, self.seed)
        self.lengths = lengths_all[sh.lpt_shards(lengths_all, world)[rank]]
        self.total_rows = int(self.lengths.sum())
        self.n_utts = len(self.lengths)
        self.node = sy.EMBEDDING_NODE[self.topology]
        self.py_model = sy.make_model(self.topology, seed=42)      # the oracle reads this same object
        self.model_dir = tempfile.mkdtemp(prefix=f"cvx_bench_{wid}_")
        self.model_path = os.path.join(self.model_dir, "final.raw")
        kf.write_nnet3(self.model_path, self.py_model, binary=True)
        self.model = cvx.Model(self.model_path, self.node)
        starts, lens, utts = [], [], []
        base = 0
        for u, T in enumerate(self.lengths.tolist()):
            for s, n in cvx.chunk_plan(T, self.chunk_size, MIN_CHUNK_SIZE, False):
                starts.append(base + s); lens.append(n); utts.append(u)
            base += T
        self.starts, self.lens, self.utts = np.array(starts, np.int64), np.array(lens, np.int32), np.array(utts, np.int32)
        self.audio_sec = self.total_rows / 100.0
        self.utt_row0 = np.concatenate([[0], np.cumsum(self.lengths)[:-1]]).astype(np.int64)

    def cleanup(self):
        shutil.rmtree(self.model_dir, ignore_errors=True)


def parity_check(wl: Workload, host_feats, out: np.ndarray, k: int, rank: int) -> dict:
    """k utterances of this shard (the longest, the shortest, k-2 random ones) through the fp64 oracle,
    compared with the embeddings the timed steps produced."""
    from oracle.nnet3_oracle import Oracle, extract_utt, parity_metrics
    t0 = time.perf_counter()
    orc = Oracle(wl.py_model, wl.node, np.float64)
    names = wl.model.input_names
    order = np.argsort(wl.lengths, kind="stable")
    pick = [int(order[-1]), int(order[0])][:k]
    rng = np.random.default_rng(777 + rank)
    for u in rng.permutation(wl.n_utts).tolist():
        if len(pick) >= k:
            break
        if u not in pick:
            pick.append(int(u))
    ref = []
    for u in pick:
        r0, T = int(wl.utt_row0[u]), int(wl.lengths[u])
        e = extract_utt(orc, {n: a[r0:r0 + T] for n, a in zip(names, host_feats)}, chunk_size=wl.chunk_size,
                        min_chunk_size=MIN_CHUNK_SIZE)
        ref.append(e)
    rel, cos = parity_metrics(out[pick], np.stack(ref))
    return {"max_rel": rel, "min_cos": cos, "n": len(pick), "frames": int(wl.lengths[pick].sum()),
            "oracle": "oracle/nnet3_oracle.py fp64", "seconds": round(time.perf_counter() - t0, 1)}


def cli_leg(wl: Workload, host_feats, out_ref: np.ndarray, local_rank: int, mods, barrier) -> dict:
    """The shard through the drop-in program: ark file in (page cache) -> ark,scp out."""
    cvx, kf, sy, sh = mods
    base = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None
    d = tempfile.mkdtemp(prefix="cvx_cli_", dir=base)
    try:
        ark, scp = os.path.join(d, "feats.ark"), os.path.join(d, "feats.scp")
        keys = sy.utt_keys(wl.n_utts)
        x = host_feats[0]
        # the table lists every utterance CLI_REPLICAS times (different keys, same archive offsets): the job is
        # long enough (~1 s of device time) for the process start-up not to be all there is to measure
        lines = [[] for _ in range(CLI_REPLICAS)]
        with open(ark, "wb") as f:
            for k, r0, T in zip(keys, wl.utt_row0.tolist(), wl.lengths.tolist()):
                f.write(k.encode() + b" ")
                for r in range(CLI_REPLICAS):
                    lines[r].append(f"{k}_r{r} {ark}:{f.tell()}\n")
                f.write(b"\0BFM \x04" + np.int32(T).tobytes() + b"\x04" + np.int32(x.shape[1]).tobytes())
                f.write(x[r0:r0 + T].tobytes())
        with open(scp, "w") as sf:
            for r in range(CLI_REPLICAS):
                sf.writelines(lines[r])
        exe = os.path.join(cvx.BIN_DIR, "nnet3-xvector-compute")
        # no device hint, like the recipe's GPU branch (extract_xvectors_new.sh:83-89): each of the N processes picks
        # a free GPU through the lock files and makes it the only device its CUDA context sees
        # (a CUDA_VISIBLE_DEVICES restriction of this process is inherited and respected)
        env = {k: v for k, v in os.environ.items() if k != "CVX_DEVICE"}
        cmd = [exe, "--use-gpu=yes", "--verbose=1", f"--min-chunk-size={MIN_CHUNK_SIZE}",
               f"--chunk-size={wl.chunk_size}", f"--output-node={wl.node}", wl.model_path, f"scp:{scp}",
               f"ark,scp:{d}/xvector.ark,{d}/xvector.scp"]
        barrier()
        t0 = time.perf_counter()
        r = subprocess.run(cmd, capture_output=True, text=True, env=env)
        wall = time.perf_counter() - t0
        if r.returncode != 0:
            raise RuntimeError("drop-in program failed: " + r.stderr[-2000:])
        in_loop, picked, phases = None, None, []
        for ln in r.stderr.splitlines():
            if "Time taken" in ln:
                in_loop = float(ln.split("Time taken")[1].split("s:")[0])
            if "Selected GPU" in ln or "sharing GPU" in ln:
                picked = ln.split(") ", 1)[-1]
            if "start-up:" in ln or "tear-down" in ln or "lane timing" in ln:
                phases.append(ln.split(") ", 1)[-1])
        got = np.stack([v for _, v in kf.read_ark(f"{d}/xvector.ark")])
        same = bool(got.shape == (CLI_REPLICAS * out_ref.shape[0], out_ref.shape[1]) and
                    all(float(np.max(np.abs(got[r * len(out_ref):(r + 1) * len(out_ref)] - out_ref))) <= 5e-4 * float(np.max(np.abs(out_ref)))
                        for r in range(CLI_REPLICAS)))   # other device batches: other summation order / activation form
        return {"wall_s": wall, "in_loop_s": in_loop, "matches_c_abi_output": same, "ark_bytes": os.path.getsize(ark),
                "device": picked, "phases": phases}
    finally:
        shutil.rmtree(d, ignore_errors=True)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg4", choices=sorted(WORKLOADS),
                    help="cfg4 = the headline workload; the others are also timed briefly as other_workloads")
    ap.add_argument("--utts-per-gpu", type=int, default=0)
    ap.add_argument("--precision", default="fp16", choices=["fp16", "bf16"])
    ap.add_argument("--max-batch-rows", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true",
                    help="timed legs only: no parity / CLI / other-workload / pooling / CPU legs (for ncu runs)")
    ap.add_argument("--parity-utts", type=int, default=-1, help="utterances checked against the oracle (default 8; 2 per rank for N>1)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("for --gpus N > 1 launch with torch.distributed.run --nproc-per-node N (one rank per GPU)")
        raise SystemExit(f"--gpus {args.gpus} does not match WORLD_SIZE {world}")

    import torch
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce(v: float, op: str) -> float:
        if dist is None:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op={"max": dist.ReduceOp.MAX, "min": dist.ReduceOp.MIN, "sum": dist.ReduceOp.SUM}[op])
        return float(t.item())

    def gather(obj):
        if dist is None:
            return [obj]
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out

    cvx = importlib.import_module(PKG)
    mods = (cvx, importlib.import_module(PKG + ".kaldi_format"), importlib.import_module(PKG + ".synth"),
            importlib.import_module(PKG + ".shard"))
    lib = cvx.load_library()
    import ctypes as C

    # ---- workload: this rank's LPT shard of N x 1250 utterances -------------------
    wl = Workload(args.workload, world, rank, args.utts_per_gpu, mods)
    model, info, total_rows, n_utts = wl.model, wl.model.info, wl.total_rows, wl.n_utts
    pinned = [cvx.PinnedArray((total_rows, d)) for d in model.input_dims]
    rng = np.random.default_rng(wl.seed * 1000 + rank)
    step_rows = 1 << 20
    for pa, d in zip(pinned, model.input_dims):
        for r0 in range(0, total_rows, step_rows):
            r1 = min(total_rows, r0 + step_rows)
            pa.array[r0:r1] = rng.standard_normal((r1 - r0, d), dtype=np.float32)
    host_feats = [pa.array for pa in pinned]

    ex = cvx.Extractor(model, device=local_rank, precision=args.precision, max_batch_rows=args.max_batch_rows,
                       profile=True)
    dptrs, nbytes = [], 0
    for pa, d in zip(pinned, model.input_dims):
        dp = C.c_void_p()
        nb = total_rows * d * 4
        assert lib.cvx_device_alloc(local_rank, C.byref(dp), nb) == 0, lib.cvx_last_error()
        assert lib.cvx_memcpy_h2d(local_rank, dp, pa.array.ctypes.data_as(C.c_void_p), nb) == 0
        dptrs.append(dp.value)
        nbytes += nb
    out_pinned = cvx.PinnedArray((n_utts, info.embedding_dim))     # results come back with an asynchronous copy
    out = out_pinned.array
    out[:] = 0

    def step_device():
        ex.extract_chunks(dptrs, total_rows, wl.starts, wl.lens, wl.utts, None, n_utts, out=out, flags=cvx.FLAG_DEVICE_INPUT)
        return ex.last_extract_ms()

    def step_host():
        t0 = time.perf_counter()
        ex.extract_chunks(host_feats, total_rows, wl.starts, wl.lens, wl.utts, None, n_utts, out=out)
        return (time.perf_counter() - t0) * 1e3

    # ---- device-resident leg -------------------------------------------------------
    for _ in range(args.warmup):
        step_device()
    ex.reset_profile()
    launches0 = ex.launch_count()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    t_wall0 = time.perf_counter()
    dev_ms = [step_device() for _ in range(args.steps)]
    barrier()
    wall_ms = (time.perf_counter() - t_wall0) * 1e3
    launches = ex.launch_count() - launches0
    prof = {p["name"]: p for p in ex.profile()}
    emb_dev = out.copy()
    total_ms = reduce(float(sum(dev_ms)), "max")
    total_audio = reduce(wl.audio_sec, "sum") * args.steps
    value = total_audio / (total_ms / 1e3)

    # ---- end-to-end leg (host buffers) ----------------------------------------------
    for _ in range(2):
        step_host()
    barrier()
    e2e_ms = [step_host() for _ in range(args.steps)]
    barrier()
    clocks = sampler.stop()                             # sampled over both timed regions
    assert np.array_equal(out, emb_dev), "host-input and device-input passes disagree"
    e2e_total_ms = reduce(float(sum(e2e_ms)), "max")
    e2e_value = total_audio / (e2e_total_ms / 1e3)
    total_launches = int(reduce(float(launches), "sum"))
    finite = bool(np.isfinite(out).all() and np.abs(out).sum() > 0)
    per_rank = gather({"rank": rank, "frames": total_rows, "utts": n_utts, "ms_per_step": float(np.mean(dev_ms)),
                       "e2e_ms_per_step": float(np.mean(e2e_ms)), "sm_mhz_median": clocks.get("sm_mhz"),
                       "sm_mhz_min": clocks.get("sm_min_mhz"), "power_w_max": clocks.get("power_w_max"),
                       "reasons": clocks.get("reasons"),
                       "gemm_ms_per_step": prof["tdnn_gemm"]["total_ms"] / args.steps})

    # ---- parity of the timed output against the oracle (outside the timed region) ----
    parity = None
    if not args.quick:
        k = args.parity_utts if args.parity_utts >= 0 else (8 if world == 1 else 2)
        if k > 0:
            pr = parity_check(wl, host_feats, emb_dev, k, rank)
            parity = {"max_rel": reduce(pr["max_rel"], "max"), "min_cos": reduce(pr["min_cos"], "min"),
                      "n": int(reduce(float(pr["n"]), "sum")), "frames": int(reduce(float(pr["frames"]), "sum")),
                      "tolerance": {"max_rel": TOL_REL, "min_cos": TOL_COS}, "oracle": pr["oracle"],
                      "what": "embeddings produced by the timed device-resident steps vs the oracle on the same features "
                              "(longest + shortest + random utterances of every rank's shard)",
                      "seconds": pr["seconds"]}
            parity["ok"] = bool(parity["max_rel"] <= TOL_REL and parity["min_cos"] >= TOL_COS and finite)

    # ---- the drop-in program on the same shard (one process per GPU, like the recipe's nj jobs) ----
    e2e_cli = None
    if not args.quick and len(host_feats) == 1:
        c = cli_leg(wl, host_feats, emb_dev, local_rank, mods, barrier)
        cli_wall = reduce(c["wall_s"], "max")
        cli_loop = reduce(c["in_loop_s"] or 0.0, "max")
        all_same = reduce(1.0 if c["matches_c_abi_output"] else 0.0, "min") == 1.0
        cli_audio = total_audio / args.steps * CLI_REPLICAS
        e2e_cli = {"value": cli_audio / cli_wall, "unit": UNIT, "wall_s": cli_wall,
                   "in_loop_value": (cli_audio / cli_loop) if cli_loop > 0 else None, "in_loop_s": cli_loop,
                   "utts_per_gpu": wl.n_utts * CLI_REPLICAS, "frames_per_gpu": total_rows * CLI_REPLICAS,
                   "startup_s": cli_wall - cli_loop,
                   "what": "bin/nnet3-xvector-compute --use-gpu=yes (no device hint: lock-file selection) --chunk-size=.. model scp:<table over an archive in page cache> "
                           "ark,scp:..; one process per GPU on its shard listed 4 times (the recipe's nj jobs), wall clock of the slowest process "
                           "incl. CUDA context + model load + engine build (startup_s); in_loop = the program's own 'Time taken' line",
                   "ark_bytes_per_gpu": c["ark_bytes"], "matches_c_abi_output": all_same,
                   "device_selection": [g["device"] for g in gather({"device": c["device"]})],
                   "rank0_phases": c["phases"]}

    # ---- the other BASELINE configs, device-resident leg only ----------------------------
    other = None
    if not args.quick and world == 1 and args.workload == "cfg4":
        other = {}
        for wid in ("cfg1", "cfg2", "cfg3", "cfg5x", "cfg5c"):
            w2 = Workload(wid, 1, 0, 0, mods)
            g = torch.Generator(device="cuda").manual_seed(w2.seed)
            dev = [torch.randn((w2.total_rows, d), generator=g, device="cuda", dtype=torch.float32) for d in w2.model.input_dims]
            o2_pinned = cvx.PinnedArray((w2.n_utts, w2.model.info.embedding_dim))
            o2 = o2_pinned.array
            o2[:] = 0
            # timed without the per-launch profiling events (they cost ~4 % on the short-segment configs: ~400
            # launches per step); the kernel breakdown comes from a second, profiled engine
            res = {}
            for prof_on in (False, True):
                ex2 = cvx.Extractor(w2.model, device=local_rank, precision=args.precision, profile=prof_on)
                ms = []
                for i in range(3 + 3):
                    if i == 3:
                        ex2.reset_profile()
                    ex2.extract_chunks([t.data_ptr() for t in dev], w2.total_rows, w2.starts, w2.lens, w2.utts, None, w2.n_utts,
                                       out=o2, flags=cvx.FLAG_DEVICE_INPUT)
                    if i >= 3:
                        ms.append(ex2.last_extract_ms())
                res[prof_on] = (float(np.mean(ms)), {p["name"]: p for p in ex2.profile()} if prof_on else None)
                if not prof_on:
                    ex2.close()
            ms_plain, (ms_prof, p2) = res[False][0], res[True]
            gm = p2["tdnn_gemm"]
            other[wid] = {"value": w2.audio_sec / (ms_plain / 1e3), "ms_per_step": ms_plain, "ms_per_step_profiled": ms_prof,
                          "gemm_ms_per_step": gm["total_ms"] / 3, "gemm_frac": gm["flops"] / (gm["total_ms"] * 1e-3) / 1e12 / peaks()["tflops"],
                          # of the SUSTAINED measured peak: millisecond steps run at burst clocks and can exceed 1
                          "gemm_frac_of": "sustained bf16 GEMM peak (MEASURED_PEAKS.json)",
                          "topology": w2.topology, "frames": w2.total_rows, "utts": w2.n_utts,
                          "outputs_finite": bool(np.isfinite(o2).all() and np.abs(o2).sum() > 0)}
            ex2.close()
            del dev, o2
            o2_pinned.close()
            w2.model.close()
            w2.cleanup()
            torch.cuda.empty_cache()

    if rank == 0:
        pk = peaks()
        g = prof["tdnn_gemm"]
        gemm_tflops = g["flops"] / (g["total_ms"] * 1e-3) / 1e12
        roofline_pool = None
        if not args.quick:
            # default path: pooling is fused into the pooled layer's GEMM epilogue.  The stand-alone
            # HBM-bound kernel (CVX_FUSED_POOL=0) is measured in a separate, untimed pass.
            os.environ["CVX_FUSED_POOL"] = "0"
            ex2 = cvx.Extractor(model, device=local_rank, precision=args.precision, max_batch_rows=args.max_batch_rows,
                                profile=True)
            for i in range(3):
                if i == 1:
                    ex2.reset_profile()
                ex2.extract_chunks(dptrs, total_rows, wl.starts, wl.lens, wl.utts, None, n_utts, out=np.zeros_like(out),
                                   flags=cvx.FLAG_DEVICE_INPUT)
            pl = {p["name"]: p for p in ex2.profile()}["stats_pool_partial"]
            ex2.close()
            del os.environ["CVX_FUSED_POOL"]
            pool_gbs = pl["bytes"] / (pl["total_ms"] * 1e-3) / 1e9
            roofline_pool = {"kernel": "stats_pool_partial", "bound": "hbm", "achieved": pool_gbs, "peak": pk["hbm"],
                             "unit": "GB/s", "frac": pool_gbs / pk["hbm"], "traffic": None, "launches": int(pl["launches"]),
                             "note": "pooling is fused into the pooled layer's GEMM epilogue on the timed path; this is the "
                                     "stand-alone kernel (CVX_FUSED_POOL=0) measured in a separate pass of the same workload"}
        # ncu-measured companions of the roofline (profiles/, committed): per-layer DRAM bytes vs algorithmic
        # bytes and the tensor-pipe utilisation, the second half of BASELINE.json's metric
        traffic, tensor_pipe = None, None
        tp = os.path.join(ROOT, "profiles", "tensor_pipe.json")
        if os.path.exists(tp):
            t = json.load(open(tp))
            traffic = t.get("dram_bytes_per_launch_avg")
            tensor_pipe = {"flop_weighted_pct": t.get("tensor_pipe_pct_flop_weighted"),
                           "time_weighted_pct": t.get("tensor_pipe_pct_time_weighted"),
                           "dram_over_algorithmic": t.get("dram_over_algorithmic"),
                           "per_layer": t.get("per_layer"), "source": t.get("source")}
        kernel_ms = {k: round(v["total_ms"] / args.steps, 3) for k, v in prof.items() if not k.startswith("layer:")}
        # per frame layer inside the same timed steps: ms per step and achieved TFLOP/s (sustained clocks)
        layer_ms = {k[6:]: {"ms": round(v["total_ms"] / args.steps, 3),
                            "tflops": round(v["flops"] / max(v["total_ms"], 1e-9) * 1e-9, 1)}
                    for k, v in prof.items() if k.startswith("layer:")}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.precision, "data": "synthetic",
            "config": {"workload": wl.desc, "workload_id": args.workload,
                       "topology": wl.topology, "utts_per_gpu": wl.utts_per_gpu, "frames_per_gpu": total_rows,
                       "chunk_size": wl.chunk_size, "min_chunk_size": MIN_CHUNK_SIZE, "parallelism": f"utterance-shards x{world} (LPT, no collective)",
                       "operands": f"{args.precision} operands, fp32 accumulate; pooling+tdnn6 fp32",
                       "l2": f"inputs larger than L2 ({nbytes / 1e9:.2f} GB features, multi-GB activations per step)",
                       "mac_per_frame": int(info.mac_per_frame), "outputs_finite": finite},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(nbytes), "d2h_bytes_per_step": int(out.nbytes),
                    "ms_per_step": e2e_total_ms / args.steps, "timer": "wall clock around the synchronous C-ABI call, max over ranks"},
            "e2e_cli": e2e_cli,
            "parity": parity,
            "gpu_launches": total_launches,
            "wall_ms_per_step": wall_ms / args.steps,
            "kernel_ms_per_step": kernel_ms,
            "layer_ms_per_step": layer_ms,
            "roofline": {"kernel": "tdnn_gemm (tcgen05 implicit GEMM, all frame layers)", "bound": "tensor",
                         "achieved": gemm_tflops, "peak": pk["tflops"], "unit": "TFLOP/s", "frac": gemm_tflops / pk["tflops"],
                         "traffic": traffic, "tensor_pipe_pct": tensor_pipe["flop_weighted_pct"] if tensor_pipe else None,
                         "tensor_pipe": tensor_pipe, "peak_source": pk["source"],
                         "launches": int(g["launches"]), "avg_launch_ms": g["total_ms"] / max(g["launches"], 1)},
            "roofline_pool": roofline_pool,
            "per_rank": per_rank,
            "other_workloads": other,
        }
        if world == 1 and not args.no_cpu_baseline and not args.quick:
            from oracle import cpu_baseline
            cores = os.cpu_count() or 1
            r = cpu_baseline.run(wl.topology, cores, 12.0, wl.min_frames, wl.max_frames, wl.chunk_size, MIN_CHUNK_SIZE)
            line["cpu_baseline"] = {"value": r["audio_sec_per_sec"], "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{r['utts']} utterances / {r['frames']} frames of the same workload, one "
                                              f"single-threaded fp32 job per core ({cores}), {r['wall_s']:.1f} s wall"}
        else:
            line["cpu_baseline"] = None
        if parity is not None and not parity["ok"]:
            sys.stderr.write("bench.py: PARITY FAILURE, no number reported: " + json.dumps(parity) + "\n")
        else:
            print(json.dumps(line), flush=True)
    wl.cleanup()
    if parity is not None and not parity["ok"]:          # every rank holds the same reduced verdict
        if dist is not None:
            dist.destroy_process_group()
        raise SystemExit(1)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
