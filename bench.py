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

  value  audio-seconds (frames/100) processed per second by all ranks, features
         already resident in HBM, timed with CUDA events on the engine's stream
         (max over ranks of the summed step times)
  e2e    the same pass from pinned HOST buffers (H2D of the features and D2H of the
         embeddings inside the timed region), wall clock around the synchronous call
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
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
TOPOLOGY, UTTS_PER_GPU, MIN_FRAMES, MAX_FRAMES, CHUNK_SIZE, LENGTH_SEED, WORKLOAD_DESC = WORKLOADS["cfg4"]


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
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]),
                "power_w_max": max(float(r[2]) for r in self.rows), "samples": len(self.rows), "reasons": reasons}


def run_reference(args, rank: int) -> None:
    """--impl reference: the reference's CPU path (restated: Kaldi is not shipped and
    cannot be built offline) with one single-threaded job per host core."""
    if rank != 0:
        return
    from oracle import cpu_baseline
    cores = os.cpu_count() or 1
    secs = 8.0
    tot_frames, tot_wall, per_step = 0, 0.0, []
    for i in range(args.warmup + args.steps):
        r = cpu_baseline.run(TOPOLOGY, cores, secs, MIN_FRAMES, MAX_FRAMES, CHUNK_SIZE, MIN_CHUNK_SIZE)
        if i >= args.warmup:
            tot_frames += r["frames"]
            tot_wall += r["wall_s"]
            per_step.append(r["wall_s"] * 1e3)
    value = tot_frames / 100.0 / tot_wall
    sample = (f"each step: {cores} single-threaded jobs x ~{secs:.0f} s of v5 c-vector extraction "
              f"(utterances U[{MIN_FRAMES},{MAX_FRAMES}] frames, chunk-size {CHUNK_SIZE}); numpy/BLAS fp32 oracle")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": float(np.mean(per_step)), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD_DESC + " (bounded sample)", "workload_id": args.workload,
                   "topology": TOPOLOGY, "chunk_size": CHUNK_SIZE, "min_chunk_size": MIN_CHUNK_SIZE},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}), flush=True)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg4", choices=sorted(WORKLOADS),
                    help="cfg4 = the headline workload; the others are extra measurements for BASELINE.md's table")
    ap.add_argument("--utts-per-gpu", type=int, default=0)
    ap.add_argument("--precision", default="fp16", choices=["fp16", "bf16"])
    ap.add_argument("--max-batch-rows", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    global TOPOLOGY, UTTS_PER_GPU, MIN_FRAMES, MAX_FRAMES, CHUNK_SIZE, LENGTH_SEED, WORKLOAD_DESC
    TOPOLOGY, UTTS_PER_GPU, MIN_FRAMES, MAX_FRAMES, CHUNK_SIZE, LENGTH_SEED, WORKLOAD_DESC = WORKLOADS[args.workload]
    if args.utts_per_gpu <= 0:
        args.utts_per_gpu = UTTS_PER_GPU

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
        dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
        return float(t.item())

    cvx = importlib.import_module(PKG)
    kf = importlib.import_module(PKG + ".kaldi_format")
    sy = importlib.import_module(PKG + ".synth")
    sh = importlib.import_module(PKG + ".shard")

    # ---- workload: this rank's LPT shard of N x 1250 utterances -------------------
    lengths_all = sy.make_lengths(world * args.utts_per_gpu, MIN_FRAMES, MAX_FRAMES, LENGTH_SEED)
    mine = sh.lpt_shards(lengths_all, world)[rank]
    lengths = lengths_all[mine]
    total_rows = int(lengths.sum())
    tmp = os.environ.get("TMPDIR", "/tmp")
    model_path = os.path.join(tmp, f"cvx_bench_{TOPOLOGY}_{os.getpid()}.raw")
    kf.write_nnet3(model_path, sy.make_model(TOPOLOGY, seed=42), binary=True)
    model = cvx.Model(model_path, sy.EMBEDDING_NODE[TOPOLOGY])
    os.unlink(model_path)
    info = model.info
    pinned = [cvx.PinnedArray((total_rows, d)) for d in model.input_dims]
    rng = np.random.default_rng(LENGTH_SEED * 1000 + rank)
    step_rows = 1 << 20
    for pa, d in zip(pinned, model.input_dims):
        for r0 in range(0, total_rows, step_rows):
            r1 = min(total_rows, r0 + step_rows)
            pa.array[r0:r1] = rng.standard_normal((r1 - r0, d), dtype=np.float32)
    starts, lens, utts = [], [], []
    base = 0
    for u, T in enumerate(lengths.tolist()):
        for s, n in cvx.chunk_plan(T, CHUNK_SIZE, MIN_CHUNK_SIZE, False):
            starts.append(base + s); lens.append(n); utts.append(u)
        base += T
    starts, lens, utts = np.array(starts, np.int64), np.array(lens, np.int32), np.array(utts, np.int32)
    n_utts = len(lengths)
    audio_sec = total_rows / 100.0

    ex = cvx.Extractor(model, device=local_rank, precision=args.precision, max_batch_rows=args.max_batch_rows,
                       profile=True)
    lib = cvx.load_library()
    import ctypes as C
    dptrs, nbytes = [], 0
    for pa, d in zip(pinned, model.input_dims):
        dp = C.c_void_p()
        nb = total_rows * d * 4
        assert lib.cvx_device_alloc(local_rank, C.byref(dp), nb) == 0, lib.cvx_last_error()
        assert lib.cvx_memcpy_h2d(local_rank, dp, pa.array.ctypes.data_as(C.c_void_p), nb) == 0
        dptrs.append(dp.value)
        nbytes += nb
    out = np.zeros((n_utts, info.embedding_dim), np.float32)

    def step_device():
        ex.extract_chunks(dptrs, total_rows, starts, lens, utts, None, n_utts, out=out, flags=cvx.FLAG_DEVICE_INPUT)
        return ex.last_extract_ms()

    def step_host():
        t0 = time.perf_counter()
        ex.extract_chunks([pa.array for pa in pinned], total_rows, starts, lens, utts, None, n_utts, out=out)
        return (time.perf_counter() - t0) * 1e3

    # ---- device-resident leg -------------------------------------------------------
    for _ in range(args.warmup):
        step_device()
    ex.reset_profile()
    launches0 = ex.launch_count()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    barrier()
    if sampler:
        sampler.start()
    t_wall0 = time.perf_counter()
    dev_ms = [step_device() for _ in range(args.steps)]
    barrier()
    wall_ms = (time.perf_counter() - t_wall0) * 1e3
    launches = ex.launch_count() - launches0
    prof = {p["name"]: p for p in ex.profile()}
    emb_dev = out.copy()
    total_ms = reduce(float(sum(dev_ms)), "max")
    total_audio = reduce(audio_sec, "sum") * args.steps
    value = total_audio / (total_ms / 1e3)

    # ---- end-to-end leg (host buffers) ----------------------------------------------
    for _ in range(2):
        step_host()
    barrier()
    e2e_ms = [step_host() for _ in range(args.steps)]
    barrier()
    clocks = sampler.stop() if sampler else None     # sampled over both timed regions
    assert np.array_equal(out, emb_dev), "host-input and device-input passes disagree"
    e2e_total_ms = reduce(float(sum(e2e_ms)), "max")
    e2e_value = total_audio / (e2e_total_ms / 1e3)
    total_launches = int(reduce(float(launches), "sum"))
    finite = bool(np.isfinite(out).all() and np.abs(out).sum() > 0)

    if rank == 0:
        pk = peaks()
        g = prof["tdnn_gemm"]
        gemm_tflops = g["flops"] / (g["total_ms"] * 1e-3) / 1e12
        pl = prof["stats_pool_partial"]
        pool_note = "stand-alone pooling kernel inside the timed steps"
        if pl["launches"] == 0:
            # default path: pooling is fused into the pooled layer's GEMM epilogue.  The stand-alone
            # HBM-bound kernel (CVX_FUSED_POOL=0) is measured in a separate, untimed pass.
            os.environ["CVX_FUSED_POOL"] = "0"
            ex2 = cvx.Extractor(model, device=local_rank, precision=args.precision, max_batch_rows=args.max_batch_rows,
                                profile=True)
            for i in range(3):
                if i == 1:
                    ex2.reset_profile()
                ex2.extract_chunks(dptrs, total_rows, starts, lens, utts, None, n_utts, out=np.zeros_like(out),
                                   flags=cvx.FLAG_DEVICE_INPUT)
            pl = {p["name"]: p for p in ex2.profile()}["stats_pool_partial"]
            ex2.close()
            del os.environ["CVX_FUSED_POOL"]
            pool_note = ("pooling is fused into the pooled layer's GEMM epilogue on the timed path; this is the "
                         "stand-alone kernel (CVX_FUSED_POOL=0) measured in a separate pass of the same workload")
        pool_gbs = pl["bytes"] / (pl["total_ms"] * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get("tdnn_gemm_dram_bytes_per_launch")
        kernel_ms = {k: round(v["total_ms"] / args.steps, 3) for k, v in prof.items() if not k.startswith("layer:")}
        # per frame layer inside the same timed steps: ms per step and achieved TFLOP/s (sustained clocks)
        layer_ms = {k[6:]: {"ms": round(v["total_ms"] / args.steps, 3),
                            "tflops": round(v["flops"] / max(v["total_ms"], 1e-9) * 1e-9, 1)}
                    for k, v in prof.items() if k.startswith("layer:")}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.precision, "data": "synthetic",
            "config": {"workload": WORKLOAD_DESC, "workload_id": args.workload,
                       "topology": TOPOLOGY, "utts_per_gpu": args.utts_per_gpu, "frames_per_gpu": total_rows,
                       "chunk_size": CHUNK_SIZE, "min_chunk_size": MIN_CHUNK_SIZE, "parallelism": f"utterance-shards x{world} (LPT, no collective)",
                       "operands": f"{args.precision} operands, fp32 accumulate; pooling+tdnn6 fp32",
                       "l2": f"inputs larger than L2 ({nbytes / 1e9:.2f} GB features, multi-GB activations per step)",
                       "mac_per_frame": int(info.mac_per_frame), "outputs_finite": finite},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(nbytes), "d2h_bytes_per_step": int(out.nbytes),
                    "ms_per_step": e2e_total_ms / args.steps, "timer": "wall clock around the synchronous C-ABI call, max over ranks"},
            "gpu_launches": total_launches,
            "wall_ms_per_step": wall_ms / args.steps,
            "kernel_ms_per_step": kernel_ms,
            "layer_ms_per_step": layer_ms,
            "roofline": {"kernel": "tdnn_gemm (tcgen05 implicit GEMM, all frame layers)", "bound": "tensor",
                         "achieved": gemm_tflops, "peak": pk["tflops"], "unit": "TFLOP/s", "frac": gemm_tflops / pk["tflops"],
                         "traffic": traffic, "peak_source": pk["source"],
                         "launches": int(g["launches"]), "avg_launch_ms": g["total_ms"] / max(g["launches"], 1)},
            "roofline_pool": {"kernel": "stats_pool_partial", "bound": "hbm", "achieved": pool_gbs, "peak": pk["hbm"],
                              "unit": "GB/s", "frac": pool_gbs / pk["hbm"], "traffic": None,
                              "launches": int(pl["launches"]), "note": pool_note},
        }
        if world == 1 and not args.no_cpu_baseline:
            from oracle import cpu_baseline
            cores = os.cpu_count() or 1
            r = cpu_baseline.run(TOPOLOGY, cores, 12.0, MIN_FRAMES, MAX_FRAMES, CHUNK_SIZE, MIN_CHUNK_SIZE)
            line["cpu_baseline"] = {"value": r["audio_sec_per_sec"], "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{r['utts']} utterances / {r['frames']} frames of the same workload, one "
                                              f"single-threaded fp32 job per core ({cores}), {r['wall_s']:.1f} s wall"}
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
