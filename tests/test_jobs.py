"""Job fan-out of the drop-in programs (host logic, no GPU): length-balanced shards and the
device lock protocol that spreads a recipe stage's nj processes over the GPUs of a box
(R/egs/sre/v5/sid/nnet3/xvector/extract_xvectors_new.sh:72-74, 83-89)."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("n,shards,seed", [(0, 3, 0), (1, 4, 1), (7, 7, 2), (1250, 8, 3), (10000, 8, 4), (999, 5, 5)])
def test_c_abi_lpt_matches_python_sharding(cvx, n, shards, seed):
    """cvx_lpt_shards (what --shard=J/N uses) == shard.lpt_shards (what bench.py uses)."""
    import importlib
    sh = importlib.import_module(cvx.__name__ + ".shard")
    rng = np.random.default_rng(seed)
    lengths = rng.integers(1000, 30001, size=n).astype(np.int64)
    if n > 10:
        lengths[5] = lengths[3]                      # ties
    got = cvx.lpt_shards(lengths, shards)
    want = np.zeros(n, np.int32)
    for s, idx in enumerate(sh.lpt_shards(lengths, shards)):
        want[idx] = s
    assert np.array_equal(got, want)
    if n >= shards * 50:
        loads = np.bincount(got, weights=lengths, minlength=shards)
        assert loads.max() - loads.min() <= lengths.max()


def test_lpt_rejects_bad_arguments(cvx):
    lib = cvx.load_library()
    out = np.zeros(3, np.int32)
    ln = np.ones(3, np.int64)
    assert lib.cvx_lpt_shards(ln.ctypes.data_as(C.c_void_p), 3, 0, out.ctypes.data_as(C.c_void_p)) == -1
    assert lib.cvx_lpt_shards(None, 3, 2, out.ctypes.data_as(C.c_void_p)) == -1


_HOLDER = r"""
import ctypes as C, importlib, sys
sys.path.insert(0, {root!r})
cvx = importlib.import_module({pkg!r})
lib = cvx.load_library()
d, s = C.c_int32(-1), C.c_int32(-1)
rc = lib.cvx_test_acquire_lock({dir!r}.encode(), {ndev}, 0, C.byref(d), C.byref(s))
print(rc, d.value, s.value, flush=True)
sys.stdin.readline()
"""


def test_device_locks_spread_concurrent_jobs(cvx, tmp_path):
    """five jobs on a two-GPU box: the first two get a free device each, the others share evenly;
    when a job exits its device becomes free again."""
    code = _HOLDER.format(root=ROOT, pkg=cvx.__name__, dir=str(tmp_path), ndev=2)
    procs, got = [], []
    try:
        for _ in range(5):
            p = subprocess.Popen([sys.executable, "-c", code], stdin=subprocess.PIPE, stdout=subprocess.PIPE, text=True)
            procs.append(p)
            rc, d, s = map(int, p.stdout.readline().split())
            assert rc == 0
            got.append((d, s))
        assert got == [(0, 0), (1, 0), (0, 1), (1, 1), (0, 2)]
        # the job holding device 1 exclusively ends: the next job gets device 1 for itself
        procs[1].stdin.write("\n"); procs[1].stdin.flush(); procs[1].wait(timeout=30)
        p = subprocess.Popen([sys.executable, "-c", code], stdin=subprocess.PIPE, stdout=subprocess.PIPE, text=True)
        procs.append(p)
        assert tuple(map(int, p.stdout.readline().split())) == (0, 1, 0)
    finally:
        for p in procs:
            if p.poll() is None:
                p.stdin.write("\n"); p.stdin.flush()
                p.wait(timeout=30)


_HOLDER_BY_ID = r"""
import ctypes as C, importlib, os, sys
sys.path.insert(0, {root!r})
os.environ["CVX_LOCK_DIR"] = {dir!r}
cvx = importlib.import_module({pkg!r})
lib = cvx.load_library()
ids = {ids!r}
arr = (C.c_char_p * len(ids))(*[i.encode() for i in ids])
lib.cvx_acquire_device_by_id.argtypes = [C.POINTER(C.c_char_p), C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
d, s = C.c_int32(-1), C.c_int32(-1)
rc = lib.cvx_acquire_device_by_id(arr, len(ids), 0, C.byref(d), C.byref(s))
print(rc, d.value, s.value, flush=True)
sys.stdin.readline()
"""


def test_pre_cuda_selection_shares_lock_names_with_cuda_enumeration(cvx, tmp_path):
    """a job that picks its GPU from /proc's PCI ids before initialising CUDA (lower-case hex) and one that
    enumerates through CUDA (upper-case hex) must exclude each other on the same device."""
    procs, got = [], []
    try:
        for ids in (["0000:1b:00.0", "0000:2c:00.0"], ["0000:1B:00.0", "0000:2C:00.0"], ["0000:1b:00.0", "0000:2c:00.0"]):
            code = _HOLDER_BY_ID.format(root=ROOT, pkg=cvx.__name__, dir=str(tmp_path), ids=ids)
            p = subprocess.Popen([sys.executable, "-c", code], stdin=subprocess.PIPE, stdout=subprocess.PIPE, text=True)
            procs.append(p)
            got.append(tuple(map(int, p.stdout.readline().split())))
        assert got == [(0, 0, 0), (0, 1, 0), (0, 0, 1)]
    finally:
        for p in procs:
            if p.poll() is None:
                p.stdin.write("\n"); p.stdin.flush()
                p.wait(timeout=30)


def test_ivector_mean_writes_text_by_default(cvx, kf, tmp_path):
    """upstream's option is --binary-write with default false; --binary stays as an alias."""
    vecs = [(f"u{i}", np.arange(4, dtype=np.float32) + i) for i in range(3)]
    kf.write_ark(str(tmp_path / "v.ark"), vecs)
    exe = os.path.join(cvx.BIN_DIR, "ivector-mean")
    r = subprocess.run([exe, f"ark:{tmp_path}/v.ark", f"{tmp_path}/mean.vec"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    txt = open(tmp_path / "mean.vec", "rb").read()
    assert txt.startswith(b" [") and np.allclose([float(x) for x in txt.decode().strip(" []\n").split()], [1, 2, 3, 4])
    for flag in ("--binary-write=true", "--binary=true"):
        r = subprocess.run([exe, flag, f"ark:{tmp_path}/v.ark", f"{tmp_path}/mean.bin"], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        assert open(tmp_path / "mean.bin", "rb").read().startswith(b"\0BFV ")


def test_shard_option_is_validated(cvx):
    exe = os.path.join(cvx.BIN_DIR, "nnet3-xvector-compute")
    for bad in ("3", "0/4", "5/4", "a/b", "1/0"):
        r = subprocess.run([exe, f"--shard={bad}", "final.raw", "ark:/dev/null", "ark:/dev/null"], capture_output=True, text=True)
        assert r.returncode != 0 and "--shard expects J/N" in r.stderr, (bad, r.stderr)


def test_selection_among_cuda_visible_devices_entries_happens_before_cuda(cvx, tmp_path):
    """With CUDA_VISIBLE_DEVICES set (container runtimes do), a job picks one ENTRY of the list through the lock
    files and makes it the only visible device -- no CUDA call is needed for the choice (it is logged before
    the model is even opened, so this runs without a GPU)."""
    import fcntl
    exe = os.path.join(cvx.BIN_DIR, "nnet3-xvector-compute")
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="3, GPU-AbC", CVX_LOCK_DIR=str(tmp_path))
    env.pop("CVX_DEVICE", None)
    env.pop("CVX_SELECT_WITH_CUDA", None)

    def run():
        r = subprocess.run([exe, "--use-gpu=yes", str(tmp_path / "missing.raw"), "ark:/dev/null", "ark:/dev/null"],
                           capture_output=True, text=True, env=env)
        assert r.returncode != 0  # no model (and here no GPU): the job fails AFTER the selection
        return [ln for ln in r.stderr.splitlines() if "GPU" in ln and ("Selected" in ln or "sharing" in ln)]

    assert "Selected GPU 3 (entry 0 of CUDA_VISIBLE_DEVICES=3, GPU-AbC" in run()[0]
    held = []
    try:
        for name in ("cvd-3.0", "cvd-gpu-abc.0"):  # lock names are lower-cased ids; round 0 = exclusive use
            f = open(tmp_path / f"cvx-gpu-{os.getuid()}-{name}", "a+")
            fcntl.flock(f, fcntl.LOCK_EX | fcntl.LOCK_NB)
            held.append(f)
            line = run()[0]
            if len(held) == 1:
                assert "Selected GPU GPU-AbC (entry 1 of" in line, line
            else:
                assert "sharing GPU 3 with 1 other job(s)" in line, line
    finally:
        for f in held:
            f.close()
    # a list CUDA itself would cut short is left to the CUDA path
    env["CUDA_VISIBLE_DEVICES"] = "-1"
    assert run() == []
