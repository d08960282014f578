"""B200-native c-vector / x-vector extraction: Python host-side mirror.

Thin ctypes binding over the C ABI in include/cvx.h (lib/libcvx.so).  It mirrors
the reference's interface for this path -- the command line of
`nnet3-xvector-compute` driven by
egs/sre/v5/sid/nnet3/xvector/extract_xvectors_new.sh:84-93 -- as a function
(`Extractor.extract`) with the same knobs (`chunk_size`, `min_chunk_size`,
`pad_input`, output node), so tests read like the reference's call sites.

There is no CPU fallback: if lib/libcvx.so is missing, `load_library()` raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# CVX_LIB: another build of the library (A/B runs of kernel variants); the default is the in-tree build
LIB_PATH = os.environ.get("CVX_LIB") or os.path.join(_HERE, "lib", "libcvx.so")
BIN_DIR = os.path.join(_HERE, "bin")

PRECISION = {"fp16": 0, "bf16": 1, "fp32": 2}
FLAG_DEVICE_INPUT = 1
FLAG_DEVICE_OUTPUT = 2
FLAG_APPLY_EXP = 4


class CvxError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"cvx error {code}: {msg}")
        self.code = code


class ModelInfo(C.Structure):
    _fields_ = [("num_inputs", C.c_int32), ("embedding_dim", C.c_int32), ("left_context", C.c_int32),
                ("right_context", C.c_int32), ("pooled_dim", C.c_int32), ("num_frame_layers", C.c_int32),
                ("num_segment_layers", C.c_int32), ("mac_per_frame", C.c_int64), ("mac_per_chunk", C.c_int64)]


class EngineOptions(C.Structure):
    _fields_ = [("device", C.c_int32), ("precision", C.c_int32), ("max_batch_rows", C.c_int64),
                ("profile", C.c_int32), ("reserved", C.c_int32 * 7)]


class KernelProfile(C.Structure):
    _fields_ = [("name", C.c_char * 48), ("launches", C.c_int64), ("total_ms", C.c_double),
                ("flops", C.c_double), ("bytes", C.c_double)]


class VectorChain(C.Structure):
    _fields_ = [("mean", C.c_void_p), ("transform", C.c_void_p), ("transform_rows", C.c_int32),
                ("transform_cols", C.c_int32), ("normalize_length", C.c_int32), ("scaleup", C.c_int32)]


class PldaModel(C.Structure):
    _fields_ = [("mean", C.c_void_p), ("transform", C.c_void_p), ("psi", C.c_void_p), ("dim_in", C.c_int32),
                ("dim", C.c_int32), ("normalize_length", C.c_int32), ("simple_length_norm", C.c_int32)]


def _make_chain(mean, transform, normalize_length, scaleup):
    """-> (VectorChain, keep-alive arrays)."""
    ch = VectorChain()
    keep = []
    if mean is not None:
        m = np.ascontiguousarray(mean, np.float32)
        keep.append(m)
        ch.mean = m.ctypes.data
    if transform is not None:
        t = np.ascontiguousarray(transform, np.float32)
        keep.append(t)
        ch.transform = t.ctypes.data
        ch.transform_rows, ch.transform_cols = t.shape
    ch.normalize_length = int(bool(normalize_length))
    ch.scaleup = int(bool(scaleup))
    return ch, keep


#: every symbol include/cvx.h declares (tests check the library exports them all)
ABI_SYMBOLS = [
    "cvx_last_error", "cvx_version", "cvx_model_load", "cvx_model_load_frames", "cvx_model_from_plan", "cvx_model_free",
    "cvx_model_get_info", "cvx_model_input_name", "cvx_model_input_dim", "cvx_model_describe",
    "cvx_engine_options_default", "cvx_engine_create", "cvx_engine_destroy", "cvx_extract",
    "cvx_engine_last_chunk_embeddings", "cvx_extract_raw", "cvx_compute_frames", "cvx_engine_set_vector_chain", "cvx_engine_output_dim", "cvx_transform_vectors", "cvx_plda_score",
    "cvx_test_prestep", "cvx_host_alloc", "cvx_host_free", "cvx_device_alloc",
    "cvx_device_free", "cvx_memcpy_h2d", "cvx_device_count", "cvx_engine_get_profile",
    "cvx_engine_reset_profile", "cvx_engine_launch_count", "cvx_engine_last_extract_ms", "cvx_test_tdnn_layer", "cvx_test_tdnn_layer_bench", "cvx_test_stats_pool",
    "cvx_acquire_device", "cvx_acquire_device_by_id", "cvx_test_acquire_lock", "cvx_lpt_shards", "cvx_engine_saturation_counts",
    "cvx_test_segment_affine",
]

_lib = None


def build(verbose: bool = False) -> None:
    """Compile lib/libcvx.so and the drop-in programs in-tree (nvcc, sm_100a)."""
    r = subprocess.run(["make", "-j8", "-C", _HERE], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:])
        print(r.stderr[-8000:])
    if r.returncode != 0:
        raise RuntimeError("building libcvx.so failed")


def load_library() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback for this path)")
    lib = C.CDLL(LIB_PATH)
    lib.cvx_last_error.restype = C.c_char_p
    lib.cvx_version.restype = C.c_char_p
    lib.cvx_model_load.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(C.c_void_p)]
    lib.cvx_model_load_frames.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(C.c_void_p)]
    lib.cvx_model_free.argtypes = [C.c_void_p]
    lib.cvx_model_get_info.argtypes = [C.c_void_p, C.POINTER(ModelInfo)]
    lib.cvx_model_input_name.argtypes = [C.c_void_p, C.c_int32]
    lib.cvx_model_input_name.restype = C.c_char_p
    lib.cvx_model_input_dim.argtypes = [C.c_void_p, C.c_int32]
    lib.cvx_model_describe.argtypes = [C.c_void_p]
    lib.cvx_model_describe.restype = C.c_char_p
    lib.cvx_engine_options_default.argtypes = [C.POINTER(EngineOptions)]
    lib.cvx_engine_create.argtypes = [C.c_void_p, C.POINTER(EngineOptions), C.POINTER(C.c_void_p)]
    lib.cvx_engine_destroy.argtypes = [C.c_void_p]
    lib.cvx_extract.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_int32, C.c_int64, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_uint32]
    lib.cvx_compute_frames.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_int32, C.c_int64, C.c_void_p, C.c_int32,
                                       C.c_void_p, C.c_uint32]
    lib.cvx_engine_set_vector_chain.argtypes = [C.c_void_p, C.POINTER(VectorChain)]
    lib.cvx_engine_output_dim.argtypes = [C.c_void_p]
    lib.cvx_engine_output_dim.restype = C.c_int32
    lib.cvx_transform_vectors.argtypes = [C.c_int32, C.c_void_p, C.c_int64, C.c_int32, C.POINTER(VectorChain), C.c_void_p,
                                          C.c_void_p]
    lib.cvx_plda_score.argtypes = [C.c_int32, C.POINTER(PldaModel), C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32,
                                   C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    lib.cvx_engine_last_chunk_embeddings.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
    lib.cvx_host_alloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
    lib.cvx_host_free.argtypes = [C.c_void_p]
    lib.cvx_device_alloc.argtypes = [C.c_int32, C.POINTER(C.c_void_p), C.c_size_t]
    lib.cvx_device_free.argtypes = [C.c_int32, C.c_void_p]
    lib.cvx_memcpy_h2d.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_size_t]
    lib.cvx_device_count.argtypes = [C.POINTER(C.c_int32)]
    lib.cvx_engine_get_profile.argtypes = [C.c_void_p, C.POINTER(KernelProfile), C.c_int32, C.POINTER(C.c_int32)]
    lib.cvx_engine_reset_profile.argtypes = [C.c_void_p]
    lib.cvx_engine_saturation_counts.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(C.c_int32)]
    lib.cvx_engine_launch_count.argtypes = [C.c_void_p]
    lib.cvx_engine_launch_count.restype = C.c_int64
    lib.cvx_engine_last_extract_ms.argtypes = [C.c_void_p]
    lib.cvx_engine_last_extract_ms.restype = C.c_double
    lib.cvx_test_tdnn_layer.argtypes = [C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int32,
                                        C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32,
                                        C.c_void_p]
    lib.cvx_test_tdnn_layer_bench.argtypes = [C.c_int32, C.c_int32, C.c_int64, C.c_int32, C.c_void_p, C.c_int32, C.c_int32,
                                              C.c_int32, C.POINTER(C.c_double)]
    lib.cvx_test_stats_pool.argtypes = [C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p,
                                        C.c_int64, C.c_float, C.c_void_p]
    lib.cvx_acquire_device.argtypes = [C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    lib.cvx_test_acquire_lock.argtypes = [C.c_char_p, C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    lib.cvx_lpt_shards.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]
    _lib = lib
    return lib


def _check(code: int) -> None:
    if code != 0:
        raise CvxError(code, load_library().cvx_last_error().decode(errors="replace"))


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def device_count() -> int:
    n = C.c_int32(0)
    load_library().cvx_device_count(C.byref(n))
    return n.value


def lpt_shards(lengths: Sequence[int], num_shards: int) -> np.ndarray:
    """shard index per utterance from the C ABI's cvx_lpt_shards (the partition the drop-in
    programs use for --shard=J/N --utt2num-frames=...; same rule as shard.lpt_shards)."""
    ln = np.ascontiguousarray(lengths, np.int64)
    out = np.zeros(len(ln), np.int32)
    _check(load_library().cvx_lpt_shards(_ptr(ln), len(ln), num_shards, _ptr(out)))
    return out


class PinnedArray:
    """fp32 matrix in pinned host memory (cvx_host_alloc) exposed as numpy."""

    def __init__(self, shape: Tuple[int, ...]):
        self._lib = load_library()
        n = int(np.prod(shape)) * 4
        p = C.c_void_p()
        _check(self._lib.cvx_host_alloc(C.byref(p), n))
        self._p = p
        buf = (C.c_float * (n // 4)).from_address(p.value)
        self.array = np.frombuffer(buf, dtype=np.float32).reshape(shape)

    def close(self):
        if self._p:
            self.array = None
            self._lib.cvx_host_free(self._p)
            self._p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Model:
    """nnet3 raw model lowered to the extraction plan (cvx_model_load), or with
    ``frames=True`` to the frame-output plan nnet3-compute runs (cvx_model_load_frames)."""

    def __init__(self, rxfilename: str, output_node: Optional[str] = None, frames: bool = False):
        self._lib = load_library()
        h = C.c_void_p()
        load = self._lib.cvx_model_load_frames if frames else self._lib.cvx_model_load
        _check(load(rxfilename.encode(), output_node.encode() if output_node else None, C.byref(h)))
        self.frames = frames
        self._h = h
        info = ModelInfo()
        _check(self._lib.cvx_model_get_info(h, C.byref(info)))
        self.info = info
        self.input_names = [self._lib.cvx_model_input_name(h, i).decode() for i in range(info.num_inputs)]
        self.input_dims = [self._lib.cvx_model_input_dim(h, i) for i in range(info.num_inputs)]

    def describe(self) -> str:
        return self._lib.cvx_model_describe(self._h).decode()

    def close(self):
        if self._h:
            self._lib.cvx_model_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def chunk_plan(T: int, chunk_size: int, min_chunk_size: int, pad_input: bool) -> List[Tuple[int, int]]:
    """(start, rows) of the chunks an utterance of T rows is cut into; mirrors the
    chunk loop of nnet3-xvector-compute -- same rule as
    csrc/bin/extract_main.cc."""
    if T == 0 or (not pad_input and T < min_chunk_size):
        return []
    c = chunk_size if (chunk_size != -1 and T >= chunk_size) else T
    out = []
    for i in range((T + c - 1) // c):
        n = min(c, T - i * c)
        if n < min_chunk_size and not pad_input:
            continue
        out.append((i * c, n))
    return out


class Extractor:
    """One GPU engine (cvx_engine) for one model."""

    def __init__(self, model: Model, device: int = 0, precision: str = "fp16", max_batch_rows: int = 0,
                 profile: bool = False, check_saturation: bool = False):
        self._lib = load_library()
        self.model = model
        o = EngineOptions()
        self._lib.cvx_engine_options_default(C.byref(o))
        o.device, o.precision, o.max_batch_rows, o.profile = device, PRECISION[precision], max_batch_rows, int(profile)
        o.reserved[1] = int(check_saturation)
        h = C.c_void_p()
        _check(self._lib.cvx_engine_create(model._h, C.byref(o), C.byref(h)))
        self._h = h
        self.device = device

    def close(self):
        if self._h:
            self._lib.cvx_engine_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- embedding post-step ----------------------------------------------
    def set_vector_chain(self, mean=None, transform=None, normalize_length: bool = True, scaleup: bool = True,
                         clear: bool = False) -> None:
        """Install `ivector-subtract-global-mean | transform-vec | ivector-normalize-length` as a device
        post-step of extract()/extract_chunks() (cvx_engine_set_vector_chain); clear=True removes it."""
        if clear:
            _check(self._lib.cvx_engine_set_vector_chain(self._h, None))
            return
        ch, keep = _make_chain(mean, transform, normalize_length, scaleup)
        _check(self._lib.cvx_engine_set_vector_chain(self._h, C.byref(ch)))

    @property
    def output_dim(self) -> int:
        return int(self._lib.cvx_engine_output_dim(self._h))

    # -- raw ABI call ------------------------------------------------------
    def extract_chunks(self, feats: Sequence, total_rows: int, chunk_start: np.ndarray, chunk_len: np.ndarray,
                       chunk_utt: np.ndarray, chunk_weight: Optional[np.ndarray], num_utts: int,
                       out=None, flags: int = 0) -> np.ndarray:
        """feats: list of np.float32 [total_rows, dim] arrays (host) or raw device
        pointers (ints) with FLAG_DEVICE_INPUT."""
        n_s = len(feats)
        ptrs = (C.c_void_p * n_s)()
        keep = []
        for i, f in enumerate(feats):
            if isinstance(f, np.ndarray):
                assert f.dtype == np.float32 and f.flags.c_contiguous
                keep.append(f)
                ptrs[i] = f.ctypes.data
            else:
                ptrs[i] = int(f)
        cs = np.ascontiguousarray(chunk_start, dtype=np.int64)
        cl = np.ascontiguousarray(chunk_len, dtype=np.int32)
        cu = np.ascontiguousarray(chunk_utt, dtype=np.int32)
        cw = None if chunk_weight is None else np.ascontiguousarray(chunk_weight, dtype=np.float32)
        E = self.output_dim
        if out is None:
            out = np.zeros((num_utts, E), dtype=np.float32)
        out_ptr = out.ctypes.data_as(C.c_void_p) if isinstance(out, np.ndarray) else C.c_void_p(int(out))
        _check(self._lib.cvx_extract(self._h, ptrs, n_s, total_rows, _ptr(cs), _ptr(cl), _ptr(cu), _ptr(cw),
                                     len(cs), num_utts, out_ptr, flags))
        return out

    # -- reference-shaped call --------------------------------------------
    def extract(self, feats: Dict[str, np.ndarray], lengths: Sequence[int], chunk_size: int = -1,
                min_chunk_size: int = 100, pad_input: bool = False) -> Tuple[np.ndarray, np.ndarray]:
        """feats: input-node name -> packed fp32 rows of all utterances; lengths:
        rows per utterance.  Returns ([n_utts, dim], ok mask): what
        `nnet3-xvector-compute --chunk-size= --min-chunk-size= [--pad-input=]` writes
        (failed utterances: ok False, zeros)."""
        names = self.model.input_names
        mats = [np.ascontiguousarray(feats[n], dtype=np.float32) for n in names]
        starts, lens, utts, wts = [], [], [], []
        ok = np.zeros(len(lengths), dtype=bool)
        ctx = self.model.info.left_context + self.model.info.right_context + 1
        if pad_input:
            # materialise replicated edge rows chunk by chunk, like the C++ driver
            rows_out = [[] for _ in names]
            base, pos = 0, 0
            for u, T in enumerate(lengths):
                for (s, n) in chunk_plan(T, chunk_size, min_chunk_size, True):
                    lp = (min_chunk_size - n) // 2 if n < min_chunk_size else 0
                    rp = (min_chunk_size - n - lp) if n < min_chunk_size else 0
                    idx = np.concatenate([np.full(lp, base + s), np.arange(base + s, base + s + n),
                                          np.full(rp, base + s + n - 1)]).astype(np.int64)
                    if len(idx) >= ctx:
                        starts.append(pos); lens.append(len(idx)); utts.append(u); wts.append(float(n))
                        ok[u] = True
                    for k in range(len(names)):
                        rows_out[k].append(mats[k][idx])
                    pos += len(idx)
                base += T
            mats = [np.ascontiguousarray(np.concatenate(r, axis=0)) if r else np.zeros((0, m.shape[1]), np.float32)
                    for r, m in zip(rows_out, mats)]
            total = pos
        else:
            base = 0
            for u, T in enumerate(lengths):
                for (s, n) in chunk_plan(T, chunk_size, min_chunk_size, False):
                    if n >= ctx:
                        starts.append(base + s); lens.append(n); utts.append(u); wts.append(float(n))
                        ok[u] = True
                base += T
            total = base
        out = self.extract_chunks(mats, total, np.array(starts, np.int64), np.array(lens, np.int32),
                                  np.array(utts, np.int32), np.array(wts, np.float32), len(lengths))
        return out, ok

    def compute_frames(self, feats: Dict[str, np.ndarray], lengths: Sequence[int], apply_exp: bool = False,
                       flags: int = 0, out: Optional[np.ndarray] = None) -> np.ndarray:
        """Frame-output models: the output layer at every frame of every utterance, what
        `nnet3-compute [--apply-exp=true]` writes row for row (cvx_compute_frames).
        feats: input-node name -> packed fp32 rows; lengths: rows per utterance; out: optional
        preallocated [rows, dim] fp32 result (a PinnedArray's .array for full PCIe speed)."""
        names = self.model.input_names
        mats = [np.ascontiguousarray(feats[n], dtype=np.float32) for n in names]
        start = np.zeros(len(lengths) + 1, dtype=np.int64)
        np.cumsum(np.asarray(lengths, dtype=np.int64), out=start[1:])
        total = int(start[-1])
        ptrs = (C.c_void_p * len(mats))()
        for i, m in enumerate(mats):
            ptrs[i] = m.ctypes.data
        if out is None:
            out = np.zeros((total, self.model.info.embedding_dim), dtype=np.float32)
        assert out.dtype == np.float32 and out.flags.c_contiguous and out.shape == (total, self.model.info.embedding_dim)
        _check(self._lib.cvx_compute_frames(self._h, ptrs, len(mats), total, _ptr(start), len(lengths), _ptr(out),
                                            flags | (FLAG_APPLY_EXP if apply_exp else 0)))
        return out

    def last_chunk_embeddings(self, num_chunks: int) -> np.ndarray:
        out = np.zeros((num_chunks, self.model.info.embedding_dim), dtype=np.float32)
        _check(self._lib.cvx_engine_last_chunk_embeddings(self._h, _ptr(out), num_chunks))
        return out

    def profile(self) -> List[dict]:
        recs = (KernelProfile * 64)()   # kernel classes first, then one "layer:<name>" record per frame layer
        n = C.c_int32(0)
        _check(self._lib.cvx_engine_get_profile(self._h, recs, 64, C.byref(n)))
        return [dict(name=recs[i].name.decode(), launches=recs[i].launches, total_ms=recs[i].total_ms,
                     flops=recs[i].flops, bytes=recs[i].bytes) for i in range(n.value)]

    def saturation_counts(self) -> List[int]:
        """check_saturation=True engines: [staged inputs, frame layer 0, 1, ...] values clipped at the fp16 limit."""
        buf = (C.c_int64 * 64)()
        n = C.c_int32(0)
        _check(self._lib.cvx_engine_saturation_counts(self._h, buf, 64, C.byref(n)))
        return [int(buf[i]) for i in range(n.value)]

    def reset_profile(self) -> None:
        self._lib.cvx_engine_reset_profile(self._h)

    def launch_count(self) -> int:
        return int(self._lib.cvx_engine_launch_count(self._h))

    def last_extract_ms(self) -> float:
        return float(self._lib.cvx_engine_last_extract_ms(self._h))


# -- single-kernel test entries ------------------------------------------------

def test_tdnn_layer(x: np.ndarray, offsets: Sequence[int], weight: np.ndarray, bias=None, relu=False, scale=None,
                    offset=None, precision: str = "fp16", device: int = 0, compact: bool = False) -> np.ndarray:
    lib = load_library()
    x = np.ascontiguousarray(x, np.float32)
    w = np.ascontiguousarray(weight, np.float32)
    offs = np.ascontiguousarray(offsets, np.int32)
    out = np.zeros((x.shape[0], w.shape[0]), np.float32)
    f = lambda a: None if a is None else np.ascontiguousarray(a, np.float32)
    b, s, o = f(bias), f(scale), f(offset)
    _check(lib.cvx_test_tdnn_layer(device, PRECISION[precision] | (0x100 if compact else 0), _ptr(x), x.shape[0],
                                   x.shape[1], _ptr(offs), len(offs), _ptr(w), _ptr(b), int(relu), _ptr(s), _ptr(o),
                                   w.shape[0], _ptr(out)))
    return out


def bench_tdnn_layer(rows: int, in_dim: int, offsets: Sequence[int], out_dim: int, iters: int = 10,
                     precision: str = "fp16", device: int = 0) -> float:
    """avg ms per launch of the production TDNN GEMM kernel for one layer shape."""
    lib = load_library()
    offs = np.ascontiguousarray(offsets, np.int32)
    ms = C.c_double(0.0)
    _check(lib.cvx_test_tdnn_layer_bench(device, PRECISION[precision], rows, in_dim, _ptr(offs), len(offs), out_dim,
                                         iters, C.byref(ms)))
    return ms.value


def transform_vectors(vectors: np.ndarray, mean=None, transform=None, normalize_length: bool = True,
                      scaleup: bool = True, device: int = 0) -> Tuple[np.ndarray, np.ndarray]:
    """The back-end's per-vector chain on the device (cvx_transform_vectors):
    subtract mean -> transform-vec -> ivector-normalize-length.  Returns (vectors, ratios)."""
    lib = load_library()
    v = np.ascontiguousarray(vectors, np.float32)
    ch, keep = _make_chain(mean, transform, normalize_length, scaleup)
    od = ch.transform_rows if transform is not None else v.shape[1]
    out = np.zeros((v.shape[0], od), np.float32)
    ratios = np.zeros(v.shape[0], np.float32)
    _check(lib.cvx_transform_vectors(device, _ptr(v), v.shape[0], v.shape[1], C.byref(ch), _ptr(out), _ptr(ratios)))
    return out, ratios


def plda_score(mean, transform, psi, train, test, trial_train, trial_test, train_num_utts=None,
               normalize_length: bool = True, simple_length_norm: bool = False, device: int = 0) -> np.ndarray:
    """PLDA log-likelihood ratios of a trial list on the device (cvx_plda_score), what
    `ivector-plda-scoring [--normalize-length=true] [--num-utts=...]` writes per trial."""
    lib = load_library()
    m = PldaModel()
    mean = np.ascontiguousarray(mean, np.float64)
    tf = np.ascontiguousarray(transform, np.float64)
    psi = np.ascontiguousarray(psi, np.float64)
    m.mean, m.transform, m.psi = mean.ctypes.data, tf.ctypes.data, psi.ctypes.data
    m.dim, m.dim_in = tf.shape
    m.normalize_length, m.simple_length_norm = int(bool(normalize_length)), int(bool(simple_length_norm))
    tr = np.ascontiguousarray(train, np.float32)
    te = np.ascontiguousarray(test, np.float32)
    ti = np.ascontiguousarray(trial_train, np.int32)
    tj = np.ascontiguousarray(trial_test, np.int32)
    nu = None if train_num_utts is None else np.ascontiguousarray(train_num_utts, np.int32)
    scores = np.zeros(len(ti), np.float32)
    _check(lib.cvx_plda_score(device, C.byref(m), _ptr(tr), _ptr(nu), tr.shape[0], _ptr(te), te.shape[0], _ptr(ti), _ptr(tj),
                              len(ti), _ptr(scores)))
    return scores


def test_segment_affine(x: np.ndarray, weight: np.ndarray, bias=None, relu=False, scale=None, offset=None,
                        precision: str = "fp16", device: int = 0) -> np.ndarray:
    """One segment-level layer through the production path (cvx_test_segment_affine)."""
    lib = load_library()
    x = np.ascontiguousarray(x, np.float32)
    w = np.ascontiguousarray(weight, np.float32)
    f = lambda a: None if a is None else np.ascontiguousarray(a, np.float32)
    b, s, o = f(bias), f(scale), f(offset)
    out = np.zeros((x.shape[0], w.shape[0]), np.float32)
    lib.cvx_test_segment_affine.argtypes = [C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p,
                                            C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
    _check(lib.cvx_test_segment_affine(device, PRECISION[precision], _ptr(x), x.shape[0], x.shape[1], _ptr(w), _ptr(b),
                                       int(relu), _ptr(s), _ptr(o), w.shape[0], _ptr(out)))
    return out


def test_stats_pool(x: np.ndarray, begin: Sequence[int], end: Sequence[int], precision: str = "fp16",
                    variance_floor: float = 1e-10, device: int = 0) -> np.ndarray:
    lib = load_library()
    x = np.ascontiguousarray(x, np.float32)
    b = np.ascontiguousarray(begin, np.int64)
    e = np.ascontiguousarray(end, np.int64)
    out = np.zeros((len(b), 2 * x.shape[1]), np.float32)
    _check(lib.cvx_test_stats_pool(device, PRECISION[precision], _ptr(x), x.shape[0], x.shape[1], _ptr(b), _ptr(e),
                                   len(b), variance_floor, _ptr(out)))
    return out
