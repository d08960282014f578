"""Python side of the device feature pre-step (cvx_extract_raw / cvx_test_prestep):
sliding-window CMN + voiced-frame selection, the step right before the hot path in
egs/sre/v5/sid/nnet3/xvector/extract_xvectors_new.sh:79."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence, Tuple

import numpy as np

from . import Extractor, _check, _ptr, chunk_plan, load_library


class Prestep(C.Structure):
    _fields_ = [("cmn_window", C.c_int32), ("cmn_center", C.c_int32), ("min_cmn_window", C.c_int32),
                ("norm_vars", C.c_int32)]


def _index(lengths: Sequence[int], vad: Optional[Sequence[np.ndarray]]):
    raw_start = np.concatenate([[0], np.cumsum(lengths)]).astype(np.int64)
    if vad is None:
        return raw_start, None, raw_start.copy()
    keep = [np.flatnonzero(np.asarray(v) != 0).astype(np.int32) for v in vad]
    kept_start = np.concatenate([[0], np.cumsum([len(k) for k in keep])]).astype(np.int64)
    return raw_start, (np.concatenate(keep) if keep else np.zeros(0, np.int32)).astype(np.int32), kept_start


def run_prestep(ex: Extractor, raw: np.ndarray, lengths: Sequence[int], vad: Optional[Sequence[np.ndarray]],
                cmn_window: int = 300, center: bool = True, min_cmn_window: int = 100) -> np.ndarray:
    """The pre-step alone (test entry): returns the selected, normalised rows."""
    lib = load_library()
    lib.cvx_test_prestep.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(Prestep), C.c_int64, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
    raw = np.ascontiguousarray(raw, np.float32)
    raw_start, keep, kept_start = _index(lengths, vad)
    out = np.zeros((int(kept_start[-1]), raw.shape[1]), np.float32)
    ps = Prestep(cmn_window, int(center), min_cmn_window, 0)
    _check(lib.cvx_test_prestep(ex._h, _ptr(raw), raw.shape[1], C.byref(ps), raw.shape[0], _ptr(raw_start), _ptr(keep),
                                _ptr(kept_start), len(lengths), _ptr(out)))
    return out


def extract_raw(ex: Extractor, raw: Dict[str, np.ndarray], lengths: Sequence[int], vad: Optional[Sequence[np.ndarray]],
                cmn: Dict[str, int], chunk_size: int = -1, min_chunk_size: int = 100,
                center: bool = True) -> Tuple[np.ndarray, np.ndarray]:
    """`nnet3-xvector-compute` on RAW features with the reference's feature pipe fused in:
    per-stream sliding CMN window (`cmn[name]`, 0 = none) and VAD selection, then chunking on the
    selected frames.  Returns ([n_utts, dim], ok)."""
    lib = load_library()
    lib.cvx_extract_raw.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(Prestep), C.c_int32, C.c_int64,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_int64, C.c_void_p, C.c_uint32]
    names = ex.model.input_names
    mats = [np.ascontiguousarray(raw[n], np.float32) for n in names]
    raw_start, keep, kept_start = _index(lengths, vad)
    ptrs = (C.c_void_p * len(mats))(*[m.ctypes.data for m in mats])
    pres = (Prestep * len(mats))(*[Prestep(int(cmn.get(n, 0)), int(center), 100, 0) for n in names])
    starts, lens, utts = [], [], []
    ok = np.zeros(len(lengths), bool)
    ctx = ex.model.info.left_context + ex.model.info.right_context + 1
    for u in range(len(lengths)):
        T = int(kept_start[u + 1] - kept_start[u])
        for s, n in chunk_plan(T, chunk_size, min_chunk_size, False):
            if n >= ctx:
                starts.append(int(kept_start[u]) + s); lens.append(n); utts.append(u)
                ok[u] = True
    cs, cl, cu = np.array(starts, np.int64), np.array(lens, np.int32), np.array(utts, np.int32)
    out = np.zeros((len(lengths), ex.output_dim), np.float32)
    _check(lib.cvx_extract_raw(ex._h, ptrs, pres, len(mats), mats[0].shape[0], _ptr(raw_start), _ptr(keep),
                               _ptr(kept_start), len(lengths), _ptr(cs), _ptr(cl), _ptr(cu), None, len(cs), _ptr(out), 0))
    return out, ok
