"""CPU ORACLE for the feature pre-step -- TEST INFRASTRUCTURE ONLY.

Restates the two upstream Kaldi programs the reference pipes features through before
the hot path (egs/sre/v5/sid/nnet3/xvector/extract_xvectors_new.sh:79), neither shipped
by the reference (PARITY UNPINNED, as for oracle/nnet3_oracle.py):

  sliding_cmn      apply-cmvn-sliding = SlidingWindowCmn (src/feat/feature-functions.cc):
                   per frame t a window [ws, we) -- centred: ws = t - W/2, we = ws + W; else
                   ws = t - W, we = t + 1 -- shifted right if ws < 0, (non-centred: grown to
                   min-cmn-window), shifted left if we > T; out[t] = x[t] - mean(x[ws:we]),
                   sums in double.
  select_voiced    select-voiced-frames: rows whose VAD value is non-zero, in order -- the
                   same rule the reference's own src/ivectorbin/select-voiced-ali.cc:58-91
                   applies to alignments; an utterance with no voiced frame is skipped.
"""
from __future__ import annotations

from typing import Optional

import numpy as np


def sliding_cmn(x: np.ndarray, cmn_window: int = 300, center: bool = True, min_cmn_window: int = 100) -> np.ndarray:
    T = x.shape[0]
    xd = x.astype(np.float64)
    out = np.empty_like(xd)
    for t in range(T):                      # plain loop on purpose: this is the checker
        if center:
            ws = t - cmn_window // 2
            we = ws + cmn_window
        else:
            ws = t - cmn_window
            we = t + 1
        if ws < 0:
            we -= ws
            ws = 0
        if not center and we > t:
            we = max(t + 1, min_cmn_window)
        if we > T:
            ws -= we - T
            we = T
            ws = max(ws, 0)
        out[t] = xd[t] - xd[ws:we].mean(axis=0)
    return out.astype(np.float32)


def select_voiced(x: np.ndarray, vad: Optional[np.ndarray]) -> np.ndarray:
    if vad is None:
        return x
    if len(vad) != x.shape[0]:
        raise ValueError("VAD length does not match the number of frames")
    return x[np.asarray(vad) != 0]
