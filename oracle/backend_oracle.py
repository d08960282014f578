"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference back-end's per-vector post-step
(SURVEY.md 8f-2).  Only tests/ may import this; the product path never does.

PARITY UNPINNED: the reference ships none of these programs (they are upstream Kaldi binaries it
calls from R/egs/sre/v5/run_sre10.sh:107-129 and R/egs/sre/v5/sid/nnet3/xvector/
extract_xvectors_new.sh:102-108); the rules below restate upstream from memory
[UPSTREAM-RECALL ivectorbin/ivector-mean.cc, ivectorbin/ivector-subtract-global-mean.cc,
bin/transform-vec.cc, ivectorbin/ivector-normalize-length.cc] and no golden vectors exist.
"""
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np


def subtract_global_mean(v: np.ndarray, mean: Optional[np.ndarray] = None) -> np.ndarray:
    """ivector-subtract-global-mean: two-argument form subtracts the mean of the table itself,
    three-argument form a mean read from a file (run_sre10.sh:118,127-128)."""
    v = np.asarray(v, np.float64)
    m = v.mean(axis=0) if mean is None else np.asarray(mean, np.float64)
    return v - m


def transform_vec(v: np.ndarray, t: np.ndarray) -> np.ndarray:
    """transform-vec: y = T v, or T [v; 1] when T has dim + 1 columns (run_sre10.sh:123,127)."""
    v, t = np.asarray(v, np.float64), np.asarray(t, np.float64)
    if t.shape[1] == v.shape[1]:
        return v @ t.T
    if t.shape[1] == v.shape[1] + 1:
        return v @ t[:, :-1].T + t[:, -1]
    raise ValueError(f"transform has {t.shape[1]} columns, vectors have dim {v.shape[1]}")


def normalize_length(v: np.ndarray, normalize: bool = True, scaleup: bool = True) -> Tuple[np.ndarray, np.ndarray]:
    """ivector-normalize-length: ratio = ||v|| / sqrt(dim) (scaleup) or ||v||; v /= ratio unless it
    is zero.  Returns (vectors, ratios)."""
    v = np.asarray(v, np.float64)
    norm = np.sqrt((v * v).sum(axis=1))
    ratio = norm / np.sqrt(v.shape[1]) if scaleup else norm
    out = v.copy()
    nz = ratio != 0
    if normalize:
        out[nz] = v[nz] / ratio[nz, None]
    return out, ratio


def vector_chain(v, mean=None, transform=None, normalize: bool = True, scaleup: bool = True):
    y = np.asarray(v, np.float64)
    if mean is not None:
        y = y - np.asarray(mean, np.float64)
    if transform is not None:
        y = transform_vec(y, transform)
    return normalize_length(y, normalize, scaleup)


def speaker_means(spk2utt: Sequence[Tuple[str, Sequence[str]]], vectors: Dict[str, np.ndarray]):
    """ivector-mean <spk2utt> <ivectors> <spk-ivectors> <num-utts> (extract_xvectors_new.sh:102-108):
    per speaker the plain average of the utterance vectors that exist; speakers with none are
    skipped.  -> ([(spk, mean)], [(spk, n)], n_missing_utts)."""
    out, counts, missing = [], [], 0
    for spk, utts in spk2utt:
        have = [np.asarray(vectors[u], np.float64) for u in utts if u in vectors]
        missing += len(utts) - len(have)
        if not have:
            continue
        out.append((spk, np.mean(have, axis=0)))
        counts.append((spk, len(have)))
    return out, counts, missing


def plda_transform(x, mean, transform, psi, num_examples=1, normalize_length=True, simple_length_norm=False):
    """Plda::TransformIvector + GetNormalizationFactor [UPSTREAM-RECALL ivector/plda.cc]:
    y = transform (x - mean); y *= sqrt(dim / sum y^2 / (psi + 1/n)) (or sqrt(dim) / ||y||)."""
    y = (np.asarray(x, np.float64) - np.asarray(mean, np.float64)) @ np.asarray(transform, np.float64).T
    n = np.broadcast_to(np.asarray(num_examples, np.float64), y.shape[:1])
    if simple_length_norm:
        f = np.sqrt(y.shape[1]) / np.sqrt((y * y).sum(axis=1))
    else:
        f = np.sqrt(y.shape[1] / (y * y / (np.asarray(psi, np.float64) + 1.0 / n[:, None])).sum(axis=1))
    return y * f[:, None] if normalize_length else y


def plda_llr(train, n, test, psi):
    """Plda::LogLikelihoodRatio for one (transformed) enrolment vector averaged over n utterances and
    one transformed test vector."""
    psi = np.asarray(psi, np.float64)
    mean = n * psi / (n * psi + 1.0) * train
    var = 1.0 + psi / (n * psi + 1.0)
    given = -0.5 * (np.log(var).sum() + np.log(2 * np.pi) * len(psi) + ((test - mean) ** 2 / var).sum())
    var0 = psi + 1.0
    without = -0.5 * (np.log(var0).sum() + np.log(2 * np.pi) * len(psi) + (test ** 2 / var0).sum())
    return given - without
