"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY, NOT PRODUCT CODE.

numpy restatement of the reference's embedding-extraction hot path: the nnet3
forward pass + chunk averaging done by upstream Kaldi's `nnet3-xvector-compute`
as driven by egs/sre/v5/sid/nnet3/xvector/extract_xvectors_new.sh:57-94.

PARITY UNPINNED: the algorithm lives in a third-party dependency that is absent
from /root/reference (Kaldi, github.com/kaldi-asr/kaldi, version unpinned --
hints `.travis.yml:23-25` `kaldi_52`/master, `steps/libs/nnet3/train/
frame_level_objf_20180105/` => early-2018 master, Kaldi 5.3/5.4), and the
reference ships no model, no features, no expected embeddings and no tests for
this path.  What *is* pinned by the reference and
checked in tests/: the descriptor-string grammar
(steps/libs/nnet3/xconfig/utils.py:599-631), the component/node templates
(basic_layers.py:758-834, stats_layer.py:96-132), the chunk parameters and CLI
(extract_xvectors_new.sh:50-51,62-68,86-93) and the topologies.
Everything else restates Kaldi's published algorithm:

  forward_chunk        nnet3 NnetComputer over the sub-graph reachable from the
                       requested output (src/nnet3/nnet-compute.cc); a frame is
                       computed only where all its inputs exist (no padding)
  _affine              AffineComponent::Propagate   (nnet-simple-component.cc)  y = x W^T + b
  _relu                RectifiedLinearComponent::Propagate                       max(x, 0)
  _batchnorm           BatchNormComponent::ComputeDerived/Propagate, test mode
                       (nnet-normalize-component.cc)  s = rms*(max(var,0)+eps)^-1/2, o = -mean*s
  _stats               StatisticsExtractionComponent + StatisticsPoolingComponent
                       (nnet-general-component.cc)   mean, sqrt(max(E[x^2]-mean^2, floor))
  extract_utt          main() of src/nnet3bin/nnet3-xvector-compute.cc: chunking,
                       min-chunk handling (old skip / new --pad-input), frame-weighted average

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.
"""
from __future__ import annotations

import importlib
import math
import re
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

_pkg = importlib.import_module("speaker-embedding-with-phonetic-information_b200.kaldi_format")
Nnet3Model = _pkg.Nnet3Model
read_nnet3 = _pkg.read_nnet3

# --------------------------------------------------------------------------
# descriptors
# --------------------------------------------------------------------------


def parse_descriptor(s: str):
    """nnet3 descriptor string -> nested tuples:
    ('node', name) | ('Offset', d, k) | ('Append', [d..]) | ('Round', d, k)."""
    toks = re.findall(r"[(),]|[^\s(),]+", s)
    pos = 0

    def parse():
        nonlocal pos
        t = toks[pos]
        pos += 1
        if pos < len(toks) and toks[pos] == "(":
            pos += 1
            args = []
            while True:
                if re.match(r"^-?\d+$", toks[pos]) and t in ("Offset", "Round"):
                    args.append(int(toks[pos]))
                    pos += 1
                else:
                    args.append(parse())
                if toks[pos] == ",":
                    pos += 1
                elif toks[pos] == ")":
                    pos += 1
                    break
                else:
                    raise ValueError(f"bad descriptor {s!r}")
            if t == "Append":
                return ("Append", args)
            if t == "Offset":
                if len(args) != 2:
                    raise ValueError("Offset(d, t) only (no x offsets on this path)")
                return ("Offset", args[0], args[1])
            if t == "Round":
                return ("Round", args[0], args[1])
            raise ValueError(f"descriptor function {t} is not produced by the reference's recipes")
        return ("node", t)

    d = parse()
    if pos != len(toks):
        raise ValueError(f"trailing tokens in {s!r}")
    return d


def _parse_config_line(line: str) -> Tuple[str, Dict[str, str]]:
    first, rest = line.split(None, 1)
    # values may contain spaces (descriptors): split on ' key=' boundaries
    parts = re.split(r"\s+(?=[A-Za-z][\w-]*=)", rest.strip())
    kv = {}
    for p in parts:
        k, v = p.split("=", 1)
        kv[k] = v.strip()
    return first, kv


@dataclass
class Node:
    kind: str               # 'input' | 'component' | 'output'
    name: str
    dim: int = 0            # for input
    component: str = ""     # for component
    desc: object = None     # parsed input descriptor


class Graph:
    def __init__(self, model: Nnet3Model, output_node: Optional[str] = None):
        self.model = model
        self.nodes: Dict[str, Node] = {}
        for ln in model.config_lines:
            first, kv = _parse_config_line(ln)
            if first == "input-node":
                self.nodes[kv["name"]] = Node("input", kv["name"], dim=int(kv["dim"]))
            elif first == "component-node":
                self.nodes[kv["name"]] = Node("component", kv["name"], component=kv["component"],
                                              desc=parse_descriptor(kv["input"]))
            elif first == "output-node":
                self.nodes[kv["name"]] = Node("output", kv["name"], desc=parse_descriptor(kv["input"]))
            else:
                raise ValueError(f"unsupported config line: {ln}")
        if output_node is not None:
            # what `nnet3-copy --nnet-config=extract.config` does (extract_xvectors_new.sh:58-59)
            self.nodes["output"] = Node("output", "output", desc=parse_descriptor(output_node))
        self.input_names = [n.name for n in self.nodes.values() if n.kind == "input"]


# --------------------------------------------------------------------------
# forward
# --------------------------------------------------------------------------

Quant = Optional[Callable[[np.ndarray], np.ndarray]]


def round_bf16(a: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even to bfloat16, returned in the input's float type
    (emulates the operand rounding of the tensor-core path)."""
    f = np.ascontiguousarray(a, dtype=np.float32)
    u = f.view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000).astype(np.uint32)
    return r.view(np.float32).astype(a.dtype).reshape(a.shape)


def round_fp16(a: np.ndarray) -> np.ndarray:
    return np.asarray(a).astype(np.float16).astype(a.dtype)


class Oracle:
    """Graph-driven forward of one chunk.  dtype float64 is the reference mode,
    float32 restates Kaldi's BaseFloat arithmetic.  `quant` (optional) rounds
    weights and layer outputs like the GPU path does, to predict its error."""

    def __init__(self, model: Nnet3Model, output_node: Optional[str] = None, dtype=np.float64,
                 quant: Quant = None):
        self.g = Graph(model, output_node)
        self.dtype = dtype
        self.quant = quant
        self._w: Dict[str, Tuple[np.ndarray, np.ndarray]] = {}
        self._bn: Dict[str, Tuple[np.ndarray, np.ndarray]] = {}

    # -- component math ---------------------------------------------------
    def _affine(self, cname: str, x: np.ndarray, frame_level: bool) -> np.ndarray:
        if cname not in self._w:
            f = self.g.model.components[cname].fields
            w = np.asarray(f["LinearParams"], dtype=self.dtype)
            b = np.asarray(f["BiasParams"], dtype=self.dtype)
            self._w[cname] = (w, b)
        w, b = self._w[cname]
        if self.quant is not None and frame_level:
            w = self.quant(w)
        return x @ w.T + b

    def _batchnorm(self, cname: str, x: np.ndarray) -> np.ndarray:
        if cname not in self._bn:
            f = self.g.model.components[cname].fields
            dim, block = int(f["Dim"]), int(f.get("BlockDim", f["Dim"]))
            if block != dim:
                raise ValueError("BatchNorm block-dim != dim not used by these recipes")
            count = float(f["Count"])
            if count <= 0:
                raise ValueError("BatchNorm <Count>==0: upstream Kaldi fabricates random stats in test "
                                 "mode; fixtures must carry count>0 (SURVEY 8a-3)")
            eps, rms = float(f["Epsilon"]), float(f["TargetRms"])
            mean = np.asarray(f["StatsMean"], dtype=self.dtype)
            var = np.asarray(f["StatsVar"], dtype=self.dtype)
            s = (rms * (np.maximum(var, 0) + self.dtype(eps)) ** self.dtype(-0.5)).astype(self.dtype)
            self._bn[cname] = (s, (-mean * s).astype(self.dtype))
        s, o = self._bn[cname]
        return x * s + o

    # -- graph evaluation: returns (values[t_lo..t_hi], t_lo) -------------
    def forward_chunk(self, inputs: Dict[str, np.ndarray]) -> np.ndarray:
        y, lo, fl = self._evaluate(inputs)
        if fl:
            raise ValueError("requested output is frame-level; the extraction path expects one vector per chunk")
        return y[0]

    def forward_frames(self, inputs: Dict[str, np.ndarray]) -> np.ndarray:
        """Frame-level output at every input frame: what `nnet3-compute` writes for one
        utterance (R/egs/sre/v5/sid/nnet3_cvector/cvector/extract_am_embedding.sh:79-85,
        extract_log_post.sh:80-85).  Frames beyond the ends of the input are copies of the
        first / last frame -- upstream's DecodableNnetSimple edge rule [UPSTREAM-RECALL
        nnet3/nnet-am-decodable-simple.cc; not in the reference tree]."""
        left, right = self.frame_context()
        padded = {}
        for n, x in inputs.items():
            x = np.asarray(x)
            padded[n] = np.concatenate([np.repeat(x[:1], left, axis=0), x, np.repeat(x[-1:], right, axis=0)], axis=0)
        y, lo, fl = self._evaluate(padded)
        if not fl:
            raise ValueError("requested output is segment-level")
        T = next(iter(inputs.values())).shape[0]
        assert lo == left and y.shape[0] == T, (lo, left, y.shape, T)
        return y

    def _evaluate(self, inputs: Dict[str, np.ndarray]):
        T = None
        for n in self.g.input_names:
            if n in inputs:
                T = inputs[n].shape[0] if T is None else T
                if inputs[n].shape[0] != T:
                    raise ValueError("all input streams must have equal row counts")
        memo: Dict[str, Tuple[np.ndarray, int]] = {}

        def ev_node(name: str) -> Tuple[np.ndarray, int, bool]:
            """-> (values, t_lo, frame_level)."""
            if name in memo:
                return memo[name]
            nd = self.g.nodes[name]
            if nd.kind == "input":
                x = np.asarray(inputs[name], dtype=self.dtype)
                if x.shape[1] != nd.dim:
                    raise ValueError(f"input {name}: dim {x.shape[1]} != {nd.dim}")
                if self.quant is not None:
                    x = self.quant(x)
                res = (x, 0, True)
            elif nd.kind == "output":
                res = ev_desc(nd.desc)
            else:
                comp = self.g.model.components[nd.component]
                x, lo, fl = ev_desc(nd.desc)
                ct = comp.type
                if ct in ("NaturalGradientAffineComponent", "AffineComponent", "FixedAffineComponent"):
                    y = self._affine(nd.component, x, fl)
                elif ct == "RectifiedLinearComponent":
                    y = np.maximum(x, 0)
                elif ct == "BatchNormComponent":
                    y = self._batchnorm(nd.component, x)
                    if self.quant is not None and fl:
                        y = self.quant(y)
                elif ct == "LogSoftmaxComponent":
                    if not fl:
                        raise ValueError("log-softmax of a segment-level node")
                    m = x.max(axis=1, keepdims=True)
                    y = x - m - np.log(np.exp(x - m).sum(axis=1, keepdims=True))
                elif ct in ("NoOpComponent", "DropoutComponent", "GeneralDropoutComponent"):
                    y = x                                       # test mode
                elif ct == "StatisticsExtractionComponent":
                    f = comp.fields
                    if int(f["InputPeriod"]) != 1 or int(f["OutputPeriod"]) != 1:
                        raise ValueError("stats periods != 1 not used by these recipes")
                    inc_var = bool(f.get("IncludeVarinance", f.get("IncludeVariance", True)))
                    cols = [np.ones((x.shape[0], 1), dtype=self.dtype), x]
                    if inc_var:
                        cols.append(x * x)
                    y = np.concatenate(cols, axis=1)
                elif ct == "StatisticsPoolingComponent":
                    f = comp.fields
                    lc, rc = int(f["LeftContext"]), int(f["RightContext"])
                    if int(f["NumLogCountFeatures"]) != 0:
                        raise ValueError("num-log-count-features>0 not used by these recipes")
                    floor = float(f.get("VarianceFloor", 1e-10))
                    hi = lo + x.shape[0] - 1
                    a, b = max(lo, -lc), min(hi, rc)          # frames of the window [0-lc, 0+rc] that exist
                    if a > b:
                        raise ValueError("chunk too short: no computable frame reaches the pooling layer")
                    seg = x[a - lo:b - lo + 1]
                    tot = seg.sum(axis=0, dtype=self.dtype)
                    cnt = tot[0]
                    d = (x.shape[1] - 1) // 2 if bool(f["OutputStddevs"]) else x.shape[1] - 1
                    mean = tot[1:1 + d] / cnt
                    if bool(f["OutputStddevs"]):
                        var = tot[1 + d:1 + 2 * d] / cnt - mean * mean
                        y = np.concatenate([mean, np.sqrt(np.maximum(var, self.dtype(floor)))])[None, :]
                    else:
                        y = mean[None, :]
                    y = y.astype(self.dtype)
                    lo, fl = 0, False
                else:
                    raise ValueError(f"component type {ct} not on the extraction path")
                res = (y.astype(self.dtype, copy=False), lo, fl)
            memo[name] = res
            return res

        def ev_desc(d) -> Tuple[np.ndarray, int, bool]:
            if d[0] == "node":
                return ev_node(d[1])
            if d[0] == "Offset":
                x, lo, fl = ev_desc(d[1])
                if not fl:
                    raise ValueError("Offset of a segment-level node")
                return x, lo - d[2], fl          # value at t is the source at t+k
            if d[0] == "Round":
                x, lo, fl = ev_desc(d[1])
                return x, lo, fl
            if d[0] == "Append":
                parts = [ev_desc(a) for a in d[1]]
                fls = {p[2] for p in parts}
                if len(fls) != 1:
                    raise ValueError("Append mixes frame-level and segment-level inputs")
                lo = max(p[1] for p in parts)
                hi = min(p[1] + p[0].shape[0] - 1 for p in parts)
                if hi < lo:
                    raise ValueError("chunk too short for the network's context")
                return (np.concatenate([p[0][lo - p[1]:hi - p[1] + 1] for p in parts], axis=1), lo,
                        parts[0][2])
            raise ValueError(d)

        return ev_node("output")

    # -- context ----------------------------------------------------------
    def frame_context(self) -> Tuple[int, int]:
        """(left, right) context of a frame-level output node."""
        def walk(d) -> Tuple[int, int]:
            if d[0] == "node":
                nd = self.g.nodes[d[1]]
                return (0, 0) if nd.kind == "input" else walk(nd.desc)
            if d[0] == "Offset":
                l, r = walk(d[1])
                return l - d[2], r + d[2]
            if d[0] == "Round":
                return walk(d[1])
            ls, rs = zip(*[walk(a) for a in d[1]])
            return max(ls), max(rs)
        l, r = walk(self.g.nodes["output"].desc)
        return max(l, 0), max(r, 0)

    def context(self) -> Tuple[int, int]:
        """(left, right) context of the pooling input, by probing (N_valid = T-L-R)."""
        def walk(d) -> Tuple[int, int]:
            if d[0] == "node":
                nd = self.g.nodes[d[1]]
                if nd.kind == "input":
                    return 0, 0
                return walk(nd.desc)
            if d[0] == "Offset":
                l, r = walk(d[1])
                return l - d[2], r + d[2]
            if d[0] == "Round":
                return walk(d[1])
            ls, rs = zip(*[walk(a) for a in d[1]])
            return max(ls), max(rs)

        def find_pool(d):
            if d[0] == "node":
                nd = self.g.nodes[d[1]]
                if nd.kind == "component" and \
                        self.g.model.components[nd.component].type == "StatisticsPoolingComponent":
                    return nd
                if nd.kind == "input":
                    return None
                return find_pool(nd.desc)
            if d[0] in ("Offset", "Round"):
                return find_pool(d[1])
            for a in d[1]:
                r = find_pool(a)
                if r is not None:
                    return r
            return None

        pool = find_pool(self.g.nodes["output"].desc)
        return walk(pool.desc)


# --------------------------------------------------------------------------
# chunking / averaging  (nnet3-xvector-compute main())
# --------------------------------------------------------------------------


def chunk_plan(T: int, chunk_size: int, min_chunk_size: int, pad_input: bool) -> List[Tuple[int, int]]:
    """[(start_row, num_rows)] of the chunks that get an embedding; [] if the
    utterance fails.  chunk_size == -1 means "whole utterance"."""
    if T == 0:
        return []
    if not pad_input and T < min_chunk_size:
        return []
    c = chunk_size if (chunk_size != -1 and T >= chunk_size) else T
    out = []
    for i in range(int(math.ceil(T / c))):
        off = min(c, T - i * c)
        if off < min_chunk_size and not pad_input:
            continue
        out.append((i * c, off))
    return out


def pad_chunk(x: np.ndarray, min_chunk_size: int) -> np.ndarray:
    """Newer-Kaldi `--pad-input=true`: replicate first/last row up to min size."""
    n = x.shape[0]
    if n >= min_chunk_size:
        return x
    left = (min_chunk_size - n) // 2
    right = min_chunk_size - n - left
    return np.concatenate([np.repeat(x[:1], left, 0), x, np.repeat(x[-1:], right, 0)], axis=0)


def extract_utt(oracle: Oracle, feats: Dict[str, np.ndarray], chunk_size: int = 10000,
                min_chunk_size: int = 25, pad_input: bool = False) -> Optional[np.ndarray]:
    T = next(iter(feats.values())).shape[0]
    plan = chunk_plan(T, chunk_size, min_chunk_size, pad_input)
    if not plan:
        return None
    acc, tot = None, 0.0
    for start, n in plan:
        xi = {k: v[start:start + n] for k, v in feats.items()}
        if pad_input:
            xi = {k: pad_chunk(v, min_chunk_size) for k, v in xi.items()}
        e = oracle.forward_chunk(xi)
        acc = n * e if acc is None else acc + n * e
        tot += n
    return (acc / tot).astype(oracle.dtype)


def extract_all(oracle: Oracle, feats: Dict[str, np.ndarray], lengths: Sequence[int], **kw):
    """feats: packed rows per stream; returns ([n_utts, dim] array, ok mask)."""
    outs, ok, r = [], [], 0
    dim = None
    for n in lengths:
        e = extract_utt(oracle, {k: v[r:r + n] for k, v in feats.items()}, **kw)
        r += n
        ok.append(e is not None)
        outs.append(e)
        if e is not None:
            dim = e.shape[0]
    arr = np.zeros((len(outs), dim or 0), dtype=oracle.dtype)
    for i, e in enumerate(outs):
        if e is not None:
            arr[i] = e
    return arr, np.array(ok)


def parity_metrics(a: np.ndarray, b: np.ndarray) -> Tuple[float, float]:
    """(max over vectors of max|a-b|/max|b|, min cosine) -- the north-star tolerance
    definition (SURVEY section 7 'Precision vs tolerance')."""
    a = np.asarray(a, dtype=np.float64).reshape(-1, a.shape[-1])
    b = np.asarray(b, dtype=np.float64).reshape(-1, b.shape[-1])
    rel = np.max(np.abs(a - b), axis=1) / np.max(np.abs(b), axis=1)
    cos = np.sum(a * b, axis=1) / (np.linalg.norm(a, axis=1) * np.linalg.norm(b, axis=1))
    return float(rel.max()), float(cos.min())
