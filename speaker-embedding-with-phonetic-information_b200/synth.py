"""Synthetic workload generator: random-init nnet3 models of the reference's
topologies and synthetic MFCC-shaped features.

The reference ships no trained model and no data, so every
test / bench input is produced here with fixed seeds.  Layer stacks are copied
*as numbers* from the reference's xconfig heredocs:

  v2  x-vector     egs/sre/v2/local/nnet3/xvector/run_xvector_new.sh:94-114
  v3  multitask    egs/sre/v3/local/nnet3_cvector/cvector/prepare_nnet3_xconfig{,_2share,_3share,_4share}.sh:46-70
  v4  AM           egs/sre/v4/local/nnet3_cvector/cvector/train_am.sh:30-38
  v4  c-vector     egs/sre/v4/local/nnet3_cvector/cvector/train_xvector_with_am.sh:44-57
  v5  c-vector     egs/sre/v5/local/nnet3_cvector/cvector/train_cvector_with_am.sh:65-89
  v5  "expand"     egs/sre/v5/sid/nnet3_cvector/cvector/train_cvector_with_am_expand.sh:73-96

and turned into nnet3 node/component lines the way
egs/sre/v5/steps/libs/nnet3/xconfig/basic_layers.py:758-834 and
stats_layer.py:96-132 do (that library is Python-2 only and cannot be imported
here).  Files are written in genuine Kaldi formats so a Kaldi
owner can replay them through the real `nnet3-xvector-compute`.
"""
from __future__ import annotations

import re
from collections import OrderedDict
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .kaldi_format import Component, Nnet3Model

# --------------------------------------------------------------------------
# xconfig descriptor strings -> nnet3 descriptor strings
# (grammar pinned by egs/sre/v5/steps/libs/nnet3/xconfig/utils.py:599-631)
# --------------------------------------------------------------------------

_TOKEN_RE = re.compile(r"\s*([(),@]|[^\s(),@]+)")


def tokenize_descriptor(s: str, prev_names: Optional[Sequence[str]] = None,
                        layer_outputs: Optional[Dict[str, str]] = None) -> List[str]:
    toks = _TOKEN_RE.findall(s)
    out = []
    for t in toks:
        m = re.match(r"^\[(-?\d+)\](.*)$", t)
        if m:
            if not prev_names:
                raise ValueError("relative layer reference without previous layers")
            out.append(prev_names[int(m.group(1))] + m.group(2))
        elif layer_outputs and t in layer_outputs:
            out.append(layer_outputs[t])       # xconfig layer name -> its output node
        else:
            out.append(t)
    return out


def normalize_descriptor(s: str, prev_names: Optional[Sequence[str]] = None,
                         layer_outputs: Optional[Dict[str, str]] = None) -> str:
    """`Append(-3,0,3)` -> `Append(Offset(prev, -3), prev, Offset(prev, 3))`,
    `input@-3` -> `Offset(input, -3)` etc."""
    toks = tokenize_descriptor(s, prev_names, layer_outputs)
    pos = 0

    def is_int(t: str) -> bool:
        return re.match(r"^-?\d+$", t) is not None

    def parse() -> str:
        nonlocal pos
        t = toks[pos]
        pos += 1
        if pos < len(toks) and toks[pos] == "(":
            pos += 1
            args = []
            while True:
                args.append(parse_arg(t))
                if toks[pos] == ",":
                    pos += 1
                    continue
                if toks[pos] == ")":
                    pos += 1
                    break
                raise ValueError(f"bad descriptor {s!r}")
            d = f"{t}({', '.join(args)})"
        elif is_int(t):
            if not prev_names:
                raise ValueError("bare integer offset without previous layer")
            d = prev_names[-1] if int(t) == 0 else f"Offset({prev_names[-1]}, {int(t)})"
        else:
            d = t
        if pos < len(toks) and toks[pos] == "@":
            off = int(toks[pos + 1])
            pos += 2
            if off != 0:
                d = f"Offset({d}, {off})"
        return d

    def parse_arg(fn: str) -> str:
        nonlocal pos
        # integer arguments of Offset/Round/ReplaceIndex stay literal
        if fn in ("Offset", "Round", "ReplaceIndex") and (is_int(toks[pos]) or toks[pos] in ("t", "x")) \
                and toks[pos + 1] in (",", ")"):
            v = toks[pos]
            pos += 1
            return v
        return parse()

    d = parse()
    if pos != len(toks):
        raise ValueError(f"trailing tokens in descriptor {s!r}")
    return d


# --------------------------------------------------------------------------
# mini xconfig -> model
# --------------------------------------------------------------------------


class _Builder:
    def __init__(self, rng: np.random.Generator, bias_std: float = 0.1):
        self.rng = rng
        self.bias_std = bias_std
        self.lines: List[str] = []
        self.comps: "OrderedDict[str, Component]" = OrderedDict()
        self.dims: Dict[str, int] = {}          # node name -> dim
        self.prev: List[str] = []               # xconfig layer output names
        self.outs: Dict[str, str] = {}          # xconfig layer name -> output node

    # -- helpers ---------------------------------------------------------
    def _desc_dim(self, desc: str) -> int:
        d = desc.strip()
        m = re.match(r"^(\w+)\((.*)\)$", d)
        if not m:
            return self.dims[d]
        fn, body = m.group(1), m.group(2)
        args, depth, cur = [], 0, ""
        for ch in body:
            if ch == "(":
                depth += 1
            if ch == ")":
                depth -= 1
            if ch == "," and depth == 0:
                args.append(cur)
                cur = ""
            else:
                cur += ch
        args.append(cur)
        if fn == "Append":
            return sum(self._desc_dim(a) for a in args)
        if fn in ("Offset", "Round", "ReplaceIndex", "IfDefined"):
            return self._desc_dim(args[0])
        if fn in ("Sum", "Failover", "Switch"):
            return self._desc_dim(args[0])
        raise ValueError(f"unknown descriptor function {fn}")

    def _affine(self, name: str, in_dim: int, out_dim: int, zero: bool = False) -> None:
        w = np.zeros((out_dim, in_dim), np.float32) if zero else \
            (self.rng.standard_normal((out_dim, in_dim)) / np.sqrt(in_dim)).astype(np.float32)
        b = np.zeros(out_dim, np.float32) if zero else \
            (self.rng.standard_normal(out_dim) * self.bias_std).astype(np.float32)
        self.comps[name] = Component("NaturalGradientAffineComponent", OrderedDict([
            ("MaxChange", 0.75), ("LearningRate", 0.001),
            ("LinearParams", w), ("BiasParams", b),
            ("RankIn", 20), ("RankOut", 80), ("UpdatePeriod", 4),
            ("NumSamplesHistory", 2000.0), ("Alpha", 4.0)]))

    # -- layers ----------------------------------------------------------
    def input(self, name: str, dim: int) -> None:
        self.lines.append(f"input-node name={name} dim={dim}")
        self.dims[name] = dim
        self.prev.append(name)

    def relu_batchnorm_layer(self, name: str, dim: int, input: str = "[-1]") -> None:
        desc = normalize_descriptor(input, self.prev, self.outs)
        in_dim = self._desc_dim(desc)
        self._affine(f"{name}.affine", in_dim, dim)
        self.lines.append(f"component-node name={name}.affine component={name}.affine input={desc}")
        self.comps[f"{name}.relu"] = Component("RectifiedLinearComponent", OrderedDict([
            ("Dim", dim),
            ("ValueAvg", np.zeros(0, np.float32)), ("DerivAvg", np.zeros(0, np.float32)),
            ("Count", 0.0), ("NumDimsSelfRepaired", 0.0), ("NumDimsProcessed", 0.0),
            ("SelfRepairScale", 1e-5)]))
        self.lines.append(f"component-node name={name}.relu component={name}.relu input={name}.affine")
        mean = (0.3 + 0.1 * self.rng.standard_normal(dim)).astype(np.float32)
        var = self.rng.uniform(0.05, 0.5, dim).astype(np.float32)
        self.comps[f"{name}.batchnorm"] = Component("BatchNormComponent", OrderedDict([
            ("Dim", dim), ("BlockDim", dim), ("Epsilon", 1e-3), ("TargetRms", 1.0),
            ("TestMode", False), ("Count", 1e5), ("StatsMean", mean), ("StatsVar", var)]))
        self.lines.append(f"component-node name={name}.batchnorm component={name}.batchnorm input={name}.relu")
        for suf in ("affine", "relu", "batchnorm"):
            self.dims[f"{name}.{suf}"] = dim
        self.prev.append(f"{name}.batchnorm")
        self.outs[name] = f"{name}.batchnorm"

    def stats_layer(self, name: str, left: int, right: int, input: str = "[-1]") -> None:
        desc = normalize_descriptor(input, self.prev, self.outs)
        d = self._desc_dim(desc)
        ex, po = f"{name}-extraction-{left}-{right}", f"{name}-pooling-{left}-{right}"
        self.comps[ex] = Component("StatisticsExtractionComponent", OrderedDict([
            ("InputDim", d), ("InputPeriod", 1), ("OutputPeriod", 1), ("IncludeVarinance", True)]))
        self.lines.append(f"component-node name={ex} component={ex} input={desc}")
        self.comps[po] = Component("StatisticsPoolingComponent", OrderedDict([
            ("InputDim", 2 * d + 1), ("InputPeriod", 1), ("LeftContext", left), ("RightContext", right),
            ("NumLogCountFeatures", 0), ("OutputStddevs", True), ("VarianceFloor", 1e-10)]))
        self.lines.append(f"component-node name={po} component={po} input={ex}")
        self.dims[ex] = 2 * d + 1
        self.dims[po] = 2 * d
        out = f"Round({po}, 1)"
        self.dims[out] = 2 * d
        self.prev.append(out)
        self.outs[name] = out

    def output_layer(self, name: str, dim: int, input: str = "[-1]") -> None:
        desc = normalize_descriptor(input, self.prev, self.outs)
        in_dim = self._desc_dim(desc)
        # xconfig initialises output layers with param-stddev=0 bias-stddev=0; random_outputs gives
        # trained-looking heads for the frame-level (nnet3-compute) tests
        self._affine(f"{name}.affine", in_dim, dim, zero=not getattr(self, "random_outputs", False))
        self.lines.append(f"component-node name={name}.affine component={name}.affine input={desc}")
        self.comps[f"{name}.log-softmax"] = Component("LogSoftmaxComponent", OrderedDict([
            ("Dim", dim), ("ValueAvg", np.zeros(0, np.float32)), ("DerivAvg", np.zeros(0, np.float32)),
            ("Count", 0.0), ("NumDimsSelfRepaired", 0.0), ("NumDimsProcessed", 0.0)]))
        self.lines.append(f"component-node name={name}.log-softmax component={name}.log-softmax "
                          f"input={name}.affine")
        self.lines.append(f"output-node name={name} input={name}.log-softmax objective=linear")
        self.dims[f"{name}.affine"] = dim

    def model(self) -> Nnet3Model:
        return Nnet3Model(list(self.lines), OrderedDict(self.comps))


def _am_stack(b: _Builder, prefix: str = "tdnn", suffix: str = "", inp: str = "input",
              hidden: int = 650, bottleneck: int = 128) -> str:
    """v4 train_am.sh:31-36 (names tdnn1..tdnn5); returns the bottleneck node."""
    n = lambda i: f"{prefix}{i}{suffix}"
    b.relu_batchnorm_layer(n(1), hidden, f"Append({inp}@-2,{inp}@-1,{inp}@0,{inp}@1,{inp}@2)")
    b.relu_batchnorm_layer(n(2), hidden, "Append(-1,0,1)")
    b.relu_batchnorm_layer(n(3), hidden, "Append(-1,0,1)")
    b.relu_batchnorm_layer(n(4), hidden, "Append(-3,0,3)")
    b.relu_batchnorm_layer(n(5), bottleneck, "Append(-6,-3,0)")
    return f"{n(5)}.batchnorm"


TOPOLOGIES = ("v2_xvector", "v3_multitask", "v3_multitask_2share", "v3_multitask_3share",
              "v3_multitask_4share", "v4_cvector", "v4_two_input", "v4_pa_wo_pretrain", "v5_cvector",
              "v5_expand")

#: node that extract_xvectors_new.sh redirects `output` to, per topology
EMBEDDING_NODE = {
    "v2_xvector": "tdnn6.affine",                 # v2/run_sre16.sh:59
    "v3_multitask": "tdnn6_xvec.affine",          # v3/run_sre10.sh:127
    "v3_multitask_2share": "tdnn6_xvec.affine",
    "v3_multitask_3share": "tdnn6_xvec.affine",
    "v3_multitask_4share": "tdnn6_xvec.affine",
    "v4_cvector": "tdnn6_xvec.affine",            # v4/run_sre10.sh:81
    "v4_two_input": "tdnn6_xvec.affine",
    "v4_pa_wo_pretrain": "tdnn6.affine",          # v4/run_sre10_pa_wo_pretrain.sh:199-201
    "v5_cvector": "tdnn6_xvec.affine",            # v5/run_sre10.sh:83
    "v5_expand": "tdnn6_xvec.affine",
}

#: frame-level multiply-accumulates per input frame of the extracted sub-graph
MAC_PER_FRAME = {
    "v2_xvector": 2_661_888, "v3_multitask": 2_661_888, "v3_multitask_2share": 2_661_888,
    "v3_multitask_3share": 2_661_888, "v3_multitask_4share": 2_661_888,
    "v4_cvector": 6_980_738, "v4_pa_wo_pretrain": 6_980_738, "v5_cvector": 6_980_738,
    "v4_two_input": 7_035_988,
    # AM stack 4 126 850 + tdnn1_share 755*512 + tdnn2_share/tdnn3_xvec 2*1536*512 + tdnn4/5_xvec 2*512*512
    "v5_expand": 6_610_562,
}
SEGMENT_MAC_PER_CHUNK = 3000 * 512


def make_model(topology: str, seed: int = 42, feat_dim: int = 23, am_feat_dim: int = 40,
               num_speakers: int = 64, num_senones: int = 96, max_chunk: int = 10000,
               hidden: int = 512, stats_dim: int = 1500, am_hidden: int = 650, am_bottleneck: int = 128,
               random_outputs: bool = False) -> Nnet3Model:
    """Random-init model file content for one of TOPOLOGIES (W~N(0,1/K),
    b~N(0,0.1^2), BN mean~N(0.3,0.1^2), var~U(0.05,0.5), count=1e5, eps=1e-3).
    hidden / stats_dim / am_* default to the reference's sizes; smaller values give
    same-shaped nets for small committed fixtures."""
    rng = np.random.default_rng(seed)
    b = _Builder(rng)
    b.random_outputs = random_outputs
    H, S = hidden, stats_dim
    am = dict(hidden=am_hidden, bottleneck=am_bottleneck)
    if topology == "v2_xvector":
        b.input("input", feat_dim)
        b.relu_batchnorm_layer("tdnn1", H, "Append(-2,-1,0,1,2)")
        b.relu_batchnorm_layer("tdnn2", H, "Append(-2,0,2)")
        b.relu_batchnorm_layer("tdnn3", H, "Append(-3,0,3)")
        b.relu_batchnorm_layer("tdnn4", H)
        b.relu_batchnorm_layer("tdnn5", S)
        b.stats_layer("stats", 0, max_chunk)
        b.relu_batchnorm_layer("tdnn6", H)
        b.relu_batchnorm_layer("tdnn7", H)
        b.output_layer("output", num_speakers)
    elif topology.startswith("v3_multitask"):
        nshare = {"v3_multitask": 1, "v3_multitask_2share": 2, "v3_multitask_3share": 3,
                  "v3_multitask_4share": 4}[topology]
        splice = {1: "Append({p}@-2,{p}@-1,{p}@0,{p}@1,{p}@2)", 2: "Append({p}@-2,{p}@0,{p}@2)",
                  3: "Append({p}@-3,{p}@0,{p}@3)", 4: "{p}", 5: "{p}"}
        b.input("input", feat_dim)
        last = "input"
        for i in range(1, nshare + 1):
            b.relu_batchnorm_layer(f"tdnn{i}", H, splice[i].format(p=last))
            last = f"tdnn{i}"
        shared = last
        for branch, nlayers in (("am", 7), ("xvec", 5)):
            last = shared
            for i in range(nshare + 1, nlayers + 1):
                dim = S if (branch == "xvec" and i == 5) else H
                sp = splice[i] if i <= 5 else "{p}"
                b.relu_batchnorm_layer(f"tdnn{i}_{branch}", dim, sp.format(p=last))
                last = f"tdnn{i}_{branch}"
            if branch == "am":
                b.output_layer("output_am", num_senones, last)
        b.stats_layer("stats_xvec", 0, max_chunk, "tdnn5_xvec")
        b.relu_batchnorm_layer("tdnn6_xvec", H)
        b.relu_batchnorm_layer("tdnn7_xvec", H, "tdnn6_xvec")
        b.output_layer("output", num_speakers, "tdnn7_xvec")
    elif topology in ("v4_cvector", "v4_two_input"):
        b.input("input", feat_dim)
        am_in = "input"
        if topology == "v4_two_input":
            b.input("input_am", am_feat_dim)
            am_in = "input_am"
        am_node = _am_stack(b, inp=am_in, **am)
        b.output_layer("output_am", num_senones, "tdnn5")       # the pretrained AM's own head (orphan)
        b.relu_batchnorm_layer("tdnn1_xvec", H, "Append(input@-2,input@-1,input@0,input@1,input@2)")
        b.relu_batchnorm_layer("tdnn2_xvec", H, "Append(-2,0,2)")
        b.relu_batchnorm_layer("tdnn3_xvec", H, "Append(-3,0,3)")
        b.relu_batchnorm_layer("tdnn4_xvec", H)
        b.relu_batchnorm_layer("tdnn5_xvec", S, f"Append(tdnn4_xvec,{am_node})")
        b.stats_layer("stats", 0, max_chunk)
        b.relu_batchnorm_layer("tdnn6_xvec", H)
        b.relu_batchnorm_layer("tdnn7_xvec", H)
        b.output_layer("output", num_speakers)
    elif topology == "v4_pa_wo_pretrain":
        b.input("input", feat_dim)
        am_node = _am_stack(b, suffix="_am", **am)
        b.relu_batchnorm_layer("tdnn1", H, "Append(input@-2,input@-1,input@0,input@1,input@2)")
        b.relu_batchnorm_layer("tdnn2", H, "Append(-2,0,2)")
        b.relu_batchnorm_layer("tdnn3", H, "Append(-3,0,3)")
        b.relu_batchnorm_layer("tdnn4", H)
        b.relu_batchnorm_layer("tdnn5", S, f"Append(tdnn4,{am_node})")
        b.stats_layer("stats", 0, max_chunk)
        b.relu_batchnorm_layer("tdnn6", H)
        b.relu_batchnorm_layer("tdnn7", H)
        b.output_layer("output", num_speakers)
    elif topology == "v5_cvector":
        b.input("input", feat_dim)
        am_node = _am_stack(b, **am)                                 # pretrained AM, names tdnn1..tdnn5
        b.relu_batchnorm_layer("tdnn1_share", H, "Append(input@-2,input@-1,input@0,input@1,input@2)")
        last = "tdnn1_share"
        for i, sp in ((2, "Append({p}@-2,{p}@0,{p}@2)"), (3, "Append({p}@-3,{p}@0,{p}@3)"),
                      (4, "{p}"), (5, "{p}"), (6, "{p}"), (7, "{p}")):
            b.relu_batchnorm_layer(f"tdnn{i}_am", H, sp.format(p=last))
            last = f"tdnn{i}_am"
        b.output_layer("output_am", num_senones, "tdnn7_am")
        b.relu_batchnorm_layer("tdnn2_xvec", H, "Append(tdnn1_share@-2,tdnn1_share@0,tdnn1_share@2)")
        b.relu_batchnorm_layer("tdnn3_xvec", H, "Append(tdnn2_xvec@-3,tdnn2_xvec@0,tdnn2_xvec@3)")
        b.relu_batchnorm_layer("tdnn4_xvec", H, "tdnn3_xvec")
        b.relu_batchnorm_layer("tdnn5_xvec", S, f"Append(tdnn4_xvec,{am_node})")
        b.stats_layer("stats_xvec", 0, max_chunk, "tdnn5_xvec")
        b.relu_batchnorm_layer("tdnn6_xvec", H)
        b.relu_batchnorm_layer("tdnn7_xvec", H, "tdnn6_xvec")
        b.output_layer("output", num_speakers, "tdnn7_xvec")
    elif topology == "v5_expand":
        # train_cvector_with_am_expand.sh:73-96: the phonetic bottleneck is spliced into the FIRST
        # speaker layer, tap by tap, next to the raw features; the pooled layer is 512 wide
        b.input("input", feat_dim)
        am_node = _am_stack(b, **am)
        b.relu_batchnorm_layer("tdnn1_share", H, "Append(" + ",".join(
            f"input@{k},{am_node}@{k}" for k in (-2, -1, 0, 1, 2)) + ")")
        b.relu_batchnorm_layer("tdnn2_share", H, "Append(tdnn1_share@-2,tdnn1_share@0,tdnn1_share@2)")
        last = "tdnn2_share"
        for i, sp in ((3, "Append({p}@-3,{p}@0,{p}@3)"), (4, "{p}"), (5, "{p}"), (6, "{p}"), (7, "{p}")):
            b.relu_batchnorm_layer(f"tdnn{i}_am", H, sp.format(p=last))
            last = f"tdnn{i}_am"
        b.output_layer("output_am", num_senones, "tdnn7_am")
        b.relu_batchnorm_layer("tdnn3_xvec", H, "Append(tdnn2_share@-3,tdnn2_share@0,tdnn2_share@3)")
        b.relu_batchnorm_layer("tdnn4_xvec", H, "tdnn3_xvec")
        b.relu_batchnorm_layer("tdnn5_xvec", H, "tdnn4_xvec")
        b.stats_layer("stats_xvec", 0, max_chunk, "tdnn5_xvec")
        b.relu_batchnorm_layer("tdnn6_xvec", H)
        b.relu_batchnorm_layer("tdnn7_xvec", H, "tdnn6_xvec")
        b.output_layer("output", num_speakers, "tdnn7_xvec")
    else:
        raise ValueError(f"unknown topology {topology!r}; choose from {TOPOLOGIES}")
    return b.model()


def make_tiny_model(seed: int = 7, feat_dim: int = 5, hidden: int = 16, pooled: int = 24,
                    emb: int = 8, two_branch: bool = False) -> Nnet3Model:
    """Small net with the same operator set (for pure-Python-loop checks)."""
    rng = np.random.default_rng(seed)
    b = _Builder(rng)
    b.input("input", feat_dim)
    if two_branch:
        b.relu_batchnorm_layer("am1", 12, "Append(input@-1,input@0,input@1)")
        b.relu_batchnorm_layer("am2", 6, "Append(-4,-2,0)")
    b.relu_batchnorm_layer("tdnn1", hidden, "Append(input@-2,input@-1,input@0,input@1,input@2)")
    b.relu_batchnorm_layer("tdnn2", hidden, "Append(-2,0,2)")
    b.relu_batchnorm_layer("tdnn3", pooled, "Append(tdnn2,am2.batchnorm)" if two_branch else "tdnn2")
    b.stats_layer("stats", 0, 10000)
    b.relu_batchnorm_layer("tdnn6", emb)
    b.output_layer("output", 4)
    return b.model()


# --------------------------------------------------------------------------
# features
# --------------------------------------------------------------------------


def make_lengths(n_utts: int, lo: int, hi: int, seed: int) -> np.ndarray:
    """Utterance lengths in frames, uniform integer on [lo, hi]."""
    return np.random.default_rng(seed).integers(lo, hi + 1, size=n_utts).astype(np.int64)


def make_features(lengths: Sequence[int], dim: int, seed: int) -> np.ndarray:
    """Post-VAD post-CMN frames ~ N(0,1), all utterances packed row-wise."""
    rng = np.random.default_rng(seed)
    return rng.standard_normal((int(np.sum(lengths)), dim), dtype=np.float32)


def utt_keys(n: int) -> List[str]:
    return [f"utt{i:07d}" for i in range(n)]
