"""SECOND, INDEPENDENT RESTATEMENT OF THE ORACLE -- TEST INFRASTRUCTURE ONLY.

Same algorithm as oracle/nnet3_oracle.py (the forward pass of upstream Kaldi's
`nnet3-xvector-compute`) reached by a different route,
so the two must agree before either is trusted to judge the CUDA path
(SURVEY.md 8c "Oracle plan" item 2):

  * a spliced TDNN layer  Append(Offset(x,-d), x, Offset(x,d))  is a *dilated 1-D
    convolution* (torch.nn.functional.conv1d, channel-first), not a row-splice + matmul;
  * test-mode BatchNorm is torch.nn.functional.batch_norm(training=False) on the
    stored StatsMean / StatsVar, not a precomputed scale/offset;
  * pooling uses torch.mean / torch.var(unbiased=False) (two-pass), not
    E[x^2] - mean^2;
  * node strings are handled with regular expressions, not the descriptor parser.

float64 throughout.
"""
from __future__ import annotations

import re
from typing import Dict, List, Tuple

import numpy as np
import torch
import torch.nn.functional as F


def _node_table(model, output_node: str):
    nodes = {}
    for ln in model.config_lines:
        kind = ln.split()[0]
        name = re.search(r"name=(\S+)", ln).group(1)
        if kind == "input-node":
            nodes[name] = ("input", int(re.search(r"dim=(\d+)", ln).group(1)))
        elif kind == "component-node":
            comp = re.search(r"component=(\S+)", ln).group(1)
            inp = re.search(r"input=(.*?)(?:\s+\w[\w-]*=|$)", ln).group(1).strip()
            nodes[name] = ("component", comp, inp)
    return nodes


def _taps(desc: str) -> List[Tuple[str, int]]:
    """'Append(Offset(a, -2), a, Offset(b, 3))' -> [(a,-2),(a,0),(b,3)]; Round(x, 1) -> [(x,0)]."""
    d = desc.strip()
    m = re.fullmatch(r"Round\((.*),\s*\d+\)", d)
    if m:
        d = m.group(1).strip()
    if d.startswith("Append(") and d.endswith(")"):
        d = d[len("Append("):-1]
    out = []
    for part in re.findall(r"Offset\([^()]+\)|[^,\s()]+", d):
        m = re.fullmatch(r"Offset\(\s*([^,\s]+)\s*,\s*(-?\d+)\s*\)", part)
        out.append((m.group(1), int(m.group(2))) if m else (part, 0))
    return out


def forward_chunk_torch(model, output_node: str, inputs: Dict[str, np.ndarray]) -> np.ndarray:
    nodes = _node_table(model, output_node)
    comps = model.components
    memo = {}

    def ev(name):
        """-> (tensor [C, T'] for frame level or [C] for segment level, t_lo or None)."""
        if name in memo:
            return memo[name]
        nd = nodes[name]
        if nd[0] == "input":
            res = (torch.from_numpy(np.asarray(inputs[name], np.float64).T.copy()), 0)
        else:
            comp = comps[nd[1]]
            f = comp.fields
            if comp.type.endswith("AffineComponent"):
                W = torch.from_numpy(np.asarray(f["LinearParams"], np.float64))
                b = torch.from_numpy(np.asarray(f["BiasParams"], np.float64))
                taps = _taps(nd[2])
                srcs = [ev(n) for n, _ in taps]
                if srcs[0][1] is None:                       # segment level
                    assert len(taps) == 1
                    res = (F.linear(srcs[0][0], W, b), None)
                else:
                    # group consecutive taps by source; each group is one dilated conv
                    groups, col = [], 0
                    for (n, off), (x, lo) in zip(taps, srcs):
                        c = x.shape[0]
                        if groups and groups[-1]["name"] == n:
                            groups[-1]["offs"].append(off)
                            groups[-1]["cols"].append((col, col + c))
                        else:
                            groups.append(dict(name=n, x=x, lo=lo, offs=[off], cols=[(col, col + c)]))
                        col += c
                    assert col == W.shape[1]
                    outs = []
                    for g in groups:
                        offs = g["offs"]
                        k = len(offs)
                        dil = offs[1] - offs[0] if k > 1 else 1
                        assert all(offs[i + 1] - offs[i] == dil for i in range(k - 1)) and dil > 0
                        w = torch.stack([W[:, a:b_] for a, b_ in g["cols"]], dim=2)      # [N, C, k]
                        y = F.conv1d(g["x"][None], w, dilation=dil)[0]                   # [N, T' - (k-1)dil]
                        outs.append((y, g["lo"] - offs[0]))                              # first valid output t
                    lo = max(o[1] for o in outs)
                    hi = min(o[1] + o[0].shape[1] - 1 for o in outs)
                    y = sum(o[0][:, lo - o[1]:hi - o[1] + 1] for o in outs) + b[:, None]
                    res = (y, lo)
            elif comp.type == "RectifiedLinearComponent":
                x, lo = ev(_taps(nd[2])[0][0])
                res = (F.relu(x), lo)
            elif comp.type == "BatchNormComponent":
                x, lo = ev(_taps(nd[2])[0][0])
                mean = torch.from_numpy(np.asarray(f["StatsMean"], np.float64))
                var = torch.from_numpy(np.asarray(f["StatsVar"], np.float64))
                rms = float(f["TargetRms"])
                # batch_norm wants [B, C, L]
                xb = (x[None] if lo is not None else x[None, :, None])
                y = F.batch_norm(xb, mean, var.clamp(min=0), weight=torch.full_like(mean, rms), bias=None,
                                 training=False, eps=float(f["Epsilon"]))
                res = (y[0] if lo is not None else y[0, :, 0], lo)
            elif comp.type == "StatisticsExtractionComponent":
                res = ev(_taps(nd[2])[0][0])                 # pooling below works on x directly
            elif comp.type == "StatisticsPoolingComponent":
                x, lo = ev(_taps(nd[2])[0][0])
                lc, rc = int(f["LeftContext"]), int(f["RightContext"])
                hi = lo + x.shape[1] - 1
                a, b_ = max(lo, -lc), min(hi, rc)
                seg = x[:, a - lo:b_ - lo + 1]
                mean = seg.mean(dim=1)
                var = seg.var(dim=1, unbiased=False)
                floor = float(f.get("VarianceFloor", 1e-10))
                res = (torch.cat([mean, var.clamp(min=floor).sqrt()]), None)
            else:
                raise ValueError(comp.type)
        memo[name] = res
        return res

    y, lo = ev(_taps(output_node)[0][0])
    assert lo is None, "output must be segment level"
    return y.numpy()
