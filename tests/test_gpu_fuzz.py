"""Randomised topologies: the loader / lowering / kernels against the oracle on networks the recipes
do not contain -- random tap sets (including unsorted-looking spans up to 8, repeated sources,
raw-input taps mixed with layer taps as in train_cvector_with_am_expand.sh:77), widths that are not
multiples of 16 / 32 / 64, two-source Appends, one or two segment layers -- so that every padding,
K-block, halo-grouping and tile-splitting path is taken with shapes nobody tuned for."""
import importlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def random_model(synth, seed):
    rng = np.random.default_rng(seed)
    b = synth._Builder(rng)
    feat_dim = int(rng.integers(5, 41))
    b.input("input", feat_dim)
    outs = ["input"]
    n_layers = int(rng.integers(2, 6))
    for i in range(n_layers):
        k = 1 if len(outs) == 1 else int(rng.integers(1, 3))
        pick = rng.choice(len(outs), size=min(k, len(outs)), replace=False)
        if i > 0 and (len(outs) - 1) not in pick.tolist():
            pick[0] = len(outs) - 1                      # keep the chain connected: always tap the previous layer
        taps = []
        for s in sorted(pick.tolist()):
            offs = sorted(set(rng.integers(-4, 5, size=int(rng.integers(1, 4))).tolist()))
            taps += [f"{outs[s]}@{o}" for o in offs]
        dim = int(rng.choice([8, 17, 33, 64, 100, 129, 200, 257, 300]))
        b.relu_batchnorm_layer(f"l{i}", dim, "Append(" + ",".join(taps) + ")" if len(taps) > 1 else taps[0])
        outs.append(f"l{i}")
    b.stats_layer("stats", 0, 10000, outs[-1])
    emb = int(rng.choice([12, 64, 100]))
    b.relu_batchnorm_layer("seg1", emb)
    two_seg = bool(rng.integers(0, 2))
    if two_seg:
        b.relu_batchnorm_layer("seg2", emb + 4, "seg1")
    b.output_layer("output", 4)
    return b.model(), ("seg2.affine" if two_seg else "seg1.affine"), feat_dim


@pytest.mark.parametrize("seed", list(range(100, 132)))
def test_random_topology_matches_oracle(cvx, kf, synth, require_gpu, tmp_path, seed):
    from oracle.nnet3_oracle import Oracle, extract_all, parity_metrics
    model, node, feat_dim = random_model(synth, seed)
    path = str(tmp_path / "net.raw")
    kf.write_nnet3(path, model, True)
    L, R = Oracle(model, node, np.float64).context()
    try:
        mod = cvx.Model(path, node)
    except cvx.CvxError as e:
        # Nets in which some layer is needed before the first / after the last input frame (a forward-only
        # source tapped backwards, a pooled layer that only looks back) are REFUSED, never computed differently:
        # the packed layout holds frames 0..T-1 of every layer (csrc/plan.cc).  About 4 in 10 random nets.
        assert e.code == -4 and "computable on frames" in str(e), str(e)
        return
    assert (mod.info.left_context, mod.info.right_context) == (max(L, 0), R)     # nnet3's own context of the pooled layer
    ctx = mod.info.left_context + mod.info.right_context + 1
    rng = np.random.default_rng(seed + 1)
    lengths = [ctx, ctx + 3] + rng.integers(ctx + 40, 900, size=4).tolist()
    feats = {"input": synth.make_features(lengths, feat_dim, seed)}
    chunk = int(rng.choice([-1, 250, 400]))
    minc = min(25, ctx)
    ref, ok_ref = extract_all(Oracle(model, node, np.float64), feats, lengths, chunk_size=chunk, min_chunk_size=minc)
    for precision, rel, cos in (("fp32", 5e-5, 0.999999), ("fp16", 2e-3, 0.9999)):
        got, ok = cvx.Extractor(mod, precision=precision).extract(feats, lengths, chunk_size=chunk, min_chunk_size=minc)
        assert ok.tolist() == ok_ref.tolist(), (seed, precision)
        # fp16: the two shortest utterances pool 1 and 4 frames (no averaging of the rounding noise, DESIGN.md
        # section 4); the tolerance is applied from the third utterance on, the short ones must stay sane
        sel = ok if precision == "fp32" else ok & (np.arange(len(lengths)) >= 2)
        r, c = parity_metrics(got[sel], ref[sel])
        assert r <= rel and c >= cos, f"seed {seed} {precision}: rel {r} cos {c}\n{mod.describe()}"
        r_all, c_all = parity_metrics(got[ok], ref[ok])
        assert r_all <= 2e-2 and c_all >= 0.999, f"seed {seed} {precision} short utterances: rel {r_all} cos {c_all}"
