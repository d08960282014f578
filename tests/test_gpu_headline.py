"""Oracle parity at the shapes the bench runs ("parity untested at headline shapes"):
full-width nets, --chunk-size=10000 on utterances up to 30 000 frames
(R/egs/sre/v5/sid/nnet3/xvector/extract_xvectors_new.sh:62-68, run_xvector_new.sh:78-88), the
short-segment sweep (BASELINE configs[4]), the two-input config (configs[2]) and the remaining
reference topologies (prepare_nnet3_xconfig_{2,3,4}share.sh:50-63, train_cvector_with_am_expand.sh:77).
Every comparison is CUDA (through the C ABI) vs the fp64 oracle on the same seeded inputs."""
import numpy as np
import pytest

from conftest import embedding_node

pytestmark = pytest.mark.gpu

# north-star tolerance (BASELINE.json): max|a-b|/max|b| <= 1e-3 and cosine >= 0.9999
TOL_REL, TOL_COS = 1e-3, 0.9999


def _oracle(model, node):
    from oracle.nnet3_oracle import Oracle
    return Oracle(model, node, np.float64)


def _inputs(synth, mod, lengths, seed):
    return {n: synth.make_features(lengths, d, seed + i) for i, (n, d) in enumerate(zip(mod.input_names, mod.input_dims))}


@pytest.mark.parametrize("topology", ["v5_cvector", "v2_xvector"])
def test_chunk_size_10000_long_utterances(cvx, synth, model_dir, require_gpu, topology):
    """cfg4 shapes: one 30 000-frame utterance (3 full chunks), one of exactly 10 000, a ragged
    12 345 (10 000 + 2 345), one whose tail (20 frames) is dropped by min-chunk-size, a short one."""
    from oracle.nnet3_oracle import extract_all, parity_metrics
    path, model = model_dir(topology)
    node = embedding_node(synth, topology)
    mod = cvx.Model(path, node)
    lengths = [30000, 10000, 12345, 20020, 1000]
    feats = _inputs(synth, mod, lengths, 101)
    ex = cvx.Extractor(mod, precision="fp16")
    got, ok = ex.extract(feats, lengths, chunk_size=10000, min_chunk_size=25)
    ref, ok_ref = extract_all(_oracle(model, node), feats, lengths, chunk_size=10000, min_chunk_size=25)
    assert ok.all() and ok_ref.all()
    r, c = parity_metrics(got, ref)
    print(f"{topology} chunk 10000: rel {r:.3e} cos {c:.8f}")
    assert r <= TOL_REL and c >= TOL_COS, (r, c)
    # several device batches (chunks of 10 000 rows cut over 16 Ki-row batches) give the same answer
    small = cvx.Extractor(mod, precision="fp16", max_batch_rows=16384)
    got2, _ = small.extract(feats, lengths, chunk_size=10000, min_chunk_size=25)
    r2, c2 = parity_metrics(got2, ref)
    assert r2 <= TOL_REL and c2 >= TOL_COS, (r2, c2)


@pytest.mark.parametrize("topology", ["v2_xvector", "v5_cvector"])
def test_short_segment_sweep_sample(cvx, synth, model_dir, require_gpu, topology):
    """cfg5 shapes: 12 000 x 200-frame segments in one call (2.4 M rows: five device batches, one
    segment-layer flush of 12 000 chunks); a random 64 of them against the oracle."""
    from oracle.nnet3_oracle import parity_metrics
    path, model = model_dir(topology)
    node = embedding_node(synth, topology)
    mod = cvx.Model(path, node)
    n = 12000
    lengths = [200] * n
    feats = _inputs(synth, mod, lengths, 55)
    ex = cvx.Extractor(mod, precision="fp16")
    got, ok = ex.extract(feats, lengths, chunk_size=-1, min_chunk_size=25)
    assert ok.all() and np.isfinite(got).all()
    orc = _oracle(model, node)
    pick = np.random.default_rng(7).choice(n, 64, replace=False)
    pick = np.concatenate([pick, [0, n - 1, 2621, 2622]])        # first, last, the 512 Ki-row batch boundary
    ref = np.stack([orc.forward_chunk({k: v[u * 200:(u + 1) * 200] for k, v in feats.items()}) for u in pick])
    r, c = parity_metrics(got[pick], ref)
    print(f"{topology} 200-frame segments: rel {r:.3e} cos {c:.8f}")
    assert r <= TOL_REL and c >= TOL_COS, (r, c)


def test_two_input_config3_shapes(cvx, synth, model_dir, require_gpu):
    """cfg3: v4 net with the 40-dim hires stream feeding the phonetic TDNN, utterances of 3-30 s."""
    from oracle.nnet3_oracle import extract_all, parity_metrics
    path, model = model_dir("v4_two_input")
    node = embedding_node(synth, "v4_two_input")
    mod = cvx.Model(path, node)
    assert sorted(mod.input_names) == ["input", "input_am"]
    lengths = [3000, 300, 1777, 2999]
    feats = _inputs(synth, mod, lengths, 71)
    got, ok = cvx.Extractor(mod, precision="fp16").extract(feats, lengths, chunk_size=10000, min_chunk_size=25)
    ref, _ = extract_all(_oracle(model, node), feats, lengths, chunk_size=10000, min_chunk_size=25)
    r, c = parity_metrics(got[ok], ref[ok])
    assert ok.all() and r <= TOL_REL and c >= TOL_COS, (r, c)


@pytest.mark.parametrize("precision,rel,cos", [("fp32", 2e-5, 0.999999), ("fp16", TOL_REL, TOL_COS)])
@pytest.mark.parametrize("topology", ["v3_multitask_2share", "v3_multitask_3share", "v3_multitask_4share", "v5_expand"])
def test_remaining_reference_topologies(cvx, synth, model_dir, require_gpu, topology, precision, rel, cos):
    """the later-fork multitask nets and the "expand" net, whose first speaker layer mixes raw-feature
    taps and phonetic-bottleneck taps in one Append (staged input copy + layer-source K blocks in one
    GEMM)."""
    from oracle.nnet3_oracle import extract_all, parity_metrics
    path, model = model_dir(topology)
    node = embedding_node(synth, topology)
    mod = cvx.Model(path, node)
    ctx = mod.info.left_context + mod.info.right_context + 1
    # shortest utterances: ctx + 7 frames = 8 pooled frames (fp16) / exactly ctx = ONE pooled frame (fp32)
    lengths = [ctx + 7 if precision == "fp16" else ctx, 300, 57, 1234, 100, ctx + 12, 640]
    feats = _inputs(synth, mod, lengths, 13)
    got, ok = cvx.Extractor(mod, precision=precision).extract(feats, lengths, chunk_size=500, min_chunk_size=25)
    ref, ok_ref = extract_all(_oracle(model, node), feats, lengths, chunk_size=500, min_chunk_size=25)
    assert ok.tolist() == ok_ref.tolist()
    per_utt = [parity_metrics(got[u][None], ref[u][None]) for u in range(len(lengths)) if ok[u]]
    print(topology, precision, [(n, f"{r:.2e}") for n, (r, _) in zip(lengths, per_utt)])
    r, c = parity_metrics(got[ok], ref[ok])
    assert r <= rel and c >= cos, f"{topology}/{precision}: rel {r} cos {c} per-utt {per_utt}"


def test_single_pooled_frame_of_the_deepest_net_in_fp16(cvx, synth, model_dir, require_gpu):
    """Documented limit of fp16 operands: a chunk of exactly left+right+1 frames pools ONE
    frame, so nothing averages the operand rounding of the 10 layers the "expand" net stacks in front of
    the pooling (the fp16-rounding model of the oracle predicts 1.1e-3 for it).  The cosine stays four
    nines beyond the tolerance; max-rel is held to 3e-3 here and to 1e-3 from 8 pooled frames on (above)."""
    from oracle.nnet3_oracle import extract_all, parity_metrics
    path, model = model_dir("v5_expand")
    node = embedding_node(synth, "v5_expand")
    mod = cvx.Model(path, node)
    lengths = [35, 36, 38]
    feats = _inputs(synth, mod, lengths, 13)
    got, ok = cvx.Extractor(mod, precision="fp16").extract(feats, lengths, chunk_size=-1, min_chunk_size=25)
    ref, _ = extract_all(_oracle(model, node), feats, lengths, chunk_size=-1, min_chunk_size=25)
    r, c = parity_metrics(got, ref)
    print(f"v5_expand single pooled frame: rel {r:.3e} cos {c:.8f}")
    assert ok.all() and r <= 3e-3 and c >= 0.99999, (r, c)
