"""Frame-level outputs (the nnet3-compute drop-in) through the C ABI vs the oracle.

Reference call sites: egs/sre/v5/sid/nnet3_cvector/cvector/extract_am_embedding.sh:79-85
(bottleneck features of the phonetic branch) and cvector/extract_log_post.sh:80-85
(log-posteriors)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL_REL = 1e-3  # north-star tolerance (BASELINE.json): max|a-b|/max|b|


@pytest.fixture(scope="module")
def frame_models(tmp_path_factory, kf, synth):
    d = tmp_path_factory.mktemp("frame_models")
    out = {}
    for topo, kw in (("v4_cvector", {}), ("v5_cvector", {"num_senones": 1000}), ("v4_two_input", {})):
        m = synth.make_model(topo, seed=5, random_outputs=True, **kw)
        p = str(d / f"{topo}.raw")
        kf.write_nnet3(p, m, True)
        out[topo] = (p, m)
    return out


def _ref_frames(model, node, feats, lengths, dtype=np.float64):
    from oracle.nnet3_oracle import Oracle
    orc = Oracle(model, node, dtype)
    outs, base = [], 0
    for T in lengths:
        outs.append(orc.forward_frames({n: x[base:base + T] for n, x in feats.items()}))
        base += T
    return np.concatenate(outs, axis=0)


def _feats(synth, mod, lengths, seed):
    return {n: synth.make_features(lengths, d, seed + i) for i, (n, d) in enumerate(zip(mod.input_names, mod.input_dims))}


@pytest.mark.parametrize("precision,rel", [("fp32", 2e-5), ("fp16", TOL_REL)])
@pytest.mark.parametrize("topology,node", [("v4_cvector", "tdnn5.batchnorm"),          # AM bottleneck, 128-d
                                           ("v4_cvector", "tdnn4_xvec.batchnorm"),     # speaker branch mid layer
                                           ("v5_cvector", "output_am.log-softmax"),    # 1000 senones: 4 N tiles
                                           ("v4_two_input", "tdnn5.batchnorm")])
def test_frames_match_oracle(cvx, synth, frame_models, require_gpu, topology, node, precision, rel):
    path, model = frame_models[topology]
    mod = cvx.Model(path, node, frames=True)
    assert mod.info.pooled_dim == 0 and mod.info.num_segment_layers == 0
    lengths = [1, 7, 300, 57, 1234, 13, 2]  # shorter than the context included: edges are replicated
    feats = _feats(synth, mod, lengths, 21)
    ex = cvx.Extractor(mod, precision=precision)
    got = ex.compute_frames(feats, lengths)
    ref = _ref_frames(model, node, feats, lengths)
    assert got.shape == ref.shape == (sum(lengths), mod.info.embedding_dim)
    err = np.abs(got - ref).max() / np.abs(ref).max()
    assert err <= rel, f"{topology}/{node}/{precision}: rel {err}"
    if node.endswith("log-softmax"):
        assert np.allclose(np.exp(got.astype(np.float64)).sum(axis=1), 1.0, atol=1e-4)
        post = ex.compute_frames(feats, lengths, apply_exp=True)
        assert np.abs(post - np.exp(got)).max() <= 1e-6
        # |e^a - e^b| <= |a - b| for a, b <= 0: posteriors are at least as close as the log-posteriors
        assert np.abs(post - np.exp(ref)).max() <= np.abs(got - ref).max() * 1.01 + 1e-6


def test_frames_batching_is_invariant(cvx, synth, frame_models, require_gpu):
    """One utterance per call, all in one call, and many small batches give the same rows bit for bit."""
    path, _ = frame_models["v4_cvector"]
    mod = cvx.Model(path, "tdnn5.batchnorm", frames=True)
    lengths = [40, 333, 5, 1000, 128, 129]
    feats = _feats(synth, mod, lengths, 8)
    whole = cvx.Extractor(mod).compute_frames(feats, lengths)
    small = cvx.Extractor(mod, max_batch_rows=256).compute_frames(feats, lengths)
    assert np.array_equal(whole, small)
    base = 0
    ex = cvx.Extractor(mod)
    for T in lengths:
        one = ex.compute_frames({n: x[base:base + T] for n, x in feats.items()}, [T])
        assert np.array_equal(one, whole[base:base + T])
        base += T


def test_frame_and_extraction_entry_points_do_not_mix(cvx, synth, frame_models, require_gpu):
    path, _ = frame_models["v4_cvector"]
    fm = cvx.Model(path, "tdnn5.batchnorm", frames=True)
    em = cvx.Model(path, "tdnn6_xvec.affine")
    feats = _feats(synth, em, [100], 1)
    with pytest.raises(cvx.CvxError):
        cvx.Extractor(em).compute_frames(feats, [100])
    with pytest.raises(cvx.CvxError):
        cvx.Extractor(fm).extract(feats, [100], chunk_size=100, min_chunk_size=25)
    with pytest.raises(cvx.CvxError):
        cvx.Model(path, "tdnn6_xvec.affine", frames=True)      # segment-level node under the frame loader
    with pytest.raises(cvx.CvxError):
        cvx.Extractor(fm).compute_frames(feats, [100, 0])      # empty utterance


def test_cli_nnet3_compute(cvx, kf, synth, frame_models, require_gpu, tmp_path):
    """nnet3-compute with the recipes' argument shapes: bottleneck features to a pipe-style ark
    (extract_am_embedding.sh:79-85, output redirected by an extract.config) and log-posteriors to
    ark,scp (extract_log_post.sh:80-85); zero-length utterances fail like upstream."""
    import os
    import subprocess
    path, model = frame_models["v4_cvector"]
    lengths = [120, 3, 511, 64]
    x = synth.make_features(lengths, 23, 77)
    keys = synth.utt_keys(len(lengths))
    items, base = [], 0
    for k, T in zip(keys, lengths):
        items.append((k, x[base:base + T]))
        base += T
    items.insert(2, ("empty_utt", np.zeros((0, 23), np.float32)))
    feats_ark = str(tmp_path / "feats.ark")
    kf.write_ark(feats_ark, items)
    cfg = str(tmp_path / "extract.config")
    with open(cfg, "w") as f:
        f.write("output-node name=output input=tdnn5.batchnorm\n")
    exe = os.path.join(cvx.BIN_DIR, "nnet3-compute")
    out_ark = str(tmp_path / "bn.ark")
    r = subprocess.run([exe, "--use-gpu=yes", "--frames-per-chunk=50", f"--nnet-config={cfg}", f"cat {path} |",
                        f"ark:cat {feats_ark} |", f"ark:{out_ark}"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    assert "Done 4 utterances, failed for 1" in r.stderr and "Zero-length utterance: empty_utt" in r.stderr, r.stderr
    got = list(kf.read_ark(out_ark))
    assert [k for k, _ in got] == keys
    ref = _ref_frames(model, "tdnn5.batchnorm", {"input": x}, lengths)
    a = np.concatenate([m for _, m in got], axis=0)
    assert [m.shape[0] for _, m in got] == lengths
    assert np.abs(a - ref).max() / np.abs(ref).max() <= TOL_REL
    # log-posteriors / posteriors with ark,scp and text output
    ark, scp = str(tmp_path / "post.ark"), str(tmp_path / "post.scp")
    r = subprocess.run([exe, "--use-gpu=no", "--apply-exp=true", "--output-node=output_am.log-softmax", path,
                        f"ark:{feats_ark}", f"ark,scp:{ark},{scp}"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    post = dict(kf.read_scp(scp))
    assert list(post) == keys
    ref = _ref_frames(model, "output_am.log-softmax", {"input": x}, lengths)
    got = np.concatenate([post[k] for k in keys])
    assert np.allclose(got.sum(axis=1), 1.0, atol=1e-4)
    assert np.abs(np.log(np.maximum(got, 1e-30)) - ref)[ref > -20].max() / np.abs(ref).max() <= TOL_REL
    r = subprocess.run([exe, "--output-node=tdnn6_xvec.affine", path, f"ark:{feats_ark}", f"ark,t:{tmp_path}/x.txt"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode != 0 and "segment-level" in r.stderr


def test_frame_golden_fixture_through_the_cli(cvx, kf, require_gpu, tmp_path):
    """committed golden rows (tests/golden/v5_small_frames.npz): nnet3-compute on the committed model
    and feature archive, output redirected to the phonetic bottleneck."""
    import os
    import subprocess
    g = os.path.join(os.path.dirname(__file__), "golden")
    zf = np.load(os.path.join(g, "v5_small_frames.npz"))
    exe = os.path.join(cvx.BIN_DIR, "nnet3-compute")
    r = subprocess.run([exe, f"--output-node={zf['output_node']}", f"{g}/v5_small_final.raw", f"ark:{g}/v5_small_feats.ark",
                        f"ark:{tmp_path}/bn.ark"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    assert "Done 6 utterances, failed for 1" in r.stderr, r.stderr       # one zero-length utterance in the archive
    got = np.concatenate([m for _, m in kf.read_ark(f"{tmp_path}/bn.ark")])
    assert got.shape == zf["frames"].shape
    assert np.abs(got - zf["frames"]).max() / np.abs(zf["frames"]).max() <= TOL_REL
