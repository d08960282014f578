"""Embedding post-step: mean subtraction -> transform-vec -> length normalisation on
the device, stand-alone and fused behind the extraction, against the numpy restatement."""
import numpy as np
import pytest

from conftest import embedding_node

pytestmark = pytest.mark.gpu


def _chain_inputs(dim, out_dim, n, seed, affine=False):
    rng = np.random.default_rng(seed)
    v = rng.normal(0.2, 1.5, (n, dim)).astype(np.float32)
    mean = rng.normal(0.2, 0.3, dim).astype(np.float32)
    t = rng.normal(0, dim ** -0.5, (out_dim, dim + (1 if affine else 0))).astype(np.float32)
    return v, mean, t


@pytest.mark.parametrize("dim,out_dim,affine", [(512, 150, False), (512, 150, True), (100, 100, False), (3000, 200, False),
                                                (7, 3, True)])
def test_transform_vectors_matches_restatement(cvx, require_gpu, dim, out_dim, affine):
    from oracle.backend_oracle import vector_chain
    v, mean, t = _chain_inputs(dim, out_dim, 257, dim + out_dim, affine)
    got, ratios = cvx.transform_vectors(v, mean, t)
    ref, ref_ratio = vector_chain(v, mean, t)
    assert got.shape == (257, out_dim)
    assert np.abs(got - ref).max() <= 2e-5 * np.abs(ref).max()
    assert np.allclose(ratios, ref_ratio, rtol=2e-5)
    assert np.allclose(np.linalg.norm(got.astype(np.float64), axis=1), np.sqrt(out_dim), rtol=1e-5)
    # each stage alone, and the scaleup / normalize switches
    a, _ = cvx.transform_vectors(v, mean, None, normalize_length=False)
    assert np.allclose(a, v - mean, atol=1e-6)
    b, _ = cvx.transform_vectors(v, None, t, normalize_length=False)
    assert np.abs(b - vector_chain(v, None, t, normalize=False)[0]).max() <= 2e-5 * np.abs(b).max()
    c, rc = cvx.transform_vectors(v, None, None, normalize_length=True, scaleup=False)
    assert np.allclose(np.linalg.norm(c.astype(np.float64), axis=1), 1.0, rtol=1e-5)
    assert np.allclose(rc, np.linalg.norm(v.astype(np.float64), axis=1), rtol=1e-5)
    z, rz = cvx.transform_vectors(np.zeros((2, dim), np.float32), None, None)
    assert np.all(z == 0) and np.all(rz == 0)          # "Zero iVector": left alone
    with pytest.raises(cvx.CvxError):
        cvx.transform_vectors(v, mean, t[:, :-2])      # wrong number of columns


def test_chain_fused_behind_extraction(cvx, synth, model_dir, require_gpu):
    """cvx_engine_set_vector_chain: extract() returns what extract() + the three back-end programs
    would; failed utterances stay zero; removing the chain restores the plain embeddings."""
    from oracle.backend_oracle import vector_chain
    path, _ = model_dir("v5_cvector")
    mod = cvx.Model(path, embedding_node(synth, "v5_cvector"))
    lengths = [400, 10, 250, 1200, 90]
    feats = {n: synth.make_features(lengths, d, 5 + i) for i, (n, d) in enumerate(zip(mod.input_names, mod.input_dims))}
    ex = cvx.Extractor(mod)
    plain, ok = ex.extract(feats, lengths, chunk_size=500, min_chunk_size=25)
    assert ok.tolist() == [True, False, True, True, True]
    _, mean, t = _chain_inputs(mod.info.embedding_dim, 150, 1, 99)
    ex.set_vector_chain(mean, t)
    assert ex.output_dim == 150
    fused, ok2 = ex.extract(feats, lengths, chunk_size=500, min_chunk_size=25)
    assert fused.shape == (5, 150) and ok2.tolist() == ok.tolist()
    ref, _ = vector_chain(plain[ok], mean, t)
    assert np.abs(fused[ok] - ref).max() <= 2e-5 * np.abs(ref).max()
    assert np.all(fused[~ok] == 0)
    ex.set_vector_chain(clear=True)
    again, _ = ex.extract(feats, lengths, chunk_size=500, min_chunk_size=25)
    assert ex.output_dim == mod.info.embedding_dim and np.array_equal(again, plain)


def _write_obj(kf, path, a, binary=True):
    with open(path, "wb") as f:
        if binary:
            f.write(b"\0B")
        (kf.write_vector if a.ndim == 1 else kf.write_matrix)(f, binary, a)


def _read_one(kf, path):
    data = open(path, "rb").read()
    r = kf._Reader(data)
    binary = r.peek(2) == b"\0B"
    if binary:
        r.read(2)
    return kf.read_vector_or_matrix(r, binary)


def test_cli_backend_programs(cvx, kf, require_gpu, tmp_path):
    """ivector-mean (both forms), ivector-subtract-global-mean | transform-vec | ivector-normalize-length
    with the argument shapes of run_sre10.sh:107-129 / extract_xvectors_new.sh:102-108."""
    import os
    import subprocess
    from oracle.backend_oracle import speaker_means, subtract_global_mean, vector_chain
    B = cvx.BIN_DIR
    v, _, t = _chain_inputs(512, 150, 23, 7)
    keys = [f"utt{i:03d}" for i in range(len(v))]
    ark, scp = str(tmp_path / "xvector.ark"), str(tmp_path / "xvector.scp")
    kf.write_ark(ark, list(zip(keys, v)), scp_path=scp)
    # speaker means + num_utts; spk3 has one utterance without a vector, spk4 none at all
    spk2utt = [("spk1", keys[:10]), ("spk2", keys[10:22]), ("spk3", [keys[22], "missing_a"]), ("spk4", ["missing_b"])]
    with open(tmp_path / "spk2utt", "w") as f:
        for s, us in spk2utt:
            f.write(s + " " + " ".join(us) + "\n")
    r = subprocess.run([f"{B}/ivector-mean", f"ark:{tmp_path}/spk2utt", f"scp:{scp}",
                        f"ark,scp:{tmp_path}/spk.ark,{tmp_path}/spk.scp", f"ark,t:{tmp_path}/num_utts.ark"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    ref_means, ref_counts, missing = speaker_means(spk2utt, dict(zip(keys, v)))
    got = list(kf.read_scp(f"{tmp_path}/spk.scp"))
    assert [k for k, _ in got] == [s for s, _ in ref_means] == ["spk1", "spk2", "spk3"]
    for (_, a), (_, b) in zip(got, ref_means):
        assert np.allclose(a, b, atol=1e-6)
    assert open(tmp_path / "num_utts.ark").read().split() == [x for s, n in ref_counts for x in (s, str(n))]
    assert missing == 2 and "No iVector present in input for utterance missing_a" in r.stderr
    # global mean to a file
    r = subprocess.run([f"{B}/ivector-mean", f"scp:{scp}", f"{tmp_path}/mean.vec"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    mean = _read_one(kf, f"{tmp_path}/mean.vec")
    assert np.allclose(mean, v.astype(np.float64).mean(axis=0), atol=1e-6)
    # the scoring-time chain as a shell pipe
    _write_obj(kf, f"{tmp_path}/transform.mat", t)
    env = dict(os.environ, PATH=B + os.pathsep + os.environ["PATH"])
    pipe = (f"ivector-subtract-global-mean {tmp_path}/mean.vec scp:{scp} ark:- | "
            f"transform-vec {tmp_path}/transform.mat ark:- ark:- | "
            f"ivector-normalize-length ark:- ark,t:{tmp_path}/final.txt")
    r = subprocess.run(pipe, shell=True, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "Average ratio of iVector to expected length was" in r.stderr
    final = dict(kf.read_ark(f"{tmp_path}/final.txt"))
    ref, _ = vector_chain(v, mean, t)
    assert list(final) == keys
    assert np.abs(np.stack([final[k] for k in keys]) - ref).max() <= 3e-5 * np.abs(ref).max()
    # two-argument form: subtract the table's own mean (LDA/PLDA training input, run_sre10.sh:118)
    r = subprocess.run([f"{B}/ivector-subtract-global-mean", f"scp:{scp}", f"ark:{tmp_path}/centered.ark"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    c = np.stack([a for _, a in kf.read_ark(f"{tmp_path}/centered.ark")])
    assert np.allclose(c, subtract_global_mean(v), atol=2e-6)


def test_cli_extraction_with_fused_backend_chain(cvx, kf, require_gpu, tmp_path):
    """nnet3-xvector-compute --mean-vec --transform-mat --normalize-length writes what the plain
    program piped through the three back-end programs would."""
    import os
    import subprocess
    from oracle.backend_oracle import vector_chain
    g = os.path.join(os.path.dirname(__file__), "golden")
    z = np.load(os.path.join(g, "v5_small.npz"))
    E = z["embeddings"].shape[1]
    _, mean, t = _chain_inputs(E, 24, 1, 3)
    _write_obj(kf, f"{tmp_path}/mean.vec", mean)
    _write_obj(kf, f"{tmp_path}/transform.mat", t, binary=False)
    exe = os.path.join(cvx.BIN_DIR, "nnet3-xvector-compute")
    common = [f"--min-chunk-size={int(z['min_chunk_size'])}", f"--chunk-size={int(z['chunk_size'])}",
              f"--output-node={z['output_node']}"]
    r = subprocess.run([exe, *common, f"{g}/v5_small_final.raw", f"ark:{g}/v5_small_feats.ark", f"ark:{tmp_path}/plain.ark"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe, *common, f"--mean-vec={tmp_path}/mean.vec", f"--transform-mat={tmp_path}/transform.mat",
                        "--normalize-length=true", f"{g}/v5_small_final.raw", f"ark:{g}/v5_small_feats.ark",
                        f"ark:{tmp_path}/fused.ark"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    plain = dict(kf.read_ark(f"{tmp_path}/plain.ark"))
    fused = dict(kf.read_ark(f"{tmp_path}/fused.ark"))
    assert list(plain) == list(fused) and len(plain) == int(z["ok"].sum())
    ref, _ = vector_chain(np.stack(list(plain.values())), mean, t)
    got = np.stack(list(fused.values()))
    assert got.shape[1] == 24 and np.abs(got - ref).max() <= 3e-5 * np.abs(ref).max()


def _plda_model(dim_in, dim, seed):
    rng = np.random.default_rng(seed)
    mean = rng.normal(0.1, 0.3, dim_in)
    tf = rng.normal(0, dim_in ** -0.5, (dim, dim_in))
    psi = np.sort(rng.uniform(0.05, 8.0, dim))[::-1].copy()
    return mean, tf, psi


@pytest.mark.parametrize("normalize,simple", [(True, False), (True, True), (False, False)])
def test_plda_scoring_matches_restatement(cvx, require_gpu, normalize, simple):
    from oracle.backend_oracle import plda_llr, plda_transform
    rng = np.random.default_rng(5)
    mean, tf, psi = _plda_model(150, 150, 1)
    train = rng.normal(0.1, 1.0, (12, 150)).astype(np.float32)
    test = rng.normal(0.1, 1.0, (40, 150)).astype(np.float32)
    n = rng.integers(1, 9, 12).astype(np.int32)
    ti = rng.integers(0, 12, 500).astype(np.int32)
    tj = rng.integers(0, 40, 500).astype(np.int32)
    got = cvx.plda_score(mean, tf, psi, train, test, ti, tj, train_num_utts=n, normalize_length=normalize,
                         simple_length_norm=simple)
    yt = plda_transform(train, mean, tf, psi, n, normalize, simple)
    ye = plda_transform(test, mean, tf, psi, 1, normalize, simple)
    ref = np.array([plda_llr(yt[i], float(n[i]), ye[j], psi) for i, j in zip(ti, tj)])
    assert np.abs(got - ref).max() <= 1e-4 * max(1.0, np.abs(ref).max())
    # without num_utts every enrolment vector counts as one utterance
    got1 = cvx.plda_score(mean, tf, psi, train, test, ti, tj, normalize_length=normalize, simple_length_norm=simple)
    yt1 = plda_transform(train, mean, tf, psi, 1, normalize, simple)
    ref1 = np.array([plda_llr(yt1[i], 1.0, ye[j], psi) for i, j in zip(ti, tj)])
    assert np.abs(got1 - ref1).max() <= 1e-4 * max(1.0, np.abs(ref1).max())
    with pytest.raises(cvx.CvxError):
        cvx.plda_score(mean, tf, psi, train, test, [12], [0])          # trial refers to a missing vector


def test_cli_plda_scoring(cvx, kf, require_gpu, tmp_path):
    """ivector-plda-scoring with the recipe's argument shapes (run_sre10.sh:122-129): model through a
    pipe, --num-utts from ivector-mean's text table, enrolment and test tables through the
    subtract / transform / normalise pipes, trials through `cut`."""
    import os
    import subprocess
    from oracle.backend_oracle import plda_llr, plda_transform, vector_chain
    B = cvx.BIN_DIR
    rng = np.random.default_rng(11)
    v, gmean, lda = _chain_inputs(64, 20, 30, 13)
    mean, tf, psi = _plda_model(20, 20, 2)
    with open(tmp_path / "plda", "wb") as f:
        f.write(b"\0B<Plda> ")
        kf.write_vector(f, True, mean)
        kf.write_matrix(f, True, tf)
        kf.write_vector(f, True, psi)
        f.write(b"</Plda> ")
    _write_obj(kf, f"{tmp_path}/mean.vec", gmean)
    _write_obj(kf, f"{tmp_path}/transform.mat", lda)
    spk = [f"spk{i}" for i in range(6)]
    utt = [f"utt{i:02d}" for i in range(24)]
    kf.write_ark(f"{tmp_path}/spk.ark", list(zip(spk, v[:6])))
    kf.write_ark(f"{tmp_path}/test.ark", list(zip(utt, v[6:])))
    nu = rng.integers(1, 6, 6)
    with open(tmp_path / "num_utts.ark", "w") as f:
        for s, n in zip(spk, nu):
            f.write(f"{s} {n}\n")
    trials = [(spk[i], utt[j], "target" if (i + j) % 3 == 0 else "nontarget") for i in range(6) for j in range(0, 24, 2)]
    trials.insert(5, ("spk_missing", utt[0], "nontarget"))
    with open(tmp_path / "trials", "w") as f:
        for t in trials:
            f.write(" ".join(t) + "\n")
    env = dict(os.environ, PATH=B + os.pathsep + os.environ["PATH"])
    chain = (f"ivector-subtract-global-mean {tmp_path}/mean.vec ark:{{}} ark:- | transform-vec {tmp_path}/transform.mat ark:- ark:- | "
             f"ivector-normalize-length ark:- ark:- |")
    cmd = ["ivector-plda-scoring", "--normalize-length=true", f"--num-utts=ark:{tmp_path}/num_utts.ark",
           f"cat {tmp_path}/plda |", "ark:" + chain.format(f"{tmp_path}/spk.ark"), "ark:" + chain.format(f"{tmp_path}/test.ark"),
           f"cat '{tmp_path}/trials' | cut -d\\  --fields=1,2 |", f"{tmp_path}/scores"]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "Processed 72 trials, 1 had errors." in r.stderr and "Key spk_missing not present in training iVectors." in r.stderr
    lines = [ln.split() for ln in open(tmp_path / "scores")]
    assert [(a, b) for a, b, _ in lines] == [(a, b) for a, b, _ in trials if a != "spk_missing"]
    x, _ = vector_chain(v, gmean, lda)
    yt = plda_transform(x[:6], mean, tf, psi, nu, True, False)
    ye = plda_transform(x[6:], mean, tf, psi, 1, True, False)
    ref = [plda_llr(yt[spk.index(a)], float(nu[spk.index(a)]), ye[utt.index(b)], psi) for a, b, _ in lines]
    got = np.array([float(s) for _, _, s in lines])
    assert np.abs(got - np.array(ref)).max() <= 2e-4 * max(1.0, np.abs(ref).max())
