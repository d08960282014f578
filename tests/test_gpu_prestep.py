"""Device feature pre-step (sliding CMN + VAD selection) against its oracle, alone and
fused in front of the extraction."""
import importlib

import numpy as np
import pytest

from conftest import PKG

pytestmark = pytest.mark.gpu


def _data(lengths, dim, seed, voiced_p=0.7):
    rng = np.random.default_rng(seed)
    x = (rng.standard_normal((sum(lengths), dim)) * 3 + rng.standard_normal(dim) * 10).astype(np.float32)
    vad = [(rng.random(T) < voiced_p).astype(np.float32) for T in lengths]
    return x, vad


@pytest.mark.parametrize("center,window", [(True, 300), (False, 300), (True, 7), (True, 0)])
def test_prestep_matches_oracle(cvx, model_dir, require_gpu, center, window):
    from oracle.prestep_oracle import select_voiced, sliding_cmn
    ps = importlib.import_module(PKG + ".prestep")
    path, _ = model_dir("tiny")
    ex = cvx.Extractor(cvx.Model(path, "tdnn6.affine"), precision="fp32")
    lengths = [1, 5, 149, 150, 299, 300, 301, 1000, 37]          # around the window / half-window edges
    x, vad = _data(lengths, 5, 1)
    vad[0][:] = 1
    got = ps.run_prestep(ex, x, lengths, vad, cmn_window=window, center=center)
    want, r = [], 0
    for T, v in zip(lengths, vad):
        seg = x[r:r + T]
        seg = sliding_cmn(seg, window, center) if window > 0 else seg
        want.append(select_voiced(seg, v))
        r += T
    want = np.concatenate(want)
    assert got.shape == want.shape
    assert np.max(np.abs(got - want)) <= 2e-5 * max(1.0, np.max(np.abs(want)))
    # no VAD: every frame kept
    got_all = ps.run_prestep(ex, x, lengths, None, cmn_window=window, center=center)
    assert got_all.shape == x.shape


def test_extract_raw_equals_pipe_then_extract(cvx, synth, model_dir, require_gpu):
    """fused pre-step + extraction == (oracle pre-step) -> cvx_extract: bit identical inputs
    are not expected (double vs oracle order), so compare embeddings to tolerance, and the
    failure rule (no voiced frames / too few frames) to exactness."""
    from oracle.prestep_oracle import select_voiced, sliding_cmn
    ps = importlib.import_module(PKG + ".prestep")
    path, _ = model_dir("v2_xvector")
    ex = cvx.Extractor(cvx.Model(path, "tdnn6.affine"), precision="fp32")
    lengths = [400, 90, 700, 30, 250]
    x, vad = _data(lengths, 23, 2)
    vad[3][:] = 0                                                   # nothing voiced -> fails
    emb, ok = ps.extract_raw(ex, {"input": x}, lengths, vad, {"input": 300}, chunk_size=200, min_chunk_size=25)
    pieces, kept, r = [], [], 0
    for T, v in zip(lengths, vad):
        seg = select_voiced(sliding_cmn(x[r:r + T], 300, True), v)
        pieces.append(seg); kept.append(len(seg)); r += T
    ref, ok_ref = ex.extract({"input": np.concatenate(pieces)}, kept, chunk_size=200, min_chunk_size=25)
    assert ok.tolist() == ok_ref.tolist() and not ok[3]
    assert np.max(np.abs(emb[ok] - ref[ok])) <= 1e-4 * np.max(np.abs(ref[ok]))


def test_cli_raw_features_with_vad(cvx, kf, synth, model_dir, require_gpu, tmp_path):
    """the drop-in program reading RAW features + vad table directly:
    `--cmn-window=300 --center=true --vad-rspecifier=scp:vad.scp scp:feats.scp` replaces the
    `apply-cmvn-sliding | select-voiced-frames` pipe of extract_xvectors_new.sh:79."""
    import os
    import subprocess
    from oracle.nnet3_oracle import Oracle, extract_all, parity_metrics
    from oracle.prestep_oracle import select_voiced, sliding_cmn
    path, model = model_dir("v2_xvector")
    lengths = [500, 60, 900, 40]
    x, vad = _data(lengths, 23, 5)
    vad[3][:] = 0
    keys = synth.utt_keys(len(lengths))
    starts = np.cumsum([0] + lengths[:-1])
    kf.write_ark(str(tmp_path / "feats.ark"), [(k, x[s:s + n]) for k, s, n in zip(keys, starts, lengths)],
                 scp_path=str(tmp_path / "feats.scp"))
    kf.write_ark(str(tmp_path / "vad.ark"), [(k, v) for k, v in zip(keys, vad)], scp_path=str(tmp_path / "vad.scp"))
    exe = os.path.join(cvx.BIN_DIR, "nnet3-xvector-compute")
    cmd = [exe, "--use-gpu=no", "--min-chunk-size=25", "--chunk-size=300", "--output-node=tdnn6.affine",
           "--cmn-window=300", "--center=true", f"--vad-rspecifier=scp,s,cs:{tmp_path}/vad.scp", path,
           f"scp:{tmp_path}/feats.scp", f"ark:{tmp_path}/o.ark"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    assert "No features were judged as voiced" in r.stderr and "Done 3 utterances, failed for 1" in r.stderr
    got = dict(kf.read_ark(str(tmp_path / "o.ark")))
    pieces, kept = [], []
    for s, n, v in zip(starts, lengths, vad):
        seg = select_voiced(sliding_cmn(x[s:s + n], 300, True), v)
        pieces.append(seg); kept.append(len(seg))
    ref, ok = extract_all(Oracle(model, "tdnn6.affine", np.float64), {"input": np.concatenate(pieces)}, kept,
                          chunk_size=300, min_chunk_size=25)
    assert list(got) == [k for k, o in zip(keys, ok) if o]
    rr, cc = parity_metrics(np.stack([got[k] for k in got]), ref[ok])
    assert rr <= 1e-3 and cc >= 0.9999, (rr, cc)


def test_extract_raw_in_several_pieces(cvx, synth, model_dir, require_gpu):
    """a call large enough for the raw copies to be cut into several pieces and several device
    batches (copies overlapping compute): same embeddings as pre-processing everything on the host
    first, for every utterance, in both VAD and no-VAD form."""
    from oracle.prestep_oracle import select_voiced, sliding_cmn
    ps = importlib.import_module(PKG + ".prestep")
    path, _ = model_dir("v2_xvector")
    ex = cvx.Extractor(cvx.Model(path, "tdnn6.affine"), precision="fp16", max_batch_rows=131072)
    rng = np.random.default_rng(9)
    lengths = rng.integers(2000, 30000, 48).tolist()                # ~770 k raw rows: 3+ pieces, 4+ device batches
    x, vad = _data(lengths, 23, 3, voiced_p=0.8)
    for use_vad in (True, False):
        v = vad if use_vad else None
        emb, ok = ps.extract_raw(ex, {"input": x}, lengths, v, {"input": 300}, chunk_size=10000, min_chunk_size=25)
        pieces, kept, r = [], [], 0
        for T, m in zip(lengths, vad):
            seg = sliding_cmn(x[r:r + T], 300, True)
            seg = select_voiced(seg, m) if use_vad else seg
            pieces.append(seg); kept.append(len(seg)); r += T
        ref, ok_ref = ex.extract({"input": np.concatenate(pieces)}, kept, chunk_size=10000, min_chunk_size=25)
        assert ok.all() and ok_ref.all()
        assert np.max(np.abs(emb - ref)) <= 2e-3 * np.max(np.abs(ref)), use_vad
