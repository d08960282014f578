"""The drop-in program's job fan-out on a GPU box: --shard=J/N, reader lanes (--num-gpus with scp
input), automatic device selection for concurrent jobs, VAD looked up by key, and the
--pad-input / --chunk-size < --min-chunk-size case."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL_REL, TOL_COS = 1e-3, 0.9999


def _same(x, y):
    """the same utterance computed in two runs whose device batches differ: the fused pooling sums in another
    order (~1e-6) and the batch may run the other activation form (dense / sparse, DESIGN.md section 6: ~1e-4)."""
    return np.max(np.abs(np.asarray(x) - np.asarray(y))) <= 5e-4 * np.max(np.abs(y))


@pytest.fixture(scope="module")
def data_dir(tmp_path_factory, cvx, kf, synth, model_dir):
    """A small Kaldi-style data dir: feats.ark/scp (23-dim), utt2num_frames, vad, the v2 net."""
    d = tmp_path_factory.mktemp("data")
    path, model = model_dir("v2_xvector")
    rng = np.random.default_rng(17)
    lengths = rng.integers(120, 900, size=40).tolist()
    keys = synth.utt_keys(len(lengths))
    x = synth.make_features(lengths, 23, 4)
    starts = np.concatenate([[0], np.cumsum(lengths)[:-1]])
    mats = [(k, x[s:s + n]) for k, s, n in zip(keys, starts, lengths)]
    kf.write_ark(str(d / "feats.ark"), mats, scp_path=str(d / "feats.scp"))
    with open(d / "utt2num_frames", "w") as f:
        for k, n in zip(keys, lengths):
            f.write(f"{k} {n}\n")
    return {"dir": d, "model": path, "py_model": model, "keys": keys, "lengths": lengths, "x": x, "starts": starts}


def _run(cvx, args, **kw):
    exe = os.path.join(cvx.BIN_DIR, "nnet3-xvector-compute")
    return subprocess.run([exe, "--min-chunk-size=25", "--chunk-size=300", "--output-node=tdnn6.affine", *args],
                          capture_output=True, text=True, timeout=600, **kw)


def test_shards_partition_the_table(cvx, kf, data_dir):
    """--shard=J/3 over an scp (frame counts from the matrix headers) and over an archive with
    --utt2num-frames: the three jobs' outputs are disjoint, their union is the unsharded output, and
    shard J holds exactly the utterances cvx_lpt_shards assigns to it."""
    d = data_dir["dir"]
    r = _run(cvx, [data_dir["model"], f"scp:{d}/feats.scp", f"ark:{d}/all.ark"])
    assert r.returncode == 0, r.stderr
    whole = dict(kf.read_ark(f"{d}/all.ark"))
    assert list(whole) == data_dir["keys"]
    want = cvx.lpt_shards(data_dir["lengths"], 3)
    for mode in ("scp", "ark"):
        seen = {}
        for j in (1, 2, 3):
            opts, src = ([], f"scp:{d}/feats.scp") if mode == "scp" else ([f"--utt2num-frames={d}/utt2num_frames"], f"ark:{d}/feats.ark")
            r = _run(cvx, [f"--shard={j}/3", *opts, data_dir["model"], src, f"ark,scp:{d}/x.{j}.ark,{d}/x.{j}.scp"])
            assert r.returncode == 0, r.stderr
            got = dict(kf.read_scp(f"{d}/x.{j}.scp"))
            assert list(got) == [k for k, s in zip(data_dir["keys"], want) if s == j - 1], (mode, j)
            for k, v in got.items():
                assert k not in seen
                seen[k] = v
        assert set(seen) == set(whole)
        for k in whole:
            assert _same(seen[k], whole[k]), k
    # an archive cannot be balanced without frame counts
    r = _run(cvx, ["--shard=1/2", data_dir["model"], f"ark:{d}/feats.ark", f"ark:{d}/none.ark"])
    assert r.returncode != 0 and "--utt2num-frames" in r.stderr


def test_reader_lanes_keep_table_order(cvx, kf, data_dir, monkeypatch):
    """--num-gpus=3 with an scp input: three reader lanes (own engine, own share of the table) whose
    results the writer merges back into table order; CVX_WRAP_DEVICES lets a one-GPU box run it."""
    d = data_dir["dir"]
    monkeypatch.setenv("CVX_WRAP_DEVICES", "1")
    r1 = _run(cvx, [data_dir["model"], f"scp:{d}/feats.scp", f"ark:{d}/one.ark"])
    r3 = _run(cvx, ["--num-gpus=3", "--batch-rows=2000", data_dir["model"], f"scp:{d}/feats.scp", f"ark:{d}/three.ark"])
    assert r1.returncode == 0 and r3.returncode == 0, r3.stderr
    assert "3 reader lane(s)" in r3.stderr and "Done 40 utterances, failed for 0" in r3.stderr
    a, b = list(kf.read_ark(f"{d}/one.ark")), list(kf.read_ark(f"{d}/three.ark"))
    assert [k for k, _ in a] == [k for k, _ in b] == data_dir["keys"]
    assert all(_same(x, y) for (_, x), (_, y) in zip(a, b))
    # lanes inside a shard, and an archive input (one sequential stream: batches go to the engines in turn)
    r = _run(cvx, ["--num-gpus=2", "--shard=2/2", data_dir["model"], f"scp:{d}/feats.scp", f"ark:{d}/s22.ark"])
    assert r.returncode == 0, r.stderr
    want = cvx.lpt_shards(data_dir["lengths"], 4)
    assert [k for k, _ in kf.read_ark(f"{d}/s22.ark")] == [k for k, s in zip(data_dir["keys"], want) if s // 2 == 1]
    r = _run(cvx, ["--num-gpus=2", "--batch-rows=2000", data_dir["model"], f"ark:{d}/feats.ark", f"ark:{d}/rr.ark"])
    assert r.returncode == 0 and "reader lane" not in r.stderr, r.stderr
    assert [k for k, _ in kf.read_ark(f"{d}/rr.ark")] == data_dir["keys"]


def test_concurrent_jobs_pick_devices_through_lock_files(cvx, data_dir, tmp_path, monkeypatch):
    """the recipe's GPU branch: nj background jobs, --use-gpu=yes, no device hint
    (extract_xvectors_new.sh:83-89).  Each job takes a lock; with G visible GPUs the first G jobs
    get a free device each, the rest share."""
    d = data_dir["dir"]
    monkeypatch.setenv("CVX_LOCK_DIR", str(tmp_path))
    monkeypatch.delenv("CVX_DEVICE", raising=False)
    exe = os.path.join(cvx.BIN_DIR, "nnet3-xvector-compute")
    n = cvx.device_count()
    procs = [subprocess.Popen([exe, "--use-gpu=yes", "--min-chunk-size=25", "--chunk-size=300", "--output-node=tdnn6.affine",
                               f"--shard={j}/{n + 1}", data_dir["model"], f"scp:{d}/feats.scp", f"ark:{tmp_path}/j{j}.ark"],
                              stderr=subprocess.PIPE, text=True) for j in range(1, n + 2)]
    logs = [p.communicate(timeout=600)[1] for p in procs]
    assert all(p.returncode == 0 for p in procs), logs
    picked = ["".join(ln for ln in lg.splitlines() if "GPU" in ln and ("Selected" in ln or "sharing" in ln)) for lg in logs]
    assert all(picked), logs
    # jobs that overlapped in time never hold the same device exclusively; at least one line of each kind
    # appears when all n + 1 jobs overlap, which start-up times make the common case -- accept either
    assert sum("Selected GPU" in p for p in picked) >= min(n, 1)


def test_vad_is_looked_up_by_key(cvx, kf, synth, data_dir, tmp_path):
    """--vad-rspecifier replaces `select-voiced-frames ark:- scp,s,cs:vad.scp`: a superset VAD table works,
    an utterance without a VAD entry is skipped with a warning instead of aborting the job."""
    from oracle.nnet3_oracle import Oracle, extract_all, parity_metrics
    d = data_dir["dir"]
    keys, lengths, x, starts = data_dir["keys"][:6], data_dir["lengths"][:6], data_dir["x"], data_dir["starts"]
    rng = np.random.default_rng(5)
    vads = {k: (rng.random(n) > 0.3).astype(np.float32) for k, n in zip(keys, lengths)}
    table = [("aaa_extra", np.ones(7, np.float32))] + [(k, vads[k]) for k in keys if k != keys[2]] + [("zzz_extra", np.ones(3, np.float32))]
    kf.write_ark(str(tmp_path / "vad.ark"), sorted(table), scp_path=str(tmp_path / "vad.scp"))
    kf.write_ark(str(tmp_path / "f.ark"), [(k, x[s:s + n]) for k, s, n in zip(keys, starts, lengths)])
    r = _run(cvx, [f"--vad-rspecifier=scp:{tmp_path}/vad.scp", data_dir["model"], f"ark:{tmp_path}/f.ark", f"ark:{tmp_path}/o.ark"])
    assert r.returncode == 0, r.stderr
    assert f"No VAD entry for utterance {keys[2]}" in r.stderr and "Done 5 utterances, failed for 1" in r.stderr
    got = dict(kf.read_ark(f"{tmp_path}/o.ark"))
    assert list(got) == [k for k in keys if k != keys[2]]
    orc = Oracle(data_dir["py_model"], "tdnn6.affine", np.float64)
    for k, s, n in zip(keys, starts, lengths):
        if k == keys[2]:
            continue
        sel = x[s:s + n][vads[k] != 0]
        ref, ok = extract_all(orc, {"input": sel}, [len(sel)], chunk_size=300, min_chunk_size=25)
        rr, cc = parity_metrics(got[k][None], ref)
        assert rr <= TOL_REL and cc >= TOL_COS, (k, rr, cc)


def test_pad_input_with_chunk_size_below_min_chunk_size(cvx, kf, synth, data_dir, tmp_path):
    """--pad-input=true --chunk-size=10 --min-chunk-size=25: EVERY chunk is padded to 25 rows, far more
    rows than the utterance has (this overflowed the pinned batch in round 1)."""
    from oracle.nnet3_oracle import Oracle, extract_all, parity_metrics
    keys, lengths, x, starts = data_dir["keys"][:3], data_dir["lengths"][:3], data_dir["x"], data_dir["starts"]
    lengths = [lengths[0], 7, 3000]
    x = synth.make_features(lengths, 23, 77)
    starts = np.concatenate([[0], np.cumsum(lengths)[:-1]])
    kf.write_ark(str(tmp_path / "f.ark"), [(k, x[s:s + n]) for k, s, n in zip(keys, starts, lengths)])
    exe = os.path.join(cvx.BIN_DIR, "nnet3-xvector-compute")
    r = subprocess.run([exe, "--pad-input=true", "--chunk-size=10", "--min-chunk-size=25", "--batch-rows=4096",
                        "--output-node=tdnn6.affine", data_dir["model"], f"ark:{tmp_path}/f.ark", f"ark:{tmp_path}/o.ark"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    assert "every chunk is padded" in r.stderr
    got = np.stack([v for _, v in kf.read_ark(f"{tmp_path}/o.ark")])
    ref, ok = extract_all(Oracle(data_dir["py_model"], "tdnn6.affine", np.float64), {"input": x}, lengths, chunk_size=10,
                          min_chunk_size=25, pad_input=True)
    assert ok.all() and got.shape == ref.shape
    rr, cc = parity_metrics(got, ref)
    assert rr <= TOL_REL and cc >= TOL_COS, (rr, cc)
