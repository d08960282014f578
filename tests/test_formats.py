"""Kaldi formats: Python writer/reader round trips and the C++ reader (through the
C ABI's cvx_model_load) on files the Python side wrote.  CPU only."""
import os
import struct

import numpy as np
import pytest


def test_matrix_vector_roundtrip_binary_and_text(kf, tmp_path):
    rng = np.random.default_rng(0)
    items = [("utt1", rng.standard_normal((7, 23)).astype(np.float32)), ("utt2", rng.standard_normal(5).astype(np.float32)),
             ("u3", np.zeros((0, 0), np.float32))]
    for binary in (True, False):
        p = str(tmp_path / f"a{int(binary)}.ark")
        kf.write_ark(p, items[:2], binary=binary, scp_path=p + ".scp")
        back = list(kf.read_ark(p))
        assert [k for k, _ in back] == ["utt1", "utt2"]
        for (_, a), (_, b) in zip(items, back):
            assert np.array_equal(a.reshape(b.shape), b)
        via_scp = list(kf.read_scp(p + ".scp"))
        assert np.array_equal(via_scp[0][1], items[0][1])
        assert np.array_equal(via_scp[1][1], items[1][1])


def test_binary_layout_is_kaldi(kf, tmp_path):
    p = str(tmp_path / "x.ark")
    kf.write_ark(p, [("k", np.array([[1.0, 2.0]], np.float32))])
    raw = open(p, "rb").read()
    assert raw == b"k \0BFM \x04" + struct.pack("<i", 1) + b"\x04" + struct.pack("<i", 2) + struct.pack("<ff", 1.0, 2.0)
    kf.write_ark(p, [("k", np.array([1.5], np.float32))])
    assert open(p, "rb").read() == b"k \0BFV \x04" + struct.pack("<i", 1) + struct.pack("<f", 1.5)


def test_compressed_matrix_formats(kf):
    # CM2 (uint16) and CM3 (uint8) global-header formats, CM (per-column header)
    rows, cols = 3, 2
    hdr = struct.pack("<ffii", -1.0, 2.0, rows, cols)
    cm2 = b"CM2 " + hdr + np.array([[0, 65535], [32768, 1], [100, 200]], "<u2").tobytes()
    r = kf._Reader(cm2)
    m = kf.read_vector_or_matrix(r, True)
    assert m.shape == (3, 2) and abs(m[0, 0] + 1.0) < 1e-6 and abs(m[0, 1] - 1.0) < 1e-6
    cm3 = b"CM3 " + hdr + np.array([[0, 255], [128, 1], [100, 200]], np.uint8).tobytes()
    m = kf.read_vector_or_matrix(kf._Reader(cm3), True)
    assert abs(m[0, 1] - 1.0) < 1e-6 and abs(m[1, 0] - (-1 + 2 * 128 / 255)) < 1e-6
    colhdr = np.array([[0, 16384, 49152, 65535]] * cols, "<u2").tobytes()
    data = np.array([[0, 64, 192], [255, 128, 32]], np.uint8).tobytes()         # column-major bytes
    m = kf.read_vector_or_matrix(kf._Reader(b"CM " + hdr + colhdr + data), True)
    p0, p25, p75, p100 = [-1 + 2 * v / 65535 for v in (0, 16384, 49152, 65535)]
    assert np.allclose(m[:, 0], [p0, p25, p75], atol=1e-6)
    assert np.allclose(m[:, 1], [p100, p25 + (p75 - p25) * 0.5, p0 + (p25 - p0) * 0.5], atol=1e-6)


@pytest.mark.parametrize("binary", [True, False])
def test_nnet3_roundtrip(kf, synth, tmp_path, binary):
    m = synth.make_tiny_model(two_branch=True)
    p = str(tmp_path / "final.raw")
    kf.write_nnet3(p, m, binary)
    head = open(p, "rb").read(11)
    assert head.startswith(b"\0B<Nnet3> \n") if binary else head.startswith(b"<Nnet3> \n")
    m2 = kf.read_nnet3(p)
    assert m2.config_lines == m.config_lines
    assert list(m2.components) == list(m.components)
    for k, c in m.components.items():
        for fk, fv in c.fields.items():
            got = m2.components[k].fields[fk]
            if isinstance(fv, np.ndarray):
                assert np.array_equal(np.asarray(got).reshape(fv.shape), fv), (k, fk)
            elif isinstance(fv, bool):
                assert got == fv
            elif fk in kf._INT_FIELDS:
                assert int(got) == fv
            else:
                assert abs(float(got) - fv) <= 1e-6 * max(abs(fv), 1e-30)


@pytest.mark.parametrize("binary", [True, False])
def test_cpp_loader_reads_python_written_models(cvx, synth, model_dir, binary):
    """text and binary, every topology's graph: inputs, contexts, MAC counts."""
    for topology in synth.TOPOLOGIES if binary else ["tiny2", "v2_xvector"]:
        if topology.startswith("tiny"):
            path, _ = model_dir(topology, binary)
            mod = cvx.Model(path, "tdnn6.affine")
            assert (mod.info.left_context, mod.info.right_context) == ((5, 4) if topology == "tiny2" else (4, 4))
            continue
        path, _ = model_dir(topology, binary)
        mod = cvx.Model(path, synth.EMBEDDING_NODE[topology])
        assert mod.info.mac_per_frame == synth.MAC_PER_FRAME[topology]
        assert mod.info.mac_per_chunk == (1024 * 512 if topology == "v5_expand" else synth.SEGMENT_MAC_PER_CHUNK)
        assert mod.info.embedding_dim == 512 and mod.info.pooled_dim == (1024 if topology == "v5_expand" else 3000)
        # "expand": the bottleneck (left 13, right 7) is spliced {-2..2} into tdnn1_share, then {-2,0,2}, {-3,0,3}
        want_ctx = (7, 7) if topology.startswith(("v2", "v3")) else (20, 14) if topology == "v5_expand" else (13, 7)
        assert (mod.info.left_context, mod.info.right_context) == want_ctx
        assert mod.input_names == (["input", "input_am"] if topology == "v4_two_input" else ["input"])
        assert mod.input_dims == ([23, 40] if topology == "v4_two_input" else [23])


def test_cpp_loader_prunes_to_requested_output(cvx, synth, model_dir):
    path, _ = model_dir("v5_cvector")
    d = cvx.Model(path, "tdnn6_xvec.affine").describe()
    assert "_am" not in d and "tdnn7" not in d          # multitask head and orphans are gone
    assert "tdnn5.batchnorm 1950->128" in d             # phonetic bottleneck kept
    assert "taps={L3@0,L8@0}" in d                      # Append(tdnn4_xvec, tdnn5.batchnorm): no copy, two sources
    # a deeper output node: two segment layers
    d2 = cvx.Model(path, "tdnn7_xvec.affine").describe()
    assert "segment[0] tdnn6_xvec.batchnorm 3000->512 relu bn" in d2 and "segment[1] tdnn7_xvec.affine 512->512" in d2


def test_cpp_loader_model_through_pipe_and_errors(cvx, synth, model_dir, tmp_path):
    path, model = model_dir("tiny")
    assert cvx.Model(f"cat {path} |", "tdnn6.affine").info.embedding_dim == 8
    with pytest.raises(cvx.CvxError) as e:
        cvx.Model(path)                                  # output = log-softmax head, frame level
    assert e.value.code == -4
    with pytest.raises(cvx.CvxError) as e:
        cvx.Model(str(tmp_path / "missing.raw"), "tdnn6.affine")
    assert e.value.code == -2
    bad = tmp_path / "bad.raw"
    bad.write_bytes(b"\0B<Nnet2> ")
    with pytest.raises(cvx.CvxError):
        cvx.Model(str(bad), "x")
    with pytest.raises(cvx.CvxError) as e:
        cvx.Model(path, "no_such_node")
    assert e.value.code in (-3, -4)


def test_cpp_loader_refuses_batchnorm_without_stats(cvx, kf, synth, tmp_path):
    m = synth.make_tiny_model()
    m.components["tdnn2.batchnorm"].fields["Count"] = 0.0
    p = str(tmp_path / "c0.raw")
    kf.write_nnet3(p, m, True)
    with pytest.raises(cvx.CvxError, match="Count"):
        cvx.Model(p, "tdnn6.affine")


def test_cpp_loader_tolerates_unknown_trailing_fields(cvx, kf, synth, tmp_path):
    """version-dependent extra tokens inside component blobs are skipped generically."""
    m = synth.make_tiny_model()
    f = m.components["tdnn1.affine"].fields
    f["UpdateCount"] = 12.5
    f["IsGradient"] = False
    f["MaxChangeScaleStats"] = 0.25
    f["OrthonormalConstraint"] = 0.0
    m.components["tdnn1.relu"].fields["OderivRms"] = np.zeros(16, np.float64)
    for binary in (True, False):
        p = str(tmp_path / f"x{int(binary)}.raw")
        kf.write_nnet3(p, m, binary)
        assert cvx.Model(p, "tdnn6.affine").info.embedding_dim == 8


def test_forward_only_taps_are_refused_not_clamped(cvx, kf, synth, tmp_path):
    """a layer whose taps all look forward is computable before t = 0 under nnet3's rules; tapped again
    backwards, upstream uses those frames.  The packed layout cannot hold them: the loader refuses the
    net (CVX_ERR_UNSUPPORTED) instead of silently pooling fewer frames (found by tests/test_gpu_fuzz.py)."""
    rng = np.random.default_rng(1)
    b = synth._Builder(rng)
    b.input("input", 7)
    b.relu_batchnorm_layer("a", 16, "input@1")                       # valid t in [-1, T-2]
    b.relu_batchnorm_layer("b", 16, "Append(a@-1,a@1)")              # needs a(-1) at t = 0: computable upstream
    b.stats_layer("stats", 0, 10000, "b")
    b.relu_batchnorm_layer("seg", 8)
    b.output_layer("output", 4)
    kf.write_nnet3(str(tmp_path / "fwd.raw"), b.model(), True)
    with pytest.raises(cvx.CvxError) as e:
        cvx.Model(str(tmp_path / "fwd.raw"), "seg.affine")
    assert e.value.code == -4 and "computable on frames [0, T-1-(2)]" in str(e.value), str(e.value)
    # with a tap at 0 the first layer is computable exactly on the input's frames and all is well:
    # a: left 1, right 0; b taps a@-1, a@1 -> left 2, right 1
    b = synth._Builder(np.random.default_rng(1))
    b.input("input", 7)
    b.relu_batchnorm_layer("a", 16, "Append(input@-1,input@0)")
    b.relu_batchnorm_layer("b", 16, "Append(a@-1,a@1)")
    b.stats_layer("stats", 0, 10000, "b")
    b.relu_batchnorm_layer("seg", 8)
    b.output_layer("output", 4)
    kf.write_nnet3(str(tmp_path / "bwd.raw"), b.model(), True)
    m = cvx.Model(str(tmp_path / "bwd.raw"), "seg.affine")
    assert (m.info.left_context, m.info.right_context) == (2, 1)
