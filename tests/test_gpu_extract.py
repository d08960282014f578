"""End-to-end parity of the CUDA path (through the C ABI) against the oracle."""
import numpy as np
import pytest

from conftest import embedding_node

pytestmark = pytest.mark.gpu

# north-star tolerance (BASELINE.json): max|a-b|/max|b| <= 1e-3 and cosine >= 0.9999
TOL_REL, TOL_COS = 1e-3, 0.9999


def _oracle(model, node, dtype=np.float64):
    from oracle.nnet3_oracle import Oracle
    return Oracle(model, node, dtype)


def _inputs(synth, mod, lengths, seed):
    return {n: synth.make_features(lengths, d, seed + i) for i, (n, d) in enumerate(zip(mod.input_names, mod.input_dims))}


@pytest.mark.parametrize("precision,rel,cos", [("fp32", 2e-5, 0.999999), ("fp16", TOL_REL, TOL_COS)])
@pytest.mark.parametrize("topology", ["tiny", "tiny2", "v2_xvector", "v3_multitask", "v4_cvector", "v4_two_input",
                                      "v4_pa_wo_pretrain", "v5_cvector"])
def test_extract_matches_oracle(cvx, synth, model_dir, require_gpu, topology, precision, rel, cos):
    from oracle.nnet3_oracle import extract_all, parity_metrics
    path, model = model_dir(topology)
    node = embedding_node(synth, topology)
    mod = cvx.Model(path, node)
    lengths = [25, 300, 57, 1234, 100, 31, 640]
    feats = _inputs(synth, mod, lengths, 11)
    ex = cvx.Extractor(mod, precision=precision)
    got, ok = ex.extract(feats, lengths, chunk_size=500, min_chunk_size=25)
    ref, ok_ref = extract_all(_oracle(model, node), feats, lengths, chunk_size=500, min_chunk_size=25)
    assert ok.tolist() == ok_ref.tolist()
    r, c = parity_metrics(got[ok], ref[ok])
    assert r <= rel and c >= cos, f"{topology}/{precision}: rel {r} cos {c}"


def test_chunking_semantics(cvx, synth, model_dir, require_gpu):
    """short tails skipped (old Kaldi) vs padded (new Kaldi), too-short utterances fail,
    chunk-size -1, chunk boundaries exactly at multiples."""
    from oracle.nnet3_oracle import extract_all, parity_metrics
    path, model = model_dir("v2_xvector")
    mod = cvx.Model(path, "tdnn6.affine")
    ex = cvx.Extractor(mod, precision="fp32")
    lengths = [10, 24, 25, 200, 210, 400, 401, 0, 15]
    feats = _inputs(synth, mod, lengths, 3)
    for chunk, minc, pad in [(200, 25, False), (200, 25, True), (-1, 25, False), (100, 100, False), (100, 100, True)]:
        got, ok = ex.extract(feats, lengths, chunk_size=chunk, min_chunk_size=minc, pad_input=pad)
        ref, ok_ref = extract_all(_oracle(model, "tdnn6.affine"), feats, lengths, chunk_size=chunk,
                                  min_chunk_size=minc, pad_input=pad)
        assert ok.tolist() == ok_ref.tolist(), (chunk, minc, pad)
        r, c = parity_metrics(got[ok], ref[ok])
        assert r <= 2e-5, (chunk, minc, pad, r)
        assert np.all(got[~ok] == 0)


@pytest.mark.parametrize("mode", ["dense-standalone", "dense-fused", "default"])
def test_ragged_batch_is_bit_identical_to_single(cvx, synth, model_dir, require_gpu, monkeypatch, mode):
    """per-chunk results must not depend on what else is in the packed batch, nor on
    how the engine cuts device batches (guards the edge handling of the packed layout).
    With the stand-alone pooling kernels (CVX_FUSED_POOL=0) this holds bit for bit; the fused
    pooling epilogue sums in 32-row groups aligned to the batch, so there it holds to ~1e-6.  Both with
    the dense form pinned (CVX_FORWARD_OFFSETS=0): by default a device batch runs the sparse
    (offset-forwarding) form unless it holds a chunk that pools < 32 frames -- the 45-frame utterance
    here -- so a chunk's rounding depends on its batch; the two forms agree to ~1e-4 of the 1e-3 tolerance."""
    if mode != "default":
        monkeypatch.setenv("CVX_FORWARD_OFFSETS", "0")
        monkeypatch.setenv("CVX_FUSED_POOL", "1" if mode == "dense-fused" else "0")
    if mode == "dense-standalone":
        same = np.array_equal
    elif mode == "dense-fused":
        same = lambda x, y: np.allclose(x, y, rtol=2e-5, atol=2e-5)
    else:
        same = lambda x, y: np.max(np.abs(x - y)) <= 5e-4 * np.max(np.abs(y))
    path, model = model_dir("v5_cvector")
    mod = cvx.Model(path, "tdnn6_xvec.affine")
    lengths = [300, 45, 1000, 77, 512, 128]
    feats = _inputs(synth, mod, lengths, 21)
    big = cvx.Extractor(mod, precision="fp16")
    small = cvx.Extractor(mod, precision="fp16", max_batch_rows=512)   # forces several device batches
    a, _ = big.extract(feats, lengths, chunk_size=400, min_chunk_size=25)
    b, _ = small.extract(feats, lengths, chunk_size=400, min_chunk_size=25)
    assert same(a, b)
    base = 0
    for u, T in enumerate(lengths):
        one, _ = big.extract({k: v[base:base + T] for k, v in feats.items()}, [T], chunk_size=400, min_chunk_size=25)
        assert same(one[0], a[u]), u
        base += T


def test_sparse_and_dense_forms_agree(cvx, synth, model_dir, require_gpu, monkeypatch):
    """offset forwarding (hidden activations stored as relu(.) without the BatchNorm offset, the offset folded
    into the consumers' biases) is an exact algebraic rewrite: in fp32 the two forms agree to rounding, in fp16
    to a fraction of the tolerance; utterances that pool < 32 frames keep the dense form either way."""
    from oracle.nnet3_oracle import Oracle, extract_all, parity_metrics
    path, model = model_dir("v5_cvector")
    mod = cvx.Model(path, "tdnn6_xvec.affine")
    lengths = [300, 1000, 640, 2000]
    feats = _inputs(synth, mod, lengths, 5)
    ref, _ = extract_all(Oracle(model, "tdnn6_xvec.affine", np.float64), feats, lengths, chunk_size=500, min_chunk_size=25)
    out = {}
    for prec in ("fp32", "fp16"):
        for fwd in ("0", "1"):
            monkeypatch.setenv("CVX_FORWARD_OFFSETS", fwd)
            out[prec, fwd], _ = cvx.Extractor(mod, precision=prec).extract(feats, lengths, chunk_size=500, min_chunk_size=25)
    assert np.max(np.abs(out["fp32", "0"] - out["fp32", "1"])) <= 2e-5 * np.max(np.abs(ref))
    assert not np.array_equal(out["fp16", "0"], out["fp16", "1"])          # the sparse form really ran
    for fwd in ("0", "1"):
        r, c = parity_metrics(out["fp16", fwd], ref)
        print(f"forward_offsets={fwd}: rel {r:.3e} cos {c:.8f}")
        assert r <= TOL_REL and c >= TOL_COS, (fwd, r, c)


@pytest.mark.parametrize("name", ["v2_small", "v5_small"])
@pytest.mark.parametrize("pad", [False, True])
def test_golden_fixture(cvx, require_gpu, name, pad):
    """committed golden vectors (tests/golden, made by tests/golden/make_golden.py)."""
    import os
    from oracle.nnet3_oracle import parity_metrics
    g = os.path.join(os.path.dirname(__file__), "golden")
    z = np.load(os.path.join(g, f"{name}.npz"))
    mod = cvx.Model(os.path.join(g, f"{name}_final.raw"), str(z["output_node"]))
    ex = cvx.Extractor(mod, precision="fp16")
    got, ok = ex.extract({"input": z["feats"]}, z["lengths"].tolist(), chunk_size=int(z["chunk_size"]),
                         min_chunk_size=int(z["min_chunk_size"]), pad_input=pad)
    want, want_ok = (z["embeddings_pad_input"], z["ok_pad_input"]) if pad else (z["embeddings"], z["ok"])
    assert ok.tolist() == want_ok.tolist()
    r, c = parity_metrics(got[ok], want[ok])
    assert r <= TOL_REL and c >= TOL_COS, (r, c)


@pytest.mark.parametrize("name", ["v2_small", "v5_small"])
def test_cli_drop_in(cvx, kf, require_gpu, tmp_path, name):
    """the drop-in program with the recipe's argument shapes
    (extract_xvectors_new.sh:84-93): model through a pipe, features through a pipe,
    ark,scp output; embeddings match the golden vectors and the scp offsets resolve."""
    import os
    import subprocess
    from oracle.nnet3_oracle import parity_metrics
    g = os.path.join(os.path.dirname(__file__), "golden")
    z = np.load(os.path.join(g, f"{name}.npz"))
    exe = os.path.join(cvx.BIN_DIR, "nnet3-xvector-compute")
    ark, scp = str(tmp_path / "xvector.1.ark"), str(tmp_path / "xvector.1.scp")
    cmd = [exe, "--use-gpu=no", f"--min-chunk-size={int(z['min_chunk_size'])}", f"--chunk-size={int(z['chunk_size'])}",
           f"--nnet-config={g}/{name}_extract.config", f"cat {g}/{name}_final.raw |",
           f"ark:cat {g}/{name}_feats.ark |", f"ark,scp:{ark},{scp}"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    assert "Done 5 utterances, failed for 2" in r.stderr, r.stderr
    assert "real-time factor assuming 100 frames/sec" in r.stderr
    got = dict(kf.read_scp(scp))
    keys = [f"utt{i:07d}" for i in range(len(z["lengths"]))]
    assert list(got) == [k for k, o in zip(keys, z["ok"]) if o]            # input order, failed ones absent
    a = np.stack([got[k] for k in got])
    rr, cc = parity_metrics(a, z["embeddings"][z["ok"]])
    assert rr <= TOL_REL and cc >= TOL_COS, (rr, cc)
    # text output + --pad-input + --output-node instead of the config file
    cmd = [exe, "--pad-input=true", f"--min-chunk-size={int(z['min_chunk_size'])}", f"--chunk-size={int(z['chunk_size'])}",
           f"--output-node={z['output_node']}", f"{g}/{name}_final.raw", f"ark:{g}/{name}_feats.ark", f"ark,t:{tmp_path}/x.txt"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    got = dict(kf.read_ark(f"{tmp_path}/x.txt"))
    a = np.stack([got[k] for k in got])
    rr, cc = parity_metrics(a, z["embeddings_pad_input"][z["ok_pad_input"]])
    assert rr <= TOL_REL and cc >= TOL_COS, (rr, cc)


def test_cli_two_input(cvx, kf, synth, model_dir, require_gpu, tmp_path):
    """nnet3-cvector-compute-multiple-input --name=a,b 2 <nnet> <feats1> <feats2> <wspec>
    (extract_cvectors_with_am.sh:99-110), names given in swapped order."""
    import os
    import subprocess
    from oracle.nnet3_oracle import Oracle, extract_all, parity_metrics
    path, model = model_dir("v4_two_input")
    lengths = [120, 333, 40]
    f23 = synth.make_features(lengths, 23, 1)
    f40 = synth.make_features(lengths, 40, 2)
    starts = np.cumsum([0] + lengths[:-1])
    keys = synth.utt_keys(3)
    kf.write_ark(str(tmp_path / "a.ark"), [(k, f23[s:s + n]) for k, s, n in zip(keys, starts, lengths)])
    kf.write_ark(str(tmp_path / "b.ark"), [(k, f40[s:s + n]) for k, s, n in zip(keys, starts, lengths)], binary=False)
    exe = os.path.join(cvx.BIN_DIR, "nnet3-cvector-compute-multiple-input")
    cmd = [exe, "--use-gpu=no", "--min-chunk-size=25", "--chunk-size=200", "--output-node=tdnn6_xvec.affine",
           "--name=input_am,input", "2", path, f"ark:{tmp_path}/b.ark", f"ark:{tmp_path}/a.ark", f"ark:{tmp_path}/o.ark"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    got = np.stack([v for _, v in kf.read_ark(str(tmp_path / "o.ark"))])
    ref, ok = extract_all(Oracle(model, "tdnn6_xvec.affine", np.float64), {"input": f23, "input_am": f40}, lengths,
                          chunk_size=200, min_chunk_size=25)
    rr, cc = parity_metrics(got, ref[ok])
    assert rr <= TOL_REL and cc >= TOL_COS, (rr, cc)


def test_full_size_properties(cvx, synth, model_dir, require_gpu):
    """size-independent checks at a BASELINE-sized slice (config 4 style long utterances):
    (1) the frame-weighted average of per-chunk embeddings reproduces the utterance
    embedding; (2) duplicating an utterance gives identical embeddings; (3) device-input
    and host-input paths agree bit for bit."""
    path, model = model_dir("v5_cvector")
    mod = cvx.Model(path, "tdnn6_xvec.affine")
    lengths = [30000, 12345, 30000, 10000, 1000]
    feats = _inputs(synth, mod, lengths, 5)
    feats["input"][42345:72345] = feats["input"][0:30000]       # utt 2 := utt 0
    ex = cvx.Extractor(mod, precision="fp16", max_batch_rows=16384)
    emb, ok = ex.extract(feats, lengths, chunk_size=10000, min_chunk_size=25)
    assert ok.all()
    assert np.allclose(emb[0], emb[2], rtol=2e-5, atol=2e-5)   # bit-equal only with CVX_FUSED_POOL=0
    n_chunks = sum(len(cvx.chunk_plan(T, 10000, 25, False)) for T in lengths)
    ce = ex.last_chunk_embeddings(n_chunks)
    i = 0
    for u, T in enumerate(lengths):
        plan = cvx.chunk_plan(T, 10000, 25, False)
        w = np.array([n for _, n in plan], np.float64)
        avg = (ce[i:i + len(plan)].astype(np.float64) * w[:, None]).sum(0) / w.sum()
        assert np.allclose(avg, emb[u], rtol=1e-5, atol=1e-5)
        i += len(plan)


def test_fused_pooling_matches_standalone_kernels(cvx, synth, model_dir, require_gpu, monkeypatch):
    """the pooled layer's fused-reduction epilogue vs tdnn5 store + stats_pool kernels: same
    activations (rounded to the operand type either way), different summation order."""
    path, model = model_dir("v2_xvector")
    mod = cvx.Model(path, "tdnn6.affine")
    lengths = [25, 26, 40, 57, 300, 1000, 33, 64, 31, 2048]      # chunk edges at every alignment mod 32
    feats = _inputs(synth, mod, lengths, 77)
    out = {}
    for fused in ("0", "1"):
        monkeypatch.setenv("CVX_FUSED_POOL", fused)
        ex = cvx.Extractor(mod, precision="fp16")
        out[fused], ok = ex.extract(feats, lengths, chunk_size=100, min_chunk_size=25)
        assert ok.all()
    assert np.max(np.abs(out["0"] - out["1"])) <= 2e-5 * np.max(np.abs(out["0"]))


def test_model_from_plan_abi(cvx, synth, require_gpu):
    """level-0 ABI: a host that already holds the network hands over a cvx_plan
    (cvx_model_from_plan) instead of a file; same embeddings as the file route."""
    import ctypes as C
    from oracle.nnet3_oracle import Oracle, extract_all, parity_metrics

    class Tap(C.Structure):
        _fields_ = [("src", C.c_int32), ("offset", C.c_int32)]

    class Layer(C.Structure):
        _fields_ = [("name", C.c_char_p), ("num_taps", C.c_int32), ("taps", C.POINTER(Tap)), ("in_dim", C.c_int32),
                    ("out_dim", C.c_int32), ("weight", C.POINTER(C.c_float)), ("bias", C.POINTER(C.c_float)),
                    ("relu", C.c_int32), ("scale", C.POINTER(C.c_float)), ("offset", C.POINTER(C.c_float))]

    class Plan(C.Structure):
        _fields_ = [("num_inputs", C.c_int32), ("input_names", C.POINTER(C.c_char_p)), ("input_dims", C.POINTER(C.c_int32)),
                    ("num_frame_layers", C.c_int32), ("frame_layers", C.POINTER(Layer)), ("pool_src", C.c_int32),
                    ("pool_left_context", C.c_int32), ("pool_right_context", C.c_int32),
                    ("pool_output_stddevs", C.c_int32), ("pool_variance_floor", C.c_float),
                    ("num_segment_layers", C.c_int32), ("segment_layers", C.POINTER(Layer))]

    m = synth.make_tiny_model(feat_dim=5, hidden=16, pooled=24, emb=8)
    keep = []

    def fptr(a):
        a = np.ascontiguousarray(a, np.float32)
        keep.append(a)
        return a.ctypes.data_as(C.POINTER(C.c_float))

    def layer(name, taps, bn=True, relu=True):
        f = m.components[name + ".affine"].fields
        W, b = f["LinearParams"], f["BiasParams"]
        t = (Tap * max(len(taps), 1))(*[Tap(s, o) for s, o in taps])
        keep.append(t)
        L = Layer(name.encode(), len(taps), t, W.shape[1], W.shape[0], fptr(W), fptr(b), int(relu), None, None)
        if bn:
            g = m.components[name + ".batchnorm"].fields
            s = np.float32(g["TargetRms"]) * (np.maximum(g["StatsVar"], 0) + np.float32(g["Epsilon"])) ** np.float32(-0.5)
            L.scale, L.offset = fptr(s), fptr(-g["StatsMean"] * s)
        return L

    frames = (Layer * 3)(layer("tdnn1", [(-1, o) for o in (-2, -1, 0, 1, 2)]), layer("tdnn2", [(0, -2), (0, 0), (0, 2)]),
                         layer("tdnn3", [(1, 0)]))
    segs = (Layer * 1)(layer("tdnn6", [], bn=False, relu=False))
    names = (C.c_char_p * 1)(b"input")
    dims = (C.c_int32 * 1)(5)
    plan = Plan(1, names, dims, 3, frames, 2, 0, 10000, 1, 1e-10, 1, segs)
    lib = cvx.load_library()
    h = C.c_void_p()
    lib.cvx_model_from_plan.argtypes = [C.POINTER(Plan), C.POINTER(C.c_void_p)]
    assert lib.cvx_model_from_plan(C.byref(plan), C.byref(h)) == 0, lib.cvx_last_error()
    mod = cvx.Model.__new__(cvx.Model)
    mod._lib, mod._h = lib, h
    info = cvx.ModelInfo()
    assert lib.cvx_model_get_info(h, C.byref(info)) == 0
    mod.info, mod.input_names, mod.input_dims = info, ["input"], [5]
    assert (info.left_context, info.right_context, info.embedding_dim) == (4, 4, 8)
    lengths = [40, 100, 25]
    feats = {"input": synth.make_features(lengths, 5, 3)}
    got, ok = cvx.Extractor(mod, precision="fp32").extract(feats, lengths, chunk_size=60, min_chunk_size=25)
    ref, _ = extract_all(Oracle(m, "tdnn6.affine", np.float64), feats, lengths, chunk_size=60, min_chunk_size=25)
    r, c = parity_metrics(got[ok], ref[ok])
    assert r <= 2e-5, r
    # malformed plans are refused with a message, not a crash
    plan.pool_src = 7
    assert lib.cvx_model_from_plan(C.byref(plan), C.byref(h)) == -1 and b"pool_src" in lib.cvx_last_error()


def test_cli_reads_compressed_feature_archives(cvx, kf, synth, model_dir, require_gpu, tmp_path):
    """feats.scp archives written by steps/make_mfcc.sh hold CompressedMatrix entries
    (CM2 here: uint16 per value, global min/range header); the reader decodes them."""
    import os
    import struct
    import subprocess
    from oracle.nnet3_oracle import Oracle, extract_all, parity_metrics
    path, model = model_dir("v2_xvector")
    lengths = [120, 300]
    x = synth.make_features(lengths, 23, 9)
    lo, rng = float(x.min()), float(x.max() - x.min())
    q = np.clip(np.round((x - lo) / rng * 65535.0), 0, 65535).astype("<u2")
    deq = (lo + rng * (1.0 / 65535.0) * q.astype(np.float32)).astype(np.float32)
    with open(tmp_path / "cm.ark", "wb") as f, open(tmp_path / "cm.scp", "w") as scp:
        r = 0
        for k, T in zip(synth.utt_keys(2), lengths):
            f.write(k.encode() + b" ")
            scp.write(f"{k} {tmp_path}/cm.ark:{f.tell()}\n")
            f.write(b"\0BCM2 " + struct.pack("<ffii", lo, rng, T, 23) + q[r:r + T].tobytes())
            r += T
    exe = os.path.join(cvx.BIN_DIR, "nnet3-xvector-compute")
    res = subprocess.run([exe, "--min-chunk-size=25", "--chunk-size=200", "--output-node=tdnn6.affine", path,
                          f"scp:{tmp_path}/cm.scp", f"ark:{tmp_path}/o.ark"], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr
    got = np.stack([v for _, v in kf.read_ark(str(tmp_path / "o.ark"))])
    ref, ok = extract_all(Oracle(model, "tdnn6.affine", np.float64), {"input": deq}, lengths, chunk_size=200, min_chunk_size=25)
    rr, cc = parity_metrics(got, ref[ok])
    assert rr <= TOL_REL and cc >= TOL_COS, (rr, cc)


@pytest.mark.parametrize("fmt", ["CM", "CM3"])
def test_cli_reads_cm_and_cm3_archives(cvx, kf, synth, model_dir, require_gpu, tmp_path, fmt):
    """the other two CompressedMatrix layouts through the C++ reader: CM = per-column uint16 percentile
    header + one byte per value, column-major (what steps/make_mfcc.sh --compress=true writes); CM3 = one
    byte per value with a global min/range.  Expected values = the decoded matrices (decode formulas
    restated here, [UPSTREAM-RECALL matrix/compressed-matrix.cc])."""
    import os
    import struct
    import subprocess
    from oracle.nnet3_oracle import Oracle, extract_all, parity_metrics
    path, model = model_dir("v2_xvector")
    lengths = [150, 260]
    x = synth.make_features(lengths, 23, 19)
    lo, rng = float(x.min()), float(x.max() - x.min())
    deq = np.zeros_like(x)
    with open(tmp_path / "c.ark", "wb") as f, open(tmp_path / "c.scp", "w") as scp:
        r = 0
        for k, T in zip(synth.utt_keys(2), lengths):
            m = x[r:r + T]
            f.write(k.encode() + b" ")
            scp.write(f"{k} {tmp_path}/c.ark:{f.tell()}\n")
            if fmt == "CM3":
                q = np.clip(np.round((m - lo) / rng * 255.0), 0, 255).astype(np.uint8)
                deq[r:r + T] = np.float32(lo) + np.float32(rng) * np.float32(1.0 / 255.0) * q.astype(np.float32)
                f.write(b"\0BCM3 " + struct.pack("<ffii", lo, rng, T, 23) + q.tobytes())
            else:
                inc = np.float32(rng) * np.float32(1.0 / 65535.0)
                to16 = lambda v: np.clip(np.round((v - lo) / rng * 65535.0), 0, 65535).astype(np.uint16)
                srt = np.sort(m, axis=0)
                h = np.stack([to16(srt[0]), to16(srt[T // 4]), to16(srt[3 * T // 4]), to16(srt[-1])], axis=1)   # [cols, 4]
                h[:, 1] = np.maximum(h[:, 1], h[:, 0] + 1); h[:, 2] = np.maximum(h[:, 2], h[:, 1] + 1); h[:, 3] = np.maximum(h[:, 3], h[:, 2] + 1)
                p = np.float32(lo) + inc * h.astype(np.float32)                                                  # percentile values
                body = np.zeros((23, T), np.uint8)
                for c in range(23):
                    v = m[:, c]
                    b = np.where(v < p[c, 1], np.round((v - p[c, 0]) / (p[c, 1] - p[c, 0]) * 64.0),
                                 np.where(v < p[c, 2], 64 + np.round((v - p[c, 1]) / (p[c, 2] - p[c, 1]) * 128.0),
                                          192 + np.round((v - p[c, 2]) / (p[c, 3] - p[c, 2]) * 63.0)))
                    b = np.clip(b, 0, 255).astype(np.uint8)
                    body[c] = b
                    bf = b.astype(np.float32)
                    deq[r:r + T, c] = np.where(b <= 64, p[c, 0] + (p[c, 1] - p[c, 0]) * bf * np.float32(1 / 64.0),
                                               np.where(b <= 192, p[c, 1] + (p[c, 2] - p[c, 1]) * (bf - 64) * np.float32(1 / 128.0),
                                                        p[c, 2] + (p[c, 3] - p[c, 2]) * (bf - 192) * np.float32(1 / 63.0)))
                f.write(b"\0BCM " + struct.pack("<ffii", lo, rng, T, 23) + h.astype("<u2").tobytes() + body.tobytes())
            r += T
    assert np.max(np.abs(deq - x)) < 0.1          # the compression is lossy but sane
    exe = os.path.join(cvx.BIN_DIR, "nnet3-xvector-compute")
    res = subprocess.run([exe, "--min-chunk-size=25", "--chunk-size=200", "--output-node=tdnn6.affine", path,
                          f"scp:{tmp_path}/c.scp", f"ark:{tmp_path}/o.ark"], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr
    got = np.stack([v for _, v in kf.read_ark(str(tmp_path / "o.ark"))])
    ref, ok = extract_all(Oracle(model, "tdnn6.affine", np.float64), {"input": deq.astype(np.float32)}, lengths,
                          chunk_size=200, min_chunk_size=25)
    rr, cc = parity_metrics(got, ref[ok])
    assert rr <= TOL_REL and cc >= TOL_COS, (fmt, rr, cc)


def test_cli_num_gpus_is_checked_against_visible_devices(cvx, require_gpu, tmp_path):
    """--num-gpus beyond the visible devices is refused before any work (SURVEY 8e mode ii)."""
    import os
    import subprocess
    g = os.path.join(os.path.dirname(__file__), "golden")
    exe = os.path.join(cvx.BIN_DIR, "nnet3-xvector-compute")
    n = cvx.device_count()
    r = subprocess.run([exe, f"--num-gpus={n + 1}", "--output-node=tdnn6.affine", f"{g}/v2_small_final.raw",
                        f"ark:{g}/v2_small_feats.ark", f"ark:{tmp_path}/x.ark"], capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "CUDA device(s) are visible" in r.stderr
    # all visible devices from one process: same archive as one device
    outs = []
    for k in (1, n):
        r = subprocess.run([exe, f"--num-gpus={k}", "--batch-rows=300", "--min-chunk-size=25", "--chunk-size=200",
                            "--output-node=tdnn6.affine", f"{g}/v2_small_final.raw", f"ark:{g}/v2_small_feats.ark",
                            f"ark:{tmp_path}/x{k}.ark"], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        outs.append(open(f"{tmp_path}/x{k}.ark", "rb").read())
    assert outs[0] == outs[1]


def test_one_very_long_utterance_without_chunking(cvx, synth, model_dir, require_gpu):
    """--chunk-size=-1 on an utterance longer than a device batch (the engine grows its batch) and
    longer than the pooling window: like upstream, only frames inside the window [0, right-context]
    are pooled, so the result equals that of the utterance cut at right-context + R + 1 frames."""
    path, _ = model_dir("v2_xvector")
    mod = cvx.Model(path, "tdnn6.affine")
    ex = cvx.Extractor(mod, precision="fp16")
    T = 600_000                                    # > the 524 288-row default batch
    x = synth.make_features([T], 23, 5)
    whole, ok = ex.extract({"input": x}, [T], chunk_size=-1, min_chunk_size=25)
    assert ok.all() and np.isfinite(whole).all()
    cut = 10000 + mod.info.right_context + 1       # stats layer right-context of the recipes' nets
    head, _ = ex.extract({"input": x[:cut]}, [cut], chunk_size=-1, min_chunk_size=25)
    assert np.max(np.abs(whole - head)) <= 2e-5 * np.max(np.abs(head))
    # and chunked the recipe's way it is the frame-weighted average of 60 chunks
    chunked, _ = ex.extract({"input": x}, [T], chunk_size=10000, min_chunk_size=25)
    per_chunk = ex.last_chunk_embeddings(60)
    assert np.allclose(chunked[0], per_chunk.mean(axis=0), rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("topology", ["v2_xvector", "v5_cvector"])
def test_bf16_operands_end_to_end(cvx, synth, model_dir, require_gpu, topology):
    """the bf16 variants of every kernel (pair / halo / narrow ring / fused pooling / segment split):
    not the default -- bf16 misses the 1e-3 tolerance -- but they must stay correct."""
    from oracle.nnet3_oracle import extract_all, parity_metrics
    path, model = model_dir(topology)
    node = embedding_node(synth, topology)
    mod = cvx.Model(path, node)
    lengths = [300, 57, 1234, 100, 640]
    feats = _inputs(synth, mod, lengths, 31)
    got, ok = cvx.Extractor(mod, precision="bf16").extract(feats, lengths, chunk_size=500, min_chunk_size=25)
    ref, _ = extract_all(_oracle(model, node), feats, lengths, chunk_size=500, min_chunk_size=25)
    r, c = parity_metrics(got[ok], ref[ok])
    assert r <= 8e-3 and c >= 0.9999, (r, c)


def test_fp16_range_check_mode_counts_clipped_values(cvx, kf, synth, require_gpu, tmp_path):
    """check_saturation engines store every layer and count the values the saturating fp16 converts
    clipped at +-65504: zero for the recipes' BatchNorm-ed nets, non-zero -- with the
    layer identified -- for a net whose second layer explodes; the drop-in program warns about it."""
    import os
    import subprocess
    m = synth.make_tiny_model()
    lengths = [60, 200]
    feats = {"input": synth.make_features(lengths, 5, 3)}
    kf.write_nnet3(str(tmp_path / "ok.raw"), m, True)
    ex = cvx.Extractor(cvx.Model(str(tmp_path / "ok.raw"), "tdnn6.affine"), precision="fp16", check_saturation=True)
    a, ok = ex.extract(feats, lengths, chunk_size=100, min_chunk_size=25)
    assert ok.all() and ex.saturation_counts() == [0, 0, 0, 0]
    plain, _ = cvx.Extractor(cvx.Model(str(tmp_path / "ok.raw"), "tdnn6.affine"), precision="fp16").extract(
        feats, lengths, chunk_size=100, min_chunk_size=25)
    assert np.allclose(a, plain, rtol=2e-5, atol=2e-5)          # check mode = stand-alone pooling kernels, same numbers
    # BatchNorm of tdnn2 with a tiny variance: scale ~ 3e4, outputs far beyond the fp16 range
    m.components["tdnn2.batchnorm"].fields["StatsVar"] = np.full_like(m.components["tdnn2.batchnorm"].fields["StatsVar"], 1e-12)
    m.components["tdnn2.batchnorm"].fields["Epsilon"] = 1e-9
    kf.write_nnet3(str(tmp_path / "hot.raw"), m, True)
    ex = cvx.Extractor(cvx.Model(str(tmp_path / "hot.raw"), "tdnn6.affine"), precision="fp16", check_saturation=True)
    ex.extract(feats, lengths, chunk_size=100, min_chunk_size=25)
    c = ex.saturation_counts()
    assert c[0] == 0 and c[1] == 0 and c[2] > 0, c
    assert cvx.Extractor(cvx.Model(str(tmp_path / "hot.raw"), "tdnn6.affine"), precision="fp16").saturation_counts() == []
    kf.write_ark(str(tmp_path / "f.ark"), [("a", feats["input"][:60]), ("b", feats["input"][60:])])
    exe = os.path.join(cvx.BIN_DIR, "nnet3-xvector-compute")
    r = subprocess.run([exe, "--check-saturation=true", "--min-chunk-size=25", "--chunk-size=100", "--output-node=tdnn6.affine",
                        str(tmp_path / "hot.raw"), f"ark:{tmp_path}/f.ark", f"ark:{tmp_path}/o.ark"], capture_output=True, text=True)
    assert r.returncode == 0 and "clipped at the fp16 range limit" in r.stderr and "tdnn2" in r.stderr, r.stderr
