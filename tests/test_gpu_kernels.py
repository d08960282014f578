"""Single-kernel parity (through the C ABI test entries) against numpy."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def ref_tdnn(x, offsets, w, b, relu, scale, offset):
    rows, d = x.shape
    xs = np.zeros((rows, d * len(offsets)), np.float64)
    for i, o in enumerate(offsets):
        src = np.arange(rows) + o
        ok = (src >= 0) & (src < rows)
        xs[ok, i * d:(i + 1) * d] = x[src[ok]]
    y = xs @ w.astype(np.float64).T
    if b is not None:
        y = y + b
    if relu:
        y = np.maximum(y, 0)
    if scale is not None:
        y = y * scale + offset
    return y


CASES = [
    # rows, in_dim, offsets, out_dim
    (300, 23, [-2, -1, 0, 1, 2], 512),       # first layer of every net
    (1000, 512, [-2, 0, 2], 512),            # tdnn2
    (777, 512, [0], 1500),                   # tdnn5, N not a multiple of 256
    (515, 650, [-1, 0, 1], 650),             # AM layers: K and N need padding
    (129, 650, [-6, -3, 0], 128),            # AM bottleneck
    (64, 40, [0], 64),                       # tiny
    # widths that end in a 16-column tail (the MMAs produce roundup(out_dim, 16) columns, the store maps clip there)
    (300, 64, [0], 40),                      # one tile of 48 columns: full-chunk store clipped at 48
    (260, 96, [-1, 0, 1], 272),              # tiles 160 + 112: the second ends in a 48-column chunk
    (300, 128, [-2, 0, 2], 330),             # tiles 192 + 144: the second ends in a clipped half chunk
    (200, 64, [0], 10),                      # N = 16, the narrowest instruction
]


@pytest.mark.parametrize("precision,tol", [("fp32", 2e-5), ("fp16", 2e-3), ("bf16", 1.5e-2)])
@pytest.mark.parametrize("rows,in_dim,offsets,out_dim", CASES)
def test_tdnn_layer(cvx, require_gpu, precision, tol, rows, in_dim, offsets, out_dim):
    rng = np.random.default_rng(rows + in_dim)
    x = rng.standard_normal((rows, in_dim)).astype(np.float32)
    k = in_dim * len(offsets)
    w = (rng.standard_normal((out_dim, k)) / np.sqrt(k)).astype(np.float32)
    b = (0.1 * rng.standard_normal(out_dim)).astype(np.float32)
    scale = rng.uniform(0.5, 2.0, out_dim).astype(np.float32)
    offset = rng.standard_normal(out_dim).astype(np.float32)
    got = cvx.test_tdnn_layer(x, offsets, w, b, True, scale, offset, precision=precision)
    ref = ref_tdnn(x, offsets, w, b, True, scale, offset)
    err = np.max(np.abs(got - ref)) / np.max(np.abs(ref))
    assert err < tol, f"{precision}: max rel err {err}"


@pytest.mark.parametrize("rows,in_dim,offsets,out_dim", CASES)
def test_tdnn_layer_single_cta_kernel(cvx, require_gpu, monkeypatch, rows, in_dim, offsets, out_dim):
    monkeypatch.setenv("CVX_GEMM_1CTA", "1")
    test_tdnn_layer(cvx, require_gpu, "fp16", 2e-3, rows, in_dim, offsets, out_dim)


@pytest.mark.parametrize("precision", ["fp32", "fp16"])
def test_tdnn_layer_compact_splice_and_plain(cvx, require_gpu, precision):
    """compact first-layer staging path; no relu / no batchnorm epilogue."""
    rng = np.random.default_rng(5)
    x = rng.standard_normal((400, 23)).astype(np.float32)
    offsets = [-2, -1, 0, 1, 2]
    w = (rng.standard_normal((512, 115)) / 10).astype(np.float32)
    got = cvx.test_tdnn_layer(x, offsets, w, None, False, None, None, precision=precision, compact=True)
    ref = ref_tdnn(x, offsets, w, None, False, None, None)
    err = np.max(np.abs(got - ref)) / np.max(np.abs(ref))
    assert err < (2e-5 if precision == "fp32" else 2e-3)


@pytest.mark.parametrize("single_cta", [False, True])
def test_tdnn_layer_many_tiles(cvx, require_gpu, monkeypatch, single_cta):
    """more tiles than SMs: exercises the persistent loop, both accumulator
    stages and the smem ring wrap-around -- for the CTA-pair kernel (default) and
    the 1-CTA kernel (CVX_GEMM_1CTA=1)."""
    monkeypatch.setenv("CVX_GEMM_1CTA", "1" if single_cta else "0")
    rng = np.random.default_rng(9)
    rows = 128 * 200 + 37
    x = rng.standard_normal((rows, 128)).astype(np.float32)
    w = (rng.standard_normal((512, 384)) / 20).astype(np.float32)
    b = rng.standard_normal(512).astype(np.float32)
    got = cvx.test_tdnn_layer(x, [-3, 0, 3], w, b, True, None, None, precision="fp16")
    ref = ref_tdnn(x, [-3, 0, 3], w, b, True, None, None)
    err = np.max(np.abs(got - ref)) / np.max(np.abs(ref))
    assert err < 2e-3


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-5), ("fp16", 1e-3), ("bf16", 8e-3)])
def test_stats_pool(cvx, require_gpu, precision, tol):
    rng = np.random.default_rng(3)
    rows, dim = 5000, 1500
    x = (rng.standard_normal((rows, dim)) * rng.uniform(0.1, 3, dim) + rng.standard_normal(dim)).astype(np.float32)
    begin = [0, 7, 300, 1000, 1001, 4000]
    end = [7, 300, 1000, 1001, 4000, 5000]      # 1-row chunk, multi-slab chunks, ragged
    got = cvx.test_stats_pool(x, begin, end, precision=precision)
    if precision == "fp16":
        xq = x.astype(np.float16).astype(np.float64)
    elif precision == "bf16":
        import importlib
        xq = x.astype(np.float64)
    else:
        xq = x.astype(np.float64)
    for c, (b, e) in enumerate(zip(begin, end)):
        seg = xq[b:e]
        mean = seg.mean(0)
        std = np.sqrt(np.maximum((seg * seg).mean(0) - mean * mean, 1e-10))
        ref = np.concatenate([mean, std])
        err = np.max(np.abs(got[c] - ref)) / np.max(np.abs(ref))
        assert err < tol, f"chunk {c}: {err}"


def test_stats_pool_constant_column_hits_floor(cvx, require_gpu):
    x = np.ones((100, 64), np.float32) * 3.0
    got = cvx.test_stats_pool(x, [0], [100], precision="fp32")
    assert np.allclose(got[0, :64], 3.0)
    assert np.allclose(got[0, 64:], np.sqrt(1e-10), rtol=1e-3)


@pytest.mark.parametrize("m,k,n", [(1, 3000, 512), (37, 3000, 512), (700, 1024, 512), (5000, 3000, 512), (300, 200, 96)])
@pytest.mark.parametrize("precision,tol", [("fp16", 3e-5), ("bf16", 1e-4), ("fp32", 3e-5)])
def test_segment_affine_hi_lo_split_is_fp32_accurate(cvx, require_gpu, m, k, n, precision, tol):
    """tdnn6 in isolation: the 16-bit hi/lo split GEMM (x = hi + lo, W = Wh + Wl,
    y = hi Wh + lo Wh + hi Wl on tcgen05, fp32 epilogue) against fp64 numpy -- fp32-class accuracy (an fp32 sum over K = 3000 terms is itself good to ~3e-6)
    (fp16 split: ~2^-22; bf16 split keeps 16 mantissa bits: ~2^-16), with and without ReLU + BatchNorm."""
    rng = np.random.default_rng(m + k)
    x = rng.standard_normal((m, k)).astype(np.float32) * np.float32(3.0)
    x[:, : k // 2] = np.abs(x[:, : k // 2])                       # pooled vectors: means of post-ReLU values, stddevs > 0
    w = (rng.standard_normal((n, k)) / np.sqrt(k)).astype(np.float32)
    b = (0.1 * rng.standard_normal(n)).astype(np.float32)
    ref = x.astype(np.float64) @ w.astype(np.float64).T + b
    got = cvx.test_segment_affine(x, w, b, precision=precision)
    assert np.max(np.abs(got - ref)) <= tol * np.max(np.abs(ref)), np.max(np.abs(got - ref)) / np.max(np.abs(ref))
    s = rng.uniform(0.5, 2.0, n).astype(np.float32)
    o = (0.2 * rng.standard_normal(n)).astype(np.float32)
    ref2 = np.maximum(ref, 0) * s + o
    got2 = cvx.test_segment_affine(x, w, b, relu=True, scale=s, offset=o, precision=precision)
    assert np.max(np.abs(got2 - ref2)) <= 2 * tol * np.max(np.abs(ref2))
