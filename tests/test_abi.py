"""The C-ABI library loads and exports everything include/cvx.h declares; no compute
calls here (no GPU in this container)."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_are_exported(cvx):
    lib = cvx.load_library()
    hdr = open(os.path.join(ROOT, "include", "cvx.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(cvx_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(cvx.ABI_SYMBOLS)
    for s in declared:
        assert hasattr(lib, s), s
    nm = subprocess.run(["nm", "-D", "--defined-only", cvx.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (cvx_[a-z0-9_]+)", nm))
    assert declared <= exported


def test_header_is_plain_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "cvx.h"\nint main(void){cvx_engine_options o; cvx_plan p; (void)o; (void)p; return CVX_OK;}\n')
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(src), "-o",
                        str(tmp_path / "t.o")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_library_has_no_torch_dependency(cvx):
    out = subprocess.run(["ldd", cvx.LIB_PATH], capture_output=True, text=True).stdout
    assert "torch" not in out and "libcudart" in out


def test_sass_proves_blackwell_native_kernel(cvx):
    """tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM, TMA -> UTMALDG/UTMASTG (B200_PROFILING.md)."""
    sass = subprocess.run(["cuobjdump", "-sass", cvx.LIB_PATH], capture_output=True, text=True).stdout
    assert re.search(r"UTC\w*MMA", sass), "no tcgen05 MMA in SASS"
    assert "LDTM" in sass and "UTMALDG" in sass and "UTMASTG" in sass
    assert "sm_100a" in sass


def test_no_cpu_fallback(cvx, model_dir):
    """without a CUDA device engine creation fails loudly with CVX_ERR_CUDA."""
    if cvx.device_count() > 0:
        pytest.skip("a GPU is present")
    path, _ = model_dir("tiny")
    mod = cvx.Model(path, "tdnn6.affine")
    with pytest.raises(cvx.CvxError) as e:
        cvx.Extractor(mod)
    assert e.value.code == -5 and "no CPU fallback" in str(e.value)


def test_product_code_does_not_touch_the_oracle():
    """only tests/, __graft_entry__.smoke() and bench.py's baseline legs may import, call,
    link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "speaker-embedding-with-phonetic-information_b200")
    pat = re.compile(r"(import\s+oracle|from\s+oracle|oracle/|oracle\.(nnet3_oracle|torch_route|cpu_baseline)|import_module\([^)]*oracle)")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cc", ".cu", ".h", ".cuh")) or f == "Makefile":
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert not pat.search(txt), f"{f} uses the oracle"
    # bench.py: only inside run_reference(), the cpu_baseline leg, and parity_check() -- the oracle as the
    # CHECKER of the timed output (outside the timed region), never as the thing measured
    b = open(os.path.join(ROOT, "bench.py")).read()
    assert len(re.findall(r"from oracle import cpu_baseline", b)) == 2
    assert len(re.findall(r"(?:from|import) oracle", b)) == 3
    body = b[b.index("def parity_check("):b.index("def cli_leg(")]
    assert "from oracle.nnet3_oracle import" in body and b.count("nnet3_oracle import") == 1


def test_model_from_plan_and_invalid_arguments(cvx):
    import ctypes as C
    lib = cvx.load_library()
    h = C.c_void_p()
    assert lib.cvx_model_load(None, None, C.byref(h)) == -1
    assert b"NULL" in lib.cvx_last_error()
    assert lib.cvx_extract(None, None, 1, 0, None, None, None, None, 0, 0, None, 0) == -1
    o = cvx.EngineOptions()
    lib.cvx_engine_options_default(C.byref(o))
    assert o.precision == 0 and o.device == 0 and o.max_batch_rows == 0


def test_cli_usage_and_option_errors(cvx):
    exe = os.path.join(cvx.BIN_DIR, "nnet3-xvector-compute")
    r = subprocess.run([exe, "--help"], capture_output=True, text=True)
    assert r.returncode == 0 and "Usage: nnet3-xvector-compute" in r.stderr and "--min-chunk-size" in r.stderr
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 1 and "Usage" in r.stderr
    r = subprocess.run([exe, "--no-such-option=1", "a", "b", "c"], capture_output=True, text=True)
    assert r.returncode != 0 and "invalid option" in r.stderr
    r = subprocess.run([exe, "--use-gpu=maybe", "a", "ark:b", "ark:c"], capture_output=True, text=True)
    assert r.returncode != 0
    r = subprocess.run([exe, "--use-gpu=no", "/nonexistent/final.raw", "ark:/dev/null", "ark:/dev/null"],
                       capture_output=True, text=True)
    assert r.returncode != 0 and "ERROR" in r.stderr
    exe2 = os.path.join(cvx.BIN_DIR, "nnet3-cvector-compute-multiple-input")
    r = subprocess.run([exe2, "--help"], capture_output=True, text=True)
    assert r.returncode == 0 and "--name" in r.stderr
