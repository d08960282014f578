"""Tuning aid: process start-up cost of the drop-in program (CUDA context, model load, engine build,
pinned buffers), measured by running it on an empty feature archive."""
import ctypes
import importlib
import os
import subprocess
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
P = "speaker-embedding-with-phonetic-information_b200"
cvx = importlib.import_module(P)
kf = importlib.import_module(P + ".kaldi_format")
sy = importlib.import_module(P + ".synth")
d = tempfile.mkdtemp(prefix="cvx_start_")
model = os.path.join(d, "final.raw")
kf.write_nnet3(model, sy.make_model("v5_cvector"), binary=True)
empty = os.path.join(d, "empty.ark")
open(empty, "wb").close()
exe = os.path.join(cvx.BIN_DIR, "nnet3-xvector-compute")


def run(label, *args):
    t0 = time.perf_counter()
    subprocess.run([exe, *args], capture_output=True)
    print(f"{label}: {time.perf_counter() - t0:.2f} s", flush=True)


run("--help (no CUDA)", "--help")
for i in range(3):
    run("default", "--output-node=tdnn6_xvec.affine", model, f"ark:{empty}", "ark:/dev/null")
run("--batch-rows=65536", "--batch-rows=65536", "--output-node=tdnn6_xvec.affine", model, f"ark:{empty}", "ark:/dev/null")
t0 = time.perf_counter()
ctypes.CDLL("libcudart.so.12").cudaFree(0)
print(f"bare CUDA context in this process: {time.perf_counter() - t0:.2f} s")
