import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PKG = "speaker-embedding-with-phonetic-information_b200"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def cvx():
    return importlib.import_module(PKG)


@pytest.fixture(scope="session")
def kf():
    return importlib.import_module(PKG + ".kaldi_format")


@pytest.fixture(scope="session")
def synth():
    return importlib.import_module(PKG + ".synth")


@pytest.fixture(scope="session")
def model_dir(tmp_path_factory, kf, synth):
    """Random-init model files of every topology, written once per session."""
    d = tmp_path_factory.mktemp("models")
    paths = {}

    def get(topology, binary=True):
        key = (topology, binary)
        if key not in paths:
            if topology == "tiny":
                m = synth.make_tiny_model()
            elif topology == "tiny2":
                m = synth.make_tiny_model(two_branch=True)
            else:
                m = synth.make_model(topology)
            p = str(d / f"{topology}.{'raw' if binary else 'txt'}")
            kf.write_nnet3(p, m, binary)
            paths[key] = (p, m)
        return paths[key]

    return get


def embedding_node(synth, topology):
    return "tdnn6.affine" if topology.startswith("tiny") else synth.EMBEDDING_NODE[topology]


@pytest.fixture(scope="session")
def require_gpu(cvx):
    if cvx.device_count() < 1:
        pytest.fail("no CUDA device: -m gpu tests must run on the GPU box (there is no CPU fallback)")
    return True
