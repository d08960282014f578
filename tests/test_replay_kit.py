"""The Kaldi replay kit (tools/make_replay_kit.py, tools/replay_with_kaldi.sh): the generated full-width
fixtures are reproducible (sha256 manifest) and the committed expectations are what the oracle -- and
on a GPU box the CUDA path -- produce for them."""
import importlib
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "replay")


@pytest.fixture(scope="module")
def kit(tmp_path_factory):
    out = tmp_path_factory.mktemp("replay_kit")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "make_replay_kit.py"), "--out", str(out)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr      # manifest + expected tables verified by the script itself
    return out


def test_kit_is_reproducible_and_oracle_matches_committed_expectations(kit):
    assert sorted(os.listdir(kit)) == ["v2", "v5"]
    assert os.path.getsize(kit / "v5" / "final.raw") > 20e6          # full-width model, not the small golden one


def test_compare_tool_and_script_are_wired(kit, kf):
    exp = os.path.join(GOLD, "v2_expected.txt")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "compare_vectors.py"), exp, exp], capture_output=True, text=True)
    assert r.returncode == 0 and "MATCH" in r.stdout
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "compare_vectors.py"), exp,
                        os.path.join(GOLD, "v2_expected_pad.txt")], capture_output=True, text=True)
    assert r.returncode == 1 and "KEYS DIFFER" in r.stdout        # --pad-input=true also embeds the 24-frame utterance
    sh = open(os.path.join(ROOT, "tools", "replay_with_kaldi.sh")).read()
    for needle in ("nnet3-copy --nnet-config=", "apply-cmvn-sliding --norm-vars=false --center=true --cmn-window=300",
                   "select-voiced-frames ark:- scp,s,cs:vad.scp", "nnet3-xvector-compute --use-gpu=no", "nnet3-compute --use-gpu=no"):
        assert needle in sh


@pytest.mark.gpu
@pytest.mark.parametrize("net,node", [("v2", "tdnn6.affine"), ("v5", "tdnn6_xvec.affine")])
def test_cuda_path_matches_the_replay_expectations(cvx, kf, kit, net, node, require_gpu):
    """the drop-in programs on the kit, invoked like tools/replay_with_kaldi.sh invokes Kaldi's, against the
    same committed files a Kaldi owner compares with (fp16 operands: the north-star tolerance)."""
    from oracle.nnet3_oracle import parity_metrics
    d = kit / net
    exe = os.path.join(cvx.BIN_DIR, "nnet3-xvector-compute")
    runs = [(["--pad-input=false", f"ark:{d}/feats.ark"], "expected"), (["--pad-input=true", f"ark:{d}/feats.ark"], "expected_pad"),
            (["--cmn-window=300", f"--vad-rspecifier=scp:{d}/vad.scp", f"scp:{d}/raw_feats.scp"], "expected_rawpipe")]
    for args, tag in runs:
        out = d / f"cvx_{tag}.txt"
        r = subprocess.run([exe, "--use-gpu=no", "--min-chunk-size=25", "--chunk-size=1000", f"--nnet-config={d}/extract.config",
                            *args[:-1], f"{d}/final.raw", args[-1], f"ark,t:{out}"], capture_output=True, text=True, cwd=d, timeout=600)
        assert r.returncode == 0, r.stderr
        got, want = dict(kf.read_ark(str(out))), dict(kf.read_ark(os.path.join(GOLD, f"{net}_{tag}.txt")))
        assert list(got) == list(want), tag
        rel, cos = parity_metrics(np.stack(list(got.values())), np.stack(list(want.values())))
        print(net, tag, f"rel {rel:.3e} cos {cos:.8f}")
        assert rel <= 1e-3 and cos >= 0.9999, (tag, rel, cos)
    if net == "v5":
        exe2 = os.path.join(cvx.BIN_DIR, "nnet3-compute")
        open(d / "extract_bn.config", "w").write("output-node name=output input=tdnn5.batchnorm\n")
        one = [ln for ln in open(d / "feats.scp") if ln.startswith("utt0000004 ")]
        open(d / "one.scp", "w").writelines(one)
        r = subprocess.run([exe2, "--use-gpu=no", f"--nnet-config={d}/extract_bn.config", f"{d}/final.raw", f"scp:{d}/one.scp",
                            f"ark,t:{d}/cvx_bn.txt"], capture_output=True, text=True, cwd=d, timeout=600)
        assert r.returncode == 0, r.stderr
        got = np.array(list(kf.read_ark(str(d / "cvx_bn.txt")))[0][1])
        want = np.array(list(kf.read_ark(os.path.join(GOLD, "v5_expected_bn_frames.txt")))[0][1])
        assert got.shape == want.shape and np.max(np.abs(got - want)) <= 2e-3 * np.max(np.abs(want))
