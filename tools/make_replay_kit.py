"""Replay kit for a Kaldi owner: FULL-WIDTH v2 x-vector and v5 c-vector fixtures that the REAL
`nnet3-xvector-compute` / `nnet3-compute` can run exactly the way the reference's recipe does
(R/egs/sre/v5/sid/nnet3/xvector/extract_xvectors_new.sh:57-59,79,91-93), with the embeddings this
repo's oracle expects committed next to it -- the missing link that would pin the oracle
(parity unpinned because Kaldi cannot be built here).

  python tools/make_replay_kit.py [--out replay_kit] [--check]

writes, per net (v2, v5):   <out>/<net>/final.raw  extract.config  min_chunk_size  max_chunk_size
                             feats.ark (post-CMN, post-VAD 23-dim)  raw_feats.ark + vad.ark (for the
                             recipe's apply-cmvn-sliding | select-voiced-frames pipe)  feats.scp ...
The models and features are generated from fixed seeds (28 MB / 18 MB binaries: regenerated, not
committed); what IS committed under tests/golden/replay/ is small:
  manifest.json                 sha256 of every generated file (so a regenerated kit is provably the same)
  <net>_expected.txt            oracle fp64 embeddings, text archive (ark,t), --pad-input=false (older Kaldi)
  <net>_expected_pad.txt        the same with --pad-input=true (newer Kaldi's default)
  <net>_expected_rawpipe.txt    from raw_feats.ark + vad.ark through the restated sliding CMN + selection
  v5_expected_bn_frames.txt     nnet3-compute of the phonetic bottleneck (tdnn5.batchnorm) for one utterance
`--check` regenerates into a temp dir and verifies the manifest + expected files (tests/test_replay_kit.py)."""
import argparse
import hashlib
import importlib
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
P = "speaker-embedding-with-phonetic-information_b200"
kf = importlib.import_module(P + ".kaldi_format")
sy = importlib.import_module(P + ".synth")
from oracle.nnet3_oracle import Oracle, extract_all  # noqa: E402
from oracle.prestep_oracle import select_voiced, sliding_cmn  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden", "replay")
NETS = {"v2": ("v2_xvector", "tdnn6.affine"), "v5": ("v5_cvector", "tdnn6_xvec.affine")}
LENGTHS = [1234, 25, 24, 300, 57, 2100, 0, 45]          # 24 < min-chunk-size fails; 0 rows fails; 2100 = 2 chunks + a 100-frame tail
RAW_LENGTHS = [900, 400, 2500]
CHUNK, MIN_CHUNK, MAX_CHUNK = 1000, 25, 10000


def sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def write_text_vectors(path, items):
    with open(path, "w") as f:
        for k, v in items:
            f.write(k + "  [ " + " ".join(f"{x:.8g}" for x in v) + " ]\n")


def write_text_matrix(path, key, m):
    with open(path, "w") as f:
        f.write(key + "  [\n")
        for i, row in enumerate(m):
            f.write("  " + " ".join(f"{x:.8g}" for x in row) + (" ]\n" if i == len(m) - 1 else "\n"))


def build(out):
    manifest, expected = {}, {}
    for net, (topology, node) in NETS.items():
        d = os.path.join(out, net)
        os.makedirs(d, exist_ok=True)
        kf.write_nnet3(os.path.join(d, "final.raw"), sy.make_model(topology, seed=42), binary=True)
        model = kf.read_nnet3(os.path.join(d, "final.raw"))      # what every reader of the file sees
        open(os.path.join(d, "extract.config"), "w").write(f"output-node name=output input={node}\n")
        open(os.path.join(d, "min_chunk_size"), "w").write(f"{MIN_CHUNK}\n")
        open(os.path.join(d, "max_chunk_size"), "w").write(f"{MAX_CHUNK}\n")
        keys = sy.utt_keys(len(LENGTHS))
        feats = sy.make_features(LENGTHS, 23, seed=2024)
        starts = np.cumsum([0] + LENGTHS[:-1])
        kf.write_ark(os.path.join(d, "feats.ark"), [(k, feats[s:s + n]) for k, s, n in zip(keys, starts, LENGTHS)],
                     scp_path=os.path.join(d, "feats.scp"), ark_name_in_scp="feats.ark")
        # raw features with an energy-like offset and a VAD table, for the recipe's own feature pipe
        rkeys = [f"raw{i:04d}" for i in range(len(RAW_LENGTHS))]
        rng = np.random.default_rng(99)
        raw = [(3.0 + 2.0 * rng.standard_normal((n, 23))).astype(np.float32) for n in RAW_LENGTHS]
        vad = [(rng.random(n) > 0.25).astype(np.float32) for n in RAW_LENGTHS]
        kf.write_ark(os.path.join(d, "raw_feats.ark"), list(zip(rkeys, raw)), scp_path=os.path.join(d, "raw_feats.scp"),
                     ark_name_in_scp="raw_feats.ark")
        kf.write_ark(os.path.join(d, "vad.ark"), list(zip(rkeys, vad)), scp_path=os.path.join(d, "vad.scp"), ark_name_in_scp="vad.ark")
        orc = Oracle(model, node, np.float64)
        for tag, pad in (("expected", False), ("expected_pad", True)):
            emb, ok = extract_all(orc, {"input": feats}, LENGTHS, chunk_size=CHUNK, min_chunk_size=MIN_CHUNK, pad_input=pad)
            expected[f"{net}_{tag}.txt"] = [(k, e) for k, e, o in zip(keys, emb, ok) if o]
        sel = [select_voiced(sliding_cmn(x, 300, True), v) for x, v in zip(raw, vad)]
        emb, ok = extract_all(orc, {"input": np.concatenate(sel)}, [len(s) for s in sel], chunk_size=CHUNK, min_chunk_size=MIN_CHUNK)
        expected[f"{net}_expected_rawpipe.txt"] = [(k, e) for k, e, o in zip(rkeys, emb, ok) if o]
        if net == "v5":
            bn = Oracle(model, "tdnn5.batchnorm", np.float64).forward_frames({"input": feats[starts[4]:starts[4] + LENGTHS[4]]})
            expected["v5_expected_bn_frames.txt"] = (keys[4], bn)
        for f in sorted(os.listdir(d)):
            manifest[f"{net}/{f}"] = sha(os.path.join(d, f))
    return manifest, expected


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "replay_kit"))
    ap.add_argument("--check", action="store_true", help="regenerate into a temp dir and verify against tests/golden/replay/")
    ap.add_argument("--write-golden", action="store_true", help="(re)write tests/golden/replay/ from this run")
    a = ap.parse_args()
    out = tempfile.mkdtemp(prefix="cvx_replay_") if a.check else a.out
    manifest, expected = build(out)
    if a.write_golden:
        os.makedirs(GOLD, exist_ok=True)
        json.dump(manifest, open(os.path.join(GOLD, "manifest.json"), "w"), indent=1, sort_keys=True)
        for name, items in expected.items():
            if isinstance(items, tuple):
                write_text_matrix(os.path.join(GOLD, name), *items)
            else:
                write_text_vectors(os.path.join(GOLD, name), items)
    want = json.load(open(os.path.join(GOLD, "manifest.json")))
    bad = [k for k in want if manifest.get(k) != want[k]]
    if bad:
        raise SystemExit(f"generated files differ from tests/golden/replay/manifest.json: {bad}")
    for name, items in expected.items():
        if isinstance(items, tuple):
            got = np.array(list(kf.read_ark(os.path.join(GOLD, name)))[0][1])
            assert np.allclose(got, items[1], rtol=1e-6, atol=1e-6), name
        else:
            got = dict(kf.read_ark(os.path.join(GOLD, name)))
            assert list(got) == [k for k, _ in items], name
            for k, e in items:
                assert np.allclose(got[k], e, rtol=1e-6, atol=1e-6), (name, k)
    print(f"replay kit in {out}: {len(manifest)} files match the manifest; {len(expected)} expected tables verified")


if __name__ == "__main__":
    main()
