#!/bin/bash
# Pins this repo's oracle to REAL Kaldi: runs upstream's binaries on the full-width replay fixtures
# exactly as the reference's recipe does and compares with the committed expectations.
#
#   tools/replay_with_kaldi.sh <kaldi-root>           (needs a built Kaldi: src/nnet3bin, src/featbin, src/ivectorbin)
#
# Recipe lines replayed (R = the reference repo):
#   R/egs/sre/v5/sid/nnet3/xvector/extract_xvectors_new.sh:57-59   nnet="nnet3-copy --nnet-config=$dir/extract.config $srcdir/final.raw - |"
#   :79   feat="ark:apply-cmvn-sliding --norm-vars=false --center=true --cmn-window=300 scp:feats.scp ark:- | select-voiced-frames ark:- scp,s,cs:vad.scp ark:- |"
#   :91-93 nnet3-xvector-compute --use-gpu=no --min-chunk-size=$min --chunk-size=$chunk "$nnet" "$feat" ark,scp:...
#   R/egs/sre/v5/sid/nnet3_cvector/cvector/extract_am_embedding.sh:79-85   nnet3-compute --use-gpu=no "$nnet" ... (frame-level)
# A MATCH on every line below turns "parity unpinned"  into "pinned".
set -e
KALDI=${1:?usage: $0 <kaldi-root>}
HERE=$(cd "$(dirname "$0")/.." && pwd)
KIT=${2:-$HERE/replay_kit}
export PATH=$KALDI/src/nnet3bin:$KALDI/src/featbin:$KALDI/src/ivectorbin:$KALDI/src/bin:$PATH
python "$HERE/tools/make_replay_kit.py" --out "$KIT"
GOLD=$HERE/tests/golden/replay
status=0
for net in v2 v5; do
  dir=$KIT/$net
  min=$(cat $dir/min_chunk_size); chunk=1000
  nnet="nnet3-copy --nnet-config=$dir/extract.config $dir/final.raw - |"
  echo "== $net: pre-normalised features, --chunk-size=$chunk --min-chunk-size=$min"
  if nnet3-xvector-compute --use-gpu=no --pad-input=false --min-chunk-size=$min --chunk-size=$chunk "$nnet" ark:$dir/feats.ark ark,t:$dir/kaldi.txt 2>$dir/kaldi.log; then
    python "$HERE/tools/compare_vectors.py" $dir/kaldi.txt $GOLD/${net}_expected.txt || status=1
    nnet3-xvector-compute --use-gpu=no --pad-input=true --min-chunk-size=$min --chunk-size=$chunk "$nnet" ark:$dir/feats.ark ark,t:$dir/kaldi_pad.txt 2>>$dir/kaldi.log
    python "$HERE/tools/compare_vectors.py" $dir/kaldi_pad.txt $GOLD/${net}_expected_pad.txt || status=1
  else
    echo "   (this Kaldi predates --pad-input: short tails are skipped)"
    nnet3-xvector-compute --use-gpu=no --min-chunk-size=$min --chunk-size=$chunk "$nnet" ark:$dir/feats.ark ark,t:$dir/kaldi.txt 2>$dir/kaldi.log
    python "$HERE/tools/compare_vectors.py" $dir/kaldi.txt $GOLD/${net}_expected.txt || status=1
  fi
  echo "== $net: the recipe's feature pipe (apply-cmvn-sliding | select-voiced-frames)"
  ( cd $dir
    feat="ark:apply-cmvn-sliding --norm-vars=false --center=true --cmn-window=300 scp:raw_feats.scp ark:- | select-voiced-frames ark:- scp,s,cs:vad.scp ark:- |"
    nnet3-xvector-compute --use-gpu=no --min-chunk-size=$min --chunk-size=$chunk "$nnet" "$feat" ark,t:kaldi_rawpipe.txt 2>>kaldi.log )
  python "$HERE/tools/compare_vectors.py" $dir/kaldi_rawpipe.txt $GOLD/${net}_expected_rawpipe.txt || status=1
done
echo "== v5: nnet3-compute, phonetic bottleneck tdnn5.batchnorm at every frame of utt0000004"
dir=$KIT/v5
echo "output-node name=output input=tdnn5.batchnorm" > $dir/extract_bn.config
nnet3-copy-feats --help >/dev/null 2>&1 || true
( cd $dir && grep '^utt0000004 ' feats.scp > one.scp &&
  nnet3-compute --use-gpu=no "nnet3-copy --nnet-config=extract_bn.config final.raw - |" scp:one.scp ark,t:kaldi_bn.txt 2>>kaldi.log )
python "$HERE/tools/compare_vectors.py" $dir/kaldi_bn.txt $GOLD/v5_expected_bn_frames.txt || status=1
[ $status -eq 0 ] && echo "ALL MATCH: the oracle restates this Kaldi build's nnet3 extraction" || echo "MISMATCH (see above; logs in $KIT/*/kaldi.log)"
exit $status
