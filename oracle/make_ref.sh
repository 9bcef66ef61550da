#!/bin/sh
# TEST / BASELINE INFRASTRUCTURE -- stages the UNMODIFIED reference files the AttFind path lives in under
# oracle/_ref/ (git-ignored, NOT gpurun-ignored: it travels to the GPU box like our own built .so).
# Nothing is edited and nothing is committed: the files are read from where they lie under /root/reference.
#   stylex_train.py            Generator / GeneratorBlock / RGBBlock / Conv2DMod / Blur  (ST:144-153, 604-840)
#   resnet_classifier.py       ResNet.classify_images        (resnet_classifier.py:56-71)
#   mobilenet_classifier.py    MobileNet.classify_images     (mobilenet_classifier.py:57-73)
#   run_attfind_combined.ipynb attfind_extraction (cell 5), find_significant_styles (cell 15), cells 11/14/16
#   version.py, diff_augment.py, debug_encoders.py   module-level imports of stylex_train.py (ST:34,35,58), not executed
# Used by `bench.py --impl reference` (cpu_baseline.kind = "reference": the VERBATIM notebook loop on the host cores)
# and, when /root/reference is absent, by the tests that compare against the live reference.
set -e
SRC="${1:-/root/reference}"
HERE="$(cd "$(dirname "$0")" && pwd)"
DST="$HERE/_ref/stylex"
[ -f "$SRC/stylex/stylex_train.py" ] || { echo "make_ref.sh: no reference under $SRC (keeping $DST as is)"; exit 0; }
mkdir -p "$DST"
for f in stylex_train.py resnet_classifier.py mobilenet_classifier.py run_attfind_combined.ipynb version.py diff_augment.py debug_encoders.py; do
  cp -f "$SRC/stylex/$f" "$DST/$f"
  chmod u+w "$DST/$f"
done
( cd "$DST" && sha256sum stylex_train.py resnet_classifier.py mobilenet_classifier.py run_attfind_combined.ipynb version.py diff_augment.py debug_encoders.py > SHA256SUMS )
echo "make_ref.sh: staged $(ls "$DST" | wc -l) files under $DST"
