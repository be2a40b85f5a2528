#!/bin/bash
# gpurun with a slim snapshot: the server plans (314 MB) and the sample videos stay here unless named in KEEP, so that the push
# (which is charged to the GPU budget) takes seconds.  usage: [KEEP="V4__ch_det test_en"] tools/gpurun_slim.sh <timeout_s> '<command>'
cd "$(dirname "$0")/.."
cp .gpurunignore /tmp/gpurunignore.bak.$$
trap 'cp /tmp/gpurunignore.bak.$$ .gpurunignore; rm -f /tmp/gpurunignore.bak.$$' EXIT
for f in video_subtitle_extractor_b200/weights/V2__ch_rec.vsep video_subtitle_extractor_b200/weights/V4__ch_det.vsep \
         video_subtitle_extractor_b200/weights/V4__ch_rec.vsep tests/golden/_videos/test_en.mp4 tests/golden/_videos/test_cn.mp4 \
         tests/golden/_videos/test_japan.mp4 tests/golden/_videos/test_korean.flv; do
  keep=0
  for k in $KEEP; do case "$f" in *$k*) keep=1;; esac; done
  [ $keep = 0 ] && echo "$f" >> .gpurunignore
done
/usr/local/graft/bin/gpurun --timeout "$1" -- "$2"
