#!/bin/bash
# gpurun with a small snapshot: the push is charged GPU time (~2 MB/s), so iterations leave the big packed plans and videos
# at home.  usage: tools/gpurun_lite.sh <timeout> '<command>' [extra paths to KEEP, e.g. V4__ch_det]
T=$1; CMD=$2; KEEP=${3:-}
cp .gpurunignore /tmp/gpurunignore.full
{
  cat /tmp/gpurunignore.full
  for f in video_subtitle_extractor_b200/weights/V4__ch_det.vsep video_subtitle_extractor_b200/weights/V4__ch_rec.vsep \
           video_subtitle_extractor_b200/weights/V2__ch_rec.vsep video_subtitle_extractor_b200/weights/V3__japan_rec_fast.vsep \
           video_subtitle_extractor_b200/weights/V3__korean_rec_fast.vsep tests/golden/_videos/test_japan.mp4 \
           tests/golden/_videos/test_korean.flv tests/golden/_videos/test_cn.mp4 tests/golden/_videos/test_en.mp4; do
    case " $KEEP " in *" $(basename $f) "*) ;; *) echo $f;; esac
  done
} > .gpurunignore
gpurun --timeout $T -- "$CMD"; RC=$?
cp /tmp/gpurunignore.full .gpurunignore
exit $RC
