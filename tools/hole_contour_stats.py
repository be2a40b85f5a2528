"""How often does a HOLE contour of cv2.findContours(RETR_LIST) survive DBPostProcess (min size 3, box score >= 0.6)?
Runs the CPU oracle's detector on frames of the reference's own test videos (needs /root/reference) and on synthetic
frames; quoted in DESIGN.md section 7.  Test infrastructure only."""
import sys, glob, os; sys.path.insert(0,'/root/repo')
import numpy as np, cv2
from oracle.pipeline import OraclePipeline
from oracle import hostlogic as hl
from video_subtitle_extractor_b200 import weights
orc = OraclePipeline.from_plans(weights.load_plan_blob("V4/ch_det_fast"), None)
def stats(img):
    pred, shape = orc.det_prob_map(img)
    bitmap = (pred > 0.3)
    cs, hier = cv2.findContours((bitmap*255).astype(np.uint8), cv2.RETR_CCOMP, cv2.CHAIN_APPROX_SIMPLE)
    n_outer = n_hole = n_hole_box = 0
    for c, h in zip(cs, hier[0] if hier is not None else []):
        is_hole = h[3] >= 0
        if not is_hole: n_outer += 1; continue
        n_hole += 1
        pts, ss = hl.get_mini_boxes(c)
        if ss < 3: continue
        sc = hl.box_score_fast(pred, np.array(pts).reshape(-1,2))
        if sc < 0.6: continue
        n_hole_box += 1
    return n_outer, n_hole, n_hole_box
tot = np.zeros(3, int); nf = 0
vids = ['/root/reference/test/test_en.mp4','/root/reference/test/test_cn.mp4','/root/reference/test/test_japan.mp4','/root/reference/test/test_korean.flv']
for v in vids:
    cap = cv2.VideoCapture(v); n = int(cap.get(cv2.CAP_PROP_FRAME_COUNT))
    for i in range(0, n, max(1, n // 25)):
        cap.set(cv2.CAP_PROP_POS_FRAMES, i); ok, fr = cap.read()
        if not ok: continue
        tot += stats(fr); nf += 1
    print(v, nf, tot)
from video_subtitle_extractor_b200.synth import SynthStream
s = np.zeros(3,int)
for i in range(0, 600, 15): s += stats(SynthStream(1080,1920).frame(i))
print('synthetic', s)
