"""CPU checks of the PRODUCT geometry source (csrc/geom.cuh + csrc/dbpost_core.cuh) against cv2 and the oracle.

The device kernels in csrc/postproc.cu call these very functions; tests/native/geom_host.cpp compiles them for the
CPU so the restatement of cv::convexHull / cv::minAreaRect / cv::boxPoints / cv::fillPoly / Clipper offset /
cv::warpPerspective(INTER_CUBIC) can be pinned without a GPU.  (The reference holds no tests for this path —
SURVEY.md §4 — so cv2 itself, the reference's own dependency, is the golden source here.)
"""
import ctypes as C
import os
import subprocess

import cv2
import numpy as np
import pytest

from oracle import hostlogic as hl

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "native", "geom_host.cpp")
OUT = os.path.join(ROOT, "tests", "native", "_build", "libgeom_host.so")


def _build():
    deps = [SRC] + [os.path.join(ROOT, "video_subtitle_extractor_b200", "csrc", f) for f in ("geom.cuh", "dbpost_core.cuh")]
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in deps):
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-o", OUT, SRC])
    lib = C.CDLL(OUT)
    lib.gh_mini_box_from_hull.restype = C.c_float
    return lib


@pytest.fixture(scope="module")
def lib():
    return _build()


ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
fp = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))


def random_prob_map(rng, h=160, w=320):
    m = np.zeros((h, w), np.float32)
    for _ in range(rng.integers(1, 6)):
        cx, cy = rng.integers(20, w - 20), rng.integers(10, h - 10)
        bw, bh = rng.integers(4, 120), rng.integers(3, 30)
        ang = rng.uniform(-30, 30) if rng.random() < 0.5 else 0
        box = cv2.boxPoints(((float(cx), float(cy)), (float(bw), float(bh)), float(ang)))
        cv2.fillPoly(m, [box.astype(np.int32)], 1.0)
    m = cv2.GaussianBlur(m, (0, 0), rng.uniform(0.5, 3))
    m += rng.normal(0, 0.08, m.shape).astype(np.float32)
    return np.clip(m, 0, 1)


def outer_components(bitmap):
    """[(contour, xl, xr, ymin)] for every OUTER contour of cv2.findContours, with the component's row extents."""
    cs, hier = cv2.findContours(bitmap * 255, cv2.RETR_CCOMP, cv2.CHAIN_APPROX_SIMPLE)
    _, lab = cv2.connectedComponents(bitmap, connectivity=8)
    out = []
    for ci, c in enumerate(cs):
        if hier[0][ci][3] != -1:
            continue
        x0, y0 = c[0, 0]
        ys, xs = np.nonzero(lab == lab[y0, x0])
        ymin, ymax = ys.min(), ys.max()
        xl = np.array([xs[ys == y].min() for y in range(ymin, ymax + 1)], np.int32)
        xr = np.array([xs[ys == y].max() for y in range(ymin, ymax + 1)], np.int32)
        out.append((c, xl, xr, int(ymin)))
    return out


def test_convex_hull_port_matches_cv2(lib):
    rng = np.random.default_rng(5)
    for _ in range(300):
        n = int(rng.integers(3, 40))
        pts = rng.integers(0, 30, size=(n, 2)).astype(np.int32)
        ref = cv2.convexHull(pts.reshape(-1, 1, 2), clockwise=False, returnPoints=False).ravel()
        hull = np.zeros(n, np.int32)
        m = lib.gh_convex_hull(ip(pts), n, ip(hull))
        # duplicated points may be represented by either index: compare coordinates
        assert m == len(ref)
        assert np.array_equal(pts[hull[:m]], pts[ref])


def test_component_hull_and_min_area_rect_match_cv2(lib):
    rng = np.random.default_rng(1)
    tot = bad_hull = bad_rect = bad_box = 0
    worst = 0.0
    for _ in range(120):
        bm = (random_prob_map(rng) > 0.3).astype(np.uint8)
        for c, xl, xr, ymin in outer_components(bm):
            rows = len(xl)
            out = np.zeros((2 * rows + 4, 2), np.int32)
            m = lib.gh_hull_rows(ip(xl), ip(xr), rows, ymin, ip(out))
            mine = np.ascontiguousarray(out[:m])
            ref = cv2.convexHull(c, clockwise=False, returnPoints=True).reshape(-1, 2)
            tot += 1
            if mine.shape != ref.shape or not (mine == ref).all():
                bad_hull += 1
                # same vertex set even when the start differs
                assert sorted(map(tuple, mine.tolist())) == sorted(map(tuple, ref.tolist()))
            box8, rect5 = np.zeros(8, np.float32), np.zeros(5, np.float32)
            ss = lib.gh_mini_box_from_hull(ip(mine), m, fp(box8), fp(rect5))
            r = cv2.minAreaRect(c)
            exact = False
            if m > 2:
                ref5 = np.array([r[0][0], r[0][1], r[1][0], r[1][1]], np.float32)
                exact = np.array_equal(ref5, rect5[:4]) and rect5[4] == np.float32(r[2])
                bad_rect += not exact
            exact = m > 2 and exact
            rbox, rss = hl.get_mini_boxes(c)
            rbox = np.array(rbox, np.float32).ravel()
            if exact:
                d = float(np.abs(rbox - box8).max())
                worst = max(worst, d)
                bad_box += d > 0
                assert np.float32(rss) == np.float32(ss)
    assert tot > 1000
    assert bad_hull <= 0.01 * tot          # only contours that revisit pixels start the hull elsewhere
    assert bad_rect == 0                   # centre/size/angle bit-exact
    assert worst == 0 and bad_box == 0     # and so are the get_mini_boxes corners


def test_min_area_rect_rotated_shapes_bit_exact(lib):
    """Rotated text boxes: cv2.minAreaRect / boxPoints bit for bit (the score mask truncates these corners to ints, so
    a 1-ulp difference can move a mask edge by a pixel)."""
    rng = np.random.default_rng(3)
    tot = 0
    for _ in range(3000):
        cx, cy = rng.uniform(50, 900), rng.uniform(50, 500)
        w, h, ang = rng.uniform(5, 400), rng.uniform(3, 60), rng.uniform(-90, 90)
        bp = cv2.boxPoints(((cx, cy), (w, h), ang)).astype(np.int32)
        m = np.zeros((600, 1000), np.uint8)
        cv2.fillPoly(m, [bp], 1)
        cs, _ = cv2.findContours(m, cv2.RETR_LIST, cv2.CHAIN_APPROX_SIMPLE)
        if not cs:
            continue
        c = cs[0]
        hull = np.ascontiguousarray(cv2.convexHull(c, clockwise=False, returnPoints=True).reshape(-1, 2).astype(np.int32))
        if len(hull) < 3:
            continue
        r = cv2.minAreaRect(c)
        box8, rect5 = np.zeros(8, np.float32), np.zeros(5, np.float32)
        lib.gh_mini_box_from_hull(ip(hull), len(hull), fp(box8), fp(rect5))
        ref5 = np.array([r[0][0], r[0][1], r[1][0], r[1][1], r[2]], np.float32)
        assert np.array_equal(ref5, rect5), (hull.tolist(), ref5, rect5)
        rbox, _ = hl.get_mini_boxes(c)
        assert np.array_equal(np.array(rbox, np.float32).ravel(), box8)
        tot += 1
    assert tot > 2500


def _fill_rows(lib, box, W, H):
    win, qx, qy = np.zeros(4, np.int32), np.zeros(4, np.int32), np.zeros(4, np.int32)
    lib.gh_score_window(fp(np.ascontiguousarray(box)), W, H, ip(win), ip(qx), ip(qy))
    xmin, ymin, xmax, ymax = win
    mask = np.zeros((ymax - ymin + 1, xmax - xmin + 1), np.uint8)
    b = box.copy()
    b[:, 0] -= xmin
    b[:, 1] -= ymin
    cv2.fillPoly(mask, b.reshape(1, -1, 2).astype("int32"), 1)
    assert np.array_equal(b.astype("int32")[:, 0], qx) and np.array_equal(b.astype("int32")[:, 1], qy)
    mine = np.zeros_like(mask)
    xa, xb = C.c_int(0), C.c_int(0)
    for y in range(mask.shape[0]):
        if lib.gh_quad_row_span(ip(qx), ip(qy), mask.shape[1], mask.shape[0], y, C.byref(xa), C.byref(xb)):
            mine[y, xa.value:xb.value + 1] = 1
    return mine, mask


def test_fill_poly_rows_match_cv2(lib):
    """box_score_fast mask: bit-exact with cv2.fillPoly, also for boxes cut by the map border (cv::clipLine restarts the
    Bresenham walk there and the scan-line edges take their x from the clipped end points)."""
    rng = np.random.default_rng(2)
    for _ in range(600):
        cx, cy = rng.uniform(60, 200), rng.uniform(50, 150)
        box = cv2.boxPoints(((cx, cy), (rng.uniform(3, 90), rng.uniform(3, 30)), rng.uniform(-90, 0)))
        box = np.array(hl.get_mini_boxes(box.reshape(-1, 1, 2))[0], np.float32)
        mine, mask = _fill_rows(lib, box, 260, 200)
        assert np.array_equal(mine, mask)
    for _ in range(300):
        cx, cy = rng.uniform(0, 128), rng.uniform(0, 64)
        box = cv2.boxPoints(((cx, cy), (rng.uniform(10, 90), rng.uniform(5, 30)), rng.uniform(-90, 0)))
        box = np.array(hl.get_mini_boxes(box.reshape(-1, 1, 2))[0], np.float32)
        mine, mask = _fill_rows(lib, box, 128, 64)
        assert np.array_equal(mine, mask)
    # arbitrary integer quads far outside small windows (the quad_row_span contract itself, not only mini boxes)
    xa, xb = C.c_int(0), C.c_int(0)
    for _ in range(4000):
        W, H = int(rng.integers(20, 140)), int(rng.integers(8, 50))
        cx, cy = rng.uniform(-10, W + 10), rng.uniform(-5, H + 5)
        ang = rng.uniform(-30, 30) if rng.random() < 0.6 else 0
        q = cv2.boxPoints(((cx, cy), (rng.uniform(10, 150), rng.uniform(4, 40)), ang)).astype(np.int32)
        ref = np.zeros((H, W), np.uint8)
        cv2.fillPoly(ref, [q], 1)
        qx, qy = np.ascontiguousarray(q[:, 0]), np.ascontiguousarray(q[:, 1])
        mine = np.zeros_like(ref)
        for y in range(H):
            if lib.gh_quad_row_span(ip(qx), ip(qy), W, H, y, C.byref(xa), C.byref(xb)):
                mine[y, xa.value:xb.value + 1] = 1
        assert np.array_equal(mine, ref), (W, H, q.tolist())


def test_db_candidate_pipeline_matches_oracle(lib):
    """component -> box -> (score by the oracle) -> unclip -> frame-space integer quad, vs oracle/hostlogic."""
    rng = np.random.default_rng(3)
    tot = bad = 0
    for _ in range(150):
        pm = random_prob_map(rng)
        H, W = pm.shape
        src_h, src_w = int(H * rng.uniform(1.0, 2.2)), int(W * rng.uniform(1.0, 2.2))
        ref_boxes, _ = hl.db_postprocess(pm, (src_h, src_w, 0, 0), box_thresh=0.0)
        ref_final = hl.filter_tag_det_res(ref_boxes, (src_h, src_w, 3))
        mine = []
        comps = outer_components((pm > 0.3).astype(np.uint8))
        for c, xl, xr, ymin in comps:
            box8 = np.zeros(8, np.float32)
            if not lib.gh_db_stage1(ip(xl), ip(xr), len(xl), ymin, fp(box8)):
                continue
            quad8, ipts = np.zeros(8, np.float32), np.zeros(8, np.int32)
            if lib.gh_db_stage3(fp(box8), C.c_float(1.5), W, H, src_w, src_h, fp(quad8), ip(ipts)):
                mine.append(quad8.reshape(4, 2).copy())
        # the oracle also walks hole contours (RETR_LIST); compare as sets of quads
        ref_set = {tuple(b.ravel().tolist()) for b in ref_final}
        mine_set = {tuple(b.ravel().tolist()) for b in mine}
        tot += len(mine_set)
        bad += len(mine_set - ref_set)
    assert tot > 300
    assert bad <= 0.005 * tot


def test_warp_cubic_matches_cv2(lib):
    rng = np.random.default_rng(4)
    frame = rng.integers(0, 256, (120, 200, 3), dtype=np.uint8)
    frame = cv2.GaussianBlur(frame, (0, 0), 1.2)
    for k in range(40):
        cx, cy = rng.uniform(40, 160), rng.uniform(30, 90)
        ang = 0.0 if k % 4 == 0 else rng.uniform(-25, 25)
        box = cv2.boxPoints(((cx, cy), (rng.uniform(20, 120), rng.uniform(8, 40)), ang))
        quad = hl.order_points_clockwise(np.round(box).astype(np.float32))
        ref = hl.get_rotate_crop_image(frame, quad.copy())
        if ref.shape[0] * 1.0 / ref.shape[1] >= 1.5:
            continue  # rot90 is applied by the caller
        cw, ch = C.c_int(0), C.c_int(0)
        lib.gh_crop_size(fp(np.ascontiguousarray(quad)), C.byref(cw), C.byref(ch))
        assert (ch.value, cw.value) == ref.shape[:2]
        out = np.zeros((ch.value, cw.value, 3), np.uint8)
        lib.gh_warp_crop(frame.ctypes.data_as(C.c_void_p), frame.shape[0], frame.shape[1], frame.strides[0],
                         fp(np.ascontiguousarray(quad)), out.ctypes.data_as(C.c_void_p), cw.value, ch.value)
        diff = np.abs(out.astype(int) - ref.astype(int))
        assert diff.max() <= 1 and (diff > 0).mean() < 0.002, (k, diff.max(), (diff > 0).mean())
