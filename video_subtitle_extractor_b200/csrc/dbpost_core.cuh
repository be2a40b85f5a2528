// dbpost_core.cuh — per-candidate stages of the DB post-process and the bicubic crop sampler (host/device).
//
// Restates paddleocr 2.10 DBPostProcess.boxes_from_bitmap (score_mode 'fast', box_type 'quad'), TextDetector.
// filter_tag_det_res and TextSystem.get_rotate_crop_image as driven by reference backend/tools/ocr.py:27 and
// backend/tools/subtitle_detect.py:24-26, on top of geom.cuh:
//   stage 1  component row extents -> convex hull -> minAreaRect -> get_mini_boxes          (one thread)
//   stage 2  box_score_fast: mean probability under cv::fillPoly(quad)                       (whole block, postproc.cu)
//   stage 3  unclip (Clipper round offset) -> minAreaRect -> scale/round/clip -> ordered quad (one thread)
// and cv::warpPerspective(INTER_CUBIC, BORDER_REPLICATE) for the text-line crops.
// The same source is compiled for the CPU by tests/native/geom_host.cpp and checked against cv2.
#pragma once
#include "geom.cuh"

namespace vse {
namespace dbpost {

using geom::P2f;
using geom::P2i;

struct Candidate {
    float box[8];    // get_mini_boxes corners (map space): tl, tr, br, bl
    float sside;
    float score;
    float quad[8];   // final ordered quad in frame pixels (float32 integers)
    int ipts[8];     // DBPostProcess integer corners before order_points_clockwise (debug)
};

// `hull` needs 2*rows + 2 points, `work` 5 * (2*rows + 2) floats
VSE_HD inline bool stage1_component_box(const int* xl, const int* xr, int rows, int y0, P2i* hull, float* work,
                                        float min_size, Candidate* c) {
    int m = geom::hull_from_row_extents(xl, xr, rows, y0, hull);
    P2f* hf = reinterpret_cast<P2f*>(work);
    for (int i = 0; i < m; i++) { hf[i].x = (float)hull[i].x; hf[i].y = (float)hull[i].y; }
    geom::RotRect r = geom::cv_min_area_rect(hf, m, work + 2 * m);
    P2f bp[4], box[4];
    geom::cv_box_points(r, bp);
    geom::mini_box_order(bp, box);
    for (int i = 0; i < 4; i++) { c->box[2 * i] = box[i].x; c->box[2 * i + 1] = box[i].y; }
    c->sside = r.w < r.h ? r.w : r.h;
    return !(c->sside < min_size);
}

// box_score_fast window: clipped floor/ceil bounding box of the 4 corners and the window-relative integer quad
VSE_HD inline void score_window(const P2f* box, int rw, int rh, int* win /*xmin,ymin,xmax,ymax*/, int* qx, int* qy) {
    float minx = box[0].x, maxx = box[0].x, miny = box[0].y, maxy = box[0].y;
    for (int i = 1; i < 4; i++) {
        minx = fminf(minx, box[i].x); maxx = fmaxf(maxx, box[i].x);
        miny = fminf(miny, box[i].y); maxy = fmaxf(maxy, box[i].y);
    }
    auto clampi = [](int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); };
    int xmin = clampi((int)floorf(minx), 0, rw - 1), xmax = clampi((int)ceilf(maxx), 0, rw - 1);
    int ymin = clampi((int)floorf(miny), 0, rh - 1), ymax = clampi((int)ceilf(maxy), 0, rh - 1);
    win[0] = xmin; win[1] = ymin; win[2] = xmax; win[3] = ymax;
    for (int i = 0; i < 4; i++) {
        qx[i] = (int)(box[i].x - (float)xmin);  // float32 subtract, astype(int32) truncation
        qy[i] = (int)(box[i].y - (float)ymin);
    }
}

VSE_HD inline long long floor_div(long long a, long long b) {  // b > 0
    long long q = a / b;
    return (a % b != 0 && a < 0) ? q - 1 : q;
}

// cv::clipLine(Size(w, h), pt1, pt2) on 64-bit points (imgproc/drawing.cpp): Cohen-Sutherland with double-precision
// intersections truncated toward zero.  Returns true when (part of) the segment is inside; the points are updated in place
// exactly as OpenCV leaves them (also when the answer is false).
VSE_HD inline bool cv_clip_line(long long w, long long h, long long& x1, long long& y1, long long& x2, long long& y2) {
    if (w <= 0 || h <= 0) return false;
    const long long right = w - 1, bottom = h - 1;
    int c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8;
    int c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8;
    if ((c1 & c2) == 0 && (c1 | c2) != 0) {
        long long a;
        if (c1 & 12) {
            a = c1 < 8 ? 0 : bottom;
            x1 += (long long)((double)(a - y1) * (double)(x2 - x1) / (double)(y2 - y1));
            y1 = a;
            c1 = (x1 < 0) + (x1 > right) * 2;
        }
        if (c2 & 12) {
            a = c2 < 8 ? 0 : bottom;
            x2 += (long long)((double)(a - y2) * (double)(x2 - x1) / (double)(y2 - y1));
            y2 = a;
            c2 = (x2 < 0) + (x2 > right) * 2;
        }
        if ((c1 & c2) == 0 && (c1 | c2) != 0) {
            if (c1) {
                a = c1 == 1 ? 0 : right;
                y1 += (long long)((double)(a - x1) * (double)(y2 - y1) / (double)(x2 - x1));
                x1 = a;
                c1 = 0;
            }
            if (c2) {
                a = c2 == 1 ? 0 : right;
                y2 += (long long)((double)(a - x2) * (double)(y2 - y1) / (double)(x2 - x1));
                x2 = a;
                c2 = 0;
            }
        }
    }
    return (c1 | c2) == 0;
}

// Row `y` of cv::fillPoly(mask(h x w), quad) for a window-relative integer quad whose corners may lie outside the
// window (a text box cut by the map border) — bit for bit what OpenCV 4.13 draws (checked on 20 000 random quads,
// tests/test_geom_cpu.py): per edge
//   * the 8-connected boundary line is drawn between the cv::clipLine'd end points (cv::LineIterator, left to right,
//     ties keep the minor coordinate) — clipping restarts the Bresenham walk;
//   * the scan-line edge takes its x from the clipped end points ALWAYS and its y from them only when the clipped points
//     differ in y; slope = truncated 16.16 quotient; start x extrapolated back to the unclipped top row;
//   * a row is filled from ceil(x_left) to floor(x_right) (16.16), clamped to the window.
// For a convex quad the union on one row is a single interval [xa, xb], already clipped to the window.
VSE_HD inline bool quad_row_span(const int* qx, const int* qy, int w, int h, int y, int* xa, int* xb) {
    if (y < 0 || y >= h) return false;
    bool any = false;
    long long lo = 0, hi = 0;
    auto add = [&](long long a, long long b) {
        if (a < 0) a = 0;
        if (b > w - 1) b = w - 1;
        if (a > b) return;
        if (!any) { lo = a; hi = b; any = true; }
        else { if (a < lo) lo = a; if (b > hi) hi = b; }
    };
    long long ex[4];
    int ne = 0;
    for (int i = 0; i < 4; i++) {
        const int j = (i + 3) & 3;  // pt0 = v[i-1], pt1 = v[i]
        const long long x0 = qx[j], y0 = qy[j], x1 = qx[i], y1 = qy[i];
        const bool outside = x0 < 0 || x0 >= w || x1 < 0 || x1 >= w || y0 < 0 || y0 >= h || y1 < 0 || y1 >= h;
        long long cx0 = x0, cy0 = y0, cx1 = x1, cy1 = y1;
        bool visible = true;
        if (outside) visible = cv_clip_line(w, h, cx0, cy0, cx1, cy1);
        // ---- boundary line between the clipped end points
        if (visible) {
            long long ax = cx0, ay = cy0, bx = cx1, by = cy1;
            if (bx < ax) { long long t = ax; ax = bx; bx = t; t = ay; ay = by; by = t; }
            const long long adx = bx - ax, dy = by - ay, ady = dy < 0 ? -dy : dy;
            const int sy = dy < 0 ? -1 : 1;
            const long long t = (long long)(y - ay) * sy;
            if (t >= 0 && t <= ady) {
                if (adx >= ady) {  // x-major
                    if (ady == 0) {
                        add(ax, bx);
                    } else {
                        long long kmin = floor_div(adx * (2 * t - 1), 2 * ady) + 1;
                        long long kmax = floor_div(adx * (2 * t + 1), 2 * ady);
                        if (kmin < 0) kmin = 0;
                        if (kmax > adx) kmax = adx;
                        add(ax + kmin, ax + kmax);
                    }
                } else {  // y-major: one pixel per row
                    const long long num = 2 * t * adx - ady, den = 2 * ady;
                    long long m = -floor_div(-num, den);  // ceil
                    if (m < 0) m = 0;
                    add(ax + m, ax + m);
                }
            }
        }
        // ---- scan-line edge
        if (y0 == y1) continue;
        long long p0x = x0 << 16, p1x = x1 << 16, p0y = y0, p1y = y1;
        if (outside) {
            if (cy0 != cy1) { p0y = cy0; p1y = cy1; }
            p0x = cx0 << 16;
            p1x = cx1 << 16;
        }
        const long long dx = (p1x - p0x) / (p1y - p0y);   // C++ division: truncates toward zero, like OpenCV's
        long long ya, yb, xs;
        if (y0 < y1) { ya = y0; yb = y1; xs = p0x + (y0 - p0y) * dx; }
        else { ya = y1; yb = y0; xs = p1x + (y1 - p1y) * dx; }
        if (y >= ya && y < yb) ex[ne++] = xs + (long long)(y - ya) * dx;
    }
    if (ne >= 2) {
        long long x1 = ex[0], x2 = ex[0];
        for (int i = 1; i < ne; i++) { if (ex[i] < x1) x1 = ex[i]; if (ex[i] > x2) x2 = ex[i]; }
        const long long l = (x1 + 65535) >> 16, r = x2 >> 16;
        if (l < w && r >= 0) add(l, r);
    }
    if (!any) return false;
    *xa = (int)lo;
    *xb = (int)hi;
    return true;
}

// stage 3: unclip, second rectangle, rescale to the frame, order/clip/filter.
VSE_HD inline bool stage3_unclip_scale(Candidate* c, float unclip_ratio, float min_size, int rw, int rh, int src_w, int src_h) {
    P2f box[4];
    for (int i = 0; i < 4; i++) { box[i].x = c->box[2 * i]; box[i].y = c->box[2 * i + 1]; }
    double area, length;
    geom::polygon_area_length(box, 4, &area, &length);
    double distance = area * (double)unclip_ratio / length;
    const int CAP = 96;
    P2i pts[CAP];
    int n = geom::clipper_offset_round(box, 4, distance, pts, CAP);
    if (n < 3) return false;
    int order[CAP], stack[CAP + 2], hull[CAP];
    int m = geom::cv_convex_hull(pts, n, order, stack, hull);
    P2f hf[CAP];
    float work[3 * CAP];
    for (int i = 0; i < m; i++) { hf[i].x = (float)pts[hull[i]].x; hf[i].y = (float)pts[hull[i]].y; }
    geom::RotRect r = geom::cv_min_area_rect(hf, m, work);
    float sside = r.w < r.h ? r.w : r.h;
    if (sside < min_size + 2.f) return false;
    P2f bp[4], b2[4];
    geom::cv_box_points(r, bp);
    geom::mini_box_order(bp, b2);
    P2i ip[4];
    for (int i = 0; i < 4; i++) {
        float fx = rintf((b2[i].x / (float)rw) * (float)src_w);
        float fy = rintf((b2[i].y / (float)rh) * (float)src_h);
        fx = fminf(fmaxf(fx, 0.f), (float)src_w);
        fy = fminf(fmaxf(fy, 0.f), (float)src_h);
        ip[i].x = (int)fx;
        ip[i].y = (int)fy;
        c->ipts[2 * i] = ip[i].x;
        c->ipts[2 * i + 1] = ip[i].y;
    }
    P2f q[4];
    bool keep = geom::order_clip_filter(ip, src_h, src_w, q);
    for (int i = 0; i < 4; i++) { c->quad[2 * i] = q[i].x; c->quad[2 * i + 1] = q[i].y; }
    return keep;
}

// ------------------------------------------------------------------------------------------------------------------
// cv::warpPerspective(INTER_CUBIC, BORDER_REPLICATE) on 8-bit pixels
// ------------------------------------------------------------------------------------------------------------------
#ifndef VSE_CUBIC_FIX_K0
#define VSE_CUBIC_FIX_K0 2
#endif
inline int sat_short_round(float v) {
    double r = nearbyint((double)v);
    if (r < -32768.) r = -32768.;
    if (r > 32767.) r = 32767.;
    return (int)r;
}

// 32x32 table of 4x4 15-bit weights (cv::initInterTab2D(INTER_CUBIC, fixpt=true)); host-only, uploaded once
inline void build_cubic_table(short* itab /*32*32*16*/) {
    float tab1[32 * 4];
    const float A = -0.75f;
    const float scale = 1.f / 32;
    for (int i = 0; i < 32; i++) {
        float x = i * scale;
        float* co = tab1 + i * 4;
        co[0] = ((A * (x + 1) - 5 * A) * (x + 1) + 8 * A) * (x + 1) - 4 * A;
        co[1] = ((A + 2) * x - (A + 3)) * x * x + 1;
        co[2] = ((A + 2) * (1 - x) - (A + 3)) * (1 - x) * (1 - x) + 1;
        co[3] = 1.f - co[0] - co[1] - co[2];
    }
    const int ksize = 4;
    for (int i = 0; i < 32; i++)
        for (int j = 0; j < 32; j++) {
            short* it = itab + (i * 32 + j) * 16;
            int isum = 0;
            for (int k1 = 0; k1 < ksize; k1++) {
                float vy = tab1[i * ksize + k1];
                for (int k2 = 0; k2 < ksize; k2++) {
                    float v = vy * tab1[j * ksize + k2];
                    int iv = sat_short_round(v * 32768.f);
                    it[k1 * ksize + k2] = (short)iv;
                    isum += iv;
                }
            }
            if (isum != 32768) {
                int diff = isum - 32768;
                const int k0 = VSE_CUBIC_FIX_K0;  // first row/column of the 2x2 block searched for the fix-up tap
                int Mk1 = 2, Mk2 = 2, mk1 = 2, mk2 = 2;
                for (int k1 = k0; k1 < k0 + 2; k1++)
                    for (int k2 = k0; k2 < k0 + 2; k2++) {
                        if (it[k1 * ksize + k2] < it[mk1 * ksize + mk2]) { mk1 = k1; mk2 = k2; }
                        else if (it[k1 * ksize + k2] > it[Mk1 * ksize + Mk2]) { Mk1 = k1; Mk2 = k2; }
                    }
                if (diff < 0) it[Mk1 * ksize + Mk2] = (short)(it[Mk1 * ksize + Mk2] - diff);
                else it[mk1 * ksize + mk2] = (short)(it[mk1 * ksize + mk2] - diff);
            }
        }
}

VSE_HD inline int round_half_even_d(double v) {
#if defined(__CUDA_ARCH__)
    return __double2int_rn(v);
#else
    return (int)nearbyint(v);
#endif
}

// one destination pixel (x, y) of the crop; M maps crop -> frame coordinates; `pix` = bytes per frame pixel (3 or 4)
VSE_HD inline void warp_cubic_pixel(const unsigned char* frame, int fh, int fw, int stride, int pix, const double* M,
                                    const short* itab, int x, int y, unsigned char* out3) {
    double X0 = M[0] * x + M[1] * y + M[2];
    double Y0 = M[3] * x + M[4] * y + M[5];
    double W = M[6] * x + M[7] * y + M[8];
    W = W ? 32. / W : 0;
    double fX = fmax(-2147483648., fmin(2147483647., X0 * W));
    double fY = fmax(-2147483648., fmin(2147483647., Y0 * W));
    int X = round_half_even_d(fX), Y = round_half_even_d(fY);
    int sx = (X >> 5) - 1, sy = (Y >> 5) - 1;
    if (sx < -32769) sx = -32769; if (sx > 32766) sx = 32766;   // saturate_cast<short>(X >> 5) - 1
    if (sy < -32769) sy = -32769; if (sy > 32766) sy = 32766;
    const short* w = itab + ((Y & 31) * 32 + (X & 31)) * 16;
    int sum[3] = {0, 0, 0};
    for (int k1 = 0; k1 < 4; k1++) {
        int yy = sy + k1;
        yy = yy < 0 ? 0 : (yy > fh - 1 ? fh - 1 : yy);
        const unsigned char* row = frame + (size_t)yy * stride;
        for (int k2 = 0; k2 < 4; k2++) {
            int xx = sx + k2;
            xx = xx < 0 ? 0 : (xx > fw - 1 ? fw - 1 : xx);
            int wt = w[k1 * 4 + k2];
            const unsigned char* p = row + xx * pix;
            sum[0] += p[0] * wt;
            sum[1] += p[1] * wt;
            sum[2] += p[2] * wt;
        }
    }
    for (int c = 0; c < 3; c++) {
        int v = (sum[c] + (1 << 14)) >> 15;
        out3[c] = (unsigned char)(v < 0 ? 0 : (v > 255 ? 255 : v));
    }
}

}  // namespace dbpost
}  // namespace vse
