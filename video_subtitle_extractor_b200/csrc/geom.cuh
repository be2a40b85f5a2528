// geom.cuh — host/device geometry used by the device-side DB post-process and the crop kernel.
//
// Everything here restates, operation by operation, what the reference's detector post-process asks of
// OpenCV / pyclipper (paddleocr 2.10 DBPostProcess, driven by reference backend/tools/subtitle_detect.py:24-26 and
// backend/tools/ocr.py:27):
//   cv::convexHull (Sklansky)            -> cv_convex_hull / hull_from_row_extents
//   cv::minAreaRect (rotating calipers)  -> cv_min_area_rect
//   cv::boxPoints                        -> cv_box_points
//   DBPostProcess.get_mini_boxes         -> mini_box_order
//   pyclipper JT_ROUND closed-polygon offset (ClipperLib 6.4.2 DoOffset/DoRound) -> clipper_offset_round
//   TextDetector.order_points_clockwise / filter_tag_det_res / sorted_boxes key
// The functions are plain C++ (`VSE_HD`) so that tests/native/geom_host.cpp compiles THE SAME source for the CPU and
// checks it against cv2 on thousands of random shapes (tests/test_geom_cpu.py); the CUDA kernels in postproc.cu call
// them from one thread per candidate.  Compile with FMA contraction off (nvcc --fmad=false, gcc -ffp-contract=off):
// OpenCV's float code is built without contraction and the box corners must match to the last bit.
#pragma once
#include <float.h>
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define VSE_HD __host__ __device__
#else
#define VSE_HD
#endif

namespace vse {
namespace geom {

struct P2i { int x, y; };
struct P2f { float x, y; };
struct RotRect { float cx, cy, w, h, angle; };

#ifndef VSE_AREA_CMP
#define VSE_AREA_CMP(a, m) ((a) <= (m))
#endif
#define VSE_PI 3.1415926535897932384626433832795

VSE_HD inline int sgn_ll(long long v) { return (v > 0) - (v < 0); }
VSE_HD inline int sgn_i(int v) { return (v > 0) - (v < 0); }

// ------------------------------------------------------------------------------------------------------------------
// cv::convexHull(points, hull, clockwise=false, returnPoints=false) for integer points.
// `order` (n ints) and `stack` (n + 2 ints) are scratch; `hull` receives indices into pts. Returns the hull size.
// ------------------------------------------------------------------------------------------------------------------
VSE_HD inline bool hull_less(const P2i* pts, int a, int b) {
    if (pts[a].x != pts[b].x) return pts[a].x < pts[b].x;
    if (pts[a].y != pts[b].y) return pts[a].y < pts[b].y;
    return a < b;
}

VSE_HD inline int sklansky(const P2i* pts, const int* order, int start, int end, int* stack, int nsign, int sign2) {
    int incr = end > start ? 1 : -1;
    int pprev = start, pcur = pprev + incr, pnext = pcur + incr;
    int stacksize = 3;
    if (start == end || (pts[order[start]].x == pts[order[end]].x && pts[order[start]].y == pts[order[end]].y)) {
        stack[0] = start;
        return 1;
    }
    stack[0] = pprev;
    stack[1] = pcur;
    stack[2] = pnext;
    end += incr;
    while (pnext != end) {
        int cury = pts[order[pcur]].y;
        int nexty = pts[order[pnext]].y;
        int by = nexty - cury;
        if (sgn_i(by) != nsign) {
            int ax = pts[order[pcur]].x - pts[order[pprev]].x;
            int bx = pts[order[pnext]].x - pts[order[pcur]].x;
            int ay = cury - pts[order[pprev]].y;
            long long convexity = (long long)ay * bx - (long long)ax * by;
            if (sgn_ll(convexity) == sign2 && (ax != 0 || ay != 0)) {
                pprev = pcur;
                pcur = pnext;
                pnext += incr;
                stack[stacksize] = pnext;
                stacksize++;
            } else {
                if (pprev == start) {
                    pcur = pnext;
                    stack[1] = pcur;
                    pnext += incr;
                    stack[2] = pnext;
                } else {
                    stack[stacksize - 2] = pnext;
                    pcur = pprev;
                    pprev = stack[stacksize - 4];
                    stacksize--;
                }
            }
        } else {
            pnext += incr;
            stack[stacksize - 1] = pnext;
        }
    }
    return --stacksize;
}

VSE_HD inline int cv_convex_hull(const P2i* pts, int n, int* order, int* stack, int* hull) {
    if (n <= 0) return 0;
    // sort indices by (x, y, index): insertion sort (n is small wherever this generic path is used)
    for (int i = 0; i < n; i++) {
        int v = i, j = i - 1;
        while (j >= 0 && hull_less(pts, v, order[j])) {
            order[j + 1] = order[j];
            j--;
        }
        order[j + 1] = v;
    }
    int miny_ind = 0, maxy_ind = 0;
    for (int i = 1; i < n; i++) {
        int y = pts[order[i]].y;
        if (pts[order[miny_ind]].y > y) miny_ind = i;
        if (pts[order[maxy_ind]].y < y) maxy_ind = i;
    }
    int nout = 0;
    if (pts[order[0]].x == pts[order[n - 1]].x && pts[order[0]].y == pts[order[n - 1]].y) {
        hull[nout++] = order[0];
        return nout;
    }
    // upper half
    int* tl_stack = stack;
    int tl_count = sklansky(pts, order, 0, maxy_ind, tl_stack, -1, 1);
    int* tr_stack = stack + tl_count;
    int tr_count = sklansky(pts, order, n - 1, maxy_ind, tr_stack, -1, -1);
    {   // clockwise == false: swap
        int* t = tl_stack; tl_stack = tr_stack; tr_stack = t;
        int c = tl_count; tl_count = tr_count; tr_count = c;
    }
    for (int i = 0; i < tl_count - 1; i++) hull[nout++] = order[tl_stack[i]];
    for (int i = tr_count - 1; i > 0; i--) hull[nout++] = order[tr_stack[i]];
    int stop_idx = tr_count > 2 ? tr_stack[1] : tl_count > 2 ? tl_stack[tl_count - 2] : -1;
    // the lower half reuses `stack`, so remember what stop_idx points at
    P2i stop_pt = {0, 0};
    if (stop_idx >= 0) stop_pt = pts[order[stop_idx]];
    // lower half
    int* bl_stack = stack;
    int bl_count = sklansky(pts, order, 0, miny_ind, bl_stack, 1, -1);
    int* br_stack = stack + bl_count;
    int br_count = sklansky(pts, order, n - 1, miny_ind, br_stack, 1, 1);
    if (stop_idx >= 0) {
        int check_idx = bl_count > 2 ? bl_stack[1] : bl_count + br_count > 2 ? br_stack[2 - bl_count] : -1;
        if (check_idx == stop_idx ||
            (check_idx >= 0 && pts[order[check_idx]].x == stop_pt.x && pts[order[check_idx]].y == stop_pt.y)) {
            bl_count = bl_count < 2 ? bl_count : 2;
            br_count = br_count < 2 ? br_count : 2;
        }
    }
    for (int i = 0; i < bl_count - 1; i++) hull[nout++] = order[bl_stack[i]];
    for (int i = br_count - 1; i > 0; i--) hull[nout++] = order[br_stack[i]];
    // cyclic shift so that the indices form an ascending / descending sequence
    if (nout >= 3) {
        int min_idx = 0, max_idx = 0, lt = 0, i;
        for (i = 1; i < nout; i++) {
            int idx = hull[i];
            lt += hull[i - 1] < idx;
            if (lt > 1 && lt <= i - 2) break;
            if (idx < hull[min_idx]) min_idx = i;
            if (idx > hull[max_idx]) max_idx = i;
        }
        int mmdist = max_idx - min_idx;
        if (mmdist < 0) mmdist = -mmdist;
        if ((mmdist == 1 || mmdist == nout - 1) && (lt <= 1 || lt >= nout - 2)) {
            int ascending = (max_idx + 1) % nout == min_idx;
            int i0 = ascending ? min_idx : max_idx, j = i0;
            if (i0 > 0) {
                for (i = 0; i < nout; i++) {
                    int curr_idx = stack[i] = hull[j];
                    int next_j = j + 1 < nout ? j + 1 : 0;
                    int next_idx = hull[next_j];
                    if (i < nout - 1 && (ascending != (curr_idx < next_idx))) break;
                    j = next_j;
                }
                if (i == nout)
                    for (i = 0; i < nout; i++) hull[i] = stack[i];
            }
        }
    }
    return nout;
}

// ------------------------------------------------------------------------------------------------------------------
// Convex hull of an 8-connected component given its per-row extents xl[y], xr[y] (y = 0..rows-1, offset by y0).
// Output order = what cv::convexHull returns for the component's outer cv::findContours contour in the generic case
// (contour traced from the top-left-most pixel, counter-clockwise on screen; hull indices descending => the hull starts
// at the screen-clockwise neighbour of the start pixel and ends with the start pixel).  `out` needs 2*rows + 2 entries.
// ------------------------------------------------------------------------------------------------------------------
VSE_HD inline long long cross3(P2i a, P2i b, P2i c) {
    return (long long)(b.x - a.x) * (c.y - b.y) - (long long)(b.y - a.y) * (c.x - b.x);
}

VSE_HD inline int hull_from_row_extents(const int* xl, const int* xr, int rows, int y0, P2i* out) {
    // closed walk: left side top->bottom, right side bottom->top (screen counter-clockwise, cross < 0 at convex corners)
    int m = 0;
    auto push = [&](P2i p, int floor_m) {
        if (m > 0 && out[m - 1].x == p.x && out[m - 1].y == p.y) return;
        while (m - floor_m >= 2 && cross3(out[m - 2], out[m - 1], p) >= 0) m--;
        out[m++] = p;
    };
    for (int r = 0; r < rows; r++) push(P2i{xl[r], y0 + r}, 0);
    push(P2i{xr[rows - 1], y0 + rows - 1}, 0);
    int floor_m = m - 1;  // the bottom-right point is a hull vertex: the right chain never pops below it
    for (int r = rows - 2; r >= 0; r--) push(P2i{xr[r], y0 + r}, floor_m);
    // closing edge back to the start: drop trailing points that are not strictly convex w.r.t. out[0]
    while (m - floor_m >= 2 && m >= 3 && cross3(out[m - 2], out[m - 1], out[0]) >= 0) m--;
    if (m >= 2 && out[m - 1].x == out[0].x && out[m - 1].y == out[0].y) m--;
    // also the first vertex after the start may be collinear with (last, start)
    if (m >= 3 && cross3(out[m - 1], out[0], out[1]) >= 0) {
        // start pixel is always an extreme point; collinearity here means a degenerate (flat) hull
    }
    // reverse into OpenCV's order: [H[m-1], ..., H[1], H[0]]
    for (int i = 0, j = m - 1; i < j; i++, j--) {
        P2i t = out[i]; out[i] = out[j]; out[j] = t;
    }
    return m;
}

// ------------------------------------------------------------------------------------------------------------------
// cv::minAreaRect on an already-computed hull (float points in cv::convexHull order).  `work` needs 3*n floats.
// ------------------------------------------------------------------------------------------------------------------
VSE_HD inline void rotating_calipers_minarea(const P2f* points, int n, float* work, float* out /*6*/) {
    float minarea = FLT_MAX;
    float buf_f[7];
    int buf_left = 0, buf_bottom = 0;
    for (int i = 0; i < 7; i++) buf_f[i] = 0.f;
    float* inv_vect_length = work;
    P2f* vect = reinterpret_cast<P2f*>(work + n);
    int left = 0, bottom = 0, right = 0, top = 0;
    int seq[4] = {-1, -1, -1, -1};
    float orientation = 0;
    float base_a;
    float base_b = 0;
    float left_x, right_x, top_y, bottom_y;
    P2f pt0 = points[0];
    left_x = right_x = pt0.x;
    top_y = bottom_y = pt0.y;
    for (int i = 0; i < n; i++) {
        double dx, dy;
        if (pt0.x < left_x) left_x = pt0.x, left = i;
        if (pt0.x > right_x) right_x = pt0.x, right = i;
        if (pt0.y > top_y) top_y = pt0.y, top = i;
        if (pt0.y < bottom_y) bottom_y = pt0.y, bottom = i;
        P2f pt = points[(i + 1) & (i + 1 < n ? -1 : 0)];
        dx = (double)pt.x - (double)pt0.x;
        dy = (double)pt.y - (double)pt0.y;
        vect[i].x = (float)dx;
        vect[i].y = (float)dy;
        inv_vect_length[i] = (float)(1. / sqrt(dx * dx + dy * dy));
        pt0 = pt;
    }
    {
        double ax = vect[n - 1].x;
        double ay = vect[n - 1].y;
        for (int i = 0; i < n; i++) {
            double bx = vect[i].x;
            double by = vect[i].y;
            double convexity = ax * by - ay * bx;
            if (convexity != 0) {
                orientation = (convexity > 0) ? 1.f : (-1.f);
                break;
            }
            ax = bx;
            ay = by;
        }
    }
    base_a = orientation;
    seq[0] = bottom;
    seq[1] = right;
    seq[2] = top;
    seq[3] = left;
    for (int k = 0; k < n; k++) {
        // OpenCV >= 4.5.2 (rotcalipers.cpp): the caliper side with the smallest angle to its polygon edge is found by
        // cross-product sign tests between the edge vectors rotated into the frame of side 0, not by comparing cosines
        P2f rv[4];
        rv[0] = vect[seq[0]];
        rv[1].x = vect[seq[1]].y;  rv[1].y = -vect[seq[1]].x;   // rotate90CW
        rv[2].x = -vect[seq[2]].x; rv[2].y = -vect[seq[2]].y;   // rotate180
        rv[3].x = -vect[seq[3]].y; rv[3].y = vect[seq[3]].x;    // rotate90CCW
        int main_element = 0;
        for (int i = 1; i < 4; ++i) {
            // firstVecIsRight(rv[i], rv[main_element]): rotate90CW(rv[i]) . rv[main] < 0
            float tx = rv[i].y, ty = -rv[i].x;
            if (tx * rv[main_element].x + ty * rv[main_element].y < 0) main_element = i;
        }
        {
            int pindex = seq[main_element];
            float lead_x = vect[pindex].x * inv_vect_length[pindex];
            float lead_y = vect[pindex].y * inv_vect_length[pindex];
            switch (main_element) {
                case 0: base_a = lead_x; base_b = lead_y; break;
                case 1: base_a = lead_y; base_b = -lead_x; break;
                case 2: base_a = -lead_x; base_b = -lead_y; break;
                default: base_a = -lead_y; base_b = lead_x; break;
            }
        }
        seq[main_element] += 1;
        seq[main_element] = (seq[main_element] == n) ? 0 : seq[main_element];
        {
            float dx = points[seq[1]].x - points[seq[3]].x;
            float dy = points[seq[1]].y - points[seq[3]].y;
            float width = dx * base_a + dy * base_b;
            dx = points[seq[2]].x - points[seq[0]].x;
            dy = points[seq[2]].y - points[seq[0]].y;
            float height = -dx * base_b + dy * base_a;
            float area = width * height;
            if (VSE_AREA_CMP(area, minarea)) {
                minarea = area;
                buf_left = seq[3];
                buf_f[1] = base_a;
                buf_f[2] = width;
                buf_f[3] = base_b;
                buf_f[4] = height;
                buf_bottom = seq[0];
                buf_f[6] = area;
            }
        }
    }
    {
        float A1 = buf_f[1];
        float B1 = buf_f[3];
        float A2 = -buf_f[3];
        float B2 = buf_f[1];
        float C1 = A1 * points[buf_left].x + points[buf_left].y * B1;
        float C2 = A2 * points[buf_bottom].x + points[buf_bottom].y * B2;
        float idet = 1.f / (A1 * B2 - A2 * B1);
        float px = (C1 * B2 - C2 * B1) * idet;
        float py = (A1 * C2 - A2 * C1) * idet;
        out[0] = px;
        out[1] = py;
        out[2] = A1 * buf_f[2];
        out[3] = B1 * buf_f[2];
        out[4] = A2 * buf_f[4];
        out[5] = B2 * buf_f[4];
    }
}

VSE_HD inline RotRect cv_min_area_rect(const P2f* hpoints, int n, float* work) {
    RotRect box = {0.f, 0.f, 0.f, 0.f, 0.f};
    if (n > 2) {
        float out[6];
        rotating_calipers_minarea(hpoints, n, work, out);
        box.cx = out[0] + (out[2] + out[4]) * 0.5f;
        box.cy = out[1] + (out[3] + out[5]) * 0.5f;
        // OpenCV >= 4.5.1 reports the side whose direction lies in [-90, 0) degrees as "width": while the first side
        // vector points into [0, 180], step to the next side (a, b) -> (-b, a) and take 90 degrees off the angle, which
        // is kept in double until the final cast.  Checked against cv2 4.13 on axis-aligned and rotated shapes: centre,
        // size and angle bit-exact (tests/test_geom_cpu.py).
        float ax = out[2], ay = out[3], bx = out[4], by = out[5];
        double deg = atan2((double)ay, (double)ax) * 180 / VSE_PI;
        for (int it = 0; it < 4 && deg >= 0; it++) {
            float tx = ax, ty = ay;
            ax = -bx; ay = -by; bx = tx; by = ty;
            deg -= 90;
        }
        box.w = (float)sqrt((double)ax * ax + (double)ay * ay);
        box.h = (float)sqrt((double)bx * bx + (double)by * by);
        box.angle = (float)deg;
        return box;
    } else if (n == 2) {
        box.cx = (hpoints[0].x + hpoints[1].x) * 0.5f;
        box.cy = (hpoints[0].y + hpoints[1].y) * 0.5f;
        double dx = (double)hpoints[1].x - (double)hpoints[0].x;
        double dy = (double)hpoints[1].y - (double)hpoints[0].y;
        box.w = (float)sqrt(dx * dx + dy * dy);
        box.h = 0;
        box.angle = (float)atan2(dy, dx);
    } else if (n == 1) {
        box.cx = hpoints[0].x;
        box.cy = hpoints[0].y;
    }
    box.angle = (float)(box.angle * 180 / VSE_PI);
    return box;
}

// cv::boxPoints / RotatedRect::points
VSE_HD inline void cv_box_points(const RotRect& r, P2f* pt /*4*/) {
    double _angle = r.angle * VSE_PI / 180.;
    float b = (float)cos(_angle) * 0.5f;
    float a = (float)sin(_angle) * 0.5f;
    pt[0].x = r.cx - a * r.h - b * r.w;
    pt[0].y = r.cy + b * r.h - a * r.w;
    pt[1].x = r.cx + a * r.h - b * r.w;
    pt[1].y = r.cy - b * r.h - a * r.w;
    pt[2].x = 2 * r.cx - pt[0].x;
    pt[2].y = 2 * r.cy - pt[0].y;
    pt[3].x = 2 * r.cx - pt[1].x;
    pt[3].y = 2 * r.cy - pt[1].y;
}

// DBPostProcess.get_mini_boxes ordering: sort the 4 corners by x (stable), then pick [tl, tr, br, bl]
VSE_HD inline void mini_box_order(const P2f* in /*4*/, P2f* box /*4*/) {
    P2f p[4] = {in[0], in[1], in[2], in[3]};
    for (int i = 1; i < 4; i++) {  // stable insertion sort by x (Python sorted(key=x))
        P2f v = p[i];
        int j = i - 1;
        while (j >= 0 && p[j].x > v.x) {
            p[j + 1] = p[j];
            j--;
        }
        p[j + 1] = v;
    }
    int i1, i2, i3, i4;
    if (p[1].y > p[0].y) { i1 = 0; i4 = 1; } else { i1 = 1; i4 = 0; }
    if (p[3].y > p[2].y) { i2 = 2; i3 = 3; } else { i2 = 3; i3 = 2; }
    box[0] = p[i1];
    box[1] = p[i2];
    box[2] = p[i3];
    box[3] = p[i4];
}

// shapely Polygon(box).area / .length on float64 copies of the 4 float corners
VSE_HD inline void polygon_area_length(const P2f* b, int n, double* area, double* length) {
    double a = 0, l = 0;
    for (int i = 0; i < n; i++) {
        int j = (i + 1) % n;
        a += (double)b[i].x * (double)b[j].y - (double)b[j].x * (double)b[i].y;
        double dx = (double)b[j].x - (double)b[i].x, dy = (double)b[j].y - (double)b[i].y;
        l += sqrt(dx * dx + dy * dy);
    }
    *area = fabs(a * 0.5);
    *length = l;
}

// ------------------------------------------------------------------------------------------------------------------
// ClipperLib 6.4.2 ClipperOffset (JT_ROUND, ET_CLOSEDPOLYGON, ArcTolerance 0.25) on one polygon whose float corners
// pyclipper truncates to integers.  Returns the number of integer vertices written (0 = degenerate, -1 = overflow).
// ------------------------------------------------------------------------------------------------------------------
VSE_HD inline long long clipper_round(double v) { return (v < 0) ? (long long)(v - 0.5) : (long long)(v + 0.5); }

VSE_HD inline int clipper_offset_round(const P2f* path, int npath, double delta, P2i* dest, int cap) {
    P2i contour[8];
    int n = 0;
    if (npath > 8) return -1;
    {
        P2i pts[8];
        for (int i = 0; i < npath; i++) { pts[i].x = (int)path[i].x; pts[i].y = (int)path[i].y; }
        int high = npath - 1;
        while (high > 0 && pts[0].x == pts[high].x && pts[0].y == pts[high].y) high--;
        contour[n++] = pts[0];
        for (int i = 1; i <= high; i++)
            if (contour[n - 1].x != pts[i].x || contour[n - 1].y != pts[i].y) contour[n++] = pts[i];
    }
    if (n < 3) return 0;
    {
        double area = 0.0;
        int j = n - 1;
        for (int i = 0; i < n; i++) {
            area += ((double)contour[j].x + contour[i].x) * ((double)contour[j].y - contour[i].y);
            j = i;
        }
        area = -area * 0.5;
        if (area < 0)
            for (int i = 0, k = n - 1; i < k; i++, k--) { P2i t = contour[i]; contour[i] = contour[k]; contour[k] = t; }
    }
    int m = 0;
    if (fabs(delta) < 1e-20) {
        for (int i = 0; i < n && m < cap; i++) dest[m++] = contour[i];
        return m;
    }
    double y = 0.25;
    if (y > fabs(delta) * 0.25) y = fabs(delta) * 0.25;
    double steps = VSE_PI / acos(1 - y / fabs(delta));
    if (steps > fabs(delta) * VSE_PI) steps = fabs(delta) * VSE_PI;
    double m_sin = sin(2 * VSE_PI / steps);
    double m_cos = cos(2 * VSE_PI / steps);
    double steps_per_rad = steps / (2 * VSE_PI);
    if (delta < 0) m_sin = -m_sin;
    double nx[8], ny[8];
    for (int i = 0; i < n; i++) {
        P2i p1 = contour[i], p2 = contour[(i + 1) % n];
        if (p1.x == p2.x && p1.y == p2.y) { nx[i] = ny[i] = 0; continue; }
        double dx = (double)(p2.x - p1.x), dy = (double)(p2.y - p1.y);
        double f = 1.0 / sqrt(dx * dx + dy * dy);
        nx[i] = dy * f;
        ny[i] = -dx * f;
    }
    auto put = [&](long long X, long long Y) {
        if (m < cap) { dest[m].x = (int)X; dest[m].y = (int)Y; }
        m++;
    };
    int k = n - 1;
    for (int j = 0; j < n; j++) {
        double sx = contour[j].x, sy = contour[j].y;
        double sin_a = nx[k] * ny[j] - nx[j] * ny[k];
        bool done = false;
        if (fabs(sin_a * delta) < 1.0) {
            double cos_a = nx[k] * nx[j] + ny[j] * ny[k];
            if (cos_a > 0) {
                put(clipper_round(sx + nx[k] * delta), clipper_round(sy + ny[k] * delta));
                done = true;
            }
        } else if (sin_a > 1.0) sin_a = 1.0;
        else if (sin_a < -1.0) sin_a = -1.0;
        if (!done) {
            if (sin_a * delta < 0) {
                put(clipper_round(sx + nx[k] * delta), clipper_round(sy + ny[k] * delta));
                put((long long)sx, (long long)sy);
                put(clipper_round(sx + nx[j] * delta), clipper_round(sy + ny[j] * delta));
            } else {
                double a = atan2(sin_a, nx[k] * nx[j] + ny[k] * ny[j]);
                long long nsteps = clipper_round(steps_per_rad * fabs(a));
                if (nsteps < 1) nsteps = 1;
                double X = nx[k], Y = ny[k];
                for (long long s = 0; s < nsteps; s++) {
                    put(clipper_round(sx + X * delta), clipper_round(sy + Y * delta));
                    double X2 = X;
                    X = X * m_cos - m_sin * Y;
                    Y = X2 * m_sin + Y * m_cos;
                }
                put(clipper_round(sx + nx[j] * delta), clipper_round(sy + ny[j] * delta));
            }
        }
        k = j;
    }
    return m <= cap ? m : -1;
}

// ------------------------------------------------------------------------------------------------------------------
// TextDetector.order_points_clockwise + clip (filter_tag_det_res); returns false when the box is dropped.
// `pts` are the integer corners produced by DBPostProcess; `img_w/h` the frame size.
// ------------------------------------------------------------------------------------------------------------------
VSE_HD inline bool order_clip_filter(const P2i* pts /*4*/, int img_h, int img_w, P2f* out /*4*/) {
    int s[4], amin = 0, amax = 0;
    for (int i = 0; i < 4; i++) s[i] = pts[i].x + pts[i].y;
    for (int i = 1; i < 4; i++) {
        if (s[i] < s[amin]) amin = i;   // np.argmin: first minimum
        if (s[i] > s[amax]) amax = i;   // np.argmax: first maximum
    }
    P2i tmp[4];
    int nt = 0;
    for (int i = 0; i < 4; i++)
        if (i != amin && i != amax) tmp[nt++] = pts[i];
    // amin == amax happens only when all sums are equal: np.delete then removes a single row
    if (nt == 3) {
        // numpy would fail on this degenerate input upstream as well (diff/argmin still work on 3 rows);
        // keep the first two rows after the deleted one, matching np.delete((i, i))
        nt = 3;
    }
    int d0 = tmp[0].y - tmp[0].x, d1 = tmp[1].y - tmp[1].x;
    int i1 = 0, i3 = 0;
    if (nt == 2) {
        i1 = (d1 < d0) ? 1 : 0;
        i3 = (d1 > d0) ? 1 : 0;
    } else {
        int d2 = tmp[2].y - tmp[2].x;
        int d[3] = {d0, d1, d2};
        for (int i = 1; i < 3; i++) {
            if (d[i] < d[i1]) i1 = i;
            if (d[i] > d[i3]) i3 = i;
        }
    }
    P2i r[4] = {pts[amin], tmp[i1], pts[amax], tmp[i3]};
    for (int i = 0; i < 4; i++) {
        int x = r[i].x, y = r[i].y;
        x = x < 0 ? 0 : (x > img_w - 1 ? img_w - 1 : x);
        y = y < 0 ? 0 : (y > img_h - 1 ? img_h - 1 : y);
        out[i].x = (float)x;
        out[i].y = (float)y;
    }
    // np.linalg.norm on float32 rows -> float32
    float dx = out[0].x - out[1].x, dy = out[0].y - out[1].y;
    int rect_width = (int)sqrtf(dx * dx + dy * dy);
    dx = out[0].x - out[3].x; dy = out[0].y - out[3].y;
    int rect_height = (int)sqrtf(dx * dx + dy * dy);
    return !(rect_width <= 3 || rect_height <= 3);
}

// crop size of get_rotate_crop_image (float32 norms, int() truncation)
VSE_HD inline void crop_size(const P2f* q /*4*/, int* cw, int* ch) {
    auto norm = [](P2f a, P2f b) { float dx = a.x - b.x, dy = a.y - b.y; return sqrtf(dx * dx + dy * dy); };
    float w = fmaxf(norm(q[0], q[1]), norm(q[2], q[3]));
    float h = fmaxf(norm(q[0], q[3]), norm(q[1], q[2]));
    *cw = (int)w;
    *ch = (int)h;
}

// ------------------------------------------------------------------------------------------------------------------
// Inverse perspective map of get_rotate_crop_image: 3x3 double matrix taking crop pixel (x, y) to frame coordinates,
// i.e. inv(cv::getPerspectiveTransform(quad, [[0,0],[cw,0],[cw,ch],[0,ch]])), computed in closed form as the
// homography rect -> quad (Heckbert).  Differences to OpenCV's SVD solve + LU inverse are ~1e-13 relative and only
// matter for coordinates that fall exactly on a 1/32-pixel rounding tie.
// ------------------------------------------------------------------------------------------------------------------
VSE_HD inline void rect_to_quad_homography(const P2f* q /*4: tl,tr,br,bl*/, int cw, int ch, double* M /*9*/) {
    double x0 = q[0].x, y0 = q[0].y, x1 = q[1].x, y1 = q[1].y, x2 = q[2].x, y2 = q[2].y, x3 = q[3].x, y3 = q[3].y;
    double dx1 = x1 - x2, dx2 = x3 - x2, dx3 = x0 - x1 + x2 - x3;
    double dy1 = y1 - y2, dy2 = y3 - y2, dy3 = y0 - y1 + y2 - y3;
    double a, b, c, d, e, f, g, h;
    if (dx3 == 0.0 && dy3 == 0.0) {
        a = x1 - x0; b = x3 - x0; c = x0;
        d = y1 - y0; e = y3 - y0; f = y0;
        g = 0; h = 0;
    } else {
        double det = dx1 * dy2 - dx2 * dy1;
        g = (dx3 * dy2 - dx2 * dy3) / det;
        h = (dx1 * dy3 - dx3 * dy1) / det;
        a = x1 - x0 + g * x1; b = x3 - x0 + h * x3; c = x0;
        d = y1 - y0 + g * y1; e = y3 - y0 + h * y3; f = y0;
    }
    // unit square -> quad; pre-scale so that (cw, ch) maps like (1, 1)
    double sx = cw > 0 ? 1.0 / cw : 0.0, sy = ch > 0 ? 1.0 / ch : 0.0;
    M[0] = a * sx; M[1] = b * sy; M[2] = c;
    M[3] = d * sx; M[4] = e * sy; M[5] = f;
    M[6] = g * sx; M[7] = h * sy; M[8] = 1.0;
}

}  // namespace geom
}  // namespace vse
