// geom_host.cpp — TEST HARNESS: compiles the product's host/device geometry (csrc/geom.cuh, csrc/dbpost_core.cuh)
// for the CPU so that tests/test_geom_cpu.py can check the very same source against cv2 without a GPU.
// Not part of libvse_b200.so; built by __graft_entry__.build() into tests/native/_build/libgeom_host.so.
#include <cstring>
#include <vector>

#include "../../video_subtitle_extractor_b200/csrc/geom.cuh"
#include "../../video_subtitle_extractor_b200/csrc/dbpost_core.cuh"

using namespace vse::geom;

extern "C" {

int gh_convex_hull(const int* pts_xy, int n, int* hull_idx) {
    std::vector<int> order(n), stack(n + 2);
    return cv_convex_hull(reinterpret_cast<const P2i*>(pts_xy), n, order.data(), stack.data(), hull_idx);
}

// hull of a component from its row extents, in cv::convexHull order; returns count, points in out_xy
int gh_hull_rows(const int* xl, const int* xr, int rows, int y0, int* out_xy) {
    std::vector<P2i> out(2 * rows + 4);
    int m = hull_from_row_extents(xl, xr, rows, y0, out.data());
    std::memcpy(out_xy, out.data(), sizeof(P2i) * m);
    return m;
}

// minAreaRect + boxPoints + get_mini_boxes on an ordered hull (int points). box8 = 4 corners, returns sside
float gh_mini_box_from_hull(const int* hull_xy, int m, float* box8, float* rect5) {
    std::vector<P2f> hp(m);
    for (int i = 0; i < m; i++) { hp[i].x = (float)hull_xy[2 * i]; hp[i].y = (float)hull_xy[2 * i + 1]; }
    std::vector<float> work(3 * m + 8);
    RotRect r = cv_min_area_rect(hp.data(), m, work.data());
    if (rect5) { rect5[0] = r.cx; rect5[1] = r.cy; rect5[2] = r.w; rect5[3] = r.h; rect5[4] = r.angle; }
    P2f bp[4], box[4];
    cv_box_points(r, bp);
    mini_box_order(bp, box);
    std::memcpy(box8, box, sizeof(box));
    return r.w < r.h ? r.w : r.h;
}

int gh_unclip(const float* box8, double delta, int* out_xy, int cap) {
    return clipper_offset_round(reinterpret_cast<const P2f*>(box8), 4, delta, reinterpret_cast<P2i*>(out_xy), cap);
}

// stage 1 of the candidate pipeline (component -> mini box); returns 1 if it survives the min-size test
int gh_db_stage1(const int* xl, const int* xr, int rows, int y0, float* box8) {
    std::vector<P2i> hull(2 * rows + 4);
    std::vector<float> work(3 * (2 * rows + 4) + 8);
    vse::dbpost::Candidate c;
    bool ok = vse::dbpost::stage1_component_box(xl, xr, rows, y0, hull.data(), work.data(), 3.0f, &c);
    std::memcpy(box8, c.box, sizeof(c.box));
    return ok ? 1 : 0;
}

// score mask rows: for the window-relative integer quad and the (w x h) window, row y -> [xa, xb] clipped to the window
// (returns 0 if the row is empty)
int gh_quad_row_span(const int* qx, const int* qy, int w, int h, int y, int* xa, int* xb) {
    return vse::dbpost::quad_row_span(qx, qy, w, h, y, xa, xb) ? 1 : 0;
}

void gh_score_window(const float* box8, int rw, int rh, int* win /*xmin,ymin,xmax,ymax*/, int* qx, int* qy) {
    vse::dbpost::score_window(reinterpret_cast<const P2f*>(box8), rw, rh, win, qx, qy);
}

// stage 3: mini box -> unclip -> second box -> frame-space integer corners -> ordered float quad
int gh_db_stage3(const float* box8, float unclip_ratio, int rw, int rh, int src_w, int src_h, float* quad8, int* ipts8) {
    vse::dbpost::Candidate c;
    std::memcpy(c.box, box8, sizeof(c.box));
    bool ok = vse::dbpost::stage3_unclip_scale(&c, unclip_ratio, 3.0f, rw, rh, src_w, src_h);
    std::memcpy(quad8, c.quad, sizeof(c.quad));
    if (ipts8) std::memcpy(ipts8, c.ipts, sizeof(c.ipts));
    return ok ? 1 : 0;
}

void gh_crop_size(const float* quad8, int* cw, int* ch) { crop_size(reinterpret_cast<const P2f*>(quad8), cw, ch); }

void gh_homography(const float* quad8, int cw, int ch, double* M9) {
    rect_to_quad_homography(reinterpret_cast<const P2f*>(quad8), cw, ch, M9);
}

// bicubic perspective sample of one crop pixel (same routine the crop kernel calls)
void gh_warp_crop(const unsigned char* frame, int fh, int fw, int stride, const float* quad8, unsigned char* out, int cw, int ch) {
    double M[9];
    rect_to_quad_homography(reinterpret_cast<const P2f*>(quad8), cw, ch, M);
    static short tab[32 * 32 * 16];
    static bool init = false;
    if (!init) { vse::dbpost::build_cubic_table(tab); init = true; }
    for (int y = 0; y < ch; y++)
        for (int x = 0; x < cw; x++) {
            unsigned char px[3];
            vse::dbpost::warp_cubic_pixel(frame, fh, fw, stride, 3, M, tab, x, y, px);
            out[(y * cw + x) * 3 + 0] = px[0];
            out[(y * cw + x) * 3 + 1] = px[1];
            out[(y * cw + x) * 3 + 2] = px[2];
        }
}

}  // extern "C"

extern "C" void gh_calipers_raw(const int* hull_xy, int m, float* out6) {
    std::vector<P2f> hp(m);
    for (int i = 0; i < m; i++) { hp[i].x = (float)hull_xy[2 * i]; hp[i].y = (float)hull_xy[2 * i + 1]; }
    std::vector<float> work(3 * m + 8);
    rotating_calipers_minarea(hp.data(), m, work.data(), out6);
}
