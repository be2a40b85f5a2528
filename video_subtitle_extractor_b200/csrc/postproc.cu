// postproc.cu — device-side DB post-process, text-line crops and CTC greedy decode (sm_100a).
//
// Replaces, on the GPU, what paddleocr 2.10 does on the host between and after the two predictor runs
// (reference call sites: backend/tools/ocr.py:27, backend/tools/subtitle_detect.py:24-26):
//   cv2.findContours(pred > 0.3)  ->  union-find connected-component labelling (8-connectivity) + per-row extents
//   get_mini_boxes / box_score_fast / unclip / filter_tag_det_res / sorted_boxes  ->  one block per component
//   get_rotate_crop_image  ->  bicubic perspective gather;   CTCLabelDecode  ->  one block per text line.
// The per-candidate arithmetic lives in dbpost_core.cuh / geom.cuh and is unit-tested on the CPU against cv2.
// Compiled with --fmad=false: OpenCV's float code is not FMA-contracted and box corners must match bit for bit.
//
// cv2.findContours(RETR_LIST) also reports HOLE borders (the foreground pixels 4-adjacent to an enclosed background
// region); upstream treats them like any contour, and on real video some survive the score / size filters (frame 1391 of
// the reference's test_cn.mp4 yields two such boxes).  They are found here by labelling the background of the blocks that
// hold foreground (4-connectivity) and keeping the components that never reach the image border or a block without
// foreground; such a component that swallowed a whole empty block could only score far below box_thresh and is dropped.
#include "postproc.cuh"
#include "pdl.cuh"

#include <cfloat>

#include "dbpost_core.cuh"

namespace vse {

// ------------------------------------------------------------------------------------------------
// connected components
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int uf_find(volatile int* L, int a) {
    int p = L[a];
    while (p != a) {
        a = p;
        p = L[a];
    }
    return a;
}

__device__ __forceinline__ void uf_union(int* L, int a, int b) {
    bool done;
    do {
        a = uf_find(L, a);
        b = uf_find(L, b);
        if (a < b) {
            int old = atomicMin(&L[b], a);
            done = (old == b);
            b = old;
        } else if (b < a) {
            int old = atomicMin(&L[a], b);
            done = (old == a);
            a = old;
        } else {
            done = true;
        }
    } while (!done);
}

// block (32, 8): a warp covers 32 consecutive pixels of one row; the initial label is the start of the pixel's
// run inside that 32-pixel segment (one ballot, no chains along rows)
__global__ void __launch_bounds__(256) db_label_init(const float* __restrict__ prob, const DetFrame* __restrict__ frames,
                                                     float thresh, int* __restrict__ labels, int* __restrict__ status,
                                                     int* __restrict__ fg_count, int* __restrict__ fg_list,
                                                     int* __restrict__ blabels, int* __restrict__ bopen, int* __restrict__ blk_listed) {
    pdl_wait();      // pdl.cuh: nothing below may run before the previous kernel of the stream has completed
    pdl_trigger();
    const DetFrame fr = frames[blockIdx.z];
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    const bool in = x < fr.rw && y < fr.rh;
    int g = 0;
    bool fg = false;
    if (in) {
        g = fr.map_off + y * fr.rw + x;
        const float pv = prob[g];
        fg = pv > thresh;
        // a probability is a sigmoid output: anything non-finite means the network overflowed its activation type
        if (!(fabsf(pv) <= 1.5f)) atomicOr(&status[blockIdx.z], 4);
    }
    const unsigned mask = __ballot_sync(0xffffffffu, fg);
    // subtitle frames are > 99 % background: the later labelling passes skip blocks without foreground
    const int any = __syncthreads_or(fg ? 1 : 0);
    if (any && threadIdx.x == 0 && threadIdx.y == 0) {
        const int slot = atomicAdd(fg_count, 1);      // order of the list does not matter: the passes below are order-free
        fg_list[slot] = (blockIdx.z << 20) | (blockIdx.y << 10) | blockIdx.x;
        blk_listed[(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = 1;
    }
    if (!in) return;
    const int lane = threadIdx.x;
    if (fg) {
        const unsigned below = ~mask & ((1u << lane) - 1u);
        const int start = below ? (32 - __clz(below)) : 0;
        labels[g] = g - (lane - start);
    } else {
        labels[g] = -1;
    }
    if (any) {   // background labels of a listed block: start of the pixel's background run inside the 32-pixel segment
        if (fg) {
            blabels[g] = -1;
        } else {
            const unsigned below = mask & ((1u << lane) - 1u);
            const int start = below ? (32 - __clz(below)) : 0;
            blabels[g] = g - (lane - start);
        }
        bopen[g] = 0;
    }
}

__device__ __forceinline__ bool blk_is_listed(const int* __restrict__ blk_listed, int f, int gx, int gy, int x, int y) {
    return blk_listed[(f * gy + (y >> 3)) * gx + (x >> 5)] != 0;
}

__global__ void __launch_bounds__(256) db_label_merge(const DetFrame* __restrict__ frames, int* labels,
                                                      const int* __restrict__ fg_count, const int* __restrict__ fg_list,
                                                      int* blabels, const int* __restrict__ blk_listed, int gx, int gy) {
    pdl_wait();      // pdl.cuh: nothing below may run before the previous kernel of the stream has completed
    pdl_trigger();
  const int n_fg = *fg_count;
  for (int it = blockIdx.x; it < n_fg; it += gridDim.x) {
    const int code = fg_list[it], bz = code >> 20, by = (code >> 10) & 1023, bx = code & 1023;
    const DetFrame fr = frames[bz];
    const int x = bx * 32 + threadIdx.x, y = by * blockDim.y + threadIdx.y;
    if (x >= fr.rw || y >= fr.rh) continue;
    const int g = fr.map_off + y * fr.rw + x;
    if (labels[g] < 0) {
        // background, 4-connectivity: the run inside the segment is already one label; join across the segment boundary and
        // with the pixel above — where that neighbour lies in a listed block (otherwise the component is open: db_label_flatten)
        if (threadIdx.x == 0 && x > 0 && blk_is_listed(blk_listed, bz, gx, gy, x - 1, y) && blabels[g - 1] >= 0) uf_union(blabels, g, g - 1);
        if (y > 0 && (threadIdx.y != 0 || blk_is_listed(blk_listed, bz, gx, gy, x, y - 1)) && blabels[g - fr.rw] >= 0)
            uf_union(blabels, g, g - fr.rw);
        continue;
    }
    const int rw = fr.rw;
    const bool W = x > 0 && labels[g - 1] >= 0;
    const bool up = y > 0;
    const bool N = up && labels[g - rw] >= 0;
    const bool NE = up && x + 1 < rw && labels[g - rw + 1] >= 0;
    if (W) {
        if (threadIdx.x == 0) uf_union(labels, g, g - 1);  // run continues across the 32-pixel segment boundary
        if (!N && NE) uf_union(labels, g, g - rw + 1);
    } else {
        if (N) {
            uf_union(labels, g, g - rw);
        } else {
            const bool NW = up && x > 0 && labels[g - rw - 1] >= 0;
            if (NW) uf_union(labels, g, g - rw - 1);
            if (NE) uf_union(labels, g, g - rw + 1);
        }
    }
  }
}

__global__ void __launch_bounds__(256) db_label_flatten(const DetFrame* __restrict__ frames, int* labels, int* slot_of,
                                                        int* n_comp, int* roots, int* bbox,
                                                        const int* __restrict__ fg_count, const int* __restrict__ fg_list,
                                                        int* blabels, int* bopen, const int* __restrict__ blk_listed, int gx, int gy) {
    pdl_wait();      // pdl.cuh: nothing below may run before the previous kernel of the stream has completed
    pdl_trigger();
  const int n_fg = *fg_count;
  for (int it = blockIdx.x; it < n_fg; it += gridDim.x) {
    const int code = fg_list[it], f = code >> 20, by = (code >> 10) & 1023, bx = code & 1023;
    const DetFrame fr = frames[f];
    const int x = bx * 32 + threadIdx.x, y = by * blockDim.y + threadIdx.y;
    if (x >= fr.rw || y >= fr.rh) continue;
    const int g = fr.map_off + y * fr.rw + x;
    if (labels[g] < 0) {
        const int r = uf_find(blabels, g);
        blabels[g] = r;
        // reaches the image border (cv2 frames the image with background) or a block without foreground: not a hole
        const bool open = x == 0 || y == 0 || x == fr.rw - 1 || y == fr.rh - 1 ||
                          !blk_is_listed(blk_listed, f, gx, gy, x - 1, y) || !blk_is_listed(blk_listed, f, gx, gy, x + 1, y) ||
                          !blk_is_listed(blk_listed, f, gx, gy, x, y - 1) || !blk_is_listed(blk_listed, f, gx, gy, x, y + 1);
        if (open) bopen[r] = 1;
        continue;
    }
    const int r = uf_find(labels, g);
    labels[g] = r;
    if (r == g) {
        const int slot = atomicAdd(&n_comp[f], 1);
        if (slot < kSlotCap) {
            roots[f * kSlotCap + slot] = g;
            slot_of[g] = slot;
            int* b = bbox + (size_t(f) * kSlotCap + slot) * 4;
            b[0] = x; b[1] = y; b[2] = x; b[3] = y;
        } else {
            slot_of[g] = -1;
        }
    }
  }
}

// closed background components become contours of their own (hole borders)
__global__ void __launch_bounds__(256) db_hole_slots(const DetFrame* __restrict__ frames, const int* __restrict__ labels,
                                                     const int* __restrict__ blabels, const int* __restrict__ bopen, int* slot_of,
                                                     int* n_comp, int* roots, int* bbox, const int* __restrict__ fg_count,
                                                     const int* __restrict__ fg_list) {
    pdl_wait();
    pdl_trigger();
  const int n_fg = *fg_count;
  for (int it = blockIdx.x; it < n_fg; it += gridDim.x) {
    const int code = fg_list[it], f = code >> 20, by = (code >> 10) & 1023, bx = code & 1023;
    const DetFrame fr = frames[f];
    const int x = bx * 32 + threadIdx.x, y = by * blockDim.y + threadIdx.y;
    if (x >= fr.rw || y >= fr.rh) continue;
    const int g = fr.map_off + y * fr.rw + x;
    if (labels[g] >= 0 || blabels[g] != g) continue;
    if (bopen[g]) { slot_of[g] = -1; continue; }
    const int slot = atomicAdd(&n_comp[f], 1);
    if (slot < kSlotCap) {
        roots[f * kSlotCap + slot] = g | kHoleFlag;
        slot_of[g] = slot;
        int* b = bbox + (size_t(f) * kSlotCap + slot) * 4;
        b[0] = x; b[1] = y; b[2] = x; b[3] = y;
    } else {
        slot_of[g] = -1;
    }
  }
}

__global__ void __launch_bounds__(256) db_bbox(const DetFrame* __restrict__ frames, const int* __restrict__ labels,
                                               const int* __restrict__ slot_of, int* bbox,
                                               const int* __restrict__ fg_count, const int* __restrict__ fg_list,
                                               const int* __restrict__ blabels) {
    pdl_wait();      // pdl.cuh: nothing below may run before the previous kernel of the stream has completed
    pdl_trigger();
  const int n_fg = *fg_count;
  for (int it = blockIdx.x; it < n_fg; it += gridDim.x) {
    const int code = fg_list[it], f = code >> 20, by = (code >> 10) & 1023, bx = code & 1023;
    const DetFrame fr = frames[f];
    const int x = bx * 32 + threadIdx.x, y = by * blockDim.y + threadIdx.y;
    if (x >= fr.rw || y >= fr.rh) continue;
    const int g = fr.map_off + y * fr.rw + x;
    const int r = labels[g];
    if (r < 0) {
        const int slot = slot_of[blabels[g]];      // hole pixels are few: every one updates the box of its hole
        if (slot >= 0) {
            int* b = bbox + (size_t(f) * kSlotCap + slot) * 4;
            atomicMin(b + 0, x); atomicMin(b + 1, y); atomicMax(b + 2, x); atomicMax(b + 3, y);
        }
        continue;
    }
    const bool W = x > 0 && labels[g - 1] >= 0;
    const bool E = x + 1 < fr.rw && labels[g + 1] >= 0;
    if (W && E) continue;
    const int slot = slot_of[r];
    if (slot < 0) continue;
    int* b = bbox + (size_t(f) * kSlotCap + slot) * 4;
    if (!W) {
        atomicMin(b + 0, x);
        atomicMin(b + 1, y);
        atomicMax(b + 3, y);
    }
    if (!E) atomicMax(b + 2, x);
  }
}

// cv2.findContours(RETR_LIST) returns contours in reverse discovery order.  The raster scan discovers an outer border at the
// component's first pixel and a hole border at the pixel LEFT of the hole's first pixel (outer border first when both start
// at the same pixel): key = 2 * start pixel (+ 1 for a hole), descending.
__global__ void __launch_bounds__(256) db_sort_components(const int* __restrict__ n_comp, const int* __restrict__ roots,
                                                          int* order, int* status, int* ckey) {
    pdl_wait();      // pdl.cuh: nothing below may run before the previous kernel of the stream has completed
    pdl_trigger();
    const int f = blockIdx.x;
    int n = n_comp[f];
    if (n > kSlotCap) {
        if (threadIdx.x == 0) atomicOr(&status[f], 1);
        n = kSlotCap;
    }
    const int* r = roots + size_t(f) * kSlotCap;
    int* key = ckey + size_t(f) * kSlotCap;
    for (int t = threadIdx.x; t < n; t += blockDim.x) {
        const int v = r[t];
        key[t] = (v & kHoleFlag) ? 2 * ((v & ~kHoleFlag) - 1) + 1 : 2 * v;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < n; t += blockDim.x) {
        const int mine = key[t];
        int rank = 0;
        for (int j = 0; j < n; j++) rank += key[j] > mine;
        order[size_t(f) * kSlotCap + rank] = t;
    }
}

// ------------------------------------------------------------------------------------------------
// one block per component: row extents -> box -> score -> unclip -> frame-space quad
// ------------------------------------------------------------------------------------------------
static constexpr int kCandThreads = 128;
static constexpr int kCandFields = 10;

size_t db_candidate_smem_bytes(int max_rh) {
    size_t rows = size_t(max_rh);
    return rows * 2 * sizeof(int) + (2 * rows + 4) * sizeof(geom::P2i) + 5 * (2 * rows + 4) * sizeof(float) + 64;
}

__global__ void __launch_bounds__(kCandThreads) db_candidates(const float* __restrict__ prob, const DetFrame* __restrict__ frames,
                                                              const int* __restrict__ labels, const int* __restrict__ n_comp,
                                                              const int* __restrict__ roots, const int* __restrict__ bbox,
                                                              const int* __restrict__ order, DbParams p, int max_rh, float* cand,
                                                              const int* __restrict__ blabels, const int* __restrict__ blk_listed,
                                                              int gx, int gy) {
    pdl_wait();      // pdl.cuh: nothing below may run before the previous kernel of the stream has completed
    pdl_trigger();
    extern __shared__ __align__(16) unsigned char smem[];
    int* xl = reinterpret_cast<int*>(smem);
    int* xr = xl + max_rh;
    geom::P2i* hull = reinterpret_cast<geom::P2i*>(xr + max_rh);
    float* work = reinterpret_cast<float*>(hull + 2 * max_rh + 4);
    __shared__ dbpost::Candidate c;
    __shared__ int win[4], qx[4], qy[4], ok;
    __shared__ double wsum[kCandThreads / 32];
    __shared__ int wcnt[kCandThreads / 32];

    const int f = blockIdx.y;
    const DetFrame fr = frames[f];
    int n = min(n_comp[f], kSlotCap);
    n = min(n, p.max_candidates);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int k = blockIdx.x; k < n; k += gridDim.x) {
        const int slot = order[size_t(f) * kSlotCap + k];
        const int root_raw = roots[size_t(f) * kSlotCap + slot];
        const bool hole = (root_raw & kHoleFlag) != 0;
        const int root = root_raw & ~kHoleFlag;
        const int* b = bbox + (size_t(f) * kSlotCap + slot) * 4;
        int xmin = b[0], ymin = b[1], xmax = b[2], ymax = b[3];
        if (hole) {
            // hole border = the foreground pixels 4-adjacent to the hole's pixels: one row / column around the hole's box (a
            // closed background component never touches the image border, so the ring is inside the map)
            const int hx0 = xmin, hy0 = ymin, hx1 = xmax, hy1 = ymax;
            ymin -= 1; ymax += 1;
            const int rows_h = ymax - ymin + 1;
            for (int r = threadIdx.x; r < rows_h; r += kCandThreads) { xl[r] = INT_MAX; xr[r] = -1; }
            __syncthreads();
            const int hw = hx1 - hx0 + 1, hn = hw * (hy1 - hy0 + 1);
            for (int i = threadIdx.x; i < hn; i += kCandThreads) {
                const int y = hy0 + i / hw, x = hx0 + i % hw;
                if (!blk_is_listed(blk_listed, f, gx, gy, x, y)) continue;      // labels of unlisted blocks are stale
                const int g = fr.map_off + y * fr.rw + x;
                if (blabels[g] != root) continue;
                const int r = y - ymin;
                if (labels[g - 1] >= 0) { atomicMin(&xl[r], x - 1); atomicMax(&xr[r], x - 1); }
                if (labels[g + 1] >= 0) { atomicMin(&xl[r], x + 1); atomicMax(&xr[r], x + 1); }
                if (labels[g - fr.rw] >= 0) { atomicMin(&xl[r - 1], x); atomicMax(&xr[r - 1], x); }
                if (labels[g + fr.rw] >= 0) { atomicMin(&xl[r + 1], x); atomicMax(&xr[r + 1], x); }
            }
        }
        const int rows = ymax - ymin + 1;
        // per-row extents of the component inside its bounding box
        for (int r = warp; r < rows && !hole; r += kCandThreads / 32) {
            const int* L = labels + fr.map_off + (ymin + r) * fr.rw;
            int lo = INT_MAX, hi = -1;
            for (int x0 = xmin; x0 <= xmax; x0 += 32) {
                const int x = x0 + lane;
                const bool hit = x <= xmax && L[x] == root;
                const unsigned m = __ballot_sync(0xffffffffu, hit);
                if (m) {
                    lo = min(lo, x0 + __ffs(m) - 1);
                    hi = max(hi, x0 + 31 - __clz(m));
                }
            }
            if (lane == 0) { xl[r] = lo; xr[r] = hi; }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            ok = dbpost::stage1_component_box(xl, xr, rows, ymin, hull, work, 3.0f, &c) ? 1 : 0;
            if (ok) dbpost::score_window(reinterpret_cast<const geom::P2f*>(c.box), fr.rw, fr.rh, win, qx, qy);
        }
        __syncthreads();
        if (ok) {
            // box_score_fast: mean of pred under fillPoly(quad) inside the clipped window (double accumulation)
            const int ww = win[2] - win[0] + 1, wh = win[3] - win[1] + 1;
            double s = 0.0;
            int cnt = 0;
            for (int y = warp; y < wh; y += kCandThreads / 32) {
                int xa, xb;
                if (!dbpost::quad_row_span(qx, qy, ww, wh, y, &xa, &xb)) continue;
                const float* row = prob + fr.map_off + (win[1] + y) * fr.rw + win[0];
                float part = 0.f;
                for (int x = xa + lane; x <= xb; x += 32) part += row[x];
                s += (double)part;
                if (lane == 0) cnt += xb - xa + 1;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) { wsum[warp] = s; wcnt[warp] = cnt; }
            __syncthreads();
            if (threadIdx.x == 0) {
                double tot = 0.0;
                int tc = 0;
                for (int w = 0; w < kCandThreads / 32; w++) { tot += wsum[w]; tc += wcnt[w]; }
                c.score = tc > 0 ? (float)(tot / (double)tc) : 0.f;
                if (p.box_thresh > c.score) ok = 0;
                else ok = dbpost::stage3_unclip_scale(&c, p.unclip_ratio, 3.0f, fr.rw, fr.rh, fr.src_w, fr.src_h) ? 1 : 0;
            }
        }
        __syncthreads();
        if (threadIdx.x < kCandFields) {
            float* o = cand + (size_t(f) * p.max_candidates + k) * kCandFields;
            float v;
            if (threadIdx.x == 0) v = ok ? 1.f : 0.f;
            else if (threadIdx.x == 1) v = c.score;
            else v = c.quad[threadIdx.x - 2];
            o[threadIdx.x] = v;
        }
        __syncthreads();
    }
}

__global__ void db_clear_overflow(int* status, int n) {
    pdl_wait();
    pdl_trigger();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) status[i] &= ~2;
}

// keep valid candidates in contour order, optionally re-order like TextSystem.sorted_boxes
__global__ void db_compact(const int* __restrict__ n_comp, const float* __restrict__ cand, DbParams p, int* n_boxes,
                           float* quads, float* scores, int* status) {
    pdl_wait();      // pdl.cuh: nothing below may run before the previous kernel of the stream has completed
    pdl_trigger();
    const int f = blockIdx.x;
    if (threadIdx.x != 0) return;
    int n = min(min(n_comp[f], kSlotCap), p.max_candidates);
    float* q = quads + size_t(f) * p.max_boxes * 8;
    float* sc = scores + size_t(f) * p.max_boxes;
    int m = 0;
    for (int k = 0; k < n; k++) {
        const float* c = cand + (size_t(f) * p.max_candidates + k) * kCandFields;
        if (c[0] == 0.f) continue;
        if (m >= p.max_boxes) {
            atomicOr(&status[f], 2);
            break;
        }
        sc[m] = c[1];
        for (int i = 0; i < 8; i++) q[m * 8 + i] = c[2 + i];
        m++;
    }
    if (p.sort_reading_order && m > 1) {
        // sorted(dt_boxes, key=(y0, x0)) — stable
        for (int i = 1; i < m; i++) {
            float v[8], vs = sc[i];
            for (int t = 0; t < 8; t++) v[t] = q[i * 8 + t];
            int j = i - 1;
            while (j >= 0 && (q[j * 8 + 1] > v[1] || (q[j * 8 + 1] == v[1] && q[j * 8] > v[0]))) {
                for (int t = 0; t < 8; t++) q[(j + 1) * 8 + t] = q[j * 8 + t];
                sc[j + 1] = sc[j];
                j--;
            }
            for (int t = 0; t < 8; t++) q[(j + 1) * 8 + t] = v[t];
            sc[j + 1] = vs;
        }
        // adjacent lines whose tops differ by < 10 px are ordered left to right
        for (int i = 0; i < m - 1; i++) {
            for (int j = i; j >= 0; j--) {
                if (fabsf(q[(j + 1) * 8 + 1] - q[j * 8 + 1]) < 10.f && q[(j + 1) * 8] < q[j * 8]) {
                    for (int t = 0; t < 8; t++) { float tmp = q[j * 8 + t]; q[j * 8 + t] = q[(j + 1) * 8 + t]; q[(j + 1) * 8 + t] = tmp; }
                    float ts = sc[j]; sc[j] = sc[j + 1]; sc[j + 1] = ts;
                } else {
                    break;
                }
            }
        }
    }
    n_boxes[f] = m;
}

void launch_db_postprocess(const float* prob, const DetFrame* frames_dev, const DetFrame* frames_host, int n_frames,
                           int max_rh, int max_rw, const DbParams& p, const DbWorkspace& ws, cudaStream_t st,
                           int64_t* launches) {
    (void)frames_host;
    if (n_frames <= 0) return;
    cudaMemsetAsync(ws.n_comp, 0, sizeof(int) * n_frames, st);
    cudaMemsetAsync(ws.status, 0, sizeof(int) * n_frames, st);
    dim3 blk(32, 8), grid((max_rw + 31) / 32, (max_rh + 7) / 8, n_frames);
    // (32 x 8)-pixel blocks that hold foreground are listed by the first pass; the other passes walk that list with a
    // machine-sized grid (subtitle maps are > 99 % background, and launching 65 k empty blocks costs more than the work)
    if (grid.x > 1023 || grid.y > 1023 || n_frames > 2047) return;   // list code = frame << 20 | block_y << 10 | block_x (checked by the caller)
    cudaMemsetAsync(ws.fg_count, 0, sizeof(int), st);
    const int gx = int(grid.x), gy = int(grid.y);
    cudaMemsetAsync(ws.blk_listed, 0, sizeof(int) * size_t(n_frames) * gx * gy, st);
    pdl_launch(db_label_init, grid, blk, 0, st, prob, frames_dev, p.thresh, ws.labels, ws.status, ws.fg_count, ws.fg_list,
               ws.blabels, ws.bopen, ws.blk_listed);
    const int pgrid = 148 * 4;
    pdl_launch(db_label_merge, pgrid, blk, 0, st, frames_dev, ws.labels, ws.fg_count, ws.fg_list, ws.blabels, ws.blk_listed, gx, gy);
    pdl_launch(db_label_flatten, pgrid, blk, 0, st, frames_dev, ws.labels, ws.slot_of, ws.n_comp, ws.roots, ws.bbox, ws.fg_count,
               ws.fg_list, ws.blabels, ws.bopen, ws.blk_listed, gx, gy);
    pdl_launch(db_hole_slots, pgrid, blk, 0, st, frames_dev, ws.labels, ws.blabels, ws.bopen, ws.slot_of, ws.n_comp, ws.roots, ws.bbox,
               ws.fg_count, ws.fg_list);
    pdl_launch(db_bbox, pgrid, blk, 0, st, frames_dev, ws.labels, ws.slot_of, ws.bbox, ws.fg_count, ws.fg_list, ws.blabels);
    pdl_launch(db_sort_components, n_frames, 256, 0, st, ws.n_comp, ws.roots, ws.order, ws.status, ws.ckey);
    size_t smem = db_candidate_smem_bytes(max_rh);
    {   // the opt-in belongs to the current device's copy of the kernel: one high-water mark per device
        static size_t configured[64] = {};
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev < 0 || dev >= 64 || smem > configured[dev]) {
            if (cudaFuncSetAttribute(db_candidates, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return;
            if (dev >= 0 && dev < 64) configured[dev] = smem;
        }
    }
    pdl_launch(db_candidates, dim3(16, n_frames), kCandThreads, smem, st, prob, frames_dev, ws.labels, ws.n_comp, ws.roots, ws.bbox,
                                                                  ws.order, p, max_rh, ws.cand, ws.blabels, ws.blk_listed, gx, gy);
    pdl_launch(db_compact, n_frames, 32, 0, st, ws.n_comp, ws.cand, p, ws.n_boxes, ws.quads, ws.scores, ws.status);
    if (launches) *launches += 8;
}

// the compaction alone, with p.max_boxes rows per frame (the candidates of the last launch_db_postprocess are reused)
void launch_db_compact(int n_frames, const DbParams& p, const DbWorkspace& ws, cudaStream_t st, int64_t* launches) {
    pdl_launch(db_clear_overflow, (n_frames + 255) / 256, 256, 0, st, ws.status, n_frames);
    pdl_launch(db_compact, n_frames, 32, 0, st, ws.n_comp, ws.cand, p, ws.n_boxes, ws.quads, ws.scores, ws.status);
    if (launches) *launches += 2;
}

// ------------------------------------------------------------------------------------------------
// crops
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) crop_warp_kernel(const CropJob* __restrict__ jobs, const short* __restrict__ tab,
                                                        uint8_t* __restrict__ dst) {
    pdl_wait();      // pdl.cuh: nothing below may run before the previous kernel of the stream has completed
    pdl_trigger();
    const CropJob& j = jobs[blockIdx.y];
    const int ow = j.rot90 ? j.ch : j.cw, oh = j.rot90 ? j.cw : j.ch;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= ow * oh) return;
    const int oy = idx / ow, ox = idx - oy * ow;
    // np.rot90(a)[i][j] = a[j][W - 1 - i]
    const int x = j.rot90 ? (j.cw - 1 - oy) : ox;
    const int y = j.rot90 ? ox : oy;
    unsigned char px[3];
    dbpost::warp_cubic_pixel(j.frame, j.fh, j.fw, j.stride, j.pix, j.M, tab, x, y, px);
    reinterpret_cast<uchar4*>(dst)[j.dst_off + idx] = make_uchar4(px[0], px[1], px[2], 0);
}

void launch_crops(const CropJob* jobs_dev, int n_jobs, int max_pix, const short* cubic_tab, uint8_t* dst, cudaStream_t st) {
    if (n_jobs <= 0 || max_pix <= 0) return;
    dim3 grid((max_pix + 255) / 256, n_jobs);
    pdl_launch(crop_warp_kernel, grid, 256, 0, st, jobs_dev, cubic_tab, dst);
}

// ------------------------------------------------------------------------------------------------
// CTC greedy decode: argmax/max per time step, collapse repeats, drop blank (0), mean of kept max-probs.
// logits != 0: the rows are the class LOGITS of the recogniser head (row pitch cs), the plan's final softmax is not run at all —
// argmax is the same, and the probability of the winning class is 1 / sum_c exp(x_c - max) (one online pass:
// running maximum and rescaled sum per lane, merged over the warp), so the normalised [T, C] matrix never exists in HBM.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) ctc_decode_kernel(const float* __restrict__ probs, int C, int cs, int logits,
                                                         const int* __restrict__ toff, const int* __restrict__ tlen, int max_t,
                                                         int* __restrict__ ids, int* __restrict__ id_len, float* __restrict__ score) {
    pdl_wait();      // pdl.cuh: nothing below may run before the previous kernel of the stream has completed
    pdl_trigger();
    __shared__ int bad;
    if (threadIdx.x == 0) bad = 0;
    __syncthreads();
    extern __shared__ __align__(8) unsigned char sm[];
    int* bi = reinterpret_cast<int*>(sm);
    float* bp = reinterpret_cast<float*>(bi + max_t);
    const int n = blockIdx.x;
    const int T = tlen[n];
    const float* base = probs + size_t(toff[n]) * cs;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int t = warp; t < T; t += 4) {
        const float* row = base + size_t(t) * cs;
        float best = -FLT_MAX, sum = 0.f;       // sum: sum of exp(x - best) over this lane's classes (logits mode)
        int idx = 0x7fffffff;
        bool nonfinite = false;
        for (int c = lane; c < C; c += 32) {
            const float v = row[c];
            nonfinite |= !(fabsf(v) <= FLT_MAX);
            if (v > best) {
                if (logits) sum = sum * __expf(best - v) + 1.f;
                best = v;
                idx = c;
            } else if (logits) {
                sum += __expf(v - best);
            }
        }
        if (nonfinite) bad = 1;   // NaN / Inf in the class scores (fp16 overflow inside the recogniser plan)
        const float lane_best = best;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
            if (ov > best || (ov == best && oi < idx)) { best = ov; idx = oi; }
        }
        if (logits) {
            float s = idx == 0x7fffffff ? 0.f : sum * __expf(lane_best - best);   // lanes without a class (C < 32) hold best = -FLT_MAX, sum = 0
            if (lane_best == -FLT_MAX) s = 0.f;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) { bi[t] = idx < C ? idx : 0; bp[t] = 1.f / s; }
        } else if (lane == 0) {
            bi[t] = idx < C ? idx : 0;
            bp[t] = best;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (bad) {   // reported as id_len = -1: run_frames turns it into a StateError that names the fix
            id_len[n] = -1;
            score[n] = 0.f;
            return;
        }
        int m = 0;
        float s = 0.f;
        int prev = -1;
        for (int t = 0; t < T; t++) {
            const int i = bi[t];
            if (i != prev && i != 0) {
                ids[size_t(n) * max_t + m] = i;
                s += bp[t];
                m++;
            }
            prev = i;
        }
        id_len[n] = m;
        score[n] = m > 0 ? s / float(m) : 0.f;
    }
}

void launch_ctc_decode(const float* probs, int C, int cs, bool logits, const int* toff, const int* tlen, int n, int max_t, int* ids,
                       int* id_len, float* score, cudaStream_t st) {
    if (n <= 0) return;
    size_t smem = size_t(max_t) * 8;
    if (smem > 48 * 1024) {
        static size_t configured[64] = {};
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev < 0 || dev >= 64 || smem > configured[dev]) {
            if (cudaFuncSetAttribute(ctc_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return;
            if (dev >= 0 && dev < 64) configured[dev] = smem;
        }
    }
    pdl_launch(ctc_decode_kernel, n, 128, smem, st, probs, C, cs, logits ? 1 : 0, toff, tlen, max_t, ids, id_len, score);
}

}  // namespace vse
