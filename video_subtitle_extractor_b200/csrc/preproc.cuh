// preproc.cuh — image geometry kernels shared by the pipeline: OpenCV-exact bilinear resize of uint8 images.
//
// Restates cv::resize(INTER_LINEAR) for 8-bit images (the fixed-point path: 11-bit coefficients, horizontal pass
// in int32, vertical pass ((b*(S>>4))>>16 summed, +2, >>2)), which is what the reference's pre-processing calls:
//   det: DetResizeForTest -> cv2.resize(img, (rw, rh))                (paddleocr 2.10)
//   rec: resize_norm_img  -> cv2.resize(crop, (resized_w, imgH))      (paddleocr 2.10)
// Checked bit-exact against cv2 4.13 in tests/test_gpu_parity.py (and the numpy restatement in tests/test_host_cpu.py).
#pragma once
#include <cuda_runtime.h>
#include "pdl.cuh"
#include <cstdint>

namespace vse {

struct ResizeJob {
    const uint8_t* src;  // BGR or BGRX rows
    int src_h, src_w, src_stride, src_pix;  // src_pix = bytes per source pixel (3 or 4)
    int dst_h, dst_w;    // resized size
    int out_w;           // row length of the destination image in pixels (>= dst_w; columns beyond are left untouched)
    long long dst_off;   // destination pixel offset (BGRX pixels)
};

__device__ __forceinline__ void resize_coeff(int d, int dn, int sn, bool clamp_edges, int& s, int& a0, int& a1) {
    // double math with explicit rounding steps = what the host code does (no FMA contraction)
    double inv = __ddiv_rn(double(dn), double(sn));
    double scale = __ddiv_rn(1.0, inv);
    float f = float(__dadd_rn(__dmul_rn(__dadd_rn(double(d), 0.5), scale), -0.5));
    int si = int(floorf(f));
    f -= float(si);
    if (clamp_edges) {
        if (si < 0) { f = 0.f; si = 0; }
        if (si >= sn - 1) { f = 0.f; si = sn - 1; }
    }
    s = si;
    a0 = __float2int_rn((1.f - f) * 2048.f);
    a1 = __float2int_rn(f * 2048.f);
}

// one thread per destination pixel; writes BGRX (X = 0)
__global__ void __launch_bounds__(256) resize_bilinear_u8_kernel(const ResizeJob* jobs, uint8_t* dst, int max_pix) {
    pdl_wait();      // pdl.cuh: nothing below may run before the previous kernel of the stream has completed
    pdl_trigger();
    const ResizeJob j = jobs[blockIdx.y];
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= j.dst_h * j.dst_w) return;
    const int dy = idx / j.dst_w, dx = idx - dy * j.dst_w;
    int sx, a0, a1, sy, b0, b1;
    resize_coeff(dx, j.dst_w, j.src_w, true, sx, a0, a1);
    resize_coeff(dy, j.dst_h, j.src_h, false, sy, b0, b1);
    const int sx1 = min(sx + 1, j.src_w - 1);
    const int y0 = min(max(sy, 0), j.src_h - 1), y1 = min(max(sy + 1, 0), j.src_h - 1);
    const uint8_t* r0 = j.src + size_t(y0) * j.src_stride;
    const uint8_t* r1 = j.src + size_t(y1) * j.src_stride;
    uchar4 o;
    unsigned char* op = reinterpret_cast<unsigned char*>(&o);
#pragma unroll
    for (int c = 0; c < 3; c++) {
        int h0 = int(r0[sx * j.src_pix + c]) * a0 + int(r0[sx1 * j.src_pix + c]) * a1;
        int h1 = int(r1[sx * j.src_pix + c]) * a0 + int(r1[sx1 * j.src_pix + c]) * a1;
        int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
        op[c] = (unsigned char)min(max(v, 0), 255);
    }
    op[3] = 0;
    reinterpret_cast<uchar4*>(dst)[j.dst_off + (long long)dy * j.out_w + dx] = o;
    (void)max_pix;
}

}  // namespace vse
