// fast_kernels.h — shape-specialised fp16 kernels (see fast_kernels.cu).  Every launcher returns false when the step
// does not have the shape it was written for; the engine then uses the generic kernel of nn_kernels.cu.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "nn_kernels.h"

namespace vse {

// depthwise KxK (K = 3 | 5, strides 1/2 per axis, "same" padding).  max_strip_units = max over images of
// out_h * ceil(out_w / 4).
bool launch_dwconv_fast(const ConvArgs& a, int max_strip_units, cudaStream_t st);

// same layers, shared-memory tiled with cp.async staging (max_out_h / max_out_w: largest output image of the batch)
bool launch_dwconv_tiled(const ConvArgs& a, int max_out_h, int max_out_w, cudaStream_t st);

// same layers, register tiled: one thread = 3|4 x 4 outputs of one channel pair, filter and tile in registers, no shared memory
bool launch_dwconv_reg(const ConvArgs& a, int max_out_h, int max_out_w, cudaStream_t st, int prec = 0);   // prec 1: fp32 activations

// 3x3 stride-2 stem on uint8 BGRX input, 16 output channels.  max_out_pix_pairs = max over images of out_h * ceil(out_w / 2).
bool launch_stem_fast(const ConvArgs& a, int cout, int max_out_pix_pairs, cudaStream_t st, int prec = 0);

// fused DB head (two 2x2 stride-2 transposed convolutions, 24 -> 24 -> 1, ReLU / sigmoid), fp32 map out
bool launch_db_head_fused(const void* in, int in_cs, int c, const float* w1, const float* b1, const float* w2, const float* b2,
                          float* out, int out_cs, const ImgTab* tin, const ImgTab* tout, int n_img, int max_in_pix,
                          cudaStream_t st, int prec = 0);

// squeeze-excite gate from the partial sums written by launch_gpool_partial (nn_kernels): mean -> FC+act -> FC+act.
// w1 / w2 are the TRANSPOSED (input-major) matrices: w1[c][cm], w2[cm][c]; cm <= 512.
void launch_se_gate(const float* partial, int splits, int c_pad, int c, int cm, const ImgTab* tin, const float* w1,
                    const float* b1, int act1, float slope1, float offset1, const float* w2, const float* b2, int act2,
                    float slope2, float offset2, float* out, int n_img, cudaStream_t st, const float* pre_w = nullptr,
                    const float* pre_b = nullptr, int cx = 0, int cx_pad = 0, int pre_ld = 0);

// one source of a concat gather: a channel slice of the output filled from `in` (nearest-upsampled by scale_px, and/or
// multiplied by a per-image channel gate as CHSCALE does).  Sources must be ordered by ascending slice offset.
struct GatherSrc {
    const void* in; int in_cs; const ImgTab* tin; int scale_px; int shift;   // shift: log2(scale_px) or -1 (set by the launcher)
    const float* scale; int scale_c; int residual;     // optional channel gate (nullptr = plain copy)
    void* out; int out_cs; int cvecs;                  // slice base pointer, row pitch of the concat buffer, slice width / 8
};
void launch_concat_gather(const GatherSrc* src, int n, const ImgTab* tout, int n_img, int max_out_h, int max_out_w, cudaStream_t st,
                          int prec = 0);

}  // namespace vse
