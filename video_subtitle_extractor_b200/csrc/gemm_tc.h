// gemm_tc.h — tcgen05 / TMA implicit-GEMM convolution (sm_100a): launch interface.
// See gemm_tc.cu for the kernel. One TcConv describes one CONV step on one batch geometry.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <vector>

#include "nn_kernels.h"

namespace vse {

struct TcWeights {          // built once per plan step (host), uploaded as fp16
    int n_chunk = 0;        // accumulator columns per tile (multiple of 16, <= 256)
    int n_chunks = 0;       // output-channel chunks (tiles along N)
    int k_pad = 0;          // per-tap K extent in the B matrix (multiple of 64)
    int taps = 1;
    int tf32 = 0;           // 1: fp32 elements (tcgen05 kind::tf32), k_pad a multiple of 32; 0: fp16, multiple of 64
    int split = 0;          // 1: fp16 hi | lo halves of fp32 weights, 64 columns per 32 input channels (see gemm_tc.cu, "split" mode)
    float w_scale = 1.f;    // split: power of two the weights were multiplied by before the split
    int stack = 0;          // split, narrow K-heavy layers: hi and lo weights stacked along N (rows [0, n) = hi, [n, 2n) = lo), 32 columns
                            // (64 bytes, SWIZZLE_64B) per 32 input channels — see gemm_tc.cu, "stacked split"
    std::vector<uint16_t> b;  // raw 16-bit words: fp16 bits (or 2 words per fp32), [n_chunks * n_chunk][taps * k_pad], K-major, zero padded
};

// cout / cin: real channel counts; weights fp32 [cout][kh*kw][cin]
enum TcMode { TC_F16 = 0, TC_TF32 = 1, TC_SPLIT = 2 };
TcWeights tc_pack_weights(const float* w, int cout, int cin, int taps, int mode = TC_F16);

// Pixel-packed 1x1 conv: `pack` consecutive pixels (channel stride in_cs, pack * in_cs == 64) form ONE GEMM row of K = 64
// and the weights become block-diagonal [pack * out_cs][64]: row g*out_cs + co, column g*in_cs + ci = w[co][ci].  The
// activation rows are then 128 bytes wide (TMA moves narrow 32/64-byte pixel rows at a fraction of its line rate) and the
// `pack` output pixels of a row are contiguous in memory.
TcWeights tc_pack_weights_pixelpacked(const float* w, int cout, int cin, int in_cs, int out_cs, int pack);

float tc_split_activation_scale();

struct TcConv {
    CUtensorMap map_a;      // activations (2-D flat for 1x1, 4-D [C][W][H][N] for KxK)
    CUtensorMap map_b;      // weights
    CUtensorMap map_o;      // output (TMA stores), encoded lazily by launch_conv_tc
    const void* map_o_ptr = nullptr;
    int map_o_cs = 0, map_o_n = 0;
    int spatial = 0;        // 0: 1x1 over a flat pixel list; 1: KxK stride 1 over equal-sized images
    int M = 0;              // flat: number of pixels
    int n_img = 0, H = 0, W = 0, tiles_x = 0, tiles_y = 0;
    int tile_w = 16, tile_h = 8;   // spatial tile (128 pixels): 8 x 16 halo, 16 x 8, 128 x 1 for maps of height 1
    int kh = 1, kw = 1, ph = 0, pw = 0;
    int cin = 0;            // real input channels (the MMA loop skips the all-zero tail of the last K block)
    int rowbox = 0;         // KxK: A boxes span 8 + kh_g - 1 image rows and serve kh_g vertical taps (see gemm_tc.cu)
    int kh_g = 0;           // rowbox: vertical taps per box (= kh when everything fits; 9x9 kernels with streamed weights: 3)
    int tf32 = 0;           // fp32 activations / weights through kind::tf32 MMAs, fp32 output
    int split = 0;          // fp32 activations split in place into fp16 hi | lo, three kind::f16 MMAs per product, fp32 output
    float w_scale = 1.f;
    int stack = 0;          // split with hi / lo weights stacked along N (TcWeights::stack)
    float a_scale = 0.f;    // split: power of two for the operand rows (0 = the default of tc_split_activation_scale())
    int pack = 0;           // >0: pixel-packed flat conv (rows of `pack` pixels)
    // Fused depthwise -> pointwise (split mode): the TMA boxes are (16 + dw - 1) x (8 + dw - 1) pixel windows of the DEPTHWISE
    // convolution's input; the operand rows of the 1x1 convolution are computed from them in shared memory by the transform
    // warps (depthwise taps, bias, activation, post-affine, then the hi | lo split).  dw = depthwise kernel size (0 = none)
    int dw = 0, dw_cp = 0, dw_act = 0;       // dw_cp: channel pitch of the [tap][channel] depthwise filter
    const float* dw_w = nullptr;
    const float* dw_bias = nullptr;
    const float* dw_ps = nullptr;            // optional post-affine of the depthwise step
    const float* dw_pt = nullptr;
    int direct1 = 0;        // flat fp32 convolution with ONE output channel stored densely (a probability map): the epilogue writes
                            // out[pixel * out_cs] itself — a 4-byte pixel pitch is not addressable by a TMA store
    int b_resident = 0;     // the whole weight matrix stays in shared memory for the kernel's lifetime
    int halo = 0;           // KxK: one (16 + kh - 1) x (8 + kw - 1) box per k-block serves every tap (see gemm_tc.cu)
    int num_kb = 1;         // 64-channel K blocks per tap
    int k_pad = 64;
    int n_chunk = 16, n_chunks = 1, n_store = 8;
    int num_m_tiles = 0;
    void* out = nullptr;
    int out_cs = 0;
    // spatial launches: element strides of the OUTPUT tensor between consecutive pixels / rows / images (0 = dense:
    // out_cs, W * out_cs, H * W * out_cs).  A 2x2 stride-2 transposed convolution is four 1x1 convolutions whose outputs
    // interleave: position (dy, dx) writes every other pixel of every other row (engine.cu, OP_DECONV2)
    long long o_px = 0, o_row = 0, o_img = 0;
    Epilogue epi;
    bool valid = false;
};

// Fills map_a/map_b + tile counts. `in`: activation base (fp16, channel stride in_cs), `wdev`: packed weights on device.
// Returns an empty string on success, else the reason the step cannot use the tensor-core path.
std::string tc_conv_setup(TcConv& t, const void* in, int in_cs, int cin, const void* wdev, const TcWeights& w, bool flat,
                          int64_t pixels, int n_img, int H, int W, int kh, int kw, int ph, int pw, bool allow_rowbox = true,
                          bool allow_halo = true);

// returns an empty string on success, else why the launch was not possible (nothing launched)
std::string launch_conv_tc(TcConv& t, int sm_count, cudaStream_t st);
// fused depthwise(dw x dw, stride 1, same padding) -> 1x1 convolution over equal-sized images, split mode only: `in` is the
// DEPTHWISE input; the depthwise parameters go into t.dw_* before the launch (see TcConv::dw)
std::string tc_conv_setup_dwpw(TcConv& t, const void* in, int in_cs, int cin, const void* wdev, const TcWeights& w, int n_img, int H, int W,
                               int dw);

// one group of a ragged launch as the kernel reads it
struct TcGroupDev {
    int tile_begin, n_img, H, W, tiles_x, tiles_y;
    long long pix_off;      // first pixel of the group inside the step's values (residual addressing)
};
// Ragged batch (recogniser crops of different padded widths): every group (run of equal-sized images; each TcConv set up by
// tc_conv_setup on its own slice, out / epi filled) of one KxK convolution in ONE launch.  `dev`: tc_groups_dev_bytes(n) bytes
// of device memory owned by the caller for the lifetime of the context; *uploaded: the table there is current.
size_t tc_groups_dev_bytes(int n_groups);
std::string launch_conv_tc_groups(TcConv* const* groups, const long long* pix_off, int n_groups, void* dev, bool* uploaded,
                                  int sm_count, cudaStream_t st);

}  // namespace vse
