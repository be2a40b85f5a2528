// nn_kernels.h — launch interface of the generic (CUDA-core) step kernels.
// Activations are channel-last "pixel-major" [pixels][cstride]; a batch is a list of images
// (ImgTab rows) that may differ in width (ragged recogniser batches).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace vse {

struct ImgTab {
    int off;  // first pixel of this image in the value's pixel-major buffer
    int h, w;
    int vw;   // valid width (u8 input only): columns >= vw are zero after normalisation
};

struct Epilogue {
    const float* bias = nullptr;        // [>= channels computed], zero padded
    const float* post_scale = nullptr;  // optional
    const float* post_shift = nullptr;
    const void* res = nullptr;          // optional residual, same geometry as the output
    int res_cs = 0;
    int act = 0, act2 = 0;
    float hs_slope = 0.f, hs_offset = 0.f;
    // tensor-core path only: per-image squeeze-excite gate applied as v += v * gate[img][channel] (fused RSE block)
    const float* gate = nullptr;
    int gate_c = 0;     // channels per image in `gate` (multiple of 16)
    int gate_rows = 0;  // GEMM rows (pixels, or packed pixel groups) per image
};

struct ConvArgs {
    const void* in = nullptr;
    void* out = nullptr;
    const float* w = nullptr;  // CONV/STEM: [tap][w_ci][w_co]; DWCONV: [tap][c_pad]; DECONV2: [pos][cout_pad][cin_pad]
    Epilogue epi;
    const ImgTab* tin = nullptr;
    const ImgTab* tout = nullptr;
    int n_img = 0;
    int max_out_pix = 0;  // max over images of ho*wo
    int cin_pad = 0, in_cs = 0;
    int cout_store = 0, out_cs = 0;
    int w_ci = 0, w_co = 0;
    int kh = 1, kw = 1, sh = 1, sw = 1, ph = 0, pw = 0;
    int in_u8 = 0;     // STEM: input is uint8 BGRX
    int out_f32 = 0;   // output value is dense float32
    float nscale[3] = {1, 1, 1}, nshift[3] = {0, 0, 0};
};

// prec: 0 = __half activations, 1 = float activations
void launch_conv_simt(const ConvArgs& a, int prec, cudaStream_t st);
void launch_dwconv(const ConvArgs& a, int prec, cudaStream_t st);
void launch_deconv2(const ConvArgs& a, int cout, int prec, cudaStream_t st);

void launch_gpool(const void* in, int in_cs, int c_pad, const ImgTab* tin, int n_img, int max_pix, float* partial,
                  int splits, float* out, int out_c, int prec, cudaStream_t st);
// first stage only: partial[img][split][c_pad] channel sums (deterministic; the caller finishes the reduction)
void launch_gpool_partial(const void* in, int in_cs, int c_pad, const ImgTab* tin, int n_img, float* partial, int splits,
                          int prec, cudaStream_t st);
void launch_veclin(const float* in, int cin, float* out, int cout, const float* w, const Epilogue& epi, int n_img,
                   cudaStream_t st);
void launch_chscale(const void* in, int in_cs, void* out, int out_cs, int c_pad, const float* scale, int scale_c,
                    int residual, const ImgTab* tab, int n_img, int max_pix, int prec, cudaStream_t st);
void launch_pool(const void* in, int in_cs, void* out, int out_cs, int c_pad, const ImgTab* tin, const ImgTab* tout,
                 int n_img, int max_out_pix, int kh, int kw, int sh, int sw, int ph, int pw, int is_max, int exclusive,
                 int prec, cudaStream_t st);
void launch_upsample(const void* in, int in_cs, const void* add, int add_cs, void* out, int out_cs, int c_pad,
                     const ImgTab* tin, const ImgTab* tout, int n_img, int max_out_pix, int scale, int prec,
                     cudaStream_t st);
// flat elementwise family over total pixels
void launch_add(const void* a, int a_cs, const void* b, int b_cs, void* out, int out_cs, int c_pad, int c_real,
                int64_t pixels, int act, int out_f32, int prec, cudaStream_t st);
void launch_eltwise(const void* in, int in_cs, void* out, int out_cs, int c_pad, int c_real, int64_t pixels,
                    const float* scale, const float* shift, int act, float hs_slope, float hs_offset, int out_f32,
                    int prec, cudaStream_t st);
void launch_copy(const void* in, int in_cs, void* out, int out_cs, int c_pad, int64_t pixels, int prec, cudaStream_t st);
// `out` already points at the (unaligned) first channel of the slice; writes c channels + zeros up to c_fill
void launch_copy_unaligned(const void* in, int in_cs, void* out, int out_cs, int c, int c_fill, int64_t pixels, int prec,
                           cudaStream_t st);
void launch_layernorm(const void* in, int in_cs, void* out, int out_cs, int c, int64_t pixels, const float* gamma,
                      const float* beta, float eps, int prec, cudaStream_t st);
void launch_attention(const void* qkv, int qkv_cs, void* out, int out_cs, int heads, int dim, float qscale,
                      const ImgTab* tab, int n_img, int max_t, int prec, cudaStream_t st);
void launch_softmax(const void* in, int in_cs, float* out, int c, int64_t pixels, int prec, cudaStream_t st);
// value -> dense float32 [pixels][c] (debug dumps)
void launch_to_float(const void* in, int in_cs, int dtype_is_f32, float* out, int c, int64_t pixels, int prec,
                     cudaStream_t st);

int attention_smem_bytes(int max_t, int dim);

// lstm.cu — recurrent half of a (bi)LSTM layer: gates [pixels][ndir*4*hidden] (W_ih x + b, from the preceding CONV step)
// -> out [pixels][ndir*hidden]; every image (H == 1) is one sequence over its padded width.  One 8-CTA cluster per
// (direction, 4 lines) keeps W_hh resident in distributed shared memory.
bool lstm_supported(int hidden);
size_t lstm_packed_weight_floats(int hidden, int ndir);
void lstm_pack_weights(const float* w_hh, int hidden, int ndir, float* dst);
cudaError_t launch_lstm(const void* gates, int gates_cs, void* out, int out_cs, const float* w_packed, const ImgTab* tab, int n_img,
                        int ndir, int prec, cudaStream_t st);

// host table -> device through kernel parameters (no copy-engine traffic); returns the number of launches
int launch_upload(void* dst, const void* src_host, size_t bytes, cudaStream_t st);

}  // namespace vse
