// nn_kernels.cu — generic CUDA-core kernels for every plan step (sm_100a).
// These are the always-correct baseline path of the engine: fp32 accumulate, activations stored as
// __half (production) or float (parity-debug).  The GEMM-shaped steps (1x1 / KxK convolutions with
// enough channels) are taken over by the tcgen05/TMA kernels in gemm_tc.cu; everything here is
// HBM-bound, so the rules that matter are 16-byte vector accesses along the channel axis and fused epilogues.
//
// Reference semantics followed (restated from the shipped graphs):
//   conv2d / depthwise_conv2d / conv2d_transpose(2x2,s2) / pool2d / nearest_interp_v2 / layer_norm /
//   softmax / SVTR attention (backend/models/V4/en_rec_fast/inference.pdmodel ops #227-284).
#include "nn_kernels.h"
#include "pdl.cuh"

#include <cuda_fp16.h>
#include <cfloat>
#include <cstring>

#include "plan.h"

namespace vse {

// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float act_apply(float x, int act, float slope, float offset) {
    switch (act) {
        case ACT_RELU: return fmaxf(x, 0.f);
        case ACT_HSWISH: return x * fminf(fmaxf(x + 3.f, 0.f), 6.f) * (1.f / 6.f);   // no IEEE divide on the hot path
        case ACT_HSIGMOID: return fminf(fmaxf(x * slope + offset, 0.f), 1.f);
        case ACT_SWISH: return x / (1.f + __expf(-x));
        case ACT_SIGMOID: return 1.f / (1.f + __expf(-x));
        case ACT_RELU6: return fminf(fmaxf(x, 0.f), 6.f);
        default: return x;
    }
}

template <typename T> struct V8;
template <> struct V8<__half> {
    static __device__ __forceinline__ void load(const __half* p, float* v) {
        uint4 u = *reinterpret_cast<const uint4*>(p);
        const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            float2 f = __half22float2(h[i]);
            v[2 * i] = f.x;
            v[2 * i + 1] = f.y;
        }
    }
    static __device__ __forceinline__ void store(__half* p, const float* v) {
        uint4 u;
        __half2* h = reinterpret_cast<__half2*>(&u);
#pragma unroll
        for (int i = 0; i < 4; i++) h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
        *reinterpret_cast<uint4*>(p) = u;
    }
    static __device__ __forceinline__ void load4(const __half* p, float* v) {
        uint2 u = *reinterpret_cast<const uint2*>(p);
        const __half2* h = reinterpret_cast<const __half2*>(&u);
        float2 a = __half22float2(h[0]), b = __half22float2(h[1]);
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    }
};
template <> struct V8<float> {
    static __device__ __forceinline__ void load(const float* p, float* v) {
        float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
    static __device__ __forceinline__ void store(float* p, const float* v) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
    static __device__ __forceinline__ void load4(const float* p, float* v) {
        float4 a = *reinterpret_cast<const float4*>(p);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    }
};

__device__ __forceinline__ float to_f(__half x) { return __half2float(x); }
__device__ __forceinline__ float to_f(float x) { return x; }
template <typename T> __device__ __forceinline__ T from_f(float x);
template <> __device__ __forceinline__ __half from_f<__half>(float x) { return __float2half_rn(x); }
template <> __device__ __forceinline__ float from_f<float>(float x) { return x; }

struct EpiDev {
    const float* bias; const float* post_scale; const float* post_shift; const void* res;
    int res_cs, act, act2; float hs_slope, hs_offset;
};
static EpiDev to_dev(const Epilogue& e) {
    return EpiDev{e.bias, e.post_scale, e.post_shift, e.res, e.res_cs, e.act, e.act2, e.hs_slope, e.hs_offset};
}

// epilogue on one value: channel co, output pixel index pix (global)
template <typename T>
__device__ __forceinline__ float epi_apply(const EpiDev& e, float acc, int co, size_t pix) {
    float v = acc + (e.bias ? e.bias[co] : 0.f);
    v = act_apply(v, e.act, e.hs_slope, e.hs_offset);
    if (e.post_scale) v = v * e.post_scale[co] + e.post_shift[co];
    if (e.res) v += to_f(static_cast<const T*>(e.res)[pix * e.res_cs + co]);
    return act_apply(v, e.act2, 0.f, 0.f);
}

static inline int cdiv(int64_t a, int64_t b) { return int((a + b - 1) / b); }

// ------------------------------------------------------------------------------------------------
// CONV / STEM: SIMT implicit GEMM.  Tile 64 output pixels x (16*TN) output channels, K step 16.
// ------------------------------------------------------------------------------------------------
struct ConvDev {
    const void* in; void* out; const float* w; EpiDev epi; const ImgTab* tin; const ImgTab* tout;
    int cin_pad, in_cs, cout_store, out_cs, w_ci, w_co, kh, kw, sh, sw, ph, pw, out_f32;
    float nscale[3], nshift[3];
};
static ConvDev to_dev(const ConvArgs& a) {
    ConvDev d{a.in, a.out, a.w, to_dev(a.epi), a.tin, a.tout, a.cin_pad, a.in_cs, a.cout_store, a.out_cs, a.w_ci, a.w_co,
              a.kh, a.kw, a.sh, a.sw, a.ph, a.pw, a.out_f32, {a.nscale[0], a.nscale[1], a.nscale[2]},
              {a.nshift[0], a.nshift[1], a.nshift[2]}};
    return d;
}

template <typename InT>
__device__ __forceinline__ void conv_load_a(const ConvDev& p, const ImgTab& ti, bool valid, int iy, int ix, int ci, float* v) {
    v[0] = v[1] = v[2] = v[3] = 0.f;
    if (!valid || ci >= p.cin_pad) return;
    size_t pix = size_t(ti.off) + size_t(iy) * ti.w + ix;
    V8<InT>::load4(static_cast<const InT*>(p.in) + pix * p.in_cs + ci, v);
}
template <>
__device__ __forceinline__ void conv_load_a<unsigned char>(const ConvDev& p, const ImgTab& ti, bool valid, int iy, int ix,
                                                           int ci, float* v) {
    v[0] = v[1] = v[2] = v[3] = 0.f;
    if (!valid || ci != 0 || ix >= ti.vw) return;
    size_t pix = size_t(ti.off) + size_t(iy) * ti.w + ix;
    uchar4 u = *reinterpret_cast<const uchar4*>(static_cast<const unsigned char*>(p.in) + pix * 4);
    v[0] = float(u.x) * p.nscale[0] + p.nshift[0];
    v[1] = float(u.y) * p.nscale[1] + p.nshift[1];
    v[2] = float(u.z) * p.nscale[2] + p.nshift[2];
}

template <typename InT, typename OutT, int TN>
__global__ void __launch_bounds__(256) conv_simt_kernel(ConvDev p) {
    pdl_wait();      // pdl.cuh: nothing below may run before the previous kernel of the stream has completed
    pdl_trigger();
    constexpr int BM = 64, BK = 16, BN = 16 * TN, LDA = BM + 4;
    __shared__ __align__(16) float As[BK][LDA];
    __shared__ __align__(16) float Bs[BK][BN];
    const int img = blockIdx.z;
    const ImgTab ti = p.tin[img], to = p.tout[img];
    const int npix = to.h * to.w;
    const int m0 = blockIdx.x * BM;
    if (m0 >= npix) return;
    const int n0 = blockIdx.y * BN;
    const int t = threadIdx.x;
    // A-load role
    const int a_pl = t >> 2, a_cg = (t & 3) * 4;
    const int a_m = m0 + a_pl;
    const bool a_in = a_m < npix;
    const int a_oy = a_in ? a_m / to.w : 0, a_ox = a_in ? a_m % to.w : 0;
    // B-load role
    const int b_k = t >> 4, b_n = (t & 15) * TN;
    // compute role
    const int tx = t & 15, ty = t >> 4;
    float acc[4][TN];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < TN; j++) acc[i][j] = 0.f;

    const int nchunk = (p.cin_pad + BK - 1) / BK;
    const int n_it = p.kh * p.kw * nchunk;
    float ra[4], rb[TN];
    auto gload = [&](int it) {
        int tap = it / nchunk, c0 = (it - tap * nchunk) * BK;
        int ky = tap / p.kw, kx = tap - ky * p.kw;
        int iy = a_oy * p.sh - p.ph + ky, ix = a_ox * p.sw - p.pw + kx;
        bool valid = a_in && iy >= 0 && iy < ti.h && ix >= 0 && ix < ti.w;
        conv_load_a<InT>(p, ti, valid, iy, ix, c0 + a_cg, ra);
        const float* wp = p.w + (size_t(tap) * p.w_ci + c0 + b_k) * p.w_co + n0 + b_n;
        if constexpr (TN == 4) {
            float4 w4 = *reinterpret_cast<const float4*>(wp);
            rb[0] = w4.x; rb[1] = w4.y; rb[2] = w4.z; rb[3] = w4.w;
        } else {
            float2 w2 = *reinterpret_cast<const float2*>(wp);
            rb[0] = w2.x; rb[1] = w2.y;
        }
    };
    gload(0);
    for (int it = 0; it < n_it; it++) {
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 4; j++) As[a_cg + j][a_pl] = ra[j];
#pragma unroll
        for (int j = 0; j < TN; j++) Bs[b_k][b_n + j] = rb[j];
        __syncthreads();
        if (it + 1 < n_it) gload(it + 1);
#pragma unroll
        for (int kk = 0; kk < BK; kk++) {
            float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            float a[4] = {a4.x, a4.y, a4.z, a4.w};
            float b[TN];
#pragma unroll
            for (int j = 0; j < TN; j++) b[j] = Bs[kk][tx * TN + j];
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < TN; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
    }
    // epilogue
    const int co0 = n0 + tx * TN;
    if (co0 >= p.cout_store) return;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        int m = m0 + ty * 4 + i;
        if (m >= npix) continue;
        size_t pix = size_t(to.off) + m;
        float v[TN];
#pragma unroll
        for (int j = 0; j < TN; j++) v[j] = epi_apply<OutT>(p.epi, acc[i][j], co0 + j, pix);
        if (p.out_f32) {
            float* o = static_cast<float*>(p.out) + pix * p.out_cs + co0;
#pragma unroll
            for (int j = 0; j < TN; j++)
                if (co0 + j < p.out_cs) o[j] = v[j];
        } else {
            OutT* o = static_cast<OutT*>(p.out) + pix * p.out_cs + co0;
#pragma unroll
            for (int j = 0; j < TN; j++) o[j] = from_f<OutT>(v[j]);
        }
    }
}

template <typename InT, typename OutT>
static void conv_simt_dispatch(const ConvArgs& a, cudaStream_t st) {
    ConvDev d = to_dev(a);
    int tiles_m = cdiv(a.max_out_pix, 64);
    if (a.cout_store <= 32) {
        dim3 grid(tiles_m, cdiv(a.cout_store, 32), a.n_img);
        pdl_launch(conv_simt_kernel<InT, OutT, 2>, grid, 256, 0, st, d);
    } else {
        dim3 grid(tiles_m, cdiv(a.cout_store, 64), a.n_img);
        pdl_launch(conv_simt_kernel<InT, OutT, 4>, grid, 256, 0, st, d);
    }
}

void launch_conv_simt(const ConvArgs& a, int prec, cudaStream_t st) {
    if (a.in_u8) {
        if (prec == 0) conv_simt_dispatch<unsigned char, __half>(a, st);
        else conv_simt_dispatch<unsigned char, float>(a, st);
    } else {
        if (prec == 0) conv_simt_dispatch<__half, __half>(a, st);
        else conv_simt_dispatch<float, float>(a, st);
    }
}

// ------------------------------------------------------------------------------------------------
// DWCONV: one thread = one output pixel x 8 channels
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) dwconv_kernel(ConvDev p, int cvecs) {
    pdl_wait();      // pdl.cuh: nothing below may run before the previous kernel of the stream has completed
    pdl_trigger();
    const int img = blockIdx.y;
    const ImgTab ti = p.tin[img], to = p.tout[img];
    const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    const int64_t total = int64_t(to.h) * to.w * cvecs;
    if (idx >= total) return;
    const int cv = int(idx % cvecs);
    const int m = int(idx / cvecs);
    const int oy = m / to.w, ox = m - oy * to.w;
    const int c0 = cv * 8;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; j++) acc[j] = 0.f;
    const T* in = static_cast<const T*>(p.in);
    for (int ky = 0; ky < p.kh; ky++) {
        int iy = oy * p.sh - p.ph + ky;
        if (iy < 0 || iy >= ti.h) continue;
        for (int kx = 0; kx < p.kw; kx++) {
            int ix = ox * p.sw - p.pw + kx;
            if (ix < 0 || ix >= ti.w) continue;
            float x[8];
            V8<T>::load(in + (size_t(ti.off) + size_t(iy) * ti.w + ix) * p.in_cs + c0, x);
            const float* wp = p.w + size_t(ky * p.kw + kx) * p.cin_pad + c0;
            float4 w0 = __ldg(reinterpret_cast<const float4*>(wp)), w1 = __ldg(reinterpret_cast<const float4*>(wp + 4));
            acc[0] = fmaf(x[0], w0.x, acc[0]); acc[1] = fmaf(x[1], w0.y, acc[1]);
            acc[2] = fmaf(x[2], w0.z, acc[2]); acc[3] = fmaf(x[3], w0.w, acc[3]);
            acc[4] = fmaf(x[4], w1.x, acc[4]); acc[5] = fmaf(x[5], w1.y, acc[5]);
            acc[6] = fmaf(x[6], w1.z, acc[6]); acc[7] = fmaf(x[7], w1.w, acc[7]);
        }
    }
    size_t pix = size_t(to.off) + m;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; j++) v[j] = epi_apply<T>(p.epi, acc[j], c0 + j, pix);
    V8<T>::store(static_cast<T*>(p.out) + pix * p.out_cs + c0, v);
}

void launch_dwconv(const ConvArgs& a, int prec, cudaStream_t st) {
    ConvDev d = to_dev(a);
    int cvecs = a.cin_pad / 8;
    dim3 grid(cdiv(int64_t(a.max_out_pix) * cvecs, 256), a.n_img);
    if (prec == 0) pdl_launch(dwconv_kernel<__half>, grid, 256, 0, st, d, cvecs);
    else pdl_launch(dwconv_kernel<float>, grid, 256, 0, st, d, cvecs);
}

// ------------------------------------------------------------------------------------------------
// DECONV2 (conv_transpose 2x2 stride 2): one thread = one OUTPUT pixel x up to 8 output channels
// weights: [pos = dy*2+dx][cout_pad8][cin_pad]
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) deconv2_kernel(ConvDev p, int cout, int cgroups, int cout_pad) {
    pdl_wait();      // pdl.cuh: nothing below may run before the previous kernel of the stream has completed
    pdl_trigger();
    const int img = blockIdx.y;
    const ImgTab ti = p.tin[img], to = p.tout[img];
    const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    const int64_t total = int64_t(to.h) * to.w * cgroups;
    if (idx >= total) return;
    const int cg = int(idx % cgroups);
    const int m = int(idx / cgroups);
    const int oy = m / to.w, ox = m - oy * to.w;
    const int iy = oy >> 1, ix = ox >> 1, pos = (oy & 1) * 2 + (ox & 1);
    const T* in = static_cast<const T*>(p.in) + (size_t(ti.off) + size_t(iy) * ti.w + ix) * p.in_cs;
    const int co0 = cg * 8;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; j++) acc[j] = 0.f;
    for (int c = 0; c < p.cin_pad; c += 8) {
        float x[8];
        V8<T>::load(in + c, x);
#pragma unroll
        for (int j = 0; j < 8; j++) {
            if (co0 + j < cout) {
                const float* wp = p.w + (size_t(pos) * cout_pad + co0 + j) * p.cin_pad + c;
                float4 w0 = __ldg(reinterpret_cast<const float4*>(wp)), w1 = __ldg(reinterpret_cast<const float4*>(wp + 4));
                acc[j] += x[0] * w0.x + x[1] * w0.y + x[2] * w0.z + x[3] * w0.w + x[4] * w1.x + x[5] * w1.y + x[6] * w1.z +
                          x[7] * w1.w;
            }
        }
    }
    size_t pix = size_t(to.off) + m;
    if (p.out_f32) {
        float* o = static_cast<float*>(p.out) + pix * p.out_cs;
#pragma unroll
        for (int j = 0; j < 8; j++)
            if (co0 + j < cout) o[co0 + j] = epi_apply<T>(p.epi, acc[j], co0 + j, pix);
    } else {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; j++) v[j] = (co0 + j < cout) ? epi_apply<T>(p.epi, acc[j], co0 + j, pix) : 0.f;
        V8<T>::store(static_cast<T*>(p.out) + pix * p.out_cs + co0, v);
    }
}

void launch_deconv2(const ConvArgs& a, int cout, int prec, cudaStream_t st) {
    ConvDev d = to_dev(a);
    int cgroups = cdiv(cout, 8);
    dim3 grid(cdiv(int64_t(a.max_out_pix) * cgroups, 256), a.n_img);
    if (prec == 0) pdl_launch(deconv2_kernel<__half>, grid, 256, 0, st, d, cout, cgroups, cgroups * 8);
    else pdl_launch(deconv2_kernel<float>, grid, 256, 0, st, d, cout, cgroups, cgroups * 8);
}

// ------------------------------------------------------------------------------------------------
// GPOOL: global average pool, deterministic two-stage reduction
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) gpool_partial_kernel(const T* in, int in_cs, int cvecs, const ImgTab* tin,
                                                           float* partial, int splits) {
    pdl_wait();      // pdl.cuh: nothing below may run before the previous kernel of the stream has completed
    pdl_trigger();
    __shared__ float red[256][8];
    const int img = blockIdx.y, split = blockIdx.x;
    const ImgTab ti = tin[img];
    const int npix = ti.h * ti.w;
    const int per = (npix + splits - 1) / splits;
    const int p0 = split * per, p1 = min(npix, p0 + per);
    const int lanes = 256 / cvecs;  // pixel lanes per channel-vector pass (cvecs <= 256)
    for (int cvb = 0; cvb < cvecs; cvb += 256) {  // cvecs > 256 handled in passes
        const int ncv = min(256, cvecs - cvb);
        const int nl = 256 / ncv;
        const int cv = threadIdx.x % ncv, pl = threadIdx.x / ncv;
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; j++) acc[j] = 0.f;
        if (pl < nl) {
            // four independent loads in flight per thread (the sum order per thread stays fixed: deterministic)
            float acc1[8], acc2[8], acc3[8];
#pragma unroll
            for (int j = 0; j < 8; j++) acc1[j] = acc2[j] = acc3[j] = 0.f;
            const T* base = in + size_t(ti.off) * in_cs + (cvb + cv) * 8;
            int pidx = p0 + pl;
            for (; pidx + 3 * nl < p1; pidx += 4 * nl) {
                float x0[8], x1[8], x2[8], x3[8];
                V8<T>::load(base + size_t(pidx) * in_cs, x0);
                V8<T>::load(base + size_t(pidx + nl) * in_cs, x1);
                V8<T>::load(base + size_t(pidx + 2 * nl) * in_cs, x2);
                V8<T>::load(base + size_t(pidx + 3 * nl) * in_cs, x3);
#pragma unroll
                for (int j = 0; j < 8; j++) { acc[j] += x0[j]; acc1[j] += x1[j]; acc2[j] += x2[j]; acc3[j] += x3[j]; }
            }
            for (; pidx < p1; pidx += nl) {
                float x[8];
                V8<T>::load(base + size_t(pidx) * in_cs, x);
#pragma unroll
                for (int j = 0; j < 8; j++) acc[j] += x[j];
            }
#pragma unroll
            for (int j = 0; j < 8; j++) acc[j] = (acc[j] + acc1[j]) + (acc2[j] + acc3[j]);
        }
#pragma unroll
        for (int j = 0; j < 8; j++) red[threadIdx.x][j] = acc[j];
        __syncthreads();
        if (pl == 0) {
            for (int l = 1; l < nl; l++)
#pragma unroll
                for (int j = 0; j < 8; j++) acc[j] += red[l * ncv + cv][j];
            float* o = partial + (size_t(img) * splits + split) * (cvecs * 8) + (cvb + cv) * 8;
#pragma unroll
            for (int j = 0; j < 8; j++) o[j] = acc[j];
        }
        __syncthreads();
    }
    (void)lanes;
}

__global__ void gpool_final_kernel(const float* partial, int splits, int c_pad, const ImgTab* tin, float* out, int out_c) {
    pdl_wait();      // pdl.cuh: nothing below may run before the previous kernel of the stream has completed
    pdl_trigger();
    const int img = blockIdx.x;
    const float inv = 1.f / float(tin[img].h * tin[img].w);
    for (int c = threadIdx.x; c < out_c; c += blockDim.x) {
        float s = 0.f;
        for (int k = 0; k < splits; k++) s += partial[(size_t(img) * splits + k) * c_pad + c];
        out[size_t(img) * out_c + c] = s * inv;
    }
}

void launch_gpool_partial(const void* in, int in_cs, int c_pad, const ImgTab* tin, int n_img, float* partial, int splits,
                          int prec, cudaStream_t st) {
    dim3 grid(splits, n_img);
    if (prec == 0) pdl_launch(gpool_partial_kernel<__half>, grid, 256, 0, st, static_cast<const __half*>(in), in_cs, c_pad / 8, tin, partial, splits);
    else pdl_launch(gpool_partial_kernel<float>, grid, 256, 0, st, static_cast<const float*>(in), in_cs, c_pad / 8, tin, partial, splits);
}

void launch_gpool(const void* in, int in_cs, int c_pad, const ImgTab* tin, int n_img, int max_pix, float* partial,
                  int splits, float* out, int out_c, int prec, cudaStream_t st) {
    launch_gpool_partial(in, in_cs, c_pad, tin, n_img, partial, splits, prec, st);
    pdl_launch(gpool_final_kernel, n_img, 128, 0, st, partial, splits, c_pad, tin, out, out_c);
    (void)max_pix;
}

// ------------------------------------------------------------------------------------------------
// VECLIN: tiny fully-connected layer on pooled vectors (SE blocks)
// ------------------------------------------------------------------------------------------------
__global__ void veclin_kernel(const float* in, int cin, float* out, int cout, const float* w, EpiDev e) {
    pdl_wait();      // pdl.cuh: nothing below may run before the previous kernel of the stream has completed
    pdl_trigger();
    extern __shared__ float xin[];
    const int img = blockIdx.x;
    for (int i = threadIdx.x; i < cin; i += blockDim.x) xin[i] = in[size_t(img) * cin + i];
    __syncthreads();
    for (int co = threadIdx.x; co < cout; co += blockDim.x) {
        const float* wr = w + size_t(co) * cin;
        float s = 0.f;
        for (int i = 0; i < cin; i++) s = fmaf(wr[i], xin[i], s);
        float v = s + (e.bias ? e.bias[co] : 0.f);
        v = act_apply(v, e.act, e.hs_slope, e.hs_offset);
        if (e.post_scale) v = v * e.post_scale[co] + e.post_shift[co];
        out[size_t(img) * cout + co] = act_apply(v, e.act2, 0.f, 0.f);
    }
}

// Wide layers (the server models' squeeze-excite FCs, up to 2048 x 2048): one warp = NO output channels x NI images, the
// weight rows are read as coalesced float4 (once per NI images instead of once per image by one strided thread), the
// input vectors come from L1/L2, and the lanes' partial sums meet in a shuffle tree.
template <int NO, int NI>
__global__ void __launch_bounds__(256) veclin_warp_kernel(const float* __restrict__ in, int cin, float* __restrict__ out, int cout,
                                                          const float* __restrict__ w, EpiDev e, int n_img) {
    pdl_wait();      // pdl.cuh: nothing below may run before the previous kernel of the stream has completed
    pdl_trigger();
    const int lane = threadIdx.x & 31;
    const int co0 = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * NO, i0 = blockIdx.y * NI;
    if (co0 >= cout) return;
    float acc[NO][NI];
#pragma unroll
    for (int o = 0; o < NO; o++)
#pragma unroll
        for (int i = 0; i < NI; i++) acc[o][i] = 0.f;
    for (int k = lane * 4; k < cin; k += 128) {
        float4 wv[NO], xv[NI];
#pragma unroll
        for (int o = 0; o < NO; o++)
            wv[o] = co0 + o < cout ? __ldg(reinterpret_cast<const float4*>(w + size_t(co0 + o) * cin + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < NI; i++)
            xv[i] = i0 + i < n_img ? *reinterpret_cast<const float4*>(in + size_t(i0 + i) * cin + k) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int o = 0; o < NO; o++)
#pragma unroll
            for (int i = 0; i < NI; i++) {
                acc[o][i] = fmaf(wv[o].x, xv[i].x, acc[o][i]);
                acc[o][i] = fmaf(wv[o].y, xv[i].y, acc[o][i]);
                acc[o][i] = fmaf(wv[o].z, xv[i].z, acc[o][i]);
                acc[o][i] = fmaf(wv[o].w, xv[i].w, acc[o][i]);
            }
    }
#pragma unroll
    for (int o = 0; o < NO; o++)
#pragma unroll
        for (int i = 0; i < NI; i++)
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) acc[o][i] += __shfl_xor_sync(0xffffffffu, acc[o][i], off);
    if (lane == 0) {
#pragma unroll
        for (int o = 0; o < NO; o++) {
            const int co = co0 + o;
            if (co >= cout) break;
#pragma unroll
            for (int i = 0; i < NI; i++) {
                if (i0 + i >= n_img) break;
                float v = acc[o][i] + (e.bias ? e.bias[co] : 0.f);
                v = act_apply(v, e.act, e.hs_slope, e.hs_offset);
                if (e.post_scale) v = v * e.post_scale[co] + e.post_shift[co];
                out[size_t(i0 + i) * cout + co] = act_apply(v, e.act2, 0.f, 0.f);
            }
        }
    }
}

void launch_veclin(const float* in, int cin, float* out, int cout, const float* w, const Epilogue& epi, int n_img,
                   cudaStream_t st) {
    if (cin >= 128 && cin % 4 == 0 && !(reinterpret_cast<uintptr_t>(w) & 15) && !(reinterpret_cast<uintptr_t>(in) & 15)) {
        constexpr int NO = 4, NI = 4;
        dim3 grid(cdiv(cout, NO * 8), cdiv(n_img, NI));
        pdl_launch(veclin_warp_kernel<NO, NI>, grid, 256, 0, st, in, cin, out, cout, w, to_dev(epi), n_img);
        return;
    }
    pdl_launch(veclin_kernel, n_img, 128, cin * sizeof(float), st, in, cin, out, cout, w, to_dev(epi));
}

// ------------------------------------------------------------------------------------------------
// CHSCALE: y = x * s[img][c]  (+ x)
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) chscale_kernel(const T* in, int in_cs, T* out, int out_cs, int cvecs, const float* scale,
                                                     int scale_c, int residual, const ImgTab* tab) {
    pdl_wait();      // pdl.cuh: nothing below may run before the previous kernel of the stream has completed
    pdl_trigger();
    const int img = blockIdx.y;
    const ImgTab ti = tab[img];
    const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= int64_t(ti.h) * ti.w * cvecs) return;
    const int cv = int(idx % cvecs);
    const size_t pix = size_t(ti.off) + size_t(idx / cvecs);
    float x[8], v[8];
    V8<T>::load(in + pix * in_cs + cv * 8, x);
#pragma unroll
    for (int j = 0; j < 8; j++) {
        int c = cv * 8 + j;
        float s = c < scale_c ? scale[size_t(img) * scale_c + c] : 0.f;
        v[j] = residual ? x[j] + x[j] * s : x[j] * s;
    }
    V8<T>::store(out + pix * out_cs + cv * 8, v);
}

void launch_chscale(const void* in, int in_cs, void* out, int out_cs, int c_pad, const float* scale, int scale_c,
                    int residual, const ImgTab* tab, int n_img, int max_pix, int prec, cudaStream_t st) {
    int cvecs = c_pad / 8;
    dim3 grid(cdiv(int64_t(max_pix) * cvecs, 256), n_img);
    if (prec == 0) pdl_launch(chscale_kernel<__half>, grid, 256, 0, st, (const __half*)in, in_cs, (__half*)out, out_cs, cvecs, scale, scale_c, residual, tab);
    else pdl_launch(chscale_kernel<float>, grid, 256, 0, st, (const float*)in, in_cs, (float*)out, out_cs, cvecs, scale, scale_c, residual, tab);
}

// ------------------------------------------------------------------------------------------------
// POOL (max / avg, ceil handled by the output geometry, exclusive = divide by valid count)
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) pool_kernel(const T* in, int in_cs, T* out, int out_cs, int cvecs, const ImgTab* tin,
                                                  const ImgTab* tout, int kh, int kw, int sh, int sw, int ph, int pw, int is_max,
                                                  int exclusive) {
    pdl_wait();      // pdl.cuh: nothing below may run before the previous kernel of the stream has completed
    pdl_trigger();
    const int img = blockIdx.y;
    const ImgTab ti = tin[img], to = tout[img];
    const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= int64_t(to.h) * to.w * cvecs) return;
    const int cv = int(idx % cvecs);
    const int m = int(idx / cvecs);
    const int oy = m / to.w, ox = m - oy * to.w;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; j++) acc[j] = is_max ? -FLT_MAX : 0.f;
    int cnt = 0;
    for (int ky = 0; ky < kh; ky++) {
        int iy = oy * sh - ph + ky;
        if (iy < 0 || iy >= ti.h) continue;
        for (int kx = 0; kx < kw; kx++) {
            int ix = ox * sw - pw + kx;
            if (ix < 0 || ix >= ti.w) continue;
            float x[8];
            V8<T>::load(in + (size_t(ti.off) + size_t(iy) * ti.w + ix) * in_cs + cv * 8, x);
#pragma unroll
            for (int j = 0; j < 8; j++) acc[j] = is_max ? fmaxf(acc[j], x[j]) : acc[j] + x[j];
            cnt++;
        }
    }
    if (!is_max) {
        // non-exclusive: windows clipped by ceil_mode still divide by the part inside the padded image
        int div = cnt;
        if (!exclusive) {
            int y0 = oy * sh - ph, x0 = ox * sw - pw;
            int y1 = min(y0 + kh, ti.h + ph), x1 = min(x0 + kw, ti.w + pw);
            div = (y1 - y0) * (x1 - x0);
        }
        float inv = 1.f / float(max(div, 1));
#pragma unroll
        for (int j = 0; j < 8; j++) acc[j] *= inv;
    }
    V8<T>::store(out + (size_t(to.off) + m) * out_cs + cv * 8, acc);
}

void launch_pool(const void* in, int in_cs, void* out, int out_cs, int c_pad, const ImgTab* tin, const ImgTab* tout,
                 int n_img, int max_out_pix, int kh, int kw, int sh, int sw, int ph, int pw, int is_max, int exclusive,
                 int prec, cudaStream_t st) {
    int cvecs = c_pad / 8;
    dim3 grid(cdiv(int64_t(max_out_pix) * cvecs, 256), n_img);
    if (prec == 0) pdl_launch(pool_kernel<__half>, grid, 256, 0, st, (const __half*)in, in_cs, (__half*)out, out_cs, cvecs, tin, tout, kh, kw, sh, sw, ph, pw, is_max, exclusive);
    else pdl_launch(pool_kernel<float>, grid, 256, 0, st, (const float*)in, in_cs, (float*)out, out_cs, cvecs, tin, tout, kh, kw, sh, sw, ph, pw, is_max, exclusive);
}

// ------------------------------------------------------------------------------------------------
// UPSAMPLE nearest (integer scale) [+ add]
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) upsample_kernel(const T* in, int in_cs, const T* add, int add_cs, T* out, int out_cs,
                                                      int cvecs, const ImgTab* tin, const ImgTab* tout, int scale) {
    pdl_wait();      // pdl.cuh: nothing below may run before the previous kernel of the stream has completed
    pdl_trigger();
    const int img = blockIdx.y;
    const ImgTab ti = tin[img], to = tout[img];
    const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= int64_t(to.h) * to.w * cvecs) return;
    const int cv = int(idx % cvecs);
    const int m = int(idx / cvecs);
    const int oy = m / to.w, ox = m - oy * to.w;
    const int iy = min(oy / scale, ti.h - 1), ix = min(ox / scale, ti.w - 1);
    float x[8];
    V8<T>::load(in + (size_t(ti.off) + size_t(iy) * ti.w + ix) * in_cs + cv * 8, x);
    const size_t pix = size_t(to.off) + m;
    if (add) {
        float y[8];
        V8<T>::load(add + pix * add_cs + cv * 8, y);
#pragma unroll
        for (int j = 0; j < 8; j++) x[j] += y[j];
    }
    V8<T>::store(out + pix * out_cs + cv * 8, x);
}

void launch_upsample(const void* in, int in_cs, const void* add, int add_cs, void* out, int out_cs, int c_pad,
                     const ImgTab* tin, const ImgTab* tout, int n_img, int max_out_pix, int scale, int prec,
                     cudaStream_t st) {
    int cvecs = c_pad / 8;
    dim3 grid(cdiv(int64_t(max_out_pix) * cvecs, 256), n_img);
    if (prec == 0) pdl_launch(upsample_kernel<__half>, grid, 256, 0, st, (const __half*)in, in_cs, (const __half*)add, add_cs, (__half*)out, out_cs, cvecs, tin, tout, scale);
    else pdl_launch(upsample_kernel<float>, grid, 256, 0, st, (const float*)in, in_cs, (const float*)add, add_cs, (float*)out, out_cs, cvecs, tin, tout, scale);
}

// ------------------------------------------------------------------------------------------------
// flat elementwise family
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) add_kernel(const T* a, int a_cs, const T* b, int b_cs, void* out, int out_cs, int cvecs,
                                                 int c_real, int64_t pixels, int act, int out_f32) {
    pdl_wait();      // pdl.cuh: nothing below may run before the previous kernel of the stream has completed
    pdl_trigger();
    const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= pixels * cvecs) return;
    const int cv = int(idx % cvecs);
    const size_t pix = size_t(idx / cvecs);
    float x[8], y[8];
    V8<T>::load(a + pix * a_cs + cv * 8, x);
    V8<T>::load(b + pix * b_cs + cv * 8, y);
#pragma unroll
    for (int j = 0; j < 8; j++) x[j] = act_apply(x[j] + y[j], act, 0.f, 0.f);
    if (out_f32) {
        float* o = static_cast<float*>(out) + pix * out_cs;
#pragma unroll
        for (int j = 0; j < 8; j++)
            if (cv * 8 + j < c_real) o[cv * 8 + j] = x[j];
    } else {
        V8<T>::store(static_cast<T*>(out) + pix * out_cs + cv * 8, x);
    }
}

void launch_add(const void* a, int a_cs, const void* b, int b_cs, void* out, int out_cs, int c_pad, int c_real,
                int64_t pixels, int act, int out_f32, int prec, cudaStream_t st) {
    int cvecs = c_pad / 8;
    int grid = cdiv(pixels * cvecs, 256);
    if (prec == 0) pdl_launch(add_kernel<__half>, grid, 256, 0, st, (const __half*)a, a_cs, (const __half*)b, b_cs, out, out_cs, cvecs, c_real, pixels, act, out_f32);
    else pdl_launch(add_kernel<float>, grid, 256, 0, st, (const float*)a, a_cs, (const float*)b, b_cs, out, out_cs, cvecs, c_real, pixels, act, out_f32);
}

template <typename T>
__global__ void __launch_bounds__(256) eltwise_kernel(const T* in, int in_cs, void* out, int out_cs, int cvecs, int c_real,
                                                     int64_t pixels, const float* scale, const float* shift, int act,
                                                     float hs_slope, float hs_offset, int out_f32) {
    pdl_wait();      // pdl.cuh: nothing below may run before the previous kernel of the stream has completed
    pdl_trigger();
    const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= pixels * cvecs) return;
    const int cv = int(idx % cvecs);
    const size_t pix = size_t(idx / cvecs);
    float x[8];
    V8<T>::load(in + pix * in_cs + cv * 8, x);
#pragma unroll
    for (int j = 0; j < 8; j++) {
        int c = cv * 8 + j;
        float v = c < c_real ? x[j] * scale[c] + shift[c] : 0.f;
        x[j] = c < c_real ? act_apply(v, act, hs_slope, hs_offset) : 0.f;
    }
    if (out_f32) {
        float* o = static_cast<float*>(out) + pix * out_cs;
#pragma unroll
        for (int j = 0; j < 8; j++)
            if (cv * 8 + j < c_real) o[cv * 8 + j] = x[j];
    } else {
        V8<T>::store(static_cast<T*>(out) + pix * out_cs + cv * 8, x);
    }
}

void launch_eltwise(const void* in, int in_cs, void* out, int out_cs, int c_pad, int c_real, int64_t pixels,
                    const float* scale, const float* shift, int act, float hs_slope, float hs_offset, int out_f32,
                    int prec, cudaStream_t st) {
    int cvecs = c_pad / 8;
    int grid = cdiv(pixels * cvecs, 256);
    if (prec == 0) pdl_launch(eltwise_kernel<__half>, grid, 256, 0, st, (const __half*)in, in_cs, out, out_cs, cvecs, c_real, pixels, scale, shift, act, hs_slope, hs_offset, out_f32);
    else pdl_launch(eltwise_kernel<float>, grid, 256, 0, st, (const float*)in, in_cs, out, out_cs, cvecs, c_real, pixels, scale, shift, act, hs_slope, hs_offset, out_f32);
}

template <typename T>
__global__ void __launch_bounds__(256) copy_kernel(const T* in, int in_cs, T* out, int out_cs, int cvecs, int64_t pixels) {
    pdl_wait();      // pdl.cuh: nothing below may run before the previous kernel of the stream has completed
    pdl_trigger();
    const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= pixels * cvecs) return;
    const int cv = int(idx % cvecs);
    const size_t pix = size_t(idx / cvecs);
    float x[8];
    V8<T>::load(in + pix * in_cs + cv * 8, x);
    V8<T>::store(out + pix * out_cs + cv * 8, x);
}

// concat slice at a channel offset that is not a multiple of 8 (e.g. the 1 + 64 channel concat of PFHeadLocal in
// reference backend/models/V4/ch_det): scalar copy of `c` channels, then zeros up to the next multiple of 8 so that the
// consumer's padded channels hold finite values
template <typename T>
__global__ void __launch_bounds__(256) copy_unaligned_kernel(const T* in, int in_cs, T* out, int out_cs, int c, int c_fill,
                                                            int64_t pixels) {
    pdl_wait();      // pdl.cuh: nothing below may run before the previous kernel of the stream has completed
    pdl_trigger();
    const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= pixels * c_fill) return;
    const int ch = int(idx % c_fill);
    const size_t pix = size_t(idx / c_fill);
    out[pix * out_cs + ch] = ch < c ? in[pix * in_cs + ch] : from_f<T>(0.f);
}

void launch_copy_unaligned(const void* in, int in_cs, void* out, int out_cs, int c, int c_fill, int64_t pixels, int prec,
                           cudaStream_t st) {
    int grid = cdiv(pixels * c_fill, 256);
    if (prec == 0) pdl_launch(copy_unaligned_kernel<__half>, grid, 256, 0, st, (const __half*)in, in_cs, (__half*)out, out_cs, c, c_fill, pixels);
    else pdl_launch(copy_unaligned_kernel<float>, grid, 256, 0, st, (const float*)in, in_cs, (float*)out, out_cs, c, c_fill, pixels);
}

void launch_copy(const void* in, int in_cs, void* out, int out_cs, int c_pad, int64_t pixels, int prec, cudaStream_t st) {
    int cvecs = c_pad / 8;
    int grid = cdiv(pixels * cvecs, 256);
    if (prec == 0) pdl_launch(copy_kernel<__half>, grid, 256, 0, st, (const __half*)in, in_cs, (__half*)out, out_cs, cvecs, pixels);
    else pdl_launch(copy_kernel<float>, grid, 256, 0, st, (const float*)in, in_cs, (float*)out, out_cs, cvecs, pixels);
}

// ------------------------------------------------------------------------------------------------
// LAYERNORM over channels: one warp per pixel row
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) layernorm_kernel(const T* in, int in_cs, T* out, int out_cs, int c, int64_t pixels,
                                                       const float* gamma, const float* beta, float eps) {
    pdl_wait();      // pdl.cuh: nothing below may run before the previous kernel of the stream has completed
    pdl_trigger();
    const int64_t row = int64_t(blockIdx.x) * (blockDim.x / 32) + (threadIdx.x >> 5);
    if (row >= pixels) return;
    const int lane = threadIdx.x & 31;
    const T* x = in + size_t(row) * in_cs;
    float s = 0.f;
    for (int i = lane; i < c; i += 32) s += to_f(x[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / float(c);
    float vs = 0.f;
    for (int i = lane; i < c; i += 32) {
        float d = to_f(x[i]) - mean;
        vs += d * d;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) vs += __shfl_xor_sync(0xffffffffu, vs, o);
    const float rstd = rsqrtf(vs / float(c) + eps);
    T* y = out + size_t(row) * out_cs;
    const int c_pad = (c + 7) / 8 * 8;
    for (int i = lane; i < c_pad; i += 32)
        y[i] = i < c ? from_f<T>((to_f(x[i]) - mean) * rstd * gamma[i] + beta[i]) : from_f<T>(0.f);
}

void launch_layernorm(const void* in, int in_cs, void* out, int out_cs, int c, int64_t pixels, const float* gamma,
                      const float* beta, float eps, int prec, cudaStream_t st) {
    int grid = cdiv(pixels, 8);
    if (prec == 0) pdl_launch(layernorm_kernel<__half>, grid, 256, 0, st, (const __half*)in, in_cs, (__half*)out, out_cs, c, pixels, gamma, beta, eps);
    else pdl_launch(layernorm_kernel<float>, grid, 256, 0, st, (const float*)in, in_cs, (float*)out, out_cs, c, pixels, gamma, beta, eps);
}

// ------------------------------------------------------------------------------------------------
// ATTENTION (SVTR global mixing): qkv [T][3][heads][dim] per image -> ctx [T][heads*dim]
// grid (query tiles of 32, heads, images); K and V of one head staged in shared memory.
// ------------------------------------------------------------------------------------------------
int attention_smem_bytes(int max_t, int dim) { return (2 * max_t * dim + 8 * max_t) * int(sizeof(float)); }

template <typename T>
__global__ void __launch_bounds__(256) attention_kernel(const T* qkv, int qkv_cs, T* out, int out_cs, int heads, int dim,
                                                       float qscale, const ImgTab* tab) {
    pdl_wait();      // pdl.cuh: nothing below may run before the previous kernel of the stream has completed
    pdl_trigger();
    extern __shared__ float sm[];
    const int img = blockIdx.z, head = blockIdx.y;
    const ImgTab ti = tab[img];
    const int Tn = ti.h * ti.w;
    const int q0 = blockIdx.x * 32;
    if (q0 >= Tn) return;
    float* Ks = sm;
    float* Vs = sm + size_t(Tn) * dim;
    float* Ps = Vs + size_t(Tn) * dim;  // [8 warps][Tn]
    const T* base = qkv + size_t(ti.off) * qkv_cs;
    const int hd = heads * dim;
    for (int i = threadIdx.x; i < Tn * dim; i += blockDim.x) {
        int tt = i / dim, d = i - tt * dim;
        Ks[i] = to_f(base[size_t(tt) * qkv_cs + hd + head * dim + d]);
        Vs[i] = to_f(base[size_t(tt) * qkv_cs + 2 * hd + head * dim + d]);
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* P = Ps + size_t(warp) * Tn;
    __shared__ float Qs[8][32];
    for (int qi = q0 + warp; qi < min(q0 + 32, Tn); qi += 8) {
        // scaled q row of this warp -> shared memory (dim <= 32)
        Qs[warp][lane] = lane < dim ? to_f(base[size_t(qi) * qkv_cs + head * dim + lane]) * qscale : 0.f;
        __syncwarp();
        float mx = -FLT_MAX;
        for (int j = lane; j < Tn; j += 32) {
            float s = 0.f;
            for (int d = 0; d < dim; d++) s = fmaf(Qs[warp][d], Ks[j * dim + d], s);
            P[j] = s;
            mx = fmaxf(mx, s);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float sum = 0.f;
        for (int j = lane; j < Tn; j += 32) {
            float e = __expf(P[j] - mx);
            P[j] = e;
            sum += e;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        __syncwarp();
        const float inv = 1.f / sum;
        if (lane < dim) {
            float acc = 0.f;
            for (int j = 0; j < Tn; j++) acc = fmaf(P[j], Vs[j * dim + lane], acc);
            out[(size_t(ti.off) + qi) * out_cs + head * dim + lane] = from_f<T>(acc * inv);
        }
        __syncwarp();
    }
}

void launch_attention(const void* qkv, int qkv_cs, void* out, int out_cs, int heads, int dim, float qscale,
                      const ImgTab* tab, int n_img, int max_t, int prec, cudaStream_t st) {
    int smem = attention_smem_bytes(max_t, dim);
    dim3 grid(cdiv(max_t, 32), heads, n_img);
    if (prec == 0) {
        cudaFuncSetAttribute(attention_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        pdl_launch(attention_kernel<__half>, grid, 256, smem, st, (const __half*)qkv, qkv_cs, (__half*)out, out_cs, heads, dim, qscale, tab);
    } else {
        cudaFuncSetAttribute(attention_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        pdl_launch(attention_kernel<float>, grid, 256, smem, st, (const float*)qkv, qkv_cs, (float*)out, out_cs, heads, dim, qscale, tab);
    }
}

// ------------------------------------------------------------------------------------------------
// SOFTMAX over channels -> dense float32 probabilities. One block (128 threads) per row.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128) softmax_kernel(const T* in, int in_cs, float* out, int c, int64_t pixels) {
    pdl_wait();      // pdl.cuh: nothing below may run before the previous kernel of the stream has completed
    pdl_trigger();
    __shared__ float red[4];
    const int64_t row = blockIdx.x;
    const T* x = in + size_t(row) * in_cs;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float mx = -FLT_MAX;
    for (int i = threadIdx.x; i < c; i += 128) mx = fmaxf(mx, to_f(x[i]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
    __syncthreads();
    float s = 0.f;
    for (int i = threadIdx.x; i < c; i += 128) s += __expf(to_f(x[i]) - mx);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    s = red[0] + red[1] + red[2] + red[3];
    const float inv = 1.f / s;
    float* y = out + size_t(row) * c;
    for (int i = threadIdx.x; i < c; i += 128) y[i] = __expf(to_f(x[i]) - mx) * inv;
}

void launch_softmax(const void* in, int in_cs, float* out, int c, int64_t pixels, int prec, cudaStream_t st) {
    if (pixels <= 0) return;
    if (prec == 0) pdl_launch(softmax_kernel<__half>, (unsigned)pixels, 128, 0, st, (const __half*)in, in_cs, out, c, pixels);
    else pdl_launch(softmax_kernel<float>, (unsigned)pixels, 128, 0, st, (const float*)in, in_cs, out, c, pixels);
}

// ------------------------------------------------------------------------------------------------
// debug: any value -> dense float32
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void to_float_kernel(const T* in, int in_cs, float* out, int c, int64_t pixels) {
    pdl_wait();      // pdl.cuh: nothing below may run before the previous kernel of the stream has completed
    pdl_trigger();
    const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= pixels * c) return;
    const int64_t pix = idx / c;
    const int ch = int(idx - pix * c);
    out[idx] = to_f(in[size_t(pix) * in_cs + ch]);
}

void launch_to_float(const void* in, int in_cs, int dtype_is_f32, float* out, int c, int64_t pixels, int prec,
                     cudaStream_t st) {
    if (pixels * c <= 0) return;
    int grid = cdiv(pixels * c, 256);
    if (dtype_is_f32 || prec == 1) pdl_launch(to_float_kernel<float>, grid, 256, 0, st, (const float*)in, in_cs, out, c, pixels);
    else pdl_launch(to_float_kernel<__half>, grid, 256, 0, st, (const __half*)in, in_cs, out, c, pixels);
}

// ------------------------------------------------------------------------------------------------
// small host -> device tables (job lists, image tables): carried in the kernel PARAMETERS instead of a
// cudaMemcpyAsync, so they never queue on the copy engine behind a multi-megabyte frame prefetch, and the host
// buffer is free again as soon as the launch call returns.
// ------------------------------------------------------------------------------------------------
struct UploadBlob { uint32_t w[1000]; };
__global__ void upload_kernel(UploadBlob b, uint32_t* dst, int nwords) {
    pdl_wait();      // pdl.cuh: nothing below may run before the previous kernel of the stream has completed
    pdl_trigger();
    for (int i = threadIdx.x; i < nwords; i += blockDim.x) dst[i] = b.w[i];
}

int launch_upload(void* dst, const void* src_host, size_t bytes, cudaStream_t st) {
    const size_t chunk = sizeof(UploadBlob);
    int n = 0;
    for (size_t off = 0; off < bytes; off += chunk, n++) {
        UploadBlob b;
        const size_t m = bytes - off < chunk ? bytes - off : chunk;
        std::memcpy(b.w, static_cast<const char*>(src_host) + off, m);
        // destination buffers are DevBuf allocations (256-byte slack): rounding the tail up to 4 bytes stays inside
        pdl_launch(upload_kernel, 1, 256, 0, st, b, reinterpret_cast<uint32_t*>(static_cast<char*>(dst) + off), int((m + 3) / 4));
    }
    return n;
}

}  // namespace vse
