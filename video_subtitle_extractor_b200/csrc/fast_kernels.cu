// fast_kernels.cu — shape-specialised fp16 kernels for the HBM-bound steps of the shipped "mobile" graphs
// (reference backend/models/V4/ch_det_fast, V4/*_rec_fast: PP-LCNetV3 blocks = depthwise KxK + 1x1, DB head = two
// conv_transpose 2x2 s2).  Same arithmetic as the generic kernels in nn_kernels.cu (fp32
// accumulate, fp16 storage); what changes is the work per thread:
//   * depthwise KxK: one thread = a strip of TW output pixels x 8 channels, so every input vector is loaded once per
//     strip row instead of once per tap, and the filter row lives in registers;
//   * stem 3x3 s2 (u8 BGRX -> 16 ch): two output pixels per thread, filter bank in shared memory, normalisation fused;
//   * DB head: conv_transpose(24->24)+ReLU and conv_transpose(24->1)+sigmoid fused — one thread turns one neck pixel
//     into its 4x4 patch of the probability map; the 272x480x24 intermediate never goes to HBM.
// The engine falls back to the generic kernels for any other shape.
#include "fast_kernels.h"
#include "pdl.cuh"

#include <cuda_fp16.h>

#include <algorithm>
#include <cstdlib>

#include "plan.h"

namespace vse {

static inline int cdiv_i(int64_t a, int64_t b) { return int((a + b - 1) / b); }

__device__ __forceinline__ void load8h(const __half* p, float* v) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const float2 f = __half22float2(h[i]);
        v[2 * i] = f.x;
        v[2 * i + 1] = f.y;
    }
}
__device__ __forceinline__ void store8h(__half* p, const float* v) {
    uint4 u;
    __half2* h = reinterpret_cast<__half2*>(&u);
#pragma unroll
    for (int i = 0; i < 4; i++) h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = u;
}

// storage-generic 8-channel vectors: fp16 activations (16 bytes) or fp32 activations (32 bytes)
template <typename T> __device__ __forceinline__ void load8(const T* p, float* v);
template <> __device__ __forceinline__ void load8<__half>(const __half* p, float* v) { load8h(p, v); }
template <> __device__ __forceinline__ void load8<float>(const float* p, float* v) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <typename T> __device__ __forceinline__ void store8(T* p, const float* v);
template <> __device__ __forceinline__ void store8<__half>(__half* p, const float* v) { store8h(p, v); }
template <> __device__ __forceinline__ void store8<float>(float* p, const float* v) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
// one channel PAIR of a pixel: a packed half2 word or a float2
template <typename T> struct PairOf;
template <> struct PairOf<__half> {
    typedef unsigned raw;
    static __device__ __forceinline__ raw zero() { return 0u; }
    static __device__ __forceinline__ raw ld(const char* p) { return __ldg(reinterpret_cast<const unsigned*>(p)); }
    static __device__ __forceinline__ float2 f2(raw r) { return __half22float2(*reinterpret_cast<const __half2*>(&r)); }
    static __device__ __forceinline__ void st(char* p, float2 v) { *reinterpret_cast<__half2*>(p) = __floats2half2_rn(v.x, v.y); }
};
template <> struct PairOf<float> {
    typedef float2 raw;
    static __device__ __forceinline__ raw zero() { return make_float2(0.f, 0.f); }
    static __device__ __forceinline__ raw ld(const char* p) { return __ldg(reinterpret_cast<const float2*>(p)); }
    static __device__ __forceinline__ float2 f2(raw r) { return r; }
    static __device__ __forceinline__ void st(char* p, float2 v) { *reinterpret_cast<float2*>(p) = v; }
};

template <int ACT>
__device__ __forceinline__ float fact(float x) {
    if constexpr (ACT == ACT_RELU) return fmaxf(x, 0.f);
    else if constexpr (ACT == ACT_HSWISH) return x * fminf(fmaxf(x + 3.f, 0.f), 6.f) * (1.f / 6.f);
    else return x;
}

// packed fp32x2 arithmetic (sm_100: FFMA2 / FMUL2 / FADD2 — two IEEE fp32 operations per issued instruction, each lane
// rounded exactly like the scalar instruction) and a one-instruction 64-bit address  base + a * b  (IMAD.WIDE.U32)
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(*reinterpret_cast<unsigned long long*>(&a)),
        "l"(*reinterpret_cast<unsigned long long*>(&b)), "l"(*reinterpret_cast<unsigned long long*>(&c)));
    return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
    unsigned long long d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
    return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
    unsigned long long d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
    return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ const char* addr_mad(const char* base, unsigned a, unsigned b) {
    unsigned long long d;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(d) : "r"(a), "r"(b), "l"(reinterpret_cast<unsigned long long>(base)));
    return reinterpret_cast<const char*>(d);
}
// ------------------------------------------------------------------------------------------------
// depthwise KxK
// ------------------------------------------------------------------------------------------------
struct DwDev {
    const __half* in; __half* out; const float* w; const float* bias; const float* ps; const float* pt;
    const ImgTab* tin; const ImgTab* tout;
    int in_cs, out_cs, cvecs, cp, ph, pw;
};

template <int K, int SH, int SW, int TW, int ACT>
__global__ void __launch_bounds__(256) dwconv_fast_kernel(DwDev p) {
    pdl_wait();      // pdl.cuh: nothing below may run before the previous kernel of the stream has completed
    pdl_trigger();
    const int img = blockIdx.y;
    const ImgTab ti = p.tin[img], to = p.tout[img];
    const int strips = (to.w + TW - 1) / TW;
    const int64_t idx = int64_t(blockIdx.x) * 256 + threadIdx.x;
    if (idx >= int64_t(to.h) * strips * p.cvecs) return;
    const int cv = int(idx % p.cvecs);
    const int s = int(idx / p.cvecs);
    const int oy = s / strips, ox0 = (s - oy * strips) * TW;
    const int c0 = cv * 8;
    constexpr int IW = (TW - 1) * SW + K;
    float acc[TW][8];
#pragma unroll
    for (int t = 0; t < TW; t++)
#pragma unroll
        for (int j = 0; j < 8; j++) acc[t][j] = 0.f;
    const int iy0 = oy * SH - p.ph, ix0 = ox0 * SW - p.pw;
#pragma unroll
    for (int ky = 0; ky < K; ky++) {
        const int iy = iy0 + ky;
        if (iy < 0 || iy >= ti.h) continue;
        float w[K][8];
#pragma unroll
        for (int kx = 0; kx < K; kx++) {
            const float* wp = p.w + size_t(ky * K + kx) * p.cp + c0;
            const float4 a = __ldg(reinterpret_cast<const float4*>(wp)), b = __ldg(reinterpret_cast<const float4*>(wp + 4));
            w[kx][0] = a.x; w[kx][1] = a.y; w[kx][2] = a.z; w[kx][3] = a.w;
            w[kx][4] = b.x; w[kx][5] = b.y; w[kx][6] = b.z; w[kx][7] = b.w;
        }
        const __half* rowp = p.in + (size_t(ti.off) + size_t(iy) * ti.w) * p.in_cs + c0;
#pragma unroll
        for (int i = 0; i < IW; i++) {
            const int ix = ix0 + i;
            if (ix < 0 || ix >= ti.w) continue;
            float x[8];
            load8h(rowp + size_t(ix) * p.in_cs, x);
#pragma unroll
            for (int t = 0; t < TW; t++) {
                const int kx = i - t * SW;   // compile-time after unrolling
                if (kx >= 0 && kx < K) {
#pragma unroll
                    for (int j = 0; j < 8; j++) acc[t][j] = fmaf(x[j], w[kx][j], acc[t][j]);
                }
            }
        }
    }
    float b[8], sc[8], sh[8];
    {
        const float4 a0 = __ldg(reinterpret_cast<const float4*>(p.bias + c0)), a1 = __ldg(reinterpret_cast<const float4*>(p.bias + c0 + 4));
        b[0] = a0.x; b[1] = a0.y; b[2] = a0.z; b[3] = a0.w; b[4] = a1.x; b[5] = a1.y; b[6] = a1.z; b[7] = a1.w;
    }
    const bool post = p.ps != nullptr;
    if (post) {
        const float4 s0 = __ldg(reinterpret_cast<const float4*>(p.ps + c0)), s1 = __ldg(reinterpret_cast<const float4*>(p.ps + c0 + 4));
        const float4 t0 = __ldg(reinterpret_cast<const float4*>(p.pt + c0)), t1 = __ldg(reinterpret_cast<const float4*>(p.pt + c0 + 4));
        sc[0] = s0.x; sc[1] = s0.y; sc[2] = s0.z; sc[3] = s0.w; sc[4] = s1.x; sc[5] = s1.y; sc[6] = s1.z; sc[7] = s1.w;
        sh[0] = t0.x; sh[1] = t0.y; sh[2] = t0.z; sh[3] = t0.w; sh[4] = t1.x; sh[5] = t1.y; sh[6] = t1.z; sh[7] = t1.w;
    }
    __half* orow = p.out + (size_t(to.off) + size_t(oy) * to.w) * p.out_cs + c0;
#pragma unroll
    for (int t = 0; t < TW; t++) {
        const int ox = ox0 + t;
        if (ox >= to.w) break;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            float x = fact<ACT>(acc[t][j] + b[j]);
            if (post) x = x * sc[j] + sh[j];
            v[j] = x;
        }
        store8h(orow + size_t(ox) * p.out_cs, v);
    }
}

template <int K, int SH, int SW>
static bool dw_launch_act(const DwDev& d, const ConvArgs& a, int max_out_w_strips_pix, cudaStream_t st) {
    constexpr int TW = 4;
    // upper bound of threads per image: every image has at most max_out_pix pixels; strips <= ceil(w / TW) per row, so
    // h * strips <= (pix + h * (TW - 1)) / TW <= pix (for TW >= 1) — use pix / TW + h_max as a safe bound via max_out_pix
    dim3 grid(cdiv_i(int64_t(max_out_w_strips_pix) * d.cvecs, 256), a.n_img);
    switch (a.epi.act) {
        case ACT_NONE: pdl_launch(dwconv_fast_kernel<K, SH, SW, TW, ACT_NONE>, grid, 256, 0, st, d); return true;
        case ACT_RELU: pdl_launch(dwconv_fast_kernel<K, SH, SW, TW, ACT_RELU>, grid, 256, 0, st, d); return true;
        case ACT_HSWISH: pdl_launch(dwconv_fast_kernel<K, SH, SW, TW, ACT_HSWISH>, grid, 256, 0, st, d); return true;
        default: return false;
    }
}

bool launch_dwconv_fast(const ConvArgs& a, int max_strip_units, cudaStream_t st) {
    if (a.epi.res || a.epi.act2 != ACT_NONE || a.out_f32 || a.kh != a.kw) return false;
    if (2 * a.ph != a.kh - 1 || 2 * a.pw != a.kw - 1) return false;
    DwDev d{static_cast<const __half*>(a.in), static_cast<__half*>(a.out), a.w, a.epi.bias, a.epi.post_scale, a.epi.post_shift,
            a.tin, a.tout, a.in_cs, a.out_cs, a.cin_pad / 8, a.cin_pad, a.ph, a.pw};
    if (!d.bias) return false;
    const int key = a.kh * 100 + a.sh * 10 + a.sw;
    switch (key) {
        case 311: return dw_launch_act<3, 1, 1>(d, a, max_strip_units, st);
        case 322: return dw_launch_act<3, 2, 2>(d, a, max_strip_units, st);
        case 321: return dw_launch_act<3, 2, 1>(d, a, max_strip_units, st);
        case 312: return dw_launch_act<3, 1, 2>(d, a, max_strip_units, st);
        case 511: return dw_launch_act<5, 1, 1>(d, a, max_strip_units, st);
        case 522: return dw_launch_act<5, 2, 2>(d, a, max_strip_units, st);
        case 521: return dw_launch_act<5, 2, 1>(d, a, max_strip_units, st);
        case 512: return dw_launch_act<5, 1, 2>(d, a, max_strip_units, st);
        default: return false;
    }
}

// ------------------------------------------------------------------------------------------------
// depthwise KxK, shared-memory tiled: a CTA stages the input tile of (TH x 16 outputs) x (CBV * 8 channels) with
// cp.async (every thread issues its 16-byte pieces back to back, out-of-image pieces are zero-filled = the padding),
// waits once, and computes strips of 4 outputs from shared memory.  The strip kernel above pays a dependent global-load
// round trip per filter row; on the small, L2-resident maps of the deep stages that latency is the whole run time.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src, bool valid) {
    const int n = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

template <int K, int SH, int SW, int TH, int CBV, int ACT>
__global__ void __launch_bounds__(TH * 4 * CBV) dwconv_tile_kernel(DwDev p, int tiles_x) {
    pdl_wait();      // pdl.cuh: nothing below may run before the previous kernel of the stream has completed
    pdl_trigger();
    constexpr int TW = 16, ST = 4;                       // 16 output columns per tile, strips of 4
    constexpr int IH = (TH - 1) * SH + K, IWT = (TW - 1) * SW + K;
    constexpr int NT = TH * 4 * CBV;
    extern __shared__ __align__(16) unsigned char dw_smem[];
    __half* sin = reinterpret_cast<__half*>(dw_smem);                                   // [IH][IWT][CBV * 8]
    float* sw = reinterpret_cast<float*>(dw_smem + size_t(IH) * IWT * CBV * 16);        // [K*K][CBV * 8]
    const int img = blockIdx.y;
    const ImgTab ti = p.tin[img], to = p.tout[img];
    const int ty0 = (blockIdx.x / tiles_x) * TH, tx0 = (blockIdx.x % tiles_x) * TW;
    if (ty0 >= to.h || tx0 >= to.w) return;
    const int cv0 = blockIdx.z * CBV;
    const int ncv = min(CBV, p.cvecs - cv0);
    const int iy0 = ty0 * SH - p.ph, ix0 = tx0 * SW - p.pw;
    const uint32_t sin_addr = (uint32_t)__cvta_generic_to_shared(sin);
    for (int i = threadIdx.x; i < IH * IWT * CBV; i += NT) {
        const int cv = i % CBV, pix = i / CBV;
        const int r = pix / IWT, c = pix - r * IWT;
        const int iy = iy0 + r, ix = ix0 + c;
        const bool ok = cv < ncv && iy >= 0 && iy < ti.h && ix >= 0 && ix < ti.w;
        const __half* src = ok ? p.in + (size_t(ti.off) + size_t(iy) * ti.w + ix) * p.in_cs + (cv0 + cv) * 8 : p.in;
        cp_async_16(sin_addr + uint32_t(i) * 16, src, ok);
    }
    for (int i = threadIdx.x; i < K * K * CBV * 8; i += NT) {
        const int c = i % (CBV * 8), t = i / (CBV * 8);
        sw[i] = (cv0 * 8 + c < p.cp) ? p.w[size_t(t) * p.cp + cv0 * 8 + c] : 0.f;
    }
    cp_async_wait_all();
    __syncthreads();
    const int cv = threadIdx.x % CBV;
    const int st = (threadIdx.x / CBV) % 4, oyl = threadIdx.x / (CBV * 4);
    const int oy = ty0 + oyl, ox0 = tx0 + st * ST;
    if (cv >= ncv || oy >= to.h || ox0 >= to.w) return;
    constexpr int IW = (ST - 1) * SW + K;
    float acc[ST][8];
#pragma unroll
    for (int t = 0; t < ST; t++)
#pragma unroll
        for (int j = 0; j < 8; j++) acc[t][j] = 0.f;
#pragma unroll
    for (int ky = 0; ky < K; ky++) {
        float w[K][8];
#pragma unroll
        for (int kx = 0; kx < K; kx++) {
            const float4* wp = reinterpret_cast<const float4*>(sw + (ky * K + kx) * (CBV * 8) + cv * 8);
            const float4 a = wp[0], b = wp[1];
            w[kx][0] = a.x; w[kx][1] = a.y; w[kx][2] = a.z; w[kx][3] = a.w;
            w[kx][4] = b.x; w[kx][5] = b.y; w[kx][6] = b.z; w[kx][7] = b.w;
        }
        const __half* rowp = sin + (size_t(oyl * SH + ky) * IWT + st * ST * SW) * (CBV * 8) + cv * 8;
#pragma unroll
        for (int i = 0; i < IW; i++) {
            float x[8];
            load8h(rowp + size_t(i) * (CBV * 8), x);
#pragma unroll
            for (int t = 0; t < ST; t++) {
                const int kx = i - t * SW;   // compile-time after unrolling
                if (kx >= 0 && kx < K) {
#pragma unroll
                    for (int j = 0; j < 8; j++) acc[t][j] = fmaf(x[j], w[kx][j], acc[t][j]);
                }
            }
        }
    }
    const int c0 = (cv0 + cv) * 8;
    float b[8], sc[8], sh[8];
    {
        const float4 a0 = __ldg(reinterpret_cast<const float4*>(p.bias + c0)), a1 = __ldg(reinterpret_cast<const float4*>(p.bias + c0 + 4));
        b[0] = a0.x; b[1] = a0.y; b[2] = a0.z; b[3] = a0.w; b[4] = a1.x; b[5] = a1.y; b[6] = a1.z; b[7] = a1.w;
    }
    const bool post = p.ps != nullptr;
    if (post) {
        const float4 s0 = __ldg(reinterpret_cast<const float4*>(p.ps + c0)), s1 = __ldg(reinterpret_cast<const float4*>(p.ps + c0 + 4));
        const float4 t0 = __ldg(reinterpret_cast<const float4*>(p.pt + c0)), t1 = __ldg(reinterpret_cast<const float4*>(p.pt + c0 + 4));
        sc[0] = s0.x; sc[1] = s0.y; sc[2] = s0.z; sc[3] = s0.w; sc[4] = s1.x; sc[5] = s1.y; sc[6] = s1.z; sc[7] = s1.w;
        sh[0] = t0.x; sh[1] = t0.y; sh[2] = t0.z; sh[3] = t0.w; sh[4] = t1.x; sh[5] = t1.y; sh[6] = t1.z; sh[7] = t1.w;
    }
    __half* orow = p.out + (size_t(to.off) + size_t(oy) * to.w) * p.out_cs + c0;
#pragma unroll
    for (int t = 0; t < ST; t++) {
        const int ox = ox0 + t;
        if (ox >= to.w) break;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            float x = fact<ACT>(acc[t][j] + b[j]);
            if (post) x = x * sc[j] + sh[j];
            v[j] = x;
        }
        store8h(orow + size_t(ox) * p.out_cs, v);
    }
}

template <int K, int SH, int SW, int TH, int CBV>
static bool dw_tile_launch(const DwDev& d, const ConvArgs& a, int max_h, int max_w, cudaStream_t st) {
    constexpr int IH = (TH - 1) * SH + K, IWT = 15 * SW + K;
    const size_t smem = size_t(IH) * IWT * CBV * 16 + size_t(K) * K * CBV * 8 * sizeof(float);
    const int tiles_x = (max_w + 15) / 16, tiles_y = (max_h + TH - 1) / TH;
    dim3 grid(tiles_x * tiles_y, a.n_img, (d.cvecs + CBV - 1) / CBV);
    auto go = [&](auto kern) {
        // > 48 KB of dynamic shared memory needs the opt-in (per device and per kernel instantiation; the call is cheap)
        if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
        pdl_launch(kern, grid, TH * 4 * CBV, smem, st, d, tiles_x);
    };
    switch (a.epi.act) {
        case ACT_NONE: go(dwconv_tile_kernel<K, SH, SW, TH, CBV, ACT_NONE>); return true;
        case ACT_RELU: go(dwconv_tile_kernel<K, SH, SW, TH, CBV, ACT_RELU>); return true;
        case ACT_HSWISH: go(dwconv_tile_kernel<K, SH, SW, TH, CBV, ACT_HSWISH>); return true;
        default: return false;
    }
}

bool launch_dwconv_tiled(const ConvArgs& a, int max_out_h, int max_out_w, cudaStream_t st) {
    if (a.epi.res || a.epi.act2 != ACT_NONE || a.out_f32 || a.kh != a.kw) return false;
    if (2 * a.ph != a.kh - 1 || 2 * a.pw != a.kw - 1) return false;
    DwDev d{static_cast<const __half*>(a.in), static_cast<__half*>(a.out), a.w, a.epi.bias, a.epi.post_scale, a.epi.post_shift,
            a.tin, a.tout, a.in_cs, a.out_cs, a.cin_pad / 8, a.cin_pad, a.ph, a.pw};
    if (!d.bias) return false;
    const int key = a.kh * 100 + a.sh * 10 + a.sw;
    // channel block per CTA: 8 vectors (64 channels, 128 B per pixel), 4 for 32-channel layers, 2 for 16/24 channels
#define VSE_DW_TILE(KK, SHH, SWW, THW, THN)                                                                              \
    (d.cvecs >= 5   ? (max_out_h <= 4 ? dw_tile_launch<KK, SHH, SWW, 4, 8>(d, a, max_out_h, max_out_w, st)               \
                                      : dw_tile_launch<KK, SHH, SWW, THW, 8>(d, a, max_out_h, max_out_w, st))            \
     : d.cvecs == 4 ? dw_tile_launch<KK, SHH, SWW, THN, 4>(d, a, max_out_h, max_out_w, st)                               \
                    : dw_tile_launch<KK, SHH, SWW, THN, 2>(d, a, max_out_h, max_out_w, st))
    switch (key) {
        case 311: return VSE_DW_TILE(3, 1, 1, 8, 8);
        case 322: return VSE_DW_TILE(3, 2, 2, 4, 8);
        case 321: return VSE_DW_TILE(3, 2, 1, 4, 8);
        case 312: return VSE_DW_TILE(3, 1, 2, 8, 8);
        case 511: return VSE_DW_TILE(5, 1, 1, 8, 8);
        case 522: return VSE_DW_TILE(5, 2, 2, 4, 8);
        case 521: return VSE_DW_TILE(5, 2, 1, 4, 8);
        case 512: return VSE_DW_TILE(5, 1, 2, 8, 8);
        default: return false;
    }
#undef VSE_DW_TILE
}

// ------------------------------------------------------------------------------------------------
// depthwise KxK, register tiled: one thread = TH x TW output pixels of ONE channel pair (half2).  The whole filter of the
// pair (K*K float2) and the output tile live in registers; the input patch is read row by row straight from global
// memory (4-byte loads: a warp reads 128 contiguous bytes = 64 channels of one pixel), converted once, and every
// converted value feeds up to K*K FMAs — no shared memory, no per-strip filter reloads.  The kernels above issue one
// 16-byte shared/global load per ~10 FMAs plus the fp16->fp32 conversions of 8 channels per strip column and are
// instruction-issue bound on the deep (>= 96 channel, 5x5) layers; here ~80 % of the issued instructions are FMAs.
// Accumulation order per output (ky, then kx ascending, fp32 fmaf from 0) is the same as in the other depthwise kernels,
// so the results are bit-identical to theirs.
// ------------------------------------------------------------------------------------------------
template <int ACT>
__device__ __forceinline__ float2 fact2(float2 v) {
    if constexpr (ACT == ACT_RELU) return make_float2(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f));
    else if constexpr (ACT == ACT_HSWISH) {   // x * min(max(x + 3, 0), 6) * (1/6), evaluated left to right as fact<>()
        const float2 t = fadd2(v, make_float2(3.f, 3.f));
        const float2 c = make_float2(fminf(fmaxf(t.x, 0.f), 6.f), fminf(fmaxf(t.y, 0.f), 6.f));
        return fmul2(fmul2(v, c), make_float2(1.f / 6.f, 1.f / 6.f));
    } else return v;
}

template <typename T, int K, int SH, int SW, int TH, int TW, int ACT>
// resident CTAs per SM (register caps 128 / 168 / 255): the kernel is latency bound (ncu: 0.46 eligible warps per scheduler at 8 warps
// per SM, profiles/r02_ncu_summary.txt), so every variant takes the highest occupancy ptxas reaches without spilling
__global__ void __launch_bounds__(128, K == 3 ? (SH * SW < 4 ? 4 : 3) : (sizeof(T) == 2 || SH * SW == 1) ? 3 : 2) dwconv_reg_kernel(DwDev p, int tiles_x, int tiles_y) {
    typedef PairOf<T> PR;
    constexpr int IH = (TH - 1) * SH + K, IW = (TW - 1) * SW + K;
    const int img = blockIdx.y;
    const int cp2 = p.cvecs * 4;                          // channel pairs per pixel
    const unsigned idx = blockIdx.x * 128u + threadIdx.x;   // the launcher keeps tiles * pairs below 2^31
    const unsigned tile = idx / unsigned(cp2);
    const int pair = int(idx - tile * unsigned(cp2));
    if (tile >= unsigned(tiles_x * tiles_y)) return;
    // filter, bias and post-affine of this channel pair are static weights: loaded BEFORE the dependency wait, so under
    // programmatic dependent launch (pdl.cuh) this round trip overlaps the tail of the previous step's kernel
    float2 w[K * K];
    {
        const char* wp = reinterpret_cast<const char*>(p.w + pair * 2);
        const unsigned wstep = unsigned(p.cp) * 4u;
#pragma unroll
        for (int t = 0; t < K * K; t++) w[t] = __ldg(reinterpret_cast<const float2*>(addr_mad(wp, unsigned(t), wstep)));
    }
    const float2 bias = __ldg(reinterpret_cast<const float2*>(p.bias + pair * 2));
    const bool post = p.ps != nullptr;
    float2 sc = make_float2(1.f, 1.f), sh = make_float2(0.f, 0.f);
    if (post) {
        sc = __ldg(reinterpret_cast<const float2*>(p.ps + pair * 2));
        sh = __ldg(reinterpret_cast<const float2*>(p.pt + pair * 2));
    }
    pdl_wait();      // image tables and activations come from earlier kernels of the stream
    pdl_trigger();
    const ImgTab ti = p.tin[img], to = p.tout[img];
    const int ty = int(tile / unsigned(tiles_x)), tx = int(tile) - ty * tiles_x;
    const int oy0 = ty * TH, ox0 = tx * TW;
    if (oy0 >= to.h || ox0 >= to.w) return;
    const int c0 = pair * 2;
    const int iy0 = oy0 * SH - p.ph, ix0 = ox0 * SW - p.pw;
    const char* base = reinterpret_cast<const char*>(reinterpret_cast<const T*>(p.in) + size_t(ti.off) * p.in_cs + c0);
    const unsigned cstep = unsigned(p.in_cs) * unsigned(sizeof(T));        // bytes per pixel
    // The whole input patch goes into registers first (packed half2, one register per tap): all IH*IW loads of a thread
    // are in flight together, so a thread pays ONE memory round trip instead of one per patch row — on these small,
    // L2-resident maps the dependent round trips were half the run time of the row-by-row variants.  Every address is one
    // IMAD.WIDE.  Out-of-image taps are predicated loads that yield 0 (the FMAs then add +0).
    typename PR::raw xin[IH][IW];
    const bool interior = iy0 >= 0 && iy0 + IH <= ti.h && ix0 >= 0 && ix0 + IW <= ti.w;
    if (interior) {
        const char* pp = addr_mad(base, unsigned(iy0 * ti.w + ix0), cstep);
        const unsigned rstep = unsigned(ti.w) * cstep;    // bytes per image row (< 2^32: checked by the launcher)
#pragma unroll
        for (int r = 0; r < IH; r++) {
            const char* rowp = addr_mad(pp, unsigned(r), rstep);
#pragma unroll
            for (int i = 0; i < IW; i++) xin[r][i] = PR::ld(addr_mad(rowp, unsigned(i), cstep));
        }
    } else {
#pragma unroll
        for (int r = 0; r < IH; r++) {
            const int iy = iy0 + r;
            const bool row_ok = iy >= 0 && iy < ti.h;
            const int roff = (row_ok ? iy : 0) * ti.w;   // in-image pixel offsets fit 32 bits (engine: pixel index < 2^31)
#pragma unroll
            for (int i = 0; i < IW; i++) {
                const int ix = ix0 + i;
                typename PR::raw v = PR::zero();
                if (row_ok && ix >= 0 && ix < ti.w) v = PR::ld(addr_mad(base, unsigned(roff + ix), cstep));
                xin[r][i] = v;
            }
        }
    }
    float2 acc[TH][TW];
#pragma unroll
    for (int a = 0; a < TH; a++)
#pragma unroll
        for (int b = 0; b < TW; b++) acc[a][b] = make_float2(0.f, 0.f);
#pragma unroll
    for (int r = 0; r < IH; r++) {
        float2 x[IW];
#pragma unroll
        for (int i = 0; i < IW; i++) x[i] = PR::f2(xin[r][i]);
#pragma unroll
        for (int a = 0; a < TH; a++) {
            const int ky = r - a * SH;   // compile-time after unrolling
            if (ky < 0 || ky >= K) continue;
#pragma unroll
            for (int i = 0; i < IW; i++) {
#pragma unroll
                for (int b = 0; b < TW; b++) {
                    const int kx = i - b * SW;
                    if (kx >= 0 && kx < K) acc[a][b] = ffma2(x[i], w[ky * K + kx], acc[a][b]);
                }
            }
        }
    }
    char* obase = reinterpret_cast<char*>(reinterpret_cast<T*>(p.out) + (size_t(to.off) + size_t(oy0) * to.w + ox0) * p.out_cs + c0);
    const unsigned ocstep = unsigned(p.out_cs) * unsigned(sizeof(T)), orstep = unsigned(to.w) * ocstep;
    const bool full = oy0 + TH <= to.h && ox0 + TW <= to.w;
#pragma unroll
    for (int a = 0; a < TH; a++) {
        if (!full && oy0 + a >= to.h) break;
        char* orow = const_cast<char*>(addr_mad(obase, unsigned(a), orstep));
#pragma unroll
        for (int b = 0; b < TW; b++) {
            if (!full && ox0 + b >= to.w) break;
            float2 v = fact2<ACT>(fadd2(acc[a][b], bias));
            if (post) v = ffma2(v, sc, sh);
            PR::st(const_cast<char*>(addr_mad(orow, unsigned(b), ocstep)), v);
        }
    }
}

template <typename T, int K, int SH, int SW, int TH, int TW>
static bool dw_reg_launch(const DwDev& d, const ConvArgs& a, int max_h, int max_w, cudaStream_t st) {
    const int tiles_x = (max_w + TW - 1) / TW, tiles_y = (max_h + TH - 1) / TH;
    if (int64_t(tiles_x) * tiles_y * d.cvecs * 4 > 0x7fffff00LL) return false;
    if (int64_t(max_w) * std::max(SW, 1) * std::max(a.in_cs, a.out_cs) * int(sizeof(T)) * 2 > 0x7fffffffLL) return false;   // 32-bit row strides
    dim3 grid(cdiv_i(int64_t(tiles_x) * tiles_y * d.cvecs * 4, 128), a.n_img);
    switch (a.epi.act) {
        case ACT_NONE: pdl_launch(dwconv_reg_kernel<T, K, SH, SW, TH, TW, ACT_NONE>, grid, 128, 0, st, d, tiles_x, tiles_y); return true;
        case ACT_RELU: pdl_launch(dwconv_reg_kernel<T, K, SH, SW, TH, TW, ACT_RELU>, grid, 128, 0, st, d, tiles_x, tiles_y); return true;
        case ACT_HSWISH: pdl_launch(dwconv_reg_kernel<T, K, SH, SW, TH, TW, ACT_HSWISH>, grid, 128, 0, st, d, tiles_x, tiles_y); return true;
        default: return false;
    }
}

bool launch_dwconv_reg(const ConvArgs& a, int max_out_h, int max_out_w, cudaStream_t st, int prec) {
    if (a.epi.res || a.epi.act2 != ACT_NONE || a.out_f32 || a.kh != a.kw) return false;
    if (2 * a.ph != a.kh - 1 || 2 * a.pw != a.kw - 1) return false;
    DwDev d{static_cast<const __half*>(a.in), static_cast<__half*>(a.out), a.w, a.epi.bias, a.epi.post_scale, a.epi.post_shift,
            a.tin, a.tout, a.in_cs, a.out_cs, a.cin_pad / 8, a.cin_pad, a.ph, a.pw};
    if (!d.bias) return false;
    if (prec == 1) {
        // fp32 activations: a channel pair is a float2 (two registers per tap), so the tiles are smaller
        switch (a.kh * 100 + a.sh * 10 + a.sw) {
            case 311: return dw_reg_launch<float, 3, 1, 1, 4, 4>(d, a, max_out_h, max_out_w, st);
            case 322: return dw_reg_launch<float, 3, 2, 2, 3, 4>(d, a, max_out_h, max_out_w, st);
            case 321: return dw_reg_launch<float, 3, 2, 1, 3, 4>(d, a, max_out_h, max_out_w, st);
            case 312: return dw_reg_launch<float, 3, 1, 2, 4, 3>(d, a, max_out_h, max_out_w, st);
            case 511: return dw_reg_launch<float, 5, 1, 1, 2, 4>(d, a, max_out_h, max_out_w, st);
            case 522: return dw_reg_launch<float, 5, 2, 2, 2, 3>(d, a, max_out_h, max_out_w, st);
            case 521: return dw_reg_launch<float, 5, 2, 1, 2, 4>(d, a, max_out_h, max_out_w, st);
            case 512: return dw_reg_launch<float, 5, 1, 2, 2, 3>(d, a, max_out_h, max_out_w, st);
            default: return false;
        }
    }
    // rows per tile: 3 or 4, whichever pads the tallest image less (ties: 4)
    const bool th3 = ((max_out_h + 2) / 3) * 3 < ((max_out_h + 3) / 4) * 4;
    const int key = a.kh * 100 + a.sh * 10 + a.sw;
#define VSE_DW_REG(KK, SHH, SWW) \
    (th3 ? dw_reg_launch<__half, KK, SHH, SWW, 3, 4>(d, a, max_out_h, max_out_w, st) : dw_reg_launch<__half, KK, SHH, SWW, 4, 4>(d, a, max_out_h, max_out_w, st))
    switch (key) {
        case 311: return VSE_DW_REG(3, 1, 1);
        case 322: return VSE_DW_REG(3, 2, 2);
        case 321: return VSE_DW_REG(3, 2, 1);
        case 312: return VSE_DW_REG(3, 1, 2);
        case 511: return VSE_DW_REG(5, 1, 1);
        case 522: return dw_reg_launch<__half, 5, 2, 2, 2, 4>(d, a, max_out_h, max_out_w, st);   // 7 x 11 patch: the 4-row tile (11 x 11) would spill
        case 521: return VSE_DW_REG(5, 2, 1);
        case 512: return VSE_DW_REG(5, 1, 2);
        default: return false;
    }
#undef VSE_DW_REG
}

// ------------------------------------------------------------------------------------------------
// stem: 3x3 stride 2 pad 1, uint8 BGRX -> 16 channels per blockIdx.z (the server detector's stem has 64: four channel blocks
// re-read the small u8 input from L2), (x * nscale + nshift) fused into the load
// ------------------------------------------------------------------------------------------------
struct StemDev {
    const unsigned char* in; __half* out; const float* w; const float* bias; const ImgTab* tin; const ImgTab* tout;
    int out_cs, w_ci, w_co, act;
    float nscale[3], nshift[3];
};

template <typename T, int ACT>
__global__ void __launch_bounds__(128) stem_fast_kernel(StemDev p) {
    pdl_wait();      // pdl.cuh: nothing below may run before the previous kernel of the stream has completed
    pdl_trigger();
    __shared__ __align__(16) float sw[27 * 16];   // [(ky*3+kx)*3 + ci][co]
    __shared__ float sb[16];
    const int co_base = blockIdx.z * 16;
    for (int i = threadIdx.x; i < 27 * 16; i += blockDim.x) {
        const int co = i & 15, r = i >> 4, ci = r % 3, tap = r / 3;
        sw[i] = p.w[(size_t(tap) * p.w_ci + ci) * p.w_co + co_base + co];
    }
    if (threadIdx.x < 16) sb[threadIdx.x] = p.bias[co_base + threadIdx.x];
    __syncthreads();
    const int img = blockIdx.y;
    const ImgTab ti = p.tin[img], to = p.tout[img];
    const int pairs = (to.w + 1) >> 1;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= to.h * pairs) return;
    const int oy = idx / pairs, ox0 = (idx - oy * pairs) * 2;
    // accumulators as channel pairs: one FFMA2 (fma.rn.f32x2) per two output channels, each lane rounded like the scalar FMA
    float2 acc[2][8];
#pragma unroll
    for (int t = 0; t < 2; t++)
#pragma unroll
        for (int j = 0; j < 8; j++) acc[t][j] = make_float2(sb[2 * j], sb[2 * j + 1]);
    const int iy0 = oy * 2 - 1, ix0 = ox0 * 2 - 1;
#pragma unroll
    for (int ky = 0; ky < 3; ky++) {
        const int iy = iy0 + ky;
        if (iy < 0 || iy >= ti.h) continue;
        const uchar4* rowp = reinterpret_cast<const uchar4*>(p.in) + size_t(ti.off) + size_t(iy) * ti.w;
        float px[5][3];
#pragma unroll
        for (int i = 0; i < 5; i++) {
            const int ix = ix0 + i;
            if (ix >= 0 && ix < ti.w && ix < ti.vw) {
                const uchar4 u = __ldg(rowp + ix);
                px[i][0] = float(u.x) * p.nscale[0] + p.nshift[0];
                px[i][1] = float(u.y) * p.nscale[1] + p.nshift[1];
                px[i][2] = float(u.z) * p.nscale[2] + p.nshift[2];
            } else {
                px[i][0] = px[i][1] = px[i][2] = 0.f;
            }
        }
#pragma unroll
        for (int kx = 0; kx < 3; kx++) {
#pragma unroll
            for (int ci = 0; ci < 3; ci++) {
                const float4* wp = reinterpret_cast<const float4*>(sw + ((ky * 3 + kx) * 3 + ci) * 16);
                float2 w[8];
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const float4 f = wp[q];
                    w[2 * q] = make_float2(f.x, f.y);
                    w[2 * q + 1] = make_float2(f.z, f.w);
                }
#pragma unroll
                for (int t = 0; t < 2; t++) {
                    const float x = px[kx + 2 * t][ci];
                    const float2 xx = make_float2(x, x);
#pragma unroll
                    for (int j = 0; j < 8; j++) acc[t][j] = ffma2(xx, w[j], acc[t][j]);
                }
            }
        }
    }
#pragma unroll
    for (int t = 0; t < 2; t++) {
        const int ox = ox0 + t;
        if (ox >= to.w) break;
        float v[16];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            v[2 * j] = fact<ACT>(acc[t][j].x);
            v[2 * j + 1] = fact<ACT>(acc[t][j].y);
        }
        T* o = reinterpret_cast<T*>(p.out) + (size_t(to.off) + size_t(oy) * to.w + ox) * p.out_cs + co_base;
        store8<T>(o, v);
        store8<T>(o + 8, v + 8);
    }
}

bool launch_stem_fast(const ConvArgs& a, int cout, int max_out_pix_pairs, cudaStream_t st, int prec) {
    if (!a.in_u8 || a.kh != 3 || a.kw != 3 || a.sh != 2 || a.sw != 2 || a.ph != 1 || a.pw != 1 || cout % 16 || cout > 256 || a.out_f32) return false;
    if (a.epi.res || a.epi.post_scale || a.epi.act2 != ACT_NONE || !a.epi.bias || a.out_cs < cout) return false;
    StemDev d{static_cast<const unsigned char*>(a.in), static_cast<__half*>(a.out), a.w, a.epi.bias, a.tin, a.tout,
              a.out_cs, a.w_ci, a.w_co, a.epi.act, {a.nscale[0], a.nscale[1], a.nscale[2]}, {a.nshift[0], a.nshift[1], a.nshift[2]}};
    dim3 grid(cdiv_i(max_out_pix_pairs, 128), a.n_img, cout / 16);
    if (prec == 1) {
        switch (a.epi.act) {
            case ACT_NONE: pdl_launch(stem_fast_kernel<float, ACT_NONE>, grid, 128, 0, st, d); return true;
            case ACT_RELU: pdl_launch(stem_fast_kernel<float, ACT_RELU>, grid, 128, 0, st, d); return true;
            case ACT_HSWISH: pdl_launch(stem_fast_kernel<float, ACT_HSWISH>, grid, 128, 0, st, d); return true;
            default: return false;
        }
    }
    switch (a.epi.act) {
        case ACT_NONE: pdl_launch(stem_fast_kernel<__half, ACT_NONE>, grid, 128, 0, st, d); return true;
        case ACT_RELU: pdl_launch(stem_fast_kernel<__half, ACT_RELU>, grid, 128, 0, st, d); return true;
        case ACT_HSWISH: pdl_launch(stem_fast_kernel<__half, ACT_HSWISH>, grid, 128, 0, st, d); return true;
        default: return false;
    }
}

// ------------------------------------------------------------------------------------------------
// DB head: conv_transpose2x2s2(C->C)+ReLU -> conv_transpose2x2s2(C->1)+sigmoid, fused.  fp32 probability map out.
// w1: [pos][C][C] (cout-major rows of cin), w2: [pos][8][C] (row 0 is the single output channel)
// ------------------------------------------------------------------------------------------------
struct HeadDev {
    const __half* in; float* out; const float* w1; const float* b1; const float* w2; const float* b2;
    const ImgTab* tin; const ImgTab* tout;
    int in_cs;
};

template <typename T, int C>
__global__ void __launch_bounds__(128, 4) db_head_fused_kernel(HeadDev p) {
    pdl_wait();      // pdl.cuh: nothing below may run before the previous kernel of the stream has completed
    pdl_trigger();
    __shared__ __align__(16) float sw1[4 * C * C];
    __shared__ __align__(16) float sw2[4 * C];
    __shared__ float sb1[C];
    for (int i = threadIdx.x; i < 4 * C * C; i += blockDim.x) sw1[i] = p.w1[i];
    for (int i = threadIdx.x; i < 4 * C; i += blockDim.x) sw2[i] = p.w2[size_t(i / C) * 8 * C + (i % C)];
    if (threadIdx.x < C) sb1[threadIdx.x] = p.b1[threadIdx.x];
    __syncthreads();
    const float b2 = p.b2[0];
    const int img = blockIdx.y;
    const ImgTab ti = p.tin[img], to = p.tout[img];
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= ti.h * ti.w) return;
    const int iy = idx / ti.w, ix = idx - iy * ti.w;
    float x[C];
    const T* ip = reinterpret_cast<const T*>(p.in) + (size_t(ti.off) + idx) * p.in_cs;
#pragma unroll
    for (int c = 0; c < C; c += 8) load8<T>(ip + c, x + c);
    float o[4][4];   // [row 2*dy1+dy2][col 2*dx1+dx2]
#pragma unroll
    for (int pos1 = 0; pos1 < 4; pos1++) {
        float mid[C];
#pragma unroll
        for (int co = 0; co < C; co++) {
            // packed fp32x2 FMAs (FFMA2) over input-channel pairs: even channels accumulate in .x (starting from the
            // bias), odd channels in .y, summed at the end — half the FMA issue slots of the scalar chain
            const float4* wr = reinterpret_cast<const float4*>(sw1 + (pos1 * C + co) * C);
            float2 m = make_float2(sb1[co], 0.f);
#pragma unroll
            for (int q = 0; q < C / 4; q++) {
                const float4 w = wr[q];
                m = ffma2(make_float2(x[4 * q], x[4 * q + 1]), make_float2(w.x, w.y), m);
                m = ffma2(make_float2(x[4 * q + 2], x[4 * q + 3]), make_float2(w.z, w.w), m);
            }
            mid[co] = fmaxf(m.x + m.y, 0.f);
        }
#pragma unroll
        for (int pos2 = 0; pos2 < 4; pos2++) {
            const float4* wr = reinterpret_cast<const float4*>(sw2 + pos2 * C);
            float s = b2;
#pragma unroll
            for (int q = 0; q < C / 4; q++) {
                const float4 w = wr[q];
                s = fmaf(mid[4 * q], w.x, s); s = fmaf(mid[4 * q + 1], w.y, s);
                s = fmaf(mid[4 * q + 2], w.z, s); s = fmaf(mid[4 * q + 3], w.w, s);
            }
            o[2 * (pos1 >> 1) + (pos2 >> 1)][2 * (pos1 & 1) + (pos2 & 1)] = 1.f / (1.f + __expf(-s));
        }
    }
    float* op = p.out + size_t(to.off) + size_t(4 * iy) * to.w + 4 * ix;
#pragma unroll
    for (int r = 0; r < 4; r++)
        *reinterpret_cast<float4*>(op + size_t(r) * to.w) = make_float4(o[r][0], o[r][1], o[r][2], o[r][3]);
}

bool launch_db_head_fused(const void* in, int in_cs, int c, const float* w1, const float* b1, const float* w2, const float* b2,
                          float* out, int out_cs, const ImgTab* tin, const ImgTab* tout, int n_img, int max_in_pix,
                          cudaStream_t st, int prec) {
    if (c != 24 || out_cs != 1 || (in_cs & 7) || !b1 || !b2) return false;
    if (reinterpret_cast<uintptr_t>(out) & 15) return false;
    HeadDev d{static_cast<const __half*>(in), out, w1, b1, w2, b2, tin, tout, in_cs};
    dim3 grid(cdiv_i(max_in_pix, 128), n_img);
    if (prec == 1) pdl_launch(db_head_fused_kernel<float, 24>, grid, 128, 0, st, d);
    else pdl_launch(db_head_fused_kernel<__half, 24>, grid, 128, 0, st, d);
    return true;
}

// ------------------------------------------------------------------------------------------------
// squeeze-excite gate: (partial channel sums of the global average pool) -> mean -> FC(C->Cm)+act1 -> FC(Cm->C)+act2,
// one CTA per image.  Replaces gpool_final + 2 x veclin (3 launches of a latency-bound chain) by one.
// ------------------------------------------------------------------------------------------------
struct SeDev {
    const float* partial; int splits, c_pad, c, cm;
    const ImgTab* tin;
    const float* w1; const float* b1; const float* w2; const float* b2;
    int act1, act2; float slope1, offset1, slope2, offset2;
    float* out;
    // optional pre-stage: the pooled tensor is y = W x + b of a 1x1 convolution, pooled from its INPUT x
    // (mean(y) = W mean(x) + b): partial holds sums of x (cx channels, cx_pad stride), pre_w is [cx][pre_ld] (co contiguous)
    const float* pre_w; const float* pre_b; int cx, cx_pad, pre_ld;
};

__device__ __forceinline__ float act_rt(float x, int act, float slope, float offset) {
    switch (act) {
        case ACT_RELU: return fmaxf(x, 0.f);
        case ACT_HSWISH: return x * fminf(fmaxf(x + 3.f, 0.f), 6.f) * (1.f / 6.f);
        case ACT_HSIGMOID: return fminf(fmaxf(x * slope + offset, 0.f), 1.f);
        case ACT_SWISH: return x / (1.f + __expf(-x));
        case ACT_SIGMOID: return 1.f / (1.f + __expf(-x));
        case ACT_RELU6: return fminf(fmaxf(x, 0.f), 6.f);
        default: return x;
    }
}

// One small fully-connected layer inside a CTA: out[co] = act(bias[co] + sum_i w[i * ld + co] * in[i]).  The weights are
// input-major (consecutive threads own consecutive outputs: coalesced reads, no shuffles) and the K range is split over
// blockDim / n_out thread groups whose partial sums meet in shared memory.  in / part live in shared memory.
__device__ __forceinline__ void cta_fc(const float* in, int n_in, const float* __restrict__ w, int ld, const float* __restrict__ bias,
                                       int n_out, float* part, float* out, int act, float slope, float offset) {
    const int nt = blockDim.x;
    const int groups = n_out >= nt ? 1 : nt / n_out;
    for (int co0 = 0; co0 < n_out; co0 += nt) {          // n_out > blockDim: several passes (groups == 1)
        const int co = co0 + int(threadIdx.x) % (n_out >= nt ? nt : n_out), g = n_out >= nt ? 0 : int(threadIdx.x) / n_out;
        float s = 0.f;
        if (g < groups && co < n_out) {
            const int chunk = (n_in + groups - 1) / groups;
            const int i0 = g * chunk, i1 = min(n_in, i0 + chunk);
#pragma unroll 8
            for (int i = i0; i < i1; i++) s = fmaf(__ldg(w + size_t(i) * ld + co), in[i], s);
        }
        part[threadIdx.x] = s;
        __syncthreads();
        if (g == 0 && co < n_out) {
            float t = bias[co];
            for (int q = 0; q < groups; q++) t += part[q * n_out + (co - co0)];
            out[co] = act_rt(t, act, slope, offset);
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(1024) se_gate_kernel(SeDev p) {
    pdl_wait();      // pdl.cuh: nothing below may run before the previous kernel of the stream has completed
    pdl_trigger();
    extern __shared__ float sm[];
    float* mean = sm;                 // [c]
    float* hid = sm + p.c;            // [cm]
    float* part = hid + p.cm;         // [blockDim.x]
    const int img = blockIdx.x;
    const float inv = 1.f / float(p.tin[img].h * p.tin[img].w);
    if (p.pre_w) {
        float* mean_x = part + blockDim.x;    // [cx]
        for (int c = threadIdx.x; c < p.cx; c += blockDim.x) {
            float s = 0.f;
            for (int k = 0; k < p.splits; k++) s += p.partial[(size_t(img) * p.splits + k) * p.cx_pad + c];
            mean_x[c] = s * inv;
        }
        __syncthreads();
        cta_fc(mean_x, p.cx, p.pre_w, p.pre_ld, p.pre_b, p.c, part, mean, ACT_NONE, 0.f, 0.f);   // mean(W x + b) = W mean(x) + b
    } else {
        for (int c = threadIdx.x; c < p.c; c += blockDim.x) {
            float s = 0.f;
            for (int k = 0; k < p.splits; k++) s += p.partial[(size_t(img) * p.splits + k) * p.c_pad + c];
            mean[c] = s * inv;
        }
        __syncthreads();
    }
    cta_fc(mean, p.c, p.w1, p.cm, p.b1, p.cm, part, hid, p.act1, p.slope1, p.offset1);
    cta_fc(hid, p.cm, p.w2, p.c, p.b2, p.c, part, p.out + size_t(img) * p.c, p.act2, p.slope2, p.offset2);
}

void launch_se_gate(const float* partial, int splits, int c_pad, int c, int cm, const ImgTab* tin, const float* w1,
                    const float* b1, int act1, float slope1, float offset1, const float* w2, const float* b2, int act2,
                    float slope2, float offset2, float* out, int n_img, cudaStream_t st, const float* pre_w, const float* pre_b,
                    int cx, int cx_pad, int pre_ld) {
    SeDev d{partial, splits, c_pad, c, cm, tin, w1, b1, w2, b2, act1, act2, slope1, offset1, slope2, offset2, out,
            pre_w, pre_b, cx, cx_pad, pre_ld};
    const int threads = 1024;   // cm <= 512 <= threads is checked by the caller; more threads = shorter dependent-load chains per FC
    pdl_launch(se_gate_kernel, n_img, threads, size_t(c + cm + threads + (pre_w ? cx : 0)) * sizeof(float), st, d);
}

// ------------------------------------------------------------------------------------------------
// concat gather: several CHSCALE / nearest-UPSAMPLE steps that each fill one channel slice of the same concat buffer
// run as ONE kernel that writes whole pixels (all slices, contiguous 16-byte pieces) — instead of one pass per slice of
// 48-byte partial-sector writes into a buffer that is larger than L2.
// ------------------------------------------------------------------------------------------------
struct GatherDev {
    GatherSrc s[4];
    int n, total_cvecs;
    const ImgTab* tout;
};

template <typename T>
__global__ void __launch_bounds__(256) concat_gather_kernel(GatherDev p) {
    pdl_wait();      // pdl.cuh: nothing below may run before the previous kernel of the stream has completed
    pdl_trigger();
    // grid: x over (column, 16-byte piece) of one output row, y = row, z = image: one integer division per thread
    const int img = blockIdx.z, oy = blockIdx.y;
    const ImgTab to = p.tout[img];
    if (oy >= to.h) return;
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= to.w * p.total_cvecs) return;
    const int ox = l / p.total_cvecs;
    int piece = l - ox * p.total_cvecs;
    int k = 0;
    while (k + 1 < p.n && piece >= p.s[k].cvecs) { piece -= p.s[k].cvecs; k++; }
    const GatherSrc& g = p.s[k];
    const ImgTab ti = g.tin[img];
    const int iy = g.shift >= 0 ? oy >> g.shift : oy / g.scale_px, ix = g.shift >= 0 ? ox >> g.shift : ox / g.scale_px;
    float x[8];
    load8<T>(reinterpret_cast<const T*>(g.in) + (size_t(ti.off) + size_t(iy) * ti.w + ix) * g.in_cs + piece * 8, x);
    if (g.scale) {
        const float* sp = g.scale + size_t(img) * g.scale_c;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int c = piece * 8 + j;
            const float sc = c < g.scale_c ? sp[c] : 0.f;
            x[j] = g.residual ? x[j] + x[j] * sc : x[j] * sc;
        }
    }
    store8<T>(reinterpret_cast<T*>(g.out) + (size_t(to.off) + size_t(oy) * to.w + ox) * g.out_cs + piece * 8, x);
}

void launch_concat_gather(const GatherSrc* src, int n, const ImgTab* tout, int n_img, int max_out_h, int max_out_w, cudaStream_t st,
                          int prec) {
    GatherDev d{};
    d.n = n;
    d.tout = tout;
    for (int i = 0; i < n; i++) {
        d.s[i] = src[i];
        d.total_cvecs += src[i].cvecs;
        int sh = -1;
        for (int b2 = 0; b2 < 8; b2++)
            if ((1 << b2) == src[i].scale_px) sh = b2;
        d.s[i].shift = sh;
    }
    dim3 grid(cdiv_i(int64_t(max_out_w) * d.total_cvecs, 256), max_out_h, n_img);
    if (prec == 1) pdl_launch(concat_gather_kernel<float>, grid, 256, 0, st, d);
    else pdl_launch(concat_gather_kernel<__half>, grid, 256, 0, st, d);
}

}  // namespace vse
