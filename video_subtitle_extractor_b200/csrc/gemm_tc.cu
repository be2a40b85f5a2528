// gemm_tc.cu — tcgen05 / TMA / TMEM implicit-GEMM convolution for sm_100a.
//
// The dense contractions of the shipped graphs (reference backend/models/*/inference.pdmodel: `conv2d` 1x1 / KxK stride 1,
// `conv2d_transpose` 2x2 stride 2 as four 1x1 launches, `matmul_v2` — SURVEY.md Appendix B/F) run here on the 5th-generation
// tensor cores:
//   A  = pixel-major activations loaded by TMA straight into 128B-swizzled shared memory.  1x1: 2-D tiles [128 pixels x one
//        128-byte K row]; KxK: 4-D boxes shifted per filter tap (out-of-bounds rows / columns are zero-filled by TMA, which IS
//        the convolution padding) — one HALO box per k-block serving every tap (kernels up to 5x5), ROW boxes serving a group
//        of vertical taps (7x7, 9x9), per-tap boxes otherwise; maps of height 1 (text lines) use 128 x 1 tiles;
//   B  = K-major weights [cout][tap][cin], packed once per plan; resident in shared memory when small, else streamed per stage;
//   D  = fp32 accumulators in a TMEM ring, so the epilogue of tile i overlaps the MMAs of the following tiles;
//   epilogue (8 warps, two per TMEM lane quarter): tcgen05.ld -> bias -> act -> affine -> (+gate) -> +residual -> act ->
//        128B-swizzled staging tile -> TMA store into the (possibly concat-aliased, possibly strided) output slice.
// Operand modes: fp16 activations (kind::f16), fp32 activations as tf32 (kind::tf32), and the SPLIT mode the engine runs by
// default — fp32 activations rewritten in shared memory by four transform warps as fp16 hi | lo rows, three kind::f16 MMAs per
// product (or the stacked form: two operand fetches, four MMAs) — see conv_tc_kernel / umma_kblock.
// Persistent CTAs (one per SM) walk the tile list; warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner, warps 2-9 =
// epilogue, warps 10-13 = operand transform (split mode; optionally the fused depthwise stage).  The fast nets are HBM-bound
// (35-59 FLOP/B): the point of the tensor cores is to get the FMA work out of the way so that the kernel streams activations
// at memory speed; the server models (309-649 FLOP/B) are bound by the tensor pipe and the L2 -> shared-memory operand traffic.
#include "gemm_tc.h"
#include "pdl.cuh"

#include <cuda_fp16.h>

#include <cmath>
#include <cstdlib>
#include <cstring>

#include "plan.h"

namespace vse {

static constexpr int BLOCK_M = 128, BLOCK_K = 64;   // BLOCK_K: fp16 elements per 128-byte K row (32 for fp32 / tf32)
static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
static constexpr int kEpiWarps = 8;
static constexpr int kXfWarps = 4;              // split mode: warps that turn fp32 operand rows into fp16 hi | lo in place
static constexpr int kThreadsBase = 64 + 32 * kEpiWarps;
static constexpr int kThreads = kThreadsBase + 32 * kXfWarps;   // launch bound; plain fp16 / tf32 launches use kThreadsBase
static constexpr int kMaxAccStages = 8;         // TMEM accumulator ring (512 columns / n_chunk, at most 8)
static constexpr int kBarRegion = 512;          // shared-memory bytes reserved for the mbarriers + TMEM slot
static constexpr int kBResidentMax = 80 * 1024; // weight matrices up to this size stay in shared memory for the whole kernel
static constexpr int kBResidentMaxSplit = 112 * 1024;   // split mode: hi | lo weights are twice the bytes, the staging ring shrinks instead
static constexpr int kOutBufs = 4;               // ring of [128 rows x 128 B] staging tiles for the TMA stores
static constexpr int kOutBufBytes = BLOCK_M * 128;
static constexpr int kParamSmemMaxCh = 1024;   // bias/scale/shift of up to this many channels are staged in shared memory
static constexpr int kMaxSmem = 227 * 1024;

struct TcParams {
    int spatial, M, n_img, H, W, tiles_x, tiles_y, kh, kw, ph, pw, num_kb, k_pad, n_chunk, n_chunks, n_store, num_m_tiles,
        stages, tmem_cols, cin, rowbox, a_bytes, acc_stages, b_resident, b_total, halo, a_tx, tf32, kb_elems, split, kbb, out_bufs,
        stack, acc_cols, kh_g, direct1, dw, dw_cp, dw_act, tw_shift, tile_h;   // dw: fused depthwise kernel size (0 = none)   // kh_g: rowbox, vertical taps per A box;   // stack: see 'stacked split' (umma_kblock); acc_cols: TMEM columns per accumulator; spatial tiles: 2^tw_shift columns x tile_h rows = 128 pixels (halo 8 x 16, default 16 x 8, text-line maps 128 x 1)
    void* out;
    int out_cs;
    const float* bias;
    const float* post_scale;
    const float* post_shift;
    const void* res;
    int res_cs, act, act2;
    float hs_slope, hs_offset;
    const float* gate;   // optional per-image channel gate (fused squeeze-excite): v += v * gate[row / gate_rows][channel]
    int gate_c, gate_rows;
    float a_scale;     // split mode: operand rows are multiplied by this power of two before the fp16 split (see transform warps)
    float acc_scale;   // accumulator -> output units (1 / (a_scale * weight scale); 1 outside split mode)
    // ragged batches (recogniser crops of different widths): several groups of equal-sized images in ONE launch; group g has
    // its own activation / output tensor maps (gmaps[2g], gmaps[2g + 1], device memory) and geometry (gtab[g])
    const CUtensorMap* gmaps;
    const TcGroupDev* gtab;
    int n_groups;
    int n_total;       // n_chunks * n_chunk (bias / post arrays are readable up to here)
    int param_smem;    // 1: bias/scale/shift staged in shared memory
    const float *dw_w, *dw_bias, *dw_ps, *dw_pt;   // fused depthwise: [tap][dw_cp] filter, per-channel bias / post-affine (ps may be null)
};

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t addr, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    return ok != 0;
}
// bounded wait: a protocol bug must trap (launch error on the host) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t spins = 0;
    long long t0 = 0;
    while (!mbar_try_wait(addr, parity)) {
        if ((++spins & 0x3FFu) == 0) {
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 4000000000LL) __trap();
        }
    }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(m)),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(m)),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ void st_shared_16(uint32_t addr, uint4 v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// K-major operand tile in 128B-swizzled shared memory: rows of 128 bytes, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= uint64_t((saddr & 0x3FFFFu) >> 4);     // start address
    d |= uint64_t(1) << 16;                     // leading byte offset (unused for swizzled K-major; canonical 1)
    d |= uint64_t(1024 >> 4) << 32;             // stride byte offset: next 8-row group
    d |= uint64_t(1) << 46;                     // descriptor version (Blackwell)
    d |= uint64_t(2) << 61;                     // SWIZZLE_128B
    return d;
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// the descriptor's high word is constant (SBO = 1024 B, version 1, SWIZZLE_128B); the low word is (addr >> 4) | LBO field
static constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t saddr) { return ((saddr & 0x3FFFFu) >> 4) | (1u << 16); }
__device__ __forceinline__ void umma_f16_words(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        ".reg .b64 da, db;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "mov.b64 da, {%1, %5};\n"
        "mov.b64 db, {%2, %5};\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(kDescHi)
        : "memory");
}
__device__ __forceinline__ void umma_f16_words2(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t idesc,
                                                uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        ".reg .b64 da, db;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "mov.b64 da, {%1, %6};\n"
        "mov.b64 db, {%2, %5};\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(kDescHi), "r"(a_hi)
        : "memory");
}
// both descriptors' high words given (the stacked split reads its weights through a SWIZZLE_64B descriptor)
__device__ __forceinline__ void umma_f16_words3(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                                uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        ".reg .b64 da, db;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "mov.b64 da, {%1, %6};\n"
        "mov.b64 db, {%2, %5};\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(b_hi), "r"(a_hi)
        : "memory");
}
// kind::tf32: fp32 operands in shared memory (rounded to tf32 by the tensor core), 8 elements (32 bytes) of K per instruction
__device__ __forceinline__ void umma_tf32_words2(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t idesc,
                                                 uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        ".reg .b64 da, db;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "mov.b64 da, {%1, %6};\n"
        "mov.b64 db, {%2, %5};\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(kDescHi), "r"(a_hi)
        : "memory");
}
template <bool TF32>
__device__ __forceinline__ void umma_words(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t idesc,
                                           uint32_t accumulate) {
    if constexpr (TF32) umma_tf32_words2(d_tmem, a_lo, a_hi, b_lo, idesc, accumulate);
    else umma_f16_words2(d_tmem, a_lo, a_hi, b_lo, idesc, accumulate);
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// activation with the kind known at compile time (the epilogue loop is instantiated per activation)
template <int ACT>
__device__ __forceinline__ float act_t(float x, float slope, float offset) {
    if constexpr (ACT == ACT_RELU) return fmaxf(x, 0.f);
    else if constexpr (ACT == ACT_HSWISH) return x * __saturatef(fmaf(x, 1.f / 6.f, 0.5f));   // x * clip(x + 3, 0, 6) / 6
    else if constexpr (ACT == ACT_HSIGMOID) return __saturatef(fmaf(x, slope, offset));
    else if constexpr (ACT == ACT_SWISH) return __fdividef(x, 1.f + __expf(-x));
    else if constexpr (ACT == ACT_SIGMOID) return __fdividef(1.f, 1.f + __expf(-x));
    else if constexpr (ACT == ACT_RELU6) return fminf(fmaxf(x, 0.f), 6.f);
    else return x;
}

__device__ __forceinline__ float tc_act(float x, int act, float slope, float offset) {
    switch (act) {
        case ACT_RELU: return fmaxf(x, 0.f);
        case ACT_HSWISH: return x * __saturatef(fmaf(x, 1.f / 6.f, 0.5f));
        case ACT_HSIGMOID: return __saturatef(fmaf(x, slope, offset));
        case ACT_SWISH: return __fdividef(x, 1.f + __expf(-x));
        case ACT_SIGMOID: return __fdividef(1.f, 1.f + __expf(-x));
        case ACT_RELU6: return fminf(fmaxf(x, 0.f), 6.f);
        default: return x;
    }
}

// geometry of the group a tile belongs to (ragged launches); tiles come in increasing order per CTA, so the search resumes
// where it stopped.  A group's tensor maps live in global memory and are rewritten by an upload kernel for every new batch
// geometry: the thread that hands them to TMA acquires them through the tensormap proxy first.
struct TileGeo {
    const CUtensorMap* ma;
    const CUtensorMap* mo;
    int H, W, tiles_x, tiles_y, m_local, g;
    long long pix_off;
};
__device__ __forceinline__ void tensormap_acquire(const CUtensorMap* m) {
    asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tile_geo(const TcParams& p, const CUtensorMap* ma0, const CUtensorMap* mo0, int m_tile, TileGeo& c,
                                         bool acquire_a, bool acquire_o) {
    if (p.n_groups == 0) {
        c.ma = ma0; c.mo = mo0; c.H = p.H; c.W = p.W; c.tiles_x = p.tiles_x; c.tiles_y = p.tiles_y; c.m_local = m_tile; c.pix_off = 0;
        return;
    }
    int g = c.g < 0 ? 0 : c.g;
    while (g + 1 < p.n_groups && m_tile >= p.gtab[g + 1].tile_begin) g++;
    if (g != c.g) {
        c.g = g;
        const TcGroupDev t = p.gtab[g];
        c.ma = p.gmaps + 2 * g; c.mo = p.gmaps + 2 * g + 1;
        c.H = t.H; c.W = t.W; c.tiles_x = t.tiles_x; c.tiles_y = t.tiles_y; c.pix_off = t.pix_off;
        if (acquire_a) tensormap_acquire(c.ma);
        if (acquire_o) tensormap_acquire(c.mo);
    }
    c.m_local = m_tile - p.gtab[g].tile_begin;
}

// ------------------------------------------------------------------------------------------------
// epilogue: 8 warps.  Warp w may only touch TMEM lanes [32 * (w % 4), +32) (hardware rule), so q = w % 4 picks the
// 32 output pixels (one per lane) and `half` splits the accumulator columns in 32-column pairs between the two warps
// that share a lane quarter.  Per 16 columns: tcgen05.ld -> +bias -> act -> affine -> (+residual -> act2) -> fp16.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// What the epilogue needs from TcParams, read ONCE into registers by epilogue_loop (TcParams lives in local memory once
// it is passed by reference into the non-inlined loops: every p.field inside the column loop was a local-memory load).
struct EpiRegs {
    const float *pb, *ps, *pt;     // per-channel bias / post scale / post shift (generic pointers: global or shared)
    uint32_t pb_s, ps_s, pt_s;     // the same as shared-memory addresses when they are staged there (PSM)
    const void* res;
    int res_cs, n_store, act, act2, gate_c;
    float hs_slope, hs_offset, acc_scale;
};

// per-channel constants: PSM = staged in shared memory -> ld.shared (the generic loads the compiler has to emit for a
// pointer that may be global or shared cost a descriptor set-up per load on top of the load)
template <bool PSM>
__device__ __forceinline__ float4 ld_const4(const float* g, uint32_t s_addr, int idx) {
    if constexpr (PSM) {
        float4 v;
        asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(s_addr + uint32_t(idx) * 4u));
        return v;
    } else {
        return ld4(g + idx);
    }
}

// 16 accumulator columns of one output pixel -> 16 output values: fp16 (two 16-byte pieces) or fp32 (four pieces).
// ACT < 0: activation kind read at run time (the rare layers whose constants do not fit shared memory).
template <int ACT, bool POST, bool OUTF32, bool PSM, bool GATE>
__device__ __forceinline__ void epi_chunk16(const EpiRegs& e, const uint32_t* raw, int cb, long long pix, uint4* out4, const float* grow) {
    float v[16];
    const float* gp = nullptr;   // gate row of this pixel's image, at this chunk's first channel (one modulo per 16 columns)
    if constexpr (GATE) gp = grow ? grow + (cb % e.gate_c) : nullptr;
#pragma unroll
    for (int j4 = 0; j4 < 4; j4++) {
        const float4 b = ld_const4<PSM>(e.pb, e.pb_s, cb + 4 * j4);
        // acc * acc_scale + bias: acc_scale is 1 outside split mode (one rounding either way: identical to acc + bias)
        float x0 = fmaf(__uint_as_float(raw[4 * j4 + 0]), e.acc_scale, b.x), x1 = fmaf(__uint_as_float(raw[4 * j4 + 1]), e.acc_scale, b.y);
        float x2 = fmaf(__uint_as_float(raw[4 * j4 + 2]), e.acc_scale, b.z), x3 = fmaf(__uint_as_float(raw[4 * j4 + 3]), e.acc_scale, b.w);
        if constexpr (ACT >= 0) {
            x0 = act_t<ACT>(x0, e.hs_slope, e.hs_offset); x1 = act_t<ACT>(x1, e.hs_slope, e.hs_offset);
            x2 = act_t<ACT>(x2, e.hs_slope, e.hs_offset); x3 = act_t<ACT>(x3, e.hs_slope, e.hs_offset);
        } else {
            x0 = tc_act(x0, e.act, e.hs_slope, e.hs_offset); x1 = tc_act(x1, e.act, e.hs_slope, e.hs_offset);
            x2 = tc_act(x2, e.act, e.hs_slope, e.hs_offset); x3 = tc_act(x3, e.act, e.hs_slope, e.hs_offset);
        }
        if constexpr (POST) {
            const float4 sc = ld_const4<PSM>(e.ps, e.ps_s, cb + 4 * j4), sh = ld_const4<PSM>(e.pt, e.pt_s, cb + 4 * j4);
            x0 = fmaf(x0, sc.x, sh.x); x1 = fmaf(x1, sc.y, sh.y); x2 = fmaf(x2, sc.z, sh.z); x3 = fmaf(x3, sc.w, sh.w);
        }
        if constexpr (GATE) {
            if (gp) {   // residual squeeze-excite: y + y * gate
                const float4 g = __ldg(reinterpret_cast<const float4*>(gp + 4 * j4));
                x0 = fmaf(x0, g.x, x0); x1 = fmaf(x1, g.y, x1); x2 = fmaf(x2, g.z, x2); x3 = fmaf(x3, g.w, x3);
            }
        }
        v[4 * j4 + 0] = x0; v[4 * j4 + 1] = x1; v[4 * j4 + 2] = x2; v[4 * j4 + 3] = x3;
    }
#pragma unroll
    for (int h8 = 0; h8 < 2; h8++) {
        if (e.res && pix >= 0 && cb + h8 * 8 < e.n_store) {
            if constexpr (OUTF32) {
                const float* rp = static_cast<const float*>(e.res) + size_t(pix) * e.res_cs + cb + h8 * 8;
                const float4 r0 = ld4(rp), r1 = ld4(rp + 4);
                v[h8 * 8 + 0] += r0.x; v[h8 * 8 + 1] += r0.y; v[h8 * 8 + 2] += r0.z; v[h8 * 8 + 3] += r0.w;
                v[h8 * 8 + 4] += r1.x; v[h8 * 8 + 5] += r1.y; v[h8 * 8 + 6] += r1.z; v[h8 * 8 + 7] += r1.w;
            } else {
                const uint4 r4 = *reinterpret_cast<const uint4*>(static_cast<const __half*>(e.res) + size_t(pix) * e.res_cs + cb + h8 * 8);
                const __half2* rh = reinterpret_cast<const __half2*>(&r4);
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const float2 f = __half22float2(rh[j]);
                    v[h8 * 8 + 2 * j] += f.x;
                    v[h8 * 8 + 2 * j + 1] += f.y;
                }
            }
        }
        if (e.act2 != ACT_NONE) {
#pragma unroll
            for (int j = 0; j < 8; j++) v[h8 * 8 + j] = tc_act(v[h8 * 8 + j], e.act2, 0.f, 0.f);
        }
        if constexpr (OUTF32) {
            out4[2 * h8] = make_uint4(__float_as_uint(v[h8 * 8 + 0]), __float_as_uint(v[h8 * 8 + 1]), __float_as_uint(v[h8 * 8 + 2]),
                                      __float_as_uint(v[h8 * 8 + 3]));
            out4[2 * h8 + 1] = make_uint4(__float_as_uint(v[h8 * 8 + 4]), __float_as_uint(v[h8 * 8 + 5]), __float_as_uint(v[h8 * 8 + 6]),
                                          __float_as_uint(v[h8 * 8 + 7]));
        } else {
            __half2* hh = reinterpret_cast<__half2*>(&out4[h8]);
#pragma unroll
            for (int j = 0; j < 4; j++) hh[j] = __floats2half2_rn(v[h8 * 8 + 2 * j], v[h8 * 8 + 2 * j + 1]);
        }
    }
}

// The 8 epilogue warps turn the accumulator into fp16 in 64-column sub-tiles: every thread writes its pixel's 32 columns
// into a 128B-swizzled [128 pixels x 128 B] staging tile, the warps meet at a named barrier, and one thread hands the tile
// to TMA (cp.async.bulk.tensor store): full-line, coalesced global writes, rows past the end of the tensor clipped by the
// hardware.  A ring of kOutBufs staging tiles keeps kOutBufs - 1 stores in flight.
template <int ACT, bool POST, bool OUTF32, bool PSM, bool GATE>
__device__ __noinline__ void epilogue_loop(const TcParams& p, const CUtensorMap* map_o, uint8_t* sout, uint32_t tmem_base,
                                           uint64_t* tmem_full, uint64_t* tmem_empty, const float* pb, const float* ps,
                                           const float* pt, int q, int half, int lane, bool issuer) {
    EpiRegs e;
    e.pb = pb; e.ps = ps; e.pt = pt;
    e.pb_s = PSM ? smem_u32(pb) : 0u; e.ps_s = PSM ? smem_u32(ps) : 0u; e.pt_s = PSM ? smem_u32(pt) : 0u;
    e.res = p.res; e.res_cs = p.res_cs; e.n_store = p.n_store; e.act = p.act; e.act2 = p.act2; e.gate_c = p.gate_c;
    e.hs_slope = p.hs_slope; e.hs_offset = p.hs_offset; e.acc_scale = p.acc_scale;
    const int row = q * 32 + lane;
    const int total_tiles = p.num_m_tiles * p.n_chunks;
    constexpr int SUBC = OUTF32 ? 32 : 64;        // columns per 128-byte staging row
    const int n_sub = (p.n_chunk + SUBC - 1) / SUBC;
    const uint32_t sout_addr = smem_u32(sout);
    const int out_bufs = p.out_bufs;
    int acc = 0, slot = 0;
    uint32_t acc_phase = 0;
    TileGeo tg;
    tg.g = -1;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int m_tile = tile / p.n_chunks, n_idx = tile - m_tile * p.n_chunks;
        long long pix = -1;
        int img = 0, y0 = 0, x0 = 0;
        if (p.spatial) {
            tile_geo(p, nullptr, map_o, m_tile, tg, false, issuer);
            const int per_img = tg.tiles_x * tg.tiles_y;
            img = tg.m_local / per_img;
            const int r = tg.m_local - img * per_img;
            y0 = (r / tg.tiles_x) * p.tile_h;
            x0 = (r % tg.tiles_x) << p.tw_shift;
            const int y = y0 + (row >> p.tw_shift), x = x0 + (row & ((1 << p.tw_shift) - 1));
            if (y < tg.H && x < tg.W) pix = tg.pix_off + (long long)img * tg.H * tg.W + (long long)y * tg.W + x;
        } else {
            const long long m = (long long)m_tile * BLOCK_M + row;
            if (m < p.M) pix = m;
        }
        const float* grow = (GATE && p.gate && pix >= 0) ? p.gate + size_t(pix / p.gate_rows) * p.gate_c : nullptr;
        mbar_wait(&tmem_full[acc], acc_phase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(acc * p.acc_cols);
        const int ch0 = n_idx * p.n_chunk;
        for (int sub = 0; sub < n_sub; sub++) {
            if (ch0 + sub * SUBC >= p.n_store) break;                  // uniform over the 8 warps
            const int c0 = sub * SUBC + half * (SUBC / 2);              // this warp's columns of the sub-tile
            const uint32_t buf = sout_addr + uint32_t(slot * kOutBufBytes) + uint32_t(row * 128);
            if (c0 < p.n_chunk && ch0 + c0 < p.n_store) {               // warp-uniform
                const int j0 = half * 4;                                // 16-byte piece index inside the 128-byte row
                if constexpr (OUTF32) {
                    uint32_t raw0[16];
                    tmem_ld16_nowait(taddr + uint32_t(c0), raw0);       // .sync.aligned: whole (converged) warp
                    if (p.stack) {                                      // stacked split: hi * Wl lives n_chunk columns further
                        uint32_t rawl[16];
                        tmem_ld16_nowait(taddr + uint32_t(p.n_chunk + c0), rawl);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 16; j++) raw0[j] = __float_as_uint(__uint_as_float(raw0[j]) + __uint_as_float(rawl[j]));
                    } else {
                        tmem_ld_wait();
                    }
                    uint4 o[4];
                    epi_chunk16<ACT, POST, true, PSM, GATE>(e, raw0, ch0 + c0, pix, o, grow);
                    if (p.direct1) {    // one dense fp32 channel: lane = pixel, a warp writes 128 contiguous bytes
                        if (pix >= 0) static_cast<float*>(p.out)[pix * p.out_cs] = __uint_as_float(o[0].x);
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; j++) st_shared_16(buf + uint32_t(((j0 + j) ^ (row & 7)) << 4), o[j]);
                    }
                } else {
                    const bool two = c0 + 16 < p.n_chunk && ch0 + c0 + 16 < p.n_store;
                    uint32_t raw0[16], raw1[16];
                    tmem_ld16_nowait(taddr + uint32_t(c0), raw0);
                    if (two) tmem_ld16_nowait(taddr + uint32_t(c0 + 16), raw1);
                    tmem_ld_wait();
                    uint4 o[2];
                    epi_chunk16<ACT, POST, false, PSM, GATE>(e, raw0, ch0 + c0, pix, o, grow);
                    st_shared_16(buf + uint32_t(((j0 + 0) ^ (row & 7)) << 4), o[0]);
                    st_shared_16(buf + uint32_t(((j0 + 1) ^ (row & 7)) << 4), o[1]);
                    if (two) {
                        epi_chunk16<ACT, POST, false, PSM, GATE>(e, raw1, ch0 + c0 + 16, pix, o, grow);
                        st_shared_16(buf + uint32_t(((j0 + 2) ^ (row & 7)) << 4), o[0]);
                        st_shared_16(buf + uint32_t(((j0 + 3) ^ (row & 7)) << 4), o[1]);
                    }
                }
                fence_proxy_async();                                    // generic-proxy writes -> visible to the TMA store
            }
            if (p.direct1) continue;                                    // (warp-uniform) nothing staged, nothing for TMA
            epi_barrier();
            if (issuer) {
                const void* src = sout + size_t(slot) * kOutBufBytes;
                if (p.spatial) tma_store_4d(tg.mo, src, ch0 + sub * SUBC, x0, y0, img);
                else tma_store_2d(map_o, src, ch0 + sub * SUBC, m_tile * BLOCK_M);
                tma_store_commit();
                // the tile written two barriers from now is free again (see ring note); shorter rings wait for more stores
                if (out_bufs >= 4) tma_store_wait_read<2>();
                else if (out_bufs == 3) tma_store_wait_read<1>();
                else tma_store_wait_read<0>();
            }
            if (++slot == out_bufs) slot = 0;
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        if (++acc == p.acc_stages) { acc = 0; acc_phase ^= 1u; }
    }
    if (issuer) tma_store_wait_all();
}

// Instantiations: the per-channel constants are in shared memory for every layer of up to kParamSmemMaxCh output channels
// (compile-time activation, ld.shared constants); the fused squeeze-excite gate only exists on plain fp16 1x1 convolutions
// (engine.cu: act none, no post-affine); wider layers (the CTC class projections) take the run-time-activation variant
// that reads its constants from global memory.
template <bool POST>
__device__ __forceinline__ void epilogue_dispatch(const TcParams& p, const CUtensorMap* map_o, uint8_t* sout, uint32_t tmem_base,
                                                  uint64_t* tmem_full, uint64_t* tmem_empty, const float* pb, const float* ps,
                                                  const float* pt, int q, int half, int lane, bool issuer) {
#define VSE_EPI_ARGS p, map_o, sout, tmem_base, tmem_full, tmem_empty, pb, ps, pt, q, half, lane, issuer
    const bool of32 = p.tf32 || p.split;   // fp32 activations in, fp32 out
    if (!p.param_smem || (p.gate && (POST || p.tf32 || p.act != ACT_NONE))) {
        if (p.gate) {   // not produced by the engine; kept correct rather than fast
            if (of32) epilogue_loop<-1, POST, true, false, true>(VSE_EPI_ARGS);
            else epilogue_loop<-1, POST, false, false, true>(VSE_EPI_ARGS);
        } else {
            if (of32) epilogue_loop<-1, POST, true, false, false>(VSE_EPI_ARGS);
            else epilogue_loop<-1, POST, false, false, false>(VSE_EPI_ARGS);
        }
        return;
    }
    if (p.gate) {
        if constexpr (!POST) {
            if (of32) epilogue_loop<ACT_NONE, false, true, true, true>(VSE_EPI_ARGS);
            else epilogue_loop<ACT_NONE, false, false, true, true>(VSE_EPI_ARGS);
        }
        return;
    }
#define VSE_EPI(A)                                                                  \
    do {                                                                            \
        if (of32) epilogue_loop<A, POST, true, true, false>(VSE_EPI_ARGS);          \
        else epilogue_loop<A, POST, false, true, false>(VSE_EPI_ARGS);              \
    } while (0)
    switch (p.act) {
        case ACT_RELU: VSE_EPI(ACT_RELU); break;
        case ACT_HSWISH: VSE_EPI(ACT_HSWISH); break;
        case ACT_HSIGMOID: VSE_EPI(ACT_HSIGMOID); break;
        case ACT_SWISH: VSE_EPI(ACT_SWISH); break;
        case ACT_SIGMOID: VSE_EPI(ACT_SIGMOID); break;
        case ACT_RELU6: VSE_EPI(ACT_RELU6); break;
        default: VSE_EPI(ACT_NONE); break;
    }
#undef VSE_EPI
#undef VSE_EPI_ARGS
}

// ------------------------------------------------------------------------------------------------
// MMA issuer warp.  The whole warp walks the loop (warp-uniform control flow and addresses, so the descriptor arithmetic
// stays on the uniform datapath); one elected lane issues the tcgen05 instructions.  With N as small as 32 an MMA occupies
// the tensor pipe for ~16 cycles, so the scalar instructions spent per k-iteration here bound the kernel: the loop is
// instantiated per addressing mode (MODE 0: one operand tile per k-iteration; 1: row box, KH vertical taps per tile;
// 2: halo box, KH x KW taps per tile) with the common 3x3 shapes unrolled (KH / KW = 0: run-time extents), and every
// per-iteration quantity is a running sum instead of a product.
// ------------------------------------------------------------------------------------------------
// one k-block of one filter tap: up to four K = 32-byte MMAs — or, in split mode, the six MMAs of the 3-term product: the
// operand row is [hi(32 ch) | lo(32 ch)] (fp16, written in place by the transform warps), the weight row [Wh(32) | Wl(32)]:
//   hi * Wh (2 MMAs), lo * Wh (2), hi * Wl (2); lo * Wl (2^-22 relative) is dropped.
// Stacked split (narrow, K-heavy layers: N <= 32, where an MMA's time is the fetch of its 128 A rows, not the math): the weight
// slice of a k-block is [2n rows x 64 B] (SWIZZLE_64B) — rows [0, n) hold Wh, rows [n, 2n) Wl — so ONE N = 2n MMA on the hi rows
// yields hi * Wh (columns [0, n)) and hi * Wl (columns [n, 2n)), and one N = n MMA on the lo rows adds lo * Wh to the first n
// columns: four MMAs / two A fetches per 16 channels x 2 instead of six / three.  The epilogue adds the two column groups.
static constexpr uint32_t kDescHi64 = (512u >> 4) | (1u << 14) | (4u << 29);   // SBO = 512 B (8 rows of 64 B), version 1, SWIZZLE_64B
template <bool TF32, bool SPLIT>
__device__ __forceinline__ void umma_kblock(uint32_t d_tmem, uint32_t a, uint32_t a_hi, uint32_t b, uint32_t idesc, uint32_t acc_first, int ks,
                                            uint32_t idesc2 = 0) {
    if constexpr (SPLIT) {
        // ks = 1: the k-block holds at most 16 real channels (K tail) — the second K = 16 slice of hi and of lo is all zero
        if (idesc2 != 0) {
            umma_f16_words3(d_tmem, a, a_hi, b, kDescHi64, idesc2, acc_first);
            if (ks > 1) umma_f16_words3(d_tmem, a + 2, a_hi, b + 2, kDescHi64, idesc2, 1u);
            umma_f16_words3(d_tmem, a + 4, a_hi, b, kDescHi64, idesc, 1u);
            if (ks > 1) umma_f16_words3(d_tmem, a + 6, a_hi, b + 2, kDescHi64, idesc, 1u);
        } else {
            umma_f16_words2(d_tmem, a, a_hi, b, idesc, acc_first);
            if (ks > 1) umma_f16_words2(d_tmem, a + 2, a_hi, b + 2, idesc, 1u);
            umma_f16_words2(d_tmem, a + 4, a_hi, b, idesc, 1u);
            if (ks > 1) umma_f16_words2(d_tmem, a + 6, a_hi, b + 2, idesc, 1u);
            umma_f16_words2(d_tmem, a, a_hi, b + 4, idesc, 1u);
            if (ks > 1) umma_f16_words2(d_tmem, a + 2, a_hi, b + 6, idesc, 1u);
        }
    } else {
        umma_words<TF32>(d_tmem, a, a_hi, b, idesc, acc_first);
        if (ks > 1) umma_words<TF32>(d_tmem, a + 2, a_hi, b + 2, idesc, 1u);
        if (ks > 2) umma_words<TF32>(d_tmem, a + 4, a_hi, b + 4, idesc, 1u);
        if (ks > 3) umma_words<TF32>(d_tmem, a + 6, a_hi, b + 6, idesc, 1u);
    }
}

template <int MODE, int KH_, int KW_, bool TF32, bool SPLIT>
__device__ __forceinline__ void mma_warp_loop(const TcParams& p, uint64_t* full, uint64_t* empty, uint64_t* tmem_full,
                                              uint64_t* tmem_empty, uint32_t tmem_base, uint32_t a_lo0, uint32_t bres_lo,
                                              uint32_t stage16, int k_iters) {
    const int KH = KH_ ? KH_ : (MODE == 1 ? p.kh_g : p.kh), KW = KW_ ? KW_ : p.kw;
    // instruction descriptor: D = f32, A/B = f16 (kind::f16) or tf32 (kind::tf32, format code 2), N >> 3, M >> 4
    const uint32_t idesc = (1u << 4) | (TF32 ? (2u << 7) | (2u << 10) : 0u) | (uint32_t(p.n_chunk >> 3) << 17) | (uint32_t(BLOCK_M >> 4) << 24);
    const uint32_t idesc2 = (SPLIT && p.stack) ? ((1u << 4) | (uint32_t((2 * p.n_chunk) >> 3) << 17) | (uint32_t(BLOCK_M >> 4) << 24)) : 0u;
    const int total_tiles = p.num_m_tiles * p.n_chunks;
    const int stages = p.stages, acc_stages = p.acc_stages, num_kb = p.num_kb, n_chunk = p.n_chunk;
    const bool resident = p.b_resident != 0;
    const uint32_t b_step = uint32_t(n_chunk * 8);                  // one [n_chunk x 64] weight slice in 16-byte units
    const uint32_t a16 = uint32_t(p.a_bytes >> 4);
    const uint32_t op16 = p.dw ? a16 : 0u;                        // fused depthwise: the operand tile follows the input window
    const uint32_t bo16 = a16 + (p.dw ? uint32_t(A_BYTES >> 4) : 0u);   // streamed weights follow the operand
    const int umma_k = p.kb_elems / 4;   // K elements per MMA (32 bytes): 16 halves or 8 floats
    // all-zero K tail skipped (split: 32 channels per k-block = two K = 16 slices of hi and of lo; ks 4 stands for 'both')
    const int ks_last = SPLIT ? ((p.cin - (num_kb - 1) * 32) <= 16 ? 1 : 4) : min(4, (p.cin - (num_kb - 1) * p.kb_elems + umma_k - 1) / umma_k);
    // strides between the weight slices of consecutive taps
    const uint32_t b_ky_step = resident ? uint32_t(KW * num_kb) * b_step : b_step;     // MODE 1
    const uint32_t b_tap_step = resident ? uint32_t(num_kb) * b_step : b_step;         // MODE 2
    const uint32_t bw = uint32_t(8 + KW - 1);                                          // MODE 2: box width in pixels
    // MODE 2: SBO = one image row of the box.  The base-offset field stays 0: measured on B200, the 128B-swizzle XOR is
    // taken from the absolute shared-memory address bits (as TMA writes them), so a start address that is not 1024-byte
    // aligned needs no phase correction (setting the field breaks parity).
    const uint32_t halo_hi = uint32_t((bw * 128) >> 4) | (1u << 14) | (2u << 29);
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0, a_lo = a_lo0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + uint32_t(acc * p.acc_cols);
        int kb = 0, kx1 = 0, ky0 = 0;                                // MODE 1: horizontal tap and first vertical tap of the box
        uint32_t b_it = bres_lo;                                     // resident slice of (k-iteration `it`)
        for (int it = 0; it < k_iters; it++) {
            mbar_wait(&full[stage], phase);
            tc_fence_after();
            const int ks = (kb == num_kb - 1) ? ks_last : 4;
            uint32_t b_first = resident ? b_it : a_lo + bo16;
            if constexpr (MODE == 1)     // iteration = (vertical-tap group, kx, k-block): first resident slice = tap (ky0, kx1)
                if (resident) b_first = bres_lo + uint32_t((ky0 * p.kw + kx1) * num_kb + kb) * b_step;
            if (elect_one()) {
                if constexpr (MODE == 0) {
                    umma_kblock<TF32, SPLIT>(d_tmem, a_lo + op16, kDescHi, b_first, idesc, it != 0 ? 1u : 0u, ks, idesc2);
                } else if constexpr (MODE == 1) {
                    const int cnt = KH_ ? KH : min(KH, p.kh - ky0);    // KH = taps per box (p.kh_g); the last group may be short
#pragma unroll
                    for (int ky = 0; ky < cnt; ky++) {
                        const uint32_t a_t = a_lo + uint32_t(ky * 128);           // next image row of the box: 16 px * 128 B
                        const uint32_t b_t = b_first + uint32_t(ky) * b_ky_step;
                        umma_kblock<TF32, SPLIT>(d_tmem, a_t, kDescHi, b_t, idesc, (it | ky) != 0 ? 1u : 0u, ks, idesc2);
                    }
                } else {
#pragma unroll
                    for (int ky = 0; ky < KH; ky++)
#pragma unroll
                        for (int kx = 0; kx < KW; kx++) {
                            const uint32_t a_t = a_lo + (uint32_t(ky) * bw + uint32_t(kx)) * 8u;   // pixel rows of 128 B
                            const uint32_t b_t = b_first + uint32_t(ky * KW + kx) * b_tap_step;
                            umma_kblock<TF32, SPLIT>(d_tmem, a_t, halo_hi, b_t, idesc, (it | ky | kx) != 0 ? 1u : 0u, ks, idesc2);
                        }
                }
                umma_commit(&empty[stage]);                            // smem slot is free once these MMAs have read it
                if (it == k_iters - 1) umma_commit(&tmem_full[acc]);   // accumulator complete -> epilogue
            }
            __syncwarp();
            b_it += b_step;
            if (++kb == num_kb) {
                kb = 0;
                if constexpr (MODE == 1)
                    if (++kx1 == p.kw) { kx1 = 0; ky0 += KH; }
            }
            a_lo += stage16;
            if (++stage == stages) { stage = 0; phase ^= 1u; a_lo = a_lo0; }
        }
        if (++acc == acc_stages) { acc = 0; acc_phase ^= 1u; }
    }
}

// ------------------------------------------------------------------------------------------------
// Fused depthwise -> pointwise: one operand stage.  The stage holds a (16 + K - 1) x (8 + K - 1) pixel window of the depthwise
// INPUT (fp32 rows of 32 channels, 128B-swizzled by TMA; out-of-image pixels arrive as zeros = the convolution's padding).
// The 128 threads of the transform warps compute the depthwise output of the 16 x 8 tile for these 32 channels and write it
// as the [hi(32) | lo(32)] fp16 operand rows of the 1x1 convolution.  Thread t: channels 4q .. 4q + 3 (q = t & 7: one 16-byte
// chunk; eight neighbouring lanes read one whole 128-byte pixel row, four rows per warp instruction: conflict free) of the
// pixels r0 + 16 i (r0 = t >> 3, i < 8).  Accumulation per output: taps in (ky, kx) order, fmaf from 0, then + bias,
// activation, post-affine — the arithmetic of dwconv_reg_kernel (fast_kernels.cu), so the fused path is bit-identical to it.
// dwp: [K * K taps][C32] filter, then bias[C32], post scale[C32], post shift[C32] in shared memory (C32 = num_kb * 32).
// ------------------------------------------------------------------------------------------------
template <int K>
__device__ __noinline__ void dw_operand_stage(const TcParams& p, uint32_t box, uint32_t opnd, const float* dwp, int kb, int t, float asc) {
    constexpr int BW = 8 + K - 1;
    const int q = t & 7, r0 = t >> 3;
    const int C32 = p.num_kb * 32;
    const uint32_t wq = smem_u32(dwp) + uint32_t(kb * 32 + 4 * q) * 4u;      // this thread's 4 channels of tap 0
    auto ldw = [&](int row) {
        float4 v;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(wq + uint32_t(row * C32) * 4u));
        return v;
    };
    const int px = r0 & 7, py0 = r0 >> 3;
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
#pragma unroll 1
    for (int ky = 0; ky < K; ky++) {      // (not unrolled: the 25 x 8 window addresses of a 5x5 filter would all be kept in registers)
#pragma unroll
        for (int kx = 0; kx < K; kx++) {
            const float4 w = ldw(ky * K + kx);
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const uint32_t row = uint32_t((py0 + 2 * i + ky) * BW + px + kx);
                float4 x;
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w)
                             : "r"(box + row * 128u + ((uint32_t(q) ^ (row & 7u)) << 4)));
                acc[i][0] = fmaf(x.x, w.x, acc[i][0]); acc[i][1] = fmaf(x.y, w.y, acc[i][1]);
                acc[i][2] = fmaf(x.z, w.z, acc[i][2]); acc[i][3] = fmaf(x.w, w.w, acc[i][3]);
            }
        }
    }
    const float4 b4 = ldw(K * K), s4 = ldw(K * K + 1), t4 = ldw(K * K + 2);
    const float bb[4] = {b4.x, b4.y, b4.z, b4.w}, ss[4] = {s4.x, s4.y, s4.z, s4.w}, tt[4] = {t4.x, t4.y, t4.z, t4.w};
    const bool post = p.dw_ps != nullptr;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint32_t hi[2], lo[2];
        float v[4];
#pragma unroll
        for (int c = 0; c < 4; c++) {
            float x = __fadd_rn(acc[i][c], bb[c]);
            if (p.dw_act == ACT_RELU) x = fmaxf(x, 0.f);
            else if (p.dw_act == ACT_HSWISH)   // x * min(max(x + 3, 0), 6) * (1/6), left to right (fact2<ACT_HSWISH>)
                x = __fmul_rn(__fmul_rn(x, fminf(fmaxf(__fadd_rn(x, 3.f), 0.f), 6.f)), 1.f / 6.f);
            if (post) x = fmaf(x, ss[c], tt[c]);
            v[c] = x * asc;
        }
        const __half2 h0 = __floats2half2_rn(v[0], v[1]), h1 = __floats2half2_rn(v[2], v[3]);
        const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
        const __half2 l0 = __floats2half2_rn(v[0] - f0.x, v[1] - f0.y), l1 = __floats2half2_rn(v[2] - f1.x, v[3] - f1.y);
        hi[0] = *reinterpret_cast<const uint32_t*>(&h0); hi[1] = *reinterpret_cast<const uint32_t*>(&h1);
        lo[0] = *reinterpret_cast<const uint32_t*>(&l0); lo[1] = *reinterpret_cast<const uint32_t*>(&l1);
        const uint32_t r = uint32_t(r0 + 16 * i);
        const uint32_t rowa = opnd + r * 128u + uint32_t(q & 1) * 8u;
        asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(rowa + ((uint32_t(q >> 1) ^ (r & 7u)) << 4)), "r"(hi[0]), "r"(hi[1]) : "memory");
        asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(rowa + ((uint32_t(4 + (q >> 1)) ^ (r & 7u)) << 4)), "r"(lo[0]), "r"(lo[1]) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
// SPLIT: the fp32-activation / 3-term fp16 product variant (p.split; 4 extra operand-transform warps).  A separate
// instantiation so that the plain fp16 / tf32 kernel keeps its 320-thread register budget.
template <bool SPLIT>
__global__ void __launch_bounds__(SPLIT ? kThreads : kThreadsBase, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const __grid_constant__ CUtensorMap map_o, const TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    // rowbox (KxK): one A box of 8 + kh - 1 image rows per (kx, k-block) serves all kh vertical taps; B holds their kh slices
    // halo (KxK, kernels up to 5x5): ONE box of (16 + kh - 1) x (8 + kw - 1) pixels per k-block serves every tap — the MMA
    // descriptors of tap (ky, kx) start (ky * box_w + kx) pixel rows into the box (tile = 16 rows x 8 columns, so that the
    // 8-row groups of the operand are one image row each, a uniform stride apart)
    const int b_bytes = p.b_resident ? 0 : p.n_chunk * 128 * (p.halo ? p.kh * p.kw : p.rowbox ? p.kh_g : 1);
    const int opnd_bytes = p.dw ? A_BYTES : 0;                      // fused depthwise: the operand tile next to the input window
    const int stage_bytes = p.a_bytes + opnd_bytes + b_bytes;
    uint8_t* bres = smem + size_t(p.stages) * stage_bytes;          // resident weights: [tap][k-block][n_chunk x 64]
    uint8_t* sout = bres + p.b_total;                               // p.out_bufs staging tiles for the TMA stores
    uint64_t* full = reinterpret_cast<uint64_t*>(sout + p.out_bufs * kOutBufBytes);
    uint64_t* empty = full + p.stages;
    uint64_t* tmem_full = empty + p.stages;
    uint64_t* tmem_empty = tmem_full + kMaxAccStages;
    uint64_t* b_full = tmem_empty + kMaxAccStages;
    uint64_t* xf = b_full + 1;                                      // split mode: operand rows of a stage are transformed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xf + p.stages);
    float* sparam = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full) + kBarRegion);
    const float *pb = p.bias, *ps = p.post_scale, *pt = p.post_shift;
    if (p.param_smem) {
        // per-channel epilogue constants: one copy per CTA in shared memory
        for (int i = threadIdx.x; i < p.n_total; i += blockDim.x) {
            sparam[i] = p.bias[i];
            if (p.post_scale) {
                sparam[p.n_total + i] = p.post_scale[i];
                sparam[2 * p.n_total + i] = p.post_shift[i];
            }
        }
        pb = sparam; ps = sparam + p.n_total; pt = sparam + 2 * p.n_total;
    }
    // fused depthwise: filter taps, bias and post-affine of all input channels, zero beyond the real channels
    float* dwp = sparam + (p.param_smem ? 3 * p.n_total : 0);
    if (p.dw) {
        const int C32 = p.num_kb * 32, taps_dw = p.dw * p.dw;
        for (int i = threadIdx.x; i < (taps_dw + 3) * C32; i += blockDim.x) {
            const int r = i / C32, c = i - r * C32;
            float v = 0.f;
            if (c < p.cin) {
                if (r < taps_dw) v = p.dw_w[size_t(r) * p.dw_cp + c];
                else if (r == taps_dw) v = p.dw_bias[c];
                else if (p.dw_ps) v = r == taps_dw + 1 ? p.dw_ps[c] : p.dw_pt[c];
            }
            dwp[i] = v;
        }
    }

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_b);
        tma_prefetch_desc(&map_o);
        for (int s = 0; s < p.stages; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
            mbar_init(&xf[s], kXfWarps);
        }
        for (int s = 0; s < p.acc_stages; s++) {
            mbar_init(&tmem_full[s], 1);
            mbar_init(&tmem_empty[s], kEpiWarps);
        }
        mbar_init(b_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, uint32_t(p.tmem_cols));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int total_tiles = p.num_m_tiles * p.n_chunks;
    const int taps = p.kh * p.kw;
    const int n_kg = p.rowbox ? (p.kh + p.kh_g - 1) / p.kh_g : 1;     // rowbox: groups of vertical taps, one A box each
    const int k_iters = (p.halo ? 1 : p.rowbox ? p.kw * n_kg : taps) * p.num_kb;

    // Everything up to here (and the resident weight load below) touches only kernel parameters and static weights, so
    // under programmatic dependent launch (pdl.cuh) it overlaps the tail of the previous step's kernel; activations,
    // residuals and gates are first touched after pdl_wait().
    if (warp == 0 && lane == 0 && p.b_resident) {
        mbar_expect_tx(b_full, uint32_t(p.b_total));
        for (int tp = 0; tp < taps; tp++)
            for (int kb = 0; kb < p.num_kb; kb++)
                tma_load_2d(bres + size_t(tp * p.num_kb + kb) * p.n_chunk * 128, &map_b, b_full, tp * p.k_pad + kb * p.kbb, 0);
    }
    pdl_wait();
    pdl_trigger();

    if (warp == 0) {
        if (lane == 0) {
            // ---------------- TMA producer ----------------
            int stage = 0;
            uint32_t phase = 0;
            TileGeo tg;
            tg.g = -1;
            tg.ma = &map_a;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int m_tile = tile / p.n_chunks, n_idx = tile - m_tile * p.n_chunks;
                int img = 0, y0 = 0, x0 = 0;
                if (p.spatial) {
                    tile_geo(p, &map_a, nullptr, m_tile, tg, true, false);
                    const int per_img = tg.tiles_x * tg.tiles_y;
                    img = tg.m_local / per_img;
                    const int r = tg.m_local - img * per_img;
                    y0 = (r / tg.tiles_x) * p.tile_h;
                    x0 = (r % tg.tiles_x) << p.tw_shift;
                }
                for (int it = 0; it < k_iters; it++) {
                    const int tap = it / p.num_kb, kb = it - tap * p.num_kb;   // rowbox: tap = kx
                    mbar_wait(&empty[stage], phase ^ 1u);
                    int tx = p.a_tx + b_bytes;
                    if (p.rowbox && !p.b_resident) tx = p.a_tx + min(p.kh_g, p.kh - (tap / p.kw) * p.kh_g) * p.n_chunk * 128;   // ragged last group
                    mbar_expect_tx(&full[stage], uint32_t(tx));
                    uint8_t* a_dst = smem + size_t(stage) * stage_bytes;
                    if (p.dw) {
                        // window of the depthwise input for this tile and k-block; the 1x1 weights follow the operand tile
                        tma_load_4d(a_dst, tg.ma, &full[stage], kb * p.kb_elems, x0 - p.pw, y0 - p.ph, img);
                        if (!p.b_resident)
                            tma_load_2d(a_dst + p.a_bytes + opnd_bytes, &map_b, &full[stage], kb * p.kbb, n_idx * p.n_chunk);
                    } else if (p.halo) {
                        tma_load_4d(a_dst, tg.ma, &full[stage], kb * p.kb_elems, x0 - p.pw, y0 - p.ph, img);
                        for (int tp = 0; tp < taps && !p.b_resident; tp++)
                            tma_load_2d(a_dst + p.a_bytes + tp * p.n_chunk * 128, &map_b, &full[stage], tp * p.k_pad + kb * p.kbb,
                                        n_idx * p.n_chunk);
                    } else if (p.rowbox) {
                        // tap = gy * kw + kx: the box of vertical-tap group gy (taps ky0 .. ky0 + cnt - 1) at horizontal tap kx
                        const int gy = tap / p.kw, kx = tap - gy * p.kw, ky0 = gy * p.kh_g, cnt = min(p.kh_g, p.kh - ky0);
                        tma_load_4d(a_dst, tg.ma, &full[stage], kb * p.kb_elems, x0 + kx - p.pw, y0 - p.ph + ky0, img);
                        for (int j = 0; j < cnt && !p.b_resident; j++)
                            tma_load_2d(a_dst + p.a_bytes + j * p.n_chunk * 128, &map_b, &full[stage],
                                        ((ky0 + j) * p.kw + kx) * p.k_pad + kb * p.kbb, n_idx * p.n_chunk);
                    } else {
                        if (p.spatial) {
                            const int ky = tap / p.kw, kx = tap - ky * p.kw;
                            tma_load_4d(a_dst, tg.ma, &full[stage], kb * p.kb_elems, x0 + kx - p.pw, y0 + ky - p.ph, img);
                        } else {
                            tma_load_2d(a_dst, &map_a, &full[stage], kb * p.kb_elems, m_tile * BLOCK_M);
                        }
                        if (!p.b_resident)
                            tma_load_2d(a_dst + p.a_bytes, &map_b, &full[stage], tap * p.k_pad + kb * p.kbb, n_idx * p.n_chunk);
                    }
                    if (++stage == p.stages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer (see mma_warp_loop) ----------------
        if (p.b_resident) {
            mbar_wait(b_full, 0);
            tc_fence_after();
        }
        const uint32_t a_lo0 = umma_desc_lo(smem_u32(smem)), bres_lo = umma_desc_lo(smem_u32(bres));
        const uint32_t stage16 = uint32_t(stage_bytes >> 4);
#define VSE_MMA(M, A, B)                                                                                                      \
    do {                                                                                                                      \
        if constexpr (SPLIT) mma_warp_loop<M, A, B, false, true>(p, xf, empty, tmem_full, tmem_empty, tmem_base, a_lo0, bres_lo, stage16, k_iters);   \
        else if (p.tf32) mma_warp_loop<M, A, B, true, false>(p, full, empty, tmem_full, tmem_empty, tmem_base, a_lo0, bres_lo, stage16, k_iters);  \
        else mma_warp_loop<M, A, B, false, false>(p, full, empty, tmem_full, tmem_empty, tmem_base, a_lo0, bres_lo, stage16, k_iters);       \
    } while (0)
        if (p.halo) {
            if (p.kh == 3 && p.kw == 3) VSE_MMA(2, 3, 3);
            else VSE_MMA(2, 0, 0);
        } else if (p.rowbox) {
            if (p.kh == 3 && p.kh_g == 3) VSE_MMA(1, 3, 0);
            else VSE_MMA(1, 0, 0);
        } else {
            VSE_MMA(0, 1, 1);
        }
#undef VSE_MMA
    } else if (SPLIT && warp >= 2 + kEpiWarps) {
        // ---------------- split mode: operand transform, warps 10..13 ----------------
        // A stage arrives as fp32 rows of 32 channels (128 bytes, 128B-swizzled by TMA).  Each thread owns whole rows: it reads
        // the row's eight 16-byte chunks, splits every value into fp16 hi = rn(x) and lo = rn(x - hi), and writes the row back
        // in place as [hi(32) | lo(32)] in the same swizzle — a K-major fp16 operand row of 64 columns for kind::f16 MMAs.
        // 16-byte accesses of 8 consecutive rows cover all 32 banks (the swizzle XOR), so the pass is conflict free.
        // The tensor core flushes fp16 SUBNORMAL inputs to zero (measured: without scaling the products are only good to
        // 2^-12, exactly what losing every lo < 2^-14 predicts), so both operands are scaled by powers of two first: the
        // activations by p.a_scale, the weights per layer so that max |w| sits just below 2^14 (tc_pack_weights); the
        // epilogue multiplies the accumulator by the exact inverse.
        const int xt = threadIdx.x - 32 * (2 + kEpiWarps);
        const int rows = p.a_tx >> 7;
        const float asc = p.a_scale;
        int stage = 0;
        uint32_t phase = 0;
        const long long n_it = (long long)((total_tiles - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x)) * k_iters;
        for (long long it = 0; it < n_it; it++) {
            mbar_wait(&full[stage], phase);
            const uint32_t base = smem_u32(smem + size_t(stage) * stage_bytes);
            if (p.dw) {
                const int kb = int(it % p.num_kb);
                if (p.dw == 3) dw_operand_stage<3>(p, base, base + uint32_t(p.a_bytes), dwp, kb, xt, asc);
                else dw_operand_stage<5>(p, base, base + uint32_t(p.a_bytes), dwp, kb, xt, asc);
            }
            for (int r = xt; r < rows && !p.dw; r += 32 * kXfWarps) {
                const uint32_t row = base + uint32_t(r) * 128u, sw = uint32_t(r & 7);
                float4 v[8];
#pragma unroll
                for (int q = 0; q < 8; q++)
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[q].x), "=f"(v[q].y), "=f"(v[q].z), "=f"(v[q].w)
                                 : "r"(row + ((uint32_t(q) ^ sw) << 4)));
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    const float sx = v[q].x * asc, sy = v[q].y * asc, sz = v[q].z * asc, sw4 = v[q].w * asc;
                    const __half2 h0 = __floats2half2_rn(sx, sy), h1 = __floats2half2_rn(sz, sw4);
                    const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
                    const __half2 l0 = __floats2half2_rn(sx - f0.x, sy - f0.y), l1 = __floats2half2_rn(sz - f1.x, sw4 - f1.y);
                    hi[2 * q] = *reinterpret_cast<const uint32_t*>(&h0); hi[2 * q + 1] = *reinterpret_cast<const uint32_t*>(&h1);
                    lo[2 * q] = *reinterpret_cast<const uint32_t*>(&l0); lo[2 * q + 1] = *reinterpret_cast<const uint32_t*>(&l1);
                }
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    st_shared_16(row + ((uint32_t(q) ^ sw) << 4), make_uint4(hi[4 * q], hi[4 * q + 1], hi[4 * q + 2], hi[4 * q + 3]));
                    st_shared_16(row + ((uint32_t(q + 4) ^ sw) << 4), make_uint4(lo[4 * q], lo[4 * q + 1], lo[4 * q + 2], lo[4 * q + 3]));
                }
            }
            fence_proxy_async();                                      // generic-proxy writes -> visible to the tensor core's reads
            __syncwarp();
            if (lane == 0) mbar_arrive(&xf[stage]);
            if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
    } else {
        // ---------------- epilogue: warps 2..9 ----------------
        const int q = warp & 3, half = (warp - 2) >> 2;
        const bool issuer = threadIdx.x == 64;   // first epilogue thread: owns the bulk-store groups
        if (p.post_scale) epilogue_dispatch<true>(p, &map_o, sout, tmem_base, tmem_full, tmem_empty, pb, ps, pt, q, half, lane, issuer);
        else epilogue_dispatch<false>(p, &map_o, sout, tmem_base, tmem_full, tmem_empty, pb, ps, pt, q, half, lane, issuer);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tc_fence_after();
        tmem_dealloc(tmem_base, uint32_t(p.tmem_cols));
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static inline int round_up_i(int x, int m) { return (x + m - 1) / m * m; }

// split mode, layers without a calibrated input range (vse_set_conv_input_ranges): activations are multiplied by
// 2^VSE_SPLIT_ASHIFT (default 2^2) before the fp16 split, so that the lo halves of everything above 2^-14 * 2^11 / 2^2 = 2^-5
// stay normal fp16 numbers; the price is the range: |x| must stay below 65504 / 2^2 (violations turn into non-finite
// outputs, which the DB post-process / CTC decode report)
float tc_split_activation_scale() {
    static const float s = [] {
        const char* e = getenv("VSE_SPLIT_ASHIFT");
        int sh = e ? atoi(e) : 2;
        sh = std::max(0, std::min(12, sh));
        return std::ldexp(1.f, sh);
    }();
    return s;
}

TcWeights tc_pack_weights(const float* w, int cout, int cin, int taps, int mode) {
    const bool tf32 = mode == TC_TF32;
    TcWeights t;
    t.tf32 = tf32 ? 1 : 0;
    if (mode == TC_SPLIT) {
        // fp32 weights as fp16 hi = rn(w) and lo = rn(w - hi); per 32 input channels one 128-byte row [hi(32) | lo(32)]
        t.split = 1;
        const int n_mma = round_up_i(cout, 16);
        t.n_chunks = (n_mma + 255) / 256;
        t.n_chunk = round_up_i((n_mma + t.n_chunks - 1) / t.n_chunks, t.n_chunks > 1 ? 64 : 16);
        const int num_kb = (cin + 31) / 32;
        // narrow K-heavy layers (the 3x3 96 -> 24 convolutions of the FPN, the 9x9 / 3x3 64-channel convolutions of the server
        // detector, the recogniser's 1x3 480 -> 60 convolutions): hi / lo stacked along N, see umma_kblock
        t.stack = (t.n_chunks == 1 && ((t.n_chunk <= 32 && taps * cin >= 256) || (t.n_chunk <= 64 && taps * cin >= 512)) && !getenv("VSE_NO_STACK")) ? 1 : 0;
        t.k_pad = num_kb * (t.stack ? 32 : 64);    // columns (halves) per tap
        t.taps = taps;
        const size_t rows = size_t(t.n_chunks) * t.n_chunk * (t.stack ? 2 : 1), cols = size_t(taps) * t.k_pad;
        t.b.assign(rows * cols, 0);
        // power-of-two weight scale: max |w| * scale in [2^13, 2^14) keeps hi AND lo = w - hi of every weight that matters in
        // the fp16 normal range (the tensor core flushes subnormal inputs)
        float wmax = 0.f;
        for (size_t i = 0; i < size_t(cout) * taps * cin; i++) wmax = std::max(wmax, std::fabs(w[i]));
        int e = 0;
        if (wmax > 0.f) {
            std::frexp(wmax, &e);               // wmax = m * 2^e, m in [0.5, 1)
            e = 14 - e;                         // wmax * 2^e in [2^13, 2^14)
        }
        e = std::max(-24, std::min(24, e));
        t.w_scale = std::ldexp(1.f, e);
        for (int co = 0; co < cout; co++)
            for (int tp = 0; tp < taps; tp++)
                for (int ci = 0; ci < cin; ci++) {
                    const float f = w[(size_t(co) * taps + tp) * cin + ci] * t.w_scale;
                    const __half h = __float2half_rn(f);
                    const __half l = __float2half_rn(f - __half2float(h));
                    if (t.stack) {        // row co = hi, row n_chunk + co = lo; 32 columns per k-block
                        const size_t at = size_t(co) * cols + size_t(tp) * t.k_pad + ci;
                        std::memcpy(&t.b[at], &h, 2);
                        std::memcpy(&t.b[at + size_t(t.n_chunk) * cols], &l, 2);
                        continue;
                    }
                    const size_t at = size_t(co) * cols + size_t(tp) * t.k_pad + size_t(ci / 32) * 64 + (ci % 32);
                    std::memcpy(&t.b[at], &h, 2);
                    std::memcpy(&t.b[at + 32], &l, 2);
                }
        return t;
    }
    const int kb = tf32 ? 32 : BLOCK_K;            // elements per 128-byte K row
    const int n_mma = round_up_i(cout, 16);
    t.n_chunks = (n_mma + 255) / 256;
    t.n_chunk = round_up_i((n_mma + t.n_chunks - 1) / t.n_chunks, t.n_chunks > 1 ? 64 : 16);   // store boxes are 64 columns wide
    t.k_pad = round_up_i(cin, kb);
    t.taps = taps;
    const size_t rows = size_t(t.n_chunks) * t.n_chunk, cols = size_t(taps) * t.k_pad;
    const int wpe = tf32 ? 2 : 1;                  // 16-bit words per element
    t.b.assign(rows * cols * wpe, 0);
    for (int co = 0; co < cout; co++)
        for (int tp = 0; tp < taps; tp++)
            for (int ci = 0; ci < cin; ci++) {
                const float f = w[(size_t(co) * taps + tp) * cin + ci];
                const size_t at = size_t(co) * cols + size_t(tp) * t.k_pad + ci;
                if (tf32) {
                    std::memcpy(&t.b[at * 2], &f, 4);
                } else {
                    __half h = __float2half_rn(f);
                    std::memcpy(&t.b[at], &h, 2);
                }
            }
    return t;
}

TcWeights tc_pack_weights_pixelpacked(const float* w, int cout, int cin, int in_cs, int out_cs, int pack) {
    TcWeights t;
    const int n_real = pack * out_cs;
    const int n_mma = round_up_i(n_real, 16);
    t.n_chunks = (n_mma + 255) / 256;
    t.n_chunk = round_up_i((n_mma + t.n_chunks - 1) / t.n_chunks, t.n_chunks > 1 ? 64 : 16);
    t.k_pad = BLOCK_K;
    t.taps = 1;
    const size_t rows = size_t(t.n_chunks) * t.n_chunk, cols = BLOCK_K;
    t.b.assign(rows * cols, 0);
    for (int g = 0; g < pack; g++)
        for (int co = 0; co < cout; co++)
            for (int ci = 0; ci < cin; ci++) {
                __half h = __float2half_rn(w[size_t(co) * cin + ci]);
                uint16_t bits;
                std::memcpy(&bits, &h, 2);
                t.b[size_t(g * out_cs + co) * cols + size_t(g * in_cs + ci)] = bits;
            }
    return t;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

static std::string encode(CUtensorMap* m, void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                          const cuuint32_t* box, bool f32 = false, bool swizzle64 = false) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return "cuTensorMapEncodeTiled unavailable";
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    const CUtensorMapL2promotion pr = CU_TENSOR_MAP_L2_PROMOTION_L2_128B;   // 64B / 256B / none measured equal on these layers
    CUresult r = fn(m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, cuuint32_t(rank), base, dims, strides_bytes, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, pr, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return "cuTensorMapEncodeTiled failed (" + std::to_string(int(r)) + ")";
    return "";
}

std::string tc_conv_setup(TcConv& t, const void* in, int in_cs, int cin, const void* wdev, const TcWeights& w, bool flat,
                          int64_t pixels, int n_img, int H, int W, int kh, int kw, int ph, int pw, bool allow_rowbox, bool allow_halo) {
    t.valid = false;
    const bool a32 = w.tf32 || w.split;          // fp32 activations
    if ((reinterpret_cast<uintptr_t>(in) & 15) || (!a32 && (in_cs & 7))) return "activation view not 16-byte aligned";
    if (pixels <= 0) return "empty";
    t.kh = kh; t.kw = kw; t.ph = ph; t.pw = pw;
    t.k_pad = w.k_pad;
    t.cin = cin;
    t.rowbox = 0;
    t.halo = 0;
    t.tf32 = w.tf32;
    t.split = w.split;
    t.stack = w.split ? w.stack : 0;
    t.w_scale = w.split ? w.w_scale : 1.f;
    const int es = a32 ? 4 : 2;                  // activation element size
    const int es_b = w.tf32 ? 4 : 2;             // weight element size (split: fp16 hi | lo)
    const cuuint32_t kbe = cuuint32_t(128 / es);   // activation elements per 128-byte K row (64 halves or 32 floats)
    if (a32 && (in_cs & 3)) return "fp32 activation view not 16-byte aligned";
    t.num_kb = (cin + int(kbe) - 1) / int(kbe);
    t.n_chunk = w.n_chunk;
    t.n_chunks = w.n_chunks;
    // weights resident in shared memory when the whole (single-chunk) matrix is small: tiles then stream activations only
    const int b_all = kh * kw * t.num_kb * w.n_chunk * 128;
    t.b_resident = (w.n_chunks == 1 && b_all <= (w.split ? kBResidentMaxSplit : kBResidentMax)) ? 1 : 0;
    std::string err;
    if (flat) {
        t.spatial = 0;
        t.M = int(pixels);
        t.num_m_tiles = int((pixels + BLOCK_M - 1) / BLOCK_M);
        cuuint64_t dims[2] = {cuuint64_t(cin), cuuint64_t(pixels)};
        cuuint64_t strides[1] = {cuuint64_t(in_cs) * es};
        cuuint32_t box[2] = {kbe, BLOCK_M};
        err = encode(&t.map_a, const_cast<void*>(in), 2, dims, strides, box, a32);
    } else {
        t.spatial = 1;
        t.n_img = n_img; t.H = H; t.W = W;
        // halo tiles (16 rows x 8 columns) for kernels up to 5x5: needs >= 3 stages of one box (+ the taps' weight slices)
        const int h_box = (16 + kh - 1) * (8 + kw - 1) * 128;
        const int slack = w.split ? 4096 : 16384;   // barriers, epilogue constants, alignment
        const int h_stage = round_up_i(h_box, 1024) + (t.b_resident ? 0 : kh * kw * w.n_chunk * 128);
        t.halo = (allow_halo && kh > 1 && kh <= 5 && kw <= 5 && kw > 1 &&
                  3 * h_stage + (t.b_resident ? b_all : 0) <= kMaxSmem - slack - (w.split ? 2 : kOutBufs) * kOutBufBytes) ? 1 : 0;
        // tile shape: 8 x 16 (halo), 16 x 8, or — maps of height 1 (the recogniser's sequence part: 1 x kw convolutions over
        // text lines) — one row of 128 columns: a 16 x 8 tile would spend 7/8 of its rows on nothing
        t.tile_w = t.halo ? 8 : (H == 1 && kh == 1) ? 128 : 16;
        t.tile_h = 128 / t.tile_w;
        t.tiles_x = (W + t.tile_w - 1) / t.tile_w;
        t.tiles_y = (H + t.tile_h - 1) / t.tile_h;
        t.num_m_tiles = n_img * t.tiles_x * t.tiles_y;
        cuuint64_t dims[4] = {cuuint64_t(cin), cuuint64_t(W), cuuint64_t(H), cuuint64_t(n_img)};
        cuuint64_t strides[3] = {cuuint64_t(in_cs) * es, cuuint64_t(W) * in_cs * es, cuuint64_t(H) * W * in_cs * es};
        // rowbox: one box of 8 + kh - 1 rows per (kx, k-block) instead of one 8-row box per tap — kh x less L2->smem
        // traffic for the activations.  Needs >= 3 pipeline stages of (box + kh weight slices) in shared memory.
        // Kernels too tall for that (9x9 with streamed weights: 9 weight slices per stage) split their vertical taps into equal
        // groups of kh_g, one box of 8 + kh_g - 1 rows each: still kh_g x less activation traffic than per-tap boxes.
        const int budget_rb = kMaxSmem - slack - (w.split ? 2 : kOutBufs) * kOutBufBytes;
        t.kh_g = 0;
        if (!t.halo && t.tile_w == 16 && allow_rowbox && kh > 1) {
            for (int groups = 1; groups < kh && !t.kh_g; groups++) {
                const int g = (kh + groups - 1) / groups;
                if (g < 2) break;
                const int a_box = (8 + g - 1) * 16 * 128;
                const int rb_need = t.b_resident ? 3 * a_box + b_all : 3 * (a_box + g * w.n_chunk * 128);
                if (rb_need <= budget_rb) t.kh_g = g;
            }
        }
        t.rowbox = t.kh_g ? 1 : 0;
        cuuint32_t box[4] = {kbe, cuuint32_t(t.halo ? 8 + kw - 1 : t.tile_w), cuuint32_t(t.halo ? 16 + kh - 1 : t.rowbox ? 8 + t.kh_g - 1 : t.tile_h), 1};
        err = encode(&t.map_a, const_cast<void*>(in), 4, dims, strides, box, a32);
    }
    if (!err.empty()) return err;
    {
        // stacked split: [2 n_chunk rows (hi | lo)] x [32 columns = 64 bytes per k-block], SWIZZLE_64B
        cuuint64_t dims[2] = {cuuint64_t(w.taps) * w.k_pad, cuuint64_t(w.n_chunks) * w.n_chunk * (t.stack ? 2 : 1)};
        cuuint64_t strides[1] = {cuuint64_t(w.taps) * w.k_pad * es_b};
        cuuint32_t box[2] = {cuuint32_t((t.stack ? 64 : 128) / es_b), cuuint32_t(w.n_chunk * (t.stack ? 2 : 1))};
        err = encode(&t.map_b, const_cast<void*>(wdev), 2, dims, strides, box, w.tf32, t.stack != 0);
        if (!err.empty()) return err;
    }
    t.valid = true;
    return "";
}

std::string tc_conv_setup_dwpw(TcConv& t, const void* in, int in_cs, int cin, const void* wdev, const TcWeights& w, int n_img, int H, int W,
                               int dw) {
    t.valid = false;
    if (!w.split || w.stack || w.n_chunks != 1 || w.taps != 1) return "fused depthwise needs a single-chunk split-mode 1x1 convolution";
    if (dw != 3 && dw != 5) return "fused depthwise: 3x3 or 5x5 only";
    if ((reinterpret_cast<uintptr_t>(in) & 15) || (in_cs & 3)) return "activation view not 16-byte aligned";
    t.kh = t.kw = 1;
    t.ph = t.pw = dw / 2;             // offsets of the input window (the depthwise padding)
    t.dw = dw;
    t.k_pad = w.k_pad;
    t.cin = cin;
    t.rowbox = 0; t.halo = 0; t.kh_g = 0; t.pack = 0; t.direct1 = 0;
    t.tf32 = 0; t.split = 1; t.stack = 0;
    t.w_scale = w.w_scale;
    t.num_kb = (cin + 31) / 32;
    t.n_chunk = w.n_chunk;
    t.n_chunks = 1;
    const int b_all = t.num_kb * w.n_chunk * 128;
    t.b_resident = b_all <= kBResidentMaxSplit ? 1 : 0;
    t.spatial = 1;
    t.n_img = n_img; t.H = H; t.W = W;
    t.tile_w = 8; t.tile_h = 16;
    t.tiles_x = (W + 7) / 8;
    t.tiles_y = (H + 15) / 16;
    t.num_m_tiles = n_img * t.tiles_x * t.tiles_y;
    // two stages of (window + operand tile + streamed weight slice) must fit next to the resident weights and two store tiles
    const int win = round_up_i((16 + dw - 1) * (8 + dw - 1) * 128, 1024);
    const int stage = win + A_BYTES + (t.b_resident ? 0 : w.n_chunk * 128);
    const int dw_param = (dw * dw + 3) * t.num_kb * 32 * 4;
    if (2 * stage + (t.b_resident ? b_all : 0) + 2 * kOutBufBytes + dw_param + 3 * w.n_chunk * 4 + 1024 + kBarRegion > kMaxSmem) return "fused depthwise: stages do not fit";
    cuuint64_t dims[4] = {cuuint64_t(cin), cuuint64_t(W), cuuint64_t(H), cuuint64_t(n_img)};
    cuuint64_t strides[3] = {cuuint64_t(in_cs) * 4, cuuint64_t(W) * in_cs * 4, cuuint64_t(H) * W * in_cs * 4};
    cuuint32_t box[4] = {32, cuuint32_t(8 + dw - 1), cuuint32_t(16 + dw - 1), 1};
    std::string err = encode(&t.map_a, const_cast<void*>(in), 4, dims, strides, box, true);
    if (!err.empty()) return err;
    {
        cuuint64_t bdims[2] = {cuuint64_t(w.k_pad), cuuint64_t(w.n_chunk)};
        cuuint64_t bstrides[1] = {cuuint64_t(w.k_pad) * 2};
        cuuint32_t bbox[2] = {64, cuuint32_t(w.n_chunk)};
        err = encode(&t.map_b, const_cast<void*>(wdev), 2, bdims, bstrides, bbox, false);
        if (!err.empty()) return err;
    }
    t.valid = true;
    return "";
}

// output tensor map (TMA stores): [n_store channels] x pixels, row pitch out_cs; re-encoded only when the target changes
static std::string tc_output_map(TcConv& t, bool* changed = nullptr) {
    if (t.map_o_ptr == t.out && t.map_o_cs == t.out_cs && t.map_o_n == t.n_store) return "";
    if (changed) *changed = true;
    std::string err;
    const int es = (t.tf32 || t.split) ? 4 : 2;
    const cuuint32_t cols = cuuint32_t(128 / es);   // one staging row = 64 fp16 or 32 fp32 columns
    if (t.spatial) {
        cuuint64_t dims[4] = {cuuint64_t(t.n_store), cuuint64_t(t.W), cuuint64_t(t.H), cuuint64_t(t.n_img)};
        cuuint64_t strides[3] = {cuuint64_t(t.o_px ? t.o_px : t.out_cs) * es, cuuint64_t(t.o_row ? t.o_row : (long long)t.W * t.out_cs) * es,
                                 cuuint64_t(t.o_img ? t.o_img : (long long)t.H * t.W * t.out_cs) * es};
        cuuint32_t box[4] = {cols, cuuint32_t(t.tile_w), cuuint32_t(t.tile_h), 1};
        err = encode(&t.map_o, t.out, 4, dims, strides, box, es == 4);
    } else {
        cuuint64_t dims[2] = {cuuint64_t(t.n_store), cuuint64_t(t.M)};
        cuuint64_t strides[1] = {cuuint64_t(t.out_cs) * es};
        cuuint32_t box[2] = {cols, BLOCK_M};
        err = encode(&t.map_o, t.out, 2, dims, strides, box, es == 4);
    }
    if (err.empty()) { t.map_o_ptr = t.out; t.map_o_cs = t.out_cs; t.map_o_n = t.n_store; }
    return err;
}

static std::string launch_impl(TcConv& t, int sm_count, cudaStream_t st, const CUtensorMap* gmaps, const TcGroupDev* gtab,
                               int n_groups, int total_m_tiles);

std::string launch_conv_tc(TcConv& t, int sm_count, cudaStream_t st) {
    const bool of32 = t.tf32 || t.split;
    if (t.direct1) {
        if (!of32 || t.spatial || t.pack || t.n_store != 1 || t.n_chunks != 1) return "direct single-channel output needs a flat fp32 convolution";
        t.map_o = t.map_a;      // never used for a store; the kernel only prefetches the descriptor
        return launch_impl(t, sm_count, st, nullptr, nullptr, 0, t.num_m_tiles);
    }
    if ((reinterpret_cast<uintptr_t>(t.out) & 15) || (t.out_cs & (of32 ? 3 : 7))) return "output view not 16-byte aligned";
    {
        std::string err = tc_output_map(t);
        if (!err.empty()) return err;
    }
    return launch_impl(t, sm_count, st, nullptr, nullptr, 0, t.num_m_tiles);
}

// Ragged batch: the groups (runs of equal-sized images, each with its own tensor maps) of one KxK convolution in ONE launch.
// `dev` receives [2 maps per group][TcGroupDev per group] (tc_groups_dev_bytes(n) bytes, 128-byte aligned); `uploaded` caches
// that the table on the device is current (same context, same output views).
size_t tc_groups_dev_bytes(int n_groups) { return size_t(n_groups) * (2 * sizeof(CUtensorMap) + sizeof(TcGroupDev)) + 256; }

std::string launch_conv_tc_groups(TcConv* const* gs, const long long* pix_off, int n_groups, void* dev, bool* uploaded, int sm_count,
                                  cudaStream_t st) {
    if (n_groups <= 0) return "no groups";
    const bool of32 = gs[0]->tf32 || gs[0]->split;
    bool changed = !*uploaded;
    for (int g = 0; g < n_groups; g++) {
        TcConv& t = *gs[g];
        if (!t.spatial || t.halo != gs[0]->halo || t.rowbox != gs[0]->rowbox || t.kh_g != gs[0]->kh_g || t.tile_w != gs[0]->tile_w || t.b_resident != gs[0]->b_resident || t.num_kb != gs[0]->num_kb)
            return "groups with different kernel shapes";
        if ((reinterpret_cast<uintptr_t>(t.out) & 15) || (t.out_cs & (of32 ? 3 : 7))) return "output view not 16-byte aligned";
        std::string err = tc_output_map(t, &changed);
        if (!err.empty()) return err;
    }
    static_assert(sizeof(CUtensorMap) == 128, "CUtensorMap size");
    const size_t map_bytes = size_t(n_groups) * 2 * sizeof(CUtensorMap);
    int tiles = 0;
    if (changed) {
        std::vector<unsigned char> host(map_bytes + size_t(n_groups) * sizeof(TcGroupDev));
        for (int g = 0; g < n_groups; g++) {
            const TcConv& t = *gs[g];
            std::memcpy(host.data() + size_t(2 * g) * sizeof(CUtensorMap), &t.map_a, sizeof(CUtensorMap));
            std::memcpy(host.data() + size_t(2 * g + 1) * sizeof(CUtensorMap), &t.map_o, sizeof(CUtensorMap));
            TcGroupDev d{tiles, t.n_img, t.H, t.W, t.tiles_x, t.tiles_y, pix_off[g]};
            std::memcpy(host.data() + map_bytes + size_t(g) * sizeof(TcGroupDev), &d, sizeof(d));
            tiles += t.num_m_tiles;
        }
        launch_upload(dev, host.data(), host.size(), st);
        *uploaded = true;
    } else {
        for (int g = 0; g < n_groups; g++) tiles += gs[g]->num_m_tiles;
    }
    const CUtensorMap* gmaps = static_cast<const CUtensorMap*>(dev);
    const TcGroupDev* gtab = reinterpret_cast<const TcGroupDev*>(static_cast<const char*>(dev) + map_bytes);
    return launch_impl(*gs[0], sm_count, st, gmaps, gtab, n_groups, tiles);
}

static std::string launch_impl(TcConv& t, int sm_count, cudaStream_t st, const CUtensorMap* gmaps, const TcGroupDev* gtab,
                               int n_groups, int total_m_tiles) {
    const bool of32 = t.tf32 || t.split;
    TcParams p{};
    p.gmaps = gmaps; p.gtab = gtab; p.n_groups = n_groups;
    p.spatial = t.spatial; p.M = t.M; p.n_img = t.n_img; p.H = t.H; p.W = t.W; p.tiles_x = t.tiles_x; p.tiles_y = t.tiles_y;
    p.kh = t.kh; p.kw = t.kw; p.ph = t.ph; p.pw = t.pw; p.num_kb = t.num_kb; p.k_pad = t.k_pad;
    p.n_chunk = t.n_chunk; p.n_chunks = t.n_chunks; p.n_store = t.n_store; p.num_m_tiles = total_m_tiles;
    p.cin = t.cin;
    p.tf32 = t.tf32;
    p.split = t.split;
    p.kb_elems = of32 ? 32 : BLOCK_K;              // activation elements per k-block
    p.kbb = t.split ? (t.stack ? 32 : 64) : p.kb_elems;   // weight columns per k-block
    p.stack = t.stack;
    p.direct1 = t.direct1;
    p.acc_cols = t.n_chunk * (t.stack ? 2 : 1);
    p.rowbox = t.rowbox;
    p.kh_g = t.rowbox ? t.kh_g : 0;
    p.halo = t.halo;
    p.tile_h = t.tile_h;
    p.tw_shift = t.tile_w == 8 ? 3 : t.tile_w == 16 ? 4 : 7;
    p.a_tx = t.dw ? (16 + t.dw - 1) * (8 + t.dw - 1) * 128 : t.halo ? (16 + t.kh - 1) * (8 + t.kw - 1) * 128 : t.rowbox ? (8 + t.kh_g - 1) * 16 * 128 : A_BYTES;
    p.dw = t.dw; p.dw_cp = t.dw_cp; p.dw_act = t.dw_act;
    p.dw_w = t.dw_w; p.dw_bias = t.dw_bias; p.dw_ps = t.dw_ps; p.dw_pt = t.dw_pt;
    if (t.dw && (!t.split || !t.dw_w || !t.dw_bias)) return "fused depthwise: parameters missing";
    p.a_bytes = round_up_i(p.a_tx, 1024);
    p.n_total = t.n_chunk * t.n_chunks;
    p.param_smem = p.n_total <= kParamSmemMaxCh ? 1 : 0;
    const int param_bytes = (p.param_smem ? 3 * p.n_total * int(sizeof(float)) : 0) + (t.dw ? (t.dw * t.dw + 3) * t.num_kb * 32 * int(sizeof(float)) : 0);
    p.b_resident = t.b_resident;
    p.b_total = t.b_resident ? t.kh * t.kw * t.num_kb * t.n_chunk * 128 : 0;
    const int stage_bytes = p.a_bytes + (t.dw ? A_BYTES : 0) + (p.b_resident ? 0 : t.n_chunk * 128 * (t.halo ? t.kh * t.kw : t.rowbox ? t.kh_g : 1));
    // staging ring of the TMA stores: 4 tiles; split mode (operand stages and weights are twice the bytes) gives tiles back
    // to the operand pipeline until it holds 3 stages
    p.out_bufs = kOutBufs;
    int budget = kMaxSmem - 1024 - kBarRegion - param_bytes - p.b_total - p.out_bufs * kOutBufBytes;
    while (t.split && p.out_bufs > 2 && budget / stage_bytes < 3) {
        p.out_bufs--;
        budget += kOutBufBytes;
    }
    p.stages = std::max(2, std::min(8, budget / stage_bytes));
    // TMEM accumulator ring: the MMA warp runs up to acc_stages tiles ahead of the epilogue (hides the commit -> wait ->
    // tcgen05.ld -> arrive round trip, which dominates layers with one k-iteration per tile)
    p.acc_stages = std::max(2, std::min(kMaxAccStages, 512 / p.acc_cols));
    int cols = 32;
    while (cols < p.acc_stages * p.acc_cols) cols *= 2;
    p.tmem_cols = cols;
    p.out = t.out; p.out_cs = t.out_cs;
    p.bias = t.epi.bias; p.post_scale = t.epi.post_scale; p.post_shift = t.epi.post_shift;
    p.res = t.epi.res; p.res_cs = t.epi.res_cs; p.act = t.epi.act; p.act2 = t.epi.act2;
    p.hs_slope = t.epi.hs_slope; p.hs_offset = t.epi.hs_offset;
    p.gate = t.epi.gate; p.gate_c = t.epi.gate_c; p.gate_rows = std::max(t.epi.gate_rows, 1);
    p.a_scale = t.split ? (t.a_scale > 0.f ? t.a_scale : tc_split_activation_scale()) : 1.f;
    p.acc_scale = t.split ? 1.f / (p.a_scale * t.w_scale) : 1.f;
    const size_t smem = size_t(p.stages) * stage_bytes + p.b_total + p.out_bufs * kOutBufBytes + 1024 + kBarRegion + param_bytes;
    if (smem > size_t(kMaxSmem)) return "operand stages do not fit shared memory";
    static bool configured[64] = {};   // per device: the attribute belongs to the device's copy of the kernel
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        cudaFuncSetAttribute(conv_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
        cudaFuncSetAttribute(conv_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    const int total = total_m_tiles * t.n_chunks;
    const int grid = std::max(1, std::min(total, sm_count));
    if (t.split) pdl_launch(conv_tc_kernel<true>, grid, kThreads, smem, st, t.map_a, t.map_b, t.map_o, p);
    else pdl_launch(conv_tc_kernel<false>, grid, kThreadsBase, smem, st, t.map_a, t.map_b, t.map_o, p);
    return "";
}

}  // namespace vse
