// lstm.cu — recurrent half of a (bi)LSTM layer (reference backend/models/V2/ch_rec, Paddle `rnn` op#140, mode=LSTM,
// hidden 256, 2 layers, bidirectional).  The input GEMM W_ih x_t + b_ih + b_hh of every time step and
// both directions is a plain 1x1 CONV step (plan.py::_op_rnn) and runs on the tensor-core conv path; this kernel walks the
// time axis:   g_t = xg_t + W_hh h_{t-1};  c_t = s(f) c_{t-1} + s(i) tanh(g);  h_t = s(o) tanh(c_t)   (gate order i,f,g,o).
//
// B200 mapping: one thread-block CLUSTER of 8 CTAs per (direction, group of 4 text lines).  W_hh of one direction
// (4H x H fp32 = 1 MB) does not fit one SM, so CTA r of the cluster keeps the 4*H/8 gate rows of its H/8 hidden units
// resident in shared memory (128 KB, loaded once) for the whole sequence; per time step each CTA computes its rows for
// the 4 lines from the full h_{t-1} in its own shared memory, updates its slice of c/h, and pushes the new h slice into
// the (double-buffered) h vector of all 8 CTAs through distributed shared memory with st.async: the 16-byte stores
// complete a transaction count on the RECEIVER's mbarrier, so a CTA starts step t+1 as soon as the 4 KB of h_t have
// landed in its own shared memory — no cluster-wide barrier and no memory fence in the loop (a cluster barrier's release
// fence also waits for the h_t stores to HBM: 45 % of the stall samples of the first version).  Nothing but the gate
// pre-activations (read once) and h_t (written once) touches HBM inside the time loop.
#include <cooperative_groups.h>
#include <cuda_fp16.h>

#include "nn_kernels.h"

namespace cg = cooperative_groups;

namespace vse {

namespace {

constexpr int kCL = 8;     // CTAs per cluster
constexpr int kSeq = 4;    // text lines per cluster
constexpr int kThreads = 256;

__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(__half v) { return __half2float(v); }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint32_t map_to_cta(uint32_t addr, uint32_t rank) {   // shared::cta address -> shared::cluster address of CTA `rank`
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 16-byte store into another CTA's shared memory that completes 16 bytes on that CTA's mbarrier
__device__ __forceinline__ void st_async_16(uint32_t remote_addr, float4 v, uint32_t remote_bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(remote_addr),
                 "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "r"(remote_bar)
                 : "memory");
}
__device__ __forceinline__ void stf(float* p, float v) { *p = v; }
__device__ __forceinline__ void stf(__half* p, float v) { *p = __float2half_rn(v); }

__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }

// H = hidden size (multiple of 64, <= 256).  Shared memory: Wt [H][4*HU] | hT [2][H][kSeq] | part [4][4*HU][kSeq] | hst [HU][kSeq] | hbar [2]
template <typename T, int H>
__global__ void __launch_bounds__(kThreads, 1) lstm_recurrent_kernel(const T* __restrict__ gates, int gates_cs, T* __restrict__ out,
                                                                      int out_cs, const float* __restrict__ w_packed,
                                                                      const ImgTab* __restrict__ tab, int n_img) {
    constexpr int HU = H / kCL;       // hidden units owned by this CTA
    constexpr int R = 4 * HU;         // gate rows owned by this CTA
    constexpr int KQ = H / 4;         // k range per warp quarter
    static_assert(R == 128 && HU == 32, "thread mapping below assumes 128 gate rows per CTA");
    extern __shared__ __align__(16) unsigned char lstm_smem[];
    float* Wt = reinterpret_cast<float*>(lstm_smem);            // [H][R]
    float* hT = Wt + size_t(H) * R;                             // [2][H][kSeq]
    float* part = hT + 2 * H * kSeq;                            // [4][R][kSeq]
    float* hst = part + 4 * R * kSeq;                           // [HU][kSeq]
    uint64_t* hbar = reinterpret_cast<uint64_t*>(hst + HU * kSeq);   // [2]: "h[buf] of the next step has arrived"
    constexpr uint32_t kPushBytes = kCL * HU * kSeq * sizeof(float);   // what the 8 CTAs push into one h buffer per step

    cg::cluster_group cluster = cg::this_cluster();
    const int rank = int(cluster.block_rank());
    const int group = blockIdx.y, dir = blockIdx.z;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // resident weights: this CTA's slice was packed contiguously by the host ([dir][rank][k][R])
    {
        const float4* src = reinterpret_cast<const float4*>(w_packed + (size_t(dir) * kCL + rank) * H * R);
        float4* dst = reinterpret_cast<float4*>(Wt);
        for (int i = tid; i < H * R / 4; i += kThreads) dst[i] = __ldg(src + i);
    }
    for (int i = tid; i < 2 * H * kSeq; i += kThreads) hT[i] = 0.f;
    if (tid == 0) {
        mbar_init(&hbar[0], 1);
        mbar_init(&hbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_arrive_expect_tx(&hbar[1], kPushBytes);   // step 0 pushes h_0 into buffer 1
    }

    // the (up to) 4 lines of this group
    int seq_off[kSeq], seq_T[kSeq], max_T = 0;
#pragma unroll
    for (int s = 0; s < kSeq; s++) {
        const int img = group * kSeq + s;
        seq_off[s] = 0;
        seq_T[s] = 0;
        if (img < n_img) {
            const ImgTab t = tab[img];
            seq_off[s] = t.off;
            seq_T[s] = t.h * t.w;   // H == 1 views: the padded width is the sequence
        }
        max_T = max(max_T, seq_T[s]);
    }
    // phase B role: thread (u, s) owns hidden unit rank*HU + u of line s
    const int bu = tid & 31, bs = tid >> 5;
    int my_off = 0, my_T = 0;
#pragma unroll
    for (int s = 0; s < kSeq; s++)
        if (bs == s) { my_off = seq_off[s]; my_T = seq_T[s]; }
    float c_state = 0.f, h_state = 0.f;
    // phase A role: warp -> (k quarter, row half), lane -> two adjacent gate rows
    const int kq = warp & 3, r0 = (warp >> 2) * 64 + lane * 2;

    cluster.sync();   // weights, zeroed h and armed barriers visible; every CTA of the cluster is running (DSMEM is addressable)

    // phase C role: thread (u, dest) pushes unit u of this CTA's slice into CTA dest
    const int cu = tid & 31, cdest = tid >> 5;
    const uint32_t push_dst = map_to_cta(smem_u32(hT + (size_t(rank) * HU + cu) * kSeq), uint32_t(cdest));
    const uint32_t push_bar = map_to_cta(smem_u32(hbar), uint32_t(cdest));

    for (int t = 0; t < max_T; t++) {
        const int cur = t & 1, nxt = cur ^ 1;
        // gate pre-activations of this step (independent of h: in flight during the wait and phase A; converted in phase B)
        T xg[4];
        int pixel = 0;
        const bool active = tid < HU * kSeq && t < my_T;
        if (active) {
            pixel = my_off + (dir == 0 ? t : my_T - 1 - t);
            const T* gp = gates + size_t(pixel) * gates_cs + dir * 4 * H + rank * HU + bu;
#pragma unroll
            for (int g = 0; g < 4; g++) xg[g] = __ldg(gp + g * H);
        }
        // h_{t-1} of all 8 CTAs has landed in h[cur] (pushed during step t-1; each barrier completes every other step)
        if (t > 0) mbar_wait_cluster(&hbar[cur], uint32_t((t - 1) >> 1) & 1u);
        // phase A: partial dot products  W_slice[r][k] * h[k][s]  over this warp's quarter of k
        {
            float acc0[kSeq] = {0.f, 0.f, 0.f, 0.f}, acc1[kSeq] = {0.f, 0.f, 0.f, 0.f};
            const float* wp = Wt + size_t(kq * KQ) * R + r0;
            const float4* hp = reinterpret_cast<const float4*>(hT + (size_t(cur) * H + kq * KQ) * kSeq);
#pragma unroll 8
            for (int k = 0; k < KQ; k++) {
                const float2 w = *reinterpret_cast<const float2*>(wp + size_t(k) * R);
                const float4 h = hp[k];
                acc0[0] = fmaf(w.x, h.x, acc0[0]); acc0[1] = fmaf(w.x, h.y, acc0[1]);
                acc0[2] = fmaf(w.x, h.z, acc0[2]); acc0[3] = fmaf(w.x, h.w, acc0[3]);
                acc1[0] = fmaf(w.y, h.x, acc1[0]); acc1[1] = fmaf(w.y, h.y, acc1[1]);
                acc1[2] = fmaf(w.y, h.z, acc1[2]); acc1[3] = fmaf(w.y, h.w, acc1[3]);
            }
            float4* pp = reinterpret_cast<float4*>(part + (size_t(kq) * R + r0) * kSeq);
            pp[0] = make_float4(acc0[0], acc0[1], acc0[2], acc0[3]);
            pp[1] = make_float4(acc1[0], acc1[1], acc1[2], acc1[3]);
        }
        __syncthreads();
        // phase B: gates -> c, h for (unit bu, line bs)
        if (tid < HU * kSeq) {
            if (active) {
                float pre[4];
#pragma unroll
                for (int g = 0; g < 4; g++) {
                    const int r = g * HU + bu;
                    float v = part[(0 * R + r) * kSeq + bs];
                    v += part[(1 * R + r) * kSeq + bs];
                    v += part[(2 * R + r) * kSeq + bs];
                    v += part[(3 * R + r) * kSeq + bs];
                    pre[g] = to_f(xg[g]) + v;
                }
                c_state = sigmoid_f(pre[1]) * c_state + sigmoid_f(pre[0]) * tanhf(pre[2]);
                h_state = sigmoid_f(pre[3]) * tanhf(c_state);
                stf(out + size_t(pixel) * out_cs + dir * H + rank * HU + bu, h_state);
            }
            hst[bu * kSeq + bs] = h_state;
        }
        __syncthreads();
        // phase C: push this CTA's h slice into h[nxt] of every CTA (itself included); the receiver's barrier counts the bytes.
        // Write-after-read: a receiver still reads h[nxt] only during ITS step t-1, and we are in step t because its
        // step t-1 push (issued after that read) has arrived here.  Nothing is pushed after the last step, so no store
        // is in flight towards a CTA that has left the loop.
        if (t + 1 < max_T) {
            if (tid == 0) mbar_arrive_expect_tx(&hbar[cur], kPushBytes);   // arms the barrier of step t+1's pushes (into h[cur])
            const float4 v = *reinterpret_cast<const float4*>(hst + cu * kSeq);
            st_async_16(push_dst + uint32_t(nxt) * uint32_t(H * kSeq * sizeof(float)), v, push_bar + uint32_t(nxt) * 8u);
        }
    }
    cluster.sync();   // no CTA of the cluster exits while a sibling could still address its shared memory
}

}  // namespace

size_t lstm_packed_weight_floats(int hidden, int ndir) { return size_t(ndir) * 4 * hidden * hidden; }

// w_hh [ndir][4*hidden][hidden] (gate order i,f,g,o) -> [dir][rank][k][g*HU + u] so that each CTA's resident slice is one
// contiguous block, k-major (conflict-free shared-memory reads with one row pair per lane)
void lstm_pack_weights(const float* w_hh, int hidden, int ndir, float* dst) {
    const int HU = hidden / kCL, R = 4 * HU;
    for (int d = 0; d < ndir; d++)
        for (int rank = 0; rank < kCL; rank++)
            for (int k = 0; k < hidden; k++)
                for (int g = 0; g < 4; g++)
                    for (int u = 0; u < HU; u++)
                        dst[((size_t(d) * kCL + rank) * hidden + k) * R + g * HU + u] =
                            w_hh[(size_t(d) * 4 * hidden + g * hidden + rank * HU + u) * hidden + k];
}

bool lstm_supported(int hidden) { return hidden == 256; }

template <typename T>
static cudaError_t launch_lstm_t(const void* gates, int gates_cs, void* out, int out_cs, const float* w_packed, const ImgTab* tab,
                                 int n_img, int ndir, cudaStream_t st) {
    constexpr int H = 256, HU = H / kCL, R = 4 * HU;
    const size_t smem = (size_t(H) * R + 2 * H * kSeq + 4 * R * kSeq + HU * kSeq) * sizeof(float) + 2 * sizeof(uint64_t);
    auto kern = lstm_recurrent_kernel<T, H>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(kCL, (n_img + kSeq - 1) / kSeq, ndir);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kCL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<const T*>(gates), gates_cs, static_cast<T*>(out), out_cs, w_packed, tab, n_img);
}

// returns cudaSuccess or the launch error; the caller checked lstm_supported(hidden)
cudaError_t launch_lstm(const void* gates, int gates_cs, void* out, int out_cs, const float* w_packed, const ImgTab* tab, int n_img,
                        int ndir, int prec, cudaStream_t st) {
    if (n_img <= 0) return cudaSuccess;
    return prec == 0 ? launch_lstm_t<__half>(gates, gates_cs, out, out_cs, w_packed, tab, n_img, ndir, st)
                     : launch_lstm_t<float>(gates, gates_cs, out, out_cs, w_packed, tab, n_img, ndir, st);
}

}  // namespace vse
