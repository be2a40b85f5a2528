// engine.h — internal C++ interface of the B200 engine (plan executor + frame pipeline).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/vse_b200.h"
#include "gemm_tc.h"
#include "nn_kernels.h"
#include "plan.h"
#include "postproc.cuh"

namespace vse {

struct CudaError {
    std::string msg;
};
#define VSE_CUDA(expr)                                                                                    \
    do {                                                                                                  \
        cudaError_t _e = (expr);                                                                          \
        if (_e != cudaSuccess)                                                                            \
            throw ::vse::CudaError{std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" + __FILE__ + ":" + \
                                   std::to_string(__LINE__) + ")"};                                       \
    } while (0)

struct InvalidArg {
    std::string msg;
};
struct CapacityError {
    std::string msg;
};
struct StateError {
    std::string msg;
};

// growable device buffer
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    void reserve(size_t n) {
        if (n <= cap) return;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = n + (n >> 3) + 256;
        VSE_CUDA(cudaMalloc(&p, want));
        cap = want;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T> T* as() const { return static_cast<T*>(p); }
};

struct PinnedBuf {
    void* p = nullptr;
    size_t cap = 0;
    void reserve(size_t n) {
        if (n <= cap) return;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        VSE_CUDA(cudaMallocHost(&p, n + (n >> 3) + 256));
        cap = n + (n >> 3) + 256;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T> T* as() const { return static_cast<T*>(p); }
};

// device-side parameters of one step, in kernel-ready layout
struct StepDev {
    const float* w = nullptr;
    const float* bias = nullptr;
    const float* post_scale = nullptr;
    const float* post_shift = nullptr;
    const float* gamma = nullptr;
    const float* beta = nullptr;
    const float* scale = nullptr;
    const float* shift = nullptr;
    const void* w_tc = nullptr;  // half, K-major [cout_pad][k_pad] for the tensor-core path
    // pixel-packed variant of a 1x1 conv (see Engine::load_plan): per-channel constants repeated `pack` times
    const float* w_t = nullptr;  // VECLIN: transposed (input-major) matrix for the squeeze-excite gate kernel
    const float* bias_pk = nullptr;
    const float* post_scale_pk = nullptr;
    const float* post_shift_pk = nullptr;
    int pack = 0;                // pixels per packed GEMM row (0 = no packed variant)
    int w_ci = 0, w_co = 0, k_pad = 0, n_pad = 0;
};

struct Geo {
    std::vector<ImgTab> tab;
    int64_t total = 0;
    int max_pix = 0;
    size_t tab_off = 0;  // offset (in ImgTab units) inside the context's device table
};

struct ValueRt {
    int geo = -1;       // index into ExecContext::geos (images) ; -1 for vectors
    size_t off = 0;     // byte offset of the ROOT buffer in the arena
    size_t bytes = 0;   // size of the root buffer
    bool live = false;
};

struct LoadedPlan {
    PlanData data;
    std::vector<StepDev> dev;
    DevBuf weights;  // all device parameters
    // tensor-core path: fp16 K-major weight matrices of the CONV steps (gemm_tc.cu)
    std::vector<TcWeights> tcw;      // per step (b emptied after upload; n_chunk == 0 => not packed)
    std::vector<size_t> tcw_off;     // byte offset into tc_weights
    std::vector<TcWeights> tcw_pk;   // pixel-packed block-diagonal variants (n_chunk == 0 => none)
    std::vector<size_t> tcw_pk_off;
    DevBuf tc_weights;
    // 2x2 stride-2 transposed convolutions on the tensor-core path: one weight matrix [cout][cin] per output position
    std::vector<TcWeights> tcw_dc;   // [step * 4 + pos] (n_chunk == 0 => not packed)
    std::vector<size_t> tcw_dc_off;
    std::vector<float> a_scale;      // per step: power of two applied to a convolution's operand rows in split mode
    bool loaded = false;
};

// one execution of a plan on a concrete batch geometry
struct ExecContext {
    std::vector<Geo> geos;
    std::vector<ValueRt> vals;
    DevBuf tabs;      // all ImgTab tables
    size_t arena_bytes = 0;
    size_t scratch_off = 0, scratch_bytes = 0;
    int n_img = 0;
    std::vector<TcConv> tc;   // per step; valid => the step runs on the tcgen05 kernel for this geometry
    // KxK convolutions over a RAGGED batch (recogniser crops of different padded widths): one tensor-core launch per run
    // of consecutive equal-sized images (the reference pads every <= 6-crop batch to one width, so a ragged batch is a
    // short list of uniform groups).  Non-empty => the step runs as these launches.
    struct TcGroup {
        TcConv tc;
        int64_t pix_off = 0;   // first pixel of the group in the step's input / output / residual values
    };
    std::vector<std::vector<TcGroup>> tc_groups;
    std::vector<TcConv> tc_dc;      // [step * 4 + pos]: the four 1x1 launches of a transposed convolution (valid => tensor cores)
    DevBuf tc_gdev;                 // per ragged step: the groups' tensor maps + geometry table on the device (gemm_tc.h)
    std::vector<size_t> tc_goff;    // byte offset of a step's table in tc_gdev
    std::vector<char> tc_gup;       // 1 = the table on the device is current
    // per step: 1 = depthwise convolution computed inside the following 1x1 convolution's kernel (gemm_tc.h, TcConv::dw): the
    // depthwise output is never materialised; the fused launch happens at the depthwise step, the 1x1 step is skipped
    std::vector<char> dwpw;
    std::vector<char> se_conv; // per step: 1 = 1x1 conv heading a fused residual squeeze-excite group (build_context)
    std::vector<int> kind;    // per step, filled by exec_steps: which kernel family ran (see Engine::time_steps)
};

class Engine {
  public:
    explicit Engine(const vse_config& cfg);
    ~Engine();
    void load_plan(int which, const void* blob, size_t n);
    void set_conv_input_ranges(int which, const float* absmax, int n_steps);

    // Builds geometry + memory plan for `which` on images (h, w[i], valid_w[i]) and runs it.
    // Input pixels (uint8 BGRX) must already be in `input_dev` laid out image after image.
    void run_plan(int which, const std::vector<ImgTab>& in_tab, const uint8_t* input_dev, bool keep_all);
    int64_t get_value(int which, int vid, float* out, int64_t cap, int32_t* channels);
    const void* value_ptr(int which, int vid, int* cs, const Geo** geo);

    void prefetch_frames(const uint8_t* const* frames, const int32_t* h, const int32_t* w, const int32_t* stride, int n,
                         int mem_kind);
    void run_frames(const uint8_t* const* frames, const int32_t* h, const int32_t* w, const int32_t* stride, int n,
                    int mem_kind, vse_result* out, bool det_only);

    // debug hooks
    int time_steps(int which, int reps, float* ms, int64_t* info, int cap);
    void debug_run_plan(int which, const uint8_t* const* images, int n, int h, const int32_t* w, const int32_t* valid_w,
                        bool keep_all);
    void debug_resize(const uint8_t* src, int sh, int sw, int stride, uint8_t* dst, int dh, int dw);
    void debug_db_post(const float* prob, int rh, int rw, int src_h, int src_w, float* quads, float* scores, int cap,
                       int* n_out);
    void debug_crop(const uint8_t* frame, int h, int w, const float* quad, uint8_t* out, int cap, int* oh, int* ow);

    // pipeline state (pipeline.cu)
    struct Pipeline;
    void ensure_pipeline();
    // DB post-process of a device-resident probability map; results land in the pipeline's pinned h_out buffer
    void db_post_device(const float* prob, const std::vector<DetFrame>& frames, bool reading_order);

    std::string last_error;
    int64_t launches = 0;
    vse_config cfg;
    cudaStream_t stream = nullptr;
    int sm_count = 148;
    int64_t tc_launches = 0;
    // Recogniser plans end in SOFTMAX over the class logits, whose only reader is the CTC decode: run_frames decodes straight from
    // the logits (postproc.cu, ctc_decode_kernel) and exec_steps skips the step while this is set.  logits_vid(): the value that
    // feeds that SOFTMAX when the fold applies to the loaded plan (float activations, single reader), else -1
    bool fold_final_softmax = false;
    int logits_vid(int which) const;

  private:
    void prepare_plan(int which, LoadedPlan& lp);
    void build_context(int which, const std::vector<ImgTab>& in_tab, bool keep_all);
    void exec_steps(int which, std::vector<cudaEvent_t>* step_events = nullptr);
    size_t elt_size(int which, const ValueRec& v) const;
    int plan_prec_[2] = {0, 0};   // VSE_PRECISION_* per plan (see load_plan)
    int value_cs(const PlanData& pd, int vid) const;  // channel stride in elements
    void* vptr(int which, int vid) const;
    bool launch_conv(int which, int step, const ConvArgs& a, int prec);   // true: ran on the tcgen05 kernel
    const uint8_t* input_ptr_[2] = {nullptr, nullptr};

    LoadedPlan plans_[2];
    ExecContext ctx_[2];
    DevBuf arena_[2];
    DevBuf dbg_;
    PinnedBuf pin_;

    Pipeline* pipe_ = nullptr;
    std::vector<ImgTab> last_tab_[2];
    bool last_keep_all_[2] = {false, false};
    // Contexts of recent batch geometries (per plan).  A subtitle stays on screen for dozens of frames, so consecutive batches
    // of a video repeat a small set of recogniser geometries (same boxes -> same padded crop widths): shape inference, arena
    // plan and ~100 tensor-map encodings are reused instead of rebuilt between the detector's and the recogniser's kernels.
    struct CtxSlot {
        ExecContext cx;
        std::vector<ImgTab> tab;
        bool keep_all = false;
        uint64_t stamp = 0;
    };
    std::vector<CtxSlot> ctx_store_[2];
    uint64_t ctx_clock_ = 0;
    void purge_contexts(int which);
};

}  // namespace vse
