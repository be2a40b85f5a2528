// plan.h — binary plan format shared with video_subtitle_extractor_b200/plan.py (Plan.serialize).
// A plan is the fused, channel-last restatement of one of the reference's shipped Paddle graphs
// (reference backend/models/<ver>/<name>/inference.pdmodel); see plan.py for how it is derived.
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace vse {

enum Op : int32_t {
    OP_CONV = 0, OP_DWCONV, OP_DECONV2, OP_GPOOL, OP_VECLIN, OP_CHSCALE, OP_POOL, OP_UPSAMPLE, OP_ADD, OP_COPY,
    OP_LAYERNORM, OP_ATTN, OP_ELTWISE, OP_SOFTMAX, OP_STEM, OP_LSTM, OP_COUNT
};
static const char* const kOpNames[OP_COUNT] = {"CONV", "DWCONV", "DECONV2", "GPOOL", "VECLIN", "CHSCALE", "POOL",
                                               "UPSAMPLE", "ADD", "COPY", "LAYERNORM", "ATTN", "ELTWISE", "SOFTMAX",
                                               "STEM", "LSTM"};

enum Act : int32_t { ACT_NONE = 0, ACT_RELU, ACT_HSWISH, ACT_HSIGMOID, ACT_SWISH, ACT_SIGMOID, ACT_RELU6 };
enum Kind : int32_t { KIND_IMG = 0, KIND_VEC = 1 };
enum DType : int32_t { DT_ACT = 0, DT_F32 = 1, DT_U8 = 2 };

// integer parameter slots
enum PSlot {
    P_KH = 0, P_KW, P_SH, P_SW, P_PH, P_PW, P_CIN, P_COUT, P_ACT, P_ACT2, P_HAS_POST, P_HAS_RES, P_SCALE, P_HAS_ADD,
    P_RESIDUAL, P_IS_MAX, P_CEIL, P_EXCLUSIVE, P_HEADS, P_DIM, P_COUNT
};
enum FSlot { F_HS_SLOPE = 0, F_HS_OFFSET, F_EPS, F_QSCALE, F_COUNT };
enum WSlot { W_WEIGHT = 0, W_BIAS, W_POST_SCALE, W_POST_SHIFT, W_GAMMA, W_BETA, W_SCALE, W_SHIFT, W_COUNT };

static constexpr uint32_t kPlanMagic = 0x50455356u;  // 'VSEP'
static constexpr uint32_t kPlanVersion = 3;

#pragma pack(push, 1)
struct PlanHeader {
    uint32_t magic, version, n_values, n_steps, r0, r1;
    int64_t n_weights;
    int32_t input_vid;
    int32_t output_vids[4];
    int32_t h1_values[8];
    float norm_scale[3];
    float norm_shift[3];
    char name[64];
};
struct ValueRec {
    int32_t channels, cstride, kind, dtype, alias_of, alias_coff, first_def, last_use;
};
struct StepRec {
    int32_t op;
    int32_t ins[4];
    int32_t out;
    int32_t p[P_COUNT];
    float f[F_COUNT];
    int64_t woff[W_COUNT];
    int64_t wsize[W_COUNT];
};
#pragma pack(pop)
static_assert(sizeof(PlanHeader) == 172, "PlanHeader layout");
static_assert(sizeof(ValueRec) == 32, "ValueRec layout");
static_assert(sizeof(StepRec) == 248, "StepRec layout");

struct PlanData {
    PlanHeader hdr{};
    std::vector<ValueRec> values;
    std::vector<StepRec> steps;
    std::vector<float> weights;

    // returns empty string on success
    std::string parse(const void* blob, size_t n) {
        const uint8_t* p = static_cast<const uint8_t*>(blob);
        if (n < sizeof(PlanHeader)) return "plan blob too small";
        std::memcpy(&hdr, p, sizeof(PlanHeader));
        if (hdr.magic != kPlanMagic) return "bad plan magic";
        if (hdr.version != kPlanVersion) return "plan version mismatch (rebuild the plan with this package)";
        size_t need = sizeof(PlanHeader) + size_t(hdr.n_values) * sizeof(ValueRec) + size_t(hdr.n_steps) * sizeof(StepRec) +
                      size_t(hdr.n_weights) * sizeof(float);
        if (n != need) return "plan blob size mismatch";
        p += sizeof(PlanHeader);
        values.resize(hdr.n_values);
        std::memcpy(values.data(), p, values.size() * sizeof(ValueRec));
        p += values.size() * sizeof(ValueRec);
        steps.resize(hdr.n_steps);
        std::memcpy(steps.data(), p, steps.size() * sizeof(StepRec));
        p += steps.size() * sizeof(StepRec);
        weights.resize(hdr.n_weights);
        std::memcpy(weights.data(), p, weights.size() * sizeof(float));
        for (auto& s : steps) {
            if (s.op < 0 || s.op >= OP_COUNT) return "bad op code";
            if (s.out < 0 || s.out >= (int)hdr.n_values) return "bad step output";
            for (int i = 0; i < 4; i++)
                if (s.ins[i] >= (int)hdr.n_values) return "bad step input";
            for (int i = 0; i < W_COUNT; i++)
                if (s.woff[i] >= 0 && s.woff[i] + s.wsize[i] > hdr.n_weights) return "weight slice out of range";
        }
        return "";
    }
    const float* w(const StepRec& s, int slot) const { return s.woff[slot] >= 0 ? weights.data() + s.woff[slot] : nullptr; }
};

}  // namespace vse
