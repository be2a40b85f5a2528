// engine.cu — plan loader, run-time shape inference, arena planner and step executor.
// Replaces the paddle.inference predictor runs behind reference backend/tools/ocr.py:27 and
// backend/tools/subtitle_detect.py:25 (see include/vse_b200.h for the boundary).
#include "engine.h"
#include "fast_kernels.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace vse {

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
static inline int pad8(int c) { return round_up(c, 8); }

Engine::Engine(const vse_config& c) : cfg(c) {
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) throw CudaError{"no CUDA device available (this engine has no CPU fallback)"};
    if (cfg.device < 0 || cfg.device >= ndev) throw InvalidArg{"device ordinal out of range"};
    VSE_CUDA(cudaSetDevice(cfg.device));
    VSE_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    cudaDeviceProp prop{};
    VSE_CUDA(cudaGetDeviceProperties(&prop, cfg.device));
    sm_count = prop.multiProcessorCount;
    if (prop.major != 10) throw CudaError{"this engine is built for sm_100a (B200); found compute capability " +
                                          std::to_string(prop.major) + "." + std::to_string(prop.minor)};
}

size_t Engine::elt_size(int which, const ValueRec& v) const {
    if (v.dtype == DT_U8) return 1;
    if (v.dtype == DT_F32) return 4;
    return plan_prec_[which] == VSE_PRECISION_FP16 ? 2 : 4;
}

int Engine::value_cs(const PlanData& pd, int vid) const {
    const ValueRec& v = pd.values[vid];
    const ValueRec& r = v.alias_of >= 0 ? pd.values[v.alias_of] : v;
    if (r.kind == KIND_VEC) return r.channels;
    if (r.dtype == DT_U8) return 4;            // BGRX
    if (r.dtype == DT_F32) return r.channels;  // dense outputs
    return pad8(r.channels);
}

// ------------------------------------------------------------------------------------------------
// plan loading: re-lay every parameter for the kernels
// ------------------------------------------------------------------------------------------------
void Engine::load_plan(int which, const void* blob, size_t n) {
    if (which < 0 || which > 1) throw InvalidArg{"plan index must be 0 (det) or 1 (rec)"};
    LoadedPlan& lp = plans_[which];
    lp.loaded = false;
    purge_contexts(which);
    std::string err = lp.data.parse(blob, n);
    if (!err.empty()) throw InvalidArg{err};
    // activation type of this plan: the engine's, except that VSE_FLAG_DET_FP32 keeps the detector in fp32 (server detector
    // V4/ch_det: activations beyond the fp16 range) while the recogniser stays on the fp16 tensor-core path
    plan_prec_[which] = cfg.precision;
    if (which == 0 && (cfg.flags & VSE_FLAG_DET_FP32)) plan_prec_[which] = VSE_PRECISION_FP32;
    if (which == 0 && (cfg.flags & VSE_FLAG_DET_TF32)) plan_prec_[which] = VSE_PRECISION_TF32;
    if (which == 0 && (cfg.flags & VSE_FLAG_DET_FP32_TC)) plan_prec_[which] = VSE_PRECISION_FP32_TC;
    if (plan_prec_[which] < VSE_PRECISION_FP16 || plan_prec_[which] > VSE_PRECISION_FP32_TC) throw InvalidArg{"bad precision"};
    prepare_plan(which, lp);
    lp.loaded = true;
}

// split mode (VSE_PRECISION_FP32_TC): per-step operand scale from calibrated input ranges — the power of two that brings the
// largest expected |x| into [2^13, 2^14) (4x headroom below the fp16 limit 65504); see gemm_tc.cu, transform warps
void Engine::set_conv_input_ranges(int which, const float* absmax, int n_steps) {
    if (which < 0 || which > 1 || !plans_[which].loaded) throw StateError{"set_conv_input_ranges: plan not loaded"};
    LoadedPlan& lp = plans_[which];
    if (n_steps != int(lp.data.steps.size())) throw InvalidArg{"set_conv_input_ranges: one value per plan step expected"};
    for (int k = 0; k < n_steps; k++) {
        float s = tc_split_activation_scale();
        if (absmax[k] > 0.f && std::isfinite(absmax[k])) {
            int e = 0;
            std::frexp(absmax[k], &e);                       // absmax = m * 2^e, m in [0.5, 1)
            s = std::ldexp(1.f, std::max(-30, std::min(14, 14 - e)));
        }
        lp.a_scale[k] = s;
    }
    purge_contexts(which);                                   // contexts cache the scale: rebuild on the next run
}

void Engine::purge_contexts(int which) {
    last_tab_[which].clear();
    for (auto& sl : ctx_store_[which]) { sl.cx.tabs.release(); sl.cx.tc_gdev.release(); }
    ctx_store_[which].clear();
}

void Engine::prepare_plan(int which, LoadedPlan& lp) {
    const PlanData& pd = lp.data;
    std::vector<float> host;
    host.reserve(pd.weights.size() * 2 + 4096);
    struct Off { size_t w_t = SIZE_MAX, bias_pk = SIZE_MAX, ps_pk = SIZE_MAX, pb_pk = SIZE_MAX, w = SIZE_MAX, bias = SIZE_MAX, ps = SIZE_MAX, pb = SIZE_MAX, g = SIZE_MAX, b = SIZE_MAX, sc = SIZE_MAX, sh = SIZE_MAX; };
    std::vector<Off> offs(pd.steps.size());
    lp.dev.assign(pd.steps.size(), StepDev{});
    lp.a_scale.assign(pd.steps.size(), tc_split_activation_scale());
    auto alloc = [&](size_t nfloats) {
        size_t o = host.size();
        host.resize(o + round_up(int(nfloats), 4), 0.f);
        return o;
    };
    auto put_vec = [&](const float* src, int64_t n, int padded) -> size_t {
        size_t o = alloc(padded);
        if (src) std::memcpy(host.data() + o, src, sizeof(float) * size_t(std::min<int64_t>(n, padded)));
        return o;
    };
    for (size_t k = 0; k < pd.steps.size(); k++) {
        const StepRec& s = pd.steps[k];
        StepDev& d = lp.dev[k];
        Off& o = offs[k];
        const int cin = s.p[P_CIN], cout = s.p[P_COUT];
        switch (s.op) {
            case OP_CONV:
            case OP_STEM: {
                const int kh = s.p[P_KH], kw = s.p[P_KW];
                const int cin_pad = s.op == OP_STEM ? 4 : pad8(cin);
                d.w_ci = round_up(cin_pad, 16);
                d.w_co = round_up(cout, 64);
                if (s.wsize[W_WEIGHT] != int64_t(cout) * kh * kw * cin) throw InvalidArg{"conv weight size mismatch"};
                o.w = alloc(size_t(kh) * kw * d.w_ci * d.w_co);
                const float* src = pd.w(s, W_WEIGHT);  // [cout][kh][kw][cin]
                for (int co = 0; co < cout; co++)
                    for (int t = 0; t < kh * kw; t++)
                        for (int ci = 0; ci < cin; ci++)
                            host[o.w + (size_t(t) * d.w_ci + ci) * d.w_co + co] = src[(size_t(co) * kh * kw + t) * cin + ci];
                // zero padded past the tensor-core tile grid (n_chunks * n_chunk <= cout + 16 * n_chunks): the epilogue reads
                // these per 4 channels without bounds checks
                const int vec_pad = d.w_co + 512;
                o.bias = put_vec(pd.w(s, W_BIAS), cout, vec_pad);
                if (s.p[P_HAS_POST]) {
                    o.ps = put_vec(pd.w(s, W_POST_SCALE), cout, vec_pad);
                    o.pb = put_vec(pd.w(s, W_POST_SHIFT), cout, vec_pad);
                }
                // pixel-packed variant (gemm_tc.h): narrow 1x1 convs whose input pixels are 32 / 64 contiguous bytes and whose
                // output is a dense value (not a slice of a concat buffer)
                if (s.op == OP_CONV && kh == 1 && kw == 1 && s.p[P_SH] == 1 && s.p[P_SW] == 1 && s.p[P_PH] == 0 && s.p[P_PW] == 0 &&
                    plan_prec_[which] == VSE_PRECISION_FP16 && !(cfg.flags & (VSE_FLAG_NO_TENSOR_CORES | VSE_FLAG_NO_PIXEL_PACK)) &&
                    s.ins[0] != pd.hdr.input_vid && pd.values[s.out].dtype != DT_F32) {
                    const int ics = value_cs(pd, s.ins[0]), ocs = value_cs(pd, s.out);
                    const bool res_ok = !s.p[P_HAS_RES] || value_cs(pd, s.ins[1]) == ocs;
                    if ((ics == 16 || ics == 32) && cin <= ics && ocs == pad8(cout) && res_ok && (64 / ics) * ocs <= 1024) {
                        d.pack = 64 / ics;
                        const int n = d.pack * ocs, npad = n + 512;
                        auto rep = [&](const float* src) -> size_t {
                            size_t off = alloc(npad);
                            for (int g = 0; g < d.pack; g++)
                                for (int c = 0; c < cout; c++) host[off + size_t(g) * ocs + c] = src ? src[c] : 0.f;
                            return off;
                        };
                        o.bias_pk = rep(pd.w(s, W_BIAS));
                        if (s.p[P_HAS_POST]) {
                            o.ps_pk = rep(pd.w(s, W_POST_SCALE));
                            o.pb_pk = rep(pd.w(s, W_POST_SHIFT));
                        }
                    }
                }
                break;
            }
            case OP_DWCONV: {
                const int kh = s.p[P_KH], kw = s.p[P_KW], cp = pad8(cin);
                if (s.wsize[W_WEIGHT] != int64_t(kh) * kw * cin) throw InvalidArg{"dwconv weight size mismatch"};
                o.w = alloc(size_t(kh) * kw * cp);
                const float* src = pd.w(s, W_WEIGHT);  // [kh][kw][c]
                for (int t = 0; t < kh * kw; t++)
                    for (int c = 0; c < cin; c++) host[o.w + size_t(t) * cp + c] = src[size_t(t) * cin + c];
                o.bias = put_vec(pd.w(s, W_BIAS), cout, cp);
                if (s.p[P_HAS_POST]) {
                    o.ps = put_vec(pd.w(s, W_POST_SCALE), cout, cp);
                    o.pb = put_vec(pd.w(s, W_POST_SHIFT), cout, cp);
                }
                break;
            }
            case OP_DECONV2: {
                const int cip = pad8(cin), cop = pad8(cout);
                if (s.wsize[W_WEIGHT] != int64_t(4) * cout * cin) throw InvalidArg{"deconv weight size mismatch"};
                o.w = alloc(size_t(4) * cop * cip);
                const float* src = pd.w(s, W_WEIGHT);  // [kh][kw][cout][cin]
                for (int pos = 0; pos < 4; pos++)
                    for (int co = 0; co < cout; co++)
                        for (int ci = 0; ci < cin; ci++)
                            host[o.w + (size_t(pos) * cop + co) * cip + ci] = src[(size_t(pos) * cout + co) * cin + ci];
                o.bias = put_vec(pd.w(s, W_BIAS), cout, cop);
                if (s.p[P_HAS_POST]) {
                    o.ps = put_vec(pd.w(s, W_POST_SCALE), cout, cop);
                    o.pb = put_vec(pd.w(s, W_POST_SHIFT), cout, cop);
                }
                break;
            }
            case OP_VECLIN: {
                if (s.wsize[W_WEIGHT] != int64_t(cout) * cin) throw InvalidArg{"veclin weight size mismatch"};
                o.w = put_vec(pd.w(s, W_WEIGHT), int64_t(cout) * cin, cout * cin);
                {
                    o.w_t = alloc(size_t(cout) * cin);
                    const float* src = pd.w(s, W_WEIGHT);   // [cout][cin]
                    for (int co = 0; co < cout; co++)
                        for (int ci = 0; ci < cin; ci++) host[o.w_t + size_t(ci) * cout + co] = src[size_t(co) * cin + ci];
                }
                o.bias = put_vec(pd.w(s, W_BIAS), cout, cout);
                if (s.p[P_HAS_POST]) {
                    o.ps = put_vec(pd.w(s, W_POST_SCALE), cout, cout);
                    o.pb = put_vec(pd.w(s, W_POST_SHIFT), cout, cout);
                }
                break;
            }
            case OP_LAYERNORM: {
                int c = pd.values[s.out].channels;
                o.g = put_vec(pd.w(s, W_GAMMA), c, pad8(c));
                o.b = put_vec(pd.w(s, W_BETA), c, pad8(c));
                break;
            }
            case OP_ELTWISE: {
                int c = pd.values[s.out].channels;
                o.sc = put_vec(pd.w(s, W_SCALE), c, pad8(c));
                o.sh = put_vec(pd.w(s, W_SHIFT), c, pad8(c));
                break;
            }
            case OP_LSTM: {
                const int hidden = s.p[P_HEADS], ndir = s.p[P_SCALE];
                if (!lstm_supported(hidden) || ndir < 1 || ndir > 2) throw InvalidArg{"LSTM step: unsupported hidden size / directions"};
                if (s.wsize[W_WEIGHT] != int64_t(ndir) * 4 * hidden * hidden || cin != ndir * 4 * hidden || cout != ndir * hidden)
                    throw InvalidArg{"LSTM weight size mismatch"};
                o.w = alloc(lstm_packed_weight_floats(hidden, ndir));
                lstm_pack_weights(pd.w(s, W_WEIGHT), hidden, ndir, host.data() + o.w);
                break;
            }
            default:
                break;
        }
    }
    lp.weights.reserve(host.size() * sizeof(float));
    VSE_CUDA(cudaMemcpy(lp.weights.p, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice));
    // fp16 K-major weight matrices for the tensor-core path
    lp.tcw.assign(pd.steps.size(), TcWeights{});
    lp.tcw_off.assign(pd.steps.size(), 0);
    lp.tcw_pk.assign(pd.steps.size(), TcWeights{});
    lp.tcw_pk_off.assign(pd.steps.size(), 0);
    lp.tcw_dc.assign(pd.steps.size() * 4, TcWeights{});
    lp.tcw_dc_off.assign(pd.steps.size() * 4, 0);
    if (plan_prec_[which] != VSE_PRECISION_FP32 && !(cfg.flags & VSE_FLAG_NO_TENSOR_CORES)) {
        const int tc_mode = plan_prec_[which] == VSE_PRECISION_TF32 ? TC_TF32 : plan_prec_[which] == VSE_PRECISION_FP32_TC ? TC_SPLIT : TC_F16;
        std::vector<uint16_t> all;
        auto append = [&](TcWeights& t) -> size_t {
            while (all.size() % 512) all.push_back(0);   // 1024-byte aligned matrices
            size_t off = all.size() * sizeof(uint16_t);
            all.insert(all.end(), t.b.begin(), t.b.end());
            t.b.clear();
            t.b.shrink_to_fit();
            return off;
        };
        lp.tcw_dc.assign(pd.steps.size() * 4, TcWeights{});
        lp.tcw_dc_off.assign(pd.steps.size() * 4, 0);
        for (size_t k = 0; k < pd.steps.size(); k++) {
            const StepRec& s = pd.steps[k];
            if (s.op == OP_DECONV2 && s.p[P_COUT] % 8 == 0 && pd.values[s.out].dtype != DT_F32 && s.ins[0] != pd.hdr.input_vid) {
                const int cin = s.p[P_CIN], cout = s.p[P_COUT];
                for (int pos = 0; pos < 4; pos++) {     // canonical layout [kh][kw][cout][cin]: position pos = dy * 2 + dx
                    lp.tcw_dc[k * 4 + pos] = tc_pack_weights(pd.w(s, W_WEIGHT) + size_t(pos) * cout * cin, cout, cin, 1, tc_mode);
                    lp.tcw_dc_off[k * 4 + pos] = append(lp.tcw_dc[k * 4 + pos]);
                }
                continue;
            }
            if (s.op != OP_CONV || s.p[P_SH] != 1 || s.p[P_SW] != 1) continue;
            lp.tcw[k] = tc_pack_weights(pd.w(s, W_WEIGHT), s.p[P_COUT], s.p[P_CIN], s.p[P_KH] * s.p[P_KW], tc_mode);
            lp.tcw_off[k] = append(lp.tcw[k]);
            if (lp.dev[k].pack > 0) {
                lp.tcw_pk[k] = tc_pack_weights_pixelpacked(pd.w(s, W_WEIGHT), s.p[P_COUT], s.p[P_CIN], value_cs(pd, s.ins[0]),
                                                           value_cs(pd, s.out), lp.dev[k].pack);
                lp.tcw_pk_off[k] = append(lp.tcw_pk[k]);
            }
        }
        if (!all.empty()) {
            lp.tc_weights.reserve(all.size() * sizeof(uint16_t));
            VSE_CUDA(cudaMemcpy(lp.tc_weights.p, all.data(), all.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
        }
    }
    const float* base = lp.weights.as<float>();
    auto ptr = [&](size_t off) -> const float* { return off == SIZE_MAX ? nullptr : base + off; };
    for (size_t k = 0; k < pd.steps.size(); k++) {
        StepDev& d = lp.dev[k];
        const Off& o = offs[k];
        d.w = ptr(o.w); d.bias = ptr(o.bias); d.post_scale = ptr(o.ps); d.post_shift = ptr(o.pb);
        d.gamma = ptr(o.g); d.beta = ptr(o.b); d.scale = ptr(o.sc); d.shift = ptr(o.sh);
        d.w_t = ptr(o.w_t);
        d.bias_pk = ptr(o.bias_pk); d.post_scale_pk = ptr(o.ps_pk); d.post_shift_pk = ptr(o.pb_pk);
    }
}

// ------------------------------------------------------------------------------------------------
// geometry inference + arena planning
// ------------------------------------------------------------------------------------------------
static Geo derive_geo(const Geo& in, int kh, int kw, int sh, int sw, int ph, int pw, bool ceil_mode) {
    Geo g;
    g.tab.resize(in.tab.size());
    int64_t off = 0;
    for (size_t i = 0; i < in.tab.size(); i++) {
        const ImgTab& t = in.tab[i];
        int ho, wo;
        if (!ceil_mode) {
            ho = (t.h + 2 * ph - kh) / sh + 1;
            wo = (t.w + 2 * pw - kw) / sw + 1;
        } else {
            ho = (t.h + 2 * ph - kh + sh - 1) / sh + 1;
            wo = (t.w + 2 * pw - kw + sw - 1) / sw + 1;
            if ((ho - 1) * sh >= t.h + ph) ho--;
            if ((wo - 1) * sw >= t.w + pw) wo--;
        }
        if (ho < 1 || wo < 1) throw InvalidArg{"image too small for this network"};
        g.tab[i] = ImgTab{int(off), ho, wo, wo};
        off += int64_t(ho) * wo;
        g.max_pix = std::max(g.max_pix, ho * wo);
    }
    if (off > 0x7fffffffLL) throw InvalidArg{"batch too large (pixel index overflow)"};
    g.total = off;
    return g;
}

static Geo scale_geo(const Geo& in, int scale) {
    Geo g;
    g.tab.resize(in.tab.size());
    int64_t off = 0;
    for (size_t i = 0; i < in.tab.size(); i++) {
        int ho = in.tab[i].h * scale, wo = in.tab[i].w * scale;
        g.tab[i] = ImgTab{int(off), ho, wo, wo};
        off += int64_t(ho) * wo;
        g.max_pix = std::max(g.max_pix, ho * wo);
    }
    if (off > 0x7fffffffLL) throw InvalidArg{"batch too large (pixel index overflow)"};
    g.total = off;
    return g;
}

static inline bool keep_all_disables_tc(bool) { return false; }

static bool same_geo(const Geo& a, const Geo& b) {
    if (a.tab.size() != b.tab.size()) return false;
    for (size_t i = 0; i < a.tab.size(); i++)
        if (a.tab[i].h != b.tab[i].h || a.tab[i].w != b.tab[i].w) return false;
    return true;
}

void Engine::build_context(int which, const std::vector<ImgTab>& in_tab, bool keep_all) {
    LoadedPlan& lp = plans_[which];
    if (!lp.loaded) throw InvalidArg{which == 0 ? "detection plan not loaded" : "recognition plan not loaded"};
    const PlanData& pd = lp.data;
    ExecContext& cx = ctx_[which];
    cx.geos.clear();
    cx.vals.assign(pd.values.size(), ValueRt{});
    cx.n_img = int(in_tab.size());
    auto add_geo = [&](Geo&& g) -> int {
        for (size_t i = 0; i < cx.geos.size(); i++)
            if (same_geo(cx.geos[i], g)) return int(i);
        cx.geos.push_back(std::move(g));
        return int(cx.geos.size()) - 1;
    };
    {
        Geo g;
        g.tab = in_tab;
        int64_t off = 0;
        for (auto& t : g.tab) {
            t.off = int(off);
            off += int64_t(t.h) * t.w;
            g.max_pix = std::max(g.max_pix, t.h * t.w);
        }
        g.total = off;
        cx.geos.push_back(std::move(g));  // geo 0 keeps valid widths; never merged with others
        cx.vals[pd.hdr.input_vid].geo = 0;
    }
    auto root_of = [&](int v) { return pd.values[v].alias_of >= 0 ? pd.values[v].alias_of : v; };
    size_t scratch = 0;
    for (size_t k = 0; k < pd.steps.size(); k++) {
        const StepRec& s = pd.steps[k];
        const int gin = s.ins[0] >= 0 ? cx.vals[s.ins[0]].geo : -1;
        int gout = -1;
        switch (s.op) {
            case OP_CONV: case OP_STEM: case OP_DWCONV: {
                if (gin < 0) throw InvalidArg{"conv input has no geometry"};
                Geo g = derive_geo(cx.geos[gin], s.p[P_KH], s.p[P_KW], s.p[P_SH], s.p[P_SW], s.p[P_PH], s.p[P_PW], false);
                gout = add_geo(std::move(g));
                break;
            }
            case OP_POOL: {
                Geo g = derive_geo(cx.geos[gin], s.p[P_KH], s.p[P_KW], s.p[P_SH], s.p[P_SW], s.p[P_PH], s.p[P_PW], s.p[P_CEIL] != 0);
                gout = add_geo(std::move(g));
                break;
            }
            case OP_DECONV2: gout = add_geo(scale_geo(cx.geos[gin], 2)); break;
            case OP_UPSAMPLE: gout = add_geo(scale_geo(cx.geos[gin], s.p[P_SCALE])); break;
            case OP_GPOOL: {
                int cp = pad8(pd.values[s.ins[0]].channels);
                scratch = std::max(scratch, size_t(cx.n_img) * 64 * cp * sizeof(float));
                gout = -1;
                break;
            }
            case OP_VECLIN: gout = -1; break;
            default: {
                // same geometry as input; geo 0 (input with valid widths) is never propagated
                if (gin == 0) {
                    Geo g = cx.geos[0];
                    for (auto& t : g.tab) t.vw = t.w;
                    gout = add_geo(std::move(g));
                } else gout = gin;
                break;
            }
        }
        // outputs that alias into a concat root share the root's geometry
        cx.vals[s.out].geo = gout;
        int r = root_of(s.out);
        if (r != s.out) {
            if (cx.vals[r].geo >= 0 && gout >= 0 && !same_geo(cx.geos[cx.vals[r].geo], cx.geos[gout]))
                throw InvalidArg{"concat inputs with different geometry"};
            cx.vals[r].geo = gout;
        }
        if (s.op == OP_COPY) cx.vals[s.out].geo = gin;
    }
    // views of a root that were assigned before the root got its geometry
    for (size_t v = 0; v < pd.values.size(); v++) {
        int r = root_of(int(v));
        if (r != int(v) && cx.vals[v].geo < 0) cx.vals[v].geo = cx.vals[r].geo;
        if (r != int(v) && cx.vals[r].geo < 0) cx.vals[r].geo = cx.vals[v].geo;
    }
    // H==1 requirement of [B,T,C] views
    for (int i = 0; i < 8; i++) {
        int v = pd.hdr.h1_values[i];
        if (v < 0) continue;
        int g = cx.vals[v].geo;
        if (g >= 0)
            for (auto& t : cx.geos[g].tab)
                if (t.h != 1) throw InvalidArg{"recognition input height does not reduce to 1 (wrong rec_image_h?)"};
    }
    // device tables
    size_t ntab = 0;
    for (auto& g : cx.geos) { g.tab_off = ntab; ntab += g.tab.size(); }
    pin_.reserve(ntab * sizeof(ImgTab));
    for (auto& g : cx.geos) std::memcpy(pin_.as<ImgTab>() + g.tab_off, g.tab.data(), g.tab.size() * sizeof(ImgTab));
    cx.tabs.reserve(ntab * sizeof(ImgTab));
    launch_upload(cx.tabs.p, pin_.p, ntab * sizeof(ImgTab), stream);   // kernel-parameter upload: no copy engine, no sync

    // arena planning (first-fit free list over root buffers)
    struct Block { size_t off, size; };
    std::vector<Block> free_list;
    size_t top = 0;
    auto arena_alloc = [&](size_t n) -> size_t {
        n = (n + 255) & ~size_t(255);
        for (size_t i = 0; i < free_list.size(); i++) {
            if (free_list[i].size >= n) {
                size_t o = free_list[i].off;
                free_list[i].off += n;
                free_list[i].size -= n;
                if (free_list[i].size == 0) free_list.erase(free_list.begin() + i);
                return o;
            }
        }
        size_t o = top;
        top += n;
        return o;
    };
    auto arena_free = [&](size_t off, size_t n) {
        n = (n + 255) & ~size_t(255);
        free_list.push_back(Block{off, n});
        std::sort(free_list.begin(), free_list.end(), [](const Block& a, const Block& b) { return a.off < b.off; });
        for (size_t i = 0; i + 1 < free_list.size();) {
            if (free_list[i].off + free_list[i].size == free_list[i + 1].off) {
                free_list[i].size += free_list[i + 1].size;
                free_list.erase(free_list.begin() + i + 1);
            } else i++;
        }
        if (!free_list.empty() && free_list.back().off + free_list.back().size == top) {
            top = free_list.back().off;
            free_list.pop_back();
        }
    };
    size_t peak = 0;
    auto root_bytes = [&](int v) -> size_t {
        const ValueRec& r = pd.values[v];
        if (r.kind == KIND_VEC) return size_t(cx.n_img) * r.channels * sizeof(float);
        int g = cx.vals[v].geo;
        if (g < 0) throw InvalidArg{"value without geometry: " + std::to_string(v)};
        return size_t(cx.geos[g].total) * value_cs(pd, v) * elt_size(which, r);
    };
    const int nsteps = int(pd.steps.size());
    // Fused residual squeeze-excite around a 1x1 convolution (RSE in-convs of the FPN): conv(act none) -> GPOOL -> FC -> FC
    // -> CHSCALE(residual).  mean(conv(x)) = W mean(x) + b, so the gate is computed from the conv's INPUT and applied in the
    // conv's epilogue: the conv output is never pooled nor re-read.  The group's results (gate vector, CHSCALE output) are
    // then produced at the conv's step, so their arena lifetime starts there.
    std::vector<int> fd(pd.values.size());
    for (size_t v = 0; v < pd.values.size(); v++) fd[v] = pd.values[v].first_def;
    cx.se_conv.assign(pd.steps.size(), 0);
    if (!keep_all && (plan_prec_[which] == VSE_PRECISION_FP16 || plan_prec_[which] == VSE_PRECISION_FP32_TC) &&
        !(cfg.flags & (VSE_FLAG_NO_FAST_KERNELS | VSE_FLAG_NO_TENSOR_CORES | VSE_FLAG_NO_SE_CONV))) {
        auto readers = [&](int vid) {
            int n = 0;
            for (const StepRec& q : pd.steps)
                for (int i = 0; i < 4; i++) n += q.ins[i] == vid;
            return n;
        };
        auto plain = [](const StepRec& f) { return !f.p[P_HAS_POST] && !f.p[P_HAS_RES] && f.p[P_ACT2] == ACT_NONE; };
        for (int k = 0; k + 4 < nsteps; k++) {
            const StepRec& c = pd.steps[k];
            const StepRec &gp = pd.steps[k + 1], &f1 = pd.steps[k + 2], &f2 = pd.steps[k + 3], &ch = pd.steps[k + 4];
            if (c.op != OP_CONV || gp.op != OP_GPOOL || f1.op != OP_VECLIN || f2.op != OP_VECLIN || ch.op != OP_CHSCALE) continue;
            if (c.p[P_KH] != 1 || c.p[P_KW] != 1 || c.p[P_SH] != 1 || c.p[P_SW] != 1 || c.p[P_PH] || c.p[P_PW]) continue;
            if (c.p[P_ACT] != ACT_NONE || !plain(c) || !plain(f1) || !plain(f2) || lp.tcw[k].n_chunk == 0) continue;
            if (gp.ins[0] != c.out || f1.ins[0] != gp.out || f2.ins[0] != f1.out || ch.ins[0] != c.out || ch.ins[1] != f2.out) continue;
            if (!ch.p[P_RESIDUAL] || c.ins[0] == pd.hdr.input_vid) continue;
            const ValueRec &vc = pd.values[c.out], &vo = pd.values[ch.out];
            if (vc.alias_of >= 0 || vo.alias_of >= 0 || vc.dtype != DT_ACT || vo.dtype != DT_ACT) continue;
            const int cout = c.p[P_COUT];
            if (cout % 16 || f1.p[P_CIN] != cout || f2.p[P_COUT] != cout || f2.p[P_CIN] != f1.p[P_COUT] || f1.p[P_COUT] > 512 ||
                c.p[P_CIN] > 2048 || vo.channels != cout)
                continue;
            if (readers(c.out) != 2 || readers(gp.out) != 1 || readers(f1.out) != 1 || readers(f2.out) != 1) continue;
            const Geo& g = cx.geos[cx.vals[c.out].geo];
            bool uniform = true;
            for (auto& t : g.tab) uniform = uniform && t.h == g.tab[0].h && t.w == g.tab[0].w;
            if (!uniform || !lp.dev[k + 2].w_t || !lp.dev[k + 3].w_t || !lp.dev[k + 2].bias || !lp.dev[k + 3].bias || !lp.dev[k].bias) continue;
            cx.se_conv[k] = 1;
            scratch = std::max(scratch, size_t(cx.n_img) * 64 * pad8(c.p[P_CIN]) * sizeof(float));   // partial sums of the conv input
            fd[f2.out] = k;
            fd[ch.out] = k;
        }
    }
    // Fused depthwise -> pointwise (opt-in: VSE_FLAG_DWPW_FUSION; fp32 tensor-core mode, equal-sized images): DWCONV (3x3 / 5x5,
    // stride 1, same padding) whose only reader is the next step's 1x1 convolution.  The depthwise output is never allocated.
    // Measured slower than the separate kernels, hence not the default.
    cx.dwpw.assign(pd.steps.size(), 0);
    if (!keep_all && plan_prec_[which] == VSE_PRECISION_FP32_TC && (cfg.flags & VSE_FLAG_DWPW_FUSION) &&
        !(cfg.flags & (VSE_FLAG_NO_FAST_KERNELS | VSE_FLAG_NO_TENSOR_CORES))) {
        auto readers = [&](int vid) {
            int n = 0;
            for (const StepRec& q : pd.steps)
                for (int i = 0; i < 4; i++) n += q.ins[i] == vid;
            return n;
        };
        for (int k = 0; k + 1 < nsteps; k++) {
            const StepRec &dwc = pd.steps[k], &c = pd.steps[k + 1];
            if (dwc.op != OP_DWCONV || c.op != OP_CONV || c.ins[0] != dwc.out || cx.se_conv[k + 1]) continue;
            const int kd = dwc.p[P_KH];
            if ((kd != 3 && kd != 5) || dwc.p[P_KW] != kd || dwc.p[P_SH] != 1 || dwc.p[P_SW] != 1 || dwc.p[P_PH] != kd / 2 || dwc.p[P_PW] != kd / 2) continue;
            if (dwc.p[P_HAS_RES] || dwc.p[P_ACT2] != ACT_NONE || (dwc.p[P_ACT] != ACT_NONE && dwc.p[P_ACT] != ACT_RELU && dwc.p[P_ACT] != ACT_HSWISH)) continue;
            if (c.p[P_KH] != 1 || c.p[P_KW] != 1 || c.p[P_SH] != 1 || c.p[P_SW] != 1 || c.p[P_PH] || c.p[P_PW]) continue;
            const ValueRec &vd = pd.values[dwc.out], &vc = pd.values[c.out];
            if (vd.alias_of >= 0 || vd.dtype != DT_ACT || vc.dtype != DT_ACT || readers(dwc.out) != 1 || dwc.ins[0] == pd.hdr.input_vid) continue;
            if (lp.tcw[k + 1].n_chunk == 0 || lp.tcw[k + 1].n_chunks != 1 || lp.tcw[k + 1].stack || !lp.dev[k].bias || !lp.dev[k].w) continue;
            bool is_out = false;
            for (int i = 0; i < 4; i++) is_out = is_out || pd.hdr.output_vids[i] == dwc.out;
            if (is_out) continue;
            const Geo& g = cx.geos[cx.vals[dwc.ins[0]].geo];
            bool uniform = !g.tab.empty();
            for (auto& t : g.tab) uniform = uniform && t.h == g.tab[0].h && t.w == g.tab[0].w;
            if (!uniform) continue;
            // shared-memory budget (mirrors tc_conv_setup_dwpw)
            const int nkb = (dwc.p[P_CIN] + 31) / 32, nck = lp.tcw[k + 1].n_chunk;
            const int win = ((16 + kd - 1) * (8 + kd - 1) * 128 + 1023) / 1024 * 1024;
            const int b_all = nkb * nck * 128;
            const bool res = b_all <= 112 * 1024;
            const int need = 2 * (win + 16384 + (res ? 0 : nck * 128)) + (res ? b_all : 0) + 2 * 16384 + (kd * kd + 3) * nkb * 128 + 3 * nck * 4 + 1536;
            if (need > 227 * 1024) continue;
            cx.dwpw[k] = 1;
            fd[dwc.out] = -3;           // never materialised
            // the fused kernel runs at the depthwise step: its output must be live BEFORE the depthwise input is released at the
            // end of that step (allocated one step later it could land in the very memory the kernel is still reading)
            fd[c.out] = std::min(fd[c.out], k);
        }
    }
    for (int k = -1; k <= nsteps; k++) {
        for (size_t v = 0; v < pd.values.size(); v++) {
            const ValueRec& r = pd.values[v];
            if (r.alias_of >= 0 || fd[v] != k || r.first_def == -2) continue;
            if (int(v) == pd.hdr.input_vid) continue;  // the input lives in caller memory
            ValueRt& rt = cx.vals[v];
            rt.bytes = root_bytes(int(v));
            rt.off = arena_alloc(rt.bytes);
            rt.live = true;
            peak = std::max(peak, top);
        }
        if (!keep_all && k >= 0)
            for (size_t v = 0; v < pd.values.size(); v++) {
                const ValueRec& r = pd.values[v];
                if (r.alias_of >= 0 || r.first_def == -2 || r.last_use != k || !cx.vals[v].live) continue;
                if (int(v) == pd.hdr.input_vid) continue;
                arena_free(cx.vals[v].off, cx.vals[v].bytes);
            }
    }
    cx.scratch_off = (peak + 255) & ~size_t(255);
    cx.scratch_bytes = scratch;
    cx.arena_bytes = cx.scratch_off + scratch + 256;
    arena_[which].reserve(cx.arena_bytes);

    // tensor-core plans (need the final arena addresses for the TMA descriptors)
    cx.tc.assign(pd.steps.size(), TcConv{});
    cx.tc_groups.assign(pd.steps.size(), {});
    for (size_t k = 0; k < pd.steps.size() && !keep_all_disables_tc(keep_all); k++) {
        const StepRec& s = pd.steps[k];
        if (s.op != OP_CONV || lp.tcw[k].n_chunk == 0) continue;
        const ValueRec& vo = pd.values[s.out];
        const int kh = s.p[P_KH], kw = s.p[P_KW], ph = s.p[P_PH], pw = s.p[P_PW];
        // fetched fp32 outputs are dense [pixels][channels]: only the single-channel map of a 1x1 head (V4/ch_det's last
        // convolution) is taken, in the fp32 tensor-core mode, through the kernel's direct-store epilogue
        const bool direct1 = vo.dtype == DT_F32 && vo.kind == KIND_IMG && s.p[P_COUT] == 1 && kh == 1 && kw == 1 && ph == 0 && pw == 0 &&
                             plan_prec_[which] == VSE_PRECISION_FP32_TC && !s.p[P_HAS_RES];
        if ((vo.dtype == DT_F32 && !direct1) || s.ins[0] == pd.hdr.input_vid) continue;
        const Geo& gi = cx.geos[cx.vals[s.ins[0]].geo];
        const bool flat = kh == 1 && kw == 1 && ph == 0 && pw == 0;
        bool uniform = true;
        for (auto& t : gi.tab) uniform = uniform && t.h == gi.tab[0].h && t.w == gi.tab[0].w;
        if (!flat && !uniform && 2 * ph == kh - 1 && 2 * pw == kw - 1 && !(cfg.flags & VSE_FLAG_NO_TENSOR_CORES)) {
            // ragged batch: one launch per run of equal-sized images (see ExecContext::tc_groups); all groups or none
            const size_t es = plan_prec_[which] == VSE_PRECISION_FP16 ? 2 : 4;
            const void* wdev = static_cast<const char*>(lp.tc_weights.p) + lp.tcw_off[k];
            std::vector<ExecContext::TcGroup> groups;
            bool ok = true;
            for (size_t i0 = 0; i0 < gi.tab.size() && ok;) {
                size_t i1 = i0 + 1;
                while (i1 < gi.tab.size() && gi.tab[i1].h == gi.tab[i0].h && gi.tab[i1].w == gi.tab[i0].w) i1++;
                ExecContext::TcGroup g;
                g.pix_off = gi.tab[i0].off;
                const char* base = static_cast<const char*>(vptr(which, s.ins[0])) + size_t(g.pix_off) * value_cs(pd, s.ins[0]) * es;
                const int64_t gpix = int64_t(i1 - i0) * gi.tab[i0].h * gi.tab[i0].w;
                ok = tc_conv_setup(g.tc, base, value_cs(pd, s.ins[0]), s.p[P_CIN], wdev, lp.tcw[k], false, gpix, int(i1 - i0), gi.tab[i0].h,
                                   gi.tab[i0].w, kh, kw, ph, pw, !(cfg.flags & VSE_FLAG_NO_ROWBOX), !(cfg.flags & VSE_FLAG_NO_HALO)).empty();
                g.tc.a_scale = lp.a_scale[k];
                groups.push_back(std::move(g));
                i0 = i1;
                if (groups.size() > 64) ok = false;
            }
            if (ok) cx.tc_groups[k] = std::move(groups);
            continue;
        }
        if (!flat && !(uniform && 2 * ph == kh - 1 && 2 * pw == kw - 1)) continue;
        if (flat && lp.dev[k].pack > 0 && lp.tcw_pk[k].n_chunk > 0 && gi.total % lp.dev[k].pack == 0) {
            // pixel-packed: `pack` pixels per GEMM row, K = 64, block-diagonal weights
            const void* wpk = static_cast<const char*>(lp.tc_weights.p) + lp.tcw_pk_off[k];
            std::string why = tc_conv_setup(cx.tc[k], vptr(which, s.ins[0]), 64, 64, wpk, lp.tcw_pk[k], true, gi.total / lp.dev[k].pack,
                                            cx.n_img, 1, 1, 1, 1, 0, 0, false);
            cx.tc[k].pack = why.empty() ? lp.dev[k].pack : 0;
            cx.tc[k].a_scale = lp.a_scale[k];
            if (why.empty()) continue;
        }
        const void* wdev = static_cast<const char*>(lp.tc_weights.p) + lp.tcw_off[k];
        if (k > 0 && cx.dwpw[k - 1]) {
            // fused depthwise -> pointwise: tensor maps over the DEPTHWISE input (ExecContext::dwpw)
            const StepRec& dwc = pd.steps[k - 1];
            const Geo& gd = cx.geos[cx.vals[dwc.ins[0]].geo];
            TcConv& t = cx.tc[k];
            std::string why = tc_conv_setup_dwpw(t, vptr(which, dwc.ins[0]), value_cs(pd, dwc.ins[0]), s.p[P_CIN], wdev, lp.tcw[k], cx.n_img,
                                                 gd.tab[0].h, gd.tab[0].w, dwc.p[P_KH]);
            if (!why.empty()) throw StateError{"fused depthwise -> pointwise: " + why + " (clear VSE_FLAG_DWPW_FUSION)"};
            t.dw_w = lp.dev[k - 1].w; t.dw_cp = pad8(dwc.p[P_CIN]); t.dw_bias = lp.dev[k - 1].bias; t.dw_act = dwc.p[P_ACT];
            t.dw_ps = dwc.p[P_HAS_POST] ? lp.dev[k - 1].post_scale : nullptr;
            t.dw_pt = dwc.p[P_HAS_POST] ? lp.dev[k - 1].post_shift : nullptr;
            t.a_scale = lp.a_scale[k];
            continue;
        }
        std::string why = tc_conv_setup(cx.tc[k], vptr(which, s.ins[0]), value_cs(pd, s.ins[0]), s.p[P_CIN], wdev, lp.tcw[k], flat,
                                        gi.total, cx.n_img, gi.tab[0].h, gi.tab[0].w, kh, kw, ph, pw, !(cfg.flags & VSE_FLAG_NO_ROWBOX),
                                        !(cfg.flags & VSE_FLAG_NO_HALO));
        if (!why.empty()) cx.tc[k].valid = false;
        cx.tc[k].a_scale = lp.a_scale[k];
        cx.tc[k].direct1 = direct1 ? 1 : 0;
    }
    // transposed convolutions: four spatial 1x1 launches over the (uniform) input geometry
    cx.tc_dc.assign(pd.steps.size() * 4, TcConv{});
    for (size_t k = 0; k < pd.steps.size(); k++) {
        const StepRec& s = pd.steps[k];
        if (s.op != OP_DECONV2 || lp.tcw_dc[k * 4].n_chunk == 0) continue;
        const Geo& gi = cx.geos[cx.vals[s.ins[0]].geo];
        bool uniform = true;
        for (auto& t : gi.tab) uniform = uniform && t.h == gi.tab[0].h && t.w == gi.tab[0].w;
        if (!uniform) continue;
        bool ok = true;
        for (int pos = 0; pos < 4 && ok; pos++) {
            TcConv& t = cx.tc_dc[k * 4 + pos];
            const void* wdev = static_cast<const char*>(lp.tc_weights.p) + lp.tcw_dc_off[k * 4 + pos];
            ok = tc_conv_setup(t, vptr(which, s.ins[0]), value_cs(pd, s.ins[0]), s.p[P_CIN], wdev, lp.tcw_dc[k * 4 + pos], false, gi.total,
                               cx.n_img, gi.tab[0].h, gi.tab[0].w, 1, 1, 0, 0, false, false).empty();
            t.a_scale = lp.a_scale[k];
        }
        if (!ok)
            for (int pos = 0; pos < 4; pos++) cx.tc_dc[k * 4 + pos].valid = false;
    }
    // device tables of the ragged steps (one launch per step: launch_conv_tc_groups)
    cx.tc_goff.assign(pd.steps.size(), 0);
    cx.tc_gup.assign(pd.steps.size(), 0);
    size_t gbytes = 0;
    for (size_t k = 0; k < pd.steps.size(); k++)
        if (!cx.tc_groups[k].empty()) {
            cx.tc_goff[k] = gbytes;
            gbytes += (tc_groups_dev_bytes(int(cx.tc_groups[k].size())) + 255) & ~size_t(255);
        }
    if (gbytes) cx.tc_gdev.reserve(gbytes);
}

void* Engine::vptr(int which, int vid) const {
    const PlanData& pd = plans_[which].data;
    const ValueRec& v = pd.values[vid];
    int r = v.alias_of >= 0 ? v.alias_of : vid;
    const ValueRt& rt = ctx_[which].vals[r];
    char* base = static_cast<char*>(arena_[which].p) + rt.off;
    if (v.alias_of >= 0) base += size_t(v.alias_coff) * elt_size(which, pd.values[r]);
    return base;
}

const void* Engine::value_ptr(int which, int vid, int* cs, const Geo** geo) {
    const PlanData& pd = plans_[which].data;
    if (cs) *cs = value_cs(pd, vid);
    if (geo) {
        int g = ctx_[which].vals[vid].geo;
        *geo = g >= 0 ? &ctx_[which].geos[g] : nullptr;
    }
    return vptr(which, vid);
}

int Engine::logits_vid(int which) const {
    const LoadedPlan& lp = plans_[which];
    if (!lp.loaded || lp.data.steps.empty() || plan_prec_[which] == VSE_PRECISION_FP16 || (cfg.flags & VSE_FLAG_NO_FAST_KERNELS)) return -1;
    const PlanData& pd = lp.data;
    const StepRec& s = pd.steps.back();
    if (s.op != OP_SOFTMAX || s.out != pd.hdr.output_vids[0] || s.ins[0] < 0 || pd.values[s.ins[0]].dtype != DT_ACT) return -1;
    int readers = 0;
    for (const StepRec& q : pd.steps)
        for (int i = 0; i < 4; i++) readers += q.ins[i] == s.ins[0];
    return readers == 1 ? s.ins[0] : -1;
}

// ------------------------------------------------------------------------------------------------
// execution
// ------------------------------------------------------------------------------------------------
void Engine::run_plan(int which, const std::vector<ImgTab>& in_tab, const uint8_t* input_dev, bool keep_all) {
    // geometry + arena plan are reused when the batch has the same shapes as the previous call (the usual case for
    // the detector: every frame of a video resizes to the same map)
    auto same_tab = [&](const std::vector<ImgTab>& t, bool ka) {
        bool same = ka == keep_all && t.size() == in_tab.size() && !t.empty();
        for (size_t i = 0; same && i < in_tab.size(); i++) same = t[i].h == in_tab[i].h && t[i].w == in_tab[i].w && t[i].vw == in_tab[i].vw;
        return same;
    };
    if (!(plans_[which].loaded && same_tab(last_tab_[which], last_keep_all_[which]))) {
        static const int kCtxCache = [] { const char* e = getenv("VSE_CTX_CACHE"); return e ? atoi(e) : 6; }();
        auto& store = ctx_store_[which];
        // park the active context, then look for a parked one with this geometry
        if (!last_tab_[which].empty() && kCtxCache > 0) {
            CtxSlot sl;
            sl.cx = std::move(ctx_[which]);
            sl.tab = std::move(last_tab_[which]);
            sl.keep_all = last_keep_all_[which];
            sl.stamp = ++ctx_clock_;
            store.push_back(std::move(sl));
            ctx_[which] = ExecContext{};
        } else {
            ctx_[which].tabs.release();
            ctx_[which].tc_gdev.release();
            ctx_[which] = ExecContext{};
        }
        last_tab_[which].clear();
        int hit = -1;
        for (size_t i = 0; i < store.size() && hit < 0; i++)
            if (same_tab(store[i].tab, store[i].keep_all)) hit = int(i);
        if (hit >= 0) {
            ctx_[which] = std::move(store[hit].cx);
            store.erase(store.begin() + hit);
        } else {
            while (int(store.size()) >= std::max(kCtxCache, 1)) {       // evict the least recently used geometry
                size_t old = 0;
                for (size_t i = 1; i < store.size(); i++)
                    if (store[i].stamp < store[old].stamp) old = i;
                store[old].cx.tabs.release();
                store[old].cx.tc_gdev.release();
                store.erase(store.begin() + old);
            }
            const void* arena_before = arena_[which].p;
            build_context(which, in_tab, keep_all);
            if (arena_[which].p != arena_before) {
                // the arena moved (it only ever grows): parked contexts hold tensor maps into the old allocation
                for (auto& sl : store) { sl.cx.tabs.release(); sl.cx.tc_gdev.release(); }
                store.clear();
            }
        }
        last_tab_[which] = in_tab;
        last_keep_all_[which] = keep_all;
    }
    // the input value is external memory: stash its pointer via a fake arena offset trick
    input_ptr_[which] = input_dev;
    exec_steps(which);
}

void Engine::exec_steps(int which, std::vector<cudaEvent_t>* step_events) {
    LoadedPlan& lp = plans_[which];
    const PlanData& pd = lp.data;
    ExecContext& cx = ctx_[which];
    const int prec = plan_prec_[which] == VSE_PRECISION_FP16 ? 0 : 1;   // storage type of activations: __half / float
    // the specialised kernels exist for fp16 storage and — as the same templates on float — for the fp32 tensor-core mode;
    // VSE_PRECISION_FP32 / TF32 keep the generic kernels
    const bool fast_ok = (prec == 0 || plan_prec_[which] == VSE_PRECISION_FP32_TC) && !(cfg.flags & VSE_FLAG_NO_FAST_KERNELS);
    const ImgTab* dtab = cx.tabs.as<ImgTab>();
    auto tab_of = [&](int vid) -> const ImgTab* {
        int g = cx.vals[vid].geo;
        return g >= 0 ? dtab + cx.geos[g].tab_off : nullptr;
    };
    auto geo_of = [&](int vid) -> const Geo& { return cx.geos[cx.vals[vid].geo]; };
    auto ptr_of = [&](int vid) -> void* {
        if (vid == pd.hdr.input_vid) return const_cast<uint8_t*>(input_ptr_[which]);
        return vptr(which, vid);
    };
    auto fill_epi = [&](const StepRec& s, const StepDev& d, Epilogue& e) {
        e.bias = d.bias;
        e.post_scale = s.p[P_HAS_POST] ? d.post_scale : nullptr;
        e.post_shift = s.p[P_HAS_POST] ? d.post_shift : nullptr;
        e.act = s.p[P_ACT];
        e.act2 = s.p[P_ACT2];
        e.hs_slope = s.f[F_HS_SLOPE];
        e.hs_offset = s.f[F_HS_OFFSET];
        if (s.p[P_HAS_RES]) {
            e.res = ptr_of(s.ins[1]);
            e.res_cs = value_cs(pd, s.ins[1]);
        }
    };
    if (step_events) cudaEventRecord((*step_events)[0], stream);
    cx.kind.assign(pd.steps.size(), 0);   // 0 generic kernel, 1 tcgen05 GEMM, 2 specialised kernel, 3 fused into the previous step
    for (size_t k = 0; k < pd.steps.size(); k++) {
        if (step_events && k > 0) cudaEventRecord((*step_events)[k], stream);
        const StepRec& s = pd.steps[k];
        const StepDev& d = lp.dev[k];
        const ValueRec& vo = pd.values[s.out];
        const int out_f32 = vo.dtype == DT_F32 && vo.kind == KIND_IMG;
        // concat gather: a run of CHSCALE / plain nearest-UPSAMPLE steps filling slices of one concat buffer -> one kernel
        if (fast_ok && !(cfg.flags & VSE_FLAG_NO_CONCAT_GATHER) && vo.alias_of >= 0 &&
            (s.op == OP_CHSCALE || (s.op == OP_UPSAMPLE && !s.p[P_HAS_ADD]))) {
            GatherSrc src[4];
            int n = 0;
            size_t j = k;
            for (; j < pd.steps.size() && n < 4; j++) {
                const StepRec& sj = pd.steps[j];
                const ValueRec& vj = pd.values[sj.out];
                const bool okop = sj.op == OP_CHSCALE || (sj.op == OP_UPSAMPLE && !sj.p[P_HAS_ADD]);
                if (!okop || vj.alias_of != vo.alias_of || cx.vals[sj.out].geo != cx.vals[s.out].geo || vj.dtype != DT_ACT) break;
                bool dep = false;   // a later member must not read an earlier member's output
                for (size_t q = k; q < j; q++) dep = dep || sj.ins[0] == pd.steps[q].out || sj.ins[1] == pd.steps[q].out;
                if (dep) break;
                GatherSrc& g = src[n++];
                g.in = ptr_of(sj.ins[0]);
                g.in_cs = value_cs(pd, sj.ins[0]);
                g.tin = tab_of(sj.ins[0]);
                g.scale_px = sj.op == OP_UPSAMPLE ? sj.p[P_SCALE] : 1;
                g.shift = -1;
                g.scale = sj.op == OP_CHSCALE ? static_cast<const float*>(ptr_of(sj.ins[1])) : nullptr;
                g.scale_c = sj.op == OP_CHSCALE ? pd.values[sj.ins[1]].channels : 0;
                g.residual = sj.op == OP_CHSCALE ? sj.p[P_RESIDUAL] : 0;
                g.out = ptr_of(sj.out);
                g.out_cs = value_cs(pd, sj.out);
                g.cvecs = pad8(vj.channels) / 8;
            }
            if (n >= 2) {
                std::sort(src, src + n, [](const GatherSrc& a, const GatherSrc& b) { return static_cast<const char*>(a.out) < static_cast<const char*>(b.out); });
                int mh = 0, mw = 0;
                for (const ImgTab& t : geo_of(s.out).tab) { mh = std::max(mh, t.h); mw = std::max(mw, t.w); }
                launch_concat_gather(src, n, tab_of(s.out), cx.n_img, mh, mw, stream, prec);
                launches++;
                cx.kind[k] = 2;
                for (size_t q = k + 1; q < j; q++) {
                    cx.kind[q] = 3;
                    if (step_events) cudaEventRecord((*step_events)[q], stream);
                }
                k = j - 1;
                continue;
            }
        }
        switch (s.op) {
            case OP_CONV: case OP_STEM: case OP_DWCONV: case OP_DECONV2: {
                ConvArgs a;
                a.in = ptr_of(s.ins[0]);
                a.out = ptr_of(s.out);
                a.w = d.w;
                fill_epi(s, d, a.epi);
                a.tin = tab_of(s.ins[0]);
                a.tout = tab_of(s.out);
                a.n_img = cx.n_img;
                a.max_out_pix = geo_of(s.out).max_pix;
                a.in_u8 = s.op == OP_STEM;
                a.cin_pad = a.in_u8 ? 4 : pad8(s.p[P_CIN]);
                a.in_cs = value_cs(pd, s.ins[0]);
                a.cout_store = pad8(s.p[P_COUT]);
                a.out_cs = value_cs(pd, s.out);
                a.w_ci = d.w_ci;
                a.w_co = d.w_co;
                a.kh = s.p[P_KH]; a.kw = s.p[P_KW]; a.sh = s.p[P_SH]; a.sw = s.p[P_SW]; a.ph = s.p[P_PH]; a.pw = s.p[P_PW];
                a.out_f32 = out_f32;
                for (int i = 0; i < 3; i++) { a.nscale[i] = pd.hdr.norm_scale[i]; a.nshift[i] = pd.hdr.norm_shift[i]; }
                const bool fast = fast_ok;
                auto max_units = [&](int tw) {   // max over images of out_h * ceil(out_w / tw)
                    int m = 0;
                    for (const ImgTab& t : geo_of(s.out).tab) m = std::max(m, t.h * ((t.w + tw - 1) / tw));
                    return m;
                };
                if (s.op == OP_DWCONV && cx.dwpw[k]) {
                    // fused depthwise -> pointwise: the following 1x1 convolution's kernel computes this step on the fly
                    const StepRec& c = pd.steps[k + 1];
                    const StepDev& dc = lp.dev[k + 1];
                    ConvArgs b;
                    b.out = ptr_of(c.out);
                    fill_epi(c, dc, b.epi);
                    b.cout_store = pad8(c.p[P_COUT]);
                    b.out_cs = value_cs(pd, c.out);
                    if (!launch_conv(which, int(k + 1), b, prec)) throw StateError{"fused depthwise -> pointwise lost its tensor-core plan"};
                    cx.kind[k] = 1;
                    k++;
                    cx.kind[k] = 3;
                    if (step_events) cudaEventRecord((*step_events)[k], stream);
                } else if (s.op == OP_DWCONV) {
                    if (out_f32) throw InvalidArg{"depthwise conv cannot produce a fetched output"};
                    // 0 strip, 1 shared-memory tiled, 2 (default) register tiled, 3 per layer: register tiled from
                    // VSE_DW_REG_MIN_C channels up, shared-memory tiled below
                    static const int dw_mode = [] { const char* e = getenv("VSE_DW_MODE"); return e ? atoi(e) : 2; }();
                    static const int dw_reg_min_c = [] { const char* e = getenv("VSE_DW_REG_MIN_C"); return e ? atoi(e) : 64; }();
                    int mh = 0, mw = 0;
                    for (const ImgTab& t : geo_of(s.out).tab) { mh = std::max(mh, t.h); mw = std::max(mw, t.w); }
                    const bool use_reg = dw_mode == 2 || (dw_mode == 3 && a.cin_pad >= dw_reg_min_c);
                    if (fast && !(cfg.flags & VSE_FLAG_NO_FAST_DW) && (use_reg || prec == 1) && launch_dwconv_reg(a, mh, mw, stream, prec)) cx.kind[k] = 2;
                    else if (fast && prec == 0 && !(cfg.flags & VSE_FLAG_NO_FAST_DW) && dw_mode >= 1 && launch_dwconv_tiled(a, mh, mw, stream)) cx.kind[k] = 2;
                    else if (fast && prec == 0 && !(cfg.flags & VSE_FLAG_NO_FAST_DW) && launch_dwconv_fast(a, max_units(4), stream)) cx.kind[k] = 2;
                    else launch_dwconv(a, prec, stream);
                } else if (s.op == OP_DECONV2) {
                    // DB head: deconv(C->C)+ReLU feeding only a deconv(C->1)+sigmoid that is a fetched fp32 map -> one kernel
                    bool fused = false;
                    if (fast && !(cfg.flags & VSE_FLAG_NO_FUSED_HEAD) && !last_keep_all_[which] && k + 1 < pd.steps.size()) {
                        const StepRec& s2 = pd.steps[k + 1];
                        const ValueRec& v2 = pd.values[s2.out];
                        if (s2.op == OP_DECONV2 && s2.ins[0] == s.out && pd.values[s.out].last_use == int(k + 1) &&
                            pd.values[s.out].alias_of < 0 && s.p[P_CIN] == s.p[P_COUT] && s2.p[P_CIN] == s.p[P_COUT] &&
                            s2.p[P_COUT] == 1 && v2.dtype == DT_F32 && v2.kind == KIND_IMG && s.p[P_ACT] == ACT_RELU &&
                            s2.p[P_ACT] == ACT_SIGMOID && !s.p[P_HAS_POST] && !s2.p[P_HAS_POST] && !s.p[P_HAS_RES] &&
                            !s2.p[P_HAS_RES] && s.p[P_ACT2] == ACT_NONE && s2.p[P_ACT2] == ACT_NONE) {
                            const StepDev& d2 = lp.dev[k + 1];
                            fused = launch_db_head_fused(a.in, a.in_cs, s.p[P_CIN], d.w, d.bias, d2.w, d2.bias,
                                                         static_cast<float*>(ptr_of(s2.out)), value_cs(pd, s2.out), a.tin,
                                                         tab_of(s2.out), cx.n_img, geo_of(s.ins[0]).max_pix, stream, prec);
                        }
                    }
                    if (fused) {
                        cx.kind[k] = 2;
                        k++;   // the second transposed convolution ran inside the fused kernel
                        cx.kind[k] = 3;
                        if (step_events) cudaEventRecord((*step_events)[k], stream);
                    } else {
                        bool on_tc = cx.tc_dc[k * 4].valid && !a.epi.res;
                        if (on_tc) {
                            // out[2y + dy][2x + dx] = act(W[dy][dx] x[y][x] + b): four 1x1 convolutions whose outputs interleave
                            const Geo& go = geo_of(s.out);
                            const int wo = go.tab[0].w, ho = go.tab[0].h;
                            const size_t es = prec == 0 ? 2 : 4;
                            for (int pos = 0; pos < 4 && on_tc; pos++) {
                                TcConv& t = cx.tc_dc[k * 4 + pos];
                                const int dy = pos >> 1, dx = pos & 1;
                                t.out = static_cast<char*>(a.out) + (size_t(dy) * wo + dx) * a.out_cs * es;
                                t.out_cs = a.out_cs;
                                t.n_store = a.cout_store;
                                t.o_px = 2LL * a.out_cs;
                                t.o_row = 2LL * wo * a.out_cs;
                                t.o_img = (long long)ho * wo * a.out_cs;
                                t.epi = a.epi;
                                if (!launch_conv_tc(t, sm_count, stream).empty()) {
                                    if (pos > 0) throw StateError{"transposed convolution failed after launching part of its positions"};
                                    on_tc = false;
                                }
                            }
                            if (on_tc) {
                                tc_launches += 4;
                                launches += 3;
                                cx.kind[k] = 1;
                            } else {
                                for (int pos = 0; pos < 4; pos++) cx.tc_dc[k * 4 + pos].valid = false;
                            }
                        }
                        if (!on_tc) launch_deconv2(a, s.p[P_COUT], prec, stream);
                    }
                } else if (s.op == OP_STEM && fast && !(cfg.flags & VSE_FLAG_NO_FAST_STEM) && launch_stem_fast(a, s.p[P_COUT], max_units(2), stream, prec)) {
                    cx.kind[k] = 2;
                } else if (s.op == OP_CONV && cx.se_conv[k] && cx.tc[k].valid) {
                    // fused residual squeeze-excite (see build_context): pool the conv INPUT, gate from W mean(x) + b, conv
                    // epilogue applies y + y * gate and writes the CHSCALE output
                    const StepRec &f1 = pd.steps[k + 2], &f2 = pd.steps[k + 3], &chs = pd.steps[k + 4];
                    const int cin = s.p[P_CIN], cout = s.p[P_COUT], cxp = pad8(cin);
                    const Geo& gx = geo_of(s.ins[0]);
                    int splits = std::max(1, std::min({64, gx.max_pix / 64, std::max((4 * sm_count + cx.n_img - 1) / std::max(cx.n_img, 1), gx.max_pix / 1024)}));
                    float* partial = reinterpret_cast<float*>(static_cast<char*>(arena_[which].p) + cx.scratch_off);
                    if (size_t(cx.n_img) * splits * cxp * sizeof(float) > cx.scratch_bytes) throw StateError{"squeeze-excite scratch too small"};
                    launch_gpool_partial(a.in, a.in_cs, cxp, a.tin, cx.n_img, partial, splits, prec, stream);
                    float* gate = static_cast<float*>(ptr_of(f2.out));
                    launch_se_gate(partial, splits, cxp, cout, f1.p[P_COUT], a.tin, lp.dev[k + 2].w_t, lp.dev[k + 2].bias, f1.p[P_ACT],
                                   f1.f[F_HS_SLOPE], f1.f[F_HS_OFFSET], lp.dev[k + 3].w_t, lp.dev[k + 3].bias, f2.p[P_ACT],
                                   f2.f[F_HS_SLOPE], f2.f[F_HS_OFFSET], gate, cx.n_img, stream, d.w, d.bias, cin, cxp, d.w_co);
                    a.out = ptr_of(chs.out);
                    a.out_cs = value_cs(pd, chs.out);
                    a.epi.gate = gate;
                    a.epi.gate_c = cout;
                    const int pack = std::max(cx.tc[k].pack, 1);
                    a.epi.gate_rows = gx.max_pix / pack;
                    if (gx.max_pix % pack) throw StateError{"squeeze-excite fusion: image size not a multiple of the pixel pack"};
                    if (!launch_conv(which, int(k), a, prec)) throw StateError{"squeeze-excite fusion lost its tensor-core plan"};
                    launches += 2;
                    cx.kind[k] = 1;
                    for (int j = 0; j < 4; j++) {
                        k++;
                        cx.kind[k] = 3;
                        if (step_events) cudaEventRecord((*step_events)[k], stream);
                    }
                } else {
                    if (out_f32) a.cout_store = s.p[P_COUT];
                    if (launch_conv(which, int(k), a, prec)) cx.kind[k] = 1;
                }
                launches++;
                break;
            }
            case OP_GPOOL: {
                const int cp = pad8(pd.values[s.ins[0]].channels);
                const Geo& g = geo_of(s.ins[0]);
                // enough CTAs to fill the machine (n_img x splits >= ~2 waves), at least 64 pixels per split
                int splits = std::max(1, std::min({64, g.max_pix / 64, std::max((4 * sm_count + cx.n_img - 1) / std::max(cx.n_img, 1), g.max_pix / 1024)}));
                float* partial = reinterpret_cast<float*>(static_cast<char*>(arena_[which].p) + cx.scratch_off);
                // squeeze-excite: GPOOL -> VECLIN -> VECLIN (each the sole consumer of the previous) -> one gate kernel
                bool fused = false;
                if (fast_ok && !(cfg.flags & VSE_FLAG_NO_SE_FUSION) && !last_keep_all_[which] && k + 2 < pd.steps.size()) {
                    const StepRec& f1 = pd.steps[k + 1];
                    const StepRec& f2 = pd.steps[k + 2];
                    auto plain = [](const StepRec& f) { return !f.p[P_HAS_POST] && !f.p[P_HAS_RES] && f.p[P_ACT2] == ACT_NONE; };
                    if (f1.op == OP_VECLIN && f2.op == OP_VECLIN && f1.ins[0] == s.out && f2.ins[0] == f1.out &&
                        pd.values[s.out].last_use == int(k + 1) && pd.values[f1.out].last_use == int(k + 2) && plain(f1) &&
                        plain(f2) && f1.p[P_CIN] == vo.channels && f2.p[P_CIN] == f1.p[P_COUT] && f2.p[P_COUT] == f1.p[P_CIN] &&
                        lp.dev[k + 1].bias && lp.dev[k + 2].bias && lp.dev[k + 1].w_t && lp.dev[k + 2].w_t && f1.p[P_COUT] <= 512 &&
                        f1.p[P_CIN] + f1.p[P_COUT] <= 8192) {
                        launch_gpool_partial(ptr_of(s.ins[0]), value_cs(pd, s.ins[0]), cp, tab_of(s.ins[0]), cx.n_img, partial,
                                             splits, prec, stream);
                        launch_se_gate(partial, splits, cp, f1.p[P_CIN], f1.p[P_COUT], tab_of(s.ins[0]), lp.dev[k + 1].w_t,
                                       lp.dev[k + 1].bias, f1.p[P_ACT], f1.f[F_HS_SLOPE], f1.f[F_HS_OFFSET], lp.dev[k + 2].w_t,
                                       lp.dev[k + 2].bias, f2.p[P_ACT], f2.f[F_HS_SLOPE], f2.f[F_HS_OFFSET],
                                       static_cast<float*>(ptr_of(f2.out)), cx.n_img, stream);
                        fused = true;
                    }
                }
                if (fused) {
                    cx.kind[k] = 2;
                    for (int j = 0; j < 2; j++) {
                        k++;
                        cx.kind[k] = 3;
                        if (step_events) cudaEventRecord((*step_events)[k], stream);
                    }
                } else {
                    launch_gpool(ptr_of(s.ins[0]), value_cs(pd, s.ins[0]), cp, tab_of(s.ins[0]), cx.n_img, g.max_pix, partial,
                                 splits, static_cast<float*>(ptr_of(s.out)), vo.channels, prec, stream);
                }
                launches += 2;
                break;
            }
            case OP_VECLIN: {
                Epilogue e;
                fill_epi(s, d, e);
                launch_veclin(static_cast<const float*>(ptr_of(s.ins[0])), s.p[P_CIN], static_cast<float*>(ptr_of(s.out)),
                              s.p[P_COUT], d.w, e, cx.n_img, stream);
                launches++;
                break;
            }
            case OP_CHSCALE: {
                const Geo& g = geo_of(s.out);
                launch_chscale(ptr_of(s.ins[0]), value_cs(pd, s.ins[0]), ptr_of(s.out), value_cs(pd, s.out), pad8(vo.channels),
                               static_cast<const float*>(ptr_of(s.ins[1])), pd.values[s.ins[1]].channels, s.p[P_RESIDUAL],
                               tab_of(s.out), cx.n_img, g.max_pix, prec, stream);
                launches++;
                break;
            }
            case OP_POOL: {
                launch_pool(ptr_of(s.ins[0]), value_cs(pd, s.ins[0]), ptr_of(s.out), value_cs(pd, s.out), pad8(vo.channels),
                            tab_of(s.ins[0]), tab_of(s.out), cx.n_img, geo_of(s.out).max_pix, s.p[P_KH], s.p[P_KW], s.p[P_SH],
                            s.p[P_SW], s.p[P_PH], s.p[P_PW], s.p[P_IS_MAX], s.p[P_EXCLUSIVE], prec, stream);
                launches++;
                break;
            }
            case OP_UPSAMPLE: {
                const void* add = s.p[P_HAS_ADD] ? ptr_of(s.ins[1]) : nullptr;
                int add_cs = s.p[P_HAS_ADD] ? value_cs(pd, s.ins[1]) : 0;
                launch_upsample(ptr_of(s.ins[0]), value_cs(pd, s.ins[0]), add, add_cs, ptr_of(s.out), value_cs(pd, s.out),
                                pad8(vo.channels), tab_of(s.ins[0]), tab_of(s.out), cx.n_img, geo_of(s.out).max_pix,
                                s.p[P_SCALE], prec, stream);
                launches++;
                break;
            }
            case OP_ADD: {
                launch_add(ptr_of(s.ins[0]), value_cs(pd, s.ins[0]), ptr_of(s.ins[1]), value_cs(pd, s.ins[1]), ptr_of(s.out),
                           value_cs(pd, s.out), pad8(vo.channels), vo.channels, geo_of(s.out).total, s.p[P_ACT], out_f32, prec,
                           stream);
                launches++;
                break;
            }
            case OP_ELTWISE: {
                if (vo.kind == KIND_VEC) throw InvalidArg{"standalone elementwise op on a pooled vector"};
                launch_eltwise(ptr_of(s.ins[0]), value_cs(pd, s.ins[0]), ptr_of(s.out), value_cs(pd, s.out), pad8(vo.channels),
                               vo.channels, geo_of(s.out).total, d.scale, d.shift, s.p[P_ACT], s.f[F_HS_SLOPE], s.f[F_HS_OFFSET],
                               out_f32, prec, stream);
                launches++;
                break;
            }
            case OP_COPY: {
                // copy ins[0] into channels [coff, coff+c) of the concat root `out`
                const int coff = s.p[P_SCALE], c = s.p[P_COUT];
                char* dst = static_cast<char*>(ptr_of(s.out)) + size_t(coff) * elt_size(which, vo);
                if (coff % 8 == 0) {
                    launch_copy(ptr_of(s.ins[0]), value_cs(pd, s.ins[0]), dst, value_cs(pd, s.out), pad8(c), geo_of(s.ins[0]).total,
                                prec, stream);
                } else {
                    const int fill = std::min(pad8(coff + c), value_cs(pd, s.out)) - coff;   // zero the pad channels behind the slice
                    launch_copy_unaligned(ptr_of(s.ins[0]), value_cs(pd, s.ins[0]), dst, value_cs(pd, s.out), c, fill,
                                          geo_of(s.ins[0]).total, prec, stream);
                }
                launches++;
                break;
            }
            case OP_LAYERNORM: {
                launch_layernorm(ptr_of(s.ins[0]), value_cs(pd, s.ins[0]), ptr_of(s.out), value_cs(pd, s.out), vo.channels,
                                 geo_of(s.out).total, d.gamma, d.beta, s.f[F_EPS], prec, stream);
                launches++;
                break;
            }
            case OP_ATTN: {
                const Geo& g = geo_of(s.out);
                if (s.p[P_DIM] > 32) throw InvalidArg{"attention head dim > 32"};
                if (attention_smem_bytes(g.max_pix, s.p[P_DIM]) > 200 * 1024) throw InvalidArg{"text line too long for attention kernel"};
                launch_attention(ptr_of(s.ins[0]), value_cs(pd, s.ins[0]), ptr_of(s.out), value_cs(pd, s.out), s.p[P_HEADS],
                                 s.p[P_DIM], s.f[F_QSCALE], tab_of(s.out), cx.n_img, g.max_pix, prec, stream);
                launches++;
                break;
            }
            case OP_SOFTMAX: {
                if (fold_final_softmax && !last_keep_all_[which] && k + 1 == pd.steps.size() && logits_vid(which) == s.ins[0]) {
                    cx.kind[k] = 3;      // the CTC decode works on the logits (Engine::fold_final_softmax)
                    break;
                }
                launch_softmax(ptr_of(s.ins[0]), value_cs(pd, s.ins[0]), static_cast<float*>(ptr_of(s.out)), vo.channels,
                               geo_of(s.out).total, prec, stream);
                launches++;
                break;
            }
            case OP_LSTM: {
                if (value_cs(pd, s.ins[0]) < s.p[P_CIN] || value_cs(pd, s.out) < s.p[P_COUT]) throw InvalidArg{"LSTM value strides"};
                VSE_CUDA(launch_lstm(ptr_of(s.ins[0]), value_cs(pd, s.ins[0]), ptr_of(s.out), value_cs(pd, s.out), d.w, tab_of(s.out),
                                     cx.n_img, s.p[P_SCALE], prec, stream));
                launches++;
                cx.kind[k] = 2;
                break;
            }
            default:
                throw InvalidArg{std::string("unsupported step ") + kOpNames[s.op]};
        }
    }
    if (step_events) cudaEventRecord((*step_events)[pd.steps.size()], stream);
    VSE_CUDA(cudaGetLastError());
}

// Re-executes the steps of the last run of plan `which` (same buffers, same geometry) with a CUDA event between
// consecutive steps; ms[k] = mean device time of step k over `reps` runs.  info[k] = {op, in_pixels, out_pixels,
// cin, cout, kh*kw, act_bytes, out_elt_bytes} lets the caller compute each step's algorithmic bytes / FLOPs.
int Engine::time_steps(int which, int reps, float* ms, int64_t* info, int cap) {
    if (which < 0 || which > 1 || !plans_[which].loaded || last_tab_[which].empty()) throw StateError{"no previous run of this plan"};
    const PlanData& pd = plans_[which].data;
    const int n = int(pd.steps.size());
    if (cap < n) throw CapacityError{"time_steps: capacity too small"};
    std::vector<cudaEvent_t> ev(n + 1);
    for (auto& e : ev) VSE_CUDA(cudaEventCreate(&e));
    std::vector<double> acc(n, 0.0);
    struct FoldGuard { bool& f; bool old; ~FoldGuard() { f = old; } } fold_guard{fold_final_softmax, fold_final_softmax};
    fold_final_softmax = which == VSE_PLAN_REC;     // what run_frames does
    exec_steps(which);  // warm
    VSE_CUDA(cudaStreamSynchronize(stream));
    for (int r = 0; r < reps; r++) {
        exec_steps(which, &ev);
        VSE_CUDA(cudaStreamSynchronize(stream));
        for (int k = 0; k < n; k++) {
            float t = 0.f;
            cudaEventElapsedTime(&t, ev[k], ev[k + 1]);
            acc[k] += t;
        }
    }
    for (auto& e : ev) cudaEventDestroy(e);
    const ExecContext& cx = ctx_[which];
    for (int k = 0; k < n; k++) {
        ms[k] = float(acc[k] / std::max(reps, 1));
        const StepRec& s = pd.steps[k];
        auto pixels = [&](int vid) -> int64_t {
            if (vid < 0) return 0;
            if (pd.values[vid].kind == KIND_VEC) return cx.n_img;
            int g = cx.vals[vid].geo;
            return g >= 0 ? cx.geos[g].total : 0;
        };
        int64_t* o = info + size_t(k) * 8;
        o[0] = s.op | (int64_t(k < int(cx.kind.size()) ? cx.kind[k] : 0) << 8);
        o[1] = pixels(s.ins[0]);
        o[2] = pixels(s.out);
        o[3] = s.ins[0] >= 0 ? pd.values[s.ins[0]].channels : 0;
        o[4] = pd.values[s.out].channels;
        o[5] = (s.op == OP_CONV || s.op == OP_STEM || s.op == OP_DWCONV) ? s.p[P_KH] * s.p[P_KW] : (s.op == OP_DECONV2 ? 4 : 1);
        o[6] = s.ins[0] >= 0 ? int64_t(elt_size(which, pd.values[s.ins[0]])) : 0;
        o[7] = int64_t(elt_size(which, pd.values[s.out]));
    }
    return n;
}

bool Engine::launch_conv(int which, int step, const ConvArgs& a, int prec) {
    TcConv& t = ctx_[which].tc[step];
    auto& groups = ctx_[which].tc_groups[step];
    if (!groups.empty()) {   // ragged KxK convolution: one tensor-core launch per run of equal-sized images
        const size_t es = plan_prec_[which] == VSE_PRECISION_FP16 ? 2 : 4;
        ExecContext& cx = ctx_[which];
        std::vector<TcConv*> gp;
        std::vector<long long> off;
        for (auto& g : groups) {
            TcConv& tg = g.tc;
            tg.out = static_cast<char*>(a.out) + size_t(g.pix_off) * a.out_cs * es;   // same-padding stride 1: output pixels = input pixels
            tg.out_cs = a.out_cs;
            tg.n_store = a.cout_store;
            tg.epi = a.epi;       // residual: the value's base pointer — the kernel adds the group's pixel offset
            gp.push_back(&tg);
            off.push_back(g.pix_off);
        }
        bool up = cx.tc_gup[step] != 0;
        const std::string err = launch_conv_tc_groups(gp.data(), off.data(), int(gp.size()), static_cast<char*>(cx.tc_gdev.p) + cx.tc_goff[step],
                                                      &up, sm_count, stream);
        cx.tc_gup[step] = up ? 1 : 0;
        if (err.empty()) {
            tc_launches++;
            return true;
        }
        groups.clear();   // e.g. an output view the TMA store cannot address: CUDA-core kernel from now on
    }
    (void)prec;
    if (t.valid) {   // built only for plans whose precision uses the tensor cores (fp16 / tf32)
        t.out = a.out;
        t.out_cs = a.out_cs;
        t.n_store = a.cout_store;
        t.epi = a.epi;
        if (t.pack > 0) {   // one GEMM row = `pack` consecutive pixels: their outputs (and residuals) are contiguous
            const StepDev& d = plans_[which].dev[step];
            t.out_cs = t.pack * a.out_cs;
            t.n_store = t.pack * a.out_cs;
            t.epi.bias = d.bias_pk;
            if (a.epi.post_scale) { t.epi.post_scale = d.post_scale_pk; t.epi.post_shift = d.post_shift_pk; }
            t.epi.res_cs = t.pack * a.epi.res_cs;
        }
        if (launch_conv_tc(t, sm_count, stream).empty()) {
            tc_launches++;
            return true;
        }
        t.valid = false;   // e.g. an output view the TMA store cannot address: CUDA-core kernel from now on
    }
    launch_conv_simt(a, prec, stream);
    return false;
}

int64_t Engine::get_value(int which, int vid, float* out, int64_t cap, int32_t* channels) {
    const PlanData& pd = plans_[which].data;
    if (vid < 0 || vid >= int(pd.values.size())) throw InvalidArg{"value id out of range"};
    const ValueRec& v = pd.values[vid];
    const ExecContext& cx = ctx_[which];
    int r = v.alias_of >= 0 ? v.alias_of : vid;
    if (vid != pd.hdr.input_vid && !cx.vals[r].live) return 0;  // fused away
    if (channels) *channels = v.channels;
    int64_t pixels = v.kind == KIND_VEC ? cx.n_img : cx.geos[cx.vals[vid].geo].total;
    int64_t n = pixels * v.channels;
    if (!out) return n;
    if (cap < n) throw InvalidArg{"output buffer too small"};
    if (v.dtype == DT_U8) throw InvalidArg{"cannot dump the uint8 input"};
    dbg_.reserve(size_t(n) * sizeof(float));
    const int prec = plan_prec_[which] == VSE_PRECISION_FP16 ? 0 : 1;   // storage type of activations: __half / float
    bool is_f32 = v.dtype == DT_F32 || v.kind == KIND_VEC;
    launch_to_float(vptr(which, vid), value_cs(pd, vid), is_f32, dbg_.as<float>(), v.channels, pixels, prec, stream);
    VSE_CUDA(cudaMemcpyAsync(out, dbg_.p, size_t(n) * sizeof(float), cudaMemcpyDeviceToHost, stream));
    VSE_CUDA(cudaStreamSynchronize(stream));
    return n;
}

}  // namespace vse
