// capi.cu — the extern "C" boundary (include/vse_b200.h). Nothing throws across it.
#include <cstring>
#include <new>
#include <vector>

#include "engine.h"

struct vse_engine {
    vse::Engine* impl;
};

static thread_local std::string g_create_error;

template <typename F>
static int guarded(vse_engine* e, F&& f) {
    try {
        // the engine's streams, buffers and kernels belong to its device: every entry point selects it, so calls from another
        // thread (or after an engine on another GPU was created in this thread) do not run against the wrong current device
        if (e && cudaSetDevice(e->impl->cfg.device) != cudaSuccess) throw vse::CudaError{"cudaSetDevice failed"};
        f();
        return VSE_OK;
    } catch (const vse::CudaError& ce) {
        if (e) e->impl->last_error = ce.msg; else g_create_error = ce.msg;
        return VSE_ERR_CUDA;
    } catch (const vse::InvalidArg& ia) {
        if (e) e->impl->last_error = ia.msg; else g_create_error = ia.msg;
        return VSE_ERR_INVALID;
    } catch (const vse::CapacityError& ce) {
        if (e) e->impl->last_error = ce.msg; else g_create_error = ce.msg;
        return VSE_ERR_CAPACITY;
    } catch (const vse::StateError& se) {
        if (e) e->impl->last_error = se.msg; else g_create_error = se.msg;
        return VSE_ERR_STATE;
    } catch (const std::bad_alloc&) {
        if (e) e->impl->last_error = "host allocation failed"; else g_create_error = "host allocation failed";
        return VSE_ERR_INVALID;
    } catch (const std::exception& ex) {
        if (e) e->impl->last_error = ex.what(); else g_create_error = ex.what();
        return VSE_ERR_INVALID;
    }
}

extern "C" {

void vse_default_config(vse_config* cfg) {
    if (!cfg) return;
    std::memset(cfg, 0, sizeof(*cfg));
    cfg->device = 0;
    cfg->precision = VSE_PRECISION_FP32_TC;   // the mode that reproduces the reference's boxes and texts
    cfg->det_limit_side_len = 960;
    cfg->det_thresh = 0.3f;
    cfg->det_box_thresh = 0.6f;
    cfg->det_unclip_ratio = 1.5f;
    cfg->det_max_candidates = 1000;
    cfg->rec_image_h = 48;
    cfg->rec_image_w = 320;
    cfg->rec_batch_num = 6;
    cfg->max_boxes_per_frame = 128;
    cfg->flags = 0;
}

int vse_abi_version(void) { return VSE_ABI_VERSION; }

int vse_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int vse_create(const vse_config* cfg, vse_engine** out) {
    if (!cfg || !out) {
        g_create_error = "null argument";
        return VSE_ERR_INVALID;
    }
    *out = nullptr;
    vse_engine* e = nullptr;
    int rc = guarded(nullptr, [&] {
        if (vse_device_count() == 0) throw vse::CudaError{"no CUDA device available (this engine has no CPU fallback)"};
        vse::Engine* impl = new vse::Engine(*cfg);
        e = new vse_engine{impl};
    });
    if (rc == VSE_ERR_CUDA && vse_device_count() == 0) rc = VSE_ERR_NO_DEVICE;
    if (rc == VSE_OK) *out = e;
    return rc;
}

void vse_destroy(vse_engine* e) {
    if (!e) return;
    cudaSetDevice(e->impl->cfg.device);
    delete e->impl;
    delete e;
}

const char* vse_last_error(const vse_engine* e) { return e ? e->impl->last_error.c_str() : g_create_error.c_str(); }

int vse_load_plan(vse_engine* e, int32_t which, const void* blob, size_t nbytes) {
    if (!e || !blob) return VSE_ERR_INVALID;
    return guarded(e, [&] { e->impl->load_plan(which, blob, nbytes); });
}

int vse_set_conv_input_ranges(vse_engine* e, int32_t which, const float* absmax, int32_t n_steps) {
    if (!e || !absmax || n_steps < 0) return VSE_ERR_INVALID;
    return guarded(e, [&] { e->impl->set_conv_input_ranges(which, absmax, n_steps); });
}

int vse_run(vse_engine* e, const uint8_t* const* frames, const int32_t* h, const int32_t* w, const int32_t* row_stride,
            int32_t n_frames, int32_t mem_kind, vse_result* out) {
    if (!e || !out || (n_frames > 0 && (!frames || !h || !w))) return VSE_ERR_INVALID;
    return guarded(e, [&] { e->impl->run_frames(frames, h, w, row_stride, n_frames, mem_kind, out, false); });
}

int vse_prefetch(vse_engine* e, const uint8_t* const* frames, const int32_t* h, const int32_t* w, const int32_t* row_stride,
                 int32_t n_frames, int32_t mem_kind) {
    if (!e || (n_frames > 0 && (!frames || !h || !w))) return VSE_ERR_INVALID;
    return guarded(e, [&] { e->impl->prefetch_frames(frames, h, w, row_stride, n_frames, mem_kind); });
}

int vse_det_only(vse_engine* e, const uint8_t* const* frames, const int32_t* h, const int32_t* w, const int32_t* row_stride,
                 int32_t n_frames, int32_t mem_kind, vse_result* out) {
    if (!e || !out || (n_frames > 0 && (!frames || !h || !w))) return VSE_ERR_INVALID;
    return guarded(e, [&] { e->impl->run_frames(frames, h, w, row_stride, n_frames, mem_kind, out, true); });
}

int64_t vse_launch_count(const vse_engine* e) { return e ? e->impl->launches : 0; }

int64_t vse_tc_launch_count(const vse_engine* e) { return e ? e->impl->tc_launches : 0; }

int vse_debug_run_plan(vse_engine* e, int32_t which, const uint8_t* const* images, int32_t n, int32_t h, const int32_t* w,
                       const int32_t* valid_w, int32_t keep_all) {
    if (!e || !images || !w || n <= 0) return VSE_ERR_INVALID;
    return guarded(e, [&] { e->impl->debug_run_plan(which, images, n, h, w, valid_w, keep_all != 0); });
}

int64_t vse_debug_get_value(vse_engine* e, int32_t which, int32_t vid, float* out, int64_t capacity, int32_t* channels) {
    if (!e) return VSE_ERR_INVALID;
    int64_t n = 0;
    int rc = guarded(e, [&] { n = e->impl->get_value(which, vid, out, capacity, channels); });
    return rc == VSE_OK ? n : rc;
}

int vse_debug_resize_bilinear(vse_engine* e, const uint8_t* src, int32_t sh, int32_t sw, int32_t src_stride, uint8_t* dst_bgrx,
                              int32_t dh, int32_t dw) {
    if (!e || !src || !dst_bgrx) return VSE_ERR_INVALID;
    return guarded(e, [&] { e->impl->debug_resize(src, sh, sw, src_stride, dst_bgrx, dh, dw); });
}

int vse_debug_db_postprocess(vse_engine* e, const float* prob, int32_t rh, int32_t rw, int32_t src_h, int32_t src_w, float* quads,
                             float* scores, int32_t capacity, int32_t* n_out) {
    if (!e || !prob || !quads || !scores || !n_out) return VSE_ERR_INVALID;
    return guarded(e, [&] { e->impl->debug_db_post(prob, rh, rw, src_h, src_w, quads, scores, capacity, n_out); });
}

int vse_debug_crop(vse_engine* e, const uint8_t* frame, int32_t h, int32_t w, const float* quad, uint8_t* out_bgr, int32_t capacity,
                   int32_t* out_h, int32_t* out_w) {
    if (!e || !frame || !quad || !out_bgr || !out_h || !out_w) return VSE_ERR_INVALID;
    return guarded(e, [&] { e->impl->debug_crop(frame, h, w, quad, out_bgr, capacity, out_h, out_w); });
}

int vse_debug_time_steps(vse_engine* e, int32_t which, int32_t reps, float* ms, int64_t* info, int32_t capacity) {
    if (!e || !ms || !info) return VSE_ERR_INVALID;
    int n = 0;
    int rc = guarded(e, [&] { n = e->impl->time_steps(which, reps, ms, info, capacity); });
    return rc == VSE_OK ? n : rc;
}

}  // extern "C"
