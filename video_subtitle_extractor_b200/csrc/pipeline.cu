// pipeline.cu — the per-frame det + rec pipeline behind vse_run / vse_det_only, all on the device:
//   frames (u8 BGR) -> det resize -> det plan -> DB post-process -> crops -> rec resize -> rec plan -> CTC decode.
// Mirrors paddleocr 2.10 TextSystem.__call__ as driven by reference backend/tools/ocr.py:27 and
// TextDetector.__call__ as driven by backend/tools/subtitle_detect.py:25.
// Two host synchronisations per call: after the DB post-process (the number and size of the text lines decide the
// recogniser's ragged batch) and after the CTC decode.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <mutex>

#include "dbpost_core.cuh"
#include "pdl.cuh"
#include "engine.h"
#include "postproc.cuh"
#include "preproc.cuh"

namespace vse {

struct Engine::Pipeline {
    // frame staging is double-buffered: vse_prefetch copies later batches (copy stream) into the buffer that no pending
    // prefetch occupies while the current batch computes.  vse_run is synchronous, so a buffer is free as soon as the
    // run that used it has returned.
    struct Pending {
        int buf;
        std::vector<const uint8_t*> src, dev;
        std::vector<int> h, w, stride;
    };
    DevBuf fbuf[2], dbg_frames, det_in, jobs, det_frames;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t fdone[2] = {nullptr, nullptr};
    std::vector<Pending> pending;             // oldest first, at most 2
    // vse_prefetch may be called from a producer thread while vse_run computes on the consumer thread: `mu` guards the
    // pending list and the choice of staging buffer; `busy` is the buffer the running batch reads (never handed out)
    std::mutex mu;
    int busy = -1;
    DevBuf labels, slot_of, n_comp, roots, bbox, order, cand, status, n_boxes, quads, scores, blk_fg, blabels, bopen, blk_listed, ckey;
    DevBuf cubic_tab, crop_jobs, crop_buf, rec_jobs, rec_in;
    DevBuf ctc_meta, ctc_ids, ctc_len, ctc_score;
    PinnedBuf h_in, h_out;
    cudaEvent_t ev[9] = {};
    bool ev_ready = false, tab_ready = false;
    std::vector<DevBuf*> all() {
        return {&fbuf[0], &fbuf[1], &dbg_frames, &det_in, &jobs, &det_frames, &labels, &slot_of, &n_comp, &roots, &bbox, &order, &cand, &status, &blk_fg, &blabels, &bopen, &blk_listed, &ckey,
                &n_boxes, &quads, &scores, &cubic_tab, &crop_jobs, &crop_buf, &rec_jobs, &rec_in, &ctc_meta, &ctc_ids,
                &ctc_len, &ctc_score};
    }
};

Engine::~Engine() {
    for (int i = 0; i < 2; i++) {
        plans_[i].weights.release();
        ctx_[i].tabs.release();
        ctx_[i].tc_gdev.release();
        for (auto& sl : ctx_store_[i]) { sl.cx.tabs.release(); sl.cx.tc_gdev.release(); }
        arena_[i].release();
    }
    dbg_.release();
    pin_.release();
    if (pipe_) {
        for (DevBuf* b : pipe_->all()) b->release();
        pipe_->h_in.release();
        pipe_->h_out.release();
        if (pipe_->ev_ready)
            for (auto& e : pipe_->ev) cudaEventDestroy(e);
        for (auto& e : pipe_->fdone)
            if (e) cudaEventDestroy(e);
        if (pipe_->copy_stream) cudaStreamDestroy(pipe_->copy_stream);
        delete pipe_;
    }
    if (stream) cudaStreamDestroy(stream);
}

// ------------------------------------------------------------------------------------------------
// host-side shape rules (upstream DetResizeForTest / TextRecognizer batching)
// ------------------------------------------------------------------------------------------------
static void det_resize_shape(int h, int w, int limit, int* rh, int* rw) {
    double ratio = 1.0;
    if (std::max(h, w) > limit) ratio = h > w ? double(limit) / h : double(limit) / w;
    int a = int(h * ratio), b = int(w * ratio);
    a = std::max(int(std::nearbyint(a / 32.0) * 32), 32);  // Python round(): half to even
    b = std::max(int(std::nearbyint(b / 32.0) * 32), 32);
    *rh = a;
    *rw = b;
}

struct CropPlan {
    int frame, box;       // source frame / box row inside the frame
    int cw, ch, rot;      // warped crop size, rot90 flag
    int H, W;             // crop size after the optional rot90
    bool direct;          // axis-aligned integer box: the crop is a plain copy of frame pixels
    int x0, y0;
    int img_w, resized_w; // recogniser input width (padded) / resized width
    long long crop_off;   // BGRX pixel offset in crop_buf (non-direct only)
    long long rec_off;    // BGRX pixel offset in rec_in
};

static DbWorkspace make_ws(Engine::Pipeline* p);

void Engine::ensure_pipeline() {
    if (!pipe_) pipe_ = new Pipeline();
    if (!pipe_->ev_ready) {
        for (auto& e : pipe_->ev) VSE_CUDA(cudaEventCreate(&e));
        pipe_->ev_ready = true;
    }
    if (!pipe_->copy_stream) {
        VSE_CUDA(cudaStreamCreateWithFlags(&pipe_->copy_stream, cudaStreamNonBlocking));
        for (auto& e : pipe_->fdone) VSE_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    if (!pipe_->tab_ready) {
        std::vector<short> tab(32 * 32 * 16);
        dbpost::build_cubic_table(tab.data());
        pipe_->cubic_tab.reserve(tab.size() * sizeof(short));
        VSE_CUDA(cudaMemcpy(pipe_->cubic_tab.p, tab.data(), tab.size() * sizeof(short), cudaMemcpyHostToDevice));
        pipe_->tab_ready = true;
    }
}

static DbWorkspace make_ws(Engine::Pipeline* p) {
    DbWorkspace ws;
    ws.labels = p->labels.as<int>();
    ws.slot_of = p->slot_of.as<int>();
    ws.blabels = p->blabels.as<int>();
    ws.bopen = p->bopen.as<int>();
    ws.blk_listed = p->blk_listed.as<int>();
    ws.ckey = p->ckey.as<int>();
    ws.fg_count = p->blk_fg.as<int>();
    ws.fg_list = p->blk_fg.as<int>() + 4;
    ws.n_comp = p->n_comp.as<int>();
    ws.roots = p->roots.as<int>();
    ws.bbox = p->bbox.as<int>();
    ws.order = p->order.as<int>();
    ws.cand = p->cand.as<float>();
    ws.status = p->status.as<int>();
    ws.n_boxes = p->n_boxes.as<int>();
    ws.quads = p->quads.as<float>();
    ws.scores = p->scores.as<float>();
    return ws;
}

// Runs the DB post-process on a probability map already on the device and brings n_boxes / quads / scores to h_out.
// Layout of h_out: int n_boxes[n] | int status[n] | float scores[n][mb] | float quads[n][mb][8]
void Engine::db_post_device(const float* prob, const std::vector<DetFrame>& frames, bool reading_order) {
    Pipeline* P = pipe_;
    const int n = int(frames.size());
    int max_rh = 0, max_rw = 0;
    size_t total = 0;
    for (auto& f : frames) {
        max_rh = std::max(max_rh, f.rh);
        max_rw = std::max(max_rw, f.rw);
        total = std::max(total, size_t(f.map_off) + size_t(f.rh) * f.rw);
    }
    if (max_rh > 2048) throw InvalidArg{"detection map taller than 2048 rows"};
    if (max_rw > 32 * 1023 || n > 2047) throw InvalidArg{"detection batch beyond the post-process limits (2047 frames, 32736 columns)"};
    const int mc = std::min(std::max(cfg.det_max_candidates, 1), kSlotCap);
    cfg.max_boxes_per_frame = std::min(std::max(cfg.max_boxes_per_frame, 1), mc);   // upstream keeps at most max_candidates boxes
    int mb = cfg.max_boxes_per_frame;
    P->labels.reserve(total * sizeof(int));
    P->slot_of.reserve(total * sizeof(int));
    P->blk_fg.reserve((size_t(n) * ((max_rh + 7) / 8) * ((max_rw + 31) / 32) + 4) * sizeof(int));
    P->blabels.reserve(total * sizeof(int));
    P->bopen.reserve(total * sizeof(int));
    P->blk_listed.reserve(size_t(n) * ((max_rh + 7) / 8) * ((max_rw + 31) / 32) * sizeof(int));
    P->ckey.reserve(size_t(n) * kSlotCap * sizeof(int));
    P->n_comp.reserve(n * sizeof(int));
    P->status.reserve(n * sizeof(int));
    P->roots.reserve(size_t(n) * kSlotCap * sizeof(int));
    P->bbox.reserve(size_t(n) * kSlotCap * 4 * sizeof(int));
    P->order.reserve(size_t(n) * kSlotCap * sizeof(int));
    P->cand.reserve(size_t(n) * mc * 10 * sizeof(float));
    P->n_boxes.reserve(n * sizeof(int));
    P->quads.reserve(size_t(n) * mb * 8 * sizeof(float));
    P->scores.reserve(size_t(n) * mb * sizeof(float));
    P->det_frames.reserve(n * sizeof(DetFrame));
    P->h_in.reserve(n * sizeof(DetFrame));
    std::memcpy(P->h_in.p, frames.data(), n * sizeof(DetFrame));
    launch_upload(P->det_frames.p, P->h_in.p, n * sizeof(DetFrame), stream);
    DbParams dp{cfg.det_thresh, cfg.det_box_thresh, cfg.det_unclip_ratio, mc, mb, reading_order ? 1 : 0};
    launch_db_postprocess(prob, P->det_frames.as<DetFrame>(), frames.data(), n, max_rh, max_rw, dp, make_ws(P), stream, &launches);
    VSE_CUDA(cudaGetLastError());
    char* h = nullptr;
    const int* status = nullptr;
    for (;;) {
        const size_t bytes = size_t(n) * (2 * sizeof(int) + mb * sizeof(float) * 9);
        P->h_out.reserve(bytes);
        h = P->h_out.as<char>();
        VSE_CUDA(cudaMemcpyAsync(h, P->n_boxes.p, n * sizeof(int), cudaMemcpyDeviceToHost, stream));
        VSE_CUDA(cudaMemcpyAsync(h + n * sizeof(int), P->status.p, n * sizeof(int), cudaMemcpyDeviceToHost, stream));
        VSE_CUDA(cudaMemcpyAsync(h + 2 * n * sizeof(int), P->scores.p, size_t(n) * mb * sizeof(float), cudaMemcpyDeviceToHost, stream));
        VSE_CUDA(cudaMemcpyAsync(h + 2 * n * sizeof(int) + size_t(n) * mb * sizeof(float), P->quads.p, size_t(n) * mb * 8 * sizeof(float),
                                 cudaMemcpyDeviceToHost, stream));
        VSE_CUDA(cudaStreamSynchronize(stream));
        status = reinterpret_cast<const int*>(h + n * sizeof(int));
        bool overflow = false;
        for (int i = 0; i < n; i++) overflow |= (status[i] & 2) != 0;
        if (!overflow || mb >= mc) break;
        // a frame holds more boxes than the per-frame rows (dense text, credits): the candidates are still on the device,
        // so only the compaction is repeated with more rows — up to max_candidates, which is all upstream ever keeps
        mb = std::min(mc, mb * 4);
        cfg.max_boxes_per_frame = mb;
        P->quads.reserve(size_t(n) * mb * 8 * sizeof(float));
        P->scores.reserve(size_t(n) * mb * sizeof(float));
        dp.max_boxes = mb;
        launch_db_compact(n, dp, make_ws(P), stream, &launches);
        VSE_CUDA(cudaGetLastError());
    }
    for (int i = 0; i < n; i++) {
        if (status[i] & 4)
            throw StateError{"frame " + std::to_string(i) + ": the detection map holds non-finite values — this model's activations "
                             "exceed the fp16 range (e.g. V4/ch_det: LK-PAN outputs reach 1.5e5); create the engine with "
                             "VSE_FLAG_DET_FP32 (or VSE_PRECISION_FP32) for it"};
        if (status[i] & 1) throw InvalidArg{"frame " + std::to_string(i) + ": more than 4096 connected components in the detection map"};
        if (status[i] & 2) throw CapacityError{"frame " + std::to_string(i) + ": more boxes than max_boxes_per_frame"};   // unreachable: rows == max_candidates
    }
}

// ------------------------------------------------------------------------------------------------
// vse_run / vse_det_only
// ------------------------------------------------------------------------------------------------
static void frame_strides(const uint8_t* const* frames, const int32_t* h, const int32_t* w, const int32_t* stride, int n,
                          std::vector<int>& fstride) {
    for (int i = 0; i < n; i++) {
        if (h[i] <= 0 || w[i] <= 0 || !frames[i]) throw InvalidArg{"empty frame"};
        if (h[i] + w[i] < 64) throw InvalidArg{"frame smaller than 64 px in total is not supported"};
        fstride[i] = stride && stride[i] > 0 ? stride[i] : w[i] * 3;
        if (fstride[i] < w[i] * 3) throw InvalidArg{"row stride smaller than 3*width"};
    }
}

// host frames -> one staging buffer, 256-byte aligned slots, async on `st`
static void stage_frames(const uint8_t* const* frames, const int32_t* h, const int32_t* w, const std::vector<int>& fstride, int n,
                         DevBuf& buf, cudaStream_t st, std::vector<const uint8_t*>& fdev) {
    size_t total = 0;
    for (int i = 0; i < n; i++) total += (size_t(h[i]) * fstride[i] + 255) & ~size_t(255);
    buf.reserve(total);
    // frames that follow each other in host memory (a decoded batch in one pinned block) go up in ONE copy: a 1080p batch
    // of 32 is 32 driver calls otherwise, ~0.5 ms of host time per step
    size_t off = 0, run_off = 0, run_bytes = 0;
    const uint8_t* run_src = nullptr;
    auto flush = [&]() {
        if (run_bytes) VSE_CUDA(cudaMemcpyAsync(buf.as<uint8_t>() + run_off, run_src, run_bytes, cudaMemcpyHostToDevice, st));
        run_bytes = 0;
    };
    for (int i = 0; i < n; i++) {
        // a strided view (sub-area of a larger frame) ends with its last pixel, not with a full row pitch: reading
        // h * stride bytes would run past the end of the caller's frame for an area that touches its bottom edge
        const size_t bytes = size_t(h[i] - 1) * fstride[i] + size_t(w[i]) * 3;
        if (run_bytes && frames[i] == run_src + run_bytes && off == run_off + run_bytes) {
            run_bytes += bytes;
        } else {
            flush();
            run_src = frames[i];
            run_off = off;
            run_bytes = bytes;
        }
        fdev[i] = buf.as<uint8_t>() + off;
        off += (bytes + 255) & ~size_t(255);
    }
    flush();
}

// vse_prefetch: start the host->device copy of the NEXT batch on the copy stream while the current vse_run computes.
// The following vse_run / vse_det_only with the same frame pointers and sizes finds the frames resident.
void Engine::prefetch_frames(const uint8_t* const* frames, const int32_t* h, const int32_t* w, const int32_t* stride, int n,
                             int mem_kind) {
    if (n <= 0 || mem_kind == VSE_MEM_DEVICE) return;
    if (mem_kind < VSE_MEM_HOST || mem_kind > VSE_MEM_DEVICE) throw InvalidArg{"bad mem_kind"};
    ensure_pipeline();
    Pipeline* P = pipe_;
    std::vector<int> fstride(n);
    frame_strides(frames, h, w, stride, n, fstride);
    std::lock_guard<std::mutex> lock(P->mu);
    // buffers a prefetch may use: not the one the running batch reads; an unused older prefetch gives way
    auto holds = [&](int b) { for (auto& q : P->pending) if (q.buf == b) return true; return false; };
    int buf = -1;
    for (int b = 0; b < 2 && buf < 0; b++)
        if (b != P->busy && !holds(b)) buf = b;
    if (buf < 0) {
        VSE_CUDA(cudaStreamSynchronize(P->copy_stream));
        for (size_t q = 0; q < P->pending.size() && buf < 0; q++)
            if (P->pending[q].buf != P->busy) {
                buf = P->pending[q].buf;
                P->pending.erase(P->pending.begin() + q);
            }
        if (buf < 0) throw StateError{"no staging buffer free for vse_prefetch"};
    }
    Pipeline::Pending pe;
    pe.buf = buf;
    pe.dev.assign(n, nullptr);
    // growing the buffer frees the old allocation (implicit device sync); the run that last used it has returned
    stage_frames(frames, h, w, fstride, n, P->fbuf[pe.buf], P->copy_stream, pe.dev);
    VSE_CUDA(cudaEventRecord(P->fdone[pe.buf], P->copy_stream));
    pe.src.assign(frames, frames + n);
    pe.h.assign(h, h + n);
    pe.w.assign(w, w + n);
    pe.stride = fstride;
    P->pending.push_back(std::move(pe));
}

void Engine::run_frames(const uint8_t* const* frames, const int32_t* h, const int32_t* w, const int32_t* stride, int n, int mem_kind,
                        vse_result* out, bool det_only) {
    if (n < 0) throw InvalidArg{"negative frame count"};
    if (!out->n_boxes) throw InvalidArg{"result.n_boxes is null"};
    for (int i = 0; i < 8; i++) out->timings_ms[i] = 0.f;
    if (n == 0) return;
    if (!plans_[0].loaded) throw StateError{"detection plan not loaded"};
    if (!det_only && !plans_[1].loaded) throw StateError{"recognition plan not loaded"};
    if (mem_kind < VSE_MEM_HOST || mem_kind > VSE_MEM_DEVICE) throw InvalidArg{"bad mem_kind"};
    ensure_pipeline();
    Pipeline* P = pipe_;
    VSE_CUDA(cudaEventRecord(P->ev[0], stream));

    // 1. frames -> device (already there when vse_prefetch staged exactly this batch)
    std::vector<const uint8_t*> fdev(n);
    std::vector<int> fstride(n);
    frame_strides(frames, h, w, stride, n, fstride);
    struct BusyGuard {      // the staging buffer of this batch is released on every way out of the call
        Pipeline* P;
        ~BusyGuard() { std::lock_guard<std::mutex> lock(P->mu); P->busy = -1; }
    } busy_guard{P};
    if (mem_kind == VSE_MEM_DEVICE) {
        for (int i = 0; i < n; i++) fdev[i] = frames[i];
    } else {
        std::lock_guard<std::mutex> lock(P->mu);
        int hit = -1;
        for (size_t q = 0; q < P->pending.size() && hit < 0; q++) {
            const Pipeline::Pending& pe = P->pending[q];
            bool same = int(pe.src.size()) == n;
            for (int i = 0; same && i < n; i++)
                same = pe.src[i] == frames[i] && pe.h[i] == h[i] && pe.w[i] == w[i] && pe.stride[i] == fstride[i];
            if (same) hit = int(q);
        }
        if (hit >= 0) {
            VSE_CUDA(cudaStreamWaitEvent(stream, P->fdone[P->pending[hit].buf], 0));
            fdev = P->pending[hit].dev;
            P->busy = P->pending[hit].buf;
            P->pending.erase(P->pending.begin() + hit);
        } else {
            int buf = 0;
            if (P->pending.size() >= 2) {          // both buffers hold batches nobody asked for: drop them
                VSE_CUDA(cudaStreamSynchronize(P->copy_stream));
                P->pending.clear();
            } else if (P->pending.size() == 1) {
                buf = 1 - P->pending[0].buf;
            }
            P->busy = buf;
            stage_frames(frames, h, w, fstride, n, P->fbuf[buf], stream, fdev);
        }
    }
    VSE_CUDA(cudaEventRecord(P->ev[1], stream));

    // 2. detector pre-process: OpenCV-exact bilinear resize to (rh, rw), BGRX
    std::vector<ImgTab> det_tab(n);
    {
        std::vector<ResizeJob> jobs(n);
        long long off = 0;
        int max_pix = 0;
        for (int i = 0; i < n; i++) {
            int rh, rw;
            det_resize_shape(h[i], w[i], cfg.det_limit_side_len, &rh, &rw);
            det_tab[i] = ImgTab{0, rh, rw, rw};
            jobs[i] = ResizeJob{fdev[i], h[i], w[i], fstride[i], 3, rh, rw, rw, off};
            off += (long long)rh * rw;
            max_pix = std::max(max_pix, rh * rw);
        }
        P->det_in.reserve(size_t(off) * 4);
        P->jobs.reserve(n * sizeof(ResizeJob));
        P->h_in.reserve(n * sizeof(ResizeJob));
        std::memcpy(P->h_in.p, jobs.data(), n * sizeof(ResizeJob));
        launch_upload(P->jobs.p, P->h_in.p, n * sizeof(ResizeJob), stream);   // (kernel parameters: h_in is free again)
        dim3 grid((max_pix + 255) / 256, n);
        pdl_launch(resize_bilinear_u8_kernel, grid, 256, 0, stream, P->jobs.as<ResizeJob>(), P->det_in.as<uint8_t>(), max_pix);
        launches++;
        VSE_CUDA(cudaGetLastError());
    }
    VSE_CUDA(cudaEventRecord(P->ev[2], stream));

    // 3. detector network
    run_plan(VSE_PLAN_DET, det_tab, P->det_in.as<uint8_t>(), false);
    VSE_CUDA(cudaEventRecord(P->ev[3], stream));

    // 4. DB post-process on the device
    const int det_vid = plans_[0].data.hdr.output_vids[0];
    int cs = 0;
    const Geo* geo = nullptr;
    const float* prob = static_cast<const float*>(value_ptr(VSE_PLAN_DET, det_vid, &cs, &geo));
    if (!geo || cs != 1) throw InvalidArg{"detection plan output is not a 1-channel map"};
    std::vector<DetFrame> dfr(n);
    for (int i = 0; i < n; i++) dfr[i] = DetFrame{geo->tab[i].off, geo->tab[i].h, geo->tab[i].w, h[i], w[i]};
    db_post_device(prob, dfr, !det_only);
    const int mb = cfg.max_boxes_per_frame;      // (db_post_device grows it when a frame holds more boxes)
    VSE_CUDA(cudaEventRecord(P->ev[4], stream));
    const char* hb = P->h_out.as<char>();
    const int* nb = reinterpret_cast<const int*>(hb);
    const float* hscore = reinterpret_cast<const float*>(hb + 2 * n * sizeof(int));
    const float* hquad = hscore + size_t(n) * mb;
    int total_boxes = 0;
    for (int i = 0; i < n; i++) total_boxes += nb[i];
    if (total_boxes > out->box_capacity) throw CapacityError{"result.box_capacity too small for " + std::to_string(total_boxes) + " boxes"};
    if (total_boxes > 0 && (!out->quads || !out->det_score)) throw InvalidArg{"result.quads / det_score is null"};
    {
        int row = 0;
        for (int i = 0; i < n; i++) {
            out->n_boxes[i] = nb[i];
            for (int b = 0; b < nb[i]; b++, row++) {
                std::memcpy(out->quads + size_t(row) * 8, hquad + (size_t(i) * mb + b) * 8, 8 * sizeof(float));
                out->det_score[row] = hscore[size_t(i) * mb + b];
                if (out->id_len) out->id_len[row] = 0;
                if (out->rec_score) out->rec_score[row] = 0.f;
                if (out->rec_width) out->rec_width[row] = 0;
            }
        }
    }
    auto finish_timings = [&](int last) {
        VSE_CUDA(cudaEventRecord(P->ev[last], stream));
        VSE_CUDA(cudaEventSynchronize(P->ev[last]));
        for (int i = 0; i < last && i < 7; i++) cudaEventElapsedTime(&out->timings_ms[i], P->ev[i], P->ev[i + 1]);
        cudaEventElapsedTime(&out->timings_ms[7], P->ev[0], P->ev[last]);
    };
    if (det_only || total_boxes == 0) {
        finish_timings(5);
        return;
    }
    if (!out->ids || !out->id_len || !out->rec_score) throw InvalidArg{"result.ids / id_len / rec_score is null"};

    // 5. crops and recogniser pre-process (upstream batching rules decide each crop's padded width)
    std::vector<CropPlan> crops;
    crops.reserve(total_boxes);
    const int RH = cfg.rec_image_h, RW = cfg.rec_image_w, RB = std::max(cfg.rec_batch_num, 1);
    long long crop_pix = 0, rec_pix = 0;
    int max_crop_pix = 0, max_rec_pix = 0;
    for (int i = 0; i < n; i++) {
        const int first = int(crops.size());
        for (int b = 0; b < nb[i]; b++) {
            const geom::P2f* q = reinterpret_cast<const geom::P2f*>(hquad + (size_t(i) * mb + b) * 8);
            CropPlan c{};
            c.frame = i;
            c.box = b;
            geom::crop_size(q, &c.cw, &c.ch);
            if (c.cw <= 0 || c.ch <= 0) throw InvalidArg{"degenerate text box"};
            c.rot = (c.ch * 1.0 / c.cw >= 1.5) ? 1 : 0;
            c.H = c.rot ? c.cw : c.ch;
            c.W = c.rot ? c.ch : c.cw;
            c.x0 = int(q[0].x);
            c.y0 = int(q[0].y);
            c.direct = !c.rot && q[0].y == q[1].y && q[1].x == q[2].x && q[2].y == q[3].y && q[3].x == q[0].x &&
                       q[1].x - q[0].x == float(c.cw) && q[3].y - q[0].y == float(c.ch);
            crops.push_back(c);
        }
        // TextRecognizer.__call__: sort by aspect ratio, batches of rec_batch_num, per-batch padded width
        const int m = nb[i];
        std::vector<int> order(m);
        std::vector<double> ratio(m);
        for (int k = 0; k < m; k++) {
            order[k] = k;
            ratio[k] = crops[first + k].W / double(crops[first + k].H);
        }
        std::stable_sort(order.begin(), order.end(), [&](int a, int b2) { return ratio[a] < ratio[b2]; });
        for (int beg = 0; beg < m; beg += RB) {
            const int end = std::min(m, beg + RB);
            double max_wh = double(RW) / RH;
            for (int k = beg; k < end; k++) max_wh = std::max(max_wh, ratio[order[k]]);
            const int img_w = int(RH * max_wh);
            for (int k = beg; k < end; k++) {
                CropPlan& c = crops[first + order[k]];
                const int want = int(std::ceil(RH * ratio[order[k]]));
                c.img_w = img_w;
                c.resized_w = want > img_w ? img_w : want;
            }
        }
    }
    for (auto& c : crops) {
        if (!c.direct) {
            c.crop_off = crop_pix;
            crop_pix += (long long)c.H * c.W;
            max_crop_pix = std::max(max_crop_pix, c.H * c.W);
        }
        c.rec_off = rec_pix;
        rec_pix += (long long)RH * c.img_w;
        max_rec_pix = std::max(max_rec_pix, RH * c.resized_w);
    }
    const int nC = int(crops.size());
    std::vector<ImgTab> rec_tab(nC);
    {
        std::vector<CropJob> cj;
        std::vector<ResizeJob> rj(nC);
        P->crop_buf.reserve(size_t(std::max<long long>(crop_pix, 1)) * 4);
        P->rec_in.reserve(size_t(rec_pix) * 4);
        for (int k = 0; k < nC; k++) {
            const CropPlan& c = crops[k];
            rec_tab[k] = ImgTab{0, RH, c.img_w, c.resized_w};
            if (c.direct) {
                const uint8_t* src = fdev[c.frame] + size_t(c.y0) * fstride[c.frame] + size_t(c.x0) * 3;
                rj[k] = ResizeJob{src, c.ch, c.cw, fstride[c.frame], 3, RH, c.resized_w, c.img_w, c.rec_off};
            } else {
                CropJob j{};
                j.frame = fdev[c.frame];
                j.fh = h[c.frame];
                j.fw = w[c.frame];
                j.stride = fstride[c.frame];
                j.pix = 3;
                const geom::P2f* q = reinterpret_cast<const geom::P2f*>(hquad + (size_t(c.frame) * mb + c.box) * 8);
                geom::rect_to_quad_homography(q, c.cw, c.ch, j.M);
                j.cw = c.cw;
                j.ch = c.ch;
                j.rot90 = c.rot;
                j.dst_off = c.crop_off;
                cj.push_back(j);
                rj[k] = ResizeJob{P->crop_buf.as<uint8_t>() + size_t(c.crop_off) * 4, c.H, c.W, c.W * 4, 4, RH, c.resized_w, c.img_w, c.rec_off};
            }
        }
        const size_t cj_bytes = cj.size() * sizeof(CropJob), rj_bytes = rj.size() * sizeof(ResizeJob);
        P->h_in.reserve(cj_bytes + rj_bytes + 16);
        if (!cj.empty()) std::memcpy(P->h_in.p, cj.data(), cj_bytes);
        std::memcpy(P->h_in.as<char>() + cj_bytes, rj.data(), rj_bytes);
        P->crop_jobs.reserve(std::max<size_t>(cj_bytes, 16));
        P->rec_jobs.reserve(rj_bytes);
        if (!cj.empty()) launch_upload(P->crop_jobs.p, P->h_in.p, cj_bytes, stream);
        launch_upload(P->rec_jobs.p, P->h_in.as<char>() + cj_bytes, rj_bytes, stream);
        if (!cj.empty()) {
            launch_crops(P->crop_jobs.as<CropJob>(), int(cj.size()), max_crop_pix, P->cubic_tab.as<short>(), P->crop_buf.as<uint8_t>(), stream);
            launches++;
        }
        dim3 grid((max_rec_pix + 255) / 256, nC);
        pdl_launch(resize_bilinear_u8_kernel, grid, 256, 0, stream, P->rec_jobs.as<ResizeJob>(), P->rec_in.as<uint8_t>(), max_rec_pix);
        launches++;
        VSE_CUDA(cudaGetLastError());
        // no synchronisation here: launch_upload copies its source into the kernel's parameters at launch time, so h_in
        // (and the engine's table staging, run_plan) may be rewritten as soon as the call returns
    }
    VSE_CUDA(cudaEventRecord(P->ev[5], stream));

    // 6. recogniser network on the ragged batch (its final softmax is folded into the CTC decode when the plan allows it)
    {
        struct FoldGuard { bool& f; ~FoldGuard() { f = false; } } fold_guard{fold_final_softmax};
        fold_final_softmax = true;
        run_plan(VSE_PLAN_REC, rec_tab, P->rec_in.as<uint8_t>(), false);
    }
    VSE_CUDA(cudaEventRecord(P->ev[6], stream));

    // 7. CTC greedy decode + results
    {
        const int rec_vid = plans_[1].data.hdr.output_vids[0];
        const int lg_vid = logits_vid(VSE_PLAN_REC);      // >= 0: the softmax step was skipped, decode from the logits
        int C = 0, cs = 0;
        const Geo* rg = nullptr;
        const float* probs = static_cast<const float*>(value_ptr(VSE_PLAN_REC, rec_vid, &C, &rg));
        cs = C;
        if (lg_vid >= 0) probs = static_cast<const float*>(value_ptr(VSE_PLAN_REC, lg_vid, &cs, nullptr));
        if (!rg) throw InvalidArg{"recognition plan output has no geometry"};
        int max_t = 1;
        std::vector<int> meta(2 * nC);
        for (int k = 0; k < nC; k++) {
            meta[k] = rg->tab[k].off;
            meta[nC + k] = rg->tab[k].h * rg->tab[k].w;
            max_t = std::max(max_t, meta[nC + k]);
        }
        P->ctc_meta.reserve(meta.size() * sizeof(int));
        P->ctc_ids.reserve(size_t(nC) * max_t * sizeof(int));
        P->ctc_len.reserve(nC * sizeof(int));
        P->ctc_score.reserve(nC * sizeof(float));
        P->h_in.reserve(meta.size() * sizeof(int));
        std::memcpy(P->h_in.p, meta.data(), meta.size() * sizeof(int));
        launch_upload(P->ctc_meta.p, P->h_in.p, meta.size() * sizeof(int), stream);
        launch_ctc_decode(probs, C, cs, lg_vid >= 0, P->ctc_meta.as<int>(), P->ctc_meta.as<int>() + nC, nC, max_t, P->ctc_ids.as<int>(),
                          P->ctc_len.as<int>(), P->ctc_score.as<float>(), stream);
        launches++;
        VSE_CUDA(cudaGetLastError());
        const size_t ids_bytes = size_t(nC) * max_t * sizeof(int);
        P->h_out.reserve(ids_bytes + nC * 8 + size_t(n) * (2 * sizeof(int) + mb * sizeof(float) * 9));
        // keep the det results (already consumed into `out`) – reuse h_out from the start
        char* ho = P->h_out.as<char>();
        VSE_CUDA(cudaMemcpyAsync(ho, P->ctc_ids.p, ids_bytes, cudaMemcpyDeviceToHost, stream));
        VSE_CUDA(cudaMemcpyAsync(ho + ids_bytes, P->ctc_len.p, nC * sizeof(int), cudaMemcpyDeviceToHost, stream));
        VSE_CUDA(cudaMemcpyAsync(ho + ids_bytes + nC * sizeof(int), P->ctc_score.p, nC * sizeof(float), cudaMemcpyDeviceToHost, stream));
        VSE_CUDA(cudaStreamSynchronize(stream));
        const int* hid = reinterpret_cast<const int*>(ho);
        const int* hlen = reinterpret_cast<const int*>(ho + ids_bytes);
        const float* hsc = reinterpret_cast<const float*>(ho + ids_bytes + nC * sizeof(int));
        for (int k = 0; k < nC; k++) {
            if (hlen[k] < 0)
                throw StateError{"text line " + std::to_string(k) + ": the recogniser produced non-finite class probabilities — its "
                                 "activations exceed the fp16 range; create the engine with VSE_PRECISION_FP32 or VSE_PRECISION_TF32"};
            if (hlen[k] > out->max_text_len) throw CapacityError{"result.max_text_len too small for a line of " + std::to_string(hlen[k]) + " symbols"};
            std::memcpy(out->ids + size_t(k) * out->max_text_len, hid + size_t(k) * max_t, hlen[k] * sizeof(int));
            out->id_len[k] = hlen[k];
            out->rec_score[k] = hsc[k];
            if (out->rec_width) out->rec_width[k] = crops[k].img_w;
        }
    }
    finish_timings(7);
}

// ------------------------------------------------------------------------------------------------
// debug hooks
// ------------------------------------------------------------------------------------------------
void Engine::debug_run_plan(int which, const uint8_t* const* images, int n, int h, const int32_t* w, const int32_t* valid_w,
                            bool keep_all) {
    ensure_pipeline();
    std::vector<ImgTab> tab(n);
    size_t total = 0;
    for (int i = 0; i < n; i++) {
        if (w[i] <= 0 || h <= 0) throw InvalidArg{"empty image"};
        tab[i] = ImgTab{0, h, w[i], valid_w ? valid_w[i] : w[i]};
        total += size_t(h) * w[i];
    }
    pipe_->det_in.reserve(total * 4);
    size_t off = 0;
    for (int i = 0; i < n; i++) {
        size_t bytes = size_t(h) * w[i] * 4;
        VSE_CUDA(cudaMemcpyAsync(pipe_->det_in.as<uint8_t>() + off, images[i], bytes, cudaMemcpyHostToDevice, stream));
        off += bytes;
    }
    run_plan(which, tab, pipe_->det_in.as<uint8_t>(), keep_all);
    VSE_CUDA(cudaStreamSynchronize(stream));
}

void Engine::debug_resize(const uint8_t* src, int sh, int sw, int stride, uint8_t* dst, int dh, int dw) {
    ensure_pipeline();
    if (stride <= 0) stride = sw * 3;
    pipe_->dbg_frames.reserve(size_t(sh) * stride);
    pipe_->det_in.reserve(size_t(dh) * dw * 4);
    pipe_->jobs.reserve(sizeof(ResizeJob));
    VSE_CUDA(cudaMemcpyAsync(pipe_->dbg_frames.p, src, size_t(sh) * stride, cudaMemcpyHostToDevice, stream));
    ResizeJob j{pipe_->dbg_frames.as<uint8_t>(), sh, sw, stride, 3, dh, dw, dw, 0};
    VSE_CUDA(cudaMemcpyAsync(pipe_->jobs.p, &j, sizeof(j), cudaMemcpyHostToDevice, stream));
    dim3 grid((dh * dw + 255) / 256, 1);
    pdl_launch(resize_bilinear_u8_kernel, grid, 256, 0, stream, pipe_->jobs.as<ResizeJob>(), pipe_->det_in.as<uint8_t>(), dh * dw);
    launches++;
    VSE_CUDA(cudaGetLastError());
    VSE_CUDA(cudaMemcpyAsync(dst, pipe_->det_in.p, size_t(dh) * dw * 4, cudaMemcpyDeviceToHost, stream));
    VSE_CUDA(cudaStreamSynchronize(stream));
}

void Engine::debug_db_post(const float* prob, int rh, int rw, int src_h, int src_w, float* quads, float* scores, int cap, int* n_out) {
    ensure_pipeline();
    pipe_->det_in.reserve(size_t(rh) * rw * sizeof(float));
    VSE_CUDA(cudaMemcpyAsync(pipe_->det_in.p, prob, size_t(rh) * rw * sizeof(float), cudaMemcpyHostToDevice, stream));
    std::vector<DetFrame> fr{DetFrame{0, rh, rw, src_h, src_w}};
    db_post_device(pipe_->det_in.as<float>(), fr, false);
    const char* hb = pipe_->h_out.as<char>();
    const int nb = *reinterpret_cast<const int*>(hb);
    if (nb > cap) throw CapacityError{"debug_db_post: capacity too small"};
    const float* hscore = reinterpret_cast<const float*>(hb + 2 * sizeof(int));
    const float* hquad = hscore + cfg.max_boxes_per_frame;
    std::memcpy(scores, hscore, nb * sizeof(float));
    std::memcpy(quads, hquad, size_t(nb) * 8 * sizeof(float));
    *n_out = nb;
}

void Engine::debug_crop(const uint8_t* frame, int h, int w, const float* quad, uint8_t* out, int cap, int* oh, int* ow) {
    ensure_pipeline();
    const geom::P2f* q = reinterpret_cast<const geom::P2f*>(quad);
    CropJob j{};
    geom::crop_size(q, &j.cw, &j.ch);
    if (j.cw <= 0 || j.ch <= 0) throw InvalidArg{"degenerate quad"};
    j.rot90 = (j.ch * 1.0 / j.cw >= 1.5) ? 1 : 0;
    const int H = j.rot90 ? j.cw : j.ch, W = j.rot90 ? j.ch : j.cw;
    if (size_t(H) * W * 3 > size_t(cap)) throw CapacityError{"debug_crop: capacity too small"};
    pipe_->dbg_frames.reserve(size_t(h) * w * 3);
    VSE_CUDA(cudaMemcpyAsync(pipe_->dbg_frames.p, frame, size_t(h) * w * 3, cudaMemcpyHostToDevice, stream));
    j.frame = pipe_->dbg_frames.as<uint8_t>();
    j.fh = h; j.fw = w; j.stride = w * 3; j.pix = 3;
    geom::rect_to_quad_homography(q, j.cw, j.ch, j.M);
    j.dst_off = 0;
    pipe_->crop_jobs.reserve(sizeof(CropJob));
    pipe_->crop_buf.reserve(size_t(H) * W * 4);
    VSE_CUDA(cudaMemcpyAsync(pipe_->crop_jobs.p, &j, sizeof(j), cudaMemcpyHostToDevice, stream));
    launch_crops(pipe_->crop_jobs.as<CropJob>(), 1, H * W, pipe_->cubic_tab.as<short>(), pipe_->crop_buf.as<uint8_t>(), stream);
    launches++;
    VSE_CUDA(cudaGetLastError());
    std::vector<uint8_t> tmp(size_t(H) * W * 4);
    VSE_CUDA(cudaMemcpyAsync(tmp.data(), pipe_->crop_buf.p, tmp.size(), cudaMemcpyDeviceToHost, stream));
    VSE_CUDA(cudaStreamSynchronize(stream));
    for (size_t i = 0; i < size_t(H) * W; i++) {
        out[i * 3] = tmp[i * 4];
        out[i * 3 + 1] = tmp[i * 4 + 1];
        out[i * 3 + 2] = tmp[i * 4 + 2];
    }
    *oh = H;
    *ow = W;
}

}  // namespace vse
