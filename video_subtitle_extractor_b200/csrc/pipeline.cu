// pipeline.cu — the per-frame det + rec pipeline behind vse_run / vse_det_only, all on the device:
//   frames (u8 BGR) -> det resize -> det plan -> DB post-process -> crops -> rec resize -> rec plan -> CTC decode.
// Mirrors paddleocr 2.10 TextSystem.__call__ as driven by reference backend/tools/ocr.py:27 and
// TextDetector.__call__ as driven by backend/tools/subtitle_detect.py:25.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "engine.h"
#include "preproc.cuh"

namespace vse {

struct Engine::Pipeline {
    DevBuf frames;     // uploaded source frames
    DevBuf det_in;     // resized BGRX det inputs
    DevBuf jobs;       // ResizeJob table
    DevBuf scratch;
};

Engine::~Engine() {
    for (int i = 0; i < 2; i++) {
        plans_[i].weights.release();
        ctx_[i].tabs.release();
        arena_[i].release();
    }
    dbg_.release();
    pin_.release();
    if (pipe_) {
        pipe_->frames.release();
        pipe_->det_in.release();
        pipe_->jobs.release();
        pipe_->scratch.release();
        delete pipe_;
    }
    if (stream) cudaStreamDestroy(stream);
}

void Engine::debug_run_plan(int which, const uint8_t* const* images, int n, int h, const int32_t* w, const int32_t* valid_w,
                            bool keep_all) {
    if (!pipe_) pipe_ = new Pipeline();
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
    if (!pipe_) pipe_ = new Pipeline();
    if (stride <= 0) stride = sw * 3;
    pipe_->frames.reserve(size_t(sh) * stride);
    pipe_->det_in.reserve(size_t(dh) * dw * 4);
    pipe_->jobs.reserve(sizeof(ResizeJob));
    VSE_CUDA(cudaMemcpyAsync(pipe_->frames.p, src, size_t(sh) * stride, cudaMemcpyHostToDevice, stream));
    ResizeJob j{pipe_->frames.as<uint8_t>(), sh, sw, stride, 3, dh, dw, dw, 0};
    VSE_CUDA(cudaMemcpyAsync(pipe_->jobs.p, &j, sizeof(j), cudaMemcpyHostToDevice, stream));
    dim3 grid((dh * dw + 255) / 256, 1);
    resize_bilinear_u8_kernel<<<grid, 256, 0, stream>>>(pipe_->jobs.as<ResizeJob>(), pipe_->det_in.as<uint8_t>(), dh * dw);
    launches++;
    VSE_CUDA(cudaGetLastError());
    VSE_CUDA(cudaMemcpyAsync(dst, pipe_->det_in.p, size_t(dh) * dw * 4, cudaMemcpyDeviceToHost, stream));
    VSE_CUDA(cudaStreamSynchronize(stream));
}

void Engine::run_frames(const uint8_t* const*, const int32_t*, const int32_t*, const int32_t*, int, int, vse_result*, bool) {
    throw InvalidArg{"vse_run: pipeline not built yet"};
}
void Engine::debug_db_post(const float*, int, int, int, int, float*, float*, int, int*) {
    throw InvalidArg{"db post-process not built yet"};
}
void Engine::debug_crop(const uint8_t*, int, int, const float*, uint8_t*, int, int*, int*) {
    throw InvalidArg{"crop not built yet"};
}

}  // namespace vse
