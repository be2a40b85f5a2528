/*
 * vse_b200.h — C ABI of libvse_b200.so, the B200-native per-frame subtitle-OCR engine.
 *
 * This is the drop-in boundary for the reference's det/rec predictor calls.  Every entry
 * point below replaces a call the reference makes into paddleocr / paddle.inference
 * (all paths relative to the reference repository, eritpchy/video-subtitle-extractor):
 *
 *   vse_create + vse_load_plan  <->  PaddleOCR(...) construction        backend/tools/ocr.py:88-113
 *                                    TextDetector(args) construction    backend/tools/subtitle_detect.py:10-22
 *   vse_run                     <->  self.recogniser(image, cls=False)  backend/tools/ocr.py:27
 *   vse_det_only                <->  self.text_detector(img)            backend/tools/subtitle_detect.py:24-26
 *   vse_device_count            <->  paddle.is_compiled_with_cuda() /
 *                                    paddle.static.cuda_places()        backend/tools/hardware_accelerator.py:26-32
 *   vse_last_error              <->  the exception the worker prints    backend/tools/subtitle_ocr.py:155-157
 *
 * Plain C types only; the caller owns every buffer; no callbacks; one engine per
 * (process, GPU); calls on one engine are not re-entrant (the reference calls predict()
 * from exactly one thread per process: backend/tools/subtitle_ocr.py:231, backend/main.py:307).
 * All functions return 0 on success, a negative vse_status otherwise; the message is
 * available from vse_last_error() (engine may be NULL for creation failures).
 */
#ifndef VSE_B200_H
#define VSE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VSE_ABI_VERSION 1

typedef struct vse_engine vse_engine;

typedef enum {
    VSE_OK = 0,
    VSE_ERR_INVALID = -1,   /* bad argument / malformed plan          */
    VSE_ERR_CUDA = -2,      /* CUDA runtime failure                    */
    VSE_ERR_CAPACITY = -3,  /* caller-provided result buffers too small */
    VSE_ERR_STATE = -4,     /* plan not loaded                         */
    VSE_ERR_NO_DEVICE = -5  /* no CUDA device: there is no CPU fallback */
} vse_status;

enum { VSE_MEM_HOST = 0, VSE_MEM_PINNED = 1, VSE_MEM_DEVICE = 2 };
enum { VSE_PLAN_DET = 0, VSE_PLAN_REC = 1 };
enum {
    VSE_PRECISION_FP16 = 0, /* fp16 activations, fp32 accumulate: tcgen05 kind::f16 convolutions + specialised kernels   */
    VSE_PRECISION_FP32 = 1, /* fp32 activations, CUDA-core kernels only: the exact-parity mode                             */
    VSE_PRECISION_TF32 = 2, /* fp32 activations (full range), convolutions on tcgen05 kind::tf32 (10-bit mantissa operands) */
    VSE_PRECISION_FP32_TC = 3 /* fp32 activations; convolutions on tcgen05 kind::f16 with both operands split into fp16 hi + lo
                               * (three MMAs per product, ~22-bit operands): the fp32 engine's results at tensor-core speed —
                               * the mode that meets the parity bar on real video and the one bench.py times.  Needs activations
                               * inside the fp16 range (like VSE_PRECISION_FP16; violations are reported, never silent) */
};
enum {
    VSE_FLAG_NO_TENSOR_CORES = 1, /* vse_config.flags: keep every conv on the CUDA-core kernels (A/B checks)      */
    VSE_FLAG_NO_FAST_KERNELS = 2, /* keep depthwise / stem / DB-head / SE steps on the generic kernels (A/B checks) */
    /* finer A/B switches (bisecting numerical differences); each disables one specialised path */
    VSE_FLAG_NO_FUSED_HEAD = 4, VSE_FLAG_NO_SE_FUSION = 8, VSE_FLAG_NO_ROWBOX = 16, VSE_FLAG_NO_FAST_DW = 32,
    VSE_FLAG_NO_FAST_STEM = 64, VSE_FLAG_NO_PIXEL_PACK = 128,
    VSE_FLAG_NO_CONCAT_GATHER = 256, VSE_FLAG_NO_HALO = 512, VSE_FLAG_NO_SE_CONV = 4096,
    /* not an A/B switch: run the DETECTOR plan with fp32 activations whatever vse_config.precision says (V4/ch_det, the
     * accurate-mode detector of backend/tools/paddle_model_config.py:60,70, exceeds the fp16 range) */
    VSE_FLAG_DET_FP32 = 1024,
    VSE_FLAG_DET_TF32 = 2048,  /* same, with the detector's convolutions on the tf32 tensor-core path */
    /* the DETECTOR plan in VSE_PRECISION_FP32_TC whatever vse_config.precision says: with precision = VSE_PRECISION_FP16 the
     * recogniser keeps fp16 activations (its bar is CER <= 1e-3 on class ids, which fp16 meets on the reference's videos) while
     * the detector — whose 0.3 threshold crossing needs ~1e-5 on the probability map — keeps fp32 activations */
    VSE_FLAG_DET_FP32_TC = 8192,
    /* opt-in experiment (fp32 tensor-core mode, equal-sized images): compute a stride-1 3x3 / 5x5 depthwise convolution inside the
     * kernel of the 1x1 convolution that follows it (its output is never stored).  Bit-identical results, but measured SLOWER
     * than the two separate kernels on B200 (the depthwise arithmetic gets 4 warps per SM instead of 16) */
    VSE_FLAG_DWPW_FUSION = 16384
};

/* Mirrors the knobs the reference passes to PaddleOCR / TextDetector (ocr.py:91-113) and the
 * upstream defaults it relies on (utility.parse_args(): SURVEY.md Appendix D.8 item 3). */
typedef struct {
    int32_t device;              /* CUDA ordinal                                  */
    int32_t precision;           /* VSE_PRECISION_*: storage type of activations  */
    int32_t det_limit_side_len;  /* 960                                           */
    float   det_thresh;          /* 0.3                                           */
    float   det_box_thresh;      /* 0.6                                           */
    float   det_unclip_ratio;    /* 1.5                                           */
    int32_t det_max_candidates;  /* 1000                                          */
    int32_t rec_image_h;         /* 48 (V2 models: 32)                            */
    int32_t rec_image_w;         /* 320                                           */
    int32_t rec_batch_num;       /* 6   (config.recBatchNumber, ocr.py:99)        */
    int32_t max_boxes_per_frame; /* device-side candidate capacity per frame      */
    int32_t flags;               /* VSE_FLAG_* bits                               */
} vse_config;

/* Results, caller-allocated.  Boxes of frame f occupy rows [sum(n_boxes[0..f)), +n_boxes[f]). */
typedef struct {
    int32_t  box_capacity;   /* rows available in the per-box arrays                       */
    int32_t  max_text_len;   /* columns of ids[]                                           */
    int32_t* n_boxes;        /* [n_frames]                                                 */
    float*   quads;          /* [cap][4][2] clockwise from top-left, frame pixels          */
    float*   det_score;      /* [cap] mean probability inside the (pre-unclip) box         */
    int32_t* ids;            /* [cap][max_text_len] CTC class ids (0 = blank never stored) */
    int32_t* id_len;         /* [cap]                                                      */
    float*   rec_score;      /* [cap] mean of the kept max-probabilities (0 if none)       */
    int32_t* rec_width;      /* [cap] padded width each crop was recognised at             */
    float    timings_ms[8];  /* 0 h2d, 1 det-pre, 2 det-net, 3 det-post, 4 crop, 5 rec-net, 6 ctc+d2h, 7 total */
} vse_result;

/* Defaults: the reference's knobs as listed above, precision = VSE_PRECISION_FP32_TC (the parity mode), flags = 0. */
void vse_default_config(vse_config* cfg);
int  vse_abi_version(void);
int  vse_device_count(void);

int  vse_create(const vse_config* cfg, vse_engine** out);
void vse_destroy(vse_engine* e);
const char* vse_last_error(const vse_engine* e);

/* Packed plan (steps + fp32 weights) produced by video_subtitle_extractor_b200/plan.py from the
 * reference's inference.pdmodel/.pdiparams; on multi-GPU jobs rank 0 builds it and broadcasts the bytes. */
int  vse_load_plan(vse_engine* e, int32_t which, const void* blob, size_t nbytes);

/* Optional, VSE_PRECISION_FP32_TC only: absmax[k] = the largest |activation| the INPUT of plan step k is expected to hold
 * (calibrated offline: tools/calibrate_ranges.py -> video_subtitle_extractor_b200/calibration/<model>.json; <= 0 = unknown).
 * The engine scales each convolution's operand rows by the power of two that brings absmax[k] just below 2^14 before the
 * fp16 hi/lo split (4x headroom to the fp16 limit; an input that still overflows is reported as non-finite output, never
 * silently).  Without the call every convolution uses a conservative 2^2.  Must follow vse_load_plan of that plan. */
int  vse_set_conv_input_ranges(vse_engine* e, int32_t which, const float* absmax, int32_t n_steps);

/* det + rec on n_frames BGR uint8 HWC frames (row_stride in bytes, may be NULL for tight rows).
 * TextSystem order: boxes sorted top-to-bottom / left-to-right per frame.
 * A frame may be a VIEW of a larger image (the reference's half-frame / subtitle-area crop, frame_preprocess,
 * backend/tools/subtitle_ocr.py:270-289): pass the address of the view's first pixel, the view's h / w and the full image's
 * row pitch; the engine reads (h - 1) * row_stride + 3 * w bytes from there and returns boxes in the view's coordinates. */
int  vse_run(vse_engine* e, const uint8_t* const* frames, const int32_t* h, const int32_t* w,
             const int32_t* row_stride, int32_t n_frames, int32_t mem_kind, vse_result* out);

/* Optional: start the host->device copy of the NEXT batch (pinned or pageable host frames) on the engine's copy stream
 * and return at once, so that it overlaps the vse_run of the current batch.  The next vse_run / vse_det_only called with
 * the same frame pointers, sizes and strides uses the staged copy; any other call simply ignores it.  The host frames
 * must stay valid and unchanged until that call returns.  The reference has no counterpart (it feeds one frame per call,
 * backend/tools/subtitle_ocr.py:30); its producer thread (:164-208) is where a caller would issue this. */
int  vse_prefetch(vse_engine* e, const uint8_t* const* frames, const int32_t* h, const int32_t* w,
                  const int32_t* row_stride, int32_t n_frames, int32_t mem_kind);

/* det only: TextDetector order (contour order), quads/det_score/n_boxes filled. */
int  vse_det_only(vse_engine* e, const uint8_t* const* frames, const int32_t* h, const int32_t* w,
                  const int32_t* row_stride, int32_t n_frames, int32_t mem_kind, vse_result* out);

/* Number of kernels this engine has launched since creation (bench.py's gpu_launches). */
int64_t vse_launch_count(const vse_engine* e);
/* ... of which launches of the tcgen05/TMA implicit-GEMM kernel (gemm_tc.cu). */
int64_t vse_tc_launch_count(const vse_engine* e);

/* ---- test / profiling hooks (used by tests/ and bench.py only) --------------------------- */

/* Run one loaded plan on a batch of uint8 BGRX (4 bytes/pixel) host images of equal height and
 * per-image width w[i] (valid_w[i] <= w[i]: columns beyond are "zero after normalisation").
 * keep_all != 0 keeps every intermediate value alive for vse_debug_get_value. */
int  vse_debug_run_plan(vse_engine* e, int32_t which, const uint8_t* const* images, int32_t n, int32_t h,
                        const int32_t* w, const int32_t* valid_w, int32_t keep_all);
/* Copy value `vid` of the last debug run to host as dense float32 [pixels][channels];
 * returns the number of floats (or a negative status); out may be NULL to query the size. */
int64_t vse_debug_get_value(vse_engine* e, int32_t which, int32_t vid, float* out, int64_t capacity,
                            int32_t* channels);
/* Individual host-logic kernels, for parity tests against cv2. */
int  vse_debug_resize_bilinear(vse_engine* e, const uint8_t* src, int32_t sh, int32_t sw, int32_t src_stride,
                               uint8_t* dst_bgrx, int32_t dh, int32_t dw);
int  vse_debug_db_postprocess(vse_engine* e, const float* prob, int32_t rh, int32_t rw, int32_t src_h,
                              int32_t src_w, float* quads, float* scores, int32_t capacity, int32_t* n_out);
int  vse_debug_crop(vse_engine* e, const uint8_t* frame, int32_t h, int32_t w, const float* quad,
                    uint8_t* out_bgr, int32_t capacity, int32_t* out_h, int32_t* out_w);

/* Re-runs the steps of the last run of plan `which` with a CUDA event between steps (on the engine's stream);
 * ms[k] = mean device time of step k; info[k][8] = {op | kernel_kind << 8 (0 generic, 1 tcgen05, 2 specialised, 3 fused into the previous step), in_pixels, out_pixels, cin, cout, taps, in_elt_bytes,
 * out_elt_bytes}.  Returns the number of steps.  bench.py derives the per-kernel roofline from this. */
int  vse_debug_time_steps(vse_engine* e, int32_t which, int32_t reps, float* ms, int64_t* info, int32_t capacity);

#ifdef __cplusplus
}
#endif
#endif /* VSE_B200_H */
