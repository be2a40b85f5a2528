// postproc.cuh — launch interface of the device-side host-logic kernels (postproc.cu):
//   DB post-process  (paddleocr DBPostProcess + TextDetector.filter_tag_det_res + TextSystem.sorted_boxes)
//   text-line crops  (TextSystem.get_rotate_crop_image)
//   CTC greedy decode (CTCLabelDecode)
// See dbpost_core.cuh / geom.cuh for the per-candidate arithmetic and SURVEY.md Appendix D for the upstream algorithm.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace vse {

static constexpr int kSlotCap = 4096;   // contours (components + hole borders) tracked per frame (more => error, never silent)
static constexpr int kHoleFlag = 0x40000000;   // DbWorkspace::roots entry of a hole border: first hole pixel | kHoleFlag

struct DetFrame {
    int map_off;   // first pixel of this frame's probability map in the det output value
    int rh, rw;    // map size (= resized det input)
    int src_h, src_w;
};

struct DbParams {
    float thresh, box_thresh, unclip_ratio;
    int max_candidates;     // contours examined per frame (upstream: first 1000 in cv2 order)
    int max_boxes;          // output rows per frame
    int sort_reading_order; // 1: TextSystem.sorted_boxes order; 0: TextDetector (contour) order
};

struct DbWorkspace {
    int* labels;      // [total map pixels] union-find parents -> flattened roots (-1 = background)
    int* slot_of;     // [total map pixels] component slot, valid at root pixels
    int* n_comp;      // [n_frames]
    int* roots;       // [n_frames][kSlotCap] root pixel (global index) per slot
    int* bbox;        // [n_frames][kSlotCap][4] xmin, ymin, xmax, ymax (map coordinates)
    int* order;       // [n_frames][kSlotCap] slots in cv2 contour order
    float* cand;      // [n_frames][max_candidates][10]: valid, score, quad[8]
    int* status;      // [n_frames] bit 0: too many components, bit 1: too many boxes, bit 2: non-finite probability
    // hole borders (cv2.findContours RETR_LIST also reports them): background pixels of the listed blocks are labelled too
    // (4-connectivity); a background component that never touches the image border or an unlisted block is a hole
    int* blabels;     // [total map pixels] union-find over background pixels (valid inside listed blocks; -1 = foreground)
    int* bopen;       // [total map pixels] at background roots: 1 = connected to the outside (not a hole)
    int* blk_listed;  // [n_frames * blocks_y * blocks_x] 1 = block holds foreground (its blabels / bopen are valid)
    int* ckey;        // [n_frames][kSlotCap] discovery key of a contour: 2 * start pixel (+ 1 for a hole border)
    int* fg_count;    // [1] number of (32 x 8)-pixel blocks that hold foreground
    int* fg_list;     // [n_frames * blocks_y * blocks_x] their codes (frame << 20 | block_y << 10 | block_x)
    // outputs
    int* n_boxes;     // [n_frames]
    float* quads;     // [n_frames][max_boxes][8]
    float* scores;    // [n_frames][max_boxes]
};

size_t db_candidate_smem_bytes(int max_rh);
// prob: det plan output (float32 [pixels], frames back to back as DetFrame::map_off says)
void launch_db_postprocess(const float* prob, const DetFrame* frames_dev, const DetFrame* frames_host, int n_frames,
                           int max_rh, int max_rw, const DbParams& p, const DbWorkspace& ws, cudaStream_t st,
                           int64_t* launches);

void launch_db_compact(int n_frames, const DbParams& p, const DbWorkspace& ws, cudaStream_t st, int64_t* launches);

struct CropJob {
    const uint8_t* frame;   // source frame (device), BGR or BGRX rows
    int fh, fw, stride, pix;
    double M[9];            // crop pixel -> frame coordinates
    int cw, ch;             // size of the warped crop (before the optional rot90)
    int rot90;              // 1: output is np.rot90(crop) with shape (cw, ch)
    long long dst_off;      // BGRX pixel offset of the output crop
};
void launch_crops(const CropJob* jobs_dev, int n_jobs, int max_pix, const short* cubic_tab, uint8_t* dst, cudaStream_t st);

// probs: float32 [sum T][C]; toff/tlen per crop; ids [n][max_t]
// probs: [time steps][cs] rows of C class probabilities — or, logits = true, of class LOGITS (the softmax is folded into the decode)
void launch_ctc_decode(const float* probs, int C, int cs, bool logits, const int* toff, const int* tlen, int n, int max_t, int* ids,
                       int* id_len, float* score, cudaStream_t st);

}  // namespace vse
