// nms.cuh -- parameter block of the segmented NMS kernels (see nms.cu).
#pragma once
#include "common.cuh"

namespace b200 {

struct Cand;

struct NmsParams {
    // ---- source A: caller arrays, segments given by offsets (b200_nms) --------------------
    const float* boxes;      // [T,4] xyxy, 16 B aligned
    const float* scores;     // [T]
    const int* labels;       // [T] or nullptr
    const int* seg_offsets;  // [S+1] (or [S] starts when seg_counts is given)
    const int* seg_counts;   // [S] or nullptr: explicit segment lengths (fixed-stride layouts)
    long long* keep;         // [T]
    int* keep_count;         // [S]
    int* labels_out;         // [T] or nullptr
    // ---- source B: unordered candidate slab of the fused decode kernel --------------------
    const Cand* slab;        // [S, cap]
    const int* count;        // [S] true candidate counts
    int cap;
    float4* cbox;            // [S, cap] canonical (ascending anchor) candidate arrays, written
    float* cscore;           //          only by the canonicalise-only mode (mode < 0)
    int* clabel;
    int* canchor;
    float* det;              // [S, max_det, 6]
    int* det_keep;           // [S, max_det] or nullptr
    int* det_anchor;         // [S, max_det] or nullptr
    int* det_count;          // [S]
    int* cand_count_out;     // [S] or nullptr
    int max_det;
    int* status;
    // ---- configuration -----------------------------------------------------------------------
    int mode;                // B200_NMS_* ; < 0 = canonicalise the slab only
    float thr_f;             // MAJORITY: threshold rounded to fp32 (tensor-vs-scalar compare)
    double thr_d;            // TV modes: compared against (double)iou
    int from_slab;
    const float* given_unit;  // [S] or nullptr: coordinate-trick units supplied by the caller (segments that are
                              // slices of one torchvision call share the unit of the whole call)
    long long auto_limit;    // TV_AUTO: segments with 4*n > auto_limit use the per-class arithmetic
    int anchor_space;        // slab path: flat anchor indices are < anchor_space (0 = unknown)
    int num_segments;
    int max_seg;             // upper bound of any segment length (host-known)
    int max_words;           // cdiv(max_seg, 64): row stride of the dominator bitmask
    int max_tiles;           // nt*(nt+1)/2 with nt = max_words: work items per segment at most
    // ---- global scratch ---------------------------------------------------------------------
    unsigned long long* gkey;   // [2*T] sort keys when a sort does not fit in shared memory
    int* gval;                  // [2*T] sort payload
    int* gperm;                 // [T] binned position -> index inside the segment
    unsigned long long* dom;    // [S * max_seg * max_words]: bit q of row p = "the box at position q
                                //  precedes the box at position p in (score desc, index asc) order
                                //  and suppresses it"
    unsigned long long* nzmask; // [T] per row (segments of <= 4096 boxes): which words of the row are non-zero
    float* shift_unit;          // [S] coordinate-trick offset unit (max coordinate + 1)
    int2* work;                 // [S * max_tiles] (segment, row_tile << 16 | col_tile)
    int* work_count;            // [2] number of work items, queue cursor
    int* gsup;                  // [T] first suppressor | vote flag (MAJORITY, slow path)
    int* gklist;                // [T] kept positions (slow path)
    int* gnewlab;               // [T] label of each kept box after the majority vote (slow path)
    long long* prof;            // debug: [S, 8] clock64 stamps of the resolve phases (nullptr = off)
    // ---- single-launch path (nms_fused.cu, segments of <= 4096 boxes) ---------------------------
    int* f_ctl;                 // [S] team arrival counters (zeroed before every launch)
    // private scratch of every CTA (slot = blockIdx.x * max_seg): its current segment in RANK order
    float4* f_rbox;             // boxes by rank (score desc, index asc), shifted in coordinate-trick modes
    float* f_rarea;
    int* f_rlabel;
    unsigned long long* f_rkey; // (~orderable(score) << 32) | tie << 12 | index inside the segment
    int force_general;          // debug: 1 = always take the three-launch path
    int serial;                 // the call is not overlapped with the decode of another batch (b200_yolo_postprocess):
                                // the large segments of a slab go to the 1024-thread single-launch kernel, not the 256-thread one
    int split_lo, split_hi;     // this launch only handles segments with split_lo < n <= split_hi (others are left alone)
};

// scratch bytes needed for T boxes in S segments of at most max_seg boxes each
size_t nms_scratch_bytes(size_t total, size_t segments, size_t max_seg);
// carve the scratch block (256 B aligned) into the pointers above; returns false if too small
bool nms_carve_scratch(NmsParams* P, size_t total, size_t segments, size_t max_seg, void* base, size_t bytes);

int launch_nms(NmsParams& P, int num_segments, cudaStream_t stream);
// nms_fused.cu
bool nms_fused_eligible(const NmsParams& P);
int launch_nms_fused(NmsParams& P, int num_segments, cudaStream_t stream);
int nms_fused_slots();

}  // namespace b200
