// nms.cuh -- parameter block of the per-segment NMS kernel (see nms.cu).
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
    float4* cbox;            // [S, cap] canonical (ascending anchor) candidate arrays
    float* cscore;
    int* clabel;
    int* canchor;
    float* det;              // [S, max_det, 6]
    int* det_keep;           // [S, max_det]
    int* det_anchor;         // [S, max_det] or nullptr
    int* det_count;          // [S]
    int* cand_count_out;     // [S] or nullptr
    int max_det;
    int* status;
    // ---- configuration -----------------------------------------------------------------------
    int mode;                // B200_NMS_* ; < 0 = canonicalise the slab only
    float thr_f;             // MAJORITY: threshold rounded to fp32 (tensor-vs-scalar compare)
    double thr_d;            // TV modes: compared against (double)iou
    int fast_reject;         // inter == 0 can never suppress (thr > 0 resp. >= 0)
    int smem_cap;            // power of two; larger segments use the global scratch below
    // ---- global scratch for oversized segments ------------------------------------------------
    unsigned long long* gkey;  // [2*T]
    float4* gbox;              // [T]
    float* garea;
    int* glabel;
    int* gsup;
    int* gcidx;
};

static constexpr int kNmsSmemCap = 4096;

size_t nms_smem_bytes(int smem_cap);
int launch_nms(const NmsParams& P, int num_segments, bool from_slab, cudaStream_t stream);

}  // namespace b200
