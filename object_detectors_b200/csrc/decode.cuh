// decode.cuh -- device-side descriptors shared by the YOLO decode kernels and the C ABI.
#pragma once
#include "common.cuh"

namespace b200 {

// one candidate detection as emitted by the fused decode+filter kernel (32 B, two 128-bit stores)
struct __align__(16) Cand {
    float x1, y1, x2, y2;
    float score;
    int32_t label;
    int32_t anchor;  // flat anchor index n (scale offset + (h*W+w)*A + a)
    int32_t pad;
};
static_assert(sizeof(Cand) == 32, "Cand must be 32 bytes");

struct ScaleDev {
    const float* head;  // [B, A*(5+C), grid, grid]
    int grid;           // W == H
    int hw;             // grid*grid
    int vec;            // cells per lane: 4 when every plane row is 16 B aligned, else 1
    int tiles;          // warp tiles per (b, a) = cdiv(hw, 128): 32 lanes x 4 cells
    int task_begin;     // first warp task of this scale
    int anchor_off;     // flat anchor index of (h=0,w=0,a=0) of this scale
    float inw;          // (float)grid                       (yolo_forw.py:116)
    float stride;       // fp32(img_size / inw)              (yolo_forw.py:164)
    float anc[B200_MAX_ANCHORS][2];  // cxypwh[:,2:4] per anchor (yolo_forw.py:108-113)
};

struct DecodeParams {
    ScaleDev sc[B200_MAX_SCALES];
    int num_scales, A, C, B, N;
    int total_tasks;
    const float* idf;  // [C] or nullptr
    float thr;
    Cand* slab;    // [B, cap]
    int cap;
    int* count;    // [B] true candidate count
    int* status;   // |= 1 on slab overflow
};

// host: fills everything except the output pointers; returns B200_OK or B200_ERR_INVALID
int make_decode_params(const b200_yolo_layout* L, const float* const* heads, const float* idf,
                       DecodeParams* p);

}  // namespace b200
