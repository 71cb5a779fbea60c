// nms_dev.cuh -- device helpers shared by the NMS kernels (nms.cu: general three-launch path,
// nms_fused.cu: single-launch path for segments of <= 4096 boxes).
#pragma once
#include "decode.cuh"
#include "nms.cuh"

namespace b200 {

__device__ __forceinline__ int next_pow2(int n) {
    int p = 1;
    while (p < n) p <<= 1;
    return p;
}

__device__ __forceinline__ void segment_range(const NmsParams& P, int seg, long long& off, int& n, int& n_true) {
    if (P.from_slab) {
        n_true = P.count[seg];
        n = min(n_true, P.cap);
        off = (long long)seg * P.cap;
    } else {
        off = P.seg_offsets[seg];
        n = P.seg_counts ? P.seg_counts[seg] : P.seg_offsets[seg + 1] - (int)off;
        n_true = n;
    }
    if (n > P.max_seg) n = P.max_seg;   // host bound; never exceeded when the caller is honest
    if (n < 0) n = 0;
}

// ------------------------------------------------------------------------------------------
// one box as the NMS kernels see it
// ------------------------------------------------------------------------------------------
struct Item {
    float4 b;                // xyxy (shifted by label*unit in coordinate-trick mode)
    float area;              // (x2-x1)*(y2-y1) of b
    unsigned long long key;  // (~orderable(score) << 32) | tie : smaller key = earlier in NMS order
    int label;
};

// coordinate-trick arithmetic: always (TV_TRICK) or per segment (TV_AUTO, unit == 0 where the segment is
// large enough for torchvision to switch to its per-class loop)
__device__ __forceinline__ bool uses_shift(const NmsParams& P) {
    return P.mode == B200_NMS_TV_TRICK || P.mode == B200_NMS_TV_AUTO;
}

// by index i inside the segment
template <bool SLAB>
__device__ __forceinline__ Item load_raw(const NmsParams& P, long long off, int i, float unit) {
    Item it;
    float score;
    unsigned tie;
    if (SLAB) {
        const float4* rec = reinterpret_cast<const float4*>(P.slab + off + i);
        it.b = rec[0];
        const float4 m = rec[1];
        score = m.x;
        it.label = __float_as_int(m.y);
        tie = (unsigned)__float_as_int(m.z);        // flat anchor index: the canonical order
    } else {
        it.b = reinterpret_cast<const float4*>(P.boxes)[off + i];
        score = P.scores[off + i];
        it.label = P.labels ? P.labels[off + i] : 0;
        tie = (unsigned)i;
    }
    if (uses_shift(P)) {
        // boxes + idxs.to(boxes) * (boxes.max() + 1)   (torchvision boxes.py coordinate trick)
        const float sh = __fmul_rn((float)it.label, unit);
        it.b = make_float4(__fadd_rn(it.b.x, sh), __fadd_rn(it.b.y, sh), __fadd_rn(it.b.z, sh), __fadd_rn(it.b.w, sh));
    }
    it.area = __fmul_rn(__fsub_rn(it.b.z, it.b.x), __fsub_rn(it.b.w, it.b.y));
    it.key = ((unsigned long long)(~orderable(score)) << 32) | tie;
    return it;
}
// by binned position p
template <bool SLAB>
__device__ __forceinline__ Item load_item(const NmsParams& P, long long off, int p, float unit) {
    return load_raw<SLAB>(P, off, P.gperm[off + p], unit);
}

// exact test, S = the box that precedes (picked), T = the later one (remaining)
template <int MODE>
__device__ __forceinline__ bool suppresses_exact(const NmsParams& P, const float4& bs, float as, const float4& bt,
                                                 float at, bool* vote) {
    float w = __fsub_rn(fminf(bs.z, bt.z), fmaxf(bs.x, bt.x));
    float h = __fsub_rn(fminf(bs.w, bt.w), fmaxf(bs.y, bt.y));
    w = fmaxf(w, 0.f);
    h = fmaxf(h, 0.f);
    const float inter = __fmul_rn(w, h);
    // helper.py:361-366  union = (area_T - inter) + area_S ;  torchvision: (area_i + area_j) - inter
    const float den = MODE == B200_NMS_MAJORITY ? __fadd_rn(__fsub_rn(at, inter), as)
                                                : __fsub_rn(__fadd_rn(as, at), inter);
    const float iou = __fdiv_rn(inter, den);
    if (MODE == B200_NMS_MAJORITY) {
        if (vote) *vote = iou > P.thr_f;                 // helper.py:369
        return !(iou < P.thr_f);                         // helper.py:368 (NaN and == thr are removed)
    }
    return (double)iou > P.thr_d;                        // torchvision: double threshold
}

// Unordered pair test.  The intersection is symmetric; only the MAJORITY union
// (area_T - inter) + area_S depends on which box comes first.
template <int MODE>
__device__ __forceinline__ bool pair_hit(const NmsParams& P, const float4& bi, float ai, int li,
                                         const float4& bj, float aj, int lj, bool i_first) {
    if (MODE == B200_NMS_TV_CLASS && li != lj) return false;
    float w = __fsub_rn(fminf(bi.z, bj.z), fmaxf(bi.x, bj.x));
    float h = __fsub_rn(fminf(bi.w, bj.w), fmaxf(bi.y, bj.y));
    w = fmaxf(w, 0.f);
    h = fmaxf(h, 0.f);
    const float inter = __fmul_rn(w, h);
    float den;
    if (MODE == B200_NMS_MAJORITY) {
        const float as = i_first ? ai : aj, at = i_first ? aj : ai;
        den = __fadd_rn(__fsub_rn(at, inter), as);
    } else {
        den = __fsub_rn(__fadd_rn(ai, aj), inter);
    }
    // fl(inter/den) differs from inter/den by < 2^-24 relative: outside a 1e-6 band around thr*den
    // the threshold comparison is decided without dividing.
    const float t = __fmul_rn(den, P.thr_f);
    if (den > 1e-30f) {
        if (inter > __fmul_rn(t, 1.000001f)) return true;
        if (inter < __fmul_rn(t, 0.999999f)) return false;
    }
    const float iou = __fdiv_rn(inter, den);
    if (MODE == B200_NMS_MAJORITY) return !(iou < P.thr_f);
    return (double)iou > P.thr_d;
}

__device__ __forceinline__ float block_reduce_max(float v, float* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFullMask, v, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = red[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) r = fmaxf(r, red[w]);
    return r;
}

__device__ __forceinline__ int box_bin(const Item& it, float ymin, float yscale) {
    const bool ok = it.area > 0.f && it.area < 3.0e38f;
    if (!ok) return 0;
    const int e = ((__float_as_int(it.area) >> 23) & 0xff) - 127;       // floor(log2(area))
    // one octave of area per class: IoU >= thr needs an area ratio >= thr, so boxes two classes apart cannot
    // suppress each other and tiles of different classes are pruned by their area ranges (19 % fewer tile
    // pairs survive than with two-octave classes on the C2 workload)
    const int cls = min(max(e - 6, 0), 15);
    const float cy = 0.5f * (it.b.y + it.b.w);
    int yb = (int)((cy - ymin) * yscale);
    yb = min(max(yb, 0), 15);
    return cls * 16 + yb;
}

}  // namespace b200
