// iou.cu -- pairwise IoU family and IoU-based target matching (sm_100a).
//
//  k_box_iou       : [M,N] matrix of helper.bbox_iou (IoU/GIoU/DIoU/CIoU) or torchvision box_iou.
//  k_box_iou_pair  : element-wise version for the loss-side call (yolo_forw.py:125).
//  k_iou_match     : YOLOForw.get_target's IoU + reductions (yolo_forw.py:186-201) without ever
//                    writing the [M,N] matrix: every thread keeps a few anchors in registers,
//                    walks the image's ground-truth boxes (staged in shared memory), tracks
//                    "all IoUs below the ignore threshold" per anchor and feeds a per-GT first
//                    argmax through redux.sync -> shared 64-bit atomicMax -> global atomicMax.
//  k_match_finish  : unpacks the argmax keys and clears the matched anchors in the mask.
// Bound: SM issue rate (about 35 fp32 lane-ops per GIoU pair), not memory.
#include "common.cuh"

namespace b200 {

__device__ __forceinline__ Box load_box(const float* p, int xcycwh) {
    const float4 v = *reinterpret_cast<const float4*>(p);
    if (xcycwh) return abs_coord(v.x, v.y, v.z, v.w);
    return Box{v.x, v.y, v.z, v.w};
}

__global__ void __launch_bounds__(256)
k_box_iou(const float* __restrict__ b1, int M, const float* __restrict__ b2, int N, int kind,
          int xcycwh, float* __restrict__ out) {
    __shared__ Box rows[16];
    const int m0 = blockIdx.y * 16;
    const int n = blockIdx.x * 256 + threadIdx.x;
    if (threadIdx.x < 16 && m0 + threadIdx.x < M) rows[threadIdx.x] = load_box(b1 + 4 * (size_t)(m0 + threadIdx.x), xcycwh);
    __syncthreads();
    if (n >= N) return;
    const Box q = load_box(b2 + 4 * (size_t)n, xcycwh);
    const int mr = min(16, M - m0);
    for (int r = 0; r < mr; ++r) out[(size_t)(m0 + r) * N + n] = pair_iou(rows[r], q, kind);
}

__global__ void __launch_bounds__(256)
k_box_iou_pair(const float* __restrict__ b1, const float* __restrict__ b2, int K, int kind,
               int xcycwh, float* __restrict__ out) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= K) return;
    out[i] = pair_iou(load_box(b1 + 4 * (size_t)i, xcycwh), load_box(b2 + 4 * (size_t)i, xcycwh), kind);
}

// ------------------------------------------------------------------------------------------
// backward of the paired IoU family (loss side, yolo_forw.py:125,143-146)
// ------------------------------------------------------------------------------------------
// d out / d (both boxes) for out = helper.bbox_iou(b1, b2, kind) on K pairs, derived from the reference's own
// expression graph (helper.py:244-277) so that it matches what torch autograd produces for it: min/max split
// the gradient evenly on ties, clamp(min=0) passes it where the argument is >= 0, CIoU's alpha is a constant
// (computed under no_grad, :273-274).  One thread per pair; plain fp32 (gradients are not threshold inputs).
struct Corner8 { float ax1, ay1, ax2, ay2, bx1, by1, bx2, by2; };

__device__ __forceinline__ void d_min(float a, float b, float g, float& ga, float& gb) {   // d min(a,b)
    if (a < b) ga += g; else if (b < a) gb += g; else { ga += 0.5f * g; gb += 0.5f * g; }
}
__device__ __forceinline__ void d_max(float a, float b, float g, float& ga, float& gb) {   // d max(a,b)
    if (a > b) ga += g; else if (b > a) gb += g; else { ga += 0.5f * g; gb += 0.5f * g; }
}

__global__ void __launch_bounds__(256)
k_box_iou_pair_bwd(const float* __restrict__ b1, const float* __restrict__ b2, const float* __restrict__ gout,
                   int K, int kind, int xcycwh, float* __restrict__ g1, float* __restrict__ g2) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= K) return;
    const Box p = load_box(b1 + 4 * (size_t)i, xcycwh), q = load_box(b2 + 4 * (size_t)i, xcycwh);
    const float go = gout[i];
    // forward quantities
    const float mnx2 = fminf(p.x2, q.x2), mxx1 = fmaxf(p.x1, q.x1), mny2 = fminf(p.y2, q.y2), mxy1 = fmaxf(p.y1, q.y1);
    const float dw = mnx2 - mxx1, dh = mny2 - mxy1;
    const float iw = fmaxf(dw, 0.f), ih = fmaxf(dh, 0.f);
    const float inter = iw * ih;
    const float w1 = p.x2 - p.x1, h1 = p.y2 - p.y1, w2 = q.x2 - q.x1, h2 = q.y2 - q.y1;
    const float uni = (w1 * h1 + 1e-16f) + w2 * h2 - inter;
    const float iou = inter / uni;
    // adjoints of the scalar intermediates: out = iou [- penalty]
    float g_inter = go / uni, g_uni = -go * inter / (uni * uni);
    float g_cw = 0.f, g_ch = 0.f;
    float gw1 = 0.f, gh1 = 0.f, gw2 = 0.f, gh2 = 0.f;
    Corner8 g{0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const float mxx2 = fmaxf(p.x2, q.x2), mnx1 = fminf(p.x1, q.x1), mxy2 = fmaxf(p.y2, q.y2), mny1 = fminf(p.y1, q.y1);
    const float cw = mxx2 - mnx1, ch = mxy2 - mny1;
    if (kind == B200_GIOU) {
        // out = iou - (c - uni) / c = iou - 1 + uni / c
        const float c = cw * ch + 1e-16f;
        g_uni += go / c;
        const float g_c = -go * uni / (c * c);
        g_cw += g_c * ch; g_ch += g_c * cw;
    } else if (kind == B200_DIOU || kind == B200_CIOU) {
        const float c2 = cw * cw + ch * ch + 1e-16f;
        const float sx = (q.x1 + q.x2) - (p.x1 + p.x2), sy = (q.y1 + q.y2) - (p.y1 + p.y2);
        const float rho2 = sx * sx * 0.25f + sy * sy * 0.25f;
        // out -= rho2 / c2
        const float g_rho = -go / c2, g_c2 = go * rho2 / (c2 * c2);
        g_cw += g_c2 * 2.f * cw; g_ch += g_c2 * 2.f * ch;
        const float gsx = g_rho * 0.5f * sx, gsy = g_rho * 0.5f * sy;
        g.bx1 += gsx; g.bx2 += gsx; g.ax1 -= gsx; g.ax2 -= gsx;
        g.by1 += gsy; g.by2 += gsy; g.ay1 -= gsy; g.ay2 -= gsy;
        if (kind == B200_CIOU) {
            // out -= v * alpha, v = 4/pi^2 (atan(w2/h2) - atan(w1/h1))^2, alpha = v / (1 - iou + v) held constant
            const float kv = 0.40528473456935116f;
            const float da = atanf(w2 / h2) - atanf(w1 / h1);
            const float v = kv * da * da;
            const float alpha = v / (1.f - iou + v);
            const float g_da = -go * alpha * kv * 2.f * da;
            // d atan(w/h) = (h dw - w dh) / (w^2 + h^2)
            const float n2 = w2 * w2 + h2 * h2, n1 = w1 * w1 + h1 * h1;
            gw2 += g_da * h2 / n2; gh2 -= g_da * w2 / n2;
            gw1 -= g_da * h1 / n1; gh1 += g_da * w1 / n1;
        }
    }
    // enclosing box extents
    d_max(p.x2, q.x2, g_cw, g.ax2, g.bx2);  d_min(p.x1, q.x1, -g_cw, g.ax1, g.bx1);
    d_max(p.y2, q.y2, g_ch, g.ay2, g.by2);  d_min(p.y1, q.y1, -g_ch, g.ay1, g.by1);
    // union = w1*h1 + 1e-16 + w2*h2 - inter
    gw1 += g_uni * h1; gh1 += g_uni * w1; gw2 += g_uni * h2; gh2 += g_uni * w2;
    g_inter -= g_uni;
    // inter = clamp(dw, 0) * clamp(dh, 0)
    const float g_dw = dw >= 0.f ? g_inter * ih : 0.f, g_dh = dh >= 0.f ? g_inter * iw : 0.f;
    d_min(p.x2, q.x2, g_dw, g.ax2, g.bx2);  d_max(p.x1, q.x1, -g_dw, g.ax1, g.bx1);
    d_min(p.y2, q.y2, g_dh, g.ay2, g.by2);  d_max(p.y1, q.y1, -g_dh, g.ay1, g.by1);
    // widths / heights
    g.ax2 += gw1; g.ax1 -= gw1; g.ay2 += gh1; g.ay1 -= gh1;
    g.bx2 += gw2; g.bx1 -= gw2; g.by2 += gh2; g.by1 -= gh2;
    // back to the input format: x1 = xc - w/2, x2 = xc + w/2 (helper.py:203-217)
    float4 o1, o2;
    if (xcycwh) {
        o1 = make_float4(g.ax1 + g.ax2, g.ay1 + g.ay2, 0.5f * (g.ax2 - g.ax1), 0.5f * (g.ay2 - g.ay1));
        o2 = make_float4(g.bx1 + g.bx2, g.by1 + g.by2, 0.5f * (g.bx2 - g.bx1), 0.5f * (g.by2 - g.by1));
    } else {
        o1 = make_float4(g.ax1, g.ay1, g.ax2, g.ay2);
        o2 = make_float4(g.bx1, g.by1, g.bx2, g.by2);
    }
    if (g1) reinterpret_cast<float4*>(g1)[i] = o1;
    if (g2) reinterpret_cast<float4*>(g2)[i] = o2;
}

int launch_box_iou_pair_bwd(const float* b1, const float* b2, const float* gout, int K, int kind, int xcycwh,
                            float* g1, float* g2, cudaStream_t st) {
    if (K <= 0) return B200_OK;
    k_box_iou_pair_bwd<<<cdiv(K, 256), 256, 0, st>>>(b1, b2, gout, K, kind, xcycwh, g1, g2);
    return cudaGetLastError() == cudaSuccess ? B200_OK : B200_ERR_CUDA;
}

static constexpr int kMatchThreads = 256;
static constexpr int kMatchItems = 4;    // anchors per thread
static constexpr int kGtChunk = 128;     // ground-truth boxes staged per pass

// IoU / GIoU of helper.bbox_iou with the per-box terms hoisted out of the pair loop: a1e = w1*h1 + 1e-16 of the
// ground-truth box and a2 = w2*h2 of the anchor are rounded exactly as in pair_iou (common.cuh), so the result is
// bit-identical; KIND is a compile-time constant (0 IoU, 1 GIoU), other kinds go through pair_iou.
template <int KIND>
__device__ __forceinline__ float match_iou(const Box& p, float a1e, const Box& q, float a2, int kind) {
    if (KIND < 0) return pair_iou(p, q, kind);
    const float iw = fmaxf(__fsub_rn(fminf(p.x2, q.x2), fmaxf(p.x1, q.x1)), 0.0f);
    const float ih = fmaxf(__fsub_rn(fminf(p.y2, q.y2), fmaxf(p.y1, q.y1)), 0.0f);
    const float inter = __fmul_rn(iw, ih);
    const float uni = __fsub_rn(__fadd_rn(a1e, a2), inter);               // helper.py:255
    const float iou = __fdiv_rn(inter, uni);
    if (KIND == 0) return iou;
    const float cw = __fsub_rn(fmaxf(p.x2, q.x2), fminf(p.x1, q.x1));
    const float ch = __fsub_rn(fmaxf(p.y2, q.y2), fminf(p.y1, q.y1));
    const float c_area = __fadd_rn(__fmul_rn(cw, ch), 1e-16f);            // :259-263
    return __fsub_rn(iou, __fdiv_rn(__fsub_rn(c_area, uni), c_area));
}

// SKIP (IoU / GIoU with a positive ignore threshold): a pair whose boxes do not intersect has IoU == +0 and
// GIoU <= 0, so it can neither clear the no-object flag (v < thr) nor beat a POSITIVE best value -- it is dropped
// after the 6-instruction intersection test and only intersecting pairs (a few percent, coherent across a warp
// because consecutive anchors are neighbouring cells) pay for the IEEE divisions.  A ground-truth box whose best
// intersecting value is not positive (or that intersects no anchor) is redone exhaustively by k_iou_match_redo, so
// the result is bit-identical to the exhaustive evaluation.
template <int KIND, bool SKIP>
__global__ void __launch_bounds__(kMatchThreads)
k_iou_match(const float* __restrict__ gt, const int* __restrict__ gt_count, int max_gt,
            const float* __restrict__ anchors, int N, int kind, float ignore_thr,
            unsigned long long* __restrict__ best_key, uint8_t* __restrict__ noobj) {
    __shared__ Box sgt[kGtChunk];
    __shared__ float sarea[kGtChunk];
    __shared__ unsigned long long sbest[kGtChunk];
    const int b = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31;
    const int M = min(gt_count[b], max_gt);
    const int n0 = blockIdx.x * (kMatchThreads * kMatchItems);

    Box anc[kMatchItems];
    float a2[kMatchItems];
    bool valid[kMatchItems], free_[kMatchItems];
#pragma unroll
    for (int k = 0; k < kMatchItems; ++k) {
        const int n = n0 + k * kMatchThreads + tid;
        valid[k] = n < N;
        free_[k] = true;
        anc[k] = valid[k] ? load_box(anchors + 4 * (size_t)n, 1) : Box{0.f, 0.f, 0.f, 0.f};
        a2[k] = __fmul_rn(__fsub_rn(anc[k].x2, anc[k].x1), __fsub_rn(anc[k].y2, anc[k].y1));
    }

    // SKIP: the extent of this warp's anchors (consecutive anchors are neighbouring cells: a strip of the image).  A
    // ground-truth box that misses the extent intersects none of the 128 anchors: one warp-uniform test replaces 128
    // pair tests.  A NaN coordinate in the warp switches the shortcut off (NaN pairs must reach the exact path).
    float ex1 = INFINITY, ey1 = INFINITY, ex2 = -INFINITY, ey2 = -INFINITY;
    bool warp_extent = false;
    if (SKIP) {
        bool nan = false;
#pragma unroll
        for (int k = 0; k < kMatchItems; ++k)
            if (valid[k]) {
                ex1 = fminf(ex1, anc[k].x1); ey1 = fminf(ey1, anc[k].y1);
                ex2 = fmaxf(ex2, anc[k].x2); ey2 = fmaxf(ey2, anc[k].y2);
                nan |= !(anc[k].x1 == anc[k].x1) || !(anc[k].y1 == anc[k].y1) || !(anc[k].x2 == anc[k].x2) || !(anc[k].y2 == anc[k].y2);
            }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            ex1 = fminf(ex1, __shfl_xor_sync(kFullMask, ex1, o)); ey1 = fminf(ey1, __shfl_xor_sync(kFullMask, ey1, o));
            ex2 = fmaxf(ex2, __shfl_xor_sync(kFullMask, ex2, o)); ey2 = fmaxf(ey2, __shfl_xor_sync(kFullMask, ey2, o));
        }
        warp_extent = !__any_sync(kFullMask, nan);
    }

    for (int mc = 0; mc < M; mc += kGtChunk) {
        const int mm = min(kGtChunk, M - mc);
        __syncthreads();
        for (int i = tid; i < mm; i += kMatchThreads) {
            const Box g = load_box(gt + ((size_t)b * max_gt + mc + i) * 4, 1);
            sgt[i] = g;
            sarea[i] = __fadd_rn(__fmul_rn(__fsub_rn(g.x2, g.x1), __fsub_rn(g.y2, g.y1)), 1e-16f);
            sbest[i] = 0ull;
        }
        __syncthreads();
        for (int i = 0; i < mm; ++i) {
            const Box g = sgt[i];
            // min(g.x2, a.x2) <= g.x2 <= a.x1 <= max(g.x1, a.x1) for every anchor of the warp: w <= 0 everywhere (same for
            // the other three sides); comparisons with a NaN box are false, so it is not skipped
            if (SKIP && warp_extent && (g.x2 <= ex1 || g.x1 >= ex2 || g.y2 <= ey1 || g.y1 >= ey2)) continue;
            const float a1e = sarea[i];
            unsigned bk = 0u;          // orderable(best iou) over this thread's anchors
            int bkk = -1;              // which of its anchors
#pragma unroll
            for (int k = 0; k < kMatchItems; ++k) {
                if (valid[k]) {
                    if (SKIP) {
                        const float w = __fsub_rn(fminf(g.x2, anc[k].x2), fmaxf(g.x1, anc[k].x1));
                        const float h = __fsub_rn(fminf(g.y2, anc[k].y2), fmaxf(g.y1, anc[k].y1));
                        if (w <= 0.f || h <= 0.f) continue;               // inter == 0 (NaN coordinates fall through)
                    }
                    const float v = match_iou<KIND>(g, a1e, anc[k], a2[k], kind);
                    free_[k] = free_[k] && (v < ignore_thr);
                    const unsigned ok = orderable(v);
                    if (ok > bk) { bk = ok; bkk = k; }                    // ascending n: first max
                }
            }
            if (SKIP && !__any_sync(kFullMask, bkk >= 0)) continue;       // no lane of this warp met the box
            const unsigned bn = bkk >= 0 ? (unsigned)(n0 + bkk * kMatchThreads + tid) : 0xffffffffu;
            const unsigned wmax = __reduce_max_sync(kFullMask, bk);
            const unsigned wn = __reduce_min_sync(kFullMask, bk == wmax ? bn : 0xffffffffu);
            if (lane == 0 && wn != 0xffffffffu)
                atomicMax(&sbest[i], ((unsigned long long)wmax << 32) | (unsigned long long)(~wn));
        }
        __syncthreads();
        for (int i = tid; i < mm; i += kMatchThreads)
            if (sbest[i]) atomicMax(best_key + (size_t)b * max_gt + mc + i, sbest[i]);
    }
#pragma unroll
    for (int k = 0; k < kMatchItems; ++k) {
        const int n = n0 + k * kMatchThreads + tid;
        if (valid[k]) noobj[(size_t)b * N + n] = free_[k] ? 1 : 0;
    }
}

// Exhaustive redo of the ground-truth boxes the skipping kernel could not decide (best intersecting value <= +0, or
// no intersecting anchor): one CTA per image walks its flagged boxes; normally there are none.
__global__ void __launch_bounds__(256)
k_iou_match_redo(const float* __restrict__ gt, const int* __restrict__ gt_count, int max_gt,
                 const float* __restrict__ anchors, int N, int kind, unsigned long long* __restrict__ best_key) {
    __shared__ unsigned long long s_best;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
    const int M = min(gt_count[b], max_gt);
    for (int m = 0; m < M; ++m) {
        const unsigned long long cur = best_key[(size_t)b * max_gt + m];
        if ((unsigned)(cur >> 32) > 0x80000000u) continue;              // a positive best value: decided
        if (tid == 0) s_best = 0ull;
        __syncthreads();
        const Box g = load_box(gt + ((size_t)b * max_gt + m) * 4, 1);
        unsigned bk = 0u, bn = 0xffffffffu;
        for (int n = tid; n < N; n += 256) {
            const unsigned ok = orderable(pair_iou(g, load_box(anchors + 4 * (size_t)n, 1), kind));
            if (ok > bk) { bk = ok; bn = (unsigned)n; }
        }
        const unsigned wmax = __reduce_max_sync(kFullMask, bk);
        const unsigned wn = __reduce_min_sync(kFullMask, bk == wmax ? bn : 0xffffffffu);
        if (lane == 0 && wn != 0xffffffffu) atomicMax(&s_best, ((unsigned long long)wmax << 32) | (unsigned long long)(~wn));
        __syncthreads();
        if (tid == 0) best_key[(size_t)b * max_gt + m] = s_best;
        __syncthreads();
    }
}

__global__ void k_match_finish(const unsigned long long* __restrict__ best_key,
                               const int* __restrict__ gt_count, int max_gt, int N,
                               long long* __restrict__ best_anchor, uint8_t* __restrict__ noobj) {
    const int b = blockIdx.x;
    const int M = min(gt_count[b], max_gt);
    for (int m = threadIdx.x; m < max_gt; m += blockDim.x) {
        long long idx = 0;
        if (m < M) {
            idx = (long long)(~(unsigned)(best_key[(size_t)b * max_gt + m] & 0xffffffffull));
            if (idx >= 0 && idx < N) noobj[(size_t)b * N + idx] = 0;   // yolo_forw.py:201
        }
        best_anchor[(size_t)b * max_gt + m] = idx;
    }
}

// ------------------------------------------------------------------------------------------
int launch_box_iou(const float* b1, int m, const float* b2, int n, int kind, int xcycwh, float* out,
                   cudaStream_t stream) {
    if (m <= 0 || n <= 0) return B200_OK;
    dim3 grid(cdiv(n, 256), cdiv(m, 16));
    k_box_iou<<<grid, 256, 0, stream>>>(b1, m, b2, n, kind, xcycwh, out);
    return cudaGetLastError() == cudaSuccess ? B200_OK : B200_ERR_CUDA;
}

int launch_box_iou_pair(const float* b1, const float* b2, int k, int kind, int xcycwh, float* out,
                        cudaStream_t stream) {
    if (k <= 0) return B200_OK;
    k_box_iou_pair<<<cdiv(k, 256), 256, 0, stream>>>(b1, b2, k, kind, xcycwh, out);
    return cudaGetLastError() == cudaSuccess ? B200_OK : B200_ERR_CUDA;
}

int launch_iou_match(const float* gt, const int* gt_count, int batch, int max_gt, const float* anchors,
                     int n, int kind, float ignore_thr, long long* best_anchor, uint8_t* noobj,
                     unsigned long long* best_key, cudaStream_t stream) {
    if (cudaMemsetAsync(best_key, 0, sizeof(unsigned long long) * (size_t)batch * max_gt, stream) != cudaSuccess)
        return B200_ERR_CUDA;
    dim3 grid(cdiv(n, kMatchThreads * kMatchItems), batch);
    const bool skip = ignore_thr > 0.f && (kind == B200_IOU || kind == B200_GIOU);
    if (kind == B200_IOU && skip)
        k_iou_match<0, true><<<grid, kMatchThreads, 0, stream>>>(gt, gt_count, max_gt, anchors, n, kind, ignore_thr, best_key, noobj);
    else if (kind == B200_GIOU && skip)
        k_iou_match<1, true><<<grid, kMatchThreads, 0, stream>>>(gt, gt_count, max_gt, anchors, n, kind, ignore_thr, best_key, noobj);
    else if (kind == B200_IOU)
        k_iou_match<0, false><<<grid, kMatchThreads, 0, stream>>>(gt, gt_count, max_gt, anchors, n, kind, ignore_thr, best_key, noobj);
    else if (kind == B200_GIOU)
        k_iou_match<1, false><<<grid, kMatchThreads, 0, stream>>>(gt, gt_count, max_gt, anchors, n, kind, ignore_thr, best_key, noobj);
    else
        k_iou_match<-1, false><<<grid, kMatchThreads, 0, stream>>>(gt, gt_count, max_gt, anchors, n, kind, ignore_thr, best_key, noobj);
    if (skip) k_iou_match_redo<<<batch, 256, 0, stream>>>(gt, gt_count, max_gt, anchors, n, kind, best_key);
    k_match_finish<<<batch, 128, 0, stream>>>(best_key, gt_count, max_gt, n, best_anchor, noobj);
    return cudaGetLastError() == cudaSuccess ? B200_OK : B200_ERR_CUDA;
}

}  // namespace b200
