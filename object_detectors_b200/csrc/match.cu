// match.cu -- box_iou + Matcher / SSDMatcher fused: the [M, N] quality matrix is never written (sm_100a).
//
// Replaces, for the target assignment of the torchvision path (rpn.py:179-213, roi_heads.py:627-651,
// retinanet.py:409-410, ssd.py:371-372):
//     match_quality_matrix = box_ops.box_iou(gt_boxes, anchors)        # [M, 268 569] fp32
//     matched_idxs = self.proposal_matcher(match_quality_matrix)       # Matcher.__call__, _utils.py:271-344
// with two passes over the (ground truth, box) PAIRS that keep everything in registers / shared memory:
//   k_match_cols  per box n: max / first argmax over the ground-truth boxes (matched_vals, matches = q.max(dim=0)),
//                 the BELOW_LOW / BETWEEN thresholds, and -- warp-reduced, then one atomicMax per warp and ground
//                 truth -- the row maxima q.max(dim=1) with their first index;
//   k_match_rows  (allow_low_quality_matches) per box n: restored iff it ties some ground truth's row maximum
//                 (set_low_quality_matches_, :315-344);   k_match_ssd: SSDMatcher's per-ground-truth override (:347-361).
// IoU arithmetic is torchvision's, operation by operation (inter / ((area1 + area2) - inter), appendix A.3-3), so
// the results equal Matcher()(box_iou(gt, boxes)) bit for bit.  Pairs that do not intersect have quality +0 exactly
// and are skipped after the 6-instruction intersection test (unless both areas vanish: 0/0 = NaN is kept).
#include "common.cuh"

namespace b200 {

static constexpr int kMT = 256;
static constexpr int kMGt = 256;       // ground-truth boxes staged per pass

__device__ __forceinline__ float tv_iou(const float4& g, float ag, const float4& b, float ab, bool& hit) {
    const float w = __fsub_rn(fminf(g.z, b.z), fmaxf(g.x, b.x));
    const float h = __fsub_rn(fminf(g.w, b.w), fmaxf(g.y, b.y));
    hit = true;
    if ((w <= 0.f || h <= 0.f) && __fadd_rn(ag, ab) > 0.f) { hit = false; return 0.f; }      // inter == 0, union > 0
    const float inter = __fmul_rn(fmaxf(w, 0.f), fmaxf(h, 0.f));
    return __fdiv_rn(inter, __fsub_rn(__fadd_rn(ag, ab), inter));
}

__global__ void __launch_bounds__(kMT)
k_match_cols(const float4* __restrict__ gt, int M, const float4* __restrict__ boxes, int N, float high, float low,
             long long* __restrict__ matches, long long* __restrict__ all_matches, float* __restrict__ matched_vals,
             unsigned long long* __restrict__ row_key) {
    __shared__ float4 sg[kMGt];
    __shared__ float sa[kMGt];
    __shared__ unsigned long long sbest[kMGt];
    const int tid = threadIdx.x, lane = tid & 31;
    const int n = blockIdx.x * kMT + tid;
    const bool valid = n < N;
    const float4 b = valid ? boxes[n] : make_float4(0.f, 0.f, 0.f, 0.f);
    const float ab = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
    float best = 0.f;
    int arg = 0;
    bool first = true;
    for (int m0 = 0; m0 < M; m0 += kMGt) {
        const int mm = min(kMGt, M - m0);
        __syncthreads();
        for (int i = tid; i < mm; i += kMT) {
            const float4 g = gt[m0 + i];
            sg[i] = g;
            sa[i] = __fmul_rn(__fsub_rn(g.z, g.x), __fsub_rn(g.w, g.y));
            sbest[i] = 0ull;
        }
        __syncthreads();
        for (int i = 0; i < mm; ++i) {
            bool hit = false;
            float v = 0.f;
            if (valid) v = tv_iou(sg[i], sa[i], b, ab, hit);
            if (valid) {
                // quality.max(dim=0): first maximum, NaN propagates (torch.max)
                if (first) { best = v; arg = m0 + i; first = false; }
                else if (v > best || (v != v && best == best)) { best = v; arg = m0 + i; }
            }
            // row maximum with its first index: only intersecting pairs can exceed +0
            if (!__any_sync(kFullMask, valid && hit)) continue;
            const unsigned key = valid && hit ? orderable(v) : 0u;
            const unsigned wmax = __reduce_max_sync(kFullMask, key);
            const unsigned wn = __reduce_min_sync(kFullMask, key == wmax && valid && hit ? (unsigned)n : 0xffffffffu);
            if (lane == 0 && wn != 0xffffffffu) atomicMax(&sbest[i], ((unsigned long long)wmax << 32) | (unsigned long long)(~wn));
        }
        __syncthreads();
        for (int i = tid; i < mm; i += kMT)
            if (sbest[i]) atomicMax(row_key + m0 + i, sbest[i]);
    }
    if (!valid) return;
    if (all_matches) all_matches[n] = arg;
    if (matched_vals) matched_vals[n] = best;
    long long r = arg;
    if (best < low) r = -1;                                   // BELOW_LOW_THRESHOLD   (:300-306)
    else if (best >= low && best < high) r = -2;              // BETWEEN_THRESHOLDS
    matches[n] = r;
}

// row maximum of ground truth m as a float (+0 when no box intersects it or every intersecting value is +0)
__device__ __forceinline__ float row_max_of(unsigned long long key) {
    const unsigned k = (unsigned)(key >> 32);
    return k > 0x80000000u ? from_orderable(k) : 0.f;
}

__global__ void __launch_bounds__(kMT)
k_match_rows(const float4* __restrict__ gt, int M, const float4* __restrict__ boxes, int N,
             const unsigned long long* __restrict__ row_key, const long long* __restrict__ all_matches,
             long long* __restrict__ matches) {
    __shared__ float4 sg[kMGt];
    __shared__ float sa[kMGt];
    __shared__ float smax[kMGt];
    const int tid = threadIdx.x;
    const int n = blockIdx.x * kMT + tid;
    const bool valid = n < N;
    const float4 b = valid ? boxes[n] : make_float4(0.f, 0.f, 0.f, 0.f);
    const float ab = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
    bool restore = false;
    for (int m0 = 0; m0 < M; m0 += kMGt) {
        const int mm = min(kMGt, M - m0);
        __syncthreads();
        for (int i = tid; i < mm; i += kMT) {
            const float4 g = gt[m0 + i];
            sg[i] = g;
            sa[i] = __fmul_rn(__fsub_rn(g.z, g.x), __fsub_rn(g.w, g.y));
            smax[i] = row_max_of(row_key[m0 + i]);
        }
        __syncthreads();
        if (valid && !restore)
            for (int i = 0; i < mm; ++i) {
                bool hit;
                const float v = tv_iou(sg[i], sa[i], b, ab, hit);
                if (v == smax[i]) { restore = true; break; }        // ties included (:327-329); NaN never equals
            }
    }
    if (valid && restore) matches[n] = all_matches[n];
}

// SSDMatcher: matches[argmax_n q[m, :]] = m for every ground truth, in ascending m (later m wins a shared box)
__global__ void k_match_ssd(const unsigned long long* __restrict__ row_key, int M, long long* __restrict__ matches) {
    if (threadIdx.x == 0 && blockIdx.x == 0)
        for (int m = 0; m < M; ++m) {
            const unsigned long long key = row_key[m];
            const unsigned idx = (unsigned)(key >> 32) > 0x80000000u ? ~(unsigned)(key & 0xffffffffull) : 0u;   // all-zero row: index 0
            matches[idx] = m;
        }
}

// dense form for SSDMatcher.__call__(match_quality_matrix): row arg-max (first index) of a materialised matrix
__global__ void __launch_bounds__(256)
k_row_argmax(const float* __restrict__ q, int M, int N, unsigned long long* __restrict__ row_key) {
    __shared__ unsigned long long red[8];
    const int m = blockIdx.x;
    const float* row = q + (size_t)m * N;
    unsigned long long best = 0ull;
    for (int n = threadIdx.x; n < N; n += 256) {
        const float v = row[n];
        // NaN sorts above everything for torch.max; orderable() already places positive NaNs at the top
        const unsigned long long key = ((unsigned long long)orderable(v) << 32) | (unsigned long long)(~(unsigned)n);
        best = key > best ? key : best;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const unsigned long long u = __shfl_xor_sync(kFullMask, best, o); best = u > best ? u : best; }
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) best = red[w] > best ? red[w] : best;
        // store in the k_match_ssd convention: any value counts (also <= 0): flag the key as "positive"
        row_key[m] = (0xffffffffull << 32) | (best & 0xffffffffull);
    }
}

}  // namespace b200

using namespace b200;

extern "C" {

size_t b200_match_boxes_workspace_bytes(int32_t m, int32_t n) {
    if (m < 1 || n < 1) return 0;
    return align_up(sizeof(unsigned long long) * (size_t)m, 256) + align_up(sizeof(long long) * (size_t)n, 256);
}

int b200_match_boxes(const float* gt_boxes, int32_t m, const float* boxes, int32_t n, float high_thr, float low_thr,
                     int32_t allow_low_quality, int32_t ssd, int64_t* matches, float* matched_vals, void* workspace,
                     size_t workspace_bytes, void* stream) {
    if (m < 1 || n < 1 || !gt_boxes || !boxes || !matches) return B200_ERR_INVALID;
    if ((reinterpret_cast<uintptr_t>(gt_boxes) & 15u) || (reinterpret_cast<uintptr_t>(boxes) & 15u)) return B200_ERR_INVALID;
    if (!workspace || (reinterpret_cast<uintptr_t>(workspace) & 255u) || workspace_bytes < b200_match_boxes_workspace_bytes(m, n))
        return B200_ERR_WORKSPACE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    unsigned long long* row_key = reinterpret_cast<unsigned long long*>(workspace);
    long long* all_matches = reinterpret_cast<long long*>(reinterpret_cast<unsigned char*>(workspace) +
                                                          align_up(sizeof(unsigned long long) * (size_t)m, 256));
    B200_CUDA_TRY(cudaMemsetAsync(row_key, 0, sizeof(unsigned long long) * (size_t)m, st));
    const int grid = cdiv(n, kMT);
    k_match_cols<<<grid, kMT, 0, st>>>(reinterpret_cast<const float4*>(gt_boxes), m, reinterpret_cast<const float4*>(boxes), n,
                                       high_thr, low_thr, reinterpret_cast<long long*>(matches),
                                       allow_low_quality ? all_matches : nullptr, matched_vals, row_key);
    if (allow_low_quality)
        k_match_rows<<<grid, kMT, 0, st>>>(reinterpret_cast<const float4*>(gt_boxes), m, reinterpret_cast<const float4*>(boxes), n,
                                           row_key, all_matches, reinterpret_cast<long long*>(matches));
    if (ssd) k_match_ssd<<<1, 32, 0, st>>>(row_key, m, reinterpret_cast<long long*>(matches));
    return cudaGetLastError() == cudaSuccess ? B200_OK : B200_ERR_CUDA;
}

// SSDMatcher.__call__ on a materialised quality matrix: Matcher.__call__ (b200_matcher) followed by this override
int b200_matcher_ssd_override(const float* quality, int32_t m, int32_t n, int64_t* matches, void* workspace,
                              size_t workspace_bytes, void* stream) {
    if (m < 1 || n < 1 || !quality || !matches) return B200_ERR_INVALID;
    if (!workspace || (reinterpret_cast<uintptr_t>(workspace) & 7u) || workspace_bytes < sizeof(unsigned long long) * (size_t)m)
        return B200_ERR_WORKSPACE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    unsigned long long* row_key = reinterpret_cast<unsigned long long*>(workspace);
    k_row_argmax<<<m, 256, 0, st>>>(quality, m, n, row_key);
    k_match_ssd<<<1, 32, 0, st>>>(row_key, m, reinterpret_cast<long long*>(matches));
    return cudaGetLastError() == cudaSuccess ? B200_OK : B200_ERR_CUDA;
}

}  // extern "C"
