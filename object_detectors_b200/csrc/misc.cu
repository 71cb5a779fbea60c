// misc.cu -- small element-wise / reduction kernels that complete the drop-in surface (sm_100a):
//   k_abs_coord      helper.get_abs_coord                      (yolo/utilities/helper.py:203-217)
//   k_boxcoder       BoxCoder.decode_single                    (tvision/_utils.py:186-223)
//   k_boxcoder_encode encode_boxes / BoxCoder.encode_single    (tvision/_utils.py:80-125, 160-166)
//   k_matcher_*      Matcher.__call__ incl. low-quality ties   (tvision/_utils.py:271-344)
// All are memory-bound one-pass kernels; arithmetic follows the reference operation by operation.
#include "common.cuh"

namespace b200 {

__global__ void __launch_bounds__(256)
k_abs_coord(const float4* __restrict__ in, long long n, float4* __restrict__ out) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const float4 v = in[i];
    const Box b = abs_coord(v.x, v.y, v.z, v.w);
    out[i] = make_float4(b.x1, b.y1, b.x2, b.y2);
}

// rel_codes [n, 4k], boxes [n,4] -> out [n, 4k]; one thread per (row, class slot)
__global__ void __launch_bounds__(256)
k_boxcoder(const float4* __restrict__ rel, const float4* __restrict__ boxes, long long n, int k, float wx,
           float wy, float ww, float wh, float clip, float4* __restrict__ out) {
    const long long e = (long long)blockIdx.x * 256 + threadIdx.x;
    if (e >= n * k) return;
    const long long row = e / k;
    const float4 a = boxes[row];
    const float4 d = rel[e];
    const float w = __fsub_rn(a.z, a.x), h = __fsub_rn(a.w, a.y);                       // :199-200
    const float cx = __fadd_rn(a.x, __fmul_rn(0.5f, w)), cy = __fadd_rn(a.y, __fmul_rn(0.5f, h));
    const float dx = __fdiv_rn(d.x, wx), dy = __fdiv_rn(d.y, wy);                       // :205-208
    const float dw = fminf(__fdiv_rn(d.z, ww), clip), dh = fminf(__fdiv_rn(d.w, wh), clip);   // :211-212
    const float pcx = __fadd_rn(__fmul_rn(dx, w), cx), pcy = __fadd_rn(__fmul_rn(dy, h), cy);
    const float pw = __fmul_rn(expf(dw), w), ph = __fmul_rn(expf(dh), h);
    out[e] = make_float4(__fsub_rn(pcx, __fmul_rn(0.5f, pw)), __fsub_rn(pcy, __fmul_rn(0.5f, ph)),
                         __fadd_rn(pcx, __fmul_rn(0.5f, pw)), __fadd_rn(pcy, __fmul_rn(0.5f, ph)));
}

// encode_boxes (tvision/_utils.py:80-125): regression targets of proposals w.r.t. their matched reference boxes
__global__ void __launch_bounds__(256)
k_boxcoder_encode(const float4* __restrict__ ref, const float4* __restrict__ prop, long long n, float wx, float wy,
                  float ww, float wh, float4* __restrict__ out) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const float4 p = prop[i], g = ref[i];
    const float ew = __fsub_rn(p.z, p.x), eh = __fsub_rn(p.w, p.y);                          // :108-109
    const float ex = __fadd_rn(p.x, __fmul_rn(0.5f, ew)), ey = __fadd_rn(p.y, __fmul_rn(0.5f, eh));
    const float gw = __fsub_rn(g.z, g.x), gh = __fsub_rn(g.w, g.y);                          // :113-114
    const float gx = __fadd_rn(g.x, __fmul_rn(0.5f, gw)), gy = __fadd_rn(g.y, __fmul_rn(0.5f, gh));
    out[i] = make_float4(__fdiv_rn(__fmul_rn(wx, __fsub_rn(gx, ex)), ew),                   // :118-121
                         __fdiv_rn(__fmul_rn(wy, __fsub_rn(gy, ey)), eh),
                         __fmul_rn(ww, logf(__fdiv_rn(gw, ew))),
                         __fmul_rn(wh, logf(__fdiv_rn(gh, eh))));
}

int launch_boxcoder_encode(const float* ref, const float* prop, long long n, const float* weights, float* out,
                           cudaStream_t st) {
    if (n <= 0) return B200_OK;
    k_boxcoder_encode<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(
        reinterpret_cast<const float4*>(ref), reinterpret_cast<const float4*>(prop), n, weights[0], weights[1], weights[2],
        weights[3], reinterpret_cast<float4*>(out));
    return cudaGetLastError() == cudaSuccess ? B200_OK : B200_ERR_CUDA;
}

// column pass: matched_vals, matches = quality.max(dim=0) (first maximum), then the thresholds
__global__ void __launch_bounds__(256)
k_matcher_cols(const float* __restrict__ q, int M, int N, float high, float low, long long* __restrict__ matches,
               long long* __restrict__ all_matches) {
    const int n = blockIdx.x * 256 + threadIdx.x;
    if (n >= N) return;
    float best = q[n];
    int arg = 0;
    for (int m = 1; m < M; ++m) {
        const float v = q[(size_t)m * N + n];
        if (v > best || (v != v && best == best)) { best = v; arg = m; }     // torch.max propagates NaN
    }
    if (all_matches) all_matches[n] = arg;
    long long r = arg;
    if (best < low) r = -1;                                   // BELOW_LOW_THRESHOLD   (:300-306)
    else if (best >= low && best < high) r = -2;              // BETWEEN_THRESHOLDS
    matches[n] = r;
}

// row pass: highest quality per ground truth; every prediction that ties it gets its match back
__global__ void __launch_bounds__(256)
k_matcher_rows(const float* __restrict__ q, int M, int N, const long long* __restrict__ all_matches,
               long long* __restrict__ matches) {
    __shared__ float red[8];
    const int m = blockIdx.x;
    const float* row = q + (size_t)m * N;
    float best = -INFINITY;
    for (int n = threadIdx.x; n < N; n += 256) best = fmaxf(best, row[n]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = fmaxf(best, __shfl_xor_sync(kFullMask, best, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = best;
    __syncthreads();
    best = red[0];
    for (int w = 1; w < 8; ++w) best = fmaxf(best, red[w]);
    for (int n = threadIdx.x; n < N; n += 256)
        if (row[n] == best) matches[n] = all_matches[n];       // set_low_quality_matches_ (:315-344)
}

int launch_abs_coord(const float* in, long long n, float* out, cudaStream_t st) {
    if (n <= 0) return B200_OK;
    k_abs_coord<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(reinterpret_cast<const float4*>(in), n,
                                                              reinterpret_cast<float4*>(out));
    return cudaGetLastError() == cudaSuccess ? B200_OK : B200_ERR_CUDA;
}

int launch_boxcoder(const float* rel, const float* boxes, long long n, int k, const float* weights, float clip,
                    float* out, cudaStream_t st) {
    if (n <= 0 || k <= 0) return B200_OK;
    k_boxcoder<<<(unsigned)((n * k + 255) / 256), 256, 0, st>>>(
        reinterpret_cast<const float4*>(rel), reinterpret_cast<const float4*>(boxes), n, k, weights[0], weights[1],
        weights[2], weights[3], clip, reinterpret_cast<float4*>(out));
    return cudaGetLastError() == cudaSuccess ? B200_OK : B200_ERR_CUDA;
}

int launch_matcher(const float* q, int M, int N, float high, float low, int allow_low_quality, long long* matches,
                   long long* all_matches_ws, cudaStream_t st) {
    k_matcher_cols<<<cdiv(N, 256), 256, 0, st>>>(q, M, N, high, low, matches, allow_low_quality ? all_matches_ws : nullptr);
    if (allow_low_quality) k_matcher_rows<<<M, 256, 0, st>>>(q, M, N, all_matches_ws, matches);
    return cudaGetLastError() == cudaSuccess ? B200_OK : B200_ERR_CUDA;
}

}  // namespace b200
