// roi.cu -- (a9) legacy per-head YOLO decode and (a13) ROI-head box post-process front end (sm_100a).
//
//  k_legacy_decode : YOLOLoss.forward(input, targets=None) of yolo/nets/yolo_loss.py:34-105 -- one head,
//                    output rows ordered (a, h, w), xy = (sigmoid + grid) * stride, wh = exp * anchor * stride,
//                    sigmoid objectness and classes.  Per (b, a) it is a [5+C, H*W] -> [H*W, 5+C] transpose
//                    with an element-wise transform: 32x32 tiles through shared memory, both sides coalesced,
//                    every byte read once and written once (HBM bound).
//  k_roi_candidates: RoIHeads.postprocess_detections (torchvision_models/tvision/roi_heads.py:715-767) up to
//                    the NMS: class scores (softmax / gombit / sigmoid of tfidf*logits, :724-729), per-class box
//                    decode (BoxCoder.decode_single, _utils.py:186-223), clip to the image, drop the background
//                    column, score > thr, small-box filter -- fused, one warp per proposal row, candidates
//                    compacted into the per-image slab the shared NMS kernels consume (nms.cu, labels = class).
//                    The reference materialises [R, C] scores and [R, C, 4] boxes before it filters.
#include "decode.cuh"
#include "nms.cuh"

namespace b200 {

// ------------------------------------------------------------------------------------------ a9
__global__ void __launch_bounds__(256)
k_legacy_decode(const float* __restrict__ head, int A, int CH, int H, int W, float stride_w, float stride_h,
                const float* __restrict__ anchors /* [A][2] scaled */, float* __restrict__ out) {
    __shared__ float tile[32][33];
    const int HW = H * W;
    const int ba = blockIdx.z, a = ba % A;
    const int cell0 = blockIdx.x * 32, ch0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;            // 32 x 8
    const float* src = head + (size_t)ba * CH * HW;
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int ch = ch0 + r, cell = cell0 + tx;
        tile[r][tx] = (ch < CH && cell < HW) ? ldg_stream_f32(src + (size_t)ch * HW + cell) : 0.f;
    }
    __syncthreads();
    float* dst = out + (size_t)ba * HW * CH;                              // row n = a*H*W + h*W + w of image b
    const float aw = anchors[2 * a], ah = anchors[2 * a + 1];
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int cell = cell0 + r, ch = ch0 + tx;
        if (cell >= HW || ch >= CH) continue;
        const float t = tile[tx][r];
        float v;
        if (ch >= 4) v = sigmoid_ref(t);                                               // conf, classes (:80-81)
        else if (ch == 0) v = __fmul_rn(__fadd_rn(sigmoid_ref(t), (float)(cell % W)), stride_w);   // :97,:103
        else if (ch == 1) v = __fmul_rn(__fadd_rn(sigmoid_ref(t), (float)(cell / W)), stride_h);   // :98
        else if (ch == 2) v = __fmul_rn(__fmul_rn(expf(t), aw), stride_w);                        // :99
        else v = __fmul_rn(__fmul_rn(expf(t), ah), stride_h);                                     // :100
        dst[(size_t)cell * CH + ch] = v;
    }
}

int launch_legacy_dense(const float* head, int B, int A, int C, int H, int W, float stride_w, float stride_h,
                        const float* anchors_dev, float* out, cudaStream_t stream);

int launch_legacy_decode(const float* head, int B, int A, int C, int H, int W, float stride_w, float stride_h,
                         const float* anchors_dev, float* out, cudaStream_t st) {
    // the staged 64-cell kernel of decode.cu when the channel count fits in shared memory, else the tiled transpose
    const int rc = launch_legacy_dense(head, B, A, C, H, W, stride_w, stride_h, anchors_dev, out, st);
    if (rc != 1) return rc;
    const int CH = 5 + C, HW = H * W;
    dim3 grid(cdiv(HW, 32), cdiv(CH, 32), B * A);
    if (grid.y > 65535 || grid.z > 65535) return B200_ERR_INVALID;
    k_legacy_decode<<<grid, 256, 0, st>>>(head, A, CH, H, W, stride_w, stride_h, anchors_dev, out);
    return cudaGetLastError() == cudaSuccess ? B200_OK : B200_ERR_CUDA;
}

// ------------------------------------------------------------------------------------------ a13
struct RoiParams {
    const float* logits;      // [R, C]
    const float* deltas;      // [R, 4C]
    const float* proposals;   // [R, 4]
    const int* row_offsets;   // [B+1] rows of image b = [row_offsets[b], row_offsets[b+1])
    const float* image_hw;    // [B, 2]
    const float* tfidf;       // [C] or nullptr (== 1)
    int R, C, B, activation;  // 0 softmax, 1 gombit, 2 sigmoid
    float wx, wy, ww, wh, clip, score_thr, min_size;
    Cand* slab;
    int cap;
    int* count;
    int* status;
    int shared_deltas;        // 1: deltas [R, 4], one box per row for every class (SSD); 0: [R, 4C] (ROI heads)
    int class_major;          // canonical candidate order: 1 = (class, row) (SSD's per-class loop), 0 = (row, class)
};

__device__ __forceinline__ int image_of_row(const int* off, int B, int r) {
    int lo = 0, hi = B;                       // largest b with off[b] <= r
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (off[mid] <= r) lo = mid; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(256)
k_roi_candidates(const __grid_constant__ RoiParams p) {
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= p.R) return;
    const int C = p.C;
    const float* row = p.logits + (size_t)r * C;
    // softmax statistics over ALL classes (background included), roi_heads.py:725
    float m = -INFINITY, s = 0.f;
    if (p.activation == 0) {
        for (int c = lane; c < C; c += 32) {
            const float x = p.tfidf ? __fmul_rn(p.tfidf[c], row[c]) : row[c];
            m = fmaxf(m, x);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(kFullMask, m, o));
        for (int c = lane; c < C; c += 32) {
            const float x = p.tfidf ? __fmul_rn(p.tfidf[c], row[c]) : row[c];
            s = __fadd_rn(s, expf(__fsub_rn(x, m)));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s = __fadd_rn(s, __shfl_xor_sync(kFullMask, s, o));
    }
    const int b = image_of_row(p.row_offsets, p.B, r);
    const float img_h = p.image_hw[2 * b], img_w = p.image_hw[2 * b + 1];
    const float4 a = reinterpret_cast<const float4*>(p.proposals)[r];
    const float w = __fsub_rn(a.z, a.x), h = __fsub_rn(a.w, a.y);                       // _utils.py:199-202
    const float cx = __fadd_rn(a.x, __fmul_rn(0.5f, w)), cy = __fadd_rn(a.y, __fmul_rn(0.5f, h));
    const unsigned lt = (1u << lane) - 1u;
    for (int c0 = 1; c0 < C; c0 += 32) {                 // class 0 is the background column (:752-754)
        const int c = c0 + lane;
        bool pass = false;
        float score = 0.f;
        Box q{0.f, 0.f, 0.f, 0.f};
        if (c < C) {
            const float x = p.tfidf ? __fmul_rn(p.tfidf[c], row[c]) : row[c];
            if (p.activation == 0) score = __fdiv_rn(expf(__fsub_rn(x, m)), s);
            else if (p.activation == 2) score = sigmoid_ref(x);
            else {
                // gombit: 1/exp(exp(-tfidf*(logit-1.96)))   (:727)
                const float z = p.tfidf ? __fmul_rn(-p.tfidf[c], __fsub_rn(row[c], 1.96f)) : -__fsub_rn(row[c], 1.96f);
                score = __fdiv_rn(1.0f, expf(expf(z)));
            }
            if (score > p.score_thr) {                                                    // :763
                const float4 d = reinterpret_cast<const float4*>(p.deltas)[p.shared_deltas ? (size_t)r : (size_t)r * C + c];
                const float dx = __fdiv_rn(d.x, p.wx), dy = __fdiv_rn(d.y, p.wy);          // _utils.py:205-208
                const float dw = fminf(__fdiv_rn(d.z, p.ww), p.clip), dh = fminf(__fdiv_rn(d.w, p.wh), p.clip);
                const float pcx = __fadd_rn(__fmul_rn(dx, w), cx), pcy = __fadd_rn(__fmul_rn(dy, h), cy);
                const float pw = __fmul_rn(expf(dw), w), ph = __fmul_rn(expf(dh), h);
                q.x1 = __fsub_rn(pcx, __fmul_rn(0.5f, pw)); q.y1 = __fsub_rn(pcy, __fmul_rn(0.5f, ph));
                q.x2 = __fadd_rn(pcx, __fmul_rn(0.5f, pw)); q.y2 = __fadd_rn(pcy, __fmul_rn(0.5f, ph));
                // clip_boxes_to_image: x in [0, width], y in [0, height]               (:746)
                q.x1 = fminf(fmaxf(q.x1, 0.f), img_w); q.x2 = fminf(fmaxf(q.x2, 0.f), img_w);
                q.y1 = fminf(fmaxf(q.y1, 0.f), img_h); q.y2 = fminf(fmaxf(q.y2, 0.f), img_h);
                // remove_small_boxes(min_size)                                           (:767)
                pass = (__fsub_rn(q.x2, q.x1) >= p.min_size) && (__fsub_rn(q.y2, q.y1) >= p.min_size);
            }
        }
        const unsigned bal = __ballot_sync(kFullMask, pass);
        if (bal == 0u) continue;
        int slot0 = 0;
        if (lane == 0) slot0 = atomicAdd(p.count + b, __popc(bal));
        slot0 = __shfl_sync(kFullMask, slot0, 0);
        if (!pass) continue;
        const int slot = slot0 + __popc(bal & lt);
        if (slot >= p.cap) { atomicOr(p.status, 1); continue; }
        float4* d4 = reinterpret_cast<float4*>(p.slab + (size_t)b * (size_t)p.cap + (size_t)slot);
        d4[0] = make_float4(q.x1, q.y1, q.x2, q.y2);
        // canonical order inside the image = the reference's flattened (row, class) order
        const int rows_b = p.row_offsets[b + 1] - p.row_offsets[b];
        const int flat = p.class_major ? (c - 1) * rows_b + (r - p.row_offsets[b]) : (r - p.row_offsets[b]) * (C - 1) + (c - 1);
        d4[1] = make_float4(score, __int_as_float(c), __int_as_float(flat), 0.f);
    }
}

// SSD keeps at most `topk` candidates PER CLASS (score.topk(min(topk_candidates, n)), ssd.py:407-409) before the
// NMS.  One CTA per image: class histogram of the slab; classes above the limit (rare) rank their members by
// (score desc, canonical index asc) and drop the tail; the slab is compacted in place.
__global__ void __launch_bounds__(1024, 1)
k_class_topk(Cand* __restrict__ slab_all, int* __restrict__ count, int cap, int C, int topk, int* __restrict__ status) {
    extern __shared__ int cls_cnt[];                  // [C]
    __shared__ int s_over, s_scan[32], s_total;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    Cand* slab = slab_all + (size_t)b * cap;
    const int n = min(count[b], cap);
    for (int c = tid; c < C; c += 1024) cls_cnt[c] = 0;
    if (tid == 0) s_over = 0;
    __syncthreads();
    for (int i = tid; i < n; i += 1024) atomicAdd(&cls_cnt[slab[i].label], 1);
    __syncthreads();
    for (int c = tid; c < C; c += 1024) if (cls_cnt[c] > topk) s_over = 1;
    __syncthreads();
    if (!s_over) return;
    // rank inside the class, one warp per member of an over-full class
    for (int i = warp; i < n; i += 32) {
        const Cand ci = slab[i];
        int keep = 1;
        if (cls_cnt[ci.label] > topk) {
            int better = 0;
            for (int j0 = 0; j0 < n; j0 += 32) {
                const int j = j0 + lane;
                bool bt = false;
                if (j < n) {
                    const Cand cj = slab[j];
                    bt = cj.label == ci.label && (cj.score > ci.score || (cj.score == ci.score && cj.anchor < ci.anchor));
                }
                better += __popc(__ballot_sync(kFullMask, bt));
            }
            keep = better < topk;
        }
        if (lane == 0) slab[i].pad = keep;
    }
    __syncthreads();
    // in-place compaction: chunks of 1024 rows are read completely before they are written (destination <= source)
    int at = 0;
    for (int i0 = 0; i0 < n; i0 += 1024) {
        const int i = i0 + tid;
        Cand c{};
        bool keep = false;
        if (i < n) { c = slab[i]; keep = c.pad != 0; }
        const unsigned bal = __ballot_sync(kFullMask, keep);
        if (lane == 0) s_scan[warp] = __popc(bal);
        __syncthreads();
        int before = 0, total = 0;
        for (int w = 0; w < 32; ++w) { const int v = s_scan[w]; if (w < warp) before += v; total += v; }
        if (keep) { c.pad = 0; slab[at + before + __popc(bal & ((1u << lane) - 1u))] = c; }
        at += total;
        __syncthreads();
    }
    // an overflowed slab keeps its TRUE count (status bit 0 is set): the caller sizes the retry from it
    if (tid == 0 && count[b] <= cap) count[b] = at;
    (void)status; (void)s_total;
}

int launch_roi_candidates(const RoiParams& p, cudaStream_t st) {
    if (p.R <= 0) return B200_OK;
    k_roi_candidates<<<cdiv(p.R, 8), 256, 0, st>>>(p);
    return cudaGetLastError() == cudaSuccess ? B200_OK : B200_ERR_CUDA;
}

}  // namespace b200

// ------------------------------------------------------------------------------------------ C ABI
using namespace b200;

extern "C" {

int b200_yolo_legacy_decode(const float* head, int32_t batch, int32_t num_anchors, int32_t num_classes,
                            int32_t in_h, int32_t in_w, float stride_w, float stride_h,
                            const float* anchors_scaled, float* out, void* stream) {
    if (!head || !anchors_scaled || !out || batch < 1 || num_anchors < 1 || num_classes < 1 || in_h < 1 || in_w < 1)
        return B200_ERR_INVALID;
    return launch_legacy_decode(head, batch, num_anchors, num_classes, in_h, in_w, stride_w, stride_h,
                                anchors_scaled, out, static_cast<cudaStream_t>(stream));
}

size_t b200_roi_workspace_bytes(int32_t batch, int32_t capacity) {
    if (batch < 1 || capacity < 1) return 0;
    const size_t T = (size_t)batch * (size_t)capacity;
    return align_up(sizeof(int) * (size_t)batch, 256) + align_up(sizeof(Cand) * T, 256) +
           nms_scratch_bytes(T, (size_t)batch, (size_t)capacity) + 512;
}

int b200_roi_postprocess(const float* class_logits, const float* box_regression, const float* proposals,
                         const int32_t* row_offsets, int32_t batch, int32_t total_rows, int32_t num_classes,
                         const float* image_hw, const float* tfidf, int32_t activation,
                         const float* weights_host, float xform_clip, float score_thr, float min_size,
                         double nms_thr, int32_t nms_mode, int32_t capacity, int32_t max_det, float* det,
                         int32_t* det_keep, int32_t* det_count, int32_t* cand_count, int32_t* status,
                         void* workspace, size_t workspace_bytes, void* stream) {
    if (!class_logits || !box_regression || !proposals || !row_offsets || !image_hw || !weights_host || !det ||
        !det_count || !status || batch < 1 || total_rows < 0 || num_classes < 2 || capacity < 1 || max_det < 1)
        return B200_ERR_INVALID;
    if (activation < 0 || activation > 2) return B200_ERR_INVALID;
    if (nms_mode != B200_NMS_TV_CLASS && nms_mode != B200_NMS_TV_TRICK && nms_mode != B200_NMS_TV_AUTO) return B200_ERR_INVALID;
    if ((reinterpret_cast<uintptr_t>(proposals) & 15u) || (reinterpret_cast<uintptr_t>(box_regression) & 15u))
        return B200_ERR_INVALID;
    if (!workspace || (reinterpret_cast<uintptr_t>(workspace) & 255u) ||
        workspace_bytes < b200_roi_workspace_bytes(batch, capacity))
        return B200_ERR_WORKSPACE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t T = (size_t)batch * (size_t)capacity;
    unsigned char* q = reinterpret_cast<unsigned char*>(workspace);
    int* count = reinterpret_cast<int*>(q);                q += align_up(sizeof(int) * (size_t)batch, 256);
    Cand* slab = reinterpret_cast<Cand*>(q);               q += align_up(sizeof(Cand) * T, 256);
    const size_t nms_bytes = nms_scratch_bytes(T, (size_t)batch, (size_t)capacity);
    B200_CUDA_TRY(cudaMemsetAsync(count, 0, sizeof(int) * (size_t)batch, st));
    RoiParams p{};
    p.logits = class_logits; p.deltas = box_regression; p.proposals = proposals; p.row_offsets = row_offsets;
    p.image_hw = image_hw; p.tfidf = tfidf; p.R = total_rows; p.C = num_classes; p.B = batch; p.activation = activation;
    p.wx = weights_host[0]; p.wy = weights_host[1]; p.ww = weights_host[2]; p.wh = weights_host[3];
    p.clip = xform_clip; p.score_thr = score_thr; p.min_size = min_size;
    p.slab = slab; p.cap = capacity; p.count = count; p.status = status;
    const int rc = launch_roi_candidates(p, st);
    if (rc != B200_OK) return rc;
    NmsParams np{};
    if (!nms_carve_scratch(&np, T, (size_t)batch, (size_t)capacity, q, nms_bytes)) return B200_ERR_WORKSPACE;
    np.mode = nms_mode;
    np.thr_f = (float)nms_thr;
    np.thr_d = nms_thr;
    np.status = status;
    np.det = det; np.det_keep = det_keep; np.det_anchor = nullptr; np.det_count = det_count;
    np.cand_count_out = cand_count;
    np.max_det = max_det;
    np.slab = slab; np.count = count; np.cap = capacity; np.from_slab = 1; np.serial = 1;
    np.anchor_space = 0;                 // flat (row, class) indices: rank by counting
    np.max_seg = capacity;
    return launch_nms(np, batch, st);
}

// SSD.postprocess_detections (ssd.py:386-430)
int b200_ssd_postprocess(const float* cls_logits, const float* bbox_regression, const float* anchors,
                         const int32_t* row_offsets, int32_t batch, int32_t total_rows, int32_t num_classes,
                         const float* image_hw, const float* tfidf, const float* weights_host, float xform_clip,
                         float score_thr, int32_t topk_per_class, double nms_thr, int32_t nms_mode, int32_t capacity,
                         int32_t max_det, float* det, int32_t* det_count, int32_t* cand_count, int32_t* status,
                         void* workspace, size_t workspace_bytes, void* stream) {
    if (!cls_logits || !bbox_regression || !anchors || !row_offsets || !image_hw || !weights_host || !det || !det_count ||
        !status || batch < 1 || total_rows < 0 || num_classes < 2 || num_classes > 8192 || capacity < 1 || max_det < 1 ||
        topk_per_class < 1)
        return B200_ERR_INVALID;
    if (nms_mode != B200_NMS_TV_CLASS && nms_mode != B200_NMS_TV_TRICK && nms_mode != B200_NMS_TV_AUTO) return B200_ERR_INVALID;
    if ((reinterpret_cast<uintptr_t>(anchors) & 15u) || (reinterpret_cast<uintptr_t>(bbox_regression) & 15u)) return B200_ERR_INVALID;
    if (!workspace || (reinterpret_cast<uintptr_t>(workspace) & 255u) || workspace_bytes < b200_roi_workspace_bytes(batch, capacity))
        return B200_ERR_WORKSPACE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t T = (size_t)batch * (size_t)capacity;
    unsigned char* q = reinterpret_cast<unsigned char*>(workspace);
    int* count = reinterpret_cast<int*>(q);                q += align_up(sizeof(int) * (size_t)batch, 256);
    Cand* slab = reinterpret_cast<Cand*>(q);               q += align_up(sizeof(Cand) * T, 256);
    const size_t nms_bytes = nms_scratch_bytes(T, (size_t)batch, (size_t)capacity);
    B200_CUDA_TRY(cudaMemsetAsync(count, 0, sizeof(int) * (size_t)batch, st));
    RoiParams p{};
    p.logits = cls_logits; p.deltas = bbox_regression; p.proposals = anchors; p.row_offsets = row_offsets;
    p.image_hw = image_hw; p.tfidf = tfidf; p.R = total_rows; p.C = num_classes; p.B = batch; p.activation = 0;
    p.wx = weights_host[0]; p.wy = weights_host[1]; p.ww = weights_host[2]; p.wh = weights_host[3];
    p.clip = xform_clip; p.score_thr = score_thr; p.min_size = -INFINITY;       // SSD has no small-box filter
    p.slab = slab; p.cap = capacity; p.count = count; p.status = status;
    p.shared_deltas = 1; p.class_major = 1;
    const int rc = launch_roi_candidates(p, st);
    if (rc != B200_OK) return rc;
    k_class_topk<<<batch, 1024, sizeof(int) * (size_t)num_classes, st>>>(slab, count, capacity, num_classes, topk_per_class, status);
    NmsParams np{};
    if (!nms_carve_scratch(&np, T, (size_t)batch, (size_t)capacity, q, nms_bytes)) return B200_ERR_WORKSPACE;
    np.mode = nms_mode;
    np.thr_f = (float)nms_thr;
    np.thr_d = nms_thr;
    np.status = status;
    np.det = det; np.det_keep = nullptr; np.det_anchor = nullptr; np.det_count = det_count;
    np.cand_count_out = cand_count;
    np.max_det = max_det;
    np.slab = slab; np.count = count; np.cap = capacity; np.from_slab = 1; np.serial = 1;
    np.anchor_space = 0;
    np.max_seg = capacity;
    return launch_nms(np, batch, st);
}

}  // extern "C"
