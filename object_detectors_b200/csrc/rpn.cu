// rpn.cu -- RPN proposal filter (sm_100a): per-level top-k -> decode -> clip/filter -> NMS -> top-n.
// Replaces BoxCoder.decode_single + RegionProposalNetwork._get_top_n_idx / filter_proposals
// (torchvision_models/tvision/_utils.py:186-223, rpn.py:215-280).
//
//  k_rpn_select : one CTA per (image, level).  8-bit MSB radix select over the level's raw
//                 objectness logits (4 histogram passes, L2-resident after the first) finds the
//                 k-th largest key; the selected <= k entries are sorted in shared memory
//                 (score desc, index asc == Tensor.topk order for tie-free input) and only those
//                 are decoded / clipped / filtered -- the reference decodes all 268 569 anchors
//                 first (rpn.py:355).  Survivors are written in order at a fixed stride.
//  k_rpn_units  : (coordinate-trick strategy) the image's shift unit for its per-level segments (see below).
//  NMS          : the shared per-segment kernel of nms.cu (segments = image x level for the
//                 vanilla strategy, = image for the coordinate trick).
//  k_rpn_finish : merges an image's kept lists by score and emits the first post_nms_top_n.
#include "nms.cuh"

namespace b200 {

static constexpr int kSelThreads = 1024;
static constexpr int kSelWarps = kSelThreads / 32;
static constexpr int kMaxLevels = 16;
static constexpr float kXformClip = 4.135166556742356f;  // fp32(log(1000/16)), _utils.py:134

struct RpnParams {
    const float* obj;       // [B, total]
    const float* deltas;    // [B, total, 4]
    const float* anchors;   // [total, 4]
    const float* proposals; // [B, total, 4] already decoded boxes (then deltas / anchors are unused)
    const float* image_hw;  // [B, 2]
    int B, total, L, Ktot, pre_k, post_k;
    int level_off[kMaxLevels], level_n[kMaxLevels], level_k[kMaxLevels], level_koff[kMaxLevels];
    float score_thr, min_size;
    // survivors at [b*Ktot + koff_l + r]
    float4* box;
    float* score;
    int* label;
    int* aidx;
    int* seg_start;   // [B*L]
    int* seg_count;   // [B*L]
    long long* keep;  // [B*Ktot] NMS output (relative to segment start)
    int* keep_count;  // [B*L] or [B]
    int segs_per_img;
    float* out_boxes; float* out_scores; int* out_index; int* out_count;
};

__device__ __forceinline__ void bitonic_sort_u64(unsigned long long* key, int P) {
    for (int k = 2; k <= P; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < (P >> 1); t += blockDim.x) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int ixj = i | j;
                const bool asc = (i & k) == 0;
                const unsigned long long a = key[i], b = key[ixj];
                if ((a > b) == asc) { key[i] = b; key[ixj] = a; }
            }
            __syncthreads();
        }
}

__global__ void __launch_bounds__(kSelThreads, 1)
k_rpn_select(const __grid_constant__ RpnParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long* sel = reinterpret_cast<unsigned long long*>(smem_raw);  // [pow2(k)]
    __shared__ int hist[256];
    __shared__ unsigned s_prefix, s_mask;
    __shared__ int s_remaining, s_cnt, s_tie;
    __shared__ int s_scan[kSelWarps];

    const int l = blockIdx.x, b = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = P.level_n[l], k = P.level_k[l];
    const float* src = P.obj + (size_t)b * P.total + P.level_off[l];
    int Ppad = 1;
    while (Ppad < k) Ppad <<= 1;

    // ---- radix select: key of the k-th largest logit ------------------------------------------
    if (tid == 0) { s_prefix = 0u; s_mask = 0u; s_remaining = k; s_cnt = 0; s_tie = 0; }
    __syncthreads();
    if (n > k) {
        for (int pass = 3; pass >= 0; --pass) {
            for (int i = tid; i < 256; i += kSelThreads) hist[i] = 0;
            __syncthreads();
            const unsigned prefix = s_prefix, mask = s_mask;
            const int shift = 8 * pass;
            for (int i = tid; i < n; i += kSelThreads) {
                const unsigned key = orderable(__ldg(src + i));
                if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1);
            }
            __syncthreads();
            if (warp == 0) {
                // bins 255 .. 0, 8 per lane (lane 0 owns the top bins); find the bin holding rank `remaining`
                int local[8], sum = 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) { local[j] = hist[255 - (lane * 8 + j)]; sum += local[j]; }
                int incl = sum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_up_sync(kFullMask, incl, o);
                    if (lane >= o) incl += v;
                }
                int before = incl - sum;
                const int rem = s_remaining;
                if (before < rem && rem <= incl) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (before < rem && rem <= before + local[j]) {
                            s_prefix = prefix | ((unsigned)(255 - (lane * 8 + j)) << shift);
                            s_mask = mask | (255u << shift);
                            s_remaining = rem - before;
                        }
                        before += local[j];
                    }
                }
            }
            __syncthreads();
        }
    }
    const unsigned T = s_prefix;           // key of the k-th largest (exact after 4 passes)
    const int take_ties = s_remaining;     // how many entries equal to T are still needed
    // ---- collect the selected entries --------------------------------------------------------
    for (int i0 = 0; i0 < n; i0 += kSelThreads) {
        const int i = i0 + tid;
        if (i < n) {
            const unsigned key = orderable(__ldg(src + i));
            bool take = n <= k || key > T;
            if (!take && key == T) take = atomicAdd(&s_tie, 1) < take_ties;
            if (take) sel[atomicAdd(&s_cnt, 1)] = ((unsigned long long)(~key) << 32) | (unsigned)i;
        }
    }
    __syncthreads();
    const int got = s_cnt;  // == k
    for (int i = got + tid; i < Ppad; i += kSelThreads) sel[i] = ~0ull;
    __syncthreads();
    bitonic_sort_u64(sel, Ppad);

    // ---- decode / clip / filter the selected anchors, keep order ------------------------------
    const float img_h = P.image_hw[2 * b], img_w = P.image_hw[2 * b + 1];
    const size_t out0 = (size_t)b * P.Ktot + P.level_koff[l];
    int running = 0;
    for (int r0 = 0; r0 < got; r0 += kSelThreads) {
        const int r = r0 + tid;
        bool ok = false;
        float4 box = make_float4(0.f, 0.f, 0.f, 0.f);
        float prob = 0.f;
        int a = 0;
        if (r < got) {
            const int i = (int)(unsigned)sel[r];
            a = P.level_off[l] + i;
            prob = sigmoid_ref(src[i]);                                            // rpn.py:255
            float x1, y1, x2, y2;
            if (P.proposals) {
                const float4 pb = *reinterpret_cast<const float4*>(P.proposals + ((size_t)b * P.total + a) * 4);
                x1 = pb.x; y1 = pb.y; x2 = pb.z; y2 = pb.w;
            } else {
                const float4 an = *reinterpret_cast<const float4*>(P.anchors + 4 * (size_t)a);
                const float4 d = *reinterpret_cast<const float4*>(P.deltas + ((size_t)b * P.total + a) * 4);
                // BoxCoder.decode_single, weights (1,1,1,1)                           _utils.py:199-221
                const float w = __fsub_rn(an.z, an.x), h = __fsub_rn(an.w, an.y);
                const float cx = __fadd_rn(an.x, __fmul_rn(0.5f, w)), cy = __fadd_rn(an.y, __fmul_rn(0.5f, h));
                const float dw = fminf(d.z, kXformClip), dh = fminf(d.w, kXformClip);
                const float pcx = __fadd_rn(__fmul_rn(d.x, w), cx), pcy = __fadd_rn(__fmul_rn(d.y, h), cy);
                const float pw = __fmul_rn(expf(dw), w), ph = __fmul_rn(expf(dh), h);
                x1 = __fsub_rn(pcx, __fmul_rn(0.5f, pw)); y1 = __fsub_rn(pcy, __fmul_rn(0.5f, ph));
                x2 = __fadd_rn(pcx, __fmul_rn(0.5f, pw)); y2 = __fadd_rn(pcy, __fmul_rn(0.5f, ph));
            }
            // clip_boxes_to_image (rpn.py:260)
            x1 = fminf(fmaxf(x1, 0.f), img_w); x2 = fminf(fmaxf(x2, 0.f), img_w);
            y1 = fminf(fmaxf(y1, 0.f), img_h); y2 = fminf(fmaxf(y2, 0.f), img_h);
            box = make_float4(x1, y1, x2, y2);
            // remove_small_boxes + score threshold (rpn.py:263-269)
            ok = (__fsub_rn(x2, x1) >= P.min_size) && (__fsub_rn(y2, y1) >= P.min_size) && (prob >= P.score_thr);
        }
        const unsigned bal = __ballot_sync(kFullMask, ok);
        if (lane == 0) s_scan[warp] = __popc(bal);
        __syncthreads();
        int before = 0, total = 0;
        for (int w2 = 0; w2 < kSelWarps; ++w2) { const int c = s_scan[w2]; if (w2 < warp) before += c; total += c; }
        if (ok) {
            const size_t o = out0 + running + before + __popc(bal & ((1u << lane) - 1u));
            P.box[o] = box;
            P.score[o] = prob;
            P.label[o] = l;
            P.aidx[o] = a;
        }
        running += total;
        __syncthreads();
    }
    if (tid == 0) {
        P.seg_start[b * P.L + l] = (int)out0;
        P.seg_count[b * P.L + l] = running;
    }
}

// merge the kept lists of an image by score, emit the first post_k (rpn.py:272-278)
__global__ void __launch_bounds__(1024, 1)
k_rpn_finish(const __grid_constant__ RpnParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long* key = reinterpret_cast<unsigned long long*>(smem_raw);
    __shared__ int s_total;
    const int b = blockIdx.x, tid = threadIdx.x;
    const size_t base = (size_t)b * P.Ktot;
    if (tid == 0) s_total = 0;
    __syncthreads();
    const int* starts = P.seg_start;
    for (int s = 0; s < P.segs_per_img; ++s) {
        const int seg = b * P.segs_per_img + s;
        const int kc = P.keep_count[seg];
        const int st = starts[seg];
        const int at = s_total;
        for (int i = tid; i < kc; i += 1024) {
            const int pos = st + (int)P.keep[st + i];           // absolute row
            key[at + i] = ((unsigned long long)(~orderable(P.score[pos])) << 32) | (unsigned)(pos - (int)base);
        }
        __syncthreads();
        if (tid == 0) s_total = at + kc;
        __syncthreads();
    }
    const int total = s_total;
    {
        int Ppad = 1;
        while (Ppad < total) Ppad <<= 1;
        for (int i = total + tid; i < Ppad; i += 1024) key[i] = ~0ull;
        __syncthreads();
        bitonic_sort_u64(key, Ppad);
    }
    const int nout = min(total, P.post_k);
    for (int i = tid; i < nout; i += 1024) {
        const size_t pos = base + (unsigned)key[i];
        reinterpret_cast<float4*>(P.out_boxes)[(size_t)b * P.post_k + i] = P.box[pos];
        P.out_scores[(size_t)b * P.post_k + i] = P.score[pos];
        if (P.out_index) P.out_index[(size_t)b * P.post_k + i] = P.aidx[pos];
    }
    if (tid == 0) P.out_count[b] = nout;
}

// Coordinate-trick arithmetic on per-level segments: torchvision shifts level l by l * (max coordinate of the
// image's boxes + 1), so boxes of different levels never intersect and the one big NMS over the image decomposes
// exactly into one NMS per level on the SHIFTED boxes.  This kernel computes the image's unit and hands it to
// every (image, level) segment; the NMS then runs on L small segments instead of one of up to L*k boxes.
__global__ void __launch_bounds__(256)
k_rpn_units(const RpnParams P, float* __restrict__ units) {
    __shared__ float red[8];
    const int b = blockIdx.x, tid = threadIdx.x;
    float mx = -INFINITY;
    for (int l = 0; l < P.L; ++l) {
        const int n = P.seg_count[b * P.L + l];
        const float4* box = P.box + (size_t)b * P.Ktot + P.level_koff[l];
        for (int i = tid; i < n; i += 256) {
            const float4 v = box[i];
            mx = fmaxf(mx, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(kFullMask, mx, o));
    if ((tid & 31) == 0) red[tid >> 5] = mx;
    __syncthreads();
    mx = red[0];
    for (int w = 1; w < 8; ++w) mx = fmaxf(mx, red[w]);
    const float unit = __fadd_rn(mx, 1.0f);
    for (int l = tid; l < P.L; l += 256) units[b * P.L + l] = unit;
}

// ------------------------------------------------------------------------------------------ host
namespace {
struct RpnWs {
    float4* box; float* score; int* label; int* aidx;
    int* seg_start; int* seg_count;
    long long* keep; int* keep_count; float* units;
    void* nms; size_t nms_bytes;
};
// NMS scratch: B*L segments of <= pre_k boxes for either batched_nms strategy
size_t rpn_nms_bytes(int batch, int levels, int pre_k) {
    const size_t T = (size_t)batch * levels * pre_k;
    return nms_scratch_bytes(T, (size_t)batch * levels, (size_t)pre_k);
}
size_t rpn_carve(int batch, int levels, int pre_k, void* base, size_t bytes, RpnWs* w) {
    const size_t T = (size_t)batch * levels * pre_k;   // worst case rows (>= B*Ktot)
    unsigned char* p = reinterpret_cast<unsigned char*>(base);
    size_t used = 0;
    auto take = [&](size_t b) { b = align_up(b, 256); void* r = base ? p + used : nullptr; used += b; return r; };
    RpnWs t;
    t.box = (float4*)take(16 * T); t.score = (float*)take(4 * T); t.label = (int*)take(4 * T); t.aidx = (int*)take(4 * T);
    t.seg_start = (int*)take(4 * (size_t)batch * levels); t.seg_count = (int*)take(4 * (size_t)batch * levels);
    t.keep = (long long*)take(8 * T); t.keep_count = (int*)take(4 * (size_t)batch * levels);
    t.units = (float*)take(4 * (size_t)batch * levels);
    t.nms_bytes = rpn_nms_bytes(batch, levels, pre_k);
    t.nms = take(t.nms_bytes);
    if (base && used > bytes) return 0;
    if (w) *w = t;
    return used;
}
}  // namespace

size_t rpn_workspace_bytes(int batch, int total, int num_levels, int pre_k) {
    (void)total;
    return rpn_carve(batch, num_levels, pre_k, nullptr, 0, nullptr) + 256;
}

int launch_rpn_filter(const float* objectness, const float* deltas, const float* anchors, const float* proposals, int batch,
                      int total, const int* level_sizes_host, int num_levels, const float* image_hw,
                      int pre_k, int post_k, double nms_thr, float score_thr, float min_size, int nms_mode,
                      float* out_boxes, float* out_scores, int* out_index, int* out_count,
                      void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    RpnParams P{};
    P.obj = objectness; P.deltas = deltas; P.anchors = anchors; P.proposals = proposals; P.image_hw = image_hw;
    P.B = batch; P.total = total; P.L = num_levels; P.pre_k = pre_k; P.post_k = post_k;
    P.score_thr = score_thr; P.min_size = min_size;
    int off = 0, koff = 0, kmax = 0;
    for (int l = 0; l < num_levels; ++l) {
        const int n = level_sizes_host[l];
        if (n < 1) return B200_ERR_INVALID;
        P.level_off[l] = off; P.level_n[l] = n;
        P.level_k[l] = n < pre_k ? n : pre_k;
        P.level_koff[l] = koff;
        off += n; koff += P.level_k[l];
        kmax = kmax > P.level_k[l] ? kmax : P.level_k[l];
    }
    if (off != total) return B200_ERR_INVALID;
    P.Ktot = koff;
    if (koff > 16384 || kmax > 8192) return B200_ERR_INVALID;   // shared-memory sort capacity
    RpnWs w;
    if (!rpn_carve(batch, num_levels, pre_k, workspace, workspace_bytes, &w)) return B200_ERR_WORKSPACE;
    P.box = w.box; P.score = w.score; P.label = w.label; P.aidx = w.aidx;
    P.seg_start = w.seg_start; P.seg_count = w.seg_count;
    P.keep = w.keep; P.keep_count = w.keep_count;
    P.out_boxes = out_boxes; P.out_scores = out_scores; P.out_index = out_index; P.out_count = out_count;

    int pp = 1;
    while (pp < kmax) pp <<= 1;
    const size_t sel_smem = sizeof(unsigned long long) * (size_t)pp;
    static SmemOptIn optin1, optin2;
    if (optin1.ensure(k_rpn_select, 8192 * 8) != cudaSuccess) return B200_ERR_CUDA;
    k_rpn_select<<<dim3(num_levels, batch), kSelThreads, sel_smem, stream>>>(P);

    NmsParams np{};
    const size_t T = (size_t)batch * P.Ktot;
    const bool trick = nms_mode == B200_NMS_TV_TRICK;
    const int nseg = batch * num_levels;             // one segment per (image, level) for either strategy
    const int max_seg = kmax;
    if (!nms_carve_scratch(&np, T, (size_t)nseg, (size_t)max_seg, w.nms, w.nms_bytes)) return B200_ERR_WORKSPACE;
    np.boxes = reinterpret_cast<const float*>(w.box); np.scores = w.score; np.labels = w.label;
    np.keep = w.keep; np.labels_out = nullptr; np.keep_count = w.keep_count;
    np.thr_f = (float)nms_thr; np.thr_d = nms_thr;
    np.from_slab = 0;
    np.max_seg = max_seg;
    np.seg_offsets = w.seg_start; np.seg_counts = w.seg_count;
    P.segs_per_img = num_levels;
    if (trick) {
        // same segments, torchvision's shifted-coordinate arithmetic (labels = level, unit of the whole image)
        k_rpn_units<<<batch, 256, 0, stream>>>(P, w.units);
        np.mode = B200_NMS_TV_TRICK;
        np.given_unit = w.units;
    } else {
        np.mode = B200_NMS_TV;                        // one level per segment
    }
    const int rc = launch_nms(np, nseg, stream);
    if (rc != B200_OK) return rc;
    int fp = 1;
    while (fp < P.Ktot) fp <<= 1;
    if (optin2.ensure(k_rpn_finish, 16384 * 8) != cudaSuccess) return B200_ERR_CUDA;
    k_rpn_finish<<<batch, 1024, sizeof(unsigned long long) * (size_t)fp, stream>>>(P);
    return cudaGetLastError() == cudaSuccess ? B200_OK : B200_ERR_CUDA;
}

}  // namespace b200
