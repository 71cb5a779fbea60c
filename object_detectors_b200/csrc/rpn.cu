// rpn.cu -- RPN proposal filter (sm_100a): per-level top-k -> decode -> clip/filter -> NMS -> top-n.
// Replaces BoxCoder.decode_single + RegionProposalNetwork._get_top_n_idx / filter_proposals
// (torchvision_models/tvision/_utils.py:186-223, rpn.py:215-280).
//
//  k_rpn_select_cluster : a CLUSTER of 8 CTAs per (image, level).  Every CTA pulls its eighth of the level's raw
//                 objectness logits into shared memory ONCE (<= 100 KB of order-preserving keys; level 0 of an
//                 800 x 1344 image is 201 600 logits), so HBM is read exactly once; the 8-bit MSB radix select then
//                 runs out of shared memory, the eight local histograms are summed into CTA 0's through distributed
//                 shared memory and every CTA takes the same decision (3 cluster barriers per pass).  The selected
//                 <= k entries are compacted into one global list, sorted by CTA 0 (score desc, index asc ==
//                 Tensor.topk order for tie-free input) and only those are decoded / clipped / filtered -- the
//                 reference decodes all 268 569 anchors first (rpn.py:355).  Survivors are written in order at a
//                 fixed stride.  With `topk_out` the kernel stops after the sort and emits the index list of
//                 RegionProposalNetwork._get_top_n_idx (rpn.py:215-228).
//  k_rpn_select : fallback for levels that do not fit eight shared-memory slices (> 204 800 anchors): one CTA per
//                 (image, level), four histogram passes over global memory.
//  k_rpn_units  : (coordinate-trick strategy) the image's shift unit for its per-level segments (see below).
//  NMS          : the shared per-segment kernel of nms.cu (segments = image x level for the
//                 vanilla strategy, = image for the coordinate trick).
//  k_rpn_finish : merges an image's kept lists by score and emits the first post_nms_top_n.
#include <cooperative_groups.h>

#include "nms.cuh"

namespace cg = cooperative_groups;

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
    long long* topk_out;   // [B, Ktot] or nullptr: emit the sorted top-k anchor indices and stop (_get_top_n_idx)
    // sliced select
    int level_slices[kMaxLevels], level_coff[kMaxLevels];   // slices per level, candidate offset of the level
    int cand_stride;                  // candidates per image = sum_l slices_l * k_l (multi-slice levels only)
    unsigned long long* cand;         // [B, cand_stride] (~key << 32 | index inside the level)
    int* sel_counter;                 // [B * L] slice arrival counters (zeroed before the launch)
    // sampled threshold (levels much larger than k): candidates = everything at least as good as a sample quantile
    int level_mode[kMaxLevels];       // 0: one slice, 1: slices keep their local top k, 2: sampled threshold
    int level_cap[kMaxLevels];        // mode 2: capacity of the level's candidate list
    int level_rank[kMaxLevels];       // mode 2: rank of the threshold inside the sample
    int level_chunks[kMaxLevels];     // mode 2: CTAs of k_rpn_pass (4096 logits each), 0 for the other modes
    unsigned* thr;                    // [B * L] mode 2: threshold in key space (smaller = better)
    int* cand_count;                  // [B * L] mode 2: candidates appended (may exceed the capacity: then fallback)
    // many-class heads (RetinaNet): a level is [anchors_l, C] logits, flattened; C == 1 for the RPN
    int C;                            // classes per anchor
    const float* class_scale;         // [C] or nullptr: logits are multiplied by it before the sigmoid (tfidf_post)
    int strict_thr;                   // 1: score >  score_thr (retinanet.py:437), 0: score >= score_thr (rpn.py:268)
    int label_class;                  // 1: NMS label = class (index % C), 0: NMS label = level
    int level_aoff[kMaxLevels];       // first anchor of the level (= level_off / C)
    int atotal;                       // anchors per image (= total / C): row count of deltas / proposals
    int* out_labels;                  // [B, post_k] or nullptr
    int* img_start; int* img_count;   // [B] image-wide segments (k_rpn_compact)
    long long* prof;                  // debug: [CTAs of the sliced select, 8] globaltimer stamps (b200_debug_set_rpn_prof)
};

__device__ __forceinline__ void rpn_stamp(const RpnParams& P, int slot) {
    if (P.prof && threadIdx.x == 0) {
        long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        P.prof[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 8 + slot] = t;
    }
}

// the logit the selection orders by: raw objectness, or tfidf_post[c] * logit for many-class heads
__device__ __forceinline__ float level_logit(const RpnParams& P, const float* src, int j) {
    const float x = __ldg(src + j);             // (not the volatile streaming load: loops over logits must be free to batch their loads)
    if (!P.class_scale) return x;
    return __fmul_rn(__ldg(P.class_scale + (j - (j / P.C) * P.C)), x);
}

// Ascending bitonic sort of key[0..N) in shared memory, N a power of two; every thread of the CTA calls it.
// Thread t keeps the E = max(1, N / 1024) keys [t*E, t*E + E) in registers: compare-exchange steps with a partner
// distance below E are register moves, below 32*E warp shuffles, and only the remaining steps (15 of the 91 for 8192
// keys) go through shared memory, with the array itself as the exchange buffer (transposed: conflict-free).
template <int E>
__device__ __forceinline__ void bitonic_regs(unsigned long long* key, int N) {
    const int tid = threadIdx.x;
    const int T = N / E;                       // active threads: a multiple of 32 (N >= 64)
    const bool active = tid < T;
    unsigned long long a[E];
#pragma unroll
    for (int e = 0; e < E; ++e) a[e] = active ? key[tid * E + e] : ~0ull;
    __syncthreads();
    for (int k = 2; k <= N; k <<= 1) {
        int j = k >> 1;
        for (; j >= 32 * E; j >>= 1) {         // partner in another warp
            const int m = j / E;
            const bool keep_min = ((tid & m) == 0) == (((tid * E) & k) == 0);
            if (active) {
#pragma unroll
                for (int e = 0; e < E; ++e) key[e * T + tid] = a[e];
            }
            __syncthreads();
            if (active) {
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const unsigned long long o = key[e * T + (tid ^ m)];
                    a[e] = keep_min ? (a[e] < o ? a[e] : o) : (a[e] < o ? o : a[e]);
                }
            }
            __syncthreads();
        }
        if (active) {
            for (; j >= E; j >>= 1) {          // partner in another lane
                const int m = j / E;
                const bool keep_min = ((tid & m) == 0) == (((tid * E) & k) == 0);
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const unsigned long long o = __shfl_xor_sync(kFullMask, a[e], m);
                    a[e] = keep_min ? (a[e] < o ? a[e] : o) : (a[e] < o ? o : a[e]);
                }
            }
#pragma unroll
            for (int jj = E >> 1; jj > 0; jj >>= 1) {      // partner in this thread
                if (jj < k) {
#pragma unroll
                    for (int e = 0; e < E; ++e) {
                        if ((e & jj) == 0) {
                            const bool asc = ((tid * E + e) & k) == 0;
                            const unsigned long long x = a[e], y = a[e | jj];
                            if ((x > y) == asc) { a[e] = y; a[e | jj] = x; }
                        }
                    }
                }
            }
        }
    }
    if (active) {
#pragma unroll
        for (int e = 0; e < E; ++e) key[tid * E + e] = a[e];
    }
    __syncthreads();
}

__device__ __noinline__ void bitonic_sort_u64(unsigned long long* key, int P) {
    const int tid = threadIdx.x;
    if (P < 64) {
        for (int k = 2; k <= P; k <<= 1)
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int t = tid; t < (P >> 1); t += blockDim.x) {
                    const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                    const int ixj = i | j;
                    const unsigned long long a = key[i], b = key[ixj];
                    if ((a > b) == ((i & k) == 0)) { key[i] = b; key[ixj] = a; }
                }
                __syncthreads();
            }
        return;
    }
    if (P <= 1024) bitonic_regs<1>(key, P);
    else if (P == 2048) bitonic_regs<2>(key, P);
    else if (P == 4096) bitonic_regs<4>(key, P);
    else if (P == 8192) bitonic_regs<8>(key, P);
    else bitonic_regs<16>(key, P);
}

// sorted selection -> (optional) index list, else decode / clip / filter in order (s_scan: 2 * kSelWarps ints).  `sel` holds `got` keys
// (~orderable(score) << 32 | index inside the level) in shared memory, with room for Ppad (a power of two >= got).
__device__ void rpn_finish_level(const RpnParams& P, int b, int l, unsigned long long* sel, int got, int Ppad, const float* src,
                                 int* s_scan) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (Ppad) {                                        // Ppad == 0: sel[0..got) already is sorted
        for (int i = got + tid; i < Ppad; i += kSelThreads) sel[i] = ~0ull;
        __syncthreads();
        bitonic_sort_u64(sel, Ppad);
    }
    rpn_stamp(P, 4);
    const size_t out0 = (size_t)b * P.Ktot + P.level_koff[l];
    if (P.topk_out) {
        for (int r = tid; r < got; r += kSelThreads) P.topk_out[out0 + r] = P.level_off[l] + (int)(unsigned)sel[r];
        return;
    }
    // ---- decode / clip / filter the selected anchors, keep order ------------------------------
    // two rows per thread and round (r and r + 1024): their gathers are in flight together
    const float img_h = P.image_hw[2 * b], img_w = P.image_hw[2 * b + 1];
    int running = 0;
    for (int r0 = 0; r0 < got; r0 += 2 * kSelThreads) {
        bool ok[2];
        float4 box[2];
        float prob[2];
        int a[2], cls[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int r = r0 + h * kSelThreads + tid;
            ok[h] = false; box[h] = make_float4(0.f, 0.f, 0.f, 0.f); prob[h] = 0.f; a[h] = 0; cls[h] = 0;
            if (r < got) {
                const int i = (int)(unsigned)sel[r];
                const int an = i / P.C;
                cls[h] = i - an * P.C;
                a[h] = P.level_aoff[l] + an;
                prob[h] = sigmoid_ref(level_logit(P, src, i));                        // rpn.py:255 / retinanet.py:436
                float x1, y1, x2, y2;
                if (P.proposals) {
                    const float4 pb = *reinterpret_cast<const float4*>(P.proposals + ((size_t)b * P.atotal + a[h]) * 4);
                    x1 = pb.x; y1 = pb.y; x2 = pb.z; y2 = pb.w;
                } else {
                    const float4 an4 = *reinterpret_cast<const float4*>(P.anchors + 4 * (size_t)a[h]);
                    const float4 d = *reinterpret_cast<const float4*>(P.deltas + ((size_t)b * P.atotal + a[h]) * 4);
                    // BoxCoder.decode_single, weights (1,1,1,1)                           _utils.py:199-221
                    const float w = __fsub_rn(an4.z, an4.x), hh = __fsub_rn(an4.w, an4.y);
                    const float cx = __fadd_rn(an4.x, __fmul_rn(0.5f, w)), cy = __fadd_rn(an4.y, __fmul_rn(0.5f, hh));
                    const float dw = fminf(d.z, kXformClip), dh = fminf(d.w, kXformClip);
                    const float pcx = __fadd_rn(__fmul_rn(d.x, w), cx), pcy = __fadd_rn(__fmul_rn(d.y, hh), cy);
                    const float pw = __fmul_rn(expf(dw), w), ph = __fmul_rn(expf(dh), hh);
                    x1 = __fsub_rn(pcx, __fmul_rn(0.5f, pw)); y1 = __fsub_rn(pcy, __fmul_rn(0.5f, ph));
                    x2 = __fadd_rn(pcx, __fmul_rn(0.5f, pw)); y2 = __fadd_rn(pcy, __fmul_rn(0.5f, ph));
                }
                // clip_boxes_to_image (rpn.py:260)
                x1 = fminf(fmaxf(x1, 0.f), img_w); x2 = fminf(fmaxf(x2, 0.f), img_w);
                y1 = fminf(fmaxf(y1, 0.f), img_h); y2 = fminf(fmaxf(y2, 0.f), img_h);
                box[h] = make_float4(x1, y1, x2, y2);
                // remove_small_boxes + score threshold (rpn.py:263-269)
                ok[h] = (__fsub_rn(x2, x1) >= P.min_size) && (__fsub_rn(y2, y1) >= P.min_size) &&
                        (P.strict_thr ? prob[h] > P.score_thr : prob[h] >= P.score_thr);
            }
        }
        const unsigned bal0 = __ballot_sync(kFullMask, ok[0]), bal1 = __ballot_sync(kFullMask, ok[1]);
        if (lane == 0) { s_scan[warp] = __popc(bal0); s_scan[kSelWarps + warp] = __popc(bal1); }
        __syncthreads();
        int before0 = 0, total0 = 0, before1 = 0, total1 = 0;
        for (int w2 = 0; w2 < kSelWarps; ++w2) {
            const int c0 = s_scan[w2], c1 = s_scan[kSelWarps + w2];
            if (w2 < warp) { before0 += c0; before1 += c1; }
            total0 += c0; total1 += c1;
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            if (ok[h]) {
                const unsigned bal = h ? bal1 : bal0;
                const size_t o = out0 + running + (h ? total0 + before1 : before0) + __popc(bal & ((1u << lane) - 1u));
                P.box[o] = box[h];
                P.score[o] = prob[h];
                P.label[o] = P.label_class ? cls[h] : l;
                P.aidx[o] = a[h];
            }
        }
        running += total0 + total1;
        __syncthreads();
    }
    if (tid == 0) {
        P.seg_start[b * P.L + l] = (int)out0;
        P.seg_count[b * P.L + l] = running;
    }
}


// ------------------------------------------------------------------------------------------ sliced select
static constexpr int kSliceMax = 20480;       // logits one CTA holds in shared memory (80 KB of 32-bit keys)
static constexpr int kNarrowMax = 3072;       // entries of the threshold bin refined out of a compact list (24 KB)

struct SelShared {
    int hist[256];
    unsigned prefix, mask;
    int remaining;
    int c_out, c_mid, last;
    int bin, above, in_bin;
    int scan[kSelWarps];
};

// one histogram pass: entries of v[0..m) whose masked bits equal `prefix`, binned by byte `shift / 8`.
// Float keys concentrate in a handful of exponent bins: lanes that hit the same bin are merged (match.any) into
// ONE shared-memory atomic instead of up to 32 serialised ones.
__device__ __forceinline__ void hist_pass(const unsigned* v, int m, unsigned prefix, unsigned mask, int shift, int* hist) {
    const int tid = threadIdx.x, lane = tid & 31;
    for (int i = tid; i - lane < m; i += kSelThreads) {
        const unsigned key = i < m ? v[i] : 0u;
        const bool in = i < m && (key & mask) == prefix;
        const unsigned bin = in ? (key >> shift) & 255u : 256u;
        const unsigned peers = __match_any_sync(kFullMask, bin);
        if (in && lane == __ffs(peers) - 1) atomicAdd(&hist[bin], __popc(peers));
    }
}

// warp 0: the bin (ascending) that holds rank `rem` (1-based) of the histogram; entries before it
__device__ __forceinline__ void find_bin(const int* hist, int rem, int& bin, int& before_bin, int& in_bin) {
    const int lane = threadIdx.x & 31;
    int local[8], sum = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) { local[j] = hist[lane * 8 + j]; sum += local[j]; }
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(kFullMask, incl, o);
        if (lane >= o) incl += u;
    }
    int before = incl - sum, fb = -1, fbefore = 0, fin = 0;
    if (before < rem && rem <= incl) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (fb < 0 && rem <= before + local[j]) { fb = lane * 8 + j; fbefore = before; fin = local[j]; }
            before += local[j];
        }
    }
    const unsigned who = __ballot_sync(kFullMask, fb >= 0);
    const int src = who ? __ffs(who) - 1 : 0;
    bin = __shfl_sync(kFullMask, fb, src);
    before_bin = __shfl_sync(kFullMask, fbefore, src);
    in_bin = __shfl_sync(kFullMask, fin, src);
}

// ------------------------------------------------------------------------------------------ bucket sort
static constexpr int kBuckets = 2048;        // value-linear buckets between the best and the worst candidate
static constexpr int kBucketRun = 1024;      // longest bucket that is still ranked by counting

// The kk = min(k, cnt) smallest of cnt keys (value << 32 | index; value = ~orderable(logit)) in ascending order into
// out[0..kk) -- exactly what a full sort would put there -- without a sorting network: a histogram over kBuckets
// buckets that are linear in the LOGIT (a monotone map: the order of the buckets is the order of the keys), a scan, a
// scatter of the buckets up to the one that holds rank kk, and a rank-by-counting inside each (short) bucket.
// Eight barriers instead of the ~90 steps of a bitonic network.  Returns kk, or -1 when the distribution defeats
// the buckets (non-finite or all-equal values, a bucket longer than kBucketRun, more than tmp_cap entries up to the
// threshold bucket): the caller then sorts the slow way.  `out` may alias the storage `load` reads from.
template <typename Load>
__device__ int bucket_sort_topk(Load load, int cnt, int k, unsigned long long* tmp, int tmp_cap, unsigned long long* out,
                                int* start, int* fill, SelShared& S) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int kk = min(k, cnt);
    unsigned vmin = ~0u, vmax = 0u;
    for (int i = tid; i < cnt; i += kSelThreads) {
        const unsigned v = (unsigned)(load(i) >> 32);
        vmin = min(vmin, v); vmax = max(vmax, v);
    }
    vmin = __reduce_min_sync(kFullMask, vmin);
    vmax = __reduce_max_sync(kFullMask, vmax);
    if (lane == 0) { S.scan[warp] = (int)vmin; S.hist[warp] = (int)vmax; }
    for (int i = tid; i < kBuckets; i += kSelThreads) fill[i] = 0;
    if (tid == 0) { S.bin = -1; S.c_out = 0; }
    __syncthreads();
    for (int w = 0; w < kSelWarps; ++w) { vmin = min(vmin, (unsigned)S.scan[w]); vmax = max(vmax, (unsigned)S.hist[w]); }
    const float xbest = from_orderable(~vmin), xworst = from_orderable(~vmax);
    const float scale = __fdiv_rn((float)(kBuckets - 1), __fsub_rn(xbest, xworst));
    if (!(xbest > xworst) || !isfinite(xbest) || !isfinite(xworst) || !isfinite(scale)) return -1;
    auto bucket_of = [&](unsigned v) {
        const float t = __fmul_rn(__fsub_rn(xbest, from_orderable(~v)), scale);       // >= 0, grows as the logit falls
        return min(kBuckets - 1, (int)t);
    };
    __syncthreads();                                                                   // S.scan / S.hist are reused below
    for (int i = tid; i < cnt; i += kSelThreads) atomicAdd(&fill[bucket_of((unsigned)(load(i) >> 32))], 1);
    __syncthreads();
    // exclusive scan of the 2048 counts, two per thread
    {
        const int c0 = fill[2 * tid], c1 = fill[2 * tid + 1];
        int incl = c0 + c1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(kFullMask, incl, o);
            if (lane >= o) incl += u;
        }
        if (lane == 31) S.scan[warp] = incl;
        __syncthreads();
        int before = 0;
        for (int w = 0; w < warp; ++w) before += S.scan[w];
        const int e0 = before + incl - (c0 + c1), e1 = e0 + c0, e2 = e1 + c1;
        start[2 * tid] = e0; start[2 * tid + 1] = e1;
        if (tid == kSelThreads - 1) start[kBuckets] = e2;
        fill[2 * tid] = 0; fill[2 * tid + 1] = 0;
        // the bucket that holds rank kk (1-based); longer-than-allowed buckets at or before it
        if (c0 > 0 && e0 < kk && kk <= e1) S.bin = 2 * tid;
        if (c1 > 0 && e1 < kk && kk <= e2) S.bin = 2 * tid + 1;
        if ((c0 > kBucketRun && e0 < kk) || (c1 > kBucketRun && e1 < kk)) S.c_out = 1;
    }
    __syncthreads();
    const int bk = S.bin;
    if (bk < 0 || S.c_out) return -1;
    const int m = start[bk + 1];
    if (m > tmp_cap) return -1;
    for (int i = tid; i < cnt; i += kSelThreads) {
        const unsigned long long key = load(i);
        const int b = bucket_of((unsigned)(key >> 32));
        if (b <= bk) tmp[start[b] + atomicAdd(&fill[b], 1)] = key;
    }
    __syncthreads();
    for (int p = tid; p < m; p += kSelThreads) {
        const unsigned long long key = tmp[p];
        const int b = bucket_of((unsigned)(key >> 32));
        const int s0 = start[b], s1 = start[b + 1];
        int r = 0;
        for (int q = s0; q < s1; ++q) r += tmp[q] < key;
        if (s0 + r < kk) out[s0 + r] = key;
    }
    __syncthreads();
    return kk;
}

// Block-wide MSB radix select over v[0..m) in shared memory: the threshold T and the number of entries equal to T
// that belong to the k SMALLEST values (smaller value = larger score: v = ~orderable(logit)).  m > k >= 1.
// `prefix0` / `mask0` / `first_pass`: bytes already decided by the caller.
__device__ void pick_smallest(const unsigned* v, int m, int k, SelShared& S, unsigned prefix0, unsigned mask0, int first_pass,
                              unsigned& T, int& take_eq) {
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) { S.prefix = prefix0; S.mask = mask0; S.remaining = k; }
    __syncthreads();
    for (int pass = first_pass; pass >= 0; --pass) {
        if (tid < 256) S.hist[tid] = 0;
        __syncthreads();
        const unsigned prefix = S.prefix, mask = S.mask;
        const int shift = 8 * pass;
        hist_pass(v, m, prefix, mask, shift, S.hist);
        __syncthreads();
        if (warp == 0) {
            int bin, before, in_bin;
            find_bin(S.hist, S.remaining, bin, before, in_bin);
            if (tid == 0) {
                S.prefix = prefix | ((unsigned)bin << shift);
                S.mask = mask | (255u << shift);
                S.remaining -= before;
            }
        }
        __syncthreads();
    }
    T = S.prefix;
    take_eq = S.remaining;
}

// The min(k, cnt) smallest entries of vals[0..cnt) (shared memory) as (value << 32 | idx_of(i)) into out[] (any
// order; equal values: lowest i first).  One histogram pass over everything finds the byte-3 bin of the k-th value;
// one compaction pass sends the entries of better bins straight to out[] and the (few) entries of that bin to a
// compact list, where the remaining three bytes are decided.  A bin too crowded for the list falls back to
// histogram passes over the whole array.
template <typename IdxOf>
__device__ void select_block(const unsigned* vals, int cnt, int k, unsigned long long* out, IdxOf idx_of, SelShared& S,
                             unsigned* nv, unsigned* ni) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (cnt <= k) {
        for (int i = tid; i < cnt; i += kSelThreads) out[i] = ((unsigned long long)vals[i] << 32) | idx_of(i);
        __syncthreads();
        return;
    }
    if (tid < 256) S.hist[tid] = 0;
    if (tid == 0) { S.c_out = 0; S.c_mid = 0; }
    __syncthreads();
    hist_pass(vals, cnt, 0u, 0u, 24, S.hist);
    __syncthreads();
    if (warp == 0) {
        int bin, before, in_bin;
        find_bin(S.hist, k, bin, before, in_bin);
        if (tid == 0) { S.bin = bin; S.above = before; S.in_bin = in_bin; }
    }
    __syncthreads();
    const unsigned b1 = (unsigned)S.bin;
    const int above = S.above, in_bin = S.in_bin, need = k - above;      // 1 <= need <= in_bin
    const bool narrow = in_bin <= kNarrowMax;
    // warp-aggregated compaction: entries of better bins -> out[0..above), the threshold bin -> the compact list
    for (int i = tid; i - lane < cnt; i += kSelThreads) {
        const unsigned v = i < cnt ? vals[i] : ~0u;
        const bool hi = i < cnt && (v >> 24) < b1, mid = narrow && i < cnt && (v >> 24) == b1;
        const unsigned bh = __ballot_sync(kFullMask, hi), bm = __ballot_sync(kFullMask, mid);
        int base_h = 0, base_m = 0;
        if (lane == 0) {
            if (bh) base_h = atomicAdd(&S.c_out, __popc(bh));
            if (bm) base_m = atomicAdd(&S.c_mid, __popc(bm));
        }
        base_h = __shfl_sync(kFullMask, base_h, 0);
        base_m = __shfl_sync(kFullMask, base_m, 0);
        const unsigned lt = (1u << lane) - 1u;
        if (hi) out[base_h + __popc(bh & lt)] = ((unsigned long long)v << 32) | idx_of(i);
        if (mid) { const int p = base_m + __popc(bm & lt); nv[p] = v; ni[p] = (unsigned)i; }     // ascending i inside a warp
    }
    __syncthreads();
    unsigned T;
    int take_eq;
    const unsigned* rv = narrow ? nv : vals;
    const int rn = narrow ? in_bin : cnt;
    if (narrow && in_bin == need) {
        T = ~0u; take_eq = 0;                                   // the whole bin is selected
        for (int p = tid; p < in_bin; p += kSelThreads) out[above + p] = ((unsigned long long)nv[p] << 32) | idx_of((int)ni[p]);
        __syncthreads();
        return;
    }
    pick_smallest(rv, rn, narrow ? need : k, S, narrow ? (b1 << 24) : 0u, narrow ? 0xff000000u : 0u, narrow ? 2 : 3, T, take_eq);
    // entries below T inside the threshold bin, then the first take_eq entries equal to T in index order
    if (tid == 0) S.c_mid = 0;
    __syncthreads();
    int lt_cnt = 0;
    for (int p = tid; p < rn; p += kSelThreads) { const unsigned v = rv[p]; lt_cnt += (v >> 24) == b1 && v < T; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lt_cnt += __shfl_xor_sync(kFullMask, lt_cnt, o);
    if (lane == 0 && lt_cnt) atomicAdd(&S.c_mid, lt_cnt);
    __syncthreads();
    const int lt_total = S.c_mid;                               // == need - take_eq
    __syncthreads();
    if (tid == 0) S.c_mid = 0;
    __syncthreads();
    for (int p = tid; p - lane < rn; p += kSelThreads) {
        const unsigned v = p < rn ? rv[p] : ~0u;
        const bool in = p < rn && (v >> 24) == b1 && v < T;
        const unsigned bal = __ballot_sync(kFullMask, in);
        int base = 0;
        if (lane == 0 && bal) base = atomicAdd(&S.c_mid, __popc(bal));
        base = __shfl_sync(kFullMask, base, 0);
        if (in) out[above + base + __popc(bal & ((1u << lane) - 1u))] = ((unsigned long long)v << 32) | idx_of(narrow ? (int)ni[p] : p);
    }
    // ties at the threshold.  Usually every entry equal to T is taken (tie-free input: T itself): append them all.
    if (tid == 0) S.c_out = 0;
    __syncthreads();
    int eq_cnt = 0;
    for (int p = tid; p < rn; p += kSelThreads) eq_cnt += rv[p] == T;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) eq_cnt += __shfl_xor_sync(kFullMask, eq_cnt, o);
    if (lane == 0 && eq_cnt) atomicAdd(&S.c_out, eq_cnt);
    __syncthreads();
    const int eq_total = S.c_out;
    __syncthreads();
    if (eq_total <= take_eq) {
        if (tid == 0) S.c_out = 0;
        __syncthreads();
        for (int p = tid; p - lane < rn; p += kSelThreads) {
            const bool in = p < rn && rv[p] == T;
            const unsigned bal = __ballot_sync(kFullMask, in);
            int base = 0;
            if (lane == 0 && bal) base = atomicAdd(&S.c_out, __popc(bal));
            base = __shfl_sync(kFullMask, base, 0);
            if (in) out[above + lt_total + base + __popc(bal & ((1u << lane) - 1u))] = ((unsigned long long)T << 32) | idx_of(narrow ? (int)ni[p] : p);
        }
    } else if (warp == 0) {
        // more equal entries than places: the lowest indices win.  The compact list is ascending in i only inside
        // each warp's chunk, so equal entries are ranked by index explicitly.
        int taken = 0;
        for (int base = 0; base < rn && taken < take_eq; base += 32) {
            const int p = base + lane;
            const bool is = p < rn && rv[p] == T;
            const unsigned bal = __ballot_sync(kFullMask, is);
            if (!narrow) {
                const int pos = taken + __popc(bal & ((1u << lane) - 1u));
                if (is && pos < take_eq) out[above + lt_total + pos] = ((unsigned long long)T << 32) | idx_of(p);
            } else if (is) {
                const unsigned mine = ni[p];
                int rank = 0;
                for (int q = 0; q < rn; ++q) rank += rv[q] == T && ni[q] < mine;
                if (rank < take_eq) out[above + lt_total + rank] = ((unsigned long long)T << 32) | idx_of((int)mine);
            }
            taken += narrow ? 0 : __popc(bal);
        }
    }
    __syncthreads();
}

__device__ int select_level_global(const RpnParams& P, int b, int l, unsigned long long* sel);

static constexpr int kSample = 4096;          // logits sampled per level for the threshold estimate

// Levels much larger than k (mode 2): a sample of kSample evenly spaced logits estimates a threshold that keeps about
// twice the wanted count (the order statistic `level_rank` of the sample).  ONE streaming pass over the level
// (k_rpn_pass: no shared-memory slice, no histogram) then yields a candidate list of a few thousand entries that surely
// holds the top k, and the level's CTA of the select kernel just sorts that list.  "Surely" is statistics (k sits 6-7
// standard deviations below the expected count), so it is verified: a level whose list ends up shorter than k or
// longer than its capacity is redone by select_level_global.
// The threshold itself comes from the same value-linear buckets as the sort: histogram of the sample, the bucket that
// holds rank `level_rank`, and the worst sampled value up to that bucket (an order statistic of rank >= level_rank).
// The kernel also zeroes the level's arrival and candidate counters.
__global__ void __launch_bounds__(kSelThreads, 1)
k_rpn_sample(const __grid_constant__ RpnParams P) {
    __shared__ unsigned keys[kSample];
    __shared__ int hist[kBuckets];
    __shared__ SelShared S;
    const int l = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { P.sel_counter[b * P.L + l] = 0; P.cand_count[b * P.L + l] = 0; }
    if (P.level_mode[l] != 2) return;
    const int n = P.level_n[l], rank = P.level_rank[l];
    const float* src = P.obj + (size_t)b * P.total + P.level_off[l];
    constexpr int kPer = kSample / kSelThreads;
    unsigned v[kPer], vmin = ~0u, vmax = 0u;
#pragma unroll
    for (int j = 0; j < kPer; ++j) {
        const int at = j * kSelThreads + tid;
        v[j] = ~orderable(level_logit(P, src, (int)((long long)at * n / kSample)));
        keys[at] = v[j];
        vmin = min(vmin, v[j]); vmax = max(vmax, v[j]);
    }
    vmin = __reduce_min_sync(kFullMask, vmin);
    vmax = __reduce_max_sync(kFullMask, vmax);
    if (lane == 0) { S.scan[warp] = (int)vmin; S.hist[warp] = (int)vmax; }
    hist[2 * tid] = 0; hist[2 * tid + 1] = 0;
    if (tid == 0) S.bin = -1;
    __syncthreads();
    for (int w = 0; w < kSelWarps; ++w) { vmin = min(vmin, (unsigned)S.scan[w]); vmax = max(vmax, (unsigned)S.hist[w]); }
    const float xbest = from_orderable(~vmin), xworst = from_orderable(~vmax);
    const float scale = __fdiv_rn((float)(kBuckets - 1), __fsub_rn(xbest, xworst));
    if (!(xbest > xworst) || !isfinite(xbest) || !isfinite(xworst) || !isfinite(scale)) {
        unsigned T;
        int take_eq;
        pick_smallest(keys, kSample, rank, S, 0u, 0u, 3, T, take_eq);      // ties / non-finite logits: exact order statistic
        if (tid == 0) P.thr[b * P.L + l] = T;
        return;
    }
    int bk[kPer];
#pragma unroll
    for (int j = 0; j < kPer; ++j) {
        bk[j] = min(kBuckets - 1, (int)__fmul_rn(__fsub_rn(xbest, from_orderable(~v[j])), scale));
        atomicAdd(&hist[bk[j]], 1);
    }
    __syncthreads();
    const int c0 = hist[2 * tid], c1 = hist[2 * tid + 1];
    int incl = c0 + c1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(kFullMask, incl, o);
        if (lane >= o) incl += u;
    }
    if (lane == 31) S.scan[warp] = incl;
    __syncthreads();
    int before = 0;
    for (int w = 0; w < warp; ++w) before += S.scan[w];
    const int e0 = before + incl - (c0 + c1), e1 = e0 + c0, e2 = e1 + c1;
    if (c0 > 0 && e0 < rank && rank <= e1) S.bin = 2 * tid;
    if (c1 > 0 && e1 < rank && rank <= e2) S.bin = 2 * tid + 1;
    __syncthreads();
    const int cut = S.bin;
    unsigned worst = 0u;
#pragma unroll
    for (int j = 0; j < kPer; ++j) if (bk[j] <= cut) worst = max(worst, v[j]);
    worst = __reduce_max_sync(kFullMask, worst);
    if (lane == 0) S.hist[warp] = (int)worst;
    __syncthreads();
    if (tid == 0) {
        for (int w = 0; w < kSelWarps; ++w) worst = max(worst, (unsigned)S.hist[w]);
        P.thr[b * P.L + l] = worst;
    }
}

static constexpr int kPassThreads = 256, kPassPer = 16, kPassChunk = kPassThreads * kPassPer;

// mode-2 levels: append every logit at least as good as the level's sampled threshold to the level's candidate list
// as (~key << 32 | index).  A CTA takes 4096 logits, 16 independent loads per thread, compacts inside the CTA and
// reserves its run of the list with ONE global atomic.
__global__ void __launch_bounds__(kPassThreads)
k_rpn_pass(const __grid_constant__ RpnParams P) {
    __shared__ int s_warp[kPassThreads / 32];
    __shared__ int s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, b = blockIdx.y;
    int l = 0, c = blockIdx.x;
    while (c >= P.level_chunks[l]) { c -= P.level_chunks[l]; ++l; }
    const int n = P.level_n[l];
    const float* src = P.obj + (size_t)b * P.total + P.level_off[l];
    const unsigned T0 = P.thr[b * P.L + l];
    const int i0 = c * kPassChunk;
    // all 16 loads first (clamped index, no branch between them), then the arithmetic: one memory round trip per CTA
    float x[kPassPer];
#pragma unroll
    for (int j = 0; j < kPassPer; ++j) x[j] = __ldg(src + min(i0 + j * kPassThreads + tid, n - 1));
    if (P.class_scale) {
#pragma unroll
        for (int j = 0; j < kPassPer; ++j) {
            const int i = min(i0 + j * kPassThreads + tid, n - 1);
            x[j] = __fmul_rn(__ldg(P.class_scale + (i - (i / P.C) * P.C)), x[j]);
        }
    }
    unsigned v[kPassPer];
#pragma unroll
    for (int j = 0; j < kPassPer; ++j) v[j] = ~orderable(x[j]);
    unsigned hit = 0;
#pragma unroll
    for (int j = 0; j < kPassPer; ++j) hit |= (unsigned)(i0 + j * kPassThreads + tid < n && v[j] <= T0) << j;
    const int mine = __popc(hit);
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(kFullMask, incl, o);
        if (lane >= o) incl += u;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (tid == 0) {
        int tot = 0;
#pragma unroll
        for (int w = 0; w < kPassThreads / 32; ++w) { const int t = s_warp[w]; s_warp[w] = tot; tot += t; }
        s_base = tot ? atomicAdd(P.cand_count + b * P.L + l, tot) : 0;
    }
    __syncthreads();
    if (!hit) return;
    int at = s_base + s_warp[warp] + incl - mine;
    const int capL = P.level_cap[l];
    unsigned long long* cand = P.cand + ((size_t)b * P.cand_stride + P.level_coff[l]);
#pragma unroll
    for (int j = 0; j < kPassPer; ++j)
        if (hit >> j & 1u) {
            if (at < capL) cand[at] = ((unsigned long long)v[j] << 32) | (unsigned)(i0 + j * kPassThreads + tid);
            ++at;
        }
}

// Level l of image b is cut into R_l = ceil(n_l / kSliceMax) slices, one CTA each (sampled levels: one CTA that
// sorts the candidates k_rpn_pass left).  A CTA reads its slice ONCE
// (order-preserving keys into shared memory), selects its local top min(k, m) there -- a superset of its share of
// the level's top k -- and appends them to the level's candidate list.  The CTA that arrives last at the level's
// counter (nobody waits) selects the top k among the R_l * k candidates, sorts them and runs the level's decode /
// clip / filter (or emits the index list).  Single-slice levels skip the candidate round trip.
__global__ void __launch_bounds__(kSelThreads, 1)
k_rpn_select_sliced(const __grid_constant__ RpnParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned* nv = reinterpret_cast<unsigned*>(smem_raw);                       // [kNarrowMax] compact list: values
    unsigned* ni = nv + kNarrowMax;                                             // [kNarrowMax]               indices
    unsigned* keys = ni + kNarrowMax;                                           // slice keys / candidate keys
    unsigned char* after = reinterpret_cast<unsigned char*>(keys);
    __shared__ SelShared S;
    const int tid = threadIdx.x;
    const int b = blockIdx.y;
    int l = 0, r = blockIdx.x;
    while (r >= P.level_slices[l]) { r -= P.level_slices[l]; ++l; }
    const int R = P.level_slices[l];
    const int n = P.level_n[l], k = P.level_k[l];
    const float* src = P.obj + (size_t)b * P.total + P.level_off[l];
    const int per = (((n + R - 1) / R) + 3) & ~3;
    const int i0 = min(n, r * per), m = min(n, i0 + per) - i0;

    int got;
    unsigned long long* sel;
    bool presorted = false, bucket_scratch = true;
    int* bstart = reinterpret_cast<int*>(nv);                                   // bucket sort: [kBuckets + 1] and [kBuckets]
    int* bfill = bstart + kBuckets + 4;
    rpn_stamp(P, 0);
    if (P.prof && tid == 0) { P.prof[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 8 + 6] = l; P.prof[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 8 + 7] = r; }
    if (P.level_mode[l] == 2) {
        // ---- sampled threshold: k_rpn_pass left the level's candidates; sort them ------------------------------------
        const int capL = P.level_cap[l];
        const unsigned long long* cand = P.cand + ((size_t)b * P.cand_stride + P.level_coff[l]);
        const int cnt = P.cand_count[b * P.L + l];
        unsigned long long* tmp = reinterpret_cast<unsigned long long*>(after);
        const int tmp_cap = k + 1024;
        sel = tmp + tmp_cap;
        got = -1;
        if (cnt >= min(k, n) && cnt <= capL)
            got = bucket_sort_topk([&](int i) { return __ldcg(cand + i); }, cnt, k, tmp, tmp_cap, sel, bstart, bfill, S);
        if (got >= 0) {
            presorted = true;
        } else {
            // the sample misjudged the level (too few candidates, or more than the list holds), or the values defeat
            // the buckets (ties, non-finite logits): robust path
            sel = tmp;
            got = select_level_global(P, b, l, sel);
            bucket_scratch = false;                    // the selection sits where the bucket sort would scatter
        }
    } else {
    // ---- the one read of the logits ---------------------------------------------------------------------------------
#pragma unroll 4
    for (int i = tid; i < m; i += kSelThreads) keys[i] = ~orderable(level_logit(P, src, i0 + i));
    __syncthreads();
    rpn_stamp(P, 2);

    got = -1;
    if (R == 1 && m >= 256) {
        // the slice is the level: bucket-sort its top k straight out of the keys (select and sort in one go)
        unsigned long long* tmp = reinterpret_cast<unsigned long long*>(after + (((size_t)m * 4 + 15) & ~(size_t)15));
        const int tmp_cap = k + 1024;
        sel = tmp + tmp_cap;
        got = bucket_sort_topk([&](int i) { return ((unsigned long long)keys[i] << 32) | (unsigned)(i0 + i); }, m, k, tmp, tmp_cap,
                               sel, bstart, bfill, S);
        presorted = got >= 0;
    }
    if (presorted) {
        // nothing left to select
    } else if (R == 1) {
        // (values that defeat the buckets, tiny levels) select into the sort buffer behind the keys
        const size_t room = max(((size_t)m * 4 + 15) & ~(size_t)15, ((size_t)k * 8 + 15) & ~(size_t)15);
        sel = reinterpret_cast<unsigned long long*>(after + room);                // leaves k keys of scratch in front
        select_block(keys, m, k, sel, [&](int i) { return (unsigned)(i0 + i); }, S, nv, ni);
        got = min(k, m);
    } else {
        unsigned long long* cand = P.cand + ((size_t)b * P.cand_stride + P.level_coff[l]) + (size_t)r * k;
        select_block(keys, m, k, cand, [&](int i) { return (unsigned)(i0 + i); }, S, nv, ni);
        __threadfence();
        __syncthreads();
        if (tid == 0) S.last = atomicAdd(P.sel_counter + b * P.L + l, 1) == R - 1;
        __syncthreads();
        if (!S.last) return;
        __threadfence();
        // ---- last CTA of the level: top k of the R * k candidates (slice r' holds min(k, m_r') of them) ----------
        const unsigned long long* all_cand = P.cand + ((size_t)b * P.cand_stride + P.level_coff[l]);
        int total = 0;
        for (int rr = 0; rr < R; ++rr) {
            const int j0 = min(n, rr * per), mm = min(n, j0 + per) - j0;
            const int c = min(k, mm);
            for (int i = tid; i < c; i += kSelThreads) keys[total + i] = (unsigned)(__ldcg(all_cand + (size_t)rr * k + i) >> 32);
            total += c;
        }
        __syncthreads();
        sel = reinterpret_cast<unsigned long long*>(after + (((size_t)total * 4 + 15) & ~(size_t)15));
        // candidate i of the concatenation lives in slice i / c_full; the slices are equally long except the last
        const int c_full = min(k, per);
        select_block(keys, total, k, sel, [&](int i) {
            const int rr = min(i / c_full, R - 1);
            return (unsigned)__ldcg(all_cand + (size_t)rr * k + (i - rr * c_full));
        }, S, nv, ni);
        got = min(k, total);
    }
    }
    rpn_stamp(P, 3);
    if (!presorted && bucket_scratch && got >= 256) {
        // the slice keys / candidate keys are dead: scratch for the bucket sort of the selection
        unsigned long long* picked = sel;
        presorted = bucket_sort_topk([&](int i) { return picked[i]; }, got, got, reinterpret_cast<unsigned long long*>(after), got,
                                     sel, bstart, bfill, S) >= 0;
    }
    int Ppad = 1;
    while (Ppad < got) Ppad <<= 1;
    rpn_finish_level(P, b, l, sel, got, presorted ? 0 : Ppad, src, S.hist);
    rpn_stamp(P, 5);
}

// Robust single-CTA select over the level in GLOBAL memory (four histogram passes + one collection pass): the general
// fallback -- levels whose shared-memory plan does not fit, and sampled levels whose sample misjudged the threshold.
// Fills sel[0 .. got) with (~key << 32 | index inside the level) and returns got = min(k, n).
__device__ int select_level_global(const RpnParams& P, int b, int l, unsigned long long* sel) {
    __shared__ int hist[256];
    __shared__ unsigned s_prefix, s_mask;
    __shared__ int s_remaining, s_cnt, s_tie;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = P.level_n[l], k = P.level_k[l];
    const float* src = P.obj + (size_t)b * P.total + P.level_off[l];
    // ---- radix select: key of the k-th largest logit ------------------------------------------
    if (tid == 0) { s_prefix = 0u; s_mask = 0u; s_remaining = k; s_cnt = 0; s_tie = 0; }
    __syncthreads();
    if (n > k) {
        for (int pass = 3; pass >= 0; --pass) {
            for (int i = tid; i < 256; i += kSelThreads) hist[i] = 0;
            __syncthreads();
            const unsigned prefix = s_prefix, mask = s_mask;
            const int shift = 8 * pass;
            for (int i = tid; i - lane < n; i += kSelThreads) {
                const unsigned key = i < n ? orderable(level_logit(P, src, i)) : 0u;
                const bool in = i < n && (key & mask) == prefix;
                const unsigned bin = in ? (key >> shift) & 255u : 256u;
                const unsigned peers = __match_any_sync(kFullMask, bin);          // one atomic per distinct bin and warp
                if (in && lane == __ffs(peers) - 1) atomicAdd(&hist[bin], __popc(peers));
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
            const unsigned key = orderable(level_logit(P, src, i));
            bool take = n <= k || key > T;
            if (!take && key == T) take = atomicAdd(&s_tie, 1) < take_ties;
            if (take) sel[atomicAdd(&s_cnt, 1)] = ((unsigned long long)(~key) << 32) | (unsigned)i;
        }
    }
    __syncthreads();
    return s_cnt;  // == min(k, n)
}

__global__ void __launch_bounds__(kSelThreads, 1)
k_rpn_select(const __grid_constant__ RpnParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long* sel = reinterpret_cast<unsigned long long*>(smem_raw);  // [pow2(k)]
    __shared__ int s_scan[2 * kSelWarps];
    const int l = blockIdx.x, b = blockIdx.y;
    const float* src = P.obj + (size_t)b * P.total + P.level_off[l];
    const int got = select_level_global(P, b, l, sel);
    int Ppad = 1;
    while (Ppad < got) Ppad <<= 1;
    rpn_finish_level(P, b, l, sel, got, Ppad, src, s_scan);
}

// merge the kept lists of an image by score, emit the first post_k (rpn.py:272-278).  Every segment's kept list
// already is in descending score (the NMS emits in that order), so an element's place in the merged order is its own
// position plus, for every other segment, the number of that segment's elements in front of it (binary search):
// no sort.  kFinishParts CTAs per image: each stages all keys (independent, unrolled gathers) and places its share.
static constexpr int kFinishParts = 4;

__global__ void __launch_bounds__(1024, 1)
k_rpn_finish(const __grid_constant__ RpnParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long* key = reinterpret_cast<unsigned long long*>(smem_raw);     // [sum of kept] grouped by segment
    __shared__ int s_at[kMaxLevels + 1], s_st[kMaxLevels];
    const int b = blockIdx.x, part = blockIdx.y, tid = threadIdx.x;
    const size_t base = (size_t)b * P.Ktot;
    const int S = P.segs_per_img;
    if (tid < S) { s_st[tid] = P.seg_start[b * S + tid]; s_at[tid + 1] = P.keep_count[b * S + tid]; }   // one round trip
    __syncthreads();
    if (tid == 0) {
        int at = 0;
        for (int s = 0; s < S; ++s) { const int c = s_at[s + 1]; s_at[s] = at; at += c; }
        s_at[S] = at;
    }
    __syncthreads();
    const int total = s_at[S];
#pragma unroll 4
    for (int e = tid; e < total; e += 1024) {
        int s = 0;
        while (e >= s_at[s + 1]) ++s;
        const int st = s_st[s];
        const int pos = st + (int)P.keep[st + e - s_at[s]];     // absolute row
        key[e] = ((unsigned long long)(~orderable(P.score[pos])) << 32) | (unsigned)(pos - (int)base);
    }
    __syncthreads();
    const int nout = min(total, P.post_k);
    for (int e = part * 1024 + tid; e < total; e += 1024 * kFinishParts) {
        int s = 0;
        while (e >= s_at[s + 1]) ++s;
        const unsigned long long mine = key[e];
        int rank = e - s_at[s];
        for (int o = 0; o < S; ++o) {
            if (o == s) continue;
            int lo = s_at[o], hi = s_at[o + 1];               // first element of segment o that is not in front of mine
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (key[mid] < mine) lo = mid + 1; else hi = mid;
            }
            rank += lo - s_at[o];
        }
        if (rank < nout) {
            const size_t pos = base + (unsigned)mine;
            reinterpret_cast<float4*>(P.out_boxes)[(size_t)b * P.post_k + rank] = P.box[pos];
            P.out_scores[(size_t)b * P.post_k + rank] = from_orderable(~(unsigned)(mine >> 32));
            if (P.out_index) P.out_index[(size_t)b * P.post_k + rank] = P.aidx[pos];
            if (P.out_labels) P.out_labels[(size_t)b * P.post_k + rank] = P.label[pos];
        }
    }
    if (tid == 0 && part == 0) P.out_count[b] = nout;
}

// image-wide segments (class-aware NMS across the levels): the per-level survivor lists of an image, written at a
// fixed stride, are moved together so that the image is one contiguous segment
__global__ void __launch_bounds__(1024, 1)
k_rpn_compact(const __grid_constant__ RpnParams P) {
    const int b = blockIdx.x, tid = threadIdx.x;
    const size_t base = (size_t)b * P.Ktot;
    int at = P.seg_count[b * P.L];
    for (int l = 1; l < P.L; ++l) {
        const int n = P.seg_count[b * P.L + l];
        const size_t from = base + P.level_koff[l];
        if ((size_t)at != (size_t)P.level_koff[l]) {
            // destination precedes the source; chunks of 1024 rows are read completely before they are written
            for (int i0 = 0; i0 < n; i0 += 1024) {
                const int i = i0 + tid;
                float4 bx; float sc = 0.f; int lb = 0, ai = 0;
                if (i < n) { bx = P.box[from + i]; sc = P.score[from + i]; lb = P.label[from + i]; ai = P.aidx[from + i]; }
                __syncthreads();
                if (i < n) { P.box[base + at + i] = bx; P.score[base + at + i] = sc; P.label[base + at + i] = lb; P.aidx[base + at + i] = ai; }
                __syncthreads();
            }
        }
        at += n;
    }
    if (tid == 0) { P.img_start[b] = (int)base; P.img_count[b] = at; }
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
static size_t rpn_select_ws_bytes(int batch, int total, int num_levels, int pre_k);
namespace {
struct RpnWs {
    float4* box; float* score; int* label; int* aidx;
    int* seg_start; int* seg_count;
    long long* keep; int* keep_count; float* units;
    void* nms; size_t nms_bytes;
    void* sel; size_t sel_bytes;
    int* img_start; int* img_count;
};
// NMS scratch: B*L segments of <= pre_k boxes for either batched_nms strategy
size_t rpn_nms_bytes(int batch, int levels, int pre_k) {
    const size_t T = (size_t)batch * levels * pre_k;
    const size_t per_level = nms_scratch_bytes(T, (size_t)batch * levels, (size_t)pre_k);
    // many-class heads run ONE segment of up to levels * pre_k boxes per image (bounded by the 16384-row sort capacity)
    const size_t ktot = (size_t)levels * pre_k < 16384 ? (size_t)levels * pre_k : 16384;
    const size_t per_image = nms_scratch_bytes(T, (size_t)batch, ktot);
    return per_level > per_image ? per_level : per_image;
}
size_t rpn_carve(int batch, int total, int levels, int pre_k, void* base, size_t bytes, RpnWs* w) {
    const size_t T = (size_t)batch * levels * pre_k;   // worst case rows (>= B*Ktot)
    unsigned char* p = reinterpret_cast<unsigned char*>(base);
    size_t used = 0;
    auto take = [&](size_t b) { b = align_up(b, 256); void* r = base ? p + used : nullptr; used += b; return r; };
    RpnWs t;
    t.box = (float4*)take(16 * T); t.score = (float*)take(4 * T); t.label = (int*)take(4 * T); t.aidx = (int*)take(4 * T);
    t.seg_start = (int*)take(4 * (size_t)batch * levels); t.seg_count = (int*)take(4 * (size_t)batch * levels);
    t.keep = (long long*)take(8 * T); t.keep_count = (int*)take(4 * (size_t)batch * levels);
    t.units = (float*)take(4 * (size_t)batch * levels);
    t.img_start = (int*)take(4 * (size_t)batch); t.img_count = (int*)take(4 * (size_t)batch);
    t.nms_bytes = rpn_nms_bytes(batch, levels, pre_k);
    t.nms = take(t.nms_bytes);
    t.sel_bytes = rpn_select_ws_bytes(batch, total, levels, pre_k);
    t.sel = take(t.sel_bytes);
    if (base && used > bytes) return 0;
    if (w) *w = t;
    return used;
}
}  // namespace

// per-level top-k (+ decode / filter unless P.topk_out): sliced kernel when the shared-memory plan fits
long long* g_rpn_prof = nullptr;
static int launch_rpn_select(RpnParams& P, int kmax, void* sel_ws, size_t sel_ws_bytes, cudaStream_t stream) {
    P.prof = g_rpn_prof;
    int pp = 1;
    while (pp < kmax) pp <<= 1;
    static SmemOptIn optin1, optin3;
    // shared-memory plan of the sliced kernel: the slice keys, or (last CTA) the candidates' keys + the sort buffer
    int slices = 0, coff = 0, chunks = 0;
    size_t smem = 0;
    bool any_sampled = false;
    for (int l = 0; l < P.L; ++l) {
        const int n = P.level_n[l], k = P.level_k[l];
        int kp = 1;
        while (kp < k) kp <<= 1;
        size_t need;
        P.level_coff[l] = coff;
        P.level_chunks[l] = 0;
        if (n <= kSliceMax) {
            P.level_mode[l] = 0;
            P.level_slices[l] = 1;
            need = (((size_t)n * 4 + 15) & ~(size_t)15);
            need = need > (((size_t)k * 8 + 15) & ~(size_t)15) ? need : (((size_t)k * 8 + 15) & ~(size_t)15);
            need += (size_t)kp * 8;
            const size_t direct = (((size_t)n * 4 + 15) & ~(size_t)15) + 2 * sizeof(unsigned long long) * (size_t)(k + 1024);
            need = need > direct ? need : direct;
        } else if ((long long)n >= 8ll * k && n >= 4 * kSample) {
            // sampled threshold: keep ~2k (+ a margin of 32 sample ranks), list capacity twice the expectation
            P.level_mode[l] = 2;
            const int rank = (int)((2ll * k * kSample + n - 1) / n) + 32;
            P.level_rank[l] = rank < kSample - 1 ? rank : kSample - 1;
            const long long expect = (long long)P.level_rank[l] * n / kSample;
            long long cap = 2 * expect > 4ll * k ? 2 * expect : 4ll * k;
            if (cap > 16384) cap = 16384;
            P.level_cap[l] = (int)cap;
            P.level_slices[l] = 1;
            P.level_chunks[l] = (n + kPassChunk - 1) / kPassChunk;
            chunks += P.level_chunks[l];
            coff += P.level_cap[l];
            need = 2 * sizeof(unsigned long long) * (size_t)(k + 1024);          // bucket-sort scratch + the sorted list
            need = need > (size_t)kp * 8 ? need : (size_t)kp * 8;                  // (the fallback sorts out of the same buffer)
            any_sampled = true;
        } else {
            P.level_mode[l] = 1;
            const int R = (n + kSliceMax - 1) / kSliceMax;
            const int per = (((n + R - 1) / R) + 3) & ~3;
            P.level_slices[l] = R;
            coff += R * k;
            need = (size_t)per * 4;
            const size_t merge = (((size_t)R * (k < per ? k : per) * 4 + 15) & ~(size_t)15) + (size_t)kp * 8;
            need = need > merge ? need : merge;
        }
        smem = smem > need ? smem : need;
        slices += P.level_slices[l];
    }
    smem += 2 * sizeof(unsigned) * (size_t)kNarrowMax;            // the compact refinement list
    P.cand_stride = coff;
    const size_t cand_bytes = align_up(sizeof(unsigned long long) * (size_t)P.B * (size_t)(coff > 0 ? coff : 1), 256);
    const size_t cnt_bytes = align_up(sizeof(int) * 3 * (size_t)P.B * P.L, 256);
    if (smem <= 200 * 1024 && slices <= 65535 && sel_ws && sel_ws_bytes >= cand_bytes + cnt_bytes) {
        P.cand = reinterpret_cast<unsigned long long*>(sel_ws);
        P.sel_counter = reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(sel_ws) + cand_bytes);
        P.cand_count = P.sel_counter + (size_t)P.B * P.L;
        P.thr = reinterpret_cast<unsigned*>(P.cand_count + (size_t)P.B * P.L);
        if (!any_sampled) {
            if (cudaMemsetAsync(P.sel_counter, 0, sizeof(int) * 2 * (size_t)P.B * P.L, stream) != cudaSuccess) return B200_ERR_CUDA;
        } else {                                      // the sample kernel zeroes the counters
            k_rpn_sample<<<dim3(P.L, P.B), kSelThreads, 0, stream>>>(P);
            k_rpn_pass<<<dim3(chunks, P.B), kPassThreads, 0, stream>>>(P);
        }
        if (optin3.ensure(k_rpn_select_sliced, 200 * 1024) != cudaSuccess) return B200_ERR_CUDA;
        k_rpn_select_sliced<<<dim3(slices, P.B), kSelThreads, smem, stream>>>(P);
    } else {
        if (optin1.ensure(k_rpn_select, 8192 * 8) != cudaSuccess) return B200_ERR_CUDA;
        k_rpn_select<<<dim3(P.L, P.B), kSelThreads, sizeof(unsigned long long) * (size_t)pp, stream>>>(P);
    }
    return cudaGetLastError() == cudaSuccess ? B200_OK : B200_ERR_CUDA;
}

// scratch of the sliced select: candidate lists of the multi-slice levels (<= 8 bytes x pre_k per started slice) + counters
static size_t rpn_select_ws_bytes(int batch, int total, int num_levels, int pre_k) {
    const size_t slices = (size_t)(total + kSliceMax - 1) / kSliceMax + (size_t)num_levels;
    const size_t per_image = slices * (size_t)pre_k + 16384 * (size_t)num_levels;      // local-top-k lists or sampled lists
    return align_up(sizeof(unsigned long long) * (size_t)batch * per_image, 256) +
           align_up(sizeof(int) * 3 * (size_t)batch * num_levels, 256) + 256;
}

size_t rpn_workspace_bytes(int batch, int total, int num_levels, int pre_k) {
    return rpn_carve(batch, total, num_levels, pre_k, nullptr, 0, nullptr) + 256;
}

// `classes` == 1: the RPN filter (labels = level, one NMS segment per image and level).  `classes` > 1: the many-class
// head of RetinaNet (retinanet.py:414-472): levels are [anchors_l, classes] logits scaled by `class_scale`, strict
// score threshold, label = class, ONE class-aware NMS segment per image, labels returned.
int launch_rpn_filter_ex(const float* objectness, const float* deltas, const float* anchors, const float* proposals, int batch,
                         int total, const int* level_sizes_host, int num_levels, const float* image_hw,
                         int pre_k, int post_k, double nms_thr, float score_thr, float min_size, int nms_mode,
                         float* out_boxes, float* out_scores, int* out_index, int* out_count,
                         void* workspace, size_t workspace_bytes, cudaStream_t stream, int classes,
                         const float* class_scale, int* out_labels) {
    RpnParams P{};
    P.C = classes; P.atotal = total / classes; P.class_scale = class_scale; P.strict_thr = classes > 1; P.label_class = classes > 1;
    P.out_labels = out_labels;
    P.obj = objectness; P.deltas = deltas; P.anchors = anchors; P.proposals = proposals; P.image_hw = image_hw;
    P.B = batch; P.total = total; P.L = num_levels; P.pre_k = pre_k; P.post_k = post_k;
    P.score_thr = score_thr; P.min_size = min_size;
    int off = 0, koff = 0, kmax = 0;
    for (int l = 0; l < num_levels; ++l) {
        const int n = level_sizes_host[l];
        if (n < 1) return B200_ERR_INVALID;
        if (n % classes) return B200_ERR_INVALID;
        P.level_off[l] = off; P.level_n[l] = n;
        P.level_aoff[l] = off / classes;
        P.level_k[l] = n < pre_k ? n : pre_k;
        P.level_koff[l] = koff;
        off += n; koff += P.level_k[l];
        kmax = kmax > P.level_k[l] ? kmax : P.level_k[l];
    }
    if (off != total) return B200_ERR_INVALID;
    P.Ktot = koff;
    if (koff > 16384 || kmax > 8192) return B200_ERR_INVALID;   // shared-memory sort capacity
    RpnWs w;
    if (!rpn_carve(batch, total, num_levels, pre_k, workspace, workspace_bytes, &w)) return B200_ERR_WORKSPACE;
    P.box = w.box; P.score = w.score; P.label = w.label; P.aidx = w.aidx;
    P.seg_start = w.seg_start; P.seg_count = w.seg_count;
    P.keep = w.keep; P.keep_count = w.keep_count;
    P.img_start = w.img_start; P.img_count = w.img_count;
    P.out_boxes = out_boxes; P.out_scores = out_scores; P.out_index = out_index; P.out_count = out_count;

    const int rcs = launch_rpn_select(P, kmax, w.sel, w.sel_bytes, stream);
    if (rcs != B200_OK) return rcs;

    NmsParams np{};
    const size_t T = (size_t)batch * P.Ktot;
    if (classes > 1) {
        // class-aware NMS over the whole image (batched_nms(image_boxes, image_scores, image_labels), retinanet.py:463)
        k_rpn_compact<<<batch, 1024, 0, stream>>>(P);
        if (!nms_carve_scratch(&np, T, (size_t)batch, (size_t)P.Ktot, w.nms, w.nms_bytes)) return B200_ERR_WORKSPACE;
        np.boxes = reinterpret_cast<const float*>(w.box); np.scores = w.score; np.labels = w.label;
        np.keep = w.keep; np.labels_out = nullptr; np.keep_count = w.keep_count;
        np.thr_f = (float)nms_thr; np.thr_d = nms_thr;
        np.from_slab = 0;
        np.max_seg = P.Ktot;
        np.seg_offsets = w.img_start; np.seg_counts = w.img_count;
        np.mode = nms_mode;
        P.segs_per_img = 1;
        P.seg_start = w.img_start;
        const int rc1 = launch_nms(np, batch, stream);
        if (rc1 != B200_OK) return rc1;
        static SmemOptIn optin4;
        if (optin4.ensure(k_rpn_finish, 16384 * 8) != cudaSuccess) return B200_ERR_CUDA;
        k_rpn_finish<<<dim3(batch, kFinishParts), 1024, sizeof(unsigned long long) * (size_t)(P.Ktot > 0 ? P.Ktot : 1), stream>>>(P);
        return cudaGetLastError() == cudaSuccess ? B200_OK : B200_ERR_CUDA;
    }
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
    static SmemOptIn optin2;
    if (optin2.ensure(k_rpn_finish, 16384 * 8) != cudaSuccess) return B200_ERR_CUDA;
    k_rpn_finish<<<dim3(batch, kFinishParts), 1024, sizeof(unsigned long long) * (size_t)(P.Ktot > 0 ? P.Ktot : 1), stream>>>(P);
    return cudaGetLastError() == cudaSuccess ? B200_OK : B200_ERR_CUDA;
}

int launch_rpn_filter(const float* objectness, const float* deltas, const float* anchors, const float* proposals, int batch,
                      int total, const int* level_sizes_host, int num_levels, const float* image_hw,
                      int pre_k, int post_k, double nms_thr, float score_thr, float min_size, int nms_mode,
                      float* out_boxes, float* out_scores, int* out_index, int* out_count,
                      void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    return launch_rpn_filter_ex(objectness, deltas, anchors, proposals, batch, total, level_sizes_host, num_levels, image_hw, pre_k,
                                post_k, nms_thr, score_thr, min_size, nms_mode, out_boxes, out_scores, out_index, out_count,
                                workspace, workspace_bytes, stream, 1, nullptr, nullptr);
}

// RegionProposalNetwork._get_top_n_idx (rpn.py:215-228): per level the indices of the top min(pre_k, n_l) raw objectness
// logits in descending order, offset by the level start; out [B, sum_l min(pre_k, n_l)] int64
int launch_rpn_topk(const float* objectness, int batch, int total, const int* level_sizes_host, int num_levels, int pre_k,
                    long long* out_index, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    RpnParams P{};
    P.C = 1; P.atotal = total;
    P.obj = objectness; P.B = batch; P.total = total; P.L = num_levels; P.pre_k = pre_k;
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
    if (off != total || kmax > 8192) return B200_ERR_INVALID;
    P.Ktot = koff;
    if (!workspace || workspace_bytes < rpn_select_ws_bytes(batch, total, num_levels, pre_k)) return B200_ERR_WORKSPACE;
    P.topk_out = out_index;
    return launch_rpn_select(P, kmax, workspace, workspace_bytes, stream);
}

size_t rpn_topk_workspace_bytes(int batch, int total, int num_levels, int pre_k) {
    return rpn_select_ws_bytes(batch, total, num_levels, pre_k);
}

}  // namespace b200
