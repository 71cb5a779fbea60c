// nms.cu -- segmented greedy NMS without sorting the candidates (sm_100a).
// A segment is one image (YOLO post-process, torchvision nms / coordinate-trick batched_nms) or
// one image x level (RPN).
//
// Greedy NMS in (score desc, index asc) order has a unique characterisation that needs no serial
// walk: box i is kept  <=>  no KEPT box that precedes i suppresses it.  So
//
//  k_nms_pairs   (all SMs) builds the "dominator" bitmask on the boxes in whatever order they
//                arrive: bit j of row i is set iff j precedes i and suppresses it.  64x64 tiles of
//                unordered pairs, each pair evaluated once with the roles (picked S / remaining T)
//                chosen by comparing (score, index) keys, the IoU arithmetic reproduced operation
//                by operation per flavour (appendix A.3).  The IEEE division is only executed when
//                inter is within 1e-6 (relative) of thr*union; outside that band the comparison of
//                the rounded quotient with the threshold is already decided.
//  k_nms_resolve (one CTA per segment) iterates the fixed point in parallel rounds over bitsets in
//                shared memory (a box becomes KEPT when all its dominators are removed, REMOVED as
//                soon as one of them is kept; the number of rounds is the depth of the suppression
//                chains, a handful in practice), then sorts only the KEPT boxes by score, finds
//                each removed box's first suppressor, applies the majority relabel
//                (helper.py:368-375) and emits.
//  k_nms_canon   (YOLO stage API only) sorts the unordered candidate slab by flat anchor index and
//                writes the reference's ascending-anchor candidate list (test_one_epoch.py:27-28).
#include "decode.cuh"
#include "nms.cuh"

namespace b200 {

static constexpr int kCanonThreads = 512;
static constexpr int kSortSmemKeys = 4096;
static constexpr int kPairThreads = 256;
static constexpr int kResolveThreads = 512;
static constexpr int kResolveWarps = kResolveThreads / 32;
static constexpr int kKeptSmem = 2048;          // kept boxes sorted in shared memory up to this many
static constexpr int kVoteFlag = 1 << 30;
static constexpr int kVoteListCap = 128;

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

template <bool SLAB>
__device__ __forceinline__ Item load_item(const NmsParams& P, long long off, int i, float unit) {
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
    if (P.mode == B200_NMS_TV_TRICK) {
        // boxes + idxs.to(boxes) * (boxes.max() + 1)   (torchvision boxes.py coordinate trick)
        const float sh = __fmul_rn((float)it.label, unit);
        it.b = make_float4(__fadd_rn(it.b.x, sh), __fadd_rn(it.b.y, sh), __fadd_rn(it.b.z, sh), __fadd_rn(it.b.w, sh));
    }
    it.area = __fmul_rn(__fsub_rn(it.b.z, it.b.x), __fsub_rn(it.b.w, it.b.y));
    it.key = ((unsigned long long)(~orderable(score)) << 32) | tie;
    return it;
}

// S = the box that precedes (picked), T = the later one (remaining).  exact: always divide.
template <int MODE, bool EXACT>
__device__ __forceinline__ bool suppresses(const NmsParams& P, const float4& bs, float as, int ls,
                                           const float4& bt, float at, int lt, bool* vote) {
    if (MODE == B200_NMS_TV_CLASS && ls != lt) return false;
    float w = __fsub_rn(fminf(bs.z, bt.z), fmaxf(bs.x, bt.x));
    float h = __fsub_rn(fminf(bs.w, bt.w), fmaxf(bs.y, bt.y));
    w = fmaxf(w, 0.f);
    h = fmaxf(h, 0.f);
    const float inter = __fmul_rn(w, h);
    // helper.py:361-366  union = (area_T - inter) + area_S ;  torchvision: (area_i + area_j) - inter
    const float den = MODE == B200_NMS_MAJORITY ? __fadd_rn(__fsub_rn(at, inter), as)
                                                : __fsub_rn(__fadd_rn(as, at), inter);
    if (!EXACT) {
        // fl(inter/den) differs from inter/den by < 2^-24 relative: outside a 1e-6 band around
        // thr*den the threshold comparison is decided without dividing.
        const float t = __fmul_rn(den, P.thr_f);
        if (den > 1e-30f) {
            if (inter > __fmul_rn(t, 1.000001f)) return true;
            if (inter < __fmul_rn(t, 0.999999f)) return false;
        }
    }
    const float iou = __fdiv_rn(inter, den);
    if (MODE == B200_NMS_MAJORITY) {
        if (vote) *vote = iou > P.thr_f;                 // helper.py:369
        return !(iou < P.thr_f);                         // helper.py:368 (NaN and == thr are removed)
    }
    return (double)iou > P.thr_d;                        // torchvision: double threshold
}

// ascending bitonic sort of key[0..P) (+ optional payload), P a power of two, whole CTA
__device__ void bitonic_sort(unsigned long long* key, int* val, int P) {
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < (P >> 1); t += blockDim.x) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int ixj = i | j;
                const bool asc = (i & k) == 0;
                const unsigned long long a = key[i], b = key[ixj];
                if ((a > b) == asc) {
                    key[i] = b; key[ixj] = a;
                    if (val) { const int va = val[i]; val[i] = val[ixj]; val[ixj] = va; }
                }
            }
            __syncthreads();
        }
    }
}

// ------------------------------------------------------------------------------------------
// canonicalise the slab (stage API b200_yolo_decode_filter)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kCanonThreads)
k_nms_canon(const __grid_constant__ NmsParams P) {
    __shared__ unsigned long long skey[kSortSmemKeys];
    const int seg = blockIdx.x, tid = threadIdx.x;
    long long off;
    int n, n_true;
    segment_range(P, seg, off, n, n_true);
    if (tid == 0 && P.cand_count_out) P.cand_count_out[seg] = n_true;
    if (n == 0) return;
    const int Ppad = next_pow2(n);
    unsigned long long* key = Ppad <= kSortSmemKeys ? skey : P.gkey + 2 * off;
    const Cand* slab = P.slab + off;
    for (int i = tid; i < Ppad; i += kCanonThreads)
        key[i] = i < n ? (((unsigned long long)(unsigned)slab[i].anchor << 32) | (unsigned)i) : ~0ull;
    __syncthreads();
    bitonic_sort(key, nullptr, Ppad);
    for (int p = tid; p < n; p += kCanonThreads) {
        const Cand c = slab[(unsigned)key[p]];
        P.cbox[off + p] = make_float4(c.x1, c.y1, c.x2, c.y2);
        P.cscore[off + p] = c.score;
        P.clabel[off + p] = c.label;
        P.canchor[off + p] = c.anchor;
    }
}

// coordinate trick: per-segment max coordinate + 1
template <bool SLAB>
__global__ void __launch_bounds__(256)
k_nms_trick_prep(const __grid_constant__ NmsParams P) {
    __shared__ float red[8];
    const int seg = blockIdx.x, tid = threadIdx.x;
    long long off;
    int n, n_true;
    segment_range(P, seg, off, n, n_true);
    float mx = -INFINITY;
    for (int i = tid; i < n; i += 256) {
        const float4 b = SLAB ? reinterpret_cast<const float4*>(P.slab + off + i)[0]
                              : reinterpret_cast<const float4*>(P.boxes)[off + i];
        mx = fmaxf(mx, fmaxf(fmaxf(b.x, b.y), fmaxf(b.z, b.w)));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(kFullMask, mx, o));
    if ((tid & 31) == 0) red[tid >> 5] = mx;
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < 8; ++w) mx = fmaxf(mx, red[w]);
        P.shift_unit[seg] = __fadd_rn(mx, 1.0f);
    }
}

// ------------------------------------------------------------------------------------------
// dominator bitmask: 64x64 tiles of unordered pairs, dynamically scheduled over all SMs
// ------------------------------------------------------------------------------------------
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
    const float t = __fmul_rn(den, P.thr_f);
    if (den > 1e-30f) {
        if (inter > __fmul_rn(t, 1.000001f)) return true;
        if (inter < __fmul_rn(t, 0.999999f)) return false;
    }
    const float iou = __fdiv_rn(inter, den);
    if (MODE == B200_NMS_MAJORITY) return !(iou < P.thr_f);
    return (double)iou > P.thr_d;
}

// one CTA (1024 threads): tiles per segment -> exclusive prefix, and the work counter reset
__global__ void __launch_bounds__(1024)
k_nms_plan(const __grid_constant__ NmsParams P) {
    __shared__ int warp_sum[32];
    __shared__ int carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { carry = 0; *P.work_counter = 0; }
    __syncthreads();
    for (int s0 = 0; s0 < P.num_segments; s0 += 1024) {
        const int s = s0 + tid;
        int tiles = 0;
        if (s < P.num_segments) {
            long long off; int n, n_true;
            segment_range(P, s, off, n, n_true);
            const int nt = cdiv(n, 64);
            tiles = nt * (nt + 1) / 2;
        }
        int incl = tiles;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(kFullMask, incl, o); if (lane >= o) incl += v; }
        if (lane == 31) warp_sum[warp] = incl;
        __syncthreads();
        int before = carry;
        for (int w = 0; w < warp; ++w) before += warp_sum[w];
        if (s < P.num_segments) P.tile_prefix[s] = before + incl - tiles;
        __syncthreads();
        if (tid == 1023) carry = before + incl;
        __syncthreads();
    }
    if (tid == 0) P.tile_prefix[P.num_segments] = carry;
}

template <int MODE, bool SLAB>
__device__ __forceinline__ void pair_tile(const NmsParams& P, int seg, int local, float4* rb, float4* cb, float* ra,
                                          float* ca, unsigned long long* rk, unsigned long long* ck,
                                          unsigned long long* trans, int* rl, int* cl) {
    long long off;
    int n, n_true;
    segment_range(P, seg, off, n, n_true);
    const float unit = P.mode == B200_NMS_TV_TRICK ? P.shift_unit[seg] : 0.f;
    const int nt = cdiv(n, 64);
    int rt = 0, rem = local;
    while (rem >= nt - rt) { rem -= nt - rt; ++rt; }
    const int ct = rt + rem;
    const int tid = threadIdx.x;
    const int r = tid >> 2, cg = tid & 3;
    unsigned long long* dom = P.dom + (size_t)seg * (size_t)P.max_seg * (size_t)P.max_words;
    if (tid < 64) {
        const int i = rt * 64 + tid;
        if (i < n) { const Item it = load_item<SLAB>(P, off, i, unit); rb[tid] = it.b; ra[tid] = it.area; rk[tid] = it.key; rl[tid] = it.label; }
    } else if (tid < 128) {
        const int j = ct * 64 + tid - 64;
        if (j < n) { const Item it = load_item<SLAB>(P, off, j, unit); cb[tid - 64] = it.b; ca[tid - 64] = it.area; ck[tid - 64] = it.key; cl[tid - 64] = it.label; }
    } else if (tid < 192) {
        trans[tid - 128] = 0ull;
    }
    __syncthreads();
    const int i = rt * 64 + r;
    unsigned long long bits = 0ull;
    if (i < n) {
        const float4 bi = rb[r];
        const float ai = ra[r];
        const unsigned long long ki = rk[r];
        const int li = rl[r];
#pragma unroll 4
        for (int c = 0; c < 16; ++c) {
            const int cc = c * 4 + cg;            // the 4 threads of a row read adjacent columns
            const int j = ct * 64 + cc;
            if (j < n && (rt != ct || cc > r)) {
                const bool i_first = ki < ck[cc];
                if (pair_hit<MODE>(P, bi, ai, li, cb[cc], ca[cc], cl[cc], i_first)) {
                    if (i_first) atomicOr(&trans[cc], 1ull << r);   // i dominates j
                    else bits |= 1ull << cc;                        // j dominates i
                }
            }
        }
    }
    bits |= __shfl_xor_sync(kFullMask, bits, 1);
    bits |= __shfl_xor_sync(kFullMask, bits, 2);
    __syncthreads();
    if (rt == ct) {
        if (cg == 0 && i < n) dom[(size_t)i * P.max_words + ct] = bits | trans[r];
    } else {
        if (cg == 0 && i < n) dom[(size_t)i * P.max_words + ct] = bits;
        if (tid < 64 && ct * 64 + tid < n) dom[(size_t)(ct * 64 + tid) * P.max_words + rt] = trans[tid];
    }
}

template <bool SLAB>
__global__ void __launch_bounds__(kPairThreads)
k_nms_pairs(const __grid_constant__ NmsParams P) {
    __shared__ float4 rb[64], cb[64];
    __shared__ float ra[64], ca[64];
    __shared__ unsigned long long rk[64], ck[64], trans[64];
    __shared__ int rl[64], cl[64];
    __shared__ int s_work;
    const int total = P.tile_prefix[P.num_segments];
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_work = atomicAdd(P.work_counter, 1);
        __syncthreads();
        const int t = s_work;
        if (t >= total) break;
        int lo = 0, hi = P.num_segments - 1;          // last segment with tile_prefix[seg] <= t
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (P.tile_prefix[mid] <= t) lo = mid; else hi = mid - 1;
        }
        const int local = t - P.tile_prefix[lo];
        switch (P.mode) {
            case B200_NMS_MAJORITY: pair_tile<B200_NMS_MAJORITY, SLAB>(P, lo, local, rb, cb, ra, ca, rk, ck, trans, rl, cl); break;
            case B200_NMS_TV_CLASS: pair_tile<B200_NMS_TV_CLASS, SLAB>(P, lo, local, rb, cb, ra, ca, rk, ck, trans, rl, cl); break;
            default:                pair_tile<B200_NMS_TV, SLAB>(P, lo, local, rb, cb, ra, ca, rk, ck, trans, rl, cl); break;
        }
    }
}

// ------------------------------------------------------------------------------------------
// resolve + order the kept boxes + vote + emit
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int block_rank(bool flag, int* scratch, int& running) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned bal = __ballot_sync(kFullMask, flag);
    if (lane == 0) scratch[warp] = __popc(bal);
    __syncthreads();
    int before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kResolveWarps; ++w) {
        const int c = scratch[w];
        if (w < warp) before += c;
        total += c;
    }
    const int rank = running + before + __popc(bal & ((1u << lane) - 1u));
    running += total;
    __syncthreads();
    return rank;
}

template <int MODE, bool SLAB>
__device__ void resolve_body(const NmsParams& P, int seg, unsigned char* smem_raw, int* sm_scan, int* sm_vote,
                             unsigned long long* skey, int* sval) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    long long off;
    int n, n_true;
    segment_range(P, seg, off, n, n_true);
    if (tid == 0 && SLAB && P.cand_count_out) P.cand_count_out[seg] = n_true;
    if (n == 0) {
        if (tid == 0) {
            if (P.keep_count) P.keep_count[seg] = 0;
            if (P.det_count) P.det_count[seg] = 0;
        }
        return;
    }
    const float unit = P.mode == B200_NMS_TV_TRICK ? P.shift_unit[seg] : 0.f;
    const int nw = cdiv(n, 64);
    unsigned long long* Kset = reinterpret_cast<unsigned long long*>(smem_raw);   // [max_words]
    unsigned long long* Rset = Kset + P.max_words;                                // [max_words]
    const unsigned long long* dom = P.dom + (size_t)seg * (size_t)P.max_seg * (size_t)P.max_words;
    for (int w = tid; w < nw; w += kResolveThreads) { Kset[w] = 0ull; Rset[w] = 0ull; }
    __syncthreads();

    // ---- fixed point: kept <=> every dominator removed ; removed <=> some dominator kept ---------
    int pending;
    do {
        int undecided = 0;
        for (int i = tid; i < n; i += kResolveThreads) {
            const int wi = i >> 6;
            const unsigned long long bit = 1ull << (i & 63);
            if ((Kset[wi] | Rset[wi]) & bit) continue;
            const unsigned long long* row = dom + (size_t)i * P.max_words;
            bool any_kept = false, all_removed = true;
            for (int w = 0; w < nw; ++w) {
                const unsigned long long d = row[w];
                if (d) {
                    if (d & Kset[w]) { any_kept = true; break; }
                    if (d & ~Rset[w]) all_removed = false;
                }
            }
            if (any_kept) atomicOr(&Rset[wi], bit);
            else if (all_removed) atomicOr(&Kset[wi], bit);
            else ++undecided;
        }
        pending = __syncthreads_count(undecided > 0);
    } while (pending > 0);

    // ---- kept boxes, ordered by (score desc, canonical index asc) ------------------------------------
    int running = 0;
    int* klist = P.gklist + off;
    for (int i0 = 0; i0 < n; i0 += kResolveThreads) {
        const int i = i0 + tid;
        const bool kept = i < n && ((Kset[i >> 6] >> (i & 63)) & 1ull);
        const int k = block_rank(kept, sm_scan, running);
        if (kept) klist[k] = i;
    }
    const int K = running;
    __syncthreads();
    const int Pk = next_pow2(K);
    unsigned long long* key = Pk <= kKeptSmem ? skey : P.gkey + 2 * off;   // Pk < 2K <= 2n
    int* val = Pk <= kKeptSmem ? sval : P.gval + 2 * off;
    for (int t = tid; t < Pk; t += kResolveThreads) {
        if (t < K) {
            const int i = klist[t];
            key[t] = load_item<SLAB>(P, off, i, unit).key;
            val[t] = i;
        } else {
            key[t] = ~0ull;
            val[t] = -1;
        }
    }
    __syncthreads();
    bitonic_sort(key, val, Pk);

    int* newlab = P.gnewlab + off;   // indexed by position
    if (MODE == B200_NMS_MAJORITY) {
        // ---- first suppressor of every removed box = its kept dominator that comes first ----------
        int* sup = P.gsup + off;
        for (int j = tid; j < n; j += kResolveThreads) {
            int s = -1;
            if ((Rset[j >> 6] >> (j & 63)) & 1ull) {
                const unsigned long long* row = dom + (size_t)j * P.max_words;
                unsigned long long best = ~0ull;
                int besti = -1;
                for (int w = 0; w < nw; ++w) {
                    unsigned long long d = row[w] & Kset[w];
                    while (d) {
                        const int i = w * 64 + __ffsll((long long)d) - 1;
                        d &= d - 1ull;
                        const unsigned long long k = load_item<SLAB>(P, off, i, unit).key;
                        if (k < best) { best = k; besti = i; }
                    }
                }
                if (besti >= 0) {
                    const Item S = load_item<SLAB>(P, off, besti, unit);
                    const Item T = load_item<SLAB>(P, off, j, unit);
                    bool vote = false;
                    suppresses<MODE, true>(P, S.b, S.area, S.label, T.b, T.area, T.label, &vote);
                    s = besti | (vote ? kVoteFlag : 0);
                }
            }
            sup[j] = s;
        }
        __syncthreads();
        // ---- majority relabel (helper.py:368-375): one warp per kept box ----------------------------
        int* list = sm_vote + warp * kVoteListCap;
        for (int t = warp; t < K; t += kResolveWarps) {
            const int i = klist[t];
            const int want = i | kVoteFlag;
            int L = 0;
            for (int j0 = 0; j0 < n; j0 += 32) {
                const int j = j0 + lane;
                const bool v = j < n && sup[j] == want;
                const unsigned bal = __ballot_sync(kFullMask, v);
                if (v) {
                    const int pos = L + __popc(bal & ((1u << lane) - 1u));
                    if (pos < kVoteListCap) list[pos] = load_item<SLAB>(P, off, j, unit).label;
                }
                L += __popc(bal);
            }
            __syncwarp();
            int label = load_item<SLAB>(P, off, i, unit).label;
            if (L >= 2) {
                int best_cnt = 0, best_lab = 0x7fffffff;
                if (L <= kVoteListCap) {
                    for (int a = lane; a < L; a += 32) {
                        const int la = list[a];
                        int cnt = 0;
                        for (int b = 0; b < L; ++b) cnt += (list[b] == la);
                        if (cnt > best_cnt || (cnt == best_cnt && la < best_lab)) { best_cnt = cnt; best_lab = la; }
                    }
                } else {   // rare: more voters than the list holds -> recount by rescanning
                    for (int j = lane; j < n; j += 32) {
                        if (sup[j] != want) continue;
                        const int la = load_item<SLAB>(P, off, j, unit).label;
                        int cnt = 0;
                        for (int b = 0; b < n; ++b)
                            cnt += (sup[b] == want && load_item<SLAB>(P, off, b, unit).label == la);
                        if (cnt > best_cnt || (cnt == best_cnt && la < best_lab)) { best_cnt = cnt; best_lab = la; }
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const int oc = __shfl_xor_sync(kFullMask, best_cnt, o);
                    const int ol = __shfl_xor_sync(kFullMask, best_lab, o);
                    if (oc > best_cnt || (oc == best_cnt && ol < best_lab)) { best_cnt = oc; best_lab = ol; }
                }
                if (best_cnt < L) label = best_lab;   // more than one distinct class among the voters
            }
            if (lane == 0) newlab[i] = label;
            __syncwarp();
        }
        __syncthreads();
    }

    // ---- emit in score order -----------------------------------------------------------------------------
    for (int t = tid; t < K; t += kResolveThreads) {
        const int i = val[t];
        if (SLAB) {
            if (t < P.max_det) {
                const Cand c = P.slab[off + i];
                float* d = P.det + ((size_t)seg * P.max_det + t) * 6;
                d[0] = c.x1; d[1] = c.y1; d[2] = c.x2; d[3] = c.y2;
                d[4] = c.score;
                d[5] = (float)(MODE == B200_NMS_MAJORITY ? newlab[i] : c.label);
                if (P.det_anchor) P.det_anchor[(size_t)seg * P.max_det + t] = c.anchor;
                if (P.det_keep) {
                    // index in the reference's candidate list = rank of the flat anchor index
                    int rank = 0;
                    for (int j = 0; j < n; ++j) rank += (P.slab[off + j].anchor < c.anchor);
                    P.det_keep[(size_t)seg * P.max_det + t] = rank;
                }
            }
        } else {
            P.keep[off + t] = i;
            if (P.labels_out) P.labels_out[off + t] = MODE == B200_NMS_MAJORITY ? newlab[i] : (P.labels ? P.labels[off + i] : 0);
        }
    }
    if (tid == 0) {
        if (SLAB) {
            P.det_count[seg] = min(K, P.max_det);
            if (K > P.max_det && P.status) atomicOr(P.status, 2);
        } else {
            P.keep_count[seg] = K;
        }
    }
}

// ------------------------------------------------------------------------------------------
// fast resolve: the whole segment lives in shared memory (n <= kFastN boxes, <= kFastE edges)
// ------------------------------------------------------------------------------------------
static constexpr int kFastN = 2048;
static constexpr int kFastE = 40960;

struct FastSmem {
    float4* box;               // [kFastN]
    float* area;               // [kFastN]
    unsigned long long* key;   // [kFastN]
    int* lab;                  // [kFastN]
    int* off;                  // [kFastN+1] dominator list offsets
    unsigned short* edge;      // [kFastE]   dominator indices; later reused as sort keys + payload
    unsigned char* state;      // [kFastN]   0 undecided, 1 kept, 2 removed
    int* sup;                  // [kFastN]   first suppressor | vote flag
    int* voff;                 // [kFastN+1] voter list offsets
    int* vlab;                 // [kFastN]   voter labels
    int* newlab;               // [kFastN]
    int* klist;                // [kFastN]
    int* scan;                 // [32]
};
__host__ __device__ inline size_t fast_smem_carve(FastSmem* f, unsigned char* base) {
    size_t o = 0;
    auto take = [&](size_t bytes) { unsigned char* r = base ? base + o : nullptr; o += align_up(bytes, 16); return r; };
    unsigned char* p;
    p = take(16 * kFastN);       if (f) f->box = reinterpret_cast<float4*>(p);
    p = take(4 * kFastN);        if (f) f->area = reinterpret_cast<float*>(p);
    p = take(8 * kFastN);        if (f) f->key = reinterpret_cast<unsigned long long*>(p);
    p = take(4 * kFastN);        if (f) f->lab = reinterpret_cast<int*>(p);
    p = take(4 * (kFastN + 1));  if (f) f->off = reinterpret_cast<int*>(p);
    p = take(2 * kFastE);        if (f) f->edge = reinterpret_cast<unsigned short*>(p);
    p = take(kFastN);            if (f) f->state = p;
    p = take(4 * kFastN);        if (f) f->sup = reinterpret_cast<int*>(p);
    p = take(4 * (kFastN + 1));  if (f) f->voff = reinterpret_cast<int*>(p);
    p = take(4 * kFastN);        if (f) f->vlab = reinterpret_cast<int*>(p);
    p = take(4 * kFastN);        if (f) f->newlab = reinterpret_cast<int*>(p);
    p = take(4 * kFastN);        if (f) f->klist = reinterpret_cast<int*>(p);
    p = take(4 * 32);            if (f) f->scan = reinterpret_cast<int*>(p);
    return o;
}

// in-place exclusive scan of a[0..n) (whole CTA), returns the total
__device__ __forceinline__ int block_exclusive_scan(int* a, int n, int* scratch) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int carry = 0;
    for (int base = 0; base < n; base += kResolveThreads) {
        const int i = base + tid;
        const int v = i < n ? a[i] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(kFullMask, incl, o); if (lane >= o) incl += u; }
        if (lane == 31) scratch[warp] = incl;
        __syncthreads();
        int before = 0, total = 0;
#pragma unroll
        for (int w = 0; w < kResolveWarps; ++w) { const int c = scratch[w]; if (w < warp) before += c; total += c; }
        if (i < n) a[i] = carry + before + incl - v;
        carry += total;
        __syncthreads();
    }
    return carry;
}

// returns false (before touching any output) when the dominator lists do not fit -> slow path
template <int MODE, bool SLAB>
__device__ bool resolve_fast(const NmsParams& P, int seg, long long off, int n, unsigned char* smem_raw) {
    FastSmem f;
    fast_smem_carve(&f, smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float unit = P.mode == B200_NMS_TV_TRICK ? P.shift_unit[seg] : 0.f;
    const int nw = cdiv(n, 64);
    const unsigned long long* dom = P.dom + (size_t)seg * (size_t)P.max_seg * (size_t)P.max_words;

    // ---- A. boxes + in-degree ---------------------------------------------------------------------
    for (int i = tid; i < n; i += kResolveThreads) {
        const Item it = load_item<SLAB>(P, off, i, unit);
        f.box[i] = it.b; f.area[i] = it.area; f.key[i] = it.key; f.lab[i] = it.label;
        f.state[i] = 0;
        const unsigned long long* row = dom + (size_t)i * P.max_words;
        int deg = 0;
        for (int w = 0; w < nw; ++w) deg += __popcll(row[w]);
        f.off[i] = deg;
    }
    if (tid == 0) f.off[n] = 0;
    __syncthreads();
    const int E = block_exclusive_scan(f.off, n + 1, f.scan);
    if (E > kFastE) return false;
    // ---- B. dominator index lists -----------------------------------------------------------------
    for (int i = tid; i < n; i += kResolveThreads) {
        const unsigned long long* row = dom + (size_t)i * P.max_words;
        int e = f.off[i];
        if (f.off[i + 1] == e) continue;
        for (int w = 0; w < nw; ++w) {
            unsigned long long d = row[w];
            while (d) {
                f.edge[e++] = (unsigned short)(w * 64 + __ffsll((long long)d) - 1);
                d &= d - 1ull;
            }
        }
    }
    __syncthreads();
    // ---- C. fixed point ---------------------------------------------------------------------------
    int pending;
    do {
        int undecided = 0;
        for (int i = tid; i < n; i += kResolveThreads) {
            if (f.state[i]) continue;
            bool any_kept = false, all_removed = true;
            for (int e = f.off[i], e1 = f.off[i + 1]; e < e1; ++e) {
                const int s = f.state[f.edge[e]];
                if (s == 1) { any_kept = true; break; }
                if (s == 0) all_removed = false;
            }
            if (any_kept) f.state[i] = 2;
            else if (all_removed) f.state[i] = 1;
            else ++undecided;
        }
        pending = __syncthreads_count(undecided > 0);
    } while (pending > 0);

    if (MODE == B200_NMS_MAJORITY) {
        // ---- D. first suppressor + vote (helper.py:368-369), voters gathered per kept box ---------
        for (int i = tid; i <= n; i += kResolveThreads) f.voff[i] = 0;
        __syncthreads();
        for (int j = tid; j < n; j += kResolveThreads) {
            int s = -1;
            if (f.state[j] == 2) {
                unsigned long long best = ~0ull;
                int besti = -1;
                for (int e = f.off[j], e1 = f.off[j + 1]; e < e1; ++e) {
                    const int i = f.edge[e];
                    if (f.state[i] == 1 && f.key[i] < best) { best = f.key[i]; besti = i; }
                }
                bool vote = false;
                suppresses<MODE, true>(P, f.box[besti], f.area[besti], 0, f.box[j], f.area[j], 0, &vote);
                s = besti | (vote ? kVoteFlag : 0);
                if (vote) atomicAdd(&f.voff[besti], 1);
            }
            f.sup[j] = s;
        }
        __syncthreads();
        block_exclusive_scan(f.voff, n + 1, f.scan);
        for (int i = tid; i < n; i += kResolveThreads) f.klist[i] = f.voff[i];    // fill cursors
        __syncthreads();
        for (int j = tid; j < n; j += kResolveThreads) {
            const int s = f.sup[j];
            if (s >= 0 && (s & kVoteFlag)) f.vlab[atomicAdd(&f.klist[s & ~kVoteFlag], 1)] = f.lab[j];
        }
        __syncthreads();
        // majority relabel (helper.py:370-375): one warp per kept box
        for (int i = warp; i < n; i += kResolveWarps) {
            if (f.state[i] != 1) continue;
            const int v0 = f.voff[i], L = f.voff[i + 1] - v0;
            int label = f.lab[i];
            if (L >= 2) {
                int best_cnt = 0, best_lab = 0x7fffffff;
                for (int a = lane; a < L; a += 32) {
                    const int la = f.vlab[v0 + a];
                    int cnt = 0;
                    for (int b = 0; b < L; ++b) cnt += (f.vlab[v0 + b] == la);
                    if (cnt > best_cnt || (cnt == best_cnt && la < best_lab)) { best_cnt = cnt; best_lab = la; }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const int oc = __shfl_xor_sync(kFullMask, best_cnt, o);
                    const int ol = __shfl_xor_sync(kFullMask, best_lab, o);
                    if (oc > best_cnt || (oc == best_cnt && ol < best_lab)) { best_cnt = oc; best_lab = ol; }
                }
                if (best_cnt < L) label = best_lab;   // more than one distinct class among the voters
            }
            if (lane == 0) f.newlab[i] = label;
        }
        __syncthreads();
    }

    // ---- E. kept boxes in (score desc, canonical index asc) order ------------------------------------
    int running = 0;
    for (int i0 = 0; i0 < n; i0 += kResolveThreads) {
        const int i = i0 + tid;
        const bool kept = i < n && f.state[i] == 1;
        const int k = block_rank(kept, f.scan, running);
        if (kept) f.klist[k] = i;
    }
    const int K = running;
    const int Pk = next_pow2(K);
    unsigned long long* skey = reinterpret_cast<unsigned long long*>(f.edge);      // edges are dead now
    int* sval = reinterpret_cast<int*>(skey + kFastN);
    for (int t = tid; t < Pk; t += kResolveThreads) {
        skey[t] = t < K ? f.key[f.klist[t]] : ~0ull;
        sval[t] = t < K ? f.klist[t] : -1;
    }
    __syncthreads();
    bitonic_sort(skey, sval, Pk);

    // ---- F. emit -------------------------------------------------------------------------------------
    for (int t = tid; t < K; t += kResolveThreads) {
        const int i = sval[t];
        const int lab = MODE == B200_NMS_MAJORITY ? f.newlab[i] : f.lab[i];
        if (SLAB) {
            if (t < P.max_det) {
                const float4 b = reinterpret_cast<const float4*>(P.slab + off + i)[0];   // unshifted box
                const unsigned anchor = (unsigned)f.key[i];
                float* d = P.det + ((size_t)seg * P.max_det + t) * 6;
                d[0] = b.x; d[1] = b.y; d[2] = b.z; d[3] = b.w;
                d[4] = from_orderable(~(unsigned)(f.key[i] >> 32));
                d[5] = (float)lab;
                if (P.det_anchor) P.det_anchor[(size_t)seg * P.max_det + t] = (int)anchor;
                if (P.det_keep) {
                    // index in the reference's candidate list = rank of the flat anchor index
                    int rank = 0;
                    for (int j = 0; j < n; ++j) rank += ((unsigned)f.key[j] < anchor);
                    P.det_keep[(size_t)seg * P.max_det + t] = rank;
                }
            }
        } else {
            P.keep[off + t] = i;
            if (P.labels_out) P.labels_out[off + t] = lab;
        }
    }
    if (tid == 0) {
        if (SLAB) {
            P.det_count[seg] = min(K, P.max_det);
            if (K > P.max_det && P.status) atomicOr(P.status, 2);
        } else {
            P.keep_count[seg] = K;
        }
    }
    return true;
}

// dynamic shared memory: max(fast layout, slow layout)
__host__ __device__ inline size_t slow_smem_bytes(int max_words) {
    return align_up(sizeof(unsigned long long) * 2 * (size_t)max_words, 16) + sizeof(unsigned long long) * kKeptSmem +
           sizeof(int) * kKeptSmem + sizeof(int) * kResolveWarps * kVoteListCap + sizeof(int) * 32;
}

template <bool SLAB>
__global__ void __launch_bounds__(kResolveThreads, 1)
k_nms_resolve(const __grid_constant__ NmsParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int seg = blockIdx.x;
    long long off;
    int n, n_true;
    segment_range(P, seg, off, n, n_true);
    if (n > 0 && n <= kFastN) {
        const bool done = P.mode == B200_NMS_MAJORITY ? resolve_fast<B200_NMS_MAJORITY, SLAB>(P, seg, off, n, smem_raw)
                                                      : resolve_fast<B200_NMS_TV, SLAB>(P, seg, off, n, smem_raw);
        if (done) {
            if (threadIdx.x == 0 && SLAB && P.cand_count_out) P.cand_count_out[seg] = n_true;
            return;
        }
        __syncthreads();
    }
    unsigned char* q = smem_raw + align_up(sizeof(unsigned long long) * 2 * (size_t)P.max_words, 16);
    unsigned long long* skey = reinterpret_cast<unsigned long long*>(q); q += sizeof(unsigned long long) * kKeptSmem;
    int* sval = reinterpret_cast<int*>(q);                                q += sizeof(int) * kKeptSmem;
    int* sm_vote = reinterpret_cast<int*>(q);                             q += sizeof(int) * kResolveWarps * kVoteListCap;
    int* sm_scan = reinterpret_cast<int*>(q);
    if (P.mode == B200_NMS_MAJORITY)
        resolve_body<B200_NMS_MAJORITY, SLAB>(P, seg, smem_raw, sm_scan, sm_vote, skey, sval);
    else
        resolve_body<B200_NMS_TV, SLAB>(P, seg, smem_raw, sm_scan, sm_vote, skey, sval);
}

// ------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------
namespace {
struct Carve {
    unsigned char* p;
    size_t used;
    bool query;
    void* take(size_t bytes) {
        bytes = align_up(bytes, 256);
        void* r = query ? nullptr : p + used;
        used += bytes;
        return r;
    }
};
size_t carve_all(NmsParams* P, size_t T, size_t S, size_t max_seg, void* base, bool query) {
    if (T == 0) T = 1;
    if (S == 0) S = 1;
    if (max_seg == 0) max_seg = 1;
    const size_t words = (max_seg + 63) / 64;
    Carve c{reinterpret_cast<unsigned char*>(base), 0, query};
    NmsParams tmp{};
    NmsParams& o = P ? *P : tmp;
    o.gkey = (unsigned long long*)c.take(16 * T);
    o.gval = (int*)c.take(8 * T);
    o.gsup = (int*)c.take(4 * T);
    o.gklist = (int*)c.take(4 * T);
    o.gnewlab = (int*)c.take(4 * T);
    o.shift_unit = (float*)c.take(4 * S);
    o.tile_prefix = (int*)c.take(4 * (S + 1));
    o.work_counter = (int*)c.take(4);
    o.dom = (unsigned long long*)c.take(8 * S * max_seg * words);
    return c.used;
}
}  // namespace

size_t nms_scratch_bytes(size_t total, size_t segments, size_t max_seg) {
    return carve_all(nullptr, total, segments, max_seg, nullptr, true);
}

bool nms_carve_scratch(NmsParams* P, size_t total, size_t segments, size_t max_seg, void* base, size_t bytes) {
    if (!base || (reinterpret_cast<uintptr_t>(base) & 255u)) return false;
    if (carve_all(nullptr, total, segments, max_seg, nullptr, true) > bytes) return false;
    carve_all(P, total, segments, max_seg, base, false);
    return true;
}

int launch_nms(NmsParams& P, int num_segments, cudaStream_t stream) {
    if (num_segments <= 0) return B200_OK;
    if (num_segments > 65535) return B200_ERR_INVALID;
    if (P.max_seg < 1) P.max_seg = 1;
    P.max_words = cdiv(P.max_seg, 64);
    if (P.mode < 0) {
        if (!P.from_slab) return B200_ERR_INVALID;
        k_nms_canon<<<num_segments, kCanonThreads, 0, stream>>>(P);
        return cudaGetLastError() == cudaSuccess ? B200_OK : B200_ERR_CUDA;
    }
    if (P.mode == B200_NMS_TV_TRICK) {
        if (P.from_slab) k_nms_trick_prep<true><<<num_segments, 256, 0, stream>>>(P);
        else             k_nms_trick_prep<false><<<num_segments, 256, 0, stream>>>(P);
    }
    static int sms = 0;
    if (sms == 0) {
        int dev = 0, v = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0)
            sms = v;
        else
            sms = 148;
    }
    P.num_segments = num_segments;
    k_nms_plan<<<1, 1024, 0, stream>>>(P);
    const int pair_ctas = 8 * sms;                       // 8 x 256 threads per SM, tiles pulled from a queue
    if (P.from_slab) k_nms_pairs<true><<<pair_ctas, kPairThreads, 0, stream>>>(P);
    else             k_nms_pairs<false><<<pair_ctas, kPairThreads, 0, stream>>>(P);

    const size_t fast_bytes = fast_smem_carve(nullptr, nullptr);
    const size_t slow_bytes = slow_smem_bytes(P.max_words);
    const size_t smem = fast_bytes > slow_bytes ? fast_bytes : slow_bytes;
    if (smem > 227 * 1024) return B200_ERR_INVALID;   // max_seg beyond ~700k boxes
    static size_t attr_bytes[2] = {0, 0};
    if (smem > attr_bytes[P.from_slab ? 1 : 0]) {
        const cudaError_t e = P.from_slab
            ? cudaFuncSetAttribute(k_nms_resolve<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
            : cudaFuncSetAttribute(k_nms_resolve<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return B200_ERR_CUDA;
        attr_bytes[P.from_slab ? 1 : 0] = smem;
    }
    if (P.from_slab) k_nms_resolve<true><<<num_segments, kResolveThreads, smem, stream>>>(P);
    else             k_nms_resolve<false><<<num_segments, kResolveThreads, smem, stream>>>(P);
    return cudaGetLastError() == cudaSuccess ? B200_OK : B200_ERR_CUDA;
}

}  // namespace b200
