// nms.cu -- segmented greedy NMS without sorting the candidates (sm_100a).
// A segment is one image (YOLO post-process, torchvision nms / coordinate-trick batched_nms) or
// one image x level (RPN).
//
// Greedy NMS in (score desc, index asc) order has a characterisation that needs no serial walk:
// box i is kept  <=>  no KEPT box that precedes i suppresses it.  Pipeline per batch:
//
//  k_nms_plan    (one CTA per segment) counting-sorts the boxes into (size class, y band) bins so
//                that 64 consecutive positions are spatially coherent, summarises every 64-tile
//                (union box, area range), zeroes the bitmask rows and emits only those tile pairs
//                that can contain a suppressing pair (union boxes intersect and the area ranges
//                allow IoU >= thr) into a global work list.
//  k_nms_pairs   (all SMs, persistent CTAs popping the work list) builds the "dominator" bitmask:
//                bit q of row p is set iff q precedes p and suppresses it.  Each unordered pair is
//                evaluated once, roles (picked S / remaining T) chosen by comparing (score, index)
//                keys, IoU arithmetic reproduced operation by operation per flavour (appendix
//                A.3).  The IEEE division is only executed when inter is within 1e-6 (relative) of
//                thr*union; outside that band the comparison of the rounded quotient is decided.
//  k_nms_resolve (one CTA per segment) brings boxes and mask rows into shared memory and iterates
//                the fixed point in parallel rounds over bitsets (KEPT when every dominator is
//                removed, REMOVED as soon as one is kept; rounds = depth of the suppression chains),
//                sorts only the KEPT boxes by score, finds each removed box's first suppressor,
//                applies the majority relabel (helper.py:368-375) and emits.
//  k_nms_canon   (YOLO stage API only) sorts the unordered candidate slab by flat anchor index and
//                writes the reference's ascending-anchor candidate list (test_one_epoch.py:27-28).
#include "nms_dev.cuh"

namespace b200 {

static constexpr int kCanonThreads = 512;
static constexpr int kSortSmemKeys = 4096;
static constexpr int kPlanThreads = 512;
static constexpr int kPlanTiles = 512;          // tile summaries kept in shared memory (n <= 32768)
static constexpr int kBins = 256;               // 16 size classes (one octave of area each) x 16 y bands
static constexpr int kBinsPerLane = kBins / 32;
static constexpr int kPairThreads = 256;
#ifndef RESOLVE_MINB
#define RESOLVE_MINB 2
#endif
static constexpr int kResolveMaxThreads = 1024;     // resolve CTAs are launched with g_resolve_threads <= this
static constexpr int kResolveMaxWarps = kResolveMaxThreads / 32;
#define kResolveThreads ((int)blockDim.x)
#define kResolveWarps ((int)(blockDim.x >> 5))
static constexpr int kKeptSmem = 2048;          // slow path: kept boxes sorted in shared memory
static constexpr int kRankSortMax = 1024;       // kept boxes ordered by counting below this, bitonic network above
static constexpr int kVoteFlag = 1 << 30;
static constexpr int kVoteListCap = 128;
static constexpr int kSplitBoxes = 1500;        // largest segment the shared-memory resolve holds at its default 112 KB
int g_serial_split = 1500;                      // split point of serial calls (b200_debug_set_serial_split)



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

// exclusive rank of a flag over the CTA (nwarps warps, runtime); `running` accumulates the total
__device__ __forceinline__ int block_rank_rt(bool flag, int* scratch, int& running, int nwarps) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned bal = __ballot_sync(kFullMask, flag);
    if (lane == 0) scratch[warp] = __popc(bal);
    __syncthreads();
    int before = 0, total = 0;
    for (int w = 0; w < nwarps; ++w) {
        const int c = scratch[w];
        if (w < warp) before += c;
        total += c;
    }
    const int rank = running + before + __popc(bal & ((1u << lane) - 1u));
    running += total;
    __syncthreads();
    return rank;
}

// exclusive rank of a flag over the CTA (NW warps); `running` accumulates the total
template <int NW>
__device__ __forceinline__ int block_rank(bool flag, int* scratch, int& running) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned bal = __ballot_sync(kFullMask, flag);
    if (lane == 0) scratch[warp] = __popc(bal);
    __syncthreads();
    int before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
        const int c = scratch[w];
        if (w < warp) before += c;
        total += c;
    }
    const int rank = running + before + __popc(bal & ((1u << lane) - 1u));
    running += total;
    __syncthreads();
    return rank;
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

// ------------------------------------------------------------------------------------------
// plan: spatial binning, tile summaries, pruned work list
// ------------------------------------------------------------------------------------------

template <bool SLAB>
__global__ void __launch_bounds__(kPlanThreads)
k_nms_plan(const __grid_constant__ NmsParams P) {
    __shared__ float red[kPlanThreads / 32];
    __shared__ int hist[kBins];
    __shared__ float4 tbox[kPlanTiles];
    __shared__ float2 tarea[kPlanTiles];
    __shared__ unsigned char tbad[kPlanTiles];
    __shared__ int s_scan[kPlanThreads / 32];
    __shared__ int s_base;
    const int seg = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    long long off;
    int n, n_true;
    segment_range(P, seg, off, n, n_true);
    if (n == 0 || n > P.split_hi) return;         // larger segments belong to the single-launch path

    // ---- coordinate-trick unit, y range of the box centres ------------------------------------
    float unit = 0.f;
    // torchvision.ops.batched_nms: coordinate trick unless boxes.numel() > limit (then per-class "vanilla")
    if (P.given_unit) {
        unit = P.given_unit[seg];
    } else if (P.mode == B200_NMS_TV_TRICK || (P.mode == B200_NMS_TV_AUTO && 4ll * n_true <= P.auto_limit)) {
        float mx = -INFINITY;
        for (int i = tid; i < n; i += kPlanThreads) {
            const float4 b = SLAB ? reinterpret_cast<const float4*>(P.slab + off + i)[0]
                                  : reinterpret_cast<const float4*>(P.boxes)[off + i];
            mx = fmaxf(mx, fmaxf(fmaxf(b.x, b.y), fmaxf(b.z, b.w)));
        }
        unit = __fadd_rn(block_reduce_max(mx, red), 1.0f);
    }
    if (uses_shift(P) && tid == 0) P.shift_unit[seg] = unit;
    float lo = INFINITY, hi = -INFINITY;
    for (int i = tid; i < n; i += kPlanThreads) {
        const Item it = load_raw<SLAB>(P, off, i, unit);
        const float cy = 0.5f * (it.b.y + it.b.w);
        if (cy == cy) { lo = fminf(lo, cy); hi = fmaxf(hi, cy); }
    }
    const float ymax = block_reduce_max(hi, red);
    const float ymin = -block_reduce_max(-lo, red);
    const float range = ymax - ymin;
    const float yscale = range > 0.f && range < 3.0e38f ? 16.0f / range : 0.f;

    // ---- counting sort into (size class, y band) bins -> gperm ----------------------------------
    for (int b = tid; b < kBins; b += kPlanThreads) hist[b] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += kPlanThreads) atomicAdd(&hist[box_bin(load_raw<SLAB>(P, off, i, unit), ymin, yscale)], 1);
    __syncthreads();
    if (warp == 0) {                      // exclusive scan of the bins, kBinsPerLane per lane
        int v[kBinsPerLane], sum = 0;
#pragma unroll
        for (int k = 0; k < kBinsPerLane; ++k) { v[k] = hist[lane * kBinsPerLane + k]; sum += v[k]; }
        int incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(kFullMask, incl, o); if (lane >= o) incl += u; }
        int run = incl - sum;
#pragma unroll
        for (int k = 0; k < kBinsPerLane; ++k) { hist[lane * kBinsPerLane + k] = run; run += v[k]; }
    }
    __syncthreads();
    int* perm = P.gperm + off;
    for (int i = tid; i < n; i += kPlanThreads)
        perm[atomicAdd(&hist[box_bin(load_raw<SLAB>(P, off, i, unit), ymin, yscale)], 1)] = i;
    __syncthreads();

    // ---- tile summaries + zeroed bitmask rows -----------------------------------------------------
    const int nt = cdiv(n, 64);
    const bool summarise = nt <= kPlanTiles && P.thr_f > 0.f;
    if (summarise) {
        for (int t = warp; t < nt; t += kPlanThreads / 32) {
            float x1 = INFINITY, y1 = INFINITY, x2 = -INFINITY, y2 = -INFINITY, amin = INFINITY, amax = -INFINITY;
            bool bad = false;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int p = t * 64 + h * 32 + lane;
                if (p < n) {
                    const Item it = load_raw<SLAB>(P, off, perm[p], unit);
                    x1 = fminf(x1, it.b.x); y1 = fminf(y1, it.b.y); x2 = fmaxf(x2, it.b.z); y2 = fmaxf(y2, it.b.w);
                    amin = fminf(amin, it.area); amax = fmaxf(amax, it.area);
                    // degenerate boxes can yield NaN IoU (removed by the majority rule): never prune them
                    bad |= !(it.area > 0.f) || !(it.area < 3.0e38f) || !(it.b.x == it.b.x) || !(it.b.y == it.b.y) ||
                           !(it.b.z == it.b.z) || !(it.b.w == it.b.w);
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                x1 = fminf(x1, __shfl_xor_sync(kFullMask, x1, o)); y1 = fminf(y1, __shfl_xor_sync(kFullMask, y1, o));
                x2 = fmaxf(x2, __shfl_xor_sync(kFullMask, x2, o)); y2 = fmaxf(y2, __shfl_xor_sync(kFullMask, y2, o));
                amin = fminf(amin, __shfl_xor_sync(kFullMask, amin, o)); amax = fmaxf(amax, __shfl_xor_sync(kFullMask, amax, o));
            }
            bad = __any_sync(kFullMask, bad);
            if (lane == 0) { tbox[t] = make_float4(x1, y1, x2, y2); tarea[t] = make_float2(amin, amax); tbad[t] = bad ? 1 : 0; }
        }
    }
    // rows of up to 64 words keep a per-row "non-zero word" mask and are never read where it is clear;
    // longer rows (slow resolve path) are cleared in full
    if (nt <= 64) {
        for (int p = tid; p < n; p += kPlanThreads) P.nzmask[off + p] = 0ull;
    } else {
        unsigned long long* dom = P.dom + (size_t)seg * (size_t)P.max_seg * (size_t)P.max_words;
        for (long long e = tid; e < (long long)n * nt; e += kPlanThreads) {
            const long long p = e / nt;
            dom[(size_t)p * P.max_words + (size_t)(e - p * nt)] = 0ull;
        }
    }
    __syncthreads();

    // ---- work list: diagonal tiles always, off-diagonal tiles unless provably empty --------------
    const int total = nt * (nt + 1) / 2;
    const float slack = 0.99f * P.thr_f;
    for (int t0 = 0; t0 < total; t0 += kPlanThreads) {
        const int t = t0 + tid;
        bool keep = false;
        int rt = 0, ct = 0;
        if (t < total) {
            // row-major upper triangle: prefix(r) = r*nt - r*(r-1)/2
            const float fb = (float)(2 * nt + 1);
            rt = (int)((fb - sqrtf(fmaxf(fb * fb - 8.0f * (float)t, 0.f))) * 0.5f);
            rt = min(max(rt, 0), nt - 1);
            while (rt > 0 && rt * nt - rt * (rt - 1) / 2 > t) --rt;
            while ((rt + 1) * nt - (rt + 1) * rt / 2 <= t) ++rt;
            ct = rt + (t - (rt * nt - rt * (rt - 1) / 2));
            keep = true;
            if (summarise && rt != ct && !tbad[rt] && !tbad[ct]) {
                const float4 a = tbox[rt], b = tbox[ct];
                const float2 ra = tarea[rt], rb = tarea[ct];
                const bool disjoint = a.z <= b.x || b.z <= a.x || a.w <= b.y || b.w <= a.y;   // inter == 0
                const bool mismatch = ra.y < slack * rb.x || rb.y < slack * ra.x;             // IoU <= min/max area
                keep = !(disjoint || mismatch);
            }
        }
        int running = 0;
        const int k = block_rank<kPlanThreads / 32>(keep, s_scan, running);
        if (tid == 0) s_base = running > 0 ? atomicAdd(&P.work_count[0], running) : 0;
        __syncthreads();
        if (keep) P.work[s_base + k] = make_int2(seg, (rt << 16) | ct);
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// dominator bitmask: 64x64 tiles of unordered pairs
// ------------------------------------------------------------------------------------------
struct PairSmem {
    float4 rb[64], cb[64];
    float ra[64], ca[64];
    unsigned long long rk[64], ck[64];
    unsigned trans[128];      // transposed hits, 32-bit halves (native shared-memory atomics)
    int rl[64], cl[64];
};

template <int MODE, bool SLAB>
__device__ __forceinline__ void pair_tile(const NmsParams& P, int seg, int rt, int ct, PairSmem& S) {
    long long off;
    int n, n_true;
    segment_range(P, seg, off, n, n_true);
    const float unit = uses_shift(P) ? P.shift_unit[seg] : 0.f;
    const int tid = threadIdx.x;
    const int r = tid >> 2, cg = tid & 3;
    unsigned long long* dom = P.dom + (size_t)seg * (size_t)P.max_seg * (size_t)P.max_words;
    // degenerate boxes (zero / negative / non-finite area, NaN coordinates) can yield NaN IoU, which the majority rule
    // treats as "removed": a tile that holds one never uses the cheap reject below
    int bad = 0;
    if (tid < 64) {
        const int i = rt * 64 + tid;
        if (i < n) {
            const Item it = load_item<SLAB>(P, off, i, unit);
            S.rb[tid] = it.b; S.ra[tid] = it.area; S.rk[tid] = it.key; S.rl[tid] = it.label;
            bad = !(it.area > 0.f) || !(it.area < 3.0e38f) || !(it.b.x == it.b.x) || !(it.b.y == it.b.y) || !(it.b.z == it.b.z) || !(it.b.w == it.b.w);
        }
    } else if (tid < 128) {
        const int j = ct * 64 + tid - 64;
        if (j < n) {
            const Item it = load_item<SLAB>(P, off, j, unit);
            S.cb[tid - 64] = it.b; S.ca[tid - 64] = it.area; S.ck[tid - 64] = it.key; S.cl[tid - 64] = it.label;
            bad = !(it.area > 0.f) || !(it.area < 3.0e38f) || !(it.b.x == it.b.x) || !(it.b.y == it.b.y) || !(it.b.z == it.b.z) || !(it.b.w == it.b.w);
        }
    } else {
        S.trans[tid - 128] = 0u;
    }
    const bool filter = __syncthreads_or(bad) == 0 && P.thr_f > 0.f;
    const float fthr = 0.999f * P.thr_f;
    const int i = rt * 64 + r;
    unsigned long long bits = 0ull;
    if (i < n) {
        const float4 bi = S.rb[r];
        const float ai = S.ra[r];
        const unsigned long long ki = S.rk[r];
        const int li = S.rl[r];
        const float tai = fthr * ai;
#pragma unroll 4
        for (int c = 0; c < 16; ++c) {
            const int cc = c * 4 + cg;            // the 4 threads of a row read adjacent columns
            const int j = ct * 64 + cc;
            if (j < n && (rt != ct || cc > r)) {
                const float4 bj = S.cb[cc];
                const float aj = S.ca[cc];
                if (filter) {
                    // IoU <= inter / max(area): a pair whose intersection is clearly below thr * max(area) cannot
                    // suppress (margin 1e-3, far above any fp32 rounding).  One clamp is enough: with w >= 0 a
                    // negative h makes the product <= 0, below the positive bound.
                    const float w = fmaxf(fminf(bi.z, bj.z) - fmaxf(bi.x, bj.x), 0.f);
                    const float h = fminf(bi.w, bj.w) - fmaxf(bi.y, bj.y);
                    if (w * h < fmaxf(tai, fthr * aj)) continue;
                }
                const bool i_first = ki < S.ck[cc];
                if (pair_hit<MODE>(P, bi, ai, li, bj, aj, S.cl[cc], i_first)) {
                    if (i_first) atomicOr(&S.trans[2 * cc + (r >> 5)], 1u << (r & 31));   // i dominates j
                    else bits |= 1ull << cc;                                              // j dominates i
                }
            }
        }
    }
    bits |= __shfl_xor_sync(kFullMask, bits, 1);
    bits |= __shfl_xor_sync(kFullMask, bits, 2);
    __syncthreads();
    const bool track = n <= 64 * 64;          // rows of <= 64 words: maintain the non-zero word mask
    unsigned long long* nzm = P.nzmask + off;
    if (rt == ct) {
        if (cg == 0 && i < n) {
            const unsigned long long word = bits | ((unsigned long long)S.trans[2 * r + 1] << 32) | S.trans[2 * r];
            if (word) {
                dom[(size_t)i * P.max_words + ct] = word;
                if (track) atomicOr(&nzm[i], 1ull << ct);
            }
        }
    } else {
        if (cg == 0 && i < n && bits) {
            dom[(size_t)i * P.max_words + ct] = bits;
            if (track) atomicOr(&nzm[i], 1ull << ct);
        }
        if (tid < 64 && ct * 64 + tid < n) {
            const unsigned long long word = ((unsigned long long)S.trans[2 * tid + 1] << 32) | S.trans[2 * tid];
            if (word) {
                dom[(size_t)(ct * 64 + tid) * P.max_words + rt] = word;
                if (track) atomicOr(&nzm[ct * 64 + tid], 1ull << rt);
            }
        }
    }
}

template <bool SLAB>
__global__ void __launch_bounds__(kPairThreads)
k_nms_pairs(const __grid_constant__ NmsParams P) {
    __shared__ PairSmem S;
    __shared__ int s_work;
    const int total = P.work_count[0];
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_work = atomicAdd(&P.work_count[1], 1);
        __syncthreads();
        const int t = s_work;
        if (t >= total) break;
        const int2 w = P.work[t];
        const int rt = w.y >> 16, ct = w.y & 0xffff;
        switch (P.mode) {
            case B200_NMS_MAJORITY: pair_tile<B200_NMS_MAJORITY, SLAB>(P, w.x, rt, ct, S); break;
            case B200_NMS_TV_AUTO:    // shifted boxes of different labels never intersect: the label test is exact for both
            case B200_NMS_TV_CLASS: pair_tile<B200_NMS_TV_CLASS, SLAB>(P, w.x, rt, ct, S); break;
            default:                pair_tile<B200_NMS_TV, SLAB>(P, w.x, rt, ct, S); break;   // TV, TV_TRICK
        }
    }
}

// ------------------------------------------------------------------------------------------
// resolve, fast path: the segment lives in shared memory
// ------------------------------------------------------------------------------------------
struct FastSmem {
    float4* box;                 // [n]
    float* area;                 // [n]
    unsigned long long* key;     // [n]
    unsigned long long* nz;      // [n]   which words of the box's dominator row are non-zero
    int* lab;                    // [n]
    int* sup;                    // [n]   first suppressor | vote flag
    int* voff;                   // [n+1] voter list offsets
    int* vlab;                   // [n]   voter labels
    int* newlab;                 // [n]
    int* klist;                  // [n]
    unsigned long long* Kset;    // [nw]  kept bitset
    unsigned long long* Rset;    // [nw]  removed bitset
    int* scan;                   // [32]
    unsigned char* u;            // union: compact mask rows [4n] | sort keys + payload [P2]
};
static constexpr int kCompactWords = 4;   // non-zero row words cached per box (spatial binning keeps rows sparse)
__host__ __device__ inline size_t fast_fixed_bytes(int n, int nw) {
    return 16 * 14 + 60 * (size_t)n + 4 + 16 * (size_t)nw + 128;
}
__device__ __forceinline__ void fast_carve(FastSmem& f, unsigned char* base, int n, int nw) {
    size_t o = 0;
    auto take = [&](size_t bytes) { unsigned char* r = base + o; o += align_up(bytes, 16); return r; };
    f.box = reinterpret_cast<float4*>(take(16 * (size_t)n));
    f.area = reinterpret_cast<float*>(take(4 * (size_t)n));
    f.key = reinterpret_cast<unsigned long long*>(take(8 * (size_t)n));
    f.nz = reinterpret_cast<unsigned long long*>(take(8 * (size_t)n));
    f.lab = reinterpret_cast<int*>(take(4 * (size_t)n));
    f.sup = reinterpret_cast<int*>(take(4 * (size_t)n));
    f.voff = reinterpret_cast<int*>(take(4 * ((size_t)n + 1)));
    f.vlab = reinterpret_cast<int*>(take(4 * (size_t)n));
    f.newlab = reinterpret_cast<int*>(take(4 * (size_t)n));
    f.klist = reinterpret_cast<int*>(take(4 * (size_t)n));
    f.Kset = reinterpret_cast<unsigned long long*>(take(8 * (size_t)nw));
    f.Rset = reinterpret_cast<unsigned long long*>(take(8 * (size_t)nw));
    f.scan = reinterpret_cast<int*>(take(4 * 32));
    f.u = base + o;
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
        for (int w = 0; w < kResolveWarps; ++w) { const int c = scratch[w]; if (w < warp) before += c; total += c; }
        if (i < n) a[i] = carry + before + incl - v;
        carry += total;
        __syncthreads();
    }
    return carry;
}

__device__ __forceinline__ void set_bit(unsigned long long* set, int p) {
    atomicOr(reinterpret_cast<unsigned*>(set) + (p >> 5), 1u << (p & 31));   // little endian halves
}
__device__ __forceinline__ bool get_bit(const unsigned long long* set, int p) {
    return (set[p >> 6] >> (p & 63)) & 1ull;
}

// word w of the dominator row of box p: from the compact cache when it is there, else from global
struct RowReader {
    const unsigned long long* cw;     // [kCompactWords * n] or nullptr
    const unsigned long long* dom;
    size_t stride;
    __device__ __forceinline__ unsigned long long get(int p, unsigned long long nz, int w) const {
        const int k = __popcll(nz & ((1ull << w) - 1ull));      // rank of word w among the non-zero ones
        if (cw && k < kCompactWords) return cw[(size_t)p * kCompactWords + k];
        return dom[(size_t)p * stride + w];
    }
};

template <int MODE, bool SLAB>
__device__ void resolve_fast(const NmsParams& P, int seg, long long off, int n, unsigned char* smem_raw, bool compact) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nw = cdiv(n, 64);      // <= 64 on this path
    FastSmem f;
    fast_carve(f, smem_raw, n, nw);
    const float unit = uses_shift(P) ? P.shift_unit[seg] : 0.f;
    const unsigned long long* dom = P.dom + (size_t)seg * (size_t)P.max_seg * (size_t)P.max_words;
    unsigned long long* cw = compact ? reinterpret_cast<unsigned long long*>(f.u) : nullptr;
    const int* perm = P.gperm + off;

    for (int w = tid; w < nw; w += kResolveThreads) { f.Kset[w] = 0ull; f.Rset[w] = 0ull; }
    __syncthreads();
    if (P.prof && tid == 0) P.prof[seg * 16 + 0] = clock64();
    // ---- A. stage boxes and the non-zero structure of the mask rows --------------------------------
    for (int p = tid; p < n; p += kResolveThreads) {
        const Item it = load_raw<SLAB>(P, off, perm[p], unit);
        f.box[p] = it.b; f.area[p] = it.area; f.key[p] = it.key; f.lab[p] = it.label;
        const unsigned long long* row = dom + (size_t)p * P.max_words;
        const unsigned long long nz = P.nzmask[off + p];
        if (cw && nz) {
            // first kCompactWords non-zero words, loads issued together
            int wi[kCompactWords];
            unsigned long long it2 = nz;
#pragma unroll
            for (int k = 0; k < kCompactWords; ++k) {
                wi[k] = it2 ? __ffsll((long long)it2) - 1 : -1;
                it2 &= it2 - 1ull;
            }
            unsigned long long d[kCompactWords];
#pragma unroll
            for (int k = 0; k < kCompactWords; ++k) d[k] = wi[k] >= 0 ? row[wi[k]] : 0ull;
#pragma unroll
            for (int k = 0; k < kCompactWords; ++k) cw[(size_t)p * kCompactWords + k] = d[k];
        }
        f.nz[p] = nz;
        if (!nz) atomicOr(reinterpret_cast<unsigned*>(f.Kset) + (p >> 5), 1u << (p & 31));   // no dominator: kept
    }
    __syncthreads();
    const RowReader R{cw, dom, (size_t)P.max_words};

    if (P.prof && tid == 0) P.prof[seg * 16 + 1] = clock64();
    // ---- B. fixed point over bitsets ----------------------------------------------------------------
    int pending;
    do {
        int undecided = 0;
        for (int p = tid; p < n; p += kResolveThreads) {
            if (get_bit(f.Kset, p) || get_bit(f.Rset, p)) continue;
            const unsigned long long nz = f.nz[p];
            unsigned long long hitK = 0ull, alive = 0ull, it = nz;
            while (it) {
                const int w = __ffsll((long long)it) - 1;
                it &= it - 1ull;
                const unsigned long long d = R.get(p, nz, w);
                hitK |= d & f.Kset[w];
                alive |= d & ~f.Rset[w];
            }
            if (hitK) set_bit(f.Rset, p);
            else if (!alive) set_bit(f.Kset, p);
            else ++undecided;
        }
        pending = __syncthreads_count(undecided > 0);
    } while (pending > 0);

    if (P.prof && tid == 0) P.prof[seg * 16 + 2] = clock64();
    if (MODE == B200_NMS_MAJORITY) {
        // ---- C. first suppressor + vote (helper.py:368-369), voters gathered per kept box ---------
        for (int p = tid; p <= n; p += kResolveThreads) f.voff[p] = 0;
        __syncthreads();
        for (int j = tid; j < n; j += kResolveThreads) {
            int s = -1;
            if (get_bit(f.Rset, j)) {
                const unsigned long long nz = f.nz[j];
                unsigned long long best = ~0ull, it = nz;
                int besti = -1;
                while (it) {
                    const int w = __ffsll((long long)it) - 1;
                    it &= it - 1ull;
                    unsigned long long d = R.get(j, nz, w) & f.Kset[w];
                    while (d) {
                        const int i = w * 64 + __ffsll((long long)d) - 1;
                        d &= d - 1ull;
                        if (f.key[i] < best) { best = f.key[i]; besti = i; }
                    }
                }
                bool vote = false;
                suppresses_exact<MODE>(P, f.box[besti], f.area[besti], f.box[j], f.area[j], &vote);
                s = besti | (vote ? kVoteFlag : 0);
                if (vote) atomicAdd(&f.voff[besti], 1);
            }
            f.sup[j] = s;
        }
        __syncthreads();
        if (P.prof && tid == 0) P.prof[seg * 16 + 8] = clock64();
        block_exclusive_scan(f.voff, n + 1, f.scan);
        if (P.prof && tid == 0) P.prof[seg * 16 + 9] = clock64();
        for (int p = tid; p < n; p += kResolveThreads) f.klist[p] = f.voff[p];    // fill cursors
        __syncthreads();
        for (int j = tid; j < n; j += kResolveThreads) {
            const int s = f.sup[j];
            if (s >= 0 && (s & kVoteFlag)) f.vlab[atomicAdd(&f.klist[s & ~kVoteFlag], 1)] = f.lab[j];
        }
        __syncthreads();
        if (P.prof && tid == 0) P.prof[seg * 16 + 10] = clock64();
        // majority relabel (helper.py:370-375): one warp per kept box that has at least two voters
        for (int p = warp; p < n; p += kResolveWarps) {
            if (!get_bit(f.Kset, p)) continue;
            const int v0 = f.voff[p], L = f.voff[p + 1] - v0;
            int label = f.lab[p];
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
            if (lane == 0) f.newlab[p] = label;
        }
        __syncthreads();
    }

    if (P.prof && tid == 0) P.prof[seg * 16 + 3] = clock64();
    // ---- D. kept boxes in (score desc, canonical index asc) order ------------------------------------
    // rank of a kept box among the kept ones = popcount of the kept bitset below it: one warp prefix-sums the
    // (<= 64) word popcounts, then every kept position places itself -- no block-wide scans
    if (warp == 0) {
        const int w0 = 2 * lane, w1 = 2 * lane + 1;
        const int c0 = w0 < nw ? __popcll(f.Kset[w0]) : 0, c1 = w1 < nw ? __popcll(f.Kset[w1]) : 0;
        int incl = c0 + c1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(kFullMask, incl, o); if (lane >= o) incl += u; }
        const int excl = incl - (c0 + c1);
        if (w0 < nw) f.voff[w0] = excl;                  // voff (n+1 ints) is free again after the vote phase
        if (w1 < nw) f.voff[w1] = excl + c0;
        if (lane == 31) f.scan[0] = incl;
    }
    __syncthreads();
    const int K = f.scan[0];
    for (int p = tid; p < n; p += kResolveThreads) {
        const unsigned long long word = f.Kset[p >> 6];
        if ((word >> (p & 63)) & 1ull) f.klist[f.voff[p >> 6] + __popcll(word & ((1ull << (p & 63)) - 1ull))] = p;
    }
    if (P.prof && tid == 0) P.prof[seg * 16 + 11] = clock64();
    const int Pk = next_pow2(K);
    unsigned long long* skey = reinterpret_cast<unsigned long long*>(f.u);      // compact rows are dead now
    int* sval = reinterpret_cast<int*>(skey + Pk);
    __syncthreads();
    if (K <= kRankSortMax) {
        // rank by counting: keys are unique, so the position of a kept box in the sorted order is the number of
        // kept boxes with a smaller key.  K broadcast reads per thread and ONE barrier instead of the
        // log2(K)^2 / 2 barrier-separated passes of a bitonic network (36 for K = 200).
        for (int t = tid; t < K; t += kResolveThreads) skey[t] = f.key[f.klist[t]];
        __syncthreads();
        if (P.prof && tid == 0) P.prof[seg * 16 + 12] = clock64();
        for (int t = tid; t < K; t += kResolveThreads) {
            const unsigned long long mine = skey[t];
            int rank = 0;
            for (int u = 0; u < K; ++u) rank += skey[u] < mine;
            sval[rank] = f.klist[t];
        }
        __syncthreads();
    } else {
        for (int t = tid; t < Pk; t += kResolveThreads) {
            skey[t] = t < K ? f.key[f.klist[t]] : ~0ull;
            sval[t] = t < K ? f.klist[t] : -1;
        }
        __syncthreads();
        bitonic_sort(skey, sval, Pk);
    }

    if (P.prof && tid == 0) P.prof[seg * 16 + 4] = clock64();
    // ---- E. emit -----------------------------------------------------------------------------------------
    const int Kout = SLAB ? min(K, P.max_det) : K;
    if (SLAB && P.det_keep) {
        // index in the reference's candidate list = rank of the flat anchor index among all candidates
        const int abits = P.anchor_space;                          // flat anchor indices are < abits
        const int awords = cdiv(abits, 32);
        if (abits > 0 && (size_t)awords * 8 <= 12 * (size_t)n) {
            // bitmap of the candidates' anchor indices + prefix popcounts: rank = #set bits below mine.
            // sup / voff / vlab (12n contiguous bytes) are free again.
            unsigned* bm = reinterpret_cast<unsigned*>(f.sup);
            int* pre = reinterpret_cast<int*>(bm + awords);
            for (int w = tid; w < awords; w += kResolveThreads) bm[w] = 0u;
            __syncthreads();
            for (int j = tid; j < n; j += kResolveThreads) {
                const unsigned a = (unsigned)f.key[j];
                atomicOr(&bm[a >> 5], 1u << (a & 31));
            }
            __syncthreads();
            for (int w = tid; w < awords; w += kResolveThreads) pre[w] = __popc(bm[w]);
            __syncthreads();
            block_exclusive_scan(pre, awords, f.scan);
            for (int t = tid; t < Kout; t += kResolveThreads) {
                const unsigned a = (unsigned)f.key[sval[t]];
                P.det_keep[(size_t)seg * P.max_det + t] = pre[a >> 5] + __popc(bm[a >> 5] & ((1u << (a & 31)) - 1u));
            }
        } else {
            // one warp per kept box, lanes sweep consecutive candidates (conflict-free shared reads)
            for (int t = warp; t < Kout; t += kResolveWarps) {
                const unsigned anchor = (unsigned)f.key[sval[t]];
                int cnt = 0;
                for (int j0 = 0; j0 < n; j0 += 32) {
                    const int j = j0 + lane;
                    cnt += __popc(__ballot_sync(kFullMask, j < n && (unsigned)f.key[j] < anchor));
                }
                if (lane == 0) P.det_keep[(size_t)seg * P.max_det + t] = cnt;
            }
        }
    }
    for (int t = tid; t < Kout; t += kResolveThreads) {
        const int p = sval[t];
        const int i = perm[p];
        const int lab = MODE == B200_NMS_MAJORITY ? f.newlab[p] : f.lab[p];
        if (SLAB) {
            const float4 b = reinterpret_cast<const float4*>(P.slab + off + i)[0];   // unshifted box
            float* d = P.det + ((size_t)seg * P.max_det + t) * 6;
            d[0] = b.x; d[1] = b.y; d[2] = b.z; d[3] = b.w;
            d[4] = from_orderable(~(unsigned)(f.key[p] >> 32));
            d[5] = (float)lab;
            if (P.det_anchor) P.det_anchor[(size_t)seg * P.max_det + t] = (int)(unsigned)f.key[p];
        } else {
            P.keep[off + t] = i;
            if (P.labels_out) P.labels_out[off + t] = lab;
        }
    }
    if (P.prof && tid == 0) P.prof[seg * 16 + 5] = clock64();
    if (P.prof && tid == 0) { P.prof[seg * 16 + 6] = n; P.prof[seg * 16 + 7] = K; }
    if (tid == 0) {
        if (SLAB) {
            P.det_count[seg] = Kout;
            if (K > P.max_det && P.status) atomicOr(P.status, 2);
        } else {
            P.keep_count[seg] = K;
        }
    }
}

// ------------------------------------------------------------------------------------------
// resolve, slow path: segments too large for shared memory work on global scratch
// ------------------------------------------------------------------------------------------
template <int MODE, bool SLAB>
__device__ void resolve_slow(const NmsParams& P, int seg, long long off, int n, unsigned char* smem_raw) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float unit = uses_shift(P) ? P.shift_unit[seg] : 0.f;
    const int nw = cdiv(n, 64);
    unsigned char* q = smem_raw;
    unsigned long long* Kset = reinterpret_cast<unsigned long long*>(q); q += sizeof(unsigned long long) * P.max_words;
    unsigned long long* Rset = reinterpret_cast<unsigned long long*>(q); q += sizeof(unsigned long long) * P.max_words;
    unsigned long long* skey = reinterpret_cast<unsigned long long*>(q); q += sizeof(unsigned long long) * kKeptSmem;
    int* sval = reinterpret_cast<int*>(q);                               q += sizeof(int) * kKeptSmem;
    int* sm_vote = reinterpret_cast<int*>(q);                            q += sizeof(int) * kResolveMaxWarps * kVoteListCap;
    int* sm_scan = reinterpret_cast<int*>(q);
    const unsigned long long* dom = P.dom + (size_t)seg * (size_t)P.max_seg * (size_t)P.max_words;
    const int* perm = P.gperm + off;
    for (int w = tid; w < nw; w += kResolveThreads) { Kset[w] = 0ull; Rset[w] = 0ull; }
    __syncthreads();

    int pending;
    do {
        int undecided = 0;
        for (int p = tid; p < n; p += kResolveThreads) {
            if (get_bit(Kset, p) || get_bit(Rset, p)) continue;
            const unsigned long long* row = dom + (size_t)p * P.max_words;
            const unsigned long long nz = nw <= 64 ? P.nzmask[off + p] : ~0ull;   // short rows: only flagged words are valid
            unsigned long long hitK = 0ull, alive = 0ull;
            for (int w = 0; w < nw; ++w) {
                const unsigned long long d = (nw > 64 || ((nz >> w) & 1ull)) ? row[w] : 0ull;
                hitK |= d & Kset[w];
                alive |= d & ~Rset[w];
            }
            if (hitK) set_bit(Rset, p);
            else if (!alive) set_bit(Kset, p);
            else ++undecided;
        }
        pending = __syncthreads_count(undecided > 0);
    } while (pending > 0);

    int running = 0;
    int* klist = P.gklist + off;
    for (int p0 = 0; p0 < n; p0 += kResolveThreads) {
        const int p = p0 + tid;
        const bool kept = p < n && get_bit(Kset, p);
        const int k = block_rank_rt(kept, sm_scan, running, kResolveWarps);
        if (kept) klist[k] = p;
    }
    const int K = running;
    __syncthreads();
    const int Pk = next_pow2(K);
    unsigned long long* key = Pk <= kKeptSmem ? skey : P.gkey + 2 * off;   // Pk < 2K <= 2n
    int* val = Pk <= kKeptSmem ? sval : P.gval + 2 * off;
    for (int t = tid; t < Pk; t += kResolveThreads) {
        if (t < K) { key[t] = load_item<SLAB>(P, off, klist[t], unit).key; val[t] = klist[t]; }
        else       { key[t] = ~0ull; val[t] = -1; }
    }
    __syncthreads();
    bitonic_sort(key, val, Pk);

    int* newlab = P.gnewlab + off;   // indexed by position
    if (MODE == B200_NMS_MAJORITY) {
        int* sup = P.gsup + off;
        for (int j = tid; j < n; j += kResolveThreads) {
            int s = -1;
            if (get_bit(Rset, j)) {
                const unsigned long long* row = dom + (size_t)j * P.max_words;
                const unsigned long long nz = nw <= 64 ? P.nzmask[off + j] : ~0ull;
                unsigned long long best = ~0ull;
                int besti = -1;
                for (int w = 0; w < nw; ++w) {
                    unsigned long long d = (nw > 64 || ((nz >> w) & 1ull)) ? (row[w] & Kset[w]) : 0ull;
                    while (d) {
                        const int i = w * 64 + __ffsll((long long)d) - 1;
                        d &= d - 1ull;
                        const unsigned long long k = load_item<SLAB>(P, off, i, unit).key;
                        if (k < best) { best = k; besti = i; }
                    }
                }
                const Item S = load_item<SLAB>(P, off, besti, unit);
                const Item T = load_item<SLAB>(P, off, j, unit);
                bool vote = false;
                suppresses_exact<MODE>(P, S.b, S.area, T.b, T.area, &vote);
                s = besti | (vote ? kVoteFlag : 0);
            }
            sup[j] = s;
        }
        __syncthreads();
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
                if (best_cnt < L) label = best_lab;
            }
            if (lane == 0) newlab[i] = label;
            __syncwarp();
        }
        __syncthreads();
    }

    for (int t = tid; t < K; t += kResolveThreads) {
        const int p = val[t];
        const int i = perm[p];
        if (SLAB) {
            if (t < P.max_det) {
                const Cand c = P.slab[off + i];
                float* d = P.det + ((size_t)seg * P.max_det + t) * 6;
                d[0] = c.x1; d[1] = c.y1; d[2] = c.x2; d[3] = c.y2;
                d[4] = c.score;
                d[5] = (float)(MODE == B200_NMS_MAJORITY ? newlab[p] : c.label);
                if (P.det_anchor) P.det_anchor[(size_t)seg * P.max_det + t] = c.anchor;
                if (P.det_keep) {
                    int rank = 0;
                    for (int j = 0; j < n; ++j) rank += (P.slab[off + j].anchor < c.anchor);
                    P.det_keep[(size_t)seg * P.max_det + t] = rank;
                }
            }
        } else {
            P.keep[off + t] = i;
            if (P.labels_out) P.labels_out[off + t] = MODE == B200_NMS_MAJORITY ? newlab[p] : (P.labels ? P.labels[off + i] : 0);
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

__host__ __device__ inline size_t slow_smem_bytes(int max_words) {
    return sizeof(unsigned long long) * 2 * (size_t)max_words + sizeof(unsigned long long) * kKeptSmem +
           sizeof(int) * kKeptSmem + sizeof(int) * kResolveMaxWarps * kVoteListCap + sizeof(int) * 32;
}

template <bool SLAB>
__global__ void __launch_bounds__(kResolveMaxThreads, RESOLVE_MINB)
k_nms_resolve(const __grid_constant__ NmsParams P, const unsigned smem_bytes) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int seg = blockIdx.x;
    long long off;
    int n, n_true;
    segment_range(P, seg, off, n, n_true);
    if (threadIdx.x == 0 && SLAB && P.cand_count_out) P.cand_count_out[seg] = n_true;
    if (n == 0) {
        if (threadIdx.x == 0) {
            if (P.keep_count) P.keep_count[seg] = 0;
            if (P.det_count) P.det_count[seg] = 0;
        }
        return;
    }
    if (n > P.split_hi) return;                   // resolved by the single-launch path
    const int nw = cdiv(n, 64);
    const size_t fixed = fast_fixed_bytes(n, nw);
    const size_t sort_bytes = 12 * (size_t)next_pow2(n);
    const size_t cw_bytes = 8 * (size_t)kCompactWords * (size_t)n;
    const bool fast = nw <= 64 && fixed + sort_bytes <= smem_bytes;
    const bool staged = fast && fixed + (cw_bytes > sort_bytes ? cw_bytes : sort_bytes) <= smem_bytes;
    if (P.mode == B200_NMS_MAJORITY) {
        if (fast) resolve_fast<B200_NMS_MAJORITY, SLAB>(P, seg, off, n, smem_raw, staged);
        else      resolve_slow<B200_NMS_MAJORITY, SLAB>(P, seg, off, n, smem_raw);
    } else {
        if (fast) resolve_fast<B200_NMS_TV, SLAB>(P, seg, off, n, smem_raw, staged);
        else      resolve_slow<B200_NMS_TV, SLAB>(P, seg, off, n, smem_raw);
    }
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
    const size_t tiles = words * (words + 1) / 2;
    Carve c{reinterpret_cast<unsigned char*>(base), 0, query};
    NmsParams tmp{};
    NmsParams& o = P ? *P : tmp;
    o.gkey = (unsigned long long*)c.take(16 * T);
    o.gval = (int*)c.take(8 * T);
    o.gperm = (int*)c.take(4 * T);
    o.nzmask = (unsigned long long*)c.take(8 * T);
    o.gsup = (int*)c.take(4 * T);
    o.gklist = (int*)c.take(4 * T);
    o.gnewlab = (int*)c.take(4 * T);
    o.shift_unit = (float*)c.take(4 * S);
    o.work_count = (int*)c.take(8);
    o.work = (int2*)c.take(8 * S * tiles);
    o.dom = (unsigned long long*)c.take(8 * S * max_seg * words);
    // single-launch path: arrival counters + one private rank-order slot per CTA
    const size_t FS = max_seg <= 4096 ? (size_t)nms_fused_slots() * max_seg : 1;
    o.f_ctl = (int*)c.take(4 * S);
    o.f_rbox = (float4*)c.take(16 * FS);
    o.f_rarea = (float*)c.take(4 * FS);
    o.f_rlabel = (int*)c.take(4 * FS);
    o.f_rkey = (unsigned long long*)c.take(8 * FS);
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

// optional profiling events (b200_debug_set_timeline): recorded after plan, pairs and resolve
cudaEvent_t g_nms_timeline[3] = {nullptr, nullptr, nullptr};
long long* g_resolve_prof = nullptr;

long long g_batched_nms_auto_limit = 100000;
int g_nms_force_general = -1;   // b200_debug_set_nms_path: -1 (default) = by workload, 1 = three-launch path, 0 = single-launch path
// launch shape of the resolve CTAs (b200_debug_set_resolve): 1024 threads at <= 32 registers and 112 KB leave room
// for a decode CTA on the same SM; 112 KB stage every segment of up to ~1200 boxes in shared memory
int g_resolve_threads = 1024;
int g_resolve_smem_kb = 112;   // torchvision 0.26 on CUDA (4000 on CPU)

int launch_nms(NmsParams& P, int num_segments, cudaStream_t stream) {
    P.auto_limit = g_batched_nms_auto_limit;
    if (num_segments <= 0) return B200_OK;
    if (num_segments > 65535) return B200_ERR_INVALID;
    if (P.max_seg < 1) P.max_seg = 1;
    P.max_words = cdiv(P.max_seg, 64);
    if (P.max_words > 65535) return B200_ERR_INVALID;      // tile coordinates are packed in 16 bits
    P.max_tiles = P.max_words * (P.max_words + 1) / 2;
    P.num_segments = num_segments;
    if (P.mode < 0) {
        if (!P.from_slab) return B200_ERR_INVALID;
        k_nms_canon<<<num_segments, kCanonThreads, 0, stream>>>(P);
        return cudaGetLastError() == cudaSuccess ? B200_OK : B200_ERR_CUDA;
    }
    // Default choice by workload: candidate slabs of the YOLO post-process run beside the streaming decode kernel of
    // the next batch, where the three small-footprint kernels of the general path co-reside with it and its spatially
    // pruned tile pairs do half the work; array inputs (torchvision nms / batched_nms, RPN levels, ROI heads) are
    // stand-alone calls on boxes spread all over the image, where one launch without a work queue wins.
    P.force_general = g_nms_force_general;
    P.prof = g_resolve_prof;
    P.split_lo = 0;
    P.split_hi = 0x7fffffff;
    const bool fused_ok = nms_fused_eligible(P);
    if (fused_ok && P.force_general == 2) {                 // debug: the 256-thread instantiation for every segment
        P.split_lo = -1;
        const int rc = launch_nms_fused(P, num_segments, stream);
        if (g_nms_timeline[2]) cudaEventRecord(g_nms_timeline[2], stream);
        return rc;
    }
    if (fused_ok && (P.force_general == 0 || (P.force_general < 0 && !P.from_slab))) {
        const int rc = launch_nms_fused(P, num_segments, stream);
        if (g_nms_timeline[2]) cudaEventRecord(g_nms_timeline[2], stream);
        return rc;
    }
    // Candidate slabs (default): both paths, split by segment size.  The three kernels below take the segments their
    // shared-memory resolve holds (<= kSplitBoxes boxes: the usual case, small CTAs that co-reside with the streaming
    // decode kernel of the next batch); the single-launch kernel takes the larger ones, which would otherwise fall
    // back to the global-memory resolve (20x slower).  Each kernel returns at once from segments that are not its own.
    const bool split = fused_ok && P.force_general < 0 && P.from_slab;
    const int split_at = P.serial ? (g_serial_split < kSplitBoxes ? g_serial_split : kSplitBoxes) : kSplitBoxes;
    if (split) P.split_hi = split_at;
    const int sms = current_sm_count();
    if (cudaMemsetAsync(P.work_count, 0, 2 * sizeof(int), stream) != cudaSuccess) return B200_ERR_CUDA;
    if (P.from_slab) k_nms_plan<true><<<num_segments, kPlanThreads, 0, stream>>>(P);
    else             k_nms_plan<false><<<num_segments, kPlanThreads, 0, stream>>>(P);
    if (g_nms_timeline[0]) cudaEventRecord(g_nms_timeline[0], stream);
    const int pair_ctas = 8 * sms;                       // 8 x 256 threads per SM, tiles pulled from a queue
    if (P.from_slab) k_nms_pairs<true><<<pair_ctas, kPairThreads, 0, stream>>>(P);
    else             k_nms_pairs<false><<<pair_ctas, kPairThreads, 0, stream>>>(P);

    if (g_nms_timeline[1]) cudaEventRecord(g_nms_timeline[1], stream);
    P.prof = g_resolve_prof;
    // Resolve CTAs (one per segment) are latency bound; they are kept small enough (threads, shared memory) to
    // share an SM with the streaming decode kernel of the next pipeline stage instead of waiting for it.
    const int resolve_threads = g_resolve_threads;
    const size_t resolve_smem = (size_t)g_resolve_smem_kb * 1024;
    const size_t slow_bytes = slow_smem_bytes(P.max_words);
    const size_t smem = resolve_smem > slow_bytes ? resolve_smem : slow_bytes;
    if (smem > 227 * 1024) return B200_ERR_INVALID;   // max_seg beyond ~1.4M boxes
    static SmemOptIn optin[2];
    if ((P.from_slab ? optin[1].ensure(k_nms_resolve<true>, smem) : optin[0].ensure(k_nms_resolve<false>, smem)) != cudaSuccess)
        return B200_ERR_CUDA;
    if (P.from_slab) k_nms_resolve<true><<<num_segments, resolve_threads, smem, stream>>>(P, (unsigned)smem);
    else             k_nms_resolve<false><<<num_segments, resolve_threads, smem, stream>>>(P, (unsigned)smem);
    if (g_nms_timeline[2]) cudaEventRecord(g_nms_timeline[2], stream);
    if (cudaGetLastError() != cudaSuccess) return B200_ERR_CUDA;
    if (split) {
        P.split_lo = split_at;
        P.split_hi = 0x7fffffff;
        return launch_nms_fused(P, num_segments, stream);
    }
    return B200_OK;
}

}  // namespace b200
