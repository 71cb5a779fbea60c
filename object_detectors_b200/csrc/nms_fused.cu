// nms_fused.cu -- segmented greedy NMS in ONE launch for segments of <= 4096 boxes (sm_100a).
//
// The general path (nms.cu) needs three dependent launches (plan -> pairs -> resolve), a global work queue and a
// fixed-point resolve whose latency is set by the heaviest segment.  Post-processing batches are many small segments
// (C2: 64 images x ~640 candidates, RPN: 80 levels x <= 2000), so latency -- not arithmetic -- is what the NMS stage
// costs.  Here one grid of (at most) one CTA per SM does everything without any CTA ever waiting for another one:
//
//   TEAMS   Every CTA reads the segment sizes and derives the same assignment: segment s gets a team of g_s CTAs,
//           g_s proportional to its pair count (largest-remainder rounding, so every CTA of the grid is used).  With
//           more segments than CTAs every CTA owns whole segments.
//   RANK    Each team member sorts the segment by (score desc, index asc) -- bucket sort, bitonic network as fallback -- and
//           writes the boxes in RANK order to its private scratch.  The work is redundant inside a team (a few
//           microseconds) and buys independence: nobody waits for a "planner".
//   STRIPS  The dominator bitmask lives in rank space: bit q of row r <=> q precedes r and suppresses it, so only
//           words up to the diagonal exist and row tile rt needs column tiles 0..rt.  The nt*(nt+1)/2 tiles are one
//           flat list; member m evaluates a contiguous 1/g share of it.  A cheap exact-safe prefilter
//           (IoU <= inter / max(area)) collects candidates, the reference's own operation-by-operation IoU test
//           (nms_dev.cuh) decides them.  Every mask word has exactly one writer, so words are plain 64-bit stores: no
//           zero-initialisation and no global atomics.
//   RESOLVE The member that arrives last at the team counter (fence + atomicAdd, nobody spins) walks the rows in
//           blocks of 64: words of earlier blocks are tested in parallel against the kept set (register-prefetched
//           from L2 two blocks ahead), the diagonal word serially over just the rows that have one; first
//           suppressors fall out of the same pass.  Majority relabel (helper.py:368-375) and emission follow in rank
//           order -- no second sort.
#include "nms_dev.cuh"

namespace b200 {

// Threads per CTA is a template parameter NT: 1024 for stand-alone calls (array inputs), 256 when the kernel only takes
// the large segments of a candidate slab next to the general path (it then has to start beside the streaming decode
// kernel of the next batch: 256 threads x 64 registers and 41 KB co-reside with it, 1024 threads need the whole SM).
// NT / 64 threads share one row of a row tile; the same number of column tiles is staged per batch.
static constexpr int kMaxParts = 16;
static constexpr int kHeavyCap = 256;           // kept boxes with many voters, relabelled by a whole warp
static constexpr int kHeavyVoters = 12;
static constexpr int kPoolInts = 9216;          // resolve scratch in shared memory (36 KB)
static constexpr int kSmemBytes = 40 * 1024;
static constexpr int kNone = 0x7fffffff;
static constexpr int kFVoteFlag = 1 << 30;

__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// in-place exclusive scan of a[0..n) by the whole CTA (a may live in shared or global memory)
template <int NT>
__device__ __forceinline__ void f_exclusive_scan(int* a, int n, int* scratch) {
    constexpr int kFT = NT, kFW = NT / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int carry = 0;
    for (int base = 0; base < n; base += kFT) {
        const int i = base + tid;
        const int v = i < n ? a[i] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(kFullMask, incl, o); if (lane >= o) incl += u; }
        if (lane == 31) scratch[warp] = incl;
        __syncthreads();
        int before = 0, total = 0;
#pragma unroll
        for (int w = 0; w < kFW; ++w) { const int c = scratch[w]; if (w < warp) before += c; total += c; }
        if (i < n) a[i] = carry + before + incl - v;
        carry += total;
        __syncthreads();
    }
}

// ascending bitonic sort of key[0..P2) in shared memory, P2 a power of two >= 64.  Steps with a partner distance
// below 64 stay inside aligned blocks of 64 keys and are done by one warp per block without CTA barriers.
template <int NT>
__device__ __forceinline__ void f_bitonic(unsigned long long* key, int P2) {
    constexpr int kFT = NT, kFW = NT / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int base = warp * 64; base < P2; base += kFW * 64) {
        unsigned long long* kk = key + base;
        for (int k = 2; k <= 64; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                const int i = ((lane & ~(j - 1)) << 1) | (lane & (j - 1));
                const int ixj = i | j;
                const unsigned long long a = kk[i], b = kk[ixj];
                // direction from the GLOBAL index so that runs alternate as the later merges expect
                if ((a > b) == (((base + i) & k) == 0)) { kk[i] = b; kk[ixj] = a; }
                __syncwarp();
            }
        }
    }
    __syncthreads();
    for (int k = 128; k <= P2; k <<= 1) {
        for (int j = k >> 1; j >= 64; j >>= 1) {
            for (int t = tid; t < (P2 >> 1); t += kFT) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int ixj = i | j;
                const unsigned long long a = key[i], b = key[ixj];
                if ((a > b) == ((i & k) == 0)) { key[i] = b; key[ixj] = a; }
            }
            __syncthreads();
        }
        for (int base = warp * 64; base < P2; base += kFW * 64) {
            unsigned long long* kk = key + base;
            const bool asc = (base & k) == 0;
            for (int j = 32; j > 0; j >>= 1) {
                const int i = ((lane & ~(j - 1)) << 1) | (lane & (j - 1));
                const int ixj = i | j;
                const unsigned long long a = kk[i], b = kk[ixj];
                if ((a > b) == asc) { kk[i] = b; kk[ixj] = a; }
                __syncwarp();
            }
        }
        __syncthreads();
    }
}

// Sort by buckets instead of a sorting network (the idea of rpn.cu's bucket_sort_topk, whole array): a histogram over
// NB buckets linear in the SCORE between the best and the worst key -- a monotone map, so bucket order is key order --
// a scan, a scatter into `tmp` (this CTA's private global scratch) and a rank-by-counting inside each (short) bucket
// leave key[0..n) ascending, exactly as the network would (keys are unique).  Seven barriers instead of 66 steps for
// 2048 keys.  Returns false -- key[] untouched -- when the scores defeat the buckets (non-finite, all equal, a run of
// more than kFBucketRun equal-ish scores): the caller then runs the network.
static constexpr int kFBucketRun = 256;

template <int NT, int NB>
__device__ bool f_bucket_sort(unsigned long long* key, int n, unsigned long long* tmp, int* start, int* fill, int* sc) {
    constexpr int kFW = NT / 32, kPer = NB / NT;
    static_assert(NB % NT == 0 && kPer >= 1, "buckets per thread");
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned vmin = ~0u, vmax = 0u;
    for (int i = tid; i < n; i += NT) {
        const unsigned v = (unsigned)(key[i] >> 32);
        vmin = min(vmin, v); vmax = max(vmax, v);
    }
    vmin = __reduce_min_sync(kFullMask, vmin);
    vmax = __reduce_max_sync(kFullMask, vmax);
    __syncthreads();
    if (lane == 0) { sc[warp] = (int)vmin; start[warp] = (int)vmax; }
    __syncthreads();
    for (int w = 0; w < kFW; ++w) { vmin = min(vmin, (unsigned)sc[w]); vmax = max(vmax, (unsigned)start[w]); }
    const float xbest = from_orderable(~vmin), xworst = from_orderable(~vmax);
    const float scale = __fdiv_rn((float)(NB - 1), __fsub_rn(xbest, xworst));
    if (!(xbest > xworst) || !(fabsf(xbest) < 3.0e38f) || !(fabsf(xworst) < 3.0e38f) || !(scale < 3.0e38f)) return false;
    auto bucket_of = [&](unsigned long long k) {
        const float t = __fmul_rn(__fsub_rn(xbest, from_orderable(~(unsigned)(k >> 32))), scale);
        return min(NB - 1, (int)t);
    };
    __syncthreads();                                                   // sc / start are reused below
    for (int i = tid; i < NB; i += NT) fill[i] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += NT) atomicAdd(&fill[bucket_of(key[i])], 1);
    __syncthreads();
    // exclusive scan of the counts, kPer consecutive buckets per thread
    int cnt[kPer], sum = 0, too_long = 0;
#pragma unroll
    for (int j = 0; j < kPer; ++j) { cnt[j] = fill[tid * kPer + j]; sum += cnt[j]; too_long |= cnt[j] > kFBucketRun; }
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(kFullMask, incl, o); if (lane >= o) incl += u; }
    if (lane == 31) sc[warp] = incl;
    const bool fail = __syncthreads_or(too_long) != 0;
    if (fail) return false;
    int at = incl - sum;
    for (int w = 0; w < warp; ++w) at += sc[w];
#pragma unroll
    for (int j = 0; j < kPer; ++j) { start[tid * kPer + j] = at; at += cnt[j]; fill[tid * kPer + j] = 0; }
    __syncthreads();
    for (int i = tid; i < n; i += NT) {
        const unsigned long long k = key[i];
        const int b = bucket_of(k);
        tmp[start[b] + atomicAdd(&fill[b], 1)] = k;
    }
    __syncthreads();
    for (int p = tid; p < n; p += NT) {
        const unsigned long long k = tmp[p];
        const int b = bucket_of(k);
        const int s0 = start[b], s1 = b + 1 < NB ? start[b + 1] : n;
        int r = 0;
        for (int q = s0; q < s1; ++q) r += tmp[q] < k;
        key[s0 + r] = k;
    }
    __syncthreads();
    return true;
}

// private scratch of one CTA: the current segment in rank order
struct RankArrays {
    float4* box;
    float* area;
    int* label;
    unsigned long long* key;
};

// ---------------------------------------------------------------------------------------------- RANK
// returns true when the segment holds a degenerate box (zero / negative / non-finite area, NaN coordinates): such
// boxes can yield NaN IoU, which the majority rule treats as "removed", so the segment never uses the prefilter
template <bool SLAB, int NT>
__device__ bool fused_rank(const NmsParams& P, int seg, long long off, int n, int n_true, const RankArrays& R,
                           unsigned char* smem, float* red, bool first_member) {
    constexpr int kFT = NT;
    const int tid = threadIdx.x;
    // ---- coordinate-trick unit (torchvision.ops.batched_nms: boxes + idxs * (boxes.max() + 1)) ------------------
    float unit = 0.f;
    if (P.given_unit) {
        unit = P.given_unit[seg];
    } else if (P.mode == B200_NMS_TV_TRICK || (P.mode == B200_NMS_TV_AUTO && 4ll * n_true <= P.auto_limit)) {
        float mx = -INFINITY;
        for (int i = tid; i < n; i += kFT) {
            const float4 b = SLAB ? reinterpret_cast<const float4*>(P.slab + off + i)[0]
                                  : reinterpret_cast<const float4*>(P.boxes)[off + i];
            mx = fmaxf(mx, fmaxf(fmaxf(b.x, b.y), fmaxf(b.z, b.w)));
        }
        unit = __fadd_rn(block_reduce_max(mx, red), 1.0f);
    }
    if (uses_shift(P) && tid == 0 && first_member) P.shift_unit[seg] = unit;

    // ---- sort (score desc, canonical index asc); the low 12 bits carry the index inside the segment ------------
    unsigned long long* key = reinterpret_cast<unsigned long long*>(smem);
    const int P2 = max(next_pow2(n), 64);
    for (int i = tid; i < P2; i += kFT) {
        unsigned long long k = ~0ull;
        if (i < n) {
            if (SLAB) {
                const float4 m = reinterpret_cast<const float4*>(P.slab + off + i)[1];
                k = ((unsigned long long)(~orderable(m.x)) << 32) | ((unsigned long long)(unsigned)__float_as_int(m.z) << 12) | (unsigned)i;
            } else {
                k = ((unsigned long long)(~orderable(P.scores[off + i])) << 32) | (unsigned)i;
            }
        }
        key[i] = k;
    }
    __syncthreads();
    bool sorted = false;
    {
        // the 40 KB scratch holds the keys and the bucket counters: 16 + 16 KB up to 2048 boxes, 32 + 8 KB up to 4096
        int* sc = reinterpret_cast<int*>(red);
        if (n >= 256 && n <= 2048)
            sorted = f_bucket_sort<NT, 2048>(key, n, R.key, reinterpret_cast<int*>(smem + 16384), reinterpret_cast<int*>(smem + 16384) + 2048, sc);
        else if (n > 2048 && n <= 4096)
            sorted = f_bucket_sort<NT, 1024>(key, n, R.key, reinterpret_cast<int*>(smem + 32768), reinterpret_cast<int*>(smem + 32768) + 1024, sc);
    }
    if (!sorted) f_bitonic<NT>(key, P2);
    int bad = 0;
    for (int r = tid; r < n; r += kFT) {
        const unsigned long long k = key[r];
        const Item it = load_raw<SLAB>(P, off, (int)(k & 0xfffu), unit);
        R.box[r] = it.b;
        R.area[r] = it.area;
        R.label[r] = it.label;
        R.key[r] = k;
        bad |= !(it.area > 0.f) || !(it.area < 3.0e38f) || !(it.b.x == it.b.x) || !(it.b.y == it.b.y) ||
               !(it.b.z == it.b.z) || !(it.b.w == it.b.w);
    }
    return __syncthreads_or(bad) != 0;
}

// ---------------------------------------------------------------------------------------------- STRIPS
struct StripSmem {
    float4 cb[kMaxParts * 64];
    float ca[kMaxParts * 64];
    float cta[kMaxParts * 64];                   // 0.999 * thr * area (prefilter)
    int clab[kMaxParts * 64];
    unsigned wbuf[64 * kMaxParts * 2];           // [64 rows][NT / 64 tiles] 64-bit words as 32-bit halves
};
static_assert(sizeof(StripSmem) <= kSmemBytes, "strip staging must fit");

// row tile rt against column tiles [c0, c1), c1 <= rt + 1
template <int MODE, int NT>
__device__ void fused_strip(const NmsParams& P, int seg, int n, int rt, int c0, int c1, const RankArrays& R,
                            unsigned char* smem, bool nofilter) {
    constexpr int kParts = NT / 64, kBatchTiles = kParts;
    StripSmem& S = *reinterpret_cast<StripSmem*>(smem);
    const int tid = threadIdx.x;
    unsigned long long* dom = P.dom + (size_t)seg * (size_t)P.max_seg * (size_t)P.max_words;
    const size_t W = (size_t)P.max_words;
    const int rr = tid & 63, part = tid >> 6;
    const int ri = rt * 64 + rr;
    const bool vi = ri < n;
    const float fthr = 0.999f * P.thr_f;
    float4 bi = make_float4(0.f, 0.f, 0.f, 0.f);
    float ai = 0.f;
    int li = 0;
    if (vi) { bi = R.box[ri]; ai = R.area[ri]; li = R.label[ri]; }
    const float tai = fthr * ai;
    for (int b0 = c0; b0 < c1; b0 += kBatchTiles) {
        const int ntl = min(kBatchTiles, c1 - b0);
        __syncthreads();                       // previous batch fully consumed
        {
            const int q = b0 * 64 + tid;       // column rank staged by this thread
            if (tid < ntl * 64 && q < n) {
                const float a = R.area[q];
                S.cb[tid] = R.box[q]; S.ca[tid] = a; S.cta[tid] = fthr * a; S.clab[tid] = R.label[q];
            }
            S.wbuf[2 * tid] = 0u;
            S.wbuf[2 * tid + 1] = 0u;
        }
        __syncthreads();
        if (vi) {
            // the batch's ntl*64 columns are split evenly over the 16 threads that share a row; only columns
            // that precede the row (q < ri) form a pair
            const int per = ntl * (64 / kParts), s0 = part * per;
            const int lim = ri - (b0 * 64 + s0);
            if (lim > 0) {
                unsigned long long cand = 0ull;
                if (nofilter) {
                    cand = ~0ull;
                } else {
                    for (int u0 = 0; u0 < per; u0 += 4) {
                        unsigned m = 0u;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const float4 bj = S.cb[s0 + u0 + k];
                            // IoU <= inter / max(area): a pair whose intersection is clearly below thr * max(area)
                            // cannot suppress.  The margin (1e-3) dwarfs every fp32 rounding.  One clamp is enough:
                            // with w >= 0 a negative h makes the product <= 0, below the (positive) bound.
                            const float w = fmaxf(fminf(bi.z, bj.z) - fmaxf(bi.x, bj.x), 0.f);
                            const float h = fminf(bi.w, bj.w) - fmaxf(bi.y, bj.y);
                            if (!(w * h < fmaxf(tai, S.cta[s0 + u0 + k]))) m |= 1u << k;
                        }
                        cand |= (unsigned long long)m << u0;
                    }
                }
                if (lim < 64) cand &= (1ull << lim) - 1ull;
                if (per < 64) cand &= (1ull << per) - 1ull;
                while (cand) {
                    const int s = s0 + __ffsll((long long)cand) - 1;
                    cand &= cand - 1ull;
                    // the column precedes the row: S = column (picked), T = row (remaining)
                    if (pair_hit<MODE>(P, bi, ai, li, S.cb[s], S.ca[s], S.clab[s], false))
                        atomicOr(&S.wbuf[2 * (rr * kBatchTiles + (s >> 6)) + ((s >> 5) & 1)], 1u << (s & 31));
                }
            }
        }
        __syncthreads();
        {
            const int row = tid / kBatchTiles, k = tid % kBatchTiles;
            const int r = rt * 64 + row;
            if (k < ntl && r < n)
                dom[(size_t)r * W + b0 + k] = ((unsigned long long)S.wbuf[2 * tid + 1] << 32) | S.wbuf[2 * tid];
        }
    }
}

// ---------------------------------------------------------------------------------------------- RESOLVE
struct ResolveSmem {
    unsigned long long Kept[64];
    unsigned long long diag[64];
    int esup[64];
    int kpre[64];
    int heavy_n;
    int pad[3];
    int pool[kPoolInts];
};
static_assert(sizeof(ResolveSmem) <= kSmemBytes, "resolve scratch must fit");

// blocks of 64 rows in rank order.  Thread (row = tid / P, g = tid % P), P = NT / 64, holds words g + P*k (k < NK) of its row;
// three register sets rotate so that the words of block b + 2 are in flight while block b is decided.
template <int NK, bool MAJ, int NT>
__device__ __noinline__ void resolve_blocks(const unsigned long long* __restrict__ dom, size_t W, int n, int nw,
                                            ResolveSmem& M, int* sup) {
    constexpr int kParts = NT / 64;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int row = tid / kParts, g = tid % kParts;
    struct Set { unsigned long long v[NK]; };
    auto load = [&](int blk, Set& s) {
        const int r = blk * 64 + row;
#pragma unroll
        for (int k = 0; k < NK; ++k) {
            const int w = g + kParts * k;
            s.v[k] = (r < n && w <= blk) ? __ldcg(dom + (size_t)r * W + w) : 0ull;
        }
    };
    auto decide = [&](int blk, const Set& s) {
        int first = kNone;
#pragma unroll
        for (int k = 0; k < NK; ++k) {
            const int w = g + kParts * k;
            if (w < blk) {
                const unsigned long long h = s.v[k] & M.Kept[w];
                if (h) first = min(first, w * 64 + __ffsll((long long)h) - 1);
            } else if (w == blk) {
                M.diag[row] = s.v[k];
            }
        }
#pragma unroll
        for (int o = 1; o < kParts; o <<= 1) first = min(first, __shfl_xor_sync(kFullMask, first, o));
        if (g == 0) M.esup[row] = first;
        __syncthreads();
        if (warp == 0) {
            const int r0 = blk * 64 + lane, r1 = r0 + 32;
            const bool v0 = r0 < n, v1 = r1 < n;
            const int e0 = M.esup[lane], e1 = M.esup[lane + 32];
            const unsigned long long d0 = v0 ? M.diag[lane] : 0ull, d1 = v1 ? M.diag[lane + 32] : 0ull;
            const unsigned long long valid = ((unsigned long long)__ballot_sync(kFullMask, v1) << 32) | __ballot_sync(kFullMask, v0);
            const unsigned long long ext = ((unsigned long long)__ballot_sync(kFullMask, v1 && e1 != kNone) << 32) |
                                           __ballot_sync(kFullMask, v0 && e0 != kNone);
            const unsigned long long hd = ((unsigned long long)__ballot_sync(kFullMask, d1 != 0ull) << 32) |
                                          __ballot_sync(kFullMask, d0 != 0ull);
            unsigned long long keptw = valid & ~ext & ~hd;
            unsigned long long pend = valid & ~ext & hd;
            int ds0 = -1, ds1 = -1;
            while (pend) {                              // uniform across the warp: rows that have a diagonal word
                const int k = __ffsll((long long)pend) - 1;
                pend &= pend - 1ull;
                const unsigned long long da = __shfl_sync(kFullMask, d0, k & 31), db = __shfl_sync(kFullMask, d1, k & 31);
                const unsigned long long hit = (k < 32 ? da : db) & keptw;
                if (!hit) keptw |= 1ull << k;
                else if (lane == (k & 31)) {
                    const int sp = blk * 64 + __ffsll((long long)hit) - 1;
                    if (k < 32) ds0 = sp; else ds1 = sp;
                }
            }
            if (lane == 0) M.Kept[blk] = keptw;
            if (MAJ) {
                if (v0) sup[r0] = e0 != kNone ? e0 : ds0;
                if (v1) sup[r1] = e1 != kNone ? e1 : ds1;
            }
        }
        __syncthreads();
    };
    if (NK > 4) {
        // many words per thread (long rows handled by few threads): no register rotation, the words of a block are
        // loaded when the block is decided
        for (int blk = 0; blk < nw; ++blk) {
            Set A;
            load(blk, A);
            decide(blk, A);
        }
        return;
    }
    Set A, B, C;
    load(0, A);
    load(1, B);
    for (int blk = 0; blk < nw; blk += 3) {
        load(blk + 2, C);
        decide(blk, A);
        if (blk + 1 < nw) { load(blk + 3, A); decide(blk + 1, B); }
        if (blk + 2 < nw) { load(blk + 4, B); decide(blk + 2, C); }
    }
}

template <int MODE, bool SLAB, int NT>
__device__ void fused_resolve(const NmsParams& P, int seg, long long off, int n, const RankArrays& R,
                              unsigned char* smem, int* scan, unsigned long long* prof) {
    constexpr int kFT = NT, kFW = NT / 32, kParts = NT / 64;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nw = cdiv(n, 64);
    const size_t W = (size_t)P.max_words;
    const unsigned long long* dom = P.dom + (size_t)seg * (size_t)P.max_seg * W;
    ResolveSmem& M = *reinterpret_cast<ResolveSmem*>(smem);
    int* pool = M.pool;
    int* sup = pool;                                                                  // [n] (MAJORITY)
    constexpr bool MAJ = MODE == B200_NMS_MAJORITY;
    if (tid == 0) M.heavy_n = 0;
    if (nw <= kParts)          resolve_blocks<1, MAJ, NT>(dom, W, n, nw, M, sup);
    else if (nw <= 2 * kParts) resolve_blocks<2, MAJ, NT>(dom, W, n, nw, M, sup);
    else if (nw <= 4 * kParts) resolve_blocks<4, MAJ, NT>(dom, W, n, nw, M, sup);
    else                       resolve_blocks<64 / kParts, MAJ, NT>(dom, W, n, nw, M, sup);
    if (prof && tid == 0) prof[5] = gtime();

    // ---- output slot of a kept rank = number of kept ranks below it ------------------------------------------------
    if (warp == 0) {
        const int w0 = 2 * lane, w1 = 2 * lane + 1;
        const int c0 = w0 < nw ? __popcll(M.Kept[w0]) : 0, c1 = w1 < nw ? __popcll(M.Kept[w1]) : 0;
        int incl = c0 + c1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(kFullMask, incl, o); if (lane >= o) incl += u; }
        const int excl = incl - (c0 + c1);
        M.kpre[w0] = excl;
        M.kpre[w1] = excl + c0;
        if (lane == 31) scan[31] = incl;
    }
    __syncthreads();
    const int K = scan[31];
    const int Kout = SLAB ? min(K, P.max_det) : K;
    auto is_kept = [&](int r) { return ((M.Kept[r >> 6] >> (r & 63)) & 1ull) != 0ull; };
    auto slot_of = [&](int r) {
        const unsigned long long word = M.Kept[r >> 6];
        return M.kpre[r >> 6] + __popcll(word & ((1ull << (r & 63)) - 1ull));
    };
    auto emit = [&](int r, int t, int lab) {
        if (t >= Kout) return;
        const unsigned long long k = R.key[r];
        const int i = (int)(k & 0xfffu);
        if (SLAB) {
            const float4 b = reinterpret_cast<const float4*>(P.slab + off + i)[0];     // unshifted box
            float* d = P.det + ((size_t)seg * P.max_det + t) * 6;
            d[0] = b.x; d[1] = b.y; d[2] = b.z; d[3] = b.w;
            d[4] = from_orderable(~(unsigned)(k >> 32));
            d[5] = (float)lab;
            if (P.det_anchor) P.det_anchor[(size_t)seg * P.max_det + t] = (int)((unsigned)k >> 12);
        } else {
            P.keep[off + t] = i;
            if (P.labels_out) P.labels_out[off + t] = lab;
        }
    };

    if (MAJ) {
        // ---- first suppressor's vote (helper.py:368-369), voters gathered per kept box, majority relabel --------------
        const bool in_smem = 4 * n + 1 + kHeavyCap <= kPoolInts;
        int* voff = pool + n;                                               // [n+1]
        int* vlab = in_smem ? pool + 2 * n + 1 : P.gsup + off;              // [n]
        int* fill = in_smem ? pool + 3 * n + 1 : P.gklist + off;            // [n] scatter cursors, then the new labels
        int* heavy = in_smem ? pool + 4 * n + 1 : pool + 2 * n + 1;         // [kHeavyCap]
        for (int p = tid; p <= n; p += kFT) voff[p] = 0;
        __syncthreads();
        for (int j = tid; j < n; j += kFT) {
            const int s = sup[j];
            if (s >= 0) {
                bool vote = false;
                suppresses_exact<MODE>(P, R.box[s], R.area[s], R.box[j], R.area[j], &vote);
                if (vote) { sup[j] = s | kFVoteFlag; atomicAdd(&voff[s], 1); }
            }
        }
        __syncthreads();
        f_exclusive_scan<NT>(voff, n + 1, scan);
        for (int p = tid; p < n; p += kFT) fill[p] = voff[p];
        __syncthreads();
        for (int j = tid; j < n; j += kFT) {
            const int s = sup[j];
            if (s >= 0 && (s & kFVoteFlag)) vlab[atomicAdd(&fill[s & ~kFVoteFlag], 1)] = R.label[j];
        }
        __syncthreads();
        if (prof && tid == 0) prof[6] = gtime();
        // majority label of a voter list: most frequent class, smallest class id on ties (unique()/argmax)
        for (int r = tid; r < n; r += kFT) {
            if (!is_kept(r)) continue;
            const int v0 = voff[r], L = voff[r + 1] - v0;
            int label = R.label[r];
            if (L >= 2) {
                int slot = kHeavyCap;
                if (L > kHeavyVoters) slot = atomicAdd(&M.heavy_n, 1);
                if (slot < kHeavyCap) {
                    heavy[slot] = r;                     // a whole warp counts this one below
                } else {
                    int best_cnt = 0, best_lab = 0x7fffffff;
                    for (int a = 0; a < L; ++a) {
                        const int la = vlab[v0 + a];
                        int cnt = 0;
                        for (int b = 0; b < L; ++b) cnt += (vlab[v0 + b] == la);
                        if (cnt > best_cnt || (cnt == best_cnt && la < best_lab)) { best_cnt = cnt; best_lab = la; }
                    }
                    if (best_cnt < L) label = best_lab;   // more than one distinct class among the voters
                }
            }
            fill[r] = label;
        }
        __syncthreads();
        const int nh = min(M.heavy_n, kHeavyCap);
        for (int hI = warp; hI < nh; hI += kFW) {
            const int r = heavy[hI];
            const int v0 = voff[r], L = voff[r + 1] - v0;
            int best_cnt = 0, best_lab = 0x7fffffff;
            for (int a = lane; a < L; a += 32) {
                const int la = vlab[v0 + a];
                int cnt = 0;
                for (int b = 0; b < L; ++b) cnt += (vlab[v0 + b] == la);
                if (cnt > best_cnt || (cnt == best_cnt && la < best_lab)) { best_cnt = cnt; best_lab = la; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const int oc = __shfl_xor_sync(kFullMask, best_cnt, o);
                const int ol = __shfl_xor_sync(kFullMask, best_lab, o);
                if (oc > best_cnt || (oc == best_cnt && ol < best_lab)) { best_cnt = oc; best_lab = ol; }
            }
            if (lane == 0 && best_cnt < L) fill[r] = best_lab;
        }
        __syncthreads();
        for (int r = tid; r < n; r += kFT)
            if (is_kept(r)) emit(r, slot_of(r), fill[r]);
    } else {
        for (int r = tid; r < n; r += kFT)
            if (is_kept(r)) emit(r, slot_of(r), R.label[r]);
    }
    if (prof && tid == 0) prof[7] = gtime();

    // ---- index in the reference's candidate list = rank of the flat anchor index among all candidates ------------
    if (SLAB && P.det_keep) {
        __syncthreads();                                   // pool is free again
        const int abits = P.anchor_space, awords = cdiv(abits, 32);
        if (abits > 0 && 2 * awords <= kPoolInts) {
            unsigned* bm = reinterpret_cast<unsigned*>(pool);
            int* pre = pool + awords;
            for (int w = tid; w < awords; w += kFT) bm[w] = 0u;
            __syncthreads();
            for (int j = tid; j < n; j += kFT) {
                const unsigned a = (unsigned)R.key[j] >> 12;
                atomicOr(&bm[a >> 5], 1u << (a & 31));
            }
            __syncthreads();
            for (int w = tid; w < awords; w += kFT) pre[w] = __popc(bm[w]);
            __syncthreads();
            f_exclusive_scan<NT>(pre, awords, scan);
            for (int r = tid; r < n; r += kFT) {
                if (!is_kept(r)) continue;
                const int t = slot_of(r);
                if (t >= Kout) continue;
                const unsigned a = (unsigned)R.key[r] >> 12;
                P.det_keep[(size_t)seg * P.max_det + t] = pre[a >> 5] + __popc(bm[a >> 5] & ((1u << (a & 31)) - 1u));
            }
        } else {
            for (int r = warp; r < n; r += kFW) {
                if (!is_kept(r)) continue;
                const int t = slot_of(r);
                if (t >= Kout) continue;
                const unsigned a = (unsigned)R.key[r] >> 12;
                int cnt = 0;
                for (int j0 = 0; j0 < n; j0 += 32) {
                    const int j = j0 + lane;
                    cnt += __popc(__ballot_sync(kFullMask, j < n && ((unsigned)R.key[j] >> 12) < a));
                }
                if (lane == 0) P.det_keep[(size_t)seg * P.max_det + t] = cnt;
            }
        }
    }
    if (tid == 0) {
        if (SLAB) {
            P.det_count[seg] = Kout;
            if (K > P.max_det && P.status) atomicOr(P.status, 2);
        } else {
            P.keep_count[seg] = K;
        }
    }
}

// ---------------------------------------------------------------------------------------------- one segment
template <bool SLAB, int NT>
__device__ void fused_segment(const NmsParams& P, int seg, int member, int team, const RankArrays& R,
                              unsigned char* smem, float* red, int* scan, int* cmd) {
    const int tid = threadIdx.x;
    long long off;
    int n, n_true;
    segment_range(P, seg, off, n, n_true);
    if (P.split_lo > 0 && n <= P.split_lo) return;          // the general path's segment (it also writes n == 0)
    if (tid == 0 && member == 0 && SLAB && P.cand_count_out) P.cand_count_out[seg] = n_true;
    if (n == 0) {
        if (tid == 0 && member == 0) {
            if (P.keep_count) P.keep_count[seg] = 0;
            if (P.det_count) P.det_count[seg] = 0;
        }
        return;
    }
    unsigned long long* prof = P.prof ? reinterpret_cast<unsigned long long*>(P.prof) + (size_t)blockIdx.x * 8 : nullptr;
    if (prof && tid == 0) { prof[0] = gtime(); prof[1] = ((unsigned long long)seg << 40) | ((unsigned long long)team << 32) | (unsigned)n; }
    const bool nofilter = fused_rank<SLAB, NT>(P, seg, off, n, n_true, R, smem, red, member == 0) || !(P.thr_f > 0.f);
    if (prof && tid == 0) prof[2] = gtime();

    // the nt*(nt+1)/2 tiles in row-major order (row tile rt holds column tiles 0..rt): this member's contiguous share
    const int nt = cdiv(n, 64);
    const int T = nt * (nt + 1) / 2;
    const int t_begin = (int)((long long)T * member / team), t_end = (int)((long long)T * (member + 1) / team);
    for (int t = t_begin; t < t_end;) {
        int rt = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
        while (rt * (rt + 1) / 2 > t) --rt;
        while ((rt + 1) * (rt + 2) / 2 <= t) ++rt;
        const int c0 = t - rt * (rt + 1) / 2;
        const int c1 = min(rt + 1, c0 + (t_end - t));
        switch (P.mode) {
            case B200_NMS_MAJORITY: fused_strip<B200_NMS_MAJORITY, NT>(P, seg, n, rt, c0, c1, R, smem, nofilter); break;
            case B200_NMS_TV_AUTO:   // shifted boxes of different labels never intersect: the label test is exact for both
            case B200_NMS_TV_CLASS: fused_strip<B200_NMS_TV_CLASS, NT>(P, seg, n, rt, c0, c1, R, smem, nofilter); break;
            default:                fused_strip<B200_NMS_TV, NT>(P, seg, n, rt, c0, c1, R, smem, nofilter); break;      // TV, TV_TRICK
        }
        t += c1 - c0;
    }
    if (prof && tid == 0) prof[3] = gtime();
    bool last = true;
    if (team > 1) {
        __threadfence();                                         // release: this member's mask words
        __syncthreads();
        if (tid == 0) cmd[0] = atomicAdd(P.f_ctl + seg, 1) == team - 1;
        __syncthreads();
        last = cmd[0] != 0;
        if (last) __threadfence();                               // acquire: every member's mask words
    }
    __syncthreads();
    if (prof && tid == 0) prof[4] = gtime();
    if (last) {
        if (P.mode == B200_NMS_MAJORITY) fused_resolve<B200_NMS_MAJORITY, SLAB, NT>(P, seg, off, n, R, smem, scan, prof);
        else                             fused_resolve<B200_NMS_TV, SLAB, NT>(P, seg, off, n, R, smem, scan, prof);
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------- kernel
template <bool SLAB, int NT>
__global__ void __launch_bounds__(NT, 1024 / NT)
k_nms_fused(const __grid_constant__ NmsParams P) {
    constexpr int kFT = NT, kFW = NT / 32;
    __shared__ __align__(16) unsigned char smem[kSmemBytes];
    __shared__ float red[32];
    __shared__ int scan[32];
    __shared__ int cmd[4];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int S = P.num_segments, G = (int)gridDim.x;
    const size_t slot = (size_t)blockIdx.x * (size_t)P.max_seg;
    const RankArrays R{P.f_rbox + slot, P.f_rarea + slot, P.f_rlabel + slot, P.f_rkey + slot};

    if (S >= G) {
        for (int seg = blockIdx.x; seg < S; seg += G) fused_segment<SLAB, NT>(P, seg, 0, 1, R, smem, red, scan, cmd);
        return;
    }
    // ---- fewer segments than CTAs: teams.  Every CTA derives the same assignment from the segment sizes (same
    //      instructions on the same inputs): one CTA per segment, the G - S others spread in proportion to
    //      n_s * (n_s + 256) ~ pair count + per-box work, remainders handed out largest first.
    float* share = reinterpret_cast<float*>(smem);                 // [S]   (S < G <= kFT)
    int* team_of = reinterpret_cast<int*>(smem + 4096);            // [S]
    int* first = reinterpret_cast<int*>(smem + 8192);              // [S+1]
    float cost = 0.f;
    if (tid < S) {
        long long off;
        int n, n_true;
        segment_range(P, tid, off, n, n_true);
        cost = n > P.split_lo ? (float)n * (float)(n + 256) : 0.f;       // segments of the other path cost nothing
    }
    float tot = cost;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(kFullMask, tot, o);
    if (lane == 0) red[warp] = tot;
    __syncthreads();
    tot = 0.f;
    for (int w = 0; w < kFW; ++w) tot += red[w];                  // same order in every thread of every CTA
    const int extra = G - S;
    float sh = tot > 0.f ? (float)extra * 0.999999f * cost / tot : 0.f;
    const float fl = floorf(sh);
    if (tid < S) share[tid] = sh - fl;
    int mine = tid < S ? 1 + (int)fl : 0;
    int used = mine;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) used += __shfl_xor_sync(kFullMask, used, o);
    __syncthreads();                                               // red[] consumed, share[] visible
    if (lane == 0) scan[warp] = used;
    __syncthreads();
    used = 0;
    for (int w = 0; w < kFW; ++w) used += scan[w];
    const int left = G - used;                                     // CTAs not yet handed out (0 <= left < S + 1)
    if (tid < S && left > 0 && cost > 0.f) {
        const float rem = share[tid];
        int rank = 0;
        for (int s = 0; s < S; ++s) {
            const float o = share[s];
            rank += (o > rem) || (o == rem && s < tid);
        }
        if (rank < left) ++mine;
    }
    if (tid < S) team_of[tid] = mine;
    __syncthreads();
    if (tid < S) first[tid] = team_of[tid];
    if (tid == S) first[tid] = 0;
    __syncthreads();
    f_exclusive_scan<NT>(first, S + 1, scan);                          // first[s] = first CTA of segment s, first[S] <= G
    __syncthreads();
    // binary search: the segment whose CTA range holds blockIdx.x
    const int me = blockIdx.x;
    int seg = -1, member = 0, team = 1;
    if (me < first[S]) {
        int lo = 0, hi = S - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (first[mid] <= me) lo = mid; else hi = mid - 1;
        }
        seg = lo;
        member = me - first[lo];
        team = first[lo + 1] - first[lo];
    }
    __syncthreads();
    if (seg >= 0) fused_segment<SLAB, NT>(P, seg, member, team, R, smem, red, scan, cmd);
}

// ---------------------------------------------------------------------------------------------- host
static constexpr int kFusedGridMax = 320;       // private rank-array slots carved per launch (>= 2 x SMs of a B200)
int nms_fused_slots() { return kFusedGridMax; }

bool nms_fused_eligible(const NmsParams& P) {
    if (P.mode < 0 || P.max_seg > 4096) return false;
    if (P.from_slab && (P.anchor_space <= 0 || P.anchor_space > (1 << 20) || P.cap > 4096)) return false;
    return P.f_ctl != nullptr;
}

int launch_nms_fused(NmsParams& P, int num_segments, cudaStream_t stream) {
    if (cudaMemsetAsync(P.f_ctl, 0, sizeof(int) * (size_t)num_segments, stream) != cudaSuccess) return B200_ERR_CUDA;
    // a lone small segment does not need the whole chip: at most one CTA per row tile
    const long long useful = (long long)num_segments * (P.max_words > 0 ? P.max_words : 1);
    const int sms = current_sm_count();
    if (P.split_lo != 0 && !P.serial) {
        // beside the general path (large segments of a candidate slab only): small CTAs, two per SM, that start next
        // to the streaming decode kernel and leave at once when the batch has no large segment
        int grid = 2 * sms;
        if (grid > kFusedGridMax) grid = kFusedGridMax;
        if (useful < grid) grid = (int)useful;
        if (P.from_slab) k_nms_fused<true, 256><<<grid, 256, 0, stream>>>(P);
        else             k_nms_fused<false, 256><<<grid, 256, 0, stream>>>(P);
    } else {
        int grid = sms;
        if (grid > kFusedGridMax) grid = kFusedGridMax;
        if (useful < grid) grid = (int)useful;
        if (P.from_slab) k_nms_fused<true, 1024><<<grid, 1024, 0, stream>>>(P);
        else             k_nms_fused<false, 1024><<<grid, 1024, 0, stream>>>(P);
    }
    return cudaGetLastError() == cudaSuccess ? B200_OK : B200_ERR_CUDA;
}

}  // namespace b200
