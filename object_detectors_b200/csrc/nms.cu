// nms.cu -- per-segment greedy NMS, one CTA per segment (image, or image x level), sm_100a.
//
// One launch handles every segment of a batch.  A CTA
//   1. (YOLO path) canonicalises its image's unordered candidate slab: bitonic sort by flat anchor
//      index -> the reference's ascending-anchor candidate list (test_one_epoch.py:27-28),
//   2. sorts by (score desc, index asc) -- helper.py:308 / torchvision's stable sort,
//   3. runs greedy suppression block-serially: rows are taken 64 at a time; the 64x64 diagonal
//      block is turned into ballot bitmasks and resolved by one warp with register-resident rows,
//      then all threads test the kept rows of the block against every later, still-alive column.
//      IoUs are computed on the fly (no O(n^2) mask in memory) and only for kept rows, which is
//      also what makes "first suppressor" (needed for the majority vote) fall out for free,
//   4. (MAJORITY) relabels each kept box by the vote of the boxes it removed (helper.py:368-375),
//   5. compacts the kept rows in score order into the outputs.
// Segments up to kSmemCap boxes live entirely in shared memory; larger ones run the same code on
// global-memory scratch (correct, slow; a multi-CTA path for huge segments is future work).
#include "decode.cuh"
#include "nms.cuh"

namespace b200 {

static constexpr int kThreads = 512;
static constexpr int kWarps = kThreads / 32;
static constexpr int kVoteFlag = 1 << 30;
static constexpr int kVoteListCap = 128;

struct SegStore {
    unsigned long long* key;  // [P] sort keys
    float4* box;              // [n] sorted boxes (shifted in TRICK mode)
    float* area;              // [n] sorted areas; reused as int new-label after suppression
    int* label;               // [n] sorted labels (snapshot)
    int* sup;                 // [n] -1 = alive/kept, else first suppressor (sorted idx) | vote flag
    int* cidx;                // [n] sorted position -> canonical index
};

__device__ __forceinline__ int next_pow2(int n) {
    int p = 1;
    while (p < n) p <<= 1;
    return p;
}

// ascending bitonic sort of key[0..P), P a power of two, whole CTA
__device__ void bitonic_sort(unsigned long long* key, int P) {
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < (P >> 1); t += blockDim.x) {
                // t-th pair of this stage: insert a zero bit at position log2(j)
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int ixj = i | j;
                const bool asc = (i & k) == 0;
                const unsigned long long a = key[i], b = key[ixj];
                if ((a > b) == asc) { key[i] = b; key[ixj] = a; }
            }
            __syncthreads();
        }
    }
}

// exclusive prefix over a per-thread flag, processed in rounds of blockDim.x items.
// returns the running total after the round; `scratch` holds kWarps+1 ints.
__device__ __forceinline__ int block_rank(bool flag, int* scratch, int& running) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned bal = __ballot_sync(kFullMask, flag);
    if (lane == 0) scratch[warp] = __popc(bal);
    __syncthreads();
    int before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
        const int c = scratch[w];
        if (w < warp) before += c;
        total += c;
    }
    const int rank = running + before + __popc(bal & ((1u << lane) - 1u));
    running += total;
    __syncthreads();
    return rank;
}

template <int MODE>
__device__ __forceinline__ bool suppresses(const NmsParams& P, const float4& bi, float ai, int li,
                                           const float4& bj, float aj, int lj, bool& vote) {
    vote = false;
    if (MODE == B200_NMS_TV_CLASS && li != lj) return false;
    float w = __fsub_rn(fminf(bi.z, bj.z), fmaxf(bi.x, bj.x));
    float h = __fsub_rn(fminf(bi.w, bj.w), fmaxf(bi.y, bj.y));
    w = fmaxf(w, 0.f);
    h = fmaxf(h, 0.f);
    const float inter = __fmul_rn(w, h);
    if (MODE == B200_NMS_MAJORITY) {
        // helper.py:361-369: union = (area_T - inter) + area_S, T = remaining (j), S = picked (i)
        const float den = __fadd_rn(__fsub_rn(aj, inter), ai);
        if (inter == 0.f && P.fast_reject && den != 0.f) return false;   // IoU == +-0 < thr
        const float iou = __fdiv_rn(inter, den);
        vote = iou > P.thr_f;
        return !(iou < P.thr_f);
    } else {
        const float den = __fsub_rn(__fadd_rn(ai, aj), inter);
        if (inter == 0.f && P.fast_reject && den != 0.f) return false;
        const float iou = __fdiv_rn(inter, den);
        return (double)iou > P.thr_d;
    }
}

template <int MODE>
__device__ void nms_core(const NmsParams& P, const SegStore& S, int n, int* sm_small,
                         unsigned long long* dsup, unsigned long long* dvote) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int* klist = sm_small;        // [64] kept rows of the current block (sorted idx)
    int* kcount = sm_small + 64;  // [1]

    for (int i = tid; i < n; i += kThreads) S.sup[i] = -1;
    __syncthreads();

    for (int base = 0; base < n; base += 64) {
        const int m = min(64, n - base);
        // ---- (a) diagonal 64x64 block -> ballot masks; row r only needs columns > r -----------
        for (int r = warp; r < 64; r += kWarps) {
            unsigned lo = 0, hi = 0, vlo = 0, vhi = 0;
            if (r < m) {
                const float4 bi = S.box[base + r];
                const float ai = S.area[base + r];
                const int li = S.label[base + r];
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int c = half * 32 + lane;
                    bool sup = false, vote = false;
                    if (c > r && c < m)
                        sup = suppresses<MODE>(P, bi, ai, li, S.box[base + c], S.area[base + c],
                                               S.label[base + c], vote);
                    const unsigned bs = __ballot_sync(kFullMask, sup);
                    const unsigned bv = __ballot_sync(kFullMask, vote);
                    if (half == 0) { lo = bs; vlo = bv; } else { hi = bs; vhi = bv; }
                }
            }
            if (lane == 0) {
                dsup[r] = ((unsigned long long)hi << 32) | lo;
                dvote[r] = ((unsigned long long)vhi << 32) | vlo;
            }
        }
        __syncthreads();
        // ---- (b) one warp resolves the block; lane l keeps rows l, l+32 in registers ------------
        if (warp == 0) {
            const bool a0 = lane < m && S.sup[base + lane] < 0;
            const bool a1 = lane + 32 < m && S.sup[base + lane + 32] < 0;
            unsigned long long alive = ((unsigned long long)__ballot_sync(kFullMask, a1) << 32) |
                                       __ballot_sync(kFullMask, a0);
            const unsigned long long row0 = dsup[lane], row1 = dsup[lane + 32];
            const unsigned long long vot0 = dvote[lane], vot1 = dvote[lane + 32];
            int sup0 = -1, sup1 = -1;  // suppressor of columns lane / lane+32 found in this block
#pragma unroll 8
            for (int k = 0; k < 64; ++k) {
                const unsigned long long src_r = (k < 32) ? row0 : row1;
                const unsigned long long src_v = (k < 32) ? vot0 : vot1;
                const unsigned long long rk = __shfl_sync(kFullMask, src_r, k & 31);
                const unsigned long long vk = __shfl_sync(kFullMask, src_v, k & 31);
                if ((alive >> k) & 1ull) {
                    const unsigned long long hit = rk & alive;
                    if (hit) {
                        if ((hit >> lane) & 1ull) sup0 = (base + k) | (((vk >> lane) & 1ull) ? kVoteFlag : 0);
                        if ((hit >> (lane + 32)) & 1ull) sup1 = (base + k) | (((vk >> (lane + 32)) & 1ull) ? kVoteFlag : 0);
                        alive &= ~hit;
                    }
                }
            }
            if (sup0 >= 0) S.sup[base + lane] = sup0;
            if (sup1 >= 0) S.sup[base + lane + 32] = sup1;
            // rows still alive are kept; list them in order for phase (c)
            const unsigned klo = (unsigned)alive, khi = (unsigned)(alive >> 32);
            const unsigned lt = (1u << lane) - 1u;
            if ((klo >> lane) & 1u) klist[__popc(klo & lt)] = base + lane;
            if ((khi >> lane) & 1u) klist[__popc(klo) + __popc(khi & lt)] = base + lane + 32;
            if (lane == 0) *kcount = __popc(klo) + __popc(khi);
        }
        __syncthreads();
        // ---- (c) kept rows of this block against all later, still-alive columns ------------------
        const int kc = *kcount;
        for (int j = base + 64 + tid; j < n; j += kThreads) {
            if (S.sup[j] >= 0) continue;
            const float4 bj = S.box[j];
            const float aj = S.area[j];
            const int lj = S.label[j];
            for (int t = 0; t < kc; ++t) {
                const int i = klist[t];
                bool vote;
                if (suppresses<MODE>(P, S.box[i], S.area[i], S.label[i], bj, aj, lj, vote)) {
                    S.sup[j] = i | (vote ? kVoteFlag : 0);
                    break;
                }
            }
        }
        __syncthreads();
    }
}

// helper.py:368-375 -- kept box i is relabelled to the most frequent class among the boxes it
// removed with IoU > thr, if those hold more than one distinct class (ties -> smallest id).
__device__ void majority_relabel(const SegStore& S, int n, int* newlab, int* vote_list) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int* list = vote_list + warp * kVoteListCap;
    for (int i = warp; i < n; i += kWarps) {
        if (S.sup[i] >= 0) continue;  // not kept (warp-uniform)
        const int want = i | kVoteFlag;
        // pass 1: collect voter labels
        int L = 0;
        for (int j0 = i + 1; j0 < n; j0 += 32) {
            const int j = j0 + lane;
            const bool v = j < n && S.sup[j] == want;
            const unsigned bal = __ballot_sync(kFullMask, v);
            if (v) {
                const int pos = L + __popc(bal & ((1u << lane) - 1u));
                if (pos < kVoteListCap) list[pos] = S.label[j];
            }
            L += __popc(bal);
        }
        __syncwarp();
        int label = S.label[i];
        if (L >= 2) {
            int best_cnt = 0, best_lab = 0x7fffffff;
            if (L <= kVoteListCap) {
                for (int a = lane; a < L; a += 32) {
                    const int la = list[a];
                    int cnt = 0;
                    for (int b = 0; b < L; ++b) cnt += (list[b] == la);
                    if (cnt > best_cnt || (cnt == best_cnt && la < best_lab)) { best_cnt = cnt; best_lab = la; }
                }
            } else {
                // rare: more voters than the list holds -> recount by rescanning the segment
                for (int j = i + 1 + lane; j < n; j += 32) {
                    if (S.sup[j] != want) continue;
                    const int la = S.label[j];
                    int cnt = 0;
                    for (int b = i + 1; b < n; ++b) cnt += (S.sup[b] == want && S.label[b] == la);
                    if (cnt > best_cnt || (cnt == best_cnt && la < best_lab)) { best_cnt = cnt; best_lab = la; }
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const int oc = __shfl_xor_sync(kFullMask, best_cnt, o);
                const int ol = __shfl_xor_sync(kFullMask, best_lab, o);
                if (oc > best_cnt || (oc == best_cnt && ol < best_lab)) { best_cnt = oc; best_lab = ol; }
            }
            if (best_cnt < L) label = best_lab;  // more than one distinct class among the voters
        }
        if (lane == 0) newlab[i] = label;
        __syncwarp();
    }
}

template <int MODE, bool FROM_SLAB>
__device__ void segment_body(const NmsParams& P, int seg, unsigned char* smem_raw) {
    const int tid = threadIdx.x;
    __shared__ int sm_small[80];
    __shared__ int sm_scan[kWarps + 1];
    __shared__ unsigned long long dsup[64], dvote[64];
    __shared__ float sm_red[kWarps];

    long long off;
    int n, n_true;
    if (FROM_SLAB) {
        n_true = P.count[seg];
        n = min(n_true, P.cap);
        off = (long long)seg * P.cap;
        if (tid == 0 && P.cand_count_out) P.cand_count_out[seg] = n_true;
    } else {
        off = P.seg_offsets[seg];
        n = P.seg_counts ? P.seg_counts[seg] : P.seg_offsets[seg + 1] - (int)off;
        n_true = n;
    }
    if (n <= 0) {
        if (tid == 0) {
            if (P.keep_count) P.keep_count[seg] = 0;
            if (P.det_count) P.det_count[seg] = 0;
        }
        return;
    }
    const int Ppad = next_pow2(n);

    SegStore S;
    if (n <= P.smem_cap) {
        unsigned char* q = smem_raw;
        S.key = reinterpret_cast<unsigned long long*>(q); q += sizeof(unsigned long long) * (size_t)P.smem_cap;
        S.box = reinterpret_cast<float4*>(q);              q += sizeof(float4) * (size_t)P.smem_cap;
        S.area = reinterpret_cast<float*>(q);              q += sizeof(float) * (size_t)P.smem_cap;
        S.label = reinterpret_cast<int*>(q);               q += sizeof(int) * (size_t)P.smem_cap;
        S.sup = reinterpret_cast<int*>(q);                 q += sizeof(int) * (size_t)P.smem_cap;
        S.cidx = reinterpret_cast<int*>(q);
    } else {
        S.key = P.gkey + 2 * off;
        S.box = P.gbox + off;
        S.area = P.garea + off;
        S.label = P.glabel + off;
        S.sup = P.gsup + off;
        S.cidx = P.gcidx + off;
    }
    int* vote_list = reinterpret_cast<int*>(smem_raw + (size_t)P.smem_cap * 40);

    // canonical (ascending anchor / input order) views of this segment
    const float4* cbox;
    const float* cscore;
    const int* clabel;
    if (FROM_SLAB) {
        // ---- 1. sort the unordered slab by flat anchor index ---------------------------------
        const Cand* slab = P.slab + off;
        for (int i = tid; i < Ppad; i += kThreads)
            S.key[i] = i < n ? (((unsigned long long)(unsigned)slab[i].anchor << 32) | (unsigned)i) : ~0ull;
        __syncthreads();
        bitonic_sort(S.key, Ppad);
        for (int p = tid; p < n; p += kThreads) {
            const Cand c = slab[(unsigned)S.key[p]];
            P.cbox[off + p] = make_float4(c.x1, c.y1, c.x2, c.y2);
            P.cscore[off + p] = c.score;
            P.clabel[off + p] = c.label;
            P.canchor[off + p] = c.anchor;
        }
        __syncthreads();  // global writes above are re-read by this CTA below
        cbox = P.cbox + off;
        cscore = P.cscore + off;
        clabel = P.clabel + off;
    } else {
        cbox = reinterpret_cast<const float4*>(P.boxes) + off;
        cscore = P.scores + off;
        clabel = P.labels ? P.labels + off : nullptr;
    }
    if (P.mode < 0) return;  // canonicalise only (b200_yolo_decode_filter)

    // ---- 2. sort by (score desc, canonical index asc) ------------------------------------------
    for (int i = tid; i < Ppad; i += kThreads)
        S.key[i] = i < n ? (((unsigned long long)(~orderable(cscore[i])) << 32) | (unsigned)i) : ~0ull;
    __syncthreads();
    bitonic_sort(S.key, Ppad);

    float shift_unit = 0.f;
    if (MODE == B200_NMS_TV_TRICK) {
        // offsets = idxs.to(boxes) * (boxes.max() + 1)   (torchvision boxes.py coordinate trick)
        float mx = -INFINITY;
        for (int i = tid; i < n; i += kThreads) {
            const float4 b = cbox[i];
            mx = fmaxf(mx, fmaxf(fmaxf(b.x, b.y), fmaxf(b.z, b.w)));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(kFullMask, mx, o));
        if ((tid & 31) == 0) sm_red[tid >> 5] = mx;
        __syncthreads();
        mx = sm_red[0];
#pragma unroll
        for (int w = 1; w < kWarps; ++w) mx = fmaxf(mx, sm_red[w]);
        shift_unit = __fadd_rn(mx, 1.0f);
    }
    for (int q = tid; q < n; q += kThreads) {
        const int p = (int)(unsigned)S.key[q];
        float4 b = cbox[p];
        const int lab = clabel ? clabel[p] : 0;
        if (MODE == B200_NMS_TV_TRICK) {
            const float sh = __fmul_rn((float)lab, shift_unit);
            b = make_float4(__fadd_rn(b.x, sh), __fadd_rn(b.y, sh), __fadd_rn(b.z, sh), __fadd_rn(b.w, sh));
        }
        S.box[q] = b;
        S.area[q] = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
        S.label[q] = lab;
        S.cidx[q] = p;
    }
    __syncthreads();

    // ---- 3. greedy suppression -----------------------------------------------------------------
    nms_core<MODE>(P, S, n, sm_small, dsup, dvote);

    // ---- 4. majority relabel ---------------------------------------------------------------------
    int* newlab = reinterpret_cast<int*>(S.area);
    if (MODE == B200_NMS_MAJORITY) {
        majority_relabel(S, n, newlab, vote_list);
    } else {
        for (int q = tid; q < n; q += kThreads) newlab[q] = S.label[q];
    }
    __syncthreads();

    // ---- 5. compact kept rows in score order -----------------------------------------------------
    int running = 0;
    for (int q0 = 0; q0 < n; q0 += kThreads) {
        const int q = q0 + tid;
        const bool kept = q < n && S.sup[q] < 0;
        const int k = block_rank(kept, sm_scan, running);
        if (!kept) continue;
        const int p = S.cidx[q];
        if (FROM_SLAB) {
            if (k < P.max_det) {
                const float4 b = cbox[p];
                float* d = P.det + ((size_t)seg * P.max_det + k) * 6;
                d[0] = b.x; d[1] = b.y; d[2] = b.z; d[3] = b.w;
                d[4] = cscore[p];
                d[5] = (float)newlab[q];
                P.det_keep[(size_t)seg * P.max_det + k] = p;
                if (P.det_anchor) P.det_anchor[(size_t)seg * P.max_det + k] = P.canchor[off + p];
            }
        } else {
            P.keep[off + k] = p;
            if (P.labels_out) P.labels_out[off + k] = newlab[q];
        }
    }
    if (tid == 0) {
        if (FROM_SLAB) {
            P.det_count[seg] = min(running, P.max_det);
            if (running > P.max_det && P.status) atomicOr(P.status, 2);
        } else {
            P.keep_count[seg] = running;
        }
    }
}

template <bool FROM_SLAB>
__global__ void __launch_bounds__(kThreads, 1)
k_nms_segments(const __grid_constant__ NmsParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int seg = blockIdx.x;
    switch (P.mode) {
        case B200_NMS_MAJORITY: segment_body<B200_NMS_MAJORITY, FROM_SLAB>(P, seg, smem_raw); break;
        case B200_NMS_TV:       segment_body<B200_NMS_TV, FROM_SLAB>(P, seg, smem_raw); break;
        case B200_NMS_TV_CLASS: segment_body<B200_NMS_TV_CLASS, FROM_SLAB>(P, seg, smem_raw); break;
        case B200_NMS_TV_TRICK: segment_body<B200_NMS_TV_TRICK, FROM_SLAB>(P, seg, smem_raw); break;
        default:                segment_body<B200_NMS_TV, FROM_SLAB>(P, seg, smem_raw); break;  // mode<0
    }
}

size_t nms_smem_bytes(int smem_cap) {
    return (size_t)smem_cap * 40 + (size_t)kWarps * kVoteListCap * sizeof(int);
}

int launch_nms(const NmsParams& P, int num_segments, bool from_slab, cudaStream_t stream) {
    if (num_segments <= 0) return B200_OK;
    const size_t smem = nms_smem_bytes(P.smem_cap);
    static bool attr_set[2] = {false, false};
    if (!attr_set[from_slab ? 1 : 0]) {
        cudaError_t e = from_slab
            ? cudaFuncSetAttribute(k_nms_segments<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
            : cudaFuncSetAttribute(k_nms_segments<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return B200_ERR_CUDA;
        attr_set[from_slab ? 1 : 0] = true;
    }
    if (from_slab) k_nms_segments<true><<<num_segments, kThreads, smem, stream>>>(P);
    else           k_nms_segments<false><<<num_segments, kThreads, smem, stream>>>(P);
    return cudaGetLastError() == cudaSuccess ? B200_OK : B200_ERR_CUDA;
}

}  // namespace b200
