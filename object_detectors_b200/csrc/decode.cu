// decode.cu -- YOLO head decode kernels (sm_100a).
//
//  k_decode_filter : fused decode + score threshold + warp-ballot stream compaction.  Streams the
//                    raw head tensors exactly once (HBM-bound), never materialises [B,N,5+C].
//                    Replaces yolo_forw.py:93-119,163-176 + helper.py:203-217 +
//                    test_one_epoch.py:25-28,35.
//  k_decode_dense  : the [B,N,5+C] tensor YOLOForw.forward returns (drop-in for callers that
//                    want the dense output), smem-transposed so reads and writes are coalesced.
//
// Layout reminder: head s is [B, A*(5+C), H, W]; for a fixed (b, a) the 5+C channel planes are
// rows of HW contiguous floats.  A warp task owns 32*VEC consecutive cells of one (b, a) and
// walks down the planes, so every load is a fully coalesced 512 B (VEC=4) / 128 B (VEC=1) row
// segment and each thread keeps the running class maximum / softmax denominator of its own cells
// in registers (online softmax, one ex2 per logit).
#include "decode.cuh"

namespace b200 {

// ------------------------------------------------------------------------------------------
// host: geometry
// ------------------------------------------------------------------------------------------
int make_decode_params(const b200_yolo_layout* L, const float* const* heads, const float* idf,
                       DecodeParams* p) {
    if (!L || !heads || !p) return B200_ERR_INVALID;
    if (L->num_scales < 1 || L->num_scales > B200_MAX_SCALES) return B200_ERR_INVALID;
    if (L->num_anchors < 1 || L->num_anchors > B200_MAX_ANCHORS) return B200_ERR_INVALID;
    if (L->num_classes < 1 || L->batch < 1) return B200_ERR_INVALID;
    p->num_scales = L->num_scales;
    p->A = L->num_anchors;
    p->C = L->num_classes;
    p->B = L->batch;
    p->idf = idf;
    int task = 0, anchor_off = 0;
    for (int s = 0; s < L->num_scales; ++s) {
        ScaleDev& d = p->sc[s];
        if (!heads[s] || L->grid[s] < 1) return B200_ERR_INVALID;
        d.head = heads[s];
        d.grid = L->grid[s];
        d.hw = d.grid * d.grid;
        // 128-bit loads need every plane row (hw floats apart) 16 B aligned
        const bool aligned = (d.hw % 4 == 0) && ((reinterpret_cast<uintptr_t>(heads[s]) & 15u) == 0);
        d.vec = aligned ? 4 : 1;
        d.tiles = cdiv(d.hw, 128);   // 32 lanes x 4 cells per warp task
        d.task_begin = task;
        d.anchor_off = anchor_off;
        d.inw = (float)d.grid;
        d.stride = L->img_size / d.inw;  // fp32 division, as `self.img_size / inw_inh`
        for (int a = 0; a < L->num_anchors; ++a) {
            d.anc[a][0] = L->anchor_rel[s][a][0];
            d.anc[a][1] = L->anchor_rel[s][a][1];
        }
        task += L->batch * L->num_anchors * d.tiles;
        anchor_off += d.hw * L->num_anchors;
    }
    p->total_tasks = task;
    p->N = anchor_off;
    return B200_OK;
}

// ------------------------------------------------------------------------------------------
// fused decode + filter
// ------------------------------------------------------------------------------------------
// `sink` folds every loaded word into a register that is (never, but unprovably) stored at the end
// of the task.  Without it ptxas sinks the class-plane loads under the live-cell predicate -- the
// inline-asm `volatile` does not bind ptxas -- and the kernel silently degrades from a coalesced
// stream into sector-granular gathers (measured: 184 MB of DRAM reads instead of 495 MB, same time).
// Every lane owns kCellsPerLane = 4 cells: 4 consecutive ones fetched by one 128-bit load when the
// plane rows are 16 B aligned (VEC = 4), else the cells lane, lane+32, lane+64, lane+96 of the tile
// fetched by four coalesced 32-bit loads (VEC = 1, odd grids).
static constexpr int kCellsPerLane = 4;
template <int VEC, bool STREAM>
__device__ __forceinline__ void load_row(const float* p, const bool (&in)[kCellsPerLane], float (&v)[kCellsPerLane],
                                         unsigned& sink) {
    if constexpr (VEC == 4) {
        float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
        if (in[0]) r = ldg_stream_v4(p);
        v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w;
    } else {
#pragma unroll
        for (int k = 0; k < kCellsPerLane; ++k) v[k] = in[k] ? ldg_stream_f32(p + 32 * k) : 0.f;
    }
    if constexpr (STREAM) sink ^= __float_as_uint(v[0]) ^ __float_as_uint(v[1]) ^ __float_as_uint(v[2]) ^ __float_as_uint(v[3]);
}

static constexpr float kLog2e = 1.4426950408889634f;
#ifndef DF_U
#define DF_U 8        // class planes loaded per chunk (tuned on B200: 8 planes x 6 CTAs/SM)
#endif
#ifndef DF_MINB
#define DF_MINB 6     // resident CTAs per SM the register budget is set for
#endif

// One warp task: 32*VEC cells of one (scale, b, a).
template <int VEC, bool SOFTMAX, bool HAS_IDF, int U, bool STREAM>
__device__ __forceinline__ void decode_filter_task(const DecodeParams& p, const ScaleDev& sc, int b,
                                                   int a, int tile, int lane) {
    constexpr int NC = kCellsPerLane;
    // cell k of this lane: VEC=4 -> hw0 + k ; VEC=1 -> hw0 + 32*k
    const int hw0 = VEC == 4 ? (tile * 32 + lane) * 4 : tile * 128 + lane;
    constexpr int kStep = VEC == 4 ? 1 : 32;
    bool in[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) in[k] = hw0 + k * kStep < sc.hw;   // VEC=4: hw % 4 == 0, all or nothing
    const int C = p.C;
    const size_t plane = (size_t)sc.hw;
    const float* base = sc.head + ((size_t)(b * p.A + a) * (size_t)(5 + C)) * plane + (size_t)hw0;

    // --- objectness plane: which cells can still pass?  score = conf*maxp <= conf ------------
    unsigned sink = 0u;
    float t4[NC];
    load_row<VEC, STREAM>(base + 4 * plane, in, t4, sink);
    float conf[NC];
    bool live[NC];
    bool any_live = false;
#pragma unroll
    for (int k = 0; k < NC; ++k) {
        conf[k] = sigmoid_ref(t4[k]);
        live[k] = in[k] && (conf[k] > p.thr);
        any_live |= live[k];
    }

    // GATED variant (STREAM == false): class and box planes are fetched only where a cell can still
    // pass -- per lane for the 128-bit path (one load spans the lane's 4 cells), per cell otherwise.
    bool want[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) want[k] = STREAM ? in[k] : (VEC == 4 ? (in[k] && any_live) : live[k]);

    // --- class sweep: running max (first index wins) and softmax denominator -----------------
    float m[NC], s[NC];
    int arg[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) { m[k] = -INFINITY; s[k] = 0.f; arg[k] = 0; }

    const float* cls = base + 5 * plane;
    int c = 0;
    for (; c + U <= C; c += U) {
        float v[U][NC];
#pragma unroll
        for (int u = 0; u < U; ++u) load_row<VEC, STREAM>(cls + (size_t)(c + u) * plane, want, v[u], sink);
        if (any_live) {
            float w[U];
            if constexpr (HAS_IDF) {
#pragma unroll
                for (int u = 0; u < U; ++u) w[u] = __ldg(p.idf + c + u);
            }
#pragma unroll
            for (int k = 0; k < NC; ++k) {
                if (live[k]) {
                    float x[U];
                    const float m_old = m[k];
                    float mk = m_old;
                    int ak = arg[k];
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        x[u] = HAS_IDF ? __fmul_rn(w[u], v[u][k]) : v[u][k];
                        if (x[u] > mk) {
                            // the reference takes the first maximum of the PROBABILITIES (test_one_epoch.py:35): where
                            // the fp32 sigmoid is flat (saturated at 1.0f above ~16.6) a larger logit is not a new maximum
                            if (SOFTMAX || mk == -INFINITY || sigmoid_ref(x[u]) > sigmoid_ref(mk)) ak = c + u;
                            mk = x[u];
                        }
                    }
                    m[k] = mk;
                    arg[k] = ak;
                    if constexpr (SOFTMAX) {
                        // s <- s * exp(m_old - m_new) + sum_u exp(x_u - m_new)
                        float acc = __fmul_rn(s[k], ex2_approx(__fmul_rn(__fsub_rn(m_old, mk), kLog2e)));
#pragma unroll
                        for (int u = 0; u < U; ++u)
                            acc = __fadd_rn(acc, ex2_approx(__fmul_rn(__fsub_rn(x[u], mk), kLog2e)));
                        s[k] = acc;
                    }
                }
            }
        }
    }
    for (; c < C; ++c) {  // tail classes (C % U)
        float v[NC];
        load_row<VEC, STREAM>(cls + (size_t)c * plane, want, v, sink);
        if (any_live) {
            const float w = HAS_IDF ? __ldg(p.idf + c) : 1.0f;
#pragma unroll
            for (int k = 0; k < NC; ++k) {
                if (live[k]) {
                    const float x = HAS_IDF ? __fmul_rn(w, v[k]) : v[k];
                    const float m_old = m[k];
                    if (x > m_old) {
                        if (SOFTMAX || m_old == -INFINITY || sigmoid_ref(x) > sigmoid_ref(m_old)) arg[k] = c;
                        m[k] = x;
                    }
                    if constexpr (SOFTMAX) {
                        float acc = __fmul_rn(s[k], ex2_approx(__fmul_rn(__fsub_rn(m_old, m[k]), kLog2e)));
                        s[k] = __fadd_rn(acc, ex2_approx(__fmul_rn(__fsub_rn(x, m[k]), kLog2e)));
                    }
                }
            }
        }
    }

    // --- box planes (last chunk of the stream) ------------------------------------------------
    float tb[4][NC];
#pragma unroll
    for (int u = 0; u < 4; ++u) load_row<VEC, STREAM>(base + (size_t)u * plane, want, tb[u], sink);
    // never true for finite thresholds; keeps every load above architecturally visible
    if (STREAM && sink == 0x9e3779b9u && p.thr < -3.0e38f) atomicOr(p.status, (int)sink);

    // --- threshold + compaction ----------------------------------------------------------------
    float score[NC];
    bool pass[NC];
    unsigned ballots[NC];
    int total = 0;
    const unsigned lt = (1u << lane) - 1u;
#pragma unroll
    for (int k = 0; k < NC; ++k) {
        // max_c softmax = exp(0)/sum ; max_c sigmoid = sigmoid(max logit)
        float best = 0.f;
        if (live[k]) best = SOFTMAX ? __fdiv_rn(1.0f, s[k]) : sigmoid_ref(m[k]);
        score[k] = __fmul_rn(conf[k], best);                // test_one_epoch.py:25
        pass[k] = live[k] && (score[k] > p.thr);            // :26 (strict, fp32)
        ballots[k] = __ballot_sync(kFullMask, pass[k]);
        total += __popc(ballots[k]);
    }
    if (total == 0) return;

    int slot0 = 0;
    if (lane == 0) slot0 = atomicAdd(p.count + b, total);
    slot0 = __shfl_sync(kFullMask, slot0, 0);

    int before = 0;
#pragma unroll
    for (int k = 0; k < NC; ++k) {
        if (pass[k]) {
            const int slot = slot0 + before + __popc(ballots[k] & lt);
            if (slot < p.cap) {
                const int hw = hw0 + k * kStep;
                const int gy_i = hw / sc.grid;
                const int gx_i = hw - gy_i * sc.grid;
                // cxypwh[:, :2] = (idx + 0.5) / in_w   (yolo_forw.py:104-107)
                const float cx = __fdiv_rn((float)gx_i + 0.5f, sc.inw);
                const float cy = __fdiv_rn((float)gy_i + 0.5f, sc.inw);
                // xy = (sigmoid(t) + cxy*inw - 0.5) * stride                       (:166)
                const float bx = __fmul_rn(__fsub_rn(__fadd_rn(sigmoid_ref(tb[0][k]), __fmul_rn(cx, sc.inw)), 0.5f), sc.stride);
                const float by = __fmul_rn(__fsub_rn(__fadd_rn(sigmoid_ref(tb[1][k]), __fmul_rn(cy, sc.inw)), 0.5f), sc.stride);
                // wh = exp(t) * cwh * inw * stride  (left to right)                 (:167)
                const float bw = __fmul_rn(__fmul_rn(__fmul_rn(expf(tb[2][k]), sc.anc[a][0]), sc.inw), sc.stride);
                const float bh = __fmul_rn(__fmul_rn(__fmul_rn(expf(tb[3][k]), sc.anc[a][1]), sc.inw), sc.stride);
                const Box q = abs_coord(bx, by, bw, bh);                              // helper.py:203
                Cand* dst = p.slab + (size_t)b * (size_t)p.cap + (size_t)slot;
                float4* d4 = reinterpret_cast<float4*>(dst);
                d4[0] = make_float4(q.x1, q.y1, q.x2, q.y2);
                d4[1] = make_float4(score[k], __int_as_float(arg[k]),
                                    __int_as_float(sc.anchor_off + hw * p.A + a), 0.f);
            } else {
                atomicOr(p.status, 1);
            }
        }
        before += __popc(ballots[k]);
    }
}

template <bool SOFTMAX, bool HAS_IDF, bool STREAM>
__global__ void __launch_bounds__(128, DF_MINB)
k_decode_filter(const __grid_constant__ DecodeParams p) {
    const int lane = threadIdx.x & 31;
    const int task = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (task >= p.total_tasks) return;
    int s = 0;
#pragma unroll
    for (int i = 1; i < B200_MAX_SCALES; ++i)
        if (i < p.num_scales && task >= p.sc[i].task_begin) s = i;
    const ScaleDev& sc = p.sc[s];
    const int local = task - sc.task_begin;
    const int tile = local % sc.tiles;
    const int ba = local / sc.tiles;
    const int a = ba % p.A;
    const int b = ba / p.A;
    if (sc.vec == 4)
        decode_filter_task<4, SOFTMAX, HAS_IDF, DF_U, STREAM>(p, sc, b, a, tile, lane);
    else
        decode_filter_task<1, SOFTMAX, HAS_IDF, 8, STREAM>(p, sc, b, a, tile, lane);
}

template <bool STREAM>
static void launch_df(const DecodeParams& p, bool softmax, bool idf, int blocks, int threads, cudaStream_t stream) {
    if (softmax) {
        if (idf) k_decode_filter<true, true, STREAM><<<blocks, threads, 0, stream>>>(p);
        else     k_decode_filter<true, false, STREAM><<<blocks, threads, 0, stream>>>(p);
    } else {
        if (idf) k_decode_filter<false, true, STREAM><<<blocks, threads, 0, stream>>>(p);
        else     k_decode_filter<false, false, STREAM><<<blocks, threads, 0, stream>>>(p);
    }
}

// gate = 0: every byte of the head tensors is read (coalesced stream, input independent);
// gate = 1: class / box planes are only fetched for lanes that hold a live cell (sector gathers).
int launch_decode_filter(const DecodeParams& p, bool softmax, int gate, cudaStream_t stream) {
    const int warps_per_block = 4;
    const int blocks = cdiv(p.total_tasks, warps_per_block);
    const bool idf = p.idf != nullptr;
    if (gate) launch_df<false>(p, softmax, idf, blocks, 32 * warps_per_block, stream);
    else      launch_df<true>(p, softmax, idf, blocks, 32 * warps_per_block, stream);
    return cudaGetLastError() == cudaSuccess ? B200_OK : B200_ERR_CUDA;
}



// ------------------------------------------------------------------------------------------
// dense decode: out[b, n, 0:5+C]
// ------------------------------------------------------------------------------------------
// CTA = (scale, b, tile of kCells cells), all A anchors.  Channels are processed in chunks of
// kChunk planes: coalesced plane reads -> smem [a][chunk][cell] -> transposed writes where each
// warp writes a run of contiguous output floats.  Softmax needs the row max / denominator before
// any probability can be written, so class statistics are gathered in a first sweep (registers of
// the thread that owns the (cell, a) pair) and the planes are read a second time (L2 hits: the
// CTA's working set is kCells*A*(5+C)*4 B) for the write sweep.
static constexpr int kCells = 32;
static constexpr int kChunk = 32;

struct DenseParams {
    DecodeParams d;
    float* out;      // [B, N, 5+C]
    int softmax;
    int tile_begin[B200_MAX_SCALES + 1];  // CTA index ranges per scale
};

__global__ void __launch_bounds__(256)
k_decode_dense(const __grid_constant__ DenseParams q) {
    const DecodeParams& p = q.d;
    extern __shared__ float smem[];
    const int A = p.A, C = p.C, CH = 5 + C;
    int s = 0;
    for (int i = 1; i < p.num_scales; ++i)
        if ((int)blockIdx.x >= q.tile_begin[i]) s = i;
    const ScaleDev& sc = p.sc[s];
    const int tiles = cdiv(sc.hw, kCells);
    const int local = blockIdx.x - q.tile_begin[s];
    const int tile = local % tiles;
    const int b = local / tiles;
    const int cell0 = tile * kCells;
    const int ncell = min(kCells, sc.hw - cell0);
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarp = nthr >> 5;

    float* stat_m = smem;                       // [A*kCells] class max
    float* stat_r = stat_m + A * kCells;        // [A*kCells] softmax denominator
    float* tile_s = stat_r + A * kCells;        // [A][kChunk][kCells+1]
    const int ldc = kCells + 1;

    const size_t plane = (size_t)sc.hw;
    const float* img = sc.head + (size_t)b * (size_t)A * (size_t)CH * plane;

    // ---- sweep 1: per (cell, a) class max and softmax denominator -------------------------
    if (q.softmax) {
        for (int item = tid; item < A * kCells; item += nthr) {
            const int a = item / kCells, cl = item % kCells;
            float m = -INFINITY, sum = 0.f;
            if (cl < ncell) {
                const float* col = img + ((size_t)a * CH + 5) * plane + cell0 + cl;
                for (int c = 0; c < C; ++c) {
                    const float w = p.idf ? __ldg(p.idf + c) : 1.0f;
                    const float x = __fmul_rn(w, __ldg(col + (size_t)c * plane));
                    if (x > m) {
                        sum = __fmul_rn(sum, ex2_approx(__fmul_rn(__fsub_rn(m, x), kLog2e)));
                        m = x;
                    }
                    sum = __fadd_rn(sum, ex2_approx(__fmul_rn(__fsub_rn(x, m), kLog2e)));
                }
            }
            stat_m[item] = m;
            stat_r[item] = sum;
        }
    }
    __syncthreads();

    // ---- sweep 2: chunks of channels, transposed through smem -----------------------------
    float* out_img = q.out + ((size_t)b * (size_t)p.N + (size_t)sc.anchor_off + (size_t)cell0 * A) * (size_t)CH;
    for (int ch0 = 0; ch0 < CH; ch0 += kChunk) {
        const int nch = min(kChunk, CH - ch0);
        // load: one warp per (a, channel) row of ncell floats
        for (int row = warp; row < A * nch; row += nwarp) {
            const int a = row / nch, cc = row % nch;
            float v = 0.f;
            if (lane < ncell) v = __ldg(img + ((size_t)a * CH + ch0 + cc) * plane + cell0 + lane);
            tile_s[(a * kChunk + cc) * ldc + lane] = v;
        }
        __syncthreads();
        // store: element e of the [ncell][A][nch] block -> out row (cell*A + a), column ch0+cc
        const int total = ncell * A * nch;
        for (int e = tid; e < total; e += nthr) {
            const int cc = e % nch;
            const int ra = e / nch;       // cell*A + a
            const int a = ra % A, cl = ra / A;
            const float t = tile_s[(a * kChunk + cc) * ldc + cl];
            const int ch = ch0 + cc;
            float r;
            if (ch >= 5) {
                const float w = p.idf ? __ldg(p.idf + (ch - 5)) : 1.0f;
                const float x = __fmul_rn(w, t);
                if (q.softmax) {
                    const float e_ = expf(__fsub_rn(x, stat_m[a * kCells + cl]));
                    r = __fdiv_rn(e_, stat_r[a * kCells + cl]);
                } else {
                    r = sigmoid_ref(x);
                }
            } else if (ch == 4) {
                r = sigmoid_ref(t);
            } else if (ch >= 2) {
                r = __fmul_rn(__fmul_rn(__fmul_rn(expf(t), sc.anc[a][ch - 2]), sc.inw), sc.stride);
            } else {
                const int hw = cell0 + cl;
                const int gy_i = hw / sc.grid, gx_i = hw - gy_i * sc.grid;
                const float g = __fdiv_rn((float)(ch == 0 ? gx_i : gy_i) + 0.5f, sc.inw);
                r = __fmul_rn(__fsub_rn(__fadd_rn(sigmoid_ref(t), __fmul_rn(g, sc.inw)), 0.5f), sc.stride);
            }
            out_img[(size_t)ra * CH + ch] = r;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// dense decode, staged variant (5+C <= kDenseMaxCh): one CTA = 64 cells of one (scale, b, a)
// ------------------------------------------------------------------------------------------
// The (5+C) x 64 tile is read once with coalesced row loads into shared memory (row stride 65 floats: both
// the row-wise and the column-wise accesses below are bank-conflict free), the softmax statistics of a cell
// are computed by 4 threads over disjoint class quarters, and the output rows (5+C contiguous floats each)
// are written by warps with lanes across channels.  Every input byte is read once and every output byte
// written once; 8 CTAs per SM hide the latency.
// probabilities of the dense outputs: ex2.approx (2 ulp) + correctly rounded reciprocal -- well inside the 1e-5
// relative contract and a third of the instructions of expf + IEEE division; box coordinates keep expf
__device__ __forceinline__ float sigmoid_fast(float x) {
    return __frcp_rn(__fadd_rn(1.0f, ex2_approx(__fmul_rn(-x, kLog2e))));
}

static constexpr int kDenseCells = 64;
static constexpr int kDenseLd = kDenseCells + 1;
static constexpr int kDenseMaxCh = 184;          // (5+C) * 65 * 4 B <= 48 KB

struct Dense2Params {
    DecodeParams d;
    float* out;
    int softmax;
    int cta_begin[B200_MAX_SCALES + 1];
    int tiles[B200_MAX_SCALES];
    unsigned long long tiles_inv[B200_MAX_SCALES], a_inv;      // ceil(2^40 / d): n / d == (n * inv) >> 40 for n * d < 2^40
    // legacy (YOLOLoss) mode: one scale, rows ordered (a, h, w), xy = (sigmoid + grid) * stride, sigmoid classes
    int legacy, in_w;
    float stride_w, stride_h;
    const float* legacy_anchors;                 // device [A][2], scaled
};

__global__ void __launch_bounds__(256)
k_decode_dense2(const __grid_constant__ Dense2Params q) {
    extern __shared__ float dsm[];
    const DecodeParams& p = q.d;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int C = p.C, CH = 5 + C;
    float* tile = dsm;                            // [CH][65]
    float* part = tile + CH * kDenseLd;           // [4][64] partial max / partial sums
    float* st_m = part + 4 * kDenseCells;         // [64]
    float* st_r = st_m + kDenseCells;             // [64] 1 / denominator
    int s = 0;
#pragma unroll
    for (int i = 1; i < B200_MAX_SCALES; ++i)
        if (i < p.num_scales && (int)blockIdx.x >= q.cta_begin[i]) s = i;
    const ScaleDev& sc = p.sc[s];
    const int local = blockIdx.x - q.cta_begin[s];
    // (divisions by run-time values cost ~100 instructions per warp -- 8 % of this kernel -- hence the reciprocals)
    const int ba = (int)(((unsigned long long)local * q.tiles_inv[s]) >> 40), tl = local - ba * q.tiles[s];
    const int b = (int)(((unsigned long long)ba * q.a_inv) >> 40), a = ba - b * p.A;
    const int cell0 = tl * kDenseCells;
    const int ncell = min(kDenseCells, sc.hw - cell0);
    const float* src = sc.head + (size_t)ba * (size_t)CH * (size_t)sc.hw + (size_t)cell0;

    // ---- A. stage the tile: one warp per plane row, 2 cells per lane; 6 rows (12 loads) in flight per warp ----
    for (int r0 = warp; r0 < CH; r0 += 48) {
        float v[6][2];
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            const int r = r0 + 8 * j;
            const float* row = src + (size_t)r * (size_t)sc.hw;
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int c = lane + 32 * k;
                v[j][k] = (r < CH && c < ncell) ? ldg_stream_f32(row + c) : 0.f;
            }
        }
        if (p.idf) {
            // class planes are staged already multiplied by their IDF weight (one scalar per row = per warp)
#pragma unroll
            for (int j = 0; j < 6; ++j) {
                const int r = r0 + 8 * j;
                if (r >= 5 && r < CH) {
                    const float wgt = __ldg(p.idf + (r - 5));
                    v[j][0] = __fmul_rn(wgt, v[j][0]);
                    v[j][1] = __fmul_rn(wgt, v[j][1]);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            const int r = r0 + 8 * j;
            if (r < CH) {
                tile[r * kDenseLd + lane] = v[j][0];
                tile[r * kDenseLd + lane + 32] = v[j][1];
            }
        }
    }
    __syncthreads();

    // ---- B. softmax statistics: thread (quarter, cell) -----------------------------------------------------
    const bool softmax = q.softmax && !q.legacy;
    if (softmax) {
        const int cell = tid & 63, qtr = tid >> 6;
        float m = -INFINITY;
        for (int c = qtr; c < C; c += 4) m = fmaxf(m, tile[(5 + c) * kDenseLd + cell]);
        part[qtr * kDenseCells + cell] = m;
        __syncthreads();
        m = fmaxf(fmaxf(part[cell], part[kDenseCells + cell]), fmaxf(part[2 * kDenseCells + cell], part[3 * kDenseCells + cell]));
        __syncthreads();
        float sum = 0.f;
        for (int c = qtr; c < C; c += 4) {
            // the exponential is kept in place: phase D only scales it by 1 / denominator
            const float e = ex2_approx(__fmul_rn(__fsub_rn(tile[(5 + c) * kDenseLd + cell], m), kLog2e));
            tile[(5 + c) * kDenseLd + cell] = e;
            sum = __fadd_rn(sum, e);
        }
        part[qtr * kDenseCells + cell] = sum;
        __syncthreads();
        if (qtr == 0) {
            const float tot = __fadd_rn(__fadd_rn(part[cell], part[kDenseCells + cell]),
                                        __fadd_rn(part[2 * kDenseCells + cell], part[3 * kDenseCells + cell]));
            st_m[cell] = m;
            st_r[cell] = __fdiv_rn(1.0f, tot);
        }
        __syncthreads();
    }

    // ---- C. box and objectness planes in place: thread = (plane, cell), so a warp runs ONE of the five
    //      formulas on 32 cells (with lanes across channels the first five lanes would serialise five different
    //      transcendental sequences for every cell) ------------------------------------------------------------
    for (int item = tid; item < 5 * kDenseCells; item += 256) {
        const int ch = item >> 6, cl = item & 63;
        if (cl >= ncell) continue;
        const int hw = cell0 + cl;
        const float t = tile[ch * kDenseLd + cl];
        float r;
        if (q.legacy) {
            if (ch == 0) r = __fmul_rn(__fadd_rn(sigmoid_ref(t), (float)(hw % q.in_w)), q.stride_w);        // yolo_loss.py:97,103
            else if (ch == 1) r = __fmul_rn(__fadd_rn(sigmoid_ref(t), (float)(hw / q.in_w)), q.stride_h);   // :98
            else if (ch == 2) r = __fmul_rn(__fmul_rn(expf(t), q.legacy_anchors[2 * a]), q.stride_w);       // :99
            else if (ch == 3) r = __fmul_rn(__fmul_rn(expf(t), q.legacy_anchors[2 * a + 1]), q.stride_h);   // :100
            else r = sigmoid_fast(t);
        } else {
            if (ch < 2) {
                const int gy_i = hw / sc.grid, gx_i = hw - gy_i * sc.grid;
                // xy = (sigmoid(t) + cxy*inw - 0.5) * stride, cxy = (idx + 0.5) / in_w     (yolo_forw.py:104-107,166)
                const float g = __fmul_rn(__fdiv_rn((float)(ch == 0 ? gx_i : gy_i) + 0.5f, sc.inw), sc.inw);
                r = __fmul_rn(__fsub_rn(__fadd_rn(sigmoid_ref(t), g), 0.5f), sc.stride);
            } else if (ch < 4) {
                // wh = exp(t) * cwh * inw * stride  (left to right)                          (:167)
                r = __fmul_rn(__fmul_rn(__fmul_rn(expf(t), sc.anc[a][ch - 2]), sc.inw), sc.stride);
            } else {
                r = sigmoid_fast(t);
            }
        }
        tile[ch * kDenseLd + cl] = r;
    }
    __syncthreads();

    // ---- D. output rows: one warp per cell, lanes across the 5+C channels ---------------------------------
    for (int cl = warp; cl < ncell; cl += 8) {
        const int hw = cell0 + cl;
        float* dst = q.legacy ? q.out + ((size_t)ba * (size_t)sc.hw + (size_t)hw) * (size_t)CH            // n = a*H*W + h*W + w
                              : q.out + ((size_t)b * (size_t)p.N + (size_t)sc.anchor_off + (size_t)hw * p.A + a) * (size_t)CH;
        const float rinv = softmax ? st_r[cl] : 0.f;
        for (int ch = lane; ch < CH; ch += 32) {
            const float t = tile[ch * kDenseLd + cl];
            float r = t;                                     // planes 0-4 are final already
            if (ch >= 5) r = softmax ? __fmul_rn(t, rinv) : sigmoid_fast(t);      // t: IDF-scaled logit, or its exponential
            dst[ch] = r;
        }
    }
}

static int launch_dense2(Dense2Params& q, cudaStream_t stream) {
    const DecodeParams& p = q.d;
    int t = 0;
    for (int s = 0; s < B200_MAX_SCALES; ++s) {
        q.cta_begin[s] = t;
        q.tiles[s] = 1;
        if (s < p.num_scales) {
            q.tiles[s] = cdiv(p.sc[s].hw, kDenseCells);
            t += p.B * p.A * q.tiles[s];
        }
        q.tiles_inv[s] = ((1ull << 40) + (unsigned long long)q.tiles[s] - 1) / (unsigned long long)q.tiles[s];
    }
    q.cta_begin[B200_MAX_SCALES] = t;
    q.a_inv = ((1ull << 40) + (unsigned long long)p.A - 1) / (unsigned long long)p.A;
    if ((long long)t * (long long)(p.A > q.tiles[0] ? p.A : q.tiles[0]) >= (1ll << 40)) return B200_ERR_INVALID;
    const size_t smem = (size_t)((5 + p.C) * kDenseLd + 6 * kDenseCells) * sizeof(float);
    static SmemOptIn optin;
    if (optin.ensure(k_decode_dense2, smem) != cudaSuccess) return B200_ERR_CUDA;
    k_decode_dense2<<<t, 256, smem, stream>>>(q);
    return cudaGetLastError() == cudaSuccess ? B200_OK : B200_ERR_CUDA;
}

// legacy YOLOLoss layer through the same kernel (one scale, own row order and xy/wh formulas)
int launch_legacy_dense(const float* head, int B, int A, int C, int H, int W, float stride_w, float stride_h,
                        const float* anchors_dev, float* out, cudaStream_t stream) {
    if (5 + C > kDenseMaxCh) return 1;
    Dense2Params q{};
    q.d.num_scales = 1; q.d.A = A; q.d.C = C; q.d.B = B; q.d.N = A * H * W; q.d.idf = nullptr;
    q.d.sc[0].head = head; q.d.sc[0].grid = W; q.d.sc[0].hw = H * W; q.d.sc[0].anchor_off = 0;
    q.d.sc[0].inw = (float)W; q.d.sc[0].stride = stride_w;
    q.out = out; q.softmax = 0; q.legacy = 1; q.in_w = W; q.stride_w = stride_w; q.stride_h = stride_h;
    q.legacy_anchors = anchors_dev;
    return launch_dense2(q, stream);
}

int launch_decode_dense(const DecodeParams& p, bool softmax, float* out, cudaStream_t stream) {
    if (5 + p.C <= kDenseMaxCh) {
        Dense2Params q2{};
        q2.d = p;
        q2.out = out;
        q2.softmax = softmax ? 1 : 0;
        return launch_dense2(q2, stream);
    }
    DenseParams q;
    q.d = p;
    q.out = out;
    q.softmax = softmax ? 1 : 0;
    int t = 0;
    for (int s = 0; s < p.num_scales; ++s) {
        q.tile_begin[s] = t;
        t += p.B * cdiv(p.sc[s].hw, kCells);
    }
    q.tile_begin[p.num_scales] = t;
    const size_t smem = (size_t)(2 * p.A * kCells + p.A * kChunk * (kCells + 1)) * sizeof(float);
    k_decode_dense<<<t, 256, smem, stream>>>(q);
    return cudaGetLastError() == cudaSuccess ? B200_OK : B200_ERR_CUDA;
}

}  // namespace b200
