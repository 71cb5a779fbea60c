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
        d.tiles = cdiv(d.hw, 32 * d.vec);
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
template <int VEC>
__device__ __forceinline__ void load_row(const float* p, bool in, float (&v)[VEC]) {
    if constexpr (VEC == 4) {
        float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
        if (in) r = ldg_stream_v4(p);
        v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w;
    } else {
        v[0] = in ? ldg_stream_f32(p) : 0.f;
    }
}

static constexpr float kLog2e = 1.4426950408889634f;

// One warp task: 32*VEC cells of one (scale, b, a).
template <int VEC, bool SOFTMAX, bool HAS_IDF, int U>
__device__ __forceinline__ void decode_filter_task(const DecodeParams& p, const ScaleDev& sc, int b,
                                                   int a, int tile, int lane) {
    const int hw0 = (tile * 32 + lane) * VEC;
    const bool in = hw0 < sc.hw;  // hw % VEC == 0, so a lane is entirely in or out
    const int C = p.C;
    const size_t plane = (size_t)sc.hw;
    const float* base = sc.head + ((size_t)(b * p.A + a) * (size_t)(5 + C)) * plane + (size_t)hw0;

    // --- objectness plane: which cells can still pass?  score = conf*maxp <= conf ------------
    float t4[VEC];
    load_row<VEC>(base + 4 * plane, in, t4);
    float conf[VEC];
    bool live[VEC];
    bool any_live = false;
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
        conf[k] = sigmoid_ref(t4[k]);
        live[k] = in && (conf[k] > p.thr);
        any_live |= live[k];
    }

    // --- class sweep: running max (first index wins) and softmax denominator -----------------
    float m[VEC], s[VEC];
    int arg[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) { m[k] = -INFINITY; s[k] = 0.f; arg[k] = 0; }

    const float* cls = base + 5 * plane;
    int c = 0;
    for (; c + U <= C; c += U) {
        float v[U][VEC];
#pragma unroll
        for (int u = 0; u < U; ++u) load_row<VEC>(cls + (size_t)(c + u) * plane, in, v[u]);
        if (any_live) {
            float w[U];
            if constexpr (HAS_IDF) {
#pragma unroll
                for (int u = 0; u < U; ++u) w[u] = __ldg(p.idf + c + u);
            }
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                if (live[k]) {
                    float x[U];
                    const float m_old = m[k];
                    float mk = m_old;
                    int ak = arg[k];
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        x[u] = HAS_IDF ? __fmul_rn(w[u], v[u][k]) : v[u][k];
                        if (x[u] > mk) { mk = x[u]; ak = c + u; }
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
        float v[VEC];
        load_row<VEC>(cls + (size_t)c * plane, in, v);
        if (any_live) {
            const float w = HAS_IDF ? __ldg(p.idf + c) : 1.0f;
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                if (live[k]) {
                    const float x = HAS_IDF ? __fmul_rn(w, v[k]) : v[k];
                    const float m_old = m[k];
                    if (x > m_old) { m[k] = x; arg[k] = c; }
                    if constexpr (SOFTMAX) {
                        float acc = __fmul_rn(s[k], ex2_approx(__fmul_rn(__fsub_rn(m_old, m[k]), kLog2e)));
                        s[k] = __fadd_rn(acc, ex2_approx(__fmul_rn(__fsub_rn(x, m[k]), kLog2e)));
                    }
                }
            }
        }
    }

    // --- box planes (last chunk of the stream) ------------------------------------------------
    float tb[4][VEC];
#pragma unroll
    for (int u = 0; u < 4; ++u) load_row<VEC>(base + (size_t)u * plane, in, tb[u]);

    // --- threshold + compaction ----------------------------------------------------------------
    float score[VEC];
    bool pass[VEC];
    unsigned ballots[VEC];
    int total = 0;
    const unsigned lt = (1u << lane) - 1u;
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
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
    for (int k = 0; k < VEC; ++k) {
        if (pass[k]) {
            const int slot = slot0 + before + __popc(ballots[k] & lt);
            if (slot < p.cap) {
                const int hw = hw0 + k;
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

template <bool SOFTMAX, bool HAS_IDF>
__global__ void __launch_bounds__(128, 8)
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
        decode_filter_task<4, SOFTMAX, HAS_IDF, 4>(p, sc, b, a, tile, lane);
    else
        decode_filter_task<1, SOFTMAX, HAS_IDF, 8>(p, sc, b, a, tile, lane);
}

int launch_decode_filter(const DecodeParams& p, bool softmax, cudaStream_t stream) {
    const int warps_per_block = 4;
    const int blocks = cdiv(p.total_tasks, warps_per_block);
    const bool idf = p.idf != nullptr;
    if (softmax) {
        if (idf) k_decode_filter<true, true><<<blocks, 32 * warps_per_block, 0, stream>>>(p);
        else     k_decode_filter<true, false><<<blocks, 32 * warps_per_block, 0, stream>>>(p);
    } else {
        if (idf) k_decode_filter<false, true><<<blocks, 32 * warps_per_block, 0, stream>>>(p);
        else     k_decode_filter<false, false><<<blocks, 32 * warps_per_block, 0, stream>>>(p);
    }
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

int launch_decode_dense(const DecodeParams& p, bool softmax, float* out, cudaStream_t stream) {
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
