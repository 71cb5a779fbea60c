// decode_ring.cu -- fused decode + filter as a persistent, TMA-fed stream (sm_100a).
//
// Same contract as k_decode_filter (decode.cu): replaces yolo_forw.py:93-119,163-176 +
// helper.py:203-217 + test_one_epoch.py:25-28,35 and reads every byte of the head tensors exactly
// once, whatever the input looks like.  What differs is who moves the bytes:
//
//   * persistent CTAs (one per SM) whose warps pull 64-cell tiles of one (scale, b, a) from a global
//     ticket counter, so SMs that are busy with another stream's NMS kernels simply take fewer tiles;
//   * every warp is its own producer and consumer.  Its elected lane describes the tile to the TMA
//     unit: the head tensor of a scale is a 2-D tensor map [B*A*(5+C) rows, H*W cells] and a tile is
//     two boxes of 32 cells x R rows (128 B rows, 128 B swizzle) fetched by two
//     cp.async.bulk.tensor.2d instructions (UTMALDG in SASS) that complete on the stage's mbarrier,
//     L2 evict-first since nothing is read twice.  Six warps x 23 KB stages keep ~100-140 KB per SM in
//     flight -- two to three times the ~45 KB the HBM latency x bandwidth product asks for -- and no
//     register or issue slot is spent on holding loads; there is no CTA-wide barrier after set-up;
//   * when its stage lands the warp wakes on the mbarrier.  A lane owns two cells: one sigmoid of the
//     objectness logit decides whether the cell can still pass (score = conf * max_c p <= conf).  The
//     few live cells of the tile (~1-3 %) are then swept warp-cooperatively out of shared memory --
//     lanes across classes (the swizzle spreads a column over 8 banks), exact two-pass softmax inside
//     a chunk, online rescale across chunks -- so there is no divergent per-lane class loop and the
//     SM's issue slots stay free for the NMS kernels of the neighbouring pipeline stages.  The stage
//     is refilled by the same warp the moment it has been read.
//
// Scales whose plane rows are not 16 B aligned (odd grids: 13x13, 19x19, ...) cannot be described by
// a tensor map; their tiles are staged row by row with 1-D cp.async.bulk copies from the aligned
// address below each row and the consumers add the row's 0-3 float shift when they index the stage.
// Class counts that do not fit one stage (LVIS, 1208 rows) stream through in chunks of R rows.
#include <cuda.h>
#include <math.h>

#include "decode.cuh"

namespace b200 {

// tiles are TC = 32 or 64 cells (template parameter of the kernel)
static constexpr int kBoxCells = 32;               // one TMA box: 32 cells (128 B) x R rows
// 1-D mode: floats per staged row = TC cells + <=3 floats of alignment shift, padded to a 16 B multiple
__host__ __device__ constexpr int row_stride_b(int tc) { return tc + 4; }
static constexpr int kMaxWarps = 8;                // warps per CTA (launch parameter, 1..8)
static constexpr int kMaxSlots = 32;               // warps x stages per warp
static constexpr int kLiveBatch = 4;               // live cells swept concurrently by a warp
static constexpr int kRowsPerLane = 3;             // class rows of one chunk per lane: chunks hold <= 96 rows
static constexpr float kLog2eR = 1.4426950408889634f;

struct __align__(64) RingParams {
    CUtensorMap tmap[B200_MAX_SCALES];     // by scale index; valid where tensor[s] != 0
    DecodeParams d;
    int tile_begin[B200_MAX_SCALES + 1];   // ticket ranges, in processing order (largest grid first)
    int order[B200_MAX_SCALES];            // processing slot -> scale index
    int tiles_per_ba[B200_MAX_SCALES];     // by scale index
    unsigned tpb_magic[B200_MAX_SCALES];   // by processing slot: ceil(2^32 / tiles_per_ba), 0 when it is 1
    float logit_screen;                    // objectness logits at or below this can never reach the threshold
    int tensor[B200_MAX_SCALES];           // 1: 2-D tensor-map boxes, 0: 1-D row copies
    int rows_per_chunk, chunks, slots_per_warp, total_tiles;
    int box_bytes;                         // ceil8(R) * 128: one swizzled box, 1024 B multiple
    int stage_bytes;                       // max over both staging modes, 1024 B multiple
    int* tile_counter;                     // zeroed by the host before every launch
};

__device__ __forceinline__ unsigned s_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    unsigned done = 0;
    while (!done)
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar,
                                         unsigned long long policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(policy) : "memory");
}
__device__ __forceinline__ void tma_2d_g2s(unsigned dst, const CUtensorMap* map, int c0, int c1, unsigned bar,
                                           unsigned long long policy) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
                 "[%0], [%1, {%3, %4}], [%2], %5;"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "l"(policy) : "memory");
}

// what a warp needs to know about the tile sitting in one of its stages (written at issue time)
struct __align__(16) TileMeta {
    int tile;      // ticket, < 0: no more tiles
    int s;         // scale index
    int ba;        // b * A + a
    int cell0;     // first cell of the tile inside the (b, a) plane
};

__device__ __forceinline__ TileMeta tile_geom(const RingParams& q, int tile, int tile_cells) {
    const DecodeParams& p = q.d;
    int k = 0;
#pragma unroll
    for (int i = 1; i < B200_MAX_SCALES; ++i)
        if (i < p.num_scales && tile >= q.tile_begin[i]) k = i;
    TileMeta g;
    g.tile = tile;
    g.s = q.order[k];
    const unsigned local = (unsigned)(tile - q.tile_begin[k]);
    // local / tiles_per_ba: multiply-high by ceil(2^32 / d) is exact while local * d < 2^32 (checked on the host,
    // true for every realistic shape); 0 = (d == 1), 1 = fall back to the division
    const unsigned magic = q.tpb_magic[k];
    g.ba = magic > 1u ? (int)__umulhi(local, magic) : magic == 1u ? (int)(local / (unsigned)q.tiles_per_ba[g.s]) : (int)local;
    g.cell0 = ((int)local - g.ba * q.tiles_per_ba[g.s]) * tile_cells;
    return g;
}

// The lane's class logits of one live cell: rows rl0 + lane + 32k of the staged chunk (k < kRowsPerLane),
// scaled by the class weights, -inf where the chunk has no such row.
template <bool TENSOR, bool HAS_IDF, int TC>
__device__ __forceinline__ void load_logits(const unsigned char* st, int box_bytes, int cell, int rl0, int nrows, int lane,
                                            unsigned shift0, unsigned hw_u, const float (&w)[kRowsPerLane],
                                            float (&x)[kRowsPerLane]) {
    const int rl = rl0 + lane;
    if (TENSOR) {
        // 128 B rows, 16 B chunks XOR-swizzled with (row & 7); rows 32 apart share the swizzle term
        const int cc = cell & 31;
        const unsigned char* ptr = st + (cell >> 5) * box_bytes + rl * 128 + ((((cc >> 2) ^ rl) & 7) << 4) + ((cc & 3) << 2);
#pragma unroll
        for (int k = 0; k < kRowsPerLane; ++k) {
            float v = 0.f;
            if (lane + 32 * k < nrows) v = *reinterpret_cast<const float*>(ptr + k * 32 * 128);
            x[k] = v;
        }
    } else {
        const float* f = reinterpret_cast<const float*>(st);
#pragma unroll
        for (int k = 0; k < kRowsPerLane; ++k) {
            float v = 0.f;
            // shift0 = (float index of staged row 0 of this chunk) mod 4; row rl sits rl*hw floats further
            if (lane + 32 * k < nrows) {
                const int r = rl + 32 * k;
                v = f[r * row_stride_b(TC) + (int)((shift0 + (unsigned)r * hw_u) & 3u) + cell];
            }
            x[k] = v;
        }
    }
#pragma unroll
    for (int k = 0; k < kRowsPerLane; ++k) {
        float v = HAS_IDF ? __fmul_rn(w[k], x[k]) : x[k];
        v = __fadd_rn(v, 0.f);                                   // -0 -> +0: one ordered key per value
        x[k] = lane + 32 * k < nrows ? v : -INFINITY;
    }
}

template <bool SOFTMAX, bool HAS_IDF, int TC>
__global__ void __launch_bounds__(32 * kMaxWarps, 1)
k_decode_filter_ring(const __grid_constant__ RingParams q) {
    extern __shared__ unsigned char ring_raw[];
    __shared__ __align__(8) unsigned long long full_bar[kMaxSlots];
    __shared__ TileMeta meta[kMaxSlots];

    const DecodeParams& p = q.d;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int NW = blockDim.x >> 5, SW = q.slots_per_warp;
    const int R = q.rows_per_chunk, CH = 5 + p.C;
    // swizzled boxes want 1024 B alignment
    unsigned char* ring = ring_raw + ((1024u - (s_u32(ring_raw) & 1023u)) & 1023u);
    float* idf_s = reinterpret_cast<float*>(ring + (size_t)NW * SW * q.stage_bytes);   // [C] class scale, staged once per CTA
    // the live cells of this warp's current tile: cell, argmax | max, sum, t_x, t_y, t_w, t_h, t_obj
    float* live_f = idf_s + ((p.C + 3) & ~3) + (size_t)warp * 9 * TC;
    int* my_cell = reinterpret_cast<int*>(live_f);
    int* my_arg = my_cell + TC;
    float* my_m = live_f + 2 * TC;
    float* my_s = live_f + 3 * TC;
    float* my_t = live_f + 4 * TC;          // [5][TC]

    if (tid == 0) {
        for (int i = 0; i < NW * SW; ++i) mbar_init(s_u32(&full_bar[i]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (HAS_IDF)
        for (int c = tid; c < p.C; c += blockDim.x) idf_s[c] = __ldg(p.idf + c);
    __syncthreads();

    // Every warp is its own producer and consumer: it owns SW stages, keeps SW loads in flight and
    // refills a stage as soon as it has finished reading it.  No CTA-wide barrier below this line.
    unsigned char* my_ring = ring + (size_t)warp * SW * q.stage_bytes;
    unsigned long long* my_bar = full_bar + warp * SW;
    TileMeta* my_meta = meta + warp * SW;
    unsigned long long policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));

    // ---- issue cursor ---------------------------------------------------------------------------------
    // One global ticket counter hands out tiles in memory order, so the ~600-1200 tiles in flight across
    // the chip at any moment are neighbours: they touch the same DRAM pages of the same plane rows.
    // (Splitting the counter into slices was measured: fewer same-address atomics, but 15-20 % slower
    // because that locality is lost.)  The ticket for the next tile is requested one tile ahead.
    int ticket = 0;                              // next tile of this warp (valid in lane 0)
    if (lane == 0) ticket = atomicAdd(q.tile_counter, 1);
    int islot = 0, i_chunk = 0;
    bool exhausted = false;
    TileMeta ig{};
    auto issue_next = [&]() {
        const int slot = islot;
        islot = islot + 1 == SW ? 0 : islot + 1;
        if (!exhausted && i_chunk == 0) {
            const int t = __shfl_sync(kFullMask, ticket, 0);
            if (t >= q.total_tiles) exhausted = true;
            else {
                if (lane == 0) ticket = atomicAdd(q.tile_counter, 1);
                ig = tile_geom(q, t, TC);
            }
        }
        if (exhausted) {
            if (lane == 0) my_meta[slot].tile = -1;
            return;
        }
        const ScaleDev& sc = p.sc[ig.s];
        const int ncell = min(TC, sc.hw - ig.cell0);
        const unsigned bar = s_u32(&my_bar[slot]);
        const unsigned dst0 = s_u32(my_ring + (size_t)slot * q.stage_bytes);
        const int r0 = i_chunk * R;
        if (q.tensor[ig.s]) {
            if (lane == 0) {
                const int nbox = (TC > kBoxCells && ncell > kBoxCells) ? 2 : 1;   // a box that starts past the row end is skipped
                my_meta[slot] = ig;
                mbar_arrive_expect_tx(bar, (unsigned)(nbox * R * kBoxCells * 4));
                for (int j = 0; j < nbox; ++j)
                    tma_2d_g2s(dst0 + (unsigned)(j * q.box_bytes), &q.tmap[ig.s], ig.cell0 + j * kBoxCells,
                               ig.ba * CH + r0, bar, policy);
            }
        } else {
            const size_t hw = (size_t)sc.hw;
            const float* base = sc.head + (size_t)ig.ba * (size_t)CH * hw + (size_t)ig.cell0;
            const int nr = min(R, CH - r0);
            // A row copy starts at the 16 B aligned address below the row and ends at the aligned address above it.
            // For the first row of the tensor that start lies before the buffer, for the last row the end lies behind
            // it: those (at most two) rows are copied with ordinary loads instead, so that nothing outside the
            // caller's buffer is ever touched.
            const uintptr_t t_lo = reinterpret_cast<uintptr_t>(sc.head);
            const uintptr_t t_hi = t_lo + (size_t)p.B * (size_t)p.A * (size_t)CH * hw * sizeof(float);
            unsigned bytes = 0;
            for (int r = lane; r < nr; r += 32) {
                const uintptr_t addr = reinterpret_cast<uintptr_t>(base + (size_t)(r0 + r) * hw);
                const unsigned nb = ((unsigned)(addr & 15u) + (unsigned)ncell * 4u + 15u) & ~15u;
                const uintptr_t lo = addr & ~(uintptr_t)15u;
                if (lo >= t_lo && lo + nb <= t_hi) bytes += nb;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) bytes += __shfl_xor_sync(kFullMask, bytes, o);
            if (lane == 0) {
                my_meta[slot] = ig;
                mbar_arrive_expect_tx(bar, bytes);
            }
            __syncwarp();
            for (int r = lane; r < nr; r += 32) {
                const float* row = base + (size_t)(r0 + r) * hw;
                const uintptr_t addr = reinterpret_cast<uintptr_t>(row);
                const unsigned nb = ((unsigned)(addr & 15u) + (unsigned)ncell * 4u + 15u) & ~15u;
                const uintptr_t lo = addr & ~(uintptr_t)15u;
                if (lo >= t_lo && lo + nb <= t_hi) {
                    bulk_g2s(dst0 + (unsigned)r * (row_stride_b(TC) * 4u), reinterpret_cast<const void*>(lo), nb, bar, policy);
                } else {
                    float* drow = reinterpret_cast<float*>(my_ring + (size_t)slot * q.stage_bytes) + r * row_stride_b(TC) +
                                  (int)((addr & 15u) >> 2);
                    for (int c = 0; c < ncell; ++c) drow[c] = ldg_stream_f32(row + c);
                }
            }
        }
        i_chunk = i_chunk + 1 == q.chunks ? 0 : i_chunk + 1;
    };
    for (int k = 0; k < SW; ++k) issue_next();

    // ---- consume ---------------------------------------------------------------------------------------
    const unsigned lt = (1u << lane) - 1u;
    const float screen = q.logit_screen;
    const int box_bytes = q.box_bytes;
    int nl = 0, chunk = 0;
    float w[kRowsPerLane];
#pragma unroll
    for (int k = 0; k < kRowsPerLane; ++k) w[k] = 1.0f;
    int cslot = 0;
    unsigned cphase = 0;
    for (bool first = true;; first = false) {
        const int slot = cslot;
        __syncwarp();
        const TileMeta g = my_meta[slot];
        if (g.tile < 0) break;
        const ScaleDev& sc = p.sc[g.s];
        const bool tensor = q.tensor[g.s] != 0;
        const unsigned hw_u = (unsigned)sc.hw;
        const int ncell = min(TC, sc.hw - g.cell0);
        const int r0 = chunk * R, nr = min(R, CH - r0);
        const int c_lo = max(r0, 5), c_hi = r0 + nr;           // class rows of this chunk
        // 1-D mode: float offset of staged row rl inside its 16 B aligned copy = (shift0 + rl*hw) mod 4
        const unsigned shift0 = (unsigned)(reinterpret_cast<uintptr_t>(sc.head + ((size_t)g.ba * (size_t)CH + (size_t)r0) * (size_t)hw_u +
                                                                       (size_t)g.cell0) >> 2);
        if (HAS_IDF && (first || q.chunks > 1)) {
#pragma unroll
            for (int k = 0; k < kRowsPerLane; ++k) {
                const int c = c_lo - 5 + lane + 32 * k;
                w[k] = c < p.C ? idf_s[c] : 1.0f;
            }
        }
        mbar_wait(s_u32(&my_bar[slot]), cphase);
        if (++cslot == SW) { cslot = 0; cphase ^= 1u; }
        const unsigned char* st = my_ring + (size_t)slot * q.stage_bytes;
        auto at = [&](int r, int cell) -> float {      // element (tensor row r, tile cell) of this stage
            const int rl = r - r0;
            if (tensor) {
                const int cc = cell & 31;
                return *reinterpret_cast<const float*>(st + (cell >> 5) * box_bytes + rl * 128 +
                                                       ((((cc >> 2) ^ rl) & 7) << 4) + ((cc & 3) << 2));
            }
            return reinterpret_cast<const float*>(st)[rl * row_stride_b(TC) + (int)((shift0 + (unsigned)rl * hw_u) & 3u) + cell];
        };

        if (chunk == 0) {
            // ---- objectness screen: score = conf * max_c p <= conf = sigmoid(t_obj), so a cell whose logit
            //      is below logit(thr) (minus a margin far wider than fp32 rounding) can never pass ----------
            nl = 0;
#pragma unroll
            for (int k = 0; k < TC / 32; ++k) {
                const int cell = lane + 32 * k;
                float t4 = 0.f;
                bool live = false;
                if (cell < ncell) {
                    t4 = at(4, cell);
                    live = t4 > screen;
                }
                const unsigned bal = __ballot_sync(kFullMask, live);
                if (live) {
                    const int i = nl + __popc(bal & lt);
                    my_cell[i] = cell; my_m[i] = -INFINITY; my_s[i] = 0.f; my_arg[i] = 0;
                    my_t[4 * TC + i] = t4;
#pragma unroll
                    for (int u = 0; u < 4; ++u) my_t[u * TC + i] = at(u, cell);
                }
                nl += __popc(bal);
            }
            __syncwarp();
        }

        // ---- class sweep of the live cells: lanes across the chunk's class rows, one pass over shared memory.
        //      kLiveBatch cells are in flight at once so that their load -> max -> redux -> exp -> shuffle
        //      chains overlap (a single warp has nothing else to hide that latency with). ----------------------
        if (c_lo < c_hi && nl > 0) {
            const int rl0 = c_lo - r0, nrows = c_hi - c_lo;
            for (int i0 = 0; i0 < nl; i0 += kLiveBatch) {
                float x[kLiveBatch][kRowsPerLane];
#pragma unroll
                for (int u = 0; u < kLiveBatch; ++u) {
                    const int cell = my_cell[min(i0 + u, nl - 1)];       // tail slots repeat the last cell, results unused
                    if (tensor) load_logits<true, HAS_IDF, TC>(st, box_bytes, cell, rl0, nrows, lane, shift0, hw_u, w, x[u]);
                    else        load_logits<false, HAS_IDF, TC>(st, box_bytes, cell, rl0, nrows, lane, shift0, hw_u, w, x[u]);
                }
                float m_new[kLiveBatch], s_new[kLiveBatch];
                int amin[kLiveBatch];
                bool up[kLiveBatch];
#pragma unroll
                for (int u = 0; u < kLiveBatch; ++u) {
                    float vmax = x[u][0];
                    int varg = 0;
#pragma unroll
                    for (int k = 1; k < kRowsPerLane; ++k)
                        if (x[u][k] > vmax) { vmax = x[u][k]; varg = k; }        // first maximum wins inside a lane
                    varg = c_lo - 5 + lane + 32 * varg;
                    const unsigned key = orderable(vmax);
                    const unsigned kmax = __reduce_max_sync(kFullMask, key);
                    unsigned mine = key == kmax ? (unsigned)varg : 0x7fffffffu;
                    vmax = from_orderable(kmax);
                    const float m_old = my_m[min(i0 + u, nl - 1)];
                    up[u] = vmax > m_old;                                    // strict: earlier chunks win ties
                    if (!SOFTMAX && vmax > 5.0f) {
                        // The reference takes the first maximum of the PROBABILITIES (test_one_epoch.py:35).  Where the
                        // fp32 sigmoid is flat -- it saturates at 1.0f above ~16.6 and its steps widen to 0.2 at 15 --
                        // smaller logits tie with the largest one and the lowest class index wins.
                        const float pm = sigmoid_ref(vmax);
#pragma unroll
                        for (int k = 0; k < kRowsPerLane; ++k)
                            if (x[u][k] > fminf(vmax, 17.0f) - 8.0f && sigmoid_ref(x[u][k]) == pm)
                                mine = min(mine, (unsigned)(c_lo - 5 + lane + 32 * k));
                        if (up[u] && m_old > -INFINITY && sigmoid_ref(m_old) == pm) up[u] = false;    // an earlier chunk already holds it
                    }
                    amin[u] = (int)__reduce_min_sync(kFullMask, mine);       // ... and across lanes
                    m_new[u] = up[u] ? vmax : m_old;
                    s_new[u] = 0.f;
                    if (SOFTMAX) {
                        float sum = 0.f;
#pragma unroll
                        for (int k = 0; k < kRowsPerLane; ++k)
                            sum = __fadd_rn(sum, ex2_approx(__fmul_rn(__fsub_rn(x[u][k], m_new[u]), kLog2eR)));
                        // s <- s * exp(m_old - m_new) + sum_c exp(x_c - m_new): the carry rides in lane 0's partial sum
                        if (chunk != 0 && lane == 0)
                            sum = __fadd_rn(sum, __fmul_rn(my_s[min(i0 + u, nl - 1)],
                                                           ex2_approx(__fmul_rn(__fsub_rn(m_old, m_new[u]), kLog2eR))));
                        s_new[u] = sum;
                    }
                }
                if (SOFTMAX) {
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
                        for (int u = 0; u < kLiveBatch; ++u) s_new[u] = __fadd_rn(s_new[u], __shfl_xor_sync(kFullMask, s_new[u], o));
                }
                __syncwarp();
                if (lane == 0) {
#pragma unroll
                    for (int u = 0; u < kLiveBatch; ++u) {
                        const int i = i0 + u;
                        if (i < nl) {
                            my_m[i] = m_new[u];
                            if (up[u]) my_arg[i] = amin[u];
                            if (SOFTMAX) my_s[i] = s_new[u];
                        }
                    }
                }
                __syncwarp();
            }
        }

        if (chunk == q.chunks - 1) {
            // ---- threshold + compaction: lane i finishes live cell i ---------------------------------------
            const int a = g.ba % p.A, b = g.ba / p.A;
            for (int i0 = 0; i0 < nl; i0 += 32) {
                const int i = i0 + lane;
                bool pass = false;
                float score = 0.f;
                if (i < nl) {
                    // max_c softmax = exp(0)/sum ; max_c sigmoid = sigmoid(max logit)
                    const float best = SOFTMAX ? __fdiv_rn(1.0f, my_s[i]) : sigmoid_ref(my_m[i]);
                    score = __fmul_rn(sigmoid_ref(my_t[4 * TC + i]), best);   // test_one_epoch.py:25
                    pass = score > p.thr;                                             // :26 (strict, fp32)
                }
                const unsigned bal = __ballot_sync(kFullMask, pass);
                if (bal == 0u) continue;
                int slot0 = 0;
                if (lane == 0) slot0 = atomicAdd(p.count + b, __popc(bal));
                slot0 = __shfl_sync(kFullMask, slot0, 0);
                if (!pass) continue;
                const int out = slot0 + __popc(bal & lt);
                if (out >= p.cap) { atomicOr(p.status, 1); continue; }
                const int hw = g.cell0 + my_cell[i];
                const int gy_i = hw / sc.grid, gx_i = hw - gy_i * sc.grid;
                // cxypwh[:, :2] = (idx + 0.5) / in_w                             (yolo_forw.py:104-107)
                const float cx = __fdiv_rn((float)gx_i + 0.5f, sc.inw);
                const float cy = __fdiv_rn((float)gy_i + 0.5f, sc.inw);
                // xy = (sigmoid(t) + cxy*inw - 0.5) * stride                     (:166)
                const float bx = __fmul_rn(__fsub_rn(__fadd_rn(sigmoid_ref(my_t[0 * TC + i]), __fmul_rn(cx, sc.inw)), 0.5f), sc.stride);
                const float by = __fmul_rn(__fsub_rn(__fadd_rn(sigmoid_ref(my_t[1 * TC + i]), __fmul_rn(cy, sc.inw)), 0.5f), sc.stride);
                // wh = exp(t) * cwh * inw * stride  (left to right)               (:167)
                const float bw = __fmul_rn(__fmul_rn(__fmul_rn(expf(my_t[2 * TC + i]), sc.anc[a][0]), sc.inw), sc.stride);
                const float bh = __fmul_rn(__fmul_rn(__fmul_rn(expf(my_t[3 * TC + i]), sc.anc[a][1]), sc.inw), sc.stride);
                const Box bb = abs_coord(bx, by, bw, bh);                           // helper.py:203
                float4* d4 = reinterpret_cast<float4*>(p.slab + (size_t)b * (size_t)p.cap + (size_t)out);
                d4[0] = make_float4(bb.x1, bb.y1, bb.x2, bb.y2);
                d4[1] = make_float4(score, __int_as_float(my_arg[i]),
                                    __int_as_float(sc.anchor_off + hw * p.A + a), 0.f);
            }
        }
        chunk = chunk + 1 == q.chunks ? 0 : chunk + 1;
        // every lane's reads of the stage are done; order them before the async-proxy refill
        __syncwarp();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        issue_next();
    }
}

// Launch configuration knobs (b200_debug_set_ring): warps per CTA, stages per warp, CTAs per SM, tile cells.
// Default = the best stand-alone launch (8 warps x 64-cell tiles: 93 us for the C2 batch); a caller that keeps
// several launches in flight (bench.py: three) uses 4 warps x 32-cell tiles per launch instead.
static int g_ring_warps = 8;
static int g_ring_slots = 1;
static int g_ring_ctas_per_sm = 1;
static int g_ring_tile_cells = 64;
static const int kStageBudget = 23 * 1024;         // for 64-cell tiles; halves with the tile
void ring_set_tile_cells(int tc) {
    if (tc == 32 || tc == 64) g_ring_tile_cells = tc;
}
void ring_set_tuning(int warps, int slots_per_warp, int ctas_per_sm) {
    if (warps >= 1 && warps <= kMaxWarps) g_ring_warps = warps;
    if (slots_per_warp >= 1 && slots_per_warp * g_ring_warps <= kMaxSlots) g_ring_slots = slots_per_warp;
    if (ctas_per_sm >= 1 && ctas_per_sm <= 4) g_ring_ctas_per_sm = ctas_per_sm;
}

namespace {
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* f = nullptr;
        cudaDriverEntryPointQueryResult st;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &st) == cudaSuccess &&
            st == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(f);
    }
    return fn;
}
}  // namespace

// returns B200_OK, or 1 when the configuration cannot be launched (caller falls back to the register path)
int launch_decode_filter_ring(const DecodeParams& p, bool softmax, int* tile_counter, cudaStream_t stream) {
    if (!tile_counter) return 1;
    RingParams q;
    q.d = p;
    const int CH = 5 + p.C;
    // rows per chunk: as many as the stage budget holds, evened out over the chunks
    const int TC = g_ring_tile_cells;
    int rmax = kStageBudget / (row_stride_b(64) * 4);
    if (rmax > 32 * kRowsPerLane) rmax = 32 * kRowsPerLane;   // a lane holds kRowsPerLane class rows of a chunk
    if (rmax < 8) return 1;
    q.chunks = cdiv(CH, rmax);
    const int R = cdiv(CH, q.chunks);
    q.rows_per_chunk = R;
    q.slots_per_warp = g_ring_slots;
    q.tile_counter = tile_counter;
    {   // sigmoid(t) > thr  =>  t > logit(thr) - delta, delta chosen so that the implied relative gap in conf
        // ((1 - thr) * delta >= 1e-4) dwarfs the rounding of the fp32 sigmoid (~1e-6)
        const double thr = (double)p.thr;
        if (!(thr > 0.0)) q.logit_screen = -INFINITY;            // thr <= 0 (or NaN): every cell stays live
        else if (thr >= 1.0) q.logit_screen = INFINITY;          // conf <= 1: nothing can pass
        else q.logit_screen = (float)(log(thr) - log1p(-thr) - (0.01 + 1e-4 / (1.0 - thr))) - 1e-6f;
    }
    q.box_bytes = ((R + 7) / 8) * 8 * 128;
    const int stage_a = (TC / kBoxCells) * q.box_bytes, stage_b = R * row_stride_b(TC) * 4;
    q.stage_bytes = (int)align_up((size_t)(stage_a > stage_b ? stage_a : stage_b), 1024);

    // processing order: largest grid first, so the last tickets are the small tiles
    int idx[B200_MAX_SCALES];
    for (int s = 0; s < p.num_scales; ++s) idx[s] = s;
    for (int i = 0; i < p.num_scales; ++i)
        for (int j = i + 1; j < p.num_scales; ++j)
            if (p.sc[idx[j]].hw > p.sc[idx[i]].hw) { const int t = idx[i]; idx[i] = idx[j]; idx[j] = t; }
    int t = 0;
    for (int k = 0; k < B200_MAX_SCALES; ++k) {
        q.tile_begin[k] = t;
        q.order[k] = 0;
        q.tpb_magic[k] = 0u;
        if (k < p.num_scales) {
            const int s = idx[k];
            q.order[k] = s;
            q.tiles_per_ba[s] = cdiv(p.sc[s].hw, TC);
            const unsigned long long d = (unsigned long long)q.tiles_per_ba[s], n_max = (unsigned long long)p.B * p.A * d;
            q.tpb_magic[k] = d <= 1 ? 0u : (n_max * d < 0x100000000ull ? (unsigned)((0x100000000ull + d - 1) / d) : 1u);
            t += p.B * p.A * q.tiles_per_ba[s];
        }
    }
    q.tile_begin[B200_MAX_SCALES] = t;
    q.total_tiles = t;
    for (int s = 0; s < B200_MAX_SCALES; ++s) {
        q.tensor[s] = 0;
        if (s >= p.num_scales) { q.tiles_per_ba[s] = 1; continue; }
        const ScaleDev& sc = p.sc[s];
        const bool aligned = (sc.hw % 4 == 0) && ((reinterpret_cast<uintptr_t>(sc.head) & 15u) == 0);
        EncodeTiledFn enc = aligned ? encode_tiled() : nullptr;
        if (!enc) continue;
        const cuuint64_t dims[2] = {(cuuint64_t)sc.hw, (cuuint64_t)p.B * (cuuint64_t)p.A * (cuuint64_t)CH};
        const cuuint64_t strides[1] = {(cuuint64_t)sc.hw * sizeof(float)};
        const cuuint32_t box[2] = {(cuuint32_t)kBoxCells, (cuuint32_t)R};
        const cuuint32_t estr[2] = {1, 1};
        const CUresult r = enc(&q.tmap[s], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(sc.head), dims, strides,
                               box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        q.tensor[s] = r == CUDA_SUCCESS ? 1 : 0;
    }

    const int warps = g_ring_warps, threads = 32 * warps;
    const size_t smem = (size_t)warps * q.slots_per_warp * q.stage_bytes + align_up((size_t)p.C * sizeof(float), 16) +
                        (size_t)warps * 9 * TC * sizeof(float) + 1024;
    if (smem > 220 * 1024) return 1;
    const int sms = current_sm_count();
    int ctas = sms * g_ring_ctas_per_sm;
    if (ctas > q.total_tiles) ctas = q.total_tiles;
    const bool idf = p.idf != nullptr;
    typedef void (*Kern)(const RingParams);
    static const Kern kerns[8] = {
        k_decode_filter_ring<false, false, 32>, k_decode_filter_ring<false, true, 32>,
        k_decode_filter_ring<true, false, 32>,  k_decode_filter_ring<true, true, 32>,
        k_decode_filter_ring<false, false, 64>, k_decode_filter_ring<false, true, 64>,
        k_decode_filter_ring<true, false, 64>,  k_decode_filter_ring<true, true, 64>};
    static SmemOptIn optin[8];
    const int which = (TC == 64 ? 4 : 0) + (softmax ? 2 : 0) + (idf ? 1 : 0);
    if (optin[which].ensure(kerns[which], smem) != cudaSuccess) return B200_ERR_CUDA;
    kerns[which]<<<ctas, threads, smem, stream>>>(q);
    return cudaGetLastError() == cudaSuccess ? B200_OK : B200_ERR_CUDA;
}

}  // namespace b200
