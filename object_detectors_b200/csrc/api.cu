// api.cu -- the extern "C" boundary of libb200det.so (see include/b200det.h).
// Host-side only: argument validation, workspace carving, kernel sequencing.  No allocation on
// the stream-ordered entry points; the *_host convenience entry owns a small device pool.
#include <mutex>

#include "decode.cuh"
#include "decode_ring.cuh"
#include "nms.cuh"

namespace b200 {
int launch_decode_filter(const DecodeParams& p, bool softmax, int gate, cudaStream_t stream);
int launch_decode_filter_ring(const DecodeParams& p, bool softmax, int* tile_counter, cudaStream_t stream);
void ring_set_tuning(int warps, int slots_per_warp, int ctas_per_sm);
void ring_set_tile_cells(int tc);
int launch_decode_dense(const DecodeParams& p, bool softmax, float* out, cudaStream_t stream);
int launch_box_iou(const float*, int, const float*, int, int, int, float*, cudaStream_t);
int launch_box_iou_pair(const float*, const float*, int, int, int, float*, cudaStream_t);
int launch_box_iou_pair_bwd(const float*, const float*, const float*, int, int, int, float*, float*, cudaStream_t);
int launch_iou_match(const float*, const int*, int, int, const float*, int, int, float, long long*,
                     uint8_t*, unsigned long long*, cudaStream_t);
int launch_abs_coord(const float* in, long long n, float* out, cudaStream_t st);
int launch_boxcoder(const float* rel, const float* boxes, long long n, int k, const float* weights, float clip,
                    float* out, cudaStream_t st);
int launch_boxcoder_encode(const float* ref, const float* prop, long long n, const float* weights, float* out,
                           cudaStream_t st);
int launch_matcher(const float* q, int M, int N, float high, float low, int allow_low_quality, long long* matches,
                   long long* all_matches_ws, cudaStream_t st);
int launch_rpn_filter(const float* objectness, const float* deltas, const float* anchors, const float* proposals, int batch,
                      int total, const int* level_sizes_host, int num_levels, const float* image_hw,
                      int pre_k, int post_k, double nms_thr, float score_thr, float min_size, int nms_mode,
                      float* out_boxes, float* out_scores, int* out_index, int* out_count,
                      void* workspace, size_t workspace_bytes, cudaStream_t stream);
size_t rpn_workspace_bytes(int batch, int total, int num_levels, int pre_k);
int launch_rpn_filter_ex(const float* objectness, const float* deltas, const float* anchors, const float* proposals, int batch,
                         int total, const int* level_sizes_host, int num_levels, const float* image_hw,
                         int pre_k, int post_k, double nms_thr, float score_thr, float min_size, int nms_mode,
                         float* out_boxes, float* out_scores, int* out_index, int* out_count,
                         void* workspace, size_t workspace_bytes, cudaStream_t stream, int classes,
                         const float* class_scale, int* out_labels);
size_t rpn_topk_workspace_bytes(int batch, int total, int num_levels, int pre_k);
int launch_rpn_topk(const float* objectness, int batch, int total, const int* level_sizes_host, int num_levels, int pre_k,
                    long long* out_index, void* workspace, size_t workspace_bytes, cudaStream_t stream);

namespace {

struct Carver {
    unsigned char* p;
    size_t left;
    bool ok = true;
    template <typename T>
    T* take(size_t count) {
        const size_t bytes = align_up(count * sizeof(T), 256);
        if (bytes > left) { ok = false; return nullptr; }
        T* r = reinterpret_cast<T*>(p);
        p += bytes;
        left -= bytes;
        return r;
    }
};

struct YoloWs {
    int* count;
    int* ticket;        // tile ticket counter of the ring decode kernel (directly behind `count`)
    Cand* slab;
    float4* cbox;
    float* cscore;
    int* clabel;
    int* canchor;
    void* nms;          // scratch of the NMS kernels (nms_carve_scratch)
    size_t nms_bytes;
};

size_t yolo_ws_layout(int batch, int cap, void* base, size_t bytes, YoloWs* w) {
    // base == nullptr: size query
    const size_t T = (size_t)batch * (size_t)cap;
    Carver c{reinterpret_cast<unsigned char*>(base), base ? bytes : (size_t)-1};
    YoloWs tmp;
    YoloWs& o = w ? *w : tmp;
    const unsigned char* start = c.p;
    o.count = c.take<int>((size_t)batch);
    o.ticket = c.take<int>(kTicketInts);
    o.slab = c.take<Cand>(T);
    o.cbox = c.take<float4>(T);
    o.cscore = c.take<float>(T);
    o.clabel = c.take<int>(T);
    o.canchor = c.take<int>(T);
    o.nms_bytes = nms_scratch_bytes(T, (size_t)batch, (size_t)cap);
    o.nms = c.take<unsigned char>(o.nms_bytes);
    if (!c.ok) return 0;
    return (size_t)(c.p - start);
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace
}  // namespace b200


namespace b200 { extern cudaEvent_t g_nms_timeline[3]; extern long long* g_resolve_prof; extern long long* g_rpn_prof; extern int g_serial_split; extern long long g_batched_nms_auto_limit; extern int g_resolve_threads, g_resolve_smem_kb, g_nms_force_general; }
using namespace b200;

extern "C" {

int b200_abi_version(void) { return B200_ABI_VERSION; }

const char* b200_error_string(int code) {
    switch (code) {
        case B200_OK: return "ok";
        case B200_ERR_INVALID: return "invalid argument";
        case B200_ERR_CUDA: return "CUDA runtime error (no device, launch failure or out of memory)";
        case B200_ERR_WORKSPACE: return "workspace too small or misaligned";
        default: return "unknown error";
    }
}

int b200_device_info(int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor) {
    int dev = 0;
    B200_CUDA_TRY(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    B200_CUDA_TRY(cudaGetDeviceProperties(&prop, dev));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    return B200_OK;
}

// ------------------------------------------------------------------------------------------ YOLO
int b200_yolo_decode_dense(const b200_yolo_layout* layout, const float* const* heads,
                           const float* idf, float* out, void* stream) {
    if (!out) return B200_ERR_INVALID;
    DecodeParams p{};
    const int rc = make_decode_params(layout, heads, idf, &p);
    if (rc != B200_OK) return rc;
    return launch_decode_dense(p, layout->softmax != 0, out, static_cast<cudaStream_t>(stream));
}

size_t b200_yolo_workspace_bytes(const b200_yolo_layout* layout, int32_t capacity) {
    if (!layout || capacity < 1 || layout->batch < 1) return 0;
    return yolo_ws_layout(layout->batch, capacity, nullptr, 0, nullptr) + 256;
}

// Optional profiling hook (bench.py): events recorded on the call's stream right before / after
// the fused decode+filter kernel, so its duration can be measured inside a timed region.
static int g_decode_variant = B200_DECODE_RING;
int b200_set_decode_variant(int variant) {
    if (variant < B200_DECODE_GATED || variant > B200_DECODE_RING || variant == 2) return B200_ERR_INVALID;   // 2 was the retired BULK prototype
    g_decode_variant = variant;
    return B200_OK;
}
int b200_debug_set_ring(int warps, int stages_per_warp, int ctas_per_sm) {
    if (ctas_per_sm >= 100) {          // hundreds digit selects the tile: 1xx = 32 cells, otherwise 64
        ring_set_tile_cells(32);
        ctas_per_sm -= 100;
    } else if (ctas_per_sm > 0) {
        ring_set_tile_cells(64);
    }
    ring_set_tuning(warps, stages_per_warp, ctas_per_sm);
    return B200_OK;
}
int b200_debug_set_timeline(void* after_plan, void* after_pairs, void* after_resolve) {
    b200::g_nms_timeline[0] = static_cast<cudaEvent_t>(after_plan);
    b200::g_nms_timeline[1] = static_cast<cudaEvent_t>(after_pairs);
    b200::g_nms_timeline[2] = static_cast<cudaEvent_t>(after_resolve);
    return B200_OK;
}
int b200_set_batched_nms_auto_limit(int64_t numel) {
    if (numel < 0) return B200_ERR_INVALID;
    b200::g_batched_nms_auto_limit = numel;
    return B200_OK;
}
int b200_debug_set_resolve(int threads, int smem_kb) {
    if (threads >= 64 && threads <= 1024 && (threads & 31) == 0) b200::g_resolve_threads = threads;
    if (smem_kb >= 16 && smem_kb <= 200) b200::g_resolve_smem_kb = smem_kb;
    return B200_OK;
}
int b200_debug_set_nms_path(int general) { b200::g_nms_force_general = general < 0 ? -1 : (general > 2 ? 1 : general); return B200_OK; }
int b200_debug_set_resolve_prof(void* buf) { b200::g_resolve_prof = static_cast<long long*>(buf); return B200_OK; }
int b200_debug_set_serial_split(int boxes) { b200::g_serial_split = boxes < 1 ? 1 : boxes; return B200_OK; }
int b200_debug_set_rpn_prof(void* buf) { b200::g_rpn_prof = static_cast<long long*>(buf); return B200_OK; }
static void* g_ev_decode_begin = nullptr;
static void* g_ev_decode_end = nullptr;
int b200_debug_set_decode_events(void* ev_begin, void* ev_end) {
    g_ev_decode_begin = ev_begin;
    g_ev_decode_end = ev_end;
    return B200_OK;
}

// phase 1: slab cursors + ticket counter cleared, fused decode+filter into the slab
static int yolo_decode_phase(const b200_yolo_layout* layout, const float* const* heads, const float* idf,
                             float conf_thr, int capacity, int* count_buf, int* status, const YoloWs& w,
                             cudaStream_t st, int* anchor_space) {
    DecodeParams p{};
    const int rc = make_decode_params(layout, heads, idf, &p);
    if (rc != B200_OK) return rc;
    p.thr = conf_thr;
    p.slab = w.slab;
    p.cap = capacity;
    p.count = count_buf;
    p.status = status;
    if (anchor_space) *anchor_space = p.N;
    if (count_buf == w.count) {
        // slab cursors and the ring kernel's ticket counter sit in one block: one memset node
        B200_CUDA_TRY(cudaMemsetAsync(w.count, 0, (size_t)(reinterpret_cast<unsigned char*>(w.ticket + kTicketInts) -
                                                           reinterpret_cast<unsigned char*>(w.count)), st));
    } else {
        B200_CUDA_TRY(cudaMemsetAsync(count_buf, 0, sizeof(int) * (size_t)layout->batch, st));
        if (g_decode_variant == B200_DECODE_RING) B200_CUDA_TRY(cudaMemsetAsync(w.ticket, 0, sizeof(int) * kTicketInts, st));
    }
    if (g_ev_decode_begin) B200_CUDA_TRY(cudaEventRecord(static_cast<cudaEvent_t>(g_ev_decode_begin), st));
    int rc2 = 1;
    if (g_decode_variant == B200_DECODE_RING) rc2 = launch_decode_filter_ring(p, layout->softmax != 0, w.ticket, st);
    if (rc2 == 1)   // register path asked for, or the staged tile does not fit in shared memory
        rc2 = launch_decode_filter(p, layout->softmax != 0, g_decode_variant == B200_DECODE_GATED ? 1 : 0, st);
    if (rc2 != B200_OK) return rc2;
    if (g_ev_decode_end) B200_CUDA_TRY(cudaEventRecord(static_cast<cudaEvent_t>(g_ev_decode_end), st));
    return B200_OK;
}

// phase 2: order + NMS of the slab
static int yolo_nms_phase(const b200_yolo_layout* layout, int capacity, const int* count_buf, NmsParams& np,
                          const YoloWs& w, int anchor_space, cudaStream_t st) {
    if (!nms_carve_scratch(&np, (size_t)layout->batch * capacity, (size_t)layout->batch, (size_t)capacity,
                           w.nms, w.nms_bytes))
        return B200_ERR_WORKSPACE;
    np.slab = w.slab;
    np.count = count_buf;
    np.cap = capacity;
    np.from_slab = 1;
    np.anchor_space = anchor_space;
    np.max_seg = capacity;
    return launch_nms(np, layout->batch, st);
}

static int yolo_run(const b200_yolo_layout* layout, const float* const* heads, const float* idf,
                    float conf_thr, int capacity, int* count_buf, NmsParams& np, const YoloWs& w,
                    cudaStream_t st) {
    int anchor_space = 0;
    const int rc = yolo_decode_phase(layout, heads, idf, conf_thr, capacity, count_buf, np.status, w, st, &anchor_space);
    if (rc != B200_OK) return rc;
    return yolo_nms_phase(layout, capacity, count_buf, np, w, anchor_space, st);
}

int b200_yolo_decode_filter(const b200_yolo_layout* layout, const float* const* heads,
                            const float* idf, float conf_thr, int32_t capacity, float* cand_box,
                            float* cand_score, int32_t* cand_label, int32_t* cand_anchor,
                            int32_t* cand_count, int32_t* status, void* workspace,
                            size_t workspace_bytes, void* stream) {
    if (!layout || !cand_box || !cand_score || !cand_label || !cand_anchor || !cand_count || !status ||
        capacity < 1 || !aligned16(cand_box))
        return B200_ERR_INVALID;
    if (!workspace || (reinterpret_cast<uintptr_t>(workspace) & 255u)) return B200_ERR_WORKSPACE;
    YoloWs w;
    if (!yolo_ws_layout(layout->batch, capacity, workspace, workspace_bytes, &w)) return B200_ERR_WORKSPACE;
    NmsParams np{};
    np.mode = -1;
    np.status = status;
    np.cbox = reinterpret_cast<float4*>(cand_box);
    np.cscore = cand_score;
    np.clabel = cand_label;
    np.canchor = cand_anchor;
    // the caller's count array doubles as the atomic slab cursor
    return yolo_run(layout, heads, idf, conf_thr, capacity, cand_count, np, w, static_cast<cudaStream_t>(stream));
}

int b200_yolo_postprocess(const b200_yolo_layout* layout, const float* const* heads,
                          const float* idf, float conf_thr, double nms_thr, int32_t nms_mode,
                          int32_t capacity, int32_t max_det, float* det, int32_t* det_keep,
                          int32_t* det_anchor, int32_t* det_count, int32_t* cand_count,
                          int32_t* status, void* workspace, size_t workspace_bytes, void* stream) {
    if (!layout || !det || !det_count || !status || capacity < 1 || max_det < 1)
        return B200_ERR_INVALID;
    if (nms_mode < B200_NMS_MAJORITY || nms_mode > B200_NMS_TV_AUTO) return B200_ERR_INVALID;
    if (!workspace || (reinterpret_cast<uintptr_t>(workspace) & 255u)) return B200_ERR_WORKSPACE;
    YoloWs w;
    if (!yolo_ws_layout(layout->batch, capacity, workspace, workspace_bytes, &w)) return B200_ERR_WORKSPACE;
    NmsParams np{};
    np.mode = nms_mode;
    np.thr_f = (float)nms_thr;
    np.thr_d = nms_thr;
    np.status = status;
    np.cbox = w.cbox; np.cscore = w.cscore; np.clabel = w.clabel; np.canchor = w.canchor;
    np.det = det; np.det_keep = det_keep; np.det_anchor = det_anchor; np.det_count = det_count;
    np.cand_count_out = cand_count;
    np.max_det = max_det;
    np.serial = 1;                     // decode and NMS of ONE batch back to back on one stream: nothing to co-reside with
    return yolo_run(layout, heads, idf, conf_thr, capacity, w.count, np, w, static_cast<cudaStream_t>(stream));
}

int b200_yolo_postprocess_decode(const b200_yolo_layout* layout, const float* const* heads, const float* idf,
                                 float conf_thr, int32_t capacity, int32_t* status, void* workspace,
                                 size_t workspace_bytes, void* stream) {
    if (!layout || !status || capacity < 1) return B200_ERR_INVALID;
    if (!workspace || (reinterpret_cast<uintptr_t>(workspace) & 255u)) return B200_ERR_WORKSPACE;
    YoloWs w;
    if (!yolo_ws_layout(layout->batch, capacity, workspace, workspace_bytes, &w)) return B200_ERR_WORKSPACE;
    return yolo_decode_phase(layout, heads, idf, conf_thr, capacity, w.count, status, w,
                             static_cast<cudaStream_t>(stream), nullptr);
}

int b200_yolo_postprocess_nms(const b200_yolo_layout* layout, double nms_thr, int32_t nms_mode, int32_t capacity,
                              int32_t max_det, float* det, int32_t* det_keep, int32_t* det_anchor,
                              int32_t* det_count, int32_t* cand_count, int32_t* status, void* workspace,
                              size_t workspace_bytes, void* stream) {
    if (!layout || !det || !det_count || !status || capacity < 1 || max_det < 1) return B200_ERR_INVALID;
    if (layout->num_scales < 1 || layout->num_scales > B200_MAX_SCALES) return B200_ERR_INVALID;
    if (nms_mode < B200_NMS_MAJORITY || nms_mode > B200_NMS_TV_AUTO) return B200_ERR_INVALID;
    if (!workspace || (reinterpret_cast<uintptr_t>(workspace) & 255u)) return B200_ERR_WORKSPACE;
    YoloWs w;
    if (!yolo_ws_layout(layout->batch, capacity, workspace, workspace_bytes, &w)) return B200_ERR_WORKSPACE;
    NmsParams np{};
    np.mode = nms_mode;
    np.thr_f = (float)nms_thr;
    np.thr_d = nms_thr;
    np.status = status;
    np.cbox = w.cbox; np.cscore = w.cscore; np.clabel = w.clabel; np.canchor = w.canchor;
    np.det = det; np.det_keep = det_keep; np.det_anchor = det_anchor; np.det_count = det_count;
    np.cand_count_out = cand_count;
    np.max_det = max_det;
    int anchor_space = 0;
    for (int s = 0; s < layout->num_scales; ++s) anchor_space += layout->grid[s] * layout->grid[s] * layout->num_anchors;
    return yolo_nms_phase(layout, capacity, w.count, np, w, anchor_space, static_cast<cudaStream_t>(stream));
}

// ----------------------------------------------------------------------------- host-buffer e2e
namespace {
struct HostPool {
    std::mutex mu;
    void* heads[B200_MAX_SCALES] = {nullptr, nullptr, nullptr, nullptr};
    size_t head_bytes[B200_MAX_SCALES] = {0, 0, 0, 0};
    void* idf = nullptr; size_t idf_bytes = 0;
    void* ws = nullptr; size_t ws_bytes = 0;
    void* out = nullptr; size_t out_bytes = 0;
    cudaStream_t copy = nullptr, compute = nullptr;
    cudaEvent_t ev[64];
    bool init = false;
    cudaError_t grow(void** p, size_t* have, size_t want) {
        if (*have >= want) return cudaSuccess;
        if (*p) cudaFree(*p);
        *p = nullptr; *have = 0;
        cudaError_t e = cudaMalloc(p, want);
        if (e == cudaSuccess) *have = want;
        return e;
    }
};
HostPool g_pool;
}  // namespace

int b200_yolo_postprocess_host(const b200_yolo_layout* layout, const float* const* heads_host,
                               const float* idf_host, float conf_thr, double nms_thr,
                               int32_t nms_mode, int32_t capacity, int32_t max_det,
                               float* det_host, int32_t* det_keep_host, int32_t* det_count_host,
                               int32_t* status_host) {
    if (!layout || !heads_host || !det_host || !det_count_host || capacity < 1 || max_det < 1)
        return B200_ERR_INVALID;
    HostPool& P = g_pool;
    std::lock_guard<std::mutex> lock(P.mu);
    const int B = layout->batch, A = layout->num_anchors, CH = 5 + layout->num_classes;
    if (!P.init) {
        B200_CUDA_TRY(cudaStreamCreateWithFlags(&P.copy, cudaStreamNonBlocking));
        B200_CUDA_TRY(cudaStreamCreateWithFlags(&P.compute, cudaStreamNonBlocking));
        for (auto& e : P.ev) B200_CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        P.init = true;
    }
    size_t per_img[B200_MAX_SCALES];
    for (int s = 0; s < layout->num_scales; ++s) {
        per_img[s] = (size_t)A * CH * layout->grid[s] * layout->grid[s] * sizeof(float);
        B200_CUDA_TRY(P.grow(&P.heads[s], &P.head_bytes[s], per_img[s] * B));
    }
    // sub-batches: copies of chunk i+1 overlap the kernels of chunk i
    const int chunk = B >= 16 ? (B + 7) / 8 : B;
    const int nchunk = (B + chunk - 1) / chunk;
    if (nchunk > 64) return B200_ERR_INVALID;
    b200_yolo_layout sub = *layout;
    sub.batch = chunk;
    const size_t ws_one = b200_yolo_workspace_bytes(&sub, capacity);
    B200_CUDA_TRY(P.grow(&P.ws, &P.ws_bytes, ws_one * 2));
    const size_t det_b = (size_t)B * max_det * 6 * sizeof(float);
    const size_t keep_b = (size_t)B * max_det * sizeof(int);
    const size_t cnt_b = align_up((size_t)B * sizeof(int), 256);
    B200_CUDA_TRY(P.grow(&P.out, &P.out_bytes, det_b + keep_b + cnt_b + 256));
    float* d_det = reinterpret_cast<float*>(P.out);
    int* d_keep = reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(P.out) + det_b);
    int* d_cnt = reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(P.out) + det_b + keep_b);
    int* d_status = reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(P.out) + det_b + keep_b + cnt_b);
    const float* d_idf = nullptr;
    if (idf_host) {
        B200_CUDA_TRY(P.grow(&P.idf, &P.idf_bytes, sizeof(float) * layout->num_classes));
        B200_CUDA_TRY(cudaMemcpyAsync(P.idf, idf_host, sizeof(float) * layout->num_classes,
                                      cudaMemcpyHostToDevice, P.copy));
        d_idf = reinterpret_cast<const float*>(P.idf);
    }
    B200_CUDA_TRY(cudaMemsetAsync(d_status, 0, sizeof(int), P.copy));
    for (int c = 0; c < nchunk; ++c) {
        const int b0 = c * chunk, nb = (b0 + chunk <= B) ? chunk : B - b0;
        const float* dev_heads[B200_MAX_SCALES] = {nullptr, nullptr, nullptr, nullptr};
        for (int s = 0; s < layout->num_scales; ++s) {
            unsigned char* dst = reinterpret_cast<unsigned char*>(P.heads[s]) + per_img[s] * b0;
            const unsigned char* src = reinterpret_cast<const unsigned char*>(heads_host[s]) + per_img[s] * b0;
            B200_CUDA_TRY(cudaMemcpyAsync(dst, src, per_img[s] * nb, cudaMemcpyHostToDevice, P.copy));
            dev_heads[s] = reinterpret_cast<const float*>(dst);
        }
        B200_CUDA_TRY(cudaEventRecord(P.ev[c], P.copy));
        B200_CUDA_TRY(cudaStreamWaitEvent(P.compute, P.ev[c], 0));
        sub.batch = nb;
        void* ws = reinterpret_cast<unsigned char*>(P.ws) + ws_one * (c & 1);
        const int rc = b200_yolo_postprocess(&sub, dev_heads, d_idf, conf_thr, nms_thr, nms_mode, capacity,
                                             max_det, d_det + (size_t)b0 * max_det * 6,
                                             d_keep + (size_t)b0 * max_det, nullptr, d_cnt + b0, nullptr,
                                             d_status, ws, ws_one, P.compute);
        if (rc != B200_OK) return rc;
    }
    B200_CUDA_TRY(cudaMemcpyAsync(det_host, d_det, det_b, cudaMemcpyDeviceToHost, P.compute));
    if (det_keep_host) B200_CUDA_TRY(cudaMemcpyAsync(det_keep_host, d_keep, keep_b, cudaMemcpyDeviceToHost, P.compute));
    B200_CUDA_TRY(cudaMemcpyAsync(det_count_host, d_cnt, sizeof(int) * B, cudaMemcpyDeviceToHost, P.compute));
    if (status_host) B200_CUDA_TRY(cudaMemcpyAsync(status_host, d_status, sizeof(int), cudaMemcpyDeviceToHost, P.compute));
    B200_CUDA_TRY(cudaStreamSynchronize(P.compute));
    return B200_OK;
}

// ------------------------------------------------------------------------------------------- NMS
size_t b200_nms_workspace_bytes(int64_t total_boxes, int32_t num_segments, int32_t max_segment) {
    if (total_boxes < 0 || num_segments < 0 || max_segment < 0) return 0;
    const size_t ms = max_segment > 0 ? (size_t)max_segment : (size_t)total_boxes;
    return nms_scratch_bytes((size_t)total_boxes, (size_t)num_segments, ms) + 256;
}

int b200_nms(const float* boxes, const float* scores, const int32_t* labels, const int32_t* seg_offsets,
             int32_t num_segments, int64_t total_boxes, int32_t max_segment, double iou_thr, int32_t mode,
             int64_t* keep, int32_t* keep_count, int32_t* labels_out, void* workspace, size_t workspace_bytes,
             void* stream) {
    if (num_segments < 0 || total_boxes < 0 || max_segment < 0) return B200_ERR_INVALID;
    if (num_segments == 0) return B200_OK;
    if (!seg_offsets || !keep_count) return B200_ERR_INVALID;
    if (total_boxes > 0 && (!boxes || !scores || !keep || !aligned16(boxes))) return B200_ERR_INVALID;
    if (mode < B200_NMS_MAJORITY || mode > B200_NMS_TV_AUTO) return B200_ERR_INVALID;
    if (mode != B200_NMS_TV && !labels && total_boxes > 0) return B200_ERR_INVALID;
    const size_t ms = max_segment > 0 ? (size_t)max_segment : (size_t)total_boxes;
    NmsParams np{};
    if (!nms_carve_scratch(&np, (size_t)total_boxes, (size_t)num_segments, ms, workspace, workspace_bytes))
        return B200_ERR_WORKSPACE;
    np.boxes = boxes; np.scores = scores; np.labels = labels; np.seg_offsets = seg_offsets;
    np.keep = reinterpret_cast<long long*>(keep); np.keep_count = keep_count; np.labels_out = labels_out;
    np.mode = mode;
    np.thr_f = (float)iou_thr;
    np.thr_d = iou_thr;
    np.from_slab = 0;
    np.max_seg = (int)ms;
    return launch_nms(np, num_segments, static_cast<cudaStream_t>(stream));
}

// ------------------------------------------------------------------------------------------- IoU
int b200_box_iou(const float* boxes1, int32_t m, const float* boxes2, int32_t n, int32_t kind,
                 int32_t xcycwh, float* out, void* stream) {
    if (m < 0 || n < 0 || kind < B200_IOU || kind > B200_IOU_TV) return B200_ERR_INVALID;
    if (m == 0 || n == 0) return B200_OK;
    if (!boxes1 || !boxes2 || !out || !aligned16(boxes1) || !aligned16(boxes2)) return B200_ERR_INVALID;
    if (kind == B200_IOU_TV && xcycwh) return B200_ERR_INVALID;
    return launch_box_iou(boxes1, m, boxes2, n, kind, xcycwh, out, static_cast<cudaStream_t>(stream));
}

int b200_box_iou_paired(const float* boxes1, const float* boxes2, int32_t k, int32_t kind,
                        int32_t xcycwh, float* out, void* stream) {
    if (k < 0 || kind < B200_IOU || kind > B200_IOU_TV) return B200_ERR_INVALID;
    if (k == 0) return B200_OK;
    if (!boxes1 || !boxes2 || !out || !aligned16(boxes1) || !aligned16(boxes2)) return B200_ERR_INVALID;
    return launch_box_iou_pair(boxes1, boxes2, k, kind, xcycwh, out, static_cast<cudaStream_t>(stream));
}

int b200_box_iou_paired_backward(const float* boxes1, const float* boxes2, const float* grad_out, int32_t k,
                                 int32_t kind, int32_t xcycwh, float* grad_boxes1, float* grad_boxes2, void* stream) {
    if (k < 0 || kind < B200_IOU || kind > B200_CIOU) return B200_ERR_INVALID;
    if (k == 0) return B200_OK;
    if (!boxes1 || !boxes2 || !grad_out || (!grad_boxes1 && !grad_boxes2) || !aligned16(boxes1) || !aligned16(boxes2) ||
        (grad_boxes1 && !aligned16(grad_boxes1)) || (grad_boxes2 && !aligned16(grad_boxes2)))
        return B200_ERR_INVALID;
    return launch_box_iou_pair_bwd(boxes1, boxes2, grad_out, k, kind, xcycwh, grad_boxes1, grad_boxes2,
                                   static_cast<cudaStream_t>(stream));
}

size_t b200_iou_match_workspace_bytes(int32_t batch, int32_t max_gt) {
    if (batch < 1 || max_gt < 1) return 0;
    return align_up(sizeof(unsigned long long) * (size_t)batch * max_gt, 256);
}

int b200_iou_match(const float* gt, const int32_t* gt_count, int32_t batch, int32_t max_gt,
                   const float* anchors, int32_t n, int32_t kind, float ignore_thr, int64_t* best_anchor,
                   uint8_t* noobj, void* workspace, size_t workspace_bytes, void* stream) {
    if (!gt || !gt_count || !anchors || !best_anchor || !noobj || batch < 1 || max_gt < 1 || n < 1)
        return B200_ERR_INVALID;
    if (kind < B200_IOU || kind > B200_CIOU || !aligned16(gt) || !aligned16(anchors)) return B200_ERR_INVALID;
    if (!workspace || (reinterpret_cast<uintptr_t>(workspace) & 255u) ||
        workspace_bytes < b200_iou_match_workspace_bytes(batch, max_gt))
        return B200_ERR_WORKSPACE;
    return launch_iou_match(gt, gt_count, batch, max_gt, anchors, n, kind, ignore_thr,
                            reinterpret_cast<long long*>(best_anchor), noobj,
                            reinterpret_cast<unsigned long long*>(workspace), static_cast<cudaStream_t>(stream));
}

// ------------------------------------------------------------------------------------------- RPN
size_t b200_rpn_workspace_bytes(int32_t batch, int32_t total_anchors, int32_t num_levels,
                                int32_t pre_nms_top_n) {
    if (batch < 1 || total_anchors < 1 || num_levels < 1 || pre_nms_top_n < 1) return 0;
    return rpn_workspace_bytes(batch, total_anchors, num_levels, pre_nms_top_n);
}

int b200_rpn_filter(const float* objectness, const float* deltas, const float* anchors, int32_t batch,
                    int32_t total_anchors, const int32_t* level_sizes_host, int32_t num_levels,
                    const float* image_hw, int32_t pre_nms_top_n, int32_t post_nms_top_n, double nms_thr,
                    float score_thr, float min_size, int32_t nms_mode, float* out_boxes, float* out_scores,
                    int32_t* out_index, int32_t* out_count, void* workspace, size_t workspace_bytes,
                    void* stream) {
    if (!objectness || !deltas || !anchors || !level_sizes_host || !image_hw || !out_boxes || !out_scores ||
        !out_count || batch < 1 || total_anchors < 1 || num_levels < 1 || num_levels > 16 ||
        pre_nms_top_n < 1 || post_nms_top_n < 1)
        return B200_ERR_INVALID;
    if (nms_mode != B200_NMS_TV_CLASS && nms_mode != B200_NMS_TV_TRICK) return B200_ERR_INVALID;
    if (!aligned16(deltas) || !aligned16(anchors) || !aligned16(out_boxes)) return B200_ERR_INVALID;
    if (!workspace || (reinterpret_cast<uintptr_t>(workspace) & 255u) ||
        workspace_bytes < b200_rpn_workspace_bytes(batch, total_anchors, num_levels, pre_nms_top_n))
        return B200_ERR_WORKSPACE;
    return launch_rpn_filter(objectness, deltas, anchors, nullptr, batch, total_anchors, level_sizes_host, num_levels,
                             image_hw, pre_nms_top_n, post_nms_top_n, nms_thr, score_thr, min_size, nms_mode,
                             out_boxes, out_scores, out_index, out_count, workspace, workspace_bytes,
                             static_cast<cudaStream_t>(stream));
}

size_t b200_rpn_top_n_idx_workspace_bytes(int32_t batch, int32_t total_anchors, int32_t num_levels, int32_t pre_nms_top_n) {
    if (batch < 1 || total_anchors < 1 || num_levels < 1 || pre_nms_top_n < 1) return 0;
    return rpn_topk_workspace_bytes(batch, total_anchors, num_levels, pre_nms_top_n);
}

int b200_rpn_top_n_idx(const float* objectness, int32_t batch, int32_t total_anchors, const int32_t* level_sizes_host,
                       int32_t num_levels, int32_t pre_nms_top_n, int64_t* out_index, void* workspace, size_t workspace_bytes,
                       void* stream) {
    if (!objectness || !level_sizes_host || !out_index || batch < 1 || total_anchors < 1 || num_levels < 1 ||
        num_levels > 16 || pre_nms_top_n < 1)
        return B200_ERR_INVALID;
    if (!workspace || (reinterpret_cast<uintptr_t>(workspace) & 255u)) return B200_ERR_WORKSPACE;
    return launch_rpn_topk(objectness, batch, total_anchors, level_sizes_host, num_levels, pre_nms_top_n,
                           reinterpret_cast<long long*>(out_index), workspace, workspace_bytes,
                           static_cast<cudaStream_t>(stream));
}

int b200_rpn_filter_proposals(const float* objectness, const float* proposals, int32_t batch,
                              int32_t total_anchors, const int32_t* level_sizes_host, int32_t num_levels,
                              const float* image_hw, int32_t pre_nms_top_n, int32_t post_nms_top_n, double nms_thr,
                              float score_thr, float min_size, int32_t nms_mode, float* out_boxes,
                              float* out_scores, int32_t* out_index, int32_t* out_count, void* workspace,
                              size_t workspace_bytes, void* stream) {
    if (!objectness || !proposals || !level_sizes_host || !image_hw || !out_boxes || !out_scores || !out_count ||
        batch < 1 || total_anchors < 1 || num_levels < 1 || num_levels > 16 || pre_nms_top_n < 1 ||
        post_nms_top_n < 1)
        return B200_ERR_INVALID;
    if (nms_mode != B200_NMS_TV_CLASS && nms_mode != B200_NMS_TV_TRICK) return B200_ERR_INVALID;
    if (!aligned16(proposals) || !aligned16(out_boxes)) return B200_ERR_INVALID;
    if (!workspace || (reinterpret_cast<uintptr_t>(workspace) & 255u) ||
        workspace_bytes < b200_rpn_workspace_bytes(batch, total_anchors, num_levels, pre_nms_top_n))
        return B200_ERR_WORKSPACE;
    return launch_rpn_filter(objectness, nullptr, nullptr, proposals, batch, total_anchors, level_sizes_host,
                             num_levels, image_hw, pre_nms_top_n, post_nms_top_n, nms_thr, score_thr, min_size,
                             nms_mode, out_boxes, out_scores, out_index, out_count, workspace, workspace_bytes,
                             static_cast<cudaStream_t>(stream));
}

// RetinaNet.postprocess_detections (retinanet.py:414-472)
size_t b200_retinanet_workspace_bytes(int32_t batch, int32_t total_anchors, int32_t num_classes, int32_t num_levels,
                                      int32_t topk_candidates) {
    if (batch < 1 || total_anchors < 1 || num_classes < 1 || num_levels < 1 || topk_candidates < 1) return 0;
    return rpn_workspace_bytes(batch, total_anchors * num_classes, num_levels, topk_candidates);
}

int b200_retinanet_postprocess(const float* cls_logits, const float* bbox_regression, const float* anchors, int32_t batch,
                               int32_t total_anchors, int32_t num_classes, const int32_t* level_anchors_host,
                               int32_t num_levels, const float* tfidf, const float* image_hw, int32_t topk_candidates,
                               float score_thr, double nms_thr, int32_t nms_mode, int32_t detections_per_img,
                               float* out_boxes, float* out_scores, int32_t* out_labels, int32_t* out_count,
                               void* workspace, size_t workspace_bytes, void* stream) {
    if (!cls_logits || !bbox_regression || !anchors || !level_anchors_host || !image_hw || !out_boxes || !out_scores ||
        !out_labels || !out_count || batch < 1 || total_anchors < 1 || num_classes < 1 || num_levels < 1 || num_levels > 16 ||
        topk_candidates < 1 || detections_per_img < 1)
        return B200_ERR_INVALID;
    if (nms_mode != B200_NMS_TV_CLASS && nms_mode != B200_NMS_TV_TRICK && nms_mode != B200_NMS_TV_AUTO) return B200_ERR_INVALID;
    if (!aligned16(bbox_regression) || !aligned16(anchors) || !aligned16(out_boxes)) return B200_ERR_INVALID;
    if ((long long)total_anchors * num_classes > 0x7fffffffll) return B200_ERR_INVALID;
    if (!workspace || (reinterpret_cast<uintptr_t>(workspace) & 255u) ||
        workspace_bytes < b200_retinanet_workspace_bytes(batch, total_anchors, num_classes, num_levels, topk_candidates))
        return B200_ERR_WORKSPACE;
    int32_t level_sizes[16];
    for (int l = 0; l < num_levels; ++l) level_sizes[l] = level_anchors_host[l] * num_classes;
    // no small-box filter in RetinaNet: min_size = -inf
    return launch_rpn_filter_ex(cls_logits, bbox_regression, anchors, nullptr, batch, total_anchors * num_classes, level_sizes,
                                num_levels, image_hw, topk_candidates, detections_per_img, nms_thr, score_thr, -INFINITY,
                                nms_mode, out_boxes, out_scores, nullptr, out_count, workspace, workspace_bytes,
                                static_cast<cudaStream_t>(stream), num_classes, tfidf, out_labels);
}

// ------------------------------------------------------------------------- element-wise drop-ins
int b200_abs_coord(const float* boxes, int64_t n, float* out, void* stream) {
    if (n < 0) return B200_ERR_INVALID;
    if (n == 0) return B200_OK;
    if (!boxes || !out || !aligned16(boxes) || !aligned16(out)) return B200_ERR_INVALID;
    return launch_abs_coord(boxes, n, out, static_cast<cudaStream_t>(stream));
}

int b200_boxcoder_decode(const float* rel_codes, const float* boxes, int64_t n, int32_t boxes_per_row,
                         const float* weights_host, float xform_clip, float* out, void* stream) {
    if (n < 0 || boxes_per_row < 1 || !weights_host) return B200_ERR_INVALID;
    if (n == 0) return B200_OK;
    if (!rel_codes || !boxes || !out || !aligned16(rel_codes) || !aligned16(boxes) || !aligned16(out))
        return B200_ERR_INVALID;
    return launch_boxcoder(rel_codes, boxes, n, boxes_per_row, weights_host, xform_clip, out,
                           static_cast<cudaStream_t>(stream));
}

int b200_boxcoder_encode(const float* reference_boxes, const float* proposals, int64_t n, const float* weights_host,
                         float* out, void* stream) {
    if (n < 0 || !weights_host) return B200_ERR_INVALID;
    if (n == 0) return B200_OK;
    if (!reference_boxes || !proposals || !out || !aligned16(reference_boxes) || !aligned16(proposals) || !aligned16(out))
        return B200_ERR_INVALID;
    return launch_boxcoder_encode(reference_boxes, proposals, n, weights_host, out, static_cast<cudaStream_t>(stream));
}

int b200_matcher(const float* quality, int32_t m, int32_t n, float high_thr, float low_thr,
                 int32_t allow_low_quality, int64_t* matches, void* workspace, size_t workspace_bytes,
                 void* stream) {
    if (m < 1 || n < 1 || !quality || !matches) return B200_ERR_INVALID;
    if (allow_low_quality && (!workspace || workspace_bytes < sizeof(long long) * (size_t)n ||
                              (reinterpret_cast<uintptr_t>(workspace) & 7u)))
        return B200_ERR_WORKSPACE;
    return launch_matcher(quality, m, n, high_thr, low_thr, allow_low_quality, reinterpret_cast<long long*>(matches),
                          reinterpret_cast<long long*>(workspace), static_cast<cudaStream_t>(stream));
}

}  // extern "C"
