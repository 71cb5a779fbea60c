// exchange.cu -- the path's one exchange step: all-gather of the variable-length kept-detection lists.
//
// Replaces the reference's per-rank pickle files + barrier (yolo/procedures/eval_results.py:12-31,
// yolo/main.py:102-105) and its size-exchange / pad / gather `utils.all_gather`
// (torchvision_models/detection/utils.py:75-115).
//
// One-sided design (NVLink 5 / NVSwitch peer memory, no collective kernel, no rendezvous on the data path):
//   * every rank owns ONE device buffer  data[slots][world][message] + flags[slots][world] + acks[world], exported
//     with cudaIpcGetMemHandle and mapped by every peer;
//   * b200_exchange_push   (one CTA per image x peer) stores this rank's kept lists of the current step straight into
//     every peer's data[step % slots][rank] -- only the count and the rows that exist travel (~3 KB per image instead
//     of the 6 KB fixed-capacity record) -- and, once a peer's copy is complete (system-scope fence, last CTA),
//     releases flags[step % slots][rank] = step + 1 in that peer's memory;
//   * b200_exchange_wait   (one warp) completes on the stream when the flags of all ranks for the next un-waited step
//     have arrived, and tells every peer (acks[rank] = step) that the slots of all EARLIER steps may be reused: what
//     the rank enqueued on the wait stream before this wait -- its reads of the previous step -- has completed.
// The step numbers live on the device (the kernels advance them), so both calls can be captured in CUDA graphs.
// A rank never waits for a peer except (i) in b200_exchange_wait, which is the semantics of a gather, and (ii) in
// b200_exchange_push when a peer is more than `slots` steps behind (flow control).
//
// b200_allgather_dets is the NCCL form of the same exchange (ncclAllGather bound at run time from the NCCL library
// the process has already loaded), kept as the fallback when peer mapping is not available.
#include <dlfcn.h>

#include <cstring>
#include <new>

#include "common.cuh"

namespace b200 {

struct ExchangeDev {
    float* data[B200_MAX_RANKS];                   // peer p's data region (own pointer at p == rank)
    unsigned long long* flags[B200_MAX_RANKS];     // peer p's flags[slots][world]
    unsigned long long* acks[B200_MAX_RANKS];      // peer p's acks[world]
    unsigned long long* push_step;                 // local: steps pushed so far
    unsigned long long* wait_step;                 // local: steps waited for so far
    int* arrive;                                   // local: [world + 1] CTA arrival counters of the running push
    int rank, world, batch, max_det, slots;
    long long msg;                                 // floats per message = batch * (1 + 6 * max_det)
};

__device__ __forceinline__ unsigned long long ld_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// grid (batch, world): CTA (b, p) copies image b's kept list into peer p's slot
__global__ void __launch_bounds__(128)
k_exchange_push(const __grid_constant__ ExchangeDev X, const float* __restrict__ det, const int* __restrict__ cnt) {
    const int b = blockIdx.x, p = blockIdx.y, tid = threadIdx.x;
    const unsigned long long step = *X.push_step;                  // advanced only by the last CTA of this launch
    const int slot = (int)(step % (unsigned long long)X.slots);
    if (tid == 0) {
        // flow control: the peer must have consumed the message that lived in this slot `slots` steps ago
        while (ld_sys(X.acks[X.rank] + p) + (unsigned long long)X.slots < step + 1ull) __nanosleep(200);
    }
    __syncthreads();
    const int stride = 1 + 6 * X.max_det;
    const int n = min(cnt[b], X.max_det);
    float* dst = X.data[p] + ((size_t)slot * X.world + X.rank) * (size_t)X.msg + (size_t)b * stride;
    const float* src = det + (size_t)b * X.max_det * 6;
    if (tid == 0) dst[0] = __int_as_float(n);
    for (int i = tid; i < n * 6; i += blockDim.x) dst[1 + i] = src[i];
    __threadfence_system();                                        // this thread's stores are visible at the peer
    __syncthreads();
    if (tid == 0) {
        const int c = atomicAdd(X.arrive + p, 1);
        bool s_last = false;
        if (c == X.batch - 1) {
            X.arrive[p] = 0;
            __threadfence_system();
            st_sys(X.flags[p] + (size_t)slot * X.world + X.rank, step + 1ull);       // release: message complete
            if (atomicAdd(X.arrive + X.world, 1) == X.world - 1) s_last = true;
        }
        if (s_last) {
            X.arrive[X.world] = 0;
            __threadfence();
            *X.push_step = step + 1ull;
        }
    }
}

// one warp: lane p waits for rank p's message of the next un-waited step, then acknowledges to rank p
__global__ void __launch_bounds__(32)
k_exchange_wait(const __grid_constant__ ExchangeDev X) {
    const int p = threadIdx.x;
    const unsigned long long step = *X.wait_step;
    const int slot = (int)(step % (unsigned long long)X.slots);
    if (p < X.world) {
        const unsigned long long* f = X.flags[X.rank] + (size_t)slot * X.world + p;
        while (ld_sys(f) != step + 1ull) __nanosleep(200);
    }
    __threadfence_system();                                        // acquire: the messages behind the flags
    __syncwarp();
    // Everything this rank enqueued on this stream before this launch has finished, in particular its reads of the
    // PREVIOUS step's messages: steps < `step` are consumed, rank p may overwrite their slots.  (The message of
    // `step` itself stays valid until the next wait runs.)
    if (p < X.world) st_sys(X.acks[p] + X.rank, step);
    if (p == 0) *X.wait_step = step + 1ull;
}

static __global__ void k_pack(const float* __restrict__ det, const int* __restrict__ cnt, int max_det,
                              float* __restrict__ msg) {
    const int b = blockIdx.x;
    const int stride = 1 + max_det * 6;
    const int n = min(cnt[b], max_det);
    float* dst = msg + (size_t)b * stride;
    if (threadIdx.x == 0) dst[0] = __int_as_float(n);
    for (int i = threadIdx.x; i < max_det * 6; i += blockDim.x)
        dst[1 + i] = i < n * 6 ? det[(size_t)b * max_det * 6 + i] : 0.f;
}

}  // namespace b200

using namespace b200;

struct b200_exchange {
    ExchangeDev dev;
    void* local = nullptr;                  // the cudaMalloc'ed block this rank exports
    void* peer_base[B200_MAX_RANKS] = {};   // mapped blocks of the peers (nullptr for self / not connected)
    size_t data_bytes = 0, flag_bytes = 0, ack_bytes = 0, total_bytes = 0;
    int connected = 0;
};

namespace {
void carve_peer(b200_exchange* x, int p, void* base) {
    unsigned char* q = reinterpret_cast<unsigned char*>(base);
    x->dev.data[p] = reinterpret_cast<float*>(q);
    x->dev.flags[p] = reinterpret_cast<unsigned long long*>(q + x->data_bytes);
    x->dev.acks[p] = reinterpret_cast<unsigned long long*>(q + x->data_bytes + x->flag_bytes);
}
}  // namespace

extern "C" {

int b200_exchange_create(int32_t rank, int32_t world, int32_t batch, int32_t max_det, int32_t slots,
                         b200_exchange** out) {
    if (!out || world < 1 || world > B200_MAX_RANKS || rank < 0 || rank >= world || batch < 1 || max_det < 1 || slots < 2)
        return B200_ERR_INVALID;
    b200_exchange* x = new (std::nothrow) b200_exchange();
    if (!x) return B200_ERR_INVALID;
    ExchangeDev& d = x->dev;
    d.rank = rank; d.world = world; d.batch = batch; d.max_det = max_det; d.slots = slots;
    d.msg = (long long)batch * (1 + 6ll * max_det);
    x->data_bytes = align_up(sizeof(float) * (size_t)slots * world * (size_t)d.msg, 256);
    x->flag_bytes = align_up(sizeof(unsigned long long) * (size_t)slots * world, 256);
    x->ack_bytes = align_up(sizeof(unsigned long long) * (size_t)world, 256);
    const size_t local_bytes = 256 + 256 + align_up(sizeof(int) * (size_t)(world + 1), 256);
    x->total_bytes = x->data_bytes + x->flag_bytes + x->ack_bytes + local_bytes;
    if (cudaMalloc(&x->local, x->total_bytes) != cudaSuccess) { delete x; return B200_ERR_CUDA; }
    if (cudaMemset(x->local, 0, x->total_bytes) != cudaSuccess) { cudaFree(x->local); delete x; return B200_ERR_CUDA; }
    carve_peer(x, rank, x->local);
    unsigned char* q = reinterpret_cast<unsigned char*>(x->local) + x->data_bytes + x->flag_bytes + x->ack_bytes;
    d.push_step = reinterpret_cast<unsigned long long*>(q);
    d.wait_step = reinterpret_cast<unsigned long long*>(q + 256);
    d.arrive = reinterpret_cast<int*>(q + 512);
    x->connected = 1;
    *out = x;
    return B200_OK;
}

int b200_exchange_handle(b200_exchange* x, void* handle64) {
    if (!x || !handle64) return B200_ERR_INVALID;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handles are 64 bytes");
    cudaIpcMemHandle_t h;
    B200_CUDA_TRY(cudaIpcGetMemHandle(&h, x->local));
    memcpy(handle64, &h, sizeof(h));
    return B200_OK;
}

int b200_exchange_connect(b200_exchange* x, int32_t peer, const void* handle64) {
    if (!x || !handle64 || peer < 0 || peer >= x->dev.world || peer == x->dev.rank || x->peer_base[peer])
        return B200_ERR_INVALID;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    void* base = nullptr;
    B200_CUDA_TRY(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
    x->peer_base[peer] = base;
    carve_peer(x, peer, base);
    ++x->connected;
    return B200_OK;
}

int b200_exchange_push(b200_exchange* x, const float* det, const int32_t* det_count, void* stream) {
    if (!x || !det || !det_count || x->connected != x->dev.world) return B200_ERR_INVALID;
    const dim3 grid((unsigned)x->dev.batch, (unsigned)x->dev.world);
    k_exchange_push<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(x->dev, det, det_count);
    return cudaGetLastError() == cudaSuccess ? B200_OK : B200_ERR_CUDA;
}

int b200_exchange_wait(b200_exchange* x, void* stream) {
    if (!x || x->connected != x->dev.world) return B200_ERR_INVALID;
    k_exchange_wait<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(x->dev);
    return cudaGetLastError() == cudaSuccess ? B200_OK : B200_ERR_CUDA;
}

const float* b200_exchange_message(b200_exchange* x, int64_t step, int32_t src_rank) {
    if (!x || step < 0 || src_rank < 0 || src_rank >= x->dev.world) return nullptr;
    const size_t slot = (size_t)(step % x->dev.slots);
    return x->dev.data[x->dev.rank] + (slot * x->dev.world + src_rank) * (size_t)x->dev.msg;
}

int b200_exchange_read(b200_exchange* x, int64_t step, float* gathered, void* stream) {
    if (!x || !gathered || step < 0) return B200_ERR_INVALID;
    const float* src = b200_exchange_message(x, step, 0);
    B200_CUDA_TRY(cudaMemcpyAsync(gathered, src, sizeof(float) * (size_t)x->dev.world * (size_t)x->dev.msg,
                                  cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)));
    return B200_OK;
}

int b200_exchange_steps(b200_exchange* x, int64_t* pushed, int64_t* waited) {
    if (!x) return B200_ERR_INVALID;
    unsigned long long v[2] = {0, 0};
    B200_CUDA_TRY(cudaMemcpy(&v[0], x->dev.push_step, sizeof(v[0]), cudaMemcpyDeviceToHost));
    B200_CUDA_TRY(cudaMemcpy(&v[1], x->dev.wait_step, sizeof(v[1]), cudaMemcpyDeviceToHost));
    if (pushed) *pushed = (int64_t)v[0];
    if (waited) *waited = (int64_t)v[1];
    return B200_OK;
}

int b200_exchange_destroy(b200_exchange* x) {
    if (!x) return B200_OK;
    for (int p = 0; p < B200_MAX_RANKS; ++p)
        if (x->peer_base[p]) cudaIpcCloseMemHandle(x->peer_base[p]);
    if (x->local) cudaFree(x->local);
    delete x;
    return B200_OK;
}

// ------------------------------------------------------------------------------------------ pack + NCCL form
int b200_pack_detections(const float* det, const int32_t* det_count, int32_t batch, int32_t max_det,
                         float* message, void* stream) {
    if (!det || !det_count || !message || batch < 1 || max_det < 1) return B200_ERR_INVALID;
    b200::k_pack<<<batch, 256, 0, static_cast<cudaStream_t>(stream)>>>(det, det_count, max_det, message);
    return cudaGetLastError() == cudaSuccess ? B200_OK : B200_ERR_CUDA;
}

int b200_allgather_dets(const float* det, const int32_t* det_count, int32_t batch, int32_t max_det,
                        float* message, float* gathered, void* nccl_comm, void* stream) {
    if (!message || !gathered || !nccl_comm) return B200_ERR_INVALID;
    typedef int (*AllGatherFn)(const void*, void*, size_t, int, void*, cudaStream_t);
    static AllGatherFn fn = nullptr;
    if (!fn) {
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);       // the copy the process already uses (torch's)
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW);
        if (h) fn = reinterpret_cast<AllGatherFn>(dlsym(h, "ncclAllGather"));
        if (!fn) return B200_ERR_INVALID;
    }
    const int rc = b200_pack_detections(det, det_count, batch, max_det, message, stream);
    if (rc != B200_OK) return rc;
    const size_t count = (size_t)batch * (1 + 6 * (size_t)max_det);
    return fn(message, gathered, count, /*ncclFloat32*/ 7, nccl_comm, static_cast<cudaStream_t>(stream)) == 0
               ? B200_OK : B200_ERR_CUDA;
}

}  // extern "C"
