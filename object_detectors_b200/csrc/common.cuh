// common.cuh -- shared device helpers for libb200det (sm_100a).
//
// Arithmetic policy: every box / IoU expression that feeds a
// threshold decision is written with explicit round-to-nearest intrinsics so that it is evaluated
// exactly like the reference's unfused fp32 tensor ops (one rounding per operation, no FMA
// contraction) regardless of compiler flags.  The library is also compiled with --fmad=false.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/b200det.h"

#define B200_CUDA_TRY(expr)                          \
    do {                                             \
        cudaError_t _e = (expr);                     \
        if (_e != cudaSuccess) return B200_ERR_CUDA; \
    } while (0)

namespace b200 {

static constexpr unsigned kFullMask = 0xffffffffu;

__host__ __device__ inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// Opt a kernel in to `bytes` of dynamic shared memory on the CURRENT device.  Function attributes live in the
// device's context, so the "already raised" cache is per (call site, device): a process that drives several GPUs
// (the reference spawns one process per GPU, but nothing forbids it) gets the attribute set on each of them.
struct SmemOptIn {
    size_t have[32] = {};
    template <typename Kernel>
    cudaError_t ensure(Kernel kernel, size_t bytes) {
        if (bytes <= 48 * 1024) return cudaSuccess;
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        const int slot = dev & 31;
        if (bytes <= have[slot]) return cudaSuccess;
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e == cudaSuccess) have[slot] = bytes;
        return e;
    }
};
// number of SMs of the current device (148 if the query fails)
inline int current_sm_count() {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0)
        return v;
    return 148;
}
__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---------------------------------------------------------------- streaming loads (read-once)
// ld.global.nc + L1::no_allocate: the head tensors are touched exactly once, keep them out of L1.
// `volatile` pins the load where it is written (the dense kernel must read every byte even for
// cells whose result is later discarded; without it ptxas sinks loads into the live-cell branch).
__device__ __forceinline__ float4 ldg_stream_v4(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ float ldg_stream_f32(const float* p) {
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// ---------------------------------------------------------------- order-preserving float keys
// maps fp32 to uint32 so that unsigned order == float order (-inf < ... < +inf; NaNs at the ends)
__device__ __forceinline__ uint32_t orderable(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float from_orderable(uint32_t k) {
    uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(u);
}

// ---------------------------------------------------------------- exact transcendental wrappers
// sigmoid as the reference evaluates it: 1 / (1 + exp(-x)), IEEE division, accurate expf.
__device__ __forceinline__ float sigmoid_ref(float x) {
    return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x)));
}

// ---------------------------------------------------------------- IoU flavours (appendix A.3)
struct Box { float x1, y1, x2, y2; };

__device__ __forceinline__ float box_area(const Box& b) {
    return __fmul_rn(__fsub_rn(b.x2, b.x1), __fsub_rn(b.y2, b.y1));
}

// intersection with the "clamp(min=0)" of the reference (NaN-propagation of torch.clamp is not
// reproduced: coordinates are assumed finite)
__device__ __forceinline__ float box_inter(const Box& a, const Box& b) {
    float w = __fsub_rn(fminf(a.x2, b.x2), fmaxf(a.x1, b.x1));
    float h = __fsub_rn(fminf(a.y2, b.y2), fmaxf(a.y1, b.y1));
    w = fmaxf(w, 0.0f);
    h = fmaxf(h, 0.0f);
    return __fmul_rn(w, h);
}

// helper.nms_majority (helper.py:339-366): S = picked box (area_s), T = remaining box (area_t)
__device__ __forceinline__ float iou_majority(float inter, float area_s, float area_t) {
    return __fdiv_rn(inter, __fadd_rn(__fsub_rn(area_t, inter), area_s));
}
// torchvision nms kernel: inter / (area_i + area_j - inter)
__device__ __forceinline__ float iou_tv(float inter, float area_i, float area_j) {
    return __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_i, area_j), inter));
}

// helper.get_abs_coord (helper.py:203-217)
__device__ __forceinline__ Box abs_coord(float xc, float yc, float w, float h) {
    const float hw = __fmul_rn(w, 0.5f);  // w/2 is exact either way
    const float hh = __fmul_rn(h, 0.5f);
    return Box{__fsub_rn(xc, hw), __fsub_rn(yc, hh), __fadd_rn(xc, hw), __fadd_rn(yc, hh)};
}

// helper.bbox_iou on corner boxes (helper.py:244-277); kind: B200_IOU / GIOU / DIOU / CIOU / IOU_TV
__device__ __forceinline__ float pair_iou(const Box& p, const Box& q, int kind) {
    float iw = fmaxf(__fsub_rn(fminf(p.x2, q.x2), fmaxf(p.x1, q.x1)), 0.0f);
    float ih = fmaxf(__fsub_rn(fminf(p.y2, q.y2), fmaxf(p.y1, q.y1)), 0.0f);
    const float inter = __fmul_rn(iw, ih);
    const float w1 = __fsub_rn(p.x2, p.x1), h1 = __fsub_rn(p.y2, p.y1);
    const float w2 = __fsub_rn(q.x2, q.x1), h2 = __fsub_rn(q.y2, q.y1);
    if (kind == B200_IOU_TV) {
        return __fdiv_rn(inter, __fsub_rn(__fadd_rn(__fmul_rn(w1, h1), __fmul_rn(w2, h2)), inter));
    }
    // union = (w1*h1 + 1e-16) + w2*h2 - inter      (:255)
    const float uni = __fsub_rn(__fadd_rn(__fadd_rn(__fmul_rn(w1, h1), 1e-16f), __fmul_rn(w2, h2)), inter);
    const float iou = __fdiv_rn(inter, uni);
    if (kind == B200_IOU) return iou;
    const float cw = __fsub_rn(fmaxf(p.x2, q.x2), fminf(p.x1, q.x1));
    const float ch = __fsub_rn(fmaxf(p.y2, q.y2), fminf(p.y1, q.y1));
    if (kind == B200_GIOU) {
        const float c_area = __fadd_rn(__fmul_rn(cw, ch), 1e-16f);
        return __fsub_rn(iou, __fdiv_rn(__fsub_rn(c_area, uni), c_area));
    }
    // c2 = cw**2 + ch**2 + 1e-16 ; rho2 = ((b2x1+b2x2)-(b1x1+b1x2))**2/4 + (...)**2/4   (:266-268)
    const float c2 = __fadd_rn(__fadd_rn(__fmul_rn(cw, cw), __fmul_rn(ch, ch)), 1e-16f);
    const float dx = __fsub_rn(__fadd_rn(q.x1, q.x2), __fadd_rn(p.x1, p.x2));
    const float dy = __fsub_rn(__fadd_rn(q.y1, q.y2), __fadd_rn(p.y1, p.y2));
    const float rho2 = __fadd_rn(__fmul_rn(__fmul_rn(dx, dx), 0.25f), __fmul_rn(__fmul_rn(dy, dy), 0.25f));
    if (kind == B200_DIOU) return __fsub_rn(iou, __fdiv_rn(rho2, c2));
    // CIoU: v = (4/pi^2) * (atan(w2/h2) - atan(w1/h1))^2 ; alpha = v / (1 - iou + v)   (:271-275)
    const float da = __fsub_rn(atanf(__fdiv_rn(w2, h2)), atanf(__fdiv_rn(w1, h1)));
    const float v = __fmul_rn(0.40528473456935116f, __fmul_rn(da, da));
    const float alpha = __fdiv_rn(v, __fadd_rn(__fsub_rn(1.0f, iou), v));
    return __fsub_rn(iou, __fadd_rn(__fdiv_rn(rho2, c2), __fmul_rn(v, alpha)));
}

}  // namespace b200
