// emit.cu -- result emission of the YOLO evaluation loop as ONE kernel and one packed record array
// (replaces yolo/procedures/test_one_epoch.py:41-66 + yolo/utilities/helper.py:16-24 of the reference: per image,
// rescale the kept boxes to the original image, xyxy -> xywh, area, 80 -> 91 category ids, image id; the reference
// does it with ~10 tensor ops and five .tolist() round trips per image).
//
// Reference behaviour kept on request (strict_reference != 0): images without detections are dropped from the
// prediction list BEFORE it is matched with `targets` by position (:37,:41-47), so the k-th NON-EMPTY image is
// scaled with the size -- and labelled with the id -- of targets[k].
#include "common.cuh"

namespace b200 {

// one CTA per image; records of image b start at the sum of the counts of the images before it
__global__ void __launch_bounds__(128)
k_emit_results(const float* __restrict__ det, const int* __restrict__ cnt, int batch, int max_det,
               const float* __restrict__ img_hw, const long long* __restrict__ image_id, float inp_dim,
               const int* __restrict__ class_map, int num_map, int strict, float* __restrict__ rec,
               int* __restrict__ rec_cat, long long* __restrict__ rec_img, int* __restrict__ total) {
    __shared__ int s_off, s_owner;
    const int b = blockIdx.x, tid = threadIdx.x;
    if (tid < 32) {
        int off = 0, nonempty = 0;
        for (int i = tid; i < b; i += 32) {
            const int c = min(cnt[i], max_det);
            off += c;
            nonempty += c > 0;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            off += __shfl_xor_sync(kFullMask, off, o);
            nonempty += __shfl_xor_sync(kFullMask, nonempty, o);
        }
        if (tid == 0) { s_off = off; s_owner = strict ? nonempty : b; }
    }
    __syncthreads();
    const int n = min(cnt[b], max_det);
    if (b == batch - 1 && tid == 0) *total = s_off + n;
    if (n == 0) return;
    const int owner = s_owner;                       // < batch: at most b non-empty images precede image b
    const float sh = img_hw[2 * owner], sw = img_hw[2 * owner + 1];
    const long long id = image_id[owner];
    for (int t = tid; t < n; t += blockDim.x) {
        const float* d = det + ((size_t)b * max_det + t) * 6;
        // xmin = atrbs[:,0] / inp_dim * img_size[1]  (test_one_epoch.py:42-45), w = xmax - xmin (:46-47)
        const float x1 = __fmul_rn(__fdiv_rn(d[0], inp_dim), sw), y1 = __fmul_rn(__fdiv_rn(d[1], inp_dim), sh);
        const float x2 = __fmul_rn(__fdiv_rn(d[2], inp_dim), sw), y2 = __fmul_rn(__fdiv_rn(d[3], inp_dim), sh);
        const float w = __fsub_rn(x2, x1), h = __fsub_rn(y2, y1);
        float* r = rec + (size_t)(s_off + t) * 6;
        r[0] = x1; r[1] = y1; r[2] = w; r[3] = h;
        r[4] = __fmul_rn(w, h);                                            // areas = bboxes[:,2] * bboxes[:,3] (:59)
        r[5] = d[4];
        const int lab = (int)d[5];                                         // atrbs[:,5].long()
        // coco: helper.torch80_to_91 (helper.py:16-24); otherwise labels + 1 (:55-56)
        rec_cat[s_off + t] = class_map ? class_map[min(max(lab, 0), num_map - 1)] : lab + 1;
        rec_img[s_off + t] = id;
    }
}

// torchvision.ops.boxes.clip_boxes_to_image (tvision/boxes.py, called at rpn.py:260, roi_heads.py:746, retinanet.py:452,
// ssd.py:397): x coordinates clamped to [0, width], y coordinates to [0, height]; any leading shape, 4 floats per box
__global__ void __launch_bounds__(256)
k_clip_boxes(const float4* __restrict__ boxes, long long n, float height, float width, float4* __restrict__ out) {
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        float4 b = boxes[i];
        // torch.clamp(min, max) = min(max(x, lo), hi); NaN propagates
        b.x = b.x != b.x ? b.x : fminf(fmaxf(b.x, 0.f), width);  b.z = b.z != b.z ? b.z : fminf(fmaxf(b.z, 0.f), width);
        b.y = b.y != b.y ? b.y : fminf(fmaxf(b.y, 0.f), height); b.w = b.w != b.w ? b.w : fminf(fmaxf(b.w, 0.f), height);
        out[i] = b;
    }
}

// torchvision.ops.boxes.remove_small_boxes: indices (ascending) of the boxes with w >= min_size and h >= min_size.
// One CTA walks the boxes 1024 at a time with a running offset (an ordered compaction; the callers hold a few thousand
// boxes per image); count[0] = number of indices written.
__global__ void __launch_bounds__(1024, 1)
k_small_box_keep(const float4* __restrict__ boxes, int n, float min_size, long long* __restrict__ keep, int* __restrict__ count) {
    __shared__ int s_warp[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int running = 0;
    for (int i0 = 0; i0 < n; i0 += 1024) {
        const int i = i0 + tid;
        bool ok = false;
        if (i < n) {
            const float4 b = boxes[i];
            ok = __fsub_rn(b.z, b.x) >= min_size && __fsub_rn(b.w, b.y) >= min_size;
        }
        const unsigned bal = __ballot_sync(kFullMask, ok);
        if (lane == 0) s_warp[warp] = __popc(bal);
        __syncthreads();
        int before = 0, total = 0;
        for (int w = 0; w < 32; ++w) { const int c = s_warp[w]; if (w < warp) before += c; total += c; }
        if (ok) keep[running + before + __popc(bal & ((1u << lane) - 1u))] = i;
        running += total;
        __syncthreads();
    }
    if (tid == 0) *count = running;
}

}  // namespace b200

extern "C" int b200_clip_boxes_to_image(const float* boxes, int64_t n, float height, float width, float* out, void* stream) {
    if (n < 0 || (n > 0 && (!boxes || !out))) return B200_ERR_INVALID;
    if (n == 0) return B200_OK;
    if ((reinterpret_cast<uintptr_t>(boxes) | reinterpret_cast<uintptr_t>(out)) & 15) return B200_ERR_INVALID;
    const long long ctas = (n + 255) / 256;
    b200::k_clip_boxes<<<(unsigned)(ctas < 4096 ? ctas : 4096), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float4*>(boxes), n, height, width, reinterpret_cast<float4*>(out));
    return cudaGetLastError() == cudaSuccess ? B200_OK : B200_ERR_CUDA;
}

extern "C" int b200_remove_small_boxes(const float* boxes, int32_t n, float min_size, int64_t* keep, int32_t* count, void* stream) {
    if (n < 0 || !count || (n > 0 && (!boxes || !keep))) return B200_ERR_INVALID;
    if (n > 0 && (reinterpret_cast<uintptr_t>(boxes) & 15)) return B200_ERR_INVALID;
    b200::k_small_box_keep<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float4*>(boxes), n, min_size, reinterpret_cast<long long*>(keep), count);
    return cudaGetLastError() == cudaSuccess ? B200_OK : B200_ERR_CUDA;
}

extern "C" int b200_emit_results(const float* det, const int32_t* det_count, int32_t batch, int32_t max_det,
                                 const float* img_hw, const int64_t* image_id, float inp_dim, const int32_t* class_map,
                                 int32_t num_map, int32_t strict_reference, float* records, int32_t* category,
                                 int64_t* image, int32_t* total, void* stream) {
    if (!det || !det_count || !img_hw || !image_id || !records || !category || !image || !total || batch < 1 || max_det < 1 ||
        !(inp_dim > 0.f) || (class_map && num_map < 1))
        return B200_ERR_INVALID;
    b200::k_emit_results<<<batch, 128, 0, static_cast<cudaStream_t>(stream)>>>(
        det, det_count, batch, max_det, img_hw, reinterpret_cast<const long long*>(image_id), inp_dim, class_map, num_map,
        strict_reference, records, category, reinterpret_cast<long long*>(image), total);
    return cudaGetLastError() == cudaSuccess ? B200_OK : B200_ERR_CUDA;
}
