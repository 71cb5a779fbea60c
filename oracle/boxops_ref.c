/*
 * oracle/boxops_ref.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Scalar C restatement of the sequential loops on the detection box-ops hot path of
 * kostas1515/object_detectors, used as the CPU checker for the CUDA kernels at sizes
 * where the torch-level restatement (oracle/yolo_ref.py, oracle/tv_ref.py) is too slow.
 * Each function cites the reference lines whose arithmetic it follows.  All arithmetic
 * is IEEE fp32, one rounding per operation, no contraction: build with
 *     gcc -O2 -ffp-contract=off -fno-fast-math -shared -fPIC
 * (see oracle/Makefile).  Pinned against the Python reference / torchvision CPU ops by
 * tests/test_oracle_golden.py and tests/test_oracle_pin.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---------------------------------------------------------------------------------- */
/* stable descending order of scores, ties -> lower original index first               */
/* ---------------------------------------------------------------------------------- */
typedef struct { float s; int i; } sc_t;
static int cmp_desc(const void* a, const void* b) {
    const sc_t* x = (const sc_t*)a; const sc_t* y = (const sc_t*)b;
    if (x->s > y->s) return -1;
    if (x->s < y->s) return 1;
    /* NaN scores: keep original order among incomparable entries */
    return (x->i > y->i) - (x->i < y->i);
}
static int* order_desc(const float* scores, int n) {
    sc_t* t = (sc_t*)malloc(sizeof(sc_t) * (size_t)(n > 0 ? n : 1));
    int* o = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; ++i) { t[i].s = scores[i]; t[i].i = i; }
    qsort(t, (size_t)n, sizeof(sc_t), cmp_desc);
    for (int i = 0; i < n; ++i) o[i] = t[i].i;
    free(t);
    return o;
}

static inline float fmaxf_(float a, float b) { return a > b ? a : b; }
static inline float fminf_(float a, float b) { return a < b ? a : b; }

/* ---------------------------------------------------------------------------------- */
/* helper.nms_majority  (yolo/utilities/helper.py:280-382)                             */
/*   det6: [n,6] rows x1,y1,x2,y2,score,label(float).  Class-agnostic greedy NMS in     */
/*   descending score.  For the picked box S and every remaining box T:                 */
/*     inter = clamp(min(x2)-max(x1),0) * clamp(min(y2)-max(y1),0)        (:339-358)    */
/*     union = (area_T - inter) + area_S ; IoU = inter/union              (:361-365)    */
/*   T survives iff IoU <  thr (:368); T votes iff IoU > thr (:369); thr is compared    */
/*   in fp32.  If the voters hold >1 distinct class, S takes the most frequent one      */
/*   (ties -> smallest id, torch.unique is sorted and max returns the first) (:370-375) */
/*   Votes use the labels as they were on entry (snapshot at :301).                     */
/*   Outputs: keep_idx[k] (index into det6 rows), keep_label[k] (label after relabel).  */
/*   Returns K.                                                                         */
/* ---------------------------------------------------------------------------------- */
int ref_nms_majority(const float* det6, int n, float thr, int num_classes,
                     int32_t* keep_idx, int32_t* keep_label) {
    if (n <= 0) return 0;
    float* sc = (float*)malloc(sizeof(float) * (size_t)n);
    float* area = (float*)malloc(sizeof(float) * (size_t)n);
    int* cls = (int*)malloc(sizeof(int) * (size_t)n);
    unsigned char* gone = (unsigned char*)calloc((size_t)n, 1);
    int* hist = (int*)calloc((size_t)(num_classes > 0 ? num_classes : 1), sizeof(int));
    for (int i = 0; i < n; ++i) {
        const float* r = det6 + 6 * (size_t)i;
        sc[i] = r[4];
        area[i] = (r[2] - r[0]) * (r[3] - r[1]);                              /* :304 */
        cls[i] = (int)r[5];                                                   /* :301 */
    }
    int* ord = order_desc(sc, n);
    int K = 0;
    for (int a = 0; a < n; ++a) {
        int s = ord[a];
        if (gone[s]) continue;
        const float* S = det6 + 6 * (size_t)s;
        int label = (int)S[5];
        int voters = 0;
        for (int b = a + 1; b < n; ++b) {
            int t = ord[b];
            if (gone[t]) continue;
            const float* T = det6 + 6 * (size_t)t;
            float xx1 = fmaxf_(T[0], S[0]), yy1 = fmaxf_(T[1], S[1]);
            float xx2 = fminf_(T[2], S[2]), yy2 = fminf_(T[3], S[3]);
            float w = xx2 - xx1, h = yy2 - yy1;
            if (!(w >= 0.0f)) w = (w != w) ? w : 0.0f;   /* clamp(min=0) keeps NaN */
            if (!(h >= 0.0f)) h = (h != h) ? h : 0.0f;
            float inter = w * h;
            float uni = (area[t] - inter) + area[s];
            float iou = inter / uni;
            if (!(iou < thr)) {                     /* removed (also NaN, also == thr) */
                gone[t] = 1;
                if (iou > thr) {                    /* voter */
                    if (cls[t] >= 0 && cls[t] < num_classes) hist[cls[t]]++;
                    voters++;
                }
            }
        }
        if (voters > 0) {
            int distinct = 0, best = -1, bestc = 0;
            for (int c = 0; c < num_classes; ++c) {
                if (hist[c] > 0) {
                    distinct++;
                    if (hist[c] > bestc) { bestc = hist[c]; best = c; }
                    hist[c] = 0;
                }
            }
            if (distinct > 1 && best != label) label = best;
        }
        keep_idx[K] = s;
        keep_label[K] = label;
        ++K;
    }
    free(sc); free(area); free(cls); free(gone); free(hist); free(ord);
    return K;
}

/* ---------------------------------------------------------------------------------- */
/* torchvision.ops.nms  (third-party, torchvision 0.26.0 csrc/ops/cpu/nms_kernel.cpp,   */
/* published algorithm; call sites yolo/benchmark.py:100, telemetry.py:207,248):        */
/* stable descending sort; box i suppresses later box j iff                             */
/*   inter / (area_i + area_j - inter) > thr   with thr a double, ovr an fp32 value.    */
/* labels != NULL -> class-aware ("vanilla" batched_nms semantics, boxes.py             */
/* _batched_nms_vanilla): only equal labels interact; output is in descending score.    */
/* Returns K, keep[k] = original index.                                                 */
/* ---------------------------------------------------------------------------------- */
int ref_nms_tv(const float* boxes, const float* scores, const int64_t* labels, int n,
               double thr, int64_t* keep) {
    if (n <= 0) return 0;
    float* area = (float*)malloc(sizeof(float) * (size_t)n);
    unsigned char* gone = (unsigned char*)calloc((size_t)n, 1);
    for (int i = 0; i < n; ++i) {
        const float* r = boxes + 4 * (size_t)i;
        area[i] = (r[2] - r[0]) * (r[3] - r[1]);
    }
    int* ord = order_desc(scores, n);
    int K = 0;
    for (int a = 0; a < n; ++a) {
        int i = ord[a];
        if (gone[i]) continue;
        keep[K++] = i;
        const float* I = boxes + 4 * (size_t)i;
        for (int b = a + 1; b < n; ++b) {
            int j = ord[b];
            if (gone[j]) continue;
            if (labels && labels[i] != labels[j]) continue;
            const float* J = boxes + 4 * (size_t)j;
            float xx1 = fmaxf_(I[0], J[0]), yy1 = fmaxf_(I[1], J[1]);
            float xx2 = fminf_(I[2], J[2]), yy2 = fminf_(I[3], J[3]);
            float w = fmaxf_(0.0f, xx2 - xx1), h = fmaxf_(0.0f, yy2 - yy1);
            float inter = w * h;
            float ovr = inter / (area[i] + area[j] - inter);
            if ((double)ovr > thr) gone[j] = 1;
        }
    }
    free(area); free(gone); free(ord);
    return K;
}

/* ---------------------------------------------------------------------------------- */
/* helper.bbox_iou on relative xc,yc,w,h boxes (helper.py:221-263) fused with the       */
/* reductions of YOLOForw.get_target (yolo_forw.py:186-201):                            */
/*   best[m]  = first argmax_n iou(m,n)                                    (:187)       */
/*   noobj[n] = all_m (iou(m,n) < ignore_thr), then cleared at best[m]     (:200-201)   */
/* kind 0 = IoU, 1 = GIoU.  gt [M,4], anc [N,4] are xc,yc,w,h.                          */
/* ---------------------------------------------------------------------------------- */
static inline void to_xyxy(const float* b, float* o) {
    /* helper.get_abs_coord (helper.py:203-217): x -/+ w/2 */
    o[0] = b[0] - b[2] / 2.0f; o[1] = b[1] - b[3] / 2.0f;
    o[2] = b[0] + b[2] / 2.0f; o[3] = b[1] + b[3] / 2.0f;
}
float ref_pair_iou(const float* p, const float* q, int kind) {
    /* p, q are xyxy.  helper.py:249-263 */
    float iw = fminf_(p[2], q[2]) - fmaxf_(p[0], q[0]);
    float ih = fminf_(p[3], q[3]) - fmaxf_(p[1], q[1]);
    if (iw < 0.0f) iw = 0.0f;
    if (ih < 0.0f) ih = 0.0f;
    float inter = iw * ih;
    float w1 = p[2] - p[0], h1 = p[3] - p[1];
    float w2 = q[2] - q[0], h2 = q[3] - q[1];
    float uni = ((w1 * h1 + 1e-16f) + w2 * h2) - inter;
    float iou = inter / uni;
    if (kind == 1) {
        float cw = fmaxf_(p[2], q[2]) - fminf_(p[0], q[0]);
        float chh = fmaxf_(p[3], q[3]) - fminf_(p[1], q[1]);
        float carea = cw * chh + 1e-16f;
        return iou - (carea - uni) / carea;
    }
    return iou;
}
void ref_iou_match(const float* gt, int M, const float* anc, int N, int kind, float ignore_thr,
                   int64_t* best, unsigned char* noobj, float* iou_out /* [M,N] or NULL */) {
    float* a_xyxy = (float*)malloc(sizeof(float) * 4 * (size_t)(N > 0 ? N : 1));
    for (int n = 0; n < N; ++n) to_xyxy(anc + 4 * (size_t)n, a_xyxy + 4 * (size_t)n);
    for (int n = 0; n < N; ++n) noobj[n] = 1;
    for (int m = 0; m < M; ++m) {
        float g[4];
        to_xyxy(gt + 4 * (size_t)m, g);
        float bv = -INFINITY; int64_t bi = 0; int have = 0;
        for (int n = 0; n < N; ++n) {
            float v = ref_pair_iou(g, a_xyxy + 4 * (size_t)n, kind);
            if (iou_out) iou_out[(size_t)m * (size_t)N + (size_t)n] = v;
            if (!have || v > bv) { bv = v; bi = n; have = 1; }
            if (!(v < ignore_thr)) noobj[n] = 0;
        }
        best[m] = bi;
    }
    for (int m = 0; m < M; ++m) noobj[best[m]] = 0;
    free(a_xyxy);
}

/* ---------------------------------------------------------------------------------- */
/* torchvision.ops.box_iou (boxes.py _box_inter_union): xyxy,                           */
/*   union = (area1 + area2) - inter ; iou = inter/union                                */
/* ---------------------------------------------------------------------------------- */
void ref_box_iou_tv(const float* b1, int M, const float* b2, int N, float* out) {
    for (int m = 0; m < M; ++m) {
        const float* p = b1 + 4 * (size_t)m;
        float a1 = (p[2] - p[0]) * (p[3] - p[1]);
        for (int n = 0; n < N; ++n) {
            const float* q = b2 + 4 * (size_t)n;
            float a2 = (q[2] - q[0]) * (q[3] - q[1]);
            float w = fminf_(p[2], q[2]) - fmaxf_(p[0], q[0]);
            float h = fminf_(p[3], q[3]) - fmaxf_(p[1], q[1]);
            if (w < 0.0f) w = 0.0f;
            if (h < 0.0f) h = 0.0f;
            float inter = w * h;
            out[(size_t)m * (size_t)N + (size_t)n] = inter / ((a1 + a2) - inter);
        }
    }
}
