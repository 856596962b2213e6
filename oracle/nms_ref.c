/* ORACLE — TEST / BASELINE INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Plain-C restatement of tf.image.combined_non_max_suppression as the reference calls it
 * (/root/reference/custom_layers.py:290-297: max_output_size_per_class = max_total_size = 100, pad_per_class=False,
 * clip_boxes=True) — the same semantics as oracle/y4_oracle.py:combined_nms, in compiled code so that
 * the CPU baseline of bench.py is not dominated by Python loops (TF's own kernel is C++ and shards (image, class) pairs
 * over the host threads; so does this file with OpenMP).  PARITY UNPINNED like the Python oracle: TensorFlow's source is
 * not in /root/reference; tests/test_oracle.py checks this file against the Python restatement bit for bit.
 *
 *   gcc -O2 -fPIC -shared -ffp-contract=off -fopenmp oracle/nms_ref.c -o oracle/libnmsref.so -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float score; int32_t box; int32_t cls; } cand_t;

/* TF's IOU: corners re-ordered with min/max, area <= 0 -> 0 */
static float iou_tf(const float* a, const float* b) {
    const float ya0 = fminf(a[0], a[2]), ya1 = fmaxf(a[0], a[2]), xa0 = fminf(a[1], a[3]), xa1 = fmaxf(a[1], a[3]);
    const float yb0 = fminf(b[0], b[2]), yb1 = fmaxf(b[0], b[2]), xb0 = fminf(b[1], b[3]), xb1 = fmaxf(b[1], b[3]);
    const float area_a = (ya1 - ya0) * (xa1 - xa0), area_b = (yb1 - yb0) * (xb1 - xb0);
    if (area_a <= 0.f || area_b <= 0.f) return 0.f;
    const float iy0 = fmaxf(ya0, yb0), ix0 = fmaxf(xa0, xb0), iy1 = fminf(ya1, yb1), ix1 = fminf(xa1, xb1);
    const float inter = fmaxf(iy1 - iy0, 0.f) * fmaxf(ix1 - ix0, 0.f);
    return inter / ((area_a + area_b) - inter);
}

static int by_score_desc_box_asc(const void* pa, const void* pb) {
    const cand_t* a = (const cand_t*)pa; const cand_t* b = (const cand_t*)pb;
    if (a->score != b->score) return a->score > b->score ? -1 : 1;
    return a->box < b->box ? -1 : (a->box > b->box);
}
static int by_score_desc_cls_box_asc(const void* pa, const void* pb) {
    const cand_t* a = (const cand_t*)pa; const cand_t* b = (const cand_t*)pb;
    if (a->score != b->score) return a->score > b->score ? -1 : 1;
    if (a->cls != b->cls) return a->cls < b->cls ? -1 : 1;
    return a->box < b->box ? -1 : (a->box > b->box);
}

/* boxes (B,N,4) normalised x1,y1,x2,y2; scores (B,N,C).  Outputs as y4_oracle.combined_nms: nmsed boxes (B,T,4) clipped to
 * [0,1], scores (B,T), classes (B,T) float, valid (B), cand_idx (B,T) (-1 padded).  Returns 0. */
int y4ref_combined_nms(const float* boxes, const float* scores, int32_t B, int32_t N, int32_t C, float iou_thr, float score_thr,
                       int32_t max_per_class, int32_t T, float* out_boxes, float* out_scores, float* out_classes,
                       int32_t* out_valid, int32_t* out_idx) {
    memset(out_boxes, 0, sizeof(float) * 4 * (size_t)B * T);
    memset(out_scores, 0, sizeof(float) * (size_t)B * T);
    memset(out_classes, 0, sizeof(float) * (size_t)B * T);
    for (size_t i = 0; i < (size_t)B * T; i++) out_idx[i] = -1;
    cand_t* picked_all = (cand_t*)malloc(sizeof(cand_t) * (size_t)B * C * max_per_class);
    int32_t* npicked = (int32_t*)calloc((size_t)B * C, sizeof(int32_t));
#pragma omp parallel
    {
        cand_t* cand = (cand_t*)malloc(sizeof(cand_t) * (size_t)N);
#pragma omp for collapse(2) schedule(dynamic, 4)
        for (int32_t b = 0; b < B; b++)
            for (int32_t c = 0; c < C; c++) {
                const float* sc = scores + (size_t)b * N * C;
                const float* bx = boxes + (size_t)b * N * 4;
                int32_t n = 0;
                for (int32_t i = 0; i < N; i++) {
                    const float s = sc[(size_t)i * C + c];
                    if (s > score_thr) { cand[n].score = s; cand[n].box = i; cand[n].cls = c; n++; }      /* strict > */
                }
                if (!n) continue;
                qsort(cand, (size_t)n, sizeof(cand_t), by_score_desc_box_asc);
                cand_t* sel = picked_all + ((size_t)b * C + c) * max_per_class;
                int32_t ns = 0;
                for (int32_t j = 0; j < n && ns < max_per_class; j++) {
                    int keep = 1;
                    for (int32_t m = ns - 1; m >= 0; m--)                                   /* newest selected first */
                        if (iou_tf(bx + 4 * (size_t)cand[j].box, bx + 4 * (size_t)sel[m].box) > iou_thr) { keep = 0; break; }   /* strict > */
                    if (keep) sel[ns++] = cand[j];
                }
                npicked[(size_t)b * C + c] = ns;
            }
        free(cand);
    }
#pragma omp parallel for schedule(dynamic, 1)
    for (int32_t b = 0; b < B; b++) {
        int32_t total = 0;
        for (int32_t c = 0; c < C; c++) total += npicked[(size_t)b * C + c];
        cand_t* all = (cand_t*)malloc(sizeof(cand_t) * (size_t)(total > 0 ? total : 1));
        int32_t k = 0;
        for (int32_t c = 0; c < C; c++) {
            memcpy(all + k, picked_all + ((size_t)b * C + c) * max_per_class, sizeof(cand_t) * (size_t)npicked[(size_t)b * C + c]);
            k += npicked[(size_t)b * C + c];
        }
        qsort(all, (size_t)total, sizeof(cand_t), by_score_desc_cls_box_asc);
        const int32_t nv = total < T ? total : T;
        out_valid[b] = nv;
        for (int32_t i = 0; i < nv; i++) {
            const float* bb = boxes + ((size_t)b * N + all[i].box) * 4;
            for (int q = 0; q < 4; q++) out_boxes[((size_t)b * T + i) * 4 + q] = fminf(fmaxf(bb[q], 0.f), 1.f);   /* clip on output only */
            out_scores[(size_t)b * T + i] = all[i].score;
            out_classes[(size_t)b * T + i] = (float)all[i].cls;
            out_idx[(size_t)b * T + i] = all[i].box;
        }
        free(all);
    }
    free(picked_all); free(npicked);
    return 0;
}
