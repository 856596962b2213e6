// Head decode + score filter (get_boxes / nms flattening, custom_layers.py:221-284) and
// tf.image.combined_non_max_suppression (custom_layers.py:290-297; semantics: SURVEY.md App. D.8).
// All arithmetic fp32 with explicit round-to-nearest intrinsics (no FMA contraction), so that every
// threshold decision is made on the same values a CPU evaluation of the reference formulas produces.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace y4 {

constexpr int kCandCap = 8192;     // Y4_MAX_CANDIDATES: candidate-list capacity per image of the FAST path; images with more
                                   // candidates are handled exactly by nms_overflow_kernel (no limit, as in TF)
constexpr int kSelCap = 8192;      // >= num_classes * max_boxes
constexpr int kMaxBoxesCap = 128;  // max_boxes upper bound (per-warp selected list in smem)

struct DecodeParams {
    const float* head[3];
    int ld[3];            // floats per cell row (255 packed, or padded ld)
    int padded[3];        // 1: padded-flat (g+2)x(g+2) with halo, 0: packed (B,g,g,ld)
    int g[3];
    float stride[3];
    float xyscale[3];     // float32(xyscale)
    float xyoff[3];       // float32(0.5 * (xyscale - 1)) evaluated in double (custom_layers.py:251)
    float anchors[18];
    int cell_off[4];      // cumulative cells per image
    int box_off[3];       // first flat box index of each scale: 0, 3*g0^2, 3*(g0^2+g1^2)
    int nc, C;            // classes, 5 + classes
    int N;                // boxes per image
    int batch;
    float img_size;
    float score_thr;
    unsigned long long* cand_keys;   // [batch][kCandCap]
    int* cand_count;                 // [batch]
    float4* boxes;                   // [batch][N] normalised x1,y1,x2,y2 (written only for boxes with a candidate)
};

__device__ __forceinline__ float sigmoid_rn(float x) {
    return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x)));
}

// sort key: class(8) | ~score_bits(32) | box(24): ascending == class asc, score desc, box asc
__device__ __forceinline__ unsigned long long cand_key(int cls, float score, int box) {
    return ((unsigned long long)cls << 56) | ((unsigned long long)(~__float_as_uint(score)) << 24) | (unsigned long long)box;
}
// merge key: ~score_bits(32) | class(8) | box(24): ascending == score desc, class asc, box asc
__device__ __forceinline__ unsigned long long merge_key(int cls, float score, int box) {
    return ((unsigned long long)(~__float_as_uint(score)) << 32) | ((unsigned long long)cls << 24) | (unsigned long long)box;
}

// The 5 + nc logits of flat box n = off_scale + (row*g + col)*3 + a of image img (custom_layers.py:232-237, 274-280).
__device__ __forceinline__ const float* box_logits(const DecodeParams& p, int img, int n, int& s, int& a, int& row, int& col) {
    s = n < p.box_off[1] ? 0 : (n < p.box_off[2] ? 1 : 2);
    const int ln = n - p.box_off[s];
    const int lc = ln / 3;
    a = ln - lc * 3;
    const int g = p.g[s];
    row = lc / g; col = lc - row * g;
    const float* cellp = p.padded[s]
        ? p.head[s] + (((long long)img * (g + 2) + row + 1) * (g + 2) + col + 1) * p.ld[s]
        : p.head[s] + (((long long)img * g + row) * g + col) * p.ld[s];
    return cellp + a * p.C;
}

// One thread per box (cell, anchor): objectness test first (score = obj * cls <= obj since cls <= 1 and the product is
// rounded to nearest, so a box whose objectness fails can produce no candidate).  The few boxes that pass are then expanded
// by the whole warp, one after the other: lanes take the classes, candidates are appended with one atomic per ballot, and
// the box is decoded once if any class passed.  (The earlier warp-per-cell version spent its time launching 240 k warps per
// batch that exit after three loads.)
__global__ void __launch_bounds__(256) decode_filter_kernel(DecodeParams p) {
    const int lane = threadIdx.x & 31;
    const long long gt = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)p.batch * p.N;
    int img = 0, n = 0, s = 0, a = 0, row = 0, col = 0;
    const float* q = nullptr;
    float obj = 0.f;
    bool pass = false;
    if (gt < total) {
        img = (int)(gt / p.N);
        n = (int)(gt - (long long)img * p.N);
        q = box_logits(p, img, n, s, a, row, col);
        obj = sigmoid_rn(q[4]);
        pass = obj > p.score_thr;
    }
    unsigned todo = __ballot_sync(0xffffffffu, pass);
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const float* bq = reinterpret_cast<const float*>(__shfl_sync(0xffffffffu, (unsigned long long)q, src));
        const float bobj = __shfl_sync(0xffffffffu, obj, src);
        const int bimg = __shfl_sync(0xffffffffu, img, src), bn = __shfl_sync(0xffffffffu, n, src);
        bool any = false;
        for (int f0 = 0; f0 < p.nc; f0 += 32) {
            const int f = f0 + lane;
            bool cand = false;
            float score = 0.f;
            if (f < p.nc) {
                score = __fmul_rn(bobj, sigmoid_rn(bq[5 + f]));          // confidence * class_probabilities (custom_layers.py:282)
                cand = score > p.score_thr;                               // strict >
            }
            const unsigned mask = __ballot_sync(0xffffffffu, cand);
            if (mask) {
                any = true;
                int base = 0;
                if (lane == 0) base = atomicAdd(&p.cand_count[bimg], __popc(mask));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (cand) {
                    const int pos = base + __popc(mask & ((1u << lane) - 1u));
                    if (pos < kCandCap) p.cand_keys[(long long)bimg * kCandCap + pos] = cand_key(f, score, bn);
                }
            }
        }
        if (any && lane == src) {
            const float sx = sigmoid_rn(q[0]), sy = sigmoid_rn(q[1]);
            const float bx = __fmul_rn(__fadd_rn(__fsub_rn(__fmul_rn(sx, p.xyscale[s]), p.xyoff[s]), (float)col), p.stride[s]);
            const float by = __fmul_rn(__fadd_rn(__fsub_rn(__fmul_rn(sy, p.xyscale[s]), p.xyoff[s]), (float)row), p.stride[s]);
            const float bw = __fmul_rn(expf(q[2]), p.anchors[(s * 3 + a) * 2 + 0]);
            const float bh = __fmul_rn(expf(q[3]), p.anchors[(s * 3 + a) * 2 + 1]);
            const float hw = __fmul_rn(bw, 0.5f), hh = __fmul_rn(bh, 0.5f);   // box_wh / 2 (exact)
            float4 b;
            b.x = __fdiv_rn(__fsub_rn(bx, hw), p.img_size);                   // boxes / input_shape[0]  (custom_layers.py:284)
            b.y = __fdiv_rn(__fsub_rn(by, hh), p.img_size);
            b.z = __fdiv_rn(__fadd_rn(bx, hw), p.img_size);
            b.w = __fdiv_rn(__fadd_rn(by, hh), p.img_size);
            p.boxes[(long long)img * p.N + n] = b;
        }
    }
}

// TF's IOU (non_max_suppression_op.cc): coordinates re-ordered with min/max, area <= 0 -> 0.
__device__ __forceinline__ float iou_tf(const float4 a, const float4 b) {
    const float ya0 = fminf(a.x, a.z), ya1 = fmaxf(a.x, a.z), xa0 = fminf(a.y, a.w), xa1 = fmaxf(a.y, a.w);
    const float yb0 = fminf(b.x, b.z), yb1 = fmaxf(b.x, b.z), xb0 = fminf(b.y, b.w), xb1 = fmaxf(b.y, b.w);
    const float area_a = __fmul_rn(__fsub_rn(ya1, ya0), __fsub_rn(xa1, xa0));
    const float area_b = __fmul_rn(__fsub_rn(yb1, yb0), __fsub_rn(xb1, xb0));
    if (area_a <= 0.f || area_b <= 0.f) return 0.f;
    const float iy0 = fmaxf(ya0, yb0), ix0 = fmaxf(xa0, xb0), iy1 = fminf(ya1, yb1), ix1 = fminf(xa1, xb1);
    const float inter = __fmul_rn(fmaxf(__fsub_rn(iy1, iy0), 0.f), fmaxf(__fsub_rn(ix1, ix0), 0.f));
    return __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_a, area_b), inter));
}

struct NmsParams {
    const unsigned long long* cand_keys;   // [batch][kCandCap]   decode output, unordered
    const int* cand_count;                 // [batch]
    const float4* boxes;                   // [batch][N]
    int N, nc, max_boxes;
    float iou_thr;
    // workspace
    unsigned long long* bucket_keys;       // [batch][kCandCap]   grouped by class (order inside a class arbitrary)
    unsigned long long* sorted_keys;       // [batch][kCandCap]   only used by class segments longer than kClassSmemKeys
    int* seg_start;                        // [batch][257]        first bucket position of each class
    unsigned long long* win_keys;          // [batch][nc][max_boxes]  per-class NMS survivors as merge keys, score descending
    int* nwin;                             // [batch][256]
    float* out_boxes;      // [batch][max_boxes][4]
    float* out_scores;     // [batch][max_boxes]
    float* out_classes;    // [batch][max_boxes]
    int* out_valid;        // [batch]
    int* out_idx;          // [batch][max_boxes]
};

// combined_non_max_suppression as three small kernels whose parallelism is (image, class), not image:
//   nms_bucket_kernel  one CTA per image: counting sort of the candidate keys by class (smem histogram + scatter)
//   nms_class_kernel   one CTA per (image, class): rank sort of the segment (score desc, box asc; keys are unique),
//                      then warp 0 runs TF's greedy scan (iou > thr strict, against already selected boxes)
//                      (segments longer than kClassSmemKeys: greedy by repeated block-wide arg-max, no sort)
//   nms_overflow_kernel one CTA per (image, class), only for images whose candidates did not fit kCandCap: the same greedy
//                      selection by repeated arg-max straight from the head tensors -- exact for ANY number of candidates
//   nms_merge_kernel   one CTA per image: k-way merge of the per-class survivor lists, first max_boxes, clipped to [0,1]
constexpr int kBucketThreads = 1024;
constexpr int kClassThreads = 128;
constexpr int kClassSmemKeys = 1024;
constexpr int kMergeThreads = 256;                            // >= num_classes (255 max): one thread per class list

// block-wide minimum of a 64-bit key, result in every thread (T threads, all participate; contains two barriers)
template <int T>
__device__ __forceinline__ unsigned long long block_min_u64(unsigned long long v) {
    __shared__ unsigned long long wmin_[T / 32];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) { const unsigned long long o = __shfl_xor_sync(0xffffffffu, v, d); v = o < v ? o : v; }
    __syncthreads();                                        // previous round's readers are done with wmin_
    if ((threadIdx.x & 31) == 0) wmin_[threadIdx.x >> 5] = v;
    __syncthreads();
    unsigned long long best = wmin_[0];
#pragma unroll
    for (int w = 1; w < T / 32; w++) best = wmin_[w] < best ? wmin_[w] : best;
    return best;
}

__global__ void __launch_bounds__(kBucketThreads) nms_bucket_kernel(NmsParams p) {
    __shared__ int hist[256], cursor[256];
    const int img = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
    int cnt = p.cand_count[img];
    if (cnt > kCandCap) cnt = kCandCap;                     // this image is redone exactly by nms_overflow_kernel
    if (tid < 256) hist[tid] = 0;
    __syncthreads();
    constexpr int PER = kCandCap / kBucketThreads;
    unsigned long long k[PER];
#pragma unroll
    for (int r = 0; r < PER; r++) {
        const int i = tid + r * kBucketThreads;
        if (i < cnt) { k[r] = p.cand_keys[(long long)img * kCandCap + i]; atomicAdd(&hist[(int)(k[r] >> 56)], 1); }
    }
    __syncthreads();
    if (tid < 32) {                                        // exclusive scan of the 256 bins, 8 per lane
        int v[8], s = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) { v[j] = hist[lane * 8 + j]; s += v[j]; }
        int incl = s;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
        int run = incl - s;
#pragma unroll
        for (int j = 0; j < 8; j++) { cursor[lane * 8 + j] = run; p.seg_start[img * 257 + lane * 8 + j] = run; run += v[j]; }
        if (lane == 31) p.seg_start[img * 257 + 256] = run;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < PER; r++) {
        const int i = tid + r * kBucketThreads;
        if (i < cnt) p.bucket_keys[(long long)img * kCandCap + atomicAdd(&cursor[(int)(k[r] >> 56)], 1)] = k[r];
    }
}

__global__ void __launch_bounds__(kClassThreads) nms_class_kernel(NmsParams p) {
    __shared__ unsigned long long sin[kClassSmemKeys], ssorted[kClassSmemKeys];
    __shared__ float4 selbox[kMaxBoxesCap];
    const int img = blockIdx.x / p.nc, c = blockIdx.x - img * p.nc;
    const int tid = threadIdx.x, lane = tid & 31;
    if (p.cand_count[img] > kCandCap) return;              // truncated candidate list: nms_overflow_kernel owns this image
    const int lo = p.seg_start[img * 257 + c], L = p.seg_start[img * 257 + c + 1] - lo;
    if (L == 0) { if (tid == 0) p.nwin[img * 256 + c] = 0; return; }
    const unsigned long long* in = p.bucket_keys + (long long)img * kCandCap + lo;
    unsigned long long* sorted = p.sorted_keys + (long long)img * kCandCap + lo;
    if (L > kClassSmemKeys) {
        // long segment (up to kCandCap keys of one class): sorting it would cost O(L^2) (rank sort) for at most max_boxes picks;
        // greedy selection by arg-max rounds instead: round r suppresses against the box picked in round r-1 and finds the
        // best remaining key, <= max_boxes passes over the segment, alive bits in shared memory
        __shared__ unsigned dead[kCandCap / 32];
        const int words = (L + 31) >> 5;
        for (int w = tid; w < words; w += kClassThreads) dead[w] = (w * 32 + 32 <= L) ? 0u : ~((1u << (L - w * 32)) - 1u);
        __syncthreads();
        const float4* bx = p.boxes + (long long)img * p.N;
        unsigned long long* wk = p.win_keys + ((long long)img * p.nc + c) * p.max_boxes;
        float4 sel = make_float4(0.f, 0.f, 0.f, 0.f);
        int ns = 0;
        while (ns < p.max_boxes) {
            unsigned long long mine = ~0ull;
            int mypos = 0;
            for (int w = tid; w < words; w += kClassThreads) {
                unsigned m = ~dead[w], kill = 0u;
                while (m) {
                    const int b = __ffs(m) - 1;
                    m &= m - 1;
                    const unsigned long long k = in[w * 32 + b];
                    if (ns > 0 && iou_tf(bx[(int)(k & 0xFFFFFFull)], sel) > p.iou_thr) { kill |= 1u << b; continue; }   // strict >
                    if (k < mine) { mine = k; mypos = w * 32 + b; }
                }
                dead[w] |= kill;
            }
            const unsigned long long best = block_min_u64<kClassThreads>(mine);
            if (best == ~0ull) break;
            const int box = (int)(best & 0xFFFFFFull);
            if (mine == best) {                             // keys are unique: exactly one owner, and it owns that word of dead[]
                wk[ns] = merge_key(c, __uint_as_float(~(unsigned)((best >> 24) & 0xFFFFFFFFull)), box);
                dead[mypos >> 5] |= 1u << (mypos & 31);
            }
            sel = bx[box];
            ns++;
            __syncthreads();
        }
        if (tid == 0) p.nwin[img * 256 + c] = ns;
        return;
    }
    if (L <= kClassSmemKeys) {
        for (int i = tid; i < L; i += kClassThreads) sin[i] = in[i];
        __syncthreads();
        in = sin; sorted = ssorted;
    }
    for (int i = tid; i < L; i += kClassThreads) {        // keys are unique (box index): rank = number of smaller keys
        const unsigned long long k = in[i];
        int r = 0;
        for (int j = 0; j < L; j++) r += in[j] < k;
        sorted[r] = k;
    }
    __syncthreads();
    if (tid >= 32) return;

    const float4* boxes = p.boxes + (long long)img * p.N;
    unsigned long long* win = p.win_keys + ((long long)img * p.nc + c) * p.max_boxes;
    int nsel = 0;
    for (int base = 0; base < L && nsel < p.max_boxes; base += 32) {
        const int mine = base + lane;
        unsigned long long mykey = 0ull;
        float4 mybx = make_float4(0.f, 0.f, 0.f, 0.f);
        if (mine < L) { mykey = sorted[mine]; mybx = boxes[(int)(mykey & 0xFFFFFFull)]; }
        const int cnt32 = L - base < 32 ? L - base : 32;
        for (int j = 0; j < cnt32 && nsel < p.max_boxes; j++) {
            float4 b;
            b.x = __shfl_sync(0xffffffffu, mybx.x, j); b.y = __shfl_sync(0xffffffffu, mybx.y, j);
            b.z = __shfl_sync(0xffffffffu, mybx.z, j); b.w = __shfl_sync(0xffffffffu, mybx.w, j);
            const unsigned long long key = __shfl_sync(0xffffffffu, mykey, j);
            bool sup = false;
            for (int t = lane; t < nsel; t += 32) sup |= iou_tf(b, selbox[t]) > p.iou_thr;   // strict >
            if (!__any_sync(0xffffffffu, sup)) {
                if (lane == 0) {
                    selbox[nsel] = b;
                    const float score = __uint_as_float(~(unsigned)((key >> 24) & 0xFFFFFFFFull));
                    win[nsel] = merge_key(c, score, (int)(key & 0xFFFFFFull));
                }
                nsel++;
                __syncwarp();
            }
        }
    }
    if (lane == 0) p.nwin[img * 256 + c] = nsel;
}

// Exact path for images with more than kCandCap candidates (low score thresholds, e.g. mAP export at 0.001; TF's
// CombinedNonMaxSuppression has no limit): one CTA per (image, class) works straight from the head tensors.  Pass 0 marks the
// boxes whose class score passes the threshold (same arithmetic as decode_filter_kernel, so the same candidate set); then at
// most max_boxes rounds, each ONE sweep over the still-alive boxes: suppress against the box selected in the previous round
// (iou > thr strict) and find the best remaining (score desc, box asc) -- identical to TF's sorted greedy scan, without ever
// materialising or sorting the candidate list.  Launched after nms_class_kernel on every step; returns at once for images
// that fitted the fast path.
constexpr int kOverflowThreads = 256;
__global__ void __launch_bounds__(kOverflowThreads) nms_overflow_kernel(DecodeParams d, NmsParams p) {
    extern __shared__ unsigned ov_dead[];                   // (N + 31) / 32 words: 1 = not a candidate / suppressed / selected
    const int img = blockIdx.x / p.nc, c = blockIdx.x - img * p.nc;
    if (p.cand_count[img] <= kCandCap) return;
    const int tid = threadIdx.x;
    const int words = (p.N + 31) >> 5;
    auto class_score = [&](int n) {
        int s, a, row, col;
        const float* q = box_logits(d, img, n, s, a, row, col);
        return __fmul_rn(sigmoid_rn(q[4]), sigmoid_rn(q[5 + c]));
    };
    for (int w = tid; w < words; w += kOverflowThreads) {
        unsigned alive = 0u;
        for (int b = 0; b < 32; b++) {
            const int n = w * 32 + b;
            if (n < p.N && class_score(n) > d.score_thr) alive |= 1u << b;      // strict >
        }
        ov_dead[w] = ~alive;
    }
    __syncthreads();
    const float4* bx = p.boxes + (long long)img * p.N;      // decode_filter_kernel wrote every box that has a candidate
    unsigned long long* wk = p.win_keys + ((long long)img * p.nc + c) * p.max_boxes;
    float4 sel = make_float4(0.f, 0.f, 0.f, 0.f);
    int ns = 0;
    while (ns < p.max_boxes) {
        unsigned long long best = ~0ull;
        for (int w = tid; w < words; w += kOverflowThreads) {
            unsigned m = ~ov_dead[w], kill = 0u;
            while (m) {
                const int b = __ffs(m) - 1;
                m &= m - 1;
                const int n = w * 32 + b;
                if (ns > 0 && iou_tf(bx[n], sel) > p.iou_thr) { kill |= 1u << b; continue; }
                const unsigned long long k = ((unsigned long long)(~__float_as_uint(class_score(n))) << 32) | (unsigned long long)n;
                best = k < best ? k : best;
            }
            ov_dead[w] |= kill;
        }
        best = block_min_u64<kOverflowThreads>(best);
        if (best == ~0ull) break;
        const int box = (int)(best & 0xFFFFFFFFull);
        if (tid == 0) {
            wk[ns] = merge_key(c, __uint_as_float(~(unsigned)(best >> 32)), box);
            ov_dead[box >> 5] |= 1u << (box & 31);
        }
        sel = bx[box];
        ns++;
        __syncthreads();
    }
    if (tid == 0) p.nwin[img * 256 + c] = ns;
}

__global__ void __launch_bounds__(kMergeThreads) nms_merge_kernel(NmsParams p) {
    // k-way merge of the per-class survivor lists (each already score-descending = merge-key ascending): thread c holds the
    // head of class c; max_boxes rounds of a block-wide min pick the output in order.  (Ranking every survivor against every
    // class list cost 135 us per batch; the output only needs the first max_boxes of the merged order.)
    __shared__ unsigned long long wmin[kMergeThreads / 32];
    __shared__ unsigned long long outkeys[kMaxBoxesCap];
    const int img = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned long long kNone = ~0ull;
    const unsigned long long* list = p.win_keys + ((long long)img * p.nc + tid) * p.max_boxes;
    const int cnt = tid < p.nc ? p.nwin[img * 256 + tid] : 0;
    int cur = 0;
    unsigned long long head = cnt > 0 ? list[0] : kNone;
    unsigned long long nxt = cnt > 1 ? list[1] : kNone;       // one element of lookahead hides the global-load latency
    int nvalid = 0;
    for (int k = 0; k < p.max_boxes; k++) {
        unsigned long long m = head;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) { const unsigned long long o = __shfl_xor_sync(0xffffffffu, m, d); m = o < m ? o : m; }
        if (lane == 0) wmin[warp] = m;
        __syncthreads();
        unsigned long long best = wmin[0];
#pragma unroll
        for (int w = 1; w < kMergeThreads / 32; w++) best = wmin[w] < best ? wmin[w] : best;
        __syncthreads();
        if (best == kNone) break;                             // block-uniform: every list is exhausted
        if (tid == 0) outkeys[k] = best;
        nvalid = k + 1;
        if (head == best) {                                   // keys are unique: exactly one owner
            cur++;
            head = nxt;
            nxt = cur + 1 < cnt ? list[cur + 1] : kNone;
        }
    }
    __syncthreads();
    const float4* boxes = p.boxes + (long long)img * p.N;
    if (tid == 0) p.out_valid[img] = nvalid;
    for (int k = tid; k < p.max_boxes; k += kMergeThreads) {
        float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
        float score = 0.f, cls = 0.f;
        int idx = -1;
        if (k < nvalid) {
            const unsigned long long key = outkeys[k];
            idx = (int)(key & 0xFFFFFFull);
            cls = (float)((key >> 24) & 0xFFull);
            score = __uint_as_float(~(unsigned)(key >> 32));
            b = boxes[idx];
            b.x = fminf(fmaxf(b.x, 0.f), 1.f); b.y = fminf(fmaxf(b.y, 0.f), 1.f);   // clip_boxes=True
            b.z = fminf(fmaxf(b.z, 0.f), 1.f); b.w = fminf(fmaxf(b.w, 0.f), 1.f);
        }
        const long long o = (long long)img * p.max_boxes + k;
        reinterpret_cast<float4*>(p.out_boxes)[o] = b;
        p.out_scores[o] = score;
        p.out_classes[o] = cls;
        p.out_idx[o] = idx;
    }
}

}  // namespace y4
