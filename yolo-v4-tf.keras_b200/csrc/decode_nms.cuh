// Head decode + score filter (get_boxes / nms flattening, custom_layers.py:221-284) and
// tf.image.combined_non_max_suppression (custom_layers.py:290-297; semantics: SURVEY.md App. D.8).
// All arithmetic fp32 with explicit round-to-nearest intrinsics (no FMA contraction), so that every
// threshold decision is made on the same values a CPU evaluation of the reference formulas produces.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace y4 {

constexpr int kCandCap = 8192;     // Y4_MAX_CANDIDATES
constexpr int kSelCap = 8192;      // >= num_classes * max_boxes
constexpr int kMaxBoxesCap = 128;  // max_boxes upper bound (per-warp selected list in smem)
constexpr int kNmsThreads = 1024;

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

// One warp per grid cell (3 anchors x (5+nc) logits, contiguous in NHWC): coalesced read, sigmoid only
// where obj > threshold can still pass (score = obj*cls <= obj), warp-aggregated append.
__global__ void __launch_bounds__(256) decode_filter_kernel(DecodeParams p) {
    const int lane = threadIdx.x & 31;
    const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int cells = p.cell_off[3];
    if (gw >= (long long)p.batch * cells) return;
    const int img = (int)(gw / cells);
    const int cell = (int)(gw - (long long)img * cells);
    const int s = cell < p.cell_off[1] ? 0 : (cell < p.cell_off[2] ? 1 : 2);
    const int lc = cell - p.cell_off[s];
    const int g = p.g[s];
    const int row = lc / g, col = lc - row * g;
    const float* ptr = p.padded[s]
        ? p.head[s] + (((long long)img * (g + 2) + row + 1) * (g + 2) + col + 1) * p.ld[s]
        : p.head[s] + (((long long)img * g + row) * g + col) * p.ld[s];

    float obj[3];
#pragma unroll
    for (int a = 0; a < 3; a++) obj[a] = sigmoid_rn(ptr[a * p.C + 4]);
    // score = obj * cls <= obj (cls <= 1, round-to-nearest product): nothing in this cell can pass -> warp-uniform exit
    if (!(obj[0] > p.score_thr || obj[1] > p.score_thr || obj[2] > p.score_thr)) return;
    unsigned anymask = 0;
    const int total = 3 * p.C;
    const int iters = (total + 31) >> 5;
    for (int j = 0; j < iters; j++) {
        const int e = lane + 32 * j;
        bool cand = false;
        float score = 0.f;
        int a = 0, f = 0;
        if (e < total) {
            a = e / p.C;
            f = e - a * p.C;
            if (f >= 5 && obj[a] > p.score_thr) {
                score = __fmul_rn(obj[a], sigmoid_rn(ptr[e]));       // confidence * class_probabilities (custom_layers.py:282)
                cand = score > p.score_thr;                           // strict >
            }
        }
        const unsigned mask = __ballot_sync(0xffffffffu, cand);
        if (mask) {
            int base = 0;
            if (lane == 0) base = atomicAdd(&p.cand_count[img], __popc(mask));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (cand) {
                const int pos = base + __popc(mask & ((1u << lane) - 1u));
                const int nbox = p.box_off[s] + lc * 3 + a;
                if (pos < kCandCap) p.cand_keys[(long long)img * kCandCap + pos] = cand_key(f - 5, score, nbox);
            }
#pragma unroll
            for (int k = 0; k < 3; k++)
                if (__ballot_sync(0xffffffffu, cand && a == k)) anymask |= 1u << k;
        }
    }
    if (anymask && lane < 3 && ((anymask >> lane) & 1u)) {
        const int a = lane;
        const float* q = ptr + a * p.C;
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
        p.boxes[(long long)img * p.N + p.box_off[s] + lc * 3 + a] = b;
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

__device__ __forceinline__ void bitonic_sort_smem(unsigned long long* keys, int P) {
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < P; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long a = keys[i], b = keys[ixj];
                    const bool asc = (i & k) == 0;
                    if ((a > b) == asc) { keys[i] = b; keys[ixj] = a; }
                }
            }
            __syncthreads();
        }
    }
}

struct NmsParams {
    const unsigned long long* cand_keys;   // [batch][kCandCap]
    const int* cand_count;                 // [batch]
    const float4* boxes;                   // [batch][N]
    int N, nc, max_boxes;
    float iou_thr;
    float* out_boxes;      // [batch][max_boxes][4]
    float* out_scores;     // [batch][max_boxes]
    float* out_classes;    // [batch][max_boxes]
    int* out_valid;        // [batch]
    int* out_idx;          // [batch][max_boxes]
    int* overflow;         // set to 1 if any image exceeded kCandCap
};

// One CTA per image.  Candidates are grouped by class with a counting sort (histogram + scatter), each class segment
// is ordered (score desc, box asc) by a warp rank-sort (segments are short: ~candidates/80), one warp per class runs the
// greedy suppression (boxes fetched 32 at a time and broadcast by shuffle; every lane tests the candidate against a
// strided subset of the already-selected boxes, warp vote), and the per-class winner lists - already in score order -
// are merged by global ranking (score desc, class asc, box asc) into the top max_boxes, clipped to [0,1].
// Segments longer than kRankSortMax are rank-sorted by the whole CTA (cost ~ L^2 / 1024 per thread).
constexpr int kRankSortMax = 256;

__global__ void __launch_bounds__(kNmsThreads) nms_kernel(NmsParams p) {
    extern __shared__ __align__(16) unsigned char nms_smem[];
    unsigned long long* bufA = reinterpret_cast<unsigned long long*>(nms_smem);      // keys, finally class-sorted
    unsigned long long* bufB = bufA + kCandCap;                                       // scatter target, then winner lists
    float4* wbox = reinterpret_cast<float4*>(bufB + kSelCap);                         // [32 warps][kMaxBoxesCap]
    __shared__ int hist[256], start[257], cursor[256], nwin[256];
    __shared__ int big;

    const int img = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int cnt = p.cand_count[img];
    if (cnt > kCandCap) { if (tid == 0) *p.overflow = 1; cnt = kCandCap; }
    if (tid < 256) { hist[tid] = 0; nwin[tid] = 0; }
    if (tid == 0) big = 0;
    __syncthreads();
    for (int i = tid; i < cnt; i += blockDim.x) {
        const unsigned long long k = p.cand_keys[(long long)img * kCandCap + i];
        bufA[i] = k;
        atomicAdd(&hist[(int)(k >> 56)], 1);
    }
    __syncthreads();
    if (tid == 0) {
        int acc = 0, mx = 0;
        for (int c = 0; c < p.nc; c++) { start[c] = acc; cursor[c] = acc; acc += hist[c]; mx = hist[c] > mx ? hist[c] : mx; }
        start[p.nc] = acc;
        big = mx > kRankSortMax;
    }
    __syncthreads();
    for (int i = tid; i < cnt; i += blockDim.x) {
        const unsigned long long k = bufA[i];
        bufB[atomicAdd(&cursor[(int)(k >> 56)], 1)] = k;
    }
    __syncthreads();
    // rank sort inside each class segment: keys are unique (box index), rank = #smaller keys.
    // short segments: one warp each; segments longer than kRankSortMax: the whole CTA, one segment after the other
    for (int c = warp; c < p.nc; c += (kNmsThreads >> 5)) {
        const int lo = start[c], L = hist[c];
        if (L > kRankSortMax) continue;
        for (int i = lane; i < L; i += 32) {
            const unsigned long long k = bufB[lo + i];
            int r = 0;
            for (int j = 0; j < L; j++) r += bufB[lo + j] < k;
            bufA[lo + r] = k;
        }
    }
    if (big) {
        for (int c = 0; c < p.nc; c++) {
            const int lo = start[c], L = hist[c];
            if (L <= kRankSortMax) continue;                     // block-uniform
            for (int i = tid; i < L; i += blockDim.x) {
                const unsigned long long k = bufB[lo + i];
                int r = 0;
                for (int j = 0; j < L; j++) r += bufB[lo + j] < k;
                bufA[lo + r] = k;
            }
        }
    }
    __syncthreads();
    const unsigned long long* keys = bufA;

    const float4* boxes = p.boxes + (long long)img * p.N;
    float4* mybox = wbox + warp * kMaxBoxesCap;
    for (int c = warp; c < p.nc; c += (kNmsThreads >> 5)) {
        const int lo = start[c], hi = start[c] + hist[c];
        unsigned long long* win = bufB + (size_t)c * p.max_boxes;       // winners of class c, in score order
        int nsel = 0;
        for (int base = lo; base < hi && nsel < p.max_boxes; base += 32) {
            const int mine = base + lane;
            unsigned long long mykey = 0ull;
            float4 mybx = make_float4(0.f, 0.f, 0.f, 0.f);
            if (mine < hi) { mykey = keys[mine]; mybx = boxes[(int)(mykey & 0xFFFFFFull)]; }
            const int cnt32 = hi - base < 32 ? hi - base : 32;
            for (int j = 0; j < cnt32 && nsel < p.max_boxes; j++) {
                float4 b;
                b.x = __shfl_sync(0xffffffffu, mybx.x, j); b.y = __shfl_sync(0xffffffffu, mybx.y, j);
                b.z = __shfl_sync(0xffffffffu, mybx.z, j); b.w = __shfl_sync(0xffffffffu, mybx.w, j);
                const unsigned long long key = __shfl_sync(0xffffffffu, mykey, j);
                bool sup = false;
                for (int t = lane; t < nsel; t += 32) sup |= iou_tf(b, mybox[t]) > p.iou_thr;   // strict >
                if (!__any_sync(0xffffffffu, sup)) {
                    if (lane == 0) {
                        mybox[nsel] = b;
                        const float score = __uint_as_float(~(unsigned)((key >> 24) & 0xFFFFFFFFull));
                        win[nsel] = merge_key(c, score, (int)(key & 0xFFFFFFull));
                    }
                    nsel++;
                    __syncwarp();
                }
            }
        }
        if (lane == 0) nwin[c] = nsel;
    }
    __syncthreads();

    // merge: compact the per-class winner lists, then every winner computes its global rank (number of smaller merge
    // keys = score desc, class asc, box asc; keys are unique) and the first max_boxes ranks are the output order
    __shared__ unsigned long long outkeys[kMaxBoxesCap];
    __shared__ int nout;
    if (tid == 0) {
        int acc = 0;
        for (int c = 0; c < p.nc; c++) { cursor[c] = acc; acc += nwin[c]; }
        nout = acc;
    }
    __syncthreads();
    const int ntot = nout;
    for (int c = warp; c < p.nc; c += (kNmsThreads >> 5))
        for (int i = lane; i < nwin[c]; i += 32) bufA[cursor[c] + i] = bufB[(size_t)c * p.max_boxes + i];
    __syncthreads();
    for (int i = tid; i < ntot; i += blockDim.x) {
        const unsigned long long k = bufA[i];
        int r = 0;
        for (int j = 0; j < ntot && r < p.max_boxes; j++) r += bufA[j] < k;
        if (r < p.max_boxes) outkeys[r] = k;
    }
    __syncthreads();
    if (tid == 0) nout = ntot < p.max_boxes ? ntot : p.max_boxes;
    __syncthreads();

    const int nvalid = nout;
    if (tid == 0) p.out_valid[img] = nvalid;
    for (int k = tid; k < p.max_boxes; k += blockDim.x) {
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

constexpr size_t kNmsSmemBytes = (size_t)(kCandCap + kSelCap) * 8 + (size_t)(kNmsThreads / 32) * kMaxBoxesCap * 16;

}  // namespace y4
