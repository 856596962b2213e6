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
constexpr int kMaxBoxesCap = 128;  // max_boxes upper bound (selected lists in smem, per-part survivor lists in global memory)

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
    float logit_lo;                  // a logit below this cannot pass score_thr: logit(thr) minus a margin (host, double; -inf / +inf when thr
                                     // is outside (0,1)).  sigmoid_rn is monotonic up to 1e-6 relative, the margin is 1e-3: the exact
                                     // sigmoid + compare runs only for logits at or above it, i.e. the decision is unchanged
    const float* obj[3];             // optional compact planar copy of the objectness logits, [3][obj_rows[s]] over the padded-flat
    long long obj_rows[3];           // pixel rows (written by the head convs' epilogue): the first pass then reads coalesced
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

// Two phases per CTA of 256 boxes of ONE image (grid.y = image):
//  1. one thread per box (warp-interleaved over the image, see below) tests objectness (score = obj * cls <= obj since cls <= 1 and the product is rounded to nearest, so a
//     box whose objectness fails can produce no candidate); the logits come from the compact planar copy the head convs write
//     (coalesced) when there is one.  The boxes that pass are compacted into a shared list.
//  2. the CTA's eight warps share that list evenly: one warp per box reads the whole 5 + nc row with every load in flight at
//     once (3 coalesced loads at 80 classes), lanes take the classes, lanes 0-3 decode the box.  Candidates are staged in shared
//     memory and appended to the image's list with ONE global atomic per CTA (keys that do not fit the stage -- more than
//     kStageKeys candidates among 256 boxes -- go out one by one).
// 64-bit keys `class | ~score_bits | box`; the order inside the list is arbitrary (nms_image_kernel sorts).
constexpr int kFilterThreads = 256;
constexpr int kStageKeys = 1024;
__global__ void __launch_bounds__(kFilterThreads) decode_filter_kernel(DecodeParams p) {
    __shared__ unsigned long long s_keys[kStageKeys];
    __shared__ int s_list[kFilterThreads];
    __shared__ int s_npass, s_nkeys, s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned lt = (1u << lane) - 1u;
    const int img = blockIdx.y;
    // warp w of CTA b takes the 32 consecutive boxes of global warp index w * gridDim.x + b: coalesced objectness reads, and every CTA
    // samples eight positions spread over the three scales (contiguous 256-box CTAs left the few CTAs that cover the 19^2 grid,
    // where detections are as numerous as on the 76^2 grid, with ten times the passing boxes of the others: a long tail)
    const int n = (warp * (int)gridDim.x + (int)blockIdx.x) * 32 + lane;
    if (tid == 0) { s_npass = 0; s_nkeys = 0; }
    __syncthreads();
    bool pass = false;
    if (n < p.N) {
        int s, a, row, col;
        const float* q = box_logits(p, img, n, s, a, row, col);
        const float lo = p.obj[s] ? p.obj[s][a * p.obj_rows[s] + ((long long)img * (p.g[s] + 2) + row + 1) * (p.g[s] + 2) + col + 1] : q[4];
        pass = !(lo < p.logit_lo) && sigmoid_rn(lo) > p.score_thr;
    }
    {
        const unsigned m = __ballot_sync(0xffffffffu, pass);
        if (m) {
            int base = 0;
            if (lane == 0) base = atomicAdd(&s_npass, __popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (pass) s_list[base + __popc(m & lt)] = n;
        }
    }
    __syncthreads();
    const int npass = s_npass;
    for (int it = warp; it < npass; it += kFilterThreads / 32) {
        const int bn = s_list[it];
        int s, a, row, col;
        const float* q = box_logits(p, img, bn, s, a, row, col);
        bool any = false;
        float v0 = 0.f, bobj = 0.f;
        for (int k0 = 0; k0 < p.C; k0 += 96) {
            float v[3];
#pragma unroll
            for (int k = 0; k < 3; k++) { const int idx = k0 + 32 * k + lane; v[k] = idx < p.C ? q[idx] : 0.f; }
            if (k0 == 0) { v0 = v[0]; bobj = sigmoid_rn(__shfl_sync(0xffffffffu, v0, 4)); }
            unsigned mk[3];
            float sc[3];
            int tot = 0;
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const int f = k0 + 32 * k + lane - 5;
                bool cand = false;
                sc[k] = 0.f;
                if (f >= 0 && f < p.nc && !(v[k] < p.logit_lo)) {             // score <= sigmoid(class logit): same pre-test
                    sc[k] = __fmul_rn(bobj, sigmoid_rn(v[k]));                // confidence * class_probabilities (custom_layers.py:282)
                    cand = sc[k] > p.score_thr;                               // strict >
                }
                mk[k] = __ballot_sync(0xffffffffu, cand);
                tot += __popc(mk[k]);
            }
            if (tot) {
                any = true;
                int base = 0;
                if (lane == 0) base = atomicAdd(&s_nkeys, tot);
                base = __shfl_sync(0xffffffffu, base, 0);
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    if ((mk[k] >> lane) & 1u) {
                        const int pos = base + __popc(mk[k] & lt);
                        const unsigned long long key = cand_key(k0 + 32 * k + lane - 5, sc[k], bn);
                        if (pos < kStageKeys) s_keys[pos] = key;
                        else {                                                // stage full: straight to the image's list
                            const int gpos = atomicAdd(&p.cand_count[img], 1);
                            if (gpos < kCandCap) p.cand_keys[(long long)img * kCandCap + gpos] = key;
                        }
                    }
                    base += __popc(mk[k]);
                }
            }
        }
        if (any) {                                                            // warp-uniform
            const float t = lane < 2 ? sigmoid_rn(v0) : expf(v0);             // lanes 0..3 hold the x, y, w, h logits
            const float sx = __shfl_sync(0xffffffffu, t, 0), sy = __shfl_sync(0xffffffffu, t, 1);
            const float ew = __shfl_sync(0xffffffffu, t, 2), eh = __shfl_sync(0xffffffffu, t, 3);
            if (lane == 0) {
                const float bx = __fmul_rn(__fadd_rn(__fsub_rn(__fmul_rn(sx, p.xyscale[s]), p.xyoff[s]), (float)col), p.stride[s]);
                const float by = __fmul_rn(__fadd_rn(__fsub_rn(__fmul_rn(sy, p.xyscale[s]), p.xyoff[s]), (float)row), p.stride[s]);
                const float bw = __fmul_rn(ew, p.anchors[(s * 3 + a) * 2 + 0]);
                const float bh = __fmul_rn(eh, p.anchors[(s * 3 + a) * 2 + 1]);
                const float hw = __fmul_rn(bw, 0.5f), hh = __fmul_rn(bh, 0.5f);   // box_wh / 2 (exact)
                float4 b;
                b.x = __fdiv_rn(__fsub_rn(bx, hw), p.img_size);                   // boxes / input_shape[0]  (custom_layers.py:284)
                b.y = __fdiv_rn(__fsub_rn(by, hh), p.img_size);
                b.z = __fdiv_rn(__fadd_rn(bx, hw), p.img_size);
                b.w = __fdiv_rn(__fadd_rn(by, hh), p.img_size);
                p.boxes[(long long)img * p.N + bn] = b;
            }
        }
    }
    __syncthreads();
    const int nk = s_nkeys < kStageKeys ? s_nkeys : kStageKeys;
    if (tid == 0 && nk > 0) s_base = atomicAdd(&p.cand_count[img], nk);
    __syncthreads();
    for (int i = tid; i < nk; i += kFilterThreads) {
        const int pos = s_base + i;
        if (pos < kCandCap) p.cand_keys[(long long)img * kCandCap + pos] = s_keys[i];
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

// Necessary condition for iou_tf(a, b) > thr when thr >= 0: a strictly positive overlap in both axes (same min / max re-ordering;
// a rounded difference is positive exactly when the true one is).
__device__ __forceinline__ bool boxes_intersect(const float4 a, const float4 b) {
    const float iy0 = fmaxf(fminf(a.x, a.z), fminf(b.x, b.z)), iy1 = fminf(fmaxf(a.x, a.z), fmaxf(b.x, b.z));
    const float ix0 = fmaxf(fminf(a.y, a.w), fminf(b.y, b.w)), ix1 = fminf(fmaxf(a.y, a.w), fmaxf(b.y, b.w));
    return iy1 > iy0 && ix1 > ix0;
}

struct NmsParams {
    const unsigned long long* cand_keys;   // [batch][kCandCap]   decode output, unordered
    const int* cand_count;                 // [batch]
    const float4* boxes;                   // [batch][N]
    int N, nc, max_boxes;
    int merge_batch;                       // images in this launch (nms_merge_kernel: several images per CTA)
    float iou_thr;
    // workspace
    unsigned long long* win_keys;          // [batch][nc][max_boxes]  per-class NMS survivors as merge keys, score descending
    int* nwin;                             // [batch][256]              (both only for images on the overflow path)
    unsigned long long* part_keys;         // [batch][kMaxParts][kMaxBoxesCap]  nms_image_kernel: each part's first max_boxes survivors
    int* ticket;                           // [batch]  parts finished (zero between launches)
    float* out_boxes;      // [batch][max_boxes][4]
    float* out_scores;     // [batch][max_boxes]
    float* out_classes;    // [batch][max_boxes]
    int* out_valid;        // [batch]
    int* out_idx;          // [batch][max_boxes]
};

// combined_non_max_suppression:
//   nms_image_kernel    `parts` CTAs per image (as many as keep batch * parts <= the SM count, at most kMaxParts), everything in
//                       shared memory, for images whose candidates fit kCandCap (all of them at the usual thresholds).  CTA k takes
//                       the classes c % parts == k: compaction of their keys -> bitonic sort (class asc, score desc, box asc) ->
//                       class segments -> one warp per class runs TF's greedy scan (iou > thr strict against already selected boxes;
//                       the pairwise tests of a short segment run lane-parallel) -> survivors become merge keys in place -> second
//                       bitonic sort -> the part's first max_boxes go to global memory; the image's last CTA (atomic ticket) merges
//                       the parts' sorted lists with one more sort and writes the first max_boxes, clipped to [0,1].
//                       (Round 2a ran this as bucket / class / merge kernels with (image, class) CTAs: 5 dependent launches and
//                       three global round trips; one CTA per image was issue-bound on 32 of the 148 SMs.)
//   nms_overflow_kernel one CTA per (image, class), only for images whose candidates did not fit kCandCap: greedy selection by
//                       repeated arg-max straight from the head tensors -- exact for ANY number of candidates
//   nms_merge_kernel    one warp per overflow image: k-way merge of the per-class survivor lists, first max_boxes, clipped
constexpr int kImgThreads = 1024;
constexpr int kMaxParts = 8;                                  // CTAs per image (kMaxParts * kMaxBoxesCap keys fit the final sort)
constexpr int kMergeThreads = 256;
constexpr size_t kImgSmemBytes = (size_t)kCandCap * (8 + 16) + kMaxBoxesCap * sizeof(unsigned short);

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

// Ascending bitonic sort of s[0..n) in shared memory, n a power of two >= 32, called by all T threads.  Exchange distances of 32
// and more go through shared memory (one barrier each); the distances below 32 of a merge step stay inside a warp and run on
// shuffles (one barrier per step): 15 + 10 barriers instead of 55 at n = 1024.
template <int T>
__device__ __forceinline__ void bitonic_sort_smem(unsigned long long* s, int n) {
    const int tid = threadIdx.x;
    for (int k = 2; k <= n; k <<= 1) {
        int j = k >> 1;
        for (; j >= 32; j >>= 1) {
            for (int t = tid; t < (n >> 1); t += T) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));      // zero bit inserted at log2(j)
                const int l = i | j;
                const unsigned long long a = s[i], b = s[l];
                if ((a > b) == ((i & k) == 0)) { s[i] = b; s[l] = a; }
            }
            __syncthreads();
        }
        for (int i = tid; i < n; i += T) {                                // whole warps: n and T are multiples of 32
            unsigned long long v = s[i];
            const bool asc = (i & k) == 0;
            for (int jj = j; jj > 0; jj >>= 1) {
                const unsigned long long o = __shfl_xor_sync(0xffffffffu, v, jj);
                const bool keep_min = ((i & jj) == 0) == asc;
                v = keep_min ? (o < v ? o : v) : (o > v ? o : v);
            }
            s[i] = v;
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kImgThreads, 1) nms_image_kernel(NmsParams p) {
    extern __shared__ __align__(16) unsigned char nms_smem[];
    unsigned long long* skey = reinterpret_cast<unsigned long long*>(nms_smem);                 // [n_pad] sort keys, later merge keys
    float4* sbox = reinterpret_cast<float4*>(nms_smem + (size_t)kCandCap * 8);                  // [cnt] boxes in sorted order
    unsigned short* selpos_all = reinterpret_cast<unsigned short*>(nms_smem + (size_t)kCandCap * 24);
    __shared__ int seg_lo[256], seg_hi[256];
    __shared__ unsigned s_sup[kImgThreads / 32][128];       // per warp: 64 candidates x 64-bit suppression mask
    __shared__ unsigned short s_pq[kImgThreads / 32][64];        // per warp: queue of intersecting pairs (i << 8 | j)
    __shared__ unsigned char long_list[256];
    __shared__ int next_class, s_cnt, s_last, n_long;
    const int img = blockIdx.y, part = blockIdx.x, K = gridDim.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned lt = (1u << lane) - 1u;
    const int cnt = p.cand_count[img];
    if (cnt > kCandCap) return;                            // nms_overflow_kernel + nms_merge_kernel own this image
    if (tid < 256) { seg_lo[tid] = 0; seg_hi[tid] = 0; }
    if (tid == 0) { next_class = part; s_cnt = 0; n_long = 0; }
    __syncthreads();
    // ---- 1. the candidates of this part's classes (class % K == part), compacted into shared memory
    for (int i0 = 0; i0 < cnt; i0 += kImgThreads) {
        const int i = i0 + tid;
        const unsigned long long key = i < cnt ? p.cand_keys[(long long)img * kCandCap + i] : 0ull;
        const bool mine = i < cnt && (int)(key >> 56) % K == part;
        const unsigned m = __ballot_sync(0xffffffffu, mine);
        if (m) {
            int base = 0;
            if (lane == 0) base = atomicAdd(&s_cnt, __popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (mine) skey[base + __popc(m & lt)] = key;
        }
    }
    __syncthreads();
    const int my = s_cnt;
    int n_pad = 32;
    while (n_pad < my) n_pad <<= 1;
    for (int i = my + tid; i < n_pad; i += kImgThreads) skey[i] = ~0ull;
    __syncthreads();
    bitonic_sort_smem<kImgThreads>(skey, n_pad);
    for (int i = tid; i < my; i += kImgThreads) {
        const unsigned long long k = skey[i];
        const int c = (int)(k >> 56);
        sbox[i] = p.boxes[(long long)img * p.N + (int)(k & 0xFFFFFFull)];
        if (i == 0 || (int)(skey[i - 1] >> 56) != c) seg_lo[c] = i;
        if (i == my - 1 || (int)(skey[i + 1] >> 56) != c) seg_hi[c] = i + 1;
    }
    __syncthreads();
    // ---- 2. TF's greedy scan per class, one warp per class, classes handed out dynamically
    unsigned short* selpos = selpos_all;                    // sequential fallback (warp 0 only)
    unsigned* sup = s_sup[warp];
    unsigned short* pq = s_pq[warp];
    while (true) {
        int c = 0;
        if (lane == 0) c = atomicAdd(&next_class, K);
        c = __shfl_sync(0xffffffffu, c, 0);
        if (c >= p.nc) break;
        const int lo = seg_lo[c], L = seg_hi[c] - lo;
        if (L == 0) continue;
        if (L > 64) {                                       // long segment: done by the whole CTA below
            if (lane == 0) long_list[atomicAdd(&n_long, 1)] = (unsigned char)c;
            continue;
        }
        // The usual case, one warp per class.  Candidate i is dropped iff an earlier SELECTED candidate j has iou > thr.  The pair
        // tests do not depend on the selection, so they run lane-parallel and leave a 64-bit mask per candidate
        // (sup[i] = { j < i : iou(i, j) > thr }); the sequential part is then a few integer instructions per candidate.
        //   pass 1: for every i, lanes j < i test whether the two boxes intersect at all (a dozen instructions; a pair that does
        //           not has iou = 0) and queue the pairs that do;
        //   pass 2: the exact TF iou (with its IEEE division) runs on full warps of queued pairs only -- in clustered detections
        //           that is ~10 % of the pairs.
        sup[lane] = 0u; sup[lane + 32] = 0u; sup[lane + 64] = 0u; sup[lane + 96] = 0u;
        const float4 bj0 = lane < L ? sbox[lo + lane] : make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 bj1 = lane + 32 < L ? sbox[lo + lane + 32] : make_float4(0.f, 0.f, 0.f, 0.f);
        __syncwarp();
        int nq = 0;
        auto drain = [&](int first, int n) {               // exact test of queued pairs [first, first + n), n <= 32
            if (lane < n) {
                const unsigned pr = pq[first + lane];
                const int i = (int)(pr >> 8), j = (int)(pr & 0xFFu);
                if (iou_tf(sbox[lo + i], sbox[lo + j]) > p.iou_thr) atomicOr(&sup[2 * i + (j >> 5)], 1u << (j & 31));   // strict >
            }
        };
        const bool all_pairs = !(p.iou_thr >= 0.f);         // a negative (or NaN) threshold: iou = 0 decides too, nothing may be skipped
        for (int i = 1; i < L; i++) {
            const float4 bi = sbox[lo + i];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                if (32 * h >= i) break;                     // warp-uniform
                const int j = 32 * h + lane;
                const float4 bj = h ? bj1 : bj0;
                const bool ok = j < i && (all_pairs || boxes_intersect(bi, bj));
                const unsigned m = __ballot_sync(0xffffffffu, ok);
                if (ok) pq[nq + __popc(m & lt)] = (unsigned short)((i << 8) | j);
                nq += __popc(m);
                __syncwarp();
                if (nq >= 32) { nq -= 32; drain(nq, 32); __syncwarp(); }
            }
        }
        drain(0, nq);
        __syncwarp();
        const unsigned S0a = sup[2 * lane], S0b = sup[2 * lane + 1], S1a = sup[2 * lane + 64], S1b = sup[2 * lane + 65];
        unsigned sel0 = 0u, sel1 = 0u;
        int nsel = 0;
        for (int j = 0; j < L && nsel < p.max_boxes; j++) {
            const unsigned sa = __shfl_sync(0xffffffffu, j < 32 ? S0a : S1a, j & 31);
            const unsigned sb = __shfl_sync(0xffffffffu, j < 32 ? S0b : S1b, j & 31);
            if (((sa & sel0) | (sb & sel1)) == 0u) { if (j < 32) sel0 |= 1u << j; else sel1 |= 1u << (j - 32); nsel++; }
        }
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int i = 32 * h + lane;
            if (i < L) {
                const unsigned long long k = skey[lo + i];
                skey[lo + i] = (((h ? sel1 : sel0) >> lane) & 1u)
                    ? merge_key(c, __uint_as_float(~(unsigned)((k >> 24) & 0xFFFFFFFFull)), (int)(k & 0xFFFFFFull)) : ~0ull;
            }
        }
        __syncwarp();
    }
    __syncthreads();
    // ---- 2b. segments of more than 64 candidates, one after the other, by the whole CTA: all threads fill the bit matrix
    // M[i] = { j < i : iou(i, j) > thr } (L^2 / 2 tests, 1024 at a time), then warp 0 walks the candidates in order with the
    // selected set as a bit vector across its lanes (one shared-memory read + one vote per candidate).  The matrix lives in the
    // unused upper half of the box array (64 KB: segments up to 704 candidates while this part holds <= kCandCap / 2 keys); longer
    // ones fall back to the sequential scan by warp 0.
    const int nl = n_long;
    unsigned* M = reinterpret_cast<unsigned*>(sbox + kCandCap / 2);
    for (int li = 0; li < nl; li++) {
        const int c = long_list[li];
        const int lo = seg_lo[c], L = seg_hi[c] - lo;
        const int W = (L + 31) >> 5;
        if (my <= kCandCap / 2 && (size_t)L * W * 4 <= (size_t)(kCandCap / 2) * sizeof(float4)) {      // L <= 704
            for (int i = tid; i < L * W; i += kImgThreads) M[i] = 0u;
            __syncthreads();
            for (int pp = tid; pp < L * L; pp += kImgThreads) {
                const int i = pp / L, j = pp - i * L;
                if (j < i && (!(p.iou_thr >= 0.f) || boxes_intersect(sbox[lo + i], sbox[lo + j])) && iou_tf(sbox[lo + i], sbox[lo + j]) > p.iou_thr)
                    atomicOr(&M[i * W + (j >> 5)], 1u << (j & 31));   // strict >
            }
            __syncthreads();
            if (warp == 0) {
                unsigned sel = 0u;                          // lane w: selected candidates 32w .. 32w+31
                int nsel = 0;
                for (int i = 0; i < L && nsel < p.max_boxes; i++) {
                    const unsigned hit = lane < W ? (M[i * W + lane] & sel) : 0u;
                    if (!__any_sync(0xffffffffu, hit != 0u)) { if (lane == (i >> 5)) sel |= 1u << (i & 31); nsel++; }
                }
                if (lane < W) M[lane] = sel;                // row 0 is no longer needed
            }
            __syncthreads();
            for (int i = tid; i < L; i += kImgThreads) {
                const unsigned long long k = skey[lo + i];
                skey[lo + i] = ((M[i >> 5] >> (i & 31)) & 1u)
                    ? merge_key(c, __uint_as_float(~(unsigned)((k >> 24) & 0xFFFFFFFFull)), (int)(k & 0xFFFFFFull)) : ~0ull;
            }
            __syncthreads();
            continue;
        }
        if (warp == 0) {
            int nsel = 0, base = 0;
            for (; base < L && nsel < p.max_boxes; base += 32) {
                const int mine = base + lane;
                const float4 mb = mine < L ? sbox[lo + mine] : make_float4(0.f, 0.f, 0.f, 0.f);
                const int cnt32 = L - base < 32 ? L - base : 32;
                unsigned keep = 0u;
                for (int j = 0; j < cnt32 && nsel < p.max_boxes; j++) {
                    float4 b;
                    b.x = __shfl_sync(0xffffffffu, mb.x, j); b.y = __shfl_sync(0xffffffffu, mb.y, j);
                    b.z = __shfl_sync(0xffffffffu, mb.z, j); b.w = __shfl_sync(0xffffffffu, mb.w, j);
                    bool supd = false;
                    for (int t = lane; t < nsel; t += 32) supd |= iou_tf(b, sbox[lo + selpos[t]]) > p.iou_thr;   // strict >
                    if (!__any_sync(0xffffffffu, supd)) {
                        if (lane == 0) selpos[nsel] = (unsigned short)(base + j);
                        keep |= 1u << j;
                        nsel++;
                        __syncwarp();
                    }
                }
                if (mine < L) {
                    const unsigned long long k = skey[lo + mine];
                    skey[lo + mine] = ((keep >> lane) & 1u)
                        ? merge_key(c, __uint_as_float(~(unsigned)((k >> 24) & 0xFFFFFFFFull)), (int)(k & 0xFFFFFFull)) : ~0ull;
                }
            }
            for (int i = base + lane; i < L; i += 32) skey[lo + i] = ~0ull;       // max_boxes reached: the rest of the class is out
        }
        __syncthreads();
    }
    // ---- 3. this part's first max_boxes survivors in merge order (score desc, class asc, box asc; ~0 = none) -> global
    bitonic_sort_smem<kImgThreads>(skey, n_pad);
    unsigned long long* mine_out = p.part_keys + ((long long)img * kMaxParts + part) * kMaxBoxesCap;
    for (int k = tid; k < p.max_boxes; k += kImgThreads) mine_out[k] = k < n_pad ? skey[k] : ~0ull;
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = atomicAdd(&p.ticket[img], 1) == K - 1;
    __syncthreads();
    if (!s_last) return;
    // ---- 4. the image's last CTA merges the K sorted lists and writes the result
    __threadfence();
    const int total = K * p.max_boxes;                      // <= kMaxParts * kMaxBoxesCap = 1024
    int n2 = 32;
    while (n2 < total) n2 <<= 1;
    for (int i = tid; i < n2; i += kImgThreads) {
        const int pk = i / p.max_boxes, k = i - pk * p.max_boxes;
        skey[i] = i < total ? __ldcg(p.part_keys + ((long long)img * kMaxParts + pk) * kMaxBoxesCap + k) : ~0ull;
    }
    if (tid == 0) p.ticket[img] = 0;                        // ready for the next launch
    __syncthreads();
    bitonic_sort_smem<kImgThreads>(skey, n2);
    const float4* boxes = p.boxes + (long long)img * p.N;
    if (tid == 0 && skey[0] == ~0ull) p.out_valid[img] = 0;
    for (int k = tid; k < p.max_boxes; k += kImgThreads) {
        const unsigned long long key = k < n2 ? skey[k] : ~0ull;
        float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
        float score = 0.f, cls = 0.f;
        int idx = -1;
        if (key != ~0ull) {
            idx = (int)(key & 0xFFFFFFull);
            cls = (float)((key >> 24) & 0xFFull);
            score = __uint_as_float(~(unsigned)(key >> 32));
            b = boxes[idx];
            b.x = fminf(fmaxf(b.x, 0.f), 1.f); b.y = fminf(fmaxf(b.y, 0.f), 1.f);   // clip_boxes=True
            b.z = fminf(fmaxf(b.z, 0.f), 1.f); b.w = fminf(fmaxf(b.w, 0.f), 1.f);
            const unsigned long long nk = (k + 1 < p.max_boxes && k + 1 < n2) ? skey[k + 1] : ~0ull;
            if (nk == ~0ull) p.out_valid[img] = k + 1;      // last valid entry among the first max_boxes
        }
        const long long o = (long long)img * p.max_boxes + k;
        reinterpret_cast<float4*>(p.out_boxes)[o] = b;
        p.out_scores[o] = score;
        p.out_classes[o] = cls;
        p.out_idx[o] = idx;
    }
}

// Exact path for images with more than kCandCap candidates (low score thresholds, e.g. mAP export at 0.001; TF's
// CombinedNonMaxSuppression has no limit): one CTA per (image, class) works straight from the head tensors.  Pass 0 marks the
// boxes whose class score passes the threshold (same arithmetic as decode_filter_kernel, so the same candidate set); then at
// most max_boxes rounds, each ONE sweep over the still-alive boxes: suppress against the box selected in the previous round
// (iou > thr strict) and find the best remaining (score desc, box asc) -- identical to TF's sorted greedy scan, without ever
// materialising or sorting the candidate list.  Launched after nms_image_kernel on every step; returns at once for images
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
    // k-way merge of the per-class survivor lists (each already score-descending = merge-key ascending) by ONE WARP per image:
    // lane l holds the heads of classes l, l + 32, ... (<= 8 lists per lane at 255 classes); max_boxes rounds of a warp-wide
    // 64-bit min by shuffles pick the output in order -- no block barriers (the 256-thread version spent 200 of them per image).
    // kMergeThreads / 32 images per CTA.
    __shared__ unsigned long long outkeys[kMergeThreads / 32][kMaxBoxesCap];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int img = blockIdx.x * (kMergeThreads / 32) + w;
    if (img >= p.merge_batch || p.cand_count[img] <= kCandCap) return;     // nms_image_kernel did this image
    const unsigned long long kNone = ~0ull;
    constexpr int PER = 8;                                    // ceil(255 / 32)
    unsigned long long head[PER], nxt[PER];
    int cur[PER], cnt[PER];
#pragma unroll
    for (int i = 0; i < PER; i++) {
        const int c = lane + 32 * i;
        cnt[i] = c < p.nc ? p.nwin[img * 256 + c] : 0;
        const unsigned long long* list = p.win_keys + ((long long)img * p.nc + (c < p.nc ? c : 0)) * p.max_boxes;
        cur[i] = 0;
        head[i] = cnt[i] > 0 ? list[0] : kNone;
        nxt[i] = cnt[i] > 1 ? list[1] : kNone;             // one element of lookahead hides the global-load latency
    }
    int nvalid = 0;
    for (int k = 0; k < p.max_boxes; k++) {
        unsigned long long m = head[0];
#pragma unroll
        for (int i = 1; i < PER; i++) m = head[i] < m ? head[i] : m;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) { const unsigned long long o = __shfl_xor_sync(0xffffffffu, m, d); m = o < m ? o : m; }
        if (m == kNone) break;                                // warp-uniform: every list is exhausted
        if (lane == 0) outkeys[w][k] = m;
        nvalid = k + 1;
#pragma unroll
        for (int i = 0; i < PER; i++)
            if (head[i] == m) {                               // keys are unique: exactly one owner
                const int c = lane + 32 * i;
                const unsigned long long* list = p.win_keys + ((long long)img * p.nc + c) * p.max_boxes;
                cur[i]++;
                head[i] = nxt[i];
                nxt[i] = cur[i] + 1 < cnt[i] ? list[cur[i] + 1] : kNone;
            }
    }
    __syncwarp();
    const float4* boxes = p.boxes + (long long)img * p.N;
    if (lane == 0) p.out_valid[img] = nvalid;
    for (int k = lane; k < p.max_boxes; k += 32) {
        float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
        float score = 0.f, cls = 0.f;
        int idx = -1;
        if (k < nvalid) {
            const unsigned long long key = outkeys[w][k];
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
