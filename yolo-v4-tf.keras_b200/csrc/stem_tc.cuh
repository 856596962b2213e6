// conv 0 + conv 1 as ONE kernel (custom_layers.py:101-102: 3x3 'same' 3->32 leaky, then ZeroPadding2D(((1,0),(1,0))) +
// 3x3 stride-2 'valid' 32->64 leaky): image in, conv-1 output out.  conv 0's 762 MB output (batch 32, 608^2) is consumed
// by conv 1 only; run separately the two layers cost 0.69 ms against an HBM floor of 0.08 ms for image-in / c1-out.
//
// One tile = 16 x 8 conv-1 outputs of one image = 128 rows of the conv-1 GEMM.  It needs the 33 x 17 conv-0 outputs
// (2*oh0-1 .., 2*ow0-1 ..), computed here (9.6 % recompute on the tile borders) and kept in shared memory:
//   G0  the 35 x 19 image pixels under the patch are staged in shared memory as fp16 (c0, c1, c2, 0), zero outside the image;
//   G   every thread builds conv-0 im2col rows (27 halfs + 5 zeros = one 64 B K-major row, K order as conv0_tc.cuh: 9 LDS.64 +
//       8 PRMT per row instead of 27 predicated global loads) for the 561 patch pixels, ordered PLANE-MAJOR: the patch is stored as its four (row parity, column parity) planes, 17x9, 17x8,
//       16x9, 16x8 pixels, because tap (kh, kw) of the stride-2 conv reads plane (kh&1, kw&1) at offset (kh>>1, kw>>1):
//       8 consecutive outputs of a tile row are 8 CONSECUTIVE plane pixels = one 8-row core group of a K-major UMMA operand,
//       and the 16 tile rows are 16 groups one plane pitch apart (the descriptor's stride-byte-offset);
//   M0  5 x (M=128, N=32, K=32) MMAs -> TMEM columns [0, 160);
//   E0  TMEM -> bias + leaky -> fp16, written back IN PLACE over the im2col rows (row m of conv 0's GEMM is pixel m of the
//       plane-major patch); pixels in conv 1's top / left zero padding (y or x = -1) become zero rows;
//   M1  9 taps x 2 K-steps of (M=128, N=64, K=16) against the resident conv-1 weights (36 KB) -> TMEM columns [192, 256);
//   E1  TMEM -> bias + leaky -> fp16 -> 128 B-swizzled slab -> one 4-D TMA box store {64 ch, 8, 16, 1} into c1.
// Same K order and per-element arithmetic as conv0_tc_kernel followed by conv 1's tcgen05 plans (tap outer, channel inner), so
// the result is bit-identical to the two-kernel path (tests/test_gpu_determinism.py).  2 CTAs of 256 threads per SM overlap
// each other's phases.
#pragma once
#include "conv_tc.cuh"

namespace y4 {

struct StemParams {
    const float* img;        // (N, S, S, 3) float32
    const __half* w0;        // conv 0: [32][32] fp16, row = cout, K index (kh*3+kw)*3 + c zero padded 27 -> 32
    const float* bias0;      // [32]
    const float* bias1;      // [64]
    CUtensorMap tmW1;        // conv 1 weights [64][288] fp16 (K index tap*32 + c), box {32, 64}, SWIZZLE_64B
    CUtensorMap tmOut;       // c1 padded-flat (N, H1+2, H1+2, 64) as 4-D {64, W1p, H1p, N}, box {64, 8, 16, 1}, SWIZZLE_128B
    int N, S, H1;            // H1 = S / 2
    int tiles_w, tiles_h, num_tiles;
};

constexpr int kStemThreads = 256;
constexpr int kStemRows = 561;                      // 33 x 17 conv-0 pixels per tile
constexpr int kStemRowsPad = 640;                   // 5 MMA tiles of 128 rows
constexpr uint32_t kStemA = 0;                      // [640][64 B]  im2col rows, then the conv-0 outputs (in place)
constexpr uint32_t kStemSlab = kStemRowsPad * 64;   // [128][128 B] conv-1 output tile, SWIZZLE_128B
constexpr uint32_t kStemB1 = kStemSlab + 128 * 128; // 9 x [64][64 B]  conv-1 weights, SWIZZLE_64B
constexpr uint32_t kStemB0 = kStemB1 + 9 * 64 * 64; // [32][64 B]      conv-0 weights, SWIZZLE_64B
constexpr uint32_t kStemBars = kStemB0 + 32 * 64;   // 3 mbarriers + tmem slot, then the biases
constexpr uint32_t kStemPatch = kStemBars + 64 + 96 * 4;         // [35][19] image pixels as 4 x fp16 (c0, c1, c2, 0), zero outside the image
constexpr uint32_t kStemPatchBytes = 35 * 19 * 8;               // two of them: this tile's and the next one's
constexpr uint32_t kStemSmem = kStemPatch + 2 * kStemPatchBytes + 1024;  // + alignment slack

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
        ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
// K-major SWIZZLE_64B operand whose 8-row groups are `sbo_bytes` apart (a plane pitch instead of the dense 512 B)
__device__ __forceinline__ uint64_t make_smem_desc_sw64_sbo(uint32_t saddr, uint32_t sbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46) | (4ull << 61);
}

// plane-major patch index m -> (patch row i, patch column j) of the 33 x 17 conv-0 patch
__device__ __forceinline__ void stem_row_to_patch(int m, int& i, int& j) {
    int r, pw_, ph, pc;
    if (m < 153) { r = m; pw_ = 9; ph = 0; pc = 0; }
    else if (m < 289) { r = m - 153; pw_ = 8; ph = 0; pc = 1; }
    else if (m < 433) { r = m - 289; pw_ = 9; ph = 1; pc = 0; }
    else { r = m - 433; pw_ = 8; ph = 1; pc = 1; }
    const int prow = r / pw_;
    i = 2 * prow + ph;
    j = 2 * (r - prow * pw_) + pc;
}

// image pixels idx = tid, tid + 256, tid + 512 of the 35 x 19 patch of `tile` -> registers (zero outside the image)
__device__ __forceinline__ void stem_patch_load(const StemParams& p, int tile, int tid, float (&v)[3][3]) {
    const int twi = tile % p.tiles_w;
    const int t2 = tile / p.tiles_w;
    const int thi = t2 % p.tiles_h, n = t2 / p.tiles_h;
    const int y0 = 32 * thi - 1, x0 = 16 * twi - 1;
    const float* img = p.img + (long long)n * p.S * p.S * 3;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const int idx = tid + k * kStemThreads;
        const int r = idx / 19, c = idx - r * 19;
        const int iy = y0 - 1 + r, ix = x0 - 1 + c;
        const bool ok = idx < 35 * 19 && iy >= 0 && iy < p.S && ix >= 0 && ix < p.S;
        const float* px = img + ((long long)(ok ? iy : 0) * p.S + (ok ? ix : 0)) * 3;
        v[k][0] = ok ? __ldg(px) : 0.f; v[k][1] = ok ? __ldg(px + 1) : 0.f; v[k][2] = ok ? __ldg(px + 2) : 0.f;
    }
}
__device__ __forceinline__ void stem_patch_store(unsigned char* buf, int tid, const float (&v)[3][3]) {
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const int idx = tid + k * kStemThreads;
        if (idx < 35 * 19) {
            __half2 h01 = __floats2half2_rn(v[k][0], v[k][1]);
            __half2 h2z = __floats2half2_rn(v[k][2], 0.f);
            reinterpret_cast<uint2*>(buf)[idx] = make_uint2(*reinterpret_cast<uint32_t*>(&h01), *reinterpret_cast<uint32_t*>(&h2z));
        }
    }
}

__global__ void __launch_bounds__(kStemThreads, 2) stem_tc_kernel(const __grid_constant__ StemParams p) {
    extern __shared__ unsigned char stem_smem[];
    const uint32_t raw = smem_u32(stem_smem);
    const uint32_t base = (raw + 1023u) & ~1023u;
    unsigned char* sm = stem_smem + (base - raw);
    const uint32_t a_addr = base + kStemA, slab_addr = base + kStemSlab, b1_addr = base + kStemB1, b0_addr = base + kStemB0;
    const uint32_t bar0 = base + kStemBars, bar1 = bar0 + 8, bar_w = bar0 + 16, tmem_slot = bar0 + 24;
    float* sbias0 = reinterpret_cast<float*>(sm + kStemBars + 64);
    float* sbias1 = sbias0 + 32;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr uint32_t IDESC0 = make_idesc(128, 32), IDESC1 = make_idesc(128, 64);
    pdl_launch_dependents();
    if (tid == 0) {
        tma_prefetch_desc(&p.tmW1); tma_prefetch_desc(&p.tmOut);
        mbar_init(bar0, 1); mbar_init(bar1, 1); mbar_init(bar_w, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(bar_w, 9u * 64u * 64u);            // weights are written at load time, never by a preceding kernel
        for (int t = 0; t < 9; t++) tma_load_2d(b1_addr + (uint32_t)t * 4096u, &p.tmW1, bar_w, t * 32, 0);
    }
    if (warp == 0) tmem_alloc(tmem_slot, 256);
    if (tid < 32) {
        sbias0[tid] = p.bias0[tid];
        const uint4* src = reinterpret_cast<const uint4*>(p.w0 + tid * 32);
#pragma unroll
        for (int j = 0; j < 4; j++) *reinterpret_cast<uint4*>(sm + kStemB0 + tid * 64 + ((j ^ ((tid >> 1) & 3)) << 4)) = src[j];
    }
    if (tid >= 64 && tid < 128) sbias1[tid - 64] = p.bias1[tid - 64];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    pdl_wait();                                            // the image may be written by the preceding kernel (preprocess)

    const int S = p.S;
    const int q = warp & 3, set = warp >> 2;               // TMEM lane quarter of this warp; which half of the work it takes
    uint32_t phase = 0, pbuf = 0;
    bool w_ready = false;
    if ((int)blockIdx.x < p.num_tiles) {                   // prologue: the first tile's patch
        float pre0[3][3];
        stem_patch_load(p, (int)blockIdx.x, tid, pre0);
        stem_patch_store(sm + kStemPatch, tid, pre0);
    }
    __syncthreads();
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, pbuf ^= 1u) {
        const int twi = tile % p.tiles_w;
        const int t2 = tile / p.tiles_w;
        const int thi = t2 % p.tiles_h, n = t2 / p.tiles_h;
        const int oh0 = thi * 16, ow0 = twi * 8;
        const int y0 = 2 * oh0 - 1, x0 = 2 * ow0 - 1;     // conv-0 pixel of patch position (0, 0)

        // ---- G0 (software-pipelined): the 35 x 19 image pixels under the NEXT tile's patch are requested here (global loads into
        // registers, nothing waits on them) and written to the other patch buffer at the end of this tile, fp16 (c0, c1, c2, 0),
        // zero outside the image ('same'); this tile's patch was staged during the previous one (prologue for the first).
        const uint2* spatch = reinterpret_cast<const uint2*>(sm + kStemPatch + (pbuf ? kStemPatchBytes : 0u));
        float pre[3][3];
        const int ntile = tile + (int)gridDim.x;
        if (ntile < p.num_tiles) stem_patch_load(p, ntile, tid, pre);
        // ---- G: conv-0 im2col rows, plane-major: K index (kh*3 + kw)*3 + c (as conv0_tc_kernel), 27 halfs packed from 9 pixels
        for (int m = tid; m < kStemRows; m += kStemThreads) {
            int i, j;
            stem_row_to_patch(m, i, j);
            uint32_t a[9], b[9];
#pragma unroll
            for (int kh = 0; kh < 3; kh++)
#pragma unroll
                for (int kw = 0; kw < 3; kw++) {
                    const uint2 v = spatch[(i + kh) * 19 + j + kw];
                    a[kh * 3 + kw] = v.x; b[kh * 3 + kw] = v.y;
                }
            uint32_t w[16];
#pragma unroll
            for (int t = 0; t < 4; t++) {
                w[3 * t] = a[2 * t];
                w[3 * t + 1] = __byte_perm(b[2 * t], a[2 * t + 1], 0x5410);        // (c2 of pixel 2t, c0 of pixel 2t+1)
                w[3 * t + 2] = __byte_perm(a[2 * t + 1], b[2 * t + 1], 0x5432);    // (c1, c2 of pixel 2t+1)
            }
            w[12] = a[8]; w[13] = b[8]; w[14] = 0u; w[15] = 0u;                     // b[8] = (c2, 0): k = 26, 27
#pragma unroll
            for (int c4 = 0; c4 < 4; c4++)
                *reinterpret_cast<uint4*>(sm + kStemA + m * 64 + ((c4 ^ ((m >> 1) & 3)) << 4)) = make_uint4(w[4 * c4], w[4 * c4 + 1], w[4 * c4 + 2], w[4 * c4 + 3]);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        tc_fence_before();
        __syncthreads();
        // ---- M0: conv 0, five 128-row tiles
        if (warp == 0) {                                   // whole warp, elect-predicated issue (conv_tc.cuh elect_one)
            tc_fence_after();
            const uint64_t db = make_smem_desc<64>(b0_addr);
#pragma unroll
            for (int t = 0; t < 5; t++) {
                const uint64_t da = make_smem_desc<64>(a_addr + (uint32_t)t * 8192u);
                umma_f16_elect(tmem_base + (uint32_t)(32 * t), da, db, IDESC0, 0u);
                umma_f16_elect(tmem_base + (uint32_t)(32 * t), da + 2ull, db + 2ull, IDESC0, 1u);
            }
            umma_commit_elect(bar0);
        }
        mbar_wait(bar0, phase);
        tc_fence_after();
        // ---- E0: bias + leaky -> fp16, in place; conv 1's zero padding (y = -1 or x = -1) -> zero rows
        for (int t = set; t < 5; t += 2) {
            const int m = t * 128 + q * 32 + lane;
            uint32_t acc[32];
            tmem_ld32_issue(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(32 * t), acc);
            tmem_ld_wait(acc);
            if (m < kStemRows) {
                int i, j;
                stem_row_to_patch(m, i, j);
                const bool pad = (y0 + i) < 0 || (x0 + j) < 0;
#pragma unroll
                for (int c4 = 0; c4 < 4; c4++) {
                    __half2 h[4];
#pragma unroll
                    for (int t4 = 0; t4 < 4; t4++) {                                               // packed fp32 (FADD2 / FMUL2): same bits as the scalar form
                        const unsigned long long u = add2(pk2(__uint_as_float(acc[8 * c4 + 2 * t4]), __uint_as_float(acc[8 * c4 + 2 * t4 + 1])),
                                                          *reinterpret_cast<const unsigned long long*>(sbias0 + 8 * c4 + 2 * t4));
                        float u0, u1, y0, y1;
                        upk2(u, u0, u1); upk2(mul2(u, pk2(0.1f, 0.1f)), y0, y1);
                        h[t4] = __floats2half2_rn(fmaxf(u0, y0), fmaxf(u1, y1));                   // leaky (custom_layers.py:101)
                    }
                    uint4 o;
                    o.x = *reinterpret_cast<uint32_t*>(&h[0]); o.y = *reinterpret_cast<uint32_t*>(&h[1]);
                    o.z = *reinterpret_cast<uint32_t*>(&h[2]); o.w = *reinterpret_cast<uint32_t*>(&h[3]);
                    if (pad) o = make_uint4(0u, 0u, 0u, 0u);
                    *reinterpret_cast<uint4*>(sm + kStemA + m * 64 + ((c4 ^ ((m >> 1) & 3)) << 4)) = o;
                }
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        tc_fence_before();
        __syncthreads();
        // ---- M1: conv 1, tap outer / channel inner; A = plane (kh&1, kw&1) shifted by (kh>>1, kw>>1), 16 groups one plane pitch apart
        if (warp == 0) {
            if (!w_ready) { mbar_wait(bar_w, 0u); w_ready = true; }
            tc_fence_after();
#pragma unroll
            for (int tap = 0; tap < 9; tap++) {
                const int kh = tap / 3, kw = tap - kh * 3;
                const int pl = (kh & 1) * 2 + (kw & 1);
                const int pbase = pl == 0 ? 0 : (pl == 1 ? 153 : (pl == 2 ? 289 : 433));
                const int pitch = (kw & 1) ? 8 : 9;
                const uint32_t row0 = (uint32_t)(pbase + (kh >> 1) * pitch + (kw >> 1));
                const uint64_t da = make_smem_desc_sw64_sbo(a_addr + row0 * 64u, (uint32_t)pitch * 64u);
                const uint64_t db = make_smem_desc<64>(b1_addr + (uint32_t)tap * 4096u);
                umma_f16_elect(tmem_base + 192u, da, db, IDESC1, tap ? 1u : 0u);
                umma_f16_elect(tmem_base + 192u, da + 2ull, db + 2ull, IDESC1, 1u);
            }
            umma_commit_elect(bar1);
            if (lane == 0) bulk_wait_read<0>();            // the previous tile's TMA store (issued by thread 0) has finished reading the slab
            __syncwarp();
        }
        mbar_wait(bar1, phase);
        phase ^= 1u;
        tc_fence_after();
        // ---- E1: bias + leaky -> fp16 -> slab (row m = th*8 + tw, 128 B rows, SWIZZLE_128B); warps 0-3 / 4-7: channels 0-31 / 32-63
        {
            const int m = q * 32 + lane;
            uint32_t acc[32];
            tmem_ld32_issue(tmem_base + ((uint32_t)(q * 32) << 16) + 192u + (uint32_t)(32 * set), acc);
            tmem_ld_wait(acc);
            __syncthreads();                               // thread 0 is past bulk_wait_read: the slab may be overwritten
#pragma unroll
            for (int c4 = 0; c4 < 4; c4++) {
                __half2 h[4];
#pragma unroll
                for (int t4 = 0; t4 < 4; t4++) {
                    const unsigned long long u = add2(pk2(__uint_as_float(acc[8 * c4 + 2 * t4]), __uint_as_float(acc[8 * c4 + 2 * t4 + 1])),
                                                      *reinterpret_cast<const unsigned long long*>(sbias1 + 32 * set + 8 * c4 + 2 * t4));
                    float u0, u1, y0, y1;
                    upk2(u, u0, u1); upk2(mul2(u, pk2(0.1f, 0.1f)), y0, y1);
                    h[t4] = __floats2half2_rn(fmaxf(u0, y0), fmaxf(u1, y1));                       // leaky (custom_layers.py:102)
                }
                uint4 o;
                o.x = *reinterpret_cast<uint32_t*>(&h[0]); o.y = *reinterpret_cast<uint32_t*>(&h[1]);
                o.z = *reinterpret_cast<uint32_t*>(&h[2]); o.w = *reinterpret_cast<uint32_t*>(&h[3]);
                const int chunk = 4 * set + c4;
                *reinterpret_cast<uint4*>(sm + kStemSlab + m * 128 + ((chunk ^ (m & 7)) << 4)) = o;
            }
        }
        if (ntile < p.num_tiles) stem_patch_store(sm + kStemPatch + (pbuf ? 0u : kStemPatchBytes), tid, pre);   // the next tile's patch
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        tc_fence_before();
        __syncthreads();                                   // slab + next patch complete; TMEM and the A region are free for the next tile
        if (tid == 0) { tma_store_4d(&p.tmOut, slab_addr, 0, ow0 + 1, oh0 + 1, n); bulk_commit(); }
    }
    if (tid == 0) bulk_wait_all();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 256); }
}

struct StemPlan {
    StemParams p;
    bool ok = false;
};

// c1 must be a plain 64-channel buffer (ld 64, channel offset 0); S a multiple of 32 (16 x 8 tiles of the S/2 map)
inline bool stem_plan(StemPlan* pl, const float* bias0, const __half* w0, const float* bias1, const __half* w1, void* c1_buf, int c1_ld,
                      int c1_choff, int S, int max_batch, std::string* err) {
    pl->ok = false;
    if (c1_ld != 64 || c1_choff != 0 || S % 32 != 0) return false;
    StemParams& p = pl->p;
    memset(&p, 0, sizeof(p));
    p.w0 = w0; p.bias0 = bias0; p.bias1 = bias1; p.S = S; p.H1 = S / 2;
    p.tiles_w = p.H1 / 8; p.tiles_h = p.H1 / 16;
    {
        cuuint64_t dims[2] = {288, 64};
        cuuint64_t str[1] = {288 * 2};
        cuuint32_t box[2] = {32, 64};
        if (!encode_map(&p.tmW1, const_cast<__half*>(w1), 2, dims, str, box, 64, err)) return false;
    }
    {
        const cuuint64_t W1p = (cuuint64_t)p.H1 + 2;
        cuuint64_t dims[4] = {64, W1p, W1p, (cuuint64_t)max_batch};
        cuuint64_t str[3] = {64 * 2, W1p * 64 * 2, W1p * W1p * 64 * 2};
        cuuint32_t box[4] = {64, 8, 16, 1};
        if (!encode_map(&p.tmOut, c1_buf, 4, dims, str, box, 128, err)) return false;
    }
    pl->ok = true;
    return true;
}

inline int stem_launch(const StemPlan& pl_in, const float* img, int batch, cudaStream_t st) {
    StemPlan pl = pl_in;
    pl.p.img = img; pl.p.N = batch;
    pl.p.num_tiles = batch * pl.p.tiles_h * pl.p.tiles_w;
    static DeviceOnce once;
    if (once.first_use()) {
        if (cudaFuncSetAttribute(stem_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kStemSmem) != cudaSuccess) { once.forget(); return -1; }
    }
    const int max_ctas = sm_count() * 2;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(pl.p.num_tiles < max_ctas ? pl.p.num_tiles : max_ctas));
    cfg.blockDim = dim3(kStemThreads); cfg.dynamicSmemBytes = kStemSmem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, stem_tc_kernel, pl.p) == cudaSuccess ? 0 : -1;
}

}  // namespace y4
