// conv 0 (3 -> 32, 3x3 'same', leaky) on the tensor cores, straight from the caller's float32 NHWC image.
//
// K = 27 padded to 32 = one 64-byte K-major row per pixel, so the whole layer is M=128 x N=32 x K=32 tiles:
// each thread gathers the 27 inputs of ITS pixel from global memory (im2col in registers, L1-served: neighbouring
// pixels share 2/3 of their window), converts to fp16 and writes the row into shared memory in the 64B-swizzled
// K-major layout the UMMA descriptor names (chunk j of row r lands at r*64 + ((j ^ ((r >> 1) & 3)) << 4)).  Two
// tcgen05.mma K-steps per tile, accumulator in TMEM, epilogue = bias + leaky -> fp16, 64 B per pixel, consecutive
// lanes write consecutive pixels (2 KB coalesced per warp).  The layer is then bound by writing its own output.
#pragma once
#include "conv_tc.cuh"

namespace y4 {

struct Conv0TcParams {
    const float* img;        // (N, S, S, 3) float32
    const __half* w;         // [32][32] fp16: row = cout, K index (kh*3+kw)*3 + c, zero padded 27 -> 32
    const float* bias;       // [>=32]
    __half* out;             // padded-flat (N, S+2, S+2, 32)
    int N, S, tiles_per_row; // tiles of 128 pixels along x
    int num_tiles;
};

constexpr int kC0Threads = 128;

__global__ void __launch_bounds__(kC0Threads) conv0_tc_kernel(const Conv0TcParams p) {
    __shared__ __align__(1024) unsigned char sA[128 * 64];     // 128 pixels x 32 fp16, SW64
    __shared__ __align__(1024) unsigned char sB[32 * 64];      // 32 couts  x 32 fp16, SW64
    __shared__ __align__(8) unsigned long long mma_bar;
    __shared__ uint32_t tmem_slot;
    __shared__ float sbias[32];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t a_addr = smem_u32(sA), b_addr = smem_u32(sB), bar = smem_u32(&mma_bar);
    constexpr uint32_t IDESC = make_idesc(128, 32);

    if (tid == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), 32);
    if (tid < 32) {
        sbias[tid] = p.bias[tid];
        const uint4* src = reinterpret_cast<const uint4*>(p.w + tid * 32);
#pragma unroll
        for (int j = 0; j < 4; j++)
            *reinterpret_cast<uint4*>(sB + tid * 64 + ((j ^ ((tid >> 1) & 3)) << 4)) = src[j];
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    const int S = p.S, Sp = S + 2;
    uint32_t phase = 0;

    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int tx = tile % p.tiles_per_row;
        const int row = tile / p.tiles_per_row;          // n * S + y
        const int y = row % S, n = row / S;
        const int x = tx * 128 + tid;
        // ---- im2col in registers: k = (kh*3 + kw)*3 + c
        float v[32];
#pragma unroll
        for (int k = 27; k < 32; k++) v[k] = 0.f;
#pragma unroll
        for (int kh = 0; kh < 3; kh++) {
            const int yy = y + kh - 1;
            const bool yok = yy >= 0 && yy < S && x < S;
            const float* rp = p.img + ((long long)(n * S + (yok ? yy : 0)) * S) * 3;
#pragma unroll
            for (int kw = 0; kw < 3; kw++) {
                const int xx = x + kw - 1;
                const bool ok = yok && xx >= 0 && xx < S;
#pragma unroll
                for (int c = 0; c < 3; c++) v[(kh * 3 + kw) * 3 + c] = ok ? __ldg(rp + xx * 3 + c) : 0.f;
            }
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            __half2 h0 = __floats2half2_rn(v[8 * j + 0], v[8 * j + 1]);
            __half2 h1 = __floats2half2_rn(v[8 * j + 2], v[8 * j + 3]);
            __half2 h2 = __floats2half2_rn(v[8 * j + 4], v[8 * j + 5]);
            __half2 h3 = __floats2half2_rn(v[8 * j + 6], v[8 * j + 7]);
            uint4 u;
            u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
            u.z = *reinterpret_cast<uint32_t*>(&h2); u.w = *reinterpret_cast<uint32_t*>(&h3);
            *reinterpret_cast<uint4*>(sA + tid * 64 + ((j ^ ((tid >> 1) & 3)) << 4)) = u;
        }
        // generic-proxy smem writes -> visible to the async proxy the tensor core reads through
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            const uint64_t da = make_smem_desc<64>(a_addr), db = make_smem_desc<64>(b_addr);
            umma_f16(tmem_base, da, db, IDESC, 0u);
            umma_f16(tmem_base, da + 2ull, db + 2ull, IDESC, 1u);
            umma_commit(bar);
        }
        mbar_wait(bar, phase);
        phase ^= 1u;
        tc_fence_after();
        uint32_t acc[32];
        tmem_ld32_issue(tmem_base + ((uint32_t)(warp * 32) << 16), acc);
        tmem_ld_wait(acc);
        if (x < S) {
            uint4 o[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                __half2 h[4];
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    float u0 = __uint_as_float(acc[8 * q + 2 * t]) + sbias[8 * q + 2 * t];
                    float u1 = __uint_as_float(acc[8 * q + 2 * t + 1]) + sbias[8 * q + 2 * t + 1];
                    h[t] = __floats2half2_rn(fmaxf(u0, 0.1f * u0), fmaxf(u1, 0.1f * u1));     // leaky (custom_layers.py:101)
                }
                o[q].x = *reinterpret_cast<uint32_t*>(&h[0]); o[q].y = *reinterpret_cast<uint32_t*>(&h[1]);
                o[q].z = *reinterpret_cast<uint32_t*>(&h[2]); o[q].w = *reinterpret_cast<uint32_t*>(&h[3]);
            }
            uint4* op = reinterpret_cast<uint4*>(p.out + (((long long)n * Sp + y + 1) * Sp + x + 1) * 32);
#pragma unroll
            for (int q = 0; q < 4; q++) op[q] = o[q];
        }
        tc_fence_before();
        __syncthreads();                                   // TMEM + sA are reused by the next tile
    }
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 32); }
}

}  // namespace y4
