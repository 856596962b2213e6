// CUDA-core kernels: fp32 implicit-GEMM conv (parity mode / debug reference for the tcgen05 path),
// SPP max-pool, synthetic input fill, L2 flush.
//
// Activation layout ("padded-flat"): every activation tensor is stored NHWC with a 1-pixel zero halo,
// rows = N*(H+2)*(W+2), row r = (n*(H+2) + h+1)*(W+2) + (w+1), `ld` channels per row (a tensor may be a
// channel slice [ch_off, ch_off+C) of a wider concat buffer).  The halo is zeroed once at allocation and
// never written, so a 3x3 'same' conv is a plain GEMM whose A rows are the output row shifted by
// (kh-1)*(W+2) + (kw-1), and ZeroPadding2D(((1,0),(1,0))) of the stride-2 convs (custom_layers.py:10)
// is the halo itself.
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace y4 {

enum { ACT_LINEAR = 0, ACT_LEAKY = 1, ACT_MISH = 2 };

struct SimtConvParams {
    const void* in;            // padded-flat activations (TIn) or raw image (float, RAW=true)
    int in_ld, in_choff;       // channels per row of the underlying buffer, first channel of the view
    int in_Hp, in_Wp;          // padded dims of the input (RAW: unpadded H, W)
    const float* w;            // [K][cout_pad] fp32, K index = (kh*k + kw)*cin + c   (HWIO, utils.py:42)
    const float* bias;         // [cout_pad] folded BN bias / head bias
    int K, cin, ksize, stride, cout, cout_pad;
    void* out;                 // TOut (or float when out_f32)
    int out_ld, out_choff;
    int N, OH, OW;             // logical output dims; M = N*(OH+2)*(OW+2) rows are swept, halo rows skipped
    const void* res;           // optional residual (same dims as out), added AFTER the activation (custom_layers.py:44)
    int res_ld, res_choff;
    int act;
    int upsample;              // 1: write each output pixel to the 2x2 block of a (2*OH, 2*OW) destination (UpSampling2D)
    int out_f32;               // heads: destination is float regardless of TOut
    long long M;
};

__device__ __forceinline__ float ld_as_float(const float* p) { return *p; }
__device__ __forceinline__ float ld_as_float(const __half* p) { return __half2float(*p); }
__device__ __forceinline__ void st_from_float(float* p, float v) { *p = v; }
__device__ __forceinline__ void st_from_float(__half* p, float v) { *p = __float2half_rn(v); }

// Reference activation semantics (custom_layers.py:6-7,27-30), fp32, accurate libm functions.
__device__ __forceinline__ float act_exact(float x, int act) {
    if (act == ACT_MISH) {
        float sp = (x > 20.f) ? x : log1pf(expf(x));
        return x * tanhf(sp);
    }
    if (act == ACT_LEAKY) return x > 0.f ? x : 0.1f * x;
    return x;
}

template <typename TIn, typename TOut, bool RAW>
__global__ void __launch_bounds__(256) conv_simt_kernel(SimtConvParams p) {
    __shared__ float As[16][64 + 4];
    __shared__ float Bs[16][64];
    const int tid = threadIdx.x;
    const long long m0 = (long long)blockIdx.x * 64;
    const int n0 = blockIdx.y * 64;
    const int OHp = p.OH + 2, OWp = p.OW + 2;

    // A-load role: one output row per 4 threads
    const int la_r = tid >> 2, la_k = (tid & 3) * 4;
    const long long r = m0 + la_r;
    int rn = 0, rhp = 0, rwp = 0;
    bool interior = false;
    if (r < p.M) {
        rwp = (int)(r % OWp);
        long long t = r / OWp;
        rhp = (int)(t % OHp);
        rn = (int)(t / OHp);
        interior = (rhp >= 1 && rhp <= p.OH && rwp >= 1 && rwp <= p.OW);
    }
    const int pad = p.ksize / 2;
    const TIn* in = reinterpret_cast<const TIn*>(p.in);

    const int ty = tid >> 4, tx = tid & 15;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < p.K; k0 += 16) {
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int kk = k0 + la_k + j;
            float v = 0.f;
            if (interior && kk < p.K) {
                int tap = kk / p.cin, c = kk - tap * p.cin;
                int kh = tap / p.ksize, kw = tap - kh * p.ksize;
                if (RAW) {
                    int ih = (rhp - 1) + kh - pad, iw = (rwp - 1) + kw - pad;   // 'same', stride 1, unpadded source
                    if (ih >= 0 && ih < p.in_Hp && iw >= 0 && iw < p.in_Wp)
                        v = ld_as_float(in + (((long long)rn * p.in_Hp + ih) * p.in_Wp + iw) * p.in_ld + p.in_choff + c);
                } else {
                    int ihp = (rhp - 1) * p.stride + kh - pad + 1;
                    int iwp = (rwp - 1) * p.stride + kw - pad + 1;
                    v = ld_as_float(in + (((long long)rn * p.in_Hp + ihp) * p.in_Wp + iwp) * p.in_ld + p.in_choff + c);
                }
            }
            As[la_k + j][la_r] = v;
        }
        {
            int kk = k0 + (tid >> 4);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                int col = n0 + (tid & 15) * 4 + j;
                Bs[tid >> 4][(tid & 15) * 4 + j] = (kk < p.K && col < p.cout_pad) ? p.w[(long long)kk * p.cout_pad + col] : 0.f;
            }
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; kk++) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; i++) a[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; j++) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }

    // epilogue: bias -> activation -> (+ residual) -> store
#pragma unroll
    for (int i = 0; i < 4; i++) {
        long long rr = m0 + ty * 4 + i;
        if (rr >= p.M) continue;
        int wp = (int)(rr % OWp);
        long long t = rr / OWp;
        int hp = (int)(t % OHp);
        int n = (int)(t / OHp);
        if (!(hp >= 1 && hp <= p.OH && wp >= 1 && wp <= p.OW)) continue;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int col = n0 + tx * 4 + j;
            if (col >= p.cout) continue;
            float v = act_exact(acc[i][j] + p.bias[col], p.act);
            if (p.res) v += ld_as_float(reinterpret_cast<const TOut*>(p.res) + rr * p.res_ld + p.res_choff + col);
            if (p.upsample) {
                int DHp = 2 * p.OH + 2, DWp = 2 * p.OW + 2;
#pragma unroll
                for (int dy = 0; dy < 2; dy++)
#pragma unroll
                    for (int dx = 0; dx < 2; dx++) {
                        long long dr = ((long long)n * DHp + 2 * (hp - 1) + 1 + dy) * DWp + 2 * (wp - 1) + 1 + dx;
                        st_from_float(reinterpret_cast<TOut*>(p.out) + dr * p.out_ld + p.out_choff + col, v);
                    }
            } else if (p.out_f32) {
                reinterpret_cast<float*>(p.out)[rr * p.out_ld + p.out_choff + col] = v;
            } else {
                st_from_float(reinterpret_cast<TOut*>(p.out) + rr * p.out_ld + p.out_choff + col, v);
            }
        }
    }
}

// conv 0 (cin = 3, 3x3 'same', leaky) straight from the caller's float32 NHWC image: direct convolution on CUDA
// cores, fp32 accumulate, fp16 padded-flat output.  A block computes a 32 x 16 pixel tile x 32 channels; each thread
// owns two pixels (rows y and y+8) so every broadcast weight read feeds two FMAs.  Output rows are 64 B and
// consecutive lanes write consecutive pixels: fully coalesced 2 KB per warp store.
struct Conv0Params {
    const float* img;      // (N, S, S, 3)
    const float* w;        // [27][cout_pad] folded, K index = (kh*3+kw)*3 + c
    const float* bias;     // [cout_pad]
    __half* out;           // padded-flat (N, S+2, S+2, 32)
    __half* out_lo;        // split precision: low-order plane (nullptr otherwise)
    int N, S, cout_pad;
};

__global__ void __launch_bounds__(256) conv0_direct_kernel(Conv0Params p) {
    constexpr int TW = 32, TH = 16;
    __shared__ float patch[TH + 2][(TW + 2) * 3];
    __shared__ __align__(16) float ws[27][32];
    __shared__ float bs[32];
    const int tid = threadIdx.x;
    const int n = blockIdx.z, y0 = blockIdx.y * TH, x0 = blockIdx.x * TW;
    for (int i = tid; i < 27 * 32; i += 256) ws[i >> 5][i & 31] = p.w[(i >> 5) * p.cout_pad + (i & 31)];
    if (tid < 32) bs[tid] = p.bias[tid];
    const float* img = p.img + (long long)n * p.S * p.S * 3;
    for (int i = tid; i < (TH + 2) * (TW + 2) * 3; i += 256) {
        const int r = i / ((TW + 2) * 3), c = i - r * ((TW + 2) * 3);
        const int iy = y0 + r - 1, ix3 = x0 * 3 + c - 3;          // column index in floats of the image row
        float v = 0.f;
        if (iy >= 0 && iy < p.S && ix3 >= 0 && ix3 < p.S * 3) v = img[(long long)iy * p.S * 3 + ix3];
        patch[r][c] = v;
    }
    __syncthreads();
    const int tx = tid & 31, ty = tid >> 5;                      // ty in 0..7, pixels (ty, tx) and (ty+8, tx)
    float a0[32], a1[32];
#pragma unroll
    for (int c = 0; c < 32; c++) { a0[c] = 0.f; a1[c] = 0.f; }
#pragma unroll
    for (int kh = 0; kh < 3; kh++)
#pragma unroll
        for (int j = 0; j < 9; j++) {                            // j = kw*3 + c
            const float v0 = patch[ty + kh][tx * 3 + j];
            const float v1 = patch[ty + 8 + kh][tx * 3 + j];
            const float4* wr = reinterpret_cast<const float4*>(ws[kh * 9 + j]);
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const float4 w4 = wr[q];
                a0[4 * q + 0] = fmaf(v0, w4.x, a0[4 * q + 0]); a1[4 * q + 0] = fmaf(v1, w4.x, a1[4 * q + 0]);
                a0[4 * q + 1] = fmaf(v0, w4.y, a0[4 * q + 1]); a1[4 * q + 1] = fmaf(v1, w4.y, a1[4 * q + 1]);
                a0[4 * q + 2] = fmaf(v0, w4.z, a0[4 * q + 2]); a1[4 * q + 2] = fmaf(v1, w4.z, a1[4 * q + 2]);
                a0[4 * q + 3] = fmaf(v0, w4.w, a0[4 * q + 3]); a1[4 * q + 3] = fmaf(v1, w4.w, a1[4 * q + 3]);
            }
        }
    const int Sp = p.S + 2;
#pragma unroll
    for (int half = 0; half < 2; half++) {
        const int y = y0 + ty + half * 8, x = x0 + tx;
        if (y >= p.S || x >= p.S) continue;
        const float* a = half ? a1 : a0;
        uint4 o[4], ol[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            __half2 h[4], hl[4];
#pragma unroll
            for (int t = 0; t < 4; t++) {
                float u0 = a[8 * q + 2 * t] + bs[8 * q + 2 * t], u1 = a[8 * q + 2 * t + 1] + bs[8 * q + 2 * t + 1];
                u0 = u0 > 0.f ? u0 : 0.1f * u0; u1 = u1 > 0.f ? u1 : 0.1f * u1;          // leaky (custom_layers.py:101 default act)
                if (p.out_lo) { u0 *= 256.f; u1 *= 256.f; }                               // split planes are stored * 2^8
                h[t] = __floats2half2_rn(u0, u1);
                const float2 back = __half22float2(h[t]);
                hl[t] = __floats2half2_rn(u0 - back.x, u1 - back.y);
            }
            o[q].x = *reinterpret_cast<uint32_t*>(&h[0]); o[q].y = *reinterpret_cast<uint32_t*>(&h[1]);
            o[q].z = *reinterpret_cast<uint32_t*>(&h[2]); o[q].w = *reinterpret_cast<uint32_t*>(&h[3]);
            ol[q].x = *reinterpret_cast<uint32_t*>(&hl[0]); ol[q].y = *reinterpret_cast<uint32_t*>(&hl[1]);
            ol[q].z = *reinterpret_cast<uint32_t*>(&hl[2]); ol[q].w = *reinterpret_cast<uint32_t*>(&hl[3]);
        }
        const long long off = (((long long)n * Sp + y + 1) * Sp + x + 1) * 32;
        uint4* op = reinterpret_cast<uint4*>(p.out + off);
#pragma unroll
        for (int q = 0; q < 4; q++) op[q] = o[q];
        if (p.out_lo) {
            uint4* oq = reinterpret_cast<uint4*>(p.out_lo + off);
#pragma unroll
            for (int q = 0; q < 4; q++) oq[q] = ol[q];
        }
    }
}

struct SppParams {
    void* buf;            // concat buffer (N, H+2, W+2, ld)
    int ld, N, H, W, C;   // C = channels of x; x lives at channel offset 3*C, outputs at 0, C, 2*C
};

// fp16 SPP, 8 channels per thread (16 B loads, __hmax2): nested 5 c 9 c 13 windows in one sweep.
__global__ void spp_half8_kernel(SppParams p) {
    const int C8 = p.C >> 3;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)p.N * p.H * p.W * C8;
    if (i >= total) return;
    const int c8 = (int)(i % C8);
    long long t = i / C8;
    const int w = (int)(t % p.W); t /= p.W;
    const int h = (int)(t % p.H);
    const int n = (int)(t / p.H);
    const int Hp = p.H + 2, Wp = p.W + 2;
    const __half* base = reinterpret_cast<const __half*>(p.buf);
    const __half2 ninf = __float2half2_rn(-INFINITY);
    __half2 m5[4], m9[4], m13[4];
#pragma unroll
    for (int k = 0; k < 4; k++) { m5[k] = ninf; m9[k] = ninf; m13[k] = ninf; }
    for (int dy = -6; dy <= 6; dy++) {
        const int hh = h + dy;
        if (hh < 0 || hh >= p.H) continue;
        const int ady = dy < 0 ? -dy : dy;
        for (int dx = -6; dx <= 6; dx++) {
            const int ww = w + dx;
            if (ww < 0 || ww >= p.W) continue;
            const uint4 u = *reinterpret_cast<const uint4*>(base + (((long long)n * Hp + hh + 1) * Wp + ww + 1) * p.ld + 3 * p.C + c8 * 8);
            const __half2* v = reinterpret_cast<const __half2*>(&u);
            const int adx = dx < 0 ? -dx : dx;
            const int d = ady > adx ? ady : adx;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                m13[k] = __hmax2(m13[k], v[k]);
                if (d <= 4) m9[k] = __hmax2(m9[k], v[k]);
                if (d <= 2) m5[k] = __hmax2(m5[k], v[k]);
            }
        }
    }
    __half* ob = reinterpret_cast<__half*>(p.buf) + (((long long)n * Hp + h + 1) * Wp + w + 1) * p.ld + c8 * 8;
    *reinterpret_cast<uint4*>(ob) = *reinterpret_cast<uint4*>(m13);
    *reinterpret_cast<uint4*>(ob + p.C) = *reinterpret_cast<uint4*>(m9);
    *reinterpret_cast<uint4*>(ob + 2 * p.C) = *reinterpret_cast<uint4*>(m5);
}

// Separable fp16 SPP: one CTA per (image, 32-channel block).  The block's H x W x 32 slice of x is staged in shared memory,
// a horizontal pass builds the running maxima of widths 5 / 9 / 13 (nested windows), a vertical pass finishes them:
// 13 + 27 shared-memory reads per output position instead of 169 global ones (the sweep kernel above ran 16x above its
// HBM floor).  max() is exact, so the result is identical whatever the order.
constexpr int kSppChunk = 32;                                 // channels per CTA (4 x 16 B)
__global__ void __launch_bounds__(256) spp_sep_kernel(SppParams p) {
    extern __shared__ __align__(16) unsigned char spp_smem[];
    const int HW = p.H * p.W;
    uint4* sx = reinterpret_cast<uint4*>(spp_smem);           // [HW][4]
    uint4* r5 = sx + HW * 4; uint4* r9 = r5 + HW * 4; uint4* r13 = r9 + HW * 4;
    const int blocks_c = p.C / kSppChunk;
    const int n = blockIdx.x / blocks_c, cb = blockIdx.x - n * blocks_c;
    const int Hp = p.H + 2, Wp = p.W + 2;
    __half* base = reinterpret_cast<__half*>(p.buf);
    const int items = HW * 4;
    for (int i = threadIdx.x; i < items; i += blockDim.x) {
        const int q = i & 3, pos = i >> 2;
        const int h = pos / p.W, w = pos - h * p.W;
        sx[i] = *reinterpret_cast<const uint4*>(base + (((long long)n * Hp + h + 1) * Wp + w + 1) * p.ld + 3 * p.C + cb * kSppChunk + q * 8);
    }
    __syncthreads();
    auto hmax4 = [](uint4 a, const uint4 b) {
        __half2* x = reinterpret_cast<__half2*>(&a);
        const __half2* y = reinterpret_cast<const __half2*>(&b);
#pragma unroll
        for (int k = 0; k < 4; k++) x[k] = __hmax2(x[k], y[k]);
        return a;
    };
    for (int i = threadIdx.x; i < items; i += blockDim.x) {
        const int q = i & 3, pos = i >> 2;
        const int h = pos / p.W, w = pos - h * p.W;
        uint4 m = sx[i];
        for (int d = 1; d <= 2; d++) { if (w - d >= 0) m = hmax4(m, sx[i - 4 * d]); if (w + d < p.W) m = hmax4(m, sx[i + 4 * d]); }
        r5[i] = m;
        for (int d = 3; d <= 4; d++) { if (w - d >= 0) m = hmax4(m, sx[i - 4 * d]); if (w + d < p.W) m = hmax4(m, sx[i + 4 * d]); }
        r9[i] = m;
        for (int d = 5; d <= 6; d++) { if (w - d >= 0) m = hmax4(m, sx[i - 4 * d]); if (w + d < p.W) m = hmax4(m, sx[i + 4 * d]); }
        r13[i] = m;
        (void)h; (void)q;
    }
    __syncthreads();
    const int rowq = p.W * 4;
    for (int i = threadIdx.x; i < items; i += blockDim.x) {
        const int q = i & 3, pos = i >> 2;
        const int h = pos / p.W, w = pos - h * p.W;
        uint4 m5 = r5[i], m9 = r9[i], m13 = r13[i];
        for (int d = 1; d <= 6; d++) {
            if (h - d >= 0) { m13 = hmax4(m13, r13[i - d * rowq]); if (d <= 4) m9 = hmax4(m9, r9[i - d * rowq]); if (d <= 2) m5 = hmax4(m5, r5[i - d * rowq]); }
            if (h + d < p.H) { m13 = hmax4(m13, r13[i + d * rowq]); if (d <= 4) m9 = hmax4(m9, r9[i + d * rowq]); if (d <= 2) m5 = hmax4(m5, r5[i + d * rowq]); }
        }
        __half* ob = base + (((long long)n * Hp + h + 1) * Wp + w + 1) * p.ld + cb * kSppChunk + q * 8;
        *reinterpret_cast<uint4*>(ob) = m13;
        *reinterpret_cast<uint4*>(ob + p.C) = m9;
        *reinterpret_cast<uint4*>(ob + 2 * p.C) = m5;
    }
}

// SPP (custom_layers.py:130-134): mp13 | mp9 | mp5 | x, stride 1, 'same' (out-of-range cells ignored).
// One thread per (n,h,w,c): nested windows 5 c 9 c 13 share one sweep.

template <typename T>
__global__ void spp_kernel(SppParams p) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)p.N * p.H * p.W * p.C;
    if (i >= total) return;
    int c = (int)(i % p.C);
    long long t = i / p.C;
    int w = (int)(t % p.W); t /= p.W;
    int h = (int)(t % p.H);
    int n = (int)(t / p.H);
    const int Hp = p.H + 2, Wp = p.W + 2;
    T* base = reinterpret_cast<T*>(p.buf);
    float m5 = -INFINITY, m9 = -INFINITY, m13 = -INFINITY;
    for (int dy = -6; dy <= 6; dy++) {
        int hh = h + dy;
        if (hh < 0 || hh >= p.H) continue;
        for (int dx = -6; dx <= 6; dx++) {
            int ww = w + dx;
            if (ww < 0 || ww >= p.W) continue;
            float v = ld_as_float(base + (((long long)n * Hp + hh + 1) * Wp + ww + 1) * p.ld + 3 * p.C + c);
            int ady = dy < 0 ? -dy : dy, adx = dx < 0 ? -dx : dx;
            int d = ady > adx ? ady : adx;
            m13 = fmaxf(m13, v);
            if (d <= 4) m9 = fmaxf(m9, v);
            if (d <= 2) m5 = fmaxf(m5, v);
        }
    }
    long long o = (((long long)n * Hp + h + 1) * Wp + w + 1) * p.ld + c;
    st_from_float(base + o, m13);
    st_from_float(base + o + p.C, m9);
    st_from_float(base + o + 2 * p.C, m5);
}

// split-precision SPP: values are hi + lo pairs; max picks one of its inputs, so re-splitting is exact
__global__ void spp_split_kernel(SppParams p, void* buf_lo) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)p.N * p.H * p.W * p.C;
    if (i >= total) return;
    int c = (int)(i % p.C);
    long long t = i / p.C;
    int w = (int)(t % p.W); t /= p.W;
    int h = (int)(t % p.H);
    int n = (int)(t / p.H);
    const int Hp = p.H + 2, Wp = p.W + 2;
    __half* hi = reinterpret_cast<__half*>(p.buf);
    __half* lo = reinterpret_cast<__half*>(buf_lo);
    float m5 = -INFINITY, m9 = -INFINITY, m13 = -INFINITY;
    for (int dy = -6; dy <= 6; dy++) {
        int hh = h + dy;
        if (hh < 0 || hh >= p.H) continue;
        for (int dx = -6; dx <= 6; dx++) {
            int ww = w + dx;
            if (ww < 0 || ww >= p.W) continue;
            const long long o = (((long long)n * Hp + hh + 1) * Wp + ww + 1) * p.ld + 3 * p.C + c;
            float v = __half2float(hi[o]) + __half2float(lo[o]);
            int ady = dy < 0 ? -dy : dy, adx = dx < 0 ? -dx : dx;
            int d = ady > adx ? ady : adx;
            m13 = fmaxf(m13, v);
            if (d <= 4) m9 = fmaxf(m9, v);
            if (d <= 2) m5 = fmaxf(m5, v);
        }
    }
    const long long o = (((long long)n * Hp + h + 1) * Wp + w + 1) * p.ld + c;
    const float vals[3] = {m13, m9, m5};
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const __half a = __float2half_rn(vals[k]);
        hi[o + k * p.C] = a;
        lo[o + k * p.C] = __float2half_rn(vals[k] - __half2float(a));
    }
}

__global__ void gather_view_split_kernel(const __half* hi, const __half* lo, float* dst, int N, int H, int W, int C, int ld, int choff, float inv_scale) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)N * H * W * C;
    if (i >= total) return;
    int c = (int)(i % C);
    long long t = i / C;
    int w = (int)(t % W); t /= W;
    int h = (int)(t % H);
    int n = (int)(t / H);
    const long long o = (((long long)n * (H + 2) + h + 1) * (W + 2) + w + 1) * ld + choff + c;
    dst[i] = (__half2float(hi[o]) + __half2float(lo[o])) * inv_scale;
}

// splitmix64-style hash -> [0,1) with 24 random bits; bit-identical to oracle/y4_oracle.py:hash_uniform.
__device__ __forceinline__ float hash_uniform(uint64_t idx, uint64_t seed) {
    uint64_t z = idx + seed * 0x9E3779B97F4A7C15ull;
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    return (float)(z >> 40) * (1.0f / 16777216.0f);
}

__global__ void synth_fill_kernel(float* img, uint64_t seed, uint64_t first_elem, long long count) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < count; i += stride) img[i] = hash_uniform(first_elem + (uint64_t)i, seed);
}

__global__ void l2_flush_kernel(float4* buf, long long n, float v) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) buf[i] = make_float4(v, v, v, v);
}

// packed user heads (B,g,g,C) -> padded-flat fp32 head buffer (ld = ld_out)
__global__ void scatter_head_kernel(const float* src, float* dst, int N, int g, int C, int ld_out, float* obj, long long obj_rows, int nc5) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)N * g * g * C;
    if (i >= total) return;
    int c = (int)(i % C);
    long long t = i / C;
    int w = (int)(t % g); t /= g;
    int h = (int)(t % g);
    int n = (int)(t / g);
    const long long prow = ((long long)n * (g + 2) + h + 1) * (g + 2) + w + 1;
    dst[prow * ld_out + c] = src[i];
    if (obj && c % nc5 == 4) obj[(c / nc5) * obj_rows + prow] = src[i];      // compact objectness copy (decode_nms.cuh DecodeParams::obj)
}

// padded-flat (any T) view -> packed NHWC float (debug / y4_forward_heads)
template <typename T>
__global__ void gather_view_kernel(const T* src, float* dst, int N, int H, int W, int C, int ld, int choff) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)N * H * W * C;
    if (i >= total) return;
    int c = (int)(i % C);
    long long t = i / C;
    int w = (int)(t % W); t /= W;
    int h = (int)(t % H);
    int n = (int)(t / H);
    dst[i] = ld_as_float(src + (((long long)n * (H + 2) + h + 1) * (W + 2) + w + 1) * ld + choff + c);
}

}  // namespace y4
