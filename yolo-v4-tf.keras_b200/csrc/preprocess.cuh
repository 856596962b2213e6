// preprocess_img (models.py:95-98) on the GPU: cv2.resize(img, (S, S)) [INTER_LINEAR, 8-bit] then / 255.
// OpenCV resizes 8-bit images in fixed point (imgproc/resize.cpp, HResizeLinear / VResizeLinear<uchar,int,short>):
//   fx = (float)((dx + 0.5) * (double)w / S - 0.5); sx = floor(fx); fx -= sx; clamp to the image (fx = 0 at the borders)
//   coefficients as shorts: a1 = round_half_even(fx * 2048), a0 = round_half_even((1 - fx) * 2048)   (same for rows, unclamped fy)
//   H pass: T = S[sx] * a0 + S[sx+1] * a1          V pass: dst = (((b0 * (T0 >> 4)) >> 16) + ((b1 * (T1 >> 4)) >> 16) + 2) >> 2
// The oracle restates the same arithmetic in numpy and is pinned bit-exactly against cv2.resize on the reference's images
// (tests/test_oracle.py).  "/ 255." is a float64 division that Keras casts to float32: a 256-entry table built on the host.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace y4 {

struct PreImage { long long offset; int h, w; };            // byte offset of the image in the staging buffer

__constant__ float c_div255[256];

// one thread per output pixel: 4 source pixels x 3 channels -> 3 floats (NHWC float32, the engine's input layout)
__global__ void __launch_bounds__(256) preprocess_u8_kernel(const uint8_t* __restrict__ src, const PreImage* __restrict__ imgs,
                                                             float* __restrict__ dst, int S, int batch, int reverse_channels) {
    const int dx = blockIdx.x * blockDim.x + threadIdx.x;
    const int dy = blockIdx.y;
    const int b = blockIdx.z;
    if (dx >= S) return;
    const PreImage im = imgs[b];
    const double scale_x = (double)im.w / (double)S, scale_y = (double)im.h / (double)S;
    float fx = (float)(((double)dx + 0.5) * scale_x - 0.5);
    int sx = (int)floorf(fx);
    fx -= (float)sx;
    if (sx < 0) { fx = 0.f; sx = 0; }
    if (sx >= im.w - 1) { fx = 0.f; sx = im.w - 1; }
    const int a1 = __float2int_rn(fx * 2048.f), a0 = __float2int_rn((1.f - fx) * 2048.f);
    const int sx1 = min(sx + 1, im.w - 1);
    float fy = (float)(((double)dy + 0.5) * scale_y - 0.5);
    const int sy = (int)floorf(fy);
    fy -= (float)sy;
    const int b1 = __float2int_rn(fy * 2048.f), b0 = __float2int_rn((1.f - fy) * 2048.f);
    const int sy0 = min(max(sy, 0), im.h - 1), sy1 = min(max(sy + 1, 0), im.h - 1);
    const uint8_t* base = src + im.offset;
    const uint8_t* r0 = base + (long long)sy0 * im.w * 3;
    const uint8_t* r1 = base + (long long)sy1 * im.w * 3;
    float out[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const int t0 = (int)r0[sx * 3 + c] * a0 + (int)r0[sx1 * 3 + c] * a1;
        const int t1 = (int)r1[sx * 3 + c] * a0 + (int)r1[sx1 * 3 + c] * a1;
        int v = (((b0 * (t0 >> 4)) >> 16) + ((b1 * (t1 >> 4)) >> 16) + 2) >> 2;
        v = min(max(v, 0), 255);
        out[c] = c_div255[v];
    }
    float* o = dst + (((long long)b * S + dy) * S + dx) * 3;
    if (reverse_channels) { o[0] = out[2]; o[1] = out[1]; o[2] = out[0]; }
    else { o[0] = out[0]; o[1] = out[1]; o[2] = out[2]; }
}

}  // namespace y4
