// img_encoder.cu — the convolutional front of the actor's image encoder (row f1 / f2, BASELINE cfg 4 with USE_IMG) as one kernel.
//
// Reference: ImgEncoder of MultiObsEmbedding (src/model/network.py:198-299 with the shipped switches: no batch norm, residual
// on, tanh): two blocks  out = maxpool2(tanh(conv3x3(x))) + avgpool2(conv1x1(x)),  3 -> 4 -> 8 channels, on the 3 x 64 x 64 image
// / 255, then flatten to 2048 features in (channel, row, column) order.  As cuDNN calls these two blocks take 58 ms per
// 65 536 images (4- and 8-channel convolutions are far below any library tile), 68 % of the 4-modal policy forward.
//
// Here a CTA of 256 threads owns one image: the uint8 image becomes a zero-bordered float32 tile in shared memory, block 1 runs
// with 4 pooled positions per thread (the 4 x 4 input patch of a 2 x 2 pooling window is loaded once and feeds 4 positions x
// 4 channels x 27 products), its 4 x 32 x 32 output stays in shared memory (zero-bordered again) and block 2 runs with one
// pooled position per thread.  All weights travel in the kernel parameter block (1.9 KB: constant-bank operands, no loads).
// float32 arithmetic, bf16 features out (what the following Linear(2048, 256) consumes under autocast): 12 KB in, 4 KB out per image.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/hope_b200.h"

namespace hope_img {

constexpr int HW = HOPE_IMG_HW;        // 64
constexpr int C0 = HOPE_IMG_C;         // 3
constexpr int C1 = 4, C2 = 8;
constexpr int P1 = HW + 2, S1 = HW + 4;           // bordered rows / row stride of the input tile (even stride: float2 loads stay aligned)
constexpr int H1 = HW / 2, P2 = H1 + 2, S2 = H1 + 4;
constexpr int H2 = H1 / 2;
constexpr int THREADS = 256;
static_assert(H2 * H2 == THREADS, "block 2: one pooled position per thread");

struct Smem {
    float in[C0][P1][S1];    // image / 255 with a zero border
    float mid[C1][P2][S2];   // block 1 output with a zero border
};

__device__ __forceinline__ float tanh_fast(float v) {
    float r;
    asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}

// one 2 x 2 pooling window of a block: patch[ci][4][4] = the bordered input around it; returns maxpool(tanh(conv3x3)) + avgpool(conv1x1) per output channel
template <int CI, int CO>
__device__ __forceinline__ void block_window(const float (&patch)[CI][4][4], const float *__restrict__ w3 /* [CO][CI][3][3] */, const float *__restrict__ b3,
                                             const float *__restrict__ w1 /* [CO][CI] */, const float *__restrict__ b1, float (&out)[CO]) {
    float avg[CI];
#pragma unroll
    for (int ci = 0; ci < CI; ++ci) avg[ci] = 0.25f * ((patch[ci][1][1] + patch[ci][1][2]) + (patch[ci][2][1] + patch[ci][2][2]));
#pragma unroll
    for (int co = 0; co < CO; ++co) {
        float best = -2.f;  // tanh > -1
#pragma unroll
        for (int py = 0; py < 2; ++py)
#pragma unroll
            for (int px = 0; px < 2; ++px) {
                float acc = b3[co];
#pragma unroll
                for (int ci = 0; ci < CI; ++ci)
#pragma unroll
                    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                        for (int dx = 0; dx < 3; ++dx) acc = fmaf(w3[((co * CI + ci) * 3 + dy) * 3 + dx], patch[ci][py + dy][px + dx], acc);
                best = fmaxf(best, tanh_fast(acc));
            }
        float sc = b1[co];
#pragma unroll
        for (int ci = 0; ci < CI; ++ci) sc = fmaf(w1[co * CI + ci], avg[ci], sc);
        out[co] = best + sc;
    }
}

__global__ void __launch_bounds__(THREADS, 3) k_img_conv(int n, const uint8_t *__restrict__ img, hope_img_conv_weights W, __nv_bfloat16 *__restrict__ feat) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem &sm = *reinterpret_cast<Smem *>(smem_raw);
    const int tid = threadIdx.x, i = blockIdx.x;
    if (i >= n) return;
    // zero the borders (cheap: the whole tiles), then the image: 12 288 bytes as 768 16-byte loads
    for (int k = tid; k < (int)(sizeof(Smem) / 16); k += THREADS) reinterpret_cast<float4 *>(smem_raw)[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    const uint4 *src = reinterpret_cast<const uint4 *>(img + (size_t)i * C0 * HW * HW);
    for (int k = tid; k < C0 * HW * HW / 16; k += THREADS) {
        const uint4 v = __ldg(src + k);
        const int c = k / (HW * HW / 16), rem = k - c * (HW * HW / 16), y = rem / (HW / 16), x0 = (rem - y * (HW / 16)) * 16;
        float *dst = &sm.in[c][y + 1][x0 + 1];
        const uint32_t ws[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int b = 0; b < 4; ++b) dst[4 * q + b] = (float)((ws[q] >> (8 * b)) & 0xffu) * (1.0f / 255.0f);  // observation_processor.py:13-17
    }
    __syncthreads();
    // ---- block 1: 32 x 32 pooled positions, 4 per thread -------------------------------------------------------
    for (int p = tid; p < H1 * H1; p += THREADS) {
        const int py = p / H1, px = p - py * H1;
        float patch[C0][4][4];
#pragma unroll
        for (int c = 0; c < C0; ++c)
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const float2 a = *reinterpret_cast<const float2 *>(&sm.in[c][2 * py + r][2 * px]), b = *reinterpret_cast<const float2 *>(&sm.in[c][2 * py + r][2 * px + 2]);
                patch[c][r][0] = a.x; patch[c][r][1] = a.y; patch[c][r][2] = b.x; patch[c][r][3] = b.y;
            }
        float o[C1];
        block_window<C0, C1>(patch, W.conv1_w, W.conv1_b, W.short1_w, W.short1_b, o);
#pragma unroll
        for (int c = 0; c < C1; ++c) sm.mid[c][py + 1][px + 1] = o[c];
    }
    __syncthreads();
    // ---- block 2: 16 x 16 pooled positions, one per thread -----------------------------------------------------
    {
        const int py = tid / H2, px = tid - py * H2;
        float patch[C1][4][4];
#pragma unroll
        for (int c = 0; c < C1; ++c)
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const float2 a = *reinterpret_cast<const float2 *>(&sm.mid[c][2 * py + r][2 * px]), b = *reinterpret_cast<const float2 *>(&sm.mid[c][2 * py + r][2 * px + 2]);
                patch[c][r][0] = a.x; patch[c][r][1] = a.y; patch[c][r][2] = b.x; patch[c][r][3] = b.y;
            }
        float o[C2];
        block_window<C1, C2>(patch, W.conv2_w, W.conv2_b, W.short2_w, W.short2_b, o);
        __nv_bfloat16 *dst = feat + (size_t)i * (C2 * H2 * H2);
#pragma unroll
        for (int c = 0; c < C2; ++c) dst[c * (H2 * H2) + tid] = __float2bfloat16(o[c]);  // nn.Flatten of (8, 16, 16): channel, row, column
    }
}

}  // namespace hope_img

extern "C" {

int hope_img_conv_forward(int n, const uint8_t *d_img, const hope_img_conv_weights *w, void *d_feat_bf16, void *stream) {
    if (n <= 0 || !d_img || !w || !d_feat_bf16) return HOPE_ERR_INVALID;
    using namespace hope_img;
    const int smem = (int)sizeof(Smem);
    if (cudaFuncSetAttribute(k_img_conv, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return HOPE_ERR_CUDA;
    k_img_conv<<<n, THREADS, smem, static_cast<cudaStream_t>(stream)>>>(n, d_img, *w, static_cast<__nv_bfloat16 *>(d_feat_bf16));
    return cudaGetLastError() == cudaSuccess ? HOPE_OK : HOPE_ERR_CUDA;
}

}  // extern "C"
