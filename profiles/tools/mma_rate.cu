// Microbenchmark: throughput of the legacy tensor path (mma.sync.m16n8k16 bf16 -> f32) on this GPU, per SM.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
__global__ void __launch_bounds__(256) k(float *out, int iters, uint32_t a0, uint32_t b0) {
    float c[8][4];
    for (int i = 0; i < 8; ++i) for (int q = 0; q < 4; ++q) c[i][q] = 0.f;
    uint32_t a[4] = {a0, a0 + 1, a0 + 2, a0 + 3}, b[2] = {b0, b0 + 1};
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    float s = 0.f;
    for (int i = 0; i < 8; ++i) for (int q = 0; q < 4; ++q) s += c[i][q];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float *out; cudaMalloc(&out, sizeof(float) * sms * 8 * 256);
    for (int warps_per_sm : {4, 8, 16, 32, 64}) {
        const int blocks = sms * warps_per_sm / 8, iters = 20000;
        k<<<blocks, 256>>>(out, 100, 0, 0);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0); k<<<blocks, 256>>>(out, iters, 0x3f803f80u, 0x3f803f80u); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double mmas = (double)blocks * 8 * iters * 8, flops = mmas * 2 * 16 * 8 * 16;
        printf("{\"warps_per_sm\": %d, \"ms\": %.3f, \"tflops\": %.1f, \"mma_per_sm_per_us\": %.1f}\n", warps_per_sm, ms, flops / ms / 1e9, mmas / sms / (ms * 1e3));
    }
    return 0;
}
