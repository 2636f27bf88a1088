// zerocopy_rate.cu — how fast can SMs write straight into pinned host memory (no copy engine)?  Decides whether k_observe's
// packed outputs could go to the host directly in hope_step_host.   nvcc -arch=sm_100a -O3 -o zerocopy_rate zerocopy_rate.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

// one warp per "env": writes `len` doubles at offset env * stride (8-byte aligned, not sector aligned when stride % 4 != 0)
__global__ void k_rows(double *dst, int n, int len, int stride) {
    const int env = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (env >= n) return;
    double *row = dst + (size_t)env * stride;
    for (int j = lane; j < len; j += 32) row[j] = (double)(env + j);
}
// the wire format of one env as k_observe would write it: kept values (len doubles at an unaligned offset) + a 16-byte flag record +
// a 4-byte offset + 42 step-count bytes, each into its own array
__global__ void k_env_records(double *vals, uint4 *bits, unsigned *off, unsigned char *steps, int n, int len, int stride, int hdr) {
    const int env = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (env >= n) return;
    double *row = vals + (size_t)env * stride;
    for (int j = lane; j < len; j += 32) row[j] = (double)(env + j);
    if (hdr & 1) { if (lane == 0) { bits[env] = make_uint4(env, 1, 2, 3); off[env] = env; } }
    if (hdr & 2) { steps[(size_t)env * 42 + lane] = (unsigned char)lane; if (lane < 10) steps[(size_t)env * 42 + 32 + lane] = 1; }
}
// the same data volume as 16-byte vector stores, fully coalesced
__global__ void k_stream(double2 *dst, size_t n2) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x) dst[i] = make_double2((double)i, 1.0);
}

int main() {
    const int n = 65536, len = 68, stride = 69;   // 544 B per env, rows start at odd multiples of 8 B
    const size_t bytes = (size_t)n * stride * 8;
    double *h, *d;
    CK(cudaHostAlloc(&h, bytes, cudaHostAllocMapped));
    CK(cudaMalloc(&d, bytes));
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    float ms;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(a)); CK(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost)); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
        CK(cudaEventElapsedTime(&ms, a, b)); printf("copy engine D2H        %6.1f MB  %.3f ms  %.1f GB/s\n", bytes / 1e6, ms, bytes / ms / 1e6);
        CK(cudaEventRecord(a)); k_rows<<<n / 4, 128>>>(d, n, len, stride); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
        CK(cudaEventElapsedTime(&ms, a, b)); printf("rows -> HBM            %6.1f MB  %.3f ms  %.1f GB/s\n", n * len * 8 / 1e6, ms, n * len * 8.0 / ms / 1e6);
        CK(cudaEventRecord(a)); k_rows<<<n / 4, 128>>>(h, n, len, stride); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
        CK(cudaEventElapsedTime(&ms, a, b)); printf("rows -> host (unaligned)%5.1f MB  %.3f ms  %.1f GB/s\n", n * len * 8 / 1e6, ms, n * len * 8.0 / ms / 1e6);
        CK(cudaEventRecord(a)); k_rows<<<n / 4, 128>>>(h, n, 64, 64); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
        CK(cudaEventElapsedTime(&ms, a, b)); printf("rows -> host (512 B aligned)%5.1f MB  %.3f ms  %.1f GB/s\n", n * 64 * 8 / 1e6, ms, n * 64 * 8.0 / ms / 1e6);
        for (int hdr = 1; hdr <= 3; hdr += 2) {
            static uint4 *hb = nullptr; static unsigned *ho = nullptr; static unsigned char *hs = nullptr;
            if (!hb) { CK(cudaHostAlloc(&hb, 16 * n, cudaHostAllocMapped)); CK(cudaHostAlloc(&ho, 4 * n, cudaHostAllocMapped)); CK(cudaHostAlloc(&hs, 42 * n, cudaHostAllocMapped)); }
            CK(cudaEventRecord(a)); k_env_records<<<n / 2, 64>>>(h, hb, ho, hs, n, len, stride, hdr); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
            const double tot = n * (len * 8.0 + 20 + (hdr & 2 ? 42 : 0));
            CK(cudaEventElapsedTime(&ms, a, b)); printf("env records -> host (hdr %d) %5.1f MB  %.3f ms  %.1f GB/s\n", hdr, tot / 1e6, ms, tot / ms / 1e6);
        }
        CK(cudaEventRecord(a)); k_stream<<<148 * 4, 256>>>((double2 *)h, bytes / 16); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
        CK(cudaEventElapsedTime(&ms, a, b)); printf("stream -> host         %6.1f MB  %.3f ms  %.1f GB/s\n", bytes / 1e6, ms, bytes / ms / 1e6);
        CK(cudaEventRecord(a)); k_stream<<<16, 256>>>((double2 *)h, bytes / 16); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
        CK(cudaEventElapsedTime(&ms, a, b)); printf("stream -> host, 16 CTAs %5.1f MB  %.3f ms  %.1f GB/s\n", bytes / 1e6, ms, bytes / ms / 1e6);
    }
    double s = 0; for (size_t i = 0; i < bytes / 8; i += 4097) s += h[i];
    printf("checksum %g\n", s);
    return 0;
}
