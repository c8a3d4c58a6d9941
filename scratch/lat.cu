// scratch micro-benchmark (not part of the product): memory / instruction-fetch latencies on the B200
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long gt() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ long long clk() { return clock64(); }

// one warp: `nb` batches of 8 loads per lane; lane addresses = base + (batch*256 + u*32 + lane) * stride
__global__ void k_batched(const uint32_t *buf, size_t stride_words, int nb, long long *out, uint32_t *sink) {
    const int lane = threadIdx.x & 31;
    long long t0 = clk();
    uint32_t acc = 0;
    for (int b = 0; b < nb; ++b) {
        uint32_t v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __ldcg(buf + ((size_t)b * 256 + u * 32 + lane) * stride_words);
#pragma unroll
        for (int u = 0; u < 8; ++u) acc += v[u];
    }
    long long t1 = clk();
    if (acc == 0x12345678) *sink = acc;
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}
// dependent chain of single loads (pointer chase style, fixed stride)
__global__ void k_chain(const uint32_t *buf, size_t stride_words, int n, long long *out, uint32_t *sink) {
    long long t0 = clk();
    uint32_t idx = 0, acc = 0;
    for (int i = 0; i < n; ++i) { uint32_t v = __ldcg(buf + (size_t)(i + (idx & 1)) * stride_words); idx = v; acc += v; }
    long long t1 = clk();
    if (acc == 0x12345678) *sink = acc;
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}
__global__ void k_atomic_chain(uint32_t *ctr, int n, long long *out, uint32_t *sink) {
    long long t0 = clk();
    uint32_t acc = 0;
    for (int i = 0; i < n; ++i) acc += atomicAdd(ctr + (acc & 1), 1u);
    long long t1 = clk();
    if (acc == 0x12345678) *sink = acc;
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}
int main() {
    const size_t bytes = (size_t)1 << 30;
    uint32_t *buf, *sink; long long *out;
    cudaMalloc(&buf, bytes); cudaMemset(buf, 0, bytes); cudaMalloc(&sink, 64); cudaMalloc(&out, 1024);
    long long h[8];
    int clkrate; cudaDeviceGetAttribute(&clkrate, cudaDevAttrClockRate, 0);
    printf("clock %d kHz\n", clkrate);
    const size_t strides[] = {4, 32, 8192, 8192 + 32, 16384};   // words: 16 B, 128 B, 32 KB, 32 KB + 128 B, 64 KB
    for (size_t s : strides) {
        for (int rep = 0; rep < 2; ++rep) {
            // flush L2 by touching another region? (buffer is 1 GB, region used is small; rep 1 = warm L2)
            k_batched<<<1, 32>>>(buf, s, 6, out, sink);
            cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost);
            printf("batched stride %6zu B rep %d: %lld cycles for 6 batches of 8x32 loads (%.0f per batch)\n", s * 4, rep, h[0], h[0] / 6.0);
        }
    }
    for (size_t s : strides) {
        k_chain<<<1, 32>>>(buf + (256u << 20) / 4, s, 32, out, sink);   // cold region
        cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost);
        printf("chain cold  stride %6zu B: %.0f cycles per load\n", s * 4, h[0] / 32.0);
        k_chain<<<1, 32>>>(buf + (256u << 20) / 4, s, 32, out, sink);
        cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost);
        printf("chain warm  stride %6zu B: %.0f cycles per load\n", s * 4, h[0] / 32.0);
    }
    k_atomic_chain<<<1, 32>>>(buf, 64, out, sink);
    cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost);
    printf("atomic-with-return chain: %.0f cycles each\n", h[0] / 64.0);
    // many CTAs doing the column pattern concurrently (like 4 warps of one CTA)
    k_batched<<<1, 128>>>(buf + (512u << 20) / 4, 8192, 6, out, sink);
    cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost);
    printf("batched 4 warps same addresses cold: %lld cycles\n", h[0]);
    cudaError_t e = cudaDeviceSynchronize();
    printf("done: %s\n", cudaGetErrorString(e));
    return 0;
}
