// scratch: latency of batched random 16-byte loads vs the size of the region they fall in (TLB reach)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }
// every thread: nb batches of 8 independent random loads inside [0, region)
__global__ void k_rand(const uint4 *buf, size_t region_elems, int nb, long long *out, uint32_t *sink, uint32_t seed) {
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    long long t0 = clock64();
    uint32_t acc = 0;
    for (int b = 0; b < nb; ++b) {
        uint4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const uint64_t r = ((uint64_t)hash32(seed + gid * 131u + b * 8 + u) << 20) ^ hash32(gid + u * 77u + seed);
            v[u] = __ldcg(buf + (r % region_elems));
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) acc += v[u].x + v[u].w;
    }
    long long t1 = clock64();
    if (acc == 0x12345678) *sink = acc;
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}
int main() {
    const size_t bytes = (size_t)6 << 30;
    uint4 *buf; uint32_t *sink; long long *out;
    if (cudaMalloc(&buf, bytes) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMemset(buf, 0, bytes); cudaMalloc(&sink, 64); cudaMalloc(&out, 8192);
    long long h[1024];
    for (int grid : {1, 148, 296}) for (int thr : {32, 256}) {
        for (size_t mb : {16, 128, 512, 2048, 6000}) {
            const size_t elems = (mb << 20) / 16;
            uint32_t seed = 1;
            for (int rep = 0; rep < 2; ++rep) {
                k_rand<<<grid, thr>>>(buf, elems, 4, out, sink, seed += 977);
                cudaMemcpy(h, out, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
                long long mx = 0; double avg = 0; for (int i = 0; i < grid; ++i) { mx = h[i] > mx ? h[i] : mx; avg += h[i]; }
                if (rep) printf("grid %3d thr %3d region %5zu MB: avg %.0f max %lld cycles per 4 batches -> %.0f per batch\n", grid, thr, mb, avg / grid, mx, avg / grid / 4);
            }
        }
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
