// scratch: cost of executing cold straight-line code once per launch vs a loop with the same dynamic instruction count
#include <cstdio>
#include <cuda_runtime.h>
#define S1(k) x = fmaf(x, a, 0.5f + (float)(k));
#define S10(k) S1(k) S1(k+1) S1(k+2) S1(k+3) S1(k+4) S1(k+5) S1(k+6) S1(k+7) S1(k+8) S1(k+9)
#define S100(k) S10(k) S10(k+10) S10(k+20) S10(k+30) S10(k+40) S10(k+50) S10(k+60) S10(k+70) S10(k+80) S10(k+90)
#define S1000(k) S100(k) S100(k+100) S100(k+200) S100(k+300) S100(k+400) S100(k+500) S100(k+600) S100(k+700) S100(k+800) S100(k+900)
template <int N> struct Unroll {
    static __device__ __forceinline__ float run(float x, float a) {
        S1000(0)
        if (N >= 4000) { S1000(1000) S1000(2000) S1000(3000) }
        if (N >= 8000) { S1000(4000) S1000(5000) S1000(6000) S1000(7000) }
        return x;
    }
};
template <int N> __global__ void k_straight(float *out, float a, long long *cyc) {
    long long t0 = clock64();
    float x = Unroll<N>::run((float)threadIdx.x, a);
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_loop(float *out, float a, int n, long long *cyc) {
    long long t0 = clock64();
    float x = (float)threadIdx.x;
#pragma unroll 4
    for (int i = 0; i < n; ++i) x = fmaf(x, a, 0.5f);
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int N> void test(float *out, long long *cyc, int grid, int threads, const char *tag) {
    long long h[1024];
    for (int rep = 0; rep < 3; ++rep) {
        k_straight<N><<<grid, threads>>>(out, 1.0001f, cyc);
        cudaMemcpy(h, cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
        long long mx = 0, mn = 1ll << 60; for (int i = 0; i < grid; ++i) { mx = h[i] > mx ? h[i] : mx; mn = h[i] < mn ? h[i] : mn; }
        printf("%s straight N=%d grid=%d thr=%d rep %d: min %lld max %lld cycles (%.2f / instr)\n", tag, N, grid, threads, rep, mn, mx, (double)mx / N);
        // evict: run a different big kernel in between? (done by caller ordering)
    }
}
int main() {
    float *out; long long *cyc;
    cudaMalloc(&out, 4 << 20); cudaMalloc(&cyc, 8192);
    long long h[1024];
    for (int threads : {32, 288}) {
        for (int grid : {1, 296}) {
            test<1000>(out, cyc, grid, threads, "A");
            test<4000>(out, cyc, grid, threads, "B");
            test<1000>(out, cyc, grid, threads, "A-again");
            test<8000>(out, cyc, grid, threads, "C");
            test<1000>(out, cyc, grid, threads, "A-after-C");
            k_loop<<<grid, threads>>>(out, 1.0001f, 4000, cyc);
            cudaMemcpy(h, cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
            printf("loop n=4000 grid=%d thr=%d: %lld cycles\n", grid, threads, h[0]);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
