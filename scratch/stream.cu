// scratch micro-benchmark (not part of the product): what bounds the class-score stream of the
// predicted-semantics ingest?  16 envs x 40 planes x 256x256 f32 (168 MB per frame, 4 rotating frames).
//   plain     grid-stride LDG.128 read of the whole frame (upper bound for a read-only stream)
//   direct    one thread = 4 pixels, loops over the planes with 8 LDG.128 in flight (register staged)
//   ring      persistent CTAs, producer lane issues cp.async.bulk into an smem ring, 8 consumer warps
//             mode 0: consumers only wait + release   1: exact argmax (3 FSETP form)   2: GT + FADD NaN probe
//             map 0: contiguous tile range per CTA    1: round-robin tiles
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scratch/stream scratch/stream.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#define B 16
#define NCLS 40
#define HW 65536

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
        ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}
__device__ __forceinline__ float4 ld_cs_v4(const float *p) {
    float4 v;
    asm volatile("ld.global.cs.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

// all 8 loads must have been ISSUED before any is consumed: an empty volatile asm that names every loaded register
#define PIN8(v)                                                                                                   \
    asm volatile("" : "+f"(v[0].x), "+f"(v[0].y), "+f"(v[0].z), "+f"(v[0].w), "+f"(v[1].x), "+f"(v[1].y), "+f"(v[1].z), \
                 "+f"(v[1].w), "+f"(v[2].x), "+f"(v[2].y), "+f"(v[2].z), "+f"(v[2].w), "+f"(v[3].x), "+f"(v[3].y),      \
                 "+f"(v[3].z), "+f"(v[3].w));                                                                      \
    asm volatile("" : "+f"(v[4].x), "+f"(v[4].y), "+f"(v[4].z), "+f"(v[4].w), "+f"(v[5].x), "+f"(v[5].y), "+f"(v[5].z), \
                 "+f"(v[5].w), "+f"(v[6].x), "+f"(v[6].y), "+f"(v[6].z), "+f"(v[6].w), "+f"(v[7].x), "+f"(v[7].y),      \
                 "+f"(v[7].z), "+f"(v[7].w));
#define ARGMAX_EXACT(v, k, best, arg) \
    if (((v) > (best)) || ((v) != (v) && (best) == (best))) { (best) = (v); (arg) = (k); }
#define ARGMAX_FAST(v, k, best, arg, s) \
    { (s) = __fadd_rn((s), (v)); if ((v) > (best)) { (best) = (v); (arg) = (k); } }

__global__ void __launch_bounds__(256) k_plain(const float4 *__restrict__ in, size_t n4, float *sink) {
    float acc = 0.f;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 7 * stride < n4; i += 8 * stride) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = ld_cs_v4(reinterpret_cast<const float *>(in + i + u * stride));
        PIN8(v)
#pragma unroll
        for (int u = 0; u < 8; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
    }
    for (; i < n4; i += stride) { float4 v = in[i]; acc += v.x + v.y + v.z + v.w; }
    if (acc == 1.2345e-30f) *sink = acc;
}

// mode 0: sum only; 1: exact argmax; 2: fast argmax
template <int MODE>
__global__ void __launch_bounds__(256) k_direct(const float *__restrict__ logits, uint8_t *__restrict__ labels, float *sink) {
    const int b = blockIdx.y;
    const int pix0 = (blockIdx.x * 256 + threadIdx.x) * 4;
    const float *lp = logits + (size_t)b * NCLS * HW + pix0;
    float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY}, s[4] = {0, 0, 0, 0};
    int arg[4] = {0, 0, 0, 0};
    for (int k = 0; k < NCLS; k += 8) {
        float4 v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = ld_cs_v4(lp + (size_t)(k + q) * HW);
        PIN8(v)
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const float e[4] = {v[q].x, v[q].y, v[q].z, v[q].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (MODE == 0) s[j] += e[j];
                else if (MODE == 1) { ARGMAX_EXACT(e[j], k + q, best[j], arg[j]); }
                else { ARGMAX_FAST(e[j], k + q, best[j], arg[j], s[j]); }
            }
        }
    }
    if (MODE == 0) { if (s[0] + s[1] + s[2] + s[3] == 1.2345e-30f) *sink = s[0]; return; }
    if (MODE == 2 && (s[0] + s[1]) + (s[2] + s[3]) != (s[0] + s[1]) + (s[2] + s[3])) arg[0] = 255;  // slow path stand-in
    uchar4 o;
    o.x = (uint8_t)arg[0]; o.y = (uint8_t)arg[1]; o.z = (uint8_t)arg[2]; o.w = (uint8_t)arg[3];
    *reinterpret_cast<uchar4 *>(labels + (size_t)b * HW + pix0) = o;
}

// TILE pixels per tile = 256 consumer threads x PX; SP planes per stage; NSTAGE stages
template <int PX, int SP, int NSTAGE, int MODE>
__global__ void __launch_bounds__(288) k_ring(const float *__restrict__ logits, uint8_t *__restrict__ labels, int map) {
    constexpr int TILE = 256 * PX;
    extern __shared__ __align__(128) unsigned char dyn[];
    float(*ring)[SP][TILE] = reinterpret_cast<float(*)[SP][TILE]>(dyn);
    __shared__ __align__(8) uint64_t full[NSTAGE], empty[NSTAGE];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int tpe = HW / TILE;
    const int total = B * tpe;
    int t0, t1, tstep;
    if (map == 0) {
        t0 = (int)((long long)blockIdx.x * total / gridDim.x);
        t1 = (int)((long long)(blockIdx.x + 1) * total / gridDim.x);
        tstep = 1;
    } else {
        t0 = blockIdx.x; t1 = total; tstep = gridDim.x;
    }
    if (tid == 0) {
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp == 8) {
        if (lane == 0) {
            uint64_t policy;
            asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
            int slot = 0;
            uint32_t round = 0;
            for (int tile = t0; tile < t1; tile += tstep) {
                const int eb = tile / tpe, tp0 = (tile - eb * tpe) * TILE;
                const float *src = logits + (size_t)eb * NCLS * HW + tp0;
                for (int p0 = 0; p0 < NCLS; p0 += SP) {
                    if (round > 0) mbar_wait(&empty[slot], (round - 1) & 1u);
                    mbar_expect_tx(&full[slot], (uint32_t)(SP * TILE * sizeof(float)));
                    for (int p = 0; p < SP; ++p)
                        bulk_g2s(&ring[slot][p][0], src + (size_t)(p0 + p) * HW, TILE * sizeof(float), &full[slot], policy);
                    if (++slot == NSTAGE) { slot = 0; ++round; }
                }
            }
        }
    } else {
        int slot = 0;
        uint32_t round = 0;
        for (int tile = t0; tile < t1; tile += tstep) {
            const int b = tile / tpe, tp0 = (tile - b * tpe) * TILE;
            float best[PX], s[PX];
            int arg[PX];
#pragma unroll
            for (int j = 0; j < PX; ++j) { best[j] = -INFINITY; s[j] = 0.f; arg[j] = 0; }
            for (int p0 = 0; p0 < NCLS; p0 += SP) {
                mbar_wait(&full[slot], round & 1u);
                if (MODE != 0) {
                    float v[SP][PX];
#pragma unroll
                    for (int p = 0; p < SP; ++p) {
                        if (PX == 2) {
                            const float2 t = *reinterpret_cast<const float2 *>(&ring[slot][p][tid * 2]);
                            v[p][0] = t.x; v[p][1 % PX] = t.y;
                        } else {
#pragma unroll
                            for (int h = 0; h < PX / 4; ++h) {
                                // 16-byte lane stride inside each 128-pixel group keeps LDS.128 conflict-free
                                const float4 t = *reinterpret_cast<const float4 *>(&ring[slot][p][h * 1024 + tid * 4]);
                                v[p][(4 * h) % PX] = t.x; v[p][(4 * h + 1) % PX] = t.y; v[p][(4 * h + 2) % PX] = t.z; v[p][(4 * h + 3) % PX] = t.w;
                            }
                        }
                    }
#pragma unroll
                    for (int p = 0; p < SP; ++p)
#pragma unroll
                        for (int j = 0; j < PX; ++j) {
                            if (MODE == 1) { ARGMAX_EXACT(v[p][j], p0 + p, best[j], arg[j]); }
                            else { ARGMAX_FAST(v[p][j], p0 + p, best[j], arg[j], s[j]); }
                        }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[slot]);
                if (++slot == NSTAGE) { slot = 0; ++round; }
            }
            if (MODE != 0) {
                uint8_t *out = labels + (size_t)b * HW + tp0;
                if (MODE == 2) {
                    float t = 0.f;
#pragma unroll
                    for (int j = 0; j < PX; ++j) t += s[j];
                    if (t != t) arg[0] = 255;
                }
                if (PX == 2) {
                    uchar2 o; o.x = (uint8_t)arg[0]; o.y = (uint8_t)arg[1 % PX];
                    *reinterpret_cast<uchar2 *>(out + tid * 2) = o;
                } else {
#pragma unroll
                    for (int h = 0; h < PX / 4; ++h) {
                        uchar4 o;
                        o.x = (uint8_t)arg[(4 * h) % PX]; o.y = (uint8_t)arg[(4 * h + 1) % PX];
                        o.z = (uint8_t)arg[(4 * h + 2) % PX]; o.w = (uint8_t)arg[(4 * h + 3) % PX];
                        *reinterpret_cast<uchar4 *>(out + h * 1024 + tid * 4) = o;
                    }
                }
            }
        }
    }
}

static float *g_frames[4];
static uint8_t *g_labels;
static float *g_sink;
static const size_t FRAME = (size_t)B * NCLS * HW;

template <class F>
static void run(const char *name, F launch) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 4; ++i) launch(g_frames[i & 3]);
    cudaEventRecord(e0);
    const int N = 40;
    for (int i = 0; i < N; ++i) launch(g_frames[i & 3]);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double us = ms * 1000.0 / N;
    printf("%-44s %8.2f us  %7.1f GB/s  %s\n", name, us, FRAME * 4.0 / us * 1e-3, err == cudaSuccess ? "" : cudaGetErrorString(err));
    fflush(stdout);
}

template <int PX, int SP, int NSTAGE, int MODE>
static void run_ring(int ctas_per_sm, int map) {
    const size_t smem = (size_t)NSTAGE * SP * 256 * PX * 4;
    auto fn = k_ring<PX, SP, NSTAGE, MODE>;
    cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, 288, smem);
    char name[128];
    snprintf(name, sizeof(name), "ring tile=%d sp=%d st=%d (%zuK) cta/sm=%d(occ %d) map=%d mode=%d", 256 * PX, SP, NSTAGE, smem >> 10,
             ctas_per_sm, occ, map, MODE);
    if (occ < ctas_per_sm) { printf("%s: skipped\n", name); return; }
    const int grid = 148 * ctas_per_sm;
    run(name, [&](const float *f) { fn<<<grid, 288, smem>>>(f, g_labels, map); });
}

int main() {
    for (int i = 0; i < 4; ++i) {
        cudaMalloc(&g_frames[i], FRAME * 4);
        cudaMemset(g_frames[i], 0, FRAME * 4);
    }
    cudaMalloc(&g_labels, (size_t)B * HW);
    cudaMalloc(&g_sink, 64);
    // some structure so the argmax is not trivially constant: plane k gets value (k * 7919 % 40) in a few spots
    {
        float *h = (float *)malloc(FRAME * 4);
        uint32_t x = 12345;
        for (size_t i = 0; i < FRAME; ++i) { x = x * 1664525u + 1013904223u; h[i] = (float)(x >> 8) * (1.0f / 16777216.0f) - 0.5f; }
        for (int i = 0; i < 4; ++i) cudaMemcpy(g_frames[i], h, FRAME * 4, cudaMemcpyHostToDevice);
        free(h);
    }
    run("plain LDG.128 read, grid 148x8", [&](const float *f) { k_plain<<<148 * 8, 256>>>((const float4 *)f, FRAME / 4, g_sink); });
    run("plain LDG.128 read, grid 148x4", [&](const float *f) { k_plain<<<148 * 4, 256>>>((const float4 *)f, FRAME / 4, g_sink); });
    run("direct 4px/thread sum", [&](const float *f) { k_direct<0><<<dim3(HW / 1024, B), 256>>>(f, g_labels, g_sink); });
    run("direct 4px/thread exact argmax", [&](const float *f) { k_direct<1><<<dim3(HW / 1024, B), 256>>>(f, g_labels, g_sink); });
    run("direct 4px/thread fast argmax", [&](const float *f) { k_direct<2><<<dim3(HW / 1024, B), 256>>>(f, g_labels, g_sink); });
    for (int map = 0; map < 2; ++map) {
        run_ring<2, 8, 6, 0>(2, map);
        run_ring<2, 8, 6, 1>(2, map);
        run_ring<2, 8, 6, 2>(2, map);
        run_ring<4, 4, 6, 0>(2, map);
        run_ring<4, 4, 6, 1>(2, map);
        run_ring<4, 4, 6, 2>(2, map);
        run_ring<4, 8, 3, 0>(2, map);
        run_ring<4, 8, 3, 2>(2, map);
        run_ring<8, 4, 3, 0>(2, map);
        run_ring<8, 4, 3, 2>(2, map);
        run_ring<4, 8, 6, 0>(1, map);
        run_ring<4, 8, 6, 2>(1, map);
        run_ring<8, 4, 6, 0>(1, map);
        run_ring<8, 4, 6, 2>(1, map);
        run_ring<2, 8, 3, 0>(4, map);
        run_ring<2, 8, 3, 2>(4, map);
    }
    printf("done: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
