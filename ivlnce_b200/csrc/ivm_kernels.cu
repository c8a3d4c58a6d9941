// ivm_kernels.cu -- sm_100a kernels + C ABI of the semantic-map update.
//
// One step (iterative mode) on the caller's stream is ONE launch of the persistent kernel k_step_overlap (see its
// header below; all CTAs co-resident, grid barriers between the phases, programmatically serialised behind the
// previous step) whenever the image tiles evenly and the inputs are aligned.  Otherwise the same phases run as
// four kernels:
//   K1 k_ingest_scatter  per-env O(1) reset / store re-centring / pose matrices (mapper.py:310-326, 127-138), then
//      (_bulk)           [argmax ->] unproject -> transform -> filter -> half-cell ->
//                        64-bit atomicMax into the frame-candidate plane   (mapper.py:381-474, core.py:117-230)
//   K2 k_ingest_resolve  the owning pixel of each candidate merges into the world store
//   K3 k_fixup           edge-collision fix-up of both de-dup stages        (mapper.py:461-474 quirk)
//   K4 k_raster          band filter -> ego transform -> cell -> smem max/or -> u8 maps (mapper.py:555-617, 884-901)
// No point cloud is ever written to HBM; the only per-pixel state is the 8-byte candidate word of the touched
// half-cells.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include <mutex>

#include "../../include/ivln_map.h"
#include "ivm_core.h"

#define IVM_THREADS 256
#define IVM_RASTER_THREADS 128
#define IVM_NSTAGES 5
#define IVM_EVPOOL 64
#define IVM_MAX_DEVICES 64

// ------------------------------------------------------------------ helpers
__device__ __forceinline__ int warp_min(int v) { return __reduce_min_sync(0xffffffffu, v); }
__device__ __forceinline__ int warp_max(int v) { return __reduce_max_sync(0xffffffffu, v); }
__device__ __forceinline__ unsigned warp_sum(unsigned v) { return __reduce_add_sync(0xffffffffu, v); }

__device__ __forceinline__ float4 ld_stream4(const float *p) { return __ldcs(reinterpret_cast<const float4 *>(p)); }

// ------------------------------------------------------------------ one-time init
__global__ void k_init(IvmParams P) {
    const size_t n = (size_t)P.hmask + 1;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        P.hkeys[i] = IVM_EMPTY_KEY; P.hxord[i] = IVM_EMPTY_KEY; P.hbest[i] = 0u;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) ivm_reset_step_globals(P.g);
}

// ------------------------------------------------------------------ K1: prep + scatter
// Head of every ingest CTA: decide the env's reset / store origin locally, get the pose matrices
// (derived here from the angles, or loaded), and -- in the env's first CTA only -- publish the
// new env state for the kernels that follow.
struct K1Shared {
    float T[12];
    float cs[2];
    int32_t origin_r, origin_c, reset;
    int32_t bb[4];
    unsigned valid;
};

__device__ __forceinline__ void k1_prologue(const IvmParams &P, int b, K1Shared &sh, bool publisher) {
    if (threadIdx.x == 0) {
        const IvmEnvPrep q = ivm_env_decide(P, b);
        sh.origin_r = q.origin_r; sh.origin_c = q.origin_c; sh.reset = q.reset;
        sh.bb[0] = INT32_MAX; sh.bb[1] = INT32_MIN; sh.bb[2] = INT32_MAX; sh.bb[3] = INT32_MIN;
        sh.valid = 0;
    }
    if (P.orient != nullptr) {
        if (threadIdx.x == 32) ivm_pose_matrices(P, b, sh.T, sh.cs);
    } else if (threadIdx.x >= 32 && threadIdx.x < 44) {
        sh.T[threadIdx.x - 32] = P.T12[12 * b + threadIdx.x - 32];
    }
    __syncthreads();
    if (publisher) {
        IvmEnvPrep q;
        q.reset = sh.reset; q.origin_r = sh.origin_r; q.origin_c = sh.origin_c;
        ivm_env_publish<IvmAtomics>(P, b, q, threadIdx.x, blockDim.x);
        if (P.orient != nullptr) {
            if (threadIdx.x < 12) P.T12_buf[12 * b + threadIdx.x] = sh.T[threadIdx.x];
            if (threadIdx.x < 2) P.cs_buf[2 * b + threadIdx.x] = sh.cs[threadIdx.x];
        }
    }
}

// unproject VEC pixels of one image row, scatter their candidates, fold the frame bbox
template <int VEC>
__device__ __forceinline__ void k1_scatter_pixels(const IvmParams &P, int b, int pix0, const float *d, K1Shared &sh) {
    int rmin = INT32_MAX, rmax = INT32_MIN, cmin = INT32_MAX, cmax = INT32_MIN;
    unsigned nvalid = 0;
    if (pix0 < P.HW) {
        const float h = P.pose[3 * b + 1];
        const int v = pix0 / P.W, u0 = pix0 - v * P.W;
        const float ysv = P.ys[v];
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            IvmPoint p;
            const int ok = ivm_unproject(d[j], P.xs[u0 + j], ysv, sh.T, h, P.half_res, P.inv_half_res, p);
            if (ok == 0) continue;
            size_t idx;
            if (ok == 2 || !ivm_store_index(P, sh.origin_r, sh.origin_c, b, p.r, p.c, idx)) {
                atomicOr(&P.g->err, IVM_ERR_STORE_OVERFLOW);
                continue;
            }
            ivm_cand_insert<IvmAtomics>(P, b, (uint32_t)(idx - (size_t)b * P.SR * P.SC), ivm_cand_key(P, p.y, (uint32_t)(pix0 + j)));
            rmin = min(rmin, p.r); rmax = max(rmax, p.r); cmin = min(cmin, p.c); cmax = max(cmax, p.c);
            ++nvalid;
        }
    }
    // frame bbox over ALL envs (the reference subtracts batch-global minima, mapper.py:465)
    const unsigned wv = warp_sum(nvalid);
    if (wv) {  // warp-uniform
        rmin = warp_min(rmin); rmax = warp_max(rmax); cmin = warp_min(cmin); cmax = warp_max(cmax);
        if ((threadIdx.x & 31) == 0) {
            atomicMin(&sh.bb[0], rmin); atomicMax(&sh.bb[1], rmax); atomicMin(&sh.bb[2], cmin); atomicMax(&sh.bb[3], cmax);
            atomicAdd(&sh.valid, wv);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0 && sh.valid) {
        atomicMin(&P.g->loc[0], sh.bb[0]); atomicMax(&P.g->loc[1], sh.bb[1]);
        atomicMin(&P.g->loc[2], sh.bb[2]); atomicMax(&P.g->loc[3], sh.bb[3]);
        atomicAdd(&P.g->acc_valid, (unsigned long long)sh.valid);
    }
}

__device__ __forceinline__ float4 ld_cs_v4(const float *p) {  // volatile: keeps the batch of loads together
    float4 v;
    asm volatile("ld.global.cs.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

#define IVM_ARGMAX_STEP(v, k, best, arg) \
    if (((v) > (best)) || ((v) != (v) && (best) == (best))) { (best) = (v); (arg) = (k); }

// One thread = VEC consecutive pixels of one image row (128-bit loads for VEC=4).
template <bool PRED, int VEC>
__global__ void __launch_bounds__(IVM_THREADS)
k_ingest_scatter(IvmParams P, const float *__restrict__ logits, int ncls, uint8_t *__restrict__ labels_out) {
    const int b = blockIdx.y;
    __shared__ K1Shared sh;
    if (b >= P.B) {  // paused env (mapper.py:315-318): the first CTA wipes it, the rest have nothing to do
        if (blockIdx.x == 0) { IvmEnvPrep q; q.reset = 1; q.origin_r = 0; q.origin_c = 0; ivm_env_publish<IvmAtomics>(P, b, q, threadIdx.x, blockDim.x); }
        return;
    }
    const int pix0 = (blockIdx.x * IVM_THREADS + threadIdx.x) * VEC;
    // issue the depth load before the prologue's barrier
    float d[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) d[j] = 2.0f;
    const size_t base = (size_t)b * P.HW + pix0;
    if (pix0 < P.HW) {
        if (VEC == 4) {
            const float4 v = ld_stream4(P.depth + base);
            d[0] = v.x; d[1 % VEC] = v.y; d[2 % VEC] = v.z; d[3 % VEC] = v.w;
        } else {
            d[0] = P.depth[base];
        }
    }
    k1_prologue(P, b, sh, blockIdx.x == 0);
    if (PRED && pix0 < P.HW) {
        // PredictSemantics tail (mapper.py:795-798): argmax over class planes, first max wins,
        // NaN counts as maximal (torch.argmax).  Planes are streamed with evict-first loads,
        // 8 independent 128-bit loads in flight per thread.
        const float *lp = logits + (size_t)b * ncls * P.HW + pix0;
        float best[VEC];
        int arg[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) { best[j] = 0.f; arg[j] = 0; }
        if (VEC == 4) {
            const float4 v = ld_cs_v4(lp);
            best[0] = v.x; best[1 % VEC] = v.y; best[2 % VEC] = v.z; best[3 % VEC] = v.w;
        } else {
            best[0] = __ldcs(lp);
        }
        int k = 1;
        constexpr int U = 8;
        for (; k + U <= ncls; k += U) {
            if (VEC == 4) {
                float4 vals[U];
#pragma unroll
                for (int q = 0; q < U; ++q) vals[q] = ld_cs_v4(lp + (size_t)(k + q) * P.HW);
#pragma unroll
                for (int q = 0; q < U; ++q) {
                    IVM_ARGMAX_STEP(vals[q].x, k + q, best[0], arg[0]);
                    IVM_ARGMAX_STEP(vals[q].y, k + q, best[1 % VEC], arg[1 % VEC]);
                    IVM_ARGMAX_STEP(vals[q].z, k + q, best[2 % VEC], arg[2 % VEC]);
                    IVM_ARGMAX_STEP(vals[q].w, k + q, best[3 % VEC], arg[3 % VEC]);
                }
            } else {
                float vals[U];
#pragma unroll
                for (int q = 0; q < U; ++q) vals[q] = __ldcs(lp + (size_t)(k + q) * P.HW);
#pragma unroll
                for (int q = 0; q < U; ++q) { IVM_ARGMAX_STEP(vals[q], k + q, best[0], arg[0]); }
            }
        }
        for (; k < ncls; ++k) {
            if (VEC == 4) {
                const float4 v = ld_cs_v4(lp + (size_t)k * P.HW);
                IVM_ARGMAX_STEP(v.x, k, best[0], arg[0]);
                IVM_ARGMAX_STEP(v.y, k, best[1 % VEC], arg[1 % VEC]);
                IVM_ARGMAX_STEP(v.z, k, best[2 % VEC], arg[2 % VEC]);
                IVM_ARGMAX_STEP(v.w, k, best[3 % VEC], arg[3 % VEC]);
            } else {
                const float v = __ldcs(lp + (size_t)k * P.HW);
                IVM_ARGMAX_STEP(v, k, best[0], arg[0]);
            }
        }
        if (VEC == 4) {
            uchar4 o;
            o.x = (uint8_t)arg[0]; o.y = (uint8_t)arg[1 % VEC]; o.z = (uint8_t)arg[2 % VEC]; o.w = (uint8_t)arg[3 % VEC];
            *reinterpret_cast<uchar4 *>(labels_out + base) = o;
        } else {
            labels_out[base] = (uint8_t)arg[0];
        }
    }
    k1_scatter_pixels<VEC>(P, b, pix0, d, sh);
}

// ---- persistent bulk-async ingest of the predicted-semantics path (the default when the image
// tiles evenly).  2 CTAs per SM stay resident and walk a contiguous range of 512-pixel tiles; the
// score planes are staged through an 8-deep, 64 KB shared-memory ring by cp.async.bulk (the TMA
// engine's 1-D bulk copy, SASS UBLKCP) completing on mbarriers.  The ring never drains between
// tiles, so the loads of the next tile are in flight while a tile's labels are written and its
// points are scattered: ~128 KB per SM outstanding at all times, no registers held by loads, no
// wave tail.  Threads read their 2 pixels per plane from shared memory (LDS.64).
#define IVM_BULK_TILE 512    // pixels per tile (= 256 threads x 2)
#define IVM_BULK_SP 4        // planes per stage (8 KB)
#define IVM_BULK_NSTAGE 8    // ring depth (64 KB)
#define IVM_BULK_CTAS_PER_SM 2

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
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

__global__ void __launch_bounds__(IVM_THREADS, IVM_BULK_CTAS_PER_SM)
k_ingest_scatter_bulk(IvmParams P, const float *__restrict__ logits, int ncls, uint8_t *__restrict__ labels_out,
                      int nenv_total) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float(*ring)[IVM_BULK_SP][IVM_BULK_TILE] = reinterpret_cast<float(*)[IVM_BULK_SP][IVM_BULK_TILE]>(smem_raw);
    __shared__ __align__(8) uint64_t full[IVM_BULK_NSTAGE];
    __shared__ K1Shared sh;
    const int tpe = P.HW / IVM_BULK_TILE;                 // tiles per env
    const long long total = (long long)P.B * tpe;
    const int t0 = (int)((long long)blockIdx.x * total / gridDim.x);
    const int t1 = (int)((long long)(blockIdx.x + 1) * total / gridDim.x);
    const int nchunks = (ncls + IVM_BULK_SP - 1) / IVM_BULK_SP;
    const int my_chunks = (t1 - t0) * nchunks;
    uint64_t policy = 0;
    auto issue = [&](int q) {  // thread 0 only
        const int tile = t0 + q / nchunks, ch = q - (q / nchunks) * nchunks;
        const int eb = tile / tpe, tp0 = (tile - eb * tpe) * IVM_BULK_TILE;
        const int slot = q % IVM_BULK_NSTAGE;
        const int p0 = ch * IVM_BULK_SP;
        const int np = min(IVM_BULK_SP, ncls - p0);
        const float *src = logits + ((size_t)eb * ncls + p0) * P.HW + tp0;
        mbar_expect_tx(&full[slot], (uint32_t)(np * IVM_BULK_TILE * sizeof(float)));
        for (int p = 0; p < np; ++p)
            bulk_g2s(&ring[slot][p][0], src + (size_t)p * P.HW, IVM_BULK_TILE * sizeof(float), &full[slot], policy);
    };
    if (threadIdx.x == 0) {
        for (int s = 0; s < IVM_BULK_NSTAGE; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
        for (int q = 0; q < IVM_BULK_NSTAGE && q < my_chunks; ++q) issue(q);
    }
    // paused envs (mapper.py:315-318) are wiped by the last CTA
    if (blockIdx.x == gridDim.x - 1)
        for (int b = P.B; b < nenv_total; ++b) {
            IvmEnvPrep q; q.reset = 1; q.origin_r = 0; q.origin_c = 0;
            ivm_env_publish<IvmAtomics>(P, b, q, threadIdx.x, blockDim.x);
        }
    int rmin = INT32_MAX, rmax = INT32_MIN, cmin = INT32_MAX, cmax = INT32_MIN;
    unsigned nvalid = 0;
    int cur_env = -1;
    __syncthreads();  // barriers initialised
    for (int tile = t0; tile < t1; ++tile) {
        const int b = tile / tpe, tp0 = (tile - b * tpe) * IVM_BULK_TILE;
        const int pix0 = tp0 + threadIdx.x * 2;
        const size_t base = (size_t)b * P.HW + pix0;
        const float2 dv = __ldcs(reinterpret_cast<const float2 *>(P.depth + base));
        if (b != cur_env) {  // block-uniform: at most twice per CTA (its tile range is contiguous)
            __syncthreads();  // the previous env's matrices are no longer read
            if (threadIdx.x == 0) {
                const IvmEnvPrep q = ivm_env_decide(P, b);
                sh.origin_r = q.origin_r; sh.origin_c = q.origin_c; sh.reset = q.reset;
            }
            if (P.orient != nullptr) {
                if (threadIdx.x == 32) ivm_pose_matrices(P, b, sh.T, sh.cs);
            } else if (threadIdx.x >= 32 && threadIdx.x < 44) {
                sh.T[threadIdx.x - 32] = P.T12[12 * b + threadIdx.x - 32];
            }
            __syncthreads();
            cur_env = b;
        }
        if (tp0 == 0) {  // the CTA that owns an env's first tile publishes the env's new state
            IvmEnvPrep q;
            q.reset = sh.reset; q.origin_r = sh.origin_r; q.origin_c = sh.origin_c;
            ivm_env_publish<IvmAtomics>(P, b, q, threadIdx.x, blockDim.x);
            if (P.orient != nullptr) {
                if (threadIdx.x < 12) P.T12_buf[12 * b + threadIdx.x] = sh.T[threadIdx.x];
                if (threadIdx.x < 2) P.cs_buf[2 * b + threadIdx.x] = sh.cs[threadIdx.x];
            }
        }
        // PredictSemantics tail (mapper.py:795-798): running argmax over the planes, first max wins,
        // NaN counts as maximal (torch.argmax)
        float best0 = 0.f, best1 = 0.f;
        int a0 = 0, a1 = 0;
        for (int ch = 0; ch < nchunks; ++ch) {
            const int q = (tile - t0) * nchunks + ch;
            const int slot = q % IVM_BULK_NSTAGE;
            mbar_wait(&full[slot], (uint32_t)((q / IVM_BULK_NSTAGE) & 1));
            const int p0 = ch * IVM_BULK_SP;
            const int np = min(IVM_BULK_SP, ncls - p0);
#pragma unroll
            for (int p = 0; p < IVM_BULK_SP; ++p) {
                if (p < np) {
                    const float2 v = *reinterpret_cast<const float2 *>(&ring[slot][p][threadIdx.x * 2]);
                    const int k = p0 + p;
                    if (k == 0) {
                        best0 = v.x; best1 = v.y;
                    } else {
                        IVM_ARGMAX_STEP(v.x, k, best0, a0);
                        IVM_ARGMAX_STEP(v.y, k, best1, a1);
                    }
                }
            }
            __syncthreads();  // every thread is done with this slot
            if (threadIdx.x == 0 && q + IVM_BULK_NSTAGE < my_chunks) issue(q + IVM_BULK_NSTAGE);
        }
        uchar2 o;
        o.x = (uint8_t)a0; o.y = (uint8_t)a1;
        *reinterpret_cast<uchar2 *>(labels_out + base) = o;
        // unproject + scatter this thread's two pixels (the next tile's planes are already in flight)
        const float h = P.pose[3 * b + 1];
        const int v = pix0 / P.W, u0 = pix0 - v * P.W;
        const float ysv = P.ys[v];
        const float dd[2] = {dv.x, dv.y};
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            IvmPoint p;
            const int ok = ivm_unproject(dd[j], P.xs[u0 + j], ysv, sh.T, h, P.half_res, P.inv_half_res, p);
            if (ok == 0) continue;
            size_t idx;
            if (ok == 2 || !ivm_store_index(P, sh.origin_r, sh.origin_c, b, p.r, p.c, idx)) {
                atomicOr(&P.g->err, IVM_ERR_STORE_OVERFLOW);
                continue;
            }
            ivm_cand_insert<IvmAtomics>(P, b, (uint32_t)(idx - (size_t)b * P.SR * P.SC), ivm_cand_key(P, p.y, (uint32_t)(pix0 + j)));
            rmin = min(rmin, p.r); rmax = max(rmax, p.r); cmin = min(cmin, p.c); cmax = max(cmax, p.c);
            ++nvalid;
        }
    }
    // frame bbox over ALL envs (mapper.py:465), one flush per CTA
    __syncthreads();
    if (threadIdx.x == 0) { sh.bb[0] = INT32_MAX; sh.bb[1] = INT32_MIN; sh.bb[2] = INT32_MAX; sh.bb[3] = INT32_MIN; sh.valid = 0; }
    __syncthreads();
    const unsigned wv = warp_sum(nvalid);
    if (wv) {
        rmin = warp_min(rmin); rmax = warp_max(rmax); cmin = warp_min(cmin); cmax = warp_max(cmax);
        if ((threadIdx.x & 31) == 0) {
            atomicMin(&sh.bb[0], rmin); atomicMax(&sh.bb[1], rmax); atomicMin(&sh.bb[2], cmin); atomicMax(&sh.bb[3], cmax);
            atomicAdd(&sh.valid, wv);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0 && sh.valid) {
        atomicMin(&P.g->loc[0], sh.bb[0]); atomicMax(&P.g->loc[1], sh.bb[1]);
        atomicMin(&P.g->loc[2], sh.bb[2]); atomicMax(&P.g->loc[3], sh.bb[3]);
        atomicAdd(&P.g->acc_valid, (unsigned long long)sh.valid);
    }
}

// ------------------------------------------------------------------ K2: resolve
// Only ~1 pixel in 4 survives the depth/height filters, scattered over the image, so the filters
// are evaluated for all pixels first (cheap) and the survivors are compacted per warp through a
// shared-memory queue; the expensive part (two IEEE divisions, candidate check, world-record
// read-modify-write) then runs on dense warps.
template <int VEC>
__global__ void __launch_bounds__(IVM_THREADS) k_ingest_resolve(IvmParams P) {
    const int b = blockIdx.y;
    const int pix0 = (blockIdx.x * IVM_THREADS + threadIdx.x) * VEC;
    __shared__ float sT[12];
    __shared__ int32_t sloc[4];
    __shared__ int32_t sbox[5];   // rmin, rmax, cmin, cmax, n of the cells this CTA newly occupied
    __shared__ unsigned s_local;
    __shared__ uint32_t q_pix[IVM_THREADS / 32][32 * VEC];
    __shared__ float q_d[IVM_THREADS / 32][32 * VEC];
    __shared__ uint8_t q_lab[IVM_THREADS / 32][32 * VEC];
    if (threadIdx.x < 12) sT[threadIdx.x] = P.T12[12 * b + threadIdx.x];
    if (threadIdx.x >= 32 && threadIdx.x < 36) sloc[threadIdx.x - 32] = P.g->loc[threadIdx.x - 32];
    if (threadIdx.x == 64) { sbox[0] = INT32_MAX; sbox[1] = INT32_MIN; sbox[2] = INT32_MAX; sbox[3] = INT32_MIN; sbox[4] = 0; s_local = 0; }
    // issue the pixel loads before the barrier
    float d[VEC];
    uint8_t lab[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) { d[j] = 2.0f; lab[j] = 0; }
    if (pix0 < P.HW) {
        const size_t base = (size_t)b * P.HW + pix0;
        if (VEC == 4) {
            const float4 v = *reinterpret_cast<const float4 *>(P.depth + base);
            d[0] = v.x; d[1 % VEC] = v.y; d[2 % VEC] = v.z; d[3 % VEC] = v.w;
            const uchar4 l = *reinterpret_cast<const uchar4 *>(P.labels + base);
            lab[0] = l.x; lab[1 % VEC] = l.y; lab[2 % VEC] = l.z; lab[3 % VEC] = l.w;
        } else {
            d[0] = P.depth[base];
            lab[0] = P.labels[base];
        }
    }
    const IvmEnv *e = &P.env[b];
    const int32_t origin_r = e->origin_r, origin_c = e->origin_c;
    const float h = P.pose[3 * b + 1];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // pass 1: filters only (depth band, then the height of the transformed point), compact survivors
    int count = 0;
    {
        const int v = pix0 < P.HW ? pix0 / P.W : 0, u0 = pix0 - v * P.W;
        const float ysv = P.ys[v];
        const float hlo = ivm_sub(h, 1.0f), hhi = ivm_add(h, 0.5f);
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            bool ok = pix0 < P.HW && d[j] > 0.01f && d[j] < 0.99f;
            if (ok) {
                const float z = ivm_mul(d[j], 10.0f);
                float acc = ivm_mul(sT[4], ivm_mul(z, P.xs[u0 + j]));
                acc = ivm_fma(sT[5], ivm_mul(z, ysv), acc);
                acc = ivm_fma(sT[6], z, acc);
                acc = ivm_fma(sT[7], 1.0f, acc);
                ok = acc > hlo && acc < hhi;
            }
            const unsigned m = __ballot_sync(0xffffffffu, ok);
            if (ok) {
                const int pos = count + __popc(m & ((1u << lane) - 1u));
                q_pix[warp][pos] = (uint32_t)(pix0 + j); q_d[warp][pos] = d[j]; q_lab[warp][pos] = lab[j];
            }
            count += __popc(m);
        }
    }
    __syncwarp();
    // pass 2: dense -- one queued pixel per lane
    unsigned nlocal = 0;
    IvmBoxAcc acc;
    acc.clear();
    for (int i = lane; i < count; i += 32) {
        const uint32_t pix = q_pix[warp][i];
        const int v = (int)pix / P.W, u = (int)pix - v * P.W;
        IvmPoint p;
        if (ivm_unproject(q_d[warp][i], P.xs[u], P.ys[v], sT, h, P.half_res, P.inv_half_res, p) != 1) continue;
        nlocal += (unsigned)ivm_resolve_pixel<IvmAtomics>(P, b, pix, p, q_lab[warp][i], sloc, origin_r, origin_c, acc);
    }
    // newly occupied cells: warp -> block -> 5 global atomics per CTA
    const unsigned wn = warp_sum((unsigned)acc.n);
    if (wn) {
        const int r0 = warp_min(acc.rmin), r1 = warp_max(acc.rmax), c0 = warp_min(acc.cmin), c1 = warp_max(acc.cmax);
        if (lane == 0) {
            atomicMin(&sbox[0], r0); atomicMax(&sbox[1], r1); atomicMin(&sbox[2], c0); atomicMax(&sbox[3], c1);
            atomicAdd(&sbox[4], (int)wn);
        }
    }
    const unsigned wl = warp_sum(nlocal);
    if (wl && lane == 0) atomicAdd(&s_local, wl);
    __syncthreads();
    if (threadIdx.x == 0) {
        if (sbox[4] > 0) {
            IvmBoxAcc t;
            t.rmin = sbox[0]; t.rmax = sbox[1]; t.cmin = sbox[2]; t.cmax = sbox[3]; t.n = sbox[4];
            ivm_box_flush<IvmAtomics>(&P.env[b], t);
        }
        if (s_local) atomicAdd(&P.g->acc_local, (unsigned long long)s_local);
    }
}

// ------------------------------------------------------------------ K3: fix-up
#define IVM_FIX_SMALL 512
__global__ void __launch_bounds__(1024) k_fixup(const __grid_constant__ IvmParams P) {
    __shared__ unsigned long long s_key[IVM_FIX_SMALL], s_xo[IVM_FIX_SMALL], s_l[2];
    __shared__ uint32_t s_ord[IVM_FIX_SMALL];
    __shared__ int32_t s_i[8];
    IvmFixScratch S;
    S.key = s_key; S.xo = s_xo; S.ord = s_ord; S.cap = IVM_FIX_SMALL; S.ibuf = s_i; S.lbuf = s_l;
    S.release = nullptr; S.release_add = 0u;
    ivm_fixup_program<IvmAtomics>(P, S, threadIdx.x, blockDim.x);
    // (the persistent kernel's record of its last edge-line scan says nothing about a store this path has merged into)
    if (threadIdx.x == 0) P.g->scan_valid = 0u;
}

// ------------------------------------------------------------------ K4: raster
// One thread group (a whole CTA in the stand-alone kernel, a 128-thread half of the CTA in the
// fused step kernel) = one ego tile of one env (output-stationary).  The store half-rows under the
// rotated tile are cut into 32-record chunks; warps take chunks round-robin, IVM_RASTER_MLP at a
// time, so that every lane has that many independent 16-byte loads in flight.
#define IVM_RASTER_MLP 2   // (measured: 2 beats 4 by 0.4 us per step, 1 loses 1-4 us, 8 spills)

// named barrier with a compile-time id (a register id would make ptxas reserve all 16 barriers)
__device__ __forceinline__ void group_bar(int bar_id, int nthr) {
    switch (bar_id) {
        case 0: asm volatile("bar.sync 0, %0;" ::"r"(nthr) : "memory"); break;
        case 1: asm volatile("bar.sync 1, %0;" ::"r"(nthr) : "memory"); break;
        case 2: asm volatile("bar.sync 2, %0;" ::"r"(nthr) : "memory"); break;
        case 3: asm volatile("bar.sync 3, %0;" ::"r"(nthr) : "memory"); break;
        case 4: asm volatile("bar.sync 4, %0;" ::"r"(nthr) : "memory"); break;
        case 6: asm volatile("bar.sync 6, %0;" ::"r"(nthr) : "memory"); break;
        default: asm volatile("bar.sync 5, %0;" ::"r"(nthr) : "memory"); break;
    }
}

// MLP records of one lane together: the common path is branch-free (the rare ambiguous quotient, ivm_rint_mul,
// sends all of them through the true division once), so that the independent chains interleave.
template <int MLP, bool KNOWN>
__device__ __forceinline__ void raster_records(const IvmParams &P, const uint4 *raw, const bool *have, uint32_t rs1, float px, float pz,
                                               float c, float s, float ylo, float yhi, float fr0, float fr1, float fc0, float fc1,
                                               int r0, int c0, int tc, const uint32_t *cell, uint32_t *skey, uint8_t *socc, unsigned &n_in) {
    float ar[MLP], ac[MLP], rf[MLP], cf[MLP];
    bool amb = false;
#pragma unroll
    for (int u = 0; u < MLP; ++u) {
        ivm_ego_numerators(P, __uint_as_float(raw[u].x), __uint_as_float(raw[u].z), px, pz, c, s, ar[u], ac[u]);
        rf[u] = ivm_rint_mul(ar[u], P.inv_res, amb);
        cf[u] = ivm_rint_mul(ac[u], P.inv_res, amb);
    }
    if (amb) {
#pragma unroll
        for (int u = 0; u < MLP; ++u) { rf[u] = rintf(ivm_div(ar[u], P.res)); cf[u] = rintf(ivm_div(ac[u], P.res)); }
    }
#pragma unroll
    for (int u = 0; u < MLP; ++u) {
        const float y = __uint_as_float(raw[u].y);
        bool ok = have[u] && y > ylo && y < yhi && rf[u] >= fr0 && rf[u] < fr1 && cf[u] >= fc0 && cf[u] < fc1;
        if (!KNOWN) ok = ok && (raw[u].w >> 8) >= rs1;   // live: written since the env's last reset (rs1 >= 1: never-written cells fail)
        if (ok) {
            ++n_in;
            const int t = ((int)rf[u] - r0) * tc + ((int)cf[u] - c0);
            socc[t] = 1;  // OccupancyStatus.OCCUPIED
            const uint32_t label = raw[u].w & 0xFFu;
            // last point in list order wins (mapper.py:569-571): list order within an env is (half-row, half-col)
            // lexicographic / the npz index in known mode; labels 0 are excluded (mapper.py:611)
            if (label) atomicMax(&skey[t], KNOWN ? raw[u].w : ((cell[u] << 8) | label));
        }
    }
}

__device__ __forceinline__ void raster_record(const IvmParams &P, const uint4 raw, bool have, uint32_t reset_stamp, float px,
                                              float h, float pz, float c, float s, int r0, int r1, int c0, int c1, int tc,
                                              uint32_t cell, uint32_t *skey, uint8_t *socc, unsigned &n_in) {
    int row, col;
    bool ok = ivm_ego_cell(P, __uint_as_float(raw.x), __uint_as_float(raw.y), __uint_as_float(raw.z), px, h, pz, c, s, row, col);
    ok = ok && have && ivm_live(raw.w, reset_stamp) && row >= r0 && row < r1 && col >= c0 && col < c1;
    if (ok) {
        ++n_in;
        const int t = (row - r0) * tc + (col - c0);
        socc[t] = 1;  // OccupancyStatus.OCCUPIED
        const uint32_t label = raw.w & 0xFFu;
        // last point in list order wins (mapper.py:569-571); list order within an env is (half-row,
        // half-col) lexicographic; labels 0 are excluded (mapper.py:611)
        if (label) atomicMax(&skey[t], (cell << 8) | label);
    }
}

// bytes of shared memory one raster group needs: keys + occupancy of the tile, the span table, and (stage_cap > 0)
// a staging area of stage_cap records + the per-row offsets into it
static inline size_t raster_fixed_bytes(int tile_r, int tile_c, int max_rows) {
    return (((size_t)tile_r * tile_c * 5 + (size_t)max_rows * 8 + 16) + 15) & ~(size_t)15;
}
static inline size_t raster_smem_bytes(int tile_r, int tile_c, int max_rows, int stage_cap = 0) {
    size_t n = raster_fixed_bytes(tile_r, tile_c, max_rows);
    if (stage_cap > 0) n += (((size_t)(max_rows + 1) * 4 + 15) & ~(size_t)15) + (size_t)stage_cap * 16;
    return n;
}
__device__ __forceinline__ size_t raster_fixed_bytes_dev(int tile_r, int tile_c, int max_rows) {
    return (((size_t)tile_r * tile_c * 5 + (size_t)max_rows * 8 + 16) + 15) & ~(size_t)15;
}

#define RTRACE_DECL unsigned long long rt_[4] = {0ull, 0ull, 0ull, 0ull}
#define RTRACE(k)                                                                                          \
    do {                                                                                                   \
        if (!KNOWN && tid == 0 && bar_id == 2 && blockIdx.x < IVM_TRACE_CTAS) {                            \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(rt_[k]));                                     \
            if ((k) > 0) P.cta_trace[(size_t)blockIdx.x * IVM_TRACE_SLOTS + 12 + (k)] += rt_[k] - rt_[(k) - 1]; \
        }                                                                                                  \
    } while (0)
// safe_only: the tile is rastered only if nothing under it can still be changed by the edge fix-up, i.e. no
// frame-edge winner is pending in it (tile_dirty stamp) and its store footprint stays clear of the edge lines of
// the batch-global bounding box, where the stage-2 collisions live.  Those lines are known (as narrow bands,
// `gband`, see ovl_global_bands) as soon as grid barrier 2 has passed.  Returns false if the tile was skipped.
// stage_cap > 0: the records under the tile are first copied into shared memory with cp.async (every copy of the
// tile in flight at once, no registers held), so that a tile costs ONE memory round trip instead of one per
// batch of rows; tiles with more records than stage_cap take the direct path.
template <bool KNOWN>
__device__ __forceinline__ bool raster_tile(const IvmParams &P, int max_rows, int b, int r0, int c0, uint32_t *smem, int tid,
                                            int nthr, int bar_id, unsigned &n_in, bool safe_only = false, int stage_cap = 0,
                                            const int32_t *gband = nullptr) {
    const int tr = P.tile_r, tc = P.tile_c;
    uint32_t *skey = smem;
    int32_t *s_clo = reinterpret_cast<int32_t *>(skey + tr * tc);   // first store column of each half-row's span
    int32_t *s_len = s_clo + max_rows;                              // span length (0 = nothing to read)
    int32_t *s_maxlen = s_len + max_rows;
    uint8_t *socc = reinterpret_cast<uint8_t *>(s_maxlen + 4);
    const int r1 = min(r0 + tr, P.R), c1 = min(c0 + tc, P.C);
    RTRACE_DECL;
    group_bar(bar_id, nthr);  // the group's previous tile has been written out
    RTRACE(0);
    for (int i = tid; i < tr * tc; i += nthr) { skey[i] = 0u; socc[i] = 0; }
    if (tid == 0) { s_maxlen[0] = 0; s_maxlen[1] = 0; }
    const IvmEnv e = P.env[b];
    const float px = P.pose[3 * b + 0], h = P.pose[3 * b + 1], pz = P.pose[3 * b + 2];
    const float c = P.cs[2 * b + 0], s = P.cs[2 * b + 1];
    IvmTileGeom G;
    ivm_tile_geom(P, px, pz, c, s, r0, r1, c0, c1, G);
    const int lane = tid & 31, warp = tid >> 5, nwarps = nthr >> 5;
    (void)lane;
    const uint32_t rs1 = max(e.reset_stamp, 1u);
    const float ylo = ivm_sub(h, 1.25f), yhi = ivm_add(h, 0.75f);   // FilterPointCloudByRobotHeight, mapper.py:884-901
    const float fr0 = (float)r0, fr1 = (float)r1, fc0 = (float)c0, fc1 = (float)c1;
    int row_lo = max(G.row_lo, KNOWN ? e.origin_r : e.rmin);
    int row_hi = min(G.row_hi, KNOWN ? e.origin_r + P.SR - 1 : e.rmax);
    if (e.count <= 0) row_hi = row_lo - 1;
    if (row_hi - row_lo + 1 > max_rows) row_hi = row_lo + max_rows - 1;  // cannot happen: max_rows bounds the tile diagonal
    const int nrows = row_hi - row_lo + 1;
    const int col_lo = KNOWN ? e.origin_c : e.cmin, col_hi = KNOWN ? e.origin_c + P.SC - 1 : e.cmax;
    group_bar(bar_id, nthr);
    // phase 1: one thread per half-row solves the column span under the rotated tile
    for (int i = tid; i < nrows; i += nthr) {
        int clo, chi;
        ivm_row_span(G, row_lo + i, clo, chi);
        clo = max(clo, col_lo); chi = min(chi, col_hi);
        const int len = chi >= clo ? chi - clo + 1 : 0;
        s_clo[i] = clo; s_len[i] = len;
        if (len > 0) atomicMax(s_maxlen, len);
        if (!KNOWN && safe_only && len > 0) {
            const int row = row_lo + i;
            if ((row >= gband[0] && row <= gband[1]) || (row >= gband[2] && row <= gband[3]) ||
                (chi >= gband[4] && clo <= gband[5]) || (chi >= gband[6] && clo <= gband[7]))
                s_maxlen[1] = 1;
        }
    }
    if (!KNOWN && safe_only && tid == 0) {
        const int tiles_x = (P.C + P.tile_c - 1) / P.tile_c, tiles_y = (P.R + P.tile_r - 1) / P.tile_r;
        if (__ldcg(&P.tile_dirty[(size_t)b * (tiles_x * tiles_y) + (r0 / P.tile_r) * tiles_x + c0 / P.tile_c]) == P.step) s_maxlen[1] = 1;
    }
    group_bar(bar_id, nthr);
    RTRACE(1);
    if (!KNOWN && safe_only && s_maxlen[1]) return false;  // uniform
    bool staged = false;
    if (!KNOWN && stage_cap > 0) {
        int32_t *s_off = reinterpret_cast<int32_t *>(reinterpret_cast<unsigned char *>(smem) + raster_fixed_bytes_dev(tr, tc, max_rows));
        uint4 *stage = reinterpret_cast<uint4 *>(s_off + ((max_rows + 1 + 3) & ~3));
        if (warp == 0) {  // exclusive prefix of the span lengths
            int carry = 0;
            for (int i0 = 0; i0 < nrows; i0 += 32) {
                const int i = i0 + lane;
                const int len = i < nrows ? s_len[i] : 0;
                int x = len;
                for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
                if (i < nrows) s_off[i] = carry + x - len;
                carry += __shfl_sync(0xffffffffu, x, 31);
            }
            if (lane == 0) s_maxlen[2] = carry;
        }
        group_bar(bar_id, nthr);
        const int total = s_maxlen[2];
        if (total <= stage_cap) {
            staged = true;
            const IvmRecord *env_store = P.store + (size_t)b * P.SR * P.SC;
            const int hw = tid >> 4, nhw = nthr >> 4, l16 = tid & 15;
            const unsigned long long pol = ivm_policy_keep();
            for (int i = hw; i < nrows; i += nhw) {
                const int len = s_len[i];
                const IvmRecord *src = env_store + (size_t)(row_lo + i - e.origin_r) * P.SC + (size_t)(s_clo[i] - e.origin_c);
                uint4 *dst = stage + s_off[i];
                for (int k = l16; k < len; k += 16)
                    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(smem_u32(dst + k)), "l"(src + k), "l"(pol) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            group_bar(bar_id, nthr);
            // every record of the tile is in shared memory: one half-warp per half-row, as in the direct path
            for (int i = hw; i < nrows; i += nhw) {
                const int len = s_len[i];
                const uint32_t rowbase = (uint32_t)(row_lo + i - e.origin_r) * (uint32_t)P.SC + (uint32_t)(s_clo[i] - e.origin_c);
                const uint4 *srow = stage + s_off[i];
                for (int k0 = 0; k0 < len; k0 += 16 * IVM_RASTER_MLP) {
                    uint4 raw[IVM_RASTER_MLP];
                    bool have[IVM_RASTER_MLP];
#pragma unroll
                    for (int u = 0; u < IVM_RASTER_MLP; ++u) {
                        const int off = k0 + 16 * u + l16;
                        have[u] = off < len;
                        raw[u] = have[u] ? srow[off] : make_uint4(0, 0, 0, 0);
                    }
#pragma unroll
                    for (int u = 0; u < IVM_RASTER_MLP; ++u)
                        raster_record(P, raw[u], have[u], e.reset_stamp, px, h, pz, c, s, r0, r1, c0, c1, tc,
                                      rowbase + (uint32_t)(k0 + 16 * u + l16), skey, socc, n_in);
                }
            }
        }
    }
    if (KNOWN || staged) {
    } else if (!KNOWN) {
        // phase 2: one half-warp per store half-row (a span under a 16x16 tile holds ~40 records);
        // its 16 lanes read up to IVM_RASTER_MLP x 16 consecutive records (independent 16-byte
        // loads, 256 contiguous bytes per half-warp and load) before any of them is processed
        const IvmRecord *env_store = P.store + (size_t)b * P.SR * P.SC;
        const int hw = tid >> 4, nhw = nthr >> 4, l16 = tid & 15;
        // pull the whole footprint of the tile towards L2 first (one prefetch per 128-byte line)
        for (int j = tid; j < nrows * 8; j += nthr) {
            const int i = j >> 3, seg = j & 7;  // a span of <= 64 records covers <= 9 lines; longer spans are only partly prefetched
            if (seg * 8 < s_len[i]) {
                const size_t first = (size_t)(row_lo + i - e.origin_r) * P.SC + (size_t)(s_clo[i] - e.origin_c);
                asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(env_store + first + seg * 8));
            }
        }
        // the two half-warps of a warp take neighbouring half-rows and stay in step (warp-uniform loops), so that
        // a 16-record chunk neither of them has is skipped by the whole warp: spans under a rotated tile are
        // 1..45 records long, and a chunk costs ~80 instructions per lane whether its lanes hold records or not.
        // (Measured: prefetching the next rows' records while a batch is evaluated does not pay inside the
        // 96-register budget of the persistent kernel -- the loop is bound by its dependent arithmetic.)
        (void)hw; (void)nhw;
        // (Measured: two pairs of half-rows per iteration -- 8 record loads in flight per lane -- spills inside the 72-register
        // budget of the persistent kernel and loses: record loop 7.4 -> 8.1 us.)
        for (int i0 = 2 * warp; i0 < nrows; i0 += 2 * nwarps) {
            const int i = i0 + (lane >> 4);
            const int len = i < nrows ? s_len[i] : 0;
            const int lenmax = max(len, __shfl_xor_sync(0xffffffffu, len, 16));
            const uint32_t rowbase = len > 0 ? (uint32_t)(row_lo + i - e.origin_r) * (uint32_t)P.SC + (uint32_t)(s_clo[i] - e.origin_c) : 0u;
            for (int k0 = 0; k0 < lenmax; k0 += 16 * IVM_RASTER_MLP) {
                uint4 raw[IVM_RASTER_MLP];
                bool have[IVM_RASTER_MLP];
                uint32_t cellv[IVM_RASTER_MLP];
#pragma unroll
                for (int u = 0; u < IVM_RASTER_MLP; ++u) {
                    const int off = k0 + 16 * u + l16;
                    have[u] = off < len;
                    cellv[u] = rowbase + (uint32_t)off;
                    raw[u] = make_uint4(0, 0, 0, 0);
                    // L2-only load: in the persistent kernel the records were written earlier in the same launch
                    if (have[u]) {
                        const IvmRecord q = ivm_load_record(env_store + rowbase + off);
                        raw[u] = make_uint4(__float_as_uint(q.x), __float_as_uint(q.y), __float_as_uint(q.z), q.meta);
                    }
                }
                raster_records<IVM_RASTER_MLP, false>(P, raw, have, rs1, px, pz, c, s, ylo, yhi, fr0, fr1, fc0, fc1, r0, c0, tc, cellv,
                                                      skey, socc, n_in);
            }
        }
    }
    if (KNOWN) {
        for (int i = warp; i < nrows; i += nwarps) {
            if (s_len[i] <= 0) continue;
            const int rr = row_lo + i, clo = s_clo[i], chi = s_clo[i] + s_len[i] - 1;
            const uint32_t *off = P.koff + (size_t)b * ((size_t)P.SR * P.SC + 1) + (size_t)(rr - e.origin_r) * P.SC;
            const uint32_t p0 = off[clo - e.origin_c], p1 = off[chi - e.origin_c + 1];
            const IvmRecord *pts = P.kpts + (size_t)b * P.kcap;
            for (uint32_t q = p0 + lane; q < p1; q += 32) {
                const uint4 raw = __ldg(reinterpret_cast<const uint4 *>(pts + q));
                int row, col;
                if (!ivm_ego_cell(P, __uint_as_float(raw.x), __uint_as_float(raw.y), __uint_as_float(raw.z), px, h, pz, c,
                                  s, row, col))
                    continue;
                if (row < r0 || row >= r1 || col < c0 || col >= c1) continue;
                ++n_in;
                const int t = (row - r0) * tc + (col - c0);
                socc[t] = 1;
                // raw.w = (npz index << 8) | label: later points overwrite earlier ones
                if (raw.w & 0xFFu) atomicMax(&skey[t], raw.w);
            }
        }
    }
    group_bar(bar_id, nthr);
    RTRACE(2);
    const int wr = r1 - r0, wc = c1 - c0;
    for (int i = tid; i < wr * wc; i += nthr) {
        const int rr = i / wc, cc = i - rr * wc;
        const size_t o = ((size_t)b * P.R + (size_t)(r0 + rr)) * P.C + (size_t)(c0 + cc);
        P.occ[o] = socc[rr * tc + cc];
        P.sem[o] = (uint8_t)(skey[rr * tc + cc] & 0xFFu);
    }
    RTRACE(3);
    return true;
}

template <bool KNOWN>
__global__ void __launch_bounds__(IVM_RASTER_THREADS) k_raster(IvmParams P, int max_rows) {
    extern __shared__ __align__(16) uint32_t raster_smem[];
    unsigned n_in = 0;
    raster_tile<KNOWN>(P, max_rows, blockIdx.z, blockIdx.y * P.tile_r, blockIdx.x * P.tile_c, raster_smem, threadIdx.x,
                       blockDim.x, 0, n_in);
    const unsigned wn = warp_sum(n_in);
    if (wn && (threadIdx.x & 31) == 0) atomicAdd(&P.g->stats[IVM_STAT_IN], (unsigned long long)wn);
}

// Known-map raster (BASELINE config 5): one 128-thread CTA per ego tile of one env, over the env's CSR store (points
// sorted by half-cell).  The step is latency-bound (a few hundred KB per env), so everything is arranged for few
// dependent round trips: (1) one thread per store half-row solves its column span and reads the two CSR offsets;
// (2) a scan turns the row counts into one flat index space; (3) every thread then loads IVM_KNOWN_MLP points that
// are independent of each other before it evaluates any (flat index -> row by binary search in shared memory).
// The whole known-map step is this ONE kernel: thread 0 of a CTA derives (cos, sin)(-heading) from the angles itself
// (mapper.py:38-48, in the angles' dtype) while the others clear the tile.
#define IVM_KNOWN_MLP 8
__global__ void __launch_bounds__(IVM_RASTER_THREADS, 8) k_raster_known(const __grid_constant__ IvmParams P, int max_rows, int parity) {
    extern __shared__ __align__(16) uint32_t ksm[];
    const int tr = P.tile_r, tc = P.tile_c, tid = threadIdx.x, nthr = blockDim.x;
    uint32_t *skey = ksm;
    uint32_t *s_p0 = skey + tr * tc;
    int32_t *s_off = reinterpret_cast<int32_t *>(s_p0 + max_rows);   // [max_rows + 1] exclusive prefix of the row counts
    uint8_t *socc = reinterpret_cast<uint8_t *>(s_off + max_rows + 1);
    const int b = blockIdx.z, r0 = blockIdx.y * tr, c0 = blockIdx.x * tc;
    const int r1 = min(r0 + tr, P.R), c1 = min(c0 + tc, P.C);
    __shared__ float s_cs[2];
    if (tid == 0) {
        if (P.orient != nullptr) {
            float T[12];
            ivm_pose_matrices(P, b, T, s_cs);
        } else {
            s_cs[0] = P.cs[2 * b + 0]; s_cs[1] = P.cs[2 * b + 1];
        }
        if (blockIdx.x == 0 && blockIdx.y == 0 && b == 0) P.g->known_in[parity ^ 1] = 0ull;  // the next step's counter
    }
    for (int i = tid; i < tr * tc; i += nthr) { skey[i] = 0u; socc[i] = 0; }
    const IvmEnv e = P.env[b];
    const float px = P.pose[3 * b + 0], h = P.pose[3 * b + 1], pz = P.pose[3 * b + 2];
    __syncthreads();
    const float c = s_cs[0], s = s_cs[1];
    IvmTileGeom G;
    ivm_tile_geom(P, px, pz, c, s, r0, r1, c0, c1, G);
    int row_lo = max(G.row_lo, e.origin_r), row_hi = min(G.row_hi, e.origin_r + P.SR - 1);
    if (e.count <= 0) row_hi = row_lo - 1;
    if (row_hi - row_lo + 1 > max_rows) row_hi = row_lo + max_rows - 1;  // cannot happen: max_rows bounds the tile diagonal
    const int nrows = max(row_hi - row_lo + 1, 0);
    const uint32_t *off = P.koff + (size_t)b * ((size_t)P.SR * P.SC + 1);
    for (int i = tid; i < nrows; i += nthr) {
        int clo, chi;
        ivm_row_span(G, row_lo + i, clo, chi);
        clo = max(clo, e.origin_c); chi = min(chi, e.origin_c + P.SC - 1);
        uint32_t p0 = 0u, p1 = 0u;
        if (chi >= clo) {
            const uint32_t *o = off + (size_t)(row_lo + i - e.origin_r) * P.SC;
            p0 = __ldg(o + (clo - e.origin_c)); p1 = __ldg(o + (chi - e.origin_c + 1));
        }
        s_p0[i] = p0; s_off[i + 1] = (int32_t)(p1 - p0);
    }
    if (tid == 0) s_off[0] = 0;
    __syncthreads();
    if (tid < 32) {  // inclusive scan of the row counts
        int carry = 0;
        for (int i0 = 0; i0 < nrows; i0 += 32) {
            const int i = i0 + tid;
            int x = i < nrows ? s_off[i + 1] : 0;
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (tid >= o) x += y; }
            if (i < nrows) s_off[i + 1] = carry + x;
            carry += __shfl_sync(0xffffffffu, x, 31);
        }
    }
    __syncthreads();
    const int total = nrows > 0 ? s_off[nrows] : 0;
    const IvmRecord *pts = P.kpts + (size_t)b * P.kcap;
    unsigned n_in = 0;
    const float ylo = ivm_sub(h, 1.25f), yhi = ivm_add(h, 0.75f);
    int row = 0;  // the row of this thread's current flat index: indices only grow, so the row is walked forward
    for (int base = 0; base < total; base += nthr * IVM_KNOWN_MLP) {
        uint4 raw[IVM_KNOWN_MLP];
        bool have[IVM_KNOWN_MLP];
#pragma unroll
        for (int u = 0; u < IVM_KNOWN_MLP; ++u) {
            const int f = base + u * nthr + tid;
            have[u] = f < total;
            raw[u] = make_uint4(0, 0, 0, 0);
            if (have[u]) {
                while (s_off[row + 1] <= f) ++row;            // last row with s_off[row] <= f
                raw[u] = __ldg(reinterpret_cast<const uint4 *>(pts + (s_p0[row] + (uint32_t)(f - s_off[row]))));
            }
        }
        // raw.w = (npz index << 8) | label: later points overwrite earlier ones (mapper.py:569-571), label 0 excluded
        raster_records<IVM_KNOWN_MLP, true>(P, raw, have, 0u, px, pz, c, s, ylo, yhi, (float)r0, (float)r1, (float)c0, (float)c1, r0, c0,
                                            tc, nullptr, skey, socc, n_in);
    }
    __syncthreads();
    const int wr = r1 - r0, wc = c1 - c0;
    for (int i = tid; i < wr * wc; i += nthr) {
        const int rr = i / wc, cc = i - rr * wc;
        const size_t o = ((size_t)b * P.R + (size_t)(r0 + rr)) * P.C + (size_t)(c0 + cc);
        P.occ[o] = socc[rr * tc + cc];
        P.sem[o] = (uint8_t)(skey[rr * tc + cc] & 0xFFu);
    }
    const unsigned wn = warp_sum(n_in);
    if (wn && (tid & 31) == 0) atomicAdd(&P.g->known_in[parity], (unsigned long long)wn);
}

__global__ void k_pose(IvmParams P) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b == 0) P.g->stats[IVM_STAT_IN] = 0ull;  // known-map mode: this step's rastered-record count
    if (b < P.B && P.orient != nullptr) ivm_pose_matrices(P, b, P.T12_buf + 12 * b, P.cs_buf + 2 * b);
}

// ------------------------------------------------------------------ persistent step kernel: helpers
#define IVM_F_TILE 512         // pixels per tile of the persistent step kernel
#define IVM_F_GROUP 128        // threads of one raster group (two groups per CTA)

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Grid barrier over the co-resident CTAs of one launch (the host sizes the grid by occupancy), split into ARRIVE and WAIT so
// that independent work can sit between the two.  `target` = arrivals expected on the monotone
// counter.  grid_wait returns false on time-out (never observed; guards against a hang).
// grid_arrive: called by ONE thread after a CTA-level barrier that covers the writes to publish.
__device__ __forceinline__ void grid_arrive(uint32_t *bar) {
    __threadfence();  // release: this CTA's writes (made visible to this thread by the bar.sync) before the arrival
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(bar), "r"(1u) : "memory");
}
__device__ __forceinline__ bool grid_wait(uint32_t *bar, uint32_t target, int *s_flag) {
    if (threadIdx.x == 0) {
        int ok = 1;
        uint32_t spins = 0;
        for (;;) {
            // RELAXED polling: an acquire load would invalidate this SM's L1 (CCTL.IVALL) on every
            // iteration and stall the memory pipeline of the CTA that shares the SM and is still working
            uint32_t v;
            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
            if ((int32_t)(v - target) >= 0) break;
            if (++spins > (1u << 22)) { ok = 0; break; }
            __nanosleep(spins < 8 ? 64 : 256);
        }
        __threadfence();  // acquire, once: later loads of this CTA see the other CTAs' writes
        *s_flag = ok;
    }
    __syncthreads();
    return *s_flag != 0;
}
__device__ __forceinline__ bool grid_barrier(uint32_t *bar, uint32_t target, int *s_flag) {
    __syncthreads();
    if (threadIdx.x == 0) grid_arrive(bar);
    return grid_wait(bar, target, s_flag);
}

// ------------------------------------------------------------------ persistent step kernel
// The whole map update as ONE launch of co-resident CTAs.  With a class-score stream (PRED): 2 CTAs per SM of 13
// warps (8 geometry + 4 argmax + 1 producer); with GT labels: 3 CTAs per SM of 9 warps (8 geometry + 1 that joins the
// resolve).  CTA t owns the 512-pixel tiles t, t + grid, ... (valid pixels cluster in a few image rows, so tiles are
// dealt round-robin) and holds the queues of up to NSLOT of them (a chunk) in shared memory.  The score stream runs
// beside the geometry phases on its own warps from the first cycle.
//   A1 depth filter (warps 0..7): every pixel of the chunk -- depth -> height of the world point -> strict depth /
//        height filters; survivors compacted into the tile's queue as (pixel, depth).  Reads INPUTS only, so with
//        pipelined stepping it runs before the dependency wait on the previous step (a counter of finished CTAs).
//   G1 scatter (warps 0..7), one queue entry per thread across all tiles of the chunk: A2 = world x / z, half-cell,
//        store cell, L2 prefetch of the candidate word and the world record; B = ONE 64-bit RED.MAX per entry into
//        the candidate plane.  -> grid barrier 1
//   G3 resolve (warps 0..7; 0..8 with GT labels), 4 queue entries in flight per thread: candidate word + world record
//        (+ GT label) loads first, THEN the tile's labels are awaited (mbarrier); winners merge into the world store,
//        frame-edge winners decide against their partner cells (ivm_core.h).  -> grid barrier 2, launch_dependents
//   argmax (warps 8..11, PRED): running argmax over the staged planes (4 pixels per thread, LDS.128), labels to
//        shared memory (for G3) and to labels_out
//   producer (warp 12, one lane, PRED): keeps a 3 x 16 KB ring full with cp.async.bulk (TMA 1-D copies, SASS
//        UBLKCP), full/empty mbarriers; no block-wide barrier inside the stream
//   edge fix-up beside the raster: a small team of CTAs resolves the collisions on the edge lines of the world
//        bounding box (direct flow: scan only when a collision can have appeared since the last scan; generic flow
//        for degenerate boxes) while every other CTA already rasters the ego tiles the fix-up cannot touch;
//        128-thread groups (3 per CTA with PRED, 2 with GT), one ego tile each at a time, handed out dynamically.
#define IVM_O_THREADS_GT 288    // GT labels: 8 geometry warps + a ninth that joins the resolve
#define IVM_O_THREADS_PRED 416  // score stream: 8 geometry warps + 4 argmax warps + 1 producer warp
#define IVM_O_RGROUPS_PRED 3    // raster groups of 128 threads per CTA
#define IVM_O_RGROUPS_GT 2
#define IVM_O_TILE 512         // pixels per tile
#define IVM_O_SP 8             // planes per ring stage (16 KB)
#define IVM_O_NSTAGE 3         // ring depth: 48 KB per CTA; <= 64 KB per SM in flight keeps HBM saturated (measured) without
                               // building multi-microsecond queues in front of the geometry warps' loads and atomics
#define IVM_O_SLOTS_PRED 8     // tiles whose queue / labels a CTA holds at a time
#define IVM_O_SLOTS_GT 12     // 12 x 512 x 10 B of queues = 60 KB per CTA, three CTAs per SM
#define IVM_O_CTAS_PER_SM_PRED 2  // with the score stream: two CTAs per SM (ring + queues = 76 KB each)
#define IVM_O_CTAS_PER_SM_GT 3    // GT labels: every phase is latency-bound, a third CTA per SM (72 registers) pays: measured 85 -> 79 us at 32 envs
#define IVM_O_TILE_CTR 40      // word of the barrier block (its second 128-byte line) that hands out raster tiles
#define IVM_O_FIX_FLAG 48      // ... = step once CTA 0 has finished stage 1 of the edge fix-up
#define IVM_O_FIX_ARRIVE 56    // ... arrivals of the fix-up team after the edge-line scan (monotone)
#define IVM_O_DONE 64          // word of the barrier block (its third 128-byte line): CTAs that are done with the map state (monotone)

struct OvlSlot {               // one tile of a chunk
    float T[12];
    float cs[2];
    int32_t b, tp0, origin_r, origin_c, reset;
    uint32_t reset_stamp;
    float h, hlo, hhi;         // camera height and the strict height band (mapper.py:416-424)
    uint32_t qn;               // queued (valid) pixels of the tile
    int32_t box[5];            // bbox + count of the cells this CTA newly occupied in the slot's env
};
template <int NSLOT>
struct OvlShared {
    OvlSlot slot[NSLOT];
    int32_t bb[4];
    unsigned valid;
    int flag;
    uint64_t full[IVM_O_NSTAGE], empty[IVM_O_NSTAGE];
    uint64_t lab_full[NSLOT], lab_empty[NSLOT];
    int32_t next[4];           // raster: the group's next ego tile (dynamic distribution)
    int32_t gflag[4];
    int32_t pend[4][32];       // raster: tiles of the group that wait for the edge fix-up
    unsigned long long t_start; // %globaltimer when the CTA became resident
    int32_t qoff[NSLOT + 1];   // prefix of the chunk's queue lengths
    uint8_t qblk[NSLOT * IVM_O_TILE / 64];  // slot that holds queue entry 64 * j (the flat loops start their slot search there)
    int32_t glob[4], loc4[4];  // world box over the env boxes at grid barrier 2 / frame box of the step
    int32_t direct;            // both boxes allow the direct resolution of the edge collisions
    int32_t skipscan;          // ... and the edge lines are as the last scan left them (IvmGlobal::scan_glob)
    int32_t segcnt[2];         // direct edge-line scan: segments, longest segment
    int32_t gband[8];          // raster: bands of half-rows / half-cols that hold the global bbox edge lines (ovl_global_bands)
};

static inline size_t ovl_stream_bytes(bool pred) {
    const size_t nslot = pred ? IVM_O_SLOTS_PRED : IVM_O_SLOTS_GT;
    return (pred ? (size_t)IVM_O_NSTAGE * IVM_O_SP * IVM_O_TILE * sizeof(float) : 0) + nslot * IVM_O_TILE * (4 + 4 + 2 + (pred ? 1 : 0));
}

// A frame winner on the frame bbox edge: the fix-up decides its cell later.  The ego tiles of the pending point
// and of the record it may replace are rastered only after the fix-up.  Rare, so kept out of line.
__device__ __noinline__ void ovl_defer_edge(const IvmParams &P, int b, uint32_t pix, float x, float y, float z, int32_t r, int32_t c,
                                            uint32_t label, size_t idx, float ox, float oy, float oz, uint32_t ometa,
                                            uint32_t reset_stamp, float h) {
    IvmPoint pt;
    pt.x = x; pt.y = y; pt.z = z; pt.r = r; pt.c = c;
    IvmRecord old;
    old.x = ox; old.y = oy; old.z = oz; old.meta = ometa;
    ivm_push_edge1<IvmAtomics>(P, b, pix, pt, label, idx);
    const float epx = P.pose[3 * b + 0], epz = P.pose[3 * b + 2];
    const float ec = __ldcg(&P.cs[2 * b + 0]), es = __ldcg(&P.cs[2 * b + 1]);
    ivm_mark_tile(P, b, pt.x, pt.y, pt.z, epx, h, epz, ec, es);
    if (ivm_live(old.meta, reset_stamp)) ivm_mark_tile(P, b, old.x, old.y, old.z, epx, h, epz, ec, es);
}

__device__ __noinline__ bool ovl_frame_edge_loses(const IvmParams &P, int b, int32_t r, int32_t c, uint32_t ord,
                                                  unsigned long long xorder, const int32_t *loc) {
    return ivm_frame_edge_loses(P, b, r, c, ord, xorder, loc);
}

// wait of a thread GROUP (named barrier 1, nthr threads, leader = thread 0) on the grid barrier
__device__ __forceinline__ bool grid_wait_group(uint32_t *bar, uint32_t target, int *s_flag, int nthr, int bar_id) {
    if (threadIdx.x == 0) {
        int ok = 1;
        uint32_t spins = 0;
        for (;;) {
            uint32_t v;
            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
            if ((int32_t)(v - target) >= 0) break;
            if (++spins > (1u << 22)) { ok = 0; break; }
            __nanosleep(spins < 8 ? 32 : 128);
        }
        __threadfence();
        *s_flag = ok;
    }
    group_bar(bar_id, nthr);
    return *s_flag != 0;
}

// Where can the edge lines of the stage-2 (world) bounding box lie?  The box that the fix-up will compute
// (ivm_fixup_stage1) is the union of the env boxes as they stand at grid barrier 2 and of the frame-edge winners
// that stage 1 still merges; those lie inside the frame box `loc`.  So its first row is in
// [min(Rb2, loc.rmin), Rb2] with Rb2 = first row over the env boxes at barrier 2 (exactly Rb2 when the frame did
// not reach beyond the known world: the usual case), and likewise for the other three sides.  Reads that race
// with stage 1's merges only move Rb2 inside that band.  Called by one warp right after grid barrier 2;
// band[0..1] first-row band, [2..3] last-row band, [4..5] first-col band, [6..7] last-col band.
__device__ __forceinline__ void ovl_global_bands(const IvmParams &P, int lane, int32_t *band, int32_t *glob, int32_t *loc4, int32_t *direct,
                                                 int32_t *skipscan) {
    int rmin = INT32_MAX, rmax = INT32_MIN, cmin = INT32_MAX, cmax = INT32_MIN;
    for (int b = lane; b < P.B; b += 32) {
        const IvmEnv *e = &P.env[b];
        const int4 bx = __ldcg(reinterpret_cast<const int4 *>(&e->rmin));
        if (__ldcg(&e->count) > 0) { rmin = min(rmin, bx.x); rmax = max(rmax, bx.y); cmin = min(cmin, bx.z); cmax = max(cmax, bx.w); }
    }
    rmin = warp_min(rmin); rmax = warp_max(rmax); cmin = warp_min(cmin); cmax = warp_max(cmax);
    if (lane == 0) {
        const int32_t l0 = __ldcg(&P.g->loc[0]), l1 = __ldcg(&P.g->loc[1]), l2 = __ldcg(&P.g->loc[2]), l3 = __ldcg(&P.g->loc[3]);
        band[0] = min(rmin, l0); band[1] = rmin; band[2] = rmax; band[3] = max(rmax, l1);
        band[4] = min(cmin, l2); band[5] = cmin; band[6] = cmax; band[7] = max(cmax, l3);
        // direct resolution of the edge collisions (ivm_core.h): the frame box allowed the resolve to merge every frame
        // winner at once (no stage-1 list is pending), so the env boxes are final and the world box is their union
        glob[0] = rmin; glob[1] = rmax; glob[2] = cmin; glob[3] = cmax;
        loc4[0] = l0; loc4[1] = l1; loc4[2] = l2; loc4[3] = l3;
        const bool frame_direct = l0 > l1 || ivm_box_direct(loc4);
        *direct = (frame_direct && rmin <= rmax && ivm_box_direct(glob)) ? 1 : 0;
        // nothing has touched the edge lines of the same world box since the last direct scan: no class can have gained a member
        const int4 sg = __ldcg(reinterpret_cast<const int4 *>(P.g->scan_glob));
        *skipscan = (*direct && __ldcg(&P.g->scan_valid) != 0u && __ldcg(&P.g->scan_B) == P.B && __ldcg(&P.g->edge_touched) == 0u &&
                     sg.x == rmin && sg.y == rmax && sg.z == cmin && sg.w == cmax) ? 1 : 0;
    }
}

// image row / column of a pixel index (a shift and a mask for the usual power-of-two widths, P.w_shift >= 0)
__device__ __forceinline__ void ovl_row_col(const IvmParams &P, int pix, int &v, int &u) {
    if (P.w_shift >= 0) { v = pix >> P.w_shift; u = pix & (P.W - 1); }
    else { v = pix / P.W; u = pix - v * P.W; }
}

#define OVL_STEPLOG(k) P.cta_trace[(size_t)(512 + (P.step & 255u)) * IVM_TRACE_SLOTS + (k)]  // per-step log of CTA 0 (rows 512..767)
#define OVL_STAMP(k, who)                                                                              \
    do {                                                                                               \
        if (tid == (who) && blockIdx.x < IVM_TRACE_CTAS && trace_on)                                   \
            P.cta_trace[(size_t)blockIdx.x * IVM_TRACE_SLOTS + (k)] = global_timer();                  \
    } while (0)

template <bool PRED>
__global__ void __launch_bounds__(PRED ? IVM_O_THREADS_PRED : IVM_O_THREADS_GT, PRED ? IVM_O_CTAS_PER_SM_PRED : IVM_O_CTAS_PER_SM_GT)
k_step_overlap(const __grid_constant__ IvmParams P, const float *__restrict__ logits, int ncls, uint8_t *__restrict__ labels_out,
               int nenv_total, uint32_t bar_base, int max_rows, int raster_group_bytes, int stage_cap, int team, uint32_t team_base,
               int pipelined, uint32_t done_target) {
    constexpr int NG1 = 256;                               // G1 (depth scatter): warps 0..7 in both modes
    constexpr int NG = PRED ? 256 : IVM_O_THREADS_GT;      // G3 (resolve) threads: warps 0..7, and the ninth warp of the GT
                                                           // layout (idle in G1: one more warp makes most CTAs' queues ONE round)
    constexpr int GB3 = PRED ? 1 : 6;                      // named barrier of the G3 group (the GT group is not the G1 group)
    constexpr int RG = PRED ? IVM_O_RGROUPS_PRED : IVM_O_RGROUPS_GT;  // raster groups per CTA
    constexpr int NGW = NG / 32;
    constexpr int NSLOT = PRED ? IVM_O_SLOTS_PRED : IVM_O_SLOTS_GT;
    constexpr int NCW = 4;                                 // argmax warps (PRED)
    constexpr size_t RING_BYTES = PRED ? (size_t)IVM_O_NSTAGE * IVM_O_SP * IVM_O_TILE * sizeof(float) : 0;
    extern __shared__ __align__(128) unsigned char dyn[];
    __shared__ __align__(8) OvlShared<NSLOT> sh;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tpe = P.HW / IVM_O_TILE;                     // tiles per env
    const int total = P.B * tpe;                           // < 2^31 (checked by the host)
    // tiles are dealt round-robin: tile = CTA + j * grid (valid pixels cluster in a few image rows)
    const int grid_n = (int)gridDim.x, cta = (int)blockIdx.x;
    const int my_tiles = (total - cta + grid_n - 1) / grid_n;
    const bool trace_on = !(P.debug & 4096) || (P.step & 31u) == 20u;  // debug bit 4096: trace one step in 32 (a mid-burst step)
    const int nslot = (P.debug & 128) ? 2 : NSLOT;         // tiles per chunk (debug bit 128: tiny chunks, to test the multi-chunk path)
    const bool keep_queue = my_tiles <= nslot;             // the queue built by G1 is still there in G3
    uint32_t *qcell = reinterpret_cast<uint32_t *>(dyn + RING_BYTES);
    float *qd = reinterpret_cast<float *>(dyn + RING_BYTES + (size_t)NSLOT * IVM_O_TILE * 4);
    uint16_t *qpix = reinterpret_cast<uint16_t *>(dyn + RING_BYTES + (size_t)NSLOT * IVM_O_TILE * 8);
    uint8_t *slab = dyn + RING_BYTES + (size_t)NSLOT * IVM_O_TILE * 10;
    IvmGlobal *g = P.g;
    const bool geo = warp < NGW;
    const int nstage = ((P.debug >> 8) & 15) ? min((P.debug >> 8) & 15, IVM_O_NSTAGE) : IVM_O_NSTAGE;  // experiment: shallower ring

    // ---- geometry warps (0..7): the depth scatter (G1) and the resolve (G3) work on a CHUNK of tiles (<= NSLOT, whose
    // queues fit shared memory) in phases that span all tiles of the chunk, so that nothing is serialised per tile:
    //   A1  every pixel: depth -> height of the world point -> strict depth / height filters (mapper.py:416-424); the
    //       survivors (about one pixel in four) are compacted into the tile's queue as (pixel, depth).  Inputs only.
    //   A2  one queue entry per thread: world x / z, half-cell (two IEEE divisions), store cell; the candidate word and
    //       the world record of the cell are PREFETCHED into L2; frame bbox
    //   B   ONE 64-bit RED.MAX per queue entry into the candidate plane                          -> grid barrier 1
    //   G3  one queue entry per thread: candidate word + world record (+ GT label) loads in flight together, then the
    //       tile's labels are awaited and the winners merge into the world store                 -> grid barrier 2
    constexpr int F4T = IVM_O_TILE / 4;                    // 128-bit depth loads per tile
    constexpr int A1B = 4;                                 // depth loads in flight per thread
    constexpr int E3 = PRED ? 2 : 1;                       // queue entries in flight per thread in G3 (measured: 4 in flight spill in
                                                           // the merge and lose 2 us (PRED) / 7 us (GT, 32 envs) per step)
    const int cn0 = min(nslot, my_tiles);                  // tiles of the first chunk
    // input half of the slot prep (pose, camera height, pose matrices): 4 slots per warp at a time, 8 lanes each
    auto prep_inputs = [&](int c0, int cn) {
        const int role = lane & 7, k = warp + (lane >> 3) * (NG1 / 32);
        if (k < cn) {
            OvlSlot &sl = sh.slot[k];
            const int tile = cta + (c0 + k) * grid_n;
            const int b = tile / tpe;
            if (role == 0) {
                const float h = P.pose[3 * b + 1];
                sl.b = b; sl.tp0 = (tile - b * tpe) * IVM_O_TILE;
                sl.h = h; sl.hlo = ivm_sub(h, 1.0f); sl.hhi = ivm_add(h, 0.5f);
                sl.qn = 0u;
                sl.box[0] = INT32_MAX; sl.box[1] = INT32_MIN; sl.box[2] = INT32_MAX; sl.box[3] = INT32_MIN; sl.box[4] = 0;
            } else if (role == 1) {
                if (P.orient != nullptr) ivm_pose_matrices(P, b, sl.T, sl.cs);
                else
                    for (int i = 0; i < 12; ++i) sl.T[i] = P.T12[12 * b + i];
            }
        }
    };
    // state half: env decision (reset / store origin, mapper.py:310-326), or the state the env's first CTA published
    auto prep_state = [&](int c0, int cn, bool published) {
        const int role = lane & 7, k = warp + (lane >> 3) * (NG1 / 32);
        if (k < cn && role == 2) {
            OvlSlot &sl = sh.slot[k];
            const int b = (cta + (c0 + k) * grid_n) / tpe;
            const IvmEnv *e = &P.env[b];
            const uint32_t m = P.masks[b];
            const int32_t cnt = __ldcg(&e->count), eor = __ldcg(&e->origin_r), eoc = __ldcg(&e->origin_c);
            const uint32_t ers = __ldcg(&e->reset_stamp);
            if (published) {
                sl.origin_r = eor; sl.origin_c = eoc; sl.reset = 0; sl.reset_stamp = ers;
            } else {
                const IvmEnvPrep q = ivm_env_decide_vals(P, m, cnt, eor, eoc, P.pose[3 * b + 0], P.pose[3 * b + 2]);
                sl.origin_r = q.origin_r; sl.origin_c = q.origin_c; sl.reset = q.reset;
                sl.reset_stamp = q.reset ? P.step : ers;
            }
        }
    };
    auto a1_load = [&](int c0, int cn, int f0, float4 *dv) {
#pragma unroll
        for (int q = 0; q < A1B; ++q) {
            const int f = f0 + q * NG1 + tid;
            dv[q] = make_float4(2.0f, 2.0f, 2.0f, 2.0f);
            if (f < cn * F4T)
                dv[q] = __ldcg(reinterpret_cast<const float4 *>(P.depth + (size_t)(cta + (c0 + f / F4T) * grid_n) * IVM_O_TILE + (f % F4T) * 4));
        }
    };
    auto a1_filter = [&](int cn, int f0, const float4 *dv) {
#pragma unroll
        for (int q = 0; q < A1B; ++q) {
            if (f0 + q * NG1 + (tid & ~31) >= cn * F4T) continue;  // warp-uniform: a warp's 32 loads lie in one tile
            const int f = f0 + q * NG1 + tid, k = f / F4T, gt = f % F4T;
            OvlSlot &sl = sh.slot[k];
            const int pix0 = sl.tp0 + gt * 4;
            int v, u0;
            ovl_row_col(P, pix0, v, u0);
            const float ysv = P.ys[v];
            const float4 xq = __ldg(reinterpret_cast<const float4 *>(P.xs + u0));  // (u0 is a multiple of 4, the table 16-byte aligned)
            const float xsv[4] = {xq.x, xq.y, xq.z, xq.w};
            const float dd[4] = {dv[q].x, dv[q].y, dv[q].z, dv[q].w};
            bool ok[4];
            unsigned mk[4];
            int n = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float y = ivm_world_y(dd[j], xsv[j], ysv, sl.T);
                ok[j] = dd[j] > 0.01f && dd[j] < 0.99f && y > sl.hlo && y < sl.hhi;
                mk[j] = __ballot_sync(0xffffffffu, ok[j]);
                n += __popc(mk[j]);
            }
            if (n) {  // warp-uniform
                unsigned pos = 0;
                if (lane == 0) pos = atomicAdd(&sl.qn, (unsigned)n);
                pos = __shfl_sync(0xffffffffu, pos, 0);
                const unsigned below = (1u << lane) - 1u;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (ok[j]) {
                        const unsigned o = (unsigned)k * IVM_O_TILE + pos + __popc(mk[j] & below);
                        qpix[o] = (uint16_t)(gt * 4 + j); qd[o] = dd[j];
                    }
                    pos += __popc(mk[j]);
                }
            }
        }
    };
    auto a1_rest = [&](int c0, int cn, int f_first, float4 *dv) {  // the batches after the first (or all of them)
        for (int f0 = f_first; f0 < cn * F4T; f0 += A1B * NG1) {
            a1_load(c0, cn, f0, dv);
            a1_filter(cn, f0, dv);
        }
    };
    auto queue_prefix = [&](int cn) {  // warp 0, after a barrier that covers A1
        if (warp == 0) {
            unsigned x = lane < cn ? sh.slot[lane].qn : 0u;
            for (int o = 1; o < 32; o <<= 1) { const unsigned y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
            if (lane < cn) sh.qoff[lane + 1] = (int)x;
            if (lane == 0) sh.qoff[0] = 0;
            __syncwarp();
            const int totalq = sh.qoff[cn];
            for (int j = lane; j * 64 < totalq; j += 32) {
                int k = 0;
                while (j * 64 >= sh.qoff[k + 1]) ++k;
                sh.qblk[j] = (uint8_t)k;
            }
        }
    };
    auto slot_of = [&](int e) {  // the slot (tile of the chunk) that holds flat queue entry e
        int k = (int)sh.qblk[e >> 6];
        while (e >= sh.qoff[k + 1]) ++k;
        return k;
    };
    float4 dv[A1B];
    if (tid < NG1) {
        if (!pipelined) asm volatile("griddepcontrol.wait;" ::: "memory");  // nothing is read before the predecessor has completed
        // the first batch of depth loads goes out before anything else (ahead of the score stream's first 48 KB)
        a1_load(0, cn0, 0, dv);
        if ((lane & 7) == 0)  // the later batches are pulled towards L2: one prefetch per 128-byte line
            for (int f = A1B * NG1 + tid; f < cn0 * F4T; f += NG1)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(P.depth + (size_t)(cta + (f / F4T) * grid_n) * IVM_O_TILE + (f % F4T) * 4));
    } else if (!pipelined) {
        asm volatile("griddepcontrol.wait;" ::: "memory");
    }
    // In pipelined mode everything up to the dependency wait below touches only this step's INPUTS, shared memory and
    // labels_out: the launch is programmatically serialised behind the previous step's kernel (which releases its
    // dependents once it is past its grid barrier 2), so these CTAs become resident while the previous step still
    // rasters -- the score stream, the pose matrices and the depth filter of this step run beside that tail.
    if (tid == 0) {
        sh.t_start = global_timer();
        if (PRED) {
            for (int s = 0; s < IVM_O_NSTAGE; ++s) { mbar_init(&sh.full[s], 1); mbar_init(&sh.empty[s], NCW); }
            for (int s = 0; s < NSLOT; ++s) { mbar_init(&sh.lab_full[s], NCW); mbar_init(&sh.lab_empty[s], NGW); }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        sh.bb[0] = INT32_MAX; sh.bb[1] = INT32_MIN; sh.bb[2] = INT32_MAX; sh.bb[3] = INT32_MIN; sh.valid = 0;
    }
    __syncthreads();

    if (tid < NG1) {
        // ============================================================ G1: depth scatter (warps 0..7)
        int rmin = INT32_MAX, rmax = INT32_MIN, cmin = INT32_MAX, cmax = INT32_MIN;
        unsigned nvalid = 0;
        prep_inputs(0, cn0);
        group_bar(1, NG1);
        a1_filter(cn0, 0, dv);
        a1_rest(0, cn0, A1B * NG1, dv);
        OVL_STAMP(12, 0);
        // ---- the previous step: its kernel has completed (griddepcontrol.wait above), or -- pipelined -- every CTA of
        //      it has signalled that it is done with the map state
        if (tid == 0) {
            if (pipelined) {
                uint32_t spins = 0;
                for (;;) {
                    uint32_t v;
                    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(&P.bar[IVM_O_DONE]) : "memory");
                    if ((int32_t)(v - done_target) >= 0) break;
                    if (++spins > (1u << 22)) { atomicOr(&g->err, IVM_ERR_GRID_BARRIER); break; }
                    __nanosleep(64);
                }
                __threadfence();  // acquire; also drops L1 lines this SM may have cached while the previous kernel was still writing
            }
            if (blockIdx.x == 0) {
                g->tstamp[0] = sh.t_start; g->tstamp[5] = 0ull; g->tstamp[6] = global_timer();
                OVL_STEPLOG(0) = g->tstamp[6]; OVL_STEPLOG(5) = sh.t_start; OVL_STEPLOG(4) = 0ull; OVL_STEPLOG(3) = 0ull;
                g->stat_in[(P.step + 1u) & 3u] = 0ull;  // the NEXT step's rastered-record count (this step's was zeroed a step ago)
                P.bar[IVM_O_TILE_CTR] = (unsigned)RG * (gridDim.x - (gridDim.x > 1 ? (unsigned)team : 0u));  // raster tiles handed out statically
            }
        }
        group_bar(1, NG1);  // covers A1's queues as well
        if (tid == 64 && blockIdx.x < IVM_TRACE_CTAS && trace_on) {
            P.cta_trace[(size_t)blockIdx.x * IVM_TRACE_SLOTS + 7] = sh.t_start;
            P.cta_trace[(size_t)blockIdx.x * IVM_TRACE_SLOTS + 4] = global_timer();
            for (int k = 13; k < 16; ++k) P.cta_trace[(size_t)blockIdx.x * IVM_TRACE_SLOTS + k] = 0ull;
        }
        // paused envs (mapper.py:315-318) are wiped by the last CTA
        if (blockIdx.x == gridDim.x - 1)
            for (int b = P.B; b < nenv_total; ++b) {
                IvmEnvPrep q; q.reset = 1; q.origin_r = 0; q.origin_c = 0;
                ivm_env_publish<IvmAtomics>(P, b, q, tid, NG1);
            }
        for (int c0 = 0; c0 < my_tiles; c0 += nslot) {
            const int cn = min(nslot, my_tiles - c0);
            if (c0) {
                group_bar(1, NG1);  // the previous chunk's slots and queues are no longer read
                prep_inputs(c0, cn);
                group_bar(1, NG1);
                a1_rest(c0, cn, 0, dv);
                group_bar(1, NG1);
            }
            prep_state(c0, cn, false);
            queue_prefix(cn);
            group_bar(1, NG1);
            if (c0 == 0) OVL_STAMP(1, 0);
            // the CTA that owns an env's first tile publishes the env's new state (read after barrier 1 only)
            for (int k = 0; k < cn; ++k) {
                const OvlSlot &sl = sh.slot[k];
                if (sl.tp0 != 0) continue;  // uniform
                IvmEnvPrep q;
                q.reset = sl.reset; q.origin_r = sl.origin_r; q.origin_c = sl.origin_c;
                ivm_env_publish<IvmAtomics>(P, sl.b, q, tid, NG1);
                if (P.orient != nullptr) {
                    if (tid < 12) P.T12_buf[12 * sl.b + tid] = sl.T[tid];
                    if (tid < 2) P.cs_buf[2 * sl.b + tid] = sl.cs[tid];
                }
            }
            const int totalq = sh.qoff[cn];
            // A2: one queue entry per thread
#pragma unroll 2
            for (int e = tid; e < totalq; e += NG1) {
                const int k = slot_of(e);
                const unsigned o = (unsigned)k * IVM_O_TILE + (unsigned)(e - sh.qoff[k]);
                const OvlSlot &sl = sh.slot[k];
                const int pix = sl.tp0 + (int)qpix[o];
                int v, u;
                ovl_row_col(P, pix, v, u);
                float x, y, z;
                ivm_world_xyz(qd[o], P.xs[u], P.ys[v], sl.T, x, y, z);
                const float rf = ivm_rint_div(z, P.half_res, P.inv_half_res);
                const float cf = ivm_rint_div(x, P.half_res, P.inv_half_res);
                const bool rep = fabsf(rf) < 1.0e9f && fabsf(cf) < 1.0e9f;
                const int32_t r = rep ? (int32_t)rf : 0, c = rep ? (int32_t)cf : 0;
                const int32_t rr = r - sl.origin_r, cc = c - sl.origin_c;
                uint32_t cellv = 0xFFFFFFFFu;
                if (rep && rr >= 0 && rr < P.SR && cc >= 0 && cc < P.SC) {
                    cellv = ((uint32_t)rr << 16) | (uint32_t)cc;
                    const size_t w = (size_t)sl.b * P.SR * P.SC + (size_t)rr * P.SC + (size_t)cc;
                    asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(&P.cplane[w]));
                    asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(&P.store[w]));
                    rmin = min(rmin, r); rmax = max(rmax, r); cmin = min(cmin, c); cmax = max(cmax, c);
                    ++nvalid;
                } else {
                    atomicOr(&g->err, IVM_ERR_STORE_OVERFLOW);
                }
                qcell[o] = cellv;
            }
            // B: ONE 64-bit RED.MAX per queue entry into the candidate plane (its line is in L2 or on its way)
#pragma unroll 2
            for (int e = tid; e < totalq; e += NG1) {
                const int k = slot_of(e);
                const unsigned o = (unsigned)k * IVM_O_TILE + (unsigned)(e - sh.qoff[k]);
                const uint32_t cellv = qcell[o];
                if (cellv == 0xFFFFFFFFu) continue;
                const OvlSlot &sl = sh.slot[k];
                const int pix = sl.tp0 + (int)qpix[o];
                int v, u;
                ovl_row_col(P, pix, v, u);
                const float y = ivm_world_y(qd[o], P.xs[u], P.ys[v], sl.T);
                unsigned long long *w = &P.cplane[(size_t)sl.b * P.SR * P.SC + (size_t)(cellv >> 16) * P.SC + (size_t)(cellv & 0xFFFFu)];
                const unsigned long long key = ivm_cand_key(P, y, (uint32_t)pix);
                asm volatile("red.relaxed.gpu.global.max.u64.L2::cache_hint [%0], %1, %2;" ::"l"(w), "l"(key), "l"(ivm_policy_keep()) : "memory");
            }
        }
        OVL_STAMP(2, 0);
        // frame bbox over ALL envs (mapper.py:465), one flush per CTA
        const unsigned wv = warp_sum(nvalid);
        if (wv) {
            rmin = warp_min(rmin); rmax = warp_max(rmax); cmin = warp_min(cmin); cmax = warp_max(cmax);
            if (lane == 0) {
                atomicMin(&sh.bb[0], rmin); atomicMax(&sh.bb[1], rmax); atomicMin(&sh.bb[2], cmin); atomicMax(&sh.bb[3], cmax);
                atomicAdd(&sh.valid, wv);
            }
        }
        group_bar(1, NG1);  // every G1 thread's REDs, queue entries and the env publication are issued
        if (tid == 0) {
            if (sh.valid) {
                atomicMin(&g->loc[0], sh.bb[0]); atomicMax(&g->loc[1], sh.bb[1]);
                atomicMin(&g->loc[2], sh.bb[2]); atomicMax(&g->loc[3], sh.bb[3]);
                atomicAdd(&g->acc_valid, (unsigned long long)sh.valid);
            }
            grid_arrive(P.bar);  // barrier 1
        }
    }
    if (geo) {
        OVL_STAMP(10, 0);
        if (!grid_wait_group(P.bar, bar_base + 1u * gridDim.x, &sh.flag, NG, GB3) && tid == 0) atomicOr(&g->err, IVM_ERR_GRID_BARRIER);
        if (blockIdx.x == 0 && tid == 0) { g->tstamp[1] = global_timer(); OVL_STEPLOG(1) = g->tstamp[1]; }
        OVL_STAMP(0, 0);

        // ============================================================ G3: resolve
        const int32_t loc[4] = {__ldcg(&g->loc[0]), __ldcg(&g->loc[1]), __ldcg(&g->loc[2]), __ldcg(&g->loc[3])};
        const bool direct1 = loc[0] <= loc[1] && ivm_box_direct(loc);  // frame-edge winners decide for themselves (ivm_core.h)
        unsigned nlocal = 0, nedge1 = 0;
        int chunk = 0;
        for (int c0 = 0; c0 < my_tiles; c0 += nslot, ++chunk) {
            const int cn = min(nslot, my_tiles - c0);
            if (!keep_queue) {
                // more tiles than slots: rebuild this chunk's slots (from the published env state) and queues
                group_bar(GB3, NG);
                if (tid < NG1) {  // (these are written for the G1 group)
                    prep_inputs(c0, cn);
                    prep_state(c0, cn, true);
                }
                group_bar(GB3, NG);
                if (tid < NG1) a1_rest(c0, cn, 0, dv);
                group_bar(GB3, NG);
                queue_prefix(cn);
                group_bar(GB3, NG);
                const int tq = sh.qoff[cn];
                for (int e = tid; e < tq; e += NG) {
                    const int k = slot_of(e);
                    const unsigned o = (unsigned)k * IVM_O_TILE + (unsigned)(e - sh.qoff[k]);
                    const OvlSlot &sl = sh.slot[k];
                    const int pix = sl.tp0 + (int)qpix[o];
                    int v, u;
                    ovl_row_col(P, pix, v, u);
                    float x, y, z;
                    ivm_world_xyz(qd[o], P.xs[u], P.ys[v], sl.T, x, y, z);
                    const float rf = ivm_rint_div(z, P.half_res, P.inv_half_res);
                    const float cf = ivm_rint_div(x, P.half_res, P.inv_half_res);
                    const bool rep = fabsf(rf) < 1.0e9f && fabsf(cf) < 1.0e9f;
                    const int32_t rr = (rep ? (int32_t)rf : 0) - sl.origin_r, cc = (rep ? (int32_t)cf : 0) - sl.origin_c;
                    qcell[o] = (rep && rr >= 0 && rr < P.SR && cc >= 0 && cc < P.SC) ? (((uint32_t)rr << 16) | (uint32_t)cc) : 0xFFFFFFFFu;
                }
                group_bar(GB3, NG);
            }
            const int totalq = sh.qoff[cn];
            for (int e0 = 0; e0 < totalq; e0 += E3 * NG) {
                bool act[E3];
                int kk[E3];
                uint32_t ce[E3], px[E3], lab[E3];
                float d[E3];
                unsigned long long cw[E3];
                IvmRecord old[E3];
                // all loads of the round first: candidate word, world record (speculative), GT label
#pragma unroll
                for (int u = 0; u < E3; ++u) {
                    const int e = e0 + u * NG + tid;
                    act[u] = e < totalq;
                    kk[u] = 0; ce[u] = 0u; px[u] = 0u; lab[u] = 0u; d[u] = 2.0f; cw[u] = 0ull;
                    old[u].x = old[u].y = old[u].z = 0.f; old[u].meta = 0u;
                    if (act[u]) {
                        const int k = slot_of(e);
                        const unsigned o = (unsigned)k * IVM_O_TILE + (unsigned)(e - sh.qoff[k]);
                        kk[u] = k; ce[u] = qcell[o]; px[u] = qpix[o]; d[u] = qd[o];
                        act[u] = ce[u] != 0xFFFFFFFFu;  // outside the store window: flagged by the scatter
                        if (act[u]) {
                            const OvlSlot &sl = sh.slot[k];
                            const size_t w = (size_t)sl.b * P.SR * P.SC + (size_t)(ce[u] >> 16) * P.SC + (size_t)(ce[u] & 0xFFFFu);
                            cw[u] = ivm_load_ull(&P.cplane[w]);
                            old[u] = ivm_load_record(&P.store[w]);
                            if (!PRED) lab[u] = __ldcg(P.labels + (size_t)sl.b * P.HW + sl.tp0 + px[u]);
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < E3; ++u) {
                    if (!act[u]) continue;
                    const int k = kk[u];
                    OvlSlot &sl = sh.slot[k];
                    const uint32_t pix = (uint32_t)sl.tp0 + px[u];
                    int v, uu;
                    ovl_row_col(P, (int)pix, v, uu);
                    IvmPoint pt;
                    ivm_world_xyz(d[u], P.xs[uu], P.ys[v], sl.T, pt.x, pt.y, pt.z);
                    if (cw[u] != ivm_cand_key(P, pt.y, pix)) continue;  // another pixel owns the cell
                    pt.r = sl.origin_r + (int32_t)(ce[u] >> 16); pt.c = sl.origin_c + (int32_t)(ce[u] & 0xFFFFu);
                    const size_t w = (size_t)sl.b * P.SR * P.SC + (size_t)(ce[u] >> 16) * P.SC + (size_t)(ce[u] & 0xFFFFu);
                    uint32_t label = lab[u];
                    if (PRED) {
                        mbar_wait(&sh.lab_full[k], (uint32_t)(chunk & 1));  // the tile's labels are in shared memory
                        label = (uint32_t)slab[k * IVM_O_TILE + px[u]];
                    }
                    if (ivm_on_frame_edge(pt, loc)) {
                        ++nedge1;
                        if (!direct1) {  // tiny frame box: classes of any size, resolved by CTA 0 after grid barrier 2
                            ovl_defer_edge(P, sl.b, pix, pt.x, pt.y, pt.z, pt.r, pt.c, label, w, old[u].x, old[u].y, old[u].z,
                                           old[u].meta, sl.reset_stamp, sl.h);
                            continue;
                        }
                        // stage 1 of the edge fix-up, resolved here: the <= 5 cells that share this cell's key are looked up
                        // in the candidate plane; the class winner merges like any other winner, the others are dropped
                        if (ovl_frame_edge_loses(P, sl.b, pt.r, pt.c, ivm_orderable(pt.y), (unsigned long long)sl.b * (unsigned long long)P.HW + pix, loc))
                            continue;
                    }
                    IvmBoxAcc acc;
                    acc.clear();
                    ivm_merge_record<IvmAtomics>(P, sl.b, w, pt.r, pt.c, pt.x, pt.y, pt.z, label, old[u], sl.reset_stamp,
                                                 sl.origin_r, sl.origin_c, acc);
                    ++nlocal;
                    if (acc.n) {  // a newly occupied cell: fold into the slot's box
                        atomicMin(&sl.box[0], acc.rmin); atomicMax(&sl.box[1], acc.rmax);
                        atomicMin(&sl.box[2], acc.cmin); atomicMax(&sl.box[3], acc.cmax);
                        atomicAdd(&sl.box[4], 1);
                        // on an edge line of the last scanned world box?  Then this step scans again (IvmGlobal::scan_glob)
                        const int4 sg = __ldcg(reinterpret_cast<const int4 *>(g->scan_glob));
                        if (pt.r == sg.x || pt.r == sg.y || pt.c == sg.z || pt.c == sg.w) g->edge_touched = 1u;
                    }
                }
            }
            if (PRED) {
                // every label of the chunk has been read by this warp (it waits for the whole chunk's labels first, so that
                // the argmax warps' protocol does not depend on which tiles had winners)
                for (int k = 0; k < cn; ++k) mbar_wait(&sh.lab_full[k], (uint32_t)(chunk & 1));
                __syncwarp();
                if (lane == 0)
                    for (int k = 0; k < cn; ++k) mbar_arrive(&sh.lab_empty[k]);
            }
            group_bar(GB3, NG);
            if (tid < cn && sh.slot[tid].box[4] > 0) {
                IvmBoxAcc t;
                const OvlSlot &sl = sh.slot[tid];
                t.rmin = sl.box[0]; t.rmax = sl.box[1]; t.cmin = sl.box[2]; t.cmax = sl.box[3]; t.n = sl.box[4];
                ivm_box_flush<IvmAtomics>(&P.env[sl.b], t);
            }
        }
        const unsigned wl = warp_sum(nlocal);
        if (wl && lane == 0) atomicAdd(&g->acc_local, (unsigned long long)wl);
        const unsigned we = warp_sum(nedge1);
        if (we && lane == 0) atomicAdd(&g->acc_e1, (unsigned long long)we);
        OVL_STAMP(3, 0);
        group_bar(GB3, NG);
        if (tid == 0) grid_arrive(P.bar);  // barrier 2
    } else if (PRED && warp < NGW + NCW) {
        // ============================================================ argmax warps: PredictSemantics tail
        // (mapper.py:795-798): running argmax over the planes, first max wins, NaN counts as maximal (torch.argmax)
        float(*ring)[IVM_O_SP][IVM_O_TILE] = reinterpret_cast<float(*)[IVM_O_SP][IVM_O_TILE]>(dyn);
        const int ct = tid - NG;  // 0..127, 4 pixels each
        int slot = 0;
        uint32_t round = 0;
        for (int j = 0; j < my_tiles; ++j) {
            const int tile = cta + j * grid_n;
            const int ls = j % nslot, chunk = j / nslot;
            float b0 = 0.f, b1 = 0.f, b2 = 0.f, b3 = 0.f;
            int a0 = 0, a1 = 0, a2 = 0, a3 = 0;
            for (int p0 = 0; p0 < ncls; p0 += IVM_O_SP) {
                mbar_wait(&sh.full[slot], round & 1u);
                const int np = min(IVM_O_SP, ncls - p0);
                if (np == IVM_O_SP) {
                    float4 v[IVM_O_SP];
#pragma unroll
                    for (int p = 0; p < IVM_O_SP; ++p) v[p] = *reinterpret_cast<const float4 *>(&ring[slot][p][ct * 4]);
#pragma unroll
                    for (int p = 0; p < IVM_O_SP; ++p) {
                        if (p0 + p == 0) { b0 = v[p].x; b1 = v[p].y; b2 = v[p].z; b3 = v[p].w; }
                        else {
                            IVM_ARGMAX_STEP(v[p].x, p0 + p, b0, a0); IVM_ARGMAX_STEP(v[p].y, p0 + p, b1, a1);
                            IVM_ARGMAX_STEP(v[p].z, p0 + p, b2, a2); IVM_ARGMAX_STEP(v[p].w, p0 + p, b3, a3);
                        }
                    }
                } else {
                    for (int p = 0; p < np; ++p) {
                        const float4 v = *reinterpret_cast<const float4 *>(&ring[slot][p][ct * 4]);
                        if (p0 + p == 0) { b0 = v.x; b1 = v.y; b2 = v.z; b3 = v.w; }
                        else {
                            IVM_ARGMAX_STEP(v.x, p0 + p, b0, a0); IVM_ARGMAX_STEP(v.y, p0 + p, b1, a1);
                            IVM_ARGMAX_STEP(v.z, p0 + p, b2, a2); IVM_ARGMAX_STEP(v.w, p0 + p, b3, a3);
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&sh.empty[slot]);  // this warp is done with the stage
                if (++slot == nstage) { slot = 0; ++round; }
            }
            if (chunk > 0) mbar_wait(&sh.lab_empty[ls], (uint32_t)((chunk - 1) & 1));  // G3 is done with the slot's previous labels
            uchar4 o;
            o.x = (uint8_t)a0; o.y = (uint8_t)a1; o.z = (uint8_t)a2; o.w = (uint8_t)a3;
            *reinterpret_cast<uchar4 *>(slab + ls * IVM_O_TILE + ct * 4) = o;
            *reinterpret_cast<uchar4 *>(labels_out + (size_t)tile * IVM_O_TILE + ct * 4) = o;
            __syncwarp();
            if (lane == 0) mbar_arrive(&sh.lab_full[ls]);
        }
        OVL_STAMP(6, NG);
    } else if (PRED && warp == NGW + NCW) {
        // ============================================================ producer warp: one lane keeps the ring full
        if (lane == 0) {
            float(*ring)[IVM_O_SP][IVM_O_TILE] = reinterpret_cast<float(*)[IVM_O_SP][IVM_O_TILE]>(dyn);
            uint64_t policy;
            asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
            int slot = 0;
            uint32_t round = 0;  // how many times the ring has wrapped
            for (int tile = cta; tile < total; tile += grid_n) {
                const int eb = tile / tpe, tp0 = (tile - eb * tpe) * IVM_O_TILE;
                const float *src = logits + (size_t)eb * ncls * P.HW + tp0;
                for (int p0 = 0; p0 < ncls; p0 += IVM_O_SP) {
                    const int np = min(IVM_O_SP, ncls - p0);
                    if (round > 0) mbar_wait(&sh.empty[slot], (round - 1) & 1u);
                    mbar_expect_tx(&sh.full[slot], (uint32_t)(np * IVM_O_TILE * sizeof(float)));
                    for (int p = 0; p < np; ++p)
                        bulk_g2s(&ring[slot][p][0], src + (size_t)(p0 + p) * P.HW, IVM_O_TILE * sizeof(float), &sh.full[slot],
                                 policy);
                    if (++slot == nstage) { slot = 0; ++round; }
                }
            }
        }
    }
    __syncthreads();  // ring and queues are drained: their memory becomes fix-up scratch / raster tiles
    if (!grid_wait(P.bar, bar_base + 2u * gridDim.x, &sh.flag)) {
        if (tid == 0) {
            atomicOr(&g->err, IVM_ERR_GRID_BARRIER);
            asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(&P.bar[IVM_O_DONE]), "r"(1u) : "memory");
        }
        return;
    }
    if (blockIdx.x == 0 && tid == 0) { g->tstamp[2] = global_timer(); OVL_STEPLOG(2) = g->tstamp[2]; }
    OVL_STAMP(5, 0);
    // every resolve of this step is done: the next step's kernel may become resident as CTAs of this one exit (it
    // touches nothing but its own inputs until this kernel has completed)
    if (tid == 0) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (warp == 0) ovl_global_bands(P, lane, sh.gband, sh.glob, sh.loc4, &sh.direct, &sh.skipscan);  // before anything of the fix-up can have moved an env box for good
    __syncthreads();

    // ================================================================ edge fix-up beside the raster
    // CTA 0 runs the whole fix-up program (stage-1 classes + merges, edge-line scan, stage-2 classes) on its
    // own -- no grid barrier inside -- while every other CTA already rasters the ego tiles the fix-up cannot
    // touch; then CTA 0 releases everybody with ONE add worth the three remaining arrivals, and the few
    // tiles that had to wait are rastered by the groups that hold them.  Tiles are handed out dynamically
    // (one static tile per group, then an atomic counter): their cost varies with the records under them.
    const int first = grid_n > 1 ? team : 0;              // CTAs [first, grid) start rastering at once
    const int tiles_x = (P.C + P.tile_c - 1) / P.tile_c, tiles_y = (P.R + P.tile_r - 1) / P.tile_r;
    const int per_env = tiles_x * tiles_y, units = P.B * per_env;
    const int group = tid / IVM_F_GROUP, gtid = tid - group * IVM_F_GROUP;
    uint32_t *gsm = reinterpret_cast<uint32_t *>(dyn + (size_t)group * raster_group_bytes);
    const uint32_t release_target = bar_base + 5u * gridDim.x;
    if (cta < team) {
        // The fix-up team: CTA 0 runs stages 1 and 2; the edge-line scan between them (thousands of scattered
        // 4-byte loads, more misses than one SM keeps in flight) is shared by the team's CTAs.
        IvmFixScratch S;
        S.key = reinterpret_cast<unsigned long long *>(dyn);
        S.xo = S.key + IVM_FIX_SMALL;
        S.ord = reinterpret_cast<uint32_t *>(S.xo + IVM_FIX_SMALL);
        S.cap = IVM_FIX_SMALL;
        S.ibuf = reinterpret_cast<int32_t *>(S.ord + IVM_FIX_SMALL);
        S.lbuf = reinterpret_cast<unsigned long long *>(S.ibuf + 8);
        S.release = P.bar; S.release_add = 3u * gridDim.x;  // stage 2 releases the deferred tiles itself
        auto spin_until = [&](uint32_t *word, uint32_t want, bool equal) {  // thread 0 only
            uint32_t spins = 0;
            for (;;) {
                uint32_t v;
                asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(word) : "memory");
                if (equal ? v == want : (int32_t)(v - want) >= 0) break;
                if (++spins > (1u << 24)) { atomicOr(&g->err, IVM_ERR_GRID_BARRIER); break; }
            }
            __threadfence();
        };
        bool direct = sh.direct != 0;
        constexpr int SEGCAP = 512;
        int32_t *seg_hdr = reinterpret_cast<int32_t *>(dyn + 12 * 1024);  // beyond the fix-up scratch (10.3 KB)
        if (direct) {
            // Direct flow: nothing of stage 1 is pending and the world box is known, so the team scans the edge lines at
            // once; every live record on them looks up the cells that share its key and lists itself if it loses.
            uint8_t *seg_tb = reinterpret_cast<uint8_t *>(dyn + 12 * 1024 + SEGCAP * IVM_SCAN_HDR * 4);  // P.B bytes (<= 4096 here)
            if (warp == 0) {
                if (sh.skipscan) { if (lane == 0) { sh.segcnt[0] = 0; sh.segcnt[1] = 0; } }
                else if (P.B <= 4096) ivm_direct_segments(P, sh.glob, seg_hdr, SEGCAP, sh.segcnt, lane, seg_tb);
                else if (lane == 0) sh.segcnt[0] = -1;
            }
            __syncthreads();
            direct = sh.segcnt[0] >= 0;  // (else: more segments than the table holds -- the generic flow below)
        }
        if (direct) {
            if (cta == 0) IVM_TRACE(g, 0, tid);
            const bool scan = sh.segcnt[0] > 0;  // (the same in every team CTA; usually nothing has to be scanned)
            if (scan) {
                const int nlive = ivm_direct_scan<IvmAtomics>(P, seg_hdr, sh.segcnt[0], sh.segcnt[1], sh.glob, sh.loc4, cta, team, tid, (int)blockDim.x);
                const unsigned wn2 = warp_sum((unsigned)nlive);
                if (wn2 && lane == 0) atomicAdd(&g->acc_e2, (unsigned long long)wn2);
                __syncthreads();
            }
            if (cta == 0) IVM_TRACE(g, 1, tid);
            // every decision of the team is made (round 1 of the team's arrival counter); then the listed losers are
            // deleted by the whole team (round 2) -- decisions must all read the store as it was.  The counter always
            // receives two arrivals per team CTA and step (the host advances its base by that much).
            if (team > 1) {
                if (tid == 0) {
                    __threadfence();
                    asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(&P.bar[IVM_O_FIX_ARRIVE]), "r"(1u) : "memory");
                    if (scan) spin_until(&P.bar[IVM_O_FIX_ARRIVE], team_base + (uint32_t)team, false);
                }
                __syncthreads();
            }
            const uint32_t nlose = scan ? min(__ldcg(&g->n_e2), P.ecap) : 0u;
            if (nlose) ivm_delete_losers<IvmAtomics>(P, nlose, (uint32_t)(cta * (int)blockDim.x + tid), (uint32_t)(team * (int)blockDim.x));
            __syncthreads();
            if (team > 1 && tid == 0) {
                __threadfence();
                asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(&P.bar[IVM_O_FIX_ARRIVE]), "r"(1u) : "memory");
            }
            if (cta == 0) {
                IVM_TRACE(g, 2, tid); IVM_TRACE(g, 3, tid);
                if (team > 1 && nlose) {
                    if (tid == 0) spin_until(&P.bar[IVM_O_FIX_ARRIVE], team_base + 2u * (uint32_t)team, false);
                    __syncthreads();
                }
                if (tid == 0) {
                    g->glob[0] = sh.glob[0]; g->glob[1] = sh.glob[1]; g->glob[2] = sh.glob[2]; g->glob[3] = sh.glob[3];
                    g->n_seg = (uint32_t)sh.segcnt[0];
                    g->scan_chunks = (uint32_t)((sh.segcnt[1] + IVM_SCAN_CHUNK - 1) / IVM_SCAN_CHUNK);
                }
                __syncthreads();
                ivm_fixup_stage2<IvmAtomics>(P, S, tid, blockDim.x, false, (uint32_t)__ldcg(&g->acc_e1), (uint32_t)__ldcg(&g->acc_e2), true);
                __syncthreads();
                if (tid == 0) {
                    // the edge lines of this world box are collision-free from here on (every class has one survivor)
                    g->scan_glob[0] = sh.glob[0]; g->scan_glob[1] = sh.glob[1]; g->scan_glob[2] = sh.glob[2]; g->scan_glob[3] = sh.glob[3];
                    g->scan_B = P.B; g->edge_touched = 0u; g->scan_valid = 1u;
                    g->tstamp[3] = global_timer(); g->tstamp[4] = g->tstamp[3]; OVL_STEPLOG(3) = g->tstamp[3];
                }
                __syncthreads();
            }
        } else {
        if (cta == 0) {
            ivm_fixup_stage1<IvmAtomics>(P, S, tid, blockDim.x);
            __syncthreads();
            if (team > 1 && tid == 0) {
                __threadfence();
                asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(&P.bar[IVM_O_FIX_FLAG]), "r"(P.step) : "memory");
            }
        } else {
            ivm_fixup_scan_prefetch(P, cta - 1, team - 1, tid, blockDim.x);  // while CTA 0 is in stage 1
            if (tid == 0) spin_until(&P.bar[IVM_O_FIX_FLAG], P.step, true);
            __syncthreads();
        }
        if (cta == 0) IVM_TRACE(g, 3, tid);
        ivm_fixup_scan_block<IvmAtomics>(P, S, cta, team, tid, blockDim.x);
        __syncthreads();
        if (team > 1 && tid == 0) {
            __threadfence();  // (two arrivals per team CTA and step, as in the direct flow: the host advances the base by 2 x team)
            asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(&P.bar[IVM_O_FIX_ARRIVE]), "r"(2u) : "memory");
        }
        if (cta == 0) {
            if (team > 1) {
                if (tid == 0) spin_until(&P.bar[IVM_O_FIX_ARRIVE], team_base + 2u * (uint32_t)team, false);
                __syncthreads();
            }
            ivm_fixup_stage2<IvmAtomics>(P, S, tid, blockDim.x);
            __syncthreads();
            if (tid == 0) { g->scan_valid = 0u; g->tstamp[3] = global_timer(); g->tstamp[4] = g->tstamp[3]; }
            __syncthreads();
        }
        }
    }
    if (warp < RG * IVM_F_GROUP / 32) {
        // group-level wait for the release (leader polls, named barrier of the group)
        auto wait_release = [&]() {
            if (gtid == 0) {
                int ok = 1;
                uint32_t spins = 0;
                for (;;) {
                    uint32_t v;
                    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(P.bar) : "memory");
                    if ((int32_t)(v - release_target) >= 0) break;
                    if (++spins > (1u << 22)) { ok = 0; break; }
                    __nanosleep(spins < 8 ? 32 : 128);
                }
                __threadfence();
                if (!ok) atomicOr(&g->err, IVM_ERR_GRID_BARRIER);
            }
            group_bar(2 + group, IVM_F_GROUP);
        };
        unsigned n_in = 0;
        bool released = cta == 0;                         // CTA 0 has just run the fix-up itself
        int npend = 0;
        int u = (cta - first) * RG + group;
        if (cta < first) {                                // the fix-up team joins late: no static tile
            if (gtid == 0) sh.next[group] = (int32_t)atomicAdd(&P.bar[IVM_O_TILE_CTR], 1u);
            group_bar(2 + group, IVM_F_GROUP);
            u = sh.next[group];
        }
        while (u < units) {
            int32_t unext = 0;
            if (gtid == 0) unext = (int32_t)atomicAdd(&P.bar[IVM_O_TILE_CTR], 1u);  // consumed after this tile
            const int b = u / per_env, w = u - b * per_env;
            const int ty = w / tiles_x, tx = w - ty * tiles_x;
            if (released) {
                raster_tile<false>(P, max_rows, b, ty * P.tile_r, tx * P.tile_c, gsm, gtid, IVM_F_GROUP, 2 + group, n_in, false, stage_cap);
            } else if (!raster_tile<false>(P, max_rows, b, ty * P.tile_r, tx * P.tile_c, gsm, gtid, IVM_F_GROUP, 2 + group, n_in, true, stage_cap, sh.gband)) {
                if (npend < 32) {
                    if (gtid == 0) sh.pend[group][npend] = u;
                    ++npend;
                } else {                                  // no room to remember it: wait here
                    wait_release();
                    released = true;
                    raster_tile<false>(P, max_rows, b, ty * P.tile_r, tx * P.tile_c, gsm, gtid, IVM_F_GROUP, 2 + group, n_in, false, stage_cap);
                }
            }
            if (gtid == 0) sh.next[group] = unext;
            group_bar(2 + group, IVM_F_GROUP);
            u = sh.next[group];
        }
        if (tid == 0) OVL_STAMP(11, 0);
        if (!released && npend > 0) wait_release();  // nothing pending: this group is done, whatever the fix-up still does
        if (tid == 0) OVL_STAMP(8, 0);
        for (int i = 0; i < npend; ++i) {
            const int up = sh.pend[group][i];
            const int b = up / per_env, w = up - b * per_env;
            const int ty = w / tiles_x, tx = w - ty * tiles_x;
            raster_tile<false>(P, max_rows, b, ty * P.tile_r, tx * P.tile_c, gsm, gtid, IVM_F_GROUP, 2 + group, n_in, false, stage_cap);
        }
        const unsigned wn = warp_sum(n_in);
        if (wn && lane == 0) atomicAdd(&g->stat_in[P.step & 3u], (unsigned long long)wn);
    }
    __syncthreads();
    OVL_STAMP(9, 0);
    if (tid == 0) {
        atomicMax(&g->tstamp[5], global_timer());
        atomicMax(&OVL_STEPLOG(4), global_timer());
        // this CTA is done with the map state: a pipelined successor waits for all of these instead of for the
        // completion of the whole kernel (which is signalled several microseconds after the last CTA has left).
        // Only the fix-up team has written anything the successor reads (store deletions, env boxes, the step's
        // published figures): the raster CTAs have only READ the map state -- their loads have returned -- and need no
        // fence in front of the signal (their maps and counts are not read by the successor).
        if (cta < team) __threadfence();
        asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(&P.bar[IVM_O_DONE]), "r"(1u) : "memory");
    }
}

// ------------------------------------------------------------------ known-map store build
__global__ void k_known_reset_env(IvmParams P, int b, long long n, int origin_r, int origin_c) {
    IvmEnv *e = &P.env[b];
    e->origin_r = origin_r; e->origin_c = origin_c;
    e->count = (int32_t)n; e->known_n = (int32_t)n;
    e->rmin = origin_r; e->rmax = origin_r + P.SR - 1; e->cmin = origin_c; e->cmax = origin_c + P.SC - 1;
    e->reset_stamp = 0; e->dirty = 0;
}

__device__ __forceinline__ bool known_cell(const IvmParams &P, const float *xyz, long long i, int origin_r, int origin_c,
                                           uint32_t &cell) {
    const float rf = ivm_rint_div(xyz[3 * i + 2], P.half_res, P.inv_half_res);
    const float cf = ivm_rint_div(xyz[3 * i + 0], P.half_res, P.inv_half_res);
    if (!(fabsf(rf) < 1.0e9f && fabsf(cf) < 1.0e9f)) return false;
    const int rr = (int)rf - origin_r, cc = (int)cf - origin_c;
    if (rr < 0 || rr >= P.SR || cc < 0 || cc >= P.SC) return false;
    cell = (uint32_t)rr * (uint32_t)P.SC + (uint32_t)cc;
    return true;
}

__global__ void k_known_hist(IvmParams P, int b, long long n, const float *xyz, int origin_r, int origin_c, uint32_t *hist) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        uint32_t cell;
        if (known_cell(P, xyz, i, origin_r, origin_c, cell)) atomicAdd(&hist[cell + 1], 1u);
        else atomicOr(&P.g->err, IVM_ERR_KNOWN_OVERFLOW);
    }
}

// inclusive scan of `n` u32 in chunks of 4096 per block (3 phases)
__global__ void __launch_bounds__(1024) k_scan_local(uint32_t *data, size_t n, uint32_t *totals) {
    __shared__ uint32_t sw[32];
    const size_t base = (size_t)blockIdx.x * 4096 + (size_t)threadIdx.x * 4;
    uint32_t v[4];
    for (int j = 0; j < 4; ++j) v[j] = (base + j < n) ? data[base + j] : 0u;
    v[1] += v[0]; v[2] += v[1]; v[3] += v[2];
    uint32_t x = v[3];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) sw[warp] = x;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = sw[lane];
        for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += y; }
        sw[lane] = w;
    }
    __syncthreads();
    const uint32_t prefix = (x - v[3]) + (warp ? sw[warp - 1] : 0u);
    for (int j = 0; j < 4; ++j) if (base + j < n) data[base + j] = v[j] + prefix;
    if (threadIdx.x == 1023) totals[blockIdx.x] = x + (warp ? sw[warp - 1] : 0u);
}
__global__ void __launch_bounds__(1024) k_scan_totals(uint32_t *totals, int nblocks) {
    __shared__ uint32_t sw[32];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int start = 0; start < nblocks; start += 1024) {
        const int i = start + threadIdx.x;
        uint32_t x = (i < nblocks) ? totals[i] : 0u;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (lane == 31) sw[warp] = x;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = sw[lane];
            for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += y; }
            sw[lane] = w;
        }
        __syncthreads();
        const uint32_t incl = x + (warp ? sw[warp - 1] : 0u) + carry;
        if (i < nblocks) totals[i] = incl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = incl;
        __syncthreads();
    }
}
__global__ void __launch_bounds__(1024) k_scan_add(uint32_t *data, size_t n, const uint32_t *totals) {
    if (blockIdx.x == 0) return;
    const uint32_t add = totals[blockIdx.x - 1];
    const size_t base = (size_t)blockIdx.x * 4096 + (size_t)threadIdx.x * 4;
    for (int j = 0; j < 4; ++j) if (base + j < n) data[base + j] += add;
}

__global__ void k_known_scatter(IvmParams P, int b, long long n, const float *xyz, const uint8_t *sem, int origin_r,
                                int origin_c, const uint32_t *off, uint32_t *fill) {
    IvmRecord *pts = P.kpts + (size_t)b * P.kcap;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        uint32_t cell;
        if (!known_cell(P, xyz, i, origin_r, origin_c, cell)) continue;
        const uint32_t pos = off[cell] + atomicAdd(&fill[cell], 1u);
        IvmRecord r;
        r.x = xyz[3 * i]; r.y = xyz[3 * i + 1]; r.z = xyz[3 * i + 2];
        r.meta = ((uint32_t)i << 8) | (uint32_t)sem[i];
        pts[pos] = r;
    }
}

// ------------------------------------------------------------------ export / maintenance
// one warp per (env, store row): count, then ordered write
__global__ void k_export_count(IvmParams P, uint32_t *rowlive) {
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (gw >= P.B * P.SR) return;
    const int b = gw / P.SR, rr = gw - b * P.SR;
    const IvmEnv e = P.env[b];
    unsigned n = 0;
    if (e.count > 0 && P.rowcount[(size_t)b * P.SR + rr] > 0) {
        const IvmRecord *row = P.store + ((size_t)b * P.SR + rr) * P.SC;
        for (int cc = lane; cc < P.SC; cc += 32) n += ivm_live(row[cc].meta, e.reset_stamp) ? 1u : 0u;
    }
    n = warp_sum(n);
    if (lane == 0) rowlive[gw] = n;
}
__global__ void k_export_write(IvmParams P, const uint32_t *rowoff_incl, long long cap, long long *env_out, float *xyz_out,
                               uint8_t *label_out, unsigned long long *key_out, unsigned long long *count_dev) {
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int total_rows = P.B * P.SR;
    if (gw == 0 && lane == 0) *count_dev = total_rows ? rowoff_incl[total_rows - 1] : 0ull;
    if (gw >= total_rows) return;
    const int b = gw / P.SR, rr = gw - b * P.SR;
    const IvmEnv e = P.env[b];
    if (e.count <= 0 || P.rowcount[(size_t)b * P.SR + rr] <= 0) return;
    long long pos = gw ? rowoff_incl[gw - 1] : 0;
    const IvmRecord *row = P.store + ((size_t)b * P.SR + rr) * P.SC;
    const IvmGlobal *g = P.g;
    for (int c0 = 0; c0 < P.SC; c0 += 32) {
        const int cc = c0 + lane;
        IvmRecord rec;
        rec.meta = 0;
        if (cc < P.SC) rec = row[cc];
        const bool live = cc < P.SC && ivm_live(rec.meta, e.reset_stamp);
        const unsigned m = __ballot_sync(0xffffffffu, live);
        if (live) {
            const long long o = pos + __popc(m & ((1u << lane) - 1u));
            if (o < cap) {
                env_out[o] = b;
                xyz_out[3 * o] = rec.x; xyz_out[3 * o + 1] = rec.y; xyz_out[3 * o + 2] = rec.z;
                label_out[o] = (uint8_t)(rec.meta & 0xFFu);
                key_out[o] = ivm_list_key(b, e.origin_r + rr, e.origin_c + cc, g->prev_rmin, g->prev_cmin, g->prev_R, g->prev_C);
            }
        }
        pos += __popc(m);
    }
}

__global__ void k_rebase(IvmParams P) {
    const size_t n = (size_t)P.maxB * P.SR * P.SC;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int b = (int)(i / ((size_t)P.SR * P.SC));
        const uint32_t meta = P.store[i].meta;
        if (meta == 0u) continue;
        P.store[i].meta = ivm_live(meta, P.env[b].reset_stamp) ? ((1u << 8) | (meta & 0xFFu)) : 0u;
    }
}
__global__ void k_rebase_env(IvmParams P) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < P.maxB) P.env[b].reset_stamp = 1u;
}

// ------------------------------------------------------------------ map features
// SemanticMapEncoder.generate_map_features (models/encoders/map_encoder.py:85-90): occupancy plane + one-hot of the
// semantic map, float32 [B, 1 + K, R, C].  Pure write stream (4 (1 + K) bytes out per 2 bytes in): one thread per
// 4 consecutive cells, one 128-bit store per thread and channel plane (coalesced), evict-first.
__global__ void __launch_bounds__(256) k_map_features(const uint8_t *__restrict__ occ, const uint8_t *__restrict__ sem, long long ncell,
                                                      int planes_cells, int num_classes, float *__restrict__ out,
                                                      uint32_t *__restrict__ err) {
    // ncell = B * R * C (multiple of 4 here), planes_cells = R * C (multiple of 4); blockIdx.y = output channel
    // (0 = occupancy, 1 + k = class k): the 2 input bytes per cell are re-read per channel from L1 / L2, the
    // 4 output bytes per cell and channel are what the kernel is bound by
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long i0 = q * 4;
    if (i0 >= ncell) return;
    const int ch = blockIdx.y;
    const long long b = i0 / planes_cells, within = i0 - b * planes_cells;
    float *dst = out + (b * (1 + num_classes) + ch) * (long long)planes_cells + within;
    if (ch == 0) {
        const uchar4 o = *reinterpret_cast<const uchar4 *>(occ + i0);
        __stcs(reinterpret_cast<float4 *>(dst), make_float4((float)o.x, (float)o.y, (float)o.z, (float)o.w));
        return;
    }
    const uchar4 l = *reinterpret_cast<const uchar4 *>(sem + i0);
    const int k = ch - 1;
    __stcs(reinterpret_cast<float4 *>(dst), make_float4(l.x == k ? 1.f : 0.f, l.y == k ? 1.f : 0.f, l.z == k ? 1.f : 0.f, l.w == k ? 1.f : 0.f));
    // F.one_hot raises on class values >= num_classes: flag it (the planes of such a cell are all zero)
    if (ch == 1 && err != nullptr && (l.x >= num_classes || l.y >= num_classes || l.z >= num_classes || l.w >= num_classes)) atomicOr(err, 1u);
}
// scalar variant for maps whose plane size is not a multiple of 4 (or unaligned pointers)
__global__ void __launch_bounds__(256) k_map_features_scalar(const uint8_t *__restrict__ occ, const uint8_t *__restrict__ sem, long long ncell,
                                                             int planes_cells, int num_classes, float *__restrict__ out,
                                                             uint32_t *__restrict__ err) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ncell) return;
    const int o = occ[i], l = sem[i];
    const long long b = i / planes_cells, within = i - b * planes_cells;
    float *dst = out + (b * (1 + num_classes)) * (long long)planes_cells + within;
    dst[0] = (float)o;
    for (int k = 0; k < num_classes; ++k) dst[(long long)(k + 1) * planes_cells] = l == k ? 1.f : 0.f;
    if (err != nullptr && l >= num_classes) atomicOr(err, 1u);
}

// ------------------------------------------------------------------ segmentation front end
// PredictSemantics.forward up to the network (mapper.py:715-736, 788-793): rgb u8 -> /255 -> bilinear resize to the depth
// size (F.interpolate, align_corners=False) -> (x - mean) / std per channel, and depth -> (d - 0.213) / 0.285, as ONE
// kernel: one thread per output pixel and all three channels (the 4 source texels of a pixel are 3-byte neighbours in
// the NHWC sensor layout, read through arbitrary element strides so that an NCHW tensor works too).  Same operation
// order as torch's upsample_bilinear2d: h0*(w0*v00 + w1*v01) + h1*(w0*v10 + w1*v11).
struct IvmPreArgs {
    const uint8_t *rgb; long long sb, sc, sy, sx;   // element strides of [B,3,h,w]
    int B, h, w, H, W;
    const float *depth; float *rgb_out, *depth_out;
    float mean[3], std[3], dmean, dstd;
    float scale_y, scale_x;
};
__global__ void __launch_bounds__(256) k_rednet_preprocess(const IvmPreArgs a) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long hw = (long long)a.H * a.W;
    if (i >= hw * a.B) return;
    const int b = (int)(i / hw);
    const int p = (int)(i - (long long)b * hw);
    const int y = p / a.W, x = p - y * a.W;
    if (a.rgb) {
        // area_pixel_compute_source_index(scale, dst, align_corners=false): scale * (dst + 0.5) - 0.5, clamped at 0
        float fy = a.scale_y * ((float)y + 0.5f) - 0.5f; fy = fy < 0.f ? 0.f : fy;
        float fx = a.scale_x * ((float)x + 0.5f) - 0.5f; fx = fx < 0.f ? 0.f : fx;
        const int y0 = min((int)fy, a.h - 1), x0 = min((int)fx, a.w - 1);
        const int y1 = y0 + (y0 < a.h - 1 ? 1 : 0), x1 = x0 + (x0 < a.w - 1 ? 1 : 0);
        const float ly = fy - (float)y0, lx = fx - (float)x0, hy = 1.f - ly, hx = 1.f - lx;
        const uint8_t *base = a.rgb + (long long)b * a.sb;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const uint8_t *pc = base + (long long)c * a.sc;
            const float v00 = (float)pc[y0 * a.sy + x0 * a.sx] / 255.0f, v01 = (float)pc[y0 * a.sy + x1 * a.sx] / 255.0f;
            const float v10 = (float)pc[y1 * a.sy + x0 * a.sx] / 255.0f, v11 = (float)pc[y1 * a.sy + x1 * a.sx] / 255.0f;
            const float v = hy * (hx * v00 + lx * v01) + ly * (hx * v10 + lx * v11);
            a.rgb_out[((long long)b * 3 + c) * hw + p] = (v - a.mean[c]) / a.std[c];
        }
    }
    if (a.depth) a.depth_out[i] = (a.depth[i] - a.dmean) / a.dstd;
}

// ------------------------------------------------------------------ host side
struct ivm_ctx {
    ivm_config cfg;
    IvmParams P;       // persistent part filled at create
    uint32_t step;     // 24-bit stamp of the last call
    unsigned long long cstep;  // calls so far (never rebased): the candidate-plane stamp cycles with it
    int hi_water;      // envs [0, hi_water) may hold records
    int first_call;
    int bulk_attr_set;
    int num_sms;
    int coop;             // device supports cooperative launches
    int ovl_grid[2];      // co-resident CTAs of k_step_overlap<false/true> (0 = not queried yet)
    size_t ovl_smem[2];
    uint32_t bar_base;    // value of IvmGlobal.bar_count before the next fused launch
    uint32_t team_base;   // same for the fix-up team's arrival counter (k_step_overlap)
    uint32_t kstep;       // known-map mode: steps so far (parity selects the rastered-points counter)
    uint32_t done_base;   // same for the CTAs-done counter: CTAs of all fused launches so far
    int pipelined;        // ivm_set_pipelined: consecutive steps may overlap (see ivln_map.h)
    int last_fused;       // the last iterative step ran as the persistent kernel (its rastered-record count is in stat_in)
    int64_t launches;
    // known-mode scratch
    uint32_t *kfill, *ktotals;
    size_t kcells;
    // export scratch
    uint32_t *rowlive, *rowtotals;
    // timing
    int timing;
    cudaEvent_t ev[IVM_EVPOOL][IVM_NSTAGES][2];
    int ev_used;
    int ev_created;
    float stage_ms[IVM_NSTAGES];
    int stage_launches[IVM_NSTAGES];
    char err[256];
};

static size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

struct Carver {
    char *base; size_t off;
    template <class T> T *take(size_t count) {
        T *p = base ? reinterpret_cast<T *>(base + off) : nullptr;
        off += align_up(count * sizeof(T));
        return p;
    }
};

static uint32_t edge_capacity(const ivm_config *c) {
    long long cap = (long long)c->max_envs * 8192;
    if (cap > (1ll << 20)) cap = 1ll << 20;
    return (uint32_t)cap;
}
static uint32_t hash_size(uint32_t ecap) { uint32_t h = 1024; while (h < 2 * ecap) h <<= 1; return h; }

// Ego tile of one raster group.  0 = choose.  Tiles are handed out one per group (then dynamically), so what counts
// is the number of WAVES the context's envs need on the raster groups of one B200 (2 x 148 x 3 groups with a score
// stream, 3 x 148 x 2 with GT labels, less the fix-up team's: ~840) and the cost of one tile: a fixed part (zeroing,
// spans, write-out, ~1.6 us) plus the half-cells under the rotated tile (its area with a margin of ~3 cells all
// round).  The grid of ny x nx tiles per env with the smallest waves x cost is taken; a partial wave counts as a
// whole one below three waves.  Measured (us per step, 16 / 32 envs, round 1): 16x16 65.5 / 78.5, 24x16 61.4 / 75.9,
// 32x16 61.4 / 80.6, 32x32 67.2 / 75.4; 1 env: 8x8 37.2, 16x16 39.6.  This picks 22x16 (16 envs), 26x26 (32), 8x8 (1).
static void tile_dims(const ivm_config *c, int *tr_out, int *tc_out) {
    int tr = c->tile_rows, tc = c->tile_cols;
    if ((tr <= 0 || tc <= 0) && c->mode == 1) {
        tr = 16; tc = 16;  // known-map mode: one 128-thread CTA per tile, 16 CTAs per SM -- many small tiles keep the SMs full
    } else if (tr <= 0 || tc <= 0) {
        const double groups = 840.0, fixed = 150.0, margin = 6.0;
        double best = 1.0e300;
        tr = c->map_rows < 64 ? c->map_rows : 64; tc = c->map_cols < 64 ? c->map_cols : 64;
        for (int ny = 1; ny <= 32; ++ny)
            for (int nx = 1; nx <= 32; ++nx) {
                const int r = (c->map_rows + ny - 1) / ny, q = (c->map_cols + nx - 1) / nx;
                if (r > 64 || q > 64 || (r < 8 && r < c->map_rows) || (q < 8 && q < c->map_cols)) continue;
                if ((c->map_rows + r - 1) / r != ny || (c->map_cols + q - 1) / q != nx) continue;  // same tile, fewer of them
                const double waves = (double)c->max_envs * ny * nx / groups;
                const double w = waves < 3.0 ? ceil(waves) : waves + 0.5;
                const double cost = w * (fixed + (r + margin) * (q + margin)) * (1.0 + 0.02 * fabs((double)r - q) / (r + q));
                if (cost < best) { best = cost; tr = r; tc = q; }
            }
    }
    if (tr > c->map_rows) tr = c->map_rows;
    if (tc > c->map_cols) tc = c->map_cols;
    *tr_out = tr; *tc_out = tc;
}

static void carve(const ivm_config *c, void *ws, IvmParams *P, ivm_ctx *ctx, size_t *total) {
    Carver cv{(char *)ws, 0};
    const size_t B = c->max_envs, SR = c->store_rows, SC = c->store_cols;
    const uint32_t ecap = edge_capacity(c), hs = hash_size(ecap);
    IvmParams q;
    memset(&q, 0, sizeof(q));
    q.g = cv.take<IvmGlobal>(1);
    q.bar = cv.take<uint32_t>(128);
    q.cta_trace = cv.take<unsigned long long>((size_t)IVM_TRACE_CTAS * IVM_TRACE_SLOTS);
    q.env = cv.take<IvmEnv>(B);
    float *xs = cv.take<float>(c->width > 0 ? c->width : 1);
    float *ys = cv.take<float>(c->height > 0 ? c->height : 1);
    q.xs = xs; q.ys = ys;
    q.T12_buf = cv.take<float>(12 * B);
    q.cs_buf = cv.take<float>(2 * B);
    q.rowcount = cv.take<int32_t>(B * SR);
    q.colcount = cv.take<int32_t>(B * SC);
    q.segs = cv.take<int32_t>(16 * B);
    {
        int tr = c->tile_rows, tc = c->tile_cols;
        tile_dims(c, &tr, &tc);
        q.tile_dirty = cv.take<uint32_t>(B * (size_t)((c->map_rows + tr - 1) / tr) * (size_t)((c->map_cols + tc - 1) / tc));
    }
    q.e1 = cv.take<IvmEdge>(ecap);
    q.e2 = cv.take<IvmEdge>(ecap);
    q.ecap = ecap;
    q.hkeys = cv.take<unsigned long long>(hs);
    q.hxord = cv.take<unsigned long long>(hs);
    q.hbest = cv.take<uint32_t>(hs);
    q.hmask = hs - 1;
    uint32_t *rowlive = cv.take<uint32_t>(B * SR);
    uint32_t *rowtotals = cv.take<uint32_t>((B * SR + 4095) / 4096 + 1);
    uint32_t *kfill = nullptr, *ktotals = nullptr;
    if (c->mode == 0) {
        q.store = cv.take<IvmRecord>(B * SR * SC);
        q.cplane = cv.take<unsigned long long>(B * SR * SC);
    } else {
        q.kcap = c->known_capacity;
        q.kpts = cv.take<IvmRecord>(B * (size_t)c->known_capacity);
        q.koff = cv.take<uint32_t>(B * (SR * SC + 1));
        kfill = cv.take<uint32_t>(SR * SC);
        ktotals = cv.take<uint32_t>((SR * SC + 1 + 4095) / 4096 + 1);
    }
    if (P) *P = q;
    if (ctx) { ctx->kfill = kfill; ctx->ktotals = ktotals; ctx->kcells = SR * SC; ctx->rowlive = rowlive; ctx->rowtotals = rowtotals; }
    if (total) *total = cv.off;
}

static int valid_config(const ivm_config *c) {
    if (!c) return 0;
    if (c->max_envs < 1 || c->map_rows < 1 || c->map_cols < 1) return 0;
    if (c->store_rows < 8 || c->store_cols < 8) return 0;
    if ((long long)c->store_rows * c->store_cols > (1ll << 24)) return 0;  // raster key = cell index << 8 | label
    if (!(c->res > 0.f) || !(c->half_res > 0.f)) return 0;
    if (c->mode == 0 && (c->height < 1 || c->width < 1 || (long long)c->height * c->width > (1ll << 24))) return 0;
    if (c->mode == 1 && (c->known_capacity < 1 || c->known_capacity >= (1ll << 24))) return 0;
    if (c->mode != 0 && c->mode != 1) return 0;
    return 1;
}

extern "C" {

const char *ivm_version(void) { return "ivlnmap 0.1 (sm_100a)"; }

size_t ivm_workspace_bytes(const ivm_config *cfg) {
    if (!valid_config(cfg)) return 0;
    size_t total = 0;
    carve(cfg, nullptr, nullptr, nullptr, &total);
    return total;
}

int ivm_create(const ivm_config *cfg, void *workspace_dev, size_t workspace_bytes, ivm_ctx **out) {
    if (!out || !valid_config(cfg) || !workspace_dev) return IVM_E_INVALID;
    if (((uintptr_t)workspace_dev & 255) != 0) return IVM_E_WORKSPACE;
    size_t need = 0;
    carve(cfg, nullptr, nullptr, nullptr, &need);
    if (workspace_bytes < need) return IVM_E_WORKSPACE;
    ivm_ctx *ctx = new ivm_ctx();
    memset(ctx, 0, sizeof(*ctx));
    ctx->cfg = *cfg;
    carve(cfg, workspace_dev, &ctx->P, ctx, nullptr);
    IvmParams &P = ctx->P;
    P.H = cfg->height; P.W = cfg->width; P.HW = cfg->height * cfg->width;
    P.R = cfg->map_rows; P.C = cfg->map_cols;
    P.res = cfg->res; P.half_res = cfg->half_res; P.half_h = cfg->half_h; P.half_w = cfg->half_w;
    P.inv_res = 1.0f / cfg->res; P.inv_half_res = 1.0f / cfg->half_res;  // IEEE single division on the host
    P.SR = cfg->store_rows; P.SC = cfg->store_cols; P.maxB = cfg->max_envs;
    int tr = 0, tc = 0;
    tile_dims(cfg, &tr, &tc);
    P.tile_r = tr; P.tile_c = tc;
    P.debug = cfg->reserved[1];
    P.pix_bits = ivm_pix_bits((long long)cfg->height * cfg->width);
    P.w_shift = -1;
    for (int sft = 0; sft < 30; ++sft)
        if ((1 << sft) == cfg->width) P.w_shift = sft;
    ctx->first_call = 1;
    ctx->num_sms = 148;
    {
        int dev = 0, sms = 0;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0)
            ctx->num_sms = sms;
        int coop = 0;
        if (cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev) == cudaSuccess) ctx->coop = coop;
        cudaGetLastError();
    }
    *out = ctx;
    return IVM_OK;
}

int ivm_destroy(ivm_ctx *ctx) {
    if (!ctx) return IVM_E_INVALID;
    if (ctx->ev_created)
        for (int i = 0; i < IVM_EVPOOL; ++i)
            for (int s = 0; s < IVM_NSTAGES; ++s) { cudaEventDestroy(ctx->ev[i][s][0]); cudaEventDestroy(ctx->ev[i][s][1]); }
    delete ctx;
    return IVM_OK;
}

static int cuda_fail(ivm_ctx *ctx, cudaError_t e, const char *where) {
    snprintf(ctx->err, sizeof(ctx->err), "%s: %s", where, cudaGetErrorString(e));
    return IVM_E_CUDA;
}
#define IVM_CHECK_LAUNCH(where)                                   \
    do {                                                          \
        cudaError_t _e = cudaGetLastError();                      \
        if (_e != cudaSuccess) return cuda_fail(ctx, _e, where);  \
    } while (0)

const char *ivm_last_cuda_error(const ivm_ctx *ctx) { return ctx ? ctx->err : "null context"; }
int64_t ivm_kernel_launches(const ivm_ctx *ctx) { return ctx ? ctx->launches : 0; }

int ivm_set_camera(ivm_ctx *ctx, const float *xs_dev, const float *ys_dev, ivm_stream_t stream) {
    if (!ctx || !xs_dev || !ys_dev) return IVM_E_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemcpyAsync((void *)ctx->P.xs, xs_dev, sizeof(float) * ctx->P.W, cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "set_camera xs");
    e = cudaMemcpyAsync((void *)ctx->P.ys, ys_dev, sizeof(float) * ctx->P.H, cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "set_camera ys");
    return IVM_OK;
}

// ---- timing helpers
static void timing_flush(ivm_ctx *ctx) {
    for (int i = 0; i < ctx->ev_used; ++i)
        for (int s = 0; s < IVM_NSTAGES; ++s) {
            cudaEventSynchronize(ctx->ev[i][s][1]);
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, ctx->ev[i][s][0], ctx->ev[i][s][1]) == cudaSuccess) {
                ctx->stage_ms[s] += ms;
                ctx->stage_launches[s] += 1;
            }
        }
    ctx->ev_used = 0;
    cudaGetLastError();
}
static int timing_slot(ivm_ctx *ctx) {
    if (!ctx->timing) return -1;
    if (!ctx->ev_created) {
        for (int i = 0; i < IVM_EVPOOL; ++i)
            for (int s = 0; s < IVM_NSTAGES; ++s) { cudaEventCreate(&ctx->ev[i][s][0]); cudaEventCreate(&ctx->ev[i][s][1]); }
        ctx->ev_created = 1;
    }
    if (ctx->ev_used == IVM_EVPOOL) timing_flush(ctx);
    return ctx->ev_used++;
}
#define T_BEGIN(stage) do { if (slot >= 0) cudaEventRecord(ctx->ev[slot][stage][0], st); } while (0)
#define T_END(stage) do { if (slot >= 0) cudaEventRecord(ctx->ev[slot][stage][1], st); } while (0)

int ivm_set_timing(ivm_ctx *ctx, int32_t enabled) {
    if (!ctx) return IVM_E_INVALID;
    if (!enabled && ctx->timing) timing_flush(ctx);
    ctx->timing = enabled ? 1 : 0;
    return IVM_OK;
}
int ivm_stage_times(ivm_ctx *ctx, float *ms_out5, int32_t *launches_out5, int32_t reset) {
    if (!ctx) return IVM_E_INVALID;
    timing_flush(ctx);
    for (int s = 0; s < IVM_NSTAGES; ++s) {
        if (ms_out5) ms_out5[s] = ctx->stage_ms[s];
        if (launches_out5) launches_out5[s] = ctx->stage_launches[s];
        if (reset) { ctx->stage_ms[s] = 0.f; ctx->stage_launches[s] = 0; }
    }
    return IVM_OK;
}

static int next_step(ivm_ctx *ctx) {
    if (ctx->step >= 0xFFFFFFu) return IVM_E_STEP_OVERFLOW;
    ctx->step += 1;
    return IVM_OK;
}

static int raster_max_rows(const IvmParams &P);
static void launch_raster(ivm_ctx *ctx, const IvmParams &P, cudaStream_t st, bool known) {
    dim3 grid((P.C + P.tile_c - 1) / P.tile_c, (P.R + P.tile_r - 1) / P.tile_r, P.B);
    const int max_rows = raster_max_rows(P);
    const size_t smem = raster_smem_bytes(P.tile_r, P.tile_c, max_rows);
    if (known) {
        const size_t ksmem = (size_t)P.tile_r * P.tile_c * 5 + (size_t)(2 * max_rows + 1) * 4 + 32;
        k_raster_known<<<grid, IVM_RASTER_THREADS, ksmem, st>>>(P, max_rows, (int)(ctx->kstep & 1u));
    } else {
        k_raster<false><<<grid, IVM_RASTER_THREADS, smem, st>>>(P, max_rows);
    }
    ctx->launches += 1;
}

static int raster_max_rows(const IvmParams &P) {
    // half-rows under a rotated tile: its diagonal in half-cells + the conservative margins
    const float diag = sqrtf((float)(P.tile_r * P.tile_r + P.tile_c * P.tile_c)) * (P.res / P.half_res);
    return (int)diag + 12;
}

// The fused persistent kernel applies when the image tiles evenly and the inputs are aligned.
static bool fused_applies(const ivm_ctx *ctx, const IvmParams &P, const float *depth, const uint8_t *labels, const float *logits) {
    if (!ctx->coop || ctx->cfg.reserved[0] != 0) return false;
    if (P.HW % IVM_F_TILE != 0 || (P.W & 1) || P.HW > (1 << 24)) return false;
    if (((uintptr_t)depth & 7) || ((uintptr_t)labels & 1) || (logits && ((uintptr_t)logits & 15))) return false;
    const size_t rb = 2 * raster_smem_bytes(P.tile_r, P.tile_c, raster_max_rows(P));
    return rb <= 96 * 1024;
}

// The overlapped kernel additionally needs 4-pixel groups inside one image row, 16-byte aligned depth and
// store coordinates that fit 16 bits each.
static bool overlap_applies(const ivm_ctx *ctx, const IvmParams &P, const float *depth, const uint8_t *labels, const float *logits,
                            const uint8_t *labels_out) {
    if (!fused_applies(ctx, P, depth, labels, logits)) return false;
    if ((P.W & 3) || ((uintptr_t)depth & 15) || ((uintptr_t)P.xs & 15) || P.SR > 65535 || P.SC > 65535) return false;
    if ((long long)P.B * (P.HW / IVM_O_TILE) > (1ll << 30)) return false;
    if (logits && ((uintptr_t)labels_out & 3)) return false;
    return true;
}

static int launch_overlap(ivm_ctx *ctx, IvmParams &P, const float *logits, int ncls, uint8_t *labels_out, int nenv_total,
                          cudaStream_t st) {
    const int pred = logits ? 1 : 0;
    int max_rows = raster_max_rows(P);
    // staging area of a raster group: the half-cells under a rotated tile (its area in half-cells + a margin of one
    // half-cell all round) with 25 % head-room, within what two CTAs per SM leave (tiles that still exceed it take
    // the direct path)
    const double hc = (double)P.res / (double)P.half_res;
    int stage_cap = (int)(1.25 * (P.tile_r * hc + 3.0) * (P.tile_c * hc + 3.0)) + 64;
    const int cap_max = (int)((110 * 1024 / 2 - (long)raster_smem_bytes(P.tile_r, P.tile_c, max_rows, 1)) / 16);
    if (stage_cap > cap_max) stage_cap = cap_max;
    if (stage_cap < 0 || ctx->cfg.reserved[2] != 2) stage_cap = 0;  // measured slower than the direct path (the raster is ALU-bound): opt-in only
    int group_bytes = (int)raster_smem_bytes(P.tile_r, P.tile_c, max_rows, stage_cap);
    size_t smem = (size_t)(pred ? IVM_O_RGROUPS_PRED : IVM_O_RGROUPS_GT) * group_bytes;
    const size_t scratch = (size_t)IVM_FIX_SMALL * 20 + 1024;  // fix-up scratch
    if (smem < scratch) smem = scratch;
    if (smem < ovl_stream_bytes(pred != 0)) smem = ovl_stream_bytes(pred != 0);
    const void *fn = pred ? (const void *)k_step_overlap<true> : (const void *)k_step_overlap<false>;
    if (!ctx->ovl_grid[pred] || ctx->ovl_smem[pred] != smem) {
        // The attribute belongs to the function and the device, not to this context: it only ever grows (a second
        // context with smaller tiles must not lower it under a context that launches with more).
        cudaError_t e = cudaSuccess;
        {
            static std::mutex mu;
            static size_t most[2][IVM_MAX_DEVICES];
            int dev = 0;
            cudaGetDevice(&dev);
            std::lock_guard<std::mutex> lock(mu);
            if (dev < 0 || dev >= IVM_MAX_DEVICES || most[pred][dev] < smem) {
                e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                if (e == cudaSuccess && dev >= 0 && dev < IVM_MAX_DEVICES) most[pred][dev] = smem;
            }
        }
        if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaFuncSetAttribute(k_step_overlap)");
        int per_sm = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, pred ? IVM_O_THREADS_PRED : IVM_O_THREADS_GT, smem);
        if (e != cudaSuccess) return cuda_fail(ctx, e, "occupancy(k_step_overlap)");
        if (per_sm < 1) { snprintf(ctx->err, sizeof(ctx->err), "k_step_overlap does not fit on an SM"); return IVM_E_CUDA; }
        const int want = pred ? IVM_O_CTAS_PER_SM_PRED : IVM_O_CTAS_PER_SM_GT;
        if (per_sm > want) per_sm = want;
        ctx->ovl_grid[pred] = per_sm * ctx->num_sms;
        ctx->ovl_smem[pred] = smem;
    }
    long long tiles = (long long)P.B * (P.HW / IVM_O_TILE);
    int grid = ctx->ovl_grid[pred];
    if (grid > tiles) grid = (int)tiles;
    uint32_t bar_base = ctx->bar_base;
    int team = grid / 16;  // CTAs that share the edge-line scan of the fix-up
    if (team > 16) team = 16;
    if (team < 1) team = 1;
    uint32_t team_base = ctx->team_base;
    int pipelined = ctx->pipelined;
    uint32_t done_target = ctx->done_base;
    void *args[] = {(void *)&P, (void *)&logits, (void *)&ncls, (void *)&labels_out, (void *)&nenv_total,
                    (void *)&bar_base, (void *)&max_rows, (void *)&group_bytes, (void *)&stage_cap, (void *)&team, (void *)&team_base,
                    (void *)&pipelined, (void *)&done_target};
    // A plain launch of at most as many CTAs as are co-resident (the grid barriers need that, and nothing else of
    // a cooperative launch), programmatically serialised behind the previous kernel of the stream: when that is
    // the previous step, this step's CTAs move in while its last ego tiles are still being rastered.
    cudaLaunchConfig_t lc;
    memset(&lc, 0, sizeof(lc));
    lc.gridDim = dim3(grid); lc.blockDim = dim3(pred ? IVM_O_THREADS_PRED : IVM_O_THREADS_GT);
    lc.dynamicSmemBytes = smem; lc.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = at; lc.numAttrs = (ctx->cfg.reserved[1] & 64) ? 0 : 1;  // debug bit 64: fully serialised launches
    cudaError_t e = cudaLaunchKernelExC(&lc, fn, args);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaLaunchKernelExC(k_step_overlap)");
    ctx->bar_base += 5u * (uint32_t)grid;
    ctx->done_base += (uint32_t)grid;
    if (team > 1) ctx->team_base += 2u * (uint32_t)team;
    ctx->launches += 1;
    return IVM_OK;
}

int ivm_step_iterative(ivm_ctx *ctx, int32_t num_envs, const float *depth, const uint8_t *labels, const float *logits,
                       int32_t num_classes, uint8_t *labels_out, const float *T12, const float *pose, const float *cs,
                       const void *orientation, int32_t orientation_is_f64, const uint8_t *masks, uint8_t *occ,
                       uint8_t *sem, ivm_stream_t stream) {
    if (!ctx || ctx->cfg.mode != 0) return IVM_E_INVALID;
    if (num_envs < 1 || num_envs > ctx->cfg.max_envs) return IVM_E_INVALID;
    if (!depth || !pose || !masks || !occ || !sem) return IVM_E_INVALID;
    if (!orientation && (!T12 || !cs)) return IVM_E_INVALID;
    if (!labels && !(logits && labels_out && num_classes >= 1 && num_classes <= 256)) return IVM_E_INVALID;
    int rc = next_step(ctx);
    if (rc != IVM_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    IvmParams P = ctx->P;
    P.B = num_envs; P.step = ctx->step;
    ctx->cstep += 1ull;
    const uint32_t period = ivm_stamp_period(P.pix_bits, (uint32_t)(ctx->cfg.reserved[3] > 0 ? ctx->cfg.reserved[3] : 0));
    P.cstamp = (uint32_t)((ctx->cstep - 1ull) % period) + 1u;
    if (P.cstamp == 1u && ctx->cstep > 1ull) {  // the stamp wrapped: forget the candidates of the last `period` steps
        cudaError_t e = cudaMemsetAsync(P.cplane, 0, sizeof(unsigned long long) * (size_t)ctx->cfg.max_envs * P.SR * P.SC,
                                        (cudaStream_t)stream);
        if (e != cudaSuccess) return cuda_fail(ctx, e, "candidate plane clear");
    }
    P.depth = depth; P.labels = labels ? labels : labels_out; P.pose = pose; P.masks = masks;
    P.occ = occ; P.sem = sem;
    if (orientation) { P.orient = orientation; P.orient_f64 = orientation_is_f64; P.T12 = P.T12_buf; P.cs = P.cs_buf; }
    else { P.orient = nullptr; P.T12 = T12; P.cs = cs; }
    const int slot = timing_slot(ctx);

    if (ctx->first_call) {
        k_init<<<64, 256, 0, st>>>(P);
        IVM_CHECK_LAUNCH("k_init");
        ctx->first_call = 0;
        ctx->launches += 1;
    }
    const int nenv = num_envs > ctx->hi_water ? num_envs : ctx->hi_water;  // envs >= num_envs get wiped
    ctx->hi_water = num_envs;

    if (ctx->cfg.reserved[0] == 0 && overlap_applies(ctx, P, depth, P.labels, logits, labels_out)) {
        if (!ctx->last_fused) {  // (the slots are kept zeroed by the fused steps themselves, one step ahead)
            cudaError_t e = cudaMemsetAsync(P.g->stat_in, 0, sizeof(P.g->stat_in), st);
            if (e != cudaSuccess) return cuda_fail(ctx, e, "memset(stat_in)");
            ctx->last_fused = 1;
        }
        T_BEGIN(1);
        rc = launch_overlap(ctx, P, logits, num_classes, labels_out, nenv, st);
        T_END(1);
        return rc;
    }
    ctx->last_fused = 0;
    const bool vec4 = (P.W % 4 == 0) && (((uintptr_t)depth & 15) == 0) && (((uintptr_t)P.labels & 3) == 0) &&
                      (!logits || ((uintptr_t)logits & 15) == 0);
    const int vec = vec4 ? 4 : 1;
    dim3 grid((P.HW + IVM_THREADS * vec - 1) / (IVM_THREADS * vec), nenv);
    const bool bulk = logits && vec4 && (P.HW % IVM_BULK_TILE == 0) && ctx->cfg.reserved[0] != 1;  // 1 = register-staged, 2 = bulk
    T_BEGIN(1);
    if (bulk) {
        const size_t smem = (size_t)IVM_BULK_NSTAGE * IVM_BULK_SP * IVM_BULK_TILE * sizeof(float);
        if (!ctx->bulk_attr_set) {
            cudaError_t e = cudaFuncSetAttribute(k_ingest_scatter_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaFuncSetAttribute(k_ingest_scatter_bulk)");
            ctx->bulk_attr_set = 1;
        }
        long long tiles = (long long)num_envs * (P.HW / IVM_BULK_TILE);
        int ctas = ctx->num_sms * IVM_BULK_CTAS_PER_SM;
        if (ctas > tiles) ctas = (int)tiles;
        k_ingest_scatter_bulk<<<ctas, IVM_THREADS, smem, st>>>(P, logits, num_classes, labels_out, nenv);
    } else if (logits) {
        if (vec4) k_ingest_scatter<true, 4><<<grid, IVM_THREADS, 0, st>>>(P, logits, num_classes, labels_out);
        else k_ingest_scatter<true, 1><<<grid, IVM_THREADS, 0, st>>>(P, logits, num_classes, labels_out);
    } else {
        if (vec4) k_ingest_scatter<false, 4><<<grid, IVM_THREADS, 0, st>>>(P, nullptr, 0, nullptr);
        else k_ingest_scatter<false, 1><<<grid, IVM_THREADS, 0, st>>>(P, nullptr, 0, nullptr);
    }
    T_END(1);
    IVM_CHECK_LAUNCH("k_ingest_scatter");
    grid.y = num_envs;
    T_BEGIN(2);
    if (vec4) k_ingest_resolve<4><<<grid, IVM_THREADS, 0, st>>>(P);
    else k_ingest_resolve<1><<<grid, IVM_THREADS, 0, st>>>(P);
    T_END(2);
    IVM_CHECK_LAUNCH("k_ingest_resolve");
    T_BEGIN(3);
    k_fixup<<<1, 1024, 0, st>>>(P);
    T_END(3);
    IVM_CHECK_LAUNCH("k_fixup");
    T_BEGIN(4);
    launch_raster(ctx, P, st, false);
    T_END(4);
    IVM_CHECK_LAUNCH("k_raster");
    ctx->launches += 3;
    return IVM_OK;
}

int ivm_known_clear(ivm_ctx *ctx, int32_t env, ivm_stream_t stream) {
    if (!ctx || ctx->cfg.mode != 1 || env < 0 || env >= ctx->cfg.max_envs) return IVM_E_INVALID;
    k_known_reset_env<<<1, 1, 0, (cudaStream_t)stream>>>(ctx->P, env, 0, 0, 0);
    ctx->launches += 1;
    IVM_CHECK_LAUNCH("k_known_reset_env");
    return IVM_OK;
}

int ivm_known_load(ivm_ctx *ctx, int32_t env, int64_t n, const float *xyz, const uint8_t *sem, int32_t origin_row,
                   int32_t origin_col, ivm_stream_t stream) {
    if (!ctx || ctx->cfg.mode != 1 || env < 0 || env >= ctx->cfg.max_envs) return IVM_E_INVALID;
    if (n < 0 || n > ctx->cfg.known_capacity || (n > 0 && (!xyz || !sem))) return IVM_E_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    const IvmParams &P = ctx->P;
    const size_t ncell1 = ctx->kcells + 1;
    uint32_t *off = P.koff + (size_t)env * ncell1;
    cudaError_t e = cudaMemsetAsync(off, 0, ncell1 * sizeof(uint32_t), st);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "known_load memset off");
    e = cudaMemsetAsync(ctx->kfill, 0, ctx->kcells * sizeof(uint32_t), st);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "known_load memset fill");
    k_known_reset_env<<<1, 1, 0, st>>>(P, env, n, origin_row, origin_col);
    if (n > 0) {
        const int blocks = (int)((n + 255) / 256 < 4096 ? (n + 255) / 256 : 4096);
        k_known_hist<<<blocks, 256, 0, st>>>(P, env, n, xyz, origin_row, origin_col, off);
        const int nb = (int)((ncell1 + 4095) / 4096);
        k_scan_local<<<nb, 1024, 0, st>>>(off, ncell1, ctx->ktotals);
        k_scan_totals<<<1, 1024, 0, st>>>(ctx->ktotals, nb);
        k_scan_add<<<nb, 1024, 0, st>>>(off, ncell1, ctx->ktotals);
        k_known_scatter<<<blocks, 256, 0, st>>>(P, env, n, xyz, sem, origin_row, origin_col, off, ctx->kfill);
        ctx->launches += 5;
    }
    ctx->launches += 1;
    IVM_CHECK_LAUNCH("known_load");
    return IVM_OK;
}

int ivm_step_known(ivm_ctx *ctx, int32_t num_envs, const float *pose, const float *cs, const void *orientation,
                   int32_t orientation_is_f64, uint8_t *occ, uint8_t *sem, ivm_stream_t stream) {
    if (!ctx || ctx->cfg.mode != 1) return IVM_E_INVALID;
    if (num_envs < 1 || num_envs > ctx->cfg.max_envs || !pose || !occ || !sem) return IVM_E_INVALID;
    if (!orientation && !cs) return IVM_E_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    IvmParams P = ctx->P;
    P.B = num_envs; P.pose = pose; P.cs = cs; P.occ = occ; P.sem = sem;
    const int slot = timing_slot(ctx);
    P.orient = nullptr;
    if (orientation) { P.orient = orientation; P.orient_f64 = orientation_is_f64; }
    ctx->kstep += 1u;
    T_BEGIN(4);
    launch_raster(ctx, P, st, true);
    T_END(4);
    IVM_CHECK_LAUNCH("k_raster<known>");
    return IVM_OK;
}

int ivm_export_world(ivm_ctx *ctx, int32_t num_envs, int64_t cap, int64_t *env_out, float *xyz_out, uint8_t *label_out,
                     uint64_t *key_out, uint64_t *count_dev, ivm_stream_t stream) {
    if (!ctx || ctx->cfg.mode != 0 || num_envs < 1 || num_envs > ctx->cfg.max_envs) return IVM_E_INVALID;
    if (!env_out || !xyz_out || !label_out || !key_out || !count_dev || cap < 0) return IVM_E_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    IvmParams P = ctx->P;
    P.B = num_envs;
    const int rows = num_envs * P.SR;
    const int blocks = (rows * 32 + 255) / 256;
    k_export_count<<<blocks, 256, 0, st>>>(P, ctx->rowlive);
    const int nb = (rows + 4095) / 4096;
    k_scan_local<<<nb, 1024, 0, st>>>(ctx->rowlive, (size_t)rows, ctx->rowtotals);
    k_scan_totals<<<1, 1024, 0, st>>>(ctx->rowtotals, nb);
    k_scan_add<<<nb, 1024, 0, st>>>(ctx->rowlive, (size_t)rows, ctx->rowtotals);
    k_export_write<<<blocks, 256, 0, st>>>(P, ctx->rowlive, cap, (long long *)env_out, xyz_out, label_out,
                                           (unsigned long long *)key_out, (unsigned long long *)count_dev);
    ctx->launches += 5;
    IVM_CHECK_LAUNCH("export_world");
    return IVM_OK;
}

int ivm_read_status(ivm_ctx *ctx, ivm_status *host_out, ivm_stream_t stream) {
    if (!ctx || !host_out) return IVM_E_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    IvmGlobal g;
    cudaError_t e = cudaMemcpyAsync(&g, ctx->P.g, sizeof(g), cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "read_status memcpy");
    e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "read_status sync");
    host_out->error_flags = g.err;
    host_out->pad = 0;
    for (int i = 0; i < 8; ++i) host_out->stats[i] = g.stats[i];
    if (ctx->cfg.mode == 1) host_out->stats[IVM_STAT_IN] = g.known_in[ctx->kstep & 1u];
    else if (ctx->last_fused) host_out->stats[IVM_STAT_IN] = g.stat_in[ctx->step & 3u];
    return IVM_OK;
}

int ivm_copy_error_flags_async(ivm_ctx *ctx, uint32_t *host_pinned_out, ivm_stream_t stream) {
    if (!ctx || !host_pinned_out) return IVM_E_INVALID;
    cudaError_t e = cudaMemcpyAsync(host_pinned_out, &ctx->P.g->err, sizeof(uint32_t), cudaMemcpyDeviceToHost, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "copy_error_flags_async");
    return IVM_OK;
}

int ivm_clear_error_flags(ivm_ctx *ctx, ivm_stream_t stream) {
    if (!ctx) return IVM_E_INVALID;
    cudaError_t e = cudaMemsetAsync(&ctx->P.g->err, 0, sizeof(uint32_t), (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "clear_error_flags");
    return IVM_OK;
}

int ivm_read_phase_ns(ivm_ctx *ctx, uint64_t *ns_out24, ivm_stream_t stream) {
    if (!ctx || !ns_out24) return IVM_E_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    IvmGlobal g;
    cudaError_t e = cudaMemcpyAsync(&g, ctx->P.g, sizeof(g), cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "read_phase_ns memcpy");
    e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "read_phase_ns sync");
    for (int i = 0; i < 8; ++i) ns_out24[i] = g.tstamp[i];
    for (int i = 0; i < 16; ++i) ns_out24[8 + i] = g.ttrace[i];
    return IVM_OK;
}

int ivm_read_cta_trace(ivm_ctx *ctx, uint64_t *ns_out, int32_t num_ctas, ivm_stream_t stream) {
    if (!ctx || !ns_out || num_ctas < 1 || num_ctas > IVM_TRACE_CTAS) return IVM_E_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemcpyAsync(ns_out, ctx->P.cta_trace, sizeof(uint64_t) * IVM_TRACE_SLOTS * (size_t)num_ctas,
                                    cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "read_cta_trace memcpy");
    e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "read_cta_trace sync");
    return IVM_OK;
}

int ivm_map_features(const uint8_t *occ, const uint8_t *sem, int32_t num_envs, int32_t rows, int32_t cols, int32_t num_classes,
                     float *out, uint32_t *err_flag_dev, ivm_stream_t stream) {
    if (!occ || !sem || !out || num_envs < 1 || rows < 1 || cols < 1 || num_classes < 1 || num_classes > 255) return IVM_E_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    const long long plane = (long long)rows * cols, ncell = plane * num_envs;
    if (plane > (1ll << 30)) return IVM_E_INVALID;
    const bool vec = (plane % 4 == 0) && (((uintptr_t)occ | (uintptr_t)sem) % 4 == 0) && ((uintptr_t)out % 16 == 0);
    if (vec) {
        const long long threads = ncell / 4;
        k_map_features<<<dim3((unsigned)((threads + 255) / 256), (unsigned)(1 + num_classes)), 256, 0, st>>>(occ, sem, ncell, (int)plane, num_classes, out, err_flag_dev);
    } else {
        k_map_features_scalar<<<(unsigned)((ncell + 255) / 256), 256, 0, st>>>(occ, sem, ncell, (int)plane, num_classes, out, err_flag_dev);
    }
    return cudaGetLastError() == cudaSuccess ? IVM_OK : IVM_E_CUDA;
}

int ivm_rednet_preprocess(const uint8_t *rgb, const int64_t *rgb_strides4, int32_t num_envs, int32_t rgb_h, int32_t rgb_w,
                          const float *depth, int32_t height, int32_t width, float *rgb_out, float *depth_out,
                          ivm_stream_t stream) {
    if (num_envs < 1 || height < 1 || width < 1) return IVM_E_INVALID;
    if (rgb && (!rgb_strides4 || !rgb_out || rgb_h < 1 || rgb_w < 1)) return IVM_E_INVALID;
    if (depth && !depth_out) return IVM_E_INVALID;
    if (!rgb && !depth) return IVM_OK;
    IvmPreArgs a;
    memset(&a, 0, sizeof(a));
    a.rgb = rgb; a.depth = depth; a.rgb_out = rgb_out; a.depth_out = depth_out;
    if (rgb) { a.sb = rgb_strides4[0]; a.sc = rgb_strides4[1]; a.sy = rgb_strides4[2]; a.sx = rgb_strides4[3]; }
    a.B = num_envs; a.h = rgb_h; a.w = rgb_w; a.H = height; a.W = width;
    a.mean[0] = 0.485f; a.mean[1] = 0.456f; a.mean[2] = 0.406f;   // mapper.py:725-727
    a.std[0] = 0.229f; a.std[1] = 0.224f; a.std[2] = 0.225f;
    a.dmean = 0.213f; a.dstd = 0.285f;                           // mapper.py:731-733
    a.scale_y = rgb ? (float)rgb_h / (float)height : 1.f;         // torch: area_pixel_compute_scale (no scale factor given)
    a.scale_x = rgb ? (float)rgb_w / (float)width : 1.f;
    const long long n = (long long)num_envs * height * width;
    k_rednet_preprocess<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a);
    return cudaGetLastError() == cudaSuccess ? IVM_OK : IVM_E_CUDA;
}

int ivm_copy_state(ivm_ctx *dst, const ivm_ctx *src, ivm_stream_t stream) {
    if (!dst || !src) return IVM_E_INVALID;
    const ivm_config &a = dst->cfg, &b = src->cfg;
    if (a.max_envs < b.max_envs || a.mode != b.mode || a.store_rows != b.store_rows || a.store_cols != b.store_cols ||
        a.height != b.height || a.width != b.width || a.known_capacity != b.known_capacity)
        return IVM_E_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t B = b.max_envs, SR = b.store_rows, SC = b.store_cols;
    ivm_ctx *ctx = dst;
#define IVM_COPY(dptr, sptr, bytes)                                                              \
    do {                                                                                         \
        cudaError_t _e = cudaMemcpyAsync((void *)(dptr), (const void *)(sptr), (bytes), cudaMemcpyDeviceToDevice, st); \
        if (_e != cudaSuccess) return cuda_fail(ctx, _e, "copy_state");                          \
    } while (0)
    IVM_COPY(dst->P.g, src->P.g, sizeof(IvmGlobal));
    dst->last_fused = 0;
    IVM_COPY(dst->P.env, src->P.env, sizeof(IvmEnv) * B);
    IVM_COPY(dst->P.xs, src->P.xs, sizeof(float) * (b.width > 0 ? b.width : 1));
    IVM_COPY(dst->P.ys, src->P.ys, sizeof(float) * (b.height > 0 ? b.height : 1));
    IVM_COPY(dst->P.rowcount, src->P.rowcount, sizeof(int32_t) * B * SR);
    IVM_COPY(dst->P.colcount, src->P.colcount, sizeof(int32_t) * B * SC);
    if (b.mode == 0) {
        IVM_COPY(dst->P.store, src->P.store, sizeof(IvmRecord) * B * SR * SC);
    } else {
        IVM_COPY(dst->P.kpts, src->P.kpts, sizeof(IvmRecord) * B * (size_t)b.known_capacity);
        IVM_COPY(dst->P.koff, src->P.koff, sizeof(uint32_t) * B * (SR * SC + 1));
    }
#undef IVM_COPY
    dst->step = src->step;
    dst->hi_water = src->hi_water;
    return IVM_OK;
}

int ivm_set_pipelined(ivm_ctx *ctx, int32_t enabled) {
    if (!ctx) return IVM_E_INVALID;
    ctx->pipelined = enabled ? 1 : 0;
    return IVM_OK;
}

int ivm_debug_set_step(ivm_ctx *ctx, uint32_t step) {
    if (!ctx || ctx->step != 0 || step >= 0xFFFFFFu) return IVM_E_INVALID;  // before the first step only
    ctx->step = step;
    return IVM_OK;
}

int ivm_rebase_stamps(ivm_ctx *ctx, ivm_stream_t stream) {
    if (!ctx || ctx->cfg.mode != 0) return IVM_E_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    k_rebase<<<148 * 8, 256, 0, st>>>(ctx->P);
    k_rebase_env<<<(ctx->P.maxB + 255) / 256, 256, 0, st>>>(ctx->P);
    ctx->launches += 2;
    ctx->step = 1;
    ctx->last_fused = 0;  // (the step numbering restarts: the per-step count slots are cleared before the next fused step)
    IVM_CHECK_LAUNCH("rebase");
    return IVM_OK;
}

}  // extern "C"
