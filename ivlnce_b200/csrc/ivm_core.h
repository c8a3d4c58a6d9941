// ivm_core.h -- data layout and per-thread logic of the B200 semantic-map update.
//
// Everything that decides a result bit lives here as host/device inline code:
// the CUDA kernels (ivm_kernels.cu) call it from parallel drivers, and the
// test-only serial emulator (tests/emu/emulator.cpp) calls the very same
// functions from plain loops, so the algorithm can be checked against the
// oracle on a machine without a GPU.
//
// What is computed (reference: ivlnce_baselines/common/mapping_module/mapper.py):
//   * the reference keeps a world POINT LIST, de-duplicated to the highest point
//     per half-resolution cell (mapper.py:428-474) and re-rasterised every step
//     (mapper.py:555-617);
//   * here the same state is a dense per-env grid of half-cells ("world store"),
//     one 16-byte record per cell, and the per-step work is bounded to the cells
//     a frame touches (ingest) and the cells under the ego window (raster);
//   * the reference's de-dup key is `b*(Rmax*Cmax) + r*Cmax + c` with strides
//     max instead of max+1 (mapper.py:468-469), so cells on the bounding-box
//     edge of a batch collide and are merged.  Only edge cells can collide (see
//     DESIGN.md), so the dense store reproduces this with a small "edge fix-up"
//     program per de-dup stage (ivm_fixup_program below).
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define IVM_HD __host__ __device__ __forceinline__
// Out-of-line on the device: rarely executed or called from several places.  The step kernel runs
// every phase once per launch, so its cost is dominated by first-touch instruction fetches; code
// that is not executed must not sit in the fetched path, and code used twice must exist once.
#define IVM_HD_COLD __host__ __device__ __noinline__
#else
#define IVM_HD inline
#define IVM_HD_COLD inline
#endif

#if defined(__CUDA_ARCH__)
#define IVM_TRACE(g, k, tid)                                                      \
    do {                                                                          \
        if ((tid) == 0) {                                                         \
            unsigned long long _t;                                                \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(_t));                \
            (g)->ttrace[k] = _t;                                                  \
        }                                                                         \
    } while (0)
#else
#define IVM_TRACE(g, k, tid) do { } while (0)
#endif

// ---------------------------------------------------------------------------
// Exactly rounded fp32 primitives.  The device versions are intrinsics that the
// compiler never contracts into FMAs; the host versions rely on
// -ffp-contract=off.  (SURVEY.md Appendix A pins where the reference fuses.)
#if defined(__CUDA_ARCH__)
IVM_HD float ivm_mul(float a, float b) { return __fmul_rn(a, b); }
IVM_HD float ivm_add(float a, float b) { return __fadd_rn(a, b); }
IVM_HD float ivm_sub(float a, float b) { return __fsub_rn(a, b); }
IVM_HD float ivm_div(float a, float b) { return __fdiv_rn(a, b); }
IVM_HD float ivm_fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
IVM_HD uint32_t ivm_f2u(float f) { return __float_as_uint(f); }
#else
IVM_HD float ivm_mul(float a, float b) { volatile float r = a * b; return r; }
IVM_HD float ivm_add(float a, float b) { volatile float r = a + b; return r; }
IVM_HD float ivm_sub(float a, float b) { volatile float r = a - b; return r; }
IVM_HD float ivm_div(float a, float b) { volatile float r = a / b; return r; }
IVM_HD float ivm_fma(float a, float b, float c) { return fmaf(a, b, c); }
IVM_HD uint32_t ivm_f2u(float f) { union { float f; uint32_t u; } v; v.f = f; return v.u; }
#endif

// rint(a / d) for a divisor that is fixed for the life of a context (cell size), without the IEEE division on the
// common path.  The reference computes q = fl(a / d) and rounds it to the nearest integer (ties to even).  With
// inv = fl(1 / d) and t = fl(a * inv):  |t - a/d| <= 2^-23 |a/d| (1 + eps)  and  |q - a/d| <= 2^-24 |a/d|,  hence
// |t - q| < 1.8e-7 |t| + (denormal slack).  If t lies further than 1e-6 + 4e-7 |t| from every half-integer, q lies in
// the same rounding interval and rint(q) == rint(t); otherwise (about 2 values in 10^6, and every NaN) `amb` is set
// and the caller redoes the value with the true division.  So the result is ALWAYS the reference's integer.
IVM_HD float ivm_rint_mul(float a, float inv, bool &amb) {
    const float t = ivm_mul(a, inv);
    const float r = rintf(t);
    const float e = fabsf(ivm_sub(t, r));
    amb = amb || !(e < fmaf(fabsf(t), -4.0e-7f, 0.499999f));
    return r;
}
IVM_HD float ivm_rint_div(float a, float d, float inv) {
    bool amb = false;
    const float r = ivm_rint_mul(a, inv, amb);
    return amb ? rintf(ivm_div(a, d)) : r;
}

// monotone float -> uint map (x < y  <=>  ord(x) < ord(y)); -0 and +0 map to the
// same value because the reference compares heights with a float '>'.
IVM_HD uint32_t ivm_orderable(float y) {
    uint32_t u = ivm_f2u(ivm_add(y, 0.0f));
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// ---------------------------------------------------------------------------
// Device-resident layout.

// One half-cell of an env's world store.  meta = (stamp << 8) | label; stamp is
// the 24-bit step counter of the call that wrote the record, 0 = never written /
// deleted.  A record is live iff stamp != 0 && stamp >= env.reset_stamp.
struct IvmRecord {
    float x, y, z;
    uint32_t meta;
};

// Frame candidate plane: one 64-bit word per half-cell of every env's store window,
//   word = stamp << (32 + pb) | orderable(height) << pb | (2^pb - 1 - pixel index),   pb = bits of a pixel index
// stamp = 1..2^(32-pb)-1 cycles with the step counter, so words of earlier steps lose every atomicMax of the
// current step: the plane is never cleaned between steps (the host clears it when the stamp wraps: once per
// 65 535 steps for 256x256 images).  Offering a point is ONE fire-and-forget 64-bit RED.MAX (no tag to claim,
// no dependent round trip); a frame touches ~11 k distinct cells per env, so the live part of the
// plane is a few hundred KB per env and stays in L2 between the scatter and the resolve.

struct IvmEnv {                   // 64 bytes
    int32_t origin_r, origin_c;   // absolute half-cell index of store cell (0,0)
    uint32_t reset_stamp;         // records older than this are dead (O(1) reset)
    int32_t count;                // live records
    int32_t rmin, rmax, cmin, cmax; // exact bbox of live records (absolute), if count>0
    int32_t dirty;                // bbox must be rebuilt from rowcount/colcount
    int32_t known_n;              // known-map mode: points in this env's CSR store
    int32_t nb[4];                // bbox under reconstruction (ivm_rebuild_dirty_boxes): rmin, rmax, cmin, cmax
    int32_t pad[2];
};

// A de-dup winner sitting on a bounding-box edge cell, waiting for the fix-up.
struct IvmEdge {
    float x, y, z;
    uint32_t label;
    int32_t b, r, c;              // env, absolute half-row / half-col
    uint32_t slot;                // hash slot of its collision class
    unsigned long long xorder;    // position in the reference's point list (tie-break)
    unsigned long long addr;      // record index in the store (stage 2)
};

#define IVM_STAT_VALID 0   // pixels surviving the depth+height filters (this step)
#define IVM_STAT_LOCAL 1   // survivors of the frame de-dup
#define IVM_STAT_WORLD 2   // live world records after the step
#define IVM_STAT_IN 3      // world records rasterised into the ego window
#define IVM_STAT_E1 4      // stage-1 edge candidates
#define IVM_STAT_E2 5      // stage-2 edge candidates
#define IVM_STAT_MERGED 6  // records deleted by edge collisions (cumulative)
#define IVM_NSTATS 8

#define IVM_ERR_STORE_OVERFLOW 1u  // a point fell outside an env's world store window
#define IVM_ERR_EDGE_OVERFLOW 2u   // edge list / hash capacity exceeded
#define IVM_ERR_KNOWN_OVERFLOW 4u  // known-map cloud larger than capacity / index range
#define IVM_ERR_GRID_BARRIER 8u    // a grid barrier of the fused step kernel timed out (results invalid)

struct IvmGlobal {
    int32_t loc[4];               // frame (stage-1) bbox over all envs: rmin,rmax,cmin,cmax
    int32_t glob[4];              // world (stage-2) bbox over all envs
    int32_t prev_valid, prev_rmin, prev_cmin, pad0;
    long long prev_R, prev_C;     // stage-2 bbox extents of the previous step (list order)
    uint32_t n_e1, n_e2;
    uint32_t err;
    uint32_t any_dirty;
    uint32_t n_seg;
    uint32_t scan_chunks;         // chunks of IVM_SCAN_CHUNK cells per edge-line segment (longest segment)
    uint32_t prev_n_seg, prev_scan_chunks;  // last step's edge-line segments (still in P.segs): prefetch hints for the scan
    unsigned long long acc_valid, acc_local;  // per-step accumulators (K1 / K2+F), published and zeroed by F
    unsigned long long acc_e1, acc_e2;        // direct path: frame-edge winners examined / live records on the world edge lines
    unsigned long long known_in[2];           // known-map mode: rastered points of the even / odd steps (one is counted into while
                                              // the other is zeroed for the next step)
    unsigned long long stats[IVM_NSTATS];     // published figures of the last step; stats[IN] accumulates in K4
    // fused step kernel only
    unsigned long long tstamp[8];             // %globaltimer at the phase boundaries of the last fused step:
                                              // 0 start, 1 ingest done, 2 resolve done, 3 fix-up done, 4 raster released,
                                              // 5 max over CTAs of the end time
    unsigned long long ttrace[16];            // %globaltimer at the milestones inside the fix-up program (device only)
    // fused step kernel, direct edge resolution: the edge-line scan of a step is skipped while nothing that could create
    // a collision has happened since the last one.  After a scan every key class on the world-box edge lines has one
    // survivor; a new member needs a cell ON one of those lines to become live (edge_touched), another world box
    // (scan_glob) or another batch (scan_B) -- resets and deletions only remove members.
    int32_t scan_glob[4];                     // world box of the last completed direct scan (16-byte aligned)
    uint32_t scan_valid;                      // scan_glob / scan_B describe the store as the last direct scan left it
    int32_t scan_B;
    uint32_t edge_touched;                    // a cell on an edge line of scan_glob became live since that scan
    uint32_t pad1;
    unsigned long long stat_in[4];            // fused step kernel: rastered records of step & 3 (the slot of the NEXT step is zeroed
                                              // at the start of a step, so that no CTA has to fence its count before it leaves)
};

static_assert(offsetof(IvmGlobal, scan_glob) % 16 == 0, "scan_glob is read with one 128-bit load");

struct IvmParams {
    // geometry
    int32_t H, W, HW;
    int32_t R, C;                 // ego map rows / cols
    float res, half_res, half_h, half_w;
    float inv_res, inv_half_res;  // fl(1 / res), fl(1 / half_res): ivm_rint_mul
    int32_t SR, SC;               // world store rows / cols (half-cells) per env
    int32_t maxB;
    int32_t tile_r, tile_c;       // ego tile of one raster CTA
    // persistent device memory
    IvmRecord *store;             // [maxB][SR][SC]
    unsigned long long *cplane;   // [maxB][SR][SC] frame candidate plane (see above)
    uint32_t cstamp;              // per step: 1..stamp period (see the candidate plane above)
    int32_t pix_bits;             // pb: bits of a pixel index, ceil(log2(HW)) (<= 24)
    int32_t w_shift;              // log2(W) if W is a power of two, else -1
    IvmEnv *env;                  // [maxB]
    int32_t *rowcount, *colcount; // [maxB][SR], [maxB][SC] live records per store row / col
    IvmGlobal *g;
    uint32_t *bar;                // grid-barrier arrival counter of the fused step kernel, alone in its own 256-byte
                                  // block (CTAs spin on it; nothing else may share its L2 slice line), monotone
                                  // across launches (the host tracks the base)
    unsigned long long *cta_trace; // [IVM_TRACE_CTAS][IVM_TRACE_SLOTS] %globaltimer stamps per CTA of the fused kernel
    IvmEdge *e1, *e2;             // edge lists, capacity ecap each
    uint32_t ecap;
    int32_t *segs;                // [4*maxB][4] edge-line segments to scan: b, is_col, line, pad
    unsigned long long *hkeys;    // open-addressing hash of collision classes
    uint32_t *hbest;              // best height per class
    unsigned long long *hxord;    // first list position among the best
    uint32_t hmask;
    const float *xs, *ys;         // camera scale tables [W], [H]
    float *T12_buf, *cs_buf;      // [maxB][12], [maxB][2]: matrices derived in K0 when the caller passes angles
    // known-map mode (CSR store)
    IvmRecord *kpts;              // [maxB][kcap] points sorted by half-cell
    uint32_t *koff;               // [maxB][SR*SC+1]
    long long kcap;
    // per step
    int32_t B;
    uint32_t step;
    const float *depth;           // [B][H][W] normalised
    const uint8_t *labels;        // [B][H][W]
    const float *T12;             // [B][12] rows 0..2 of camera->world
    const float *pose;            // [B][3]
    const float *cs;              // [B][2] cos(-heading), sin(-heading)
    const uint8_t *masks;         // [B] 0 = reset this env before ingesting
    const void *orient;           // [B][2] (elevation, heading), f64 or f32; NULL if T12/cs are given
    int32_t orient_f64;
    int32_t debug;                // profiling experiments only (config.reserved[1]); results are invalid when non-zero:
                                  // 1 = ingest phase without candidate inserts, 2 = without world-record prefetch
    uint8_t *occ, *sem;           // [B][R][C]
    uint32_t *tile_dirty;         // [maxB][ego tiles]: == step if the tile holds a record the edge fix-up may still change
};

#define IVM_EMPTY_KEY 0xFFFFFFFFFFFFFFFFull
#define IVM_TRACE_CTAS 1024
#define IVM_TRACE_SLOTS 16

// Record / candidate loads.  In the fused step kernel these locations are written by other SMs
// earlier in the SAME launch, so the device versions read through L2 (ld.global.cg) and never
// from a possibly stale L1 line.
// The world store and the candidate plane are touched sparsely but in (nearly) the same places step after
// step, while hundreds of MB of class scores stream through L2 in between: their accesses carry an
// L2 evict_last policy (the score stream is evict_first), so that the working set of a few MB per env
// stays L2-resident across steps instead of being re-fetched from HBM at loaded latency.
#if defined(__CUDACC__)
__device__ __forceinline__ unsigned long long ivm_policy_keep() {
    unsigned long long p;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
#endif
#if defined(__CUDA_ARCH__)
IVM_HD IvmRecord ivm_load_record(const IvmRecord *p) {
    uint4 v;
    asm volatile("ld.global.cg.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(ivm_policy_keep()));
    IvmRecord r;
    r.x = __uint_as_float(v.x); r.y = __uint_as_float(v.y); r.z = __uint_as_float(v.z); r.meta = v.w;
    return r;
}
IVM_HD void ivm_store_record(IvmRecord *p, const IvmRecord &r) {
    asm volatile("st.global.L2::cache_hint.v4.u32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "r"(__float_as_uint(r.x)),
                 "r"(__float_as_uint(r.y)), "r"(__float_as_uint(r.z)), "r"(r.meta), "l"(ivm_policy_keep()) : "memory");
}
IVM_HD unsigned long long ivm_load_ull(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.global.cg.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(v) : "l"(p), "l"(ivm_policy_keep()));
    return v;
}
IVM_HD uint32_t ivm_load_u32(const uint32_t *p) { return __ldcg(p); }
IVM_HD uint32_t ivm_load_meta(const IvmRecord *p) { return __ldcg(&p->meta); }
#else
IVM_HD uint32_t ivm_load_meta(const IvmRecord *p) { return p->meta; }
IVM_HD IvmRecord ivm_load_record(const IvmRecord *p) { return *p; }
IVM_HD void ivm_store_record(IvmRecord *p, const IvmRecord &r) { *p = r; }
IVM_HD unsigned long long ivm_load_ull(const unsigned long long *p) { return *p; }
IVM_HD uint32_t ivm_load_u32(const uint32_t *p) { return *p; }
#endif

IVM_HD bool ivm_live(uint32_t meta, uint32_t reset_stamp) {
    return meta != 0u && (meta >> 8) >= reset_stamp;
}

// ---------------------------------------------------------------------------
// Atomics policy.  Device: hardware atomics.  Host emulator: plain serial ops.
#if defined(__CUDA_ARCH__)
struct IvmAtomics {
    static __device__ __forceinline__ void add_i(int32_t *p, int32_t v) { atomicAdd(p, v); }
    static __device__ __forceinline__ uint32_t add_u(uint32_t *p, uint32_t v) { return atomicAdd(p, v); }
    static __device__ __forceinline__ void add_ull(unsigned long long *p, unsigned long long v) { atomicAdd(p, v); }
    static __device__ __forceinline__ void min_i(int32_t *p, int32_t v) { atomicMin(p, v); }
    static __device__ __forceinline__ void max_i(int32_t *p, int32_t v) { atomicMax(p, v); }
    static __device__ __forceinline__ void max_u(uint32_t *p, uint32_t v) { atomicMax(p, v); }
    static __device__ __forceinline__ void or_u(uint32_t *p, uint32_t v) { atomicOr(p, v); }
    static __device__ __forceinline__ void max_ull(unsigned long long *p, unsigned long long v) { atomicMax(p, v); }
    static __device__ __forceinline__ void min_ull(unsigned long long *p, unsigned long long v) { atomicMin(p, v); }
    static __device__ __forceinline__ unsigned long long cas_ull(unsigned long long *p, unsigned long long c,
                                                                  unsigned long long v) { return atomicCAS(p, c, v); }
    static __device__ __forceinline__ uint32_t cas_u(uint32_t *p, uint32_t c, uint32_t v) { return atomicCAS(p, c, v); }
    static __device__ __forceinline__ void sync() { __syncthreads(); }
};
#else
struct IvmAtomics {
    static void add_i(int32_t *p, int32_t v) { *p += v; }
    static uint32_t add_u(uint32_t *p, uint32_t v) { uint32_t o = *p; *p += v; return o; }
    static void add_ull(unsigned long long *p, unsigned long long v) { *p += v; }
    static void min_i(int32_t *p, int32_t v) { if (v < *p) *p = v; }
    static void max_i(int32_t *p, int32_t v) { if (v > *p) *p = v; }
    static void max_u(uint32_t *p, uint32_t v) { if (v > *p) *p = v; }
    static void or_u(uint32_t *p, uint32_t v) { *p |= v; }
    static void max_ull(unsigned long long *p, unsigned long long v) { if (v > *p) *p = v; }
    static void min_ull(unsigned long long *p, unsigned long long v) { if (v < *p) *p = v; }
    static unsigned long long cas_ull(unsigned long long *p, unsigned long long c, unsigned long long v) {
        unsigned long long o = *p; if (o == c) *p = v; return o;
    }
    static uint32_t cas_u(uint32_t *p, uint32_t c, uint32_t v) { uint32_t o = *p; if (o == c) *p = v; return o; }
    static void sync() {}
};
#endif

// ---------------------------------------------------------------------------
// Unprojection of one depth pixel: reference mapper.py:381-384 (x10),
// projector/core.py:117-149 (x = z*xs, y = z*ys), core.py:171 (T * [x y z 1],
// evaluated as mul + 3 fma per row, the order the reference's BLAS uses),
// filters mapper.py:416-424 (strict), de-dup cell mapper.py:464.
struct IvmPoint {
    float x, y, z;
    int32_t r, c;  // absolute half-row (from z) and half-col (from x)
};

// world coordinates of one depth pixel (shared by every caller, so that a point recomputed in a later
// phase is bit-identical to the one offered to the candidate plane)
IVM_HD void ivm_world_xyz(float d, float xs_u, float ys_v, const float *T, float &x, float &y, float &z) {
    const float zc = ivm_mul(d, 10.0f);
    const float xc = ivm_mul(zc, xs_u);
    const float yc = ivm_mul(zc, ys_v);
    float w[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        float acc = ivm_mul(T[4 * r + 0], xc);
        acc = ivm_fma(T[4 * r + 1], yc, acc);
        acc = ivm_fma(T[4 * r + 2], zc, acc);
        acc = ivm_fma(T[4 * r + 3], 1.0f, acc);
        w[r] = acc;
    }
    x = w[0]; y = w[1]; z = w[2];
}

// the height (row 1 of T) alone: the same operations in the same order as ivm_world_xyz, so the same bits
IVM_HD float ivm_world_y(float d, float xs_u, float ys_v, const float *T) {
    const float zc = ivm_mul(d, 10.0f);
    const float xc = ivm_mul(zc, xs_u);
    const float yc = ivm_mul(zc, ys_v);
    float acc = ivm_mul(T[4], xc);
    acc = ivm_fma(T[5], yc, acc);
    acc = ivm_fma(T[6], zc, acc);
    acc = ivm_fma(T[7], 1.0f, acc);
    return acc;
}

// returns 0 = filtered out, 1 = valid point, 2 = valid but its cell index is not representable
// (non-finite / absurd coordinates; the caller flags IVM_ERR_STORE_OVERFLOW)
IVM_HD int ivm_unproject(float d, float xs_u, float ys_v, const float *T, float h, float half_res, float inv, IvmPoint &p) {
    if (!(d > 0.01f && d < 0.99f)) return 0;
    float w[3];
    ivm_world_xyz(d, xs_u, ys_v, T, w[0], w[1], w[2]);
    if (!(w[1] > ivm_sub(h, 1.0f) && w[1] < ivm_add(h, 0.5f))) return 0;
    const float rf = ivm_rint_div(w[2], half_res, inv);
    const float cf = ivm_rint_div(w[0], half_res, inv);
    if (!(fabsf(rf) < 1.0e9f && fabsf(cf) < 1.0e9f)) return 2;
    p.x = w[0]; p.y = w[1]; p.z = w[2];
    p.r = (int32_t)rf; p.c = (int32_t)cf;
    return 1;
}

// frame de-dup candidate: highest point wins, lowest pixel index on ties
// (first-index rule of scatter_max; pixels are listed in (v,u) order, mapper.py:32-35).
IVM_HD unsigned long long ivm_cand_key(const IvmParams &P, float y, uint32_t pix) {
    const int pb = P.pix_bits;
    return ((unsigned long long)P.cstamp << (32 + pb)) | ((unsigned long long)ivm_orderable(y) << pb) |
           (unsigned long long)(((1u << pb) - 1u) - pix);
}
// bits of a pixel index and the longest stamp period they leave in a 64-bit candidate word
IVM_HD int ivm_pix_bits(long long hw) { int pb = 1; while (pb < 24 && (1ll << pb) < hw) ++pb; return pb; }
IVM_HD uint32_t ivm_stamp_period(int pix_bits, uint32_t override_period) {
    const uint32_t most = (uint32_t)((1ull << (32 - pix_bits)) - 1ull);
    return (override_period >= 1u && override_period < most) ? override_period : most;
}

IVM_HD bool ivm_store_index(const IvmParams &P, int32_t origin_r, int32_t origin_c, int b, int32_t r, int32_t c,
                            size_t &idx) {
    const int32_t rr = r - origin_r, cc = c - origin_c;
    if (rr < 0 || rr >= P.SR || cc < 0 || cc >= P.SC) return false;
    idx = ((size_t)b * P.SR + (size_t)rr) * P.SC + (size_t)cc;
    return true;
}

// the reference's flattened de-dup key (mapper.py:468-469) for a bbox with
// minima (rmin,cmin) and extents Rx = rows.max(), Cx = cols.max()
IVM_HD unsigned long long ivm_list_key(int b, int32_t r, int32_t c, int32_t rmin, int32_t cmin, long long Rx, long long Cx) {
    return (unsigned long long)((long long)b * (Rx * Cx) + (long long)(r - rmin) * Cx + (long long)(c - cmin));
}

// Per-thread accumulator of newly occupied cells (flushed to the env's count / bbox by the caller:
// warp + block reduction on the device, so that a step issues a handful of same-address atomics
// per CTA instead of five per new cell).
struct IvmBoxAcc {
    int32_t rmin, rmax, cmin, cmax, n;
    IVM_HD void clear() { rmin = INT32_MAX; rmax = INT32_MIN; cmin = INT32_MAX; cmax = INT32_MIN; n = 0; }
    IVM_HD void add(int32_t r, int32_t c) {
        rmin = r < rmin ? r : rmin; rmax = r > rmax ? r : rmax;
        cmin = c < cmin ? c : cmin; cmax = c > cmax ? c : cmax; ++n;
    }
};
template <class A>
IVM_HD void ivm_box_flush(IvmEnv *e, const IvmBoxAcc &a) {
    if (a.n <= 0) return;
    A::add_i(&e->count, a.n);
    A::min_i(&e->rmin, a.rmin); A::max_i(&e->rmax, a.rmax);
    A::min_i(&e->cmin, a.cmin); A::max_i(&e->cmax, a.cmax);
}

// Insert the frame/edge survivor into the world store ("world.concatenate(local)" +
// second keep_highest for a cell that does not collide): the older record wins ties
// because world points precede local points in the list (mapper.py:226-230, 299-308).
// `old` = the cell's current record, `reset_stamp` / origin = the env's (callers that already hold
// them pass them in, so that the hot loops issue no dependent loads of the env struct).
template <class A>
IVM_HD void ivm_merge_record(const IvmParams &P, int b, size_t idx, int32_t r, int32_t c, float x, float y, float z,
                             uint32_t label, const IvmRecord &old, uint32_t reset_stamp, int32_t origin_r, int32_t origin_c,
                             IvmBoxAcc &acc) {
    const bool live = ivm_live(old.meta, reset_stamp);
    if (live && !(y > old.y)) return;
    IvmRecord rec;
    rec.x = x; rec.y = y; rec.z = z; rec.meta = (P.step << 8) | (label & 0xFFu);
    ivm_store_record(&P.store[idx], rec);
    if (!live) {
        A::add_i(&P.rowcount[(size_t)b * P.SR + (r - origin_r)], 1);
        A::add_i(&P.colcount[(size_t)b * P.SC + (c - origin_c)], 1);
        acc.add(r, c);
    }
}
template <class A>
IVM_HD void ivm_merge_into_world(const IvmParams &P, int b, size_t idx, int32_t r, int32_t c, float x, float y, float z,
                                 uint32_t label, IvmBoxAcc &acc) {
    const IvmEnv *e = &P.env[b];
    const IvmRecord old = ivm_load_record(&P.store[idx]);
    ivm_merge_record<A>(P, b, idx, r, c, x, y, z, label, old, e->reset_stamp, e->origin_r, e->origin_c, acc);
}

// ---------------------------------------------------------------------------
// Per-env preparation, folded into the head of the ingest-scatter kernel.  Envs whose mask is 0
// (episode/tour finished, mapper.py:320-326), whose index is >= B (paused, mapper.py:315-318) or
// that hold nothing are wiped in O(1) by advancing reset_stamp, and the store window is
// re-centred on the pose.  Every CTA of an env DECIDES locally (same inputs, same answer);
// only the env's first CTA PUBLISHES the new state, which nobody reads before the next kernel.
struct IvmEnvPrep {
    int32_t reset;
    int32_t origin_r, origin_c;
};
// the decision from already loaded values (callers that batch the loads)
IVM_HD IvmEnvPrep ivm_env_decide_vals(const IvmParams &P, uint32_t mask, int32_t count, int32_t e_origin_r, int32_t e_origin_c,
                                      float pose_x, float pose_z) {
    IvmEnvPrep q;
    q.reset = (mask == 0u) || (count <= 0);
    if (q.reset) {
        const float pr = rintf(ivm_div(pose_z, P.half_res));
        const float pc = rintf(ivm_div(pose_x, P.half_res));
        q.origin_r = (fabsf(pr) < 1.0e9f ? (int32_t)pr : 0) - P.SR / 2;
        q.origin_c = (fabsf(pc) < 1.0e9f ? (int32_t)pc : 0) - P.SC / 2;
    } else {
        q.origin_r = e_origin_r; q.origin_c = e_origin_c;
    }
    return q;
}
IVM_HD IvmEnvPrep ivm_env_decide(const IvmParams &P, int b) {  // b < P.B
    const IvmEnv *e = &P.env[b];
    return ivm_env_decide_vals(P, P.masks[b], e->count, e->origin_r, e->origin_c, P.pose[3 * b + 0], P.pose[3 * b + 2]);
}
// called by all `nthreads` threads of the publishing CTA; for b >= P.B the env is simply wiped
template <class A>
IVM_HD void ivm_env_publish(const IvmParams &P, int b, const IvmEnvPrep &q, int tid, int nthreads) {
    const bool dropped = b >= P.B;
    if (!dropped && !q.reset) return;
    for (int i = tid; i < P.SR; i += nthreads) P.rowcount[(size_t)b * P.SR + i] = 0;
    for (int i = tid; i < P.SC; i += nthreads) P.colcount[(size_t)b * P.SC + i] = 0;
    if (tid == 0) {
        IvmEnv *e = &P.env[b];
        e->reset_stamp = P.step;
        e->count = 0;
        e->rmin = INT32_MAX; e->rmax = INT32_MIN; e->cmin = INT32_MAX; e->cmax = INT32_MIN;
        e->dirty = 0;
        if (!dropped) { e->origin_r = q.origin_r; e->origin_c = q.origin_c; }
    }
}

// Pose matrices of one env from (elevation, heading): rows 0..2 of Rx(elevation+pi)*Ry(heading)|pose
// (projector/core.py:6-37 as called by mapper.py:132-138) and (cos,sin)(-heading) (mapper.py:38-48,
// 264-266).  Trig and products are evaluated in the angles' dtype and rounded to f32 when stored,
// exactly like the reference's assignments into float32 tensors; on the GPU these are the same
// libdevice sin/cos that torch's CUDA kernels call.
IVM_HD double ivm_cos(double a) { return cos(a); }
IVM_HD double ivm_sin(double a) { return sin(a); }
IVM_HD float ivm_cos(float a) { return cosf(a); }
IVM_HD float ivm_sin(float a) { return sinf(a); }
// both at once: on the device sincos() shares the argument reduction of sin() and cos() and evaluates the same
// two polynomials, so the values are the ones sin() / cos() return (checked bitwise on the GPU by
// tests/test_gpu_parity.py::test_device_trig_f64_matches and the kernel-trig parity tests)
#if defined(__CUDA_ARCH__)
IVM_HD void ivm_sincos(double a, double &sn, double &cs) { sincos(a, &sn, &cs); }
IVM_HD void ivm_sincos(float a, float &sn, float &cs) { sincosf(a, &sn, &cs); }
#else
IVM_HD void ivm_sincos(double a, double &sn, double &cs) { sn = sin(a); cs = cos(a); }
IVM_HD void ivm_sincos(float a, float &sn, float &cs) { sn = sinf(a); cs = cosf(a); }
#endif
template <class F>
IVM_HD void ivm_pose_matrices_t(const float *pose, F elevation, F heading, float *T, float *cs) {
    const F ex = elevation + (F)3.141592653589793238462643383279502884;
    F cx, sx, cy, sy;
    ivm_sincos(ex, sx, cx);
    ivm_sincos(heading, sy, cy);
    T[0] = (float)cy; T[1] = (float)(sx * sy); T[2] = (float)(cx * sy); T[3] = pose[0];
    T[4] = 0.0f;      T[5] = (float)cx;        T[6] = (float)(-sx);     T[7] = pose[1];
    T[8] = (float)(-sy); T[9] = (float)(cy * sx); T[10] = (float)(cy * cx); T[11] = pose[2];
    // cos(-a) = cos(a) and sin(-a) = -sin(a) hold exactly for libm / libdevice (odd/even symmetric
    // implementations), so the reference's cos(-heading), sin(-heading) need no second evaluation
    cs[0] = (float)cy; cs[1] = (float)(-sy);
}
IVM_HD_COLD void ivm_pose_matrices(const IvmParams &P, int b, float *T, float *cs) {
    if (P.orient_f64) {
        const double *o = (const double *)P.orient;
        ivm_pose_matrices_t<double>(P.pose + 3 * b, o[2 * b], o[2 * b + 1], T, cs);
    } else {
        const float *o = (const float *)P.orient;
        ivm_pose_matrices_t<float>(P.pose + 3 * b, o[2 * b], o[2 * b + 1], T, cs);
    }
}

// Per-step scratch of IvmGlobal: reset once at context start and then by the fix-up program at
// the end of every step (so that no kernel has to run before the ingest of the next step).
IVM_HD void ivm_reset_step_globals(IvmGlobal *g, bool reset_rastered = true) {
    g->loc[0] = INT32_MAX; g->loc[1] = INT32_MIN; g->loc[2] = INT32_MAX; g->loc[3] = INT32_MIN;
    g->n_e1 = 0; g->n_e2 = 0; g->n_seg = 0; g->any_dirty = 0;
    g->acc_valid = 0; g->acc_local = 0; g->acc_e1 = 0; g->acc_e2 = 0;
    // (the persistent step kernel rasters BESIDE the fix-up: it zeroes this one at its start instead)
    if (reset_rastered) g->stats[IVM_STAT_IN] = 0;
}

// ---------------------------------------------------------------------------
// K1b per-pixel resolve.  `p` is the pixel's point (already unprojected and
// valid).  The pixel whose key sits in the candidate slot of its cell is the frame
// de-dup winner of that cell.  Winners on the frame bbox edge go to the edge
// list (they may collide with other cells); all others are
// merged into the world store directly.  Returns 1 if the pixel was a
// non-edge winner (for the LOCAL statistic).
// Offer a point to its cell's word of env b's candidate plane.  `cell` = index of the half-cell
// inside the env's store window.
template <class A>
IVM_HD void ivm_cand_insert(const IvmParams &P, int b, uint32_t cell, unsigned long long key) {
    A::max_ull(&P.cplane[(size_t)b * P.SR * P.SC + cell], key);
}
// The winning key of a cell after all inserts of the step (a stale or zero word if nothing was offered).
IVM_HD unsigned long long ivm_cand_lookup(const IvmParams &P, int b, uint32_t cell) {
    return ivm_load_ull(&P.cplane[(size_t)b * P.SR * P.SC + cell]);
}
IVM_HD bool ivm_on_frame_edge(const IvmPoint &p, const int32_t *loc) {
    return p.r == loc[0] || p.r == loc[1] || p.c == loc[2] || p.c == loc[3];
}
// a frame winner on the frame bbox edge waits in the edge list for the fix-up
template <class A>
IVM_HD void ivm_push_edge1(const IvmParams &P, int b, uint32_t pix, const IvmPoint &p, uint32_t label, size_t idx) {
    const uint32_t k = A::add_u(&P.g->n_e1, 1u);
    if (k >= P.ecap) { A::or_u(&P.g->err, IVM_ERR_EDGE_OVERFLOW); return; }
    IvmEdge ed;
    ed.x = p.x; ed.y = p.y; ed.z = p.z; ed.label = label;
    ed.b = b; ed.r = p.r; ed.c = p.c; ed.slot = 0;
    ed.xorder = (unsigned long long)b * (unsigned long long)P.HW + pix;  // position in the frame point list
    ed.addr = idx;
    P.e1[k] = ed;
}
template <class A>
IVM_HD int ivm_resolve_pixel(const IvmParams &P, int b, uint32_t pix, const IvmPoint &p, uint32_t label,
                             const int32_t *loc, int32_t origin_r, int32_t origin_c, IvmBoxAcc &acc) {
    size_t idx;
    if (!ivm_store_index(P, origin_r, origin_c, b, p.r, p.c, idx)) return 0;  // overflow was flagged by the scatter
    const uint32_t cell = (uint32_t)(idx - (size_t)b * P.SR * P.SC);
    if (ivm_cand_lookup(P, b, cell) != ivm_cand_key(P, p.y, pix)) return 0;
    if (ivm_on_frame_edge(p, loc)) {
        ivm_push_edge1<A>(P, b, pix, p, label, idx);
        return 0;
    }
    ivm_merge_into_world<A>(P, b, idx, p.r, p.c, p.x, p.y, p.z, label, acc);
    return 1;
}

// the same with the direct resolution of frame-edge collisions (ivm_frame_edge_loses, defined below): an edge winner
// merges at once or is dropped; nothing is deferred.  Only valid for a direct frame box (ivm_box_direct(loc)).
IVM_HD bool ivm_frame_edge_loses(const IvmParams &P, int b, int32_t r, int32_t c, uint32_t ord, unsigned long long xorder,
                                 const int32_t *loc);
template <class A>
IVM_HD int ivm_resolve_pixel_direct(const IvmParams &P, int b, uint32_t pix, const IvmPoint &p, uint32_t label,
                                    const int32_t *loc, int32_t origin_r, int32_t origin_c, IvmBoxAcc &acc) {
    size_t idx;
    if (!ivm_store_index(P, origin_r, origin_c, b, p.r, p.c, idx)) return 0;
    const uint32_t cell = (uint32_t)(idx - (size_t)b * P.SR * P.SC);
    if (ivm_cand_lookup(P, b, cell) != ivm_cand_key(P, p.y, pix)) return 0;
    if (ivm_on_frame_edge(p, loc)) {
        A::add_ull(&P.g->acc_e1, 1ull);
        if (ivm_frame_edge_loses(P, b, p.r, p.c, ivm_orderable(p.y), (unsigned long long)b * (unsigned long long)P.HW + pix, loc)) return 0;
    }
    ivm_merge_into_world<A>(P, b, idx, p.r, p.c, p.x, p.y, p.z, label, acc);
    return 1;
}

// ---------------------------------------------------------------------------
// Collision-class resolution shared by both de-dup stages: every edge entry is
// hashed by the reference's flattened key; per class the highest point wins,
// ties go to the entry that comes first in the reference's list (xorder).
// Runs inside ONE thread block (tid / nthreads), phases separated by A::sync().
IVM_HD uint32_t ivm_mix(unsigned long long k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
    return (uint32_t)k;
}

// Block-shared scratch of the fix-up program (shared memory on the device, plain arrays in the
// emulator).  key/ord/xo serve the small-class fast path.
struct IvmFixScratch {
    unsigned long long *key, *xo;  // [cap]
    uint32_t *ord;                 // [cap]
    uint32_t cap;
    int32_t *ibuf;                 // [8]: 0..3 world bbox, 4 segment count, 5 any-dirty, 6 longest segment
    unsigned long long *lbuf;      // [2]: 0 live-record total
    uint32_t *release;             // persistent step kernel: counter that releases the deferred raster tiles as soon as
    uint32_t release_add;          // the store is final (before the bookkeeping that follows); NULL elsewhere
};

// many entries: open-addressing hash in global memory
template <class A>
IVM_HD_COLD void ivm_resolve_classes_hash(const IvmParams &P, IvmEdge *E, uint32_t n, int32_t rmin, int32_t cmin,
                                          long long Rx, long long Cx, int tid, int nthreads) {
    for (uint32_t i = tid; i < n; i += nthreads) {
        const unsigned long long k = ivm_list_key(E[i].b, E[i].r, E[i].c, rmin, cmin, Rx, Cx);
        uint32_t s = ivm_mix(k) & P.hmask;
        for (;;) {
            const unsigned long long prev = A::cas_ull(&P.hkeys[s], IVM_EMPTY_KEY, k);
            if (prev == IVM_EMPTY_KEY || prev == k) break;
            s = (s + 1) & P.hmask;
        }
        E[i].slot = s;
        A::max_u(&P.hbest[s], ivm_orderable(E[i].y));
    }
    A::sync();
    for (uint32_t i = tid; i < n; i += nthreads)
        if (ivm_orderable(E[i].y) == P.hbest[E[i].slot]) A::min_ull(&P.hxord[E[i].slot], E[i].xorder);
    A::sync();
    for (uint32_t i = tid; i < n; i += nthreads) {
        const uint32_t s = E[i].slot;
        const bool win = ivm_orderable(E[i].y) == P.hbest[s] && E[i].xorder == P.hxord[s];
        E[i].addr = win ? E[i].addr : (E[i].addr | 0x8000000000000000ull);  // tag losers, keep slot for cleanup
    }
    A::sync();
    for (uint32_t i = tid; i < n; i += nthreads) {
        const uint32_t s = E[i].slot;
        P.hkeys[s] = IVM_EMPTY_KEY; P.hbest[s] = 0u; P.hxord[s] = IVM_EMPTY_KEY;
        const bool lose = (E[i].addr >> 63) != 0ull;
        E[i].addr &= 0x7FFFFFFFFFFFFFFFull;
        E[i].slot = lose ? 0xFFFFFFFFu : 0u;
    }
    A::sync();
}

// on return E[i].slot = 0xFFFFFFFF for losers, anything else for winners.  One out-of-line copy
// serves both de-dup stages (the second call finds the code already fetched).
template <class A>
IVM_HD_COLD void ivm_resolve_classes(const IvmParams &P, IvmEdge *E, uint32_t n, int32_t rmin, int32_t cmin, long long Rx,
                                     long long Cx, const IvmFixScratch &S, int tid, int nthreads) {
    if (n > S.cap) {
        ivm_resolve_classes_hash<A>(P, E, n, rmin, cmin, Rx, Cx, tid, nthreads);
        return;
    }
    // few entries (the usual case: a bbox edge holds a handful of points): all-pairs in block memory
    for (uint32_t i = tid; i < n; i += nthreads) {
        S.key[i] = ivm_list_key(E[i].b, E[i].r, E[i].c, rmin, cmin, Rx, Cx);
        S.ord[i] = ivm_orderable(E[i].y);
        S.xo[i] = E[i].xorder;
    }
    A::sync();
    for (uint32_t i = tid; i < n; i += nthreads) {
        const unsigned long long ki = S.key[i], xi = S.xo[i];
        const uint32_t oi = S.ord[i];
        bool lose = false;
        for (uint32_t j = 0; j < n; ++j)
            if (S.key[j] == ki && (S.ord[j] > oi || (S.ord[j] == oi && S.xo[j] < xi))) { lose = true; break; }
        E[i].slot = lose ? 0xFFFFFFFFu : 0u;
    }
    A::sync();
}

// scan one store cell of an edge line; live records join the stage-2 edge list
template <class A>
IVM_HD_COLD void ivm_scan_edge_cell(const IvmParams &P, int b, int32_t r, int32_t c, const int32_t *loc) {
    const IvmEnv &e = P.env[b];
    size_t idx;
    if (!ivm_store_index(P, e.origin_r, e.origin_c, b, r, c, idx)) return;
    const IvmRecord rec = ivm_load_record(&P.store[idx]);
    if (!ivm_live(rec.meta, e.reset_stamp)) return;
    const uint32_t k = A::add_u(&P.g->n_e2, 1u);
    if (k >= P.ecap) { A::or_u(&P.g->err, IVM_ERR_EDGE_OVERFLOW); return; }
    const IvmGlobal *g = P.g;
    const bool fresh = (rec.meta >> 8) == P.step;
    IvmEdge ed;
    ed.x = rec.x; ed.y = rec.y; ed.z = rec.z; ed.label = rec.meta & 0xFFu;
    ed.b = b; ed.r = r; ed.c = c; ed.slot = 0; ed.addr = idx;
    // position in the concatenated list [world(t-1) ; frame survivors]: old records keep
    // the order of the previous sort (previous stage-2 key), fresh ones follow in
    // frame-key order (mapper.py:226-230, 471-474).
    if (fresh)
        ed.xorder = (1ull << 62) | ivm_list_key(b, r, c, loc[0], loc[2], (long long)loc[1] - loc[0], (long long)loc[3] - loc[2]);
    else
        ed.xorder = ivm_list_key(b, r, c, g->prev_rmin, g->prev_cmin, g->prev_R, g->prev_C);
    P.e2[k] = ed;
}

// exact bbox of the envs that lost records, from the per-row / per-column live counts.  Deletions only shrink a
// box, so each side walks INWARD from its old extreme to the first row / column that still holds a record (usually
// the extreme itself): one thread per (dirty env, side).  The fields are stored one by one: in the persistent step
// kernel other CTAs raster ego tiles meanwhile and read rmin..cmax; they must see the old or the new value of a field
// (both bound the live records), never an intermediate one.
template <class A>
IVM_HD_COLD void ivm_rebuild_dirty_boxes(const IvmParams &P, int tid, int nthreads) {
    for (int i = tid; i < 4 * P.B; i += nthreads) {
        const int b = i >> 2, side = i & 3;
        IvmEnv *e = &P.env[b];
        if (!e->dirty) continue;
        const int32_t *cnt = (side < 2) ? P.rowcount + (size_t)b * P.SR : P.colcount + (size_t)b * P.SC;
        const int32_t org = (side < 2) ? e->origin_r : e->origin_c;
        const int32_t lo = (side < 2) ? e->rmin : e->cmin, hi = (side < 2) ? e->rmax : e->cmax;
        int32_t v;
        if ((side & 1) == 0) { v = lo; while (v <= hi && ivm_load_u32((const uint32_t *)&cnt[v - org]) == 0u) ++v; if (v > hi) v = INT32_MAX; }
        else { v = hi; while (v >= lo && ivm_load_u32((const uint32_t *)&cnt[v - org]) == 0u) --v; if (v < lo) v = INT32_MIN; }
        e->nb[side] = v;
    }
    A::sync();
    for (int b = tid; b < P.B; b += nthreads) {
        IvmEnv *e = &P.env[b];
        if (!e->dirty) continue;
        e->rmin = e->nb[0]; e->rmax = e->nb[1]; e->cmin = e->nb[2]; e->cmax = e->nb[3];
        e->dirty = 0;
    }
}

// ---------------------------------------------------------------------------
// Direct resolution of the edge collisions (the fast path of both de-dup stages).
// With the reference's key  b*(R'C') + r'*C' + c'  (strides max, not max+1) and a box of at least 3 rows and
// 2 columns (R' >= 2, C' >= 1), a key class has at most three member cells, and they can be written down:
// for env b2 in {b-1, b, b+1} the value M = m - (b2-b)*R'C' (m = r'C' + c') decomposes as r2*C' + c2 in at most two
// ways -- (M / C', M % C') and, when the remainder is 0, (M / C' - 1, C').  So instead of grouping edge entries
// by key (lists, hashing, one thread block), every edge cell LOOKS UP its few partner cells and decides on its
// own whether it survives: it loses iff a partner holds a point that is higher, or as high and earlier in the
// reference's list.  That is a strict total order, so exactly the class winner survives, as in the reference.
struct IvmPartner { int32_t b, r, c; };   // env, ABSOLUTE half-row / half-col
#define IVM_MAX_PARTNERS 5
IVM_HD bool ivm_box_direct(const int32_t *box) {  // box = rmin, rmax, cmin, cmax (a valid box)
    return (long long)box[1] - box[0] >= 2 && (long long)box[3] - box[2] >= 1;
}
IVM_HD int ivm_partners(int b, int32_t r, int32_t c, const int32_t *box, int B, IvmPartner *out) {
    const long long Rx = (long long)box[1] - box[0], Cx = (long long)box[3] - box[2], S = Rx * Cx;
    const long long rr = (long long)r - box[0], cc = (long long)c - box[2], m = rr * Cx + cc;
    int n = 0;
    for (int db = -1; db <= 1; ++db) {
        const int b2 = b + db;
        if (b2 < 0 || b2 >= B) continue;
        const long long M = m - (long long)db * S;
        if (M < 0 || M > S + Cx) continue;
        // (the quotient of two 64-bit integers is slow on the device: the values fit 32 bits unless envs lie far apart)
        const long long r2 = (M < 0x7FFFFFFFll && Cx < 0x7FFFFFFFll) ? (long long)((uint32_t)M / (uint32_t)Cx) : M / Cx;
        const long long c2 = M - r2 * Cx;
        if (r2 <= Rx && !(db == 0 && r2 == rr && c2 == cc)) {
            out[n].b = b2; out[n].r = (int32_t)(r2 + box[0]); out[n].c = (int32_t)(c2 + box[2]); ++n;
        }
        if (c2 == 0 && r2 >= 1 && r2 - 1 <= Rx && !(db == 0 && r2 - 1 == rr && Cx == cc)) {
            out[n].b = b2; out[n].r = (int32_t)(r2 - 1 + box[0]); out[n].c = (int32_t)(Cx + box[2]); ++n;
        }
    }
    return n;
}
// stage 1: does the frame winner (height order `ord`, list position `xorder` = b*HW + pixel) of an edge cell of
// the frame box `loc` lose its collision class?  The partners' winners sit in the candidate plane (every scatter
// of the step has completed).  The envs' store windows are read as published this step.
IVM_HD bool ivm_frame_edge_loses(const IvmParams &P, int b, int32_t r, int32_t c, uint32_t ord, unsigned long long xorder,
                                 const int32_t *loc) {
    IvmPartner pt[IVM_MAX_PARTNERS];
    const int n = ivm_partners(b, r, c, loc, P.B, pt);
    const int pb = P.pix_bits;
    bool lose = false;
    for (int i = 0; i < n; ++i) {
        const IvmEnv *e = &P.env[pt[i].b];
        size_t idx;
        if (!ivm_store_index(P, (int32_t)ivm_load_u32((const uint32_t *)&e->origin_r), (int32_t)ivm_load_u32((const uint32_t *)&e->origin_c),
                             pt[i].b, pt[i].r, pt[i].c, idx)) continue;
        const unsigned long long w = ivm_load_ull(&P.cplane[idx]);
        if ((uint32_t)(w >> (32 + pb)) != P.cstamp) continue;  // nothing offered to that cell this step
        const uint32_t po = (uint32_t)(w >> pb);
        const unsigned long long px = (unsigned long long)pt[i].b * (unsigned long long)P.HW +
                                      (unsigned long long)(((1u << pb) - 1u) - (uint32_t)(w & ((1ull << pb) - 1ull)));
        lose = lose || po > ord || (po == ord && px < xorder);
    }
    return lose;
}
// position of a live world record in the concatenated list [world(t-1) ; frame survivors] (see ivm_scan_edge_cell)
IVM_HD unsigned long long ivm_world_xorder(const IvmParams &P, int b, int32_t r, int32_t c, uint32_t meta, const int32_t *loc) {
    const IvmGlobal *g = P.g;
    if ((meta >> 8) == P.step)
        return (1ull << 62) | ivm_list_key(b, r, c, loc[0], loc[2], (long long)loc[1] - loc[0], (long long)loc[3] - loc[2]);
    return ivm_list_key(b, r, c, g->prev_rmin, g->prev_cmin, g->prev_R, g->prev_C);
}
// stage 2: does the live record `rec` of edge cell (b, r, c) of the world box `glob` lose its collision class?
IVM_HD bool ivm_world_edge_loses(const IvmParams &P, int b, int32_t r, int32_t c, const IvmRecord &rec, const int32_t *glob,
                                 const int32_t *loc) {
    IvmPartner pt[IVM_MAX_PARTNERS];
    const int n = ivm_partners(b, r, c, glob, P.B, pt);
    if (n == 0) return false;
    const uint32_t ord = ivm_orderable(rec.y);
    const unsigned long long xo = ivm_world_xorder(P, b, r, c, rec.meta, loc);
    bool lose = false;
    for (int i = 0; i < n; ++i) {
        const IvmEnv *e = &P.env[pt[i].b];
        size_t idx;
        if (!ivm_store_index(P, e->origin_r, e->origin_c, pt[i].b, pt[i].r, pt[i].c, idx)) continue;
        const IvmRecord q = ivm_load_record(&P.store[idx]);
        if (!ivm_live(q.meta, e->reset_stamp)) continue;
        const uint32_t po = ivm_orderable(q.y);
        lose = lose || po > ord || (po == ord && ivm_world_xorder(P, pt[i].b, pt[i].r, pt[i].c, q.meta, loc) < xo);
    }
    return lose;
}
// a loser of stage 2 joins the deletion list (the e2 list, every entry marked as a loser)
template <class A>
IVM_HD void ivm_push_loser(const IvmParams &P, int b, int32_t r, int32_t c, size_t idx) {
    const uint32_t k = A::add_u(&P.g->n_e2, 1u);
    if (k >= P.ecap) { A::or_u(&P.g->err, IVM_ERR_EDGE_OVERFLOW); return; }
    IvmEdge ed;
    ed.x = ed.y = ed.z = 0.f; ed.label = 0u; ed.b = b; ed.r = r; ed.c = c; ed.slot = 0xFFFFFFFFu; ed.xorder = 0ull;
    ed.addr = idx;
    P.e2[k] = ed;
}
// one store cell of an edge line of the world box, direct path: a live record that loses is listed for deletion;
// returns 1 if the cell holds a live record (statistics)
template <class A>
IVM_HD_COLD int ivm_direct_edge_cell(const IvmParams &P, int b, int32_t r, int32_t c, int32_t origin_r, int32_t origin_c,
                                uint32_t reset_stamp, const int32_t *glob, const int32_t *loc) {
    size_t idx;
    if (!ivm_store_index(P, origin_r, origin_c, b, r, c, idx)) return 0;
    const IvmRecord rec = ivm_load_record(&P.store[idx]);
    if (!ivm_live(rec.meta, reset_stamp)) return 0;
    if (ivm_world_edge_loses(P, b, r, c, rec, glob, loc)) ivm_push_loser<A>(P, b, r, c, idx);
    return 1;
}

// F: the edge fix-up of both de-dup stages + bbox bookkeeping, in three parts so that the fused
// step kernel can spread the middle one (the edge-line scan) over the whole grid:
//   ivm_fixup_stage1  one thread block: stage-1 classes + merges, world bbox, edge-line segments
//   ivm_fixup_scan    any number of blocks: live records on the edge lines -> stage-2 edge list
//   ivm_fixup_stage2  one thread block: stage-2 classes, deletions, bookkeeping, publish
// State handed from part to part lives in IvmGlobal (glob[], n_seg, scan_chunks) and P.segs.
#define IVM_SCAN_CHUNK 256   // cells of one segment handled by one block pass

template <class A>
IVM_HD void ivm_fixup_stage1(const IvmParams &P, const IvmFixScratch &S, int tid, int nthreads) {
    IvmGlobal *g = P.g;
    const uint32_t n1 = g->n_e1 < P.ecap ? g->n_e1 : P.ecap;
    const int32_t loc[4] = {g->loc[0], g->loc[1], g->loc[2], g->loc[3]};
    if (tid == 0) {
        S.ibuf[0] = INT32_MAX; S.ibuf[1] = INT32_MIN; S.ibuf[2] = INT32_MAX; S.ibuf[3] = INT32_MIN;
        S.ibuf[4] = 0; S.ibuf[5] = 0; S.ibuf[6] = 0;
    }
    IVM_TRACE(g, 0, tid);
    // ---- stage 1: collisions on the frame bbox edge (mapper.py:840-842)
    if (n1 > 0) {
        ivm_resolve_classes<A>(P, P.e1, n1, loc[0], loc[2], (long long)loc[1] - loc[0], (long long)loc[3] - loc[2], S, tid,
                               nthreads);
        for (uint32_t i = tid; i < n1; i += nthreads) {
            const IvmEdge ed = P.e1[i];
            if (ed.slot == 0xFFFFFFFFu) continue;
            A::add_ull(&g->acc_local, 1ull);
            IvmBoxAcc acc;
            acc.clear();
            ivm_merge_into_world<A>(P, ed.b, (size_t)ed.addr, ed.r, ed.c, ed.x, ed.y, ed.z, ed.label, acc);
            ivm_box_flush<A>(&P.env[ed.b], acc);
        }
    }
    A::sync();
    IVM_TRACE(g, 1, tid);
    // ---- stage-2 bbox over all live records of all envs (mapper.py:461-469 on world + frame survivors)
    for (int b = tid; b < P.B; b += nthreads) {
        const IvmEnv &e = P.env[b];
        if (e.count > 0) {
            A::min_i(&S.ibuf[0], e.rmin); A::max_i(&S.ibuf[1], e.rmax);
            A::min_i(&S.ibuf[2], e.cmin); A::max_i(&S.ibuf[3], e.cmax);
        }
    }
    A::sync();
    const int32_t grmin = S.ibuf[0], grmax = S.ibuf[1], gcmin = S.ibuf[2], gcmax = S.ibuf[3];
    const bool alive = grmin <= grmax;  // else nothing alive: the reference skips keep_highest on an empty cloud
    if (alive) {
        // ---- which edge lines hold records?  an env has cells on a global edge line only if its own
        //      bbox touches that line.  segs[q] = (env, is_col, line, cells to scan)
        for (int b = tid; b < P.B; b += nthreads) {
            const IvmEnv &e = P.env[b];
            if (e.count <= 0) continue;
            const int lines[4] = {grmin, grmax, gcmin, gcmax};
            const bool touch[4] = {e.rmin == grmin, e.rmax == grmax && grmax != grmin, e.cmin == gcmin,
                                   e.cmax == gcmax && gcmax != gcmin};
            for (int s = 0; s < 4; ++s)
                if (touch[s]) {
                    const int k = (int)A::add_u((uint32_t *)&S.ibuf[4], 1u);
                    const int len = (s >> 1) ? e.rmax - e.rmin + 1 : e.cmax - e.cmin + 1;
                    P.segs[4 * k + 0] = b; P.segs[4 * k + 1] = s >> 1; P.segs[4 * k + 2] = lines[s]; P.segs[4 * k + 3] = len;
                    A::max_i(&S.ibuf[6], len);
                }
        }
    }
    A::sync();
    if (tid == 0) {
        g->glob[0] = grmin; g->glob[1] = grmax; g->glob[2] = gcmin; g->glob[3] = gcmax;
        g->n_seg = alive ? (uint32_t)S.ibuf[4] : 0u;
        g->scan_chunks = alive ? (uint32_t)((S.ibuf[6] + IVM_SCAN_CHUNK - 1) / IVM_SCAN_CHUNK) : 0u;
    }
    IVM_TRACE(g, 2, tid);
}

// Block `blk` of `nblk` visits the (segment, chunk) units blk, blk + nblk, ...: one cell per thread
// and unit, so that on the device the whole scan is a single round of independent loads.
template <class A>
IVM_HD void ivm_fixup_scan(const IvmParams &P, int blk, int nblk, int tid, int nthreads) {
    const IvmGlobal *g = P.g;
    const int nseg = (int)g->n_seg, nchunks = (int)g->scan_chunks;
    if (nseg <= 0 || nchunks <= 0) return;
    const int32_t grmin = g->glob[0], grmax = g->glob[1];
    const int32_t loc[4] = {g->loc[0], g->loc[1], g->loc[2], g->loc[3]};
    const long long units = (long long)nseg * nchunks;
    for (long long u = blk; u < units; u += nblk) {
        const int q = (int)(u % nseg), j = (int)(u / nseg);
        const int b = P.segs[4 * q + 0], is_col = P.segs[4 * q + 1], line = P.segs[4 * q + 2], len = P.segs[4 * q + 3];
        const IvmEnv &e = P.env[b];
        const int32_t lo = is_col ? e.rmin : e.cmin;
        for (int t = tid; t < IVM_SCAN_CHUNK; t += nthreads) {
            const int off = j * IVM_SCAN_CHUNK + t;
            if (off >= len) break;
            const int32_t v = lo + off;
            size_t idx;
            if (!ivm_store_index(P, e.origin_r, e.origin_c, b, is_col ? v : line, is_col ? line : v, idx)) continue;
            if (!ivm_live(ivm_load_meta(&P.store[idx]), e.reset_stamp)) continue;
            if (is_col) {
                if (v != grmin && v != grmax) ivm_scan_edge_cell<A>(P, b, v, line, loc);  // corners belong to the row scans
            } else {
                ivm_scan_edge_cell<A>(P, b, line, v, loc);
            }
        }
    }
}

#if defined(__CUDACC__)
// Pull LAST step's edge lines towards L2 (the world bbox rarely moves between steps, so they are almost always
// this step's lines as well): run by the scan team while stage 1 of the fix-up is still busy.  Hints only.
__device__ __forceinline__ void ivm_fixup_scan_prefetch(const IvmParams &P, int blk, int nblk, int tid, int nthreads) {
    const IvmGlobal *g = P.g;
    const int nseg = (int)g->prev_n_seg, nchunks = (int)g->prev_scan_chunks;
    if (nseg <= 0 || nchunks <= 0 || nseg > 4 * P.maxB || (long long)nchunks * IVM_SCAN_CHUNK * nseg > (1ll << 24)) return;
    const int span = nchunks * IVM_SCAN_CHUNK, total = span * nseg;
    for (int i = blk * nthreads + tid; i < total; i += nblk * nthreads) {
        const int q = i / span, off = i - q * span;
        const int b = P.segs[4 * q + 0], is_col = P.segs[4 * q + 1], line = P.segs[4 * q + 2], len = P.segs[4 * q + 3];
        if (b < 0 || b >= P.B || off >= len) continue;
        if (!is_col && (off & 1)) continue;  // a row segment is contiguous: one prefetch per 32-byte sector
        const IvmEnv &e = P.env[b];
        const int32_t v = (is_col ? e.rmin : e.cmin) + off;
        size_t idx;
        if (!ivm_store_index(P, e.origin_r, e.origin_c, b, is_col ? v : line, is_col ? line : v, idx)) continue;
        asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(&P.store[idx]));
    }
}
#endif

// The same scan by one thread block (or a few: block `blk` of `nblk`).  The segment headers (and what is needed of their envs) are staged in
// block memory first; then every thread loads the metas of IVM_SCAN_MLP cells together (independent loads)
// before any live cell is appended, so the whole scan costs a few memory round trips.
#define IVM_SCAN_MLP 16
#define IVM_SCAN_HDR 8   // ints per staged segment: b, is_col, line, len, lo, origin_r, origin_c, reset_stamp
template <class A>
IVM_HD void ivm_fixup_scan_block(const IvmParams &P, const IvmFixScratch &S, int blk, int nblk, int tid, int nthreads) {
    const IvmGlobal *g = P.g;
    const int nseg = (int)g->n_seg, nchunks = (int)g->scan_chunks;
    if (nseg <= 0 || nchunks <= 0) return;  // block-uniform
    if ((size_t)nseg * IVM_SCAN_HDR * sizeof(int32_t) > (size_t)S.cap * sizeof(unsigned long long) ||
        (long long)nchunks * IVM_SCAN_CHUNK * nseg > (1ll << 30)) {
        ivm_fixup_scan<A>(P, blk, nblk, tid, nthreads);  // more segments than the scratch holds
        return;
    }
    int32_t *hdr = reinterpret_cast<int32_t *>(S.key);  // free between the two class resolutions
    for (int q = tid; q < nseg; q += nthreads) {
        const int b = P.segs[4 * q + 0], is_col = P.segs[4 * q + 1];
        const IvmEnv &e = P.env[b];
        hdr[IVM_SCAN_HDR * q + 0] = b; hdr[IVM_SCAN_HDR * q + 1] = is_col;
        hdr[IVM_SCAN_HDR * q + 2] = P.segs[4 * q + 2]; hdr[IVM_SCAN_HDR * q + 3] = P.segs[4 * q + 3];
        hdr[IVM_SCAN_HDR * q + 4] = is_col ? e.rmin : e.cmin;
        hdr[IVM_SCAN_HDR * q + 5] = e.origin_r; hdr[IVM_SCAN_HDR * q + 6] = e.origin_c;
        hdr[IVM_SCAN_HDR * q + 7] = (int32_t)e.reset_stamp;
    }
    A::sync();
    const int32_t grmin = g->glob[0], grmax = g->glob[1];
    const int32_t loc[4] = {g->loc[0], g->loc[1], g->loc[2], g->loc[3]};
    const int span = nchunks * IVM_SCAN_CHUNK;  // padded cells per segment
    const int total = span * nseg;
    const int gthreads = nthreads * nblk, gtid = blk * nthreads + tid;  // `nblk` blocks share the cells
    for (int base = 0; base < total; base += gthreads * IVM_SCAN_MLP) {
        uint32_t meta[IVM_SCAN_MLP];
#pragma unroll
        for (int j = 0; j < IVM_SCAN_MLP; ++j) {
            meta[j] = 0u;
            const int i = base + j * gthreads + gtid;
            if (i >= total) continue;
            const int q = i / span, off = i - q * span;
            const int32_t *h = hdr + IVM_SCAN_HDR * q;
            if (off >= h[3]) continue;
            const int32_t v = h[4] + off;
            if (h[1] && (v == grmin || v == grmax)) continue;  // corners belong to the row scans
            size_t idx;
            if (!ivm_store_index(P, h[5], h[6], h[0], h[1] ? v : h[2], h[1] ? h[2] : v, idx)) continue;
            meta[j] = ivm_load_meta(&P.store[idx]);
        }
#pragma unroll
        for (int j = 0; j < IVM_SCAN_MLP; ++j) {
            if (meta[j] == 0u) continue;
            const int i = base + j * gthreads + gtid;
            const int q = i / span, off = i - q * span;
            const int32_t *h = hdr + IVM_SCAN_HDR * q;
            if (!ivm_live(meta[j], (uint32_t)h[7])) continue;
            const int32_t v = h[4] + off;
            ivm_scan_edge_cell<A>(P, h[0], h[1] ? v : h[2], h[1] ? h[2] : v, loc);
        }
    }
    A::sync();  // the staged headers are dead: the scratch goes back to the class resolution
}

// Direct path of the edge-line scan.  The segments (which env touches which edge line of the world box `glob`, and
// where) are derived from the env boxes by ONE warp-sized group of threads (tid < 32) in a fixed order, so that every
// block of a team builds the same table on its own: hdr[q] = b, is_col, line, len, lo, origin_r, origin_c,
// reset_stamp (IVM_SCAN_HDR ints).  Returns the segment count through counts[0] and the longest segment through
// counts[1]; counts[0] = -1 if the table would not fit `cap` segments.  `tb` = scratch of P.B bytes (block memory).
#if defined(__CUDA_ARCH__)
#define IVM_BALLOT(p) __ballot_sync(0xffffffffu, (p))
#define IVM_LANES 32
#else
#define IVM_BALLOT(p) ((p) ? 1u : 0u)
#define IVM_LANES 1
#endif
IVM_HD_COLD void ivm_direct_segments(const IvmParams &P, const int32_t *glob, int32_t *hdr, int cap, int32_t *counts, int lane,
                                     uint8_t *tb) {
    // pass 1: which edge lines does each env touch?  (its box is exact, so touching = holding a live record there)
    // bit 0 first row, bit 1 last row, bit 2 first column, bit 3 last column of the world box
    for (int b = lane; b < P.B; b += IVM_LANES) {
        const IvmEnv *e = &P.env[b];
        unsigned t = 0;
        if (e->count > 0)
            t = (e->rmin == glob[0] ? 1u : 0u) | ((e->rmax == glob[1] && glob[1] != glob[0]) ? 2u : 0u) |
                (e->cmin == glob[2] ? 4u : 0u) | ((e->cmax == glob[3] && glob[3] != glob[2]) ? 8u : 0u);
        tb[b] = (uint8_t)t;
    }
#if defined(__CUDA_ARCH__)
    __syncwarp();
#endif
    // pass 2: a line of env b has to be scanned only if a cell that can share a key with one of its cells may be live.
    // Keys are shared between (b: last column) ~ (b: first column), and between the last row / last column of env b and
    // the first row / first column of env b+1 (ivm_partners); so the "max" lines (last row, last column) of env b matter
    // iff env b touches the first column or env b+1 touches a "min" line, and the "min" lines of env b iff env b touches
    // the last column or env b-1 touches a "max" line.  In the usual case -- the four edge lines are touched by four
    // unrelated envs -- nothing has to be scanned at all.
    int nseg = 0, longest = 0;
    bool fits = true;
    const int lines[4] = {glob[0], glob[1], glob[2], glob[3]};
    for (int b0 = 0; b0 < P.B; b0 += IVM_LANES) {
        const int b = b0 + lane;
        bool touch[4] = {false, false, false, false};
        IvmEnv e;
        e.count = 0; e.rmin = e.rmax = e.cmin = e.cmax = 0; e.origin_r = e.origin_c = 0; e.reset_stamp = 0;
        if (b < P.B) {
            const unsigned t = tb[b], tn = (b + 1 < P.B) ? tb[b + 1] : 0u, tp = (b >= 1) ? tb[b - 1] : 0u;
            const bool need_max = (t & 4u) || (tn & 5u), need_min = (t & 8u) || (tp & 10u);
            touch[0] = (t & 1u) && need_min; touch[1] = (t & 2u) && need_max;
            touch[2] = (t & 4u) && need_min; touch[3] = (t & 8u) && need_max;
            if (touch[0] || touch[1] || touch[2] || touch[3]) e = P.env[b];
        }
        for (int s = 0; s < 4; ++s) {
            const unsigned m = IVM_BALLOT(touch[s]);
            if (touch[s]) {
                const int q = nseg + __builtin_popcount(m & ((1u << lane) - 1u));
                if (q < cap) {
                    int32_t *h = hdr + IVM_SCAN_HDR * q;
                    h[0] = b; h[1] = s >> 1; h[2] = lines[s];
                    h[3] = (s >> 1) ? e.rmax - e.rmin + 1 : e.cmax - e.cmin + 1;
                    h[4] = (s >> 1) ? e.rmin : e.cmin; h[5] = e.origin_r; h[6] = e.origin_c; h[7] = (int32_t)e.reset_stamp;
                } else {
                    fits = false;
                }
            }
            nseg += __builtin_popcount(m);
            int len = touch[s] ? ((s >> 1) ? e.rmax - e.rmin + 1 : e.cmax - e.cmin + 1) : 0;
#if defined(__CUDA_ARCH__)
            len = __reduce_max_sync(0xffffffffu, len);
            fits = __all_sync(0xffffffffu, fits);
#endif
            longest = len > longest ? len : longest;
        }
    }
    if (lane == 0) { counts[0] = fits ? nseg : -1; counts[1] = longest; }
}
// The cells of the staged segments, shared by `nblk` blocks: metas first (independent loads), then every live cell
// decides for itself (ivm_direct_edge_cell).  Returns the live records this thread found.
template <class A>
IVM_HD_COLD int ivm_direct_scan(const IvmParams &P, const int32_t *hdr, int nseg, int longest, const int32_t *glob, const int32_t *loc,
                           int blk, int nblk, int tid, int nthreads) {
    if (nseg <= 0 || longest <= 0) return 0;
    const int span = ((longest + IVM_SCAN_CHUNK - 1) / IVM_SCAN_CHUNK) * IVM_SCAN_CHUNK;
    const long long total = (long long)span * nseg;
    const int gthreads = nthreads * nblk, gtid = blk * nthreads + tid;
    int nlive = 0;
    for (long long base = 0; base < total; base += (long long)gthreads * IVM_SCAN_MLP) {
        uint32_t meta[IVM_SCAN_MLP];
#pragma unroll
        for (int j = 0; j < IVM_SCAN_MLP; ++j) {
            meta[j] = 0u;
            const long long i = base + (long long)j * gthreads + gtid;
            if (i >= total) continue;
            const int q = (int)(i / span), off = (int)(i - (long long)q * span);
            const int32_t *h = hdr + IVM_SCAN_HDR * q;
            if (off >= h[3]) continue;
            const int32_t v = h[4] + off;
            if (h[1] && (v == glob[0] || v == glob[1])) continue;  // corners belong to the row scans
            size_t idx;
            if (!ivm_store_index(P, h[5], h[6], h[0], h[1] ? v : h[2], h[1] ? h[2] : v, idx)) continue;
            meta[j] = ivm_load_meta(&P.store[idx]);
        }
#pragma unroll
        for (int j = 0; j < IVM_SCAN_MLP; ++j) {
            if (meta[j] == 0u) continue;
            const long long i = base + (long long)j * gthreads + gtid;
            const int q = (int)(i / span), off = (int)(i - (long long)q * span);
            const int32_t *h = hdr + IVM_SCAN_HDR * q;
            if (!ivm_live(meta[j], (uint32_t)h[7])) continue;
            const int32_t v = h[4] + off;
            nlive += ivm_direct_edge_cell<A>(P, h[0], h[1] ? v : h[2], h[1] ? h[2] : v, h[5], h[6], (uint32_t)h[7], glob, loc);
        }
    }
    return nlive;
}

// the listed losers (slot == 0xFFFFFFFF) of the e2 list leave the store: entries first, first + stride, ...
template <class A>
IVM_HD void ivm_delete_losers(const IvmParams &P, uint32_t n2, uint32_t first, uint32_t stride) {
    IvmGlobal *g = P.g;
    for (uint32_t i = first; i < n2; i += stride) {
        const IvmEdge ed = P.e2[i];
        if (ed.slot != 0xFFFFFFFFu) continue;
        IvmEnv *e = &P.env[ed.b];
        P.store[(size_t)ed.addr].meta = 0u;  // merged away for good
        A::add_i(&P.rowcount[(size_t)ed.b * P.SR + (ed.r - e->origin_r)], -1);
        A::add_i(&P.colcount[(size_t)ed.b * P.SC + (ed.c - e->origin_c)], -1);
        A::add_i(&e->count, -1);
        e->dirty = 1;
        g->any_dirty = 1u;
        A::add_ull(&g->stats[IVM_STAT_MERGED], 1ull);
    }
}

// stage 2 in two halves: `resolve` = the class resolution over the e2 list (generic path); `finish` = deletion of the
// listed losers, bbox rebuild, release of the deferred raster tiles, bookkeeping.  The direct path lists only losers
// (already marked) and calls the second half alone; e1_stat / e2_stat = the figures to publish.
template <class A>
IVM_HD void ivm_fixup_stage2(const IvmParams &P, const IvmFixScratch &S, int tid, int nthreads, bool resolve = true,
                             uint32_t e1_stat = 0xFFFFFFFFu, uint32_t e2_stat = 0xFFFFFFFFu, bool deleted = false) {
    IvmGlobal *g = P.g;
    const int32_t grmin = g->glob[0], grmax = g->glob[1], gcmin = g->glob[2], gcmax = g->glob[3];
    const bool alive = grmin <= grmax;
    const uint32_t n1 = e1_stat != 0xFFFFFFFFu ? e1_stat : (g->n_e1 < P.ecap ? g->n_e1 : P.ecap);
    uint32_t n2 = 0;
    if (tid == 0) { S.ibuf[5] = 0; S.lbuf[0] = 0ull; }
    A::sync();
    IVM_TRACE(g, 4, tid);
    if (alive) {
        // ---- stage 2: collisions on the world bbox edge (mapper.py:844-847)
        n2 = g->n_e2 < P.ecap ? g->n_e2 : P.ecap;
        if (resolve ? n2 > 1 : n2 > 0) {
            if (resolve)
                ivm_resolve_classes<A>(P, P.e2, n2, grmin, gcmin, (long long)grmax - grmin, (long long)gcmax - gcmin, S, tid,
                                       nthreads);
            if (!deleted) ivm_delete_losers<A>(P, n2, (uint32_t)tid, (uint32_t)nthreads);  // (else: the team has done it)
            A::sync();
            // ---- rebuild the bbox of envs that lost records
            if (ivm_load_u32(&g->any_dirty)) ivm_rebuild_dirty_boxes<A>(P, tid, nthreads);
        }
        IVM_TRACE(g, 5, tid);
    }
#if defined(__CUDA_ARCH__)
    // the world store and the env boxes are final from here on: let the deferred ego tiles go before the bookkeeping
    A::sync();
    if (S.release != nullptr && tid == 0) {
        __threadfence();
        asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(S.release), "r"(S.release_add) : "memory");
    }
#endif
    if (alive) {
        for (int b = tid; b < P.B; b += nthreads) {
            const int32_t c = P.env[b].count;
            if (c > 0) A::add_ull(&S.lbuf[0], (unsigned long long)c);
        }
    }
    A::sync();
    // ---- publish the step's figures, remember its list order (the next step's tie-breaks need
    //      it), and reset the per-step scratch for the next step's ingest
    if (tid == 0) {
        g->prev_valid = alive ? 1 : 0;
        if (alive) {
            g->prev_rmin = grmin; g->prev_cmin = gcmin;
            g->prev_R = (long long)grmax - grmin; g->prev_C = (long long)gcmax - gcmin;
        }
        g->stats[IVM_STAT_VALID] = g->acc_valid;
        g->stats[IVM_STAT_LOCAL] = g->acc_local;
        g->stats[IVM_STAT_WORLD] = S.lbuf[0];
        g->stats[IVM_STAT_E1] = n1;
        g->stats[IVM_STAT_E2] = e2_stat != 0xFFFFFFFFu ? e2_stat : n2;
        g->stats[7] = ((unsigned long long)g->n_seg << 32) | ((unsigned long long)g->scan_chunks * IVM_SCAN_CHUNK);
        g->prev_n_seg = g->n_seg; g->prev_scan_chunks = g->scan_chunks;
        ivm_reset_step_globals(g, S.release == nullptr);
    }
    IVM_TRACE(g, 6, tid);
}

// the whole fix-up in one thread block (multi-kernel step, emulator)
template <class A>
IVM_HD void ivm_fixup_program(const IvmParams &P, const IvmFixScratch &S, int tid, int nthreads) {
    ivm_fixup_stage1<A>(P, S, tid, nthreads);
    A::sync();
    IVM_TRACE(P.g, 3, tid);
    ivm_fixup_scan_block<A>(P, S, 0, 1, tid, nthreads);
    A::sync();
    ivm_fixup_stage2<A>(P, S, tid, nthreads);
}

// ---------------------------------------------------------------------------
// K2 geometry: which store columns of half-row `rr` can hold records that land
// in ego tile rows [r0,r1) x cols [c0,c1)?  Conservative (never drops a record
// that maps into the tile); exactness comes from evaluating the reference
// arithmetic per record afterwards.
struct IvmTileGeom {
    float px, pz, c, s;
    float xe_lo, xe_hi, ze_lo, ze_hi;  // ego-frame metric bounds of the tile, with slack
    float hr;
    float inv_hr, inv_a0, inv_a1;      // reciprocals used by ivm_row_span (spans are conservative bounds, not results:
                                       // an ulp of difference against a true division disappears in the slack)
    int32_t row_lo, row_hi;            // absolute half-rows to visit
};

IVM_HD void ivm_tile_geom(const IvmParams &P, float px, float pz, float c, float s, int r0, int r1, int c0, int c1,
                          IvmTileGeom &G) {
    const float slack = 2.0e-3f;  // metres; covers fp32 rounding of the reference transform (<1e-5 m at 100 m)
    G.px = px; G.pz = pz; G.c = c; G.s = s; G.hr = P.half_res;
    G.inv_hr = 1.0f / P.half_res;
    G.inv_a0 = fabsf(c) < 1.0e-6f ? 0.0f : 1.0f / c;
    G.inv_a1 = fabsf(s) < 1.0e-6f ? 0.0f : 1.0f / (-s);
    G.xe_lo = ((float)c0 - 0.5f) * P.res - P.half_w - slack;
    G.xe_hi = ((float)c1 - 0.5f) * P.res - P.half_w + slack;
    G.ze_lo = ((float)r0 - 0.5f) * P.res - P.half_h - slack;
    G.ze_hi = ((float)r1 - 0.5f) * P.res - P.half_h + slack;
    // world z of the four corners: z1 = s*xe + c*ze
    float zmin = 3.0e38f, zmax = -3.0e38f;
    const float xe[2] = {G.xe_lo, G.xe_hi}, ze[2] = {G.ze_lo, G.ze_hi};
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j) {
            const float z1 = s * xe[i] + c * ze[j] + pz;
            zmin = fminf(zmin, z1); zmax = fmaxf(zmax, z1);
        }
    const float q0 = zmin / G.hr - 0.55f, q1 = zmax / G.hr + 0.55f;
    G.row_lo = (fabsf(q0) < 1.0e9f) ? (int32_t)floorf(q0) : 0;
    G.row_hi = (fabsf(q1) < 1.0e9f) ? (int32_t)ceilf(q1) : -1;
}

// one slab  L <= a*x1 + b*z1 <= U  with z1 in [za, zb]: tighten [xlo, xhi]  (inv_a = 1/a, or 0 if a ~ 0)
IVM_HD void ivm_slab(float a, float inv_a, float b, float L, float U, float za, float zb, float &xlo, float &xhi) {
    const float bz0 = b * za, bz1 = b * zb;
    const float bzmin = fminf(bz0, bz1), bzmax = fmaxf(bz0, bz1);
    if (inv_a == 0.0f) {
        if (bzmax < L || bzmin > U) { xlo = 1.0f; xhi = -1.0f; }  // empty
        return;
    }
    const float p = (L - bzmax) * inv_a, q = (U - bzmin) * inv_a;
    xlo = fmaxf(xlo, a < 0.0f ? q : p); xhi = fminf(xhi, a < 0.0f ? p : q);
}

IVM_HD void ivm_row_span(const IvmTileGeom &G, int32_t rr, int32_t &clo, int32_t &chi) {
    // records of half-row rr have z/hr within rr +- 0.5 (rint of an fp32 quotient)
    const float za = ((float)rr - 0.55f) * G.hr - G.pz, zb = ((float)rr + 0.55f) * G.hr - G.pz;
    float xlo = -1.0e30f, xhi = 1.0e30f;
    ivm_slab(G.c, G.inv_a0, G.s, G.xe_lo, G.xe_hi, za, zb, xlo, xhi);    // xe =  c*x1 + s*z1
    ivm_slab(-G.s, G.inv_a1, G.c, G.ze_lo, G.ze_hi, za, zb, xlo, xhi);   // ze = -s*x1 + c*z1
    if (!(xlo <= xhi)) { clo = 0; chi = -1; return; }
    const float q0 = (xlo + G.px) * G.inv_hr - 0.55f, q1 = (xhi + G.px) * G.inv_hr + 0.55f;
    clo = (q0 > -1.0e9f) ? (int32_t)floorf(q0) : -1000000000;
    chi = (q1 < 1.0e9f) ? (int32_t)ceilf(q1) : 1000000000;
}

// One world record against the ego map: band filter (mapper.py:884-901), translate
// + rotate (mapper.py:255-267: x+(-px); unfused c*x+s*z), cell index
// (mapper.py:101-114).  Returns true and (row, col) if the record is inside the map.
// the ego-frame quotients' numerators: (ze + half_h, xe + half_w)
IVM_HD void ivm_ego_numerators(const IvmParams &P, float x, float z, float px, float pz, float c, float s, float &ar, float &ac) {
    const float x1 = ivm_add(x, -px);
    const float z1 = ivm_add(z, -pz);
    const float xe = ivm_add(ivm_mul(c, x1), ivm_mul(s, z1));
    const float ze = ivm_add(ivm_mul(-s, x1), ivm_mul(c, z1));
    ar = ivm_add(ze, P.half_h); ac = ivm_add(xe, P.half_w);
}
IVM_HD bool ivm_ego_cell(const IvmParams &P, float x, float y, float z, float px, float h, float pz, float c, float s,
                         int32_t &row, int32_t &col) {
    const bool band = y > ivm_sub(h, 1.25f) && y < ivm_add(h, 0.75f);
    float ar, ac;
    ivm_ego_numerators(P, x, z, px, pz, c, s, ar, ac);
    const float rf = ivm_rint_div(ar, P.res, P.inv_res);
    const float cf = ivm_rint_div(ac, P.res, P.inv_res);
    const bool in = band && rf >= 0.0f && rf < (float)P.R && cf >= 0.0f && cf < (float)P.C;
    row = in ? (int32_t)rf : 0; col = in ? (int32_t)cf : 0;
    return in;
}

// The ego tile a world record falls into (if any) still depends on the edge fix-up: stamp it.
IVM_HD void ivm_mark_tile(const IvmParams &P, int b, float x, float y, float z, float px, float h, float pz, float c, float s) {
    int32_t row, col;
    if (!ivm_ego_cell(P, x, y, z, px, h, pz, c, s, row, col)) return;
    const int tiles_x = (P.C + P.tile_c - 1) / P.tile_c, tiles_y = (P.R + P.tile_r - 1) / P.tile_r;
    P.tile_dirty[(size_t)b * (tiles_x * tiles_y) + (row / P.tile_r) * tiles_x + col / P.tile_c] = P.step;
}
