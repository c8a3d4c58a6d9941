/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle for the MapCMA semantic-map update.
 *
 * A plain-C, single-threaded, point-list restatement of the reference
 * algorithm in /root/reference/ivlnce_baselines/common/mapping_module/
 * (mapper.py, projector/core.py, projector/point_cloud.py).  It is the checker
 * the CUDA path is compared with.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it; the product
 * package never does.
 *
 * Parity pinning: the reference has no tests or golden vectors for this path
 * (SURVEY.md section 4, 8c).  This oracle is pinned instead against outputs of
 * the unmodified reference code run in the build container
 * (tests/golden/make_golden.py writes tests/golden/<name>.npz; see tests/test_oracle_golden.py).
 * The one third-party op on the path, torch_scatter.scatter_max
 * (torch-scatter==2.0.9, reference requirements.txt:24), is not vendored; its
 * published CPU semantics (serial loop, strict '>' update, so the FIRST element
 * attaining the maximum wins; empty groups dropped) are restated in
 * keep_highest() below.
 *
 * Arithmetic follows SURVEY.md Appendix A: fp32, round-to-nearest-even, fused
 * multiply-add exactly where the reference's BLAS call fuses.  Compile with
 * -ffp-contract=off so the compiler adds no fusion of its own.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int64_t n, cap;
    int64_t *b;
    float *x, *y, *z;
    uint8_t *sem;
} cloud_t;

typedef struct {
    int H, W;
    float *xs, *ys;        /* per-column / per-row scale tables, core.py:86-107 */
    float res, half_res;   /* (float)res, (float)(res/2) */
    float half_h, half_w;  /* (float)(height_m/2), (float)(width_m/2) */
    int R, C;              /* MapDimensions.num_rows / num_cols, mapper.py:97-99 */
    cloud_t world;
    int world_is_none;     /* WorldSemanticPointcloud.xyz is None, mapper.py:300 */
    uint8_t *occ, *sem;    /* [map_batch, R, C] */
    int map_batch;
    /* counters of the last step */
    int64_t n_valid, n_local, n_world, n_band, n_in, n_ties;
} oracle_t;

/* ------------------------------------------------------------------ clouds */
static void cloud_reserve(cloud_t *c, int64_t cap) {
    if (cap <= c->cap) return;
    int64_t ncap = c->cap ? c->cap : 1024;
    while (ncap < cap) ncap *= 2;
    c->b = (int64_t *)realloc(c->b, ncap * sizeof(int64_t));
    c->x = (float *)realloc(c->x, ncap * sizeof(float));
    c->y = (float *)realloc(c->y, ncap * sizeof(float));
    c->z = (float *)realloc(c->z, ncap * sizeof(float));
    c->sem = (uint8_t *)realloc(c->sem, ncap * sizeof(uint8_t));
    c->cap = ncap;
}
static void cloud_push(cloud_t *c, int64_t b, float x, float y, float z, uint8_t s) {
    cloud_reserve(c, c->n + 1);
    c->b[c->n] = b; c->x[c->n] = x; c->y[c->n] = y; c->z[c->n] = z; c->sem[c->n] = s;
    c->n++;
}
static void cloud_free(cloud_t *c) {
    free(c->b); free(c->x); free(c->y); free(c->z); free(c->sem);
    memset(c, 0, sizeof(*c));
}
/* keep element i at position j (j <= i) */
static inline void cloud_move(cloud_t *c, int64_t j, int64_t i) {
    c->b[j] = c->b[i]; c->x[j] = c->x[i]; c->y[j] = c->y[i]; c->z[j] = c->z[i]; c->sem[j] = c->sem[i];
}

/* ------------------------------------------------------------------ create */
oracle_t *orc_create(int H, int W, const float *xs, const float *ys, float res, float half_res,
                     float half_h, float half_w, int R, int C) {
    oracle_t *o = (oracle_t *)calloc(1, sizeof(oracle_t));
    o->H = H; o->W = W;
    o->xs = (float *)malloc(sizeof(float) * (W > 0 ? W : 1));
    o->ys = (float *)malloc(sizeof(float) * (H > 0 ? H : 1));
    if (xs) memcpy(o->xs, xs, sizeof(float) * W);
    if (ys) memcpy(o->ys, ys, sizeof(float) * H);
    o->res = res; o->half_res = half_res; o->half_h = half_h; o->half_w = half_w;
    o->R = R; o->C = C;
    o->world_is_none = 1;
    return o;
}
void orc_destroy(oracle_t *o) {
    if (!o) return;
    cloud_free(&o->world);
    free(o->xs); free(o->ys); free(o->occ); free(o->sem);
    free(o);
}

/* ---------------------------------------------------------------- clearing */
/* WorldSemanticPointcloud.clear_completed_episode_data, mapper.py:310-326:
 * drop points of envs >= num_envs (paused), then of envs whose mask == 0. */
void orc_clear(oracle_t *o, int B, const uint8_t *masks) {
    if (o->world_is_none) return;
    cloud_t *w = &o->world;
    int64_t j = 0;
    for (int64_t i = 0; i < w->n; ++i) {
        int64_t b = w->b[i];
        if (b >= B) continue;
        if (masks[b] == 0) continue;
        if (j != i) cloud_move(w, j, i);
        ++j;
    }
    w->n = j;
}

/* -------------------------------------------------------- highest-point dedup */
static void radix_sort_pairs(uint64_t *key, int64_t *val, int64_t n) {
    uint64_t *k2 = (uint64_t *)malloc(n * sizeof(uint64_t));
    int64_t *v2 = (int64_t *)malloc(n * sizeof(int64_t));
    uint64_t all_or = 0;
    for (int64_t i = 0; i < n; ++i) all_or |= key[i];
    uint64_t *ka = key, *kb = k2;
    int64_t *va = val, *vb = v2;
    for (int pass = 0; pass < 8; ++pass) {
        int shift = pass * 8;
        if (((all_or >> shift) & 0xff) == 0) continue; /* digit is zero everywhere */
        int64_t cnt[257];
        memset(cnt, 0, sizeof(cnt));
        for (int64_t i = 0; i < n; ++i) cnt[((ka[i] >> shift) & 0xff) + 1]++;
        for (int d = 0; d < 256; ++d) cnt[d + 1] += cnt[d];
        for (int64_t i = 0; i < n; ++i) {
            int64_t p = cnt[(ka[i] >> shift) & 0xff]++;
            kb[p] = ka[i]; vb[p] = va[i];
        }
        uint64_t *tk = ka; ka = kb; kb = tk;
        int64_t *tv = va; va = vb; vb = tv;
    }
    if (ka != key) { memcpy(key, ka, n * sizeof(uint64_t)); memcpy(val, va, n * sizeof(int64_t)); }
    free(k2); free(v2);
}

/* KeepHighestSemanticPointcloud.forward, mapper.py:433-474.
 *  rows = round(z/(res/2)) - min, cols = round(x/(res/2)) - min (min over ALL
 *  points of ALL envs), key = b*(rows.max()*cols.max()) + rows*cols.max() + cols
 *  (sic: strides are max, not max+1, so bbox-edge cells collide -- kept as is),
 *  per key keep the highest point, first index on ties, output in key order. */
static void keep_highest(oracle_t *o, cloud_t *c) {
    int64_t n = c->n;
    if (n == 0) return;
    int64_t *ri = (int64_t *)malloc(n * sizeof(int64_t));
    int64_t *ci = (int64_t *)malloc(n * sizeof(int64_t));
    int64_t rmin = INT64_MAX, cmin = INT64_MAX;
    for (int64_t i = 0; i < n; ++i) {
        ri[i] = (int64_t)rintf(c->z[i] / o->half_res); /* mapper.py:464 */
        ci[i] = (int64_t)rintf(c->x[i] / o->half_res);
        if (ri[i] < rmin) rmin = ri[i];
        if (ci[i] < cmin) cmin = ci[i];
    }
    int64_t rmax = 0, cmax = 0;
    for (int64_t i = 0; i < n; ++i) {
        ri[i] -= rmin; ci[i] -= cmin; /* mapper.py:465 */
        if (ri[i] > rmax) rmax = ri[i];
        if (ci[i] > cmax) cmax = ci[i];
    }
    uint64_t *key = (uint64_t *)malloc(n * sizeof(uint64_t));
    int64_t *idx = (int64_t *)malloc(n * sizeof(int64_t));
    for (int64_t i = 0; i < n; ++i) {
        key[i] = (uint64_t)(c->b[i] * (rmax * cmax) + ri[i] * cmax + ci[i]); /* mapper.py:469 */
        idx[i] = i;
    }
    radix_sort_pairs(key, idx, n); /* stable: equal keys stay in index order */
    /* scatter_max + drop empty groups, mapper.py:471-474 */
    int64_t *keep = (int64_t *)malloc(n * sizeof(int64_t));
    int64_t m = 0;
    for (int64_t s = 0; s < n;) {
        int64_t e = s;
        float best_h = -3.402823466e+38F; /* numeric_limits<float>::lowest() */
        int64_t best = -1, n_at_max = 0;
        while (e < n && key[e] == key[s]) {
            float h = c->y[idx[e]];
            if (h > best_h) { best_h = h; best = idx[e]; n_at_max = 1; }
            else if (h == best_h && best >= 0) { n_at_max++; }
            ++e;
        }
        if (best >= 0) keep[m++] = best;
        if (n_at_max > 1) o->n_ties++;
        s = e;
    }
    /* semantic_pointcloud.index(argmax_order): gather in key order */
    cloud_t out;
    memset(&out, 0, sizeof(out));
    cloud_reserve(&out, m > 0 ? m : 1);
    for (int64_t j = 0; j < m; ++j) {
        int64_t i = keep[j];
        out.b[j] = c->b[i]; out.x[j] = c->x[i]; out.y[j] = c->y[i]; out.z[j] = c->z[i]; out.sem[j] = c->sem[i];
    }
    out.n = m;
    cloud_free(c);
    *c = out;
    free(ri); free(ci); free(key); free(idx); free(keep);
}

/* --------------------------------------------------------------- ingest */
/* UpdateWorldSemanticPointcloud.forward, mapper.py:825-848 (after the clear):
 *  GenerateSemanticPointCloud (mapper.py:398-425) -> keep_highest(local) ->
 *  world.concatenate(local) -> keep_highest(world).
 *  depth  [B,H,W] f32 normalised; labels [B,H,W] u8; T [B,16] row-major 4x4
 *  camera->world (projector/core.py:6-37 with elevation+pi, mapper.py:132-138);
 *  pose [B,3]. */
void orc_ingest(oracle_t *o, int B, const float *depth, const uint8_t *labels, const float *T,
                const float *pose) {
    const int H = o->H, W = o->W;
    cloud_t local;
    memset(&local, 0, sizeof(local));
    o->n_ties = 0;
    for (int b = 0; b < B; ++b) {
        const float *Tb = T + 16 * b;
        const float h = pose[3 * b + 1];               /* RobotCurrentState.height, mapper.py:128-130 */
        const float lo = h - 1.0f, hi = h + 0.5f;      /* mapper.py:420-424, 248-252 */
        for (int v = 0; v < H; ++v) {
            for (int u = 0; u < W; ++u) {
                const int64_t p = ((int64_t)b * H + v) * W + u;
                const float d = depth[p];
                /* remove_invalid_depth_values(0.01, 0.99), strict, on the normalised depth */
                if (!(d > 0.01f && d < 0.99f)) continue;
                const float zm = d * 10.0f;            /* to_depth_meters, mapper.py:381-384 */
                const float z = zm / 1.0f;             /* core.py:142 */
                const float xc = z * o->xs[u];         /* core.py:143 */
                const float yc = z * o->ys[v];         /* core.py:144 */
                float w3[3];
                for (int r = 0; r < 3; ++r) {          /* bmm(T, xyz1), core.py:171 */
                    float acc = Tb[4 * r + 0] * xc;
                    acc = fmaf(Tb[4 * r + 1], yc, acc);
                    acc = fmaf(Tb[4 * r + 2], z, acc);
                    acc = fmaf(Tb[4 * r + 3], 1.0f, acc);
                    w3[r] = acc - 0.0f;                /* world_shift_origin = 0, core.py:215 */
                }
                if (!(w3[1] > lo && w3[1] < hi)) continue;
                cloud_push(&local, b, w3[0], w3[1], w3[2], labels[p]);
            }
        }
    }
    o->n_valid = local.n;
    keep_highest(o, &local);
    o->n_local = local.n;
    /* world.concatenate(local): world first, mapper.py:299-308 */
    if (o->world_is_none) {
        cloud_free(&o->world);
        o->world = local;
        o->world_is_none = 0;
    } else {
        cloud_t *w = &o->world;
        cloud_reserve(w, w->n + local.n);
        for (int64_t i = 0; i < local.n; ++i)
            cloud_push(w, local.b[i], local.x[i], local.y[i], local.z[i], local.sem[i]);
        cloud_free(&local);
    }
    keep_highest(o, &o->world);
    o->n_world = o->world.n;
}

/* known-map mode: GetGTWorldSemanticPointcloud.forward, mapper.py:862-881.
 * The caller (python wrapper) clears first, then appends the scene cloud of each
 * reset env in increasing env order.  No de-duplication happens in this mode. */
void orc_append_cloud(oracle_t *o, int b, int64_t n, const float *xyz, const uint8_t *sem) {
    cloud_t *w = &o->world;
    o->world_is_none = 0;
    cloud_reserve(w, w->n + n);
    for (int64_t i = 0; i < n; ++i) cloud_push(w, b, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], sem[i]);
    o->n_world = w->n;
}

/* --------------------------------------------------------------- raster */
/* FilterPointCloudByRobotHeight (mapper.py:884-901) + OccupancySemanticMapMemory
 * .update (mapper.py:625-636): DenseMap.update (555-567) -> shift_origin
 * (255-267) -> project_to_map_indices (101-114) -> update_map (569-571).
 *  pose [B,3]; cs [B,2] = (f32 cos(-heading), f32 sin(-heading)). */
void orc_raster(oracle_t *o, int B, const float *pose, const float *cs) {
    const int R = o->R, C = o->C;
    if (o->map_batch != B) { /* DenseMap.create_data, mapper.py:533-553 (values are re-zeroed anyway) */
        o->occ = (uint8_t *)realloc(o->occ, (size_t)B * R * C + 1);
        o->sem = (uint8_t *)realloc(o->sem, (size_t)B * R * C + 1);
        o->map_batch = B;
    }
    memset(o->occ, 0, (size_t)B * R * C); /* data.fill_(0) */
    memset(o->sem, 0, (size_t)B * R * C);
    o->n_band = 0; o->n_in = 0;
    if (o->world_is_none) return;
    const cloud_t *w = &o->world;
    for (int64_t i = 0; i < w->n; ++i) {
        const int64_t b = w->b[i];
        if (b >= B) continue; /* cannot happen after orc_clear */
        const float h = pose[3 * b + 1];
        const float y = w->y[i];
        if (!(y > h - 1.25f && y < h + 0.75f)) continue; /* mapper.py:885, 248-252 */
        o->n_band++;
        const float x1 = w->x[i] + (-pose[3 * b + 0]); /* translate(-pose), mapper.py:255-256 */
        const float z1 = w->z[i] + (-pose[3 * b + 2]);
        const float c = cs[2 * b + 0], s = cs[2 * b + 1];
        /* bmm([P,3,3],[P,3,1]) small-matrix path: unfused products, sequential adds */
        const float xe = c * x1 + s * z1;
        const float ze = (-s) * x1 + c * z1;
        const float rf = rintf((ze + o->half_h) / o->res); /* mapper.py:101-114 */
        const float cf = rintf((xe + o->half_w) / o->res);
        if (!(rf >= 0.0f && rf < (float)R && cf >= 0.0f && cf < (float)C)) continue;
        o->n_in++;
        const size_t cell = ((size_t)b * R + (size_t)rf) * C + (size_t)cf;
        o->occ[cell] = 1;                               /* OccupancyStatus.OCCUPIED */
        if (w->sem[i] != 0) o->sem[cell] = w->sem[i];   /* exclude FLOOR(0); last write wins */
    }
}

/* PredictSemantics tail, mapper.py:795-798: scores.argmax(1).to(uint8); first
 * maximal index wins, NaN counts as maximal (torch.argmax).  scores [B,Cls,HW]. */
void orc_argmax_labels(const float *scores, int B, int Cls, int64_t HW, uint8_t *out) {
    for (int b = 0; b < B; ++b)
        for (int64_t p = 0; p < HW; ++p) {
            const float *s = scores + (int64_t)b * Cls * HW + p;
            float best = s[0];
            int arg = 0;
            for (int k = 1; k < Cls; ++k) {
                const float v = s[(int64_t)k * HW];
                if ((v > best) || (v != v && best == best)) { best = v; arg = k; }
            }
            out[(int64_t)b * HW + p] = (uint8_t)arg;
        }
}

/* ----------------------------------------------------------------- getters */
const uint8_t *orc_occupancy(const oracle_t *o) { return o->occ; }
const uint8_t *orc_semantic(const oracle_t *o) { return o->sem; }
int64_t orc_world_size(const oracle_t *o) { return o->world_is_none ? 0 : o->world.n; }
void orc_world_export(const oracle_t *o, int64_t *b, float *xyz, uint8_t *sem) {
    if (o->world_is_none) return;
    const cloud_t *w = &o->world;
    for (int64_t i = 0; i < w->n; ++i) {
        b[i] = w->b[i];
        xyz[3 * i] = w->x[i]; xyz[3 * i + 1] = w->y[i]; xyz[3 * i + 2] = w->z[i];
        sem[i] = w->sem[i];
    }
}
void orc_counters(const oracle_t *o, int64_t *out6) {
    out6[0] = o->n_valid; out6[1] = o->n_local; out6[2] = o->n_world;
    out6[3] = o->n_band; out6[4] = o->n_in; out6[5] = o->n_ties;
}
