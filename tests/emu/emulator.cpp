// TEST INFRASTRUCTURE ONLY -- serial host emulator of the CUDA step pipeline.
//
// Runs the SAME per-thread logic as the kernels (ivlnce_b200/csrc/ivm_core.h) from
// plain loops over host memory, so the dense-store + edge-fix-up algorithm can be
// diffed against the oracle on a machine without a GPU.  It is never loaded by the
// product package (the product has no CPU path); only tests/ build and use it.
// Pixels can be visited in reverse or strided order to show the result does not
// depend on thread scheduling.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../ivlnce_b200/csrc/ivm_core.h"

struct Emu {
    IvmParams P;
    int mode;
    uint32_t step;
    unsigned long long cstep;
    uint32_t stamp_period;  // 0 = longest
    uint32_t last_cstamp;
    int hi_water;
    int order;  // 0 forward, 1 reverse, 2 strided
    int fix_cap;  // capacity of the fix-up's small-class fast path (0 forces the hash path)
    int direct;   // 1 = direct (partner look-up) resolution of the edge collisions where the boxes allow it
    std::vector<IvmRecord> store;
    std::vector<unsigned long long> cand;
    std::vector<IvmEnv> env;
    std::vector<int32_t> rowcount, colcount, segs;
    IvmGlobal g;
    std::vector<IvmEdge> e1, e2;
    std::vector<unsigned long long> hkeys, hxord;
    std::vector<uint32_t> hbest;
    std::vector<float> xs, ys, T12_buf, cs_buf;
    std::vector<IvmRecord> kpts;
    std::vector<uint32_t> koff;
};

extern "C" {

Emu *emu_create(int H, int W, const float *xs, const float *ys, float res, float half_res, float half_h, float half_w,
                int R, int C, int SR, int SC, int maxB, int tile_r, int tile_c, int mode, long long kcap) {
    Emu *m = new Emu();
    memset(&m->P, 0, sizeof(m->P));
    memset(&m->g, 0, sizeof(m->g));
    m->mode = mode; m->step = 0; m->cstep = 0; m->stamp_period = 0; m->last_cstamp = 0; m->hi_water = 0; m->order = 0; m->fix_cap = 512; m->direct = 0;
    ivm_reset_step_globals(&m->g);
    IvmParams &P = m->P;
    P.H = H; P.W = W; P.HW = H * W; P.R = R; P.C = C;
    P.res = res; P.half_res = half_res; P.half_h = half_h; P.half_w = half_w;
    P.inv_res = ivm_div(1.0f, res); P.inv_half_res = ivm_div(1.0f, half_res);
    P.SR = SR; P.SC = SC; P.maxB = maxB;
    P.pix_bits = ivm_pix_bits((long long)H * W);
    P.tile_r = tile_r > R ? R : tile_r; P.tile_c = tile_c > C ? C : tile_c;
    const size_t cells = (size_t)maxB * SR * SC;
    m->env.assign(maxB, IvmEnv());
    memset(m->env.data(), 0, sizeof(IvmEnv) * maxB);
    m->rowcount.assign((size_t)maxB * SR, 0);
    m->colcount.assign((size_t)maxB * SC, 0);
    m->segs.assign((size_t)16 * maxB, 0);
    uint32_t ecap = (uint32_t)((long long)maxB * 8192 > (1ll << 20) ? (1ll << 20) : (long long)maxB * 8192);
    uint32_t hs = 1024;
    while (hs < 2 * ecap) hs <<= 1;
    m->e1.resize(ecap); m->e2.resize(ecap);
    m->hkeys.assign(hs, IVM_EMPTY_KEY); m->hxord.assign(hs, IVM_EMPTY_KEY); m->hbest.assign(hs, 0u);
    if (xs) m->xs.assign(xs, xs + W);
    if (ys) m->ys.assign(ys, ys + H);
    if (mode == 0) {
        IvmRecord z; z.x = z.y = z.z = 0.f; z.meta = 0;
        m->store.assign(cells, z);
        m->cand.assign(cells, 0ull);
    } else {
        IvmRecord z; z.x = z.y = z.z = 0.f; z.meta = 0;
        m->kpts.assign((size_t)maxB * kcap, z);
        m->koff.assign((size_t)maxB * ((size_t)SR * SC + 1), 0u);
        P.kcap = kcap;
    }
    P.store = m->store.data(); P.cplane = m->cand.data(); P.env = m->env.data();
    P.rowcount = m->rowcount.data(); P.colcount = m->colcount.data(); P.g = &m->g;
    P.e1 = m->e1.data(); P.e2 = m->e2.data(); P.ecap = ecap; P.segs = m->segs.data();
    P.hkeys = m->hkeys.data(); P.hbest = m->hbest.data(); P.hxord = m->hxord.data(); P.hmask = hs - 1;
    P.xs = m->xs.data(); P.ys = m->ys.data();
    m->T12_buf.assign((size_t)12 * maxB, 0.f); m->cs_buf.assign((size_t)2 * maxB, 0.f);
    P.T12_buf = m->T12_buf.data(); P.cs_buf = m->cs_buf.data();
    P.kpts = m->kpts.data(); P.koff = m->koff.data();
    return m;
}

void emu_destroy(Emu *m) { delete m; }
void emu_set_order(Emu *m, int order) { m->order = order; }
void emu_set_step(Emu *m, unsigned step) { m->step = step; m->cstep = step; }  // before the first step only (tests the stamp wrap)
void emu_set_stamp_period(Emu *m, unsigned period) { m->stamp_period = period; }
void emu_set_fix_cap(Emu *m, int cap) { m->fix_cap = cap < 1 ? 1 : cap; }
void emu_set_direct(Emu *m, int on) { m->direct = on; }

static inline int visit(const Emu *m, int i, int n) {
    if (m->order == 1) return n - 1 - i;
    if (m->order == 2) { const int stride = 7919; return (int)(((long long)i * stride) % n); }  // prime stride: a permutation if gcd(n,7919)=1
    return i;
}

// raster of one env tile by tile (k_raster)
static void emu_raster(Emu *m, const IvmParams &P, bool known) {
    const int tr = P.tile_r, tc = P.tile_c;
    std::vector<uint32_t> skey((size_t)tr * tc);
    std::vector<uint8_t> socc((size_t)tr * tc);
    for (int b = 0; b < P.B; ++b) {
        const IvmEnv e = P.env[b];
        const float px = P.pose[3 * b], h = P.pose[3 * b + 1], pz = P.pose[3 * b + 2];
        const float c = P.cs[2 * b], s = P.cs[2 * b + 1];
        for (int r0 = 0; r0 < P.R; r0 += tr)
            for (int c0 = 0; c0 < P.C; c0 += tc) {
                const int r1 = r0 + tr < P.R ? r0 + tr : P.R, c1 = c0 + tc < P.C ? c0 + tc : P.C;
                std::fill(skey.begin(), skey.end(), 0u);
                std::fill(socc.begin(), socc.end(), (uint8_t)0);
                if (e.count > 0) {
                    IvmTileGeom G;
                    ivm_tile_geom(P, px, pz, c, s, r0, r1, c0, c1, G);
                    const int row_lo = G.row_lo > (known ? e.origin_r : e.rmin) ? G.row_lo : (known ? e.origin_r : e.rmin);
                    const int rh = known ? e.origin_r + P.SR - 1 : e.rmax;
                    const int row_hi = G.row_hi < rh ? G.row_hi : rh;
                    for (int rr = row_lo; rr <= row_hi; ++rr) {
                        int clo, chi;
                        ivm_row_span(G, rr, clo, chi);
                        const int cl = known ? e.origin_c : e.cmin, ch = known ? e.origin_c + P.SC - 1 : e.cmax;
                        if (clo < cl) clo = cl;
                        if (chi > ch) chi = ch;
                        if (clo > chi) continue;
                        const size_t rowbase = ((size_t)b * P.SR + (size_t)(rr - e.origin_r)) * P.SC;
                        if (!known) {
                            for (int cc = clo; cc <= chi; ++cc) {
                                const int ccr = cc - e.origin_c;
                                const IvmRecord rec = P.store[rowbase + ccr];
                                if (!ivm_live(rec.meta, e.reset_stamp)) continue;
                                int row, col;
                                if (!ivm_ego_cell(P, rec.x, rec.y, rec.z, px, h, pz, c, s, row, col)) continue;
                                if (row < r0 || row >= r1 || col < c0 || col >= c1) continue;
                                m->g.stats[IVM_STAT_IN]++;
                                const int t = (row - r0) * tc + (col - c0);
                                socc[t] = 1;
                                const uint32_t label = rec.meta & 0xFFu;
                                if (label) {
                                    const uint32_t key = ((uint32_t)((rr - e.origin_r) * P.SC + ccr) << 8) | label;
                                    if (key > skey[t]) skey[t] = key;
                                }
                            }
                        } else {
                            const uint32_t *off = P.koff + (size_t)b * ((size_t)P.SR * P.SC + 1) + (size_t)(rr - e.origin_r) * P.SC;
                            const uint32_t p0 = off[clo - e.origin_c], p1 = off[chi - e.origin_c + 1];
                            const IvmRecord *pts = P.kpts + (size_t)b * P.kcap;
                            for (uint32_t q = p0; q < p1; ++q) {
                                const IvmRecord rec = pts[q];
                                int row, col;
                                if (!ivm_ego_cell(P, rec.x, rec.y, rec.z, px, h, pz, c, s, row, col)) continue;
                                if (row < r0 || row >= r1 || col < c0 || col >= c1) continue;
                                m->g.stats[IVM_STAT_IN]++;
                                const int t = (row - r0) * tc + (col - c0);
                                socc[t] = 1;
                                if ((rec.meta & 0xFFu) && rec.meta > skey[t]) skey[t] = rec.meta;
                            }
                        }
                    }
                }
                for (int rr = 0; rr < r1 - r0; ++rr)
                    for (int cc = 0; cc < c1 - c0; ++cc) {
                        const size_t o = ((size_t)b * P.R + (size_t)(r0 + rr)) * P.C + (size_t)(c0 + cc);
                        P.occ[o] = socc[rr * tc + cc];
                        P.sem[o] = (uint8_t)(skey[rr * tc + cc] & 0xFFu);
                    }
            }
    }
}

int emu_step_iterative(Emu *m, int B, const float *depth, const uint8_t *labels, const float *T12, const float *pose,
                       const float *cs, const void *orient, int orient_f64, const uint8_t *masks, uint8_t *occ,
                       uint8_t *sem) {
    if (m->step >= 0xFFFFFFu) return 4;
    m->step += 1;
    IvmParams P = m->P;
    P.B = B; P.step = m->step;
    m->cstep += 1ull;
    const uint32_t period = ivm_stamp_period(P.pix_bits, m->stamp_period);
    P.cstamp = (uint32_t)((m->cstep - 1ull) % period) + 1u;
    m->last_cstamp = P.cstamp;
    if (P.cstamp == 1u && m->cstep > 1ull) std::fill(m->cand.begin(), m->cand.end(), 0ull);
    P.depth = depth; P.labels = labels; P.T12 = T12; P.pose = pose; P.cs = cs; P.masks = masks;
    P.occ = occ; P.sem = sem;
    if (orient) {  // the K0 path that derives the matrices from the angles
        P.orient = orient; P.orient_f64 = orient_f64; P.T12 = P.T12_buf; P.cs = P.cs_buf;
        for (int b = 0; b < B; ++b) ivm_pose_matrices(P, b, P.T12_buf + 12 * b, P.cs_buf + 2 * b);
        T12 = P.T12; cs = P.cs;
    }
    // K1 head: every env decides its reset / origin; the "first CTA" publishes it (dropped envs are wiped)
    const int nenv = B > m->hi_water ? B : m->hi_water;
    std::vector<IvmEnvPrep> prep(nenv);
    for (int b = 0; b < nenv; ++b) {
        if (b < B) prep[b] = ivm_env_decide(P, b);
        else { prep[b].reset = 1; prep[b].origin_r = 0; prep[b].origin_c = 0; }
    }
    for (int b = 0; b < nenv; ++b) ivm_env_publish<IvmAtomics>(P, b, prep[b], 0, 1);
    m->hi_water = B;
    // K1: scatter
    const int n = B * P.HW;
    for (int i = 0; i < n; ++i) {
        const int gp = visit(m, i, n);
        const int b = gp / P.HW, pix = gp - b * P.HW;
        const int v = pix / P.W, u = pix - v * P.W;
        IvmPoint p;
        const int ok = ivm_unproject(depth[gp], P.xs[u], P.ys[v], T12 + 12 * b, pose[3 * b + 1], P.half_res, P.inv_half_res, p);
        if (ok == 0) continue;
        size_t idx;
        if (ok == 2 || !ivm_store_index(P, prep[b].origin_r, prep[b].origin_c, b, p.r, p.c, idx)) { m->g.err |= IVM_ERR_STORE_OVERFLOW; continue; }
        ivm_cand_insert<IvmAtomics>(P, b, (uint32_t)(idx - (size_t)b * P.SR * P.SC), ivm_cand_key(P, p.y, (uint32_t)pix));
        IvmAtomics::min_i(&m->g.loc[0], p.r); IvmAtomics::max_i(&m->g.loc[1], p.r);
        IvmAtomics::min_i(&m->g.loc[2], p.c); IvmAtomics::max_i(&m->g.loc[3], p.c);
        m->g.acc_valid++;
    }
    // the frame box is final: may stage 1 take the direct path?
    const bool direct1 = m->direct && m->g.loc[0] <= m->g.loc[1] && ivm_box_direct(m->g.loc);
    // K2: resolve (one accumulator per simulated CTA of 1024 pixels, flushed like the kernel does)
    for (int i = 0; i < n; ++i) {
        const int gp = visit(m, n - 1 - i, n);
        const int b = gp / P.HW, pix = gp - b * P.HW;
        const int v = pix / P.W, u = pix - v * P.W;
        IvmPoint p;
        if (ivm_unproject(depth[gp], P.xs[u], P.ys[v], T12 + 12 * b, pose[3 * b + 1], P.half_res, P.inv_half_res, p) != 1) continue;
        IvmBoxAcc acc;
        acc.clear();
        if (direct1)
            m->g.acc_local += (unsigned)ivm_resolve_pixel_direct<IvmAtomics>(P, b, (uint32_t)pix, p, labels[gp], m->g.loc,
                                                                             P.env[b].origin_r, P.env[b].origin_c, acc);
        else
            m->g.acc_local += (unsigned)ivm_resolve_pixel<IvmAtomics>(P, b, (uint32_t)pix, p, labels[gp], m->g.loc,
                                                                      P.env[b].origin_r, P.env[b].origin_c, acc);
        ivm_box_flush<IvmAtomics>(&P.env[b], acc);
    }
    // K3: fix-up
    {
        std::vector<unsigned long long> key(m->fix_cap), xo(m->fix_cap);
        std::vector<uint32_t> ord(m->fix_cap);
        int32_t ibuf[8];
        unsigned long long lbuf[2];
        IvmFixScratch S;
        S.key = key.data(); S.xo = xo.data(); S.ord = ord.data(); S.cap = (uint32_t)m->fix_cap; S.ibuf = ibuf; S.lbuf = lbuf;
        S.release = nullptr; S.release_add = 0u;
        // world box over the env boxes as they stand after the resolve (the direct stage 1 has merged everything)
        int32_t glob[4] = {INT32_MAX, INT32_MIN, INT32_MAX, INT32_MIN};
        for (int b = 0; b < B; ++b) {
            const IvmEnv &e = P.env[b];
            if (e.count > 0) {
                glob[0] = std::min(glob[0], e.rmin); glob[1] = std::max(glob[1], e.rmax);
                glob[2] = std::min(glob[2], e.cmin); glob[3] = std::max(glob[3], e.cmax);
            }
        }
        const bool novalid = m->g.loc[0] > m->g.loc[1];
        if (m->direct && (direct1 || novalid) && glob[0] <= glob[1] && ivm_box_direct(glob)) {
            std::vector<int32_t> hdr((size_t)IVM_SCAN_HDR * 4 * B);
            int32_t counts[2];
            std::vector<uint8_t> tb((size_t)B + 1);
            ivm_direct_segments(P, glob, hdr.data(), 4 * B, counts, 0, tb.data());
            const int nlive = ivm_direct_scan<IvmAtomics>(P, hdr.data(), counts[0], counts[1], glob, m->g.loc, 0, 1, 0, 1);
            m->g.glob[0] = glob[0]; m->g.glob[1] = glob[1]; m->g.glob[2] = glob[2]; m->g.glob[3] = glob[3];
            m->g.n_seg = (uint32_t)counts[0]; m->g.scan_chunks = (uint32_t)((counts[1] + IVM_SCAN_CHUNK - 1) / IVM_SCAN_CHUNK);
            ivm_fixup_stage2<IvmAtomics>(P, S, 0, 1, false, (uint32_t)m->g.acc_e1, (uint32_t)nlive);
        } else {
            ivm_fixup_program<IvmAtomics>(P, S, 0, 1);
        }
    }
    // K4: raster
    emu_raster(m, P, false);
    return 0;
}

int emu_known_load(Emu *m, int b, long long n, const float *xyz, const uint8_t *sem, int origin_r, int origin_c) {
    IvmParams &P = m->P;
    if (n > P.kcap) return 1;
    IvmEnv *e = &P.env[b];
    e->origin_r = origin_r; e->origin_c = origin_c; e->count = (int32_t)n; e->known_n = (int32_t)n;
    e->rmin = origin_r; e->rmax = origin_r + P.SR - 1; e->cmin = origin_c; e->cmax = origin_c + P.SC - 1;
    e->reset_stamp = 0;
    const size_t ncell = (size_t)P.SR * P.SC;
    uint32_t *off = P.koff + (size_t)b * (ncell + 1);
    memset(off, 0, sizeof(uint32_t) * (ncell + 1));
    std::vector<uint32_t> cell(n, 0xFFFFFFFFu);
    for (long long i = 0; i < n; ++i) {
        const float rf = ivm_rint_div(xyz[3 * i + 2], P.half_res, P.inv_half_res), cf = ivm_rint_div(xyz[3 * i], P.half_res, P.inv_half_res);
        const int rr = (int)rf - origin_r, cc = (int)cf - origin_c;
        if (rr < 0 || rr >= P.SR || cc < 0 || cc >= P.SC) { m->g.err |= IVM_ERR_KNOWN_OVERFLOW; continue; }
        cell[i] = (uint32_t)rr * P.SC + cc;
        off[cell[i] + 1]++;
    }
    for (size_t i = 1; i <= ncell; ++i) off[i] += off[i - 1];
    std::vector<uint32_t> fill(ncell, 0);
    IvmRecord *pts = P.kpts + (size_t)b * P.kcap;
    for (long long i = n - 1; i >= 0; --i) {  // any order within a cell is fine: the list index rides in meta
        if (cell[i] == 0xFFFFFFFFu) continue;
        IvmRecord r; r.x = xyz[3 * i]; r.y = xyz[3 * i + 1]; r.z = xyz[3 * i + 2];
        r.meta = ((uint32_t)i << 8) | sem[i];
        pts[off[cell[i]] + fill[cell[i]]++] = r;
    }
    return 0;
}

void emu_known_clear(Emu *m, int b) { m->P.env[b].count = 0; m->P.env[b].known_n = 0; }

int emu_step_known(Emu *m, int B, const float *pose, const float *cs, uint8_t *occ, uint8_t *sem) {
    IvmParams P = m->P;
    P.B = B; P.pose = pose; P.cs = cs; P.occ = occ; P.sem = sem;
    m->g.stats[IVM_STAT_IN] = 0;
    emu_raster(m, P, true);
    return 0;
}

// live records in (env, half-row, half-col) order + the reference list key
long long emu_export_world(Emu *m, int B, long long cap, long long *env_out, float *xyz_out, uint8_t *label_out,
                           unsigned long long *key_out) {
    const IvmParams &P = m->P;
    long long n = 0;
    for (int b = 0; b < B; ++b) {
        const IvmEnv &e = P.env[b];
        if (e.count <= 0) continue;
        for (int rr = 0; rr < P.SR; ++rr)
            for (int cc = 0; cc < P.SC; ++cc) {
                const IvmRecord rec = P.store[((size_t)b * P.SR + rr) * P.SC + cc];
                if (!ivm_live(rec.meta, e.reset_stamp)) continue;
                if (n < cap) {
                    env_out[n] = b;
                    xyz_out[3 * n] = rec.x; xyz_out[3 * n + 1] = rec.y; xyz_out[3 * n + 2] = rec.z;
                    label_out[n] = (uint8_t)(rec.meta & 0xFFu);
                    key_out[n] = ivm_list_key(b, e.origin_r + rr, e.origin_c + cc, m->g.prev_rmin, m->g.prev_cmin, m->g.prev_R,
                                              m->g.prev_C);
                }
                ++n;
            }
    }
    return n;
}

void emu_status(const Emu *m, uint32_t *err, unsigned long long *stats8) {
    *err = m->g.err;
    for (int i = 0; i < IVM_NSTATS; ++i) stats8[i] = m->g.stats[i];
}

// words of the candidate plane that carry the stamp of the last step (= half-cells the last frame touched)
long long emu_cand_current(const Emu *m) {
    long long n = 0;
    const int sh = 32 + m->P.pix_bits;
    for (size_t i = 0; i < m->cand.size(); ++i) n += (uint32_t)(m->cand[i] >> sh) == m->last_cstamp;
    return n;
}

}  // extern "C"

// Exhaustive check of ivm_rint_div against rint of the true fp32 quotient: every float a with lo <= |a| <= hi
// (both signs).  Returns the number of values whose results differ (must be 0); *ambiguous receives how many
// took the true-division fallback.
extern "C" long long emu_check_rint_div(float d, float lo, float hi, long long *ambiguous) {
    const float inv = ivm_div(1.0f, d);
    union { float f; uint32_t u; } a, b;
    a.f = lo; b.f = hi;
    long long bad = 0, amb_n = 0;
    for (uint32_t u = a.u; u <= b.u; ++u) {
        union { float f; uint32_t u; } v;
        v.u = u;
        for (int sgn = 0; sgn < 2; ++sgn) {
            const float x = sgn ? -v.f : v.f;
            bool amb = false;
            (void)ivm_rint_mul(x, inv, amb);
            amb_n += amb;
            const float want = rintf(ivm_div(x, d));
            const float got = ivm_rint_div(x, d, inv);
            bad += !(got == want);
        }
    }
    if (ambiguous) *ambiguous = amb_n;
    return bad;
}
