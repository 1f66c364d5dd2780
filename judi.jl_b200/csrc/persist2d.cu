// persist2d.cu -- 2-D time loops as ONE persistent cooperative kernel: isotropic acoustic with
// constant density (acoustic2d_persist) and the flux-form physics (flux2d_persist, below).
//
// A 2-D grid of the reference's test/FWI sizes (481x201 points) is ~1 us of HBM time per step, so a
// launch-per-step schedule is pure launch latency.  This kernel keeps the whole loop of one sweep
// (forward, adjoint, Born, adjoint + imaging condition) on the device with one grid-wide barrier
// per time step.  Per step it executes exactly the reference schedule (operators.py:131,188,231):
//   interpolation of level t  ->  stencil  ->  injection into the new level  ->  Born source /
//   imaging condition with the post-injection value  ->  snapshot save / illumination.
// (Free-surface models keep the launch-per-step path.)
// Injection is "owner computes": the thread that updates a grid point looks the point up in a
// cell map and adds the sorted (trace, weight) contributions itself, so u.dt2 / v.dt2 seen by the
// Born source and the imaging condition are post-injection by construction.
#include "common.cuh"
#include "propagate.h"
#include "tma_util.cuh"
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

struct P2Args {
    Grid g;
    Coefs cf;
    float *b[2];       // background field buffers; level t is b[cur]
    float *l[2];       // linearised field buffers (Born) or nullptr
    int cur;
    const float *m, *dx, *dz, *dm, *dms;   // dms: dm as the factor of u.dt2 (judi_model::dm_src)
    float sdamp, irho_c;
    int t0, t1, dir;   // time indices t0, t0+dir, ..., t1
    // injected traces (sorted cell layout, see sparse.cu)
    const int *cellmap;
    const int *cell_start, *ent_trace;
    const float *ent_w, *cell_scale, *cell_born;
    const float *src;
    int nsrc;
    // interpolated traces: set 1 samples the linearised field when Born, else the background
    int nrec;
    const long long *coff;
    const float *cw;
    float *rec;
    float *rec_bg;     // second output sampling the background field (nlind), or nullptr
    // fused extras
    float *snap;       // snapshot pool (slot s at snap + s*slot) or nullptr
    const float *usnap;
    float *grad, *illum;
    long long slot;
    int t_sub;
    float gcoef;
    int t_base;        // snapshot slot of time index t is (t - t_base) / t_sub
    int ic;            // JUDI_IC_*: imaging condition of the gradient and of the Born source
    Region reg;
    // space-time source of the background field / extended receiver (fields_exprs.py:58-103)
    const float *qext, *qwt, *qwf;
    float qbase;
    float *wrec;
    const float *wrt;
    float wdt;
    int fs;            // free surface: mirror the new level about z = 0 (fields_exprs.py:106-137)
};

template <int R>
__device__ __forceinline__ float lap2d(const Grid &g, const Coefs &cf, const float *__restrict__ u,
                                       long long i)
{
    float acc = cf.c0 * u[i];
#pragma unroll
    for (int k = 1; k <= R; k++) acc += cf.cz[k] * (u[i + k * g.sz] + u[i - k * g.sz]);
#pragma unroll
    for (int k = 1; k <= R; k++) acc += cf.cx[k] * (u[i + k] + u[i - k]);
    return acc;
}

// staggered first derivative D+ of a Grid field along stride s (value at i + 1/2)
template <int R>
__device__ __forceinline__ float dplus2d(const float *__restrict__ f, long long i, long long s, const float *w)
{
    float acc = 0.f;
#pragma unroll
    for (int k = 1; k <= R; k++) acc += w[k] * (f[i + k * s] - f[i - (k - 1) * s]);
    return acc;
}

// grad(u_s, +1/2) . grad(v, +1/2) with the snapshot read through its region (zero outside):
// the isic / fwi imaging conditions (sensitivity.py:106-122, 212-228)
template <int R>
__device__ float inner_grad2d(const Grid &g, const Coefs &cf, const Region &reg, const float *us,
                              const float *__restrict__ v, int x, int z)
{
    float gux = 0.f, guz = 0.f;
#pragma unroll
    for (int k = 1; k <= R; k++) {
        const float xp = reg.has(x + k, 0, z) ? us[reg.at(x + k, 0, z)] : 0.f;
        const float xm = reg.has(x - (k - 1), 0, z) ? us[reg.at(x - (k - 1), 0, z)] : 0.f;
        const float zp = reg.has(x, 0, z + k) ? us[reg.at(x, 0, z + k)] : 0.f;
        const float zm = reg.has(x, 0, z - (k - 1)) ? us[reg.at(x, 0, z - (k - 1))] : 0.f;
        gux += cf.wx[k] * (xp - xm);
        guz += cf.wz[k] * (zp - zm);
    }
    const long long i = g.at(x, 0, z);
    return gux * dplus2d<R>(v, i, 1, cf.wx) + guz * dplus2d<R>(v, i, g.sz, cf.wz);
}

// div(dm grad(u, +1/2), -1/2): the isic / fwi Born sources (sensitivity.py:191-209), constant irho
template <int R>
__device__ float lap_weighted2d(const Grid &g, const Coefs &cf, const float *__restrict__ u,
                                const float *__restrict__ c, long long i)
{
    float acc = 0.f;
#pragma unroll
    for (int k = 1; k <= R; k++) {
        const long long xp = i + (k - 1), xm = i - k, zp = i + (k - 1) * g.sz, zm = i - k * g.sz;
        acc += cf.wx[k] * (c[xp] * dplus2d<R>(u, xp, 1, cf.wx) - c[xm] * dplus2d<R>(u, xm, 1, cf.wx));
        acc += cf.wz[k] * (c[zp] * dplus2d<R>(u, zp, g.sz, cf.wz) - c[zm] * dplus2d<R>(u, zm, g.sz, cf.wz));
    }
    return acc;
}

// lb / nb: index of this CTA among the nb CTAs that work on the shot described by `a`.
// (Staging the haloed tile of u(t) in shared memory was measured slower in the batched launch, 55
// vs 34 ms for the forward sweeps of 16 C2 shots: the L1 already serves the stencil's reuse.)
// XIC: compiled with the isic / fwi imaging conditions and Born sources (more registers: the "as"
// instantiation keeps its occupancy)
// VEC: every thread owns 4 consecutive x-points (128-bit loads and stores, a 128 x 8 tile per CTA
// visit).  The one-point-per-thread body spends two thirds of its instructions on index arithmetic
// and runs 16 C2 shots in lockstep at 60 Gpt/s, close to its own issue limit; the vector body
// amortises that arithmetic over 4 points.  Plain / save / imaging condition "as" / illumination
// of the constant-density kernel without free surface, space-time sources or Born (host: p2_vec_ok).
template <int R, bool XIC, int VEC = 0, bool VBORN = false>
__device__ __forceinline__ void acoustic2d_body(const P2Args &a, const int lb, const int nb)
{
    cg::grid_group grid = cg::this_grid();
    const Grid &g = a.g;
    const int TX = 32, TZ = 8;
    const int ntx = (g.n[0] + TX - 1) / TX, ntz = (g.n[2] + TZ - 1) / TZ;
    const int ntiles = ntx * ntz;
    const int tx = threadIdx.x % TX, tz = threadIdx.x / TX;
    const long long gtid = (long long)lb * blockDim.x + threadIdx.x;
    const long long gthreads = (long long)nb * blockDim.x;
    int cur = a.cur;
    // snapshot phase and slot offset of time index t, advanced step by step (no division in the loop)
    int ph = (a.t0 - a.t_base) % a.t_sub;
    long long sidx = (long long)((a.t0 - a.t_base) / a.t_sub) * a.slot;
    // grid coordinates of this thread's first tile (the only one when the grid has a CTA per tile)
    const int x0 = (lb % ntx) * TX + tx, z0 = (lb / ntx) * TZ + tz;
    for (int t = a.t0;; t += a.dir) {
        const float *uc = a.b[cur];
        float *un = a.b[1 - cur];
        const float *lc = a.l[0] ? a.l[cur] : nullptr;
        float *ln = a.l[0] ? a.l[1 - cur] : nullptr;
        // ---- receivers sample level t ------------------------------------------------------
        for (long long p = gtid; p < a.nrec; p += gthreads) {
            float acc = 0.f, acc2 = 0.f;
            const float *f1 = lc ? lc : uc;
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const long long o = a.coff[p * 4 + c];
                if (o >= 0) {
                    const float w = a.cw[p * 4 + c];
                    acc += w * f1[o];
                    if (a.rec_bg) acc2 += w * uc[o];
                }
            }
            a.rec[(long long)t * a.nrec + p] = acc;
            if (a.rec_bg) a.rec_bg[(long long)t * a.nrec + p] = acc2;
        }
        const bool fused_t = ph == 0;
        float *snap_t = (a.snap && fused_t) ? a.snap + sidx : nullptr;
        const float *usnap_t = (a.grad && fused_t) ? a.usnap + sidx : nullptr;
        const float *src_row = a.src ? a.src + (long long)t * a.nsrc : nullptr;
        if (a.dir > 0) { if (++ph == a.t_sub) { ph = 0; sidx += a.slot; } }
        else if (ph == 0) { ph = a.t_sub - 1; sidx -= a.slot; }
        else ph--;
        // the snapshot the NEXT imaging step reads (sidx already points at it): its lines are pulled
        // into L2 while this step runs -- the adjoint sweep reads the whole pool from DRAM, one
        // dependent load per tile visit (profiles/r02_p2d_resident_experiment.md)
        const float *usnap_n = (usnap_t && t != a.t1 && sidx >= 0) ? a.usnap + sidx : nullptr;
        // ---- stencil + injection + fused extras, tile by tile --------------------------------
        if constexpr (VEC) {
            constexpr int NV = (R + 3) / 4, HRV = 4 * NV, NW = 4 + 2 * HRV;
            const int ntxv = (g.n[0] + 4 * TX - 1) / (4 * TX);
            const int ntilesv = ntxv * ntz;
            const bool reg_al = (a.reg.lo[0] & 3) == 0;
            for (int tile = lb; tile < ntilesv; tile += nb) {
                const int x = (tile % ntxv) * (4 * TX) + 4 * tx, z = (tile / ntxv) * TZ + tz;
                if (x >= g.n[0] || z >= g.n[2]) continue;
                const long long i = g.at(x, 0, z);
                float w[NW];
#pragma unroll
                for (int v = 0; v < NW / 4; v++) {
                    const float4 t4 = *(const float4 *)(uc + i - HRV + 4 * v);
                    w[4 * v] = t4.x; w[4 * v + 1] = t4.y; w[4 * v + 2] = t4.z; w[4 * v + 3] = t4.w;
                }
                float lap[4];
#pragma unroll
                for (int e = 0; e < 4; e++) lap[e] = a.cf.c0 * w[HRV + e];
#pragma unroll
                for (int k = 1; k <= R; k++) {
                    const float4 p4 = *(const float4 *)(uc + i + k * g.sz), m4 = *(const float4 *)(uc + i - k * g.sz);
                    lap[0] += a.cf.cz[k] * (p4.x + m4.x); lap[1] += a.cf.cz[k] * (p4.y + m4.y);
                    lap[2] += a.cf.cz[k] * (p4.z + m4.z); lap[3] += a.cf.cz[k] * (p4.w + m4.w);
                }
#pragma unroll
                for (int k = 1; k <= R; k++)
#pragma unroll
                    for (int e = 0; e < 4; e++) lap[e] += a.cf.cx[k] * (w[HRV + e + k] + w[HRV + e - k]);
                const float4 dx4 = *(const float4 *)(a.dx + x), mm4 = *(const float4 *)(a.m + i);
                const float4 up4 = *(const float4 *)(un + i);
                const float dxs[4] = {dx4.x, dx4.y, dx4.z, dx4.w}, mms[4] = {mm4.x, mm4.y, mm4.z, mm4.w};
                const float ups[4] = {up4.x, up4.y, up4.z, up4.w};
                const float dzv = a.dz[z];
                int rows[4] = {-1, -1, -1, -1};
                if (a.cellmap) {
                    const int4 r4 = *(const int4 *)(a.cellmap + i);
                    rows[0] = r4.x; rows[1] = r4.y; rows[2] = r4.z; rows[3] = r4.w;
                }
                float dl[4], unew[4];
                [[maybe_unused]] float sv[4], rd[4];
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    const float ucv = w[HRV + e];
                    const float s_ = ((dxs[e] + 0.f) + dzv) * a.sdamp;
                    const float rden = 1.0f / (mms[e] + s_);
                    if (VBORN) { sv[e] = s_; rd[e] = rden; }
                    const float d1 = ucv - ups[e];
                    const float delta = (lap[e] - s_ * d1) * rden;
                    float inj = 0.f;
                    if (rows[e] >= 0 && x + e < g.n[0]) {
                        float acc = 0.f;
                        for (int en = a.cell_start[rows[e]]; en < a.cell_start[rows[e] + 1]; en++)
                            acc += a.ent_w[en] * src_row[a.ent_trace[en]];
                        inj = acc * a.cell_scale[rows[e]];
                    }
                    unew[e] = ((ucv + d1) + delta) + inj;
                    dl[e] = delta + inj;
                }
                if (x + 3 < g.n[0]) *(float4 *)(un + i) = make_float4(unew[0], unew[1], unew[2], unew[3]);
                else {
#pragma unroll
                    for (int e = 0; e < 4; e++)
                        if (x + e < g.n[0]) un[i + e] = unew[e];
                }
                if constexpr (VBORN) {
                    // linearised field with the Born source -dm (u+ - 2u + u-), post-injection
                    // (sensitivity.py:174-188), same arithmetic as the one-point-per-thread body
                    float wl[NW];
#pragma unroll
                    for (int v = 0; v < NW / 4; v++) {
                        const float4 t4 = *(const float4 *)(lc + i - HRV + 4 * v);
                        wl[4 * v] = t4.x; wl[4 * v + 1] = t4.y; wl[4 * v + 2] = t4.z; wl[4 * v + 3] = t4.w;
                    }
                    float lapl[4];
#pragma unroll
                    for (int e = 0; e < 4; e++) lapl[e] = a.cf.c0 * wl[HRV + e];
#pragma unroll
                    for (int k = 1; k <= R; k++) {
                        const float4 p4 = *(const float4 *)(lc + i + k * g.sz), m4 = *(const float4 *)(lc + i - k * g.sz);
                        lapl[0] += a.cf.cz[k] * (p4.x + m4.x); lapl[1] += a.cf.cz[k] * (p4.y + m4.y);
                        lapl[2] += a.cf.cz[k] * (p4.z + m4.z); lapl[3] += a.cf.cz[k] * (p4.w + m4.w);
                    }
#pragma unroll
                    for (int k = 1; k <= R; k++)
#pragma unroll
                        for (int e = 0; e < 4; e++) lapl[e] += a.cf.cx[k] * (wl[HRV + e + k] + wl[HRV + e - k]);
                    const float4 lp4 = *(const float4 *)(ln + i), dm4 = *(const float4 *)(a.dms + i);
                    const float lps[4] = {lp4.x, lp4.y, lp4.z, lp4.w}, dmv[4] = {dm4.x, dm4.y, dm4.z, dm4.w};
                    float lnew[4];
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        const float lcv = wl[HRV + e];
                        const float d1l = lcv - lps[e];
                        const float q = -dmv[e] * dl[e];
                        lnew[e] = (lcv + d1l) + (lapl[e] - sv[e] * d1l + q) * rd[e];
                    }
                    if (x + 3 < g.n[0]) *(float4 *)(ln + i) = make_float4(lnew[0], lnew[1], lnew[2], lnew[3]);
                    else {
#pragma unroll
                        for (int e = 0; e < 4; e++)
                            if (x + e < g.n[0]) ln[i + e] = lnew[e];
                    }
                }
                if ((snap_t || usnap_t || a.illum) && z >= a.reg.lo[2] && z < a.reg.hi[2] &&
                    x + 3 >= a.reg.lo[0] && x < a.reg.hi[0]) {
                    const long long ri = a.reg.at(x, 0, z);
                    if (reg_al && x >= a.reg.lo[0] && x + 3 < a.reg.hi[0]) {
                        if (snap_t) *(float4 *)(snap_t + ri) = make_float4(w[HRV], w[HRV + 1], w[HRV + 2], w[HRV + 3]);
                        if (a.illum) {
                            float4 I4 = *(float4 *)(a.illum + ri);
                            I4.x += a.cf.dt * w[HRV] * w[HRV]; I4.y += a.cf.dt * w[HRV + 1] * w[HRV + 1];
                            I4.z += a.cf.dt * w[HRV + 2] * w[HRV + 2]; I4.w += a.cf.dt * w[HRV + 3] * w[HRV + 3];
                            *(float4 *)(a.illum + ri) = I4;
                        }
                        if (usnap_t) {
                            if (usnap_n) prefetch_l2(usnap_n + ri);
                            const float4 us4 = *(const float4 *)(usnap_t + ri);
                            float4 G4 = *(float4 *)(a.grad + ri);
                            G4.x -= a.gcoef * a.irho_c * us4.x * dl[0]; G4.y -= a.gcoef * a.irho_c * us4.y * dl[1];
                            G4.z -= a.gcoef * a.irho_c * us4.z * dl[2]; G4.w -= a.gcoef * a.irho_c * us4.w * dl[3];
                            *(float4 *)(a.grad + ri) = G4;
                        }
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; e++) {
                            if (x + e < a.reg.lo[0] || x + e >= a.reg.hi[0]) continue;
                            if (snap_t) snap_t[ri + e] = w[HRV + e];
                            if (a.illum) a.illum[ri + e] += a.cf.dt * w[HRV + e] * w[HRV + e];
                            if (usnap_t) a.grad[ri + e] -= a.gcoef * a.irho_c * usnap_t[ri + e] * dl[e];
                        }
                    }
                }
            }
        } else
        for (int tile = lb; tile < ntiles; tile += nb) {
            int x = x0, z = z0;
            if (tile != lb) { x = (tile % ntx) * TX + tx; z = (tile / ntx) * TZ + tz; }
            if (x >= g.n[0] || z >= g.n[2]) continue;
            const long long i = g.at(x, 0, z);
            const float ucv = uc[i];
            float lap = lap2d<R>(g, a.cf, uc, i);
            if (a.qext) lap += (a.qbase * a.qwt[t]) * a.qext[i];
            else if (a.qwf) lap += a.qbase * a.qwf[((long long)t * g.n[2] + z) * g.n[0] + x];
            const float s = ((a.dx[x] + 0.f) + a.dz[z]) * a.sdamp;
            const float mm = a.m[i];
            const float rden = 1.0f / (mm + s);
            const float d1 = ucv - un[i];
            float delta = (lap - s * d1) * rden;
            float inj = 0.f;
            int row = -1;
            if (a.cellmap && (row = a.cellmap[i]) >= 0) {
                float acc = 0.f;
                for (int e = a.cell_start[row]; e < a.cell_start[row + 1]; e++)
                    acc += a.ent_w[e] * src_row[a.ent_trace[e]];
                inj = acc * a.cell_scale[row];
            }
            const float unew = ((ucv + d1) + delta) + inj;
            // free surface after the injection: u(-k) = -u(k-1), k = 1..R, then u(0) = 0 -- every
            // mirrored value comes from the thread that owns its source point
            if (a.fs && z < R) un[i - (long long)(2 * z + 1) * g.sz] = -unew;
            un[i] = (a.fs && z == 0) ? 0.f : unew;
            if (ln) {
                const float lcv = lc[i];
                const float lapl = lap2d<R>(g, a.cf, lc, i);
                const float d1l = lcv - ln[i];
                // Born source from the post-injection u+ - 2u + u-  (sensitivity.py:174-209)
                float q;
                if (!XIC || a.ic == JUDI_IC_AS) q = -a.dms[i] * (delta + inj);
                else {
                    const float ics = a.ic == JUDI_IC_FWI ? -1.f : 1.f;
                    q = -(a.dms[i] * mm * (delta + inj) -
                          ics * a.cf.dt2 * lap_weighted2d<R>(g, a.cf, uc, a.dm, i));
                }
                const float lnew = (lcv + d1l) + (lapl - s * d1l + q) * rden;
                if (a.fs && z < R) ln[i - (long long)(2 * z + 1) * g.sz] = -lnew;
                ln[i] = (a.fs && z == 0) ? 0.f : lnew;
            }
            if ((snap_t || usnap_t || a.illum || a.wrec) && a.reg.has(x, 0, z)) {
                const long long ri = a.reg.at(x, 0, z);
                if (a.wrec) a.wrec[ri] += (a.wdt * a.wrt[t]) * ucv;
                if (snap_t) snap_t[ri] = ucv;
                if (a.illum) a.illum[ri] += a.cf.dt * ucv * ucv;
                if (usnap_t) {
                    if (!XIC || a.ic == JUDI_IC_AS) a.grad[ri] -= a.gcoef * a.irho_c * usnap_t[ri] * (delta + inj);
                    else {
                        const float ics = a.ic == JUDI_IC_FWI ? -1.f : 1.f;
                        const float ig = inner_grad2d<R>(g, a.cf, a.reg, usnap_t, uc, x, z);
                        a.grad[ri] -= a.gcoef * a.irho_c * (usnap_t[ri] * (delta + inj) * mm + ics * a.cf.dt2 * ig);
                    }
                }
            }
        }
        grid.sync();
        cur = 1 - cur;
        if (t == a.t1) break;
    }
}

// VEC: 0 = one point per thread; n > 0 = four points per thread, compiled for n CTAs per SM
template <int R, bool XIC, int VEC = 0>
__global__ void __launch_bounds__(256, VEC ? VEC : (XIC ? 2 : 4)) acoustic2d_persist(P2Args a)
{
    acoustic2d_body<R, XIC, VEC>(a, blockIdx.x, gridDim.x);
}

// Several shots of the same model and time axis in lockstep: CTA b works on shot b / per_shot.
// One grid barrier per time step serves all of them (judi_fg_multishot on small 2-D grids).
template <int R, bool XIC, int VEC = 0, bool VBORN = false>
__global__ void __launch_bounds__(256, VEC ? (VBORN ? 2 : VEC) : (XIC ? 2 : 4)) acoustic2d_persist_batch(const P2Args *__restrict__ batch, int per_shot)
{
    // the shot's arguments go to shared memory once: read through the global pointer, every stencil
    // coefficient would be a load from L1/L2 instead of a constant-bank operand
    __shared__ P2Args sa;
    {
        const int *src = (const int *)(batch + blockIdx.x / per_shot);
        int *dst = (int *)&sa;
        for (int i = threadIdx.x; i < (int)(sizeof(P2Args) / sizeof(int)); i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    acoustic2d_body<R, XIC, VEC, VBORN>(sa, blockIdx.x % per_shot, per_shot);
}

// ------------------------------------------------------------------------------------------
// The same persistent time loop for the flux-form physics in 2-D (variable density, TTI):
// per step  receivers  ->  flux pass (P, M on the grid extended by R nodes)  -> grid barrier ->
// divergence + update + injection + fused extras -> grid barrier.  Two barriers per step instead
// of three launches (flux_kernel, flux_update, sparse_kernel: 12.8 us per step measured on the C2
// grid).  Plain / save / imaging condition "as" / illumination; Born keeps the launch-per-step path.
struct P2FluxArgs {
    Grid g;
    Coefs cf;
    float *b[2][2];       // [field][buffer]; level t of field f is b[f][cur]
    float *P[2], *Mx[2];  // flux scratch: P_x, P_z and M_x, M_z
    int cur;
    const float *m, *irho, *eps, *delta, *st, *ct;
    const float *dx, *dz;
    int t0, t1, dir;
    const int *cellmap;
    const int *cell_start, *ent_trace;
    const float *ent_w, *cell_scale;
    const float *src;
    int nsrc;
    int nrec;
    const long long *coff;
    const float *cw;
    float *rec;
    float *snap;          // snapshot pool: slot s of field f at snap + (s * nf + f) * slot
    const float *usnap;
    float *grad, *illum;
    long long slot;
    int t_sub;
    float gcoef;
    int t_base;
    Region reg;
};

template <int R, bool TTI>
__device__ __forceinline__ void flux2d_body(const P2FluxArgs &a, const int lb, const int nb)
{
    cg::grid_group grid = cg::this_grid();
    const Grid &g = a.g;
    const Coefs &cf = a.cf;
    constexpr int NF = TTI ? 2 : 1;
    const int TX = 32, TZ = 8;
    const int tx = threadIdx.x % TX, tz = threadIdx.x / TX;
    const long long gtid = (long long)lb * blockDim.x + threadIdx.x;
    const long long gthreads = (long long)nb * blockDim.x;
    // flux nodes: x in [-R, nx + R), z in [-R, nz + R)
    const int fnx = g.n[0] + 2 * R, fnz = g.n[2] + 2 * R;
    const int ftx = (fnx + TX - 1) / TX, ftiles = ftx * ((fnz + TZ - 1) / TZ);
    const int ntx = (g.n[0] + TX - 1) / TX, ntiles = ntx * ((g.n[2] + TZ - 1) / TZ);
    int cur = a.cur;
    for (int t = a.t0;; t += a.dir) {
        const float *u1 = a.b[0][cur], *u2 = TTI ? a.b[1][cur] : nullptr;
        // ---- receivers sample level t of the first field -------------------------------------
        for (long long p = gtid; p < a.nrec; p += gthreads) {
            float acc = 0.f;
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const long long o = a.coff[p * 4 + c];
                if (o >= 0) acc += a.cw[p * 4 + c] * u1[o];
            }
            a.rec[(long long)t * a.nrec + p] = acc;
        }
        // ---- pass 1: rotated fluxes (FD_utils.py:24-105) or irho * grad ------------------------
        for (int tile = lb; tile < ftiles; tile += nb) {
            const int x = (tile % ftx) * TX + tx - R, z = (tile / ftx) * TZ + tz - R;
            if (x >= g.n[0] + R || z >= g.n[2] + R) continue;
            const long long i = g.at(x, 0, z);
            float g1x = 0.f, g1z = 0.f, g2x = 0.f, g2z = 0.f;
#pragma unroll
            for (int k = 1; k <= R; k++) {
                g1x += cf.wx[k] * (u1[i + k] - u1[i - (k - 1)]);
                g1z += cf.wz[k] * (u1[i + k * g.sz] - u1[i - (k - 1) * g.sz]);
                if (TTI) {
                    g2x += cf.wx[k] * (u2[i + k] - u2[i - (k - 1)]);
                    g2z += cf.wz[k] * (u2[i + k * g.sz] - u2[i - (k - 1) * g.sz]);
                }
            }
            const float bb = a.irho[i];
            if (!TTI) {
                a.P[0][i] = bb * g1x;
                a.P[1][i] = bb * g1z;
            } else {
                // closed form through the symmetry axis r = (sin theta, cos theta) (flux3d.cu)
                const float e = a.eps[i], dl = a.delta[i], rx = a.st[i], rz = a.ct[i];
                const float A0 = bb * dl, B0 = bb * sqrtf((e - dl) * dl), C0 = bb * (e - dl);
                const float d1 = rx * g1x + rz * g1z, d2 = rx * g2x + rz * g2z;
                const float sP = (bb - A0) * d1 - B0 * d2, sM = -(B0 * d1 + C0 * d2);
                a.P[0][i] = (A0 * g1x + B0 * g2x) + sP * rx;
                a.P[1][i] = (A0 * g1z + B0 * g2z) + sP * rz;
                a.Mx[0][i] = (B0 * g1x + C0 * g2x) + sM * rx;
                a.Mx[1][i] = (B0 * g1z + C0 * g2z) + sM * rz;
            }
        }
        grid.sync();
        // ---- pass 2: divergence, update, injection, fused extras -------------------------------
        const bool fused_t = ((t - a.t_base) % a.t_sub) == 0;
        const long long sidx = (long long)((t - a.t_base) / a.t_sub) * NF * a.slot;
        float *snap_t = (a.snap && fused_t) ? a.snap + sidx : nullptr;
        const float *usnap_t = (a.grad && fused_t) ? a.usnap + sidx : nullptr;
        const float *src_row = a.src ? a.src + (long long)t * a.nsrc : nullptr;
        for (int tile = lb; tile < ntiles; tile += nb) {
            const int x = (tile % ntx) * TX + tx, z = (tile / ntx) * TZ + tz;
            if (x >= g.n[0] || z >= g.n[2]) continue;
            const long long i = g.at(x, 0, z);
            const float ir = a.irho[i];
            const float sd = cf.dt * ((a.dx[x] + 0.f) + a.dz[z]);
            const float rden = 1.0f / (a.m[i] * ir + sd);
            float inj = 0.f;
            int row = -1;
            if (a.cellmap && (row = a.cellmap[i]) >= 0) {
                float acc = 0.f;
                for (int e = a.cell_start[row]; e < a.cell_start[row + 1]; e++)
                    acc += a.ent_w[e] * src_row[a.ent_trace[e]];
                inj = acc * a.cell_scale[row];
            }
            float ucv[NF], delta[NF];
#pragma unroll
            for (int f = 0; f < NF; f++) {
                const float *Fx = f == 0 ? a.P[0] : a.Mx[0], *Fz = f == 0 ? a.P[1] : a.Mx[1];
                float L = 0.f, Lz = 0.f;
#pragma unroll
                for (int k = 1; k <= R; k++) {
                    L += cf.wx[k] * (Fx[i + (k - 1)] - Fx[i - k]);
                    Lz += cf.wz[k] * (Fz[i + (k - 1) * g.sz] - Fz[i - k * g.sz]);
                }
                L += Lz;
                const float *uc = a.b[f][cur];
                float *un = a.b[f][1 - cur];
                ucv[f] = uc[i];
                const float d1 = ucv[f] - un[i];
                delta[f] = (cf.dt2 * L - sd * d1) * rden;
                // point sources drive the first field only (geom_utils.py:81)
                const float add = f == 0 ? inj : 0.f;
                un[i] = ((ucv[f] + d1) + delta[f]) + add;
                delta[f] += add;     // post-injection u+ - 2u + u-
            }
            if ((snap_t || usnap_t || a.illum) && a.reg.has(x, 0, z)) {
                const long long ri = a.reg.at(x, 0, z);
                float e2 = 0.f, ic = 0.f;
#pragma unroll
                for (int f = 0; f < NF; f++) {
                    if (snap_t) snap_t[f * a.slot + ri] = ucv[f];
                    e2 += ucv[f] * ucv[f];
                    if (usnap_t) ic += usnap_t[f * a.slot + ri] * delta[f];
                }
                if (a.illum) a.illum[ri] += cf.dt * e2;
                if (usnap_t) a.grad[ri] -= a.gcoef * ir * ic;
            }
        }
        grid.sync();
        cur = 1 - cur;
        if (t == a.t1) break;
    }
}

template <int R, bool TTI>
__global__ void __launch_bounds__(256) flux2d_persist(P2FluxArgs a)
{
    flux2d_body<R, TTI>(a, blockIdx.x, gridDim.x);
}

// several shots in lockstep (see acoustic2d_persist_batch)
template <int R, bool TTI>
__global__ void __launch_bounds__(256) flux2d_persist_batch(const P2FluxArgs *__restrict__ batch, int per_shot)
{
    __shared__ P2FluxArgs sa;
    {
        const int *src = (const int *)(batch + blockIdx.x / per_shot);
        int *dst = (int *)&sa;
        for (int i = threadIdx.x; i < (int)(sizeof(P2FluxArgs) / sizeof(int)); i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    flux2d_body<R, TTI>(sa, blockIdx.x % per_shot, per_shot);
}

// ------------------------------------------------------------------------------------------
// The persistent 2-D time loop of the visco-acoustic SLS physics (kernels.cu has the equations):
// forward  [flux of p] -> barrier -> r, p update ;  adjoint  r-, w -> barrier -> [flux of w] ->
// barrier -> p update.  The flux pass exists only with a density field.
struct P2ViscoArgs {
    Grid g;
    Coefs cf;
    float *b[2];          // pressure buffers; level t is b[cur]
    float *r, *w;         // memory variable, adjoint product field
    float *P[2];          // flux scratch (density field)
    int cur;
    const float *m, *irho, *tt, *ts;
    float irho_c;
    const float *dx, *dz;
    int t0, t1, dir, adj;
    const int *cellmap;
    const int *cell_start, *ent_trace;
    const float *ent_w, *cell_scale;
    const float *src;
    int nsrc;
    int nrec;
    const long long *coff;
    const float *cw;
    float *rec;
    float *snap;
    const float *usnap;
    float *grad, *illum;
    long long slot;
    int t_sub;
    float gcoef;
    int t_base;
    Region reg;
};

template <int R>
__device__ __forceinline__ void visco2d_body(const P2ViscoArgs &a, const int lb, const int nb)
{
    cg::grid_group grid = cg::this_grid();
    const Grid &g = a.g;
    const Coefs &cf = a.cf;
    const int TX = 32, TZ = 8;
    const int tx = threadIdx.x % TX, tz = threadIdx.x / TX;
    const long long gtid = (long long)lb * blockDim.x + threadIdx.x;
    const long long gthreads = (long long)nb * blockDim.x;
    const int fnx = g.n[0] + 2 * R, fnz = g.n[2] + 2 * R;
    const int ftx = (fnx + TX - 1) / TX, ftiles = ftx * ((fnz + TZ - 1) / TZ);
    const int ntx = (g.n[0] + TX - 1) / TX, ntiles = ntx * ((g.n[2] + TZ - 1) / TZ);
    int cur = a.cur;
    for (int t = a.t0;; t += a.dir) {
        const float *pcur = a.b[cur];
        float *pnew = a.b[1 - cur];
        for (long long p = gtid; p < a.nrec; p += gthreads) {
            float acc = 0.f;
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const long long o = a.coff[p * 4 + c];
                if (o >= 0) acc += a.cw[p * 4 + c] * pcur[o];
            }
            a.rec[(long long)t * a.nrec + p] = acc;
        }
        if (a.adj) {
            // r- = (1 - dl) (r - dt (b p + r / ts)),  w = (1 + tt) p + (tt / (b ts)) r-
            for (int tile = lb; tile < ntiles; tile += nb) {
                const int x = (tile % ntx) * TX + tx, z = (tile / ntx) * TZ + tz;
                if (x >= g.n[0] || z >= g.n[2]) continue;
                const long long i = g.at(x, 0, z);
                const float bb = a.irho ? a.irho[i] : a.irho_c;
                const float dl = (a.dx[x] + 0.f) + a.dz[z];
                const float tt = a.tt[i], ts = a.ts[i], pc = pcur[i], rc = a.r[i];
                const float rn = (1.0f - dl) * (rc - cf.dt * (bb * pc + rc / ts));
                a.r[i] = rn;
                a.w[i] = (1.0f + tt) * pc + (tt / (bb * ts)) * rn;
            }
            grid.sync();
        }
        const float *Ls = a.adj ? a.w : pcur;
        if (a.irho) {
            for (int tile = lb; tile < ftiles; tile += nb) {
                const int x = (tile % ftx) * TX + tx - R, z = (tile / ftx) * TZ + tz - R;
                if (x >= g.n[0] + R || z >= g.n[2] + R) continue;
                const long long i = g.at(x, 0, z);
                float gx = 0.f, gz = 0.f;
#pragma unroll
                for (int k = 1; k <= R; k++) {
                    gx += cf.wx[k] * (Ls[i + k] - Ls[i - (k - 1)]);
                    gz += cf.wz[k] * (Ls[i + k * g.sz] - Ls[i - (k - 1) * g.sz]);
                }
                const float bb = a.irho[i];
                a.P[0][i] = bb * gx;
                a.P[1][i] = bb * gz;
            }
            grid.sync();
        }
        const bool fused_t = ((t - a.t_base) % a.t_sub) == 0;
        const long long sidx = (long long)((t - a.t_base) / a.t_sub) * a.slot;
        float *snap_t = (a.snap && fused_t) ? a.snap + sidx : nullptr;
        const float *usnap_t = (a.grad && fused_t) ? a.usnap + sidx : nullptr;
        const float *src_row = a.src ? a.src + (long long)t * a.nsrc : nullptr;
        for (int tile = lb; tile < ntiles; tile += nb) {
            const int x = (tile % ntx) * TX + tx, z = (tile / ntx) * TZ + tz;
            if (x >= g.n[0] || z >= g.n[2]) continue;
            const long long i = g.at(x, 0, z);
            const float bb = a.irho ? a.irho[i] : a.irho_c;
            const float mb = a.m[i] * bb;
            const float dl = (a.dx[x] + 0.f) + a.dz[z];
            const float sd = cf.dt * dl;
            float L2;
            if (a.irho) {
                float L = 0.f, Lz = 0.f;
#pragma unroll
                for (int k = 1; k <= R; k++) {
                    L += cf.wx[k] * (a.P[0][i + (k - 1)] - a.P[0][i - k]);
                    Lz += cf.wz[k] * (a.P[1][i + (k - 1) * g.sz] - a.P[1][i - k * g.sz]);
                }
                L2 = cf.dt2 * (L + Lz);
            } else L2 = bb * lap2d<R>(g, cf, Ls, i);
            float S2 = L2;
            if (!a.adj) {
                const float tt = a.tt[i], ts = a.ts[i], rc = a.r[i];
                const float rn = (1.0f - dl) * (rc + ((tt / ts) * (L2 / bb) * cf.idt - cf.dt * (rc / ts)));
                a.r[i] = rn;
                S2 = (1.0f + tt) * L2 - cf.dt2 * (bb * rn);
            }
            float inj = 0.f;
            int row = -1;
            if (a.cellmap && (row = a.cellmap[i]) >= 0) {
                float acc = 0.f;
                for (int e = a.cell_start[row]; e < a.cell_start[row + 1]; e++)
                    acc += a.ent_w[e] * src_row[a.ent_trace[e]];
                inj = acc * a.cell_scale[row];
            }
            const float pc = pcur[i];
            const float d1 = pc - pnew[i];
            const float delta = (S2 - sd * d1) / (mb + sd);
            pnew[i] = ((pc + d1) + delta) + inj;
            if ((snap_t || usnap_t || a.illum) && a.reg.has(x, 0, z)) {
                const long long ri = a.reg.at(x, 0, z);
                if (snap_t) snap_t[ri] = pc;
                if (a.illum) a.illum[ri] += cf.dt * pc * pc;
                if (usnap_t) a.grad[ri] -= a.gcoef * bb * usnap_t[ri] * (delta + inj);
            }
        }
        grid.sync();
        cur = 1 - cur;
        if (t == a.t1) break;
    }
}

template <int R>
__global__ void __launch_bounds__(256) visco2d_persist(P2ViscoArgs a)
{
    visco2d_body<R>(a, blockIdx.x, gridDim.x);
}

template <int R>
__global__ void __launch_bounds__(256) visco2d_persist_batch(const P2ViscoArgs *__restrict__ batch, int per_shot)
{
    __shared__ P2ViscoArgs sa;
    {
        const int *src = (const int *)(batch + blockIdx.x / per_shot);
        int *dst = (int *)&sa;
        for (int i = threadIdx.x; i < (int)(sizeof(P2ViscoArgs) / sizeof(int)); i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    visco2d_body<R>(sa, blockIdx.x % per_shot, per_shot);
}

// ------------------------------------------------------------------------------------------
bool persist2d_supported(const judi_model *M, int ic, bool born)
{
    const bool flux = M->tti || M->irho_field || M->visco;
    // flux form / visco: so = 8, no Born, "as" only; constant density: every imaging condition
    if (flux && (born || M->fs || M->r != 4 || ic != JUDI_IC_AS || (M->tti && !M->r3[0].p))) return false;
    return M->ndim == 2 && (M->r == 2 || M->r == 4 || M->r == 8) && M->ctx->opt_persist2d;
}

static void p2flux_fill(const judi_model *M, Wavefield &W, int t0, int t1, int dir,
                        const SparseSet *S_in, const Traces *in, const SparseSet *S_out,
                        Traces *out, const P2Fuse &fz, P2FluxArgs &a)
{
    memset(&a, 0, sizeof(a));
    a.g = M->g; a.cf = M->cf;
    const int nf = M->tti ? 2 : 1;
    for (int f = 0; f < nf; f++) { a.b[f][0] = W.b[f][0].p; a.b[f][1] = W.b[f][1].p; }
    a.P[0] = W.flux[0].p; a.P[1] = W.flux[2].p;
    if (M->tti) { a.Mx[0] = W.flux[3].p; a.Mx[1] = W.flux[5].p; }
    a.cur = W.cur;
    a.m = M->m.p; a.irho = M->irho.p; a.eps = M->eps.p; a.delta = M->delta.p;
    a.st = M->r3[0].p; a.ct = M->r3[2].p;
    a.dx = M->damp[0].p; a.dz = M->damp[2].p;
    a.t0 = t0; a.t1 = t1; a.dir = dir;
    if (S_in && in && S_in->ncell > 0) {
        a.cellmap = S_in->cellmap.p;
        a.cell_start = S_in->cell_start.p;
        a.ent_trace = S_in->ent_trace.p;
        a.ent_w = S_in->ent_w.p;
        a.cell_scale = S_in->cell_scale.p;
        a.src = in->buf.p;
        a.nsrc = in->np;
    }
    if (S_out && out && S_out->np > 0) {
        a.nrec = S_out->np;
        a.coff = S_out->corner_off.p;
        a.cw = S_out->corner_w.p;
        a.rec = out->buf.p;
    }
    a.snap = fz.snap; a.usnap = fz.usnap; a.grad = fz.grad; a.illum = fz.illum;
    a.slot = fz.slot / nf;   // P2Fuse::slot counts all fields of a snapshot
    a.t_sub = fz.t_sub > 0 ? fz.t_sub : 1; a.gcoef = fz.gcoef;
    a.reg = fz.reg;
    a.t_base = fz.t_base;
}

static void persist2d_run_flux(const judi_model *M, Wavefield &W, int t0, int t1, int dir,
                               const SparseSet *S_in, const Traces *in, const SparseSet *S_out,
                               Traces *out, const P2Fuse &fz)
{
    judi_ctx *ctx = M->ctx;
    const int nsteps = (t1 - t0) * dir + 1;
    P2FluxArgs a;
    p2flux_fill(M, W, t0, t1, dir, S_in, in, S_out, out, fz, a);
    void *kern = M->tti ? (void *)flux2d_persist<4, true> : (void *)flux2d_persist<4, false>;
    static int max_blocks[2] = {0, 0};
    int &mb = max_blocks[M->tti ? 1 : 0];
    if (!mb) {
        int per_sm = 0;
        if (M->tti) CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, flux2d_persist<4, true>, 256, 0));
        else CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, flux2d_persist<4, false>, 256, 0));
        mb = per_sm * ctx->num_sms;
    }
    const int ftiles = ((M->N[0] + 8 + 31) / 32) * ((M->N[2] + 8 + 7) / 8);
    int nb = std::min(ftiles, mb);
    if (ctx->opt_persist_blocks > 0) nb = std::min(nb, ctx->opt_persist_blocks);
    void *args[] = {(void *)&a};
    CUDA_CHECK(cudaLaunchCooperativeKernel(kern, dim3(nb), dim3(256), args, 0, ctx->stream));
    launch_count(ctx, "flux2d_persist");
    if (nsteps & 1) W.swap();
    if (ctx->timing) {
        ctx->stencil_launches += 1;
        ctx->stencil_points += (double)nsteps * M->N[0] * M->N[2];
    }
}

static void p2visco_fill(const judi_model *M, Wavefield &W, int t0, int t1, int dir,
                         const SparseSet *S_in, const Traces *in, const SparseSet *S_out,
                         Traces *out, const P2Fuse &fz, P2ViscoArgs &a)
{
    memset(&a, 0, sizeof(a));
    a.g = M->g; a.cf = M->cf;
    a.b[0] = W.b[0][0].p; a.b[1] = W.b[0][1].p;
    a.r = W.rmem.p; a.w = W.wtmp.p;
    if (M->irho_field) { a.P[0] = W.flux[0].p; a.P[1] = W.flux[2].p; a.irho = M->irho.p; }
    a.irho_c = M->irho_c;
    a.cur = W.cur;
    a.m = M->m.p; a.tt = M->vtt.p; a.ts = M->vts.p;
    a.dx = M->damp[0].p; a.dz = M->damp[2].p;
    a.t0 = t0; a.t1 = t1; a.dir = dir;
    a.adj = dir < 0;      // the reverse sweep is the adjoint scheme
    if (S_in && in && S_in->ncell > 0) {
        a.cellmap = S_in->cellmap.p;
        a.cell_start = S_in->cell_start.p;
        a.ent_trace = S_in->ent_trace.p;
        a.ent_w = S_in->ent_w.p;
        a.cell_scale = S_in->cell_scale.p;
        a.src = in->buf.p;
        a.nsrc = in->np;
    }
    if (S_out && out && S_out->np > 0) {
        a.nrec = S_out->np;
        a.coff = S_out->corner_off.p;
        a.cw = S_out->corner_w.p;
        a.rec = out->buf.p;
    }
    a.snap = fz.snap; a.usnap = fz.usnap; a.grad = fz.grad; a.illum = fz.illum;
    a.slot = fz.slot;
    a.t_sub = fz.t_sub > 0 ? fz.t_sub : 1; a.gcoef = fz.gcoef;
    a.reg = fz.reg;
    a.t_base = fz.t_base;
}

static int p2visco_blocks(const judi_model *M, bool batch)
{
    static int mb[2] = {0, 0};
    if (!mb[batch]) {
        int per_sm = 0;
        if (batch) CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, visco2d_persist_batch<4>, 256, 0));
        else CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, visco2d_persist<4>, 256, 0));
        mb[batch] = per_sm * M->ctx->num_sms;
    }
    return mb[batch];
}

static void persist2d_run_visco(const judi_model *M, Wavefield &W, int t0, int t1, int dir,
                                const SparseSet *S_in, const Traces *in, const SparseSet *S_out,
                                Traces *out, const P2Fuse &fz)
{
    judi_ctx *ctx = M->ctx;
    const int nsteps = (t1 - t0) * dir + 1;
    P2ViscoArgs a;
    p2visco_fill(M, W, t0, t1, dir, S_in, in, S_out, out, fz, a);
    const int ftiles = ((M->N[0] + 8 + 31) / 32) * ((M->N[2] + 8 + 7) / 8);
    int nb = std::min(ftiles, p2visco_blocks(M, false));
    if (ctx->opt_persist_blocks > 0) nb = std::min(nb, ctx->opt_persist_blocks);
    void *args[] = {(void *)&a};
    CUDA_CHECK(cudaLaunchCooperativeKernel((void *)visco2d_persist<4>, dim3(nb), dim3(256), args, 0, ctx->stream));
    launch_count(ctx, "visco2d_persist");
    if (nsteps & 1) W.swap();
    if (ctx->timing) {
        ctx->stencil_launches += 1;
        ctx->stencil_points += (double)nsteps * M->N[0] * M->N[2];
    }
}

static void p2_fill(const judi_model *M, Wavefield &W, Wavefield *WL, int t0, int t1, int dir,
                    const SparseSet *S_in, const Traces *in, const SparseSet *S_out, Traces *out,
                    Traces *out_bg, const P2Fuse &fz, P2Args &a)
{
    memset(&a, 0, sizeof(a));
    a.g = M->g;
    a.cf = M->cf;
    a.b[0] = W.b[0][0].p; a.b[1] = W.b[0][1].p;
    if (WL) { a.l[0] = WL->b[0][0].p; a.l[1] = WL->b[0][1].p; }
    a.cur = W.cur;
    a.m = M->m.p; a.dx = M->damp[0].p; a.dz = M->damp[2].p; a.dm = M->dm.p; a.dms = M->dm_src();
    a.sdamp = M->dt / M->irho_c;
    a.irho_c = M->irho_c;
    a.t0 = t0; a.t1 = t1; a.dir = dir;
    if (S_in && in && S_in->ncell > 0) {
        a.cellmap = S_in->cellmap.p;
        a.cell_start = S_in->cell_start.p;
        a.ent_trace = S_in->ent_trace.p;
        a.ent_w = S_in->ent_w.p;
        a.cell_scale = S_in->cell_scale.p;
        a.cell_born = S_in->cell_born.p;
        a.src = in->buf.p;
        a.nsrc = in->np;
    }
    if (S_out && out && S_out->np > 0) {
        a.nrec = S_out->np;
        a.coff = S_out->corner_off.p;
        a.cw = S_out->corner_w.p;
        a.rec = out->buf.p;
        a.rec_bg = out_bg ? out_bg->buf.p : nullptr;
    }
    a.snap = fz.snap; a.usnap = fz.usnap; a.grad = fz.grad; a.illum = fz.illum;
    a.slot = fz.slot; a.t_sub = fz.t_sub > 0 ? fz.t_sub : 1; a.gcoef = fz.gcoef;
    a.reg = fz.reg;
    a.t_base = fz.t_base;
    a.ic = fz.ic;
    a.qext = fz.qext; a.qwt = fz.qwt; a.qwf = fz.qwf; a.qbase = fz.qbase;
    a.wrec = fz.wrec; a.wrt = fz.wrt; a.wdt = fz.wdt;
    a.fs = M->fs ? 1 : 0;
}

// what the 4-points-per-thread body covers
static bool p2_vec_ok(const P2Args &a, bool born = false)
{
    return !a.fs && !a.qext && !a.qwf && !a.wrec && (born || !a.l[0]) && a.ic == JUDI_IC_AS;
}

// Runs time indices t0, t0+dir, ..., t1 of one sweep.  W (and WL for Born) are advanced
// accordingly.  S_in/in: injected traces; S_out/out: interpolated traces (of WL when Born, with
// out_bg sampling W for nlind).
void persist2d_run(const judi_model *M, Wavefield &W, Wavefield *WL, int t0, int t1, int dir,
                   const SparseSet *S_in, const Traces *in, const SparseSet *S_out, Traces *out,
                   Traces *out_bg, const P2Fuse &fz)
{
    judi_ctx *ctx = M->ctx;
    int nsteps = (t1 - t0) * dir + 1;
    if (nsteps <= 0) return;
    if (M->visco) {
        JUDI_REQUIRE(WL == nullptr && out_bg == nullptr, "persistent 2-D visco-acoustic kernel has no Born mode");
        persist2d_run_visco(M, W, t0, t1, dir, S_in, in, S_out, out, fz);
        return;
    }
    if (M->tti || M->irho_field) {
        JUDI_REQUIRE(WL == nullptr && out_bg == nullptr, "persistent 2-D flux kernel has no Born mode");
        persist2d_run_flux(M, W, t0, t1, dir, S_in, in, S_out, out, fz);
        return;
    }
    P2Args a;
    p2_fill(M, W, WL, t0, t1, dir, S_in, in, S_out, out, out_bg, fz, a);
    auto launch = [&](auto rtag, auto xtag, auto vtag) {
        constexpr int R = decltype(rtag)::value;
        constexpr int VEC = decltype(vtag)::value;
        auto kern = acoustic2d_persist<R, decltype(xtag)::value, VEC>;
        static int max_blocks = 0;
        if (!max_blocks) {
            int per_sm = 0;
            CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, 0));
            max_blocks = per_sm * ctx->num_sms;
        }
        int ntiles = ((M->N[0] + (VEC ? 127 : 31)) / (VEC ? 128 : 32)) * ((M->N[2] + 7) / 8);
        int nb = std::min(ntiles, max_blocks);
        if (ctx->opt_persist_blocks > 0) nb = std::min(nb, ctx->opt_persist_blocks);
        void *args[] = {(void *)&a};
        CUDA_CHECK(cudaLaunchCooperativeKernel((void *)kern, dim3(nb), dim3(256), args, 0,
                                               ctx->stream));
        launch_count(ctx, "acoustic2d_persist");
    };
    auto launch_r = [&](auto xtag, auto vtag) {
        switch (M->r) {
        case 2: launch(std::integral_constant<int, 2>(), xtag, vtag); break;
        case 4: launch(std::integral_constant<int, 4>(), xtag, vtag); break;
        default: launch(std::integral_constant<int, 8>(), xtag, vtag); break;
        }
    };
    typedef std::integral_constant<int, 0> V0;
    if (a.ic != JUDI_IC_AS && (a.grad || WL)) launch_r(std::true_type(), V0());
    else if (ctx->opt_persist_vec == 2 && p2_vec_ok(a)) launch_r(std::false_type(), std::integral_constant<int, 2>());
    else launch_r(std::false_type(), V0());
    if (nsteps & 1) { W.swap(); if (WL) WL->swap(); }
    if (ctx->timing) {
        ctx->stencil_launches += 1;
        ctx->stencil_points += (double)nsteps * M->N[0] * M->N[2] * (WL ? 2 : 1);
    }
}

// ---- several shots in one cooperative launch -------------------------------------------------------
bool persist2d_batch_supported(const judi_model *M, int ic)
{
    return persist2d_supported(M, ic, false);
}

// One sweep (t0, t0+dir, ..., t1) of `n` shots of the same model in lockstep.  W[s] is advanced.
void persist2d_run_batch(const judi_model *M, int n, Wavefield *const *W, int t0, int t1, int dir,
                         const SparseSet *const *S_in, const Traces *const *in,
                         const SparseSet *const *S_out, Traces *const *out, const P2Fuse *fz,
                         Wavefield *const *WL, Traces *const *out_bg)
{
    judi_ctx *ctx = M->ctx;
    const int nsteps = (t1 - t0) * dir + 1;
    if (nsteps <= 0 || n <= 0) return;
    if (M->visco) {
        JUDI_REQUIRE(WL == nullptr, "persistent 2-D visco-acoustic kernel has no Born mode");
        std::vector<P2ViscoArgs> vargs((size_t)n);
        for (int s = 0; s < n; s++)
            p2visco_fill(M, *W[s], t0, t1, dir, S_in ? S_in[s] : nullptr, in ? in[s] : nullptr,
                         S_out ? S_out[s] : nullptr, out ? out[s] : nullptr, fz[s], vargs[(size_t)s]);
        DevArr<P2ViscoArgs> dva(ctx, (size_t)n, false);
        dva.upload(vargs);
        const int ftiles = ((M->N[0] + 8 + 31) / 32) * ((M->N[2] + 8 + 7) / 8);
        int cap = p2visco_blocks(M, true);
        if (ctx->opt_persist_blocks > 0) cap = std::min(cap, ctx->opt_persist_blocks);
        int per_shot = std::max(1, std::min(ftiles, cap / n));
        const P2ViscoArgs *dp = dva.p;
        void *kargs[] = {(void *)&dp, (void *)&per_shot};
        CUDA_CHECK(cudaLaunchCooperativeKernel((void *)visco2d_persist_batch<4>, dim3(per_shot * n), dim3(256),
                                               kargs, 0, ctx->stream));
        launch_count(ctx, "visco2d_persist_batch");
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        if (nsteps & 1)
            for (int s = 0; s < n; s++) W[s]->swap();
        if (ctx->timing) {
            ctx->stencil_launches += 1;
            ctx->stencil_points += (double)n * nsteps * M->N[0] * M->N[2];
        }
        return;
    }
    if (M->tti || M->irho_field) {
        JUDI_REQUIRE(WL == nullptr, "persistent 2-D flux kernel has no Born mode");
        std::vector<P2FluxArgs> fargs((size_t)n);
        for (int s = 0; s < n; s++)
            p2flux_fill(M, *W[s], t0, t1, dir, S_in ? S_in[s] : nullptr, in ? in[s] : nullptr,
                        S_out ? S_out[s] : nullptr, out ? out[s] : nullptr, fz[s], fargs[(size_t)s]);
        DevArr<P2FluxArgs> dfa(ctx, (size_t)n, false);
        dfa.upload(fargs);
        void *kern = M->tti ? (void *)flux2d_persist_batch<4, true> : (void *)flux2d_persist_batch<4, false>;
        static int max_blocks[2] = {0, 0};
        int &mb = max_blocks[M->tti ? 1 : 0];
        if (!mb) {
            int per_sm = 0;
            if (M->tti) CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, flux2d_persist_batch<4, true>, 256, 0));
            else CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, flux2d_persist_batch<4, false>, 256, 0));
            mb = per_sm * ctx->num_sms;
        }
        const int ftiles = ((M->N[0] + 8 + 31) / 32) * ((M->N[2] + 8 + 7) / 8);
        int cap = mb;
        if (ctx->opt_persist_blocks > 0) cap = std::min(cap, ctx->opt_persist_blocks);
        int per_shot = std::max(1, std::min(ftiles, cap / n));
        const P2FluxArgs *dp = dfa.p;
        void *kargs[] = {(void *)&dp, (void *)&per_shot};
        CUDA_CHECK(cudaLaunchCooperativeKernel(kern, dim3(per_shot * n), dim3(256), kargs, 0, ctx->stream));
        launch_count(ctx, "flux2d_persist_batch");
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        if (nsteps & 1)
            for (int s = 0; s < n; s++) W[s]->swap();
        if (ctx->timing) {
            ctx->stencil_launches += 1;
            ctx->stencil_points += (double)n * nsteps * M->N[0] * M->N[2];
        }
        return;
    }
    std::vector<P2Args> args((size_t)n);
    for (int s = 0; s < n; s++)
        p2_fill(M, *W[s], WL ? WL[s] : nullptr, t0, t1, dir, S_in ? S_in[s] : nullptr, in ? in[s] : nullptr,
                S_out ? S_out[s] : nullptr, out ? out[s] : nullptr, out_bg ? out_bg[s] : nullptr, fz[s],
                args[(size_t)s]);
    DevArr<P2Args> dargs(ctx, (size_t)n, false);
    dargs.upload(args);
    auto launch = [&](auto rtag, auto xtag, auto vtag, auto btag) {
        constexpr int R = decltype(rtag)::value;
        constexpr int VEC = decltype(vtag)::value;
        constexpr bool VBORN = decltype(btag)::value;
        auto kern = acoustic2d_persist_batch<R, decltype(xtag)::value, VEC, VBORN>;
        static int max_blocks = 0;
        if (!max_blocks) {
            int per_sm = 0;
            CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, 0));
            max_blocks = per_sm * ctx->num_sms;
        }
        const int ntiles = ((M->N[0] + (VEC ? 127 : 31)) / (VEC ? 128 : 32)) * ((M->N[2] + 7) / 8);
        int cap = max_blocks;
        if (ctx->opt_persist_blocks > 0) cap = std::min(cap, ctx->opt_persist_blocks);
        int per_shot = std::max(1, std::min(ntiles, cap / n));
        // as few CTAs as the same number of tile rounds needs (whole rounds, no idle tail)
        if (VEC) per_shot = (ntiles + (ntiles + per_shot - 1) / per_shot - 1) / ((ntiles + per_shot - 1) / per_shot);
        const P2Args *dp = dargs.p;
        void *kargs[] = {(void *)&dp, (void *)&per_shot};
        CUDA_CHECK(cudaLaunchCooperativeKernel((void *)kern, dim3(per_shot * n), dim3(256), kargs, 0,
                                               ctx->stream));
        launch_count(ctx, "acoustic2d_persist_batch");
    };
    auto launch_r = [&](auto xtag, auto vtag, auto btag) {
        switch (M->r) {
        case 2: launch(std::integral_constant<int, 2>(), xtag, vtag, btag); break;
        case 4: launch(std::integral_constant<int, 4>(), xtag, vtag, btag); break;
        default: launch(std::integral_constant<int, 8>(), xtag, vtag, btag); break;
        }
    };
    bool vec = ctx->opt_persist_vec != 0;
    for (int s = 0; s < n; s++) vec = vec && p2_vec_ok(args[(size_t)s], WL != nullptr);
    // option persist_vec: 1 / 2 = vector body at 3 CTAs per SM (80 registers), 3 = at 2 CTAs per SM (118),
    // 4 = at 4 CTAs per SM (64 registers, 84 bytes of spills)
    typedef std::integral_constant<int, 0> V0;
    if (args[0].ic != JUDI_IC_AS && (args[0].grad || WL)) launch_r(std::true_type(), V0(), std::false_type());
    else if (vec && WL) launch_r(std::false_type(), std::integral_constant<int, 3>(), std::true_type());
    else if (vec && ctx->opt_persist_vec == 3) launch_r(std::false_type(), std::integral_constant<int, 2>(), std::false_type());
    else if (vec && ctx->opt_persist_vec == 4) launch_r(std::false_type(), std::integral_constant<int, 4>(), std::false_type());
    else if (vec) launch_r(std::false_type(), std::integral_constant<int, 3>(), std::false_type());
    else launch_r(std::false_type(), V0(), std::false_type());
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));   // dargs is freed on return
    if (nsteps & 1)
        for (int s = 0; s < n; s++) {
            W[s]->swap();
            if (WL) WL[s]->swap();
        }
    if (ctx->timing) {
        ctx->stencil_launches += 1;
        ctx->stencil_points += (double)n * nsteps * M->N[0] * M->N[2];
    }
}
