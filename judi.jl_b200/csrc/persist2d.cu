// persist2d.cu -- 2-D isotropic acoustic (constant density) time loop as ONE persistent
// cooperative kernel.
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
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

struct P2Args {
    Grid g;
    Coefs cf;
    float *b[2];       // background field buffers; level t is b[cur]
    float *l[2];       // linearised field buffers (Born) or nullptr
    int cur;
    const float *m, *dx, *dz, *dm;
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
    Region reg;
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

template <int R>
__global__ void __launch_bounds__(256) acoustic2d_persist(P2Args a)
{
    cg::grid_group grid = cg::this_grid();
    const Grid &g = a.g;
    const int TX = 32, TZ = 8;
    const int ntx = (g.n[0] + TX - 1) / TX, ntz = (g.n[2] + TZ - 1) / TZ;
    const int ntiles = ntx * ntz;
    const int tx = threadIdx.x % TX, tz = threadIdx.x / TX;
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long gthreads = (long long)gridDim.x * blockDim.x;
    int cur = a.cur;
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
        const bool fused_t = ((t - a.t_base) % a.t_sub) == 0;
        const long long sidx = (long long)((t - a.t_base) / a.t_sub) * a.slot;
        float *snap_t = (a.snap && fused_t) ? a.snap + sidx : nullptr;
        const float *usnap_t = (a.grad && fused_t) ? a.usnap + sidx : nullptr;
        const float *src_row = a.src ? a.src + (long long)t * a.nsrc : nullptr;
        // ---- stencil + injection + fused extras, tile by tile --------------------------------
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int x = (tile % ntx) * TX + tx, z = (tile / ntx) * TZ + tz;
            if (x >= g.n[0] || z >= g.n[2]) continue;
            const long long i = g.at(x, 0, z);
            const float ucv = uc[i];
            const float lap = lap2d<R>(g, a.cf, uc, i);
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
            un[i] = ((ucv + d1) + delta) + inj;
            if (ln) {
                const float lcv = lc[i];
                const float lapl = lap2d<R>(g, a.cf, lc, i);
                const float d1l = lcv - ln[i];
                // Born source from the post-injection u+ - 2u + u-  (sensitivity.py:174-188)
                const float q = -a.dm[i] * (delta + inj);
                ln[i] = (lcv + d1l) + (lapl - s * d1l + q) * rden;
            }
            if ((snap_t || usnap_t || a.illum) && a.reg.has(x, 0, z)) {
                const long long ri = a.reg.at(x, 0, z);
                if (snap_t) snap_t[ri] = ucv;
                if (a.illum) a.illum[ri] += a.cf.dt * ucv * ucv;
                if (usnap_t) a.grad[ri] -= a.gcoef * a.irho_c * usnap_t[ri] * (delta + inj);
            }
        }
        grid.sync();
        cur = 1 - cur;
        if (t == a.t1) break;
    }
}

// ------------------------------------------------------------------------------------------
bool persist2d_supported(const judi_model *M, int ic)
{
    return M->ndim == 2 && !M->tti && !M->visco && !M->irho_field && !M->fs && ic == JUDI_IC_AS &&
           (M->r == 2 || M->r == 4 || M->r == 8) && M->ctx->opt_persist2d;
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
    P2Args a;
    memset(&a, 0, sizeof(a));
    a.g = M->g;
    a.cf = M->cf;
    a.b[0] = W.b[0][0].p; a.b[1] = W.b[0][1].p;
    if (WL) { a.l[0] = WL->b[0][0].p; a.l[1] = WL->b[0][1].p; }
    a.cur = W.cur;
    a.m = M->m.p; a.dx = M->damp[0].p; a.dz = M->damp[2].p; a.dm = M->dm.p;
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
    auto launch = [&](auto rtag) {
        constexpr int R = decltype(rtag)::value;
        auto kern = acoustic2d_persist<R>;
        static int max_blocks = 0;
        if (!max_blocks) {
            int per_sm = 0;
            CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, 0));
            max_blocks = per_sm * ctx->num_sms;
        }
        int ntiles = ((M->N[0] + 31) / 32) * ((M->N[2] + 7) / 8);
        int nb = std::min(ntiles, max_blocks);
        void *args[] = {(void *)&a};
        CUDA_CHECK(cudaLaunchCooperativeKernel((void *)kern, dim3(nb), dim3(256), args, 0,
                                               ctx->stream));
        launch_count(ctx, "acoustic2d_persist");
    };
    switch (M->r) {
    case 2: launch(std::integral_constant<int, 2>()); break;
    case 4: launch(std::integral_constant<int, 4>()); break;
    default: launch(std::integral_constant<int, 8>()); break;
    }
    if (nsteps & 1) { W.swap(); if (WL) WL->swap(); }
    if (ctx->timing) {
        ctx->stencil_launches += 1;
        ctx->stencil_points += (double)nsteps * M->N[0] * M->N[2] * (WL ? 2 : 1);
    }
}
