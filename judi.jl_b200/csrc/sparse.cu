// sparse.cu -- source injection and receiver interpolation.
//
// Replaces Devito's SparseTimeFunction.inject / .interpolate as used by
// src/pysource/geom_utils.py:31-96 (injection `src*dt^2/m` :82, interpolation :94).
//
// Layout: traces are sorted once per shot by the grid cell they touch.  Injection is a
// segmented sum -- one thread per *unique* cell adds all (trace, corner) contributions in a fixed
// order and does a single non-atomic read-modify-write -- so it is deterministic and conflict
// free.  An atomicAdd path (one thread per (trace, corner)) is kept behind
// judi_ctx_set_option("atomic_inject", 1).  Interpolation is a gather, one thread per trace,
// written trace-fastest so that a time step writes one contiguous row.
#include "common.cuh"
#include <algorithm>
#include <cmath>
#include <numeric>

// Position arithmetic exactly as Devito generates it: fp32
//   pos = (-o + coord) / h ; i = (int)floor(pos) ; p = -floor(pos) + pos
//   w = prod_d (r_d * p_d + (1 - r_d) * (1 - p_d))
// `volatile` pins every intermediate to fp32 (no contraction, no excess precision).
void sparse_locate_host(const judi_model *M, int np, const float *coords, std::vector<int> &ijk,
                        std::vector<float> &w, std::vector<int> &valid)
{
    int nd = M->ndim, nc = 1 << nd;
    ijk.assign((size_t)np * nc * 3, 0);
    w.assign((size_t)np * nc, 0.f);
    valid.assign((size_t)np * nc, 0);
    const int dm2[2] = {0, 2};
    for (int p = 0; p < np; p++) {
        int base[3] = {0, 0, 0};
        float frac[3] = {0, 0, 0};
        for (int d = 0; d < nd; d++) {
            int gd = nd == 3 ? d : dm2[d];
            volatile float num = -M->opml[gd] + coords[(size_t)d * np + p];
            volatile float pos = num / M->d[gd];
            float fl = floorf(pos);
            base[gd] = (int)fl;
            volatile float fr = -fl + pos;
            frac[gd] = fr;
        }
        for (int c = 0; c < nc; c++) {
            int rr[3] = {0, 0, 0};
            if (nd == 3) { rr[0] = (c >> 2) & 1; rr[1] = (c >> 1) & 1; rr[2] = c & 1; }
            else { rr[0] = (c >> 1) & 1; rr[2] = c & 1; }
            volatile float ww = 1.0f;
            for (int d = 0; d < nd; d++) {
                int gd = nd == 3 ? d : dm2[d];
                volatile float a = (float)rr[gd] * frac[gd];
                volatile float om = 1.0f - frac[gd];
                volatile float b = (float)(1 - rr[gd]) * om;
                volatile float wd = a + b;
                ww = ww * wd;
            }
            size_t k = (size_t)p * nc + c;
            int ok = 1;
            for (int gd = 0; gd < 3; gd++) {
                int v = base[gd] + rr[gd];
                ijk[3 * k + gd] = v;
                if (v < 0 || v >= M->N[gd]) ok = 0;
            }
            w[k] = ww;
            valid[k] = ok;
        }
    }
}

// per-unique-cell factors: scale = dt^2/(m*irho), born = -dm*irho/(m*irho + dt*damp)
__global__ void sparse_cell_setup(int ncell, const long long *__restrict__ off,
                                  const int *__restrict__ ijk, const float *__restrict__ m,
                                  const float *__restrict__ irho, float irho_c,
                                  const float *__restrict__ dm, const float *__restrict__ dx,
                                  const float *__restrict__ dy, const float *__restrict__ dz,
                                  float dt, float *__restrict__ scale, float *__restrict__ born)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ncell) return;
    long long o = off[i];
    float ir = irho ? irho[o] : irho_c;
    float mi = m[o] * ir;
    scale[i] = dt * dt / mi;
    if (born) {
        float damp = (dx[ijk[3 * i]] + dy[ijk[3 * i + 1]]) + dz[ijk[3 * i + 2]];
        born[i] = dm ? -dm[o] * ir / (mi + dt * damp) : 0.0f;
    }
}

__global__ void cellmap_kernel(int ncell, const long long *__restrict__ off, int *map)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < ncell) map[off[i]] = i;
}

void sparse_build(const judi_model *M, int np, const float *coords, SparseSet &S)
{
    judi_ctx *ctx = M->ctx;
    int nc = 1 << M->ndim;
    std::vector<int> ijk, valid;
    std::vector<float> w;
    sparse_locate_host(M, np, coords, ijk, w, valid);
    S.np = np;
    S.nc = nc;
    size_t ntot = (size_t)np * nc;
    std::vector<long long> coff(ntot);
    for (size_t k = 0; k < ntot; k++)
        coff[k] = valid[k] ? M->g.at(ijk[3 * k], ijk[3 * k + 1], ijk[3 * k + 2]) : -1;
    // sort valid entries by (cell offset, trace, corner): stable sort of the natural order
    std::vector<int> order;
    order.reserve(ntot);
    for (size_t k = 0; k < ntot; k++)
        if (valid[k]) order.push_back((int)k);
    std::stable_sort(order.begin(), order.end(),
                     [&](int a, int b) { return coff[a] < coff[b]; });
    std::vector<long long> cell_off;
    std::vector<int> cell_start, cell_ijk, ent_trace;
    std::vector<float> ent_w;
    ent_trace.reserve(order.size());
    ent_w.reserve(order.size());
    for (size_t e = 0; e < order.size(); e++) {
        int k = order[e];
        if (cell_off.empty() || cell_off.back() != coff[k]) {
            cell_off.push_back(coff[k]);
            cell_start.push_back((int)e);
            cell_ijk.push_back(ijk[3 * k]);
            cell_ijk.push_back(ijk[3 * k + 1]);
            cell_ijk.push_back(ijk[3 * k + 2]);
        }
        ent_trace.push_back(k / nc);
        ent_w.push_back(w[k]);
    }
    cell_start.push_back((int)order.size());
    S.ncell = (int)cell_off.size();
    S.nent = (int)order.size();

    S.corner_off = DevArr<long long>(ctx, ntot, false);
    S.corner_off.upload(coff);
    S.corner_w = DevBuf(ctx, ntot, false);
    S.corner_w.upload(w);
    if (S.ncell > 0) {
        S.cell_off = DevArr<long long>(ctx, S.ncell, false);
        S.cell_off.upload(cell_off);
        S.cell_ijk = DevArr<int>(ctx, (size_t)S.ncell * 3, false);
        S.cell_ijk.upload(cell_ijk);
        S.ent_trace = DevArr<int>(ctx, S.nent, false);
        S.ent_trace.upload(ent_trace);
        S.ent_w = DevBuf(ctx, S.nent, false);
        S.ent_w.upload(ent_w);
    }
    S.cell_start = DevArr<int>(ctx, (size_t)S.ncell + 1, false);
    S.cell_start.upload(cell_start);
    S.cell_scale = DevBuf(ctx, std::max(S.ncell, 1), false);
    S.cell_born = DevBuf(ctx, std::max(S.ncell, 1), true);
    if (S.ncell > 0) {
        sparse_cell_setup<<<(S.ncell + 127) / 128, 128, 0, ctx->stream>>>(
            S.ncell, S.cell_off.p, S.cell_ijk.p, M->m.p, M->irho.p, M->irho_c,
            M->has_dm ? (M->visco ? M->dm.p : M->dm_src()) : nullptr, M->damp[0].p, M->damp[1].p, M->damp[2].p, M->dt,
            S.cell_scale.p, S.cell_born.p);
        launch_count(ctx, "sparse_cell_setup");
    }
    if (M->ndim == 2 && ctx->opt_persist2d) {
        // point -> CSR row map for the owner-computes injection of the persistent 2-D kernel
        S.cellmap = DevArr<int>(ctx, (size_t)M->g.nalloc, false);
        CUDA_CHECK(cudaMemsetAsync(S.cellmap.p, 0xFF, (size_t)M->g.nalloc * sizeof(int), ctx->stream));
        if (S.ncell > 0) {
            cellmap_kernel<<<(S.ncell + 127) / 128, 128, 0, ctx->stream>>>(S.ncell, S.cell_off.p,
                                                                            S.cellmap.p);
            launch_count(ctx, "cellmap");
        }
    }
}

// ------------------------------------------------------------------------------------------
struct SparseArgs {
    // injection (segmented)
    int ncell;
    const long long *cell_off;
    const int *cell_start;
    const int *cell_ijk;
    const int *ent_trace;
    const float *ent_w;
    const float *cell_scale;
    const float *cell_born;
    const float *src_row;
    float *dst;
    int nblk_inj;
    // corrections
    float *ul_next;
    int born_scale_m;
    float *grad;
    const float *usnap;
    float gcoef;   // already multiplied by irho when irho is constant
    const float *irho;
    const float *m;
    int scale_m;
    Region reg;
    // interpolation (two receiver sets)
    int np1, nc;
    const long long *coff1;
    const float *cw1;
    const float *cur1;
    float *row1;
    int nblk_rec1;
    int np2;
    const long long *coff2;
    const float *cw2;
    const float *cur2;
    float *row2;
};

__global__ void __launch_bounds__(128) sparse_kernel(SparseArgs a)
{
    int b = blockIdx.x;
    if (b < a.nblk_inj) {
        int i = b * blockDim.x + threadIdx.x;
        if (i >= a.ncell) return;
        int e0 = a.cell_start[i], e1 = a.cell_start[i + 1];
        float acc = 0.f;
        for (int e = e0; e < e1; e++) acc += a.ent_w[e] * a.src_row[a.ent_trace[e]];
        float inj = acc * a.cell_scale[i];
        long long o = a.cell_off[i];
        a.dst[o] += inj;
        if (a.ul_next) a.ul_next[o] += a.cell_born[i] * inj * (a.born_scale_m ? a.m[o] : 1.0f);
        if (a.grad) {
            int x = a.cell_ijk[3 * i], y = a.cell_ijk[3 * i + 1], z = a.cell_ijk[3 * i + 2];
            if (a.reg.has(x, y, z)) {
                long long ri = a.reg.at(x, y, z);
                float gc = a.gcoef * (a.irho ? a.irho[o] : 1.0f);
                if (a.scale_m) gc *= a.m[o];
                a.grad[ri] -= gc * a.usnap[ri] * inj;
            }
        }
        return;
    }
    b -= a.nblk_inj;
    if (b < a.nblk_rec1) {
        int p = b * blockDim.x + threadIdx.x;
        if (p >= a.np1) return;
        float acc = 0.f;
        for (int c = 0; c < a.nc; c++) {
            long long o = a.coff1[(size_t)p * a.nc + c];
            if (o >= 0) acc += a.cw1[(size_t)p * a.nc + c] * a.cur1[o];
        }
        a.row1[p] = acc;
        return;
    }
    b -= a.nblk_rec1;
    {
        int p = b * blockDim.x + threadIdx.x;
        if (p >= a.np2) return;
        float acc = 0.f;
        for (int c = 0; c < a.nc; c++) {
            long long o = a.coff2[(size_t)p * a.nc + c];
            if (o >= 0) acc += a.cw2[(size_t)p * a.nc + c] * a.cur2[o];
        }
        a.row2[p] = acc;
    }
}

// atomic injection path: one thread per sorted entry
__global__ void inject_atomic_kernel(int ncell, const long long *__restrict__ cell_off,
                                     const int *__restrict__ cell_start,
                                     const int *__restrict__ ent_trace,
                                     const float *__restrict__ ent_w,
                                     const float *__restrict__ cell_scale,
                                     const float *__restrict__ src_row, float *dst)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ncell) return;
    for (int e = cell_start[i]; e < cell_start[i + 1]; e++)
        atomicAdd(dst + cell_off[i], ent_w[e] * src_row[ent_trace[e]] * cell_scale[i]);
}

void sparse_step(const judi_model *M, const SparseSet *S_src, const float *src_row, float *dst,
                 const SparseCorr *corr, const SparseSet *S_rec, const float *cur, float *rec_row,
                 const SparseSet *S_rec2, const float *cur2, float *rec_row2)
{
    judi_ctx *ctx = M->ctx;
    SparseArgs a;
    memset(&a, 0, sizeof(a));
    const int T = 128;
    bool atomic = ctx->opt_atomic_inject && !(corr && (corr->ul_next || corr->grad));
    if (S_src && S_src->ncell > 0 && src_row && dst) {
        if (atomic) {
            inject_atomic_kernel<<<(S_src->ncell + T - 1) / T, T, 0, ctx->stream>>>(
                S_src->ncell, S_src->cell_off.p, S_src->cell_start.p, S_src->ent_trace.p,
                S_src->ent_w.p, S_src->cell_scale.p, src_row, dst);
            launch_count(ctx, "inject_atomic");
        } else {
            a.ncell = S_src->ncell;
            a.cell_off = S_src->cell_off.p;
            a.cell_start = S_src->cell_start.p;
            a.cell_ijk = S_src->cell_ijk.p;
            a.ent_trace = S_src->ent_trace.p;
            a.ent_w = S_src->ent_w.p;
            a.cell_scale = S_src->cell_scale.p;
            a.cell_born = S_src->cell_born.p;
            a.src_row = src_row;
            a.dst = dst;
            a.nblk_inj = (S_src->ncell + T - 1) / T;
            if (corr) {
                a.ul_next = corr->ul_next;
                a.born_scale_m = corr->born_scale_m;
                a.grad = corr->grad;
                a.usnap = corr->usnap;
                a.gcoef = corr->gcoef * (M->irho.p ? 1.0f : M->irho_c);
                a.irho = M->irho.p;
                a.m = M->m.p;
                a.scale_m = corr->scale_m;
                a.reg = corr->reg;
            }
        }
    }
    a.nc = 1 << M->ndim;
    if (S_rec && rec_row && S_rec->np > 0) {
        a.np1 = S_rec->np;
        a.coff1 = S_rec->corner_off.p;
        a.cw1 = S_rec->corner_w.p;
        a.cur1 = cur;
        a.row1 = rec_row;
        a.nblk_rec1 = (S_rec->np + T - 1) / T;
    }
    int nblk2 = 0;
    if (S_rec2 && rec_row2 && S_rec2->np > 0) {
        a.np2 = S_rec2->np;
        a.coff2 = S_rec2->corner_off.p;
        a.cw2 = S_rec2->corner_w.p;
        a.cur2 = cur2;
        a.row2 = rec_row2;
        nblk2 = (S_rec2->np + T - 1) / T;
    }
    int nblk = a.nblk_inj + a.nblk_rec1 + nblk2;
    if (nblk == 0) return;
    sparse_kernel<<<nblk, T, 0, ctx->stream>>>(a);
    launch_count(ctx, "sparse");
}
