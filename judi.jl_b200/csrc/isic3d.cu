// isic3d.cu -- the isic / fwi imaging condition of the 3-D constant-density acoustic gradient
// (src/pysource/sensitivity.py:106-122, 212-228):
//
//   gradm -= w ( u_s v.dt2 m + ics grad(u_s, +1/2) . grad(v, +1/2) ),   w = dt_save irho
//
// as ONE vectorised pass after the TMA time step of the adjoint field (which also writes
// delta = dt^2 v.dt2, flavour F_DDOUT).  The one-point-per-thread kernel it replaces gathers ~190
// scalars per grid point (2.0 ms per imaging step on the C3 grid); here every thread owns 4
// consecutive x-points, marches along z with the z-lines of both fields in register queues and
// takes the x / y lines as 128-bit loads through L1.  The snapshot is first expanded from its
// compact region layout into a Grid-layout scratch field, whose zero halo is exactly the "zero
// outside the region" rule of the reference's subsampled wavefield.
#include "common.cuh"

// snapshot slot (region layout) -> Grid-layout field (domain points only; halo stays zero)
__global__ void region_to_grid_kernel(Grid g, Region reg, const float *__restrict__ src, float *__restrict__ dst)
{
    const int x = 4 * (blockIdx.x * blockDim.x + threadIdx.x), y = blockIdx.y, z = blockIdx.z;
    if (x >= g.n[0]) return;
    const long long i = g.at(x, y, z);
    if (reg.has(x, y, z) && reg.has(x + 3, y, z) && ((reg.lo[0] & 3) == 0) && x + 3 < g.n[0]) {
        *(float4 *)(dst + i) = *(const float4 *)(src + reg.at(x, y, z));
        return;
    }
#pragma unroll
    for (int e = 0; e < 4; e++)
        if (x + e < g.n[0]) dst[i + e] = reg.has(x + e, y, z) ? src[reg.at(x + e, y, z)] : 0.f;
}

struct IsicArgs {
    Grid g;
    Coefs cf;
    Region reg;
    const float *us, *v, *dd, *m;   // Grid layout: snapshot, adjoint field (level t), delta, m
    float *grad;                    // region layout
    float gc, ics;
    int zchunk;
};

// staggered D+ along x at the 4 nodes of a group from the 12-float window (l, c, r)
template <int R>
__device__ __forceinline__ void dplus_x4(const float *wx, const float4 &l, const float4 &c, const float4 &r, float *o)
{
    const float w[12] = {l.x, l.y, l.z, l.w, c.x, c.y, c.z, c.w, r.x, r.y, r.z, r.w};
#pragma unroll
    for (int e = 0; e < 4; e++) {
        float acc = 0.f;
#pragma unroll
        for (int k = 1; k <= R; k++) acc += wx[k] * (w[4 + e + k] - w[4 + e - k + 1]);
        o[e] = acc;
    }
}

template <int R>
__global__ void __launch_bounds__(256, 2) isic_ic3d(const IsicArgs a)
{
    const Grid &g = a.g;
    const Coefs &cf = a.cf;
    constexpr int NQ = 2 * R;
    const int x = 4 * (blockIdx.x * 32 + threadIdx.x), y = blockIdx.y * 8 + threadIdx.y;
    if (x >= g.n[0] || y >= g.n[1]) return;
    const int z0 = blockIdx.z * a.zchunk, z1 = min(z0 + a.zchunk, g.n[2]);
    long long i = g.at(x, y, z0);
    const long long sy = g.sy, sz = g.sz;
    // q[k] holds plane z - R + 1 + k of the own column
    float4 qu[NQ], qv[NQ];
#pragma unroll
    for (int k = 1; k < NQ; k++) {
        qu[k] = *(const float4 *)(a.us + i + (long long)(k - R) * sz);
        qv[k] = *(const float4 *)(a.v + i + (long long)(k - R) * sz);
    }
    const bool xfull = x + 3 < g.n[0];
    const bool vec = xfull && (a.reg.lo[0] & 3) == 0 && x >= a.reg.lo[0] && x + 3 < a.reg.hi[0] &&
                     y >= a.reg.lo[1] && y < a.reg.hi[1];
    for (int z = z0; z < z1; z++, i += sz) {
#pragma unroll
        for (int k = 0; k < NQ - 1; k++) { qu[k] = qu[k + 1]; qv[k] = qv[k + 1]; }
        qu[NQ - 1] = *(const float4 *)(a.us + i + (long long)R * sz);
        qv[NQ - 1] = *(const float4 *)(a.v + i + (long long)R * sz);
        const float4 uc = qu[R - 1], vc = qv[R - 1];
        float gu[3][4], gv[3][4];
        dplus_x4<R>(cf.wx, *(const float4 *)(a.us + i - 4), uc, *(const float4 *)(a.us + i + 4), gu[0]);
        dplus_x4<R>(cf.wx, *(const float4 *)(a.v + i - 4), vc, *(const float4 *)(a.v + i + 4), gv[0]);
#pragma unroll
        for (int e = 0; e < 4; e++) gu[1][e] = gu[2][e] = gv[1][e] = gv[2][e] = 0.f;
#pragma unroll
        for (int k = 1; k <= R; k++) {
            const float4 up = *(const float4 *)(a.us + i + k * sy);
            const float4 um = (k == 1) ? uc : *(const float4 *)(a.us + i - (k - 1) * sy);
            const float4 vp = *(const float4 *)(a.v + i + k * sy);
            const float4 vm = (k == 1) ? vc : *(const float4 *)(a.v + i - (k - 1) * sy);
            const float4 uzp = qu[R - 1 + k], uzm = qu[R - k], vzp = qv[R - 1 + k], vzm = qv[R - k];
            gu[1][0] += cf.wy[k] * (up.x - um.x); gu[1][1] += cf.wy[k] * (up.y - um.y);
            gu[1][2] += cf.wy[k] * (up.z - um.z); gu[1][3] += cf.wy[k] * (up.w - um.w);
            gv[1][0] += cf.wy[k] * (vp.x - vm.x); gv[1][1] += cf.wy[k] * (vp.y - vm.y);
            gv[1][2] += cf.wy[k] * (vp.z - vm.z); gv[1][3] += cf.wy[k] * (vp.w - vm.w);
            gu[2][0] += cf.wz[k] * (uzp.x - uzm.x); gu[2][1] += cf.wz[k] * (uzp.y - uzm.y);
            gu[2][2] += cf.wz[k] * (uzp.z - uzm.z); gu[2][3] += cf.wz[k] * (uzp.w - uzm.w);
            gv[2][0] += cf.wz[k] * (vzp.x - vzm.x); gv[2][1] += cf.wz[k] * (vzp.y - vzm.y);
            gv[2][2] += cf.wz[k] * (vzp.z - vzm.z); gv[2][3] += cf.wz[k] * (vzp.w - vzm.w);
        }
        if (z < a.reg.lo[2] || z >= a.reg.hi[2]) continue;
        const float4 d4 = *(const float4 *)(a.dd + i), m4 = *(const float4 *)(a.m + i);
        const float us[4] = {uc.x, uc.y, uc.z, uc.w}, dl[4] = {d4.x, d4.y, d4.z, d4.w};
        const float mm[4] = {m4.x, m4.y, m4.z, m4.w};
        float inc[4];
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const float ig = gu[0][e] * gv[0][e] + gu[1][e] * gv[1][e] + gu[2][e] * gv[2][e];
            inc[e] = a.gc * (us[e] * dl[e] * mm[e] + a.ics * cf.dt2 * ig);
        }
        const long long ri = a.reg.at(x, y, z);
        if (vec) {
            float4 G = *(float4 *)(a.grad + ri);
            G.x -= inc[0]; G.y -= inc[1]; G.z -= inc[2]; G.w -= inc[3];
            *(float4 *)(a.grad + ri) = G;
        } else if (y >= a.reg.lo[1] && y < a.reg.hi[1]) {
#pragma unroll
            for (int e = 0; e < 4; e++)
                if (x + e < g.n[0] && x + e >= a.reg.lo[0] && x + e < a.reg.hi[0]) a.grad[ri + e] -= inc[e];
        }
    }
}

bool isic3d_supported(const judi_model *M)
{
    return M->ndim == 3 && !M->tti && !M->irho_field && !M->visco && (M->r == 2 || M->r == 4);
}

// gradm -= gcoef irho ( u_s delta m + ics dt^2 inner_grad(u_s, v) ) for the snapshot `usnap` (region
// layout), the adjoint field `v` at level t and delta = v+ - 2v + v- (both Grid layout); `usg` is a
// Grid-layout scratch field (zero halo) of the caller.
void isic3d_apply(const judi_model *M, const Fuse &f, const float *v, const float *dd, float *usg)
{
    judi_ctx *ctx = M->ctx;
    {
        dim3 blk(128), grd((M->N[0] + 511) / 512, M->N[1], M->N[2]);
        region_to_grid_kernel<<<grd, blk, 0, ctx->stream>>>(M->g, f.reg, f.usnap[0], usg);
        launch_count(ctx, "region_to_grid");
    }
    IsicArgs a;
    a.g = M->g; a.cf = M->cf; a.reg = f.reg;
    a.us = usg; a.v = v; a.dd = dd; a.m = M->m.p;
    a.grad = f.grad;
    a.gc = f.gcoef * M->irho_c;
    a.ics = f.ic == JUDI_IC_FWI ? -1.f : 1.f;
    const int nbx = (M->N[0] + 127) / 128, nby = (M->N[1] + 7) / 8;
    // z-chunks: enough CTAs for ~4 waves of 2 CTAs per SM, chunks of at least 16 planes
    int nzc = std::max(1, std::min(M->N[2] / 16, (8 * ctx->num_sms + nbx * nby - 1) / (nbx * nby)));
    a.zchunk = (M->N[2] + nzc - 1) / nzc;
    nzc = (M->N[2] + a.zchunk - 1) / a.zchunk;
    dim3 blk(32, 8), grd(nbx, nby, nzc);
    if (M->r == 2) isic_ic3d<2><<<grd, blk, 0, ctx->stream>>>(a);
    else isic_ic3d<4><<<grd, blk, 0, ctx->stream>>>(a);
    launch_count(ctx, "isic_ic3d");
}
