// flux3d.cu -- vectorised 3-D kernels of the div(grad) forms: variable density
// (src/pysource/FD_utils.py:11-21) and the TTI rotated self-adjoint operator
// (FD_utils.py:24-105, kernels.py:144-182).  Two passes per time step:
//   pass 1  flux3d_vec:    P = R^T (A R G u1 + B R G u2), M = R^T (B R G u1 + C R G u2)
//                          (or P = irho * G u) on the grid extended by R nodes
//   pass 2  fluxdiv3d_vec: H = sum_d D-_d P_d, time update of 1 or 2 fields + fused extras
// Every thread owns 4 consecutive x points; all global accesses are 128-bit and aligned (x is a
// multiple of 4 relative to the 128-byte aligned grid origin).  The same arithmetic as the
// one-point-per-thread kernels in kernels.cu (which remain the 2-D path).
#include "common.cuh"

struct Flux3Args {
    Grid g;
    Coefs cf;
    const float *u1, *u2;
    const float *irho, *eps, *delta, *theta, *phi;
    float *P[3], *M[3];
};

__device__ __forceinline__ void ld4(const float *p, float *w)
{
    const float4 t = *(const float4 *)p;
    w[0] = t.x; w[1] = t.y; w[2] = t.z; w[3] = t.w;
}
__device__ __forceinline__ void st4(float *p, const float *w)
{
    *(float4 *)p = make_float4(w[0], w[1], w[2], w[3]);
}

// staggered gradient (D+x, D+y, D+z) of one field at the 4 nodes x..x+3 of row (y,z)
template <int R, int XE>
__device__ __forceinline__ void grad4(const Grid &g, const Coefs &cf, const float *__restrict__ u,
                                      long long i, float *gx, float *gy, float *gz)
{
    float w[4 + 2 * XE];
#pragma unroll
    for (int v = 0; v < 1 + 2 * (XE / 4); v++) ld4(u + i - XE + 4 * v, w + 4 * v);
#pragma unroll
    for (int e = 0; e < 4; e++) {
        float a = 0.f;
#pragma unroll
        for (int k = 1; k <= R; k++) a += cf.wx[k] * (w[XE + e + k] - w[XE + e - k + 1]);
        gx[e] = a;
        gy[e] = 0.f;
        gz[e] = 0.f;
    }
#pragma unroll
    for (int k = 1; k <= R; k++) {
        float p[4], q[4];
        ld4(u + i + k * g.sy, p);
        if (k == 1) { p[0] -= w[XE]; p[1] -= w[XE + 1]; p[2] -= w[XE + 2]; p[3] -= w[XE + 3]; }
        else { ld4(u + i - (k - 1) * g.sy, q); p[0] -= q[0]; p[1] -= q[1]; p[2] -= q[2]; p[3] -= q[3]; }
#pragma unroll
        for (int e = 0; e < 4; e++) gy[e] += cf.wy[k] * p[e];
        ld4(u + i + k * g.sz, p);
        if (k == 1) { p[0] -= w[XE]; p[1] -= w[XE + 1]; p[2] -= w[XE + 2]; p[3] -= w[XE + 3]; }
        else { ld4(u + i - (k - 1) * g.sz, q); p[0] -= q[0]; p[1] -= q[1]; p[2] -= q[2]; p[3] -= q[3]; }
#pragma unroll
        for (int e = 0; e < 4; e++) gz[e] += cf.wz[k] * p[e];
    }
}

template <int R, bool TTI>
__global__ void __launch_bounds__(128) flux3d_vec(Flux3Args a)
{
    constexpr int XE = ((R + 3) / 4) * 4;
    const Grid &g = a.g;
    const int x = -XE + 4 * (blockIdx.x * blockDim.x + threadIdx.x);
    const int y = (int)(blockIdx.y * blockDim.y + threadIdx.y) - R;
    const int z = (int)blockIdx.z - R;
    if (x >= g.n[0] + R || y >= g.n[1] + R) return;
    const long long i = g.at(x, y, z);
    const Coefs &cf = a.cf;
    float gx[4], gy[4], gz[4], b[4];
    grad4<R, XE>(g, cf, a.u1, i, gx, gy, gz);
    ld4(a.irho + i, b);
    if (!TTI) {
        float o[4];
#pragma unroll
        for (int e = 0; e < 4; e++) o[e] = b[e] * gx[e];
        st4(a.P[0] + i, o);
#pragma unroll
        for (int e = 0; e < 4; e++) o[e] = b[e] * gy[e];
        st4(a.P[1] + i, o);
#pragma unroll
        for (int e = 0; e < 4; e++) o[e] = b[e] * gz[e];
        st4(a.P[2] + i, o);
        return;
    }
    float hx[4], hy[4], hz[4], ep[4], dl[4], th[4], ph[4];
    grad4<R, XE>(g, cf, a.u2, i, hx, hy, hz);
    ld4(a.eps + i, ep);
    ld4(a.delta + i, dl);
    ld4(a.theta + i, th);
    ld4(a.phi + i, ph);
    float Px[4], Py[4], Pz[4], Mx[4], My[4], Mz[4];
#pragma unroll
    for (int e = 0; e < 4; e++) {
        float st, ct, sp, cp;
        sincosf(th[e], &st, &ct);
        sincosf(ph[e], &sp, &cp);
        // R = rot_axis2(theta) * rot_axis3(phi)  (FD_utils.py:24-47, sympy conventions)
        const float R00 = ct * cp, R01 = ct * sp, R02 = -st;
        const float R10 = -sp, R11 = cp;
        const float R20 = st * cp, R21 = st * sp, R22 = ct;
        const float A0 = b[e] * dl[e], A2 = b[e];
        const float B0 = b[e] * sqrtf((ep[e] - dl[e]) * dl[e]);
        const float C0 = b[e] * (ep[e] - dl[e]);
        const float r1x = R00 * gx[e] + R01 * gy[e] + R02 * gz[e];
        const float r1y = R10 * gx[e] + R11 * gy[e];
        const float r1z = R20 * gx[e] + R21 * gy[e] + R22 * gz[e];
        const float r2x = R00 * hx[e] + R01 * hy[e] + R02 * hz[e];
        const float r2y = R10 * hx[e] + R11 * hy[e];
        const float px = A0 * r1x + B0 * r2x, py = A0 * r1y + B0 * r2y, pz = A2 * r1z;
        const float mx = B0 * r1x + C0 * r2x, my = B0 * r1y + C0 * r2y;
        Px[e] = R00 * px + R10 * py + R20 * pz;
        Py[e] = R01 * px + R11 * py + R21 * pz;
        Pz[e] = R02 * px + R22 * pz;
        Mx[e] = R00 * mx + R10 * my;
        My[e] = R01 * mx + R11 * my;
        Mz[e] = R02 * mx;
    }
    st4(a.P[0] + i, Px); st4(a.P[1] + i, Py); st4(a.P[2] + i, Pz);
    st4(a.M[0] + i, Mx); st4(a.M[1] + i, My); st4(a.M[2] + i, Mz);
}

// ------------------------------------------------------------------------------------------
struct FluxDiv3Args {
    Grid g;
    Coefs cf;
    const float *u[2];
    float *up[2];
    const float *P[3], *M[3];
    const float *m, *irho, *dm;
    const float *dx, *dy, *dz;
    float *dd_out[2];
    const float *dd_in[2];
    const float *qext;
    float qext_scale;
    Fuse f;
};

template <int R, int XE>
__device__ __forceinline__ void div4(const Grid &g, const Coefs &cf, const float *const *F,
                                     long long i, float *L)
{
    float w[4 + 2 * XE];
#pragma unroll
    for (int v = 0; v < 1 + 2 * (XE / 4); v++) ld4(F[0] + i - XE + 4 * v, w + 4 * v);
#pragma unroll
    for (int e = 0; e < 4; e++) {
        float a = 0.f;
#pragma unroll
        for (int k = 1; k <= R; k++) a += cf.wx[k] * (w[XE + e + k - 1] - w[XE + e - k]);
        L[e] = a;
    }
    float ay[4] = {0.f, 0.f, 0.f, 0.f}, az[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 1; k <= R; k++) {
        float p[4], q[4];
        ld4(F[1] + i + (k - 1) * g.sy, p);
        ld4(F[1] + i - k * g.sy, q);
#pragma unroll
        for (int e = 0; e < 4; e++) ay[e] += cf.wy[k] * (p[e] - q[e]);
        ld4(F[2] + i + (k - 1) * g.sz, p);
        ld4(F[2] + i - k * g.sz, q);
#pragma unroll
        for (int e = 0; e < 4; e++) az[e] += cf.wz[k] * (p[e] - q[e]);
    }
    // same association as the scalar kernel: (Dx + Dz) + Dy
#pragma unroll
    for (int e = 0; e < 4; e++) L[e] = (L[e] + az[e]) + ay[e];
}

template <int R, int NF>
__global__ void __launch_bounds__(128) fluxdiv3d_vec(FluxDiv3Args a)
{
    constexpr int XE = ((R + 3) / 4) * 4;
    const Grid &g = a.g;
    const int x = 4 * (blockIdx.x * blockDim.x + threadIdx.x);
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int z = blockIdx.z;
    if (x >= g.n[0] || y >= g.n[1]) return;
    const long long i = g.at(x, y, z);
    const Coefs &cf = a.cf;
    float ir[4], mm[4], dxv[4], rden[4], sd[4];
    ld4(a.irho + i, ir);
    ld4(a.m + i, mm);
    ld4(a.dx + x, dxv);
    const float dyz = a.dy[y];
    const float dzv = a.dz[z];
#pragma unroll
    for (int e = 0; e < 4; e++) {
        sd[e] = cf.dt * ((dxv[e] + dyz) + dzv);
        rden[e] = 1.0f / (mm[e] * ir[e] + sd[e]);
    }
    const bool xfull = x + 3 < g.n[0];
    float uc[NF][4], delta[NF][4];
#pragma unroll
    for (int f = 0; f < NF; f++) {
        float L[4], upv[4], un[4], dmv[4], ddv[4];
        div4<R, XE>(g, cf, f == 0 ? a.P : a.M, i, L);
        ld4(a.u[f] + i, uc[f]);
        ld4(a.up[f] + i, upv);
        const bool born = a.dd_in[f] != nullptr;
        if (born) { ld4(a.dm + i, dmv); ld4(a.dd_in[f] + i, ddv); }
        float qe[4] = {0.f, 0.f, 0.f, 0.f};
        if (a.qext) ld4(a.qext + i, qe);
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const float d1 = uc[f][e] - upv[e];
            const float q = (born ? -dmv[e] * ir[e] * ddv[e] : 0.f) + a.qext_scale * qe[e];
            delta[f][e] = (cf.dt2 * L[e] + q - sd[e] * d1) * rden[e];
            un[e] = (uc[f][e] + d1) + delta[f][e];
        }
        if (xfull) {
            st4(a.up[f] + i, un);
            if (a.dd_out[f]) st4(a.dd_out[f] + i, delta[f]);
        } else {
#pragma unroll
            for (int e = 0; e < 4; e++)
                if (x + e < g.n[0]) {
                    a.up[f][i + e] = un[e];
                    if (a.dd_out[f]) a.dd_out[f][i + e] = delta[f][e];
                }
        }
    }
    const Fuse &fz = a.f;
    if ((fz.snap[0] || fz.grad || fz.illum || fz.wrec) && y >= fz.reg.lo[1] && y < fz.reg.hi[1] &&
        z >= fz.reg.lo[2] && z < fz.reg.hi[2]) {
        const long long ri = fz.reg.at(x, y, z);
#pragma unroll
        for (int e = 0; e < 4; e++) {
            if (x + e < fz.reg.lo[0] || x + e >= fz.reg.hi[0]) continue;
            float e2 = 0.f, ic = 0.f;
#pragma unroll
            for (int f = 0; f < NF; f++) {
                if (fz.snap[f]) fz.snap[f][ri + e] = uc[f][e];
                e2 += uc[f][e] * uc[f][e];
                if (fz.grad) ic += fz.usnap[f][ri + e] * delta[f][e];
            }
            if (fz.illum) fz.illum[ri + e] += cf.dt * e2;
            if (fz.wrec) fz.wrec[ri + e] += fz.wrec_scale * (NF == 2 ? uc[0][e] + uc[1][e] : uc[0][e]);
            if (fz.grad) fz.grad[ri + e] -= fz.gcoef * ir[e] * ic;
        }
    }
}

// ------------------------------------------------------------------------------------------
template <int R>
static void flux3d_one(const judi_model *M, Wavefield *W, const float *const *u, float *const *up,
                       float *const *dd_out, const float *const *dd_in, const Fuse &fuse,
                       const float *qext = nullptr, float qext_scale = 0.f)
{
    constexpr int XE = ((R + 3) / 4) * 4;
    judi_ctx *ctx = M->ctx;
    const int nf = M->tti ? 2 : 1;
    Flux3Args fa;
    memset(&fa, 0, sizeof(fa));
    fa.g = M->g; fa.cf = M->cf;
    fa.u1 = u[0]; fa.u2 = nf == 2 ? u[1] : nullptr;
    fa.irho = M->irho.p; fa.eps = M->eps.p; fa.delta = M->delta.p;
    fa.theta = M->theta.p; fa.phi = M->phi.p;
    for (int d = 0; d < 3; d++) { fa.P[d] = W->flux[d].p; fa.M[d] = W->flux[3 + d].p; }
    dim3 blk(32, 4, 1);
    int nxt = (M->N[0] + R + XE + 3) / 4;  // threads along x covering [-XE, nx + R)
    dim3 grd((nxt + 31) / 32, (M->N[1] + 2 * R + 3) / 4, M->N[2] + 2 * R);
    if (M->tti) flux3d_vec<R, true><<<grd, blk, 0, ctx->stream>>>(fa);
    else flux3d_vec<R, false><<<grd, blk, 0, ctx->stream>>>(fa);
    launch_count(ctx, "flux3d_vec");
    FluxDiv3Args ua;
    memset(&ua, 0, sizeof(ua));
    ua.g = M->g; ua.cf = M->cf;
    for (int f = 0; f < nf; f++) {
        ua.u[f] = u[f]; ua.up[f] = up[f];
        ua.dd_out[f] = dd_out ? dd_out[f] : nullptr;
        ua.dd_in[f] = dd_in ? dd_in[f] : nullptr;
    }
    for (int d = 0; d < 3; d++) { ua.P[d] = W->flux[d].p; ua.M[d] = W->flux[3 + d].p; }
    ua.m = M->m.p; ua.irho = M->irho.p; ua.dm = M->dm.p;
    ua.dx = M->damp[0].p; ua.dy = M->damp[1].p; ua.dz = M->damp[2].p;
    ua.qext = qext;
    ua.qext_scale = qext_scale;
    ua.f = fuse;
    dim3 grd2(((M->N[0] + 3) / 4 + 31) / 32, (M->N[1] + 3) / 4, M->N[2]);
    if (nf == 2) fluxdiv3d_vec<R, 2><<<grd2, blk, 0, ctx->stream>>>(ua);
    else fluxdiv3d_vec<R, 1><<<grd2, blk, 0, ctx->stream>>>(ua);
    launch_count(ctx, "fluxdiv3d_vec");
}

// One time step of the flux-form physics in 3-D (plain or Born).  Returns false if the radius is
// not instantiated (caller falls back to the scalar kernels).
bool flux3d_step(const judi_model *M, const StepArgs &s)
{
    Wavefield *W = s.scratch;
    JUDI_REQUIRE(W != nullptr, "flux scratch not bound");
    auto run = [&](auto rtag) {
        constexpr int R = decltype(rtag)::value;
        if (!s.born) {
            flux3d_one<R>(M, W, s.u, s.up, nullptr, nullptr, s.f, s.qext, s.qext_scale);
            return;
        }
        JUDI_REQUIRE(s.born_ic == JUDI_IC_AS,
                     "isic/fwi Born sources are only implemented for constant-density acoustics");
        float *dd[2] = {W->dd[0].p, W->dd[1].p};
        const float *ddc[2] = {W->dd[0].p, W->dd[1].p};
        flux3d_one<R>(M, W, s.u, s.up, dd, nullptr, s.f, s.qext, s.qext_scale);
        Fuse none;
        memset(&none, 0, sizeof(none));
        flux3d_one<R>(M, W, s.ul, s.ulp, nullptr, ddc, none);
    };
    switch (M->r) {
    case 2: run(std::integral_constant<int, 2>()); return true;
    case 4: run(std::integral_constant<int, 4>()); return true;
    case 8: run(std::integral_constant<int, 8>()); return true;
    default: return false;
    }
}
