// kernels.cu -- generic stencil kernels and the per-time-step dispatcher.
//
//  * acoustic_generic<NDIM,R>: isotropic acoustic, constant density (Laplacian form), one thread
//    per grid point, optional fused Born field / snapshot save / imaging condition / illumination.
//    Used for 2-D grids, for cross-checking, and as the Born / ISIC path in 3-D.
//  * flux_kernel + flux_update: the div(grad) forms -- variable density and TTI
//    (src/pysource/FD_utils.py:11-21, 24-105; kernels.py:144-182) in two passes.
//  The headline 3-D constant-density kernel is in tma3d.cu.
//
// Time update (restates kernels.py:61-69), written in the increment form
//     u+ = u + (u - u-) + delta,   delta = (dt^2 (L + q) - dt damp (u - u-)) / (m irho + dt damp)
// which is algebraically identical to solve(m irho u.dt2 + damp u.dt - L - q, u.forward) and
// gives delta = u+ - 2u + u- = dt^2 u.dt2 for free (imaging condition, Born source).
#include "common.cuh"

// ------------------------------------------------------------------------------------------
struct AcArgs {
    Grid g;
    Coefs cf;
    const float *u;
    float *up;
    const float *m;
    const float *dx, *dy, *dz;  // damping profiles
    float sdamp;                // dt / irho_c
    float irho_c;
    const float *ul;            // Born (nullptr: off)
    float *ulp;
    const float *dm;
    const float *dms;           // dm as the factor of u.dt2 (judi_model::dm_src)
    int born_ic;
    const float *qext;          // extended / wavefield source (nullptr: off)
    float qext_scale;
    Fuse f;
};

template <int NDIM, int R>
__device__ __forceinline__ float lap_at(const Grid &g, const Coefs &cf, const float *__restrict__ u,
                                        long long i)
{
    // same summation order as the TMA kernel (z, then x, then y)
    float acc = cf.c0 * u[i];
#pragma unroll
    for (int k = 1; k <= R; k++) acc += cf.cz[k] * (u[i + k * g.sz] + u[i - k * g.sz]);
#pragma unroll
    for (int k = 1; k <= R; k++) acc += cf.cx[k] * (u[i + k] + u[i - k]);
    if (NDIM == 3) {
#pragma unroll
        for (int k = 1; k <= R; k++) acc += cf.cy[k] * (u[i + k * g.sy] + u[i - k * g.sy]);
    }
    return acc;
}

// staggered derivative D+ along stride s at node i (value lives at i + 1/2)
template <int R>
__device__ __forceinline__ float dplus(const float *__restrict__ f, long long i, long long s,
                                       const float *w)
{
    float a = 0.f;
#pragma unroll
    for (int k = 1; k <= R; k++) a += w[k] * (f[i + k * s] - f[i - (k - 1) * s]);
    return a;
}

template <int R>
__device__ __forceinline__ float dminus(const float *__restrict__ f, long long i, long long s,
                                        const float *w)
{
    float a = 0.f;
#pragma unroll
    for (int k = 1; k <= R; k++) a += w[k] * (f[i + (k - 1) * s] - f[i - k * s]);
    return a;
}

// snapshot value with zero outside the region (only used by the ISIC / FWI conditions)
__device__ __forceinline__ float snap_at(const Region &r, const float *s, int x, int y, int z)
{
    return r.has(x, y, z) ? s[r.at(x, y, z)] : 0.f;
}

template <int NDIM, int R>
__device__ float inner_grad_snap(const Grid &g, const Coefs &cf, const Region &reg,
                                 const float *us, const float *__restrict__ v, int x, int y, int z)
{
    long long i = g.at(x, y, z);
    float gux = 0.f, guy = 0.f, guz = 0.f;
#pragma unroll
    for (int k = 1; k <= R; k++) {
        gux += cf.wx[k] * (snap_at(reg, us, x + k, y, z) - snap_at(reg, us, x - (k - 1), y, z));
        if (NDIM == 3)
            guy += cf.wy[k] * (snap_at(reg, us, x, y + k, z) - snap_at(reg, us, x, y - (k - 1), z));
        guz += cf.wz[k] * (snap_at(reg, us, x, y, z + k) - snap_at(reg, us, x, y, z - (k - 1)));
    }
    float s = gux * dplus<R>(v, i, 1, cf.wx) + guz * dplus<R>(v, i, g.sz, cf.wz);
    if (NDIM == 3) s += guy * dplus<R>(v, i, g.sy, cf.wy);
    return s;
}

// div(c * grad(u,+1/2),-1/2) with c = dm (constant irho factored out): ISIC Born source
template <int NDIM, int R>
__device__ float lap_weighted(const Grid &g, const Coefs &cf, const float *__restrict__ u,
                              const float *__restrict__ c, long long i)
{
    float acc = 0.f;
    const long long st[3] = {1, g.sy, g.sz};
    const float *ws[3] = {cf.wx, cf.wy, cf.wz};
#pragma unroll
    for (int d = 0; d < 3; d++) {
        if (d == 1 && NDIM == 2) continue;
#pragma unroll
        for (int k = 1; k <= R; k++) {
            long long ip = i + (k - 1) * st[d], im = i - k * st[d];
            acc += ws[d][k] * (c[ip] * dplus<R>(u, ip, st[d], ws[d]) -
                               c[im] * dplus<R>(u, im, st[d], ws[d]));
        }
    }
    return acc;
}

// the same with the node-valued coefficient c = c1 * c2 (dm * irho as fields: variable density, TTI)
template <int NDIM, int R>
__device__ float lap_weighted2(const Grid &g, const Coefs &cf, const float *__restrict__ u,
                               const float *__restrict__ c1, const float *__restrict__ c2, long long i)
{
    float acc = 0.f;
    const long long st[3] = {1, g.sy, g.sz};
    const float *ws[3] = {cf.wx, cf.wy, cf.wz};
#pragma unroll
    for (int d = 0; d < 3; d++) {
        if (d == 1 && NDIM == 2) continue;
#pragma unroll
        for (int k = 1; k <= R; k++) {
            long long ip = i + (k - 1) * st[d], im = i - k * st[d];
            acc += ws[d][k] * (c1[ip] * c2[ip] * dplus<R>(u, ip, st[d], ws[d]) -
                               c1[im] * c2[im] * dplus<R>(u, im, st[d], ws[d]));
        }
    }
    return acc;
}

template <int NDIM, int R>
__global__ void __launch_bounds__(256) acoustic_generic(AcArgs a)
{
    const Grid &g = a.g;
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y, z;
    if (NDIM == 3) { y = blockIdx.y * blockDim.y + threadIdx.y; z = blockIdx.z; }
    else { y = 0; z = blockIdx.y * blockDim.y + threadIdx.y; }
    if (x >= g.n[0] || y >= g.n[1] || z >= g.n[2]) return;
    long long i = g.at(x, y, z);
    float uc = a.u[i];
    float lap = lap_at<NDIM, R>(g, a.cf, a.u, i);
    if (a.qext) lap += a.qext_scale * a.qext[i];
    float s = ((a.dx[x] + a.dy[y]) + a.dz[z]) * a.sdamp;
    float mm = a.m[i];
    float rden = 1.0f / (mm + s);
    float upv = a.up[i];
    float d1 = uc - upv;
    float delta = (lap - s * d1) * rden;
    a.up[i] = (uc + d1) + delta;
    if (a.ul) {
        float ulc = a.ul[i];
        float lapl = lap_at<NDIM, R>(g, a.cf, a.ul, i);
        float ulpv = a.ulp[i];
        float d1l = ulc - ulpv;
        float q;  // (dt^2/irho) * lin_src   (sensitivity.py:174-209)
        if (a.born_ic == JUDI_IC_AS) q = -a.dms[i] * delta;
        else {
            float ics = a.born_ic == JUDI_IC_FWI ? -1.f : 1.f;
            q = -(a.dms[i] * mm * delta -
                  ics * a.cf.dt2 * lap_weighted<NDIM, R>(g, a.cf, a.u, a.dm, i));
        }
        float deltal = (lapl - s * d1l + q) * rden;
        a.ulp[i] = (ulc + d1l) + deltal;
    }
    const Fuse &f = a.f;
    if ((f.snap[0] || f.grad || f.illum || f.wrec) && f.reg.has(x, y, z)) {
        long long ri = f.reg.at(x, y, z);
        if (f.wrec) f.wrec[ri] += f.wrec_scale * uc;
        if (f.snap[0]) f.snap[0][ri] = uc;
        if (f.illum) f.illum[ri] += a.cf.dt * uc * uc;
        if (f.grad) {
            float us = f.usnap[0][ri];
            float gc = f.gcoef * a.irho_c;
            if (f.ic == JUDI_IC_AS) f.grad[ri] -= gc * us * delta;
            else {
                float ics = f.ic == JUDI_IC_FWI ? -1.f : 1.f;
                float ig = inner_grad_snap<NDIM, R>(g, a.cf, f.reg, f.usnap[0], a.u, x, y, z);
                f.grad[ri] -= gc * (us * delta * mm + ics * a.cf.dt2 * ig);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// Flux form, pass 1: P = R^T (A R G u1 + B R G u2), M = R^T (B R G u1 + C R G u2) on the grid
// extended by R nodes; for plain variable density P = irho * G u.
struct FluxArgs {
    Grid g;
    Coefs cf;
    const float *u1, *u2;
    const float *irho, *eps, *delta, *theta, *phi;
    float *P[3], *M[3];
    int tti;
};

template <int NDIM, int R>
__global__ void __launch_bounds__(256) flux_kernel(FluxArgs a)
{
    const Grid &g = a.g;
    int x = blockIdx.x * blockDim.x + threadIdx.x - R;
    int y, z;
    if (NDIM == 3) { y = blockIdx.y * blockDim.y + threadIdx.y - R; z = (int)blockIdx.z - R; }
    else { y = 0; z = blockIdx.y * blockDim.y + threadIdx.y - R; }
    if (x >= g.n[0] + R || z >= g.n[2] + R) return;
    if (NDIM == 3 && y >= g.n[1] + R) return;
    long long i = g.at(x, y, z);
    const Coefs &cf = a.cf;
    float b = a.irho[i];
    float gx = dplus<R>(a.u1, i, 1, cf.wx);
    float gy = NDIM == 3 ? dplus<R>(a.u1, i, g.sy, cf.wy) : 0.f;
    float gz = dplus<R>(a.u1, i, g.sz, cf.wz);
    if (!a.tti) {
        a.P[0][i] = b * gx;
        if (NDIM == 3) a.P[1][i] = b * gy;
        a.P[2][i] = b * gz;
        return;
    }
    float hx = dplus<R>(a.u2, i, 1, cf.wx);
    float hy = NDIM == 3 ? dplus<R>(a.u2, i, g.sy, cf.wy) : 0.f;
    float hz = dplus<R>(a.u2, i, g.sz, cf.wz);
    float st, ct, sp = 0.f, cp = 1.f;
    sincosf(a.theta[i], &st, &ct);
    if (NDIM == 3) sincosf(a.phi[i], &sp, &cp);
    // R = rot_axis2(theta) * rot_axis3(phi)  (FD_utils.py:24-47, sympy conventions)
    float R00 = ct * cp, R01 = ct * sp, R02 = -st;
    float R10 = -sp, R11 = cp;
    float R20 = st * cp, R21 = st * sp, R22 = ct;
    float e = a.eps[i], dl = a.delta[i];
    float A0 = b * dl, A2 = b;
    float B0 = b * sqrtf((e - dl) * dl);
    float C0 = b * (e - dl);
    float r1x = R00 * gx + R01 * gy + R02 * gz;
    float r1y = R10 * gx + R11 * gy;
    float r1z = R20 * gx + R21 * gy + R22 * gz;
    float r2x = R00 * hx + R01 * hy + R02 * hz;
    float r2y = R10 * hx + R11 * hy;
    float r2z = R20 * hx + R21 * hy + R22 * hz;
    (void)r2z;
    float px = A0 * r1x + B0 * r2x, py = A0 * r1y + B0 * r2y, pz = A2 * r1z;
    float mx = B0 * r1x + C0 * r2x, my = B0 * r1y + C0 * r2y;
    if (NDIM == 2) { py = 0.f; my = 0.f; }
    a.P[0][i] = R00 * px + R10 * py + R20 * pz;
    a.P[2][i] = R02 * px + R22 * pz;
    a.M[0][i] = R00 * mx + R10 * my;
    a.M[2][i] = R02 * mx;
    if (NDIM == 3) {
        a.P[1][i] = R01 * px + R11 * py + R21 * pz;
        a.M[1][i] = R01 * mx + R11 * my;
    }
}

// Flux form, pass 2: divergence + time update of NF fields (+ fused epilogue).
struct FluxUpdArgs {
    Grid g;
    Coefs cf;
    const float *u[2];
    float *up[2];
    const float *P[3], *M[3];
    const float *m, *irho, *dm, *dms;
    const float *dx, *dy, *dz;
    float *dd_out[2];        // write delta of each field (background pass of Born)
    const float *dd_in[2];   // read delta of the background field (linearised pass of Born)
    const float *qext;       // extended / wavefield source, fed to every field
    float qext_scale;
    int born_ic;             // imaging condition of the Born source (linearised pass)
    const float *ub[2];      // isic / fwi Born source: background field(s), level t
    Fuse f;
};

template <int NDIM, int R, int NF>
__global__ void __launch_bounds__(256) flux_update(FluxUpdArgs a)
{
    const Grid &g = a.g;
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y, z;
    if (NDIM == 3) { y = blockIdx.y * blockDim.y + threadIdx.y; z = blockIdx.z; }
    else { y = 0; z = blockIdx.y * blockDim.y + threadIdx.y; }
    if (x >= g.n[0] || y >= g.n[1] || z >= g.n[2]) return;
    long long i = g.at(x, y, z);
    const Coefs &cf = a.cf;
    float ir = a.irho[i];
    float mm = a.m[i];
    float sd = cf.dt * ((a.dx[x] + a.dy[y]) + a.dz[z]);
    float rden = 1.0f / (mm * ir + sd);
    float uc[NF], delta[NF];
#pragma unroll
    for (int f = 0; f < NF; f++) {
        const float *const *F = f == 0 ? a.P : a.M;
        float L = dminus<R>(F[0], i, 1, cf.wx) + dminus<R>(F[2], i, g.sz, cf.wz);
        if (NDIM == 3) L += dminus<R>(F[1], i, g.sy, cf.wy);
        uc[f] = a.u[f][i];
        float upv = a.up[f][i];
        float d1 = uc[f] - upv;
        float q = 0.f;
        if (a.dd_in[f]) {
            // dt^2 * lin_src: "as" -dm irho u.dt2; isic / fwi -(dm irho u.dt2 m -+ laplacian(u, dm irho))
            // (sensitivity.py:174-209)
            if (a.born_ic == JUDI_IC_AS) q = -a.dms[i] * ir * a.dd_in[f][i];
            else {
                float ics = a.born_ic == JUDI_IC_FWI ? -1.f : 1.f;
                q = -(a.dms[i] * ir * mm * a.dd_in[f][i] -
                      ics * cf.dt2 * lap_weighted2<NDIM, R>(g, cf, a.ub[f], a.dm, a.irho, i));
            }
        }
        if (a.qext) q += a.qext_scale * a.qext[i];
        delta[f] = (cf.dt2 * L + q - sd * d1) * rden;
        a.up[f][i] = (uc[f] + d1) + delta[f];
        if (a.dd_out[f]) a.dd_out[f][i] = delta[f];
    }
    const Fuse &fz = a.f;
    if ((fz.snap[0] || fz.grad || fz.illum || fz.wrec) && fz.reg.has(x, y, z)) {
        long long ri = fz.reg.at(x, y, z);
        float e2 = 0.f, ic = 0.f;
        if (fz.wrec) fz.wrec[ri] += fz.wrec_scale * (NF == 2 ? uc[0] + uc[1] : uc[0]);
#pragma unroll
        for (int f = 0; f < NF; f++) {
            if (fz.snap[f]) fz.snap[f][ri] = uc[f];
            e2 += uc[f] * uc[f];
            if (fz.grad) {
                if (fz.ic == JUDI_IC_AS) ic += fz.usnap[f][ri] * delta[f];
                else {
                    // isic / fwi (sensitivity.py:106-122): u v.dt2 m +- grad u . grad v per field
                    float ics = fz.ic == JUDI_IC_FWI ? -1.f : 1.f;
                    float ig = inner_grad_snap<NDIM, R>(g, cf, fz.reg, fz.usnap[f], a.u[f], x, y, z);
                    ic += fz.usnap[f][ri] * delta[f] * mm + ics * cf.dt2 * ig;
                }
            }
        }
        if (fz.illum) fz.illum[ri] += cf.dt * e2;
        if (fz.grad) fz.grad[ri] -= fz.gcoef * ir * ic;
    }
}

// ------------------------------------------------------------------------------------------
// Visco-acoustic SLS physics (src/pysource/kernels.py:80-141), one point per thread.  Pressure p,
// memory variable r (updated in place).  With b = irho, mb = m b, dl = damp (the layer profile;
// the reference's mask-type field is 1 - dl), L(.) = laplacian(., b) (FD_utils.py:11-21):
//   forward  r+ = (1 - dl) (r + dt ((tt/ts) L(p)/b - r/ts))
//            p+ = p + (p - p-) + [dt^2 ((1 + tt) L(p) - b r+ + q) - dt dl (p - p-)] / (mb + dt dl)
//   adjoint  r- = (1 - dl) (r - dt (b p + r/ts)),   w = (1 + tt) p + (tt / (b ts)) r-
//            p- = p + (p - p+) + [dt^2 (L(w) + q) - dt dl (p - p+)] / (mb + dt dl)
// (the increment form of solve(.) with the one-sided u.dt of SURVEY A.3; forward and adjoint are
// exact discrete adjoints of each other: the dot test holds to 1e-15 in fp64.)
struct ViscoArgs {
    Grid g;
    Coefs cf;
    const float *p;          // level t
    float *pp;               // level t-+1 in, level t+-1 out
    float *r;                // memory variable
    float *w;                // adjoint: product field (pre pass writes, update reads)
    const float *P[3];       // flux form: b grad(p or w); else nullptr and the Laplacian of Lin is taken
    const float *Lin;
    const float *m, *irho, *tt, *ts;
    float irho_c;
    const float *dx, *dy, *dz;
    int adj;
    const float *qext;
    float qext_scale;
    const float *dd_in, *dm;
    float *dd_out;
    int born_ic;             // imaging condition of the Born source (linearised pass)
    const float *ub;         // isic / fwi Born source: background pressure, level t
    Fuse f;
};

template <int NDIM>
__global__ void __launch_bounds__(256) visco_pre_adjoint(ViscoArgs a)
{
    const Grid &g = a.g;
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y, z;
    if (NDIM == 3) { y = blockIdx.y * blockDim.y + threadIdx.y; z = blockIdx.z; }
    else { y = 0; z = blockIdx.y * blockDim.y + threadIdx.y; }
    if (x >= g.n[0] || y >= g.n[1] || z >= g.n[2]) return;
    long long i = g.at(x, y, z);
    const float b = a.irho ? a.irho[i] : a.irho_c;
    const float dl = (a.dx[x] + a.dy[y]) + a.dz[z];
    const float tt = a.tt[i], ts = a.ts[i];
    const float pc = a.p[i], rc = a.r[i];
    const float rn = (1.0f - dl) * (rc - a.cf.dt * (b * pc + rc / ts));
    a.r[i] = rn;
    a.w[i] = (1.0f + tt) * pc + (tt / (b * ts)) * rn;
}

template <int NDIM, int R>
__global__ void __launch_bounds__(256) visco_update(ViscoArgs a)
{
    const Grid &g = a.g;
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y, z;
    if (NDIM == 3) { y = blockIdx.y * blockDim.y + threadIdx.y; z = blockIdx.z; }
    else { y = 0; z = blockIdx.y * blockDim.y + threadIdx.y; }
    if (x >= g.n[0] || y >= g.n[1] || z >= g.n[2]) return;
    long long i = g.at(x, y, z);
    const Coefs &cf = a.cf;
    const float b = a.irho ? a.irho[i] : a.irho_c;
    const float mb = a.m[i] * b;
    const float dl = (a.dx[x] + a.dy[y]) + a.dz[z];
    const float sd = cf.dt * dl;
    float L2;   // dt^2 * laplacian(Lin, b)
    if (a.P[0]) {
        float L = dminus<R>(a.P[0], i, 1, cf.wx) + dminus<R>(a.P[2], i, g.sz, cf.wz);
        if (NDIM == 3) L += dminus<R>(a.P[1], i, g.sy, cf.wy);
        L2 = cf.dt2 * L;
    } else L2 = b * lap_at<NDIM, R>(g, cf, a.Lin, i);
    float S2;   // dt^2 * (right-hand side without the time-derivative terms)
    if (!a.adj) {
        const float tt = a.tt[i], ts = a.ts[i];
        const float rc = a.r[i];
        const float rn = (1.0f - dl) * (rc + ((tt / ts) * (L2 / b) * cf.idt - cf.dt * (rc / ts)));
        a.r[i] = rn;
        S2 = (1.0f + tt) * L2 - cf.dt2 * (b * rn);
    } else S2 = L2;
    if (a.dd_in) {
        // dt^2 * lin_src of the background pressure (sensitivity.py:174-209)
        if (a.born_ic == JUDI_IC_AS) S2 += -a.dm[i] * b * a.dd_in[i];
        else {
            const float ics = a.born_ic == JUDI_IC_FWI ? -1.f : 1.f;
            const float lw = a.irho ? lap_weighted2<NDIM, R>(g, cf, a.ub, a.dm, a.irho, i)
                                    : b * lap_weighted<NDIM, R>(g, cf, a.ub, a.dm, i);
            S2 += -(a.dm[i] * b * a.m[i] * a.dd_in[i] - ics * cf.dt2 * lw);
        }
    }
    if (a.qext) S2 += a.qext_scale * a.qext[i];
    const float pc = a.p[i];
    const float d1 = pc - a.pp[i];
    const float delta = (S2 - sd * d1) / (mb + sd);
    a.pp[i] = (pc + d1) + delta;
    if (a.dd_out) a.dd_out[i] = delta;
    const Fuse &fz = a.f;
    if ((fz.snap[0] || fz.grad || fz.illum || fz.wrec) && fz.reg.has(x, y, z)) {
        long long ri = fz.reg.at(x, y, z);
        if (fz.wrec) fz.wrec[ri] += fz.wrec_scale * pc;
        if (fz.snap[0]) fz.snap[0][ri] = pc;
        if (fz.illum) fz.illum[ri] += cf.dt * pc * pc;
        if (fz.grad) {
            const float us = fz.usnap[0][ri];
            if (fz.ic == JUDI_IC_AS) fz.grad[ri] -= fz.gcoef * b * us * delta;
            else {
                const float ics = fz.ic == JUDI_IC_FWI ? -1.f : 1.f;
                const float ig = inner_grad_snap<NDIM, R>(g, cf, fz.reg, fz.usnap[0], a.p, x, y, z);
                fz.grad[ri] -= fz.gcoef * b * (us * delta * a.m[i] + ics * cf.dt2 * ig);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
void tma3d_encode(const judi_model *M, const float *base, CUtensorMap *halo_map, CUtensorMap *tile_map);
bool tma3d_supported(const judi_model *M);
bool tma3d_fuse_supported(const StepArgs &s);
void tma3d_step(const judi_model *M, const StepArgs &a);
bool tma3d_step_born(const judi_model *M, const StepArgs &a);
bool flux3d_step(const judi_model *M, const StepArgs &s);
bool visco3d_step(const judi_model *M, const StepArgs &s);
bool tti3d_supported(const judi_model *M);
void tti3d_encode(const judi_model *M, const float *base, CUtensorMap *halo_map, CUtensorMap *tile_map);
void tti3d_encode_coef(const judi_model *M, CUtensorMap *maps);
bool tti3d_step(const judi_model *M, const StepArgs &s);
bool varc3d_supported(const judi_model *M);
void varc3d_encode(const judi_model *M, const float *base, CUtensorMap *halo_map, CUtensorMap *tile_map);
void varc3d_encode_coef(const judi_model *M, const float *base, CUtensorMap *map);
bool varc3d_step(const judi_model *M, const StepArgs &s);
bool isic3d_supported(const judi_model *M);
void isic3d_apply(const judi_model *M, const Fuse &f, const float *v, const float *dd, float *usg);
bool varc3d_lapw_supported(const judi_model *M);
void varc3d_lapw(const judi_model *M, const CUtensorMap &umap, const CUtensorMap &cmap, float *out);

void Wavefield::init(const judi_model *M, bool born_scratch)
{
    ctx = M->ctx;
    nf = M->tti ? 2 : 1;
    for (int f = 0; f < nf; f++)
        for (int s = 0; s < 2; s++) b[f][s] = DevBuf(ctx, (size_t)M->g.nalloc, true);
    cur = 0;
    bool ff = M->tti || M->irho_field;
    if (M->visco) {
        rmem = DevBuf(ctx, (size_t)M->g.nalloc, true);
        wtmp = DevBuf(ctx, (size_t)M->g.nalloc, true);
        if (born_scratch && !ff) dd[0] = DevBuf(ctx, (size_t)M->g.nalloc, true);
    }
    if (ff) {
        int nflux = M->tti ? 6 : 3;
        for (int k = 0; k < nflux; k++) flux[k] = DevBuf(ctx, (size_t)M->g.nalloc, true);
        if (born_scratch)
            for (int f = 0; f < nf; f++) dd[f] = DevBuf(ctx, (size_t)M->g.nalloc, true);
    } else if (born_scratch && tma3d_supported(M)) {
        dd[0] = DevBuf(ctx, (size_t)M->g.nalloc, true);  // two-launch Born of the TMA kernel
    }
    has_tmap = false;
    if (tma3d_supported(M)) {
        for (int f = 0; f < nf; f++)
            for (int s = 0; s < 2; s++) tma3d_encode(M, b[f][s].p, &tmap_halo[f][s], &tmap_tile[f][s]);
        tma3d_encode(M, M->m.p, nullptr, &tmap_m);
        if (M->dm.p) tma3d_encode(M, M->dm_src(), nullptr, &tmap_dm);
        has_tmap = true;
        if (born_scratch && M->dm.p && varc3d_lapw_supported(M)) {
            // isic / fwi Born source on the single-pass flux machinery (varc3d_lapw)
            lapw = DevBuf(ctx, (size_t)M->g.nalloc, true);
            for (int s = 0; s < 2; s++) varc3d_encode(M, b[0][s].p, &tmap_halo_v[s], nullptr);
            varc3d_encode_coef(M, M->dm.p, &tmap_dmc);
            has_lapw = true;
        }
    } else if (varc3d_supported(M)) {
        for (int s = 0; s < 2; s++) varc3d_encode(M, b[0][s].p, &tmap_halo[0][s], &tmap_tile[0][s]);
        varc3d_encode(M, M->m.p, nullptr, &tmap_m);
        varc3d_encode(M, M->irho.p, nullptr, &tmap_irho);
        varc3d_encode_coef(M, M->irho.p, &tmap_coef);
        has_tmap = true;
    } else if (tti3d_supported(M)) {
        for (int f = 0; f < nf; f++)
            for (int s = 0; s < 2; s++) tti3d_encode(M, b[f][s].p, &tmap_halo[f][s], &tmap_tile[f][s]);
        tti3d_encode(M, M->irho_field ? M->tti_mi.p : M->m.p, nullptr, &tmap_m);
        tti3d_encode_coef(M, tmap_par);
        has_tmap = true;
    }
}

template <int NDIM, int R>
static void launch_acoustic_generic(const judi_model *M, const StepArgs &s)
{
    judi_ctx *ctx = M->ctx;
    AcArgs a;
    memset(&a, 0, sizeof(a));
    a.g = M->g;
    a.cf = M->cf;
    a.u = s.u[0];
    a.up = s.up[0];
    a.m = M->m.p;
    a.dx = M->damp[0].p; a.dy = M->damp[1].p; a.dz = M->damp[2].p;
    a.sdamp = M->dt / M->irho_c;
    a.irho_c = M->irho_c;
    if (s.born) { a.ul = s.ul[0]; a.ulp = s.ulp[0]; a.dm = M->dm.p; a.dms = M->dm_src(); a.born_ic = s.born_ic; }
    a.qext = s.qext;
    a.qext_scale = s.qext_scale;
    a.f = s.f;
    dim3 blk(64, 4, 1), grd;
    if (NDIM == 3) grd = dim3((M->N[0] + 63) / 64, (M->N[1] + 3) / 4, M->N[2]);
    else grd = dim3((M->N[0] + 63) / 64, (M->N[2] + 3) / 4, 1);
    acoustic_generic<NDIM, R><<<grd, blk, 0, ctx->stream>>>(a);
    launch_count(ctx, "acoustic_generic");
}

template <int NDIM, int R>
static void flux_one(const judi_model *M, Wavefield *W, const float *const *u, float *const *up,
                     float *const *dd_out, const float *const *dd_in, const Fuse &fuse,
                     const float *qext = nullptr, float qext_scale = 0.f, int born_ic = JUDI_IC_AS,
                     const float *const *ub = nullptr)
{
    judi_ctx *ctx = M->ctx;
    int nf = M->tti ? 2 : 1;
    FluxArgs fa;
    memset(&fa, 0, sizeof(fa));
    fa.g = M->g; fa.cf = M->cf;
    fa.u1 = u[0]; fa.u2 = nf == 2 ? u[1] : nullptr;
    fa.irho = M->irho.p; fa.eps = M->eps.p; fa.delta = M->delta.p;
    fa.theta = M->theta.p; fa.phi = M->phi.p;
    for (int d = 0; d < 3; d++) { fa.P[d] = W->flux[d].p; fa.M[d] = W->flux[3 + d].p; }
    fa.tti = M->tti;
    dim3 blk(64, 4, 1), grd;
    if (NDIM == 3)
        grd = dim3((M->N[0] + 2 * R + 63) / 64, (M->N[1] + 2 * R + 3) / 4, M->N[2] + 2 * R);
    else grd = dim3((M->N[0] + 2 * R + 63) / 64, (M->N[2] + 2 * R + 3) / 4, 1);
    flux_kernel<NDIM, R><<<grd, blk, 0, ctx->stream>>>(fa);
    launch_count(ctx, "flux_kernel");
    FluxUpdArgs ua;
    memset(&ua, 0, sizeof(ua));
    ua.g = M->g; ua.cf = M->cf;
    for (int f = 0; f < nf; f++) {
        ua.u[f] = u[f]; ua.up[f] = up[f];
        ua.dd_out[f] = dd_out ? dd_out[f] : nullptr;
        ua.dd_in[f] = dd_in ? dd_in[f] : nullptr;
        ua.ub[f] = ub ? ub[f] : nullptr;
    }
    ua.born_ic = born_ic;
    for (int d = 0; d < 3; d++) { ua.P[d] = W->flux[d].p; ua.M[d] = W->flux[3 + d].p; }
    ua.m = M->m.p; ua.irho = M->irho.p; ua.dm = M->dm.p; ua.dms = M->dm_src();
    ua.dx = M->damp[0].p; ua.dy = M->damp[1].p; ua.dz = M->damp[2].p;
    ua.qext = qext;
    ua.qext_scale = qext_scale;
    ua.f = fuse;
    if (NDIM == 3) grd = dim3((M->N[0] + 63) / 64, (M->N[1] + 3) / 4, M->N[2]);
    else grd = dim3((M->N[0] + 63) / 64, (M->N[2] + 3) / 4, 1);
    if (nf == 2) flux_update<NDIM, R, 2><<<grd, blk, 0, ctx->stream>>>(ua);
    else flux_update<NDIM, R, 1><<<grd, blk, 0, ctx->stream>>>(ua);
    launch_count(ctx, "flux_update");
}

template <int NDIM, int R>
static void step_flux(const judi_model *M, const StepArgs &s)
{
    Wavefield *W = s.scratch;
    JUDI_REQUIRE(W != nullptr, "flux scratch not bound");
    if (!s.born) {
        flux_one<NDIM, R>(M, W, s.u, s.up, nullptr, nullptr, s.f, s.qext, s.qext_scale);
        return;
    }
    float *dd[2] = {W->dd[0].p, W->dd[1].p};
    const float *ddc[2] = {W->dd[0].p, W->dd[1].p};
    flux_one<NDIM, R>(M, W, s.u, s.up, dd, nullptr, s.f, s.qext, s.qext_scale);
    Fuse none;
    memset(&none, 0, sizeof(none));
    // the background pass wrote level t+1 over level t-1: s.u still holds level t
    flux_one<NDIM, R>(M, W, s.ul, s.ulp, nullptr, ddc, none, nullptr, 0.f, s.born_ic, s.u);
}

// One visco-acoustic step (forward or adjoint, plain or Born): flux pass (density field) + update.
template <int NDIM, int R>
static void visco_one(const judi_model *M, Wavefield *W, const float *p, float *pp, float *r, int adj,
                      const float *dd_in, float *dd_out, const Fuse &fuse, const float *qext,
                      float qext_scale, int born_ic = JUDI_IC_AS, const float *ub = nullptr)
{
    judi_ctx *ctx = M->ctx;
    ViscoArgs a;
    memset(&a, 0, sizeof(a));
    a.g = M->g; a.cf = M->cf;
    a.p = p; a.pp = pp; a.r = r; a.w = W->wtmp.p;
    a.m = M->m.p; a.irho = M->irho_field ? M->irho.p : nullptr; a.irho_c = M->irho_c;
    a.tt = M->vtt.p; a.ts = M->vts.p;
    a.dx = M->damp[0].p; a.dy = M->damp[1].p; a.dz = M->damp[2].p;
    a.adj = adj;
    a.qext = qext; a.qext_scale = qext_scale;
    a.dd_in = dd_in; a.dm = M->dm.p; a.dd_out = dd_out;
    a.born_ic = born_ic; a.ub = ub;
    a.f = fuse;
    dim3 blk(64, 4, 1), grd;
    if (NDIM == 3) grd = dim3((M->N[0] + 63) / 64, (M->N[1] + 3) / 4, M->N[2]);
    else grd = dim3((M->N[0] + 63) / 64, (M->N[2] + 3) / 4, 1);
    if (adj) {
        visco_pre_adjoint<NDIM><<<grd, blk, 0, ctx->stream>>>(a);
        launch_count(ctx, "visco_pre_adjoint");
    }
    a.Lin = adj ? W->wtmp.p : p;
    if (M->irho_field) {
        FluxArgs fa;
        memset(&fa, 0, sizeof(fa));
        fa.g = M->g; fa.cf = M->cf;
        fa.u1 = a.Lin;
        fa.irho = M->irho.p;
        for (int d = 0; d < 3; d++) { fa.P[d] = W->flux[d].p; a.P[d] = W->flux[d].p; }
        dim3 gf;
        if (NDIM == 3) gf = dim3((M->N[0] + 2 * R + 63) / 64, (M->N[1] + 2 * R + 3) / 4, M->N[2] + 2 * R);
        else gf = dim3((M->N[0] + 2 * R + 63) / 64, (M->N[2] + 2 * R + 3) / 4, 1);
        flux_kernel<NDIM, R><<<gf, blk, 0, ctx->stream>>>(fa);
        launch_count(ctx, "flux_kernel");
    }
    visco_update<NDIM, R><<<grd, blk, 0, ctx->stream>>>(a);
    launch_count(ctx, "visco_update");
}

template <int NDIM, int R>
static void step_visco(const judi_model *M, const StepArgs &s)
{
    Wavefield *W = s.scratch;
    JUDI_REQUIRE(W != nullptr && W->rmem.p, "visco-acoustic scratch not bound");
    if (!s.born) {
        visco_one<NDIM, R>(M, W, s.u[0], s.up[0], W->rmem.p, s.adj, nullptr, nullptr, s.f, s.qext, s.qext_scale);
        return;
    }
    JUDI_REQUIRE(s.lin && s.lin->rmem.p && W->dd[0].p, "visco-acoustic Born scratch not bound");
    visco_one<NDIM, R>(M, W, s.u[0], s.up[0], W->rmem.p, 0, nullptr, W->dd[0].p, s.f, s.qext, s.qext_scale);
    Fuse none;
    memset(&none, 0, sizeof(none));
    // the background pass wrote level t+1 over level t-1: s.u[0] still holds level t
    visco_one<NDIM, R>(M, W, s.ul[0], s.ulp[0], s.lin->rmem.p, 0, W->dd[0].p, nullptr, none, nullptr, 0.f,
                       s.born_ic, s.u[0]);
}

template <int NDIM>
static void step_ndim(const judi_model *M, const StepArgs &s)
{
    bool ff = M->tti || M->irho_field;
#define DISPATCH_R(RR)                                                              \
    case RR:                                                                        \
        if (M->visco) step_visco<NDIM, RR>(M, s);                                   \
        else if (ff) step_flux<NDIM, RR>(M, s);                                     \
        else launch_acoustic_generic<NDIM, RR>(M, s);                               \
        break;
    switch (M->r) {
        DISPATCH_R(1)
        DISPATCH_R(2)
        DISPATCH_R(4)
        DISPATCH_R(6)
        DISPATCH_R(8)
    default:
        throw JudiError("unsupported space order");
    }
#undef DISPATCH_R
}

bool flux_born_single_pass(const judi_model *M)
{
    return M->ctx->opt_kernel3d == 1 && (tti3d_supported(M) || varc3d_supported(M));
}

// One leapfrog step of every field of the model (+ fused extras), timed with CUDA events when
// the context asks for it.
void stencil_step(const judi_model *M, const StepArgs &s)
{
    judi_ctx *ctx = M->ctx;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    // timing = k > 0: bracket every k-th stencil launch with CUDA events (k = 1: all of them).
    // Event records cost ~4 us of stream time each, so long runs sample (k coprime with t_sub
    // visits every launch flavour in proportion).
    const bool timed = ctx->timing > 0 && (ctx->timing_seq++ % ctx->timing) == 0;
    if (timed) {
        auto get = [&]() {
            cudaEvent_t e;
            if (!ctx->free_events.empty()) { e = ctx->free_events.back(); ctx->free_events.pop_back(); }
            else CUDA_CHECK(cudaEventCreate(&e));
            return e;
        };
        e0 = get(); e1 = get();
        CUDA_CHECK(cudaEventRecord(e0, ctx->stream));
    }
    bool ff = M->tti || M->irho_field;
    bool use_tma = M->ndim == 3 && !ff && !M->visco && ctx->opt_kernel3d == 1 && tma3d_supported(M) &&
                   (s.f.ic == JUDI_IC_AS || !s.f.grad) && s.tmap != nullptr && tma3d_fuse_supported(s) &&
                   (!s.born || (s.lin && s.scratch && s.scratch->dd[0].p &&
                                (s.born_ic == JUDI_IC_AS || (s.scratch->has_lapw && !s.qext))));
    bool born_done = false;
    // isic / fwi imaging step of the 3-D acoustic gradient: TMA step that also writes delta, then
    // one vectorised imaging pass (isic3d.cu) instead of the one-point-per-thread kernel
    const bool isic_fast = M->ndim == 3 && !ff && !M->visco && ctx->opt_kernel3d == 1 && tma3d_supported(M) &&
                           isic3d_supported(M) && s.f.grad && s.f.ic != JUDI_IC_AS && !s.born && !s.qext &&
                           !s.f.wrec && !s.f.snap[0] && s.tmap != nullptr && s.scratch != nullptr &&
                           ctx->opt_isic_fast;
    if (isic_fast) {
        Wavefield *W = s.scratch;
        if (!W->dd[0].p) W->dd[0] = DevBuf(ctx, (size_t)M->g.nalloc, true);
        if (!W->lapw.p) W->lapw = DevBuf(ctx, (size_t)M->g.nalloc, true);
        StepArgs a1 = s;
        a1.f.grad = nullptr;
        a1.f.usnap[0] = nullptr;
        a1.dd_out = W->dd[0].p;
        tma3d_step(M, a1);
        // s.u[0] is still level t (the step wrote the new level over level t-+1)
        isic3d_apply(M, s.f, s.u[0], W->dd[0].p, W->lapw.p);
        born_done = true;
    } else
    if (use_tma && s.born && ctx->opt_born_fused && s.born_ic == JUDI_IC_AS) {
        // fused Born (32 B/pt): both fields in one launch, delta handed over in shared memory
        StepArgs a = s;
        a.born = 0;
        a.tmap2 = &s.lin->tmap_halo[0][s.lin->cur];
        a.tmap_prev2 = &s.lin->tmap_tile[0][1 - s.lin->cur];
        a.tmap_dm = &s.lin->tmap_dm;
        born_done = tma3d_step_born(M, a);
    }
    if (born_done) { /* one launch */ }
    else if (use_tma && s.born) {
        // Born as two launches of the TMA kernel (44 B/pt): the background step also writes
        // delta = u+ - 2u + u-, the linearised step reads it as the source -dm * delta
        StepArgs a1 = s;
        a1.born = 0;
        a1.dd_out = s.scratch->dd[0].p;
        tma3d_step(M, a1);
        StepArgs a2 = s;
        memset(&a2.f, 0, sizeof(a2.f));
        a2.born = 0;
        a2.qext = nullptr;      // the space-time source drives the background field only
        a2.u[0] = s.ul[0];
        a2.up[0] = s.ulp[0];
        a2.tmap = &s.lin->tmap_halo[0][s.lin->cur];
        a2.tmap_prev = &s.lin->tmap_tile[0][1 - s.lin->cur];
        a2.qsrc = s.scratch->dd[0].p;
        if (s.born_ic != JUDI_IC_AS) {
            // isic / fwi: + ics dt^2 div(dm grad u(t)) from one single-pass flux launch on u(t)
            Wavefield *W = s.scratch;
            varc3d_lapw(M, W->tmap_halo_v[W->cur], W->tmap_dmc, W->lapw.p);
            a2.sfield = W->lapw.p;
        }
        tma3d_step(M, a2);
    } else if (use_tma) tma3d_step(M, s);
    else if (M->ndim == 3 && ff && !M->visco && ctx->opt_kernel3d == 1 && !(s.scratch && s.scratch->generic_only) &&
             !(s.scratch && s.scratch->two_pass_only) && (tti3d_step(M, s) || varc3d_step(M, s))) { /* single pass */ }
    else if (M->ndim == 3 && ff && !M->visco && ctx->opt_kernel3d == 1 && !(s.scratch && s.scratch->generic_only) &&
             flux3d_step(M, s)) { /* vectorised */ }
    else if (M->ndim == 3 && M->visco && ctx->opt_kernel3d == 1 && visco3d_step(M, s)) { /* vectorised */ }
    else if (M->ndim == 3) step_ndim<3>(M, s);
    else step_ndim<2>(M, s);
    if (timed) {
        CUDA_CHECK(cudaEventRecord(e1, ctx->stream));
        int kind = s.born ? JUDI_KIND_BORN : s.f.grad ? JUDI_KIND_IC : s.f.snap[0] ? JUDI_KIND_SAVE
                                                     : JUDI_KIND_PLAIN;
        ctx->pending.push_back({e0, e1, kind});
        ctx->stencil_launches++;
        ctx->stencil_points += (double)M->N[0] * M->N[1] * M->N[2] * (s.born ? 2 : 1);
        if (ctx->pending.size() >= 4096) ctx->flush_timing();
    }
}
