// flux3d.cu -- vectorised 3-D kernels of the div(grad) forms: variable density
// (src/pysource/FD_utils.py:11-21) and the TTI rotated self-adjoint operator
// (FD_utils.py:24-105, kernels.py:144-182).  Two passes per time step:
//   pass 1  flux3d_vec:    P = R^T (A R G u1 + B R G u2), M = R^T (B R G u1 + C R G u2)
//                          (or P = irho * G u) on the grid extended by R nodes
//   pass 2  fluxdiv3d_vec: H = sum_d D-_d P_d, time update of 1 or 2 fields + fused extras
// Every thread owns 4 consecutive x points; all global accesses are 128-bit and aligned (x is a
// multiple of 4 relative to the 128-byte aligned grid origin).
//
// The rotation is applied in closed form.  With r = third row of R = rot_y(theta) rot_z(phi)
// = (sin(theta) cos(phi), sin(theta) sin(phi), cos(theta)) (the symmetry axis),
//     R^T diag(1,1,0) R = I - r r^T,   R^T diag(0,0,1) R = r r^T,
// hence (A = b diag(d,d,1), B = b sqrt((e-d) d) diag(1,1,0), C = b (e-d) diag(1,1,0)):
//     P = A0 g1 + B0 g2 + [(b - A0)(r.g1) - B0 (r.g2)] r
//     M = B0 g1 + C0 g2 - [ B0 (r.g1)     + C0 (r.g2)] r
// which needs the unit vector r (precomputed once per model, model.cu) instead of two sincos and
// two 3x3 rotations per point and step (same operator as the one-point-per-thread kernels in
// kernels.cu, which remain the 2-D path, up to rounding).  The algebra is packed fp32x2.
// (A z-marching variant with the z-lines in register queues was measured slower, 1.42 vs
// 1.15 ms/step: 237 registers -> 8 warps/SM cannot hide the latency of synchronous loads.)
#include "common.cuh"

struct Flux3Args {
    Grid g;
    Coefs cf;
    const float *u1, *u2;
    const float *irho, *eps, *delta, *r3[3];
    float irho_c;
    float *P[3], *M[3];
    int z0, z1;   // flux planes [z0, z1) of this launch
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

template <int R, bool TTI, bool IRHOF, int MINB = 8>
__global__ void __launch_bounds__(128, MINB) flux3d_vec(Flux3Args a)
{
    constexpr int XE = ((R + 3) / 4) * 4;
    const Grid &g = a.g;
    const int x = -XE + 4 * (blockIdx.x * blockDim.x + threadIdx.x);
    const int y = (int)(blockIdx.y * blockDim.y + threadIdx.y) - R;
    const int z = a.z0 + (int)(blockIdx.z * blockDim.z + threadIdx.z);
    if (x >= g.n[0] + R || y >= g.n[1] + R || z >= a.z1) return;
    const long long i = g.at(x, y, z);
    const Coefs &cf = a.cf;
    float gx[4], gy[4], gz[4], b[4];
    grad4<R, XE>(g, cf, a.u1, i, gx, gy, gz);
    if (IRHOF) ld4(a.irho + i, b);
    else b[0] = b[1] = b[2] = b[3] = a.irho_c;
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
    float hx[4], hy[4], hz[4], ep[4], dl[4], rx[4], ry[4], rz[4];
    grad4<R, XE>(g, cf, a.u2, i, hx, hy, hz);
    ld4(a.eps + i, ep);
    ld4(a.delta + i, dl);
    ld4(a.r3[0] + i, rx);
    ld4(a.r3[1] + i, ry);
    ld4(a.r3[2] + i, rz);
    float Px[4], Py[4], Pz[4], Mx[4], My[4], Mz[4];
#pragma unroll
    for (int h = 0; h < 4; h += 2) {
#define P2(v) pk(v[h], v[h + 1])
        const f2 bb = P2(b), e2 = P2(ep), d2_ = P2(dl), r_x = P2(rx), r_y = P2(ry), r_z = P2(rz);
        const f2 g1x = P2(gx), g1y = P2(gy), g1z = P2(gz), g2x = P2(hx), g2y = P2(hy), g2z = P2(hz);
#undef P2
        const f2 emd = sub2(e2, d2_);
        const f2 A0 = mul2(bb, d2_);
        const f2 pr = mul2(emd, d2_);
        const f2 B0 = mul2(bb, pk(sqrtf(pr.x), sqrtf(pr.y)));
        const f2 C0 = mul2(bb, emd);
        const f2 d1 = fma2(r_z, g1z, fma2(r_y, g1y, mul2(r_x, g1x)));
        const f2 d2 = fma2(r_z, g2z, fma2(r_y, g2y, mul2(r_x, g2x)));
        // sP = (b - A0) d1 - B0 d2,   sM = -(B0 d1 + C0 d2)
        const f2 nB0 = mul2s(-1.f, B0);
        const f2 sP = fma2(nB0, d2, mul2(sub2(bb, A0), d1));
        const f2 sM = fma2(nB0, d1, mul2(mul2s(-1.f, C0), d2));
        const f2 px = fma2(sP, r_x, fma2(B0, g2x, mul2(A0, g1x)));
        const f2 py = fma2(sP, r_y, fma2(B0, g2y, mul2(A0, g1y)));
        const f2 pz = fma2(sP, r_z, fma2(B0, g2z, mul2(A0, g1z)));
        const f2 mx = fma2(sM, r_x, fma2(C0, g2x, mul2(B0, g1x)));
        const f2 my = fma2(sM, r_y, fma2(C0, g2y, mul2(B0, g1y)));
        const f2 mz = fma2(sM, r_z, fma2(C0, g2z, mul2(B0, g1z)));
        Px[h] = px.x; Px[h + 1] = px.y; Py[h] = py.x; Py[h + 1] = py.y; Pz[h] = pz.x; Pz[h + 1] = pz.y;
        Mx[h] = mx.x; Mx[h + 1] = mx.y; My[h] = my.x; My[h + 1] = my.y; Mz[h] = mz.x; Mz[h + 1] = mz.y;
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
    float irho_c;
    const float *dx, *dy, *dz;
    float *dd_out[2];
    const float *dd_in[2];
    const float *qext;
    float qext_scale;
    Fuse f;
    int z0, z1;   // output planes [z0, z1) of this launch
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

template <int R, int NF, bool IRHOF, int MINB = 8>
__global__ void __launch_bounds__(128, MINB) fluxdiv3d_vec(FluxDiv3Args a)
{
    constexpr int XE = ((R + 3) / 4) * 4;
    const Grid &g = a.g;
    const int x = 4 * (blockIdx.x * blockDim.x + threadIdx.x);
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int z = a.z0 + (int)(blockIdx.z * blockDim.z + threadIdx.z);
    if (x >= g.n[0] || y >= g.n[1] || z >= a.z1) return;
    const long long i = g.at(x, y, z);
    const Coefs &cf = a.cf;
    float ir[4], mm[4], dxv[4], rden[4], sd[4];
    if (IRHOF) ld4(a.irho + i, ir);
    else ir[0] = ir[1] = ir[2] = ir[3] = a.irho_c;
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
// Visco-acoustic SLS step in 3-D (src/pysource/kernels.py:80-141; the one-point-per-thread form
// and the equations are in kernels.cu), four x-points per thread with 128-bit accesses.  The
// Laplacian is either the divergence of the fluxes written by flux3d_vec (density field) or the
// centred stencil times the constant irho.
struct Visco3Args {
    Grid g;
    Coefs cf;
    const float *p;          // level t
    float *pp;               // level t-+1 in, level t+-1 out
    float *r;                // memory variable
    float *w;                // adjoint: product field written by the pre pass
    const float *P[3];       // flux form: irho grad(p or w)
    const float *Lin;        // constant density: field the Laplacian is taken of
    const float *m, *irho, *tt, *ts;
    float irho_c;
    const float *dx, *dy, *dz;
    int adj;
    const float *qext;
    float qext_scale;
    const float *dd_in, *dm;
    float *dd_out;
    Fuse f;
};

__global__ void __launch_bounds__(128) visco3d_pre_adjoint_vec(Visco3Args a)
{
    const Grid &g = a.g;
    const int x = 4 * (blockIdx.x * blockDim.x + threadIdx.x);
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int z = blockIdx.z * blockDim.z + threadIdx.z;
    if (x >= g.n[0] || y >= g.n[1] || z >= g.n[2]) return;
    const long long i = g.at(x, y, z);
    float pc[4], rc[4], tt[4], ts[4], b[4], dxv[4], rn[4], w[4];
    ld4(a.p + i, pc); ld4(a.r + i, rc); ld4(a.tt + i, tt); ld4(a.ts + i, ts); ld4(a.dx + x, dxv);
    if (a.irho) ld4(a.irho + i, b);
    else b[0] = b[1] = b[2] = b[3] = a.irho_c;
    const float dyz = a.dy[y], dzv = a.dz[z];
#pragma unroll
    for (int e = 0; e < 4; e++) {
        const float dl = (dxv[e] + dyz) + dzv;
        rn[e] = (1.0f - dl) * (rc[e] - a.cf.dt * (b[e] * pc[e] + rc[e] / ts[e]));
        w[e] = (1.0f + tt[e]) * pc[e] + (tt[e] / (b[e] * ts[e])) * rn[e];
    }
    // the halo of r and w stays zero: only domain points are written
    if (x + 3 < g.n[0]) { st4(a.r + i, rn); st4(a.w + i, w); }
    else
#pragma unroll
        for (int e = 0; e < 4; e++)
            if (x + e < g.n[0]) { a.r[i + e] = rn[e]; a.w[i + e] = w[e]; }
}

template <int R, bool FLUX>
__global__ void __launch_bounds__(128) visco3d_update_vec(Visco3Args a)
{
    constexpr int XE = ((R + 3) / 4) * 4;
    const Grid &g = a.g;
    const Coefs &cf = a.cf;
    const int x = 4 * (blockIdx.x * blockDim.x + threadIdx.x);
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int z = blockIdx.z * blockDim.z + threadIdx.z;
    if (x >= g.n[0] || y >= g.n[1] || z >= g.n[2]) return;
    const long long i = g.at(x, y, z);
    float b[4], L2[4];
    if (a.irho) ld4(a.irho + i, b);
    else b[0] = b[1] = b[2] = b[3] = a.irho_c;
    if (FLUX) {
        div4<R, XE>(g, cf, a.P, i, L2);
#pragma unroll
        for (int e = 0; e < 4; e++) L2[e] *= cf.dt2;
    } else {
        // same summation order as lap_at (kernels.cu): centre, z, x, y
        const float *u = a.Lin;
        float wv[4 + 2 * XE], c[4];
#pragma unroll
        for (int v = 0; v < 1 + 2 * (XE / 4); v++) ld4(u + i - XE + 4 * v, wv + 4 * v);
#pragma unroll
        for (int e = 0; e < 4; e++) c[e] = cf.c0 * wv[XE + e];
#pragma unroll
        for (int k = 1; k <= R; k++) {
            float pz[4], qz[4];
            ld4(u + i + k * g.sz, pz); ld4(u + i - k * g.sz, qz);
#pragma unroll
            for (int e = 0; e < 4; e++) c[e] += cf.cz[k] * (pz[e] + qz[e]);
        }
#pragma unroll
        for (int k = 1; k <= R; k++)
#pragma unroll
            for (int e = 0; e < 4; e++) c[e] += cf.cx[k] * (wv[XE + e + k] + wv[XE + e - k]);
#pragma unroll
        for (int k = 1; k <= R; k++) {
            float py[4], qy[4];
            ld4(u + i + k * g.sy, py); ld4(u + i - k * g.sy, qy);
#pragma unroll
            for (int e = 0; e < 4; e++) c[e] += cf.cy[k] * (py[e] + qy[e]);
        }
#pragma unroll
        for (int e = 0; e < 4; e++) L2[e] = b[e] * c[e];
    }
    float mm[4], dxv[4], pc[4], po[4], pn[4], delta[4], rn[4], rc[4], tt[4], ts[4];
    ld4(a.m + i, mm); ld4(a.dx + x, dxv); ld4(a.p + i, pc); ld4(a.pp + i, po);
    if (!a.adj) { ld4(a.r + i, rc); ld4(a.tt + i, tt); ld4(a.ts + i, ts); }
    float ddv[4] = {0.f, 0.f, 0.f, 0.f}, dmv[4] = {0.f, 0.f, 0.f, 0.f}, qe[4] = {0.f, 0.f, 0.f, 0.f};
    if (a.dd_in) { ld4(a.dd_in + i, ddv); ld4(a.dm + i, dmv); }
    if (a.qext) ld4(a.qext + i, qe);
    const float dyz = a.dy[y], dzv = a.dz[z];
#pragma unroll
    for (int e = 0; e < 4; e++) {
        const float dl = (dxv[e] + dyz) + dzv;
        const float sd = cf.dt * dl;
        float S2;
        if (!a.adj) {
            rn[e] = (1.0f - dl) * (rc[e] + ((tt[e] / ts[e]) * (L2[e] / b[e]) * cf.idt - cf.dt * (rc[e] / ts[e])));
            S2 = (1.0f + tt[e]) * L2[e] - cf.dt2 * (b[e] * rn[e]);
        } else S2 = L2[e];
        if (a.dd_in) S2 += -dmv[e] * b[e] * ddv[e];
        if (a.qext) S2 += a.qext_scale * qe[e];
        const float d1 = pc[e] - po[e];
        delta[e] = (S2 - sd * d1) / (mm[e] * b[e] + sd);
        pn[e] = (pc[e] + d1) + delta[e];
    }
    if (x + 3 < g.n[0]) {
        st4(a.pp + i, pn);
        if (!a.adj) st4(a.r + i, rn);
        if (a.dd_out) st4(a.dd_out + i, delta);
    } else {
#pragma unroll
        for (int e = 0; e < 4; e++)
            if (x + e < g.n[0]) {
                a.pp[i + e] = pn[e];
                if (!a.adj) a.r[i + e] = rn[e];
                if (a.dd_out) a.dd_out[i + e] = delta[e];
            }
    }
    const Fuse &fz = a.f;
    if ((fz.snap[0] || fz.grad || fz.illum || fz.wrec) && y >= fz.reg.lo[1] && y < fz.reg.hi[1] &&
        z >= fz.reg.lo[2] && z < fz.reg.hi[2]) {
        const long long ri = fz.reg.at(x, y, z);
#pragma unroll
        for (int e = 0; e < 4; e++) {
            if (x + e < fz.reg.lo[0] || x + e >= fz.reg.hi[0]) continue;
            if (fz.wrec) fz.wrec[ri + e] += fz.wrec_scale * pc[e];
            if (fz.snap[0]) fz.snap[0][ri + e] = pc[e];
            if (fz.illum) fz.illum[ri + e] += cf.dt * pc[e] * pc[e];
            if (fz.grad) fz.grad[ri + e] -= fz.gcoef * b[e] * fz.usnap[0][ri + e] * delta[e];
        }
    }
}

template <int R>
static void visco3d_one(const judi_model *M, Wavefield *W, const float *p, float *pp, float *r, int adj,
                        const float *dd_in, float *dd_out, const Fuse &fuse, const float *qext,
                        float qext_scale)
{
    constexpr int XE = ((R + 3) / 4) * 4;
    judi_ctx *ctx = M->ctx;
    Visco3Args a;
    memset(&a, 0, sizeof(a));
    a.g = M->g; a.cf = M->cf;
    a.p = p; a.pp = pp; a.r = r; a.w = W->wtmp.p;
    a.m = M->m.p; a.irho = M->irho_field ? M->irho.p : nullptr; a.irho_c = M->irho_c;
    a.tt = M->vtt.p; a.ts = M->vts.p;
    a.dx = M->damp[0].p; a.dy = M->damp[1].p; a.dz = M->damp[2].p;
    a.adj = adj;
    a.qext = qext; a.qext_scale = qext_scale;
    a.dd_in = dd_in; a.dm = M->dm.p; a.dd_out = dd_out;
    a.f = fuse;
    const dim3 blk(32, 4, 1);
    const dim3 grd(((M->N[0] + 3) / 4 + blk.x - 1) / blk.x, (M->N[1] + blk.y - 1) / blk.y, M->N[2]);
    if (adj) {
        visco3d_pre_adjoint_vec<<<grd, blk, 0, ctx->stream>>>(a);
        launch_count(ctx, "visco3d_pre_adjoint_vec");
    }
    a.Lin = adj ? W->wtmp.p : p;
    if (M->irho_field) {
        Flux3Args fa;
        memset(&fa, 0, sizeof(fa));
        fa.g = M->g; fa.cf = M->cf;
        fa.u1 = a.Lin;
        fa.irho = M->irho.p; fa.irho_c = M->irho_c;
        for (int d = 0; d < 3; d++) { fa.P[d] = W->flux[d].p; a.P[d] = W->flux[d].p; }
        fa.z0 = -R; fa.z1 = M->N[2] + R;
        const int nxt = (M->N[0] + R + XE + 3) / 4;
        const dim3 gf((nxt + blk.x - 1) / blk.x, (M->N[1] + 2 * R + blk.y - 1) / blk.y, M->N[2] + 2 * R);
        flux3d_vec<R, false, true><<<gf, blk, 0, ctx->stream>>>(fa);
        launch_count(ctx, "flux3d_vec");
        visco3d_update_vec<R, true><<<grd, blk, 0, ctx->stream>>>(a);
    } else visco3d_update_vec<R, false><<<grd, blk, 0, ctx->stream>>>(a);
    launch_count(ctx, "visco3d_update_vec");
}

// One 3-D visco-acoustic step with the vectorised kernels; false if the combination is not covered
// (isic / fwi conditions and sources, space orders without an instantiation): kernels.cu then runs it.
bool visco3d_step(const judi_model *M, const StepArgs &s)
{
    Wavefield *W = s.scratch;
    if (!M->visco || M->ndim != 3 || !W || !W->rmem.p) return false;
    if ((s.f.grad && s.f.ic != JUDI_IC_AS) || (s.born && s.born_ic != JUDI_IC_AS)) return false;
    if (s.born && (!s.lin || !s.lin->rmem.p || !W->dd[0].p)) return false;
    auto run = [&](auto rtag) {
        constexpr int R = decltype(rtag)::value;
        if (!s.born) {
            visco3d_one<R>(M, W, s.u[0], s.up[0], W->rmem.p, s.adj, nullptr, nullptr, s.f, s.qext, s.qext_scale);
            return;
        }
        visco3d_one<R>(M, W, s.u[0], s.up[0], W->rmem.p, 0, nullptr, W->dd[0].p, s.f, s.qext, s.qext_scale);
        Fuse none;
        memset(&none, 0, sizeof(none));
        visco3d_one<R>(M, W, s.ul[0], s.ulp[0], s.lin->rmem.p, 0, W->dd[0].p, nullptr, none, nullptr, 0.f);
    };
    switch (M->r) {
    case 2: run(std::integral_constant<int, 2>()); return true;
    case 4: run(std::integral_constant<int, 4>()); return true;
    case 8: run(std::integral_constant<int, 8>()); return true;
    default: return false;
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
    fa.irho = M->irho_field ? M->irho.p : nullptr;
    fa.irho_c = M->irho_c;
    fa.eps = M->eps.p; fa.delta = M->delta.p;
    for (int d = 0; d < 3; d++) {
        fa.r3[d] = M->r3[d].p;
        fa.P[d] = W->flux[d].p; fa.M[d] = W->flux[3 + d].p;
    }
    // 3-D thread blocks: the 2R-wide y- and z-lines of neighbouring threads overlap inside the CTA
    // and are served by the L1 instead of L2 (tunable: option "flux_block" = bx | by<<8 | bz<<16)
    int fb = ctx->opt_flux_block;
    dim3 blk(fb & 255, (fb >> 8) & 255, (fb >> 16) & 255);
    int nxt = (M->N[0] + R + XE + 3) / 4;  // threads along x covering [-XE, nx + R)
    dim3 grd((nxt + blk.x - 1) / blk.x, (M->N[1] + 2 * R + blk.y - 1) / blk.y, 1);
    FluxDiv3Args ua;
    memset(&ua, 0, sizeof(ua));
    ua.g = M->g; ua.cf = M->cf;
    for (int f = 0; f < nf; f++) {
        ua.u[f] = u[f]; ua.up[f] = up[f];
        ua.dd_out[f] = dd_out ? dd_out[f] : nullptr;
        ua.dd_in[f] = dd_in ? dd_in[f] : nullptr;
    }
    for (int d = 0; d < 3; d++) { ua.P[d] = W->flux[d].p; ua.M[d] = W->flux[3 + d].p; }
    ua.m = M->m.p; ua.dm = M->dm_src();
    ua.irho = M->irho_field ? M->irho.p : nullptr;
    ua.irho_c = M->irho_c;
    ua.dx = M->damp[0].p; ua.dy = M->damp[1].p; ua.dz = M->damp[2].p;
    ua.qext = qext;
    ua.qext_scale = qext_scale;
    ua.f = fuse;
    dim3 grd2(((M->N[0] + 3) / 4 + blk.x - 1) / blk.x, (M->N[1] + blk.y - 1) / blk.y, 1);
    // The two passes alternate over slabs of z planes so that the flux fields written by pass 1
    // are still in L2 (126 MB) when pass 2 reads them: pass 2 of slab k needs the flux planes
    // [kS - R, (k+1)S + R - 1]; pass 1 never reads what pass 2 writes (u(t-1) -> u(t+1) in place),
    // so the interleaving is free of hazards.
    const int S = ctx->opt_flux_slab > 0 ? ctx->opt_flux_slab : M->N[2];
    int fdone = -R;   // flux planes [-R, fdone) are computed
    for (int zs = 0; zs < M->N[2]; zs += S) {
        const int ze = std::min(zs + S, M->N[2]);
        fa.z0 = fdone;
        fa.z1 = ze == M->N[2] ? M->N[2] + R : ze + R;
        fdone = fa.z1;
        grd.z = (fa.z1 - fa.z0 + blk.z - 1) / blk.z;
        if (M->tti) {
            if (M->irho_field) flux3d_vec<R, true, true><<<grd, blk, 0, ctx->stream>>>(fa);
            else if (ctx->opt_flux_minb == 8) flux3d_vec<R, true, false, 8><<<grd, blk, 0, ctx->stream>>>(fa);
            else if (ctx->opt_flux_minb == 10) flux3d_vec<R, true, false, 10><<<grd, blk, 0, ctx->stream>>>(fa);
            else if (ctx->opt_flux_minb == 12) flux3d_vec<R, true, false, 12><<<grd, blk, 0, ctx->stream>>>(fa);
            else flux3d_vec<R, true, false><<<grd, blk, 0, ctx->stream>>>(fa);
        } else flux3d_vec<R, false, true><<<grd, blk, 0, ctx->stream>>>(fa);
        launch_count(ctx, "flux3d_vec");
        ua.z0 = zs;
        ua.z1 = ze;
        grd2.z = (ze - zs + blk.z - 1) / blk.z;
        if (nf == 2) {
            if (M->irho_field) fluxdiv3d_vec<R, 2, true><<<grd2, blk, 0, ctx->stream>>>(ua);
            else if (ctx->opt_flux_minb == 8) fluxdiv3d_vec<R, 2, false, 8><<<grd2, blk, 0, ctx->stream>>>(ua);
            else if (ctx->opt_flux_minb == 10) fluxdiv3d_vec<R, 2, false, 10><<<grd2, blk, 0, ctx->stream>>>(ua);
            else if (ctx->opt_flux_minb == 12) fluxdiv3d_vec<R, 2, false, 12><<<grd2, blk, 0, ctx->stream>>>(ua);
            else fluxdiv3d_vec<R, 2, false><<<grd2, blk, 0, ctx->stream>>>(ua);
        } else fluxdiv3d_vec<R, 1, true><<<grd2, blk, 0, ctx->stream>>>(ua);
        launch_count(ctx, "fluxdiv3d_vec");
    }
}

// One time step of the flux-form physics in 3-D (plain or Born).  Returns false if the radius is
// not instantiated (caller falls back to the scalar kernels).
bool flux3d_step(const judi_model *M, const StepArgs &s)
{
    Wavefield *W = s.scratch;
    JUDI_REQUIRE(W != nullptr, "flux scratch not bound");
    // isic / fwi imaging conditions and Born sources live in the generic kernels (kernels.cu)
    if ((s.f.grad && s.f.ic != JUDI_IC_AS) || (s.born && s.born_ic != JUDI_IC_AS)) return false;
    auto run = [&](auto rtag) {
        constexpr int R = decltype(rtag)::value;
        if (!s.born) {
            flux3d_one<R>(M, W, s.u, s.up, nullptr, nullptr, s.f, s.qext, s.qext_scale);
            return;
        }
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
