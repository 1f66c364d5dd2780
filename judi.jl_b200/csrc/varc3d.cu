// varc3d.cu -- single-pass 3-D variable-density acoustic time step for sm_100a (space order 8).
// Replaces the Devito loop nest of acoustic_kernel with a density field
// (src/pysource/kernels.py:46-77, FD_utils.py:16-17: `div(irho * grad(u, .5), -.5)`), and the
// two-pass form of flux3d.cu whose three flux fields cost 24 of its 52 B per grid point.
//
//   L(u) = sum_d D-_d ( c D+_d u ),   c = irho (node valued)
//
// Unlike the TTI operator there are no cross terms: the flux F_d needs only the derivative along
// its own axis, so the halo flux nodes are cheap (one 8-tap line each) and need no z history.
// One CTA owns a 64 x 16 (x,y) tile and marches along z, two phases per plane around ONE named
// barrier:
//   F  every thread (4 consecutive x of one column): F_z from the register queue of u (kept in a
//      second register queue), F_x and F_y of its own nodes from the shared u plane (TMA ring,
//      haloed 80 x 32 boxes; c from a haloed 72 x 24 box) into a double-buffered shared flux plane;
//      threads 0..31 add the 2 x-halo groups of a row, threads 0..127 the 8 y-halo rows;
//   D  D-_x F_x + D-_y F_y from the shared plane, parked for R-1 planes in registers until D-_z
//      completes, then the leapfrog update in place (+ snapshot save / imaging condition /
//      illumination).
// HBM traffic: 20 B per grid point (u, u(t-1), m, irho read; u(t+1) written) + halos that miss L2.
// Measured (C5-size grid, 40.8 M points): 0.38 ms/step against 0.50 ms for the two passes.  ncu:
// 183 M warp instructions, 40% issue-active with the 8 worker warps of the single resident CTA
// (167 registers, 80 of them the three register queues; 148 KB shared): latency-, not HBM-bound.
#include "tma_util.cuh"

namespace varc {
constexpr int R = 4, XE = 4;
constexpr int TXT = 16, TY = 16, TX = 4 * TXT;
constexpr int UPX = TX + 4 * XE, UPY = TY + 4 * R;     // haloed u plane 80 x 32 (halo 2R)
constexpr int USLOT = UPX * UPY;
constexpr int CPX = TX + 2 * XE, CPY = TY + 2 * R;     // haloed c plane 72 x 24 (halo R)
constexpr int CSLOT = CPX * CPY;
constexpr int PSLOT = TX * TY;
constexpr int NS = 8, NC = 3, NP = 3;                  // ring depths: u planes, c planes, tiles
constexpr int NQ = 2 * R;
constexpr int NT = TXT * TY;                           // 256 worker threads
constexpr int NCW = NT / 32;
constexpr int NTHREADS = NT + 32;
constexpr int FXW = (TXT + 2) * 4;                     // x-extended flux plane: 16 rows x 72
constexpr int FX = TY * FXW;
constexpr int FY = (TY + 2 * R) * TX;                  // y-extended flux plane: 24 rows x 64
constexpr int FBUF = FX + FY;
constexpr size_t SMEM = (size_t)(NS * USLOT + NC * CSLOT + NP * 3 * PSLOT + 2 * FBUF) * 4 +
                        (2 * NS + 2 * NC + 2 * NP) * 8 + 128;
}  // namespace varc

struct VarcMaps {
    CUtensorMap U, C;          // haloed planes of u(t) and of the coefficient c
    CUtensorMap P, M, I;       // tiles of u(t-1), m, irho
};

struct VarcArgs {
    Grid g;
    Coefs cf;
    float *up;
    const float *dx, *dy, *dz;
    int zchunk;
    Fuse f;
};

enum { VF_SNAP = 1, VF_GRAD = 2, VF_ILLUM = 4 };

// D+ along x of the 4 nodes of a group from the 12-float row window centred on it
__device__ __forceinline__ void varc_dplus_x(const Coefs &cf, const float *w, f2 &g0, f2 &g1)
{
    float o[4];
#pragma unroll
    for (int e = 0; e < 4; e++) {
        float acc = 0.f;
#pragma unroll
        for (int k = 1; k <= varc::R; k++)
            acc = __fmaf_rn(cf.wx[k], __fsub_rn(w[4 + e + k], w[4 + e - k + 1]), acc);
        o[e] = acc;
    }
    g0 = pk(o[0], o[1]);
    g1 = pk(o[2], o[3]);
}

__device__ __forceinline__ void varc_window(const float *p, float *w)
{
    const float4 l = *(const float4 *)(p - 4), c = *(const float4 *)p, r = *(const float4 *)(p + 4);
    w[0] = l.x; w[1] = l.y; w[2] = l.z; w[3] = l.w;
    w[4] = c.x; w[5] = c.y; w[6] = c.z; w[7] = c.w;
    w[8] = r.x; w[9] = r.y; w[10] = r.z; w[11] = r.w;
}

template <int FUSE>
__global__ void __launch_bounds__(varc::NTHREADS, 1)
    varc3d_tma(const __grid_constant__ VarcMaps tm, const VarcArgs a)
{
    using namespace varc;
    extern __shared__ unsigned char smem_raw[];
    float *ringU = (float *)(smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u));
    float *ringC = ringU + NS * USLOT;
    float *ringP = ringC + NC * CSLOT;         // [NP][3][PSLOT]: u(t-1), m, irho
    float *fbuf = ringP + NP * 3 * PSLOT;      // [2][FBUF]
    uint64_t *fullU = (uint64_t *)(fbuf + 2 * FBUF);
    uint64_t *emptyU = fullU + NS;
    uint64_t *fullC = emptyU + NS;
    uint64_t *emptyC = fullC + NC;
    uint64_t *fullP = emptyC + NC;
    uint64_t *emptyP = fullP + NP;

    const Grid &g = a.g;
    const Coefs &cf = a.cf;
    const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
    const int z0 = blockIdx.z * a.zchunk;
    const int z1 = min(z0 + a.zchunk, g.n[2]);
    const int nzc = z1 - z0;
    // iteration j brings in u plane z0-2R+1+j; flux plane zf = z0-3R+1+j (from j = 2R-1, flux
    // counter jf = j-2R+1); output plane zo = zf-R+1 = z0-4R+2+j (from j = 4R-2)
    const int total = nzc + 4 * R - 2;
    const int nflux = nzc + 2 * R - 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; s++) { mbar_init(&fullU[s], 1); mbar_init(&emptyU[s], NCW); }
        for (int s = 0; s < NC; s++) { mbar_init(&fullC[s], 1); mbar_init(&emptyC[s], NCW); }
        for (int s = 0; s < NP; s++) { mbar_init(&fullP[s], 1); mbar_init(&emptyP[s], NCW); }
        fence_barrier_init();
    }
    __syncthreads();

    if (threadIdx.x >= NT) {
        // ===== producer warp ======================================================================
        if ((threadIdx.x & 31) == 0) {
            const int bx = g.hx + x0 - 2 * XE, by = g.hy + y0 - 2 * R, bz = g.hz + z0 - 2 * R + 1;
            const int cbx = g.hx + x0 - XE, cby = g.hy + y0 - R, cbz = g.hz + z0 - R;
            const int tx0 = g.hx + x0, ty0 = g.hy + y0, tz0 = g.hz + z0;
            for (int j = 0; j < total; j++) {
                const int s = j % NS;
                if (j >= NS) mbar_wait(&emptyU[s], (uint32_t)(((j / NS) - 1) & 1));
                mbar_arrive_expect_tx(&fullU[s], USLOT * 4);
                tma_load_3d(ringU + s * USLOT, &tm.U, bx, by, bz + j, &fullU[s]);
                // coefficient plane of flux counter jc, used at iteration jc + 2R-1: one early
                const int jc = j - (2 * R - 1) + 1;
                if (jc >= 0 && jc < nflux) {
                    const int sc = jc % NC;
                    if (jc >= NC) mbar_wait(&emptyC[sc], (uint32_t)(((jc / NC) - 1) & 1));
                    mbar_arrive_expect_tx(&fullC[sc], CSLOT * 4);
                    tma_load_3d(ringC + sc * CSLOT, &tm.C, cbx, cby, cbz + jc, &fullC[sc]);
                }
                // tiles of output plane i, used at iteration i + 4R-2: one early
                const int i = j - (4 * R - 2) + 1;
                if (i >= 0 && i < nzc) {
                    const int sp = i % NP;
                    if (i >= NP) mbar_wait(&emptyP[sp], (uint32_t)(((i / NP) - 1) & 1));
                    mbar_arrive_expect_tx(&fullP[sp], 3 * PSLOT * 4);
                    tma_load_3d(ringP + sp * 3 * PSLOT, &tm.P, tx0, ty0, tz0 + i, &fullP[sp]);
                    tma_load_3d(ringP + sp * 3 * PSLOT + PSLOT, &tm.M, tx0, ty0, tz0 + i, &fullP[sp]);
                    tma_load_3d(ringP + sp * 3 * PSLOT + 2 * PSLOT, &tm.I, tx0, ty0, tz0 + i, &fullP[sp]);
                }
            }
        }
        return;
    }

    // ===== worker threads ===========================================================================
    const int tid = threadIdx.x, lane = tid & 31;
    const int tx = tid % TXT, ty = tid / TXT;
    const int x = x0 + 4 * tx, y = y0 + ty;
    const bool act = x < g.n[0] && y < g.n[1];
    const bool xfull = x + 3 < g.n[0];
    const int su = (ty + 2 * R) * UPX + 4 * tx + 2 * XE;     // own group in a u plane
    const int scn = (ty + R) * CPX + 4 * tx + XE;            // ... in a c plane
    const int sfx = ty * FXW + 4 * (tx + 1);                 // ... in the x-extended flux plane
    const int sfy = (ty + R) * TX + 4 * tx;                  // ... in the y-extended flux plane
    const int st = ty * TX + 4 * tx;                         // ... in a tile
    // extra halo work: threads 0..31 one x-halo group (row hxr, left/right), threads 0..127 one
    // y-halo group (rows -4..-1 and 16..19)
    const bool hx_on = tid < 2 * TY, hy_on = tid < TXT * 2 * R;
    const int hxr = tid >> 1, hxg = (tid & 1) ? TXT : -1;
    const int hx_u = (hxr + 2 * R) * UPX + 4 * hxg + 2 * XE, hx_c = (hxr + R) * CPX + 4 * hxg + XE;
    const int hx_f = hxr * FXW + 4 * (hxg + 1);
    const int hyh = tid / TXT, hyg = tid % TXT;
    const int hyy = hyh < R ? hyh - R : hyh - R + TY;
    const int hy_u = (hyy + 2 * R) * UPX + 4 * hyg + 2 * XE, hy_c = (hyy + R) * CPX + 4 * hyg + XE;
    const int hy_f = (hyy + R) * TX + 4 * hyg;
    long long gout = g.at(x, y, z0);

    float4 qu[NQ], qf[NQ], hq[R];
#pragma unroll
    for (int k = 0; k < NQ; k++) qu[k] = qf[k] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < R; k++) hq[k] = make_float4(0.f, 0.f, 0.f, 0.f);

    f2 dxy0 = pk1(0.f), dxy1 = pk1(0.f);
    if (act) {
        const float4 d4 = *(const float4 *)(a.dx + x);
        dxy0 = add2(lo2(d4), pk1(a.dy[y]));
        dxy1 = add2(hi2(d4), pk1(a.dy[y]));
    }
    constexpr bool f_snap = (FUSE & VF_SNAP) != 0, f_grad = (FUSE & VF_GRAD) != 0;
    constexpr bool f_il = (FUSE & VF_ILLUM) != 0, f_reg = FUSE != 0;
    [[maybe_unused]] const Fuse &fz = a.f;
    [[maybe_unused]] const bool f_xy = f_reg && act && y >= fz.reg.lo[1] && y < fz.reg.hi[1] &&
                                       x + 3 >= fz.reg.lo[0] && x < fz.reg.hi[0];
    [[maybe_unused]] const bool f_vec = f_xy && (fz.reg.lo[0] & 3) == 0 && x >= fz.reg.lo[0] &&
                                        x + 3 < fz.reg.hi[0];
    [[maybe_unused]] long long ri = f_reg ? fz.reg.at(x, y, z0) : 0;
    [[maybe_unused]] const long long rstride = f_reg ? (long long)fz.reg.pitch * fz.reg.ext[1] : 0;

    for (int j = 0; j < total; j++) {
#pragma unroll
        for (int k = 0; k < NQ - 1; k++) qu[k] = qu[k + 1];
        const int s_in = j % NS;
        mbar_wait(&fullU[s_in], (uint32_t)((j / NS) & 1));
        qu[NQ - 1] = *(const float4 *)(ringU + s_in * USLOT + su);
        if (j < 2 * R - 1) {
            if (j >= R) {     // planes that are never a flux plane of this chunk
                __syncwarp();
                if (lane == 0) mbar_arrive(&emptyU[(j + NS - R) % NS]);
            }
            continue;
        }
        // ======== F phase: flux plane zf ==========================================================
        const int jf = j - (2 * R - 1);
        const int sc = (j + NS - R) % NS, scc = jf % NC;
        const float *pu = ringU + sc * USLOT;
        float *fb = fbuf + (j & 1) * FBUF;
        mbar_wait(&fullC[scc], (uint32_t)((jf / NC) & 1));
        const float *pc = ringC + scc * CSLOT;
        {
            const float4 c4 = *(const float4 *)(pc + scn);
            const f2 c0 = lo2(c4), c1 = hi2(c4);
            // F_z = c D+z u from the register queue
            f2 gz0 = pk1(0.f), gz1 = pk1(0.f);
#pragma unroll
            for (int k = 1; k <= R; k++) {
                const float4 zp = qu[R - 1 + k], zm = qu[R - k];
                gz0 = fma2s(cf.wz[k], sub2(lo2(zp), lo2(zm)), gz0);
                gz1 = fma2s(cf.wz[k], sub2(hi2(zp), hi2(zm)), gz1);
            }
#pragma unroll
            for (int k = 0; k < NQ - 1; k++) qf[k] = qf[k + 1];
            qf[NQ - 1] = cat4(mul2(c0, gz0), mul2(c1, gz1));
            // F_x, F_y of the own nodes
            float w[12];
            varc_window(pu + su, w);
            f2 gx0, gx1;
            varc_dplus_x(cf, w, gx0, gx1);
            *(float4 *)(fb + sfx) = cat4(mul2(c0, gx0), mul2(c1, gx1));
            const float4 cen = qu[R - 1];
            f2 gy0 = pk1(0.f), gy1 = pk1(0.f);
#pragma unroll
            for (int k = 1; k <= R; k++) {
                const float4 p = *(const float4 *)(pu + su + k * UPX);
                const float4 m = (k == 1) ? cen : *(const float4 *)(pu + su - (k - 1) * UPX);
                gy0 = fma2s(cf.wy[k], sub2(lo2(p), lo2(m)), gy0);
                gy1 = fma2s(cf.wy[k], sub2(hi2(p), hi2(m)), gy1);
            }
            *(float4 *)(fb + FX + sfy) = cat4(mul2(c0, gy0), mul2(c1, gy1));
        }
        if (hx_on) {
            const float4 c4 = *(const float4 *)(pc + hx_c);
            float w[12];
            varc_window(pu + hx_u, w);
            f2 gx0, gx1;
            varc_dplus_x(cf, w, gx0, gx1);
            *(float4 *)(fb + hx_f) = cat4(mul2(lo2(c4), gx0), mul2(hi2(c4), gx1));
        }
        if (hy_on) {
            const float4 c4 = *(const float4 *)(pc + hy_c);
            f2 gy0 = pk1(0.f), gy1 = pk1(0.f);
#pragma unroll
            for (int k = 1; k <= R; k++) {
                const float4 p = *(const float4 *)(pu + hy_u + k * UPX);
                const float4 m = *(const float4 *)(pu + hy_u - (k - 1) * UPX);
                gy0 = fma2s(cf.wy[k], sub2(lo2(p), lo2(m)), gy0);
                gy1 = fma2s(cf.wy[k], sub2(hi2(p), hi2(m)), gy1);
            }
            *(float4 *)(fb + FX + hy_f) = cat4(mul2(lo2(c4), gy0), mul2(hi2(c4), gy1));
        }
        __syncwarp();
        if (lane == 0) { mbar_arrive(&emptyU[sc]); mbar_arrive(&emptyC[scc]); }
        named_bar_sync(1, NT);
        // ======== D phase =========================================================================
        {
            float w[12];
            varc_window(fb + sfx, w);
            float o[4];
#pragma unroll
            for (int e = 0; e < 4; e++) {
                float acc = 0.f;
#pragma unroll
                for (int k = 1; k <= R; k++)
                    acc = __fmaf_rn(cf.wx[k], __fsub_rn(w[4 + e + k - 1], w[4 + e - k]), acc);
                o[e] = acc;
            }
            const float *py = fb + FX + sfy;
            f2 ly0 = pk1(0.f), ly1 = pk1(0.f);
#pragma unroll
            for (int k = 1; k <= R; k++) {
                const float4 p = *(const float4 *)(py + (k - 1) * TX);
                const float4 m = *(const float4 *)(py - k * TX);
                ly0 = fma2s(cf.wy[k], sub2(lo2(p), lo2(m)), ly0);
                ly1 = fma2s(cf.wy[k], sub2(hi2(p), hi2(m)), ly1);
            }
#pragma unroll
            for (int k = 0; k < R - 1; k++) hq[k] = hq[k + 1];
            hq[R - 1] = cat4(add2(pk(o[0], o[1]), ly0), add2(pk(o[2], o[3]), ly1));
        }
        if (j >= 4 * R - 2) {
            // ---- output plane zo = zf-R+1: hq[0] is its D-x + D-y, qf[] its z-line, qu[0] its u ----
            const int i = j - (4 * R - 2);
            const int z = z0 + i;
            const int sp = i % NP;
            [[maybe_unused]] bool fin = false, fvec = false;
            [[maybe_unused]] float4 us4, g4, il4;
            if (f_reg) {
                const bool zin = z >= fz.reg.lo[2] && z < fz.reg.hi[2];
                fin = f_xy && zin;
                fvec = f_vec && zin;
                if (f_grad && f_vec && z + 3 >= fz.reg.lo[2] && z + 3 < fz.reg.hi[2]) {
                    prefetch_l2(fz.usnap[0] + ri + 3 * rstride);
                    prefetch_l2(fz.grad + ri + 3 * rstride);
                }
                if (fvec) {
                    if (f_grad) {
                        us4 = __ldg((const float4 *)(fz.usnap[0] + ri));
                        g4 = *(const float4 *)(fz.grad + ri);
                    }
                    if (f_il) il4 = *(const float4 *)(fz.illum + ri);
                }
            }
            f2 L0 = lo2(hq[0]), L1 = hi2(hq[0]);
            {
                // D-z F_z at zo: planes zo+k-1 (queue index R-1+k) and zo-k (index R-k) of the
                // flux queue, whose newest entry is plane zf = zo+R-1
                f2 lz0 = pk1(0.f), lz1 = pk1(0.f);
#pragma unroll
                for (int k = 1; k <= R; k++) {
                    const float4 p = qf[R - 1 + k], m = qf[R - k];
                    lz0 = fma2s(cf.wz[k], sub2(lo2(p), lo2(m)), lz0);
                    lz1 = fma2s(cf.wz[k], sub2(hi2(p), hi2(m)), lz1);
                }
                L0 = add2(L0, lz0);
                L1 = add2(L1, lz1);
            }
            mbar_wait(&fullP[sp], (uint32_t)((i / NP) & 1));
            const float *pt = ringP + sp * 3 * PSLOT + st;
            const float4 up4 = *(const float4 *)pt, m4 = *(const float4 *)(pt + PSLOT);
            const float4 ir4 = *(const float4 *)(pt + 2 * PSLOT);
            __syncwarp();
            if (lane == 0) mbar_arrive(&emptyP[sp]);
            const float4 uc = qu[0];
            // delta = (dt^2 L - sd (u - u-)) / (m irho + sd),  u+ = (u + (u - u-)) + delta
            const f2 dzv = pk1(a.dz[z]);
            const f2 sd0 = mul2s(cf.dt, add2(dxy0, dzv)), sd1 = mul2s(cf.dt, add2(dxy1, dzv));
            const f2 den0 = fma2(lo2(m4), lo2(ir4), sd0), den1 = fma2(hi2(m4), hi2(ir4), sd1);
            const f2 r0 = pk(__frcp_rn(den0.x), __frcp_rn(den0.y));
            const f2 r1 = pk(__frcp_rn(den1.x), __frcp_rn(den1.y));
            const f2 d10 = sub2(lo2(uc), lo2(up4)), d11 = sub2(hi2(uc), hi2(up4));
            const f2 dl0 = mul2(fma2(mul2s(-1.f, sd0), d10, mul2s(cf.dt2, L0)), r0);
            const f2 dl1 = mul2(fma2(mul2s(-1.f, sd1), d11, mul2s(cf.dt2, L1)), r1);
            const f2 un0 = add2(add2(lo2(uc), d10), dl0), un1 = add2(add2(hi2(uc), d11), dl1);
            if (act) {
                float *o = a.up + gout;
                if (xfull) *(float4 *)o = cat4(un0, un1);
                else {
                    const float v[4] = {un0.x, un0.y, un1.x, un1.y};
#pragma unroll
                    for (int e = 0; e < 4; e++)
                        if (x + e < g.n[0]) o[e] = v[e];
                }
            }
            if (f_reg && fin) {
                if (fvec) {
                    if (f_snap) __stcs((float4 *)(fz.snap[0] + ri), uc);
                    if (f_il)
                        *(float4 *)(fz.illum + ri) = cat4(fma2(mul2s(cf.dt, lo2(uc)), lo2(uc), lo2(il4)),
                                                          fma2(mul2s(cf.dt, hi2(uc)), hi2(uc), hi2(il4)));
                    if (f_grad) {
                        // gradm -= gcoef * irho * u_s * delta   (fluxdiv3d_vec)
                        const f2 w0 = mul2(mul2s(-fz.gcoef, lo2(ir4)), lo2(us4));
                        const f2 w1 = mul2(mul2s(-fz.gcoef, hi2(ir4)), hi2(us4));
                        *(float4 *)(fz.grad + ri) = cat4(fma2(w0, dl0, lo2(g4)), fma2(w1, dl1, hi2(g4)));
                    }
                } else {
                    const float ucs[4] = {uc.x, uc.y, uc.z, uc.w};
                    const float dls[4] = {dl0.x, dl0.y, dl1.x, dl1.y};
                    const float irs[4] = {ir4.x, ir4.y, ir4.z, ir4.w};
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        if (x + e < fz.reg.lo[0] || x + e >= fz.reg.hi[0]) continue;
                        if (f_snap) fz.snap[0][ri + e] = ucs[e];
                        if (f_il) fz.illum[ri + e] += cf.dt * ucs[e] * ucs[e];
                        if (f_grad) fz.grad[ri + e] -= fz.gcoef * irs[e] * fz.usnap[0][ri + e] * dls[e];
                    }
                }
            }
            gout += g.sz;
            if (f_reg) ri += rstride;
        }
    }
}

// ---- host side ---------------------------------------------------------------------------------
bool varc3d_supported(const judi_model *M)
{
    return M->ndim == 3 && !M->tti && M->irho_field && M->r == varc::R && M->ctx->encode_tiled != nullptr &&
           M->ctx->opt_varc_fused;
}

void varc3d_encode(const judi_model *M, const float *base, CUtensorMap *halo_map, CUtensorMap *tile_map)
{
    if (halo_map) tma_encode_box(M, base, varc::UPX, varc::UPY, halo_map);
    if (tile_map) tma_encode_box(M, base, varc::TX, varc::TY, tile_map);
}

void varc3d_encode_coef(const judi_model *M, const float *base, CUtensorMap *map)
{
    tma_encode_box(M, base, varc::CPX, varc::CPY, map);
}

template <int FUSE>
static void varc3d_launch(const judi_model *M, const StepArgs &s)
{
    judi_ctx *ctx = M->ctx;
    auto kern = varc3d_tma<FUSE>;
    static bool configured = false;
    if (!configured) {
        CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)varc::SMEM));
        configured = true;
    }
    Wavefield *W = s.scratch;
    VarcMaps tm;
    tm.U = W->tmap_halo[0][W->cur];
    tm.P = W->tmap_tile[0][1 - W->cur];
    tm.M = W->tmap_m;
    tm.C = W->tmap_coef;
    tm.I = W->tmap_irho;
    VarcArgs a;
    memset(&a, 0, sizeof(a));
    a.g = M->g; a.cf = M->cf;
    a.up = s.up[0];
    a.dx = M->damp[0].p; a.dy = M->damp[1].p; a.dz = M->damp[2].p;
    a.f = s.f;
    const int ntx = (M->N[0] + varc::TX - 1) / varc::TX, nty = (M->N[1] + varc::TY - 1) / varc::TY;
    int nzc = ctx->opt_zchunks;
    if (nzc <= 0) {
        // one CTA per SM: minimise waves x (planes per chunk + the 4R-2 warm-up planes of a chunk)
        long long best = -1;
        for (int c = 1; c <= 8 && c * 32 <= M->N[2]; c++) {
            const long long waves = ((long long)ntx * nty * c + ctx->num_sms - 1) / ctx->num_sms;
            const long long cost = waves * ((M->N[2] + c - 1) / c + 4 * varc::R - 2);
            if (best < 0 || cost < best) { best = cost; nzc = c; }
        }
        if (nzc <= 0) nzc = 1;
    }
    nzc = std::min(nzc, M->N[2]);
    a.zchunk = (M->N[2] + nzc - 1) / nzc;
    nzc = (M->N[2] + a.zchunk - 1) / a.zchunk;
    kern<<<dim3(ntx, nty, nzc), varc::NTHREADS, varc::SMEM, ctx->stream>>>(tm, a);
    launch_count(ctx, "varc3d_tma");
}

// One fused variable-density step; false if this combination has no single-pass instantiation.
bool varc3d_step(const judi_model *M, const StepArgs &s)
{
    if (!varc3d_supported(M) || s.born || s.qext || s.f.wrec || !s.scratch || !s.scratch->has_tmap) return false;
    if (s.f.grad && s.f.ic != JUDI_IC_AS) return false;
    const int mask = (s.f.snap[0] ? VF_SNAP : 0) | (s.f.grad ? VF_GRAD : 0) | (s.f.illum ? VF_ILLUM : 0);
    switch (mask) {
    case 0: varc3d_launch<0>(M, s); return true;
    case VF_SNAP: varc3d_launch<VF_SNAP>(M, s); return true;
    case VF_GRAD: varc3d_launch<VF_GRAD>(M, s); return true;
    case VF_ILLUM: varc3d_launch<VF_ILLUM>(M, s); return true;
    case VF_SNAP | VF_ILLUM: varc3d_launch<VF_SNAP | VF_ILLUM>(M, s); return true;
    case VF_GRAD | VF_ILLUM: varc3d_launch<VF_GRAD | VF_ILLUM>(M, s); return true;
    default: return false;
    }
}
