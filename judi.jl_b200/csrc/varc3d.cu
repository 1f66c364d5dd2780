// varc3d.cu -- single-pass 3-D variable-density acoustic time step for sm_100a (space orders 4 and 8:
// the kernel is a template on the stencil radius R; figures below are for R = 4).
// Replaces the Devito loop nest of acoustic_kernel with a density field
// (src/pysource/kernels.py:46-77, FD_utils.py:16-17: `div(irho * grad(u, .5), -.5)`), and the
// two-pass form of flux3d.cu whose three flux fields cost 24 of its 52 B per grid point.
//
//   L(u) = sum_d D-_d ( c D+_d u ),   c = irho (node valued)
//
// Unlike the TTI operator there are no cross terms: the flux F_d needs only the derivative along
// its own axis, so the halo flux nodes are cheap (one 8-tap line each) and need no z history.
// Same structure as tti3d.cu: one CTA of 12 warps owns a 32 x 16 (x,y) tile and marches along z;
// a producer warp streams three TMA rings (haloed 48 x 30 planes of u, haloed 40 x 23 planes of c,
// tiles of u(t-1), m, irho).
//  F warps (7)  tile threads (4 warps, one per group of 4 x-points) keep the z-line of u in a
//     register queue and publish F_x, F_y, F_z and u of the output plane; the x-halo warp adds F_x
//     of one group left / right of every row, the two y-halo warps F_y of rows -4..-1, 16..18;
//  D warps (4)  D-_x F_x + D-_y F_y from the shared plane into the register accumulator of that
//     plane, F_z scattered into the 8 accumulators of output planes zf-3 .. zf+4, then the leapfrog
//     update in place (+ snapshot save / imaging condition / illumination).
// The roles hand the double-buffered flux plane over through named barriers (bar.arrive / bar.sync).
// ~80 registers and 95 KB of shared memory: two CTAs per SM.
// HBM traffic: 20 B per grid point (u, u(t-1), m, irho read; u(t+1) written) + halos that miss L2.
// Measured (C5-size grid, 40.8 M points): plain 0.315 ms, save 0.345, imaging condition 0.48 ms
// per step against 0.50 / 0.52 / 0.55 ms for the two passes.  ncu (profiles/r01_ncu_varc3d.md): 206 M
// warp instructions, 63% issue-active with 24 warps per SM, DRAM 0.95 GB = 3.0 TB/s: bound by
// instruction issue (only 29% of the instructions are arithmetic: register-queue moves 15%,
// barrier / mbarrier / address bookkeeping the rest), not by HBM (20 B/pt = 0.125 ms).
// History: the first version (64 x 16 tile, F and D in the same 8 warps behind one barrier, flux
// and partial-sum queues in 167 registers, one CTA per SM) ran 0.38 ms/step; ncu: 40%
// issue-active, latency-bound.
#include "tma_util.cuh"

// Geometry of one CTA for stencil radius R = space_order / 2 (2 or 4).  The x layout works in groups
// of 4 points and keeps one halo group per side whatever R <= 4 is; the y / z extents follow R.
template <int R_>
struct VarcC {
    static constexpr int R = R_;
    static constexpr int TXG = 8, TX = 4 * TXG, TY = 16;      // tile: 8 groups of 4 x-points, 16 rows
    static constexpr int UPX = TX + 16, UPY = TY + 4 * R - 2; // haloed u plane, origin (x0-8, y0-(2R-1))
    static constexpr int USLOT = ((UPX * UPY + 31) / 32) * 32;
    static constexpr int CPX = TX + 8, CPY = TY + 2 * R - 1;  // coefficient plane, origin (x0-4, y0-R)
    static constexpr int CSLOT = ((CPX * CPY + 31) / 32) * 32;
    static constexpr int PSLOT = TX * TY;
    static constexpr int NS = 8, NC = 3, NP = 3;              // ring depths: u planes, c planes, tiles
    static constexpr int NQ = 2 * R;
    static constexpr int NI = TXG * TY;                       // 128 tile groups
    static constexpr int NYH = (((2 * R - 1) * TXG + 31) / 32) * 32;  // y-halo threads (whole warps)
    static constexpr int NF = NI + 32 + NYH;                  // F warps: tile, x halo, y halo
    static constexpr int ND = NI;                             // D warps: one thread per tile group
    static constexpr int NTHREADS = NF + ND + 32;             // + the producer warp
    static constexpr int FXW = TX + 8;                        // row pitch of the x-extended flux plane (40)
    static constexpr int FX = TY * FXW;
    static constexpr int FY = CPY * TX;                       // y-extended flux plane (rows -R .. TY+R-2)
    static constexpr int FBUF = FX + FY + 2 * PSLOT;          // F_x, F_y of one plane + F_z, u tiles
    static constexpr int LEADU = 3, LEADC = 2;                // producer run-ahead (planes)
    static constexpr int WARM = 4 * R - 2;                    // iterations before the first output plane
    static constexpr int FIRST = 2 * R - 1;                   // first flux iteration
    static constexpr size_t SMEM = (size_t)(NS * USLOT + NC * CSLOT + NP * 3 * PSLOT + 2 * FBUF) * 4 +
                                   (2 * NS + 2 * NC + 2 * NP) * 8 + 128;
    static_assert(R == 2 || R == 4, "x halo of one 4-point group: R <= 4");
    static_assert(R + 1 + LEADU <= NS, "u ring too short");
};

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
    float *dd_out;           // Born background pass: write delta = u+ - 2u + u-
    const float *dd_in;      // Born linearised pass: delta of the background field ...
    const float *dm;         // ... and the model perturbation: q = -dm irho delta
    float *lap_out;          // VF_LAPOUT: dt^2 div(c grad u)  (Grid layout)
};

// VF_LAPOUT: no time update at all -- the launch writes dt^2 div(c grad u) of the haloed field to
// `lap_out` (the isic / fwi Born source `laplacian(u, dm irho)`, sensitivity.py:191-209, with the
// perturbation dm as the coefficient plane); the tile ring (u(t-1), m, irho) is not streamed.
enum { VF_SNAP = 1, VF_GRAD = 2, VF_ILLUM = 4, VF_DDOUT = 8, VF_QSRC = 16, VF_LAPOUT = 32 };

// staggered first derivative along x at the 4 nodes of a group from the 12-float row window centred
// on it: PLUS: D+ (node + 1/2), sum_k w_k (u[e+k] - u[e-k+1]); else D- (node - 1/2)
template <bool PLUS, int R>
__device__ __forceinline__ void varc_dx(const float *wx, const float4 &l, const float4 &c,
                                        const float4 &r, f2 &o0, f2 &o1)
{
    const float w[12] = {l.x, l.y, l.z, l.w, c.x, c.y, c.z, c.w, r.x, r.y, r.z, r.w};
    float o[4];
#pragma unroll
    for (int e = 0; e < 4; e++) {
        float acc = 0.f;
#pragma unroll
        for (int k = 1; k <= R; k++) {
            const float d = PLUS ? __fsub_rn(w[4 + e + k], w[4 + e - k + 1])
                                 : __fsub_rn(w[4 + e + k - 1], w[4 + e - k]);
            acc = (k == 1) ? __fmul_rn(wx[1], d) : __fmaf_rn(wx[k], d, acc);
        }
        o[e] = acc;
    }
    o0 = pk(o[0], o[1]);
    o1 = pk(o[2], o[3]);
}

template <int R, int FUSE, bool ROTD>
__global__ void __launch_bounds__(VarcC<R>::NTHREADS, 2)
    varc3d_tma(const __grid_constant__ VarcMaps tm, const VarcArgs a)
{
    typedef VarcC<R> C;
    constexpr int TXG = C::TXG, TX = C::TX, TY = C::TY, UPX = C::UPX, USLOT = C::USLOT, CPX = C::CPX;
    constexpr int CPY = C::CPY, CSLOT = C::CSLOT, PSLOT = C::PSLOT, NS = C::NS, NC = C::NC, NP = C::NP;
    constexpr int NQ = C::NQ, NI = C::NI, NF = C::NF, ND = C::ND, FXW = C::FXW, FX = C::FX, FY = C::FY;
    constexpr int FBUF = C::FBUF, LEADU = C::LEADU, LEADC = C::LEADC, WARM = C::WARM, FIRST = C::FIRST;
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
    // iteration j brings in u plane z0-(2R-1)+j; flux plane zf = z0-(3R-1)+j (from j = 2R-1); output
    // plane zo = zf-(R-1) = z0-(4R-2)+j (from j = 4R-2)
    const int total = nzc + WARM;
    constexpr int NSYNC = NF + ND;

    if (threadIdx.x == 0) {
        // only the tile warps read the incoming plane and the u plane of the z-line; every F warp
        // reads the flux plane's lines and the coefficient plane
        for (int s = 0; s < NS; s++) { mbar_init(&fullU[s], 1); mbar_init(&emptyU[s], NF / 32); }
        for (int s = 0; s < NC; s++) { mbar_init(&fullC[s], 1); mbar_init(&emptyC[s], NF / 32); }
        for (int s = 0; s < NP; s++) { mbar_init(&fullP[s], 1); mbar_init(&emptyP[s], ND / 32); }
        fence_barrier_init();
    }
    __syncthreads();
    const int tid = threadIdx.x, lane = tid & 31;

    if (tid >= NF + ND) {
        // ===== producer warp ======================================================================
        if (lane == 0) {
            const int bx = g.hx + x0 - 8, by = g.hy + y0 - (2 * R - 1), bz = g.hz + z0 - (2 * R - 1);
            const int cx0 = g.hx + x0 - 4, cy0 = g.hy + y0 - R, cz0 = g.hz + z0 - R;
            const int tx0 = g.hx + x0, ty0 = g.hy + y0, tz0 = g.hz + z0;
            for (int j = -LEADU; j < total; j++) {
                const int pj = j + LEADU;                 // u plane (consumer iteration pj)
                if (pj < total) {
                    const int s = pj % NS;
                    if (pj >= NS) mbar_wait_sleep(&emptyU[s], (uint32_t)(((pj / NS) - 1) & 1));
                    mbar_arrive_expect_tx(&fullU[s], UPX * C::UPY * 4);
                    tma_load_3d(ringU + s * USLOT, &tm.U, bx, by, bz + pj, &fullU[s]);
                }
                const int c = j + LEADC - FIRST;          // flux plane counter (iteration c + 7)
                if (c >= 0 && c < total - FIRST) {
                    const int s = c % NC;
                    if (c >= NC) mbar_wait_sleep(&emptyC[s], (uint32_t)(((c / NC) - 1) & 1));
                    mbar_arrive_expect_tx(&fullC[s], CPX * CPY * 4);
                    tma_load_3d(ringC + s * CSLOT, &tm.C, cx0, cy0, cz0 + c, &fullC[s]);
                }
                const int i = j + LEADC - WARM;           // output plane counter (iteration i + 14)
                if (!(FUSE & VF_LAPOUT) && i >= 0 && i < nzc) {
                    const int s = i % NP;
                    if (i >= NP) mbar_wait_sleep(&emptyP[s], (uint32_t)(((i / NP) - 1) & 1));
                    mbar_arrive_expect_tx(&fullP[s], 3 * PSLOT * 4);
                    tma_load_3d(ringP + s * 3 * PSLOT, &tm.P, tx0, ty0, tz0 + i, &fullP[s]);
                    tma_load_3d(ringP + s * 3 * PSLOT + PSLOT, &tm.M, tx0, ty0, tz0 + i, &fullP[s]);
                    tma_load_3d(ringP + s * 3 * PSLOT + 2 * PSLOT, &tm.I, tx0, ty0, tz0 + i, &fullP[s]);
                }
            }
        }
        return;
    }

    if (tid < NI) {
        // ===== F role, tile warps: F_x, F_y, F_z of the own nodes ====================================
        const int cx = tid & (TXG - 1), cy = tid / TXG;
        const int su = (cy + 2 * R - 1) * UPX + 4 * cx + 8;       // node group inside a u plane
        const int sc = (cy + R) * CPX + 4 * cx + 4;               // ... inside a coefficient plane
        const int sfx = cy * FXW + 4 * (cx + 1);                  // ... inside the x-extended flux plane
        const int sfy = (cy + R) * TX + 4 * cx;                   // ... inside the y-extended flux plane
        const int st = cy * TX + 4 * cx;                          // ... inside a tile
        float4 q[NQ];   // u of planes zf-3 .. zf+4 at this node group
#pragma unroll
        for (int k = 0; k < NQ; k++) q[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int j = 0; j < FIRST; j++) {
#pragma unroll
            for (int k = 0; k < NQ - 1; k++) q[k] = q[k + 1];
            mbar_wait(&fullU[j], 0);
            q[NQ - 1] = *(const float4 *)(ringU + j * USLOT + su);
            if (j < R - 1) {   // the first three planes of a chunk are never a flux plane
                __syncwarp();
                mbar_arrive_if(&emptyU[j], lane == 0);
            }
        }
        int cslot = 0, cpar = 0;
        for (int j = FIRST; j < total; j++) {
#pragma unroll
            for (int k = 0; k < NQ - 1; k++) q[k] = q[k + 1];
            const int s_in = j & (NS - 1), s_c = (j + NS - R) & (NS - 1);
            mbar_wait(&fullU[s_in], (uint32_t)((j / NS) & 1));
            q[NQ - 1] = *(const float4 *)(ringU + s_in * USLOT + su);
            const int zf = z0 - (3 * R - 1) + j;
            const bool need_xy = zf >= z0 && zf < z1;       // its x / y flux is used
            float *fb = fbuf + (j & 1) * FBUF;
            mbar_wait(&fullC[cslot], (uint32_t)cpar);
            const float *pl = ringU + s_c * USLOT + su;
            const float4 c4 = *(const float4 *)(ringC + cslot * CSLOT + sc);
            const float4 cen = q[R - 1];
            f2 gx0, gx1, gy0, gy1, gz0, gz1;
            varc_dx<true, R>(cf.wx, *(const float4 *)(pl - 4), cen, *(const float4 *)(pl + 4), gx0, gx1);
#pragma unroll
            for (int k = 1; k <= R; k++) {
                const float4 p = *(const float4 *)(pl + k * UPX);
                const float4 m = (k == 1) ? cen : *(const float4 *)(pl - (k - 1) * UPX);
                const f2 dy0 = sub2(lo2(p), lo2(m)), dy1 = sub2(hi2(p), hi2(m));
                const float4 zp = q[R - 1 + k], zm = q[R - k];
                const f2 dz0 = sub2(lo2(zp), lo2(zm)), dz1 = sub2(hi2(zp), hi2(zm));
                if (k == 1) {
                    gy0 = mul2s(cf.wy[1], dy0); gy1 = mul2s(cf.wy[1], dy1);
                    gz0 = mul2s(cf.wz[1], dz0); gz1 = mul2s(cf.wz[1], dz1);
                } else {
                    gy0 = fma2s(cf.wy[k], dy0, gy0); gy1 = fma2s(cf.wy[k], dy1, gy1);
                    gz0 = fma2s(cf.wz[k], dz0, gz0); gz1 = fma2s(cf.wz[k], dz1, gz1);
                }
            }
            __syncwarp();
            mbar_arrive_if(&emptyU[s_c], lane == 0); mbar_arrive_if(&emptyC[cslot], lane == 0);
            if (++cslot == NC) { cslot = 0; cpar ^= 1; }
            // the flux buffer of this parity was read by the D warps two planes ago
            if (j >= FIRST + 2) named_bar_sync(3 + (j & 1), NSYNC);
            if (need_xy) {
                *(float4 *)(fb + sfx) = cat4(mul2(lo2(c4), gx0), mul2(hi2(c4), gx1));
                *(float4 *)(fb + FX + sfy) = cat4(mul2(lo2(c4), gy0), mul2(hi2(c4), gy1));
            }
            float *fz_ = fb + FX + FY + st;
            *(float4 *)fz_ = cat4(mul2(lo2(c4), gz0), mul2(hi2(c4), gz1));
            *(float4 *)(fz_ + PSLOT) = q[0];      // u of the output plane zf-3
            __threadfence_block();
            named_bar_arrive(1 + (j & 1), NSYNC);
        }
        return;
    }
    if (tid < NF) {
        // ===== F role, halo warps: F_x one group left / right of every row (warp 4), F_y of rows
        // -4..-1 and 16..18 (warps 5, 6).  No z history: they only ever read the flux plane.
        const bool xh = tid < NI + 32;
        int cx, cy;
        bool wr = true;
        if (xh) {
            const int h = tid - NI;
            cx = (h & 1) ? TXG : -1;
            cy = h >> 1;
        } else {
            const int h = tid - NI - 32, row = h / TXG;   // rows beyond 2R-2 are idle (mirror the last one)
            cx = h & (TXG - 1);
            cy = row < R ? row - R : TY + min(row, 2 * R - 2) - R;
            wr = row < 2 * R - 1;
        }
        const int su = (cy + 2 * R - 1) * UPX + 4 * cx + 8;
        const int sc = (cy + R) * CPX + 4 * cx + 4;
        const int sf = xh ? cy * FXW + 4 * (cx + 1) : FX + (cy + R) * TX + 4 * cx;
        if (lane == 0)
            for (int j = 0; j < R - 1; j++) mbar_arrive(&emptyU[j]);   // planes they never read
        int cslot = 0, cpar = 0;
        for (int j = FIRST; j < total; j++) {
            const int s_c = (j + NS - R) & (NS - 1);
            const int zf = z0 - (3 * R - 1) + j;
            const bool need_xy = zf >= z0 && zf < z1;
            float *fb = fbuf + (j & 1) * FBUF;
            mbar_wait(&fullU[s_c], (uint32_t)(((j - R) / NS) & 1));
            mbar_wait(&fullC[cslot], (uint32_t)cpar);
            f2 g0 = pk1(0.f), g1 = pk1(0.f);
            float4 c4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (need_xy) {
                const float *pl = ringU + s_c * USLOT + su;
                c4 = *(const float4 *)(ringC + cslot * CSLOT + sc);
                const float4 cen = *(const float4 *)pl;
                if (xh) varc_dx<true, R>(cf.wx, *(const float4 *)(pl - 4), cen, *(const float4 *)(pl + 4), g0, g1);
                else {
#pragma unroll
                    for (int k = 1; k <= R; k++) {
                        const float4 p = *(const float4 *)(pl + k * UPX);
                        const float4 m = (k == 1) ? cen : *(const float4 *)(pl - (k - 1) * UPX);
                        const f2 d0 = sub2(lo2(p), lo2(m)), d1 = sub2(hi2(p), hi2(m));
                        if (k == 1) { g0 = mul2s(cf.wy[1], d0); g1 = mul2s(cf.wy[1], d1); }
                        else { g0 = fma2s(cf.wy[k], d0, g0); g1 = fma2s(cf.wy[k], d1, g1); }
                    }
                }
            }
            __syncwarp();
            mbar_arrive_if(&emptyU[s_c], lane == 0); mbar_arrive_if(&emptyC[cslot], lane == 0);
            if (++cslot == NC) { cslot = 0; cpar ^= 1; }
            if (j >= FIRST + 2) named_bar_sync(3 + (j & 1), NSYNC);
            if (need_xy && wr) *(float4 *)(fb + sf) = cat4(mul2(lo2(c4), g0), mul2(hi2(c4), g1));
            __threadfence_block();
            named_bar_arrive(1 + (j & 1), NSYNC);
        }
        return;
    }

    // ===== D role: divergence, time update and fused extras of one tile group ======================
    const int dt_ = tid - NF;
    const int cx = dt_ & (TXG - 1), cy = dt_ / TXG;
    const int sfx = cy * FXW + 4 * (cx + 1), sfy = (cy + R) * TX + 4 * cx, st = cy * TX + 4 * cx;
    const int ox = x0 + 4 * cx, oy = y0 + cy;
    const bool act = ox < g.n[0] && oy < g.n[1];
    const bool xfull = ox + 3 < g.n[0];
    long long gout = g.at(ox, oy, z0);
    f2 dxy0 = pk1(0.f), dxy1 = pk1(0.f);
    if (act) {
        const float4 d4 = *(const float4 *)(a.dx + ox);
        dxy0 = add2(lo2(d4), pk1(a.dy[oy]));
        dxy1 = add2(hi2(d4), pk1(a.dy[oy]));
    }
    constexpr bool f_snap = (FUSE & VF_SNAP) != 0, f_grad = (FUSE & VF_GRAD) != 0;
    constexpr bool f_il = (FUSE & VF_ILLUM) != 0, f_reg = (FUSE & (VF_SNAP | VF_GRAD | VF_ILLUM)) != 0;
    constexpr bool f_dd = (FUSE & VF_DDOUT) != 0, f_q = (FUSE & VF_QSRC) != 0;
    [[maybe_unused]] const Fuse &fz = a.f;
    [[maybe_unused]] const bool f_xy = f_reg && act && oy >= fz.reg.lo[1] && oy < fz.reg.hi[1] &&
                                       ox + 3 >= fz.reg.lo[0] && ox < fz.reg.hi[0];
    [[maybe_unused]] const bool f_vec = f_xy && (fz.reg.lo[0] & 3) == 0 && ox >= fz.reg.lo[0] &&
                                        ox + 3 < fz.reg.hi[0];
    [[maybe_unused]] long long ri = f_reg ? fz.reg.at(ox, oy, z0) : 0;
    [[maybe_unused]] const long long rstride = f_reg ? (long long)fz.reg.pitch * fz.reg.ext[1] : 0;

    float4 acc[NQ];   // divergence accumulators of the output planes zf-3 .. zf+4
#pragma unroll
    for (int k = 0; k < NQ; k++) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    int pslot = 0, ppar = 0;

    // accumulator slot of output plane zf + d, d = -3 .. 4: rotated at compile time (loop unrolled by
    // 8) or shifted by 7 register moves per plane
#define AS(d) (ROTD ? ((ph + (d) + 2 * NQ) % NQ) : ((d) + R - 1))
    constexpr int UNR = ROTD ? NQ : 1;
    for (int j0 = FIRST; j0 < total; j0 += UNR) {
#pragma unroll
      for (int ph = 0; ph < UNR; ph++) {
        const int j = j0 + ph;
        if (ROTD && j >= total) break;
        const int zf = z0 - (3 * R - 1) + j;
        const bool need_xy = zf >= z0 && zf < z1;
        const bool out = j >= WARM;
        const int z = z0 + j - WARM;                  // output plane zo = zf - 3
        const float *fb = fbuf + (j & 1) * FBUF;
        // ---- everything that comes from global memory is requested before the hand-over wait -----
        float dz_s = 0.f;
        [[maybe_unused]] bool fin = false, fvec = false;
        [[maybe_unused]] float4 us4, G4, I4;
        [[maybe_unused]] float4 dq4 = make_float4(0.f, 0.f, 0.f, 0.f), dm4 = dq4;
        if (out) {
            dz_s = __ldg(a.dz + z);
            if (f_q && act) {
                dq4 = *(const float4 *)(a.dd_in + gout);
                dm4 = __ldg((const float4 *)(a.dm + gout));
            }
            if (f_reg) {
                const bool zin = z >= fz.reg.lo[2] && z < fz.reg.hi[2];
                fin = f_xy && zin;
                fvec = f_vec && zin;
                if ((f_grad || f_il) && f_vec && z + 3 >= fz.reg.lo[2] && z + 3 < fz.reg.hi[2]) {
                    if (f_grad) {
                        prefetch_l2(fz.usnap[0] + ri + 3 * rstride);
                        prefetch_l2(fz.grad + ri + 3 * rstride);
                    }
                    if (f_il) prefetch_l2(fz.illum + ri + 3 * rstride);
                }
                if (fvec && f_grad) {
                    us4 = __ldg((const float4 *)(fz.usnap[0] + ri));
                    G4 = *(const float4 *)(fz.grad + ri);
                }
            }
        }
        named_bar_sync(1 + (j & 1), NSYNC);
        // ---- D-_x F_x + D-_y F_y of flux plane zf go to the accumulator of output plane zf --------
        if (need_xy) {
            const float *px = fb + sfx;
            f2 lx0, lx1;
            varc_dx<false, R>(cf.wx, *(const float4 *)(px - 4), *(const float4 *)px, *(const float4 *)(px + 4),
                           lx0, lx1);
            const float *py = fb + FX + sfy;
            f2 ly0, ly1;
#pragma unroll
            for (int k = 1; k <= R; k++) {
                const float4 p = *(const float4 *)(py + (k - 1) * TX);
                const float4 m = *(const float4 *)(py - k * TX);
                const f2 d0 = sub2(lo2(p), lo2(m)), d1 = sub2(hi2(p), hi2(m));
                if (k == 1) { ly0 = mul2s(cf.wy[1], d0); ly1 = mul2s(cf.wy[1], d1); }
                else { ly0 = fma2s(cf.wy[k], d0, ly0); ly1 = fma2s(cf.wy[k], d1, ly1); }
            }
            acc[AS(0)] = cat4(add2(lo2(acc[AS(0)]), add2(lx0, ly0)), add2(hi2(acc[AS(0)]), add2(lx1, ly1)));
        }
        // ---- z part: D-_z out[z] = sum_k wz[k] (F[z+k-1] - F[z-k]), scattered from F_z of plane zf ---
        const float *fz_ = fb + FX + FY + st;
        {
            const float4 Fz = *(const float4 *)fz_;
#pragma unroll
            for (int k = 1; k <= R; k++) {
                float4 &ap = acc[AS(1 - k)];  // plane zf-k+1
                ap = cat4(fma2s(cf.wz[k], lo2(Fz), lo2(ap)), fma2s(cf.wz[k], hi2(Fz), hi2(ap)));
                float4 &bp = acc[AS(k)];      // plane zf+k
                if (k == R) bp = cat4(mul2s(-cf.wz[k], lo2(Fz)), mul2s(-cf.wz[k], hi2(Fz)));
                else bp = cat4(fma2s(-cf.wz[k], lo2(Fz), lo2(bp)), fma2s(-cf.wz[k], hi2(Fz), hi2(bp)));
            }
        }
        if (out && (FUSE & VF_LAPOUT)) {
            const float4 H = acc[AS(1 - R)];
            const f2 h0 = mul2s(cf.dt2, lo2(H)), h1 = mul2s(cf.dt2, hi2(H));
            if (act) {
                float *o = a.lap_out + gout;
                if (xfull) *(float4 *)o = cat4(h0, h1);
                else {
                    const float v[4] = {h0.x, h0.y, h1.x, h1.y};
#pragma unroll
                    for (int e = 0; e < 4; e++)
                        if (ox + e < g.n[0]) o[e] = v[e];
                }
            }
            gout += g.sz;
        } else if (out) {
            const float4 uc = *(const float4 *)(fz_ + PSLOT);
            const float4 H = acc[AS(1 - R)];
            mbar_wait(&fullP[pslot], (uint32_t)ppar);
            const float *pt = ringP + pslot * 3 * PSLOT + st;
            const float4 up4 = *(const float4 *)pt, m4 = *(const float4 *)(pt + PSLOT);
            const float4 ir4 = *(const float4 *)(pt + 2 * PSLOT);
            __syncwarp();
            mbar_arrive_if(&emptyP[pslot], lane == 0);
            if (++pslot == NP) { pslot = 0; ppar ^= 1; }
            // delta = (dt^2 L - sd (u - u-)) / (m irho + sd),  u+ = (u + (u - u-)) + delta
            const f2 dzv = pk1(dz_s);
            const f2 sd0 = mul2s(cf.dt, add2(dxy0, dzv)), sd1 = mul2s(cf.dt, add2(dxy1, dzv));
            const f2 den0 = fma2(lo2(m4), lo2(ir4), sd0), den1 = fma2(hi2(m4), hi2(ir4), sd1);
            const f2 r0 = pk(__frcp_rn(den0.x), __frcp_rn(den0.y));
            const f2 r1 = pk(__frcp_rn(den1.x), __frcp_rn(den1.y));
            const f2 d10 = sub2(lo2(uc), lo2(up4)), d11 = sub2(hi2(uc), hi2(up4));
            // one explicit fma for the update: ptxas contracts a packed mul + add pair in the flavours
            // that do not also use delta, and the field must not depend on the flavour
            f2 h0 = mul2s(cf.dt2, lo2(H)), h1 = mul2s(cf.dt2, hi2(H));
            if (f_q) {
                // dt^2 * lin_src = -dm irho (u+ - 2u + u-) of the background field
                // (sensitivity.py:174-188; the post-injection part is added by sparse.cu)
                h0 = fma2(mul2(mul2s(-1.f, lo2(dm4)), lo2(ir4)), lo2(dq4), h0);
                h1 = fma2(mul2(mul2s(-1.f, hi2(dm4)), hi2(ir4)), hi2(dq4), h1);
            }
            const f2 t0 = fma2(mul2s(-1.f, sd0), d10, h0);
            const f2 t1 = fma2(mul2s(-1.f, sd1), d11, h1);
            [[maybe_unused]] const f2 dl0 = mul2(t0, r0), dl1 = mul2(t1, r1);
            const f2 un0 = fma2(t0, r0, add2(lo2(uc), d10)), un1 = fma2(t1, r1, add2(hi2(uc), d11));
            if (act) {
                float *o = a.up + gout;
                if (xfull) {
                    *(float4 *)o = cat4(un0, un1);
                    if (f_dd) *(float4 *)(a.dd_out + gout) = cat4(dl0, dl1);
                } else {
                    const float v[4] = {un0.x, un0.y, un1.x, un1.y};
                    const float dv[4] = {dl0.x, dl0.y, dl1.x, dl1.y};
#pragma unroll
                    for (int e = 0; e < 4; e++)
                        if (ox + e < g.n[0]) {
                            o[e] = v[e];
                            if (f_dd) a.dd_out[gout + e] = dv[e];
                        }
                }
            }
            if (f_reg && fin) {
                if (fvec) {
                    if (f_snap) __stcs((float4 *)(fz.snap[0] + ri), uc);
                    if (f_il) {
                        I4 = *(const float4 *)(fz.illum + ri);
                        *(float4 *)(fz.illum + ri) = cat4(fma2(mul2s(cf.dt, lo2(uc)), lo2(uc), lo2(I4)),
                                                          fma2(mul2s(cf.dt, hi2(uc)), hi2(uc), hi2(I4)));
                    }
                    if (f_grad) {
                        // gradm -= gcoef * irho * u_s * delta   (sensitivity.py:55-69)
                        const f2 w0 = mul2(mul2s(-fz.gcoef, lo2(ir4)), lo2(us4));
                        const f2 w1 = mul2(mul2s(-fz.gcoef, hi2(ir4)), hi2(us4));
                        *(float4 *)(fz.grad + ri) = cat4(fma2(w0, dl0, lo2(G4)), fma2(w1, dl1, hi2(G4)));
                    }
                } else {
                    const float ucs[4] = {uc.x, uc.y, uc.z, uc.w};
                    const float dls[4] = {dl0.x, dl0.y, dl1.x, dl1.y};
                    const float irs[4] = {ir4.x, ir4.y, ir4.z, ir4.w};
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        if (ox + e < fz.reg.lo[0] || ox + e >= fz.reg.hi[0]) continue;
                        if (f_snap) fz.snap[0][ri + e] = ucs[e];
                        if (f_il) fz.illum[ri + e] = __fmaf_rn(__fmul_rn(cf.dt, ucs[e]), ucs[e], fz.illum[ri + e]);
                        if (f_grad)
                            fz.grad[ri + e] = __fmaf_rn(__fmul_rn(__fmul_rn(-fz.gcoef, irs[e]), fz.usnap[0][ri + e]),
                                                        dls[e], fz.grad[ri + e]);
                    }
                }
            }
            gout += g.sz;
            if (f_reg) ri += rstride;
        }
        // this flux buffer may be overwritten by the F warps two planes from now
        if (j + 2 < total) named_bar_arrive(3 + (j & 1), NSYNC);
        if (!ROTD) {
#pragma unroll
            for (int k = 0; k < NQ - 1; k++) acc[k] = acc[k + 1];
        }
      }
    }
#undef AS
}

// ---- host side ---------------------------------------------------------------------------------
bool varc3d_supported(const judi_model *M)
{
    return M->ndim == 3 && !M->tti && !M->visco && M->irho_field && (M->r == 4 || M->r == 2) && M->ctx->encode_tiled != nullptr &&
           M->ctx->opt_varc_fused;
}

void varc3d_encode(const judi_model *M, const float *base, CUtensorMap *halo_map, CUtensorMap *tile_map)
{
    const int upy = M->r == 2 ? VarcC<2>::UPY : VarcC<4>::UPY;
    if (halo_map) tma_encode_box(M, base, VarcC<4>::UPX, upy, halo_map);
    if (tile_map) tma_encode_box(M, base, VarcC<4>::TX, VarcC<4>::TY, tile_map);
}

void varc3d_encode_coef(const judi_model *M, const float *base, CUtensorMap *map)
{
    tma_encode_box(M, base, VarcC<4>::CPX, M->r == 2 ? VarcC<2>::CPY : VarcC<4>::CPY, map);
}

// lin: launch on the linearised wavefield (Born source in); otherwise on the background field
template <int R, int FUSE>
static void varc3d_launch_r(const judi_model *M, const StepArgs &s, bool lin)
{
    typedef VarcC<R> C;
    judi_ctx *ctx = M->ctx;
    // option varc_fused: 1 shifted accumulators, 2 accumulators rotated at compile time
    const int var = ctx->opt_varc_fused == 2 ? 1 : 0;
    void (*kern)(const VarcMaps, const VarcArgs) = var ? varc3d_tma<R, FUSE, true> : varc3d_tma<R, FUSE, false>;
    static bool configured[2] = {false, false};
    if (!configured[var]) {
        CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
        configured[var] = true;
    }
    Wavefield *W = lin ? s.lin : s.scratch;
    VarcMaps tm;
    tm.U = W->tmap_halo[0][W->cur];
    tm.P = W->tmap_tile[0][1 - W->cur];
    tm.M = W->tmap_m;
    tm.C = W->tmap_coef;
    tm.I = W->tmap_irho;
    VarcArgs a;
    memset(&a, 0, sizeof(a));
    a.g = M->g; a.cf = M->cf;
    a.up = lin ? s.ulp[0] : s.up[0];
    a.dx = M->damp[0].p; a.dy = M->damp[1].p; a.dz = M->damp[2].p;
    if (!lin) a.f = s.f;
    if (FUSE & VF_DDOUT) a.dd_out = s.scratch->dd[0].p;
    if (FUSE & VF_QSRC) { a.dd_in = s.scratch->dd[0].p; a.dm = M->dm_src(); }
    const int ntx = (M->N[0] + C::TX - 1) / C::TX, nty = (M->N[1] + C::TY - 1) / C::TY;
    int nzc = ctx->opt_zchunks;
    if (nzc <= 0) {
        // two CTAs per SM: minimise waves x (planes per chunk + the 4R-2 warm-up planes of a chunk)
        long long best = -1;
        for (int c = 1; c <= 8 && c * 32 <= M->N[2]; c++) {
            const long long slots = 2LL * ctx->num_sms;
            const long long waves = ((long long)ntx * nty * c + slots - 1) / slots;
            // (a single chunk per tile column leaves the launch waiting for its slowest CTA:
            // measured 0.323 vs 0.315 ms on the C5-size grid, hence the 6% handicap)
            const long long cost = waves * ((M->N[2] + c - 1) / c + C::WARM) * (c == 1 ? 106 : 100);
            if (best < 0 || cost < best) { best = cost; nzc = c; }
        }
        if (nzc <= 0) nzc = 1;
    }
    nzc = std::min(nzc, M->N[2]);
    a.zchunk = (M->N[2] + nzc - 1) / nzc;
    nzc = (M->N[2] + a.zchunk - 1) / a.zchunk;
    kern<<<dim3(ntx, nty, nzc), C::NTHREADS, C::SMEM, ctx->stream>>>(tm, a);
    launch_count(ctx, "varc3d_tma");
}

template <int FUSE>
static void varc3d_launch(const judi_model *M, const StepArgs &s, bool lin = false)
{
    if (M->r == 2) varc3d_launch_r<2, FUSE>(M, s, lin);
    else varc3d_launch_r<4, FUSE>(M, s, lin);
}

// dt^2 div(c grad u) of a constant-density model's field u (tensor map `umap`: haloed varc box over
// the field, `cmap`: coefficient box over c) into `out`: the isic / fwi Born source term.
template <int R>
static void varc3d_lapw_r(const judi_model *M, const CUtensorMap &umap, const CUtensorMap &cmap, float *out)
{
    typedef VarcC<R> C;
    judi_ctx *ctx = M->ctx;
    void (*kern)(const VarcMaps, const VarcArgs) = varc3d_tma<R, VF_LAPOUT, false>;
    static bool configured = false;
    if (!configured) {
        CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
        configured = true;
    }
    VarcMaps tm;
    tm.U = umap; tm.C = cmap;
    tm.P = umap; tm.M = umap; tm.I = umap;   // not streamed by this flavour
    VarcArgs a;
    memset(&a, 0, sizeof(a));
    a.g = M->g; a.cf = M->cf;
    a.dx = M->damp[0].p; a.dy = M->damp[1].p; a.dz = M->damp[2].p;
    a.lap_out = out;
    const int ntx = (M->N[0] + C::TX - 1) / C::TX, nty = (M->N[1] + C::TY - 1) / C::TY;
    int nzc = 1;
    {
        long long best = -1;
        for (int c = 1; c <= 8 && c * 32 <= M->N[2]; c++) {
            const long long slots = 2LL * ctx->num_sms;
            const long long waves = ((long long)ntx * nty * c + slots - 1) / slots;
            const long long cost = waves * ((M->N[2] + c - 1) / c + C::WARM) * (c == 1 ? 106 : 100);
            if (best < 0 || cost < best) { best = cost; nzc = c; }
        }
    }
    a.zchunk = (M->N[2] + nzc - 1) / nzc;
    nzc = (M->N[2] + a.zchunk - 1) / a.zchunk;
    kern<<<dim3(ntx, nty, nzc), C::NTHREADS, C::SMEM, ctx->stream>>>(tm, a);
    launch_count(ctx, "varc3d_tma");
}

bool varc3d_lapw_supported(const judi_model *M)
{
    return M->ndim == 3 && (M->r == 4 || M->r == 2) && M->ctx->encode_tiled != nullptr;
}

void varc3d_lapw(const judi_model *M, const CUtensorMap &umap, const CUtensorMap &cmap, float *out)
{
    if (M->r == 2) varc3d_lapw_r<2>(M, umap, cmap, out);
    else varc3d_lapw_r<4>(M, umap, cmap, out);
}

// One fused variable-density step; false if this combination has no single-pass instantiation.
bool varc3d_step(const judi_model *M, const StepArgs &s)
{
    if (!varc3d_supported(M) || s.qext || s.f.wrec || !s.scratch || !s.scratch->has_tmap) return false;
    if (s.f.grad && s.f.ic != JUDI_IC_AS) return false;
    const int mask = (s.f.snap[0] ? VF_SNAP : 0) | (s.f.grad ? VF_GRAD : 0) | (s.f.illum ? VF_ILLUM : 0);
    if (s.born) {
        // Born step: two single-pass launches (background writes delta, linearised field reads it)
        if (s.born_ic != JUDI_IC_AS || s.f.grad || !s.lin || !s.lin->has_tmap || !s.scratch->dd[0].p || !M->dm.p)
            return false;
        switch (mask) {
        case 0: varc3d_launch<VF_DDOUT>(M, s); break;
        case VF_SNAP: varc3d_launch<VF_DDOUT | VF_SNAP>(M, s); break;
        case VF_ILLUM: varc3d_launch<VF_DDOUT | VF_ILLUM>(M, s); break;
        case VF_SNAP | VF_ILLUM: varc3d_launch<VF_DDOUT | VF_SNAP | VF_ILLUM>(M, s); break;
        default: return false;
        }
        varc3d_launch<VF_QSRC>(M, s, true);
        return true;
    }
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
