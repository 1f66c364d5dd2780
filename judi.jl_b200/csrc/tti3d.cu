// tti3d.cu -- single-pass 3-D TTI time step for sm_100a (space orders 4 and 8: a template on the
// stencil radius R; constant density or, RHO, a density field; the figures below are for R = 4,
// constant density).
// Replaces the Devito loop nest of tti_kernel (src/pysource/kernels.py:144-182 with the rotated
// self-adjoint operator of FD_utils.py:24-105) -- and the two-pass form of flux3d.cu, whose six
// flux fields cost 48 of its 119 B per grid point in HBM round trips and whose synchronous
// global loads leave it latency-bound.
//
//   H0 = sum_d D-_d P_d,  H1 = sum_d D-_d M_d,   (P, M) = R^T [[A,B],[B,C]] R (G u1, G u2)
//
// needs the flux one stencil radius around every output point, and every flux node needs the full
// staggered gradient of both fields.  One CTA (12 warps) owns a 32 x 16 (x,y) tile and marches
// along z.  Everything it reads arrives by TMA (one producer warp, three mbarrier rings): the
// haloed planes of u1, u2 (48 x 30), the six parameter planes A0, B0, C0, r (40 x 23, precomputed
// per model, model.cu) and the tiles of u1(t-1), u2(t-1), m.
//
//  F warps (7)  216 threads, one per group of 4 x-points of the plus-shaped region (tile 128 +
//     x halo 32 + y halo 56), keep the z-lines of u1, u2 of their own column in register queues,
//     take the x/y lines from the shared plane, evaluate the closed-form rotation (flux3d.cu) and
//     publish P_x, M_x (x-extended), P_y, M_y (y-extended), P_z, M_z and u1, u2 of the output plane
//     (tile only) in a double-buffered shared plane;
//  D warps (4)  128 threads, one per tile group: D-_x + D-_y of both fields from the shared plane go
//     to the register accumulator of that plane; the z part never meets memory again: P_z, M_z of
//     flux plane zf are scattered into the 8 accumulators of the output planes zf-3 .. zf+4, so
//     plane zf-3 is complete when flux plane zf is, and is updated in place (+ fused snapshot save /
//     imaging condition / illumination, whose global loads are issued before the hand-over wait).
//  The two roles run one plane apart and hand the flux plane over through two pairs of named
//  barriers (bar.arrive / bar.sync: "full" F -> D, "free" D -> F), never a CTA-wide barrier.
//
// HBM traffic: 48 B per grid point (u1, u2, their previous level, m, six parameters: read; two
// writes) + the plane halos that miss L2; the price is evaluating the flux on 864 instead of 512
// nodes per tile plane (1.69x).
//
// Measured (C5 grid, 40.8 M points): 0.69-0.72 ms/step plain, 0.74 save, 0.83 imaging condition,
// against 1.15 / 1.15 / 1.17 ms for the two passes.  ncu (profiles/r01_ncu_tti3d.md): 339 M warp
// instructions, 46% issue-active with 12 warps per SM (160-168 registers: the cap at 384 threads),
// shared-memory pipe 60% busy (1344 wavefronts per tile plane: 26 LDS.128 per flux group, 29 per
// output group), DRAM 3.7 TB/s: bound by shared-memory bandwidth and issue together, not by HBM.
//
// History: (1) the first single-pass kernel of the round (64 x 8 tile, 2.1x redundant flux, flux z
// ring and parked partial sums in shared memory, D phase split by field) executed 660 M warp
// instructions per step -- 115 M of them register-queue moves, 50 M spin-wait branches, local-
// memory traffic for the dynamically indexed rings -- and ran 1.29 ms; (2) the same mathematics
// with the bookkeeping removed but F and D in the same 4 tile warps behind one barrier (245
// registers, 8 warps) ran 0.85 ms: the tile warps' serial F + D chain at 0.2 IPC was the critical
// path; (3) splitting the roles over 11 compute warps gave 0.72 ms.  The compile-time rotated
// queues (ROT, option tti_fused=1) avoid 64 register moves per thread and plane but run slower
// (0.80 ms) than the shifted queues (tti_fused=2, the default): 8x the code.  Round 2 bracketed the
// design from both sides (profiles/r02_tti_tune.log): thinner threads (tti3d_h below: twice the warps,
// +15 % shared-memory wavefronts) are 3-7 % slower, fatter ones (two rows per D thread: -7 %
// wavefronts, half the D warps) 1.4-2.3x slower -- F and D are a two-stage pipeline one plane
// apart, each stage a serial per-warp chain, under a shared-memory pipe that is 72 % busy.
#include "tma_util.cuh"

// Geometry of one CTA for stencil radius R = space_order / 2 (2 or 4).  The x layout works in groups
// of 4 points and keeps one halo group per side whatever R <= 4 is; the y / z extents follow R.
// RHO: the model has a density field -- a seventh parameter plane carries b = 1/rho to the flux
// nodes, the tile ring carries m * b instead of m, and the D role reads b of its own points from
// global memory where a flavour needs it (imaging condition, Born source).
template <int R_, bool RHO_ = false>
struct TtiC {
    static constexpr int R = R_;
    static constexpr int TXG = 8, TX = 4 * TXG, TY = 16;      // tile: 8 groups of 4 x-points, 16 rows
    static constexpr int UPX = TX + 16, UPY = TY + 4 * R - 2; // haloed u plane, origin (x0-8, y0-(2R-1))
    static constexpr int USLOT = ((UPX * UPY + 31) / 32) * 32;
    static constexpr int CPX = TX + 8, CPY = TY + 2 * R - 1;  // parameter plane, origin (x0-4, y0-R)
    static constexpr int CSLOT = ((CPX * CPY + 31) / 32) * 32;  // 128-byte aligned slots
    static constexpr int NPAR = RHO_ ? 7 : 6;                 // A0, B0, C0, rx, ry, rz [, b]
    static constexpr int PSLOT = TX * TY;
    static constexpr int NS = 8, NC = 3, NP = 3;              // ring depths: u planes, parameter planes, tiles
    static constexpr int NQ = 2 * R;                          // z-queue: planes zf-(R-1) .. zf+R
    static constexpr int NI = TXG * TY;                       // 128 tile groups
    static constexpr int NYH = (((2 * R - 1) * TXG + 31) / 32) * 32;  // y-halo threads (whole warps)
    static constexpr int NF = NI + 32 + NYH;                  // F warps: tile, x halo, y halo
    static constexpr int ND = NI;                             // D warps: one thread per tile group
    static constexpr int NTHREADS = NF + ND + 32;             // + the producer warp
    static constexpr int FXW = TX + 8;                        // row pitch of the x-extended flux plane (40)
    static constexpr int FX = TY * FXW;                       // floats per field
    static constexpr int FY = CPY * TX;                       // y-extended flux plane (rows -R .. TY+R-2)
    static constexpr int FBUF = 2 * FX + 2 * FY + 4 * PSLOT;  // P_x, M_x, P_y, M_y of one plane + P_z, M_z, u1, u2 tiles
    static constexpr int LEADU = 3, LEADC = 2;                // producer run-ahead (planes)
    static constexpr int WARM = 4 * R - 2;                    // iterations before the first output plane
    static constexpr size_t SMEM = (size_t)(NS * 2 * USLOT + NC * NPAR * CSLOT + NP * 3 * PSLOT + 2 * FBUF) * 4 +
                                   (2 * NS + 2 * NC + 2 * NP) * 8 + 128;
    static_assert(R == 2 || R == 4, "x halo of one 4-point group: R <= 4");
    static_assert(R + 1 + LEADU <= NS, "u ring too short");
};
constexpr int TTI_NPAR = 7;

struct TtiMaps {
    CUtensorMap U1, U2;        // haloed planes of u1(t), u2(t)
    CUtensorMap C[TTI_NPAR];  // parameter planes
    CUtensorMap P1, P2, M;     // tiles of u1(t-1), u2(t-1), m
};

struct TtiArgs {
    Grid g;
    Coefs cf;
    float *up1, *up2;
    const float *dx, *dy, *dz;
    float irho_c;
    int zchunk;
    Fuse f;
    float *dd_out[2];        // Born background pass: write delta = u+ - 2u + u- of both fields
    const float *dd_in[2];   // Born linearised pass: delta of the background fields ...
    const float *dm;         // ... and the model perturbation: q_f = -dm irho delta_f
    const float *irho;       // RHO: b = 1/rho (Grid layout), read by the D role of the IC / Born flavours
};

// what rides on the step (compile-time mask): snapshot save, imaging condition, illumination,
// Born delta out (background field) / Born source in (linearised field)
enum { TF_SNAP = 1, TF_GRAD = 2, TF_ILLUM = 4, TF_DDOUT = 8, TF_QSRC = 16 };

// staggered first derivative along x at the 4 nodes of a group from the 12-float row window
// centred on it: PLUS: D+ (node + 1/2), sum_k w_k (u[e+k] - u[e-k+1]); else D- (node - 1/2),
// sum_k w_k (F[e+k-1] - F[e-k]).  Odd offsets straddle register pairs: scalar arithmetic.
template <bool PLUS, int R>
__device__ __forceinline__ void tti_dx(const float *wx, const float4 &l, const float4 &c,
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

// UF / UD: how the register queues of the F / D role advance: 1 = shifted by one slot every plane
// (2R-1 float4 moves per field), 2 = main loop unrolled by two with one spare slot, shifted by two
// every second plane (half the moves, twice the code), NQ = rotated at compile time (no moves,
// NQ times the code).
template <int R, int FUSE, int UF, int UD, bool RHO = false>
__global__ void __launch_bounds__(TtiC<R>::NTHREADS, 1)
    tti3d_tma(const __grid_constant__ TtiMaps tm, const TtiArgs a)
{
    typedef TtiC<R, RHO> C;
    constexpr int TXG = C::TXG, TX = C::TX, TY = C::TY, UPX = C::UPX, USLOT = C::USLOT, CPX = C::CPX;
    constexpr int CPY = C::CPY, CSLOT = C::CSLOT, NPAR = C::NPAR, PSLOT = C::PSLOT, NS = C::NS, NC = C::NC;
    constexpr int NP = C::NP, NQ = C::NQ, NI = C::NI, NF = C::NF, ND = C::ND, FXW = C::FXW, FX = C::FX;
    constexpr int FY = C::FY, FBUF = C::FBUF, LEADU = C::LEADU, LEADC = C::LEADC, WARM = C::WARM;
    static_assert(!(UF == NQ || UD == NQ) || NS == NQ, "rotated queues need ring length == queue length");
    extern __shared__ unsigned char smem_raw[];
    float *ringU = (float *)(smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u));
    float *ringC = ringU + NS * 2 * USLOT;       // [NC][NPAR][CSLOT]
    float *ringP = ringC + NC * NPAR * CSLOT;    // [NP][3][PSLOT]: u1-, u2-, m
    float *fbuf = ringP + NP * 3 * PSLOT;        // [2][FBUF]
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
    constexpr int FIRST = 2 * R - 1;             // first flux iteration
    constexpr int NSYNC = NF + ND;               // threads of the F <-> D hand-over barriers

    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; s++) { mbar_init(&fullU[s], 1); mbar_init(&emptyU[s], NF / 32); }
        for (int s = 0; s < NC; s++) { mbar_init(&fullC[s], 1); mbar_init(&emptyC[s], NF / 32); }
        for (int s = 0; s < NP; s++) { mbar_init(&fullP[s], 1); mbar_init(&emptyP[s], ND / 32); }
        fence_barrier_init();
    }
    __syncthreads();
    const int tid = threadIdx.x, lane = tid & 31;

    if (tid >= NF + ND) {
        // ===== producer warp: one lane issues the three streams, each a fixed distance ahead ======
        if (lane == 0) {
            const int bx = g.hx + x0 - 8, by = g.hy + y0 - (2 * R - 1), bz = g.hz + z0 - (2 * R - 1);
            const int cx0 = g.hx + x0 - 4, cy0 = g.hy + y0 - R, cz0 = g.hz + z0 - R;
            const int tx0 = g.hx + x0, ty0 = g.hy + y0, tz0 = g.hz + z0;
            for (int j = -LEADU; j < total; j++) {
                const int pj = j + LEADU;                 // u plane (consumer iteration pj)
                if (pj < total) {
                    const int s = pj % NS;
                    if (pj >= NS) mbar_wait_sleep(&emptyU[s], (uint32_t)(((pj / NS) - 1) & 1));
                    mbar_arrive_expect_tx(&fullU[s], 2 * UPX * C::UPY * 4);
                    tma_load_3d(ringU + s * 2 * USLOT, &tm.U1, bx, by, bz + pj, &fullU[s]);
                    tma_load_3d(ringU + s * 2 * USLOT + USLOT, &tm.U2, bx, by, bz + pj, &fullU[s]);
                }
                const int c = j + LEADC - FIRST;          // flux plane counter (iteration c + 7)
                if (c >= 0 && c < total - FIRST) {
                    const int s = c % NC;
                    if (c >= NC) mbar_wait_sleep(&emptyC[s], (uint32_t)(((c / NC) - 1) & 1));
                    mbar_arrive_expect_tx(&fullC[s], NPAR * CPX * CPY * 4);
#pragma unroll
                    for (int p = 0; p < NPAR; p++)
                        tma_load_3d(ringC + (s * NPAR + p) * CSLOT, &tm.C[p], cx0, cy0, cz0 + c, &fullC[s]);
                }
                const int i = j + LEADC - WARM;           // output plane counter (iteration i + 14)
                if (i >= 0 && i < nzc) {
                    const int s = i % NP;
                    if (i >= NP) mbar_wait_sleep(&emptyP[s], (uint32_t)(((i / NP) - 1) & 1));
                    mbar_arrive_expect_tx(&fullP[s], 3 * PSLOT * 4);
                    tma_load_3d(ringP + s * 3 * PSLOT, &tm.P1, tx0, ty0, tz0 + i, &fullP[s]);
                    tma_load_3d(ringP + s * 3 * PSLOT + PSLOT, &tm.P2, tx0, ty0, tz0 + i, &fullP[s]);
                    tma_load_3d(ringP + s * 3 * PSLOT + 2 * PSLOT, &tm.M, tx0, ty0, tz0 + i, &fullP[s]);
                }
            }
        }
        return;
    }

    // register slot of plane zf + k, k = -3 .. 4 (rotated at compile time, or shifted every plane)
#define SL(k) (ROT ? ((ph + (k) + 2 * NQ - R) % NQ) : (ph + (k) + R - 1))

    if (tid < NF) {
        constexpr bool ROT = UF == NQ;
        constexpr int UNR = UF;
        constexpr int NQL = ROT ? NQ : NQ + UNR - 1;
        // ===== F role: flux node group cx in -1..8 (groups of 4 x-points), cy in -4..18 ============
        const bool tile = tid < NI;                    // warps 0-3 (warp uniform)
        int cx, cy;
        bool wr_x = false, wr_y = false;
        if (tile) { cx = tid & (TXG - 1); cy = tid / TXG; wr_x = wr_y = true; }
        else if (tid < NI + 32) {
            const int h = tid - NI;
            cx = (h & 1) ? TXG : -1;
            cy = h >> 1;
            wr_x = true;
        } else {
            const int h = tid - NI - 32, row = h / TXG;   // rows beyond 2R-2 are idle (mirror the last one)
            cx = h & (TXG - 1);
            cy = row < R ? row - R : TY + min(row, 2 * R - 2) - R;
            wr_y = row < 2 * R - 1;
        }
        const int su = (cy + 2 * R - 1) * UPX + 4 * cx + 8;       // node group inside a u plane
        const int sc = (cy + R) * CPX + 4 * cx + 4;               // ... inside a parameter plane
        const int sfx = cy * FXW + 4 * (cx + 1);                  // ... inside the x-extended flux plane
        const int sfy = (cy + R) * TX + 4 * cx;                   // ... inside the y-extended flux plane
        const int st = cy * TX + 4 * cx;                          // ... inside a tile (tile threads)
        const f2 bb = pk1(a.irho_c);

        float4 q1[NQL], q2[NQL];   // u1, u2 of planes zf-3 .. zf+4 at this node group
#pragma unroll
        for (int k = 0; k < NQL; k++) q1[k] = q2[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        int cslot = 0, cpar = 0;   // parameter ring position / parity

        for (int j0 = 0; j0 < total; j0 += UNR) {
#pragma unroll
            for (int ph = 0; ph < UNR; ph++) {
                const int j = j0 + ph;
                if (UNR > 1 && j >= total) break;
                if (!ROT && ph == 0) {
#pragma unroll
                    for (int k = 0; k < NQL - UNR; k++) { q1[k] = q1[k + UNR]; q2[k] = q2[k + UNR]; }
                }
                const int s_in = ROT ? ph : (j & (NS - 1));
                mbar_wait(&fullU[s_in], (uint32_t)((j / NS) & 1));
                q1[SL(R)] = *(const float4 *)(ringU + s_in * 2 * USLOT + su);
                q2[SL(R)] = *(const float4 *)(ringU + s_in * 2 * USLOT + USLOT + su);
                if (j < FIRST) {
                    // the first three planes of a chunk are never a flux plane: release them now
                    if (j < R - 1) {
                        __syncwarp();
                        mbar_arrive_if(&emptyU[s_in], lane == 0);
                    }
                    continue;
                }
                // ---- flux at plane zf (brought in R iterations ago) ----------------------------
                const int zf = z0 - (3 * R - 1) + j;
                const bool need_xy = zf >= z0 && zf < z1;       // its x / y flux is used
                const bool work = tile || need_xy;
                const int s_c = ROT ? (ph + NS - R) % NS : ((j + NS - R) & (NS - 1));
                float *fb = fbuf + (j & 1) * FBUF;
                f2 P[3][2], Mf[3][2];
                mbar_wait(&fullC[cslot], (uint32_t)cpar);
                if (work) {
                    f2 G[2][3][2];   // [field][x,y,z][pair]
#pragma unroll
                    for (int f = 0; f < 2; f++) {
                        const float *pl = ringU + s_c * 2 * USLOT + f * USLOT + su;
                        const float4 *q = f ? q2 : q1;
                        const float4 c = q[SL(0)];
                        tti_dx<true, R>(cf.wx, *(const float4 *)(pl - 4), c, *(const float4 *)(pl + 4),
                                     G[f][0][0], G[f][0][1]);
                        f2 gy0, gy1, gz0, gz1;
#pragma unroll
                        for (int k = 1; k <= R; k++) {
                            const float4 p = *(const float4 *)(pl + k * UPX);
                            const float4 m = (k == 1) ? c : *(const float4 *)(pl - (k - 1) * UPX);
                            const f2 dy0 = sub2(lo2(p), lo2(m)), dy1 = sub2(hi2(p), hi2(m));
                            const float4 zp = q[SL(k)], zm = q[SL(1 - k)];
                            const f2 dz0 = sub2(lo2(zp), lo2(zm)), dz1 = sub2(hi2(zp), hi2(zm));
                            if (k == 1) {
                                gy0 = mul2s(cf.wy[1], dy0); gy1 = mul2s(cf.wy[1], dy1);
                                gz0 = mul2s(cf.wz[1], dz0); gz1 = mul2s(cf.wz[1], dz1);
                            } else {
                                gy0 = fma2s(cf.wy[k], dy0, gy0); gy1 = fma2s(cf.wy[k], dy1, gy1);
                                gz0 = fma2s(cf.wz[k], dz0, gz0); gz1 = fma2s(cf.wz[k], dz1, gz1);
                            }
                        }
                        G[f][1][0] = gy0; G[f][1][1] = gy1;
                        G[f][2][0] = gz0; G[f][2][1] = gz1;
                    }
                    // closed-form rotation (flux3d.cu): P = A0 g1 + B0 g2 + sP r, M = B0 g1 + C0 g2 + sM r
                    const float *pc = ringC + cslot * NPAR * CSLOT + sc;
                    const float4 A4 = *(const float4 *)pc, B4 = *(const float4 *)(pc + CSLOT);
                    const float4 C4 = *(const float4 *)(pc + 2 * CSLOT);
                    const float4 rx4 = *(const float4 *)(pc + 3 * CSLOT), ry4 = *(const float4 *)(pc + 4 * CSLOT);
                    const float4 rz4 = *(const float4 *)(pc + 5 * CSLOT);
                    [[maybe_unused]] float4 b4;
                    if (RHO) b4 = *(const float4 *)(pc + 6 * CSLOT);
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const f2 A0 = h ? hi2(A4) : lo2(A4), B0 = h ? hi2(B4) : lo2(B4);
                        const f2 C0 = h ? hi2(C4) : lo2(C4);
                        const f2 rx = h ? hi2(rx4) : lo2(rx4), ry = h ? hi2(ry4) : lo2(ry4);
                        const f2 rz = h ? hi2(rz4) : lo2(rz4);
                        const f2 d1 = fma2(rz, G[0][2][h], fma2(ry, G[0][1][h], mul2(rx, G[0][0][h])));
                        const f2 d2 = fma2(rz, G[1][2][h], fma2(ry, G[1][1][h], mul2(rx, G[1][0][h])));
                        const f2 nB0 = mul2s(-1.f, B0);
                        const f2 bh = RHO ? (h ? hi2(b4) : lo2(b4)) : bb;
                        const f2 sP = fma2(nB0, d2, mul2(sub2(bh, A0), d1));
                        const f2 sM = fma2(nB0, d1, mul2(mul2s(-1.f, C0), d2));
                        P[0][h] = fma2(sP, rx, fma2(B0, G[1][0][h], mul2(A0, G[0][0][h])));
                        P[1][h] = fma2(sP, ry, fma2(B0, G[1][1][h], mul2(A0, G[0][1][h])));
                        P[2][h] = fma2(sP, rz, fma2(B0, G[1][2][h], mul2(A0, G[0][2][h])));
                        Mf[0][h] = fma2(sM, rx, fma2(C0, G[1][0][h], mul2(B0, G[0][0][h])));
                        Mf[1][h] = fma2(sM, ry, fma2(C0, G[1][1][h], mul2(B0, G[0][1][h])));
                        Mf[2][h] = fma2(sM, rz, fma2(C0, G[1][2][h], mul2(B0, G[0][2][h])));
                    }
                }
                // plane zf and its parameters are dead for this warp
                __syncwarp();
                mbar_arrive_if(&emptyU[s_c], lane == 0); mbar_arrive_if(&emptyC[cslot], lane == 0);
                if (++cslot == NC) { cslot = 0; cpar ^= 1; }
                // the flux buffer of this parity was read by the D warps two planes ago
                if (j >= FIRST + 2) named_bar_sync(3 + (j & 1), NSYNC);
                if (work) {
                    if (need_xy) {
                        if (wr_x) {
                            *(float4 *)(fb + sfx) = cat4(P[0][0], P[0][1]);
                            *(float4 *)(fb + FX + sfx) = cat4(Mf[0][0], Mf[0][1]);
                        }
                        if (wr_y) {
                            *(float4 *)(fb + 2 * FX + sfy) = cat4(P[1][0], P[1][1]);
                            *(float4 *)(fb + 2 * FX + FY + sfy) = cat4(Mf[1][0], Mf[1][1]);
                        }
                    }
                    if (tile) {
                        float *fz_ = fb + 2 * FX + 2 * FY + st;
                        *(float4 *)fz_ = cat4(P[2][0], P[2][1]);
                        *(float4 *)(fz_ + PSLOT) = cat4(Mf[2][0], Mf[2][1]);
                        // u1, u2 of the output plane zf-3 (the oldest queue entries)
                        *(float4 *)(fz_ + 2 * PSLOT) = q1[SL(1 - R)];
                        *(float4 *)(fz_ + 3 * PSLOT) = q2[SL(1 - R)];
                    }
                }
                __threadfence_block();
                named_bar_arrive(1 + (j & 1), NSYNC);
            }
        }
        return;
    }

    // ===== D role: divergence, time update and fused extras of one tile group ======================
    constexpr bool ROT = UD == NQ;
    constexpr int UNR = UD;
    constexpr int NQL = ROT ? NQ : NQ + UNR - 1;
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
    constexpr bool f_snap = (FUSE & TF_SNAP) != 0, f_grad = (FUSE & TF_GRAD) != 0;
    constexpr bool f_il = (FUSE & TF_ILLUM) != 0, f_reg = (FUSE & (TF_SNAP | TF_GRAD | TF_ILLUM)) != 0;
    constexpr bool f_dd = (FUSE & TF_DDOUT) != 0, f_q = (FUSE & TF_QSRC) != 0;
    [[maybe_unused]] const Fuse &fz = a.f;
    [[maybe_unused]] const bool f_xy = f_reg && act && oy >= fz.reg.lo[1] && oy < fz.reg.hi[1] &&
                                       ox + 3 >= fz.reg.lo[0] && ox < fz.reg.hi[0];
    [[maybe_unused]] const bool f_vec = f_xy && (fz.reg.lo[0] & 3) == 0 && ox >= fz.reg.lo[0] &&
                                        ox + 3 < fz.reg.hi[0];
    [[maybe_unused]] long long ri = f_reg ? fz.reg.at(ox, oy, z0) : 0;
    [[maybe_unused]] const long long rstride = f_reg ? (long long)fz.reg.pitch * fz.reg.ext[1] : 0;
    [[maybe_unused]] const float gc = f_grad ? fz.gcoef * a.irho_c : 0.f;

    float4 aP[NQL], aM[NQL];   // divergence accumulators of the output planes zf-3 .. zf+4
#pragma unroll
    for (int k = 0; k < NQL; k++) aP[k] = aM[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    int pslot = 0, ppar = 0;   // tile ring position / parity

    for (int j0 = FIRST; j0 < total; j0 += UNR) {
#pragma unroll
        for (int ph = 0; ph < UNR; ph++) {
            const int j = j0 + ph;
            if (UNR > 1 && j >= total) break;
            const int zf = z0 - (3 * R - 1) + j;
            const bool need_xy = zf >= z0 && zf < z1;
            const bool out = j >= WARM;
            const int z = z0 + j - WARM;                  // output plane zo = zf - 3
            const float *fb = fbuf + (j & 1) * FBUF;
            // ---- everything that comes from global memory is requested before the hand-over wait
            float dz_s = 0.f;
            [[maybe_unused]] bool fin = false, fvec = false;
            [[maybe_unused]] float4 us1, us2, G4, I4;
            [[maybe_unused]] float4 dq1 = make_float4(0.f, 0.f, 0.f, 0.f), dq2 = dq1, dm4 = dq1;
            [[maybe_unused]] float4 ir4 = make_float4(a.irho_c, a.irho_c, a.irho_c, a.irho_c);
            if (out) {
                dz_s = __ldg(a.dz + z);
                // (a partial group at the x edge reads into the replicated halo of the allocation)
                if (RHO && (f_q || f_grad) && act) ir4 = __ldg((const float4 *)(a.irho + gout));
                if (f_q && act) {
                    // (a partial group at the x edge reads into the zero halo of the allocation)
                    dq1 = *(const float4 *)(a.dd_in[0] + gout);
                    dq2 = *(const float4 *)(a.dd_in[1] + gout);
                    dm4 = __ldg((const float4 *)(a.dm + gout));
                }
                if (f_reg) {
                    const bool zin = z >= fz.reg.lo[2] && z < fz.reg.hi[2];
                    fin = f_xy && zin;
                    fvec = f_vec && zin;
                    if ((f_grad || f_il) && f_vec && z + 3 >= fz.reg.lo[2] && z + 3 < fz.reg.hi[2]) {
                        if (f_grad) {
                            prefetch_l2(fz.usnap[0] + ri + 3 * rstride);
                            prefetch_l2(fz.usnap[1] + ri + 3 * rstride);
                            prefetch_l2(fz.grad + ri + 3 * rstride);
                        }
                        if (f_il) prefetch_l2(fz.illum + ri + 3 * rstride);
                    }
                    if (fvec) {
                        if (f_grad) {
                            us1 = __ldg((const float4 *)(fz.usnap[0] + ri));
                            us2 = __ldg((const float4 *)(fz.usnap[1] + ri));
                            G4 = *(const float4 *)(fz.grad + ri);
                        }
                        if (f_il) I4 = *(const float4 *)(fz.illum + ri);
                    }
                }
            }
            named_bar_sync(1 + (j & 1), NSYNC);
            // ---- D-_x + D-_y of flux plane zf go to the accumulator of output plane zf ------------
            if (need_xy) {
#pragma unroll
                for (int f = 0; f < 2; f++) {
                    const float *px = fb + f * FX + sfx;
                    f2 lx0, lx1;
                    tti_dx<false, R>(cf.wx, *(const float4 *)(px - 4), *(const float4 *)px,
                                  *(const float4 *)(px + 4), lx0, lx1);
                    const float *py = fb + 2 * FX + f * FY + sfy;
                    f2 ly0, ly1;
#pragma unroll
                    for (int k = 1; k <= R; k++) {
                        const float4 p = *(const float4 *)(py + (k - 1) * TX);
                        const float4 m = *(const float4 *)(py - k * TX);
                        const f2 d0 = sub2(lo2(p), lo2(m)), d1 = sub2(hi2(p), hi2(m));
                        if (k == 1) { ly0 = mul2s(cf.wy[1], d0); ly1 = mul2s(cf.wy[1], d1); }
                        else { ly0 = fma2s(cf.wy[k], d0, ly0); ly1 = fma2s(cf.wy[k], d1, ly1); }
                    }
                    float4 &acc = f ? aM[SL(0)] : aP[SL(0)];
                    acc = cat4(add2(lo2(acc), add2(lx0, ly0)), add2(hi2(acc), add2(lx1, ly1)));
                }
            }
            // ---- z part: D-_z out[z] = sum_k wz[k] (F[z+k-1] - F[z-k]), scattered from F = P_z, M_z
            {
                const float *fz_ = fb + 2 * FX + 2 * FY + st;
                const float4 Pz = *(const float4 *)fz_, Mz = *(const float4 *)(fz_ + PSLOT);
#pragma unroll
                for (int k = 1; k <= R; k++) {
                    float4 &ap = aP[SL(1 - k)], &am = aM[SL(1 - k)];
                    ap = cat4(fma2s(cf.wz[k], lo2(Pz), lo2(ap)), fma2s(cf.wz[k], hi2(Pz), hi2(ap)));
                    am = cat4(fma2s(cf.wz[k], lo2(Mz), lo2(am)), fma2s(cf.wz[k], hi2(Mz), hi2(am)));
                    float4 &bp = aP[SL(k)], &bm = aM[SL(k)];
                    if (k == R) {   // first contribution to plane zf+4: initialises its accumulator
                        bp = cat4(mul2s(-cf.wz[k], lo2(Pz)), mul2s(-cf.wz[k], hi2(Pz)));
                        bm = cat4(mul2s(-cf.wz[k], lo2(Mz)), mul2s(-cf.wz[k], hi2(Mz)));
                    } else {
                        bp = cat4(fma2s(-cf.wz[k], lo2(Pz), lo2(bp)), fma2s(-cf.wz[k], hi2(Pz), hi2(bp)));
                        bm = cat4(fma2s(-cf.wz[k], lo2(Mz), lo2(bm)), fma2s(-cf.wz[k], hi2(Mz), hi2(bm)));
                    }
                }
            }
            if (out) {
                // ---- output plane zo = zf - 3: its divergence is complete -------------------------
                const float *fz_ = fb + 2 * FX + 2 * FY + st;
                const float4 uc1 = *(const float4 *)(fz_ + 2 * PSLOT), uc2 = *(const float4 *)(fz_ + 3 * PSLOT);
                const float4 H1 = aP[SL(1 - R)], H2 = aM[SL(1 - R)];
                mbar_wait(&fullP[pslot], (uint32_t)ppar);
                const float *pt = ringP + pslot * 3 * PSLOT + st;
                const float4 up1 = *(const float4 *)pt, up2 = *(const float4 *)(pt + PSLOT);
                const float4 m4 = *(const float4 *)(pt + 2 * PSLOT);
                __syncwarp();
                mbar_arrive_if(&emptyP[pslot], lane == 0);
                if (++pslot == NP) { pslot = 0; ppar ^= 1; }
                // delta = (dt^2 H - sd (u - u-)) / (m irho + sd),  u+ = (u + (u - u-)) + delta
                const f2 dzv = pk1(dz_s);
                const f2 sd0 = mul2s(cf.dt, add2(dxy0, dzv)), sd1 = mul2s(cf.dt, add2(dxy1, dzv));
                // RHO: the tile ring carries m * b
                const f2 den0 = RHO ? add2(lo2(m4), sd0) : fma2s(a.irho_c, lo2(m4), sd0);
                const f2 den1 = RHO ? add2(hi2(m4), sd1) : fma2s(a.irho_c, hi2(m4), sd1);
                const f2 r0 = pk(__frcp_rn(den0.x), __frcp_rn(den0.y));
                const f2 r1 = pk(__frcp_rn(den1.x), __frcp_rn(den1.y));
                const f2 nsd0 = mul2s(-1.f, sd0), nsd1 = mul2s(-1.f, sd1);
                f2 dl[2][2];
#pragma unroll
                for (int f = 0; f < 2; f++) {
                    const float4 uc = f ? uc2 : uc1, up = f ? up2 : up1, H = f ? H2 : H1;
                    const f2 d10 = sub2(lo2(uc), lo2(up)), d11 = sub2(hi2(uc), hi2(up));
                    f2 h0 = mul2s(cf.dt2, lo2(H)), h1 = mul2s(cf.dt2, hi2(H));
                    if (f_q) {
                        // dt^2 * lin_src = -dm irho (u+ - 2u + u-) of the background field
                        // (sensitivity.py:174-188; the post-injection part is added by sparse.cu)
                        const float4 dq = f ? dq2 : dq1;
                        const f2 w0 = RHO ? mul2(mul2s(-1.f, lo2(ir4)), lo2(dm4)) : mul2s(-a.irho_c, lo2(dm4));
                        const f2 w1 = RHO ? mul2(mul2s(-1.f, hi2(ir4)), hi2(dm4)) : mul2s(-a.irho_c, hi2(dm4));
                        h0 = fma2(w0, lo2(dq), h0);
                        h1 = fma2(w1, hi2(dq), h1);
                    }
                    // one explicit fma for the update: ptxas contracts a packed mul + add pair in the
                    // flavours that do not also store delta, and the field must not depend on the flavour
                    const f2 t0 = fma2(nsd0, d10, h0), t1 = fma2(nsd1, d11, h1);
                    dl[f][0] = mul2(t0, r0);
                    dl[f][1] = mul2(t1, r1);
                    const f2 un0 = fma2(t0, r0, add2(lo2(uc), d10)), un1 = fma2(t1, r1, add2(hi2(uc), d11));
                    if (act) {
                        float *o = (f ? a.up2 : a.up1) + gout;
                        if (xfull) {
                            *(float4 *)o = cat4(un0, un1);
                            if (f_dd) *(float4 *)(a.dd_out[f] + gout) = cat4(dl[f][0], dl[f][1]);
                        } else {
                            const float v[4] = {un0.x, un0.y, un1.x, un1.y};
                            const float dv[4] = {dl[f][0].x, dl[f][0].y, dl[f][1].x, dl[f][1].y};
#pragma unroll
                            for (int e = 0; e < 4; e++)
                                if (ox + e < g.n[0]) {
                                    o[e] = v[e];
                                    if (f_dd) a.dd_out[f][gout + e] = dv[e];
                                }
                        }
                    }
                }
                if (f_reg && fin) {
                    if (fvec) {
                        if (f_snap) {
                            __stcs((float4 *)(fz.snap[0] + ri), uc1);
                            __stcs((float4 *)(fz.snap[1] + ri), uc2);
                        }
                        if (f_grad) {
                            // gradient -= gcoef irho (us1 delta1 + us2 delta2)   (sensitivity.py:55-69)
                            const f2 ic0 = fma2(lo2(us2), dl[1][0], mul2(lo2(us1), dl[0][0]));
                            const f2 ic1 = fma2(hi2(us2), dl[1][1], mul2(hi2(us1), dl[0][1]));
                            if (RHO)
                                *(float4 *)(fz.grad + ri) = cat4(fma2(mul2s(-fz.gcoef, lo2(ir4)), ic0, lo2(G4)),
                                                                 fma2(mul2s(-fz.gcoef, hi2(ir4)), ic1, hi2(G4)));
                            else
                                *(float4 *)(fz.grad + ri) = cat4(fma2s(-gc, ic0, lo2(G4)), fma2s(-gc, ic1, hi2(G4)));
                        }
                        if (f_il) {
                            // illumination += dt (u1^2 + u2^2)   (fields_exprs.py:240-255)
                            const f2 e0 = fma2(lo2(uc2), lo2(uc2), mul2(lo2(uc1), lo2(uc1)));
                            const f2 e1 = fma2(hi2(uc2), hi2(uc2), mul2(hi2(uc1), hi2(uc1)));
                            *(float4 *)(fz.illum + ri) = cat4(fma2s(cf.dt, e0, lo2(I4)), fma2s(cf.dt, e1, hi2(I4)));
                        }
                    } else {
                        const float u1v[4] = {uc1.x, uc1.y, uc1.z, uc1.w}, u2v[4] = {uc2.x, uc2.y, uc2.z, uc2.w};
                        const float d1v[4] = {dl[0][0].x, dl[0][0].y, dl[0][1].x, dl[0][1].y};
                        const float d2v[4] = {dl[1][0].x, dl[1][0].y, dl[1][1].x, dl[1][1].y};
#pragma unroll
                        for (int e = 0; e < 4; e++) {
                            if (ox + e < fz.reg.lo[0] || ox + e >= fz.reg.hi[0]) continue;
                            if (f_snap) { fz.snap[0][ri + e] = u1v[e]; fz.snap[1][ri + e] = u2v[e]; }
                            if (f_grad) {
                                const float ic = __fmaf_rn(fz.usnap[1][ri + e], d2v[e],
                                                           __fmul_rn(fz.usnap[0][ri + e], d1v[e]));
                                const float irv[4] = {ir4.x, ir4.y, ir4.z, ir4.w};
                                const float gce = RHO ? __fmul_rn(fz.gcoef, irv[e]) : gc;
                                fz.grad[ri + e] = __fmaf_rn(-gce, ic, fz.grad[ri + e]);
                            }
                            if (f_il) {
                                const float e2 = __fmaf_rn(u2v[e], u2v[e], __fmul_rn(u1v[e], u1v[e]));
                                fz.illum[ri + e] = __fmaf_rn(cf.dt, e2, fz.illum[ri + e]);
                            }
                        }
                    }
                }
                gout += g.sz;
                if (f_reg) ri += rstride;
            }
            // this flux buffer may be overwritten by the F warps two planes from now
            if (j + 2 < total) named_bar_arrive(3 + (j & 1), NSYNC);
            if (!ROT && ph == UNR - 1) {
#pragma unroll
                for (int k = 0; k < NQL - UNR; k++) { aP[k] = aP[k + UNR]; aM[k] = aM[k + UNR]; }
            }
        }
    }
#undef SL
}

// ================================================================================================
// tti3d_h: the same step with HALF-WIDTH threads (2 x-points per thread instead of 4), R = 4.
// tti3d_tma is latency-bound, not issue-bound: 12 warps per SM at 160 registers, 46 % of the issue
// slots used, and halving its register-queue moves does not change its time (option tti_fused =
// 5/6/7).  What raises the number of warps is fewer registers per thread, and the z-queues of the
// two fields are 64 of them: with 2 x-points per thread they are 32, the kernel fits 23 warps
// (448 F + 256 D + 32 producer threads) under an 88-register cap.  The price: the x windows are no
// longer amortised over 4 points (+15 % shared-memory wavefronts).  Same shared-memory layout,
// producer and hand-over protocol as tti3d_tma<4>; plain / save / imaging-condition flavours.
namespace ttih {
typedef TtiC<4> C;
constexpr int R = 4;
constexpr int TXP = C::TX / 2;                  // 16 pairs per tile row
constexpr int NFT = TXP * C::TY;                // 256 tile threads
constexpr int NFX = 4 * C::TY;                  // x halo: pairs -2, -1, 16, 17 of every row
constexpr int NFY = 128;                        // y halo: 7 rows x 16 pairs (112) in whole warps
constexpr int NF = NFT + NFX + NFY;             // 448
constexpr int ND = NFT;                         // 256
constexpr int NTHREADS = NF + ND + 32;          // 736
}  // namespace ttih

// D+ (PLUS) / D- along x at the 2 nodes of a pair from the 10-float window w (w[4] = own.x)
template <bool PLUS>
__device__ __forceinline__ f2 ttih_dx(const float *wx, const float2 &l0, const float2 &l1, const float2 &c,
                                      const float2 &r0, const float2 &r1)
{
    const float w[10] = {l0.x, l0.y, l1.x, l1.y, c.x, c.y, r0.x, r0.y, r1.x, r1.y};
    float o[2];
#pragma unroll
    for (int e = 0; e < 2; e++) {
        float acc = 0.f;
#pragma unroll
        for (int k = 1; k <= ttih::R; k++) {
            const float d = PLUS ? __fsub_rn(w[4 + e + k], w[4 + e - k + 1])
                                 : __fsub_rn(w[4 + e + k - 1], w[4 + e - k]);
            acc = (k == 1) ? __fmul_rn(wx[1], d) : __fmaf_rn(wx[k], d, acc);
        }
        o[e] = acc;
    }
    return pk(o[0], o[1]);
}

template <int FUSE>
__global__ void __launch_bounds__(ttih::NTHREADS, 1)
    tti3d_h(const __grid_constant__ TtiMaps tm, const TtiArgs a)
{
    typedef TtiC<4> C;
    constexpr int R = 4;
    constexpr int TX = C::TX, TY = C::TY, UPX = C::UPX, USLOT = C::USLOT, CPX = C::CPX;
    constexpr int CPY = C::CPY, CSLOT = C::CSLOT, NPAR = C::NPAR, PSLOT = C::PSLOT, NS = C::NS, NC = C::NC;
    constexpr int NP = C::NP, NQ = C::NQ, FXW = C::FXW, FX = C::FX;
    constexpr int FY = C::FY, FBUF = C::FBUF, LEADU = C::LEADU, LEADC = C::LEADC, WARM = C::WARM;
    constexpr int TXP = ttih::TXP, NFT = ttih::NFT, NFX = ttih::NFX, NF = ttih::NF, ND = ttih::ND;
    extern __shared__ unsigned char smem_raw[];
    float *ringU = (float *)(smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u));
    float *ringC = ringU + NS * 2 * USLOT;
    float *ringP = ringC + NC * NPAR * CSLOT;
    float *fbuf = ringP + NP * 3 * PSLOT;
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
    const int total = nzc + WARM;
    constexpr int FIRST = 2 * R - 1;
    constexpr int NSYNC = NF + ND;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; s++) { mbar_init(&fullU[s], 1); mbar_init(&emptyU[s], NF / 32); }
        for (int s = 0; s < NC; s++) { mbar_init(&fullC[s], 1); mbar_init(&emptyC[s], NF / 32); }
        for (int s = 0; s < NP; s++) { mbar_init(&fullP[s], 1); mbar_init(&emptyP[s], ND / 32); }
        fence_barrier_init();
    }
    __syncthreads();
    const int tid = threadIdx.x, lane = tid & 31;

    if (tid >= NF + ND) {
        // ===== producer warp (identical to tti3d_tma<4>) ===========================================
        if (lane == 0) {
            const int bx = g.hx + x0 - 8, by = g.hy + y0 - (2 * R - 1), bz = g.hz + z0 - (2 * R - 1);
            const int cx0 = g.hx + x0 - 4, cy0 = g.hy + y0 - R, cz0 = g.hz + z0 - R;
            const int tx0 = g.hx + x0, ty0 = g.hy + y0, tz0 = g.hz + z0;
            for (int j = -LEADU; j < total; j++) {
                const int pj = j + LEADU;
                if (pj < total) {
                    const int s = pj % NS;
                    if (pj >= NS) mbar_wait_sleep(&emptyU[s], (uint32_t)(((pj / NS) - 1) & 1));
                    mbar_arrive_expect_tx(&fullU[s], 2 * UPX * C::UPY * 4);
                    tma_load_3d(ringU + s * 2 * USLOT, &tm.U1, bx, by, bz + pj, &fullU[s]);
                    tma_load_3d(ringU + s * 2 * USLOT + USLOT, &tm.U2, bx, by, bz + pj, &fullU[s]);
                }
                const int c = j + LEADC - FIRST;
                if (c >= 0 && c < total - FIRST) {
                    const int s = c % NC;
                    if (c >= NC) mbar_wait_sleep(&emptyC[s], (uint32_t)(((c / NC) - 1) & 1));
                    mbar_arrive_expect_tx(&fullC[s], NPAR * CPX * CPY * 4);
#pragma unroll
                    for (int p = 0; p < NPAR; p++)
                        tma_load_3d(ringC + (s * NPAR + p) * CSLOT, &tm.C[p], cx0, cy0, cz0 + c, &fullC[s]);
                }
                const int i = j + LEADC - WARM;
                if (i >= 0 && i < nzc) {
                    const int s = i % NP;
                    if (i >= NP) mbar_wait_sleep(&emptyP[s], (uint32_t)(((i / NP) - 1) & 1));
                    mbar_arrive_expect_tx(&fullP[s], 3 * PSLOT * 4);
                    tma_load_3d(ringP + s * 3 * PSLOT, &tm.P1, tx0, ty0, tz0 + i, &fullP[s]);
                    tma_load_3d(ringP + s * 3 * PSLOT + PSLOT, &tm.P2, tx0, ty0, tz0 + i, &fullP[s]);
                    tma_load_3d(ringP + s * 3 * PSLOT + 2 * PSLOT, &tm.M, tx0, ty0, tz0 + i, &fullP[s]);
                }
            }
        }
        return;
    }

    if (tid < NF) {
        // ===== F role: flux at the pair cx2 in -2..17 (x = 2 cx2), row cy in -4..18 ==================
        const bool tile = tid < NFT;
        int cx2, cy;
        bool wr_x = false, wr_y = false;
        if (tile) { cx2 = tid & (TXP - 1); cy = tid / TXP; wr_x = wr_y = true; }
        else if (tid < NFT + NFX) {
            const int h = tid - NFT, pp = h & 3;
            cx2 = pp < 2 ? pp - 2 : TXP - 2 + pp;
            cy = h >> 2;
            wr_x = true;
        } else {
            const int h = tid - NFT - NFX, row = h / TXP;   // rows beyond 2R-2 are idle (mirror the last one)
            cx2 = h & (TXP - 1);
            cy = row < R ? row - R : TY + min(row, 2 * R - 2) - R;
            wr_y = row < 2 * R - 1;
        }
        const int su = (cy + 2 * R - 1) * UPX + 2 * cx2 + 8;
        const int sc = (cy + R) * CPX + 2 * cx2 + 4;
        const int sfx = cy * FXW + 2 * cx2 + 4;
        const int sfy = (cy + R) * TX + 2 * cx2;
        const int st = cy * TX + 2 * cx2;
        const f2 bb = pk1(a.irho_c);

        float2 q1[NQ], q2[NQ];   // u1, u2 of planes zf-3 .. zf+4 at this pair
#pragma unroll
        for (int k = 0; k < NQ; k++) q1[k] = q2[k] = make_float2(0.f, 0.f);
        int cslot = 0, cpar = 0;

        for (int j = 0; j < total; j++) {
#pragma unroll
            for (int k = 0; k < NQ - 1; k++) { q1[k] = q1[k + 1]; q2[k] = q2[k + 1]; }
            const int s_in = j & (NS - 1);
            mbar_wait(&fullU[s_in], (uint32_t)((j / NS) & 1));
            q1[NQ - 1] = *(const float2 *)(ringU + s_in * 2 * USLOT + su);
            q2[NQ - 1] = *(const float2 *)(ringU + s_in * 2 * USLOT + USLOT + su);
            if (j < FIRST) {
                if (j < R - 1) {
                    __syncwarp();
                    mbar_arrive_if(&emptyU[s_in], lane == 0);
                }
                continue;
            }
            const int zf = z0 - (3 * R - 1) + j;
            const bool need_xy = zf >= z0 && zf < z1;
            const bool work = tile || need_xy;
            const int s_c = (j + NS - R) & (NS - 1);
            float *fb = fbuf + (j & 1) * FBUF;
            f2 G[2][3];
            mbar_wait(&fullC[cslot], (uint32_t)cpar);
            if (work) {
#pragma unroll
                for (int f = 0; f < 2; f++) {
                    const float *pl = ringU + s_c * 2 * USLOT + f * USLOT + su;
                    const float2 *q = f ? q2 : q1;
                    const float2 c = q[R - 1];
                    G[f][0] = ttih_dx<true>(cf.wx, *(const float2 *)(pl - 4), *(const float2 *)(pl - 2), c,
                                            *(const float2 *)(pl + 2), *(const float2 *)(pl + 4));
                    f2 gy, gz;
#pragma unroll
                    for (int k = 1; k <= R; k++) {
                        const float2 p = *(const float2 *)(pl + k * UPX);
                        const float2 m = (k == 1) ? c : *(const float2 *)(pl - (k - 1) * UPX);
                        const f2 dy = sub2(p, m);
                        const f2 dz = sub2(q[R - 1 + k], q[R - k]);
                        if (k == 1) { gy = mul2s(cf.wy[1], dy); gz = mul2s(cf.wz[1], dz); }
                        else { gy = fma2s(cf.wy[k], dy, gy); gz = fma2s(cf.wz[k], dz, gz); }
                    }
                    G[f][1] = gy;
                    G[f][2] = gz;
                }
            }
            // the flux buffer of this parity was read by the D warps two planes ago
            if (j >= FIRST + 2) named_bar_sync(3 + (j & 1), NSYNC);
            if (work) {
                // closed-form rotation: P = A0 g1 + B0 g2 + sP r, M = B0 g1 + C0 g2 + sM r; every
                // component goes to shared memory as soon as it exists (few live registers)
                const float *pc = ringC + cslot * NPAR * CSLOT + sc;
                const f2 A0 = *(const f2 *)pc, B0 = *(const f2 *)(pc + CSLOT), C0 = *(const f2 *)(pc + 2 * CSLOT);
                const f2 rx = *(const f2 *)(pc + 3 * CSLOT), ry = *(const f2 *)(pc + 4 * CSLOT);
                const f2 rz = *(const f2 *)(pc + 5 * CSLOT);
                const f2 d1 = fma2(rz, G[0][2], fma2(ry, G[0][1], mul2(rx, G[0][0])));
                const f2 d2 = fma2(rz, G[1][2], fma2(ry, G[1][1], mul2(rx, G[1][0])));
                const f2 nB0 = mul2s(-1.f, B0);
                const f2 sP = fma2(nB0, d2, mul2(sub2(bb, A0), d1));
                const f2 sM = fma2(nB0, d1, mul2(mul2s(-1.f, C0), d2));
                if (need_xy && wr_x) {
                    *(f2 *)(fb + sfx) = fma2(sP, rx, fma2(B0, G[1][0], mul2(A0, G[0][0])));
                    *(f2 *)(fb + FX + sfx) = fma2(sM, rx, fma2(C0, G[1][0], mul2(B0, G[0][0])));
                }
                if (need_xy && wr_y) {
                    *(f2 *)(fb + 2 * FX + sfy) = fma2(sP, ry, fma2(B0, G[1][1], mul2(A0, G[0][1])));
                    *(f2 *)(fb + 2 * FX + FY + sfy) = fma2(sM, ry, fma2(C0, G[1][1], mul2(B0, G[0][1])));
                }
                if (tile) {
                    float *fz_ = fb + 2 * FX + 2 * FY + st;
                    *(f2 *)fz_ = fma2(sP, rz, fma2(B0, G[1][2], mul2(A0, G[0][2])));
                    *(f2 *)(fz_ + PSLOT) = fma2(sM, rz, fma2(C0, G[1][2], mul2(B0, G[0][2])));
                    *(float2 *)(fz_ + 2 * PSLOT) = q1[0];     // u1, u2 of the output plane zf-3
                    *(float2 *)(fz_ + 3 * PSLOT) = q2[0];
                }
            }
            // plane zf and its parameters are dead for this warp
            __syncwarp();
            mbar_arrive_if(&emptyU[s_c], lane == 0); mbar_arrive_if(&emptyC[cslot], lane == 0);
            if (++cslot == NC) { cslot = 0; cpar ^= 1; }
            __threadfence_block();
            named_bar_arrive(1 + (j & 1), NSYNC);
        }
        return;
    }

    // ===== D role: divergence, time update and fused extras of one pair =============================
    const int dt_ = tid - NF;
    const int cx2 = dt_ & (TXP - 1), cy = dt_ / TXP;
    const int sfx = cy * FXW + 2 * cx2 + 4, sfy = (cy + R) * TX + 2 * cx2, st = cy * TX + 2 * cx2;
    const int ox = x0 + 2 * cx2, oy = y0 + cy;
    const bool act = ox < g.n[0] && oy < g.n[1];
    const bool xfull = ox + 1 < g.n[0];
    long long gout = g.at(ox, oy, z0);
    f2 dxy = pk1(0.f);
    if (act) {
        const float2 d2_ = *(const float2 *)(a.dx + ox);
        dxy = add2(d2_, pk1(a.dy[oy]));
    }
    constexpr bool f_snap = (FUSE & TF_SNAP) != 0, f_grad = (FUSE & TF_GRAD) != 0;
    constexpr bool f_reg = f_snap || f_grad;
    [[maybe_unused]] const Fuse &fz = a.f;
    [[maybe_unused]] const bool f_xy = f_reg && act && oy >= fz.reg.lo[1] && oy < fz.reg.hi[1] &&
                                       ox + 1 >= fz.reg.lo[0] && ox < fz.reg.hi[0];
    [[maybe_unused]] const bool f_vec = f_xy && (fz.reg.lo[0] & 1) == 0 && ox >= fz.reg.lo[0] &&
                                        ox + 1 < fz.reg.hi[0];
    [[maybe_unused]] long long ri = f_reg ? fz.reg.at(ox, oy, z0) : 0;
    [[maybe_unused]] const long long rstride = f_reg ? (long long)fz.reg.pitch * fz.reg.ext[1] : 0;
    [[maybe_unused]] const float gc = f_grad ? fz.gcoef * a.irho_c : 0.f;

    float2 aP[NQ], aM[NQ];
#pragma unroll
    for (int k = 0; k < NQ; k++) aP[k] = aM[k] = make_float2(0.f, 0.f);
    int pslot = 0, ppar = 0;

    for (int j = FIRST; j < total; j++) {
        const int zf = z0 - (3 * R - 1) + j;
        const bool need_xy = zf >= z0 && zf < z1;
        const bool out = j >= WARM;
        const int z = z0 + j - WARM;
        const float *fb = fbuf + (j & 1) * FBUF;
        float dz_s = 0.f;
        [[maybe_unused]] bool fin = false, fvec = false;
        [[maybe_unused]] float2 us1, us2, G2;
        if (out) {
            dz_s = __ldg(a.dz + z);
            if (f_reg) {
                const bool zin = z >= fz.reg.lo[2] && z < fz.reg.hi[2];
                fin = f_xy && zin;
                fvec = f_vec && zin;
                if (f_grad && f_vec && (cx2 & 1) == 0 && z + 3 >= fz.reg.lo[2] && z + 3 < fz.reg.hi[2]) {
                    prefetch_l2(fz.usnap[0] + ri + 3 * rstride);
                    prefetch_l2(fz.usnap[1] + ri + 3 * rstride);
                    prefetch_l2(fz.grad + ri + 3 * rstride);
                }
                if (fvec && f_grad) {
                    us1 = __ldg((const float2 *)(fz.usnap[0] + ri));
                    us2 = __ldg((const float2 *)(fz.usnap[1] + ri));
                    G2 = *(const float2 *)(fz.grad + ri);
                }
            }
        }
        named_bar_sync(1 + (j & 1), NSYNC);
        if (need_xy) {
#pragma unroll
            for (int f = 0; f < 2; f++) {
                const float *px = fb + f * FX + sfx;
                const f2 lx = ttih_dx<false>(cf.wx, *(const float2 *)(px - 4), *(const float2 *)(px - 2),
                                             *(const float2 *)px, *(const float2 *)(px + 2), *(const float2 *)(px + 4));
                const float *py = fb + 2 * FX + f * FY + sfy;
                f2 ly;
#pragma unroll
                for (int k = 1; k <= R; k++) {
                    const f2 d = sub2(*(const f2 *)(py + (k - 1) * TX), *(const f2 *)(py - k * TX));
                    if (k == 1) ly = mul2s(cf.wy[1], d);
                    else ly = fma2s(cf.wy[k], d, ly);
                }
                float2 &acc = f ? aM[R - 1] : aP[R - 1];
                acc = add2(acc, add2(lx, ly));
            }
        }
        {
            const float *fz_ = fb + 2 * FX + 2 * FY + st;
            const f2 Pz = *(const f2 *)fz_, Mz = *(const f2 *)(fz_ + PSLOT);
#pragma unroll
            for (int k = 1; k <= R; k++) {
                aP[R - k] = fma2s(cf.wz[k], Pz, aP[R - k]);
                aM[R - k] = fma2s(cf.wz[k], Mz, aM[R - k]);
                if (k == R) {
                    aP[R - 1 + k] = mul2s(-cf.wz[k], Pz);
                    aM[R - 1 + k] = mul2s(-cf.wz[k], Mz);
                } else {
                    aP[R - 1 + k] = fma2s(-cf.wz[k], Pz, aP[R - 1 + k]);
                    aM[R - 1 + k] = fma2s(-cf.wz[k], Mz, aM[R - 1 + k]);
                }
            }
        }
        if (out) {
            const float *fz_ = fb + 2 * FX + 2 * FY + st;
            const f2 uc1 = *(const f2 *)(fz_ + 2 * PSLOT), uc2 = *(const f2 *)(fz_ + 3 * PSLOT);
            const f2 H1 = aP[0], H2 = aM[0];
            mbar_wait(&fullP[pslot], (uint32_t)ppar);
            const float *pt = ringP + pslot * 3 * PSLOT + st;
            const f2 up1 = *(const f2 *)pt, up2 = *(const f2 *)(pt + PSLOT), m2 = *(const f2 *)(pt + 2 * PSLOT);
            __syncwarp();
            mbar_arrive_if(&emptyP[pslot], lane == 0);
            if (++pslot == NP) { pslot = 0; ppar ^= 1; }
            const f2 sd = mul2s(cf.dt, add2(dxy, pk1(dz_s)));
            const f2 den = fma2s(a.irho_c, m2, sd);
            const f2 r = pk(__frcp_rn(den.x), __frcp_rn(den.y));
            const f2 nsd = mul2s(-1.f, sd);
            f2 dl[2];
#pragma unroll
            for (int f = 0; f < 2; f++) {
                const f2 uc = f ? uc2 : uc1, up = f ? up2 : up1, H = f ? H2 : H1;
                const f2 d1 = sub2(uc, up);
                const f2 t = fma2(nsd, d1, mul2s(cf.dt2, H));
                dl[f] = mul2(t, r);
                const f2 un = fma2(t, r, add2(uc, d1));
                if (act) {
                    float *o = (f ? a.up2 : a.up1) + gout;
                    if (xfull) *(f2 *)o = un;
                    else o[0] = un.x;
                }
            }
            if (f_reg && fin) {
                if (fvec) {
                    if (f_snap) {
                        __stcs((float2 *)(fz.snap[0] + ri), uc1);
                        __stcs((float2 *)(fz.snap[1] + ri), uc2);
                    }
                    if (f_grad) {
                        const f2 ic = fma2(us2, dl[1], mul2(us1, dl[0]));
                        *(f2 *)(fz.grad + ri) = fma2s(-gc, ic, G2);
                    }
                } else {
                    const float u1v[2] = {uc1.x, uc1.y}, u2v[2] = {uc2.x, uc2.y};
                    const float d1v[2] = {dl[0].x, dl[0].y}, d2v[2] = {dl[1].x, dl[1].y};
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                        if (ox + e < fz.reg.lo[0] || ox + e >= fz.reg.hi[0] || ox + e >= g.n[0]) continue;
                        if (f_snap) { fz.snap[0][ri + e] = u1v[e]; fz.snap[1][ri + e] = u2v[e]; }
                        if (f_grad) {
                            const float ic = __fmaf_rn(fz.usnap[1][ri + e], d2v[e], __fmul_rn(fz.usnap[0][ri + e], d1v[e]));
                            fz.grad[ri + e] = __fmaf_rn(-gc, ic, fz.grad[ri + e]);
                        }
                    }
                }
            }
            gout += g.sz;
            if (f_reg) ri += rstride;
        }
        if (j + 2 < total) named_bar_arrive(3 + (j & 1), NSYNC);
#pragma unroll
        for (int k = 0; k < NQ - 1; k++) { aP[k] = aP[k + 1]; aM[k] = aM[k + 1]; }
    }
}

// ---- host side ---------------------------------------------------------------------------------
bool tti3d_supported(const judi_model *M)
{
    return M->ndim == 3 && M->tti && !M->visco && (M->r == 4 || M->r == 2) && M->ctx->encode_tiled != nullptr &&
           M->ctx->opt_tti_fused && M->ttiABC[0].p != nullptr && (!M->irho_field || M->tti_mi.p != nullptr);
}

void tti3d_encode(const judi_model *M, const float *base, CUtensorMap *halo_map, CUtensorMap *tile_map)
{
    if (halo_map) tma_encode_box(M, base, TtiC<4>::UPX, M->r == 2 ? TtiC<2>::UPY : TtiC<4>::UPY, halo_map);
    if (tile_map) tma_encode_box(M, base, TtiC<4>::TX, TtiC<4>::TY, tile_map);
}

void tti3d_encode_coef(const judi_model *M, CUtensorMap *maps)
{
    const int cpy = M->r == 2 ? TtiC<2>::CPY : TtiC<4>::CPY;
    for (int p = 0; p < 3; p++) tma_encode_box(M, M->ttiABC[p].p, TtiC<4>::CPX, cpy, &maps[p]);
    for (int p = 0; p < 3; p++) tma_encode_box(M, M->r3[p].p, TtiC<4>::CPX, cpy, &maps[3 + p]);
    if (M->irho_field) tma_encode_box(M, M->irho.p, TtiC<4>::CPX, cpy, &maps[6]);
    else maps[6] = maps[0];
}

// z-chunks per tile column: whole waves of one CTA per SM, counting the 14 warm-up planes
static int tti3d_chunks(const judi_ctx *ctx, int ntiles, int nz, int warm)
{
    if (ctx->opt_zchunks > 0) return std::min(ctx->opt_zchunks, nz);
    int best = 1;
    double best_cost = 1e30;
    for (int c = 1; c <= 8 && c <= nz; c++) {
        const int zc = (nz + c - 1) / c;
        const long long ctas = (long long)ntiles * ((nz + zc - 1) / zc);
        const double waves = (double)((ctas + ctx->num_sms - 1) / ctx->num_sms);
        const double cost = waves * (zc + warm);
        if (cost < best_cost * 0.999) { best_cost = cost; best = c; }
    }
    return best;
}

// lin: launch on the linearised wavefield (Born source in); otherwise on the background field
template <int R, int FUSE, bool RHO = false>
static void tti3d_launch_r(const judi_model *M, const StepArgs &s, bool lin)
{
    typedef TtiC<R, RHO> C;
    judi_ctx *ctx = M->ctx;
    // register queues: 2 shifted every plane (default), 1 rotated at compile time, 3 / 4 only the
    // F / D role rotated
    // option tti_fused: 2 shifted queues, 1 rotated (needs ring length == queue length: R = 4 only),
    // 5 / 6 / 7: loops unrolled by two in both roles / in F only / in D only
    constexpr int ROTQ = C::NS == C::NQ ? C::NQ : 1;
    int var = ctx->opt_tti_fused;
    if (var == 1 && ROTQ == 1) var = 2;
    if (var != 1 && var != 5 && var != 6 && var != 7) var = 2;
    void (*kern)(const TtiMaps, const TtiArgs);
    if constexpr (RHO) {
        var = 2;                  // the density-field kernel exists with shifted queues only
        kern = tti3d_tma<R, FUSE, 1, 1, true>;
    } else {
        kern = var == 1 ? tti3d_tma<R, FUSE, ROTQ, ROTQ> : var == 5 ? tti3d_tma<R, FUSE, 2, 2>
               : var == 6 ? tti3d_tma<R, FUSE, 2, 1> : var == 7 ? tti3d_tma<R, FUSE, 1, 2> : tti3d_tma<R, FUSE, 1, 1>;
    }
    static bool configured[8] = {false, false, false, false, false, false, false, false};
    if (!configured[var]) {
        CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
        configured[var] = true;
    }
    // option tti_fused = 8: the half-width-thread kernel (R = 4; plain / save / imaging condition)
    constexpr bool HOK = !RHO && R == 4 && (FUSE == 0 || FUSE == TF_SNAP || FUSE == TF_GRAD);
    bool half = false;
    if constexpr (HOK) {
        if (ctx->opt_tti_fused == 8) {
            half = true;
            kern = tti3d_h<FUSE>;
            static bool hconf = false;
            if (!hconf) {
                CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
                hconf = true;
            }
        }
    }
    Wavefield *W = lin ? s.lin : s.scratch;
    TtiMaps tm;
    tm.U1 = W->tmap_halo[0][W->cur]; tm.U2 = W->tmap_halo[1][W->cur];
    tm.P1 = W->tmap_tile[0][1 - W->cur]; tm.P2 = W->tmap_tile[1][1 - W->cur];
    tm.M = W->tmap_m;
    for (int p = 0; p < TTI_NPAR; p++) tm.C[p] = W->tmap_par[p];
    TtiArgs a;
    memset(&a, 0, sizeof(a));
    a.g = M->g; a.cf = M->cf;
    a.up1 = lin ? s.ulp[0] : s.up[0]; a.up2 = lin ? s.ulp[1] : s.up[1];
    if (FUSE & TF_DDOUT) { a.dd_out[0] = s.scratch->dd[0].p; a.dd_out[1] = s.scratch->dd[1].p; }
    if (FUSE & TF_QSRC) {
        a.dd_in[0] = s.scratch->dd[0].p; a.dd_in[1] = s.scratch->dd[1].p;
        a.dm = M->dm_src();
    }
    a.dx = M->damp[0].p; a.dy = M->damp[1].p; a.dz = M->damp[2].p;
    a.irho_c = M->irho_c;
    a.irho = M->irho.p;
    if (!lin) a.f = s.f;
    const int ntx = (M->N[0] + C::TX - 1) / C::TX, nty = (M->N[1] + C::TY - 1) / C::TY;
    int nzc = tti3d_chunks(ctx, ntx * nty, M->N[2], C::WARM);
    a.zchunk = (M->N[2] + nzc - 1) / nzc;
    nzc = (M->N[2] + a.zchunk - 1) / a.zchunk;
    const dim3 grid(ntx, nty, nzc);
    kern<<<grid, half ? ttih::NTHREADS : C::NTHREADS, C::SMEM, ctx->stream>>>(tm, a);
    launch_count(ctx, "tti3d_tma");
}

template <int FUSE>
static void tti3d_launch(const judi_model *M, const StepArgs &s, bool lin = false)
{
    if (M->irho_field) {
        if (M->r == 2) tti3d_launch_r<2, FUSE, true>(M, s, lin);
        else tti3d_launch_r<4, FUSE, true>(M, s, lin);
    } else if (M->r == 2) tti3d_launch_r<2, FUSE>(M, s, lin);
    else tti3d_launch_r<4, FUSE>(M, s, lin);
}

// One fused TTI step; false if this combination has no single-pass instantiation.
bool tti3d_step(const judi_model *M, const StepArgs &s)
{
    if (!tti3d_supported(M) || s.qext || s.f.wrec || !s.scratch || !s.scratch->has_tmap) return false;
    if (s.f.grad && s.f.ic != JUDI_IC_AS) return false;
    const int mask = (s.f.snap[0] ? TF_SNAP : 0) | (s.f.grad ? TF_GRAD : 0) | (s.f.illum ? TF_ILLUM : 0);
    if (s.born) {
        // Born step: the background field writes its delta, the linearised field reads it back as
        // its source -- two single-pass launches instead of four two-pass ones
        if (s.born_ic != JUDI_IC_AS || s.f.grad || !s.lin || !s.lin->has_tmap || !s.scratch->dd[0].p ||
            !s.scratch->dd[1].p || !M->dm.p)
            return false;
        switch (mask) {
        case 0: tti3d_launch<TF_DDOUT>(M, s); break;
        case TF_SNAP: tti3d_launch<TF_DDOUT | TF_SNAP>(M, s); break;
        case TF_ILLUM: tti3d_launch<TF_DDOUT | TF_ILLUM>(M, s); break;
        case TF_SNAP | TF_ILLUM: tti3d_launch<TF_DDOUT | TF_SNAP | TF_ILLUM>(M, s); break;
        default: return false;
        }
        tti3d_launch<TF_QSRC>(M, s, true);
        return true;
    }
    switch (mask) {
    case 0: tti3d_launch<0>(M, s); return true;
    case TF_SNAP: tti3d_launch<TF_SNAP>(M, s); return true;
    case TF_GRAD: tti3d_launch<TF_GRAD>(M, s); return true;
    case TF_ILLUM: tti3d_launch<TF_ILLUM>(M, s); return true;
    case TF_SNAP | TF_ILLUM: tti3d_launch<TF_SNAP | TF_ILLUM>(M, s); return true;
    case TF_GRAD | TF_ILLUM: tti3d_launch<TF_GRAD | TF_ILLUM>(M, s); return true;
    default: return false;
    }
}
