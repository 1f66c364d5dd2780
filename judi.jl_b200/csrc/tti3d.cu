// tti3d.cu -- single-pass 3-D TTI time step for sm_100a (space order 8, constant density).
// Replaces the Devito loop nest of tti_kernel (src/pysource/kernels.py:144-182 with the rotated
// self-adjoint operator of FD_utils.py:24-105) -- and the two-pass form of flux3d.cu, whose six
// flux fields cost 48 of its 119 B per grid point in HBM round trips.
//
//   H0 = sum_d D-_d P_d,  H1 = sum_d D-_d M_d,   (P, M) = R^T [[A,B],[B,C]] R (G u1, G u2)
//
// needs the flux one stencil radius around every output point, and every flux node needs the full
// staggered gradient of both fields.  One CTA owns a 64 x 8 (x,y) tile and marches along z; per
// z plane it runs two phases separated by ONE named barrier:
//
//  F  272 threads, one per group of 4 x-points of the plus-shaped region (tile + one group
//     left/right + 4 rows below/above), keep the z-lines of u1, u2 of their own column in register
//     queues, read the x/y-lines from the shared plane (TMA ring, haloed 80 x 24 boxes), evaluate
//     the closed-form rotation (flux3d.cu) and publish P_x, M_x (x-extended) and P_y, M_y
//     (y-extended) in a double-buffered shared plane; the 128 tile threads also append P_z, M_z
//     to a shared ring and hand u2 of the output plane to the thread that updates it;
//  D  256 threads, one per (tile group, field): D-_x + D-_y from the shared flux plane, parked
//     for R-1 planes in a private ring until the z part is complete, D-_z from the P_z / M_z
//     ring, time update of that field in place (+ fused snapshot save / imaging condition /
//     illumination; field 1 adds its share to the gradient one plane later, behind the next
//     barrier, so the two read-modify-writes of an address are ordered).
//
// HBM traffic: 44 B per grid point (u1, u2, their previous level, m, eps, delta, r: read; two
// writes) + the plane halos that miss L2; the price is evaluating the flux on 1088 instead of
// 512 nodes per tile plane.
//
// STATUS: EXPERIMENTAL, OFF BY DEFAULT (option "tti_fused").  It is correct (the 3-D TTI parity,
// dot-product and full-size tests pass with it) but slower than the two-pass form: 1.29 ms vs
// 1.12 ms per step on the C5 grid.  ncu (gpurun_out/r01_tti3d*.ncu-rep): 540-660 M warp
// instructions per step against 495 M for the two passes together, 25-31% issue-active with 9
// flux warps per SM, 168 registers (the cap at 10-12 warps) with the z-queues alone taking 64;
// an unbalanced variant (tile threads do the whole D phase) ran 1.27 ms barrier-stalled, and
// register-prefetching the parameters spilled (1.85 ms).  The 2.1x redundant flux evaluation that
// a 64 x 8 tile costs is not paid back by the saved HBM traffic because neither form is
// HBM-bound.  What would change that: flux halos exchanged between the CTAs of a thread-block
// cluster through distributed shared memory instead of recomputed (redundancy ~1.2x).
#include "tma_util.cuh"

namespace tti {
constexpr int R = 4, XE = 4;             // stencil radius, flux halo in x (one float4 group)
constexpr int TXT = 16, TY = 8;          // tile: 16 groups of 4 x-points, 8 rows
constexpr int TX = 4 * TXT;
constexpr int UPX = TX + 4 * XE;         // haloed u plane: 80 x 24
constexpr int UPY = TY + 4 * R;
constexpr int USLOT = UPX * UPY;         // floats per field plane
constexpr int NS = 8, NP = 3;            // ring depths (u planes, tiles)
constexpr int PSLOT = TX * TY;
constexpr int NQ = 2 * R;                // z-queue: planes zf-R+1 .. zf+R
constexpr int NI = TXT * TY;             // 128 tile groups
constexpr int NY = TXT * 2 * R;          // 128 y-halo groups
constexpr int NX = 2 * TY;               // 16 x-halo groups
constexpr int ND = 2 * NI;               // D-phase threads: (tile group, field)
constexpr int NFLUX = NI + NY + 32;      // 9 warps (the last one half empty)
constexpr int NCW = NFLUX / 32;
constexpr int NTHREADS = NFLUX + 32;     // + the producer warp
constexpr int FXW = (TXT + 2) * 4;       // row pitch of the x-extended flux plane (72)
constexpr int FX = TY * FXW;             // floats per component
constexpr int FY = (TY + 2 * R) * TX;    // y-extended flux plane
constexpr int FBUF = 2 * FX + 2 * FY;    // P_x, M_x, P_y, M_y of one plane
constexpr int NZ = NQ + 1;               // ring of P_z, M_z (one spare: written by another thread)
constexpr int FZ = NZ * 2 * PSLOT;
constexpr int HQ = R * 2 * PSLOT;        // parked D-x + D-y partial sums (thread private)
constexpr int UCX = 2 * PSLOT;           // u2 of the output plane, double buffered
constexpr size_t SMEM = (size_t)(NS * 2 * USLOT + NP * 3 * PSLOT + 2 * FBUF + FZ + HQ + UCX) * 4 +
                        (2 * NS + 2 * NP) * 8 + 128;
}  // namespace tti

struct TtiMaps {
    CUtensorMap U1, U2;        // haloed planes of u1(t), u2(t)
    CUtensorMap P1, P2, M;     // tiles of u1(t-1), u2(t-1), m
};

struct TtiArgs {
    Grid g;
    Coefs cf;
    float *up1, *up2;
    const float *eps, *delta, *r3x, *r3y, *r3z;
    const float *dx, *dy, *dz;
    float irho_c;
    int zchunk;
    Fuse f;
};

enum { TF_SNAP = 1, TF_GRAD = 2, TF_ILLUM = 4 };

template <int FUSE>
__global__ void __launch_bounds__(tti::NTHREADS, 1)
    tti3d_tma(const __grid_constant__ TtiMaps tm, const TtiArgs a)
{
    using namespace tti;
    extern __shared__ unsigned char smem_raw[];
    float *ringU = (float *)(smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u));
    float *ringP = ringU + NS * 2 * USLOT;     // [NP][3][PSLOT]: u1-, u2-, m
    float *fbuf = ringP + NP * 3 * PSLOT;      // [2][FBUF]
    float *fzr = fbuf + 2 * FBUF;              // [NZ][2][PSLOT]
    float *hq = fzr + FZ;                      // [R][2][PSLOT]
    float *ucx = hq + HQ;                      // [2][PSLOT]
    uint64_t *fullU = (uint64_t *)(ucx + UCX);
    uint64_t *emptyU = fullU + NS;
    uint64_t *fullP = emptyU + NS;
    uint64_t *emptyP = fullP + NP;

    const Grid &g = a.g;
    const Coefs &cf = a.cf;
    const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
    const int z0 = blockIdx.z * a.zchunk;
    const int z1 = min(z0 + a.zchunk, g.n[2]);
    const int nzc = z1 - z0;
    // iteration j brings in u plane z0-2R+1+j; flux plane zf = z0-3R+1+j (from j = 2R-1);
    // output plane zo = zf-R+1 = z0-4R+2+j (from j = 4R-2)
    const int total = nzc + 4 * R - 2;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; s++) { mbar_init(&fullU[s], 1); mbar_init(&emptyU[s], NCW); }
        for (int s = 0; s < NP; s++) { mbar_init(&fullP[s], 1); mbar_init(&emptyP[s], ND / 32); }
        fence_barrier_init();
    }
    __syncthreads();

    if (threadIdx.x >= NFLUX) {
        // ===== producer warp: one lane issues both streams ========================================
        if ((threadIdx.x & 31) == 0) {
            const int bx = g.hx + x0 - 2 * XE, by = g.hy + y0 - 2 * R, bz = g.hz + z0 - 2 * R + 1;
            const int tx0 = g.hx + x0, ty0 = g.hy + y0, tz0 = g.hz + z0;
            for (int j = 0; j < total; j++) {
                const int s = j % NS;
                if (j >= NS) mbar_wait(&emptyU[s], (uint32_t)(((j / NS) - 1) & 1));
                mbar_arrive_expect_tx(&fullU[s], 2 * USLOT * 4);
                tma_load_3d(ringU + s * 2 * USLOT, &tm.U1, bx, by, bz + j, &fullU[s]);
                tma_load_3d(ringU + s * 2 * USLOT + USLOT, &tm.U2, bx, by, bz + j, &fullU[s]);
                // tiles of output plane z0+i are consumed at iteration i + 4R-2: issue one early (the
                // slot was released three iterations ago, so this wait does not stall the u stream)
                const int i = j - (4 * R - 2) + 1;
                if (i >= 0 && i < nzc) {
                    const int sp = i % NP;
                    if (i >= NP) mbar_wait(&emptyP[sp], (uint32_t)(((i / NP) - 1) & 1));
                    mbar_arrive_expect_tx(&fullP[sp], 3 * PSLOT * 4);
                    tma_load_3d(ringP + sp * 3 * PSLOT, &tm.P1, tx0, ty0, tz0 + i, &fullP[sp]);
                    tma_load_3d(ringP + sp * 3 * PSLOT + PSLOT, &tm.P2, tx0, ty0, tz0 + i, &fullP[sp]);
                    tma_load_3d(ringP + sp * 3 * PSLOT + 2 * PSLOT, &tm.M, tx0, ty0, tz0 + i, &fullP[sp]);
                }
            }
        }
        return;
    }

    // ===== F role: flux node group (cx in -1..16 groups of 4 x-points, cy in -4..11) =============
    const int tid = threadIdx.x, lane = tid & 31;
    int cx, cy;
    bool tile = false, wr_x = false, wr_y = false, idle = false;
    if (tid < NI) { cx = tid % TXT; cy = tid / TXT; tile = wr_x = wr_y = true; }
    else if (tid < NI + NY) {
        const int h = (tid - NI) / TXT;
        cx = (tid - NI) % TXT;
        cy = h < R ? h - R : h - R + TY;      // rows -4..-1 and 8..11
        wr_y = true;
    } else if (tid < NI + NY + NX) {
        const int h = tid - NI - NY;
        cx = (h & 1) ? TXT : -1;
        cy = h >> 1;
        wr_x = true;
    } else { cx = 0; cy = 0; idle = true; }
    const int x = x0 + 4 * cx, y = y0 + cy;
    // parameters are defined on the grid and its halo; nodes beyond it carry no flux anyone reads
    const bool inb = !idle && x >= -g.hx && x + 3 < g.px - g.hx && y >= -g.hy && y < g.py - g.hy;
    const int su = (cy + 2 * R) * UPX + (4 * cx + 2 * XE);   // node group inside a u plane
    const int sfx = cy * FXW + 4 * (cx + 1);                 // ... inside the x-extended flux plane
    const int sfy = (cy + R) * TX + 4 * cx;                  // ... inside the y-extended flux plane
    const int st = cy * TX + 4 * cx;                         // ... inside a tile (tile threads)
    long long gnode = g.at(x, y, z0 - R);                    // flux node of the first flux plane

    // ===== D role: (tile group, field) ============================================================
    const bool dth = tid < ND;
    const int fld = tid >= NI ? 1 : 0;
    const int dg = tid & (NI - 1);
    const int dgx = dg % TXT, dgy = dg / TXT;
    const int ox = x0 + 4 * dgx, oy = y0 + dgy;
    const bool act = dth && ox < g.n[0] && oy < g.n[1];
    const bool xfull = ox + 3 < g.n[0];
    const int dfx = dgy * FXW + 4 * (dgx + 1), dfy = (dgy + R) * TX + 4 * dgx, dt_ = dgy * TX + 4 * dgx;
    long long gout = g.at(ox, oy, z0);
    float *const outp = fld ? a.up2 : a.up1;
    f2 dxy0 = pk1(0.f), dxy1 = pk1(0.f);
    if (act) {
        const float4 d4 = *(const float4 *)(a.dx + ox);
        dxy0 = add2(lo2(d4), pk1(a.dy[oy]));
        dxy1 = add2(hi2(d4), pk1(a.dy[oy]));
    }
    constexpr bool f_snap = (FUSE & TF_SNAP) != 0, f_grad = (FUSE & TF_GRAD) != 0;
    constexpr bool f_il = (FUSE & TF_ILLUM) != 0, f_reg = FUSE != 0;
    [[maybe_unused]] const Fuse &fz = a.f;
    [[maybe_unused]] const bool f_xy = f_reg && act && oy >= fz.reg.lo[1] && oy < fz.reg.hi[1] &&
                                       ox + 3 >= fz.reg.lo[0] && ox < fz.reg.hi[0];
    [[maybe_unused]] const bool f_vec = f_xy && (fz.reg.lo[0] & 3) == 0 && ox >= fz.reg.lo[0] &&
                                        ox + 3 < fz.reg.hi[0];
    [[maybe_unused]] long long ri = f_reg ? fz.reg.at(ox, oy, z0) : 0;
    [[maybe_unused]] const long long rstride = f_reg ? (long long)fz.reg.pitch * fz.reg.ext[1] : 0;
    [[maybe_unused]] const float gc = f_grad ? fz.gcoef * a.irho_c : 0.f;
    // field 1 applies its read-modify-writes (gradient, illumination) one plane late
    [[maybe_unused]] float4 pend_g = make_float4(0.f, 0.f, 0.f, 0.f), pend_i = pend_g;
    [[maybe_unused]] long long pend_ri = -1;
    [[maybe_unused]] bool pend_vec = false;
    auto flush_pending = [&]() {
        if (!(f_grad || f_il) || fld == 0 || pend_ri < 0) return;
        const float pg[4] = {pend_g.x, pend_g.y, pend_g.z, pend_g.w};
        const float pi[4] = {pend_i.x, pend_i.y, pend_i.z, pend_i.w};
        if (pend_vec) {
            if (f_grad) {
                float4 G = *(float4 *)(fz.grad + pend_ri);
                G.x = __fadd_rn(G.x, pg[0]); G.y = __fadd_rn(G.y, pg[1]);
                G.z = __fadd_rn(G.z, pg[2]); G.w = __fadd_rn(G.w, pg[3]);
                *(float4 *)(fz.grad + pend_ri) = G;
            }
            if (f_il) {
                float4 I = *(float4 *)(fz.illum + pend_ri);
                I.x = __fadd_rn(I.x, pi[0]); I.y = __fadd_rn(I.y, pi[1]);
                I.z = __fadd_rn(I.z, pi[2]); I.w = __fadd_rn(I.w, pi[3]);
                *(float4 *)(fz.illum + pend_ri) = I;
            }
        } else {
#pragma unroll
            for (int e = 0; e < 4; e++) {
                if (ox + e < fz.reg.lo[0] || ox + e >= fz.reg.hi[0]) continue;
                if (f_grad) fz.grad[pend_ri + e] = __fadd_rn(fz.grad[pend_ri + e], pg[e]);
                if (f_il) fz.illum[pend_ri + e] = __fadd_rn(fz.illum[pend_ri + e], pi[e]);
            }
        }
        pend_ri = -1;
    };

    float4 q1[NQ], q2[NQ];
#pragma unroll
    for (int k = 0; k < NQ; k++) q1[k] = q2[k] = make_float4(0.f, 0.f, 0.f, 0.f);

    for (int j = 0; j < total; j++) {
        // ---- incoming u plane -> z-queues ---------------------------------------------------------
#pragma unroll
        for (int k = 0; k < NQ - 1; k++) { q1[k] = q1[k + 1]; q2[k] = q2[k + 1]; }
        const int s_in = j % NS;
        const bool do_flux = j >= 2 * R - 1;
        // parameters of this flux node: loaded now, used after the gradient; the lines of the
        // next two planes are pulled into L2 so that these loads cost an L2 hit
        float4 ep4, dl4, rx4, ry4, rz4;
        ep4 = dl4 = rx4 = ry4 = rz4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (do_flux && inb) {
            ep4 = __ldg((const float4 *)(a.eps + gnode));
            dl4 = __ldg((const float4 *)(a.delta + gnode));
            rx4 = __ldg((const float4 *)(a.r3x + gnode));
            ry4 = __ldg((const float4 *)(a.r3y + gnode));
            rz4 = __ldg((const float4 *)(a.r3z + gnode));
            if ((lane & 7) == 0 && j + 2 < total) {     // one prefetch per 128-byte line
                prefetch_l2(a.eps + gnode + 2 * g.sz);
                prefetch_l2(a.delta + gnode + 2 * g.sz);
                prefetch_l2(a.r3x + gnode + 2 * g.sz);
                prefetch_l2(a.r3y + gnode + 2 * g.sz);
                prefetch_l2(a.r3z + gnode + 2 * g.sz);
            }
        }
        mbar_wait(&fullU[s_in], (uint32_t)((j / NS) & 1));
        if (!idle) {
            q1[NQ - 1] = *(const float4 *)(ringU + s_in * 2 * USLOT + su);
            q2[NQ - 1] = *(const float4 *)(ringU + s_in * 2 * USLOT + USLOT + su);
        }
        if (!do_flux) {
            // planes that are never a flux plane of this chunk: release them in step
            if (j >= R) {
                __syncwarp();
                if (lane == 0) mbar_arrive(&emptyU[(j + NS - R) % NS]);
            }
            continue;
        }
        // ======== F phase: flux at plane zf = (incoming plane) - R ================================
        const int sc = (j + NS - R) % NS;
        const int jf = j - (2 * R - 1);          // flux plane counter
        float *fb = fbuf + (j & 1) * FBUF;
        {
            f2 G[2][3][2];   // [field][x,y,z][pair]
#pragma unroll
            for (int f = 0; f < 2; f++) {
                const float *pl = ringU + sc * 2 * USLOT + f * USLOT + su;
                const float4 *q = f ? q2 : q1;
                const float4 c = q[R - 1];
                float w[12];
                {
                    const float4 l = *(const float4 *)(pl - 4), r = *(const float4 *)(pl + 4);
                    w[0] = l.x; w[1] = l.y; w[2] = l.z; w[3] = l.w;
                    w[4] = c.x; w[5] = c.y; w[6] = c.z; w[7] = c.w;
                    w[8] = r.x; w[9] = r.y; w[10] = r.z; w[11] = r.w;
                }
                float o[4];
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    float acc = 0.f;
#pragma unroll
                    for (int k = 1; k <= R; k++)
                        acc = __fmaf_rn(cf.wx[k], __fsub_rn(w[XE + e + k], w[XE + e - k + 1]), acc);
                    o[e] = acc;
                }
                G[f][0][0] = pk(o[0], o[1]);
                G[f][0][1] = pk(o[2], o[3]);
                f2 gy0 = pk1(0.f), gy1 = pk1(0.f), gz0 = pk1(0.f), gz1 = pk1(0.f);
#pragma unroll
                for (int k = 1; k <= R; k++) {
                    const float4 p = *(const float4 *)(pl + k * UPX);
                    const float4 m = (k == 1) ? c : *(const float4 *)(pl - (k - 1) * UPX);
                    gy0 = fma2s(cf.wy[k], sub2(lo2(p), lo2(m)), gy0);
                    gy1 = fma2s(cf.wy[k], sub2(hi2(p), hi2(m)), gy1);
                    const float4 zp = q[R - 1 + k], zm = q[R - k];
                    gz0 = fma2s(cf.wz[k], sub2(lo2(zp), lo2(zm)), gz0);
                    gz1 = fma2s(cf.wz[k], sub2(hi2(zp), hi2(zm)), gz1);
                }
                G[f][1][0] = gy0; G[f][1][1] = gy1;
                G[f][2][0] = gz0; G[f][2][1] = gz1;
            }
            // the plane zf is dead for this warp's x/y lines
            __syncwarp();
            if (lane == 0) mbar_arrive(&emptyU[sc]);
            // closed-form rotation (flux3d.cu): P = A0 g1 + B0 g2 + sP r, M = B0 g1 + C0 g2 + sM r
            f2 P[3][2], Mf[3][2];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const f2 ep = h ? hi2(ep4) : lo2(ep4), dl = h ? hi2(dl4) : lo2(dl4);
                const f2 rx = h ? hi2(rx4) : lo2(rx4), ry = h ? hi2(ry4) : lo2(ry4);
                const f2 rz = h ? hi2(rz4) : lo2(rz4);
                const f2 bb = pk1(a.irho_c);
                const f2 emd = sub2(ep, dl);
                const f2 A0 = mul2(bb, dl);
                const f2 pr = mul2(emd, dl);
                const f2 B0 = mul2(bb, pk(sqrtf(pr.x), sqrtf(pr.y)));
                const f2 C0 = mul2(bb, emd);
                const f2 d1 = fma2(rz, G[0][2][h], fma2(ry, G[0][1][h], mul2(rx, G[0][0][h])));
                const f2 d2 = fma2(rz, G[1][2][h], fma2(ry, G[1][1][h], mul2(rx, G[1][0][h])));
                const f2 nB0 = mul2s(-1.f, B0);
                const f2 sP = fma2(nB0, d2, mul2(sub2(bb, A0), d1));
                const f2 sM = fma2(nB0, d1, mul2(mul2s(-1.f, C0), d2));
                P[0][h] = fma2(sP, rx, fma2(B0, G[1][0][h], mul2(A0, G[0][0][h])));
                P[1][h] = fma2(sP, ry, fma2(B0, G[1][1][h], mul2(A0, G[0][1][h])));
                P[2][h] = fma2(sP, rz, fma2(B0, G[1][2][h], mul2(A0, G[0][2][h])));
                Mf[0][h] = fma2(sM, rx, fma2(C0, G[1][0][h], mul2(B0, G[0][0][h])));
                Mf[1][h] = fma2(sM, ry, fma2(C0, G[1][1][h], mul2(B0, G[0][1][h])));
                Mf[2][h] = fma2(sM, rz, fma2(C0, G[1][2][h], mul2(B0, G[0][2][h])));
            }
            if (wr_x) {
                *(float4 *)(fb + sfx) = cat4(P[0][0], P[0][1]);
                *(float4 *)(fb + FX + sfx) = cat4(Mf[0][0], Mf[0][1]);
            }
            if (wr_y) {
                *(float4 *)(fb + 2 * FX + sfy) = cat4(P[1][0], P[1][1]);
                *(float4 *)(fb + 2 * FX + FY + sfy) = cat4(Mf[1][0], Mf[1][1]);
            }
            if (tile) {
                float *zr = fzr + (jf % NZ) * 2 * PSLOT + st;
                *(float4 *)zr = cat4(P[2][0], P[2][1]);
                *(float4 *)(zr + PSLOT) = cat4(Mf[2][0], Mf[2][1]);
                // u2 of the output plane of this iteration (the oldest queue entry) for the
                // thread that updates field 1 of this group
                *(float4 *)(ucx + (j & 1) * PSLOT + st) = q2[0];
            }
        }
        gnode += g.sz;
        named_bar_sync(1, NFLUX);
        // ======== D phase: (tile group, field) ====================================================
        if (dth) {
            flush_pending();
            // D-x + D-y of this flux plane, parked until its z part is complete
            {
                const float *px = fb + fld * FX + dfx;
                float w[12];
                {
                    const float4 l = *(const float4 *)(px - 4), c = *(const float4 *)px;
                    const float4 r = *(const float4 *)(px + 4);
                    w[0] = l.x; w[1] = l.y; w[2] = l.z; w[3] = l.w;
                    w[4] = c.x; w[5] = c.y; w[6] = c.z; w[7] = c.w;
                    w[8] = r.x; w[9] = r.y; w[10] = r.z; w[11] = r.w;
                }
                float o[4];
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    float acc = 0.f;
#pragma unroll
                    for (int k = 1; k <= R; k++)
                        acc = __fmaf_rn(cf.wx[k], __fsub_rn(w[XE + e + k - 1], w[XE + e - k]), acc);
                    o[e] = acc;
                }
                const float *py = fb + 2 * FX + fld * FY + dfy;
                f2 ly0 = pk1(0.f), ly1 = pk1(0.f);
#pragma unroll
                for (int k = 1; k <= R; k++) {
                    const float4 p = *(const float4 *)(py + (k - 1) * TX);
                    const float4 m = *(const float4 *)(py - k * TX);
                    ly0 = fma2s(cf.wy[k], sub2(lo2(p), lo2(m)), ly0);
                    ly1 = fma2s(cf.wy[k], sub2(hi2(p), hi2(m)), ly1);
                }
                *(float4 *)(hq + (jf % R) * 2 * PSLOT + fld * PSLOT + dt_) =
                    cat4(add2(pk(o[0], o[1]), ly0), add2(pk(o[2], o[3]), ly1));
            }
            if (j >= 4 * R - 2) {
                // ---- output plane zo = zf - R + 1 -------------------------------------------------
                const int i = j - (4 * R - 2);
                const int z = z0 + i;
                const int jo = jf - (R - 1);         // flux-plane counter of plane zo
                const int sp = i % NP;
                [[maybe_unused]] bool fin = false, fvec = false;
                [[maybe_unused]] float4 us4;
                if (f_reg) {
                    const bool zin = z >= fz.reg.lo[2] && z < fz.reg.hi[2];
                    fin = f_xy && zin;
                    fvec = f_vec && zin;
                    if (f_grad && f_vec && z + 3 >= fz.reg.lo[2] && z + 3 < fz.reg.hi[2]) {
                        prefetch_l2(fz.usnap[fld] + ri + 3 * rstride);
                        if (fld == 0) prefetch_l2(fz.grad + ri + 3 * rstride);
                    }
                    if (f_grad && fvec) us4 = __ldg((const float4 *)(fz.usnap[fld] + ri));
                }
                f2 L0, L1;
                {
                    const float4 h4 = *(const float4 *)(hq + (jo % R) * 2 * PSLOT + fld * PSLOT + dt_);
                    f2 lz0 = pk1(0.f), lz1 = pk1(0.f);
#pragma unroll
                    for (int k = 1; k <= R; k++) {
                        // planes zo+k-1 and zo-k: flux counters jo+k-1, jo-k (all >= 0 here)
                        const float4 p = *(const float4 *)(fzr + ((jo + k - 1) % NZ) * 2 * PSLOT + fld * PSLOT + dt_);
                        const float4 m = *(const float4 *)(fzr + ((jo - k + NZ) % NZ) * 2 * PSLOT + fld * PSLOT + dt_);
                        lz0 = fma2s(cf.wz[k], sub2(lo2(p), lo2(m)), lz0);
                        lz1 = fma2s(cf.wz[k], sub2(hi2(p), hi2(m)), lz1);
                    }
                    L0 = add2(lo2(h4), lz0);
                    L1 = add2(hi2(h4), lz1);
                }
                // plane zo is the oldest queue entry of the tile thread of this group
                const float4 uc = fld ? *(const float4 *)(ucx + (j & 1) * PSLOT + dt_) : q1[0];
                mbar_wait(&fullP[sp], (uint32_t)((i / NP) & 1));
                const float *pt = ringP + sp * 3 * PSLOT + dt_;
                const float4 up = *(const float4 *)(pt + fld * PSLOT);
                const float4 m4 = *(const float4 *)(pt + 2 * PSLOT);
                __syncwarp();
                if (lane == 0) mbar_arrive(&emptyP[sp]);
                // delta = (dt^2 H - sd (u - u-)) / (m irho + sd),  u+ = (u + (u - u-)) + delta
                const f2 dzv = pk1(a.dz[z]);
                const f2 sd0 = mul2s(cf.dt, add2(dxy0, dzv)), sd1 = mul2s(cf.dt, add2(dxy1, dzv));
                const f2 den0 = fma2s(a.irho_c, lo2(m4), sd0), den1 = fma2s(a.irho_c, hi2(m4), sd1);
                const f2 r0 = pk(__frcp_rn(den0.x), __frcp_rn(den0.y));
                const f2 r1 = pk(__frcp_rn(den1.x), __frcp_rn(den1.y));
                const f2 d10 = sub2(lo2(uc), lo2(up)), d11 = sub2(hi2(uc), hi2(up));
                const f2 dl0 = mul2(fma2(mul2s(-1.f, sd0), d10, mul2s(cf.dt2, L0)), r0);
                const f2 dl1 = mul2(fma2(mul2s(-1.f, sd1), d11, mul2s(cf.dt2, L1)), r1);
                const f2 un0 = add2(add2(lo2(uc), d10), dl0), un1 = add2(add2(hi2(uc), d11), dl1);
                if (act) {
                    float *o = outp + gout;
                    if (xfull) *(float4 *)o = cat4(un0, un1);
                    else {
                        const float v[4] = {un0.x, un0.y, un1.x, un1.y};
#pragma unroll
                        for (int e = 0; e < 4; e++)
                            if (ox + e < g.n[0]) o[e] = v[e];
                    }
                }
                if (f_reg && fin) {
                    // contributions of this field: gradient -gc us delta, illumination dt u^2
                    [[maybe_unused]] float4 cg, ci;
                    if (f_grad) {
                        float4 us = us4;
                        if (!fvec) {
                            float t[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                            for (int e = 0; e < 4; e++)
                                if (ox + e >= fz.reg.lo[0] && ox + e < fz.reg.hi[0]) t[e] = fz.usnap[fld][ri + e];
                            us = make_float4(t[0], t[1], t[2], t[3]);
                        }
                        cg = cat4(mul2(mul2s(-gc, lo2(us)), dl0), mul2(mul2s(-gc, hi2(us)), dl1));
                    }
                    if (f_il) ci = cat4(mul2(mul2s(cf.dt, lo2(uc)), lo2(uc)), mul2(mul2s(cf.dt, hi2(uc)), hi2(uc)));
                    if (f_snap) {
                        if (fvec) __stcs((float4 *)(fz.snap[fld] + ri), uc);
                        else {
                            const float v[4] = {uc.x, uc.y, uc.z, uc.w};
#pragma unroll
                            for (int e = 0; e < 4; e++)
                                if (ox + e >= fz.reg.lo[0] && ox + e < fz.reg.hi[0]) fz.snap[fld][ri + e] = v[e];
                        }
                    }
                    if (f_grad || f_il) {
                        // field 0 updates now, field 1 after the next barrier (or at the end)
                        if (f_grad) pend_g = cg;
                        if (f_il) pend_i = ci;
                        pend_ri = ri;
                        pend_vec = fvec;
                        if (fld == 0) {
                            const int keep = fld;   // flush_pending() skips field 0: do it inline
                            (void)keep;
                            const float pg[4] = {pend_g.x, pend_g.y, pend_g.z, pend_g.w};
                            const float pi[4] = {pend_i.x, pend_i.y, pend_i.z, pend_i.w};
                            if (fvec) {
                                if (f_grad) {
                                    float4 G = *(float4 *)(fz.grad + ri);
                                    G.x = __fadd_rn(G.x, pg[0]); G.y = __fadd_rn(G.y, pg[1]);
                                    G.z = __fadd_rn(G.z, pg[2]); G.w = __fadd_rn(G.w, pg[3]);
                                    *(float4 *)(fz.grad + ri) = G;
                                }
                                if (f_il) {
                                    float4 I = *(float4 *)(fz.illum + ri);
                                    I.x = __fadd_rn(I.x, pi[0]); I.y = __fadd_rn(I.y, pi[1]);
                                    I.z = __fadd_rn(I.z, pi[2]); I.w = __fadd_rn(I.w, pi[3]);
                                    *(float4 *)(fz.illum + ri) = I;
                                }
                            } else {
#pragma unroll
                                for (int e = 0; e < 4; e++) {
                                    if (ox + e < fz.reg.lo[0] || ox + e >= fz.reg.hi[0]) continue;
                                    if (f_grad) fz.grad[ri + e] = __fadd_rn(fz.grad[ri + e], pg[e]);
                                    if (f_il) fz.illum[ri + e] = __fadd_rn(fz.illum[ri + e], pi[e]);
                                }
                            }
                            pend_ri = -1;
                        }
                    }
                }
                gout += g.sz;
                if (f_reg) ri += rstride;
            }
        }
    }
    if (f_grad || f_il) {
        // the last plane of field 1: order it behind field 0's update of the same addresses
        named_bar_sync(1, NFLUX);
        if (dth) flush_pending();
    }
}

// ---- host side ---------------------------------------------------------------------------------
bool tti3d_supported(const judi_model *M)
{
    return M->ndim == 3 && M->tti && !M->irho_field && M->r == tti::R && M->ctx->encode_tiled != nullptr &&
           M->ctx->opt_tti_fused;
}

void tti3d_encode(const judi_model *M, const float *base, CUtensorMap *halo_map, CUtensorMap *tile_map)
{
    if (halo_map) tma_encode_box(M, base, tti::UPX, tti::UPY, halo_map);
    if (tile_map) tma_encode_box(M, base, tti::TX, tti::TY, tile_map);
}

template <int FUSE>
static void tti3d_launch(const judi_model *M, const StepArgs &s)
{
    judi_ctx *ctx = M->ctx;
    auto kern = tti3d_tma<FUSE>;
    static bool configured = false;
    if (!configured) {
        CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tti::SMEM));
        configured = true;
    }
    Wavefield *W = s.scratch;
    TtiMaps tm;
    tm.U1 = W->tmap_halo[0][W->cur]; tm.U2 = W->tmap_halo[1][W->cur];
    tm.P1 = W->tmap_tile[0][1 - W->cur]; tm.P2 = W->tmap_tile[1][1 - W->cur];
    tm.M = W->tmap_m;
    TtiArgs a;
    memset(&a, 0, sizeof(a));
    a.g = M->g; a.cf = M->cf;
    a.up1 = s.up[0]; a.up2 = s.up[1];
    a.eps = M->eps.p; a.delta = M->delta.p;
    a.r3x = M->r3[0].p; a.r3y = M->r3[1].p; a.r3z = M->r3[2].p;
    a.dx = M->damp[0].p; a.dy = M->damp[1].p; a.dz = M->damp[2].p;
    a.irho_c = M->irho_c;
    a.f = s.f;
    const int ntx = (M->N[0] + tti::TX - 1) / tti::TX, nty = (M->N[1] + tti::TY - 1) / tti::TY;
    int nzc = ctx->opt_zchunks;
    if (nzc <= 0) {
        // ~4 waves of one CTA per SM, chunks of at least 48 planes (14 warm-up planes each)
        nzc = (4 * ctx->num_sms + ntx * nty - 1) / (ntx * nty);
        nzc = std::max(1, std::min(nzc, M->N[2] / 48));
    }
    nzc = std::min(nzc, M->N[2]);
    a.zchunk = (M->N[2] + nzc - 1) / nzc;
    nzc = (M->N[2] + a.zchunk - 1) / a.zchunk;
    kern<<<dim3(ntx, nty, nzc), tti::NTHREADS, tti::SMEM, ctx->stream>>>(tm, a);
    launch_count(ctx, "tti3d_tma");
}

// One fused TTI step; false if this combination has no single-pass instantiation.
bool tti3d_step(const judi_model *M, const StepArgs &s)
{
    if (!tti3d_supported(M) || s.born || s.qext || s.f.wrec || !s.scratch || !s.scratch->has_tmap) return false;
    if (s.f.grad && s.f.ic != JUDI_IC_AS) return false;
    const int mask = (s.f.snap[0] ? TF_SNAP : 0) | (s.f.grad ? TF_GRAD : 0) | (s.f.illum ? TF_ILLUM : 0);
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
