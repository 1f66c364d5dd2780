// tma3d.cu -- the headline kernel: 3-D isotropic acoustic leapfrog step, constant density,
// for sm_100a.  Replaces the Devito-generated loop nest of acoustic_kernel
// (src/pysource/kernels.py:46-77 with FD_utils.py:19 `irho * u.laplace`) including the fused
// damping term, and optionally fuses the snapshot save (fields_exprs.py:12-36), the "as" imaging
// condition (sensitivity.py:55-69) and the illumination (fields_exprs.py:240-255).
//
// Design:
//  * 2.5-D blocking: a CTA owns a (TX x TY) tile of the (x,y) plane and marches along z (the
//    slowest axis) over one z-chunk.  x is the fastest axis = Julia's native order.
//  * Warp-specialised TMA pipeline.  One producer warp issues every global->shared transfer as
//    cp.async.bulk.tensor.3d: the haloed (TX+2HR) x (TY+2R) plane of u(t) into a ring of NS
//    slots, and the (TX x TY) tiles of u(t-1) and m of the output plane into a ring of NP slots.
//    Completion is signalled on "full" mbarriers; the 8 consumer warps hand slots back through
//    "empty" mbarriers (one elected arrive per warp).  There is no CTA-wide barrier in the main
//    loop, so warps drift apart and hide each other's latency.  Out-of-range box parts are
//    zero-filled by the TMA unit, which is exactly the Dirichlet halo of the reference.
//  * The z-neighbours live in a per-thread register queue of 2R+1 float4 (4 consecutive x per
//    thread) that is rotated at compile time (main loop unrolled by 2R+1); x and y neighbours
//    are read from the shared plane with 128-bit loads.
//  * u(t+1) overwrites u(t-1) in place with 128-bit stores; damping is separable (three 1-D
//    profiles) so it costs no HBM traffic, and tiles/planes outside the absorbing layer take a
//    branch-free fast path.
//  Algorithmic HBM traffic: 16 B per grid-point update (u(t), u(t-1), m read; u(t+1) written).
#include "common.cuh"

#include "tma_util.cuh"

// ---- kernel ------------------------------------------------------------------------------
struct TmaArgs {
    Grid g;
    Coefs cf;
    float *up;
    const float *dx, *dy, *dz;
    float sdamp;     // dt / irho
    float irho_c;
    int zchunk;      // planes per z-chunk
    int dlo[3], dhi[3];  // [dlo, dhi): indices where the damping profile of each axis is zero
    Fuse f;
    float *dd_out;       // Born, background pass: write delta = u+ - 2u + u-   (Grid layout)
    const float *qsrc;   // Born, linearised pass: delta of the background field ...
    const float *dm;     // ... and the perturbation: source term -dm * qsrc
    float *ulp;          // fused Born (DUAL): output of the linearised field, in place
    const float *qext;   // F_QEXT: space-time source of this step (Grid layout), added as
    float qext_scale;    //   qext_scale * qext  (= dt^2 / irho * wavelet[t], fields_exprs.py:58-81)
    const float *sfield; // F_QISIC: dt^2 div(dm grad u) of the background field (varc3d_lapw) ...
    float ics;           // ... and its sign: +1 isic, -1 fwi  (sensitivity.py:191-209)
};

// DUAL: fused Born step.  Two groups of consumer warps share one CTA: group 0 steps the background
// field u, group 1 the linearised field ul; both planes travel in one ring slot, the tiles
// (u-, m, ul-, dm) in one tile slot, and the Born source -dm (u+ - 2u + u-) goes from group 0 to
// group 1 through a shared-memory tile ring guarded by per-warp mbarriers.  32 B per grid point
// of HBM traffic instead of the 44 B of two launches that exchange delta through HBM.
// YR > 0: Y role.  NYT = tile points / (4 YR) extra threads each own YR consecutive rows of one
// 4-point group and evaluate the y part of the Laplacian of the centre plane for all of them from
// ONE window of YR + 2R row vectors (so = 16: 20 LDS.128 per 4 rows instead of 16 per row); the
// result reaches the consumer threads through a small shared-memory ring (one LDS.128).  Same
// operations in the same order as the consumer's own y loop: bit-identical wavefields.
template <int R, int TXT, int TYT, int NS, int NP, int NPW = 1, bool DUAL = false, int YR = 0>
struct TmaCfg {
    static constexpr int HR = ((R + 3) / 4) * 4;  // x halo in shared memory (float4 aligned)
    static constexpr int TX = 4 * TXT, TY = TYT;
    static constexpr int PX = TX + 2 * HR, PY = TY + 2 * R;
    static constexpr int SLOT = ((PX * PY * 4 + 127) / 128) * 128 / 4;  // floats per u slot
    static constexpr int PSLOT = TX * TY;                               // floats per tile
    static constexpr int NQ = 2 * R + 1;
    static constexpr int NCONS = TXT * TYT;       // consumer threads per group
    static constexpr int NG = DUAL ? 2 : 1;       // consumer groups
    static constexpr int ND = 2;                  // delta-exchange slots (DUAL)
    static constexpr int USTRIDE = NG * SLOT;     // floats per ring slot (u [, ul] plane)
    static constexpr int PSTRIDE = 2 * NG * PSLOT;  // floats per tile slot (u-, m [, ul-, dm])
    static constexpr int NYT = (YR > 0) * (NCONS / (YR > 0 ? YR : 1));   // Y-role threads
    static constexpr int NYW = NYT / 32;
    static constexpr int NY = 3;                      // y-sum ring slots
    static constexpr int NTHREADS = NG * NCONS + 32 * NPW + NYT;   // + producer warp(s) [+ Y warps]
    static constexpr int NBAR = 2 * NS + 2 * NP + (DUAL ? 2 * ND * (NCONS / 32) : 0) + (YR ? 2 * NY : 0);
    static constexpr size_t SMEM = (size_t)NS * USTRIDE * 4 + (size_t)NP * PSTRIDE * 4 +
                                   (DUAL ? (size_t)ND * PSLOT * 4 : 0) + (YR ? (size_t)NY * PSLOT * 4 : 0) +
                                   NBAR * 8 + 128;
    static_assert(YR == 0 || (!DUAL && TYT % YR == 0 && NYT % 32 == 0), "Y role: whole warps, no DUAL");
};

// FUSE is a compile-time mask of what rides on the step, so that each flavour of launch carries
// only its own instructions (the all-in-one fused instantiation executed 33% more instructions
// than the plain one for a single extra store):
// F_QEXT: extended / full-wavefield source read from one more field (+4 B per point);
// F_WREC: extended receiver  wrec += dt wavelet[t] u[t]  (fields_exprs.py:85-103, +8 B per point)
// F_QISIC (with F_QSRC): the isic / fwi Born source  -(dm m delta) + ics dt^2 div(dm grad u)
enum { F_SNAP = 1, F_GRAD = 2, F_ILLUM = 4, F_DDOUT = 8, F_QSRC = 16, F_QEXT = 32, F_WREC = 64, F_QISIC = 128 };

struct TmaMaps {   // tensor maps of one launch (kernel parameter, __grid_constant__)
    CUtensorMap U, P, M;      // haloed u(t), tile u(t-1), tile m
    CUtensorMap U2, P2, D;    // DUAL: haloed ul(t), tile ul(t-1), tile dm
};

template <int R, int TXT, int TYT, int NS, int NP, int MINB, int FUSE, int NPW, bool ROT, bool DUAL, int YR = 0>
__global__ void __launch_bounds__((DUAL ? 2 : 1) * TXT * TYT + 32 * NPW + (YR > 0) * (TXT * TYT / (YR > 0 ? YR : 1)), MINB)
    acoustic3d_tma(const __grid_constant__ TmaMaps tm, const TmaArgs a)
{
    typedef TmaCfg<R, TXT, TYT, NS, NP, NPW, DUAL, YR> C;
    constexpr int HR = C::HR, TX = C::TX, TY = C::TY, PX = C::PX, PY = C::PY, SLOT = C::SLOT;
    constexpr int PSLOT = C::PSLOT, NQ = C::NQ, NCW = C::NCONS / 32;
    constexpr int NG = C::NG, ND = C::ND, USTRIDE = C::USTRIDE, PSTRIDE = C::PSTRIDE;
    static_assert(!DUAL || NPW == 2, "the fused Born variant uses two producer warps");
    // STATIC: ring length == queue length, so ring slots are compile-time constants inside the
    // unrolled main loop and every shared-memory address is an immediate offset
    constexpr bool STATIC = ROT && (NS == NQ) && (NQ % NP == 0);
    // (ring lengths need not be powers of two: `% NS` with a constant is a multiply-shift)
    extern __shared__ unsigned char smem_raw[];
    float *ringU = (float *)(smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u));
    float *ringP = ringU + NS * USTRIDE;              // [NP][2 NG][PSLOT]: u(t-1), m [, ul(t-1), dm]
    float *ringD = ringP + NP * PSTRIDE;              // DUAL: [ND][PSLOT] delta of the background
    [[maybe_unused]] float *ringY = ringD + (DUAL ? ND * PSLOT : 0);   // YR: [NY][PSLOT] y sums
    uint64_t *fullU = (uint64_t *)(ringY + (YR ? C::NY * PSLOT : 0));
    uint64_t *emptyU = fullU + NS;
    uint64_t *fullP = emptyU + NS;
    uint64_t *emptyP = fullP + NP;
    [[maybe_unused]] uint64_t *fullD = emptyP + NP;   // DUAL: [ND][NCW], one per consumer warp
    [[maybe_unused]] uint64_t *emptyD = fullD + ND * NCW;
    [[maybe_unused]] uint64_t *fullY = emptyD + (DUAL ? ND * NCW : 0);   // YR: [NY]
    [[maybe_unused]] uint64_t *emptyY = fullY + C::NY;
    constexpr int NYW = C::NYW, NY = C::NY;

    const Grid &g = a.g;
    const Coefs &cf = a.cf;
    const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
    const int z0 = blockIdx.z * a.zchunk;
    const int z1 = min(z0 + a.zchunk, g.n[2]);
    const int nzc = z1 - z0;
    const int total = nzc + 2 * R;  // u planes this CTA consumes: z0-R .. z1-1+R

    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; s++) { mbar_init(&fullU[s], 1); mbar_init(&emptyU[s], NG * NCW + NYW); }
        if (YR)
            for (int s = 0; s < NY; s++) { mbar_init(&fullY[s], NYW); mbar_init(&emptyY[s], NCW); }
        for (int s = 0; s < NP; s++) { mbar_init(&fullP[s], 1); mbar_init(&emptyP[s], NG * NCW); }
        if (DUAL)
            for (int s = 0; s < ND * NCW; s++) { mbar_init(&fullD[s], 1); mbar_init(&emptyD[s], 1); }
        fence_barrier_init();
    }
    __syncthreads();

    if (YR > 0 && threadIdx.x >= NG * C::NCONS + 32 * NPW) {
        // ===== Y warps: y part of the Laplacian of every centre plane, YR rows per thread ========
        constexpr int YRR = YR > 0 ? YR : 1;
        const int yid = (int)threadIdx.x - (NG * C::NCONS + 32 * NPW);
        const int ytx = yid % TXT, yr0 = (yid / TXT) * YRR;
        const int ylane = threadIdx.x & 31;
        const int win = yr0 * PX + HR + 4 * ytx;        // first row of the window in a u slot
        const int yout = yr0 * TX + 4 * ytx;
        // planes 0 .. R-1 of the chunk are never centres (the consumers release them unread too)
        if (ylane == 0)
            for (int j = 0; j < R && j < total; j++) mbar_arrive(&emptyU[j % NS]);
        for (int i = 0; i < nzc; i++) {
            const int j = i + R, s = j % NS;
            mbar_wait(&fullU[s], (uint32_t)((j / NS) & 1));
            const float *col = ringU + s * USTRIDE + win;
            float4 w[YRR + 2 * R];
#pragma unroll
            for (int k = 0; k < YRR + 2 * R; k++) w[k] = *(const float4 *)(col + k * PX);
            float4 o[YRR];
#pragma unroll
            for (int e = 0; e < YRR; e++) {
                f2 ly0, ly1;
                {
                    const float4 p = w[e + R + 1], m_ = w[e + R - 1];
                    ly0 = mul2s(cf.cy[1], add2(lo2(p), lo2(m_)));
                    ly1 = mul2s(cf.cy[1], add2(hi2(p), hi2(m_)));
                }
#pragma unroll
                for (int k = 2; k <= R; k++) {
                    const float4 p = w[e + R + k], m_ = w[e + R - k];
                    ly0 = fma2s(cf.cy[k], add2(lo2(p), lo2(m_)), ly0);
                    ly1 = fma2s(cf.cy[k], add2(hi2(p), hi2(m_)), ly1);
                }
                o[e] = cat4(ly0, ly1);
            }
            // the plane is dead for this warp (its values are in registers)
            __syncwarp();
            if (ylane == 0) mbar_arrive(&emptyU[s]);
            const int sy = i % NY;
            if (i >= NY) mbar_wait(&emptyY[sy], (uint32_t)(((i / NY) - 1) & 1));
            float *yo = ringY + sy * PSLOT + yout;
#pragma unroll
            for (int e = 0; e < YRR; e++) *(float4 *)(yo + e * TX) = o[e];
            __syncwarp();
            if (ylane == 0) mbar_arrive(&fullY[sy]);
        }
        return;
    }
    if (threadIdx.x >= NG * C::NCONS) {
        // ===== producer warp(s): one lane issues the TMA loads of this CTA =====================
        // NPW == 1: a single lane issues both streams; NPW == 2: warp 0 streams the haloed u(t)
        // planes, warp 1 the u(t-1)/m tiles, so neither ring can stall the other.
        const int pw = (threadIdx.x - NG * C::NCONS) >> 5;
        if ((threadIdx.x & 31) == 0) {
            const int bx = g.hx + x0 - HR, by = g.hy + y0 - R, bz = g.hz + z0 - R;
            const int tx0 = g.hx + x0, ty0 = g.hy + y0, tz0 = g.hz + z0;
            constexpr uint32_t UBYTES = NG * PX * PY * 4, PBYTES = 2 * NG * PSLOT * 4;
            const CUtensorMap *tmU = &tm.U, *tmP = &tm.P, *tmM = &tm.M;
            if (NPW == 1) {
                for (int j = 0; j < total; j++) {
                    const int s = j % NS;
                    if (j >= NS) mbar_wait(&emptyU[s], (uint32_t)(((j / NS) - 1) & 1));
                    mbar_arrive_expect_tx(&fullU[s], UBYTES);
                    tma_load_3d(ringU + s * USTRIDE, tmU, bx, by, bz + j, &fullU[s]);
                    // tiles of output plane z0+i are consumed at iteration i+2R: issue them one
                    // iteration early (the ring depth NP then allows NP-2 planes of run-ahead)
                    const int i = j - 2 * R + 1;
                    if (i >= 0 && i < nzc) {
                        const int sp = i % NP;
                        if (i >= NP) mbar_wait(&emptyP[sp], (uint32_t)(((i / NP) - 1) & 1));
                        mbar_arrive_expect_tx(&fullP[sp], PBYTES);
                        tma_load_3d(ringP + sp * PSTRIDE, tmP, tx0, ty0, tz0 + i, &fullP[sp]);
                        tma_load_3d(ringP + sp * PSTRIDE + PSLOT, tmM, tx0, ty0, tz0 + i, &fullP[sp]);
                    }
                }
            } else if (pw == 0) {
                for (int j = 0; j < total; j++) {
                    const int s = j % NS;
                    if (j >= NS) mbar_wait(&emptyU[s], (uint32_t)(((j / NS) - 1) & 1));
                    mbar_arrive_expect_tx(&fullU[s], UBYTES);
                    tma_load_3d(ringU + s * USTRIDE, tmU, bx, by, bz + j, &fullU[s]);
                    if (DUAL) tma_load_3d(ringU + s * USTRIDE + SLOT, &tm.U2, bx, by, bz + j, &fullU[s]);
                }
            } else {
                for (int i = 0; i < nzc; i++) {
                    const int sp = i % NP;
                    if (i >= NP) mbar_wait(&emptyP[sp], (uint32_t)(((i / NP) - 1) & 1));
                    mbar_arrive_expect_tx(&fullP[sp], PBYTES);
                    tma_load_3d(ringP + sp * PSTRIDE, tmP, tx0, ty0, tz0 + i, &fullP[sp]);
                    tma_load_3d(ringP + sp * PSTRIDE + PSLOT, tmM, tx0, ty0, tz0 + i, &fullP[sp]);
                    if (DUAL) {
                        tma_load_3d(ringP + sp * PSTRIDE + 2 * PSLOT, &tm.P2, tx0, ty0, tz0 + i, &fullP[sp]);
                        tma_load_3d(ringP + sp * PSTRIDE + 3 * PSLOT, &tm.D, tx0, ty0, tz0 + i, &fullP[sp]);
                    }
                }
            }
        }
        return;
    }

    // ===== consumer warps ===================================================================
    // grp 0 steps the background field (and is the only group when !DUAL), grp 1 the linearised one
    const int grp = DUAL ? (int)threadIdx.x / C::NCONS : 0;
    const int ctid = (int)threadIdx.x - grp * C::NCONS;
    const int tx = ctid % TXT, ty = ctid / TXT;
    const int lane = threadIdx.x & 31;
    [[maybe_unused]] const int wg = ctid >> 5;          // warp index inside the group
    const float *ringUg = ringU + grp * SLOT;           // this group's plane inside a ring slot
    float *const outp = (DUAL && grp == 1) ? a.ulp : a.up;
    const int x = x0 + 4 * tx, y = y0 + ty;
    const bool act = x < g.n[0] && y < g.n[1];
    const bool xfull = x + 3 < g.n[0];
    float4 q[NQ];
#pragma unroll
    for (int k = 0; k < NQ; k++) q[k] = make_float4(0.f, 0.f, 0.f, 0.f);

    // damping along x and y for this thread's points: (dx + dy), z added per plane
    f2 sxy0 = pk1(0.f), sxy1 = pk1(0.f);
    if (act) {
        const float4 d4 = *(const float4 *)(a.dx + x);  // profile arrays are padded 8 past n[0]
        const float dyv = a.dy[y];
        sxy0 = add2(lo2(d4), pk1(dyv));
        sxy1 = add2(hi2(d4), pk1(dyv));
    }
    const bool tile_interior = x0 >= a.dlo[0] && x0 + TX <= a.dhi[0] && y0 >= a.dlo[1] &&
                               y0 + TY <= a.dhi[1];
    const int sm_own = (ty + R) * PX + HR + 4 * tx;  // this thread's first point in a u slot
    const int sm_tile = ty * TX + 4 * tx;            // ... and in a tile slot
    long long goff = g.at(x, y, z0);                 // advanced by sz per output plane
    // fused extras: everything that does not depend on z is decided once per thread; the region
    // offset advances by one region plane per output plane (no 64-bit index math in the loop)
    [[maybe_unused]] const Fuse &fz = a.f;
    constexpr bool f_snap = (FUSE & F_SNAP) != 0, f_grad = (FUSE & F_GRAD) != 0;
    constexpr bool f_il = (FUSE & F_ILLUM) != 0, f_dd = (FUSE & F_DDOUT) != 0;
    constexpr bool f_wr = (FUSE & F_WREC) != 0;
    constexpr bool f_reg = f_snap || f_grad || f_il || f_wr;   // anything stored in the region layout
    [[maybe_unused]] const bool f_born = (FUSE & F_QSRC) != 0 && act;
    [[maybe_unused]] const bool f_qx = (FUSE & F_QEXT) != 0 && act;
    [[maybe_unused]] const bool f_xy = f_reg && act && grp == 0 && y >= fz.reg.lo[1] &&
                                       y < fz.reg.hi[1] && x + 3 >= fz.reg.lo[0] && x < fz.reg.hi[0];
    [[maybe_unused]] const bool f_vec = f_xy && (fz.reg.lo[0] & 3) == 0 && x >= fz.reg.lo[0] &&
                                        x + 3 < fz.reg.hi[0];
    [[maybe_unused]] long long ri = f_reg ? fz.reg.at(x, y, z0) : 0;
    [[maybe_unused]] const long long rstride = f_reg ? (long long)fz.reg.pitch * fz.reg.ext[1] : 0;
    [[maybe_unused]] const float gc = f_grad ? fz.gcoef * a.irho_c : 0.f;
    [[maybe_unused]] const int rz0 = f_reg ? fz.reg.lo[2] : 0, rz1 = f_reg ? fz.reg.hi[2] : 0;

    // ROT: the z-queue is rotated at compile time (main loop unrolled by NQ, no register moves,
    // ~9x the code); !ROT: the queue is shifted by NQ-1 register moves per plane (compact code).
    constexpr int UNR = ROT ? NQ : 1;
    for (int jt0 = 0; jt0 < total; jt0 += UNR) {
#pragma unroll
        for (int ph = 0; ph < UNR; ph++) {
            const int jt = jt0 + ph;
            if (jt >= total) break;
            // queue element of z-offset k (k = R is the centre, k = 2R the incoming plane)
#define Q(k) q[ROT ? (((ph) + 1 + (k)) % NQ) : (k)]
            if (!ROT) {
#pragma unroll
                for (int k = 0; k < NQ - 1; k++) q[k] = q[k + 1];
            }
            // ---- incoming plane z + R ------------------------------------------------------
            const int st = STATIC ? ph : (jt % NS);                     // slot of the incoming plane
            const int sc = STATIC ? (ph + NQ - R) % NQ : ((jt + NS - R) % NS);  // slot of the centre
            // fused extras: issue their point-wise global loads before waiting on the pipeline so
            // that their latency hides behind the stencil arithmetic of this plane
            [[maybe_unused]] bool fin = false, fvec = false;
            [[maybe_unused]] float4 us4, g4, il4, qs4, dm4, qx4, wr4, sf4;
            if (FUSE != 0 && jt >= 2 * R) {
                const int z = z0 + jt - 2 * R;
                const bool zin = z >= rz0 && z < rz1;
                fin = f_xy && zin;
                fvec = f_vec && zin;
                // the imaging condition streams the snapshot and the gradient through the consumer
                // threads: pull the lines of plane z + 3 into L2 now so the loads below cost an
                // L2 hit, not a DRAM round trip, when that plane comes up
                if (f_grad && f_vec && z + 3 >= rz0 && z + 3 < rz1) {
                    prefetch_l2(fz.usnap[0] + ri + 3 * rstride);
                    prefetch_l2(fz.grad + ri + 3 * rstride);
                }
                if ((f_grad || f_il || f_wr) && fvec) {
                    if (f_grad) {
                        us4 = __ldg((const float4 *)(fz.usnap[0] + ri));
                        g4 = *(const float4 *)(fz.grad + ri);
                    }
                    if (f_il) il4 = *(const float4 *)(fz.illum + ri);
                    if (f_wr) wr4 = *(const float4 *)(fz.wrec + ri);
                }
                // (a partial group at the x edge reads into the zero halo of the allocation)
                if (f_qx) qx4 = __ldg((const float4 *)(a.qext + goff));
                if (f_born) {
                    qs4 = *(const float4 *)(a.qsrc + goff);
                    dm4 = __ldg((const float4 *)(a.dm + goff));
                    if (FUSE & F_QISIC) sf4 = *(const float4 *)(a.sfield + goff);
                }
            }
            mbar_wait(&fullU[st], (uint32_t)((jt / NS) & 1));
            Q(2 * R) = *(const float4 *)(ringUg + st * USTRIDE + sm_own);
            if (jt >= 2 * R) {
                const int i = jt - 2 * R;  // output plane index inside the chunk
                const int z = z0 + i;
                const float *cen = ringUg + sc * USTRIDE;
                // All point-wise arithmetic is packed fp32x2 on the pairs (x, x+1), (x+2, x+3):
                // z part (register queue), x part (row), y part (column): three accumulators
                const float4 cq = Q(R);
                const f2 uc0 = lo2(cq), uc1 = hi2(cq);
                f2 lz0 = mul2s(cf.c0, uc0), lz1 = mul2s(cf.c0, uc1);
#pragma unroll
                for (int k = 1; k <= R; k++) {
                    const float4 p = Q(R + k), m_ = Q(R - k);
                    lz0 = fma2s(cf.cz[k], add2(lo2(p), lo2(m_)), lz0);
                    lz1 = fma2s(cf.cz[k], add2(hi2(p), hi2(m_)), lz1);
                }
                f2 lxp[2];
                {
                    const float *crow = cen + sm_own - HR;
                    float w[4 + 2 * HR];
#pragma unroll
                    for (int v = 0; v < 1 + 2 * (HR / 4); v++) {
                        if (v == HR / 4) {  // the centre vector is already in the queue
                            w[4 * v] = cq.x; w[4 * v + 1] = cq.y; w[4 * v + 2] = cq.z; w[4 * v + 3] = cq.w;
                        } else {
                            const float4 t = *(const float4 *)(crow + 4 * v);
                            w[4 * v] = t.x; w[4 * v + 1] = t.y; w[4 * v + 2] = t.z; w[4 * v + 3] = t.w;
                        }
                    }
                    // even offsets keep the register pairs aligned (packed); odd ones straddle
                    // two pairs and stay scalar
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const int b = HR + 2 * h;
                        lxp[h].x = __fmul_rn(cf.cx[1], __fadd_rn(w[b + 1], w[b - 1]));
                        lxp[h].y = __fmul_rn(cf.cx[1], __fadd_rn(w[b + 2], w[b]));
#pragma unroll
                        for (int k = 2; k <= R; k++) {
                            if (k & 1) {
                                lxp[h].x = __fmaf_rn(cf.cx[k], __fadd_rn(w[b + k], w[b - k]), lxp[h].x);
                                lxp[h].y = __fmaf_rn(cf.cx[k], __fadd_rn(w[b + 1 + k], w[b + 1 - k]), lxp[h].y);
                            } else {
                                lxp[h] = fma2s(cf.cx[k], add2(pk(w[b + k], w[b + k + 1]),
                                                             pk(w[b - k], w[b - k + 1])), lxp[h]);
                            }
                        }
                    }
                }
                f2 ly0, ly1;
                if (YR > 0) {
                    // the Y warps have the y sums of this plane (or will in a moment)
                    const int sy = i % NY;
                    mbar_wait(&fullY[sy], (uint32_t)((i / NY) & 1));
                    const float4 t = *(const float4 *)(ringY + sy * PSLOT + sm_tile);
                    ly0 = lo2(t); ly1 = hi2(t);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&emptyY[sy]);
                } else {
                    const float *ccol = cen + sm_own;
                    {
                        const float4 p = *(const float4 *)(ccol + PX), m_ = *(const float4 *)(ccol - PX);
                        ly0 = mul2s(cf.cy[1], add2(lo2(p), lo2(m_)));
                        ly1 = mul2s(cf.cy[1], add2(hi2(p), hi2(m_)));
                    }
#pragma unroll
                    for (int k = 2; k <= R; k++) {
                        const float4 p = *(const float4 *)(ccol + k * PX);
                        const float4 m_ = *(const float4 *)(ccol - k * PX);
                        ly0 = fma2s(cf.cy[k], add2(lo2(p), lo2(m_)), ly0);
                        ly1 = fma2s(cf.cy[k], add2(hi2(p), hi2(m_)), ly1);
                    }
                }
                // the centre plane is dead for this warp: hand its slot back to the producer
                __syncwarp();
                if (lane == 0) mbar_arrive(&emptyU[sc]);
                // ---- u(t-1) and m tiles of the output plane ----------------------------------
                const int sp = STATIC ? (ph + 4 * NQ - 2 * R) % NP : (i % NP);
                mbar_wait(&fullP[sp], (uint32_t)((i / NP) & 1));
                const float *pt = ringP + sp * PSTRIDE + sm_tile;
                const float4 up4 = *(const float4 *)(pt + (DUAL ? 2 * grp * PSLOT : 0));
                const float4 m4 = *(const float4 *)(pt + PSLOT);
                [[maybe_unused]] float4 dmt4;
                if (DUAL && grp == 1) dmt4 = *(const float4 *)(pt + 3 * PSLOT);
                __syncwarp();
                if (lane == 0) mbar_arrive(&emptyP[sp]);
                if (DUAL && grp == 1) {
                    // Born source of the linearised field: -dm * delta of the background, handed
                    // over by the same warp of group 0 through shared memory
                    const int sd = i % ND;
                    mbar_wait(&fullD[sd * NCW + wg], (uint32_t)((i / ND) & 1));
                    const float4 dd4 = *(const float4 *)(ringD + sd * PSLOT + sm_tile);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&emptyD[sd * NCW + wg]);
                    lz0.x = __fmaf_rn(-dmt4.x, dd4.x, lz0.x); lz0.y = __fmaf_rn(-dmt4.y, dd4.y, lz0.y);
                    lz1.x = __fmaf_rn(-dmt4.z, dd4.z, lz1.x); lz1.y = __fmaf_rn(-dmt4.w, dd4.w, lz1.y);
                }
                if (f_born && (FUSE & F_QISIC)) {
                    // isic / fwi Born source (dt^2/irho) q = -(dm m delta) + ics dt^2 div(dm grad u)
                    const f2 q0 = fma2s(a.ics, lo2(sf4), mul2(mul2(mul2s(-1.f, lo2(dm4)), lo2(m4)), lo2(qs4)));
                    const f2 q1 = fma2s(a.ics, hi2(sf4), mul2(mul2(mul2s(-1.f, hi2(dm4)), hi2(m4)), hi2(qs4)));
                    lz0 = add2(lz0, q0);
                    lz1 = add2(lz1, q1);
                } else if (f_born) {  // Born source (dt^2/irho) q = -dm * (u+ - 2u + u-)
                    lz0.x = __fmaf_rn(-dm4.x, qs4.x, lz0.x); lz0.y = __fmaf_rn(-dm4.y, qs4.y, lz0.y);
                    lz1.x = __fmaf_rn(-dm4.z, qs4.z, lz1.x); lz1.y = __fmaf_rn(-dm4.w, qs4.w, lz1.y);
                }
                if (f_qx) {    // + dt^2 / irho * q(x, t)
                    lz0 = fma2s(a.qext_scale, lo2(qx4), lz0);
                    lz1 = fma2s(a.qext_scale, hi2(qx4), lz1);
                }
                // explicit rn operations: the plain and the fused instantiation must produce
                // bit-identical wavefields (checkpoint recomputation relies on it)
                const f2 lap0 = add2(add2(lz0, lxp[0]), ly0), lap1 = add2(add2(lz1, lxp[1]), ly1);
                const f2 d10 = sub2(uc0, lo2(up4)), d11 = sub2(uc1, hi2(up4));
                // u+ = (u + (u - u-)) + t * r with t the numerator and r the reciprocal of the
                // denominator of the increment; written as ONE explicit fma so that ptxas cannot
                // contract it in one instantiation and not in the other (it fuses mul.rn.f32x2 +
                // add.rn.f32x2 when the product has no other use)
                f2 t0, t1, r0, r1;
                if (tile_interior && z >= a.dlo[2] && z < a.dhi[2]) {
                    t0 = lap0; t1 = lap1;
                    r0 = pk(rcp_approx(m4.x), rcp_approx(m4.y));
                    r1 = pk(rcp_approx(m4.z), rcp_approx(m4.w));
                } else {
                    // ns = -(dx + dy + dz) * dt / irho: the sign rides on the scalar (exact)
                    const f2 szv = pk1(a.dz[z]);
                    const f2 ns0 = mul2s(-a.sdamp, add2(sxy0, szv)), ns1 = mul2s(-a.sdamp, add2(sxy1, szv));
                    const f2 den0 = sub2(lo2(m4), ns0), den1 = sub2(hi2(m4), ns1);
                    t0 = fma2(ns0, d10, lap0); t1 = fma2(ns1, d11, lap1);
                    r0 = pk(rcp_approx(den0.x), rcp_approx(den0.y));
                    r1 = pk(rcp_approx(den1.x), rcp_approx(den1.y));
                }
                const f2 un0 = fma2(t0, r0, add2(uc0, d10)), un1 = fma2(t1, r1, add2(uc1, d11));
                // delta = u+ - 2u + u- (imaging condition, Born source): fused variants only
                [[maybe_unused]] f2 dl0, dl1;
                if (f_grad || f_dd || DUAL) { dl0 = mul2(t0, r0); dl1 = mul2(t1, r1); }
                if (DUAL && grp == 0) {
                    const int sd = i % ND;
                    if (i >= ND) mbar_wait(&emptyD[sd * NCW + wg], (uint32_t)(((i / ND) - 1) & 1));
                    *(float4 *)(ringD + sd * PSLOT + sm_tile) = cat4(dl0, dl1);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&fullD[sd * NCW + wg]);
                }
                if (act) {
                    float *gout = outp + goff;
                    if (xfull) *(float4 *)gout = cat4(un0, un1);
                    else {
                        const float un[4] = {un0.x, un0.y, un1.x, un1.y};
#pragma unroll
                        for (int e = 0; e < 4; e++)
                            if (x + e < g.n[0]) gout[e] = un[e];
                    }
                    if (f_dd) {
                        float *dout = a.dd_out + goff;
                        if (xfull) *(float4 *)dout = cat4(dl0, dl1);
                        else {
                            const float delta[4] = {dl0.x, dl0.y, dl1.x, dl1.y};
#pragma unroll
                            for (int e = 0; e < 4; e++)
                                if (x + e < g.n[0]) dout[e] = delta[e];
                        }
                    }
                }
                if (f_reg) {
                    if (fin) {
                        if (fvec) {
                            // written once, read a whole sweep later: streaming store
                            if (f_snap) __stcs((float4 *)(fz.snap[0] + ri), cq);
                            if (f_il) {
                                const f2 I0 = fma2(mul2s(cf.dt, uc0), uc0, lo2(il4));
                                const f2 I1 = fma2(mul2s(cf.dt, uc1), uc1, hi2(il4));
                                *(float4 *)(fz.illum + ri) = cat4(I0, I1);
                            }
                            if (f_grad) {
                                const f2 G0 = fma2(mul2s(-gc, lo2(us4)), dl0, lo2(g4));
                                const f2 G1 = fma2(mul2s(-gc, hi2(us4)), dl1, hi2(g4));
                                *(float4 *)(fz.grad + ri) = cat4(G0, G1);
                            }
                            if (f_wr)
                                *(float4 *)(fz.wrec + ri) = cat4(fma2s(fz.wrec_scale, uc0, lo2(wr4)),
                                                                 fma2s(fz.wrec_scale, uc1, hi2(wr4)));
                        } else {
                            const float uc[4] = {cq.x, cq.y, cq.z, cq.w};
                            const float delta[4] = {dl0.x, dl0.y, dl1.x, dl1.y};
#pragma unroll
                            for (int e = 0; e < 4; e++) {
                                if (x + e < fz.reg.lo[0] || x + e >= fz.reg.hi[0]) continue;
                                if (f_snap) fz.snap[0][ri + e] = uc[e];
                                if (f_il)
                                    fz.illum[ri + e] = __fmaf_rn(__fmul_rn(cf.dt, uc[e]), uc[e], fz.illum[ri + e]);
                                if (f_grad)
                                    fz.grad[ri + e] = __fmaf_rn(__fmul_rn(-gc, fz.usnap[0][ri + e]), delta[e],
                                                                fz.grad[ri + e]);
                                if (f_wr) fz.wrec[ri + e] = __fmaf_rn(fz.wrec_scale, uc[e], fz.wrec[ri + e]);
                            }
                        }
                    }
                }
                goff += g.sz;
                if (f_reg) ri += rstride;
            } else if (jt >= R) {
                // planes below the chunk are never centres: release them as soon as R newer
                // planes have arrived (keeps the producer's slot accounting uniform)
                __syncwarp();
                if (lane == 0) mbar_arrive(&emptyU[sc]);
            }
#undef Q
        }
    }
}

// ---- host side -------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                  const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct TileVariant { int txt, tyt; };
static const TileVariant kVariants[] = {{16, 16}, {32, 8}, {16, 8}, {32, 4},
                                        {16, 16}, {32, 8}, {16, 16}, {16, 8},
                                        {16, 16}, {32, 8}};
static const int kNumVariants = 10;

static int variant_of(const judi_model *M)
{
    int v = M->ctx->opt_tile;
    if ((v != 6 && v != 9) || (M->r != 4 && v == 9)) v = 8;
    return v;
}

// the combinations of fused extras launch_variant has an instantiation for
bool tma3d_fuse_supported(const StepArgs &s)
{
    const int m = (s.f.snap[0] ? F_SNAP : 0) | (s.f.grad ? F_GRAD : 0) | (s.f.illum ? F_ILLUM : 0) |
                  (s.qext ? F_QEXT : 0) | (s.f.wrec ? F_WREC : 0);
    if (!(m & (F_QEXT | F_WREC))) return true;
    if (s.born) return m == F_QEXT || m == (F_QEXT | F_SNAP);
    switch (m) {
    case F_QEXT: case F_QEXT | F_SNAP: case F_QEXT | F_ILLUM: case F_QEXT | F_SNAP | F_ILLUM:
    case F_WREC: case F_WREC | F_ILLUM:
        return true;
    default:
        return false;
    }
}

bool tma3d_supported(const judi_model *M)
{
    return M->ndim == 3 && !M->tti && !M->irho_field && !M->visco && (M->r == 2 || M->r == 4 || M->r == 8) &&
           M->ctx->encode_tiled != nullptr;
}

void tma_encode_box(const judi_model *M, const float *base, int bx, int by, CUtensorMap *out)
{
    const Grid &g = M->g;
    cuuint64_t dims[3] = {(cuuint64_t)g.px, (cuuint64_t)g.py, (cuuint64_t)g.pz};
    cuuint64_t strides[2] = {(cuuint64_t)g.sy * 4, (cuuint64_t)g.sz * 4};
    cuuint32_t box[3] = {(cuuint32_t)bx, (cuuint32_t)by, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult rc = ((EncodeTiledFn)M->ctx->encode_tiled)(
        out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void *)base, dims, strides, box, estr,
        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) {
        char b[128];
        snprintf(b, sizeof(b), "cuTensorMapEncodeTiled failed with CUresult %d", (int)rc);
        throw JudiError(b);
    }
}

// halo-box map (whole haloed plane of u(t)) and tile-box map (u(t-1), m) of one buffer
void tma3d_encode(const judi_model *M, const float *base, CUtensorMap *halo_map, CUtensorMap *tile_map)
{
    int v = variant_of(M);
    int R = M->r, HR = ((R + 3) / 4) * 4;
    int tx = 4 * kVariants[v].txt, ty = kVariants[v].tyt;
    if (halo_map) tma_encode_box(M, base, tx + 2 * HR, ty + 2 * R, halo_map);
    if (tile_map) tma_encode_box(M, base, tx, ty, tile_map);
}

template <int R, int TXT, int TYT, int NS, int NP, int MINB, int NPW, bool ROT, int FUSE,
          bool DUAL = false, int YR = 0>
static void launch_fused(const judi_model *M, const StepArgs &s)
{
    typedef TmaCfg<R, TXT, TYT, NS, NP, NPW, DUAL, YR> C;
    judi_ctx *ctx = M->ctx;
    auto kern = acoustic3d_tma<R, TXT, TYT, NS, NP, MINB, FUSE, NPW, ROT, DUAL, YR>;
    static bool configured = false;
    if (!configured) {
        CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)C::SMEM));
        configured = true;
    }
    TmaArgs a;
    memset(&a, 0, sizeof(a));
    a.g = M->g;
    a.cf = M->cf;
    a.up = s.up[0];
    a.dx = M->damp[0].p; a.dy = M->damp[1].p; a.dz = M->damp[2].p;
    a.sdamp = M->dt / M->irho_c;
    a.irho_c = M->irho_c;
    for (int d = 0; d < 3; d++) {
        a.dlo[d] = M->padl[d] > 3 ? M->padl[d] - 3 : 0;
        a.dhi[d] = M->N[d] - (M->padr[d] > 3 ? M->padr[d] - 3 : 0);
    }
    a.f = s.f;
    a.dd_out = s.dd_out;
    a.qsrc = s.qsrc;
    a.dm = s.qsrc ? M->dm_src() : nullptr;
    a.qext = s.qext;
    a.qext_scale = s.qext_scale;
    a.sfield = s.sfield;
    a.ics = s.born_ic == JUDI_IC_FWI ? -1.f : 1.f;
    TmaMaps tm;
    tm.U = *s.tmap; tm.P = *s.tmap_prev; tm.M = *s.tmap_m;
    if (DUAL) {
        JUDI_REQUIRE(s.tmap2 && s.tmap_prev2 && s.tmap_dm && s.ulp[0], "fused Born: maps missing");
        tm.U2 = *s.tmap2; tm.P2 = *s.tmap_prev2; tm.D = *s.tmap_dm;
        a.ulp = s.ulp[0];
    } else {
        tm.U2 = tm.U; tm.P2 = tm.P; tm.D = tm.M;
    }
    int ntx = (M->N[0] + C::TX - 1) / C::TX, nty = (M->N[1] + C::TY - 1) / C::TY;
    int nzc = ctx->opt_zchunks;
    if (nzc <= 0) {
        // aim at ~6 waves of CTAs so the tail of the last wave is small, chunks >= 24 planes
        int slots = ctx->num_sms * MINB;
        nzc = (6 * slots + ntx * nty - 1) / (ntx * nty);
        int maxc = M->N[2] / 24;
        if (nzc > maxc) nzc = maxc;
        if (nzc < 1) nzc = 1;
    }
    if (nzc > M->N[2]) nzc = M->N[2];
    a.zchunk = (M->N[2] + nzc - 1) / nzc;
    nzc = (M->N[2] + a.zchunk - 1) / a.zchunk;
    dim3 grd(ntx, nty, nzc);
    kern<<<grd, C::NTHREADS, C::SMEM, ctx->stream>>>(tm, a);
    launch_count(ctx, "acoustic3d_tma");
}

// what the step carries -> the instantiation compiled for exactly that
static int fuse_mask(const StepArgs &s)
{
    return (s.f.snap[0] ? F_SNAP : 0) | (s.f.grad ? F_GRAD : 0) | (s.f.illum ? F_ILLUM : 0) |
           (s.dd_out ? F_DDOUT : 0) | (s.qsrc ? F_QSRC : 0) | (s.qext ? F_QEXT : 0) |
           (s.f.wrec ? F_WREC : 0) | ((s.qsrc && s.sfield) ? F_QISIC : 0);
}

template <int R, int TXT, int TYT, int NS, int NP, int MINB, int NPW = 1, bool ROT = true, int YR = 0>
static void launch_variant(const judi_model *M, const StepArgs &s)
{
#define FUSE_CASE(F) case F: launch_fused<R, TXT, TYT, NS, NP, MINB, NPW, ROT, F, false, YR>(M, s); return;
    switch (fuse_mask(s)) {
        FUSE_CASE(0)
        FUSE_CASE(F_SNAP)
        FUSE_CASE(F_GRAD)
        FUSE_CASE(F_ILLUM)
        FUSE_CASE(F_SNAP | F_ILLUM)
        FUSE_CASE(F_GRAD | F_ILLUM)
        FUSE_CASE(F_DDOUT)
        FUSE_CASE(F_DDOUT | F_SNAP)
        FUSE_CASE(F_DDOUT | F_ILLUM)
        FUSE_CASE(F_DDOUT | F_SNAP | F_ILLUM)
        FUSE_CASE(F_QSRC)
        FUSE_CASE(F_QSRC | F_QISIC)
        // extended / full-wavefield sources (forward_rec_w, forward_wf_src, J_adjoint(ws=...),
        // born_rec_w) and the extended receiver of adjoint_w (interface.py:50-204, 244-276)
        FUSE_CASE(F_QEXT)
        FUSE_CASE(F_QEXT | F_SNAP)
        FUSE_CASE(F_QEXT | F_ILLUM)
        FUSE_CASE(F_QEXT | F_SNAP | F_ILLUM)
        FUSE_CASE(F_QEXT | F_DDOUT)
        FUSE_CASE(F_QEXT | F_DDOUT | F_SNAP)
        FUSE_CASE(F_WREC)
        FUSE_CASE(F_WREC | F_ILLUM)
    default:
        throw JudiError("acoustic3d_tma: unsupported combination of fused extras");
    }
#undef FUSE_CASE
}

// Fused Born step (background + linearised field in one launch); false if this radius / tile
// has no fused instantiation (the caller then runs the two-launch form).
bool tma3d_step_born(const judi_model *M, const StepArgs &s)
{
    if (variant_of(M) != 8 || s.f.grad || s.dd_out || s.qsrc || s.qext || s.f.wrec) return false;
    const int mask = (s.f.snap[0] ? F_SNAP : 0) | (s.f.illum ? F_ILLUM : 0);
#define BORN_CASE(RR, F) \
    case F: launch_fused<RR, 16, 16, 8, 4, 1, 2, false, F, true>(M, s); return true;
    if (M->r == 4) {
        switch (mask) { BORN_CASE(4, 0) BORN_CASE(4, F_SNAP) BORN_CASE(4, F_ILLUM) BORN_CASE(4, F_SNAP | F_ILLUM) }
    } else if (M->r == 2) {
        switch (mask) { BORN_CASE(2, 0) BORN_CASE(2, F_SNAP) BORN_CASE(2, F_ILLUM) BORN_CASE(2, F_SNAP | F_ILLUM) }
    }
#undef BORN_CASE
    return false;
}

void tma3d_step(const judi_model *M, const StepArgs &s)
{
    JUDI_REQUIRE(s.tmap && s.tmap_prev && s.tmap_m, "tensor maps missing");
    int v = variant_of(M);
    // so = 4: two producer warps (the u(t-1) / m tiles then run NP planes ahead instead of one) are
    // 3 % faster than one (0.175 vs 0.180 ms on the C3 grid, profiles/r02_so16_sweep.log)
    if (M->r == 2) { launch_variant<2, 16, 16, 8, 4, 2, 2, false>(M, s); return; }
    if (M->r == 8) {
        // so = 16: 80 x 32 haloed planes, 16-slot ring, one CTA per SM, two producer warps, shifted
        // queue: 0.30 ms/step on the C3 grid; every other shape measures the same 0.33 ms as the
        // one-producer form (profiles/r02_so16_sweep.log): 32 x 16 or 64 x 8 tiles at two CTAs per
        // SM, a 12-slot ring, rotated vs shifted queue.
        // Y role (ctx option "tile" 34: Y warps, 4 rows per thread) -- measured, not the default: the
        // consumers' 16 y loads become one and the shared-memory wavefronts of a plane drop by 36 %
        // (ncu: data pipe 65 % -> 41 % busy), results bit-identical, and the step is 5 % SLOWER
        // (0.318 vs 0.302 ms): this kernel is bound by instruction issue and latency at 2.5 warps per
        // scheduler (161 registers, one CTA per SM, 56 % issue-active, 19 % of the instructions are
        // register-queue moves), not by the shared-memory pipe; a tile ring of 6 - 8 slots instead
        // of 4 is 6 - 10 % slower too (profiles/r02_so16_yrole.md).
        if (M->ctx->opt_tile == 34) launch_variant<8, 16, 16, 16, 4, 1, 2, false, 4>(M, s);
        else launch_variant<8, 16, 16, 16, 4, 1, 2, false>(M, s);
        return;
    }
    switch (v) {
    case 6: launch_variant<4, 16, 16, 8, 4, 2, 2>(M, s); break;          // rotated queue (x9 code)
    case 9: launch_variant<4, 32, 8, 8, 4, 2, 2, false>(M, s); break;   // 128 x 8 tile
    default: launch_variant<4, 16, 16, 8, 4, 2, 2, false>(M, s); break; // 8: shifted queue
    }
}
