// tma3d.cu -- the headline kernel: 3-D isotropic acoustic leapfrog step, constant density,
// for sm_100a.  Replaces the Devito-generated loop nest of acoustic_kernel
// (src/pysource/kernels.py:46-77 with FD_utils.py:19 `irho * u.laplace`) including the fused
// damping term, and optionally fuses the snapshot save (fields_exprs.py:12-36), the "as" imaging
// condition (sensitivity.py:55-69) and the illumination (fields_exprs.py:240-255).
//
// Design:
//  * 2.5-D blocking: a CTA owns a (TX x TY) tile of the (x,y) plane and marches along z (the
//    slowest axis) over one z-chunk.  x is the fastest axis = Julia's native order.
//  * Each (TX+2HR) x (TY+2R) haloed plane is fetched ONCE per CTA by a single
//    cp.async.bulk.tensor.3d (TMA) into a ring of NS shared-memory slots, completion signalled on
//    an mbarrier; out-of-range box parts are zero-filled by the TMA unit, which is exactly the
//    Dirichlet halo.  NS - R - 1 planes are always in flight ahead of the compute.
//  * The z-neighbours live in a per-thread register queue of 2R+1 float4 (4 consecutive x per
//    thread); x and y neighbours are read from the shared plane with 128-bit loads.
//  * u(t-1) is read and u(t+1) written in place with 128-bit global accesses; damping is
//    separable, three 1-D profiles, so it costs no HBM traffic.
//  Algorithmic HBM traffic: 16 B per grid-point update (u(t), u(t-1), m read; u(t+1) written).
#include "common.cuh"

// ---- PTX helpers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *tm, int c0, int c1,
                                            int c2, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(dst)),
        "l"((uint64_t)tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
        : "memory");
}

// ---- kernel ------------------------------------------------------------------------------
struct TmaArgs {
    Grid g;
    Coefs cf;
    float *up;
    const float *m;
    const float *dx, *dy, *dz;
    float sdamp;   // dt / irho
    float irho_c;
    int zchunk;    // planes per z-chunk
    Fuse f;
};

template <int R, int TXT, int TYT, int RY, int NS>
struct TmaCfg {
    static constexpr int HR = ((R + 3) / 4) * 4;  // x halo in shared memory (float4 aligned)
    static constexpr int TX = 4 * TXT, TY = TYT * RY;
    static constexpr int PX = TX + 2 * HR, PY = TY + 2 * R;
    static constexpr int SLOT = ((PX * PY * 4 + 127) / 128) * 128 / 4;  // floats per ring slot
    static constexpr int NQ = 2 * R + 1;
    static constexpr int NTHREADS = TXT * TYT;
    static constexpr size_t SMEM = (size_t)NS * SLOT * 4 + NS * 8 + 128;
};

__device__ __forceinline__ float f4get(const float4 &v, int e)
{
    return e == 0 ? v.x : (e == 1 ? v.y : (e == 2 ? v.z : v.w));
}

template <int R, int TXT, int TYT, int RY, int NS, int MINB>
__global__ void __launch_bounds__(TXT *TYT, MINB)
    acoustic3d_tma(const __grid_constant__ CUtensorMap tm, const TmaArgs a)
{
    typedef TmaCfg<R, TXT, TYT, RY, NS> C;
    constexpr int HR = C::HR, TX = C::TX, TY = C::TY, PX = C::PX, PY = C::PY, SLOT = C::SLOT;
    constexpr int NQ = C::NQ;
    extern __shared__ unsigned char smem_raw[];
    // 128-byte aligned ring (TMA destination alignment)
    float *ring = (float *)(smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u));
    uint64_t *bars = (uint64_t *)(ring + NS * SLOT);

    const Grid &g = a.g;
    const Coefs &cf = a.cf;
    const int tx = threadIdx.x % TXT, ty = threadIdx.x / TXT;
    const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
    const int z0 = blockIdx.z * a.zchunk;
    const int z1 = min(z0 + a.zchunk, g.n[2]);
    const int total = (z1 - z0) + 2 * R;  // planes this CTA consumes: z0-R .. z1-1+R
    const int bx = g.hx + x0 - HR, by = g.hy + y0 - R, bz = g.hz + z0 - R;
    constexpr uint32_t BOX_BYTES = PX * PY * 4;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; s++) mbar_init(&bars[s], 1);
        fence_barrier_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int npre = total < NS ? total : NS;
        for (int j = 0; j < npre; j++) {
            mbar_arrive_expect_tx(&bars[j], BOX_BYTES);
            tma_load_3d(ring + j * SLOT, &tm, bx, by, bz + j, &bars[j]);
        }
    }

    const int x = x0 + 4 * tx;
    const bool xact = x < g.n[0];
    const bool xfull = x + 3 < g.n[0];
    float4 q[RY][NQ];
#pragma unroll
    for (int j = 0; j < RY; j++)
#pragma unroll
        for (int k = 0; k < NQ; k++) q[j][k] = make_float4(0.f, 0.f, 0.f, 0.f);

    // damping profile along x for this thread's 4 points
    float sx[4], sy[RY];
    {
        float4 d4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (xact) d4 = *(const float4 *)(a.dx + x);  // profile arrays are padded 8 past n[0]
        sx[0] = d4.x; sx[1] = d4.y; sx[2] = d4.z; sx[3] = d4.w;
#pragma unroll
        for (int j = 0; j < RY; j++) {
            int y = y0 + ty * RY + j;
            sy[j] = y < g.n[1] ? a.dy[y] : 0.f;
        }
    }
    const Fuse &fz = a.f;
    const bool fuse_any = fz.snap[0] || fz.grad || fz.illum;
    const bool reg_vec = (fz.reg.lo[0] & 3) == 0;

    for (int jt = 0; jt < total; ++jt) {
        const int z = z0 + jt - 2 * R;  // output plane of this iteration (if jt >= 2R)
        const bool out = jt >= 2 * R;
        // ---- issue the point-wise global loads of the output plane early --------------------
        float4 upv[RY], mv[RY];
        long long gi[RY];
        bool act[RY];
#pragma unroll
        for (int j = 0; j < RY; j++) {
            int y = y0 + ty * RY + j;
            act[j] = out && xact && y < g.n[1];
            gi[j] = g.at(x, y, z);
            if (act[j]) {
                upv[j] = *(const float4 *)(a.up + gi[j]);
                mv[j] = __ldg((const float4 *)(a.m + gi[j]));
            }
        }
        // ---- wait for the incoming plane (z + R) and append it to the z-queue --------------
        mbar_wait(&bars[jt % NS], (uint32_t)((jt / NS) & 1));
        const float *tail = ring + (jt % NS) * SLOT;
#pragma unroll
        for (int j = 0; j < RY; j++) {
#pragma unroll
            for (int k = 0; k < NQ - 1; k++) q[j][k] = q[j][k + 1];
            q[j][NQ - 1] = *(const float4 *)(tail + (ty * RY + j + R) * PX + HR + 4 * tx);
        }
        if (out) {
            const float *cen = ring + ((jt - R) % NS) * SLOT;
            float lap[RY][4];
            // z part (register queue)
#pragma unroll
            for (int j = 0; j < RY; j++)
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    float acc = cf.c0 * f4get(q[j][R], e);
#pragma unroll
                    for (int k = 1; k <= R; k++)
                        acc += cf.cz[k] * (f4get(q[j][R + k], e) + f4get(q[j][R - k], e));
                    lap[j][e] = acc;
                }
            // x part (shared plane, row of the point)
#pragma unroll
            for (int j = 0; j < RY; j++) {
                const float *crow = cen + (ty * RY + j + R) * PX + 4 * tx;
                float w[4 + 2 * HR];
#pragma unroll
                for (int v = 0; v < 1 + 2 * (HR / 4); v++) {
                    float4 t = *(const float4 *)(crow + 4 * v);
                    w[4 * v] = t.x; w[4 * v + 1] = t.y; w[4 * v + 2] = t.z; w[4 * v + 3] = t.w;
                }
#pragma unroll
                for (int e = 0; e < 4; e++)
#pragma unroll
                    for (int k = 1; k <= R; k++)
                        lap[j][e] += cf.cx[k] * (w[HR + e + k] + w[HR + e - k]);
            }
            // y part (shared plane, column of the points; rows shared by the RY outputs)
#pragma unroll
            for (int rr = 0; rr < RY + 2 * R; rr++) {
                float4 t = *(const float4 *)(cen + (ty * RY + rr) * PX + HR + 4 * tx);
#pragma unroll
                for (int j = 0; j < RY; j++) {
                    const int k = rr - R - j;
                    if (k != 0 && k >= -R && k <= R) {
                        const float c = cf.cy[k < 0 ? -k : k];
                        lap[j][0] += c * t.x; lap[j][1] += c * t.y;
                        lap[j][2] += c * t.z; lap[j][3] += c * t.w;
                    }
                }
            }
            // ---- time update + fused extras ------------------------------------------------
            const float sz = a.dz[z];
#pragma unroll
            for (int j = 0; j < RY; j++) {
                if (!act[j]) continue;
                const int y = y0 + ty * RY + j;
                float uc[4] = {q[j][R].x, q[j][R].y, q[j][R].z, q[j][R].w};
                float up4[4] = {upv[j].x, upv[j].y, upv[j].z, upv[j].w};
                float m4[4] = {mv[j].x, mv[j].y, mv[j].z, mv[j].w};
                float un[4], delta[4];
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    float s = ((sx[e] + sy[j]) + sz) * a.sdamp;
                    float rden = __frcp_rn(m4[e] + s);
                    float d1 = uc[e] - up4[e];
                    delta[e] = (lap[j][e] - s * d1) * rden;
                    un[e] = (uc[e] + d1) + delta[e];
                }
                if (xfull) *(float4 *)(a.up + gi[j]) = make_float4(un[0], un[1], un[2], un[3]);
                else {
#pragma unroll
                    for (int e = 0; e < 4; e++)
                        if (x + e < g.n[0]) a.up[gi[j] + e] = un[e];
                }
                if (fuse_any && y >= fz.reg.lo[1] && y < fz.reg.hi[1] && z >= fz.reg.lo[2] &&
                    z < fz.reg.hi[2] && x + 3 >= fz.reg.lo[0] && x < fz.reg.hi[0]) {
                    const long long ri = fz.reg.at(x, y, z);
                    const bool vec = reg_vec && x >= fz.reg.lo[0] && x + 3 < fz.reg.hi[0];
                    if (vec) {
                        if (fz.snap[0])
                            *(float4 *)(fz.snap[0] + ri) = make_float4(uc[0], uc[1], uc[2], uc[3]);
                        if (fz.illum) {
                            float4 I = *(float4 *)(fz.illum + ri);
                            I.x += cf.dt * uc[0] * uc[0]; I.y += cf.dt * uc[1] * uc[1];
                            I.z += cf.dt * uc[2] * uc[2]; I.w += cf.dt * uc[3] * uc[3];
                            *(float4 *)(fz.illum + ri) = I;
                        }
                        if (fz.grad) {
                            const float gc = fz.gcoef * a.irho_c;
                            float4 us = __ldg((const float4 *)(fz.usnap[0] + ri));
                            float4 G = *(float4 *)(fz.grad + ri);
                            G.x -= gc * us.x * delta[0]; G.y -= gc * us.y * delta[1];
                            G.z -= gc * us.z * delta[2]; G.w -= gc * us.w * delta[3];
                            *(float4 *)(fz.grad + ri) = G;
                        }
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; e++) {
                            if (x + e < fz.reg.lo[0] || x + e >= fz.reg.hi[0]) continue;
                            if (fz.snap[0]) fz.snap[0][ri + e] = uc[e];
                            if (fz.illum) fz.illum[ri + e] += cf.dt * uc[e] * uc[e];
                            if (fz.grad)
                                fz.grad[ri + e] -= fz.gcoef * a.irho_c * fz.usnap[0][ri + e] * delta[e];
                        }
                    }
                }
            }
        }
        // ---- the centre plane is now dead: refill its slot NS planes ahead ------------------
        __syncthreads();
        if (threadIdx.x == 0) {
            const int jc = jt - R;
            if (jc >= 0 && jc + NS < total) {
                const int s = jc % NS;
                mbar_arrive_expect_tx(&bars[s], BOX_BYTES);
                tma_load_3d(ring + s * SLOT, &tm, bx, by, bz + jc + NS, &bars[s]);
            }
        }
    }
}

// ---- host side -------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                  const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct TileVariant { int txt, tyt, ry; };
static const TileVariant kVariants[] = {{16, 16, 1}, {32, 8, 2}, {16, 16, 2}, {32, 8, 1}};

static void variant_box(const judi_model *M, int &px, int &py)
{
    int v = M->ctx->opt_tile;
    if (v < 0 || v > 3 || M->r != 4) v = 0;
    int R = M->r, HR = ((R + 3) / 4) * 4;
    px = 4 * kVariants[v].txt + 2 * HR;
    py = kVariants[v].tyt * kVariants[v].ry + 2 * R;
}

bool tma3d_supported(const judi_model *M)
{
    return M->ndim == 3 && !M->tti && !M->irho_field && (M->r == 2 || M->r == 4 || M->r == 8) &&
           M->ctx->encode_tiled != nullptr;
}

void tma3d_encode(const judi_model *M, float *base, CUtensorMap *out)
{
    const Grid &g = M->g;
    int px, py;
    variant_box(M, px, py);
    cuuint64_t dims[3] = {(cuuint64_t)g.px, (cuuint64_t)g.py, (cuuint64_t)g.pz};
    cuuint64_t strides[2] = {(cuuint64_t)g.sy * 4, (cuuint64_t)g.sz * 4};
    cuuint32_t box[3] = {(cuuint32_t)px, (cuuint32_t)py, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult rc = ((EncodeTiledFn)M->ctx->encode_tiled)(
        out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides, box, estr,
        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) {
        char b[128];
        snprintf(b, sizeof(b), "cuTensorMapEncodeTiled failed with CUresult %d", (int)rc);
        throw JudiError(b);
    }
}

template <int R, int TXT, int TYT, int RY, int NS, int MINB>
static void launch_variant(const judi_model *M, const StepArgs &s, const CUtensorMap *tm)
{
    typedef TmaCfg<R, TXT, TYT, RY, NS> C;
    judi_ctx *ctx = M->ctx;
    auto kern = acoustic3d_tma<R, TXT, TYT, RY, NS, MINB>;
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
    a.m = M->m.p;
    a.dx = M->damp[0].p; a.dy = M->damp[1].p; a.dz = M->damp[2].p;
    a.sdamp = M->dt / M->irho_c;
    a.irho_c = M->irho_c;
    a.f = s.f;
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
    kern<<<grd, C::NTHREADS, C::SMEM, ctx->stream>>>(*tm, a);
    launch_count(ctx, "acoustic3d_tma");
}

void tma3d_step(const judi_model *M, const StepArgs &s)
{
    const CUtensorMap *tm = s.tmap;
    JUDI_REQUIRE(tm != nullptr, "tensor map missing");
    int v = M->ctx->opt_tile;
    if (M->r == 2) { launch_variant<2, 16, 16, 1, 6, 2>(M, s, tm); return; }
    if (M->r == 8) { launch_variant<8, 16, 16, 1, 12, 1>(M, s, tm); return; }
    switch (v) {
    case 1: launch_variant<4, 32, 8, 2, 8, 1>(M, s, tm); break;
    case 2: launch_variant<4, 16, 16, 2, 8, 1>(M, s, tm); break;
    case 3: launch_variant<4, 32, 8, 1, 8, 2>(M, s, tm); break;
    default: launch_variant<4, 16, 16, 1, 8, 2>(M, s, tm); break;
    }
}
