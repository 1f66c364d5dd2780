// model.cu -- physical model resident in HBM: padded grid, edge-replicated parameters, separable
// damping profiles, CFL time step, stencil coefficients.
//
// Replaces src/pysource/models.py of the reference: Model.__init__ :157-248 (grid + nbl :162-180),
// _gen_phys_param :301-326 (edge padding), damp_op / initialize_damp :54-120, padsizes :269-272,
// _cfl_coeff :440-456, _thomsen_scale :459-463, critical_dt :466-482.
#include "common.cuh"
#include <cmath>
#include <algorithm>

// ------------------------------------------------------------------------------------------
void *judi_ctx::alloc_bytes(size_t nbytes, bool zero)
{
    void *p = nullptr;
    if (nbytes == 0) nbytes = 16;
    CUDA_CHECK(cudaMallocFromPoolAsync(&p, nbytes, pool, stream));
    if (zero) CUDA_CHECK(cudaMemsetAsync(p, 0, nbytes, stream));
    return p;
}

void judi_ctx::free(void *p)
{
    if (p) cudaFreeAsync(p, stream);
}

float *judi_ctx::arena(size_t nfloat)
{
    if (nfloat <= arena_floats && arena_p) return arena_p;
    if (arena_p) {
        CUDA_CHECK(cudaStreamSynchronize(stream));
        CUDA_CHECK(cudaFree(arena_p));
        arena_p = nullptr;
        arena_floats = 0;
    }
    // plain cudaMalloc: outside the pool, so pool reuse patterns cannot split it
    const size_t bytes = std::max<size_t>(nfloat, 4) * sizeof(float);
    cudaError_t e = cudaMalloc((void **)&arena_p, bytes);
    if (e == cudaErrorMemoryAllocation) {
        // give the pool's cached blocks back to the driver and retry
        (void)cudaGetLastError();
        CUDA_CHECK(cudaStreamSynchronize(stream));
        CUDA_CHECK(cudaMemPoolTrimTo(pool, 0));
        e = cudaMalloc((void **)&arena_p, bytes);
    }
    if (e != cudaSuccess) arena_p = nullptr;
    CUDA_CHECK(e);
    arena_floats = nfloat;
    return arena_p;
}

void launch_count(judi_ctx *ctx, const char *what)
{
    ctx->launches++;
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) {
        cudaGetLastError();
        throw JudiError(std::string("kernel launch failed: ") + what + ": " + cudaGetErrorString(e));
    }
}

void judi_ctx::flush_timing()
{
    if (pending.empty()) return;
    CUDA_CHECK(cudaStreamSynchronize(stream));
    for (auto &pr : pending) {
        float ms = 0;
        CUDA_CHECK(cudaEventElapsedTime(&ms, pr.e0, pr.e1));
        stencil_ms += ms;
        kind_ms[pr.kind] += ms;
        kind_launches[pr.kind]++;
        free_events.push_back(pr.e0);
        free_events.push_back(pr.e1);
    }
    pending.clear();
}

// ------------------------------------------------------------------------------------------
Grid make_grid(int ndim, const int *N, int halo)
{
    Grid g;
    g.ndim = ndim;
    g.n[0] = N[0]; g.n[1] = N[1]; g.n[2] = N[2];
    g.hx = JUDI_HALO_XL;
    g.px = ((JUDI_HALO_XL + N[0] + halo + 31) / 32) * 32;
    if (ndim == 3) { g.hy = halo; g.py = N[1] + 2 * halo; }
    else { g.hy = 0; g.py = 1; }
    g.hz = halo;
    g.pz = N[2] + 2 * halo;
    g.sy = g.px;
    g.sz = (long long)g.px * g.py;
    g.off0 = g.hx + (long long)g.hy * g.sy + (long long)g.hz * g.sz;
    g.nalloc = g.sz * g.pz + 64;  // slack: vector loads of the last row may run a few floats over
    return g;
}

// Fornberg (1988): weights of derivative orders 0..m at z from nodes x[0..n-1]; c is (m+1) x n.
void fornberg_weights(double z, const double *x, int n, int m, double *c)
{
    double c1 = 1.0, c4 = x[0] - z;
    for (int i = 0; i < (m + 1) * n; i++) c[i] = 0.0;
    c[0] = 1.0;
    for (int i = 1; i < n; i++) {
        int mn = std::min(i, m);
        double c2 = 1.0, c5 = c4;
        c4 = x[i] - z;
        for (int j = 0; j < i; j++) {
            double c3 = x[i] - x[j];
            c2 *= c3;
            if (j == i - 1) {
                for (int k = mn; k >= 1; k--)
                    c[k * n + i] = c1 * (k * c[(k - 1) * n + i - 1] - c5 * c[k * n + i - 1]) / c2;
                c[i] = -c1 * c5 * c[i - 1] / c2;
            }
            for (int k = mn; k >= 1; k--)
                c[k * n + j] = (c4 * c[k * n + j] - k * c[(k - 1) * n + j]) / c3;
            c[j] = c4 * c[j] / c3;
        }
        c1 = c2;
    }
}

// ------------------------------------------------------------------------------------------
// dst (Grid layout, every allocated cell incl. halo) = src (physical n, x fastest) with the index
// clamped to the physical box: edge replication into the absorbing layer and the halo
// (models.py:321-322, initialize_function(..., pad)).
__global__ void pad_edge_kernel(Grid g, float *__restrict__ dst, const float *__restrict__ src,
                                int nx, int ny, int nz, int plx, int ply, int plz)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.nalloc) return;
    int ax = (int)(i % g.px);
    long long t = i / g.px;
    int ay = (int)(t % g.py);
    int az = (int)(t / g.py);
    int x = ax - g.hx - plx, y = ay - g.hy - ply, z = az - g.hz - plz;
    x = min(max(x, 0), nx - 1);
    y = min(max(y, 0), ny - 1);
    z = min(max(z, 0), nz - 1);
    dst[i] = src[x + (long long)nx * (y + (long long)ny * z)];
}

__global__ void fill_kernel(float *dst, long long n, float v)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = v;
}

// y = a*x + b elementwise (1 + 2*epsilon, 1/rho with a = 0 handled separately)
__global__ void affine_kernel(float *x, long long n, float a, float b)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = a * x[i] + b;
}

__global__ void recip_kernel(float *x, long long n)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = 1.0f / x[i];
}

// min / max over the domain points of a Grid field (positive floats: ordered as ints)
__global__ void minmax_kernel(Grid g, const float *__restrict__ f, float *out /* [min,max] */)
{
    long long n = (long long)g.n[0] * g.n[1] * g.n[2];
    float lo = INFINITY, hi = -INFINITY;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        int x = (int)(i % g.n[0]);
        long long t = i / g.n[0];
        int y = (int)(t % g.n[1]);
        int z = (int)(t / g.n[1]);
        float v = f[g.at(x, y, z)];
        lo = fminf(lo, v);
        hi = fmaxf(hi, v);
    }
    for (int o = 16; o > 0; o >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0) {
        // floats compare like signed ints when both are >= 0; handle negatives via CAS loop
        int *pi = (int *)out;
        int old = pi[0];
        while (__int_as_float(old) > lo) {
            int prev = atomicCAS(&pi[0], old, __float_as_int(lo));
            if (prev == old) break;
            old = prev;
        }
        old = pi[1];
        while (__int_as_float(old) < hi) {
            int prev = atomicCAS(&pi[1], old, __float_as_int(hi));
            if (prev == old) break;
            old = prev;
        }
    }
}

static void field_minmax(const judi_model *M, const float *f, float &lo, float &hi)
{
    judi_ctx *ctx = M->ctx;
    DevBuf out(ctx, 2, false);
    float init[2] = {INFINITY, -INFINITY};
    CUDA_CHECK(cudaMemcpyAsync(out.p, init, sizeof(init), cudaMemcpyHostToDevice, ctx->stream));
    minmax_kernel<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(M->g, f, out.p);
    launch_count(ctx, "minmax");
    float h[2];
    CUDA_CHECK(cudaMemcpyAsync(h, out.p, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    lo = h[0];
    hi = h[1];
}

// Upload one physical parameter into a Grid-layout field.  Scalars become constant fields when
// `force_field` (m, Thomsen parameters); otherwise the caller keeps the scalar.
__global__ void zero_surface_row_kernel(Grid g, float *__restrict__ f)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    if (x >= g.n[0] || y >= g.n[1]) return;
    f[g.at(x, y, 0)] = 0.f;
}

void model_update_dm_src(judi_model *M)
{
    if (!M->fs || !M->dm.p) { M->dm_src_buf = DevBuf(); return; }
    judi_ctx *ctx = M->ctx;
    if (!M->dm_src_buf.p) M->dm_src_buf = DevBuf(ctx, (size_t)M->g.nalloc, false);
    CUDA_CHECK(cudaMemcpyAsync(M->dm_src_buf.p, M->dm.p, (size_t)M->g.nalloc * sizeof(float),
                               cudaMemcpyDeviceToDevice, ctx->stream));
    dim3 blk(128, 1), grd((M->N[0] + 127) / 128, M->N[1]);
    zero_surface_row_kernel<<<grd, blk, 0, ctx->stream>>>(M->g, M->dm_src_buf.p);
    launch_count(ctx, "zero_surface_row");
}

void model_set_param(judi_model *M, DevBuf &dst, const judi_param &p, int on_device,
                     float default_value, bool force_field)
{
    judi_ctx *ctx = M->ctx;
    const Grid &g = M->g;
    if (p.kind == JUDI_PARAM_ARRAY) {
        JUDI_REQUIRE(p.data != nullptr, "array parameter without data");
        size_t nphys = (size_t)M->n[0] * M->n[1] * M->n[2];
        if (!dst.p) dst = DevBuf(ctx, (size_t)g.nalloc, false);
        const float *src = p.data;
        DevBuf tmp;
        if (!on_device) {
            tmp = DevBuf(ctx, nphys, false);
            CUDA_CHECK(cudaMemcpyAsync(tmp.p, p.data, nphys * sizeof(float),
                                       cudaMemcpyHostToDevice, ctx->stream));
            src = tmp.p;
        }
        int nb = (int)((g.nalloc + 255) / 256);
        pad_edge_kernel<<<nb, 256, 0, ctx->stream>>>(g, dst.p, src, M->n[0], M->n[1], M->n[2],
                                                     M->padl[0], M->padl[1], M->padl[2]);
        launch_count(ctx, "pad_edge");
        if (!on_device) CUDA_CHECK(cudaStreamSynchronize(ctx->stream));  // host buffer borrowed
    } else if (force_field) {
        float v = p.kind == JUDI_PARAM_SCALAR ? p.value : default_value;
        if (!dst.p) dst = DevBuf(ctx, (size_t)g.nalloc, false);
        int nb = (int)((g.nalloc + 255) / 256);
        fill_kernel<<<nb, 256, 0, ctx->stream>>>(dst.p, g.nalloc, v);
        launch_count(ctx, "fill");
    }
}

// One side of the damping profile (models.py:74-94), outermost cell first.
static std::vector<float> damp_side(int npad, float h)
{
    int nb = npad - 3;  // "3 point buffer to avoid weird interaction with abc"
    std::vector<float> v;
    if (nb <= 0) return v;
    v.resize(nb);
    float coeff = (float)(1.5 * std::log(1.0 / 0.001) / nb);
    for (int i = 0; i < nb; i++) {
        float pos = fabsf(((float)nb - (float)i + 1.0f) / (float)nb);
        float val = coeff * (pos - sinf((float)(2 * M_PI) * pos) / (float)(2 * M_PI));
        v[i] = val / h;
    }
    return v;
}

static Region make_region(const int lo[3], const int hi[3])
{
    Region r;
    for (int d = 0; d < 3; d++) { r.lo[d] = lo[d]; r.hi[d] = hi[d]; r.ext[d] = hi[d] - lo[d]; }
    r.pitch = ((r.ext[0] + 3) / 4) * 4;
    r.slot = (long long)r.pitch * r.ext[1] * r.ext[2];
    return r;
}

// symmetry axis of the TTI medium = third row of rot_y(theta) rot_z(phi) (FD_utils.py:24-47)
__global__ void tti_axis_kernel(const float *__restrict__ theta, const float *__restrict__ phi,
                                long long n, float *__restrict__ rx, float *__restrict__ ry,
                                float *__restrict__ rz)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float st, ct, sp = 0.f, cp = 1.f;
    sincosf(theta[i], &st, &ct);
    if (phi) sincosf(phi[i], &sp, &cp);   // 2-D: no azimuth, r = (sin theta, 0, cos theta)
    rx[i] = st * cp;
    ry[i] = st * sp;
    rz[i] = ct;
}

// point-wise TTI coefficients of the closed-form rotation (flux3d.cu), constant density b:
// the same fp32 expressions the two-pass kernel evaluates per node and step
// (bf: b as a field, or nullptr for the constant; mi: optional output m * b)
__global__ void tti_abc_kernel(const float *__restrict__ eps, const float *__restrict__ delta, float bc,
                               const float *__restrict__ bf, const float *__restrict__ m,
                               long long n, float *__restrict__ A0, float *__restrict__ B0,
                               float *__restrict__ C0, float *__restrict__ mi)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float b = bf ? bf[i] : bc;
    if (mi) { mi[i] = __fmul_rn(m[i], b); return; }
    const float e = eps[i], d = delta[i];
    const float emd = __fsub_rn(e, d);
    A0[i] = __fmul_rn(b, d);
    B0[i] = __fmul_rn(b, sqrtf(__fmul_rn(emd, d)));
    C0[i] = __fmul_rn(b, emd);
}

// relaxation parameters of the SLS model (kernels.py:107-116), in the written expression order
__global__ void visco_params_kernel(const float *__restrict__ qp, float f0, long long n,
                                    float *__restrict__ tt, float *__restrict__ ts)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float q = qp[i];
    const float t_s = __fdiv_rn(__fsub_rn(sqrtf(__fadd_rn(1.f, __fdiv_rn(1.f, __fmul_rn(q, q)))), __fdiv_rn(1.f, q)), f0);
    const float t_ep = __fdiv_rn(1.f, __fmul_rn(__fmul_rn(f0, f0), t_s));
    ts[i] = t_s;
    tt[i] = __fsub_rn(__fdiv_rn(t_ep, t_s), 1.f);
}

// Visco-acoustic model (models.py:208-214): quality factor field qp (NULL keeps the current one) and
// the peak frequency f0 the reference passes with every call (propagators.py:46).
void model_set_visco(judi_model *M, const judi_param *qp, float f0, int on_device)
{
    judi_ctx *ctx = M->ctx;
    JUDI_REQUIRE(!M->tti, "a model cannot be TTI and visco-acoustic");
    JUDI_REQUIRE(f0 > 0.f, "f0 must be positive");
    if (qp && qp->kind != JUDI_PARAM_NONE) model_set_param(M, M->qp, *qp, on_device, 1.0f, true);
    JUDI_REQUIRE(M->qp.p != nullptr, "visco-acoustic model without qp");
    M->visco = true;
    M->f0 = f0;
    if (!M->vtt.p) {
        M->vtt = DevBuf(ctx, (size_t)M->g.nalloc, false);
        M->vts = DevBuf(ctx, (size_t)M->g.nalloc, false);
    }
    int nb = (int)((M->g.nalloc + 255) / 256);
    visco_params_kernel<<<nb, 256, 0, ctx->stream>>>(M->qp.p, f0, M->g.nalloc, M->vtt.p, M->vts.p);
    launch_count(ctx, "visco_params");
}

void model_build(judi_model *M, judi_ctx *ctx, const judi_model_desc *desc)
{
    JUDI_REQUIRE(desc->ndim == 2 || desc->ndim == 3, "ndim must be 2 or 3");
    JUDI_REQUIRE(desc->space_order == 4 || desc->space_order == 8 || desc->space_order == 16 ||
                     desc->space_order == 2 || desc->space_order == 12,
                 "space_order must be one of 2, 4, 8, 12, 16");
    JUDI_REQUIRE(desc->m.kind != JUDI_PARAM_NONE, "squared slowness m is required");
    JUDI_REQUIRE(desc->nbl >= 0, "nbl must be >= 0");
    M->ctx = ctx;
    M->ndim = desc->ndim;
    M->so = desc->space_order;
    M->r = M->so / 2;
    M->nbl = desc->nbl;
    M->fs = desc->fs;
    M->dt_user = desc->dt;
    const int dm2[2] = {0, 2};
    for (int k = 0; k < desc->ndim; k++) {
        int gd = desc->ndim == 3 ? k : dm2[k];
        JUDI_REQUIRE(desc->shape[k] >= 1, "shape must be positive");
        M->n[gd] = desc->shape[k];
        M->d[gd] = desc->spacing[k];
        M->o[gd] = desc->origin[k];
        M->padl[gd] = M->padr[gd] = desc->nbl;
        // models.py:172: origin_pml = dtype(o - s*nbl), evaluated in double from the fp32 inputs
        M->opml[gd] = (float)((double)desc->origin[k] - (double)desc->spacing[k] * desc->nbl);
        if (desc->fs && gd == 2) {
            // free surface (models.py:174-176, 269-272): no layer above z = 0, origin unshifted
            M->padl[gd] = 0;
            M->opml[gd] = desc->origin[k];
        }
    }
    for (int d = 0; d < 3; d++) M->N[d] = M->n[d] + M->padl[d] + M->padr[d];
    // the two-pass flux form (TTI / variable density) and the isic / fwi Born sources of every
    // physics differentiate twice: halo of 2R
    M->g = make_grid(M->ndim, M->N, 2 * M->r > JUDI_HALO ? 2 * M->r : JUDI_HALO);
    {
        int lo[3] = {0, 0, 0};
        M->full = make_region(lo, M->N);
        int plo[3] = {M->padl[0], M->padl[1], M->padl[2]};
        int phi[3] = {M->padl[0] + M->n[0], M->padl[1] + M->n[1], M->padl[2] + M->n[2]};
        M->phys = make_region(plo, phi);
    }

    // ---- parameters --------------------------------------------------------------------
    int dev = desc->params_on_device;
    model_set_param(M, M->m, desc->m, dev, 1.0f, true);
    // density (models.py:250-266): scalar -> Constant irho, array -> field irho = 1/rho
    M->irho_c = 1.0f;
    M->irho_field = false;
    if (desc->rho.kind == JUDI_PARAM_SCALAR) {
        M->irho_c = 1.0f / desc->rho.value;
    } else if (desc->rho.kind == JUDI_PARAM_ARRAY) {
        model_set_param(M, M->irho, desc->rho, dev, 1.0f, true);
        int nb = (int)((M->g.nalloc + 255) / 256);
        recip_kernel<<<nb, 256, 0, ctx->stream>>>(M->irho.p, M->g.nalloc);
        launch_count(ctx, "recip");
        M->irho_field = true;
    } else if (desc->b.kind == JUDI_PARAM_SCALAR) {
        M->irho_c = desc->b.value;
    } else if (desc->b.kind == JUDI_PARAM_ARRAY) {
        model_set_param(M, M->irho, desc->b, dev, 1.0f, true);
        M->irho_field = true;
    }
    M->tti = desc->epsilon.kind != JUDI_PARAM_NONE || desc->delta.kind != JUDI_PARAM_NONE ||
             desc->theta.kind != JUDI_PARAM_NONE || desc->phi.kind != JUDI_PARAM_NONE;
    float eps_max = 1.0f;
    if (M->tti) {
        // models.py:216-225: epsilon <- 1 + 2 epsilon, delta <- 1 + 2 delta
        int nb = (int)((M->g.nalloc + 255) / 256);
        model_set_param(M, M->eps, desc->epsilon, dev, 0.0f, true);
        affine_kernel<<<nb, 256, 0, ctx->stream>>>(M->eps.p, M->g.nalloc, 2.0f, 1.0f);
        launch_count(ctx, "affine");
        model_set_param(M, M->delta, desc->delta, dev, 0.0f, true);
        affine_kernel<<<nb, 256, 0, ctx->stream>>>(M->delta.p, M->g.nalloc, 2.0f, 1.0f);
        launch_count(ctx, "affine");
        model_set_param(M, M->theta, desc->theta, dev, 0.0f, true);
        if (M->ndim == 3) model_set_param(M, M->phi, desc->phi, dev, 0.0f, true);
        {
            for (int d = 0; d < 3; d++) M->r3[d] = DevBuf(ctx, (size_t)M->g.nalloc, true);
            tti_axis_kernel<<<nb, 256, 0, ctx->stream>>>(M->theta.p, M->ndim == 3 ? M->phi.p : nullptr,
                                                         M->g.nalloc, M->r3[0].p, M->r3[1].p, M->r3[2].p);
            launch_count(ctx, "tti_axis");
        }
        if (M->ndim == 3) {
            if (M->so == 8 || M->so == 4) {
                for (int d = 0; d < 3; d++) M->ttiABC[d] = DevBuf(ctx, (size_t)M->g.nalloc, false);
                tti_abc_kernel<<<nb, 256, 0, ctx->stream>>>(M->eps.p, M->delta.p, M->irho_c,
                                                            M->irho_field ? M->irho.p : nullptr, M->m.p, M->g.nalloc,
                                                            M->ttiABC[0].p, M->ttiABC[1].p, M->ttiABC[2].p,
                                                            nullptr);
                launch_count(ctx, "tti_abc");
                if (M->irho_field) {
                    // the single-pass kernel's tile ring carries m * b for a density field
                    M->tti_mi = DevBuf(ctx, (size_t)M->g.nalloc, false);
                    tti_abc_kernel<<<nb, 256, 0, ctx->stream>>>(M->eps.p, M->delta.p, M->irho_c, M->irho.p, M->m.p,
                                                                M->g.nalloc, nullptr, nullptr, nullptr, M->tti_mi.p);
                    launch_count(ctx, "tti_abc");
                }
            }
        }
        if (M->tti && !M->irho_field) {
            // the flux-form kernels read irho as a field
            M->irho = DevBuf(ctx, (size_t)M->g.nalloc, false);
            fill_kernel<<<nb, 256, 0, ctx->stream>>>(M->irho.p, M->g.nalloc, M->irho_c);
            launch_count(ctx, "fill");
        }
        float lo, hi;
        field_minmax(M, M->eps.p, lo, hi);
        eps_max = hi;
    }
    M->has_dm = false;
    if (desc->dm.kind == JUDI_PARAM_ARRAY ||
        (desc->dm.kind == JUDI_PARAM_SCALAR && desc->dm.value != 0.0f)) {
        model_set_param(M, M->dm, desc->dm, dev, 0.0f, true);
        M->has_dm = true;
        model_update_dm_src(M);
    }

    // ---- damping profiles (separable: damp = sum_d profile_d, models.py:54-96) -------------
    for (int d = 0; d < 3; d++) {
        std::vector<float> prof(M->N[d] + 8, 0.0f);
        if (!(M->ndim == 2 && d == 1)) {
            std::vector<float> l = damp_side(M->padl[d], M->d[d]);
            std::vector<float> r = damp_side(M->padr[d], M->d[d]);
            for (size_t i = 0; i < l.size(); i++) prof[i] += l[i];
            for (size_t i = 0; i < r.size(); i++) prof[M->N[d] - 1 - i] += r[i];
        }
        M->damp_h[d] = prof;
        M->damp[d] = DevBuf(ctx, prof.size(), false);
        M->damp[d].upload(prof);
    }

    // ---- critical dt (models.py:440-482) ------------------------------------------------
    float mmin, mmax;
    field_minmax(M, M->m.p, mmin, mmax);
    JUDI_REQUIRE(mmin > 0.0f, "squared slowness must be positive");
    {
        int so = std::max(M->so / 2, 4);
        int npts = 2 * so;  // range(-so, so)
        std::vector<double> x(npts), c(3 * npts);
        for (int i = 0; i < npts; i++) x[i] = i - so;
        fornberg_weights(0.0, x.data(), npts, 2, c.data());
        double sabs = 0;
        for (int i = 0; i < npts; i++) sabs += std::fabs(c[2 * npts + i]);
        double cfl = 0.9 * std::sqrt(4.0 / ((double)M->ndim * sabs));
        float vmax = sqrtf(1.0f / mmin);
        float scale = M->tti ? sqrtf(1.0f + 2.0f * eps_max) : 1.0f;
        float hmin = M->d[0];
        if (M->ndim == 3) hmin = std::min(hmin, M->d[1]);
        hmin = std::min(hmin, M->d[2]);
        double dtc = cfl * (double)hmin / (double)(scale * vmax);
        char buf[64];
        snprintf(buf, sizeof(buf), "%.3e", dtc);
        float dt = strtof(buf, nullptr);
        if (M->dt_user > 0 && !(M->dt_user > dt)) {
            snprintf(buf, sizeof(buf), "%.3e", (double)M->dt_user);
            dt = strtof(buf, nullptr);
        }
        M->dt = dt;
    }

    // ---- stencil coefficients ---------------------------------------------------------------
    {
        Coefs &cf = M->cf;
        memset(&cf, 0, sizeof(cf));
        int r = M->r;
        cf.r = r;
        double dt = (double)M->dt;
        cf.dt = M->dt;
        cf.dt2 = (float)(dt * dt);
        cf.idt = (float)(1.0 / dt);
        cf.idt2 = (float)(1.0 / (dt * dt));
        int npts = 2 * r + 1;
        std::vector<double> x(npts), c(3 * npts);
        for (int i = 0; i < npts; i++) x[i] = i - r;
        fornberg_weights(0.0, x.data(), npts, 2, c.data());
        int ns = 2 * r;
        std::vector<double> xs(ns), cs(2 * ns);
        for (int i = 0; i < ns; i++) xs[i] = (i - r) + 0.5;
        fornberg_weights(0.0, xs.data(), ns, 1, cs.data());
        double c0 = 0;
        for (int d = 0; d < 3; d++) {
            if (M->ndim == 2 && d == 1) continue;
            double h = (double)M->d[d];
            float *cc = d == 0 ? cf.cx : (d == 1 ? cf.cy : cf.cz);
            float *ww = d == 0 ? cf.wx : (d == 1 ? cf.wy : cf.wz);
            c0 += dt * dt * c[2 * npts + r] / (h * h);
            for (int k = 1; k <= r; k++) {
                cc[k] = (float)(dt * dt * c[2 * npts + r + k] / (h * h));
                ww[k] = (float)(cs[ns + r + k - 1] / h);
            }
        }
        cf.c0 = (float)c0;
    }
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
}
