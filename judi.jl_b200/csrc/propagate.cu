// propagate.cu -- per-shot time loops: the schedule of the reference's three Devito operators.
//
//   forward  (operators.py:81-135):  pde ; inject src[t] -> level t+1 ; rec[t] = interp(level t) ;
//                                    save level t ; illumination.     t = 1 .. nt-2
//   adjoint  (same operator, fw=False): mirror image, t = nt-2 .. 1, writes level t-1
//   born     (operators.py:138-192): pde(u) ; inject ; interp(u) [nlind] ; interp(ul) ;
//                                    pde(ul ; q = lin_src(u)) ; save ; illumination
//   gradient (operators.py:195-240): pde(v, backward) ; inject residual[t] -> v[t-1] ;
//                                    gradm -= w * IC(u[t], v) ; illumination
// Loop bounds and ordering follow SURVEY.md Appendix A.2 / A.6; J_adjoint composition follows
// src/pysource/interface.py:411-470 (standard) and :473-562 (checkpointing, here as fixed-stride
// HBM checkpoints + segment recompute instead of pyrevolve).
#include "common.cuh"
#include "propagate.h"
#include <algorithm>
#include <chrono>
#include <cmath>

// phase timing of one call (option "profile"): host clock after a stream synchronize
struct PhaseClock {
    judi_ctx *ctx;
    std::chrono::steady_clock::time_point t0;
    explicit PhaseClock(judi_ctx *c) : ctx(c) { if (ctx->opt_profile) mark(nullptr); }
    void mark(const char *what)
    {
        if (!ctx->opt_profile) return;
        cudaStreamSynchronize(ctx->stream);
        auto t1 = std::chrono::steady_clock::now();
        if (what)
            fprintf(stderr, "[judi_b200] %-28s %9.3f ms\n", what,
                    std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

// ---- small data-movement kernels -------------------------------------------------------------
// out[c][r] = in[r][c]
__global__ void transpose_kernel(const float *__restrict__ in, float *__restrict__ out, int rows,
                                 int cols)
{
    __shared__ float tile[32][33];
    int c = blockIdx.x * 32 + threadIdx.x, r0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        int r = r0 + j;
        if (r < rows && c < cols) tile[j][threadIdx.x] = in[(size_t)r * cols + c];
    }
    __syncthreads();
    int r = r0 + threadIdx.x, c0 = blockIdx.x * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        int cc = c0 + j;
        if (r < rows && cc < cols) out[(size_t)cc * rows + r] = tile[threadIdx.x][j];
    }
}

static void transpose(judi_ctx *ctx, const float *in, float *out, int rows, int cols)
{
    if (rows == 0 || cols == 0) return;
    dim3 blk(32, 8), grd((cols + 31) / 32, (rows + 31) / 32);
    transpose_kernel<<<grd, blk, 0, ctx->stream>>>(in, out, rows, cols);
    launch_count(ctx, "transpose");
}

void Traces::alloc(judi_ctx *ctx, int nt_, int np_)
{
    nt = nt_; np = np_;
    buf = DevBuf(ctx, (size_t)nt * std::max(np, 1), true);
}

// user layout is time fastest (data[p*nt + t]); device rows are trace fastest
void Traces::load(judi_ctx *ctx, const float *user, int nt_, int np_, bool on_device)
{
    alloc(ctx, nt_, np_);
    size_t n = (size_t)nt * np;
    if (n == 0) return;
    if (on_device) {
        transpose(ctx, user, buf.p, np, nt);
    } else {
        DevBuf tmp(ctx, n, false);
        CUDA_CHECK(cudaMemcpyAsync(tmp.p, user, n * sizeof(float), cudaMemcpyHostToDevice,
                                   ctx->stream));
        transpose(ctx, tmp.p, buf.p, np, nt);
        // no stream sync: the staging buffer is freed in stream order, and a copy from pageable
        // host memory has been staged by the time cudaMemcpyAsync returns; page-locked caller
        // arrays must stay untouched until the call returns (every run_* ends with a sync)
    }
}

void Traces::store(judi_ctx *ctx, float *user, bool on_device) const
{
    size_t n = (size_t)nt * np;
    if (n == 0) return;
    if (on_device) {
        transpose(ctx, buf.p, user, nt, np);
    } else {
        DevBuf tmp(ctx, n, false);
        transpose(ctx, buf.p, tmp.p, nt, np);
        CUDA_CHECK(cudaMemcpyAsync(user, tmp.p, n * sizeof(float), cudaMemcpyDeviceToHost,
                                   ctx->stream));
        // completed by the stream synchronisation that ends every run_* call (or by the caller
        // before it reads `user` on the host)
    }
}

// region layout -> padded-grid compact array (x fastest), zero outside the region
__global__ void region_to_padded_kernel(Region reg, const float *__restrict__ src, int nx, int ny,
                                        int nz, float *__restrict__ dst, int accumulate)
{
    long long n = (long long)nx * ny * nz;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int x = (int)(i % nx);
    long long t = i / nx;
    int y = (int)(t % ny), z = (int)(t / ny);
    float v = reg.has(x, y, z) ? src[reg.at(x, y, z)] : 0.f;
    if (accumulate) dst[i] += v;
    else dst[i] = v;
}

void region_output(const judi_model *M, const Region &reg, const float *src, float *user,
                   bool on_device, bool accumulate)
{
    judi_ctx *ctx = M->ctx;
    long long n = (long long)M->N[0] * M->N[1] * M->N[2];
    int nb = (int)((n + 255) / 256);
    if (on_device) {
        region_to_padded_kernel<<<nb, 256, 0, ctx->stream>>>(reg, src, M->N[0], M->N[1], M->N[2],
                                                             user, accumulate ? 1 : 0);
        launch_count(ctx, "region_to_padded");
        return;
    }
    DevBuf tmp(ctx, (size_t)n, false);
    if (accumulate)
        CUDA_CHECK(cudaMemcpyAsync(tmp.p, user, n * sizeof(float), cudaMemcpyHostToDevice,
                                   ctx->stream));
    region_to_padded_kernel<<<nb, 256, 0, ctx->stream>>>(reg, src, M->N[0], M->N[1], M->N[2],
                                                         tmp.p, accumulate ? 1 : 0);
    launch_count(ctx, "region_to_padded");
    CUDA_CHECK(cudaMemcpyAsync(user, tmp.p, n * sizeof(float), cudaMemcpyDeviceToHost,
                               ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
}

// Grid-layout field -> padded-grid compact array
__global__ void grid_to_padded_kernel(Grid g, const float *__restrict__ src,
                                      float *__restrict__ dst)
{
    long long n = (long long)g.n[0] * g.n[1] * g.n[2];
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int x = (int)(i % g.n[0]);
    long long t = i / g.n[0];
    int y = (int)(t % g.n[1]), z = (int)(t / g.n[1]);
    dst[i] = src[g.at(x, y, z)];
}

void grid_output(const judi_model *M, const float *field, float *user, bool on_device)
{
    judi_ctx *ctx = M->ctx;
    long long n = (long long)M->N[0] * M->N[1] * M->N[2];
    int nb = (int)((n + 255) / 256);
    if (on_device) {
        grid_to_padded_kernel<<<nb, 256, 0, ctx->stream>>>(M->g, field, user);
        launch_count(ctx, "grid_to_padded");
        return;
    }
    DevBuf tmp(ctx, (size_t)n, false);
    grid_to_padded_kernel<<<nb, 256, 0, ctx->stream>>>(M->g, field, tmp.p);
    launch_count(ctx, "grid_to_padded");
    CUDA_CHECK(cudaMemcpyAsync(user, tmp.p, n * sizeof(float), cudaMemcpyDeviceToHost,
                               ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
}

// residual = a + b - obs (b optional), f += 0.5 * dt * sum(residual^2)    (sensitivity.py:264-276)
__global__ void loss_kernel(long long n, float *__restrict__ a, const float *__restrict__ b,
                            const float *__restrict__ obs, int is_residual, double *fout,
                            float half_dt)
{
    double acc = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        float r;
        if (is_residual) r = obs[i];
        else r = b ? a[i] - (obs[i] - b[i]) : a[i] - obs[i];
        a[i] = r;
        acc += (double)r * (double)r;
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    __shared__ double sh[8];
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) s += sh[w];
        atomicAdd(fout, s * (double)half_dt);
    }
}

// *dst (+)= the objective value: from the device scalar of the loss kernel, or the host value of a
// misfit callback
__global__ void scalar_out_kernel(float *dst, const double *src_dev, double host_val, int accumulate)
{
    const float v = (float)(src_dev ? *src_dev : host_val);
    *dst = accumulate ? *dst + v : v;
}

// Free surface (fields_exprs.py:106-137, applied to the new level after the injection, see the
// `extra` position in operators.py:131,188,231): u(-k) = -u(k-1) for k = 1..so/2, then u(0) = 0.
__global__ void free_surface_kernel(Grid g, float *__restrict__ u, int r)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    if (x >= g.n[0] || y >= g.n[1]) return;
    long long i0 = g.at(x, y, 0);
    for (int k = 1; k <= r; k++) u[i0 - k * g.sz] = -u[i0 + (k - 1) * g.sz];
    u[i0] = 0.f;
}

static void free_surface(const judi_model *M, float *u)
{
    dim3 blk(128, 1), grd((M->N[0] + 127) / 128, M->N[1]);
    free_surface_kernel<<<grd, blk, 0, M->ctx->stream>>>(M->g, u, M->r);
    launch_count(M->ctx, "free_surface");
}

// padded-grid compact array -> Grid-layout field (halo untouched, must be zero already)
__global__ void padded_to_grid_kernel(Grid g, const float *__restrict__ src, float *__restrict__ dst)
{
    long long n = (long long)g.n[0] * g.n[1] * g.n[2];
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int x = (int)(i % g.n[0]);
    long long t = i / g.n[0];
    int y = (int)(t % g.n[1]), z = (int)(t / g.n[1]);
    dst[g.at(x, y, z)] = src[i];
}

static void padded_to_grid(const judi_model *M, const float *src_dev, float *dst)
{
    long long n = (long long)M->N[0] * M->N[1] * M->N[2];
    padded_to_grid_kernel<<<(int)((n + 255) / 256), 256, 0, M->ctx->stream>>>(M->g, src_dev, dst);
    launch_count(M->ctx, "padded_to_grid");
}

// ---- Shot ------------------------------------------------------------------------------------
Shot::Shot(judi_model *M_) : M(M_), ctx(M_->ctx) {}

static std::vector<float> first_column(judi_ctx *ctx, const float *user, int nt, bool on_device)
{
    std::vector<float> h(nt);
    if (on_device) {
        CUDA_CHECK(cudaMemcpyAsync(h.data(), user, nt * sizeof(float), cudaMemcpyDeviceToHost,
                                   ctx->stream));
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    } else memcpy(h.data(), user, nt * sizeof(float));
    return h;
}

// Upload the space-time source / extended receiver description of this call.
void Shot::bind_ext(const ExtSpec *ext, const float *wavelet, int nt, bool on_device)
{
    if (!ext) return;
    size_t ng = (size_t)M->N[0] * M->N[1] * M->N[2];
    auto to_dev = [&](const float *user, size_t n, DevBuf &owned) -> const float * {
        if (on_device) return user;
        owned = DevBuf(ctx, n, false);
        CUDA_CHECK(cudaMemcpyAsync(owned.p, user, n * sizeof(float), cudaMemcpyHostToDevice,
                                   ctx->stream));
        return owned.p;
    };
    if (ext->ws) {
        JUDI_REQUIRE(wavelet != nullptr, "extended source needs a wavelet");
        DevBuf tmp;
        const float *d = to_dev(ext->ws, ng, tmp);
        ws_grid = DevBuf(ctx, (size_t)M->g.nalloc, true);
        padded_to_grid(M, d, ws_grid.p);
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        wt_host = first_column(ctx, wavelet, nt, on_device);
        wt_dev = DevBuf(ctx, (size_t)nt, false);
        wt_dev.upload(wt_host);
    }
    if (ext->wr_wavelet) {
        JUDI_REQUIRE(ext->w_out != nullptr, "extended receiver needs an output array");
        wrec = DevBuf(ctx, (size_t)M->full.slot, true);
        wrt_host = first_column(ctx, ext->wr_wavelet, nt, on_device);
        wrt_dev = DevBuf(ctx, (size_t)nt, false);
        wrt_dev.upload(wrt_host);
    }
    if (ext->qwf) {
        qwf_dev = to_dev(ext->qwf, ng * (size_t)nt, qwf_owned);
        qgrid = DevBuf(ctx, (size_t)M->g.nalloc, true);
    }
}

void Shot::apply_ext(StepArgs &s, int t)
{
    const bool flux = M->tti || M->irho_field || M->visco;
    const float base = flux ? M->cf.dt2 : M->cf.dt2 / M->irho_c;  // see the update formula
    if (ws_grid.p) {
        s.qext = ws_grid.p;
        s.qext_scale = base * wt_host[t];
    } else if (qwf_dev) {
        size_t ng = (size_t)M->N[0] * M->N[1] * M->N[2];
        padded_to_grid(M, qwf_dev + (size_t)t * ng, qgrid.p);
        s.qext = qgrid.p;
        s.qext_scale = base;
    }
    if (wrec.p) {
        if (!(s.f.snap[0] || s.f.grad || s.f.illum)) s.f.reg = M->full;
        s.f.wrec = wrec.p;
        s.f.wrec_scale = M->dt * wrt_host[t];
    }
}

// The same description for the persistent 2-D kernel (constant-density acoustic only).
bool persist2d_bind_ext(const judi_model *M, const Shot &sh, P2Fuse &fz)
{
    if (!sh.ext_active()) return true;
    if (M->tti || M->irho_field || M->visco) return false;
    fz.qbase = M->cf.dt2 / M->irho_c;
    if (sh.ws_grid.p) { fz.qext = sh.ws_grid.p; fz.qwt = sh.wt_dev.p; }
    else if (sh.qwf_dev) fz.qwf = sh.qwf_dev;
    if (sh.wrec.p) {
        fz.wrec = sh.wrec.p;
        fz.wrt = sh.wrt_dev.p;
        fz.wdt = M->dt;
        if (!(fz.snap || fz.grad || fz.illum)) fz.reg = M->full;
        else if (!(fz.reg.slot == M->full.slot)) return false;   // wrec is laid out on the full region
    }
    return true;
}

void Shot::finish_ext(const ExtSpec *ext, bool on_device)
{
    if (ext && ext->w_out && wrec.p) region_output(M, M->full, wrec.p, ext->w_out, on_device, false);
}

void Shot::make_args(StepArgs &s, Wavefield &W)
{
    memset(&s, 0, sizeof(s));
    for (int f = 0; f < W.nf; f++) { s.u[f] = W.level(f); s.up[f] = W.other(f); }
    s.tmap = W.has_tmap ? &W.tmap_halo[0][W.cur] : nullptr;
    s.tmap_prev = W.has_tmap ? &W.tmap_tile[0][1 - W.cur] : nullptr;
    s.tmap_m = W.has_tmap ? &W.tmap_m : nullptr;
    s.scratch = &W;
    s.adj = adjoint ? 1 : 0;
}

// One forward (dir=+1) or backward (dir=-1) step of the background field at time index t.
void Shot::step_plain(Wavefield &W, int t, const Fuse *fuse, const SparseSet *S_in,
                      const Traces *in, const SparseCorr *corr, const SparseSet *S_out,
                      Traces *out)
{
    StepArgs s;
    make_args(s, W);
    if (fuse) s.f = *fuse;
    if (ext_active()) apply_ext(s, t);
    stencil_step(M, s);
    sparse_step(M, S_in, in ? in->row(t) : nullptr, W.other(0), corr, S_out, W.level(0),
                out ? out->row(t) : nullptr, nullptr, nullptr, nullptr);
    if (M->fs && !M->visco)   // SLS_2nd_order carries no mirror equations (kernels.py:141)
        for (int f = 0; f < W.nf; f++) free_surface(M, W.other(f));
    W.swap();
}

void Shot::step_born(Wavefield &W, Wavefield &WL, int t, int ic, const Fuse *fuse,
                     const SparseSet *S_src, const Traces *q, const SparseSet *S_rec,
                     Traces *recl, Traces *recnl)
{
    StepArgs s;
    make_args(s, W);
    s.born = 1;
    s.born_ic = ic;
    for (int f = 0; f < WL.nf; f++) { s.ul[f] = WL.level(f); s.ulp[f] = WL.other(f); }
    s.lin = &WL;
    if (fuse) s.f = *fuse;
    if (ext_active()) apply_ext(s, t);
    stencil_step(M, s);
    SparseCorr corr;
    corr.ul_next = WL.other(0);
    corr.born_scale_m = ic != JUDI_IC_AS;
    sparse_step(M, S_src, S_src ? q->row(t) : nullptr, W.other(0), &corr, S_rec, WL.level(0),
                recl ? recl->row(t) : nullptr, recnl ? S_rec : nullptr, W.level(0),
                recnl ? recnl->row(t) : nullptr);
    if (M->fs && !M->visco)   // SLS_2nd_order carries no mirror equations (kernels.py:141)
        for (int f = 0; f < W.nf; f++) { free_surface(M, W.other(f)); free_surface(M, WL.other(f)); }
    W.swap();
    WL.swap();
}

static Fuse no_fuse()
{
    Fuse f;
    memset(&f, 0, sizeof(f));
    return f;
}

// ---- forward_rec / adjoint --------------------------------------------------------------------
void run_forward_rec(judi_model *M, int fw, int nsrc, const float *src_coords,
                     const float *wavelet, int nt, int nrec, const float *rec_coords,
                     float *rec_out, float *illum_out, int flags, const ExtSpec *ext)
{
    judi_ctx *ctx = M->ctx;
    bool dev = flags & JUDI_DATA_ON_DEVICE;
    JUDI_REQUIRE(nt >= 3, "need at least 3 time samples");
    Shot sh(M);
    sh.bind_ext(ext, wavelet, nt, dev);
    SparseSet S_src, S_rec;
    if (nsrc > 0) sparse_build(M, nsrc, src_coords, S_src);
    if (nrec > 0) sparse_build(M, nrec, rec_coords, S_rec);
    Traces q, rec;
    if (nsrc > 0) q.load(ctx, wavelet, nt, nsrc, dev);
    rec.alloc(ctx, nt, nrec);
    Wavefield W;
    W.init(M, false);
    DevBuf illum;
    Fuse fz = no_fuse();
    fz.reg = M->full;
    if (illum_out) { illum = DevBuf(ctx, (size_t)M->full.slot, true); fz.illum = illum.p; }
    const SparseSet *pS = nsrc > 0 ? &S_src : nullptr;
    sh.adjoint = !fw;
    P2Fuse pf;
    pf.reg = M->full;
    pf.illum = fz.illum;
    if (persist2d_supported(M, JUDI_IC_AS) && persist2d_bind_ext(M, sh, pf)) {
        persist2d_run(M, W, nullptr, fw ? 1 : nt - 2, fw ? nt - 2 : 1, fw ? 1 : -1, pS, &q,
                      nrec > 0 ? &S_rec : nullptr, &rec, nullptr, pf);
    } else if (fw) {
        for (int t = 1; t <= nt - 2; t++)
            sh.step_plain(W, t, illum_out ? &fz : nullptr, pS, &q, nullptr,
                          nrec > 0 ? &S_rec : nullptr, &rec);
    } else {
        for (int t = nt - 2; t >= 1; t--)
            sh.step_plain(W, t, illum_out ? &fz : nullptr, pS, &q, nullptr,
                          nrec > 0 ? &S_rec : nullptr, &rec);
    }
    sh.finish_ext(ext, dev);
    if (rec_out && nrec > 0) rec.store(ctx, rec_out, dev);
    if (illum_out) region_output(M, M->full, illum.p, illum_out, dev, false);
    ctx->flush_timing();
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
}

// ---- forward_no_rec: full wavefield history -----------------------------------------------------
void run_forward_no_rec(judi_model *M, int fw, int nsrc, const float *src_coords,
                        const float *wavelet, int nt, float *u_out, float *illum_out, int flags,
                        const ExtSpec *ext)
{
    judi_ctx *ctx = M->ctx;
    bool dev = flags & JUDI_DATA_ON_DEVICE;
    JUDI_REQUIRE(nt >= 3, "need at least 3 time samples");
    Shot sh(M);
    sh.bind_ext(ext, wavelet, nt, dev);
    SparseSet S_src;
    if (nsrc > 0) sparse_build(M, nsrc, src_coords, S_src);
    const SparseSet *pS = nsrc > 0 ? &S_src : nullptr;
    Traces q;
    if (nsrc > 0) q.load(ctx, wavelet, nt, nsrc, dev);
    Wavefield W;
    W.init(M, false);
    DevBuf illum;
    Fuse fz = no_fuse();
    fz.reg = M->full;
    if (illum_out) { illum = DevBuf(ctx, (size_t)M->full.slot, true); fz.illum = illum.p; }
    size_t ng = (size_t)M->N[0] * M->N[1] * M->N[2];
    // u.data[t] for t = 0..nt-1 (fields.py:41-42, save=nt): levels 0,1 (fw) / nt-1,nt-2 (adjoint)
    // are never written and stay zero.
    auto emit = [&](int t, const float *field) {
        grid_output(M, field, u_out + (size_t)t * ng, dev);
    };
    if (dev) CUDA_CHECK(cudaMemsetAsync(u_out, 0, (size_t)nt * ng * sizeof(float), ctx->stream));
    else memset(u_out, 0, (size_t)nt * ng * sizeof(float));
    sh.adjoint = !fw;
    if (fw) {
        for (int t = 1; t <= nt - 2; t++) {
            sh.step_plain(W, t, illum_out ? &fz : nullptr, pS, &q, nullptr, nullptr, nullptr);
            emit(t + 1, W.level(0));
        }
    } else {
        for (int t = nt - 2; t >= 1; t--) {
            sh.step_plain(W, t, illum_out ? &fz : nullptr, pS, &q, nullptr, nullptr, nullptr);
            emit(t - 1, W.level(0));
        }
    }
    if (illum_out) region_output(M, M->full, illum.p, illum_out, dev, false);
    ctx->flush_timing();
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
}

// ---- born_rec ------------------------------------------------------------------------------------
void run_born_rec(judi_model *M, int nsrc, const float *src_coords, const float *wavelet, int nt,
                  int nrec, const float *rec_coords, int ic, float *rec_out, float *illum_out,
                  int flags, const ExtSpec *ext)
{
    judi_ctx *ctx = M->ctx;
    bool dev = flags & JUDI_DATA_ON_DEVICE;
    JUDI_REQUIRE(nt >= 3, "need at least 3 time samples");
    size_t nr = (size_t)nt * nrec;
    if (!M->has_dm) {
        // propagators.py:181-194 / time_modeling_serial.jl:22-24: J*0 is exactly zero
        if (rec_out) {
            if (dev) CUDA_CHECK(cudaMemsetAsync(rec_out, 0, nr * sizeof(float), ctx->stream));
            else memset(rec_out, 0, nr * sizeof(float));
        }
        if (illum_out)
            run_forward_rec(M, 1, nsrc, src_coords, wavelet, nt, 0, nullptr, nullptr, illum_out,
                            flags, ext);
        return;
    }
    Shot sh(M);
    sh.bind_ext(ext, wavelet, nt, dev);
    SparseSet S_src, S_rec;
    if (nsrc > 0) sparse_build(M, nsrc, src_coords, S_src);
    sparse_build(M, nrec, rec_coords, S_rec);
    const SparseSet *pS = nsrc > 0 ? &S_src : nullptr;
    Traces q, rec;
    if (nsrc > 0) q.load(ctx, wavelet, nt, nsrc, dev);
    rec.alloc(ctx, nt, nrec);
    Wavefield W, WL;
    W.init(M, true);
    WL.init(M, false);
    DevBuf illum;
    Fuse fz = no_fuse();
    fz.reg = M->full;
    if (illum_out) { illum = DevBuf(ctx, (size_t)M->full.slot, true); fz.illum = illum.p; }
    P2Fuse pf;
    pf.reg = M->full;
    pf.ic = ic;
    pf.illum = fz.illum;
    if (persist2d_supported(M, ic, true) && persist2d_bind_ext(M, sh, pf)) {
        persist2d_run(M, W, &WL, 1, nt - 2, 1, pS, &q, &S_rec, &rec, nullptr, pf);
    } else {
        for (int t = 1; t <= nt - 2; t++)
            sh.step_born(W, WL, t, ic, illum_out ? &fz : nullptr, pS, &q, &S_rec, &rec, nullptr);
    }
    rec.store(ctx, rec_out, dev);
    if (illum_out) region_output(M, M->full, illum.p, illum_out, dev, false);
    ctx->flush_timing();
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
}

__global__ void batch_sum_kernel(double *acc, const double *f, int n)
{
    double s = *acc;
    for (int i = 0; i < n; i++) s += f[i];
    *acc = s;
}

// ---- J_adjoint -----------------------------------------------------------------------------------
static int num_slots(int nt, int t_sub)
{
    // fields.py:141-144
    return t_sub == 1 ? nt : (int)std::ceil((double)(nt + t_sub) / (double)t_sub);
}

// ---- several 2-D shots in lockstep (judi_fg_multishot fast path) -----------------------------------
// Small 2-D shots are bound by the grid barrier of the persistent kernel, not by bandwidth: all
// shots of the call (same model, same time axis) step together inside one cooperative launch per
// sweep, one barrier per time step for all of them.  Covers the plain FWI gradient of a constant-
// density acoustic model (imaging condition "as", stored snapshots, on-device L2 misfit); anything
// else returns false and the caller loops over run_J_adjoint.
bool run_fg_batch2d(judi_model *M, int nshots, const judi_shot *shots, const judi_jopts *o,
                    float *f_dev, float *g_dev, int flags)
{
    judi_ctx *ctx = M->ctx;
    if (nshots < 2 || !persist2d_batch_supported(M, o->ic) || o->checkpointing || o->nfreq > 0 ||
        o->ws || o->illum || !ctx->opt_persist2d)
        return false;
    // lsrtm_objective: Born forward sweeps (constant-density kernel only)
    const bool born = o->born_fwd && M->has_dm;
    if (born && (M->tti || M->irho_field || M->visco)) return false;
    if (o->nlind && !o->born_fwd) return false;
    const int nt = shots[0].nt;
    if (nt < 3) return false;
    for (int s = 0; s < nshots; s++)
        if (shots[s].nt != nt || shots[s].nsrc <= 0 || shots[s].nrec <= 0) return false;
    const bool dev = flags & JUDI_DATA_ON_DEVICE;
    const int k = std::max(1, o->t_sub);
    // ISIC/FWI differentiate the snapshot: it must cover the whole padded grid
    const Region reg = (o->snapshot_region == JUDI_SNAP_PHYSICAL && o->ic == JUDI_IC_AS) ? M->phys : M->full;
    const int nslots = num_slots(nt, k);
    const int nf = M->tti ? 2 : 1;
    const size_t slot_floats = (size_t)reg.slot * nf;
    // how many shots fit at once: snapshots + wavefields (+ flux scratch) + gradient per shot
    size_t free_b = 0, total_b = 0;
    CUDA_CHECK(cudaMemGetInfo(&free_b, &total_b));
    unsigned long long reserved = 0, used = 0;
    cudaMemPoolGetAttribute(ctx->pool, cudaMemPoolAttrReservedMemCurrent, &reserved);
    cudaMemPoolGetAttribute(ctx->pool, cudaMemPoolAttrUsedMemCurrent, &used);
    const double avail = 0.8 * ((double)free_b + (double)(reserved - used));
    const double per_shot = 4.0 * ((double)nslots * slot_floats + (4.0 * nf + 12.0) * M->g.nalloc + reg.slot) +
                            12.0 * (double)nt * shots[0].nrec;
    const int chunk = (int)std::max(1.0, std::min((double)std::min(nshots, 32), avail / per_shot));
    if (chunk < 2) return false;
    DevArr<double> fsum_dev(ctx, 1, true);
    PhaseClock pc(ctx);
    for (int s0 = 0; s0 < nshots; s0 += chunk) {
        const int n = std::min(chunk, nshots - s0);
        std::vector<SparseSet> S_src((size_t)n), S_rec((size_t)n);
        std::vector<Traces> q((size_t)n), obs((size_t)n), rec((size_t)n), recnl((size_t)n);
        std::vector<Wavefield> W((size_t)n), V((size_t)n), WLs((size_t)n);
        std::vector<Wavefield *> pWL((size_t)n);
        std::vector<Traces *> pnl((size_t)n);
        std::vector<DevBuf> snaps((size_t)n), grad((size_t)n);
        std::vector<P2Fuse> pf((size_t)n);
        std::vector<Wavefield *> pW((size_t)n), pV((size_t)n);
        std::vector<const SparseSet *> pSs((size_t)n), pSr((size_t)n);
        std::vector<const Traces *> pq((size_t)n), pres((size_t)n);
        std::vector<Traces *> prec((size_t)n);
        for (int s = 0; s < n; s++) {
            const judi_shot &sh = shots[s0 + s];
            sparse_build(M, sh.nsrc, sh.src_coords, S_src[s]);
            sparse_build(M, sh.nrec, sh.rec_coords, S_rec[s]);
            q[s].load(ctx, sh.wavelet, nt, sh.nsrc, dev);
            obs[s].load(ctx, sh.recin, nt, sh.nrec, dev);
            rec[s].alloc(ctx, nt, sh.nrec);
            W[s].init(M, false);
            if (born) {
                WLs[s].init(M, false);
                if (o->nlind) recnl[s].alloc(ctx, nt, sh.nrec);
            }
            pWL[s] = &WLs[s];
            pnl[s] = &recnl[s];
            snaps[s] = DevBuf(ctx, (size_t)nslots * slot_floats, false);
            grad[s] = DevBuf(ctx, (size_t)reg.slot, true);
            pf[s].reg = reg;
            pf[s].ic = o->ic;
            pf[s].snap = snaps[s].p;
            pf[s].slot = (long long)slot_floats;
            pf[s].t_sub = k;
            pf[s].t_base = 0;
            pW[s] = &W[s]; pSs[s] = &S_src[s]; pSr[s] = &S_rec[s]; pq[s] = &q[s]; prec[s] = &rec[s];
        }
        pc.mark("batch setup (sparse, traces, alloc)");
        // forward sweeps with snapshots
        persist2d_run_batch(M, n, pW.data(), 1, nt - 2, 1, pSs.data(), pq.data(), pSr.data(), prec.data(),
                            pf.data(), born ? pWL.data() : nullptr, (born && o->nlind) ? pnl.data() : nullptr);
        pc.mark("batch forward sweep");
        // born_fwd without a perturbation: zero linearised record (propagators.py:181-194)
        if (o->born_fwd && !born && !o->nlind)
            for (int s = 0; s < n; s++)
                CUDA_CHECK(cudaMemsetAsync(rec[s].buf.p, 0, (size_t)nt * shots[s0 + s].nrec * sizeof(float),
                                           ctx->stream));
        // misfit and residual (sensitivity.py:236-276), one small launch per shot
        DevArr<double> fd(ctx, (size_t)n, true);
        for (int s = 0; s < n; s++) {
            const long long nn = (long long)nt * shots[s0 + s].nrec;
            const int nb = (int)std::max<long long>(1, std::min<long long>((nn + 255) / 256, 1024));
            loss_kernel<<<nb, 256, 0, ctx->stream>>>(nn, rec[s].buf.p,
                                                     (born && o->nlind) ? recnl[s].buf.p : nullptr, obs[s].buf.p,
                                                     o->is_residual ? 1 : 0, fd.p + s, 0.5f * M->dt);
            launch_count(ctx, "loss");
        }
        // adjoint sweeps with the imaging condition
        for (int s = 0; s < n; s++) {
            V[s].init(M, false);
            W[s] = Wavefield();           // the forward fields are no longer needed
            WLs[s] = Wavefield();
            pV[s] = &V[s];
            pres[s] = &rec[s];
            pf[s] = P2Fuse();
            pf[s].reg = reg;
            pf[s].ic = o->ic;
            pf[s].grad = grad[s].p;
            pf[s].usnap = snaps[s].p;
            pf[s].slot = (long long)slot_floats;
            pf[s].t_sub = k;
            pf[s].t_base = 0;
            pf[s].gcoef = (float)((double)k / (double)M->dt);
        }
        pc.mark("batch misfit + adjoint setup");
        persist2d_run_batch(M, n, pV.data(), nt - 2, 1, -1, pSr.data(), pres.data(), nullptr, nullptr,
                            pf.data());
        pc.mark("batch adjoint sweep");
        for (int s = 0; s < n; s++) region_output(M, reg, grad[s].p, g_dev, true, true);
        batch_sum_kernel<<<1, 1, 0, ctx->stream>>>(fsum_dev.p, fd.p, n);
        launch_count(ctx, "batch_sum");
        pc.mark("batch outputs");
    }
    // f_dev += (float)(sum of the shots' objectives), on the device
    scalar_out_kernel<<<1, 1, 0, ctx->stream>>>(f_dev, fsum_dev.p, 0.0, 1);
    launch_count(ctx, "scalar_out");
    ctx->flush_timing();
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    return true;
}

void run_J_adjoint(judi_model *M, int nsrc, const float *src_coords, const float *wavelet, int nt,
                   int nrec, const float *rec_coords, const float *recin, const judi_jopts *o,
                   judi_misfit_fn misfit, void *user, float *f_out, float *grad_out,
                   float *illum_u, float *illum_v, int flags)
{
    judi_ctx *ctx = M->ctx;
    bool dev = flags & JUDI_DATA_ON_DEVICE;
    bool accumulate = flags & JUDI_GRAD_ACCUMULATE;
    bool dev_out = dev || (flags & JUDI_OUT_ON_DEVICE);
    JUDI_REQUIRE(nt >= 3, "need at least 3 time samples");
    JUDI_REQUIRE(o->t_sub >= 1, "t_sub must be >= 1");
    JUDI_REQUIRE(!(o->nlind && !o->born_fwd), "nlind requires born_fwd");
    JUDI_REQUIRE(o->ic >= JUDI_IC_AS && o->ic <= JUDI_IC_FWI, "unknown imaging condition");
    bool born = o->born_fwd && M->has_dm;
    const int nf = M->tti ? 2 : 1;
    const int k = o->checkpointing ? 1 : o->t_sub;
    // on-the-fly DFT (J_adjoint_freq, interface.py:348-408): frequency slices instead of snapshots
    // checkpointing wins over frequencies, as in the reference's dispatch (interface.py:331-345)
    const bool dft = o->nfreq > 0 && !o->checkpointing;
    const int dfac = std::max(o->dft_sub, 1);
    if (dft) {
        JUDI_REQUIRE(o->freq_list != nullptr, "freq_list is NULL");
    }
    // ISIC/FWI differentiate the snapshot: it must cover the whole padded grid
    const Region reg = (o->snapshot_region == JUDI_SNAP_PHYSICAL && o->ic == JUDI_IC_AS)
                           ? M->phys : M->full;

    PhaseClock pc(ctx);
    Shot sh(M);
    ExtSpec jext;
    jext.ws = o->ws;
    sh.bind_ext(o->ws ? &jext : nullptr, wavelet, nt, dev);
    SparseSet S_src0, S_rec;
    if (nsrc > 0) sparse_build(M, nsrc, src_coords, S_src0);
    sparse_build(M, nrec, rec_coords, S_rec);
    const SparseSet *pSrc = nsrc > 0 ? &S_src0 : nullptr;
    Traces q, obs, rec, recnl;
    if (nsrc > 0) q.load(ctx, wavelet, nt, nsrc, dev);
    obs.load(ctx, recin, nt, nrec, dev);
    rec.alloc(ctx, nt, nrec);
    if (born && o->nlind) recnl.alloc(ctx, nt, nrec);

    Wavefield W, WL;
    W.init(M, born);
    if (born) WL.init(M, false);
    // Checkpointing recomputes the background field with plain steps and must reproduce the Born
    // forward sweep bit for bit, so every step of this wavefield takes the kernel family the Born
    // step takes: single pass where it has a Born flavour (tti3d.cu), else the two-pass kernels;
    // the one-point-per-thread kernels for the isic / fwi Born sources of the flux-form physics.
    {
        const bool fluxp = M->tti || M->irho_field;
        W.generic_only = born && fluxp && o->ic != JUDI_IC_AS;
        W.two_pass_only = born && !W.generic_only && !flux_born_single_pass(M);
    }
    DevBuf Iu, Iv, grad(ctx, (size_t)reg.slot, true);
    if (illum_u) Iu = DevBuf(ctx, (size_t)reg.slot, true);
    if (illum_v) Iv = DevBuf(ctx, (size_t)reg.slot, true);

    // ---- snapshot storage ------------------------------------------------------------------
    const size_t slot_floats = (size_t)reg.slot * nf;
    // Checkpointing in HBM (interface.py:473-562 replaces pyrevolve's schedule): the time steps
    // 1 .. T are split into   [ segments of `seg` steps, each with a checkpointed state ] [ the
    // last `ndirect` steps ].  The snapshots of the last `ndirect` levels are written by the
    // first forward sweep and stay in HBM (they are the first ones the reverse sweep needs); only
    // the segments are recomputed, one at a time, into a buffer of `seg` snapshots.  With 180 GB
    // of HBM the whole history often fits (C3: 1040 physical-domain levels = 135 GB): then
    // ndirect = T, nothing is checkpointed and nothing recomputed -- what Revolve's schedule
    // degenerates to when it is given as many checkpoints as time steps.  A user-given
    // n_checkpoints keeps its meaning (a budget of states): seg = nt / n_checkpoints, ndirect = 0.
    int seg = 0, nseg = 0;          // checkpointing: segment length / count
    int ndirect = 0;                // checkpointing: trailing time levels kept as snapshots
    int nslots;
    std::vector<DevBuf> ckpt;       // per segment: level t and level t-1 of every field
    std::vector<DevBuf> ckpt_r;     // ... and the visco-acoustic memory variable
    double avail = 0.0, fixed_need = 0.0;
    {
        size_t free_b = 0, total_b = 0;
        CUDA_CHECK(cudaMemGetInfo(&free_b, &total_b));
        unsigned long long reserved = 0, used = 0;
        cudaMemPoolGetAttribute(ctx->pool, cudaMemPoolAttrReservedMemCurrent, &reserved);
        cudaMemPoolGetAttribute(ctx->pool, cudaMemPoolAttrUsedMemCurrent, &used);
        avail = (double)free_b + (double)(reserved - used) + 4.0 * (double)ctx->arena_floats;
        // what the sweeps allocate besides snapshots: the adjoint wavefield V (W and WL exist
        // already), the flux-form or Born scratch of each wavefield, gradient and illumination
        const bool fluxp = M->tti || M->irho_field;
        const double field = (double)M->g.nalloc * 4.0;
        const double scratch = (fluxp ? (M->tti ? 6.0 : 3.0) : 0.0) + (born ? 2.0 * nf : 0.0);
        fixed_need = (2.0 * nf + scratch) * field + 3.0 * (double)reg.slot * 4.0 +
                     (dft ? 2.0 * o->nfreq * slot_floats * 4.0 : 0.0);
    }
    if (dft) {
        nslots = 0;
    } else if (!o->checkpointing) {
        nslots = num_slots(nt, k);
    } else {
        const int T = nt - 2;
        const double state = (2.0 * nf + (M->visco ? 1.0 : 0.0)) * (double)M->g.nalloc * 4.0;
        const double snapb = (double)slot_floats * 4.0;
        const double budget = 0.90 * avail - fixed_need;
        // (option ckpt_direct = N > 1 forces a tail of exactly N levels: test hook of the hybrid path)
        const int forced = ctx->opt_ckpt_direct > 1 ? std::min(ctx->opt_ckpt_direct, T - 1) : -1;
        if (o->n_checkpoints <= 0 && ctx->opt_ckpt_direct == 1 && (double)T * snapb <= budget) {
            ndirect = T;            // everything fits: store, never recompute
        } else {
            // fixed-stride checkpoints + segment recompute: memory ~ seg snapshots + (T/seg) states;
            // seg ~ sqrt(T * state/snapshot) balances the two
            seg = (int)std::ceil(std::sqrt((double)nt * state / snapb));
            if (o->n_checkpoints > 0) seg = std::max(1, (nt + o->n_checkpoints - 1) / o->n_checkpoints);
            seg = std::max(1, std::min(seg, T));
            if (o->n_checkpoints <= 0 && ctx->opt_ckpt_direct) {
                // whatever HBM is left after the recompute buffer and the states holds the tail
                const double left = budget - (double)seg * snapb - std::ceil((double)T / seg) * state;
                ndirect = (int)std::max(0.0, std::min((double)(T - 1), std::floor(left / snapb)));
                if (forced >= 0) ndirect = forced;
            }
            nseg = (T - ndirect + seg - 1) / seg;
        }
        nslots = seg + ndirect;
    }
    {
        const double field = (double)M->g.nalloc * 4.0;
        double need = (double)nslots * slot_floats * 4.0 +
                      (double)nseg * (2.0 * nf + (M->visco ? 1.0 : 0.0)) * field + fixed_need;
        if (need > 0.97 * avail) {
            char b[256];
            snprintf(b, sizeof(b),
                     "wavefield snapshots need %.1f GB but only %.1f GB of HBM are free: raise "
                     "subsampling_factor, use snapshot_region=physical or optimal_checkpointing",
                     need / 1e9, avail / 1e9);
            throw JudiError(b);
        }
    }
    struct { float *p; } snaps = {ctx->arena((size_t)nslots * slot_floats)};
    auto slot_ptr = [&](int s, int f) { return snaps.p + ((size_t)s * nf + f) * reg.slot; };
    // DFT: uf[field][freq][re,im] on the region, and per-step phase tables
    //   forward  (fields_exprs.py:164-167): uf += factor exp(-i w t) u[t]        when t % factor == 0
    //   gradient (sensitivity.py:94-103):   gradm -= w_f Re(uf exp(i w t)) v[t],  w_f = -(2 pi f)^2 / time_M
    // on EVERY step: grad_expr passes dft_sub to the imaging condition as `factor=` (sensitivity.py:50),
    // a keyword crosscorr_freq does not have, so the adjoint correlation is never subsampled.
    // isic_freq / fwi_freq (sensitivity.py:125-154) do read it: they run when t % factor == 0, with
    // gradm -= Re(w_f U v m - w2 inner_grad(U, v)),  U = uf e^{i w t},  w2 = +-factor / time_M.
    DevBuf uf, cs_tab, wc_tab, ec_tab, dftB;
    const float dft_w2 = (o->ic == JUDI_IC_FWI ? -1.f : 1.f) * (float)dfac / (float)(nt - 2);
    const int ndft = dft ? (nt / dfac + 1) : 0;
    if (dft) {
        uf = DevBuf(ctx, (size_t)2 * o->nfreq * slot_floats, true);
        std::vector<float> cs((size_t)ndft * 2 * o->nfreq), wc((size_t)nt * 2 * o->nfreq), ec(wc.size());
        for (int q_ = 0; q_ < o->nfreq; q_++) {
            const double f = (double)o->freq_list[q_];
            const double w = -(2.0 * M_PI * f) * (2.0 * M_PI * f) / (double)(nt - 2);
            for (int j = 0; j < ndft; j++) {
                const double wt = 2.0 * M_PI * f * (double)(j * dfac) * (double)M->dt;
                cs[((size_t)j * o->nfreq + q_) * 2] = (float)((double)dfac * std::cos(wt));
                cs[((size_t)j * o->nfreq + q_) * 2 + 1] = (float)(-(double)dfac * std::sin(wt));
            }
            for (int t = 0; t < nt; t++) {
                const double wt = 2.0 * M_PI * f * (double)t * (double)M->dt;
                wc[((size_t)t * o->nfreq + q_) * 2] = (float)(w * std::cos(wt));
                wc[((size_t)t * o->nfreq + q_) * 2 + 1] = (float)(-w * std::sin(wt));
                ec[((size_t)t * o->nfreq + q_) * 2] = (float)std::cos(wt);
                ec[((size_t)t * o->nfreq + q_) * 2 + 1] = (float)(-std::sin(wt));
            }
        }
        cs_tab = DevBuf(ctx, cs.size(), false);
        cs_tab.upload(cs);
        wc_tab = DevBuf(ctx, wc.size(), false);
        wc_tab.upload(wc);
        if (o->ic != JUDI_IC_AS) {
            ec_tab = DevBuf(ctx, ec.size(), false);
            ec_tab.upload(ec);
            dftB = DevBuf(ctx, (size_t)reg.slot, false);
        }
    }
    auto uf_ptr = [&](int f) { return uf.p + (size_t)f * 2 * o->nfreq * reg.slot; };

    pc.mark("setup (sparse, traces, alloc)");
    Fuse fsave = no_fuse();
    fsave.reg = reg;
    fsave.illum = illum_u ? Iu.p : nullptr;

    // ---- forward sweep ---------------------------------------------------------------------
    auto fwd_step = [&](int t, int save_slot) {
        sh.adjoint = false;
        Fuse f = fsave;
        if (save_slot >= 0)
            for (int ff = 0; ff < nf; ff++) f.snap[ff] = slot_ptr(save_slot, ff);
        bool any = f.snap[0] || f.illum;
        if (born)
            sh.step_born(W, WL, t, o->ic, any ? &f : nullptr, pSrc, &q, &S_rec, &rec,
                         o->nlind ? &recnl : nullptr);
        else
            sh.step_plain(W, t, any ? &f : nullptr, pSrc, &q, nullptr, &S_rec, &rec);
    };
    // A range of forward steps.  mode 0: no snapshots; 1: slot t/k when t % k == 0; 2: slot t-base
    // (segment recomputation).  2-D constant-density models run the whole range inside one
    // persistent kernel (persist2d.cu), everything else steps kernel by kernel.
    P2Fuse pf_ext;      // J_adjoint(ws=...): the extended source drives the forward sweeps
    const bool p2 = persist2d_supported(M, o->ic, born) && persist2d_bind_ext(M, sh, pf_ext) && !dft;
    auto fwd_range = [&](int ta, int tb, int mode, int base, bool record) {
        if (tb < ta) return;
        if (p2) {
            P2Fuse pf = pf_ext;
            pf.reg = reg;
            pf.ic = o->ic;
            pf.illum = fsave.illum;
            if (mode) {
                pf.snap = snaps.p;
                pf.slot = (long long)slot_floats;
                pf.t_sub = mode == 1 ? k : 1;
                pf.t_base = mode == 1 ? 0 : base;
            }
            persist2d_run(M, W, (born && record) ? &WL : nullptr, ta, tb, 1, pSrc, &q,
                          record ? &S_rec : nullptr, &rec,
                          (record && born && o->nlind) ? &recnl : nullptr, pf);
            return;
        }
        for (int t = ta; t <= tb; t++) {
            int slot = mode == 1 ? ((t % k) == 0 ? t / k : -1) : (mode == 2 ? t - base : -1);
            if (dft) slot = -1;
            if (record) fwd_step(t, slot);
            else {
                sh.adjoint = false;
                Fuse f = fsave;
                for (int ff = 0; ff < nf; ff++) f.snap[ff] = slot >= 0 ? slot_ptr(slot, ff) : nullptr;
                sh.step_plain(W, t, &f, pSrc, &q, nullptr, nullptr, nullptr);
            }
            if (dft && (t % dfac) == 0)   // level t sits in the "other" buffer after the swap
                for (int ff = 0; ff < nf; ff++)
                    dft_accumulate(M, reg, W.other(ff), uf_ptr(ff), o->nfreq,
                                   cs_tab.p + (size_t)(t / dfac) * 2 * o->nfreq);
        }
    };
    if (!o->checkpointing) {
        fwd_range(1, nt - 2, 1, 0, true);
    } else {
        // with born_fwd only the background field is checkpointed: the recomputation needs the
        // snapshots of u, not the linearised field
        ckpt.resize((size_t)nseg * 2 * nf);
        if (M->visco) ckpt_r.resize((size_t)nseg);   // the memory variable is part of the state
        const int tseg_end = nt - 2 - ndirect;     // last time step covered by a segment
        for (int s = 0; s < nseg; s++) {
            int ts = 1 + s * seg, te = std::min(ts + seg - 1, tseg_end);
            if (M->visco) {
                ckpt_r[s] = DevBuf(ctx, (size_t)M->g.nalloc, false);
                CUDA_CHECK(cudaMemcpyAsync(ckpt_r[s].p, W.rmem.p, (size_t)M->g.nalloc * sizeof(float),
                                           cudaMemcpyDeviceToDevice, ctx->stream));
            }
            for (int ff = 0; ff < nf; ff++) {
                for (int lv = 0; lv < 2; lv++) {
                    DevBuf &c = ckpt[((size_t)s * nf + ff) * 2 + lv];
                    c = DevBuf(ctx, (size_t)M->g.nalloc, false);
                    CUDA_CHECK(cudaMemcpyAsync(c.p, lv == 0 ? W.level(ff) : W.other(ff),
                                               (size_t)M->g.nalloc * sizeof(float),
                                               cudaMemcpyDeviceToDevice, ctx->stream));
                }
            }
            fwd_range(ts, te, 0, 0, true);
        }
        // the trailing levels go straight to their slots (behind the recompute buffer of `seg` slots)
        if (ndirect > 0) fwd_range(tseg_end + 1, nt - 2, 2, tseg_end + 1 - seg, true);
        fsave.illum = nullptr;  // the recomputation must not add to the illumination again
    }

    pc.mark("forward sweep");
    // born_fwd without a perturbation (propagators.py:181-194): the reference runs the plain
    // forward operator for the wavefield and returns the untouched, zero linearised record -- the
    // background data only reach the Loss through `rnl` when nlind is set.
    if (o->born_fwd && !born && !o->nlind)
        CUDA_CHECK(cudaMemsetAsync(rec.buf.p, 0, (size_t)nt * std::max(nrec, 1) * sizeof(float), ctx->stream));
    // ---- misfit (sensitivity.py:236-276) ---------------------------------------------------
    double fval = 0.0;
    DevArr<double> fd;      // the on-device misfit value (stays in HBM when the output does)
    if (misfit) {
        size_t n = (size_t)nt * nrec;
        std::vector<float> hsyn(n), hobs(n), hres(n);
        if (born && o->nlind) {
            // misfit(dsyn[0], dobs - dsyn[1]) with dsyn = (rcvl, rnl)
            std::vector<float> hnl(n);
            rec.store(ctx, hsyn.data(), false);
            recnl.store(ctx, hnl.data(), false);
            obs.store(ctx, hobs.data(), false);
            CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
            for (size_t i = 0; i < n; i++) hobs[i] -= hnl[i];
        } else {
            rec.store(ctx, hsyn.data(), false);
            obs.store(ctx, hobs.data(), false);
        }
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        float fv = misfit(hsyn.data(), hobs.data(), hres.data(), nt, nrec, user);
        fval = (double)M->dt * (double)fv;
        rec.load(ctx, hres.data(), nt, nrec, false);
    } else {
        fd = DevArr<double>(ctx, 1, true);
        long long n = (long long)nt * nrec;
        int nb = (int)std::min<long long>((n + 255) / 256, 1024);
        if (nb < 1) nb = 1;
        loss_kernel<<<nb, 256, 0, ctx->stream>>>(n, rec.buf.p,
                                                 (born && o->nlind) ? recnl.buf.p : nullptr,
                                                 obs.buf.p, o->is_residual ? 1 : 0, fd.p,
                                                 0.5f * M->dt);
        launch_count(ctx, "loss");
        if (f_out && !dev_out) {
            CUDA_CHECK(cudaMemcpyAsync(&fval, fd.p, sizeof(double), cudaMemcpyDeviceToHost,
                                       ctx->stream));
            CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        }
    }
    // rec now holds the adjoint source (residual)

    pc.mark("misfit");
    // ---- adjoint sweep with the imaging condition ---------------------------------------------
    Wavefield V;
    V.init(M, false);
    Fuse fadj = no_fuse();
    fadj.reg = reg;
    fadj.illum = illum_v ? Iv.p : nullptr;
    fadj.ic = o->ic;
    fadj.gcoef = (float)((double)k / (double)M->dt);
    auto adj_step = [&](int t, int slot) {
        Fuse f = fadj;
        SparseCorr corr;
        if (slot >= 0) {
            f.grad = grad.p;
            for (int ff = 0; ff < nf; ff++) f.usnap[ff] = slot_ptr(slot, ff);
            corr.grad = grad.p;
            corr.usnap = slot_ptr(slot, 0);
            corr.gcoef = fadj.gcoef;
            corr.scale_m = o->ic != JUDI_IC_AS;
            corr.reg = reg;
        }
        bool any = f.grad || f.illum;
        sh.adjoint = true;
        sh.ext_on = false;  // the extended source drives the forward sweep only
        sh.step_plain(V, t, any ? &f : nullptr, &S_rec, &rec, slot >= 0 ? &corr : nullptr, nullptr,
                      nullptr);
        sh.ext_on = true;
    };
    auto adj_range = [&](int ta, int tb, int mode, int base) {  // ta >= tb, descending
        if (ta < tb) return;
        if (p2) {
            P2Fuse pf;
            pf.reg = reg;
            pf.ic = o->ic;
            pf.illum = fadj.illum;
            pf.grad = grad.p;
            pf.usnap = snaps.p;
            pf.slot = (long long)slot_floats;
            pf.t_sub = mode == 1 ? k : 1;
            pf.t_base = mode == 1 ? 0 : base;
            pf.gcoef = fadj.gcoef;
            persist2d_run(M, V, nullptr, ta, tb, -1, &S_rec, &rec, nullptr, nullptr, nullptr, pf);
            return;
        }
        for (int t = ta; t >= tb; t--) {
            adj_step(t, dft ? -1 : (mode == 1 ? ((t % k) == 0 ? t / k : -1) : t - base));
            if (dft && o->ic == JUDI_IC_AS)
                for (int ff = 0; ff < nf; ff++)
                    dft_gradient(M, reg, V.other(ff), uf_ptr(ff), o->nfreq,
                                 wc_tab.p + (size_t)t * 2 * o->nfreq, grad.p);
            else if (dft && (t % dfac) == 0)
                for (int ff = 0; ff < nf; ff++)
                    dft_gradient_isic(M, reg, V.other(ff), uf_ptr(ff), o->nfreq,
                                      wc_tab.p + (size_t)t * 2 * o->nfreq, ec_tab.p + (size_t)t * 2 * o->nfreq,
                                      dft_w2, dftB.p, grad.p);
        }
    };
    if (!o->checkpointing) {
        adj_range(nt - 2, 1, 1, 0);
    } else {
        const int tseg_end = nt - 2 - ndirect;
        if (ndirect > 0) adj_range(nt - 2, tseg_end + 1, 2, tseg_end + 1 - seg);
        for (int s = nseg - 1; s >= 0; s--) {
            int ts = 1 + s * seg, te = std::min(ts + seg - 1, tseg_end);
            // restore the state at the start of the segment and recompute its snapshots
            W.cur = 0;
            if (M->visco) {
                CUDA_CHECK(cudaMemcpyAsync(W.rmem.p, ckpt_r[s].p, (size_t)M->g.nalloc * sizeof(float),
                                           cudaMemcpyDeviceToDevice, ctx->stream));
                ckpt_r[s].reset();
            }
            for (int ff = 0; ff < nf; ff++)
                for (int lv = 0; lv < 2; lv++) {
                    DevBuf &c = ckpt[((size_t)s * nf + ff) * 2 + lv];
                    CUDA_CHECK(cudaMemcpyAsync(lv == 0 ? W.level(ff) : W.other(ff), c.p,
                                               (size_t)M->g.nalloc * sizeof(float),
                                               cudaMemcpyDeviceToDevice, ctx->stream));
                    c.reset();
                }
            fwd_range(ts, te, 2, ts, false);
            adj_range(te, ts, 2, ts);
        }
    }

    pc.mark("adjoint sweep");
    // ---- outputs -----------------------------------------------------------------------------
    region_output(M, reg, grad.p, grad_out, dev_out, accumulate);
    if (illum_u) region_output(M, reg, Iu.p, illum_u, dev_out, false);
    if (illum_v) region_output(M, reg, Iv.p, illum_v, dev_out, false);
    if (f_out) {
        if (dev_out) {
            // device scalar: the objective never leaves HBM (one-thread add, no stream sync)
            scalar_out_kernel<<<1, 1, 0, ctx->stream>>>(f_out, fd.p, fval, accumulate ? 1 : 0);
            launch_count(ctx, "scalar_out");
        } else {
            if (accumulate) *f_out += (float)fval;
            else *f_out = (float)fval;
        }
    }
    ctx->flush_timing();
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    pc.mark("outputs");
}
