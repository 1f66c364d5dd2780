// capi.cu -- the extern "C" boundary declared in include/judi_b200.h.
// Every function converts C++ exceptions into a non-zero return code + judi_last_error().
#include "common.cuh"
#include "propagate.h"

static thread_local std::string g_last_error;

#define API_BEGIN try {
#define API_END(ret_fail)                                   \
    }                                                       \
    catch (const std::exception &e) {                       \
        g_last_error = e.what();                            \
        return ret_fail;                                    \
    }                                                       \
    catch (...) {                                           \
        g_last_error = "unknown C++ exception";             \
        return ret_fail;                                    \
    }

extern "C" {

int judi_abi_version(void) { return JUDI_B200_ABI_VERSION; }

const char *judi_last_error(void) { return g_last_error.c_str(); }

judi_ctx *judi_ctx_create(int device)
{
    API_BEGIN
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        throw JudiError("no CUDA device available: libjudi_b200 has no CPU fallback");
    JUDI_REQUIRE(device >= 0 && device < ndev, "device index out of range");
    CUDA_CHECK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        throw JudiError(std::string("libjudi_b200 is built for sm_100a (B200) only; found ") +
                        prop.name);
    judi_ctx *ctx = new judi_ctx();
    // a failure below must not leak the half-built context
    struct Guard { judi_ctx *c; ~Guard() { if (c) judi_ctx_destroy(c); } } guard{ctx};
    ctx->device = device;
    ctx->num_sms = prop.multiProcessorCount;
    CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    // a PRIVATE stream-ordered pool: the release threshold below must not change the behaviour of
    // other cudaMallocAsync users of the host process (CUDA.jl, torch's async allocator)
    cudaMemPoolProps pp;
    memset(&pp, 0, sizeof(pp));
    pp.allocType = cudaMemAllocationTypePinned;
    pp.handleTypes = cudaMemHandleTypeNone;
    pp.location.type = cudaMemLocationTypeDevice;
    pp.location.id = device;
    CUDA_CHECK(cudaMemPoolCreate(&ctx->pool, &pp));
    ctx->own_pool = true;
    unsigned long long thresh = ~0ull;  // keep freed blocks: shots reuse the same buffers
    CUDA_CHECK(cudaMemPoolSetAttribute(ctx->pool, cudaMemPoolAttrReleaseThreshold, &thresh));
    cudaDriverEntryPointQueryResult qres;
    void *fn = nullptr;
    e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) ctx->encode_tiled = fn;
    else cudaGetLastError();
    // tuning switches from the environment, e.g. JUDI_B200_OPTIONS="tile=6,zchunks=8" (the keys of
    // judi_ctx_set_option): lets a caller that does not own the context (bench.py, a Julia session)
    // select a kernel variant without a code change
    if (const char *env = getenv("JUDI_B200_OPTIONS")) {
        std::string str(env);
        size_t pos = 0;
        while (pos < str.size()) {
            size_t end = str.find(',', pos);
            if (end == std::string::npos) end = str.size();
            const std::string kv = str.substr(pos, end - pos);
            const size_t eq = kv.find('=');
            if (eq != std::string::npos && judi_ctx_set_option(ctx, kv.substr(0, eq).c_str(), atoi(kv.c_str() + eq + 1)) != 0) {
                throw JudiError(std::string("JUDI_B200_OPTIONS: ") + judi_last_error());
            }
            pos = end + 1;
        }
    }
    guard.c = nullptr;
    return ctx;
    API_END(nullptr)
}

void judi_ctx_destroy(judi_ctx *ctx)
{
    if (!ctx) return;
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (auto &p : ctx->pending) { cudaEventDestroy(p.e0); cudaEventDestroy(p.e1); }
    for (auto e : ctx->free_events) cudaEventDestroy(e);
    comm_destroy(ctx);
    if (ctx->red_p) cudaFreeAsync(ctx->red_p, ctx->stream);
    if (ctx->arena_p) cudaFree(ctx->arena_p);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->own_pool) cudaMemPoolDestroy(ctx->pool);
    delete ctx;
}

__global__ void axpy_kernel(float *__restrict__ y, const float *__restrict__ x, size_t n)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] += x[i];
}
static void axpy_device(judi_ctx *ctx, float *y, const float *x, size_t n)
{
    axpy_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(y, x, n);
    launch_count(ctx, "axpy");
}

int judi_ctx_synchronize(judi_ctx *ctx)
{
    API_BEGIN
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    return 0;
    API_END(1)
}

void *judi_ctx_stream(judi_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

int judi_host_register(void *ptr, size_t nbytes)
{
    API_BEGIN
    JUDI_REQUIRE(ptr != nullptr && nbytes > 0, "judi_host_register: empty range");
    cudaError_t e = cudaHostRegister(ptr, nbytes, cudaHostRegisterPortable);
    if (e == cudaErrorHostMemoryAlreadyRegistered) { (void)cudaGetLastError(); return 0; }
    CUDA_CHECK(e);
    return 0;
    API_END(1)
}

int judi_host_unregister(void *ptr)
{
    API_BEGIN
    cudaError_t e = cudaHostUnregister(ptr);
    if (e == cudaErrorHostMemoryNotRegistered) { (void)cudaGetLastError(); return 0; }
    CUDA_CHECK(e);
    return 0;
    API_END(1)
}

long long judi_ctx_launch_count(judi_ctx *ctx) { return ctx ? ctx->launches : 0; }

int judi_ctx_stencil_stats(judi_ctx *ctx, double *ms, long long *launches, double *points)
{
    API_BEGIN
    ctx->flush_timing();
    if (ms) *ms = ctx->stencil_ms;
    if (launches) *launches = ctx->stencil_launches;
    if (points) *points = ctx->stencil_points;
    return 0;
    API_END(1)
}

int judi_ctx_stencil_stats_kind(judi_ctx *ctx, int kind, double *ms, long long *launches)
{
    API_BEGIN
    JUDI_REQUIRE(ctx && kind >= 0 && kind < JUDI_NUM_KINDS, "unknown launch kind");
    ctx->flush_timing();
    if (ms) *ms = ctx->kind_ms[kind];
    if (launches) *launches = ctx->kind_launches[kind];
    return 0;
    API_END(1)
}

int judi_ctx_reset_stats(judi_ctx *ctx)
{
    API_BEGIN
    ctx->flush_timing();
    for (int k = 0; k < JUDI_NUM_KINDS; k++) { ctx->kind_ms[k] = 0; ctx->kind_launches[k] = 0; }
    ctx->stencil_ms = 0;
    ctx->stencil_launches = 0;
    ctx->stencil_points = 0;
    ctx->launches = 0;
    return 0;
    API_END(1)
}

int judi_ctx_set_option(judi_ctx *ctx, const char *key, int value)
{
    API_BEGIN
    std::string k(key);
    if (k == "acoustic3d_kernel") ctx->opt_kernel3d = value;
    else if (k == "zchunks") ctx->opt_zchunks = value;
    else if (k == "tile") ctx->opt_tile = value;
    else if (k == "atomic_inject") ctx->opt_atomic_inject = value;
    else if (k == "timing") ctx->timing = value;
    else if (k == "persist2d") ctx->opt_persist2d = value;
    else if (k == "persist_vec") ctx->opt_persist_vec = value;
    else if (k == "ckpt_direct") ctx->opt_ckpt_direct = value;
    else if (k == "isic_fast") ctx->opt_isic_fast = value;
    else if (k == "profile") ctx->opt_profile = value;
    else if (k == "born_fused") ctx->opt_born_fused = value;
    else if (k == "tti_fused") ctx->opt_tti_fused = value;
    else if (k == "varc_fused") ctx->opt_varc_fused = value;
    else if (k == "persist_blocks") ctx->opt_persist_blocks = value;
    else if (k == "flux_slab") ctx->opt_flux_slab = value;
    else if (k == "flux_minb") ctx->opt_flux_minb = value;
    else if (k == "flux_block") ctx->opt_flux_block = value;
    else throw JudiError("unknown option " + k);
    return 0;
    API_END(1)
}

judi_model *judi_model_create(judi_ctx *ctx, const judi_model_desc *desc)
{
    judi_model *M = nullptr;
    try {
        JUDI_REQUIRE(ctx != nullptr && desc != nullptr, "null argument");
        CUDA_CHECK(cudaSetDevice(ctx->device));
        M = new judi_model();
        model_build(M, ctx, desc);
        return M;
    } catch (const std::exception &e) {
        g_last_error = e.what();
        delete M;
        return nullptr;
    }
}

void judi_model_destroy(judi_model *model)
{
    if (!model) return;
    cudaStreamSynchronize(model->ctx->stream);
    delete model;
}

float judi_model_critical_dt(const judi_model *model) { return model ? model->dt : 0.f; }

int judi_model_padsizes(const judi_model *model, int *out)
{
    API_BEGIN
    JUDI_REQUIRE(model && out, "null argument");
    const int dm2[2] = {0, 2};
    for (int k = 0; k < model->ndim; k++) {
        int gd = model->ndim == 3 ? k : dm2[k];
        out[2 * k] = model->padl[gd];
        out[2 * k + 1] = model->padr[gd];
    }
    return 0;
    API_END(1)
}

int judi_model_padded_shape(const judi_model *model, int *out)
{
    API_BEGIN
    JUDI_REQUIRE(model && out, "null argument");
    const int dm2[2] = {0, 2};
    for (int k = 0; k < model->ndim; k++) out[k] = model->N[model->ndim == 3 ? k : dm2[k]];
    return 0;
    API_END(1)
}

int judi_model_set_visco(judi_model *model, const judi_param *qp, float f0, int on_device)
{
    API_BEGIN
    JUDI_REQUIRE(model != nullptr, "model is NULL");
    model_set_visco(model, qp, f0, on_device);
    CUDA_CHECK(cudaStreamSynchronize(model->ctx->stream));
    return 0;
    API_END(1)
}

int judi_model_set_dm(judi_model *model, const judi_param *dm, int on_device)
{
    API_BEGIN
    JUDI_REQUIRE(model && dm, "null argument");
    CUDA_CHECK(cudaSetDevice(model->ctx->device));
    if (dm->kind == JUDI_PARAM_NONE || (dm->kind == JUDI_PARAM_SCALAR && dm->value == 0.f)) {
        model->has_dm = false;
        return 0;
    }
    model_set_param(model, model->dm, *dm, on_device, 0.f, true);
    model->has_dm = true;
    model_update_dm_src(model);
    CUDA_CHECK(cudaStreamSynchronize(model->ctx->stream));
    return 0;
    API_END(1)
}

__global__ void damp_field_kernel(Grid g, const float *dx, const float *dy, const float *dz,
                                  float *out)
{
    long long n = (long long)g.n[0] * g.n[1] * g.n[2];
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int x = (int)(i % g.n[0]);
    long long t = i / g.n[0];
    int y = (int)(t % g.n[1]), z = (int)(t / g.n[1]);
    out[i] = (dx[x] + dy[y]) + dz[z];
}

int judi_model_get_field(const judi_model *model, const char *name, float *out)
{
    API_BEGIN
    JUDI_REQUIRE(model && name && out, "null argument");
    judi_ctx *ctx = model->ctx;
    CUDA_CHECK(cudaSetDevice(ctx->device));
    std::string k(name);
    const float *f = nullptr;
    if (k == "m") f = model->m.p;
    else if (k == "irho") f = model->irho.p;
    else if (k == "epsilon") f = model->eps.p;
    else if (k == "delta") f = model->delta.p;
    else if (k == "theta") f = model->theta.p;
    else if (k == "phi") f = model->phi.p;
    else if (k == "dm") f = model->dm.p;
    else if (k == "damp") {
        long long n = (long long)model->N[0] * model->N[1] * model->N[2];
        DevBuf tmp(ctx, (size_t)n, false);
        damp_field_kernel<<<(int)((n + 255) / 256), 256, 0, ctx->stream>>>(
            model->g, model->damp[0].p, model->damp[1].p, model->damp[2].p, tmp.p);
        launch_count(ctx, "damp_field");
        CUDA_CHECK(cudaMemcpyAsync(out, tmp.p, n * sizeof(float), cudaMemcpyDeviceToHost,
                                   ctx->stream));
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        return 0;
    } else throw JudiError("unknown field " + k);
    JUDI_REQUIRE(f != nullptr, "field not present in this model");
    grid_output(model, f, out, false);
    return 0;
    API_END(1)
}

int judi_sparse_locate(const judi_model *model, int npoint, const float *coords, int *ijk, float *w,
                       int *valid)
{
    API_BEGIN
    JUDI_REQUIRE(model && coords && ijk && w && valid, "null argument");
    std::vector<int> vi, vv;
    std::vector<float> vw;
    sparse_locate_host(model, npoint, coords, vi, vw, vv);
    memcpy(ijk, vi.data(), vi.size() * sizeof(int));
    memcpy(w, vw.data(), vw.size() * sizeof(float));
    memcpy(valid, vv.data(), vv.size() * sizeof(int));
    return 0;
    API_END(1)
}

int judi_forward_rec(judi_model *model, int fw, int nsrc, const float *src_coords,
                     const float *wavelet, int nt, int nrec, const float *rec_coords,
                     float *rec_out, float *illum_out, int flags)
{
    API_BEGIN
    JUDI_REQUIRE(model && src_coords && wavelet, "null argument");
    CUDA_CHECK(cudaSetDevice(model->ctx->device));
    run_forward_rec(model, fw, nsrc, src_coords, wavelet, nt, nrec, rec_coords, rec_out, illum_out,
                    flags);
    return 0;
    API_END(1)
}

int judi_forward_no_rec(judi_model *model, int fw, int nsrc, const float *src_coords,
                        const float *wavelet, int nt, float *u_out, float *illum_out, int flags)
{
    API_BEGIN
    JUDI_REQUIRE(model && src_coords && wavelet && u_out, "null argument");
    CUDA_CHECK(cudaSetDevice(model->ctx->device));
    run_forward_no_rec(model, fw, nsrc, src_coords, wavelet, nt, u_out, illum_out, flags);
    return 0;
    API_END(1)
}

int judi_born_rec(judi_model *model, int nsrc, const float *src_coords, const float *wavelet,
                  int nt, int nrec, const float *rec_coords, int ic, float *rec_out,
                  float *illum_out, int flags)
{
    API_BEGIN
    JUDI_REQUIRE(model && src_coords && wavelet && rec_coords && rec_out, "null argument");
    CUDA_CHECK(cudaSetDevice(model->ctx->device));
    run_born_rec(model, nsrc, src_coords, wavelet, nt, nrec, rec_coords, ic, rec_out, illum_out,
                 flags);
    return 0;
    API_END(1)
}

int judi_J_adjoint(judi_model *model, int nsrc, const float *src_coords, const float *wavelet,
                   int nt, int nrec, const float *rec_coords, const float *recin,
                   const judi_jopts *opts, judi_misfit_fn misfit, void *user, float *f_out,
                   float *grad_out, float *illum_u, float *illum_v, int flags)
{
    API_BEGIN
    JUDI_REQUIRE(model && wavelet && rec_coords && recin && opts && grad_out, "null argument");
    JUDI_REQUIRE((src_coords && nsrc > 0) || opts->ws, "need point sources or extended-source weights");
    CUDA_CHECK(cudaSetDevice(model->ctx->device));
    judi_jopts o = *opts;
    if (!o.illum) { illum_u = nullptr; illum_v = nullptr; }
    run_J_adjoint(model, nsrc, src_coords, wavelet, nt, nrec, rec_coords, recin, &o, misfit, user,
                  f_out, grad_out, illum_u, illum_v, flags);
    return 0;
    API_END(1)
}

int judi_fg_multishot(judi_model *model, int nshots, const judi_shot *shots, const judi_jopts *opts,
                      judi_misfit_fn misfit, void *user, float *f_out, float *grad_out, int flags)
{
    API_BEGIN
    JUDI_REQUIRE(model && opts && grad_out && f_out && (shots || nshots == 0), "null argument");
    JUDI_REQUIRE(nshots >= 0, "negative shot count");
    judi_ctx *ctx = model->ctx;
    CUDA_CHECK(cudaSetDevice(ctx->device));
    const size_t ng = (size_t)model->N[0] * model->N[1] * model->N[2];
    const bool out_dev = (flags & JUDI_OUT_ON_DEVICE) || (flags & JUDI_DATA_ON_DEVICE);
    const bool accumulate = flags & JUDI_GRAD_ACCUMULATE;
    judi_jopts o = *opts;
    o.illum = 0;
    // the accumulator of this call: the caller's device buffers, or a temporary one in HBM
    DevBuf acc;
    float *g_dev = grad_out, *f_dev = f_out;
    if (!out_dev) {
        acc = DevBuf(ctx, ng + 1, true);
        g_dev = acc.p;
        f_dev = acc.p + ng;
    } else if (!accumulate) {
        CUDA_CHECK(cudaMemsetAsync(g_dev, 0, ng * sizeof(float), ctx->stream));
        CUDA_CHECK(cudaMemsetAsync(f_dev, 0, sizeof(float), ctx->stream));
    }
    const int shot_flags = (flags & JUDI_DATA_ON_DEVICE) | JUDI_OUT_ON_DEVICE | JUDI_GRAD_ACCUMULATE;
    // small 2-D shots: all of them in lockstep inside one cooperative launch per sweep
    const bool batched = !misfit && run_fg_batch2d(model, nshots, shots, &o, f_dev, g_dev, shot_flags);
    for (int i = 0; i < nshots && !batched; i++) {
        const judi_shot &sh = shots[i];
        JUDI_REQUIRE(sh.wavelet && sh.rec_coords && sh.recin, "shot with a null pointer");
        JUDI_REQUIRE((sh.src_coords && sh.nsrc > 0) || o.ws, "shot without a source");
        run_J_adjoint(model, sh.nsrc, sh.src_coords, sh.wavelet, sh.nt, sh.nrec, sh.rec_coords, sh.recin,
                      &o, misfit, user, f_dev, g_dev, nullptr, nullptr, shot_flags);
    }
    if (!out_dev) {
        std::vector<float> fsum(1);
        CUDA_CHECK(cudaMemcpyAsync(fsum.data(), f_dev, sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
        if (accumulate) {
            // add to the caller's host arrays: stage through a second device buffer
            DevBuf prev(ctx, ng, false);
            CUDA_CHECK(cudaMemcpyAsync(prev.p, grad_out, ng * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
            axpy_device(ctx, g_dev, prev.p, ng);
        }
        CUDA_CHECK(cudaMemcpyAsync(grad_out, g_dev, ng * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        *f_out = accumulate ? *f_out + fsum[0] : fsum[0];
    }
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    return 0;
    API_END(1)
}

int judi_fg_multishot_reduced(judi_model *model, int nshots, const judi_shot *shots, const judi_jopts *opts,
                              judi_misfit_fn misfit, void *user, int true_adjoint, float *f_out,
                              float *grad_out, int flags)
{
    API_BEGIN
    JUDI_REQUIRE(model && opts && grad_out && f_out, "null argument");
    judi_ctx *ctx = model->ctx;
    CUDA_CHECK(cudaSetDevice(ctx->device));
    const size_t ng = (size_t)model->N[0] * model->N[1] * model->N[2];
    DevBuf acc(ctx, ng + 1, true);
    // this rank's shots, summed in HBM (a rank may hold no shot at all: it still joins the reduction)
    if (nshots > 0) {
        int rc = judi_fg_multishot(model, nshots, shots, opts, misfit, user, acc.p + ng, acc.p,
                                   (flags & JUDI_DATA_ON_DEVICE) | JUDI_OUT_ON_DEVICE | JUDI_GRAD_ACCUMULATE);
        if (rc != 0) return rc;
    }
    run_fg_reduce(model, acc.p, acc.p + ng, true_adjoint, grad_out, f_out, flags & JUDI_OUT_ON_DEVICE);
    return 0;
    API_END(1)
}

int judi_forward_rec_w(judi_model *model, int fw, const float *weight, const float *wavelet, int nt,
                       int nrec, const float *rec_coords, float *rec_out, float *illum_out,
                       int flags)
{
    API_BEGIN
    JUDI_REQUIRE(model && weight && wavelet && rec_coords && rec_out, "null argument");
    CUDA_CHECK(cudaSetDevice(model->ctx->device));
    ExtSpec ext;
    ext.ws = weight;
    run_forward_rec(model, fw, 0, nullptr, wavelet, nt, nrec, rec_coords, rec_out, illum_out, flags,
                    &ext);
    return 0;
    API_END(1)
}

int judi_adjoint_w(judi_model *model, int fw, int nrec, const float *rec_coords, const float *data,
                   const float *wavelet, int nt, float *w_out, float *illum_out, int flags)
{
    API_BEGIN
    JUDI_REQUIRE(model && rec_coords && data && wavelet && w_out, "null argument");
    CUDA_CHECK(cudaSetDevice(model->ctx->device));
    ExtSpec ext;
    ext.wr_wavelet = wavelet;
    ext.w_out = w_out;
    run_forward_rec(model, fw, nrec, rec_coords, data, nt, 0, nullptr, nullptr, illum_out, flags,
                    &ext);
    return 0;
    API_END(1)
}

int judi_born_rec_w(judi_model *model, const float *weight, const float *wavelet, int nt, int nrec,
                    const float *rec_coords, int ic, float *rec_out, float *illum_out, int flags)
{
    API_BEGIN
    JUDI_REQUIRE(model && weight && wavelet && rec_coords && rec_out, "null argument");
    CUDA_CHECK(cudaSetDevice(model->ctx->device));
    ExtSpec ext;
    ext.ws = weight;
    run_born_rec(model, 0, nullptr, wavelet, nt, nrec, rec_coords, ic, rec_out, illum_out, flags,
                 &ext);
    return 0;
    API_END(1)
}

int judi_forward_wf_src(judi_model *model, int fw, const float *u_src, int nt, int nrec,
                        const float *rec_coords, float *rec_out, float *illum_out, int flags)
{
    API_BEGIN
    JUDI_REQUIRE(model && u_src && rec_coords && rec_out, "null argument");
    CUDA_CHECK(cudaSetDevice(model->ctx->device));
    ExtSpec ext;
    ext.qwf = u_src;
    run_forward_rec(model, fw, 0, nullptr, nullptr, nt, nrec, rec_coords, rec_out, illum_out, flags,
                    &ext);
    return 0;
    API_END(1)
}

int judi_forward_wf_src_norec(judi_model *model, int fw, const float *u_src, int nt, float *u_out,
                              float *illum_out, int flags)
{
    API_BEGIN
    JUDI_REQUIRE(model && u_src && u_out, "null argument");
    CUDA_CHECK(cudaSetDevice(model->ctx->device));
    ExtSpec ext;
    ext.qwf = u_src;
    run_forward_no_rec(model, fw, 0, nullptr, nullptr, nt, u_out, illum_out, flags, &ext);
    return 0;
    API_END(1)
}

int judi_remove_padding(const judi_model *model, const float *padded, int true_adjoint, float *out,
                        int flags)
{
    API_BEGIN
    JUDI_REQUIRE(model && padded && out, "null argument");
    CUDA_CHECK(cudaSetDevice(model->ctx->device));
    run_remove_padding(model, padded, true_adjoint, out, flags);
    return 0;
    API_END(1)
}

int judi_time_resample(judi_ctx *ctx, const float *in, int nt_in, int ntrace, float t0_in,
                       float dt_in, float *out, int nt_out, float t0_out, float dt_out, int flags)
{
    API_BEGIN
    JUDI_REQUIRE(ctx && in && out, "null argument");
    CUDA_CHECK(cudaSetDevice(ctx->device));
    run_time_resample(ctx, in, nt_in, ntrace, t0_in, dt_in, out, nt_out, t0_out, dt_out, flags);
    return 0;
    API_END(1)
}

int judi_ctx_trim(judi_ctx *ctx)
{
    API_BEGIN
    JUDI_REQUIRE(ctx, "null context");
    CUDA_CHECK(cudaSetDevice(ctx->device));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    if (ctx->red_p) { CUDA_CHECK(cudaFreeAsync(ctx->red_p, ctx->stream)); ctx->red_p = nullptr; ctx->red_floats = 0; }
    if (ctx->arena_p) { CUDA_CHECK(cudaFree(ctx->arena_p)); ctx->arena_p = nullptr; ctx->arena_floats = 0; }
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    CUDA_CHECK(cudaMemPoolTrimTo(ctx->pool, 0));
    return 0;
    API_END(1)
}

int judi_comm_unique_id(void *id128)
{
    API_BEGIN
    JUDI_REQUIRE(id128, "null argument");
    comm_unique_id(id128);
    return 0;
    API_END(1)
}

int judi_comm_init(judi_ctx *ctx, int nranks, int rank, const void *id128)
{
    API_BEGIN
    JUDI_REQUIRE(ctx && (id128 || nranks == 1), "null argument");
    CUDA_CHECK(cudaSetDevice(ctx->device));
    comm_init(ctx, nranks, rank, id128);
    return 0;
    API_END(1)
}

int judi_comm_destroy(judi_ctx *ctx)
{
    API_BEGIN
    JUDI_REQUIRE(ctx, "null context");
    CUDA_CHECK(cudaSetDevice(ctx->device));
    comm_destroy(ctx);
    return 0;
    API_END(1)
}

int judi_comm_size(judi_ctx *ctx) { return ctx ? ctx->comm_size : 0; }
int judi_comm_rank(judi_ctx *ctx) { return ctx ? ctx->comm_rank : -1; }

int judi_comm_allreduce(judi_ctx *ctx, float *dev, size_t n)
{
    API_BEGIN
    JUDI_REQUIRE(ctx && dev, "null argument");
    CUDA_CHECK(cudaSetDevice(ctx->device));
    comm_allreduce(ctx, dev, n);
    return 0;
    API_END(1)
}

int judi_fg_reduce(judi_model *model, const float *grad_padded_dev, const float *f_dev, int true_adjoint,
                   float *grad_out, float *f_out, int flags)
{
    API_BEGIN
    JUDI_REQUIRE(model && grad_padded_dev && grad_out, "null argument");
    CUDA_CHECK(cudaSetDevice(model->ctx->device));
    run_fg_reduce(model, grad_padded_dev, f_dev, true_adjoint, grad_out, f_out, flags);
    return 0;
    API_END(1)
}

}  // extern "C"
