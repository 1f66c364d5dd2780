// comm.cu -- the one collective of the path behind the C-ABI: the sum over shots / workers of
// (objective, gradient) that the reference does with `reduce!` / `run_and_reduce`, a binary tree
// over Julia Futures (src/TimeModeling/Modeling/distributed.jl:60-95, propagation.jl:23-50).
//
// One process per GPU; each rank accumulates its shots in HBM; judi_fg_reduce crops (or folds) the
// padded accumulator on the device, appends the objective and issues ONE ncclAllReduce of
// prod(n_physical)+1 floats on the context's stream (C3: 129 MB instead of the 260 MB padded
// grid).  NCCL is bound at run time with dlopen -- the library has no link-time dependency on it,
// and a caller that never creates a communicator never loads it.  The host process supplies the
// out-of-band channel for the 128-byte unique id (Julia: Distributed; Python: the rendezvous
// store), exactly as with any NCCL program.
#include "common.cuh"
#include "propagate.h"
#include <dlfcn.h>

namespace {

// minimal NCCL ABI (nccl.h 2.x: ncclUniqueId is 128 opaque bytes passed by value)
struct NcclId { char internal[128]; };
typedef void *NcclComm;
typedef int (*fn_GetUniqueId)(NcclId *);
typedef int (*fn_CommInitRank)(NcclComm *, int, NcclId, int);
typedef int (*fn_CommDestroy)(NcclComm);
typedef int (*fn_AllReduce)(const void *, void *, size_t, int, int, NcclComm, cudaStream_t);
typedef const char *(*fn_GetErrorString)(int);
typedef int (*fn_GetVersion)(int *);
const int kNcclFloat32 = 7, kNcclSum = 0;

struct NcclApi {
    void *h = nullptr;
    fn_GetUniqueId GetUniqueId = nullptr;
    fn_CommInitRank CommInitRank = nullptr;
    fn_CommDestroy CommDestroy = nullptr;
    fn_AllReduce AllReduce = nullptr;
    fn_GetErrorString GetErrorString = nullptr;
    fn_GetVersion GetVersion = nullptr;
};

// loaded once per process (thread-safe: function-local static initialised by a lambda)
NcclApi load_nccl();
NcclApi &nccl()
{
    static NcclApi api = load_nccl();
    return api;
}

NcclApi load_nccl()
{
    NcclApi api;
    const char *env = getenv("JUDI_NCCL_LIB");
    const char *names[] = {env, "libnccl.so.2", "libnccl.so"};
    std::string tried;
    for (const char *n : names) {
        if (!n || !*n) continue;
        api.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (api.h) break;
        tried += std::string(n) + ": " + dlerror() + "; ";
    }
    if (!api.h) throw JudiError("cannot load NCCL (set JUDI_NCCL_LIB): " + tried);
    auto sym = [&](const char *s) {
        void *p = dlsym(api.h, s);
        if (!p) throw JudiError(std::string("NCCL symbol missing: ") + s);
        return p;
    };
    api.GetUniqueId = (fn_GetUniqueId)sym("ncclGetUniqueId");
    api.CommInitRank = (fn_CommInitRank)sym("ncclCommInitRank");
    api.CommDestroy = (fn_CommDestroy)sym("ncclCommDestroy");
    api.AllReduce = (fn_AllReduce)sym("ncclAllReduce");
    api.GetErrorString = (fn_GetErrorString)sym("ncclGetErrorString");
    api.GetVersion = (fn_GetVersion)sym("ncclGetVersion");
    return api;
}

void nccl_check(int rc, const char *what)
{
    if (rc != 0)
        throw JudiError(std::string("NCCL error in ") + what + ": " + nccl().GetErrorString(rc));
}

// one thread: acc[n] = f (the objective rides along as the last element of the reduced buffer)
__global__ void set_tail_kernel(float *buf, size_t n, const float *f_dev) { buf[n] = f_dev ? *f_dev : 0.f; }

}  // namespace

void comm_unique_id(void *id128)
{
    NcclId id;
    nccl_check(nccl().GetUniqueId(&id), "ncclGetUniqueId");
    memcpy(id128, &id, sizeof(id));
}

void comm_init(judi_ctx *ctx, int nranks, int rank, const void *id128)
{
    JUDI_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "bad rank / nranks");
    JUDI_REQUIRE(ctx->comm == nullptr, "this context already has a communicator");
    ctx->comm_size = nranks;
    ctx->comm_rank = rank;
    if (nranks == 1) return;   // a single rank reduces with itself: nothing to load
    NcclId id;
    memcpy(&id, id128, sizeof(id));
    NcclComm c = nullptr;
    nccl_check(nccl().CommInitRank(&c, nranks, id, rank), "ncclCommInitRank");
    ctx->comm = c;
}

void comm_destroy(judi_ctx *ctx)
{
    if (ctx->comm) {
        cudaStreamSynchronize(ctx->stream);
        nccl().CommDestroy((NcclComm)ctx->comm);
        ctx->comm = nullptr;
    }
    ctx->comm_size = 1;
    ctx->comm_rank = 0;
}

void comm_allreduce(judi_ctx *ctx, float *dev, size_t n)
{
    if (ctx->comm_size <= 1 || n == 0) return;
    JUDI_REQUIRE(ctx->comm != nullptr, "no communicator: call judi_comm_init first");
    nccl_check(nccl().AllReduce(dev, dev, n, kNcclFloat32, kNcclSum, (NcclComm)ctx->comm, ctx->stream),
               "ncclAllReduce");
    ctx->collectives++;
}

// multi_src_fg!'s reduction (propagation.jl:99-106 + distributed.jl:60-95) and post-processing
// (misfit_fg.jl:77 remove_padding) in one call: padded accumulator in HBM -> physical-grid
// gradient + objective summed over all ranks.
void run_fg_reduce(judi_model *M, const float *grad_padded_dev, const float *f_dev, int true_adjoint,
                   float *grad_out, float *f_out, int flags)
{
    judi_ctx *ctx = M->ctx;
    const bool out_dev = flags & JUDI_OUT_ON_DEVICE;
    const size_t np = (size_t)M->n[0] * M->n[1] * M->n[2];
    if (ctx->red_floats < np + 1) {
        if (ctx->red_p) ctx->free(ctx->red_p);
        ctx->red_p = ctx->alloc(np + 1, false);
        ctx->red_floats = np + 1;
    }
    float *red = ctx->red_p;
    run_remove_padding(M, grad_padded_dev, true_adjoint, red, JUDI_DATA_ON_DEVICE | JUDI_OUT_ON_DEVICE);
    set_tail_kernel<<<1, 1, 0, ctx->stream>>>(red, np, f_dev);
    launch_count(ctx, "set_tail");
    comm_allreduce(ctx, red, np + 1);
    if (out_dev) {
        CUDA_CHECK(cudaMemcpyAsync(grad_out, red, np * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
        if (f_out)
            CUDA_CHECK(cudaMemcpyAsync(f_out, red + np, sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
    } else {
        CUDA_CHECK(cudaMemcpyAsync(grad_out, red, np * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
        if (f_out)
            CUDA_CHECK(cudaMemcpyAsync(f_out, red + np, sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    }
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
}
