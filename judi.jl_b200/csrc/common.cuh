// common.cuh -- shared host/device declarations of libjudi_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <stdexcept>
#include "../../include/judi_b200.h"

#define JUDI_MAXR 8       // space_order <= 16
#define JUDI_HALO 8       // zero halo (cells) kept around every padded-grid field in y and z
#define JUDI_HALO_XL 32   // left x halo: keeps domain point x=0 128-byte aligned
#define JUDI_HALO_XR 8

// ------------------------------------------------------------------------------------------
// Device layout of one padded-grid scalar field: x fastest (Julia's native order), zero halo.
struct Grid {
    int ndim;
    int n[3];          // padded-domain points along x, y, z (n[1] == 1 in 2-D)
    int hx, hy, hz;    // index of the first domain point inside the allocation
    int px, py, pz;    // allocated extents
    long long sy, sz;  // strides (elements) of y and z
    long long off0;    // element offset of domain point (0,0,0)
    long long nalloc;  // elements allocated
    __host__ __device__ inline long long at(int x, int y, int z) const {
        return off0 + x + (long long)y * sy + (long long)z * sz;
    }
};

// Box of the padded grid over which snapshots / gradient / illumination are produced, with the
// compact layout they are stored in (x fastest, pitch multiple of 4, no halo).
struct Region {
    int lo[3], hi[3];  // [lo, hi) in padded-domain coordinates
    int ext[3];        // hi - lo
    int pitch;         // x pitch of the compact layout
    long long slot;    // elements per field per snapshot slot
    __host__ __device__ inline long long at(int x, int y, int z) const {
        return (x - lo[0]) + (long long)pitch * ((y - lo[1]) + (long long)ext[1] * (z - lo[2]));
    }
    __host__ __device__ inline bool has(int x, int y, int z) const {
        return x >= lo[0] && x < hi[0] && y >= lo[1] && y < hi[1] && z >= lo[2] && z < hi[2];
    }
};

// Coefficients shared by every stencil kernel (passed by value: kernel parameters live in the
// constant bank, so these are uniform loads).
struct Coefs {
    int r;                      // space_order / 2
    float c0;                   // sum over dims of dt^2 * c2[0] / h_d^2
    float cx[JUDI_MAXR + 1];    // dt^2 * c2[k] / hx^2   (centred 2nd derivative, k = 1..r)
    float cy[JUDI_MAXR + 1];
    float cz[JUDI_MAXR + 1];
    float wx[JUDI_MAXR + 1];    // w1[k] / hx            (staggered 1st derivative, k = 1..r)
    float wy[JUDI_MAXR + 1];
    float wz[JUDI_MAXR + 1];
    float dt, dt2, idt, idt2;
};

// What the update epilogue fuses besides the leapfrog step (all optional).
struct Fuse {
    float *snap[2];         // SAVE: snapshot slot(s) to write level t into (forward)
    const float *usnap[2];  // IC:   snapshot slot(s) of the forward field(s) (adjoint)
    float *grad;            // IC:   gradient accumulator (region layout)
    float *illum;           // ILLUM: illumination accumulator (region layout)
    float gcoef;            // IC:   t_sub / dt  (multiplied by irho in the kernel)
    float *wrec;            // extended receiver accumulator (region layout): += wrec_scale * sum(u)
    float wrec_scale;       //   dt * wavelet[t]   (fields_exprs.py:85-103)
    int ic;                 // JUDI_IC_*
    Region reg;
};

// ------------------------------------------------------------------------------------------
// Packed fp32x2 arithmetic (sm_100a FADD2 / FMUL2 / FFMA2): two IEEE-rn operations per issue
// slot on an aligned register pair.  The stencil kernels are bound by instruction issue / the
// FMA pipe, not by HBM, when written with scalar FFMA (ncu: 77% issue-active on the plain 3-D
// step), so every point-wise expression is evaluated on pairs of x-neighbours.  Each packed op
// rounds exactly like its scalar twin, so results stay bit-identical to the scalar kernels.
#ifdef __CUDACC__
typedef float2 f2;
__device__ __forceinline__ f2 pk(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ f2 pk1(float a) { return make_float2(a, a); }
__device__ __forceinline__ f2 lo2(const float4 &v) { return make_float2(v.x, v.y); }
__device__ __forceinline__ f2 hi2(const float4 &v) { return make_float2(v.z, v.w); }
__device__ __forceinline__ float4 cat4(f2 a, f2 b) { return make_float4(a.x, a.y, b.x, b.y); }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ f2 sub2(f2 a, f2 b) { return __ffma2_rn(b, make_float2(-1.f, -1.f), a); }
__device__ __forceinline__ f2 mul2s(float c, f2 a) { return __fmul2_rn(make_float2(c, c), a); }
__device__ __forceinline__ f2 fma2s(float c, f2 a, f2 acc) { return __ffma2_rn(make_float2(c, c), a, acc); }
#endif

// ------------------------------------------------------------------------------------------
// error handling: C++ exceptions inside the library, converted to return codes at the C-ABI
struct JudiError : public std::runtime_error {
    explicit JudiError(const std::string &s) : std::runtime_error(s) {}
};

#define CUDA_CHECK(expr)                                                                     \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            char _b[512];                                                                    \
            snprintf(_b, sizeof(_b), "CUDA error %s (%s) at %s:%d", cudaGetErrorName(_e),    \
                     cudaGetErrorString(_e), __FILE__, __LINE__);                            \
            throw JudiError(_b);                                                             \
        }                                                                                    \
    } while (0)

#define JUDI_REQUIRE(cond, msg)                                                              \
    do {                                                                                     \
        if (!(cond)) throw JudiError(std::string(msg) + " [" #cond "]");                     \
    } while (0)

// ------------------------------------------------------------------------------------------
struct judi_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaMemPool_t pool = nullptr;
    bool own_pool = false;
    long long launches = 0;
    // stencil timing (CUDA events on `stream`)
    int timing = 0;             // 0 off, k: time every k-th stencil launch
    long long timing_seq = 0;
    struct Timed { cudaEvent_t e0, e1; int kind; };
    std::vector<Timed> pending;
    std::vector<cudaEvent_t> free_events;
    double stencil_ms = 0;
    long long stencil_launches = 0;
    double stencil_points = 0;
    // the same, split by launch flavour (JUDI_KIND_*)
    double kind_ms[JUDI_NUM_KINDS] = {0};
    long long kind_launches[JUDI_NUM_KINDS] = {0};
    // options
    int opt_kernel3d = 1;       // 0 generic, 1 TMA z-marching
    int opt_zchunks = 0;        // 0 = auto
    int opt_tile = 8;           // TMA kernel variant (8: 64x16 tile, 2 producer warps, shifted queue)
    int opt_atomic_inject = 0;  // 1: atomicAdd injection instead of the sorted segmented path
    int opt_flux_block = 32 | (4 << 8) | (1 << 16);  // thread block of the 3-D flux kernels
    int opt_flux_minb = 8;      // tuning: occupancy target of the 3-D TTI flux kernels (8: 64 registers)
    int opt_flux_slab = 0;      // 3-D flux kernels: z planes per pass-1 / pass-2 slab (0: whole grid)
    int opt_varc_fused = 1;     // 3-D variable density so=8: single-pass kernel (0: two passes)
    int opt_tti_fused = 2;      // 3-D TTI so=8: single-pass kernel (tti3d.cu), 2: shifted queues, 1: rotated; 0: two passes
    int opt_born_fused = 1;     // 3-D Born: one fused launch (0: two launches through HBM)
    int opt_profile = 0;        // print per-phase wall times of J_adjoint to stderr
    int opt_persist2d = 1;      // 2-D: whole time loop in one persistent cooperative kernel
    int opt_persist_blocks = 0; // ... with at most this many CTAs (0: as many tiles as fit)
    int opt_isic_fast = 1;      // 3-D acoustic isic / fwi imaging step: TMA step + vectorised pass (0: generic kernel)
    int opt_ckpt_direct = 1;    // checkpointing: keep as many trailing snapshots in HBM as fit (0: always recompute)
    int opt_persist_vec = 1;    // 2-D acoustic: 4 points per thread; 1: lockstep batches only, 2: single shots too, 0: off
    int num_sms = 148;
    void *encode_tiled = nullptr;  // cuTensorMapEncodeTiled entry point

    void *alloc_bytes(size_t nbytes, bool zero = true);
    float *alloc(size_t nfloat, bool zero = true) {
        return (float *)alloc_bytes(nfloat * sizeof(float), zero);
    }
    void free(void *p);
    void flush_timing();
    // Snapshot arena: ONE buffer that persists between calls and only ever grows.  Tens of GB of
    // snapshots allocated and freed per shot fragment the stream-ordered pool: when the next
    // shot's model arrays carve pieces out of the freed block, the following snapshot request needs
    // a fresh driver allocation (measured: +450 ms per C5 shot, 74 GB of snapshots).
    float *arena_p = nullptr;
    size_t arena_floats = 0;
    float *arena(size_t nfloat);
    // multi-GPU reduction (comm.cu): NCCL communicator of this rank, reduction staging buffer
    void *comm = nullptr;
    int comm_size = 1, comm_rank = 0;
    long long collectives = 0;
    float *red_p = nullptr;
    size_t red_floats = 0;
};

// RAII device buffer on the context's stream-ordered pool
template <typename T>
struct DevArr {
    judi_ctx *ctx = nullptr;
    T *p = nullptr;
    size_t n = 0;
    DevArr() {}
    DevArr(judi_ctx *c, size_t count, bool zero = true) : ctx(c), n(count) {
        p = (T *)c->alloc_bytes(count * sizeof(T), zero);
    }
    DevArr(const DevArr &) = delete;
    DevArr &operator=(const DevArr &) = delete;
    DevArr(DevArr &&o) noexcept : ctx(o.ctx), p(o.p), n(o.n) { o.p = nullptr; }
    DevArr &operator=(DevArr &&o) noexcept {
        if (this != &o) { reset(); ctx = o.ctx; p = o.p; n = o.n; o.p = nullptr; }
        return *this;
    }
    void reset() { if (p && ctx) ctx->free(p); p = nullptr; n = 0; }
    void upload(const std::vector<T> &h) {
        // h is pageable host memory (a std::vector, possibly a temporary): cudaMemcpyAsync copies it
        // to the driver's staging buffer before it returns, so no stream synchronisation is needed
        // (CUDA runtime, "API synchronization behavior": pageable host -> device)
        CUDA_CHECK(cudaMemcpyAsync(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice,
                                   ctx->stream));
    }
    ~DevArr() { reset(); }
};
typedef DevArr<float> DevBuf;

// ------------------------------------------------------------------------------------------
struct judi_model {
    judi_ctx *ctx = nullptr;
    int ndim = 0, so = 0, r = 0, nbl = 0, fs = 0;
    int n[3] = {1, 1, 1};        // physical shape (x,y,z), y == 1 in 2-D
    int N[3] = {1, 1, 1};        // padded shape
    int padl[3] = {0, 0, 0}, padr[3] = {0, 0, 0};
    float d[3] = {1, 1, 1};      // spacing
    float o[3] = {0, 0, 0};      // physical origin
    float opml[3] = {0, 0, 0};   // origin of the padded grid (fp32, models.py:172)
    float dt_user = 0, dt = 0;   // dt = critical_dt actually used
    Grid g;
    Coefs cf;
    bool tti = false, irho_field = false, has_dm = false;
    bool visco = false;          // visco-acoustic SLS (kernels.py:80-141)
    float f0 = 0.015f;           // ... peak frequency (kHz) of the relaxation times
    DevBuf qp, vtt, vts;         // ... quality factor, tt = t_ep/t_s - 1 and t_s per point
    float irho_c = 1.f;
    DevBuf m, irho, eps, delta, theta, phi, dm;  // fields on `g` (edge-replicated into halo)
    // dm as the factor of u.dt2 in the Born source.  With a free surface the reference evaluates the
    // source after the mirror equations (operators.py:188), where u(z=0) is identically zero: the
    // surface row contributes nothing.  dm_src = dm with that row zeroed (dm itself when !fs); the
    // div(dm grad u) part of the isic / fwi sources keeps reading dm.
    DevBuf dm_src_buf;
    const float *dm_src() const { return dm_src_buf.p ? dm_src_buf.p : dm.p; }
    DevBuf r3[3];                // 3-D TTI: symmetry axis (sin th cos ph, sin th sin ph, cos th)
    DevBuf ttiABC[3];            // 3-D TTI: A0 = b delta~, B0 = b sqrt((eps~-delta~) delta~), C0 = b (eps~-delta~)
    DevBuf tti_mi;               // 3-D TTI with a density field: m * b (tile ring of the single-pass kernel)
    DevBuf damp[3];              // 1-D damping profiles per dimension (value of `damp`)
    std::vector<float> damp_h[3];
    Region phys, full;
};

// sparse points of one geometry (sources or receivers) prepared for injection/interpolation
struct SparseSet {
    int np = 0, nc = 0;           // traces, corners per trace
    int ncell = 0, nent = 0;      // unique injected cells, valid (trace, corner) entries
    DevArr<long long> corner_off; // [np*nc] element offsets on Grid, -1 if outside
    DevBuf corner_w;              // [np*nc]
    DevArr<long long> cell_off;   // [ncell] element offset of each unique cell (sorted)
    DevArr<int> cell_start;       // [ncell+1] CSR row pointers into the entries
    DevArr<int> cell_ijk;         // [ncell*3]
    DevArr<int> ent_trace;        // [nent] sorted by (cell, trace, corner)
    DevBuf ent_w;                 // [nent]
    DevBuf cell_scale;            // [ncell] dt^2 / (m*irho)
    DevBuf cell_born;             // [ncell] -dm*irho / (m*irho + dt*damp)   (if model has dm)
    DevArr<int> cellmap;          // 2-D only: [nalloc] row of the cell in the CSR, or -1
};

// optional post-injection corrections applied by the sparse kernel
struct SparseCorr {
    float *ul_next = nullptr;     // Born: ul_next[cell] += cell_born * inj
    int born_scale_m = 0;         // isic/fwi Born source: also multiply by m[cell]
    float *grad = nullptr;        // IC:   grad[reg(cell)] -= gcoef*irho * usnap[reg(cell)] * inj
    const float *usnap = nullptr;
    float gcoef = 0;
    int scale_m = 0;              // isic/fwi: also multiply by m[cell]
    Region reg;
};

// ---- model.cu
Grid make_grid(int ndim, const int *N, int halo);
void model_update_dm_src(judi_model *M);
void fornberg_weights(double z, const double *x, int n, int m, double *c);
void model_build(judi_model *M, judi_ctx *ctx, const judi_model_desc *desc);
void model_set_param(judi_model *M, DevBuf &dst, const judi_param &p, int on_device,
                     float default_value, bool force_field);
void model_set_visco(judi_model *M, const judi_param *qp, float f0, int on_device);
void launch_count(judi_ctx *ctx, const char *what);

// ---- comm.cu
void comm_unique_id(void *id128);
void comm_init(judi_ctx *ctx, int nranks, int rank, const void *id128);
void comm_destroy(judi_ctx *ctx);
void comm_allreduce(judi_ctx *ctx, float *dev, size_t n);
void run_fg_reduce(judi_model *M, const float *grad_padded_dev, const float *f_dev, int true_adjoint,
                   float *grad_out, float *f_out, int flags);

// ---- sparse.cu
void sparse_locate_host(const judi_model *M, int np, const float *coords, std::vector<int> &ijk,
                        std::vector<float> &w, std::vector<int> &valid);
void sparse_build(const judi_model *M, int np, const float *coords, SparseSet &S);
// One launch: inject row `src_row` ([np_src], device) of S_src into field `dst` (+ corrections),
// and interpolate field `cur` at S_rec into rec_row (and `cur2` at S_rec2 into rec_row2).
void sparse_step(const judi_model *M, const SparseSet *S_src, const float *src_row, float *dst,
                 const SparseCorr *corr, const SparseSet *S_rec, const float *cur, float *rec_row,
                 const SparseSet *S_rec2, const float *cur2, float *rec_row2);

// ---- kernels.cu (generic kernels) / tma3d.cu (z-marching TMA kernel)
struct StepArgs {
    const float *u[2];   // level t of field(s)
    float *up[2];        // level t-+1 in, level t+-1 out (in place)
    const float *ul[2];  // Born: linearised field(s), level t
    float *ulp[2];
    int born;            // 1: also step the linearised field(s) with the Born source
    int born_ic;         // imaging condition of the Born source
    Fuse f;
    const CUtensorMap *tmap;       // halo-box tensor map of u[0] (TMA kernel), or nullptr
    const CUtensorMap *tmap_prev;  // tile-box tensor map of up[0]
    const CUtensorMap *tmap_m;     // tile-box tensor map of m
    const CUtensorMap *tmap2;      // fused Born: halo-box map of ul[0], tile-box map of ulp[0] ...
    const CUtensorMap *tmap_prev2;
    const CUtensorMap *tmap_dm;    // ... and tile-box map of dm
    struct Wavefield *scratch; // owner of the flux-form / Born scratch buffers
    struct Wavefield *lin;     // Born: the linearised wavefield (tensor maps of ul)
    const float *qext;         // space-time source field of this step (Grid layout) or nullptr:
    float qext_scale;          //   q = qext_scale * qext  (already multiplied by dt^2 [/ irho])
    float *dd_out;             // TMA kernel, Born background pass: write delta
    const float *qsrc;         // TMA kernel, Born linearised pass: read delta of the background
    const float *sfield;       // ... isic / fwi: dt^2 div(dm grad u) of the background (varc3d_lapw)
    int adj;                   // 1: adjoint sweep (only the visco-acoustic step is not its own mirror image)
};
void stencil_step(const judi_model *M, const StepArgs &a);
// true if a Born step of this model (imaging condition "as") runs on the single-pass flux kernels
bool flux_born_single_pass(const judi_model *M);

// Working set of one propagation: wavefield buffers + scratch for the flux form + tensor maps
struct Wavefield {
    judi_ctx *ctx = nullptr;
    int nf = 1;
    DevBuf b[2][2];            // [field][slot]
    int cur = 0;
    DevBuf flux[6];            // P_x,P_y,P_z,M_x,M_y,M_z (flux form only)
    DevBuf dd[2];              // u+ - 2u + u- of the background field (flux-form Born)
    DevBuf lapw;               // dt^2 div(dm grad u): isic / fwi Born source of the 3-D acoustic kernel
    CUtensorMap tmap_halo_v[2];   // ... haloed boxes of the background field in the varc3d geometry
    CUtensorMap tmap_dmc;         // ... and the coefficient box over dm
    bool has_lapw = false;
    DevBuf rmem, wtmp;         // visco-acoustic: memory variable, adjoint product field
    CUtensorMap tmap_halo[2][2];  // [field][slot]: haloed plane box
    CUtensorMap tmap_tile[2][2];  // [field][slot]: tile box
    CUtensorMap tmap_m;           // tile box over the model's m
    CUtensorMap tmap_dm;          // tile box over the model's dm (Born)
    CUtensorMap tmap_coef;        // variable density: haloed box over irho ...
    CUtensorMap tmap_irho;        // ... and its tile box
    CUtensorMap tmap_par[7];      // 3-D TTI single pass: parameter planes A0, B0, C0, r [, b]
    bool has_tmap = false;
    bool two_pass_only = false;   // flux-form physics: never take the single-pass kernels
    bool generic_only = false;    // flux-form physics: one-point-per-thread kernels only (isic / fwi Born)
    void init(const judi_model *M, bool born_scratch);
    void swap() { cur = 1 - cur; }
    float *level(int f) { return b[f][cur].p; }
    float *other(int f) { return b[f][1 - cur].p; }
};
