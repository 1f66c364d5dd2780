/*
 * judi_b200.h -- C-ABI of libjudi_b200.so: B200-native (sm_100a) finite-difference wave
 * propagators that replace JUDI's `src/pysource` + Devito hot path.
 *
 * Every entry point below replaces one reference interface (paths relative to the reference
 * repository slimgroup/JUDI.jl v4.1.7).  The Julia driver reaches the reference through
 * `wrapcall_data(ac.<fn>, modelPy, ...)` (src/TimeModeling/Modeling/python_interface.jl:9-27);
 * the binding a maintainer adds instead is a `ccall` per function -- see INTEGRATION.md.
 *
 * Conventions (chosen so Julia arrays are passed as they are, without copies):
 *   - grids (m, rho, dm, gradients, illumination): fp32, x fastest (Julia column-major
 *     Array{Float32,N} of size n / padded size), i.e. element (ix,iy,iz) at ix + nx*(iy + ny*iz);
 *   - coordinates: `hcat(x[,y],z)` column-major (auxiliaryFunctions.jl:319-321), i.e. SoA
 *     coords[d*npoint + p], absolute metres, fp32;
 *   - traces (wavelets, shot records, residuals): (nt x ntrace) column-major, i.e. time fastest:
 *     data[p*nt + t];
 *   - all functions return 0 on success, non-zero on failure with the message available from
 *     judi_last_error() (thread local).  The Julia wrapper turns non-zero into `error(...)`, so the
 *     reference's `retry` wrapper (time_modeling_serial.jl:95, misfit_fg.jl:91) keeps working.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef JUDI_B200_H
#define JUDI_B200_H

#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

#define JUDI_B200_ABI_VERSION 1

typedef struct judi_ctx judi_ctx;     /* one per (process, GPU): stream, memory pool          */
typedef struct judi_model judi_model; /* padded physical model resident in HBM                */

/* how a physical parameter is given (pysource Model accepts array | scalar | None,
 * src/pysource/models.py:301-326) */
enum { JUDI_PARAM_NONE = 0, JUDI_PARAM_SCALAR = 1, JUDI_PARAM_ARRAY = 2 };
typedef struct {
    int kind;          /* JUDI_PARAM_* */
    float value;       /* used when kind == JUDI_PARAM_SCALAR */
    const float *data; /* used when kind == JUDI_PARAM_ARRAY: prod(shape) floats, x fastest */
} judi_param;

/* replaces `pm.Model(origin, spacing, shape; nbl, space_order, dt, fs, dm, m=, rho=, ...)`
 * (src/TimeModeling/Utils/auxiliaryFunctions.jl:26-37 -> src/pysource/models.py:157-248) */
typedef struct {
    int ndim;         /* 2 or 3 */
    int shape[3];     /* n: (nx, nz) or (nx, ny, nz) */
    float spacing[3]; /* d, metres */
    float origin[3];  /* o, metres */
    int nbl;          /* absorbing layer width in cells */
    int space_order;  /* 4, 8 or 16 */
    int fs;           /* Options(free_surface): no layer above z=0 + mirror (fields_exprs.py:106) */
    float dt;         /* user dt_comp (ms); <= 0 means "use the CFL value" */
    judi_param m;     /* squared slowness s^2/km^2 (required) */
    judi_param rho;   /* density; scalar -> constant-density Laplacian, array -> div(1/rho grad) */
    judi_param b;     /* buoyancy 1/rho (alternative to rho) */
    judi_param epsilon, delta, theta, phi; /* Thomsen / tilt / azimuth: any present => TTI */
    judi_param dm;    /* squared-slowness perturbation for Born modelling */
    int params_on_device; /* 1: the `data` pointers are device pointers (inputs resident in HBM) */
} judi_model_desc;

/* imaging condition, src/pysource/sensitivity.py:231-233 */
enum { JUDI_IC_AS = 0, JUDI_IC_ISIC = 1, JUDI_IC_FWI = 2 };

/* flags of the propagation calls */
enum {
    JUDI_DATA_ON_DEVICE = 1, /* wavelet / recin / rec_out / grad_out / illum pointers are device
                                pointers (same layouts); coordinates are always host pointers */
    JUDI_GRAD_ACCUMULATE = 2, /* grad_out += g and *f_out += f instead of overwriting: the local
                                reduction of src/TimeModeling/Modeling/distributed.jl:6-22 */
    JUDI_OUT_ON_DEVICE = 4   /* J_adjoint: f_out / grad_out / illum are device pointers although the
                                traces are host arrays: a worker sums its shots in HBM and copies
                                the gradient back once, after the all-reduce */
};

/* where forward-wavefield snapshots are kept for the imaging condition */
enum {
    JUDI_SNAP_PADDED = 0,  /* whole padded grid: reference semantics, gradient defined in the
                              absorbing layer too (needed by Options(sum_padding=true)) */
    JUDI_SNAP_PHYSICAL = 1 /* physical domain only: the returned gradient is zero in the layer,
                              which `remove_padding` crops anyway (auxiliaryFunctions.jl:79-89) */
};

/* options of J_adjoint: the keyword arguments of src/pysource/interface.py:279-282 that this
 * path supports */
typedef struct {
    int is_residual;   /* recin is already the residual (J' as an operator, python_interface.jl:119) */
    int checkpointing; /* Options(optimal_checkpointing): recompute segments from HBM checkpoints */
    int n_checkpoints; /* <= 0: choose from free HBM */
    int t_sub;         /* Options(subsampling_factor) */
    int ic;            /* JUDI_IC_* */
    int born_fwd;      /* lsrtm_objective: linearised forward (interface.py:565) */
    int nlind;         /* remove the non-linear data (only with born_fwd) */
    int illum;         /* also return illumination of u and v */
    int snapshot_region; /* JUDI_SNAP_* */
    const float *ws;   /* extended source weights (padded grid) instead of point sources, or NULL:
                          J_adjoint(ws=...) of python_interface.jl:179-198 */
    /* on-the-fly DFT gradient (J_adjoint_freq, interface.py:348-408): nfreq > 0 replaces the
     * wavefield snapshots by nfreq complex frequency slices accumulated every dft_sub steps
     * (fields_exprs.py:140-169) and the imaging condition by sensitivity.py:72-104, which the
     * reference applies on every adjoint step whatever dft_sub is (grad_expr, sensitivity.py:50);
     * with ic = isic / fwi by sensitivity.py:125-154, applied every dft_sub steps.
     * freq_list: host array of nfreq frequencies in kHz. */
    int nfreq;
    int dft_sub;       /* DFT time subsampling factor (>= 1) */
    const float *freq_list;
} judi_jopts;

/* optional host misfit callback, replaces `misfit=pyjl(runtime_misfit)`
 * (src/TimeModeling/Modeling/misfit_fg.jl:55-64, src/pysource/sensitivity.py:255-263):
 * given synthetic and observed data (nt x nrec, time fastest) it writes the adjoint source into
 * `residual` and returns the objective value (the library multiplies it by dt like the reference).
 * NULL selects the on-device L2 misfit of src/TimeModeling/Modeling/losses.jl:15-19. */
typedef float (*judi_misfit_fn)(const float *dsyn, const float *dobs, float *residual, int nt,
                                int nrec, void *user);

/* ---- context --------------------------------------------------------------------------- */
int judi_abi_version(void);
const char *judi_last_error(void);
judi_ctx *judi_ctx_create(int device);
void judi_ctx_destroy(judi_ctx *ctx);
int judi_ctx_synchronize(judi_ctx *ctx);
void *judi_ctx_stream(judi_ctx *ctx); /* the cudaStream_t all work of this context runs on */
/* kernels launched by this context since creation (bench.py's gpu_launches) */
long long judi_ctx_launch_count(judi_ctx *ctx);
/* accumulated device time (ms, CUDA events on the context stream) of the stencil kernels, and
 * the number of stencil launches and grid-point updates they covered; reset clears them */
int judi_ctx_stencil_stats(judi_ctx *ctx, double *ms, long long *launches, double *points);
/* the same split by the flavour of the time step (what was fused into the stencil launch) */
#define JUDI_KIND_PLAIN 0 /* leapfrog step only (+ illumination) */
#define JUDI_KIND_SAVE 1  /* + snapshot save (forward sweep of a gradient) */
#define JUDI_KIND_IC 2    /* + imaging condition (adjoint sweep of a gradient) */
#define JUDI_KIND_BORN 3  /* background + linearised field */
#define JUDI_NUM_KINDS 4
int judi_ctx_stencil_stats_kind(judi_ctx *ctx, int kind, double *ms, long long *launches);
int judi_ctx_reset_stats(judi_ctx *ctx);
/* tuning knobs: "acoustic3d_kernel" (0 generic, 1 TMA z-marching), "zchunks", "tile", ...; the
 * same "key=value,key=value" list is read from the environment variable JUDI_B200_OPTIONS when a
 * context is created (judi_ctx_create fails on an unknown key) */
int judi_ctx_set_option(judi_ctx *ctx, const char *key, int value);
/* Page-lock a caller-owned host array (a Julia Array that outlives many shots: the model's m, the
 * observed data, the gradient the per-shot results are summed into) so that the copies of every
 * later call run at PCIe speed instead of through the driver's pageable staging (~3.5 GB/s
 * measured: 75 ms for one padded C3 gradient).  Replaces nothing in the reference -- PyCall hands
 * numpy views of Julia memory to Devito, which runs on the host -- it is what the `ccall` binding
 * does once per array in place of that zero-copy wrap (python_interface.jl:9-27).
 * Registering an already registered range is not an error. */
int judi_host_register(void *ptr, size_t nbytes);
int judi_host_unregister(void *ptr);

/* Give the memory this context caches (snapshot arena, reduction buffer, freed pool blocks) back
 * to the driver, e.g. before another library of the host process needs the HBM.  The context
 * allocates from a private stream-ordered pool, never from the device's default one. */
int judi_ctx_trim(judi_ctx *ctx);

/* ---- multi-GPU reduction ---------------------------------------------------------------
 * Replaces `reduce!` / `run_and_reduce` (src/TimeModeling/Modeling/distributed.jl:60-95,
 * propagation.jl:23-50): the binary tree over Julia Futures that sums (objective, gradient) of
 * the workers.  One process per GPU; rank 0 calls judi_comm_unique_id and hands the 128 bytes
 * to every rank over the host's own channel (Distributed.jl / MPI / a TCP store); every rank
 * then calls judi_comm_init on its context.  NCCL is loaded at run time (dlopen of
 * $JUDI_NCCL_LIB, else libnccl.so.2); nranks == 1 needs no NCCL at all. */
#define JUDI_COMM_ID_BYTES 128
int judi_comm_unique_id(void *id128);
int judi_comm_init(judi_ctx *ctx, int nranks, int rank, const void *id128);
int judi_comm_destroy(judi_ctx *ctx);
int judi_comm_size(judi_ctx *ctx);
int judi_comm_rank(judi_ctx *ctx);
/* in-place sum over the ranks of n floats in HBM, on the context stream (ncclAllReduce) */
int judi_comm_allreduce(judi_ctx *ctx, float *dev, size_t n);

/* ---- model ----------------------------------------------------------------------------- */
judi_model *judi_model_create(judi_ctx *ctx, const judi_model_desc *desc);
void judi_model_destroy(judi_model *model);
/* modelPy.critical_dt (models.py:466-482; read at python_interface.jl:33, misfit_fg.jl:40) */
float judi_model_critical_dt(const judi_model *model);
/* modelPy.padsizes (models.py:269-272): out[2*d] = left, out[2*d+1] = right */
int judi_model_padsizes(const judi_model *model, int *out);
/* padded grid shape (Devito grid.shape, models.py:173) */
int judi_model_padded_shape(const judi_model *model, int *out);
/* replace the perturbation dm of an existing model (born / lsrtm with a new dm) */
int judi_model_set_dm(judi_model *model, const judi_param *dm, int on_device);
/* make the model visco-acoustic (`Model(n, d, o, m, rho, qp)`, src/pysource/models.py:208-214 and
 * kernels.py:80-141): quality factor field qp (NULL: keep the current one) and the peak frequency
 * f0 in kHz that the reference passes with every call (`f0=options.f0`, python_interface.jl:42;
 * propagators.py:46).  Call again to change f0 between calls. */
int judi_model_set_visco(judi_model *model, const judi_param *qp, float f0, int on_device);
/* the padded damping field (models.py:54-120) and padded m, for tests: out is padded-grid sized */
int judi_model_get_field(const judi_model *model, const char *name, float *out);

/* ---- sparse geometry: bit-exact indexing contract ---------------------------------------- */
/* corner cells and weights of Devito's linear SparseTimeFunction for `npoint` coordinates
 * (src/pysource/geom_utils.py:74-94).  ijk: [npoint][2^ndim][3] ints (x,y,z; y = 0 in 2-D),
 * w: [npoint][2^ndim] floats, valid: [npoint][2^ndim] ints (0 = corner outside the padded grid) */
int judi_sparse_locate(const judi_model *model, int npoint, const float *coords, int *ijk,
                       float *w, int *valid);

/* ---- propagators (src/pysource/interface.py) --------------------------------------------- */
/* forward_rec (interface.py:17-47; called at python_interface.jl:42): rec = Pr*F*Ps'*q, or the
 * adjoint F' with fw = 0.  illum_out (padded grid) may be NULL. */
int judi_forward_rec(judi_model *model, int fw, int nsrc, const float *src_coords,
                     const float *wavelet, int nt, int nrec, const float *rec_coords,
                     float *rec_out, float *illum_out, int flags);

/* born_rec (interface.py:208-241; python_interface.jl:97): d_lin = J*dm with the model's dm */
int judi_born_rec(judi_model *model, int nsrc, const float *src_coords, const float *wavelet,
                  int nt, int nrec, const float *rec_coords, int ic, float *rec_out,
                  float *illum_out, int flags);

/* J_adjoint (interface.py:279-345 -> :411-470 standard, :473-562 checkpointing; called at
 * python_interface.jl:116-120 and misfit_fg.jl:69-73).  f_out may be NULL (return_obj=false).
 * grad_out / illum_u / illum_v are padded-grid sized (the caller crops with remove_padding). */
int judi_J_adjoint(judi_model *model, int nsrc, const float *src_coords, const float *wavelet,
                   int nt, int nrec, const float *rec_coords, const float *recin,
                   const judi_jopts *opts, judi_misfit_fn misfit, void *user, float *f_out,
                   float *grad_out, float *illum_u, float *illum_v, int flags);

/* One source experiment of a multi-shot objective: geometry, wavelet and observed data (or residual)
 * as the per-shot arguments of judi_J_adjoint. */
typedef struct {
    int nsrc;
    const float *src_coords; /* nsrc x ndim */
    const float *wavelet;    /* nt x nsrc, time fastest */
    int nt;
    int nrec;
    const float *rec_coords; /* nrec x ndim */
    const float *recin;      /* nt x nrec, time fastest */
} judi_shot;

/* Objective and gradient of `nshots` source experiments on this context's GPU, summed in HBM:
 * the per-worker part of multi_src_fg! (src/TimeModeling/Modeling/propagation.jl:93-107) with the
 * local reduction of distributed.jl:6-22 -- one call per worker instead of one per shot, one copy
 * of the gradient back instead of one per shot.  f_out / grad_out (padded grid) receive the sums;
 * with JUDI_OUT_ON_DEVICE they are device pointers, so the caller can all-reduce over GPUs (NCCL)
 * before anything crosses PCIe; JUDI_GRAD_ACCUMULATE adds to their current content.  The shots'
 * own pointers are host pointers unless JUDI_DATA_ON_DEVICE is set. */
int judi_fg_multishot(judi_model *model, int nshots, const judi_shot *shots, const judi_jopts *opts,
                      judi_misfit_fn misfit, void *user, float *f_out, float *grad_out, int flags);

/* forward_no_rec (interface.py:83-111; python_interface.jl:56): the full wavefield history,
 * u_out is nt x padded grid (time slowest).  TTI returns the first field like the reference. */
int judi_forward_no_rec(judi_model *model, int fw, int nsrc, const float *src_coords,
                        const float *wavelet, int nt, float *u_out, float *illum_out, int flags);

/* ---- extended and wavefield sources -------------------------------------- */
/* forward_rec_w (interface.py:50-80; python_interface.jl:139): rec = Pr*F*Pw'*w.  `weight` is the
 * spatial source distribution on the PADDED grid (Julia zero-pads it, python_interface.jl:129),
 * `wavelet` its (nt x 1) time signature. */
int judi_forward_rec_w(judi_model *model, int fw, const float *weight, const float *wavelet, int nt,
                       int nrec, const float *rec_coords, float *rec_out, float *illum_out,
                       int flags);

/* adjoint_w (interface.py:174-204; python_interface.jl:156): w = Pw*F'*Pr'*d.  The data are
 * injected at the receivers, the wavefield is correlated with `wavelet`; w_out is padded-grid
 * sized. */
int judi_adjoint_w(judi_model *model, int fw, int nrec, const float *rec_coords, const float *data,
                   const float *wavelet, int nt, float *w_out, float *illum_out, int flags);

/* born_rec_w (interface.py:244-276; python_interface.jl:174) */
int judi_born_rec_w(judi_model *model, const float *weight, const float *wavelet, int nt, int nrec,
                    const float *rec_coords, int ic, float *rec_out, float *illum_out, int flags);

/* forward_wf_src (interface.py:114-142; python_interface.jl:69): rec = Pr*F*u with a full
 * space-time source u_src (nt x padded grid, time slowest) */
int judi_forward_wf_src(judi_model *model, int fw, const float *u_src, int nt, int nrec,
                        const float *rec_coords, float *rec_out, float *illum_out, int flags);

/* forward_wf_src_norec (interface.py:145-171; python_interface.jl:79): u_out = F*u_src */
int judi_forward_wf_src_norec(judi_model *model, int fw, const float *u_src, int nt, float *u_out,
                              float *illum_out, int flags);

/* ---- the steps around every propagator call ----------------------------- */
/* remove_padding (src/TimeModeling/Utils/auxiliaryFunctions.jl:79-89; called at
 * time_modeling_serial.jl:70-71, misfit_fg.jl:77): crop a padded-grid array to the physical
 * grid; true_adjoint != 0 first sums every pad cell into the nearest edge cell (the adjoint of
 * edge-replication padding, `sum_padding=true`).  flags: JUDI_DATA_ON_DEVICE = `padded` is a
 * device pointer, JUDI_OUT_ON_DEVICE = `out` is a device pointer. */
int judi_remove_padding(const judi_model *model, const float *padded, int true_adjoint, float *out,
                        int flags);

/* time_resample (src/TimeModeling/Utils/time_utilities.jl:32-48, SincInterpolation :84; called
 * at python_interface.jl:33-39 on the wavelet / data and time_modeling_serial.jl:78-79 on the
 * shot record): traces are time fastest, in = [ntrace][nt_in] sampled at t0_in + i*dt_in,
 * out = [ntrace][nt_out] at t0_out + j*dt_out.  Equal steps copy, integer ratios decimate,
 * everything else is sinc interpolation.  Flags as for judi_remove_padding. */
int judi_time_resample(judi_ctx *ctx, const float *in, int nt_in, int ntrace, float t0_in,
                       float dt_in, float *out, int nt_out, float t0_out, float dt_out, int flags);

/* The reduction + post-processing of multi_src_fg! (propagation.jl:99-106, distributed.jl:60-95,
 * misfit_fg.jl:77) in one call: `grad_padded_dev` (padded grid, device) and `f_dev` (device scalar
 * or NULL) are this rank's accumulators as filled by judi_J_adjoint / judi_fg_multishot with
 * JUDI_OUT_ON_DEVICE | JUDI_GRAD_ACCUMULATE.  The gradient is cropped (true_adjoint: pad folded
 * into the edges) on the device, the objective appended, and ONE all-reduce of prod(n)+1 floats
 * sums them over the ranks of the context's communicator (none: the local values); the result is
 * written to `grad_out` (physical grid, x fastest) and `f_out` -- host pointers, or device
 * pointers with JUDI_OUT_ON_DEVICE. */
int judi_fg_reduce(judi_model *model, const float *grad_padded_dev, const float *f_dev,
                   int true_adjoint, float *grad_out, float *f_out, int flags);

/* multi_src_fg! + reduce! for one worker in ONE call (SURVEY 8b `judi_fg_multishot`): this
 * rank's `nshots` experiments (0 is allowed: the rank still takes part in the reduction) are
 * summed in HBM as by judi_fg_multishot, then reduced over the communicator as by
 * judi_fg_reduce.  Every rank receives the gradient on the PHYSICAL grid and the objective, summed
 * over all shots of all ranks (host pointers, or device pointers with JUDI_OUT_ON_DEVICE). */
int judi_fg_multishot_reduced(judi_model *model, int nshots, const judi_shot *shots,
                              const judi_jopts *opts, judi_misfit_fn misfit, void *user,
                              int true_adjoint, float *f_out, float *grad_out, int flags);

#ifdef __cplusplus
}
#endif
#endif /* JUDI_B200_H */
