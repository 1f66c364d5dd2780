// propagate.h -- per-shot drivers (see propagate.cu)
#pragma once
#include "common.cuh"

// device-resident traces, one row per time sample (trace fastest)
struct Traces {
    DevBuf buf;
    int nt = 0, np = 0;
    float *row(int t) const { return buf.p + (size_t)t * np; }
    void alloc(judi_ctx *ctx, int nt, int np);
    void load(judi_ctx *ctx, const float *user, int nt, int np, bool on_device);
    void store(judi_ctx *ctx, float *user, bool on_device) const;
};

// optional space-time sources / extended receiver of one call (user pointers, see judi_b200.h)
struct ExtSpec {
    const float *ws = nullptr;          // extended source weights, padded grid   (interface.py:50)
    const float *wr_wavelet = nullptr;  // extended receiver: wavelet to correlate with ...
    float *w_out = nullptr;             // ... and the padded-grid output           (interface.py:174)
    const float *qwf = nullptr;         // full wavefield source, nt x padded grid  (interface.py:114)
};

struct Shot {
    judi_model *M;
    judi_ctx *ctx;
    // device-side state of the ExtSpec of the current call
    DevBuf ws_grid, qgrid, wrec;
    std::vector<float> wt_host, wrt_host;
    DevBuf wt_dev, wrt_dev;    // the same time series on the device (persistent 2-D kernel)
    const float *qwf_dev = nullptr;
    DevBuf qwf_owned;
    bool ext_on = true;
    bool adjoint = false;      // direction of the sweep being stepped (visco-acoustic steps differ)
    bool ext_active() const { return ext_on && (ws_grid.p || qwf_dev || wrec.p); }
    void bind_ext(const ExtSpec *ext, const float *wavelet, int nt, bool on_device);
    void apply_ext(StepArgs &s, int t);
    void finish_ext(const ExtSpec *ext, bool on_device);
    explicit Shot(judi_model *M);
    void make_args(StepArgs &s, Wavefield &W);
    void step_plain(Wavefield &W, int t, const Fuse *fuse, const SparseSet *S_in, const Traces *in,
                    const SparseCorr *corr, const SparseSet *S_out, Traces *out);
    void step_born(Wavefield &W, Wavefield &WL, int t, int ic, const Fuse *fuse,
                   const SparseSet *S_src, const Traces *q, const SparseSet *S_rec, Traces *recl,
                   Traces *recnl);
};

// fused extras of the persistent 2-D kernel (persist2d.cu)
struct P2Fuse {
    float *snap = nullptr;        // snapshot pool base (save) ...
    const float *usnap = nullptr; // ... or (imaging condition) the pool to read
    float *grad = nullptr, *illum = nullptr;
    long long slot = 0;           // floats per snapshot slot
    int t_sub = 1, t_base = 0;    // slot of time index t is (t - t_base) / t_sub
    float gcoef = 0;
    int ic = JUDI_IC_AS;          // imaging condition of the gradient / Born source
    Region reg;
    // space-time sources / extended receiver (fields_exprs.py:58-103), constant-density kernel only
    const float *qext = nullptr, *qwt = nullptr;   // q = qbase * qwt[t] * qext(x)   (Grid layout)
    const float *qwf = nullptr;                    // q = qbase * qwf[t](x)          (padded layout, nt slices)
    float qbase = 0;
    float *wrec = nullptr;                         // wrec += wdt * wrt[t] * u[t]    (region `reg`)
    const float *wrt = nullptr;
    float wdt = 0;
};
struct Shot;
// fills the extended-source members of fz from the shot's bound ExtSpec; false when the kernel cannot take it
bool persist2d_bind_ext(const judi_model *M, const Shot &sh, P2Fuse &fz);
bool persist2d_supported(const judi_model *M, int ic, bool born = false);
void persist2d_run(const judi_model *M, Wavefield &W, Wavefield *WL, int t0, int t1, int dir,
                   const SparseSet *S_in, const Traces *in, const SparseSet *S_out, Traces *out,
                   Traces *out_bg, const P2Fuse &fz);

bool persist2d_batch_supported(const judi_model *M, int ic);
void persist2d_run_batch(const judi_model *M, int n, Wavefield *const *W, int t0, int t1, int dir,
                         const SparseSet *const *S_in, const Traces *const *in,
                         const SparseSet *const *S_out, Traces *const *out, const P2Fuse *fz,
                         Wavefield *const *WL = nullptr, Traces *const *out_bg = nullptr);
// judi_fg_multishot fast path: all shots of a 2-D constant-density model in lockstep; false if the
// combination is not covered (the caller then loops over run_J_adjoint)
bool run_fg_batch2d(judi_model *M, int nshots, const judi_shot *shots, const judi_jopts *o,
                    float *f_dev, float *g_dev, int flags);

void region_output(const judi_model *M, const Region &reg, const float *src, float *user,
                   bool on_device, bool accumulate);
void grid_output(const judi_model *M, const float *field, float *user, bool on_device);

void run_forward_rec(judi_model *M, int fw, int nsrc, const float *src_coords,
                     const float *wavelet, int nt, int nrec, const float *rec_coords,
                     float *rec_out, float *illum_out, int flags, const ExtSpec *ext = nullptr);
void run_forward_no_rec(judi_model *M, int fw, int nsrc, const float *src_coords,
                        const float *wavelet, int nt, float *u_out, float *illum_out, int flags,
                        const ExtSpec *ext = nullptr);
void run_born_rec(judi_model *M, int nsrc, const float *src_coords, const float *wavelet, int nt,
                  int nrec, const float *rec_coords, int ic, float *rec_out, float *illum_out,
                  int flags, const ExtSpec *ext = nullptr);
void run_J_adjoint(judi_model *M, int nsrc, const float *src_coords, const float *wavelet, int nt,
                   int nrec, const float *rec_coords, const float *recin, const judi_jopts *o,
                   judi_misfit_fn misfit, void *user, float *f_out, float *grad_out,
                   float *illum_u, float *illum_v, int flags);

// hostops.cu
void run_remove_padding(const judi_model *M, const float *padded, int true_adjoint, float *out,
                        int flags);
void run_time_resample(judi_ctx *ctx, const float *in, int nt_in, int np, float t0_in, float dt_in,
                       float *out, int nt_out, float t0_out, float dt_out, int flags);
void dft_accumulate(const judi_model *M, const Region &reg, const float *u, float *uf, int nfreq,
                    const float *cs_dev);
void dft_gradient(const judi_model *M, const Region &reg, const float *v, const float *uf, int nfreq,
                  const float *wc_dev, float *grad);
void dft_gradient_isic(const judi_model *M, const Region &reg, const float *v, const float *uf, int nfreq,
                       const float *wc_dev, const float *ec_dev, float w2, float *B, float *grad);
