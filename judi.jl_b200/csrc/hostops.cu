// hostops.cu -- the steps that bracket every propagator call in JUDI's per-shot driver
// (src/TimeModeling/Modeling/time_modeling_serial.jl:27-79), moved to the device so that the
// gradient crosses PCIe once, already cropped, and shot records are resampled where they live:
//   * remove_padding   (src/TimeModeling/Utils/auxiliaryFunctions.jl:79-89)
//   * time_resample    (src/TimeModeling/Utils/time_utilities.jl:32-48,84)
#include "common.cuh"
#include "propagate.h"

// ---- remove_padding ----------------------------------------------------------------------------
// out[i,j,k] = sum of g over the padded points whose coordinates clamp to (i,j,k) (true_adjoint:
// the adjoint of edge-replication padding), or simply g at the shifted point (crop).
struct UnpadArgs {
    int N[3], n[3], lo[3];
    int fold;
};

__global__ void remove_padding_kernel(UnpadArgs a, const float *__restrict__ g,
                                      float *__restrict__ out)
{
    const long long np = (long long)a.n[0] * a.n[1] * a.n[2];
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= np) return;
    int c[3];
    c[0] = (int)(i % a.n[0]);
    const long long t = i / a.n[0];
    c[1] = (int)(t % a.n[1]);
    c[2] = (int)(t / a.n[1]);
    int b0[3], b1[3];
#pragma unroll
    for (int d = 0; d < 3; d++) {
        b0[d] = b1[d] = c[d] + a.lo[d];
        if (a.fold) {
            // the reference folds the left pad into row nbl and the right pad into row N-nbr-1
            // (both may hit the same row when the physical extent is 1)
            if (c[d] == 0) b0[d] = 0;
            if (c[d] == a.n[d] - 1) b1[d] = a.N[d] - 1;
        }
    }
    float acc = 0.f;
    for (int z = b0[2]; z <= b1[2]; z++)
        for (int y = b0[1]; y <= b1[1]; y++) {
            const float *row = g + (long long)a.N[0] * (y + (long long)a.N[1] * z);
            for (int x = b0[0]; x <= b1[0]; x++) acc += row[x];
        }
    out[i] = acc;
}

void run_remove_padding(const judi_model *M, const float *padded, int true_adjoint, float *out,
                        int flags)
{
    judi_ctx *ctx = M->ctx;
    const bool in_dev = flags & JUDI_DATA_ON_DEVICE, out_dev = flags & JUDI_OUT_ON_DEVICE;
    const size_t ng = (size_t)M->N[0] * M->N[1] * M->N[2];
    const size_t np = (size_t)M->n[0] * M->n[1] * M->n[2];
    DevBuf tin, tout;
    const float *src = padded;
    if (!in_dev) {
        tin = DevBuf(ctx, ng, false);
        CUDA_CHECK(cudaMemcpyAsync(tin.p, padded, ng * sizeof(float), cudaMemcpyHostToDevice,
                                   ctx->stream));
        src = tin.p;
    }
    float *dst = out;
    if (!out_dev) { tout = DevBuf(ctx, np, false); dst = tout.p; }
    UnpadArgs a;
    for (int d = 0; d < 3; d++) { a.N[d] = M->N[d]; a.n[d] = M->n[d]; a.lo[d] = M->padl[d]; }
    a.fold = true_adjoint ? 1 : 0;
    remove_padding_kernel<<<(unsigned)((np + 255) / 256), 256, 0, ctx->stream>>>(a, src, dst);
    launch_count(ctx, "remove_padding");
    if (!out_dev)
        CUDA_CHECK(cudaMemcpyAsync(out, dst, np * sizeof(float), cudaMemcpyDeviceToHost,
                                   ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
}

// ---- time_resample ---------------------------------------------------------------------------
// Traces are time fastest: in[p*nt_in + t].  out[p*nt_out + j] = data at time t0_out + j*dt_out:
//   dt equal            -> shifted copy                        (time_utilities.jl:34-36)
//   dt_out % dt_in == 0 -> decimation                          (time_utilities.jl:37-41)
//   otherwise           -> sinc interpolation  sum_i sinc((t_out_j - t_in_i)/dt_in) in[i]
//                          (time_utilities.jl:84), accumulated in fp32 like the reference's
//                          Float32 matrix product.
struct ResampleArgs {
    int nt_in, nt_out, np;
    float t0_in, dt_in, t0_out, dt_out;
    int mode;     // 0 copy / decimate with integer rate, 1 sinc
    int rate, idx0;
};

// Julia's sinc(x::Float32) = sinpi(x) / (pi * x), 1 at 0
__device__ __forceinline__ float sinc_pi(float x)
{
    if (x == 0.f) return 1.f;
    return __fdiv_rn(sinpif(x), __fmul_rn(3.14159274101257324f, x));
}

template <int TJ>
__global__ void __launch_bounds__(128) time_resample_kernel(ResampleArgs a,
                                                            const float *__restrict__ in,
                                                            float *__restrict__ out)
{
    // one CTA per (trace, block of 128*TJ output samples); the input trace is staged through
    // shared memory in chunks so every input sample is read once per CTA
    __shared__ float sh[1024], sht[1024];
    const int p = blockIdx.y;
    const int j0 = blockIdx.x * (128 * TJ) + threadIdx.x;
    const float *src = in + (size_t)p * a.nt_in;
    if (a.mode == 0) {
#pragma unroll
        for (int k = 0; k < TJ; k++) {
            const int j = j0 + 128 * k;
            if (j < a.nt_out) {
                const long long i = ((long long)j + a.idx0) * a.rate;
                out[(size_t)p * a.nt_out + j] = (i >= 0 && i < a.nt_in) ? src[i] : 0.f;
            }
        }
        return;
    }
    // time axes as Julia's Float32 StepRangeLen yields them (element computed in double, rounded
    // once), then (t_new - t_in) / dt in Float32 exactly as `sinc.((Up .- S') ./ dt)` does
    float tj[TJ], acc[TJ];
#pragma unroll
    for (int k = 0; k < TJ; k++) {
        tj[k] = (float)((double)a.t0_out + (double)(j0 + 128 * k) * (double)a.dt_out);
        acc[k] = 0.f;
    }
    for (int c0 = 0; c0 < a.nt_in; c0 += 1024) {
        const int nc = min(1024, a.nt_in - c0);
        __syncthreads();
        for (int i = threadIdx.x; i < nc; i += 128) {
            sh[i] = src[c0 + i];
            sht[i] = (float)((double)a.t0_in + (double)(c0 + i) * (double)a.dt_in);
        }
        __syncthreads();
        for (int i = 0; i < nc; i++) {
            const float ti = sht[i];
            const float v = sh[i];
#pragma unroll
            for (int k = 0; k < TJ; k++)
                acc[k] = __fmaf_rn(sinc_pi(__fdiv_rn(__fsub_rn(tj[k], ti), a.dt_in)), v, acc[k]);
        }
    }
#pragma unroll
    for (int k = 0; k < TJ; k++) {
        const int j = j0 + 128 * k;
        if (j < a.nt_out) out[(size_t)p * a.nt_out + j] = acc[k];
    }
}

void run_time_resample(judi_ctx *ctx, const float *in, int nt_in, int np, float t0_in, float dt_in,
                       float *out, int nt_out, float t0_out, float dt_out, int flags)
{
    JUDI_REQUIRE(nt_in > 0 && nt_out > 0 && np >= 0 && dt_in > 0.f && dt_out > 0.f,
                 "bad time axes");
    if (np == 0) return;
    const bool in_dev = flags & JUDI_DATA_ON_DEVICE, out_dev = flags & JUDI_OUT_ON_DEVICE;
    DevBuf tin, tout;
    const float *src = in;
    if (!in_dev) {
        tin = DevBuf(ctx, (size_t)nt_in * np, false);
        CUDA_CHECK(cudaMemcpyAsync(tin.p, in, (size_t)nt_in * np * sizeof(float),
                                   cudaMemcpyHostToDevice, ctx->stream));
        src = tin.p;
    }
    float *dst = out;
    if (!out_dev) { tout = DevBuf(ctx, (size_t)nt_out * np, false); dst = tout.p; }
    ResampleArgs a;
    a.nt_in = nt_in; a.nt_out = nt_out; a.np = np;
    a.t0_in = t0_in; a.dt_in = dt_in; a.t0_out = t0_out; a.dt_out = dt_out;
    a.mode = 1; a.rate = 1; a.idx0 = 0;
    // the same case analysis as the reference (exact comparisons on the Float32 steps)
    if (dt_out == dt_in) {
        a.mode = 0;
        a.idx0 = (int)((t0_out - t0_in) / dt_out);   // div(first(t_new) - first(t_in), dt_new)
    } else if (fmodf(dt_out, dt_in) == 0.f) {
        a.mode = 0;
        a.rate = (int)(dt_out / dt_in);
        a.idx0 = (int)((t0_out - t0_in) / dt_out);
    }
    // data[idx1:end] with idx1 < 1 throws in the reference (time_utilities.jl:35,40): a new axis
    // that starts before the data is an error, not an out-of-bounds read
    JUDI_REQUIRE(a.idx0 >= 0, "time_resample: the new time axis starts before the data");
    constexpr int TJ = 4;
    dim3 grd((nt_out + 128 * TJ - 1) / (128 * TJ), np);
    time_resample_kernel<TJ><<<grd, 128, 0, ctx->stream>>>(a, src, dst);
    launch_count(ctx, "time_resample");
    if (!out_dev)
        CUDA_CHECK(cudaMemcpyAsync(out, dst, (size_t)nt_out * np * sizeof(float),
                                   cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
}

// ---- on-the-fly DFT (fields_exprs.py:140-169, sensitivity.py:72-104) ------------------------------
// uf layout: [field][freq][re, im][region slot].  cs: [nfreq][2] = (factor cos wt, -factor sin wt).
template <typename I>
__global__ void dft_accum_kernel(Grid g, Region reg, const float *__restrict__ u,
                                 float *__restrict__ uf, int nfreq, const float *__restrict__ cs)
{
    // I = unsigned when the region has fewer than 2^31 points: 32-bit divisions
    const long long n = (long long)reg.ext[0] * reg.ext[1] * reg.ext[2];
    const long long i64 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i64 >= n) return;
    const I i = (I)i64;
    const int x = reg.lo[0] + (int)(i % (I)reg.ext[0]);
    const I t = i / (I)reg.ext[0];
    const int y = reg.lo[1] + (int)(t % (I)reg.ext[1]), z = reg.lo[2] + (int)(t / (I)reg.ext[1]);
    const float uv = u[g.at(x, y, z)];
    const long long ri = reg.at(x, y, z);
    for (int k = 0; k < nfreq; k++) {
        float *p = uf + (size_t)(2 * k) * reg.slot + ri;
        p[0] = __fmaf_rn(cs[2 * k], uv, p[0]);
        p[reg.slot] = __fmaf_rn(cs[2 * k + 1], uv, p[reg.slot]);
    }
}

// gradm -= sum_f w_f Re(uf e^{i w t}) v;   wc: [nfreq][2] = (w_f cos wt, -w_f sin wt)
template <typename I>
__global__ void dft_grad_kernel(Grid g, Region reg, const float *__restrict__ v,
                                const float *__restrict__ uf, int nfreq,
                                const float *__restrict__ wc, float *__restrict__ grad)
{
    // I = unsigned when the region has fewer than 2^31 points: 32-bit divisions
    const long long n = (long long)reg.ext[0] * reg.ext[1] * reg.ext[2];
    const long long i64 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i64 >= n) return;
    const I i = (I)i64;
    const int x = reg.lo[0] + (int)(i % (I)reg.ext[0]);
    const I t = i / (I)reg.ext[0];
    const int y = reg.lo[1] + (int)(t % (I)reg.ext[1]), z = reg.lo[2] + (int)(t / (I)reg.ext[1]);
    const long long ri = reg.at(x, y, z);
    float acc = 0.f;
    for (int k = 0; k < nfreq; k++) {
        const float *p = uf + (size_t)(2 * k) * reg.slot + ri;
        acc = __fmaf_rn(wc[2 * k], p[0], acc);
        acc = __fmaf_rn(wc[2 * k + 1], p[reg.slot], acc);
    }
    grad[ri] = __fmaf_rn(-acc, v[g.at(x, y, z)], grad[ri]);
}

void dft_accumulate(const judi_model *M, const Region &reg, const float *u, float *uf, int nfreq,
                    const float *cs_dev)
{
    const long long n = (long long)reg.ext[0] * reg.ext[1] * reg.ext[2];
    if (n < (1LL << 31))
        dft_accum_kernel<unsigned><<<(unsigned)((n + 255) / 256), 256, 0, M->ctx->stream>>>(M->g, reg, u, uf, nfreq, cs_dev);
    else
        dft_accum_kernel<long long><<<(unsigned)((n + 255) / 256), 256, 0, M->ctx->stream>>>(M->g, reg, u, uf, nfreq, cs_dev);
    launch_count(M->ctx, "dft_accum");
}

void dft_gradient(const judi_model *M, const Region &reg, const float *v, const float *uf, int nfreq,
                  const float *wc_dev, float *grad)
{
    const long long n = (long long)reg.ext[0] * reg.ext[1] * reg.ext[2];
    if (n < (1LL << 31))
        dft_grad_kernel<unsigned><<<(unsigned)((n + 255) / 256), 256, 0, M->ctx->stream>>>(M->g, reg, v, uf, nfreq, wc_dev, grad);
    else
        dft_grad_kernel<long long><<<(unsigned)((n + 255) / 256), 256, 0, M->ctx->stream>>>(M->g, reg, v, uf, nfreq, wc_dev, grad);
    launch_count(M->ctx, "dft_grad");
}

// isic / fwi in the frequency domain (sensitivity.py:125-154), on the steps with t % factor == 0:
//   gradm -= Re( w_f U v m - w2 inner_grad(U, v) ),  U = uf e^{i w t},  w2 = +-factor / time_M.
// Pass 1: gradm -= m v sum_f w_f Re(U_f) and B = sum_f Re(U_f) (w2 does not depend on f);
// pass 2: gradm += w2 inner_grad(B, v) with B read through its region (zero outside).
// ec: [nfreq][2] = (cos wt, -sin wt).
template <typename I>
__global__ void dft_grad_isic1_kernel(Grid g, Region reg, const float *__restrict__ v,
                                      const float *__restrict__ m, const float *__restrict__ uf, int nfreq,
                                      const float *__restrict__ wc, const float *__restrict__ ec,
                                      float *__restrict__ B, float *__restrict__ grad)
{
    // I = unsigned when the region has fewer than 2^31 points: 32-bit divisions
    const long long n = (long long)reg.ext[0] * reg.ext[1] * reg.ext[2];
    const long long i64 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i64 >= n) return;
    const I i = (I)i64;
    const int x = reg.lo[0] + (int)(i % (I)reg.ext[0]);
    const I t = i / (I)reg.ext[0];
    const int y = reg.lo[1] + (int)(t % (I)reg.ext[1]), z = reg.lo[2] + (int)(t / (I)reg.ext[1]);
    const long long ri = reg.at(x, y, z), gi = g.at(x, y, z);
    float aw = 0.f, ab = 0.f;
    for (int k = 0; k < nfreq; k++) {
        const float *p = uf + (size_t)(2 * k) * reg.slot + ri;
        const float re = p[0], im = p[reg.slot];
        aw = __fmaf_rn(wc[2 * k], re, aw);
        aw = __fmaf_rn(wc[2 * k + 1], im, aw);
        ab = __fmaf_rn(ec[2 * k], re, ab);
        ab = __fmaf_rn(ec[2 * k + 1], im, ab);
    }
    B[ri] = ab;
    grad[ri] = __fmaf_rn(-aw * m[gi], v[gi], grad[ri]);
}

template <typename I>
__global__ void dft_grad_isic2_kernel(Grid g, Coefs cf, int R, int ndim, Region reg,
                                      const float *__restrict__ v, const float *__restrict__ B, float w2,
                                      float *__restrict__ grad)
{
    // I = unsigned when the region has fewer than 2^31 points: 32-bit divisions
    const long long n = (long long)reg.ext[0] * reg.ext[1] * reg.ext[2];
    const long long i64 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i64 >= n) return;
    const I i = (I)i64;
    const int x = reg.lo[0] + (int)(i % (I)reg.ext[0]);
    const I t = i / (I)reg.ext[0];
    const int y = reg.lo[1] + (int)(t % (I)reg.ext[1]), z = reg.lo[2] + (int)(t / (I)reg.ext[1]);
    const long long gi = g.at(x, y, z);
    const long long st[3] = {1, g.sy, g.sz};
    const float *ws[3] = {cf.wx, cf.wy, cf.wz};
    float s = 0.f;
    for (int d = 0; d < 3; d++) {
        if (d == 1 && ndim == 2) continue;
        float gb = 0.f, gv = 0.f;
        for (int k = 1; k <= R; k++) {
            int p[3] = {x, y, z}, q[3] = {x, y, z};
            p[d] += k; q[d] -= k - 1;
            const float bp = reg.has(p[0], p[1], p[2]) ? B[reg.at(p[0], p[1], p[2])] : 0.f;
            const float bm = reg.has(q[0], q[1], q[2]) ? B[reg.at(q[0], q[1], q[2])] : 0.f;
            gb += ws[d][k] * (bp - bm);
            gv += ws[d][k] * (v[gi + k * st[d]] - v[gi - (k - 1) * st[d]]);
        }
        s += gb * gv;
    }
    const long long ri = reg.at(x, y, z);
    grad[ri] = __fmaf_rn(w2, s, grad[ri]);
}

void dft_gradient_isic(const judi_model *M, const Region &reg, const float *v, const float *uf, int nfreq,
                       const float *wc_dev, const float *ec_dev, float w2, float *B, float *grad)
{
    const long long n = (long long)reg.ext[0] * reg.ext[1] * reg.ext[2];
    const unsigned nb = (unsigned)((n + 255) / 256);
    if (n < (1LL << 31))
        dft_grad_isic1_kernel<unsigned><<<nb, 256, 0, M->ctx->stream>>>(M->g, reg, v, M->m.p, uf, nfreq, wc_dev, ec_dev, B, grad);
    else
        dft_grad_isic1_kernel<long long><<<nb, 256, 0, M->ctx->stream>>>(M->g, reg, v, M->m.p, uf, nfreq, wc_dev, ec_dev, B, grad);
    launch_count(M->ctx, "dft_grad_isic1");
    if (n < (1LL << 31))
        dft_grad_isic2_kernel<unsigned><<<nb, 256, 0, M->ctx->stream>>>(M->g, M->cf, M->r, M->ndim, reg, v, B, w2, grad);
    else
        dft_grad_isic2_kernel<long long><<<nb, 256, 0, M->ctx->stream>>>(M->g, M->cf, M->r, M->ndim, reg, v, B, w2, grad);
    launch_count(M->ctx, "dft_grad_isic2");
}
