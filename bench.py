#!/usr/bin/env python
"""bench.py -- FWI-gradient throughput of the B200 backend on BASELINE.json's headline workload.

Workload (config.workload): 3-D isotropic acoustic, space order 8, FWI gradient
(`J_adjoint(..., return_obj=True)` = forward with snapshots + adjoint with the fused imaging
condition), SURVEY.md 8(d) config C3's grid: n = 401x401x201 @ 25 m (padded 481x481x281 =
65.0 M points), 101x101 receivers, T = 2 s (nt = 1042), subsampling_factor 4.
A "step" is ONE shot per GPU (weak scaling: N GPUs -> N shots per step) followed by the
all-reduce of (gradient, objective) over the ranks -- the only collective of the path.

  value  : shots/s with model, wavelet and observed data resident in HBM before the timed region
  e2e    : the same through the host-buffer call a JUDI worker makes per shot
           (Model(...) from host arrays -> J_adjoint with host arrays -> gradient back on host)
  roofline: the stencil kernel (acoustic3d_tma), algorithmic bytes / CUDA-event time per launch
  cpu_baseline / --impl reference: the CPU oracle (a port of the reference's Devito schedule --
           Devito itself cannot be installed here) on the host cores, bounded sample.

Launch: `python bench.py --gpus 1 --steps K --warmup W`, or for N > 1
`python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...`.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SHAPE = (401, 401, 201)
SPACING = (25., 25., 25.)
NBL = 40
SO = 8
T_MS = 2000.
F0 = 0.008
DT = 1.921          # critical dt of the true model (vmax 5.5 km/s), passed as dt_comp
T_SUB = 4


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--t-sub", type=int, default=T_SUB)
    ap.add_argument("--shape", type=int, nargs=3, default=list(SHAPE))
    ap.add_argument("--tn", type=float, default=T_MS)
    ap.add_argument("--snapshot-region", default="physical", choices=["physical", "padded"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--workload", default="c3", choices=["c3", "c5", "c3fwd", "c4"],
                    help="c3 (default): the 3-D acoustic FWI gradient BASELINE.json's metric names; "
                         "c5: 3-D TTI forward + gradient on 301x301x201; c3fwd: forward modelling "
                         "of shots on the C3 grid, no collective; c4: LS-RTM on the C3 grid, Born "
                         "modelling + adjoint Jacobian with in-HBM checkpointing (secondary lines)")
    a = ap.parse_args()
    if a.workload == "c5" and a.shape == list(SHAPE):
        a.shape = [301, 301, 201]
    return a


# ---- synthetic "overthrust-shaped" model --------------------------------------------
def velocity(shape):
    nx, ny, nz = shape
    x = (np.arange(nx, dtype=np.float32) * SPACING[0]).reshape(nx, 1, 1)
    y = (np.arange(ny, dtype=np.float32) * SPACING[1]).reshape(1, ny, 1)
    z = (np.arange(nz, dtype=np.float32) * SPACING[2]).reshape(1, 1, nz)
    zr = z / np.float32((SHAPE[2] - 1) * SPACING[2])
    v = 1.5 + 4.0 * zr + 0.4 * np.sin(2 * np.pi * x / 5000.) * np.sin(2 * np.pi * y / 7000.) * zr
    v = np.clip(v, 1.5, 5.5).astype(np.float32)
    v[:, :, :21] = 1.5          # water layer
    v[0, 0, -1] = 5.5           # pins vmax so that dt = 1.921 ms for every sub-shape
    return v


def smooth_model(v):
    from scipy.ndimage import gaussian_filter1d
    v0 = gaussian_filter1d(v, sigma=10, axis=2, mode="nearest").astype(np.float32)
    v0[:, :, :21] = 1.5
    return v0


def geometry(shape, shot):
    ex = (shape[0] - 1) * SPACING[0]
    ey = (shape[1] - 1) * SPACING[1]
    gx = np.linspace(0.15 * ex, 0.85 * ex, 8)
    gy = np.linspace(0.15 * ey, 0.85 * ey, 8)
    i, j = shot % 8, (shot // 8) % 8
    src = np.array([[gx[i], gy[j], 50.]], dtype=np.float32)
    xr, yr = np.meshgrid(np.linspace(0., ex, 101), np.linspace(0., ey, 101), indexing="ij")
    rec = np.stack([xr.ravel(), yr.ravel(), np.full(xr.size, 50.)], axis=1).astype(np.float32)
    return src, rec


def ricker(nt, dt, f0):
    t = np.arange(nt, dtype=np.float32) * np.float32(dt)
    r = np.pi * f0 * (t - 1 / f0)
    return ((1 - 2 * r ** 2) * np.exp(-r ** 2)).astype(np.float32).reshape(nt, 1)


METRIC = {"c3": "fwi_gradient_shots_per_s", "c5": "fwi_gradient_shots_per_s",
          "c3fwd": "forward_modeling_shots_per_s", "c4": "lsrtm_born_plus_adjoint_shots_per_s"}


def workload_name(shape, nt, t_sub, region, workload="c3"):
    head = "3D %s so=8" % ("TTI" if workload == "c5" else "acoustic")
    tail = "n=%dx%dx%d d=25m nbl=40, 10201 receivers, nt=%d" % (shape[0], shape[1], shape[2], nt)
    if workload == "c3fwd":
        return "%s forward modelling (one shot record), %s, 1 shot/GPU/step, no collective" % (head, tail)
    if workload == "c4":
        return ("%s LS-RTM: Born modelling J*dm + adjoint Jacobian J'*d with in-HBM checkpointing "
                "(%s-domain snapshots), %s, 1 shot/GPU/step" % (head, region, tail))
    return ("%s FWI gradient (fwd+snapshots, adj+IC), %s, t_sub=%d, %s-domain snapshots, "
            "1 shot/GPU/step" % (head, tail, t_sub, region))


def config_dict(args, nt, dtc):
    """The SAME dict for the GPU arm and the CPU (`--impl reference`) arm."""
    shape = tuple(args.shape)
    npad = [n + 2 * NBL for n in shape]
    npts = float(np.prod(npad))
    return {"workload": workload_name(shape, nt, args.t_sub, args.snapshot_region, args.workload),
            "padded_grid": npad, "nt": nt, "dt_ms": dtc,
            "l2_policy": "inputs_larger_than_L2 (each field %.0f MB vs 126 MB L2)" % (4e-6 * npts),
            "parallelism": "shots partitioned over %d GPU(s), %s" % (
                args.gpus, "no collective" if args.workload == "c3fwd" else "one all-reduce per step")}


def tti_params(v):
    """eps = 0.2 (v - 1.5) / 4, delta = 0.1 (v - 1.5) / 4, theta = pi/6 (v - 1.5) / 4,
    phi = pi / 8."""
    return dict(epsilon=np.asfortranarray((0.2 * (v - 1.5) / 4).astype(np.float32)),
                delta=np.asfortranarray((0.1 * (v - 1.5) / 4).astype(np.float32)),
                theta=np.asfortranarray((np.pi / 6 * (v - 1.5) / 4).astype(np.float32)),
                phi=np.float32(np.pi / 8))


def model_dt(args):
    """dt_comp of the workload.  C5: the TTI model's critical dt, models.py:459-482 literally --
    sqrt(1 + 2 max(eps~)) with eps~ = 1 + 2 eps already applied (max eps = 0.2 here)."""
    if args.workload != "c5":
        return DT
    return float(np.float32("%.3e" % (DT / np.sqrt(1 + 2 * (1 + 2 * 0.2)))))


# ---- clocks -------------------------------------------------------------------------------------
class ClockSampler(object):
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(device), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "200"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [s.strip() for s in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                  "sw_power_cap"), c[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out["sm_mhz"] = float(np.median(sm))
            out["sm_max_mhz"] = float(max(mx))
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return out


# ---- CPU arm: the oracle on the host cores -------------------------------------------------------
def cpu_sample(shape, nt, t_sub, seconds, steps=1, warmup=0, tti=False, dt=DT, workload="c3"):
    """Times a bounded sample of the workload with the CPU oracle (all host threads) on the full
    grid and extrapolates linearly to one shot:
      c3 / c5  ns forward steps + ns adjoint+IC steps
      c3fwd    ns forward steps
      c4       Born sweep (two fields = 2 forward steps per time step) + J' with checkpointing
               (forward, recomputed forward, adjoint+IC): 4 forward steps + 1 adjoint+IC step per
               time step, from the forward-only and the forward+adjoint samples"""
    from oracle import oracle as O
    O.build()
    ncores = O.set_threads(os.cpu_count() or 1)   # torchrun sets OMP_NUM_THREADS=1: override
    v0 = smooth_model(velocity(shape))
    extra = {k_: np.ascontiguousarray(v_) for k_, v_ in tti_params(v0).items()} if tti else {}
    if tti:
        extra["phi"] = np.full(shape, extra["phi"], dtype=np.float32)
    model = O.Model((0., 0., 0.), SPACING, shape, space_order=SO, nbl=NBL,
                    m=(1 / v0 ** 2).astype(np.float32), dt=dt, precision="f32", **extra)
    grad = workload != "c3fwd"
    k = 1 if workload == "c4" else t_sub
    t1 = O.time_steps(model, 2, grad, k)          # calibration (also first-touch of memory)
    per_pair = max(t1 / 2, 1e-6) * (2.5 if workload == "c4" else 1.0)
    ns = int(max(4, min(nt - 2, seconds / per_pair)))
    ns -= ns % k
    ns = max(ns, k)
    times, fwd_times = [], []
    for _ in range(warmup + steps):
        times.append(O.time_steps(model, ns, grad, k))
        if workload == "c4":
            fwd_times.append(O.time_steps(model, ns, False, 1))
    times, fwd_times = times[warmup:], fwd_times[warmup:]
    per_step = float(np.mean(times)) / ns
    npts = float(np.prod(model.shape_pml))
    if workload == "c4":
        per_step += 3.0 * float(np.mean(fwd_times)) / ns
        what = ("%d of %d time steps of [forward + adjoint+IC] and of [forward] on the full "
                "%dx%dx%d padded grid; one LS-RTM shot = Born sweep (2 forward steps per time step) + "
                "J' with checkpointing (forward, recomputed forward, adjoint+IC), i.e. 3 x forward + "
                "(forward + adjoint+IC) per time step" % (ns, nt - 2, model.shape_pml[0],
                                                           model.shape_pml[1], model.shape_pml[2]))
        sweeps = 5
    elif workload == "c3fwd":
        what = "%d of %d forward time steps on the full %dx%dx%d padded grid" % (
            ns, nt - 2, model.shape_pml[0], model.shape_pml[1], model.shape_pml[2])
        sweeps = 1
    else:
        what = ("%d of %d forward + %d of %d adjoint+IC time steps on the full %dx%dx%d padded grid "
                "(t_sub=%d)" % (ns, nt - 2, ns, nt - 2, model.shape_pml[0], model.shape_pml[1],
                                model.shape_pml[2], t_sub))
        sweeps = 2
    sec_per_shot = per_step * (nt - 2)
    return {"value": 1.0 / sec_per_shot, "unit": "shots/s", "cores": ncores,
            "kind": "port",
            "sample": what + ", extrapolated linearly to one shot; OpenMP, all %d host threads; "
                      "oracle = C port of the reference's Devito schedule (Devito itself is not "
                      "installable offline)" % ncores,
            "gpts_per_s": sweeps * npts / per_step / 1e9,
            "ms_per_step": 1e3 * sec_per_shot}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    shape = tuple(args.shape)
    tti, dtc = args.workload == "c5", model_dt(args)
    nt = int(np.floor(args.tn / dtc)) + 1
    r = cpu_sample(shape, nt, args.t_sub, args.cpu_seconds / 2, steps=args.steps,
                   warmup=args.warmup, tti=tti, dt=dtc, workload=args.workload)
    line = {"impl": "reference", "impl_kind": "cpu_port",
            "metric": METRIC[args.workload], "value": r["value"],
            "unit": "shots/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args, nt, dtc),
            "note": "CPU arm: the oracle (a C port of the reference's Devito schedule, NOT Devito "
                    "itself), one process, all host threads, no GPU; bounded sample, extrapolated",
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "gpts_per_s": r["gpts_per_s"],
            "e2e": {"value": r["value"], "unit": "shots/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---- GPU arm ---------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    import judi_jl_b200 as J

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        # torch.distributed: rendezvous, barriers and the max-over-ranks of the timings only; the
        # data-plane all-reduce is the library's own (judi_fg_reduce, NCCL bound inside the C-ABI)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    ctx = J.get_context(local)
    J.objectives.init_comm(ctx)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)

    wl = args.workload
    shape = tuple(args.shape)
    tti, dtc = wl == "c5", model_dt(args)
    nt = int(np.floor(args.tn / dtc)) + 1
    v = velocity(shape)
    v0 = smooth_model(v)
    m_true = np.asfortranarray((1 / v ** 2).astype(np.float32))
    m0 = np.asfortranarray((1 / v0 ** 2).astype(np.float32))
    dm = np.asfortranarray(m_true - m0) if wl == "c4" else None
    src, rec = geometry(shape, rank)
    q = ricker(nt, dtc, F0)
    nrec = rec.shape[0]
    kw = dict(space_order=SO, nbl=NBL, dt=dtc)
    if tti:
        kw.update(tti_params(v0))     # the same anisotropy for the true and the starting model

    # ---- untimed setup: observed data of this rank's shot from the true model ---------------
    dobs = None
    if wl in ("c3", "c5"):
        mt = J.Model((0., 0., 0.), SPACING, shape, m=m_true, **kw)
        dobs, _ = J.forward_rec(mt, src, q, rec)
        del mt

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def to_dev(a):      # x fastest == reversed shape, contiguous
        return torch.from_numpy(np.ascontiguousarray(a.transpose(*range(a.ndim - 1, -1, -1)))).to(dev)

    with torch.cuda.stream(stream):
        m0_d, q_d = to_dev(m0), to_dev(q)
        kw_d = {k_: (to_dev(v_) if isinstance(v_, np.ndarray) and v_.ndim == 3 else v_)
                for k_, v_ in kw.items()}
        if dm is not None:
            kw_d["dm"] = to_dev(dm)
        model = J.Model((0., 0., 0.), SPACING, shape, m=m0_d, **kw_d)
        assert abs(float(model.critical_dt) - dtc) < 1e-6 * dtc, (model.critical_dt, dtc)
        npad = model.shape_pml
        ngrid, nphys_i = int(np.prod(npad)), int(np.prod(shape))
        red = torch.zeros(ngrid + 1, dtype=torch.float32, device=dev)    # padded accumulator + objective
        out = torch.zeros(nphys_i + 1, dtype=torch.float32, device=dev)  # reduced physical gradient + objective
        grad_d, f_d = red[:ngrid], red[ngrid:]
        rec_d = torch.zeros((nrec, nt), dtype=torch.float32, device=dev)
        if dobs is not None:
            rec_d.copy_(to_dev(dobs))
        ckpt = wl == "c4"
        opts = dict(return_obj=True, t_sub=1 if ckpt else args.t_sub, checkpointing=ckpt,
                    snapshot_region=args.snapshot_region, grad_out=grad_d, f_out=f_d)

        def step_device():
            if wl == "c3fwd":
                J.forward_rec(model, src, q_d, rec, rec_out=rec_d)
                return
            red.zero_()
            if wl == "c4":
                J.born_rec(model, src, q_d, rec, rec_out=rec_d)                  # d_lin = J dm
                J.J_adjoint(model, src, q_d, rec, rec_d, is_residual=True, accumulate=True, **opts)
            else:
                J.J_adjoint(model, src, q_d, rec, rec_d, accumulate=True, **opts)
            # crop on the device + ONE ncclAllReduce of prod(n)+1 floats inside the library
            J.objectives.fg_reduce(model, grad_d, f_d, out=out)

        # CUDA events around every 7th stencil launch (7 is coprime with t_sub, so the sample holds
        # the plain / save / imaging-condition flavours in their true proportions); bracketing
        # every launch costs ~4 us of stream time per event and would slow the timed region
        ctx.set_option("timing", 7)
        for _ in range(args.warmup):
            step_device()
        barrier()
        ctx.reset_stats()
        launches0 = ctx.launch_count
        clocks = ClockSampler(local)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            step_device()
        e1.record(stream)
        barrier()
        clk = clocks.stop()
        ms = e0.elapsed_time(e1)
        stats = ctx.stencil_stats()
        launches = ctx.launch_count - launches0
        if wl == "c3fwd":
            fval, gnorm = None, float(torch.linalg.vector_norm(rec_d).item())
        else:
            fval = float(out[nphys_i].item())
            gnorm = float(torch.linalg.vector_norm(out[:nphys_i]).item())
        ctx.set_option("timing", 0)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())

        # ---- e2e: host buffers in, host result out, every step ----------------------------------
        e2e = None
        if not args.no_e2e:
            def pinned(a):
                tt = torch.empty(a.shape[::-1], dtype=torch.float32).pin_memory()
                h = tt.numpy().T           # F-ordered view of pinned memory
                h[...] = a
                return tt, h
            keep_m, m0_h = pinned(m0)
            keep_q, q_h = pinned(q)
            keep_d, d_h = pinned(dobs if dobs is not None else np.zeros((nt, nrec), np.float32))
            keep_dm, dm_h = pinned(dm) if dm is not None else (None, None)
            # every model array lives in page-locked host memory (C5: the TTI parameter fields too)
            keep_kw, kw_h = [], dict(kw)
            for k_, v_ in kw.items():
                if isinstance(v_, np.ndarray) and v_.ndim == 3:
                    t_, kw_h[k_] = pinned(v_)
                    keep_kw.append(t_)
            # the public call: the objective / modelling over `world` shots, shot i -> rank i.  Each
            # rank builds its model from host arrays, propagates its shot from host wavelet / data,
            # sums in HBM, one all-reduce inside the library, and every rank copies the cropped
            # gradient (c3fwd: its shot record) back.
            shots = [(src, q_h, rec, d_h) if i == rank else None for i in range(world)]

            def step_host():
                mh = J.Model((0., 0., 0.), SPACING, shape, m=m0_h, **kw_h)
                if wl == "c3fwd":
                    return None, J.objectives.forward_modeling(mh, shots)[rank]
                if wl == "c4":
                    mh.set_dm(dm_h)
                    dl, _ = J.born_rec(mh, src, q_h, rec)
                    sh = [(src, q_h, rec, dl) if i == rank else None for i in range(world)]
                    return J.objectives.multi_src_fg(mh, sh, is_residual=True, checkpointing=True,
                                                     snapshot_region=args.snapshot_region)
                return J.objectives.fwi_objective(mh, shots, t_sub=args.t_sub,
                                                  snapshot_region=args.snapshot_region)
            for _ in range(args.warmup):      # the same W untimed steps as the device-resident region
                step_host()
            barrier()
            clocks_e2e = ClockSampler(local)
            t0 = time.perf_counter()
            for _ in range(args.steps):
                fh, gh = step_host()
            barrier()
            wall = time.perf_counter() - t0
            clk_e2e = clocks_e2e.stop()
            t = torch.tensor([wall], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            wall = float(t.item())
            h2d = world * (m0.nbytes + q.nbytes + (dobs.nbytes if dobs is not None else 0) +
                           (dm.nbytes if dm is not None else 0) +
                           sum(v_.nbytes for v_ in kw.values() if isinstance(v_, np.ndarray)))
            if wl == "c4":
                h2d += world * 4 * nt * nrec          # the Born data go back in as J' input
            d2h = world * (4 * nt * nrec if wl in ("c3fwd", "c4") else 0) + \
                (world * (4 * nphys_i + 4) if wl != "c3fwd" else 0)
            e2e = {"value": world * args.steps / wall, "unit": "shots/s",
                   "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                   "ms_per_step": 1e3 * wall / args.steps,
                   "call": {"c3fwd": "forward_modeling(Model(host arrays), host wavelet) -> host shot record",
                            "c4": "born_rec(Model(host arrays), dm) -> host data; multi_src_fg(is_residual, "
                                  "checkpointing) -> host gradient on the physical grid"}.get(
                       wl, "fwi_objective(Model(host arrays), host wavelet/data) -> host gradient on the "
                           "physical grid") + "; bytes summed over ranks",
                   "result_l2": float(np.linalg.norm(gh.ravel(order="K").astype(np.float64))),
                   "clocks": clk_e2e}
            if fval is not None:
                e2e["f_matches_device_path"] = bool(abs(float(fh) - fval) <= 1e-3 * abs(fval))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the stencil kernel over the timed region --------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    npts = float(ngrid)
    nphys = float(nphys_i) if args.snapshot_region == "physical" else npts
    k = args.t_sub
    nsteps_t = nt - 2
    nsave = len([t_ for t_ in range(1, nt - 1) if t_ % k == 0])
    # algorithmic bytes per grid point: step, + snapshot save, + imaging condition
    bpp, bsave, bic = (48.0, 8.0, 16.0) if tti else (16.0, 4.0, 12.0)
    # per launch flavour: plain step, + snapshot save, + imaging condition,
    # fused two-field Born step; the timed sample (every 7th launch) holds the flavours in the
    # proportions of the schedule that actually ran, so the average is schedule-independent
    kind_bytes = {"plain": bpp * npts, "save": bpp * npts + bsave * nphys,
                  "ic": bpp * npts + bic * nphys, "born": 2 * bpp * npts}
    tot_b = sum(kind_bytes[k_] * v_["launches"] for k_, v_ in stats["kinds"].items())
    tot_ms = sum(v_["ms"] for v_ in stats["kinds"].values())
    n_timed = sum(v_["launches"] for v_ in stats["kinds"].values())
    alg_bytes_per_launch = tot_b / max(n_timed, 1)
    ms_launch = tot_ms / max(n_timed, 1)
    achieved = alg_bytes_per_launch / (ms_launch * 1e-3) / 1e9
    nlaunch = {"c3fwd": nsteps_t, "c4": None}.get(wl, 2 * nsteps_t)
    traffic, traffic_src = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        traffic = tj.get("tti3d_tma_bytes_per_launch" if tti else "acoustic3d_tma_bytes_per_launch")
        traffic_src = tj.get("source")
    except Exception:
        pass
    kname = "tti3d_tma" if tti else "acoustic3d_tma"
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic,
                "traffic_source": "ncu --set full capture committed under profiles/ (%s); not "
                                  "re-measured in this run" % traffic_src,
                "kernel": kname + " (all stencil launches of the step averaged)",
                "algorithmic_bytes_per_launch": alg_bytes_per_launch,
                "avg_launch_ms": ms_launch, "launches_timed": stats["launches"],
                "launches_in_region": (nlaunch * args.steps) if nlaunch else 7 * int(stats["launches"]),
                "avg_launch_ms_by_flavour": {k_: v_["ms"] / v_["launches"]
                                             for k_, v_ in stats["kinds"].items()},
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else
                               "fallback 6650 GB/s (of fallback)"}

    shots = world * args.steps
    # grid sweeps per time step (c4: the Born step updates two fields; the J' sweeps are forward + adjoint
    # when the whole history stays in HBM, + one recomputation sweep otherwise)
    sweeps = {"c3fwd": 1, "c4": 5 if "plain" in stats["kinds"] else 4}.get(wl, 2)
    line = {"metric": METRIC[wl], "value": shots / (ms * 1e-3), "unit": "shots/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args, nt, dtc),
            "gpts_per_s": shots * sweeps * nsteps_t * npts / (ms * 1e-3) / 1e9,
            "clocks": clk, "gpu_launches": int(launches), "roofline": roofline,
            "collective": "none" if wl == "c3fwd" else
                          "ncclAllReduce of %d floats inside libjudi_b200 (judi_fg_reduce)" % (nphys_i + 1),
            "objective": fval, "result_l2": gnorm}
    if e2e is not None:
        line["e2e"] = e2e
    if world == 1 and not args.no_cpu:
        line["cpu_baseline"] = {kk: vv for kk, vv in
                                cpu_sample(shape, nt, args.t_sub, args.cpu_seconds, tti=tti,
                                           dt=dtc, workload=wl).items()
                                if kk in ("value", "unit", "cores", "kind", "sample")}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
