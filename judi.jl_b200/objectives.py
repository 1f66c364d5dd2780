"""Mirror of the Julia drivers above the boundary: the shot scheduler and reduction
(src/TimeModeling/Modeling/propagation.jl:23-107, distributed.jl:6-95) and the objective
wrappers (misfit_fg.jl:143-198), with the reference's process-per-worker Julia `Distributed`
pool replaced by one process per GPU and a single all-reduce of (gradient, objective).

Shots are independent until the final sum, so they are partitioned round-robin over the ranks
(shot i -> rank i mod world_size); every rank accumulates its shots locally (in HBM when the
inputs are device resident) and ONE all-reduce replaces `reduce!`'s binary tree over Futures: on
GPUs it is issued inside the C-ABI (`judi_fg_reduce`: crop on the device, ncclAllReduce of the
physical-grid gradient + objective on the library's stream; torch.distributed only carries the
128-byte NCCL id at start-up), in the CPU tests it is a gloo all-reduce of host arrays.
"""
import numpy as np


def remove_padding(g, nb, true_adjoint=False):
    """src/TimeModeling/Utils/auxiliaryFunctions.jl:79-89 (crop; true_adjoint folds the pad)."""
    g = np.array(g, copy=True)
    if true_adjoint:
        for dim, (nbl, nbr) in enumerate(nb):
            N = g.shape[dim]

            def sel(i):
                s = [slice(None)] * g.ndim
                s[dim] = i
                return tuple(s)
            g[sel(nbl)] += g[sel(slice(0, nbl))].sum(axis=dim)
            g[sel(N - nbr - 1)] += g[sel(slice(N - nbr, N))].sum(axis=dim)
    crop = tuple(slice(nbl, n - nbr) for (nbl, nbr), n in zip(nb, g.shape))
    return g[crop]


def shot_partition(nsrc, world_size, rank):
    """Static round-robin: the shots rank `rank` propagates."""
    return list(range(rank, nsrc, world_size))


def _dist():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist
    except ImportError:
        pass
    return None


_staging = {}


def pinned_gradient_buffer(shape_pml):
    """Host buffer for one padded gradient + objective in page-locked memory: returns
    (F-ordered numpy view of the gradient part, torch tensor of prod(shape)+1 floats).  Passing
    the view as `grad_out` to J_adjoint and the tensor as `pinned` to reduce_fg makes the
    reduction copy-free on the host side."""
    import torch
    n = int(np.prod(shape_pml))
    t = torch.zeros(n + 1, dtype=torch.float32).pin_memory()
    g = t[:n].numpy().reshape(tuple(shape_pml)[::-1]).T
    return g, t


def reduce_fg(f, g, pinned=None):
    """Sum (objective, gradient) over all ranks: distributed.jl:60-95 as one all-reduce.

    `g` is a numpy array (host path) or a torch tensor (device path, reduced in place over
    NCCL); the objective scalar rides along as one extra element, as in SURVEY 8(e).  `pinned`
    (see pinned_gradient_buffer) must alias `g` in its first prod(shape) elements."""
    dist = _dist()
    if dist is None or dist.get_world_size() == 1:
        return f, g
    import torch
    if hasattr(g, "is_cuda"):
        buf = torch.cat([g.reshape(-1), torch.as_tensor([float(f)], dtype=g.dtype, device=g.device)])
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)
        g.copy_(buf[:-1].reshape(g.shape))
        return float(buf[-1].item()), g
    nccl = dist.get_backend() == "nccl"
    if pinned is not None:
        pinned[-1] = float(f)
        if nccl:
            key = ("dev", pinned.numel())
            if key not in _staging:
                _staging[key] = torch.empty(pinned.numel(), dtype=torch.float32, device="cuda")
            d = _staging[key]
            d.copy_(pinned, non_blocking=True)
            dist.all_reduce(d, op=dist.ReduceOp.SUM)
            pinned.copy_(d, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        else:
            dist.all_reduce(pinned, op=dist.ReduceOp.SUM)
        return float(pinned[-1]), g
    g = np.asarray(g, dtype=np.float32)
    flat = np.concatenate([g.ravel(order="F"), np.array([f], dtype=np.float32)])
    t = torch.from_numpy(flat)
    if nccl:
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    t = t.cpu().numpy()
    return float(t[-1]), t[:-1].reshape(g.shape, order="F")


def multi_src_fg(model, shots, dm=None, lin=False, nlind=False, misfit=None, reduce=True,
                 sum_padding=False, **options):
    """multi_src_fg! (propagation.jl:93-107) over this rank's share of `shots`.

    `shots`: sequence of (src_coords, wavelet, rec_coords, dobs) per source experiment.
    Returns (f, gradient on the physical grid), summed over all shots of all ranks."""
    from . import interface
    dist = _dist()
    ws = dist.get_world_size() if dist else 1
    rk = dist.get_rank() if dist else 0
    if dm is not None:
        model.set_dm(dm)
    if getattr(model, "h", None) is not None and "grad_out" not in options:
        return _multi_src_fg_hbm(model, shots, shot_partition(len(shots), ws, rk), lin, nlind,
                                 misfit, reduce, sum_padding, options)
    f, g = 0.0, None
    for i in shot_partition(len(shots), ws, rk):
        sc, q, rc, dobs = shots[i]
        fi, gi, _, _ = interface.J_adjoint(model, sc, q, rc, dobs, return_obj=True,
                                           born_fwd=lin, nlind=nlind, misfit=misfit, **options)
        f += float(fi)
        g = gi if g is None else g + gi
    if g is None:
        g = np.zeros(model.shape_pml, dtype=np.float32, order="F")
    if reduce:
        f, g = reduce_fg(f, g)
    return np.float32(f), remove_padding(g, model.padsizes, true_adjoint=sum_padding)


def _nccl_library():
    """Path of the NCCL library the process should bind (torch's bundled copy when torch.distributed
    already runs over NCCL, so that one libnccl serves both), or None for the loader's default."""
    try:
        import nvidia.nccl as _n
        import os as _os
        for d in list(getattr(_n, "__path__", [])):
            f = _os.path.join(d, "lib", "libnccl.so.2")
            if _os.path.exists(f):
                return f
    except ImportError:
        pass
    return None


def init_comm(ctx):
    """Create the library's own NCCL communicator for this rank's context (judi_comm_init):
    rank 0 draws the unique id, the 128 bytes travel over the host-side rendezvous
    (torch.distributed's store -- control plane only, as Julia's Distributed would do it), and
    from then on the data-plane reduction is `judi_fg_reduce`, inside the C-ABI.
    No-op when already initialised or when there is one rank."""
    import ctypes
    import os
    from . import _lib as L
    dist = _dist()
    ws = dist.get_world_size() if dist else 1
    if L.lib().judi_comm_size(ctx.h) == ws and (ws == 1 or getattr(ctx, "_comm_ready", False)):
        return ws
    rk = dist.get_rank() if dist else 0
    if ws == 1:
        L.check(L.lib().judi_comm_init(ctx.h, 1, 0, None))
        ctx._comm_ready = True
        return 1
    if "JUDI_NCCL_LIB" not in os.environ:
        f = _nccl_library()
        if f:
            os.environ["JUDI_NCCL_LIB"] = f
    ident = ctypes.create_string_buffer(128)
    if rk == 0:
        L.check(L.lib().judi_comm_unique_id(ident))
    box = [ident.raw]
    dist.broadcast_object_list(box, src=0)
    ident = ctypes.create_string_buffer(box[0], 128)
    L.check(L.lib().judi_comm_init(ctx.h, ws, rk, ident))
    ctx._comm_ready = True
    return ws


def fg_reduce(model, grad_padded_dev, f_dev, sum_padding=False, out=None):
    """judi_fg_reduce: crop / fold the padded device accumulator, append the objective, ONE
    ncclAllReduce of prod(n)+1 floats on the library's stream, result in page-locked host memory.
    `grad_padded_dev`, `f_dev`: torch CUDA tensors (or raw device pointers)."""
    import ctypes
    from . import _lib as L
    ptr = lambda t: t.data_ptr() if hasattr(t, "data_ptr") else int(t)
    if out is not None and hasattr(out, "data_ptr"):
        # device output: `out` holds prod(n)+1 floats, gradient then objective (stays in HBM)
        nphys = int(np.prod(model.shape))
        assert out.numel() >= nphys + 1
        L.check(L.lib().judi_fg_reduce(model.h, ptr(grad_padded_dev), ptr(f_dev), int(bool(sum_padding)),
                                       out.data_ptr(), out.data_ptr() + 4 * nphys, L.OUT_ON_DEVICE))
        return out[nphys], out[:nphys]
    g = out if out is not None else np.zeros(model.shape, dtype=np.float32, order="F")
    assert g.flags.f_contiguous and g.shape == tuple(model.shape)
    f = ctypes.c_float(0)
    L.check(L.lib().judi_fg_reduce(model.h, ptr(grad_padded_dev), ptr(f_dev), int(bool(sum_padding)),
                                   g.ctypes.data, ctypes.addressof(f), 0))
    return np.float32(f.value), g


_hbm_acc = {}
_MULTISHOT_OPTIONS = {"is_residual", "checkpointing", "n_checkpoints", "t_sub", "ic", "snapshot_region", "f0"}


def _multi_src_fg_hbm(model, shots, mine, lin, nlind, misfit, reduce, sum_padding, options):
    """The production path: every shot of this rank adds its gradient and objective
    to one accumulator in HBM (`JUDI_OUT_ON_DEVICE | JUDI_GRAD_ACCUMULATE`), the ranks are summed
    by ONE all-reduce of prod(n)+1 floats over NCCL, and the result crosses PCIe once, already
    cropped to the physical grid.  The returned gradient is a view of a reused page-locked
    staging buffer: it is valid until the next call (copy it to keep it)."""
    import torch
    from . import interface
    dist = _dist()
    n = int(np.prod(model.shape_pml))
    dev = torch.device("cuda", model.ctx.device)
    key = (model.ctx.device, n)
    if key not in _hbm_acc:
        _hbm_acc.clear()
        _hbm_acc[key] = (torch.zeros(n + 1, dtype=torch.float32, device=dev),
                         torch.zeros(n + 1, dtype=torch.float32).pin_memory())
    red, host = _hbm_acc[key]
    stream = torch.cuda.ExternalStream(model.ctx.stream, device=dev)
    nccl = reduce and dist is not None and dist.get_world_size() > 1 and dist.get_backend() == "nccl"
    if nccl:
        init_comm(model.ctx)
    with torch.cuda.stream(stream):
        red.zero_()
        if misfit is None and not (set(options) - _MULTISHOT_OPTIONS) and mine:
            # one library call for all shots of this rank (judi_fg_multishot)
            stream.synchronize()
            interface.fg_multishot(model, [shots[i] for i in mine], red[:n], red[n:], born_fwd=lin,
                                   nlind=nlind, accumulate=True, **options)
        else:
            for i in mine:
                sc, q, rc, dobs = shots[i]
                interface.J_adjoint(model, sc, q, rc, dobs, return_obj=True, born_fwd=lin, nlind=nlind,
                                    misfit=misfit, grad_out=red[:n], f_out=red[n:], accumulate=True,
                                    **options)
        if nccl or not (reduce and dist is not None and dist.get_world_size() > 1):
            # the collective lives behind the C-ABI (judi_fg_reduce): crop (or fold the pad) on the
            # device, all-reduce prod(n_physical)+1 floats over NCCL on the library's stream, and
            # only those cross PCIe -- no torch.distributed on the data plane
            nphys = int(np.prod(model.shape))
            g = host[:nphys].numpy().reshape(tuple(model.shape)[::-1]).T
            stream.synchronize()
            f, g = fg_reduce(model, red[:n], red[n:], sum_padding=sum_padding, out=g)
            return f, g
        host.copy_(red, non_blocking=True)      # non-NCCL process group: reduce on the host
        stream.synchronize()
        dist.all_reduce(host, op=dist.ReduceOp.SUM)
    g = host[:n].numpy().reshape(tuple(model.shape_pml)[::-1]).T
    return np.float32(host[n].item()), remove_padding(g, model.padsizes, true_adjoint=sum_padding)


def forward_modeling(model, shots, born=False, **options):
    """`F*q` / `J*dm` for a list of source experiments (judiModeling / judiJacobian applied through
    `run_and_reduce` with no reduction, propagation.jl:23-50): shot i is modelled by rank
    i mod world_size; returns {shot index: record (nt, nrec)} for this rank's shots -- shot
    records stay with the worker that made them, there is no collective."""
    from . import interface
    dist = _dist()
    ws = dist.get_world_size() if dist else 1
    rk = dist.get_rank() if dist else 0
    out = {}
    for i in shot_partition(len(shots), ws, rk):
        sc, q, rc = shots[i][:3]
        fn = interface.born_rec if born else interface.forward_rec
        out[i] = fn(model, sc, q, rc, **options)[0]
    return out


def fwi_objective(model, shots, **options):
    """fwi_objective (misfit_fg.jl:143-148): f = sum_i 1/2 dt ||F(m) q_i - d_i||^2 and its
    gradient w.r.t. the squared slowness."""
    return multi_src_fg(model, shots, dm=None, lin=False, **options)


def lsrtm_objective(model, shots, dm, nlind=False, **options):
    """lsrtm_objective (misfit_fg.jl:161-166): f = sum_i 1/2 dt ||J dm - (d_i - nlind*F q_i)||^2."""
    return multi_src_fg(model, shots, dm=dm, lin=True, nlind=nlind, **options)
