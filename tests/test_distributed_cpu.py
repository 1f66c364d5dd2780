"""The N>1 path on CPU: world_size-2 gloo run of the shot scheduler + reduction
(judi.jl_b200/objectives.py, mirroring propagation.jl:93-107 and distributed.jl:60-95 and
test/test_multi_exp.jl:83-101: the distributed objective equals the sum of the single-shot ones).
The per-shot compute is replaced by the CPU oracle, so only host logic is exercised here."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, model_args, geometry

NT = 120


def _shots(args, om, nshots):
    from oracle import oracle as O
    src0, rec = geometry(args)
    dt = float(om.critical_dt)
    q = O.ricker_wavelet((NT - 1) * dt, dt, 0.02)[:NT]
    shots = []
    for i in range(nshots):
        src = src0.copy()
        src[0, 0] += 40.0 * (i - nshots / 2)
        d, _ = O.forward_rec(om, src, q, rec)
        shots.append((src, q, rec, (0.9 * d).astype(np.float32)))
    return shots


class _FakeModel(object):
    """Stands in for judi_jl_b200.Model in a process without a GPU."""
    def __init__(self, om):
        self.om = om
        self.shape_pml = om.shape_pml
        self.padsizes = om.padsizes

    def set_dm(self, dm):
        raise AssertionError("not used")


def _oracle_J_adjoint(model, sc, q, rc, dobs, return_obj=False, born_fwd=False, nlind=False,
                      misfit=None, **kw):
    from oracle import oracle as O
    f, g, Iu, Iv = O.J_adjoint(model.om, sc, q, rc, dobs, return_obj=True, t_sub=kw.get("t_sub", 1))
    return np.float32(f), np.asfortranarray(g.astype(np.float32)), Iu, Iv


def _worker(rank, world, port, nshots, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), OMP_NUM_THREADS="2")
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import judi_jl_b200 as J
    from oracle import oracle as O
    J.interface.J_adjoint = _oracle_J_adjoint
    args = model_args("iso", 2, 8, with_dm=False, n=(61, 51), nbl=12)
    om = O.Model(precision="f32", **args)
    shots = _shots(args, om, nshots)
    mine = J.objectives.shot_partition(nshots, world, rank)
    f, g = J.objectives.fwi_objective(_FakeModel(om), shots)
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), f=f, g=g, mine=np.array(mine))
    dist.destroy_process_group()


def test_shot_partition_is_a_disjoint_cover():
    import judi_jl_b200 as J
    for nsrc in (1, 5, 16, 64):
        for ws in (1, 2, 3, 8):
            parts = [J.objectives.shot_partition(nsrc, ws, r) for r in range(ws)]
            flat = sorted(i for p in parts for i in p)
            assert flat == list(range(nsrc))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_two_rank_gloo_objective_equals_serial_sum(tmp_path):
    import torch.multiprocessing as mp
    import judi_jl_b200 as J
    from oracle import oracle as O
    nshots, world, port = 5, 2, 29500 + (os.getpid() % 2000)
    mp.start_processes(_worker, args=(world, port, nshots, str(tmp_path)), nprocs=world,
                       join=True, start_method="spawn")
    r0 = np.load(tmp_path / "rank0.npz")
    r1 = np.load(tmp_path / "rank1.npz")
    assert sorted(list(r0["mine"]) + list(r1["mine"])) == list(range(nshots))
    # every rank holds the same reduced result
    assert r0["f"] == r1["f"] and np.array_equal(r0["g"], r1["g"])
    # and it equals the serial sum over all shots
    args = model_args("iso", 2, 8, with_dm=False, n=(61, 51), nbl=12)
    om = O.Model(precision="f32", **args)
    shots = _shots(args, om, nshots)
    fs, gs = 0.0, 0.0
    for sc, q, rc, dobs in shots:
        f, g, _, _ = O.J_adjoint(om, sc, q, rc, dobs, return_obj=True)
        fs += float(f)
        gs = gs + g
    gs = J.objectives.remove_padding(gs, om.padsizes)
    assert r0["g"].shape == args["shape"]
    assert abs(float(r0["f"]) - fs) <= 1e-5 * abs(fs)
    assert np.linalg.norm(r0["g"] - gs) <= 1e-5 * np.linalg.norm(gs)
