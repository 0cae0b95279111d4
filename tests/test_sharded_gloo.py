"""CPU, world_size 2 over gloo: the host logic of the row-sharded POD / projection
(pod_uqnn_b200/sharded.py) gives the single-process answer."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, mode, out):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pod_uqnn_b200 import sharded, workloads as wl
    from np_backend import NumpyBackend
    from oracle import pod_oracle as po
    w = wl.ackley(24, 20, 48, seed=5)
    _, U, _, _ = po.create_snapshots(w["x_mesh"], 1, 0, po.u_ackley, w["mu"])
    lo, hi = sharded.row_range(U.shape[0], rank, world)
    be = NumpyBackend()
    V_loc, sig, n_L = sharded.pod_sharded(be, U[lo:hi], 1e-8, 0, mode)
    if mode != "tsqr":      # the refinement decision is taken from the largest slab on every rank (no collective per call)
        assert be.refine_rows == max(b - a for a, b in (sharded.row_range(U.shape[0], p, world) for p in range(world)))
    v = sharded.project_sharded(be, V_loc, U[lo:hi]).numpy()
    gathered = [None] * world
    dist.all_gather_object(gathered, (lo, hi, V_loc))
    if rank == 0:
        V = np.zeros((U.shape[0], n_L))
        for a, b, Vp in gathered:
            V[a:b] = Vp
        np.savez(out, V=V, v=v, sig=sig, n_L=n_L)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["gram", "gram2", "tsqr"])
def test_sharded_equals_unsharded(tmp_path, mode):
    out = str(tmp_path / "res.npz")
    mp.spawn(_worker, args=(2, _free_port(), mode, out), nprocs=2, join=True)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from pod_uqnn_b200 import sharded, workloads as wl
    from np_backend import NumpyBackend
    from oracle import pod_oracle as po, compare as cmp
    w = wl.ackley(24, 20, 48, seed=5)
    _, U, _, _ = po.create_snapshots(w["x_mesh"], 1, 0, po.u_ackley, w["mu"])
    be = NumpyBackend()
    V1, sig1, n1 = sharded.pod_sharded(be, U, 1e-8, 0, mode)      # no process group: single shard
    v1 = sharded.project_sharded(be, V1, U).numpy()
    r = np.load(out)
    assert int(r["n_L"]) == n1 == po.perform_pod(U, 1e-8).shape[1]
    assert cmp.sigma_error(r["sig"], sig1) < 1e-12
    assert cmp.largest_principal_angle(r["V"], V1) < 1e-8
    Va, s = cmp.align_signs(r["V"], V1)
    assert cmp.rel_err(r["v"] * s, v1) < 1e-6


def test_row_range_partitions():
    from pod_uqnn_b200.sharded import row_range
    for n, P in ((10, 3), (160000, 8), (7, 8), (1, 1)):
        spans = [row_range(n, p, P) for p in range(P)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1


def _train_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pod_uqnn_b200 import sharded
    res = sharded.train_ensemble_sharded(_numpy_trainer, _train_members(), None)
    if rank == 1:                       # any rank holds the full ensemble afterwards
        members, losses = res
        np.savez(out, losses=losses, **{f"w{i}_{k}": a for i, w in enumerate(members) for k, a in enumerate(w)})
    dist.barrier()
    dist.destroy_process_group()


_TRAIN_LAYERS = [2, 16, 16, 3]


def _train_members():
    from oracle import ensemble as ens
    return [ens.init_weights(_TRAIN_LAYERS, seed=40 + i, bias_scale=0.1) for i in range(3)]    # 3 members on 2 ranks


def _numpy_trainer(members):
    """Test double of EnsembleTrainer: the oracle's training loop, member by member."""
    from oracle import train as otr
    rng = np.random.default_rng(9)
    X = rng.uniform(-1, 1, (40, 2)); v = rng.standard_normal((40, 3))
    outs, losses = [], []
    for w in members:
        w1, l1, _ = otr.train(X, v, [a.copy() for a in w], 3, 6, 0.01, 1e-4, 0.01, 0.5)
        outs.append(w1); losses.append(l1)
    return outs, np.stack(losses, 1)


def test_member_sharded_training_equals_single_process(tmp_path):
    """Members split 2 + 1 over two ranks, no exchange while training, all-gather of weights and losses at the end."""
    out = str(tmp_path / "train.npz")
    mp.spawn(_train_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    ref_members, ref_losses = _numpy_trainer(_train_members())
    r = np.load(out)
    np.testing.assert_array_equal(r["losses"], ref_losses)
    for i, w in enumerate(ref_members):
        for k, a in enumerate(w):
            np.testing.assert_array_equal(r[f"w{i}_{k}"], a)
