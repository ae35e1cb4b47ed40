"""Multi-rank host logic on CPU (gloo, world_size 2): contiguous instance sharding with rank-independent instance data and the
single all-gather of per-instance chi2.  The per-instance solves are done by the oracle here (no GPU in this test)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from control_box_rst_b200 import _abi as abi
from control_box_rst_b200 import distributed, problems


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, total, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import bindings

    oracle = bindings.Oracle()
    ocp = problems.van_der_pol(12)
    lo, hi = distributed.shard_bounds(total, world, rank)
    x0, xref = problems.instance_data(ocp, hi - lo, seed=5, offset=lo)
    opts = abi.LmOptions.defaults(iterations=4)
    _, chi2, _, _ = oracle.solve_batch(ocp, opts, x0, xref, threads=1)
    gathered = distributed.gather_residuals(torch.from_numpy(chi2))
    np.save(os.path.join(out_dir, f"gathered_{rank}.npy"), gathered.numpy())
    dist.destroy_process_group()


def test_sharded_solve_equals_single_process(tmp_path, oracle):
    total, world = 12, 2
    mp.spawn(_worker, args=(world, _free_port(), total, str(tmp_path)), nprocs=world, join=True)
    ocp = problems.van_der_pol(12)
    x0, xref = problems.instance_data(ocp, total, seed=5)
    _, chi2, _, _ = oracle.solve_batch(ocp, abi.LmOptions.defaults(iterations=4), x0, xref, threads=1)
    for r in range(world):
        got = np.load(tmp_path / f"gathered_{r}.npy")
        assert np.array_equal(got, chi2)  # same instances whatever the sharding, gathered in global order on every rank


def test_shard_bounds_cover_everything():
    for total in (1, 7, 4096, 65537):
        for world in (1, 2, 3, 8):
            spans = [distributed.shard_bounds(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_instance_data_is_shard_invariant():
    ocp = problems.cart_pole_shooting(10)
    full, _ = problems.instance_data(ocp, 10, seed=3)
    a, _ = problems.instance_data(ocp, 4, seed=3, offset=0)
    b, _ = problems.instance_data(ocp, 6, seed=3, offset=4)
    assert np.array_equal(np.concatenate([a, b]), full)


def test_converged_fraction():
    prev = torch.tensor([1.0, 2.0, 3.0, 4.0], dtype=torch.float64)
    now = torch.tensor([1.0, 2.0 - 1e-9, 2.0, 4.0], dtype=torch.float64)
    assert distributed.converged_fraction(now, prev, rel_tol=1e-6) == 0.75
