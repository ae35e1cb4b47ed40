"""Two-rank test of the multi-GPU data path: the stop-test gather of the per-instance chi2 fused into the LM
kernel over NVLink peer memory (b200sqp_peer_*, control_box_rst_b200.distributed.StopTestExchange).  One process per GPU, NCCL for the
comparison collective; skipped on a box with fewer than two GPUs (run with `gpurun --gpus 2`).

Checked on every rank:
  * what the fused gather delivers equals an NCCL all-gather of the same solve bit for bit, over four consecutive solves with
    different start states (the gather buffers are double-buffered by solve parity);
  * it equals the chi2 of ONE process solving the whole job (sharding does not change any instance);
  * a rank that runs ahead never overwrites what a slow rank still has to read (rank 1 stalls its stream between the wait and the
    read of solve s while rank 0 already stores solve s+1);
  * a second b200sqp_peer_attach is refused;
  * the bounded wait: a rank whose peer never solves gets the timeout flag instead of a hang;
  * detach after a barrier, handles destroyed cleanly.
"""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist

    from control_box_rst_b200 import distributed, problems, solver

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    result = {}
    ocp = problems.van_der_pol(20)
    B = 96 + 5  # not a multiple of the 32-instance tile
    lm = solver.BatchedLevenbergMarquardt(ocp, B, device=rank)
    lm.setIterations(6)
    stream = torch.cuda.Stream(device=rank)
    torch.cuda.set_stream(stream)
    lm.set_stream(stream.cuda_stream)
    exch = distributed.StopTestExchange(lm, mode="p2p", device=dev)
    assert exch.mode == "p2p"
    # a second attach must be refused (it would leak the mappings and desynchronise the arrival counters)
    try:
        lm.peer_attach([b"\0" * 64] * world)
        result["second_attach_refused"] = False
    except solver.B200SqpError as e:
        result["second_attach_refused"] = e.code == -1
    chi2_local = distributed.device_view(lm.device_pointers()["chi2"], B, dev)
    nccl_all = torch.empty(B * world, dtype=torch.float64, device=dev)
    gathered, reference = [], []
    n_solves = 4
    for s in range(n_solves):
        x0, xref = problems.instance_data(ocp, B, seed=100 + s, offset=rank * B)
        lm.set_problem_data(x0, xref)
        lm.initialize_trajectories()
        lm.solve(new_run=True, fetch=False)
        exch.wait()
        if rank == 1:
            torch.cuda._sleep(int(2e8))  # ~0.1 s: rank 0 runs ahead into its next solve and stores into this rank's other buffer
        gathered.append(exch.gathered().clone())
        dist.all_gather_into_tensor(nccl_all, chi2_local)
        reference.append(nccl_all.clone())
    torch.cuda.synchronize()
    result["fused_equals_nccl"] = [bool(torch.equal(g, r)) for g, r in zip(gathered, reference)]
    result["solves_differ"] = bool(not torch.equal(gathered[0], gathered[1]))
    result["timed_out_after_regular_solves"] = bool(exch.timed_out())
    np.save(os.path.join(out_dir, f"gathered_{rank}.npy"), torch.stack(gathered).cpu().numpy())
    dist.barrier()
    # ---- the bounded wait: rank 1 sits this solve out, rank 0's wait must give up after 2 s and raise the flag
    if rank == 0:
        lm.solve(new_run=True, fetch=False)
        exch.wait()
        torch.cuda.synchronize()
        result["timeout_flag"] = bool(exch.timed_out())
    dist.barrier()
    exch.close()
    lm.clear()
    np.save(os.path.join(out_dir, f"result_{rank}.npy"), np.array([result], dtype=object), allow_pickle=True)
    dist.destroy_process_group()


def test_fused_peer_gather_two_ranks(tmp_path):
    import torch
    import torch.multiprocessing as mp

    from control_box_rst_b200 import problems, solver

    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs on one NVSwitch box (gpurun --gpus 2)")
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    res = [np.load(tmp_path / f"result_{r}.npy", allow_pickle=True)[0] for r in range(world)]
    got = [np.load(tmp_path / f"gathered_{r}.npy") for r in range(world)]
    for r in range(world):
        assert res[r]["second_attach_refused"], r
        assert all(res[r]["fused_equals_nccl"]), (r, res[r]["fused_equals_nccl"])
        assert res[r]["solves_differ"]
        assert not res[r]["timed_out_after_regular_solves"]
    assert res[0]["timeout_flag"], "a peer that never arrives must raise the timeout flag, not hang"
    assert np.array_equal(got[0], got[1])
    # one process solving the whole job gives the same chi2 for every instance
    ocp = problems.van_der_pol(20)
    B = 96 + 5
    for s in range(got[0].shape[0]):
        x0, xref = problems.instance_data(ocp, B * world, seed=100 + s)
        lm = solver.BatchedLevenbergMarquardt(ocp, B * world)
        lm.setIterations(6)
        lm.set_problem_data(x0, xref)
        lm.initialize_trajectories()
        _, chi2 = lm.solve(new_run=True)
        lm.clear()
        assert np.array_equal(chi2, got[0][s]), s
