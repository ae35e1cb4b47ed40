"""2-GPU probe (measurement aid): cost of the stop-test exchange after a batched solve, fused peer-memory gather vs NCCL all-gather.
torchrun --nproc-per-node 2 profiles/microbench/peer_gather_probe.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import torch.distributed as dist

from control_box_rst_b200 import problems, solver

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
ocp, kw, _ = problems.config(1)
B = 4096
lm = solver.BatchedLevenbergMarquardt(ocp, B, device=lr)
lm.setIterations(10)
stream = torch.cuda.Stream(device=lr)
torch.cuda.set_stream(stream)
lm.set_stream(stream.cuda_stream)
x0, xref = problems.instance_data(ocp, B, seed=1, offset=rank * B)
lm.set_problem_data(x0, xref)
handles = [None] * world
dist.all_gather_object(handles, lm.peer_export(world, rank))
lm.peer_attach(handles)
dist.barrier()


class CA:
    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": "<f8", "data": (int(ptr), False), "version": 2}


chi2_local = torch.as_tensor(CA(lm.device_pointers()["chi2"], (B,)), device=f"cuda:{lr}")
chi2_all = torch.empty(B * world, dtype=torch.float64, device=f"cuda:{lr}")


def run(mode, steps, sync_each):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3 * steps)]
    tot = 0.0
    a = torch.cuda.Event(enable_timing=True)
    b = torch.cuda.Event(enable_timing=True)
    dist.barrier()
    torch.cuda.synchronize()
    a.record(stream)
    gaps = []
    for s in range(steps):
        lm.initialize_trajectories()
        lm.solve(new_run=True, fetch=False)
        ev[3 * s].record(stream)
        if mode == "p2p":
            lm.peer_wait()
        elif mode == "nccl":
            dist.all_gather_into_tensor(chi2_all, chi2_local)
        ev[3 * s + 1].record(stream)
        if sync_each:
            ev[3 * s + 1].synchronize()
    b.record(stream)
    torch.cuda.synchronize()
    gaps = [ev[3 * s].elapsed_time(ev[3 * s + 1]) for s in range(steps)]
    gaps.sort()
    return a.elapsed_time(b) / steps, gaps[len(gaps) // 2], gaps[-1]


for mode in ("none", "p2p", "nccl", "p2p", "nccl"):
    for sync_each in (False, True):
        run(mode, 5, sync_each)
        per, med, mx = run(mode, 100, sync_each)
        if rank == 0:
            print(f"mode={mode:5s} host-sync-each-step={sync_each!s:5s} ms/step={per:.4f} exchange median={med*1e3:.1f} us max={mx*1e3:.1f} us", flush=True)
print("rank", rank, "timed out:", lm.peer_timed_out())
dist.barrier()
lm.peer_detach()
lm.clear()
dist.destroy_process_group()
