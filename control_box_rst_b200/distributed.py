"""Host logic of the multi-GPU path: instances are independent, so a batch shards contiguously over the
ranks with no data-path collective; the only exchange is one all-gather of the per-instance chi2 (the stop-test residuals).
Works on any torch.distributed backend (nccl on the GPUs, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def shard_bounds(total, world_size, rank):
    """Contiguous shard [lo, hi) of `total` instances for `rank` (sizes differ by at most one)."""
    lo = total * rank // world_size
    hi = total * (rank + 1) // world_size
    return lo, hi


def gather_residuals(chi2_local, out=None):
    """All-gather of the per-instance chi2 of equally sized shards -> [world * B_local] on every rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return chi2_local
    world = dist.get_world_size()
    if out is None:
        out = torch.empty(world * chi2_local.numel(), dtype=chi2_local.dtype, device=chi2_local.device)
    dist.all_gather_into_tensor(out, chi2_local.contiguous())
    return out


def converged_fraction(chi2_all, chi2_prev_all, rel_tol=1e-6):
    """Batch-level stop test on the gathered residuals: fraction of instances whose chi2 moved by less than rel_tol."""
    moved = (chi2_prev_all - chi2_all).abs() > rel_tol * chi2_prev_all.abs().clamp_min(1e-300)
    return 1.0 - moved.double().mean().item()


class _CudaArray:
    """Zero-copy view of a raw device pointer for torch.as_tensor (CUDA array interface v2)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (int(ptr), False), "version": 2}


def device_view(ptr, count, device, typestr="<f8"):
    """torch tensor over `count` elements of raw device memory owned by the library (no copy)"""
    return torch.as_tensor(_CudaArray(ptr, (int(count),), typestr), device=device)


class StopTestExchange:
    """The one collective of the path: every rank ends a batched solve holding the chi2 of ALL instances of the job.

    mode "p2p" (default with more than one rank): the exchange is fused into the LM kernel -- its epilogue stores the per-instance chi2
    straight into every rank's gather buffer over NVLink peer memory (b200sqp_peer_*); torch.distributed only carries the 64-byte CUDA IPC
    handles once, here.  `wait()` enqueues the bounded, stream-ordered wait for all ranks' values of the last solve.
    mode "nccl": one all_gather_into_tensor of the handle's chi2 array after the solve (the library-collective comparison).
    Equal batch on all ranks; one process per GPU; all ranks on one NVSwitch box for "p2p".
    """

    def __init__(self, lm, mode="p2p", device=None):
        self.lm = lm
        self.world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
        self.rank = dist.get_rank() if self.world > 1 else 0
        self.mode = mode if self.world > 1 else "none"
        self.device = device if device is not None else torch.device("cuda", lm.device)
        B = lm.batch
        self._chi2_local = device_view(lm.device_pointers()["chi2"], B, self.device)
        self._chi2_all = None
        if self.mode == "p2p":
            handles = [None] * self.world
            dist.all_gather_object(handles, lm.peer_export(self.world, self.rank))
            lm.peer_attach(handles)
            dist.barrier()  # nobody solves before every rank has mapped every buffer
        elif self.mode == "nccl":
            self._chi2_all = torch.empty(B * self.world, dtype=torch.float64, device=self.device)

    def wait(self):
        """enqueue the exchange of the last solve on the current stream (call after lm.solve(fetch=False))"""
        if self.mode == "p2p":
            self.lm.peer_wait()
        elif self.mode == "nccl":
            dist.all_gather_into_tensor(self._chi2_all, self._chi2_local)

    def gathered(self):
        """chi2 of all instances of the job in global instance order, [world * batch] on this rank's GPU (valid after wait())"""
        if self.mode == "p2p":
            return device_view(self.lm.peer_gathered_ptr(), self.lm.batch * self.world, self.device)
        if self.mode == "nccl":
            return self._chi2_all
        return self._chi2_local

    def timed_out(self):
        return self.mode == "p2p" and self.lm.peer_timed_out()

    def close(self):
        if self.mode == "p2p":
            torch.cuda.synchronize()
            dist.barrier()  # nobody unmaps while a peer may still store into the mapping
            self.lm.peer_detach()
            self.mode = "none"
