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
