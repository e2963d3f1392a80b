"""
Host-side plumbing of the wavelength-sharded mode (torch.distributed; NCCL on GPUs,
gloo in the CPU tests).  Every rank owns a contiguous wavelength slice of the
tables, flux state and per-wavelength constants; T, P and the mixing ratios are
replicated.  Per sweep the only exchange is a sum of the [B][L][4] wavelength
integrals; at the end the spectrum and delta_tau slices are gathered.
"""
import numpy as np

__all__ = ['shard_range', 'shard_ranges', 'allreduce_sums', 'gather_lambda']


def shard_range(n, rank, world):
    """Contiguous slice [lo, hi) of an axis of length n owned by ``rank`` of ``world``."""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_ranges(n, world):
    return [shard_range(n, r, world) for r in range(world)]


def allreduce_sums(sums, group=None):
    """In-place sum over ranks of the per-layer wavelength integrals (any device/backend)."""
    import torch.distributed as dist
    dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    return sums


def gather_lambda(local, n_global, group=None, to_numpy=True):
    """
    All-gather wavelength slices (last axis, possibly uneven) of ``local`` [..., n_local]
    into the full [..., n_global] array on every rank.  The reassembly happens on the device
    (one collective, one permute); ``to_numpy=False`` returns the device tensor.
    """
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    sizes = shard_ranges(n_global, world)
    nmax = max(hi - lo for lo, hi in sizes)
    lead = tuple(local.shape[:-1])
    buf = torch.zeros(lead + (nmax,), dtype=local.dtype, device=local.device)
    buf[..., :local.shape[-1]] = local
    out = torch.empty((world,) + lead + (nmax,), dtype=local.dtype, device=local.device)
    if local.device.type == 'cuda':
        dist.all_gather_into_tensor(out, buf, group=group)
    else:                                                   # gloo: list form
        parts = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(parts, buf, group=group)
        out = torch.stack(parts)
    nd = out.dim()
    full = out.permute(*range(1, nd - 1), 0, nd - 1).reshape(lead + (world * nmax,))
    if any(hi - lo != nmax for lo, hi in sizes):            # drop the padding of short shards
        keep = torch.cat([torch.arange(r * nmax, r * nmax + hi - lo, device=local.device)
                          for r, (lo, hi) in enumerate(sizes)])
        full = full.index_select(full.dim() - 1, keep)
    return full.cpu().numpy() if to_numpy else full
