"""Multi-GPU glue for the hot path: one process per GPU, rows of a region sharded into bands.

Mirrors the only "collective" of the reference -- the parent process summing the per-region
statistics and N_r-weighted costs it pulled off the queue (base.py:384-396, 571-580) -- as
one all-reduce(sum) of K(1+d+d^2)+3 doubles, plus the max-reduction pygco's
down_weight_factor needs when ONE region spans several GPUs.  `torch.distributed` is the
plumbing (NCCL on GPUs; gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np


class DeviceView:
    """Expose a raw device pointer of libphmrf as a torch tensor (no copy)."""

    def __init__(self, ptr, nelem, typestr):
        self.__cuda_array_interface__ = {"shape": (int(nelem),), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 3}


def stats_tensor(region):
    import torch
    return torch.as_tensor(DeviceView(region.stats_device_ptr(), region.model.stats_len, "<f8"), device="cuda")


def absmax_tensor(region):
    """max|logp| of the band as int64 bits: for non-negative doubles integer order is numeric
    order, so a MAX all-reduce of the bit patterns is the max of the values."""
    import torch
    return torch.as_tensor(DeviceView(region.absmax_device_ptr(), 1, "<i8"), device="cuda")


def share_weight_max(region, dist):
    """Region-wide max|w| (one-off, at set-up)."""
    import torch
    t = torch.tensor([region.weight_max()], dtype=torch.float64, device="cuda" if dist.get_backend() == "nccl" else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    region.set_weight_max(float(t.item()))
    return float(t.item())


def combine_band_results(stats_flat, cost_sums, n_own, dist):
    """All-reduce one band's (statistics, cost sums, node count) and return the region-wide
    statistics dict plus the four scalars of _compute_cost_v1.  Host-tensor variant used by
    the CPU (gloo) tests and by callers that already copied the results to the host."""
    import torch
    from .engine import costs_from_sums
    buf = torch.from_numpy(np.concatenate([np.asarray(stats_flat, dtype=np.float64).ravel(),
                                           np.asarray(cost_sums, dtype=np.float64).ravel(), [float(n_own)]]))
    dist.all_reduce(buf)
    out = buf.numpy()
    n_total = out[-1]
    return out[:-4], costs_from_sums(out[-4:-1], n_total), int(round(n_total))


def global_dwf(absmax_local, wmax_local, vmax, dist):
    """pygco's down_weight_factor for a region whose bands live on different ranks."""
    import torch
    t = torch.tensor([absmax_local, wmax_local], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return max(float(t[0]), float(t[1]) * vmax) + 1e-10
