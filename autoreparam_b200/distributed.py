"""Multi-GPU plumbing: chains are independent, so an HMC job is sharded over the
ranks (one process per GPU, torch.distributed) with NO collective inside the
sampler.  Collectives appear only where the path has a real exchange step
(SURVEY.md 8e): gathering per-chain ESS for util.get_min_ess, summing accept
counts, and reducing the per-chain moments R-hat is built from.  Backend: NCCL
over NVLink on GPUs, gloo on CPU (tests)."""
from __future__ import annotations

import os

import numpy as np


def _dist():
    import torch.distributed as dist
    return dist


def rank_world():
    try:
        dist = _dist()
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except Exception:
        pass
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def local_device():
    import torch
    if torch.cuda.is_available():
        return torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    return torch.device("cpu")


def init_if_needed(backend=None):
    import torch
    dist = _dist()
    if int(os.environ.get("WORLD_SIZE", "1")) <= 1 or dist.is_initialized():
        return
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        dev = local_device()
        torch.cuda.set_device(dev)
        dist.init_process_group(backend, device_id=dev)
    else:
        dist.init_process_group(backend)


def barrier():
    dist = _dist()
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def shutdown():
    dist = _dist()
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()


def shard_range(num_chains, rank, world):
    """Contiguous near-equal shard [lo, hi) of the chain axis for `rank`.  Global
    chain ids key the Philox streams, so results do not depend on `world`."""
    base, rem = divmod(int(num_chains), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _as_tensor(arr, device):
    import torch
    t = torch.as_tensor(np.ascontiguousarray(arr))
    return t.to(device) if device is not None else t


def gather_chains(local, device=None):
    """all_gather of per-chain arrays [C_local, ...] (ragged shards allowed) -> numpy [C, ...] on every rank."""
    import torch
    dist = _dist()
    local = np.ascontiguousarray(local)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    n = torch.tensor([local.shape[0]], dtype=torch.int64, device=device if device is not None else "cpu")
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    nmax = max(sizes)
    pad = np.zeros((nmax,) + local.shape[1:], dtype=local.dtype)
    pad[:local.shape[0]] = local
    t = _as_tensor(pad, device)
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t)
    return np.concatenate([p.cpu().numpy()[:s] for p, s in zip(parts, sizes)], axis=0)


def sum_scalar(x, device=None):
    import torch
    dist = _dist()
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def is_multi():
    dist = _dist()
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def rhat_stats(chain_mean, chain_var):
    """Per-chain moments [C_local, D] (torch tensors, any device) -> packed sufficient statistics [3 D + 1] fp64:
    (sum_c mean, sum_c mean^2, sum_c var, n_chains).  Sums over chains, so shards add."""
    import torch
    m, v = chain_mean.double(), chain_var.double()
    n = torch.full((1,), float(m.shape[0]), dtype=torch.float64, device=m.device)
    return torch.cat([m.sum(0), (m * m).sum(0), v.sum(0), n])


def rhat_from_stats(packed, num_samples):
    """Packed statistics (after the all-reduce) -> R-hat [D] (torch, fp64)."""
    import torch
    D = (packed.numel() - 1) // 3
    s1, s2, sv, n = packed[:D], packed[D:2 * D], packed[2 * D:3 * D], packed[-1]
    s = float(num_samples)
    w = sv / n * s / (s - 1.0)
    b_over_n = (s2 - s1 * s1 / n) / (n - 1.0)
    return torch.sqrt(((s - 1.0) / s * w + b_over_n) / w)


def allreduce_sum_(t):
    """In-place sum over ranks of a tensor living on this rank's device (NCCL over NVLink on GPUs, gloo on CPU);
    a no-op without a process group."""
    if is_multi():
        dist = _dist()
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def rhat_allreduce(chain_mean, chain_var, num_samples, device=None):
    """R-hat from per-chain moments held shard-wise: all_reduce(sum) of
    (sum_c mean, sum_c mean^2, sum_c var, n_chains) per coordinate -- 3D+1 numbers."""
    import torch
    dist = _dist()
    m = np.asarray(chain_mean, dtype=np.float64)
    v = np.asarray(chain_var, dtype=np.float64)
    packed = np.concatenate([m.sum(0), (m * m).sum(0), v.sum(0), [float(m.shape[0])]])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        t = _as_tensor(packed, device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        packed = t.cpu().numpy()
    D = m.shape[1]
    s1, s2, sv, n = packed[:D], packed[D:2 * D], packed[2 * D:3 * D], packed[-1]
    s = float(num_samples)
    w = sv / n * s / (s - 1.0)
    b_over_n = (s2 - s1 * s1 / n) / (n - 1.0)
    return np.sqrt(((s - 1.0) / s * w + b_over_n) / w)
