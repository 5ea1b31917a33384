"""Multi-GPU plumbing: trials shard across ranks, tallies combine with one allreduce.

One process per GPU (torchrun).  Monte-Carlo trials are independent and every trial's bit stream
is keyed by its *global* trial id, so rank r of W simply takes the contiguous id range
``shard_range(total, r, W)`` of every sweep point and the summed tallies are bit-identical for
any W.  The only data-path collective is a single ``all_reduce(SUM)`` of the int64 tally vector
(NCCL over NVLink on GPUs; gloo in the CPU tests).  The reference has no equivalent: its trial
loop is serial (Pd_plotter.py:198-223).

Under ``torchrun ... Pd_plotter.py`` nobody has called ``init_process_group`` when
``run_experiment`` starts: :func:`ensure_init` does it lazily from the torchrun environment
(``WORLD_SIZE`` > 1), NCCL when a CUDA device is present, gloo otherwise.
"""
from __future__ import annotations

import os
from typing import Tuple

import numpy as np


def ensure_init() -> None:
    """Join the torchrun job if there is one and no process group exists yet (idempotent)."""
    ws = int(os.environ.get("WORLD_SIZE", "1") or 1)
    if ws <= 1:
        return
    import torch
    import torch.distributed as dist
    if not dist.is_available() or dist.is_initialized():
        return
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29500")
    import datetime
    timeout = datetime.timedelta(seconds=int(os.environ.get("MVD_DIST_TIMEOUT_S", "300")))
    if torch.cuda.is_available():
        local = int(os.environ.get("LOCAL_RANK", 0))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=timeout)
    else:
        dist.init_process_group("gloo", timeout=timeout)
    import atexit
    atexit.register(_shutdown)                # the group this module created is torn down at interpreter exit


def _shutdown() -> None:
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            dist.destroy_process_group()
    except Exception:
        pass


def world() -> Tuple[int, int]:
    """(rank, world_size): of the torch.distributed group, joining the torchrun job first if its
    environment is present (``WORLD_SIZE`` > 1); (0, 1) for a plain single process."""
    try:
        ensure_init()
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except ImportError:
        pass
    return 0, 1


def is_rank0() -> bool:
    return world()[0] == 0


def backend() -> str:
    import torch.distributed as dist
    return dist.get_backend() if dist.is_available() and dist.is_initialized() else "none"


def shard_range(total: int, rank: int, world_size: int, offset: int = 0) -> Tuple[int, int]:
    """Contiguous, balanced split of ``[offset, offset + total)``; the first ``total % W`` ranks
    get one extra trial."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank / world size")
    base, extra = divmod(int(total), world_size)
    begin = rank * base + min(rank, extra)
    end = begin + base + (1 if rank < extra else 0)
    return offset + begin, offset + end


def allreduce_sum(values: np.ndarray, device: str | None = None) -> np.ndarray:
    """Sum an integer vector that lives on the host over all ranks (no-op for a single process)."""
    rank, ws = world()
    arr = np.asarray(values)
    if ws == 1:
        return arr
    import torch
    import torch.distributed as dist

    t = torch.from_numpy(arr.astype(np.int64))
    if dist.get_backend() == "nccl":
        dev = device or f"cuda:{int(os.environ.get('LOCAL_RANK', 0))}"
        t = t.to(dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy().astype(arr.dtype)


def allreduce_sum_device(tensor):
    """In-place SUM allreduce of a device tensor (tallies left on the GPU by ``mvd_detect``)."""
    _, ws = world()
    if ws > 1:
        import torch.distributed as dist
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM)
    return tensor
