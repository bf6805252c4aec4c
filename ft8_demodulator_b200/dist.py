"""Multi-GPU sharding: independent slots (or LLR words) split contiguously across ranks.

One process per GPU (torchrun), no data-path collective - a slot is self-contained, every rank
holds its own copy of the tables.  The only exchange is an optional gather of the
fixed-size decode records (NCCL all_gather on the GPU box, gloo in the CPU tests).
"""
import os
from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) share of `n_items` for `rank`; sizes differ by at most one, in rank order."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(int(n_items), world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def env_rank_world() -> Tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment (1 process when absent)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def init_process_group(backend: str = None) -> Tuple[int, int, int]:
    rank, local_rank, world = env_rank_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local_rank, world


# fixed-size decode record: [ok, ldpc_errors, crc_extracted, crc_calculated, abs_time, abs_freq, payload[10]] as int32
RECORD_WIDTH = 16


def pack_records(out) -> torch.Tensor:
    """SlotDecodes -> int32 [B, K, RECORD_WIDTH] (+ score kept separately as float64)."""
    B, K = out.ok.shape
    rec = torch.zeros((B, K, RECORD_WIDTH), dtype=torch.int32, device=out.ok.device)
    rec[..., 0] = out.ok.to(torch.int32)
    rec[..., 1] = out.ldpc_errors
    rec[..., 2] = out.crc_extracted
    rec[..., 3] = out.crc_calculated
    rec[..., 4] = out.cand_time
    rec[..., 5] = out.cand_freq
    rec[..., 6:16] = out.payload.to(torch.int32)
    return rec


def gather_records(rec: torch.Tensor, n_total: int, rank: int, world: int) -> torch.Tensor:
    """all_gather of per-rank record blocks into the global [n_total, K, W] tensor (slot order preserved).
    Shards may differ by one slot, so blocks are padded to the largest shard for the collective."""
    if world == 1:
        return rec
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    max_n = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((max_n,) + tuple(rec.shape[1:]), dtype=rec.dtype, device=rec.device)
    pad[:rec.shape[0]] = rec
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    return torch.cat([bufs[r][:hi - lo] for r, (lo, hi) in enumerate(sizes)], dim=0)
