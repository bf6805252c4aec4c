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


def _parse_cpulist(text: str):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_host_to_gpu_numa(device_index: int, sysfs: str = "/sys") -> dict:
    """Pin this process to the CPUs of the NUMA node the GPU hangs off, BEFORE its pinned staging buffers are
    allocated (first touch then places them on that node).  With one process per GPU this keeps every rank's
    host->device stream on its own socket's memory controllers instead of crossing the inter-socket link, which is
    what bounds the end-to-end rate once several GPUs ingest audio at PCIe speed.  Best effort: returns what was done
    ({"node": n, "cpus": k}) or why nothing was ({"skipped": reason}); FT8B_NUMA_BIND=0 disables it."""
    if os.environ.get("FT8B_NUMA_BIND", "1") == "0":
        return {"skipped": "FT8B_NUMA_BIND=0"}
    try:
        pr = torch.cuda.get_device_properties(device_index)
        if all(isinstance(getattr(pr, k, None), int) for k in ("pci_domain_id", "pci_bus_id", "pci_device_id")):
            bus = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        else:
            import pynvml
            pynvml.nvmlInit()
            uuid = str(pr.uuid)
            h = pynvml.nvmlDeviceGetHandleByUUID((uuid if uuid.startswith("GPU-") else "GPU-" + uuid).encode())
            bus = pynvml.nvmlDeviceGetPciInfo(h).busId
            bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
            if len(bus.split(":")[0]) == 8:          # NVML prints an 8-digit domain, sysfs uses 4
                bus = bus[4:]
        with open(f"{sysfs}/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return {"skipped": "no NUMA affinity reported for " + bus}
        with open(f"{sysfs}/devices/system/node/node{node}/cpulist") as f:
            cpus = _parse_cpulist(f.read())
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        if not use:
            return {"skipped": f"node {node} has no CPU this process may run on"}
        os.sched_setaffinity(0, use)
        return {"node": node, "cpus": len(use), "pci": bus}
    except Exception as e:                        # containers without sysfs/NVML: leave the scheduler alone
        return {"skipped": f"{type(e).__name__}: {e}"}


def init_process_group(backend: str = None) -> Tuple[int, int, int]:
    rank, local_rank, world = env_rank_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            bind_host_to_gpu_numa(local_rank)
        kw = {"device_id": torch.device("cuda", local_rank)} if backend == "nccl" else {}
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
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
