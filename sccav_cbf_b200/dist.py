"""One process per GPU, scenario shards, no collective on the hot path.

Every vehicle / scenario is independent (own state, own obstacle set; the course is a small
read-only broadcast), so N GPUs simply own N contiguous blocks of the scenario axis
(``scenarios.shard_range``).  ``torch.distributed`` is used for exactly three things, all OFF the
timed kernels: the barrier that brackets a timed region, max / sum reductions of per-rank
scalars (device time, solve counts), and the final gather of per-shard result summaries.
Backend "nccl" on the GPU box, "gloo" in the CPU tests (tests/test_multirank_gloo.py).
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional

import torch


def bind_to_gpu_numa_node(device_index: int) -> Optional[List[int]]:
    """Pin the calling process to the CPUs closest to GPU ``device_index`` (NVML's ideal affinity: its NUMA node), so that the
    pinned host buffers allocated afterwards are local to the socket the GPU's PCIe link hangs off.  With one process per
    GPU and eight of them uploading 36 MB per step, remote-socket buffers show in the end-to-end number.  Best effort:
    returns the CPU list, or None when NVML / the affinity call is unavailable (nothing else changes)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        props = torch.cuda.get_device_properties(device_index)
        bus = "%08x:%02x:%02x.0" % (getattr(props, "pci_domain_id", 0), props.pci_bus_id, props.pci_device_id)
        h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1]
        allowed = set(os.sched_getaffinity(0))
        cpus = [c for c in cpus if c in allowed]
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None


class Shards:
    def __init__(self, backend: Optional[str] = None, device: Optional[torch.device] = None):
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.device = device
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            if not dist.is_initialized():
                backend = backend or ("nccl" if (device is not None and device.type == "cuda") else "gloo")
                kw = {"device_id": device} if backend == "nccl" and device is not None else {}
                dist.init_process_group(backend, **kw)
            self.dist = dist

    # reductions of python scalars (through a 1-element tensor on the rank's device)
    def _t(self, x: float) -> torch.Tensor:
        dev = self.device if (self.device is not None and self.dist is not None and self.dist.get_backend() == "nccl") else "cpu"
        return torch.tensor([x], dtype=torch.float64, device=dev)

    def barrier(self) -> None:
        if self.dist is not None:
            self.dist.barrier()
        if self.device is not None and self.device.type == "cuda":
            torch.cuda.synchronize(self.device)

    def max(self, x: float) -> float:
        if self.dist is None:
            return float(x)
        t = self._t(x)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum(self, x: float) -> float:
        if self.dist is None:
            return float(x)
        t = self._t(x)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def gather_summaries(self, summary: Dict[str, torch.Tensor]) -> Optional[Dict[str, torch.Tensor]]:
        """Final result gather (after timing stops): per-shard summary tensors (vehicle axis last)
        are concatenated in rank order on rank 0; other ranks get None.  Shards may differ in size."""
        if self.dist is None:
            return {k: v.detach().cpu() for k, v in summary.items()}
        objs: List[Optional[Dict[str, torch.Tensor]]] = [None] * self.world if self.rank == 0 else None
        self.dist.gather_object({k: v.detach().cpu() for k, v in summary.items()}, objs, dst=0)
        if self.rank != 0:
            return None
        return {k: torch.cat([o[k] for o in objs], dim=-1) for k in summary}

    def close(self) -> None:
        if self.dist is not None and self.dist.is_initialized():
            self.dist.destroy_process_group()
