"""Launch plumbing: binds the library's rank context and NCCL communicator to a torch.distributed job.

One process per GPU (torchrun).  torch.distributed is used for exactly two things: a gloo group that
implements the set-up allgather of make_local_matrix (hpccg_ctx_set_allgather), and the broadcast of
the ncclUniqueId with which the library creates its OWN communicator for the CG loop (halo send/recv
and scalar gathers are issued by the C++ loop, not by Python).
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from . import _capi
from ._capi import check, lib

_keepalive = []  # ctypes callbacks must outlive their registration


def install_allgather(group=None) -> None:
    """Installs a torch.distributed implementation of the set-up collective on this thread's context."""
    size = dist.get_world_size(group)

    def _allgather(_user, send, nbytes, recv):
        try:
            src = torch.frombuffer((C.c_char * max(nbytes, 1)).from_address(send), dtype=torch.uint8)[:nbytes].clone()
            outs = [torch.empty(nbytes, dtype=torch.uint8) for _ in range(size)]
            dist.all_gather(outs, src, group=group)
            for r, t in enumerate(outs):
                if nbytes:
                    C.memmove(recv + r * nbytes, t.data_ptr(), nbytes)
            return 0
        except Exception as e:  # noqa: BLE001 - reported through the C return code
            print(f"hpccg set-up allgather failed: {e}", flush=True)
            return 1

    cb = _capi.ALLGATHER_FN(_allgather)
    _keepalive.append(cb)
    check(lib.hpccg_ctx_set_allgather(cb, None), "hpccg_ctx_set_allgather")


def init_process_group_context(use_nccl: bool = True):
    """Call after torch.distributed.init_process_group.  Returns (rank, size, gloo_group)."""
    rank, size = dist.get_rank(), dist.get_world_size()
    check(lib.hpccg_ctx_set(rank, size), "hpccg_ctx_set")
    gloo = dist.new_group(backend="gloo") if dist.get_backend() != "gloo" else None
    if size > 1:
        install_allgather(gloo)
    if use_nccl and size > 1:
        ident = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = (C.c_char * 128)()
            check(lib.hpccg_nccl_unique_id(buf), "hpccg_nccl_unique_id")
            ident = torch.frombuffer(buf, dtype=torch.uint8).clone()
        dist.broadcast(ident, src=0, group=gloo)
        raw = (C.c_char * 128).from_buffer_copy(ident.numpy().tobytes())
        check(lib.hpccg_nccl_init(raw, rank, size), "hpccg_nccl_init")
    return rank, size, gloo


def bind_to_gpu_numa_node(local_rank: int) -> dict:
    """Pins this process (and so the first-touch placement of the host vectors it allocates afterwards) to the CPUs of the
    NUMA node its GPU hangs off.  With one process per GPU and page-locked host vectors of GBs, host<->device copies
    that cross the socket interconnect are what limits HPCCG()'s end-to-end rate when all ranks copy at once.
    Best effort: returns what it did; never raises."""
    import os
    info = {"bound": False}
    try:
        prop = torch.cuda.get_device_properties(local_rank)
        bdf = f"{getattr(prop, 'pci_domain_id', 0):04x}:{prop.pci_bus_id:02x}:{prop.pci_device_id:02x}.0"
        info["pci"] = bdf
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
        info["numa_node"] = node
        if node < 0:
            return info
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = cpus & set(os.sched_getaffinity(0))
        if not allowed:
            return info
        os.sched_setaffinity(0, allowed)
        info["bound"] = True
        info["cpus"] = len(allowed)
    except Exception as e:  # noqa: BLE001 - placement is an optimisation, not a requirement
        info["error"] = str(e)[:120]
    return info


def finalize() -> None:
    lib.hpccg_nccl_finalize()
