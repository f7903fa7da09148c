"""Peer-memory all-reduce of the CG scalars (no NCCL launch per dot product).

The reference's hook for distributed dot products is `dot_fn`
(`swirl_fem/linalg/cg.py:26-31`); with one process per GPU the two scalars of
an iteration (`p.Ap`, `r.z`) are all-reduced.  `ScalarExchange` does that with
`sfem_scalar_allreduce` (`csrc/sfem_halo.cu`): every rank owns a 128-byte-per-
rank CUDA-IPC region; one single-CTA kernel stores this rank's partial values
into every rank's region (NVLink), waits on the device for the others' and
adds them in ascending rank order, so all ranks obtain bitwise identical sums.
"""

from __future__ import annotations

import ctypes

import numpy as np
import torch

from swirl_fem_b200 import _lib


class ScalarExchange:
  """`sfem_scalar_exchange` handle of one rank."""

  def __init__(self, rank, world, device, region, peer_regions, mapped=()):
    self.rank, self.world, self.device = rank, world, device
    self.region, self.mapped = region, list(mapped)
    self.owns_region = True
    addrs = np.asarray(peer_regions, dtype=np.uint64)
    handle = ctypes.c_void_p()
    with torch.cuda.device(device):
      _lib._check(_lib.lib().sfem_scalar_exchange_create(
          rank, world, region, addrs.ctypes.data, ctypes.byref(handle)),
                  'sfem_scalar_exchange_create')
    self.handle = handle

  @classmethod
  def create(cls, device, group=None):
    """Collective: exports this rank's region, maps everyone else's.  Returns
    None (keep NCCL) if any rank cannot."""
    import torch.distributed as dist  # pylint: disable=g-import-not-at-top
    lib = _lib.lib()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    region, buf = ctypes.c_void_p(), ctypes.create_string_buffer(64)
    with torch.cuda.device(device):
      ok = lib.sfem_ipc_alloc(lib.sfem_scalar_region_bytes(world),
                              ctypes.byref(region), buf) == 0
    handles = [None] * world
    dist.all_gather_object(handles, (ok, bytes(buf.raw)), group=group)
    ok = all(h[0] for h in handles)
    addrs, mapped = [0] * world, []
    if ok:
      with torch.cuda.device(device):
        for q in range(world):
          if q == rank:
            addrs[q] = region.value
            continue
          ptr = ctypes.c_void_p()
          if lib.sfem_ipc_open(handles[q][1], ctypes.byref(ptr)) != 0:
            ok = False
            break
          addrs[q] = ptr.value
          mapped.append(ptr.value)
    flags = [None] * world
    dist.all_gather_object(flags, ok, group=group)
    if not all(flags):
      with torch.cuda.device(device):
        for addr in mapped:
          lib.sfem_ipc_close(addr)
        dist.barrier(group=group)
        if region.value:
          lib.sfem_ipc_free(region.value)
      return None
    out = cls(rank, world, device, region.value, addrs, mapped)
    torch.cuda.synchronize(device)
    dist.barrier(group=group)
    return out

  @staticmethod
  def create_local(world, device):
    """All ranks in ONE process on one device (tests; run the ranks'
    all-reduces on different streams)."""
    lib = _lib.lib()
    regions = []
    with torch.cuda.device(device):
      for _ in range(world):
        region, buf = ctypes.c_void_p(), ctypes.create_string_buffer(64)
        _lib._check(lib.sfem_ipc_alloc(lib.sfem_scalar_region_bytes(world),
                                       ctypes.byref(region), buf),
                    'sfem_ipc_alloc')
        regions.append(region.value)
    return [ScalarExchange(r, world, device, regions[r], regions)
            for r in range(world)]

  def allreduce_(self, values: torch.Tensor) -> torch.Tensor:
    """In place: `values` (1..4 contiguous float64 on the device) <- sums."""
    _lib.require_cuda(values)
    if values.dtype != torch.float64 or not values.is_contiguous():
      raise ValueError('scalar exchange works on contiguous float64 values')
    if not 1 <= values.numel() <= 4:
      raise ValueError('1 to 4 values per all-reduce')
    with torch.cuda.device(values.device):
      _lib._check(_lib.lib().sfem_scalar_allreduce(
          self.handle, _lib.ptr(values), values.numel(),
          _lib.stream_ptr(values.device)), 'sfem_scalar_allreduce')
    return values

  def close(self, group=None):
    """Releases the handle, the mapped peer regions and this rank's region.
    Collective when the regions were exchanged over IPC (`create`): peers must
    have stopped publishing before a region is freed."""
    h = getattr(self, 'handle', None)
    if not h:
      return
    lib = _lib.lib()
    with torch.cuda.device(self.device):
      torch.cuda.synchronize(self.device)
      lib.sfem_scalar_exchange_destroy(h)
      self.handle = None
      for addr in self.mapped:
        lib.sfem_ipc_close(addr)
      if self.mapped:
        import torch.distributed as dist  # pylint: disable=g-import-not-at-top
        dist.barrier(group=group)
      if self.owns_region and self.region:
        lib.sfem_ipc_free(self.region)
    self.mapped, self.region = [], None

  def timed_out(self) -> bool:
    with torch.cuda.device(self.device):
      return bool(_lib.lib().sfem_scalar_exchange_timed_out(
          self.handle, _lib.stream_ptr(self.device)))

  def __del__(self):
    h = getattr(self, 'handle', None)
    if h and _lib._lib is not None:
      _lib._lib.sfem_scalar_exchange_destroy(h)
      self.handle = None
