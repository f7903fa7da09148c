"""Shared-dof (halo) exchange for element-partitioned meshes: QQ^T across GPUs.

Reference semantics: `gather_scatter.exchange` under `pmap`
(`swirl_fem/core/gather_scatter.py:221-261`) sums, for every global dof that
lives on more than one partition, the copies held by all partitions, via ONE
dense `lax.psum` over all S shared dofs of all partitions (:246-248).

B200 design: one process per GPU; each pair of ranks exchanges only the dofs
the two of them share (face / edge / corner sets), ordered by global id on
both sides so no index list travels.  pack kernel -> NCCL send/recv over
NVLink (`torch.distributed.batch_isend_irecv`, one group) -> unpack-add kernel
in ascending peer order (deterministic).  The result equals the reference's
QQ^T; parity is checked against the unpartitioned oracle.
"""

from __future__ import annotations

import dataclasses

import numpy as np
import torch

from swirl_fem_b200 import _lib

SENTINEL = -1


def default_rank() -> int:
  import torch.distributed as dist  # pylint: disable=g-import-not-at-top
  return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


@dataclasses.dataclass
class HaloPlan:
  """Per-rank exchange lists.

  Attributes:
    rank, world: this partition and the number of partitions.
    peers: ascending ranks this rank shares dofs with.
    local_idx: peer -> int32 local node indices of the shared dofs, ordered by
      ascending global id (the peer's list for us has the same order).
    owned: bool (num_local_nodes,), True where this rank is the lowest rank
      holding the dof -- weights for globally consistent dot products.
  """
  rank: int
  world: int
  peers: list
  local_idx: dict
  owned: np.ndarray
  group: object = None
  _dev: dict = dataclasses.field(default_factory=dict, repr=False)

  @classmethod
  def from_node_indices(cls, node_indices: np.ndarray, rank: int) -> 'HaloPlan':
    """Builds the plan from `(P, n_max)` global ids (SENTINEL padded)."""
    node_indices = np.asarray(node_indices)
    world = len(node_indices)
    mine = node_indices[rank]
    mine_valid = mine[mine != SENTINEL]
    if len(np.unique(mine_valid)) != len(mine_valid):
      raise NotImplementedError(
          'a global dof occurs more than once in one partition '
          '(intra-partition periodicity is not supported)')
    order = np.argsort(mine_valid, kind='stable')
    sorted_ids = mine_valid[order]
    peers, local_idx = [], {}
    owned = np.ones(len(mine_valid), dtype=bool)
    for q in range(world):
      if q == rank:
        continue
      theirs = node_indices[q]
      theirs = theirs[theirs != SENTINEL]
      shared = np.intersect1d(sorted_ids, theirs, assume_unique=True)
      if not len(shared):
        continue
      pos = order[np.searchsorted(sorted_ids, shared)]
      peers.append(q)
      local_idx[q] = pos.astype(np.int32)
      if q < rank:
        owned[pos] = False
    return cls(rank=rank, world=world, peers=peers, local_idx=local_idx,
               owned=owned)

  # -- device path ---------------------------------------------------------
  def _device_lists(self, device, dtype):
    key = (str(device), dtype)
    if key not in self._dev:
      idx = {q: torch.as_tensor(v).to(device) for q, v in self.local_idx.items()}
      send = {q: torch.empty(len(v), dtype=dtype, device=device)
              for q, v in self.local_idx.items()}
      recv = {q: torch.empty(len(v), dtype=dtype, device=device)
              for q, v in self.local_idx.items()}
      self._dev[key] = (idx, send, recv)
    return self._dev[key]

  # The two device steps; the CPU (gloo) tests of the exchange *protocol*
  # override them, the product path is CUDA only.
  def _pack(self, u, idx, buf):
    _lib.halo_pack(u, idx, buf)

  def _unpack_add(self, u, idx, buf):
    _lib.halo_unpack_add(u, idx, buf)

  def _flat_lists(self, device, dtype):
    """Concatenated (all peers) index list, buffers and split sizes."""
    key = ('flat', str(device), dtype)
    if key not in self._dev:
      splits = [len(self.local_idx[q]) if q in self.local_idx else 0
                for q in range(self.world)]
      cat = (np.concatenate([self.local_idx[q] for q in self.peers])
             if self.peers else np.zeros(0, np.int32))
      idx = torch.as_tensor(cat.astype(np.int32)).to(device)
      send = torch.empty(len(cat), dtype=dtype, device=device)
      recv = torch.empty(len(cat), dtype=dtype, device=device)
      offs = np.concatenate([[0], np.cumsum(splits)])
      self._dev[key] = (idx, send, recv, splits, offs)
    return self._dev[key]

  def canonical_csr(self):
    """CSR (dofs, row_ptr, src) of the canonical unpack.

    For every unique interface dof: its contributions in ascending RANK order,
    `src >= 0` = position in the concatenated receive buffer, `src = -1` = this
    rank's own value.
    """
    if 'csr' not in self._dev:
      entries = []  # (dof, rank, src)
      off = 0
      for q in self.peers:
        loc = self.local_idx[q].astype(np.int64)
        entries.append(np.stack([loc, np.full(len(loc), q),
                                 off + np.arange(len(loc))], axis=1))
        off += len(loc)
      allp = (np.concatenate(entries) if entries
              else np.zeros((0, 3), np.int64))
      dofs = np.unique(allp[:, 0])
      own = np.stack([dofs, np.full(len(dofs), self.rank),
                      np.full(len(dofs), -1)], axis=1)
      allp = np.concatenate([allp, own])
      order = np.lexsort((allp[:, 1], allp[:, 0]))  # by dof, then rank
      allp = allp[order]
      counts = np.bincount(np.searchsorted(dofs, allp[:, 0]),
                           minlength=len(dofs))
      row_ptr = np.concatenate([[0], np.cumsum(counts)])
      self._dev['csr'] = (dofs.astype(np.int32), row_ptr.astype(np.int32),
                          allp[:, 2].astype(np.int32))
    return self._dev['csr']

  def _canonical_device(self, device):
    key = ('csr', str(device))
    if key not in self._dev:
      self._dev[key] = tuple(torch.as_tensor(a).to(device)
                             for a in self.canonical_csr())
    return self._dev[key]

  def _unpack_canonical(self, u, recv):
    dofs, row_ptr, src = self._canonical_device(u.device)
    _lib.halo_unpack_canonical(u, dofs, row_ptr, src, recv)

  def exchange_(self, u: torch.Tensor) -> torch.Tensor:
    """In-place QQ^T on this rank's `(num_local_nodes,)` vector.

    One pack kernel for all peers, ONE `all_to_all_single` (NCCL grouped
    send/recv; the host never blocks), one canonical unpack kernel: every
    holder of a dof sums all contributions in ascending rank order, so the
    replicated values are bitwise identical on all ranks.
    """
    import torch.distributed as dist  # pylint: disable=g-import-not-at-top
    if not self.peers:
      return u
    idx, send, recv, splits, _ = self._flat_lists(u.device, u.dtype)
    self._pack(u, idx, send)
    dist.all_to_all_single(recv, send, output_split_sizes=splits,
                           input_split_sizes=splits, group=self.group)
    self._unpack_canonical(u, recv)
    return u

  # -- split form, used to overlap the wire time with interior compute -------
  def side_stream(self, device):
    key = ('stream', str(device))
    if key not in self._dev:
      self._dev[key] = torch.cuda.Stream(device=device)
    return self._dev[key]

  def start_exchange(self, u: torch.Tensor):
    """Pack + all_to_all on the CURRENT stream (call it on a side stream)."""
    import torch.distributed as dist  # pylint: disable=g-import-not-at-top
    idx, send, recv, splits, _ = self._flat_lists(u.device, u.dtype)
    self._pack(u, idx, send)
    dist.all_to_all_single(recv, send, output_split_sizes=splits,
                           input_split_sizes=splits, group=self.group)

  def finish_exchange(self, u: torch.Tensor):
    """Unpack-add of the received values (after `start_exchange` completed)."""
    _, _, recv, _, _ = self._flat_lists(u.device, u.dtype)
    self._unpack_canonical(u, recv)
    return u

  def exchange(self, u: torch.Tensor) -> torch.Tensor:
    return self.exchange_(u.contiguous().clone())

  def owned_mask(self, device, dtype=torch.uint8) -> torch.Tensor:
    """Device copy of `owned` (cached: it is read by every CG update)."""
    key = ('owned', str(device), dtype)
    if key not in self._dev:
      self._dev[key] = torch.as_tensor(self.owned).to(
          device=device, dtype=dtype).contiguous()
    return self._dev[key]
