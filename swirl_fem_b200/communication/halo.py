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

import ctypes
import dataclasses

import numpy as np
import torch

from swirl_fem_b200 import _lib

SENTINEL = -1


def _round_up(v: int, m: int) -> int:
  return -(-int(v) // m) * m


def p2p_region_layout(world: int, max_recv_entries: int, esz: int):
  """Layout of a rank's peer-mapped region (the same on every rank).

  `[flag words: 2 parities x world x 8 B | receive buffer parity 0 | parity 1]`
  Returns `(flag_bytes, parity_stride_bytes, total_bytes)`.
  """
  flag_bytes = _round_up(2 * world * 8, 256)
  stride = _round_up(max(max_recv_entries, 1) * esz, 256)
  return flag_bytes, stride, flag_bytes + 2 * stride


def p2p_tables(rank: int, peers, counts, all_splits, bases, esz: int):
  """Destination tables of the peer-memory push (pure host arithmetic).

  Args:
    rank: this rank.
    peers: ascending peer ranks.
    counts: peer -> number of dofs shared with it.
    all_splits: `all_splits[q][r]` = number of dofs ranks q and r share (every
      rank's `splits`); rank q's receive buffer is the concatenation of its
      peers' segments in ascending rank order.
    bases: peer -> address of the peer's region as mapped in this process.
    esz: bytes per value.
  Returns:
    `(send_dst uint64 (num_send,), peer_flag_addr uint64 (num_peers,))`:
    the address of every send entry's slot in its peer's parity-0 receive
    buffer, and of this rank's parity-0 flag word on every peer.
  """
  world = len(all_splits)
  flag_bytes, _, _ = p2p_region_layout(world, 0, esz)
  dst, flag_addr = [], []
  for q in peers:
    n = int(counts[q])
    if int(all_splits[q][rank]) != n:
      raise ValueError(
          f'ranks {rank} and {q} disagree on the number of shared dofs '
          f'({n} vs {all_splits[q][rank]})')
    seg = int(np.sum(np.asarray(all_splits[q][:rank], dtype=np.int64)))
    start = int(bases[q]) + flag_bytes + seg * esz
    dst.append(np.uint64(start) + np.arange(n, dtype=np.uint64) * np.uint64(esz))
    flag_addr.append(int(bases[q]) + rank * 8)
  send_dst = (np.concatenate(dst) if dst else np.zeros(0, np.uint64))
  return send_dst.astype(np.uint64), np.asarray(flag_addr, dtype=np.uint64)


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
      splits = self.splits()
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

  # -- peer-memory path (NVLink P2P stores, no NCCL on the data path) -------
  def splits(self):
    return [len(self.local_idx[q]) if q in self.local_idx else 0
            for q in range(self.world)]

  def _p2p_attach(self, dtype, device, region_ptr, bases, all_splits):
    """Builds the `sfem_halo` handle once every region address is known."""
    esz = torch.empty((), dtype=dtype).element_size()
    flag_bytes, stride, _ = p2p_region_layout(
        self.world, max(int(np.sum(sp)) for sp in all_splits), esz)
    send_dst, flag_addr = p2p_tables(
        self.rank, self.peers, {q: len(v) for q, v in self.local_idx.items()},
        all_splits, bases, esz)
    idx, _, _, _, _ = self._flat_lists(device, dtype)
    dofs, row_ptr, src = self._canonical_device(device)
    send_dst_t = torch.as_tensor(send_dst.view(np.int64)).to(device)
    peer_ranks = np.asarray(self.peers, dtype=np.int32)
    desc = _lib.HaloDesc(
        dtype=_lib.dtype_code(dtype), rank=self.rank, world=self.world,
        num_peers=len(self.peers), num_send=idx.numel(),
        send_idx=idx.data_ptr(), send_dst=send_dst_t.data_ptr(),
        parity_stride_bytes=stride,
        peer_flag_addr=flag_addr.ctypes.data,
        peer_ranks=peer_ranks.ctypes.data,
        flags=region_ptr, recv=region_ptr + flag_bytes,
        num_dofs=dofs.numel(), dofs=dofs.data_ptr(),
        row_ptr=row_ptr.data_ptr(), src=src.data_ptr())
    handle = ctypes.c_void_p()
    with torch.cuda.device(device):
      _lib._check(_lib.lib().sfem_halo_create(ctypes.byref(desc),
                                              ctypes.byref(handle)),
                  'sfem_halo_create')
    self._p2p = {'handle': handle, 'dtype': dtype, 'device': device,
                 'region': region_ptr, 'keep': (send_dst_t, idx, dofs, row_ptr,
                                                src)}

  def enable_p2p(self, dtype, device, group=None) -> bool:
    """Switches the exchange of `dtype` vectors to peer-memory stores.

    Collective: every rank of `group` must call it.  Each rank exports one
    region (CUDA IPC), maps its peers' regions and builds the destination
    tables.  Returns False (and keeps the NCCL path) if peer mapping fails on
    any rank.
    """
    import torch.distributed as dist  # pylint: disable=g-import-not-at-top
    lib = _lib.lib()
    esz = torch.empty((), dtype=dtype).element_size()
    all_splits = [None] * self.world
    dist.all_gather_object(all_splits, self.splits(), group=group)
    _, _, total = p2p_region_layout(
        self.world, max(int(np.sum(sp)) for sp in all_splits), esz)
    ok, region, handle_bytes = True, ctypes.c_void_p(), bytes(64)
    buf = ctypes.create_string_buffer(64)
    with torch.cuda.device(device):
      if lib.sfem_ipc_alloc(total, ctypes.byref(region), buf) != 0:
        ok = False
      else:
        handle_bytes = bytes(buf.raw)
    handles = [None] * self.world
    dist.all_gather_object(handles, (ok, handle_bytes), group=group)
    ok = all(h[0] for h in handles)
    bases = {}
    if ok:
      with torch.cuda.device(device):
        for q in self.peers:
          mapped = ctypes.c_void_p()
          if lib.sfem_ipc_open(handles[q][1], ctypes.byref(mapped)) != 0:
            ok = False
            break
          bases[q] = mapped.value
    flags = [None] * self.world
    dist.all_gather_object(flags, ok, group=group)
    if not all(flags):
      # some rank could not export / map: release what this rank holds and
      # keep the NCCL path everywhere
      with torch.cuda.device(device):
        for addr in bases.values():
          lib.sfem_ipc_close(addr)
        dist.barrier(group=group)
        if region.value:
          lib.sfem_ipc_free(region.value)
      return False
    self.group = group   # disable_p2p's barrier runs on the same group
    if self.peers:
      self._p2p_attach(dtype, device, region.value, bases, all_splits)
      self._p2p['mapped'] = bases
    else:
      # no peers: nothing to exchange, but the region is ours to release
      self._p2p = {'handle': None, 'dtype': None, 'device': device,
                   'region': region.value, 'mapped': {}, 'keep': ()}
    torch.cuda.synchronize(device)
    dist.barrier(group=group)
    return True

  @staticmethod
  def enable_p2p_local(plans, dtype, device):
    """All ranks in ONE process on one device (tests): regions are plain
    device allocations, peers' addresses are used directly."""
    lib = _lib.lib()
    esz = torch.empty((), dtype=dtype).element_size()
    all_splits = [pl.splits() for pl in plans]
    _, _, total = p2p_region_layout(
        len(plans), max(int(np.sum(sp)) for sp in all_splits), esz)
    regions = []
    with torch.cuda.device(device):
      for _ in plans:
        region = ctypes.c_void_p()
        buf = ctypes.create_string_buffer(64)
        _lib._check(lib.sfem_ipc_alloc(total, ctypes.byref(region), buf),
                    'sfem_ipc_alloc')
        regions.append(region.value)
    for pl in plans:
      if pl.peers:
        pl._p2p_attach(dtype, device, regions[pl.rank],
                       {q: regions[q] for q in pl.peers}, all_splits)

  def disable_p2p(self):
    """Releases the peer-memory handle, the mapped peer regions and this
    rank's region (collective in effect: peers must have stopped pushing)."""
    p = getattr(self, '_p2p', None)
    if p is None:
      return
    lib = _lib.lib()
    with torch.cuda.device(p['device']):
      torch.cuda.synchronize(p['device'])
      if p['handle'] is not None:
        lib.sfem_halo_destroy(p['handle'])
      for addr in p.get('mapped', {}).values():
        lib.sfem_ipc_close(addr)
      if 'mapped' in p:
        import torch.distributed as dist  # pylint: disable=g-import-not-at-top
        dist.barrier(group=self.group)
      lib.sfem_ipc_free(p['region'])
    self._p2p = None
    sx = self._dev.pop(('sx', str(p['device'])), None)
    if sx is not None:
      sx.close(group=self.group)

  def scalar_exchange(self, device, group=None):
    """The `ScalarExchange` of this partition (collective on first use; None
    if peer mapping is unavailable).  Cached: CG solves reuse it."""
    key = ('sx', str(device))
    if key not in self._dev:
      from swirl_fem_b200.communication.scalar_exchange import ScalarExchange  # pylint: disable=g-import-not-at-top
      self._dev[key] = ScalarExchange.create(device, group=group)
    return self._dev[key]

  def p2p_handle(self, u: torch.Tensor):
    p = getattr(self, '_p2p', None)
    if (p is None or p['handle'] is None or p['dtype'] != u.dtype
        or u.dim() != 1):
      return None
    return p['handle']

  def p2p_set_option(self, key: int, value: int):
    """0: slice size; 1: where the canonical sum of a fused apply runs (0 in
    the wait kernel after the apply, 1 in the apply kernel's own CTAs, 2 in the
    wait kernel concurrently with the apply's interior elements)."""
    p = getattr(self, '_p2p', None)
    if p is not None and p['handle'] is not None:
      _lib._check(_lib.lib().sfem_halo_set_option(p['handle'], key, value),
                  'sfem_halo_set_option')

  def p2p_push(self, u: torch.Tensor):
    with torch.cuda.device(u.device):
      _lib._check(_lib.lib().sfem_halo_push(
          self.p2p_handle(u), _lib.ptr(u), _lib.stream_ptr(u.device)),
                  'sfem_halo_push')

  def p2p_wait_unpack(self, u: torch.Tensor):
    with torch.cuda.device(u.device):
      _lib._check(_lib.lib().sfem_halo_wait_unpack(
          self.p2p_handle(u), _lib.ptr(u), _lib.stream_ptr(u.device)),
                  'sfem_halo_wait_unpack')

  def p2p_timed_out(self, device) -> bool:
    p = getattr(self, '_p2p', None)
    if p is None or p['handle'] is None:
      return False
    with torch.cuda.device(device):
      return bool(_lib.lib().sfem_halo_timed_out(p['handle'],
                                                 _lib.stream_ptr(device)))

  def p2p_debug_times(self, device):
    """us offsets (from the kernel start) of the last fused apply's stamps."""
    p = getattr(self, '_p2p', None)
    out = (ctypes.c_uint64 * 8)()
    with torch.cuda.device(device):
      _lib._check(_lib.lib().sfem_halo_debug_times(
          p['handle'], out, _lib.stream_ptr(device)), 'sfem_halo_debug_times')
    t = [int(v) for v in out]
    names = ['start', 'cta0_signal', 'cta0_ready', 'cta0_pushed',
             'flags_raised', 'cta0_exit', 'cta0_peers_ready', 'cta0_unpacked']
    return {n: (t[i] - t[0]) / 1e3 if t[i] else None
            for i, n in enumerate(names)}

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
    if self.p2p_handle(u) is not None:
      # peer-memory path: NVLink stores into the peers' buffers + flags
      self.p2p_push(u)
      self.p2p_wait_unpack(u)
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
