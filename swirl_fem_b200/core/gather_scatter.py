"""Global <-> local (element / shared-dof) maps.

Public surface follows the reference's `swirl_fem/core/gather_scatter.py`:
device ops `gather` (:121-127), `scatter` (:130-133), `exchange` (:189-261)
and the host index builders `get_unique_node_indices` (:136-160),
`get_exchange_indices` (:166-186, :284-358), `group_by_partitions`
(:369-396), `get_local_elements` (:399-445).

Differences by design:
  * the device ops launch hand-written sm_100a kernels through the C ABI
    (`include/swirl_b200.h`); there is no CPU fallback -- without the CUDA
    library they raise;
  * the index builders are vectorised numpy (sort / unique / connected
    components) instead of Python `Counter`/`dict` loops, but produce
    bit-identical int32 arrays (pinned by tests/golden/connectivity.npz and
    partition.npz).
"""

from __future__ import annotations

import numpy as np
import scipy.sparse
import scipy.sparse.csgraph

SENTINEL = -1


# ----------------------------------------------------------------------------
# host index builders (integer work: bit-exact contract)
# ----------------------------------------------------------------------------


def _periodic_representatives(periodic_links: np.ndarray) -> tuple[
    np.ndarray, np.ndarray]:
  """Returns (sorted linked node ids, representative = min id of component)."""
  pairs = np.transpose(np.asarray(periodic_links), (0, 2, 1)).reshape(-1, 2)
  ids = np.unique(pairs)
  a = np.searchsorted(ids, pairs[:, 0])
  b = np.searchsorted(ids, pairs[:, 1])
  n = len(ids)
  graph = scipy.sparse.coo_matrix(
      (np.ones(len(a), dtype=np.int8), (a, b)), shape=(n, n))
  _, labels = scipy.sparse.csgraph.connected_components(graph, directed=False)
  rep = np.full(labels.max() + 1, np.iinfo(np.int64).max, dtype=np.int64)
  np.minimum.at(rep, labels, ids.astype(np.int64))
  return ids, rep[labels]


def get_unique_node_indices(node_indices: np.ndarray,
                            periodic_links: np.ndarray | None) -> np.ndarray:
  """Dedups node indices through periodic links (representative = min id)."""
  if periodic_links is None or len(periodic_links) == 0:
    return node_indices
  node_indices = np.asarray(node_indices)
  ids, rep = _periodic_representatives(periodic_links)
  pos = np.searchsorted(ids, node_indices)
  pos = np.clip(pos, 0, len(ids) - 1)
  hit = ids[pos] == node_indices
  # np.vectorize in the reference yields the platform integer type
  return np.where(hit, rep[pos], node_indices).astype(np.int64)


def _exchanged_ids(node_indices: np.ndarray) -> np.ndarray:
  """Sorted ids (excluding SENTINEL) that occur more than once."""
  flat = np.asarray(node_indices).reshape(-1)
  ids, counts = np.unique(flat[flat != SENTINEL], return_counts=True)
  return ids[counts > 1]


def get_exchange_indices(node_indices: np.ndarray):
  """Returns `(gather_indices, unique_indices)` for `exchange`."""
  node_indices = np.asarray(node_indices)
  if node_indices.ndim not in (1, 2):
    raise ValueError('node_indices must have ndim 1 or 2. Got '
                     f'{node_indices.ndim}')
  shared = _exchanged_ids(node_indices)
  if node_indices.ndim == 1:
    pos = np.searchsorted(shared, node_indices)
    pos = np.clip(pos, 0, max(len(shared) - 1, 0))
    hit = (shared[pos] == node_indices) if len(shared) else np.zeros(
        node_indices.shape, dtype=bool)
    gather_indices = np.nonzero(hit)[0].astype(np.int32)
    unique_indices = pos[hit].astype(np.int32)
    return gather_indices, unique_indices

  num_partitions = len(node_indices)
  gather_indices = np.full((num_partitions, len(shared)), SENTINEL,
                           dtype=np.int64)
  for p in range(num_partitions):
    row = node_indices[p]
    pos = np.searchsorted(shared, row)
    pos = np.clip(pos, 0, max(len(shared) - 1, 0))
    hit = (shared[pos] == row) if len(shared) else np.zeros(row.shape, bool)
    ranks = pos[hit]
    if len(np.unique(ranks)) != len(ranks):
      order = np.argsort(ranks, kind='stable')
      dup = ranks[order][1:][np.diff(ranks[order]) == 0][0]
      raise NotImplementedError(
          f'Found node_idx={shared[dup]} occurring more than once in '
          f'partition_idx={p}')
    gather_indices[p, ranks] = np.nonzero(hit)[0]
  return gather_indices, None


def _pad_evenly(indices):
  n = max(map(len, indices))
  return [np.hstack([i, np.full(n - len(i), fill_value=SENTINEL)])
          for i in indices]


def group_by_partitions(partitions: np.ndarray) -> np.ndarray:
  """`indices[p]` = ascending element ids with `partitions[i] == p`, padded."""
  partitions = np.asarray(partitions)
  assert partitions.ndim == 1, partitions.shape
  num_partitions = 1 + int(partitions.max())
  order = np.argsort(partitions, kind='stable')
  counts = np.bincount(partitions, minlength=num_partitions)
  splits = np.split(order, np.cumsum(counts)[:-1])
  return np.array(_pad_evenly(splits), dtype=np.int32)


def get_local_elements(elements: np.ndarray):
  """Renumbers partitioned elements to partition-local node ids.

  Local numbering = ascending global id within the partition (np.unique).
  Returns `(node_indices (P, n_max), local_elements)`.
  """
  elements = np.asarray(elements)
  node_lists = []
  local = []
  for part in elements:
    flat = part.reshape(-1)
    ids = np.unique(flat[flat != SENTINEL])
    pos = np.searchsorted(ids, part)
    pos = np.clip(pos, 0, max(len(ids) - 1, 0))
    loc = np.where(part != SENTINEL, pos, SENTINEL)
    node_lists.append(ids)
    local.append(loc)
  return np.stack(_pad_evenly(node_lists)), np.stack(local)


# ----------------------------------------------------------------------------
# device ops (CUDA through the C ABI; see swirl_fem_b200/_lib.py)
# ----------------------------------------------------------------------------


def gather(u, indices, fill_value=SENTINEL):
  """`out[e, n] = u[indices[e, n]]`, `fill_value` where the index is SENTINEL."""
  from swirl_fem_b200 import _lib  # pylint: disable=g-import-not-at-top
  if u.ndim != 1:
    raise ValueError(f'Expecting a rank-1 array. Got {tuple(u.shape)}')
  return _lib.gather(u, indices, fill_value)


def scatter(u, indices, num_nodes: int):
  """Zero-initialised scatter-add of `u` (same shape as `indices`)."""
  from swirl_fem_b200 import _lib  # pylint: disable=g-import-not-at-top
  assert tuple(u.shape) == tuple(indices.shape), (
      f'Got: {tuple(u.shape)} v/s {tuple(indices.shape)}')
  return _lib.scatter(u, indices, num_nodes)


def exchange(u, gather_indices, unique_indices=None, axis_name=None,
             halo=None):
  """Applies QQ^T: sums the copies of every shared dof and writes it back.

  Unpartitioned (periodic) case: one kernel pair on the device.  Partitioned
  case (`axis_name` set): `halo` is the `communication.HaloExchange` plan that
  replaces the reference's dense `lax.psum` (:246-248) by a neighbour exchange.
  """
  from swirl_fem_b200 import _lib  # pylint: disable=g-import-not-at-top
  if gather_indices is None or gather_indices.numel() == 0:
    return u
  if axis_name is not None:
    if halo is None:
      raise ValueError('partitioned exchange needs a HaloExchange plan')
    return halo.exchange(u)
  return _lib.exchange(u, gather_indices, unique_indices)
