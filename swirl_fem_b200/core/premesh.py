"""Host-side staging structure for meshes (numpy), finalised onto the GPU.

Follows the reference's `swirl_fem/core/premesh.py` (`Premesh` :37-139,
`finalize` :141-222, `_mask` :31-34).  `finalize()` of an unpartitioned
premesh returns a device-resident `Mesh`.  For a partitioned premesh the
reference `pmap`s `Mesh.create` over fake/real devices (:216); the B200 build
is one process per GPU, so `finalize(axis_name, rank=r)` returns *rank r's*
partition (same local numbering as the reference: ascending global id,
`get_local_elements`), together with the halo plan that replaces the dense
`psum` of `exchange`.
"""

from __future__ import annotations

import dataclasses
from collections.abc import Mapping

import numpy as np

from swirl_fem_b200.core import gather_scatter
from swirl_fem_b200.core.interpolation import Nodes1D
from swirl_fem_b200.core.interpolation import NodeType


def _mask(facets: np.ndarray, node_indices: np.ndarray) -> np.ndarray:
  """Boolean mask of which `node_indices` are contained in `facets`."""
  return np.isin(node_indices, np.unique(np.asarray(facets).reshape(-1)))


@dataclasses.dataclass(frozen=True)
class Premesh:
  """First- or high-order mesh staged in host memory."""

  order: int
  gridpoints_1d: Nodes1D
  node_coords: np.ndarray
  elements: np.ndarray
  physical_groups: Mapping[str, np.ndarray]
  periodic_links: np.ndarray | None = None
  partitions: np.ndarray | None = None

  @classmethod
  def create(cls, node_coords, elements, order=None, gridpoints_1d=None,
             physical_groups=None, periodic_links=None, partitions=None):
    node_coords = np.asarray(node_coords)
    elements = np.asarray(elements)
    ndim = node_coords.shape[-1]
    num_nodes_per_element = elements.shape[-1]
    if gridpoints_1d is None:
      num_points = int(round(np.exp(np.log(num_nodes_per_element) / ndim)))
      gridpoints_1d = Nodes1D.create(num_points=num_points,
                                     node_type=NodeType.NEWTON_COTES)
    if num_nodes_per_element != gridpoints_1d.num_points ** ndim:
      raise ValueError(
          'Expected the number of nodes in each element to be equal '
          f'to the number of gridpoints in {ndim} dimensions. But got '
          f'{num_nodes_per_element} != {gridpoints_1d.num_points} ** {ndim}.')
    if physical_groups is None:
      physical_groups = {}
    if order is None:
      order = gridpoints_1d.num_points - 1
    return cls(order=order, gridpoints_1d=gridpoints_1d,
               node_coords=node_coords, elements=elements,
               physical_groups=physical_groups, periodic_links=periodic_links,
               partitions=partitions)

  def replace(self, **changes) -> 'Premesh':
    return dataclasses.replace(self, **changes)

  @property
  def ndim(self) -> int:
    return self.node_coords.shape[-1]

  @property
  def num_nodes(self) -> int:
    return self.node_coords.shape[-2]

  @property
  def num_elements(self) -> int:
    return len(self.elements)

  @property
  def num_nodes_per_element(self) -> int:
    return self.elements.shape[-1]

  def is_partitioned(self) -> bool:
    return self.partitions is not None

  # -- host-only index products (no device needed; unit-testable on CPU) ----
  def finalize_host(self) -> dict:
    """Unpartitioned connectivity products as numpy arrays."""
    node_indices = gather_scatter.get_unique_node_indices(
        node_indices=np.arange(self.num_nodes, dtype=np.int32),
        periodic_links=self.periodic_links)
    physical_masks = {k: _mask(v, node_indices)
                      for k, v in self.physical_groups.items()}
    gi, ui = gather_scatter.get_exchange_indices(node_indices)
    return dict(node_indices=node_indices, physical_masks=physical_masks,
                exchange_gather_indices=gi, exchange_unique_indices=ui)

  def partition_host(self) -> dict:
    """Partitioned connectivity products for *all* partitions (numpy)."""
    assert self.partitions is not None
    element_indices = gather_scatter.group_by_partitions(self.partitions)
    mask = element_indices != gather_scatter.SENTINEL
    elements = np.where(mask[..., None], self.elements[element_indices],
                        gather_scatter.SENTINEL)
    node_indices, local_elements = gather_scatter.get_local_elements(elements)
    node_indices = gather_scatter.get_unique_node_indices(
        node_indices, periodic_links=self.periodic_links)
    gi, ui = gather_scatter.get_exchange_indices(node_indices)
    physical_masks = {k: _mask(v, node_indices)
                      for k, v in self.physical_groups.items()}
    return dict(element_indices=element_indices, node_indices=node_indices,
                local_elements=local_elements, exchange_gather_indices=gi,
                exchange_unique_indices=ui, physical_masks=physical_masks)

  def finalize(self, axis_name: str | None = None, rank: int | None = None,
               device=None, dtype=None):
    """Places the mesh on the GPU and returns a `Mesh`.

    Args:
      axis_name: required when partitioned (kept for API parity; names the
        process group axis).
      rank: partition to materialise when partitioned (default: this
        process's `torch.distributed` rank).
      device: torch device (default `cuda:<current>`).
      dtype: floating dtype of `node_coords` on the device (default: as given).
    """
    from swirl_fem_b200.core.mesh import Mesh  # pylint: disable=g-import-not-at-top
    if not self.is_partitioned():
      host = self.finalize_host()
      return Mesh.create(
          node_coords=self.node_coords, elements=self.elements,
          node_indices=host['node_indices'], gridpoints_1d=self.gridpoints_1d,
          physical_masks=host['physical_masks'],
          exchange_gather_indices=host['exchange_gather_indices'],
          exchange_unique_indices=host['exchange_unique_indices'],
          device=device, dtype=dtype)

    if not axis_name:
      raise ValueError('If partitioned, we need a non-trivial axis_name')
    from swirl_fem_b200.communication import halo as halo_lib  # pylint: disable=g-import-not-at-top
    if rank is None:
      rank = halo_lib.default_rank()
    host = self.partition_host()
    nidx = host['node_indices'][rank]
    valid = nidx != gather_scatter.SENTINEL
    coords = np.where(valid[:, None], self.node_coords[nidx], 0.0)
    local_elements = host['local_elements'][rank]
    keep = (local_elements != gather_scatter.SENTINEL).all(axis=-1)
    plan = halo_lib.HaloPlan.from_node_indices(host['node_indices'], rank)
    return Mesh.create(
        node_coords=coords[valid], elements=local_elements[keep],
        node_indices=nidx[valid], gridpoints_1d=self.gridpoints_1d,
        physical_masks={k: v[rank][valid]
                        for k, v in host['physical_masks'].items()},
        exchange_gather_indices=host['exchange_gather_indices'][rank],
        exchange_unique_indices=None, axis_name=axis_name, halo_plan=plan,
        device=device, dtype=dtype)
