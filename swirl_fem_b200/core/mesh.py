"""Device-resident mesh: node coordinates, connectivity, masks, exchange maps.

Same attributes and methods as the reference's `Mesh`
(`swirl_fem/core/mesh.py`: fields :75-88, `create` :90-133, `gather` :155-160,
`scatter` :165-168, `element_coords` :170-172, `exchange` :174-179).  Arrays
are CUDA tensors (torch is the device-memory plumbing); `gather`, `scatter`
and `exchange` launch the kernels of `csrc/sfem_gs.cu` through the C ABI.
"""

from __future__ import annotations

import dataclasses
from collections.abc import Mapping
from typing import Any

import numpy as np
import torch

from swirl_fem_b200 import _lib
from swirl_fem_b200.core import gather_scatter
from swirl_fem_b200.core.interpolation import Nodes1D
from swirl_fem_b200.core.interpolation import NodeType


def _default_device(device=None) -> torch.device:
  if device is not None:
    return torch.device(device)
  if not torch.cuda.is_available():
    raise _lib.SwirlB200Error(
        'swirl_fem_b200 has no CPU path: a CUDA device is required to place a '
        'Mesh (torch.cuda.is_available() is False)')
  return torch.device('cuda', torch.cuda.current_device())


def _to_device(x, device, dtype=None):
  if x is None:
    return None
  if isinstance(x, torch.Tensor):
    t = x.to(device)
  else:
    t = torch.as_tensor(np.ascontiguousarray(x)).to(device)
  if dtype is not None:
    t = t.to(dtype)
  return t.contiguous()


@dataclasses.dataclass(frozen=True)
class Mesh:
  """An N-dimensional tensor-product (line / quad / hex) mesh on one GPU."""

  node_coords: torch.Tensor          # (num_nodes, ndim) float
  elements: torch.Tensor             # (num_elements, nodes/element) int32
  node_indices: torch.Tensor         # (num_nodes,) unique (periodic-deduped) id
  order: int
  gridpoints_1d: Nodes1D
  physical_masks: Mapping[str, torch.Tensor] = dataclasses.field(
      default_factory=dict)
  exchange_gather_indices: torch.Tensor | None = None
  exchange_unique_indices: torch.Tensor | None = None
  axis_name: str | None = None
  halo_plan: Any = None              # communication.halo.HaloPlan (partitioned)
  _cache: dict = dataclasses.field(default_factory=dict, repr=False,
                                   compare=False)

  @classmethod
  def create(cls, node_coords, elements, node_indices=None, gridpoints_1d=None,
             physical_masks=None, exchange_gather_indices=None,
             exchange_unique_indices=None, axis_name=None, halo_plan=None,
             device=None, dtype=None) -> 'Mesh':
    """Creates a `Mesh`; host arrays are copied to the GPU."""
    shape = tuple(node_coords.shape)
    eshape = tuple(elements.shape)
    ndim = shape[-1]
    num_nodes_per_element = eshape[-1]
    if gridpoints_1d is None:
      num_points = int(round(np.exp(np.log(num_nodes_per_element) / ndim)))
      gridpoints_1d = Nodes1D.create(num_points=num_points,
                                     node_type=NodeType.NEWTON_COTES)
    if num_nodes_per_element != gridpoints_1d.num_points ** ndim:
      raise ValueError(
          'Expected the number of nodes in each element of `mesh` to be equal '
          f'to the number of gridpoints in {ndim} dimensions. But got '
          f'{num_nodes_per_element} != {gridpoints_1d.num_points} ** {ndim}.')
    device = _default_device(device)
    coords = _to_device(node_coords, device)
    if not coords.is_floating_point():
      coords = coords.to(torch.float64)
    if dtype is not None:
      coords = coords.to(dtype)
    elems = _to_device(elements, device, torch.int32)
    if node_indices is None:
      nidx = torch.arange(shape[0], dtype=torch.int32, device=device)
    else:
      nidx = _to_device(node_indices, device, torch.int32)
    masks = {k: _to_device(v, device, torch.bool)
             for k, v in (physical_masks or {}).items()}
    return cls(
        node_coords=coords, elements=elems, node_indices=nidx,
        order=gridpoints_1d.num_points - 1, gridpoints_1d=gridpoints_1d,
        physical_masks=masks,
        exchange_gather_indices=_to_device(exchange_gather_indices, device,
                                           torch.int32),
        exchange_unique_indices=_to_device(exchange_unique_indices, device,
                                           torch.int32),
        axis_name=axis_name, halo_plan=halo_plan)

  @property
  def ndim(self) -> int:
    return self.node_coords.shape[-1]

  @property
  def num_nodes(self) -> int:
    return self.node_coords.shape[-2]

  @property
  def num_elements(self) -> int:
    return self.elements.shape[-2]

  @property
  def num_nodes_per_element(self) -> int:
    return self.elements.shape[-1]

  @property
  def device(self) -> torch.device:
    return self.node_coords.device

  def gather(self, u: torch.Tensor) -> torch.Tensor:
    """Global (G,) -> element-local (E, n) values; SENTINEL slots read 0."""
    if tuple(u.shape) != (self.num_nodes,):
      raise ValueError(f'Expected `u` to have shape ({self.num_nodes},) but '
                       f'got: {tuple(u.shape)}.')
    return gather_scatter.gather(u, indices=self.elements, fill_value=0.)

  def scatter(self, u_local: torch.Tensor, deterministic: bool = False):
    """Element-local (E, n) -> global (G,) sum (direct stiffness summation).

    `deterministic=True` uses the transposed-map warp-segmented reduction
    (bitwise reproducible, no atomics) instead of RED atomics.
    """
    if deterministic:
      plan = self._cache.get('scatter_plan')
      if plan is None:
        plan = _lib.ScatterPlan(self.elements, self.num_nodes)
        self._cache['scatter_plan'] = plan
      return plan(u_local)
    return gather_scatter.scatter(u_local, indices=self.elements,
                                  num_nodes=self.num_nodes)

  def element_coords(self) -> torch.Tensor:
    """(E, n, ndim) coordinates of the nodes of every element."""
    cols = [self.gather(self.node_coords[:, k].contiguous())
            for k in range(self.ndim)]
    return torch.stack(cols, dim=-1)

  def exchange(self, u: torch.Tensor) -> torch.Tensor:
    """QQ^T: every shared (periodic / partition-interface) dof gets the sum."""
    if self.axis_name is not None:
      if self.halo_plan is None:
        raise ValueError('partitioned mesh without a halo plan')
      return self.halo_plan.exchange(u)
    return gather_scatter.exchange(
        u, gather_indices=self.exchange_gather_indices,
        unique_indices=self.exchange_unique_indices, axis_name=None)
