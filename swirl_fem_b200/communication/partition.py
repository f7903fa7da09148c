"""Block partitions of structured cube meshes, one block per rank.

The reference partitions element-wise through `Premesh.partitions`
(`swirl_fem/core/premesh.py:60-63`, `unit_cube_mesh(partitions=...)`,
`swirl_fem/common/premesh_commons.py:130-138`) and materialises every
partition from the *global* mesh.  For the >=100 M-dof config that would stage
the global mesh on every rank, so here each rank builds ITS block directly
(`box_mesh` + `refine_premesh`, local numbering = the refiner's numbering of
the block) and the shared dofs are matched through integer coordinates on the
global GLL lattice.  Parity target is the unpartitioned result on the same
global mesh (SURVEY section 5 caveat).
"""

from __future__ import annotations

import dataclasses

import numpy as np

from swirl_fem_b200.common.premesh_commons import box_mesh
from swirl_fem_b200.communication.halo import HaloPlan
from swirl_fem_b200.core.interpolation import Nodes1D
from swirl_fem_b200.core.mesh_refiner import refine_premesh
from swirl_fem_b200.core.premesh import Premesh

GRID_FOR_WORLD = {
    2: {1: (1, 1), 2: (2, 1), 4: (2, 2), 8: (4, 2)},
    3: {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)},
}


@dataclasses.dataclass
class BlockPartition:
  """One rank's block of an `ne^d` cube on `[a, b]^d`."""
  premesh: Premesh              # refined block, local numbering
  dirichlet: np.ndarray         # bool (num_local_nodes,): on the GLOBAL boundary
  interface_local: np.ndarray   # int32 local ids of dofs on inter-rank faces
  interface_global: np.ndarray  # int64 global lattice ids of the same dofs
  num_global_dofs: int
  block_index: tuple
  grid: tuple
  num_interface_elements: int = 0  # leading elements that touch other ranks


def block_partition(ne: int, ndim: int, gridpoints_1d: Nodes1D, rank: int,
                    world: int, a: float = -1.0, b: float = 1.0,
                    grid: tuple | None = None) -> BlockPartition:
  """Builds rank `rank`'s block (C-order rank -> block index)."""
  grid = tuple(grid or GRID_FOR_WORLD[ndim][world])
  assert int(np.prod(grid)) == world and all(ne % g == 0 for g in grid)
  bidx = np.unravel_index(rank, grid)
  nloc = tuple(ne // g for g in grid)
  h = (b - a) / ne
  lo = tuple(a + bidx[i] * nloc[i] * h for i in range(ndim))
  hi = tuple(b if bidx[i] == grid[i] - 1 else a + (bidx[i] + 1) * nloc[i] * h
             for i in range(ndim))
  pm = box_mesh(nloc, lo, hi)
  refined = refine_premesh(pm, gridpoints_1d)
  npts = gridpoints_1d.num_points
  p = npts - 1
  lattice = ne * p + 1  # global GLL lattice points per axis

  # element-local view (block elements are C-ordered, nodes lexicographic)
  els = refined.elements.reshape(nloc + (npts,) * ndim)
  dirichlet = np.zeros(refined.num_nodes, dtype=bool)
  iface_local, iface_global = [], []
  for axis in range(ndim):
    for side in (0, 1):
      eidx = 0 if side == 0 else nloc[axis] - 1
      nidx = 0 if side == 0 else npts - 1
      sl = [slice(None)] * (2 * ndim)
      sl[axis] = eidx
      sl[ndim + axis] = nidx
      face_nodes = els[tuple(sl)]  # (other elems..., other local nodes...)
      on_global_boundary = (bidx[axis] == 0 if side == 0
                            else bidx[axis] == grid[axis] - 1)
      if on_global_boundary:
        dirichlet[face_nodes.reshape(-1)] = True
        continue
      # global lattice coordinates of the face nodes
      other = [ax for ax in range(ndim) if ax != axis]
      coords = np.empty(face_nodes.shape + (ndim,), dtype=np.int64)
      fixed = (bidx[axis] * nloc[axis] + eidx) * p + nidx
      coords[..., axis] = fixed
      for j, ax in enumerate(other):
        e_shape = [1] * face_nodes.ndim
        e_shape[j] = nloc[ax]
        n_shape = [1] * face_nodes.ndim
        n_shape[len(other) + j] = npts
        coords[..., ax] = (
            (bidx[ax] * nloc[ax] + np.arange(nloc[ax])).reshape(e_shape) * p +
            np.arange(npts).reshape(n_shape))
      gid = np.ravel_multi_index(
          tuple(coords[..., ax].reshape(-1) for ax in range(ndim)),
          (lattice,) * ndim)
      iface_local.append(face_nodes.reshape(-1))
      iface_global.append(gid)
  if iface_local:
    loc = np.concatenate(iface_local)
    gid = np.concatenate(iface_global)
    loc, first = np.unique(loc, return_index=True)
    gid = gid[first]
  else:
    loc = np.zeros(0, dtype=np.int64)
    gid = np.zeros(0, dtype=np.int64)
  # Element order: the elements touching another rank's block come first (their
  # number rounded up to a multiple of 4 with interior elements), so that the
  # halo exchange can start after a launch over `[0, num_interface_elements)`
  # and overlap with the launch over the interior elements.
  touches = np.zeros(nloc, dtype=bool)
  for axis in range(ndim):
    for side in (0, 1):
      on_global_boundary = (bidx[axis] == 0 if side == 0
                            else bidx[axis] == grid[axis] - 1)
      if on_global_boundary:
        continue
      sl = [slice(None)] * ndim
      sl[axis] = 0 if side == 0 else nloc[axis] - 1
      touches[tuple(sl)] = True
  touches = touches.reshape(-1)
  order = np.concatenate([np.nonzero(touches)[0], np.nonzero(~touches)[0]])
  num_interface = int(touches.sum())
  num_interface = min(len(order), -(-num_interface // 4) * 4)
  refined = refined.replace(elements=refined.elements[order])
  return BlockPartition(
      premesh=refined, dirichlet=dirichlet,
      interface_local=loc.astype(np.int32), interface_global=gid,
      num_global_dofs=lattice ** ndim, block_index=tuple(int(i) for i in bidx),
      grid=grid, num_interface_elements=num_interface)


def halo_plan_from_interfaces(rank: int, interface_local: np.ndarray,
                              interface_global: np.ndarray,
                              all_global: list, num_local_nodes: int
                              ) -> HaloPlan:
  """Builds the pairwise plan from every rank's interface lattice ids."""
  order = np.argsort(interface_global, kind='stable')
  sorted_gid = interface_global[order]
  peers, local_idx = [], {}
  owned = np.ones(num_local_nodes, dtype=bool)
  for q, theirs in enumerate(all_global):
    if q == rank or not len(theirs):
      continue
    shared = np.intersect1d(sorted_gid, theirs, assume_unique=True)
    if not len(shared):
      continue
    pos = interface_local[order[np.searchsorted(sorted_gid, shared)]]
    peers.append(q)
    local_idx[q] = pos.astype(np.int32)
    if q < rank:
      owned[pos] = False
  return HaloPlan(rank=rank, world=len(all_global), peers=peers,
                  local_idx=local_idx, owned=owned)
