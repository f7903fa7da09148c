"""Shared builders for the parity tests (host-side, numpy)."""

import numpy as np

from swirl_fem_b200.common.premesh_commons import unit_cube_mesh
from swirl_fem_b200.core.interpolation import Nodes1D
from swirl_fem_b200.core.interpolation import NodeType
from swirl_fem_b200.core.mesh_refiner import refine_premesh
from swirl_fem_b200.core.premesh import Premesh

GLL = NodeType.GAUSS_LOBATTO_LEGENDRE
GL = NodeType.GAUSS_LEGENDRE
TNAME = {GLL: 'gauss_lobatto_legendre', GL: 'gauss_legendre',
         NodeType.NEWTON_COTES: 'newton_cotes'}


def deform(x):
  """Smooth non-affine map of [-1,1]^d keeping detJ > 0."""
  x = np.asarray(x, dtype=np.float64)
  ndim = x.shape[-1]
  perm = np.roll(np.arange(ndim), 1) if ndim > 1 else np.arange(ndim)
  return x + 0.08 * np.sin(np.pi * x[:, perm] + 0.3) * (1 - 0.5 * x ** 2)


def shuffled(premesh: Premesh, seed: int) -> Premesh:
  """Random element order + per-element axis permutation / flips."""
  rng = np.random.default_rng(seed)
  ndim = premesh.ndim
  elements = np.array(premesh.elements)[rng.permutation(premesh.num_elements)]
  out = []
  for el in elements:
    nd = el.reshape([2] * ndim).transpose(rng.permutation(ndim))
    flips = [ax for ax in range(ndim) if rng.integers(2)]
    nd = np.flip(nd, flips) if flips else nd
    out.append(nd.reshape(-1))
  return Premesh.create(node_coords=premesh.node_coords,
                        elements=np.array(out, dtype=np.int32),
                        physical_groups=premesh.physical_groups,
                        periodic_links=premesh.periodic_links)


def rotated_quads(premesh: Premesh, seed: int) -> Premesh:
  """Random element order + per-element 90-degree rotations of the vertex
  listing.  Unlike `shuffled`, every element keeps a POSITIVE Jacobian
  determinant (the reference integrates with the signed determinant,
  fespace.py:346, 402, so reflected elements make the operator indefinite)."""
  assert premesh.ndim == 2
  rng = np.random.default_rng(seed)
  elements = np.array(premesh.elements)[rng.permutation(premesh.num_elements)]
  out = [np.rot90(el.reshape(2, 2), k=int(rng.integers(4))).reshape(-1)
         for el in elements]
  return Premesh.create(node_coords=premesh.node_coords,
                        elements=np.array(out, dtype=np.int32),
                        physical_groups=premesh.physical_groups,
                        periodic_links=premesh.periodic_links)


def deformed_premesh(ndim, ne, n1d, seed=None, periodic_dims=(), curved=True,
                     rotate_seed=None):
  """Refined GLL premesh on [-1,1]^ndim with deformed (curved) elements."""
  pm = unit_cube_mesh(ne, ndim=ndim, a=-1., b=1., periodic_dims=periodic_dims)
  if seed is not None:
    pm = shuffled(pm, seed)
  if rotate_seed is not None:
    pm = rotated_quads(pm, rotate_seed)
  refined = refine_premesh(pm, Nodes1D.create(n1d, GLL))
  if curved and not periodic_dims:
    refined = refined.replace(node_coords=deform(refined.node_coords))
  return refined


def local_to_global(global_coords, local_coords):
  """Index of every local node in the global mesh, matched through the
  (undeformed) coordinates."""
  ndim = global_coords.shape[1]
  key = lambda c: np.round(np.asarray(c) * 1e9).astype(np.int64)  # noqa: E731
  gk, lk = key(global_coords), key(local_coords)
  order = np.lexsort(gk.T[::-1])
  rec = [('', np.int64)] * ndim
  view_g = np.ascontiguousarray(gk[order]).view(rec).ravel()
  view_l = np.ascontiguousarray(lk).view(rec).ravel()
  l2g = order[np.searchsorted(view_g, view_l)]
  assert np.array_equal(gk[l2g], lk)
  return l2g
