"""Structured premeshes.

`unit_cube_mesh` has the signature and output of the reference's
`swirl_fem/common/premesh_commons.py:67-145` (first-order line/quad/hex mesh
of `[a, b]^ndim`, 'boundary' physical group, periodic links, block
partitions), built with array arithmetic instead of nested Python products so
that the 68^3-element config-4 mesh stages in well under a second.
"""

from __future__ import annotations

from collections.abc import Sequence

import numpy as np

from swirl_fem_b200.common import facet_util
from swirl_fem_b200.common.facet_util import FacetDimType
from swirl_fem_b200.core.premesh import Premesh


def _facet_elements(num_elements, facet) -> np.ndarray:
  """Node ids of the first-order (sub)elements lying on `facet`.

  `num_elements` is the per-axis element count (tuple).  Rows follow the
  `itertools.product` order of the per-axis 1-D elements, columns the
  lexicographic order of the 2^k corner offsets.
  """
  ndim = len(facet)
  n1 = [ne + 1 for ne in num_elements]
  starts, corners = [], []
  for axis, t in enumerate(facet):
    if t == FacetDimType.INNER:
      starts.append(np.arange(num_elements[axis]))
      corners.append(np.array([0, 1]))
    elif t == FacetDimType.FIRST:
      starts.append(np.array([0]))
      corners.append(np.array([0]))
    else:
      starts.append(np.array([num_elements[axis]]))
      corners.append(np.array([0]))
  strides = [int(np.prod(n1[axis + 1:])) for axis in range(ndim)]
  sgrid = np.meshgrid(*starts, indexing='ij') if ndim else []
  cgrid = np.meshgrid(*corners, indexing='ij') if ndim else []
  base = sum(s.reshape(-1) * st for s, st in zip(sgrid, strides))
  off = sum(c.reshape(-1) * st for c, st in zip(cgrid, strides))
  return (np.asarray(base)[:, None] + np.asarray(off)[None, :]).astype(
      np.int64)


def box_mesh(num_elements: Sequence[int], lo: Sequence[float],
             hi: Sequence[float], periodic_dims: Sequence[int] = (),
             partitions: np.ndarray | None = None) -> Premesh:
  """Uniform first-order mesh of the box `prod_i [lo_i, hi_i]`.

  Generalises the reference's `unit_cube_mesh` to per-axis element counts and
  bounds (used to build one rank's block of an element-partitioned cube
  directly, without staging the global mesh on every rank).
  """
  num_elements = tuple(int(n) for n in num_elements)
  ndim = len(num_elements)
  axes = [np.linspace(lo[i], hi[i], num=num_elements[i] + 1)
          for i in range(ndim)]
  node_coords = np.stack(np.meshgrid(*axes, indexing='ij'), axis=-1).reshape(
      -1, ndim)

  elements = _facet_elements(num_elements, (FacetDimType.INNER,) * ndim)

  axis_to_facets = {axis: [] for axis in range(ndim)}
  for facet in facet_util.get_facet_types(ndim, facet_ndim=ndim - 1):
    axis = (facet.index(FacetDimType.FIRST) if FacetDimType.FIRST in facet
            else facet.index(FacetDimType.LAST))
    axis_to_facets[axis].append(_facet_elements(num_elements, facet))

  boundary, links = [], []
  for axis in range(ndim):
    if axis in periodic_dims:
      links.append(np.stack(axis_to_facets[axis], axis=1))
    else:
      boundary.extend(axis_to_facets[axis])

  physical_groups = {}
  if boundary:
    physical_groups['boundary'] = np.concatenate(boundary).astype(np.int32)
  periodic_links = (np.concatenate(links).astype(np.int32) if links else None)

  if partitions is not None:
    partitions = np.asarray(partitions)
    for axis in range(ndim):
      assert num_elements[axis] % partitions.shape[axis] == 0, (
          partitions.shape)
      partitions = np.repeat(
          partitions, num_elements[axis] // partitions.shape[axis], axis=axis)
    partitions = partitions.reshape(len(elements))

  return Premesh.create(
      node_coords=node_coords,
      elements=elements.astype(np.int32),
      periodic_links=periodic_links,
      physical_groups=physical_groups,
      partitions=partitions)


def unit_cube_mesh(
    num_elements_per_dim: int,
    ndim: int = 2,
    a: float = 0.0,
    b: float = 1.0,
    periodic_dims: Sequence[int] = (),
    partitions: np.ndarray | None = None,
) -> Premesh:
  """Uniform first-order mesh of the cube `[a, b]^ndim`."""
  return box_mesh((num_elements_per_dim,) * ndim, (a,) * ndim, (b,) * ndim,
                  periodic_dims=periodic_dims, partitions=partitions)
