"""Facets of tensor-product (line/quad/hex) elements.

Same vocabulary as the reference's `swirl_fem/common/facet_util.py`
(FacetDimType :45-50, slice_from_facet_type :53-75, get_facet_types :78-92,
get_orderings_mapping :95-143); used by the mesh refiner and the structured
premesh generator.  An n-cube has 3^n facets, each named by an n-tuple of
FIRST / LAST / INNER, iterated in `itertools.product` order of the enum
declaration order -- that iteration order *defines* the global node numbering
of refined meshes, so it is part of the bit-exact contract.
"""

from __future__ import annotations

import enum
import itertools

import numpy as np


@enum.unique
class FacetDimType(enum.Enum):
  FIRST = 'first'
  LAST = 'last'
  INNER = 'inner'


def slice_from_facet_type(facet_type, interior_nodes_only: bool):
  """Index tuple extracting a facet from an `[k+1]*n`-shaped element array."""
  inner = slice(1, -1) if interior_nodes_only else slice(None)
  table = {FacetDimType.FIRST: 0, FacetDimType.LAST: -1,
           FacetDimType.INNER: inner}
  return tuple(table[t] for t in facet_type)


def get_facet_types(ndim: int, facet_ndim: int | None = None):
  """All facet signatures of an `ndim`-cube (optionally of one dimension)."""
  facets = list(itertools.product(list(FacetDimType), repeat=ndim))
  if facet_ndim is None:
    return facets
  return [f for f in facets if f.count(FacetDimType.INNER) == facet_ndim]


def _orientations(ndim: int):
  """Yields (axis permutation, flipped axes) in the reference's order."""
  for perm in itertools.permutations(range(ndim)):
    for r in range(ndim + 1):
      for axes in itertools.combinations(range(ndim), r):
        yield perm, axes


def get_orderings_mapping(ndim: int, num_points_1d: int):
  """Maps vertex orderings of a first-order cube to high-order node orderings.

  Keys: tuples, permutations of `range(2**ndim)` reachable by axis
  permutation + flips (2^d d! of them).  Values: the matching permutation of
  the `num_points_1d ** ndim` lexicographic nodes.
  """
  source = np.arange(2 ** ndim, dtype=np.int32).reshape([2] * ndim)
  target = np.arange(num_points_1d ** ndim, dtype=np.int32).reshape(
      [num_points_1d] * ndim)
  orderings = {}
  for perm, axes in _orientations(ndim):
    key = np.flip(source.transpose(perm), axes).flatten().tolist()
    orderings[tuple(key)] = np.flip(target.transpose(perm), axes).flatten()
  return orderings


def orderings_table(ndim: int, num_points_1d: int):
  """Vectorised form of `get_orderings_mapping`.

  Returns `(codes, perms)`: `codes` sorted int64 encodings of the keys
  (`sum(key[j] * (2**ndim)**j)`) and `perms[i]` the ordering for `codes[i]`.
  """
  mapping = get_orderings_mapping(ndim, num_points_1d)
  base = 2 ** ndim
  weights = base ** np.arange(base, dtype=np.int64)
  codes = np.array([int(np.dot(np.array(k, dtype=np.int64), weights))
                    for k in mapping], dtype=np.int64)
  perms = np.stack([np.asarray(v, dtype=np.int64) for v in mapping.values()])
  order = np.argsort(codes)
  return codes[order], perms[order].reshape(len(codes), -1)
