"""Host logic of the lazy zero fill (`core.operator.lazy_zero_tables`): the
tables handed to `sfem_op_set_lazy_zero` must cover y's shared-dof prefix
exactly once, the pieces must be sorted by the chunk (= round of CTA steps)
that touches their dofs first -- the companion kernel zeroes them in that
order and a step waits for its own chunk only --, dofs no element touches go
with chunk 0, and no piece may exceed the per-warp size.  Pure index
arithmetic: runs on the CPU."""

import numpy as np
import pytest
import torch

from swirl_fem_b200.communication import partition as part
from swirl_fem_b200.core.interpolation import Nodes1D, NodeType
from swirl_fem_b200.core.operator import lazy_zero_tables
from tests import helpers

GLL = NodeType.GAUSS_LOBATTO_LEGENDRE


def _check(elements, num_nodes, step_elems, grid, piece=2048):
  E, n = elements.shape
  counts = np.bincount(elements.reshape(-1), minlength=num_nodes)
  nz = int(np.nonzero(counts != 1)[0].max()) + 1
  out = lazy_zero_tables(torch.as_tensor(elements), num_nodes, nz, step_elems,
                         grid, piece)
  assert out is not None
  pieces, chunk_ptr = out
  pieces, cp = pieces.numpy(), chunk_ptr.numpy()
  chunk_elems = grid * step_elems
  num_chunks = -(-E // chunk_elems)
  assert len(cp) == num_chunks + 1 and cp[0] == 0 and cp[-1] == len(pieces)
  assert (np.diff(cp) >= 0).all()
  start, length = pieces[:, 0], pieces[:, 1] & 0xfff
  assert (length > 0).all() and (length <= piece).all()
  cover = np.zeros(nz, dtype=int)
  owner = np.full(nz, -2)
  pchunk = np.searchsorted(cp, np.arange(len(pieces)), side='right') - 1
  assert (pieces[:, 1] >> 12 == pchunk).all()     # the label the kernel uses
  for i, (a, l) in enumerate(zip(start, length)):
    cover[a:a + l] += 1
    owner[a:a + l] = pchunk[i]
  assert (cover == 1).all()                       # every dof exactly once
  # the chunk whose range holds a piece is the chunk that touches every dof of
  # the piece FIRST (untouched dofs: chunk 0)
  first = np.full(nz, 10 ** 9)
  flat = elements.reshape(-1)
  chunk_of = np.repeat(np.arange(E) // chunk_elems, n)
  m = (flat >= 0) & (flat < nz)
  np.minimum.at(first, flat[m], chunk_of[m])
  first[first == 10 ** 9] = 0
  assert (owner == first).all()
  return len(pieces)


@pytest.mark.parametrize('case', [(8, 8, 1, 37), (10, 5, 3, 20),
                                  (10, 6, 2, 64), (12, 4, 4, 50)])
def test_tables_structured_blocks(case):
  ne, n1d, step_elems, grid = case
  blk = part.block_partition(ne, 3, Nodes1D.create(n1d, GLL), 0, 1)
  _check(blk.premesh.elements, blk.premesh.num_nodes, step_elems, grid)


def test_tables_shuffled_elements_and_small_pieces():
  """Any numbering works: shuffled element order fragments the id runs."""
  refined = helpers.deformed_premesh(3, 6, 4, seed=5, reorient=False)
  npieces = _check(np.asarray(refined.elements), refined.num_nodes, 2, 8,
                   piece=32)
  assert npieces > 100


def test_tables_refuse_tiny_meshes():
  blk = part.block_partition(2, 3, Nodes1D.create(4, GLL), 0, 1)
  el = blk.premesh.elements
  assert lazy_zero_tables(torch.as_tensor(el), blk.premesh.num_nodes, 100, 1,
                          512) is None
