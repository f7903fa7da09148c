"""Host logic of the lazy zero fill (`core.operator.lazy_zero_tables`): the
tables handed to `sfem_op_set_lazy_zero` must cover y's shared-dof prefix
exactly once, the pieces must be sorted by -- and labelled with -- the chunk
that touches their dofs first (the kernel zeroes them in that order and a step
waits for its own chunk's pieces), dofs no element touches must be zeroed
before the launch, and no piece may exceed the per-warp size.  Pure index arithmetic: runs on the CPU."""

import numpy as np
import pytest
import torch

from swirl_fem_b200.communication import partition as part
from swirl_fem_b200.core.interpolation import Nodes1D, NodeType
from swirl_fem_b200.core.operator import lazy_zero_tables
from tests import helpers

GLL = NodeType.GAUSS_LOBATTO_LEGENDRE


def _check(elements, num_nodes, epb, chunk, lookahead, piece=128):
  E, n = elements.shape
  counts = np.bincount(elements.reshape(-1), minlength=num_nodes)
  nz = int(np.nonzero(counts != 1)[0].max()) + 1
  out = lazy_zero_tables(torch.as_tensor(elements), num_nodes, nz, epb, chunk,
                         lookahead, piece)
  assert out is not None
  pieces, num_eager, chunk_ptr, S = out
  pieces, cp = pieces.numpy(), chunk_ptr.numpy()
  num_steps = -(-E // epb)
  num_chunks = -(-num_steps // S)
  assert len(cp) == num_chunks + 1 and cp[-1] == len(pieces)
  assert (np.diff(cp) >= 0).all() and num_eager == cp[lookahead]
  start, length, pchunk = pieces[:, 0], pieces[:, 1] & 0xff, pieces[:, 1] >> 8
  assert (length > 0).all() and (length <= piece).all()
  cover = np.zeros(nz, dtype=int)
  owner = np.full(nz, -2)
  for i, (a, l) in enumerate(zip(start, length)):
    cover[a:a + l] += 1
    owner[a:a + l] = pchunk[i] if i >= cp[0] else -1
  assert (cover == 1).all()                       # every dof exactly once
  # the chunk recorded in a piece is the chunk whose range holds the piece,
  # and it is the chunk that touches every dof of the piece FIRST
  for c in range(num_chunks):
    assert (pchunk[cp[c]:cp[c + 1]] == c).all()
  first = np.full(nz, 10 ** 9)
  flat = elements.reshape(-1)
  chunk_of = np.repeat(np.arange(E) // (S * epb), n)
  m = (flat >= 0) & (flat < nz)
  np.minimum.at(first, flat[m], chunk_of[m])
  touched = first < 10 ** 9
  assert (owner[touched] == first[touched]).all()
  assert (owner[~touched] == -1).all()            # zeroed before the launch
  return len(pieces), num_eager


@pytest.mark.parametrize('case', [(8, 8, 1, 64, 2), (10, 5, 3, 60, 2),
                                  (10, 6, 2, 64, 1), (12, 4, 4, 96, 3)])
def test_tables_structured_blocks(case):
  ne, n1d, epb, chunk, lookahead = case
  blk = part.block_partition(ne, 3, Nodes1D.create(n1d, GLL), 0, 1)
  _check(blk.premesh.elements, blk.premesh.num_nodes, epb, chunk, lookahead)


def test_tables_shuffled_elements_and_small_pieces():
  """Any numbering works: shuffled element order fragments the id runs."""
  refined = helpers.deformed_premesh(3, 6, 4, seed=5, reorient=False)
  npieces, _ = _check(np.asarray(refined.elements), refined.num_nodes, 2, 16,
                      1, piece=32)
  assert npieces > 100


def test_tables_refuse_tiny_meshes():
  blk = part.block_partition(2, 3, Nodes1D.create(4, GLL), 0, 1)
  el = blk.premesh.elements
  assert lazy_zero_tables(torch.as_tensor(el), blk.premesh.num_nodes, 100, 1,
                          512, 2, 128) is None
