"""Host logic of the lazy zero fill (`core.operator.lazy_zero_tables`): the
tables handed to `sfem_op_set_lazy_zero` must cover y's shared-dof prefix
exactly once, every piece must be zeroed by a duty step of the chunk
`lookahead` chunks before the dof's first touch (or before the launch), and no
piece may exceed the per-thread size.  Pure index arithmetic: runs on the CPU."""

import numpy as np
import pytest
import torch

from swirl_fem_b200.communication import partition as part
from swirl_fem_b200.core.interpolation import Nodes1D, NodeType
from swirl_fem_b200.core.operator import lazy_zero_tables
from tests import helpers

GLL = NodeType.GAUSS_LOBATTO_LEGENDRE


def _check(elements, num_nodes, epb, chunk, kd, lookahead, piece=128):
  E, n = elements.shape
  counts = np.bincount(elements.reshape(-1), minlength=num_nodes)
  nz = int(np.nonzero(counts != 1)[0].max()) + 1
  out = lazy_zero_tables(torch.as_tensor(elements), num_nodes, nz, epb, chunk,
                         lookahead, kd, piece)
  assert out is not None
  pieces, num_eager, duty_ptr, S = out
  pieces, dp = pieces.numpy(), duty_ptr.numpy()
  num_steps = -(-E // epb)
  num_duty = -(-num_steps // kd)
  assert len(dp) == num_duty + 1 and S % kd == 0
  assert dp[0] == num_eager and dp[-1] == len(pieces)
  assert (np.diff(dp) >= 0).all()
  assert (pieces[:, 1] > 0).all() and (pieces[:, 1] <= piece).all()
  cover = np.zeros(nz, dtype=int)
  zero_chunk = np.full(nz, -1)
  for a, l in pieces[:num_eager]:
    cover[a:a + l] += 1
  for q in range(num_duty):
    for a, l in pieces[dp[q]:dp[q + 1]]:
      cover[a:a + l] += 1
      zero_chunk[a:a + l] = (q * kd) // S
  assert (cover == 1).all()                       # every dof exactly once
  first = np.full(nz, 10 ** 9)
  flat = elements.reshape(-1)
  chunk_of = np.repeat(np.arange(E) // (S * epb), n)
  m = (flat >= 0) & (flat < nz)
  np.minimum.at(first, flat[m], chunk_of[m])
  lazy = zero_chunk >= 0
  assert (zero_chunk[lazy] == first[lazy] - lookahead).all()
  assert ((first[~lazy] < lookahead) | (first[~lazy] == 10 ** 9)).all()
  return len(pieces), num_eager


@pytest.mark.parametrize('case', [(8, 8, 1, 64, 8, 2), (10, 5, 3, 60, 2, 2),
                                  (10, 6, 2, 64, 4, 1), (12, 4, 4, 96, 3, 3)])
def test_tables_structured_blocks(case):
  ne, n1d, epb, chunk, kd, lookahead = case
  blk = part.block_partition(ne, 3, Nodes1D.create(n1d, GLL), 0, 1)
  _check(blk.premesh.elements, blk.premesh.num_nodes, epb, chunk, kd,
         lookahead)


def test_tables_shuffled_elements_and_small_pieces():
  """Any numbering works: shuffled element order fragments the id runs."""
  refined = helpers.deformed_premesh(3, 6, 4, seed=5, reorient=False)
  npieces, _ = _check(np.asarray(refined.elements), refined.num_nodes, 2, 16,
                      2, 1, piece=32)
  assert npieces > 100


def test_tables_refuse_tiny_meshes():
  blk = part.block_partition(2, 3, Nodes1D.create(4, GLL), 0, 1)
  el = blk.premesh.elements
  assert lazy_zero_tables(torch.as_tensor(el), blk.premesh.num_nodes, 100, 1,
                          512, 2, 8, 128) is None
