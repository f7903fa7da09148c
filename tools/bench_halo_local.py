#!/usr/bin/env python
"""Single-GPU timing of the halo-fused apply kernel (all ranks' blocks of a
2-rank partition live on cuda:0, peers' regions are local memory): isolates
the in-kernel cost of signalling / pushing from NVLink effects.

  python tools/bench_halo_local.py [ne] [order]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from swirl_fem_b200.communication import partition as part  # noqa: E402
from swirl_fem_b200.communication.halo import HaloPlan  # noqa: E402
from swirl_fem_b200.core.interpolation import Nodes1D, NodeType, Quadrature1D  # noqa: E402
from swirl_fem_b200.core.mesh import Mesh  # noqa: E402
from swirl_fem_b200.core.operator import FusedOperator  # noqa: E402


def timed(fn, reps=20, warm=5):
  for _ in range(warm):
    fn()
  torch.cuda.synchronize()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(
      enable_timing=True)
  a.record()
  for _ in range(reps):
    fn()
  b.record()
  torch.cuda.synchronize()
  return a.elapsed_time(b) / reps * 1e3


def main():
  ne = int(sys.argv[1]) if len(sys.argv) > 1 else 48
  order = int(sys.argv[2]) if len(sys.argv) > 2 else 7
  world = int(sys.argv[3]) if len(sys.argv) > 3 else 2
  dev = torch.device('cuda', 0)
  grid1d = Nodes1D.create(order + 1, NodeType.GAUSS_LOBATTO_LEGENDRE)
  quad = Quadrature1D.create_from_nodes_1d(grid1d)
  blks = [part.block_partition(ne, 3, grid1d, r, world) for r in range(world)]
  gathered = [np.sort(b.interface_global) for b in blks]
  plans = [part.halo_plan_from_interfaces(
      r, b.interface_local, b.interface_global, gathered, b.premesh.num_nodes)
           for r, b in enumerate(blks)]
  HaloPlan.enable_p2p_local(plans, torch.float64, dev)
  ops, xs, ys = [], [], []
  for b in blks[:2]:
    x0 = b.premesh.node_coords
    perm = np.roll(np.arange(3), 1)
    coords = x0 + 0.08 * np.sin(np.pi * x0[:, perm]) * (1 - x0 ** 2)
    mesh = Mesh.create(coords, b.premesh.elements, gridpoints_1d=grid1d,
                       device=dev)
    ops.append(FusedOperator(mesh, quad, dirichlet_mask=b.dirichlet,
                             with_mass=False))
    xs.append(torch.randn(mesh.num_nodes, dtype=torch.float64, device=dev))
    ys.append(torch.empty_like(xs[-1]))
  n = xs[0].numel()
  print(f'ne={ne} order={order} world={world}: rank-0 block {n} dofs, '
        f'{blks[0].num_interface_elements} interface elements of '
        f'{blks[0].premesh.num_elements}, {plans[0].splits()} shared dofs')
  t_plain = timed(lambda: ops[0].apply(xs[0], out=ys[0]))
  # fused: rank 0 pushes into rank 1's (local) region; no wait (rank 1 never
  # pushes), counters are reset by hand through a wait on rank 1's flags...
  def fused():
    for r in range(2):
      ops[r].apply_partitioned(xs[r], ys[r], plans[r],
                               blks[r].num_interface_elements, wait=False)
    for r in range(2):
      plans[r].p2p_wait_unpack(ys[r])
  def unfused():
    for r in range(2):
      ops[r].apply(xs[r], out=ys[r])
    for r in range(2):
      plans[r].p2p_push(ys[r])
    for r in range(2):
      plans[r].p2p_wait_unpack(ys[r])
  def plain2():
    for r in range(2):
      ops[r].apply(xs[r], out=ys[r])
  if world == 2:
    import itertools
    res = {}
    for rnd, mode in itertools.product(range(2), ('plain', 'fused',
                                                  'fused_nounpack', 'unfused')):
      for pl in plans:
        pl.p2p_set_option(1, 0 if mode == 'fused_nounpack' else 1)
      fn = {'plain': plain2, 'fused': fused, 'fused_nounpack': fused,
            'unfused': unfused}[mode]
      res.setdefault(mode, []).append(timed(fn))
      if mode.startswith('fused') and rnd == 0:
        fn()
        torch.cuda.synchronize()
        for r in range(2):
          print(mode, 'rank', r, 'stamps (us):', plans[r].p2p_debug_times(dev))
    print(f'plain apply rank0 {t_plain:.1f} us; both ranks: ' + '; '.join(
        f'{k} {min(v):.1f}' for k, v in res.items()))


if __name__ == '__main__':
  main()
