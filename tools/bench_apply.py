#!/usr/bin/env python
"""Kernel-level sweep: times `FusedOperator.apply` per variant / order / dtype.

Developer tool (not the driver's bench): prints one line per configuration
with the CUDA-event time of the apply (memset + kernel), GDOF/s and the
fraction of the measured HBM roofline, and checks every variant against the
generic kernel (variant 1) on the same mesh.

  python tools/bench_apply.py --dim 3 --orders 7 --ne 24 --variants 0,2

Variants (`sfem_op_set_variant`): 0 default, 1 generic kernel, 2 v1 kernels;
with a library built with `make EXTRA=-DSFEM_EXPERIMENTS`: 3-D Laplacian, any
order and precision: 9 connectivity fetched two steps ahead, 10 factors staged
with an L2 evict-first policy, 11 = 9 + 10; 3-D fp64 Laplacian, N = 5..9 only:
3/4/5 elements per CTA +1/-1/x2, 6/7 resident CTAs +1/-1, 8 factors streamed
instead of staged.
"""

from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--dim', type=int, default=3)
  ap.add_argument('--orders', default='7')
  ap.add_argument('--ne', type=int, default=0, help='0: pick ~target dofs')
  ap.add_argument('--target-dofs', type=float, default=5e6)
  ap.add_argument('--variants', default='0')
  ap.add_argument('--dtypes', default='f64')
  ap.add_argument('--reps', type=int, default=20)
  ap.add_argument('--mass', type=int, default=0)
  ap.add_argument('--check', type=int, default=1)
  args = ap.parse_args()

  import torch
  from swirl_fem_b200.common.premesh_commons import unit_cube_mesh
  from swirl_fem_b200.core.interpolation import Nodes1D, NodeType, Quadrature1D
  from swirl_fem_b200.core.mesh import Mesh
  from swirl_fem_b200.core.mesh_refiner import refine_premesh
  from swirl_fem_b200.core.operator import FusedOperator

  try:
    peak = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs']
  except OSError:
    peak = 6650.0
  gll = NodeType.GAUSS_LOBATTO_LEGENDRE
  for order in [int(o) for o in args.orders.split(',')]:
    n1d = order + 1
    ne = args.ne or max(1, int(round(
        (args.target_dofs ** (1.0 / args.dim) - 1) / order)))
    refined = refine_premesh(unit_cube_mesh(ne, ndim=args.dim, a=-1., b=1.),
                             Nodes1D.create(n1d, gll))
    x = refined.node_coords
    perm = np.roll(np.arange(args.dim), 1)
    coords = x + 0.08 * np.sin(np.pi * x[:, perm]) * (1 - x ** 2)
    bmask = refined.finalize_host()['physical_masks']['boundary']
    for dt in args.dtypes.split(','):
      dtype = torch.float64 if dt == 'f64' else torch.float32
      esz = 8 if dt == 'f64' else 4
      mesh = Mesh.create(coords, refined.elements,
                         gridpoints_1d=Nodes1D.create(n1d, gll), dtype=dtype)
      quad = Quadrature1D.create_from_nodes_1d(Nodes1D.create(n1d, gll))
      op = FusedOperator(mesh, quad, dirichlet_mask=bmask,
                         with_mass=bool(args.mass))
      g = torch.Generator(device='cuda').manual_seed(0)
      u = torch.randn(mesh.num_nodes, dtype=dtype, device='cuda', generator=g)
      y = torch.empty_like(u)
      ref = None
      if args.check:
        try:
          op.set_variant(1)
          ref = op.apply(u).double()
        except NotImplementedError:
          ref = None
      nloc = mesh.num_elements * mesh.num_nodes_per_element
      geo = args.dim * (args.dim + 1) // 2 + args.mass
      abytes = 2 * esz * mesh.num_nodes + nloc * (geo * esz + 4)
      for variant in [int(v) for v in args.variants.split(',')]:
        try:
          op.set_variant(variant)
          lam = 0.5 if args.mass else 0.0
          for _ in range(3):
            op.apply(u, lam=lam, out=y)
          torch.cuda.synchronize()
          a = torch.cuda.Event(enable_timing=True)
          b = torch.cuda.Event(enable_timing=True)
          a.record()
          for _ in range(args.reps):
            op.apply(u, lam=lam, out=y)
          b.record()
          torch.cuda.synchronize()
          ms = a.elapsed_time(b) / args.reps
          err = float('nan')
          if ref is not None and not args.mass:
            err = float((y.double() - ref).abs().max() / ref.abs().max())
          print(f'dim={args.dim} p={order} {dt} ne={ne} dofs={mesh.num_nodes} '
                f'variant={variant}: {ms * 1e3:9.1f} us  '
                f'{mesh.num_nodes / ms / 1e6:7.2f} GDOF/s  '
                f'{abytes / ms / 1e6:7.0f} GB/s  '
                f'{abytes / ms / 1e6 / peak * 100:5.1f}% of {peak:.0f}  '
                f'err_vs_generic={err:.1e}', flush=True)
        except (NotImplementedError, ValueError) as exc:
          print(f'dim={args.dim} p={order} {dt} variant={variant}: '
                f'unsupported ({exc})', flush=True)
      del op, mesh


if __name__ == '__main__':
  main()
