#!/usr/bin/env python
"""Timing of the Stokes / Navier-Stokes operators (BASELINE config 3 shape:
2-D, order 7, y-periodic channel; `--ne` elements per axis).

Developer tool (not the driver's bench): CUDA-event times of every operator
of `swirl_fem_b200.navier_stokes.StokesSEM` and of one `stokes_one_step`, with
the bytes each operator must move (x in, y out, connectivity, geometric
factors) so the numbers can be set against the measured HBM peak.

  python tools/bench_ns.py --ne 64 --order 7
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
  ap.add_argument('--ne', type=int, default=64)
  ap.add_argument('--order', type=int, default=7)
  ap.add_argument('--reps', type=int, default=20)
  ap.add_argument('--dtype', default='f64', choices=['f64', 'f32'])
  ap.add_argument('--ops-only', action='store_true',
                  help='time the operators only (profiling runs)')
  args = ap.parse_args()

  import torch
  from swirl_fem_b200.common.premesh_commons import unit_cube_mesh
  from swirl_fem_b200.navier_stokes import navier_stokes as ns

  dtype = torch.float64 if args.dtype == 'f64' else torch.float32
  esz = 8 if args.dtype == 'f64' else 4
  pm = unit_cube_mesh(args.ne, ndim=2, periodic_dims=(1,))
  x = np.asarray(pm.node_coords, dtype=np.float64)
  x = np.stack([2 * x[:, 0] - 1, 2 * np.pi * x[:, 1] - np.pi], -1)
  x[:, 0] += 0.1 * np.sin(x[:, 1]) * (1 - x[:, 0] ** 2)
  pm = pm.replace(node_coords=x)
  sem = ns.StokesSEM.create(
      pm, boundary_conditions={'boundary': (ns.BCType.DIRICHLET, 0.0)},
      order=args.order, dtype=dtype)
  vm, pmesh = sem.velocity.mesh, sem.pressure.pspace.mesh
  gen = torch.Generator(device='cuda').manual_seed(0)
  u = torch.randn(vm.num_nodes, 2, dtype=dtype, device='cuda', generator=gen)
  u = u * sem.velocity.interior_mask
  p = torch.randn(pmesh.num_nodes, dtype=dtype, device='cuda', generator=gen)
  dt, k = 1e-3, 3

  def timed(fn):
    for _ in range(3):
      fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(
        enable_timing=True)
    a.record()
    for _ in range(args.reps):
      fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / args.reps * 1e3

  e, nv = vm.num_elements, vm.num_nodes_per_element
  np_ = pmesh.num_nodes_per_element
  q = nv  # collocated GLL rule shared by both spaces
  # minimal traffic of the fused forms (bytes): fields once, connectivity,
  # the geometric data each evaluation needs
  traffic = {
      'A': 2 * 2 * vm.num_nodes * esz + e * nv * (4 + 3 * esz),
      'D': 2 * vm.num_nodes * esz + pmesh.num_nodes * esz + e * (
          nv * 4 + np_ * 4 + q * 5 * esz),
      'Dt': 2 * vm.num_nodes * esz + pmesh.num_nodes * esz + e * (
          nv * 4 + np_ * 4 + q * 5 * esz),
  }
  ops = {
      'A': lambda: sem.A(u), 'B': lambda: sem.B(u), 'Bi': lambda: sem.Bi(u),
      'C': lambda: sem.C(u), 'D': lambda: sem.D(u), 'Dt': lambda: sem.Dt(p),
      'E': lambda: sem.E(p, dt=dt, time_order=k),
      'filter': lambda: sem.filter(u),
      'exchange': lambda: sem.velocity.exchange(u),
  }
  out = {'ne': args.ne, 'order': args.order, 'dtype': args.dtype,
         'velocity_dofs': int(vm.num_nodes), 'pressure_dofs':
         int(pmesh.num_nodes), 'us': {}, 'gbs': {}}
  for name, fn in ops.items():
    t = timed(fn)
    out['us'][name] = t
    if name in traffic:
      out['gbs'][name] = traffic[name] / (t * 1e-6) / 1e9
  if args.ops_only:
    print(json.dumps(out))
    return
  us_hist = [u * (1.0 - 0.01 * i) for i in range(k)]
  ps_hist = [p * 0.0 for _ in range(k)]
  torch.cuda.synchronize()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(
      enable_timing=True)
  a.record()
  _, _, aux = sem.stokes_one_step(us_hist, ps_hist, f=0, mu=1e-2, dt=dt,
                                  time_order=k, tol=1e-5, atol=1e-4)
  b.record()
  torch.cuda.synchronize()
  out['stokes_one_step_ms'] = a.elapsed_time(b)
  out['u_star_iterations'] = aux['u_star_info']['num_iterations']
  out['dp_iterations'] = aux['dp_info']['num_iterations']
  # the same step again (kernels loaded, caches filled), and the pressure CG
  # alone for a fixed number of iterations with and without the CUDA graph
  import time
  from functools import partial
  from swirl_fem_b200.linalg import cg as cgmod
  a.record()
  t0 = time.perf_counter()
  _, _, aux = sem.stokes_one_step(us_hist, ps_hist, f=0, mu=1e-2, dt=dt,
                                  time_order=k, tol=1e-5, atol=1e-4)
  b.record()
  torch.cuda.synchronize()
  out['stokes_one_step_second_call_ms'] = a.elapsed_time(b)
  out['stokes_one_step_second_call_wall_ms'] = (time.perf_counter() - t0) * 1e3
  out['pressure_cg_last_run'] = dict(cgmod.LAST_DEVICE_STATE_RUN)
  rhs = -sem.D(u)
  precond = partial(ns._pressure_project_out_nullspace, sem)  # pylint: disable=protected-access
  rhs = precond(rhs)
  for use_graph in (True, False):
    torch.cuda.synchronize()
    a.record()
    _, info = cgmod.cg(partial(sem.E, dt=dt, time_order=k), rhs, M=precond,
                       tol=0.0, atol=0.0, maxiter=400, graph=use_graph)
    b.record()
    torch.cuda.synchronize()
    key = 'graph' if use_graph else 'eager'
    out[f'pressure_cg_400_iterations_{key}_ms'] = a.elapsed_time(b)
    out[f'pressure_cg_400_iterations_{key}_run'] = dict(
        cgmod.LAST_DEVICE_STATE_RUN)
  print(json.dumps(out))


if __name__ == '__main__':
  main()
