#!/usr/bin/env python
"""Benchmark of the B200 hot path: matrix-free Laplacian apply (+ fused CG).

Contract: `python bench.py --gpus N --steps K --warmup W` prints ONE JSON line.
A "step" is one operator apply  y = mask . Z^T K Z x  over the whole mesh
(N > 1: every rank applies its block, then the shared-dof halo exchange).
Workload (`config.workload`): BASELINE.json config 4 -- 3-D hex Poisson,
GLL order 7, ne = 68 elements per axis, 108.5 M global dofs, fp64 -- the
configuration the metric (GDOF/s apply & CG at 1/2/4/8 B200) is quoted on; it
fits one B200 (geometric factors 7.7 GB).  Total work is fixed as N grows:
`scaling: strong`.

  value      whole-job GDOF/s with inputs resident in HBM (CUDA events, max
             over ranks)
  e2e        same metric through the public API with pinned HOST buffers:
             H2D of x, apply, D2H of y inside the timed region
  roofline   algorithmic bytes of one apply / its measured duration vs the
             measured HBM peak (MEASURED_PEAKS.json)
  cpu_baseline  the numpy oracle (restated dense reference algorithm, batched
             GEMM so BLAS threads are used) on a bounded sample, rank 0, N = 1
  cg         fused PCG: ms / iteration over a fixed number of iterations

`--impl reference` times the reference's algorithm on the host cores (the
oracle port: JAX is not installable in this image).
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

METRIC = 'GDOF/s matrix-free Laplacian apply (3-D hex, GLL order 7, fp64)'


def parse_args():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=20)
  ap.add_argument('--warmup', type=int, default=5)
  ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
  ap.add_argument('--ne', type=int, default=int(os.environ.get('SFEM_NE', 68)))
  ap.add_argument('--order', type=int, default=7)
  ap.add_argument('--dim', type=int, default=3)
  ap.add_argument('--dtype', default='f64', choices=['f64', 'f32'])
  ap.add_argument('--cg-iters', type=int, default=50)
  ap.add_argument('--no-cpu-baseline', action='store_true')
  ap.add_argument('--no-e2e', action='store_true')
  ap.add_argument('--no-parity', action='store_true')
  ap.add_argument('--no-extra', action='store_true')
  return ap.parse_args()


def deform(x):
  """Smooth non-affine map so every geometric factor is populated."""
  ndim = x.shape[-1]
  perm = np.roll(np.arange(ndim), 1)
  return x + 0.08 * np.sin(np.pi * x[:, perm]) * (1 - x ** 2)


def algorithmic_bytes(num_global, num_local_nodes, ndim, esz, with_mass=False):
  """B_op of BASELINE.md section 3 for a concrete mesh: read x and write y once
  (2 s per dof), stream connectivity (4 B) and symmetric factors per local
  node."""
  g = ndim * (ndim + 1) // 2 + (1 if with_mass else 0)
  return 2 * esz * num_global + num_local_nodes * (g * esz + 4)


class ClockSampler:
  """Samples nvidia-smi clocks / throttle reasons during the timed region."""

  QUERY = ('index,clocks.sm,clocks.max.sm,power.draw,'
           'clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
           'clocks_event_reasons.hw_thermal_slowdown,'
           'clocks_event_reasons.sw_thermal_slowdown,'
           'clocks_event_reasons.sw_power_cap')

  def __init__(self, gpu_index: int):
    self.gpu_index = gpu_index
    self.proc = None
    self.lines = []

  def start(self):
    try:
      self.proc = subprocess.Popen(
          ['nvidia-smi', f'--id={self.gpu_index}',
           f'--query-gpu={self.QUERY}', '--format=csv,noheader,nounits',
           '-lms', '100'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
          text=True)
      self.thread = threading.Thread(target=self._read, daemon=True)
      self.thread.start()
    except OSError:
      self.proc = None

  def _read(self):
    for line in self.proc.stdout:
      self.lines.append(line.strip())

  def stop(self):
    if self.proc is None:
      return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['unavailable']}
    time.sleep(0.15)
    self.proc.terminate()
    try:
      self.proc.wait(timeout=2)
    except subprocess.TimeoutExpired:
      self.proc.kill()
    sm, smax, reasons = [], [], set()
    names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown',
             'sw_power_cap']
    for line in self.lines:
      parts = [p.strip() for p in line.split(',')]
      if len(parts) < 9:
        continue
      try:
        sm.append(float(parts[1]))
        smax.append(float(parts[2]))
      except ValueError:
        continue
      for name, val in zip(names, parts[5:9]):
        if val.lower().startswith('active'):
          reasons.add(name)
    return {'sm_mhz': float(np.median(sm)) if sm else None,
            'sm_max_mhz': float(max(smax)) if smax else None,
            'samples': len(sm), 'reasons': sorted(reasons)}


def measured_peak_gbs():
  path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
  try:
    with open(path) as f:
      return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
  except (OSError, KeyError, ValueError):
    return 6650.0, 'fallback (B200_PROFILING.md)'


# ----------------------------------------------------------------------------
# CPU legs (oracle port of the reference's dense algorithm)
# ----------------------------------------------------------------------------


def cpu_oracle_throughput(ndim, order, budget_s=12.0, ne_sample=None):
  """GDOF/s of the numpy oracle (dense Kronecker algorithm of the reference,
  batched GEMM) on a bounded sample of the same workload: a mesh of the same
  shape that is far larger than the host caches (3-D: ne = 16, 1.44 M dofs,
  0.4 GB of geometric data), all host cores (element chunks on a thread pool,
  BLAS pinned to one thread per chunk), median over the per-apply times."""
  from oracle import dense
  from swirl_fem_b200.common.premesh_commons import unit_cube_mesh
  from swirl_fem_b200.core.interpolation import Nodes1D, NodeType
  from swirl_fem_b200.core.mesh_refiner import refine_premesh
  try:
    from threadpoolctl import threadpool_limits
  except ImportError:  # BLAS threads then oversubscribe the chunks
    import contextlib
    threadpool_limits = lambda limits: contextlib.nullcontext()  # noqa: E731
  n1d = order + 1
  if ne_sample is None:
    ne_sample = 16 if ndim == 3 else 128
  gll = NodeType.GAUSS_LOBATTO_LEGENDRE
  refined = refine_premesh(unit_cube_mesh(ne_sample, ndim=ndim, a=-1., b=1.),
                           Nodes1D.create(n1d, gll))
  coords = deform(refined.node_coords)
  fes = dense.FESpace(coords, refined.elements, n1d, 'gauss_lobatto_legendre',
                      n1d, 'gauss_lobatto_legendre')
  u = np.random.default_rng(0).standard_normal(refined.num_nodes)
  cores = len(os.sched_getaffinity(0))
  times = []
  with threadpool_limits(limits=1):
    fes.apply_gemm(u, threads=cores)  # warm-up
    t0 = time.perf_counter()
    while True:
      t1 = time.perf_counter()
      fes.apply_gemm(u, threads=cores)
      times.append(time.perf_counter() - t1)
      if (time.perf_counter() - t0 > budget_s and len(times) >= 5) or len(
          times) >= 200:
        break
  med = float(np.median(times))
  return {
      'value': refined.num_nodes / med / 1e9,
      'unit': 'GDOF/s',
      'cores': cores,
      'kind': 'port',
      'sample': (f'{ndim}-D ne={ne_sample} order {order} '
                 f'({refined.num_nodes} dofs), median of {len(times)} applies '
                 f'({sum(times):.1f} s), numpy restatement of the reference '
                 'dense-Kronecker apply on all host cores (JAX unavailable in '
                 'image)'),
      'ms_per_step': med * 1e3,
      'spread': float((max(times) - min(times)) / med),
  }


def run_reference(args):
  rank = int(os.environ.get('RANK', 0))
  if rank != 0:
    return
  res = cpu_oracle_throughput(args.dim, args.order,
                              budget_s=max(2.0, 4.0 * args.steps / 10))
  line = {
      'impl': 'reference',
      'metric': METRIC, 'value': res['value'], 'unit': 'GDOF/s',
      'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
      'ms_per_step': res['ms_per_step'], 'higher_is_better': True,
      'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64',
      'data': 'synthetic',
      'config': bench_config(args, args.gpus),
      'note': ('reference algorithm on host cores over a bounded sample of '
               'the workload (the dense path is ~55x the flops of '
               'sum-factorisation)'),
      'cpu_baseline': {k: res[k] for k in ('value', 'unit', 'cores', 'kind',
                                           'sample')},
      'e2e': {'value': res['value'], 'unit': 'GDOF/s',
              'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
      'gpu_launches': 0,
  }
  print(json.dumps(line), flush=True)


def bench_config(args, world):
  """`config` of the JSON line: identical for both arms at the same N."""
  from swirl_fem_b200.communication.partition import GRID_FOR_WORLD
  grid = GRID_FOR_WORLD[args.dim][world]
  return {'workload': config_name(args),
          'partition': 'x'.join(str(g) for g in grid),
          'l2_policy': 'inputs larger than L2 (geometric factors alone are '
                       'tens of L2 sizes per rank)'}


def config_name(args):
  p = args.order
  return (f'{args.dim}-D hex Poisson, GLL order {p}, ne={args.ne} per axis, '
          f'{(args.ne * p + 1) ** args.dim} global dofs, deformed (non-affine) '
          'elements, homogeneous Dirichlet, collocated GLL quadrature')


# ----------------------------------------------------------------------------
# Parity of the measured path against the CPU oracle (outside every timed region)
# ----------------------------------------------------------------------------


def parity_block(rank, world, device, ndim=3, ne=4, order=7, cg_iters=20):
  """The partitioned apply and CG of THIS job -- real ranks, CUDA-IPC peer
  memory halo pushed from inside the apply kernel, scalars all-reduced inside
  the fused step kernel -- on a small mesh of the benchmark's shape, compared
  on rank 0 with `oracle.dense` on the UNPARTITIONED mesh (the parity target
  of SURVEY section 8e; reference: gather_scatter.py:246-248, cg.py:75-86).
  The oracle is the checker here, never the thing measured."""
  import torch
  import torch.distributed as dist
  from swirl_fem_b200.communication import partition as part
  from swirl_fem_b200.communication.dist_cg import distributed_cg
  from swirl_fem_b200.core.interpolation import Nodes1D, NodeType, Quadrature1D
  from swirl_fem_b200.core.mesh import Mesh
  from swirl_fem_b200.core.operator import FusedOperator

  gll = NodeType.GAUSS_LOBATTO_LEGENDRE
  grid1d = Nodes1D.create(order + 1, gll)
  quad = Quadrature1D.create_from_nodes_1d(grid1d)
  blk = part.block_partition(ne, ndim, grid1d, rank, world)
  x0 = blk.premesh.node_coords  # undeformed: identifies the global dof
  mesh = Mesh.create(deform(x0), blk.premesh.elements, gridpoints_1d=grid1d,
                     device=device, dtype=torch.float64)
  op = FusedOperator(mesh, quad, dirichlet_mask=blk.dirichlet, with_mass=True)
  halo, path = None, 'single rank'
  if world > 1:
    gathered = [None] * world
    dist.all_gather_object(gathered, np.sort(blk.interface_global))
    halo = part.halo_plan_from_interfaces(
        rank, blk.interface_local, blk.interface_global, gathered,
        mesh.num_nodes)
    if os.environ.get('SFEM_HALO', 'p2p') == 'p2p' and halo.enable_p2p(
        torch.float64, device):
      path = 'peer memory'
    else:
      path = 'nccl all_to_all'
  field = lambda c: (np.cos(1.3 * c[:, 0]) * (1.0 + 0.5 * c[:, -1]) +  # noqa: E731
                     0.2 * c[:, 1] ** 2)
  u = torch.as_tensor(field(x0)).to(device)
  y = torch.empty_like(u)
  dot = torch.zeros((), dtype=torch.float64, device=device)
  for _ in range(3):  # both epoch parities
    op.apply_partitioned(u, y, halo, blk.num_interface_elements, lam=0.3,
                         mu=1.0, dot_out=dot)
  ones = torch.ones_like(u)
  b_loc = op.apply(ones, lam=1.0, mu=0.0)
  diag = op.diag()
  if halo is not None:
    halo.exchange_(b_loc)
    halo.exchange_(diag)
  minv = torch.where(diag != 0, 1.0 / diag, torch.zeros_like(diag))
  kw = dict(minv=minv, check_every=7,
            num_interface_elements=blk.num_interface_elements)
  xs, info_fixed = distributed_cg(op, halo, b_loc, tol=0.0, maxiter=cg_iters,
                                  **kw)
  xc, info_conv = distributed_cg(op, halo, b_loc, tol=1e-6, **kw)
  torch.cuda.synchronize()
  mine = (x0, y.cpu().numpy(), float(dot), xs.cpu().numpy(), xc.cpu().numpy())
  if world > 1:
    parts = [None] * world if rank == 0 else None
    dist.gather_object(mine, parts, dst=0)
    halo.disable_p2p()
  else:
    parts = [mine]
  if rank != 0:
    return None

  from oracle import dense
  from swirl_fem_b200.common.premesh_commons import unit_cube_mesh
  from swirl_fem_b200.core.mesh_refiner import refine_premesh
  ref = refine_premesh(unit_cube_mesh(ne, ndim=ndim, a=-1., b=1.), grid1d)
  bmask = ref.finalize_host()['physical_masks']['boundary']
  interior = 1.0 - bmask
  fes = dense.FESpace(deform(ref.node_coords), ref.elements, order + 1,
                      'gauss_lobatto_legendre', order + 1,
                      'gauss_lobatto_legendre')
  gu = field(ref.node_coords)
  gy = fes.apply(gu, lam=0.3, mu=1.0, interior_mask=interior)
  gb = fes.apply(np.ones(ref.num_nodes), 1.0, 0.0, interior)
  gd = fes.stiffness_diag(interior)
  gminv = np.where(gd != 0, 1.0 / np.where(gd != 0, gd, 1.0), 0.0)
  A = lambda v: fes.apply(v, interior_mask=interior)  # noqa: E731
  gx, _ = dense.cg(A, gb, tol=0.0, maxiter=cg_iters, M=lambda r: gminv * r)
  gxc, ginfo = dense.cg(A, gb, tol=1e-6, M=lambda r: gminv * r)
  key = lambda c: np.round(np.asarray(c) * 1e9).astype(np.int64)  # noqa: E731
  gk = key(ref.node_coords)
  order_ = np.lexsort(gk.T[::-1])
  rec = [('', np.int64)] * ndim
  view_g = np.ascontiguousarray(gk[order_]).view(rec).ravel()
  apply_err = x_err = xc_err = 0.0
  total_dot = 0.0
  for c, yv, dv, xv, xcv in parts:
    view_l = np.ascontiguousarray(key(c)).view(rec).ravel()
    l2g = order_[np.searchsorted(view_g, view_l)]
    assert np.array_equal(gk[l2g], key(c))
    apply_err = max(apply_err, np.abs(yv - gy[l2g]).max() / np.abs(gy).max())
    x_err = max(x_err, np.abs(xv - gx[l2g]).max() / np.abs(gx).max())
    xc_err = max(xc_err, np.abs(xcv - gxc[l2g]).max() / np.abs(gxc).max())
    total_dot += dv
  return {
      'mesh': f'{ndim}-D ne={ne} order {order}, {ref.num_nodes} dofs, '
              f'{world} rank(s), halo: {path}',
      'checker': 'oracle.dense (numpy restatement of the reference, fp64) on '
                 'the unpartitioned mesh, rank 0',
      'rel_err': float(apply_err),
      'apply_dot_rel_err': float(abs(total_dot - gu @ gy) / abs(gu @ gy)),
      'cg_fixed_iterations': int(info_fixed['num_iterations']),
      'cg_x_rel_err_after_fixed_iterations': float(x_err),
      'cg_iterations': int(info_conv['num_iterations']),
      'cg_iterations_oracle': int(ginfo['num_iterations']),
      'cg_solution_rel_err': float(xc_err),
      'ok': bool(apply_err <= 1e-12 and x_err <= 1e-9 and abs(
          info_conv['num_iterations'] - ginfo['num_iterations']) <= 1),
  }


# ----------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------


def run_ours(args):
  import torch
  import torch.distributed as dist
  from swirl_fem_b200 import _lib
  from swirl_fem_b200.communication import partition as part
  from swirl_fem_b200.core.interpolation import Nodes1D, NodeType, Quadrature1D
  from swirl_fem_b200.core.mesh import Mesh
  from swirl_fem_b200.core.operator import FusedOperator
  from swirl_fem_b200.core.operator import JacobiPreconditioner
  from swirl_fem_b200.linalg.cg import cg

  world = int(os.environ.get('WORLD_SIZE', 1))
  rank = int(os.environ.get('RANK', 0))
  local_rank = int(os.environ.get('LOCAL_RANK', 0))
  if not torch.cuda.is_available():
    raise SystemExit('bench.py needs a CUDA device (no CPU fallback)')
  torch.cuda.set_device(local_rank)
  device = torch.device('cuda', local_rank)
  if world > 1:
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    dist.init_process_group('nccl', device_id=device)
  assert world == args.gpus or world == 1, (world, args.gpus)

  dtype = torch.float64 if args.dtype == 'f64' else torch.float32
  esz = 8 if args.dtype == 'f64' else 4
  gll = NodeType.GAUSS_LOBATTO_LEGENDRE
  grid1d = Nodes1D.create(args.order + 1, gll)

  # parity of this job's path (same ranks, same halo) before anything is timed
  parity = None
  if not args.no_parity:
    parity = parity_block(rank, world, device, ndim=args.dim, order=args.order)

  t_setup = time.perf_counter()
  blk = part.block_partition(args.ne, args.dim, grid1d, rank, world)
  coords = deform(blk.premesh.node_coords)
  mesh = Mesh.create(coords, blk.premesh.elements, gridpoints_1d=grid1d,
                     device=device, dtype=dtype)
  quad = Quadrature1D.create_from_nodes_1d(grid1d)
  # Only the fused operator is needed: no invjacs / jacdets / quad_coords.
  op = FusedOperator(mesh, quad, dirichlet_mask=blk.dirichlet, with_mass=False)
  if int(os.environ.get('SFEM_VARIANT', '0')):  # developer: tuning variants
    op.set_variant(int(os.environ['SFEM_VARIANT']))
  halo, halo_path = None, None
  if world > 1:
    gathered = [None] * world
    dist.all_gather_object(gathered, np.sort(blk.interface_global))
    halo = part.halo_plan_from_interfaces(
        rank, blk.interface_local, blk.interface_global, gathered,
        mesh.num_nodes)
    # shared-dof exchange over peer memory (NVLink stores issued from inside
    # the apply kernel); SFEM_HALO=nccl keeps the pack / all_to_all / unpack
    # path for comparison
    if os.environ.get('SFEM_HALO', 'p2p') == 'p2p':
      halo_path = ('peer memory (NVLink stores from the companion kernel '
                   'that runs next to the apply)'
                   if halo.enable_p2p(dtype, device) else
                   'nccl all_to_all (peer mapping failed)')
      # where the canonical sum runs: 1 in the apply kernel's own CTAs, 2 in
      # the wait kernel concurrently with the interior elements, 0 after the
      # apply (developer switch; the library default applies otherwise)
      mode = os.environ.get('SFEM_HALO_FUSE_UNPACK')
      if mode is not None:
        halo.p2p_set_option(1, int(mode))
        halo_path += f', canonical-sum mode {int(mode)}'
    else:
      halo_path = 'nccl all_to_all'
  torch.cuda.synchronize()
  t_setup = time.perf_counter() - t_setup

  num_global = blk.num_global_dofs
  num_local_nodes = mesh.num_elements * mesh.num_nodes_per_element
  gen = torch.Generator(device=device).manual_seed(1234 + rank)
  x = torch.randn(mesh.num_nodes, dtype=dtype, device=device, generator=gen)
  y = torch.empty_like(x)

  def step():
    # N > 1: interface elements first, halo exchange overlapped with the
    # interior elements (FusedOperator.apply_partitioned)
    op.apply_partitioned(x, y, halo, blk.num_interface_elements, lam=0.0,
                         mu=1.0)

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  # lazy zero fill (companion kernel): checked against the eager fill on this
  # very workload before anything is timed; any doubt -> eager
  zero_fill = 'eager: before the apply'
  if getattr(op, '_lazy', None) and world == 1:
    y_lazy = op.apply(x, lam=0.0, mu=1.0).clone()
    y_lazy2 = op.apply(x, lam=0.0, mu=1.0).clone()   # counters were reset
    op.disable_lazy_zero()
    y_eager = op.apply(x, lam=0.0, mu=1.0)
    scale = float(y_eager.abs().max())
    diff = max(float((y_lazy - y_eager).abs().max()),
               float((y_lazy2 - y_eager).abs().max())) / scale
    del y_lazy, y_lazy2, y_eager
    msg = None
    if diff <= (1e-13 if dtype == torch.float64 else 2e-5):
      op.enable_lazy_zero()
      op.apply(x, lam=0.0, mu=1.0, out=y)
      msg = op.lazy_zero_timed_out()
    if op._lazy and not msg:
      zero_fill = ('lazy: companion kernel next to the apply '
                   f'(sfem_op_set_lazy_zero); rel. diff vs eager {diff:.1e}')
    else:
      op.disable_lazy_zero()
      zero_fill += f' (lazy fill rejected: diff {diff:.1e}, {msg})'
      print('bench: ' + zero_fill, file=sys.stderr)

  for _ in range(max(args.warmup, 3)):
    step()
  barrier()
  sampler = ClockSampler(local_rank)
  if rank == 0:
    sampler.start()
  launches0 = _lib.launch_count()
  ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
  barrier()
  ev[0].record()
  for i in range(args.steps):
    step()
    ev[i + 1].record()
  barrier()
  launches = _lib.launch_count() - launches0
  clocks = sampler.stop() if rank == 0 else None
  total_ms = ev[0].elapsed_time(ev[-1])
  if world > 1:
    t = torch.tensor([total_ms], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
  ms_per_step = total_ms / args.steps
  value = num_global / (ms_per_step * 1e-3) / 1e9
  # where inside the last fused launch the exchange happened (globaltimer
  # stamps written by the kernel): evidence that it overlaps the interior
  halo_timeline = None
  if halo is not None and rank == 0:
    try:
      if halo.p2p_handle(x) is not None:
        halo_timeline = halo.p2p_debug_times(device)
        halo_timeline['unit'] = 'us from the start of the apply kernel (rank 0)'
    except Exception as e:  # diagnostics only  pylint: disable=broad-except
      halo_timeline = {'error': str(e)}

  # roofline of the dominant kernel: per-rank algorithmic bytes / apply time
  # (the un-fused kernel instance has not run yet when the step is the
  # halo-fused launch: warm it up so module loading is not timed)
  for _ in range(2):
    op.apply(x, lam=0.0, mu=1.0, out=y)
  torch.cuda.synchronize()
  apply_ms = []
  for i in range(min(args.steps, 10)):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(
        enable_timing=True)
    a.record()
    op.apply(x, lam=0.0, mu=1.0, out=y)
    b.record()
    torch.cuda.synchronize()
    apply_ms.append(a.elapsed_time(b))
  apply_ms = float(np.mean(apply_ms))
  # every rank's own numbers (the job runs at the pace of its slowest GPU)
  per_rank = None
  if world > 1:
    mine_ms = (float(ev[0].elapsed_time(ev[-1])) / args.steps, apply_ms)
    per_rank = [None] * world
    dist.all_gather_object(per_rank, mine_ms)
  abytes = algorithmic_bytes(mesh.num_nodes, num_local_nodes, args.dim, esz)
  peak, peak_src = measured_peak_gbs()
  # N = 1: the timed region IS this kernel (+ its zero fill), launch after
  # launch, so its own average is the kernel's average launch duration; the
  # launches timed one by one above (a synchronise between them, the GPU idles
  # and re-ramps every time) are kept as `kernel_ms_isolated`.  N > 1: the
  # timed step also holds the exchange, so the isolated local kernel is used.
  isolated_ms = apply_ms
  if world == 1:
    apply_ms = ms_per_step
  achieved = abytes / (apply_ms * 1e-3) / 1e9
  # DRAM traffic of the kernel from a recorded `ncu --set full` capture of this
  # exact workload (profiles/traffic.json); null for other workloads
  traffic, traffic_src = None, None
  try:
    with open(os.path.join(ROOT, 'profiles', 'traffic.json')) as f:
      rec = json.load(f).get(
          f'{args.dim}d_p{args.order}_{args.dtype}_ne{args.ne}_n{world}')
    if rec:
      traffic, traffic_src = rec['traffic_bytes_per_launch'], rec['source']
  except (OSError, ValueError, KeyError):
    pass
  roofline = {'bound': 'hbm', 'achieved': achieved, 'peak': peak,
              'unit': 'GB/s', 'frac': achieved / peak, 'traffic': traffic,
              'traffic_source': traffic_src,
              'peak_source': peak_src,
              'kernel': 'apply3d_v2_kernel (memset of the shared-dof prefix + '
                        'fused gather/operator/scatter)',
              'algorithmic_bytes_per_launch': abytes,
              'kernel_ms': apply_ms,
              'kernel_ms_source': (
                  'CUDA events over the timed region (mean per step)'
                  if world == 1 else
                  'CUDA events around single launches of the local kernel'),
              'kernel_ms_isolated': isolated_ms,
              'frac_of_nominal_8TBs': achieved / 8000.0}

  # end to end through the public API with pinned HOST buffers: every step
  # uploads its x, applies, downloads its y (core.operator.HostPipeline: the
  # upload of step i+1, the apply of step i and the download of step i-1
  # overlap; PCIe is full duplex)
  e2e = None
  if not args.no_e2e:
    from swirl_fem_b200.core.operator import HostPipeline
    n_e2e = max(4, min(args.steps, 8))
    xh = [torch.empty(mesh.num_nodes, dtype=dtype).pin_memory()
          for _ in range(2)]
    yh = [torch.empty(mesh.num_nodes, dtype=dtype).pin_memory()
          for _ in range(2)]
    for t in xh:
      t.copy_(x)
    # the device-resident result the pipeline's output is compared with
    op.apply_partitioned(x, y, halo, blk.num_interface_elements, lam=0.0,
                         mu=1.0)
    pipe = HostPipeline(op, halo, blk.num_interface_elements, depth=2)
    for i in range(2):
      pipe.submit(xh[i % 2], yh[i % 2])
    pipe.synchronize()
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(
        enable_timing=True)
    a.record()
    for i in range(n_e2e):
      pipe.submit(xh[i % 2], yh[i % 2])
    pipe.drain()
    b.record()
    barrier()
    e2e_ms = a.elapsed_time(b) / n_e2e
    if world > 1:
      t = torch.tensor([e2e_ms], dtype=torch.float64, device=device)
      dist.all_reduce(t, op=dist.ReduceOp.MAX)
      e2e_ms = float(t.item())
    y_check = float((yh[(n_e2e - 1) % 2].to(device) - y).abs().max())
    e2e = {'value': num_global / (e2e_ms * 1e-3) / 1e9, 'unit': 'GDOF/s',
           'h2d_bytes_per_step': int(mesh.num_nodes * esz),
           'd2h_bytes_per_step': int(mesh.num_nodes * esz),
           'ms_per_step': e2e_ms, 'steps': n_e2e,
           'pipeline': 'HostPipeline depth 2: H2D(i+1) | apply(i) | D2H(i-1) '
                       'on three streams, pinned host buffers',
           'max_abs_diff_vs_device_resident_result': y_check}
    del pipe

  # fused CG: fixed iteration count (tol = 0 never converges early)
  cg_info = None
  if args.cg_iters > 0:
    from swirl_fem_b200.communication.dist_cg import distributed_cg
    rhs = torch.where(torch.as_tensor(blk.dirichlet, device=device), 0.0,
                      1.0).to(dtype)
    diag = op.diag()
    if halo is not None:
      halo.exchange_(diag)
    minv_t = torch.where(diag != 0, 1.0 / diag, torch.zeros_like(diag))

    # the two dot products per iteration are all-reduced over peer memory
    # inside the fused step kernel (distributed_cg creates the exchange)
    sx = None

    def solve(iters):
      if world == 1:
        return cg(op.bind(0.0, 1.0), rhs, tol=0.0, maxiter=iters,
                  M=JacobiPreconditioner(minv_t), check_every=iters)
      return distributed_cg(
          op, halo, rhs, tol=0.0, maxiter=iters, minv=minv_t,
          check_every=iters,
          num_interface_elements=blk.num_interface_elements,
          scalar_exchange=sx)

    solve(3)  # warm-up
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(
        enable_timing=True)
    a.record()
    _, info = solve(args.cg_iters)
    b.record()
    barrier()
    cg_ms = a.elapsed_time(b) / max(info['num_iterations'], 1)
    if world > 1:
      t = torch.tensor([cg_ms], dtype=torch.float64, device=device)
      dist.all_reduce(t, op=dist.ReduceOp.MAX)
      cg_ms = float(t.item())
    # the same solve end to end: right-hand side from pinned host memory,
    # the fixed number of iterations, solution back to the host
    cg_e2e_ms = None
    if not args.no_e2e:
      bh = torch.empty(mesh.num_nodes, dtype=dtype).pin_memory()
      bh.copy_(rhs)
      xh_out = torch.empty(mesh.num_nodes, dtype=dtype).pin_memory()
      barrier()
      a.record()
      rhs.copy_(bh, non_blocking=True)
      xs, _ = solve(args.cg_iters)
      xh_out.copy_(xs, non_blocking=True)
      b.record()
      barrier()
      cg_e2e_ms = a.elapsed_time(b)
      if world > 1:
        t = torch.tensor([cg_e2e_ms], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        cg_e2e_ms = float(t.item())
    cg_bytes = abytes + 11 * esz * mesh.num_nodes
    cg_info = {'iterations': info['num_iterations'], 'ms_per_iteration': cg_ms,
               'gdof_per_s_iter': num_global / (cg_ms * 1e-3) / 1e9,
               'preconditioner': 'jacobi',
               'roofline_frac': cg_bytes / (cg_ms * 1e-3) / 1e9 / peak,
               'e2e_solve_ms': cg_e2e_ms,
               'e2e_note': 'b uploaded from pinned host memory, '
                           f'{args.cg_iters} iterations, x downloaded',
               'launches_per_iteration': 2 if world == 1 else 3,
               'driver': 'sfem_cg (apply + one fused step kernel per '
                         'iteration)' if world == 1 else
                         'distributed_cg -> sfem_cg_iterate (apply, '
                         'companion exchange kernel next to it, fused step '
                         'kernel with the scalar all-reduces over peer memory)'
                         if halo_path and halo_path.startswith('peer') else
                         'distributed_cg (building blocks + NCCL scalars)'}

  # configs 1, 2, 3 and a config-5 subset (N = 1 only): bench_extra.py
  extra = None
  if rank == 0 and world == 1 and not args.no_extra:
    import bench_extra
    del x, y
    torch.cuda.empty_cache()
    extra = bench_extra.run_all(device, peak, ClockSampler, local_rank)

  cpu = None
  if rank == 0 and world == 1 and not args.no_cpu_baseline:
    cpu = cpu_oracle_throughput(args.dim, args.order)
    cpu = {k: cpu[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')}

  if rank == 0:
    line = {
        'metric': METRIC if (args.dim, args.order, args.dtype) == (
            3, 7, 'f64') else (f'GDOF/s matrix-free Laplacian apply '
                               f'({args.dim}-D, order {args.order}, '
                               f'{args.dtype})'),
        'value': value, 'unit': 'GDOF/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': max(args.warmup, 3),
        'ms_per_step': ms_per_step, 'higher_is_better': True,
        'scaling': 'strong', 'vs_baseline': None, 'dtype': args.dtype,
        'data': 'synthetic',
        'config': bench_config(args, world),
        'details': {
            'halo_exchange': halo_path,
            'halo_timeline': halo_timeline,
            'local_dofs_rank0': mesh.num_nodes,
            'elements_rank0': mesh.num_elements,
            'geometric_factors_gb_per_rank': op.geom.numel() * esz / 1e9,
            'per_rank_ms': None if per_rank is None else {
                'step_with_exchange': [round(a, 5) for a, _ in per_rank],
                'local_kernel_without_exchange': [round(b, 5)
                                                  for _, b in per_rank]},
            'zero_fill': zero_fill + (
                '' if not (getattr(op, '_lazy', None) and world == 1 and
                           op.lazy_zero_timed_out())
                else ' -- A WAIT TIMED OUT DURING THE RUN: numbers invalid'),
            'setup_s': t_setup,
        },
        'clocks': clocks,
        'e2e': e2e,
        'gpu_launches': int(launches),
        'roofline': roofline,
        'cpu_baseline': cpu,
        'cg': cg_info,
        'parity': parity,
        'extra': extra,
    }
    print(json.dumps(line), flush=True)
  if world > 1:
    dist.barrier()
    dist.destroy_process_group()


def main():
  args = parse_args()
  if args.impl == 'reference':
    run_reference(args)
  else:
    run_ours(args)


if __name__ == '__main__':
  main()
