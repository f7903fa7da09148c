"""Extra measured workloads for the driver's bench line (`bench.py`, N = 1):
BASELINE.json configs 1, 2, 3 and a subset of config 5, each with its
algorithmic bytes, roofline fraction and a clock record.

Every timing: CUDA events around ONE call on the launching stream, L2 flushed
(a 512 MB device memset) before every timed call -- several of these
workloads are smaller than the 126 MB L2 -- mean over `reps` calls after
warm-up.  Nothing here is on the path of the headline number; the oracle is
used only for config 1's CPU leg (`cpu_baseline`-style, bounded).
"""

from __future__ import annotations

import time

import numpy as np


def _deform(x):
  ndim = x.shape[-1]
  perm = np.roll(np.arange(ndim), 1)
  return x + 0.08 * np.sin(np.pi * x[:, perm]) * (1 - x ** 2)


class _Timer:
  """Per-call CUDA-event timing with an L2 flush before every timed call."""

  def __init__(self, device):
    import torch
    self.torch = torch
    self.flush_buf = torch.empty(512 << 20, dtype=torch.uint8, device=device)

  def __call__(self, fn, reps=10, warm=3):
    torch = self.torch
    for _ in range(warm):
      fn()
    times = []
    for _ in range(reps):
      self.flush_buf.zero_()
      a = torch.cuda.Event(enable_timing=True)
      b = torch.cuda.Event(enable_timing=True)
      a.record()
      fn()
      b.record()
      torch.cuda.synchronize()
      times.append(a.elapsed_time(b))
    return float(np.mean(times)), float(np.min(times))


def _apply_bytes(num_nodes, num_local, ndim, esz, with_mass):
  g = ndim * (ndim + 1) // 2 + (1 if with_mass else 0)
  return 2 * esz * num_nodes + num_local * (g * esz + 4)


def _coons_disk(x):
  """Coons patch of the unit disk (examples/poisson_test.py:70-90), applied to
  0.9 * x: the patch itself is degenerate at the four corners of the square
  (det J = 0 there, "a high num_elements_per_dim may create degenerate
  elements"), which a collocated GLL rule would sample."""
  x = 0.9 * x
  r = 1 / np.sqrt(2)
  return np.stack([
      x[:, 0] * (np.cos(np.pi * x[:, 1] / 4) - r) + np.sin(np.pi * x[:, 0] / 4),
      x[:, 1] * (np.cos(np.pi * x[:, 0] / 4) - r) + np.sin(np.pi * x[:, 1] / 4),
  ], -1)


def _unstructured(premesh, seed):
  """Seeded random element order + per-element rotation of the vertex listing
  (keeps det J > 0): exercises the refiner's orientation handling."""
  from swirl_fem_b200.core.premesh import Premesh
  rng = np.random.default_rng(seed)
  el = np.asarray(premesh.elements)[rng.permutation(premesh.num_elements)]
  k = rng.integers(4, size=len(el))
  quad = el.reshape(-1, 2, 2)
  out = np.empty_like(quad)
  for r in range(4):
    sel = k == r
    out[sel] = np.rot90(quad[sel], k=r, axes=(1, 2))
  return Premesh.create(node_coords=premesh.node_coords,
                        elements=out.reshape(-1, 4).astype(np.int32),
                        physical_groups=premesh.physical_groups,
                        periodic_links=premesh.periodic_links)


def _refined(ndim, ne, n1d, kind='deformed', seed=None):
  from swirl_fem_b200.common.premesh_commons import unit_cube_mesh
  from swirl_fem_b200.core.interpolation import Nodes1D, NodeType
  from swirl_fem_b200.core.mesh_refiner import refine_premesh
  pm = unit_cube_mesh(ne, ndim=ndim, a=-1., b=1.)
  if seed is not None:
    pm = _unstructured(pm, seed)
  refined = refine_premesh(
      pm, Nodes1D.create(n1d, NodeType.GAUSS_LOBATTO_LEGENDRE))
  x = refined.node_coords
  coords = _coons_disk(x) if kind == 'disk' else _deform(x)
  bmask = refined.finalize_host()['physical_masks']['boundary']
  return refined, coords, bmask


def _operator(refined, coords, bmask, n1d, dtype, device, with_mass):
  from swirl_fem_b200.core.interpolation import Nodes1D, NodeType, Quadrature1D
  from swirl_fem_b200.core.mesh import Mesh
  from swirl_fem_b200.core.operator import FusedOperator
  grid1d = Nodes1D.create(n1d, NodeType.GAUSS_LOBATTO_LEGENDRE)
  mesh = Mesh.create(coords, refined.elements, gridpoints_1d=grid1d,
                     device=device, dtype=dtype)
  quad = Quadrature1D.create_from_nodes_1d(grid1d)
  return mesh, FusedOperator(mesh, quad, dirichlet_mask=bmask,
                             with_mass=with_mass)


def _apply_case(timer, refined, coords, bmask, ndim, n1d, dtype_name, device,
                peak, lam, mu, with_mass, cg_iters=0):
  import torch
  from swirl_fem_b200.core.operator import JacobiPreconditioner
  from swirl_fem_b200.linalg.cg import cg
  dtype = torch.float64 if dtype_name == 'f64' else torch.float32
  esz = 8 if dtype_name == 'f64' else 4
  mesh, op = _operator(refined, coords, bmask, n1d, dtype, device, with_mass)
  gen = torch.Generator(device=device).manual_seed(7)
  x = torch.randn(mesh.num_nodes, dtype=dtype, device=device, generator=gen)
  y = torch.empty_like(x)
  ms, ms_min = timer(lambda: op.apply(x, lam=lam, mu=mu, out=y))
  nloc = mesh.num_elements * mesh.num_nodes_per_element
  abytes = _apply_bytes(mesh.num_nodes, nloc, ndim, esz, with_mass)
  out = {
      'dofs': int(mesh.num_nodes), 'elements': int(mesh.num_elements),
      'apply_ms': ms, 'apply_ms_min': ms_min,
      'gdof_per_s': mesh.num_nodes / ms / 1e6,
      'algorithmic_bytes': int(abytes),
      'achieved_gbs': abytes / ms / 1e6,
      'frac': abytes / ms / 1e6 / peak,
  }
  if cg_iters:
    minv = op.jacobi_minv(lam, mu)
    rhs = torch.where(torch.as_tensor(bmask, device=device) != 0, 0.0,
                      1.0).to(dtype)
    solve = lambda: cg(op.bind(lam, mu), rhs, tol=0.0, maxiter=cg_iters,  # noqa: E731
                       M=JacobiPreconditioner(minv), check_every=cg_iters)
    cms, _ = timer(solve, reps=3, warm=1)
    cbytes = abytes + 11 * esz * mesh.num_nodes
    out['pcg'] = {
        'iterations': cg_iters, 'ms_per_iteration': cms / cg_iters,
        'algorithmic_bytes_per_iteration': int(cbytes),
        'frac': cbytes / (cms / cg_iters) / 1e6 / peak,
        'preconditioner': 'jacobi'}
  del op, mesh
  return out


def config1(device, timer):
  """C1: 2-D Poisson, structured 32x32, GLL order 4, (Jacobi-)PCG, tol 1e-5 --
  the ONE config whose whole problem also runs on the CPU oracle, so this is
  a same-problem comparison (reference algorithm restated in numpy)."""
  import torch
  from oracle import dense
  from swirl_fem_b200.core.operator import JacobiPreconditioner
  from swirl_fem_b200.linalg.cg import cg
  n1d = 5
  refined, coords, bmask = _refined(2, 32, n1d)
  mesh, op = _operator(refined, coords, bmask, n1d, torch.float64, device,
                       True)
  b = op.apply(torch.ones(mesh.num_nodes, dtype=torch.float64, device=device),
               lam=1.0, mu=0.0)
  minv = op.jacobi_minv()
  res = {}

  def solve_plain():
    res['plain'] = cg(op.bind(0.0, 1.0), b, tol=1e-5)

  def solve_jacobi():
    res['jacobi'] = cg(op.bind(0.0, 1.0), b, tol=1e-5,
                       M=JacobiPreconditioner(minv))

  ms_plain, _ = timer(solve_plain, reps=3, warm=1)
  ms_jac, _ = timer(solve_jacobi, reps=3, warm=1)
  interior = 1.0 - bmask
  fes = dense.FESpace(coords, refined.elements, n1d, 'gauss_lobatto_legendre',
                      n1d, 'gauss_lobatto_legendre')
  gb = fes.apply(np.ones(refined.num_nodes), 1.0, 0.0, interior)
  gd = fes.stiffness_diag(interior)
  gminv = np.where(gd != 0, 1.0 / np.where(gd != 0, gd, 1.0), 0.0)
  A = lambda v: fes.apply(v, interior_mask=interior)  # noqa: E731
  t0 = time.perf_counter()
  gx, ginfo = dense.cg(A, gb, tol=1e-5, M=lambda r: gminv * r)
  cpu_ms = (time.perf_counter() - t0) * 1e3
  xj = res['jacobi'][0].cpu().numpy()
  return {
      'workload': '2-D Poisson 32x32 quads, GLL order 4, 16641 dofs, fp64, '
                  'deformed elements, tol 1e-5',
      'dofs': int(mesh.num_nodes),
      'cg_iterations': int(res['plain'][1]['num_iterations']),
      'cg_solve_ms': ms_plain,
      'pcg_jacobi_iterations': int(res['jacobi'][1]['num_iterations']),
      'pcg_jacobi_solve_ms': ms_jac,
      'oracle_cpu': {'pcg_jacobi_iterations': int(ginfo['num_iterations']),
                     'solve_ms': cpu_ms, 'kind': 'port',
                     'x_rel_err_gpu_vs_oracle': float(
                         np.abs(xj - gx).max() / np.abs(gx).max())},
      'note': 'launch-bound at this size (2 launches per iteration)',
  }


def config2(device, timer, peak):
  """C2: 2-D Helmholtz on an unstructured quad mesh (Coons-patch disk, seeded
  element permutation + vertex-order rotations), order 8, 4.0 M dofs, fp64;
  H = lam*M + mu*K with the reference's lam = beta_3/dt = (11/6)/1e-3, mu = 1
  (navier_stokes.py:431)."""
  n1d = 9
  refined, coords, bmask = _refined(2, 250, n1d, kind='disk', seed=11)
  out = _apply_case(timer, refined, coords, bmask, 2, n1d, 'f64', device, peak,
                    lam=(11.0 / 6.0) / 1e-3, mu=1.0, with_mass=True,
                    cg_iters=50)
  out['workload'] = ('2-D Helmholtz, unstructured quads (disk, 250^2 elements, '
                     'random element order / vertex rotations), GLL order 8, '
                     'fp64, lam=(11/6)/1e-3, mu=1')
  return out


def config3(device, timer):
  """C3: one `stokes_one_step` (BDF3/EXT, velocity Helmholtz CG with
  M = exchange, pressure CG on E = D Q D^T with the null-space projector) on
  the y-periodic channel of navier_stokes_test.py:39-77 at 64x64 elements,
  order 7."""
  import torch
  from swirl_fem_b200.common.premesh_commons import unit_cube_mesh
  from swirl_fem_b200.navier_stokes import navier_stokes as ns
  ne, order, dt, k = 64, 7, 1e-3, 3
  pm = unit_cube_mesh(ne, ndim=2, periodic_dims=(1,))
  x = np.asarray(pm.node_coords, dtype=np.float64)
  x = np.stack([2 * x[:, 0] - 1, 2 * np.pi * x[:, 1] - np.pi], -1)
  x[:, 0] += 0.1 * np.sin(x[:, 1]) * (1 - x[:, 0] ** 2)
  pm = pm.replace(node_coords=x)
  sem = ns.StokesSEM.create(
      pm, boundary_conditions={'boundary': (ns.BCType.DIRICHLET, 0.0)},
      order=order, dtype=torch.float64)
  vm, pmesh = sem.velocity.mesh, sem.pressure.pspace.mesh
  gen = torch.Generator(device=device).manual_seed(0)
  u = torch.randn(vm.num_nodes, 2, dtype=torch.float64, device=device,
                  generator=gen) * sem.velocity.interior_mask
  p = torch.randn(pmesh.num_nodes, dtype=torch.float64, device=device,
                  generator=gen)
  us = [u * (1.0 - 0.01 * i) for i in range(k)]
  ps = [p * 0.0 for _ in range(k)]
  aux_box = {}

  def step():
    _, _, aux_box['aux'] = sem.stokes_one_step(
        us, ps, f=0, mu=1e-2, dt=dt, time_order=k, tol=1e-5, atol=1e-4)

  ms, ms_min = timer(step, reps=3, warm=1)
  aux = aux_box['aux']
  ops = {}
  for name, fn in (('A', lambda: sem.A(u)), ('D', lambda: sem.D(u)),
                   ('Dt', lambda: sem.Dt(p)), ('C', lambda: sem.C(u)),
                   ('E', lambda: sem.E(p, dt=dt, time_order=k))):
    ops[name] = timer(fn, reps=5, warm=2)[0] * 1e3
  return {
      'workload': f'2-D Stokes step, {ne}x{ne} y-periodic channel, order '
                  f'{order}, BDF{k}, fp64',
      'velocity_dofs': int(vm.num_nodes), 'pressure_dofs': int(pmesh.num_nodes),
      'stokes_one_step_ms': ms, 'stokes_one_step_ms_min': ms_min,
      'u_star_cg_iterations': int(aux['u_star_info']['num_iterations']),
      'dp_cg_iterations': int(aux['dp_info']['num_iterations']),
      'operator_us': ops,
  }


def sweep(device, timer, peak):
  """C5 subset: order 4 and 8, 2-D (4.0 M dofs as SURVEY section 8d, and 16 M
  dofs where the fixed costs of a launch no longer show) and 3-D (8.1 M dofs),
  fp32 and fp64, Laplacian apply."""
  out = {}
  for ndim, p, ne, tag in ((2, 4, 500, ''), (2, 8, 250, ''),
                           (2, 4, 1000, '_16M'), (2, 8, 500, '_16M'),
                           (3, 4, 50, ''), (3, 8, 25, '')):
    refined, coords, bmask = _refined(ndim, ne, p + 1)
    for dt in ('f64', 'f32'):
      r = _apply_case(timer, refined, coords, bmask, ndim, p + 1, dt, device,
                      peak, lam=0.0, mu=1.0, with_mass=False)
      out[f'{ndim}d_p{p}_{dt}{tag}'] = {k: r[k] for k in (
          'dofs', 'apply_ms', 'gdof_per_s', 'algorithmic_bytes', 'frac')}
    del refined, coords, bmask
  return out


def run_all(device, peak, clock_sampler_cls, gpu_index):
  """Returns the `extra` object of the bench line."""
  import torch
  timer = _Timer(device)
  out = {'timing': 'CUDA events per call, 512 MB L2 flush before every timed '
                   'call, mean over calls; % of the measured HBM peak '
                   f'{peak:.0f} GB/s'}
  for name, fn in (('c1', lambda: config1(device, timer)),
                   ('c2', lambda: config2(device, timer, peak)),
                   ('c3', lambda: config3(device, timer)),
                   ('sweep', lambda: sweep(device, timer, peak))):
    sampler = clock_sampler_cls(gpu_index)
    sampler.start()
    t0 = time.perf_counter()
    try:
      res = fn()
    except Exception as e:  # an extra must never take the headline down  pylint: disable=broad-except
      res = {'error': f'{type(e).__name__}: {e}'}
    torch.cuda.synchronize()
    clocks = sampler.stop()
    if isinstance(res, dict):
      res['clocks'] = clocks
      res['wall_s'] = time.perf_counter() - t0
    out[name] = res
    torch.cuda.empty_cache()
  return out
