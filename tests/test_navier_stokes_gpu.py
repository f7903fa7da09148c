"""GPU parity of `swirl_fem_b200.navier_stokes`.

CUDA path (C ABI kernels behind the reference's StokesSEM API) against
  * the reference's own outputs (`tests/golden/navier_stokes.npz`),
  * the numpy oracle `oracle/dense_ns.py` on curved meshes (1e-12 relative),
  * the analytical known answers of the reference's tests
    (`swirl_fem/navier_stokes/navier_stokes_test.py:73-358`, same tolerances),
  * the oracle's `stokes_one_step` (CG iteration counts +-1).
"""

import os

import numpy as np
import pytest
import torch

from oracle import dense_ns
from tests import helpers
from tests.conftest import load_golden

pytestmark = pytest.mark.gpu

DIRICHLET = None


@pytest.fixture(scope='module', autouse=True)
def _need_cuda():
  if not torch.cuda.is_available():
    pytest.skip('needs a CUDA device')
  from swirl_fem_b200 import _lib
  _lib.lib()


def dev(x, dtype=torch.float64):
  return torch.as_tensor(np.ascontiguousarray(x)).cuda().to(dtype)


def rel_err(a, b):
  a = np.asarray(a.cpu() if isinstance(a, torch.Tensor) else a, np.float64)
  b = np.asarray(b, np.float64)
  return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def _sem(premesh, order, dtype=torch.float64):
  from swirl_fem_b200.navier_stokes import navier_stokes as ns
  return ns.StokesSEM.create(
      premesh, boundary_conditions={'boundary': (ns.BCType.DIRICHLET, 0.0)},
      order=order, dtype=dtype)


def test_coefficients():
  from swirl_fem_b200.navier_stokes import navier_stokes as ns
  for k in (1, 2, 3, 4):
    np.testing.assert_allclose(ns.bdfk_coeffs(k), dense_ns.bdfk_coeffs(k),
                               atol=1e-13)
  for k in (1, 2, 3):
    np.testing.assert_allclose(ns.extk_coeffs(k), dense_ns.extk_coeffs(k),
                               atol=1e-13)


def _all_operators(sem, u, p, dt=1e-3, k=3):
  from swirl_fem_b200.navier_stokes import navier_stokes as ns
  return {
      'A': sem.A(u), 'B': sem.B(u), 'Bi': sem.Bi(u), 'C': sem.C(u),
      'D': sem.D(u), 'Dt': sem.Dt(p), 'Q': sem.Q(u, dt=dt, time_order=k),
      'E': sem.E(p, dt=dt, time_order=k), 'filter': sem.filter(u, alpha=0.05),
      'vorticity': sem.vorticity(u), 'pressure_B': sem.pressure.B(p),
      'project': ns._pressure_project_out_nullspace(sem, p),  # pylint: disable=protected-access
      'A_local': sem.velocity.A_local(sem.velocity.gather(u)),
      'D_local': sem.D_local(sem.velocity.gather(u)),
      'Dt_local': sem.Dt_local(sem.pressure.gather(p)),
  }


def test_operators_match_reference_golden():
  """Every operator on the reference's own inputs / outputs (tiny mesh)."""
  g = load_golden('navier_stokes')
  pm = helpers.stokes_vortices_premesh(int(g['ne']), curved=0.1)
  sem = _sem(pm, int(g['order']))
  assert np.array_equal(sem.velocity.mesh.elements.cpu().numpy(),
                        g['v_elements'])
  assert np.array_equal(sem.pressure.pspace.mesh.elements.cpu().numpy(),
                        g['p_elements'])
  assert rel_err(sem.velocity_mass_diag, g['velocity_mass_diag']) < 1e-13
  assert rel_err(sem.velocity.diag_qqt, g['diag_qqt']) == 0
  got = _all_operators(sem, dev(g['u']), dev(g['p']))
  for name, val in got.items():
    assert tuple(val.shape) == g[name].shape, name
    assert rel_err(val, g[name]) < 1e-12, (name, rel_err(val, g[name]))


@pytest.mark.parametrize('ne,order', [(3, 4), (2, 7), (4, 5)])
def test_operators_match_oracle_curved(ne, order):
  pm = helpers.stokes_vortices_premesh(ne, curved=0.12)
  sem = _sem(pm, order)
  vmesh, pmesh = helpers.stokes_oracle_meshes(pm, order)
  osem = dense_ns.StokesSEM(vmesh, pmesh, order)
  rng = np.random.default_rng(7)
  u = rng.standard_normal((vmesh['node_coords'].shape[0], 2))
  p = rng.standard_normal(pmesh['node_coords'].shape[0])
  dt, k = 1e-3, 3
  want = {
      'A': osem.A(u), 'B': osem.B(u), 'Bi': osem.Bi(u), 'C': osem.C(u),
      'D': osem.D(u), 'Dt': osem.Dt(p), 'Q': osem.Q(u, dt, k),
      'E': osem.E(p, dt, k), 'filter': osem.filter(u, 0.05),
      'vorticity': osem.vorticity(u), 'pressure_B': osem.pressure_B(p),
      'project': osem.project_out_nullspace(p),
      'A_local': osem.vspace.vector_stiffness_local(osem.v_gather(u)),
      'D_local': osem.D_local(osem.v_gather(u)),
      'Dt_local': osem.Dt_local(osem.pspace.gather(p)),
  }
  got = _all_operators(sem, dev(u), dev(p))
  for name in want:
    assert rel_err(got[name], want[name]) < 1e-12, (
        name, rel_err(got[name], want[name]))


def test_operators_fp32():
  pm = helpers.stokes_vortices_premesh(3, curved=0.1)
  sem = _sem(pm, 5, dtype=torch.float32)
  vmesh, pmesh = helpers.stokes_oracle_meshes(pm, 5)
  osem = dense_ns.StokesSEM(vmesh, pmesh, 5)
  rng = np.random.default_rng(3)
  u = rng.standard_normal((vmesh['node_coords'].shape[0], 2))
  p = rng.standard_normal(pmesh['node_coords'].shape[0])
  for got, want in ((sem.A(dev(u, torch.float32)), osem.A(u)),
                    (sem.D(dev(u, torch.float32)), osem.D(u)),
                    (sem.Dt(dev(p, torch.float32)), osem.Dt(p)),
                    (sem.C(dev(u, torch.float32)), osem.C(u))):
    assert got.dtype == torch.float32
    assert rel_err(got, want) < 1e-5


def test_general_covector_rejects_foreign_placeholder():
  pm = helpers.stokes_vortices_premesh(2)
  sem = _sem(pm, 3)
  vs, ps = sem.velocity.vspace, sem.pressure.pspace
  form = lambda v, q: (lambda x: q(x) * v(x)[0])  # noqa: E731
  ul = sem.velocity.gather(torch.ones(vs.mesh.num_nodes, 2, device='cuda',
                                      dtype=torch.float64))
  with pytest.raises(ValueError, match='placeholder'):
    vs.local_covector(form, (vs.vector_function(ul),
                             ps.scalar_function(None)))


# -- analytical known answers (navier_stokes_test.py), 9 x 9, order 7 -----------


@pytest.fixture(scope='module')
def vortices():
  pm = helpers.stokes_vortices_premesh(9)
  sem = _sem(pm, 7)
  vc = sem.velocity.mesh.node_coords.cpu().numpy()
  pc = sem.pressure.pspace.mesh.node_coords.cpu().numpy()

  def state(t):
    u, p = helpers.stokes_reference_soln(vc, pc, t)
    return dev(u), dev(p)
  return sem, state, pm


def test_mesh_counts(vortices):
  """navier_stokes_test.py:73-77 + refined sizes."""
  sem, _, pm = vortices
  assert pm.num_elements == 81 and pm.num_nodes == 100
  assert sem.velocity.mesh.num_elements == 81
  assert sem.pressure.pspace.mesh.num_nodes == 81 * 36


def test_analytical_momentum_divergence_bdf(vortices):
  """navier_stokes_test.py:79-131."""
  from swirl_fem_b200.navier_stokes import navier_stokes as ns
  sem, state, _ = vortices
  u, p = state(0.0)
  _, sigma = helpers.stokes_reference_soln_params()
  err = sem.velocity.exchange(sem.B(sigma * u) + sem.A(u) - sem.Dt(p))
  assert float(err.abs().max()) < 1e-7
  assert float(sem.D(u).abs().max()) < 1e-10
  k, dt = 3, 1e-3
  us, ps = zip(*[state(i * dt) for i in range(k + 1)])
  du_dt = (1 / dt) * sum(float(c) * v for c, v in zip(ns.bdfk_coeffs(k), us))
  err = sem.velocity.exchange(sem.B(du_dt) + sem.A(us[-1]) - sem.Dt(ps[-1]))
  assert float(err.abs().max()) < 1e-7


def test_fractional_step_identities_and_cg(vortices):
  """navier_stokes_test.py:133-323."""
  from swirl_fem_b200.linalg.cg import cg
  from swirl_fem_b200.navier_stokes import navier_stokes as ns
  sem, state, _ = vortices
  k, dt = 3, 1e-3
  us, ps = zip(*[state(i * dt) for i in range(k + 1)])
  us, u = us[:-1], us[-1]
  ps, p = ps[:-1], ps[-1]
  ext = ns.extk_coeffs(k=1)
  p_ext = sum(float(ext[-i]) * ps[-i] for i in range(1, len(ext) + 1))
  beta = ns.bdfk_coeffs(k)
  f = -(1 / dt) * sum(float(c) * v for c, v in zip(beta[:-1], us))
  b = sem.B(f) + sem.Dt(p_ext)
  beta_k = float(beta[-1])
  H_ = lambda v: (beta_k / dt) * sem.B(v) + sem.A(v)  # noqa: E731
  Q_ = lambda v: (dt / beta_k) * sem.Bi(v)  # noqa: E731
  dp = p - p_ext
  exch = sem.velocity.exchange
  assert float(exch(H_(u) - sem.Dt(dp) - b).abs().max()) < 1e-7
  assert float(exch(H_(u) - H_(Q_(sem.Dt(dp))) - b).abs().max()) < 10 * dt ** 2
  u_star = u - Q_(sem.Dt(dp))
  assert float(exch(H_(u_star) - b).abs().max()) < 10 * dt ** 2
  # :271-323: solve H u* = b, then u* = u - Q D^T dp
  u_cg, _ = cg(H_, b, M=exch, tol=1e-15)
  assert float(exch(H_(u_cg) - b).abs().max()) < 1e-12
  assert float((u_cg - u + Q_(sem.Dt(dp))).abs().max()) < 5 * dt ** 2


def test_stokes_one_step_analytical_and_oracle(vortices):
  """navier_stokes_test.py:325-358, and the same step on the oracle: CG
  iteration counts within +-1, fields to solver tolerance."""
  sem, state, pm = vortices
  k, dt = 3, 1e-3
  us, ps = zip(*[state(i * dt) for i in range(k + 1)])
  u, p, aux = sem.stokes_one_step(us[:-1], ps[:-1], f=0, mu=1, dt=dt,
                                  time_order=k, alpha=0.05,
                                  project_out_nullspace=True, tol=1e-12,
                                  atol=1e-12)
  assert float((u - us[-1]).abs().max()) < 5 * dt ** 2
  assert float((p - ps[-1]).abs().max()) < 50 * dt ** 2
  assert float(aux['u_star_info']['residual']) < 1e-7
  assert float(aux['dp_info']['residual']) < 1e-7
  vmesh, pmesh = helpers.stokes_oracle_meshes(pm, 7)
  osem = dense_ns.StokesSEM(vmesh, pmesh, 7)
  uo, po, auxo = osem.stokes_one_step(
      [v.cpu().numpy() for v in us[:-1]], [q.cpu().numpy() for q in ps[:-1]],
      f=0, mu=1, dt=dt, time_order=k, alpha=0.05, tol=1e-12, atol=1e-12)
  for key in ('u_star_info', 'dp_info'):
    assert abs(aux[key]['num_iterations'] -
               auxo[key]['num_iterations']) <= 1, (
                   key, aux[key]['num_iterations'],
                   auxo[key]['num_iterations'])
  assert np.abs(u.cpu().numpy() - uo).max() < 1e-9
  assert np.abs(p.cpu().numpy() - po).max() < 1e-6


def test_gmsh_file_drives_the_stokes_operators(tmp_path):
  """a `.msh` file (Gmsh cell ordering, $Periodic section)
  read by `common.mesh_reader` feeds the same operators as the generated
  premesh it was written from."""
  from swirl_fem_b200.common import mesh_reader
  pm0 = helpers.stokes_vortices_premesh(3, curved=0.1)
  path = tmp_path / 'vortices.msh'
  helpers.write_premesh_as_msh(path, pm0, version='4.1', tag_stride=2)
  pm = mesh_reader.read(path, ndim=2)
  # the reader does not carry physical groups (mesh_reader.py:108-114): take
  # the boundary facets from the generator
  pm = pm.replace(physical_groups=pm0.physical_groups)
  np.testing.assert_array_equal(pm.elements, pm0.elements)
  sem = _sem(pm, 4)
  vmesh, pmesh = helpers.stokes_oracle_meshes(pm0, 4)
  osem = dense_ns.StokesSEM(vmesh, pmesh, 4)
  rng = np.random.default_rng(11)
  u = rng.standard_normal((vmesh['node_coords'].shape[0], 2))
  p = rng.standard_normal(pmesh['node_coords'].shape[0])
  assert rel_err(sem.A(dev(u)), osem.A(u)) < 1e-12
  assert rel_err(sem.D(dev(u)), osem.D(u)) < 1e-12
  assert rel_err(sem.Dt(dev(p)), osem.Dt(p)) < 1e-12
  assert rel_err(sem.velocity.exchange(dev(u)), osem.v_exchange(u)) < 1e-14


def test_differentiable_solve_symmetric():
  """`custom_linear_solve(symmetric=True)` -- the
  cotangent of b is one more CG solve with the same operator
  (navier_stokes.py:436-452)."""
  from swirl_fem_b200.linalg.cg import cg
  from swirl_fem_b200.linalg.differentiable import custom_linear_solve
  pm = helpers.stokes_vortices_premesh(3, curved=0.1)
  sem = _sem(pm, 4)
  dt, beta_k = 1e-2, 11.0 / 6.0
  H_ = lambda v: (beta_k / dt) * sem.B(v) + sem.A(v)  # noqa: E731
  solve = lambda mv, rhs: cg(mv, rhs, M=sem.velocity.exchange, tol=1e-13)  # noqa: E731
  rng = np.random.default_rng(5)
  n = sem.velocity.mesh.num_nodes
  mask = sem.velocity.interior_mask
  b = (dev(rng.standard_normal((n, 2))) * mask).requires_grad_(True)
  w = dev(rng.standard_normal((n, 2))) * mask
  x, aux = custom_linear_solve(H_, b, solve, symmetric=True, has_aux=True)
  assert aux['num_iterations'] > 0
  loss = (w * x).sum()
  loss.backward()
  # d/db <w, H^-1 b> = H^-T w = H^-1 w
  want, _ = solve(H_, w)
  assert rel_err(b.grad, want.cpu().numpy()) < 1e-9
  with pytest.raises(ValueError, match='transpose_solve'):
    custom_linear_solve(H_, b, solve)


def test_fixed_form_kernels_equal_the_general_form_path():
  """`D`, `D^T`, `C` run as evaluation -> `sfem_pointwise` -> transposed
  evaluation; the reference's form-based formulation through the general
  `local_covector` (navier_stokes.py:238-245, 313-329) gives the same."""
  pm = helpers.stokes_vortices_premesh(3, curved=0.12)
  sem = _sem(pm, 5)
  rng = np.random.default_rng(1)
  n = sem.velocity.mesh.num_nodes
  ul = sem.velocity.gather(dev(rng.standard_normal((n, 2))))
  pl = sem.pressure.gather(dev(rng.standard_normal(
      sem.pressure.pspace.mesh.num_nodes)))
  for fast, general in ((sem.D_local(ul), sem.D_local_general(ul)),
                        (sem.Dt_local(pl), sem.Dt_local_general(pl)),
                        (sem.velocity.C_local(ul),
                         sem.velocity.C_local_general(ul))):
    assert fast.shape == general.shape
    assert rel_err(fast, general.cpu().numpy()) < 1e-13


# (orders 7, 5, 3 in 2-D and 7 in 3-D have compile-time kernel instances -- in
# 2-D the line-per-lane kernels --, the other cases take the runtime shapes)
@pytest.mark.parametrize('ndim,ne,order', [(2, 5, 7), (2, 3, 4), (3, 2, 4),
                                           (3, 2, 6), (2, 4, 5), (2, 4, 3),
                                           (3, 2, 7)])
def test_fused_div_and_gradient_match_composed(ndim, ne, order):
  """`sfem_stokes_div` / `sfem_stokes_grad_t` (one launch each) against the
  composed element-local formulation (evaluation -> pointwise -> transposed
  evaluation, itself checked against the reference goldens above), 2-D and
  3-D, curved elements, and the adjoint identity <D u, p> = <u, D^T p> on
  interior velocities."""
  from swirl_fem_b200.common.premesh_commons import unit_cube_mesh
  pm = unit_cube_mesh(ne, ndim=ndim, a=-1., b=1.)
  x = np.asarray(pm.node_coords, dtype=np.float64)
  x = x + 0.06 * np.sin(np.pi * x[:, np.roll(np.arange(ndim), 1)]) * (1 - x ** 2)
  sem = _sem(pm.replace(node_coords=x), order)
  rng = np.random.default_rng(order)
  nv = sem.velocity.mesh.num_nodes
  npr = sem.pressure.pspace.mesh.num_nodes
  u = dev(rng.standard_normal((nv, ndim)))
  p = dev(rng.standard_normal(npr))
  assert sem._fused_pair()  # pylint: disable=protected-access
  d_f, d_c = sem.D(u), sem.D_composed(u)
  g_f, g_c = sem.Dt(p), sem.Dt_composed(p)
  assert d_f.shape == d_c.shape and g_f.shape == g_c.shape
  assert rel_err(d_f, d_c.cpu().numpy()) < 1e-12
  assert rel_err(g_f, g_c.cpu().numpy()) < 1e-12
  ui = u * sem.velocity.interior_mask
  lhs = float((sem.D(ui) * p).sum())
  rhs = float((ui * sem.Dt(p)).sum())
  assert abs(lhs - rhs) <= 1e-11 * max(abs(lhs), abs(rhs), 1.0)


def test_kolmogorov_steps_match_oracle():
  """`examples/kolmogorov.solve_one_step` (niles/datagen/datagen.py:88-102)
  against the oracle over two steps on the doubly periodic square."""
  from swirl_fem_b200.common.premesh_commons import unit_cube_mesh
  from swirl_fem_b200.examples import kolmogorov
  order, k, dt = 6, 3, 1e-3
  sem = kolmogorov.create_sem(resolution=4, order=order)
  pm = unit_cube_mesh(4, ndim=2, periodic_dims=(0, 1))
  vmesh, pmesh = helpers.stokes_oracle_meshes(pm, order, boundary=None)
  osem = dense_ns.StokesSEM(vmesh, pmesh, order)
  us, ps = kolmogorov.initial_state(sem, k)
  np.testing.assert_allclose(us[0].cpu().numpy(),
                             dense_ns.kolmogorov_u_init(vmesh['node_coords']),
                             atol=1e-14)
  Cus = tuple(map(sem.C, us))
  ous = tuple(u.cpu().numpy() for u in us)
  ops_ = tuple(p.cpu().numpy() for p in ps)
  oCus = tuple(osem.C(u) for u in ous)
  kw = dict(reynolds_number=1000., dt=dt, time_order=k, tol=1e-9, atol=1e-12)
  for _ in range(2):
    u, p, Cu, aux = kolmogorov.solve_one_step(sem, us, ps, Cus, **kw)
    uo, po, Cuo, auxo = dense_ns.navier_stokes_one_step(osem, ous, ops_, oCus,
                                                        **kw)
    for key in ('u_star_info', 'dp_info'):
      assert abs(aux[key]['num_iterations'] -
                 auxo[key]['num_iterations']) <= 1
    assert rel_err(u, uo) < 1e-8 and rel_err(Cu, Cuo) < 1e-8
    assert np.abs(p.cpu().numpy() - po).max() < 1e-6
    us, ps, Cus = us[1:] + (u,), ps[1:] + (p,), Cus[1:] + (Cu,)
    ous, ops_, oCus = ous[1:] + (uo,), ops_[1:] + (po,), oCus[1:] + (Cuo,)
  assert kolmogorov.compute_dx(sem.velocity.mesh) > 0
