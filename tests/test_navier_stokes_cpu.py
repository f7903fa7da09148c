"""CPU pins of the Stokes oracle (`oracle/dense_ns.py`).

1. Against the reference's own code: `tests/golden/navier_stokes.npz` holds the
   outputs of the UNMODIFIED `swirl_fem/navier_stokes/navier_stokes.py` run
   under numpy stubs (`oracle/make_golden_ns.py`).
2. Against the analytical known answers of the reference's tests
   (`swirl_fem/navier_stokes/navier_stokes_test.py:73-358`, same mesh, order and
   tolerances).
"""

import numpy as np
import pytest

from oracle import dense
from oracle import dense_ns
from tests import helpers
from tests.conftest import load_golden


def _golden_sem():
  g = load_golden('navier_stokes')
  order = int(g['order'])
  pm = helpers.stokes_vortices_premesh(int(g['ne']), curved=0.1)
  np.testing.assert_allclose(pm.node_coords, g['premesh_coords'], atol=1e-15)
  vmesh, pmesh = helpers.stokes_oracle_meshes(pm, order)
  return g, dense_ns.StokesSEM(vmesh, pmesh, order), vmesh, pmesh


def test_time_stepping_coefficients_match_reference():
  g = load_golden('navier_stokes')
  for k in (1, 2, 3, 4):
    np.testing.assert_allclose(dense_ns.bdfk_coeffs(k), g[f'bdf{k}'],
                               rtol=1e-13, atol=1e-13)
  for k in (1, 2, 3):
    np.testing.assert_allclose(dense_ns.extk_coeffs(k), g[f'ext{k}'],
                               rtol=1e-13, atol=1e-13)
  # BDF3 (interpolation_test.py:278): [-1/3, 3/2, -3, 11/6]
  np.testing.assert_allclose(dense_ns.bdfk_coeffs(3),
                             [-1 / 3, 1.5, -3, 11 / 6], atol=1e-13)


def test_host_meshes_match_reference():
  g, sem, vmesh, pmesh = _golden_sem()
  assert np.array_equal(vmesh['elements'], g['v_elements'])
  assert np.array_equal(pmesh['elements'], g['p_elements'])
  np.testing.assert_allclose(vmesh['node_coords'], g['v_coords'], atol=1e-14)
  np.testing.assert_allclose(pmesh['node_coords'], g['p_coords'], atol=1e-14)
  np.testing.assert_allclose(sem.interior_mask, g['interior_mask'])
  np.testing.assert_allclose(sem.diag_qqt, g['diag_qqt'])
  np.testing.assert_allclose(sem.velocity_mass_diag, g['velocity_mass_diag'],
                             rtol=1e-13)


@pytest.mark.parametrize('name', ['A', 'B', 'Bi', 'C', 'D', 'Dt', 'Q', 'E',
                                  'filter', 'vorticity', 'pressure_B',
                                  'project', 'A_local', 'D_local', 'Dt_local'])
def test_oracle_operator_matches_reference(name):
  g, sem, _, _ = _golden_sem()
  u, p = g['u'], g['p']
  dt, k = 1e-3, 3
  got = {
      'A': lambda: sem.A(u), 'B': lambda: sem.B(u), 'Bi': lambda: sem.Bi(u),
      'C': lambda: sem.C(u), 'D': lambda: sem.D(u), 'Dt': lambda: sem.Dt(p),
      'Q': lambda: sem.Q(u, dt, k), 'E': lambda: sem.E(p, dt, k),
      'filter': lambda: sem.filter(u, 0.05),
      'vorticity': lambda: sem.vorticity(u),
      'pressure_B': lambda: sem.pressure_B(p),
      'project': lambda: sem.project_out_nullspace(p),
      'A_local': lambda: sem.vspace.vector_stiffness_local(sem.v_gather(u)),
      'D_local': lambda: sem.D_local(sem.v_gather(u)),
      'Dt_local': lambda: sem.Dt_local(sem.pspace.gather(p)),
  }[name]()
  want = g[name]
  assert got.shape == want.shape
  assert np.abs(got - want).max() <= 1e-12 * max(np.abs(want).max(), 1.0)


# -- analytical known answers (navier_stokes_test.py) ---------------------------


@pytest.fixture(scope='module')
def vortices():
  pm = helpers.stokes_vortices_premesh(9)
  vmesh, pmesh = helpers.stokes_oracle_meshes(pm, 7)
  sem = dense_ns.StokesSEM(vmesh, pmesh, 7)
  state = lambda t: helpers.stokes_reference_soln(  # noqa: E731
      vmesh['node_coords'], pmesh['node_coords'], t)
  return sem, state


def test_premesh_counts():
  """navier_stokes_test.py:73-77."""
  pm = helpers.stokes_vortices_premesh(9)
  assert pm.num_elements == 81 and pm.num_nodes == 100


def test_analytical_momentum_and_divergence(vortices):
  """navier_stokes_test.py:79-110: B du/dt + A u - D^T p = 0 and div u = 0."""
  sem, state = vortices
  u, p = state(0.0)
  _, sigma = helpers.stokes_reference_soln_params()
  err = sem.v_exchange(sem.B(sigma * u) + sem.A(u) - sem.Dt(p))
  assert np.abs(err).max() < 1e-7
  assert np.abs(sem.D(u)).max() < 1e-10


def _history(state, k=3, dt=1e-3):
  us, ps = zip(*[state(i * dt) for i in range(k + 1)])
  return us, ps


def test_analytical_bdf(vortices):
  """navier_stokes_test.py:112-131."""
  sem, state = vortices
  k, dt = 3, 1e-3
  us, ps = _history(state, k, dt)
  du_dt = (1 / dt) * sum(c * u for c, u in zip(dense_ns.bdfk_coeffs(k), us))
  err = sem.v_exchange(sem.B(du_dt) + sem.A(us[-1]) - sem.Dt(ps[-1]))
  assert np.abs(err).max() < 1e-7


def test_fractional_step_identities(vortices):
  """navier_stokes_test.py:133-269 (basic, approx, LU factorisation)."""
  sem, state = vortices
  k, dt = 3, 1e-3
  us, ps = _history(state, k, dt)
  us, u = us[:-1], us[-1]
  ps, p = ps[:-1], ps[-1]
  ext = dense_ns.extk_coeffs(1)
  p_ext = sum(ext[-i] * ps[-i] for i in range(1, len(ext) + 1))
  beta = dense_ns.bdfk_coeffs(k)
  f = -(1 / dt) * sum(c * v for c, v in zip(beta[:-1], us))
  b = sem.B(f) + sem.Dt(p_ext)
  H = lambda v: (beta[-1] / dt) * sem.B(v) + sem.A(v)  # noqa: E731
  Q = lambda v: (dt / beta[-1]) * sem.Bi(v)  # noqa: E731
  dp = p - p_ext
  assert np.abs(sem.v_exchange(H(u) - sem.Dt(dp) - b)).max() < 1e-7
  assert np.abs(sem.v_exchange(H(u) - H(Q(sem.Dt(dp))) - b)).max() < 10 * dt ** 2
  u_star = u - Q(sem.Dt(dp))
  assert np.abs(sem.v_exchange(H(u_star) - b)).max() < 10 * dt ** 2


def test_solve_one_step(vortices):
  """navier_stokes_test.py:325-358."""
  sem, state = vortices
  k, dt = 3, 1e-3
  us, ps = _history(state, k, dt)
  u, p, aux = sem.stokes_one_step(us[:-1], ps[:-1], f=0, mu=1, dt=dt,
                                  time_order=k, alpha=0.05, tol=1e-12,
                                  atol=1e-12)
  assert np.abs(u - us[-1]).max() < 5 * dt ** 2
  assert np.abs(p - ps[-1]).max() < 50 * dt ** 2
  assert aux['u_star_info']['residual'] < 1e-7
  assert aux['dp_info']['residual'] < 1e-7


def test_kolmogorov_step_on_the_oracle():
  """The Navier-Stokes step of the reference's data generator
  (niles/datagen/datagen.py:88-102) on the doubly periodic square: the flow
  stays divergence free (to the solver tolerance) and the initial
  Taylor-Green-like state decays slowly at high Reynolds number."""
  from swirl_fem_b200.common.premesh_commons import unit_cube_mesh
  pm = unit_cube_mesh(4, ndim=2, periodic_dims=(0, 1))
  order, k, dt = 6, 3, 1e-3
  vmesh, pmesh = helpers.stokes_oracle_meshes(pm, order, boundary=None)
  sem = dense_ns.StokesSEM(vmesh, pmesh, order)
  u0 = dense_ns.kolmogorov_u_init(vmesh['node_coords'])
  us = (u0,) * k
  ps = (np.zeros(pmesh['node_coords'].shape[0]),) * k
  Cus = tuple(sem.C(u) for u in us)
  for _ in range(2):
    u, p, Cu, aux = dense_ns.navier_stokes_one_step(
        sem, us, ps, Cus, reynolds_number=1000., dt=dt, time_order=k,
        tol=1e-9, atol=1e-12)
    us, ps, Cus = us[1:] + (u,), ps[1:] + (p,), Cus[1:] + (Cu,)
  assert np.isfinite(u).all() and np.isfinite(p).all()
  assert np.abs(sem.D(u)).max() < 1e-6
  assert 0.9 < np.abs(u).max() / np.abs(u0).max() < 1.1
