"""Pins the numpy oracle (`oracle/dense.py`) against the reference's outputs.

The goldens in tests/golden/*.npz were produced by the reference's own source
files executed under numpy stubs (oracle/make_golden.py).  CPU-only.
"""

import numpy as np
import pytest

from oracle import dense
from tests.conftest import load_golden

TNAME = {'gll': 'gauss_lobatto_legendre', 'gl': 'gauss_legendre',
         'nc': 'newton_cotes'}


def test_nodes_weights_and_1d_matrices():
  g = load_golden('interpolation')
  for t, full in TNAME.items():
    for n in range(2, 18):
      np.testing.assert_array_equal(dense.nodes_1d(n, full), g[f'nodes_{t}_{n}'])
      np.testing.assert_array_equal(dense.weights_1d(n, full),
                                    g[f'weights_{t}_{n}'])
      np.testing.assert_array_equal(dense.barycentric_weights(n, full),
                                    g[f'bary_{t}_{n}'])
      np.testing.assert_array_equal(
          dense.differentiation_matrix_1d(dense.nodes_1d(n, full), full),
          g[f'D_{t}_{n}'])
  for key in g.files:
    if key.startswith('B_'):
      _, gt, n, et, q = key.split('_')
      b = dense.interpolation_matrix_1d(
          dense.nodes_1d(int(n), TNAME[gt]), TNAME[gt],
          dense.nodes_1d(int(q), TNAME[et]))
      np.testing.assert_array_equal(b, g[key])


def test_kronecker_layout():
  g = load_golden('interpolation')
  for ndim, n, q in ((2, 3, 4), (3, 3, 2), (2, 5, 5)):
    it = dense.Interp(ndim, n, TNAME['gll'], q, TNAME['gl'])
    np.testing.assert_array_equal(it.matrix, g[f'kron_{ndim}_{n}_{q}'])
    np.testing.assert_array_equal(it.matrix_grad, g[f'krongrad_{ndim}_{n}_{q}'])


def _cases():
  g = load_golden('operator')
  return [str(n) for n in g['names']]


@pytest.mark.parametrize('name', _cases())
def test_fespace_and_operator_match_reference(name):
  g = load_golden('operator')
  p = name + '/'
  ndim, _, n, qt, q = [int(v) for v in g[p + 'meta']]
  fes = dense.FESpace(g[p + 'node_coords'], g[p + 'elements'], n, TNAME['gll'],
                      q, TNAME['gl' if qt == 1 else 'gll'])
  tol = dict(rtol=1e-13, atol=1e-13)
  np.testing.assert_allclose(fes.invjacs, g[p + 'invjacs'], **tol)
  np.testing.assert_allclose(fes.jacdets, g[p + 'jacdets'], **tol)
  np.testing.assert_allclose(fes.quad_coords, g[p + 'quad_coords'], **tol)
  u_local = dense.gather(g[p + 'u'], g[p + 'elements'])
  np.testing.assert_array_equal(u_local, g[p + 'u_local'])
  np.testing.assert_allclose(fes.eval_scalar(u_local), g[p + 'eval_u'], **tol)
  np.testing.assert_allclose(fes.eval_scalar_grad(u_local),
                             g[p + 'eval_grad_u'], **tol)
  np.testing.assert_allclose(
      fes.integrate_values(fes.eval_scalar(u_local)), g[p + 'integral_u'],
      **tol)
  np.testing.assert_allclose(
      fes.integrate_values(np.ones(fes.jacdets.shape)), g[p + 'integral_one'],
      **tol)
  np.testing.assert_allclose(fes.mass_local(u_local), g[p + 'mass_local'],
                             **tol)
  np.testing.assert_allclose(fes.stiffness_local(u_local),
                             g[p + 'stiffness_local'], rtol=1e-12, atol=1e-12)
  np.testing.assert_allclose(fes.apply(g[p + 'u']), g[p + 'stiffness'],
                             rtol=1e-12, atol=1e-12)
  np.testing.assert_allclose(fes.apply_gemm(g[p + 'u']), g[p + 'stiffness'],
                             rtol=1e-12, atol=1e-12)
  if p + 'uv' in g.files:
    uv_local = np.stack([dense.gather(g[p + 'uv'][:, k], g[p + 'elements'])
                         for k in range(ndim)], -1)
    np.testing.assert_allclose(fes.eval_vector(uv_local), g[p + 'eval_uv'],
                               **tol)
    np.testing.assert_allclose(fes.eval_vector_grad(uv_local),
                               g[p + 'eval_grad_uv'], **tol)
    np.testing.assert_allclose(fes.vector_stiffness_local(uv_local),
                               g[p + 'vstiffness_local'], rtol=1e-12,
                               atol=1e-12)
    np.testing.assert_allclose(fes.vector_mass_local(uv_local),
                               g[p + 'vmass_local'], **tol)
    div = np.trace(fes.eval_vector_grad(uv_local), axis1=-2, axis2=-1)
    np.testing.assert_allclose(fes.integrate_values(div),
                               g[p + 'integral_div'], **tol)


def test_stiffness_diag_matches_dense_matrix_diagonal():
  g = load_golden('operator')
  p = 'q2_ne2_p3_gl4/'
  fes = dense.FESpace(g[p + 'node_coords'], g[p + 'elements'], 4, TNAME['gll'],
                      4, TNAME['gl'])
  eye = np.eye(fes.num_nodes)
  diag = np.array([fes.apply(eye[i])[i] for i in range(fes.num_nodes)])
  np.testing.assert_allclose(fes.stiffness_diag(), diag, rtol=1e-12, atol=1e-13)


@pytest.mark.parametrize('tag', ['plain', 'jacobi', 'atol', 'maxiter', 'x0'])
def test_cg_matches_reference(tag):
  g = load_golden('cg')
  mat, b = g['mat'], g['b']
  dinv = 1.0 / np.diag(mat)
  kw = {
      'plain': dict(tol=1e-8),
      'jacobi': dict(tol=1e-8, M=lambda r: dinv * r),
      'atol': dict(tol=0., atol=1e-3),
      'maxiter': dict(tol=1e-14, maxiter=7),
      'x0': dict(tol=1e-6, x0=np.ones(len(b))),
  }[tag]
  x, info = dense.cg(lambda v: mat @ v, b, **kw)
  assert info['num_iterations'] == int(g[f'{tag}/num_iterations'])
  np.testing.assert_allclose(x, g[f'{tag}/x'], rtol=1e-10, atol=1e-12)
  np.testing.assert_allclose(info['residual'], g[f'{tag}/residual'], rtol=1e-6)


def test_cg_known_answers():
  # swirl_fem/linalg/cg_test.py:26-50
  g = load_golden('cg')
  b = np.arange(9.0).reshape((3, 3))
  x, info = dense.cg(lambda v: 2 * v, b)
  np.testing.assert_allclose(x, b / 2)
  assert info['num_iterations'] == int(g['kat_2x/num_iterations'])
  x, _ = dense.cg(lambda v: np.array([2 * v[0], 0 * v[1]]), 1 + np.arange(2.0),
                  M=lambda v: np.array([v[0], 0.]))
  np.testing.assert_allclose(x, [0.5, 0.])


def test_exchange_known_answers():
  # swirl_fem/core/gather_scatter_test.py:66-130
  out = dense.exchange(np.array([1., 2., 3.]), np.array([0, 2]),
                       np.array([0, 0]))
  np.testing.assert_allclose(out, [4., 2., 4.])
