"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's Stokes solver.

Follows `swirl_fem/navier_stokes/navier_stokes.py` (cited per function, paths
relative to /root/reference) on top of `oracle/dense.py` (dense Kronecker
evaluation, as the reference does).  Pinned against the reference's own code
run under numpy stubs (`oracle/make_golden_ns.py` ->
`tests/golden/navier_stokes.npz`) and against the analytical known answers of
`swirl_fem/navier_stokes/navier_stokes_test.py:73-358` by
`tests/test_oracle_golden.py` / `tests/test_navier_stokes_cpu.py`.

Only `tests/` may import this module.
"""

from __future__ import annotations

import numpy as np

from oracle import dense

GLL, GL, NC = 'gauss_lobatto_legendre', 'gauss_legendre', 'newton_cotes'


def bdfk_coeffs(k: int) -> np.ndarray:
  """navier_stokes.py:60-70: derivative at the last of k+1 equispaced points."""
  grid = dense.nodes_1d(k + 1, NC)
  b = dense.interpolation_matrix_1d(grid, NC, np.array([1.0]))
  d = dense.differentiation_matrix_1d(grid, NC)
  return (b @ d).reshape(-1) * (2 / k)


def extk_coeffs(k: int) -> np.ndarray:
  """navier_stokes.py:48-57: extrapolation to one step past the last point."""
  grid = dense.nodes_1d(k + 1, NC)
  return dense.interpolation_matrix_1d(
      grid, NC, np.array([1 + 2 / k])).reshape(-1)


def covector(space: dense.FESpace, vals=None, grads=None) -> np.ndarray:
  """Transpose of the evaluation at the quadrature points (fespace.py:458-471).

  vals: (E, q[, c]) coefficient of the placeholder's value; grads:
  (E, q, d[, c]) coefficient of its physical gradient d/dx_j.
  """
  w = space.jacdets * space.quad_weights                      # (E, q)
  out = 0.
  if vals is not None:
    wv = vals * (w if vals.ndim == 2 else w[..., None])
    out = out + np.einsum('qn,eq...->en...', space.interp.matrix, wv)
  if grads is not None:
    wg = grads * (w[..., None] if grads.ndim == 3 else w[..., None, None])
    ref = np.einsum('mqj...,mqji->mqi...', wg, space.invjacs)
    out = out + np.einsum('qni,mqi...->mn...', space.interp.matrix_grad, ref)
  return out


class StokesSEM:
  """navier_stokes.py:248-482 restated.

  Args:
    vmesh, pmesh: dicts with `node_coords`, `elements` (velocity: GLL order
      `order`; pressure: `order - 1` discontinuous Gauss-Legendre points per
      axis) and, for the velocity mesh, `interior_mask`,
      `exchange_gather_indices`, `exchange_unique_indices`.
    order: velocity order (navier_stokes.py:262-284).
  """

  def __init__(self, vmesh, pmesh, order, num_convection_overint_nodes=2):
    n = order + 1
    self.order = order
    self.vmesh, self.pmesh = vmesh, pmesh
    self.vspace = dense.FESpace(vmesh['node_coords'], vmesh['elements'], n,
                                GLL, n, GLL)
    # navier_stokes.py:174-188: same mesh, GLL rule with more points
    self.overint = dense.FESpace(vmesh['node_coords'], vmesh['elements'], n,
                                 GLL, n + num_convection_overint_nodes, GLL)
    # navier_stokes.py:112-117: pressure on GL nodes, the velocity's GLL rule
    self.pspace = dense.FESpace(pmesh['node_coords'], pmesh['elements'],
                                order - 1, GL, n, GLL)
    self.interior_mask = np.asarray(vmesh['interior_mask'],
                                    dtype=np.float64).reshape(-1, 1)
    self.gi = vmesh.get('exchange_gather_indices')
    self.ui = vmesh.get('exchange_unique_indices')
    # navier_stokes.py:189-190, 286-287
    self.diag_qqt = self.vspace.scatter(
        np.ones(self.vspace.elements.shape))
    ones = np.ones(self.vspace.elements.shape + (self.vspace.ndim,))
    self.velocity_mass_diag = self.v_scatter(
        self.vspace.vector_mass_local(ones))

  # -- velocity plumbing (navier_stokes.py:210-218) --
  def v_gather(self, u):
    return np.stack([self.vspace.gather(u[:, k])
                     for k in range(u.shape[1])], -1)

  def v_scatter(self, u_local):
    return np.stack([self.vspace.scatter(u_local[..., k])
                     for k in range(u_local.shape[-1])], -1)

  def v_exchange(self, u):
    return np.stack([dense.exchange(u[:, k], self.gi, self.ui)
                     for k in range(u.shape[1])], -1)

  # -- operators --
  def B(self, u):
    """navier_stokes.py:295-297."""
    return self.interior_mask * self.velocity_mass_diag * u

  def Bi(self, u):
    """navier_stokes.py:299-302."""
    return (1 / self.v_exchange(self.velocity_mass_diag)) * self.v_exchange(u)

  def A(self, u):
    """navier_stokes.py:304-307."""
    return self.interior_mask * self.v_scatter(
        self.vspace.vector_stiffness_local(self.v_gather(u)))

  def C_local(self, u_local):
    """navier_stokes.py:238-245: c = u_i d_i u_j v_j on the over-integration rule."""
    uq = self.overint.eval_vector(u_local)             # (E, q, d)
    gq = self.overint.eval_vector_grad(u_local)        # (E, q, j, k) = d_j u_k
    c = np.einsum('mqi,mqij->mqj', uq, gq)
    return covector(self.overint, vals=c)

  def C(self, u):
    """navier_stokes.py:201-204."""
    return self.interior_mask * self.v_scatter(self.C_local(self.v_gather(u)))

  def D_local(self, u_local):
    """navier_stokes.py:313-320: div(v) q on the pressure space."""
    gq = self.vspace.eval_vector_grad(u_local)
    return covector(self.pspace, vals=np.einsum('mqjj->mq', gq))

  def Dt_local(self, p_local):
    """navier_stokes.py:322-329: the same form transposed to the velocity."""
    pq = self.pspace.eval_scalar(p_local)              # (E, q)
    d = self.vspace.ndim
    grads = pq[..., None, None] * np.eye(d)            # coefficient of d_j v_k
    return covector(self.vspace, grads=grads)

  def D(self, u):
    """navier_stokes.py:331-333."""
    return self.pspace.scatter(self.D_local(self.v_gather(u)))

  def Dt(self, p):
    """navier_stokes.py:335-338."""
    return self.interior_mask * self.v_scatter(
        self.Dt_local(self.pspace.gather(p)))

  def Q(self, u, dt, time_order):
    """navier_stokes.py:340-343."""
    return (dt / bdfk_coeffs(time_order)[-1]) * self.Bi(u)

  def E(self, p, dt, time_order):
    """navier_stokes.py:345-348."""
    return self.D(self.Q(self.Dt(p), dt, time_order))

  def pressure_B(self, p):
    """navier_stokes.py:126-134."""
    return self.pspace.scatter(self.pspace.mass_local(self.pspace.gather(p)))

  def project_out_nullspace(self, p):
    """navier_stokes.py:73-78 (the pressure mesh has no shared dofs)."""
    q = np.ones_like(p)
    return p - (np.vdot(q, self.pressure_B(p)) /
                np.vdot(q, self.pressure_B(q))) * q

  def filter(self, u, alpha=0.05):
    """navier_stokes.py:460-482."""
    n = self.order + 1
    d = self.vspace.ndim
    low = dense.Interp(d, n, GLL, n - 1, GLL)
    high = dense.Interp(d, n - 1, GLL, n, GLL)
    ul = self.v_gather(u)
    fl = np.stack([high.interpolate(low.interpolate(ul[..., k]))
                   for k in range(d)], -1)
    filtered = (1 / self.diag_qqt[:, None]) * self.v_scatter(fl)
    return (1 - alpha) * u + alpha * filtered

  def vorticity(self, u):
    """navier_stokes.py:484-495."""
    g = self.vspace.eval_vector_grad(self.v_gather(u))
    return (1. / self.diag_qqt) * self.vspace.scatter(g[..., 1, 0] - g[..., 0, 1])

  def stokes_one_step(self, us, ps, f, mu, dt, time_order, alpha=0.05,
                      u_boundary=None, tol=1e-8, atol=0.):
    """navier_stokes.py:350-458 with the default null-space projector."""
    ext = extk_coeffs(1)
    p_ext = sum(ext[-i] * ps[-i] for i in range(1, len(ext) + 1))
    f = f + self.Dt(p_ext)
    beta = bdfk_coeffs(time_order)
    beta_hist, beta_k = beta[:-1], beta[-1]
    H = lambda u: (beta_k / dt) * self.B(u) + mu * self.A(u)  # noqa: E731
    f = f - self.B((1 / dt) * sum(c * u for c, u in zip(beta_hist, us)))
    if u_boundary is not None:
      f = f - H(u_boundary)
    dot = lambda a, b: float(np.vdot(a, b))  # noqa: E731
    u_star, info_u = dense.cg(H, f, M=self.v_exchange, tol=tol, atol=atol,
                              dot_fn=dot)
    if u_boundary is not None:
      u_star = u_star + u_boundary
    u_star = self.filter(u_star, alpha=alpha)
    dp, info_p = dense.cg(lambda p: self.E(p, dt, time_order),
                          -self.D(u_star), M=self.project_out_nullspace,
                          tol=tol, atol=atol, dot_fn=dot)
    u = u_star + self.Q(self.Dt(dp), dt, time_order)
    return u, p_ext + dp, {'u_star_info': info_u, 'dp_info': info_p}


def kolmogorov_forcing(x, u, drag_coeff=0.1):
  """niles/datagen/datagen.py:65-72."""
  f0 = np.sin(2 * np.pi * 4. * x[:, 1])
  return np.stack([f0, np.zeros_like(f0)], -1) - drag_coeff * u


def kolmogorov_u_init(x):
  """niles/datagen/datagen.py:56-62."""
  l = 2.
  return np.stack(
      [np.cos(2 * l * np.pi * x[:, 0]) * np.sin(2 * l * np.pi * x[:, 1]),
       -np.sin(2 * l * np.pi * x[:, 0]) * np.cos(2 * l * np.pi * x[:, 1])], -1)


def navier_stokes_one_step(sem: StokesSEM, us, ps, Cus, reynolds_number, dt,
                           time_order, drag_coeff=0.1, tol=1e-5, atol=1e-4):
  """`_solve_one_step` of niles/datagen/datagen.py:88-102."""
  ext = extk_coeffs(time_order - 1)
  Cu = sum(ext[-i] * Cus[-i] for i in range(1, len(ext) + 1))
  f = kolmogorov_forcing(sem.vmesh['node_coords'], us[-1], drag_coeff)
  f = -Cu + sem.B(f)
  u, p, aux = sem.stokes_one_step(us, ps, f, mu=1 / reynolds_number, dt=dt,
                                  time_order=time_order, tol=tol, atol=atol)
  return u, p, sem.C(u), aux
