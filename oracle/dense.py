"""TEST INFRASTRUCTURE ONLY -- CPU (numpy) restatement of the reference hot path.

This is the parity oracle for the CUDA path: a plain numpy restatement of the
reference's *dense Kronecker* algorithm (the reference is NOT sum-factorised,
swirl_fem/core/interpolation.py:260-261).  Every function cites the reference
file:line it follows (paths relative to /root/reference).  It is pinned against
outputs of the reference's own code run under numpy stubs
(`oracle/ref_harness.py` -> `oracle/make_golden.py` -> `tests/golden/*.npz`)
by `tests/test_oracle_golden.py`.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import this module.  Nothing under
`swirl_fem_b200/` does: the product path fails loudly without its CUDA library.
"""

from __future__ import annotations

import functools

import numpy as np
import scipy.special

SENTINEL = -1  # swirl_fem/core/gather_scatter.py:118


# ----------------------------------------------------------------------------
# 1-D tables: swirl_fem/core/interpolation.py
# ----------------------------------------------------------------------------


def nodes_1d(num_points: int, node_type: str) -> np.ndarray:
  """interpolation.py:51-76 (Nodes1D.create)."""
  if node_type == 'newton_cotes':
    return np.linspace(-1, 1, num=num_points, dtype=np.float64)
  if node_type == 'gauss_legendre':
    x, _ = np.polynomial.legendre.leggauss(deg=num_points)
    return x
  if node_type == 'gauss_lobatto_legendre':
    if num_points == 2:
      inner = np.array([], dtype=np.float64)
    else:
      inner, _ = scipy.special.roots_jacobi(num_points - 2, alpha=1, beta=1)
    return np.concatenate([[-1.], inner, [1.]])
  raise ValueError(node_type)


def weights_1d(num_points: int, node_type: str) -> np.ndarray:
  """interpolation.py:103-118 (Quadrature1D.create_from_nodes_1d)."""
  x = nodes_1d(num_points, node_type)
  if node_type == 'gauss_legendre':
    _, w = np.polynomial.legendre.leggauss(deg=num_points)
    return w
  if node_type == 'gauss_lobatto_legendre':
    return (2 / (num_points * (num_points - 1))) / np.square(
        scipy.special.eval_legendre(num_points - 1, x))
  if node_type == 'newton_cotes':
    return (1 / (num_points - 1)) * np.array(
        [1.] + (num_points - 2) * [2.] + [1.])
  raise ValueError(node_type)


def weights_nd(w: np.ndarray, ndim: int) -> np.ndarray:
  """interpolation.py:138-140."""
  return functools.reduce(np.outer, [w] * ndim).reshape(-1)


def barycentric_weights(num_points: int, node_type: str) -> np.ndarray:
  """interpolation.py:180-208."""
  if node_type == 'newton_cotes':
    order = num_points - 1
    return np.array([np.power(-1, i) * scipy.special.binom(order, i)
                     for i in range(num_points)])
  x = nodes_1d(num_points, node_type)
  w = weights_1d(num_points, node_type)
  if node_type == 'gauss_legendre':
    return np.array([np.power(-1, i) * np.sqrt((1 - np.square(xi)) * wi)
                     for i, (xi, wi) in enumerate(zip(x, w))])
  if node_type == 'gauss_lobatto_legendre':
    return np.array([np.power(-1, i) * np.sqrt(wi) for i, wi in enumerate(w)])
  raise ValueError(node_type)


def interpolation_matrix_1d(grid, grid_type, evalpoints) -> np.ndarray:
  """interpolation.py:210-228: B[q, n] = lagrange_n(eval_q), barycentric."""
  bw = barycentric_weights(len(grid), grid_type)
  out = np.zeros((len(evalpoints), len(grid)))
  for q, x in enumerate(evalpoints):
    for i in range(len(grid)):
      if x == grid[i]:  # exact comparison is intentional (222-223)
        out[q, i] = 1.
        continue
      with np.errstate(divide='ignore'):
        terms = np.array([w / (x - xj) for w, xj in zip(bw, grid)])
      out[q, i] = terms[i] / sum(terms)
  return out


def differentiation_matrix_1d(grid, grid_type) -> np.ndarray:
  """interpolation.py:230-244: D[i, j] = lagrange_j'(grid_i)."""
  bw = barycentric_weights(len(grid), grid_type)
  n = len(grid)
  d = np.zeros((n, n))
  for i in range(n):
    for j in range(n):
      if i != j:
        d[i, j] = (bw[j] / bw[i]) / (grid[i] - grid[j])
  for i in range(n):
    d[i, i] = -d[i, ...].sum()
  return d


class Interp:
  """BarycentricInterpolator restated (interpolation.py:143-292)."""

  def __init__(self, ndim, grid_n, grid_type, eval_n, eval_type):
    self.ndim = ndim
    self.grid_n, self.grid_type = grid_n, grid_type
    self.eval_n, self.eval_type = eval_n, eval_type
    self.grid = nodes_1d(grid_n, grid_type)
    self.evalpoints = nodes_1d(eval_n, eval_type)
    self.b1 = interpolation_matrix_1d(self.grid, grid_type, self.evalpoints)
    self.d1 = differentiation_matrix_1d(self.grid, grid_type)
    self.bd1 = self.b1 @ self.d1  # interpolation.py:273
    # interpolation.py:83-91: equality = same type and number of points
    self.collocated = (grid_type == eval_type and grid_n == eval_n)

  @functools.cached_property
  def matrix(self) -> np.ndarray:
    """interpolation.py:246-252: kron of B, axis 0 slowest."""
    return functools.reduce(np.kron, [self.b1] * self.ndim)

  @functools.cached_property
  def matrix_grad(self) -> np.ndarray:
    """interpolation.py:265-286: (q, n, d); d-th slice has BD at position d."""
    mats = []
    for i in range(self.ndim):
      row = [self.bd1 if i == j else self.b1 for j in range(self.ndim)]
      mats.append(functools.reduce(np.kron, row))
    return np.stack(mats, axis=-1)

  def interpolate(self, x_local: np.ndarray) -> np.ndarray:
    """interpolation.py:254-263, vmapped over elements: (E, n) -> (E, q)."""
    if self.collocated:
      return x_local
    return np.einsum('ij,ej->ei', self.matrix, x_local, optimize=True)

  def interpolate_grad(self, x_local: np.ndarray) -> np.ndarray:
    """interpolation.py:288-292, vmapped over elements: (E, n) -> (E, q, d)."""
    return np.einsum('qnd,en->eqd', self.matrix_grad, x_local,
                     optimize=True)  # same contraction, routed through BLAS


# ----------------------------------------------------------------------------
# gather / scatter / exchange: swirl_fem/core/gather_scatter.py
# ----------------------------------------------------------------------------


def gather(u, indices, fill_value=0.):
  """gather_scatter.py:121-127 (Mesh.gather passes fill_value=0, mesh.py:160)."""
  mask = indices != SENTINEL
  return np.where(mask, u[indices], fill_value)


def scatter(u_local, indices, num_nodes):
  """gather_scatter.py:130-133: zero-init scatter-add of mask*u."""
  mask = indices != SENTINEL
  out = np.zeros(num_nodes, dtype=u_local.dtype)
  np.add.at(out, indices.reshape(-1), (mask * u_local).reshape(-1))
  return out


def exchange(u, gather_indices, unique_indices=None, psum=None):
  """gather_scatter.py:189-261 (QQ^T).  `psum` stands in for lax.psum."""
  if gather_indices is None or not gather_indices.size:
    return u
  mask = gather_indices != SENTINEL
  initial = mask * u[gather_indices]
  if unique_indices is not None:
    num_unique = 1 + unique_indices.max()
    updates = np.zeros(num_unique)
    np.add.at(updates, unique_indices, initial)
  else:
    updates = initial
  if psum is not None:
    updates = psum(updates)
  if unique_indices is not None:
    updates = updates[unique_indices]
  out = np.array(u, copy=True)
  np.add.at(out, gather_indices, mask * (updates - initial))
  return out


# ----------------------------------------------------------------------------
# finite element space: swirl_fem/core/fespace.py
# ----------------------------------------------------------------------------


class FESpace:
  """FiniteElementSpace.create restated (fespace.py:306-348)."""

  def __init__(self, node_coords, elements, grid_n, grid_type, quad_n,
               quad_type, dtype=np.float64):
    """`dtype=np.float32` evaluates the SAME dense algorithm in single
    precision throughout (what the reference does with x64 disabled: the host
    tables are built in float64 and cast, every einsum / inv / det runs in
    float32); used to put the fp32 CUDA path's error next to the reference
    algorithm's own fp32 rounding."""
    self.dtype = np.dtype(dtype)
    self.node_coords = np.asarray(node_coords).astype(self.dtype)
    self.elements = np.asarray(elements)
    self.ndim = self.node_coords.shape[-1]
    self.num_nodes = self.node_coords.shape[0]
    self.interp = Interp(self.ndim, grid_n, grid_type, quad_n, quad_type)
    self.quad_weights = weights_nd(weights_1d(quad_n, quad_type), self.ndim)
    if self.dtype != np.float64:
      self.interp.matrix = self.interp.matrix.astype(self.dtype)
      self.interp.matrix_grad = self.interp.matrix_grad.astype(self.dtype)
      self.quad_weights = self.quad_weights.astype(self.dtype)
    # mesh.py:170-172
    elem_coords = np.stack(
        [gather(self.node_coords[:, k], self.elements)
         for k in range(self.ndim)], axis=-1)
    self.elem_coords = elem_coords
    # fespace.py:332-333
    self.quad_coords = np.stack(
        [self.interp.interpolate(elem_coords[..., k])
         for k in range(self.ndim)], axis=-1)
    # fespace.py:338-340: jacs[m,q,i,j] = d x_j / d xi_i
    self.jacs = np.einsum('mnj,qni->mqij', elem_coords,
                          self.interp.matrix_grad, optimize=True)
    # fespace.py:345-346 (signed det, no abs)
    self.invjacs = np.linalg.inv(self.jacs)
    self.jacdets = np.linalg.det(self.jacs)

  # -- q-function evaluation (fespace.py:178-225) --
  def eval_scalar(self, u_local):
    return self.interp.interpolate(u_local)

  def eval_scalar_grad(self, u_local):
    """fespace.py:190-195."""
    elem_grads = self.interp.interpolate_grad(u_local)
    return np.einsum('mqi,mqji->mqj', elem_grads, self.invjacs)

  def eval_vector(self, u_local):
    """fespace.py:207-209."""
    return np.stack([self.interp.interpolate(u_local[..., k])
                     for k in range(u_local.shape[-1])], axis=-1)

  def eval_vector_grad(self, u_local):
    """fespace.py:221-225: out[m,q,j,k] = d u_k / d x_j."""
    elem_grads = np.stack(
        [self.interp.interpolate_grad(u_local[..., k])
         for k in range(u_local.shape[-1])], axis=-1)
    return np.einsum('mqik,mqji->mqjk', elem_grads, self.invjacs)

  def integrate_values(self, w):
    """fespace.py:401-403."""
    return np.einsum('mq,mq,q->', w, self.jacdets, self.quad_weights)

  # -- local covectors = linear transposes (fespace.py:458-471) --
  def mass_local(self, u_local):
    """Transpose of v -> integrate(u * v), form `l` (examples/poisson.py:133-134)."""
    uq = self.eval_scalar(u_local)
    wq = uq * self.jacdets * self.quad_weights
    if self.interp.collocated:
      return wq
    return np.einsum('qn,eq->en', self.interp.matrix, wq)

  def stiffness_local(self, u_local):
    """Transpose of v -> integrate(grad u . grad v), form `a` (poisson.py:136-137)."""
    gu = self.eval_scalar_grad(u_local)                       # (E,q,d) physical
    wq = gu * (self.jacdets * self.quad_weights)[..., None]   # (E,q,d)
    ref = np.einsum('mqj,mqji->mqi', wq, self.invjacs)        # back to reference
    return np.einsum('qni,mqi->mn', self.interp.matrix_grad, ref,
                     optimize=True)

  def helmholtz_local(self, u_local, lam, mu):
    out = 0.
    if lam != 0:
      out = out + lam * self.mass_local(u_local)
    if mu != 0:
      out = out + mu * self.stiffness_local(u_local)
    return out

  def vector_stiffness_local(self, u_local):
    """navier_stokes.py:220-227 (A_local): componentwise scalar stiffness."""
    return np.stack([self.stiffness_local(u_local[..., k])
                     for k in range(u_local.shape[-1])], axis=-1)

  def vector_mass_local(self, u_local):
    """navier_stokes.py:229-236 (B_local)."""
    return np.stack([self.mass_local(u_local[..., k])
                     for k in range(u_local.shape[-1])], axis=-1)

  # -- global operators (examples/poisson.py:141-157) --
  def gather(self, u):
    return gather(u, self.elements)

  def scatter(self, u_local):
    return scatter(u_local, self.elements, self.num_nodes)

  def apply(self, u, lam=0., mu=1., interior_mask=None):
    """mask * scatter(local_covector(gather(u))) (poisson.py:141-146)."""
    y = self.scatter(self.helmholtz_local(self.gather(u), lam, mu))
    if interior_mask is not None:
      y = y * interior_mask
    return y

  def stiffness_diag(self, interior_mask=None):
    """diag(A) (K12, not in the reference; SURVEY section 2a definition)."""
    g = self.interp.matrix_grad                                 # (q,n,d)
    phys = np.einsum('qni,mqji->mqnj', g, self.invjacs)         # (E,q,n,d)
    w = self.jacdets * self.quad_weights                        # (E,q)
    diag_local = np.einsum('mq,mqnj,mqnj->mn', w, phys, phys)
    d = self.scatter(diag_local)
    if interior_mask is not None:
      d = d * interior_mask
    return d

  def mass_diag(self):
    """Lumped mass = scatter(mass_local(1)) (navier_stokes.py:286-287)."""
    return self.scatter(self.mass_local(np.ones(self.elements.shape)))

  def apply_gemm(self, u, mu=1., threads=1):
    """Same stiffness apply routed through batched GEMMs (cpu baseline leg).
    `threads` > 1: the elements are split into chunks that run concurrently
    (numpy releases the GIL inside matmul / einsum), so the CPU leg uses all
    the host cores it is given, not only the BLAS-threaded GEMMs."""
    g = self.interp.matrix_grad
    q, n, d = g.shape
    gm = np.ascontiguousarray(g.transpose(1, 0, 2).reshape(n, q * d))
    ul = self.gather(u)
    w = self.jacdets * self.quad_weights

    def local(sl):
      eg = (ul[sl] @ gm).reshape(-1, q, d)
      gu = np.einsum('mqi,mqji->mqj', eg, self.invjacs[sl])
      wq = gu * w[sl][..., None]
      ref = np.einsum('mqj,mqji->mqi', wq, self.invjacs[sl])
      return ref.reshape(-1, q * d) @ gm.T

    num = ul.shape[0]
    if threads <= 1 or num < 2 * threads:
      yl = local(slice(0, num))
    else:
      import concurrent.futures  # pylint: disable=g-import-not-at-top
      bounds = np.linspace(0, num, 4 * threads + 1).astype(int)
      with concurrent.futures.ThreadPoolExecutor(threads) as pool:
        parts = list(pool.map(lambda i: local(slice(bounds[i], bounds[i + 1])),
                              range(4 * threads)))
      yl = np.concatenate(parts)
    idx = self.elements.reshape(-1)
    ok = idx != SENTINEL
    y = np.bincount(idx[ok], weights=yl.reshape(-1)[ok],
                    minlength=self.num_nodes)
    return mu * y


# ----------------------------------------------------------------------------
# conjugate gradients: swirl_fem/linalg/cg.py:30-97
# ----------------------------------------------------------------------------


def cg(A, b, x0=None, *, tol=1e-5, atol=0.0, maxiter=None, M=None,
       dot_fn=np.vdot):
  """cg.py:54-97; stopping test on gamma = r.M r (68-73)."""
  if x0 is None:
    x0 = np.zeros_like(b)
  if maxiter is None:
    maxiter = 10 * b.size
  if M is None:
    M = lambda x: x
  bs = dot_fn(b, b)
  atol2 = max(np.square(tol) * bs, np.square(atol))
  r = b - A(x0)
  z = M(r)
  p = z
  gamma = dot_fn(r, z)
  x = x0
  k = 0
  while gamma > atol2 and k < maxiter:
    Ap = A(p)
    alpha = gamma / dot_fn(p, Ap)
    x = x + alpha * p
    r = r - alpha * Ap
    z = M(r)
    gamma_ = dot_fn(r, z)
    beta = gamma_ / gamma
    p = z + beta * p
    gamma = gamma_
    k += 1
  return x, {'residual': gamma, 'num_iterations': k}
