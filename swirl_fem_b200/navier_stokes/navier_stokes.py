"""Spectral-element Stokes / Navier-Stokes operators on the B200 path.

Same names, signatures and semantics as the reference's
`swirl_fem/navier_stokes/navier_stokes.py` (cited per member, paths relative to
/root/reference); fields are CUDA tensors: velocities `(G_v, d)` AoS, pressures
`(G_p,)`.  How the work is mapped onto the kernels:

  * `A`, `H_ = (beta_k/dt) B + mu A` -- the fused operator kernel (gather,
    sum-factorised stiffness, scatter, Dirichlet mask in one launch; `B` is the
    lumped diagonal, as in the reference).
  * `D`, `D^T`, `C` (convection with over-integration), `vorticity` -- the
    general `local_covector` path: CUDA evaluation kernels
    (`sfem_space_eval`) for values and gradients at the quadrature points, the
    pointwise form, and the transposed evaluation kernel
    (`sfem_space_eval_transpose`).
  * `filter` -- the tensor-product evaluation kernel with the 1-D matrix
    `B_high B_low` (interpolate to N-1 points and back, :460-482).
  * `exchange`, `gather`, `scatter`, `cg` -- the C-ABI kernels of the core
    path.

There is no CPU fallback.
"""

from __future__ import annotations

import ctypes
import dataclasses
import enum
from functools import partial  # pylint: disable=g-importing-member
from typing import Any, Sequence

import numpy as np
import torch

from swirl_fem_b200 import _lib
from swirl_fem_b200.core.fespace import div
from swirl_fem_b200.core.fespace import FiniteElementSpace
from swirl_fem_b200.core.fespace import grad
from swirl_fem_b200.core.interpolation import BarycentricInterpolator
from swirl_fem_b200.core.interpolation import Nodes1D
from swirl_fem_b200.core.interpolation import NodeType
from swirl_fem_b200.core.interpolation import Quadrature1D
from swirl_fem_b200.core.mesh import Mesh
from swirl_fem_b200.core.mesh_refiner import refine_premesh
from swirl_fem_b200.core.premesh import Premesh
from swirl_fem_b200.linalg.cg import cg

# pylint: disable=invalid-name


def extk_coeffs(k: int) -> np.ndarray:
  """Linear extrapolation coefficients of order k (navier_stokes.py:48-57)."""
  gridpoints = Nodes1D.create(num_points=k + 1, node_type=NodeType.NEWTON_COTES)
  h = 2 / k
  evalpoints = Nodes1D.create_single_point(
      node_value=np.array(1 + h, dtype=np.float64))
  interpolator = BarycentricInterpolator(
      ndim=1, gridpoints_1d=gridpoints, evalpoints_1d=evalpoints)
  return interpolator.interpolation_matrix().reshape((-1))


def bdfk_coeffs(k: int) -> np.ndarray:
  """Backward differentiation formula of order k (navier_stokes.py:60-70)."""
  gridpoints = Nodes1D.create(num_points=k + 1, node_type=NodeType.NEWTON_COTES)
  evalpoints = Nodes1D.create_single_point(
      node_value=np.array(1., dtype=np.float64))
  interpolator = BarycentricInterpolator(
      ndim=1, gridpoints_1d=gridpoints, evalpoints_1d=evalpoints)
  h = 2 / k
  return interpolator.interpolation_matrix_grad().reshape((-1)) * h


def _pressure_project_out_nullspace(sem, p):
  """Remove the nullspace (all 1s vector) from p (navier_stokes.py:73-78)."""
  w = sem.pressure.exchange(p)
  q = torch.ones_like(p)
  num = _lib.dot(q, sem.pressure.B(w))
  den = _lib.dot(q, sem.pressure.B(q))
  return w - (num / den).to(p.dtype) * q


@enum.unique
class BCType(enum.Enum):
  """Types of boundary conditions (navier_stokes.py:81-85)."""
  DIRICHLET = 'dirichlet'
  NEUMANN = 'neumann'


def dirichlet_bc(mesh: Mesh, boundary_conditions) -> torch.Tensor:
  """Interior mask from the boundary conditions (navier_stokes.py:88-94)."""
  interior_mask = torch.ones(mesh.num_nodes, dtype=mesh.node_coords.dtype,
                             device=mesh.device)
  for physical_group, (bctype, unused_bcvalue) in boundary_conditions.items():
    if bctype == BCType.DIRICHLET:
      interior_mask = interior_mask * (
          1 - mesh.physical_masks[physical_group].to(interior_mask.dtype))
  return interior_mask


def _vmap_last(fn, u: torch.Tensor) -> torch.Tensor:
  """`vmap(fn, in_axes=-1, out_axes=-1)(u)` for a scalar-field kernel call."""
  return torch.stack([fn(u[..., k].contiguous()) for k in range(u.shape[-1])],
                     dim=-1)


@dataclasses.dataclass
class StokesPressure:
  """Pressure space for the Stokes problem (navier_stokes.py:97-139)."""
  pspace: FiniteElementSpace

  @classmethod
  def create(cls, premesh: Premesh, quadrature: Quadrature1D, order: int,
             device=None, dtype=None) -> 'StokesPressure':
    gridpoints_1d = Nodes1D.create(
        num_points=order - 1, node_type=NodeType.GAUSS_LEGENDRE)
    pmesh = refine_premesh(premesh, gridpoints_1d=gridpoints_1d).finalize(
        device=device, dtype=dtype)
    return cls(pspace=FiniteElementSpace.create(mesh=pmesh,
                                                quadrature=quadrature))

  def gather(self, p):
    return self.pspace.mesh.gather(p)

  def scatter(self, p):
    return self.pspace.mesh.scatter(p)

  def B(self, p):
    """Apply the pressure mass matrix."""
    def l(u, v):
      return lambda x: u(x) * v(x)

    u = self.pspace.scalar_function(self.gather(p))
    v = self.pspace.scalar_function(None)
    return self.scatter(self.pspace.local_covector(l, (u, v)))

  def exchange(self, p):
    """Apply QQ^T."""
    return self.pspace.mesh.exchange(p)


@dataclasses.dataclass
class StokesVelocity:
  """Velocity space for the Stokes system (navier_stokes.py:142-245)."""
  vspace: FiniteElementSpace
  overint_space: FiniteElementSpace
  interior_mask: torch.Tensor   # (G, 1)
  diag_qqt: torch.Tensor        # (G,)
  num_convection_overint_nodes: int = 2

  @classmethod
  def create(cls, premesh: Premesh, order: int, boundary_conditions,
             num_convection_overint_nodes: int = 2, device=None,
             dtype=None) -> 'StokesVelocity':
    gridpoints_1d = Nodes1D.create(
        num_points=order + 1, node_type=NodeType.GAUSS_LOBATTO_LEGENDRE)
    vmesh = refine_premesh(premesh, gridpoints_1d=gridpoints_1d).finalize(
        device=device, dtype=dtype)
    vspace = FiniteElementSpace.create(
        mesh=vmesh,
        quadrature=Quadrature1D.create_from_nodes_1d(gridpoints_1d))
    interior_mask = dirichlet_bc(vmesh, boundary_conditions)[:, None]
    overint_gridpoints_1d = Nodes1D.create(
        num_points=gridpoints_1d.num_points + num_convection_overint_nodes,
        node_type=NodeType.GAUSS_LOBATTO_LEGENDRE)
    overint_space = FiniteElementSpace.create(
        mesh=vmesh,
        quadrature=Quadrature1D.create_from_nodes_1d(overint_gridpoints_1d))
    diag_qqt = vmesh.scatter(torch.ones(
        tuple(vmesh.elements.shape), dtype=vmesh.node_coords.dtype,
        device=vmesh.device))
    return cls(vspace=vspace, overint_space=overint_space, diag_qqt=diag_qqt,
               interior_mask=interior_mask,
               num_convection_overint_nodes=num_convection_overint_nodes)

  @property
  def local_shape(self):
    mesh = self.vspace.mesh
    return (mesh.num_elements, mesh.num_nodes_per_element, mesh.ndim)

  @property
  def mesh(self) -> Mesh:
    return self.vspace.mesh

  def C(self, u):
    """Apply the convection operator with overintegration."""
    return self.interior_mask * self.scatter(self.C_local(self.gather(u)))

  def gather(self, u):
    return _vmap_last(self.vspace.mesh.gather, u)

  def scatter(self, u):
    return _vmap_last(self.vspace.mesh.scatter, u)

  def exchange(self, u):
    """Apply QQ^T."""
    return _vmap_last(self.vspace.mesh.exchange, u)

  def A_local(self, u_local):
    """Apply the velocity stiffness operator locally."""
    def a(u, v):
      return lambda x: (grad(u)(x) * grad(v)(x)).sum((0, 1))

    u = self.vspace.vector_function(u_local)
    v = self.vspace.vector_function(None)
    return self.vspace.local_covector(a, (u, v))

  def B_local(self, u_local):
    """Apply the velocity mass operator locally."""
    def l(u, v):
      return lambda x: (u(x) * v(x)).sum(0)

    u = self.vspace.vector_function(u_local)
    v = self.vspace.vector_function(None)
    return self.vspace.local_covector(l, (u, v))

  def C_local(self, u_local):
    """Apply the local convection operator: u_i d_i w_j v_j.

    Fixed form -> three kernels (values and gradients on the over-integration
    rule, `(u . grad) u` pointwise, transposed evaluation);
    `C_local_general` keeps the reference's form-based formulation."""
    sp = self.overint_space
    d = sp.mesh.ndim
    u_local = u_local.to(sp.dtype).contiguous()
    uq = sp._eval(u_local, ncomp=d, kind=0)    # (E, q, d)
    gq = sp._eval(u_local, ncomp=d, kind=1)    # (E, q, d, d)
    cq = _lib.pointwise(2, d, uq, gq, torch.empty_like(uq))
    return sp._eval_transpose(cq, None, d)

  def C_local_general(self, u_local):
    """`C_local` through `local_covector` (navier_stokes.py:238-245)."""
    def c(u, w, v):
      def f(x):
        ux, gw, vx = u(x), grad(w)(x), v(x)
        d = ux.shape[0]
        return sum(ux[i] * gw[i, j] * vx[j]
                   for i in range(d) for j in range(d))
      return f

    u = self.overint_space.vector_function(u_local)
    v = self.overint_space.vector_function(None)
    return self.overint_space.local_covector(c, (u, u, v))


@dataclasses.dataclass
class StokesSEM:
  """Linear operators of the spectral-element Stokes solver (:248-495)."""

  velocity: StokesVelocity
  pressure: StokesPressure
  velocity_mass_diag: torch.Tensor
  _cache: dict = dataclasses.field(default_factory=dict, repr=False)

  @classmethod
  def create(cls, premesh: Premesh, boundary_conditions, order: int,
             num_convection_overint_nodes: int = 2, device=None,
             dtype=None) -> 'StokesSEM':
    if premesh.order != 1:
      raise ValueError(f'Expected mesh order 1; got {premesh.order}.')
    quadrature = Quadrature1D.create(
        num_points=order + 1, quadrature_type=NodeType.GAUSS_LOBATTO_LEGENDRE)
    pressure = StokesPressure.create(premesh, quadrature, order,
                                     device=device, dtype=dtype)
    velocity = StokesVelocity.create(
        premesh, order, boundary_conditions, num_convection_overint_nodes,
        device=device, dtype=dtype)
    ones = torch.ones(velocity.local_shape, dtype=velocity.vspace.dtype,
                      device=velocity.mesh.device)
    velocity_mass_diag = velocity.scatter(velocity.B_local(ones))
    return cls(velocity=velocity, pressure=pressure,
               velocity_mass_diag=velocity_mass_diag)

  # -- fused velocity operator (hot path) ----------------------------------
  def _velocity_operator(self):
    op = self._cache.get('vop')
    if op is None:
      dirichlet = (self.velocity.interior_mask[:, 0] == 0).to(torch.uint8)
      self._cache['dirichlet'] = dirichlet
      op = self.velocity.vspace.operator(dirichlet_mask=dirichlet,
                                         with_mass=False)
      self._cache['vop'] = op
    return op

  def B(self, u):
    """Apply the mass operator to a velocity field."""
    return self.velocity.interior_mask * self.velocity_mass_diag * u

  def Bi(self, u):
    """Apply the inverse mass operator to a velocity field."""
    diag_qqti = self._cache.get('diag_qqti')
    if diag_qqti is None:  # (the reference re-exchanges it on every call)
      diag_qqti = 1 / self.velocity.exchange(self.velocity_mass_diag)
      self._cache['diag_qqti'] = diag_qqti
    return diag_qqti * self.velocity.exchange(u)

  def A(self, u):
    """Apply the stiffness operator to a velocity field (fused kernel)."""
    return self._velocity_operator().apply(u.contiguous(), lam=0.0, mu=1.0)

  def C(self, u):
    """Apply the convection operator to a velocity field."""
    return self.velocity.C(u)

  def D_local(self, u_local):
    """Apply the local operator D: div(v) tested with the pressure basis.

    Fixed form -> three kernels (velocity gradient at the shared GLL points,
    trace, transposed evaluation on the pressure space); `D_local_general`
    keeps the reference's form-based formulation."""
    vs, ps = self.velocity.vspace, self.pressure.pspace
    d = vs.mesh.ndim
    gq = vs._eval(u_local.to(vs.dtype).contiguous(), ncomp=d, kind=1)
    div = _lib.pointwise(0, d, None, gq, torch.empty(
        gq.shape[:2], dtype=gq.dtype, device=gq.device))
    return ps._eval_transpose(div, None, 1)

  def Dt_local(self, p_local):
    """Apply the local operator D^T (pressure values times the velocity test
    functions' divergence); see `D_local`."""
    vs, ps = self.velocity.vspace, self.pressure.pspace
    d = vs.mesh.ndim
    pq = ps._eval(p_local.to(ps.dtype).contiguous(), ncomp=1, kind=0)
    coeff = _lib.pointwise(1, d, pq, None, torch.empty(
        tuple(pq.shape) + (d, d), dtype=pq.dtype, device=pq.device))
    return vs._eval_transpose(None, coeff, d)

  def D_local_general(self, u_local):
    """`D_local` through `local_covector` (navier_stokes.py:313-320)."""
    def b(v, q):
      return lambda x: div(v)(x) * q(x)

    v = self.velocity.vspace.vector_function(u_local)
    p = self.pressure.pspace.scalar_function(None)
    return self.pressure.pspace.local_covector(b, (v, p))

  def Dt_local_general(self, p_local):
    """`Dt_local` through `local_covector` (navier_stokes.py:322-329)."""
    def b(v, q):
      return lambda x: div(v)(x) * q(x)

    v = self.velocity.vspace.vector_function(None)
    p = self.pressure.pspace.scalar_function(p_local)
    return self.velocity.vspace.local_covector(b, (v, p))

  def D(self, u):
    """Velocity divergence matrix."""
    return self.pressure.scatter(self.D_local(self.velocity.gather(u)))

  def Dt(self, p):
    """Apply the pressure gradient operator."""
    return self.velocity.interior_mask * self.velocity.scatter(
        self.Dt_local(self.pressure.gather(p)))

  def Q(self, u, dt: float, time_order: int):
    """Apply the operator Q = (dt / beta_k) B^-1."""
    beta_k = bdfk_coeffs(time_order)[-1]
    return (dt / beta_k) * self.Bi(u)

  def E(self, p, dt: float, time_order: int):
    """Apply the operator E = D Q D^T."""
    Q_ = partial(self.Q, dt=dt, time_order=time_order)
    return self.D(Q_(self.Dt(p)))

  def stokes_one_step(
      self, us: Sequence[torch.Tensor], ps: Sequence[torch.Tensor], f,
      mu: float, dt: float, time_order: int, alpha: float = 0.05,
      u_boundary: torch.Tensor | None = None, pressure_preconditioner=None,
      project_out_nullspace=True, tol: float = 1e-8, atol: float = 0,
  ) -> tuple[torch.Tensor, torch.Tensor, Any]:
    """One step of the fractional-step scheme (navier_stokes.py:350-458)."""
    if pressure_preconditioner is None and project_out_nullspace:
      pressure_preconditioner = partial(_pressure_project_out_nullspace, self)

    # Extrapolate pressure linearly.
    ext_coeffs = extk_coeffs(k=1)
    p_ext = sum(
        float(ext_coeffs[-i]) * ps[-i] for i in range(1, len(ext_coeffs) + 1))
    f = f + self.Dt(p_ext)

    # Solve for H(u*) = b.
    beta_hist = bdfk_coeffs(time_order)[:-1]
    beta_k = float(bdfk_coeffs(time_order)[-1])
    op = self._velocity_operator()
    mass = self.velocity.interior_mask * self.velocity_mass_diag

    def H_(u):  # (beta_k / dt) B(u) + mu A(u), A by the fused kernel
      return (beta_k / dt) * mass * u + op.apply(u.contiguous(), lam=0.0,
                                                 mu=float(mu))

    f = f - self.B((1 / dt) * sum(float(coef) * u
                                  for coef, u in zip(beta_hist, us)))
    if u_boundary is not None:
      f = f - H_(u_boundary)

    u_star, info = cg(H_, f, M=self.velocity.exchange, tol=tol, atol=atol)
    if u_boundary is not None:
      u_star = u_star + u_boundary
    aux = dict()
    aux['u_star_info'] = info

    # Filter-based stabilization (alpha = 0.05)
    u_star = self.filter(u_star, alpha=alpha)

    # Obtain dp by solving D Q D^T (dp) = -D u*.
    dp, info = cg(partial(self.E, dt=dt, time_order=time_order),
                  -self.D(u_star), M=pressure_preconditioner, tol=tol,
                  atol=atol)
    aux['dp_info'] = info

    u = u_star + self.Q(self.Dt(dp), dt=dt, time_order=time_order)
    p = p_ext + dp
    return u, p, aux

  def _filter_space(self):
    """A space whose 'quadrature points' are the grid nodes and whose 1-D
    interpolation matrix is `B_high B_low`: evaluating on it IS the filter."""
    fs = self._cache.get('filter_space')
    if fs is None:
      mesh = self.velocity.mesh
      grid = mesh.gridpoints_1d
      low = Nodes1D.create(num_points=grid.num_points - 1,
                           node_type=grid.node_type)
      b_low, _ = BarycentricInterpolator(
          ndim=mesh.ndim, gridpoints_1d=grid, evalpoints_1d=low).matrices_1d()
      b_high, _ = BarycentricInterpolator(
          ndim=mesh.ndim, gridpoints_1d=low, evalpoints_1d=grid).matrices_1d()
      f1d = np.ascontiguousarray(b_high @ b_low)
      fs = _TensorMatrixSpace(mesh, f1d)
      self._cache['filter_space'] = fs
    return fs

  def filter(self, u, alpha=0.05):
    """Filter-based stabilization of a velocity field (:460-482)."""
    u_local = self.velocity.gather(u)
    filtered_u_local = self._filter_space().apply(u_local)
    filtered_u = (1 / self.velocity.diag_qqt[:, None]) * (
        self.velocity.scatter(filtered_u_local))
    return (1 - alpha) * u + alpha * filtered_u

  def vorticity(self, u):
    """Vorticity of a (2-D) velocity field (:484-495)."""
    uf = self.velocity.vspace.vector_function(self.velocity.gather(u))

    def _vorticity(x):
      grad_ux = grad(uf)(x)
      return grad_ux[1, 0] - grad_ux[0, 1]

    vort_local = self.velocity.vspace._evaluate(_vorticity)  # pylint: disable=protected-access
    vmesh = self.velocity.vspace.mesh
    return (1. / self.velocity.diag_qqt) * vmesh.scatter(vort_local)


class _TensorMatrixSpace:
  """`sfem_space` whose interpolation matrix is an arbitrary 1-D (N x N)
  matrix M: `apply` computes (M x ... x M) u on every element and component
  with the sum-factorised evaluation kernel (kind 0)."""

  def __init__(self, mesh: Mesh, m1d: np.ndarray):
    n = mesh.gridpoints_1d.num_points
    assert m1d.shape == (n, n)
    self.mesh = mesh
    self.desc = _lib.Desc(
        dim=mesh.ndim, n1d=n, q1d=n, dtype=mesh.node_coords.dtype,
        collocated=False, elements=mesh.elements,
        node_coords=mesh.node_coords, interp_1d=m1d,
        interp_grad_1d=np.zeros_like(m1d), quad_weights_1d=np.ones(n))
    handle = ctypes.c_void_p()
    with torch.cuda.device(mesh.device):
      _lib._check(_lib.lib().sfem_space_create(
          ctypes.byref(self.desc.c), None, None, None, ctypes.byref(handle),
          _lib.stream_ptr(mesh.device)), 'sfem_space_create')
    self.handle = handle

  def apply(self, u_local: torch.Tensor) -> torch.Tensor:
    u_local = u_local.contiguous()
    ncomp = 1 if u_local.dim() == 2 else u_local.shape[-1]
    out = torch.empty_like(u_local)
    with torch.cuda.device(u_local.device):
      _lib._check(_lib.lib().sfem_space_eval(
          self.handle, _lib.ptr(u_local), ncomp, 0, _lib.ptr(out),
          _lib.stream_ptr(u_local.device)), 'sfem_space_eval')
    return out

  def __del__(self):
    h = getattr(self, 'handle', None)
    if h and _lib._lib is not None:
      _lib._lib.sfem_space_destroy(h)
      self.handle = None
