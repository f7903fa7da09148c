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


def _history_interpolator(k: int, at: float) -> BarycentricInterpolator:
  """Lagrange basis of the last k+1 time levels, mapped to the equispaced
  points of [-1, 1] (the newest level sits at +1, the step is 2/k), evaluated
  at the abscissa `at`."""
  levels = Nodes1D.create(num_points=k + 1, node_type=NodeType.NEWTON_COTES)
  target = Nodes1D.create_single_point(
      node_value=np.array(at, dtype=np.float64))
  return BarycentricInterpolator(ndim=1, gridpoints_1d=levels,
                                 evalpoints_1d=target)


def extk_coeffs(k: int) -> np.ndarray:
  """EXT-k weights: the k+1 stored levels extrapolated one step ahead
  (reference `extk_coeffs`, navier_stokes.py:48-57)."""
  step = 2 / k
  return _history_interpolator(k, 1 + step).interpolation_matrix().reshape(-1)


def bdfk_coeffs(k: int) -> np.ndarray:
  """BDF-k weights: d/dt at the newest level in units of 1/dt, oldest level
  first (reference `bdfk_coeffs`, navier_stokes.py:60-70)."""
  step = 2 / k
  slope = _history_interpolator(k, 1.).interpolation_matrix_grad().reshape(-1)
  return slope * step


def _pressure_project_out_nullspace(sem, p):
  """Subtracts the constant mode (mass-weighted mean) from a pressure field;
  the preconditioner of the singular pressure system
  (navier_stokes.py:73-78): `w - (q.B w / q.B q) q` with `q = 1`.

  `B` (the pressure mass matrix) is symmetric, so `q . B w = (B q) . w`: the
  vector `B q` and the scalar `q . B q` do not depend on `w` and are computed
  once per `StokesSEM` (the reference re-assembles `B w` and `B q` on every
  call, i.e. in every iteration of the pressure CG); per call this is one
  device dot product and one fused update."""
  shared = sem.pressure.exchange(p)
  cache = getattr(sem, '_cache', None)
  key = ('nullspace', shared.dtype, shared.device)
  entry = None if cache is None else cache.get(key)
  if entry is None:
    constant = torch.ones_like(shared)
    b_const = sem.pressure.B(constant).contiguous()
    entry = (b_const, _lib.dot(constant, b_const))
    if cache is not None:
      cache[key] = entry
  b_const, norm = entry
  mean = _lib.dot(b_const, shared.contiguous()) / norm
  return shared - mean.to(p.dtype)


@enum.unique
class BCType(enum.Enum):
  """Kinds of boundary condition a physical group can carry (:81-85)."""
  DIRICHLET = 'dirichlet'
  NEUMANN = 'neumann'


def dirichlet_bc(mesh: Mesh, boundary_conditions) -> torch.Tensor:
  """1 on free nodes, 0 on every node of a Dirichlet group (:88-94)."""
  free = torch.ones(mesh.num_nodes, dtype=mesh.node_coords.dtype,
                    device=mesh.device)
  constrained = [name for name, (kind, _) in boundary_conditions.items()
                 if kind == BCType.DIRICHLET]
  for name in constrained:
    free = free * (1 - mesh.physical_masks[name].to(free.dtype))
  return free


# Bilinear / trilinear forms in the convention of `core.fespace` (a form takes
# q-functions and returns the pointwise integrand).
def _MASS_FORM(u, v):
  return lambda x: u(x) * v(x)


def _VECTOR_MASS_FORM(u, v):
  return lambda x: (u(x) * v(x)).sum(0)


def _VECTOR_STIFFNESS_FORM(u, v):
  return lambda x: (grad(u)(x) * grad(v)(x)).sum((0, 1))


def _DIVERGENCE_FORM(v, q):
  return lambda x: div(v)(x) * q(x)


def _CONVECTION_FORM(u, w, v):
  def integrand(x):
    ux, gw, vx = u(x), grad(w)(x), v(x)
    d = ux.shape[0]
    return sum(ux[i] * gw[i, j] * vx[j] for i in range(d) for j in range(d))
  return integrand


def _vmap_last(fn, u: torch.Tensor) -> torch.Tensor:
  """Applies a scalar-field kernel call to every component of an AoS field
  (the reference vmaps over the last axis)."""
  return torch.stack([fn(u[..., k].contiguous()) for k in range(u.shape[-1])],
                     dim=-1)


@dataclasses.dataclass
class StokesPressure:
  """Discontinuous pressure: `order - 1` Gauss-Legendre points per axis,
  integrated with the velocity's quadrature rule (navier_stokes.py:97-139)."""
  pspace: FiniteElementSpace

  @classmethod
  def create(cls, premesh: Premesh, quadrature: Quadrature1D, order: int,
             device=None, dtype=None) -> 'StokesPressure':
    gl_points = Nodes1D.create(num_points=order - 1,
                               node_type=NodeType.GAUSS_LEGENDRE)
    mesh = refine_premesh(premesh, gridpoints_1d=gl_points).finalize(
        device=device, dtype=dtype)
    return cls(pspace=FiniteElementSpace.create(mesh=mesh,
                                                quadrature=quadrature))

  def gather(self, p):
    return self.pspace.mesh.gather(p)

  def scatter(self, p):
    return self.pspace.mesh.scatter(p)

  def B(self, p):
    """Consistent pressure mass matrix (classified as a mass form -> fused
    local operator kernel)."""
    trial = self.pspace.scalar_function(self.gather(p))
    test = self.pspace.scalar_function(None)
    return self.scatter(self.pspace.local_covector(_MASS_FORM, (trial, test)))

  def exchange(self, p):
    """QQ^T on the pressure mesh (no shared dofs unless periodic)."""
    return self.pspace.mesh.exchange(p)


@dataclasses.dataclass
class StokesVelocity:
  """Continuous GLL velocity of order `order`, its collocated rule and the
  over-integration rule of the convection term (navier_stokes.py:142-245)."""
  vspace: FiniteElementSpace
  overint_space: FiniteElementSpace
  interior_mask: torch.Tensor   # (G, 1)
  diag_qqt: torch.Tensor        # (G,)
  num_convection_overint_nodes: int = 2

  @classmethod
  def create(cls, premesh: Premesh, order: int, boundary_conditions,
             num_convection_overint_nodes: int = 2, device=None,
             dtype=None) -> 'StokesVelocity':
    gll = NodeType.GAUSS_LOBATTO_LEGENDRE
    nodes = Nodes1D.create(num_points=order + 1, node_type=gll)
    mesh = refine_premesh(premesh, gridpoints_1d=nodes).finalize(
        device=device, dtype=dtype)

    def space_with(points: Nodes1D) -> FiniteElementSpace:
      return FiniteElementSpace.create(
          mesh=mesh, quadrature=Quadrature1D.create_from_nodes_1d(points))

    finer = Nodes1D.create(
        num_points=nodes.num_points + num_convection_overint_nodes,
        node_type=gll)
    # multiplicity of every node over the elements (weights of the filter and
    # of the vorticity average)
    multiplicity = mesh.scatter(torch.ones(
        tuple(mesh.elements.shape), dtype=mesh.node_coords.dtype,
        device=mesh.device))
    return cls(vspace=space_with(nodes), overint_space=space_with(finer),
               diag_qqt=multiplicity,
               interior_mask=dirichlet_bc(mesh, boundary_conditions)[:, None],
               num_convection_overint_nodes=num_convection_overint_nodes)

  @property
  def local_shape(self):
    mesh = self.vspace.mesh
    return (mesh.num_elements, mesh.num_nodes_per_element, mesh.ndim)

  @property
  def mesh(self) -> Mesh:
    return self.vspace.mesh

  def C(self, u):
    """Assembled, masked convection term `(u . grad) u` (over-integrated)."""
    return self.interior_mask * self.scatter(self.C_local(self.gather(u)))

  # The reference vmaps the scalar gather / scatter / exchange over the
  # component axis (navier_stokes.py:210-218); here all components of the AoS
  # field go through ONE kernel launch (`offset = -1` mode of the C ABI).
  def gather(self, u):
    mesh = self.vspace.mesh
    if tuple(u.shape[:1]) != (mesh.num_nodes,):
      raise ValueError(f'Expected `u` to have shape ({mesh.num_nodes}, d) but '
                       f'got: {tuple(u.shape)}.')
    return _lib.gather(u, mesh.elements, fill_value=0.)

  def scatter(self, u):
    mesh = self.vspace.mesh
    return _lib.scatter(u, mesh.elements, mesh.num_nodes)

  def exchange(self, u):
    """QQ^T of every component (periodic / shared dofs get the sum of their
    copies)."""
    mesh = self.vspace.mesh
    if mesh.axis_name is not None:
      return _vmap_last(mesh.exchange, u)
    if (mesh.exchange_gather_indices is None or
        mesh.exchange_gather_indices.numel() == 0):
      return u
    return _lib.exchange(u, mesh.exchange_gather_indices,
                         mesh.exchange_unique_indices)

  def exchange_(self, u):
    """`exchange` in place (unpartitioned meshes; `u` contiguous, owned by
    the caller): no copy of the field."""
    mesh = self.vspace.mesh
    if mesh.axis_name is not None:
      return self.exchange(u)
    if (mesh.exchange_gather_indices is None or
        mesh.exchange_gather_indices.numel() == 0):
      return u
    return _lib.exchange(u, mesh.exchange_gather_indices,
                         mesh.exchange_unique_indices, inplace=True)

  def _vector_covector(self, form, u_local):
    trial = self.vspace.vector_function(u_local)
    return self.vspace.local_covector(
        form, (trial, self.vspace.vector_function(None)))

  def A_local(self, u_local):
    """Element-local vector stiffness (classified -> fused local kernel)."""
    return self._vector_covector(_VECTOR_STIFFNESS_FORM, u_local)

  def B_local(self, u_local):
    """Element-local vector mass (classified -> fused local kernel)."""
    return self._vector_covector(_VECTOR_MASS_FORM, u_local)

  def C_local(self, u_local):
    """Element-local convection covector of `u_i d_i u_j v_j`.

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
    sp = self.overint_space
    field = sp.vector_function(u_local)
    return sp.local_covector(_CONVECTION_FORM,
                             (field, field, sp.vector_function(None)))


@dataclasses.dataclass
class StokesSEM:
  """The P_N / P_{N-2} spectral-element Stokes operators and the fractional
  step built from them (navier_stokes.py:248-495)."""

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
    velocity = StokesVelocity.create(
        premesh, order, boundary_conditions, num_convection_overint_nodes,
        device=device, dtype=dtype)
    # the pressure is integrated with the velocity's collocated GLL rule
    shared_rule = Quadrature1D.create(
        num_points=order + 1, quadrature_type=NodeType.GAUSS_LOBATTO_LEGENDRE)
    pressure = StokesPressure.create(premesh, shared_rule, order,
                                     device=device, dtype=dtype)
    # lumped (diagonal) velocity mass: rows sums of the consistent matrix
    unit_field = torch.ones(velocity.local_shape, dtype=velocity.vspace.dtype,
                            device=velocity.mesh.device)
    lumped = velocity.scatter(velocity.B_local(unit_field))
    return cls(velocity=velocity, pressure=pressure,
               velocity_mass_diag=lumped)

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
    """Lumped velocity mass with Dirichlet rows removed."""
    return self.velocity.interior_mask * self.velocity_mass_diag * u

  def Bi(self, u):
    """Inverse of the assembled lumped mass applied to an un-assembled field:
    `QQ^T u / QQ^T diag`.  The assembled diagonal is cached (the reference
    exchanges it again on every call)."""
    inverse = self._cache.get('diag_qqti')
    if inverse is None:
      inverse = 1 / self.velocity.exchange(self.velocity_mass_diag)
      self._cache['diag_qqti'] = inverse
    return inverse * self.velocity.exchange(u)

  def A(self, u):
    """Masked vector Laplacian: ONE fused launch (gather, sum-factorised
    stiffness, scatter, Dirichlet rows) for all components."""
    return self._velocity_operator().apply(u.contiguous(), lam=0.0, mu=1.0)

  def C(self, u):
    """Convection term, see `StokesVelocity.C`."""
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
    vs, ps = self.velocity.vspace, self.pressure.pspace
    return ps.local_covector(
        _DIVERGENCE_FORM,
        (vs.vector_function(u_local), ps.scalar_function(None)))

  def Dt_local_general(self, p_local):
    """`Dt_local` through `local_covector` (navier_stokes.py:322-329)."""
    vs, ps = self.velocity.vspace, self.pressure.pspace
    return vs.local_covector(
        _DIVERGENCE_FORM,
        (vs.vector_function(None), ps.scalar_function(p_local)))

  def _fused_pair(self):
    """Handles for the fused `D` / `D^T` kernels, or None when the spaces do
    not fit them (then the composed element-local path is used)."""
    ok = self._cache.get('fused_ddt')
    if ok is None:
      vs = getattr(self.velocity, 'vspace', None)
      ps = getattr(self.pressure, 'pspace', None)
      ok = (getattr(vs, '_handle', None) is not None and
            getattr(ps, '_handle', None) is not None and
            vs.mesh.ndim in (2, 3) and
            vs.mesh.num_nodes_per_element <= 1024 and
            vs.mesh.axis_name is None)
      self._cache['fused_ddt'] = ok
      if ok:
        mask = self.velocity.interior_mask[:, 0].to(vs.dtype).contiguous()
        self._cache['mask_vec'] = mask
    return ok

  def D(self, u):
    """Discrete divergence: velocity (G_v, d) -> pressure covector (G_p,).
    ONE fused launch (`sfem_stokes_div`: gather, gradient, trace with the
    inverse Jacobian, weighted pressure-basis contraction, scatter);
    `D_composed` is the element-local formulation of the reference."""
    if self._fused_pair():
      vs, ps = self.velocity.vspace, self.pressure.pspace
      return _lib.stokes_div(vs._handle.handle, ps._handle.handle,  # pylint: disable=protected-access
                             u.to(vs.dtype), ps.mesh.num_nodes)
    return self.D_composed(u)

  def Dt(self, p):
    """Discrete (weak) pressure gradient, the transpose of `D`, masked: ONE
    fused launch (`sfem_stokes_grad_t`)."""
    if self._fused_pair():
      vs, ps = self.velocity.vspace, self.pressure.pspace
      return _lib.stokes_grad_t(vs._handle.handle, ps._handle.handle,  # pylint: disable=protected-access
                                p.to(vs.dtype), self._cache['mask_vec'],
                                vs.mesh.num_nodes, vs.mesh.ndim)
    return self.Dt_composed(p)

  def D_composed(self, u):
    """`D` as scatter(D_local(gather(u))) (navier_stokes.py:331-333)."""
    return self.pressure.scatter(self.D_local(self.velocity.gather(u)))

  def Dt_composed(self, p):
    """`D^T` as mask * scatter(Dt_local(gather(p))) (navier_stokes.py:335-338)."""
    return self.velocity.interior_mask * self.velocity.scatter(
        self.Dt_local(self.pressure.gather(p)))

  def Q(self, u, dt: float, time_order: int):
    """`(dt / beta_k) B^-1`: the approximate inverse of the Helmholtz operator
    used by the pressure correction."""
    leading = bdfk_coeffs(time_order)[-1]
    scale = float(dt / leading)
    # `scale / QQ^T diag` is cached, and the assembled inverse is the same on
    # every copy of a shared dof, so it commutes with QQ^T: one fused multiply
    # and an in-place exchange instead of clone + exchange + two multiplies
    # (this runs in every iteration of the pressure CG)
    key = ('q_inverse', scale)
    inverse = self._cache.get(key)
    if inverse is None:
      base = self._cache.get('diag_qqti')
      if base is None:
        base = 1 / self.velocity.exchange(self.velocity_mass_diag)
        self._cache['diag_qqti'] = base
      inverse = (scale * base).contiguous()
      self._cache[key] = inverse
    if not isinstance(u, torch.Tensor) or u.shape != inverse.shape:
      return scale * self.Bi(u)
    # (test doubles of the velocity space only have the copying `exchange`)
    exchange = getattr(self.velocity, 'exchange_', self.velocity.exchange)
    return exchange((u * inverse).contiguous())

  def E(self, p, dt: float, time_order: int):
    """Pressure operator `D Q D^T` (symmetric, singular on constants)."""
    return self.D(self.Q(self.Dt(p), dt=dt, time_order=time_order))

  def stokes_one_step(
      self, us: Sequence[torch.Tensor], ps: Sequence[torch.Tensor], f,
      mu: float, dt: float, time_order: int, alpha: float = 0.05,
      u_boundary: torch.Tensor | None = None, pressure_preconditioner=None,
      project_out_nullspace=True, tol: float = 1e-8, atol: float = 0,
  ) -> tuple[torch.Tensor, torch.Tensor, Any]:
    """One step of the fractional-step scheme (navier_stokes.py:350-458).

    1. tentative velocity: `H u* = f + D^T p_ext - B(sum_j beta_j u^{n-j})/dt`
       with `H = (beta_k/dt) B + mu A` (CG, `M = exchange`);
    2. filter `u*`; 3. pressure increment: `E dp = -D u*` (CG with the
    null-space projector); 4. `u = u* + Q D^T dp`, `p = p_ext + dp`.
    Returns `(u, p, {'u_star_info', 'dp_info'})`.
    """
    if pressure_preconditioner is None and project_out_nullspace:
      pressure_preconditioner = partial(_pressure_project_out_nullspace, self)

    # EXT-1 pressure; its gradient joins the right-hand side
    ext = extk_coeffs(k=1)
    p_ext = sum(float(ext[-j]) * ps[-j] for j in range(1, len(ext) + 1))
    rhs = f + self.Dt(p_ext)

    # BDF-k: the stored levels move to the right-hand side
    bdf = bdfk_coeffs(time_order)
    shift = float(bdf[-1]) / dt
    history = sum(float(weight) * level for weight, level in zip(bdf[:-1], us))
    rhs = rhs - self.B((1 / dt) * history)

    fused = self._velocity_operator()
    lumped = self.velocity.interior_mask * self.velocity_mass_diag
    viscosity = float(mu)

    def helmholtz(v):  # shift * B(v) + mu * A(v); A is the fused kernel
      return shift * lumped * v + fused.apply(v.contiguous(), lam=0.0,
                                              mu=viscosity)

    if u_boundary is not None:
      rhs = rhs - helmholtz(u_boundary)
    u_star, velocity_info = cg(helmholtz, rhs, M=self.velocity.exchange,
                               tol=tol, atol=atol)
    if u_boundary is not None:
      u_star = u_star + u_boundary
    u_star = self.filter(u_star, alpha=alpha)

    dp, pressure_info = cg(partial(self.E, dt=dt, time_order=time_order),
                           -self.D(u_star), M=pressure_preconditioner,
                           tol=tol, atol=atol)
    u_new = u_star + self.Q(self.Dt(dp), dt=dt, time_order=time_order)
    return u_new, p_ext + dp, {'u_star_info': velocity_info,
                               'dp_info': pressure_info}

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
