"""Finite element space on a `Mesh`: q-functions, integration, local covectors.

Public surface mirrors the reference's `swirl_fem/core/fespace.py`:
`QFunction`/`Form` protocols (:36-72), `NodalQFunction` family (:75-225),
`grad` (:233-241), `div` (:244-248), `FiniteElementSpace.{create, integrate,
local_covector, scalar_function, vector_function}` (:256-471).

How it differs (B200-first design, same results):
  * `create` runs the geometric-factor kernel (K11) once and keeps `invjacs`,
    `jacdets`, `quad_coords` as CUDA tensors with the reference's shapes.
  * q-function evaluation (`_evaluate`) and `integrate` are sum-factorised
    CUDA kernels through the C ABI, not dense Kronecker einsums.
  * `local_covector(form, funs)`: the reference obtains the operator action by
    `jax.linear_transpose` of the quadrature integral (:458-471).  A Python
    form cannot be traced into a CUDA kernel, so the form is *classified* by
    probing it pointwise: any bilinear form in `(u, grad u; v, grad v)` with
    constant coefficients is a small matrix C recovered by evaluating the form
    on unit inputs.  Mass (`u v`), stiffness (`grad u . grad v`), Helmholtz
    (`lambda u v + mu grad u . grad v`) and their vector (component-wise)
    versions map onto the fused operator kernels.  Anything else raises
    NotImplementedError -- there is no CPU fallback.
  * point convention for user lambdas: `f(x)` is called ONCE with `x` holding
    all quadrature points component-first, `x[i]` of shape `(E, Q^d)`, and
    returns component-first values `(..., E, Q^d)`.  Pointwise lambdas written
    for the reference (`lambda x: 1 + 5 * x[0]`, `grad(f)(x)[0]`) work as is.
"""

from __future__ import annotations

import ctypes
import dataclasses
import weakref
from typing import Callable

import numpy as np
import torch

from swirl_fem_b200 import _lib
from swirl_fem_b200.core.interpolation import BarycentricInterpolator
from swirl_fem_b200.core.interpolation import Quadrature1D
from swirl_fem_b200.core.mesh import Mesh

QFunction = Callable  # f(x) -> values, see module docstring
Form = Callable       # form(*qfunctions) -> QFunction


# ----------------------------------------------------------------------------
# q-functions
# ----------------------------------------------------------------------------


@dataclasses.dataclass
class NodalQFunction:
  """A nodal function of a `FiniteElementSpace` (element-local values)."""

  fespace: 'FiniteElementSpace'
  value_shape: tuple = ()
  u_local: torch.Tensor | None = None
  # classification probe: (value, gradient) returned instead of an evaluation
  _probe: tuple | None = dataclasses.field(default=None, repr=False)
  # evaluations shared by a function and its gradient while a form is probed
  # several times (general `local_covector` path)
  _memo: dict | None = dataclasses.field(default=None, repr=False)

  def __post_init__(self):
    expected = (self.fespace.num_elements,
                self.fespace.mesh.num_nodes_per_element) + self.value_shape
    if self.u_local is not None and tuple(self.u_local.shape) != expected:
      raise ValueError('shape:', tuple(self.u_local.shape))

  def _evaluate(self) -> torch.Tensor:
    """Values on every element's quadrature points, `(E, Q^d) + shape`."""
    raise NotImplementedError

  def _probe_value(self):
    raise NotImplementedError

  def __call__(self, x):
    del x  # nodal values, not coordinates, define the function
    if self._probe is not None:
      return self._probe_value()
    if self._memo is not None and type(self) in self._memo:
      return self._memo[type(self)]
    out = self._evaluate()
    extra = out.dim() - 2
    out = out.permute(*range(2, 2 + extra), 0, 1) if extra else out
    if self._memo is not None:
      self._memo[type(self)] = out
    return out


class ScalarNodalQFunction(NodalQFunction):

  def __init__(self, fespace, u_local=None, _probe=None, _memo=None):
    super().__init__(fespace, (), u_local, _probe, _memo)

  def _evaluate(self):
    return self.fespace._eval(self.u_local, ncomp=1, kind=0)

  def _probe_value(self):
    return self._probe[0]


class ScalarNodalQFunctionGrad(NodalQFunction):

  def __init__(self, fespace, u_local=None, _probe=None, _memo=None):
    super().__init__(fespace, (), u_local, _probe, _memo)

  def _evaluate(self):
    return self.fespace._eval(self.u_local, ncomp=1, kind=1)

  def _probe_value(self):
    return self._probe[1]


class VectorNodalQFunction(NodalQFunction):

  def __init__(self, fespace, u_local=None, _probe=None, _memo=None):
    super().__init__(fespace, (fespace.mesh.ndim,), u_local, _probe, _memo)

  def _evaluate(self):
    return self.fespace._eval(self.u_local, ncomp=self.fespace.mesh.ndim,
                              kind=0)

  def _probe_value(self):
    return self._probe[0]


class VectorNodalQFunctionGrad(NodalQFunction):

  def __init__(self, fespace, u_local=None, _probe=None, _memo=None):
    super().__init__(fespace, (fespace.mesh.ndim,), u_local, _probe, _memo)

  def _evaluate(self):
    # (E, q, d, d) with [..., j, k] = d u_k / d x_j (fespace.py:224-225)
    return self.fespace._eval(self.u_local, ncomp=self.fespace.mesh.ndim,
                              kind=1)

  def _probe_value(self):
    return self._probe[1]


def _autograd_gradient(f: Callable) -> Callable:
  """Pointwise gradient of an analytic `f(x)` (stands in for `jax.grad`)."""
  def g(x):
    xr = x.detach().clone().requires_grad_(True)
    y = f(xr)
    if not isinstance(y, torch.Tensor):
      return torch.zeros_like(x)
    (gx,) = torch.autograd.grad(y.sum(), xr, allow_unused=True)
    return torch.zeros_like(x) if gx is None else gx
  return g


def grad(f: QFunction) -> QFunction:
  """Gradient of a q-function."""
  if isinstance(f, ScalarNodalQFunction):
    return ScalarNodalQFunctionGrad(f.fespace, f.u_local, f._probe, f._memo)
  if isinstance(f, VectorNodalQFunction):
    return VectorNodalQFunctionGrad(f.fespace, f.u_local, f._probe, f._memo)
  return _autograd_gradient(f)


def div(f: QFunction) -> QFunction:
  """Divergence of a vector-valued q-function: trace of its gradient."""
  def _divf(x):
    g = grad(f)(x)
    if isinstance(g, np.ndarray):
      return np.trace(g)
    return sum(g[i, i] for i in range(g.shape[0]))
  return _divf


# ----------------------------------------------------------------------------
# form classification
# ----------------------------------------------------------------------------


@dataclasses.dataclass(frozen=True)
class _FormClass:
  kind: str      # 'scalar' | 'vector'
  lam: float     # coefficient of the value-value term
  mu: float      # coefficient of the gradient-gradient term
  data_index: int
  dual_index: int


_FORM_CACHE: 'weakref.WeakKeyDictionary' = weakref.WeakKeyDictionary()


def _classify(form: Form, funs, fespace) -> _FormClass:
  """Recovers (lambda, mu) of `lam*u*v + mu*grad u . grad v` by probing."""
  is_dual = [isinstance(f, NodalQFunction) and f.u_local is None for f in funs]
  if sum(is_dual) != 1:
    raise ValueError('Exactly one `QFunction` must be a nodal function and '
                     'have `None` as nodal values')
  dual_index = is_dual.index(True)
  data = [i for i, f in enumerate(funs)
          if isinstance(f, NodalQFunction) and f.u_local is not None]
  if len(funs) != 2 or len(data) != 1:
    raise NotImplementedError(
        'local_covector on the B200 path supports bilinear forms of one data '
        'function and one placeholder (mass / stiffness / Helmholtz family); '
        f'got {len(funs)} q-functions.  No CPU fallback exists.')
  data_index = data[0]
  fu, fv = funs[data_index], funs[dual_index]
  if type(fu) is not type(fv) or fu.fespace is not fv.fespace:
    raise NotImplementedError(
        'mixed-space / mixed-rank forms (e.g. the Stokes divergence D, D^T) '
        'are not on the B200 hot path yet')
  kind = 'vector' if isinstance(fu, VectorNodalQFunction) else 'scalar'
  key = (kind, data_index, dual_index, fespace.mesh.ndim)
  try:
    cached = _FORM_CACHE.get(form, {}).get(key)
  except TypeError:
    cached = None
  if cached is not None:
    return cached

  d = fespace.mesh.ndim
  nv = 1 if kind == 'scalar' else d
  size = nv + nv * d  # value entries then gradient entries [j, k] row-major
  cls = ScalarNodalQFunction if kind == 'scalar' else VectorNodalQFunction

  def make(vec):
    if kind == 'scalar':
      return cls(fespace, None, (np.float64(vec[0]), np.array(vec[1:])))
    return cls(fespace, None, (np.array(vec[:d]),
                               np.array(vec[d:]).reshape(d, d)))

  def evaluate(uvec, vvec, x):
    args = [None, None]
    args[data_index] = make(uvec)
    args[dual_index] = make(vvec)
    return float(np.asarray(form(*args)(x)))

  rng = np.random.default_rng(12345)
  x0 = rng.uniform(-0.7, 0.7, size=d)
  x1 = rng.uniform(-0.7, 0.7, size=d)
  eye = np.eye(size)
  coeff = np.array([[evaluate(eye[a], eye[b], x0) for b in range(size)]
                    for a in range(size)])
  lam, mu = coeff[0, 0], coeff[nv, nv]
  expect = np.diag([lam] * nv + [mu] * (nv * d))
  ur, vr = rng.standard_normal(size), rng.standard_normal(size)
  bilinear = np.isclose(evaluate(ur, vr, x0), ur @ coeff @ vr, rtol=1e-10,
                        atol=1e-12)
  constant = np.isclose(evaluate(ur, vr, x1), ur @ coeff @ vr, rtol=1e-10,
                        atol=1e-12)
  if not (bilinear and constant and np.allclose(coeff, expect, rtol=1e-12,
                                                atol=1e-14)):
    raise NotImplementedError(
        'form is not of the Helmholtz family lam*u*v + mu*grad(u).grad(v) with '
        'constant coefficients; the B200 path has no kernel for it and no CPU '
        f'fallback.  Probed coefficient matrix:\n{coeff}')
  result = _FormClass(kind, float(lam), float(mu), data_index, dual_index)
  try:
    _FORM_CACHE.setdefault(form, {})[key] = result
  except TypeError:
    pass
  return result


# ----------------------------------------------------------------------------
# FiniteElementSpace
# ----------------------------------------------------------------------------


class _SpaceHandle:
  """Owns a `sfem_space*`."""

  def __init__(self, desc: _lib.Desc, invjacs, jacdets, quad_coords):
    self.desc = desc
    self.keep = (invjacs, jacdets, quad_coords)
    handle = ctypes.c_void_p()
    with torch.cuda.device(desc.device):
      _lib._check(_lib.lib().sfem_space_create(
          ctypes.byref(desc.c), _lib.ptr(invjacs), _lib.ptr(jacdets),
          _lib.ptr(quad_coords), ctypes.byref(handle),
          _lib.stream_ptr(desc.device)), 'sfem_space_create')
    self.handle = handle

  def __del__(self):
    h = getattr(self, 'handle', None)
    if h and _lib._lib is not None:
      _lib._lib.sfem_space_destroy(h)
      self.handle = None


@dataclasses.dataclass(frozen=True)
class FiniteElementSpace:
  """Nodal FE space on a mesh with a tensor-product quadrature rule."""

  mesh: Mesh
  quadrature: Quadrature1D
  interpolator: BarycentricInterpolator
  invjacs: torch.Tensor      # (E, Q^d, d, d): invjacs[e,q,j,i] = d xi_i / d x_j
  jacdets: torch.Tensor      # (E, Q^d), signed
  quad_coords: torch.Tensor  # (E, Q^d, d)
  _handle: _SpaceHandle = dataclasses.field(repr=False, compare=False,
                                            default=None)
  _cache: dict = dataclasses.field(default_factory=dict, repr=False,
                                   compare=False)

  @classmethod
  def create(cls, mesh: Mesh, quadrature: Quadrature1D) -> 'FiniteElementSpace':
    interpolator = BarycentricInterpolator(
        ndim=mesh.ndim, gridpoints_1d=mesh.gridpoints_1d,
        evalpoints_1d=quadrature.nodes)
    b, bd = interpolator.matrices_1d()
    dtype = mesh.node_coords.dtype
    desc = _lib.Desc(
        dim=mesh.ndim, n1d=mesh.gridpoints_1d.num_points,
        q1d=quadrature.num_points, dtype=dtype,
        collocated=interpolator.collocated, elements=mesh.elements,
        node_coords=mesh.node_coords, interp_1d=b, interp_grad_1d=bd,
        quad_weights_1d=quadrature.weights)
    d = mesh.ndim
    e = mesh.num_elements
    q = quadrature.num_points ** d
    dev = mesh.device
    invjacs = torch.empty((e, q, d, d), dtype=dtype, device=dev)
    jacdets = torch.empty((e, q), dtype=dtype, device=dev)
    quad_coords = torch.empty((e, q, d), dtype=dtype, device=dev)
    handle = _SpaceHandle(desc, invjacs, jacdets, quad_coords)
    return cls(mesh=mesh, quadrature=quadrature, interpolator=interpolator,
               invjacs=invjacs, jacdets=jacdets, quad_coords=quad_coords,
               _handle=handle)

  # -- sizes ---------------------------------------------------------------
  @property
  def num_elements(self) -> int:
    return self.mesh.num_elements

  @property
  def num_quadrature_points_per_element(self) -> int:
    return int(self.quadrature.num_points ** self.mesh.ndim)

  @property
  def dtype(self) -> torch.dtype:
    return self.jacdets.dtype

  # -- evaluation ----------------------------------------------------------
  def _eval(self, u_local: torch.Tensor, ncomp: int, kind: int):
    """K2-K4 through the C ABI.  Returns the reference's layouts."""
    _lib.require_cuda(u_local)
    u_local = u_local.to(self.dtype).contiguous()
    e, q, d = self.num_elements, self.num_quadrature_points_per_element, (
        self.mesh.ndim)
    shape = (e, q) if kind == 0 else (e, q, d)
    if ncomp > 1 or u_local.dim() == 3:
      shape = shape + (ncomp,)
    out = torch.empty(shape, dtype=self.dtype, device=u_local.device)
    with torch.cuda.device(u_local.device):
      _lib._check(_lib.lib().sfem_space_eval(
          self._handle.handle, _lib.ptr(u_local), ncomp, kind, _lib.ptr(out),
          _lib.stream_ptr(u_local.device)), 'sfem_space_eval')
    return out

  def _evaluate(self, f: QFunction) -> torch.Tensor:
    """Evaluates a q-function on every element's quadrature points."""
    if isinstance(f, NodalQFunction):
      return f._evaluate()
    x = self.quad_coords.permute(2, 0, 1)  # component-first view (d, E, q)
    out = f(x)
    e, q = self.num_elements, self.num_quadrature_points_per_element
    if not isinstance(out, torch.Tensor):
      out = torch.as_tensor(np.asarray(out), dtype=self.dtype,
                            device=self.mesh.device)
    out = out.to(self.dtype)
    if out.dim() < 2:
      out = out.expand(e, q) if out.dim() == 0 else out
    if out.dim() >= 2 and tuple(out.shape[-2:]) == (e, q):
      extra = out.dim() - 2
      if extra:
        out = out.permute(extra, extra + 1, *range(extra))
    return out.contiguous()

  def scalar_function(self, u_local):
    expected = (self.num_elements, self.mesh.num_nodes_per_element)
    if u_local is not None and tuple(u_local.shape) != expected:
      raise ValueError(
          f'Expecting shape {expected} but got u_local.shape='
          f'{tuple(u_local.shape)}')
    return ScalarNodalQFunction(self, u_local)

  def vector_function(self, u_local):
    expected = (self.num_elements, self.mesh.num_nodes_per_element,
                self.mesh.ndim)
    if u_local is not None and tuple(u_local.shape) != expected:
      raise ValueError(
          f'Expecting shape {expected} but got u_local.shape='
          f'{tuple(u_local.shape)}')
    return VectorNodalQFunction(self, u_local)

  def integrate(self, f: QFunction) -> torch.Tensor:
    """sum_{e,q} f[e,q] * detJ[e,q] * W[q]  (0-d tensor of the space dtype)."""
    w = self._evaluate(f)
    expected = (self.num_elements, self.num_quadrature_points_per_element)
    if tuple(w.shape) != expected:
      raise ValueError(
          'Expecting an array of shape (num elements, num quadrature points), '
          f'that is ({expected}) but got: {tuple(w.shape)}')
    out = torch.empty((), dtype=torch.float64, device=w.device)
    with torch.cuda.device(w.device):
      _lib._check(_lib.lib().sfem_space_integrate(
          self._handle.handle, _lib.ptr(w), _lib.ptr(out),
          _lib.stream_ptr(w.device)), 'sfem_space_integrate')
    return out.to(self.dtype)

  # -- operators -----------------------------------------------------------
  def operator(self, dirichlet_mask=None, with_mass: bool = True):
    """The fused global operator handle (see `core/operator.py`)."""
    from swirl_fem_b200.core.operator import FusedOperator  # pylint: disable=g-import-not-at-top
    key = ('op', id(dirichlet_mask) if dirichlet_mask is not None else None,
           bool(with_mass))
    op = self._cache.get(key)
    if op is None:
      op = FusedOperator(self.mesh, self.quadrature,
                         dirichlet_mask=dirichlet_mask, with_mass=with_mass)
      self._cache[key] = op
    return op

  def local_covector(self, form: Form, funs) -> torch.Tensor:
    """Local covector of `v -> integrate(form(..., v, ...))`, shape `(E, n[, d])`.

    Equivalent to the reference's `jax.linear_transpose` of the integral
    (:458-471) for the Helmholtz family of forms; see the module docstring.
    """
    try:
      fc = _classify(form, funs, self)
    except (NotImplementedError, TypeError):
      # not of the Helmholtz family (or written with tensor-only operations)
      return self._general_covector(form, funs)
    u_local = funs[fc.data_index].u_local
    op = self.operator(None, with_mass=True)
    ncomp = 1 if fc.kind == 'scalar' else self.mesh.ndim
    return op.apply_local(u_local, lam=fc.lam, mu=fc.mu, ncomp=ncomp)

  def _eval_transpose(self, vals, grads, ncomp: int) -> torch.Tensor:
    """Transpose of `_eval` through the C ABI (`sfem_space_eval_transpose`)."""
    e, n = self.num_elements, self.mesh.num_nodes_per_element
    dev = self.mesh.device
    shape = (e, n) if ncomp == 1 else (e, n, ncomp)
    out = torch.empty(shape, dtype=self.dtype, device=dev)
    with torch.cuda.device(dev):
      _lib._check(_lib.lib().sfem_space_eval_transpose(
          self._handle.handle, _lib.ptr(vals), _lib.ptr(grads), ncomp,
          _lib.ptr(out), _lib.stream_ptr(dev)), 'sfem_space_eval_transpose')
    return out

  def _general_covector(self, form: Form, funs) -> torch.Tensor:
    """`local_covector` for any form that is linear in its placeholder.

    The reference transposes the quadrature integral with
    `jax.linear_transpose` (:458-471).  Here the form is evaluated on the
    quadrature points with the placeholder's value / gradient set to each unit
    direction in turn (data functions are evaluated once, by the CUDA
    evaluation kernels, and memoised); that yields the pointwise coefficients
    of `v` and `grad v`, which `sfem_space_eval_transpose` contracts with the
    weighted basis functions.  Mixed-space forms (Stokes `D`, `D^T`:
    navier_stokes.py:313-338) and trilinear ones (convection, :238-245) take
    this path; data functions of another space must share this space's
    quadrature rule.
    """
    is_dual = [isinstance(f, NodalQFunction) and f.u_local is None
               for f in funs]
    dual_index = is_dual.index(True)
    dual = funs[dual_index]
    if dual.fespace is not self:
      raise ValueError('the placeholder function must belong to this space')
    e, q, d = self.num_elements, self.num_quadrature_points_per_element, (
        self.mesh.ndim)
    dev, dtype = self.mesh.device, self.dtype
    for f in funs:
      if isinstance(f, NodalQFunction) and f.u_local is not None:
        fq = f.fespace.num_quadrature_points_per_element
        if f.fespace.num_elements != e or fq != q:
          raise ValueError(
              'data functions must live on the same elements and quadrature '
              f'points as the placeholder: got {(f.fespace.num_elements, fq)} '
              f'vs {(e, q)}')
    vector = isinstance(dual, VectorNodalQFunction)
    nv = d if vector else 1
    cls = VectorNodalQFunction if vector else ScalarNodalQFunction
    x = self.quad_coords.permute(2, 0, 1)
    memos = {}
    for f in funs:
      if (isinstance(f, NodalQFunction) and f.u_local is not None
          and id(f) not in memos):
        memos[id(f)] = (f, f._memo)
        f._memo = {}

    def probe(value, gradient):
      args = list(funs)
      args[dual_index] = cls(self, None, (value, gradient))
      out = form(*args)(x)
      if not isinstance(out, torch.Tensor):
        out = torch.as_tensor(float(out), dtype=dtype, device=dev)
      return out.to(dtype).expand(e, q)

    def unit(shape, index):
      t = torch.zeros(shape + (1, 1), dtype=dtype, device=dev)
      if index is not None:
        t[index] = 1.0
      return t

    vshape = (d,) if vector else ()
    gshape = (d, d) if vector else (d,)
    try:
      vals = torch.stack(
          [probe(unit(vshape, (k,) if vector else ()), unit(gshape, None))
           for k in range(nv)], dim=-1)                       # (E, q, nv)
      grads = torch.stack([
          torch.stack(
              [probe(unit(vshape, None),
                     unit(gshape, (j, k) if vector else (j,)))
               for k in range(nv)], dim=-1)
          for j in range(d)], dim=-2)                         # (E, q, d, nv)
    finally:
      for f, old in memos.values():
        f._memo = old
    vals = vals.contiguous() if bool(vals.any()) else None
    grads = grads.contiguous() if bool(grads.any()) else None
    if vals is None and grads is None:
      n = self.mesh.num_nodes_per_element
      return torch.zeros((e, n) + ((d,) if vector else ()), dtype=dtype,
                         device=dev)
    return self._eval_transpose(vals, grads, nv)
