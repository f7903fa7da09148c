"""Conjugate gradients with the reference's semantics and signature.

`cg(A, b, x0=None, *, tol, atol, maxiter, M, dot_fn)` follows
`swirl_fem/linalg/cg.py:30-97`: x0 = 0, maxiter = 10*size, the stopping test
is on gamma = r . M r > max(tol^2 b.b, atol^2) (:65-73), and the return value
is `(x, {'residual': gamma, 'num_iterations': k})`.

Two execution paths, both CUDA:
  * fused: `A` is a `BoundOperator` (core/operator.py) and `M` is None or a
    `JacobiPreconditioner` -> one call of the C-ABI `sfem_cg`: operator apply
    with p.Ap in its epilogue, one update kernel, one direction kernel, all
    scalars and the convergence flag on the device.
  * generic: `A`, `M`, `dot_fn` are arbitrary callables on pytrees of CUDA
    tensors (e.g. `M = velocity.exchange`, navier_stokes.py:437) -> host loop
    over the same recurrence using the fused `axpby` / `dot` kernels.
"""

from __future__ import annotations

import ctypes

import torch

from swirl_fem_b200 import _lib
from swirl_fem_b200.core.operator import BoundOperator
from swirl_fem_b200.core.operator import JacobiPreconditioner


# -- minimal pytree helpers (dict / list / tuple / tensor leaves) -------------


def tree_map(fn, *trees):
  t0 = trees[0]
  if isinstance(t0, dict):
    return {k: tree_map(fn, *[t[k] for t in trees]) for k in t0}
  if isinstance(t0, (list, tuple)):
    return type(t0)(tree_map(fn, *[t[i] for t in trees])
                    for i in range(len(t0)))
  return fn(*trees)


def tree_leaves(tree):
  if isinstance(tree, dict):
    return [l for k in tree for l in tree_leaves(tree[k])]
  if isinstance(tree, (list, tuple)):
    return [l for t in tree for l in tree_leaves(t)]
  return [tree]


def _default_dot(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
  return _lib.dot(a, b)


def _vdot(a, b, dot_fn):
  return sum(tree_leaves(tree_map(dot_fn, a, b)))


def _fused_cg(A: BoundOperator, b, x0, tol, atol, maxiter, M, check_every):
  op = A.op
  _lib.require_cuda(b)
  b = b.to(op.dtype).contiguous()
  ncomp = 1 if b.dim() == 1 else b.shape[1]
  x = torch.zeros_like(b) if x0 is None else x0.to(op.dtype).clone().contiguous()
  minv = None
  if M is not None:
    minv = M.minv.to(op.dtype)
    if minv.numel() != b.numel():
      minv = minv.reshape(-1, 1).expand(b.shape[0], ncomp)
    minv = minv.contiguous()
  lib = _lib.lib()
  wbytes = lib.sfem_cg_workspace_bytes(_lib.dtype_code(op.dtype), b.numel())
  work = torch.empty(wbytes, dtype=torch.uint8, device=b.device)
  params = _lib.CgParams(
      tol=float(tol), atol=float(atol),
      maxiter=-1 if maxiter is None else int(maxiter),
      precond=0 if minv is None else 1, check_every=int(check_every),
      lam=A.lam, mu=A.mu)
  info = _lib.CgInfo()
  with torch.cuda.device(b.device):
    _lib._check(lib.sfem_cg(op.handle, _lib.ptr(b), _lib.ptr(x), ncomp,
                            _lib.ptr(minv), ctypes.byref(params),
                            _lib.ptr(work), ctypes.byref(info),
                            _lib.stream_ptr(b.device)), 'sfem_cg')
  residual = torch.tensor(info.residual, dtype=op.dtype, device=b.device)
  return x, {'residual': residual, 'num_iterations': int(info.num_iterations)}


def cg(A, b, x0=None, *, tol=1e-5, atol=0.0, maxiter=None, M=None,
       dot_fn=None, check_every=16):
  """Conjugate gradient solver; see the module docstring."""
  if (isinstance(A, BoundOperator) and isinstance(b, torch.Tensor)
      and (M is None or isinstance(M, JacobiPreconditioner))
      and dot_fn is None):
    return _fused_cg(A, b, x0, tol, atol, maxiter, M, check_every)

  if dot_fn is None:
    dot_fn = _default_dot
  for leaf in tree_leaves(b):
    _lib.require_cuda(leaf)
  if x0 is None:
    x0 = tree_map(torch.zeros_like, b)
  if maxiter is None:
    maxiter = 10 * sum(l.numel() for l in tree_leaves(b))
  if M is None:
    M = lambda v: v

  bs = _vdot(b, b, dot_fn)
  atol2 = max(float(tol) ** 2 * float(bs), float(atol) ** 2)

  def axpy(a, xs, ys):
    """ys + a * xs, out of place, through the axpby kernel."""
    def one(xl, yl):
      out = yl.contiguous().clone()
      _lib.axpby(a, xl.contiguous(), 1.0, out)
      return out
    return tree_map(one, xs, ys)

  r = axpy(-1.0, A(x0), b)
  z = M(r)
  p = z
  gamma = float(_vdot(r, z, dot_fn))
  x = x0
  k = 0
  while gamma > atol2 and k < maxiter:
    Ap = A(p)
    alpha = gamma / float(_vdot(p, Ap, dot_fn))
    x = axpy(alpha, p, x)
    r = axpy(-alpha, Ap, r)
    z = M(r)
    gamma_new = float(_vdot(r, z, dot_fn))
    beta = gamma_new / gamma
    p = axpy(beta, p, z)
    gamma = gamma_new
    k += 1
  leaf0 = tree_leaves(b)[0]
  residual = torch.tensor(gamma, dtype=leaf0.dtype, device=leaf0.device)
  return x, {'residual': residual, 'num_iterations': k}
