"""Differentiable linear solves.

The reference wraps its CG solves in `lax.custom_linear_solve(matvec, b, solve,
symmetric=True)` (`swirl_fem/navier_stokes/navier_stokes.py:436-452`) so that
reverse-mode differentiation of `x = A^-1 b` does not unroll the iteration: the
cotangent of `b` is obtained by ONE more solve with the transposed operator,
which for a symmetric operator is the same solve on the same kernels.

`custom_linear_solve(matvec, b, solve, symmetric=True)` is the
`torch.autograd` counterpart: forward `x = solve(matvec, b)` (nothing is
recorded on the tape), backward `grad_b = solve(matvec, grad_x)`.  As in the
common use of the reference, the operator itself is treated as constant.
"""

from __future__ import annotations

import torch


class _LinearSolve(torch.autograd.Function):

  @staticmethod
  def forward(ctx, b, matvec, solve, transpose_solve):
    ctx.matvec, ctx.transpose_solve = matvec, transpose_solve
    with torch.no_grad():
      x = solve(matvec, b)
    return x

  @staticmethod
  def backward(ctx, grad_x):
    with torch.no_grad():
      grad_b = ctx.transpose_solve(ctx.matvec, grad_x.contiguous())
    return grad_b, None, None, None


def custom_linear_solve(matvec, b, solve, transpose_solve=None,
                        symmetric: bool = False, has_aux: bool = False):
  """`lax.custom_linear_solve` for CUDA tensors.

  Args:
    matvec: the linear operator `A(x)`.
    b: right-hand side (may require grad).
    solve: `solve(matvec, b)` returning `x` (or `(x, aux)` with `has_aux`).
    transpose_solve: solver for `A^T`; defaults to `solve` when `symmetric`.
    symmetric: `A == A^T`.
    has_aux: `solve` returns `(x, aux)`; `aux` is passed through undifferentiated.
  """
  if transpose_solve is None:
    if not symmetric:
      raise ValueError('a non-symmetric solve needs `transpose_solve`')
    transpose_solve = solve
  aux_box = []

  def fwd(mv, rhs):
    out = solve(mv, rhs)
    if has_aux:
      aux_box.append(out[1])
      return out[0]
    return out

  def bwd(mv, rhs):
    out = transpose_solve(mv, rhs)
    return out[0] if has_aux else out

  x = _LinearSolve.apply(b, matvec, fwd, bwd)
  return (x, aux_box[0]) if has_aux else x
