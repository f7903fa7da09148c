"""Finite element Poisson solver on the B200 path.

`solve_poisson(mesh, forcing, boundary_conditions, rtol, atol)` has the
signature and discretisation of the reference's
`swirl_fem/examples/poisson.py:47-164`: Gauss-Legendre quadrature with
`order + (ndim + 1) // 2` points (:112-114), mass form `l` / stiffness form `a`
(:133-137), homogeneous Dirichlet rows zeroed by an interior mask (:119-130),
RHS `b = B(f)` (:149-157), CG on `A u = b` (:162).

Differences: the operator `A` is the fused CUDA kernel (gather + local
operator + scatter + mask in one launch) and the solve is the device-resident
fused CG (`linalg.cg`); an optional Jacobi preconditioner (not in the
reference, hook `cg(..., M=)`) is available through `preconditioner='jacobi'`.
The reference calls `jax.scipy.sparse.linalg.cg`, whose recurrence equals
`swirl_fem.linalg.cg` when M is the identity.
"""

from __future__ import annotations

import enum
from typing import Any, Mapping, Tuple

import numpy as np
import torch

from swirl_fem_b200.core.fespace import FiniteElementSpace
from swirl_fem_b200.core.fespace import grad
from swirl_fem_b200.core.interpolation import NodeType
from swirl_fem_b200.core.interpolation import Quadrature1D
from swirl_fem_b200.core.mesh import Mesh
from swirl_fem_b200.core.operator import JacobiPreconditioner
from swirl_fem_b200.linalg.cg import cg

# pylint: disable=invalid-name


@enum.unique
class BCType(enum.Enum):
  DIRICHLET = 'dirichlet'
  NEUMANN = 'neumann'


def dirichlet_mask(mesh: Mesh, boundary_conditions) -> torch.Tensor:
  """uint8 (num_nodes,): 1 on nodes constrained by a Dirichlet condition."""
  mask = torch.zeros(mesh.num_nodes, dtype=torch.bool, device=mesh.device)
  for group, (bctype, bcvalue) in boundary_conditions.items():
    if not (np.isscalar(bcvalue) and bcvalue == 0):
      raise NotImplementedError('Only scalar-valued, homogeneous boundary '
                                f'conditions are supported; got: {bcvalue}')
    if bctype == BCType.DIRICHLET:
      mask |= mesh.physical_masks[group]
  return mask.to(torch.uint8)


# The forms of poisson.py:133-137, written against this package's q-functions.
def mass_form(u, v):
  return lambda x: u(x) * v(x)


def stiffness_form(u, v):
  return lambda x: np.vdot(grad(u)(x), grad(v)(x))


def poisson_space(mesh: Mesh) -> FiniteElementSpace:
  quadrature = Quadrature1D.create(
      num_points=mesh.order + (mesh.ndim + 1) // 2,
      quadrature_type=NodeType.GAUSS_LEGENDRE)
  return FiniteElementSpace.create(mesh, quadrature)


def solve_poisson(
    mesh: Mesh,
    forcing: Any,
    boundary_conditions: Mapping[str, Tuple[BCType, Any]],
    rtol: float = 1e-5,
    atol: float = 0.,
    preconditioner: str | None = None,
    return_info: bool = False,
):
  """Solves -lap u = f with homogeneous Dirichlet conditions on `mesh`."""
  fespace = poisson_space(mesh)
  dmask = dirichlet_mask(mesh, boundary_conditions)
  op = fespace.operator(dirichlet_mask=dmask, with_mass=True)
  if not isinstance(forcing, torch.Tensor):
    forcing = torch.as_tensor(np.asarray(forcing))
  forcing = forcing.to(device=mesh.device, dtype=fespace.dtype)

  A = op.bind(lam=0.0, mu=1.0)      # stiffness, masked rows
  b = op.apply(forcing, lam=1.0, mu=0.0)  # B(f): mass operator, masked rows
  M = None
  if preconditioner == 'jacobi':
    M = JacobiPreconditioner(op.jacobi_minv(lam=0.0, mu=1.0))
  elif preconditioner is not None:
    raise ValueError(f'unknown preconditioner {preconditioner!r}')
  u, info = cg(A, b, tol=rtol, atol=atol, M=M)
  return (u, info) if return_info else u
