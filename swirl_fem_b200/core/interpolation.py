"""1-D node / quadrature tables and barycentric operators (host side, fp64).

Mirrors the public surface of the reference's
`swirl_fem/core/interpolation.py` (NodeType :28-34, Nodes1D :37-91,
Quadrature1D :94-140, BarycentricInterpolator :143-292) so drivers written
against the reference import the same names from here.

Design difference (deliberate): the reference assembles dense Kronecker
matrices `(Q^d x N^d)` / `(Q^d x N^d x d)` and applies them with einsum
(:246-292).  Here only the 1-D matrices `B (QxN)`, `D (NxN)` and `BD = B@D`
are kept; they are the shared-memory operands of the sum-factorised CUDA
kernels (`csrc/`).  The dense matrices are still constructible
(`interpolation_matrix*`) for callers that want them, but nothing on the hot
path uses them.
"""

from __future__ import annotations

import dataclasses
import enum
import functools

import numpy as np
import scipy.special


@enum.unique
class NodeType(enum.Enum):
  """Distributions of collocation / quadrature nodes on [-1, 1]."""
  NEWTON_COTES = 'newton_cotes'
  GAUSS_LEGENDRE = 'gauss_legendre'
  GAUSS_LOBATTO_LEGENDRE = 'gauss_lobatto_legendre'
  SINGLE = 'single_point'


@functools.lru_cache(maxsize=None)
def _node_values(num_points: int, node_type: NodeType) -> np.ndarray:
  if node_type == NodeType.NEWTON_COTES:
    x = np.linspace(-1, 1, num=num_points, dtype=np.float64)
  elif node_type == NodeType.GAUSS_LEGENDRE:
    x, _ = np.polynomial.legendre.leggauss(deg=num_points)
  elif node_type == NodeType.GAUSS_LOBATTO_LEGENDRE:
    # interior GLL nodes = roots of P'_{n-1} = Gauss-Jacobi(1,1) nodes
    if num_points == 2:
      inner = np.array([], dtype=np.float64)
    else:
      inner, _ = scipy.special.roots_jacobi(num_points - 2, alpha=1, beta=1)
    x = np.concatenate([[-1.], inner, [1.]])
  else:
    raise ValueError(f'Node type not recognized: {node_type}')
  x.setflags(write=False)
  return x


@dataclasses.dataclass(frozen=True, eq=False)
class Nodes1D:
  """A sequence of 1-D nodes on the reference element [-1, 1]."""

  num_points: int
  node_type: NodeType
  node_values: np.ndarray

  @classmethod
  def create_single_point(cls, node_value) -> 'Nodes1D':
    return cls(num_points=1, node_type=NodeType.SINGLE,
               node_values=np.array([node_value], dtype=np.float64))

  @classmethod
  def create(cls, num_points: int, node_type: NodeType) -> 'Nodes1D':
    return cls(num_points=num_points, node_type=node_type,
               node_values=_node_values(int(num_points), node_type))

  def is_continuous(self) -> bool:
    """Whether continuity is preserved at element boundaries."""
    return bool(self.node_values[0] == -1.0 and self.node_values[-1] == 1.0)

  def __eq__(self, other):
    # Same rule as the reference (:83-91): type and number of points.
    if not isinstance(other, Nodes1D) or self.node_type != other.node_type:
      return False
    if self.node_type == NodeType.SINGLE:
      return bool(np.allclose(self.node_values, other.node_values, rtol=0,
                              atol=np.finfo(np.float64).eps))
    return self.num_points == other.num_points

  def __hash__(self):
    return hash((self.node_type, self.num_points))


def _quadrature_weights(nodes: Nodes1D) -> np.ndarray:
  n = nodes.num_points
  if nodes.node_type == NodeType.GAUSS_LEGENDRE:
    _, w = np.polynomial.legendre.leggauss(deg=n)
    return w
  if nodes.node_type == NodeType.GAUSS_LOBATTO_LEGENDRE:
    return (2 / (n * (n - 1))) / np.square(
        scipy.special.eval_legendre(n - 1, nodes.node_values))
  if nodes.node_type == NodeType.NEWTON_COTES:
    return (1 / (n - 1)) * np.array([1.] + (n - 2) * [2.] + [1.])
  raise ValueError(f'Quadrature type not recognized: {nodes.node_type}')


@dataclasses.dataclass(frozen=True, eq=False)
class Quadrature1D:
  """A 1-D quadrature rule on [-1, 1]."""

  num_points: int
  quadrature_type: NodeType
  nodes: Nodes1D
  weights: np.ndarray

  @classmethod
  def create_from_nodes_1d(cls, nodes: Nodes1D) -> 'Quadrature1D':
    return cls(num_points=nodes.num_points, quadrature_type=nodes.node_type,
               nodes=nodes, weights=_quadrature_weights(nodes))

  @classmethod
  def create(cls, num_points: int, quadrature_type: NodeType):
    return cls.create_from_nodes_1d(
        Nodes1D.create(num_points=num_points, node_type=quadrature_type))

  def weights_nd(self, ndim: int) -> np.ndarray:
    """Tensor-product weights, axis 0 slowest."""
    return functools.reduce(np.outer, [self.weights] * ndim).reshape(-1)


def _barycentric_weights(grid: Nodes1D) -> np.ndarray:
  n = grid.num_points
  sign = np.where(np.arange(n) % 2 == 0, 1.0, -1.0)
  if grid.node_type == NodeType.NEWTON_COTES:
    return sign * scipy.special.binom(n - 1, np.arange(n))
  if grid.node_type == NodeType.GAUSS_LEGENDRE:
    w = _quadrature_weights(grid)
    return sign * np.sqrt((1 - np.square(grid.node_values)) * w)
  if grid.node_type == NodeType.GAUSS_LOBATTO_LEGENDRE:
    return sign * np.sqrt(_quadrature_weights(grid))
  raise ValueError(f'Gridpoint type not supported: {grid.node_type}')


class BarycentricInterpolator:
  """Barycentric Lagrange interpolation between 1-D node sets.

  Holds the three 1-D operators the kernels need:
    `B[q, n]  = l_n(eval_q)`,  `D[i, j] = l_j'(grid_i)`,  `BD = B @ D`.
  """

  def __init__(self, ndim: int, gridpoints_1d: Nodes1D, evalpoints_1d: Nodes1D):
    self.ndim = ndim
    self.gridpoints_1d = gridpoints_1d
    self.evalpoints_1d = evalpoints_1d

  # -- reference-named private helpers --------------------------------------
  def _barycentric_weights(self) -> np.ndarray:
    return _barycentric_weights(self.gridpoints_1d)

  @functools.cached_property
  def _b1(self) -> np.ndarray:
    bw = self._barycentric_weights()
    x = self.gridpoints_1d.node_values
    e = np.asarray(self.evalpoints_1d.node_values, dtype=np.float64)
    diff = e[:, None] - x[None, :]
    hit = diff == 0.0  # exact comparison is intentional (Berrut-Trefethen s.7)
    with np.errstate(divide='ignore', invalid='ignore'):
      terms = bw[None, :] / diff
      b = terms / terms.sum(axis=1, keepdims=True)
    # rows evaluated exactly at a grid node are unit vectors
    b = np.where(hit.any(axis=1, keepdims=True), hit.astype(np.float64), b)
    return b

  @functools.cached_property
  def _d1(self) -> np.ndarray:
    bw = self._barycentric_weights()
    x = self.gridpoints_1d.node_values
    n = len(x)
    with np.errstate(divide='ignore', invalid='ignore'):
      d = (bw[None, :] / bw[:, None]) / (x[:, None] - x[None, :])
    d[np.arange(n), np.arange(n)] = 0.0
    d[np.arange(n), np.arange(n)] = -d.sum(axis=1)
    return d

  def _interpolation_matrix_1d(self) -> np.ndarray:
    return self._b1

  def _differentiation_matrix_1d(self) -> np.ndarray:
    return self._d1

  # -- public --------------------------------------------------------------
  @property
  def collocated(self) -> bool:
    """True when grid == eval nodes: interpolation is the identity."""
    return self.gridpoints_1d == self.evalpoints_1d

  def matrices_1d(self):
    """Returns `(B, BD)` as C-contiguous fp64 `(Q, N)` arrays."""
    b = np.ascontiguousarray(self._b1, dtype=np.float64)
    bd = np.ascontiguousarray(self._b1 @ self._d1, dtype=np.float64)
    return b, bd

  def interpolation_matrix(self) -> np.ndarray:
    return functools.reduce(np.kron, [self._b1] * self.ndim)

  def interpolation_matrix_grad(self) -> np.ndarray:
    b, bd = self._b1, self._b1 @ self._d1
    mats = [
        functools.reduce(np.kron, [bd if i == j else b
                                   for j in range(self.ndim)])
        for i in range(self.ndim)
    ]
    return np.stack(mats, axis=-1)
