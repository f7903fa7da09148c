"""TEST INFRASTRUCTURE ONLY -- runs the *unmodified* reference under numpy stubs.

The reference (`/root/reference`, google-research/swirl-fem) is pure
Python/JAX and neither `jax` nor `flax` is installed in this image.  Its
host-side code (node/quadrature tables, 1-D matrices, Kronecker matrices, mesh
refiner, exchange index builders) is plain numpy/scipy/networkx, and its
device-side arithmetic is a handful of `einsum`/`inv`/`det`/index ops.  This
module installs *numpy-backed* stand-ins for the few `jax`/`flax`/
`more_itertools` names the reference imports, so that the reference's own
source files execute here and produce golden vectors.

What is stubbed (plumbing, not arithmetic):
  * `jax.numpy` -> numpy (+ `.at[].add/.set`, `precision=` ignored);
  * `jax.vmap`  -> python loop + stack;
  * `jax.custom_batching.custom_vmap`, `NodalQFunction.__call__`
    (swirl_fem/core/fespace.py:121-167): the reference uses a custom-vmap
    trick so that `f(x)` under `vmap(vmap(.))` returns `f._evaluate()`; here
    the loop-vmap keeps the current (element, quad-point) index and
    `__call__` returns `self._evaluate()[e, q]`;
  * `jax.linear_transpose` (fespace.py:471) -> probing the linear functional
    with unit vectors (exact for a linear function, tiny meshes only);
  * `lax.while_loop`, `lax.psum` (python loops), `flax.struct.dataclass`,
    `more_itertools.pairwise/powerset`.

Only `tests/` golden generation (`oracle/make_golden.py`) uses this file, and
only in the build container: `/root/reference` does not exist on the GPU box.
Nothing under `swirl_fem_b200/` imports it.
"""

from __future__ import annotations

import dataclasses
import itertools
import sys
import types

import numpy as np

REFERENCE_ROOT = '/root/reference'

_installed = False


# ----------------------------------------------------------------------------
# numpy-backed array with the `.at[idx].add(v)` / `.set(v)` functional updates
# ----------------------------------------------------------------------------


class _AtIndexer:

  def __init__(self, arr):
    self._arr = arr

  def __getitem__(self, idx):
    return _AtUpdater(self._arr, idx)


class _AtUpdater:

  def __init__(self, arr, idx):
    self._arr = arr
    self._idx = idx

  def add(self, values):
    out = np.array(self._arr, copy=True).view(JArray)
    np.add.at(out, self._idx, np.asarray(values))
    return out

  def set(self, values):
    out = np.array(self._arr, copy=True).view(JArray)
    out[self._idx] = values
    return out


class JArray(np.ndarray):
  """ndarray subclass carrying jax's `.at` property."""

  @property
  def at(self):
    return _AtIndexer(self)


def _wrap(x):
  if isinstance(x, np.ndarray) and not isinstance(x, JArray):
    return x.view(JArray)
  return x


def _wrapping(fn):
  def wrapped(*args, **kwargs):
    kwargs.pop('precision', None)
    return _wrap(fn(*args, **kwargs))
  wrapped.__name__ = getattr(fn, '__name__', 'fn')
  return wrapped


# ----------------------------------------------------------------------------
# loop-vmap with an index context (for the NodalQFunction.__call__ stand-in)
# ----------------------------------------------------------------------------

_VMAP_INDEX_STACK: list[int] = []


def _tree_map(fn, *trees):
  t0 = trees[0]
  if isinstance(t0, dict):
    return {k: _tree_map(fn, *[t[k] for t in trees]) for k in t0}
  if isinstance(t0, (list, tuple)):
    out = [_tree_map(fn, *[t[i] for t in trees]) for i in range(len(t0))]
    return type(t0)(out) if not hasattr(t0, '_fields') else type(t0)(*out)
  if t0 is None:
    return None
  return fn(*trees)


def _tree_leaves(tree):
  if isinstance(tree, dict):
    return [l for k in tree for l in _tree_leaves(tree[k])]
  if isinstance(tree, (list, tuple)):
    return [l for t in tree for l in _tree_leaves(t)]
  if tree is None:
    return []
  return [tree]


def vmap(fn, in_axes=0, out_axes=0, axis_name=None):
  del axis_name

  def mapped(*args, **kwargs):
    axes = in_axes if isinstance(in_axes, (tuple, list)) else (
        [in_axes] * len(args))
    size = None
    for a, ax in zip(args, axes):
      if ax is not None:
        size = _tree_leaves(a)[0].shape[ax]
        break
    if size is None:
      for v in kwargs.values():
        size = _tree_leaves(v)[0].shape[0]
        break
    outs = []
    for i in range(size):
      sliced = [
          a if ax is None else _tree_map(
              lambda leaf, ax=ax: _wrap(np.take(np.asarray(leaf), i,
                                                axis=ax)), a)
          for a, ax in zip(args, axes)
      ]
      skw = {k: _tree_map(lambda leaf: np.asarray(leaf)[i], v)
             for k, v in kwargs.items()}
      _VMAP_INDEX_STACK.append(i)
      try:
        outs.append(fn(*sliced, **skw))
      finally:
        _VMAP_INDEX_STACK.pop()
    return _tree_map(
        lambda *leaves: _wrap(np.stack([np.asarray(l) for l in leaves],
                                       axis=out_axes)), *outs)

  return mapped


class _CustomVmap:
  """Inert stand-in; the harness replaces NodalQFunction.__call__ instead."""

  def __init__(self, fn):
    self._fn = fn

  def def_vmap(self, rule):
    return rule

  def __call__(self, *args):
    return self._fn(*args)


def _linear_transpose(fn, *primals):
  """Transpose of a *scalar-valued linear* `fn` by probing with unit vectors."""
  (primal,) = primals

  def transposed(ct):
    shape = primal.shape
    out = np.zeros(shape, dtype=np.float64)
    flat = out.reshape(-1)
    for i in range(flat.size):
      e = np.zeros(flat.size, dtype=np.float64)
      e[i] = 1.0
      flat[i] = float(fn(_wrap(e.reshape(shape)))) * float(ct)
    return (_wrap(out),)

  return transposed


def _while_loop(cond_fun, body_fun, init):
  val = init
  while bool(cond_fun(val)):
    val = body_fun(val)
  return val


def _flax_dataclass(cls=None, **kwargs):
  del kwargs

  def deco(c):
    c = dataclasses.dataclass(c, eq=False) if '__eq__' in c.__dict__ else (
        dataclasses.dataclass(c))
    c.replace = lambda self, **kw: dataclasses.replace(self, **kw)
    return c

  return deco if cls is None else deco(cls)


def _flax_field(pytree_node=True, **kwargs):
  del pytree_node
  return dataclasses.field(**kwargs)


def _pairwise(iterable):
  a, b = itertools.tee(iterable)
  next(b, None)
  return zip(a, b)


def _powerset(iterable):
  s = list(iterable)
  return itertools.chain.from_iterable(
      itertools.combinations(s, r) for r in range(len(s) + 1))


def install():
  """Installs the stub modules and puts the reference on sys.path."""
  global _installed
  if _installed:
    return
  import os
  if not os.path.isdir(REFERENCE_ROOT):
    raise RuntimeError(
        f'{REFERENCE_ROOT} is not present: the reference harness only runs in '
        'the build container (golden generation), never on the GPU box.')

  jnp = types.ModuleType('jax.numpy')
  for name in dir(np):
    if name.startswith('_'):
      continue
    obj = getattr(np, name)
    if callable(obj) and not isinstance(obj, type):
      setattr(jnp, name, _wrapping(obj))
    else:
      setattr(jnp, name, obj)
  jnp.ndarray = np.ndarray
  jnp.array = _wrapping(np.array)
  jnp.asarray = _wrapping(np.asarray)
  linalg = types.ModuleType('jax.numpy.linalg')
  for name in ('inv', 'det', 'norm', 'solve'):
    setattr(linalg, name, _wrapping(getattr(np.linalg, name)))
  jnp.linalg = linalg
  jnp.result_type = np.result_type

  lax = types.ModuleType('jax.lax')
  lax.Precision = types.SimpleNamespace(HIGHEST='highest')
  lax.while_loop = _while_loop
  lax.psum = lambda x, axis_name: x  # single-partition harness only

  def _custom_linear_solve(matvec, b, solve, transpose_solve=None,
                           symmetric=False, has_aux=False):
    del transpose_solve, symmetric, has_aux
    return solve(matvec, b)

  lax.custom_linear_solve = _custom_linear_solve

  typing_mod = types.ModuleType('jax.typing')
  typing_mod.ArrayLike = object

  tree_mod = types.ModuleType('jax.tree')
  tree_mod.map = _tree_map
  tree_util = types.ModuleType('jax.tree_util')
  tree_util.tree_map = _tree_map
  tree_util.tree_leaves = _tree_leaves

  custom_batching = types.ModuleType('jax.custom_batching')
  custom_batching.custom_vmap = _CustomVmap

  core = types.ModuleType('jax.core')

  @dataclasses.dataclass
  class ShapedArray:
    shape: tuple
    dtype: object

  core.ShapedArray = ShapedArray

  jax = types.ModuleType('jax')
  jax.Array = np.ndarray
  jax.numpy = jnp
  jax.lax = lax
  jax.typing = typing_mod
  jax.tree = tree_mod
  jax.tree_util = tree_util
  jax.custom_batching = custom_batching
  jax.core = core
  jax.vmap = vmap
  jax.linear_transpose = _linear_transpose
  jax.jit = lambda f, **kw: f

  flax = types.ModuleType('flax')
  struct = types.ModuleType('flax.struct')
  struct.dataclass = _flax_dataclass
  struct.field = _flax_field
  struct.PyTreeNode = object
  flax.struct = struct

  more_it = types.ModuleType('more_itertools')
  more_it.pairwise = _pairwise
  more_it.powerset = _powerset

  sys.modules.update({
      'jax': jax, 'jax.numpy': jnp, 'jax.lax': lax, 'jax.typing': typing_mod,
      'jax.tree': tree_mod, 'jax.tree_util': tree_util,
      'jax.custom_batching': custom_batching, 'jax.core': core,
      'flax': flax, 'flax.struct': struct, 'more_itertools': more_it,
  })
  if REFERENCE_ROOT not in sys.path:
    sys.path.insert(0, REFERENCE_ROOT)
  _installed = True

  # Stand-in for the custom-vmap plumbing of NodalQFunction.__call__
  # (swirl_fem/core/fespace.py:121-167): inside vmap(vmap(f))(quad_coords) the
  # call returns the whole-mesh evaluation at the current (element, point).
  from swirl_fem.core import fespace as ref_fespace  # pylint: disable=g-import-not-at-top

  def _call(self, x):
    del x
    cache = self.__dict__.get('_harness_eval')
    if cache is None:
      cache = np.asarray(self._evaluate())
      self.__dict__['_harness_eval'] = cache
    e, q = _VMAP_INDEX_STACK[-2], _VMAP_INDEX_STACK[-1]
    return _wrap(cache[e, q])

  ref_fespace.NodalQFunction.__call__ = _call
  for sub in (ref_fespace.ScalarNodalQFunction,
              ref_fespace.ScalarNodalQFunctionGrad,
              ref_fespace.VectorNodalQFunction,
              ref_fespace.VectorNodalQFunctionGrad):
    sub.__call__ = _call


def reference():
  """Returns a namespace with the reference modules imported under the stubs."""
  install()
  # pylint: disable=g-import-not-at-top
  from swirl_fem.common import facet_util
  from swirl_fem.common import premesh_commons
  from swirl_fem.core import fespace
  from swirl_fem.core import gather_scatter
  from swirl_fem.core import interpolation
  from swirl_fem.core import mesh
  from swirl_fem.core import mesh_refiner
  from swirl_fem.core import premesh
  from swirl_fem.linalg import cg
  return types.SimpleNamespace(
      facet_util=facet_util, premesh_commons=premesh_commons, fespace=fespace,
      gather_scatter=gather_scatter, interpolation=interpolation, mesh=mesh,
      mesh_refiner=mesh_refiner, premesh=premesh, cg=cg, jnp=sys.modules[
          'jax.numpy'])
