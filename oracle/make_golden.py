"""Generates tests/golden/*.npz by running the reference's own code.

TEST INFRASTRUCTURE ONLY.  Run in the build container (where /root/reference
exists):  `python -m oracle.make_golden`.  The reference source files execute
unmodified under the numpy-backed jax/flax stubs of `oracle/ref_harness.py`;
the outputs are committed as small fixtures so that the CPU and GPU test
suites never need /root/reference at run time.
"""

from __future__ import annotations

import functools
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_harness  # pylint: disable=g-import-not-at-top

GOLDEN_DIR = os.path.join(ROOT, 'tests', 'golden')

R = ref_harness.reference()
jnp = R.jnp
I = R.interpolation
NodeType = I.NodeType

# The reference rebuilds the dense Kronecker matrices on every interpolate*
# call (interpolation.py:262, 291).  Memoise per interpolator instance: same
# arithmetic, tractable golden generation.
for _name in ('interpolation_matrix', 'interpolation_matrix_grad'):
  _orig = getattr(I.BarycentricInterpolator, _name)

  def _memo(self, _orig=_orig, _name=_name):
    key = '_memo_' + _name
    if key not in self.__dict__:
      self.__dict__[key] = _orig(self)
    return self.__dict__[key]

  setattr(I.BarycentricInterpolator, _name, _memo)

TYPES = {
    'gll': NodeType.GAUSS_LOBATTO_LEGENDRE,
    'gl': NodeType.GAUSS_LEGENDRE,
    'nc': NodeType.NEWTON_COTES,
}


def save(name, **arrays):
  os.makedirs(GOLDEN_DIR, exist_ok=True)
  path = os.path.join(GOLDEN_DIR, name + '.npz')
  np.savez_compressed(path, **{k: np.asarray(v) for k, v in arrays.items()})
  print('wrote', path, len(arrays), 'arrays')


# ----------------------------------------------------------------------------
# 1. interpolation tables
# ----------------------------------------------------------------------------


def golden_interpolation():
  out = {}
  with np.errstate(divide='ignore'):
    for tname, t in TYPES.items():
      for n in range(2, 18):
        nodes = I.Nodes1D.create(n, t)
        quad = I.Quadrature1D.create_from_nodes_1d(nodes)
        out[f'nodes_{tname}_{n}'] = nodes.node_values
        out[f'weights_{tname}_{n}'] = quad.weights
        interp = I.BarycentricInterpolator(1, nodes, nodes)
        out[f'bary_{tname}_{n}'] = interp._barycentric_weights()  # pylint: disable=protected-access
        out[f'D_{tname}_{n}'] = interp._differentiation_matrix_1d()  # pylint: disable=protected-access
    # grid -> eval pairs used by the drivers: GLL(N)->GL(Q), GLL(N)->GLL(Q),
    # NC(2)->GLL/GL (refiner), NC(N)->GL(Q) (fespace_test).
    pairs = []
    for n in range(2, 17):
      for q in (n - 1, n, n + 1, n + 2):
        if q >= 1:
          pairs += [('gll', n, 'gl', q)]
        if q >= 2:
          pairs += [('gll', n, 'gll', q)]
    for q in range(1, 18):
      pairs += [('nc', 2, 'gl', q)]
      if q >= 2:
        pairs += [('nc', 2, 'gll', q)]
    for n in range(2, 6):
      pairs += [('nc', n, 'gl', n), ('nc', n, 'gl', n + 1)]
    for gt, n, et, q in sorted(set(pairs)):
      interp = I.BarycentricInterpolator(
          1, I.Nodes1D.create(n, TYPES[gt]), I.Nodes1D.create(q, TYPES[et]))
      out[f'B_{gt}_{n}_{et}_{q}'] = interp._interpolation_matrix_1d()  # pylint: disable=protected-access
    # Kronecker layouts (axis 0 slowest), small cases.
    for ndim, n, q in ((2, 3, 4), (3, 3, 2), (2, 5, 5)):
      interp = I.BarycentricInterpolator(
          ndim, I.Nodes1D.create(n, TYPES['gll']),
          I.Nodes1D.create(q, TYPES['gl']))
      out[f'kron_{ndim}_{n}_{q}'] = interp.interpolation_matrix()
      out[f'krongrad_{ndim}_{n}_{q}'] = interp.interpolation_matrix_grad()
  save('interpolation', **out)


# ----------------------------------------------------------------------------
# 2. connectivity: unit_cube_mesh -> refine_premesh -> finalize
# ----------------------------------------------------------------------------


def _shuffled_premesh(premesh, seed):
  """Permutes element order and re-orients each element (same geometry).

  Exercises `get_orderings_mapping` (facet_util.py:95-143): each element's
  2^d vertices are re-listed under a random axis permutation / flips.
  """
  rng = np.random.default_rng(seed)
  ndim = premesh.node_coords.shape[-1]
  elements = np.array(premesh.elements)
  perm = rng.permutation(len(elements))
  elements = elements[perm]
  import itertools
  out = []
  for el in elements:
    nd = el.reshape([2] * ndim)
    axes = rng.permutation(ndim)
    nd = nd.transpose(axes)
    flips = [ax for ax in range(ndim) if rng.integers(2)]
    nd = np.flip(nd, flips) if flips else nd
    out.append(nd.reshape(-1))
  del itertools
  return R.premesh.Premesh.create(
      node_coords=premesh.node_coords, elements=np.array(out, dtype=np.int32),
      physical_groups=premesh.physical_groups,
      periodic_links=premesh.periodic_links), perm


def _mesh_arrays(prefix, premesh, refined, mesh):
  out = {
      prefix + 'pre_node_coords': premesh.node_coords,
      prefix + 'pre_elements': premesh.elements,
      prefix + 'elements': refined.elements,
      prefix + 'node_coords': refined.node_coords,
      prefix + 'node_indices': mesh.node_indices,
      prefix + 'exchange_gather_indices': mesh.exchange_gather_indices,
      prefix + 'exchange_unique_indices': mesh.exchange_unique_indices,
  }
  for k, v in premesh.physical_groups.items():
    out[prefix + 'pre_group_' + k] = v
  if premesh.periodic_links is not None:
    out[prefix + 'pre_periodic_links'] = premesh.periodic_links
  for k, v in refined.physical_groups.items():
    out[prefix + 'group_' + k] = v
  if refined.periodic_links is not None:
    out[prefix + 'periodic_links'] = refined.periodic_links
  for k, v in mesh.physical_masks.items():
    out[prefix + 'mask_' + k] = v
  return out


def golden_connectivity():
  cases = [
      # name, ndim, ne, N, type, periodic_dims, shuffle_seed
      ('q2_ne3_p4_yper', 2, 3, 5, 'gll', (1,), None),
      ('q2_ne4_p3', 2, 4, 4, 'gll', (), None),
      ('q2_ne2_p8', 2, 2, 9, 'gll', (), None),
      ('q2_ne3_p2_allper', 2, 3, 3, 'gll', (0, 1), None),
      ('q2_ne4_p1', 2, 4, 2, 'gll', (), None),
      ('q2_ne3_p3_nc', 2, 3, 4, 'nc', (), None),
      ('q2_ne3_gl3', 2, 3, 3, 'gl', (), None),
      ('q2_ne4_p4_shuffled', 2, 4, 5, 'gll', (), 7),
      ('h3_ne2_p3', 3, 2, 4, 'gll', (), None),
      ('h3_ne3_p2_xper', 3, 3, 3, 'gll', (0,), None),
      ('h3_ne2_p7', 3, 2, 8, 'gll', (), None),
      ('h3_ne2_p1', 3, 2, 2, 'gll', (), None),
      ('h3_ne3_p4_shuffled', 3, 3, 5, 'gll', (), 11),
      ('h3_ne2_p3_shuffled', 3, 2, 4, 'gll', (), 3),
      ('l1_ne5_p3', 1, 5, 4, 'gll', (), None),
  ]
  out = {}
  names = []
  for name, ndim, ne, n, tname, per, seed in cases:
    premesh = R.premesh_commons.unit_cube_mesh(
        ne, ndim=ndim, a=-1., b=1., periodic_dims=per)
    if seed is not None:
      premesh, _ = _shuffled_premesh(premesh, seed)
    refined = R.mesh_refiner.refine_premesh(
        premesh, I.Nodes1D.create(n, TYPES[tname]))
    mesh = refined.finalize()
    out.update(_mesh_arrays(name + '/', premesh, refined, mesh))
    out[name + '/meta'] = np.array(
        [ndim, ne, n, {'gll': 0, 'gl': 1, 'nc': 2}[tname]] + [
            int(d in per) for d in range(3)])
    names.append(name)
  out['names'] = np.array(names)
  save('connectivity', **out)


def golden_partition():
  """Index builders of the partitioned branch (premesh.py:170-222)."""
  out = {}
  gs = R.gather_scatter
  for name, ndim, ne, n, pgrid, per in (
      ('q2_ne4_p2_2x2', 2, 4, 3, (2, 2), ()),
      ('q2_ne4_p3_2x1', 2, 4, 4, (2, 1), ()),
      ('h3_ne2_p2_2x2x2', 3, 2, 3, (2, 2, 2), ()),
      ('q2_ne4_p2_4x1_xper', 2, 4, 3, (4, 1), (0,)),
  ):
    partitions = np.arange(int(np.prod(pgrid))).reshape(pgrid)
    premesh = R.premesh_commons.unit_cube_mesh(
        ne, ndim=ndim, a=-1., b=1., partitions=partitions, periodic_dims=per)
    refined = R.mesh_refiner.refine_premesh(
        premesh, I.Nodes1D.create(n, TYPES['gll']))
    element_indices = gs.group_by_partitions(refined.partitions)
    elements = np.stack([
        gs.gather(refined.elements[:, k], indices=element_indices)
        for k in range(refined.elements.shape[1])], axis=-1)
    node_indices, local_elements = gs.get_local_elements(elements)
    node_indices = gs.get_unique_node_indices(
        node_indices, periodic_links=refined.periodic_links)
    gi, ui = gs.get_exchange_indices(node_indices)
    assert ui is None
    p = name + '/'
    out[p + 'partitions'] = refined.partitions
    out[p + 'elements'] = refined.elements
    out[p + 'node_coords'] = refined.node_coords
    out[p + 'element_indices'] = element_indices
    out[p + 'node_indices'] = node_indices
    out[p + 'local_elements'] = local_elements
    out[p + 'exchange_gather_indices'] = gi
    if refined.periodic_links is not None:
      out[p + 'periodic_links'] = refined.periodic_links
    out[p + 'meta'] = np.array([ndim, ne, n] + list(pgrid))
  # known-answer vectors of gather_scatter_test.py are restated in tests/.
  save('partition', **out)


# ----------------------------------------------------------------------------
# 3. FE space + operator: the reference's own local_covector / integrate
# ----------------------------------------------------------------------------


def _deform(x):
  x = np.asarray(x, dtype=np.float64)
  ndim = x.shape[-1]
  perm = np.roll(np.arange(ndim), 1) if ndim > 1 else np.arange(ndim)
  return x + 0.08 * np.sin(np.pi * x[:, perm] + 0.3) * (1 - 0.5 * x**2)


def _forms():
  grad = R.fespace.grad

  def mass(u, v):  # examples/poisson.py:133-134
    return lambda x: u(x) * v(x)

  def stiffness(u, v):  # examples/poisson.py:136-137
    return lambda x: jnp.vdot(grad(u)(x), grad(v)(x))

  def vstiffness(u, v):  # navier_stokes.py:222-223
    return lambda x: jnp.einsum('ij,ij->', grad(u)(x), grad(v)(x))

  def vmass(u, v):  # navier_stokes.py:231-232
    return lambda x: jnp.vdot(u(x), v(x))

  return mass, stiffness, vstiffness, vmass


def golden_operator():
  mass, stiffness, vstiffness, vmass = _forms()
  cases = [
      # name, ndim, ne, N(GLL), quad type, Q, shuffle seed, deform nodes?
      ('q2_ne2_p3_gl4', 2, 2, 4, 'gl', 4, None, True),
      ('q2_ne2_p3_colloc', 2, 2, 4, 'gll', 4, None, True),
      ('q2_ne2_p4_gl5_shuffled', 2, 2, 5, 'gl', 5, 5, True),
      ('q2_ne1_p5_gll8', 2, 1, 6, 'gll', 8, None, True),
      ('h3_ne2x_p2_gl4', 3, 2, 3, 'gl', 4, None, True),
      ('h3_ne2x_p2_colloc', 3, 2, 3, 'gll', 3, None, True),
      ('l1_ne4_p3_gl4', 1, 4, 4, 'gl', 4, None, False),
  ]
  out = {}
  names = []
  rng = np.random.default_rng(1234)
  for name, ndim, ne, n, qt, q, seed, deform in cases:
    premesh = R.premesh_commons.unit_cube_mesh(ne, ndim=ndim, a=-1., b=1.)
    if name.startswith('h3_ne2x'):
      # keep it tiny: only the first two elements of the 2^3 mesh
      els = np.asarray(premesh.elements)[:2]
      used = np.unique(els)
      remap = -np.ones(premesh.num_nodes, dtype=np.int64)
      remap[used] = np.arange(len(used))
      premesh = R.premesh.Premesh.create(
          node_coords=premesh.node_coords[used],
          elements=remap[els].astype(np.int32))
    if seed is not None:
      premesh, _ = _shuffled_premesh(premesh, seed)
    refined = R.mesh_refiner.refine_premesh(
        premesh, I.Nodes1D.create(n, TYPES['gll']))
    coords = _deform(refined.node_coords) if deform else refined.node_coords
    refined = R.premesh.Premesh.create(
        node_coords=coords, elements=refined.elements,
        gridpoints_1d=refined.gridpoints_1d,
        physical_groups=refined.physical_groups)
    mesh = refined.finalize()
    quad = I.Quadrature1D.create(q, TYPES[qt])
    fes = R.fespace.FiniteElementSpace.create(mesh, quad)
    u = jnp.asarray(rng.standard_normal(mesh.num_nodes))
    u_local = mesh.gather(u)
    uf = fes.scalar_function(u_local)
    vf = fes.scalar_function(None)
    p = name + '/'
    out[p + 'node_coords'] = mesh.node_coords
    out[p + 'elements'] = mesh.elements
    out[p + 'u'] = u
    out[p + 'u_local'] = u_local
    out[p + 'invjacs'] = fes.invjacs
    out[p + 'jacdets'] = fes.jacdets
    out[p + 'quad_coords'] = fes.quad_coords
    out[p + 'eval_u'] = uf._evaluate()  # pylint: disable=protected-access
    out[p + 'eval_grad_u'] = R.fespace.grad(uf)._evaluate()  # pylint: disable=protected-access
    out[p + 'integral_u'] = fes.integrate(uf)
    out[p + 'integral_one'] = fes.integrate(lambda x: jnp.ones(()))
    out[p + 'mass_local'] = fes.local_covector(mass, (uf, vf))
    out[p + 'stiffness_local'] = fes.local_covector(stiffness, (uf, vf))
    out[p + 'stiffness'] = mesh.scatter(out[p + 'stiffness_local'])
    if ndim == 2 and ne <= 2 and n <= 4:
      uv = jnp.asarray(rng.standard_normal((mesh.num_nodes, ndim)))
      uv_local = jnp.stack([mesh.gather(uv[:, k]) for k in range(ndim)], -1)
      uvf = fes.vector_function(uv_local)
      vvf = fes.vector_function(None)
      out[p + 'uv'] = uv
      out[p + 'eval_uv'] = uvf._evaluate()  # pylint: disable=protected-access
      out[p + 'eval_grad_uv'] = R.fespace.grad(uvf)._evaluate()  # pylint: disable=protected-access
      out[p + 'vstiffness_local'] = fes.local_covector(vstiffness, (uvf, vvf))
      out[p + 'vmass_local'] = fes.local_covector(vmass, (uvf, vvf))
      out[p + 'integral_div'] = fes.integrate(R.fespace.div(uvf))
    out[p + 'meta'] = np.array(
        [ndim, ne, n, {'gll': 0, 'gl': 1}[qt], q])
    names.append(name)
    print('operator golden', name, 'done')
  out['names'] = np.array(names)
  save('operator', **out)


# ----------------------------------------------------------------------------
# 4. CG (swirl_fem/linalg/cg.py) on small dense SPD systems
# ----------------------------------------------------------------------------


def golden_cg():
  out = {}
  rng = np.random.default_rng(99)
  n = 60
  q, _ = np.linalg.qr(rng.standard_normal((n, n)))
  eig = np.logspace(0, 3, n)
  mat = (q * eig) @ q.T
  mat = 0.5 * (mat + mat.T)
  b = rng.standard_normal(n)
  dinv = 1.0 / np.diag(mat)
  A = lambda x: jnp.asarray(mat @ np.asarray(x))
  for tag, kw in (
      ('plain', dict(tol=1e-8)),
      ('jacobi', dict(tol=1e-8, M=lambda r: jnp.asarray(dinv * np.asarray(r)))),
      ('atol', dict(tol=0., atol=1e-3)),
      ('maxiter', dict(tol=1e-14, maxiter=7)),
      ('x0', dict(tol=1e-6, x0=jnp.asarray(np.ones(n)))),
  ):
    x, info = R.cg.cg(A, jnp.asarray(b), **kw)
    out[f'{tag}/x'] = x
    out[f'{tag}/residual'] = info['residual']
    out[f'{tag}/num_iterations'] = info['num_iterations']
  out['mat'] = mat
  out['b'] = b
  # reference's own known answers (linalg/cg_test.py:26-50)
  x, info = R.cg.cg(lambda x: 2 * x, jnp.arange(9.0).reshape((3, 3)))
  out['kat_2x/x'] = x
  out['kat_2x/num_iterations'] = info['num_iterations']
  x, info = R.cg.cg(
      lambda x: jnp.array([2 * x[0], 0 * x[1]]), 1 + jnp.arange(2.0),
      M=lambda x: jnp.array([x[0], 0.]))
  out['kat_singular/x'] = x
  out['kat_singular/num_iterations'] = info['num_iterations']
  save('cg', **out)


def main(argv):
  which = set(argv[1:]) or {'interpolation', 'connectivity', 'partition',
                            'operator', 'cg'}
  if 'interpolation' in which:
    golden_interpolation()
  if 'connectivity' in which:
    golden_connectivity()
  if 'partition' in which:
    golden_partition()
  if 'cg' in which:
    golden_cg()
  if 'operator' in which:
    golden_operator()


if __name__ == '__main__':
  main(sys.argv)
