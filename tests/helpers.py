"""Shared builders for the parity tests (host-side, numpy)."""

import numpy as np

from swirl_fem_b200.common.premesh_commons import unit_cube_mesh
from swirl_fem_b200.core.interpolation import Nodes1D
from swirl_fem_b200.core.interpolation import NodeType
from swirl_fem_b200.core.mesh_refiner import refine_premesh
from swirl_fem_b200.core.premesh import Premesh

GLL = NodeType.GAUSS_LOBATTO_LEGENDRE
GL = NodeType.GAUSS_LEGENDRE
TNAME = {GLL: 'gauss_lobatto_legendre', GL: 'gauss_legendre',
         NodeType.NEWTON_COTES: 'newton_cotes'}


def deform(x):
  """Smooth non-affine map of [-1,1]^d keeping detJ > 0."""
  x = np.asarray(x, dtype=np.float64)
  ndim = x.shape[-1]
  perm = np.roll(np.arange(ndim), 1) if ndim > 1 else np.arange(ndim)
  return x + 0.08 * np.sin(np.pi * x[:, perm] + 0.3) * (1 - 0.5 * x ** 2)


def shuffled(premesh: Premesh, seed: int, reorient: bool = True) -> Premesh:
  """Random element order + (reorient) per-element axis permutation / flips.

  Reference quirk (replicated bit-exactly by `core.mesh_refiner`, see the
  connectivity goldens): for re-oriented HEXES the refiner's face-orientation
  table places some shared face nodes inconsistently with the second
  element's own lexicographic positions, i.e. the refined geometry is tangled
  (det J off by up to 8x, condition numbers up to 1e5; 2-D is unaffected).
  Such meshes are kept for fp64 / connectivity tests; tests of the fp32
  tolerance use `reorient=False` in 3-D so that the geometry is a valid mesh."""
  rng = np.random.default_rng(seed)
  ndim = premesh.ndim
  elements = np.array(premesh.elements)[rng.permutation(premesh.num_elements)]
  if not reorient:
    return Premesh.create(node_coords=premesh.node_coords,
                          elements=np.array(elements, dtype=np.int32),
                          physical_groups=premesh.physical_groups,
                          periodic_links=premesh.periodic_links)
  out = []
  for el in elements:
    nd = el.reshape([2] * ndim).transpose(rng.permutation(ndim))
    flips = [ax for ax in range(ndim) if rng.integers(2)]
    nd = np.flip(nd, flips) if flips else nd
    out.append(nd.reshape(-1))
  return Premesh.create(node_coords=premesh.node_coords,
                        elements=np.array(out, dtype=np.int32),
                        physical_groups=premesh.physical_groups,
                        periodic_links=premesh.periodic_links)


def rotated_quads(premesh: Premesh, seed: int) -> Premesh:
  """Random element order + per-element 90-degree rotations of the vertex
  listing.  Unlike `shuffled`, every element keeps a POSITIVE Jacobian
  determinant (the reference integrates with the signed determinant,
  fespace.py:346, 402, so reflected elements make the operator indefinite)."""
  assert premesh.ndim == 2
  rng = np.random.default_rng(seed)
  elements = np.array(premesh.elements)[rng.permutation(premesh.num_elements)]
  out = [np.rot90(el.reshape(2, 2), k=int(rng.integers(4))).reshape(-1)
         for el in elements]
  return Premesh.create(node_coords=premesh.node_coords,
                        elements=np.array(out, dtype=np.int32),
                        physical_groups=premesh.physical_groups,
                        periodic_links=premesh.periodic_links)


def deformed_premesh(ndim, ne, n1d, seed=None, periodic_dims=(), curved=True,
                     rotate_seed=None, reorient=True):
  """Refined GLL premesh on [-1,1]^ndim with deformed (curved) elements."""
  pm = unit_cube_mesh(ne, ndim=ndim, a=-1., b=1., periodic_dims=periodic_dims)
  if seed is not None:
    pm = shuffled(pm, seed, reorient=reorient)
  if rotate_seed is not None:
    pm = rotated_quads(pm, rotate_seed)
  refined = refine_premesh(pm, Nodes1D.create(n1d, GLL))
  if curved and not periodic_dims:
    refined = refined.replace(node_coords=deform(refined.node_coords))
  return refined


def local_to_global(global_coords, local_coords):
  """Index of every local node in the global mesh, matched through the
  (undeformed) coordinates."""
  ndim = global_coords.shape[1]
  key = lambda c: np.round(np.asarray(c) * 1e9).astype(np.int64)  # noqa: E731
  gk, lk = key(global_coords), key(local_coords)
  order = np.lexsort(gk.T[::-1])
  rec = [('', np.int64)] * ndim
  view_g = np.ascontiguousarray(gk[order]).view(rec).ravel()
  view_l = np.ascontiguousarray(lk).view(rec).ravel()
  l2g = order[np.searchsorted(view_g, view_l)]
  assert np.array_equal(gk[l2g], lk)
  return l2g


# -- Stokes / Navier-Stokes fixtures (navier_stokes_test.py:39-68) ---------------


def stokes_vortices_premesh(ne=9, curved=0.0):
  """Uniform premesh on [-1, 1] x [-pi, pi], periodic in y
  (`_make_stokes_vortices_premesh`, navier_stokes_test.py:39-44); `curved`
  bends the element columns (keeps the y-periodicity)."""
  pm = unit_cube_mesh(ne, ndim=2, periodic_dims=(1,))
  x = np.asarray(pm.node_coords, dtype=np.float64)
  x = np.stack([2 * x[:, 0] - 1, 2 * np.pi * x[:, 1] - np.pi], -1)
  if curved:
    x[:, 0] += curved * np.sin(x[:, 1]) * (1 - x[:, 0] ** 2)
  return pm.replace(node_coords=x)


def stokes_oracle_meshes(premesh, order, boundary='boundary'):
  """Host meshes (numpy) of the velocity (GLL) and pressure (GL) spaces for
  `oracle.dense_ns.StokesSEM`, built with the host mesh code that the
  connectivity goldens pin bit-exactly to the reference."""
  vref = refine_premesh(premesh, Nodes1D.create(order + 1, GLL))
  pref = refine_premesh(premesh, Nodes1D.create(order - 1, GL))
  vhost = vref.finalize_host()
  vmesh = dict(node_coords=vref.node_coords, elements=vref.elements,
               interior_mask=(1.0 - vhost['physical_masks'][boundary]
                              if boundary else np.ones(vref.num_nodes)),
               exchange_gather_indices=vhost['exchange_gather_indices'],
               exchange_unique_indices=vhost['exchange_unique_indices'])
  pmesh = dict(node_coords=pref.node_coords, elements=pref.elements)
  return vmesh, pmesh


def stokes_reference_soln_params(k=1., viscosity=1.):
  """navier_stokes_test.py:47-54."""
  import scipy.optimize  # pylint: disable=g-import-not-at-top
  mu = scipy.optimize.newton(lambda x: k * np.tanh(k) + x * np.tan(x), np.pi)
  return mu, -viscosity * (np.square(k) + np.square(mu))


def stokes_reference_soln(vcoords, pcoords, t, k=1., viscosity=1.):
  """Analytical Stokes vortices (navier_stokes_test.py:57-68)."""
  mu, sigma = stokes_reference_soln_params(k, viscosity)
  f = lambda x: np.cos(mu) * np.cosh(k * x) - np.cosh(k) * np.cos(mu * x)  # noqa: E731
  g = lambda x: (1j / k) * (k * np.cos(mu) * np.sinh(k * x) +  # noqa: E731
                            mu * np.cosh(k) * np.sin(mu * x))
  h = lambda x: -(sigma / k) * np.cos(mu) * np.sinh(k * x)  # noqa: E731
  lead = lambda x: np.exp(sigma * t) * np.exp(1j * k * x[:, 1])  # noqa: E731
  u = np.real(lead(vcoords)[:, None] * np.stack(
      [f(vcoords[:, 0]), g(vcoords[:, 0])], -1))
  p = np.real(lead(pcoords) * h(pcoords[:, 0]))
  return u, p


# -- Gmsh .msh writers (fixtures for the reader tests) -----------------------------

_GMSH_FROM_LEX = {1: [0, 1], 2: [0, 2, 3, 1], 3: [0, 4, 6, 2, 1, 5, 7, 3]}
_GMSH_TYPE = {'line': 1, 'quad': 3, 'hexahedron': 5}


def write_msh41(path, points, cells, periodic=(), tag_stride=1):
  """Minimal ASCII MSH 4.1 writer.  `cells`: name -> (n, k) 0-based node
  indices in GMSH ordering; `periodic`: (entity_dim, (n, 2) node index pairs).
  Node tags are `1 + tag_stride * index` (non-contiguous for stride > 1)."""
  points = np.asarray(points, dtype=np.float64)
  pts = np.zeros((len(points), 3))
  pts[:, :points.shape[1]] = points
  tag = lambda i: 1 + tag_stride * int(i)  # noqa: E731
  out = ['$MeshFormat', '4.1 0 8', '$EndMeshFormat', '$Nodes',
         f'1 {len(pts)} 1 {tag(len(pts) - 1)}', f'2 1 0 {len(pts)}']
  out += [str(tag(i)) for i in range(len(pts))]
  out += [' '.join(repr(float(v)) for v in p) for p in pts]
  out += ['$EndNodes', '$Elements']
  total = sum(len(v) for v in cells.values())
  out.append(f'{len(cells)} {total} 1 {total}')
  etag = 1
  for k, (name, conn) in enumerate(cells.items()):
    dim = {'line': 1, 'quad': 2, 'hexahedron': 3}[name]
    out.append(f'{dim} {k + 1} {_GMSH_TYPE[name]} {len(conn)}')
    for row in conn:
      out.append(' '.join([str(etag)] + [str(tag(i)) for i in row]))
      etag += 1
  out.append('$EndElements')
  if periodic:
    out += ['$Periodic', str(len(periodic))]
    for dim, pairs in periodic:
      out += [f'{dim} 2 1', '16 1 0 0 0 0 1 0 0 0 0 1 0 0 0 0 1',
              str(len(pairs))]
      out += [f'{tag(a)} {tag(b)}' for a, b in pairs]
    out.append('$EndPeriodic')
  with open(path, 'w') as f:
    f.write('\n'.join(out) + '\n')


def write_msh22(path, points, cells, tag_stride=1):
  points = np.asarray(points, dtype=np.float64)
  pts = np.zeros((len(points), 3))
  pts[:, :points.shape[1]] = points
  tag = lambda i: 1 + tag_stride * int(i)  # noqa: E731
  out = ['$MeshFormat', '2.2 0 8', '$EndMeshFormat', '$Nodes', str(len(pts))]
  out += [' '.join([str(tag(i))] + [repr(float(v)) for v in p])
          for i, p in enumerate(pts)]
  out += ['$EndNodes', '$Elements',
          str(sum(len(v) for v in cells.values()))]
  etag = 1
  for name, conn in cells.items():
    for row in conn:
      out.append(' '.join([str(etag), str(_GMSH_TYPE[name]), '2', '1', '1'] +
                          [str(tag(i)) for i in row]))
      etag += 1
  out.append('$EndElements')
  with open(path, 'w') as f:
    f.write('\n'.join(out) + '\n')


def write_premesh_as_msh(path, premesh, version='4.1', tag_stride=1):
  """Writes a first-order `Premesh` (lexicographic cells) as a Gmsh file with
  Gmsh's cell ordering; periodic links become facet cells + `$Periodic`."""
  ndim = premesh.ndim
  name = {1: 'line', 2: 'quad', 3: 'hexahedron'}[ndim]
  cells = {name: np.asarray(premesh.elements)[:, _GMSH_FROM_LEX[ndim]]}
  periodic = []
  links = premesh.periodic_links
  if links is not None and len(links):
    links = np.asarray(links)
    fname = {2: 'line', 3: 'quad'}[ndim]
    # facet cells (any consistent ordering: the reader keeps it as listed)
    cells = {fname: np.concatenate([links[:, 0], links[:, 1]]), **cells}
    pairs = np.unique(np.stack([links[:, 0].reshape(-1),
                                links[:, 1].reshape(-1)], axis=1), axis=0)
    periodic = [(ndim - 1, pairs)]
  if version.startswith('4'):
    write_msh41(path, premesh.node_coords, cells, periodic, tag_stride)
  else:
    assert not periodic
    write_msh22(path, premesh.node_coords, cells, tag_stride)
