"""Gmsh `.msh` reader -> `Premesh` (host side, no third-party mesh library).

Same contract as the reference's `swirl_fem/common/mesh_reader.py:78-114`
(which delegates the parsing to `meshio`, not available here): `read(path,
ndim)` returns a `Premesh` with
  * `node_coords` = the first `ndim` coordinates of every node, in file order,
  * `elements` = the line / quad / hexahedron cells re-ordered from Gmsh's
    counter-clockwise listing to tensor-product (lexicographic, axis 0 slowest)
    order with the reference's permutations (`mesh_reader.py:40-44`),
  * `periodic_links` = for every `(ndim-1)`-facet cell whose nodes all have a
    periodic counterpart of dimension `ndim-1`, the pair (facet, image facet)
    (`mesh_reader.py:47-75`), or None.
The parser understands the ASCII MSH 4.1 and 2.2 formats (`$Nodes`,
`$Elements`, `$Periodic`); binary files raise `NotImplementedError`.
"""

from __future__ import annotations

import os

import numpy as np

from swirl_fem_b200.core.premesh import Premesh

# Gmsh node ordering -> tensor-product ordering (mesh_reader.py:24-44).
_NODE_ORDERING_PERMUTATIONS = {
    1: [0, 1],
    2: [0, 3, 1, 2],
    3: [0, 4, 3, 7, 1, 5, 2, 6],
}
# Gmsh element type -> (name, number of nodes); first-order cells only.
_GMSH_TYPES = {1: ('line', 2), 3: ('quad', 4), 5: ('hexahedron', 8),
               15: ('vertex', 1), 2: ('triangle', 3), 4: ('tetra', 4)}


def _sections(text: str) -> dict:
  """`{'Nodes': [lines...], ...}` for every `$Name ... $EndName` block."""
  out, name, buf = {}, None, []
  for raw in text.splitlines():
    line = raw.strip()
    if not line:
      continue
    if line.startswith('$End'):
      if name is not None:
        out.setdefault(name, buf)
      name, buf = None, []
    elif line.startswith('$'):
      name, buf = line[1:], []
    elif name is not None:
      buf.append(line)
  return out


def _parse_nodes_41(lines):
  nblocks, nnodes = (int(v) for v in lines[0].split()[:2])
  tags, coords, i = [], [], 1
  for _ in range(nblocks):
    _, _, parametric, n = (int(v) for v in lines[i].split())
    i += 1
    tags.extend(int(lines[i + k]) for k in range(n))
    i += n
    for k in range(n):
      coords.append([float(v) for v in lines[i + k].split()[:3]])
    i += n
    del parametric  # extra parametric coordinates sit past column 3
  assert len(tags) == nnodes, (len(tags), nnodes)
  return np.asarray(tags, dtype=np.int64), np.asarray(coords, dtype=np.float64)


def _parse_elements_41(lines):
  nblocks = int(lines[0].split()[0])
  cells, i = {}, 1
  for _ in range(nblocks):
    _, _, etype, n = (int(v) for v in lines[i].split())
    i += 1
    name, nn = _GMSH_TYPES.get(etype, (f'type{etype}', None))
    block = [[int(v) for v in lines[i + k].split()[1:]] for k in range(n)]
    i += n
    if nn is not None and block:
      cells.setdefault(name, []).extend(row[:nn] for row in block)
  return cells


def _parse_nodes_22(lines):
  n = int(lines[0])
  rows = [lines[1 + k].split() for k in range(n)]
  tags = np.asarray([int(r[0]) for r in rows], dtype=np.int64)
  coords = np.asarray([[float(v) for v in r[1:4]] for r in rows])
  return tags, coords


def _parse_elements_22(lines):
  n = int(lines[0])
  cells = {}
  for k in range(n):
    vals = [int(v) for v in lines[1 + k].split()]
    etype, ntags = vals[1], vals[2]
    name, nn = _GMSH_TYPES.get(etype, (f'type{etype}', None))
    if nn is not None:
      cells.setdefault(name, []).append(vals[3 + ntags:3 + ntags + nn])
  return cells


def _parse_periodic(lines):
  """List of `(entity_dim, node_pairs (n, 2) [node tag, master node tag])`."""
  nlinks, i, out = int(lines[0]), 1, []
  for _ in range(nlinks):
    entity_dim = int(lines[i].split()[0])
    i += 1
    head = lines[i].split()
    # optional affine transform: "Affine v0 ... v15" (2.2) or "16 v0 ..." (4.1)
    if head[0] == 'Affine' or (len(head) > 1 and int(float(head[0])) ==
                               len(head) - 1):
      i += 1
    n = int(lines[i])
    i += 1
    pairs = np.asarray([[int(v) for v in lines[i + k].split()[:2]]
                        for k in range(n)], dtype=np.int64).reshape(n, 2)
    i += n
    out.append((entity_dim, pairs))
  return out


def _get_periodic_links(cells, periodic, ndim: int) -> np.ndarray:
  """mesh_reader.py:47-75 on 0-based node indices."""
  src_tgt = {}
  for entity_dim, pairs in periodic:
    if entity_dim != ndim - 1:
      continue
    src_tgt.update(dict(pairs.tolist()))
  facet_type = {1: 'line', 2: 'quad'}[ndim - 1]
  links = []
  for facet in cells.get(facet_type, []):
    if all(x in src_tgt for x in facet):
      links.append(np.stack([facet, [src_tgt[x] for x in facet]]))
  return np.stack(links).astype(np.int32)


def read(path, ndim: int) -> Premesh:
  """Reads a Gmsh mesh file and parses it into a `Premesh`."""
  if ndim not in [1, 2, 3]:
    raise ValueError(f'Invalid ndim: {ndim=}. Valid spatial dimensions are '
                     '1, 2 and 3.')
  with open(os.fspath(path), 'rb') as f:
    raw = f.read()
  try:
    text = raw.decode('ascii')
  except UnicodeDecodeError as e:
    raise NotImplementedError('binary .msh files are not supported') from e
  sec = _sections(text)
  if 'MeshFormat' not in sec:
    raise ValueError(f'{path}: not a Gmsh .msh file (no $MeshFormat)')
  version, file_type = sec['MeshFormat'][0].split()[:2]
  if int(file_type) != 0:
    raise NotImplementedError('binary .msh files are not supported')
  if version.startswith('4'):
    tags, coords = _parse_nodes_41(sec['Nodes'])
    cells = _parse_elements_41(sec['Elements'])
  elif version.startswith('2'):
    tags, coords = _parse_nodes_22(sec['Nodes'])
    cells = _parse_elements_22(sec['Elements'])
  else:
    raise NotImplementedError(f'MSH format version {version}')
  # node tags -> 0-based indices in file order (tags need not be contiguous)
  index = np.full(int(tags.max()) + 1, -1, dtype=np.int64)
  index[tags] = np.arange(len(tags))
  cells = {k: index[np.asarray(v, dtype=np.int64)] for k, v in cells.items()}
  periodic = []
  if 'Periodic' in sec:
    periodic = [(dim, index[pairs]) for dim, pairs in
                _parse_periodic(sec['Periodic'])]

  elem_type = {1: 'line', 2: 'quad', 3: 'hexahedron'}[ndim]
  if elem_type not in cells:
    raise ValueError(
        f'Reading mesh of {ndim=} but cells of type {elem_type=} not found '
        f'in {sorted(cells)}')
  elements = cells[elem_type][:, _NODE_ORDERING_PERMUTATIONS[ndim]]
  periodic_links = (_get_periodic_links(cells, periodic, ndim)
                    if periodic and ndim > 1 else None)
  return Premesh.create(node_coords=coords[:, :ndim],
                        elements=elements.astype(np.int32),
                        periodic_links=periodic_links)
