"""Host-side mesh I/O and partitioning: Gmsh reader and
element partitioner with the reference's contracts
(`swirl_fem/common/mesh_reader_test.py:25-74`,
`swirl_fem/common/mesh_partitioner_test.py:37-80`).

The `.msh` fixtures are WRITTEN by this test (Gmsh node ordering, `$Periodic`
section); when the reference checkout is present (build container only) its
own `testdata/*.msh` files are read as well and must give the sizes the
reference's tests assert.
"""

import collections
import math
import os

import numpy as np
import pytest

from swirl_fem_b200.common import mesh_partitioner
from swirl_fem_b200.common import mesh_reader
from swirl_fem_b200.core.interpolation import Nodes1D, NodeType
from swirl_fem_b200.core.mesh_refiner import refine_premesh
from swirl_fem_b200.core.premesh import Premesh
from tests import helpers

REF_DATA = '/root/reference/swirl_fem/testdata'
needs_ref = pytest.mark.skipif(not os.path.isdir(REF_DATA),
                               reason='reference checkout not present')


def test_single_quad_element_is_reordered(tmp_path):
  """mesh_reader_test.py:39-57."""
  path = tmp_path / 'one.msh'
  helpers.write_msh41(path, [[0, 0], [1, 0], [1, 1], [0, 1]],
                      {'quad': [[0, 1, 2, 3]]})
  pm = mesh_reader.read(path, ndim=2)
  assert pm.node_coords.shape == (4, 2)
  np.testing.assert_array_equal(pm.node_coords,
                                [[0, 0], [1, 0], [1, 1], [0, 1]])
  np.testing.assert_array_equal(pm.elements, [[0, 3, 1, 2]])
  assert pm.periodic_links is None


@pytest.mark.parametrize('version', ['4.1', '2.2'])
@pytest.mark.parametrize('ndim', [1, 2, 3])
def test_structured_round_trip(tmp_path, ndim, version):
  """A structured mesh written in Gmsh ordering (with non-contiguous node
  tags) reads back as the lexicographic premesh it was made from."""
  ne = 3
  pm0 = helpers.unit_cube_mesh(ne, ndim=ndim, a=-1., b=1.)
  path = tmp_path / 'cube.msh'
  helpers.write_premesh_as_msh(path, pm0, version=version, tag_stride=3)
  pm = mesh_reader.read(path, ndim=ndim)
  np.testing.assert_allclose(pm.node_coords, pm0.node_coords, atol=1e-15)
  np.testing.assert_array_equal(pm.elements, pm0.elements)
  assert pm.periodic_links is None
  # positive orientation after the permutation: refinement + geometry work
  refined = refine_premesh(pm, Nodes1D.create(3, NodeType.GAUSS_LOBATTO_LEGENDRE))
  assert refined.num_nodes == (2 * ne + 1) ** ndim


@pytest.mark.parametrize('ndim', [2, 3])
def test_periodic_links_from_periodic_section(tmp_path, ndim):
  ne = 2
  pm0 = helpers.unit_cube_mesh(ne, ndim=ndim, periodic_dims=(ndim - 1,))
  path = tmp_path / 'per.msh'
  helpers.write_premesh_as_msh(path, pm0, version='4.1')
  pm = mesh_reader.read(path, ndim=ndim)
  assert pm.periodic_links.shape == (ne ** (ndim - 1), 2, 2 ** (ndim - 1))
  # every linked node pair differs by the period along the periodic axis
  x = pm.node_coords
  diff = x[pm.periodic_links[:, 1]] - x[pm.periodic_links[:, 0]]
  want = np.zeros(ndim)
  want[ndim - 1] = diff.reshape(-1, ndim)[0, ndim - 1]
  assert abs(abs(want[ndim - 1]) - 1.0) < 1e-14
  np.testing.assert_allclose(diff.reshape(-1, ndim),
                             np.broadcast_to(want, (diff.size // ndim, ndim)),
                             atol=1e-14)
  # same dof identification as the generator's own links
  a = refine_premesh(pm, Nodes1D.create(4, NodeType.GAUSS_LOBATTO_LEGENDRE))
  b = refine_premesh(pm0, Nodes1D.create(4, NodeType.GAUSS_LOBATTO_LEGENDRE))
  ha, hb = a.finalize_host(), b.finalize_host()
  assert len(np.unique(ha['node_indices'])) == len(np.unique(hb['node_indices']))


def test_errors(tmp_path):
  path = tmp_path / 'one.msh'
  helpers.write_msh41(path, [[0, 0], [1, 0], [1, 1], [0, 1]],
                      {'quad': [[0, 1, 2, 3]]})
  with pytest.raises(ValueError, match='Invalid ndim'):
    mesh_reader.read(path, ndim=4)
  with pytest.raises(ValueError, match='hexahedron'):
    mesh_reader.read(path, ndim=3)
  bad = tmp_path / 'bin.msh'
  bad.write_text('$MeshFormat\n4.1 1 8\n$EndMeshFormat\n')
  with pytest.raises(NotImplementedError):
    mesh_reader.read(bad, ndim=2)


@needs_ref
@pytest.mark.parametrize('name,ndim,nodes,elems,links', [
    ('line1d', 1, (17, 1), (16, 2), None),
    ('kovasznay', 2, (65, 2), (48, 4), (4, 2, 2)),
    ('cube', 3, (125, 3), (64, 8), None),
    ('periodic_cube', 3, (125, 3), (64, 8), (48, 2, 4)),
])
def test_reference_testdata_sizes(name, ndim, nodes, elems, links):
  """mesh_reader_test.py:27-74 on the reference's own files."""
  pm = mesh_reader.read(os.path.join(REF_DATA, name + '.msh'), ndim=ndim)
  assert pm.node_coords.shape == nodes
  assert pm.elements.shape == elems
  if links is None:
    assert pm.periodic_links is None
  else:
    assert pm.periodic_links.shape == links
  if name == 'line1d':
    np.testing.assert_array_almost_equal(np.sort(pm.node_coords.flatten()),
                                         np.linspace(0, 1, num=17))
  # every element is positively oriented in tensor-product order
  x = pm.node_coords[pm.elements]            # (E, 2^d, d)
  e0 = x[:, 2 ** (ndim - 1)] - x[:, 0]       # axis 0 is the slowest index
  if ndim == 1:
    assert (e0[:, 0] > 0).all()
  elif ndim == 2:
    e1 = x[:, 1] - x[:, 0]
    assert (e0[:, 0] * e1[:, 1] - e0[:, 1] * e1[:, 0] > 0).all()
  else:
    e1, e2 = x[:, 2] - x[:, 0], x[:, 1] - x[:, 0]
    assert (np.einsum('ei,ei->e', e0, np.cross(e1, e2)) > 0).all()


# -- partitioner ---------------------------------------------------------------


def _interval(num_elements):
  n = num_elements + 1
  return Premesh.create(
      node_coords=np.linspace(0, 1, n).reshape(n, 1),
      elements=np.array([[i, i + 1] for i in range(num_elements)]))


def _check_balanced(pm, num_partitions):
  counts = collections.Counter(np.asarray(pm.partitions).tolist())
  for pid, count in counts.items():
    assert 0 <= pid <= num_partitions - 1
    assert (math.floor(pm.num_elements / num_partitions) <= count
            <= math.ceil(pm.num_elements / num_partitions))
  assert len(pm.partitions) == pm.num_elements


@pytest.mark.parametrize('num_elements,num_partitions',
                         [(2, 2), (8, 2), (16, 4), (15, 4), (35, 8)])
def test_partition_1d(num_elements, num_partitions):
  """mesh_partitioner_test.py:37-60."""
  pm = mesh_partitioner.partition(_interval(num_elements), num_partitions)
  assert pm.num_nodes == num_elements + 1
  _check_balanced(pm, num_partitions)
  for p in range(num_partitions):
    elems = {i for i, k in enumerate(pm.partitions) if k == p}
    assert elems == set(range(min(elems), 1 + max(elems)))


@pytest.mark.parametrize('num_partitions', [2, 4, 8, 3, 5])
def test_partition_cube(num_partitions):
  """mesh_partitioner_test.py:62-80 (4 x 4 x 4 hexahedra = cube.msh's size)."""
  pm = mesh_partitioner.partition(helpers.unit_cube_mesh(4, ndim=3),
                                  num_partitions)
  assert pm.elements.shape == (64, 8)
  _check_balanced(pm, num_partitions)
  if num_partitions in (2, 4, 8):
    # structured cube: the blocks of the bench's block partition
    blocks = {2: 16 * 9 + 0, 4: None, 8: None}
    assert mesh_partitioner.edge_cut(pm) <= {2: 16 * 9, 4: 2 * 16 * 9,
                                             8: 3 * 16 * 9}[num_partitions]
    del blocks


def test_partitioned_premesh_feeds_the_index_builders():
  """The partition array drives `Premesh.partition_host` (the reference's
  `group_by_partitions` / `get_local_elements`, gather_scatter.py:355-445)."""
  pm = mesh_partitioner.partition(helpers.unit_cube_mesh(4, ndim=2), 4)
  refined = refine_premesh(pm, Nodes1D.create(3, NodeType.GAUSS_LOBATTO_LEGENDRE))
  host = refined.partition_host()
  assert host['element_indices'].shape[0] == 4
  valid = host['node_indices'] != -1
  # every global node is held by at least one partition
  assert set(np.unique(host['node_indices'][valid])) == set(
      range(refined.num_nodes))
