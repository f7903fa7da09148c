"""Host logic of `swirl_fem_b200.navier_stokes` / `examples.kolmogorov` on CPU.

The CUDA entry points are replaced by test doubles (as `CpuHaloPlan` does for
the halo protocol): element-local operators come from the numpy oracle, the
vector kernels `sfem_dot` / `sfem_axpby` from torch.  What runs for real is the
Python that composes them -- `StokesSEM.B/Bi/A/D/Dt/Q/E/filter`,
`_pressure_project_out_nullspace`, `stokes_one_step`, `linalg.cg.cg`'s generic
path and the Kolmogorov step -- and it must reproduce the oracle's
`stokes_one_step` / `navier_stokes_one_step` (iteration counts included).
The product path itself has no CPU fallback; these doubles live in the test.
"""

import numpy as np
import pytest
import torch

from oracle import dense_ns
from swirl_fem_b200 import _lib
from swirl_fem_b200.examples import kolmogorov
from swirl_fem_b200.navier_stokes import navier_stokes as ns
from tests import helpers


def t(x):
  return torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float64)


class _FakeMesh:

  def __init__(self, coords, nodes_1d):
    self.node_coords = t(coords)
    self.num_nodes = coords.shape[0]
    self.ndim = coords.shape[1]
    self.gridpoints_1d = nodes_1d


class _FakeVelocity:
  """`StokesVelocity` surface backed by the oracle."""

  def __init__(self, osem):
    self.o = osem
    self.interior_mask = t(osem.interior_mask)
    self.diag_qqt = t(osem.diag_qqt)
    self.mesh = _FakeMesh(osem.vmesh['node_coords'], None)

  def gather(self, u):
    return t(self.o.v_gather(u.numpy()))

  def scatter(self, ul):
    return t(self.o.v_scatter(ul.numpy()))

  def exchange(self, u):
    return t(self.o.v_exchange(u.numpy()))

  def C(self, u):
    return t(self.o.C(u.numpy()))


class _FakePressure:

  def __init__(self, osem):
    self.o = osem

  def gather(self, p):
    return t(self.o.pspace.gather(p.numpy()))

  def scatter(self, pl):
    return t(self.o.pspace.scatter(pl.numpy()))

  def B(self, p):
    return t(self.o.pressure_B(p.numpy()))

  def exchange(self, p):
    return p


class _FakeOperator:
  """`FusedOperator.apply` for the masked vector stiffness."""

  def __init__(self, osem):
    self.o = osem

  def apply(self, u, lam=0.0, mu=1.0):
    assert lam == 0.0
    return t(mu * self.o.A(u.numpy()))


class _FakeFilterSpace:

  def __init__(self, osem):
    from oracle import dense
    n, d = osem.order + 1, osem.vspace.ndim
    self.low = dense.Interp(d, n, dense_ns.GLL, n - 1, dense_ns.GLL)
    self.high = dense.Interp(d, n - 1, dense_ns.GLL, n, dense_ns.GLL)

  def apply(self, ul):
    ul = ul.numpy()
    return t(np.stack([self.high.interpolate(self.low.interpolate(ul[..., k]))
                       for k in range(ul.shape[-1])], -1))


class _CpuSem(ns.StokesSEM):
  """The real `StokesSEM` methods over oracle-backed element kernels."""

  def D_local(self, u_local):
    return t(self.velocity.o.D_local(u_local.numpy()))

  def Dt_local(self, p_local):
    return t(self.velocity.o.Dt_local(p_local.numpy()))


@pytest.fixture
def cpu_kernels(monkeypatch):
  """torch stand-ins for the three C-ABI vector entry points cg() uses."""
  monkeypatch.setattr(_lib, 'require_cuda', lambda *a: None)
  monkeypatch.setattr(_lib, 'dot',
                      lambda a, b: (a.double() * b.double()).sum())

  def axpby(a, x, b, y):
    y.mul_(b).add_(x, alpha=a)
  monkeypatch.setattr(_lib, 'axpby', axpby)


def _make(premesh, order, boundary='boundary'):
  vmesh, pmesh = helpers.stokes_oracle_meshes(premesh, order, boundary=boundary)
  osem = dense_ns.StokesSEM(vmesh, pmesh, order)
  sem = _CpuSem(velocity=_FakeVelocity(osem), pressure=_FakePressure(osem),
                velocity_mass_diag=t(osem.velocity_mass_diag))
  sem._cache['vop'] = _FakeOperator(osem)          # pylint: disable=protected-access
  sem._cache['filter_space'] = _FakeFilterSpace(osem)  # pylint: disable=protected-access
  return sem, osem, vmesh, pmesh


def test_operator_compositions_match_oracle(cpu_kernels):
  sem, osem, vmesh, pmesh = _make(helpers.stokes_vortices_premesh(3, 0.1), 4)
  rng = np.random.default_rng(0)
  u = rng.standard_normal((vmesh['node_coords'].shape[0], 2))
  p = rng.standard_normal(pmesh['node_coords'].shape[0])
  dt, k = 1e-3, 3
  pairs = [
      (sem.B(t(u)), osem.B(u)), (sem.Bi(t(u)), osem.Bi(u)),
      (sem.A(t(u)), osem.A(u)), (sem.D(t(u)), osem.D(u)),
      (sem.Dt(t(p)), osem.Dt(p)),
      (sem.Q(t(u), dt=dt, time_order=k), osem.Q(u, dt, k)),
      (sem.E(t(p), dt=dt, time_order=k), osem.E(p, dt, k)),
      (sem.filter(t(u), alpha=0.05), osem.filter(u, 0.05)),
      (ns._pressure_project_out_nullspace(sem, t(p)),  # pylint: disable=protected-access
       osem.project_out_nullspace(p)),
  ]
  for got, want in pairs:
    np.testing.assert_allclose(got.numpy(), want, rtol=1e-12, atol=1e-12)


def test_stokes_one_step_host_logic_matches_oracle(cpu_kernels):
  """Same step, same CG iteration counts, with and without a boundary lift."""
  pm = helpers.stokes_vortices_premesh(4)
  sem, osem, vmesh, pmesh = _make(pm, 5)
  k, dt = 3, 1e-3
  states = [helpers.stokes_reference_soln(vmesh['node_coords'],
                                          pmesh['node_coords'], i * dt)
            for i in range(k)]
  us, ps = zip(*states)
  rng = np.random.default_rng(4)
  lift = 1e-3 * rng.standard_normal(us[0].shape) * (1 - osem.interior_mask)
  for u_boundary in (None, lift):
    u, p, aux = sem.stokes_one_step(
        [t(v) for v in us], [t(q) for q in ps], f=0, mu=1, dt=dt,
        time_order=k, alpha=0.05, tol=1e-10, atol=1e-12,
        u_boundary=None if u_boundary is None else t(u_boundary))
    uo, po, auxo = osem.stokes_one_step(us, ps, f=0, mu=1, dt=dt, time_order=k,
                                        alpha=0.05, tol=1e-10, atol=1e-12,
                                        u_boundary=u_boundary)
    for key in ('u_star_info', 'dp_info'):
      assert aux[key]['num_iterations'] == auxo[key]['num_iterations']
    np.testing.assert_allclose(u.numpy(), uo, rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(p.numpy(), po, rtol=1e-7, atol=1e-9)


def test_kolmogorov_step_host_logic_matches_oracle(cpu_kernels):
  from swirl_fem_b200.common.premesh_commons import unit_cube_mesh
  pm = unit_cube_mesh(3, ndim=2, periodic_dims=(0, 1))
  sem, osem, vmesh, pmesh = _make(pm, 5, boundary=None)
  k = 3
  u0 = kolmogorov.u_init_fn(t(vmesh['node_coords']))
  np.testing.assert_allclose(
      u0.numpy(), dense_ns.kolmogorov_u_init(vmesh['node_coords']), atol=1e-15)
  np.testing.assert_allclose(
      kolmogorov.forcing(t(vmesh['node_coords']), u0).numpy(),
      dense_ns.kolmogorov_forcing(vmesh['node_coords'], u0.numpy()), atol=1e-15)
  us = (u0,) * k
  ps = (torch.zeros(pmesh['node_coords'].shape[0], dtype=torch.float64),) * k
  kw = dict(reynolds_number=500., dt=1e-3, time_order=k, tol=1e-9, atol=1e-12)
  us2, ps2, data = kolmogorov.one_cycle(sem, 0, 2, us, ps, sample_every=1, **kw)
  ous = tuple(u.numpy() for u in us)
  ops_ = tuple(p.numpy() for p in ps)
  oCus = tuple(osem.C(u) for u in ous)
  for _ in range(2):
    uo, po, Cuo, _ = dense_ns.navier_stokes_one_step(osem, ous, ops_, oCus, **kw)
    ous, ops_, oCus = ous[1:] + (uo,), ops_[1:] + (po,), oCus[1:] + (Cuo,)
  np.testing.assert_allclose(us2[-1].numpy(), ous[-1], rtol=1e-8, atol=1e-10)
  np.testing.assert_allclose(ps2[-1].numpy(), ops_[-1], rtol=1e-6, atol=1e-8)
  assert data['u'].shape[0] == 3 and len(data['t']) == 3
  np.testing.assert_allclose(data['t'], [0.0, 1e-3, 2e-3])


def test_create_wiring_with_stub_spaces(monkeypatch):
  """`StokesSEM.create` on stubs of the device classes: which node sets and
  quadrature rules are requested (navier_stokes.py:112-117, 174-190, 262-292)
  and that every attribute the constructors touch exists."""
  from swirl_fem_b200.core.interpolation import NodeType
  made = []

  class StubMesh:

    def __init__(self, premesh, gridpoints_1d):
      n = gridpoints_1d.num_points ** premesh.ndim
      self.gridpoints_1d = gridpoints_1d
      self.ndim = premesh.ndim
      self.num_elements = premesh.num_elements
      self.num_nodes_per_element = n
      self.num_nodes = 7 * n
      self.node_coords = torch.zeros(self.num_nodes, premesh.ndim,
                                     dtype=torch.float64)
      self.elements = torch.zeros(self.num_elements, n, dtype=torch.int32)
      self.device = torch.device('cpu')
      self.physical_masks = {
          'boundary': torch.arange(self.num_nodes) % 3 == 0}

    def scatter(self, u_local):
      assert tuple(u_local.shape) == tuple(self.elements.shape)
      return torch.full((self.num_nodes,), 2.0, dtype=torch.float64)

  class StubRefined:

    def __init__(self, premesh, gridpoints_1d):
      self.args = (premesh, gridpoints_1d)

    def finalize(self, device=None, dtype=None):
      return StubMesh(*self.args)

  class StubSpace:

    def __init__(self, mesh, quadrature):
      self.mesh, self.quadrature = mesh, quadrature
      self.dtype = torch.float64
      made.append(self)

    @classmethod
    def create(cls, mesh, quadrature):
      return cls(mesh, quadrature)

    def vector_function(self, u_local):
      return ('vector', u_local)

    def local_covector(self, form, funs):
      assert form is ns._VECTOR_MASS_FORM  # pylint: disable=protected-access
      assert funs[1] == ('vector', None)
      return torch.ones_like(funs[0][1])

  monkeypatch.setattr(ns, 'refine_premesh',
                      lambda premesh, gridpoints_1d: StubRefined(
                          premesh, gridpoints_1d))
  monkeypatch.setattr(ns, 'FiniteElementSpace', StubSpace)
  # the AoS scatter is ONE C-ABI launch for all components (`_lib.scatter`)

  def stub_vector_scatter(u_local, indices, num_nodes):
    assert tuple(u_local.shape[:2]) == tuple(indices.shape)
    return torch.full((num_nodes,) + tuple(u_local.shape[2:]), 2.0,
                      dtype=torch.float64)

  monkeypatch.setattr(ns._lib, 'scatter', stub_vector_scatter)  # pylint: disable=protected-access
  order = 5
  sem = ns.StokesSEM.create(
      helpers.stokes_vortices_premesh(2),
      boundary_conditions={'boundary': (ns.BCType.DIRICHLET, 0.0)},
      order=order, num_convection_overint_nodes=3)
  vs, ov, ps = sem.velocity.vspace, sem.velocity.overint_space, (
      sem.pressure.pspace)
  gll, gl = NodeType.GAUSS_LOBATTO_LEGENDRE, NodeType.GAUSS_LEGENDRE
  assert vs.mesh is ov.mesh                      # one velocity mesh, two rules
  assert vs.mesh.gridpoints_1d.num_points == order + 1
  assert vs.mesh.gridpoints_1d.node_type == gll
  assert (vs.quadrature.num_points, ov.quadrature.num_points) == (
      order + 1, order + 1 + 3)
  assert ps.mesh.gridpoints_1d.num_points == order - 1
  assert ps.mesh.gridpoints_1d.node_type == gl
  assert ps.quadrature.num_points == order + 1   # the velocity's GLL rule
  assert sem.velocity.num_convection_overint_nodes == 3
  assert tuple(sem.velocity.interior_mask.shape) == (vs.mesh.num_nodes, 1)
  assert float(sem.velocity.interior_mask.sum()) == float(
      (~vs.mesh.physical_masks['boundary']).sum())
  assert tuple(sem.velocity_mass_diag.shape) == (vs.mesh.num_nodes, 2)
  assert sem.velocity.local_shape == (4, (order + 1) ** 2, 2)
  with pytest.raises(ValueError, match='order 1'):
    from swirl_fem_b200.core.interpolation import Nodes1D
    from swirl_fem_b200.core.mesh_refiner import refine_premesh as real_refine
    ns.StokesSEM.create(
        real_refine(helpers.stokes_vortices_premesh(2), Nodes1D.create(3, gll)),
        boundary_conditions={}, order=order)
