"""N > 1 host logic on CPU: halo plans and the exchange protocol over gloo.

The pack / unpack-add steps are CUDA kernels in the product; here they are
replaced by index ops (test double) so that the *plan* and the *message
protocol* (who sends what to whom, in which order) are exercised with
world_size 2 and 4 on the CPU.  The wire results are compared with the
unpartitioned QQ^T semantics of the reference
(swirl_fem/core/gather_scatter.py:221-261) and its known answers
(swirl_fem/core/gather_scatter_test.py:157-263).
"""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from swirl_fem_b200.communication import partition as part
from swirl_fem_b200.communication.halo import HaloPlan
from swirl_fem_b200.core import gather_scatter as gs
from swirl_fem_b200.core.interpolation import Nodes1D
from swirl_fem_b200.core.interpolation import NodeType

GLL = NodeType.GAUSS_LOBATTO_LEGENDRE


class CpuHaloPlan(HaloPlan):
  """HaloPlan with the two device kernels replaced by CPU index ops."""

  def _pack(self, u, idx, buf):
    buf.copy_(u[idx.long()])

  def _unpack_add(self, u, idx, buf):
    u[idx.long()] += buf

  def _unpack_canonical(self, u, recv):
    dofs, row_ptr, src = self.canonical_csr()
    own = u.clone()
    for i, d in enumerate(dofs):
      acc = torch.zeros((), dtype=u.dtype)
      for j in range(row_ptr[i], row_ptr[i + 1]):
        acc = acc + (own[d] if src[j] < 0 else recv[src[j]])
      u[d] = acc


def _free_port():
  with socket.socket() as s:
    s.bind(('127.0.0.1', 0))
    return s.getsockname()[1]


def _worker_known_answers(rank, world, port, node_indices, u_all, expected):
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  dist.init_process_group('gloo', rank=rank, world_size=world)
  try:
    base = HaloPlan.from_node_indices(node_indices, rank)
    plan = CpuHaloPlan(**{f: getattr(base, f) for f in
                          ('rank', 'world', 'peers', 'local_idx', 'owned')})
    u = torch.tensor(u_all[rank], dtype=torch.float64)
    out = plan.exchange(u)
    np.testing.assert_array_equal(out.numpy(), expected[rank])
  finally:
    dist.destroy_process_group()


def _run(fn, world, *args):
  mp.spawn(fn, args=(world, _free_port()) + args, nprocs=world, join=True)


def test_exchange_known_answers_4_ranks():
  # gather_scatter_test.py:157-177
  ni = np.array([[0, 1, 2], [2, 3, 4], [4, 5, 6], [6, 7, 8]], dtype=np.int32)
  u = np.arange(12.).reshape(4, 3)
  expected = np.array([[0, 1, 5], [5, 4, 11], [11, 7, 17], [17, 10, 11]],
                      dtype=np.float64)
  _run(_worker_known_answers, 4, ni, u, expected)
  # periodic 0 <-> 8 (gather_scatter_test.py:200-223)
  nip = gs.get_unique_node_indices(ni, periodic_links=np.array([[[0], [8]]]))
  expected = np.array([[11, 1, 5], [5, 4, 11], [11, 7, 17], [17, 10, 11]],
                      dtype=np.float64)
  _run(_worker_known_answers, 4, nip, u, expected)


def test_exchange_doubly_periodic_4_ranks():
  # gather_scatter_test.py:225-263: every global node receives four ones
  ni = np.array([[0, 1, 3, 4], [1, 2, 4, 5], [3, 4, 6, 7], [4, 5, 7, 8]],
                dtype=np.int32)
  links = np.array([[[0, 1], [6, 7]], [[1, 2], [7, 8]], [[0, 3], [2, 5]],
                    [[3, 6], [5, 8]]], dtype=np.int32)
  nip = gs.get_unique_node_indices(ni, periodic_links=links)
  _run(_worker_known_answers, 4, nip, np.ones((4, 4)), 4 * np.ones((4, 4)))


def _worker_block(rank, world, port, ndim, ne, n1d):
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  dist.init_process_group('gloo', rank=rank, world_size=world)
  try:
    blk = part.block_partition(ne, ndim, Nodes1D.create(n1d, GLL), rank, world)
    gathered = [None] * world
    dist.all_gather_object(gathered, np.sort(blk.interface_global))
    base = part.halo_plan_from_interfaces(
        rank, blk.interface_local, blk.interface_global, gathered,
        blk.premesh.num_nodes)
    plan = CpuHaloPlan(**{f: getattr(base, f) for f in
                          ('rank', 'world', 'peers', 'local_idx', 'owned')})
    # a global field sampled at the local nodes: f(x) = 1 + x0 + 2 x1 (+3 x2)
    x = blk.premesh.node_coords
    f = 1.0 + sum((k + 1) * x[:, k] for k in range(ndim))
    u = torch.tensor(f, dtype=torch.float64)
    out = plan.exchange(u).numpy()
    # multiplicity of every local node over ranks
    mult = plan.exchange(torch.ones_like(u)).numpy()
    np.testing.assert_allclose(out, mult * f, rtol=1e-13)
    # canonical unpack: replicated dofs are BITWISE identical on all holders
    # (values chosen so that the sum depends on the association order)
    rng = np.random.default_rng(7)
    noisy = torch.tensor(f * (1.0 + 1e-3 * rng.standard_normal(len(f))) *
                         (1.0 + 0.37 * rank), dtype=torch.float64)
    summed = plan.exchange(noisy).numpy()
    mine = dict(zip(blk.interface_global.tolist(),
                    summed[blk.interface_local].tolist()))
    everyone = [None] * world
    dist.all_gather_object(everyone, mine)
    for other in everyone:
      for gid, val in other.items():
        if gid in mine:
          assert mine[gid] == val, (gid, mine[gid], val)
    # owned weights count every global dof exactly once
    total = torch.tensor([float(plan.owned.sum())], dtype=torch.float64)
    dist.all_reduce(total)
    assert int(total.item()) == blk.num_global_dofs
    # interface nodes have multiplicity >= 2, interior ones 1
    interior = np.ones(len(f), dtype=bool)
    interior[blk.interface_local] = False
    assert (mult[interior] == 1).all()
    assert (mult[blk.interface_local] >= 2).all()
    assert mult.max() <= 2 ** ndim
    # elements that touch an interface dof are stored first (overlap split)
    is_iface = np.zeros(len(f), dtype=bool)
    is_iface[blk.interface_local] = True
    touching = is_iface[blk.premesh.elements].any(axis=1)
    ni = blk.num_interface_elements
    assert ni % 4 == 0 or ni == blk.premesh.num_elements
    assert not touching[ni:].any()
    assert touching[:ni].sum() == touching.sum()
    # Dirichlet flags = global boundary only
    onb = (np.abs(np.abs(x) - 1.0) < 1e-12).any(axis=1)
    np.testing.assert_array_equal(blk.dirichlet, onb)
  finally:
    dist.destroy_process_group()


@pytest.mark.parametrize('world,ndim,ne,n1d', [(2, 2, 4, 4), (2, 3, 2, 3),
                                               (4, 3, 2, 4), (4, 2, 4, 3)])
def test_block_partition_halo(world, ndim, ne, n1d):
  _run(_worker_block, world, ndim, ne, n1d)


def test_plan_from_reference_partition_golden():
  """HaloPlan lists agree with the reference's exchange_gather_indices."""
  from tests.conftest import load_golden
  g = load_golden('partition')
  for name in ('q2_ne4_p2_2x2', 'h3_ne2_p2_2x2x2', 'q2_ne4_p3_2x1'):
    ni = g[name + '/node_indices']
    gi = g[name + '/exchange_gather_indices']
    world = len(ni)
    for r in range(world):
      plan = HaloPlan.from_node_indices(ni, r)
      shared_local = np.unique(np.concatenate(
          [plan.local_idx[q] for q in plan.peers])) if plan.peers else []
      np.testing.assert_array_equal(np.sort(gi[r][gi[r] >= 0]), shared_local)


@pytest.mark.parametrize('case', [(3, 4, 5, 8), (3, 2, 8, 2), (3, 4, 4, 4),
                                  (2, 8, 4, 4)])
def test_p2p_tables_route_every_dof_to_its_canonical_slot(case):
  """Host arithmetic of the peer-memory exchange: every (rank, peer, shared
  dof) send entry lands in a distinct slot of the peer's receive buffer, and
  the slot is the one the peer's canonical sum reads for that dof."""
  from swirl_fem_b200.communication import partition as part
  from swirl_fem_b200.communication.halo import p2p_region_layout, p2p_tables
  from swirl_fem_b200.core.interpolation import Nodes1D, NodeType
  ndim, ne, n1d, world = case
  g = Nodes1D.create(n1d, NodeType.GAUSS_LOBATTO_LEGENDRE)
  blks = [part.block_partition(ne, ndim, g, r, world) for r in range(world)]
  gathered = [np.sort(b.interface_global) for b in blks]
  plans = [part.halo_plan_from_interfaces(
      r, b.interface_local, b.interface_global, gathered, b.premesh.num_nodes)
           for r, b in enumerate(blks)]
  all_splits = [p.splits() for p in plans]
  esz = 8
  flag_bytes, stride, total = p2p_region_layout(
      world, max(sum(s) for s in all_splits), esz)
  assert flag_bytes % 256 == 0 and stride % 256 == 0
  assert total == flag_bytes + 2 * stride
  bases = {q: (q + 1) << 32 for q in range(world)}
  gid = [dict(zip(b.interface_local.tolist(), b.interface_global.tolist()))
         for b in blks]
  recv = {q: np.full(sum(all_splits[q]), -1, dtype=np.int64)
          for q in range(world)}
  for r, p in enumerate(plans):
    dst, flag_addr = p2p_tables(
        r, p.peers, {q: len(v) for q, v in p.local_idx.items()}, all_splits,
        bases, esz)
    assert [int(a) for a in flag_addr] == [bases[q] + 8 * r for q in p.peers]
    off = 0
    for q in p.peers:
      n = len(p.local_idx[q])
      slots = (dst[off:off + n].astype(np.int64) - bases[q] - flag_bytes) // esz
      assert slots.min() >= 0 and slots.max() < len(recv[q])
      assert (recv[q][slots] == -1).all()
      recv[q][slots] = [gid[r][int(l)] for l in p.local_idx[q]]
      off += n
  for q, p in enumerate(plans):
    assert (recv[q] >= 0).all()
    dofs, row_ptr, src = p.canonical_csr()
    for i, d in enumerate(dofs):
      for j in range(row_ptr[i], row_ptr[i + 1]):
        if src[j] >= 0:
          assert recv[q][src[j]] == gid[q][int(d)]


def test_p2p_tables_reject_inconsistent_counts():
  from swirl_fem_b200.communication.halo import p2p_tables
  with pytest.raises(ValueError, match='disagree'):
    p2p_tables(0, [1], {1: 5}, [[0, 5], [4, 0]], {1: 1 << 20}, 8)


# -- crystal router and prefix scan --------------------------


def _init(rank, world, port):
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  dist.init_process_group('gloo', rank=rank, world_size=world)


def _crystal_worker(rank, world, port, outdir):
  """crystal_router_test.py:36-79 with one process per axis index."""
  from swirl_fem_b200.communication.crystal_router import crystal_router_setup
  _init(rank, world, port)
  try:
    m = 7
    rng = np.random.RandomState(seed=2)
    num = rng.randint(low=m // 2, high=m + 1, size=(world,)).astype(np.int32)
    target = rng.randint(low=0, high=world, size=(world, m)).astype(np.int32)
    data = rng.randint(100, size=(world, m)).astype(np.int32)
    mask = np.arange(m) < num[:, None]
    in_src = np.where(mask, np.arange(world)[:, None], -1)
    target = np.where(mask, target, 0)
    crystal = crystal_router_setup(None)
    n_out, out, source = crystal(int(num[rank]), torch.as_tensor(data[rank]),
                                 torch.as_tensor(target[rank]))
    lexsorted = lambda *a: np.array([*a])[:, np.lexsort([*a])]  # noqa: E731
    ft, fs, fd = (target.flatten()[mask.flatten()],
                  in_src.flatten()[mask.flatten()],
                  data.flatten()[mask.flatten()])
    sel = ft == rank
    np.testing.assert_array_equal(
        lexsorted(fd[sel], fs[sel]),
        lexsorted(out.numpy()[:n_out], source.numpy()[:n_out]))
    # pytree payload + no source
    n2, out2 = crystal(int(num[rank]),
                       {'a': torch.as_tensor(data[rank]),
                        'b': torch.as_tensor(data[rank]).double()[:, None] * 2},
                       torch.as_tensor(target[rank]), return_source=False)
    assert n2 == n_out
    np.testing.assert_array_equal(np.sort(out2['a'].numpy()),
                                  np.sort(out.numpy()[:n_out]))
    np.testing.assert_array_equal(out2['b'].numpy()[:, 0],
                                  2.0 * out2['a'].numpy())
    # second invocation restores the data up to ordering
    pad = lambda t: torch.cat([t, t.new_zeros(world * m - len(t))])  # noqa: E731
    n_back, back, src_back = crystal(n_out, pad(out), pad(source))
    np.testing.assert_array_equal(
        lexsorted(data[rank, :num[rank]], target[rank, :num[rank]]),
        lexsorted(back.numpy()[:n_back], src_back.numpy()[:n_back]))
    with pytest.raises(ValueError):
      crystal(1, torch.zeros(3), torch.full((3,), world, dtype=torch.int32))
    open(os.path.join(outdir, f'ok{rank}'), 'w').close()
  finally:
    dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 4])
def test_crystal_router_semantics(world, tmp_path):
  _run(_crystal_worker, world, str(tmp_path))
  assert len(os.listdir(tmp_path)) == world


def _pscan_worker(rank, world, port, outdir):
  """pscan_test.py:60-95 with one process per axis index."""
  from swirl_fem_b200.communication.pscan import preduce, pscan
  _init(rank, world, port)
  try:
    for x in (np.arange(world), np.flip(np.arange(world)).copy()):
      for op in ('add', 'multiply', 'maximum', 'minimum', 'bitwise_and',
                 'bitwise_or', 'bitwise_xor'):
        np_op = getattr(np, op)
        mine = torch.as_tensor(x[rank:rank + 1])
        exclusive = pscan(mine, op)
        inclusive = np_op.accumulate(x)
        assert np_op(exclusive.numpy()[0], x[rank]) == inclusive[rank], op
        ex2, red = pscan(mine, op, reduction=True)
        assert torch.equal(ex2, exclusive)
        assert red.numpy()[0] == np_op.reduce(x)
        assert preduce(mine, op).numpy()[0] == np_op.reduce(x)
    # pytree, floating point, multi-entry leaves: global numbering offsets
    tree = {'count': torch.tensor([rank + 1, 2 * rank]),
            'w': torch.tensor([0.5 * (rank + 1)], dtype=torch.float64)}
    scan, red = pscan(tree, torch.add, reduction=True)
    assert scan['count'].tolist() == [rank * (rank + 1) // 2,
                                      rank * (rank - 1)]
    assert red['count'].tolist() == [world * (world + 1) // 2,
                                     world * (world - 1)]
    assert float(scan['w']) == 0.5 * rank * (rank + 1) / 2
    with pytest.raises(ValueError):
      pscan(torch.zeros(1), 'subtract')
    open(os.path.join(outdir, f'ok{rank}'), 'w').close()
  finally:
    dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_pscan_and_preduce(world, tmp_path):
  _run(_pscan_worker, world, str(tmp_path))
  assert len(os.listdir(tmp_path)) == world


# -- general (non-block) partitions: partitioner -> index builders -> halo plan ----


def _worker_general_partition(rank, world, port, seed):
  """An 'unstructured' quad mesh (shuffled elements, rotated vertex listings)
  partitioned by `common.mesh_partitioner`, localised by the reference's index
  builders (`Premesh.partition_host`: group_by_partitions / get_local_elements,
  gather_scatter.py:355-445) and exchanged with the pairwise halo plan: the
  result equals the unpartitioned QQ^T semantics (every copy of a global dof
  holds the sum over all copies)."""
  from swirl_fem_b200.common import mesh_partitioner
  from swirl_fem_b200.core.mesh_refiner import refine_premesh
  from tests import helpers
  _init(rank, world, port)
  try:
    pm = helpers.shuffled(helpers.unit_cube_mesh(6, ndim=2, a=-1., b=1.), seed)
    pm = mesh_partitioner.partition(pm, world)
    refined = refine_premesh(pm, Nodes1D.create(4, GLL))
    host = refined.partition_host()
    nidx = host['node_indices']                      # (P, n_max) global ids
    base = HaloPlan.from_node_indices(nidx, rank)
    plan = CpuHaloPlan(**{f: getattr(base, f) for f in
                          ('rank', 'world', 'peers', 'local_idx', 'owned')})
    mine = nidx[rank][nidx[rank] != -1]
    rng = np.random.default_rng(100 + rank)
    u = rng.standard_normal(len(mine))
    out = plan.exchange(torch.tensor(u, dtype=torch.float64)).numpy()
    # oracle: scatter every rank's values to the global numbering and sum
    everyone = [None] * world
    dist.all_gather_object(everyone, (mine, u))
    total = np.zeros(refined.num_nodes)
    for ids, vals in everyone:
      np.add.at(total, ids, vals)
    np.testing.assert_allclose(out, total[mine], rtol=1e-13, atol=1e-13)
    # ownership weights count every global dof once
    count = torch.tensor([float(plan.owned.sum())], dtype=torch.float64)
    dist.all_reduce(count)
    assert int(count.item()) == refined.num_nodes
  finally:
    dist.destroy_process_group()


@pytest.mark.parametrize('world,seed', [(2, 3), (4, 8), (3, 5)])
def test_general_partition_halo(world, seed):
  _run(_worker_general_partition, world, seed)
