#!/usr/bin/env python
"""`swirl_fem.communication` contracts on real GPUs over NCCL (torchrun, one rank
per GPU): the sparse all-to-all (crystal_router_test.py:36-79), the exclusive
prefix scan / reduction (pscan_test.py:60-127) and the general-partition halo
(mesh_partitioner -> index builders of gather_scatter.py:355-445 -> pairwise
halo plan) with CUDA tensors -- the same checks the CPU suite runs over gloo
(tests/test_distributed_cpu.py), here with backend 'nccl', the CUDA pack /
canonical-unpack kernels and the peer-memory (CUDA IPC) push.

  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29517 tools/check_comm_nccl.py
"""

import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from swirl_fem_b200.common import mesh_partitioner  # noqa: E402
from swirl_fem_b200.communication.crystal_router import crystal_router_setup  # noqa: E402
from swirl_fem_b200.communication.halo import HaloPlan  # noqa: E402
from swirl_fem_b200.communication.pscan import preduce, pscan  # noqa: E402
from swirl_fem_b200.core.interpolation import Nodes1D, NodeType  # noqa: E402
from swirl_fem_b200.core.mesh_refiner import refine_premesh  # noqa: E402
from tests import helpers  # noqa: E402


def check_crystal_router(rank, world, dev):
  m = 7
  rng = np.random.RandomState(seed=2)
  num = rng.randint(low=m // 2, high=m + 1, size=(world,)).astype(np.int32)
  target = rng.randint(low=0, high=world, size=(world, m)).astype(np.int32)
  data = rng.randint(100, size=(world, m)).astype(np.int32)
  mask = np.arange(m) < num[:, None]
  in_src = np.where(mask, np.arange(world)[:, None], -1)
  target = np.where(mask, target, 0)
  crystal = crystal_router_setup(None)
  n_out, out, source = crystal(int(num[rank]),
                               torch.as_tensor(data[rank]).to(dev),
                               torch.as_tensor(target[rank]).to(dev))
  assert out.is_cuda and source.is_cuda
  lexsorted = lambda *a: np.array([*a])[:, np.lexsort([*a])]  # noqa: E731
  ft, fs, fd = (target.flatten()[mask.flatten()],
                in_src.flatten()[mask.flatten()],
                data.flatten()[mask.flatten()])
  sel = ft == rank
  np.testing.assert_array_equal(
      lexsorted(fd[sel], fs[sel]),
      lexsorted(out.cpu().numpy()[:n_out], source.cpu().numpy()[:n_out]))
  # second invocation restores the data up to ordering
  pad = lambda t: torch.cat([t, t.new_zeros(world * m - len(t))])  # noqa: E731
  n_back, back, src_back = crystal(n_out, pad(out), pad(source))
  np.testing.assert_array_equal(
      lexsorted(data[rank, :num[rank]], target[rank, :num[rank]]),
      lexsorted(back.cpu().numpy()[:n_back], src_back.cpu().numpy()[:n_back]))
  return n_out


def check_pscan(rank, world, dev):
  for x in (np.arange(world), np.flip(np.arange(world)).copy()):
    for op in ('add', 'multiply', 'maximum', 'minimum', 'bitwise_and',
               'bitwise_or', 'bitwise_xor'):
      np_op = getattr(np, op)
      mine = torch.as_tensor(x[rank:rank + 1]).to(dev)
      exclusive = pscan(mine, op)
      inclusive = np_op.accumulate(x)
      assert np_op(exclusive.cpu().numpy()[0], x[rank]) == inclusive[rank], op
      _, red = pscan(mine, op, reduction=True)
      assert red.cpu().numpy()[0] == np_op.reduce(x)
      assert preduce(mine, op).cpu().numpy()[0] == np_op.reduce(x)
  tree = {'count': torch.tensor([rank + 1, 2 * rank], device=dev),
          'w': torch.tensor([0.5 * (rank + 1)], dtype=torch.float64,
                            device=dev)}
  scan, red = pscan(tree, torch.add, reduction=True)
  assert scan['count'].tolist() == [rank * (rank + 1) // 2, rank * (rank - 1)]
  assert red['count'].tolist() == [world * (world + 1) // 2,
                                   world * (world - 1)]


def check_general_partition_halo(rank, world, dev, seed):
  """'Unstructured' quads (shuffled, re-oriented), METIS-stand-in partition,
  reference index builders, then the halo exchange on the GPU: NCCL
  all_to_all path and the peer-memory path, both against the all-gathered
  sum (every copy of a global dof holds the sum over all copies)."""
  gll = NodeType.GAUSS_LOBATTO_LEGENDRE
  pm = helpers.shuffled(helpers.unit_cube_mesh(6, ndim=2, a=-1., b=1.), seed)
  pm = mesh_partitioner.partition(pm, world)
  refined = refine_premesh(pm, Nodes1D.create(4, gll))
  nidx = refined.partition_host()['node_indices']        # (P, n_max) global ids
  plan = HaloPlan.from_node_indices(nidx, rank)
  mine = nidx[rank][nidx[rank] != -1]
  u = np.random.default_rng(100 + rank).standard_normal(len(mine))
  everyone = [None] * world
  dist.all_gather_object(everyone, (mine, u))
  total = np.zeros(refined.num_nodes)
  for ids, vals in everyone:
    np.add.at(total, ids, vals)
  errs = []
  for dtype, tol in ((torch.float64, 1e-13), (torch.float32, 1e-6)):
    ud = torch.as_tensor(u).to(device=dev, dtype=dtype)
    out = plan.exchange(ud)                               # NCCL all_to_all
    errs.append(float(np.abs(out.cpu().numpy() - total[mine]).max()))
    assert errs[-1] <= tol * max(1.0, np.abs(total).max()), errs
    if plan.enable_p2p(dtype, dev):                       # CUDA IPC + NVLink
      for _ in range(3):                                  # both epoch parities
        out2 = plan.exchange(ud)
        torch.cuda.synchronize()
        assert not plan.p2p_timed_out(dev)
        # canonical (rank-ordered) sums: both paths agree bitwise
        assert torch.equal(out2, out)
      plan.disable_p2p()
  count = torch.tensor([float(plan.owned.sum())], dtype=torch.float64,
                       device=dev)
  dist.all_reduce(count)
  assert int(count.item()) == refined.num_nodes
  return max(errs)


def main():
  rank = int(os.environ['RANK'])
  world = int(os.environ['WORLD_SIZE'])
  local_rank = int(os.environ.get('LOCAL_RANK', rank))
  torch.cuda.set_device(local_rank)
  dev = torch.device('cuda', local_rank)
  dist.init_process_group('nccl', device_id=dev)
  ok = True
  try:
    n_out = check_crystal_router(rank, world, dev)
    check_pscan(rank, world, dev)
    err = check_general_partition_halo(rank, world, dev, seed=3 + world)
    print(f'[rank {rank}/{world}] nccl: crystal router OK ({n_out} received), '
          f'pscan/preduce OK, general-partition halo OK (max err {err:.1e}, '
          'NCCL and peer-memory paths bitwise equal)', flush=True)
  except Exception as e:  # pylint: disable=broad-except
    ok = False
    print(f'[rank {rank}/{world}] FAIL: {type(e).__name__}: {e}', flush=True)
    import traceback
    traceback.print_exc()
  flag = torch.tensor([1.0 if ok else 0.0], device=dev)
  dist.all_reduce(flag, op=dist.ReduceOp.MIN)
  dist.barrier()
  dist.destroy_process_group()
  sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == '__main__':
  main()
