#!/usr/bin/env python
"""Multi-GPU parity check (run under torchrun, one rank per GPU).

Partitioned apply + halo exchange and the distributed CG are compared with the
UNPARTITIONED result on the same global mesh (computed redundantly on every
rank's GPU), the parity target SURVEY section 5 names for N > 1.

  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29511 tools/check_multi_gpu.py
"""

import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from swirl_fem_b200.common.premesh_commons import unit_cube_mesh  # noqa: E402
from swirl_fem_b200.communication import partition as part  # noqa: E402
from swirl_fem_b200.communication.dist_cg import distributed_cg  # noqa: E402
from swirl_fem_b200.core.interpolation import Nodes1D, NodeType, Quadrature1D  # noqa: E402
from swirl_fem_b200.core.mesh import Mesh  # noqa: E402
from swirl_fem_b200.core.mesh_refiner import refine_premesh  # noqa: E402
from swirl_fem_b200.core.operator import FusedOperator  # noqa: E402
from swirl_fem_b200.core.operator import JacobiPreconditioner  # noqa: E402
from swirl_fem_b200.linalg.cg import cg  # noqa: E402


def deform(x):
  ndim = x.shape[-1]
  perm = np.roll(np.arange(ndim), 1)
  return x + 0.08 * np.sin(np.pi * x[:, perm]) * (1 - x ** 2)


def field(x):
  return np.cos(1.3 * x[:, 0]) * (1.0 + 0.5 * x[:, -1]) + 0.2 * x[:, 1] ** 2


def main():
  rank = int(os.environ['RANK'])
  world = int(os.environ['WORLD_SIZE'])
  local_rank = int(os.environ.get('LOCAL_RANK', rank))
  torch.cuda.set_device(local_rank)
  dev = torch.device('cuda', local_rank)
  dist.init_process_group('nccl', device_id=dev)
  gll = NodeType.GAUSS_LOBATTO_LEGENDRE
  ok = True
  for ndim, ne, n1d in ((3, 4, 5), (2, 8, 4), (3, 2, 8)):
    grid1d = Nodes1D.create(n1d, gll)
    quad = Quadrature1D.create_from_nodes_1d(grid1d)
    # ---- partitioned
    blk = part.block_partition(ne, ndim, grid1d, rank, world)
    gathered = [None] * world
    dist.all_gather_object(gathered, np.sort(blk.interface_global))
    halo = part.halo_plan_from_interfaces(
        rank, blk.interface_local, blk.interface_global, gathered,
        blk.premesh.num_nodes)
    x0 = blk.premesh.node_coords           # undeformed: global identity key
    mesh = Mesh.create(deform(x0), blk.premesh.elements, gridpoints_1d=grid1d,
                       device=dev)
    op = FusedOperator(mesh, quad, dirichlet_mask=blk.dirichlet, with_mass=True)
    u = torch.as_tensor(field(x0)).to(dev)
    y = op.apply(u, lam=0.3, mu=1.0)
    halo.exchange_(y)
    # overlapped form (interface elements first, exchange on a side stream)
    y2 = torch.empty_like(y)
    dot2 = torch.zeros((), dtype=torch.float64, device=dev)
    op.apply_partitioned(u, y2, halo, blk.num_interface_elements, lam=0.3,
                         mu=1.0, dot_out=dot2, overlap=True)
    torch.cuda.synchronize()
    assert float((y2 - y).abs().max()) <= 1e-13 * float(y.abs().max()), (
        'overlapped apply differs from apply + exchange')
    # peer-memory path (CUDA IPC + NVLink stores): fused apply + in-kernel
    # push (3-D) / apply + push kernel (2-D), then wait + canonical sum
    p2p = os.environ.get('SFEM_HALO', 'p2p') == 'p2p' and halo.enable_p2p(
        torch.float64, dev)
    if p2p:
      for rep in range(3):   # both epoch parities
        y3 = torch.empty_like(y)
        dot3 = torch.zeros((), dtype=torch.float64, device=dev)
        op.apply_partitioned(u, y3, halo, blk.num_interface_elements, lam=0.3,
                             mu=1.0, dot_out=dot3)
        torch.cuda.synchronize()
        assert not halo.p2p_timed_out(dev), 'peer flag wait timed out'
        assert float((y3 - y).abs().max()) <= 1e-13 * float(y.abs().max()), (
            f'peer-memory apply differs from apply + NCCL exchange (rep {rep}):'
            f' {float((y3 - y).abs().max())}')
        assert abs(float(dot3) - float(dot2)) <= 1e-12 * abs(float(dot2))
    if rank == 0:
      print(f'  halo path: {"peer memory (P2P)" if p2p else "NCCL"}', flush=True)
    # ---- unpartitioned reference on this GPU
    ref = refine_premesh(unit_cube_mesh(ne, ndim=ndim, a=-1., b=1.), grid1d)
    gmesh = Mesh.create(deform(ref.node_coords), ref.elements,
                        gridpoints_1d=grid1d, device=dev)
    bmask = ref.finalize_host()['physical_masks']['boundary']
    gop = FusedOperator(gmesh, quad, dirichlet_mask=bmask, with_mass=True)
    gu = torch.as_tensor(field(ref.node_coords)).to(dev)
    gy = gop.apply(gu, lam=0.3, mu=1.0).cpu().numpy()
    # match local nodes to global nodes through the undeformed coordinates
    key = lambda c: np.round(c * 1e9).astype(np.int64)  # noqa: E731
    gk = key(ref.node_coords)
    lk = key(x0)
    order = np.lexsort(gk.T[::-1])
    sorted_gk = gk[order]
    rec = [('', np.int64)] * ndim
    view_g = np.ascontiguousarray(sorted_gk).view(rec).ravel()
    view_l = np.ascontiguousarray(lk).view(rec).ravel()
    l2g = order[np.searchsorted(view_g, view_l)]
    assert np.array_equal(gk[l2g], lk)
    err = np.abs(y.cpu().numpy() - gy[l2g]).max() / np.abs(gy).max()
    # ---- CG
    rhs = torch.where(torch.as_tensor(blk.dirichlet, device=dev), 0.0,
                      1.0).double()
    b_loc = op.apply(rhs * 0 + 1.0, lam=1.0, mu=0.0)
    halo.exchange_(b_loc)
    diag = op.diag()
    halo.exchange_(diag)
    minv = torch.where(diag != 0, 1.0 / diag, torch.zeros_like(diag))
    xs, info = distributed_cg(
        op, halo, b_loc, tol=1e-9, minv=minv, check_every=7,
        num_interface_elements=blk.num_interface_elements)
    # the same solve with the dot products all-reduced over peer memory
    from swirl_fem_b200.communication.scalar_exchange import ScalarExchange  # noqa: E402
    # (opt-in: SFEM_CHECK_SCALARS=1)
    sx = (ScalarExchange.create(dev)
          if os.environ.get('SFEM_CHECK_SCALARS') == '1' else None)
    if sx is not None:
      xs2, info2 = distributed_cg(
          op, halo, b_loc, tol=1e-9, minv=minv, check_every=7,
          num_interface_elements=blk.num_interface_elements,
          scalar_exchange=sx)
      assert not sx.timed_out(), 'scalar exchange timed out'
      assert abs(info2['num_iterations'] - info['num_iterations']) <= 1, (
          info2['num_iterations'], info['num_iterations'])
      assert float((xs2 - xs).abs().max()) <= 1e-9 * float(xs.abs().max())
      if rank == 0:
        print(f'  CG scalars over peer memory: {info2["num_iterations"]} '
              f'iterations (NCCL: {info["num_iterations"]})', flush=True)
    gb = gop.apply(torch.ones_like(gu), lam=1.0, mu=0.0)
    gx, ginfo = cg(gop.bind(0.0, 1.0), gb, tol=1e-9,
                   M=JacobiPreconditioner(gop.jacobi_minv()))
    berr = float((b_loc.cpu() - gb.cpu()[l2g]).abs().max() / gb.abs().max())
    xerr = float((xs.cpu() - gx.cpu()[l2g]).abs().max() / gx.abs().max())
    it_ok = abs(info['num_iterations'] - ginfo['num_iterations']) <= 1
    good = err < 1e-12 and berr < 1e-12 and xerr < 1e-7 and it_ok
    ok = ok and good
    halo.disable_p2p()
    print(f'[rank {rank}/{world}] {ndim}-D ne={ne} N={n1d}: apply err {err:.1e} '
          f'rhs err {berr:.1e} cg x err {xerr:.1e} iters '
          f'{info["num_iterations"]} vs {ginfo["num_iterations"]} '
          f'{"OK" if good else "FAIL"}', flush=True)
  flag = torch.tensor([1.0 if ok else 0.0], device=dev)
  dist.all_reduce(flag, op=dist.ReduceOp.MIN)
  dist.barrier()
  dist.destroy_process_group()
  sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == '__main__':
  main()
