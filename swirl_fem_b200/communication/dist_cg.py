"""Element-partitioned (P)CG: one process per GPU.

Same recurrence and stopping rule as `swirl_fem/linalg/cg.py:54-97` on a
device-resident state.  Default path (peer-memory halo + `ScalarExchange`),
THREE launches per iteration and rank, no NCCL call (`sfem_cg_iterate`):

  apply kernel     local block, interface elements first (its CTAs signal
                   when they are through them), partial p.Ap in the epilogue
                   (element-wise partial sums need no ownership weights)
  companion kernel runs NEXT to the apply in the warp slots it leaves free
                   (`sfem_halo_wait_unpack`, exchange mode 3): pushes the
                   shared dofs to the peers over NVLink, waits for theirs,
                   canonical sum -- all under the interior elements
  step kernel      all-reduce of p.Ap over peer memory (warp 0 of CTA 0),
                   x, r update with partial r.z over the OWNED dofs, grid
                   arrival, all-reduce of r.z (last CTA), scalar advance +
                   convergence flag, direction update, zero fill of the next
                   apply's shared-dof prefix

Fallback (`SFEM_HALO=nccl`, or peer mapping unavailable): the building blocks
`sfem_cg_update/_direction/_advance` with two NCCL all-reduces in between.

There is no host synchronisation inside the loop: the convergence flag is read
every `check_every` iterations (it is identical on all ranks because it is
computed from all-reduced scalars).  A peer-memory wait that times out is
FATAL (`SwirlB200Error` at the next read).  The reference has no distributed
CG; its hook is `dot_fn` (`cg.py:26-31`), and the parity target is the
unpartitioned solve on the same global mesh.
"""

from __future__ import annotations

import ctypes

import torch
import torch.distributed as dist

from swirl_fem_b200 import _lib


def distributed_cg(op, halo, b, x0=None, *, tol=1e-5, atol=0.0, maxiter=None,
                   minv=None, lam=0.0, mu=1.0, check_every=16, group=None,
                   num_interface_elements=None, scalar_exchange=None):
  """Solves `A x = b` on an element-partitioned mesh.

  Args:
    op: this rank's `FusedOperator` (Dirichlet mask = global boundary only).
    halo: this rank's `HaloPlan` (None for a single rank).
    b: right-hand side, consistent across ranks at shared dofs.
    x0: initial guess (consistent), default zeros.
    minv: inverse diagonal of the *assembled* operator (exchange the local
      diagonal before inverting), or None.
    maxiter: default 10 * (global number of dofs).
    scalar_exchange: a `communication.scalar_exchange.ScalarExchange`: the two
      dot products per iteration are all-reduced over peer memory inside the
      fused step kernel (sums in rank order, bitwise identical on all ranks).
      None: one is created (and cached on `halo`) when the halo runs over peer
      memory; NCCL otherwise.
  Returns:
    `(x, {'residual', 'num_iterations'})` as `linalg.cg.cg`.
  """
  _lib.require_cuda(b)
  lib = _lib.lib()
  dtype = op.dtype
  dev = b.device
  code = _lib.dtype_code(dtype)
  b = b.to(dtype).contiguous()
  n = b.numel()
  x = torch.zeros_like(b) if x0 is None else x0.to(dtype).clone().contiguous()
  r = torch.empty_like(b)
  p = torch.empty_like(b)
  ap = torch.empty_like(b)
  state = torch.zeros(int(lib.sfem_cg_state_bytes()) // 8, dtype=torch.float64,
                      device=dev)
  world = dist.get_world_size(group) if (halo is not None and
                                         dist.is_initialized()) else 1
  owned = None
  if halo is not None:
    owned = halo.owned_mask(dev)
  if minv is not None:
    minv = minv.to(dtype).contiguous()
  if maxiter is None:
    total = torch.tensor(
        [float(n if halo is None else int(halo.owned.sum()))],
        dtype=torch.float64, device=dev)
    if world > 1:
      dist.all_reduce(total, group=group)
    maxiter = 10 * int(total.item())
  stream = _lib.stream_ptr(dev)
  p2p = halo is not None and halo.p2p_handle(b) is not None
  if p2p and world > 1:
    # ranks may arrive seconds apart (set-up, first-use module loading): the
    # bounded device-side waits must only ever see communication time
    dist.barrier(group=group)
    if scalar_exchange is None:
      scalar_exchange = halo.scalar_exchange(dev, group)
  single = halo is None or not halo.peers
  fused = single or (p2p and scalar_exchange is not None
                     and scalar_exchange.world > 1
                     and num_interface_elements is not None)

  def check_timeouts():
    if p2p and halo.p2p_timed_out(dev):
      raise _lib.SwirlB200Error(
          'halo exchange: a peer never raised its flag (4 s device-side '
          'timeout); the result is invalid')
    if scalar_exchange is not None and scalar_exchange.timed_out():
      raise _lib.SwirlB200Error(
          'scalar all-reduce: a peer never answered (4 s device-side '
          'timeout); the result is invalid')

  def apply(src, dot_out):
    if halo is not None and num_interface_elements is not None:
      # interface elements first; their exchange overlaps the interior launch
      op.apply_partitioned(src, ap, halo, num_interface_elements, lam=lam,
                           mu=mu, dot_out=dot_out)
      return
    op.apply(src, lam=lam, mu=mu, out=ap, dot_out=dot_out)
    if halo is not None:
      halo.exchange_(ap)

  def allreduce(view):
    if not single:
      if scalar_exchange is not None:
        scalar_exchange.allreduce_(view)
      else:
        dist.all_reduce(view, group=group)

  with torch.cuda.device(dev):
    apply(x, None)
    _lib._check(lib.sfem_cg_init(
        code, n, _lib.ptr(b), _lib.ptr(ap), _lib.ptr(minv), _lib.ptr(owned),
        _lib.ptr(r), _lib.ptr(p), _lib.ptr(state), float(tol), float(atol),
        int(maxiter), stream), 'sfem_cg_init')
    allreduce(state[2:4])  # gamma, b.b
    _lib._check(lib.sfem_cg_init_finish(_lib.ptr(state), stream),
                'sfem_cg_init_finish')
    info = _lib.CgInfo()
    done = ctypes.c_int32(0)
    pap = state[0:1]
    gnew = state[1:2]
    hp = None if single else halo.p2p_handle(b)
    sxh = None if single else scalar_exchange.handle
    while True:
      _lib._check(lib.sfem_cg_read(_lib.ptr(state), ctypes.byref(info),
                                   ctypes.byref(done), stream), 'sfem_cg_read')
      check_timeouts()
      if done.value == 2:
        raise _lib.SwirlB200Error('distributed CG: a peer-memory wait timed '
                                  'out inside the step kernel')
      if done.value:
        break
      iters = int(min(check_every, max(1, maxiter - info.num_iterations)))
      if fused:
        _lib._check(lib.sfem_cg_iterate(
            op.handle, hp, sxh, float(lam), float(mu),
            int(num_interface_elements or 0), 1, _lib.ptr(x), _lib.ptr(r),
            _lib.ptr(p), _lib.ptr(ap), _lib.ptr(minv), _lib.ptr(owned),
            _lib.ptr(state), iters, stream), 'sfem_cg_iterate')
        continue
      for _ in range(iters):
        apply(p, pap)
        allreduce(pap)
        _lib._check(lib.sfem_cg_update(
            code, n, _lib.ptr(x), _lib.ptr(r), _lib.ptr(p), _lib.ptr(ap),
            _lib.ptr(minv), _lib.ptr(owned), _lib.ptr(state), stream),
                    'sfem_cg_update')
        allreduce(gnew)
        _lib._check(lib.sfem_cg_direction(
            code, n, _lib.ptr(r), _lib.ptr(p), _lib.ptr(minv),
            _lib.ptr(state), stream), 'sfem_cg_direction')
        _lib._check(lib.sfem_cg_advance(_lib.ptr(state), stream),
                    'sfem_cg_advance')
  residual = torch.tensor(info.residual, dtype=dtype, device=dev)
  return x, {'residual': residual, 'num_iterations': int(info.num_iterations)}
