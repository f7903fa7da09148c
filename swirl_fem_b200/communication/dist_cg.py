"""Element-partitioned (P)CG: one process per GPU, NCCL for the two scalars.

Same recurrence and stopping rule as `swirl_fem/linalg/cg.py:54-97`, executed
by the fused CUDA building blocks (`sfem_cg_init/_update/_direction/_advance`)
on a device-resident state.  Per iteration and rank:

  apply (local block, p.Ap partial in the kernel epilogue)
  -> halo exchange of Ap (pack, NCCL send/recv over NVLink, unpack-add)
  -> all-reduce of p.Ap          (1 double; element-wise partial sums need no
                                  ownership weights)
  -> update kernel (x, r, partial r.z over OWNED dofs)
  -> all-reduce of r.z           (1 double)
  -> direction kernel, scalar advance (device side)

There is no host synchronisation inside the loop: the convergence flag is read
every `check_every` iterations (it is identical on all ranks because it is
computed from all-reduced scalars).  The reference has no distributed CG; its
hook is `dot_fn` (`cg.py:26-31`), and the parity target is the unpartitioned
solve on the same global mesh.
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
      dot products per iteration are all-reduced over peer memory by one
      single-CTA kernel each instead of an NCCL call (sums in rank order,
      bitwise identical on all ranks).  None: NCCL.
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
    if world > 1:
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
    while True:
      _lib._check(lib.sfem_cg_read(_lib.ptr(state), ctypes.byref(info),
                                   ctypes.byref(done), stream), 'sfem_cg_read')
      if done.value:
        break
      for _ in range(check_every):
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
