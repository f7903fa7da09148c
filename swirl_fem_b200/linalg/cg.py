"""Conjugate gradients with the reference's semantics and signature.

`cg(A, b, x0=None, *, tol, atol, maxiter, M, dot_fn)` follows
`swirl_fem/linalg/cg.py:30-97`: x0 = 0, maxiter = 10*size, the stopping test
is on gamma = r . M r > max(tol^2 b.b, atol^2) (:65-73), and the return value
is `(x, {'residual': gamma, 'num_iterations': k})`.

Two execution paths, both CUDA:
  * fused: `A` is a `BoundOperator` (core/operator.py) and `M` is None or a
    `JacobiPreconditioner` -> one call of the C-ABI `sfem_cg`: operator apply
    with p.Ap in its epilogue, one update kernel, one direction kernel, all
    scalars and the convergence flag on the device.
  * device-state: `A` and `M` are arbitrary callables on ONE CUDA tensor
    (e.g. `H_` with `M = velocity.exchange`, `E = D Q D^T` with the null-space
    projector, navier_stokes.py:436-452) and `dot_fn` is the default -> the
    same recurrence on the device-resident state of `sfem_cg_*`: the dot
    products land in the state, `alpha` / `beta` are formed inside the vector
    kernels, and the host only reads the convergence flag every `check_every`
    iterations (no `float()` per dot product, no clones).
  * generic: pytrees or a user `dot_fn` -> host loop over the same recurrence
    using the fused `axpby` / `dot` kernels.
"""

from __future__ import annotations

import ctypes

import torch

from swirl_fem_b200 import _lib
from swirl_fem_b200.core.operator import BoundOperator
from swirl_fem_b200.core.operator import JacobiPreconditioner


# -- minimal pytree helpers (dict / list / tuple / tensor leaves) -------------


def tree_map(fn, *trees):
  t0 = trees[0]
  if isinstance(t0, dict):
    return {k: tree_map(fn, *[t[k] for t in trees]) for k in t0}
  if isinstance(t0, (list, tuple)):
    return type(t0)(tree_map(fn, *[t[i] for t in trees])
                    for i in range(len(t0)))
  return fn(*trees)


def tree_leaves(tree):
  if isinstance(tree, dict):
    return [l for k in tree for l in tree_leaves(tree[k])]
  if isinstance(tree, (list, tuple)):
    return [l for t in tree for l in tree_leaves(t)]
  return [tree]


def _default_dot(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
  return _lib.dot(a, b)


def _vdot(a, b, dot_fn):
  return sum(tree_leaves(tree_map(dot_fn, a, b)))


def _fused_cg(A: BoundOperator, b, x0, tol, atol, maxiter, M, check_every):
  op = A.op
  _lib.require_cuda(b)
  b = b.to(op.dtype).contiguous()
  ncomp = 1 if b.dim() == 1 else b.shape[1]
  x = torch.zeros_like(b) if x0 is None else x0.to(op.dtype).clone().contiguous()
  minv = None
  if M is not None:
    minv = M.minv.to(op.dtype)
    if minv.numel() != b.numel():
      minv = minv.reshape(-1, 1).expand(b.shape[0], ncomp)
    minv = minv.contiguous()
  lib = _lib.lib()
  wbytes = lib.sfem_cg_workspace_bytes(_lib.dtype_code(op.dtype), b.numel())
  work = torch.empty(wbytes, dtype=torch.uint8, device=b.device)
  params = _lib.CgParams(
      tol=float(tol), atol=float(atol),
      maxiter=-1 if maxiter is None else int(maxiter),
      precond=0 if minv is None else 1, check_every=int(check_every),
      lam=A.lam, mu=A.mu)
  info = _lib.CgInfo()
  with torch.cuda.device(b.device):
    _lib._check(lib.sfem_cg(op.handle, _lib.ptr(b), _lib.ptr(x), ncomp,
                            _lib.ptr(minv), ctypes.byref(params),
                            _lib.ptr(work), ctypes.byref(info),
                            _lib.stream_ptr(b.device)), 'sfem_cg')
  residual = torch.tensor(info.residual, dtype=op.dtype, device=b.device)
  return x, {'residual': residual, 'num_iterations': int(info.num_iterations)}


# Diagnostics of the most recent `_device_state_cg` call (tools/bench_ns.py):
# whether the iteration body was captured into a CUDA graph.
LAST_DEVICE_STATE_RUN = {}


def _device_state_cg(A, b, x0, tol, atol, maxiter, M, check_every,
                     graph=None):
  """cg.py:54-97 for callables `A`, `M` on one CUDA tensor, scalars on the
  device.  Per iteration: A, dot(p, Ap) -> state, `sfem_cg_update` (x, r with
  alpha = gamma / p.Ap formed in the kernel), M, dot(r, z) -> state,
  `sfem_cg_direction` (p = z + beta p), `sfem_cg_advance`; once the flag is
  set the vector kernels are no-ops, so the count equals the reference's.

  `graph` (default: env SFEM_CG_GRAPH, on): the iteration body -- whatever
  kernels `A` and `M` launch plus the five above -- may be captured into a
  CUDA graph and replayed; the body is static because every scalar lives in
  the device state.  It is captured only where that pays: after 4 *
  `check_every` eager iterations, and only if the eager batches were
  host-bound (launch-bound Stokes solves on small meshes: 222 -> 90 us per
  pressure iteration at 147 k dofs; at 2.4 M dofs the eager loop is already
  GPU-bound and stays eager).  Callables that cannot be captured (host
  synchronisation inside) fall back to eager launches."""
  import os  # pylint: disable=g-import-not-at-top
  import time  # pylint: disable=g-import-not-at-top
  if graph is None:
    graph = os.environ.get('SFEM_CG_GRAPH', '1') != '0'
  lib = _lib.lib()
  dev, dtype = b.device, b.dtype
  code = _lib.dtype_code(dtype)
  b = b.contiguous()
  n = b.numel()
  if maxiter is None:
    maxiter = 10 * n
  x = torch.zeros_like(b) if x0 is None else x0.to(dtype).clone().contiguous()
  state = torch.zeros(int(lib.sfem_cg_state_bytes()) // 8, dtype=torch.float64,
                      device=dev)
  stream = _lib.stream_ptr(dev)
  identity = M is None

  def as_vec(t):
    t = t.to(dtype)
    return t if t.is_contiguous() else t.contiguous()

  with torch.cuda.device(dev):
    r = b.clone()
    _lib.axpby(-1.0, as_vec(A(x)), 1.0, r)                  # r = b - A x0
    z = r if identity else as_vec(M(r))
    p = z.clone()
    # n = 0: only (re)initialises the state (tol, atol, maxiter)
    _lib._check(lib.sfem_cg_init(code, 0, _lib.ptr(b), _lib.ptr(r), None, None,
                                 _lib.ptr(r), _lib.ptr(p), _lib.ptr(state),
                                 float(tol), float(atol), int(maxiter), stream),
                'sfem_cg_init')
    _lib._check(lib.sfem_dot(code, n, _lib.ptr(r), _lib.ptr(z),
                             _lib.ptr(state[2:3]), stream), 'sfem_dot')
    _lib._check(lib.sfem_dot(code, n, _lib.ptr(b), _lib.ptr(b),
                             _lib.ptr(state[3:4]), stream), 'sfem_dot')
    _lib._check(lib.sfem_cg_init_finish(_lib.ptr(state), stream),
                'sfem_cg_init_finish')
    info = _lib.CgInfo()
    done = ctypes.c_int32(0)
    pap, gnew = state[0:1], state[1:2]

    def iteration():
      """Enqueues one iteration on the CURRENT stream (fixed addresses only:
      x, r, p and the state are updated in place)."""
      st = _lib.stream_ptr(dev)
      ap = as_vec(A(p))
      _lib._check(lib.sfem_dot(code, n, _lib.ptr(p), _lib.ptr(ap),
                               _lib.ptr(pap), st), 'sfem_dot')
      # x += alpha p, r -= alpha Ap (its own r.r lands in gamma_new and is
      # overwritten by r.z below unless M is the identity)
      _lib._check(lib.sfem_cg_update(code, n, _lib.ptr(x), _lib.ptr(r),
                                     _lib.ptr(p), _lib.ptr(ap), None, None,
                                     _lib.ptr(state), st), 'sfem_cg_update')
      if identity:
        zz = r
      else:
        zz = as_vec(M(r))
        _lib._check(lib.sfem_dot(code, n, _lib.ptr(r), _lib.ptr(zz),
                                 _lib.ptr(gnew), st), 'sfem_dot')
      _lib._check(lib.sfem_cg_direction(code, n, _lib.ptr(zz), _lib.ptr(p),
                                        None, _lib.ptr(state), st),
                  'sfem_cg_direction')
      _lib._check(lib.sfem_cg_advance(_lib.ptr(state), st), 'sfem_cg_advance')

    def capture():
      """The iteration body as a CUDA graph, or False if it cannot be captured.
      (`CUDAGraph.capture_begin/_end` directly: the `torch.cuda.graph` context
      also runs `gc.collect()` and `torch.cuda.empty_cache()`, which costs
      ~1 s per solve once the caching allocator holds GBs -- measured on the
      Stokes step at 3.2 M velocity dofs.)"""
      cur = torch.cuda.current_stream(dev)
      side = torch.cuda.Stream(device=dev)
      side.wait_stream(cur)
      g = torch.cuda.CUDAGraph()
      ok = True
      with torch.cuda.stream(side):
        g.capture_begin()
        try:
          iteration()
        except Exception:  # pylint: disable=broad-except
          ok = False
        try:
          g.capture_end()
        except Exception:  # pylint: disable=broad-except
          ok = False
      cur.wait_stream(side)
      if not ok:
        torch.cuda.synchronize(dev)
        return False
      return g

    # Eager first: a graph pays off only when the HOST is the limit (many small
    # kernels) and the solve is long.  Every eager batch is timed -- enqueue
    # time against the time until the state read returns -- and the body is
    # captured after `capture_after` eager iterations if the host needed most
    # of the time of the last three batches to enqueue them.
    capture_after = 4 * max(1, int(check_every))
    captured = None
    eager_done = 0
    host_bound = False
    streak = 0       # consecutive host-bound eager batches
    pending = None   # (start, enqueue time) of the eager batch in flight
    while True:
      _lib._check(lib.sfem_cg_read(_lib.ptr(state), ctypes.byref(info),
                                   ctypes.byref(done), stream), 'sfem_cg_read')
      if pending is not None:
        total = time.perf_counter() - pending[0]
        host_bound = pending[1] > 0.6 * total
        streak = streak + 1 if host_bound else 0
        pending = None
      if done.value:
        break
      iters = int(min(check_every, max(1, maxiter - info.num_iterations)))
      if (graph and captured is None and streak >= 3
          and eager_done >= capture_after):
        # everything lazy (module loading, occupancy queries, cached index
        # maxima) happened in the eager iterations
        try:
          captured = capture()
        except Exception:  # pylint: disable=broad-except
          captured = False   # not capturable: stay eager
          torch.cuda.synchronize(dev)
      if captured:
        for _ in range(iters):
          captured.replay()
      else:
        t0 = time.perf_counter()
        for _ in range(iters):
          iteration()
        eager_done += iters
        pending = (t0, time.perf_counter() - t0)
  LAST_DEVICE_STATE_RUN.update(
      graph_requested=bool(graph), graph_captured=bool(captured),
      host_bound=bool(host_bound), eager_iterations=eager_done,
      iterations=int(info.num_iterations))
  residual = torch.tensor(info.residual, dtype=dtype, device=dev)
  return x, {'residual': residual, 'num_iterations': int(info.num_iterations)}


def cg(A, b, x0=None, *, tol=1e-5, atol=0.0, maxiter=None, M=None,
       dot_fn=None, check_every=16, graph=None):
  """Conjugate gradient solver; see the module docstring."""
  if (isinstance(A, BoundOperator) and isinstance(b, torch.Tensor)
      and (M is None or isinstance(M, JacobiPreconditioner))
      and dot_fn is None):
    return _fused_cg(A, b, x0, tol, atol, maxiter, M, check_every)
  if (isinstance(b, torch.Tensor) and dot_fn is None and b.is_cuda
      and b.dtype in (torch.float32, torch.float64) and b.numel() > 0
      and (x0 is None or isinstance(x0, torch.Tensor))):
    return _device_state_cg(A, b, x0, tol, atol, maxiter, M, check_every,
                            graph=graph)

  if dot_fn is None:
    dot_fn = _default_dot
  for leaf in tree_leaves(b):
    _lib.require_cuda(leaf)
  if x0 is None:
    x0 = tree_map(torch.zeros_like, b)
  if maxiter is None:
    maxiter = 10 * sum(l.numel() for l in tree_leaves(b))
  if M is None:
    M = lambda v: v

  bs = _vdot(b, b, dot_fn)
  atol2 = max(float(tol) ** 2 * float(bs), float(atol) ** 2)

  def axpy(a, xs, ys):
    """ys + a * xs, out of place, through the axpby kernel."""
    def one(xl, yl):
      out = yl.contiguous().clone()
      _lib.axpby(a, xl.contiguous(), 1.0, out)
      return out
    return tree_map(one, xs, ys)

  r = axpy(-1.0, A(x0), b)
  z = M(r)
  p = z
  gamma = float(_vdot(r, z, dot_fn))
  x = x0
  k = 0
  while gamma > atol2 and k < maxiter:
    Ap = A(p)
    alpha = gamma / float(_vdot(p, Ap, dot_fn))
    x = axpy(alpha, p, x)
    r = axpy(-alpha, Ap, r)
    z = M(r)
    gamma_new = float(_vdot(r, z, dot_fn))
    beta = gamma_new / gamma
    p = axpy(beta, p, z)
    gamma = gamma_new
    k += 1
  leaf0 = tree_leaves(b)[0]
  residual = torch.tensor(gamma, dtype=leaf0.dtype, device=leaf0.device)
  return x, {'residual': residual, 'num_iterations': k}
