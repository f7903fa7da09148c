"""Fused matrix-free operator  y = mask . Z^T (lam*Mass + mu*Stiffness) Z x.

This is the B200 replacement for the reference's operator lambdas
  `A(u) = with_bc(mesh.scatter(fespace.local_covector(a, (gather(u), v))))`
(`swirl_fem/examples/poisson.py:119-146`), `StokesSEM.A` / `H_`
(`swirl_fem/navier_stokes/navier_stokes.py:304-307, 431`): one CUDA kernel does
gather -> sum-factorised local operator -> scatter -> Dirichlet mask, and
(optionally) the `p . Ap` dot product of CG in its epilogue.

The handle owns two device buffers (torch tensors, caller-visible):
  `geom`  (E, g, Q^d): symmetric geometric factors W detJ J^-T J^-1 (+ W detJ),
  `conn`  (E, N^d) int32: connectivity with the "single occurrence" and
          "Dirichlet" flags folded into the two top bits.
"""

from __future__ import annotations

import ctypes

import numpy as np
import torch

from swirl_fem_b200 import _lib


class FusedOperator:
  """`sfem_op` handle: fused operator on a mesh + quadrature rule."""

  def __init__(self, mesh, quadrature, dirichlet_mask=None,
               with_mass: bool = True):
    """Builds the operator on `mesh` with the 1-D `quadrature` rule.

    Only the packed symmetric factors are computed (K11); no `invjacs` /
    `jacdets` arrays are materialised, so this is also the cheap way to set up
    the >=100 M-dof configuration.
    """
    from swirl_fem_b200.core.interpolation import BarycentricInterpolator  # pylint: disable=g-import-not-at-top
    self.mesh = mesh
    self.quadrature = quadrature
    interp = BarycentricInterpolator(
        ndim=mesh.ndim, gridpoints_1d=mesh.gridpoints_1d,
        evalpoints_1d=quadrature.nodes)
    b, bd = interp.matrices_1d()
    self.dtype = mesh.node_coords.dtype
    self.desc = _lib.Desc(
        dim=mesh.ndim, n1d=mesh.gridpoints_1d.num_points,
        q1d=quadrature.num_points, dtype=self.dtype,
        collocated=interp.collocated, elements=mesh.elements,
        node_coords=mesh.node_coords, interp_1d=b, interp_grad_1d=bd,
        quad_weights_1d=quadrature.weights)
    dev = mesh.device
    self.with_mass = bool(with_mass)
    if dirichlet_mask is not None:
      dm = dirichlet_mask
      if not isinstance(dm, torch.Tensor):
        dm = torch.as_tensor(np.asarray(dm))
      dm = (dm.to(dev) != 0).to(torch.uint8).contiguous()
      if tuple(dm.shape) != (mesh.num_nodes,):
        raise ValueError('dirichlet_mask must have shape (num_nodes,)')
      self.dirichlet = dm
    else:
      self.dirichlet = None
    lib = _lib.lib()
    gbytes = lib.sfem_op_geom_bytes(ctypes.byref(self.desc.c),
                                    int(self.with_mass))
    cbytes = lib.sfem_op_conn_bytes(ctypes.byref(self.desc.c))
    esz = torch.empty((), dtype=self.dtype).element_size()
    self.geom = torch.empty(max(gbytes // esz, 1), dtype=self.dtype, device=dev)
    self.conn = torch.empty(max(cbytes // 4, 1), dtype=torch.int32, device=dev)
    handle = ctypes.c_void_p()
    with torch.cuda.device(dev):
      _lib._check(lib.sfem_op_create(
          ctypes.byref(self.desc.c), _lib.ptr(self.dirichlet),
          int(self.with_mass), _lib.ptr(self.geom), _lib.ptr(self.conn),
          ctypes.byref(handle), _lib.stream_ptr(dev)), 'sfem_op_create')
    self.handle = handle
    self.num_nodes = mesh.num_nodes
    self.ndim = mesh.ndim

  def __del__(self):
    h = getattr(self, 'handle', None)
    if h and _lib._lib is not None:
      _lib._lib.sfem_op_destroy(h)
      self.handle = None

  def set_variant(self, variant: int):
    """0: auto (specialised collocated kernels), 1: generic kernel (tests)."""
    _lib._check(_lib.lib().sfem_op_set_variant(self.handle, int(variant)),
                'sfem_op_set_variant')

  def _ncomp(self, x: torch.Tensor) -> int:
    if x.dim() == 1:
      return 1
    if x.dim() == 2:
      return x.shape[1]
    raise ValueError(f'expected (G,) or (G, ncomp), got {tuple(x.shape)}')

  def apply(self, x: torch.Tensor, lam: float = 0.0, mu: float = 1.0,
            out: torch.Tensor | None = None, dot_out: torch.Tensor | None = None):
    """y = mask . scatter(local(gather(x))).  `dot_out`: 0-d float64 for x.y."""
    _lib.require_cuda(x)
    if x.shape[0] != self.num_nodes:
      raise ValueError(
          f'Expected leading dimension {self.num_nodes}, got {tuple(x.shape)}')
    x = x.to(self.dtype).contiguous()
    y = torch.empty_like(x) if out is None else out
    assert y.is_contiguous() and y.dtype == x.dtype and y.shape == x.shape
    with torch.cuda.device(x.device):
      _lib._check(_lib.lib().sfem_op_apply(
          self.handle, float(lam), float(mu), _lib.ptr(x), _lib.ptr(y),
          self._ncomp(x), _lib.ptr(dot_out), _lib.stream_ptr(x.device)),
                  'sfem_op_apply')
    return y

  def apply_range(self, x: torch.Tensor, out: torch.Tensor, elem_begin: int,
                  elem_end: int, first: bool, lam: float = 0.0,
                  mu: float = 1.0, dot_out: torch.Tensor | None = None):
    """Accumulates the contribution of elements `[elem_begin, elem_end)`."""
    _lib.require_cuda(x, out)
    assert x.is_contiguous() and out.is_contiguous() and x.dtype == self.dtype
    with torch.cuda.device(x.device):
      _lib._check(_lib.lib().sfem_op_apply_range(
          self.handle, float(lam), float(mu), _lib.ptr(x), _lib.ptr(out),
          self._ncomp(x), int(elem_begin), int(elem_end), int(bool(first)),
          _lib.ptr(dot_out), _lib.stream_ptr(x.device)), 'sfem_op_apply_range')
    return out

  def apply_partitioned(self, x: torch.Tensor, out: torch.Tensor, halo,
                        num_interface_elements: int, lam: float = 0.0,
                        mu: float = 1.0, dot_out: torch.Tensor | None = None,
                        overlap: bool = False, wait: bool = True):
    """Partitioned apply: local apply + shared-dof (halo) exchange.

    With a peer-memory halo (`HaloPlan.enable_p2p`) this is the fused path:
    `sfem_op_apply_halo` (one launch; the push to the peers happens inside the
    kernel as soon as the interface elements are done) followed by the
    wait + canonical sum (`wait=False` leaves that to the caller).

    `overlap=True`: elements `[0, num_interface_elements)` are the ones
    touching another rank's block (`communication.partition` stores them
    first); their result is complete on the shared dofs after a first launch,
    so those dofs are packed and sent on a side stream while the interior
    elements run in a second launch.  Measured on 8 B200 at 13.6 M dofs/rank
    the split costs more (two persistent launches + cross-stream events:
    0.43 ms/step) than the exchange it hides (0.33 ms/step unsplit), so the
    default is the single launch followed by the exchange; the split pays off
    only when the wire time is large compared with a launch.
    """
    ne = self.mesh.num_elements
    ni = int(num_interface_elements)
    if halo is None or not halo.peers:
      return self.apply(x, lam=lam, mu=mu, out=out, dot_out=dot_out)
    hp = halo.p2p_handle(x) if hasattr(halo, 'p2p_handle') else None
    if hp is not None and not overlap:
      # ONE launch: interface elements first, shared dofs pushed to the peers
      # over NVLink from inside the kernel while the interior elements are
      # computed; then the wait + canonical sum (sfem_halo.cu)
      _lib.require_cuda(x, out)
      assert x.is_contiguous() and out.is_contiguous()
      assert x.dtype == self.dtype and out.dtype == self.dtype
      with torch.cuda.device(x.device):
        _lib._check(_lib.lib().sfem_op_apply_halo(
            self.handle, hp, float(lam), float(mu), _lib.ptr(x),
            _lib.ptr(out), ni, _lib.ptr(dot_out), _lib.stream_ptr(x.device)),
                    'sfem_op_apply_halo')
      if wait:
        halo.p2p_wait_unpack(out)
      return out
    if not overlap:
      self.apply(x, lam=lam, mu=mu, out=out, dot_out=dot_out)
      halo.exchange_(out)
      return out
    main = torch.cuda.current_stream(x.device)
    side = halo.side_stream(x.device)
    self.apply_range(x, out, 0, ni, True, lam, mu, dot_out)
    ready = torch.cuda.Event()
    ready.record(main)
    side.wait_event(ready)
    with torch.cuda.stream(side):
      halo.start_exchange(out)          # pack + all_to_all on the side stream
      sent = torch.cuda.Event()
      sent.record(side)
    self.apply_range(x, out, ni, ne, False, lam, mu, dot_out)
    main.wait_event(sent)
    halo.finish_exchange(out)           # unpack-add on the main stream
    return out

  def apply_local(self, u_local: torch.Tensor, lam: float = 0.0,
                  mu: float = 1.0, ncomp: int = 1) -> torch.Tensor:
    """E-vector form: `(E, n[, ncomp]) -> (E, n[, ncomp])`."""
    _lib.require_cuda(u_local)
    u_local = u_local.to(self.dtype).contiguous()
    y = torch.empty_like(u_local)
    with torch.cuda.device(u_local.device):
      _lib._check(_lib.lib().sfem_op_apply_local(
          self.handle, float(lam), float(mu), _lib.ptr(u_local), _lib.ptr(y),
          int(ncomp), _lib.stream_ptr(u_local.device)), 'sfem_op_apply_local')
    return y

  def diag(self, lam: float = 0.0, mu: float = 1.0) -> torch.Tensor:
    """diag(mask . Z^T (lam M + mu K) Z): the Jacobi preconditioner's input."""
    dev = self.mesh.device
    d = torch.empty(self.num_nodes, dtype=self.dtype, device=dev)
    with torch.cuda.device(dev):
      _lib._check(_lib.lib().sfem_op_diag(
          self.handle, float(lam), float(mu), _lib.ptr(d),
          _lib.stream_ptr(dev)), 'sfem_op_diag')
    return d

  def jacobi_minv(self, lam: float = 0.0, mu: float = 1.0) -> torch.Tensor:
    """`M(r) = mask . r / diag(A)` as an inverse-diagonal vector (0 on mask)."""
    d = self.diag(lam, mu)
    return torch.where(d != 0, 1.0 / d, torch.zeros_like(d))

  def bind(self, lam: float = 0.0, mu: float = 1.0):
    """Returns the callable `A(u)` for these coefficients (for `linalg.cg`)."""
    return BoundOperator(self, lam, mu)


class HostPipeline:
  """Streams HOST vectors through the operator: `y_host = A(x_host)`.

  PCIe is full duplex, so the upload of vector i+1, the apply of vector i and
  the download of result i-1 run concurrently: two copy streams next to the
  compute stream, `depth` device buffers each for x and y, ordered by events
  only (no host synchronisation per step).  Host tensors must be pinned.
  """

  def __init__(self, op: 'FusedOperator', halo=None,
               num_interface_elements: int = 0, depth: int = 2):
    dev = op.mesh.device
    self.op, self.halo, self.ni = op, halo, int(num_interface_elements)
    self.depth, self.count = int(depth), 0
    self.s_in = torch.cuda.Stream(device=dev)
    self.s_out = torch.cuda.Stream(device=dev)
    mk = lambda: torch.empty(op.num_nodes, dtype=op.dtype, device=dev)  # noqa: E731
    self.xd = [mk() for _ in range(depth)]
    self.yd = [mk() for _ in range(depth)]
    ev = lambda: [torch.cuda.Event() for _ in range(depth)]  # noqa: E731
    self.ev_in, self.ev_apply, self.ev_out = ev(), ev(), ev()

  def submit(self, x_host: torch.Tensor, y_host: torch.Tensor,
             lam: float = 0.0, mu: float = 1.0):
    """Enqueues upload -> apply -> download of one vector; returns at once."""
    if not (x_host.is_pinned() and y_host.is_pinned()):
      raise ValueError('HostPipeline needs pinned host tensors')
    k = self.count % self.depth
    dev = self.op.mesh.device
    main = torch.cuda.current_stream(dev)
    if self.count >= self.depth:
      self.s_in.wait_event(self.ev_apply[k])   # x buffer k is free again
      main.wait_event(self.ev_out[k])          # y buffer k has been drained
    with torch.cuda.stream(self.s_in):
      self.xd[k].copy_(x_host, non_blocking=True)
      self.ev_in[k].record(self.s_in)
    main.wait_event(self.ev_in[k])
    self.op.apply_partitioned(self.xd[k], self.yd[k], self.halo, self.ni,
                              lam=lam, mu=mu)
    self.ev_apply[k].record(main)
    with torch.cuda.stream(self.s_out):
      self.s_out.wait_event(self.ev_apply[k])
      y_host.copy_(self.yd[k], non_blocking=True)
      self.ev_out[k].record(self.s_out)
    self.count += 1

  def drain(self):
    """Makes the CURRENT stream wait for every submitted download (so an event
    recorded on it afterwards closes the timed region); no host sync."""
    main = torch.cuda.current_stream(self.op.mesh.device)
    for k in range(min(self.count, self.depth)):
      main.wait_event(self.ev_out[k])

  def synchronize(self):
    self.drain()
    torch.cuda.current_stream(self.op.mesh.device).synchronize()


class BoundOperator:
  """`A(u)` with fixed (lam, mu); recognised by `linalg.cg.cg` (fused path)."""

  def __init__(self, op: FusedOperator, lam: float, mu: float):
    self.op, self.lam, self.mu = op, float(lam), float(mu)

  def __call__(self, u: torch.Tensor) -> torch.Tensor:
    return self.op.apply(u, lam=self.lam, mu=self.mu)


class JacobiPreconditioner:
  """`M(r) = minv * r`; recognised by `linalg.cg.cg` (fused path)."""

  def __init__(self, minv: torch.Tensor):
    self.minv = minv.contiguous()

  def __call__(self, r: torch.Tensor) -> torch.Tensor:
    m = self.minv
    return r * (m if r.dim() == 1 else m.reshape(r.shape[0], -1))
