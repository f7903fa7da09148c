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
import os

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
    self._lazy = None
    # SFEM_LAZY_ZERO=1: zero y's shared dofs lazily where a kernel instance
    # exists (see enable_lazy_zero); default: the eager fill before the apply
    if (os.environ.get('SFEM_LAZY_ZERO', '0') == '1' and mesh.ndim == 3
        and interp.collocated):
      self.enable_lazy_zero()

  def enable_lazy_zero(self, chunk_steps: int | None = None,
                       ahead: int | None = None, piece: int = 1024) -> bool:
    """Builds the tables of the lazy zero fill (`sfem_op_set_lazy_zero`): the
    apply's CTAs then claim their element steps from a counter, and a
    companion kernel zeroes y's shared dofs while the apply runs, each shortly
    before its first accumulation (see `lazy_zero_tables`).  `chunk_steps`:
    CTA steps per chunk; `ahead`: extra lead of the companion in steps.
    Returns False -- the eager fill stays -- when this operator has no lazy
    kernel instance or the mesh is too small."""
    lib = _lib.lib()
    se, grid, sup = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
    _lib._check(lib.sfem_op_lazy_zero_query(
        self.handle, ctypes.byref(se), ctypes.byref(grid), ctypes.byref(sup)),
                'sfem_op_lazy_zero_query')
    if not sup.value:
      return False
    if chunk_steps is None:
      chunk_steps = int(os.environ.get('SFEM_LAZY_CHUNK', 128))
    if ahead is None:
      ahead = int(os.environ.get('SFEM_LAZY_AHEAD', 0))
    if self.mesh.num_elements < 4 * grid.value * se.value:
      return False   # a couple of waves at least
    tables = lazy_zero_tables(
        self.mesh.elements, self.num_nodes,
        int(lib.sfem_op_num_zero(self.handle)), se.value, chunk_steps, piece)
    if tables is None:
      return False
    pieces, chunk_ptr = tables
    with torch.cuda.device(pieces.device):
      _lib._check(lib.sfem_op_set_lazy_zero(
          self.handle, _lib.ptr(pieces), int(pieces.shape[0]),
          _lib.ptr(chunk_ptr), int(chunk_ptr.numel() - 1), int(chunk_steps),
          int(ahead)), 'sfem_op_set_lazy_zero')
    self._lazy = (pieces, chunk_ptr)   # the C side retains these pointers
    return True

  def disable_lazy_zero(self):
    _lib._check(_lib.lib().sfem_op_set_lazy_zero(
        self.handle, None, 0, None, 0, 0, 0), 'sfem_op_set_lazy_zero')
    self._lazy = None

  def lazy_zero_timed_out(self) -> str:
    """'' or the record of the first device-side wait that hit its limit
    (every result since then is invalid).  Synchronises."""
    dev = self.geom.device
    lib = _lib.lib()
    with torch.cuda.device(dev):
      if lib.sfem_op_lazy_zero_timed_out(self.handle, _lib.stream_ptr(dev)):
        return lib.sfem_last_error().decode()
    return ''

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


def lazy_zero_tables(elements: torch.Tensor, num_nodes: int, num_zero: int,
                     step_elems: int, chunk_steps: int, piece: int = 1024):
  """Tables of the lazy zero fill (see `sfem_op_set_lazy_zero`).

  The lazy apply's CTAs claim consecutive element steps (`step_elems` elements
  each) from a counter, so the elements are cut into chunks of `chunk_steps`
  steps = `chunk_steps * step_elems` consecutive elements; every dof of the
  shared prefix [0, num_zero) belongs to the chunk that touches it FIRST (dofs
  no element touches: chunk 0).  Returns `(pieces int32 (P, 2) = {first dof,
  length | chunk << 12}, chunk_ptr int32 (num_chunks + 1,))` -- pieces sorted
  by chunk, never crossing a chunk boundary or a gap in the ids, at most
  `piece` dofs long -- or None when there are fewer than 4 chunks.  Index
  arithmetic only (torch, any device; set-up time)."""
  E, n = int(elements.shape[0]), int(elements.shape[1])
  if (step_elems <= 0 or chunk_steps <= 0 or num_zero <= 0
      or num_nodes >= 2 ** 31 or not 1 <= piece <= 4095):
    return None
  chunk_elems = chunk_steps * step_elems
  num_chunks = -(-E // chunk_elems)
  if num_chunks < 4 or num_chunks >= (1 << 19):
    return None
  dev = elements.device
  big = torch.iinfo(torch.int32).max
  nz = int(num_zero)
  first = torch.full((nz,), big, dtype=torch.int64, device=dev)
  # element blocks keep the temporaries small on 10^8-dof meshes
  blk = max(1, (1 << 26) // max(n, 1))
  for e0 in range(0, E, blk):
    ids = elements[e0:e0 + blk].reshape(-1).long()
    chunk = (torch.arange(e0, min(e0 + blk, E), device=dev)
             // chunk_elems).repeat_interleave(n)
    sel = (ids >= 0) & (ids < nz)
    first.scatter_reduce_(0, ids[sel], chunk[sel], 'amin', include_self=True)
    del ids, chunk, sel
  first = torch.where(first == big, torch.zeros_like(first), first)
  order = torch.argsort(first, stable=True)            # ids by chunk, then id
  st = first[order]
  # position in `order` of the first dof of chunk c, c = 0 .. num_chunks
  bounds = torch.searchsorted(
      st, torch.arange(0, num_chunks + 1, device=dev, dtype=st.dtype))
  cut = torch.zeros(nz + 1, dtype=torch.bool, device=dev)
  cut[0] = True
  cut[1:nz] = order[1:] != order[:-1] + 1               # gaps in the ids
  cut[bounds.clamp(max=nz)] = True                      # chunk boundaries
  seg_start = torch.nonzero(cut[:nz]).reshape(-1)
  seg_end = torch.cat([seg_start[1:], torch.tensor([nz], device=dev)])
  npieces = (seg_end - seg_start + piece - 1) // piece
  seg_of_piece = torch.repeat_interleave(
      torch.arange(seg_start.numel(), device=dev), npieces)
  first_piece = torch.cumsum(npieces, 0) - npieces
  k = torch.arange(seg_of_piece.numel(), device=dev) - first_piece[seg_of_piece]
  pos = seg_start[seg_of_piece] + k * piece
  plen = torch.minimum(torch.full_like(pos, piece),
                       seg_end[seg_of_piece] - pos)
  pchunk = st[pos]
  pieces = torch.stack([order[pos], plen + (pchunk << 12)],
                       dim=1).to(torch.int32).contiguous()
  # first piece of chunk c = number of pieces starting before its first dof
  chunk_ptr = torch.searchsorted(pos, bounds).to(torch.int32).contiguous()
  return pieces, chunk_ptr


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
