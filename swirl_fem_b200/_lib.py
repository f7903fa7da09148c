"""ctypes binding of libswirl_b200.so (the C ABI in include/swirl_b200.h).

PyTorch is used here only as plumbing: device memory (`torch.empty(...,
device='cuda')`), the current CUDA stream and dtype bookkeeping.  Every
compute call goes through the C ABI into hand-written sm_100a kernels.  There
is NO CPU fallback: if the shared library is missing, or a tensor is not on a
CUDA device, the call raises.
"""

from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# SFEM_LIB: an alternative build of the SAME library (e.g. the tuning-variant
# build `make OUTDIR=../lib_exp EXTRA=-DSFEM_EXPERIMENTS`); never a fallback.
LIB_PATH = os.environ.get('SFEM_LIB') or os.path.join(_HERE, 'lib',
                                                      'libswirl_b200.so')
CSRC_DIR = os.path.join(_HERE, 'csrc')

SFEM_F32, SFEM_F64 = 0, 1
SENTINEL = -1
MAX_1D = 18

_c_i32 = ctypes.c_int32
_c_i64 = ctypes.c_int64
_c_f64 = ctypes.c_double
_c_ptr = ctypes.c_void_p


class SpaceDesc(ctypes.Structure):
  """`sfem_space_desc` (include/swirl_b200.h)."""
  _fields_ = [
      ('dim', _c_i32), ('n1d', _c_i32), ('q1d', _c_i32), ('dtype', _c_i32),
      ('collocated', _c_i32), ('reserved', _c_i32),
      ('num_elements', _c_i64), ('num_nodes', _c_i64),
      ('elements', _c_ptr), ('node_coords', _c_ptr),
      ('interp_1d', _c_ptr), ('interp_grad_1d', _c_ptr),
      ('quad_weights_1d', _c_ptr),
  ]


class CgParams(ctypes.Structure):
  """`sfem_cg_params`."""
  _fields_ = [
      ('tol', _c_f64), ('atol', _c_f64), ('maxiter', _c_i64),
      ('precond', _c_i32), ('check_every', _c_i32),
      ('lam', _c_f64), ('mu', _c_f64),
  ]


class HaloDesc(ctypes.Structure):
  """`sfem_halo_desc`."""
  _fields_ = [
      ('dtype', _c_i32), ('rank', _c_i32), ('world', _c_i32),
      ('num_peers', _c_i32),
      ('num_send', _c_i64), ('send_idx', _c_ptr), ('send_dst', _c_ptr),
      ('parity_stride_bytes', ctypes.c_uint64), ('peer_flag_addr', _c_ptr),
      ('peer_ranks', _c_ptr), ('flags', _c_ptr), ('recv', _c_ptr),
      ('num_dofs', _c_i64), ('dofs', _c_ptr), ('row_ptr', _c_ptr),
      ('src', _c_ptr),
  ]


class CgInfo(ctypes.Structure):
  """`sfem_cg_info`."""
  _fields_ = [('residual', _c_f64), ('num_iterations', _c_i64)]


# name -> (restype, argtypes): every symbol include/swirl_b200.h declares.
SIGNATURES = {
    'sfem_last_error': (ctypes.c_char_p, []),
    'sfem_version': (ctypes.c_int, []),
    'sfem_launch_count': (_c_i64, []),
    'sfem_gather': (ctypes.c_int, [ctypes.c_int, _c_ptr, _c_ptr, _c_i64, _c_f64,
                                   _c_i32, _c_i32, _c_ptr, _c_ptr]),
    'sfem_scatter_add': (ctypes.c_int, [ctypes.c_int, _c_ptr, _c_ptr, _c_i64,
                                        _c_i64, _c_i32, _c_i32, _c_ptr,
                                        _c_ptr]),
    'sfem_scatter_plan_create': (ctypes.c_int, [_c_ptr, _c_i64, _c_i64,
                                                ctypes.POINTER(_c_ptr),
                                                _c_ptr]),
    'sfem_scatter_plan_apply': (ctypes.c_int, [_c_ptr, ctypes.c_int, _c_ptr,
                                               _c_i32, _c_i32, _c_ptr, _c_ptr]),
    'sfem_scatter_plan_destroy': (None, [_c_ptr]),
    'sfem_exchange': (ctypes.c_int, [ctypes.c_int, _c_ptr, _c_ptr, _c_ptr,
                                     _c_i64, _c_i64, _c_i32, _c_i32, _c_ptr,
                                     _c_ptr]),
    'sfem_halo_pack': (ctypes.c_int, [ctypes.c_int, _c_ptr, _c_ptr, _c_i64,
                                      _c_ptr, _c_ptr]),
    'sfem_halo_unpack_add': (ctypes.c_int, [ctypes.c_int, _c_ptr, _c_ptr,
                                            _c_i64, _c_ptr, _c_ptr]),
    'sfem_halo_unpack_canonical': (ctypes.c_int, [ctypes.c_int, _c_ptr, _c_ptr,
                                                  _c_ptr, _c_ptr, _c_i64,
                                                  _c_ptr, _c_ptr]),
    'sfem_ipc_alloc': (ctypes.c_int, [_c_i64, ctypes.POINTER(_c_ptr), _c_ptr]),
    'sfem_ipc_open': (ctypes.c_int, [_c_ptr, ctypes.POINTER(_c_ptr)]),
    'sfem_ipc_close': (ctypes.c_int, [_c_ptr]),
    'sfem_ipc_free': (ctypes.c_int, [_c_ptr]),
    'sfem_halo_create': (ctypes.c_int, [ctypes.POINTER(HaloDesc),
                                        ctypes.POINTER(_c_ptr)]),
    'sfem_halo_destroy': (None, [_c_ptr]),
    'sfem_halo_set_option': (ctypes.c_int, [_c_ptr, _c_i32, _c_i64]),
    'sfem_halo_push': (ctypes.c_int, [_c_ptr, _c_ptr, _c_ptr]),
    'sfem_halo_wait_unpack': (ctypes.c_int, [_c_ptr, _c_ptr, _c_ptr]),
    'sfem_halo_timed_out': (ctypes.c_int, [_c_ptr, _c_ptr]),
    'sfem_halo_debug_times': (ctypes.c_int, [_c_ptr, _c_ptr, _c_ptr]),
    'sfem_scalar_region_bytes': (_c_i64, [_c_i32]),
    'sfem_scalar_exchange_create': (ctypes.c_int, [_c_i32, _c_i32, _c_ptr,
                                                   _c_ptr,
                                                   ctypes.POINTER(_c_ptr)]),
    'sfem_scalar_exchange_destroy': (None, [_c_ptr]),
    'sfem_scalar_allreduce': (ctypes.c_int, [_c_ptr, _c_ptr, _c_i32, _c_ptr]),
    'sfem_scalar_exchange_timed_out': (ctypes.c_int, [_c_ptr, _c_ptr]),
    'sfem_op_apply_halo': (ctypes.c_int, [_c_ptr, _c_ptr, _c_f64, _c_f64,
                                          _c_ptr, _c_ptr, _c_i64, _c_ptr,
                                          _c_ptr]),
    'sfem_space_create': (ctypes.c_int, [ctypes.POINTER(SpaceDesc), _c_ptr,
                                         _c_ptr, _c_ptr,
                                         ctypes.POINTER(_c_ptr), _c_ptr]),
    'sfem_space_destroy': (None, [_c_ptr]),
    'sfem_space_eval': (ctypes.c_int, [_c_ptr, _c_ptr, _c_i32, _c_i32, _c_ptr,
                                       _c_ptr]),
    'sfem_space_eval_transpose': (ctypes.c_int, [_c_ptr, _c_ptr, _c_ptr,
                                                 _c_i32, _c_ptr, _c_ptr]),
    'sfem_pointwise': (ctypes.c_int, [ctypes.c_int, _c_i32, _c_i32, _c_ptr,
                                      _c_ptr, _c_i64, _c_ptr, _c_ptr]),
    'sfem_stokes_div': (ctypes.c_int, [_c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_ptr]),
    'sfem_stokes_grad_t': (ctypes.c_int, [_c_ptr, _c_ptr, _c_ptr, _c_ptr,
                                          _c_ptr, _c_ptr]),
    'sfem_space_integrate': (ctypes.c_int, [_c_ptr, _c_ptr, _c_ptr, _c_ptr]),
    'sfem_op_geom_bytes': (_c_i64, [ctypes.POINTER(SpaceDesc), _c_i32]),
    'sfem_op_conn_bytes': (_c_i64, [ctypes.POINTER(SpaceDesc)]),
    'sfem_op_create': (ctypes.c_int, [ctypes.POINTER(SpaceDesc), _c_ptr,
                                      _c_i32, _c_ptr, _c_ptr,
                                      ctypes.POINTER(_c_ptr), _c_ptr]),
    'sfem_op_destroy': (None, [_c_ptr]),
    'sfem_op_apply': (ctypes.c_int, [_c_ptr, _c_f64, _c_f64, _c_ptr, _c_ptr,
                                     _c_i32, _c_ptr, _c_ptr]),
    'sfem_op_apply_range': (ctypes.c_int, [_c_ptr, _c_f64, _c_f64, _c_ptr,
                                           _c_ptr, _c_i32, _c_i64, _c_i64,
                                           _c_i32, _c_ptr, _c_ptr]),
    'sfem_op_apply_local': (ctypes.c_int, [_c_ptr, _c_f64, _c_f64, _c_ptr,
                                           _c_ptr, _c_i32, _c_ptr]),
    'sfem_op_diag': (ctypes.c_int, [_c_ptr, _c_f64, _c_f64, _c_ptr, _c_ptr]),
    'sfem_op_set_variant': (ctypes.c_int, [_c_ptr, _c_i32]),
    'sfem_op_num_zero': (ctypes.c_int64, [_c_ptr]),
    'sfem_op_lazy_zero_query': (ctypes.c_int, [
        _c_ptr, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32),
        ctypes.POINTER(ctypes.c_int32)]),
    'sfem_op_set_lazy_zero': (ctypes.c_int, [_c_ptr, _c_ptr, _c_i64, _c_ptr,
                                             _c_i32, _c_i32, _c_i32]),
    'sfem_op_lazy_zero_timed_out': (ctypes.c_int, [_c_ptr, _c_ptr]),
    'sfem_cg_workspace_bytes': (_c_i64, [ctypes.c_int, _c_i64]),
    'sfem_cg': (ctypes.c_int, [_c_ptr, _c_ptr, _c_ptr, _c_i32, _c_ptr,
                               ctypes.POINTER(CgParams), _c_ptr,
                               ctypes.POINTER(CgInfo), _c_ptr]),
    'sfem_cg_state_bytes': (_c_i64, []),
    'sfem_cg_init': (ctypes.c_int, [ctypes.c_int, _c_i64, _c_ptr, _c_ptr,
                                    _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_ptr,
                                    _c_f64, _c_f64, _c_i64, _c_ptr]),
    'sfem_cg_init_finish': (ctypes.c_int, [_c_ptr, _c_ptr]),
    'sfem_cg_update': (ctypes.c_int, [ctypes.c_int, _c_i64, _c_ptr, _c_ptr,
                                      _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_ptr,
                                      _c_ptr]),
    'sfem_cg_direction': (ctypes.c_int, [ctypes.c_int, _c_i64, _c_ptr, _c_ptr,
                                         _c_ptr, _c_ptr, _c_ptr]),
    'sfem_cg_advance': (ctypes.c_int, [_c_ptr, _c_ptr]),
    'sfem_cg_iterate': (ctypes.c_int, [_c_ptr, _c_ptr, _c_ptr, _c_f64, _c_f64,
                                       _c_i64, _c_i32, _c_ptr, _c_ptr, _c_ptr,
                                       _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_i32,
                                       _c_ptr]),
    'sfem_cg_read': (ctypes.c_int, [_c_ptr, ctypes.POINTER(CgInfo),
                                    ctypes.POINTER(_c_i32), _c_ptr]),
    'sfem_axpby': (ctypes.c_int, [ctypes.c_int, _c_i64, _c_f64, _c_ptr, _c_f64,
                                  _c_ptr, _c_ptr]),
    'sfem_dot': (ctypes.c_int, [ctypes.c_int, _c_i64, _c_ptr, _c_ptr, _c_ptr,
                                _c_ptr]),
}

_lib = None


class SwirlB200Error(RuntimeError):
  pass


def build(verbose: bool = False) -> str:
  """Compiles the CUDA sources in-tree for sm_100a (nvcc, `make -j`)."""
  jobs = str(max(1, min(8, os.cpu_count() or 1)))
  proc = subprocess.run(['make', '-j', jobs, '-C', CSRC_DIR],
                        capture_output=True, text=True)
  if verbose or proc.returncode != 0:
    print(proc.stdout)
    print(proc.stderr)
  if proc.returncode != 0:
    raise SwirlB200Error('building libswirl_b200.so failed')
  return LIB_PATH


def lib() -> ctypes.CDLL:
  """Loads the shared library (never falls back to anything else)."""
  global _lib
  if _lib is None:
    if not os.path.exists(LIB_PATH):
      raise SwirlB200Error(
          f'{LIB_PATH} not found: build it with `python -c "import '
          '__graft_entry__ as g; g.build()"` (needs nvcc).  There is no CPU '
          'fallback.')
    handle = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
      fn = getattr(handle, name)  # AttributeError if a symbol is missing
      fn.restype = restype
      fn.argtypes = argtypes
    _lib = handle
  return _lib


def _check(rc: int, what: str):
  if rc != 0:
    msg = lib().sfem_last_error().decode('utf-8', 'replace')
    if rc == -2:
      raise NotImplementedError(f'{what}: {msg}')
    if rc == -1:
      raise ValueError(f'{what}: {msg}')
    raise SwirlB200Error(f'{what} failed ({rc}): {msg}')


def dtype_code(dtype: torch.dtype) -> int:
  if dtype == torch.float64:
    return SFEM_F64
  if dtype == torch.float32:
    return SFEM_F32
  raise TypeError(f'swirl_fem_b200 computes in float32 or float64, got {dtype}')


def require_cuda(*tensors):
  for t in tensors:
    if t is not None and not t.is_cuda:
      raise SwirlB200Error(
          'swirl_fem_b200 has no CPU path: tensors must live on a CUDA device '
          f'(got device={t.device})')


def stream_ptr(device=None) -> int:
  return torch.cuda.current_stream(device).cuda_stream


def ptr(t: torch.Tensor | None) -> int | None:
  return None if t is None else t.data_ptr()


def launch_count() -> int:
  return int(lib().sfem_launch_count())


# ----------------------------------------------------------------------------
# gather / scatter / exchange
# ----------------------------------------------------------------------------


def _as_index(indices: torch.Tensor) -> torch.Tensor:
  if indices.dtype != torch.int32:
    indices = indices.to(torch.int32)
  return indices.contiguous()


def gather(u: torch.Tensor, indices: torch.Tensor, fill_value=SENTINEL):
  """`u[indices]` (SENTINEL -> fill).  `u` of shape `(G, c)` (AoS, component
  last) is gathered to `indices.shape + (c,)` in ONE launch."""
  require_cuda(u, indices)
  u = u.contiguous()
  idx = _as_index(indices)
  ncomp = 1 if u.dim() == 1 else u.shape[-1]
  shape = tuple(idx.shape) + (() if u.dim() == 1 else (ncomp,))
  out = torch.empty(shape, dtype=u.dtype, device=u.device)
  with torch.cuda.device(u.device):
    _check(lib().sfem_gather(dtype_code(u.dtype), ptr(u), ptr(idx),
                             idx.numel(), float(fill_value), ncomp,
                             0 if u.dim() == 1 else -1, ptr(out),
                             stream_ptr(u.device)), 'sfem_gather')
  return out


def scatter(u_local: torch.Tensor, indices: torch.Tensor, num_nodes: int):
  """Zero-init scatter-add; `u_local` of shape `indices.shape + (c,)` gives
  `(num_nodes, c)` in ONE launch."""
  require_cuda(u_local, indices)
  u_local = u_local.contiguous()
  idx = _as_index(indices)
  vector = u_local.dim() == idx.dim() + 1
  ncomp = u_local.shape[-1] if vector else 1
  out = torch.empty((num_nodes, ncomp) if vector else (num_nodes,),
                    dtype=u_local.dtype, device=u_local.device)
  with torch.cuda.device(u_local.device):
    _check(lib().sfem_scatter_add(dtype_code(u_local.dtype), ptr(u_local),
                                  ptr(idx), idx.numel(), num_nodes, ncomp,
                                  -1 if vector else 0, ptr(out),
                                  stream_ptr(u_local.device)),
           'sfem_scatter_add')
  return out


class ScatterPlan:
  """Deterministic (atomic-free) scatter: sorted map + warp-segmented sums."""

  def __init__(self, indices: torch.Tensor, num_nodes: int):
    require_cuda(indices)
    self.indices = _as_index(indices)
    self.num_nodes = int(num_nodes)
    handle = _c_ptr()
    with torch.cuda.device(self.indices.device):
      _check(lib().sfem_scatter_plan_create(
          ptr(self.indices), self.indices.numel(), self.num_nodes,
          ctypes.byref(handle), stream_ptr(self.indices.device)),
             'sfem_scatter_plan_create')
    self._handle = handle

  def __call__(self, u_local: torch.Tensor) -> torch.Tensor:
    require_cuda(u_local)
    if u_local.numel() != self.indices.numel():
      raise ValueError('u_local does not match the plan')
    u_local = u_local.contiguous()
    out = torch.empty(self.num_nodes, dtype=u_local.dtype,
                      device=u_local.device)
    with torch.cuda.device(u_local.device):
      _check(lib().sfem_scatter_plan_apply(
          self._handle, dtype_code(u_local.dtype), ptr(u_local), 1, 0,
          ptr(out), stream_ptr(u_local.device)), 'sfem_scatter_plan_apply')
    return out

  def __del__(self):
    h = getattr(self, '_handle', None)
    if h and _lib is not None:
      _lib.sfem_scatter_plan_destroy(h)
      self._handle = None


_NUM_UNIQUE = {}  # (data_ptr, numel) of a unique-index tensor -> max + 1


def exchange(u: torch.Tensor, gather_indices: torch.Tensor,
             unique_indices: torch.Tensor | None, inplace: bool = False):
  """QQ^T over the (periodic) shared dofs; `(G, c)` fields in ONE call.
  `inplace`: `u` (contiguous) is overwritten instead of copied."""
  require_cuda(u, gather_indices, unique_indices)
  if inplace and not u.is_contiguous():
    raise ValueError('in-place exchange needs a contiguous tensor')
  out = u if inplace else u.contiguous().clone()
  gi = _as_index(gather_indices)
  ui = None if unique_indices is None else _as_index(unique_indices)
  count = gi.numel()
  if count == 0:
    return out
  if ui is None:
    num_unique = count
  else:
    # the only host read of the path: done once per index tensor, so that the
    # call never synchronises afterwards (CUDA-graph capturable)
    key = (ui.data_ptr(), ui.numel())
    num_unique = _NUM_UNIQUE.get(key)
    if num_unique is None:
      num_unique = int(ui.max().item()) + 1
      _NUM_UNIQUE[key] = num_unique
  ncomp = 1 if u.dim() == 1 else u.shape[-1]
  scratch = torch.empty(num_unique * ncomp, dtype=u.dtype, device=u.device)
  with torch.cuda.device(u.device):
    _check(lib().sfem_exchange(dtype_code(u.dtype), ptr(out), ptr(gi), ptr(ui),
                               count, num_unique, ncomp,
                               0 if u.dim() == 1 else -1, ptr(scratch),
                               stream_ptr(u.device)), 'sfem_exchange')
  return out


def pointwise(kind: int, dim: int, a, g, out):
  """`sfem_pointwise`: trace (0), scalar times identity (1), convection (2)."""
  require_cuda(a, g, out)
  npts = out.numel() // {0: 1, 1: dim * dim, 2: dim}[kind]
  with torch.cuda.device(out.device):
    _check(lib().sfem_pointwise(dtype_code(out.dtype), kind, dim, ptr(a),
                                ptr(g), npts, ptr(out),
                                stream_ptr(out.device)), 'sfem_pointwise')
  return out


def stokes_div(vspace_handle, pspace_handle, u, num_pressure_nodes: int):
  """`sfem_stokes_div`: fused D (velocity (G_v, d) -> pressure covector (G_p,)).
  Raises NotImplementedError when the element is too large for the fused
  kernel (the caller keeps the composed path)."""
  require_cuda(u)
  u = u.contiguous()
  out = torch.empty(num_pressure_nodes, dtype=u.dtype, device=u.device)
  with torch.cuda.device(u.device):
    _check(lib().sfem_stokes_div(vspace_handle, pspace_handle, ptr(u), ptr(out),
                                 stream_ptr(u.device)), 'sfem_stokes_div')
  return out


def stokes_grad_t(vspace_handle, pspace_handle, p, mask, num_velocity_nodes,
                  ndim):
  """`sfem_stokes_grad_t`: fused D^T (pressure (G_p,) -> velocity (G_v, d)),
  rows scaled by `mask` (G_v,)."""
  require_cuda(p, mask)
  p = p.contiguous()
  out = torch.empty((num_velocity_nodes, ndim), dtype=p.dtype, device=p.device)
  with torch.cuda.device(p.device):
    _check(lib().sfem_stokes_grad_t(vspace_handle, pspace_handle, ptr(p),
                                    ptr(mask), ptr(out), stream_ptr(p.device)),
           'sfem_stokes_grad_t')
  return out


def halo_pack(u, idx, buf):
  require_cuda(u, idx, buf)
  with torch.cuda.device(u.device):
    _check(lib().sfem_halo_pack(dtype_code(u.dtype), ptr(u), ptr(idx),
                                idx.numel(), ptr(buf), stream_ptr(u.device)),
           'sfem_halo_pack')


def halo_unpack_canonical(u, dofs, row_ptr, src, recv):
  require_cuda(u, dofs, row_ptr, src, recv)
  with torch.cuda.device(u.device):
    _check(lib().sfem_halo_unpack_canonical(
        dtype_code(u.dtype), ptr(u), ptr(dofs), ptr(row_ptr), ptr(src),
        dofs.numel(), ptr(recv), stream_ptr(u.device)),
           'sfem_halo_unpack_canonical')


def halo_unpack_add(u, idx, buf):
  require_cuda(u, idx, buf)
  with torch.cuda.device(u.device):
    _check(lib().sfem_halo_unpack_add(dtype_code(u.dtype), ptr(u), ptr(idx),
                                      idx.numel(), ptr(buf),
                                      stream_ptr(u.device)),
           'sfem_halo_unpack_add')


# ----------------------------------------------------------------------------
# vector kernels
# ----------------------------------------------------------------------------


def axpby(a: float, x: torch.Tensor, b: float, y: torch.Tensor):
  """y <- a*x + b*y (in place on y)."""
  require_cuda(x, y)
  assert x.dtype == y.dtype and x.numel() == y.numel()
  assert x.is_contiguous() and y.is_contiguous()
  with torch.cuda.device(x.device):
    _check(lib().sfem_axpby(dtype_code(x.dtype), x.numel(), float(a), ptr(x),
                            float(b), ptr(y), stream_ptr(x.device)),
           'sfem_axpby')
  return y


def dot(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
  """Returns x . y as a 0-d float64 device tensor."""
  require_cuda(x, y)
  assert x.dtype == y.dtype and x.numel() == y.numel()
  x = x.contiguous()
  y = y.contiguous()
  out = torch.empty((), dtype=torch.float64, device=x.device)
  with torch.cuda.device(x.device):
    _check(lib().sfem_dot(dtype_code(x.dtype), x.numel(), ptr(x), ptr(y),
                          ptr(out), stream_ptr(x.device)), 'sfem_dot')
  return out


# ----------------------------------------------------------------------------
# descriptor helper
# ----------------------------------------------------------------------------


class Desc:
  """Keeps the host arrays referenced by a `sfem_space_desc` alive."""

  def __init__(self, *, dim, n1d, q1d, dtype, collocated, elements,
               node_coords, interp_1d, interp_grad_1d, quad_weights_1d):
    require_cuda(elements, node_coords)
    if n1d > MAX_1D or q1d > MAX_1D:
      raise NotImplementedError(
          f'at most {MAX_1D} nodes / quadrature points per axis')
    self.elements = _as_index(elements)
    self.node_coords = node_coords.to(dtype).contiguous()
    self.b = np.ascontiguousarray(interp_1d, dtype=np.float64)
    self.bd = np.ascontiguousarray(interp_grad_1d, dtype=np.float64)
    self.w = np.ascontiguousarray(quad_weights_1d, dtype=np.float64)
    assert self.b.shape == (q1d, n1d) and self.bd.shape == (q1d, n1d)
    assert self.w.shape == (q1d,)
    self.c = SpaceDesc(
        dim=dim, n1d=n1d, q1d=q1d, dtype=dtype_code(dtype),
        collocated=int(bool(collocated)), reserved=0,
        num_elements=self.elements.shape[0],
        num_nodes=self.node_coords.shape[0],
        elements=ptr(self.elements), node_coords=ptr(self.node_coords),
        interp_1d=self.b.ctypes.data, interp_grad_1d=self.bd.ctypes.data,
        quad_weights_1d=self.w.ctypes.data)
    self.dtype = dtype
    self.device = self.node_coords.device
