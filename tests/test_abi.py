"""CPU checks of the C-ABI boundary: the library loads and exports every symbol
that include/swirl_b200.h declares (no compute calls without a GPU)."""

import ctypes
import os
import re

import pytest

from swirl_fem_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'swirl_b200.h')


def _declared_symbols():
  text = open(HEADER).read()
  text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
  return sorted(set(re.findall(r'\b(sfem_[a-z0-9_]+)\s*\(', text)))


def test_header_symbols_are_bound_and_exported():
  if not os.path.exists(_lib.LIB_PATH):
    pytest.skip('library not built yet (run __graft_entry__.build())')
  declared = _declared_symbols()
  assert len(declared) >= 25
  handle = ctypes.CDLL(_lib.LIB_PATH)
  for name in declared:
    assert hasattr(handle, name), f'{name} declared in the header, not exported'
    assert name in _lib.SIGNATURES, f'{name} has no ctypes signature'
  assert sorted(_lib.SIGNATURES) == declared


def test_library_loads_and_reports_version():
  if not os.path.exists(_lib.LIB_PATH):
    pytest.skip('library not built yet')
  lib = _lib.lib()
  assert lib.sfem_version() >= 100
  assert lib.sfem_launch_count() >= 0
  assert isinstance(lib.sfem_last_error(), bytes)


def test_struct_layouts_match_header():
  # sfem_space_desc: 6 x int32, 2 x int64, 5 pointers; sfem_cg_params:
  # 2 double, int64, 2 int32, 2 double; sfem_cg_info: double + int64
  assert ctypes.sizeof(_lib.SpaceDesc) == 6 * 4 + 2 * 8 + 5 * 8
  assert ctypes.sizeof(_lib.CgParams) == 2 * 8 + 8 + 2 * 4 + 2 * 8
  assert ctypes.sizeof(_lib.CgInfo) == 16


def test_no_cpu_path():
  import torch
  from swirl_fem_b200.core import gather_scatter as gs
  with pytest.raises(_lib.SwirlB200Error, match='no CPU path'):
    gs.gather(torch.zeros(4, dtype=torch.float64),
              torch.zeros(2, dtype=torch.int32))
  if not torch.cuda.is_available():
    from swirl_fem_b200.core.mesh import Mesh
    import numpy as np
    with pytest.raises(_lib.SwirlB200Error, match='CUDA device'):
      Mesh.create(np.zeros((2, 1)), np.array([[0, 1]]))
