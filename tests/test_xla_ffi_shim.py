"""CPU check of the `jax.ffi` binding (swirl_fem_b200/csrc/xla_ffi_shim.cc).

jax / jaxlib cannot be installed in this image, so the shim is compiled
against an inert mock of `xla/ffi/api/ffi.h` (tests/mock_xla/): g++ type-checks
every handler against its `Ffi::Bind()` operand list and every `sfem_*` call
against include/swirl_b200.h (arity and argument types), and the resulting
object is linked against libswirl_b200.so so that every C symbol the shim uses
resolves."""

import os
import re
import shutil
import subprocess

import pytest

from swirl_fem_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, 'swirl_fem_b200', 'csrc', 'xla_ffi_shim.cc')
MOCK = os.path.join(ROOT, 'tests', 'mock_xla')
CUDA_INC = '/usr/local/cuda/include'

HANDLERS = ['sfem_xla_gather', 'sfem_xla_scatter_add', 'sfem_xla_op_apply',
            'sfem_xla_op_apply_local', 'sfem_xla_op_apply_halo',
            'sfem_xla_space_eval_transpose', 'sfem_xla_exchange',
            'sfem_xla_cg', 'sfem_xla_stokes_div', 'sfem_xla_stokes_grad_t']


def _need_toolchain():
  if shutil.which('g++') is None or not os.path.isdir(CUDA_INC):
    pytest.skip('needs g++ and the CUDA runtime headers')


def test_shim_type_checks_against_mock_header():
  _need_toolchain()
  proc = subprocess.run(
      ['g++', '-std=c++17', '-fsyntax-only', '-Wall', '-Werror', '-I', MOCK,
       '-I', CUDA_INC, SHIM], capture_output=True, text=True)
  assert proc.returncode == 0, proc.stderr[-4000:]


def test_mock_rejects_a_mismatched_handler(tmp_path):
  """The mock is not vacuous: a handler whose parameter list disagrees with
  its binding must fail to compile."""
  _need_toolchain()
  bad = tmp_path / 'bad.cc'
  bad.write_text('''
#include <cuda_runtime_api.h>
#include "xla/ffi/api/ffi.h"
namespace ffi = xla::ffi;
static ffi::Error Impl(cudaStream_t, ffi::AnyBuffer, double) {
  return ffi::Error::Success();
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(bad, Impl,
    ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Arg<ffi::AnyBuffer>().Attr<int64_t>("n").Attr<double>("x"));
''')
  proc = subprocess.run(
      ['g++', '-std=c++17', '-fsyntax-only', '-I', MOCK, '-I', CUDA_INC,
       str(bad)], capture_output=True, text=True)
  assert proc.returncode != 0
  assert 'does not match' in proc.stderr


def test_shim_links_against_the_library(tmp_path):
  _need_toolchain()
  if not os.path.exists(_lib.LIB_PATH):
    pytest.skip('library not built yet (run __graft_entry__.build())')
  out = tmp_path / 'libswirl_b200_xla.so'
  proc = subprocess.run(
      ['g++', '-std=c++17', '-shared', '-fPIC', '-I', MOCK, '-I', CUDA_INC,
       SHIM, '-L', os.path.dirname(_lib.LIB_PATH), '-lswirl_b200',
       '-L', '/usr/local/cuda/lib64', '-lcudart', '-Wl,--no-undefined',
       '-Wl,--allow-shlib-undefined', '-o', str(out)],
      capture_output=True, text=True)
  assert proc.returncode == 0, proc.stderr[-4000:]
  syms = subprocess.run(['nm', '-D', '--defined-only', str(out)],
                        capture_output=True, text=True).stdout
  for name in HANDLERS:
    assert re.search(rf'\bT {name}\b', syms), f'{name} not exported'


def test_every_c_symbol_the_shim_calls_is_declared():
  text = open(SHIM).read()
  text = re.sub(r'//.*', '', text)
  used = set(re.findall(r'\b(sfem_(?!xla_)[a-z0-9_]+)\s*\(', text))
  header = open(os.path.join(ROOT, 'include', 'swirl_b200.h')).read()
  header = re.sub(r'/\*.*?\*/', '', header, flags=re.S)
  declared = set(re.findall(r'\b(sfem_[a-z0-9_]+)\s*\(', header))
  assert used and used <= declared, used - declared
  # every handler named in INTEGRATION.md exists in the shim, and vice versa
  defined = set(re.findall(r'XLA_FFI_DEFINE_HANDLER_SYMBOL\((\w+)', text))
  assert defined == set(HANDLERS)
