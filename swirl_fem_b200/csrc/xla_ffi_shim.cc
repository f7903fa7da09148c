// XLA typed-FFI shim over the C ABI of libswirl_b200 (include/swirl_b200.h), so
// that the entry points can be bound as `jax.ffi` custom calls
// (jax.ffi.register_ffi_target(name, jax.ffi.pycapsule(fn), platform="CUDA")).
//
// HEADER-GATED: jax / jaxlib (and therefore xla/ffi/api/ffi.h) are NOT
// installed in the build image and there is no network, so this file compiles
// to an empty translation unit here and is NOT part of libswirl_b200.so.  The
// CPU suite compiles it against an inert mock of that header
// (tests/mock_xla/, tests/test_xla_ffi_shim.py): every handler's signature is
// checked against its Ffi::Bind() operand list and every C entry point it
// calls against include/swirl_b200.h, and the object is linked against
// libswirl_b200.so.  It is the binding a maintainer of the reference builds
// next to jaxlib:
//   g++ -shared -fPIC -I$(python -c "import jaxlib; print(jaxlib.__path__[0])")/include
//       xla_ffi_shim.cc -L../lib -lswirl_b200 -o libswirl_b200_xla.so   (one command line)
// The tested binding of the same C symbols is ctypes + torch (swirl_fem_b200/_lib.py).
#if __has_include("xla/ffi/api/ffi.h")

#include <cuda_runtime_api.h>

#include "../../include/swirl_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

static int DtypeOf(ffi::DataType t) {
  return t == ffi::DataType::F64 ? SFEM_F64 : SFEM_F32;
}

static ffi::Error Status(int rc) {
  if (rc == SFEM_OK) return ffi::Error::Success();
  return ffi::Error(rc == SFEM_ERR_UNSUPPORTED ? ffi::ErrorCode::kUnimplemented
                                               : ffi::ErrorCode::kInternal,
                    sfem_last_error());
}

// gather_scatter.gather (swirl_fem/core/gather_scatter.py:121-127)
static ffi::Error GatherImpl(cudaStream_t stream, ffi::AnyBuffer u,
                             ffi::Buffer<ffi::DataType::S32> indices,
                             ffi::Result<ffi::AnyBuffer> out, double fill) {
  return Status(sfem_gather(DtypeOf(u.element_type()), u.untyped_data(),
                            indices.typed_data(), indices.element_count(), fill,
                            1, 0, out->untyped_data(), (sfem_stream_t)stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(sfem_xla_gather, GatherImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Arg<ffi::Buffer<ffi::DataType::S32>>()
                                  .Ret<ffi::AnyBuffer>()
                                  .Attr<double>("fill_value"));

// gather_scatter.scatter (gather_scatter.py:130-133)
static ffi::Error ScatterImpl(cudaStream_t stream, ffi::AnyBuffer u_local,
                              ffi::Buffer<ffi::DataType::S32> indices,
                              ffi::Result<ffi::AnyBuffer> out) {
  return Status(sfem_scatter_add(
      DtypeOf(u_local.element_type()), u_local.untyped_data(),
      indices.typed_data(), indices.element_count(), out->element_count(), 1, 0,
      out->untyped_data(), (sfem_stream_t)stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(sfem_xla_scatter_add, ScatterImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Arg<ffi::Buffer<ffi::DataType::S32>>()
                                  .Ret<ffi::AnyBuffer>());

// A(u) = mask . scatter(local_covector(a, (gather(u), v)))
// (swirl_fem/examples/poisson.py:141-146); `handle` = address of the sfem_op
// created once outside jit.
static ffi::Error OpApplyImpl(cudaStream_t stream, ffi::AnyBuffer x,
                              ffi::Result<ffi::AnyBuffer> y, int64_t handle,
                              double lam, double mu) {
  const auto dims = x.dimensions();
  const int ncomp = dims.size() == 2 ? (int)dims[1] : 1;
  return Status(sfem_op_apply(reinterpret_cast<const sfem_op*>(handle), lam, mu,
                              x.untyped_data(), y->untyped_data(), ncomp,
                              nullptr, (sfem_stream_t)stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(sfem_xla_op_apply, OpApplyImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>()
                                  .Attr<int64_t>("handle")
                                  .Attr<double>("lam")
                                  .Attr<double>("mu"));

// FiniteElementSpace.local_covector (swirl_fem/core/fespace.py:405-471)
static ffi::Error OpApplyLocalImpl(cudaStream_t stream, ffi::AnyBuffer u_local,
                                   ffi::Result<ffi::AnyBuffer> y_local,
                                   int64_t handle, double lam, double mu) {
  const auto dims = u_local.dimensions();
  const int ncomp = dims.size() == 3 ? (int)dims[2] : 1;
  return Status(sfem_op_apply_local(
      reinterpret_cast<const sfem_op*>(handle), lam, mu, u_local.untyped_data(),
      y_local->untyped_data(), ncomp, (sfem_stream_t)stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(sfem_xla_op_apply_local, OpApplyLocalImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>()
                                  .Attr<int64_t>("handle")
                                  .Attr<double>("lam")
                                  .Attr<double>("mu"));

// Partitioned A(u): local operator + the psum of gather_scatter.exchange
// (swirl_fem/core/gather_scatter.py:246-248) as ONE launch with the in-kernel
// peer-memory push, then the wait + canonical sum.  `op` / `halo` = addresses
// of the sfem_op / sfem_halo created once outside jit.
static ffi::Error OpApplyHaloImpl(cudaStream_t stream, ffi::AnyBuffer x,
                                  ffi::Result<ffi::AnyBuffer> y, int64_t op,
                                  int64_t halo, int64_t num_interface_elements,
                                  double lam, double mu) {
  int rc = sfem_op_apply_halo(reinterpret_cast<const sfem_op*>(op),
                              reinterpret_cast<sfem_halo*>(halo), lam, mu,
                              x.untyped_data(), y->untyped_data(),
                              num_interface_elements, nullptr,
                              (sfem_stream_t)stream);
  if (rc == SFEM_OK)
    rc = sfem_halo_wait_unpack(reinterpret_cast<sfem_halo*>(halo),
                               y->untyped_data(), (sfem_stream_t)stream);
  return Status(rc);
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(sfem_xla_op_apply_halo, OpApplyHaloImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>()
                                  .Attr<int64_t>("op")
                                  .Attr<int64_t>("halo")
                                  .Attr<int64_t>("num_interface_elements")
                                  .Attr<double>("lam")
                                  .Attr<double>("mu"));

// Transposed evaluation = jax.linear_transpose of the quadrature integral
// (swirl_fem/core/fespace.py:458-471) for a form linear in its placeholder:
// vals (E, q, c) / grads (E, q, d, c) -> covector (E, n, c).
static ffi::Error EvalTransposeImpl(cudaStream_t stream, ffi::AnyBuffer vals,
                                    ffi::AnyBuffer grads,
                                    ffi::Result<ffi::AnyBuffer> out,
                                    int64_t space, int64_t ncomp) {
  // a zero-sized operand stands for "no such coefficient field" (NULL in C)
  return Status(sfem_space_eval_transpose(
      reinterpret_cast<const sfem_space*>(space),
      vals.element_count() ? vals.untyped_data() : nullptr,
      grads.element_count() ? grads.untyped_data() : nullptr, (int32_t)ncomp,
      out->untyped_data(), (sfem_stream_t)stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(sfem_xla_space_eval_transpose, EvalTransposeImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>()
                                  .Attr<int64_t>("space")
                                  .Attr<int64_t>("ncomp"));

// gather_scatter.exchange, unpartitioned (gather_scatter.py:221-261): in place
// on a copy of u; `scratch` is an XLA-allocated (num_unique,) work buffer.
static ffi::Error ExchangeImpl(cudaStream_t stream, ffi::AnyBuffer u,
                               ffi::Buffer<ffi::DataType::S32> gather_idx,
                               ffi::Buffer<ffi::DataType::S32> unique_idx,
                               ffi::Result<ffi::AnyBuffer> out,
                               ffi::Result<ffi::AnyBuffer> scratch) {
  const size_t esz = u.element_type() == ffi::DataType::F64 ? 8 : 4;
  if (cudaMemcpyAsync(out->untyped_data(), u.untyped_data(),
                      u.element_count() * esz, cudaMemcpyDeviceToDevice,
                      stream) != cudaSuccess)
    return ffi::Error(ffi::ErrorCode::kInternal, "cudaMemcpyAsync failed");
  return Status(sfem_exchange(
      DtypeOf(u.element_type()), out->untyped_data(), gather_idx.typed_data(),
      unique_idx.element_count() ? unique_idx.typed_data() : nullptr,
      (int64_t)gather_idx.element_count(), (int64_t)scratch->element_count(), 1,
      0, scratch->untyped_data(), (sfem_stream_t)stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(sfem_xla_exchange, ExchangeImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Arg<ffi::Buffer<ffi::DataType::S32>>()
                                  .Arg<ffi::Buffer<ffi::DataType::S32>>()
                                  .Ret<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>());

// linalg.cg.cg (swirl_fem/linalg/cg.py:30-97) with A = the fused operator and
// M = Jacobi or identity: x holds x0 on entry (an aliased input/output pair in
// jax.ffi.ffi_call(..., input_output_aliases={1: 0})); `workspace` is an
// XLA-allocated byte buffer of sfem_cg_workspace_bytes(); `info` receives
// (residual, iterations) as two doubles.
static ffi::Error CgImpl(cudaStream_t stream, ffi::AnyBuffer b,
                         ffi::AnyBuffer x0, ffi::AnyBuffer minv,
                         ffi::Result<ffi::AnyBuffer> x,
                         ffi::Result<ffi::AnyBuffer> workspace,
                         ffi::Result<ffi::Buffer<ffi::DataType::F64>> info,
                         int64_t handle, double tol, double atol,
                         int64_t maxiter, double lam, double mu) {
  const auto dims = b.dimensions();
  const int ncomp = dims.size() == 2 ? (int)dims[1] : 1;
  const size_t esz = b.element_type() == ffi::DataType::F64 ? 8 : 4;
  if (x->untyped_data() != x0.untyped_data() &&
      cudaMemcpyAsync(x->untyped_data(), x0.untyped_data(),
                      b.element_count() * esz, cudaMemcpyDeviceToDevice,
                      stream) != cudaSuccess)
    return ffi::Error(ffi::ErrorCode::kInternal, "cudaMemcpyAsync failed");
  sfem_cg_params prm;
  prm.tol = tol;
  prm.atol = atol;
  prm.maxiter = maxiter;
  prm.precond = minv.element_count() ? 1 : 0;
  prm.check_every = 16;
  prm.lambda = lam;
  prm.mu = mu;
  sfem_cg_info out_info;
  const int rc = sfem_cg(
      reinterpret_cast<const sfem_op*>(handle), b.untyped_data(),
      x->untyped_data(), ncomp,
      minv.element_count() ? minv.untyped_data() : nullptr, &prm,
      workspace->untyped_data(), &out_info, (sfem_stream_t)stream);
  if (rc != SFEM_OK) return Status(rc);
  const double host[2] = {out_info.residual, (double)out_info.num_iterations};
  if (cudaMemcpyAsync(info->typed_data(), host, sizeof(host),
                      cudaMemcpyHostToDevice, stream) != cudaSuccess ||
      cudaStreamSynchronize(stream) != cudaSuccess)
    return ffi::Error(ffi::ErrorCode::kInternal, "cudaMemcpyAsync failed");
  return ffi::Error::Success();
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(sfem_xla_cg, CgImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>()
                                  .Ret<ffi::Buffer<ffi::DataType::F64>>()
                                  .Attr<int64_t>("handle")
                                  .Attr<double>("tol")
                                  .Attr<double>("atol")
                                  .Attr<int64_t>("maxiter")
                                  .Attr<double>("lam")
                                  .Attr<double>("mu"));

// StokesSEM.D / Dt (swirl_fem/navier_stokes/navier_stokes.py:313-338) as ONE
// launch each; `vspace` / `pspace` = addresses of the sfem_space handles of the
// velocity and pressure spaces, created once outside jit.
static ffi::Error StokesDivImpl(cudaStream_t stream, ffi::AnyBuffer u,
                                ffi::Result<ffi::AnyBuffer> out, int64_t vspace,
                                int64_t pspace) {
  return Status(sfem_stokes_div(reinterpret_cast<const sfem_space*>(vspace),
                                reinterpret_cast<const sfem_space*>(pspace),
                                u.untyped_data(), out->untyped_data(),
                                (sfem_stream_t)stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(sfem_xla_stokes_div, StokesDivImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>()
                                  .Attr<int64_t>("vspace")
                                  .Attr<int64_t>("pspace"));

// `mask`: the velocity interior mask (G_v,), or a zero-sized operand for none.
static ffi::Error StokesGradTImpl(cudaStream_t stream, ffi::AnyBuffer p,
                                  ffi::AnyBuffer mask,
                                  ffi::Result<ffi::AnyBuffer> out,
                                  int64_t vspace, int64_t pspace) {
  return Status(sfem_stokes_grad_t(
      reinterpret_cast<const sfem_space*>(vspace),
      reinterpret_cast<const sfem_space*>(pspace), p.untyped_data(),
      mask.element_count() ? mask.untyped_data() : nullptr,
      out->untyped_data(), (sfem_stream_t)stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(sfem_xla_stokes_grad_t, StokesGradTImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>()
                                  .Attr<int64_t>("vspace")
                                  .Attr<int64_t>("pspace"));

#endif  // __has_include("xla/ffi/api/ffi.h")
