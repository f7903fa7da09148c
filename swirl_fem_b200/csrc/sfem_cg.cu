// Fused (P)CG -- swirl_fem/linalg/cg.py:30-97 restructured for HBM traffic.
//
// Reference body (cg.py:75-86): Ap = A(p); alpha = gamma / p.Ap; x += alpha p;
// r -= alpha Ap; z = M r; gamma' = r.z; beta = gamma'/gamma; p = z + beta p.
// Here, per iteration:
//   1. operator apply with p.Ap accumulated in its epilogue   (apply kernel)
//   2. update: x, r, and gamma' = r.(minv r) in one pass       (5 reads 2 writes)
//   3. direction: p = minv r + beta p                          (3 reads 1 write)
//   4. a one-thread kernel advances the device-side scalars and the
//      convergence flag (cond_fun of cg.py:68-73: gamma > atol2 && k < maxiter).
// All scalars stay on the device; the host only polls the flag every
// `check_every` iterations, and kernels launched past convergence are no-ops,
// so the iteration count equals the reference's.
//
// The same kernels are exported as building blocks (sfem_cg_init / _update /
// _direction / _advance) operating on a device-resident state, so that a
// multi-GPU host loop can put the NCCL all-reduces of the two scalars and the
// halo exchange between them without any host synchronisation; `owned` (one
// byte per dof) weights the vector dot products so that dofs shared between
// ranks are counted once.

#include "sfem_common.cuh"

namespace sfem {

int op_apply_internal(const sfem_op* op, double lambda, double mu,
                      const void* x, void* y, int ncomp, double* dot_xy,
                      cudaStream_t stream);

namespace {

constexpr int kThreads = 256;

// Device-resident CG state.  The first four doubles are the quantities a
// distributed host all-reduces: [0] p.Ap, [1] gamma_new, [2] gamma, [3] b.b.
struct CgState {
  double pAp;
  double gamma_new;
  double gamma;
  double bs;
  double atol2;
  double tol, atol;
  double k;        // iteration count (exact in a double)
  double maxiter;
  double done;     // 0 / 1
};
static_assert(sizeof(CgState) <= 256, "state must fit the reserved 256 bytes");

inline int blocks_for(int64_t n, int per_thread) {
  int64_t b = (n + (int64_t)kThreads * per_thread - 1) /
              ((int64_t)kThreads * per_thread);
  const int64_t cap = (int64_t)num_sms() * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
cg_init_kernel(int64_t n, const T* __restrict__ b, const T* __restrict__ Ax,
               const T* __restrict__ minv, const uint8_t* __restrict__ owned,
               T* __restrict__ r, T* __restrict__ p, CgState* __restrict__ st) {
  __shared__ double red[32];
  double g = 0.0, bs = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const T bi = b[i];
    const T ri = bi - Ax[i];
    const T zi = minv ? minv[i] * ri : ri;
    r[i] = ri;
    p[i] = zi;
    if (!owned || owned[i]) {
      g += (double)ri * (double)zi;
      bs += (double)bi * (double)bi;
    }
  }
  g = block_sum(g, red);
  bs = block_sum(bs, red);
  if (threadIdx.x == 0) {
    atomicAdd(&st->gamma, g);
    atomicAdd(&st->bs, bs);
  }
}

__global__ void cg_init_scalars(CgState* st) {
  const double t2 = st->tol * st->tol * st->bs;
  const double a2 = st->atol * st->atol;
  st->atol2 = t2 > a2 ? t2 : a2;
  st->k = 0.0;
  st->pAp = 0.0;
  st->gamma_new = 0.0;
  st->done = (st->gamma > st->atol2 && st->k < st->maxiter) ? 0.0 : 1.0;
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
cg_update_kernel(int64_t n, T* __restrict__ x, T* __restrict__ r,
                 const T* __restrict__ p, const T* __restrict__ Ap,
                 const T* __restrict__ minv, const uint8_t* __restrict__ owned,
                 CgState* __restrict__ st) {
  if (st->done != 0.0) return;
  __shared__ double red[32];
  const T alpha = (T)(st->gamma / st->pAp);
  double g = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    x[i] += alpha * p[i];
    const T ri = r[i] - alpha * Ap[i];
    r[i] = ri;
    const T zi = minv ? minv[i] * ri : ri;
    if (!owned || owned[i]) g += (double)ri * (double)zi;
  }
  g = block_sum(g, red);
  if (threadIdx.x == 0) atomicAdd(&st->gamma_new, g);
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
cg_direction_kernel(int64_t n, const T* __restrict__ r, T* __restrict__ p,
                    const T* __restrict__ minv,
                    const CgState* __restrict__ st) {
  if (st->done != 0.0) return;
  const T beta = (T)(st->gamma_new / st->gamma);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const T ri = r[i];
    const T zi = minv ? minv[i] * ri : ri;
    p[i] = zi + beta * p[i];
  }
}

__global__ void cg_step_scalars(CgState* st) {
  if (st->done != 0.0) return;
  st->gamma = st->gamma_new;
  st->gamma_new = 0.0;
  st->k += 1.0;
  st->done = (st->gamma > st->atol2 && st->k < st->maxiter) ? 0.0 : 1.0;
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
axpby_kernel(int64_t n, T a, const T* __restrict__ x, T b, T* __restrict__ y) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    y[i] = b == T(0) ? a * x[i] : a * x[i] + b * y[i];
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
dot_kernel(int64_t n, const T* __restrict__ x, const T* __restrict__ y,
           double* __restrict__ result) {
  __shared__ double red[32];
  double acc = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    acc += (double)x[i] * (double)y[i];
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(result, acc);
}

int init_state(CgState* st, double tol, double atol, double maxiter,
               cudaStream_t stream) {
  CgState h{};
  h.tol = tol;
  h.atol = atol;
  h.maxiter = maxiter;
  SFEM_CUDA_CHECK(cudaMemcpyAsync(st, &h, sizeof(h), cudaMemcpyHostToDevice,
                                  stream));
  return SFEM_OK;
}

template <typename T>
int cg_impl(const sfem_op* op, const void* b_, void* x_, int ncomp,
            const void* minv_, const sfem_cg_params* prm, void* workspace,
            sfem_cg_info* info, cudaStream_t stream) {
  const int64_t n = op->base.desc.num_nodes * (int64_t)ncomp;
  const T* b = (const T*)b_;
  T* x = (T*)x_;
  const T* minv = (const T*)minv_;
  char* ws = (char*)workspace;
  CgState* st = (CgState*)ws;
  T* r = (T*)(ws + 256);
  T* p = r + n;
  T* Ap = p + n;

  int rc = init_state(st, prm->tol, prm->atol,
                      (double)(prm->maxiter > 0 ? prm->maxiter : 10 * n),
                      stream);
  if (rc) return rc;
  const int nb = blocks_for(n, 4);
  // r0 = b - A x0; p0 = z0 = M r0; gamma0 = r0.z0 (cg.py:88-92)
  rc = op_apply_internal(op, prm->lambda, prm->mu, x, Ap, ncomp, nullptr,
                         stream);
  if (rc) return rc;
  if (n > 0) {
    cg_init_kernel<T><<<nb, kThreads, 0, stream>>>(n, b, Ap, minv, nullptr, r,
                                                   p, st);
    SFEM_LAUNCH_CHECK();
  }
  cg_init_scalars<<<1, 1, 0, stream>>>(st);
  SFEM_LAUNCH_CHECK();

  const int check_every = prm->check_every > 0 ? prm->check_every : 16;
  CgState hs{};
  for (;;) {
    SFEM_CUDA_CHECK(cudaMemcpyAsync(&hs, st, sizeof(hs), cudaMemcpyDeviceToHost,
                                    stream));
    SFEM_CUDA_CHECK(cudaStreamSynchronize(stream));
    if (hs.done != 0.0) break;
    for (int it = 0; it < check_every; ++it) {
      rc = op_apply_internal(op, prm->lambda, prm->mu, p, Ap, ncomp, &st->pAp,
                             stream);
      if (rc) return rc;
      cg_update_kernel<T><<<nb, kThreads, 0, stream>>>(n, x, r, p, Ap, minv,
                                                       nullptr, st);
      SFEM_LAUNCH_CHECK();
      cg_direction_kernel<T><<<nb, kThreads, 0, stream>>>(n, r, p, minv, st);
      SFEM_LAUNCH_CHECK();
      cg_step_scalars<<<1, 1, 0, stream>>>(st);
      SFEM_LAUNCH_CHECK();
    }
  }
  if (info) {
    info->residual = hs.gamma;
    info->num_iterations = (int64_t)hs.k;
  }
  return SFEM_OK;
}

}  // namespace
}  // namespace sfem

extern "C" {

int64_t sfem_cg_workspace_bytes(int dtype, int64_t size) {
  const int64_t esz = dtype == SFEM_F64 ? 8 : 4;
  return 256 + 3 * size * esz;
}

int sfem_cg(const sfem_op* op, const void* b, void* x, int32_t ncomp,
            const void* minv, const sfem_cg_params* params, void* workspace,
            sfem_cg_info* info, sfem_stream_t stream) {
  using namespace sfem;
  SFEM_REQUIRE(op && b && x && params && workspace, "null argument");
  SFEM_REQUIRE(ncomp >= 1, "bad ncomp");
  SFEM_REQUIRE(params->precond == 0 || minv != nullptr,
               "Jacobi preconditioner needs minv");
  const void* m = params->precond ? minv : nullptr;
  return op->base.desc.dtype == SFEM_F64
             ? cg_impl<double>(op, b, x, ncomp, m, params, workspace, info,
                               (cudaStream_t)stream)
             : cg_impl<float>(op, b, x, ncomp, m, params, workspace, info,
                              (cudaStream_t)stream);
}

// ---- building blocks (distributed host loops) --------------------------------

int64_t sfem_cg_state_bytes(void) { return 256; }

int sfem_cg_init(int dtype, int64_t n, const void* b, const void* Ax,
                 const void* minv, const uint8_t* owned, void* r, void* p,
                 void* state, double tol, double atol, int64_t maxiter,
                 sfem_stream_t stream_) {
  using namespace sfem;
  cudaStream_t stream = (cudaStream_t)stream_;
  SFEM_REQUIRE(b && Ax && r && p && state, "null argument");
  CgState* st = (CgState*)state;
  int rc = init_state(st, tol, atol,
                      (double)(maxiter > 0 ? maxiter : 10 * n), stream);
  if (rc) return rc;
  if (n == 0) return SFEM_OK;
  const int nb = blocks_for(n, 4);
  if (dtype == SFEM_F64)
    cg_init_kernel<double><<<nb, kThreads, 0, stream>>>(
        n, (const double*)b, (const double*)Ax, (const double*)minv, owned,
        (double*)r, (double*)p, st);
  else
    cg_init_kernel<float><<<nb, kThreads, 0, stream>>>(
        n, (const float*)b, (const float*)Ax, (const float*)minv, owned,
        (float*)r, (float*)p, st);
  SFEM_LAUNCH_CHECK();
  return SFEM_OK;
}

int sfem_cg_init_finish(void* state, sfem_stream_t stream) {
  using namespace sfem;
  SFEM_REQUIRE(state, "null argument");
  cg_init_scalars<<<1, 1, 0, (cudaStream_t)stream>>>((CgState*)state);
  SFEM_LAUNCH_CHECK();
  return SFEM_OK;
}

int sfem_cg_update(int dtype, int64_t n, void* x, void* r, const void* p,
                   const void* Ap, const void* minv, const uint8_t* owned,
                   void* state, sfem_stream_t stream_) {
  using namespace sfem;
  cudaStream_t stream = (cudaStream_t)stream_;
  SFEM_REQUIRE(x && r && p && Ap && state, "null argument");
  if (n == 0) return SFEM_OK;
  const int nb = blocks_for(n, 4);
  if (dtype == SFEM_F64)
    cg_update_kernel<double><<<nb, kThreads, 0, stream>>>(
        n, (double*)x, (double*)r, (const double*)p, (const double*)Ap,
        (const double*)minv, owned, (CgState*)state);
  else
    cg_update_kernel<float><<<nb, kThreads, 0, stream>>>(
        n, (float*)x, (float*)r, (const float*)p, (const float*)Ap,
        (const float*)minv, owned, (CgState*)state);
  SFEM_LAUNCH_CHECK();
  return SFEM_OK;
}

int sfem_cg_direction(int dtype, int64_t n, const void* r, void* p,
                      const void* minv, void* state, sfem_stream_t stream_) {
  using namespace sfem;
  cudaStream_t stream = (cudaStream_t)stream_;
  SFEM_REQUIRE(r && p && state, "null argument");
  if (n == 0) return SFEM_OK;
  const int nb = blocks_for(n, 4);
  if (dtype == SFEM_F64)
    cg_direction_kernel<double><<<nb, kThreads, 0, stream>>>(
        n, (const double*)r, (double*)p, (const double*)minv, (CgState*)state);
  else
    cg_direction_kernel<float><<<nb, kThreads, 0, stream>>>(
        n, (const float*)r, (float*)p, (const float*)minv, (CgState*)state);
  SFEM_LAUNCH_CHECK();
  return SFEM_OK;
}

int sfem_cg_advance(void* state, sfem_stream_t stream) {
  using namespace sfem;
  SFEM_REQUIRE(state, "null argument");
  cg_step_scalars<<<1, 1, 0, (cudaStream_t)stream>>>((CgState*)state);
  SFEM_LAUNCH_CHECK();
  return SFEM_OK;
}

int sfem_cg_read(const void* state, sfem_cg_info* info, int32_t* done,
                 sfem_stream_t stream_) {
  using namespace sfem;
  cudaStream_t stream = (cudaStream_t)stream_;
  SFEM_REQUIRE(state && info && done, "null argument");
  CgState hs{};
  SFEM_CUDA_CHECK(cudaMemcpyAsync(&hs, state, sizeof(hs),
                                  cudaMemcpyDeviceToHost, stream));
  SFEM_CUDA_CHECK(cudaStreamSynchronize(stream));
  info->residual = hs.gamma;
  info->num_iterations = (int64_t)hs.k;
  *done = hs.done != 0.0;
  return SFEM_OK;
}

int sfem_axpby(int dtype, int64_t n, double a, const void* x, double b, void* y,
               sfem_stream_t stream_) {
  using namespace sfem;
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n == 0) return SFEM_OK;
  if (dtype == SFEM_F64)
    axpby_kernel<double><<<blocks_for(n, 4), kThreads, 0, stream>>>(
        n, a, (const double*)x, b, (double*)y);
  else
    axpby_kernel<float><<<blocks_for(n, 4), kThreads, 0, stream>>>(
        n, (float)a, (const float*)x, (float)b, (float*)y);
  SFEM_LAUNCH_CHECK();
  return SFEM_OK;
}

int sfem_dot(int dtype, int64_t n, const void* x, const void* y, void* result,
             sfem_stream_t stream_) {
  using namespace sfem;
  cudaStream_t stream = (cudaStream_t)stream_;
  SFEM_CUDA_CHECK(cudaMemsetAsync(result, 0, sizeof(double), stream));
  if (n == 0) return SFEM_OK;
  if (dtype == SFEM_F64)
    dot_kernel<double><<<blocks_for(n, 4), kThreads, 0, stream>>>(
        n, (const double*)x, (const double*)y, (double*)result);
  else
    dot_kernel<float><<<blocks_for(n, 4), kThreads, 0, stream>>>(
        n, (const float*)x, (const float*)y, (double*)result);
  SFEM_LAUNCH_CHECK();
  return SFEM_OK;
}

}  // extern "C"
