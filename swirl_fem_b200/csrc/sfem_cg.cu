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
                      cudaStream_t stream, bool prezeroed = false,
                      bool dot_prezeroed = false);
bool lazy_zero_applicable(const sfem_op* op, int ncomp, cudaStream_t stream);
int op_apply_halo_internal(const sfem_op* op, sfem_halo* halo, double lambda,
                           double mu, const void* x, void* y,
                           int64_t num_interface_elements, double* dot_xy,
                           bool prezeroed, cudaStream_t stream);

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
  double done;     // 0 running, 1 converged / maxiter, 2 peer wait timed out
  // fused step kernel (cg_step_kernel)
  double alpha, beta;
  unsigned long long seq;     // steps completed
  unsigned long long ready0;  // == seq + 1 once alpha of this step is published
  unsigned long long ready1;  // == seq + 1 once beta of this step is published
  unsigned arrive;            // CTAs past phase 1 of this step
  unsigned pad;
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

// ---- one fused launch per iteration ------------------------------------------
// Everything of a CG iteration that is not the operator apply, as ONE
// co-resident (cooperative) grid:
//   prologue  alpha = gamma / p.Ap.  Partitioned: warp 0 of CTA 0 all-reduces
//             this rank's partial p.Ap over peer memory (sfem_common.cuh) and
//             publishes alpha; the other CTAs wait on a flag in L2.
//   phase 1   r -= alpha Ap, partial gamma' = r.(M r) over the owned dofs
//             (3 reads, 1 write); the CTAs arrive at a counter.
//   last CTA  all-reduces gamma' (partitioned), beta = gamma'/gamma, advances
//             the scalars and the convergence flag (cg.py:68-73), zeroes the
//             dot accumulator of the next apply, publishes beta.
//   phase 2   x += alpha p, p = M r + beta p with ONE read of p (4 reads, 2
//             writes: 10 vector passes per iteration instead of 11; every CTA
//             walks ITS chunk backwards: the tail of phase 1 is still in L2),
//             then Ap[0 .. n_zero) = 0: the zero fill of the next apply's
//             shared-dof prefix (Ap is dead here).
// Replaces update + all-reduce + direction + advance + all-reduce + the zero
// fill of the next apply (6 launches, 2 of them NCCL) by one.
constexpr int kStepThreads = 512;

template <typename T> struct Vec;
template <> struct Vec<double> { using type = double2; static constexpr int W = 2; };
template <> struct Vec<float> { using type = float4; static constexpr int W = 4; };

template <typename T>
__device__ __forceinline__ void vload(const T* p, T (&v)[Vec<T>::W]) {
  const typename Vec<T>::type t = *reinterpret_cast<const typename Vec<T>::type*>(p);
  if constexpr (sizeof(T) == 8) { v[0] = t.x; v[1] = t.y; }
  else { v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
}
template <typename T>
__device__ __forceinline__ void vload_cs(const T* p, T (&v)[Vec<T>::W]) {
  const typename Vec<T>::type t =
      __ldcs(reinterpret_cast<const typename Vec<T>::type*>(p));
  if constexpr (sizeof(T) == 8) { v[0] = t.x; v[1] = t.y; }
  else { v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
}
template <typename T>
__device__ __forceinline__ typename Vec<T>::type vpack(const T (&v)[Vec<T>::W]) {
  typename Vec<T>::type t;
  if constexpr (sizeof(T) == 8) { t.x = v[0]; t.y = v[1]; }
  else { t.x = v[0]; t.y = v[1]; t.z = v[2]; t.w = v[3]; }
  return t;
}

// Bounded (~6 s of globaltimer: longer than the 4 s a peer wait may take in
// the publishing CTA): a grid that could not make progress must not hang the
// device.  Returns false on a timeout.
__device__ __forceinline__ bool spin_until(const unsigned long long* flag,
                                           unsigned long long want) {
  uint64_t t0 = 0;
  unsigned spins = 0;
  while (ld_acquire_gpu64(flag) < want) {
    __nanosleep(20);
    if ((++spins & 4095u) == 0) {
      uint64_t now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      if (now - t0 > 6000000000ull) return false;
    }
  }
  return true;
}

template <typename T>
__global__ void __launch_bounds__(kStepThreads)
cg_step_kernel(int64_t n, T* __restrict__ x, T* __restrict__ r,
               T* __restrict__ p, T* __restrict__ Ap,
               const T* __restrict__ minv, const uint8_t* __restrict__ owned,
               CgState* __restrict__ st, const ScalarDev sx, uint64_t epoch0,
               int64_t n_zero, int vec_ok) {
  constexpr int W = Vec<T>::W;
  using VT = typename Vec<T>::type;
  __shared__ double red[32];
  __shared__ int s_last;
  volatile CgState* vst = st;
  if (vst->done != 0.0) return;  // same decision in every CTA (stream order)
  const unsigned long long seq = vst->seq;
  const bool dist = sx.world > 1;

  // ---- alpha
  if (dist) {
    if (blockIdx.x == 0 && threadIdx.x < 32) {
      double v[4] = {vst->pAp, 0.0, 0.0, 0.0};
      const bool ok = scalar_allreduce_warp(sx, epoch0, v, 1);
      if (threadIdx.x == 0) {
        vst->alpha = vst->gamma / v[0];
        if (!ok) vst->done = 2.0;
        __threadfence();
        st_release_gpu(&st->ready0, seq + 1);
      }
    }
    if (threadIdx.x == 0 && !spin_until(&st->ready0, seq + 1))
      vst->done = 2.0;
    __syncthreads();
  }
  const T alpha = dist ? (T)vst->alpha : (T)(vst->gamma / vst->pAp);

  // ---- this CTA's chunk (multiple of W, so vector accesses stay aligned)
  int64_t per = (n + gridDim.x - 1) / gridDim.x;
  per = (per + W - 1) / W * W;
  const int64_t c0 = per * blockIdx.x < n ? per * blockIdx.x : n;
  const int64_t c1 = c0 + per < n ? c0 + per : n;
  // end of the vector part (none if a pointer is not 16-byte aligned)
  const int64_t cv = vec_ok ? c0 + (c1 - c0) / W * W : c0;

  // ---- phase 1: r -= alpha Ap, partial gamma' (x is updated in phase 2, where
  //      p is read anyway: one read of p less per iteration)
  double g = 0.0;
  for (int64_t i = c0 + (int64_t)threadIdx.x * W; i < cv;
       i += (int64_t)kStepThreads * W) {
    T rv[W], av[W], mv[W];
    vload(r + i, rv);
    vload_cs(Ap + i, av);
    if (minv) vload(minv + i, mv);
#pragma unroll
    for (int k = 0; k < W; ++k) {
      rv[k] -= alpha * av[k];
      const T z = minv ? mv[k] * rv[k] : rv[k];
      if (!owned || owned[i + k]) g += (double)rv[k] * (double)z;
    }
    *reinterpret_cast<VT*>(r + i) = vpack<T>(rv);
  }
  for (int64_t i = cv + threadIdx.x; i < c1; i += kStepThreads) {
    const T ri = r[i] - alpha * Ap[i];
    r[i] = ri;
    const T z = minv ? minv[i] * ri : ri;
    if (!owned || owned[i]) g += (double)ri * (double)z;
  }
  g = block_sum(g, red);
  if (threadIdx.x == 0) {
    atomicAdd(&st->gamma_new, g);
    __threadfence();
    s_last = atomicAdd(&st->arrive, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (s_last && threadIdx.x < 32) {
    __threadfence();
    double v[4] = {__ldcg(&st->gamma_new), 0.0, 0.0, 0.0};
    bool ok = true;
    if (dist) ok = scalar_allreduce_warp(sx, epoch0 + 1, v, 1);
    if (threadIdx.x == 0) {
      const double gamma = vst->gamma;
      vst->beta = v[0] / gamma;
      vst->gamma = v[0];
      vst->gamma_new = 0.0;
      vst->pAp = 0.0;
      vst->k = vst->k + 1.0;
      // a peer that never answered (here or in the prologue) is fatal: the
      // host raises when it reads done == 2
      vst->done = (!ok || vst->done == 2.0)
                      ? 2.0
                      : ((v[0] > vst->atol2 && vst->k < vst->maxiter) ? 0.0
                                                                       : 1.0);
      vst->arrive = 0u;
      vst->seq = seq + 1;
      __threadfence();
      st_release_gpu(&st->ready1, seq + 1);
    }
  }
  if (threadIdx.x == 0 && !spin_until(&st->ready1, seq + 1)) vst->done = 2.0;
  __syncthreads();
  const T beta = (T)vst->beta;

  // ---- phase 2 (backwards over the chunk): x += alpha p (the OLD direction),
  //      then p = M r + beta p
  for (int64_t i = c1 - 1 - threadIdx.x; i >= cv; i -= kStepThreads) {
    const T ri = r[i], pi = p[i];
    x[i] += alpha * pi;
    p[i] = (minv ? minv[i] * ri : ri) + beta * pi;
  }
  const int64_t nvec = (cv - c0) / W;
  for (int64_t j = nvec - 1 - threadIdx.x; j >= 0; j -= kStepThreads) {
    const int64_t i = c0 + j * W;
    T rv[W], pv[W], mv[W], xv[W];
    vload(r + i, rv);
    vload(p + i, pv);
    vload_cs(x + i, xv);
    if (minv) vload(minv + i, mv);
#pragma unroll
    for (int k = 0; k < W; ++k) {
      xv[k] += alpha * pv[k];
      pv[k] = (minv ? mv[k] * rv[k] : rv[k]) + beta * pv[k];
    }
    __stcs(reinterpret_cast<VT*>(x + i), vpack<T>(xv));
    *reinterpret_cast<VT*>(p + i) = vpack<T>(pv);
  }
  // ---- zero fill of the next apply's shared-dof prefix
  const int64_t zv = vec_ok ? n_zero / W * W : 0;
  T zero[W];
#pragma unroll
  for (int k = 0; k < W; ++k) zero[k] = T(0);
  for (int64_t i = ((int64_t)blockIdx.x * kStepThreads + threadIdx.x) * W;
       i < zv; i += (int64_t)gridDim.x * kStepThreads * W)
    *reinterpret_cast<VT*>(Ap + i) = vpack<T>(zero);
  for (int64_t i = zv + (int64_t)blockIdx.x * kStepThreads + threadIdx.x;
       i < n_zero; i += (int64_t)gridDim.x * kStepThreads)
    Ap[i] = T(0);
}

template <typename T>
int launch_cg_step(int64_t n, void* x, void* r, void* p, void* Ap,
                   const void* minv, const uint8_t* owned, CgState* st,
                   sfem_scalar_exchange* sx, int64_t n_zero,
                   cudaStream_t stream) {
  auto kernel = cg_step_kernel<T>;
  int dev = 0;
  SFEM_CUDA_CHECK(cudaGetDevice(&dev));
  static int per_sm_dev[64] = {};
  int& per_sm = per_sm_dev[dev & 63];
  if (per_sm == 0) {
    SFEM_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
        &per_sm, kernel, kStepThreads, 0));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 4) per_sm = 4;
  }
  const int64_t cap = (int64_t)num_sms() * per_sm;
  int64_t blocks = (n + (int64_t)kStepThreads * 8 - 1) / ((int64_t)kStepThreads * 8);
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  const int vec_ok = (((uintptr_t)x | (uintptr_t)r | (uintptr_t)p |
                       (uintptr_t)Ap | (uintptr_t)minv) & 15) == 0;
  uint64_t epoch0 = 0;
  if (sx && sx->world > 1) {
    epoch0 = sx->epoch + 1;
    sx->epoch += 2;
  }
  const ScalarDev sd = scalar_view(sx);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)blocks);
  cfg.blockDim = dim3(kStepThreads);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  SFEM_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kernel, n, (T*)x, (T*)r, (T*)p,
                                     (T*)Ap, (const T*)minv, owned, st, sd,
                                     epoch0, n_zero, vec_ok));
  SFEM_LAUNCH_CHECK();
  return SFEM_OK;
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

// `iters` iterations: apply (+ halo exchange) and ONE fused vector launch each.
// The first apply zero-fills its own prefix; every step zeroes it for the next.
template <typename T>
int cg_iterate_impl(const sfem_op* op, sfem_halo* halo,
                    sfem_scalar_exchange* sx, double lambda, double mu,
                    int64_t num_interface_elements, int ncomp, void* x, void* r,
                    void* p, void* Ap, const void* minv, const uint8_t* owned,
                    CgState* st, int iters, cudaStream_t stream) {
  const int64_t n = op->base.desc.num_nodes * (int64_t)ncomp;
  // lazy zero fill: the apply's companion kernel zeroes Ap's prefix while the
  // apply runs, so the step kernel does not
  const bool lazy = !halo && lazy_zero_applicable(op, ncomp, stream);
  const int64_t n_zero = lazy ? 0 : op->n_zero * (int64_t)ncomp;
  for (int it = 0; it < iters; ++it) {
    int rc;
    if (halo) {
      rc = op_apply_halo_internal(op, halo, lambda, mu, p, Ap,
                                  num_interface_elements, &st->pAp, it > 0,
                                  stream);
      if (!rc) rc = sfem_halo_wait_unpack(halo, Ap, (sfem_stream_t)stream);
    } else {
      // (every step kernel resets p.Ap for the next apply)
      rc = op_apply_internal(op, lambda, mu, p, Ap, ncomp, &st->pAp, stream,
                             it > 0 && !lazy, it > 0);
    }
    if (rc) return rc;
    rc = launch_cg_step<T>(n, x, r, p, Ap, minv, owned, st, sx, n_zero, stream);
    if (rc) return rc;
  }
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
  // vectors 16-byte aligned (see sfem_cg_workspace_bytes)
  const int64_t stride = (n * (int64_t)sizeof(T) + 15) / 16 * 16;
  T* r = (T*)(ws + 256);
  T* p = (T*)(ws + 256 + stride);
  T* Ap = (T*)(ws + 256 + 2 * stride);

  int rc = init_state(st, prm->tol, prm->atol,
                      (double)(prm->maxiter >= 0 ? prm->maxiter : 10 * n),
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
    // never enqueue more applies than iterations are left
    double left = hs.maxiter - hs.k;
    int iters = left < (double)check_every ? (int)left : check_every;
    if (iters < 1) iters = 1;
    if (n == 0) break;
    rc = cg_iterate_impl<T>(op, nullptr, nullptr, prm->lambda, prm->mu, 0,
                            ncomp, x, r, p, Ap, minv, nullptr, st, iters,
                            stream);
    if (rc) return rc;
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
  // state + r, p, Ap, each starting 16-byte aligned
  return 256 + 3 * ((size * esz + 15) / 16 * 16);
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
                      (double)(maxiter >= 0 ? maxiter : 10 * n), stream);
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

int sfem_cg_iterate(const sfem_op* op, sfem_halo* halo,
                    sfem_scalar_exchange* sx, double lambda, double mu,
                    int64_t num_interface_elements, int32_t ncomp, void* x,
                    void* r, void* p, void* Ap, const void* minv,
                    const uint8_t* owned, void* state, int32_t iters,
                    sfem_stream_t stream) {
  using namespace sfem;
  SFEM_REQUIRE(op && x && r && p && Ap && state, "null argument");
  SFEM_REQUIRE(ncomp >= 1 && (halo == nullptr || ncomp == 1), "bad ncomp");
  SFEM_REQUIRE(iters >= 0, "negative iteration count");
  SFEM_REQUIRE(halo != nullptr || sx == nullptr || sx->world == 1,
               "a scalar exchange needs the halo of the same partition");
  if (op->base.desc.num_nodes == 0) return SFEM_OK;
  return op->base.desc.dtype == SFEM_F64
             ? cg_iterate_impl<double>(op, halo, sx, lambda, mu,
                                       num_interface_elements, ncomp, x, r, p,
                                       Ap, minv, owned, (CgState*)state, iters,
                                       (cudaStream_t)stream)
             : cg_iterate_impl<float>(op, halo, sx, lambda, mu,
                                      num_interface_elements, ncomp, x, r, p,
                                      Ap, minv, owned, (CgState*)state, iters,
                                      (cudaStream_t)stream);
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
  *done = (int32_t)hs.done;  // 0 running, 1 finished, 2 peer wait timed out
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
