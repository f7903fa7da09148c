// Collocated-GLL operator apply: the hot kernels (2-D quads, 3-D hexes).
//
//   y = mask . Z^T [ lambda W detJ u + mu sum_i D_i^T ( sum_k Gf_ik D_k Z x ) ]
//
// fused gather -> sum-factorised gradient -> geometric factors -> transposed
// gradient -> scatter (+ optional x.y dot product), one pass over the packed
// connectivity (4 B/local node) and the symmetric geometric factors
// (d(d+1)/2 [+1] values/local node).  Quadrature nodes == grid nodes, so the
// interpolation matrix is the identity and G_i = I x .. x D x .. x I
// (reference: swirl_fem/core/interpolation.py:257-258 shortcut, :265-286;
// swirl_fem/core/fespace.py:190-195, 401-403, 458-471;
// swirl_fem/navier_stokes/navier_stokes.py:174-180 uses exactly this space).
//
// 2-D: one thread per node, several elements per CTA.
// 3-D: N x N threads per element sweep the slowest axis (a0) keeping the
//      a0-column of u and y in registers; the a0-derivative and its transpose
//      use D[k][m] with compile-time (k, m) -> constant-bank operands; the
//      in-plane derivatives go through a double-buffered shared-memory slice.

#pragma once
#include "sfem_common.cuh"
#include "sfem_apply3d_v2.cuh"
#include "sfem_apply2d_v2.cuh"
#include "sfem_apply2d_warp.cuh"

namespace sfem {

namespace {

template <typename T, int N>
struct DMat {
  T d[N * N];  // row-major D[i][j] = l_j'(x_i)
};

template <typename T>
__device__ __forceinline__ void store_result(T* __restrict__ y, uint32_t cn,
                                             int ncomp, int c, T v) {
  if (cn == kConnSentinel) return;
  T* dst = y + (int64_t)(cn & kConnIdMask) * ncomp + c;
  if (cn & kConnDirichlet) {
    if (cn & kConnSingle) *dst = T(0);
  } else if (cn & kConnSingle) {
    *dst = v;
  } else {
    red_add(dst, v);
  }
}

// ---------------------------------------------------------------------------
// 2-D
// ---------------------------------------------------------------------------
template <int N>
struct Cfg2D {
  static constexpr int n = N * N;
  static constexpr int epb = (256 / n) > 0 ? (256 / n) : 1;  // elements / CTA
  static constexpr int threads = ((epb * n + 31) / 32) * 32;
  static constexpr int ld = N + 1;  // padded row
};

template <typename T, int N, bool MASS, bool LOCAL>
__global__ void __launch_bounds__(Cfg2D<N>::threads)
apply2d_kernel(const __grid_constant__ DMat<T, N> dm,
               const uint32_t* __restrict__ conn, const T* __restrict__ gf,
               T lambda, T mu, const T* __restrict__ x, T* __restrict__ y,
               int ncomp, int64_t E, double* __restrict__ dot_xy) {
  using C = Cfg2D<N>;
  constexpr int n = C::n, epb = C::epb, ld = C::ld;
  constexpr int ngeom = MASS ? 4 : 3;
  __shared__ T sD[N * N];
  __shared__ T sU[epb][N][ld];
  __shared__ T sR[epb][N][ld];
  __shared__ T sS[epb][N][ld];
  __shared__ double red[32];

  for (int t = threadIdx.x; t < N * N; t += blockDim.x) sD[t] = dm.d[t];
  const int slot = threadIdx.x / n;
  const int node = threadIdx.x - slot * n;
  const int i = node / N, j = node - i * N;
  const int c = blockIdx.y;
  const bool lane_ok = slot < epb;
  double dot = 0.0;

  const int64_t nblocks = (E + epb - 1) / epb;
  for (int64_t blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
    const int64_t e = blk * epb + slot;
    const bool active = lane_ok && e < E;
    uint32_t cn = kConnSentinel;
    T u = T(0);
    T g00 = T(0), g01 = T(0), g11 = T(0), wd = T(0);
    if (active) {
      const int64_t ln = e * n + node;
      if (LOCAL) {
        u = x[ln * ncomp + c];
      } else {
        cn = ld_stream(conn + ln);
        if (cn != kConnSentinel)
          u = __ldg(x + (int64_t)(cn & kConnIdMask) * ncomp + c);
      }
      const T* g = gf + e * (int64_t)(ngeom * n) + node;
      g00 = ld_stream(g);
      g01 = ld_stream(g + n);
      g11 = ld_stream(g + 2 * n);
      if (MASS) wd = ld_stream(g + 3 * n);
    }
    __syncthreads();  // previous iteration's reads of sR/sS are done
    if (lane_ok) sU[slot][i][j] = u;
    __syncthreads();
    T ur = T(0), us = T(0);
    if (lane_ok) {
#pragma unroll
      for (int m = 0; m < N; ++m) {
        ur += sD[i * N + m] * sU[slot][m][j];
        us += sD[j * N + m] * sU[slot][i][m];
      }
      sR[slot][i][j] = mu * (g00 * ur + g01 * us);
      sS[slot][i][j] = mu * (g01 * ur + g11 * us);
    }
    __syncthreads();
    if (active) {
      T v = T(0);
#pragma unroll
      for (int m = 0; m < N; ++m) {
        v += sD[m * N + i] * sR[slot][m][j];
        v += sD[m * N + j] * sS[slot][i][m];
      }
      if (MASS) v += lambda * wd * u;
      if (LOCAL) {
        y[(e * n + node) * ncomp + c] = v;
      } else {
        store_result<T>(y, cn, ncomp, c, v);
        if (cn != kConnSentinel && !(cn & kConnDirichlet))
          dot += (double)u * (double)v;
      }
    }
  }
  if (!LOCAL && dot_xy != nullptr) {
    dot = block_sum(dot, red);
    if (threadIdx.x == 0) atomicAdd(dot_xy, dot);
  }
}

// ---------------------------------------------------------------------------
// 3-D
// ---------------------------------------------------------------------------
template <int N>
struct Cfg3D {
  static constexpr int nn = N * N;
  static constexpr int n = N * N * N;
  // elements per CTA: aim for ~256 threads, at most 8 slots
  static constexpr int epb_raw = 256 / nn;
  static constexpr int epb = epb_raw < 1 ? 1 : (epb_raw > 8 ? 8 : epb_raw);
  static constexpr int threads = ((epb * nn + 31) / 32) * 32;
  static constexpr int ld = N + 1;
  // D rows/columns of the in-plane axes live in registers for small N
  static constexpr bool dreg = N <= 8;
};

template <typename T, int N, bool MASS, bool LOCAL>
__global__ void __launch_bounds__(Cfg3D<N>::threads)
apply3d_kernel(const __grid_constant__ DMat<T, N> dm,
               const uint32_t* __restrict__ conn, const T* __restrict__ gf,
               T lambda, T mu, const T* __restrict__ x, T* __restrict__ y,
               int ncomp, int64_t E, double* __restrict__ dot_xy) {
  using C = Cfg3D<N>;
  constexpr int nn = C::nn, n = C::n, epb = C::epb, ld = C::ld;
  constexpr int ngeom = MASS ? 7 : 6;
  __shared__ T sD[N * N];
  __shared__ T sU[2][epb][N][ld];
  __shared__ T sW1[2][epb][N][ld];
  __shared__ T sW2[2][epb][N][ld];
  __shared__ double red[32];

  for (int t = threadIdx.x; t < N * N; t += blockDim.x) sD[t] = dm.d[t];
  __syncthreads();
  const int slot = threadIdx.x / nn;
  const int p = threadIdx.x - slot * nn;
  const int a1 = p / N, a2 = p - a1 * N;
  const int c = blockIdx.y;
  const bool lane_ok = slot < epb;
  double dot = 0.0;

  // in-plane rows / columns of D for this thread
  T dr1[C::dreg ? N : 1], dr2[C::dreg ? N : 1];  // D[a1][m], D[a2][m]
  T dc1[C::dreg ? N : 1], dc2[C::dreg ? N : 1];  // D[m][a1], D[m][a2]
  if (C::dreg) {
#pragma unroll
    for (int m = 0; m < N; ++m) {
      dr1[m] = sD[(lane_ok ? a1 : 0) * N + m];
      dr2[m] = sD[(lane_ok ? a2 : 0) * N + m];
      dc1[m] = sD[m * N + (lane_ok ? a1 : 0)];
      dc2[m] = sD[m * N + (lane_ok ? a2 : 0)];
    }
  }

  const int64_t nblocks = (E + epb - 1) / epb;
  for (int64_t blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
    const int64_t e = blk * epb + slot;
    const bool active = lane_ok && e < E;
    T ru[N], ry[N];
    uint32_t rc[N];
    if (active) {
#pragma unroll
      for (int k = 0; k < N; ++k) {
        const int64_t ln = e * n + k * nn + p;
        if (LOCAL) {
          rc[k] = 0;
          ru[k] = x[ln * ncomp + c];
        } else {
          rc[k] = ld_stream(conn + ln);
          ru[k] = rc[k] == kConnSentinel
                      ? T(0)
                      : __ldg(x + (int64_t)(rc[k] & kConnIdMask) * ncomp + c);
        }
        ry[k] = T(0);
      }
    } else {
#pragma unroll
      for (int k = 0; k < N; ++k) {
        ru[k] = T(0);
        ry[k] = T(0);
        rc[k] = kConnSentinel;
      }
    }
    const T* gbase = gf + (active ? e : 0) * (int64_t)(ngeom * n) + p;

#pragma unroll
    for (int k = 0; k < N; ++k) {
      const int b = k & 1;
      if (lane_ok) sU[b][slot][a1][a2] = ru[k];
      // geometric factors of this slab: issue the loads before the barrier
      T g00 = T(0), g01 = T(0), g02 = T(0), g11 = T(0), g12 = T(0),
        g22 = T(0), wd = T(0);
      if (active) {
        const T* g = gbase + k * nn;
        g00 = ld_stream(g);
        g01 = ld_stream(g + n);
        g02 = ld_stream(g + 2 * n);
        g11 = ld_stream(g + 3 * n);
        g12 = ld_stream(g + 4 * n);
        g22 = ld_stream(g + 5 * n);
        if (MASS) wd = ld_stream(g + 6 * n);
      }
      __syncthreads();
      T d0 = T(0), d1 = T(0), d2 = T(0);
      if (lane_ok) {
#pragma unroll
        for (int m = 0; m < N; ++m) {
          d0 += dm.d[k * N + m] * ru[m];
          if (C::dreg) {
            d1 += dr1[m] * sU[b][slot][m][a2];
            d2 += dr2[m] * sU[b][slot][a1][m];
          } else {
            d1 += sD[a1 * N + m] * sU[b][slot][m][a2];
            d2 += sD[a2 * N + m] * sU[b][slot][a1][m];
          }
        }
        const T w0 = mu * (g00 * d0 + g01 * d1 + g02 * d2);
        const T w1 = mu * (g01 * d0 + g11 * d1 + g12 * d2);
        const T w2 = mu * (g02 * d0 + g12 * d1 + g22 * d2);
#pragma unroll
        for (int m = 0; m < N; ++m) ry[m] += dm.d[k * N + m] * w0;
        sW1[b][slot][a1][a2] = w1;
        sW2[b][slot][a1][a2] = w2;
      }
      __syncthreads();
      if (lane_ok) {
        T v = T(0);
#pragma unroll
        for (int m = 0; m < N; ++m) {
          if (C::dreg) {
            v += dc1[m] * sW1[b][slot][m][a2];
            v += dc2[m] * sW2[b][slot][a1][m];
          } else {
            v += sD[m * N + a1] * sW1[b][slot][m][a2];
            v += sD[m * N + a2] * sW2[b][slot][a1][m];
          }
        }
        if (MASS) v += lambda * wd * ru[k];
        ry[k] += v;
      }
    }
    if (active) {
#pragma unroll
      for (int k = 0; k < N; ++k) {
        if (LOCAL) {
          y[(e * n + k * nn + p) * ncomp + c] = ry[k];
        } else {
          store_result<T>(y, rc[k], ncomp, c, ry[k]);
          if (rc[k] != kConnSentinel && !(rc[k] & kConnDirichlet))
            dot += (double)ru[k] * (double)ry[k];
        }
      }
    }
    // no trailing barrier: the double-buffered slices are only rewritten
    // after two further barriers (see the hazard analysis in DESIGN.md).
  }
  if (!LOCAL && dot_xy != nullptr) {
    dot = block_sum(dot, red);
    if (threadIdx.x == 0) atomicAdd(dot_xy, dot);
  }
}

template <typename T, int N>
DMat<T, N> make_dmat(const SpaceBase& b) {
  // collocated: BD == D (B is the identity)
  DMat<T, N> m;
  for (int i = 0; i < N * N; ++i) m.d[i] = (T)b.h_BD[i];
  return m;
}

template <typename T, int N, bool MASS, bool LOCAL>
int launch2d(const sfem_op& op, double lambda, double mu, const void* x,
             void* y, int ncomp, double* dot_xy, cudaStream_t stream) {
  // Two fused kernels with the same arithmetic: the block-synchronous
  // two-mapping kernel (v2) and the warp-autonomous one (v3).  Measured on B200
  // at 16 M dofs (profiles/r02_2d_warp_vs_block.txt): v2 wins by 2-20 % up to
  // N = 10 (both precisions), v3 from N = 12 on (+3 % at N = 12, +42 % fp64 /
  // +26 % fp32 at N = 16, where v2 fits one or two CTAs per SM).  Variant 3 /
  // 4 force v2 / v3; variant 2 is the thread-per-node kernel (v1).
  constexpr bool warp_default = N >= 12;
  if (op.variant == 3 || (op.variant != 2 && op.variant != 4 && !warp_default))
    return launch2d_v2<T, N, MASS, LOCAL>(op, lambda, mu, x, y, ncomp, dot_xy,
                                          stream);
  if (op.variant != 2)
    return launch2d_warp<T, N, MASS, LOCAL>(op, lambda, mu, x, y, ncomp,
                                            dot_xy, stream);
  using C = Cfg2D<N>;
  const int64_t E = op.base.desc.num_elements;
  const int64_t nblocks = (E + C::epb - 1) / C::epb;
  const int64_t cap = (int64_t)num_sms() * 8;
  dim3 grid((unsigned)(nblocks < cap ? nblocks : cap), ncomp);
  apply2d_kernel<T, N, MASS, LOCAL><<<grid, C::threads, 0, stream>>>(
      make_dmat<T, N>(op.base), op.conn, (const T*)op.geom, (T)lambda, (T)mu,
      (const T*)x, (T*)y, ncomp, E, dot_xy);
  SFEM_LAUNCH_CHECK();
  return SFEM_OK;
}

template <typename T, int N, bool MASS, bool LOCAL>
int launch3d(const sfem_op& op, double lambda, double mu, const void* x,
             void* y, int ncomp, double* dot_xy, cudaStream_t stream) {
  // variant 2 selects the slab-sweep kernel (v1); default is the
  // three-mapping kernel (v2)
  if (op.variant != 2)
    return launch3d_v2<T, N, MASS, LOCAL>(op, lambda, mu, x, y, ncomp, dot_xy,
                                          stream);
  using C = Cfg3D<N>;
  const int64_t E = op.base.desc.num_elements;
  const int64_t nblocks = (E + C::epb - 1) / C::epb;
  const int64_t cap = (int64_t)num_sms() * 8;
  dim3 grid((unsigned)(nblocks < cap ? nblocks : cap), ncomp);
  apply3d_kernel<T, N, MASS, LOCAL><<<grid, C::threads, 0, stream>>>(
      make_dmat<T, N>(op.base), op.conn, (const T*)op.geom, (T)lambda, (T)mu,
      (const T*)x, (T*)y, ncomp, E, dot_xy);
  SFEM_LAUNCH_CHECK();
  return SFEM_OK;
}

template <typename T, int DIM, int N>
int dispatch_n(const sfem_op& op, double lambda, double mu, const void* x,
               void* y, int ncomp, bool local, double* dot_xy,
               cudaStream_t stream) {
  // MASS selects the memory layout of the geometric factors (ngeom), so it
  // follows the handle, not the value of lambda.
  const bool mass = op.with_mass != 0;
#define SFEM_GO(FN)                                                            \
  if (mass) {                                                                  \
    return local ? FN<T, N, true, true>(op, lambda, mu, x, y, ncomp, dot_xy,   \
                                        stream)                                \
                 : FN<T, N, true, false>(op, lambda, mu, x, y, ncomp, dot_xy,  \
                                         stream);                              \
  } else {                                                                     \
    return local ? FN<T, N, false, true>(op, lambda, mu, x, y, ncomp, dot_xy,  \
                                         stream)                               \
                 : FN<T, N, false, false>(op, lambda, mu, x, y, ncomp, dot_xy, \
                                          stream);                             \
  }
  if constexpr (DIM == 2) {
    SFEM_GO(launch2d)
  } else {
    SFEM_GO(launch3d)
  }
#undef SFEM_GO
}

}  // namespace

// Returns SFEM_ERR_UNSUPPORTED when no specialised kernel exists (the caller
// then uses the generic kernel).
template <typename T, int DIM>
int launch_apply_colloc_dim(const sfem_op& op, double lambda, double mu,
                            const void* x, void* y, int ncomp, bool local,
                            double* dot_xy, cudaStream_t stream) {
  const sfem_space_desc& d = op.base.desc;
  if (d.num_elements == 0) return SFEM_OK;
  switch (d.n1d) {
#define SFEM_CASE(NN)                                                         \
  case NN:                                                                    \
    return dispatch_n<T, DIM, NN>(op, lambda, mu, x, y, ncomp, local, dot_xy, \
                                  stream);
    SFEM_CASE(2)
    SFEM_CASE(3)
    SFEM_CASE(4)
    SFEM_CASE(5)
    SFEM_CASE(6)
    SFEM_CASE(7)
    SFEM_CASE(8)
    SFEM_CASE(9)
    SFEM_CASE(10)
    SFEM_CASE(11)
    SFEM_CASE(12)
    SFEM_CASE(13)
    SFEM_CASE(14)
    SFEM_CASE(15)
    SFEM_CASE(16)
#undef SFEM_CASE
    default:
      return SFEM_ERR_UNSUPPORTED;
  }
}

}  // namespace sfem
