// 3-D collocated-GLL operator apply, "three-mapping" kernel (v2).
//
// N^2 threads per element.  The same thread index (p, q) = (t / N, t % N) is
// interpreted under three mappings, one per tensor axis, so that every 1-D
// contraction runs entirely in registers against D[i][j] with COMPILE-TIME
// (i, j) -- uniform-register / constant-bank operands, no per-thread registers
// or shared-memory traffic for the derivative matrix:
//   mapping A: thread owns the a0-column  u[:, p, q]
//   mapping B: thread owns the a1-column  u[p, :, q]
//   mapping C: thread owns the a2-column  u[p, q, :]
// Columns are exchanged through padded shared-memory tiles (u double-buffered,
// two work tiles); per element there are 4 block barriers (the slab-sweep
// kernel v1 needs 2N) and every contraction exposes N independent FMA chains.
// Nothing but the connectivity words and the a0-part of the result lives in
// registers across barriers, which keeps the fp64 kernel at <= 128 registers
// (2 CTAs of 256 threads per SM).
//
// Memory pipeline (persistent CTAs, one wave):
//   * the gather x[idx] of the CTA's NEXT element is issued with cp.async
//     (LDGSTS) straight into the other u tile while the current element is
//     computed -- no registers, latency fully overlapped;
//   * the next element's geometric factors are prefetched into L2
//     (prefetch.global.L2) a whole element ahead; phase 3 then streams them
//     from L2 in batches of KCH slabs;
//   * the next element's connectivity words are loaded at the top of the
//     iteration and consumed after the second barrier.
#pragma once

#include "sfem_common.cuh"

namespace sfem {
namespace {

template <typename T, int N>
struct DMat3 {
  T d[N * N];  // row-major D[i][j] = l_j'(x_i)
};

// EPB: elements per CTA; MINB: min CTAs per SM (register cap); KCH: slabs of
// geometric factors loaded per batch in phase 3 (bounds registers in flight).
template <typename T, int N, int EPB, int MINB, int KCH>
struct Cfg3DV2 {
  static constexpr int P = N * N;  // threads per element
  static constexpr int n = N * N * N;
  static constexpr int epb = EPB;
  static constexpr int threads = ((EPB * P + 31) / 32) * 32;
  // padded strides: S1 odd; for 8-byte words consecutive a0 planes are offset
  // by half the banks
  static constexpr int S1 = (N % 2 == 0) ? N + 1 : N;
  static constexpr int S0_raw = N * S1;
  static constexpr int S0 =
      sizeof(T) == 8 ? S0_raw + ((8 - (S0_raw % 16)) + 16) % 16 : S0_raw;
  static constexpr int tile = N * S0;
  static constexpr int tiles_per_slot = 4;  // u[2], A, B
  static constexpr int min_blocks = MINB;
  static constexpr int kch = KCH;
};

template <int N>
constexpr int default_epb() {
  return (256 / (N * N)) < 1 ? 1 : ((256 / (N * N)) > 8 ? 8 : (256 / (N * N)));
}

__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

template <typename T>
__device__ __forceinline__ void cp_async_elem(T* smem_dst, const T* gsrc) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(s), "l"(gsrc),
               "n"(sizeof(T)));
}
__device__ __forceinline__ void cp_async_commit() {
  asm volatile("cp.async.commit_group;");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

template <typename T, int N, bool MASS, bool LOCAL, int EPB, int MINB, int KCH, int JU>
__global__ void __launch_bounds__((Cfg3DV2<T, N, EPB, MINB, KCH>::threads),
                                  MINB)
apply3d_v2_kernel(const __grid_constant__ DMat3<T, N> dm,
                  const uint32_t* __restrict__ conn,
                  const T* __restrict__ gf, T lambda, T mu,
                  const T* __restrict__ x, T* __restrict__ y, int ncomp,
                  int64_t E, double* __restrict__ dot_xy) {
  using C = Cfg3DV2<T, N, EPB, MINB, KCH>;
  constexpr int P = C::P, n = C::n, epb = C::epb;
  constexpr int S0 = C::S0, S1 = C::S1;
  constexpr int ngeom = MASS ? 7 : 6;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double red[32];
  T* smem = reinterpret_cast<T*>(smem_raw);

  const int slot = threadIdx.x / P;
  const int t = threadIdx.x - slot * P;
  const int p = t / N, q = t - p * N;
  const int c = blockIdx.y;
  const bool lane_ok = slot < epb;
  T* sU0 = smem + (lane_ok ? slot : 0) * C::tiles_per_slot * C::tile;
  T* sA = sU0 + 2 * C::tile;
  T* sB = sA + C::tile;
  const int offA = p * S1 + q;       // + k * S0      (mapping A)
  const int offB = p * S0 + q;       // + m * S1      (mapping B)
  const int offC = p * S0 + q * S1;  // + m           (mapping C)
  const bool want_dot = !LOCAL && dot_xy != nullptr;
  double dot = 0.0;

  const int64_t nblocks = (E + epb - 1) / epb;
  int64_t blk = blockIdx.x;
  int64_t e = blk * epb + slot;
  bool active = lane_ok && blk < nblocks && e < E;
  uint32_t rc[N];

  // prologue: gather of the first element into u tile 0
#pragma unroll
  for (int k = 0; k < N; ++k) {
    rc[k] = kConnSentinel;
    if (lane_ok) {
      T* dst = sU0 + k * S0 + offA;
      if (active) {
        const int64_t ln = e * n + k * P + t;
        if (LOCAL) {
          rc[k] = 0;
          cp_async_elem(dst, x + ln * ncomp + c);
        } else {
          rc[k] = ld_stream(conn + ln);
          if (rc[k] != kConnSentinel)
            cp_async_elem(dst, x + (int64_t)(rc[k] & kConnIdMask) * ncomp + c);
          else
            *dst = T(0);
        }
      } else {
        *dst = T(0);
      }
    }
  }
  cp_async_commit();

  int buf = 0;
  for (; blk < nblocks; blk += gridDim.x, buf ^= 1) {
    T* sU = sU0 + buf * C::tile;
    T* sUn = sU0 + (buf ^ 1) * C::tile;
    // ---- pipeline: next element's factors -> L2, connectivity -> registers
    const int64_t blk_n = blk + gridDim.x;
    const int64_t e_n = blk_n * epb + slot;
    const bool active_n = lane_ok && blk_n < nblocks && e_n < E;
    uint32_t nrc[N];
#pragma unroll
    for (int k = 0; k < N; ++k) nrc[k] = kConnSentinel;
    if (active_n) {
      const char* g =
          reinterpret_cast<const char*>(gf + e_n * (int64_t)(ngeom * n));
      constexpr int lines = (ngeom * n * (int)sizeof(T) + 127) / 128;
#pragma unroll
      for (int l = 0; l < (lines + P - 1) / P; ++l)
        if (l * P + t < lines) prefetch_l2(g + (int64_t)(l * P + t) * 128);
      if (!LOCAL) {
#pragma unroll
        for (int k = 0; k < N; ++k)
          nrc[k] = ld_stream(conn + e_n * n + k * P + t);
      }
    }

    // ---- u tile of this element has landed (cp.async issued one element ago)
    cp_async_wait_all();
    __syncthreads();

    // ---- phase 2 (mappings B, C): a1- and a2-derivatives
    if (lane_ok) {
      T col[N];
#pragma unroll
      for (int m = 0; m < N; ++m) col[m] = sU[offB + m * S1];
#pragma unroll(JU)
      for (int j = 0; j < N; ++j) {
        T acc = T(0);
#pragma unroll
        for (int m = 0; m < N; ++m) acc += dm.d[j * N + m] * col[m];
        sA[offB + j * S1] = acc;
      }
#pragma unroll
      for (int m = 0; m < N; ++m) col[m] = sU[offC + m];
#pragma unroll(JU)
      for (int j = 0; j < N; ++j) {
        T acc = T(0);
#pragma unroll
        for (int m = 0; m < N; ++m) acc += dm.d[j * N + m] * col[m];
        sB[offC + j] = acc;
      }
    }
    __syncthreads();

    // ---- issue the gather of the next element into the other u tile (its
    //      previous contents were last read before the barrier above)
    if (lane_ok) {
#pragma unroll
      for (int k = 0; k < N; ++k) {
        T* dst = sUn + k * S0 + offA;
        if (active_n) {
          if (LOCAL)
            cp_async_elem(dst, x + (e_n * n + k * P + t) * ncomp + c);
          else if (nrc[k] != kConnSentinel)
            cp_async_elem(dst, x + (int64_t)(nrc[k] & kConnIdMask) * ncomp + c);
          else
            *dst = T(0);
        } else {
          *dst = T(0);
        }
      }
    }
    cp_async_commit();

    // ---- phase 3 (mapping A): a0-derivative, geometric factors, transposed
    //      a0-derivative
    T ry[N];
#pragma unroll
    for (int k = 0; k < N; ++k) ry[k] = T(0);
    if (lane_ok) {
      T d0[N];
      {
        T col[N];
#pragma unroll
        for (int m = 0; m < N; ++m) col[m] = sU[m * S0 + offA];
#pragma unroll
        for (int k = 0; k < N; ++k) {
          T acc = T(0);
#pragma unroll
          for (int m = 0; m < N; ++m) acc += dm.d[k * N + m] * col[m];
          d0[k] = acc;
          if (MASS) ry[k] = col[k];  // u, scaled by lambda W detJ below
        }
      }
      const T* g = gf + (active ? e : 0) * (int64_t)(ngeom * n) + t;
#pragma unroll
      for (int k0 = 0; k0 < N; k0 += KCH) {
        T gg[KCH][7];
#pragma unroll
        for (int kk = 0; kk < KCH; ++kk) {
          const int k = k0 + kk;
          if (k < N) {
#pragma unroll
            for (int s = 0; s < ngeom; ++s)
              gg[kk][s] = active ? ld_stream(g + k * P + s * n) : T(0);
          }
        }
#pragma unroll
        for (int kk = 0; kk < KCH; ++kk) {
          const int k = k0 + kk;
          if (k < N) {
            const T d1 = sA[k * S0 + offA];
            const T d2 = sB[k * S0 + offA];
            const T w0 = mu * (gg[kk][0] * d0[k] + gg[kk][1] * d1 +
                               gg[kk][2] * d2);
            sA[k * S0 + offA] = mu * (gg[kk][1] * d0[k] + gg[kk][3] * d1 +
                                      gg[kk][4] * d2);
            sB[k * S0 + offA] = mu * (gg[kk][2] * d0[k] + gg[kk][4] * d1 +
                                      gg[kk][5] * d2);
            d0[k] = w0;
            if (MASS) ry[k] *= lambda * gg[kk][6];
          }
        }
        // keep the compiler from hoisting every slab's loads to the top
        asm volatile("" ::: "memory");
      }
#pragma unroll
      for (int m = 0; m < N; ++m) {
        T acc = ry[m];
#pragma unroll
        for (int k = 0; k < N; ++k) acc += dm.d[k * N + m] * d0[k];
        ry[m] = acc;
      }
    }
    __syncthreads();

    // ---- phase 4 (mappings B, C): transposed a1-, a2-derivatives, in place
    if (lane_ok) {
      T col[N];
#pragma unroll
      for (int m = 0; m < N; ++m) col[m] = sA[offB + m * S1];
#pragma unroll(JU)
      for (int j = 0; j < N; ++j) {
        T acc = T(0);
#pragma unroll
        for (int m = 0; m < N; ++m) acc += dm.d[m * N + j] * col[m];
        sA[offB + j * S1] = acc;
      }
#pragma unroll
      for (int m = 0; m < N; ++m) col[m] = sB[offC + m];
#pragma unroll(JU)
      for (int j = 0; j < N; ++j) {
        T acc = T(0);
#pragma unroll
        for (int m = 0; m < N; ++m) acc += dm.d[m * N + j] * col[m];
        sB[offC + j] = acc;
      }
    }
    __syncthreads();

    // ---- phase 5 (mapping A): sum the three parts, scatter
    if (active) {
#pragma unroll
      for (int k = 0; k < N; ++k) {
        const T v = ry[k] + sA[k * S0 + offA] + sB[k * S0 + offA];
        if (LOCAL) {
          y[(e * n + k * P + t) * ncomp + c] = v;
        } else {
          const uint32_t cn = rc[k];
          if (cn != kConnSentinel) {
            T* dst = y + (int64_t)(cn & kConnIdMask) * ncomp + c;
            if (cn & kConnDirichlet) {
              if (cn & kConnSingle) *dst = T(0);
            } else {
              if (cn & kConnSingle)
                *dst = v;
              else
                red_add(dst, v);
              if (want_dot) dot += (double)sU[k * S0 + offA] * (double)v;
            }
          }
        }
      }
    }
    // rotate.  No barrier is needed here: the next iteration's first barrier
    // (after cp.async.wait) is reached by every thread only after its phase-5
    // reads of sA / sB / sU, and those tiles are not written before it (the
    // in-flight cp.async targets the OTHER u tile).
    e = e_n;
    active = active_n;
#pragma unroll
    for (int k = 0; k < N; ++k) rc[k] = nrc[k];
  }
  cp_async_wait_all();
  if (want_dot) {
    dot = block_sum(dot, red);
    if (threadIdx.x == 0) atomicAdd(dot_xy, dot);
  }
}

template <typename T, int N, bool MASS, bool LOCAL, int EPB, int MINB, int KCH, int JU>
int launch3d_v2_cfg(const sfem_op& op, double lambda, double mu, const void* x,
                    void* y, int ncomp, double* dot_xy, cudaStream_t stream) {
  using C = Cfg3DV2<T, N, EPB, MINB, KCH>;
  const int64_t E = op.base.desc.num_elements;
  const int64_t nblocks = (E + C::epb - 1) / C::epb;
  const size_t smem =
      (size_t)C::epb * C::tiles_per_slot * C::tile * sizeof(T);
  auto kernel = apply3d_v2_kernel<T, N, MASS, LOCAL, EPB, MINB, KCH, JU>;
  static int per_sm = 0;
  if (per_sm == 0) {
    if (smem > 48 * 1024)
      SFEM_CUDA_CHECK(cudaFuncSetAttribute(
          kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SFEM_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
        &per_sm, kernel, C::threads, smem));
    if (per_sm < 1) per_sm = 1;
  }
  // persistent CTAs: one wave, every CTA pipelines over its elements
  const int64_t cap = (int64_t)num_sms() * per_sm;
  dim3 grid((unsigned)(nblocks < cap ? nblocks : cap), ncomp);
  DMat3<T, N> dm;
  for (int i = 0; i < N * N; ++i) dm.d[i] = (T)op.base.h_BD[i];
  kernel<<<grid, C::threads, smem, stream>>>(
      dm, op.conn, (const T*)op.geom, (T)lambda, (T)mu, (const T*)x, (T*)y,
      ncomp, E, dot_xy);
  SFEM_LAUNCH_CHECK();
  return SFEM_OK;
}

template <typename T, int N, bool MASS, bool LOCAL>
int launch3d_v2(const sfem_op& op, double lambda, double mu, const void* x,
                void* y, int ncomp, double* dot_xy, cudaStream_t stream) {
  constexpr int E0 = default_epb<N>();
#ifdef SFEM_EXPERIMENTS
  // tuning variants, compiled only for the headline configuration
  if constexpr (N == 8 && !MASS && !LOCAL) {
    switch (op.variant) {
      case 3:
        return launch3d_v2_cfg<T, N, MASS, LOCAL, 4, 1, 2, 8>(
            op, lambda, mu, x, y, ncomp, dot_xy, stream);
      case 4:
        return launch3d_v2_cfg<T, N, MASS, LOCAL, 2, 3, 2, 8>(
            op, lambda, mu, x, y, ncomp, dot_xy, stream);
      case 5:
        return launch3d_v2_cfg<T, N, MASS, LOCAL, 4, 2, 2, 2>(
            op, lambda, mu, x, y, ncomp, dot_xy, stream);
      case 6:
        return launch3d_v2_cfg<T, N, MASS, LOCAL, 4, 2, 2, 1>(
            op, lambda, mu, x, y, ncomp, dot_xy, stream);
      case 7:
        return launch3d_v2_cfg<T, N, MASS, LOCAL, 4, 2, 4, 2>(
            op, lambda, mu, x, y, ncomp, dot_xy, stream);
      case 8:
        return launch3d_v2_cfg<T, N, MASS, LOCAL, 2, 4, 2, 2>(
            op, lambda, mu, x, y, ncomp, dot_xy, stream);
      case 9:
        return launch3d_v2_cfg<T, N, MASS, LOCAL, 4, 2, 2, 4>(
            op, lambda, mu, x, y, ncomp, dot_xy, stream);
      default:
        break;
    }
  }
#endif
  constexpr int MINB = (N <= 8) ? 2 : 1;
  // fp64 contractions fetch every D entry through a uniform register: full
  // unrolling of the output index spills, so fp64 unrolls by 2
  constexpr int JU = sizeof(T) == 8 ? 2 : N;
  return launch3d_v2_cfg<T, N, MASS, LOCAL, E0, MINB, 2, JU>(
      op, lambda, mu, x, y, ncomp, dot_xy, stream);
}

}  // namespace
}  // namespace sfem
