// 2-D collocated-GLL operator apply, "two-mapping" kernel (v2).
//
// N threads per element; thread t is, depending on the phase,
//   mapping A: owner of ROW    t  (u[t, :],  contiguous axis a1)
//   mapping B: owner of COLUMN t  (u[:, t])
// so both 1-D contractions run in registers (even-odd decomposition of the GLL
// differentiation matrix, matrix entries from uniform registers / the constant
// bank).  Rows/columns are exchanged through XOR-swizzled shared tiles; three
// block barriers per CTA step.
//
// A CTA step handles EPB consecutive elements, so the step's geometric factors
// and connectivity are each ONE contiguous chunk:
//   * factors:      one bulk async copy (TMA 1-D) per step into a shared stage,
//                   refilled as soon as the pointwise phase has consumed it;
//   * connectivity: one bulk copy per step into a 3-deep ring, issued two steps
//                   ahead (it feeds the gather of the NEXT step);
//   * gather x[idx] of the next step: cp.async (LDGSTS) into the other u tile.
// Persistent CTAs (one wave).  Buffers are padded by the allocator so the last
// step's copies may round their size up to 16 bytes.
#pragma once

#include "sfem_apply3d_v2.cuh"

namespace sfem {
namespace {

template <typename T, int N, int EPB>
struct Cfg2DV2 {
  static constexpr int n = N * N;
  static constexpr int epb = EPB;
  static constexpr int threads = ((EPB * N + 31) / 32) * 32;
  // row pitch / tile size / XOR swizzle from an offline bank-conflict search
  // over the row-owner and column-owner access patterns
  static constexpr int R = (sizeof(T) == 8 ? kTile2D64 : kTile2D32)[N][0];
  static constexpr int tile = (sizeof(T) == 8 ? kTile2D64 : kTile2D32)[N][1];
  static constexpr bool SWZ =
      (sizeof(T) == 8 ? kTile2D64 : kTile2D32)[N][2] != 0;
  static constexpr int tiles_per_slot = 3;   // u[2], work
  // offset (in T) of the factor stage: 16-byte aligned
  static constexpr int stage_off =
      ((EPB * tiles_per_slot * tile * (int)sizeof(T) + 15) / 16) * 16 /
      (int)sizeof(T);
};

#define SFEM_SW(a, b) (SWZ ? ((a) ^ (b)) : (a))

template <typename T, int N, bool MASS, bool LOCAL, int EPB, int MINB>
__global__ void __launch_bounds__((Cfg2DV2<T, N, EPB>::threads), MINB)
apply2d_v2_kernel(const __grid_constant__ DOps<T, N> dm,
                  const uint32_t* __restrict__ conn,
                  const T* __restrict__ gf, T lambda, T mu,
                  const T* __restrict__ x, T* __restrict__ y, int ncomp,
                  int64_t E, double* __restrict__ dot_xy) {
  using C = Cfg2DV2<T, N, EPB>;
  constexpr int n = C::n, epb = C::epb, R = C::R;
  constexpr bool SWZ = C::SWZ;
  constexpr int ngeom = MASS ? 4 : 3;
  constexpr unsigned gbytes = (unsigned)(ngeom * n * sizeof(T));
  constexpr unsigned cbytes = (unsigned)(n * sizeof(uint32_t));
  static_assert((EPB * gbytes) % 16 == 0 && (EPB * cbytes) % 16 == 0,
                "EPB must make a CTA step's chunks multiples of 16 bytes");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double red[32];
  __shared__ __align__(8) uint64_t gbar;
  __shared__ __align__(8) uint64_t cbar[3];

  // shared layout: [tiles (epb*3)] [factor stage (epb*ngeom*n)] [conn ring 3x]
  T* tiles = reinterpret_cast<T*>(smem_raw);
  T* sG0 = tiles + C::stage_off;
  uint32_t* sC0 = reinterpret_cast<uint32_t*>(sG0 + epb * ngeom * n);

  const int slot = threadIdx.x / N;
  const int t = threadIdx.x - slot * N;
  const int c = blockIdx.y;
  const bool lane_ok = slot < epb;
  const int s_ = lane_ok ? slot : 0;
  T* sU0 = tiles + s_ * C::tiles_per_slot * C::tile;
  T* sW = sU0 + 2 * C::tile;
  const T* sG = sG0 + s_ * (ngeom * n);
  const int ts = t;
  // swizzled tile index: (i, j) -> i * R + (j ^ i)
  const bool want_dot = !LOCAL && dot_xy != nullptr;
  double dot = 0.0;

  const int64_t nsteps = (E + epb - 1) / epb;
  int64_t step = blockIdx.x;

  auto step_count = [&](int64_t s) {
    const int64_t first = s * epb;
    return (E - first) < epb ? (E - first) : (int64_t)epb;
  };
  auto copy_factors = [&](int64_t s) {
    const unsigned bytes = ((unsigned)step_count(s) * gbytes + 15u) & ~15u;
    mbar_expect_tx(&gbar, bytes);
    bulk_copy_g2s(sG0, gf + s * (int64_t)epb * (ngeom * n), bytes, &gbar);
  };
  auto copy_conn = [&](int64_t s, int ring) {
    const unsigned bytes = ((unsigned)step_count(s) * cbytes + 15u) & ~15u;
    mbar_expect_tx(&cbar[ring], bytes);
    bulk_copy_g2s(sC0 + ring * (epb * n), conn + s * (int64_t)epb * n, bytes,
                  &cbar[ring]);
  };
  // gather of step `s` (slot's element) into u tile `dst`, connectivity from
  // ring entry `ring` (already landed)
  auto issue_gather = [&](int64_t s, T* dst, int ring) {
    const int64_t e = s * epb + slot;
    const bool act = lane_ok && s < nsteps && e < E;
    if (!lane_ok) return;
    const uint32_t* cn = sC0 + ring * (epb * n) + slot * n + t * N;
#pragma unroll
    for (int j = 0; j < N; ++j) {
      T* d = dst + t * R + SFEM_SW(j, ts);
      if (act) {
        if (LOCAL) {
          cp_async_elem(d, x + ((e * n + t * N + j) * (int64_t)ncomp + c));
        } else {
          const uint32_t w = cn[j];
          if (w != kConnSentinel)
            cp_async_elem(d, x + (int64_t)(w & kConnIdMask) * ncomp + c);
          else
            *d = T(0);
        }
      } else {
        *d = T(0);
      }
    }
  };

  if (threadIdx.x == 0) {
    mbar_init(&gbar, 1);
    mbar_init(&cbar[0], 1);
    mbar_init(&cbar[1], 1);
    mbar_init(&cbar[2], 1);
  }
  mbar_fence_init();
  __syncthreads();
  unsigned gphase = 0;
  unsigned cphase = 0;  // bit r = parity of ring entry r
  int ring = 0;         // ring entry holding the CURRENT step's connectivity
  if (step < nsteps) {
    if (threadIdx.x == 0) {
      copy_factors(step);
      if (!LOCAL) {
        copy_conn(step, 0);
        if (step + gridDim.x < nsteps) copy_conn(step + gridDim.x, 1);
      }
    }
    if (!LOCAL) {
      mbar_wait(&cbar[0], 0);
      cphase ^= 1u;
    }
    issue_gather(step, sU0, 0);
  }
  cp_async_commit();

  int buf = 0;
  for (; step < nsteps; step += gridDim.x, buf ^= 1) {
    T* sU = sU0 + buf * C::tile;
    T* sUn = sU0 + (buf ^ 1) * C::tile;
    const int64_t e = step * epb + slot;
    const bool active = lane_ok && e < E;
    const int64_t step_n = step + gridDim.x;
    const int64_t step_nn = step_n + gridDim.x;
    const int ring_n = (ring + 1) % 3, ring_nn = (ring + 2) % 3;

    cp_async_wait_all();
    __syncthreads();
    // connectivity two steps ahead: its ring entry was last read in phase 4 of
    // the previous step, which every thread has left
    if (!LOCAL && threadIdx.x == 0 && step_nn < nsteps)
      copy_conn(step_nn, ring_nn);

    // ---- phase 1: row derivative (A) kept in registers, column derivative (B)
    T row[N], ds[N];
    if (lane_ok) {
#pragma unroll
      for (int m = 0; m < N; ++m) row[m] = sU[t * R + SFEM_SW(m, ts)];
      eo_apply<T, N>(dm.fwd, row, ds);  // d/d(a1) along the row
      T col[N], dr[N];
#pragma unroll
      for (int m = 0; m < N; ++m) col[m] = sU[m * R + SFEM_SW(ts, m)];
      eo_apply<T, N>(dm.fwd, col, dr);  // d/d(a0) along the column
#pragma unroll
      for (int m = 0; m < N; ++m) sW[m * R + SFEM_SW(ts, m)] = dr[m];
    }
    __syncthreads();

    // ---- gather of the next step into the other u tile
    if (!LOCAL && step_n < nsteps) {
      mbar_wait(&cbar[ring_n], (cphase >> ring_n) & 1u);
      cphase ^= 1u << ring_n;
    }
    issue_gather(step_n, sUn, ring_n);
    cp_async_commit();

    // ---- phase 2 (A): geometric factors, transposed row derivative
    T yrow[N];
#pragma unroll
    for (int j = 0; j < N; ++j) yrow[j] = T(0);
    if (lane_ok) {
      mbar_wait(&gbar, gphase);
      T ws[N];
      const T* g = sG + t * N;
#pragma unroll
      for (int j = 0; j < N; ++j) {
        const T dr = sW[t * R + SFEM_SW(j, ts)];
        const T g00 = active ? g[j] : T(0);
        const T g01 = active ? g[n + j] : T(0);
        const T g11 = active ? g[2 * n + j] : T(0);
        sW[t * R + SFEM_SW(j, ts)] = mu * (g00 * dr + g01 * ds[j]);
        ws[j] = mu * (g01 * dr + g11 * ds[j]);
        if (MASS) row[j] *= lambda * (active ? g[3 * n + j] : T(0));
      }
      eo_apply<T, N>(dm.bwd, ws, yrow);
      if (MASS) {
#pragma unroll
        for (int j = 0; j < N; ++j) yrow[j] += row[j];
      }
    }
    __syncthreads();
    gphase ^= 1u;
    if (threadIdx.x == 0 && step_n < nsteps) copy_factors(step_n);

    // ---- phase 3 (B): transposed column derivative, in place
    if (lane_ok) {
      T col[N], out[N];
#pragma unroll
      for (int m = 0; m < N; ++m) col[m] = sW[m * R + SFEM_SW(ts, m)];
      eo_apply<T, N>(dm.bwd, col, out);
#pragma unroll
      for (int m = 0; m < N; ++m) sW[m * R + SFEM_SW(ts, m)] = out[m];
    }
    __syncthreads();

    // ---- phase 4 (A): sum, scatter
    if (active) {
      const uint32_t* cn = sC0 + ring * (epb * n) + slot * n + t * N;
#pragma unroll
      for (int j = 0; j < N; ++j) {
        const T v = yrow[j] + sW[t * R + SFEM_SW(j, ts)];
        if (LOCAL) {
          y[(e * n + t * N + j) * (int64_t)ncomp + c] = v;
        } else {
          const uint32_t w = cn[j];
          if (w != kConnSentinel) {
            T* dst = y + (int64_t)(w & kConnIdMask) * ncomp + c;
            if (w & kConnDirichlet) {
              if (w & kConnSingle) *dst = T(0);
            } else {
              if (w & kConnSingle)
                *dst = v;
              else
                red_add(dst, v);
              if (want_dot) dot += (double)sU[t * R + SFEM_SW(j, ts)] * (double)v;
            }
          }
        }
      }
    }
    ring = ring_n;
  }
  cp_async_wait_all();
  if (want_dot) {
    dot = block_sum(dot, red);
    if (threadIdx.x == 0) atomicAdd(dot_xy, dot);
  }
}

template <typename T, int N, bool MASS, bool LOCAL, int EPB, int MINB>
int launch2d_v2_cfg(const sfem_op& op, double lambda, double mu, const void* x,
                    void* y, int ncomp, double* dot_xy, cudaStream_t stream) {
  using C = Cfg2DV2<T, N, EPB>;
  constexpr int ngeom = MASS ? 4 : 3;
  const int64_t E = op.base.desc.num_elements;
  const int64_t nsteps = (E + EPB - 1) / EPB;
  const size_t smem =
      ((size_t)C::stage_off + (size_t)EPB * ngeom * C::n) * sizeof(T) +
      (size_t)3 * EPB * C::n * sizeof(uint32_t);
  auto kernel = apply2d_v2_kernel<T, N, MASS, LOCAL, EPB, MINB>;
  static int per_sm_dev[64] = {};
  int& per_sm = per_device_slot(per_sm_dev);
  if (per_sm == 0) {
    if (smem > 48 * 1024)
      SFEM_CUDA_CHECK(cudaFuncSetAttribute(
          kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SFEM_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
        &per_sm, kernel, C::threads, smem));
    if (per_sm < 1) per_sm = 1;
  }
  const int64_t cap = (int64_t)num_sms() * per_sm;
  dim3 grid((unsigned)(nsteps < cap ? nsteps : cap), ncomp);
  DOps<T, N> dm;
  fill_even_odd<T, N>(op.base.h_BD, false, &dm.fwd);
  fill_even_odd<T, N>(op.base.h_BD, true, &dm.bwd);
  kernel<<<grid, C::threads, smem, stream>>>(
      dm, op.conn, (const T*)op.geom, (T)lambda, (T)mu, (const T*)x, (T*)y,
      ncomp, E, dot_xy);
  SFEM_LAUNCH_CHECK();
  return SFEM_OK;
}

template <typename T, int N, bool MASS, bool LOCAL>
int launch2d_v2(const sfem_op& op, double lambda, double mu, const void* x,
                void* y, int ncomp, double* dot_xy, cudaStream_t stream) {
  // EPB: a multiple of 4 (16-byte chunks for any N), ~128 threads per CTA
  constexpr int e0 = ((128 + N - 1) / N + 3) / 4 * 4;
  constexpr int EPB = e0 < 4 ? 4 : e0;
  using C = Cfg2DV2<T, N, EPB>;
  constexpr long smem =
      ((long)C::stage_off + (long)EPB * (MASS ? 4 : 3) * C::n) *
          (long)sizeof(T) +
      3L * EPB * C::n * 4;
  constexpr int by_smem = (int)((220L * 1024) / smem);
  constexpr int est_regs_raw = 48 + (sizeof(T) == 8 ? 14 : 7) * N;
  constexpr int est_regs = est_regs_raw > 255 ? 255 : est_regs_raw;
  constexpr int by_regs = 65536 / (C::threads * est_regs);
  constexpr int m0 = by_smem < by_regs ? by_smem : by_regs;
  constexpr int MINB = m0 < 1 ? 1 : (m0 > 8 ? 8 : m0);
  return launch2d_v2_cfg<T, N, MASS, LOCAL, EPB, MINB>(op, lambda, mu, x, y,
                                                        ncomp, dot_xy, stream);
}

}  // namespace
}  // namespace sfem
