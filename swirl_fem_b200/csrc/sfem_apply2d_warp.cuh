// 2-D collocated-GLL operator apply, warp-autonomous kernel (v3).
//
// Same arithmetic and tile layouts as the two-mapping kernel
// (sfem_apply2d_v2.cuh: N threads per element, row owner <-> column owner,
// even-odd contractions), but every WARP is its own pipeline: a CTA is one
// warp that owns floor(32 / N) consecutive elements per step, its own u /
// work tiles, factor stage and connectivity ring, and synchronises with
// __syncwarp() only.  Why: ncu on the block-synchronous kernel (p = 4 / 8,
// profiles/r02_ncu_orders_pipe_utilisation.txt) shows a latency-bound kernel,
// not a bandwidth-bound one -- 2.2-4.3 short-scoreboard and 1.8-2.1 fixed-
// latency stall cycles per issue, 15-30 % warps active, DRAM at 47-55 % -- with
// 10-20 warps per SM that all stop at the same three block barriers per step.
// Autonomous warps (15-25 per SM, no block barrier anywhere) de-phase and hide
// one another's shared-memory and DRAM latency.
//
// Memory pipeline per warp and step (all cp.async / LDGSTS, one commit group
// per step, waited at the top of the next step):
//   connectivity of step s+2      -> 3-deep ring
//   gather x[idx] of step s+1     -> the other u tile
//   geometric factors of step s+1 -> the stage, as soon as the pointwise phase
//                                    of step s has consumed it (16-byte copies
//                                    when a step's chunk is 16-byte aligned,
//                                    else 8 / 4-byte ones)
#pragma once

#include "sfem_apply2d_v2.cuh"

namespace sfem {
namespace {

template <int BYTES>
__device__ __forceinline__ void cp_async_bytes(void* smem_dst, const void* gsrc) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  if constexpr (BYTES == 16)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gsrc));
  else
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(s), "l"(gsrc),
                 "n"(BYTES));
}

// All 32 lanes: asynchronous copy of `bytes` (rounded up to VEC) from global to
// shared memory; both addresses VEC-aligned.
template <int VEC>
__device__ __forceinline__ void warp_copy_async(void* smem_dst, const void* gsrc,
                                                unsigned bytes) {
  const unsigned n = (bytes + VEC - 1) / VEC;
  for (unsigned i = threadIdx.x; i < n; i += 32)
    cp_async_bytes<VEC>((char*)smem_dst + (size_t)i * VEC,
                        (const char*)gsrc + (size_t)i * VEC);
}

constexpr int vec_for(long chunk_bytes) {
  return chunk_bytes % 16 == 0 ? 16 : (chunk_bytes % 8 == 0 ? 8 : 4);
}

template <typename T, int N, bool MASS>
struct Cfg2DWarp {
  static constexpr int n = N * N;
  static constexpr int epw = 32 / N;  // elements per warp step
  static constexpr int ngeom = MASS ? 4 : 3;
  static constexpr int R = (sizeof(T) == 8 ? kTile2D64 : kTile2D32)[N][0];
  static constexpr int tile = (sizeof(T) == 8 ? kTile2D64 : kTile2D32)[N][1];
  static constexpr bool SWZ =
      (sizeof(T) == 8 ? kTile2D64 : kTile2D32)[N][2] != 0;
  static constexpr int tiles_per_slot = 3;  // u[2], work
  static constexpr long gchunk = (long)epw * ngeom * n * (long)sizeof(T);
  static constexpr long cchunk = (long)epw * n * 4;
  static constexpr int gvec = vec_for(gchunk);
  static constexpr int cvec = vec_for(cchunk);
  static constexpr long round16(long v) { return (v + 15) / 16 * 16; }
  // byte offsets of the three regions (16-byte aligned)
  static constexpr long tiles_bytes =
      round16((long)epw * tiles_per_slot * tile * (long)sizeof(T));
  static constexpr long stage_bytes = round16(gchunk + 16);
  static constexpr long ring_bytes = round16(cchunk + 16);
  static constexpr long smem_bytes = tiles_bytes + stage_bytes + 3 * ring_bytes;
};

template <typename T, int N, bool MASS, bool LOCAL, int MINB>
__global__ void __launch_bounds__(32, MINB)
apply2d_warp_kernel(const __grid_constant__ DOps<T, N> dm,
                    const uint32_t* __restrict__ conn,
                    const T* __restrict__ gf, T lambda, T mu,
                    const T* __restrict__ x, T* __restrict__ y, int ncomp,
                    int64_t E, double* __restrict__ dot_xy) {
  using C = Cfg2DWarp<T, N, MASS>;
  constexpr int n = C::n, epw = C::epw, R = C::R, ngeom = C::ngeom;
  constexpr bool SWZ = C::SWZ;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* tiles = reinterpret_cast<T*>(smem_raw);
  T* sG0 = reinterpret_cast<T*>(smem_raw + C::tiles_bytes);
  unsigned char* sCbase = smem_raw + C::tiles_bytes + C::stage_bytes;
  auto ring_ptr = [&](int r) {
    return reinterpret_cast<uint32_t*>(sCbase + (size_t)r * C::ring_bytes);
  };

  const int lane = threadIdx.x;
  const int slot = lane / N;
  const int t = lane - slot * N;
  const int ts = t;
  const int c = blockIdx.y;
  const bool lane_ok = slot < epw;
  const int s_ = lane_ok ? slot : 0;
  T* sU0 = tiles + s_ * C::tiles_per_slot * C::tile;
  T* sW = sU0 + 2 * C::tile;
  const T* sG = sG0 + s_ * (ngeom * n);
  const bool want_dot = !LOCAL && dot_xy != nullptr;
  double dot = 0.0;

  const int64_t nsteps = (E + epw - 1) / epw;
  int64_t step = blockIdx.x;
  const int64_t stride = gridDim.x;

  auto step_count = [&](int64_t s) {
    const int64_t first = s * epw;
    return (E - first) < epw ? (E - first) : (int64_t)epw;
  };
  auto copy_factors = [&](int64_t s) {
    warp_copy_async<C::gvec>(
        sG0, gf + s * (int64_t)epw * (ngeom * n),
        (unsigned)(step_count(s) * (ngeom * n) * (int64_t)sizeof(T)));
  };
  auto copy_conn = [&](int64_t s, int r) {
    warp_copy_async<C::cvec>(ring_ptr(r), conn + s * (int64_t)epw * n,
                             (unsigned)(step_count(s) * n * 4));
  };
  auto issue_gather = [&](int64_t s, T* dst, int r) {
    if (!lane_ok) return;
    const int64_t e = s * epw + slot;
    const bool act = s < nsteps && e < E;
    const uint32_t* cn = ring_ptr(r) + slot * n + t * N;
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

  // ---- prologue
  int ring = 0;  // ring entry holding the CURRENT step's connectivity
  if (step < nsteps) {
    if (!LOCAL) {
      copy_conn(step, 0);
      cp_async_commit();
      cp_async_wait_all();
      __syncwarp();
      if (step + stride < nsteps) copy_conn(step + stride, 1);
    }
    copy_factors(step);
    issue_gather(step, sU0, 0);
  }
  cp_async_commit();

  int buf = 0;
  for (; step < nsteps; step += stride, buf ^= 1) {
    T* sU = sU0 + buf * C::tile;
    T* sUn = sU0 + (buf ^ 1) * C::tile;
    const int64_t e = step * epw + slot;
    const bool active = lane_ok && e < E;
    const int64_t step_n = step + stride;
    const int64_t step_nn = step_n + stride;
    const int ring_n = (ring + 1) % 3, ring_nn = (ring + 2) % 3;

    // u tile, factors of this step and connectivity of the next have landed
    cp_async_wait_all();
    __syncwarp();
    // connectivity two steps ahead (its ring entry was last read in phase 4 of
    // the previous step)
    if (!LOCAL && step_nn < nsteps) copy_conn(step_nn, ring_nn);

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
    __syncwarp();

    // ---- gather of the next step into the other u tile
    issue_gather(step_n, sUn, ring_n);

    // ---- phase 2 (A): geometric factors, transposed row derivative
    T yrow[N];
#pragma unroll
    for (int j = 0; j < N; ++j) yrow[j] = T(0);
    if (lane_ok) {
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
    __syncwarp();
    // the stage is free: factors of the next step
    if (step_n < nsteps) copy_factors(step_n);
    cp_async_commit();

    // ---- phase 3 (B): transposed column derivative, in place
    if (lane_ok) {
      T col[N], out[N];
#pragma unroll
      for (int m = 0; m < N; ++m) col[m] = sW[m * R + SFEM_SW(ts, m)];
      eo_apply<T, N>(dm.bwd, col, out);
#pragma unroll
      for (int m = 0; m < N; ++m) sW[m * R + SFEM_SW(ts, m)] = out[m];
    }
    __syncwarp();

    // ---- phase 4 (A): sum, scatter
    if (active) {
      const uint32_t* cn = ring_ptr(ring) + slot * n + t * N;
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
    // phase-4 reads of sW / sU / the ring precede the next step's writes to
    // them (next phase 1 writes sW after the syncwarp at the top)
    ring = ring_n;
  }
  cp_async_wait_all();
  if (want_dot) {
    dot = warp_sum(dot);
    if (lane == 0) atomicAdd(dot_xy, dot);
  }
}

template <typename T, int N, bool MASS, bool LOCAL>
int launch2d_warp(const sfem_op& op, double lambda, double mu, const void* x,
                  void* y, int ncomp, double* dot_xy, cudaStream_t stream) {
  using C = Cfg2DWarp<T, N, MASS>;
  const int64_t E = op.base.desc.num_elements;
  const int64_t nsteps = (E + C::epw - 1) / C::epw;
  constexpr size_t smem = (size_t)C::smem_bytes;
  constexpr int by_smem = (int)((220L * 1024) / C::smem_bytes);
  // register estimate fitted to ptxas (fp64: 68 @ N=5, 106 @ N=9)
  constexpr int est_regs = 24 + (sizeof(T) == 8 ? 10 : 6) * N;
  constexpr int by_regs = 65536 / (32 * est_regs);
  constexpr int m0 = by_smem < by_regs ? by_smem : by_regs;
  constexpr int MINB = m0 < 1 ? 1 : (m0 > 32 ? 32 : m0);
  auto kernel = apply2d_warp_kernel<T, N, MASS, LOCAL, MINB>;
  int dev = 0;
  SFEM_CUDA_CHECK(cudaGetDevice(&dev));
  static int per_sm_dev[64] = {};
  int& per_sm = per_sm_dev[dev & 63];
  if (per_sm == 0) {
    if (smem > 48 * 1024)
      SFEM_CUDA_CHECK(cudaFuncSetAttribute(
          kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SFEM_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
        &per_sm, kernel, 32, smem));
    if (per_sm < 1) per_sm = 1;
  }
  const int64_t cap = (int64_t)num_sms() * per_sm;
  dim3 grid((unsigned)(nsteps < cap ? nsteps : cap), ncomp);
  DOps<T, N> dm;
  fill_even_odd<T, N>(op.base.h_BD, false, &dm.fwd);
  fill_even_odd<T, N>(op.base.h_BD, true, &dm.bwd);
  kernel<<<grid, 32, smem, stream>>>(dm, op.conn, (const T*)op.geom, (T)lambda,
                                     (T)mu, (const T*)x, (T*)y, ncomp, E,
                                     dot_xy);
  SFEM_LAUNCH_CHECK();
  return SFEM_OK;
}

}  // namespace
}  // namespace sfem
