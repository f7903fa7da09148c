// 3-D collocated-GLL operator apply, "three-mapping" kernel (v2).
//
// N^2 threads per element.  The same thread index (p, q) = (t / N, t % N) is
// interpreted under three mappings, one per tensor axis, so that every 1-D
// contraction runs entirely in registers against compile-time-indexed matrix
// entries (uniform-register / constant-bank operands; no per-thread registers
// or shared-memory traffic for the derivative matrix):
//   mapping A: thread owns the a0-column  u[:, p, q]
//   mapping B: thread owns the a1-column  u[p, :, q]
//   mapping C: thread owns the a2-column  u[p, q, :]
// Columns are exchanged through XOR-swizzled shared-memory tiles (u
// double-buffered + two work tiles) that are bank-conflict free under all
// three mappings; per element there are 4 block barriers (the slab-sweep
// kernel v1 needs 2N).
//
// Contractions use the even-odd decomposition of the GLL differentiation
// matrix (D is centro-antisymmetric because the nodes are symmetric):
//   E_j = sum_m Ae[j][m] (c[m] + c[N-1-m]),  O_j = sum_m Ao[j][m] (c[m] - c[N-1-m])
//   out[j] = E_j + O_j,  out[N-1-j] = O_j - E_j
// which halves the multiply count and the number of matrix-entry fetches (on
// sm_100a every fp64 entry costs one LDCU: DFMA takes no constant operand).
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
#include "sfem_tile_layouts.cuh"

namespace sfem {
namespace {

// Even-odd factors of D (apply) and D^T (transposed apply).
template <typename T, int N>
struct EvenOdd {
  static constexpr int H = N / 2;
  static constexpr int HS = H > 0 ? H : 1;
  T ae[HS * HS];   // (D[j][m] + D[j][N-1-m]) / 2
  T ao[HS * HS];   // (D[j][m] - D[j][N-1-m]) / 2
  T mid_col[HS];   // odd N: D[j][H]
  T mid_row[HS];   // odd N: D[H][m]
};

template <typename T, int N>
struct DOps {
  EvenOdd<T, N> fwd;  // out[j] = sum_m D[j][m] in[m]
  EvenOdd<T, N> bwd;  // out[j] = sum_m D[m][j] in[m]
};

template <typename T, int N>
void fill_even_odd(const double* D, bool transpose, EvenOdd<T, N>* eo) {
  constexpr int H = N / 2;
  auto at = [&](int j, int m) { return transpose ? D[m * N + j] : D[j * N + m]; };
  for (int j = 0; j < H; ++j) {
    for (int m = 0; m < H; ++m) {
      eo->ae[j * H + m] = (T)(0.5 * (at(j, m) + at(j, N - 1 - m)));
      eo->ao[j * H + m] = (T)(0.5 * (at(j, m) - at(j, N - 1 - m)));
    }
    eo->mid_col[j] = (N & 1) ? (T)at(j, H) : T(0);
    eo->mid_row[j] = (N & 1) ? (T)at(H, j) : T(0);
  }
}

// out = M in with M given by its even-odd factors.  `in` and `out` may alias
// only if the caller copies; all indices are compile-time after unrolling.
template <typename T, int N>
__device__ __forceinline__ void eo_apply(const EvenOdd<T, N>& eo,
                                         const T (&in)[N], T (&out)[N]) {
  constexpr int H = N / 2;
  if constexpr (H == 0) {
    out[0] = T(0);
    return;
  }
  T ce[H > 0 ? H : 1], co[H > 0 ? H : 1], E[H > 0 ? H : 1], O[H > 0 ? H : 1];
#pragma unroll
  for (int m = 0; m < H; ++m) {
    ce[m] = in[m] + in[N - 1 - m];
    co[m] = in[m] - in[N - 1 - m];
  }
#pragma unroll
  for (int j = 0; j < H; ++j) {
    E[j] = (N & 1) ? eo.mid_col[j] * in[H] : T(0);
    O[j] = T(0);
  }
#pragma unroll
  for (int m = 0; m < H; ++m) {
#pragma unroll
    for (int j = 0; j < H; ++j) {
      E[j] += eo.ae[j * H + m] * ce[m];
      O[j] += eo.ao[j * H + m] * co[m];
    }
  }
#pragma unroll
  for (int j = 0; j < H; ++j) {
    out[j] = E[j] + O[j];
    out[N - 1 - j] = O[j] - E[j];
  }
  if constexpr (N & 1) {
    T acc = T(0);
#pragma unroll
    for (int m = 0; m < H; ++m) acc += eo.mid_row[m] * co[m];
    out[H] = acc;
  }
}

constexpr int pow2_at_least(int n) {
  int r = 1;
  while (r < n) r *= 2;
  return r;
}

// EPB: elements per CTA; MINB: min CTAs per SM (register cap); KCH: slabs of
// geometric factors loaded per batch in phase 3 (bounds registers in flight).
template <typename T, int N, int EPB, int MINB, int KCH>
struct Cfg3DV2 {
  static constexpr int P = N * N;  // threads per element
  static constexpr int n = N * N * N;
  static constexpr int epb = EPB;
  static constexpr int threads = ((EPB * P + 31) / 32) * 32;
  // tile index: idx(a0,a1,a2) = a0*S0 + a1*R + (SWZ ? a2 ^ a1 : a2).  The
  // pitches (and whether the XOR swizzle is used: power-of-two N only) come
  // from an offline search that minimises shared-memory wavefronts over the
  // three access patterns for the actual lane -> (p, q) assignment
  // (tools/tile_layout_search.py -> sfem_tile_layouts.cuh).
  static constexpr int R = (sizeof(T) == 8 ? kTileLayout64 : kTileLayout32)[N][0];
  static constexpr int S0 = (sizeof(T) == 8 ? kTileLayout64 : kTileLayout32)[N][1];
  static constexpr bool SWZ =
      (sizeof(T) == 8 ? kTileLayout64 : kTileLayout32)[N][2] != 0;
  static constexpr int slot_pad =
      (sizeof(T) == 8 ? kTileLayout64 : kTileLayout32)[N][3];
  static constexpr int tile = N * S0;
  static constexpr int tiles_per_slot = 4;  // u[2], A, B
  static constexpr int slot_stride = tiles_per_slot * tile + slot_pad;
  // offset (in T) of the factor stage: 16-byte aligned
  static constexpr int stage_off(int epb_) {
    return ((epb_ * slot_stride * (int)sizeof(T) + 15) / 16) * 16 / (int)sizeof(T);
  }
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

// ---- bulk async copy (TMA 1-D, UBLKCP) + mbarrier, used to stage one
//      element's geometric factors (a single contiguous chunk) in shared memory
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(s),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gsrc,
                                              unsigned bytes, uint64_t* bar) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1], %2, [%3];" ::"r"(d),
      "l"(gsrc), "r"(bytes), "r"(b)
      : "memory");
}

// HALO: the CTA steps [0, hd.n_if_blocks) hold the elements that touch another
// rank's block.  A CTA that has finished its share of them signals a global
// counter; once every CTA has, the shared dofs of y are final on this rank and
// the CTAs push them (slice by slice, between two element steps) straight into
// the peers' receive buffers over NVLink while the interior elements are still
// being computed.  After its push a CTA polls the peers' flags between element
// steps; once all peers' values have arrived the CTAs also run the canonical
// sum of the shared dofs (no interior element touches them), so in the usual
// lock-step case the whole exchange is hidden and the wait kernel that follows
// finds nothing left to do.  Nothing ever spins: a CTA that exits before a
// condition holds leaves its slices to the CTAs that are still running (the
// last signaller always is) or, for the sum, to the wait kernel.
// Same copy with an L2 evict-first policy: the factors are read exactly once,
// so they should not push x / y lines (which ARE re-used by neighbouring
// elements a few steps later) out of L2.  [experimental variant, see launch3d_v2]
__device__ __forceinline__ void bulk_copy_g2s_evict_first(void* smem_dst,
                                                          const void* gsrc,
                                                          unsigned bytes,
                                                          uint64_t* bar) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
  uint64_t policy;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;"
               : "=l"(policy));
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes"
      ".L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(d),
      "l"(gsrc), "r"(bytes), "r"(b), "l"(policy)
      : "memory");
}

// CONN2 / EVICT: pipeline options chosen per (precision, N) by Tune3D below.
//   CONN2: the connectivity words are fetched TWO element steps ahead (ncu on
//          the one-step kernel: 21 % of all stall samples sat on the first use
//          of the next element's connectivity, a DRAM load issued only one
//          barrier earlier);
//   EVICT: factors staged with the evict-first copy above.
//   LAZY:  y's shared-dof prefix is zeroed by the companion lazy_zero_kernel
//          while this kernel runs (LazyDev in sfem_common.cuh): a step scatters
//          only once its chunk's counter is complete.
template <typename T, int N, bool MASS, bool LOCAL, int EPB, int MINB, int KCH,
          bool HALO = false, bool CONN2 = false, bool EVICT = false,
          bool LAZY = false>
__global__ void __launch_bounds__((Cfg3DV2<T, N, EPB, MINB, KCH>::threads),
                                  MINB)
apply3d_v2_kernel(const __grid_constant__ DOps<T, N> dm,
                  const uint32_t* __restrict__ conn,
                  const T* __restrict__ gf, T lambda, T mu,
                  const T* __restrict__ x, T* __restrict__ y, int ncomp,
                  int64_t E, double* __restrict__ dot_xy,
                  const __grid_constant__ HaloDev hd,
                  const __grid_constant__ LazyDev lz) {
  static_assert(!LAZY || (!HALO && !LOCAL), "lazy zero fill: global form only");
  using C = Cfg3DV2<T, N, EPB, MINB, KCH>;
  constexpr int P = C::P, n = C::n, epb = C::epb;
  constexpr int S0 = C::S0, R = C::R;
  constexpr bool SWZ = C::SWZ;
  constexpr int ngeom = MASS ? 7 : 6;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double red[32];
  T* smem = reinterpret_cast<T*>(smem_raw);

  const int slot = threadIdx.x / P;
  const int t = threadIdx.x - slot * P;
  const int p = t / N, q = t - p * N;
  const int c = blockIdx.y;
  const bool lane_ok = slot < epb;
  T* sU0 = smem + (lane_ok ? slot : 0) * C::slot_stride;
  T* sA = sU0 + 2 * C::tile;
  T* sB = sA + C::tile;
  // swizzled offsets (see Cfg3DV2)
  const int offA = p * R + (SWZ ? (q ^ p) : q);  // + k * S0          mapping A
  const int baseB = p * S0;                      // + m * R + sw(q, m) mapping B
  const int baseC = p * S0 + q * R;              // + sw(m, q)         mapping C
  const int qs = q;
  const bool want_dot = !LOCAL && dot_xy != nullptr;
  double dot = 0.0;

  const int64_t nblocks = (E + epb - 1) / epb;
  int64_t blk = blockIdx.x;
  int64_t e = blk * epb + slot;
  bool active = lane_ok && blk < nblocks && e < E;
  uint32_t rc[N];

  // KCH == 0: the element's geometric factors are staged in shared memory by
  // one bulk async copy (TMA 1-D) per element, completion on an mbarrier
  constexpr bool STAGE = KCH == 0;
  constexpr unsigned gbytes = (unsigned)(ngeom * n * sizeof(T));
  // a CTA step's chunk (EPB elements) must start 16-byte aligned; the last
  // step rounds its size up to 16 bytes (the allocator pads the buffer)
  static_assert(!STAGE || (EPB * gbytes) % 16 == 0,
                "bulk copy needs 16 B aligned CTA chunks");
  // (the CTA's `epb` elements are consecutive, so it is ONE copy per CTA step)
  __shared__ __align__(8) uint64_t gbar;
  T* sG0 = smem + C::stage_off(epb);
  T* sG = sG0 + (lane_ok ? slot : 0) * (ngeom * n);
  unsigned gphase = 0;
  auto stage_copy = [&](int64_t blk_id) {
    // elected thread: refill the stage with the factors of CTA step `blk_id`
    const int64_t first = blk_id * epb;
    const int64_t count = (E - first) < epb ? (E - first) : epb;
    const unsigned bytes = ((unsigned)count * gbytes + 15u) & ~15u;
    mbar_expect_tx(&gbar, bytes);
    if constexpr (EVICT)
      bulk_copy_g2s_evict_first(sG0, gf + first * (int64_t)(ngeom * n), bytes,
                                &gbar);
    else
      bulk_copy_g2s(sG0, gf + first * (int64_t)(ngeom * n), bytes, &gbar);
  };
  // LAZY: the CTA steps after the first wave are CLAIMED from a counter (three
  // steps ahead, so that connectivity and factors can be prefetched) instead
  // of strided: the CTAs then work on one tight window of consecutive steps
  // -- what the companion's pacing and the L2 residency of its zeros need.
  // s_lz[0]: "this step's chunk is zeroed", s_lz[1]: the ticket just claimed.
  __shared__ unsigned s_lz[2];
  if (LAZY && threadIdx.x == 0) s_lz[1] = atomicAdd(&lz.counters[0], 2u);
  if (STAGE) {
    if (threadIdx.x == 0) mbar_init(&gbar, 1);
    mbar_fence_init();
    __syncthreads();
    if (threadIdx.x == 0 && blk < nblocks) stage_copy(blk);
  } else if (LAZY) {
    __syncthreads();
  }
  // the steps after `blk`
  int64_t blk1 = LAZY ? (int64_t)gridDim.x + s_lz[1] : blk + gridDim.x;
  int64_t blk2 = LAZY ? blk1 + 1 : blk1 + gridDim.x;

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

  // HALO state machine (block-uniform), advanced between element steps
  enum : int {
    kHIface = 0,   // interface steps of this CTA pending
    kHSignal = 1,  // signalled; waiting for the other CTAs' interface steps
    kHIssued = 2,  // one push slice issued (loads + remote stores in flight)
    kHFenced = 3,  // ... fenced and counted; the ticket is in flight
    kHPoll = 4,    // pushed; polling the peers' flags
    kHDone = 5
  };
  __shared__ unsigned s_halo[3];  // [0] ready to push, [1] slice, [2] peers late
  int hstate = kHIface;
  uint64_t hpend = 0;    // this thread's poll in flight (kHPoll)
  unsigned hiter = 0;    // element steps spent in kHPoll
  unsigned hticket = 0;  // thread 0: this CTA's slice was the hticket-th done
  // the poll must be ONE independent load: its address is computed once
  const void* hpoll = &hd.counters[2];
  if (HALO && threadIdx.x < hd.num_peers)
    hpoll = hd.flags + hd.peer_ranks[threadIdx.x];
  const bool stamp = HALO && blockIdx.x == 0 && threadIdx.x == 0;
  if (stamp) halo_stamp(hd, 0);
  // programmatic dependent launch: whatever follows may be scheduled as this
  // grid drains; this grid itself may have started before the zero fill of y
  // finished (see sfem_op::pdl) and waits for it before its first write
  pdl_launch_dependents();
  bool first_step = true;

  // CONN2: connectivity of the NEXT step, carried across iterations
  uint32_t nrc_c[CONN2 ? N : 1];
  if constexpr (CONN2 && !LOCAL) {
    const int64_t e1 = blk1 * epb + slot;
    const bool a1 = lane_ok && blk1 < nblocks && e1 < E;
#pragma unroll
    for (int k = 0; k < N; ++k)
      nrc_c[k] = a1 ? ld_stream(conn + e1 * n + k * P + t) : kConnSentinel;
  }

  // LAZY (thread 0): chunk of this step, its count when last looked at, the
  // number of its pieces, the ticket in flight
  unsigned lz_c = 0, lz_seen = 0, lz_need = 0, lz_ticket = 0;

  int buf = 0;
  for (; blk < nblocks; buf ^= 1) {
    T* sU = sU0 + buf * C::tile;
    T* sUn = sU0 + (buf ^ 1) * C::tile;
    if (LAZY && threadIdx.x == 0) {
      lz_ticket = atomicAdd(&lz.counters[0], 1u);  // the step after blk2
      lz_c = (unsigned)blk / lz.chunk_steps;
      lz_seen = ld_relaxed_gpu(&lz.counters[2 + lz_c]);
      lz_need = (unsigned)(__ldg(lz.chunk_ptr + lz_c + 1) -
                           __ldg(lz.chunk_ptr + lz_c));
    }
    // ---- pipeline: next element's factors -> L2, connectivity -> registers
    const int64_t blk_n = blk1;
    const int64_t e_n = blk_n * epb + slot;
    const bool active_n = lane_ok && blk_n < nblocks && e_n < E;
    uint32_t nrc[N];
#pragma unroll
    for (int k = 0; k < N; ++k) nrc[k] = kConnSentinel;
    if constexpr (CONN2 && !LOCAL) {
      // this step's "next" words arrived a whole step ago; fetch the ones of
      // the step after it
      const int64_t blk_nn = blk2;
      const int64_t e_nn = blk_nn * epb + slot;
      const bool active_nn = lane_ok && blk_nn < nblocks && e_nn < E;
#pragma unroll
      for (int k = 0; k < N; ++k) {
        nrc[k] = nrc_c[k];
        nrc_c[k] = active_nn ? ld_stream(conn + e_nn * n + k * P + t)
                             : kConnSentinel;
      }
    }
    if (active_n) {
      if (!STAGE) {
        const char* g =
            reinterpret_cast<const char*>(gf + e_n * (int64_t)(ngeom * n));
        constexpr int lines = (ngeom * n * (int)sizeof(T) + 127) / 128;
#pragma unroll
        for (int l = 0; l < (lines + P - 1) / P; ++l)
          if (l * P + t < lines) prefetch_l2(g + (int64_t)(l * P + t) * 128);
      }
      if (!LOCAL && !CONN2) {
#pragma unroll
        for (int k = 0; k < N; ++k)
          nrc[k] = ld_stream(conn + e_n * n + k * P + t);
      }
    }

    if (HALO) {
      if (hstate == kHSignal) {
        if (threadIdx.x == 0)
          s_halo[0] = ld_acquire_gpu(&hd.counters[0]) >= gridDim.x;
      } else if (hstate == kHFenced) {
        // the ticket of this CTA's slice has arrived: the last one raises
        // this rank's flags on the peers
        if (threadIdx.x == 0 && hticket == hd.num_slices) halo_raise_flags(hd);
      } else if (hstate == kHPoll && (hiter < 8 || (hiter & 3) == 0) &&
                 threadIdx.x <= hd.num_peers) {
        // Waiting for the peers' values.  Thread k polls peer k's flag (the
        // thread after the last peer: this rank's count of completed push
        // slices, after which y's shared dofs are no longer read).  The poll
        // is software-pipelined: a load issued here is looked at one element
        // step later, so a late peer costs nothing but these few instructions
        // (every step at first, every fourth step after eight).
        const bool is_flag = threadIdx.x < hd.num_peers;
        if (hpend < (is_flag ? hd.epoch : (uint64_t)hd.num_slices))
          s_halo[2] = 1;  // not yet
        hpend = is_flag ? ld_relaxed_sys(reinterpret_cast<const uint64_t*>(hpoll))
                        : (uint64_t)ld_relaxed_gpu(
                              reinterpret_cast<const unsigned*>(hpoll));
      }
    }

    // ---- u tile of this element has landed (cp.async issued one element ago)
    cp_async_wait_all();
    __syncthreads();

    if (HALO) {
      if (hstate == kHIface && blk >= hd.n_if_blocks) {
        // every thread fenced its y updates at the end of the last interface
        // step and has passed the barrier above
        if (threadIdx.x == 0) {
          __threadfence();
          atomicAdd(&hd.counters[0], 1u);
        }
        // fuse_push == 0: the push, too, is done by the concurrently running
        // wait kernel; this CTA's only halo duty was the signal above
        hstate = hd.fuse_push ? kHSignal : kHDone;
        if (stamp) halo_stamp(hd, 1);
      } else if (hstate == kHSignal && s_halo[0]) {
        if (stamp) halo_stamp(hd, 2);
        hstate = halo_push_issue<T>(hd, y, &s_halo[1])
                     ? kHIssued
                     : (hd.fuse_unpack ? kHPoll : kHDone);
        if (stamp) halo_stamp(hd, 3);
      } else if (hstate == kHIssued) {
        // the remote stores were issued a whole element step ago
        __threadfence_system();
        hstate = kHFenced;  // counted after the next barrier
      } else if (hstate == kHFenced) {
        hstate = hd.fuse_unpack ? kHPoll : kHDone;
      } else if (hstate == kHPoll && (hiter < 8 || (hiter & 3) == 0) &&
                 s_halo[2] == 0) {
        // all peers' values have arrived: canonical sum of the shared dofs
        // (no interior element touches them), also hidden under the interior
        asm volatile("fence.acq_rel.sys;" ::: "memory");
        if (stamp) halo_stamp(hd, 6);
        halo_unpack_slices<T>(hd, y, &s_halo[1], 1);
        hstate = kHDone;
        if (stamp) halo_stamp(hd, 7);
      }
    }

    // ---- phase 2 (mappings B, C): a1- and a2-derivatives
    if (lane_ok) {
      T col[N], out[N];
#pragma unroll
      for (int m = 0; m < N; ++m) col[m] = sU[baseB + m * R + (SWZ ? (qs ^ m) : qs)];
      eo_apply<T, N>(dm.fwd, col, out);
#pragma unroll
      for (int j = 0; j < N; ++j) sA[baseB + j * R + (SWZ ? (qs ^ j) : qs)] = out[j];
#pragma unroll
      for (int m = 0; m < N; ++m) col[m] = sU[baseC + (SWZ ? (m ^ qs) : m)];
      eo_apply<T, N>(dm.fwd, col, out);
#pragma unroll
      for (int j = 0; j < N; ++j) sB[baseC + (SWZ ? (j ^ qs) : j)] = out[j];
    }
    __syncthreads();

    if (HALO) {
      if (hstate == kHFenced && hticket == 0) {
        // all threads' fences precede the barrier above.  (hticket == 0: the
        // state is entered before this point and left after the next step's
        // first barrier, so it passes here exactly once.)
        if (threadIdx.x == 0) hticket = atomicAdd(&hd.counters[2], 1u) + 1u;
      } else if (hstate == kHPoll) {
        if (threadIdx.x == 0) s_halo[2] = 0;
        ++hiter;
      }
    }

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
      T d0[N], ucol[N];
#pragma unroll
      for (int m = 0; m < N; ++m) ucol[m] = sU[m * S0 + offA];
      eo_apply<T, N>(dm.fwd, ucol, d0);
      const T* g = STAGE ? sG + t
                         : gf + (active ? e : 0) * (int64_t)(ngeom * n) + t;
      if (STAGE) mbar_wait(&gbar, gphase);
      constexpr int KB = STAGE ? 1 : (KCH > 0 ? KCH : 1);
#pragma unroll
      for (int k0 = 0; k0 < N; k0 += KB) {
        T gg[KB][7];
#pragma unroll
        for (int kk = 0; kk < KB; ++kk) {
          const int k = k0 + kk;
          if (k < N) {
#pragma unroll
            for (int s = 0; s < ngeom; ++s) {
              if (STAGE)
                gg[kk][s] = active ? g[k * P + s * n] : T(0);
              else
                gg[kk][s] = active ? ld_stream(g + k * P + s * n) : T(0);
            }
          }
        }
#pragma unroll
        for (int kk = 0; kk < KB; ++kk) {
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
            if (MASS) ucol[k] *= lambda * gg[kk][6];
          }
        }
        // keep the compiler from hoisting every slab's loads to the top
        if (!STAGE) asm volatile("" ::: "memory");
      }
      eo_apply<T, N>(dm.bwd, d0, ry);
      if (MASS) {
#pragma unroll
        for (int k = 0; k < N; ++k) ry[k] += ucol[k];
      }
    }
    if (LAZY && threadIdx.x == 0) {
      s_lz[0] = lz_seen >= lz_need;
      s_lz[1] = lz_ticket;
    }
    __syncthreads();
    if (STAGE) {
      // every thread of the slot is done reading the staged factors: refill
      // the stage with the next element's chunk (lands during phases 4, 5 and
      // the next element's phase 2)
      gphase ^= 1;
      if (threadIdx.x == 0 && blk_n < nblocks) stage_copy(blk_n);
    }

    // ---- phase 4 (mappings B, C): transposed a1-, a2-derivatives, in place
    if (lane_ok) {
      T col[N], out[N];
#pragma unroll
      for (int m = 0; m < N; ++m) col[m] = sA[baseB + m * R + (SWZ ? (qs ^ m) : qs)];
      eo_apply<T, N>(dm.bwd, col, out);
#pragma unroll
      for (int j = 0; j < N; ++j) sA[baseB + j * R + (SWZ ? (qs ^ j) : qs)] = out[j];
#pragma unroll
      for (int m = 0; m < N; ++m) col[m] = sB[baseC + (SWZ ? (m ^ qs) : m)];
      eo_apply<T, N>(dm.bwd, col, out);
#pragma unroll
      for (int j = 0; j < N; ++j) sB[baseC + (SWZ ? (j ^ qs) : j)] = out[j];
    }
    __syncthreads();

    // ---- phase 5 (mapping A): sum the three parts, scatter
    if (LAZY) {
      // (s_lz was written before the barrier after phase 3 and is not written
      // again before the next step's barriers.)  The relaxed poll suffices on
      // the fast path: the zeros were fenced before the count, and the REDs
      // below are L2 operations issued after the count was seen.
      if (!s_lz[0]) {
        if (threadIdx.x == 0 && ld_relaxed_gpu(&lz.counters[1]) == 0 &&
            !spin_u32_relaxed_ge(&lz.counters[2 + lz_c], lz_need, 250,
                                 2000000000ull)) {
          // sticky: later steps do not wait again.  First failure: who / what
          unsigned* dbg = lz.counters + 2 + lz.num_chunks;
          if (atomicCAS(dbg, 0u, 1u) == 0u) {
            dbg[1] = blockIdx.x;
            dbg[2] = lz_c;
            dbg[3] = ld_relaxed_gpu(&lz.counters[2 + lz_c]);
            dbg[4] = ld_relaxed_gpu(&lz.counters[0]);
          }
          lz.counters[1] = 1u;
        }
        __syncthreads();
      }
    } else if (first_step) {
      pdl_wait();
      first_step = false;
    }
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
    if (HALO && hstate == kHIface && blk_n >= hd.n_if_blocks) __threadfence();
    // next step (LAZY: s_lz[1] was written before the barrier after phase 3)
    blk = blk1;
    blk1 = blk2;
    blk2 = LAZY ? (int64_t)gridDim.x + s_lz[1] : blk2 + gridDim.x;
  }
  cp_async_wait_all();
  if (HALO) {
    // finish whatever is in flight; from here on the blocking forms are used
    if (hstate == kHIssued) {
      __threadfence_system();
      __syncthreads();
      if (threadIdx.x == 0) hticket = atomicAdd(&hd.counters[2], 1u) + 1u;
      hstate = kHFenced;
    } else if (hstate == kHFenced && hticket == 0) {
      // (cannot happen: the count is taken in the step the state is entered)
      if (threadIdx.x == 0) hticket = atomicAdd(&hd.counters[2], 1u) + 1u;
    }
    if (hstate == kHFenced) {
      if (threadIdx.x == 0 && hticket == hd.num_slices) halo_raise_flags(hd);
      hstate = hd.fuse_unpack ? kHPoll : kHDone;
    }
    if (hstate == kHIface) {  // all steps of this CTA were interface steps
      __syncthreads();
      if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(&hd.counters[0], 1u);
      }
      hstate = hd.fuse_push ? kHSignal : kHDone;
    }
    if (hstate == kHSignal) {
      __syncthreads();
      if (threadIdx.x == 0)
        s_halo[0] = ld_acquire_gpu(&hd.counters[0]) >= gridDim.x;
      __syncthreads();
      if (s_halo[0]) {
        halo_push_slices<T>(hd, y, &s_halo[1]);
        hstate = hd.fuse_unpack ? kHPoll : kHDone;
      }
    }
    if (hstate == kHPoll) {
      // one last look; what is left is done by the wait kernel
      __syncthreads();
      if (threadIdx.x == 0) s_halo[0] = halo_peers_ready(hd);
      __syncthreads();
      if (s_halo[0]) halo_unpack_slices<T>(hd, y, &s_halo[1]);
    }
    if (stamp) halo_stamp(hd, 5);
  }
  if (want_dot) {
    dot = block_sum(dot, red);
    if (threadIdx.x == 0) atomicAdd(dot_xy, dot);
  }
}

// Companion of a LAZY apply (LazyDev in sfem_common.cuh): zeroes piece after
// piece of y's shared-dof prefix in first-touch order, paced by the apply's
// progress counter.  The pieces are a work queue (one warp claims one piece at
// a time), so it does not matter how many of this kernel's CTAs find room next
// to the apply's: one resident warp is enough for progress.  Every lane fences
// its own stores, lane 0 fences again after the warp has converged and counts
// the piece on its chunk.  Launched as the apply's programmatic dependent so
// that it runs NEXT to it; it waits for the primary only once its work is done.
template <typename T>
__global__ void __launch_bounds__(64)
lazy_zero_kernel(T* __restrict__ y, const LazyDev lz) {
  constexpr int VW = 16 / (int)sizeof(T);
  const unsigned lane = threadIdx.x & 31u;
  unsigned* head = lz.counters + 2 + lz.num_chunks + 8;
  bool dead = false;
  for (;;) {
    int p = 0;
    if (lane == 0) p = (int)atomicAdd(head, 1u);
    p = __shfl_sync(0xffffffffu, p, 0);
    if (p >= lz.num_pieces) break;
    const int2 pc = __ldg(lz.pieces + p);
    const unsigned j = (unsigned)pc.y >> 12;
    // steps claimed so far = grid + counters[0] (a step is claimed three steps
    // before it runs); chunk j is due once its first step is `ahead` steps
    // from being claimed
    const uint64_t need = (uint64_t)j * lz.chunk_steps;
    const uint64_t have = (uint64_t)lz.grid + lz.ahead;
    if (need > have && !dead) {
      int ok = 1;
      if (lane == 0)
        ok = spin_u32_relaxed_ge(&lz.counters[0], (unsigned)(need - have), 500,
                                 2000000000ull);
      ok = __shfl_sync(0xffffffffu, ok, 0);
      if (!ok) {
        dead = true;  // the apply made no progress: zero the rest unpaced
        if (lane == 0) {
          unsigned* dbg = lz.counters + 2 + lz.num_chunks;
          if (atomicCAS(dbg, 0u, 2u) == 0u) {
            dbg[1] = blockIdx.x;
            dbg[2] = j;
            dbg[3] = (unsigned)(need - have);
            dbg[4] = ld_relaxed_gpu(&lz.counters[0]);
          }
          lz.counters[1] = 1u;
        }
      }
    }
    T* d = y + pc.x;
    int len = pc.y & 0xfff;
    int hd = (int)(((16u - (unsigned)((uintptr_t)d & 15u)) & 15u) / sizeof(T));
    hd = hd < len ? hd : len;
    if ((int)lane < hd) d[lane] = T(0);
    d += hd;
    len -= hd;
    const int nv = len / VW;
    float4* d4 = reinterpret_cast<float4*>(d);
    for (int i = lane; i < nv; i += 32) d4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int tail = nv * VW + (int)lane;
    if (tail < len) d[tail] = T(0);
    __threadfence();
    __syncwarp();
    if (lane == 0) {
      __threadfence();
      red_add_u32(&lz.counters[2 + j], 1u);
    }
  }
  // The apply has consumed every counter once it has completed; the LAST CTA
  // of this kernel to get here resets them (the others may still be claiming).
  pdl_wait();
  __shared__ unsigned s_last;
  __syncthreads();
  if (threadIdx.x == 0)
    s_last = atomicAdd(head + 1, 1u) + 1u == gridDim.x;
  __syncthreads();
  if (s_last) {
    for (int i = threadIdx.x; i < lz.num_chunks + 2; i += blockDim.x)
      if (i != 1) lz.counters[i] = 0u;
    if (threadIdx.x == 0) {
      head[0] = 0u;
      head[1] = 0u;
    }
  }
}

template <typename T, int N, bool MASS, bool LOCAL, int EPB, int MINB, int KCH,
          bool HALO = false, bool CONN2 = false, bool EVICT = false,
          bool LAZY = false>
int launch3d_v2_cfg(const sfem_op& op, double lambda, double mu, const void* x,
                    void* y, int ncomp, double* dot_xy, cudaStream_t stream) {
  using C = Cfg3DV2<T, N, EPB, MINB, KCH>;
  const int64_t E = op.base.desc.num_elements;
  const int64_t nblocks = (E + C::epb - 1) / C::epb;
  const size_t smem =
      ((size_t)C::stage_off(C::epb) +
       (KCH == 0 ? (size_t)C::epb * (MASS ? 7 : 6) * C::n : 0)) *
      sizeof(T);
  auto kernel = apply3d_v2_kernel<T, N, MASS, LOCAL, EPB, MINB, KCH, HALO, CONN2,
                                  EVICT, LAZY>;
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
  // persistent CTAs: one wave, every CTA pipelines over its elements
  // ... and the wave is trimmed so that every CTA gets the same number of
  // steps (+-1): with 39 304 steps on 888 slots a quarter of the CTAs would
  // run a 45th step while the others idle; 874 CTAs x 45 steps do not
  const int64_t cap = (int64_t)num_sms() * per_sm;
  const int64_t steps_per_cta = (nblocks + cap - 1) / cap;
  const int64_t even = (nblocks + steps_per_cta - 1) / steps_per_cta;
  dim3 grid((unsigned)(nblocks < cap ? nblocks : even), ncomp);
  if (op.query) {  // table set-up: report the launch geometry only
    op.query[0] = (unsigned)C::epb;
    op.query[1] = grid.x;
    return SFEM_OK;
  }
  LazyDev lz{};
  unsigned lazy_ctas = 0;
  if (LAZY) {
    if (op.lazy_step_elems != (unsigned)C::epb || ncomp != 1 ||
        op.lazy_counters == nullptr || nblocks >= (1ll << 31)) {
      set_error("lazy zero fill: tables were built for another launch geometry");
      return SFEM_ERR_INVALID;
    }
    // Load the companion BEFORE the apply is enqueued: CUDA loads a kernel at
    // its first launch, and that load can block until running kernels finish
    // -- the apply's CTAs would sit in their bounded wait for a companion
    // that cannot be launched.
    static int loaded_dev[64] = {};
    int& loaded = per_device_slot(loaded_dev);
    if (!loaded) {
      cudaFuncAttributes fa;
      SFEM_CUDA_CHECK(cudaFuncGetAttributes(&fa, lazy_zero_kernel<T>));
      SFEM_CUDA_CHECK(cudaFuncGetAttributes(&fa, kernel));
      loaded = 1;
    }
    static int lazy_ctas_env = -1;
    if (lazy_ctas_env < 0) {
      const char* e = getenv("SFEM_LAZY_CTAS");
      lazy_ctas_env = e ? atoi(e) : 0;
    }
    lazy_ctas = lazy_ctas_env > 0 ? (unsigned)lazy_ctas_env : (unsigned)num_sms();
    lz.pieces = op.lazy_pieces;
    lz.chunk_ptr = op.lazy_chunk_ptr;
    lz.counters = op.lazy_counters;
    lz.num_chunks = op.lazy_num_chunks;
    lz.grid = grid.x;
    lz.ahead = op.lazy_ahead;
    lz.chunk_steps = op.lazy_chunk_steps;
    lz.num_pieces = op.lazy_num_pieces;
  }
  DOps<T, N> dm;
  fill_even_odd<T, N>(op.base.h_BD, false, &dm.fwd);
  fill_even_odd<T, N>(op.base.h_BD, true, &dm.bwd);
  HaloDev hd{};
  if (HALO) {
    if (op.fuse == nullptr || ncomp != 1 || op.fuse->num_peers >= 32) {
      set_error("fused halo push needs a halo with < 32 peers and ncomp == 1");
      return SFEM_ERR_INVALID;
    }
    // one work item per CTA (at least 32 entries): n_if_blocks arrives as an
    // ELEMENT count
    HaloDev& f = *op.fuse;
    auto size_for = [&](int64_t total) {
      int64_t per = (total + grid.x - 1) / grid.x;
      per = ((per + 31) / 32) * 32;
      return (unsigned)(per < 32 ? 32 : per);
    };
    f.slice = size_for(f.num_send);
    f.num_slices = (unsigned)((f.num_send + f.slice - 1) / f.slice);
    f.uslice = size_for(f.num_dofs);
    f.num_uslices = (unsigned)((f.num_dofs + f.uslice - 1) / f.uslice);
    f.apply_grid = grid.x;
    hd = f;
    hd.n_if_blocks = (hd.n_if_blocks + C::epb - 1) / C::epb;
  }
  SFEM_CUDA_CHECK(launch_maybe_pdl(
      op.pdl, kernel, grid, dim3(C::threads), smem, stream, dm, op.conn,
      (const T*)op.geom, (T)lambda, (T)mu, (const T*)x, (T*)y, ncomp, E,
      dot_xy, hd, lz));
  SFEM_LAUNCH_CHECK();
  if (LAZY) {
    SFEM_CUDA_CHECK(launch_maybe_pdl(true, lazy_zero_kernel<T>, dim3(lazy_ctas),
                                     dim3(64), 0, stream, (T*)y, lz));
    SFEM_LAUNCH_CHECK();
  }
  return SFEM_OK;
}

// Launch configuration for a given elements-per-CTA count: whether the factors
// are staged (bulk async copies) and how many CTAs an SM should hold (shared
// memory footprint and a register estimate).
template <typename T, int N, bool MASS, int EPB>
struct AutoCfg3D {
  using C0 = Cfg3DV2<T, N, EPB, 1, 0>;
  static constexpr long stage_bytes =
      ((long)C0::stage_off(EPB) + (long)EPB * (MASS ? 7 : 6) * C0::n) *
      (long)sizeof(T);
  static constexpr bool staged =
      stage_bytes <= 200 * 1024 &&
      ((EPB * (MASS ? 7 : 6) * C0::n * sizeof(T)) % 16) == 0;
  static constexpr int KCH = staged ? 0 : 2;
  static constexpr long smem_bytes =
      staged ? stage_bytes : (long)C0::stage_off(EPB) * (long)sizeof(T);
  static constexpr int by_smem =
      (int)((220L * 1024) / (smem_bytes > 0 ? smem_bytes : 1));
  // register estimate fitted to ptxas output (fp64: 114 @ N=5 ... 180 @ N=9;
  // fp32: 66 @ N=5, 93 @ N=9, 118 @ N=12)
  static constexpr int est_regs_raw =
      sizeof(T) == 8 ? 26 + 17 * N : (60 + 15 * N) / 2;
  static constexpr int est_regs = est_regs_raw > 255 ? 255 : est_regs_raw;
  static constexpr int by_regs = 65536 / (C0::threads * est_regs);
  static constexpr int m0 = by_smem < by_regs ? by_smem : by_regs;
  static constexpr int MINB = m0 < 1 ? 1 : (m0 > 8 ? 8 : m0);
};

// Measured on B200 (profiles/r02_variants_9_10_11_orders.txt, r02_bench_ne68_
// variant_*.json; 16 M and 108 M dofs): both options together are worth +10 %
// at N = 8 fp64 (81 % -> 89 % of the HBM roofline at 108 M dofs), +8 % at
// N = 6 fp64, +1.5 % at N = 8 fp32, and cost 1-4 % for the other measured
// (precision, N) -- those keep the one-step / default-policy pipeline.
// `lazy`: a LAZY instance (zero fill by the companion kernel) is compiled.
template <typename T, int N>
struct Tune3D {
  static constexpr bool conn2 = false, evict = false, lazy = false;
};
template <> struct Tune3D<double, 6> {
  static constexpr bool conn2 = true, evict = true, lazy = false;
};
template <> struct Tune3D<double, 8> {
  static constexpr bool conn2 = true, evict = true, lazy = true;
};
template <> struct Tune3D<float, 8> {
  static constexpr bool conn2 = true, evict = true, lazy = true;
};

constexpr int clamp_int(int v, int lo, int hi) {
  return v < lo ? lo : (v > hi ? hi : v);
}

// Elements per CTA step of the default configuration: small CTAs (>= 64
// threads fp64, >= 128 fp32).  An element's factors need not be a multiple of
// 16 bytes (odd N in fp32, or with the mass factor): a CTA step then takes the
// next element count whose chunk is (bulk copies need 16-byte aligned chunks).
template <typename T, int N, bool MASS>
constexpr int default_epb3d() {
  constexpr int P = N * N;
  constexpr int target = sizeof(T) == 8 ? 64 : 128;
  constexpr int epb_t = clamp_int((target + P - 1) / P, 1, 16);
  constexpr long gb = (long)(MASS ? 7 : 6) * N * N * N * (long)sizeof(T);
  return (epb_t * gb) % 16 == 0         ? epb_t
         : ((epb_t + 1) * gb) % 16 == 0 ? epb_t + 1
         : ((epb_t + 2) * gb) % 16 == 0 ? epb_t + 2
                                        : epb_t + 3;
}

template <typename T, int N, bool MASS, bool LOCAL>
int launch3d_v2(const sfem_op& op, double lambda, double mu, const void* x,
                void* y, int ncomp, double* dot_xy, cudaStream_t stream) {
  constexpr int EPB = default_epb3d<T, N, MASS>();
  using A = AutoCfg3D<T, N, MASS, EPB>;
#ifdef SFEM_EXPERIMENTS
  // pipeline variants for every precision and order (Laplacian, global form)
  if constexpr (!MASS && !LOCAL) {
    switch (op.variant) {
      case 9:  // connectivity fetched two steps ahead
        return launch3d_v2_cfg<T, N, MASS, LOCAL, EPB, A::MINB, A::KCH, false,
                               true, false>(op, lambda, mu, x, y, ncomp, dot_xy,
                                            stream);
      case 10:  // factors staged with an L2 evict-first policy
        return launch3d_v2_cfg<T, N, MASS, LOCAL, EPB, A::MINB, A::KCH, false,
                               false, true>(op, lambda, mu, x, y, ncomp, dot_xy,
                                            stream);
      case 11:  // both
        return launch3d_v2_cfg<T, N, MASS, LOCAL, EPB, A::MINB, A::KCH, false,
                               true, true>(op, lambda, mu, x, y, ncomp, dot_xy,
                                           stream);
      case 12:  // neither (the round-1 pipeline)
        return launch3d_v2_cfg<T, N, MASS, LOCAL, EPB, A::MINB, A::KCH, false,
                               false, false>(op, lambda, mu, x, y, ncomp,
                                             dot_xy, stream);
      default:
        break;
    }
  }
  // tuning variants (relative to the default), Laplacian, N = 5..9 only
  if constexpr (N >= 5 && N <= 9 && !MASS && !LOCAL) {
    constexpr int Ep = clamp_int(EPB + 1, 1, 16), Em = clamp_int(EPB - 1, 1, 16);
    constexpr int E2 = clamp_int(EPB * 2, 1, 16);
    using Ap = AutoCfg3D<T, N, MASS, Ep>;
    using Am = AutoCfg3D<T, N, MASS, Em>;
    using A2 = AutoCfg3D<T, N, MASS, E2>;
    switch (op.variant) {
      case 3:  // one more element per CTA
        return launch3d_v2_cfg<T, N, MASS, LOCAL, Ep, Ap::MINB, Ap::KCH>(
            op, lambda, mu, x, y, ncomp, dot_xy, stream);
      case 4:  // one fewer
        return launch3d_v2_cfg<T, N, MASS, LOCAL, Em, Am::MINB, Am::KCH>(
            op, lambda, mu, x, y, ncomp, dot_xy, stream);
      case 5:  // twice as many
        return launch3d_v2_cfg<T, N, MASS, LOCAL, E2, A2::MINB, A2::KCH>(
            op, lambda, mu, x, y, ncomp, dot_xy, stream);
      case 6:  // one more resident CTA (tighter register cap)
        return launch3d_v2_cfg<T, N, MASS, LOCAL, EPB,
                               clamp_int(A::MINB + 1, 1, 8), A::KCH>(
            op, lambda, mu, x, y, ncomp, dot_xy, stream);
      case 7:  // one fewer resident CTA (looser register cap)
        return launch3d_v2_cfg<T, N, MASS, LOCAL, EPB,
                               clamp_int(A::MINB - 1, 1, 8), A::KCH>(
            op, lambda, mu, x, y, ncomp, dot_xy, stream);
      case 8:  // no staging: factors streamed from L2 in batches of 2 slabs
        return launch3d_v2_cfg<T, N, MASS, LOCAL, EPB, A::MINB, 2>(
            op, lambda, mu, x, y, ncomp, dot_xy, stream);
      default:
        break;
    }
  }
#endif
  using Tn = Tune3D<T, N>;
  if constexpr (Tn::lazy && !LOCAL) {
    if (op.query) op.query[2] = 1u;  // a LAZY instance exists
    if (op.lazy)
      return launch3d_v2_cfg<T, N, MASS, LOCAL, EPB, A::MINB, A::KCH, false,
                             Tn::conn2, Tn::evict, true>(
          op, lambda, mu, x, y, ncomp, dot_xy, stream);
  }
  return launch3d_v2_cfg<T, N, MASS, LOCAL, EPB, A::MINB, A::KCH, false,
                         Tn::conn2 && !LOCAL, Tn::evict>(
      op, lambda, mu, x, y, ncomp, dot_xy, stream);
}

// Default configuration of launch3d_v2 with the halo push fused in.
template <typename T, int N, bool MASS>
int launch3d_v2_halo(const sfem_op& op, double lambda, double mu, const void* x,
                     void* y, double* dot_xy, cudaStream_t stream) {
  constexpr int EPB = default_epb3d<T, N, MASS>();
  using A = AutoCfg3D<T, N, MASS, EPB>;
  using Tn = Tune3D<T, N>;
  return launch3d_v2_cfg<T, N, MASS, false, EPB, A::MINB, A::KCH, true,
                         Tn::conn2, Tn::evict>(op, lambda, mu, x, y, 1, dot_xy,
                                               stream);
}

}  // namespace
}  // namespace sfem
