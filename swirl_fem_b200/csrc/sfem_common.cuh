// Shared device/host helpers for libswirl_b200 (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdio>
#include <mutex>
#include <string>

#include "../../include/swirl_b200.h"

namespace sfem {

// ---- error plumbing ---------------------------------------------------------
void set_error(const std::string& msg);
extern std::atomic<int64_t> g_launch_count;

#define SFEM_CUDA_CHECK(expr)                                                 \
  do {                                                                        \
    cudaError_t _e = (expr);                                                  \
    if (_e != cudaSuccess) {                                                  \
      ::sfem::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));  \
      return SFEM_ERR_CUDA;                                                   \
    }                                                                         \
  } while (0)

#define SFEM_LAUNCH_CHECK()                                                   \
  do {                                                                        \
    ::sfem::g_launch_count.fetch_add(1, std::memory_order_relaxed);           \
    cudaError_t _e = cudaGetLastError();                                      \
    if (_e != cudaSuccess) {                                                  \
      ::sfem::set_error(std::string("kernel launch: ") +                      \
                        cudaGetErrorString(_e));                              \
      return SFEM_ERR_CUDA;                                                   \
    }                                                                         \
  } while (0)

#define SFEM_REQUIRE(cond, msg)                                               \
  do {                                                                        \
    if (!(cond)) {                                                            \
      ::sfem::set_error(std::string("invalid argument: ") + (msg));           \
      return SFEM_ERR_INVALID;                                                \
    }                                                                         \
  } while (0)

// Frees device allocations on every exit path of a set-up routine (the CUDA
// check macros return early).
struct DeviceFrees {
  void* ptrs[8] = {};
  int n = 0;
  void add(void* p) {
    if (p && n < 8) ptrs[n++] = p;
  }
  ~DeviceFrees() {
    for (int i = 0; i < n; ++i) cudaFree(ptrs[i]);
  }
};

// SM count of the CURRENT device (cached per device ordinal: a process may
// drive several devices, and grids are sized from this).
inline int num_sms() {
  static int sms[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  int& n = sms[dev & 63];
  if (n == 0) {
    int v = 148;
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    n = v > 0 ? v : 148;
  }
  return n;
}

// Occupancy cache slot of the current device for a kernel's launcher
// (`static int per_sm_dev[64]` in the launcher; 0 = not queried yet).
inline int& per_device_slot(int (&slots)[64]) {
  int dev = 0;
  cudaGetDevice(&dev);
  return slots[dev & 63];
}

// ---- packed connectivity ------------------------------------------------------
// One int32 per element-local node.  Low 30 bits: global node id.  Bit 31:
// the node occurs exactly once in `elements` (plain store instead of RED).
// Bit 30: Dirichlet row (result is 0).  A SENTINEL slot is encoded as
// id = 0 with both flags set and never read or written.
constexpr uint32_t kConnIdMask = 0x3fffffffu;
constexpr uint32_t kConnSingle = 0x80000000u;
constexpr uint32_t kConnDirichlet = 0x40000000u;
constexpr uint32_t kConnSentinel = 0xffffffffu;

// ---- small device helpers -----------------------------------------------------
// Fire-and-forget reduction (REDG in SASS).  Written as explicit `red` PTX:
// ptxas turns an atomicAdd with an unused result into RED only when the kernel
// has no fences -- in the halo-fused apply it kept ATOMG (a full round trip
// per update, 15 % slower kernel).
__device__ __forceinline__ void red_add(double* addr, double v) {
  asm volatile("red.relaxed.gpu.global.add.f64 [%0], %1;" ::"l"(addr), "d"(v)
               : "memory");
}
__device__ __forceinline__ void red_add(float* addr, float v) {
  asm volatile("red.relaxed.gpu.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v)
               : "memory");
}

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum (any blockDim multiple of 32, <= 1024).  Result valid in
// thread 0.  `smem` needs 32 doubles.
__device__ __forceinline__ double block_sum(double v, double* smem) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nthreads = blockDim.x * blockDim.y * blockDim.z;
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  if (warp == 0) {
    const int nw = (nthreads + 31) >> 5;
    v = lane < nw ? smem[lane] : 0.0;
    v = warp_sum(v);
  }
  __syncthreads();
  return v;
}

// streaming (read-once) loads: keep them out of L1
template <typename T>
__device__ __forceinline__ T ld_stream(const T* p) {
  return __ldcs(p);
}

// ---- peer-memory halo exchange (device side, shared by the fused apply and the
//      standalone kernels of sfem_halo.cu) ------------------------------------------
// Everything the device side needs; filled per call by the host (`sfem_halo`
// handle, sfem_halo.cu) and passed to kernels by value.
struct HaloDev {
  // send side
  const int32_t* send_idx;    // (num_send) local dof of every send entry
  const uint64_t* send_dst;   // (num_send) address of the entry's slot in the
                              // peer's receive buffer of parity 0
  int64_t num_send;
  uint64_t parity_off;        // byte offset of this epoch's receive buffers
  const uint64_t* peer_flag;  // (num_peers) address of THIS rank's flag word
                              // (parity 0) in each peer's flag array
  uint64_t flag_parity_off;   // byte offset of this epoch's flag words
  int num_peers;
  unsigned num_slices;        // ceil(num_send / slice)
  unsigned slice;             // send entries per push work item
  unsigned uslice;            // shared dofs per canonical-sum work item
  // receive side (canonical sum, see sfem_halo_unpack_canonical)
  unsigned num_uslices;       // ceil(num_dofs / slice)
  int fuse_unpack;            // fused apply: also run the canonical sum
  int fuse_push;              // fused apply: its CTAs push the shared dofs
  unsigned apply_grid;        // CTAs of the fused apply (counters[0] target)
  const uint64_t* flags;      // this rank's flag words of this epoch's parity
  const int32_t* peer_ranks;  // (num_peers)
  const void* recv;           // this epoch's receive buffer
  const int32_t* dofs;
  const int32_t* row_ptr;
  const int32_t* src;
  int64_t num_dofs;
  unsigned* counters;         // [0] CTAs past the interface elements,
                              // [1] next push slice, [2] push slices done,
                              // [3] wait-kernel CTAs done, [4] wait timed out,
                              // [5] next unpack slice; words 8..23: eight
                              // 64-bit globaltimer stamps (diagnostics)
  uint64_t epoch;
  int64_t n_if_blocks;        // CTA steps that cover the interface elements
};

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned ld_relaxed_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint64_t ld_acquire_sys(const uint64_t* p) {
  uint64_t v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint64_t ld_relaxed_sys(const uint64_t* p) {
  uint64_t v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(uint64_t* p, uint64_t v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ void halo_stamp(const HaloDev& h, int slot) {
  uint64_t now;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
  reinterpret_cast<uint64_t*>(h.counters + 8)[slot] = now;
}

// One thread, after ALL push slices are complete and fenced: raise this
// rank's flag of the epoch on every peer (release at system scope).
__device__ __forceinline__ void halo_raise_flags(const HaloDev& h) {
  __threadfence_system();
  for (int k = 0; k < h.num_peers; ++k)
    st_release_sys(
        reinterpret_cast<uint64_t*>(h.peer_flag[k] + h.flag_parity_off),
        h.epoch);
  halo_stamp(h, 4);
}

// Cooperative push of the shared dofs into the peers' receive buffers (P2P
// stores over NVLink).  Called by ALL threads of a CTA (block-uniform); CTAs
// draw slices of the send list from a global counter until none is left.  The
// CTA that completes the last slice raises this rank's flag on every peer
// (release at system scope, after every writer fenced its stores).
// `y` values were produced by other CTAs (RED / stores): read them from L2.
// `max_slices`: how many slices this call may take (the fused apply takes ONE
// per element step so that no CTA donates more time than the others: with a
// static element partition the kernel ends with its slowest CTA).
template <typename T>
__device__ __forceinline__ void halo_push_slices(const HaloDev& h, const T* y,
                                                 unsigned* s_slice,
                                                 unsigned max_slices = ~0u) {
  for (unsigned it = 0; it < max_slices; ++it) {
    __syncthreads();
    if (threadIdx.x == 0) *s_slice = atomicAdd(&h.counters[1], 1u);
    __syncthreads();
    const unsigned s = *s_slice;
    if (s >= h.num_slices) break;
    const int64_t b = (int64_t)s * h.slice;
    const int64_t e = b + h.slice < h.num_send ? b + h.slice : h.num_send;
    for (int64_t i = b + threadIdx.x; i < e; i += blockDim.x) {
      const T v = __ldcg(y + __ldg(h.send_idx + i));
      *reinterpret_cast<T*>(__ldg(h.send_dst + i) + h.parity_off) = v;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned done = atomicAdd(&h.counters[2], 1u) + 1u;
      if (done == h.num_slices) {
        halo_raise_flags(h);
      }
    }
  }
}

// Split form of one push work item for the fused apply: claim a slice and
// issue its loads and remote stores WITHOUT waiting for them (no fence).  The
// caller fences one element step later (the stores have long been
// acknowledged by then), counts the slice as done after the next barrier and
// raises the flags if it was the last one.  Returns false when no slice was
// left.  Block-uniform.
template <typename T>
__device__ __forceinline__ bool halo_push_issue(const HaloDev& h, const T* y,
                                                unsigned* s_slice) {
  __syncthreads();
  if (threadIdx.x == 0) *s_slice = atomicAdd(&h.counters[1], 1u);
  __syncthreads();
  const unsigned s = *s_slice;
  if (s >= h.num_slices) return false;
  const int64_t b = (int64_t)s * h.slice;
  const int64_t e = b + h.slice < h.num_send ? b + h.slice : h.num_send;
  for (int64_t i = b + threadIdx.x; i < e; i += blockDim.x) {
    const T v = __ldcg(y + __ldg(h.send_idx + i));
    *reinterpret_cast<T*>(__ldg(h.send_dst + i) + h.parity_off) = v;
  }
  return true;
}

// One thread: have all of this rank's push slices completed (y's shared dofs
// are no longer read) AND has every peer raised its flag of this epoch?  The
// polls are relaxed; a positive answer is followed by a system-scope fence so
// that the receive buffers may be read afterwards.
__device__ __forceinline__ bool halo_peers_ready(const HaloDev& h) {
  bool ok = ld_relaxed_gpu(&h.counters[2]) >= h.num_slices;
  for (int k = 0; k < h.num_peers; ++k)
    ok = ok && ld_relaxed_sys(h.flags + h.peer_ranks[k]) >= h.epoch;
  if (ok) __threadfence_system();
  return ok;
}

// Cooperative canonical sum (block-uniform, like halo_push_slices):
// y[dof] = sum over all holders in ascending rank order; src < 0 stands for
// this rank's own value.  Precondition: halo_peers_ready() was observed (by a
// thread of this CTA before a barrier, or by the wait kernel).
template <typename T>
__device__ __forceinline__ void halo_unpack_slices(const HaloDev& h, T* y,
                                                   unsigned* s_slice,
                                                   unsigned max_slices = ~0u) {
  const T* recv = reinterpret_cast<const T*>(h.recv);
  for (unsigned it = 0; it < max_slices; ++it) {
    __syncthreads();
    if (threadIdx.x == 0) *s_slice = atomicAdd(&h.counters[5], 1u);
    __syncthreads();
    const unsigned s = *s_slice;
    if (s >= h.num_uslices) break;
    const int64_t b = (int64_t)s * h.uslice;
    const int64_t e = b + h.uslice < h.num_dofs ? b + h.uslice : h.num_dofs;
    for (int64_t i = b + threadIdx.x; i < e; i += blockDim.x) {
      const int32_t d = __ldg(h.dofs + i);
      const T own = __ldcg(y + d);
      T acc = T(0);
      const int32_t j1 = __ldg(h.row_ptr + i + 1);
      for (int32_t j = __ldg(h.row_ptr + i); j < j1; ++j) {
        const int32_t sj = __ldg(h.src + j);
        acc += sj < 0 ? own : __ldcg(recv + sj);
      }
      y[d] = acc;
    }
  }
}

// ---- peer-memory all-reduce of a few doubles (the CG scalars) ------------------
// Region of a rank: [parity 2][source rank `world`] slots of 64 bytes.
struct ScalarSlot {
  double v[4];
  uint64_t epoch;
  uint64_t pad[3];
};
static_assert(sizeof(ScalarSlot) == 64, "slot size");

// Device view of a `sfem_scalar_exchange` handle (world == 0: no exchange).
struct ScalarDev {
  int rank;
  int world;
  char* my_region;
  const uint64_t* peer_regions;  // device array (world) of region addresses
  unsigned* timed_out;
};

__device__ __forceinline__ void st_release_gpu(unsigned long long* p,
                                               unsigned long long v) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_gpu64(
    const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// All-reduce (sum) of `count` <= 4 doubles by ONE WARP (all 32 lanes call it;
// `mine` must hold the same values on every lane).  Lane t publishes this
// rank's values into rank t's region and raises the epoch there, then waits
// (bounded: ~4 s of globaltimer) for rank t's epoch in the local region; every
// lane then adds the ranks' values in ascending rank order, so the result is
// bitwise identical on every lane and every rank.  Returns false on a timeout
// (the sticky flag is set; the sums are then meaningless).
__device__ __forceinline__ bool scalar_allreduce_warp(const ScalarDev& sx,
                                                      uint64_t epoch,
                                                      double (&vals)[4],
                                                      int count) {
  const unsigned parity = (unsigned)(epoch & 1u);
  const int lane = threadIdx.x & 31;
  bool ok = true;
  for (int t = lane; t < sx.world; t += 32) {
    ScalarSlot* dst = reinterpret_cast<ScalarSlot*>(sx.peer_regions[t]) +
                      (parity * sx.world + sx.rank);
    for (int k = 0; k < count; ++k) dst->v[k] = vals[k];
    __threadfence_system();
    st_release_sys(&dst->epoch, epoch);
  }
  for (int t = lane; t < sx.world; t += 32) {
    const ScalarSlot* src = reinterpret_cast<const ScalarSlot*>(sx.my_region) +
                            (parity * sx.world + t);
    uint64_t t0 = 0;
    unsigned spins = 0;
    while (ld_acquire_sys(&src->epoch) < epoch) {
      if ((++spins & 1023u) == 0) {
        uint64_t now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (t0 == 0) t0 = now;
        if (now - t0 > 4000000000ull) {
          atomicExch(sx.timed_out, 1u);
          ok = false;
          break;
        }
      }
    }
  }
  ok = __all_sync(0xffffffffu, ok);
  // every lane reads every rank's slot, but only lane t acquired slot t:
  // the warp barrier orders the other lanes' loads after that acquire
  __syncwarp();
  const ScalarSlot* base =
      reinterpret_cast<const ScalarSlot*>(sx.my_region) + parity * sx.world;
  for (int k = 0; k < count; ++k) {
    double acc = 0.0;
    for (int t = 0; t < sx.world; ++t) acc += __ldcg(&base[t].v[k]);
    vals[k] = acc;
  }
  return ok;
}

// ---- lazy zero fill of y's shared-dof prefix (3-D fused apply) ------------------
// The shared dofs of y are accumulated with RED, so they must be zero first.
// Filling the whole prefix before the launch costs three DRAM accesses per
// shared dof instead of one: the zeros are written, evicted (the prefix is
// 320 MB at 108 M dofs, L2 is 126 MB), fetched again by the first RED and
// written back.  Lazily, a dof is zeroed shortly before the first element
// that touches it is scattered, so the line is still in L2 when the REDs
// arrive and goes to DRAM once.
//
// Who zeroes: a COMPANION kernel (lazy_zero_kernel, 64-thread CTAs in the warp
// slots the apply leaves free, launched as the apply's programmatic dependent
// so that it runs next to it) -- not the apply's own CTAs: four in-kernel
// variants reached 0.998 x the algorithmic DRAM traffic but ran 15-45 % slower
// (fences in the element CTAs, CTAs coupled to one another).
//
// When: the LAZY apply instance does not stride its CTA steps; after the first
// wave (step b for CTA b) every CTA CLAIMS its steps from a counter, three
// steps before it runs them.  The CTAs therefore work on one tight window of
// consecutive steps, and the counter tells the companion exactly how far the
// apply is: the elements are cut into chunks of `chunk_steps` consecutive
// steps, chunk j = the dofs FIRST touched by steps [j S, (j+1) S), and a chunk
// is zeroed when its first step is about to be claimed.  (With strided steps
// the CTAs drift apart by many steps: either the companion runs far ahead and
// its zeros are evicted again, or the fast CTAs wait -- measured, no gain.)
//
// Tables (host-built, sfem_op_set_lazy_zero): `pieces` = {first dof, len |
// chunk << 12}, sorted by chunk, a work queue for the companion's warps;
// `chunk_ptr[j]` = first piece of chunk j (dofs no element touches are zeroed
// with chunk 0).  Protocol on `counters`:
//   [0]      steps claimed after the first wave (atomicAdd by the apply's CTAs)
//   [1]      sticky: a wait timed out (the result is invalid)
//   [2 + j]  pieces of chunk j that are zeroed (the warp fences, then counts);
//            a CTA scatters a step of chunk j only after [2 + j] == the number
//            of pieces of chunk j (relaxed poll issued at the top of the step,
//            looked at before phase 5)
//   [2 + num_chunks ..]  8 words: record of the first timed-out wait; then the
//            queue head and the count of companion CTAs that are done
// The apply's CTAs never fence, never zero and never wait for one another;
// the last companion CTA resets the counters after the apply has completed.
// All polls are RELAXED loads (served by L2, no L1 invalidation on SMs that
// also run the apply); the consumers of the published zeros are REDs, i.e. L2
// operations issued after the count was seen.
struct LazyDev {
  const int2* pieces;
  const int32_t* chunk_ptr;  // (num_chunks + 1)
  unsigned* counters;
  int num_chunks;
  unsigned grid;             // the apply's grid = size of the first wave
  unsigned ahead;            // extra lead of the companion, in CTA steps
  unsigned chunk_steps;      // S
  int num_pieces;
};

struct LazyHost {
  std::mutex mu;
  cudaStream_t stream = nullptr;
  bool has_stream = false;
};

// Bounded wait (~`limit_ns` of globaltimer) for *p >= want; false on a timeout.
__device__ __forceinline__ bool spin_u32_ge(const unsigned* p, unsigned want,
                                            unsigned sleep_ns,
                                            uint64_t limit_ns) {
  uint64_t t0 = 0;
  unsigned spins = 0;
  while (ld_acquire_gpu(p) < want) {
    __nanosleep(sleep_ns);
    if ((++spins & 1023u) == 0) {
      uint64_t now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      if (now - t0 > limit_ns) return false;
    }
  }
  return true;
}

// The same with relaxed loads (no L1 invalidation per poll: the pollers share
// their SMs with the apply's CTAs).  For hints, and for waits whose consumers
// only issue L2 operations (RED) on the published data afterwards.
__device__ __forceinline__ bool spin_u32_relaxed_ge(const unsigned* p,
                                                    unsigned want,
                                                    unsigned sleep_ns,
                                                    uint64_t limit_ns) {
  uint64_t t0 = 0;
  unsigned spins = 0;
  while (ld_relaxed_gpu(p) < want) {
    __nanosleep(sleep_ns);
    if ((++spins & 255u) == 0) {
      uint64_t now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      if (now - t0 > limit_ns) return false;
    }
  }
  return true;
}

__device__ __forceinline__ void red_add_u32(unsigned* addr, unsigned v) {
  asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(addr), "r"(v)
               : "memory");
}

// symmetric index of (i,k), i<=k, in the packed d(d+1)/2 layout
__host__ __device__ constexpr int sym_index(int dim, int i, int k) {
  return dim == 1 ? 0
         : dim == 2 ? (i == 0 ? k : 2)
                    : (i == 0 ? k : (i == 1 ? 2 + k : 5));
}

// ---- handles --------------------------------------------------------------------
struct SpaceBase {
  sfem_space_desc desc;
  int n;  // N^dim
  int q;  // Q^dim
  // device copies of the 1-D tables in both precisions
  double* d_tables64 = nullptr;  // [B | BD | W] (Q*N, Q*N, Q)
  float* d_tables32 = nullptr;
  double h_B[SFEM_MAX_1D * SFEM_MAX_1D];
  double h_BD[SFEM_MAX_1D * SFEM_MAX_1D];
  double h_W[SFEM_MAX_1D];
};

int space_base_init(SpaceBase* s, const sfem_space_desc* desc);
void space_base_free(SpaceBase* s);

template <typename T>
inline const T* tables(const SpaceBase& s);
template <>
inline const double* tables<double>(const SpaceBase& s) {
  return s.d_tables64;
}
template <>
inline const float* tables<float>(const SpaceBase& s) {
  return s.d_tables32;
}

}  // namespace sfem

struct sfem_space {
  sfem::SpaceBase base;
  void* invjacs;
  void* jacdets;
  void* quad_coords;
};

struct sfem_op {
  sfem::SpaceBase base;
  const uint8_t* dirichlet;
  int with_mass;
  int ngeom;        // d(d+1)/2 (+1 with mass)
  void* geom;       // (E, ngeom, Q^d) dtype
  uint32_t* conn;   // (E, N^d) packed connectivity
  int64_t n_zero;   // y[0 .. n_zero) must be zeroed before an apply
  int variant;      // 0 auto, 1 generic
  // set on a shallow copy by sfem_op_apply_halo: fuse the halo push into the
  // apply kernel (interface elements first).  The launcher sizes the work
  // items for its grid (fields slice / uslice / num_slices / num_uslices).
  sfem::HaloDev* fuse = nullptr;
  // set on a shallow copy: the preceding stream operation is zero_fill_kernel,
  // launch the apply with programmatic dependent launch (its prologue --
  // connectivity, gather, first bulk copy -- overlaps the zero fill; it waits
  // with griddepcontrol.wait before its first write to y)
  bool pdl = false;
  // set on a shallow copy by the fused CG loop: y[0 .. n_zero) and the dot
  // accumulator were already zeroed by the previous cg_step_kernel
  bool prezeroed = false;
  // lazy zero fill (sfem_op_set_lazy_zero): device tables owned by the caller,
  // counters owned by the library.  `lazy_step_elems`: elements per CTA step
  // the tables were built for (checked by the launcher).
  const int2* lazy_pieces = nullptr;
  const int32_t* lazy_chunk_ptr = nullptr;
  unsigned* lazy_counters = nullptr;
  int lazy_num_chunks = 0, lazy_num_pieces = 0;
  unsigned lazy_step_elems = 0, lazy_chunk_steps = 0, lazy_ahead = 0;
  sfem::LazyHost* lazy_host = nullptr;  // the stream the lazy launches go to
  // set on a shallow copy: the launcher only reports {elements per CTA step,
  // grid, LAZY instance compiled} of the launch it would make (table set-up)
  unsigned* query = nullptr;
  // set on a shallow copy by op_apply_internal: zero lazily in this launch
  bool lazy = false;
};

// Peer-memory all-reduce handle (sfem_halo.cu).
struct sfem_scalar_exchange {
  int rank = 0, world = 0;
  char* my_region = nullptr;
  uint64_t* d_peer_regions = nullptr;  // device (world) addresses
  unsigned* d_timeout = nullptr;
  uint64_t epoch = 0;
};

namespace sfem {
inline ScalarDev scalar_view(const sfem_scalar_exchange* h) {
  ScalarDev d{};
  if (h) {
    d.rank = h->rank;
    d.world = h->world;
    d.my_region = h->my_region;
    d.peer_regions = h->d_peer_regions;
    d.timed_out = h->d_timeout;
  }
  return d;
}
// Zero fill of y's shared-dof prefix and of the dot accumulator by ONE small
// kernel (one CTA per SM) that lets its dependents launch at once.
// `*used_kernel` = false when y is not 16-byte aligned and plain memsets were
// enqueued instead (the apply must then be launched without the attribute).
int launch_zero_fill(void* y, size_t bytes, double* dot_xy, cudaStream_t stream,
                     bool* used_kernel);
// kernel<<<grid, block, smem, stream>>>(args...) with the programmatic stream
// serialization attribute when `pdl`.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_maybe_pdl(bool pdl, void (*kernel)(KArgs...),
                                    dim3 grid, dim3 block, size_t smem,
                                    cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
__device__ __forceinline__ void pdl_wait() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
__device__ __forceinline__ void pdl_launch_dependents() {
  asm volatile("griddepcontrol.launch_dependents;");
}
}  // namespace sfem
