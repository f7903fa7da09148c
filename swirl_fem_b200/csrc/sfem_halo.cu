// Peer-memory halo exchange: the partitioned QQ^T of
// swirl_fem/core/gather_scatter.py:221-261 (one dense lax.psum under pmap)
// restated as direct NVLink stores into the peers' receive buffers.
//
//   push    (fused into the 3-D apply kernel, or the standalone kernel below):
//           recv_peer[slot] = y[dof] for every (peer, shared dof) pair, then
//           flag_peer[parity][me] = epoch          (st.release.sys)
//   wait    flag_me[parity][peer] >= epoch for all peers (ld.acquire.sys),
//   unpack  y[dof] = sum over all holders in ascending rank order (canonical:
//           every rank evaluates the same expression -> replicated dofs stay
//           bitwise identical).
//
// Two receive buffers / flag sets alternate with the epoch parity: a peer can
// only start epoch e+2 after it has seen this rank's flag of epoch e+1, which
// this rank raises after its unpack of epoch e (stream order), so a buffer is
// never overwritten while it is still being read.

#include <cstdlib>
#include <cstring>

#include "sfem_common.cuh"

struct sfem_halo {
  sfem_halo_desc desc;
  uint64_t* d_peer_flag = nullptr;  // device copy of desc.peer_flag_addr
  int32_t* d_peer_ranks = nullptr;
  unsigned* d_counters = nullptr;   // 24 words, see HaloDev::counters
  uint64_t epoch = 0;
  unsigned slice = 256;      // default work-item size (standalone kernels)
  unsigned cur_uslice = 256; // canonical-sum work-item size of this epoch
  unsigned cur_slice = 256;  // push work-item size of this epoch
  unsigned cur_grid = 0;     // CTAs of this epoch's fused apply
  int fuse_unpack = 3;  // see sfem_halo_set_option key 1
  bool last_push_fused = false;  // the current epoch's push ran inside an apply
};

namespace sfem {

namespace {

// ONE warp: the all-reduce of sfem_common.cuh as a stand-alone launch.
__global__ void __launch_bounds__(32)
scalar_allreduce_kernel(double* __restrict__ values, int count,
                        const ScalarDev sx, uint64_t epoch) {
  pdl_wait();
  double v[4] = {0.0, 0.0, 0.0, 0.0};
  for (int k = 0; k < count; ++k) v[k] = values[k];
  scalar_allreduce_warp(sx, epoch, v, count);
  if ((int)threadIdx.x < count) values[threadIdx.x] = v[threadIdx.x];
}

}  // namespace

template <typename T>
int launch_apply3d_halo(const sfem_op& op, double lambda, double mu,
                        const void* x, void* y, double* dot_xy,
                        cudaStream_t stream);
int op_apply_internal(const sfem_op* op, double lambda, double mu,
                      const void* x, void* y, int ncomp, double* dot_xy,
                      cudaStream_t stream, bool prezeroed = false,
                      bool dot_prezeroed = false);
bool pdl_enabled();

namespace {

constexpr int kThreads = 256;

template <typename T>
__global__ void __launch_bounds__(kThreads)
halo_push_kernel(const T* __restrict__ y, const __grid_constant__ HaloDev hd) {
  __shared__ unsigned s_slice;
  halo_push_slices<T>(hd, y, &s_slice);
}

// Spins (bounded: ~4 s of globaltimer) until every peer raised its flag of
// this epoch, then runs whatever is left of the canonical sum (nothing, when
// the fused apply already did it).  The last CTA to finish resets the handle's
// counters for the next epoch.
// `early`: the kernel was launched as a programmatic dependent of a HALO apply
// whose CTAs do not run the canonical sum themselves: it does NOT wait for the
// whole apply grid first -- the sum needs only this rank's pushes complete
// (y's shared dofs final and no longer read) and the peers' flags, and no
// interior element touches a shared dof -- so it runs on the SMs' spare warp
// slots while the interior elements are still being computed, and waits for
// the apply grid only at its very end (stream order for whatever follows, and
// the counter reset).
template <typename T>
__global__ void __launch_bounds__(kThreads)
halo_wait_unpack_kernel(T* __restrict__ u, const __grid_constant__ HaloDev hd,
                        int early) {
  __shared__ unsigned s_slice;
  if (!early) pdl_wait();  // programmatic dependent of the apply kernel
  if (early && !hd.fuse_push) {
    // the push is ours as well: once every CTA of the apply has signalled its
    // interface elements (bounded spin), push the shared dofs to the peers
    if (threadIdx.x == 0) {
      uint64_t t0 = 0;
      unsigned spins = 0;
      while (ld_acquire_gpu(&hd.counters[0]) < hd.apply_grid) {
        __nanosleep(100);
        if ((++spins & 1023u) == 0) {
          uint64_t now;
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
          if (t0 == 0) t0 = now;
          if (now - t0 > 4000000000ull) {
            atomicExch(&hd.counters[4], 1u);
            break;
          }
        }
      }
    }
    __syncthreads();
    halo_push_slices<T>(hd, u, &s_slice);
  }
  if (early) {
    // bounded spin (thread 0) until all of this rank's push slices completed
    if (threadIdx.x == 0) {
      uint64_t t0 = 0;
      unsigned spins = 0;
      while (ld_acquire_gpu(&hd.counters[2]) < hd.num_slices) {
        __nanosleep(100);
        if ((++spins & 1023u) == 0) {
          uint64_t now;
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
          if (t0 == 0) t0 = now;
          if (now - t0 > 4000000000ull) {
            atomicExch(&hd.counters[4], 1u);
            break;
          }
        }
      }
    }
    __syncthreads();
  }
  // other CTAs claim slices concurrently: one thread decides for the CTA
  if (threadIdx.x == 0) s_slice = ld_relaxed_gpu(&hd.counters[5]);
  __syncthreads();
  if (s_slice < hd.num_uslices) {
    for (int k = threadIdx.x; k < hd.num_peers; k += blockDim.x) {
      const uint64_t* f = hd.flags + hd.peer_ranks[k];
      uint64_t t0 = 0;
      unsigned spins = 0;
      while (ld_acquire_sys(f) < hd.epoch) {
        if ((++spins & 1023u) == 0) {
          uint64_t now;
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
          if (t0 == 0) t0 = now;
          if (now - t0 > 4000000000ull) {
            atomicExch(&hd.counters[4], 1u);
            break;
          }
        }
      }
    }
    halo_unpack_slices<T>(hd, u, &s_slice);
  }
  __syncthreads();
  if (early) pdl_wait();  // the apply grid is through: counters may be reset
  if (threadIdx.x == 0) {
    const unsigned done = atomicAdd(&hd.counters[3], 1u) + 1u;
    if (done == gridDim.x) {
      hd.counters[0] = 0;
      hd.counters[1] = 0;
      hd.counters[2] = 0;
      hd.counters[3] = 0;
      hd.counters[5] = 0;
      __threadfence();
    }
  }
}

// Device view of the handle for its CURRENT epoch.
HaloDev device_view(const sfem_halo* h, int64_t num_interface_elements) {
  const sfem_halo_desc& d = h->desc;
  const unsigned parity = (unsigned)(h->epoch & 1u);
  HaloDev hd{};
  hd.send_idx = d.send_idx;
  hd.send_dst = d.send_dst;
  hd.num_send = d.num_send;
  hd.parity_off = parity ? d.parity_stride_bytes : 0;
  hd.peer_flag = h->d_peer_flag;
  hd.flag_parity_off = parity ? (uint64_t)d.world * 8u : 0u;
  hd.num_peers = d.num_peers;
  hd.slice = h->cur_slice;
  hd.num_slices = (unsigned)((d.num_send + h->cur_slice - 1) / h->cur_slice);
  hd.uslice = h->cur_uslice;
  hd.num_uslices = (unsigned)((d.num_dofs + hd.uslice - 1) / hd.uslice);
  hd.fuse_unpack = h->fuse_unpack == 1;
  hd.fuse_push = h->fuse_unpack != 3;
  hd.apply_grid = h->cur_grid;
  hd.flags = d.flags + (parity ? d.world : 0);
  hd.peer_ranks = h->d_peer_ranks;
  hd.recv = (const char*)d.recv + (parity ? d.parity_stride_bytes : 0);
  hd.dofs = d.dofs;
  hd.row_ptr = d.row_ptr;
  hd.src = d.src;
  hd.num_dofs = d.num_dofs;
  hd.counters = h->d_counters;
  hd.epoch = h->epoch;
  hd.n_if_blocks = num_interface_elements;
  return hd;
}

HaloDev begin_epoch(sfem_halo* h, int64_t num_interface_elements) {
  h->epoch += 1;
  h->cur_uslice = h->slice;
  h->cur_slice = h->slice;
  return device_view(h, num_interface_elements);
}

int push_standalone(sfem_halo* h, const HaloDev& hd, const void* u,
                    cudaStream_t stream) {
  int blocks = (int)hd.num_slices;
  const int cap = num_sms() * 4;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  if (h->desc.dtype == SFEM_F64)
    halo_push_kernel<double><<<blocks, kThreads, 0, stream>>>((const double*)u,
                                                              hd);
  else
    halo_push_kernel<float><<<blocks, kThreads, 0, stream>>>((const float*)u,
                                                             hd);
  SFEM_LAUNCH_CHECK();
  return SFEM_OK;
}

}  // namespace
}  // namespace sfem

namespace sfem {
// One predicate for "the exchange is pushed from inside the apply kernel".
static bool halo_fused(const sfem_op* op, const sfem_halo* halo) {
  const sfem_space_desc& d = op->base.desc;
  return op->variant == 0 && d.collocated && d.dim == 3 && d.n1d >= 2 &&
         d.n1d <= 16 && d.num_elements > 0 && halo->desc.num_peers < 32;
}

int op_apply_halo_internal(const sfem_op* op, sfem_halo* halo, double lambda,
                           double mu, const void* x, void* y,
                           int64_t num_interface_elements, double* dot_xy,
                           bool prezeroed, cudaStream_t stream) {
  sfem_stream_t stream_ = (sfem_stream_t)stream;
  SFEM_REQUIRE(op && halo && x && y, "null argument");
  SFEM_REQUIRE(x != y, "sfem_op_apply_halo is out of place");
  const sfem_space_desc& d = op->base.desc;
  SFEM_REQUIRE(d.dtype == halo->desc.dtype, "operator / halo dtype mismatch");
  SFEM_REQUIRE(num_interface_elements >= 0 &&
                   num_interface_elements <= d.num_elements,
               "num_interface_elements out of range");
  SFEM_REQUIRE(lambda == 0.0 || op->with_mass,
               "operator was created without mass factors but lambda != 0");
  if (!halo_fused(op, halo)) {
    int rc = op_apply_internal(op, lambda, mu, x, y, 1, dot_xy, stream,
                               prezeroed);
    if (rc) return rc;
    return sfem_halo_push(halo, y, stream_);
  }
  const size_t esz = d.dtype == SFEM_F64 ? 8 : 4;
  bool pdl = !prezeroed && (op->n_zero > 0 || dot_xy);
  if (pdl && (!pdl_enabled() ||
              esz * (size_t)op->n_zero > ((size_t)64 << 20))) {
    pdl = false;
    if (op->n_zero > 0)
      SFEM_CUDA_CHECK(cudaMemsetAsync(y, 0, esz * (size_t)op->n_zero, stream));
    if (dot_xy)
      SFEM_CUDA_CHECK(cudaMemsetAsync(dot_xy, 0, sizeof(double), stream));
  } else if (pdl) {
    int rc = launch_zero_fill(y, esz * (size_t)op->n_zero, dot_xy, stream,
                              &pdl);
    if (rc) return rc;
  }
  // the epoch advances only once the launch has succeeded: a failed launch
  // must not shift this rank's parity against its peers
  const uint64_t epoch0 = halo->epoch;
  const unsigned uslice0 = halo->cur_uslice;
  HaloDev hd = begin_epoch(halo, num_interface_elements);
  sfem_op sub = *op;
  sub.fuse = &hd;  // the launcher sizes the work items for its grid
  sub.pdl = pdl;
  const int rc = d.dtype == SFEM_F64
                     ? launch_apply3d_halo<double>(sub, lambda, mu, x, y,
                                                   dot_xy, stream)
                     : launch_apply3d_halo<float>(sub, lambda, mu, x, y,
                                                  dot_xy, stream);
  if (rc) {
    halo->epoch = epoch0;
    halo->cur_uslice = uslice0;
    halo->cur_slice = halo->slice;
    return rc;
  }
  halo->cur_uslice = hd.uslice;
  halo->cur_slice = hd.slice;  // the wait kernel counts the same push slices
  halo->cur_grid = hd.apply_grid;
  halo->last_push_fused = true;
  return rc;
}


}  // namespace sfem

extern "C" {

int sfem_ipc_alloc(int64_t bytes, void** dev_ptr, void* handle) {
  using namespace sfem;
  SFEM_REQUIRE(bytes > 0 && dev_ptr && handle, "bad argument");
  void* p = nullptr;
  SFEM_CUDA_CHECK(cudaMalloc(&p, (size_t)bytes));
  SFEM_CUDA_CHECK(cudaMemset(p, 0, (size_t)bytes));
  SFEM_CUDA_CHECK(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  static_assert(sizeof(h) == SFEM_IPC_HANDLE_BYTES, "IPC handle size");
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    set_error(std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e));
    return SFEM_ERR_CUDA;
  }
  memcpy(handle, &h, sizeof(h));
  *dev_ptr = p;
  return SFEM_OK;
}

int sfem_ipc_open(const void* handle, void** dev_ptr) {
  using namespace sfem;
  SFEM_REQUIRE(handle && dev_ptr, "null argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  SFEM_CUDA_CHECK(
      cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return SFEM_OK;
}

int sfem_ipc_close(void* dev_ptr) {
  using namespace sfem;
  if (dev_ptr) SFEM_CUDA_CHECK(cudaIpcCloseMemHandle(dev_ptr));
  return SFEM_OK;
}

int sfem_ipc_free(void* dev_ptr) {
  using namespace sfem;
  if (dev_ptr) SFEM_CUDA_CHECK(cudaFree(dev_ptr));
  return SFEM_OK;
}

int64_t sfem_scalar_region_bytes(int32_t world) {
  return world > 0 ? (int64_t)2 * world * 64 : -1;
}

int sfem_scalar_exchange_create(int32_t rank, int32_t world, void* my_region,
                                const uint64_t* peer_regions,
                                sfem_scalar_exchange** out) {
  using namespace sfem;
  SFEM_REQUIRE(out && my_region && peer_regions, "null argument");
  SFEM_REQUIRE(world >= 1 && world <= 128 && rank >= 0 && rank < world,
               "bad rank / world");
  SFEM_REQUIRE(peer_regions[rank] == (uint64_t)(uintptr_t)my_region,
               "peer_regions[rank] must be this rank's region");
  auto* h = new sfem_scalar_exchange();
  h->rank = rank;
  h->world = world;
  h->my_region = (char*)my_region;
  cudaError_t e = cudaMalloc(&h->d_peer_regions, sizeof(uint64_t) * world);
  if (e == cudaSuccess) e = cudaMalloc(&h->d_timeout, sizeof(unsigned));
  if (e == cudaSuccess)
    e = cudaMemcpy(h->d_peer_regions, peer_regions, sizeof(uint64_t) * world,
                   cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemset(h->d_timeout, 0, sizeof(unsigned));
  if (e != cudaSuccess) {
    set_error(std::string("sfem_scalar_exchange_create: ") +
              cudaGetErrorString(e));
    sfem_scalar_exchange_destroy(h);
    return SFEM_ERR_CUDA;
  }
  *out = h;
  return SFEM_OK;
}

void sfem_scalar_exchange_destroy(sfem_scalar_exchange* h) {
  if (!h) return;
  cudaFree(h->d_peer_regions);
  cudaFree(h->d_timeout);
  delete h;
}

int sfem_scalar_allreduce(sfem_scalar_exchange* h, double* values,
                          int32_t count, sfem_stream_t stream) {
  using namespace sfem;
  SFEM_REQUIRE(h && values, "null argument");
  SFEM_REQUIRE(count >= 1 && count <= 4, "count must be 1..4");
  h->epoch += 1;
  SFEM_CUDA_CHECK(launch_maybe_pdl(
      true, scalar_allreduce_kernel, dim3(1), dim3(32), 0,
      (cudaStream_t)stream, values, (int)count, scalar_view(h), h->epoch));
  SFEM_LAUNCH_CHECK();
  return SFEM_OK;
}

int sfem_scalar_exchange_timed_out(const sfem_scalar_exchange* h,
                                   sfem_stream_t stream) {
  using namespace sfem;
  SFEM_REQUIRE(h, "null argument");
  unsigned v = 0;
  SFEM_CUDA_CHECK(cudaMemcpyAsync(&v, h->d_timeout, sizeof(v),
                                  cudaMemcpyDeviceToHost,
                                  (cudaStream_t)stream));
  SFEM_CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
  return v != 0 ? 1 : 0;
}

int sfem_halo_create(const sfem_halo_desc* desc, sfem_halo** halo) {
  using namespace sfem;
  SFEM_REQUIRE(desc && halo, "null argument");
  SFEM_REQUIRE(desc->dtype == SFEM_F32 || desc->dtype == SFEM_F64, "bad dtype");
  SFEM_REQUIRE(desc->num_peers >= 1 && desc->num_peers <= kThreads,
               "num_peers out of range");
  SFEM_REQUIRE(desc->world >= 2 && desc->rank >= 0 && desc->rank < desc->world,
               "bad rank / world");
  SFEM_REQUIRE(desc->num_send > 0 && desc->send_idx && desc->send_dst &&
                   desc->peer_flag_addr && desc->peer_ranks && desc->flags &&
                   desc->recv,
               "null send / receive arrays");
  SFEM_REQUIRE(desc->num_dofs > 0 && desc->dofs && desc->row_ptr && desc->src,
               "null canonical-sum arrays");
  auto* h = new sfem_halo();
  h->desc = *desc;
  const size_t np = (size_t)desc->num_peers;
  cudaError_t e = cudaMalloc(&h->d_peer_flag, np * sizeof(uint64_t));
  if (e == cudaSuccess) e = cudaMalloc(&h->d_peer_ranks, np * sizeof(int32_t));
  if (e == cudaSuccess) e = cudaMalloc(&h->d_counters, 24 * sizeof(unsigned));
  if (e == cudaSuccess)
    e = cudaMemcpy(h->d_peer_flag, desc->peer_flag_addr, np * sizeof(uint64_t),
                   cudaMemcpyHostToDevice);
  if (e == cudaSuccess)
    e = cudaMemcpy(h->d_peer_ranks, desc->peer_ranks, np * sizeof(int32_t),
                   cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemset(h->d_counters, 0, 24 * sizeof(unsigned));
  if (e != cudaSuccess) {
    set_error(std::string("sfem_halo_create: ") + cudaGetErrorString(e));
    sfem_halo_destroy(h);
    return SFEM_ERR_CUDA;
  }
  h->desc.peer_flag_addr = nullptr;  // host arrays are not retained
  h->desc.peer_ranks = nullptr;
  *halo = h;
  return SFEM_OK;
}

void sfem_halo_destroy(sfem_halo* halo) {
  if (!halo) return;
  cudaFree(halo->d_peer_flag);
  cudaFree(halo->d_peer_ranks);
  cudaFree(halo->d_counters);
  delete halo;
}

int sfem_halo_set_option(sfem_halo* halo, int32_t key, int64_t value) {
  using namespace sfem;
  SFEM_REQUIRE(halo, "null argument");
  switch (key) {
    case 0:
      SFEM_REQUIRE(value >= 32 && value <= (1 << 20), "slice out of range");
      halo->slice = (unsigned)value;
      return SFEM_OK;
    case 1:
      // 0: canonical sum in the wait kernel after the apply; 1: in the apply
      // kernel's own CTAs; 2: in the wait kernel, CONCURRENTLY with the apply's
      // interior elements (it is launched as a programmatic dependent and
      // only needs this rank's pushes and the peers' flags, not the interior)
      // 3: as 2, and the PUSH is done there too (after every CTA of the apply
      // has signalled its interface elements): the apply's CTAs only signal
      SFEM_REQUIRE(value >= 0 && value <= 3, "fuse_unpack must be 0..3");
      halo->fuse_unpack = (int)value;
      return SFEM_OK;
    default:
      set_error("sfem_halo_set_option: unknown key");
      return SFEM_ERR_INVALID;
  }
}

int sfem_halo_push(sfem_halo* halo, const void* u, sfem_stream_t stream) {
  using namespace sfem;
  SFEM_REQUIRE(halo && u, "null argument");
  halo->last_push_fused = false;
  const HaloDev hd = begin_epoch(halo, 0);
  return push_standalone(halo, hd, u, (cudaStream_t)stream);
}

int sfem_halo_wait_unpack(sfem_halo* halo, void* u, sfem_stream_t stream_) {
  using namespace sfem;
  cudaStream_t stream = (cudaStream_t)stream_;
  SFEM_REQUIRE(halo && u, "null argument");
  SFEM_REQUIRE(halo->epoch > 0, "wait_unpack before the first push");
  const sfem_halo_desc& d = halo->desc;
  const HaloDev hd = device_view(halo, 0);
  // sized for the case that the whole sum is still to do (one slice per CTA,
  // at most 2 CTAs per SM); when the fused apply already did it the CTAs
  // return at once
  // SFEM_WAIT_PDL=0 / SFEM_WAIT_CTAS=n: developer switches (launch the wait
  // kernel as an ordinary dependent / cap its grid)
  static const bool wait_pdl = [] {
    const char* e = getenv("SFEM_WAIT_PDL");
    return !(e && e[0] == '0');
  }();
  static const int wait_ctas = [] {
    const char* e = getenv("SFEM_WAIT_CTAS");
    return e ? atoi(e) : 0;
  }();
  int64_t b = hd.num_uslices;
  int64_t cap = (int64_t)num_sms() * 2;
  if (wait_ctas > 0) cap = wait_ctas;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  // concurrent mode: only behind a fused apply of THIS epoch (the kernel then
  // spins on counters the apply advances), one CTA per SM so that it fits the
  // spare warp slots; never without the programmatic-dependent attribute
  const int early = halo->fuse_unpack >= 2 && halo->last_push_fused && wait_pdl;
  // concurrent mode: 64-thread CTAs, so that they fit next to the apply's
  // resident CTAs (5 x 64 threads x 168 registers leave ~11 k registers per
  // SM: a 256-thread CTA would only start once apply CTAs exit -- measured:
  // flags raised AFTER the apply's exit), four per SM
  const int threads = early ? 64 : kThreads;
  if (early) {
    b = hd.num_uslices > hd.num_slices ? hd.num_uslices : hd.num_slices;
    const int64_t ecap = wait_ctas > 0 ? wait_ctas : (int64_t)num_sms() * 4;
    if (b > ecap) b = ecap;
    if (b < 1) b = 1;
  }
  if (d.dtype == SFEM_F64)
    SFEM_CUDA_CHECK(launch_maybe_pdl(wait_pdl, halo_wait_unpack_kernel<double>,
                                     dim3((unsigned)b), dim3(threads), 0,
                                     stream, (double*)u, hd, early));
  else
    SFEM_CUDA_CHECK(launch_maybe_pdl(wait_pdl, halo_wait_unpack_kernel<float>,
                                     dim3((unsigned)b), dim3(threads), 0,
                                     stream, (float*)u, hd, early));
  SFEM_LAUNCH_CHECK();
  return SFEM_OK;
}

int sfem_halo_timed_out(const sfem_halo* halo, sfem_stream_t stream) {
  using namespace sfem;
  SFEM_REQUIRE(halo, "null argument");
  unsigned v = 0;
  SFEM_CUDA_CHECK(cudaMemcpyAsync(&v, halo->d_counters + 4, sizeof(v),
                                  cudaMemcpyDeviceToHost,
                                  (cudaStream_t)stream));
  SFEM_CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
  return v != 0 ? 1 : 0;
}

int sfem_halo_debug_times(const sfem_halo* halo, uint64_t* out8,
                          sfem_stream_t stream) {
  using namespace sfem;
  SFEM_REQUIRE(halo && out8, "null argument");
  SFEM_CUDA_CHECK(cudaMemcpyAsync(out8, halo->d_counters + 8,
                                  8 * sizeof(uint64_t), cudaMemcpyDeviceToHost,
                                  (cudaStream_t)stream));
  SFEM_CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
  return SFEM_OK;
}

int sfem_op_apply_halo(const sfem_op* op, sfem_halo* halo, double lambda,
                       double mu, const void* x, void* y,
                       int64_t num_interface_elements, void* dot_xy,
                       sfem_stream_t stream) {
  return sfem::op_apply_halo_internal(op, halo, lambda, mu, x, y,
                                      num_interface_elements, (double*)dot_xy,
                                      false, (cudaStream_t)stream);
}

}  // extern "C"
