// Gather / scatter / exchange kernels (K1, K7, K9) and connectivity packing.
//
// Reference semantics: swirl_fem/core/gather_scatter.py:121-133 (gather,
// scatter), :189-261 (exchange).  These are HBM-bound index kernels: one
// coalesced pass over the int32 index stream, 128-bit loads where alignment
// allows, grid sized in multiples of the SM count.

#include <cub/cub.cuh>

#include "sfem_common.cuh"

namespace sfem {

namespace {

constexpr int kThreads = 256;

inline int blocks_for(int64_t n, int per_thread = 1) {
  int64_t b = (n + (int64_t)kThreads * per_thread - 1) /
              ((int64_t)kThreads * per_thread);
  const int64_t cap = (int64_t)num_sms() * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
gather_kernel(const T* __restrict__ u, const int32_t* __restrict__ idx,
              int64_t count, T fill, int stride, int offset,
              T* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t g = __ldcs(idx + i);
    if (offset >= 0) {
      out[i * stride + offset] =
          g == SFEM_SENTINEL ? fill : __ldg(u + (int64_t)g * stride + offset);
    } else {  // every component of the AoS field in ONE launch
      for (int c = 0; c < stride; ++c)
        out[i * stride + c] =
            g == SFEM_SENTINEL ? fill : __ldg(u + (int64_t)g * stride + c);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
scatter_add_kernel(const T* __restrict__ u_local,
                   const int32_t* __restrict__ idx, int64_t count, int stride,
                   int offset, T* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t g = __ldcs(idx + i);
    if (g != SFEM_SENTINEL) {
      if (offset >= 0) {
        red_add(out + (int64_t)g * stride + offset,
                __ldcs(u_local + i * stride + offset));
      } else {
        for (int c = 0; c < stride; ++c)
          red_add(out + (int64_t)g * stride + c,
                  __ldcs(u_local + i * stride + c));
      }
    }
  }
}

// ---- deterministic scatter: warp-segmented reduction over the sorted map -----
//
// `perm` lists the local slots sorted by global node id, `keys` the matching
// node ids.  One thread per sorted entry; a warp reduces its 32 entries with a
// segmented shuffle scan.  A segment that starts in a warp is finished by that
// warp (serial tail past the warp's window), a warp skips entries that
// continue a segment started before its window: every node is summed in
// ascending slot order by exactly one lane -> bitwise reproducible, no atomics.
template <typename T>
__global__ void __launch_bounds__(kThreads)
scatter_segmented_kernel(const T* __restrict__ u_local,
                         const int32_t* __restrict__ keys,
                         const int32_t* __restrict__ perm, int64_t nvalid,
                         int stride, int offset, T* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp_global =
      (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t base = warp_global * 32; base < nvalid; base += nwarps * 32) {
    const int64_t j = base + lane;
    const bool valid = j < nvalid;
    const int32_t key = valid ? keys[j] : -2;
    T val = valid ? u_local[(int64_t)perm[j] * stride + offset] : T(0);
    const int32_t prev_key = __shfl_up_sync(0xffffffffu, key, 1);
    const int32_t before = base > 0 ? keys[base - 1] : -3;
    // true segment heads (first slot of a node in the global sorted order)
    const bool head = key != (lane == 0 ? before : prev_key);
    const unsigned heads = __ballot_sync(0xffffffffu, head);
    const unsigned le = heads & (0xffffffffu >> (31 - lane));
    const bool owned = le != 0;  // my segment starts inside this window
    const int my_head = owned ? 31 - __clz(le) : 0;
    // segmented inclusive scan: sums restart at heads
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const T up = __shfl_up_sync(0xffffffffu, val, o);
      if (lane - o >= my_head) val += up;
    }
    const int32_t next_key = __shfl_down_sync(0xffffffffu, key, 1);
    const bool tail = valid && (lane == 31 || next_key != key);
    if (tail && owned) {
      T sum = val;
      if (lane == 31) {  // the segment may run past the window: finish it
        for (int64_t t = base + 32; t < nvalid && keys[t] == key; ++t)
          sum += u_local[(int64_t)perm[t] * stride + offset];
      }
      out[(int64_t)key * stride + offset] = sum;
    }
  }
}

__global__ void iota_kernel(int32_t* p, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    p[i] = (int32_t)i;
}

__global__ void remap_sentinel_kernel(const int32_t* in, int32_t* out,
                                      int64_t n, int32_t big) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t g = in[i];
    out[i] = g == SFEM_SENTINEL ? big : g;
  }
}

// ---- exchange -----------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads)
exchange_sum_kernel(const T* __restrict__ u, const int32_t* __restrict__ gi,
                    const int32_t* __restrict__ ui, int64_t count, int stride,
                    int offset, T* __restrict__ scratch) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t g = gi[i];
    if (g != SFEM_SENTINEL) {
      const int64_t slot = ui ? ui[i] : (int32_t)i;
      if (offset >= 0) {
        red_add(scratch + slot, u[(int64_t)g * stride + offset]);
      } else {  // scratch is (num_unique, stride)
        for (int c = 0; c < stride; ++c)
          red_add(scratch + slot * stride + c, u[(int64_t)g * stride + c]);
      }
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
exchange_write_kernel(T* __restrict__ u, const int32_t* __restrict__ gi,
                      const int32_t* __restrict__ ui, int64_t count, int stride,
                      int offset, const T* __restrict__ scratch) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t g = gi[i];
    if (g != SFEM_SENTINEL) {
      const int64_t slot = ui ? ui[i] : (int32_t)i;
      const int c0 = offset >= 0 ? offset : 0;
      const int c1 = offset >= 0 ? offset + 1 : stride;
      for (int c = c0; c < c1; ++c) {
        T* p = u + (int64_t)g * stride + c;
        const T initial = *p;
        // same expression as gather_scatter.py:261: u + (updates - initial)
        *p = initial +
             (scratch[offset >= 0 ? slot : slot * stride + c] - initial);
      }
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
pack_kernel(const T* __restrict__ u, const int32_t* __restrict__ idx,
            int64_t count, T* __restrict__ buf) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count;
       i += (int64_t)gridDim.x * blockDim.x)
    buf[i] = u[idx[i]];
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
unpack_add_kernel(T* __restrict__ u, const int32_t* __restrict__ idx,
                  int64_t count, const T* __restrict__ buf) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count;
       i += (int64_t)gridDim.x * blockDim.x)
    u[idx[i]] += buf[i];  // idx entries are unique within one message
}

// Canonical halo sum: u[dof] = sum over ALL holders of the dof (this rank and
// its peers) in ascending rank order, starting from 0.  Every rank that holds
// the dof evaluates the same floating-point expression, so the replicated
// values stay bitwise identical across ranks (a chain of in-place adds would
// associate differently on every rank for dofs with >= 3 holders).
// CSR over the unique interface dofs; src >= 0: position in `recv`, src < 0:
// this rank's own value.
template <typename T>
__global__ void __launch_bounds__(kThreads)
unpack_canonical_kernel(T* __restrict__ u, const int32_t* __restrict__ dofs,
                        const int32_t* __restrict__ row_ptr,
                        const int32_t* __restrict__ src, int64_t num_dofs,
                        const T* __restrict__ recv) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < num_dofs;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t d = dofs[i];
    const T own = u[d];
    T acc = T(0);
    for (int32_t j = row_ptr[i]; j < row_ptr[i + 1]; ++j) {
      const int32_t s = src[j];
      acc += s < 0 ? own : recv[s];
    }
    u[d] = acc;
  }
}

// ---- pointwise forms of the Stokes operators at the quadrature points ---------
// (navier_stokes.py:238-245, 313-329; AoS layouts of sfem_space_eval)
//   kind 0  trace:    out[p]       = sum_k g[p][k][k]           (div v)
//   kind 1  diagonal: out[p][j][k] = (j == k) a[p]              (q I, coefficient
//                                                                of grad v in div(v) q)
//   kind 2  convect:  out[p][k]    = sum_i a[p][i] g[p][i][k]   ((u . grad) w)
template <typename T>
__global__ void __launch_bounds__(kThreads)
pointwise_kernel(int kind, int d, const T* __restrict__ a,
                 const T* __restrict__ g, int64_t npts, T* __restrict__ out) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < npts;
       p += (int64_t)gridDim.x * blockDim.x) {
    if (kind == 0) {
      T acc = T(0);
      for (int k = 0; k < d; ++k) acc += g[(p * d + k) * d + k];
      out[p] = acc;
    } else if (kind == 1) {
      const T v = a[p];
      for (int j = 0; j < d; ++j)
        for (int k = 0; k < d; ++k) out[(p * d + j) * d + k] = j == k ? v : T(0);
    } else {
      for (int k = 0; k < d; ++k) {
        T acc = T(0);
        for (int i = 0; i < d; ++i) acc += a[p * d + i] * g[(p * d + i) * d + k];
        out[p * d + k] = acc;
      }
    }
  }
}

// ---- zero fill with early dependent launch -----------------------------------
__global__ void __launch_bounds__(256)
zero_fill_kernel(uint4* __restrict__ p16, size_t n16, unsigned char* tail,
                 int ntail, double* dot_xy) {
  pdl_launch_dependents();
  const uint4 z = make_uint4(0u, 0u, 0u, 0u);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16;
       i += (size_t)gridDim.x * blockDim.x)
    p16[i] = z;
  if (blockIdx.x == 0) {
    if ((int)threadIdx.x < ntail) tail[threadIdx.x] = 0;
    if (threadIdx.x == 0 && dot_xy) *dot_xy = 0.0;
  }
}

// ---- connectivity packing -------------------------------------------------------
__global__ void count_kernel(const int32_t* __restrict__ elements, int64_t total,
                             int32_t* __restrict__ counts) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t g = elements[i];
    if (g != SFEM_SENTINEL) atomicAdd(counts + g, 1);
  }
}

__global__ void pack_conn_kernel(const int32_t* __restrict__ elements,
                                 int64_t total,
                                 const int32_t* __restrict__ counts,
                                 const uint8_t* __restrict__ dirichlet,
                                 uint32_t* __restrict__ conn) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t g = elements[i];
    uint32_t c = kConnSentinel;
    if (g != SFEM_SENTINEL) {
      c = (uint32_t)g;
      if (counts[g] == 1) c |= kConnSingle;
      if (dirichlet && dirichlet[g]) c |= kConnDirichlet;
    }
    conn[i] = c;
  }
}

__global__ void nzero_kernel(const int32_t* __restrict__ counts, int64_t G,
                             unsigned long long* __restrict__ result) {
  unsigned long long best = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < G;
       i += (int64_t)gridDim.x * blockDim.x)
    if (counts[i] != 1) best = (unsigned long long)(i + 1);
  // i ascends per thread so `best` is this thread's max
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
    best = other > best ? other : best;
  }
  if ((threadIdx.x & 31) == 0 && best) atomicMax(result, best);
}

}  // namespace

int pack_connectivity(const sfem_space_desc& desc, int n,
                      const uint8_t* dirichlet, uint32_t* conn,
                      int64_t* n_zero, cudaStream_t stream) {
  const int64_t total = desc.num_elements * (int64_t)n;
  const int64_t G = desc.num_nodes;
  SFEM_REQUIRE(G < (int64_t)kConnIdMask, "num_nodes must be < 2^30 - 1");
  int32_t* counts = nullptr;
  unsigned long long* d_nz = nullptr;
  DeviceFrees guard;  // released on every return below
  SFEM_CUDA_CHECK(cudaMalloc(&counts, sizeof(int32_t) * (size_t)(G + 1)));
  guard.add(counts);
  SFEM_CUDA_CHECK(cudaMalloc(&d_nz, sizeof(unsigned long long)));
  guard.add(d_nz);
  SFEM_CUDA_CHECK(
      cudaMemsetAsync(counts, 0, sizeof(int32_t) * (size_t)(G + 1), stream));
  SFEM_CUDA_CHECK(cudaMemsetAsync(d_nz, 0, sizeof(unsigned long long), stream));
  if (total > 0) {
    count_kernel<<<blocks_for(total), kThreads, 0, stream>>>(desc.elements,
                                                            total, counts);
    SFEM_LAUNCH_CHECK();
    pack_conn_kernel<<<blocks_for(total), kThreads, 0, stream>>>(
        desc.elements, total, counts, dirichlet, conn);
    SFEM_LAUNCH_CHECK();
  }
  if (G > 0) {
    nzero_kernel<<<blocks_for(G), kThreads, 0, stream>>>(counts, G, d_nz);
    SFEM_LAUNCH_CHECK();
  }
  unsigned long long h_nz = 0;
  SFEM_CUDA_CHECK(cudaMemcpyAsync(&h_nz, d_nz, sizeof(h_nz),
                                  cudaMemcpyDeviceToHost, stream));
  SFEM_CUDA_CHECK(cudaStreamSynchronize(stream));
  *n_zero = (int64_t)h_nz;
  return SFEM_OK;
}

template <typename T>
static int gather_impl(const void* u, const int32_t* idx, int64_t count,
                       double fill, int stride, int offset, void* out,
                       cudaStream_t stream) {
  if (count == 0) return SFEM_OK;
  gather_kernel<T><<<blocks_for(count, 4), kThreads, 0, stream>>>(
      (const T*)u, idx, count, (T)fill, stride, offset, (T*)out);
  SFEM_LAUNCH_CHECK();
  return SFEM_OK;
}

template <typename T>
static int scatter_impl(const void* ul, const int32_t* idx, int64_t count,
                        int64_t num_nodes, int stride, int offset, void* out,
                        cudaStream_t stream) {
  if (stride == 1 || offset < 0) {
    SFEM_CUDA_CHECK(cudaMemsetAsync(
        out, 0, sizeof(T) * (size_t)num_nodes * (size_t)stride, stream));
  }  // one-component-per-call AoS callers zero the whole field once themselves
  if (count == 0) return SFEM_OK;
  scatter_add_kernel<T><<<blocks_for(count, 4), kThreads, 0, stream>>>(
      (const T*)ul, idx, count, stride, offset, (T*)out);
  SFEM_LAUNCH_CHECK();
  return SFEM_OK;
}

}  // namespace sfem

struct sfem_scatter_plan {
  int64_t count;
  int64_t nvalid;
  int64_t num_nodes;
  int32_t* keys;  // sorted node ids (count)
  int32_t* perm;  // matching local slots (count)
};

namespace sfem {

// y is a torch / XLA allocation (>= 256-byte aligned); the last bytes % 16 are
// written one by one.
int launch_zero_fill(void* y, size_t bytes, double* dot_xy,
                     cudaStream_t stream, bool* used_kernel) {
  *used_kernel = ((uintptr_t)y & 15u) == 0;
  if (!*used_kernel) {  // unaligned views: plain memsets
    if (bytes) SFEM_CUDA_CHECK(cudaMemsetAsync(y, 0, bytes, stream));
    if (dot_xy)
      SFEM_CUDA_CHECK(cudaMemsetAsync(dot_xy, 0, sizeof(double), stream));
    return SFEM_OK;
  }
  const size_t n16 = bytes / 16;
  zero_fill_kernel<<<num_sms(), 256, 0, stream>>>(
      (uint4*)y, n16, (unsigned char*)y + n16 * 16, (int)(bytes - n16 * 16),
      dot_xy);
  SFEM_LAUNCH_CHECK();
  return SFEM_OK;
}

}  // namespace sfem

extern "C" {

int sfem_pointwise(int dtype, int32_t kind, int32_t dim, const void* a,
                   const void* g, int64_t num_points, void* out,
                   sfem_stream_t stream_) {
  using namespace sfem;
  cudaStream_t stream = (cudaStream_t)stream_;
  SFEM_REQUIRE(kind >= 0 && kind <= 2, "kind must be 0, 1 or 2");
  SFEM_REQUIRE(dim >= 1 && dim <= 3, "dim must be 1, 2 or 3");
  SFEM_REQUIRE(out && (kind == 1 ? a != nullptr : g != nullptr) &&
                   (kind != 2 || a != nullptr),
               "null argument");
  if (num_points == 0) return SFEM_OK;
  if (dtype == SFEM_F64)
    pointwise_kernel<double><<<blocks_for(num_points), kThreads, 0, stream>>>(
        kind, dim, (const double*)a, (const double*)g, num_points,
        (double*)out);
  else
    pointwise_kernel<float><<<blocks_for(num_points), kThreads, 0, stream>>>(
        kind, dim, (const float*)a, (const float*)g, num_points, (float*)out);
  SFEM_LAUNCH_CHECK();
  return SFEM_OK;
}

int sfem_gather(int dtype, const void* u, const int32_t* indices, int64_t count,
                double fill_value, int32_t stride, int32_t offset, void* out,
                sfem_stream_t stream) {
  SFEM_REQUIRE(count >= 0 && stride >= 1 && offset >= -1 && offset < stride,
               "sfem_gather: bad count/stride/offset");
  return dtype == SFEM_F64
             ? sfem::gather_impl<double>(u, indices, count, fill_value, stride,
                                         offset, out, (cudaStream_t)stream)
             : sfem::gather_impl<float>(u, indices, count, fill_value, stride,
                                        offset, out, (cudaStream_t)stream);
}

int sfem_scatter_add(int dtype, const void* u_local, const int32_t* indices,
                     int64_t count, int64_t num_nodes, int32_t stride,
                     int32_t offset, void* out, sfem_stream_t stream) {
  SFEM_REQUIRE(count >= 0 && stride >= 1 && offset >= -1 && offset < stride,
               "sfem_scatter_add: bad count/stride/offset");
  return dtype == SFEM_F64
             ? sfem::scatter_impl<double>(u_local, indices, count, num_nodes,
                                          stride, offset, out,
                                          (cudaStream_t)stream)
             : sfem::scatter_impl<float>(u_local, indices, count, num_nodes,
                                         stride, offset, out,
                                         (cudaStream_t)stream);
}

int sfem_scatter_plan_create(const int32_t* indices, int64_t count,
                             int64_t num_nodes, sfem_scatter_plan** plan,
                             sfem_stream_t stream_) {
  using namespace sfem;
  cudaStream_t stream = (cudaStream_t)stream_;
  SFEM_REQUIRE(count >= 0 && count < (int64_t)0x7fffffff,
               "sfem_scatter_plan_create: count must fit int32");
  SFEM_REQUIRE(num_nodes < (int64_t)0x7fffffff, "num_nodes must fit int32");
  auto* p = new sfem_scatter_plan{count, 0, num_nodes, nullptr, nullptr};
  if (count == 0) {
    *plan = p;
    return SFEM_OK;
  }
  // the plan itself is released on every early (error) return
  struct PlanGuard {
    sfem_scatter_plan* p;
    ~PlanGuard() { sfem_scatter_plan_destroy(p); }
  } plan_guard{p};
  DeviceFrees guard;  // temporaries
  int32_t *keys_in = nullptr, *vals_in = nullptr;
  void* tmp = nullptr;
  size_t tmp_bytes = 0;
  SFEM_CUDA_CHECK(cudaMalloc(&keys_in, sizeof(int32_t) * count));
  guard.add(keys_in);
  SFEM_CUDA_CHECK(cudaMalloc(&vals_in, sizeof(int32_t) * count));
  guard.add(vals_in);
  SFEM_CUDA_CHECK(cudaMalloc(&p->keys, sizeof(int32_t) * count));
  SFEM_CUDA_CHECK(cudaMalloc(&p->perm, sizeof(int32_t) * count));
  remap_sentinel_kernel<<<blocks_for(count), kThreads, 0, stream>>>(
      indices, keys_in, count, (int32_t)num_nodes);
  SFEM_LAUNCH_CHECK();
  iota_kernel<<<blocks_for(count), kThreads, 0, stream>>>(vals_in, count);
  SFEM_LAUNCH_CHECK();
  int end_bit = 1;
  while (((int64_t)1 << end_bit) <= num_nodes && end_bit < 31) ++end_bit;
  SFEM_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(
      nullptr, tmp_bytes, keys_in, p->keys, vals_in, p->perm, (int)count, 0,
      end_bit, stream));
  SFEM_CUDA_CHECK(cudaMalloc(&tmp, tmp_bytes));
  guard.add(tmp);
  // LSD radix sort is stable: equal keys keep ascending slot order.
  SFEM_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(
      tmp, tmp_bytes, keys_in, p->keys, vals_in, p->perm, (int)count, 0,
      end_bit, stream));
  // number of non-sentinel entries = first position with key == num_nodes
  // (binary search on the host over a device array would sync per probe; do a
  // tiny reduction instead: count sentinels)
  int32_t* h_keys_tail = nullptr;
  (void)h_keys_tail;
  SFEM_CUDA_CHECK(cudaStreamSynchronize(stream));
  // binary search with few D2H probes (log2(count) <= 31)
  int64_t lo = 0, hi = count;
  while (lo < hi) {
    const int64_t mid = (lo + hi) / 2;
    int32_t k;
    SFEM_CUDA_CHECK(cudaMemcpy(&k, p->keys + mid, sizeof(k),
                               cudaMemcpyDeviceToHost));
    if (k >= (int32_t)num_nodes)
      hi = mid;
    else
      lo = mid + 1;
  }
  p->nvalid = lo;
  plan_guard.p = nullptr;  // success: ownership passes to the caller
  *plan = p;
  return SFEM_OK;
}

int sfem_scatter_plan_apply(const sfem_scatter_plan* plan, int dtype,
                            const void* u_local, int32_t stride, int32_t offset,
                            void* out, sfem_stream_t stream_) {
  using namespace sfem;
  cudaStream_t stream = (cudaStream_t)stream_;
  SFEM_REQUIRE(plan != nullptr, "null plan");
  const size_t esz = dtype == SFEM_F64 ? 8 : 4;
  if (stride == 1)
    SFEM_CUDA_CHECK(
        cudaMemsetAsync(out, 0, esz * (size_t)plan->num_nodes, stream));
  if (plan->nvalid == 0) return SFEM_OK;
  const int blocks = blocks_for(plan->nvalid);
  if (dtype == SFEM_F64)
    scatter_segmented_kernel<double><<<blocks, kThreads, 0, stream>>>(
        (const double*)u_local, plan->keys, plan->perm, plan->nvalid, stride,
        offset, (double*)out);
  else
    scatter_segmented_kernel<float><<<blocks, kThreads, 0, stream>>>(
        (const float*)u_local, plan->keys, plan->perm, plan->nvalid, stride,
        offset, (float*)out);
  SFEM_LAUNCH_CHECK();
  return SFEM_OK;
}

void sfem_scatter_plan_destroy(sfem_scatter_plan* plan) {
  if (!plan) return;
  cudaFree(plan->keys);
  cudaFree(plan->perm);
  delete plan;
}

int sfem_exchange(int dtype, void* u, const int32_t* gather_indices,
                  const int32_t* unique_indices, int64_t count,
                  int64_t num_unique, int32_t stride, int32_t offset,
                  void* scratch, sfem_stream_t stream_) {
  using namespace sfem;
  cudaStream_t stream = (cudaStream_t)stream_;
  if (count == 0) return SFEM_OK;
  SFEM_REQUIRE(stride >= 1 && offset >= -1 && offset < stride,
               "sfem_exchange: bad stride/offset");
  const size_t esz = dtype == SFEM_F64 ? 8 : 4;
  SFEM_CUDA_CHECK(cudaMemsetAsync(
      scratch, 0, esz * (size_t)num_unique * (size_t)(offset < 0 ? stride : 1),
      stream));
  const int blocks = blocks_for(count);
  if (dtype == SFEM_F64) {
    exchange_sum_kernel<double><<<blocks, kThreads, 0, stream>>>(
        (const double*)u, gather_indices, unique_indices, count, stride, offset,
        (double*)scratch);
    SFEM_LAUNCH_CHECK();
    exchange_write_kernel<double><<<blocks, kThreads, 0, stream>>>(
        (double*)u, gather_indices, unique_indices, count, stride, offset,
        (const double*)scratch);
  } else {
    exchange_sum_kernel<float><<<blocks, kThreads, 0, stream>>>(
        (const float*)u, gather_indices, unique_indices, count, stride, offset,
        (float*)scratch);
    SFEM_LAUNCH_CHECK();
    exchange_write_kernel<float><<<blocks, kThreads, 0, stream>>>(
        (float*)u, gather_indices, unique_indices, count, stride, offset,
        (const float*)scratch);
  }
  SFEM_LAUNCH_CHECK();
  return SFEM_OK;
}

int sfem_halo_pack(int dtype, const void* u, const int32_t* idx, int64_t count,
                   void* buf, sfem_stream_t stream_) {
  using namespace sfem;
  cudaStream_t stream = (cudaStream_t)stream_;
  if (count == 0) return SFEM_OK;
  if (dtype == SFEM_F64)
    pack_kernel<double><<<blocks_for(count), kThreads, 0, stream>>>(
        (const double*)u, idx, count, (double*)buf);
  else
    pack_kernel<float><<<blocks_for(count), kThreads, 0, stream>>>(
        (const float*)u, idx, count, (float*)buf);
  SFEM_LAUNCH_CHECK();
  return SFEM_OK;
}

int sfem_halo_unpack_canonical(int dtype, void* u, const int32_t* dofs,
                               const int32_t* row_ptr, const int32_t* src,
                               int64_t num_dofs, const void* recv,
                               sfem_stream_t stream_) {
  using namespace sfem;
  cudaStream_t stream = (cudaStream_t)stream_;
  if (num_dofs == 0) return SFEM_OK;
  SFEM_REQUIRE(u && dofs && row_ptr && src, "null argument");
  if (dtype == SFEM_F64)
    unpack_canonical_kernel<double><<<blocks_for(num_dofs), kThreads, 0,
                                      stream>>>(
        (double*)u, dofs, row_ptr, src, num_dofs, (const double*)recv);
  else
    unpack_canonical_kernel<float><<<blocks_for(num_dofs), kThreads, 0,
                                     stream>>>(
        (float*)u, dofs, row_ptr, src, num_dofs, (const float*)recv);
  SFEM_LAUNCH_CHECK();
  return SFEM_OK;
}

int sfem_halo_unpack_add(int dtype, void* u, const int32_t* idx, int64_t count,
                         const void* buf, sfem_stream_t stream_) {
  using namespace sfem;
  cudaStream_t stream = (cudaStream_t)stream_;
  if (count == 0) return SFEM_OK;
  if (dtype == SFEM_F64)
    unpack_add_kernel<double><<<blocks_for(count), kThreads, 0, stream>>>(
        (double*)u, idx, count, (const double*)buf);
  else
    unpack_add_kernel<float><<<blocks_for(count), kThreads, 0, stream>>>(
        (float*)u, idx, count, (const float*)buf);
  SFEM_LAUNCH_CHECK();
  return SFEM_OK;
}

}  // extern "C"
