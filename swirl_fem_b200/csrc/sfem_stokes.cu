// Fused Stokes divergence D and its transpose D^T
// (swirl_fem/navier_stokes/navier_stokes.py:313-338): the discrete divergence
// of a continuous GLL velocity tested with the discontinuous GL pressure basis,
// both integrated with the velocity's collocated GLL rule.
//
//   D(u)[m]    = sum_q  psi_m(q) W_q detJ_q  div u (q)
//   D^T(p)[n,k] = mask_n sum_q p(q) W_q detJ_q  d phi_n / d x_k (q)
//
// The reference builds both from `local_covector` of the form
// `div(v)(x) * q(x)` (a `jax.linear_transpose`); the composed path here was
// gather -> evaluation -> pointwise -> transposed evaluation -> scatter, five
// launches of the generic kernels (one CTA per element and component, several
// block barriers per contraction) plus per-component gathers: 100-135 us per
// operator at 64 x 64 elements of order 7, 4 % of the HBM roofline at 256^2.
// These kernels do each operator in ONE launch: one WARP per element (its own
// slice of shared memory, __syncwarp only), gather, sum-factorised gradient /
// interpolation, the pointwise trace with the inverse Jacobian, the transposed
// sum-factorised contraction and the scatter; the 1-D tables are loaded once
// per CTA.  Runtime (dim, N, Np): N^dim <= 1024.

#include "sfem_common.cuh"

namespace sfem {
namespace {

constexpr int kStokesMaxWarps = 4;

struct StokesShape {
  int dim, N, Np;  // velocity nodes = quadrature points per axis; pressure nodes
  int n, np;       // N^dim, Np^dim
};

constexpr int cpow_s(int b, int e) { return e <= 0 ? 1 : b * cpow_s(b, e - 1); }

// Compile-time shape (DIM_ > 0) or the runtime one: with constants the
// contractions unroll, the index arithmetic reduces to shifts / constant
// multiplies and the 1-D table rows stay in registers.
template <int DIM_, int N_, int NP_>
__device__ __forceinline__ StokesShape fix_shape(const StokesShape& s) {
  if constexpr (DIM_ > 0) {
    StokesShape f;
    f.dim = DIM_;
    f.N = N_;
    f.Np = NP_;
    f.n = cpow_s(N_, DIM_);
    f.np = cpow_s(NP_, DIM_);
    return f;
  } else {
    return s;
  }
}

__device__ __forceinline__ int ipow_s(int b, int e) {
  int r = 1;
  for (int i = 0; i < e; ++i) r *= b;
  return r;
}

// Warp-level 1-D contraction over one axis of a tensor in shared memory:
// out[a][o][c] = sum_i M[o * so + i * si] * in[a][i][c]
template <typename T>
__device__ __forceinline__ void contract_w(const T* __restrict__ M, int so,
                                           int si, const T* __restrict__ in,
                                           T* __restrict__ out, int A, int I,
                                           int O, int C) {
  const int total = A * O * C;
  for (int idx = threadIdx.x & 31; idx < total; idx += 32) {
    const int c = idx % C;
    const int o = (idx / C) % O;
    const int a = idx / (C * O);
    const T* ip = in + a * I * C + c;
    T acc = T(0);
#pragma unroll
    for (int i = 0; i < I; ++i) acc += M[o * so + i * si] * ip[i * C];
    out[idx] = acc;
  }
  __syncwarp();
}

template <typename T>
__device__ __forceinline__ T quad_w(const StokesShape& s, const T* W, int q) {
  T w = T(1);
  for (int a = 0; a < s.dim; ++a) {
    w *= W[q % s.N];
    q /= s.N;
  }
  return w;
}

// shared layout per CTA: [Dv (N*N) | Bp (N*Np) | W (N)] then per-warp slices
template <typename T>
__device__ __forceinline__ T* load_stokes_tables(const StokesShape& s,
                                                 const T* __restrict__ vtab,
                                                 const T* __restrict__ ptab,
                                                 T* smem) {
  // vtab = [B | BD | W] of the velocity space (Q = N): BD = D, W
  // ptab = [B | BD | W] of the pressure space (Q = N, N = Np): B = Bp
  T* Dv = smem;
  T* Bp = Dv + s.N * s.N;
  T* W = Bp + s.N * s.Np;
  for (int i = threadIdx.x; i < s.N * s.N; i += blockDim.x)
    Dv[i] = vtab[s.N * s.N + i];
  for (int i = threadIdx.x; i < s.N * s.Np; i += blockDim.x) Bp[i] = ptab[i];
  for (int i = threadIdx.x; i < s.N; i += blockDim.x)
    W[i] = vtab[2 * s.N * s.N + i];
  __syncthreads();
  return W + s.N;
}

template <typename T, int DIM_ = 0, int N_ = 0, int NP_ = 0>
__global__ void __launch_bounds__(32 * kStokesMaxWarps)
stokes_div_kernel(StokesShape s_rt, const T* __restrict__ vtab,
                  const T* __restrict__ ptab,
                  const int32_t* __restrict__ v_el,
                  const int32_t* __restrict__ p_el,
                  const T* __restrict__ invjacs, const T* __restrict__ jacdets,
                  const T* __restrict__ u, int64_t E, int slice,
                  T* __restrict__ out) {
  const StokesShape s = fix_shape<DIM_, N_, NP_>(s_rt);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* smem = reinterpret_cast<T*>(smem_raw);
  const T* Dv = smem;
  const T* Bp = Dv + s.N * s.N;
  const T* W = Bp + s.N * s.Np;
  T* mine = load_stokes_tables<T>(s, vtab, ptab, smem) +
            (size_t)(threadIdx.x >> 5) * slice;
  const int d = s.dim, lane = threadIdx.x & 31;
  T* U = mine;            // (d, n)
  T* H = U + d * s.n;     // (n) weighted divergence, then scratch
  T* t0 = H + s.n;        // (n)
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  // fixed shapes: the connectivity of the NEXT element is fetched while this
  // one is computed (one dependent global round trip less per element)
  constexpr bool FIXED = DIM_ > 0;
  constexpr int PASSES = FIXED ? (cpow_s(N_, DIM_) + 31) / 32 : 1;
  int32_t gnext[PASSES];
  auto fetch_conn = [&](int64_t e2) {
#pragma unroll
    for (int k = 0; k < PASSES; ++k) {
      const int i = lane + 32 * k;
      gnext[k] = (e2 < E && i < s.n) ? __ldg(v_el + e2 * s.n + i) : SFEM_SENTINEL;
    }
  };
  if (FIXED) fetch_conn(warp);
  for (int64_t e = warp; e < E; e += nwarps) {
    if constexpr (FIXED) {
#pragma unroll
      for (int k = 0; k < PASSES; ++k) {
        const int i = lane + 32 * k;
        const int32_t g = gnext[k];
        if (i < s.n) {
#pragma unroll
          for (int j = 0; j < DIM_; ++j)
            U[j * s.n + i] = g == SFEM_SENTINEL ? T(0) : u[(int64_t)g * d + j];
        }
      }
      fetch_conn(e + nwarps);
    } else {
      for (int i = lane; i < s.n; i += 32) {
        const int32_t g = v_el[e * s.n + i];
        for (int j = 0; j < d; ++j)
          U[j * s.n + i] = g == SFEM_SENTINEL ? T(0) : u[(int64_t)g * d + j];
      }
    }
    __syncwarp();
#pragma unroll
    for (int q = lane; q < s.n; q += 32) {
      const int64_t eq = e * s.n + q;
      const T* inv = invjacs + eq * d * d;
      T div = T(0);
      int rem = q, stride = 1;
      // axis a = d-1 (fastest) ... 0; stride of axis a = N^(d-1-a)
      for (int a = d - 1; a >= 0; --a) {
        const int qa = rem % s.N;
        rem /= s.N;
        const int base = q - qa * stride;
        for (int j = 0; j < d; ++j) {
          T g = T(0);
          const T* uj = U + j * s.n + base;
#pragma unroll
          for (int m = 0; m < s.N; ++m) g += Dv[qa * s.N + m] * uj[m * stride];
          div += g * inv[j * d + a];   // grad_j(u_j) = sum_a g_a(u_j) Jinv[j][a]
        }
        stride *= s.N;
      }
      H[q] = quad_w<T>(s, W, q) * jacdets[eq] * div;
    }
    __syncwarp();
    // y[m] = sum_q Bp^{(x)}[q, m] H[q]: axis by axis, Q = N -> Np
    const T* in = H;
    T* bufs[2] = {t0, H};
    int which = 0;
    for (int axis = 0; axis < d; ++axis) {
      const int A = ipow_s(s.Np, axis);
      const int C = ipow_s(s.N, d - 1 - axis);
      T* o = bufs[which];
      contract_w<T>(Bp, 1, s.Np, in, o, A, s.N, s.Np, C);
      in = o;
      which ^= 1;
    }
    for (int m = lane; m < s.np; m += 32) {
      const int32_t g = p_el[e * s.np + m];
      if (g != SFEM_SENTINEL) red_add(out + g, in[m]);
    }
    __syncwarp();
  }
}

template <typename T, int DIM_ = 0, int N_ = 0, int NP_ = 0>
__global__ void __launch_bounds__(32 * kStokesMaxWarps)
stokes_grad_t_kernel(StokesShape s_rt, const T* __restrict__ vtab,
                     const T* __restrict__ ptab,
                     const int32_t* __restrict__ v_el,
                     const int32_t* __restrict__ p_el,
                     const T* __restrict__ invjacs,
                     const T* __restrict__ jacdets, const T* __restrict__ p,
                     const T* __restrict__ mask, int64_t E, int slice,
                     T* __restrict__ out) {
  const StokesShape s = fix_shape<DIM_, N_, NP_>(s_rt);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* smem = reinterpret_cast<T*>(smem_raw);
  const T* Dv = smem;
  const T* Bp = Dv + s.N * s.N;
  const T* W = Bp + s.N * s.Np;
  T* mine = load_stokes_tables<T>(s, vtab, ptab, smem) +
            (size_t)(threadIdx.x >> 5) * slice;
  const int d = s.dim, lane = threadIdx.x & 31;
  T* P = mine;             // (n) pressure nodes, then values at the points
  T* t0 = P + s.n;         // (n)
  T* F = t0 + s.n;         // (d * d, n): F[a][k][q] = c_q Jinv[k][a]
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  constexpr bool FIXED = DIM_ > 0;
  constexpr int PPASSES = FIXED ? (cpow_s(NP_, DIM_) + 31) / 32 : 1;
  int32_t gnext[PPASSES];
  auto fetch_conn = [&](int64_t e2) {
#pragma unroll
    for (int k = 0; k < PPASSES; ++k) {
      const int m = lane + 32 * k;
      gnext[k] = (e2 < E && m < s.np) ? __ldg(p_el + e2 * s.np + m) : SFEM_SENTINEL;
    }
  };
  if (FIXED) fetch_conn(warp);
  for (int64_t e = warp; e < E; e += nwarps) {
    if constexpr (FIXED) {
#pragma unroll
      for (int k = 0; k < PPASSES; ++k) {
        const int m = lane + 32 * k;
        const int32_t g = gnext[k];
        if (m < s.np) P[m] = g == SFEM_SENTINEL ? T(0) : p[g];
      }
      fetch_conn(e + nwarps);
    } else {
      for (int m = lane; m < s.np; m += 32) {
        const int32_t g = p_el[e * s.np + m];
        P[m] = g == SFEM_SENTINEL ? T(0) : p[g];
      }
    }
    __syncwarp();
    // p(q) = sum_m Bp^{(x)}[q, m] P[m]: last axis first, Np -> N
    const T* in = P;
    T* bufs[2] = {t0, P};
    int which = 0;
    for (int axis = d - 1; axis >= 0; --axis) {
      const int A = ipow_s(s.Np, axis);
      const int C = ipow_s(s.N, d - 1 - axis);
      T* o = bufs[which];
      contract_w<T>(Bp, s.Np, 1, in, o, A, s.Np, s.N, C);
      in = o;
      which ^= 1;
    }
#pragma unroll
    for (int q = lane; q < s.n; q += 32) {
      const int64_t eq = e * s.n + q;
      const T c = quad_w<T>(s, W, q) * jacdets[eq] * in[q];
      const T* inv = invjacs + eq * d * d;
      for (int a = 0; a < d; ++a)
        for (int k = 0; k < d; ++k)
          F[(a * d + k) * s.n + q] = c * inv[k * d + a];
    }
    __syncwarp();
    // y_k[n] = sum_a sum_m D[m][n_a] F[a][k][n with n_a -> m]
    for (int i = lane; i < s.n; i += 32) {
      const int32_t g = v_el[e * s.n + i];
      T y[3] = {T(0), T(0), T(0)};
      int rem = i, stride = 1;
      for (int a = d - 1; a >= 0; --a) {
        const int na = rem % s.N;
        rem /= s.N;
        const int base = i - na * stride;
        for (int k = 0; k < d; ++k) {
          const T* f = F + (a * d + k) * s.n + base;
          T acc = T(0);
#pragma unroll
          for (int m = 0; m < s.N; ++m) acc += Dv[m * s.N + na] * f[m * stride];
          y[k] += acc;
        }
        stride *= s.N;
      }
      if (g != SFEM_SENTINEL) {
        const T w = mask ? mask[g] : T(1);
        for (int k = 0; k < d; ++k) red_add(out + (int64_t)g * d + k, w * y[k]);
      }
    }
    __syncwarp();
  }
}


// ---- 2-D, compile-time (N, NP): one lane per LINE ------------------------------
// ncu on the kernels above (profiles/r02_ncu_stokes_div_ne256.txt): shared-
// memory bound -- 401 / 278 wavefronts per element, MIO-throttle and short-
// scoreboard stalls of 9 cycles per issue -- because every POINT re-loads its
// 8 line values and 8 matrix entries for each derivative.  Here a lane owns a
// whole line (row or column) of one field: N loads feed N outputs (N^2 FMAs
// from registers), the 1-D matrices are kernel parameters (constant bank: FMA
// operands, no load instruction), tiles have a row pitch of N + 1 so that rows
// and columns are both conflict-free, and the connectivity of the next
// element is prefetched.  ~95 wavefronts per element.
template <typename T, int N, int NP>
struct StokesTabs {
  T Dv[N * N];   // Dv[q * N + m]  = d phi_m / d xi (xi_q)
  T Bp[N * NP];  // Bp[q * NP + m] = psi_m (xi_q)
  T W[N];
};

template <typename T, int N, int NP>
StokesTabs<T, N, NP> make_stokes_tabs(const sfem_space* v, const sfem_space* p) {
  StokesTabs<T, N, NP> tb;
  for (int i = 0; i < N * N; ++i) tb.Dv[i] = (T)v->base.h_BD[i];
  for (int i = 0; i < N * NP; ++i) tb.Bp[i] = (T)p->base.h_B[i];
  for (int i = 0; i < N; ++i) tb.W[i] = (T)v->base.h_W[i];
  return tb;
}

template <typename T, int N, int NP>
__global__ void __launch_bounds__(32 * kStokesMaxWarps)
stokes_div2d_kernel(const __grid_constant__ StokesTabs<T, N, NP> tb,
                    const int32_t* __restrict__ v_el,
                    const int32_t* __restrict__ p_el,
                    const T* __restrict__ invjacs,
                    const T* __restrict__ jacdets, const T* __restrict__ u,
                    int64_t E, T* __restrict__ out) {
  constexpr int n = N * N, np = NP * NP, R = N + 1, tile = N * R;
  constexpr int PASSES = (n + 31) / 32, PPASSES = (np + 31) / 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  T* mine = reinterpret_cast<T*>(smem_raw) + (size_t)(threadIdx.x >> 5) * (6 * tile);
  T* U = mine;             // [2][tile]; U[0] becomes H, U[1] the half-contracted T1
  T* G = mine + 2 * tile;  // [j * 2 + a][tile]
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  int32_t gnext[PASSES], pnext[PPASSES];
  auto fetch_conn = [&](int64_t e2) {
#pragma unroll
    for (int k = 0; k < PASSES; ++k) {
      const int i = lane + 32 * k;
      gnext[k] = (e2 < E && i < n) ? __ldg(v_el + e2 * n + i) : SFEM_SENTINEL;
    }
#pragma unroll
    for (int k = 0; k < PPASSES; ++k) {
      const int m = lane + 32 * k;
      pnext[k] = (e2 < E && m < np) ? __ldg(p_el + e2 * np + m) : SFEM_SENTINEL;
    }
  };
  fetch_conn(warp);
  for (int64_t e = warp; e < E; e += nwarps) {
    int32_t pg[PPASSES];
#pragma unroll
    for (int k = 0; k < PPASSES; ++k) pg[k] = pnext[k];
#pragma unroll
    for (int k = 0; k < PASSES; ++k) {
      const int i = lane + 32 * k;
      if (i < n) {
        const int32_t g = gnext[k];
        const int at = (i / N) * R + (i % N);
        U[at] = g == SFEM_SENTINEL ? T(0) : u[(int64_t)g * 2];
        U[tile + at] = g == SFEM_SENTINEL ? T(0) : u[(int64_t)g * 2 + 1];
      }
    }
    fetch_conn(e + nwarps);
    // the geometric data of this element's points (used after the derivatives)
    T inv[PASSES][4], jd[PASSES];
#pragma unroll
    for (int k = 0; k < PASSES; ++k) {
      const int q = lane + 32 * k;
      if (q < n) {
        const int64_t eq = e * n + q;
#pragma unroll
        for (int c = 0; c < 4; ++c) inv[k][c] = invjacs[eq * 4 + c];
        jd[k] = jacdets[eq];
      }
    }
    __syncwarp();
    // line derivatives: task (a, j, l); a = 1: row l (fastest index), a = 0:
    // column l.  (a is the slowest task index: for N = 8 a half-warp then reads
    // 16 columns -- or 16 rows -- of two tiles 8 bank pairs apart: no conflicts.)
    for (int t = lane; t < 4 * N; t += 32) {
      const int l = t % N, j = (t / N) & 1, a = t / (2 * N);
      const int st = a ? 1 : R;
      const int first = a ? l * R : l;
      const T* src = U + j * tile + first;
      T v[N];
#pragma unroll
      for (int m = 0; m < N; ++m) v[m] = src[m * st];
      T* dst = G + (j * 2 + a) * tile + first;
#pragma unroll
      for (int o = 0; o < N; ++o) {
        T acc = T(0);
#pragma unroll
        for (int m = 0; m < N; ++m) acc += tb.Dv[o * N + m] * v[m];
        dst[o * st] = acc;
      }
    }
    __syncwarp();
    // weighted divergence at the points -> H (= U[0])
#pragma unroll
    for (int k = 0; k < PASSES; ++k) {
      const int q = lane + 32 * k;
      if (q < n) {
        const int r = q / N, c = q % N, at = r * R + c;
        T div = T(0);
#pragma unroll
        for (int ja = 0; ja < 4; ++ja) div += G[ja * tile + at] * inv[k][ja];
        U[at] = tb.W[r] * tb.W[c] * jd[k] * div;
      }
    }
    __syncwarp();
    // pressure basis, axis 0: one lane per column c: T1[m0][c] = sum_q0 Bp[q0][m0] H[q0][c]
    if (lane < N) {
      T h[N];
#pragma unroll
      for (int q0 = 0; q0 < N; ++q0) h[q0] = U[q0 * R + lane];
#pragma unroll
      for (int m0 = 0; m0 < NP; ++m0) {
        T acc = T(0);
#pragma unroll
        for (int q0 = 0; q0 < N; ++q0) acc += tb.Bp[q0 * NP + m0] * h[q0];
        U[tile + m0 * R + lane] = acc;
      }
    }
    __syncwarp();
    // axis 1: one lane per row m0: y[m0][m1] = sum_q1 Bp[q1][m1] T1[m0][q1]  -> G[0] (contiguous m)
    if (lane < NP) {
      T h[N];
#pragma unroll
      for (int q1 = 0; q1 < N; ++q1) h[q1] = U[tile + lane * R + q1];
#pragma unroll
      for (int m1 = 0; m1 < NP; ++m1) {
        T acc = T(0);
#pragma unroll
        for (int q1 = 0; q1 < N; ++q1) acc += tb.Bp[q1 * NP + m1] * h[q1];
        G[lane * NP + m1] = acc;
      }
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < PPASSES; ++k) {
      const int m = lane + 32 * k;
      if (m < np && pg[k] != SFEM_SENTINEL) red_add(out + pg[k], G[m]);
    }
    __syncwarp();
  }
}

template <typename T, int N, int NP>
__global__ void __launch_bounds__(32 * kStokesMaxWarps)
stokes_grad_t2d_kernel(const __grid_constant__ StokesTabs<T, N, NP> tb,
                       const int32_t* __restrict__ v_el,
                       const int32_t* __restrict__ p_el,
                       const T* __restrict__ invjacs,
                       const T* __restrict__ jacdets, const T* __restrict__ p,
                       const T* __restrict__ mask, int64_t E,
                       T* __restrict__ out) {
  constexpr int n = N * N, np = NP * NP, R = N + 1, tile = N * R;
  constexpr int PASSES = (n + 31) / 32, PPASSES = (np + 31) / 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  T* mine = reinterpret_cast<T*>(smem_raw) + (size_t)(threadIdx.x >> 5) * (6 * tile);
  T* P = mine;             // [NP][R] pressure nodes, later V[N][R] values at the points
  T* T1 = mine + tile;     // [NP][R]
  T* F = mine + 2 * tile;  // [a * 2 + k][tile]
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  int32_t gnext[PASSES], pnext[PPASSES];
  auto fetch_conn = [&](int64_t e2) {
#pragma unroll
    for (int k = 0; k < PASSES; ++k) {
      const int i = lane + 32 * k;
      gnext[k] = (e2 < E && i < n) ? __ldg(v_el + e2 * n + i) : SFEM_SENTINEL;
    }
#pragma unroll
    for (int k = 0; k < PPASSES; ++k) {
      const int m = lane + 32 * k;
      pnext[k] = (e2 < E && m < np) ? __ldg(p_el + e2 * np + m) : SFEM_SENTINEL;
    }
  };
  fetch_conn(warp);
  for (int64_t e = warp; e < E; e += nwarps) {
    int32_t vg[PASSES];
#pragma unroll
    for (int k = 0; k < PASSES; ++k) vg[k] = gnext[k];
#pragma unroll
    for (int k = 0; k < PPASSES; ++k) {
      const int m = lane + 32 * k;
      if (m < np) {
        const int32_t g = pnext[k];
        P[(m / NP) * R + (m % NP)] = g == SFEM_SENTINEL ? T(0) : p[g];
      }
    }
    fetch_conn(e + nwarps);
    T inv[PASSES][4], jd[PASSES], w[PASSES];
#pragma unroll
    for (int k = 0; k < PASSES; ++k) {
      const int q = lane + 32 * k;
      w[k] = T(0);
      if (q < n) {
        const int64_t eq = e * n + q;
#pragma unroll
        for (int c = 0; c < 4; ++c) inv[k][c] = invjacs[eq * 4 + c];
        jd[k] = jacdets[eq];
        if (vg[k] != SFEM_SENTINEL) w[k] = mask ? mask[vg[k]] : T(1);
      }
    }
    __syncwarp();
    // interpolation, axis 1: one lane per row m0: T1[m0][q1] = sum_m1 Bp[q1][m1] P[m0][m1]
    if (lane < NP) {
      T h[NP];
#pragma unroll
      for (int m1 = 0; m1 < NP; ++m1) h[m1] = P[lane * R + m1];
#pragma unroll
      for (int q1 = 0; q1 < N; ++q1) {
        T acc = T(0);
#pragma unroll
        for (int m1 = 0; m1 < NP; ++m1) acc += tb.Bp[q1 * NP + m1] * h[m1];
        T1[lane * R + q1] = acc;
      }
    }
    __syncwarp();
    // axis 0: one lane per column q1: V[q0][q1] = sum_m0 Bp[q0][m0] T1[m0][q1]  -> P
    if (lane < N) {
      T h[NP];
#pragma unroll
      for (int m0 = 0; m0 < NP; ++m0) h[m0] = T1[m0 * R + lane];
#pragma unroll
      for (int q0 = 0; q0 < N; ++q0) {
        T acc = T(0);
#pragma unroll
        for (int m0 = 0; m0 < NP; ++m0) acc += tb.Bp[q0 * NP + m0] * h[m0];
        P[q0 * R + lane] = acc;
      }
    }
    __syncwarp();
    // F[a][k][q] = W_q detJ_q p(q) Jinv[k][a]
#pragma unroll
    for (int kk = 0; kk < PASSES; ++kk) {
      const int q = lane + 32 * kk;
      if (q < n) {
        const int r = q / N, c = q % N, at = r * R + c;
        const T cq = tb.W[r] * tb.W[c] * jd[kk] * P[at];
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
          for (int k = 0; k < 2; ++k)
            F[(a * 2 + k) * tile + at] = cq * inv[kk][k * 2 + a];
      }
    }
    __syncwarp();
    // transposed derivative along every line, in place: task (a, k, l)
    for (int t = lane; t < 4 * N; t += 32) {
      const int l = t % N, k = (t / N) & 1, a = t / (2 * N);
      const int st = a ? 1 : R;
      T* line = F + (a * 2 + k) * tile + (a ? l * R : l);
      T f[N];
#pragma unroll
      for (int m = 0; m < N; ++m) f[m] = line[m * st];
#pragma unroll
      for (int o = 0; o < N; ++o) {
        T acc = T(0);
#pragma unroll
        for (int m = 0; m < N; ++m) acc += tb.Dv[m * N + o] * f[m];
        line[o * st] = acc;
      }
    }
    __syncwarp();
#pragma unroll
    for (int kk = 0; kk < PASSES; ++kk) {
      const int i = lane + 32 * kk;
      if (i < n && vg[kk] != SFEM_SENTINEL) {
        const int at = (i / N) * R + (i % N);
#pragma unroll
        for (int k = 0; k < 2; ++k)
          red_add(out + (int64_t)vg[kk] * 2 + k,
                  w[kk] * (F[k * tile + at] + F[(2 + k) * tile + at]));
      }
    }
    __syncwarp();
  }
}

// SFEM_STOKES_LINES=0 (developer switch): the point-per-lane kernels instead.
inline bool stokes_lines_enabled() {
  static const bool on = [] {
    const char* e = getenv("SFEM_STOKES_LINES");
    return !(e && e[0] == '0');
  }();
  return on;
}

template <typename T>
const T* tables_of(const sfem_space* sp);
template <>
const double* tables_of<double>(const sfem_space* sp) {
  return sp->base.d_tables64;
}
template <>
const float* tables_of<float>(const sfem_space* sp) {
  return sp->base.d_tables32;
}

int check_pair(const sfem_space* v, const sfem_space* p, StokesShape* s) {
  SFEM_REQUIRE(v && p, "null space");
  const sfem_space_desc& dv = v->base.desc;
  const sfem_space_desc& dp = p->base.desc;
  SFEM_REQUIRE(dv.dim == dp.dim && dv.dim >= 2 && dv.dim <= 3,
               "fused Stokes operators: 2-D / 3-D spaces of equal dimension");
  SFEM_REQUIRE(dv.collocated && dv.q1d == dv.n1d,
               "the velocity space must be collocated (GLL nodes = rule)");
  SFEM_REQUIRE(dp.q1d == dv.q1d && dp.num_elements == dv.num_elements &&
                   dp.dtype == dv.dtype,
               "velocity and pressure spaces must share rule, elements, dtype");
  SFEM_REQUIRE(v->invjacs && v->jacdets, "the velocity space needs invjacs/jacdets");
  s->dim = dv.dim;
  s->N = dv.n1d;
  s->Np = dp.n1d;
  s->n = v->base.n;
  s->np = p->base.n;
  if (s->n > 1024 || s->Np > s->N) {
    set_error("fused Stokes operators: N^dim <= 1024 and Np <= N");
    return SFEM_ERR_UNSUPPORTED;
  }
  return SFEM_OK;
}

// Persistent CTAs (4 warps each) per SM; SFEM_STOKES_CTAS: developer switch.
inline int stokes_ctas_per_sm() {
  static const int v = [] {
    const char* e = getenv("SFEM_STOKES_CTAS");
    const int x = e ? atoi(e) : 0;
    return x > 0 && x <= 16 ? x : 8;
  }();
  return v;
}

template <typename T, typename K>
int launch_stokes(K kernel, const StokesShape& s, int slice_elems,
                  int64_t E, cudaStream_t stream, int* warps_out,
                  size_t* smem_out) {
  const size_t table = (size_t)(s.N * s.N + s.N * s.Np + s.N) * sizeof(T);
  int warps = kStokesMaxWarps;
  while (warps > 1 &&
         table + (size_t)warps * slice_elems * sizeof(T) > 200 * 1024)
    warps >>= 1;
  const size_t smem = table + (size_t)warps * slice_elems * sizeof(T);
  if (smem > 220 * 1024) {
    set_error("fused Stokes operators: element too large for shared memory");
    return SFEM_ERR_UNSUPPORTED;
  }
  (void)kernel;  // the attribute is set per launched instance
  *warps_out = warps;
  *smem_out = smem;
  (void)E;
  (void)stream;
  return SFEM_OK;
}

template <typename T>
int stokes_div_impl(const sfem_space* v, const sfem_space* p, const void* u,
                    void* out, cudaStream_t stream) {
  StokesShape s;
  int rc = check_pair(v, p, &s);
  if (rc) return rc;
  const int64_t E = v->base.desc.num_elements;
  SFEM_CUDA_CHECK(cudaMemsetAsync(
      out, 0, sizeof(T) * (size_t)p->base.desc.num_nodes, stream));
  if (E == 0) return SFEM_OK;
  const int slice = (s.dim + 2) * s.n;
  int warps;
  size_t smem;
  rc = launch_stokes<T>(stokes_div_kernel<T>, s, slice, E, stream, &warps, &smem);
  if (rc) return rc;
  int64_t blocks = (E + warps - 1) / warps;
  const int64_t cap = (int64_t)num_sms() * stokes_ctas_per_sm();
  if (blocks > cap) blocks = cap;
  auto go = [&](auto kernel) -> int {
    if (smem > 48 * 1024)
      SFEM_CUDA_CHECK(cudaFuncSetAttribute(
          kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kernel<<<(unsigned)blocks, 32 * warps, smem, stream>>>(
        s, tables_of<T>(v), tables_of<T>(p), v->base.desc.elements,
        p->base.desc.elements, (const T*)v->invjacs, (const T*)v->jacdets,
        (const T*)u, E, slice, (T*)out);
    return SFEM_OK;
  };
  // 2-D Stokes pairs (N, N - 2) of orders 7, 5 and 3: one lane per line
  auto go2d = [&](auto kernel, auto tabs) -> int {
    constexpr int n1 = sizeof(tabs.W) / sizeof(T);
    const size_t bytes = (size_t)4 * 6 * n1 * (n1 + 1) * sizeof(T);
    int64_t b2 = (E + 3) / 4;
    if (b2 > cap) b2 = cap;
    kernel<<<(unsigned)b2, 128, bytes, stream>>>(
        tabs, v->base.desc.elements, p->base.desc.elements,
        (const T*)v->invjacs, (const T*)v->jacdets, (const T*)u, E, (T*)out);
    return SFEM_OK;
  };
  if (s.dim == 2 && stokes_lines_enabled() && s.N == 8 && s.Np == 6)
    rc = go2d(stokes_div2d_kernel<T, 8, 6>, make_stokes_tabs<T, 8, 6>(v, p));
  else if (s.dim == 2 && stokes_lines_enabled() && s.N == 6 && s.Np == 4)
    rc = go2d(stokes_div2d_kernel<T, 6, 4>, make_stokes_tabs<T, 6, 4>(v, p));
  else if (s.dim == 2 && stokes_lines_enabled() && s.N == 4 && s.Np == 2)
    rc = go2d(stokes_div2d_kernel<T, 4, 2>, make_stokes_tabs<T, 4, 2>(v, p));
  // specialised shapes of the point-per-lane kernel
  else if (s.dim == 2 && s.N == 8 && s.Np == 6)
    rc = go(stokes_div_kernel<T, 2, 8, 6>);
  else if (s.dim == 2 && s.N == 6 && s.Np == 4)
    rc = go(stokes_div_kernel<T, 2, 6, 4>);
  else if (s.dim == 2 && s.N == 4 && s.Np == 2)
    rc = go(stokes_div_kernel<T, 2, 4, 2>);
  else if (s.dim == 3 && s.N == 8 && s.Np == 6)
    rc = go(stokes_div_kernel<T, 3, 8, 6>);
  else
    rc = go(stokes_div_kernel<T>);
  if (rc) return rc;
  SFEM_LAUNCH_CHECK();
  return SFEM_OK;
}

template <typename T>
int stokes_grad_t_impl(const sfem_space* v, const sfem_space* p,
                       const void* pr, const void* mask, void* out,
                       cudaStream_t stream) {
  StokesShape s;
  int rc = check_pair(v, p, &s);
  if (rc) return rc;
  const int64_t E = v->base.desc.num_elements;
  SFEM_CUDA_CHECK(cudaMemsetAsync(
      out, 0, sizeof(T) * (size_t)v->base.desc.num_nodes * s.dim, stream));
  if (E == 0) return SFEM_OK;
  const int slice = (2 + s.dim * s.dim) * s.n;
  int warps;
  size_t smem;
  rc = launch_stokes<T>(stokes_grad_t_kernel<T>, s, slice, E, stream, &warps,
                        &smem);
  if (rc) return rc;
  int64_t blocks = (E + warps - 1) / warps;
  const int64_t cap = (int64_t)num_sms() * stokes_ctas_per_sm();
  if (blocks > cap) blocks = cap;
  auto go = [&](auto kernel) -> int {
    if (smem > 48 * 1024)
      SFEM_CUDA_CHECK(cudaFuncSetAttribute(
          kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kernel<<<(unsigned)blocks, 32 * warps, smem, stream>>>(
        s, tables_of<T>(v), tables_of<T>(p), v->base.desc.elements,
        p->base.desc.elements, (const T*)v->invjacs, (const T*)v->jacdets,
        (const T*)pr, (const T*)mask, E, slice, (T*)out);
    return SFEM_OK;
  };
  auto go2d = [&](auto kernel, auto tabs) -> int {
    constexpr int n1 = sizeof(tabs.W) / sizeof(T);
    const size_t bytes = (size_t)4 * 6 * n1 * (n1 + 1) * sizeof(T);
    int64_t b2 = (E + 3) / 4;
    if (b2 > cap) b2 = cap;
    kernel<<<(unsigned)b2, 128, bytes, stream>>>(
        tabs, v->base.desc.elements, p->base.desc.elements,
        (const T*)v->invjacs, (const T*)v->jacdets, (const T*)pr,
        (const T*)mask, E, (T*)out);
    return SFEM_OK;
  };
  if (s.dim == 2 && stokes_lines_enabled() && s.N == 8 && s.Np == 6)
    rc = go2d(stokes_grad_t2d_kernel<T, 8, 6>, make_stokes_tabs<T, 8, 6>(v, p));
  else if (s.dim == 2 && stokes_lines_enabled() && s.N == 6 && s.Np == 4)
    rc = go2d(stokes_grad_t2d_kernel<T, 6, 4>, make_stokes_tabs<T, 6, 4>(v, p));
  else if (s.dim == 2 && stokes_lines_enabled() && s.N == 4 && s.Np == 2)
    rc = go2d(stokes_grad_t2d_kernel<T, 4, 2>, make_stokes_tabs<T, 4, 2>(v, p));
  else if (s.dim == 2 && s.N == 8 && s.Np == 6)
    rc = go(stokes_grad_t_kernel<T, 2, 8, 6>);
  else if (s.dim == 2 && s.N == 6 && s.Np == 4)
    rc = go(stokes_grad_t_kernel<T, 2, 6, 4>);
  else if (s.dim == 2 && s.N == 4 && s.Np == 2)
    rc = go(stokes_grad_t_kernel<T, 2, 4, 2>);
  else if (s.dim == 3 && s.N == 8 && s.Np == 6)
    rc = go(stokes_grad_t_kernel<T, 3, 8, 6>);
  else
    rc = go(stokes_grad_t_kernel<T>);
  if (rc) return rc;
  SFEM_LAUNCH_CHECK();
  return SFEM_OK;
}

}  // namespace
}  // namespace sfem

extern "C" {

int sfem_stokes_div(const sfem_space* vspace, const sfem_space* pspace,
                    const void* u, void* out, sfem_stream_t stream) {
  using namespace sfem;
  SFEM_REQUIRE(vspace && pspace && u && out, "null argument");
  return vspace->base.desc.dtype == SFEM_F64
             ? stokes_div_impl<double>(vspace, pspace, u, out,
                                       (cudaStream_t)stream)
             : stokes_div_impl<float>(vspace, pspace, u, out,
                                      (cudaStream_t)stream);
}

int sfem_stokes_grad_t(const sfem_space* vspace, const sfem_space* pspace,
                       const void* p, const void* mask, void* out,
                       sfem_stream_t stream) {
  using namespace sfem;
  SFEM_REQUIRE(vspace && pspace && p && out, "null argument");
  return vspace->base.desc.dtype == SFEM_F64
             ? stokes_grad_t_impl<double>(vspace, pspace, p, mask, out,
                                          (cudaStream_t)stream)
             : stokes_grad_t_impl<float>(vspace, pspace, p, mask, out,
                                         (cudaStream_t)stream);
}

}  // extern "C"
