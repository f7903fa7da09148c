// Generic (runtime N, Q, dim) sum-factorised element kernels.
//
// One CTA works on one element at a time with every intermediate tensor in
// shared memory; each 1-D contraction is a strided loop over output entries.
// This family is the *general* path: any grid/quadrature pair (B != I), any
// dimension 1..3, value / gradient evaluation, geometric-factor setup (K11),
// diagonal (K12) and the operator apply for non-collocated quadrature
// (solve_poisson's Gauss-Legendre rule, swirl_fem/examples/poisson.py:112-114).
// The collocated GLL fast kernels live in sfem_apply2d.cu / sfem_apply3d.cu and
// are cross-checked against this file in tests.
//
// Math restated (reference: swirl_fem/core/fespace.py:178-225, 338-346,
// 401-403, 458-471; interpolation.py:246-292 applies the same contractions as
// dense Kronecker matrices):
//   value  = (B x B x B) u
//   g_i    = (.. BD at axis i ..) u                reference gradient
//   J[i][j]= d x_j / d xi_i,  Jinv = J^-1,  detJ signed
//   grad_j = sum_i g_i Jinv[j][i]
//   y      = B^T ( lambda W detJ v ) + sum_i G_i^T sum_k Gf_ik g_k,
//   Gf_ik  = W detJ sum_j Jinv[j][i] Jinv[j][k]     (symmetric, d(d+1)/2)

#include "sfem_common.cuh"

namespace sfem {

namespace {

constexpr int kGenericThreads = 256;

template <typename T>
struct GenTables {
  const T* B;     // (Q, N)
  const T* BD;    // (Q, N)
  const T* W;     // (Q)
  const T* BB;    // B .* B
  const T* BDBD;  // BD .* BD
  const T* BBD;   // B .* BD
};

struct GenShape {
  int dim, N, Q, n, q, collocated;
};

__device__ __forceinline__ int ipow(int b, int e) {
  int r = 1;
  for (int i = 0; i < e; ++i) r *= b;
  return r;
}

// out[a][o][c] (+)= sum_i M[o*so + i*si] * in[a][i][c]
template <typename T>
__device__ __forceinline__ void contract(const T* __restrict__ M, int so,
                                         int si, const T* __restrict__ in,
                                         T* __restrict__ out, int A, int I,
                                         int O, int C, bool accumulate,
                                         bool identity) {
  const int total = A * O * C;
  for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
    const int c = idx % C;
    const int o = (idx / C) % O;
    const int a = idx / (C * O);
    T acc;
    if (identity) {
      acc = in[(a * I + o) * C + c];
    } else {
      acc = T(0);
      const T* ip = in + a * I * C + c;
      for (int i = 0; i < I; ++i) acc += M[o * so + i * si] * ip[i * C];
    }
    if (accumulate)
      out[idx] += acc;
    else
      out[idx] = acc;
  }
  __syncthreads();
}

// dst (Q^d) = (x_axis M_axis) src (N^d), M_axis = (axis == daxis ? BD : B).
// daxis < 0: pure interpolation.  t0, t1: scratch of max(N,Q)^d each.
template <typename T>
__device__ void forward(const GenShape& s, const GenTables<T>& tb, int daxis,
                        const T* src, T* dst, T* t0, T* t1) {
  const T* in = src;
  T* bufs[2] = {t0, t1};
  int which = 0;
  for (int axis = s.dim - 1; axis >= 0; --axis) {
    const int A = ipow(s.N, axis);
    const int C = ipow(s.Q, s.dim - 1 - axis);
    T* out = axis == 0 ? dst : bufs[which];
    const bool deriv = axis == daxis;
    contract<T>(deriv ? tb.BD : tb.B, s.N, 1, in, out, A, s.N, s.Q, C, false,
                !deriv && s.collocated);
    in = out;
    which ^= 1;
  }
}

// dst (N^d) += (x_axis M_axis^T) src (Q^d); `mats[axis]` chosen by caller.
template <typename T>
__device__ void backward(const GenShape& s, const T* const* mats,
                         const bool* ident, const T* src, T* dst, T* t0,
                         T* t1) {
  const T* in = src;
  T* bufs[2] = {t0, t1};
  int which = 0;
  for (int axis = 0; axis < s.dim; ++axis) {
    const int A = ipow(s.N, axis);
    const int C = ipow(s.Q, s.dim - 1 - axis);
    const bool last = axis == s.dim - 1;
    T* out = last ? dst : bufs[which];
    contract<T>(mats[axis], 1, s.N, in, out, A, s.Q, s.N, C, last, ident[axis]);
    in = out;
    which ^= 1;
  }
}

template <typename T>
__device__ void load_tables(const GenShape& s, const T* __restrict__ gtab,
                            T* smem, GenTables<T>* tb) {
  const int qn = s.Q * s.N;
  // global layout: [B | BD | W]; smem adds the three Hadamard tables
  for (int i = threadIdx.x; i < 2 * qn + s.Q; i += blockDim.x)
    smem[i] = gtab[i];
  __syncthreads();
  T* bb = smem + 2 * qn + s.Q;
  for (int i = threadIdx.x; i < qn; i += blockDim.x) {
    const T b = smem[i], bd = smem[qn + i];
    bb[i] = b * b;
    bb[qn + i] = bd * bd;
    bb[2 * qn + i] = b * bd;
  }
  __syncthreads();
  tb->B = smem;
  tb->BD = smem + qn;
  tb->W = smem + 2 * qn;
  tb->BB = bb;
  tb->BDBD = bb + qn;
  tb->BBD = bb + 2 * qn;
}

__host__ __device__ inline int table_elems(int N, int Q) {
  return 5 * Q * N + Q;
}

template <typename T>
__device__ __forceinline__ T quad_weight(const GenShape& s, const T* W,
                                         int qidx) {
  T w = T(1);
  for (int a = 0; a < s.dim; ++a) {
    w *= W[qidx % s.Q];
    qidx /= s.Q;
  }
  return w;
}

// d x d inverse + determinant (d <= 3), J row-major [i][j]
template <typename T>
__device__ __forceinline__ void invert(int dim, const T* J, T* inv, T* det) {
  if (dim == 1) {
    *det = J[0];
    inv[0] = T(1) / J[0];
  } else if (dim == 2) {
    const T d = J[0] * J[3] - J[1] * J[2];
    const T r = T(1) / d;
    *det = d;
    inv[0] = J[3] * r;
    inv[1] = -J[1] * r;
    inv[2] = -J[2] * r;
    inv[3] = J[0] * r;
  } else {
    const T c00 = J[4] * J[8] - J[5] * J[7];
    const T c01 = J[5] * J[6] - J[3] * J[8];
    const T c02 = J[3] * J[7] - J[4] * J[6];
    const T d = J[0] * c00 + J[1] * c01 + J[2] * c02;
    const T r = T(1) / d;
    *det = d;
    inv[0] = c00 * r;
    inv[1] = (J[2] * J[7] - J[1] * J[8]) * r;
    inv[2] = (J[1] * J[5] - J[2] * J[4]) * r;
    inv[3] = c01 * r;
    inv[4] = (J[0] * J[8] - J[2] * J[6]) * r;
    inv[5] = (J[2] * J[3] - J[0] * J[5]) * r;
    inv[6] = c02 * r;
    inv[7] = (J[1] * J[6] - J[0] * J[7]) * r;
    inv[8] = (J[0] * J[4] - J[1] * J[3]) * r;
  }
}

// ---------------------------------------------------------------------------
// K11: geometric factors.  Always computed in fp64 (differentiating the
// coordinates amplifies rounding by ~||D||), stored in the path dtype T.
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kGenericThreads)
geom_kernel(GenShape s, const double* __restrict__ gtab,
            const int32_t* __restrict__ elements,
            const T* __restrict__ node_coords, int64_t E,
            T* __restrict__ invjacs, T* __restrict__ jacdets,
            T* __restrict__ quad_coords, T* __restrict__ gf, int ngeom,
            int with_mass, double* __restrict__ jq_scratch) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* smem = reinterpret_cast<double*>(smem_raw);
  GenTables<double> tb;
  load_tables(s, gtab, smem, &tb);
  const int d = s.dim;
  const int tmax = ipow(max(s.N, s.Q), d);
  double* X = smem + table_elems(s.N, s.Q);  // one coordinate component (n)
  double* Xq = X + s.n;                      // its values at quad points (q)
  double* t0 = Xq + s.q;
  double* t1 = t0 + tmax;
  // J[i][j] at [(i*d+j)*q + p]: shared memory when it fits, else a per-CTA
  // slice of a global scratch buffer (large N in 3-D)
  double* Jq = jq_scratch ? jq_scratch + (int64_t)blockIdx.x * d * d * s.q
                          : t1 + tmax;

  for (int64_t e = blockIdx.x; e < E; e += gridDim.x) {
    for (int j = 0; j < d; ++j) {
      for (int i = threadIdx.x; i < s.n; i += blockDim.x) {
        const int32_t g = elements[e * s.n + i];
        X[i] =
            g == SFEM_SENTINEL ? 0.0 : (double)node_coords[(int64_t)g * d + j];
      }
      __syncthreads();
      if (quad_coords) {
        forward<double>(s, tb, -1, X, Xq, t0, t1);
        for (int p = threadIdx.x; p < s.q; p += blockDim.x)
          quad_coords[(e * s.q + p) * d + j] = (T)Xq[p];
      }
      for (int i = 0; i < d; ++i)
        forward<double>(s, tb, i, X, Jq + (i * d + j) * s.q, t0, t1);
    }
    for (int p = threadIdx.x; p < s.q; p += blockDim.x) {
      double J[9], inv[9], det;
      for (int a = 0; a < d * d; ++a) J[a] = Jq[a * s.q + p];
      invert<double>(d, J, inv, &det);
      const int64_t eq = e * s.q + p;
      if (invjacs)
        for (int a = 0; a < d * d; ++a) invjacs[eq * d * d + a] = (T)inv[a];
      if (jacdets) jacdets[eq] = (T)det;
      if (gf) {
        const double wd = quad_weight<double>(s, tb.W, p) * det;
        T* g = gf + e * (int64_t)ngeom * s.q + p;
        int c = 0;
        for (int i = 0; i < d; ++i)
          for (int k = i; k < d; ++k) {
            double acc = 0.0;
            for (int j = 0; j < d; ++j) acc += inv[j * d + i] * inv[j * d + k];
            g[(int64_t)c * s.q] = (T)(wd * acc);
            ++c;
          }
        if (with_mass) g[(int64_t)c * s.q] = (T)wd;
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------
// K2-K4: q-function evaluation
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kGenericThreads)
eval_kernel(GenShape s, const T* __restrict__ gtab,
            const T* __restrict__ u_local, int ncomp, int kind,
            const T* __restrict__ invjacs, int64_t E, T* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* smem = reinterpret_cast<T*>(smem_raw);
  GenTables<T> tb;
  load_tables(s, gtab, smem, &tb);
  const int d = s.dim;
  const int tmax = ipow(max(s.N, s.Q), d);
  T* U = smem + table_elems(s.N, s.Q);
  T* G = U + s.n;       // (d, q) reference gradient or (1, q) value
  T* t0 = G + d * s.q;
  T* t1 = t0 + tmax;
  const int c = blockIdx.y;

  for (int64_t e = blockIdx.x; e < E; e += gridDim.x) {
    for (int i = threadIdx.x; i < s.n; i += blockDim.x)
      U[i] = u_local[(e * s.n + i) * ncomp + c];
    __syncthreads();
    if (kind == 0) {
      forward<T>(s, tb, -1, U, G, t0, t1);
      for (int p = threadIdx.x; p < s.q; p += blockDim.x)
        out[(e * s.q + p) * ncomp + c] = G[p];
    } else {
      for (int i = 0; i < d; ++i) forward<T>(s, tb, i, U, G + i * s.q, t0, t1);
      for (int p = threadIdx.x; p < s.q; p += blockDim.x) {
        const int64_t eq = e * s.q + p;
        const T* inv = invjacs + eq * d * d;
        for (int j = 0; j < d; ++j) {
          T acc = T(0);
          for (int i = 0; i < d; ++i) acc += G[i * s.q + p] * inv[j * d + i];
          out[(eq * d + j) * ncomp + c] = acc;
        }
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------
// Transpose of K2-K4: the local covector of a functional that is linear in a
// placeholder function v (reference: jax.linear_transpose of the quadrature
// integral, swirl_fem/core/fespace.py:458-471).  Given the pointwise
// coefficients of v's value and of its physical gradient,
//   y[n] = sum_q W_q detJ_q ( a[q] phi_n(q) + sum_j b[q][j] d phi_n / d x_j (q) )
// a: (E, Q^d, ncomp) or null; b: (E, Q^d, d, ncomp) or null; y: (E, N^d, ncomp).
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kGenericThreads)
eval_transpose_kernel(GenShape s, const T* __restrict__ gtab,
                      const T* __restrict__ a, const T* __restrict__ b,
                      int ncomp, const T* __restrict__ invjacs,
                      const T* __restrict__ jacdets, int64_t E,
                      T* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* smem = reinterpret_cast<T*>(smem_raw);
  GenTables<T> tb;
  load_tables(s, gtab, smem, &tb);
  const int d = s.dim;
  const int tmax = ipow(max(s.N, s.Q), d);
  T* Y = smem + table_elems(s.N, s.Q);  // (n)
  T* H = Y + s.n;                       // (q) one weighted coefficient field
  T* t0 = H + s.q;
  T* t1 = t0 + tmax;
  const int c = blockIdx.y;

  for (int64_t e = blockIdx.x; e < E; e += gridDim.x) {
    for (int i = threadIdx.x; i < s.n; i += blockDim.x) Y[i] = T(0);
    __syncthreads();
    // part -1: value coefficient; part i >= 0: reference-gradient axis i
    for (int part = (a ? -1 : 0); part < (b ? d : 0); ++part) {
      for (int p = threadIdx.x; p < s.q; p += blockDim.x) {
        const int64_t eq = e * s.q + p;
        const T w = quad_weight<T>(s, tb.W, p) * jacdets[eq];
        T h;
        if (part < 0) {
          h = a[eq * ncomp + c];
        } else {
          // d phi / d x_j = sum_i Jinv[j][i] d phi / d xi_i
          const T* inv = invjacs + eq * d * d;
          h = T(0);
          for (int j = 0; j < d; ++j)
            h += b[(eq * d + j) * ncomp + c] * inv[j * d + part];
        }
        H[p] = w * h;
      }
      __syncthreads();
      const T* mats[3];
      bool ident[3];
      for (int axis = 0; axis < d; ++axis) {
        const bool deriv = axis == part;
        mats[axis] = deriv ? tb.BD : tb.B;
        ident[axis] = !deriv && s.collocated;
      }
      backward<T>(s, mats, ident, H, Y, t0, t1);
    }
    for (int i = threadIdx.x; i < s.n; i += blockDim.x)
      out[(e * s.n + i) * ncomp + c] = Y[i];
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------
// K5: integration
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
integrate_kernel(GenShape s, const T* __restrict__ gtab,
                 const T* __restrict__ w, const T* __restrict__ jacdets,
                 int64_t total, double* __restrict__ result) {
  __shared__ double red[32];
  __shared__ T W[SFEM_MAX_1D];
  if (threadIdx.x < s.Q) W[threadIdx.x] = gtab[2 * s.Q * s.N + threadIdx.x];
  __syncthreads();
  double acc = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int p = (int)(i % s.q);
    acc += (double)w[i] * (double)jacdets[i] * (double)quad_weight<T>(s, W, p);
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(result, acc);
}

// ---------------------------------------------------------------------------
// operator apply, generic
// ---------------------------------------------------------------------------
template <typename T, bool LOCAL>
__global__ void __launch_bounds__(kGenericThreads)
apply_generic_kernel(GenShape s, const T* __restrict__ gtab,
                     const uint32_t* __restrict__ conn,
                     const T* __restrict__ gf, int ngeom, int with_mass,
                     T lambda, T mu, const T* __restrict__ x,
                     T* __restrict__ y, int ncomp, int64_t E,
                     double* __restrict__ dot_xy) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double red[32];
  T* smem = reinterpret_cast<T*>(smem_raw);
  GenTables<T> tb;
  load_tables(s, gtab, smem, &tb);
  const int d = s.dim;
  const int tmax = ipow(max(s.N, s.Q), d);
  T* U = smem + table_elems(s.N, s.Q);
  T* Y = U + s.n;
  T* F = Y + s.n;            // (d+1, q): g_0..g_{d-1}, value
  T* t0 = F + (d + 1) * s.q;
  T* t1 = t0 + tmax;
  const int c = blockIdx.y;
  const bool do_mass = lambda != T(0);
  const bool do_stiff = mu != T(0);
  double dot = 0.0;

  for (int64_t e = blockIdx.x; e < E; e += gridDim.x) {
    for (int i = threadIdx.x; i < s.n; i += blockDim.x) {
      T v;
      if (LOCAL) {
        v = x[(e * s.n + i) * ncomp + c];
      } else {
        const uint32_t cn = conn[e * s.n + i];
        v = cn == kConnSentinel
                ? T(0)
                : x[(int64_t)(cn & kConnIdMask) * ncomp + c];
      }
      U[i] = v;
      Y[i] = T(0);
    }
    __syncthreads();
    if (do_stiff)
      for (int i = 0; i < d; ++i) forward<T>(s, tb, i, U, F + i * s.q, t0, t1);
    if (do_mass) forward<T>(s, tb, -1, U, F + d * s.q, t0, t1);
    const T* g = gf + e * (int64_t)ngeom * s.q;
    for (int p = threadIdx.x; p < s.q; p += blockDim.x) {
      if (do_stiff) {
        T gr[3], w[3] = {T(0), T(0), T(0)};
        for (int i = 0; i < d; ++i) gr[i] = F[i * s.q + p];
        for (int i = 0; i < d; ++i)
          for (int k = 0; k < d; ++k) {
            const int si = i <= k ? sym_index(d, i, k) : sym_index(d, k, i);
            w[i] += g[(int64_t)si * s.q + p] * gr[k];
          }
        for (int i = 0; i < d; ++i) F[i * s.q + p] = mu * w[i];
      }
      if (do_mass)
        F[d * s.q + p] *= lambda * g[(int64_t)(ngeom - 1) * s.q + p];
    }
    __syncthreads();
    if (do_stiff)
      for (int i = 0; i < d; ++i) {
        const T* mats[3];
        bool ident[3];
        for (int a = 0; a < d; ++a) {
          mats[a] = a == i ? tb.BD : tb.B;
          ident[a] = a != i && s.collocated;
        }
        backward<T>(s, mats, ident, F + i * s.q, Y, t0, t1);
      }
    if (do_mass) {
      const T* mats[3] = {tb.B, tb.B, tb.B};
      bool ident[3] = {(bool)s.collocated, (bool)s.collocated,
                       (bool)s.collocated};
      backward<T>(s, mats, ident, F + d * s.q, Y, t0, t1);
    }
    for (int i = threadIdx.x; i < s.n; i += blockDim.x) {
      const T v = Y[i];
      if (LOCAL) {
        y[(e * s.n + i) * ncomp + c] = v;
      } else {
        const uint32_t cn = conn[e * s.n + i];
        if (cn != kConnSentinel) {
          T* dst = y + (int64_t)(cn & kConnIdMask) * ncomp + c;
          if (cn & kConnDirichlet) {
            if (cn & kConnSingle) *dst = T(0);
          } else {
            if (cn & kConnSingle)
              *dst = v;
            else
              red_add(dst, v);
            dot += (double)U[i] * (double)v;
          }
        }
      }
    }
    __syncthreads();
  }
  if (!LOCAL && dot_xy != nullptr) {
    dot = block_sum(dot, red);
    if (threadIdx.x == 0) atomicAdd(dot_xy, dot);
  }
}

// ---------------------------------------------------------------------------
// K12: diagonal
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kGenericThreads)
diag_generic_kernel(GenShape s, const T* __restrict__ gtab,
                    const uint32_t* __restrict__ conn,
                    const T* __restrict__ gf, int ngeom, int with_mass,
                    T lambda, T mu, T* __restrict__ diag, int64_t E) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* smem = reinterpret_cast<T*>(smem_raw);
  GenTables<T> tb;
  load_tables(s, gtab, smem, &tb);
  const int d = s.dim;
  const int tmax = ipow(max(s.N, s.Q), d);
  T* Y = smem + table_elems(s.N, s.Q);
  T* F = Y + s.n;  // (q)
  T* t0 = F + s.q;
  T* t1 = t0 + tmax;
  for (int64_t e = blockIdx.x; e < E; e += gridDim.x) {
    for (int i = threadIdx.x; i < s.n; i += blockDim.x) Y[i] = T(0);
    __syncthreads();
    const T* g = gf + e * (int64_t)ngeom * s.q;
    if (mu != T(0)) {
      for (int i = 0; i < d; ++i)
        for (int k = 0; k < d; ++k) {
          const int si = i <= k ? sym_index(d, i, k) : sym_index(d, k, i);
          for (int p = threadIdx.x; p < s.q; p += blockDim.x)
            F[p] = mu * g[(int64_t)si * s.q + p];
          __syncthreads();
          const T* mats[3];
          bool ident[3];
          for (int a = 0; a < d; ++a) {
            const bool di = a == i, dk = a == k;
            mats[a] = di && dk ? tb.BDBD : (di || dk ? tb.BBD : tb.BB);
            ident[a] = false;  // Hadamard tables are explicit even if B = I
          }
          backward<T>(s, mats, ident, F, Y, t0, t1);
        }
    }
    if (lambda != T(0)) {
      for (int p = threadIdx.x; p < s.q; p += blockDim.x)
        F[p] = lambda * g[(int64_t)(ngeom - 1) * s.q + p];
      __syncthreads();
      const T* mats[3] = {tb.BB, tb.BB, tb.BB};
      bool ident[3] = {false, false, false};
      backward<T>(s, mats, ident, F, Y, t0, t1);
    }
    for (int i = threadIdx.x; i < s.n; i += blockDim.x) {
      const uint32_t cn = conn[e * s.n + i];
      if (cn == kConnSentinel) continue;
      T* dst = diag + (cn & kConnIdMask);
      if (cn & kConnDirichlet) {
        if (cn & kConnSingle) *dst = T(0);
      } else if (cn & kConnSingle) {
        *dst = Y[i];
      } else {
        red_add(dst, Y[i]);
      }
    }
    __syncthreads();
  }
}

GenShape make_shape(const SpaceBase& b) {
  GenShape s;
  s.dim = b.desc.dim;
  s.N = b.desc.n1d;
  s.Q = b.desc.q1d;
  s.n = b.n;
  s.q = b.q;
  s.collocated = b.desc.collocated;
  return s;
}

int ipow_host(int b, int e) {
  int r = 1;
  for (int i = 0; i < e; ++i) r *= b;
  return r;
}

template <typename K>
int prepare_smem(K kernel, size_t bytes) {
  if (bytes > 227 * 1024) {
    set_error("generic kernel needs " + std::to_string(bytes) +
              " B of shared memory (> 227 KB): N/Q too large for the generic "
              "path");
    return SFEM_ERR_UNSUPPORTED;
  }
  if (bytes > 48 * 1024)
    SFEM_CUDA_CHECK(cudaFuncSetAttribute(
        kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return SFEM_OK;
}

int grid_for(int64_t E, int ctas_per_sm) {
  const int64_t cap = (int64_t)num_sms() * ctas_per_sm;
  return (int)(E < cap ? (E > 0 ? E : 1) : cap);
}

}  // namespace

// ---- host launchers ----------------------------------------------------------

template <typename T>
int launch_geom(const SpaceBase& b, void* invjacs, void* jacdets,
                void* quad_coords, void* gf, int ngeom, int with_mass,
                cudaStream_t stream) {
  const GenShape s = make_shape(b);
  const int d = s.dim;
  const int tmax = ipow_host(s.N > s.Q ? s.N : s.Q, d);
  const size_t base = table_elems(s.N, s.Q) + (size_t)s.n + (size_t)s.q +
                      2 * (size_t)tmax;
  const size_t jq = (size_t)d * d * s.q;
  const int64_t E = b.desc.num_elements;
  if (E == 0) return SFEM_OK;
  const bool spill_jq = (base + jq) * sizeof(double) > 160 * 1024;
  const size_t bytes = (base + (spill_jq ? 0 : jq)) * sizeof(double);
  int rc = prepare_smem(geom_kernel<T>, bytes);
  if (rc) return rc;
  const int grid = grid_for(E, spill_jq ? 1 : 4);
  double* scratch = nullptr;
  if (spill_jq)
    SFEM_CUDA_CHECK(cudaMalloc(&scratch, sizeof(double) * jq * (size_t)grid));
  geom_kernel<T><<<grid, kGenericThreads, bytes, stream>>>(
      s, b.d_tables64, b.desc.elements, (const T*)b.desc.node_coords, E,
      (T*)invjacs, (T*)jacdets, (T*)quad_coords, (T*)gf, ngeom, with_mass,
      scratch);
  SFEM_LAUNCH_CHECK();
  if (scratch) {
    SFEM_CUDA_CHECK(cudaStreamSynchronize(stream));
    cudaFree(scratch);
  }
  return SFEM_OK;
}

template <typename T>
int launch_eval(const SpaceBase& b, const void* u_local, int ncomp, int kind,
                const void* invjacs, void* out, cudaStream_t stream) {
  const GenShape s = make_shape(b);
  const int d = s.dim;
  const int tmax = ipow_host(s.N > s.Q ? s.N : s.Q, d);
  const size_t elems =
      table_elems(s.N, s.Q) + s.n + (size_t)d * s.q + 2 * (size_t)tmax;
  const size_t bytes = elems * sizeof(T);
  int rc = prepare_smem(eval_kernel<T>, bytes);
  if (rc) return rc;
  const int64_t E = b.desc.num_elements;
  if (E == 0) return SFEM_OK;
  dim3 grid(grid_for(E, 4), ncomp);
  eval_kernel<T><<<grid, kGenericThreads, bytes, stream>>>(
      s, tables<T>(b), (const T*)u_local, ncomp, kind, (const T*)invjacs, E,
      (T*)out);
  SFEM_LAUNCH_CHECK();
  return SFEM_OK;
}

template <typename T>
int launch_eval_transpose(const SpaceBase& b, const void* a, const void* bg,
                          int ncomp, const void* invjacs, const void* jacdets,
                          void* out, cudaStream_t stream) {
  const GenShape s = make_shape(b);
  const int d = s.dim;
  const int tmax = ipow_host(s.N > s.Q ? s.N : s.Q, d);
  const size_t elems =
      table_elems(s.N, s.Q) + s.n + (size_t)s.q + 2 * (size_t)tmax;
  const size_t bytes = elems * sizeof(T);
  int rc = prepare_smem(eval_transpose_kernel<T>, bytes);
  if (rc) return rc;
  const int64_t E = b.desc.num_elements;
  if (E == 0) return SFEM_OK;
  dim3 grid(grid_for(E, 4), ncomp);
  eval_transpose_kernel<T><<<grid, kGenericThreads, bytes, stream>>>(
      s, tables<T>(b), (const T*)a, (const T*)bg, ncomp, (const T*)invjacs,
      (const T*)jacdets, E, (T*)out);
  SFEM_LAUNCH_CHECK();
  return SFEM_OK;
}

template <typename T>
int launch_integrate(const SpaceBase& b, const void* w, const void* jacdets,
                     double* result, cudaStream_t stream) {
  const GenShape s = make_shape(b);
  const int64_t total = b.desc.num_elements * (int64_t)s.q;
  SFEM_CUDA_CHECK(cudaMemsetAsync(result, 0, sizeof(double), stream));
  if (total == 0) return SFEM_OK;
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  integrate_kernel<T><<<(int)blocks, 256, 0, stream>>>(
      s, tables<T>(b), (const T*)w, (const T*)jacdets, total, result);
  SFEM_LAUNCH_CHECK();
  return SFEM_OK;
}

template <typename T>
int launch_apply_generic(const sfem_op& op, double lambda, double mu,
                         const void* x, void* y, int ncomp, bool local,
                         double* dot_xy, cudaStream_t stream) {
  const SpaceBase& b = op.base;
  const GenShape s = make_shape(b);
  const int d = s.dim;
  const int tmax = ipow_host(s.N > s.Q ? s.N : s.Q, d);
  const size_t elems = table_elems(s.N, s.Q) + 2 * (size_t)s.n +
                       (size_t)(d + 1) * s.q + 2 * (size_t)tmax;
  const size_t bytes = elems * sizeof(T);
  const int64_t E = b.desc.num_elements;
  if (E == 0) return SFEM_OK;
  dim3 grid(grid_for(E, 4), ncomp);
  if (local) {
    int rc = prepare_smem(apply_generic_kernel<T, true>, bytes);
    if (rc) return rc;
    apply_generic_kernel<T, true><<<grid, kGenericThreads, bytes, stream>>>(
        s, tables<T>(b), op.conn, (const T*)op.geom, op.ngeom, op.with_mass,
        (T)lambda, (T)mu, (const T*)x, (T*)y, ncomp, E, nullptr);
  } else {
    int rc = prepare_smem(apply_generic_kernel<T, false>, bytes);
    if (rc) return rc;
    apply_generic_kernel<T, false><<<grid, kGenericThreads, bytes, stream>>>(
        s, tables<T>(b), op.conn, (const T*)op.geom, op.ngeom, op.with_mass,
        (T)lambda, (T)mu, (const T*)x, (T*)y, ncomp, E, dot_xy);
  }
  SFEM_LAUNCH_CHECK();
  return SFEM_OK;
}

template <typename T>
int launch_diag_generic(const sfem_op& op, double lambda, double mu,
                        void* diag, cudaStream_t stream) {
  const SpaceBase& b = op.base;
  const GenShape s = make_shape(b);
  const int d = s.dim;
  const int tmax = ipow_host(s.N > s.Q ? s.N : s.Q, d);
  const size_t elems =
      table_elems(s.N, s.Q) + (size_t)s.n + (size_t)s.q + 2 * (size_t)tmax;
  const size_t bytes = elems * sizeof(T);
  int rc = prepare_smem(diag_generic_kernel<T>, bytes);
  if (rc) return rc;
  const int64_t E = b.desc.num_elements;
  if (E == 0) return SFEM_OK;
  diag_generic_kernel<T><<<grid_for(E, 4), kGenericThreads, bytes, stream>>>(
      s, tables<T>(b), op.conn, (const T*)op.geom, op.ngeom, op.with_mass,
      (T)lambda, (T)mu, (T*)diag, E);
  SFEM_LAUNCH_CHECK();
  return SFEM_OK;
}

#define SFEM_INSTANTIATE(T)                                                    \
  template int launch_geom<T>(const SpaceBase&, void*, void*, void*, void*,    \
                              int, int, cudaStream_t);                         \
  template int launch_eval<T>(const SpaceBase&, const void*, int, int,         \
                              const void*, void*, cudaStream_t);               \
  template int launch_eval_transpose<T>(const SpaceBase&, const void*,         \
                                        const void*, int, const void*,         \
                                        const void*, void*, cudaStream_t);     \
  template int launch_integrate<T>(const SpaceBase&, const void*, const void*, \
                                   double*, cudaStream_t);                     \
  template int launch_apply_generic<T>(const sfem_op&, double, double,         \
                                       const void*, void*, int, bool, double*, \
                                       cudaStream_t);                          \
  template int launch_diag_generic<T>(const sfem_op&, double, double, void*,   \
                                      cudaStream_t);
SFEM_INSTANTIATE(float)
SFEM_INSTANTIATE(double)

}  // namespace sfem
