/*
 * swirl_b200.h -- C ABI of libswirl_b200.so: the B200-native (sm_100a) hot
 * path of Swirl-FEM (matrix-free tensor-product operator apply, element <->
 * global gather-scatter, fused PCG, shared-dof exchange).
 *
 * The reference (google-research/swirl-fem) is pure Python/JAX and has NO FFI
 * of its own; each entry point below replaces one JAX-level function of the
 * reference (cited as file:line relative to the reference root) and is what a
 * `jax.ffi` custom call for that function would bind (see INTEGRATION.md and
 * swirl_fem_b200/csrc/xla_ffi_shim.cc).
 *
 * Conventions
 *   - plain C: pointers, sizes, scalars; no torch / jax types.
 *   - every array pointer is a DEVICE pointer owned by the caller unless the
 *     parameter is documented as "host".  Nothing is freed by the library
 *     except the handles it allocates (small host structs + O(N^2) device
 *     tables for the 1-D matrices).
 *   - work is enqueued on the caller's stream (`sfem_stream_t`, a
 *     cudaStream_t); no host synchronisation except where stated.
 *   - return value: 0 on success, negative `sfem_status` on error;
 *     `sfem_last_error()` gives a thread-local message.
 *   - dtype is the arithmetic type of the path: SFEM_F32 or SFEM_F64.
 *   - tensor index convention follows the reference: coordinate 0 is the
 *     SLOWEST axis of the element-local node ordering
 *     (swirl_fem/core/interpolation.py:252, 285-286).
 */
#ifndef SWIRL_B200_H_
#define SWIRL_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* sfem_stream_t;

enum sfem_dtype { SFEM_F32 = 0, SFEM_F64 = 1 };

enum sfem_status {
  SFEM_OK = 0,
  SFEM_ERR_INVALID = -1,      /* bad argument / shape                        */
  SFEM_ERR_UNSUPPORTED = -2,  /* combination not implemented (no fallback)   */
  SFEM_ERR_CUDA = -3,         /* CUDA runtime error, see sfem_last_error()   */
  SFEM_ERR_NCCL = -4
};

#define SFEM_SENTINEL (-1) /* swirl_fem/core/gather_scatter.py:118 */
#define SFEM_MAX_1D 18     /* max nodes / quadrature points per axis         */

const char* sfem_last_error(void);
int sfem_version(void);
/* Number of kernels launched by this library in this process (bench.py's
 * `gpu_launches` claim). */
int64_t sfem_launch_count(void);

/* ------------------------------------------------------------------------ */
/* K1 / K7 / K9: gather, scatter, exchange                                   */
/* ------------------------------------------------------------------------ */

/* out[i] = indices[i] == SENTINEL ? fill : u[indices[i]*stride + offset].
 * Replaces gather_scatter.gather (swirl_fem/core/gather_scatter.py:121-127)
 * and Mesh.gather (swirl_fem/core/mesh.py:155-160).  `out` is written with
 * element stride `stride` too, so a (G,d) AoS field is gathered one component
 * per call (the reference vmaps over the last axis, navier_stokes.py:210-212)
 * or, with offset = -1, all `stride` components in ONE launch.  The same
 * convention holds for sfem_scatter_add (which then zeroes the whole
 * (num_nodes, stride) field) and sfem_exchange (whose scratch then holds
 * num_unique * stride values). */
int sfem_gather(int dtype, const void* u, const int32_t* indices,
                int64_t count, double fill_value, int32_t stride,
                int32_t offset, void* out, sfem_stream_t stream);

/* out = zeros(num_nodes); out[indices[i]] += u_local[i] (SENTINEL skipped).
 * Replaces gather_scatter.scatter (gather_scatter.py:130-133) / Mesh.scatter
 * (mesh.py:165-168).  Atomic (REDG) accumulation; `out` is zeroed first. */
int sfem_scatter_add(int dtype, const void* u_local, const int32_t* indices,
                     int64_t count, int64_t num_nodes, int32_t stride,
                     int32_t offset, void* out, sfem_stream_t stream);

/* Deterministic scatter: a transposed (node -> local slots) CSR map built once
 * on the device, then a warp-segmented gather-sum with no atomics.
 * Workspace sizes are returned by sfem_scatter_plan_size(). */
typedef struct sfem_scatter_plan sfem_scatter_plan;
int sfem_scatter_plan_create(const int32_t* indices, int64_t count,
                             int64_t num_nodes, sfem_scatter_plan** plan,
                             sfem_stream_t stream);
int sfem_scatter_plan_apply(const sfem_scatter_plan* plan, int dtype,
                            const void* u_local, int32_t stride,
                            int32_t offset, void* out, sfem_stream_t stream);
void sfem_scatter_plan_destroy(sfem_scatter_plan* plan);

/* Unpartitioned QQ^T (periodic dofs), gather_scatter.exchange
 * (gather_scatter.py:189-261): u[gi] <- sum over the group `ui` of u[gi].
 * `scratch` holds `num_unique` values of dtype.  In place on `u`. */
int sfem_exchange(int dtype, void* u, const int32_t* gather_indices,
                  const int32_t* unique_indices, int64_t count,
                  int64_t num_unique, int32_t stride, int32_t offset,
                  void* scratch, sfem_stream_t stream);

/* Partitioned QQ^T building blocks (halo exchange; the wire step is NCCL /
 * NVLink peer copies driven by the host layer):
 * pack:   buf[i]  = u[send_idx[i]]
 * unpack: u[recv_idx[i]] += buf[i]                                          */
int sfem_halo_pack(int dtype, const void* u, const int32_t* idx, int64_t count,
                   void* buf, sfem_stream_t stream);
int sfem_halo_unpack_add(int dtype, void* u, const int32_t* idx, int64_t count,
                         const void* buf, sfem_stream_t stream);
/* Canonical form of the unpack: for each of the `num_dofs` unique interface
 * dofs, u[dofs[i]] = sum of the contributions src[row_ptr[i] .. row_ptr[i+1])
 * in the stored (ascending rank) order, where src >= 0 indexes `recv` and
 * src < 0 stands for this rank's own value.  All ranks evaluate the same
 * expression, so replicated dofs stay bitwise identical across ranks. */
int sfem_halo_unpack_canonical(int dtype, void* u, const int32_t* dofs,
                               const int32_t* row_ptr, const int32_t* src,
                               int64_t num_dofs, const void* recv,
                               sfem_stream_t stream);

/* ------------------------------------------------------------------------ */
/* FE space: geometric factors, q-function evaluation, integration          */
/* ------------------------------------------------------------------------ */

typedef struct {
  int32_t dim;            /* 1, 2, 3                                         */
  int32_t n1d;            /* N: grid nodes per axis (order + 1)              */
  int32_t q1d;            /* Q: quadrature points per axis                   */
  int32_t dtype;          /* SFEM_F32 / SFEM_F64                             */
  int32_t collocated;     /* 1: quadrature nodes == grid nodes (B = I)       */
  int32_t reserved;
  int64_t num_elements;   /* E                                               */
  int64_t num_nodes;      /* G                                               */
  const int32_t* elements;      /* device (E, N^dim) int32, SENTINEL allowed */
  const void* node_coords;      /* device (G, dim), dtype                    */
  const double* interp_1d;      /* host (Q, N): B[q][n] = l_n(x_q)           */
  const double* interp_grad_1d; /* host (Q, N): B @ D                        */
  const double* quad_weights_1d;/* host (Q)                                  */
} sfem_space_desc;

typedef struct sfem_space sfem_space;

/* FiniteElementSpace.create (swirl_fem/core/fespace.py:306-348).  Outputs are
 * caller-owned device arrays (any may be NULL):
 *   invjacs (E, Q^d, d, d), jacdets (E, Q^d) [signed], quad_coords (E, Q^d, d).
 * The handle keeps the descriptor, device copies of the 1-D matrices and the
 * three output pointers (they must outlive the handle). */
int sfem_space_create(const sfem_space_desc* desc, void* invjacs,
                      void* jacdets, void* quad_coords, sfem_space** space,
                      sfem_stream_t stream);
void sfem_space_destroy(sfem_space* space);

/* Scalar/VectorNodalQFunction[Grad]._evaluate (fespace.py:178-225).
 * u_local: (E, N^d, ncomp) AoS.  kind 0: values -> out (E, Q^d, ncomp);
 * kind 1: physical gradient -> out (E, Q^d, d, ncomp), out[..j,k] = d u_k/dx_j
 * (for ncomp == 1 that is the reference's (E, Q^d, d)). */
int sfem_space_eval(const sfem_space* space, const void* u_local,
                    int32_t ncomp, int32_t kind, void* out,
                    sfem_stream_t stream);

/* Transpose of sfem_space_eval: the local covector of a functional that is
 * linear in a placeholder function v, which the reference obtains with
 * jax.linear_transpose of the quadrature integral (fespace.py:458-471):
 *   out[e,n,c] = sum_q W[q] jacdets[e,q] ( vals[e,q,c] phi_n(q)
 *                + sum_j grads[e,q,j,c] d phi_n / d x_j (q) )
 * vals: (E, Q^d, ncomp) coefficient of v_c's value, or NULL; grads:
 * (E, Q^d, d, ncomp) coefficient of d v_c / d x_j, or NULL; out:
 * (E, N^d, ncomp).  Used for the mixed velocity / pressure forms of the
 * Stokes solver (D, D^T, convection; navier_stokes.py:238-245, 313-338). */
int sfem_space_eval_transpose(const sfem_space* space, const void* vals,
                              const void* grads, int32_t ncomp, void* out,
                              sfem_stream_t stream);

/* Pointwise forms of the Stokes operators at the quadrature points, between
 * sfem_space_eval and sfem_space_eval_transpose (AoS layouts of those calls):
 *   kind 0: out[p] = sum_k g[p][k][k]              div(v), navier_stokes.py:315
 *   kind 1: out[p][j][k] = (j == k) a[p]           coefficient of grad v in
 *                                                  div(v) q, navier_stokes.py:324
 *   kind 2: out[p][k] = sum_i a[p][i] g[p][i][k]   (u . grad) w, :239-240
 * a: (P) [kind 1] or (P, d) [kind 2]; g: (P, d, d) with g[j][k] = d w_k / d x_j. */
int sfem_pointwise(int dtype, int32_t kind, int32_t dim, const void* a,
                   const void* g, int64_t num_points, void* out,
                   sfem_stream_t stream);

/* FiniteElementSpace.integrate (fespace.py:381-403):
 * *result = sum_{e,q} w[e,q] * jacdets[e,q] * W[q].  result: device fp64
 * scalar (always double, zeroed by the call). */
int sfem_space_integrate(const sfem_space* space, const void* w, void* result,
                         sfem_stream_t stream);

/* ------------------------------------------------------------------------ */
/* Operator: y = mask . Z^T (lambda*Mass + mu*Stiffness) Z x                 */
/* ------------------------------------------------------------------------ */

typedef struct sfem_op sfem_op;

/* Bytes needed for the packed geometric factors and the packed connectivity
 * of an operator on `desc` (caller allocates both on the device). */
int64_t sfem_op_geom_bytes(const sfem_space_desc* desc, int32_t with_mass);
int64_t sfem_op_conn_bytes(const sfem_space_desc* desc);

/* Builds the operator of FiniteElementSpace.local_covector
 * (fespace.py:405-471) for the bilinear forms
 *   lambda * u v + mu * grad u . grad v
 * (mass `l` / stiffness `a`, swirl_fem/examples/poisson.py:133-137; Helmholtz
 * H_, swirl_fem/navier_stokes/navier_stokes.py:431), fused with Mesh.gather,
 * Mesh.scatter and the Dirichlet mask of poisson.py:119-130, 141-146.
 *   dirichlet: device (G) uint8, 1 = constrained row (zeroed), or NULL.
 *   geom / conn: caller-owned device buffers of the sizes above; filled here
 *   (one-time setup kernels, K11) and kept by the handle.
 *   with_mass: store W*detJ as well (needed when lambda != 0).              */
int sfem_op_create(const sfem_space_desc* desc, const uint8_t* dirichlet,
                   int32_t with_mass, void* geom, void* conn, sfem_op** op,
                   sfem_stream_t stream);
void sfem_op_destroy(sfem_op* op);

/* y = mask . scatter(local_op(gather(x))).  x, y: (G, ncomp) AoS of dtype;
 * y is fully overwritten.  If `dot_xy` != NULL it receives x . y (device fp64
 * scalar, always double) computed in the kernel epilogue -- the p.Ap of
 * cg.py:77.  */
int sfem_op_apply(const sfem_op* op, double lambda, double mu, const void* x,
                  void* y, int32_t ncomp, void* dot_xy, sfem_stream_t stream);

/* Same operator restricted to the elements [elem_begin, elem_end) (elem_begin
 * a multiple of 4).  `first` != 0: zero the shared-dof prefix of y and dot_xy
 * before accumulating; later calls on the same y pass first = 0.  Used to
 * overlap the halo exchange of a partitioned mesh with interior compute: the
 * rank's interface elements are stored first, applied first, their shared
 * dofs are packed and sent while the interior elements are applied. */
int sfem_op_apply_range(const sfem_op* op, double lambda, double mu,
                        const void* x, void* y, int32_t ncomp,
                        int64_t elem_begin, int64_t elem_end, int32_t first,
                        void* dot_xy, sfem_stream_t stream);

/* Local (E-vector) form: y_local = local_covector(form, (u_local, v)).
 * u_local, y_local: (E, N^d, ncomp). */
int sfem_op_apply_local(const sfem_op* op, double lambda, double mu,
                        const void* u_local, void* y_local, int32_t ncomp,
                        sfem_stream_t stream);

/* diag(mask . Z^T (lambda M + mu K) Z) (K12; Jacobi preconditioner).        */
int sfem_op_diag(const sfem_op* op, double lambda, double mu, void* diag,
                 sfem_stream_t stream);

/* Selects the kernel family: 0 = auto (specialised collocated kernels where
 * available), 1 = force the generic runtime-(N,Q) kernel.  For tests.       */
int sfem_op_set_variant(sfem_op* op, int32_t variant);

/* Lazy zero fill of y's shared-dof prefix (the part of Mesh.scatter's
 * `.at[].add` target that several elements accumulate into,
 * swirl_fem/core/mesh.py:211-227).  By default sfem_op_apply zeroes that
 * prefix before its kernel; for a prefix much larger than L2 this costs two
 * extra DRAM accesses per shared dof.  With the tables below a COMPANION
 * kernel zeroes every shared dof shortly before the first element that
 * touches it is scattered, while the apply runs, and the apply's CTAs claim
 * their element steps from a counter (one tight window of consecutive steps)
 * instead of striding through them (3-D collocated operators with a compiled
 * instance, ncomp == 1; otherwise the eager fill stays).
 *   sfem_op_lazy_zero_query: elements per CTA step and grid of the launch
 *     sfem_op_apply would make; *supported = 1 if a lazy instance exists.
 *   sfem_op_set_lazy_zero: the elements are cut into chunks of `chunk_steps`
 *     consecutive CTA steps; `pieces` device int32 (num_pieces, 2) = {first
 *     dof, length | chunk << 12} ranges (length <= 4095) covering [0, n_zero)
 *     exactly once, sorted by chunk, where chunk j = the dofs first touched
 *     by CTA steps [j * chunk_steps, (j + 1) * chunk_steps) (dofs no element
 *     touches: chunk 0); `chunk_ptr` device int32 (num_chunks + 1) = first
 *     piece of every chunk; a chunk is zeroed `ahead_steps` steps before its
 *     first step is claimed (a step is claimed three steps before it runs).
 *     The tables are retained (caller keeps them alive); chunk_ptr == NULL
 *     switches back to the eager fill.  The lazy launches of one handle must
 *     all go to ONE stream (other streams use the eager fill).
 *   sfem_op_lazy_zero_timed_out: 1 if a device-side wait of the protocol ever
 *     hit its 2 s limit (results since then are invalid; sfem_last_error
 *     tells which wait); synchronises. */
/* Number of leading entries of y that several elements accumulate into (the
 * prefix that must be zero before an apply). */
int64_t sfem_op_num_zero(const sfem_op* op);
int sfem_op_lazy_zero_query(const sfem_op* op, int32_t* step_elems,
                            int32_t* grid, int32_t* supported);
int sfem_op_set_lazy_zero(sfem_op* op, const void* pieces, int64_t num_pieces,
                          const void* chunk_ptr, int32_t num_chunks,
                          int32_t chunk_steps, int32_t ahead_steps);
int sfem_op_lazy_zero_timed_out(const sfem_op* op, sfem_stream_t stream);

/* ------------------------------------------------------------------------ */
/* Peer-memory halo exchange (partitioned QQ^T over NVLink P2P stores)       */
/* ------------------------------------------------------------------------ */

/* The reference's partitioned exchange is ONE dense lax.psum over all shared
 * dofs (swirl_fem/core/gather_scatter.py:246-248).  Here every rank owns a
 * peer-mapped region (CUDA IPC) holding its flag words and two receive
 * buffers (epoch parity); a rank writes its shared dofs directly into its
 * peers' buffers with NVLink stores and raises a flag, the receiver waits for
 * its peers' flags and sums all holders' values in ascending rank order.
 *
 * sfem_ipc_*: device memory that other processes on the node can map.
 *   alloc: cudaMalloc + zero + export (handle: 64 bytes, host);
 *   open:  map a region exported by another process (peer access enabled). */
#define SFEM_IPC_HANDLE_BYTES 64
int sfem_ipc_alloc(int64_t bytes, void** dev_ptr, void* handle);
int sfem_ipc_open(const void* handle, void** dev_ptr);
int sfem_ipc_close(void* dev_ptr);
int sfem_ipc_free(void* dev_ptr);

typedef struct {
  int32_t dtype;
  int32_t rank, world;
  int32_t num_peers;
  /* send side */
  int64_t num_send;              /* entries, all peers concatenated          */
  const int32_t* send_idx;       /* device (num_send): local dof             */
  const uint64_t* send_dst;      /* device (num_send): address of the slot in
                                    the peer's parity-0 receive buffer       */
  uint64_t parity_stride_bytes;  /* parity-1 buffers = parity-0 + this (the
                                    same on every rank)                      */
  const uint64_t* peer_flag_addr;/* HOST (num_peers): address of this rank's
                                    parity-0 flag word on each peer; the
                                    parity-1 word is `world` words further   */
  /* receive side */
  const int32_t* peer_ranks;     /* HOST (num_peers), ascending              */
  const uint64_t* flags;         /* device (2 * world): this rank's flag words
                                    [parity][source rank]                    */
  const void* recv;              /* device: parity-0 receive buffer          */
  /* canonical sum (see sfem_halo_unpack_canonical) */
  int64_t num_dofs;
  const int32_t* dofs;
  const int32_t* row_ptr;
  const int32_t* src;
} sfem_halo_desc;

typedef struct sfem_halo sfem_halo;

/* All device arrays of the descriptor must outlive the handle.  Unlike
 * sfem_op, a sfem_halo is a STATEFUL endpoint (epoch counter, device
 * counters): calls on one handle must be issued from one host thread at a
 * time and on one stream, in the same order on every rank. */
int sfem_halo_create(const sfem_halo_desc* desc, sfem_halo** halo);
void sfem_halo_destroy(sfem_halo* halo);

/* Tuning (call between epochs only).  key 0: entries per work item of the
 * cooperative push / sum (default 256); key 1: who runs the exchange of
 * sfem_op_apply_halo --
 *   0  push inside the apply kernel, canonical sum in sfem_halo_wait_unpack
 *      (runs after the apply);
 *   1  push and canonical sum inside the apply kernel's CTAs;
 *   2  push inside the apply kernel, canonical sum in the kernel of
 *      sfem_halo_wait_unpack running CONCURRENTLY with the apply's interior
 *      elements (programmatic dependent launch, 64-thread CTAs in spare slots);
 *   3  (default) push AND canonical sum in that concurrent kernel: the apply's
 *      CTAs only signal the end of their interface elements.
 * In modes 2 / 3 the exchange progresses only once sfem_halo_wait_unpack has
 * been enqueued on the apply's stream: call it right after sfem_op_apply_halo
 * (ranks sharing ONE stream of one device would wait for one another). */
int sfem_halo_set_option(sfem_halo* halo, int32_t key, int64_t value);

/* Starts a new epoch: u's shared dofs -> the peers' receive buffers, then
 * this rank's flag is raised on every peer.  Never waits. */
int sfem_halo_push(sfem_halo* halo, const void* u, sfem_stream_t stream);

/* Waits (on the device) for the current epoch's flags of all peers, then
 * u[dof] = sum over all holders in ascending rank order -- or whatever part
 * of that sum the fused apply has not already done inside its kernel.  Every
 * rank must run the same sequence of push / wait_unpack calls. */
int sfem_halo_wait_unpack(sfem_halo* halo, void* u, sfem_stream_t stream);

/* First half of y = QQ^T (mask . scatter(local_op(gather(x)))) on an
 * element-partitioned mesh whose first `num_interface_elements` elements are
 * the ones touching other ranks' blocks: ONE apply launch that pushes the
 * shared dofs of y to the peers as soon as all interface elements are done,
 * while the interior elements are still being computed (3-D collocated
 * kernels; other kernel families run the apply, then sfem_halo_push).  CTAs
 * that have pushed poll the peers' flags between element steps and, once all
 * peers' values are in, also run the canonical sum inside the same kernel.
 * The caller completes the exchange with sfem_halo_wait_unpack(halo, y),
 * which waits and does what is left of the sum (usually nothing) -- work that
 * does not touch y's shared dofs may be enqueued in between.  ncomp = 1.
 * dot_xy as in sfem_op_apply (the local, pre-exchange x . y: element-wise
 * partial sums need no ownership weights). */
int sfem_op_apply_halo(const sfem_op* op, sfem_halo* halo, double lambda,
                       double mu, const void* x, void* y,
                       int64_t num_interface_elements, void* dot_xy,
                       sfem_stream_t stream);

/* 1 if the last wait of this handle timed out (peer never raised its flag). */
int sfem_halo_timed_out(const sfem_halo* halo, sfem_stream_t stream);

/* Diagnostics of the last fused apply (host array of 8 globaltimer stamps in
 * ns, synchronises): [0] kernel start, [1] CTA 0 past its interface
 * elements, [2] CTA 0 saw all CTAs past theirs, [3] CTA 0 done pushing,
 * [4] flags raised on the peers, [5] CTA 0 exit, [6] CTA 0 saw all peers'
 * flags, [7] CTA 0 done with its share of the canonical sum. */
int sfem_halo_debug_times(const sfem_halo* halo, uint64_t* out8,
                          sfem_stream_t stream);

/* Peer-memory all-reduce of up to 4 doubles (the dot products of CG, whose
 * hook in the reference is `dot_fn`, swirl_fem/linalg/cg.py:26-31): every
 * rank owns an IPC region of sfem_scalar_region_bytes(world) zeroed bytes;
 * `peer_regions` (HOST, world entries) holds every rank's region as mapped in
 * this process ([rank] = my_region).  sfem_scalar_allreduce replaces
 * values[0..count) by their sums over all ranks, added in ascending rank
 * order (bitwise identical everywhere): ONE single-CTA kernel that stores
 * into the peers' regions, waits on the device for theirs and adds; no NCCL
 * launch, no host synchronisation.  Collective: same call sequence on every
 * rank. */
typedef struct sfem_scalar_exchange sfem_scalar_exchange;
int64_t sfem_scalar_region_bytes(int32_t world);
int sfem_scalar_exchange_create(int32_t rank, int32_t world, void* my_region,
                                const uint64_t* peer_regions,
                                sfem_scalar_exchange** out);
void sfem_scalar_exchange_destroy(sfem_scalar_exchange* h);
int sfem_scalar_allreduce(sfem_scalar_exchange* h, double* values,
                          int32_t count, sfem_stream_t stream);
int sfem_scalar_exchange_timed_out(const sfem_scalar_exchange* h,
                                   sfem_stream_t stream);

/* Fused Stokes divergence and its transpose (StokesSEM.D / Dt,
 * swirl_fem/navier_stokes/navier_stokes.py:313-338): ONE launch each instead
 * of gather -> sfem_space_eval -> sfem_pointwise -> sfem_space_eval_transpose ->
 * scatter.  `vspace`: the continuous GLL velocity space with its collocated
 * rule (created with invjacs and jacdets); `pspace`: the pressure space on the
 * same elements and rule.
 *   sfem_stokes_div:    u (G_v, dim) AoS -> out (G_p,)  (zero-initialised sum)
 *   sfem_stokes_grad_t: p (G_p,) -> out (G_v, dim), rows scaled by mask (G_v,)
 *                       (the interior mask; NULL: none).
 * SFEM_ERR_UNSUPPORTED when N^dim > 1024 (callers keep the composed path). */
int sfem_stokes_div(const sfem_space* vspace, const sfem_space* pspace,
                    const void* u, void* out, sfem_stream_t stream);
int sfem_stokes_grad_t(const sfem_space* vspace, const sfem_space* pspace,
                       const void* p, const void* mask, void* out,
                       sfem_stream_t stream);

/* ------------------------------------------------------------------------ */
/* CG (swirl_fem/linalg/cg.py:30-97)                                         */
/* ------------------------------------------------------------------------ */

typedef struct {
  double tol;          /* relative tolerance (default 1e-5)                  */
  double atol;         /* absolute tolerance                                 */
  int64_t maxiter;     /* < 0: 10 * size (0 = no iteration, as cg.py:68-73)   */
  int32_t precond;     /* 0: identity, 1: Jacobi (minv = 1/diag, 0 on mask)  */
  int32_t check_every; /* host polls the device flag every k iterations      */
  double lambda, mu;   /* operator coefficients                              */
} sfem_cg_params;

typedef struct {
  double residual;        /* gamma = r . M r at exit (cg.py:96)              */
  int64_t num_iterations;
} sfem_cg_info;

/* Bytes of device workspace sfem_cg needs for a system of `size` unknowns. */
int64_t sfem_cg_workspace_bytes(int dtype, int64_t size);

/* Solves A x = b with A = the operator (ncomp components).  x holds x0 on
 * entry (pass zeros for the reference default) and the solution on exit.
 * minv: device (G*ncomp) inverse diagonal for precond == 1, else NULL.
 * Synchronises the stream only to read the convergence flag every
 * `check_every` iterations and at exit. */
int sfem_cg(const sfem_op* op, const void* b, void* x, int32_t ncomp,
            const void* minv, const sfem_cg_params* params, void* workspace,
            sfem_cg_info* info, sfem_stream_t stream);

/* Building blocks of the same fused CG for hosts that interleave collectives
 * (one process per GPU: halo exchange after the apply, NCCL all-reduce of the
 * two scalars).  `state`: device buffer of sfem_cg_state_bytes() bytes whose
 * first four doubles are [0] p.Ap (written by sfem_op_apply's dot_xy),
 * [1] gamma_new, [2] gamma, [3] b.b -- the values a distributed host
 * all-reduces.  `owned`: device uint8 (n), 1 where this rank counts the dof in
 * dot products (NULL: all).  No call synchronises the host except
 * sfem_cg_read.  Order per solve:
 *   Ax = A x0 (+ exchange); sfem_cg_init; [allreduce state[2..3]];
 *   sfem_cg_init_finish; then per iteration: apply(p -> Ap, dot_xy = &state[0])
 *   (+ exchange) [allreduce state[0]]; sfem_cg_update [allreduce state[1]];
 *   sfem_cg_direction; sfem_cg_advance.  Kernels are no-ops once converged. */
int64_t sfem_cg_state_bytes(void);
int sfem_cg_init(int dtype, int64_t n, const void* b, const void* Ax,
                 const void* minv, const uint8_t* owned, void* r, void* p,
                 void* state, double tol, double atol, int64_t maxiter,
                 sfem_stream_t stream);
int sfem_cg_init_finish(void* state, sfem_stream_t stream);
int sfem_cg_update(int dtype, int64_t n, void* x, void* r, const void* p,
                   const void* Ap, const void* minv, const uint8_t* owned,
                   void* state, sfem_stream_t stream);
int sfem_cg_direction(int dtype, int64_t n, const void* r, void* p,
                      const void* minv, void* state, sfem_stream_t stream);
int sfem_cg_advance(void* state, sfem_stream_t stream);
/* `iters` whole iterations of the body of swirl_fem/linalg/cg.py:75-86 with
 * TWO (single GPU) or THREE (partitioned) launches each: the operator apply
 * p -> Ap with p.Ap in its epilogue (halo != NULL: sfem_op_apply_halo +
 * sfem_halo_wait_unpack), then ONE co-resident kernel that does update,
 * direction, both dot-product reductions, the scalar advance of
 * sfem_cg_advance and the zero fill of the next apply.  Partitioned: `sx`
 * all-reduces the two scalars over peer memory INSIDE that kernel (no NCCL
 * call in the iteration; NULL only with halo == NULL).  Kernels are no-ops once
 * the state's flag is set; a peer that never answers sets done = 2. */
int sfem_cg_iterate(const sfem_op* op, sfem_halo* halo,
                    sfem_scalar_exchange* sx, double lambda, double mu,
                    int64_t num_interface_elements, int32_t ncomp, void* x,
                    void* r, void* p, void* Ap, const void* minv,
                    const uint8_t* owned, void* state, int32_t iters,
                    sfem_stream_t stream);
/* Copies the state to the host (synchronises the stream).  *done: 0 running,
 * 1 converged or maxiter reached, 2 a peer-memory wait timed out (fatal). */
int sfem_cg_read(const void* state, sfem_cg_info* info, int32_t* done,
                 sfem_stream_t stream);

/* Fused vector kernels for the generic (callable-A) CG path.  All device. */
/* y = a*x + b*y */
int sfem_axpby(int dtype, int64_t n, double a, const void* x, double b,
               void* y, sfem_stream_t stream);
/* *result = x . y   (result: device fp64 scalar, always double) */
int sfem_dot(int dtype, int64_t n, const void* x, const void* y, void* result,
             sfem_stream_t stream);

#ifdef __cplusplus
} /* extern "C" */
#endif

#endif /* SWIRL_B200_H_ */
