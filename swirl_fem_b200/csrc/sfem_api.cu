// C-ABI glue: handles, argument checking and dispatch to the kernel families.
// See include/swirl_b200.h for the contract of every entry point.

#include <cstdlib>
#include <cstring>
#include <vector>

#include "sfem_common.cuh"

namespace sfem {

static thread_local std::string g_last_error;
std::atomic<int64_t> g_launch_count{0};

void set_error(const std::string& msg) { g_last_error = msg; }

// kernel-family launchers (defined in the other translation units)
template <typename T>
int launch_geom(const SpaceBase&, void*, void*, void*, void*, int, int,
                cudaStream_t);
template <typename T>
int launch_eval(const SpaceBase&, const void*, int, int, const void*, void*,
                cudaStream_t);
template <typename T>
int launch_eval_transpose(const SpaceBase&, const void*, const void*, int,
                          const void*, const void*, void*, cudaStream_t);
template <typename T>
int launch_integrate(const SpaceBase&, const void*, const void*, double*,
                     cudaStream_t);
template <typename T>
int launch_apply_generic(const sfem_op&, double, double, const void*, void*,
                         int, bool, double*, cudaStream_t);
template <typename T>
int launch_diag_generic(const sfem_op&, double, double, void*, cudaStream_t);
template <typename T, int DIM>
int launch_apply_colloc_dim(const sfem_op&, double, double, const void*, void*,
                            int, bool, double*, cudaStream_t);
int pack_connectivity(const sfem_space_desc& desc, int n,
                      const uint8_t* dirichlet, uint32_t* conn,
                      int64_t* n_zero, cudaStream_t stream);

static int ipow(int b, int e) {
  int r = 1;
  for (int i = 0; i < e; ++i) r *= b;
  return r;
}

static int check_desc(const sfem_space_desc* d) {
  SFEM_REQUIRE(d != nullptr, "null descriptor");
  SFEM_REQUIRE(d->dim >= 1 && d->dim <= 3, "dim must be 1, 2 or 3");
  SFEM_REQUIRE(d->n1d >= 2 && d->n1d <= SFEM_MAX_1D, "n1d out of range");
  SFEM_REQUIRE(d->q1d >= 1 && d->q1d <= SFEM_MAX_1D, "q1d out of range");
  SFEM_REQUIRE(d->dtype == SFEM_F32 || d->dtype == SFEM_F64, "bad dtype");
  SFEM_REQUIRE(d->num_elements >= 0 && d->num_nodes >= 0, "negative sizes");
  SFEM_REQUIRE(!d->collocated || d->n1d == d->q1d,
               "collocated space needs q1d == n1d");
  SFEM_REQUIRE(d->interp_1d && d->interp_grad_1d && d->quad_weights_1d,
               "null 1-D tables");
  SFEM_REQUIRE(d->num_elements == 0 || (d->elements && d->node_coords),
               "null mesh arrays");
  return SFEM_OK;
}

int space_base_init(SpaceBase* s, const sfem_space_desc* desc) {
  int rc = check_desc(desc);
  if (rc) return rc;
  s->desc = *desc;
  s->n = ipow(desc->n1d, desc->dim);
  s->q = ipow(desc->q1d, desc->dim);
  const int qn = desc->q1d * desc->n1d;
  std::memcpy(s->h_B, desc->interp_1d, sizeof(double) * qn);
  std::memcpy(s->h_BD, desc->interp_grad_1d, sizeof(double) * qn);
  std::memcpy(s->h_W, desc->quad_weights_1d, sizeof(double) * desc->q1d);
  s->desc.interp_1d = s->h_B;
  s->desc.interp_grad_1d = s->h_BD;
  s->desc.quad_weights_1d = s->h_W;
  const int total = 2 * qn + desc->q1d;
  std::vector<double> t64(total);
  std::vector<float> t32(total);
  for (int i = 0; i < qn; ++i) {
    t64[i] = s->h_B[i];
    t64[qn + i] = s->h_BD[i];
  }
  for (int i = 0; i < desc->q1d; ++i) t64[2 * qn + i] = s->h_W[i];
  for (int i = 0; i < total; ++i) t32[i] = (float)t64[i];
  SFEM_CUDA_CHECK(cudaMalloc(&s->d_tables64, sizeof(double) * total));
  SFEM_CUDA_CHECK(cudaMalloc(&s->d_tables32, sizeof(float) * total));
  SFEM_CUDA_CHECK(cudaMemcpy(s->d_tables64, t64.data(), sizeof(double) * total,
                             cudaMemcpyHostToDevice));
  SFEM_CUDA_CHECK(cudaMemcpy(s->d_tables32, t32.data(), sizeof(float) * total,
                             cudaMemcpyHostToDevice));
  return SFEM_OK;
}

void space_base_free(SpaceBase* s) {
  cudaFree(s->d_tables64);
  cudaFree(s->d_tables32);
  s->d_tables64 = nullptr;
  s->d_tables32 = nullptr;
}

static int ngeom_of(const sfem_space_desc& d, int with_mass) {
  return d.dim * (d.dim + 1) / 2 + (with_mass ? 1 : 0);
}

template <typename T>
static int apply_dispatch(const sfem_op& op, double lambda, double mu,
                          const void* x, void* y, int ncomp, bool local,
                          double* dot_xy, cudaStream_t stream) {
  const sfem_space_desc& d = op.base.desc;
  if (op.variant != 1 && d.collocated && d.n1d <= 16 &&
      (d.dim == 2 || d.dim == 3)) {
    return d.dim == 2 ? launch_apply_colloc_dim<T, 2>(op, lambda, mu, x, y,
                                                      ncomp, local, dot_xy,
                                                      stream)
                      : launch_apply_colloc_dim<T, 3>(op, lambda, mu, x, y,
                                                      ncomp, local, dot_xy,
                                                      stream);
  }
  return launch_apply_generic<T>(op, lambda, mu, x, y, ncomp, local, dot_xy,
                                 stream);
}

// SFEM_PDL=0 (developer switch): plain memsets + ordinary launch instead of the
// zero-fill kernel + programmatic dependent launch of the 3-D apply.
bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("SFEM_PDL");
    return !(e && e[0] == '0');
  }();
  return on;
}

// shared by sfem_op_apply and the CG driver
// `prezeroed`: y[0 .. n_zero) and *dot_xy are already zero (fused CG loop: the
// previous cg_step_kernel did it), so no fill is enqueued.
// True when this launch zero-fills lazily (tables set, supported launch, and
// the operator's lazy launches all on ONE stream: the counters are per handle).
bool lazy_zero_applicable(const sfem_op* op, int ncomp, cudaStream_t stream) {
  const sfem_space_desc& d = op->base.desc;
  if (op->lazy_chunk_ptr == nullptr || ncomp != 1 || op->variant != 0 ||
      !d.collocated || d.dim != 3 || op->fuse != nullptr)
    return false;
  LazyHost* h = op->lazy_host;
  if (h == nullptr) return false;
  std::lock_guard<std::mutex> lock(h->mu);
  if (!h->has_stream) {
    h->stream = stream;
    h->has_stream = true;
  }
  return h->stream == stream;
}

int op_apply_internal(const sfem_op* op, double lambda, double mu,
                      const void* x, void* y, int ncomp, double* dot_xy,
                      cudaStream_t stream, bool prezeroed = false,
                      bool dot_prezeroed = false) {
  const sfem_space_desc& d = op->base.desc;
  SFEM_REQUIRE(lambda == 0.0 || op->with_mass,
               "operator was created without mass factors (with_mass = 0) but "
               "lambda != 0");
  const size_t esz = d.dtype == SFEM_F64 ? 8 : 4;
  // 3-D collocated kernels: zero fill by our own kernel, apply launched as its
  // programmatic dependent (prologue overlaps the fill)
  // Measured (profiles/r02_pdl_on_off.txt): with a SHORT fill (40 MB at 13.6 M
  // dofs) the overlap is worth 1 %; with a LONG one (320 MB at 108 M dofs) the
  // apply's CTAs, all parked at griddepcontrol.wait after their first element,
  // cost 6.7 % -- so the dependent launch is used for fills up to 64 MB only.
  const bool pdl = pdl_enabled() && op->variant == 0 && d.collocated &&
                   d.dim == 3 && d.n1d <= 16 && (op->n_zero > 0 || dot_xy) &&
                   esz * (size_t)op->n_zero * ncomp <= ((size_t)64 << 20);
  sfem_op sub = *op;
  if (prezeroed) {
    sub.pdl = false;
  } else if (lazy_zero_applicable(op, ncomp, stream)) {
    // the companion kernel zeroes y's prefix while the apply runs.  (Not the
    // dot accumulator: nothing orders the companion's first instruction
    // before the apply's last.)
    sub.pdl = false;
    sub.lazy = true;
    if (dot_xy && !dot_prezeroed)
      SFEM_CUDA_CHECK(cudaMemsetAsync(dot_xy, 0, sizeof(double), stream));
  } else if (pdl) {
    int rc = launch_zero_fill(y, esz * (size_t)op->n_zero * ncomp, dot_xy,
                              stream, &sub.pdl);
    if (rc) return rc;
  } else {
    if (op->n_zero > 0)
      SFEM_CUDA_CHECK(cudaMemsetAsync(y, 0, esz * (size_t)op->n_zero * ncomp,
                                      stream));
    if (dot_xy)
      SFEM_CUDA_CHECK(cudaMemsetAsync(dot_xy, 0, sizeof(double), stream));
  }
  return d.dtype == SFEM_F64
             ? apply_dispatch<double>(sub, lambda, mu, x, y, ncomp, false,
                                      dot_xy, stream)
             : apply_dispatch<float>(sub, lambda, mu, x, y, ncomp, false,
                                     dot_xy, stream);
}

}  // namespace sfem

extern "C" {

const char* sfem_last_error(void) { return sfem::g_last_error.c_str(); }

int sfem_version(void) { return 100; }

int64_t sfem_launch_count(void) {
  return sfem::g_launch_count.load(std::memory_order_relaxed);
}

int sfem_space_create(const sfem_space_desc* desc, void* invjacs, void* jacdets,
                      void* quad_coords, sfem_space** space,
                      sfem_stream_t stream) {
  using namespace sfem;
  SFEM_REQUIRE(space != nullptr, "null output handle");
  auto* s = new sfem_space();
  int rc = space_base_init(&s->base, desc);
  if (rc) {
    delete s;
    return rc;
  }
  s->invjacs = invjacs;
  s->jacdets = jacdets;
  s->quad_coords = quad_coords;
  if (invjacs || jacdets || quad_coords) {
    rc = desc->dtype == SFEM_F64
             ? launch_geom<double>(s->base, invjacs, jacdets, quad_coords,
                                   nullptr, 0, 0, (cudaStream_t)stream)
             : launch_geom<float>(s->base, invjacs, jacdets, quad_coords,
                                  nullptr, 0, 0, (cudaStream_t)stream);
    if (rc) {
      space_base_free(&s->base);
      delete s;
      return rc;
    }
  }
  *space = s;
  return SFEM_OK;
}

void sfem_space_destroy(sfem_space* space) {
  if (!space) return;
  sfem::space_base_free(&space->base);
  delete space;
}

int sfem_space_eval(const sfem_space* space, const void* u_local, int32_t ncomp,
                    int32_t kind, void* out, sfem_stream_t stream) {
  using namespace sfem;
  SFEM_REQUIRE(space && u_local && out, "null argument");
  SFEM_REQUIRE(ncomp >= 1 && ncomp <= 65535, "bad ncomp");
  SFEM_REQUIRE(kind == 0 || kind == 1, "kind must be 0 (value) or 1 (grad)");
  SFEM_REQUIRE(kind == 0 || space->invjacs != nullptr,
               "gradient evaluation needs invjacs");
  return space->base.desc.dtype == SFEM_F64
             ? launch_eval<double>(space->base, u_local, ncomp, kind,
                                   space->invjacs, out, (cudaStream_t)stream)
             : launch_eval<float>(space->base, u_local, ncomp, kind,
                                  space->invjacs, out, (cudaStream_t)stream);
}

int sfem_space_eval_transpose(const sfem_space* space, const void* vals,
                              const void* grads, int32_t ncomp, void* out,
                              sfem_stream_t stream) {
  using namespace sfem;
  SFEM_REQUIRE(space && out, "null argument");
  SFEM_REQUIRE(ncomp >= 1 && ncomp <= 65535, "bad ncomp");
  SFEM_REQUIRE(space->jacdets != nullptr, "the covector needs jacdets");
  SFEM_REQUIRE(grads == nullptr || space->invjacs != nullptr,
               "gradient coefficients need invjacs");
  return space->base.desc.dtype == SFEM_F64
             ? launch_eval_transpose<double>(space->base, vals, grads, ncomp,
                                             space->invjacs, space->jacdets,
                                             out, (cudaStream_t)stream)
             : launch_eval_transpose<float>(space->base, vals, grads, ncomp,
                                            space->invjacs, space->jacdets,
                                            out, (cudaStream_t)stream);
}

int sfem_space_integrate(const sfem_space* space, const void* w, void* result,
                         sfem_stream_t stream) {
  using namespace sfem;
  SFEM_REQUIRE(space && result, "null argument");
  SFEM_REQUIRE(space->jacdets != nullptr, "integration needs jacdets");
  return space->base.desc.dtype == SFEM_F64
             ? launch_integrate<double>(space->base, w, space->jacdets,
                                        (double*)result, (cudaStream_t)stream)
             : launch_integrate<float>(space->base, w, space->jacdets,
                                       (double*)result, (cudaStream_t)stream);
}

int64_t sfem_op_geom_bytes(const sfem_space_desc* desc, int32_t with_mass) {
  if (!desc || desc->dim < 1 || desc->dim > 3) return -1;
  const int64_t q = sfem::ipow(desc->q1d, desc->dim);
  const int64_t esz = desc->dtype == SFEM_F64 ? 8 : 4;
  // + 64 B: bulk async copies round a CTA step's chunk up to 16 bytes
  return desc->num_elements * q * sfem::ngeom_of(*desc, with_mass) * esz + 64;
}

int64_t sfem_op_conn_bytes(const sfem_space_desc* desc) {
  if (!desc || desc->dim < 1 || desc->dim > 3) return -1;
  return desc->num_elements * (int64_t)sfem::ipow(desc->n1d, desc->dim) * 4 +
         64;
}

int sfem_op_create(const sfem_space_desc* desc, const uint8_t* dirichlet,
                   int32_t with_mass, void* geom, void* conn, sfem_op** op,
                   sfem_stream_t stream_) {
  using namespace sfem;
  cudaStream_t stream = (cudaStream_t)stream_;
  SFEM_REQUIRE(op != nullptr, "null output handle");
  auto* o = new sfem_op();
  int rc = space_base_init(&o->base, desc);
  if (rc) {
    delete o;
    return rc;
  }
  o->dirichlet = dirichlet;
  o->with_mass = with_mass ? 1 : 0;
  o->ngeom = ngeom_of(*desc, with_mass);
  o->geom = geom;
  o->conn = (uint32_t*)conn;
  o->variant = 0;
  o->n_zero = 0;
  if (desc->num_elements > 0) {
    SFEM_REQUIRE(geom && conn, "null geom/conn buffers");
    rc = desc->dtype == SFEM_F64
             ? launch_geom<double>(o->base, nullptr, nullptr, nullptr, geom,
                                   o->ngeom, o->with_mass, stream)
             : launch_geom<float>(o->base, nullptr, nullptr, nullptr, geom,
                                  o->ngeom, o->with_mass, stream);
  }
  if (!rc)
    rc = pack_connectivity(*desc, o->base.n, dirichlet, o->conn, &o->n_zero,
                           stream);
  if (rc) {
    space_base_free(&o->base);
    delete o;
    return rc;
  }
  *op = o;
  return SFEM_OK;
}

void sfem_op_destroy(sfem_op* op) {
  if (!op) return;
  if (op->lazy_counters) cudaFree(op->lazy_counters);
  delete op->lazy_host;
  sfem::space_base_free(&op->base);
  delete op;
}

int64_t sfem_op_num_zero(const sfem_op* op) { return op ? op->n_zero : -1; }

int sfem_op_lazy_zero_query(const sfem_op* op, int32_t* step_elems,
                            int32_t* grid, int32_t* supported) {
  using namespace sfem;
  SFEM_REQUIRE(op && step_elems && grid && supported, "null argument");
  const sfem_space_desc& d = op->base.desc;
  *step_elems = *grid = *supported = 0;
  if (!d.collocated || d.dim != 3 || d.n1d > 16 || d.num_elements <= 0 ||
      op->variant != 0)
    return SFEM_OK;
  unsigned q[3] = {0, 0, 0};
  sfem_op sub = *op;
  sub.query = q;
  sub.fuse = nullptr;
  sub.lazy = false;
  // (nothing is launched: the launcher returns after sizing its grid)
  const int rc =
      d.dtype == SFEM_F64
          ? apply_dispatch<double>(sub, 0.0, 1.0, op->conn, op->conn, 1, false,
                                   nullptr, nullptr)
          : apply_dispatch<float>(sub, 0.0, 1.0, op->conn, op->conn, 1, false,
                                  nullptr, nullptr);
  if (rc) return rc;
  *step_elems = (int32_t)q[0];
  *grid = (int32_t)q[1];
  *supported = (int32_t)q[2];
  return SFEM_OK;
}

int sfem_op_set_lazy_zero(sfem_op* op, const void* pieces, int64_t num_pieces,
                          const void* chunk_ptr, int32_t num_chunks,
                          int32_t chunk_steps, int32_t ahead_steps) {
  using namespace sfem;
  SFEM_REQUIRE(op != nullptr, "null argument");
  if (chunk_ptr == nullptr) {  // back to the eager fill
    op->lazy_pieces = nullptr;
    op->lazy_chunk_ptr = nullptr;
    op->lazy_num_chunks = 0;
    return SFEM_OK;
  }
  SFEM_REQUIRE(pieces && num_pieces >= 0 && num_pieces < (1ll << 31) &&
                   num_chunks >= 2 && num_chunks < (1 << 19) &&
                   chunk_steps > 0 && ahead_steps >= 0,
               "bad lazy zero tables");
  int32_t se = 0, g = 0, sup = 0;
  int rc = sfem_op_lazy_zero_query(op, &se, &g, &sup);
  if (rc) return rc;
  SFEM_REQUIRE(sup == 1, "lazy zero fill: no kernel instance for this operator");
  const int64_t steps = (op->base.desc.num_elements + se - 1) / se;
  SFEM_REQUIRE((steps + chunk_steps - 1) / chunk_steps == num_chunks,
               "lazy zero fill: num_chunks does not match chunk_steps");
  if (op->lazy_counters) {
    SFEM_CUDA_CHECK(cudaDeviceSynchronize());
    SFEM_CUDA_CHECK(cudaFree(op->lazy_counters));
    op->lazy_counters = nullptr;
  }
  // + 8 words: record of the first wait that timed out (debugging)
  // and the work-queue head + done-CTA count of the companion kernel
  const size_t bytes = ((size_t)num_chunks + 2 + 8 + 2) * sizeof(unsigned);
  SFEM_CUDA_CHECK(cudaMalloc(&op->lazy_counters, bytes));
  SFEM_CUDA_CHECK(cudaMemset(op->lazy_counters, 0, bytes));
  SFEM_CUDA_CHECK(cudaDeviceSynchronize());
  if (!op->lazy_host) op->lazy_host = new LazyHost();
  op->lazy_host->has_stream = false;
  op->lazy_pieces = (const int2*)pieces;
  op->lazy_chunk_ptr = (const int32_t*)chunk_ptr;
  op->lazy_num_chunks = num_chunks;
  op->lazy_num_pieces = (int)num_pieces;
  op->lazy_step_elems = (unsigned)se;
  op->lazy_chunk_steps = (unsigned)chunk_steps;
  op->lazy_ahead = (unsigned)ahead_steps;
  return SFEM_OK;
}

int sfem_op_lazy_zero_timed_out(const sfem_op* op, sfem_stream_t stream) {
  if (!op || !op->lazy_counters) return 0;
  unsigned v = 0, dbg[5] = {0, 0, 0, 0, 0};
  if (cudaMemcpyAsync(&v, op->lazy_counters + 1, sizeof(v),
                      cudaMemcpyDeviceToHost, (cudaStream_t)stream) != cudaSuccess)
    return 1;
  if (cudaMemcpyAsync(dbg, op->lazy_counters + 2 + op->lazy_num_chunks,
                      sizeof(dbg), cudaMemcpyDeviceToHost,
                      (cudaStream_t)stream) != cudaSuccess)
    return 1;
  if (cudaStreamSynchronize((cudaStream_t)stream) != cudaSuccess) return 1;
  if (v != 0) {
    char msg[256];
    snprintf(msg, sizeof(msg),
             "lazy zero fill: first timeout in %s %u at chunk %u (saw %u, "
             "claim counter %u; chunk steps %u, chunks %d)",
             dbg[0] == 1 ? "apply CTA" : "companion CTA", dbg[1], dbg[2],
             dbg[3], dbg[4], op->lazy_chunk_steps, op->lazy_num_chunks);
    sfem::set_error(msg);
  }
  return v != 0;
}

int sfem_op_set_variant(sfem_op* op, int32_t variant) {
  using namespace sfem;
  SFEM_REQUIRE(op && variant >= 0 && variant <= 31, "bad variant");
  op->variant = variant;
  return SFEM_OK;
}

int sfem_op_apply(const sfem_op* op, double lambda, double mu, const void* x,
                  void* y, int32_t ncomp, void* dot_xy, sfem_stream_t stream) {
  using namespace sfem;
  SFEM_REQUIRE(op && x && y, "null argument");
  SFEM_REQUIRE(x != y, "sfem_op_apply is out of place");
  SFEM_REQUIRE(ncomp >= 1 && ncomp <= 65535, "bad ncomp");
  return op_apply_internal(op, lambda, mu, x, y, ncomp, (double*)dot_xy,
                           (cudaStream_t)stream);
}

int sfem_op_apply_range(const sfem_op* op, double lambda, double mu,
                        const void* x, void* y, int32_t ncomp,
                        int64_t elem_begin, int64_t elem_end, int32_t first,
                        void* dot_xy, sfem_stream_t stream_) {
  using namespace sfem;
  cudaStream_t stream = (cudaStream_t)stream_;
  SFEM_REQUIRE(op && x && y, "null argument");
  SFEM_REQUIRE(x != y, "sfem_op_apply_range is out of place");
  SFEM_REQUIRE(ncomp >= 1 && ncomp <= 65535, "bad ncomp");
  const sfem_space_desc& d = op->base.desc;
  SFEM_REQUIRE(elem_begin >= 0 && elem_begin <= elem_end &&
                   elem_end <= d.num_elements,
               "element range out of bounds");
  SFEM_REQUIRE(elem_begin % 4 == 0 || elem_begin == elem_end,
               "elem_begin must be a multiple of 4 (16-byte aligned chunks)");
  SFEM_REQUIRE(lambda == 0.0 || op->with_mass,
               "operator was created without mass factors but lambda != 0");
  const size_t esz = d.dtype == SFEM_F64 ? 8 : 4;
  if (first) {
    if (op->n_zero > 0)
      SFEM_CUDA_CHECK(cudaMemsetAsync(y, 0, esz * (size_t)op->n_zero * ncomp,
                                      stream));
    if (dot_xy)
      SFEM_CUDA_CHECK(cudaMemsetAsync(dot_xy, 0, sizeof(double), stream));
  }
  if (elem_begin == elem_end) return SFEM_OK;
  // a shallow view of the handle restricted to [elem_begin, elem_end)
  sfem_op sub = *op;
  sub.conn = op->conn + elem_begin * (int64_t)op->base.n;
  sub.geom = (char*)op->geom +
             (size_t)elem_begin * op->ngeom * op->base.q * esz;
  sub.base.desc.num_elements = elem_end - elem_begin;
  return d.dtype == SFEM_F64
             ? apply_dispatch<double>(sub, lambda, mu, x, y, ncomp, false,
                                      (double*)dot_xy, stream)
             : apply_dispatch<float>(sub, lambda, mu, x, y, ncomp, false,
                                     (double*)dot_xy, stream);
}

int sfem_op_apply_local(const sfem_op* op, double lambda, double mu,
                        const void* u_local, void* y_local, int32_t ncomp,
                        sfem_stream_t stream) {
  using namespace sfem;
  SFEM_REQUIRE(op && u_local && y_local, "null argument");
  SFEM_REQUIRE(ncomp >= 1 && ncomp <= 65535, "bad ncomp");
  SFEM_REQUIRE(lambda == 0.0 || op->with_mass,
               "operator was created without mass factors but lambda != 0");
  return op->base.desc.dtype == SFEM_F64
             ? apply_dispatch<double>(*op, lambda, mu, u_local, y_local, ncomp,
                                      true, nullptr, (cudaStream_t)stream)
             : apply_dispatch<float>(*op, lambda, mu, u_local, y_local, ncomp,
                                     true, nullptr, (cudaStream_t)stream);
}

int sfem_op_diag(const sfem_op* op, double lambda, double mu, void* diag,
                 sfem_stream_t stream_) {
  using namespace sfem;
  cudaStream_t stream = (cudaStream_t)stream_;
  SFEM_REQUIRE(op && diag, "null argument");
  SFEM_REQUIRE(lambda == 0.0 || op->with_mass,
               "operator was created without mass factors but lambda != 0");
  const sfem_space_desc& d = op->base.desc;
  const size_t esz = d.dtype == SFEM_F64 ? 8 : 4;
  if (op->n_zero > 0)
    SFEM_CUDA_CHECK(cudaMemsetAsync(diag, 0, esz * (size_t)op->n_zero, stream));
  return d.dtype == SFEM_F64
             ? launch_diag_generic<double>(*op, lambda, mu, diag, stream)
             : launch_diag_generic<float>(*op, lambda, mu, diag, stream);
}

}  // extern "C"
