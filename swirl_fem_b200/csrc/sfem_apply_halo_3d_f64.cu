// Instantiates the halo-fused 3-D collocated apply kernels for double, N = 2..16.
#include "sfem_apply_halo_3d.cuh"

namespace sfem {
template int launch_apply3d_halo<double>(const sfem_op&, double, double,
                                      const void*, void*, double*,
                                      cudaStream_t);
}  // namespace sfem
