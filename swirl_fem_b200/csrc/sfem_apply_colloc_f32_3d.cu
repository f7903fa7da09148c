// Instantiates the collocated-GLL apply kernels for (float, 3-D), N = 2..16.
#include "sfem_apply_colloc.cuh"

namespace sfem {
template int launch_apply_colloc_dim<float, 3>(const sfem_op&, double, double,
                                               const void*, void*, int, bool,
                                               double*, cudaStream_t);
}  // namespace sfem
