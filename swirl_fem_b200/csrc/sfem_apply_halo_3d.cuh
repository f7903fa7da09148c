// 3-D collocated apply with the halo push fused in (HALO = true instances of
// apply3d_v2_kernel), dispatched on N and on the handle's factor layout.
#pragma once

#include "sfem_apply3d_v2.cuh"

namespace sfem {

template <typename T>
int launch_apply3d_halo(const sfem_op& op, double lambda, double mu,
                        const void* x, void* y, double* dot_xy,
                        cudaStream_t stream) {
  const bool mass = op.with_mass != 0;
  switch (op.base.desc.n1d) {
#define SFEM_CASE(NN)                                                        \
  case NN:                                                                   \
    return mass ? launch3d_v2_halo<T, NN, true>(op, lambda, mu, x, y, dot_xy, \
                                                stream)                      \
                : launch3d_v2_halo<T, NN, false>(op, lambda, mu, x, y,       \
                                                 dot_xy, stream);
    SFEM_CASE(2)
    SFEM_CASE(3)
    SFEM_CASE(4)
    SFEM_CASE(5)
    SFEM_CASE(6)
    SFEM_CASE(7)
    SFEM_CASE(8)
    SFEM_CASE(9)
    SFEM_CASE(10)
    SFEM_CASE(11)
    SFEM_CASE(12)
    SFEM_CASE(13)
    SFEM_CASE(14)
    SFEM_CASE(15)
    SFEM_CASE(16)
#undef SFEM_CASE
    default:
      set_error("fused halo apply: n1d out of range");
      return SFEM_ERR_UNSUPPORTED;
  }
}

}  // namespace sfem
