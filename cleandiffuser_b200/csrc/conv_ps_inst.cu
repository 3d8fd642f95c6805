// Explicit instantiations of the position-sliced tensor-core conv kernel (conv_ps.cuh); its own translation unit so that the
// build runs in parallel with the conv_tc parts.
#define CDS_PS_INSTANTIATE
#include "conv_ps.cuh"

namespace cds {
#define CDS_PS_INST(N_, R_, T_)                                                                            \
  template cudaError_t conv_ps_launch_t<N_, R_, T_>(const ConvPsLaunch&, const int*, cudaStream_t);        \
  template cudaError_t conv_ps_preload_t<N_, R_, T_>();
CDS_PS_VARIANTS(CDS_PS_INST)
#undef CDS_PS_INST
}  // namespace cds
