// Explicit instantiations of the position-sliced tensor-core conv kernel (conv_ps.cuh); its own translation unit so that the
// build runs in parallel with the conv_tc parts.
#define CDS_PS_INSTANTIATE
#include "conv_ps.cuh"

namespace cds {
template cudaError_t conv_ps_launch_t<32, false>(const ConvPsLaunch&, const int*, cudaStream_t);
template cudaError_t conv_ps_launch_t<32, true>(const ConvPsLaunch&, const int*, cudaStream_t);
template cudaError_t conv_ps_launch_t<64, false>(const ConvPsLaunch&, const int*, cudaStream_t);
template cudaError_t conv_ps_launch_t<64, true>(const ConvPsLaunch&, const int*, cudaStream_t);
template cudaError_t conv_ps_preload_t<32, false>();
template cudaError_t conv_ps_preload_t<32, true>();
template cudaError_t conv_ps_preload_t<64, false>();
template cudaError_t conv_ps_preload_t<64, true>();
}  // namespace cds
