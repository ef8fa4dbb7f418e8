// tcgen05 (3xTF32) implicit-GEMM conv — placeholder until the tensor-core kernel lands; the engine
// asks conv_tc_supported() per op and falls back to the fp32 CUDA-core kernel (conv_simt.cu).
#include "common.cuh"
namespace sb200 {
bool conv_tc_supported(const ConvArgs&) { return false; }
void launch_conv_tc(const ConvArgs& a, cudaStream_t st) { launch_conv_simt(a, st); }
}  // namespace sb200
