// tcgen05 implicit-GEMM sparse convolution (placeholder until the tile pipeline lands).
#include "common.cuh"
namespace sps {
bool conv_umma_supports(const sps_conv_args&) { return false; }
int conv_umma(const sps_conv_args&, cudaStream_t) { return SPS_ERR_UNSUPPORTED; }
}
extern "C" int sps_umma_selftest(const void*, const void*, float*, int, int, int, void*) { return SPS_ERR_UNSUPPORTED; }
