// placeholder, replaced below by the tcgen05 kernel
#include "common.cuh"
namespace ttsb {
bool conv_umma_supported(const ConvArgs&) { return false; }
cudaError_t launch_conv_umma(const ConvArgs&, cudaStream_t) { return cudaErrorNotSupported; }
void pack_umma_weights(const float*, int, int, int, int, std::vector<__nv_bfloat16>& out, int& cin_pad, int& cout_pad) { out.clear(); cin_pad = cout_pad = 0; }
}
