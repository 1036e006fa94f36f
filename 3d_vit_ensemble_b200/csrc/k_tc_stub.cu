// placeholder until the tcgen05 kernels land
#include "tc.cuh"
namespace vit3d {
bool tc_linear_supported(int, int, int, int) { return false; }
int tc_linear_fwd(const TcLinear&, cudaStream_t) { set_error("tc path not built"); return VIT3D_ERR_UNSUPPORTED; }
bool tc_patch_embed_supported(int, int, int, int, int, int, int, int) { return false; }
int tc_patch_embed_fwd(const float*, const float*, const float*, const float*, float*, int, int, int, int, int, int, int, int, cudaStream_t) { return VIT3D_ERR_UNSUPPORTED; }
bool tc_attn_supported(int, int, int) { return false; }
int tc_attn_fwd(const void*, void*, float*, int, int, int, int, cudaStream_t) { return VIT3D_ERR_UNSUPPORTED; }
}
