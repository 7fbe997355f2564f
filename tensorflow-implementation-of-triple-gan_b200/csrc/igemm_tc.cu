// igemm_tc.cu -- placeholder until the tcgen05 kernels land (next commit).
#include "common.cuh"
extern "C" int tgan_igemm_bf16(const tgan_igemm_args*, void*) { tgan::set_error("tgan_igemm_bf16: not built"); return 3; }
extern "C" int tgan_wgrad_bf16(const tgan_wgrad_args*, void*) { tgan::set_error("tgan_wgrad_bf16: not built"); return 3; }
extern "C" int64_t tgan_wgrad_workspace_bytes(const tgan_wgrad_args*) { return 0; }
extern "C" int tgan_pack_weight_bf16(const float*, void*, int, int, int, int, int, const int*, int, void*) { tgan::set_error("tgan_pack_weight_bf16: not built"); return 3; }
