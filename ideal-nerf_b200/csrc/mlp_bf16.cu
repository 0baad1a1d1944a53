// placeholder until the tcgen05 kernel lands
#include "mlp_common.cuh"
namespace inerf {
int mlp_bf16_launch(const MlpArgs&, bool, cudaStream_t) { return fail(INERF_E_UNSUPPORTED, "bf16 MLP mode not built yet"); }
int mlp_bf16_packed_bytes(const InerfNetDims*, size_t* bytes) { *bytes = 0; return fail(INERF_E_UNSUPPORTED, "bf16 MLP mode not built yet"); }
int mlp_bf16_pack(const InerfNetDims*, const float* const*, void*, cudaStream_t) { return fail(INERF_E_UNSUPPORTED, "bf16 MLP mode not built yet"); }
}
