// placeholder, replaced by the tcgen05 path
#include "regressor.h"
namespace straps {
int tc_create(straps_regressor* r) { r->tc = nullptr; return 0; }
void tc_destroy(straps_regressor* r) {}
int tc_pack(straps_regressor* r, const float* const* conv_w, cudaStream_t st) { return 0; }
int tc_encoder_forward(straps_regressor* r, const float* x, int batch, float* feat, cudaStream_t st) { set_error("tc path not built"); return 3; }
int tc_read_activation(straps_regressor* r, int buf, int batch, float* out, cudaStream_t st) { set_error("tc path not built"); return 3; }
}
