// placeholder until the Kuka kernels land
#include "common.cuh"
int kuka_alloc(srl_sim*, const void*, size_t) { srl_set_error("kuka kernels not built yet"); return 1; }
void kuka_free(srl_sim*) {}
int kuka_launch_reset(srl_sim*, const uint8_t*, const double*, float*, cudaStream_t) { return 1; }
int kuka_launch_rollout(srl_sim*, int, const void*, const float*, float*, float*, uint8_t*, float*, int32_t*, cudaStream_t) { return 1; }
int kuka_get_state(srl_sim*, int, void*, size_t) { return 1; }
int kuka_set_state(srl_sim*, int, const void*, size_t) { return 1; }
