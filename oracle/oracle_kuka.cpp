/* ORACLE (test infrastructure) -- placeholder until the Kuka restatement lands. */
#include "oracle_sim.h"
struct KukaWorld { int dummy; };
KukaWorld* oracle_kuka_create(srl_sim*, const void*, size_t) { oracle_set_error("kuka oracle not built yet"); return NULL; }
void oracle_kuka_destroy(KukaWorld* w) { delete w; }
void oracle_kuka_reset_env(srl_sim*, int, const double*) {}
void oracle_kuka_obs(const srl_sim*, int, float*) {}
void oracle_kuka_step_env(srl_sim*, int, const void*, const float*, float*, float*, uint8_t*, float*, int32_t*) {}
int oracle_kuka_get_state(srl_sim*, int, void*, size_t) { return 1; }
int oracle_kuka_set_state(srl_sim*, int, const void*, size_t) { return 1; }
