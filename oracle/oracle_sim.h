/*
 * ORACLE (test infrastructure, not product code) -- internal declarations.
 *
 * CPU restatement, in double precision, of the reference's env step/reset arithmetic for the
 * hot path named in SURVEY.md section 8.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library.  PARITY UNPINNED: the reference's
 * physics lives in third-party pybullet==1.8.6 (environment.yml:109), which is absent here, and
 * the reference's tests pin no numbers; MobileRobot is restated verbatim
 * from the reference Python (it has no physics), Kuka follows DESIGN.md's physics restatement.
 */
#ifndef ORACLE_SIM_H_
#define ORACLE_SIM_H_

#include <stdint.h>
#include <stddef.h>
#include <vector>
#include "../include/srl_sim.h"

struct MobileEnv {
    double pos[3];
    double target[2][3]; /* target_pos (and second target for the 2-target variant) */
    int current_target;
    int counter;
    int has_bumped;
    uint32_t episode;     /* index of the current episode (Philox reset counter) */
    uint32_t total_steps; /* env steps since creation (Philox action counter)    */
    double ep_ret;
    int ep_len;
};

struct KukaWorld; /* oracle_kuka.cpp */

struct srl_sim {
    int kind;
    int n;
    srl_cfg cfg;
    uint64_t seed;
    int auto_reset;
    std::vector<MobileEnv> mobile;
    KukaWorld* kuka;
    uint64_t launches;
};

void oracle_set_error(const char* fmt, ...);

/* mobile */
int  oracle_mobile_obs_dim(int kind);
int  oracle_mobile_action_dim(const srl_sim* s);
void oracle_mobile_reset_env(srl_sim* s, int i, const double* draws /*nullable, 6 values*/);
void oracle_mobile_obs(const srl_sim* s, int i, float* obs);
void oracle_mobile_step_env(srl_sim* s, int i, const void* actions, const float* noise,
                            float* obs, float* rew, uint8_t* done, float* ep_ret, int32_t* ep_len);

/* kuka */
KukaWorld* oracle_kuka_create(srl_sim* s, const void* blob, size_t bytes);
void oracle_kuka_destroy(KukaWorld* w);
void oracle_kuka_reset_env(srl_sim* s, int i, const double* draws /*nullable, 18 values*/);
void oracle_kuka_obs(const srl_sim* s, int i, float* obs);
void oracle_kuka_step_env(srl_sim* s, int i, const void* actions, const float* noise,
                          float* obs, float* rew, uint8_t* done, float* ep_ret, int32_t* ep_len);
int  oracle_kuka_get_state(srl_sim* s, int field, void* dst, size_t bytes);
int  oracle_kuka_set_state(srl_sim* s, int field, const void* src, size_t bytes);
int  oracle_kuka_scene(const srl_sim* s, int i, void* prims_out /* SrlPrim[SRL_MAX_PRIMS], csrc/render_core.h */);

#endif
