/*
 * ORACLE (test infrastructure, not product code).
 *
 * Philox4x32-10 counter-based generator (Salmon et al., "Parallel random numbers: as easy as
 * 1, 2, 3", SC'11), restated from the published algorithm.  The reference repo has no
 * counter-based RNG (its envs draw from a per-process numpy RandomState,
 * environments/srl_env.py:71-78); the batched simulator replaces that stream by one Philox
 * stream per (seed, global env index) so that results are independent of the sharding, and
 * exact-seed parity with the reference's RandomState is obtained by the host supplying the
 * draws (`reset_draws` / `noise` arguments of the C-ABI).
 *
 * Stream layout (must match robotics-rl-srl_b200/csrc/philox.cuh):
 *   key     = (seed & 0xffffffff, seed >> 32)
 *   counter = (env_global_lo, env_global_hi, index, purpose)
 *   purpose 0..7  : reset block `purpose` of episode `index`
 *   purpose 8     : step-noise draw of env step `index` (steps since creation)
 *   purpose 9     : random action of env step `index`
 */
#ifndef ORACLE_PHILOX_H_
#define ORACLE_PHILOX_H_
#include <stdint.h>

enum { PHILOX_PURPOSE_RESET0 = 0, PHILOX_PURPOSE_NOISE = 8, PHILOX_PURPOSE_ACTION = 9 };

static inline void philox4x32_10(uint64_t seed, uint64_t env, uint32_t index, uint32_t purpose,
                                 uint32_t out[4]) {
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    uint32_t c0 = (uint32_t)env, c1 = (uint32_t)(env >> 32), c2 = index, c3 = purpose;
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* 53-bit uniform in [0,1) from two 32-bit words (the construction numpy's legacy
 * random_sample uses: (a>>5, b>>6) -> (a*2^26+b)/2^53). */
static inline double philox_u01(uint32_t a, uint32_t b) {
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0);
}

#endif
