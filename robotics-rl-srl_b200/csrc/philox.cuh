// Philox4x32-10 counter-based RNG (Salmon et al., SC'11), device side.
// One stream per (seed, GLOBAL env index): results are independent of how a batch is sharded
// over GPUs.  Replaces the per-process numpy RandomState of the reference
// (environments/srl_env.py:71-78); exact-seed parity with that RandomState is obtained by the
// host supplying the draws through the `reset_draws` / `noise` arguments of the C-ABI.
//
//   key     = (seed & 0xffffffff, seed >> 32)
//   counter = (env_global_lo, env_global_hi, index, purpose)
//   purpose 0..7 : reset block `purpose` of episode `index`
//   purpose 8    : step-noise draw of env step `index`;  purpose 9 (and 10: words 4-6 of a 7-joint action) : random action of step `index`
#pragma once
#include <stdint.h>

enum { PHILOX_PURPOSE_RESET0 = 0, PHILOX_PURPOSE_NOISE = 8, PHILOX_PURPOSE_ACTION = 9 };

__device__ __forceinline__ uint4 philox4x32_10(uint64_t seed, uint64_t env, uint32_t index, uint32_t purpose) {
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    uint32_t c0 = (uint32_t)env, c1 = (uint32_t)(env >> 32), c2 = index, c3 = purpose;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}

// 53-bit uniform in [0,1); every operation is exact, so this is bit-identical on any IEEE host.
__device__ __forceinline__ double philox_u01(uint32_t a, uint32_t b) {
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0);
}
