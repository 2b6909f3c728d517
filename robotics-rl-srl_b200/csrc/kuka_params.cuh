// Kernel-parameter block of the Kuka kernels (float32 model + scene + env configuration).
//
// Passed BY VALUE as a __grid_constant__ kernel parameter: it lives in the constant bank, every
// access is warp-uniform, and ptxas folds the values into FFMA/FADD operands (c[0x0][...]) -- no load
// instructions and no registers for the robot model.  This replaces the per-reset asset loading of the
// reference (environments/kuka_gym/kuka.py:60-71, kuka_button_gym_env.py:221-239).
#pragma once
#include <stdint.h>
#include "kuka_model.h"

#define KK_NB 12           // movable bodies: PyBullet joints 0-8, 10, 11, 13
#define KK_ND 13           // + button glider
#define KK_MAXC 4          // contact rows kept per step (KM_SC_MAX_CONTACTS)
// one stored constraint row (contact normal / friction): J[14] at 0, 1 / D, target, W[14] at 16 -- 16-byte groups (eight 128-bit loads per row)
#define KK_ROW_J 0
#define KK_ROW_INVD 14
#define KK_ROW_TGT 15
#define KK_ROW_W 16
#define KK_ROWW 32

struct KukaParams {
    // ---- per body ----
    float org[KK_NB][3];   // joint origin in the parent frame
    float rot[KK_NB][9];   // parent <- child rotation at q = 0, row-major
    float axis[KK_NB][3];  // joint axis, child frame
    float mass[KK_NB];
    float com[KK_NB][3];
    float Ic[KK_NB][6];    // xx xy xz yy yz zz about the COM, body axes
    float damping[KK_NB];
    float lower[KK_NB], upper[KK_NB];
    float kp_dt[KK_NB];    // positionGain / dt
    float kd[KK_NB];
    float maxvel[KK_NB];   // <= 0: no clamp
    float maximp[KK_NB];   // force * dt
    int   tmode[KK_NB];    // 0: IK solution, others: 0 (end_effector_angle, finger_angle are identically 0)
    // saturating PGS sweep (kuka_device.cuh, "the SCALED system"): sigma_i = 2 maximp_i and the products the scaled problem needs
    float sat_sig[KK_NB], sat_isig[KK_NB], sat_isig2[KK_NB];
    float sat_ss[KK_NB * (KK_NB + 1) / 2], sat_iss[KK_NB * (KK_NB + 1) / 2];   // sigma_i sigma_j and its reciprocal, (i, j <= i) packed
    // ---- collision spheres ----
    int   nsph;
    int   sph_body[KM_MAX_SPHERES];
    float sph_c[KM_MAX_SPHERES][3];
    float sph_r[KM_MAX_SPHERES];
    int   sph_min_body;    // lowest body index that carries a sphere
    float sph_reach;       // max over spheres of |centre| + radius: no sphere surface is further from its body origin
    // ---- scene ----
    float base[3];
    float gz, dt, inv_dt;
    int   iters;
    float table_z, txmin, txmax, tymin, tymax;
    float btn_base[3];     // default button base origin
    float glider_z, gl_lo, gl_hi, btn_minv;
    float disc_r, disc_z0, disc_z1, stack_r, stack_top;
    float cdist, mu, erp, kl, ka;
    float ee_init[3];
    float box[6];          // active workspace box (small unless random_target): minx maxx miny maxy minz maxz
    float ikq[4];          // IK target orientation x y z w
    double ik_damp;
    int   ee_body, grip_body;
    float target_h, rand_x, rand_y;
    float btn_idle_imp, btn_kp_dt, btn_kd, btn_target, btn_maximp, lim_maximp, lim_eps;
    int   max_contacts;
    // ---- post-settle snapshot (state after the 500 zero-action steps of reset(), :242-247) ----
    float snap_q[KK_NB], snap_qd[KK_NB], snap_ee[3], snap_qb, snap_qdb;
    // ---- env configuration ----
    int   is_discrete, random_target, force_down, shape_reward, action_repeat, max_steps, auto_reset;
    int   action_joints;   // joint-space actions: use_inverse_kinematics = False (kuka_button_gym_env.py:238, kuka.py:158-161)
    float qinit[7];        // initial arm joint vector (kuka.py:65-66): what joint-space set-points are relative to
    int   two_buttons;     // Kuka2ButtonGymEnv: second button body, goal bookkeeping, IK damping 0.5 (kuka_2button_gym_env.py)
    float two_tgt_z;       // Z_TABLE + BUTTON_DISTANCE_HEIGHT: the z of both two-button targets
    int   moving_button;   // KukaMovingButtonGymEnv: the button slides along y (kuka_moving_button_gym_env.py:109-119)
    float max_distance;
    uint64_t seed, env_offset;
};
