/*
 * ORACLE (test infrastructure, not product code) -- Kuka button-push family in double precision.
 *
 * PARITY UNPINNED at the physics-engine boundary: the reference delegates the arithmetic of
 * `p.calculateInverseKinematics`, `p.setJointMotorControl2` and `p.stepSimulation()`
 * (environments/kuka_gym/kuka.py:144-187, environments/kuka_gym/kuka_button_gym_env.py:351) to
 * third-party pybullet==1.8.6 (environment.yml:109), which is absent from /root/reference and from
 * this image, as are the pybullet_data assets it loads.  What follows restates the published
 * algorithms Bullet's btMultiBody world uses (DESIGN.md "Physics restatement" lists every
 * semantic and its source), deliberately in a DIFFERENT formulation from the CUDA kernels:
 *
 *   oracle (this file)                              CUDA kernels (csrc/kuka_kernels.cu)
 *   ---------------------------------------------   -------------------------------------------
 *   Featherstone ABA in link coordinates, O(n)      world-frame CRBA + RNEA bias + Cholesky M^-1
 *   M^-1 J^T per row by ABA impulse response        explicit M^-1 (registers), joint-space rows
 *   PGS on accumulated delta-velocities (Bullet's   PGS on total velocities, fused multiply-add
 *     resolveSingleConstraintRowGeneric form)
 *   float64                                         float32
 *
 * so that agreement between the two is a meaningful check, not a tautology.
 *
 * Env-level logic follows the reference line by line:
 *   reset   kuka_button_gym_env.py:214-281     step/step2  :293-368
 *   _reward :428-463     _termination :422-426     Kuka.applyAction  kuka.py:118-187
 */
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <vector>
#include "oracle_sim.h"
#include "../robotics-rl-srl_b200/csrc/render_core.h"
#include "philox.h"
#include "../robotics-rl-srl_b200/csrc/kuka_model.h"

namespace {

const int NB = KM_NBODY;     /* 12 movable bodies */
const int ND = KM_NBODY + 2; /* + button glider(s): DoF NB = button 1, NB + 1 = button 2 (Kuka2ButtonGymEnv only; inert otherwise) */

/* ---------------------------------------------------------------- small linear algebra ---- */
struct V3 { double x, y, z; };
static inline V3 v3(double x, double y, double z) { V3 r = {x, y, z}; return r; }
static inline V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline V3 operator*(double s, V3 a) { return v3(s * a.x, s * a.y, s * a.z); }
static inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline V3 cross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
static inline double norm(V3 a) { return sqrt(dot(a, a)); }

struct M3 { double m[3][3]; };
static inline M3 m3_identity() { M3 r = {{{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}}; return r; }
static inline V3 operator*(const M3& A, V3 v) {
    return v3(A.m[0][0] * v.x + A.m[0][1] * v.y + A.m[0][2] * v.z, A.m[1][0] * v.x + A.m[1][1] * v.y + A.m[1][2] * v.z,
              A.m[2][0] * v.x + A.m[2][1] * v.y + A.m[2][2] * v.z);
}
static inline M3 operator*(const M3& A, const M3& B) {
    M3 r;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
        double s = 0; for (int k = 0; k < 3; ++k) s += A.m[i][k] * B.m[k][j];
        r.m[i][j] = s;
    }
    return r;
}
static inline M3 transpose(const M3& A) { M3 r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[i][j] = A.m[j][i]; return r; }
static inline M3 skew(V3 a) { M3 r = {{{0, -a.z, a.y}, {a.z, 0, -a.x}, {-a.y, a.x, 0}}}; return r; }
/* Rodrigues: rotation by angle q about unit axis a */
static M3 axis_angle(V3 a, double q) {
    const double c = cos(q), s = sin(q), t = 1 - c;
    M3 r = {{{c + t * a.x * a.x, t * a.x * a.y - s * a.z, t * a.x * a.z + s * a.y},
             {t * a.x * a.y + s * a.z, c + t * a.y * a.y, t * a.y * a.z - s * a.x},
             {t * a.x * a.z - s * a.y, t * a.y * a.z + s * a.x, c + t * a.z * a.z}}};
    return r;
}

/* spatial 6-vectors, Featherstone ordering [angular; linear] */
struct S6 { double v[6]; };
static inline S6 s6(V3 w, V3 l) { S6 r = {{w.x, w.y, w.z, l.x, l.y, l.z}}; return r; }
static inline V3 ang(const S6& a) { return v3(a.v[0], a.v[1], a.v[2]); }
static inline V3 lin(const S6& a) { return v3(a.v[3], a.v[4], a.v[5]); }
static inline S6 operator+(const S6& a, const S6& b) { S6 r; for (int i = 0; i < 6; ++i) r.v[i] = a.v[i] + b.v[i]; return r; }
static inline S6 operator-(const S6& a, const S6& b) { S6 r; for (int i = 0; i < 6; ++i) r.v[i] = a.v[i] - b.v[i]; return r; }
static inline S6 operator*(double s, const S6& a) { S6 r; for (int i = 0; i < 6; ++i) r.v[i] = s * a.v[i]; return r; }
static inline double dot6(const S6& a, const S6& b) { double s = 0; for (int i = 0; i < 6; ++i) s += a.v[i] * b.v[i]; return s; }
struct M6 { double m[6][6]; };
static inline S6 operator*(const M6& A, const S6& x) {
    S6 r; for (int i = 0; i < 6; ++i) { double s = 0; for (int j = 0; j < 6; ++j) s += A.m[i][j] * x.v[j]; r.v[i] = s; } return r;
}
/* motion cross product  v x m  and force cross product  v x* f */
static inline S6 crm(const S6& v, const S6& m) { return s6(cross(ang(v), ang(m)), cross(ang(v), lin(m)) + cross(lin(v), ang(m))); }
static inline S6 crf(const S6& v, const S6& f) { return s6(cross(ang(v), ang(f)) + cross(lin(v), lin(f)), cross(ang(v), lin(f))); }

/* Pluecker transform A -> B: E rotates A-coordinates into B-coordinates, r = origin of B in A-coordinates */
struct Xf { M3 E; V3 r; };
static inline S6 xf_motion(const Xf& X, const S6& v) { return s6(X.E * ang(v), X.E * (lin(v) - cross(X.r, ang(v)))); }
/* X^T applied to a force expressed in B: result in A */
static inline S6 xf_force_T(const Xf& X, const S6& f) {
    const M3 Et = transpose(X.E);
    const V3 fl = Et * lin(f);
    return s6(Et * ang(f) + cross(X.r, fl), fl);
}
static M6 xf_matrix(const Xf& X) { /* 6x6 motion transform [E 0; -E rx E] */
    M6 R; memset(&R, 0, sizeof(R));
    const M3 Erx = X.E * skew(X.r);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
        R.m[i][j] = X.E.m[i][j];
        R.m[i + 3][j + 3] = X.E.m[i][j];
        R.m[i + 3][j] = -Erx.m[i][j];
    }
    return R;
}
static M6 xt_I_x(const M6& X, const M6& I) { /* X^T I X */
    M6 T, R;
    for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) { double s = 0; for (int k = 0; k < 6; ++k) s += I.m[i][k] * X.m[k][j]; T.m[i][j] = s; }
    for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) { double s = 0; for (int k = 0; k < 6; ++k) s += X.m[k][i] * T.m[k][j]; R.m[i][j] = s; }
    return R;
}

/* ------------------------------------------------------------------------------- model ---- */
struct KBody {
    int parent, jtype;
    V3 origin; M3 rot; V3 axis;
    double mass; V3 com; M3 Ic;
    double lower, upper, damping, qinit;
    double kp, kd, maxforce, maxvel; int target_mode;
    M6 I; /* spatial inertia in the body frame */
};
struct KSphere { int body; V3 c; double r; };
struct KModel {
    KBody b[NB];
    int nsphere; KSphere sph[KM_MAX_SPHERES];
    double sc[KM_SCENE_SIZE];
};

static bool parse_model(const void* blob, size_t bytes, KModel& m) {
    const double* d = (const double*)blob;
    if (!blob || bytes < KM_HEADER_SIZE * sizeof(double) || d[KM_H_MAGIC] != KM_MAGIC || d[KM_H_VERSION] != KM_VERSION) {
        oracle_set_error("kuka: bad model blob (magic/version)"); return false;
    }
    if ((size_t)d[KM_H_TOTAL] * sizeof(double) != bytes || (int)d[KM_H_NBODY] != NB) {
        oracle_set_error("kuka: bad model blob (size/body count)"); return false;
    }
    m.nsphere = (int)d[KM_H_NSPHERE];
    if (m.nsphere > KM_MAX_SPHERES) { oracle_set_error("kuka: too many spheres"); return false; }
    for (int i = 0; i < NB; ++i) {
        const double* r = d + (int)d[KM_H_BODY_OFF] + i * KM_BODY_STRIDE;
        const double* c = d + (int)d[KM_H_CTRL_OFF] + i * KM_CTRL_STRIDE;
        KBody& b = m.b[i];
        b.parent = (int)r[KM_B_PARENT]; b.jtype = (int)r[KM_B_JTYPE];
        b.origin = v3(r[KM_B_ORIGIN], r[KM_B_ORIGIN + 1], r[KM_B_ORIGIN + 2]);
        for (int a = 0; a < 3; ++a) for (int k = 0; k < 3; ++k) b.rot.m[a][k] = r[KM_B_ROT + 3 * a + k];
        b.axis = v3(r[KM_B_AXIS], r[KM_B_AXIS + 1], r[KM_B_AXIS + 2]);
        b.mass = r[KM_B_MASS];
        b.com = v3(r[KM_B_COM], r[KM_B_COM + 1], r[KM_B_COM + 2]);
        const double* I = r + KM_B_INERTIA;
        M3 Ic = {{{I[0], I[1], I[2]}, {I[1], I[3], I[4]}, {I[2], I[4], I[5]}}};
        b.Ic = Ic;
        b.lower = r[KM_B_LOWER]; b.upper = r[KM_B_UPPER]; b.damping = r[KM_B_DAMPING]; b.qinit = r[KM_B_QINIT];
        b.kp = c[KM_C_KP]; b.kd = c[KM_C_KD]; b.maxforce = c[KM_C_MAXFORCE]; b.maxvel = c[KM_C_MAXVEL]; b.target_mode = (int)c[KM_C_TARGET];
        /* spatial inertia about the body-frame origin: [Ic + m cx cx^T, m cx; m cx^T, m 1] */
        const M3 cx = skew(b.com);
        const M3 cxcxT = cx * transpose(cx);
        memset(&b.I, 0, sizeof(b.I));
        for (int a = 0; a < 3; ++a) for (int k = 0; k < 3; ++k) {
            b.I.m[a][k] = Ic.m[a][k] + b.mass * cxcxT.m[a][k];
            b.I.m[a][k + 3] = b.mass * cx.m[a][k];
            b.I.m[a + 3][k] = b.mass * cx.m[k][a];
            b.I.m[a + 3][k + 3] = (a == k) ? b.mass : 0.0;
        }
        if (b.jtype != 0) { oracle_set_error("kuka: only revolute arm joints are supported"); return false; }
    }
    for (int k = 0; k < m.nsphere; ++k) {
        const double* s = d + (int)d[KM_H_SPHERE_OFF] + k * KM_SPHERE_STRIDE;
        m.sph[k].body = (int)s[KM_S_BODY];
        m.sph[k].c = v3(s[KM_S_CENTER], s[KM_S_CENTER + 1], s[KM_S_CENTER + 2]);
        m.sph[k].r = s[KM_S_RADIUS];
    }
    memcpy(m.sc, d + (int)d[KM_H_SCENE_OFF], sizeof(m.sc));
    return true;
}

/* --------------------------------------------------------------------------------- env ---- */
struct KEnv {
    double q[NB], qd[NB];
    double qb, qdb;          /* button glider */
    double ee[3], ee_angle;  /* commanded end-effector pose, kuka.py:73-74 */
    double button_base[3];   /* button base link origin */
    double button_pos[3];    /* target, frozen at reset (:273-274); y slides in the moving-button variant */
    double btn_speed;        /* kuka_moving_button_gym_env.py:33 */
    int counter, n_contacts, n_outside, terminated;
    int contact_button, contact_table; /* manifold flags of the last stepSimulation */
    /* ---- Kuka2ButtonGymEnv (kuka_2button_gym_env.py): the second button body and the goal bookkeeping ---- */
    int nbuttons;            /* 1, or 2 for the two-button kind */
    double qb2, qdb2;        /* second button glider */
    double button2_base[3];
    int contact_button_any[2]; /* contact with ANY link of button b (getContactPoints without a link index, :165) */
    int n_contacts2[2], goal_id, pressed[2];
    double gripper_pos[3], ee_pos[3];  /* link states after the last stepSimulation */
    uint32_t episode, total_steps;
    double ep_ret; int ep_len;
};

struct Kin { /* forward kinematics of one configuration */
    M3 R[NB]; V3 p[NB]; V3 a[NB]; V3 com[NB];
};

static void forward_kinematics(const KModel& m, const double* q, Kin& k) {
    const V3 base = v3(m.sc[KM_SC_BASE_POS], m.sc[KM_SC_BASE_POS + 1], m.sc[KM_SC_BASE_POS + 2]);
    for (int i = 0; i < NB; ++i) {
        const KBody& b = m.b[i];
        const M3 Rp = b.parent < 0 ? m3_identity() : k.R[b.parent];
        const V3 pp = b.parent < 0 ? base : k.p[b.parent];
        k.p[i] = pp + Rp * b.origin;
        k.R[i] = Rp * b.rot * axis_angle(b.axis, q[i]);
        k.a[i] = k.R[i] * b.axis;
        k.com[i] = k.p[i] + k.R[i] * b.com;
    }
}

/* ------------------------------------------------------------------------------------ ABA -- */
struct Aba { /* configuration-dependent quantities shared by the dynamics pass and the M^-1 solves */
    Xf Xup[NB];
    S6 S[NB], U[NB];
    double d[NB];
    M6 IA[NB];
};

static void aba_setup(const KModel& m, const double* q, Aba& A) {
    for (int i = 0; i < NB; ++i) {
        const KBody& b = m.b[i];
        /* parent frame -> child frame: rotate by (rot * Rq)^T, child origin at `origin` in the parent */
        A.Xup[i].E = transpose(b.rot * axis_angle(b.axis, q[i]));
        A.Xup[i].r = b.origin;
        A.S[i] = s6(b.axis, v3(0, 0, 0));
        A.IA[i] = b.I;
    }
    for (int i = NB - 1; i >= 0; --i) {
        A.U[i] = A.IA[i] * A.S[i];
        A.d[i] = dot6(A.S[i], A.U[i]);
        const int p = m.b[i].parent;
        if (p >= 0) {
            M6 Ia = A.IA[i];
            for (int r = 0; r < 6; ++r) for (int c = 0; c < 6; ++c) Ia.m[r][c] -= A.U[i].v[r] * A.U[i].v[c] / A.d[i];
            const M6 Xm = xf_matrix(A.Xup[i]);
            const M6 add = xt_I_x(Xm, Ia);
            for (int r = 0; r < 6; ++r) for (int c = 0; c < 6; ++c) A.IA[p].m[r][c] += add.m[r][c];
        }
    }
}

/* qdd = M^-1 tau for zero velocity and zero gravity (Bullet: calcAccelerationDeltasMultiDof) */
static void aba_minv(const KModel& m, const Aba& A, const double* tau, double* qdd) {
    S6 pA[NB]; double u[NB]; S6 acc[NB];
    for (int i = 0; i < NB; ++i) memset(&pA[i], 0, sizeof(S6));
    for (int i = NB - 1; i >= 0; --i) {
        u[i] = tau[i] - dot6(A.S[i], pA[i]);
        const int p = m.b[i].parent;
        if (p >= 0) {
            const S6 pa = pA[i] + (u[i] / A.d[i]) * A.U[i];
            pA[p] = pA[p] + xf_force_T(A.Xup[i], pa);
        }
    }
    for (int i = 0; i < NB; ++i) {
        const int p = m.b[i].parent;
        S6 ap; memset(&ap, 0, sizeof(ap));
        if (p >= 0) ap = xf_motion(A.Xup[i], acc[p]);
        qdd[i] = (u[i] - dot6(A.U[i], ap)) / A.d[i];
        acc[i] = ap + qdd[i] * A.S[i];
    }
}

/* full forward dynamics: gravity, velocity products, joint damping torque, Bullet link damping */
static void aba_forward_dynamics(const KModel& m, const Aba& A, const double* qd, double* qdd) {
    S6 v[NB], c[NB], pA[NB], acc[NB]; double u[NB];
    const double kl = m.sc[KM_SC_LIN_DAMPING], ka = m.sc[KM_SC_ANG_DAMPING];
    for (int i = 0; i < NB; ++i) {
        const KBody& b = m.b[i];
        const int p = b.parent;
        S6 vp; memset(&vp, 0, sizeof(vp));
        if (p >= 0) vp = xf_motion(A.Xup[i], v[p]);
        const S6 vj = qd[i] * A.S[i];
        v[i] = vp + vj;
        c[i] = crm(v[i], vj);
        pA[i] = crf(v[i], b.I * v[i]);
        /* btMultiBody damping terms (m_linearDamping = m_angularDamping = 0.04, K1 = K2):
           F = -kl m v_com (1 + |v_com|),  T = -ka Ic w (1 + |w|), added to the bias force */
        const V3 w = ang(v[i]);
        const V3 vcom = lin(v[i]) + cross(w, b.com);
        const V3 F = (kl * b.mass * (1.0 + norm(vcom))) * vcom;
        const V3 T = (ka * (1.0 + norm(w))) * (b.Ic * w);
        pA[i] = pA[i] + s6(T + cross(b.com, F), F);
    }
    M6 dummy; (void)dummy;
    for (int i = NB - 1; i >= 0; --i) {
        const KBody& b = m.b[i];
        const double tau = -b.damping * qd[i]; /* URDF joint damping applied as an explicit joint torque */
        u[i] = tau - dot6(A.S[i], pA[i]);
        const int p = b.parent;
        if (p >= 0) {
            M6 Ia = A.IA[i];
            for (int r = 0; r < 6; ++r) for (int cc = 0; cc < 6; ++cc) Ia.m[r][cc] -= A.U[i].v[r] * A.U[i].v[cc] / A.d[i];
            const S6 pa = pA[i] + Ia * c[i] + (u[i] / A.d[i]) * A.U[i];
            pA[p] = pA[p] + xf_force_T(A.Xup[i], pa);
        }
    }
    /* gravity as a fictitious base acceleration a0 = -g (base orientation is the identity) */
    const S6 a0 = s6(v3(0, 0, 0), v3(0, 0, -m.sc[KM_SC_GRAVITY_Z]));
    for (int i = 0; i < NB; ++i) {
        const int p = m.b[i].parent;
        const S6 ap = xf_motion(A.Xup[i], p >= 0 ? acc[p] : a0) + c[i];
        qdd[i] = (u[i] - dot6(A.U[i], ap)) / A.d[i];
        acc[i] = ap + qdd[i] * A.S[i];
    }
}

/* ------------------------------------------------------------------------ inverse kinematics */
static void quat_from_matrix(const M3& R, double q[4]) { /* (x, y, z, w) */
    const double tr = R.m[0][0] + R.m[1][1] + R.m[2][2];
    if (tr > 0) {
        const double s = sqrt(tr + 1.0) * 2;
        q[3] = 0.25 * s; q[0] = (R.m[2][1] - R.m[1][2]) / s; q[1] = (R.m[0][2] - R.m[2][0]) / s; q[2] = (R.m[1][0] - R.m[0][1]) / s;
    } else if (R.m[0][0] > R.m[1][1] && R.m[0][0] > R.m[2][2]) {
        const double s = sqrt(1.0 + R.m[0][0] - R.m[1][1] - R.m[2][2]) * 2;
        q[3] = (R.m[2][1] - R.m[1][2]) / s; q[0] = 0.25 * s; q[1] = (R.m[0][1] + R.m[1][0]) / s; q[2] = (R.m[0][2] + R.m[2][0]) / s;
    } else if (R.m[1][1] > R.m[2][2]) {
        const double s = sqrt(1.0 + R.m[1][1] - R.m[0][0] - R.m[2][2]) * 2;
        q[3] = (R.m[0][2] - R.m[2][0]) / s; q[0] = (R.m[0][1] + R.m[1][0]) / s; q[1] = 0.25 * s; q[2] = (R.m[1][2] + R.m[2][1]) / s;
    } else {
        const double s = sqrt(1.0 + R.m[2][2] - R.m[0][0] - R.m[1][1]) * 2;
        q[3] = (R.m[1][0] - R.m[0][1]) / s; q[0] = (R.m[0][2] + R.m[2][0]) / s; q[1] = (R.m[1][2] + R.m[2][1]) / s; q[2] = 0.25 * s;
    }
}

/* One damped-least-squares iteration at the CURRENT joint state (pybullet 1.8.6
   calculateInverseKinematics with orientation + jointDamping -> BussIK CalcDeltaThetasDLS2):
   dtheta = (J^T J + diag(damping))^-1 J^T e,  e = [p_target - p_ee ; angle * axis of q_t (x) q_cur^-1],
   step scaled back if max |dtheta| exceeds 45 degrees.  Returns the 7 arm joint targets. */
static void inverse_kinematics(const KModel& m, const Kin& k, const double* q, const double* target_pos, double* q_ik) {
    const int ee = (int)m.sc[KM_SC_EE_BODY];
    const int n = ee + 1;
    double J[6][NB];
    for (int j = 0; j < n; ++j) {
        const V3 l = cross(k.a[j], k.p[ee] - k.p[j]);
        J[0][j] = l.x; J[1][j] = l.y; J[2][j] = l.z;
        J[3][j] = k.a[j].x; J[4][j] = k.a[j].y; J[5][j] = k.a[j].z;
    }
    double e[6];
    e[0] = target_pos[0] - k.p[ee].x; e[1] = target_pos[1] - k.p[ee].y; e[2] = target_pos[2] - k.p[ee].z;
    double qc[4];
    quat_from_matrix(k.R[ee], qc);
    const double* qt = m.sc + KM_SC_IK_QUAT;
    /* dq = qt (x) conj(qc) */
    const double cx = -qc[0], cy = -qc[1], cz = -qc[2], cw = qc[3];
    const double dx = qt[3] * cx + qt[0] * cw + qt[1] * cz - qt[2] * cy;
    const double dy = qt[3] * cy - qt[0] * cz + qt[1] * cw + qt[2] * cx;
    const double dz = qt[3] * cz + qt[0] * cy - qt[1] * cx + qt[2] * cw;
    const double dw = qt[3] * cw - qt[0] * cx - qt[1] * cy - qt[2] * cz;
    const double vn = sqrt(dx * dx + dy * dy + dz * dz);
    double angle = 2.0 * atan2(vn, dw); /* == btQuaternion::getAngle() = 2 acos(w) for a unit quaternion */
    if (angle > M_PI) angle -= 2.0 * M_PI;
    if (vn > 1e-12) { e[3] = angle * dx / vn; e[4] = angle * dy / vn; e[5] = angle * dz / vn; }
    else { e[3] = e[4] = e[5] = 0.0; }
    /* normal equations */
    double A[NB][NB + 1];
    for (int i = 0; i < n; ++i) {
        for (int j = 0; j < n; ++j) {
            double s = 0; for (int r = 0; r < 6; ++r) s += J[r][i] * J[r][j];
            A[i][j] = s + (i == j ? m.sc[KM_SC_IK_DAMPING] : 0.0);
        }
        double s = 0; for (int r = 0; r < 6; ++r) s += J[r][i] * e[r];
        A[i][n] = s;
    }
    /* Gaussian elimination with partial pivoting */
    for (int c = 0; c < n; ++c) {
        int piv = c;
        for (int r = c + 1; r < n; ++r) if (fabs(A[r][c]) > fabs(A[piv][c])) piv = r;
        if (piv != c) for (int j = 0; j <= n; ++j) { double t = A[c][j]; A[c][j] = A[piv][j]; A[piv][j] = t; }
        for (int r = c + 1; r < n; ++r) {
            const double f = A[r][c] / A[c][c];
            for (int j = c; j <= n; ++j) A[r][j] -= f * A[c][j];
        }
    }
    double dth[NB];
    for (int r = n - 1; r >= 0; --r) {
        double s = A[r][n];
        for (int j = r + 1; j < n; ++j) s -= A[r][j] * dth[j];
        dth[r] = s / A[r][r];
    }
    double mx = 0; for (int j = 0; j < n; ++j) mx = fmax(mx, fabs(dth[j]));
    const double max_angle = 45.0 * M_PI / 180.0; /* BussIK MaxAngleDLS */
    const double scale = mx > max_angle ? max_angle / mx : 1.0;
    for (int j = 0; j < n; ++j) q_ik[j] = q[j] + scale * dth[j];
}

/* ----------------------------------------------------------------------- collision detection */
struct Contact {
    int sphere, body, shape; /* shape: 0 table, 1 button disc (button_uid link 1), 2 button base stack, 3 / 4 the same of button 2 */
    double dist; V3 n, p;    /* signed distance, normal from the shape towards the arm, point on the sphere surface */
};

/* sphere vs upright finite cylinder (axis +z through (cx, cy), z in [z0, z1], radius R) */
static void sphere_cylinder(V3 s, double r, double cx, double cy, double z0, double z1, double R, double& dist, V3& n) {
    const double dx = s.x - cx, dy = s.y - cy;
    const double rho = sqrt(dx * dx + dy * dy);
    const V3 radial = rho > 1e-12 ? v3(dx / rho, dy / rho, 0) : v3(1, 0, 0);
    double d;
    if (s.z >= z1) {
        if (rho <= R) { d = s.z - z1; n = v3(0, 0, 1); }
        else { const V3 vec = v3(dx - radial.x * R, dy - radial.y * R, s.z - z1); d = norm(vec); n = (1.0 / d) * vec; }
    } else if (s.z <= z0) {
        if (rho <= R) { d = z0 - s.z; n = v3(0, 0, -1); }
        else { const V3 vec = v3(dx - radial.x * R, dy - radial.y * R, s.z - z0); d = norm(vec); n = (1.0 / d) * vec; }
    } else {
        if (rho > R) { d = rho - R; n = radial; }
        else {
            const double d_top = z1 - s.z, d_side = R - rho;
            if (d_top <= d_side) { d = -d_top; n = v3(0, 0, 1); } else { d = -d_side; n = radial; }
        }
    }
    dist = d - r;
}

static int detect_contacts(const KModel& m, const Kin& k, const KEnv& e, Contact* out, int max_out, int& button_flag, int& table_flag) {
    const double thr = m.sc[KM_SC_CONTACT_DIST];
    const double zt = m.sc[KM_SC_TABLE_TOP_Z];
    const double bz = e.button_base[2];
    const double disc0 = bz + m.sc[KM_SC_GLIDER_Z] + e.qb + m.sc[KM_SC_DISC_Z0];
    const double disc1 = bz + m.sc[KM_SC_GLIDER_Z] + e.qb + m.sc[KM_SC_DISC_Z1];
    int n = 0;
    button_flag = 0; table_flag = 0;
    const int nshapes = e.nbuttons == 2 ? 5 : 3;
    const double b2z = e.button2_base[2];
    const double disc20 = b2z + m.sc[KM_SC_GLIDER_Z] + e.qb2 + m.sc[KM_SC_DISC_Z0];
    const double disc21 = b2z + m.sc[KM_SC_GLIDER_Z] + e.qb2 + m.sc[KM_SC_DISC_Z1];
    int any[2] = {0, 0};
    for (int s = 0; s < m.nsphere; ++s) {
        const int b = m.sph[s].body;
        const V3 c = k.p[b] + k.R[b] * m.sph[s].c;
        const double r = m.sph[s].r;
        for (int shape = 0; shape < nshapes; ++shape) {
            double dist; V3 nn;
            if (shape == 0) {
                if (c.x < m.sc[KM_SC_TABLE_XMIN] || c.x > m.sc[KM_SC_TABLE_XMAX] || c.y < m.sc[KM_SC_TABLE_YMIN] || c.y > m.sc[KM_SC_TABLE_YMAX]) continue;
                dist = c.z - zt - r; nn = v3(0, 0, 1);
            } else if (shape == 1) {
                sphere_cylinder(c, r, e.button_base[0], e.button_base[1], disc0, disc1, m.sc[KM_SC_DISC_RADIUS], dist, nn);
            } else if (shape == 2) {
                sphere_cylinder(c, r, e.button_base[0], e.button_base[1], bz, bz + m.sc[KM_SC_STACK_TOP], m.sc[KM_SC_STACK_RADIUS], dist, nn);
            } else if (shape == 3) {
                sphere_cylinder(c, r, e.button2_base[0], e.button2_base[1], disc20, disc21, m.sc[KM_SC_DISC_RADIUS], dist, nn);
            } else {
                sphere_cylinder(c, r, e.button2_base[0], e.button2_base[1], b2z, b2z + m.sc[KM_SC_STACK_TOP], m.sc[KM_SC_STACK_RADIUS], dist, nn);
            }
            if (dist > thr) continue;
            if (shape == 0) table_flag = 1;
            if (shape == 1) button_flag = 1;
            if (shape == 1 || shape == 2) any[0] = 1;
            if (shape == 3 || shape == 4) any[1] = 1;
            if (n < max_out) {
                out[n].sphere = s; out[n].body = b; out[n].shape = shape; out[n].dist = dist; out[n].n = nn;
                out[n].p = c - r * nn;
                ++n;
            }
        }
    }
    const_cast<KEnv&>(e).contact_button_any[0] = any[0];
    const_cast<KEnv&>(e).contact_button_any[1] = any[1];
    return n;
}

/* ------------------------------------------------------------------- constraint rows + PGS -- */
struct Row {
    double J[ND], W[ND]; /* Jacobian and M^-1 J^T */
    double invD, target, lo, hi;
    int friction_parent;  /* >= 0: bounds are +-mu * applied[parent] */
    double applied, rhs;
};

static bool is_ancestor_or_self(const KModel& m, int j, int body) {
    for (int b = body; b >= 0; b = m.b[b].parent) if (b == j) return true;
    return false;
}

static void fill_point_row(const KModel& m, const Kin& k, const Contact& c, V3 dir, Row& r) {
    for (int j = 0; j < ND; ++j) r.J[j] = 0.0;
    for (int j = 0; j < NB; ++j)
        if (is_ancestor_or_self(m, j, c.body)) r.J[j] = dot(dir, cross(k.a[j], c.p - k.p[j]));
    if (c.shape == 1) r.J[NB] = -dir.z; /* the button link moves along +z with the glider */
    if (c.shape == 3) r.J[NB + 1] = -dir.z;
}

static void plane_space(V3 n, V3& p, V3& q) { /* btPlaneSpace1 */
    if (fabs(n.z) > 0.7071067811865475244008443621048490) {
        const double a = n.y * n.y + n.z * n.z, k = 1.0 / sqrt(a);
        p = v3(0, -n.z * k, n.y * k);
        q = v3(a * k, -n.x * p.z, n.x * p.y);
    } else {
        const double a = n.x * n.x + n.y * n.y, k = 1.0 / sqrt(a);
        p = v3(-n.y * k, n.x * k, 0);
        q = v3(-n.z * p.y, n.z * p.x, a * k);
    }
}


struct StepScratch { std::vector<Row> rows; };

/* Motor set-points as PyBullet's setJointMotorControl2 records them (one per movable joint + the button glider). */
struct MotorCmd { double target, kp, kd, maxforce, maxvel; };
struct ButtonCmd { int position_control; double target, kp, kd, maxforce; };

/* One p.stepSimulation() with explicit motor set-points. */
static void physics_step_cmd(const KModel& m, KEnv& e, const MotorCmd* cmd, const ButtonCmd& btn, int iterations, StepScratch& sc) {
    const double dt = m.sc[KM_SC_TIMESTEP];
    Kin k;
    forward_kinematics(m, e.q, k);

    /* collision detection at the start-of-step configuration (Bullet: performDiscreteCollisionDetection precedes
       the solve; getContactPoints afterwards reports this manifold) */
    Contact contacts[8];
    const int max_contacts = (int)m.sc[KM_SC_MAX_CONTACTS];
    const int nc = detect_contacts(m, k, e, contacts, max_contacts, e.contact_button, e.contact_table);

    /* unconstrained velocity update: v = qd + dt * FD(q, qd) */
    Aba A;
    aba_setup(m, e.q, A);
    double qdd[NB], v0[ND];
    aba_forward_dynamics(m, A, e.qd, qdd);
    for (int i = 0; i < NB; ++i) v0[i] = e.qd[i] + dt * qdd[i];
    {   /* button glider: gravity + link damping on a 1-DoF prismatic body */
        const double vb = e.qdb;
        const double acc = m.sc[KM_SC_GRAVITY_Z] - m.sc[KM_SC_LIN_DAMPING] * vb * (1.0 + fabs(vb));
        v0[NB] = vb + dt * acc;
        v0[NB + 1] = 0.0;
        if (e.nbuttons == 2) {
            const double vb2 = e.qdb2;
            v0[NB + 1] = vb2 + dt * (m.sc[KM_SC_GRAVITY_Z] - m.sc[KM_SC_LIN_DAMPING] * vb2 * (1.0 + fabs(vb2)));
        }
    }
    const double minv_button = 1.0 / m.sc[KM_SC_BUTTON_MASS];

    /* columns of M^-1 (unit joint impulses) */
    double Minv[NB][NB];
    for (int i = 0; i < NB; ++i) {
        double tau[NB]; for (int j = 0; j < NB; ++j) tau[j] = (i == j) ? 1.0 : 0.0;
        aba_minv(m, A, tau, Minv[i]);
    }

    std::vector<Row>& rows = sc.rows;
    rows.clear();
    auto finish_row = [&](Row& r) {
        double tau[NB], w[NB];
        bool unit = false; int ui = -1; double us = 0;
        int nz = 0; for (int j = 0; j < NB; ++j) if (r.J[j] != 0.0) { ++nz; ui = j; us = r.J[j]; }
        unit = (nz == 1 && fabs(us) == 1.0);
        if (nz == 0) { for (int j = 0; j < NB; ++j) r.W[j] = 0.0; }
        else if (unit) { for (int j = 0; j < NB; ++j) r.W[j] = us * Minv[ui][j]; }
        else { for (int j = 0; j < NB; ++j) tau[j] = r.J[j]; aba_minv(m, A, tau, w); for (int j = 0; j < NB; ++j) r.W[j] = w[j]; }
        r.W[NB] = r.J[NB] * minv_button;
        r.W[NB + 1] = r.J[NB + 1] * minv_button;
        double D = 0; for (int j = 0; j < ND; ++j) D += r.J[j] * r.W[j];
        r.invD = 1.0 / D;
        double rel = 0; for (int j = 0; j < ND; ++j) rel += r.J[j] * v0[j];
        r.rhs = (r.target - rel) * r.invD; /* velocityImpulse (+ penetrationImpulse folded into target) */
        r.applied = 0.0;
    };
    auto joint_row = [&](int dof, double sign) {
        Row r; for (int j = 0; j < ND; ++j) r.J[j] = 0.0;
        r.J[dof] = sign; r.friction_parent = -1;
        return r;
    };

    /* (1) motors: btMultiBodyJointMotor, rhs = clamp(kp (q_des - q)/dt + kd (qd_des - qd) + qd, +-maxVelocity),
           impulse bound = force * dt.  Button first (it is loaded before the Kuka), then joints 0..11. */
    {
        Row r = joint_row(NB, 1.0);
        if (btn.position_control) { /* :347  POSITION_CONTROL targetPosition=0.1, default gains, default max force */
            r.target = btn.kp * (btn.target - e.qb) / dt + v0[NB] + btn.kd * (0.0 - v0[NB]);
            r.hi = btn.maxforce * dt;
        } else {            /* default joint motor created at load time: velocity target 0, small impulse bound */
            r.target = 0.0;
            r.hi = m.sc[KM_SC_BTN_IDLE_IMPULSE];
        }
        r.lo = -r.hi;
        finish_row(r); rows.push_back(r);
    }
    if (e.nbuttons == 2) {   /* the second button is loaded right after the first (kuka_2button_gym_env.py:58-69): same motor, next in line */
        Row r = joint_row(NB + 1, 1.0);
        if (btn.position_control) {
            r.target = btn.kp * (btn.target - e.qb2) / dt + v0[NB + 1] + btn.kd * (0.0 - v0[NB + 1]);
            r.hi = btn.maxforce * dt;
        } else { r.target = 0.0; r.hi = m.sc[KM_SC_BTN_IDLE_IMPULSE]; }
        r.lo = -r.hi;
        finish_row(r); rows.push_back(r);
    }
    for (int i = 0; i < NB; ++i) {
        const MotorCmd& b = cmd[i];
        Row r = joint_row(i, 1.0);
        double t = b.kp * (b.target - e.q[i]) / dt + v0[i] + b.kd * (0.0 - v0[i]);
        if (b.maxvel > 0.0) { if (t > b.maxvel) t = b.maxvel; if (t < -b.maxvel) t = -b.maxvel; }
        r.target = t; r.hi = b.maxforce * dt; r.lo = -r.hi;
        finish_row(r); rows.push_back(r);
    }
    /* (2) joint limits: btMultiBodyJointLimitConstraint, a row only while the limit is reached or violated */
    const double erp = m.sc[KM_SC_ERP], lim_eps = m.sc[KM_SC_LIMIT_EPS];
    auto limit_rows = [&](int dof, double qv, double lower, double upper) {
        const double pen_lo = qv - lower, pen_hi = upper - qv;
        /* Bullet tests `penetration > 0 -> no row`; a 1e-6 activation band (Bullet's own "todo: consider adding some
           safety threshold here") keeps a joint resting ON its limit from chattering on the last bit. */
        if (pen_lo <= lim_eps) { Row r = joint_row(dof, 1.0); r.target = -erp * pen_lo / dt; r.lo = 0; r.hi = m.sc[KM_SC_LIMIT_MAX_IMPULSE]; finish_row(r); rows.push_back(r); }
        if (pen_hi <= lim_eps) { Row r = joint_row(dof, -1.0); r.target = -erp * pen_hi / dt; r.lo = 0; r.hi = m.sc[KM_SC_LIMIT_MAX_IMPULSE]; finish_row(r); rows.push_back(r); }
    };
    limit_rows(NB, e.qb, m.sc[KM_SC_GLIDER_LOWER], m.sc[KM_SC_GLIDER_UPPER]);
    if (e.nbuttons == 2) limit_rows(NB + 1, e.qb2, m.sc[KM_SC_GLIDER_LOWER], m.sc[KM_SC_GLIDER_UPPER]);
    for (int i = 0; i < NB; ++i) limit_rows(i, e.q[i], m.b[i].lower, m.b[i].upper);
    /* (3) contact normals, then (4) two friction rows per contact */
    const int first_contact = (int)rows.size();
    for (int c = 0; c < nc; ++c) {
        Row r; fill_point_row(m, k, contacts[c], contacts[c].n, r);
        r.friction_parent = -1;
        const double pen = contacts[c].dist;
        r.target = pen > 0.0 ? -pen / dt : -erp * pen / dt;
        r.lo = 0; r.hi = 1e10;
        finish_row(r); rows.push_back(r);
    }
    for (int c = 0; c < nc; ++c) {
        V3 t1, t2; plane_space(contacts[c].n, t1, t2);
        for (int f = 0; f < 2; ++f) {
            Row r; fill_point_row(m, k, contacts[c], f ? t2 : t1, r);
            r.friction_parent = first_contact + c;
            r.target = 0.0; r.lo = r.hi = 0.0;
            finish_row(r); rows.push_back(r);
        }
    }

    /* projected Gauss-Seidel, Bullet's resolveSingleConstraintRowGeneric on accumulated delta-velocities,
       no warm start (btMultiBodyConstraintSolver disables it), fixed iteration count */
    double dv[ND]; for (int j = 0; j < ND; ++j) dv[j] = 0.0;
    const double mu = m.sc[KM_SC_FRICTION];
    for (int it = 0; it < iterations; ++it) {
        for (size_t ri = 0; ri < rows.size(); ++ri) {
            Row& r = rows[ri];
            double dvn = 0; for (int j = 0; j < ND; ++j) dvn += r.J[j] * dv[j];
            double delta = r.rhs - dvn * r.invD;
            double lo = r.lo, hi = r.hi;
            if (r.friction_parent >= 0) { hi = mu * rows[r.friction_parent].applied; lo = -hi; }
            const double sum = r.applied + delta;
            if (sum < lo) { delta = lo - r.applied; r.applied = lo; }
            else if (sum > hi) { delta = hi - r.applied; r.applied = hi; }
            else r.applied = sum;
            for (int j = 0; j < ND; ++j) dv[j] += r.W[j] * delta;
        }
    }
    /* write back and integrate (semi-implicit Euler): qd <- v, q <- q + dt qd */
    for (int i = 0; i < NB; ++i) { e.qd[i] = v0[i] + dv[i]; e.q[i] += dt * e.qd[i]; }
    e.qdb = v0[NB] + dv[NB]; e.qb += dt * e.qdb;
    if (e.nbuttons == 2) { e.qdb2 = v0[NB + 1] + dv[NB + 1]; e.qb2 += dt * e.qdb2; }

    /* link states after the step (getLinkState: COM of link 8; link-6 frame origin) */
    forward_kinematics(m, e.q, k);
    const int g = (int)m.sc[KM_SC_GRIPPER_BODY], eb = (int)m.sc[KM_SC_EE_BODY];
    e.gripper_pos[0] = k.com[g].x; e.gripper_pos[1] = k.com[g].y; e.gripper_pos[2] = k.com[g].z;
    e.ee_pos[0] = k.p[eb].x; e.ee_pos[1] = k.p[eb].y; e.ee_pos[2] = k.p[eb].z;
}

/* One p.stepSimulation() preceded by Kuka.applyAction's IK + motor set-points (kuka.py:142-187). */
static void physics_step(const KModel& m, KEnv& e, int button_armed, int iterations, StepScratch& sc, const double* q_joints = NULL) {
    Kin k;
    forward_kinematics(m, e.q, k);
    /* applyAction: IK at the current joint state, then the 12 POSITION_CONTROL set-points;
       use_inverse_kinematics=False (action_joints, kuka.py:158-161): the 7 arm set-points are the motor commands themselves */
    double q_ik[NB];
    if (q_joints) { for (int i = 0; i < NB; ++i) q_ik[i] = i < 7 ? q_joints[i] : 0.0; }
    else inverse_kinematics(m, k, e.q, e.ee, q_ik);
    const double finger_angle = 0.0; /* kuka_button_gym_env.py:313,334 */
    MotorCmd cmd[NB];
    for (int i = 0; i < NB; ++i) {
        const KBody& b = m.b[i];
        switch (b.target_mode) {
        case 0: cmd[i].target = q_ik[i]; break;
        case 1: cmd[i].target = e.ee_angle; break;
        case 2: cmd[i].target = -finger_angle; break;
        case 3: cmd[i].target = finger_angle; break;
        default: cmd[i].target = 0.0;
        }
        cmd[i].kp = b.kp; cmd[i].kd = b.kd; cmd[i].maxforce = b.maxforce; cmd[i].maxvel = b.maxvel;
    }
    ButtonCmd btn;
    btn.position_control = button_armed;
    btn.target = m.sc[KM_SC_BTN_TARGET]; btn.kp = m.sc[KM_SC_BTN_KP]; btn.kd = m.sc[KM_SC_BTN_KD]; btn.maxforce = m.sc[KM_SC_BTN_MAXFORCE];
    physics_step_cmd(m, e, cmd, btn, iterations, sc);
}

} /* namespace */

struct KukaWorld {
    KModel m;
    std::vector<KEnv> envs;
    KEnv snapshot; /* state after the 500 settle steps of reset() (:242-247), identical for every episode */
    StepScratch scratch;
    int iterations, max_steps;
};

namespace {

/* Kuka2ButtonGymEnv scene: button 1 at (0.5, 0.125), button 2 at (0.5, -0.125) (kuka_2button_gym_env.py:49-69), both resting on the table */
static const double TWO_BUTTON_Y = 0.125, TWO_BUTTON_RAND_Y = 0.175;
static void two_button_defaults(const KModel& m, KEnv& e) {
    e.nbuttons = 2;
    e.button_base[0] = m.sc[KM_SC_BUTTON_BASE]; e.button_base[1] = TWO_BUTTON_Y; e.button_base[2] = m.sc[KM_SC_BUTTON_BASE + 2];
    e.button2_base[0] = m.sc[KM_SC_BUTTON_BASE]; e.button2_base[1] = -TWO_BUTTON_Y; e.button2_base[2] = m.sc[KM_SC_BUTTON_BASE + 2];
}

/* Kuka.applyAction's accumulate + clip (kuka.py:134-139) */
static void apply_ee_delta(const KukaWorld& w, const srl_sim* s, KEnv& e, const double d[3]) {
    /* small_constraints = not random_target (:239); Kuka2Button always uses the large box (kuka_2button_gym_env.py:78) */
    const double* box = w.m.sc + ((s->cfg.random_target || s->kind == SRL_ENV_KUKA_2BUTTON) ? KM_SC_BOX_LARGE : KM_SC_BOX_SMALL);
    for (int a = 0; a < 3; ++a) {
        e.ee[a] += d[a];
        if (e.ee[a] < box[2 * a]) e.ee[a] = box[2 * a];
        if (e.ee[a] > box[2 * a + 1]) e.ee[a] = box[2 * a + 1];
    }
}

static void refresh_link_states(const KModel& m, KEnv& e) {
    Kin k; forward_kinematics(m, e.q, k);
    const int g = (int)m.sc[KM_SC_GRIPPER_BODY], eb = (int)m.sc[KM_SC_EE_BODY];
    e.gripper_pos[0] = k.com[g].x; e.gripper_pos[1] = k.com[g].y; e.gripper_pos[2] = k.com[g].z;
    e.ee_pos[0] = k.p[eb].x; e.ee_pos[1] = k.p[eb].y; e.ee_pos[2] = k.p[eb].z;
}

static void make_snapshot(KukaWorld& w, const srl_sim* s) {
    KEnv e; memset(&e, 0, sizeof(e));
    for (int i = 0; i < NB; ++i) { e.q[i] = w.m.b[i].qinit; e.qd[i] = 0.0; }  /* resetJointState, kuka.py:68-69 */
    for (int a = 0; a < 3; ++a) { e.ee[a] = w.m.sc[KM_SC_EE_INIT + a]; e.button_base[a] = w.m.sc[KM_SC_BUTTON_BASE + a]; }
    e.ee_angle = 0.0;
    e.nbuttons = 1;
    if (s->kind == SRL_ENV_KUKA_2BUTTON) two_button_defaults(w.m, e);
    const double zero[3] = {0, 0, 0};
    double qj[7];
    for (int j = 0; j < 7; ++j) qj[j] = w.m.b[j].qinit; /* self._kuka.joint_positions[:7] (:244) */
    for (int t = 0; t < 500; ++t) { /* :242-247 */
        if (s->cfg.action_joints) physics_step(w.m, e, 0, w.iterations, w.scratch, qj);
        else { apply_ee_delta(w, s, e, zero); physics_step(w.m, e, 0, w.iterations, w.scratch); }
    }
    w.snapshot = e;
}

} /* namespace */

KukaWorld* oracle_kuka_create(srl_sim* s, const void* blob, size_t bytes) {
    KukaWorld* w = new KukaWorld();
    if (!parse_model(blob, bytes, w->m)) { delete w; return NULL; }
    w->iterations = s->cfg.solver_iterations > 0 ? s->cfg.solver_iterations : (int)w->m.sc[KM_SC_SOLVER_ITERS];
    w->max_steps = s->cfg.max_steps > 0 ? s->cfg.max_steps : 1000; /* MAX_STEPS, :17 and kuka_rand_button_gym_env.py:3 */
    if (s->cfg.timestep > 0) w->m.sc[KM_SC_TIMESTEP] = (double)s->cfg.timestep;
    if (s->kind == SRL_ENV_KUKA_MOVING_BUTTON && s->cfg.max_steps <= 0) w->max_steps = 1500; /* kuka_moving_button_gym_env.py:3,28 */
    if (s->kind == SRL_ENV_KUKA_2BUTTON) {
        if (s->cfg.max_steps <= 0) w->max_steps = 1500; /* kuka_2button_gym_env.py:3,33 */
        /* `use_null_space = True` (:80): calculateInverseKinematics(uid, link, pos, orn, ll, ul, jr, rp) (kuka.py:147-149).  RECALLED
           pybullet 1.8.6 behaviour: the null-space task is only enabled when all four lists have one entry per JOINT (14 for
           kuka_with_gripper2; kuka.py:34-40 gives 7), so it is silently dropped -- and since this call passes no jointDamping, the
           server's default damping 0.5 per DoF replaces the 1e-5 of the single-button envs. */
        w->m.sc[KM_SC_IK_DAMPING] = 0.5;
    }
    w->envs.resize(s->n);
    memset(w->envs.data(), 0, sizeof(KEnv) * (size_t)s->n);
    make_snapshot(*w, s);
    return w;
}

void oracle_kuka_destroy(KukaWorld* w) { delete w; }

void oracle_kuka_reset_env(srl_sim* s, int i, const double* draws) {
    KukaWorld& w = *s->kuka;
    KEnv& e = w.envs[i];
    const uint32_t episode = e.episode, total = e.total_steps;
    double d[18];
    if (draws) {
        memcpy(d, draws, sizeof(d));
    } else {
        d[17] = 0.0;
        const uint64_t genv = s->cfg.global_env_offset + (uint64_t)i;
        uint32_t r[4];
        philox4x32_10(s->seed, genv, episode, PHILOX_PURPOSE_RESET0 + 0, r);
        /* :227-231  x_pos = 0.5 + 0.15 * uniform(-1, 1), y_pos = 0 + 0.3 * uniform(-1, 1) */
        d[0] = w.m.sc[KM_SC_BUTTON_BASE] + w.m.sc[KM_SC_RAND_X] * (-1.0 + 2.0 * philox_u01(r[0], r[1]));
        d[1] = w.m.sc[KM_SC_BUTTON_BASE + 1] + w.m.sc[KM_SC_RAND_Y] * (-1.0 + 2.0 * philox_u01(r[2], r[3]));
        for (int k = 0; k < 5; ++k) { /* :250-266 */
            philox4x32_10(s->seed, genv, episode, PHILOX_PURPOSE_RESET0 + 1 + k, r);
            d[2 + 3 * k] = d[3 + 3 * k] = d[4 + 3 * k] = 0.0;
            if (s->cfg.action_joints) {
                /* joints += DELTA_THETA * np_random.normal(joints.shape): ONE draw from N(loc=7, 1), broadcast to the 7 joints (:257-260) */
                const double u1 = philox_u01(r[0], r[1]), u2 = philox_u01(r[2], r[3]);
                d[2 + 3 * k] = 0.1 * (7.0 + sqrt(-2.0 * log(1.0 - u1)) * cos(2.0 * M_PI * u2));
            } else if (s->cfg.is_discrete) {
                const double sign = philox_u01(r[0], r[1]) > 0.5 ? 1.0 : -1.0; /* np_random.rand() > 0.5 */
                const int idx = (int)(((uint64_t)r[2] * 3u) >> 32);            /* np_random.randint(3) */
                d[2 + 3 * k + idx] = sign * 0.03;                               /* DELTA_V */
            } else {
                /* np_random.normal((3,)) is ONE draw from N(loc=3, 1); after L2 normalisation it is +-1 and
                   broadcasts to the three axes: action[:3] = +-DELTA_V_CONTINUOUS */
                const double u1 = philox_u01(r[0], r[1]), u2 = philox_u01(r[2], r[3]);
                const double z = sqrt(-2.0 * log(1.0 - u1)) * cos(2.0 * M_PI * u2);
                const double sign = (3.0 + z) >= 0.0 ? 1.0 : -1.0;
                d[2 + 3 * k] = d[3 + 3 * k] = d[4 + 3 * k] = sign * 0.0035;
            }
        }
        if (s->kind == SRL_ENV_KUKA_MOVING_BUTTON) { /* BUTTON_SPEED * np_random.choice([-1, 1]) */
            philox4x32_10(s->seed, genv, episode, PHILOX_PURPOSE_RESET0 + 6, r);
            d[17] = (r[0] & 1u) ? 0.001 : -0.001;
        }
    }
    const bool two = s->kind == SRL_ENV_KUKA_2BUTTON;
    if (two && !draws) {
        /* button 1 is always at (0.5, 0.125): its random placement is overwritten (kuka_2button_gym_env.py:56-57);
           button 2: x = 0.5 + 0.15 U(-1, 1), y = -0.125 + 0.175 U(-1, 0) (:63-66) */
        const uint64_t genv = s->cfg.global_env_offset + (uint64_t)i;
        uint32_t r[4];
        philox4x32_10(s->seed, genv, episode, PHILOX_PURPOSE_RESET0 + 0, r);
        d[0] = w.m.sc[KM_SC_BUTTON_BASE] + w.m.sc[KM_SC_RAND_X] * (-1.0 + 2.0 * philox_u01(r[0], r[1]));
        d[1] = -TWO_BUTTON_Y + TWO_BUTTON_RAND_Y * (-1.0 + philox_u01(r[2], r[3]));
    }
    e = w.snapshot;
    e.episode = episode + 1; e.total_steps = total;
    e.btn_speed = (s->kind == SRL_ENV_KUKA_MOVING_BUTTON) ? d[17] : 0.0;
    if (s->cfg.random_target) {
        if (two) { e.button2_base[0] = d[0]; e.button2_base[1] = d[1]; }
        else { e.button_base[0] = d[0]; e.button_base[1] = d[1]; }
    }
    for (int k = 0; k < 5; ++k) { /* N_RANDOM_ACTIONS_AT_INIT, :250-269 */
        if (s->cfg.action_joints) {
            double qj[7];
            for (int j = 0; j < 7; ++j) qj[j] = w.m.b[j].qinit + d[2 + 3 * k];
            physics_step(w.m, e, 0, w.iterations, w.scratch, qj);
        } else {
            apply_ee_delta(w, s, e, d + 2 + 3 * k);
            physics_step(w.m, e, 0, w.iterations, w.scratch);
        }
    }
    /* :273-274  button_pos = link state of the button link (COM = link origin) + BUTTON_DISTANCE_HEIGHT */
    e.button_pos[0] = e.button_base[0];
    e.button_pos[1] = e.button_base[1];
    e.button_pos[2] = e.button_base[2] + w.m.sc[KM_SC_GLIDER_Z] + e.qb + w.m.sc[KM_SC_TARGET_HEIGHT];
    if (two) e.button_pos[2] = -0.2 + w.m.sc[KM_SC_TARGET_HEIGHT]; /* button_all_pos = [x, y, Z_TABLE + BUTTON_DISTANCE_HEIGHT] (:59,69,72), not a link state */
    e.goal_id = 0; e.n_contacts2[0] = e.n_contacts2[1] = 0; e.pressed[0] = e.pressed[1] = 0;
    e.counter = 0; e.n_contacts = 0; e.n_outside = 0; e.terminated = 0;
    e.ep_ret = 0.0; e.ep_len = 0;
}

void oracle_kuka_obs(const srl_sim* s, int i, float* obs) {
    const KEnv& e = s->kuka->envs[i];
    for (int a = 0; a < 3; ++a) obs[a] = (float)(e.gripper_pos[a] - e.button_pos[a]); /* :175-186, RELATIVE_POS */
}

/* Kuka2ButtonGymEnv._reward (kuka_2button_gym_env.py:157-214) + _termination + the VecEnv epilogue */
static void two_button_reward(srl_sim* s, KukaWorld& w, KEnv& e, int i, double distance, float* obs, float* rew, uint8_t* done,
                              float* ep_ret, int32_t* ep_len) {
    double reward = 0.0;
    const int contact = e.contact_button_any[e.goal_id];      /* getContactPoints(button_uid[goal_id], kuka_uid): ANY link of that button (:165) */
    e.n_contacts2[e.goal_id] += contact;
    if (e.goal_id == 1) reward = contact;                     /* sparse reward only on the last button (:169-170) */
    if (e.n_contacts2[e.goal_id] >= 5 && !e.pressed[e.goal_id]) {   /* next button (:173-178) */
        e.pressed[e.goal_id] = 1;
        if (e.goal_id == 0) {
            e.goal_id = 1;
            e.button_pos[0] = e.button2_base[0]; e.button_pos[1] = e.button2_base[1];   /* button_all_pos[1]; z is the same constant */
        }
    }
    const int table = e.contact_table;
    if (distance > (double)s->cfg.max_distance || table) { reward = -1.0; e.n_outside += 1; } else e.n_outside = 0;
    if (table || e.n_contacts2[1] >= 5 || e.n_outside >= 5000 - 1) e.terminated = 1;   /* :188-190 */
    if (s->cfg.shape_reward) {   /* :192-212 */
        if (e.terminated && reward > 0) reward = 50;
        else if (e.n_contacts2[e.goal_id] < 5 && contact) reward = 25;
        else if (table) reward = -250;
        else if (distance > (double)s->cfg.max_distance) reward = -20;
        else reward = -distance;
    }
    const int is_done = e.terminated || e.counter > w.max_steps;
    e.ep_ret += reward; e.ep_len += 1;
    e.n_contacts = e.n_contacts2[0];   /* reported through SRL_F_COUNTERS */
    if (rew) rew[i] = (float)reward;
    if (done) done[i] = (uint8_t)is_done;
    if (is_done) {
        if (ep_ret) ep_ret[i] = (float)e.ep_ret;
        if (ep_len) ep_len[i] = e.ep_len;
        if (s->auto_reset) oracle_kuka_reset_env(s, i, NULL);
    }
    if (obs) oracle_kuka_obs(s, i, obs + 3 * (size_t)i);
}

void oracle_kuka_step_env(srl_sim* s, int i, const void* actions, const float* noise, float* obs, float* rew,
                          uint8_t* done, float* ep_ret, int32_t* ep_len) {
    KukaWorld& w = *s->kuka;
    KEnv& e = w.envs[i];
    const uint64_t genv = s->cfg.global_env_offset + (uint64_t)i;
    /* ---- step(): action decoding + noise (:293-340) ---- */
    double d[3] = {0, 0, 0};
    double nz = 0.0;
    const double noise_std = s->cfg.action_joints ? 0.002 : s->cfg.is_discrete ? 0.01 : 0.0001; /* NOISE_STD_JOINTS, NOISE_STD, NOISE_STD_CONTINUOUS (:31-33) */
    double qj[7];
    const bool joints = s->cfg.action_joints != 0;
    if (noise) nz = (double)noise[i];
    else {
        uint32_t r[4];
        philox4x32_10(s->seed, genv, e.total_steps, PHILOX_PURPOSE_NOISE, r);
        const double u1 = philox_u01(r[0], r[1]), u2 = philox_u01(r[2], r[3]);
        nz = (double)(float)(noise_std * sqrt(-2.0 * log(1.0 - u1)) * cos(2.0 * M_PI * u2));
    }
    uint32_t ra[4] = {0, 0, 0, 0};
    if (!actions) philox4x32_10(s->seed, genv, e.total_steps, PHILOX_PURPOSE_ACTION, ra);
    if (joints) {
        /* real_action = action * (DELTA_THETA + N(0, NOISE_STD_JOINTS)) + joint_positions[:7]  (:317-323); the set-points are
           relative to the INITIAL joint vector, which the reference never updates (kuka.py:65-66) */
        float a[7];
        if (actions) { for (int k = 0; k < 7; ++k) a[k] = ((const float*)actions)[7 * (size_t)i + k]; }
        else {
            uint32_t rb[4];
            philox4x32_10(s->seed, genv, e.total_steps, PHILOX_PURPOSE_ACTION + 1, rb);
            for (int k = 0; k < 4; ++k) a[k] = (float)((double)ra[k] * (2.0 / 4294967296.0) - 1.0);
            for (int k = 0; k < 3; ++k) a[4 + k] = (float)((double)rb[k] * (2.0 / 4294967296.0) - 1.0);
        }
        const double d_theta = 0.1 + nz;
        for (int k = 0; k < 7; ++k) qj[k] = (double)a[k] * d_theta + w.m.b[k].qinit;
    } else if (s->cfg.is_discrete) {
        const int a = actions ? ((const int32_t*)actions)[i] : (int)(((uint64_t)ra[0] * 6u) >> 32);
        if (a >= 0) { /* a < 0: step(None) -> zero action, no noise (:295-299) */
            const double dv = 0.03 + nz; /* DELTA_V + N(0, NOISE_STD) */
            const double dxs[6] = {-dv, dv, 0, 0, 0, 0}, dys[6] = {0, 0, -dv, dv, 0, 0};
            const double dzd[6] = {0, 0, 0, 0, -dv, -dv}, dzu[6] = {0, 0, 0, 0, -dv, dv};
            d[0] = dxs[a % 6]; d[1] = dys[a % 6]; d[2] = s->cfg.force_down ? dzd[a % 6] : dzu[a % 6];
        }
    } else {
        float a[3];
        if (actions) { for (int k = 0; k < 3; ++k) a[k] = ((const float*)actions)[3 * i + k]; }
        else { for (int k = 0; k < 3; ++k) a[k] = (float)((double)ra[k] * (2.0 / 4294967296.0) - 1.0); }
        const double dv = 0.0035 + nz; /* DELTA_V_CONTINUOUS + N(0, NOISE_STD_CONTINUOUS) */
        d[0] = (double)a[0] * dv; d[1] = (double)a[1] * dv;
        d[2] = s->cfg.force_down ? -fabs((double)a[2] * dv) : (double)a[2] * dv;
    }
    e.total_steps += 1;
    if (s->kind == SRL_ENV_KUKA_MOVING_BUTTON) {
        /* kuka_moving_button_gym_env.py:109-119: bounce at the table edge, slide the target, teleport the button base to
           (button_pos - BUTTON_DISTANCE_HEIGHT) -- i.e. x, y follow the target and z becomes the button LINK height of reset */
        if (e.button_pos[1] > 0.3 || e.button_pos[1] < -0.3) e.btn_speed = -e.btn_speed;
        e.button_pos[1] += e.btn_speed;
        e.button_base[0] = e.button_pos[0]; e.button_base[1] = e.button_pos[1];
        e.button_base[2] = e.button_pos[2] - w.m.sc[KM_SC_TARGET_HEIGHT];
    }
    /* ---- step2() (:342-368) ---- */
    for (int rep = 0; rep < s->cfg.action_repeat; ++rep) {
        if (joints) physics_step(w.m, e, 1, w.iterations, w.scratch, qj);
        else { apply_ee_delta(w, s, e, d); physics_step(w.m, e, 1, w.iterations, w.scratch); }
        if (e.terminated || e.counter > w.max_steps) break; /* _termination() */
        e.counter += 1;
    }
    /* ---- _reward() (:428-463) ---- */
    const double dx = e.button_pos[0] - e.gripper_pos[0], dy = e.button_pos[1] - e.gripper_pos[1], dz = e.button_pos[2] - e.gripper_pos[2];
    const double distance = sqrt(dx * dx + dy * dy + dz * dz);
    if (s->kind == SRL_ENV_KUKA_2BUTTON) {
        two_button_reward(s, w, e, i, distance, obs, rew, done, ep_ret, ep_len);
        return;
    }
    double reward = e.contact_button ? 1.0 : 0.0;
    e.n_contacts += e.contact_button ? 1 : 0;
    const int table = e.contact_table;
    if (distance > (double)s->cfg.max_distance || table) { reward = -1.0; e.n_outside += 1; } else e.n_outside = 0;
    if (table || e.n_contacts >= 5 || e.n_outside >= 5000) e.terminated = 1; /* N_CONTACTS_BEFORE_TERMINATION, N_STEPS_OUTSIDE_SAFETY_SPHERE */
    if (s->cfg.shape_reward) {
        if (s->cfg.is_discrete) reward = -distance;
        else if (e.terminated && reward > 0) reward = 50;
        else if (e.terminated && reward < 0) reward = -250;
        else reward = -distance;
    }
    const int is_done = e.terminated || e.counter > w.max_steps;
    e.ep_ret += reward; e.ep_len += 1;
    if (rew) rew[i] = (float)reward;
    if (done) done[i] = (uint8_t)is_done;
    if (is_done) {
        if (ep_ret) ep_ret[i] = (float)e.ep_ret;
        if (ep_len) ep_len[i] = e.ep_len;
        if (s->auto_reset) oracle_kuka_reset_env(s, i, NULL);
    }
    if (obs) oracle_kuka_obs(s, i, obs + 3 * (size_t)i);
}

/* Scene primitives of env i for the CPU checker of the image path (csrc/render_core.h holds the list builder both sides share): joint
 * frames and gripper collision spheres from this file's own float64 forward kinematics. */
int oracle_kuka_scene(const srl_sim* s, int i, void* prims_out) {
    const KukaWorld& w = *s->kuka;
    const KModel& m = w.m;
    const KEnv& e = w.envs[i];
    Kin k;
    forward_kinematics(m, e.q, k);
    float jp[NB * 3], sph[KM_MAX_SPHERES * 4];
    for (int b = 0; b < NB; ++b) { jp[3 * b] = (float)k.p[b].x; jp[3 * b + 1] = (float)k.p[b].y; jp[3 * b + 2] = (float)k.p[b].z; }
    int ns = 0;
    for (int t = 0; t < m.nsphere; ++t) {
        const KSphere& sp = m.sph[t];
        if (sp.body < 7) continue;
        const V3 c = k.p[sp.body] + k.R[sp.body] * sp.c;
        sph[4 * ns] = (float)c.x; sph[4 * ns + 1] = (float)c.y; sph[4 * ns + 2] = (float)c.z; sph[4 * ns + 3] = (float)sp.r;
        ++ns;
    }
    SrlKukaSceneConst K;
    for (int a = 0; a < 3; ++a) K.base[a] = (float)m.sc[KM_SC_BASE_POS + a];
    K.table_z = (float)m.sc[KM_SC_TABLE_TOP_Z]; K.txmin = (float)m.sc[KM_SC_TABLE_XMIN]; K.txmax = (float)m.sc[KM_SC_TABLE_XMAX];
    K.tymin = (float)m.sc[KM_SC_TABLE_YMIN]; K.tymax = (float)m.sc[KM_SC_TABLE_YMAX];
    K.glider_z = (float)m.sc[KM_SC_GLIDER_Z]; K.disc_r = (float)m.sc[KM_SC_DISC_RADIUS]; K.disc_z0 = (float)m.sc[KM_SC_DISC_Z0]; K.disc_z1 = (float)m.sc[KM_SC_DISC_Z1];
    K.stack_r = (float)m.sc[KM_SC_STACK_RADIUS]; K.stack_top = (float)m.sc[KM_SC_STACK_TOP];
    K.two_buttons = e.nbuttons == 2;
    return srl_kuka_scene(K, jp, sph, ns, (float)e.button_base[0], (float)e.button_base[1], (float)e.button_base[2], (float)e.qb,
                          (float)e.button2_base[0], (float)e.button2_base[1], (float)m.sc[KM_SC_BUTTON_BASE + 2], (float)e.qb2, static_cast<SrlPrim*>(prims_out));
}

int oracle_kuka_get_state(srl_sim* s, int field, void* dst, size_t bytes) {
    KukaWorld& w = *s->kuka;
    const size_t N = (size_t)s->n;
    auto need = [&](size_t width, size_t elem) { if (bytes != N * width * elem) { oracle_set_error("get_state: size mismatch"); return false; } return true; };
    double* D = (double*)dst; int32_t* I = (int32_t*)dst;
    for (size_t i = 0; i < N; ++i) {
        const KEnv& e = w.envs[i];
        switch (field) {
        case SRL_F_ROBOT_POS: if (!need(3, 8)) return 1; for (int a = 0; a < 3; ++a) D[3 * i + a] = e.gripper_pos[a]; break;
        case SRL_F_TARGET_POS: if (!need(3, 8)) return 1; for (int a = 0; a < 3; ++a) D[3 * i + a] = e.button_pos[a]; break;
        case SRL_F_STEP_COUNTER: if (!need(1, 4)) return 1; I[i] = e.counter; break;
        case SRL_F_JOINT_POS: if (!need(NB, 8)) return 1; for (int a = 0; a < NB; ++a) D[NB * i + a] = e.q[a]; break;
        case SRL_F_JOINT_VEL: if (!need(NB, 8)) return 1; for (int a = 0; a < NB; ++a) D[NB * i + a] = e.qd[a]; break;
        case SRL_F_EE_CMD: if (!need(3, 8)) return 1; for (int a = 0; a < 3; ++a) D[3 * i + a] = e.ee[a]; break;
        case SRL_F_EE_POS: if (!need(3, 8)) return 1; for (int a = 0; a < 3; ++a) D[3 * i + a] = e.ee_pos[a]; break;
        case SRL_F_BUTTON_GLIDER: if (!need(2, 8)) return 1; D[2 * i] = e.qb; D[2 * i + 1] = e.qdb; break;
        case SRL_F_COUNTERS: if (!need(4, 4)) return 1; I[4 * i] = e.n_contacts; I[4 * i + 1] = e.n_outside; I[4 * i + 2] = e.terminated; I[4 * i + 3] = (int32_t)e.episode; break;
        case SRL_F_EPISODE_STATS: if (!need(2, 8)) return 1; D[2 * i] = e.ep_ret; D[2 * i + 1] = (double)e.ep_len; break;
        case SRL_F_BUTTON_BASE: if (!need(3, 8)) return 1; for (int a = 0; a < 3; ++a) D[3 * i + a] = e.button_base[a]; break;
        case SRL_F_TWO_BUTTON: if (!need(8, 8)) return 1;
            D[8 * i] = e.n_contacts2[0]; D[8 * i + 1] = e.n_contacts2[1]; D[8 * i + 2] = e.goal_id;
            for (int a = 0; a < 3; ++a) D[8 * i + 3 + a] = e.button2_base[a];
            D[8 * i + 6] = e.qb2; D[8 * i + 7] = e.qdb2; break;
        default: oracle_set_error("get_state: unknown field %d", field); return 1;
        }
    }
    return 0;
}

int oracle_kuka_set_state(srl_sim* s, int field, const void* src, size_t bytes) {
    KukaWorld& w = *s->kuka;
    const size_t N = (size_t)s->n;
    auto need = [&](size_t width, size_t elem) { if (bytes != N * width * elem) { oracle_set_error("set_state: size mismatch"); return false; } return true; };
    const double* D = (const double*)src; const int32_t* I = (const int32_t*)src;
    for (size_t i = 0; i < N; ++i) {
        KEnv& e = w.envs[i];
        switch (field) {
        case SRL_F_JOINT_POS: if (!need(NB, 8)) return 1; for (int a = 0; a < NB; ++a) e.q[a] = D[NB * i + a]; refresh_link_states(w.m, e); break;
        case SRL_F_JOINT_VEL: if (!need(NB, 8)) return 1; for (int a = 0; a < NB; ++a) e.qd[a] = D[NB * i + a]; break;
        case SRL_F_EE_CMD: if (!need(3, 8)) return 1; for (int a = 0; a < 3; ++a) e.ee[a] = D[3 * i + a]; break;
        case SRL_F_TARGET_POS: if (!need(3, 8)) return 1; for (int a = 0; a < 3; ++a) e.button_pos[a] = D[3 * i + a]; break;
        case SRL_F_BUTTON_GLIDER: if (!need(2, 8)) return 1; e.qb = D[2 * i]; e.qdb = D[2 * i + 1]; break;
        case SRL_F_BUTTON_BASE: if (!need(3, 8)) return 1; for (int a = 0; a < 3; ++a) e.button_base[a] = D[3 * i + a]; break;
        case SRL_F_STEP_COUNTER: if (!need(1, 4)) return 1; e.counter = I[i]; break;
        case SRL_F_COUNTERS: if (!need(4, 4)) return 1; e.n_contacts = I[4 * i]; e.n_outside = I[4 * i + 1]; e.terminated = I[4 * i + 2]; break;
        default: oracle_set_error("set_state: field %d not settable", field); return 1;
        }
    }
    return 0;
}

/* ---- test hooks: the building blocks, exposed so tests can check them against closed forms ---- */
extern "C" {

/* forward kinematics: p[12][3] joint-frame origins, R[12][9] rotations, com[12][3] */
int oracle_kuka_fk(const void* blob, size_t bytes, const double* q, double* p, double* R, double* com) {
    KModel m; if (!parse_model(blob, bytes, m)) return 1;
    Kin k; forward_kinematics(m, q, k);
    for (int i = 0; i < NB; ++i) {
        p[3 * i] = k.p[i].x; p[3 * i + 1] = k.p[i].y; p[3 * i + 2] = k.p[i].z;
        com[3 * i] = k.com[i].x; com[3 * i + 1] = k.com[i].y; com[3 * i + 2] = k.com[i].z;
        for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) R[9 * i + 3 * a + b] = k.R[i].m[a][b];
    }
    return 0;
}

/* unconstrained forward dynamics qdd = FD(q, qd) with optional overrides of the damping terms */
int oracle_kuka_fd(const void* blob, size_t bytes, const double* q, const double* qd, int with_damping, double* qdd) {
    KModel m; if (!parse_model(blob, bytes, m)) return 1;
    if (!with_damping) { for (int i = 0; i < NB; ++i) m.b[i].damping = 0.0; m.sc[KM_SC_LIN_DAMPING] = 0.0; m.sc[KM_SC_ANG_DAMPING] = 0.0; }
    Aba A; aba_setup(m, q, A);
    aba_forward_dynamics(m, A, qd, qdd);
    return 0;
}

/* M^-1 (row-major 12x12) by ABA impulse responses */
int oracle_kuka_minv(const void* blob, size_t bytes, const double* q, double* Minv) {
    KModel m; if (!parse_model(blob, bytes, m)) return 1;
    Aba A; aba_setup(m, q, A);
    for (int i = 0; i < NB; ++i) {
        double tau[NB]; for (int j = 0; j < NB; ++j) tau[j] = (i == j) ? 1.0 : 0.0;
        aba_minv(m, A, tau, Minv + NB * i);
    }
    return 0;
}

/* one IK iteration at q towards target_pos (orientation target from the blob) */
int oracle_kuka_ik(const void* blob, size_t bytes, const double* q, const double* target_pos, double* q_ik) {
    KModel m; if (!parse_model(blob, bytes, m)) return 1;
    Kin k; forward_kinematics(m, q, k);
    for (int i = 0; i < NB; ++i) q_ik[i] = q[i];
    inverse_kinematics(m, k, q, target_pos, q_ik);
    return 0;
}

} /* extern "C" */


/* ---- low-level "bullet-like" world for tests/golden/fake_pybullet.py --------------------------------------
 * Lets the UNMODIFIED reference env classes (environments/kuka_gym/ *.py) run on the oracle's physics, so that
 * their own Python logic (action decoding, RNG order, reward, termination, reset sequencing) produces the golden
 * vectors our env-level restatement is checked against. */
extern "C" {

struct OkbWorld { KModel m; KEnv e; StepScratch sc; int iterations; };

void* okb_create(const void* blob, size_t bytes) {
    OkbWorld* w = new OkbWorld();
    if (!parse_model(blob, bytes, w->m)) { delete w; return NULL; }
    memset(&w->e, 0, sizeof(KEnv));
    w->e.nbuttons = 1;
    w->iterations = (int)w->m.sc[KM_SC_SOLVER_ITERS];
    for (int a = 0; a < 3; ++a) w->e.button_base[a] = w->m.sc[KM_SC_BUTTON_BASE + a];
    return w;
}
void okb_destroy(void* h) { delete (OkbWorld*)h; }
void okb_reset_world(void* h) {           /* p.resetSimulation(): everything back to the load-time state */
    OkbWorld* w = (OkbWorld*)h;
    memset(&w->e, 0, sizeof(KEnv));
    w->e.nbuttons = 1;
    for (int a = 0; a < 3; ++a) w->e.button_base[a] = w->m.sc[KM_SC_BUTTON_BASE + a];
    refresh_link_states(w->m, w->e);
}
void okb_set_iterations(void* h, int n) { ((OkbWorld*)h)->iterations = n; }
void okb_set_ik_damping(void* h, double d) { ((OkbWorld*)h)->m.sc[KM_SC_IK_DAMPING] = d; }   /* jointDamping given / server default */
void okb_set_button2_base(void* h, double x, double y) {   /* a second simple_button body is loaded (kuka_2button_gym_env.py:68) */
    OkbWorld* w = (OkbWorld*)h; w->e.nbuttons = 2; w->e.button2_base[0] = x; w->e.button2_base[1] = y; w->e.button2_base[2] = w->m.sc[KM_SC_BUTTON_BASE + 2];
}
void okb_get2(void* h, double* out) {   /* contact with any link of button 1 / button 2 in the last manifold; second glider */
    OkbWorld* w = (OkbWorld*)h; out[0] = w->e.contact_button_any[0]; out[1] = w->e.contact_button_any[1]; out[2] = w->e.qb2; out[3] = w->e.qdb2;
}
void okb_set_button_base(void* h, double x, double y) { OkbWorld* w = (OkbWorld*)h; w->e.button_base[0] = x; w->e.button_base[1] = y; }
void okb_set_button_base3(void* h, double x, double y, double z) { OkbWorld* w = (OkbWorld*)h; w->e.button_base[0] = x; w->e.button_base[1] = y; w->e.button_base[2] = z; }
void okb_reset_joint(void* h, int body, double q) {  /* p.resetJointState */
    OkbWorld* w = (OkbWorld*)h; w->e.q[body] = q; w->e.qd[body] = 0.0; refresh_link_states(w->m, w->e);
}
void okb_ik(void* h, const double* target_pos, double* q_out) {  /* p.calculateInverseKinematics at the current state */
    OkbWorld* w = (OkbWorld*)h;
    Kin k; forward_kinematics(w->m, w->e.q, k);
    for (int i = 0; i < NB; ++i) q_out[i] = w->e.q[i];
    inverse_kinematics(w->m, k, w->e.q, target_pos, q_out);
}
/* p.stepSimulation() with the motor table accumulated from setJointMotorControl2 calls: 12 x (target, kp, kd, force, maxvel) */
void okb_step(void* h, const double* motors, int btn_position_control, double btn_target, double btn_kp, double btn_kd, double btn_force) {
    OkbWorld* w = (OkbWorld*)h;
    MotorCmd cmd[NB];
    for (int i = 0; i < NB; ++i) { cmd[i].target = motors[5 * i]; cmd[i].kp = motors[5 * i + 1]; cmd[i].kd = motors[5 * i + 2]; cmd[i].maxforce = motors[5 * i + 3]; cmd[i].maxvel = motors[5 * i + 4]; }
    ButtonCmd btn; btn.position_control = btn_position_control; btn.target = btn_target; btn.kp = btn_kp; btn.kd = btn_kd; btn.maxforce = btn_force;
    physics_step_cmd(w->m, w->e, cmd, btn, w->iterations, w->sc);
}
/* out[0..11] q, [12..23] qd, [24] glider q, [25..27] link-8 COM, [28..30] button link origin, [31] button manifold, [32] table manifold */
void okb_get(void* h, double* out) {
    OkbWorld* w = (OkbWorld*)h; const KEnv& e = w->e;
    for (int i = 0; i < NB; ++i) { out[i] = e.q[i]; out[NB + i] = e.qd[i]; }
    out[24] = e.qb;
    for (int a = 0; a < 3; ++a) out[25 + a] = e.gripper_pos[a];
    out[28] = e.button_base[0]; out[29] = e.button_base[1]; out[30] = e.button_base[2] + w->m.sc[KM_SC_GLIDER_Z] + e.qb;
    out[31] = e.contact_button; out[32] = e.contact_table;
}

} /* extern "C" */
