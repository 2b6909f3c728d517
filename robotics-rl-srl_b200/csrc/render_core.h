// Image observations: per-pixel arithmetic of the batched ray-caster, shared by the sm_100a kernels
// (csrc/render_kernels.cu) and the CPU checker (oracle/oracle_render.cpp, test infrastructure).
//
// What the reference does: `render(mode='rgb_array')` asks PyBullet's TinyRenderer for a 224 x 224 frame of the scene through
// computeViewMatrixFromYawPitchRoll / computeProjectionMatrixFOV / getCameraImage (environments/kuka_gym/kuka_button_gym_env.py:370-420,
// environments/mobile_robot/mobile_robot_env.py:287-334), and `srl_model="raw_pixels"` returns that frame as the observation.
// What this is: a ray-caster of ANALYTIC primitives -- plane, sphere, capsule, upright cylinder, yaw-rotated box -- with ambient +
// Lambert shading, through the same pinhole camera.  The meshes and textures TinyRenderer draws (pybullet_data: kuka_iiwa/meshes, table,
// racecar, plane checker) are absent from the reference checkout and from this image, so the arm is drawn as capsules between its joint
// frames plus its collision spheres, the table / walls as boxes, the buttons / targets as the z-extruded discs their collada meshes are:
// same camera, same layout, same colours where the reference's own URDFs give them (urdf/wall.urdf, urdf/cylinder.urdf,
// urdf/simple_button.urdf), but NOT TinyRenderer's pixels -- parity with reference frames is unpinned and cannot be pinned offline.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define SRL_RHD __host__ __device__ __forceinline__
#else
#define SRL_RHD inline
#endif

#define SRL_PRIM_WORDS 16
#define SRL_MAX_PRIMS 40
enum { SRL_PRIM_PLANE = 0, SRL_PRIM_SPHERE = 1, SRL_PRIM_CAPSULE = 2, SRL_PRIM_CYL = 3, SRL_PRIM_BOX = 4 };

// One primitive = 16 floats: type, 11 geometry words, r g b, pad.
//   PLANE   a0 = z, a1 = checker period (0: plain), a2..4 = second colour of the checker
//   SPHERE  a0..2 centre, a3 radius
//   CAPSULE a0..2 end 0, a3..5 end 1, a6 radius
//   CYL     a0 cx, a1 cy, a2 z0, a3 z1, a4 radius            (upright, capped)
//   BOX     a0..2 centre, a3..5 half extents, a6 cos, a7 sin  (rotated about z)
struct SrlPrim { float type, a[11], r, g, b, pad; };

struct SrlCam { float eye[3], fwd[3], right[3], up[3], tan_half_fov, aspect; };

SRL_RHD void srl_prim_set(SrlPrim& p, int type, float r, float g, float b) {
    p.type = (float)type; p.r = r; p.g = g; p.b = b; p.pad = 0.f;
    for (int i = 0; i < 11; ++i) p.a[i] = 0.f;
}
SRL_RHD void srl_prim_capsule(SrlPrim& p, const float* e0, const float* e1, float rad, float r, float g, float b) {
    srl_prim_set(p, SRL_PRIM_CAPSULE, r, g, b);
    for (int i = 0; i < 3; ++i) { p.a[i] = e0[i]; p.a[3 + i] = e1[i]; }
    p.a[6] = rad;
}
SRL_RHD void srl_prim_sphere(SrlPrim& p, const float* c, float rad, float r, float g, float b) {
    srl_prim_set(p, SRL_PRIM_SPHERE, r, g, b);
    for (int i = 0; i < 3; ++i) p.a[i] = c[i];
    p.a[3] = rad;
}
SRL_RHD void srl_prim_cyl(SrlPrim& p, float cx, float cy, float z0, float z1, float rad, float r, float g, float b) {
    srl_prim_set(p, SRL_PRIM_CYL, r, g, b);
    p.a[0] = cx; p.a[1] = cy; p.a[2] = z0; p.a[3] = z1; p.a[4] = rad;
}
SRL_RHD void srl_prim_box(SrlPrim& p, float cx, float cy, float cz, float hx, float hy, float hz, float yaw_cos, float yaw_sin, float r, float g, float b) {
    srl_prim_set(p, SRL_PRIM_BOX, r, g, b);
    p.a[0] = cx; p.a[1] = cy; p.a[2] = cz; p.a[3] = hx; p.a[4] = hy; p.a[5] = hz; p.a[6] = yaw_cos; p.a[7] = yaw_sin;
}
SRL_RHD void srl_prim_plane(SrlPrim& p, float z, float checker, float r, float g, float b) {
    srl_prim_set(p, SRL_PRIM_PLANE, r, g, b);
    p.a[0] = z; p.a[1] = checker; p.a[2] = 0.68f; p.a[3] = 0.77f; p.a[4] = 0.93f;     // plane.urdf's checker is white / light blue (imgs/kuka.gif, imgs/mobile_robot.gif)
}

// ---- camera: pybullet's computeViewMatrixFromYawPitchRoll (upAxisIndex = 2) + computeProjectionMatrixFOV, as eye + basis (RECALLED from
//      PhysicsClientC_API.cpp: eye = target + Rz(yaw) Ry(roll) Rx(pitch) (0, -distance, 0), up = the same rotation of (0, 0, 1)) ----
SRL_RHD void srl_camera_setup(const float* target, float distance, float yaw_deg, float pitch_deg, float roll_deg, float fov_deg, float aspect, SrlCam& c) {
    const float d2r = 0.01745329251994329547f;
    const float cy = cosf(yaw_deg * d2r), sy = sinf(yaw_deg * d2r), cp = cosf(pitch_deg * d2r), sp = sinf(pitch_deg * d2r);
    const float cr = cosf(roll_deg * d2r), sr = sinf(roll_deg * d2r);
    // R = Rz(yaw) Ry(roll) Rx(pitch)
    const float R[9] = {cy * cr, cy * sr * sp - sy * cp, cy * sr * cp + sy * sp,
                        sy * cr, sy * sr * sp + cy * cp, sy * sr * cp - cy * sp,
                        -sr, cr * sp, cr * cp};
    const float e[3] = {0.f, -distance, 0.f};
    float up0[3], f[3];
    for (int i = 0; i < 3; ++i) {
        c.eye[i] = target[i] + R[3 * i] * e[0] + R[3 * i + 1] * e[1] + R[3 * i + 2] * e[2];
        up0[i] = R[3 * i + 2];
        f[i] = target[i] - c.eye[i];
    }
    const float fl = sqrtf(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]);
    for (int i = 0; i < 3; ++i) c.fwd[i] = f[i] / fl;
    // right = fwd x up0, up = right x fwd (the lookAt basis of b3ComputeViewMatrixFromPositions)
    float s[3] = {c.fwd[1] * up0[2] - c.fwd[2] * up0[1], c.fwd[2] * up0[0] - c.fwd[0] * up0[2], c.fwd[0] * up0[1] - c.fwd[1] * up0[0]};
    const float sl = sqrtf(s[0] * s[0] + s[1] * s[1] + s[2] * s[2]);
    for (int i = 0; i < 3; ++i) c.right[i] = s[i] / sl;
    c.up[0] = c.right[1] * c.fwd[2] - c.right[2] * c.fwd[1];
    c.up[1] = c.right[2] * c.fwd[0] - c.right[0] * c.fwd[2];
    c.up[2] = c.right[0] * c.fwd[1] - c.right[1] * c.fwd[0];
    c.tan_half_fov = tanf(0.5f * fov_deg * d2r);
    c.aspect = aspect;
}

// ---- ray / primitive intersections: nearest t > tmin, outward normal ----
SRL_RHD bool srl_hit_sphere(const float* o, const float* d, const float* c, float r, float& t, float* n) {
    const float ox = o[0] - c[0], oy = o[1] - c[1], oz = o[2] - c[2];
    const float b = ox * d[0] + oy * d[1] + oz * d[2], cc = ox * ox + oy * oy + oz * oz - r * r;
    const float h = b * b - cc;
    if (h < 0.f) return false;
    const float tt = -b - sqrtf(h);
    if (tt <= 1e-4f) return false;
    t = tt;
    n[0] = (ox + tt * d[0]) / r; n[1] = (oy + tt * d[1]) / r; n[2] = (oz + tt * d[2]) / r;
    return true;
}
SRL_RHD bool srl_hit_capsule(const float* o, const float* d, const float* pa, const float* pb, float r, float& t, float* n) {
    // closed form for a capped-by-spheres cylinder (the segment pa-pb swept by a sphere of radius r)
    const float ba[3] = {pb[0] - pa[0], pb[1] - pa[1], pb[2] - pa[2]}, oa[3] = {o[0] - pa[0], o[1] - pa[1], o[2] - pa[2]};
    const float baba = ba[0] * ba[0] + ba[1] * ba[1] + ba[2] * ba[2];
    if (baba < 1e-12f) return srl_hit_sphere(o, d, pa, r, t, n);
    const float bard = ba[0] * d[0] + ba[1] * d[1] + ba[2] * d[2], baoa = ba[0] * oa[0] + ba[1] * oa[1] + ba[2] * oa[2];
    const float rdoa = d[0] * oa[0] + d[1] * oa[1] + d[2] * oa[2], oaoa = oa[0] * oa[0] + oa[1] * oa[1] + oa[2] * oa[2];
    const float a = baba - bard * bard;
    float b = baba * rdoa - baoa * bard, c = baba * oaoa - baoa * baoa - r * r * baba;
    float h = b * b - a * c;
    float tt = -1.f;
    if (h >= 0.f && a > 1e-12f) {
        const float t0 = (-b - sqrtf(h)) / a;
        const float y = baoa + t0 * bard;
        if (y > 0.f && y < baba && t0 > 1e-4f) {      // body
            tt = t0;
            const float k = y / baba;
            n[0] = (oa[0] + t0 * d[0] - ba[0] * k) / r; n[1] = (oa[1] + t0 * d[1] - ba[1] * k) / r; n[2] = (oa[2] + t0 * d[2] - ba[2] * k) / r;
        }
    }
    if (tt < 0.f) {                                    // caps: the nearer of the two end spheres
        float t1, n1[3], t2, n2[3];
        const bool h1 = srl_hit_sphere(o, d, pa, r, t1, n1), h2 = srl_hit_sphere(o, d, pb, r, t2, n2);
        if (h1 && (!h2 || t1 <= t2)) { tt = t1; n[0] = n1[0]; n[1] = n1[1]; n[2] = n1[2]; }
        else if (h2) { tt = t2; n[0] = n2[0]; n[1] = n2[1]; n[2] = n2[2]; }
        else return false;
    }
    t = tt;
    return true;
}
SRL_RHD bool srl_hit_cyl(const float* o, const float* d, float cx, float cy, float z0, float z1, float r, float& t, float* n) {
    float best = 1e30f;
    const float ox = o[0] - cx, oy = o[1] - cy;
    const float a = d[0] * d[0] + d[1] * d[1];
    if (a > 1e-12f) {                                  // side
        const float b = ox * d[0] + oy * d[1], c = ox * ox + oy * oy - r * r;
        const float h = b * b - a * c;
        if (h >= 0.f) {
            const float tt = (-b - sqrtf(h)) / a;
            const float z = o[2] + tt * d[2];
            if (tt > 1e-4f && z >= z0 && z <= z1) { best = tt; n[0] = (ox + tt * d[0]) / r; n[1] = (oy + tt * d[1]) / r; n[2] = 0.f; }
        }
    }
    if (fabsf(d[2]) > 1e-12f) {                        // caps
        const float zc = d[2] < 0.f ? z1 : z0;
        const float tt = (zc - o[2]) / d[2];
        const float x = ox + tt * d[0], y = oy + tt * d[1];
        if (tt > 1e-4f && tt < best && x * x + y * y <= r * r) { best = tt; n[0] = 0.f; n[1] = 0.f; n[2] = d[2] < 0.f ? 1.f : -1.f; }
    }
    if (best > 1e29f) return false;
    t = best;
    return true;
}
SRL_RHD bool srl_hit_box(const float* o, const float* d, const float* a, float& t, float* n) {
    // into the box frame (rotation about z by the box's yaw)
    const float cs = a[6], sn = a[7];
    const float px = o[0] - a[0], py = o[1] - a[1], pz = o[2] - a[2];
    const float lo[3] = {cs * px + sn * py, -sn * px + cs * py, pz}, ld[3] = {cs * d[0] + sn * d[1], -sn * d[0] + cs * d[1], d[2]};
    float tn = -1e30f, tf = 1e30f; int axis = 0; float sign = 1.f;
    for (int k = 0; k < 3; ++k) {
        if (fabsf(ld[k]) < 1e-12f) { if (fabsf(lo[k]) > a[3 + k]) return false; continue; }
        const float inv = 1.f / ld[k];
        float t0 = (-a[3 + k] - lo[k]) * inv, t1 = (a[3 + k] - lo[k]) * inv;
        float sg = -1.f;
        if (t0 > t1) { const float tmp = t0; t0 = t1; t1 = tmp; sg = 1.f; }
        if (t0 > tn) { tn = t0; axis = k; sign = sg; }
        if (t1 < tf) tf = t1;
    }
    if (tn > tf || tn <= 1e-4f) return false;
    t = tn;
    const float ln[3] = {axis == 0 ? sign : 0.f, axis == 1 ? sign : 0.f, axis == 2 ? sign : 0.f};
    n[0] = cs * ln[0] - sn * ln[1]; n[1] = sn * ln[0] + cs * ln[1]; n[2] = ln[2];
    return true;
}

// One pixel: nearest hit over the primitive list, ambient + Lambert shading, 8-bit RGB.  Row 0 is the TOP of the image (getCameraImage).
SRL_RHD void srl_render_pixel(const SrlCam& c, const SrlPrim* prims, int np, int x, int y, int W, int H, uint8_t* rgb) {
    const float u = (2.f * ((float)x + 0.5f) / (float)W - 1.f) * c.tan_half_fov * c.aspect;
    const float v = (1.f - 2.f * ((float)y + 0.5f) / (float)H) * c.tan_half_fov;
    float d[3] = {c.fwd[0] + u * c.right[0] + v * c.up[0], c.fwd[1] + u * c.right[1] + v * c.up[1], c.fwd[2] + u * c.right[2] + v * c.up[2]};
    const float dl = 1.f / sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    d[0] *= dl; d[1] *= dl; d[2] *= dl;
    float best = 1e30f, bn[3] = {0.f, 0.f, 1.f}, col[3] = {0.84f, 0.89f, 0.95f};   // background (above the horizon)
    bool hit = false;
    for (int k = 0; k < np; ++k) {
        const SrlPrim& p = prims[k];
        const int type = (int)p.type;
        float t = 0.f, n[3] = {0.f, 0.f, 1.f};
        bool h = false;
        float cr = p.r, cg = p.g, cb = p.b;
        if (type == SRL_PRIM_PLANE) {
            if (d[2] < -1e-6f) {
                t = (p.a[0] - c.eye[2]) / d[2];
                h = t > 1e-4f;
                if (h && p.a[1] > 0.f) {
                    const float px = c.eye[0] + t * d[0], py = c.eye[1] + t * d[1];
                    const int ix = (int)floorf(px / p.a[1]), iy = (int)floorf(py / p.a[1]);
                    if ((ix + iy) & 1) { cr = p.a[2]; cg = p.a[3]; cb = p.a[4]; }
                }
            }
        } else if (type == SRL_PRIM_SPHERE) h = srl_hit_sphere(c.eye, d, p.a, p.a[3], t, n);
        else if (type == SRL_PRIM_CAPSULE) h = srl_hit_capsule(c.eye, d, p.a, p.a + 3, p.a[6], t, n);
        else if (type == SRL_PRIM_CYL) h = srl_hit_cyl(c.eye, d, p.a[0], p.a[1], p.a[2], p.a[3], p.a[4], t, n);
        else h = srl_hit_box(c.eye, d, p.a, t, n);
        if (h && t < best) { best = t; hit = true; bn[0] = n[0]; bn[1] = n[1]; bn[2] = n[2]; col[0] = cr; col[1] = cg; col[2] = cb; }
    }
    float shade = 1.f;
    if (hit) {
        const float L[3] = {0.3713907f, 0.5570860f, 0.7427814f};     // normalised (2, 3, 4): one fixed directional light, no shadows
        const float nl = bn[0] * L[0] + bn[1] * L[1] + bn[2] * L[2];
        shade = 0.55f + 0.45f * (nl > 0.f ? nl : 0.f);
    }
    for (int k = 0; k < 3; ++k) {
        float vv = col[k] * shade * 255.f + 0.5f;
        vv = vv < 0.f ? 0.f : vv > 255.f ? 255.f : vv;
        rgb[k] = (uint8_t)vv;
    }
}

// ---- scene builders (shared so that both implementations draw the same list) ------------------------------------------------------
struct SrlKukaSceneConst {      // from the model blob / KukaParams
    float base[3];
    float table_z, txmin, txmax, tymin, tymax;
    float glider_z, disc_r, disc_z0, disc_z1, stack_r, stack_top;
    int two_buttons;
};

// joint_p: world origins of the 12 movable joint frames (bodies 0..7 chain, 8-9 finger A, 10-11 finger B); sph: world centres + radii of the
// collision spheres of the gripper bodies (body >= 7).  Returns the number of primitives written (<= SRL_MAX_PRIMS).
SRL_RHD int srl_kuka_scene(const SrlKukaSceneConst& K, const float* joint_p, const float* sph, int nsph, float bbx, float bby, float bbz, float qb,
                           float bb2x, float bb2y, float b2z, float qb2, SrlPrim* out) {
    int n = 0;
    srl_prim_plane(out[n++], -1.0f, 1.0f, 1.0f, 1.0f, 1.0f);                                            // plane.urdf at z = -1 (kuka_button_gym_env.py:222)
    const float tcx = 0.5f * (K.txmin + K.txmax), tcy = 0.5f * (K.tymin + K.tymax), thx = 0.5f * (K.txmax - K.txmin), thy = 0.5f * (K.tymax - K.tymin);
    srl_prim_box(out[n++], tcx, tcy, K.table_z - 0.025f, thx, thy, 0.025f, 1.f, 0.f, 0.92f, 0.82f, 0.68f);   // table top slab (5 cm), light wood
    for (int k = 0; k < 4; ++k)                                                                           // legs down to the plane
        srl_prim_box(out[n++], tcx + ((k & 1) ? 1.f : -1.f) * (thx - 0.1f), tcy + ((k & 2) ? 1.f : -1.f) * (thy - 0.1f), 0.5f * (K.table_z - 0.05f - 1.0f),
                     0.05f, 0.05f, 0.5f * (K.table_z - 0.05f + 1.0f), 1.f, 0.f, 0.85f, 0.75f, 0.62f);
    // button(s): base + fixed cylinder stack (green), pressable disc (yellow) -- colours of urdf/simple_button.urdf
    for (int b = 0; b < (K.two_buttons ? 2 : 1); ++b) {
        const float x = b ? bb2x : bbx, y = b ? bb2y : bby, z = b ? b2z : bbz, q = b ? qb2 : qb;
        srl_prim_cyl(out[n++], x, y, z, z + K.stack_top, K.stack_r, 0.f, 1.f, 0.f);
        srl_prim_cyl(out[n++], x, y, z + K.glider_z + q + K.disc_z0, z + K.glider_z + q + K.disc_z1, K.disc_r, 1.f, 1.f, 0.f);
    }
    // arm: fixed pedestal, then a capsule per link between consecutive joint frames (iiwa orange / grey, RECALLED materials)
    srl_prim_capsule(out[n++], K.base, joint_p, 0.075f, 0.30f, 0.30f, 0.30f);
    for (int i = 0; i < 7; ++i) {
        const bool orange = (i & 1) == 0;          // orange links alternating with blue-grey ones, as in imgs/kuka.gif
        srl_prim_capsule(out[n++], joint_p + 3 * i, joint_p + 3 * (i + 1), i < 4 ? 0.065f : 0.055f, orange ? 1.0f : 0.5f, orange ? 0.42f : 0.7f, orange ? 0.04f : 1.0f);
    }
    // gripper: base to the two fingers, finger links, plus the collision spheres of the gripper bodies
    srl_prim_capsule(out[n++], joint_p + 21, joint_p + 24, 0.02f, 0.15f, 0.15f, 0.15f);
    srl_prim_capsule(out[n++], joint_p + 24, joint_p + 27, 0.012f, 0.15f, 0.15f, 0.15f);
    srl_prim_capsule(out[n++], joint_p + 21, joint_p + 30, 0.02f, 0.15f, 0.15f, 0.15f);
    srl_prim_capsule(out[n++], joint_p + 30, joint_p + 33, 0.012f, 0.15f, 0.15f, 0.15f);
    for (int k = 0; k < nsph && n < SRL_MAX_PRIMS; ++k) srl_prim_sphere(out[n++], sph + 4 * k, sph[4 * k + 3], 0.2f, 0.2f, 0.2f);
    return n;
}

// MobileRobot family: plane, four walls (urdf/wall.urdf: box 4 x 0.1 x 0.1; left red, bottom black, right green, top blue:
// mobile_robot_env.py:184-203), the robot as a box of the racecar's footprint, the target(s).
// kind: 0 base, 3 one-dimensional, 1 two targets, 2 line target (urdf/wall_target.urdf: box 4 x 0.5 x 0.1, yellow, rotated by pi / 2).
SRL_RHD int srl_mobile_scene(int kind, float rx, float ry, float t0x, float t0y, float t1x, float t1y, SrlPrim* out) {
    int n = 0;
    srl_prim_plane(out[n++], 0.f, 1.0f, 1.0f, 1.0f, 1.0f);
    srl_prim_box(out[n++], 2.f, 0.f, 0.f, 2.f, 0.05f, 0.05f, 1.f, 0.f, 0.8f, 0.f, 0.f);
    if (kind != 3) {            // the 1-D variant only has the left wall (mobile_robot_1D_env.py:84-86)
        srl_prim_box(out[n++], 4.f, 2.f, 0.f, 2.f, 0.05f, 0.05f, 0.f, 1.f, 0.f, 0.f, 0.f);
        srl_prim_box(out[n++], 2.f, 4.f, 0.f, 2.f, 0.05f, 0.05f, 1.f, 0.f, 0.f, 0.8f, 0.f);
        srl_prim_box(out[n++], 0.f, 2.f, 0.f, 2.f, 0.05f, 0.05f, 0.f, 1.f, 0.f, 0.f, 0.8f);
    }
    if (kind == 2) srl_prim_box(out[n++], t0x, 2.f, -0.045f, 2.f, 0.25f, 0.05f, 0.f, 1.f, 1.f, 1.f, 0.f);
    else {
        srl_prim_cyl(out[n++], t0x, t0y, 0.f, 0.03f, 0.18f, 1.f, 1.f, 0.f);                                // urdf/cylinder.urdf: the button disc mesh (r 0.09) scaled (2, 2, 1), yellow
        if (kind == 1) srl_prim_cyl(out[n++], t1x, t1y, 0.f, 0.03f, 0.18f, 0.8f, 0.f, 0.f);                 // second target recoloured red (mobile_robot_2target_env.py:71)
    }
    srl_prim_box(out[n++], rx, ry, 0.09f, 0.325f, 0.1f, 0.07f, 1.f, 0.f, 0.1f, 0.2f, 0.8f);                // racecar footprint ROBOT_LENGTH x ROBOT_WIDTH (:27-28), blue as in imgs/mobile_robot.gif
    srl_prim_box(out[n++], rx + 0.1f, ry, 0.17f, 0.12f, 0.08f, 0.03f, 1.f, 0.f, 0.95f, 0.95f, 0.95f);      // a white cabin so that the heading side is visible
    return n;
}
