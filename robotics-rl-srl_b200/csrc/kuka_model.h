/*
 * kuka_model.h -- layout of the flat model blob (array of float64) that the URDF loader
 * (srl_sim/model.py) produces and srl_sim_create() consumes for the Kuka env kinds.
 *
 * It carries what the reference obtains by loading assets into PyBullet at every reset
 * (environments/kuka_gym/kuka.py:60-71: kuka_with_gripper2.sdf; kuka_button_gym_env.py:221-239: plane,
 * table, simple_button.urdf, gravity) plus the controller constants of kuka.py:22-26,46-53,73,167-187.
 * A data-format definition only (no algorithm); the CPU oracle reads the same blob.
 *
 * Bodies are the 12 MOVABLE links of the 14-joint model (links behind fixed joints are merged into
 * their parent by the loader), in PyBullet joint-index order 0-8,10,11,13.  Topology is fixed:
 * a chain of 8 (arm joints 0-6, gripper yaw 7) that forks into two 2-link fingers (8->10, 11->13).
 */
#ifndef KUKA_MODEL_H_
#define KUKA_MODEL_H_

#define KM_MAGIC 1397902411.0 /* 'SRLK' */
#define KM_VERSION 1.0
#define KM_NBODY 12
#define KM_MAX_SPHERES 16

/* header (doubles) */
#define KM_H_MAGIC 0
#define KM_H_VERSION 1
#define KM_H_NBODY 2
#define KM_H_NSPHERE 3
#define KM_H_BODY_OFF 4
#define KM_H_CTRL_OFF 5
#define KM_H_SPHERE_OFF 6
#define KM_H_SCENE_OFF 7
#define KM_H_TOTAL 8
#define KM_HEADER_SIZE 16

/* body record */
#define KM_BODY_STRIDE 36
#define KM_B_PARENT 0    /* parent body index, -1 = fixed base                                    */
#define KM_B_JTYPE 1     /* 0 revolute, 1 prismatic                                              */
#define KM_B_ORIGIN 2    /* [3] joint origin in the parent body frame                            */
#define KM_B_ROT 5       /* [9] row-major rotation parent <- child at q = 0                       */
#define KM_B_AXIS 14     /* [3] joint axis in the child frame                                     */
#define KM_B_MASS 17
#define KM_B_COM 18      /* [3] centre of mass in the body frame                                  */
#define KM_B_INERTIA 21  /* [6] xx xy xz yy yz zz about the COM, body-frame axes                   */
#define KM_B_LOWER 27
#define KM_B_UPPER 28
#define KM_B_DAMPING 29  /* URDF <dynamics damping>                                               */
#define KM_B_QINIT 30    /* kuka.py:65-66                                                         */
#define KM_B_REFJOINT 31 /* PyBullet joint index (0..13) of this body's joint                     */

/* controller record per body (kuka.py:165-187) */
#define KM_CTRL_STRIDE 8
#define KM_C_KP 0
#define KM_C_KD 1
#define KM_C_MAXFORCE 2
#define KM_C_MAXVEL 3    /* <= 0: no velocity clamp                                              */
#define KM_C_TARGET 4    /* 0: IK solution, 1: end_effector_angle, 2: -finger_angle, 3: +finger_angle, 4: zero */

/* collision sphere record */
#define KM_SPHERE_STRIDE 6
#define KM_S_BODY 0
#define KM_S_CENTER 1    /* [3] body frame                                                        */
#define KM_S_RADIUS 4

/* scene record */
#define KM_SCENE_SIZE 80
#define KM_SC_BASE_POS 0        /* [3] kuka base, kuka.py:63                                        */
#define KM_SC_GRAVITY_Z 3       /* -10, kuka_button_gym_env.py:236                                  */
#define KM_SC_TIMESTEP 4        /* 1/240, :86,220                                                   */
#define KM_SC_SOLVER_ITERS 5    /* 150, :219                                                        */
#define KM_SC_TABLE_TOP_Z 6     /* world z of the table top surface                                 */
#define KM_SC_TABLE_XMIN 7
#define KM_SC_TABLE_XMAX 8
#define KM_SC_TABLE_YMIN 9
#define KM_SC_TABLE_YMAX 10
#define KM_SC_BUTTON_BASE 11    /* [3] default button base origin at rest (x, y :227-228; z settled on the table) */
#define KM_SC_GLIDER_Z 14       /* glider joint origin z in the base frame (simple_button.urdf)     */
#define KM_SC_GLIDER_LOWER 15
#define KM_SC_GLIDER_UPPER 16
#define KM_SC_BUTTON_MASS 17
#define KM_SC_DISC_RADIUS 18    /* button link collision cylinder                                   */
#define KM_SC_DISC_Z0 19        /* bottom / top of the disc in the button link frame                */
#define KM_SC_DISC_Z1 20
#define KM_SC_STACK_RADIUS 21   /* base + fixed cylinder stack (button_uid links -1, 0)             */
#define KM_SC_STACK_TOP 22      /* top of the stack in the base frame                               */
#define KM_SC_CONTACT_DIST 23   /* manifold margin: a contact point exists below this distance      */
#define KM_SC_FRICTION 24       /* combined lateral friction coefficient                            */
#define KM_SC_ERP 25            /* contact / limit error reduction                                  */
#define KM_SC_LIN_DAMPING 26    /* btMultiBody link damping                                         */
#define KM_SC_ANG_DAMPING 27
#define KM_SC_EE_INIT 28        /* [3] kuka.py:73                                                   */
#define KM_SC_BOX_SMALL 31      /* [6] minx maxx miny maxy minz maxz, kuka.py:46-49                 */
#define KM_SC_BOX_LARGE 37      /* [6] kuka.py:51-53                                                */
#define KM_SC_IK_QUAT 43        /* [4] x y z w target orientation, kuka.py:144                      */
#define KM_SC_IK_DAMPING 47     /* kuka.py:42-43                                                    */
#define KM_SC_EE_BODY 48        /* kuka.py:31                                                       */
#define KM_SC_GRIPPER_BODY 49   /* kuka.py:32 (joint index 8 -> body 8)                             */
#define KM_SC_TARGET_HEIGHT 50  /* BUTTON_DISTANCE_HEIGHT, kuka_button_gym_env.py:35                */
#define KM_SC_RAND_X 51         /* 0.15, :230                                                       */
#define KM_SC_RAND_Y 52         /* 0.3, :231                                                        */
#define KM_SC_BTN_IDLE_IMPULSE 53  /* max impulse of the default (velocity-0) joint motor           */
#define KM_SC_BTN_KP 54         /* armed POSITION_CONTROL motor, :347 (defaults kp 0.1, kd 1)       */
#define KM_SC_BTN_KD 55
#define KM_SC_BTN_TARGET 56
#define KM_SC_BTN_MAXFORCE 57
#define KM_SC_LIMIT_MAX_IMPULSE 58
#define KM_SC_MAX_CONTACTS 59   /* cap on contact rows per step                                     */
#define KM_SC_LIMIT_EPS 60      /* a joint-limit row is active while (q - limit) <= this            */

#endif
