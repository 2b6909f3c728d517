"""
Independent numpy reference for the rigid-body building blocks (test-only): mass matrix from body
Jacobians, gravity torque from the potential, Coriolis/centrifugal terms by numerically differentiating
the mass matrix.  Textbook Lagrangian mechanics -- shares no code with the oracle (ABA) or the CUDA kernels
(CRBA/RNEA), so agreement with both is a three-way check.
"""
import numpy as np


def ancestors(scene, b):
    out = []
    while b >= 0:
        out.append(b)
        b = scene.bodies[b].parent
    return out


def mass_matrix(scene, q):
    P, R = scene.forward_kinematics(q)
    n = len(scene.bodies)
    axes = [R[i] @ scene.bodies[i].axis for i in range(n)]
    M = np.zeros((n, n))
    for b, body in enumerate(scene.bodies):
        if body.mass <= 0:
            continue
        com = P[b] + R[b] @ body.com
        Jv = np.zeros((3, n)); Jw = np.zeros((3, n))
        for j in ancestors(scene, b):
            Jv[:, j] = np.cross(axes[j], com - P[j])
            Jw[:, j] = axes[j]
        Iw = R[b] @ body.inertia @ R[b].T
        M += body.mass * Jv.T @ Jv + Jw.T @ Iw @ Jw
    return M


def potential(scene, q, g=-10.0):
    P, R = scene.forward_kinematics(q)
    return sum(-b.mass * g * (P[i] + R[i] @ b.com)[2] for i, b in enumerate(scene.bodies))


def gravity_torque(scene, q, h=1e-6):
    n = len(q)
    G = np.zeros(n)
    for j in range(n):
        e = np.zeros(n); e[j] = h
        G[j] = (potential(scene, q + e) - potential(scene, q - e)) / (2 * h)
    return G


def coriolis(scene, q, qd, h=1e-5):
    """c(q, qd) = Mdot qd - 1/2 d(qd^T M qd)/dq, by central differences on M(q)."""
    n = len(q)
    dM = []
    for k in range(n):
        e = np.zeros(n); e[k] = h
        dM.append((mass_matrix(scene, q + e) - mass_matrix(scene, q - e)) / (2 * h))
    Mdot = sum(dM[k] * qd[k] for k in range(n))
    grad = np.array([qd @ dM[k] @ qd for k in range(n)])
    return Mdot @ qd - 0.5 * grad


def forward_dynamics(scene, q, qd):
    M = mass_matrix(scene, q)
    return np.linalg.solve(M, -coriolis(scene, q, qd) - gravity_torque(scene, q))
