"""Round-2 planning aid (CPU, numpy): PCG iteration counts for the C4 micro cell (8^3 elasticity, fibre mu 100/0.001,
lambda 1, rotated) with candidate preconditioners.  Not product code."""
import sys

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import numpy as np
import scipy.sparse as sp

import coefficients as Cf
from oracle import hmm_oracle as ho
from oracle import meshes, npufl

n = 8
mic = ho.MicroCell(meshes.create_unit_cube(n, n, n), "elasticity", 0)
x = np.array([0.37, 0.21, 0.05])
M = np.asarray(Cf.dtheta_rotation_3d(npufl)(x))[..., 0]
Abar = mic.element_coefficient(Cf.hooke_fibre_3d(npufl), x)
K, B = mic.assemble(Abar, M)
ne = len(mic.vol)
rhs = [mic.load(Abar, B, np.broadcast_to(E, (ne, 3, 3))) for E in ho.unit_strains(3)]
N = mic.n_per
Kd = K.tocsr()
Dinv = sp.block_diag([sp.csr_matrix(np.linalg.inv(Kd[3 * i : 3 * i + 3, 3 * i : 3 * i + 3].toarray())) for i in range(N)], format="csr")
X = mic.mesh.x
pc = np.zeros((N, 3))
for v in range(len(X)):
    pc[mic.node2per[v]] = X[v] % 1.0


def coarse(m):
    H = 1.0 / m
    cols = []
    for cz in range(m):
        for cy in range(m):
            for cx in range(m):
                c = np.array([cx, cy, cz]) * H
                d = np.abs(pc - c)
                d = np.minimum(d, 1 - d)
                w = np.prod(np.clip(1 - d / H, 0, None), axis=1)
                for k in range(3):
                    v = np.zeros(3 * N)
                    v[k::3] = w
                    cols.append(v)
    return np.array(cols).T


def pcg(b, prec, rtol=1e-8, maxit=3000):
    xk = np.zeros_like(b)
    r = b.copy()
    z = prec(r)
    p = z.copy()
    rz = r @ z
    rz0 = rz
    for it in range(1, maxit):
        Ap = K @ p
        a = rz / (p @ Ap)
        xk += a * p
        r -= a * Ap
        z = prec(r)
        rzn = r @ z
        if rzn <= rtol**2 * rz0:
            return it
        p = z + (rzn / rz) * p
        rz = rzn
    return maxit


def two_level_additive(m):
    P = coarse(m)
    E = P.T @ (K @ P)
    Ei = np.linalg.pinv(E, rcond=1e-12)
    return lambda r: Dinv @ r + P @ (Ei @ (P.T @ r))


def two_level_multiplicative(m, nu=1, omega=0.7):
    P = coarse(m)
    E = P.T @ (K @ P)
    Ei = np.linalg.pinv(E, rcond=1e-12)

    def prec(r):
        z = omega * (Dinv @ r)
        for _ in range(nu - 1):
            z += omega * (Dinv @ (r - K @ z))
        z += P @ (Ei @ (P.T @ (r - K @ z)))
        for _ in range(nu):
            z += omega * (Dinv @ (r - K @ z))
        return z

    return prec


b = rhs[3]
print("block Jacobi                         :", pcg(b, lambda r: Dinv @ r))
for m in (2, 4):
    print(f"additive two-level, coarse {m}^3 P1     :", pcg(b, two_level_additive(m)))
for m in (2, 4):
    print(f"symmetric V(1,1) two-grid, coarse {m}^3 :", pcg(b, two_level_multiplicative(m)), "(3 extra applies of K per iteration)")
