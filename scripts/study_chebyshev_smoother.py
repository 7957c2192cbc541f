"""Planning aid (CPU, numpy; not product code): would a Chebyshev smoother inside the additive two-level preconditioner
(k extra operator applications per PCG iteration, no extra reductions) pay for the cluster kernel?  Prints outer
iterations and total operator applications for the C4 micro cell.  Result (cluster kernel): no --
110 -> 82 (k=1) -> 73 (k=2) outer iterations, i.e. 1.5x / 2x the operator applications."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spl
import coefficients as Cf
from oracle import hmm_oracle as ho
from oracle import meshes, npufl
n = 8
mic = ho.MicroCell(meshes.create_unit_cube(n, n, n), "elasticity", 0)
x = np.array([0.37, 0.21, 0.05])
which = sys.argv[1] if len(sys.argv) > 1 else "fibre"
M = np.asarray(Cf.dtheta_rotation_3d(npufl)(x))[..., 0]
coef = Cf.hooke_fibre_3d(npufl) if which == "fibre" else getattr(Cf, which)(npufl)
Abar = mic.element_coefficient(coef, x)
K, B = mic.assemble(Abar, M)
ne = len(mic.vol)
rhs = [mic.load(Abar, B, np.broadcast_to(E, (ne, 3, 3))) for E in ho.unit_strains(3)]
N = mic.n_per
Kd = K.tocsr()
Dinv = sp.block_diag([sp.csr_matrix(np.linalg.inv(Kd[3*i:3*i+3, 3*i:3*i+3].toarray())) for i in range(N)], format="csr")
X = mic.mesh.x
pc = np.zeros((N, 3))
for v in range(len(X)): pc[mic.node2per[v]] = X[v] % 1.0
def coarse_semi(m, ax):
    H = 1.0 / m; cols = []
    rng = [range(m) if a != ax else range(1) for a in range(3)]
    for cz in rng[2]:
        for cy in rng[1]:
            for cx in rng[0]:
                c = np.array([cx, cy, cz]) * H
                d = np.abs(pc - c); d = np.minimum(d, 1 - d)
                wa = np.clip(1 - d / H, 0, None); wa[:, ax] = 1.0
                w = np.prod(wa, axis=1)
                for k in range(3):
                    v = np.zeros(3 * N); v[k::3] = w; cols.append(v)
    return np.array(cols).T
def pcg(b, prec, rtol=1e-8, maxit=3000):
    xk = np.zeros_like(b); r = b.copy(); z = prec(r); p = z.copy(); rz = r @ z; rz0 = rz
    for it in range(1, maxit):
        Ap = K @ p; a = rz / (p @ Ap); xk += a * p; r -= a * Ap
        z = prec(r); rzn = r @ z
        if rzn <= rtol**2 * rz0: return it
        p = z + (rzn / rz) * p; rz = rzn
    return maxit
# lambda max of Dinv K by power iteration
DK = Dinv @ Kd
v = np.random.default_rng(0).standard_normal(3*N)
for _ in range(200): v = DK @ v; v /= np.linalg.norm(v)
lmax = v @ (DK @ v)
print("lmax(Dinv K) ~", lmax)
def cheb(k, lo_frac, lmx):
    # Chebyshev polynomial approx of (Dinv K)^-1 Dinv on [lmx*lo_frac, lmx], degree k (k matvecs), returns prec
    a, b = lmx * lo_frac, lmx * 1.05
    theta, delta = (b + a) / 2, (b - a) / 2
    def prec(r):
        # standard Chebyshev iteration for K z = r with preconditioner Dinv, zero initial guess, k+1 terms
        z = np.zeros_like(r); res = r.copy()
        sigma = theta / delta; rho = 1 / sigma
        d = (Dinv @ res) / theta
        z = z + d
        for i in range(k):
            res = r - K @ z
            rho_n = 1 / (2 * sigma - rho)
            d = rho_n * rho * d + (2 * rho_n / delta) * (Dinv @ res)
            rho = rho_n
            z = z + d
        return z
    return prec
for ax in ([0, 1, 2] if which == "fibre" else [None]):
    if ax is None:
        add = lambda sm: sm
        label = "no coarse"
    else:
        P = coarse_semi(4, ax); E = P.T @ (K @ P); Einv = np.linalg.pinv(E)
        add = (lambda sm, P=P, Einv=Einv: (lambda r: sm(r) + P @ (Einv @ (P.T @ r))))
        label = f"semi ax{ax}"
    base = [pcg(b, add(lambda r: Dinv @ r)) for b in rhs[:3]]
    print(label, "jacobi+coarse iters", base, flush=True)
    for k in (1, 2, 3):
        for lo in (0.1, 0.25):
            its = [pcg(b, add(cheb(k, lo, lmax))) for b in rhs[:3]]
            print(f"  cheb k={k} lo={lo}: outer {its}  matvecs {[i*(k+1) for i in its]}", flush=True)
