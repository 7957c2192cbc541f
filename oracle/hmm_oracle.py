"""CPU restatement of the hommx hot path -- TEST INFRASTRUCTURE ONLY.

Two formulations of the same quantity, both in numpy/scipy:

* ``local_stiffness_literal``: follows ``BaseHMM._compute_local_stiffness``
  (/root/reference/src/hommx/hmm.py:334-369) line by line: ``n_b`` eps-scaled
  macro basis functions interpolated to the micro mesh (hmm.py:371-395), one
  periodic corrector per basis function (hmm.py:397-432 ->
  cell_problem.py:363-388), ``n_b^2`` integrals with the ``1/eps^2`` factor
  (hmm.py:361-364, integrands :652-667, :774-789, :905-922, :1050-1067) and the
  ``|T|/|Y|`` scaling (:367-369).
* ``cell_tensor`` / ``local_stiffness_from_tensor``: the d-RHS restatement
  (``BasePeriodicHMM.compute_effective_tensor``
  hmm.py:1219-1245 is its 1-point special case): ``d`` (Poisson) or
  ``d(d+1)/2`` (elasticity) unit-gradient/unit-strain correctors,
  ``A_hom[p,q] = 1/|Y| int A (e_q + M grad chi_q).(e_p + M grad chi_p)``,
  ``S_loc = |T| G^T A_hom G``.

Cell problems are solved with a sparse LU on the periodic system with the
constant modes pinned (the reference's tightest test does the same with MUMPS,
test/integration/test_integration_poisson.py:207-211); only gradients of the
correctors enter the result, so the pinning is immaterial.
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from . import meshes
from .quadrature import simplex_rule


# ----------------------------------------------------------------------------
# geometry helpers (hmm.py:20-28)
# ----------------------------------------------------------------------------
def simplex_volume(points):
    points = np.asarray(points, float)
    d = points.shape[0] - 1
    if d == 2:
        return 0.5 * np.linalg.norm(np.cross(points[1] - points[0], points[2] - points[0]))
    return abs(np.linalg.det([points[1] - points[0], points[2] - points[0], points[3] - points[0]])) / 6.0


def p1_gradients(verts):
    """verts (d+1, d) -> G (d, d+1): column a is grad phi_a."""
    verts = np.asarray(verts, float)
    d = verts.shape[1]
    J = (verts[1:] - verts[0]).T  # columns are edge vectors
    Jinv = np.linalg.inv(J)  # rows are grad of barycentric coords 1..d
    G = np.empty((d, d + 1))
    G[:, 1:] = Jinv.T
    G[:, 0] = -Jinv.sum(axis=0)
    return G


def unit_strains(d):
    """Basis E_q of symmetric d x d matrices: q<d -> e_q (x) e_q, then the
    pairs (0,1)[,(0,2),(1,2)] as (e_i (x) e_j + e_j (x) e_i)/2."""
    pairs = [(0, 1)] if d == 2 else [(0, 1), (0, 2), (1, 2)]
    E = []
    for q in range(d):
        m = np.zeros((d, d))
        m[q, q] = 1.0
        E.append(m)
    for i, j in pairs:
        m = np.zeros((d, d))
        m[i, j] = m[j, i] = 0.5
        E.append(m)
    return np.array(E)


def strain_coefficients(eps_mat):
    """Coefficients c with eps = sum_q c_q E_q (engineering shear)."""
    d = eps_mat.shape[0]
    pairs = [(0, 1)] if d == 2 else [(0, 1), (0, 2), (1, 2)]
    return np.array([eps_mat[q, q] for q in range(d)] + [2.0 * eps_mat[i, j] for i, j in pairs])


# ----------------------------------------------------------------------------
# micro cell: element data shared by every macro point
# ----------------------------------------------------------------------------
class MicroCell:
    """Periodic P1 micro mesh of the unit box (hmm.py:178-207).

    ``kind``: 'poisson' (bs=1) or 'elasticity' (bs=d).
    """

    def __init__(self, mesh, kind, degree, rule=None):
        """``rule`` = (points (nq, d), weights (nq,)) on the reference simplex overrides the table of ``degree``
        (e.g. ``basix.make_quadrature`` output dumped by tests/golden/make_reference_golden.py)."""
        self.mesh = mesh
        self.kind = kind
        d = self.d = mesh.dim
        self.bs = 1 if kind == "poisson" else d
        X = mesh.x[:, :d]
        cells = mesh.cells
        v = X[cells]  # (ne, d+1, d)
        J = np.transpose(v[:, 1:] - v[:, :1], (0, 2, 1))  # (ne, d, d) columns edge vectors
        self.vol = np.abs(np.linalg.det(J)) / (2 if d == 2 else 6)
        Jinv = np.linalg.inv(J)
        g = np.empty((len(cells), d + 1, d))  # g[e, a, :] = grad phi_a
        g[:, 1:, :] = Jinv
        g[:, 0, :] = -Jinv.sum(axis=1)
        self.grad = g
        qp, qw = simplex_rule(d, degree) if rule is None else (np.asarray(rule[0], float), np.asarray(rule[1], float))
        self.qw = qw / qw.sum()  # normalised: element mean
        yq = v[:, :1, :] + np.einsum("eij,qj->eqi", J, qp)  # (ne, nq, d)
        self.yq = yq
        self.Y = self.vol.sum()  # |Y| (hmm.py:101)
        master = meshes.periodic_master_map(mesh)
        uniq, inv = np.unique(master, return_inverse=True)
        self.node2per = inv  # full node -> periodic node id
        self.n_per = len(uniq)
        self.center = mesh.x.mean(axis=0)  # hmm.py:390 (mean of dof coordinates)
        # unrolled periodic dofs of each element: (ne, (d+1)*bs)
        pn = inv[cells]
        self.edofs = (pn[:, :, None] * self.bs + np.arange(self.bs)[None, None, :]).reshape(len(cells), -1)
        self.n_dof = self.n_per * self.bs

    # -- coefficient ---------------------------------------------------------
    def element_coefficient(self, A, x_macro):
        """Per-element quadrature mean of A(x_macro, y): (ne,d,d) or (ne,d,d,d,d)."""
        d = self.d
        ne, nq, _ = self.yq.shape
        y = np.zeros((3, ne * nq))
        y[:d] = self.yq.reshape(-1, d).T
        y = y[:d] if d == 2 else y
        val = A(np.asarray(x_macro, float), y)
        val = np.asarray(val, float)
        if self.kind == "poisson":
            if val.ndim <= 1:  # scalar coefficient
                val = np.broadcast_to(val.reshape(-1), (ne * nq,)) if val.ndim else np.full(ne * nq, float(val))
                val = val[None, None, :] * np.eye(d)[:, :, None]
            else:
                val = np.broadcast_to(val, (d, d, ne * nq))
            val = val.reshape(d, d, ne, nq)
            return np.einsum("ijeq,q->eij", val, self.qw)
        val = np.broadcast_to(val, (d, d, d, d, ne * nq)).reshape(d, d, d, d, ne, nq)
        return np.einsum("ijkleq,q->eijkl", val, self.qw)

    # -- element "B-matrices": strain (or gradient) of each local basis function
    def basis_fields(self, M):
        """Poisson: (ne, d+1, d) = M grad phi_a.
        Elasticity: (ne, (d+1)*d, d, d) = sym((M grad phi_a) (x) e_k)."""
        d = self.d
        Mg = np.einsum("pi,eai->eap", M, self.grad)
        if self.kind == "poisson":
            return Mg
        ne = Mg.shape[0]
        E = np.zeros((ne, d + 1, d, d, d))
        for k in range(d):
            E[:, :, k, :, k] += 0.5 * Mg  # [p, j=k]
            E[:, :, k, k, :] += 0.5 * Mg  # [j=k, p]
        return E.reshape(ne, (d + 1) * d, d, d)

    def assemble(self, Abar, M):
        """Periodic stiffness matrix (cell_problem.py:367-369) and element
        operators.  Returns (K csr, B) with B the per-element basis fields."""
        B = self.basis_fields(M)
        if self.kind == "poisson":
            # K[z_a, v_b] = |e| (M g_a) . Abar (M g_b)
            Ke = np.einsum("e,eap,epq,ebq->eab", self.vol, B, Abar, B)
        else:
            Ke = np.einsum("e,eaij,eijkl,ebkl->eab", self.vol, B, Abar, B)
        nl = Ke.shape[1]
        rows = np.repeat(self.edofs, nl, axis=1).ravel()
        cols = np.tile(self.edofs, (1, nl)).ravel()
        K = sp.coo_matrix((Ke.ravel(), (rows, cols)), shape=(self.n_dof, self.n_dof)).tocsr()
        return K, B

    def load(self, Abar, B, field):
        """-int (A field) : B_a  for a constant-per-element ``field``
        ((ne,d) gradient or (ne,d,d) strain): the cell-problem RHS
        (hmm.py:649-650, 768-772, 898-903, 1043-1048)."""
        if self.kind == "poisson":
            fe = -np.einsum("e,eap,epq,eq->ea", self.vol, B, Abar, field)
        else:
            fe = -np.einsum("e,eaij,eijkl,ekl->ea", self.vol, B, Abar, field)
        b = np.zeros(self.n_dof)
        np.add.at(b, self.edofs.ravel(), fe.ravel())
        return b

    def field_of(self, B, u):
        """M grad u (poisson) or e_D(u) (elasticity) per element for a periodic dof vector."""
        ue = u[self.edofs]
        if self.kind == "poisson":
            return np.einsum("eap,ea->ep", B, ue)
        return np.einsum("eaij,ea->eij", B, ue)

    def energy(self, Abar, F1, F2):
        """int (A F1) : F2 over the micro cell (A acts on the first slot)."""
        if self.kind == "poisson":
            return np.einsum("e,epq,eq,ep->", self.vol, Abar, F1, F2)
        return np.einsum("e,eijkl,ekl,eij->", self.vol, Abar, F1, F2)


class PinnedSolver:
    """Sparse LU of the periodic system with node 0 pinned (all components)."""

    def __init__(self, K, bs):
        n = K.shape[0]
        self.free = np.arange(bs, n)
        self.n = n
        self.lu = spla.splu(K[self.free][:, self.free].tocsc())

    def solve(self, b):
        u = np.zeros(self.n)
        u[self.free] = self.lu.solve(b[self.free])
        return u


# ----------------------------------------------------------------------------
# formulation (b): homogenised tensor
# ----------------------------------------------------------------------------
def cell_tensor(micro, A, x_macro, M=None, return_correctors=False):
    """A_hom(x_macro): (d,d) for Poisson, (m,m) with m=d(d+1)/2 for elasticity,
    in the basis returned by ``unit_strains``."""
    d = micro.d
    M = np.eye(d) if M is None else np.asarray(M, float)
    Abar = micro.element_coefficient(A, x_macro)
    K, B = micro.assemble(Abar, M)
    solver = PinnedSolver(K, micro.bs)
    ne = len(micro.vol)
    if micro.kind == "poisson":
        units = [np.broadcast_to(np.eye(d)[q], (ne, d)) for q in range(d)]
    else:
        units = [np.broadcast_to(E, (ne, d, d)) for E in unit_strains(d)]
    chis, totals = [], []
    for U in units:
        chi = solver.solve(micro.load(Abar, B, U))
        chis.append(chi)
        totals.append(U + micro.field_of(B, chi))
    m = len(units)
    Ahom = np.empty((m, m))
    for p in range(m):
        for q in range(m):
            Ahom[p, q] = micro.energy(Abar, totals[q], totals[p]) / micro.Y
    if return_correctors:
        return Ahom, chis
    return Ahom


def macro_basis_coefficients(G, kind):
    """Columns c(:, i): coefficients of grad(phi_i) (Poisson) or of the strain of
    the unrolled basis function i = a*bs + k (hmm.py:31-40) in the unit basis."""
    d, nv = G.shape
    if kind == "poisson":
        return G.copy()
    cols = []
    for a in range(nv):
        for k in range(d):
            g = np.zeros((d, d))
            g[k, :] = G[:, a]  # grad(phi_a e_k)[j,i] = delta_jk d_i phi_a
            cols.append(strain_coefficients(0.5 * (g + g.T)))
    return np.array(cols).T


def local_stiffness_from_tensor(Ahom, verts, kind):
    """S_loc[i,j] = |T| c(:,j)^T A_hom c(:,i)  (A acts on the i-slot)."""
    verts = np.asarray(verts, float)
    d = verts.shape[0] - 1
    G = p1_gradients(verts[:, :d])
    C = macro_basis_coefficients(G, kind)
    return simplex_volume(verts) * np.einsum("pj,pq,qi->ij", C, Ahom, C)


# ----------------------------------------------------------------------------
# formulation (a): literal restatement of hmm.py:334-369
# ----------------------------------------------------------------------------
def local_stiffness_literal(micro, A, verts, eps, Dtheta_t=None):
    verts = np.asarray(verts, float)  # (d+1, 3) macro cell vertex coordinates
    d = micro.d
    bs = micro.bs
    c_t = verts.mean(axis=0)  # hmm.py:350
    M = np.eye(d) if Dtheta_t is None else np.asarray(Dtheta_t(c_t), float).reshape(d, d)
    Abar = micro.element_coefficient(A, c_t)
    K, B = micro.assemble(Abar, M)
    B_plain = micro.basis_fields(np.eye(d))  # un-stratified e(.) / grad(.) for v_micro
    solver = PinnedSolver(K, bs)
    G = p1_gradients(verts[:, :d])
    nb = (d + 1) * bs
    # hmm.py:388-393: macro basis functions sampled at (y - ybar) * eps + c_T
    pts = (micro.mesh.x - micro.center) * eps + c_t
    v_fields, totals = [], []
    full_edofs = (micro.mesh.cells[:, :, None] * bs + np.arange(bs)[None, None, :]).reshape(len(micro.vol), -1)
    for i in range(nb):
        a, k = divmod(i, bs)
        phi = 1.0 / (d + 1) + (pts[:, :d] - c_t[:d]) @ G[:, a]  # affine basis function a
        v_full = np.zeros(len(pts) * bs)
        v_full[k::bs] = phi
        ve = v_full[full_edofs]
        if micro.kind == "poisson":
            Fv = np.einsum("eap,ea->ep", B_plain, ve)
        else:
            Fv = np.einsum("eaij,ea->eij", B_plain, ve)
        corr = solver.solve(micro.load(Abar, B, Fv))
        v_fields.append(Fv)
        totals.append(Fv + micro.field_of(B, corr))
    S = np.empty((nb, nb))
    for i in range(nb):
        for j in range(nb):
            S[i, j] = micro.energy(Abar, totals[i], totals[j]) / eps**2  # hmm.py:659-667
    return S * simplex_volume(verts) / micro.Y  # hmm.py:367-369


# ----------------------------------------------------------------------------
# macro assembly and solve (hmm.py:298-332, 434-491)
# ----------------------------------------------------------------------------
def unroll_dofs(dofs, bs):
    dofs = np.asarray(dofs)
    if bs == 1:
        return dofs
    return (dofs[:, None] * bs + np.arange(bs)[None, :]).ravel()


def assemble_macro(macro, micro, A, eps, Dtheta_t=None, literal=False, cells=None):
    """Global stiffness matrix (scipy CSR) from per-cell local matrices."""
    d, bs = micro.d, micro.bs
    n = len(macro.x) * bs
    rows, cols, vals = [], [], []
    cell_ids = range(len(macro.cells)) if cells is None else cells
    for c in cell_ids:
        nodes = macro.cells[c]
        verts = macro.x[nodes]
        if literal:
            S = local_stiffness_literal(micro, A, verts, eps, Dtheta_t)
        else:
            c_t = verts.mean(axis=0)
            M = None if Dtheta_t is None else np.asarray(Dtheta_t(c_t), float).reshape(d, d)
            S = local_stiffness_from_tensor(cell_tensor(micro, A, c_t, M), verts, micro.kind)
        dofs = unroll_dofs(nodes, bs)
        rows.append(np.repeat(dofs, len(dofs)))
        cols.append(np.tile(dofs, len(dofs)))
        vals.append(S.ravel())
    Amat = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n))
    return Amat.tocsr()


def assemble_rhs(macro, f, bs, degree=2):
    """b_i = int f . phi_i (hmm.py:131-133,445-450); f(x) -> scalar or (bs,) values."""
    d = macro.dim
    qp, qw = simplex_rule(d, degree)
    b = np.zeros(len(macro.x) * bs)
    X = macro.x[:, :d]
    v = X[macro.cells]
    J = np.transpose(v[:, 1:] - v[:, :1], (0, 2, 1))
    detJ = np.abs(np.linalg.det(J))
    phi = np.concatenate([1.0 - qp.sum(axis=1, keepdims=True), qp], axis=1)  # (nq, d+1)
    xq = v[:, :1, :] + np.einsum("eij,qj->eqi", J, qp)
    xx = np.zeros((3, xq.shape[0] * xq.shape[1]))
    xx[:d] = xq.reshape(-1, d).T
    fv = np.asarray(f(xx), float)
    fv = np.broadcast_to(fv.reshape(bs, -1) if fv.ndim else fv, (bs, xx.shape[1])).reshape(bs, *xq.shape[:2])
    contrib = np.einsum("keq,q,qa,e->eak", fv, qw, phi, detJ)  # (ne, d+1, bs)
    dofs = (macro.cells[:, :, None] * bs + np.arange(bs)[None, None, :])
    np.add.at(b, dofs.ravel(), contrib.ravel())
    return b


def solve_dirichlet(Amat, b, bc_dofs, bc_vals):
    """Symmetric lifting as in hmm.py:453-480 (A u_bc subtracted, rows/cols zeroed, unit diagonal)."""
    Amat = Amat.tocsr().copy()
    b = b.copy()
    bc_dofs = np.asarray(bc_dofs, dtype=np.int64)
    u_bc = np.zeros(Amat.shape[0])
    u_bc[bc_dofs] = bc_vals
    b -= Amat @ u_bc
    keep = np.ones(Amat.shape[0])
    keep[bc_dofs] = 0.0
    Dk = sp.diags(keep)
    Amat = Dk @ Amat @ Dk + sp.diags(1.0 - keep)
    b[bc_dofs] = u_bc[bc_dofs]
    return spla.spsolve(Amat.tocsc(), b)


def l2_error_squared(macro, u, exact, degree=6):
    """int (u_h - exact)^2 dx for scalar P1 u_h (test_integration_poisson.py:140-143)."""
    d = macro.dim
    qp, qw = simplex_rule(d, degree)
    X = macro.x[:, :d]
    v = X[macro.cells]
    J = np.transpose(v[:, 1:] - v[:, :1], (0, 2, 1))
    detJ = np.abs(np.linalg.det(J))
    phi = np.concatenate([1.0 - qp.sum(axis=1, keepdims=True), qp], axis=1)
    xq = v[:, :1, :] + np.einsum("eij,qj->eqi", J, qp)
    uh = np.einsum("ea,qa->eq", u[macro.cells], phi)
    xx = np.zeros((3, xq.shape[0] * xq.shape[1]))
    xx[:d] = xq.reshape(-1, d).T
    ue = np.asarray(exact(xx)).reshape(xq.shape[:2])
    return float(np.einsum("eq,q,e->", (uh - ue) ** 2, qw, detJ))


def boundary_nodes(macro, predicate=None):
    """Vertices on the bounding box boundary (PoissonHMM default BC, hmm.py:598-636),
    optionally filtered by ``predicate(x)`` with x of shape (3, N)."""
    x = macro.x
    d = macro.dim
    lo, hi = x.min(axis=0), x.max(axis=0)
    on = np.zeros(len(x), bool)
    for k in range(d):
        on |= np.isclose(x[:, k], lo[k]) | np.isclose(x[:, k], hi[k])
    if predicate is not None:
        on &= np.asarray(predicate(x.T), bool)
    return np.nonzero(on)[0]
