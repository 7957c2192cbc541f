"""The reference's example problems as workloads: BASELINE.json configs[0..3].

Coefficient factories take the ``ufl``-like namespace to build with (``hommx_b200.ufl`` for the CUDA path, real
``ufl`` under DOLFINx, ``oracle.npufl`` in the tests), so that one source text serves every back end; each cites
the reference example it restates (file:line relative to /root/reference).  ``WORKLOADS`` / ``build_solver`` are
what ``bench.py``, ``examples/`` and the strong-scaling runs share.
"""
from __future__ import annotations


def smooth_sin(ufl):
    """examples/hmm.py:15-16, examples/hmm_3d.py:14-15, test_integration_poisson.py:244-245,484-485"""

    def A(x, y):
        return 1.1 + x[0] + ufl.sin(2 * ufl.pi * y[0])

    return A


def laminate(ufl):
    """examples/diffusion/laminate.py:101-102"""

    def A(x, y):
        return ufl.conditional(ufl.cos(2 * ufl.pi * y[0]) < 0, 5, 0.05)

    return A


def circle_indicator(ufl, a, b, r=0.25):
    """examples/diffusion/inclusion.py:107-114, examples/linear_elasticity/rotated_fibers.py:23-29"""
    dx = ufl.acos(ufl.cos(2 * ufl.pi * (a - 1 / 2)))
    dy = ufl.acos(ufl.cos(2 * ufl.pi * (b - 1 / 2)))
    return (dx**2 + dy**2) < ((2 * ufl.pi) ** 2 * r**2)


def dtheta_wavy(ufl):
    """examples/diffusion/laminate.py:109-117 completed to a square matrix:
    theta(x) = (x1 - sin 2 pi x0, x0)."""

    def Dtheta(x):
        return ufl.as_matrix([[-2 * ufl.pi * ufl.cos(2 * ufl.pi * x[0]), 1.0], [1.0, 0.0]])

    return Dtheta


def dtheta_rotation_3d(ufl, W=0.4):
    """examples/linear_elasticity/rotated_fibers.py:41-63 completed to a square rotation about
    the x1 axis by gamma(x1) = pi x1 / (2 W)."""

    def Dtheta(x):
        g = 1 / 2 * ufl.pi * x[1] / W
        R = ufl.as_matrix(
            [[ufl.cos(g), 0.0, -ufl.sin(g)], [0.0, 1.0, 0.0], [ufl.sin(g), 0.0, ufl.cos(g)]]
        )
        return ufl.transpose(R)

    return Dtheta


def hooke(ufl, dim, mu, lambda_):
    """test/integration/test_integration_linear_elasticity.py:84-94,234-244;
    examples/linear_elasticity/rotated_fibers.py:66-76"""

    def A(x, y):
        I = ufl.Identity(dim)
        i, j, k, l = ufl.indices(4)
        return ufl.as_tensor(
            lambda_(x, y) * I[i, j] * I[k, l] + mu(x, y) * (I[i, k] * I[j, l] + I[i, l] * I[j, k]),
            indices=(i, j, k, l),
        )

    return A


def hooke_fibre_3d(ufl, mu_in=100, mu_out=0.001):
    """examples/linear_elasticity/rotated_fibers.py:23-38"""
    return hooke(
        ufl, 3, lambda x, y: ufl.conditional(circle_indicator(ufl, y[1], y[2]), mu_in, mu_out), lambda x, y: 1
    )


def hooke_spheres_3d(ufl, mu_in=100, mu_out=0.001):
    """Not in the reference: the fibre workload's contrast with an inclusion that varies along ALL three micro axes
    (a periodic ball of radius 0.3), so that no axis can be collapsed and no coarse axis is privileged -- reported
    next to C4 as the generic-coefficient figure."""

    def mu(x, y):
        d = [ufl.acos(ufl.cos(2 * ufl.pi * (y[k] - 1 / 2))) for k in range(3)]
        return ufl.conditional((d[0] ** 2 + d[1] ** 2 + d[2] ** 2) < ((2 * ufl.pi) ** 2 * 0.3**2), mu_in, mu_out)

    return hooke(ufl, 3, mu, lambda x, y: 1)


# name: class, dim, micro n, coefficient, Dtheta, per-GPU macro mesh (slab count multiplies the last axis)
WORKLOADS = {
    "c4": dict(cls="LinearElasticityStratifiedHMM", dim=3, kind=1, n=8, coeff="hooke_fibre_3d", dtheta="dtheta_rotation_3d",
               box=((0.0, 0.0, 0.0), (1.0, 0.4, 0.1)), cells=(40, 16, 4), strong_cells=(80, 32, 8), eps=0.01,
               desc="BASELINE configs[3]: rotated-fibre beam, Hooke mu=100/0.001 lambda=1, 8^3 micro cell, 6 RHS/point"),
    "c3": dict(cls="PoissonHMM", dim=3, kind=0, n=8, coeff="smooth_sin", dtheta=None,
               box=((0.0, 0.0, 0.0), (1.0, 1.0, 1.0)), cells=(32, 32, 32), strong_cells=(32, 32, 32), eps=2.0**-3,
               desc="BASELINE configs[2]: PoissonHMM 3D, 32^3 macro mesh per GPU, 8^3 micro cell"),
    "c2": dict(cls="PoissonStratifiedHMM", dim=2, kind=0, n=32, coeff="laminate", dtheta="dtheta_wavy",
               box=((0.0, 0.0), (1.0, 1.0)), cells=(256, 256), strong_cells=(256, 256), eps=1e-5,
               desc="BASELINE configs[1]: PoissonStratifiedHMM wavy laminate, 256x256 macro mesh per GPU, 32x32 micro cell"),
    "c1": dict(cls="PoissonHMM", dim=2, kind=0, n=16, coeff="smooth_sin", dtheta=None,
               box=((0.0, 0.0), (1.0, 1.0)), cells=(32, 32), strong_cells=(32, 32), eps=2.0**-5,
               desc="BASELINE configs[0]: PoissonHMM 2D, 32x32 macro mesh, 16x16 micro cell"),
    "c4n10": dict(cls="LinearElasticityStratifiedHMM", dim=3, kind=1, n=10, coeff="hooke_fibre_3d", dtheta="dtheta_rotation_3d",
                  box=((0.0, 0.0, 0.0), (1.0, 0.4, 0.1)), cells=(20, 8, 2), strong_cells=(20, 8, 2), eps=0.01,
                  desc="C4's beam and coefficient on a 10^3 micro cell (its vectors exceed one SM; not a BASELINE config): "
                       "cluster-resident stencil against the matrix-free kernel with its vectors in L2"),
    "c4s": dict(cls="LinearElasticityStratifiedHMM", dim=3, kind=1, n=8, coeff="hooke_spheres_3d", dtheta="dtheta_rotation_3d",
                box=((0.0, 0.0, 0.0), (1.0, 0.4, 0.1)), cells=(40, 16, 4), strong_cells=(80, 32, 8), eps=0.01,
                desc="C4's beam with a periodic stiff BALL instead of the fibre (coefficient varies along all three "
                     "micro axes; not a BASELINE config): 8^3 micro cell, 6 RHS/point"),
}  # fmt: skip


def coefficient(wl, ufl):
    w = WORKLOADS[wl]
    A = globals()[w["coeff"]](ufl)
    Dt = globals()[w["dtheta"]](ufl) if w["dtheta"] else None
    return A, Dt


def macro_cells(wl, world=1, shrink=1, scaling="weak"):
    """Macro mesh size: weak scaling gives every GPU one slab of ``cells`` (the last axis grows with ``world``);
    strong scaling keeps the named size ``strong_cells`` whatever the number of GPUs."""
    w = WORKLOADS[wl]
    cells = [max(1, c // shrink) for c in (w["strong_cells"] if scaling == "strong" else w["cells"])]
    if scaling != "strong":
        cells[-1] *= world
    return cells


def build_solver(wl, world=1, collapse=False, shrink=1, scaling="weak", rtol=1e-8, atol=1e-10, **kw):
    """The drop-in class of workload ``wl`` on its macro mesh (``hommx_b200`` stand-in meshes)."""
    import hommx_b200 as hx
    from hommx_b200 import mesh
    from hommx_b200 import ufl as pufl

    w = WORKLOADS[wl]
    cells = macro_cells(wl, world, shrink, scaling)
    msh = mesh.create_rectangle(*w["box"], cells) if w["dim"] == 2 else mesh.create_box(*w["box"], cells)
    mic = mesh.create_unit_square(w["n"], w["n"]) if w["dim"] == 2 else mesh.create_unit_cube(w["n"], w["n"], w["n"])
    A, Dt = coefficient(wl, pufl)
    f = (lambda x: 1.0) if w["kind"] == 0 else (lambda x: pufl.as_vector([0.0] * (w["dim"] - 1) + [-0.05 * 0.4**2]))
    opts = {"ksp_rtol": rtol, "ksp_atol": atol}
    cls = getattr(hx, w["cls"])
    if Dt is not None:
        return cls(msh, A, f, mic, w["eps"], Dt, petsc_options_cell_problem=opts, collapse_invariant_axes=collapse, **kw)
    return cls(msh, A, f, mic, w["eps"], petsc_options_cell_problem=opts, collapse_invariant_axes=collapse, **kw)


def kernel_jobs():
    """Cell kernels the workloads need (compiled ahead of time by __graft_entry__.build())."""
    from hommx_b200 import codegen
    from hommx_b200 import ufl as pufl

    jobs = []
    for name, w in WORKLOADS.items():
        A, Dt = coefficient(name, pufl)
        prog = codegen.build_program(A, w["dim"], w["kind"], Dt)
        jobs.append((prog, w["n"], None))
        if w["kind"] == 1 and w["dim"] == 3:  # full cell: both PCG kernels (bench: "*_cluster" / "*_pcg" entries)
            from hommx_b200 import native

            jobs.append((prog, w["n"], None, False, True, None, native.MATRIX_FREE, False))
            if native.cluster_size(prog, w["n"]) >= 2:
                jobs.append((prog, w["n"], None, False, True, None, native.CLUSTER, False))
        if prog.ydep != (1 << w["dim"]) - 1:
            jobs.append((prog, w["n"], None, False, True, None, None, True))  # axis-collapsed variant (other_workloads)
    return jobs
