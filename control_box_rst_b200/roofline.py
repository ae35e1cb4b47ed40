"""Algorithmic work of one LM outer iteration of one OCP instance: the bytes come from the library
(`b200sqp_dims.algorithmic_bytes_per_iteration`), the floating-point operations are counted here from the structure.

One outer iteration = one linearisation (values + central-difference Jacobians of every edge + J^T J, J^T(-r)), one factorisation and
solve of the block-tridiagonal normal equations, one trial-point evaluation.  Every +, -, *, / and every sin / cos call counts as ONE
operation (so the figure is a lower bound for the trigonometric models); an FMA the kernel issues counts as two.  Rejected steps
(extra factorisations and trial points) are not counted -- like the algorithmic bytes, the figure describes the work the reference's
data flow asks for, not what a particular kernel executes.
"""
from . import _abi as abi

# operations of one dynamics evaluation f(x, u)  (control_box_rst_b200/csrc/dynamics.cuh, one count per arithmetic operator / libm call)
DYNAMICS_FLOPS = {
    abi.DYN_VAN_DER_POL: 7, abi.DYN_DUFFING: 9, abi.DYN_SIMPLE_PENDULUM: 9, abi.DYN_CART_POLE: 29, abi.DYN_DOUBLE_INTEGRATOR: 1,
    abi.DYN_UNICYCLE: 4, abi.DYN_QUADROTOR: 45, abi.DYN_FREE_SPACE_ROCKET: 7, abi.DYN_MASSLESS_PENDULUM: 3, abi.DYN_TOY_EXAMPLE: 11,
    abi.DYN_ARTSTEINS_CIRCLE: 7, abi.DYN_LINEAR_2X1: 10, abi.DYN_LINEAR_3X1: 21, abi.DYN_LINEAR_4X1: 36, abi.DYN_LINEAR_4X2: 44,
    abi.DYN_TRIPLE_INTEGRATOR: 1, abi.DYN_QUAD_INTEGRATOR: 1,
}


def defect_flops(ocp):
    """operations of one defect evaluation e(x_k, u_k, x_{k+1}, dt) (dynamics.cuh defect<>)"""
    nx, f = ocp.nx, DYNAMICS_FLOPS[ocp.dynamics]
    if ocp.grid == abi.GRID_MULTIPLE_SHOOTING:
        if ocp.integrator == abi.INT_EULER:
            return f + 3 * nx
        return 4 * f + 19 * nx  # RK4: 4 stage scalings, 3 stage points, the weighted sum, / 6, + x1, - x2
    if ocp.collocation == abi.COLL_CRANK_NICOLSON:
        return 2 * f + 5 * nx
    if ocp.collocation == abi.COLL_MIDPOINT:
        return f + 5 * nx
    return f + 3 * nx


def flops_per_iteration(ocp, dims):
    nx, nu = ocp.nx, ocp.nu
    K = ocp.n_grid - 1
    vt = 1 if ocp.grid == abi.GRID_FD_NONUNIFORM_VARDT else 0
    nb = nu + vt + nx
    e = defect_flops(ocp)
    cols = 2 * nx + nu + vt                       # central-difference columns of one dynamics edge
    lin = (1 + 2 * cols) * e + 3 * nx * cols      # defect evaluations + (e2 - e1) * scalar * w per entry
    lin += 8 * (nx + nu + 2 * vt) + 3 * (nu + vt)  # diagonal lsq cost edges (value + 2 perturbed values + difference) and bound rows
    gram = 2 * nx * (nb * (nb + 1) // 2 + nb * nx + nx * (nx + 1) // 2 + nb + nx)  # G^T G, G^T A, A^T A, J^T(-r): inner dimension nx
    factor = nb ** 3 // 3 + nb * nx * nx + nx * nb * (nb + 1) + 2 * (2 * nb * nb + 2 * nb * nx)  # chol, W solve, Schur, substitutions
    trial = e + 2 * (nx + nu + vt) + nb
    return K * (lin + gram + factor + trial)
