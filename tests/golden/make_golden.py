"""Generates tests/golden/*.npz from the UNMODIFIED reference compiled out of /root/reference (oracle/_ref/libcorbo_ref.so).

Run where /root/reference exists:   python tests/golden/make_golden.py
Every array in the fixtures is an output of the reference's own classes (StructuredOptimalControlProblem, the discretization grids,
HyperGraphOptimizationProblemEdgeBased, LevenbergMarquardtSparse); nothing is computed by this repository's oracle or kernels.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from control_box_rst_b200 import _abi as abi  # noqa: E402
from control_box_rst_b200 import problems  # noqa: E402
from oracle import bindings  # noqa: E402
import cases  # noqa: E402


def main():
    bindings.build_reference()
    ref = bindings.Reference()
    only = set(sys.argv[1:])  # optional: regenerate just the named cases
    for name, (make, weights, B) in cases.CASES.items():
        if only and name not in only:
            continue
        ocp = make()
        d = ref.dims(ocp)
        x_idx, u_idx, dt_idx = ref.vertex_indices(ocp)
        x0, xref = problems.instance_data(ocp, B, seed=21)
        p_init = np.stack([ref.initial_params(ocp, x0[i], xref[i]) for i in range(B)])
        p_eval = cases.perturbed_params(ocp, p_init)
        if ocp.grid == abi.GRID_FD_NONUNIFORM_VARDT:
            p_eval[:, dt_idx] = np.abs(p_eval[:, dt_idx]) + 0.05
        # Jacobian-level fixture for instance 0
        values, J, pattern, after = ref.evaluate(ocp, x0[0], xref[0], p_eval[0], cases.EVAL_WEIGHTS)
        rows, cols = np.nonzero(pattern.T)  # column-major order = CSC
        col_idx, row_idx = rows.astype(np.int32), cols.astype(np.int32)
        col_ptr = np.zeros(d.n_params + 1, np.int32)
        np.add.at(col_ptr, col_idx + 1, 1)
        col_ptr = np.cumsum(col_ptr).astype(np.int32)
        jac_values = cases.dense_to_csc_values(J, col_ptr, row_idx)
        # solver-level fixture
        opts = abi.LmOptions.defaults(iterations=10, weights=weights)
        p_final, chi2, status, _ = ref.solve_batch(ocp, opts, x0, xref, threads=4)
        tr = ref.trace(ocp, opts, x0[0], xref[0])
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"),
            dims=np.array([d.n_params, d.m_lsq, d.m_eq, d.m_ineq, d.m_bounds, d.nnz_jacobian, d.nnz_hessian_upper,
                           d.algorithmic_bytes_per_iteration], np.int64),
            x_idx=x_idx, u_idx=u_idx, dt_idx=dt_idx,
            edges_lsq=ref.edge_table(ocp, 0), edges_eq=ref.edge_table(ocp, 1), edges_ineq=ref.edge_table(ocp, 2),
            col_ptr=col_ptr, row_idx=row_idx,
            x0=x0, xref=xref, p_init=p_init, p_eval=p_eval[0], values=values, jac_values=jac_values, p_after=after,
            p_final=p_final, chi2=chi2, status=status,
            trace_types=np.array([e[0] for e in tr["events"]], np.int32),
            trace_chi2=np.array([e[1] for e in tr["events"]]),
        )
        print(f"{name}: n={d.n_params} m={d.m} nnzJ={d.nnz_jacobian} chi2[0]={chi2[0]:.6g} events={tr['n_events']}")
    if not only or "linearize" in only:
        # SystemDynamicsInterface::getLinearA/getLinearB of the reference's own (and the two added) models, both FD rules
        out = {}
        for name, (make, _) in cases.LINEARIZE_MODELS.items():
            ocp = make()
            xs, us = cases.linearize_points(ocp)
            for method in ("forward", "central"):
                AB = [ref.linearize(ocp, xs[i], us[i], method) for i in range(len(xs))]
                out[f"{name}_{method}_A"] = np.stack([a for a, _ in AB])
                out[f"{name}_{method}_B"] = np.stack([b for _, b in AB])
            out[f"{name}_x"], out[f"{name}_u"] = xs, us
            # the second half of numerics/finite_differences: Forward/CentralDifferences::hessian (delta = 1e-5) of the same dynamics
            mult = cases.hessian_multipliers(ocp)
            for method in ("forward", "central"):
                out[f"{name}_{method}_H"] = np.stack([ref.dynamics_hessian(ocp, xs[i], us[i], None, method) for i in range(len(xs))])
                out[f"{name}_{method}_Hm"] = np.stack([ref.dynamics_hessian(ocp, xs[i], us[i], mult[i], method) for i in range(len(xs))])
        np.savez_compressed(os.path.join(HERE, "linearize.npz"), **out)
        print("linearize:", len(out), "arrays")
    if not only or "warm_start" in only:
        # moving-horizon warm start of the reference grid, isolated (corbo_ref_warm_start_shift) and inside the closed loop
        ocp = problems.van_der_pol(12)
        rng = np.random.default_rng(4)
        x0_old, x0_new, p_in, p_out = [], [], [], []
        for trial in range(12):
            xo = rng.uniform(-2, 2, 2)
            p = ref.initial_params(ocp, xo, None) + rng.uniform(-0.02, 0.02, ref.dims(ocp).n_params)
            x_idx, _, _ = ref.vertex_indices(ocp)
            s = trial % 5  # aim at the s-th state of the old trajectory (0 = start did not move at all for trial 0)
            target = xo if s == 0 else p[x_idx[s]:x_idx[s] + 2]
            xn = target.copy() if trial == 0 else target + rng.uniform(-0.01, 0.01, 2)
            x0_old.append(xo)
            x0_new.append(xn)
            p_in.append(p)
            p_out.append(ref.warm_start_shift(ocp, xo, xn, p))
        # the same on a MultipleShootingGrid (ShootingGridBase::findNearestShootingInterval + warmStartShifting)
        ms = problems.van_der_pol_shooting(12)
        ms_x0_old, ms_x0_new, ms_p_in, ms_p_out = [], [], [], []
        for trial in range(12):
            xo = rng.uniform(-2, 2, 2)
            p = ref.initial_params(ms, xo, None) + rng.uniform(-0.02, 0.02, ref.dims(ms).n_params)
            x_idx, _, _ = ref.vertex_indices(ms)
            s = trial % 5
            target = xo if s == 0 else p[x_idx[s]:x_idx[s] + 2]
            xn = target.copy() if trial == 0 else target + rng.uniform(-0.01, 0.01, 2)
            ms_x0_old.append(xo)
            ms_x0_new.append(xn)
            ms_p_in.append(p)
            ms_p_out.append(ref.warm_start_shift(ms, xo, xn, p))
        ocp20 = problems.van_der_pol(20)
        u, x = ref.closed_loop_shift(ocp20, abi.LmOptions.defaults(), np.array([1.0, 0.5]), 15)
        ms_u, ms_x = ref.closed_loop_plant(problems.van_der_pol_shooting(20), abi.LmOptions.defaults(), np.array([1.0, 0.5]), 15, "rk4", None, True)
        np.savez_compressed(os.path.join(HERE, "warm_start_shift.npz"), x0_old=np.array(x0_old), x0_new=np.array(x0_new), p_in=np.array(p_in),
                            p_out=np.array(p_out), loop_u=u, loop_x=x, ms_x0_old=np.array(ms_x0_old), ms_x0_new=np.array(ms_x0_new),
                            ms_p_in=np.array(ms_p_in), ms_p_out=np.array(ms_p_out), ms_loop_u=ms_u, ms_loop_x=ms_x)
        print("warm_start: closed loop with shifting u[:3] =", u[:3, 0])
    if not only or "plant" in only:
        # SimulatedPlant::control of the reference (both integrators) at the linearisation points, and ClosedLoopControlTask's loop
        # (PredictiveController + SimulatedPlant) for a few start states, with and without the moving-horizon warm start
        out = {}
        for name, (make, _) in cases.LINEARIZE_MODELS.items():
            ocp = make()
            xs, us = cases.linearize_points(ocp)
            for integ in ("euler", "rk4"):
                out[f"{name}_{integ}"] = ref.plant_step(ocp, xs, us, cases.PLANT_DT, integ)
        ocp20 = problems.van_der_pol(20)
        x0s = cases.closed_loop_starts()
        for integ, warm in (("euler", False), ("rk4", False), ("rk4", True)):
            ux = [ref.closed_loop_plant(ocp20, abi.LmOptions.defaults(), x0, cases.CLOSED_LOOP_STEPS, integ, None, warm) for x0 in x0s]
            tag = f"loop_{integ}_{'shift' if warm else 'keep'}"
            out[tag + "_u"] = np.stack([u for u, _ in ux], axis=1)  # [steps, B, nu]
            out[tag + "_x"] = np.stack([x for _, x in ux], axis=1)  # [steps+1, B, nx]
        out["loop_x0"] = x0s
        np.savez_compressed(os.path.join(HERE, "plant.npz"), **out)
        print("plant:", len(out), "arrays; closed loop (rk4, keep) u[:3, 0] =", out["loop_rk4_keep_u"][:3, 0, 0])
    if only:
        return
    ka = []
    for cid, stage in bindings.KNOWN_ANSWER_CASES:
        x, exp, tol = ref.known_answer(cid, stage)
        ka.append(np.concatenate([[cid, stage, len(x), tol], np.pad(x, (0, 3 - len(x))), np.pad(exp, (0, 3 - len(exp)))]))
    np.savez_compressed(os.path.join(HERE, "lm_known_answers.npz"), table=np.array(ka))
    # closed-loop plumbing (configs[0]): 15 MPC steps of the reference's own controller loop
    ocp = problems.van_der_pol(20)
    u, x = ref.closed_loop(ocp, abi.LmOptions.defaults(), np.array([1.0, 0.5]), 15)
    np.savez_compressed(os.path.join(HERE, "vdp20_closed_loop.npz"), u=u, x=x)
    print("closed loop u[:3] =", u[:3, 0])


if __name__ == "__main__":
    main()
