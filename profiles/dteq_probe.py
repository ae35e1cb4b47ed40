import sys, numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from control_box_rst_b200 import _abi as abi, problems, solver
from oracle import bindings
from test_oracle_vs_reference import DT_EQ_CASES
which = sys.argv[1]; T = int(sys.argv[2]); iters = int(sys.argv[3])
chk = bindings.Oracle()
ocp = DT_EQ_CASES[which]()
B = int(sys.argv[4]) if len(sys.argv) > 4 else 8
x0, xref = problems.instance_data(ocp, B, seed=4)
lm = solver.BatchedLevenbergMarquardt(ocp, B)
lm.set_problem_data(x0, xref); lm.initialize_trajectories()
rng = np.random.default_rng(2)
p = lm.get_params() + rng.uniform(-0.2, 0.2, (B, lm.dims.n_params))
dt_idx = solver.vertex_indices(ocp)[2]
p[:, dt_idx] = np.abs(p[:, dt_idx]) + 0.05
lm.set_params(p)
w = (2.0, 3.0, 4.0)
values, jac = lm.evaluate(w); after = lm.get_params()
v_c, J_c, _, a_c = chk.evaluate(ocp, x0[0], xref[0], p[0], w)
col_ptr, row_idx = solver.jacobian_pattern(ocp)
J = np.zeros_like(J_c)
for c in range(lm.dims.n_params):
    J[row_idx[col_ptr[c]:col_ptr[c+1]], c] = jac[0, col_ptr[c]:col_ptr[c+1]]
print("evaluate: values", np.abs(values[0]-v_c).max(), "J", np.abs(J-J_c).max(), "after", np.abs(after[0]-a_c).max(), flush=True)
if iters > 0:
    lm.setIterations(iters); lm.setPenaltyWeights(*w); lm.set_threads_per_instance(T)
    lm.initialize_trajectories()
    st, chi2 = lm.solve(new_run=True)
    opts = abi.LmOptions.defaults(iterations=iters, weights=w)
    p_c, c_c, _, _ = chk.solve_batch(ocp, opts, x0, xref, threads=2)
    print("solve T", T, "chi2 dev", chi2[:3], "oracle", c_c[:3], "traj err", np.abs(lm.get_params()-p_c).max(), lm.statistics()["inner_passes"][:4], flush=True)
