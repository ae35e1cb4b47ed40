"""B200-native Levenberg-Marquardt / SQP inner loop for control_box_rst's hypergraph-structured optimal control problems.

Host-side mirror of the reference's solver interface over the C ABI of libb200sqp.so (include/b200sqp.h):

    from control_box_rst_b200 import problems, solver
    lm = solver.BatchedLevenbergMarquardt(problems.van_der_pol(50), batch=4096)   # corbo::LevenbergMarquardtSparse's setters
    lm.setIterations(10)
    u0, chi2, status = lm.mpc_step(x0)                                            # measured states in, first controls out

There is no CPU implementation behind this package: without the CUDA library or a B200-class device construction raises.
"""
__all__ = ["problems", "solver", "distributed"]
