// Kernel instantiations: further (dynamics, collocation, grid) combinations on the finite-difference grids -- time-optimal
// (NonUniformFiniteDifferencesVariableGrid) double integrator, pendulum and cart-pole; one non-Crank-Nicolson collocation each for the
// trigonometric models.  A combination is one line here plus a reference-pinned fixture in tests/golden/cases.py.
#include "lm_kernels.cuh"

namespace b200sqp {

const KernelSet* kernelTableCombosFd(int* count)
{
    static const KernelSet table[] = {
        B200SQP_KERNEL_ENTRY(DoubleIntegrator, DEFECT_CRANK_NICOLSON, 1, 4),
        B200SQP_KERNEL_ENTRY(SimplePendulum, DEFECT_CRANK_NICOLSON, 1, 4),
        B200SQP_KERNEL_ENTRY(CartPole, DEFECT_CRANK_NICOLSON, 1, 4),
        B200SQP_KERNEL_ENTRY(SimplePendulum, DEFECT_MIDPOINT, 0, 4),
        B200SQP_KERNEL_ENTRY(CartPole, DEFECT_FORWARD, 0, 4),
        B200SQP_KERNEL_ENTRY(Unicycle, DEFECT_BACKWARD, 0, 4),
    };
    *count = (int)(sizeof(table) / sizeof(table[0]));
    return table;
}

}  // namespace b200sqp
