// Kernel instantiations: further (dynamics, integrator) combinations on the MultipleShootingGrid.
// A combination is one line here plus a reference-pinned fixture in tests/golden/cases.py.
#include "lm_kernels.cuh"

namespace b200sqp {

const KernelSet* kernelTableCombosMs(int* count)
{
    static const KernelSet table[] = {
        B200SQP_KERNEL_ENTRY(CartPole, DEFECT_EULER, 0, 4),
        B200SQP_KERNEL_ENTRY(Unicycle, DEFECT_RK4, 0, 4),
        B200SQP_KERNEL_ENTRY(Duffing, DEFECT_RK4, 0, 4),
        B200SQP_KERNEL_ENTRY(SimplePendulum, DEFECT_RK4, 0, 4),
        B200SQP_KERNEL_ENTRY(DoubleIntegrator, DEFECT_RK4, 0, 4),
        B200SQP_KERNEL_ENTRY(DoubleIntegrator, DEFECT_EULER, 0, 4),
    };
    *count = (int)(sizeof(table) / sizeof(table[0]));
    return table;
}

}  // namespace b200sqp
