// Kernel instantiations: the other 2-state benchmark systems (Duffing, pendulum, double integrator).
#include "lm_kernels.cuh"

namespace b200sqp {

const KernelSet* kernelTableOscillators(int* count)
{
    static const KernelSet table[] = {
        B200SQP_KERNEL_ENTRY(Duffing, DEFECT_CRANK_NICOLSON, 0, 4),
        B200SQP_KERNEL_ENTRY(SimplePendulum, DEFECT_CRANK_NICOLSON, 0, 4),
        B200SQP_KERNEL_ENTRY(DoubleIntegrator, DEFECT_CRANK_NICOLSON, 0, 4),
        B200SQP_KERNEL_ENTRY(DoubleIntegrator, DEFECT_FORWARD, 0, 4),
    };
    *count = (int)(sizeof(table) / sizeof(table[0]));
    return table;
}

}  // namespace b200sqp
