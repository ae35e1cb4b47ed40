// Kernel instantiations: the 2-state / 1-input benchmark systems (Van der Pol, Duffing, pendulum, double integrator).
#include "lm_kernels.cuh"

namespace b200sqp {

const KernelSet* kernelTableOscillators(int* count)
{
    static const KernelSet table[] = {
        // FiniteDifferencesGrid, all four collocation rules (configs[0], configs[1])
        B200SQP_KERNEL_ENTRY(VanDerPol, DEFECT_FORWARD, 0),
        B200SQP_KERNEL_ENTRY(VanDerPol, DEFECT_BACKWARD, 0),
        B200SQP_KERNEL_ENTRY(VanDerPol, DEFECT_MIDPOINT, 0),
        B200SQP_KERNEL_ENTRY(VanDerPol, DEFECT_CRANK_NICOLSON, 0),
        // NonUniformFiniteDifferencesVariableGrid (time-optimal stand-in used by the survey)
        B200SQP_KERNEL_ENTRY(VanDerPol, DEFECT_CRANK_NICOLSON, 1),
        // MultipleShootingGrid
        B200SQP_KERNEL_ENTRY(VanDerPol, DEFECT_EULER, 0),
        B200SQP_KERNEL_ENTRY(VanDerPol, DEFECT_RK4, 0),
        B200SQP_KERNEL_ENTRY(Duffing, DEFECT_CRANK_NICOLSON, 0),
        B200SQP_KERNEL_ENTRY(SimplePendulum, DEFECT_CRANK_NICOLSON, 0),
        B200SQP_KERNEL_ENTRY(DoubleIntegrator, DEFECT_CRANK_NICOLSON, 0),
        B200SQP_KERNEL_ENTRY(DoubleIntegrator, DEFECT_FORWARD, 0),
    };
    *count = (int)(sizeof(table) / sizeof(table[0]));
    return table;
}

}  // namespace b200sqp
