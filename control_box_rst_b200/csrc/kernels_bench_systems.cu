// Kernel instantiations: the remaining systems of the reference's nonlinear benchmark header (rocket, massless pendulum, toy example,
// Artstein's circle) on the fixed-dt FiniteDifferencesGrid, the rocket also on the time-optimal non-uniform grid; the 2-state
// LinearStateSpaceModel on all three grids.
#include "lm_kernels.cuh"

namespace b200sqp {

const KernelSet* kernelTableBenchmarkSystems(int* count)
{
    static const KernelSet table[] = {
        B200SQP_KERNEL_ENTRY(FreeSpaceRocket, DEFECT_CRANK_NICOLSON, 0, 4),
        B200SQP_KERNEL_ENTRY(FreeSpaceRocket, DEFECT_CRANK_NICOLSON, 1, 4),
        B200SQP_KERNEL_ENTRY(MasslessPendulum, DEFECT_CRANK_NICOLSON, 0, 4),
        B200SQP_KERNEL_ENTRY(ToyExample, DEFECT_CRANK_NICOLSON, 0, 4),
        B200SQP_KERNEL_ENTRY(ArtsteinsCircle, DEFECT_CRANK_NICOLSON, 0, 4),
        B200SQP_KERNEL_ENTRY(LinearStateSpace2x1, DEFECT_CRANK_NICOLSON, 0, 4),
        B200SQP_KERNEL_ENTRY(LinearStateSpace2x1, DEFECT_CRANK_NICOLSON, 1, 4),
        B200SQP_KERNEL_ENTRY(LinearStateSpace2x1, DEFECT_RK4, 0, 4),
    };
    *count = (int)(sizeof(table) / sizeof(table[0]));
    return table;
}

}  // namespace b200sqp
