// Kernel instantiations: unicycle (configs[2]: NonUniformFiniteDifferencesVariableGrid, time-optimal).
#include "lm_kernels.cuh"

namespace b200sqp {

const KernelSet* kernelTableUnicycle(int* count)
{
    static const KernelSet table[] = {
        B200SQP_KERNEL_ENTRY(Unicycle, DEFECT_CRANK_NICOLSON, 1, 8),
        B200SQP_KERNEL_ENTRY(Unicycle, DEFECT_CRANK_NICOLSON, 0, 4),
    };
    *count = (int)(sizeof(table) / sizeof(table[0]));
    return table;
}

}  // namespace b200sqp
