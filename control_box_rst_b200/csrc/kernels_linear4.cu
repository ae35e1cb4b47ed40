// Kernel instantiations: LinearStateSpaceModel with 4 states (one and two inputs).
#include "lm_kernels.cuh"

namespace b200sqp {

const KernelSet* kernelTableLinear4(int* count)
{
    static const KernelSet table[] = {
        B200SQP_KERNEL_ENTRY(LinearStateSpace4x1, DEFECT_CRANK_NICOLSON, 0, 4),
        B200SQP_KERNEL_ENTRY(LinearStateSpace4x1, DEFECT_RK4, 0, 4),
        B200SQP_KERNEL_ENTRY(LinearStateSpace4x2, DEFECT_CRANK_NICOLSON, 0, 4),
        B200SQP_KERNEL_ENTRY(LinearStateSpace4x2, DEFECT_CRANK_NICOLSON, 1, 4),
    };
    *count = (int)(sizeof(table) / sizeof(table[0]));
    return table;
}

}  // namespace b200sqp
