// Kernel instantiations: LinearStateSpaceModel with 3 states on the fixed-dt finite-difference grid, the time-optimal
// non-uniform grid and the multiple-shooting grid (RK4).
#include "lm_kernels.cuh"

namespace b200sqp {

const KernelSet* kernelTableLinear(int* count)
{
    static const KernelSet table[] = {
        B200SQP_KERNEL_ENTRY(LinearStateSpace3x1, DEFECT_CRANK_NICOLSON, 0, 4),
        B200SQP_KERNEL_ENTRY(LinearStateSpace3x1, DEFECT_CRANK_NICOLSON, 1, 4),
        B200SQP_KERNEL_ENTRY(LinearStateSpace3x1, DEFECT_RK4, 0, 4),
    };
    *count = (int)(sizeof(table) / sizeof(table[0]));
    return table;
}

}  // namespace b200sqp
