// Kernel instantiations: Van der Pol with the other collocation rules and the variable-dt grid.
#include "lm_kernels.cuh"

namespace b200sqp {

const KernelSet* kernelTableVdpFd(int* count)
{
    static const KernelSet table[] = {
        B200SQP_KERNEL_ENTRY(VanDerPol, DEFECT_FORWARD, 0, 4),
        B200SQP_KERNEL_ENTRY(VanDerPol, DEFECT_BACKWARD, 0, 4),
        B200SQP_KERNEL_ENTRY(VanDerPol, DEFECT_MIDPOINT, 0, 4),
        B200SQP_KERNEL_ENTRY(VanDerPol, DEFECT_CRANK_NICOLSON, 1, 4),
    };
    *count = (int)(sizeof(table) / sizeof(table[0]));
    return table;
}

}  // namespace b200sqp
