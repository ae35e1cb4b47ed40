// Kernel instantiations: Van der Pol on MultipleShootingGrid (Euler / RK4).
#include "lm_kernels.cuh"

namespace b200sqp {

const KernelSet* kernelTableVdpMs(int* count)
{
    static const KernelSet table[] = {
        B200SQP_KERNEL_ENTRY(VanDerPol, DEFECT_EULER, 0, 4),
        B200SQP_KERNEL_ENTRY(VanDerPol, DEFECT_RK4, 0, 4),
    };
    *count = (int)(sizeof(table) / sizeof(table[0]));
    return table;
}

}  // namespace b200sqp
