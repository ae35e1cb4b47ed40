// Kernel instantiations: cart-pole (configs[3]: MultipleShootingGrid + RK4).
#include "lm_kernels.cuh"

namespace b200sqp {

const KernelSet* kernelTableCartPole(int* count)
{
    static const KernelSet table[] = {
        B200SQP_KERNEL_ENTRY(CartPole, DEFECT_RK4, 0, 8),
        B200SQP_KERNEL_ENTRY(CartPole, DEFECT_CRANK_NICOLSON, 0, 4),
    };
    *count = (int)(sizeof(table) / sizeof(table[0]));
    return table;
}

}  // namespace b200sqp
