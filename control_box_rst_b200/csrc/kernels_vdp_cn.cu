// Kernel instantiations: Van der Pol + Crank-Nicolson on FiniteDifferencesGrid (configs[0], configs[1]: the benchmark path).
#include "lm_kernels.cuh"

namespace b200sqp {

const KernelSet* kernelTableVdpCn(int* count)
{
    static const KernelSet table[] = {
        B200SQP_KERNEL_ENTRY(VanDerPol, DEFECT_CRANK_NICOLSON, 0, 8),
    };
    *count = (int)(sizeof(table) / sizeof(table[0]));
    return table;
}

}  // namespace b200sqp
