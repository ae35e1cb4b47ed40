// Kernel instantiations: SerialIntegratorSystem of dimension 3 (also time-optimal) and 4.
#include "lm_kernels.cuh"

namespace b200sqp {

const KernelSet* kernelTableIntegrators(int* count)
{
    static const KernelSet table[] = {
        B200SQP_KERNEL_ENTRY(TripleIntegrator, DEFECT_CRANK_NICOLSON, 0, 4),
        B200SQP_KERNEL_ENTRY(TripleIntegrator, DEFECT_CRANK_NICOLSON, 1, 4),
        B200SQP_KERNEL_ENTRY(QuadIntegrator, DEFECT_CRANK_NICOLSON, 0, 4),
    };
    *count = (int)(sizeof(table) / sizeof(table[0]));
    return table;
}

}  // namespace b200sqp
