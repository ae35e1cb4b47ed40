// Kernel instantiations: 12-state quadrotor (configs[4]: FiniteDifferencesGrid).
#include "lm_kernels.cuh"
#include "lm_pipeline.cuh"

namespace b200sqp {

const KernelSet* kernelTableQuadrotor(int* count)
{
    static const KernelSet table[] = {
        B200SQP_KERNEL_ENTRY_PIPELINE(Quadrotor, DEFECT_CRANK_NICOLSON, 1),
    };
    *count = (int)(sizeof(table) / sizeof(table[0]));
    return table;
}

}  // namespace b200sqp
