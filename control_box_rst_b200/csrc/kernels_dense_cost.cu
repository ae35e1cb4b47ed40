// Kernel instantiations: the general feature set + full (non-diagonal) cost weight matrices (quadratic_cost.cpp:32-96: the lsq residual is
// U (x - xref) with the upper Cholesky factor U of Q), for a selection of (model, defect, grid) combinations.  A structure with full
// weights on any other combination answers B200SQP_ERR_UNSUPPORTED -- one line here adds it.
#include "lm_kernels.cuh"

namespace b200sqp {

const KernelSet* kernelTableDenseCost(int* count)
{
    static const KernelSet table[] = {
        B200SQP_KERNEL_ENTRY_DENSE(VanDerPol, DEFECT_CRANK_NICOLSON, 0, 4),
        B200SQP_KERNEL_ENTRY_DENSE(VanDerPol, DEFECT_RK4, 0, 2),
        B200SQP_KERNEL_ENTRY_DENSE(Duffing, DEFECT_CRANK_NICOLSON, 0, 2),
        B200SQP_KERNEL_ENTRY_DENSE(FreeSpaceRocket, DEFECT_CRANK_NICOLSON, 0, 2),
        B200SQP_KERNEL_ENTRY_DENSE(Unicycle, DEFECT_CRANK_NICOLSON, 0, 2),
        B200SQP_KERNEL_ENTRY_DENSE(LinearStateSpace4x2, DEFECT_CRANK_NICOLSON, 0, 2),
    };
    *count = (int)(sizeof(table) / sizeof(table[0]));
    return table;
}

}  // namespace b200sqp
