// Kernel instantiations: NonUniformFiniteDifferencesVariableGrid with setDtEqConstraint(true) -- one TwoScalarEqualEdge per pair of
// consecutive intervals (edges/misc_edges.h:40-67) couples the dt slots of neighbouring stage blocks (VT = 2 in lm_device.cuh Dim).
#include "lm_kernels.cuh"

namespace b200sqp {

const KernelSet* kernelTableDtEquality(int* count)
{
    static const KernelSet table[] = {
        B200SQP_KERNEL_ENTRY(Unicycle, DEFECT_CRANK_NICOLSON, 2, 8),
        B200SQP_KERNEL_ENTRY(VanDerPol, DEFECT_CRANK_NICOLSON, 2, 4),
        B200SQP_KERNEL_ENTRY(FreeSpaceRocket, DEFECT_CRANK_NICOLSON, 2, 4),
        B200SQP_KERNEL_ENTRY(DoubleIntegrator, DEFECT_CRANK_NICOLSON, 2, 4),
    };
    *count = (int)(sizeof(table) / sizeof(table[0]));
    return table;
}

}  // namespace b200sqp
