// Kernel instantiations: the warp-cooperative LM pipeline (lm_pipeline.cuh) of the 12-state quadrotor in fp64 and fp32.
#include "lm_pipeline.cuh"

namespace b200sqp {

bool launchPipelineQuadrotorCn(const DeviceOcp& P, const DeviceState& st, const PipeArraysT<double>& pa, int iterations, cudaStream_t stream)
{
    return launchPipelineT<Quadrotor, DEFECT_CRANK_NICOLSON, double>(P, st, pa, iterations, stream);
}
bool launchPipelineQuadrotorCnF32(const DeviceOcp& P, const DeviceState& st, const PipeArraysT<float>& pa, int iterations, cudaStream_t stream)
{
    return launchPipelineT<Quadrotor, DEFECT_CRANK_NICOLSON, float>(P, st, pa, iterations, stream);
}

}  // namespace b200sqp
