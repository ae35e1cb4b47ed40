// Kernel instantiations: 12-state quadrotor (configs[4]: FiniteDifferencesGrid), fused kernel; its warp-cooperative pipeline lives in
// kernels_quadrotor_pipe.cu (a translation unit of its own: the two compile in parallel and edit independently).
#include "lm_kernels.cuh"

namespace b200sqp {

bool launchPipelineQuadrotorCn(const DeviceOcp&, const DeviceState&, const PipeArraysT<double>&, int iterations, cudaStream_t);
bool launchPipelineQuadrotorCnF32(const DeviceOcp&, const DeviceState&, const PipeArraysT<float>&, int iterations, cudaStream_t);

const KernelSet* kernelTableQuadrotor(int* count)
{
    static const KernelSet table[] = {
        KernelSet{Quadrotor::ID, DEFECT_CRANK_NICOLSON, 0, Quadrotor::NX, Quadrotor::NU, 1, &launchSolve<Quadrotor, DEFECT_CRANK_NICOLSON, 0, 1>,
                  &launchEvaluate<Quadrotor, DEFECT_CRANK_NICOLSON, 0>, &launchPipelineQuadrotorCn, &launchPipelineQuadrotorCnF32},
    };
    *count = (int)(sizeof(table) / sizeof(table[0]));
    return table;
}

}  // namespace b200sqp
