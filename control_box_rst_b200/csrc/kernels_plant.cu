// Batched plant simulation: x_i <- solveIVP(x_i, u_i, dt) for a batch of independent plants -- the device counterpart of
//   SimulatedPlant::control                        src/plants/src/simulated_plant.cpp:92-146 (no dead time, no disturbances:
//                                                  one _integrator->solveIVP(_current_state, u, dt, *_dynamics, next_state), :123)
// with the reference's explicit integrators
//   IntegratorExplicitEuler::solveIVP              src/numerics/include/corbo-numerics/explicit_integrators.h:66-72  (the plant's default,
//                                                  simulated_plant.cpp:37)
//   IntegratorExplicitRungeKutta4::solveIVP        src/numerics/include/corbo-numerics/explicit_integrators.h:280-295
// (plant in the loop for Monte-Carlo closed-loop studies, BenchmarkTaskVaryingInitialState).
// The integrators are the ones the shooting defect already restates (dynamics.cuh, DEFECT_EULER / DEFECT_RK4 =
// solveIVP(x1, u1, dt) - x2): with x2 = 0 the subtraction is exact, so the plant step is that code with the same expression order
// and no FMA contraction (this TU is built with --fmad=false) -- bit-identical to the reference for the polynomial models.
// One thread per plant; states and controls are instance-major [B][nx] / [B][nu] exactly as the closed-loop log keeps them.  The
// kernel also writes the applied control into the log.  A step is microseconds (a few hundred flops per plant) next to the MPC
// solve it follows: a helper on the closed loop, not a tuned kernel.
#include "dynamics.cuh"
#include "launch.h"

namespace b200sqp {

namespace {

template <class M>
__global__ void plantStepKernel(const DynParams dyn, int integrator, double dt, int B, const double* __restrict__ xs, const double* __restrict__ us,
                                double* __restrict__ xn, double* __restrict__ u_log, const double* __restrict__ chi2_src, double* __restrict__ chi2_log,
                                const int* __restrict__ status_src, int* __restrict__ status_log)
{
    constexpr int NX = M::NX, NU = M::NU;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    double x[NX], u[NU], zero[NX], next[NX];
#pragma unroll
    for (int j = 0; j < NX; ++j)
    {
        x[j]    = xs[(size_t)i * NX + j];
        zero[j] = 0.0;
    }
#pragma unroll
    for (int j = 0; j < NU; ++j) u[j] = us[(size_t)i * NU + j];
    const StepSize h(dt);
    if (integrator == 0)
        defectCall<M, DEFECT_EULER>(dyn, x, u, zero, h, next);
    else
        defectCall<M, DEFECT_RK4>(dyn, x, u, zero, h, next);
#pragma unroll
    for (int j = 0; j < NX; ++j) xn[(size_t)i * NX + j] = next[j];
    if (u_log)
    {
#pragma unroll
        for (int j = 0; j < NU; ++j) u_log[(size_t)i * NU + j] = u[j];
    }
    // closed-loop log of the controller's chi2 / status of this step (saves two device-to-device copies per step)
    if (chi2_log) chi2_log[i] = chi2_src[i];
    if (status_log) status_log[i] = status_src[i];
}

template <class M>
void launchOne(const DynParams& dyn, int integrator, double dt, int B, const double* x, const double* u, double* xn, double* u_log,
               const double* chi2_src, double* chi2_log, const int* status_src, int* status_log, cudaStream_t st)
{
    plantStepKernel<M><<<(B + 127) / 128, 128, 0, st>>>(dyn, integrator, dt, B, x, u, xn, u_log, chi2_src, chi2_log, status_src, status_log);
}

}  // namespace

bool launchPlantStep(int dynamics, const DynParams& dyn, int integrator, double dt, int B, const double* x, const double* u, double* x_next,
                     double* u_log, const double* chi2_src, double* chi2_log, const int* status_src, int* status_log, cudaStream_t st)
{
    switch (dynamics)
    {
        case B200SQP_DYN_VAN_DER_POL: launchOne<VanDerPol>(dyn, integrator, dt, B, x, u, x_next, u_log, chi2_src, chi2_log, status_src, status_log, st); return true;
        case B200SQP_DYN_DUFFING: launchOne<Duffing>(dyn, integrator, dt, B, x, u, x_next, u_log, chi2_src, chi2_log, status_src, status_log, st); return true;
        case B200SQP_DYN_SIMPLE_PENDULUM: launchOne<SimplePendulum>(dyn, integrator, dt, B, x, u, x_next, u_log, chi2_src, chi2_log, status_src, status_log, st); return true;
        case B200SQP_DYN_CART_POLE: launchOne<CartPole>(dyn, integrator, dt, B, x, u, x_next, u_log, chi2_src, chi2_log, status_src, status_log, st); return true;
        case B200SQP_DYN_DOUBLE_INTEGRATOR: launchOne<DoubleIntegrator>(dyn, integrator, dt, B, x, u, x_next, u_log, chi2_src, chi2_log, status_src, status_log, st); return true;
        case B200SQP_DYN_UNICYCLE: launchOne<Unicycle>(dyn, integrator, dt, B, x, u, x_next, u_log, chi2_src, chi2_log, status_src, status_log, st); return true;
        case B200SQP_DYN_QUADROTOR: launchOne<Quadrotor>(dyn, integrator, dt, B, x, u, x_next, u_log, chi2_src, chi2_log, status_src, status_log, st); return true;
        case B200SQP_DYN_FREE_SPACE_ROCKET: launchOne<FreeSpaceRocket>(dyn, integrator, dt, B, x, u, x_next, u_log, chi2_src, chi2_log, status_src, status_log, st); return true;
        case B200SQP_DYN_MASSLESS_PENDULUM: launchOne<MasslessPendulum>(dyn, integrator, dt, B, x, u, x_next, u_log, chi2_src, chi2_log, status_src, status_log, st); return true;
        case B200SQP_DYN_TOY_EXAMPLE: launchOne<ToyExample>(dyn, integrator, dt, B, x, u, x_next, u_log, chi2_src, chi2_log, status_src, status_log, st); return true;
        case B200SQP_DYN_ARTSTEINS_CIRCLE: launchOne<ArtsteinsCircle>(dyn, integrator, dt, B, x, u, x_next, u_log, chi2_src, chi2_log, status_src, status_log, st); return true;
        case B200SQP_DYN_LINEAR_2X1: launchOne<LinearStateSpace2x1>(dyn, integrator, dt, B, x, u, x_next, u_log, chi2_src, chi2_log, status_src, status_log, st); return true;
        case B200SQP_DYN_LINEAR_3X1: launchOne<LinearStateSpace3x1>(dyn, integrator, dt, B, x, u, x_next, u_log, chi2_src, chi2_log, status_src, status_log, st); return true;
        case B200SQP_DYN_LINEAR_4X1: launchOne<LinearStateSpace4x1>(dyn, integrator, dt, B, x, u, x_next, u_log, chi2_src, chi2_log, status_src, status_log, st); return true;
        case B200SQP_DYN_LINEAR_4X2: launchOne<LinearStateSpace4x2>(dyn, integrator, dt, B, x, u, x_next, u_log, chi2_src, chi2_log, status_src, status_log, st); return true;
        case B200SQP_DYN_TRIPLE_INTEGRATOR: launchOne<TripleIntegrator>(dyn, integrator, dt, B, x, u, x_next, u_log, chi2_src, chi2_log, status_src, status_log, st); return true;
        case B200SQP_DYN_QUAD_INTEGRATOR: launchOne<QuadIntegrator>(dyn, integrator, dt, B, x, u, x_next, u_log, chi2_src, chi2_log, status_src, status_log, st); return true;
    }
    return false;
}

}  // namespace b200sqp
