// Batched numerical linearisation of the system dynamics: A = df/dx, B = df/du at (x_i, u_i) for a batch of points, by the
// reference's own finite-difference rules -- the device counterpart of
//   SystemDynamicsInterface::getLinearA / getLinearB      src/systems/src/system_dynamics_interface.cpp:33-59
// over
//   ForwardDifferences::jacobian (delta = 1e-9)           src/numerics/include/corbo-numerics/finite_differences.hpp:29-48
//   CentralDifferences::jacobian (delta = 1e-9)           src/numerics/include/corbo-numerics/finite_differences.hpp:167-188
// (SURVEY.md section 8 row a15).  Like the reference the perturbed vector is modified in place (x[i] += delta; ...; x[i] -= delta),
// so later columns see the rounding drift of earlier ones; expressions are compiled without FMA contraction (this TU is built
// with --fmad=false) and the results are bit-identical to the reference for the polynomial models.
// One thread per point; inputs/outputs are instance-major [B][..] in HBM exactly as the host hands them over (A, B column-major
// per point, Eigen's default), reads and writes of a warp are strided by nx / nx*nx doubles -- this is a helper off the LM hot
// path (tens of microseconds for 64k points), not a tuned kernel.
#include "dynamics.cuh"
#include "launch.h"

namespace b200sqp {

namespace {

template <class M>
__global__ void linearizeDynamicsKernel(const DynParams dyn, int method, int B, const double* __restrict__ xs, const double* __restrict__ us,
                                        double* __restrict__ As, double* __restrict__ Bs)
{
    constexpr int NX = M::NX, NU = M::NU;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    double x[NX], u[NU], f0[NX], f1[NX];
#pragma unroll
    for (int j = 0; j < NX; ++j) x[j] = xs[(size_t)i * NX + j];
#pragma unroll
    for (int j = 0; j < NU; ++j) u[j] = us[(size_t)i * NU + j];
    constexpr double delta = 1e-9, ddelta = 2 * delta;
    // ---- A: perturb a copy of x (getLinearA :40-44)
    if (As)
    {
        double xp[NX];
#pragma unroll
        for (int j = 0; j < NX; ++j) xp[j] = x[j];
        double* A = As + (size_t)i * NX * NX;
        if (method == 0)
        {
            constexpr double scalar = 1.0 / delta;
            M::f(dyn, xp, u, f0);
#pragma unroll
            for (int c = 0; c < NX; ++c)
            {
                xp[c] += delta;
                M::f(dyn, xp, u, f1);
                xp[c] += -delta;
#pragma unroll
                for (int r = 0; r < NX; ++r) A[c * NX + r] = scalar * (f1[r] - f0[r]);
            }
        }
        else
        {
            constexpr double scalar = 1.0 / ddelta;
#pragma unroll
            for (int c = 0; c < NX; ++c)
            {
                xp[c] += delta;
                M::f(dyn, xp, u, f1);
                xp[c] += -ddelta;
                M::f(dyn, xp, u, f0);
#pragma unroll
                for (int r = 0; r < NX; ++r) A[c * NX + r] = scalar * (f1[r] - f0[r]);
                xp[c] += delta;
            }
        }
    }
    // ---- B: perturb a copy of u, x stays the caller's x0 (getLinearB :54-58)
    if (Bs)
    {
        double up[NU];
#pragma unroll
        for (int j = 0; j < NU; ++j) up[j] = u[j];
        double* Bm = Bs + (size_t)i * NX * NU;
        if (method == 0)
        {
            constexpr double scalar = 1.0 / delta;
            M::f(dyn, x, up, f0);
#pragma unroll
            for (int c = 0; c < NU; ++c)
            {
                up[c] += delta;
                M::f(dyn, x, up, f1);
                up[c] += -delta;
#pragma unroll
                for (int r = 0; r < NX; ++r) Bm[c * NX + r] = scalar * (f1[r] - f0[r]);
            }
        }
        else
        {
            constexpr double scalar = 1.0 / ddelta;
#pragma unroll
            for (int c = 0; c < NU; ++c)
            {
                up[c] += delta;
                M::f(dyn, x, up, f1);
                up[c] += -ddelta;
                M::f(dyn, x, up, f0);
#pragma unroll
                for (int r = 0; r < NX; ++r) Bm[c * NX + r] = scalar * (f1[r] - f0[r]);
                up[c] += delta;
            }
        }
    }
}

template <class M>
void launchOne(const DynParams& dyn, int method, int B, const double* x, const double* u, double* A, double* Bm, cudaStream_t st)
{
    linearizeDynamicsKernel<M><<<(B + 127) / 128, 128, 0, st>>>(dyn, method, B, x, u, A, Bm);
}

}  // namespace

bool launchLinearizeDynamics(int dynamics, const DynParams& dyn, int method, int B, const double* x, const double* u, double* A, double* Bm,
                             cudaStream_t st)
{
    switch (dynamics)
    {
        case B200SQP_DYN_VAN_DER_POL: launchOne<VanDerPol>(dyn, method, B, x, u, A, Bm, st); return true;
        case B200SQP_DYN_DUFFING: launchOne<Duffing>(dyn, method, B, x, u, A, Bm, st); return true;
        case B200SQP_DYN_SIMPLE_PENDULUM: launchOne<SimplePendulum>(dyn, method, B, x, u, A, Bm, st); return true;
        case B200SQP_DYN_CART_POLE: launchOne<CartPole>(dyn, method, B, x, u, A, Bm, st); return true;
        case B200SQP_DYN_DOUBLE_INTEGRATOR: launchOne<DoubleIntegrator>(dyn, method, B, x, u, A, Bm, st); return true;
        case B200SQP_DYN_UNICYCLE: launchOne<Unicycle>(dyn, method, B, x, u, A, Bm, st); return true;
        case B200SQP_DYN_QUADROTOR: launchOne<Quadrotor>(dyn, method, B, x, u, A, Bm, st); return true;
    }
    return false;
}

}  // namespace b200sqp
