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

// Second derivatives of the system dynamics w.r.t. z = [x; u] by the reference's finite-difference Hessian rules (delta = 1e-5):
//   ForwardDifferences::hessian   src/numerics/include/corbo-numerics/finite_differences.hpp:50-104
//                                 H(i,j) = 1/d^2 sum_v (f(z+d e_i+d e_j) - f(z+d e_i) - f(z+d e_j) + f(z))_v [* multipliers_v]
//   CentralDifferences::hessian   src/numerics/include/corbo-numerics/finite_differences.hpp:190-273
//                                 H(i,i) = 1/d^2 sum_v (f(z+d e_i) - 2 f(z) + f(z-d e_i))_v,
//                                 H(i,j) = 1/(4 d^2) sum_v (f(++) - f(+-) - f(-+) + f(--))_v
// in the reference's evaluation and increment order (inc_fun modifies z in place; no symmetry is exploited, every (i, j) pair is
// evaluated; the drift of earlier pairs is seen by later ones).  One thread per point, H [B][nz*nz] with H(i,j) at [j*nz + i]
// (column-major, Eigen's default).
template <class M>
__global__ void dynamicsHessianKernel(const DynParams dyn, int method, int B, const double* __restrict__ xs, const double* __restrict__ us,
                                      const double* __restrict__ mult, double* __restrict__ Hs)
{
    constexpr int NX = M::NX, NU = M::NU, NZ = NX + NU;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= B) return;
    double z[NZ], m[NX];
#pragma unroll
    for (int j = 0; j < NX; ++j) z[j] = xs[(size_t)p * NX + j];
#pragma unroll
    for (int j = 0; j < NU; ++j) z[NX + j] = us[(size_t)p * NU + j];
#pragma unroll
    for (int j = 0; j < NX; ++j) m[j] = mult ? mult[(size_t)p * NX + j] : 1.0;
    double* H = Hs + (size_t)p * NZ * NZ;
    constexpr double delta = 1e-5, ddelta = 2 * delta;
    double fa[NX], fb[NX], fc[NX], fd[NX];
    // runtime-indexed increments of a register array would spill it: select the component with a predicated unrolled loop
    auto inc = [&](int idx, double d) {
#pragma unroll
        for (int q = 0; q < NZ; ++q)
            if (q == idx) z[q] += d;
    };
    auto eval = [&](double* out) { M::f(dyn, z, z + NX, out); };
    for (int i = 0; i < NZ; ++i)
    {
        for (int j = 0; j < NZ; ++j)
        {
            double h;
            if (method == 0)
            {
                constexpr double scalar = 1 / (delta * delta);
                inc(i, delta);
                eval(fa);  // f1 = f(x+h, y)
                inc(j, delta);
                eval(fc);  // f3 = f(x+h, y+h)
                inc(i, -delta);
                eval(fb);  // f2 = f(x, y+h)
                inc(j, -delta);
                eval(fd);  // f0
                h = mult ? scalar * (fc[0] - fa[0] - fb[0] + fd[0]) * m[0] : scalar * (fc[0] - fa[0] - fb[0] + fd[0]);
#pragma unroll
                for (int v = 1; v < NX; ++v) h += mult ? scalar * (fc[v] - fa[v] - fb[v] + fd[v]) * m[v] : scalar * (fc[v] - fa[v] - fb[v] + fd[v]);
            }
            else if (i == j)
            {
                constexpr double scalar_xx = 1 / (delta * delta);
                inc(i, delta);
                eval(fa);  // f1
                inc(i, -ddelta);
                eval(fc);  // f3
                inc(i, delta);
                eval(fb);  // f2
                h = mult ? scalar_xx * (fa[0] - 2 * fb[0] + fc[0]) * m[0] : scalar_xx * (fa[0] - 2 * fb[0] + fc[0]);
#pragma unroll
                for (int v = 1; v < NX; ++v) h += mult ? scalar_xx * (fa[v] - 2 * fb[v] + fc[v]) * m[v] : scalar_xx * (fa[v] - 2 * fb[v] + fc[v]);
            }
            else
            {
                constexpr double scalar_xy = 1 / (4.0 * delta * delta);
                inc(i, delta);
                inc(j, delta);
                eval(fa);  // f1 = f(x+h, y+h)
                inc(j, -ddelta);
                eval(fb);  // f2 = f(x+h, y-h)
                inc(i, -ddelta);
                eval(fd);  // f4 = f(x-h, y-h)
                inc(j, ddelta);
                eval(fc);  // f3 = f(x-h, y+h)
                inc(i, delta);
                inc(j, -delta);
                h = mult ? scalar_xy * (fa[0] - fb[0] - fc[0] + fd[0]) * m[0] : scalar_xy * (fa[0] - fb[0] - fc[0] + fd[0]);
#pragma unroll
                for (int v = 1; v < NX; ++v) h += mult ? scalar_xy * (fa[v] - fb[v] - fc[v] + fd[v]) * m[v] : scalar_xy * (fa[v] - fb[v] - fc[v] + fd[v]);
            }
            H[(size_t)j * NZ + i] = h;
        }
    }
}

template <class M>
void launchHess(const DynParams& dyn, int method, int B, const double* x, const double* u, const double* mult, double* H, cudaStream_t st)
{
    dynamicsHessianKernel<M><<<(B + 63) / 64, 64, 0, st>>>(dyn, method, B, x, u, mult, H);
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
        case B200SQP_DYN_FREE_SPACE_ROCKET: launchOne<FreeSpaceRocket>(dyn, method, B, x, u, A, Bm, st); return true;
        case B200SQP_DYN_MASSLESS_PENDULUM: launchOne<MasslessPendulum>(dyn, method, B, x, u, A, Bm, st); return true;
        case B200SQP_DYN_TOY_EXAMPLE: launchOne<ToyExample>(dyn, method, B, x, u, A, Bm, st); return true;
        case B200SQP_DYN_ARTSTEINS_CIRCLE: launchOne<ArtsteinsCircle>(dyn, method, B, x, u, A, Bm, st); return true;
        case B200SQP_DYN_LINEAR_2X1: launchOne<LinearStateSpace2x1>(dyn, method, B, x, u, A, Bm, st); return true;
        case B200SQP_DYN_LINEAR_3X1: launchOne<LinearStateSpace3x1>(dyn, method, B, x, u, A, Bm, st); return true;
        case B200SQP_DYN_LINEAR_4X1: launchOne<LinearStateSpace4x1>(dyn, method, B, x, u, A, Bm, st); return true;
        case B200SQP_DYN_LINEAR_4X2: launchOne<LinearStateSpace4x2>(dyn, method, B, x, u, A, Bm, st); return true;
        case B200SQP_DYN_TRIPLE_INTEGRATOR: launchOne<TripleIntegrator>(dyn, method, B, x, u, A, Bm, st); return true;
        case B200SQP_DYN_QUAD_INTEGRATOR: launchOne<QuadIntegrator>(dyn, method, B, x, u, A, Bm, st); return true;
    }
    return false;
}

bool launchDynamicsHessian(int dynamics, const DynParams& dyn, int method, int B, const double* x, const double* u, const double* mult, double* H,
                           cudaStream_t st)
{
    switch (dynamics)
    {
        case B200SQP_DYN_VAN_DER_POL: launchHess<VanDerPol>(dyn, method, B, x, u, mult, H, st); return true;
        case B200SQP_DYN_DUFFING: launchHess<Duffing>(dyn, method, B, x, u, mult, H, st); return true;
        case B200SQP_DYN_SIMPLE_PENDULUM: launchHess<SimplePendulum>(dyn, method, B, x, u, mult, H, st); return true;
        case B200SQP_DYN_CART_POLE: launchHess<CartPole>(dyn, method, B, x, u, mult, H, st); return true;
        case B200SQP_DYN_DOUBLE_INTEGRATOR: launchHess<DoubleIntegrator>(dyn, method, B, x, u, mult, H, st); return true;
        case B200SQP_DYN_UNICYCLE: launchHess<Unicycle>(dyn, method, B, x, u, mult, H, st); return true;
        case B200SQP_DYN_QUADROTOR: launchHess<Quadrotor>(dyn, method, B, x, u, mult, H, st); return true;
        case B200SQP_DYN_FREE_SPACE_ROCKET: launchHess<FreeSpaceRocket>(dyn, method, B, x, u, mult, H, st); return true;
        case B200SQP_DYN_MASSLESS_PENDULUM: launchHess<MasslessPendulum>(dyn, method, B, x, u, mult, H, st); return true;
        case B200SQP_DYN_TOY_EXAMPLE: launchHess<ToyExample>(dyn, method, B, x, u, mult, H, st); return true;
        case B200SQP_DYN_ARTSTEINS_CIRCLE: launchHess<ArtsteinsCircle>(dyn, method, B, x, u, mult, H, st); return true;
        case B200SQP_DYN_LINEAR_2X1: launchHess<LinearStateSpace2x1>(dyn, method, B, x, u, mult, H, st); return true;
        case B200SQP_DYN_LINEAR_3X1: launchHess<LinearStateSpace3x1>(dyn, method, B, x, u, mult, H, st); return true;
        case B200SQP_DYN_LINEAR_4X1: launchHess<LinearStateSpace4x1>(dyn, method, B, x, u, mult, H, st); return true;
        case B200SQP_DYN_LINEAR_4X2: launchHess<LinearStateSpace4x2>(dyn, method, B, x, u, mult, H, st); return true;
        case B200SQP_DYN_TRIPLE_INTEGRATOR: launchHess<TripleIntegrator>(dyn, method, B, x, u, mult, H, st); return true;
        case B200SQP_DYN_QUAD_INTEGRATOR: launchHess<QuadIntegrator>(dyn, method, B, x, u, mult, H, st); return true;
    }
    return false;
}

}  // namespace b200sqp
