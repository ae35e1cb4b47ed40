// Device functors of the closed system-dynamics registry (include/b200sqp.h: b200sqp_dynamics) and of the defect constraints
// built on them.
//
// The whole translation unit is compiled with --fmad=false: the reference evaluates these expressions with g++ on x86-64 without
// FMA contraction, and its central-difference Jacobians (delta = 1e-9) amplify every last-bit difference by 5e8
// (SURVEY.md "FD-noise parity").  Expression order below therefore follows the reference sources term by term; fused
// multiply-adds are only used where this project's own linear algebra asks for them explicitly via fma().
//
// Reference: corbo::SystemDynamicsInterface::dynamics, src/systems/include/corbo-systems/system_dynamics_interface.h:121
#pragma once

#include "../../include/b200sqp.h"
#include "dynamics_ids.h"
#include "lm_device_types.h"

namespace b200sqp {

// x / c.p[i] without an IEEE divide where the host found the parameter admissible (lm_device_types.h prepareDynParams): the same
// correctly rounded sequence as StepSize::div below, so results equal the reference's `x / p`.
__device__ __forceinline__ double divByParam(const DynParams& c, int i, double x)
{
    if (c.fast_div_mask & (1u << i))
    {
        const double q = x * c.rcp[i];
        const double r = fma(-q, c.p[i], x);
        return fma(r, c.rcp[i], q);
    }
    return x / c.p[i];
}

// scalar-type helpers for models that also exist in reduced precision (the fp32 variant of the warp-cooperative pipeline)
__device__ __forceinline__ void sincosT(double a, double* s, double* c) { sincos(a, s, c); }
__device__ __forceinline__ void sincosT(float a, float* s, float* c)
{
    // fp32 variant: the hardware approximation (MUFU.SIN / MUFU.COS after a range reduction to [-pi, pi], absolute error ~5e-7 there) -- the
    // finite-difference Jacobian of that variant carries 2e-5 already, and its trajectories are judged at 1e-3 against the fp64 path
    __sincosf(a, s, c);
}
__device__ __forceinline__ float divByParam(const DynParams& c, int i, float x)
{
    // fp32: the reciprocal the host prepared (rounded to float) where admissible; no bit-parity to keep in this precision
    return (c.fast_div_mask & (1u << i)) ? x * (float)c.rcp[i] : x / (float)c.p[i];
}

// VanDerPolOscillator::dynamics -- src/systems/include/corbo-systems/benchmark/nonlinear_benchmark_systems.h:52-60
struct VanDerPol
{
    static constexpr int NX = 2, NU = 1, ID = B200SQP_DYN_VAN_DER_POL;
    __device__ __forceinline__ static void f(const DynParams& c, const double* x, const double* u, double* out)
    {
        out[0] = x[1];
        out[1] = -c.p[0] * (x[0] * x[0] - 1) * x[1] - x[0] + u[0];
    }
};

// DuffingOscillator::dynamics -- nonlinear_benchmark_systems.h:108-117; p = damping, spring_alpha, spring_beta
struct Duffing
{
    static constexpr int NX = 2, NU = 1, ID = B200SQP_DYN_DUFFING;
    __device__ __forceinline__ static void f(const DynParams& c, const double* x, const double* u, double* out)
    {
        out[0] = x[1];
        out[1] = -c.p[0] * x[1] - c.p[1] * x[0] - c.p[2] * x[0] * x[0] * x[0] + u[0];
    }
};

// SimplePendulum::dynamics -- nonlinear_benchmark_systems.h:207-216; p = m, l, g, rho
struct SimplePendulum
{
    static constexpr int NX = 2, NU = 1, ID = B200SQP_DYN_SIMPLE_PENDULUM;
    // the angle x[0] enters only through its sine: finite-difference sweeps keep it across evaluations that do not move the angle
    static constexpr int NANG = 1, ANG0 = 0;
    __device__ __forceinline__ static void trig(const double* x, double* sc) { sc[0] = sin(x[0]); }
    __device__ __forceinline__ static void fTrig(const DynParams& c, const double* x, const double* u, const double* sc, double* out)
    {
        out[0] = x[1];
        out[1] = u[0] - c.p[3] / (c.p[0] * c.p[1] * c.p[1]) * x[1] - c.p[2] / c.p[1] * sc[0];
    }
    __device__ __forceinline__ static void f(const DynParams& c, const double* x, const double* u, double* out)
    {
        double sc[2];
        trig(x, sc);
        fTrig(c, x, u, sc, out);
    }
};

// CartPole::dynamics -- nonlinear_benchmark_systems.h:337-352; the reference keeps mc, mp, l, g as private constants (:390-394)
struct CartPole
{
    static constexpr int NX = 4, NU = 1, ID = B200SQP_DYN_CART_POLE;
    static constexpr unsigned X_DEPS = 0xEu;  // f never reads the cart position x[0]
    static constexpr int NANG = 1, ANG0 = 1;
    __device__ __forceinline__ static void trig(const double* x, double* sc)
    {
        sincos(x[1], &sc[0], &sc[1]);  // one argument reduction for both; same values as separate sin()/cos()
    }
    __device__ __forceinline__ static void f(const DynParams& c, const double* x, const double* u, double* out)
    {
        double sc[2];
        trig(x, sc);
        fTrig(c, x, u, sc, out);
    }
    __device__ __forceinline__ static void fTrig(const DynParams&, const double* x, const double* u, const double* sc, double* out)
    {
        const double mc = 1.0, mp = 0.3, l = 0.5, g = 9.81;
        const double s = sc[0], co = sc[1];
        double sin_phi_phidot_sq = s * x[3] * x[3];
        double denum             = mc + mp * (1 - co * co);  // std::pow(cos, 2) == cos*cos in IEEE arithmetic
        out[0]                   = x[2];
        out[1]                   = x[3];
        out[2]                   = (l * mp * sin_phi_phidot_sq + u[0] + mp * g * co * s) / denum;
        out[3]                   = -(l * mp * co * sin_phi_phidot_sq + u[0] * co + (mp + mc) * g * s) / (l * denum);
    }
};

// SerialIntegratorSystem(dimension 2)::dynamics -- linear_benchmark_systems.h:71-82; p = time constant
struct DoubleIntegrator
{
    static constexpr int NX = 2, NU = 1, ID = B200SQP_DYN_DOUBLE_INTEGRATOR;
    static constexpr unsigned X_DEPS = 0x2u;  // f reads x[1] only
    __device__ __forceinline__ static void f(const DynParams& c, const double* x, const double* u, double* out)
    {
        out[0] = x[1];
        out[1] = u[0] / c.p[0];
    }
};

// FreeSpaceRocket::dynamics -- nonlinear_benchmark_systems.h:174-183; x = (s, v, m), no parameters
struct FreeSpaceRocket
{
    static constexpr int NX = 3, NU = 1, ID = B200SQP_DYN_FREE_SPACE_ROCKET;
    static constexpr unsigned X_DEPS = 0x6u;  // f never reads the position x[0]
    __device__ __forceinline__ static void f(const DynParams&, const double* x, const double* u, double* out)
    {
        out[0] = x[1];
        out[1] = (u[0] - 0.02 * x[1] * x[1]) / x[2];
        out[2] = -0.01 * u[0] * u[0];
    }
};

// MasslessPendulum::dynamics -- nonlinear_benchmark_systems.h:281-290; p = omega0
struct MasslessPendulum
{
    static constexpr int NX = 2, NU = 1, ID = B200SQP_DYN_MASSLESS_PENDULUM;
    static constexpr int NANG = 1, ANG0 = 0;
    __device__ __forceinline__ static void trig(const double* x, double* sc) { sc[0] = sin(x[0]); }
    __device__ __forceinline__ static void fTrig(const DynParams& c, const double* x, const double* u, const double* sc, double* out)
    {
        out[0] = x[1];
        out[1] = u[0] - c.p[0] * sc[0];
    }
    __device__ __forceinline__ static void f(const DynParams& c, const double* x, const double* u, double* out)
    {
        double sc[2];
        trig(x, sc);
        fTrig(c, x, u, sc, out);
    }
};

// ToyExample::dynamics -- nonlinear_benchmark_systems.h:426-436; p = mu
struct ToyExample
{
    static constexpr int NX = 2, NU = 1, ID = B200SQP_DYN_TOY_EXAMPLE;
    __device__ __forceinline__ static void f(const DynParams& c, const double* x, const double* u, double* out)
    {
        out[0] = x[1] + u[0] * (c.p[0] + (1.0 - c.p[0]) * x[0]);
        out[1] = x[0] + u[0] * (c.p[0] - 4.0 * (1.0 - c.p[0]) * x[1]);
    }
};

// ArtsteinsCircle::dynamics -- nonlinear_benchmark_systems.h:483-492; no parameters
struct ArtsteinsCircle
{
    static constexpr int NX = 2, NU = 1, ID = B200SQP_DYN_ARTSTEINS_CIRCLE;
    __device__ __forceinline__ static void f(const DynParams&, const double* x, const double* u, double* out)
    {
        out[0] = (x[0] * x[0] - x[1] * x[1]) * u[0];
        out[1] = 2 * x[0] * x[1] * u[0];
    }
};

// LinearStateSpaceModel::dynamics, f = A x + B u -- linear_benchmark_systems.h:206-214; p = A column-major (NX*NX), then B
// column-major (NX*NU).  Summation order of Eigen's evaluation for fewer than four columns: the product A x accumulated column by
// column from zero, the product B u likewise into its own temporary, then their sum (bit-identical to the compiled reference for
// NX <= 3; from four columns on Eigen's kernel pairs the columns depending on the alignment of the destination, so NX = 4 agrees to
// rounding only).
template <int NX_, int NU_, int ID_>
struct LinearStateSpace
{
    static constexpr int NX = NX_, NU = NU_, ID = ID_;
    __device__ __forceinline__ static void f(const DynParams& c, const double* x, const double* u, double* out)
    {
#pragma unroll
        for (int i = 0; i < NX; ++i)
        {
            double ax = c.p[i] * x[0];
#pragma unroll
            for (int j = 1; j < NX; ++j) ax = ax + c.p[i + j * NX] * x[j];
            double bu = c.p[NX * NX + i] * u[0];
#pragma unroll
            for (int j = 1; j < NU; ++j) bu = bu + c.p[NX * NX + i + j * NX] * u[j];
            out[i] = ax + bu;
        }
    }
};
using LinearStateSpace2x1 = LinearStateSpace<2, 1, B200SQP_DYN_LINEAR_2X1>;
using LinearStateSpace3x1 = LinearStateSpace<3, 1, B200SQP_DYN_LINEAR_3X1>;
using LinearStateSpace4x1 = LinearStateSpace<4, 1, B200SQP_DYN_LINEAR_4X1>;
using LinearStateSpace4x2 = LinearStateSpace<4, 2, B200SQP_DYN_LINEAR_4X2>;

// SerialIntegratorSystem(dimension P)::dynamics for P = 3, 4 -- linear_benchmark_systems.h:71-82; p = time constant
template <int P, int ID_>
struct SerialIntegrator
{
    static constexpr int NX = P, NU = 1, ID = ID_;
    static constexpr unsigned X_DEPS = ((1u << P) - 1u) & ~1u;  // f never reads x[0]
    __device__ __forceinline__ static void f(const DynParams& c, const double* x, const double* u, double* out)
    {
#pragma unroll
        for (int i = 0; i < P - 1; ++i) out[i] = x[i + 1];
        out[P - 1] = u[0] / c.p[0];
    }
};
using TripleIntegrator = SerialIntegrator<3, B200SQP_DYN_TRIPLE_INTEGRATOR>;
using QuadIntegrator   = SerialIntegrator<4, B200SQP_DYN_QUAD_INTEGRATOR>;

// New model (absent from the reference; same equations as oracle/ref_models.h Unicycle)
struct Unicycle
{
    static constexpr int NX = 3, NU = 2, ID = B200SQP_DYN_UNICYCLE;
    static constexpr unsigned X_DEPS = 0x4u;  // f reads the heading x[2] only
    static constexpr int NANG = 1, ANG0 = 2;
    __device__ __forceinline__ static void trig(const double* x, double* sc)
    {
        sincos(x[2], &sc[0], &sc[1]);  // one argument reduction for both; same values as separate sin()/cos()
    }
    __device__ __forceinline__ static void fTrig(const DynParams&, const double*, const double* u, const double* sc, double* out)
    {
        out[0] = u[0] * sc[1];
        out[1] = u[0] * sc[0];
        out[2] = u[1];
    }
    __device__ __forceinline__ static void f(const DynParams& c, const double* x, const double* u, double* out)
    {
        double sc[2];
        trig(x, sc);
        fTrig(c, x, u, sc, out);
    }
};

// New model (absent from the reference; same equations as oracle/ref_models.h Quadrotor); p = m, g, Ixx, Iyy, Izz
struct Quadrotor
{
    static constexpr int NX = 12, NU = 4, ID = B200SQP_DYN_QUADROTOR;
    static constexpr unsigned X_DEPS = 0xFF8u;  // f never reads the position x[0..2]
    // The attitude angles x[3..5] enter only through their sines and cosines: callers that evaluate f at many points which share
    // most angles (the finite-difference columns of the pipeline, lm_pipeline.cuh) keep the six values and refresh one pair.
    static constexpr int NANG = 3, ANG0 = 3;
    // f is affine in u (thrust and torques enter linearly): f(x, u) = (f(x, u + d e_j) + f(x, u - d e_j)) / 2 up to rounding
    static constexpr bool CONTROL_AFFINE = true;
    // templated on the scalar type: double everywhere, float only in the reduced-precision pipeline (b200sqp_set_precision)
    template <class T>
    __device__ __forceinline__ static void trig(const T* x, T* sc /*[2*NANG]: sin, cos per angle*/)
    {
        sincosT(x[3], &sc[0], &sc[1]);  // one argument reduction per angle; same values as separate sin()/cos()
        sincosT(x[4], &sc[2], &sc[3]);
        sincosT(x[5], &sc[4], &sc[5]);
    }
    template <class T>
    __device__ __forceinline__ static void fTrig(const DynParams& c, const T* x, const T* u, const T* sc, T* out)
    {
        const T sphi = sc[0], cphi = sc[1], sth = sc[2], cth = sc[3], spsi = sc[4], cpsi = sc[5];
        const T p = x[9], q = x[10], r = x[11];
        const T tm = divByParam(c, 0, u[0]);
        out[0]     = x[6];
        out[1]     = x[7];
        out[2]     = x[8];
        const T qr = q * sphi + r * cphi;
        out[3]     = p + qr * (sth / cth);
        out[4]     = q * cphi - r * sphi;
        out[5]     = qr / cth;
        out[6]     = (cphi * sth * cpsi + sphi * spsi) * tm;
        out[7]     = (cphi * sth * spsi - sphi * cpsi) * tm;
        out[8]     = cphi * cth * tm - (T)c.p[1];
        out[9]     = divByParam(c, 2, u[1] + (T)(c.p[3] - c.p[4]) * q * r);
        out[10]    = divByParam(c, 3, u[2] + (T)(c.p[4] - c.p[2]) * p * r);
        out[11]    = divByParam(c, 4, u[3] + (T)(c.p[2] - c.p[3]) * p * q);
    }
    __device__ __forceinline__ static void f(const DynParams& c, const double* x, const double* u, double* out)
    {
        double sc[2 * NANG];
        trig(x, sc);
        fTrig(c, x, u, sc, out);
    }
};

// ---------------------------------------------------------------------------------------------------------------------------
// Defect of one interval, e(x1, u1, x2, dt): FDCollocationEdge::computeValues
// (src/optimal_control/include/corbo-optimal-control/structured_ocp/edges/finite_differences_collocation_edges.h:71-80) over
// FiniteDifferencesCollocationInterface::computeEqualityConstraint (src/numerics/include/corbo-numerics/finite_differences_collocation.h)
// and MSVariableDynamicsOnlyEdge::computeValues (.../edges/multiple_shooting_edges.h:125-134) over
// NumericalIntegratorExplicitInterface::computeEqualityConstraint (src/numerics/include/corbo-numerics/integrator_interface.h:217-222).
// DEFECT ids: 0..3 = b200sqp_collocation, 4 = explicit Euler shooting, 5 = RK4 shooting.
// ---------------------------------------------------------------------------------------------------------------------------

// x / dt for many x and one dt, correctly rounded like the reference's IEEE division but without a divide per call.
// With z = RN(1/dt) (one IEEE division per distinct dt):  q = RN(x z);  r = x - q dt (exact in one fma);  x/dt = RN(q + r z).
// This three-operation sequence returns RN(x/dt) for every x unless dt's significand is all ones (Brisebarre, Muller, Raina,
// "Accelerating correctly rounded floating-point division when the divisor is known in advance", IEEE TC 53(8), 2004, Alg. 1 /
// Markstein's theorem), barring over/underflow of the intermediates: dt with an all-ones significand or an extreme exponent
// takes the plain division.  x = +-inf gives NaN instead of +-inf; both make chi2 non-finite, which the LM loop rejects
// identically (levenberg_marquardt_sparse.cpp:171).  Profile that motivated it: the divide in the Crank-Nicolson defect was
// 27 % of all executed instructions of the LM kernel (profiles/r1c_*).  tests/test_division_by_step.py checks the sequence
// against exact rational arithmetic on hard-to-round cases.
struct StepSize
{
    double dt, rcp;
    bool fast;
    __device__ __forceinline__ explicit StepSize(double t) : dt(t)
    {
        rcp                          = 1.0 / t;
        const unsigned long long b   = (unsigned long long)__double_as_longlong(t);
        const unsigned long long man = b & 0xFFFFFFFFFFFFFull;
        const unsigned ex            = (unsigned)(b >> 52) & 0x7FFu;
        fast                         = man != 0xFFFFFFFFFFFFFull && ex > 523u && ex < 1523u;  // 2^-500 < |dt| < 2^500
    }
    __device__ __forceinline__ double div(double x) const
    {
        if (fast)
        {
            const double q = x * rcp;
            const double r = fma(-q, dt, x);
            return fma(r, rcp, q);
        }
        return x / dt;
    }
};

template <class M, int DEFECT>
__device__ __forceinline__ void defect(const DynParams& c, const double* x1, const double* u1, const double* x2, const StepSize& h, double* e)
{
    constexpr int NX = M::NX;
    const double dt  = h.dt;
    if (DEFECT == DEFECT_FORWARD)  // finite_differences_collocation.h:119-136
    {
        M::f(c, x1, u1, e);
#pragma unroll
        for (int i = 0; i < NX; ++i) e[i] -= h.div(x2[i] - x1[i]);
    }
    else if (DEFECT == DEFECT_BACKWARD)  // :153-170
    {
        M::f(c, x2, u1, e);
#pragma unroll
        for (int i = 0; i < NX; ++i) e[i] -= h.div(x2[i] - x1[i]);
    }
    else if (DEFECT == DEFECT_MIDPOINT)  // :187-204
    {
        double xm[NX];
#pragma unroll
        for (int i = 0; i < NX; ++i) xm[i] = 0.5 * (x1[i] + x2[i]);
        M::f(c, xm, u1, e);
#pragma unroll
        for (int i = 0; i < NX; ++i) e[i] -= h.div(x2[i] - x1[i]);
    }
    else if (DEFECT == DEFECT_CRANK_NICOLSON)  // :221-240
    {
        double f1[NX], f2[NX];
        M::f(c, x1, u1, f1);
        M::f(c, x2, u1, f2);
#pragma unroll
        for (int i = 0; i < NX; ++i) e[i] = h.div(x2[i] - x1[i]) - 0.5 * (f1[i] + f2[i]);
    }
    else if (DEFECT == DEFECT_EULER)  // explicit_integrators.h:66-72, then "- x2"
    {
        double k1[NX];
        M::f(c, x1, u1, k1);
#pragma unroll
        for (int i = 0; i < NX; ++i) e[i] = (k1[i] * dt + x1[i]) - x2[i];
    }
    else  // RK4: explicit_integrators.h:280-295, then "- x2"
    {
        double k1[NX], k2[NX], k3[NX], k4[NX], xt[NX];
        const StepSize six(6.0);  // `/ 6.0` correctly rounded without a divide per component (constant-folded reciprocal)
        M::f(c, x1, u1, k1);
#pragma unroll
        for (int i = 0; i < NX; ++i) k1[i] *= dt;
#pragma unroll
        for (int i = 0; i < NX; ++i) xt[i] = x1[i] + k1[i] / 2.0;
        M::f(c, xt, u1, k2);
#pragma unroll
        for (int i = 0; i < NX; ++i) k2[i] *= dt;
#pragma unroll
        for (int i = 0; i < NX; ++i) xt[i] = x1[i] + k2[i] / 2.0;
        M::f(c, xt, u1, k3);
#pragma unroll
        for (int i = 0; i < NX; ++i) k3[i] *= dt;
#pragma unroll
        for (int i = 0; i < NX; ++i) xt[i] = x1[i] + k3[i];
        M::f(c, xt, u1, k4);
#pragma unroll
        for (int i = 0; i < NX; ++i) k4[i] *= dt;
#pragma unroll
        for (int i = 0; i < NX; ++i) e[i] = (x1[i] + six.div(k1[i] + 2.0 * k2[i] + 2.0 * k3[i] + k4[i])) - x2[i];
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// The same defects split into the parts a finite-difference sweep can REUSE.  BaseEdge::computeJacobian (edge_interface.cpp:55-96)
// re-evaluates the whole edge for every perturbed component, but most of those evaluations repeat work on bit-identical inputs: with
// e.g. Crank-Nicolson, e = (x2 - x1)/dt - 0.5 (f(x1,u) + f(x2,u)), perturbing a component of x1 leaves f(x2,u) untouched, perturbing dt
// leaves both, and a component the dynamics never read (StateDeps) leaves f altogether.  Identical inputs give identical outputs, so
// evaluating each DISTINCT function part once and assembling the defects from the parts yields the reference's numbers bit for bit
// with 40-75 % fewer dynamics evaluations (Van der Pol CN 22 -> 17 per interval, unicycle CN 38 -> 19, cart-pole RK4 76 -> 44).
//   part A: the function part that reads x1 (f(x1,u); Euler k1; the RK4 increment; midpoint: f of the mean state, which also reads x2)
//   part B: the function part that reads only x2 (f(x2,u): Crank-Nicolson, backward differences)
//   assemble(): the cheap remainder, in the reference's expression order (it is the code of defect<> above, term by term)
// ---------------------------------------------------------------------------------------------------------------------------
// bit i set: f reads x[i] (default: all).  A model declares `static constexpr unsigned X_DEPS` to opt in.
template <class M, class = void>
struct StateDeps
{
    static constexpr unsigned mask = 0xFFFFFFFFu;
};
template <class M>
struct StateDeps<M, decltype((void)M::X_DEPS)>
{
    static constexpr unsigned mask = M::X_DEPS;
};

// models whose angles enter only through sines / cosines expose trig() / fTrig() and NANG angles starting at state index ANG0
template <class M, class = void>
struct HasTrig
{
    static constexpr bool value = false;
    static constexpr int slots  = 1;
};
template <class M>
struct HasTrig<M, decltype((void)M::NANG)>
{
    static constexpr bool value = true;
    static constexpr int slots  = 2 * M::NANG;
};

template <class M, int DEFECT>
struct DefectParts
{
    static constexpr int NX       = M::NX;
    // the sines / cosines behind one function part can be carried between evaluations whose angles did not move (REUSE): rules whose
    // part evaluates f at a grid state itself (not at a mean state or at inner Runge-Kutta stages)
    static constexpr bool TRIG = HasTrig<M>::value && (DEFECT == DEFECT_FORWARD || DEFECT == DEFECT_BACKWARD || DEFECT == DEFECT_CRANK_NICOLSON || DEFECT == DEFECT_EULER);
    struct Trig
    {
        double sc[HasTrig<M>::slots];
    };
    static constexpr bool hasA    = DEFECT != DEFECT_BACKWARD;
    static constexpr bool hasB    = DEFECT == DEFECT_CRANK_NICOLSON || DEFECT == DEFECT_BACKWARD;
    static constexpr bool A_on_x2 = DEFECT == DEFECT_MIDPOINT;  // part A also reads x2
    static constexpr bool A_on_dt = DEFECT == DEFECT_RK4;       // part A also reads dt

    __device__ __forceinline__ static void evalAInline(const DynParams& c, const double* x1, const double* x2, const double* u1, const StepSize& h, double* pA)
    {
        if (DEFECT == DEFECT_FORWARD || DEFECT == DEFECT_CRANK_NICOLSON || DEFECT == DEFECT_EULER)
            M::f(c, x1, u1, pA);
        else if (DEFECT == DEFECT_MIDPOINT)
        {
            double xm[NX];
#pragma unroll
            for (int i = 0; i < NX; ++i) xm[i] = 0.5 * (x1[i] + x2[i]);
            M::f(c, xm, u1, pA);
        }
        else if (DEFECT == DEFECT_RK4)
        {
            const double dt = h.dt;
            double k1[NX], k2[NX], k3[NX], k4[NX], xt[NX];
            const StepSize six(6.0);
            M::f(c, x1, u1, k1);
#pragma unroll
            for (int i = 0; i < NX; ++i) k1[i] *= dt;
#pragma unroll
            for (int i = 0; i < NX; ++i) xt[i] = x1[i] + k1[i] / 2.0;
            M::f(c, xt, u1, k2);
#pragma unroll
            for (int i = 0; i < NX; ++i) k2[i] *= dt;
#pragma unroll
            for (int i = 0; i < NX; ++i) xt[i] = x1[i] + k2[i] / 2.0;
            M::f(c, xt, u1, k3);
#pragma unroll
            for (int i = 0; i < NX; ++i) k3[i] *= dt;
#pragma unroll
            for (int i = 0; i < NX; ++i) xt[i] = x1[i] + k3[i];
            M::f(c, xt, u1, k4);
#pragma unroll
            for (int i = 0; i < NX; ++i) k4[i] *= dt;
#pragma unroll
            for (int i = 0; i < NX; ++i) pA[i] = six.div(k1[i] + 2.0 * k2[i] + 2.0 * k3[i] + k4[i]);
        }
    }
    __device__ __forceinline__ static void evalBInline(const DynParams& c, const double* x2, const double* u1, double* pB)
    {
        if (hasB) M::f(c, x2, u1, pB);
    }
    __device__ __noinline__ static void evalAOutlined(const DynParams& c, const double* x1, const double* x2, const double* u1, const StepSize& h, double* pA)
    {
        evalAInline(c, x1, x2, u1, h, pA);
    }
    __device__ __noinline__ static void evalBOutlined(const DynParams& c, const double* x2, const double* u1, double* pB) { evalBInline(c, x2, u1, pB); }
    __device__ __noinline__ static void fTrigOutlined(const DynParams& c, const double* x, const double* u1, const double* sc, double* out)
    {
        if constexpr (HasTrig<M>::value) M::fTrig(c, x, u1, sc, out);
    }
    // large models keep ONE out-of-line copy of each part (see defectOutlined below).  REUSE: the angles of the state this part reads
    // have not moved since `t` was filled.
    template <bool REUSE = false>
    __device__ __forceinline__ static void evalA(const DynParams& c, const double* x1, const double* x2, const double* u1, const StepSize& h, double* pA,
                                                 Trig& t)
    {
        if constexpr (!hasA)
            return;
        else if constexpr (TRIG)
        {
            if (!REUSE) M::trig(x1, t.sc);
            if constexpr (M::NX >= 8)
                fTrigOutlined(c, x1, u1, t.sc, pA);
            else
                M::fTrig(c, x1, u1, t.sc, pA);
        }
        else if constexpr (M::NX >= 8)
            evalAOutlined(c, x1, x2, u1, h, pA);
        else
            evalAInline(c, x1, x2, u1, h, pA);
    }
    template <bool REUSE = false>
    __device__ __forceinline__ static void evalB(const DynParams& c, const double* x2, const double* u1, double* pB, Trig& t)
    {
        if constexpr (!hasB)
            return;
        else if constexpr (TRIG)
        {
            if (!REUSE) M::trig(x2, t.sc);
            if constexpr (M::NX >= 8)
                fTrigOutlined(c, x2, u1, t.sc, pB);
            else
                M::fTrig(c, x2, u1, t.sc, pB);
        }
        else if constexpr (M::NX >= 8)
            evalBOutlined(c, x2, u1, pB);
        else
            evalBInline(c, x2, u1, pB);
    }
    // is state component c one of the model's angles?
    __device__ __forceinline__ static constexpr bool isAngle(int c)
    {
        if constexpr (HasTrig<M>::value)
            return c >= M::ANG0 && c < M::ANG0 + M::NANG;
        else
            return false;
    }
    __device__ __forceinline__ static void assemble(const double* x1, const double* x2, const StepSize& h, const double* pA, const double* pB, double* e)
    {
#pragma unroll
        for (int i = 0; i < NX; ++i)
        {
            if (DEFECT == DEFECT_FORWARD || DEFECT == DEFECT_MIDPOINT)
                e[i] = pA[i] - h.div(x2[i] - x1[i]);
            else if (DEFECT == DEFECT_BACKWARD)
                e[i] = pB[i] - h.div(x2[i] - x1[i]);
            else if (DEFECT == DEFECT_CRANK_NICOLSON)
                e[i] = h.div(x2[i] - x1[i]) - 0.5 * (pA[i] + pB[i]);
            else if (DEFECT == DEFECT_EULER)
                e[i] = (pA[i] * h.dt + x1[i]) - x2[i];
            else
                e[i] = (x1[i] + pA[i]) - x2[i];
        }
    }
};

// Large models (quadrotor: 56 defect evaluations per interval, each with two 12-state dynamics calls) are evaluated through one
// out-of-line copy of the defect instead of 57 inlined ones: the inlined form is megabytes of straight-line code that neither
// the instruction cache nor ptxas (10 minutes) copes with.  Small models keep the fully inlined, register-resident form.
template <class M, int DEFECT>
__device__ __noinline__ void defectOutlined(const DynParams& c, const double* x1, const double* u1, const double* x2, const StepSize& h, double* e)
{
    defect<M, DEFECT>(c, x1, u1, x2, h, e);
}

template <class M, int DEFECT>
__device__ __forceinline__ void defectCall(const DynParams& c, const double* x1, const double* u1, const double* x2, const StepSize& h, double* e)
{
    if constexpr (M::NX >= 8)
        defectOutlined<M, DEFECT>(c, x1, u1, x2, h, e);
    else
        defect<M, DEFECT>(c, x1, u1, x2, h, e);
}

}  // namespace b200sqp
