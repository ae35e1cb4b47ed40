// Warp-cooperative LM pipeline for models with LARGE stage blocks (nb = nu + nx >= 8, e.g. the 12-state quadrotor: 16 x 16 blocks).
//
// The fused one-thread-per-chunk kernel (lm_kernels.cuh) keeps a whole interval's Jacobian and a whole Hessian block in the
// registers of one thread; for nb = 16 that is > 1000 doubles per thread and ptxas spills 20 kB per thread to local memory.
// Here the 32 lanes of a warp share that state instead:
//
//   pipeLinearizeKernel   one warp per (instance, interval): lane c evaluates central-difference column c of the interval's
//                         Jacobian (2 nx + nu columns; BaseEdge::computeJacobian, edge_interface.cpp:55-96), the columns meet in
//                         shared memory, the lanes then share out the entries of the normal-equation blocks J^T J, J^T(-r)
//                         (levenberg_marquardt_sparse.cpp:97-100); residual norms by warp-shuffle reductions
//   pipeInitKernel        one warp per instance: chi2, ||g||inf, max diag(H) -> initial LM state (:103-126)
//   pipeFactorKernel      one warp per instance: block-tridiagonal Cholesky of (H + sum(mu) I); lane = row of the 16 x 16 block in
//                         registers, pivot columns broadcast by shuffles, the next block prefetched from HBM by TMA bulk copies
//                         (cp.async.bulk + mbarrier) while the current one is eliminated; forward and backward substitution (:135-148)
//   pipeTrialKernel       one thread per (instance, interval): trial point and its residuals (:158-167)
//   pipeControlKernel     one thread per instance: gain ratio, accept/reject, damping (:169-216), the reference's loop verbatim
//
// The host drives the passes (launchPipeline) and reads two flags per pass; one pass is milliseconds of device work for the
// batches this path is meant for, so the launch/flag latency is irrelevant.  Hessian blocks live in HBM instance-major
// ([instance][interval][entry]: a warp reads its instance's block as one contiguous run); parameters, steps and per-instance
// scalars stay in the tiled instance-minor layout of the rest of the library, so every other entry point works unchanged.
//
// Differences to the fused kernel (both within the FD-noise floor): the in-place perturbation drift of the
// parameters (+d, -2d, +d leaves ~1 ulp) is applied inside one Jacobian evaluation exactly as the reference orders it, but not
// carried into the stored parameters between evaluations; the elimination order inside a block differs.
// Reduced precision (BASELINE configs[4] names fp32; b200sqp_set_precision): every kernel below is templated on the scalar type `Real`.
// With Real = float the Jacobian columns are central differences with delta = 2^-10 evaluated in fp32 (the reference's delta = 1e-9
// does not exist in fp32), the normal equations, their Cholesky factor and the substitutions are fp32 (half the HBM traffic of the
// factorisation, twice the FMA rate, cheap sincosf), the Gram products run on the FMA pipe instead of the fp64 tensor-core tiles;
// parameters, steps, the trial-point residuals (pipeTrialKernel) and the whole LM control state stay fp64.  Parity of that variant is
// judged against the fp64 oracle at 1e-3 relative on the trajectories, not bit-wise.
// Eligibility (else the fused kernel runs): fixed-dt FD grid or shooting grid, quadratic lsq stage cost, no state bounds, no
// pinned goal components, no final-stage constraint.
#pragma once

#include <math_constants.h>

#include "launch.h"
#include "lm_device.cuh"

namespace b200sqp {

enum { PF_ACTIVE = 1, PF_LIN = 2, PF_STOP = 4 };

template <class M>
struct PipeDim
{
    static constexpr int NX = M::NX, NU = M::NU, NB = NU + NX, NV = 2 * NX + NU;
    static constexpr int ND = NB * (NB + 1) / 2, NE = NB * NX, NXX = NX * (NX + 1) / 2;
    static_assert(NV <= 31, "one lane per Jacobian column plus one lane for the unperturbed values");
    // storage stride of the A^T A triangles: padded to a multiple of 16 bytes for the bulk copies (lm_device_types.h paddedTriangle)
    template <class Real>
    struct Padded
    {
        static constexpr int per16 = 16 / (int)sizeof(Real);
        static constexpr int nxxp  = (NXX + per16 - 1) / per16 * per16;
    };
};

// finite-difference step of the Jacobian columns: the reference's 1e-9 in fp64; 2^-8 in fp32, near the optimum (3 eps)^(1/3) of a central
// difference (rounding ~ 6e-8 / 2^-8 = 1.5e-5 relative, truncation ~ delta^2 / 6 = 2.5e-6): the Jacobian carries ~2e-5 relative error
template <class Real>
struct FdStep;
template <>
struct FdStep<double>
{
    static constexpr double delta = 1e-9;
};
template <>
struct FdStep<float>
{
    static constexpr float delta = 0.00390625f;
};

// x / dt in the precision of the pipeline: the correctly rounded sequence of dynamics.cuh in fp64, a plain multiplication by the reciprocal in fp32
template <class Real>
struct StepSizeT;
template <>
struct StepSizeT<double> : public StepSize
{
    __device__ __forceinline__ explicit StepSizeT(double t) : StepSize(t) {}
};
template <>
struct StepSizeT<float>
{
    float dt, rcp;
    __device__ __forceinline__ explicit StepSizeT(double t) : dt((float)t), rcp((float)(1.0 / t)) {}
    __device__ __forceinline__ float div(float x) const { return x * rcp; }
};

__device__ __forceinline__ float pivotRsqrtT(float d)
{
    const float y = rsqrtf(d);
    return y * (1.5f - 0.5f * d * y * y);  // one Newton step on the hardware approximation
}
__device__ __forceinline__ double pivotRsqrtT(double d)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));  // BlockSolver::pivotRsqrt (lm_device.cuh)
    const double e = fma(-d * y, y, 1.0);
    const double p = fma(e, 0.375, 0.5);
    return fma(y * e, p, y);
}
__device__ __forceinline__ float fmaT(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ double fmaT(double a, double b, double c) { return fma(a, b, c); }

__device__ __forceinline__ double driftRoundTrip(double v)
{
    // what BaseEdge::computeJacobian leaves in a vertex component: +delta, -2 delta, +delta (edge_interface.cpp:78-85)
    v += 1e-9;
    v += -2e-9;
    v += 1e-9;
    return v;
}
__device__ __forceinline__ float driftRoundTrip(float v) { return v; }  // fp32 variant: no bit-parity to keep

// ---- TMA bulk copies (cp.async.bulk, 1-D) with mbarrier completion: the factor kernel prefetches the next Hessian block from HBM
//      into shared memory while the current one is eliminated (SASS: UBLKCP / SYNCS)
__device__ __forceinline__ unsigned smemAddr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbarInit(unsigned long long* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbarExpectTx(unsigned long long* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulkLoad(void* dst_smem, const void* src_global, unsigned bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smemAddr(dst_smem)),
                 "l"(src_global), "r"(bytes), "r"(smemAddr(bar))
                 : "memory");
}
// bounded wait (a lost copy must not hang the GPU): returns false after ~1e7 polls
__device__ __forceinline__ bool mbarWait(unsigned long long* bar, unsigned parity)
{
    const unsigned addr = smemAddr(bar);
    for (int spin = 0; spin < 10000000; ++spin)
    {
        unsigned ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok)
                     : "r"(addr), "r"(parity)
                     : "memory");
        if (ok) return true;
    }
    return false;
}

// fp64 tensor-core tile: D(8x8) += A(8x4) B(4x8); per lane a = A[lane/4][lane%4], b = B[lane%4][lane/4], d0/d1 = D[lane/4][2(lane%4) + {0,1}]
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// models that expose trig()/fTrig() (dynamics.cuh HasTrig): their angles' sines and cosines can be carried between evaluations
template <class M>
using PipeTrig = HasTrig<M>;

__device__ __forceinline__ size_t tiledSlot(int i, int slot, int nslots) { return ((size_t)(i >> 5) * nslots + slot) * 32 + (i & 31); }

// ---------------------------------------------------------------------------------------------------------------------------
// linearise: one warp per (instance, interval)
// ---------------------------------------------------------------------------------------------------------------------------
template <class M, int DEFECT, class Real>
__global__ void __launch_bounds__(128, 4) pipeLinearizeKernel(const __grid_constant__ DeviceOcp P, const __grid_constant__ DeviceState st,
                                                           const __grid_constant__ PipeArraysT<Real> pa)
{
    using Pd = PipeDim<M>;
    constexpr int NX = Pd::NX, NU = Pd::NU, NB = Pd::NB, NV = Pd::NV, ND = Pd::ND, NE = Pd::NE, NXXP = Pd::template Padded<Real>::nxxp;
    constexpr bool F64 = sizeof(Real) == 8;
    constexpr Real delta = FdStep<Real>::delta, neg2delta = -2 * delta, scalar = Real(1) / (2 * delta);
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int K    = P.K;
    const long long warp = (long long)blockIdx.x * 4 + wib;
    const int i = (int)(warp / K), k = (int)(warp % K);
    // Jacobian columns of the dynamics edge, [column][row]; fp64: padded by one (conflict-free column writes, scalar fragment reads);
    // fp32: rows of NX floats = whole 16-byte vectors (the Gram products read them as float4)
    constexpr int GS = F64 ? NX + 1 : NX;
    __shared__ __align__(16) Real sG[4][NV][GS];
    __shared__ __align__(16) Real sE0[4][NX];     // weighted defect at the unperturbed point
    __shared__ Real sCost[4][4][NB > NX ? NB : NX];  // 0: cost value, 1: cost Jacobian (diagonal), 2: bound value, 3: bound Jacobian, per block slot
    __shared__ unsigned char sTriR[ND], sTriC[ND];
    for (int idx = threadIdx.x; idx < ND; idx += blockDim.x)
    {
        int r = 0;
        while ((r + 1) * (r + 2) / 2 <= idx) ++r;
        sTriR[idx] = (unsigned char)r;
        sTriC[idx] = (unsigned char)(idx - r * (r + 1) / 2);
    }
    __syncthreads();
    if (i >= P.B || !(pa.flags[i] & PF_LIN)) return;

    const Real w_eq = (Real)st.w_eq, w_b = (Real)st.w_b;
    const double* z = st.z[st.cur[i]];
    const int slots = K * NB;
    const bool last = (k == K - 1);
    const bool has_xs = last ? (P.final_cost != 0) : true;
    const double* xs_w = last ? P.qf_sqrt : P.q_sqrt;

    // ---- operands: lane q < NV loads component q of [x_k | u_k | x_{k+1}], lane q < NX also the reference state; all-to-all by shuffles
    Real mine = 0, mref = 0;
    if (lane < NX)
    {
        mine = (Real)((k > 0) ? z[tiledSlot(i, (k - 1) * NB + NU + lane, slots)] : st.x0[tiledSlot(i, lane, NX)]);
        mref = (Real)st.xref[tiledSlot(i, lane, NX)];
    }
    else if (lane < NX + NU)
        mine = (Real)z[tiledSlot(i, k * NB + (lane - NX), slots)];
    else if (lane < NV)
        mine = (Real)z[tiledSlot(i, k * NB + NU + (lane - NX - NU), slots)];
    Real v[NV], xr[NX];
#pragma unroll
    for (int q = 0; q < NV; ++q) v[q] = __shfl_sync(0xffffffffu, mine, q);
#pragma unroll
    for (int q = 0; q < NX; ++q) xr[q] = __shfl_sync(0xffffffffu, mref, q);

    // ---- the vector this lane evaluates.  Lane p < NV owns column p of the dynamics edge (vertices in attachment order x_k, u_k,
    //      x_{k+1}); before that edge the lsq edges have perturbed u_k and x_{k+1} once (costs on them) and x_k twice (its own cost
    //      edge and the dynamics edge of interval k-1): computeCombinedSparseJacobian visits all lsq edges first (:1495-1525), then
    //      the equality edges in order (:1531-1559).  Components in front of p have completed this edge's round trip as well.
    //      Lanes >= NV evaluate the unperturbed point (the `values` the LM loop computed before the Jacobian).
    //      (fp32: driftRoundTrip is the identity -- there is no bit-parity to keep in that precision.)
    const int p = lane;
    const bool col_lane = p < NV && (k > 0 || p >= NX);  // x_0 is fixed: no columns for it
    Real vec[NV];
#pragma unroll
    for (int q = 0; q < NV; ++q)
    {
        Real b = v[q];
        if (p < NV)
        {
            if (q < NX)
            {
                if (k > 0) b = driftRoundTrip(driftRoundTrip(b));
            }
            else if (q < NX + NU)
                b = driftRoundTrip(b);
            else if (has_xs)
                b = driftRoundTrip(b);
            if (q < p && (k > 0 || q >= NX)) b = driftRoundTrip(b);
            if (q == p) b += delta;
        }
        vec[q] = b;
    }
    StepSizeT<Real> h(P.dt_ref);
    Real e2[NX], e1[NX];
#ifdef B200SQP_FP32_WHOLE_DEFECTS  // development switch: the fp32 variant with two whole defects per lane, as in fp64
    constexpr bool PAIRS = false;
#else
    constexpr bool PAIRS = !F64;
#endif
    if constexpr (PAIRS)
    {
        // Reduced precision: no bit-parity to keep, so the unit of work is one distinct PAIR of dynamics evaluations per lane instead of
        // two whole defects (four evaluations).  A column of x_k only moves f(x_k, u_k), a column of x_{k+1} only f(x_{k+1}, u_k); the
        // NU control columns move both and are split over two lanes: lane NX + m differences f(x_k, .), lane NV + m differences
        // f(x_{k+1}, .) and hands its part over by a shuffle.  With NV + NU = 32 every lane has exactly one pair.  The model is affine in u,
        // so the unperturbed values the LM loop needs are the MEANS of a control column's pair: no extra evaluation for them.
        static_assert(DEFECT == DEFECT_CRANK_NICOLSON && PipeTrig<M>::value && M::CONTROL_AFFINE && NV + NU <= 32,
                      "the fp32 variant exists for the trig-cached Crank-Nicolson path of a control-affine model");
        constexpr int NA = M::NANG, A0 = M::ANG0;
        const bool second = lane >= NX + NU;                                 // this lane works on f(x_{k+1}, u_k)
        const bool ucol   = (lane >= NX && lane < NX + NU) || lane >= NV;    // ... on a control column
        const int comp    = ucol ? (lane < NV ? lane - NX : lane - NV) : (second ? lane - NX - NU : lane);
        // lane-dependent choices are selects, not branches: the warp must not serialise over them
        const int xsel = ucol ? -1 : comp, usel = ucol ? comp : -1;
        Real xa[NX], ua[NU];
#pragma unroll
        for (int j = 0; j < NX; ++j) xa[j] = (second ? v[NX + NU + j] : v[j]) + (j == xsel ? delta : Real(0));
#pragma unroll
        for (int j = 0; j < NU; ++j) ua[j] = v[NX + j] + (j == usel ? delta : Real(0));
        Real sc[2 * NA], fp[NX], fm[NX];
        M::trig(xa, sc);
        M::fTrig(P.dyn, xa, ua, sc, fp);
        Real ang = 0;
#pragma unroll
        for (int q = 0; q < NX; ++q)
        {
            xa[q] += (q == xsel ? neg2delta : Real(0));
            ang = (q == xsel) ? xa[q] : ang;
        }
#pragma unroll
        for (int q = 0; q < NU; ++q) ua[q] += (q == usel ? neg2delta : Real(0));
        Real sa, ca;
        sincosT(ang, &sa, &ca);  // the one sine / cosine pair the perturbation may have touched
#pragma unroll
        for (int a = 0; a < NA; ++a)
        {
            sc[2 * a]     = (xsel == A0 + a) ? sa : sc[2 * a];
            sc[2 * a + 1] = (xsel == A0 + a) ? ca : sc[2 * a + 1];
        }
        M::fTrig(P.dyn, xa, ua, sc, fm);
        // the (x_{k+1} - x_k) / dt part of the defect moves only in row `comp` of a state column: own component +-delta against the
        // same component of the other state (lane +- (NX + NU))
        const Real other = __shfl_sync(0xffffffffu, mine, second ? lane - NX - NU : lane + NX + NU);
        const Real a1 = mine + delta, b1 = a1 + neg2delta;
        const Real dx_self = second ? h.div(a1 - other) - h.div(b1 - other) : h.div(other - a1) - h.div(other - b1);
        const bool ufirst = lane >= NX && lane < NX + NU;
#pragma unroll
        for (int j = 0; j < NX; ++j)
        {
            Real d          = fp[j] - fm[j];
            const Real mean = Real(0.5) * (fp[j] + fm[j]);
            const Real d2   = __shfl_down_sync(0xffffffffu, d, NV - NX);  // lane NX + m <- lane NV + m: the f(x_{k+1}, .) half of a control column
            const Real f2b  = __shfl_sync(0xffffffffu, mean, NV);         // f(x_{k+1}, u_k) from the first helper lane
            d += ufirst ? d2 : Real(0);
            if (lane == NX) sE0[wib][j] = (h.div(v[NX + NU + j] - v[j]) - Real(0.5) * (mean + f2b)) * w_eq;  // levenberg_marquardt_sparse.cpp:231-235
            const Real dx = (j == xsel) ? dx_self : Real(0);
            if (lane < NV) sG[wib][lane][j] = col_lane ? scalar * (dx - Real(0.5) * d) * w_eq : Real(0);
        }
    }
    else if constexpr (DEFECT == DEFECT_CRANK_NICOLSON && PipeTrig<M>::value)
    {
        // The model's angles enter only through sin/cos and a column perturbs at most one of them: keep the sines and cosines of
        // both states from the +delta evaluation and refresh the one pair the perturbation touched for the -delta evaluation (one
        // uniform sincos per lane instead of six; same arguments -> same values as evaluating f from scratch).
        constexpr int NA = M::NANG, A0 = M::ANG0;
        Real sc1[2 * NA], sc2[2 * NA];
        M::trig(vec, sc1);
        M::trig(vec + NX + NU, sc2);
        auto cn = [&](Real* e) {
            Real f1[NX], f2[NX];
            M::fTrig(P.dyn, vec, vec + NX, sc1, f1);
            M::fTrig(P.dyn, vec + NX + NU, vec + NX, sc2, f2);
#pragma unroll
            for (int j = 0; j < NX; ++j) e[j] = h.div(vec[NX + NU + j] - vec[j]) - Real(0.5) * (f1[j] + f2[j]);  // dynamics.cuh defect<>, Crank-Nicolson
        };
        cn(e2);
        Real ang = 0;
#pragma unroll
        for (int q = 0; q < NV; ++q)
            if (q == p)
            {
                vec[q] += neg2delta;
                ang = vec[q];
            }
        Real sa, ca;
        sincosT(ang, &sa, &ca);
#pragma unroll
        for (int a = 0; a < NA; ++a)
        {
            if (p == A0 + a)
            {
                sc1[2 * a]     = sa;
                sc1[2 * a + 1] = ca;
            }
            if (p == NX + NU + A0 + a)
            {
                sc2[2 * a]     = sa;
                sc2[2 * a + 1] = ca;
            }
        }
        cn(e1);
    }
    else
    {
        static_assert(F64 || (DEFECT == DEFECT_CRANK_NICOLSON && PipeTrig<M>::value), "the fp32 variant exists for the trig-cached Crank-Nicolson path");
        if constexpr (F64)
        {
            defect<M, DEFECT>(P.dyn, vec, vec + NX, vec + NX + NU, h, e2);  // inlined: the operands stay in registers
#pragma unroll
            for (int q = 0; q < NV; ++q)
                if (q == p) vec[q] += neg2delta;
            defect<M, DEFECT>(P.dyn, vec, vec + NX, vec + NX + NU, h, e1);
        }
    }
    if constexpr (!PAIRS)
    {
        if (p < NV)
        {
#pragma unroll
            for (int j = 0; j < NX; ++j) sG[wib][p][j] = col_lane ? scalar * (e2[j] - e1[j]) * w_eq : Real(0);
        }
        else if (p == NV)
        {
#pragma unroll
            for (int j = 0; j < NX; ++j) sE0[wib][j] = e2[j] * w_eq;  // levenberg_marquardt_sparse.cpp:231-235
        }
    }

    // ---- lsq cost and bound rows of the block slots [u_k | x_{k+1}] (diagonal): lane s < NB handles slot s
    double cpart = 0.0;  // residual norms are accumulated in fp64 in both variants (the gain ratio differences them)
    if (lane < NB)
    {
        const int s = lane;
        Real cv, cj, bv = 0, bj = 0;
        if (s < NU)
        {
            // QuadraticFormCost::computeNonIntegralControlTerm, lsq + diagonal (quadratic_cost.cpp:146-154), FD like any edge
            const Real rs = (Real)P.r_sqrt[s];
            Real u        = v[NX + s];
            cv            = rs * u;
            u += delta;
            const Real v2 = rs * u;
            u += neg2delta;
            const Real v1 = rs * u;
            cj            = scalar * (v2 - v1);
            if (P.u_bounded[s])
            {
                bv = (Real)(boundDist((double)v[NX + s], P.u_lb[s], P.u_ub[s])) * w_b;
                bj = (Real)boundJac((double)driftRoundTrip(driftRoundTrip(v[NX + s])), P.u_lb[s], P.u_ub[s], (double)w_b);  // bounds rows come last (:1721-1752)
            }
        }
        else
        {
            const int j = s - NU;
            Real x      = v[NX + NU + j];
            cv = cj = 0;
            if (has_xs)
            {
                const Real ws = (Real)xs_w[j];
                cv = ws * (x - xr[j]);  // quadratic_cost.cpp:105-123 / final_state_cost.cpp:73-90
                x += delta;
                const Real v2 = ws * (x - xr[j]);
                x += neg2delta;
                const Real v1 = ws * (x - xr[j]);
                cj            = scalar * (v2 - v1);
            }
        }
        sCost[wib][0][s] = cv;
        sCost[wib][1][s] = cj;
        sCost[wib][2][s] = bv;
        sCost[wib][3][s] = bj;
        cpart            = fma((double)cv, (double)cv, (double)bv * (double)bv);
    }
    if (k == 0 && lane < NX)
    {
        const double c = (double)((Real)P.q_sqrt[lane] * (v[lane] - xr[lane]));  // cost edge on the fixed start state: value only
        cpart          = fma(c, c, cpart);
    }
    __syncwarp();
    if (lane < NX) cpart = fma((double)sE0[wib][lane], (double)sE0[wib][lane], cpart);
    // residual norm of this interval: fixed-order warp-shuffle reduction
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) cpart += __shfl_down_sync(0xffffffffu, cpart, off);
    if (lane == 0) pa.cpart[(size_t)k * P.S + i] = cpart;

    // ---- normal equations: the lanes share out the entries.  Block k = G^T G over the columns of [u_k | x_{k+1}] (+ diagonal cost /
    //      bound rows), E_k = G^T A, and A^T A / A^T e go to block k-1 (stored separately: the factorisation adds them).
    const size_t blk = (size_t)i * K + k;
    Real* Dk  = pa.D + blk * ND;
    Real* gk  = pa.g + blk * NB;
    static_assert(NX % 4 == 0 && NB <= 16 && NX <= 16, "8x8 output tiles over a 16-column window, k in steps of 4 / float4 rows");
    if constexpr (F64)
    {
        // The three Gram products -- G_b G_b^T (nb x nb), G_b G_a^T (nb x nx), G_a G_a^T (nx x nx), inner dimension nx -- run on the
        // tensor cores: fp64 mma.sync m8n8k4 tiles (SASS: DMMA) with fragments read straight from the column store in shared memory.
        // Rows >= nx of a "G_a" tile are columns of G_b (the store is contiguous): those products are computed and dropped.
        const int fm = lane >> 2, fk = lane & 3;  // fragment coordinates: A(m = fm, k = fk), B(k = fk, n = fm), C(m = fm, n = 2 fk + {0,1})
        auto gram = [&](int row_base, int r0, int col_base, int c0, double& d0, double& d1) {
            d0 = 0.0;
            d1 = 0.0;
#pragma unroll
            for (int k0 = 0; k0 < NX; k0 += 4)
            {
                const double af = sG[wib][row_base + r0 + fm][k0 + fk];
                const double bf = sG[wib][col_base + c0 + fm][k0 + fk];
                dmma884(d0, d1, af, bf);
            }
        };
        // block k: G_b G_b^T, lower triangle, + the diagonal cost / bound rows
#pragma unroll
        for (int t = 0; t < 3; ++t)
        {
            const int r0 = (t == 0) ? 0 : 8, c0 = (t == 2) ? 8 : 0;
            if (r0 + 8 > NB && r0 >= NB) continue;
            double d0, d1;
            gram(NX, r0, NX, c0, d0, d1);
            const int r = r0 + fm, c = c0 + 2 * fk;
            if (r < NB)
            {
                if (c <= r)
                {
                    if (r == c) d0 = fma(sCost[wib][1][r], sCost[wib][1][r], fma(sCost[wib][3][r], sCost[wib][3][r], d0));
                    Dk[tri(r, c)] = d0;
                }
                if (c + 1 <= r)
                {
                    if (r == c + 1) d1 = fma(sCost[wib][1][r], sCost[wib][1][r], fma(sCost[wib][3][r], sCost[wib][3][r], d1));
                    Dk[tri(r, c + 1)] = d1;
                }
            }
        }
        if (k > 0)
        {
            Real* Ek  = pa.E + blk * NE;
            Real* DAk = pa.DA + blk * NXXP;
            // E_k = G_b G_a^T: rows = slots of block k, columns = x-part of block k-1
#pragma unroll
            for (int t = 0; t < 4; ++t)
            {
                const int r0 = (t & 1) * 8, c0 = (t >> 1) * 8;
                double d0, d1;
                gram(NX, r0, 0, c0, d0, d1);
                const int r = r0 + fm, c = c0 + 2 * fk;
                if (r < NB && c < NX) Ek[r * NX + c] = d0;
                if (r < NB && c + 1 < NX) Ek[r * NX + c + 1] = d1;
            }
            // A^T A of this interval (belongs to block k-1): G_a G_a^T, lower triangle
#pragma unroll
            for (int t = 0; t < 3; ++t)
            {
                const int r0 = (t == 0) ? 0 : 8, c0 = (t == 2) ? 8 : 0;
                double d0, d1;
                gram(0, r0, 0, c0, d0, d1);
                const int r = r0 + fm, c = c0 + 2 * fk;
                if (r < NX && c <= r) DAk[tri(r, c)] = d0;
                if (r < NX && c + 1 <= r) DAk[tri(r, c + 1)] = d1;
            }
        }
    }
    else
    {
        // fp32: the same three Gram products on the FMA pipe; a lane owns every 32nd entry and reads the two columns as float4 rows
        auto dot = [&](int ra, int rb) {
            float s = 0.f;
#pragma unroll
            for (int q = 0; q < NX; q += 4)
            {
                const float4 a = *reinterpret_cast<const float4*>(&sG[wib][ra][q]);
                const float4 b = *reinterpret_cast<const float4*>(&sG[wib][rb][q]);
                s = fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, fmaf(a.w, b.w, s))));
            }
            return s;
        };
        for (int idx = lane; idx < ND; idx += 32)
        {
            const int r = sTriR[idx], c = sTriC[idx];
            float s = dot(NX + r, NX + c);
            if (r == c) s = fmaf(sCost[wib][1][r], sCost[wib][1][r], fmaf(sCost[wib][3][r], sCost[wib][3][r], s));
            Dk[idx] = s;
        }
        if (k > 0)
        {
            Real* Ek  = pa.E + blk * NE;
            Real* DAk = pa.DA + blk * NXXP;
            for (int idx = lane; idx < NE; idx += 32) Ek[idx] = dot(NX + idx / NX, idx % NX);
            for (int idx = lane; idx < Pd::NXX; idx += 32) DAk[idx] = dot(sTriR[idx], sTriC[idx]);
        }
    }
    if (lane < NB)
    {
        Real s = 0;
#pragma unroll
        for (int q = 0; q < NX; ++q) s = fmaT(sG[wib][NX + lane][q], sE0[wib][q], s);
        s        = fmaT(sCost[wib][1][lane], sCost[wib][0][lane], fmaT(sCost[wib][3][lane], sCost[wib][2][lane], s));
        gk[lane] = -s;
    }
    if (k > 0 && lane < NX)
    {
        Real* gAk = pa.gA + blk * NX;
        Real s    = 0;
#pragma unroll
        for (int q = 0; q < NX; ++q) s = fmaT(sG[wib][lane][q], sE0[wib][q], s);
        gAk[lane] = -s;
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// initial LM state after the first linearisation: one warp per instance (levenberg_marquardt_sparse.cpp:103-126)
// ---------------------------------------------------------------------------------------------------------------------------
template <class M, class Real>
__global__ void __launch_bounds__(128) pipeInitKernel(const __grid_constant__ DeviceOcp P, const __grid_constant__ DeviceState st,
                                                      const __grid_constant__ PipeArraysT<Real> pa, int iterations)
{
    using Pd = PipeDim<M>;
    constexpr int NX = Pd::NX, NU = Pd::NU, NB = Pd::NB, ND = Pd::ND, NXXP = Pd::template Padded<Real>::nxxp;
    const int lane = threadIdx.x & 31;
    const int i    = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (i >= P.B) return;
    const int K = P.K;
    double ginf = 0.0, maxdiag = -CUDART_INF, chi2 = 0.0;
    for (int k = lane; k < K; k += 32)
    {
        const size_t blk = (size_t)i * K + k;
        for (int r = 0; r < NB; ++r)
        {
            double d = (double)pa.D[blk * ND + tri(r, r)], g = (double)pa.g[blk * NB + r];
            if (r >= NU && k + 1 < K)
            {
                d += (double)pa.DA[(blk + 1) * NXXP + tri(r - NU, r - NU)];
                g += (double)pa.gA[(blk + 1) * NX + (r - NU)];
            }
            maxdiag = fmax(maxdiag, d);
            ginf    = fmax(ginf, fabs(g));
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1)
    {
        ginf    = fmax(ginf, __shfl_down_sync(0xffffffffu, ginf, off));
        maxdiag = fmax(maxdiag, __shfl_down_sync(0xffffffffu, maxdiag, off));
    }
    if (lane == 0)
    {
        for (int k = 0; k < K; ++k) chi2 += pa.cpart[(size_t)k * P.S + i];  // fixed order
        constexpr double eps1 = 1e-5, tau = 1e-5;
        double mu = tau * maxdiag;
        if (mu < 0) mu = 0;
        const bool active  = iterations > 0;
        st.chi2[i]         = chi2;
        st.mu[i]           = mu;
        st.rho[i]          = 0.0;
        pa.mu_acc[i]       = mu;
        pa.last_values[i]  = chi2;
        pa.v[i]            = 2u;
        pa.k_outer[i]      = 0;
        pa.flags[i]        = (active ? PF_ACTIVE : 0) | (ginf <= eps1 ? PF_STOP : 0);
        st.n_factor[i]     = 0;
        st.n_reject[i]     = 0;
        st.n_linearize[i]  = 1;
        st.status[i]       = (ginf <= eps1) ? B200SQP_STATUS_CONVERGED : B200SQP_STATUS_EARLY_TERMINATED;
        if (st.trace) st.trace[i] = chi2;
        if (active) atomicOr(pa.any, 1);
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// (H + mu_acc I) delta = g: block-tridiagonal Cholesky, one warp per instance, lane = row of the current block
// ---------------------------------------------------------------------------------------------------------------------------
// IPW = instances per warp: a block row needs NB <= 16 lanes, so in fp32 (half the registers and shared memory per instance) two
// instances share a warp, one per half-warp (shuffles of width 16); fp64 keeps one instance per warp and the tensor-core Schur update.
template <class M, class Real, int IPW>
__global__ void __launch_bounds__(128, IPW == 1 ? 7 : 6) pipeFactorKernel(const __grid_constant__ DeviceOcp P, const __grid_constant__ DeviceState st,
                                                                        const __grid_constant__ PipeArraysT<Real> pa)
{
    using Pd = PipeDim<M>;
    constexpr int NX = Pd::NX, NU = Pd::NU, NB = Pd::NB, ND = Pd::ND, NE = Pd::NE, NXXP = Pd::template Padded<Real>::nxxp;
    constexpr bool F64 = sizeof(Real) == 8;
    constexpr int RB   = (int)sizeof(Real);
    constexpr int HL   = 32 / IPW;  // lanes per instance
    constexpr unsigned FULL = 0xffffffffu;
    static_assert((ND * RB) % 16 == 0 && (NE * RB) % 16 == 0 && (NXXP * RB) % 16 == 0, "bulk copies move multiples of 16 bytes");
    static_assert(NB <= HL && (IPW == 1 || !F64), "one lane per block row; the fp64 tensor-core tiles span the whole warp");
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int sub = lane / HL, hl = lane % HL;  // instance of this warp, lane within the instance
    const int i_raw = (blockIdx.x * 4 + wib) * IPW + sub;
    // Lane r < NB of an instance owns ROW r of the current block in registers.  The blocks arrive from HBM by TMA bulk copies into a
    // two-stage ring per instance (the next block is in flight while the current one is eliminated); shared memory also holds what
    // other lanes must read (W_k rows, the trailing block of the previous factor) and stages the coalesced stores.
    __shared__ __align__(16) Real stD[4][IPW][2][ND];    // forward: D_k        backward: L_k
    __shared__ __align__(16) Real stE[4][IPW][2][NE];    // forward: E_k        backward: W_{k+1}
    __shared__ __align__(16) Real stA[4][IPW][2][NXXP];  // forward: A^T A of interval k+1
    __shared__ __align__(8) unsigned long long bars[4][IPW][2];
    __shared__ Real sLxx[4][IPW][NX][NX + 1];  // trailing nx x nx block of the previous factor (reciprocal diagonal)
    // W_k overwrites E_k and L_k overwrites D_k in their stage (same packed layouts) once every lane holds its row in registers:
    // the coalesced stores then run straight out of the stage
    if (hl == 0)
    {
        mbarInit(&bars[wib][sub][0], 1);
        mbarInit(&bars[wib][sub][1], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    // an instance that is out of range or not active this pass still walks through the code with its half-warp (the warp's control
    // flow is shared) on the data of a valid instance, but writes nothing
    const bool valid = i_raw < P.B && (pa.flags[i_raw < P.B ? i_raw : 0] & PF_ACTIVE);
    if (!__any_sync(FULL, valid)) return;
    const int i      = i_raw < P.B ? i_raw : P.B - 1;
    const int K      = P.K;
    const Real mua   = (Real)pa.mu_acc[i];
    const double mu  = st.mu[i];
    const bool rowl  = hl < NB;
    Real* sD[2] = {stD[wib][sub][0], stD[wib][sub][1]};
    Real* sE[2] = {stE[wib][sub][0], stE[wib][sub][1]};
    Real* sA[2] = {stA[wib][sub][0], stA[wib][sub][1]};
    unsigned long long* bar[2] = {&bars[wib][sub][0], &bars[wib][sub][1]};
    unsigned phase[2] = {0u, 0u};
    bool ok           = true;
    auto issueForward = [&](int k) {
        if (hl == 0)
        {
            const size_t blk = (size_t)i * K + k;
            const int sidx   = k & 1;
            const unsigned bytes = ND * RB + (k > 0 ? NE * RB : 0) + (k + 1 < K ? NXXP * RB : 0);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // our generic-proxy writes to this stage come first
            mbarExpectTx(bar[sidx], bytes);
            bulkLoad(sD[sidx], pa.D + blk * ND, ND * RB, bar[sidx]);
            if (k > 0) bulkLoad(sE[sidx], pa.E + blk * NE, NE * RB, bar[sidx]);
            if (k + 1 < K) bulkLoad(sA[sidx], pa.DA + (blk + 1) * NXXP, NXXP * RB, bar[sidx]);
        }
    };
    issueForward(0);
    Real yp_x = 0;  // lane a < NX: x-part of the forward-substituted rhs of the previous block
    for (int k = 0; k < K; ++k)
    {
        const size_t blk = (size_t)i * K + k;
        const int sidx   = k & 1;
        if (k + 1 < K) issueForward(k + 1);  // its stage was last read two iterations ago (warp-synchronised since)
        Real y = 0;                           // lane r < NB holds rhs component r
        if (rowl)
        {
            y = pa.g[blk * NB + hl];
            if (k + 1 < K && hl >= NU) y += pa.gA[(blk + 1) * NX + (hl - NU)];
        }
        ok = mbarWait(bar[sidx], phase[sidx]) && ok;
        phase[sidx] ^= 1u;
        Real* cD = sD[sidx];
        Real* cE = sE[sidx];
        if (k > 0)
        {
            // W_k Lxx^T = E_k: row r per lane
            Real wr[NX];
#pragma unroll
            for (int a = 0; a < NX; ++a)
            {
                Real s = rowl ? cE[hl * NX + a] : Real(0);
#pragma unroll
                for (int b = 0; b < a; ++b) s = fmaT(-wr[b], sLxx[wib][sub][a][b], s);
                wr[a] = s * sLxx[wib][sub][a][a];
            }
            __syncwarp();  // every lane has read its E row
            if (rowl)
            {
#pragma unroll
                for (int a = 0; a < NX; ++a) cE[hl * NX + a] = wr[a];
            }
            __syncwarp();
            if (valid)
                for (int idx = hl; idx < NE; idx += HL) pa.W[blk * NE + idx] = cE[idx];
            if constexpr (F64)
            {
                // Schur complement S_k -= W_k W_k^T on the staged block with fp64 tensor-core tiles (DMMA m8n8k4): lower triangle = the
                // 8x8 tiles (0,0), (1,0), (1,1); fragments straight from the stage (W packed [r][a], S packed lower by rows)
                const int fm = lane >> 2, fk = lane & 3;
#pragma unroll
                for (int t = 0; t < 3; ++t)
                {
                    const int r0 = (t == 0) ? 0 : 8, c0 = (t == 2) ? 8 : 0;
                    const int r = r0 + fm, c = c0 + 2 * fk;
                    const bool in0 = r < NB && c <= r, in1 = r < NB && c + 1 <= r;
                    double d0 = in0 ? cD[tri(r, c)] : 0.0;
                    double d1 = in1 ? cD[tri(r, c + 1)] : 0.0;
#pragma unroll
                    for (int k0 = 0; k0 < NX; k0 += 4)
                    {
                        const double af = (r0 + fm < NB) ? -cE[(r0 + fm) * NX + k0 + fk] : 0.0;
                        const double bf = (c0 + fm < NB) ? cE[(c0 + fm) * NX + k0 + fk] : 0.0;
                        dmma884(d0, d1, af, bf);
                    }
                    if (in0) cD[tri(r, c)] = d0;
                    if (in1) cD[tri(r, c + 1)] = d1;
                }
            }
            else
            {
                // fp32: S_k -= W_k W_k^T on the FMA pipe: lane r updates its own row; the rows of W are read from the stage as float4
                // (all lanes of an instance read the same W_c: a broadcast)
                if (rowl)
                {
#pragma unroll
                    for (int c = 0; c < NB; ++c)
                    {
                        if (c <= hl)
                        {
                            float s = cD[tri(hl, c)];
#pragma unroll
                            for (int q = 0; q < NX; q += 4)
                            {
                                const float4 wc = *reinterpret_cast<const float4*>(&cE[c * NX + q]);
                                s = fmaf(-wr[q], wc.x, fmaf(-wr[q + 1], wc.y, fmaf(-wr[q + 2], wc.z, fmaf(-wr[q + 3], wc.w, s))));
                            }
                            cD[tri(hl, c)] = s;
                        }
                    }
                }
            }
            // rhs update
#pragma unroll
            for (int a = 0; a < NX; ++a) y = fmaT(-wr[a], __shfl_sync(FULL, yp_x, a, HL), y);
            __syncwarp();
        }
        // ---- row r of S_k = D_k - W_k W_k^T (+ A^T A of interval k+1 on the x-x part, + damping)
        Real row[NB];  // row[c], c <= hl
#pragma unroll
        for (int c = 0; c < NB; ++c)
        {
            Real v = 0;
            if (rowl && c <= hl)
            {
                v = cD[tri(hl, c)];
                if (k + 1 < K && c >= NU) v += sA[sidx][tri(hl - NU, c - NU)];
                if (c == hl) v += mua;
            }
            row[c] = v;
        }
        // ---- Cholesky of S_k in registers (right-looking; column j of L is spread over the lanes, broadcast by shuffles) fused with
        //      the forward substitution of the rhs
#pragma unroll
        for (int j = 0; j < NB; ++j)
        {
            const Real inv = pivotRsqrtT(__shfl_sync(FULL, row[j], j, HL));
            const Real yj  = __shfl_sync(FULL, y, j, HL) * inv;
            const Real l   = (hl > j) ? row[j] * inv : Real(0);
            if (hl > j) y = fmaT(-l, yj, y);
            if (hl == j) y = yj;
            row[j] = (hl == j) ? inv : l;
#pragma unroll
            for (int c = j + 1; c < NB; ++c)
            {
                const Real lc = __shfl_sync(FULL, l, c, HL);
                row[c]        = fmaT(-l, lc, row[c]);  // lanes < c hold zeros there and l = 0 for lanes <= j
            }
        }
        // ---- store the factor (coalesced, packed order) and the forward-substituted rhs, keep what the next block needs
        __syncwarp();
        if (rowl)
        {
#pragma unroll
            for (int c = 0; c < NB; ++c)
                if (c <= hl) cD[tri(hl, c)] = row[c];
            if (valid) pa.y[blk * NB + hl] = y;
            if (hl >= NU)
            {
#pragma unroll
                for (int c = NU; c < NB; ++c)
                    if (c <= hl) sLxx[wib][sub][hl - NU][c - NU] = row[c];
            }
        }
        yp_x = __shfl_sync(FULL, y, NU + (hl < NX ? hl : 0), HL);
        __syncwarp();
        if (valid)
            for (int idx = hl; idx < ND; idx += HL) pa.L[blk * ND + idx] = cD[idx];
        __syncwarp();
    }
    // ---- back-substitution, bottom-up: L_k^T delta_k = y_k - W_{k+1}^T delta_{k+1} (x-part).  The factor blocks were written through
    //      the generic proxy above and are read back by the async proxy: fence in between.  (A half-warp that writes nothing reads the
    //      factor an earlier pass left for its instance: finite data, results discarded.)
    __threadfence();
    asm volatile("fence.proxy.async;" ::: "memory");
    __syncwarp();
    auto issueBackward = [&](int k, int sidx) {
        if (hl == 0)
        {
            const size_t blk = (size_t)i * K + k;
            const unsigned bytes = ND * RB + (k + 1 < K ? NE * RB : 0);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbarExpectTx(bar[sidx], bytes);
            bulkLoad(sD[sidx], pa.L + blk * ND, ND * RB, bar[sidx]);
            if (k + 1 < K) bulkLoad(sE[sidx], pa.W + (blk + 1) * NE, NE * RB, bar[sidx]);
        }
    };
    issueBackward(K - 1, 0);
    double dn2 = 0.0, dq = 0.0;  // accumulated in fp64 in both variants
    Real dnext = 0;              // lane r holds delta_{k+1}[r]
    int step = 0;
    for (int k = K - 1; k >= 0; --k, ++step)
    {
        const size_t blk = (size_t)i * K + k;
        const int sidx   = step & 1;
        if (k > 0) issueBackward(k - 1, sidx ^ 1);
        Real d = 0, gfull = 0;
        if (rowl)
        {
            d     = pa.y[blk * NB + hl];
            gfull = pa.g[blk * NB + hl];
            if (k + 1 < K && hl >= NU) gfull += pa.gA[(blk + 1) * NX + (hl - NU)];
        }
        ok = mbarWait(bar[sidx], phase[sidx]) && ok;
        phase[sidx] ^= 1u;
        if (k + 1 < K)
        {
            Real s = 0;
            const int a = (hl >= NU && rowl) ? hl - NU : 0;
#pragma unroll
            for (int r = 0; r < NB; ++r) s = fmaT(sE[sidx][r * NX + a], __shfl_sync(FULL, dnext, r, HL), s);
            if (hl >= NU && rowl) d -= s;
        }
        // L^T delta = d, column-oriented backward: lane r needs L[j][r] for j >= r, i.e. column r of the factor
        Real colr[NB];
#pragma unroll
        for (int j = 0; j < NB; ++j) colr[j] = (rowl && j >= hl) ? sD[sidx][tri(j, hl)] : Real(0);
#pragma unroll
        for (int j = NB - 1; j >= 0; --j)
        {
            const Real dj = __shfl_sync(FULL, d, j, HL) * __shfl_sync(FULL, colr[j], j, HL);
            if (hl == j) d = dj;
            if (hl < j) d = fmaT(-colr[j], dj, d);
        }
        if (rowl)
        {
            const double dd = (double)d;
            if (valid) st.dl[tiledSlot(i, k * NB + hl, K * NB)] = dd;
            dn2 = fma(dd, dd, dn2);
            dq  = fma(dd, fma(mu, dd, (double)gfull), dq);
        }
        dnext = d;
        __syncwarp();
    }
#pragma unroll
    for (int off = HL / 2; off > 0; off >>= 1)
    {
        dn2 += __shfl_down_sync(FULL, dn2, off, HL);
        dq += __shfl_down_sync(FULL, dq, off, HL);
    }
    if (hl == 0 && valid)
    {
        // a lost bulk copy (or a non-finite factor) poisons the pass: the trial kernel then skips the instance and the control kernel
        // takes the reject branch (mu *= v), exactly as for a step that did not reduce chi2
        pa.dn2[i] = (ok && isfinite(dn2) && isfinite(dq)) ? dn2 : CUDART_NAN;
        pa.dq[i]  = dq;
        st.n_factor[i] += 1;
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// trial point and its residuals: one thread per (instance, interval)
// ---------------------------------------------------------------------------------------------------------------------------
template <class M, int DEFECT, class Real>
__global__ void __launch_bounds__(128) pipeTrialKernel(const __grid_constant__ DeviceOcp P, const __grid_constant__ DeviceState st,
                                                       const __grid_constant__ PipeArraysT<Real> pa)
{
    using Pd = PipeDim<M>;
    constexpr int NX = Pd::NX, NB = Pd::NB;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y;
    if (i >= P.B || !(pa.flags[i] & PF_ACTIVE)) return;
    constexpr double eps2 = 1e-5;
    const double dn2 = pa.dn2[i];
    if (isnan(dn2) || sqrt(dn2) <= eps2) return;  // failed factorisation, or step too small: no trial point (levenberg_marquardt_sparse.cpp:151-154)
    const Weights w{st.w_eq, st.w_ineq, st.w_b};
    const size_t tile = i >> 5, lane = i & 31;
    const size_t zoff = tile * ((size_t)P.K * NB * TILE) + lane;
    const int cur     = st.cur[i];
    const double part = trialChi2<M, DEFECT, 0, FeatLean>(P, w, st.z[cur] + zoff, st.dl + zoff, st.z[cur ^ 1] + zoff,
                                                           st.x0 + tile * ((size_t)NX * TILE) + lane, st.xref + tile * ((size_t)NX * TILE) + lane, nullptr, k, k + 1);
    pa.cpart[(size_t)k * P.S + i] = part;
}

// ---------------------------------------------------------------------------------------------------------------------------
// LM control: one thread per instance (levenberg_marquardt_sparse.cpp:151-216; same state machine as lmSolveKernel's C phase)
// ---------------------------------------------------------------------------------------------------------------------------
template <class Real>
__global__ void __launch_bounds__(128) pipeControlKernel(const __grid_constant__ DeviceOcp P, const __grid_constant__ DeviceState st,
                                                         const __grid_constant__ PipeArraysT<Real> pa, int iterations)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.B) return;
    int flags = pa.flags[i];
    if (!(flags & PF_ACTIVE)) return;
    constexpr double eps2 = 1e-5, eps3 = 1e-5, eps4 = 0;
    constexpr double goodStepUpperScale = 2. / 3., goodStepLowerScale = 1. / 3.;
    double mu = st.mu[i], mu_acc = pa.mu_acc[i], rho = st.rho[i], chi2_old = st.chi2[i], last_values = pa.last_values[i];
    unsigned v   = pa.v[i];
    int k_outer  = pa.k_outer[i];
    bool stop    = (flags & PF_STOP) != 0;
    bool lin     = false;
    const double dq = pa.dq[i];
    const bool factor_failed = isnan(pa.dn2[i]);
    if (factor_failed)
    {
        // the factorisation of this pass did not complete: reject like a step that increased chi2 (more damping, same linearisation)
        st.n_reject[i] += 1;
        rho = -1.0;
        mu  = mu * v;
        v   = 2 * v;
        if (v == 0) stop = true;
    }
    else if (sqrt(pa.dn2[i]) <= eps2)
        stop = true;
    else
    {
        double chi2_new = 0.0;
        for (int k = 0; k < P.K; ++k) chi2_new += pa.cpart[(size_t)k * P.S + i];  // fixed order
        last_values = chi2_new;
        rho         = (chi2_old - chi2_new) / dq;
        if (rho > 0 && !isnan(chi2_new) && !isinf(chi2_new))
        {
            stop = (sqrt(chi2_old) - sqrt(chi2_new) < eps4 * sqrt(chi2_old));
            st.cur[i] ^= 1;  // accept: the trial buffer becomes the current one
            if (!stop && k_outer < iterations - 1)
            {
                lin = true;
                st.n_linearize[i] += 1;
                mu_acc             = 0.0;
                const double c     = 2 * rho - 1;
                double alpha       = fmin(goodStepUpperScale, 1 - c * c * c);
                double scaleFactor = fmax(goodStepLowerScale, alpha);
                mu *= scaleFactor;
                v = 2;
            }
            chi2_old = chi2_new;
        }
        else
        {
            st.n_reject[i] += 1;
            mu = mu * v;
            v  = 2 * v;
            if (v == 0) stop = true;  // see lm_kernels.cuh: the reference's unsigned `v` wraps and its loop would never end
        }
    }
    bool active = true;
    if (!(rho <= 0 && !stop))
    {
        stop = (sqrt(last_values) <= eps3);  // :216
        ++k_outer;
        if (st.trace) st.trace[(size_t)k_outer * P.S + i] = chi2_old;
        active = k_outer < iterations;
    }
    if (active) mu_acc += mu;
    st.mu[i]          = mu;
    st.rho[i]         = rho;
    st.chi2[i]        = chi2_old;
    pa.mu_acc[i]      = mu_acc;
    pa.last_values[i] = last_values;
    pa.v[i]           = v;
    pa.k_outer[i]     = k_outer;
    pa.flags[i]       = (active ? PF_ACTIVE : 0) | (lin && active ? PF_LIN : 0) | (stop ? PF_STOP : 0);
    st.status[i]      = (stop || rho <= 0) ? B200SQP_STATUS_CONVERGED : B200SQP_STATUS_EARLY_TERMINATED;
    if (active) atomicOr(pa.any, 1);
    if (lin && active) atomicOr(pa.any + 1, 1);
}

__global__ void pipeSetFlagsKernel(int* flags, int B, int value)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B) flags[i] = value;
}

// LevenbergMarquardtSparse::solve for the whole batch, host-driven passes.  Blocks the calling thread (it reads two flags per
// pass); returns false if a pass count bound is hit (never observed: every pass either ends an outer iteration or multiplies mu).
template <class M, int DEFECT, class Real>
bool launchPipelineT(const DeviceOcp& P, const DeviceState& st, const PipeArraysT<Real>& pa, int iterations, cudaStream_t stream)
{
    const int B = P.B, K = P.K;
    const int warp_blocks_ik = (int)(((long long)B * K + 3) / 4), warp_blocks_i = (B + 3) / 4, thread_blocks_i = (B + 127) / 128;
    constexpr int IPW = (sizeof(Real) == 4 && PipeDim<M>::NB <= 16) ? 2 : 1;  // instances per warp of the factor kernel
    int any[2] = {0, 0};
    cudaMemsetAsync(pa.any, 0, 2 * sizeof(int), stream);
    pipeSetFlagsKernel<<<thread_blocks_i, 128, 0, stream>>>(pa.flags, B, PF_LIN);
    pipeLinearizeKernel<M, DEFECT, Real><<<warp_blocks_ik, 128, 0, stream>>>(P, st, pa);
    pipeInitKernel<M, Real><<<warp_blocks_i, 128, 0, stream>>>(P, st, pa, iterations);
    cudaMemcpyAsync(any, pa.any, 2 * sizeof(int), cudaMemcpyDeviceToHost, stream);
    cudaStreamSynchronize(stream);
    int passes = 0;
    const int max_passes = 64 * (iterations + 1);
    while (any[0] && passes < max_passes)
    {
        pipeFactorKernel<M, Real, IPW><<<(B + 4 * IPW - 1) / (4 * IPW), 128, 0, stream>>>(P, st, pa);
        pipeTrialKernel<M, DEFECT, Real><<<dim3(thread_blocks_i, K), 128, 0, stream>>>(P, st, pa);
        cudaMemsetAsync(pa.any, 0, 2 * sizeof(int), stream);
        pipeControlKernel<Real><<<thread_blocks_i, 128, 0, stream>>>(P, st, pa, iterations);
        cudaMemcpyAsync(any, pa.any, 2 * sizeof(int), cudaMemcpyDeviceToHost, stream);
        cudaStreamSynchronize(stream);
        if (any[0] && any[1]) pipeLinearizeKernel<M, DEFECT, Real><<<warp_blocks_ik, 128, 0, stream>>>(P, st, pa);
        ++passes;
    }
    return passes < max_passes;
}

template <class M, int DEFECT>
bool launchPipeline(const DeviceOcp& P, const DeviceState& st, const PipeArraysT<double>& pa, int iterations, cudaStream_t stream)
{
    return launchPipelineT<M, DEFECT, double>(P, st, pa, iterations, stream);
}
template <class M, int DEFECT>
bool launchPipelineF32(const DeviceOcp& P, const DeviceState& st, const PipeArraysT<float>& pa, int iterations, cudaStream_t stream)
{
    return launchPipelineT<M, DEFECT, float>(P, st, pa, iterations, stream);
}

}  // namespace b200sqp
