// Plain structs shared by host (api.cpp) and device code: the per-structure constants and the per-handle device arrays.
#pragma once

#include "../../include/b200sqp.h"

namespace b200sqp {

struct DynParams
{
    double p[B200SQP_MAX_DYN_PARAMS];
    // RN(1 / p[i]) and whether x / p[i] may be evaluated as the correctly rounded three-operation sequence of dynamics.cuh
    // (StepSize): models that divide by a parameter (quadrotor: mass, inertias) then need no IEEE divide per evaluation
    double rcp[B200SQP_MAX_DYN_PARAMS];
    unsigned fast_div_mask;
};

// host side: fill rcp / fast_div_mask from p (same admissibility rule as StepSize: finite, normal, moderate exponent, significand not all ones)
inline void prepareDynParams(DynParams& d)
{
    d.fast_div_mask = 0;
    for (int i = 0; i < B200SQP_MAX_DYN_PARAMS; ++i)
    {
        d.rcp[i] = 0.0;
        unsigned long long b;
        static_assert(sizeof(b) == sizeof(double), "");
        __builtin_memcpy(&b, &d.p[i], sizeof(b));
        const unsigned long long man = b & 0xFFFFFFFFFFFFFull;
        const unsigned ex            = (unsigned)(b >> 52) & 0x7FFu;
        if (d.p[i] != 0.0 && man != 0xFFFFFFFFFFFFFull && ex > 523u && ex < 1523u)
        {
            d.rcp[i] = 1.0 / d.p[i];
            d.fast_div_mask |= 1u << i;
        }
    }
}

struct DeviceOcp
{
    int K;       // intervals
    int B;       // instances
    int S;       // instance stride of all instance-minor arrays (B rounded up to a multiple of 32)
    int stage_cost, final_cost, tcost_every_interval;
    int xf_fixed[B200SQP_MAX_NX];
    int x_bounded[B200SQP_MAX_NX], u_bounded[B200SQP_MAX_NU], dt_bounded;
    int final_constraint;  // b200sqp_final_constraint (0 if xf is fully fixed: the reference then creates no final-stage edge)
    int q_dense, r_dense, qf_dense;  // full (non-diagonal) cost weights: the upper Cholesky factors are in DeviceState::cost_sqrt_full
    double term_xref[B200SQP_MAX_NX], term_s[B200SQP_MAX_NX], term_gamma;
    DynParams dyn;
    double dt_ref, dt_lb, dt_ub, tcost_w;
    double q_sqrt[B200SQP_MAX_NX], r_sqrt[B200SQP_MAX_NU], qf_sqrt[B200SQP_MAX_NX];
    double x_lb[B200SQP_MAX_NX], x_ub[B200SQP_MAX_NX], u_lb[B200SQP_MAX_NU], u_ub[B200SQP_MAX_NU];
};

constexpr int MAX_PEERS = 8;  // GPUs of one NVSwitch box

struct Weights
{
    double eq, ineq, b;
};

// per-handle device arrays (all instance-minor unless noted)
struct DeviceState
{
    double* z[2];     // [K*NB][S] two parameter buffers: current and trial (roles swap per instance on accept)
    double* x0;       // [NX][S]
    double* xref;     // [NX][S]   static state reference = the reference at the last grid point
    const double* cost_sqrt_full;  // [nx*nx | nu*nu | nx*nx] row-major upper factors of Q, R, Qf (structure constants, not per instance) or null
    double* xref_traj;  // [(K+1)*NX][S] time-varying state reference, row m = getReferenceCached(m), or null (static reference)
    double* D;        // [K*ND][S] diagonal blocks of J^T J (packed lower)
    double* E;        // [K*NB*NX][S] sub-diagonal coupling blocks
    double* g;        // [K*NB][S]  J^T(-r)
    double* dl;       // [K*NB][S]  step
    double* L;        // [K*ND][S]  factor, diagonal blocks (reciprocal diagonal stored)
    double* W;        // [K*NB*NX][S] factor, sub-diagonal blocks
    double* Y;        // [K*NB*NX][S] factor, spike blocks of the partitioned factorisation (coupling to the separator above a chunk)
    double* red;      // [red_blocks*(2 ND + 2 NB*NX + 2 NB)][S] reduced separator system of the partitioned factorisation: D,E,g,L,W,dl
    int red_blocks;   // separators per instance = widest cooperating-thread variant - 1 (>= 1)
    // per instance results / LM state, [S]
    double* chi2;
    double* mu;
    double* rho;
    int* status;
    int* cur;         // which of z[0]/z[1] holds the current parameters
    int* n_factor;
    int* n_reject;
    int* n_linearize;
    double* trace;    // [(max_iterations+1)][S] chi2 after every outer iteration
    // fused stop-test gather over NVLink peer memory (b200sqp_peer_*): every rank's LM kernel stores its per-instance chi2 straight
    // into every peer's gather buffer and bumps a per-rank arrival counter there; null / 0 when not attached
    double* peer_chi2[MAX_PEERS];              // peer r's gather buffer [2][world*B] (double-buffered by solve parity), own included
    unsigned long long* peer_arrivals[MAX_PEERS];  // peer r's arrival counters [world]
    int peer_world, peer_rank, peer_parity;
    long long* phase_cycles;  // [blocks][4] clock64() per phase (linearise, factor+solve, trial, control) or null (profiling off)
    double w_eq, w_ineq, w_b;  // current penalty weights (host-managed: reset / adapted per solve)
};

// device arrays of the warp-cooperative pipeline for large stage blocks (lm_pipeline.cuh); Hessian blocks instance-major.
// Real = double, or float for the reduced-precision variant of BASELINE configs[4] (b200sqp_set_precision): the normal equations and
// their factor are then evaluated, stored and factorised in fp32 (half the HBM traffic of the factorisation, twice the FMA rate),
// while parameters, steps, residual norms and the LM control state stay fp64.
template <class Real>
struct PipeArraysT
{
    Real* D;    // [B][K][nb(nb+1)/2] G^T G of block k (packed lower by rows) + diagonal cost / bound rows
    Real* E;    // [B][K][nb*nx]      coupling of block k to the x-part of block k-1
    Real* DA;   // [B][K][nxxp]       A^T A of interval k: belongs to the x-x part of block k-1 (nxxp = nx(nx+1)/2 rounded up to 16 bytes)
    Real* gA;   // [B][K][nx]         -A^T e of interval k: belongs to the x-part of g of block k-1
    Real* g;    // [B][K][nb]
    Real* L;    // [B][K][nb(nb+1)/2] Cholesky factor blocks (reciprocal diagonal)
    Real* W;    // [B][K][nb*nx]
    Real* y;    // [B][K][nb]         forward-substituted right-hand side
    double* cpart;        // [K][S] residual norm per interval (linearisation point / trial point)
    double* mu_acc;       // [S] damping accumulated on the diagonal since the last linearisation
    double* last_values;  // [S]
    double* dn2;          // [S] ||delta||^2 (NaN: the factorisation of this pass failed -> the control kernel rejects the step)
    double* dq;           // [S] delta^T (mu delta + g)
    unsigned* v;          // [S]
    int* k_outer;         // [S]
    int* flags;           // [S] PF_ACTIVE | PF_LIN | PF_STOP
    int* any;             // [2] any instance active / any instance to re-linearise
};
using PipeArrays    = PipeArraysT<double>;
using PipeArraysF32 = PipeArraysT<float>;

// packed size of an nx x nx lower triangle, padded so that consecutive blocks stay 16-byte aligned for the bulk copies
inline int paddedTriangle(int nx, int bytes_per_value)
{
    const int n = nx * (nx + 1) / 2, per16 = 16 / bytes_per_value;
    return (n + per16 - 1) / per16 * per16;
}

// grid adaptation (adapt_kernels.cu): what the kernels see of one bucket (a solver handle whose grid has K + 1 points)
struct AdaptBucketView
{
    double* z[2];
    int* cur;
    double* x0;
    double* xref;
    double* chi2;
    int* status;
    int K;
};
enum { ADAPT_NONE = 0, ADAPT_SPLIT = 1, ADAPT_MERGE = 2 };  // decision = type | interval << 2
constexpr int ADAPT_KMAX = 128;  // intervals per instance the edit scripts of the redundant-controls strategy are sized for

// structures the pipeline covers (lm_pipeline.cuh); everything else runs through the fused kernel
inline bool pipelineEligible(const DeviceOcp& P, int nx)
{
    bool ok = P.final_constraint == 0 && P.stage_cost == B200SQP_COST_QUADRATIC_LSQ;
    for (int j = 0; j < nx; ++j) ok = ok && !P.x_bounded[j] && !P.xf_fixed[j];
    return ok;
}

}  // namespace b200sqp
