// Kernels of the batched LM solver, templated on <system model, defect constraint, variable-dt flag>, plus the launch table
// that the C ABI dispatches through.  One thread per OCP instance; a thread block is one warp so that a batch spreads over as
// many SMs as possible (4096 instances = 128 warps on 148 SMs).
#pragma once

#include <math_constants.h>

#include "launch.h"
#include "lm_device.cuh"

namespace b200sqp {

// LevenbergMarquardtSparse::solve (optimization/src/solver/levenberg_marquardt_sparse.cpp:44-220) for 32 instances per thread block.
//
// Thread layout: blockDim = 32*T; thread (g = tid & 31, p = tid >> 5) is cooperating thread p of the block's instance g, i.e. warp p
// holds "lane p" of all 32 instances, so every global access of a warp is 32 consecutive doubles.  The T threads of an instance
// split the horizon into T contiguous chunks for the two embarrassingly parallel phases (linearisation with the FD Jacobians,
// trial-point evaluation); the block-tridiagonal factorisation is a sequential recursion and runs on thread p = 0 (warp 0) while
// the other warps wait at the barrier without consuming issue slots.  Phases are separated by block barriers; the per-instance LM
// state machine lives in the registers of thread p = 0 and is broadcast through shared memory.
//
// Mirrored quirks of the reference loop: damping accumulates on the Hessian diagonal across inner passes and is never removed
// (:135-138,:208) -> mu_acc; all `iterations` outer passes run (:129); `stop` is overwritten by ||values|| <= eps3 where `values`
// is whatever computeValues produced last, i.e. possibly a rejected trial point (:216); the last outer pass never re-linearises
// (:178); `v` is an unsigned int (:108); the `stop || ||g||inf <= eps1` update after a re-linearisation (:193) is dead code
// because rho > 0 leaves the inner loop and :216 overwrites `stop`, so ||g||inf is only evaluated for the initial test (:115).
struct TrueTag
{
    static constexpr bool value = true;
};
struct FalseTag
{
    static constexpr bool value = false;
};

template <class M, int DEFECT, int VT, int T, class F>
__global__ void __launch_bounds__(32 * T) lmSolveKernel(const __grid_constant__ DeviceOcp P, const __grid_constant__ DeviceState st, int iterations)
{
    using Dm = Dim<M, VT>;
    constexpr int NX = Dm::NX, XO = Dm::XO, NB = Dm::NB, ND = Dm::ND;
    const int g      = threadIdx.x & 31;
    const int p      = threadIdx.x >> 5;
    const int i      = blockIdx.x * 32 + g;
    const bool valid = i < P.B;
    const int K      = P.K;
    const int ka     = (int)((long long)p * K / T);
    const int kb     = (int)((long long)(p + 1) * K / T);
    const int S      = P.S;

    using BS = BlockSolver<M, VT>;
    // coupling slots of a block (lm_device.cuh Dim): the state, plus the dt slot when consecutive dt vertices are coupled (VT == 2)
    constexpr int NC = Dm::NC, NCC = Dm::NCC;
    constexpr bool TWISTED = T >= 2;  // two threads of an instance eliminate from both ends of the horizon
    __shared__ double s_red[3][T][32];
    __shared__ double s_muacc[32], s_mu[32];
    __shared__ double s_xc[TWISTED ? NCC + 2 * NC : 1][32];  // chain B -> A: Schur contribution; chain A -> B: coupling part of delta_m
    __shared__ double s_dn[2][T][32];                        // partial ||delta||^2 and delta^T(mu delta + g) per cooperating thread
    __shared__ int s_cur[32], s_flags[32];
    enum { F_ACTIVE = 1, F_LIN = 2 };
    const bool use_part = TWISTED && K >= 2 * T;  // partitioned factorisation: every chunk has a separator and >= 1 interior block

    const Weights w{st.w_eq, st.w_ineq, st.w_b};
    // tiled arrays: this block's tile, this thread's lane (lm_device.cuh TILE)
    constexpr int NE_   = Dm::NE;
    const size_t tile   = blockIdx.x;
    const double* x0p   = st.x0 + tile * ((size_t)NX * TILE) + g;
    const double* xrefp = st.xref + tile * ((size_t)NX * TILE) + g;
    const double* xtrajp = (F::xref_traj && st.xref_traj) ? st.xref_traj + tile * ((size_t)(K + 1) * NX * TILE) + g : nullptr;
    double* D  = st.D + tile * ((size_t)K * ND * TILE) + g;
    double* E  = st.E + tile * ((size_t)K * NE_ * TILE) + g;
    double* gg = st.g + tile * ((size_t)K * NB * TILE) + g;
    double* dl = st.dl + tile * ((size_t)K * NB * TILE) + g;
    double* L  = st.L + tile * ((size_t)K * ND * TILE) + g;
    double* W  = st.W + tile * ((size_t)K * NE_ * TILE) + g;
    double* Yf = st.Y + tile * ((size_t)K * NE_ * TILE) + g;  // spike factors of the partitioned factorisation
    // reduced (separator) system of the partitioned factorisation: st.red_blocks (= max cooperating threads - 1) blocks per instance
    double* Dr  = st.red + tile * ((size_t)st.red_blocks * (2 * ND + 2 * NE_ + 2 * NB) * TILE) + g;
    double* Er  = Dr + (size_t)st.red_blocks * ND * TILE;
    double* gr  = Er + (size_t)st.red_blocks * NE_ * TILE;
    double* Lr  = gr + (size_t)st.red_blocks * NB * TILE;
    double* Wr  = Lr + (size_t)st.red_blocks * ND * TILE;
    double* dlr = Wr + (size_t)st.red_blocks * NE_ * TILE;
    const size_t zoff = tile * ((size_t)K * NB * TILE) + g;

    constexpr double eps1 = 1e-5, eps2 = 1e-5, eps3 = 1e-5, eps4 = 0;
    constexpr double tau                = 1e-5;
    constexpr double goodStepUpperScale = 2. / 3., goodStepLowerScale = 1. / 3.;

    // LM state of instance g (meaningful in thread p == 0 only)
    int cur = 0, k_outer = 0, n_factor = 0, n_reject = 0, n_lin = 0;
    unsigned int v = 2;
    bool stop = false, active = false;
    double mu = 0, mu_acc = 0, rho = 0, chi2_old = 0, last_values = 0, dq = 0;

    // optional phase profile (b200sqp_set_phase_profile): thread 0 of the block accumulates clock64() deltas per phase
    long long prof_acc[5] = {0, 0, 0, 0, 0};
    long long prof_last   = 0;
    const bool prof       = st.phase_cycles != nullptr && threadIdx.x == 0;
    if (prof) prof_last = clock64();
    auto tick = [&](int phase) {
        if (prof)
        {
            const long long now = clock64();
            prof_acc[phase] += now - prof_last;
            prof_last = now;
        }
    };

    if (p == 0)
    {
        cur        = valid ? st.cur[i] : 0;
        s_cur[g]   = cur;
        s_flags[g] = valid ? (F_ACTIVE | F_LIN) : 0;
    }
    __syncthreads();

    // One linearisation phase for the instances flagged F_LIN; executed by the whole block because it contains barriers.
    auto linearizePhase = [&](bool first) {
        const bool do_lin = valid && (s_flags[g] & F_LIN);
        double* z         = st.z[s_cur[g]] + zoff;
        double xn_last[NX];
        double t_prev = 0.0, t_last = 0.0;  // VT == 2: dt_{ka-1} and dt_{kb-1} before this linearisation (each is rewritten by a neighbouring chunk)
        if (do_lin && kb < K)
        {
            const double* zp = z + (size_t)(kb - 1) * NB * TILE;
#pragma unroll
            for (int j = 0; j < NX; ++j) xn_last[j] = zp[(size_t)(XO + j) * TILE];
            if (Dm::DTEQ) t_last = zp[(size_t)Dm::NU * TILE];
        }
        if (Dm::DTEQ && do_lin && ka > 0) t_prev = z[(size_t)((ka - 1) * NB + Dm::NU) * TILE];
        if (T > 1) __syncthreads();  // boundary states are read before any neighbour writes its perturbed copy back
        NormalEquationSink<M, VT, F> sink(P, D, E, gg, ka, kb);
        if (do_lin) linearizeSweep<M, DEFECT, VT, F>(P, w, z, x0p, xrefp, xtrajp, ka, kb, xn_last, sink, F::dense ? st.cost_sqrt_full : nullptr, t_prev, t_last);
        if (T > 1)
        {
            __syncthreads();  // all blocks stored; now the chunk-start contributions can be added to the neighbour's last block
            if (do_lin) sink.addBoundary();
        }
        if (first)
        {
            s_red[0][p][g] = sink.chi2;
            s_red[1][p][g] = sink.ginf;
            s_red[2][p][g] = sink.maxdiag;
        }
        __syncthreads();
    };

    // LM state after the first linearisation (levenberg_marquardt_sparse.cpp:103-126)
    auto initState = [&]() {
        if (p == 0 && valid)
        {
            double ginf = 0.0, maxdiag = -CUDART_INF;
            chi2_old = 0.0;
#pragma unroll
            for (int q = 0; q < T; ++q)
            {
                chi2_old += s_red[0][q][g];
                ginf    = fmax(ginf, s_red[1][q][g]);
                maxdiag = fmax(maxdiag, s_red[2][q][g]);
            }
            ++n_lin;
            stop = ginf <= eps1;
            mu   = tau * maxdiag;
            if (mu < 0) mu = 0;
            last_values = chi2_old;
            active      = iterations > 0;
            if (st.trace) st.trace[i] = chi2_old;
            mu_acc      = mu;
            s_muacc[g]  = mu_acc;
            s_mu[g]     = mu;
            s_flags[g]  = active ? F_ACTIVE : 0;
        }
        else if (p == 0)
            s_flags[g] = 0;
        __syncthreads();
    };
    // Kernels whose sweep is large (Runge-Kutta shooting: the linearisation is most of the code) keep ONE call site for it -- the first
    // linearisation becomes the first pass of the loop -- so that the code exists once in the instruction stream; the small polynomial
    // kernels keep the first linearisation in front of the loop (measured: Van der Pol 0.352 ms against 0.376 ms with one call site).
#ifndef B200SQP_ONE_SITE_RULE
#define B200SQP_ONE_SITE_RULE (DEFECT == DEFECT_RK4)
#endif
    constexpr bool ONE_SITE = B200SQP_ONE_SITE_RULE;
    if constexpr (!ONE_SITE)
    {
        linearizePhase(true);
        tick(0);
        initState();
    }
    bool first = ONE_SITE, relin = ONE_SITE;
    while (true)
    {
        if constexpr (ONE_SITE)
        {
            // ---- L: linearise -- the initial point, later the accepted points
            if (relin) linearizePhase(first);
            tick(0);
            if (first)
            {
                first = false;
                initState();
            }
        }
        // ---- F: (H + sum(mu) I) delta = g.
        //      K >= 2T: partitioned -- every thread eliminates the interior of its chunk, the T-1 separators form a reduced
        //      block-tridiagonal system for the twisted chains, every thread back-substitutes (BlockSolver::part*).
        //      otherwise: twisted chains over the whole horizon (chain A = thread 0, chain B = thread 1), or one chain for T = 1.
        {
            const bool inst_active = valid && (s_flags[g] & F_ACTIVE);
            const double mua = s_muacc[g], mucur = s_mu[g];
            double pdn2 = 0.0, pdq = 0.0;
            if (p == 0 && inst_active) ++n_factor;
            // twisted solve of a block-tridiagonal system with Kb blocks; ACC: account ||delta||^2 and delta^T(mu delta + g)
            auto twistedSolve = [&](auto acc_tag, const double* Dq, const double* Eq, const double* gq, double* Lq, double* Wq, double* dlq, int Kb,
                                    double mu_add) {
                constexpr bool ACC = decltype(acc_tag)::value;
                const int mid      = Kb / 2;
                double Lp[ND], yp[NC], carry[NC], dx[NC];
                if (p == 0 && inst_active) BS::chainAEliminate(P, Dq, Eq, gq, Lq, Wq, dlq, mu_add, 0, mid, Lp, yp, nullptr, nullptr);
                if (p == 1 && inst_active)
                {
                    double cxx[NCC], cgx[NC];
                    BS::chainBEliminate(P, Dq, Eq, gq, Lq, Wq, dlq, mu_add, mid + 1, Kb, cxx, cgx);
#pragma unroll
                    for (int q = 0; q < NCC; ++q) s_xc[q][g] = cxx[q];
#pragma unroll
                    for (int q = 0; q < NC; ++q) s_xc[NCC + q][g] = cgx[q];
                }
                __syncthreads();
                if (p == 0 && inst_active)
                {
                    double cxx[NCC], cgx[NC];
#pragma unroll
                    for (int q = 0; q < NCC; ++q) cxx[q] = s_xc[q][g];
#pragma unroll
                    for (int q = 0; q < NC; ++q) cgx[q] = s_xc[NCC + q][g];
                    BS::chainAEliminate(P, Dq, Eq, gq, Lq, Wq, dlq, mu_add, mid, mid + 1, Lp, yp, cxx, cgx);
#pragma unroll
                    for (int q = 0; q < NC; ++q) carry[q] = 0.0;
                    BS::template chainABacksub<ACC>(P, gq, Lq, Wq, dlq, mucur, mid, mid, carry, dx, pdn2, pdq);
#pragma unroll
                    for (int q = 0; q < NC; ++q) s_xc[NCC + NC + q][g] = dx[q];
                }
                __syncthreads();
                if (p == 0 && inst_active) BS::template chainABacksub<ACC>(P, gq, Lq, Wq, dlq, mucur, mid - 1, 0, carry, nullptr, pdn2, pdq);
                if (p == 1 && inst_active)
                {
#pragma unroll
                    for (int q = 0; q < NC; ++q) dx[q] = s_xc[NCC + NC + q][g];
                    BS::template chainBSubst<ACC>(P, gq, Lq, Wq, dlq, mucur, mid + 1, Kb, dx, pdn2, pdq);
                }
            };
            if constexpr (TWISTED)
            {
                if (use_part)
                {
                    const bool has_up = p > 0, has_low = p < T - 1;
                    double cxx[NCC], cgx[NC];
                    if (inst_active) BS::partEliminate(D, E, gg, L, W, Yf, dl, Dr, Er, gr, mua, ka, kb, has_up, has_low, p, cxx, cgx);
                    __syncthreads();
                    if (inst_active && has_up) BS::partAddToUpper(Dr, gr, p - 1, cxx, cgx);
                    __syncthreads();
                    twistedSolve(FalseTag{}, Dr, Er, gr, Lr, Wr, dlr, T - 1, 0.0);
                    __syncthreads();
                    if (inst_active) BS::partBacksub(gg, L, W, Yf, dl, dlr, mucur, ka, kb, has_up, has_low, p, pdn2, pdq);
                }
                else
                    twistedSolve(TrueTag{}, D, E, gg, L, W, dl, K, mua);
            }
            else if (p == 0 && inst_active)
            {
                double Lp[ND], yp[NC], carry[NC];
                BS::chainAEliminate(P, D, E, gg, L, W, dl, mua, 0, K, Lp, yp, nullptr, nullptr);
#pragma unroll
                for (int q = 0; q < NC; ++q) carry[q] = 0.0;
                BS::chainABacksub(P, gg, L, W, dl, mucur, K - 1, 0, carry, nullptr, pdn2, pdq);
            }
            s_dn[0][p][g] = pdn2;
            s_dn[1][p][g] = pdq;
        }
        __syncthreads();
        tick(1);
        // ---- T: trial point and its chi2, all T threads
        double dn2_tot = 0.0;
#pragma unroll
        for (int q = 0; q < T; ++q) dn2_tot += s_dn[0][q][g];
        const bool step_small = sqrt(dn2_tot) <= eps2;
        {
            const bool do_trial = valid && (s_flags[g] & F_ACTIVE) && !step_small;
            double part         = 0.0;
            if (do_trial) part = trialChi2<M, DEFECT, VT, F>(P, w, st.z[s_cur[g]] + zoff, dl, st.z[s_cur[g] ^ 1] + zoff, x0p, xrefp, xtrajp, ka, kb,
                                                                 F::dense ? st.cost_sqrt_full : nullptr);
            s_red[0][p][g] = part;
        }
        __syncthreads();
        tick(2);
        // ---- C: gain ratio, accept / reject, damping update (thread p == 0)
        bool any_lin = false;
        if (p == 0)
        {
            int flags = 0;
            if (active)
            {
                dq = 0.0;
#pragma unroll
                for (int q = 0; q < T; ++q) dq += s_dn[1][q][g];
                if (step_small)
                {
                    stop = true;
                }
                else
                {
                    double chi2_new = 0.0;
#pragma unroll
                    for (int q = 0; q < T; ++q) chi2_new += s_red[0][q][g];
                    last_values = chi2_new;
                    rho         = (chi2_old - chi2_new) / dq;
                    if (rho > 0 && !isnan(chi2_new) && !isinf(chi2_new))
                    {
                        stop = (sqrt(chi2_old) - sqrt(chi2_new) < eps4 * sqrt(chi2_old));
                        cur ^= 1;  // accept: the trial buffer becomes the current one (discardBackupParameters)
                        if (!stop && k_outer < iterations - 1)
                        {
                            flags |= F_LIN;
                            ++n_lin;
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
                        ++n_reject;  // restoreBackupParameters: the current buffer was never touched
                        mu = mu * v;
                        v  = 2 * v;
                        // `v` is an unsigned int in the reference (:108): after 31 consecutive rejections it wraps to 0, mu becomes 0 and
                        // -- if the trial point keeps failing (e.g. non-finite residuals) -- LevenbergMarquardtSparse::solve never returns.
                        // A batched kernel must: the iteration is ended instead.
                        if (v == 0) stop = true;
                    }
                }
                if (!(rho <= 0 && !stop))
                {
                    // end of outer iteration k_outer (:216)
                    stop = (sqrt(last_values) <= eps3);
                    ++k_outer;
                    if (st.trace) st.trace[(size_t)k_outer * S + i] = chi2_old;
                    active = k_outer < iterations;
                }
                if (active)
                {
                    flags |= F_ACTIVE;
                    mu_acc += mu;  // what the next factorisation adds to the diagonal (:135-138)
                    s_muacc[g] = mu_acc;
                    s_mu[g]    = mu;
                }
            }
            s_cur[g]   = cur;
            s_flags[g] = flags;
            any_lin    = (flags & F_LIN) != 0;
        }
        const int any_active = __syncthreads_or(p == 0 && active);
        tick(3);
        if (!any_active) break;
        if constexpr (ONE_SITE)
            relin = __syncthreads_or(any_lin) != 0;
        else
        {
            // ---- L: re-linearise the accepted points
            if (__syncthreads_or(any_lin)) linearizePhase(false);
            tick(0);
        }
    }

    // Fused stop-test exchange: instead of a separate all-gather launch after the solve, the per-instance
    // chi2 goes straight into every rank's gather buffer through NVLink peer stores; after a system-scope fence one thread per
    // block bumps this rank's arrival counter on every peer (b200sqp_peer_wait spins on those counters, bounded).
    if (st.peer_world > 0)
    {
        if (p == 0 && valid)
        {
            const size_t slot = (size_t)st.peer_parity * st.peer_world * P.B + (size_t)st.peer_rank * P.B + i;
            for (int r = 0; r < st.peer_world; ++r) st.peer_chi2[r][slot] = chi2_old;
            __threadfence_system();
        }
        __syncthreads();
        if (threadIdx.x == 0)
            for (int r = 0; r < st.peer_world; ++r) atomicAdd_system(st.peer_arrivals[r] + st.peer_rank, 1ULL);
    }
    tick(4);  // the exchange epilogue (zero without attached peers)
    if (prof)
    {
#pragma unroll
        for (int q = 0; q < 5; ++q) st.phase_cycles[(size_t)blockIdx.x * 5 + q] = prof_acc[q];
    }
    if (p == 0 && valid)
    {
        st.cur[i]         = cur;
        st.chi2[i]        = chi2_old;
        st.mu[i]          = mu;
        st.rho[i]         = rho;
        st.status[i]      = (stop || rho <= 0) ? B200SQP_STATUS_CONVERGED : B200SQP_STATUS_EARLY_TERMINATED;
        st.n_factor[i]    = n_factor;
        st.n_reject[i]    = n_reject;
        st.n_linearize[i] = n_lin;
    }
}

// LevenbergMarquardtSparse::computeValues + computeCombinedSparseJacobian, materialised (b200sqp_evaluate)
template <class M, int DEFECT, int VT, class F>
__global__ void __launch_bounds__(32) evaluateKernel(const __grid_constant__ DeviceOcp P, const __grid_constant__ DeviceState st, double* values,
                                                     double* jac, const int* value_rows, const int* jac_pos, int v_count, int j_count)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.B) return;
    const Weights w{st.w_eq, st.w_ineq, st.w_b};
    MaterializeSink<M, VT, F> sink{P, values ? values + i : nullptr, jac ? jac + i : nullptr, value_rows, jac_pos, v_count, j_count};
    using Dm          = Dim<M, VT>;
    const size_t tile = i >> 5, lane = i & 31;
    linearizeSweep<M, DEFECT, VT, F>(P, w, st.z[st.cur[i]] + tile * ((size_t)P.K * Dm::NB * TILE) + lane, st.x0 + tile * ((size_t)Dm::NX * TILE) + lane,
                                     st.xref + tile * ((size_t)Dm::NX * TILE) + lane,
                                     st.xref_traj ? st.xref_traj + tile * ((size_t)(P.K + 1) * Dm::NX * TILE) + lane : nullptr, 0, P.K, nullptr, sink,
                                     F::dense ? st.cost_sqrt_full : nullptr);
}

// Cooperating threads per instance: small batches are latency bound (one warp per SM would leave the machine idle), so the
// horizon is split T ways; once a batch alone fills the SMs with warps T = 1 is the most work-efficient mapping.
// MAXT bounds the variants that get compiled for a (model, defect, grid) combination.
template <class M, int DEFECT, int VT, int MAXT, class F>
void launchSolveT(const DeviceOcp& P, const DeviceState& st, int iterations, int T, int blocks, cudaStream_t stream)
{
    if constexpr (MAXT >= 8)
        if (T >= 8) return (void)lmSolveKernel<M, DEFECT, VT, 8, F><<<blocks, 256, 0, stream>>>(P, st, iterations);
    if constexpr (MAXT >= 4)
        if (T >= 4) return (void)lmSolveKernel<M, DEFECT, VT, 4, F><<<blocks, 128, 0, stream>>>(P, st, iterations);
    if constexpr (MAXT >= 2)
        if (T >= 2) return (void)lmSolveKernel<M, DEFECT, VT, 2, F><<<blocks, 64, 0, stream>>>(P, st, iterations);
    lmSolveKernel<M, DEFECT, VT, 1, F><<<blocks, 32, 0, stream>>>(P, st, iterations);
}

template <class M, int DEFECT, int VT, int MAXT>
void launchSolve(const DeviceOcp& P, const DeviceState& st, int iterations, int threads_per_instance, int flags, cudaStream_t stream)
{
    const int blocks = (P.B + 31) / 32;
    int T            = threads_per_instance;
    if (T <= 0)
    {
        // measured on B200 (profiles/): the widest split wins at every batch size from 1k to 64k instances as long as every
        // thread keeps >= 3 intervals; the two-sided elimination alone is worth T = 2
        T = 1;
        while (T < 8 && P.K / (2 * T) >= 3) T *= 2;
    }
    if (T > MAXT) T = MAXT;
    // lean feature set: quadratic lsq stage cost, no state bounds, no pinned goal components, no final-stage constraint
    bool lean = P.final_constraint == 0 && P.stage_cost == B200SQP_COST_QUADRATIC_LSQ;
    for (int j = 0; j < M::NX; ++j) lean = lean && !P.x_bounded[j] && !P.xf_fixed[j];
    if (flags & SOLVE_FORCE_GENERAL_FEATURES) lean = false;
    if (st.xref_traj) lean = false;  // a time-varying state reference is a feature of the general set only  // b200sqp_set_feature_set: parity tests run both variants on one structure
    if constexpr (VT >= 1)
    {
        // time-optimal feature set: minimum-time lsq cost, no state bounds, no final-stage constraint, static reference
        bool topt = P.final_constraint == 0 && P.stage_cost == B200SQP_COST_MINIMUM_TIME_LSQ && !st.xref_traj && !(flags & SOLVE_FORCE_GENERAL_FEATURES);
        for (int j = 0; j < M::NX; ++j) topt = topt && !P.x_bounded[j];
        if (topt) return launchSolveT<M, DEFECT, VT, MAXT, FeatTimeOpt>(P, st, iterations, T, blocks, stream);
    }
    if (lean)
        launchSolveT<M, DEFECT, VT, MAXT, FeatLean>(P, st, iterations, T, blocks, stream);
    else
        launchSolveT<M, DEFECT, VT, MAXT, FeatAll>(P, st, iterations, T, blocks, stream);
}

template <class M, int DEFECT, int VT>
void launchEvaluate(const DeviceOcp& P, const DeviceState& st, double* values, double* jac, const int* value_rows, const int* jac_pos, int v_count,
                    int j_count, cudaStream_t stream)
{
    const int blocks = (P.B + 31) / 32;
    evaluateKernel<M, DEFECT, VT, FeatAll><<<blocks, 32, 0, stream>>>(P, st, values, jac, value_rows, jac_pos, v_count, j_count);
}

// the same two entry points for structures with full (non-diagonal) cost weights: the general feature set + dense cost blocks
template <class M, int DEFECT, int VT, int MAXT>
void launchSolveDense(const DeviceOcp& P, const DeviceState& st, int iterations, int threads_per_instance, int /*flags*/, cudaStream_t stream)
{
    const int blocks = (P.B + 31) / 32;
    int T            = threads_per_instance;
    if (T <= 0)
    {
        T = 1;
        while (T < 8 && P.K / (2 * T) >= 3) T *= 2;
    }
    if (T > MAXT) T = MAXT;
    launchSolveT<M, DEFECT, VT, MAXT, FeatDense>(P, st, iterations, T, blocks, stream);
}
template <class M, int DEFECT, int VT>
void launchEvaluateDense(const DeviceOcp& P, const DeviceState& st, double* values, double* jac, const int* value_rows, const int* jac_pos, int v_count,
                         int j_count, cudaStream_t stream)
{
    const int blocks = (P.B + 31) / 32;
    evaluateKernel<M, DEFECT, VT, FeatDense><<<blocks, 32, 0, stream>>>(P, st, values, jac, value_rows, jac_pos, v_count, j_count);
}

#define B200SQP_KERNEL_ENTRY_DENSE(MODEL, DEFECT, VT, MAXT) \
    KernelSet { MODEL::ID, DEFECT, VT, MODEL::NX, MODEL::NU, MAXT, &launchSolveDense<MODEL, DEFECT, VT, MAXT>, &launchEvaluateDense<MODEL, DEFECT, VT>, nullptr }
#define B200SQP_KERNEL_ENTRY(MODEL, DEFECT, VT, MAXT) \
    KernelSet { MODEL::ID, DEFECT, VT, MODEL::NX, MODEL::NU, MAXT, &launchSolve<MODEL, DEFECT, VT, MAXT>, &launchEvaluate<MODEL, DEFECT, VT>, nullptr }
}  // namespace b200sqp
