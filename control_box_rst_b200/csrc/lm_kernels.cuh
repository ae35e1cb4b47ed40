// Kernels of the batched LM solver, templated on <system model, defect constraint, variable-dt flag>, plus the launch table
// that the C ABI dispatches through.  One thread per OCP instance; a thread block is one warp so that a batch spreads over as
// many SMs as possible (4096 instances = 128 warps on 148 SMs).
#pragma once

#include <math_constants.h>

#include "launch.h"
#include "lm_device.cuh"

namespace b200sqp {

// LevenbergMarquardtSparse::solve (optimization/src/solver/levenberg_marquardt_sparse.cpp:44-220) for one instance.
// Mirrored quirks: damping accumulates on the Hessian diagonal across inner passes and is never removed (:135-138,:208) ->
// mu_acc; all `iterations` outer passes run (:129); `stop` is overwritten by ||values|| <= eps3 where `values` is whatever
// computeValues produced last, i.e. possibly a rejected trial point (:216); the last outer pass never re-linearises (:178);
// `v` is an unsigned int (:108).
template <class M, int DEFECT, int VT>
__global__ void __launch_bounds__(32) lmSolveKernel(const __grid_constant__ DeviceOcp P, const __grid_constant__ DeviceState st, int iterations)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.B) return;
    const Weights w{st.w_eq, st.w_ineq, st.w_b};
    const double* x0p   = st.x0 + i;
    const double* xrefp = st.xref + i;
    double* D  = st.D + i;
    double* E  = st.E + i;
    double* g  = st.g + i;
    double* dl = st.dl + i;
    double* L  = st.L + i;
    double* W  = st.W + i;
    int cur    = st.cur[i];

    constexpr double eps1 = 1e-5, eps2 = 1e-5, eps3 = 1e-5, eps4 = 0;
    constexpr double tau                = 1e-5;
    constexpr double goodStepUpperScale = 2. / 3., goodStepLowerScale = 1. / 3.;

    int n_factor = 0, n_reject = 0, n_lin = 0;

    double chi2_old, ginf, maxdiag;
    {
        NormalEquationSink<M, VT> sink(P, D, E, g);
        linearizeSweep<M, DEFECT, VT>(P, w, st.z[cur] + i, x0p, xrefp, sink);
        chi2_old = sink.chi2;
        ginf     = sink.ginf;
        maxdiag  = sink.maxdiag;
        ++n_lin;
    }
    unsigned int v = 2;
    bool stop      = ginf <= eps1;
    double mu      = tau * maxdiag;
    if (mu < 0) mu = 0;
    double mu_acc      = 0.0;  // what has been added to the Hessian diagonal since the last re-linearisation
    double rho         = 0;
    double last_values = chi2_old;  // squaredNorm of the reference's `_values` member
    if (st.trace) st.trace[i] = chi2_old;

    for (int k = 0; k < iterations; ++k)
    {
        do
        {
            mu_acc += mu;
            double dn2, dq;
            factorSolve<M, VT>(P, D, E, g, L, W, dl, mu_acc, mu, dn2, dq);
            ++n_factor;
            if (sqrt(dn2) <= eps2)
            {
                stop = true;
            }
            else
            {
                const double chi2_new = trialChi2<M, DEFECT, VT>(P, w, st.z[cur] + i, dl, st.z[cur ^ 1] + i, x0p, xrefp);
                last_values           = chi2_new;
                rho                   = (chi2_old - chi2_new) / dq;
                if (rho > 0 && !isnan(chi2_new) && !isinf(chi2_new))
                {
                    stop = (sqrt(chi2_old) - sqrt(chi2_new) < eps4 * sqrt(chi2_old));
                    cur ^= 1;  // accept: the trial buffer becomes the current one (discardBackupParameters)
                    if (!stop && k < iterations - 1)
                    {
                        NormalEquationSink<M, VT> sink(P, D, E, g);
                        linearizeSweep<M, DEFECT, VT>(P, w, st.z[cur] + i, x0p, xrefp, sink);
                        ++n_lin;
                        mu_acc             = 0.0;
                        stop               = stop || (sink.ginf <= eps1);
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
                }
            }
        } while (rho <= 0 && !stop);
        stop = (sqrt(last_values) <= eps3);
        if (st.trace) st.trace[(size_t)(k + 1) * P.S + i] = chi2_old;
    }
    st.cur[i]         = cur;
    st.chi2[i]        = chi2_old;
    st.mu[i]          = mu;
    st.rho[i]         = rho;
    st.status[i]      = (stop || rho <= 0) ? B200SQP_STATUS_CONVERGED : B200SQP_STATUS_EARLY_TERMINATED;
    st.n_factor[i]    = n_factor;
    st.n_reject[i]    = n_reject;
    st.n_linearize[i] = n_lin;
}

// LevenbergMarquardtSparse::computeValues + computeCombinedSparseJacobian, materialised (b200sqp_evaluate)
template <class M, int DEFECT, int VT>
__global__ void __launch_bounds__(32) evaluateKernel(const __grid_constant__ DeviceOcp P, const __grid_constant__ DeviceState st, double* values,
                                                     double* jac, const int* value_rows, const int* jac_pos, int v_count, int j_count)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.B) return;
    const Weights w{st.w_eq, st.w_ineq, st.w_b};
    MaterializeSink<M, VT> sink{P, values ? values + i : nullptr, jac ? jac + i : nullptr, value_rows, jac_pos, v_count, j_count};
    linearizeSweep<M, DEFECT, VT>(P, w, st.z[st.cur[i]] + i, st.x0 + i, st.xref + i, sink);
}

template <class M, int DEFECT, int VT>
void launchSolve(const DeviceOcp& P, const DeviceState& st, int iterations, cudaStream_t stream)
{
    const int blocks = (P.B + 31) / 32;
    lmSolveKernel<M, DEFECT, VT><<<blocks, 32, 0, stream>>>(P, st, iterations);
}

template <class M, int DEFECT, int VT>
void launchEvaluate(const DeviceOcp& P, const DeviceState& st, double* values, double* jac, const int* value_rows, const int* jac_pos, int v_count,
                    int j_count, cudaStream_t stream)
{
    const int blocks = (P.B + 31) / 32;
    evaluateKernel<M, DEFECT, VT><<<blocks, 32, 0, stream>>>(P, st, values, jac, value_rows, jac_pos, v_count, j_count);
}

#define B200SQP_KERNEL_ENTRY(MODEL, DEFECT, VT) \
    KernelSet { MODEL::ID, DEFECT, VT, MODEL::NX, MODEL::NU, &launchSolve<MODEL, DEFECT, VT>, &launchEvaluate<MODEL, DEFECT, VT> }

}  // namespace b200sqp
