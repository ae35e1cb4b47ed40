// C ABI of libb200sqp.so (include/b200sqp.h).  Host-side plumbing only: handle lifetime, device buffers, H2D/D2H copies, weight
// reset/adaptation, kernel dispatch through the launch table.  There is no CPU implementation of the hot path in this library:
// without a usable CUDA device every compute entry point fails with B200SQP_ERR_NO_DEVICE.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/b200sqp.h"
#include "launch.h"
#include "lm_device_types.h"
#include "structure.h"

using namespace b200sqp;

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string& msg)
{
    g_last_error = msg;
    return code;
}

#define CUDA_TRY(expr)                                                                                         \
    do                                                                                                         \
    {                                                                                                          \
        cudaError_t _e = (expr);                                                                               \
        if (_e != cudaSuccess) return fail(B200SQP_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)

}  // namespace

namespace b200sqp {

const KernelSet* findKernels(int dynamics, int defect, int vt)
{
#if defined(B200SQP_DEV_TRIO)  // development build (make dev3): the kernels of BASELINE configs[1..3] only
    const KernelSet* (*tables[])(int*) = {kernelTableVdpCn, kernelTableCartPole, kernelTableUnicycle};
#elif defined(B200SQP_DEV_VDP_CN_ONLY)  // development build (make dev): only the benchmark kernels, seconds instead of minutes to compile
    const KernelSet* (*tables[])(int*) = {kernelTableVdpCn};
#else
    const KernelSet* (*tables[])(int*) = {kernelTableVdpCn,       kernelTableVdpFd,    kernelTableVdpMs,   kernelTableOscillators,
                                          kernelTableCartPole,    kernelTableUnicycle, kernelTableQuadrotor, kernelTableBenchmarkSystems,
                                          kernelTableCombosFd,    kernelTableCombosMs, kernelTableLinear,   kernelTableLinear4,
                                          kernelTableIntegrators, kernelTableDtEquality};
#endif
    for (auto t : tables)
    {
        int count            = 0;
        const KernelSet* set = t(&count);
        for (int i = 0; i < count; ++i)
            if (set[i].dynamics == dynamics && set[i].defect == defect && set[i].vt == vt) return &set[i];
    }
    return nullptr;
}

const KernelSet* findDenseCostKernels(int dynamics, int defect, int vt)
{
#if defined(B200SQP_DEV_TRIO) || defined(B200SQP_DEV_VDP_CN_ONLY)
    return nullptr;
#else
    int count            = 0;
    const KernelSet* set = kernelTableDenseCost(&count);
    for (int i = 0; i < count; ++i)
        if (set[i].dynamics == dynamics && set[i].defect == defect && set[i].vt == vt) return &set[i];
    return nullptr;
#endif
}

}  // namespace b200sqp

struct b200sqp_solver
{
    Structure s;
    const KernelSet* kernels = nullptr;
    int B = 0, S = 0, device = 0;
    int max_iterations = 0;
    DeviceOcp P{};
    DeviceState st{};
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
    bool timed = false;
    int64_t launches = 0;
    bool weights_initialised = false;
    int threads_per_instance = 0;  // 0 = heuristic (launchSolve)
    int solve_flags = 0;           // SOLVE_* (launch.h)
    // staging
    double* d_params = nullptr;   // [B][n]
    double* d_x0_host_order = nullptr, *d_xref_host_order = nullptr;  // [B][nx]
    double* d_xref_traj = nullptr, *d_xref_traj_host_order = nullptr;  // time-varying reference: tiled [(K+1)*nx][S] and its staging [B][(K+1)*nx]
    double* d_u0 = nullptr;       // [B][nu]
    int* d_ref_of_internal = nullptr, *d_internal_of_ref = nullptr, *d_value_rows = nullptr, *d_jac_pos = nullptr;
    double* d_values = nullptr, *d_jac = nullptr, *d_eval_out = nullptr;
    // fused peer gather
    int peer_world = 0, peer_rank = 0;
    bool peer_attached = false;
    void* peer_local = nullptr;              // own buffer: [2*world*B] doubles, then [world] arrival counters, then 1 int timeout flag
    void* peer_mapped[MAX_PEERS] = {};       // cudaIpcOpenMemHandle results (null for own rank)
    unsigned long long peer_solves = 0;      // solves launched since attach
    int* d_num_shift = nullptr, *d_shift_plan = nullptr;
    // closed-loop log (b200sqp_closed_loop)
    double *d_loop_x = nullptr, *d_loop_u = nullptr, *d_loop_chi2 = nullptr;
    int32_t* d_loop_status = nullptr;
    int loop_steps = 0;
    PipeArrays pipe{};          // warp-cooperative pipeline (large stage blocks), allocated when the kernel set has one and the
    bool use_pipeline = false;  // structure is eligible
    PipeArraysF32 pipe32{};     // its reduced-precision twin (b200sqp_set_precision), allocated on first use
    int precision = B200SQP_PRECISION_F64;
    bool pipeline_enabled = true;
    long long* d_phase_cycles = nullptr;
    int phase_blocks = 0;
    std::vector<void*> allocations;

    template <class T>
    cudaError_t alloc(T** p, size_t count)
    {
        cudaError_t e = cudaMalloc((void**)p, sizeof(T) * (count ? count : 1));
        if (e == cudaSuccess)
        {
            allocations.push_back(*p);
            e = cudaMemsetAsync(*p, 0, sizeof(T) * (count ? count : 1), stream);
        }
        return e;
    }
};

namespace {

void fillDeviceOcp(const Structure& s, int B, int S, DeviceOcp& P)
{
    std::memset(&P, 0, sizeof(P));
    const b200sqp_ocp& o = s.ocp;
    P.K                  = s.K;
    P.B                  = B;
    P.S                  = S;
    P.stage_cost         = o.stage_cost;
    P.final_cost         = (o.final_cost == 1 && !s.xfFullyFixed()) ? 1 : 0;
    P.tcost_every_interval = s.vt;  // MinimumTime: a dt term on every interval iff !single_dt (minimum_time.h:49)
    for (int i = 0; i < s.nx; ++i)
    {
        P.xf_fixed[i]  = o.xf_fixed[i] ? 1 : 0;
        P.x_lb[i]      = o.x_lb[i];
        P.x_ub[i]      = o.x_ub[i];
        P.x_bounded[i] = (o.x_lb[i] > -kCorboInf || o.x_ub[i] < kCorboInf) ? 1 : 0;
        // QuadraticFormCost::setWeightQ -> cwiseSqrt (quadratic_cost.cpp:62), QuadraticFinalStateCost::setWeightQf (final_state_cost.cpp:66);
        // with full matrices: the diagonal mode of structure.cpp weightSqrt, or unused (dense factors travel separately)
        P.q_sqrt[i]    = (!s.q_w.dense && (int)s.q_w.w.size() == s.nx) ? s.q_w.w[i] : std::sqrt(o.q_diag[i]);
        P.qf_sqrt[i]   = (!s.qf_w.dense && (int)s.qf_w.w.size() == s.nx) ? s.qf_w.w[i] : std::sqrt(o.qf_diag[i]);
    }
    for (int i = 0; i < s.nu; ++i)
    {
        P.u_lb[i]      = o.u_lb[i];
        P.u_ub[i]      = o.u_ub[i];
        P.u_bounded[i] = (o.u_lb[i] > -kCorboInf || o.u_ub[i] < kCorboInf) ? 1 : 0;
        P.r_sqrt[i]    = (!s.r_w.dense && (int)s.r_w.w.size() == s.nu) ? s.r_w.w[i] : std::sqrt(o.r_diag[i]);
    }
    P.q_dense  = s.q_w.dense ? 1 : 0;
    P.r_dense  = s.r_w.dense ? 1 : 0;
    P.qf_dense = (s.qf_w.dense && P.final_cost) ? 1 : 0;
    P.final_constraint = s.xfFullyFixed() ? 0 : o.final_constraint;
    for (int i = 0; i < s.nx; ++i)
    {
        P.term_xref[i] = o.term_xref[i];
        P.term_s[i]    = o.term_s_diag[i];
    }
    P.term_gamma = o.term_gamma;
    P.dt_bounded = (s.vt && (o.dt_lb > -kCorboInf || o.dt_ub < kCorboInf)) ? 1 : 0;
    P.dt_ref     = o.dt_ref;
    P.dt_lb      = o.dt_lb;
    P.dt_ub      = o.dt_ub;
    P.tcost_w    = std::sqrt((double)(o.n_grid - 1));  // MinimumTime::update, lsq form (minimum_time.h:56-66)
    for (int i = 0; i < B200SQP_MAX_DYN_PARAMS; ++i) P.dyn.p[i] = o.dyn_params[i];
    prepareDynParams(P.dyn);
}

// device arrays of the warp-cooperative pipeline in precision Real; the small per-instance LM state arrays are shared between the
// fp64 and the fp32 set (only one of them is in use during a solve)
template <class Real>
cudaError_t allocPipeArrays(b200sqp_handle h, PipeArraysT<Real>& pa)
{
    const size_t K = h->s.K, nb = h->s.nb, nx = h->s.nx, S = h->S;
    const size_t nd = nb * (nb + 1) / 2, ne = nb * nx, BK = (size_t)h->B * K, nxxp = paddedTriangle((int)nx, (int)sizeof(Real));
    cudaError_t e = cudaSuccess;
    auto A = [&](auto** p, size_t cnt) {
        if (e == cudaSuccess) e = h->alloc(p, cnt);
    };
    A(&pa.D, BK * nd);
    A(&pa.E, BK * ne);
    A(&pa.DA, BK * nxxp);
    A(&pa.gA, BK * nx);
    A(&pa.g, BK * nb);
    A(&pa.L, BK * nd);
    A(&pa.W, BK * ne);
    A(&pa.y, BK * nb);
    if (h->pipe.cpart && (void*)&pa != (void*)&h->pipe)
    {
        pa.cpart = h->pipe.cpart, pa.mu_acc = h->pipe.mu_acc, pa.last_values = h->pipe.last_values, pa.dn2 = h->pipe.dn2, pa.dq = h->pipe.dq;
        pa.v = h->pipe.v, pa.k_outer = h->pipe.k_outer, pa.flags = h->pipe.flags, pa.any = h->pipe.any;
        return e;
    }
    A(&pa.cpart, K * S);
    A(&pa.mu_acc, S);
    A(&pa.last_values, S);
    A(&pa.dn2, S);
    A(&pa.dq, S);
    A(&pa.v, S);
    A(&pa.k_outer, S);
    A(&pa.flags, S);
    A(&pa.any, 2);
    return e;
}

int checkHandle(b200sqp_handle h)
{
    if (!h) return fail(B200SQP_ERR_INVALID, "null handle");
    cudaError_t e = cudaSetDevice(h->device);
    if (e != cudaSuccess) return fail(B200SQP_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
    return B200SQP_OK;
}

// resetWeights / adaptWeights (levenberg_marquardt_sparse.cpp:83-86, 264-287)
void updateWeights(b200sqp_handle h, const b200sqp_lm_options& o, bool new_run)
{
    if (new_run || !h->weights_initialised)
    {
        h->st.w_eq   = o.weight_eq;
        h->st.w_ineq = o.weight_ineq;
        h->st.w_b    = o.weight_bounds;
    }
    else
    {
        h->st.w_eq *= o.adapt_factor_eq;
        if (h->st.w_eq > o.adapt_max_eq) h->st.w_eq = o.adapt_max_eq;
        h->st.w_ineq *= o.adapt_factor_ineq;
        if (h->st.w_ineq > o.adapt_max_ineq) h->st.w_ineq = o.adapt_max_ineq;
        h->st.w_b *= o.adapt_factor_bounds;
        if (h->st.w_b > o.adapt_max_bounds) h->st.w_b = o.adapt_max_bounds;
    }
    h->weights_initialised = true;
}

int ensureTrace(b200sqp_handle h, int iterations)
{
    if (iterations <= h->max_iterations && h->st.trace) return B200SQP_OK;
    double* t = nullptr;
    CUDA_TRY(h->alloc(&t, (size_t)(iterations + 1) * h->S));
    h->st.trace       = t;
    h->max_iterations = iterations;
    return B200SQP_OK;
}

}  // namespace

extern "C" {

const char* b200sqp_last_error(void) { return g_last_error.c_str(); }

int b200sqp_device_available(void)
{
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
    {
        cudaGetLastError();
        return 0;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, 0) != cudaSuccess) return 0;
    return prop.major >= 10 ? 1 : 0;
}

int b200sqp_dims_of(const b200sqp_ocp* ocp, b200sqp_dims* out)
{
    if (!ocp || !out) return fail(B200SQP_ERR_INVALID, "null argument");
    Structure s;
    std::string err;
    int rc = buildStructure(*ocp, s, err);
    if (rc != B200SQP_OK) return fail(rc, err);
    *out = s.dims;
    return B200SQP_OK;
}

int b200sqp_vertex_indices(const b200sqp_ocp* ocp, int32_t* x_idx, int32_t* u_idx, int32_t* dt_idx)
{
    if (!ocp) return fail(B200SQP_ERR_INVALID, "null argument");
    Structure s;
    std::string err;
    int rc = buildStructure(*ocp, s, err);
    if (rc != B200SQP_OK) return fail(rc, err);
    if (x_idx) std::memcpy(x_idx, s.x_idx.data(), sizeof(int32_t) * s.x_idx.size());
    if (u_idx) std::memcpy(u_idx, s.u_idx.data(), sizeof(int32_t) * s.u_idx.size());
    if (dt_idx) std::memcpy(dt_idx, s.dt_idx.data(), sizeof(int32_t) * s.dt_idx.size());
    return B200SQP_OK;
}

int b200sqp_edge_indices(const b200sqp_ocp* ocp, int32_t* state_cost_idx, int32_t* control_cost_idx, int32_t* dt_cost_idx, int32_t* dynamics_idx,
                         int32_t* final_cost_idx)
{
    if (!ocp) return fail(B200SQP_ERR_INVALID, "null argument");
    Structure s;
    std::string err;
    int rc = buildStructure(*ocp, s, err);
    if (rc != B200SQP_OK) return fail(rc, err);
    if (state_cost_idx) std::memcpy(state_cost_idx, s.state_cost_idx.data(), sizeof(int32_t) * s.state_cost_idx.size());
    if (control_cost_idx) std::memcpy(control_cost_idx, s.control_cost_idx.data(), sizeof(int32_t) * s.control_cost_idx.size());
    if (dt_cost_idx) std::memcpy(dt_cost_idx, s.dt_cost_idx.data(), sizeof(int32_t) * s.dt_cost_idx.size());
    if (dynamics_idx) std::memcpy(dynamics_idx, s.dynamics_idx.data(), sizeof(int32_t) * s.dynamics_idx.size());
    if (final_cost_idx) *final_cost_idx = s.final_cost_idx;
    return B200SQP_OK;
}

int b200sqp_dt_equality_indices(const b200sqp_ocp* ocp, int32_t* dt_eq_idx)
{
    if (!ocp || !dt_eq_idx) return fail(B200SQP_ERR_INVALID, "null argument");
    Structure s;
    std::string err;
    int rc = buildStructure(*ocp, s, err);
    if (rc != B200SQP_OK) return fail(rc, err);
    std::memcpy(dt_eq_idx, s.dt_eq_idx.data(), sizeof(int32_t) * s.dt_eq_idx.size());
    return B200SQP_OK;
}

int b200sqp_final_constraint_indices(const b200sqp_ocp* ocp, int32_t* eq_idx, int32_t* ineq_idx)
{
    if (!ocp) return fail(B200SQP_ERR_INVALID, "null argument");
    Structure s;
    std::string err;
    int rc = buildStructure(*ocp, s, err);
    if (rc != B200SQP_OK) return fail(rc, err);
    if (eq_idx) *eq_idx = s.final_eq_idx;
    if (ineq_idx) *ineq_idx = s.final_ineq_idx;
    return B200SQP_OK;
}

int b200sqp_jacobian_pattern(const b200sqp_ocp* ocp, int32_t* col_ptr, int32_t* row_idx)
{
    if (!ocp) return fail(B200SQP_ERR_INVALID, "null argument");
    Structure s;
    std::string err;
    int rc = buildStructure(*ocp, s, err);
    if (rc != B200SQP_OK) return fail(rc, err);
    if (col_ptr) std::memcpy(col_ptr, s.col_ptr.data(), sizeof(int32_t) * s.col_ptr.size());
    if (row_idx) std::memcpy(row_idx, s.row_idx.data(), sizeof(int32_t) * s.row_idx.size());
    return B200SQP_OK;
}

int b200sqp_create(const b200sqp_ocp* ocp, int32_t batch, int32_t device, b200sqp_handle* out)
{
    if (!ocp || !out || batch < 1) return fail(B200SQP_ERR_INVALID, "null argument or batch < 1");
    *out = nullptr;
    Structure s;
    std::string err;
    int rc = buildStructure(*ocp, s, err);
    if (rc != B200SQP_OK) return fail(rc, err);
    // kernel key: 0 fixed dt, 1 one free dt per interval, 2 the same with equality edges between consecutive dt vertices
    const int vt_key    = s.vt + s.dteq;
    const KernelSet* ks = s.denseCost() ? findDenseCostKernels(ocp->dynamics, s.defect, vt_key) : findKernels(ocp->dynamics, s.defect, vt_key);
    if (!ks)
        return fail(B200SQP_ERR_UNSUPPORTED, s.denseCost() ? "full (non-diagonal) cost weights are not compiled for this (dynamics, defect, grid) combination "
                                                              "(kernels_dense_cost.cu); no CPU fallback"
                                                            : "no device kernel for this (dynamics, defect, grid) combination; no CPU fallback");
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
    {
        cudaGetLastError();
        return fail(B200SQP_ERR_NO_DEVICE, "no CUDA device visible: the LM hot path only exists as sm_100a kernels");
    }
    if (device < 0 || device >= count) return fail(B200SQP_ERR_INVALID, "device index out of range");
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(B200SQP_ERR_NO_DEVICE, "device is not sm_100 class (kernels are compiled for sm_100a only)");

    b200sqp_solver* h = new b200sqp_solver;
    h->s       = s;
    h->kernels = ks;
    h->B       = batch;
    h->S       = (batch + 31) / 32 * 32;
    h->device  = device;
    cudaError_t e = cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreate(&h->ev_begin);
    if (e == cudaSuccess) e = cudaEventCreate(&h->ev_end);
    h->stream = h->own_stream;
    fillDeviceOcp(s, h->B, h->S, h->P);

    const size_t S = h->S, K = s.K, nb = s.nb, nx = s.nx;
    const size_t nd = nb * (nb + 1) / 2, ne = nb * (nx + s.dteq), n = s.dims.n_params;  // sub-diagonal blocks: nb x coupling slots
    DeviceState& st = h->st;
    auto A = [&](auto** p, size_t cnt) {
        if (e == cudaSuccess) e = h->alloc(p, cnt);
    };
    A(&st.z[0], K * nb * S);
    A(&st.z[1], K * nb * S);
    A(&st.x0, nx * S);
    A(&st.xref, nx * S);
    A(&st.D, K * nd * S);
    A(&st.E, K * ne * S);
    A(&st.g, K * nb * S);
    A(&st.dl, K * nb * S);
    A(&st.L, K * nd * S);
    A(&st.W, K * ne * S);
    st.red_blocks = ks->max_threads > 1 ? ks->max_threads - 1 : 1;
    A(&st.Y, ks->max_threads > 1 ? K * ne * S : 1);
    A(&st.red, ks->max_threads > 1 ? (size_t)st.red_blocks * (2 * nd + 2 * ne + 2 * nb) * S : 1);
    A(&st.chi2, S);
    A(&st.mu, S);
    A(&st.rho, S);
    A(&st.status, S);
    A(&st.cur, S);
    A(&st.n_factor, S);
    A(&st.n_reject, S);
    A(&st.n_linearize, S);
    A(&h->d_params, (size_t)h->B * n);
    A(&h->d_x0_host_order, (size_t)h->B * nx);
    A(&h->d_xref_host_order, (size_t)h->B * nx);
    A(&h->d_u0, (size_t)h->B * s.nu);
    A(&h->d_ref_of_internal, s.ref_of_internal.size());
    A(&h->d_internal_of_ref, s.internal_of_ref.size());
    A(&h->d_value_rows, s.value_rows.size());
    A(&h->d_jac_pos, s.jac_pos.size());
    h->use_pipeline = ks->pipeline != nullptr && pipelineEligible(h->P, s.nx);
    if (h->use_pipeline && e == cudaSuccess) e = allocPipeArrays(h, h->pipe);
    st.trace = nullptr;
    if (s.denseCost() && e == cudaSuccess)
    {
        // upper Cholesky factors of the full weights: [nx*nx | nu*nu | nx*nx] (Q, R, Qf), zeros where a weight is diagonal
        std::vector<double> wfull(2 * nx * nx + (size_t)s.nu * s.nu, 0.0);
        if (s.q_w.dense) std::copy(s.q_w.w.begin(), s.q_w.w.end(), wfull.begin());
        if (s.r_w.dense) std::copy(s.r_w.w.begin(), s.r_w.w.end(), wfull.begin() + nx * nx);
        if (s.qf_w.dense) std::copy(s.qf_w.w.begin(), s.qf_w.w.end(), wfull.begin() + nx * nx + (size_t)s.nu * s.nu);
        double* d_w = nullptr;
        e           = h->alloc(&d_w, wfull.size());
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_w, wfull.data(), sizeof(double) * wfull.size(), cudaMemcpyHostToDevice, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);  // wfull is a local
        st.cost_sqrt_full = d_w;
    }
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(h->d_ref_of_internal, s.ref_of_internal.data(), sizeof(int) * s.ref_of_internal.size(), cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(h->d_internal_of_ref, s.internal_of_ref.data(), sizeof(int) * s.internal_of_ref.size(), cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(h->d_value_rows, s.value_rows.data(), sizeof(int) * s.value_rows.size(), cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(h->d_jac_pos, s.jac_pos.data(), sizeof(int) * s.jac_pos.size(), cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess)
    {
        std::string msg = std::string("device allocation failed: ") + cudaGetErrorString(e);
        b200sqp_destroy(h);
        return fail(B200SQP_ERR_CUDA, msg);
    }
    st.w_eq = st.w_ineq = st.w_b = 2.0;
    *out = h;
    return B200SQP_OK;
}

int b200sqp_destroy(b200sqp_handle h)
{
    if (!h) return B200SQP_OK;
    cudaSetDevice(h->device);
    if (h->own_stream) cudaStreamSynchronize(h->own_stream);
    b200sqp_peer_detach(h);
    for (void* p : h->allocations) cudaFree(p);
    if (h->ev_begin) cudaEventDestroy(h->ev_begin);
    if (h->ev_end) cudaEventDestroy(h->ev_end);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    delete h;
    return B200SQP_OK;
}

int b200sqp_get_dims(b200sqp_handle h, b200sqp_dims* out)
{
    if (!h || !out) return fail(B200SQP_ERR_INVALID, "null argument");
    *out = h->s.dims;
    return B200SQP_OK;
}

// Device view of a caller's host buffer if it is pinned (cudaHostAlloc / cudaHostRegister: mapped into the device's address space
// under unified addressing), else null.  Small per-step transfers (start states in; first controls, chi2, status out) then go
// through the kernels that consume / produce them instead of one cudaMemcpyAsync each: five DMA set-ups per step cost more than the
// bytes they move.  Pageable buffers keep the memcpy path.
static void* pinnedView(const void* p)
{
    if (!p) return nullptr;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess)
    {
        cudaGetLastError();
        return nullptr;
    }
    return (a.type == cudaMemoryTypeHost) ? a.devicePointer : nullptr;
}

// fixed goal components follow the reference: _xf.values()[i] = xref[i] (full_discretization_grid_base.cpp:102-106)
static void fillPinnedGoal(b200sqp_handle h)
{
    const int nb = h->s.nb, xo = h->s.nu + h->s.vt;
    unsigned mask = 0;
    for (int i = 0; i < h->s.nx; ++i)
        if (h->s.xfFixed(i)) mask |= 1u << i;
    if (mask)
    {
        launchFillPinned(h->st.xref, h->st.z[0], h->st.z[1], (h->s.K - 1) * nb + xo, h->s.K * nb, h->s.nx, mask, h->B, h->stream);
        h->launches += 1;
    }
}

// start states / references that are already in HBM ([B][nx], host order): the device half of b200sqp_set_problem_data
static int startStatesFromDevice(b200sqp_handle h, const double* d_x0, const double* d_xref /*null = keep the current reference*/)
{
    launchTransposeIn(d_x0, h->s.nx, h->st.x0, h->B, h->S, h->stream);
    h->launches += 1;
    if (d_xref)
    {
        launchTransposeIn(d_xref, h->s.nx, h->st.xref, h->B, h->S, h->stream);
        h->launches += 1;
    }
    // fixed goal components follow the reference: _xf.values()[i] = xref[i] (full_discretization_grid_base.cpp:102-106)
    const int nb = h->s.nb, xo = h->s.nu + h->s.vt;
    unsigned mask = 0;
    for (int i = 0; i < h->s.nx; ++i)
        if (h->s.xfFixed(i)) mask |= 1u << i;
    if (mask)
    {
        launchFillPinned(h->st.xref, h->st.z[0], h->st.z[1], (h->s.K - 1) * nb + xo, h->s.K * nb, h->s.nx, mask, h->B, h->stream);
        h->launches += 1;
    }
    CUDA_TRY(cudaGetLastError());
    return B200SQP_OK;
}

// the device half of b200sqp_warm_start_shift: d_x0_new [B][nx] in HBM
static int warmStartShiftFromDevice(b200sqp_handle h, const double* d_x0_new, int* d_num_shift)
{
    if (h->s.vt)
        return fail(B200SQP_ERR_UNSUPPORTED, "the reference never shifts a NonUniformFiniteDifferencesVariableGrid (isMovingHorizonWarmStartActive() "
                                             "is false there): use mode 1 (keep)");
    if (!h->d_shift_plan) CUDA_TRY(h->alloc(&h->d_shift_plan, (size_t)h->B));
    launchWarmStartShift(d_x0_new, h->st.x0, h->st.z[0], h->st.z[1], h->st.cur, h->s.K, h->s.nx, h->s.nu, h->d_shift_plan, d_num_shift, h->B,
                         h->stream);
    h->launches += 2;
    CUDA_TRY(cudaGetLastError());
    // fixed goal components follow the reference (full_discretization_grid_base.cpp:102-106)
    const int nb = h->s.nb, xo = h->s.nu + h->s.vt;
    unsigned mask = 0;
    for (int i = 0; i < h->s.nx; ++i)
        if (h->s.xfFixed(i)) mask |= 1u << i;
    if (mask)
    {
        launchFillPinned(h->st.xref, h->st.z[0], h->st.z[1], (h->s.K - 1) * nb + xo, h->s.K * nb, h->s.nx, mask, h->B, h->stream);
        h->launches += 1;
    }
    return B200SQP_OK;
}

int b200sqp_set_problem_data(b200sqp_handle h, const double* x0, const double* xref)
{
    int rc = checkHandle(h);
    if (rc) return rc;
    if (!x0) return fail(B200SQP_ERR_INVALID, "x0 is null");
    const size_t bytes = sizeof(double) * (size_t)h->B * h->s.nx;
    // pinned buffers are read in place by the ingest kernel; pageable ones are staged with a copy first
    const double* x0_src   = (const double*)pinnedView(x0);
    const double* xref_src = xref ? (const double*)pinnedView(xref) : nullptr;
    if (!x0_src)
    {
        CUDA_TRY(cudaMemcpyAsync(h->d_x0_host_order, x0, bytes, cudaMemcpyHostToDevice, h->stream));
        x0_src = h->d_x0_host_order;
    }
    if (xref && !xref_src)
    {
        CUDA_TRY(cudaMemcpyAsync(h->d_xref_host_order, xref, bytes, cudaMemcpyHostToDevice, h->stream));
        xref_src = h->d_xref_host_order;
    }
    // an xref argument (or a missing one = zero reference) is a STATIC reference: it ends a time-varying one
    h->st.xref_traj = nullptr;
    launchIngest(x0_src, xref_src, h->s.nx, h->st.x0, h->st.xref, h->d_x0_host_order, h->d_xref_host_order, h->B, h->stream);
    h->launches += 1;
    fillPinnedGoal(h);
    CUDA_TRY(cudaGetLastError());
    return B200SQP_OK;
}

int b200sqp_set_reference_trajectory(b200sqp_handle h, const double* xref_traj)
{
    int rc = checkHandle(h);
    if (rc) return rc;
    if (!xref_traj)
    {
        h->st.xref_traj = nullptr;
        return B200SQP_OK;
    }
    if (h->s.ocp.grid == B200SQP_GRID_MULTIPLE_SHOOTING)
        return fail(B200SQP_ERR_UNSUPPORTED, "time-varying references on the shooting grids: the reference's cold start puts xref(0) into the first "
                                             "shooting node and fixes it there (shooting_grid_base.cpp:259,278), i.e. it ignores the measured state; "
                                             "not mirrored");
    if (h->s.ocp.stage_cost == B200SQP_COST_QUADRATIC_LSQ && h->s.ocp.zero_x_ref)
        return fail(B200SQP_ERR_INVALID, "the descriptor says zero_x_ref = 1 but a reference trajectory is given");
    const int nx = h->s.nx, n_points = h->s.K + 1;
    const size_t row = (size_t)n_points * nx, bytes = sizeof(double) * (size_t)h->B * row;
    if (!h->d_xref_traj)
    {
        CUDA_TRY(h->alloc(&h->d_xref_traj, (size_t)h->S * row));
        CUDA_TRY(h->alloc(&h->d_xref_traj_host_order, (size_t)h->B * row));
    }
    CUDA_TRY(cudaMemcpyAsync(h->d_xref_traj_host_order, xref_traj, bytes, cudaMemcpyHostToDevice, h->stream));
    launchTransposeIn(h->d_xref_traj_host_order, (int)row, h->d_xref_traj, h->B, h->S, h->stream);
    // the static reference array holds the LAST grid point: goal of the initial guess, value of fixed goal components, reference of the
    // final-stage constraint (getReferenceCached(n - 1) everywhere in the reference)
    CUDA_TRY(cudaMemcpy2DAsync(h->d_xref_host_order, sizeof(double) * nx, h->d_xref_traj_host_order + (row - nx), sizeof(double) * row,
                               sizeof(double) * nx, h->B, cudaMemcpyDeviceToDevice, h->stream));
    launchTransposeIn(h->d_xref_host_order, nx, h->st.xref, h->B, h->S, h->stream);
    h->launches += 2;
    h->st.xref_traj = h->d_xref_traj;
    fillPinnedGoal(h);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(h->stream));  // the caller's buffer may be pageable and reused
    return B200SQP_OK;
}

int b200sqp_initialize_trajectories(b200sqp_handle h)
{
    int rc = checkHandle(h);
    if (rc) return rc;
    launchInitTrajectories(h->st.x0, h->st.xref, h->st.xref_traj, h->st.z[0], h->st.cur, h->s.K, h->s.nx, h->s.nu, h->s.vt, h->s.ocp.dt_ref, h->B,
                           h->S, h->stream);
    h->launches += 1;
    CUDA_TRY(cudaGetLastError());
    return B200SQP_OK;
}

int b200sqp_warm_start_shift(b200sqp_handle h, const double* x0_new, int32_t* num_shift)
{
    int rc = checkHandle(h);
    if (rc) return rc;
    if (!x0_new) return fail(B200SQP_ERR_INVALID, "x0_new is null");
    if (h->s.vt)
        return fail(B200SQP_ERR_UNSUPPORTED, "the reference never shifts a NonUniformFiniteDifferencesVariableGrid (isMovingHorizonWarmStartActive() "
                                             "is false there): use mode 1 (keep)");
    const size_t bytes = sizeof(double) * (size_t)h->B * h->s.nx;
    CUDA_TRY(cudaMemcpyAsync(h->d_x0_host_order, x0_new, bytes, cudaMemcpyHostToDevice, h->stream));
    if (num_shift && !h->d_num_shift) CUDA_TRY(h->alloc(&h->d_num_shift, (size_t)h->B));
    rc = warmStartShiftFromDevice(h, h->d_x0_host_order, num_shift ? h->d_num_shift : nullptr);
    if (rc) return rc;
    if (num_shift)
    {
        CUDA_TRY(cudaMemcpyAsync(num_shift, h->d_num_shift, sizeof(int32_t) * h->B, cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(cudaStreamSynchronize(h->stream));
    }
    return B200SQP_OK;
}

int b200sqp_set_params(b200sqp_handle h, const double* params)
{
    int rc = checkHandle(h);
    if (rc) return rc;
    if (!params) return fail(B200SQP_ERR_INVALID, "params is null");
    const int n = h->s.dims.n_params;
    CUDA_TRY(cudaMemcpyAsync(h->d_params, params, sizeof(double) * (size_t)h->B * n, cudaMemcpyHostToDevice, h->stream));
    launchPack(h->d_params, n, h->d_ref_of_internal, h->s.K * h->s.nb, nullptr, h->st.z[0], h->st.cur, h->st.z[1], h->B, h->S, h->stream);
    h->launches += 1;
    CUDA_TRY(cudaGetLastError());
    return B200SQP_OK;
}

int b200sqp_get_params(b200sqp_handle h, double* params)
{
    int rc = checkHandle(h);
    if (rc) return rc;
    if (!params) return fail(B200SQP_ERR_INVALID, "params is null");
    const int n = h->s.dims.n_params;
    launchUnpack(h->st.z[0], h->st.z[1], h->st.cur, h->d_internal_of_ref, n, h->s.K * h->s.nb, h->d_params, h->B, h->S, h->stream);
    h->launches += 1;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(params, h->d_params, sizeof(double) * (size_t)h->B * n, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return B200SQP_OK;
}

int b200sqp_get_first_controls(b200sqp_handle h, double* u0)
{
    int rc = checkHandle(h);
    if (rc) return rc;
    if (!u0) return fail(B200SQP_ERR_INVALID, "u0 is null");
    launchFirstControls(h->st.z[0], h->st.z[1], h->st.cur, h->s.nu, h->s.K * h->s.nb, h->d_u0, h->B, h->S, h->stream);
    h->launches += 1;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(u0, h->d_u0, sizeof(double) * (size_t)h->B * h->s.nu, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return B200SQP_OK;
}

int b200sqp_solve_async(b200sqp_handle h, const b200sqp_lm_options* opts, int32_t new_run)
{
    int rc = checkHandle(h);
    if (rc) return rc;
    if (!opts || opts->iterations < 0) return fail(B200SQP_ERR_INVALID, "bad options");
    rc = ensureTrace(h, opts->iterations);
    if (rc) return rc;
    updateWeights(h, *opts, new_run != 0);
    if (h->peer_attached)
    {
        h->st.peer_parity = (int)(h->peer_solves & 1);  // double-buffered: a fast peer's next solve never overwrites what we still read
        h->peer_solves += 1;
    }
    CUDA_TRY(cudaEventRecord(h->ev_begin, h->stream));
    if (h->use_pipeline && h->pipeline_enabled && !h->st.xref_traj)
    {
        // large stage blocks: warp-cooperative multi-kernel pipeline, host-driven passes (blocks until the batch is solved)
        const bool ok = h->precision == B200SQP_PRECISION_F32 ? h->kernels->pipeline_f32(h->P, h->st, h->pipe32, opts->iterations, h->stream)
                                                              : h->kernels->pipeline(h->P, h->st, h->pipe, opts->iterations, h->stream);
        if (!ok) return fail(B200SQP_ERR_CUDA, "LM pipeline exceeded its pass bound");
        h->launches += 4 * (int64_t)(opts->iterations + 1);
        if (h->peer_attached)
        {
            // the fused kernel's epilogue as a launch of its own: chi2 into every rank's gather buffer + arrival counters
            launchPeerPublish(h->st, h->B, h->stream);
            h->launches += 1;
        }
    }
    else
    {
        h->kernels->solve(h->P, h->st, opts->iterations, h->threads_per_instance, h->solve_flags, h->stream);
        h->launches += 1;
    }
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(h->ev_end, h->stream));
    h->timed = true;
    return B200SQP_OK;
}

int b200sqp_synchronize(b200sqp_handle h)
{
    int rc = checkHandle(h);
    if (rc) return rc;
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return B200SQP_OK;
}

int b200sqp_solve(b200sqp_handle h, const b200sqp_lm_options* opts, int32_t new_run, int32_t* status, double* chi2)
{
    int rc = b200sqp_solve_async(h, opts, new_run);
    if (rc) return rc;
    if (status) CUDA_TRY(cudaMemcpyAsync(status, h->st.status, sizeof(int32_t) * h->B, cudaMemcpyDeviceToHost, h->stream));
    if (chi2) CUDA_TRY(cudaMemcpyAsync(chi2, h->st.chi2, sizeof(double) * h->B, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return B200SQP_OK;
}

// first controls (when u0_out is given), chi2 and status of the last solve to the caller's host buffers: one kernel writes pinned
// buffers in place; pageable ones get a cudaMemcpyAsync each.  Asynchronous on the handle's stream.
static int exportSmallResults(b200sqp_handle h, double* u0_out, double* chi2_out, int32_t* status_out)
{
    double* u0_view   = (double*)pinnedView(u0_out);
    double* chi2_view = (double*)pinnedView(chi2_out);
    int* status_view  = (int*)pinnedView(status_out);
    if (u0_out || chi2_view || status_view)
    {
        launchExport(h->st.z[0], h->st.z[1], h->st.cur, h->s.nu, h->s.K * h->s.nb, h->st.chi2, h->st.status, u0_out ? h->d_u0 : nullptr, u0_view,
                     chi2_view, status_view, h->B, h->stream);
        h->launches += 1;
        CUDA_TRY(cudaGetLastError());
    }
    if (u0_out && !u0_view) CUDA_TRY(cudaMemcpyAsync(u0_out, h->d_u0, sizeof(double) * (size_t)h->B * h->s.nu, cudaMemcpyDeviceToHost, h->stream));
    if (status_out && !status_view) CUDA_TRY(cudaMemcpyAsync(status_out, h->st.status, sizeof(int32_t) * h->B, cudaMemcpyDeviceToHost, h->stream));
    if (chi2_out && !chi2_view) CUDA_TRY(cudaMemcpyAsync(chi2_out, h->st.chi2, sizeof(double) * h->B, cudaMemcpyDeviceToHost, h->stream));
    return B200SQP_OK;
}

int b200sqp_step(b200sqp_handle h, const b200sqp_lm_options* opts, int32_t cold_start, const double* x0, const double* xref, double* params_out,
                 double* chi2_out, int32_t* status_out)
{
    int rc = b200sqp_set_problem_data(h, x0, xref);
    if (rc) return rc;
    if (cold_start)
    {
        rc = b200sqp_initialize_trajectories(h);
        if (rc) return rc;
    }
    rc = b200sqp_solve_async(h, opts, 1);
    if (rc) return rc;
    if (params_out)
    {
        const int n = h->s.dims.n_params;
        launchUnpack(h->st.z[0], h->st.z[1], h->st.cur, h->d_internal_of_ref, n, h->s.K * h->s.nb, h->d_params, h->B, h->S, h->stream);
        h->launches += 1;
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpyAsync(params_out, h->d_params, sizeof(double) * (size_t)h->B * n, cudaMemcpyDeviceToHost, h->stream));
    }
    rc = exportSmallResults(h, nullptr, chi2_out, status_out);
    if (rc) return rc;
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return B200SQP_OK;
}

int b200sqp_mpc_step(b200sqp_handle h, const b200sqp_lm_options* opts, int32_t mode, const double* x0, const double* xref, double* u0_out,
                     double* chi2_out, int32_t* status_out)
{
    if (mode < 0 || mode > 2) return fail(B200SQP_ERR_INVALID, "mode must be 0 (cold), 1 (keep) or 2 (shift)");
    if (!u0_out) return fail(B200SQP_ERR_INVALID, "u0_out is null");
    int rc;
    if (mode == 2)
    {
        // the shift needs the previous start state, so it runs before the start/reference are replaced
        rc = b200sqp_warm_start_shift(h, x0, nullptr);
        if (rc) return rc;
    }
    rc = b200sqp_set_problem_data(h, x0, xref);
    if (rc) return rc;
    if (mode == 0)
    {
        rc = b200sqp_initialize_trajectories(h);
        if (rc) return rc;
    }
    rc = b200sqp_solve_async(h, opts, 1);
    if (rc) return rc;
    rc = exportSmallResults(h, u0_out, chi2_out, status_out);
    if (rc) return rc;
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return B200SQP_OK;
}

int b200sqp_plant_step(int32_t dynamics, const double* dyn_params, int32_t integrator, double dt, int32_t batch, const double* x, const double* u,
                       double* x_next, int32_t device)
{
    if (!dyn_params || !x || !u || !x_next || batch < 1 || (integrator != 0 && integrator != 1)) return fail(B200SQP_ERR_INVALID, "bad argument");
    int nx = 0, nu = 0;
    if (!dynamicsDimensions(dynamics, nx, nu)) return fail(B200SQP_ERR_UNSUPPORTED, "unknown dynamics id (closed functor registry; no CPU fallback)");
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
    {
        cudaGetLastError();
        return fail(B200SQP_ERR_NO_DEVICE, "no CUDA device visible: the plant simulation only exists as sm_100a kernels");
    }
    if (device < 0 || device >= count) return fail(B200SQP_ERR_INVALID, "device index out of range");
    CUDA_TRY(cudaSetDevice(device));
    DynParams dyn;
    for (int i = 0; i < B200SQP_MAX_DYN_PARAMS; ++i) dyn.p[i] = dyn_params[i];
    prepareDynParams(dyn);
    const size_t bx = sizeof(double) * (size_t)batch * nx, bu = sizeof(double) * (size_t)batch * nu;
    double *dx = nullptr, *du = nullptr, *dn = nullptr;
    cudaError_t e = cudaMalloc(&dx, bx);
    if (e == cudaSuccess) e = cudaMalloc(&du, bu);
    if (e == cudaSuccess) e = cudaMalloc(&dn, bx);
    if (e == cudaSuccess) e = cudaMemcpy(dx, x, bx, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(du, u, bu, cudaMemcpyHostToDevice);
    if (e == cudaSuccess)
    {
        launchPlantStep(dynamics, dyn, integrator, dt, batch, dx, du, dn, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(x_next, dn, bx, cudaMemcpyDeviceToHost);
    cudaFree(dx);
    cudaFree(du);
    cudaFree(dn);
    if (e != cudaSuccess) return fail(B200SQP_ERR_CUDA, std::string("plant_step: ") + cudaGetErrorString(e));
    return B200SQP_OK;
}

int b200sqp_closed_loop(b200sqp_handle h, const b200sqp_lm_options* opts, int32_t mode, int32_t integrator, double plant_dt, int32_t steps,
                        const double* x0, const double* xref, double* u_applied, double* x_closed, double* chi2_out, int32_t* status_out)
{
    int rc = checkHandle(h);
    if (rc) return rc;
    if (mode < 0 || mode > 2) return fail(B200SQP_ERR_INVALID, "mode must be 0 (cold every step), 1 (keep) or 2 (shift)");
    if (integrator != 0 && integrator != 1) return fail(B200SQP_ERR_INVALID, "integrator must be 0 (explicit Euler) or 1 (Runge-Kutta 4)");
    if (!x0 || steps < 1 || !(plant_dt > 0.0)) return fail(B200SQP_ERR_INVALID, "bad argument");
    if (mode == 2 && h->s.vt)
        return fail(B200SQP_ERR_UNSUPPORTED, "the reference never shifts a NonUniformFiniteDifferencesVariableGrid: use mode 1 (keep)");
    const int nx = h->s.nx, nu = h->s.nu, B = h->B;
    const size_t sx = (size_t)B * nx, su = (size_t)B * nu;
    // the closed-loop log lives in HBM for the whole run: states [steps+1][B][nx], applied controls [steps][B][nu], chi2 / status [steps][B]
    if (steps > h->loop_steps)
    {
        for (void* p : {(void*)h->d_loop_x, (void*)h->d_loop_u, (void*)h->d_loop_chi2, (void*)h->d_loop_status})
            if (p)
            {
                h->allocations.erase(std::remove(h->allocations.begin(), h->allocations.end(), p), h->allocations.end());
                cudaFree(p);
            }
        h->d_loop_x = h->d_loop_u = h->d_loop_chi2 = nullptr;
        h->d_loop_status                            = nullptr;
        h->loop_steps                               = 0;
        CUDA_TRY(h->alloc(&h->d_loop_x, (size_t)(steps + 1) * sx));
        CUDA_TRY(h->alloc(&h->d_loop_u, (size_t)steps * su));
        CUDA_TRY(h->alloc(&h->d_loop_chi2, (size_t)steps * B));
        CUDA_TRY(h->alloc(&h->d_loop_status, (size_t)steps * B));
        h->loop_steps = steps;
    }
    CUDA_TRY(cudaMemcpyAsync(h->d_loop_x, x0, sizeof(double) * sx, cudaMemcpyHostToDevice, h->stream));
    if (xref)
        CUDA_TRY(cudaMemcpyAsync(h->d_xref_host_order, xref, sizeof(double) * sx, cudaMemcpyHostToDevice, h->stream));
    else
        CUDA_TRY(cudaMemsetAsync(h->st.xref, 0, sizeof(double) * (size_t)h->S * nx, h->stream));
    // The reference keeps time in integer nanoseconds (corbo::Time / Duration over std::chrono, src/core/include/corbo-core/time.h:140,283:
    // fromSec truncates t * 1e9; toSec divides the tick count by 1e9) and the plant integrates over the interval its control buffer
    // hands back, (t + dt) - t in doubles (TimeValueBuffer::getValues, src/systems/src/time_value_buffer.cpp:74: `ts + dt - cur_t` with
    // cur_t = ts), which differs from dt in the last bits for t != 0.  Mirrored here step by step.
    const long long dt_ticks = (long long)(plant_dt * 1e9);
    const double dt_sec      = (double)dt_ticks / 1e9;
    for (int s = 0; s < steps; ++s)
    {
        const double t_sec    = (double)((long long)s * dt_ticks) / 1e9;
        const volatile double t_end = t_sec + dt_sec;  // volatile: keep the two roundings apart
        const double step_dt  = t_end - t_sec;
        const double* d_x = h->d_loop_x + (size_t)s * sx;  // the measurement of this step (plant output = state, no observer dynamics)
        // ---- controller: PredictiveController::step -> StructuredOptimalControlProblem::compute (grid update, solve) + getFirstControlInput
        if (s > 0 && mode == 2)
        {
            rc = warmStartShiftFromDevice(h, d_x, nullptr);  // needs the previous start state: before it is replaced
            if (rc) return rc;
        }
        if (!(s > 0 && mode == 2))  // the shift already replaced the start state by the measurement (and refreshed fixed goal components)
        {
            rc = startStatesFromDevice(h, d_x, (s == 0 && xref) ? h->d_xref_host_order : nullptr);
            if (rc) return rc;
        }
        if (s == 0 || mode == 0)
        {
            rc = b200sqp_initialize_trajectories(h);
            if (rc) return rc;
        }
        rc = b200sqp_solve_async(h, opts, 1);
        if (rc) return rc;
        launchFirstControls(h->st.z[0], h->st.z[1], h->st.cur, nu, h->s.K * h->s.nb, h->d_u0, B, h->S, h->stream);
        // ---- plant: SimulatedPlant::control, one solveIVP over plant_dt with the first control held
        if (!launchPlantStep(h->s.ocp.dynamics, h->P.dyn, integrator, step_dt, B, d_x, h->d_u0, h->d_loop_x + (size_t)(s + 1) * sx,
                             h->d_loop_u + (size_t)s * su, h->st.chi2, h->d_loop_chi2 + (size_t)s * B, h->st.status,
                             h->d_loop_status + (size_t)s * B, h->stream))
            return fail(B200SQP_ERR_UNSUPPORTED, "dynamics id not in the plant registry");
        h->launches += 2;
        CUDA_TRY(cudaGetLastError());
    }
    if (u_applied) CUDA_TRY(cudaMemcpyAsync(u_applied, h->d_loop_u, sizeof(double) * steps * su, cudaMemcpyDeviceToHost, h->stream));
    if (x_closed) CUDA_TRY(cudaMemcpyAsync(x_closed, h->d_loop_x, sizeof(double) * (steps + 1) * sx, cudaMemcpyDeviceToHost, h->stream));
    if (chi2_out) CUDA_TRY(cudaMemcpyAsync(chi2_out, h->d_loop_chi2, sizeof(double) * (size_t)steps * B, cudaMemcpyDeviceToHost, h->stream));
    if (status_out) CUDA_TRY(cudaMemcpyAsync(status_out, h->d_loop_status, sizeof(int32_t) * (size_t)steps * B, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return B200SQP_OK;
}

int b200sqp_evaluate(b200sqp_handle h, double weight_eq, double weight_ineq, double weight_bounds, double* values, double* jac_values)
{
    int rc = checkHandle(h);
    if (rc) return rc;
    const b200sqp_dims& d = h->s.dims;
    const size_t m = (size_t)d.m_lsq + d.m_eq + d.m_ineq + d.m_bounds, nnz = d.nnz_jacobian;
    if (!h->d_values)
    {
        CUDA_TRY(h->alloc(&h->d_values, m * h->S));
        CUDA_TRY(h->alloc(&h->d_jac, nnz * h->S));
        CUDA_TRY(h->alloc(&h->d_eval_out, (size_t)h->B * (m > nnz ? m : nnz)));
    }
    CUDA_TRY(cudaMemsetAsync(h->d_values, 0, sizeof(double) * m * h->S, h->stream));
    CUDA_TRY(cudaMemsetAsync(h->d_jac, 0, sizeof(double) * nnz * h->S, h->stream));
    DeviceState st = h->st;
    st.w_eq        = weight_eq;
    st.w_ineq      = weight_ineq;
    st.w_b         = weight_bounds;
    h->kernels->evaluate(h->P, st, h->d_values, jac_values ? h->d_jac : nullptr, h->d_value_rows, h->d_jac_pos, h->s.values_per_interval,
                         h->s.jac_per_interval, h->stream);
    h->launches += 1;
    CUDA_TRY(cudaGetLastError());
    if (values)
    {
        launchTransposeOut(h->d_values, (int)m, h->d_eval_out, h->B, h->S, h->stream);
        h->launches += 1;
        CUDA_TRY(cudaMemcpyAsync(values, h->d_eval_out, sizeof(double) * (size_t)h->B * m, cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(cudaStreamSynchronize(h->stream));
    }
    if (jac_values)
    {
        launchTransposeOut(h->d_jac, (int)nnz, h->d_eval_out, h->B, h->S, h->stream);
        h->launches += 1;
        CUDA_TRY(cudaMemcpyAsync(jac_values, h->d_eval_out, sizeof(double) * (size_t)h->B * nnz, cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(cudaStreamSynchronize(h->stream));
    }
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return B200SQP_OK;
}

int b200sqp_linearize_dynamics(int32_t dynamics, const double* dyn_params, int32_t method, int32_t batch, const double* x, const double* u,
                               double* A, double* Bm, int32_t device)
{
    if (!dyn_params || !x || !u || batch < 1 || (method != 0 && method != 1) || (!A && !Bm)) return fail(B200SQP_ERR_INVALID, "bad argument");
    b200sqp_ocp probe;
    std::memset(&probe, 0, sizeof(probe));
    int nx = 0, nu = 0;
    if (!dynamicsDimensions(dynamics, nx, nu)) return fail(B200SQP_ERR_UNSUPPORTED, "unknown dynamics id (closed functor registry; no CPU fallback)");
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
    {
        cudaGetLastError();
        return fail(B200SQP_ERR_NO_DEVICE, "no CUDA device visible: the linearisation only exists as sm_100a kernels");
    }
    if (device < 0 || device >= count) return fail(B200SQP_ERR_INVALID, "device index out of range");
    CUDA_TRY(cudaSetDevice(device));
    DynParams dyn;
    for (int i = 0; i < B200SQP_MAX_DYN_PARAMS; ++i) dyn.p[i] = dyn_params[i];
    prepareDynParams(dyn);
    const size_t bx = sizeof(double) * (size_t)batch * nx, bu = sizeof(double) * (size_t)batch * nu;
    const size_t bA = sizeof(double) * (size_t)batch * nx * nx, bB = sizeof(double) * (size_t)batch * nx * nu;
    double *dx = nullptr, *du = nullptr, *dA = nullptr, *dB = nullptr;
    cudaError_t e = cudaMalloc(&dx, bx);
    if (e == cudaSuccess) e = cudaMalloc(&du, bu);
    if (e == cudaSuccess && A) e = cudaMalloc(&dA, bA);
    if (e == cudaSuccess && Bm) e = cudaMalloc(&dB, bB);
    if (e == cudaSuccess) e = cudaMemcpy(dx, x, bx, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(du, u, bu, cudaMemcpyHostToDevice);
    if (e == cudaSuccess)
    {
        launchLinearizeDynamics(dynamics, dyn, method, batch, dx, du, dA, dB, nullptr);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess && A) e = cudaMemcpy(A, dA, bA, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && Bm) e = cudaMemcpy(Bm, dB, bB, cudaMemcpyDeviceToHost);
    cudaFree(dx);
    cudaFree(du);
    cudaFree(dA);
    cudaFree(dB);
    if (e != cudaSuccess) return fail(B200SQP_ERR_CUDA, std::string("linearize_dynamics: ") + cudaGetErrorString(e));
    return B200SQP_OK;
}

int b200sqp_dynamics_hessian(int32_t dynamics, const double* dyn_params, int32_t method, int32_t batch, const double* x, const double* u,
                             const double* multipliers, double* H, int32_t device)
{
    if (!dyn_params || !x || !u || !H || batch < 1 || (method != 0 && method != 1)) return fail(B200SQP_ERR_INVALID, "bad argument");
    int nx = 0, nu = 0;
    if (!dynamicsDimensions(dynamics, nx, nu)) return fail(B200SQP_ERR_UNSUPPORTED, "unknown dynamics id (closed functor registry; no CPU fallback)");
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
    {
        cudaGetLastError();
        return fail(B200SQP_ERR_NO_DEVICE, "no CUDA device visible: the finite-difference Hessian only exists as sm_100a kernels");
    }
    if (device < 0 || device >= count) return fail(B200SQP_ERR_INVALID, "device index out of range");
    CUDA_TRY(cudaSetDevice(device));
    DynParams dyn;
    for (int i = 0; i < B200SQP_MAX_DYN_PARAMS; ++i) dyn.p[i] = dyn_params[i];
    prepareDynParams(dyn);
    const int nz    = nx + nu;
    const size_t bx = sizeof(double) * (size_t)batch * nx, bu = sizeof(double) * (size_t)batch * nu, bh = sizeof(double) * (size_t)batch * nz * nz;
    double *dx = nullptr, *du = nullptr, *dm = nullptr, *dh = nullptr;
    cudaError_t e = cudaMalloc(&dx, bx);
    if (e == cudaSuccess) e = cudaMalloc(&du, bu);
    if (e == cudaSuccess && multipliers) e = cudaMalloc(&dm, bx);
    if (e == cudaSuccess) e = cudaMalloc(&dh, bh);
    if (e == cudaSuccess) e = cudaMemcpy(dx, x, bx, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(du, u, bu, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && multipliers) e = cudaMemcpy(dm, multipliers, bx, cudaMemcpyHostToDevice);
    if (e == cudaSuccess)
    {
        launchDynamicsHessian(dynamics, dyn, method, batch, dx, du, dm, dh, nullptr);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(H, dh, bh, cudaMemcpyDeviceToHost);
    cudaFree(dx);
    cudaFree(du);
    cudaFree(dm);
    cudaFree(dh);
    if (e != cudaSuccess) return fail(B200SQP_ERR_CUDA, std::string("dynamics_hessian: ") + cudaGetErrorString(e));
    return B200SQP_OK;
}

int b200sqp_get_statistics(b200sqp_handle h, int32_t* inner_passes, int32_t* rejects, int32_t* relinearizations, double* mu, double* rho)
{
    int rc = checkHandle(h);
    if (rc) return rc;
    const size_t bi = sizeof(int32_t) * h->B, bd = sizeof(double) * h->B;
    if (inner_passes) CUDA_TRY(cudaMemcpyAsync(inner_passes, h->st.n_factor, bi, cudaMemcpyDeviceToHost, h->stream));
    if (rejects) CUDA_TRY(cudaMemcpyAsync(rejects, h->st.n_reject, bi, cudaMemcpyDeviceToHost, h->stream));
    if (relinearizations) CUDA_TRY(cudaMemcpyAsync(relinearizations, h->st.n_linearize, bi, cudaMemcpyDeviceToHost, h->stream));
    if (mu) CUDA_TRY(cudaMemcpyAsync(mu, h->st.mu, bd, cudaMemcpyDeviceToHost, h->stream));
    if (rho) CUDA_TRY(cudaMemcpyAsync(rho, h->st.rho, bd, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return B200SQP_OK;
}

int b200sqp_get_chi2_trace(b200sqp_handle h, double* trace, int32_t iterations)
{
    int rc = checkHandle(h);
    if (rc) return rc;
    if (!trace || !h->st.trace || iterations > h->max_iterations) return fail(B200SQP_ERR_INVALID, "no trace of that length recorded");
    std::vector<double> tmp((size_t)(iterations + 1) * h->S);
    CUDA_TRY(cudaMemcpyAsync(tmp.data(), h->st.trace, sizeof(double) * tmp.size(), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    for (int i = 0; i < h->B; ++i)
        for (int k = 0; k <= iterations; ++k) trace[(size_t)i * (iterations + 1) + k] = tmp[(size_t)k * h->S + i];
    return B200SQP_OK;
}

int b200sqp_last_solve_ms(b200sqp_handle h, float* ms)
{
    int rc = checkHandle(h);
    if (rc) return rc;
    if (!ms || !h->timed) return fail(B200SQP_ERR_INVALID, "no solve recorded");
    CUDA_TRY(cudaEventSynchronize(h->ev_end));
    CUDA_TRY(cudaEventElapsedTime(ms, h->ev_begin, h->ev_end));
    return B200SQP_OK;
}

int b200sqp_launch_count(b200sqp_handle h, int64_t* launches)
{
    if (!h || !launches) return fail(B200SQP_ERR_INVALID, "null argument");
    *launches = h->launches;
    return B200SQP_OK;
}

int b200sqp_device_pointers(b200sqp_handle h, void** chi2, void** status, void** x0)
{
    if (!h) return fail(B200SQP_ERR_INVALID, "null handle");
    if (chi2) *chi2 = h->st.chi2;
    if (status) *status = h->st.status;
    if (x0) *x0 = h->d_x0_host_order;
    return B200SQP_OK;
}

int b200sqp_set_threads_per_instance(b200sqp_handle h, int32_t threads)
{
    if (!h || threads < -1 || threads > 8) return fail(B200SQP_ERR_INVALID, "threads per instance must be 0 (auto), 1, 2, 4, 8, or -1");
    // -1: force the fused kernel where the warp-cooperative pipeline would be chosen (comparison / debugging)
    h->pipeline_enabled = threads != -1;
    if (threads < 0) threads = 0;
    h->threads_per_instance = threads;
    return B200SQP_OK;
}

int b200sqp_set_precision(b200sqp_handle h, int32_t precision)
{
    int rc = checkHandle(h);
    if (rc) return rc;
    if (precision != B200SQP_PRECISION_F64 && precision != B200SQP_PRECISION_F32) return fail(B200SQP_ERR_INVALID, "unknown precision");
    if (precision == B200SQP_PRECISION_F32)
    {
        if (!h->use_pipeline || !h->kernels->pipeline_f32)
            return fail(B200SQP_ERR_UNSUPPORTED, "the fp32 variant exists for structures that run the warp-cooperative pipeline (large stage blocks: "
                                                 "quadrotor on the fixed-dt grid); everything else is fp64 only -- the reference's 1e-9 central "
                                                 "differences do not exist in fp32");
        if (!h->pipe32.D)
        {
            cudaError_t e = allocPipeArrays(h, h->pipe32);
            if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
            if (e != cudaSuccess) return fail(B200SQP_ERR_CUDA, std::string("fp32 pipeline arrays: ") + cudaGetErrorString(e));
        }
    }
    h->precision = precision;
    return B200SQP_OK;
}

int b200sqp_measure_fp64_peak(int32_t device, double* tflops)
{
    if (!tflops) return fail(B200SQP_ERR_INVALID, "null argument");
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
    {
        cudaGetLastError();
        return fail(B200SQP_ERR_NO_DEVICE, "no CUDA device visible");
    }
    if (device < 0 || device >= count) return fail(B200SQP_ERR_INVALID, "device index out of range");
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    const double v = measureFp64PeakTflops(prop.multiProcessorCount, nullptr);
    if (v < 0) return fail(B200SQP_ERR_CUDA, "fp64 peak measurement failed");
    *tflops = v;
    return B200SQP_OK;
}

int b200sqp_set_feature_set(b200sqp_handle h, int32_t general)
{
    if (!h) return fail(B200SQP_ERR_INVALID, "null handle");
    h->solve_flags = general ? (h->solve_flags | SOLVE_FORCE_GENERAL_FEATURES) : (h->solve_flags & ~SOLVE_FORCE_GENERAL_FEATURES);
    return B200SQP_OK;
}

static size_t peerChi2Bytes(b200sqp_handle h) { return sizeof(double) * 2 * (size_t)h->peer_world * h->B; }

int b200sqp_peer_export(b200sqp_handle h, int32_t world, int32_t rank, void* ipc_handle_out)
{
    int rc = checkHandle(h);
    if (rc) return rc;
    if (!ipc_handle_out || world < 1 || world > MAX_PEERS || rank < 0 || rank >= world) return fail(B200SQP_ERR_INVALID, "bad world/rank");
    static_assert(sizeof(cudaIpcMemHandle_t) == B200SQP_IPC_HANDLE_BYTES, "IPC handle size");
    if (h->peer_local) return fail(B200SQP_ERR_INVALID, "peer buffer already exported");
    h->peer_world = world;
    h->peer_rank  = rank;
    const size_t bytes = peerChi2Bytes(h) + sizeof(unsigned long long) * world + 64;
    CUDA_TRY(cudaMalloc(&h->peer_local, bytes));  // a dedicated allocation: IPC handles cover whole allocations
    CUDA_TRY(cudaMemset(h->peer_local, 0, bytes));
    cudaIpcMemHandle_t handle;
    CUDA_TRY(cudaIpcGetMemHandle(&handle, h->peer_local));
    std::memcpy(ipc_handle_out, &handle, sizeof(handle));
    return B200SQP_OK;
}

int b200sqp_peer_attach(b200sqp_handle h, const void* ipc_handles)
{
    int rc = checkHandle(h);
    if (rc) return rc;
    if (!ipc_handles || !h->peer_local) return fail(B200SQP_ERR_INVALID, "call b200sqp_peer_export first");
    // a second attach would leak the IPC mappings and reset the solve count while the peers' arrival counters keep counting
    if (h->peer_attached) return fail(B200SQP_ERR_INVALID, "already attached: b200sqp_peer_detach first");
    const size_t chi2_bytes = peerChi2Bytes(h);
    for (int r = 0; r < h->peer_world; ++r)
    {
        void* base = h->peer_local;
        if (r != h->peer_rank)
        {
            cudaIpcMemHandle_t handle;
            std::memcpy(&handle, (const char*)ipc_handles + (size_t)r * sizeof(handle), sizeof(handle));
            CUDA_TRY(cudaIpcOpenMemHandle(&h->peer_mapped[r], handle, cudaIpcMemLazyEnablePeerAccess));
            base = h->peer_mapped[r];
        }
        h->st.peer_chi2[r]     = (double*)base;
        h->st.peer_arrivals[r] = (unsigned long long*)((char*)base + chi2_bytes);
    }
    h->st.peer_world  = h->peer_world;
    h->st.peer_rank   = h->peer_rank;
    h->st.peer_parity = 0;
    h->peer_solves    = 0;
    h->peer_attached  = true;
    return B200SQP_OK;
}

int b200sqp_peer_wait(b200sqp_handle h)
{
    int rc = checkHandle(h);
    if (rc) return rc;
    if (!h->peer_attached || h->peer_solves == 0) return fail(B200SQP_ERR_INVALID, "not attached, or no solve launched since attach");
    const size_t chi2_bytes = peerChi2Bytes(h);
    const unsigned long long* arrivals = (const unsigned long long*)((char*)h->peer_local + chi2_bytes);
    int* timed_out                     = (int*)((char*)h->peer_local + chi2_bytes + sizeof(unsigned long long) * h->peer_world);
    const unsigned long long blocks    = (unsigned long long)((h->B + 31) / 32);
    launchPeerWait(arrivals, h->peer_world, h->peer_solves * blocks, 2000000000ull /* 2 s */, timed_out, h->stream);
    h->launches += 1;
    CUDA_TRY(cudaGetLastError());
    return B200SQP_OK;
}

int b200sqp_peer_gathered(b200sqp_handle h, void** chi2_all)
{
    if (!h || !chi2_all || !h->peer_attached || h->peer_solves == 0) return fail(B200SQP_ERR_INVALID, "not attached, or no solve yet");
    const int parity = (int)((h->peer_solves - 1) & 1);
    *chi2_all        = (double*)h->peer_local + (size_t)parity * h->peer_world * h->B;
    return B200SQP_OK;
}

int b200sqp_peer_status(b200sqp_handle h, int32_t* timed_out)
{
    int rc = checkHandle(h);
    if (rc) return rc;
    if (!timed_out || !h->peer_attached) return fail(B200SQP_ERR_INVALID, "not attached");
    const size_t off = peerChi2Bytes(h) + sizeof(unsigned long long) * h->peer_world;
    CUDA_TRY(cudaMemcpyAsync(timed_out, (char*)h->peer_local + off, sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return B200SQP_OK;
}

int b200sqp_peer_detach(b200sqp_handle h)
{
    if (!h) return B200SQP_OK;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    for (int r = 0; r < MAX_PEERS; ++r)
    {
        if (h->peer_mapped[r]) cudaIpcCloseMemHandle(h->peer_mapped[r]);
        h->peer_mapped[r]      = nullptr;
        h->st.peer_chi2[r]     = nullptr;
        h->st.peer_arrivals[r] = nullptr;
    }
    h->st.peer_world = 0;
    h->peer_attached = false;
    if (h->peer_local) cudaFree(h->peer_local);
    h->peer_local = nullptr;
    return B200SQP_OK;
}

int b200sqp_set_phase_profile(b200sqp_handle h, int32_t enable)
{
    int rc = checkHandle(h);
    if (rc) return rc;
    if (enable && !h->d_phase_cycles)
    {
        h->phase_blocks = (h->B + 31) / 32;
        CUDA_TRY(h->alloc(&h->d_phase_cycles, (size_t)h->phase_blocks * 5));
    }
    h->st.phase_cycles = enable ? h->d_phase_cycles : nullptr;
    return B200SQP_OK;
}

int b200sqp_get_phase_cycles(b200sqp_handle h, double* mean_cycles)
{
    int rc = checkHandle(h);
    if (rc) return rc;
    if (!mean_cycles || !h->st.phase_cycles) return fail(B200SQP_ERR_INVALID, "phase profile is off");
    std::vector<long long> tmp((size_t)h->phase_blocks * 5);
    CUDA_TRY(cudaMemcpyAsync(tmp.data(), h->d_phase_cycles, sizeof(long long) * tmp.size(), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    for (int q = 0; q < 5; ++q)
    {
        double s = 0.0;
        for (int b = 0; b < h->phase_blocks; ++b) s += (double)tmp[(size_t)b * 5 + q];
        mean_cycles[q] = s / h->phase_blocks;
    }
    return B200SQP_OK;
}

int b200sqp_set_stream(b200sqp_handle h, void* cuda_stream)
{
    if (!h) return fail(B200SQP_ERR_INVALID, "null handle");
    h->stream = cuda_stream ? (cudaStream_t)cuda_stream : h->own_stream;
    return B200SQP_OK;
}


// The buckets of the grid-adaptation front-end solve concurrently on one stream each.  CUDA maps streams onto
// CUDA_DEVICE_MAX_CONNECTIONS hardware queues (default 8); streams that share a queue pick up false dependencies, and 30 latency-bound
// launches then take 2.3x as long as with a queue each (measured).  The variable is read when the CUDA context is
// created, so it is set -- unless the user has set it -- when this library is loaded: load the library before the first CUDA call.
__attribute__((constructor)) static void b200sqpRequestHardwareQueues() { setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", /*overwrite=*/0); }

/* ---- grid adaptation front-end ---------------------------------------------------------------------------
 * Per-instance grid size for the time-optimal grid.  The batch is bucketed by grid size N: bucket N is a solver handle of the same OCP with
 * n_grid = N (created on first use, sized for the whole batch) whose first count[N] slots are occupied.  The grid side of an OCP
 * iteration (decision, insertion / removal of a grid point, move to the bucket of the new size) runs in adapt_kernels.cu; the host only
 * turns the per-instance decisions (4 bytes each) into slot assignments.  Buckets solve concurrently on their own streams. */
struct b200sqp_adaptive
{
    b200sqp_ocp ocp{};
    int B = 0, device = 0, n_min = 0, n_max = 0, n_lo = 0, n_hi = 0, nx = 0, nu = 0;
    double hyst = 0.1;
    int warm_start = 1;
    std::vector<b200sqp_handle> bucket;
    std::vector<int> count;                  // occupied slots per bucket
    std::vector<int> bucket_of, slot_of;     // per instance
    std::vector<AdaptBucketView> views;
    AdaptBucketView* d_views = nullptr;
    int *d_plan = nullptr, *hp_plan = nullptr;          // [5][B] (device / pinned host)
    int *d_decision = nullptr, *hp_decision = nullptr;  // [B]
    double *d_x0_master = nullptr, *d_xref_master = nullptr, *d_u0 = nullptr, *d_chi2 = nullptr;
    int* d_status = nullptr;
    double *d_tx = nullptr, *d_tu = nullptr, *d_tdt = nullptr;  // trajectory export staging
    int* d_tn = nullptr;
    int export_cap = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_ready = nullptr;
    bool first_run = true, weights_initialised = false;
    double w_eq = 2, w_ineq = 2, w_b = 2;
    int64_t launches = 0, splits = 0, merges = 0;
    std::vector<int> last_interval_changes;  // per instance: adaptations that changed the last interval (undefined in the reference)
    int strategy = 0;  // 0 adaptGridTimeBasedSingleStep, 1 adaptGridRedundantControls
    int redundant_backup = 1;
    double redundant_epsilon = 1e-3;
    int *d_ops = nullptr, *d_nops = nullptr;  // edit scripts of the redundant-controls strategy [B][ADAPT_KMAX], [B]
    double* d_scratch = nullptr;
    std::vector<void*> allocations;
};

namespace {

int adaptiveEnsureBucket(b200sqp_adaptive* a, int idx)
{
    if (a->bucket[idx]) return B200SQP_OK;
    b200sqp_ocp o = a->ocp;
    o.n_grid      = a->n_lo + idx;
    b200sqp_handle h = nullptr;
    int rc = b200sqp_create(&o, a->B, a->device, &h);
    if (rc) return rc;
    a->bucket[idx] = h;
    AdaptBucketView& v = a->views[idx];
    v.z[0] = h->st.z[0], v.z[1] = h->st.z[1], v.cur = h->st.cur, v.x0 = h->st.x0, v.xref = h->st.xref, v.chi2 = h->st.chi2, v.status = h->st.status;
    v.K = h->s.K;
    CUDA_TRY(cudaMemcpyAsync(a->d_views + idx, &v, sizeof(v), cudaMemcpyHostToDevice, a->stream));
    CUDA_TRY(cudaStreamSynchronize(a->stream));
    return B200SQP_OK;
}

void adaptiveFillPinned(b200sqp_adaptive* a)
{
    unsigned mask = 0;
    for (int i = 0; i < a->nx; ++i)
        if (a->ocp.xf_fixed[i]) mask |= 1u << i;
    if (!mask) return;
    launchAdaptFillPinned(a->d_views, a->d_plan, a->d_xref_master, mask, a->nx, a->nu, a->B, a->stream);  // one launch for all buckets
    a->launches += 1;
}

// NonUniformFiniteDifferencesVariableGrid::adaptGrid for every instance; *changed = any grid changed
int adaptiveAdapt(b200sqp_adaptive* a, bool* changed)
{
    *changed        = false;
    const double hi = a->ocp.dt_ref * (1.0 + a->hyst), lo = a->ocp.dt_ref * (1.0 - a->hyst);
    // one launch over the whole batch: every instance finds its bucket and slot through the plan
    if (a->strategy == 1)  // d_decision receives the new grid size, d_ops / d_nops the edit script
        launchAdaptDecideRedundant(a->d_views, a->d_plan, a->nx, a->nu, a->B, a->redundant_epsilon, a->redundant_backup, a->n_min, a->n_max,
                                   a->d_decision, a->d_ops, a->d_nops, a->stream);
    else
        launchAdaptDecide(a->d_views, a->d_plan, a->nx, a->nu, a->B, hi, lo, a->n_min, a->n_max, a->d_decision, a->stream);
    a->launches += 1;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(a->hp_decision, a->d_decision, sizeof(int) * a->B, cudaMemcpyDeviceToHost, a->stream));
    CUDA_TRY(cudaStreamSynchronize(a->stream));
    bool any = false;
    if (a->strategy == 1)
        for (int i = 0; i < a->B && !any; ++i) any = a->hp_decision[i] != a->n_lo + a->bucket_of[i];
    else
        for (int i = 0; i < a->B && !any; ++i) any = a->hp_decision[i] != 0;
    if (!any) return B200SQP_OK;

    const int B = a->B;
    std::vector<int> new_count(a->count.size(), 0);
    int k_max = 1;
    for (int i = 0; i < B; ++i)
    {
        const int dec = a->hp_decision[i], type = dec & 3;
        const int src = a->bucket_of[i];
        int dst;
        if (a->strategy == 1)
        {
            dst = dec - a->n_lo;  // the decision is the new grid size
            if (dst > src) a->splits += dst - src;
            if (dst < src) a->merges += src - dst;
        }
        else
        {
            dst = src + (type == ADAPT_SPLIT ? 1 : type == ADAPT_MERGE ? -1 : 0);
            a->splits += type == ADAPT_SPLIT;
            a->merges += type == ADAPT_MERGE;
            if (type != ADAPT_NONE && (dec >> 2) == a->n_lo + src - 2) a->last_interval_changes[i] += 1;  // interval K - 1 of a grid of K + 1 points
        }
        if (dst < 0 || dst >= (int)a->bucket.size()) return fail(B200SQP_ERR_INVALID, "grid adaptation left the bucket range");
        int rc = adaptiveEnsureBucket(a, dst);
        if (rc) return rc;
        a->hp_plan[i]         = src;
        a->hp_plan[B + i]     = a->slot_of[i];
        a->hp_plan[2 * B + i] = dst;
        a->hp_plan[3 * B + i] = new_count[dst]++;
        a->hp_plan[4 * B + i] = dec;
        k_max                 = std::max(k_max, a->n_lo + dst - 1);
    }
    CUDA_TRY(cudaMemcpyAsync(a->d_plan, a->hp_plan, sizeof(int) * 5 * B, cudaMemcpyHostToDevice, a->stream));
    if (a->strategy == 1)
        launchAdaptApplyOps(a->d_views, a->d_plan, a->d_ops, a->d_nops, a->d_scratch, a->d_x0_master, a->d_xref_master, a->nx, a->nu, B, a->stream);
    else
        launchAdaptMigrate(a->d_views, a->d_plan, a->d_x0_master, a->d_xref_master, a->nx, a->nu, a->warm_start ? 0 : 1, k_max, B, a->stream);
    a->launches += 2;
    CUDA_TRY(cudaGetLastError());
    for (int i = 0; i < B; ++i)
    {
        a->bucket_of[i] = a->hp_plan[2 * B + i];
        a->slot_of[i]   = a->hp_plan[3 * B + i];
    }
    a->count = new_count;
    CUDA_TRY(cudaStreamSynchronize(a->stream));  // hp_plan is rewritten by the next call
    adaptiveFillPinned(a);  // both parameter buffers of a slot carry the fixed goal components
    CUDA_TRY(cudaGetLastError());
    *changed = true;
    return B200SQP_OK;
}

}  // namespace

int b200sqp_adaptive_create(const b200sqp_ocp* ocp, int32_t batch, int32_t device, int32_t n_min, int32_t n_max, double dt_hyst_ratio,
                            int32_t warm_start, b200sqp_adaptive_handle* out)
{
    if (!ocp || !out || batch < 1) return fail(B200SQP_ERR_INVALID, "null argument or batch < 1");
    *out = nullptr;
    if (ocp->grid != B200SQP_GRID_FD_NONUNIFORM_VARDT)
        return fail(B200SQP_ERR_UNSUPPORTED, "grid adaptation is implemented for the non-uniform time-optimal grid (one dt vertex per interval)");
    if (n_min < 3 || n_max < n_min || !(dt_hyst_ratio >= 0.0 && dt_hyst_ratio < 1.0))
        return fail(B200SQP_ERR_INVALID, "need 3 <= n_min <= n_max and 0 <= dt_hyst_ratio < 1");
    if (ocp->n_grid < 3) return fail(B200SQP_ERR_INVALID, "n_grid < 3");
    b200sqp_adaptive* a = new b200sqp_adaptive;
    a->ocp = *ocp, a->B = batch, a->device = device, a->n_min = n_min, a->n_max = n_max, a->hyst = dt_hyst_ratio, a->warm_start = warm_start ? 1 : 0;
    a->nx = ocp->nx, a->nu = ocp->nu;
    // a grid only shrinks while N > n_min and only grows while N < n_max, so its size stays inside the hull of [n_min, n_max] and n_grid
    a->n_lo = std::min(n_min, ocp->n_grid), a->n_hi = std::max(n_max, ocp->n_grid);
    const int nbuckets = a->n_hi - a->n_lo + 1;
    a->bucket.assign(nbuckets, nullptr);
    a->count.assign(nbuckets, 0);
    a->views.assign(nbuckets, AdaptBucketView{});
    a->bucket_of.assign(batch, ocp->n_grid - a->n_lo);
    a->slot_of.resize(batch);
    a->last_interval_changes.assign(batch, 0);
    for (int i = 0; i < batch; ++i) a->slot_of[i] = i;
    auto destroy_and_fail = [&](int code, const std::string& msg) {
        b200sqp_adaptive_destroy(a);
        return fail(code, msg);
    };
    int dev_count = 0;
    if (cudaGetDeviceCount(&dev_count) != cudaSuccess || dev_count == 0)
    {
        cudaGetLastError();
        return destroy_and_fail(B200SQP_ERR_NO_DEVICE, "no CUDA device visible: the LM hot path only exists as sm_100a kernels");
    }
    if (device < 0 || device >= dev_count) return destroy_and_fail(B200SQP_ERR_INVALID, "device index out of range");
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&a->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&a->ev_ready, cudaEventDisableTiming);
    auto A = [&](auto** p, size_t cnt) {
        if (e == cudaSuccess) e = cudaMalloc((void**)p, sizeof(**p) * cnt);
        if (e == cudaSuccess) a->allocations.push_back(*p);
        if (e == cudaSuccess) e = cudaMemsetAsync(*p, 0, sizeof(**p) * cnt, a->stream);
    };
    A(&a->d_views, (size_t)nbuckets);
    A(&a->d_plan, (size_t)5 * batch);
    A(&a->d_decision, (size_t)batch);
    A(&a->d_x0_master, (size_t)batch * a->nx);
    A(&a->d_xref_master, (size_t)batch * a->nx);
    A(&a->d_u0, (size_t)batch * a->nu);
    A(&a->d_chi2, (size_t)batch);
    A(&a->d_status, (size_t)batch);
    if (e == cudaSuccess) e = cudaMallocHost((void**)&a->hp_plan, sizeof(int) * 5 * batch);
    if (e == cudaSuccess) e = cudaMallocHost((void**)&a->hp_decision, sizeof(int) * batch);
    if (e != cudaSuccess) return destroy_and_fail(B200SQP_ERR_CUDA, std::string("device allocation failed: ") + cudaGetErrorString(e));
    const int idx0 = ocp->n_grid - a->n_lo;
    int rc         = adaptiveEnsureBucket(a, idx0);  // also the check that the structure is in the registry
    if (rc)
    {
        const std::string msg = g_last_error;
        return destroy_and_fail(rc, msg);
    }
    a->count[idx0] = batch;
    for (int i = 0; i < batch; ++i)
    {
        a->hp_plan[i] = a->hp_plan[2 * batch + i] = idx0;
        a->hp_plan[batch + i] = a->hp_plan[3 * batch + i] = i;
        a->hp_plan[4 * batch + i]                          = 0;
    }
    e = cudaMemcpyAsync(a->d_plan, a->hp_plan, sizeof(int) * 5 * batch, cudaMemcpyHostToDevice, a->stream);
    if (e == cudaSuccess && !rc) e = cudaStreamSynchronize(a->stream);
    if (e != cudaSuccess || rc) return destroy_and_fail(B200SQP_ERR_CUDA, "initial grouping upload failed");
    *out = a;
    return B200SQP_OK;
}

int b200sqp_adaptive_destroy(b200sqp_adaptive_handle a)
{
    if (!a) return B200SQP_OK;
    cudaSetDevice(a->device);
    if (a->stream) cudaStreamSynchronize(a->stream);
    for (b200sqp_handle h : a->bucket) b200sqp_destroy(h);
    for (void* p : a->allocations) cudaFree(p);
    if (a->hp_plan) cudaFreeHost(a->hp_plan);
    if (a->hp_decision) cudaFreeHost(a->hp_decision);
    if (a->ev_ready) cudaEventDestroy(a->ev_ready);
    if (a->stream) cudaStreamDestroy(a->stream);
    delete a;
    return B200SQP_OK;
}

int b200sqp_adaptive_step(b200sqp_adaptive_handle a, const b200sqp_lm_options* opts, int32_t num_ocp_iterations, const double* x0,
                          const double* xref, double* u0_out, double* chi2_out, int32_t* status_out, int32_t* n_out)
{
    if (!a || !opts || !x0 || !xref || num_ocp_iterations < 1 || opts->iterations < 0) return fail(B200SQP_ERR_INVALID, "bad argument");
    CUDA_TRY(cudaSetDevice(a->device));
    const int B = a->B, nx = a->nx, nu = a->nu;
    for (int it = 0; it < num_ocp_iterations; ++it)
    {
        const bool new_run = it == 0;
        // FullDiscretizationGridBase / NonUniformFullDiscretizationGridBase::update: adaptGrid unless first run or new run
        if (!a->first_run && !new_run)
        {
            bool changed = false;
            int rc       = adaptiveAdapt(a, &changed);
            if (rc) return rc;
        }
        if (new_run)
        {
            CUDA_TRY(cudaMemcpyAsync(a->d_x0_master, x0, sizeof(double) * (size_t)B * nx, cudaMemcpyHostToDevice, a->stream));
            CUDA_TRY(cudaMemcpyAsync(a->d_xref_master, xref, sizeof(double) * (size_t)B * nx, cudaMemcpyHostToDevice, a->stream));
            launchAdaptScatterStart(a->d_views, a->d_plan, a->d_x0_master, a->d_xref_master, nx, B, a->stream);
            a->launches += 1;
            adaptiveFillPinned(a);
            CUDA_TRY(cudaGetLastError());
        }
        if (a->first_run || !a->warm_start)
        {
            // initializeSequences at the current grid size of every instance (n_init = _n_adapt, non_uniform_full_discretization_grid_base.cpp:161)
            for (size_t b = 0; b < a->bucket.size(); ++b)
                if (a->count[b] > 0)
                {
                    b200sqp_handle h = a->bucket[b];
                    launchInitTrajectories(h->st.x0, h->st.xref, nullptr, h->st.z[0], h->st.cur, h->s.K, nx, nu, h->s.vt, a->ocp.dt_ref, a->count[b], h->S,
                                           a->stream);
                    a->launches += 1;
                }
            CUDA_TRY(cudaGetLastError());
        }
        CUDA_TRY(cudaEventRecord(a->ev_ready, a->stream));
        // resetWeights / adaptWeights (levenberg_marquardt_sparse.cpp:83-86, 264-287): one solver object per instance, all in lockstep
        if (new_run || !a->weights_initialised)
            a->w_eq = opts->weight_eq, a->w_ineq = opts->weight_ineq, a->w_b = opts->weight_bounds;
        else
        {
            a->w_eq   = std::min(a->w_eq * opts->adapt_factor_eq, opts->adapt_max_eq);
            a->w_ineq = std::min(a->w_ineq * opts->adapt_factor_ineq, opts->adapt_max_ineq);
            a->w_b    = std::min(a->w_b * opts->adapt_factor_bounds, opts->adapt_max_bounds);
        }
        a->weights_initialised = true;
        for (size_t b = 0; b < a->bucket.size(); ++b)
            if (a->count[b] > 0)
            {
                b200sqp_handle h = a->bucket[b];
                int rc           = ensureTrace(h, opts->iterations);
                if (rc) return rc;
                h->st.w_eq = a->w_eq, h->st.w_ineq = a->w_ineq, h->st.w_b = a->w_b;
                h->weights_initialised = true;
                DeviceOcp P            = h->P;
                P.B                    = a->count[b];  // the occupied slots of the bucket
                CUDA_TRY(cudaStreamWaitEvent(h->stream, a->ev_ready, 0));
                CUDA_TRY(cudaEventRecord(h->ev_begin, h->stream));
                h->kernels->solve(P, h->st, opts->iterations, h->threads_per_instance, h->solve_flags, h->stream);
                CUDA_TRY(cudaGetLastError());
                CUDA_TRY(cudaEventRecord(h->ev_end, h->stream));
                h->timed = true;
                h->launches += 1;
                a->launches += 1;
                CUDA_TRY(cudaStreamWaitEvent(a->stream, h->ev_end, 0));
            }
        a->first_run = false;
    }
    launchAdaptGather(a->d_views, a->d_plan, nx, nu, a->d_u0, a->d_chi2, a->d_status, B, a->stream);
    a->launches += 1;
    CUDA_TRY(cudaGetLastError());
    if (u0_out) CUDA_TRY(cudaMemcpyAsync(u0_out, a->d_u0, sizeof(double) * (size_t)B * nu, cudaMemcpyDeviceToHost, a->stream));
    if (chi2_out) CUDA_TRY(cudaMemcpyAsync(chi2_out, a->d_chi2, sizeof(double) * B, cudaMemcpyDeviceToHost, a->stream));
    if (status_out) CUDA_TRY(cudaMemcpyAsync(status_out, a->d_status, sizeof(int32_t) * B, cudaMemcpyDeviceToHost, a->stream));
    CUDA_TRY(cudaStreamSynchronize(a->stream));
    if (n_out)
        for (int i = 0; i < B; ++i) n_out[i] = a->n_lo + a->bucket_of[i];
    return B200SQP_OK;
}

int b200sqp_adaptive_set_redundant_controls(b200sqp_adaptive_handle a, int32_t num_backup_nodes, double epsilon)
{
    if (!a || num_backup_nodes < 0 || !(epsilon >= 0.0)) return fail(B200SQP_ERR_INVALID, "bad argument");
    if (!a->first_run) return fail(B200SQP_ERR_INVALID, "the adaptation strategy must be chosen before the first step");
    if (a->n_hi > ADAPT_KMAX + 1) return fail(B200SQP_ERR_UNSUPPORTED, "the redundant-controls strategy is sized for grids of at most 129 points");
    CUDA_TRY(cudaSetDevice(a->device));
    if (!a->d_ops)
    {
        cudaError_t e = cudaSuccess;
        auto A        = [&](auto** p, size_t cnt) {
            if (e == cudaSuccess) e = cudaMalloc((void**)p, sizeof(**p) * cnt);
            if (e == cudaSuccess) a->allocations.push_back(*p);
            if (e == cudaSuccess) e = cudaMemsetAsync(*p, 0, sizeof(**p) * cnt, a->stream);
        };
        A(&a->d_ops, (size_t)a->B * ADAPT_KMAX);
        A(&a->d_nops, (size_t)a->B);
        A(&a->d_scratch, (size_t)a->B * ((size_t)(ADAPT_KMAX + 1) * a->nx + (size_t)ADAPT_KMAX * (a->nu + 1)));
        if (e != cudaSuccess) return fail(B200SQP_ERR_CUDA, std::string("device allocation failed: ") + cudaGetErrorString(e));
    }
    a->strategy = 1, a->redundant_backup = num_backup_nodes, a->redundant_epsilon = epsilon;
    return B200SQP_OK;
}

int b200sqp_adaptive_last_interval_changes(b200sqp_adaptive_handle a, int32_t* count)
{
    if (!a || !count) return fail(B200SQP_ERR_INVALID, "null argument");
    for (int i = 0; i < a->B; ++i) count[i] = a->last_interval_changes[i];
    return B200SQP_OK;
}

int b200sqp_adaptive_reserve(b200sqp_adaptive_handle a, int32_t n_from, int32_t n_to)
{
    if (!a || n_from > n_to) return fail(B200SQP_ERR_INVALID, "bad argument");
    CUDA_TRY(cudaSetDevice(a->device));
    for (int n = std::max(n_from, a->n_lo); n <= std::min(n_to, a->n_hi); ++n)
    {
        int rc = adaptiveEnsureBucket(a, n - a->n_lo);
        if (rc) return rc;
    }
    return B200SQP_OK;
}

int b200sqp_adaptive_get_trajectories(b200sqp_adaptive_handle a, int32_t n_cap, double* x, double* u, double* dt, int32_t* n)
{
    if (!a || n_cap < 2) return fail(B200SQP_ERR_INVALID, "bad argument");
    CUDA_TRY(cudaSetDevice(a->device));
    for (int i = 0; i < a->B; ++i)
        if (a->n_lo + a->bucket_of[i] > n_cap) return fail(B200SQP_ERR_INVALID, "n_cap is smaller than the largest grid of the batch");
    const size_t B = a->B;
    if (a->export_cap < n_cap)
    {
        cudaError_t e = cudaSuccess;
        auto A        = [&](auto** p, size_t cnt) {
            if (e == cudaSuccess) e = cudaMalloc((void**)p, sizeof(**p) * cnt);
            if (e == cudaSuccess) a->allocations.push_back(*p);
        };
        A(&a->d_tx, B * n_cap * a->nx);
        A(&a->d_tu, B * n_cap * a->nu);
        A(&a->d_tdt, B * n_cap);
        if (!a->d_tn) A(&a->d_tn, B);
        if (e != cudaSuccess) return fail(B200SQP_ERR_CUDA, std::string("device allocation failed: ") + cudaGetErrorString(e));
        a->export_cap = n_cap;
    }
    CUDA_TRY(cudaMemsetAsync(a->d_tx, 0, sizeof(double) * B * n_cap * a->nx, a->stream));
    CUDA_TRY(cudaMemsetAsync(a->d_tu, 0, sizeof(double) * B * n_cap * a->nu, a->stream));
    CUDA_TRY(cudaMemsetAsync(a->d_tdt, 0, sizeof(double) * B * n_cap, a->stream));
    launchAdaptExport(a->d_views, a->d_plan, a->d_x0_master, a->nx, a->nu, n_cap, a->d_tx, a->d_tu, a->d_tdt, a->d_tn, a->B, a->stream);
    a->launches += 1;
    CUDA_TRY(cudaGetLastError());
    if (x) CUDA_TRY(cudaMemcpyAsync(x, a->d_tx, sizeof(double) * B * n_cap * a->nx, cudaMemcpyDeviceToHost, a->stream));
    if (u) CUDA_TRY(cudaMemcpyAsync(u, a->d_tu, sizeof(double) * B * n_cap * a->nu, cudaMemcpyDeviceToHost, a->stream));
    if (dt) CUDA_TRY(cudaMemcpyAsync(dt, a->d_tdt, sizeof(double) * B * n_cap, cudaMemcpyDeviceToHost, a->stream));
    if (n) CUDA_TRY(cudaMemcpyAsync(n, a->d_tn, sizeof(int32_t) * B, cudaMemcpyDeviceToHost, a->stream));
    CUDA_TRY(cudaStreamSynchronize(a->stream));
    return B200SQP_OK;
}

int b200sqp_adaptive_statistics(b200sqp_adaptive_handle a, int32_t* occupied_buckets, int64_t* splits, int64_t* merges, int64_t* launches)
{
    if (!a) return fail(B200SQP_ERR_INVALID, "null handle");
    if (occupied_buckets)
    {
        int c = 0;
        for (int v : a->count) c += v > 0;
        *occupied_buckets = c;
    }
    if (splits) *splits = a->splits;
    if (merges) *merges = a->merges;
    if (launches) *launches = a->launches;
    return B200SQP_OK;
}

}  // extern "C"
