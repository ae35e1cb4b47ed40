// Host-visible launch table: the C ABI (api.cpp) never sees a kernel template, only these function pointers.
#pragma once

#include <cuda_runtime.h>

namespace b200sqp {

struct DeviceOcp;
struct DeviceState;
template <class Real>
struct PipeArraysT;

enum { SOLVE_FORCE_GENERAL_FEATURES = 1 };  // flags of KernelSet::solve

struct KernelSet
{
    int dynamics, defect, vt, nx, nu;
    int max_threads;  // widest cooperating-thread variant compiled for this combination
    void (*solve)(const DeviceOcp&, const DeviceState&, int iterations, int threads_per_instance /*0 = auto*/, int flags /*SOLVE_**/, cudaStream_t);
    void (*evaluate)(const DeviceOcp&, const DeviceState&, double* values, double* jac, const int* value_rows, const int* jac_pos, int v_count,
                     int j_count, cudaStream_t);
    // warp-cooperative pipeline for large stage blocks (lm_pipeline.cuh) or nullptr; blocks the host, false = pass bound hit
    bool (*pipeline)(const DeviceOcp&, const DeviceState&, const PipeArraysT<double>&, int iterations, cudaStream_t);
    // the same in reduced precision (normal equations and factor in fp32; b200sqp_set_precision) or nullptr
    bool (*pipeline_f32)(const DeviceOcp&, const DeviceState&, const PipeArraysT<float>&, int iterations, cudaStream_t);
};

// closed registry (kernels_*.cu); nullptr = combination not compiled in -> B200SQP_ERR_UNSUPPORTED, never a CPU fallback
const KernelSet* findKernels(int dynamics, int defect, int vt);

// one table per translation unit so that the heavy templates compile in parallel
const KernelSet* kernelTableVdpCn(int* count);
const KernelSet* kernelTableVdpFd(int* count);
const KernelSet* kernelTableVdpMs(int* count);
const KernelSet* kernelTableOscillators(int* count);
const KernelSet* kernelTableCartPole(int* count);
const KernelSet* kernelTableUnicycle(int* count);
const KernelSet* kernelTableQuadrotor(int* count);
const KernelSet* kernelTableBenchmarkSystems(int* count);
const KernelSet* kernelTableCombosFd(int* count);
const KernelSet* kernelTableCombosMs(int* count);
const KernelSet* kernelTableLinear(int* count);
const KernelSet* kernelTableLinear4(int* count);
const KernelSet* kernelTableIntegrators(int* count);
const KernelSet* kernelTableDtEquality(int* count);
// structures with full (non-diagonal) cost weights: their own (smaller) registry; nullptr = combination not compiled in
const KernelSet* kernelTableDenseCost(int* count);
const KernelSet* findDenseCostKernels(int dynamics, int defect, int vt);

// layout helpers (util_kernels.cu); all arrays device pointers
// params [B][n] (reference order)  <->  z (block order, tiled instance-minor [tile][K*NB][32]); pinned slots (ref index -1) are left alone
void launchPack(const double* params, int n, const int* ref_of_internal, int slots, const double* pinned_values /*[slots] or null*/, double* z,
                const int* cur_or_null, double* z_alt, int B, int S, cudaStream_t);
void launchUnpack(const double* z0, const double* z1, const int* cur, const int* internal_of_ref, int n, int slots, double* params, int B, int S,
                  cudaStream_t);
// fixed goal components <- xref
void launchFillPinned(const double* xref, double* z0, double* z1, int slot0, int slots, int nx, unsigned mask, int B, cudaStream_t);
// x0/xref: [B][nx] -> tiled [tile][nx][32]
void launchTransposeIn(const double* src, int dim, double* dst, int B, int S, cudaStream_t);
// x0 (+ xref or null = zero) [B][nx], device or pinned-host-mapped -> tiled arrays + host-order device copies, one launch
void launchIngest(const double* x0_src, const double* xref_src, int nx, double* x0_tiled, double* xref_tiled, double* x0_copy, double* xref_copy,
                  int B, cudaStream_t);
// first controls -> u0_dev (+ u0_out), chi2 -> chi2_out, status -> status_out; the *_out are device views of pinned host buffers or null
void launchExport(const double* z0, const double* z1, const int* cur, int nu, int slots, const double* chi2, const int* status, double* u0_dev,
                  double* u0_out, double* chi2_out, int* status_out, int B, cudaStream_t);
// [rows][S] -> [B][rows]
void launchTransposeOut(const double* src, int rows, double* dst, int B, int S, cudaStream_t);
// FullDiscretizationGridBase::initializeSequences (full_discretization_grid_base.cpp:134-179) on the device
void launchInitTrajectories(const double* x0 /*[nx][S]*/, const double* xref /*[nx][S]*/, const double* xtraj /*[(K+1)*nx][S] or null*/, double* z,
                            int* cur, int K, int nx, int nu, int vt, double dt_ref, int B, int S, cudaStream_t);
// u_0 of every instance -> [B][nu]
void launchFirstControls(const double* z0, const double* z1, const int* cur, int nu, int slots, double* u0, int B, int S, cudaStream_t);

// SystemDynamicsInterface::getLinearA/getLinearB by forward (method 0) or central (method 1) differences for B points (kernels_linearize.cu);
// x [B][nx], u [B][nu], A [B][nx*nx] / Bm [B][nx*nu] column-major per point (either may be null); false = dynamics id not in the registry
struct DynParams;
bool launchLinearizeDynamics(int dynamics, const DynParams& dyn, int method, int B, const double* x, const double* u, double* A, double* Bm,
                             cudaStream_t);
// ForwardDifferences::hessian (method 0) / CentralDifferences::hessian (method 1) of the dynamics w.r.t. [x; u] for B points
// (kernels_linearize.cu); mult [B][nx] or null, H [B][(nx+nu)^2] column-major per point; false = dynamics id not in the registry
bool launchDynamicsHessian(int dynamics, const DynParams& dyn, int method, int B, const double* x, const double* u, const double* mult, double* H,
                           cudaStream_t);
// SimulatedPlant::control for B plants: x_next = solveIVP(x, u, dt), integrator 0 = explicit Euler, 1 = RK4 (kernels_plant.cu);
// x, x_next [B][nx], u [B][nu]; u_log [B][nu] or null receives a copy of u, chi2_log / status_log [B] or null a copy of chi2_src /
// status_src (the closed-loop log); false = dynamics id not in the registry
bool launchPlantStep(int dynamics, const DynParams& dyn, int integrator, double dt, int B, const double* x, const double* u, double* x_next,
                     double* u_log, const double* chi2_src, double* chi2_log, const int* status_src, int* status_log, cudaStream_t);
// FullDiscretizationGridBase::warmStartShifting + findNearestState per instance, then x_seq.front() = x0_new (util_kernels.cu): a
// per-instance search kernel and a per-(instance, block) move into the instance's other parameter buffer (roles swap); plan [B] scratch
void launchWarmStartShift(const double* x0_new /*[B][nx]*/, double* x0 /*tiled*/, double* z0, double* z1, int* cur, int K, int nx, int nu,
                          int* plan /*[B]*/, int* num_shift /*[B] or null*/, int B, cudaStream_t);
// epilogue of a solve that did not run the fused LM kernel (the pipeline): stores chi2 of this rank's instances into every rank's gather
// buffer and bumps the arrival counters, with the same protocol and the same number of arriving thread blocks as lmSolveKernel's epilogue
void launchPeerPublish(const DeviceState& st, int B, cudaStream_t);
// bounded spin on the arrival counters of the fused peer-memory gather (b200sqp_peer_wait)
void launchPeerWait(const unsigned long long* arrivals, int world, unsigned long long expected, unsigned long long timeout_ns, int* timed_out,
                    cudaStream_t);

// grid adaptation of the time-optimal grids (adapt_kernels.cu; b200sqp_adaptive_*): decision per instance of one bucket, migration of every
// instance to (bucket, slot) of its new grid size, start-state scatter, result gather and trajectory export in batch order.
// plan [5][B] device: source bucket, source slot, destination bucket, destination slot, decision (ADAPT_* | interval << 2)
struct AdaptBucketView;
void launchAdaptDecide(const AdaptBucketView* views, const int* plan, int nx, int nu, int B, double hi, double lo, int n_min, int n_max, int* decision,
                       cudaStream_t);
// fixed goal components <- reference for every instance's slot (mask: bit j = component j of the goal is fixed)
void launchAdaptFillPinned(const AdaptBucketView* views, const int* plan, const double* xref_master, unsigned mask, int nx, int nu, int B, cudaStream_t);
void launchAdaptMigrate(const AdaptBucketView* views, const int* plan, double* x0_master, const double* xref_master, int nx, int nu, int keep_start,
                        int k_max, int B, cudaStream_t);
// strategy adaptGridRedundantControls: per-instance edit script from (u, dt) + new grid size; replay on the trajectory while moving
// (ops [B][ADAPT_KMAX], nops [B], scratch [B][(ADAPT_KMAX + 1) nx + ADAPT_KMAX (nu + 1)])
void launchAdaptDecideRedundant(const AdaptBucketView* views, const int* plan, int nx, int nu, int B, double eps, int backup, int n_min, int n_max,
                                int* new_n, int* ops, int* nops, cudaStream_t);
void launchAdaptApplyOps(const AdaptBucketView* views, const int* plan, const int* ops, const int* nops, double* scratch, const double* x0_master,
                         const double* xref_master, int nx, int nu, int B, cudaStream_t);
void launchAdaptScatterStart(const AdaptBucketView* views, const int* plan, const double* x0_master, const double* xref_master, int nx, int B,
                             cudaStream_t);
void launchAdaptGather(const AdaptBucketView* views, const int* plan, int nx, int nu, double* u0, double* chi2, int* status, int B, cudaStream_t);
void launchAdaptExport(const AdaptBucketView* views, const int* plan, const double* x0_master, int nx, int nu, int n_cap, double* x, double* u,
                       double* dt, int* n, int B, cudaStream_t);

// measured fp64 FMA throughput of the device (TFLOP/s, 2 flops per FMA), all SMs, best of three timed launches; < 0 on failure
double measureFp64PeakTflops(int sm_count, cudaStream_t);

}  // namespace b200sqp
