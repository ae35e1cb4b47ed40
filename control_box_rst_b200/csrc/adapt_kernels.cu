// Grid adaptation of the time-optimal grids on the device: per-instance grid size N.
//
// Instances of one batch are bucketed by their grid size; bucket N is an ordinary solver handle of the structure with n_grid = N whose
// first `count` slots are in use (api.cpp, b200sqp_adaptive_*).  The kernels here implement the grid side of one OCP iteration:
//   decide   NonUniformFiniteDifferencesVariableGrid::adaptGridTimeBasedSingleStep (non_uniform_finite_differences_variable_grid.cpp:206-257):
//            walk dt_0 .. dt_{N-2}; the first dt above dt_ref (1 + hyst) with N < n_max splits its interval, the first one below
//            dt_ref (1 - hyst) with N > n_min is merged into its successor; one change per call
//   migrate  apply the change while moving every instance to its slot in the bucket of its new size: insertion of the mid state
//            0.5 (x_i + x_{i+1}) with the control of interval i and HALF its dt (the interval itself keeps dt_i, as the reference does),
//            or removal of (x_i, u_i, dt_i) with dt_{i+1} += dt_i; removing node 0 makes the old x_1 the (fixed) start state until the next
//            measurement overwrites it (:239-240 and non_uniform_full_discretization_grid_base.cpp:111)
//   commit   swap the roles of the two parameter buffers of every written slot
// Two cases the reference leaves undefined (it indexes one past the end of its vertex vectors) are defined here: a split of the LAST
// interval uses the final state x_f as its right neighbour, a merge of the last interval drops its dt.
//
// Trajectory layout of a bucket (lm_device.cuh): K = N-1 blocks of nb = nu + 1 + nx slots (u_k, dt_k, x_{k+1}) per instance, tiled
// instance-minor; x_0 lives in the bucket's x0 array, the last block holds x_f (fixed components pinned to the reference).
#include "../../include/b200sqp.h"
#include "launch.h"
#include "lm_device_types.h"

namespace b200sqp {

namespace {

__device__ __forceinline__ size_t tiled(int i, int slot, int nslots) { return ((size_t)(i >> 5) * nslots + slot) * 32 + (i & 31); }

// one thread per instance of the batch; the instance's current (bucket, slot) are rows 2 and 3 of the plan
__global__ void adaptDecideKernel(const AdaptBucketView* __restrict__ views, const int* __restrict__ plan, int nx, int nu, int B, double hi,
                                  double lo, int n_min, int n_max, int* __restrict__ decision)
{
    const int inst = blockIdx.x * blockDim.x + threadIdx.x;
    if (inst >= B) return;
    const AdaptBucketView b = views[plan[2 * B + inst]];
    const int j = plan[3 * B + inst], K = b.K, nb = nu + 1 + nx;
    const double* src = b.cur[j] ? b.z[1] : b.z[0];
    const int slots = K * nb, n = K + 1;
    int dec = ADAPT_NONE;
    for (int k = 0; k < K; ++k)
    {
        const double dt = src[tiled(j, k * nb + nu, slots)];
        if (dt > hi && n < n_max)
        {
            dec = ADAPT_SPLIT | (k << 2);
            break;
        }
        else if (dt < lo && n > n_min)
        {
            dec = ADAPT_MERGE | (k << 2);
            break;
        }
    }
    decision[inst] = dec;
}

struct Source
{
    const double* z;
    const double* x0;  // master copy of the start state of this instance [nx]
    int slot, K, nb, nu;
    __device__ double x(int node, int j) const { return node == 0 ? x0[j] : z[tiled(slot, (node - 1) * nb + nu + 1 + j, K * nb)]; }
    __device__ double u(int k, int j) const { return z[tiled(slot, k * nb + j, K * nb)]; }
    __device__ double dt(int k) const { return z[tiled(slot, k * nb + nu, K * nb)]; }
};

// one thread per (instance, destination block)
__global__ void adaptMigrateKernel(const AdaptBucketView* __restrict__ views, const int* __restrict__ plan /*[5][B]: src bucket, src slot, dst
                                   bucket, dst slot, decision*/, double* __restrict__ x0_master, const double* __restrict__ xref_master, int nx,
                                   int nu, int keep_start, int B)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int kd = blockIdx.y;
    if (i >= B) return;
    const AdaptBucketView sb = views[plan[i]], db = views[plan[2 * B + i]];
    const int sslot = plan[B + i], dslot = plan[3 * B + i], dec = plan[4 * B + i];
    const int type = dec & 3, p = dec >> 2;
    const int Kd = db.K, nb = nu + 1 + nx;
    if (kd >= Kd) return;
    Source s{sb.cur[sslot] ? sb.z[1] : sb.z[0], x0_master + (size_t)i * nx, sslot, sb.K, nb, nu};
    double* dst = db.z[db.cur[dslot] ? 0 : 1];  // the slot's idle buffer: its current one may still be read as somebody's source
    const int dslots = Kd * nb;

    // interval kd of the new grid
    int ks = kd;  // source interval
    double dtv;
    if (type == ADAPT_SPLIT)
    {
        ks  = kd <= p ? kd : kd - 1;
        dtv = (kd == p + 1) ? 0.5 * s.dt(p) : s.dt(ks);
    }
    else if (type == ADAPT_MERGE)
    {
        ks  = kd < p ? kd : kd + 1;
        dtv = s.dt(ks);
        if (kd == p) dtv += s.dt(p);  // _dt_seq[i + 1].value() += dt
    }
    else
        dtv = s.dt(ks);
    for (int j = 0; j < nu; ++j) dst[tiled(dslot, kd * nb + j, dslots)] = s.u(ks, j);
    dst[tiled(dslot, kd * nb + nu, dslots)] = dtv;

    // node kd + 1 of the new grid (the last one is x_f)
    const int m = kd + 1;
    for (int j = 0; j < nx; ++j)
    {
        double v;
        if (type == ADAPT_SPLIT)
            v = (m == p + 1) ? 0.5 * (s.x(p, j) + s.x(p + 1, j)) : s.x(m <= p ? m : m - 1, j);
        else if (type == ADAPT_MERGE)
            v = s.x(m < p ? m : m + 1, j);
        else
            v = s.x(m, j);
        dst[tiled(dslot, kd * nb + nu + 1 + j, dslots)] = v;
    }

    if (kd == 0)
    {
        // start state and reference of the slot; a removed node 0 promotes the old x_1 (read before the master copy is overwritten: no
        // other block of this instance reads node 0 in that case)
        for (int j = 0; j < nx; ++j)
        {
            double v = x0_master[(size_t)i * nx + j];
            if (type == ADAPT_MERGE && p == 0 && !keep_start)
            {
                v                              = s.x(1, j);
                x0_master[(size_t)i * nx + j] = v;
            }
            db.x0[tiled(dslot, j, nx)]   = v;
            db.xref[tiled(dslot, j, nx)] = xref_master[(size_t)i * nx + j];
        }
    }
}

__global__ void adaptCommitKernel(const AdaptBucketView* __restrict__ views, const int* __restrict__ plan, int B)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    const AdaptBucketView db = views[plan[2 * B + i]];
    const int dslot          = plan[3 * B + i];
    db.cur[dslot]            = db.cur[dslot] ? 0 : 1;  // every (bucket, slot) is the destination of exactly one instance
}

// measured start states / references of a new run -> the buckets' tiled arrays (x_seq.front() = x0; the fixed goal components follow
// through launchFillPinned)
__global__ void adaptScatterStartKernel(const AdaptBucketView* __restrict__ views, const int* __restrict__ plan /*dst bucket, dst slot at rows 2, 3*/,
                                        const double* __restrict__ x0_master, const double* __restrict__ xref_master, int nx, int B)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    const AdaptBucketView db = views[plan[2 * B + i]];
    const int dslot          = plan[3 * B + i];
    for (int j = 0; j < nx; ++j)
    {
        db.x0[tiled(dslot, j, nx)]   = x0_master[(size_t)i * nx + j];
        db.xref[tiled(dslot, j, nx)] = xref_master[(size_t)i * nx + j];
    }
}

// fixed goal components <- reference, both parameter buffers of every instance's slot (launchFillPinned for the whole batch at once)
__global__ void adaptFillPinnedKernel(const AdaptBucketView* __restrict__ views, const int* __restrict__ plan, const double* __restrict__ xref_master,
                                      unsigned mask, int nx, int nu, int B)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    const AdaptBucketView b = views[plan[2 * B + i]];
    const int slot = plan[3 * B + i], nb = nu + 1 + nx, slots = b.K * nb, slot0 = (b.K - 1) * nb + nu + 1;
    for (int j = 0; j < nx; ++j)
        if (mask & (1u << j))
        {
            const double v                      = xref_master[(size_t)i * nx + j];
            b.z[0][tiled(slot, slot0 + j, slots)] = v;
            b.z[1][tiled(slot, slot0 + j, slots)] = v;
        }
}

// per-instance results in batch order: first control, chi2, status
__global__ void adaptGatherKernel(const AdaptBucketView* __restrict__ views, const int* __restrict__ plan, int nx, int nu, double* __restrict__ u0,
                                  double* __restrict__ chi2, int* __restrict__ status, int B)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    const AdaptBucketView b = views[plan[2 * B + i]];
    const int slot          = plan[3 * B + i];
    const double* src       = b.cur[slot] ? b.z[1] : b.z[0];
    const int nb            = nu + 1 + nx;
    for (int j = 0; j < nu; ++j) u0[(size_t)i * nu + j] = src[tiled(slot, j, b.K * nb)];
    chi2[i]   = b.chi2[slot];
    status[i] = b.status[slot];
}

// trajectories in batch order, padded to n_cap grid points: x [B][n_cap][nx], u [B][n_cap][nu], dt [B][n_cap], n [B]
__global__ void adaptExportKernel(const AdaptBucketView* __restrict__ views, const int* __restrict__ plan, const double* __restrict__ x0_master,
                                  int nx, int nu, int n_cap, double* __restrict__ x, double* __restrict__ u, double* __restrict__ dt,
                                  int* __restrict__ n, int B)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y;  // grid point
    if (i >= B) return;
    const AdaptBucketView b = views[plan[2 * B + i]];
    const int slot          = plan[3 * B + i];
    if (k == 0) n[i] = b.K + 1;
    if (k > b.K || k >= n_cap) return;
    Source s{b.cur[slot] ? b.z[1] : b.z[0], x0_master + (size_t)i * nx, slot, b.K, nu + 1 + nx, nu};
    // node 0 comes from the bucket's own start-state array: the master copy is the last measurement, which differs after a removed node 0
    for (int j = 0; j < nx; ++j) x[((size_t)i * n_cap + k) * nx + j] = k == 0 ? b.x0[tiled(slot, j, nx)] : s.x(k, j);
    if (k < b.K)
    {
        for (int j = 0; j < nu; ++j) u[((size_t)i * n_cap + k) * nu + j] = s.u(k, j);
        dt[(size_t)i * n_cap + k] = s.dt(k);
    }
}

// ---- second strategy: NonUniformFiniteDifferencesVariableGrid::adaptGridRedundantControls (non_uniform_finite_differences_variable_grid.cpp:
//      259-352).  An interval whose control repeats in its successor (every component within epsilon) or whose dt is below 1e-6 is
//      redundant (the last interval never counts).  More redundant intervals than `backup`: the surplus is removed from the back, a removed
//      interval's dt going to its predecessor, down to n_min grid points; fewer: the missing ones are created by halving the interval with
//      the largest dt (first maximum, last interval excluded), up to n_max.  Several grid points may change per call and an inserted mid
//      state can be the neighbour of the next insertion, so the edits are a per-instance SCRIPT: the decision kernel derives it from
//      (u, dt) alone -- states never influence it -- and reports the new grid size; the apply kernel replays it on the trajectory.
//      One thread per instance for both: the scripts are sequential and a few dozen steps long.
__global__ void adaptDecideRedundantKernel(const AdaptBucketView* __restrict__ views, const int* __restrict__ plan, int nx, int nu, int B, double eps,
                                           int backup, int n_min, int n_max, int* __restrict__ new_n, int* __restrict__ ops, int* __restrict__ nops)
{
    const int inst = blockIdx.x * blockDim.x + threadIdx.x;
    if (inst >= B) return;
    const AdaptBucketView b = views[plan[2 * B + inst]];
    const int j = plan[3 * B + inst], K = b.K, nb = nu + 1 + nx;
    const double* src = b.cur[j] ? b.z[1] : b.z[0];
    const int slots   = K * nb;
    int* my_ops       = ops + (size_t)inst * ADAPT_KMAX;
    int no = 0, n = K + 1;
    if (n >= 3)
    {
        unsigned char list[ADAPT_KMAX];
        int cnt = 0;
        for (int idx = 0; idx + 1 < K; ++idx)  // never the last control
        {
            bool red = src[tiled(j, idx * nb + nu, slots)] < 1e-6;
            if (!red)
            {
                red = true;
                for (int c = 0; c < nu; ++c)
                    if (!(fabs(src[tiled(j, (idx + 1) * nb + c, slots)] - src[tiled(j, idx * nb + c, slots)]) <= eps)) red = false;
            }
            if (red) list[cnt++] = (unsigned char)idx;
        }
        const int diff = cnt - backup;
        if (diff < 0)
        {
            double dtl[ADAPT_KMAX];
            int len = K;
            for (int k = 0; k < K; ++k) dtl[k] = src[tiled(j, k * nb + nu, slots)];
            for (int t = 0; t < -diff && n < n_max && len < ADAPT_KMAX; ++t)
            {
                int i = 0;
                if (n > 2)
                    for (int k = 1; k + 1 < len; ++k)  // std::max_element over [begin, end - 1): first maximum
                        if (dtl[i] < dtl[k]) i = k;
                const double new_dt = 0.5 * dtl[i];
                dtl[i]              = new_dt;
                for (int k = len; k > i + 1; --k) dtl[k] = dtl[k - 1];
                dtl[i + 1] = new_dt;
                ++len;
                ++n;
                my_ops[no++] = (i << 1) | 1;
            }
        }
        else if (diff > 0)
        {
            int it = cnt - 1;
            for (int t = 0; t < diff && n > n_min; ++t, --it)
            {
                int k = list[it];
                if (k >= n - 2) --k;
                my_ops[no++] = k << 1;
                --n;
            }
        }
    }
    new_n[inst] = n;
    nops[inst]  = no;
}

// replay of the script on the trajectory: plain per-instance scratch (nodes [K+1][nx], controls [K][nu], dt [K]), then the store into the
// idle parameter buffer of the destination slot
__global__ void adaptApplyOpsKernel(const AdaptBucketView* __restrict__ views, const int* __restrict__ plan, const int* __restrict__ ops,
                                    const int* __restrict__ nops, double* __restrict__ scratch, const double* __restrict__ x0_master,
                                    const double* __restrict__ xref_master, int nx, int nu, int B)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    const AdaptBucketView sb = views[plan[i]], db = views[plan[2 * B + i]];
    const int sslot = plan[B + i], dslot = plan[3 * B + i];
    const int nb    = nu + 1 + nx;
    Source s{sb.cur[sslot] ? sb.z[1] : sb.z[0], x0_master + (size_t)i * nx, sslot, sb.K, nb, nu};
    double* X  = scratch + (size_t)i * ((size_t)(ADAPT_KMAX + 1) * nx + (size_t)ADAPT_KMAX * (nu + 1));
    double* U  = X + (size_t)(ADAPT_KMAX + 1) * nx;
    double* DT = U + (size_t)ADAPT_KMAX * nu;
    int len    = sb.K;  // intervals; nodes = len + 1
    for (int m = 0; m <= len; ++m)
        for (int j = 0; j < nx; ++j) X[m * nx + j] = s.x(m, j);
    for (int k = 0; k < len; ++k)
    {
        for (int j = 0; j < nu; ++j) U[k * nu + j] = s.u(k, j);
        DT[k] = s.dt(k);
    }
    const int* my_ops = ops + (size_t)i * ADAPT_KMAX;
    for (int t = 0; t < nops[i]; ++t)
    {
        const int op = my_ops[t], k = op >> 1;
        if (op & 1)
        {
            // insert behind interval k: mid state, the interval's control, both halves get half the dt
            const double new_dt = 0.5 * DT[k];
            DT[k]               = new_dt;
            for (int m = len; m > k; --m)
                for (int j = 0; j < nx; ++j) X[(m + 1) * nx + j] = X[m * nx + j];
            for (int j = 0; j < nx; ++j) X[(k + 1) * nx + j] = 0.5 * (X[k * nx + j] + X[(k + 2) * nx + j]);
            for (int q = len; q > k + 1; --q)
            {
                for (int j = 0; j < nu; ++j) U[q * nu + j] = U[(q - 1) * nu + j];
                DT[q] = DT[q - 1];
            }
            for (int j = 0; j < nu; ++j) U[(k + 1) * nu + j] = U[k * nu + j];
            DT[k + 1] = new_dt;
            ++len;
        }
        else
        {
            // remove grid point k + 1: its interval's dt goes to interval k
            DT[k] += DT[k + 1];
            for (int m = k + 1; m < len; ++m)
                for (int j = 0; j < nx; ++j) X[m * nx + j] = X[(m + 1) * nx + j];
            for (int q = k + 1; q + 1 < len; ++q)
            {
                for (int j = 0; j < nu; ++j) U[q * nu + j] = U[(q + 1) * nu + j];
                DT[q] = DT[q + 1];
            }
            --len;
        }
    }
    double* dst      = db.z[db.cur[dslot] ? 0 : 1];
    const int dslots = db.K * nb;  // db.K == len by construction of the plan
    for (int k = 0; k < len && k < db.K; ++k)
    {
        for (int j = 0; j < nu; ++j) dst[tiled(dslot, k * nb + j, dslots)] = U[k * nu + j];
        dst[tiled(dslot, k * nb + nu, dslots)] = DT[k];
        for (int j = 0; j < nx; ++j) dst[tiled(dslot, k * nb + nu + 1 + j, dslots)] = X[(k + 1) * nx + j];
    }
    for (int j = 0; j < nx; ++j)
    {
        db.x0[tiled(dslot, j, nx)]   = x0_master[(size_t)i * nx + j];
        db.xref[tiled(dslot, j, nx)] = xref_master[(size_t)i * nx + j];
    }
}

}  // namespace

void launchAdaptDecide(const AdaptBucketView* views, const int* plan, int nx, int nu, int B, double hi, double lo, int n_min, int n_max, int* decision,
                       cudaStream_t stream)
{
    adaptDecideKernel<<<(B + 127) / 128, 128, 0, stream>>>(views, plan, nx, nu, B, hi, lo, n_min, n_max, decision);
}

void launchAdaptFillPinned(const AdaptBucketView* views, const int* plan, const double* xref_master, unsigned mask, int nx, int nu, int B,
                           cudaStream_t stream)
{
    if (mask) adaptFillPinnedKernel<<<(B + 127) / 128, 128, 0, stream>>>(views, plan, xref_master, mask, nx, nu, B);
}

void launchAdaptMigrate(const AdaptBucketView* views, const int* plan, double* x0_master, const double* xref_master, int nx, int nu, int keep_start,
                        int k_max, int B, cudaStream_t stream)
{
    adaptMigrateKernel<<<dim3((B + 127) / 128, k_max), 128, 0, stream>>>(views, plan, x0_master, xref_master, nx, nu, keep_start, B);
    adaptCommitKernel<<<(B + 127) / 128, 128, 0, stream>>>(views, plan, B);
}

void launchAdaptScatterStart(const AdaptBucketView* views, const int* plan, const double* x0_master, const double* xref_master, int nx, int B,
                             cudaStream_t stream)
{
    adaptScatterStartKernel<<<(B + 127) / 128, 128, 0, stream>>>(views, plan, x0_master, xref_master, nx, B);
}

void launchAdaptGather(const AdaptBucketView* views, const int* plan, int nx, int nu, double* u0, double* chi2, int* status, int B, cudaStream_t stream)
{
    adaptGatherKernel<<<(B + 127) / 128, 128, 0, stream>>>(views, plan, nx, nu, u0, chi2, status, B);
}

void launchAdaptExport(const AdaptBucketView* views, const int* plan, const double* x0_master, int nx, int nu, int n_cap, double* x, double* u,
                       double* dt, int* n, int B, cudaStream_t stream)
{
    adaptExportKernel<<<dim3((B + 127) / 128, n_cap), 128, 0, stream>>>(views, plan, x0_master, nx, nu, n_cap, x, u, dt, n, B);
}

}  // namespace b200sqp

namespace b200sqp {

void launchAdaptDecideRedundant(const AdaptBucketView* views, const int* plan, int nx, int nu, int B, double eps, int backup, int n_min, int n_max,
                                int* new_n, int* ops, int* nops, cudaStream_t stream)
{
    adaptDecideRedundantKernel<<<(B + 63) / 64, 64, 0, stream>>>(views, plan, nx, nu, B, eps, backup, n_min, n_max, new_n, ops, nops);
}

void launchAdaptApplyOps(const AdaptBucketView* views, const int* plan, const int* ops, const int* nops, double* scratch, const double* x0_master,
                         const double* xref_master, int nx, int nu, int B, cudaStream_t stream)
{
    adaptApplyOpsKernel<<<(B + 63) / 64, 64, 0, stream>>>(views, plan, ops, nops, scratch, x0_master, xref_master, nx, nu, B);
    adaptCommitKernel<<<(B + 127) / 128, 128, 0, stream>>>(views, plan, B);
}

}  // namespace b200sqp
