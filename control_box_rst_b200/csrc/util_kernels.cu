// Layout kernels around the solver: host-order <-> instance-minor conversions, trajectory initialisation, first-control gather.
// All are one-thread-per-instance (writes/reads of the instance-minor side coalesce; the host-order side is a short strided walk
// through L1/L2) -- they move a few MB per solve and are not on the roofline-relevant path.
#include "../../include/b200sqp.h"
#include "launch.h"
#include "lm_device_types.h"

namespace b200sqp {

namespace {

// tiled instance-minor addressing (lm_device.cuh TILE): element `slot` of instance i in an array with `nslots` slots per instance
__device__ __forceinline__ size_t tiled(int i, int slot, int nslots) { return ((size_t)(i >> 5) * nslots + slot) * 32 + (i & 31); }

__global__ void packKernel(const double* __restrict__ params, int n, const int* __restrict__ ref_of_internal, int slots,
                           const double* __restrict__ pinned, double* __restrict__ z, const int* __restrict__ cur, double* __restrict__ z_alt, int B,
                           int S)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    double* dst = (cur && cur[i]) ? z_alt : z;
    for (int s = 0; s < slots; ++s)
    {
        const int r = ref_of_internal[s];
        if (r >= 0) dst[tiled(i, s, slots)] = params[(size_t)i * n + r];  // pinned slots (fixed goal components) are left alone
    }
}

// block order (tiled instance-minor) -> host order [B][n]: one thread block transposes 32 instances x 32 parameters through shared
// memory so that both the reads (32 consecutive lanes of a slot) and the writes (32 consecutive parameters of an instance) coalesce
__global__ void unpackKernel(const double* __restrict__ z0, const double* __restrict__ z1, const int* __restrict__ cur,
                             const int* __restrict__ internal_of_ref, int n, int slots, double* __restrict__ params, int B, int S)
{
    __shared__ double tile[32][33];
    const int lane = threadIdx.x, ty = threadIdx.y;
    const int i0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    const int i  = i0 + lane;
    if (i < B)
    {
        const double* src = cur[i] ? z1 : z0;
        for (int rr = ty; rr < 32; rr += 8)
            if (r0 + rr < n) tile[rr][lane] = src[tiled(i, internal_of_ref[r0 + rr], slots)];
    }
    __syncthreads();
    for (int ii = ty; ii < 32; ii += 8)
        if (i0 + ii < B && r0 + lane < n) params[(size_t)(i0 + ii) * n + r0 + lane] = tile[lane][ii];
}

__global__ void transposeInKernel(const double* __restrict__ src, int dim, double* __restrict__ dst, int B, int S)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    for (int j = 0; j < dim; ++j) dst[tiled(i, j, dim)] = src[(size_t)i * dim + j];
}

// Start states and references of a step in one launch: [B][nx] (device memory, or pinned host memory read in place over PCIe) ->
// tiled [tile][nx][32], plus the host-order device copies the other entry points use.  xref_src null = zero reference.
__global__ void ingestKernel(const double* __restrict__ x0_src, const double* __restrict__ xref_src, int nx, double* __restrict__ x0_tiled,
                             double* __restrict__ xref_tiled, double* __restrict__ x0_copy, double* __restrict__ xref_copy, int B)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    for (int j = 0; j < nx; ++j)
    {
        const double v         = x0_src[(size_t)i * nx + j];
        x0_tiled[tiled(i, j, nx)] = v;
        if (x0_copy != x0_src) x0_copy[(size_t)i * nx + j] = v;
        const double r = xref_src ? xref_src[(size_t)i * nx + j] : 0.0;
        xref_tiled[tiled(i, j, nx)] = r;
        if (xref_src && xref_copy != xref_src) xref_copy[(size_t)i * nx + j] = r;
    }
}

// Small per-instance results of a step in one launch: first controls -> u0_dev [B][nu] (always) and, where the caller's buffers
// are pinned host memory, u0 / chi2 / status written in place over PCIe (any may be null)
__global__ void exportKernel(const double* __restrict__ z0, const double* __restrict__ z1, const int* __restrict__ cur, int nu, int slots,
                             const double* __restrict__ chi2, const int* __restrict__ status, double* __restrict__ u0_dev, double* __restrict__ u0_out,
                             double* __restrict__ chi2_out, int* __restrict__ status_out, int B)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    if (u0_dev || u0_out)
    {
        const double* src = cur[i] ? z1 : z0;
        for (int j = 0; j < nu; ++j)
        {
            const double v = src[tiled(i, j, slots)];
            if (u0_dev) u0_dev[(size_t)i * nu + j] = v;
            if (u0_out) u0_out[(size_t)i * nu + j] = v;
        }
    }
    if (chi2_out) chi2_out[i] = chi2[i];
    if (status_out) status_out[i] = status[i];
}

__global__ void transposeOutKernel(const double* __restrict__ src, int rows, double* __restrict__ dst, int B, int S)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    for (int j = 0; j < rows; ++j) dst[(size_t)i * rows + j] = src[(size_t)j * S + i];
}

// FullDiscretizationGridBase::initializeSequences (optimal_control/src/structured_ocp/discretization_grids/
// full_discretization_grid_base.cpp:134-179; same code in non_uniform_full_discretization_grid_base.cpp:146-190 and
// shooting_grid_base.cpp:141-200): dir = (xf - x0)/||xf - x0||, step = ||xf - x0||/(N-1), x_k = x0 + k*step*dir, u_k = uref = 0,
// dt_k = dt_ref, xf = xref.  With a non-static reference (xtraj: [(K+1)*nx] slots per instance, row m = getReferenceCached(m)) the reference
// trajectory itself is the initial guess: x_k = xref(k), k >= 1 (full_discretization_grid_base.cpp:181-228).
__global__ void initTrajectoriesKernel(const double* __restrict__ x0, const double* __restrict__ xref, const double* __restrict__ xtraj,
                                       double* __restrict__ z, int* __restrict__ cur, int K, int nx, int nu, int vt, double dt_ref, int B, int S)
{
    // one thread per (instance, interval): the batch alone (4096 threads) would leave most of the machine idle
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y;
    if (i >= B) return;
    const int nb = nu + vt + nx, slots = K * nb;
    double dist  = 0.0;
    for (int j = 0; j < nx; ++j)
    {
        const double d = xref[tiled(i, j, nx)] - x0[tiled(i, j, nx)];
        dist += d * d;
    }
    dist              = sqrt(dist);
    const double step = dist / K;
    for (int j = 0; j < nu; ++j) z[tiled(i, k * nb + j, slots)] = 0.0;
    if (vt) z[tiled(i, k * nb + nu, slots)] = dt_ref;
    for (int j = 0; j < nx; ++j)
    {
        const double a = x0[tiled(i, j, nx)], b = xref[tiled(i, j, nx)];
        double dir     = b - a;
        if (dist != 0) dir /= dist;
        // block k holds x_{k+1}; the last block holds xf = xref
        z[tiled(i, k * nb + nu + vt + j, slots)] = xtraj ? xtraj[tiled(i, (k + 1) * nx + j, (K + 1) * nx)] : ((k + 1 < K) ? a + (double)(k + 1) * step * dir : b);
    }
    if (k == 0) cur[i] = 0;
}

__global__ void firstControlsKernel(const double* __restrict__ z0, const double* __restrict__ z1, const int* __restrict__ cur, int nu, int slots,
                                    double* __restrict__ u0, int B, int S)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    const double* src = cur[i] ? z1 : z0;
    for (int j = 0; j < nu; ++j) u0[(size_t)i * nu + j] = src[tiled(i, j, slots)];
}

// fixed goal components take their value from the reference: _xf.values()[i] = xref[i] (full_discretization_grid_base.cpp:102-106)
__global__ void fillPinnedKernel(const double* __restrict__ xref, double* __restrict__ z0, double* __restrict__ z1, int slot0, int slots, int nx,
                                 unsigned mask, int B)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    for (int j = 0; j < nx; ++j)
        if (mask & (1u << j))
        {
            const double v               = xref[tiled(i, j, nx)];
            z0[tiled(i, slot0 + j, slots)] = v;
            z1[tiled(i, slot0 + j, slots)] = v;
        }
}

// Moving-horizon warm start of FullDiscretizationGridBase (full_discretization_grid_base.cpp): findNearestState (:285-318) picks
// num_shift per instance (nearest of the first min(N-2, 20) states to the new measurement, l2 norm in Eigen's reduction order,
// stop at the first non-improving state, 0 if the start did not move), warmStartShifting (:230-283) moves states and controls
// forward by num_shift and extrapolates the tail linearly (x[idx] = x[idx-2] + 2 (x[idx-1] - x[idx-2]), controls repeated), then
// update() overwrites the start with the measurement (:101).  x_seq[0] is the fixed start (x0 array), x_seq[k>=1] the x-part of
// block k-1, u_seq[k] the u-part of block k.  One thread per instance, in place on the current parameter buffer.
__device__ __forceinline__ double eigenNorm(const double* t, int n)
{
    // sqrt of Eigen's vectorised sum of squares: packets of two doubles, two accumulators, scalar tail (Eigen/src/Core/Redux.h)
    const int aligned = (n / 2) * 2, aligned2 = (n / 4) * 4;
    if (aligned == 0) return sqrt(t[0]);
    double p0a = t[0], p0b = t[1];
    if (aligned > 2)
    {
        double p1a = t[2], p1b = t[3];
        for (int i = 4; i < aligned2; i += 4)
        {
            p0a += t[i];
            p0b += t[i + 1];
            p1a += t[i + 2];
            p1b += t[i + 3];
        }
        p0a += p1a;
        p0b += p1b;
        if (aligned > aligned2)
        {
            p0a += t[aligned2];
            p0b += t[aligned2 + 1];
        }
    }
    double res = p0a + p0b;
    for (int i = aligned; i < n; ++i) res += t[i];
    return sqrt(res);
}

// Phase 1, one thread per instance: findNearestState against the current trajectory -> shift (0 = none), the buffer the trajectory
// lives in (cur) packed next to it for phase 2, and x_seq.front() = measurement (update :101).
__global__ void warmStartFindKernel(const double* __restrict__ x0_new /*[B][nx] host order*/, double* __restrict__ x0 /*tiled*/,
                                    const double* __restrict__ z0, const double* __restrict__ z1, const int* __restrict__ cur, int K, int nx, int nu,
                                    int* __restrict__ plan /*[B]: shift | src buffer << 8*/, int* __restrict__ num_shift_out, int B)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    const int nb = nu + nx, slots = K * nb, N = K + 1;
    const int src   = cur[i];
    const double* z = src ? z1 : z0;
    auto X = [&](int k, int j) -> double { return k == 0 ? x0[tiled(i, j, nx)] : z[tiled(i, (k - 1) * nb + nu + j, slots)]; };
    double xn[B200SQP_MAX_NX], sq[B200SQP_MAX_NX];
    for (int j = 0; j < nx; ++j) xn[j] = x0_new[(size_t)i * nx + j];
    auto dist = [&](int k) {
        for (int j = 0; j < nx; ++j)
        {
            const double d = xn[j] - X(k, j);
            sq[j]          = d * d;
        }
        return eigenNorm(sq, nx);
    };
    int shift               = 0;
    const double first_dist = dist(0);
    if (!(fabs(first_dist) < 1e-12))
    {
        const int lookahead = min((N - 1) - 1, 20);
        double cache        = first_dist;
        for (int k = 1; k <= lookahead; ++k)
        {
            const double d = dist(k);
            if (d < cache)
            {
                cache = d;
                shift = k;
            }
            else
                break;
        }
    }
    plan[i] = shift | (src << 8);
    if (num_shift_out) num_shift_out[i] = shift;
    // the shifted x_seq[0] is never read again (phase 2 takes its values from the old buffer): only the measurement lands here
    for (int j = 0; j < nx; ++j) x0[tiled(i, j, nx)] = xn[j];
}

// Phase 2, one thread per (instance, block k): warmStartShifting written out of place into the instance's other parameter buffer
// (the LM loop's trial buffer, free between solves), then the instance's buffer roles swap.  With s = shift, block k = [u_k, x_{k+1}]:
//   x_{j}   <- x_old_{j+s}                         for j <  N-s          (x_old_{N-1} = xf)
//   u_{k}   <- u_old_{k+s}                         for k <  N-1-s
//   x_{idx} <- a + 2 (b - a) chained from a = x_old_{N-2}, b = x_old_{N-1}   for idx = N-s .. N-1 (the reference's linear extrapolation;
//              every thread replays the chain up to its own idx: same operations in the same order, no cross-thread dependence)
//   u_{k}   <- u_old_{N-2}                         for k >= N-1-s        (u[idx-1] = u[idx-2] repeated)
// Instances with s = 0 or s > N-2 are left untouched (:233-239).
__global__ void warmStartMoveKernel(double* __restrict__ z0, double* __restrict__ z1, int* __restrict__ cur, const int* __restrict__ plan, int K,
                                    int nx, int nu, int B)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y;
    if (i >= B) return;
    const int nb = nu + nx, slots = K * nb, N = K + 1;
    const int s = plan[i] & 0xff, src = plan[i] >> 8;
    if (s <= 0 || s > N - 2) return;
    const double* zo = src ? z1 : z0;
    double* zn       = src ? z0 : z1;
    auto Xo = [&](int q, int j) -> double { return zo[tiled(i, (q - 1) * nb + nu + j, slots)]; };  // q >= 1 only
    // controls of block k
    const int ku = (k < N - 1 - s) ? k + s : N - 2;
    for (int j = 0; j < nu; ++j) zn[tiled(i, k * nb + j, slots)] = zo[tiled(i, ku * nb + j, slots)];
    // state x_{k+1}
    const int q = k + 1;
    if (q < N - s)
    {
        for (int j = 0; j < nx; ++j) zn[tiled(i, k * nb + nu + j, slots)] = Xo(q + s, j);
    }
    else
    {
        for (int j = 0; j < nx; ++j)
        {
            double a = Xo(N - 2, j), b = Xo(N - 1, j), c = 0.0;
            for (int idx = N - s; idx <= q; ++idx)
            {
                c = a + 2.0 * (b - a);
                a = b;
                b = c;
            }
            zn[tiled(i, k * nb + nu + j, slots)] = c;
        }
    }
    if (k == 0) cur[i] = 1 - src;  // nobody reads cur in this kernel (the plan carries the source buffer)
}

inline int blocksFor(int B) { return (B + 127) / 128; }

// b200sqp_peer_wait: one thread per rank spins (bounded by `timeout_ns` of %globaltimer) until that rank's arrival counter in
// OUR memory reaches `expected`, i.e. until all of its thread blocks have stored their chi2 of this solve into our gather buffer.
__global__ void peerWaitKernel(const volatile unsigned long long* arrivals, int world, unsigned long long expected, unsigned long long timeout_ns,
                               int* timed_out)
{
    const int r = threadIdx.x;
    if (r < world)
    {
        unsigned long long t0, now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        while (arrivals[r] < expected)
        {
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (now - t0 > timeout_ns)
            {
                *timed_out = 1;
                break;
            }
            __nanosleep(100);
        }
    }
    __threadfence_system();
}

// epilogue of lmSolveKernel (lm_kernels.cuh) as a kernel of its own, for solves that ran the multi-kernel pipeline: one warp per 32
// instances stores chi2 into every rank's gather buffer, fences at system scope and bumps this rank's arrival counter on every peer
__global__ void __launch_bounds__(32) peerPublishKernel(const __grid_constant__ DeviceState st, int B)
{
    const int i = blockIdx.x * 32 + threadIdx.x;
    if (i < B)
    {
        const size_t slot = (size_t)st.peer_parity * st.peer_world * B + (size_t)st.peer_rank * B + i;
        const double c    = st.chi2[i];
        for (int r = 0; r < st.peer_world; ++r) st.peer_chi2[r][slot] = c;
        __threadfence_system();
    }
    __syncwarp();
    if (threadIdx.x == 0)
        for (int r = 0; r < st.peer_world; ++r) atomicAdd_system(st.peer_arrivals[r] + st.peer_rank, 1ULL);
}

// b200sqp_measure_fp64_peak: 8 independent DFMA chains per thread, 16 warps per block, every SM loaded: the fp64 FMA issue bound
// the fused LM kernel is compared with (bench.py roofline_fp64).  fma() is explicit, so --fmad=false does not matter here.
__global__ void __launch_bounds__(512) fp64PeakKernel(double* out, double a, double b, int n)
{
    double x[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = a + j + threadIdx.x * 1e-9;
#pragma unroll 1
    for (int i = 0; i < n; ++i)
    {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] = fma(x[j], b, b);
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += x[j];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace

double measureFp64PeakTflops(int sm_count, cudaStream_t st)
{
    const int blocks = sm_count * 4, threads = 512, n = 20000;
    double* out = nullptr;
    if (cudaMalloc(&out, sizeof(double) * (size_t)blocks * threads) != cudaSuccess) return -1.0;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best = -1.0;
    for (int rep = 0; rep < 4; ++rep)  // first repetition warms the clocks up
    {
        cudaEventRecord(e0, st);
        fp64PeakKernel<<<blocks, threads, 0, st>>>(out, 1.0, 0.999999, n);
        cudaEventRecord(e1, st);
        if (cudaEventSynchronize(e1) != cudaSuccess) break;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double tflops = 2.0 * blocks * threads * (double)n * 32.0 / (ms * 1e-3) / 1e12;
        if (rep > 0 && tflops > best) best = tflops;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    return cudaGetLastError() == cudaSuccess ? best : -1.0;
}

void launchWarmStartShift(const double* x0_new, double* x0, double* z0, double* z1, int* cur, int K, int nx, int nu, int* plan, int* num_shift,
                          int B, cudaStream_t st)
{
    warmStartFindKernel<<<blocksFor(B), 128, 0, st>>>(x0_new, x0, z0, z1, cur, K, nx, nu, plan, num_shift, B);
    warmStartMoveKernel<<<dim3(blocksFor(B), K), 128, 0, st>>>(z0, z1, cur, plan, K, nx, nu, B);
}

void launchPeerPublish(const DeviceState& st, int B, cudaStream_t stream) { peerPublishKernel<<<(B + 31) / 32, 32, 0, stream>>>(st, B); }

void launchPeerWait(const unsigned long long* arrivals, int world, unsigned long long expected, unsigned long long timeout_ns, int* timed_out,
                    cudaStream_t st)
{
    peerWaitKernel<<<1, 32, 0, st>>>(arrivals, world, expected, timeout_ns, timed_out);
}

void launchPack(const double* params, int n, const int* ref_of_internal, int slots, const double* pinned, double* z, const int* cur, double* z_alt,
                int B, int S, cudaStream_t st)
{
    packKernel<<<blocksFor(B), 128, 0, st>>>(params, n, ref_of_internal, slots, pinned, z, cur, z_alt, B, S);
}
void launchUnpack(const double* z0, const double* z1, const int* cur, const int* internal_of_ref, int n, int slots, double* params, int B, int S,
                  cudaStream_t st)
{
    unpackKernel<<<dim3((B + 31) / 32, (n + 31) / 32), dim3(32, 8), 0, st>>>(z0, z1, cur, internal_of_ref, n, slots, params, B, S);
}
void launchFillPinned(const double* xref, double* z0, double* z1, int slot0, int slots, int nx, unsigned mask, int B, cudaStream_t st)
{
    fillPinnedKernel<<<blocksFor(B), 128, 0, st>>>(xref, z0, z1, slot0, slots, nx, mask, B);
}
void launchTransposeIn(const double* src, int dim, double* dst, int B, int S, cudaStream_t st)
{
    transposeInKernel<<<blocksFor(B), 128, 0, st>>>(src, dim, dst, B, S);
}
void launchIngest(const double* x0_src, const double* xref_src, int nx, double* x0_tiled, double* xref_tiled, double* x0_copy, double* xref_copy,
                  int B, cudaStream_t st)
{
    ingestKernel<<<blocksFor(B), 128, 0, st>>>(x0_src, xref_src, nx, x0_tiled, xref_tiled, x0_copy, xref_copy, B);
}
void launchExport(const double* z0, const double* z1, const int* cur, int nu, int slots, const double* chi2, const int* status, double* u0_dev,
                  double* u0_out, double* chi2_out, int* status_out, int B, cudaStream_t st)
{
    exportKernel<<<blocksFor(B), 128, 0, st>>>(z0, z1, cur, nu, slots, chi2, status, u0_dev, u0_out, chi2_out, status_out, B);
}
void launchTransposeOut(const double* src, int rows, double* dst, int B, int S, cudaStream_t st)
{
    transposeOutKernel<<<blocksFor(B), 128, 0, st>>>(src, rows, dst, B, S);
}
void launchInitTrajectories(const double* x0, const double* xref, const double* xtraj, double* z, int* cur, int K, int nx, int nu, int vt, double dt_ref,
                            int B, int S, cudaStream_t st)
{
    initTrajectoriesKernel<<<dim3(blocksFor(B), K), 128, 0, st>>>(x0, xref, xtraj, z, cur, K, nx, nu, vt, dt_ref, B, S);
}
void launchFirstControls(const double* z0, const double* z1, const int* cur, int nu, int slots, double* u0, int B, int S, cudaStream_t st)
{
    firstControlsKernel<<<blocksFor(B), 128, 0, st>>>(z0, z1, cur, nu, slots, u0, B, S);
}

}  // namespace b200sqp
