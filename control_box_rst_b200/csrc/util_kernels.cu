// Layout kernels around the solver: host-order <-> instance-minor conversions, trajectory initialisation, first-control gather.
// All are one-thread-per-instance (writes/reads of the instance-minor side coalesce; the host-order side is a short strided walk
// through L1/L2) -- they move a few MB per solve and are not on the roofline-relevant path.
#include "launch.h"

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

__global__ void unpackKernel(const double* __restrict__ z0, const double* __restrict__ z1, const int* __restrict__ cur,
                             const int* __restrict__ internal_of_ref, int n, int slots, double* __restrict__ params, int B, int S)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    const double* src = cur[i] ? z1 : z0;
    for (int r = 0; r < n; ++r) params[(size_t)i * n + r] = src[tiled(i, internal_of_ref[r], slots)];
}

__global__ void transposeInKernel(const double* __restrict__ src, int dim, double* __restrict__ dst, int B, int S)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    for (int j = 0; j < dim; ++j) dst[tiled(i, j, dim)] = src[(size_t)i * dim + j];
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
// dt_k = dt_ref, xf = xref.
__global__ void initTrajectoriesKernel(const double* __restrict__ x0, const double* __restrict__ xref, double* __restrict__ z, int* __restrict__ cur,
                                       int K, int nx, int nu, int vt, double dt_ref, int B, int S)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
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
    for (int k = 0; k < K; ++k)
    {
        for (int j = 0; j < nu; ++j) z[tiled(i, k * nb + j, slots)] = 0.0;
        if (vt) z[tiled(i, k * nb + nu, slots)] = dt_ref;
        for (int j = 0; j < nx; ++j)
        {
            const double a = x0[tiled(i, j, nx)], b = xref[tiled(i, j, nx)];
            double dir     = b - a;
            if (dist != 0) dir /= dist;
            // block k holds x_{k+1}; the last block holds xf = xref
            z[tiled(i, k * nb + nu + vt + j, slots)] = (k + 1 < K) ? a + (double)(k + 1) * step * dir : b;
        }
    }
    cur[i] = 0;
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

}  // namespace

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
    unpackKernel<<<blocksFor(B), 128, 0, st>>>(z0, z1, cur, internal_of_ref, n, slots, params, B, S);
}
void launchFillPinned(const double* xref, double* z0, double* z1, int slot0, int slots, int nx, unsigned mask, int B, cudaStream_t st)
{
    fillPinnedKernel<<<blocksFor(B), 128, 0, st>>>(xref, z0, z1, slot0, slots, nx, mask, B);
}
void launchTransposeIn(const double* src, int dim, double* dst, int B, int S, cudaStream_t st)
{
    transposeInKernel<<<blocksFor(B), 128, 0, st>>>(src, dim, dst, B, S);
}
void launchTransposeOut(const double* src, int rows, double* dst, int B, int S, cudaStream_t st)
{
    transposeOutKernel<<<blocksFor(B), 128, 0, st>>>(src, rows, dst, B, S);
}
void launchInitTrajectories(const double* x0, const double* xref, double* z, int* cur, int K, int nx, int nu, int vt, double dt_ref,
                            const int* /*xf_fixed_dev*/, int B, int S, cudaStream_t st)
{
    initTrajectoriesKernel<<<blocksFor(B), 128, 0, st>>>(x0, xref, z, cur, K, nx, nu, vt, dt_ref, B, S);
}
void launchFirstControls(const double* z0, const double* z1, const int* cur, int nu, int slots, double* u0, int B, int S, cudaStream_t st)
{
    firstControlsKernel<<<blocksFor(B), 128, 0, st>>>(z0, z1, cur, nu, slots, u0, B, S);
}

}  // namespace b200sqp
