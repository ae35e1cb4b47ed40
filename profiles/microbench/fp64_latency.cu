// Microbenchmark (measurement aid, not product code): fp64 dependent-issue latency and per-SM throughput on the B200,
// the numbers that bound the sequential block-tridiagonal recursion of the LM kernel.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_latency fp64_latency.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void chain(double* out, long long* cyc, double a, double b, int n)
{
    double x = a + threadIdx.x * 1e-9, y = b;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; ++i)
    {
#pragma unroll
        for (int j = 0; j < 16; ++j)
        {
            if (OP == 0) x = fma(x, y, y);
            if (OP == 1) x = x * y;
            if (OP == 2) x = x + y;
            if (OP == 3) x = rsqrt(x) + y;   // library rsqrt: MUFU.RSQ64H + 5 fp64 ops (4 dependent) + 1 add
            if (OP == 4) x = y / x;          // IEEE division
            if (OP == 5) x = sqrt(x) + y;
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// ILP independent chains per thread, W warps per block: throughput
template <int ILP>
__global__ void tput(double* out, long long* cyc, double a, double b, int n)
{
    double x[ILP];
#pragma unroll
    for (int j = 0; j < ILP; ++j) x[j] = a + j + threadIdx.x * 1e-9;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; ++i)
    {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int j = 0; j < ILP; ++j) x[j] = fma(x[j], b, b);
    }
    __syncthreads();
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int j = 0; j < ILP; ++j) s += x[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void ldchase(const int* next, int* out, long long* cyc, int n)
{
    int i = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 1
    for (int k = 0; k < n; ++k) i = next[i];
    long long t1 = clock64();
    out[threadIdx.x] = i;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

int main()
{
    double* out;
    long long* cyc;
    cudaMalloc(&out, 1 << 20);
    cudaMalloc(&cyc, 1 << 16);
    long long h[256];
    const int n = 2000;
    const char* names[] = {"DFMA", "DMUL", "DADD", "rsqrt()+add", "y/x (IEEE div)", "sqrt()+add"};
#define RUN(OP)                                                                               \
    chain<OP><<<1, 32>>>(out, cyc, 1.000001, 0.999999, n);                                    \
    chain<OP><<<1, 32>>>(out, cyc, 1.000001, 0.999999, n);                                    \
    cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);                                            \
    printf("dependent chain %-16s %.1f cycles per op\n", names[OP], (double)h[0] / (n * 16.0));
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5)
    for (int warps = 1; warps <= 16; warps *= 2)
    {
        tput<8><<<1, 32 * warps>>>(out, cyc, 1.0, 0.999999, n);
        tput<8><<<1, 32 * warps>>>(out, cyc, 1.0, 0.999999, n);
        cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("throughput: %2d warps x ILP 8: %.3f warp-DFMA per cycle per SM\n", warps, (double)n * 4 * 8 * warps / h[0]);
    }
    for (int warps = 1; warps <= 8; warps *= 2)
    {
        tput<1><<<1, 32 * warps>>>(out, cyc, 1.0, 0.999999, n);
        tput<1><<<1, 32 * warps>>>(out, cyc, 1.0, 0.999999, n);
        cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("throughput: %2d warps x ILP 1: %.3f warp-DFMA per cycle per SM\n", warps, (double)n * 4 * 1 * warps / h[0]);
    }
    // pointer chase: L1-resident (small), L2-resident (64 MB with stride)
    {
        const size_t N = 16u << 20;  // 64 MB of ints
        int* hn = (int*)malloc(N * 4);
        for (size_t i = 0; i < N; ++i) hn[i] = (int)((i + 1048583u) % N);  // large stride: every hop misses L1
        int *dn, *dout;
        cudaMalloc(&dn, N * 4);
        cudaMalloc(&dout, 4096);
        cudaMemcpy(dn, hn, N * 4, cudaMemcpyHostToDevice);
        ldchase<<<1, 1>>>(dn, dout, cyc, 4000);
        ldchase<<<1, 1>>>(dn, dout, cyc, 4000);
        cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("dependent load, 64 MB working set (L2 hit after warm-up / HBM): %.0f cycles per load\n", (double)h[0] / 4000);
        for (size_t i = 0; i < 1024; ++i) hn[i] = (int)((i + 33) % 1024);
        cudaMemcpy(dn, hn, 4096, cudaMemcpyHostToDevice);
        ldchase<<<1, 1>>>(dn, dout, cyc, 4000);
        ldchase<<<1, 1>>>(dn, dout, cyc, 4000);
        cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("dependent load, 4 KB working set (L1 hit): %.0f cycles per load\n", (double)h[0] / 4000);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
