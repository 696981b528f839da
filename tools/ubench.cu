// Micro-benchmarks that settle what bounds the likelihood kernel on sm_100a:
//   (1) DFMA throughput with R independent chains per thread, W warps per SM sub-partition
//   (2) DFMA interleaved with J integer instructions per DFMA (do they co-issue for free?)
//   (3) dependent-issue latency of DFMA / DADD / DMUL / MUFU.RCP64H
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ubench tools/ubench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int R, int J>
__global__ void k_mix(double *out, int *iout, int iters, double a, double b, int ia)
{
    double x[R];
    int y[8];
#pragma unroll
    for (int r = 0; r < R; ++r) x[r] = threadIdx.x + r;
#pragma unroll
    for (int j = 0; j < 8; ++j) y[j] = threadIdx.x * (j + 1);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                x[r] = __fma_rn(x[r], a, b);
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    int &v = y[(r * J + j) & 7];
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v) : "r"(ia), "r"(i));
                }
            }
        }
    }
    double s = 0;
    int t = 0;
#pragma unroll
    for (int r = 0; r < R; ++r) s += x[r];
#pragma unroll
    for (int j = 0; j < 8; ++j) t ^= y[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    iout[blockIdx.x * blockDim.x + threadIdx.x] = t;
}

template <int OP>
__global__ void k_lat(double *out, long long *cyc, int iters, double a, double b)
{
    double x = threadIdx.x + 1.5;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            if (OP == 0) x = __fma_rn(x, a, b);
            if (OP == 1) x = __dadd_rn(x, b);
            if (OP == 2) x = __dmul_rn(x, a);
            if (OP == 3) { double y; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); x = y; }
        }
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

template <int R, int J>
void run_mix(int warps_per_smsp, double *out, int *iout)
{
    const int iters = 2000;
    const int threads = warps_per_smsp * 4 * 32;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        k_mix<R, J><<<148, threads>>>(out, iout, iters, 0.999999, 1e-9, 12345);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    const double dfma = 8.0 * R * iters * (double)threads * 148;
    printf("R=%d chains, J=%d int/DFMA, %2d warps/SMSP: %7.2f TFLOP/s  (%.2f DFMA/clk/SM at 1.965 GHz, %.2f inst/clk/SMSP)\n",
           R, J, warps_per_smsp, 2 * dfma / (best * 1e-3) / 1e12, dfma / 148 / (best * 1e-3) / 1.965e9,
           dfma * (1 + J) / 32 / 592 / (best * 1e-3) / 1.965e9);
}

int main()
{
    double *out; int *iout; long long *cyc;
    cudaMalloc(&out, 148 * 1024 * sizeof(double));
    cudaMalloc(&iout, 148 * 1024 * sizeof(int));
    cudaMalloc(&cyc, 8);
    for (int w : {1, 2, 4, 8}) run_mix<1, 0>(w, out, iout);
    for (int w : {1, 2, 4, 8}) run_mix<2, 0>(w, out, iout);
    for (int w : {1, 2, 4, 8}) run_mix<4, 0>(w, out, iout);
    for (int w : {4, 8}) run_mix<4, 1>(w, out, iout);
    for (int w : {4, 8}) run_mix<4, 2>(w, out, iout);
    for (int w : {4, 8}) run_mix<4, 3>(w, out, iout);
    for (int w : {8}) run_mix<2, 1>(w, out, iout);
    for (int w : {8}) run_mix<2, 2>(w, out, iout);
    const char *names[] = {"DFMA", "DADD", "DMUL", "MUFU.RCP64H"};
    for (int op = 0; op < 4; ++op) {
        long long h = 0;
        const int iters = 1000;
        if (op == 0) k_lat<0><<<1, 32>>>(out, cyc, iters, 0.999999, 1e-9);
        if (op == 1) k_lat<1><<<1, 32>>>(out, cyc, iters, 0.999999, 1e-9);
        if (op == 2) k_lat<2><<<1, 32>>>(out, cyc, iters, 0.999999, 1e-9);
        if (op == 3) k_lat<3><<<1, 32>>>(out, cyc, iters, 0.999999, 1e-9);
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%-12s dependent-issue latency: %.2f cycles\n", names[op], (double)h / (iters * 16.0));
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
