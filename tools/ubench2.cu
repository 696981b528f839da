// Does the DFMA operand kind change co-issue?  Variants of x = fma(x, a, b) interleaved with J LOP3:
//   mode 0: a, b in registers (baseline)        mode 1: a, b from __constant__ (UR / c[] operands)
//   mode 2: three distinct rotating register operands x[r] = fma(x[r], x[r+1], x[r+2])
//   mode 3: like 1 plus FSEL (64-bit select = 2 FSEL) instead of LOP3
#include <cstdio>
#include <cuda_runtime.h>
__constant__ double kc[8] = {0.999999, 1e-9, 0.999998, 2e-9, 0.999997, 3e-9, 0.999996, 4e-9};

template <int MODE, int J>
__global__ void k(double *out, int *iout, int iters, double a, double b, int ia)
{
    constexpr int R = 4;
    double x[R + 2];
    int y[8];
    double w[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) w[r] = 1.0 + 1e-7 * (threadIdx.x + r) * a;
#pragma unroll
    for (int r = 0; r < R + 2; ++r) x[r] = 1.0 + 1e-3 * (threadIdx.x + r);
#pragma unroll
    for (int j = 0; j < 8; ++j) y[j] = threadIdx.x * (j + 1);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (MODE == 0) x[r] = __fma_rn(x[r], a, b);
                if (MODE == 1 || MODE == 3) x[r] = __fma_rn(x[r], kc[(2 * r) & 7], kc[(2 * r + 1) & 7]);
                if (MODE == 4) x[r] = __fma_rn(x[r], a, kc[(2 * r + 1) & 7]);        // one constant operand
                if (MODE == 5) x[r] = __fma_rn(x[r], a, 0.5);                        // immediate operand
                if (MODE == 6) x[r] = __fma_rn(x[r], x[(r + 1) & 3], x[(r + 2) & 3]);  // 3 distinct regs
                if (MODE == 7) x[r] = __fma_rn(x[r], a, 1.58969099521155010221e-10);   // literal -> UMOV or reg
                if (MODE == 8) x[r] = __fma_rn(x[r], a, w[r]);             // 2 fresh regs + 1 reused
                if (MODE == 9) x[r] = __fma_rn(x[r], w[r], w[r + 4]);      // 3 fresh regs (constants in regs)
                if (MODE == 10) x[r] = __dadd_rn(x[r], w[r]);              // DADD, 2 fresh
                if (MODE == 11) x[r] = __dmul_rn(x[r], w[r]);              // DMUL, 2 fresh
                if (MODE == 12) x[r] = __dmul_rn(x[r], a);                 // DMUL, 1 fresh + reused
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    int &v = y[(r * J + j) & 7];
                    if (MODE == 3) asm volatile("selp.b32 %0, %0, %1, %2;" : "+r"(v) : "r"(ia), "n"(1));
                    else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v) : "r"(ia), "r"(i));
                }
            }
        }
    }
    double s = 0; int t = 0;
#pragma unroll
    for (int r = 0; r < R + 2; ++r) s += x[r];
#pragma unroll
    for (int r = 0; r < 8; ++r) s += w[r];
#pragma unroll
    for (int j = 0; j < 8; ++j) t ^= y[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    iout[blockIdx.x * blockDim.x + threadIdx.x] = t;
}

template <int MODE, int J>
void run(double *out, int *iout, int threads, double fp64_per_iter)
{
    const int iters = 2000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        k<MODE, J><<<148, threads>>>(out, iout, iters, 0.999999, 1e-9, 12345);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    const double groups = 8.0 * 4 * iters * (double)threads * 148 / 32 / 592;  // (fp64 group) per SMSP
    printf("mode %d J=%d threads=%4d: %.2f cycles per group (%.0f FP64 + %d int)\n", MODE, J, threads,
           best * 1e-3 * 1.965e9 / groups, fp64_per_iter, J);
}

int main()
{
    double *out; int *iout;
    cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&iout, 148 * 1024 * 4);
    for (int th : {1024}) {
        run<0, 0>(out, iout, th, 1); run<0, 1>(out, iout, th, 1); run<0, 2>(out, iout, th, 1);
        run<1, 0>(out, iout, th, 1); run<1, 1>(out, iout, th, 1); run<1, 2>(out, iout, th, 1);
        run<4, 0>(out, iout, th, 1); run<4, 1>(out, iout, th, 1);
        run<5, 0>(out, iout, th, 1); run<5, 1>(out, iout, th, 1);
        run<6, 0>(out, iout, th, 1); run<6, 1>(out, iout, th, 1);
        run<7, 0>(out, iout, th, 1); run<7, 1>(out, iout, th, 1);
        run<8, 0>(out, iout, th, 1); run<8, 1>(out, iout, th, 1);
        run<9, 0>(out, iout, th, 1); run<9, 1>(out, iout, th, 1);
        run<10, 0>(out, iout, th, 1); run<10, 1>(out, iout, th, 1);
        run<11, 0>(out, iout, th, 1); run<11, 1>(out, iout, th, 1);
        run<12, 0>(out, iout, th, 1); run<12, 1>(out, iout, th, 1);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
