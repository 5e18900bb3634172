// DFMA latency / ILP micro-benchmark: W warps per SM sub-partition, C independent dependent-chains per thread.
// Reports cycles per DFMA instruction per warp and the fraction of the FP64 pipe peak -- how much instruction-level
// parallelism a kernel with 2-3 resident warps per scheduler (the row kernels) needs to fill the pipe.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/microbench_dfma tools/microbench_dfma.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int C>
__global__ void k(double *out, long long *cyc, int iters, double a, double b)
{
    double x[C];
#pragma unroll
    for (int c = 0; c < C; c++) x[c] = threadIdx.x * 1e-3 + c;
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int c = 0; c < C; c++) x[c] = fma(x[c], a, b);
    }
    const long long t1 = clock64();
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < C; c++) s += x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int C>
static void run(int warps_per_sm, double *out, long long *cyc)
{
    const int iters = 2000;
    k<C><<<148, 32 * warps_per_sm>>>(out, cyc, iters, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    k<C><<<148, 32 * warps_per_sm>>>(out, cyc, iters, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    long long h;
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double per_inst = (double)h / (iters * 8.0 * C);          // cycles per DFMA of one warp
    const double pipe = (warps_per_sm / 4.0) / per_inst / 0.5;      // fraction of 1 warp-DFMA per 2 cycles per sub-partition
    printf(" {\"chains\": %d, \"warps_per_sm\": %d, \"cycles_per_dfma_per_warp\": %.2f, \"fp64_pipe_frac\": %.2f},\n", C, warps_per_sm, per_inst, pipe);
}

int main()
{
    double *out; long long *cyc;
    cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 8);
    printf("{\"dfma\": [\n");
    for (int w : {4, 8, 12, 16, 32}) {
        run<1>(w, out, cyc); run<2>(w, out, cyc); run<4>(w, out, cyc); run<8>(w, out, cyc);
    }
    printf(" {}]}\n");
    return 0;
}
