// Pipe-rate micro-benchmarks on the bench GPU: FP64 FMA peak (the denominator of the FP64-pipe utilisation that
// BASELINE.md asks for), warp shuffles and shared-memory loads/stores.  One JSON line on stdout.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/microbench tools/microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void __launch_bounds__(256) k_dfma(double *out, double a, double b, int iters)
{
    double x[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; c++) x[c] = a + c + threadIdx.x;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int c = 0; c < CHAINS; c++) x[c] = fma(x[c], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < CHAINS; c++) s += x[c];
    out[blockIdx.x * (size_t)blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_shfl(double *out, int iters)
{
    unsigned x = threadIdx.x, y = threadIdx.x * 3, z = threadIdx.x * 7, w = threadIdx.x * 11;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
            x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31) + 1;
            y = __shfl_sync(0xffffffffu, y, (threadIdx.x + 2) & 31) + 1;
            z = __shfl_sync(0xffffffffu, z, (threadIdx.x + 3) & 31) + 1;
            w = __shfl_sync(0xffffffffu, w, (threadIdx.x + 5) & 31) + 1;
        }
    }
    out[blockIdx.x * (size_t)blockDim.x + threadIdx.x] = (double)(x + y + z + w);
}

// shared-memory read-modify-write of doubles at lane stride `stride` doubles (3 = the block layout of the assembly rows)
__global__ void __launch_bounds__(256) k_smem_rmw(double *out, int iters, int stride)
{
    extern __shared__ double s[];
    for (int x = threadIdx.x; x < 256 * 4; x += blockDim.x) s[x] = 0.0;
    __syncthreads();
    double *p = s + (threadIdx.x >> 5) * 128 + (threadIdx.x & 31) * stride % 128;
    double v = threadIdx.x;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 8; r++) { p[0] += v; p[1] += v; p[2] += v; }
    }
    __syncthreads();
    out[blockIdx.x * (size_t)blockDim.x + threadIdx.x] = s[threadIdx.x];
}

template <class F>
static float time_ms(F f, int reps = 5)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        cudaEventRecord(a);
        f();
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    return best;
}

int main()
{
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const int sms = prop.multiProcessorCount;
    double *out;
    cudaMalloc(&out, sizeof(double) * 256 * sms * 16);
    const int iters = 4096;
    const int blocks = sms * 8;
    float t8 = time_ms([&] { k_dfma<8><<<blocks, 256>>>(out, 1.0000001, 1e-9, iters); });
    float t4 = time_ms([&] { k_dfma<4><<<blocks, 256>>>(out, 1.0000001, 1e-9, iters); });
    const double fl8 = 2.0 * 8 * 8 * iters * 256.0 * blocks, fl4 = 2.0 * 4 * 8 * iters * 256.0 * blocks;
    const double tf8 = fl8 / (t8 * 1e-3) / 1e12, tf4 = fl4 / (t4 * 1e-3) / 1e12;
    float ts = time_ms([&] { k_shfl<<<blocks, 256>>>(out, iters); });
    const double shfl_per_clk_sm = 4.0 * 8 * iters * 8.0 * blocks / sms / (ts * 1e-3 * clk_khz * 1e3); // warp-instructions per clock per SM (256 thr = 8 warps)
    cudaFuncSetAttribute(k_smem_rmw, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * 4 * 8);
    float tr3 = time_ms([&] { k_smem_rmw<<<blocks, 256, 256 * 4 * 8>>>(out, iters, 3); });
    float tr1 = time_ms([&] { k_smem_rmw<<<blocks, 256, 256 * 4 * 8>>>(out, iters, 1); });
    const double rmw = 3.0 * 8 * iters * 8.0 * blocks / sms;   // warp-level RMWs (LDS.64 + DADD + STS.64) per SM
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz_attr\": %d, \"dfma_tflops_8chains\": %.2f, \"dfma_tflops_4chains\": %.2f, "
           "\"dfma_per_clk_sm_at_attr_clock\": %.1f, \"shfl_warp_instr_per_clk_sm\": %.3f, "
           "\"smem_rmw64_warp_per_clk_sm_stride3\": %.3f, \"smem_rmw64_warp_per_clk_sm_stride1\": %.3f}\n",
           prop.name, sms, clk_khz, tf8, tf4, tf8 * 1e12 / 2.0 / sms / (clk_khz * 1e3), shfl_per_clk_sm,
           rmw / (tr3 * 1e-3 * clk_khz * 1e3), rmw / (tr1 * 1e-3 * clk_khz * 1e3));
    return 0;
}
