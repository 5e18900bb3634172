// Write-order micro-benchmark: CSR rows of 1944 bytes (9 L doubles, L = 27) written once by one warp each with 16-byte
// stores, in an order that is random inside a moving WINDOW of the address space and in address order from window to
// window; optionally only every `skip`-th row is written (the rows of one bucket: the others belong to other launches).
// Answers: how compact must the set of rows in flight be for the write stream to reach the in-order rate?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/microbench_window tools/microbench_window.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(64) k_rows(double *out, size_t nrows, int len, size_t wrows, size_t mult, int skip, int rows_per_warp)
{
    const int lane = threadIdx.x & 31;
    const size_t nwarps = (size_t)gridDim.x * (blockDim.x >> 5), w = blockIdx.x * (size_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
    const size_t ntask = nrows / skip / rows_per_warp;
    for (size_t t = w; t < ntask; t += nwarps) {
        for (int r = 0; r < rows_per_warp; r++) {
            size_t c = t * rows_per_warp + r;                 // logical row index (bucket order)
            // permute inside the window of wrows rows
            const size_t win = c / wrows, in = c % wrows;
            const size_t wr = (win + 1) * wrows <= nrows / skip ? wrows : (nrows / skip - win * wrows);
            const size_t cc = (win * wrows + (in * mult) % wr) * skip;
            double *p = out + cc * (size_t)len;
            const int h = (int)((reinterpret_cast<uintptr_t>(p) >> 3) & 1);
            const int body = (len - h) & ~1;
            if (lane == 0 && h) p[0] = 0.0;
            if (lane == 1 && h + body < len) p[len - 1] = 0.0;
            double2 *q = reinterpret_cast<double2 *>(p + h);
            for (int x = lane; x < body / 2; x += 32) q[x] = make_double2(1.0, 2.0);
        }
    }
}

template <class F>
static float time_ms(F f, int reps = 4)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    return best;
}

int main()
{
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    const size_t bytes = (size_t)4 << 30;
    double *out;
    if (cudaMalloc(&out, bytes + 4096) != cudaSuccess) { printf("{\"error\": \"alloc\"}\n"); return 1; }
    const int len = 243;
    const size_t nrows = bytes / (len * 8);
    printf("{\"gpu\": \"%s\", \"row_bytes\": %d, \"results\": [\n", prop.name, len * 8);
    const size_t wins_mb[] = {0, 1, 4, 16, 64, 256, 4096};
    bool first = true;
    for (int skip = 1; skip <= 2; skip++)
        for (int rpw : {1, 10})
            for (int blocks_per_sm : {5, 8})
                for (size_t wmb : wins_mb) {
                    size_t wrows = wmb == 0 ? 1 : (wmb << 20) / (len * 8) / skip;
                    if (wrows > nrows / skip) wrows = nrows / skip;
                    const float t = time_ms([&] { k_rows<<<sms * blocks_per_sm, 64>>>(out, nrows, len, wrows, 7919, skip, rpw); });
                    const double gb = (double)(nrows / skip / rpw * rpw) * len * 8 / 1e9;
                    printf("%s {\"skip\": %d, \"rows_per_warp\": %d, \"blocks_per_sm\": %d, \"window_MB\": %zu, \"GBs\": %.0f}", first ? "" : ",\n", skip, rpw,
                           blocks_per_sm, wmb, gb / (t * 1e-3));
                    first = false;
                }
    printf("\n]}\n");
    return 0;
}
