// Write-order micro-benchmark 2: which locality does the write stream of CSR rows need to reach the in-order rate?
// Rows of `len` doubles written once by one warp each with 16-byte stores.
//   mode 0  chunks of N consecutive rows per warp, rows in order inside the chunk, chunks in random order
//   mode 1  same, but inside the chunk the even rows first, then the odd rows (two row types, same warp)
//   mode 2  chunk shared by the two warps of a block: warp 0 the even rows, warp 1 the odd rows, at the same time
//   mode 3  mode 0 with chunks in address order (reference: in-order)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/microbench_chunk tools/microbench_chunk.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void write_row(double *p, int len, int lane)
{
    const int h = (int)((reinterpret_cast<uintptr_t>(p) >> 3) & 1);
    const int body = (len - h) & ~1;
    if (lane == 0 && h) p[0] = 0.0;
    if (lane == 1 && h + body < len) p[len - 1] = 0.0;
    double2 *q = reinterpret_cast<double2 *>(p + h);
    for (int x = lane; x < body / 2; x += 32) q[x] = make_double2(1.0, 2.0);
}

__global__ void __launch_bounds__(64) k_chunks(double *out, size_t nrows, int len, int N, int mode, size_t mult)
{
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const size_t nchunks = nrows / N;
    if (mode == 2) {
        for (size_t t = blockIdx.x; t < nchunks; t += gridDim.x) {
            const size_t cc = (t * mult) % nchunks;
            for (int r = wib; r < N; r += 2) write_row(out + (cc * N + r) * (size_t)len, len, lane);
        }
        return;
    }
    const size_t nwarps = (size_t)gridDim.x * 2, w = blockIdx.x * (size_t)2 + wib;
    for (size_t t = w; t < nchunks; t += nwarps) {
        const size_t cc = mode == 3 ? t : (t * mult) % nchunks;
        if (mode == 1) {
            for (int r = 0; r < N; r += 2) write_row(out + (cc * N + r) * (size_t)len, len, lane);
            for (int r = 1; r < N; r += 2) write_row(out + (cc * N + r) * (size_t)len, len, lane);
        } else {
            for (int r = 0; r < N; r++) write_row(out + (cc * N + r) * (size_t)len, len, lane);
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
    printf("{\"gpu\": \"%s\", \"results\": [\n", prop.name);
    bool first = true;
    for (int len : {243, 256, 171})
        for (int mode : {0, 1, 2, 3})
            for (int N : {1, 2, 4, 10, 40, 160}) {
                if (mode != 0 && N == 1) continue;
                const size_t nrows = bytes / (len * 8);
                const float t = time_ms([&] { k_chunks<<<sms * 8, 64>>>(out, nrows, len, N, mode, 7919); });
                const double gb = (double)(nrows / N * N) * len * 8 / 1e9;
                printf("%s {\"row_bytes\": %d, \"mode\": %d, \"chunk_rows\": %d, \"GBs\": %.0f}", first ? "" : ",\n", len * 8, mode, N, gb / (t * 1e-3));
                first = false;
            }
    printf("\n]}\n");
    return 0;
}
