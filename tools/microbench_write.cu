// Write-path micro-benchmarks on the bench GPU: what a kernel that (almost) only WRITES can reach, for the access patterns of
// the assembly (CSR rows of ~2 KB written once, in an order that is not the address order).  One JSON line on stdout.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/microbench_write tools/microbench_write.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) k_fill(double *out, size_t n4)   // n4 = number of 32-byte groups
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        double *p = out + 4 * i;
        asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(1.0), "d"(2.0), "d"(3.0), "d"(4.0) : "memory");
    }
}

__global__ void __launch_bounds__(256) k_fill8(double *out, size_t n)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = 1.0;
}

// rows written with 16-byte stores where the row is 16-byte aligned inside (odd head / tail doubles by 8-byte stores)
__global__ void __launch_bounds__(64) k_chunks16(double *out, size_t nchunks, int len, size_t stride)
{
    const int lane = threadIdx.x & 31;
    const size_t nwarps = (size_t)gridDim.x * (blockDim.x >> 5), w = blockIdx.x * (size_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
    for (size_t c = w; c < nchunks; c += nwarps) {
        const size_t cc = (c * stride) % nchunks;
        double *p = out + cc * (size_t)len;
        const int h = (int)((reinterpret_cast<uintptr_t>(p) >> 3) & 1);
        const int body = (len - h) & ~1;
        if (lane == 0 && h) p[0] = 0.0;
        if (lane == 1 && h + body < len) p[len - 1] = 0.0;
        double2 *q = reinterpret_cast<double2 *>(p + h);
        for (int x = lane; x < body / 2; x += 32) q[x] = make_double2(1.0, 2.0);
    }
}

// a warp writes chunks of `len` doubles (8-byte aligned, like CSR rows); consecutive chunks of a warp are `stride` chunks apart
__global__ void __launch_bounds__(64) k_chunks(double *out, size_t nchunks, int len, size_t stride)
{
    const int lane = threadIdx.x & 31;
    const size_t nwarps = (size_t)gridDim.x * (blockDim.x >> 5), w = blockIdx.x * (size_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
    for (size_t c = w; c < nchunks; c += nwarps) {
        const size_t cc = (c * stride) % nchunks;
        double *p = out + cc * (size_t)len;
        for (int x = lane; x < len; x += 32) p[x] = (double)x;
    }
}

__global__ void __launch_bounds__(64) k_chunks_tma(double *out, size_t nchunks, int len, size_t stride)
{
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31;
    double *buf = sm + (size_t)(threadIdx.x >> 5) * ((len + 3) & ~1);
    const size_t nwarps = (size_t)gridDim.x * (blockDim.x >> 5), w = blockIdx.x * (size_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
    for (size_t c = w; c < nchunks; c += nwarps) {
        const size_t cc = (c * stride) % nchunks;
        double *p = out + cc * (size_t)len;
        const int h = (int)((reinterpret_cast<uintptr_t>(p) >> 3) & 1);
        for (int x = lane; x < len; x += 32) buf[x + h] = (double)x;
        __syncwarp();
        if (lane == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            const int body = (len - h) & ~1;
            if (h) p[0] = buf[h];
            const uint32_t s = (uint32_t)__cvta_generic_to_shared(buf + 2 * h);
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(p + h), "r"(s), "r"(body * 8) : "memory");
            if (h + body < len) p[len - 1] = buf[len - 1 + h];
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        __syncwarp();
    }
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
    const int sms = prop.multiProcessorCount;
    const size_t bytes = (size_t)4 << 30;
    double *out;
    if (cudaMalloc(&out, bytes + 4096) != cudaSuccess) { printf("{\"error\": \"alloc\"}\n"); return 1; }
    const float t_fill = time_ms([&] { k_fill<<<sms * 16, 256>>>(out, bytes / 32); });
    const int len = 243;                                   // 9 L doubles of an edge-node row (L = 27): 1944 bytes
    const size_t nchunks = bytes / (len * 8);
    const float t_seq = time_ms([&] { k_chunks<<<sms * 8, 64>>>(out, nchunks, len, 1); });
    const float t_str = time_ms([&] { k_chunks<<<sms * 8, 64>>>(out, nchunks, len, 7919); });
    const float t_fill8 = time_ms([&] { k_fill8<<<sms * 16, 256>>>(out, bytes / 8); });
    const float t_seq16 = time_ms([&] { k_chunks16<<<sms * 8, 64>>>(out, nchunks, len, 1); });
    const float t_str16 = time_ms([&] { k_chunks16<<<sms * 8, 64>>>(out, nchunks, len, 7919); });
    const float t_seq_w16 = time_ms([&] { k_chunks<<<sms * 16, 64>>>(out, nchunks, len, 1); });     // 32 warps per SM
    const int len2 = 585;                                  // vertex-node row (L = 65): 4680 bytes
    const size_t nchunks2 = bytes / (len2 * 8);
    const float t_seq2 = time_ms([&] { k_chunks<<<sms * 8, 64>>>(out, nchunks2, len2, 1); });
    const float t_str2 = time_ms([&] { k_chunks<<<sms * 8, 64>>>(out, nchunks2, len2, 7919); });
    cudaFuncSetAttribute(k_chunks_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * ((len + 3) & ~1) * 8);
    const float t_tma = time_ms([&] { k_chunks_tma<<<sms * 8, 64, 2 * ((len + 3) & ~1) * 8>>>(out, nchunks, len, 7919); });
    const float t_tma_seq = time_ms([&] { k_chunks_tma<<<sms * 8, 64, 2 * ((len + 3) & ~1) * 8>>>(out, nchunks, len, 1); });
    const double gb = bytes / 1e9, gbc = nchunks * (double)len * 8 / 1e9;
    printf("{\"fill_8B_GBs\": %.0f, \"rows_1944B_in_order_16Bstores_GBs\": %.0f, \"rows_1944B_strided_16Bstores_GBs\": %.0f, "
           "\"rows_1944B_in_order_32warps_GBs\": %.0f, \"rows_4680B_in_order_GBs\": %.0f, \"rows_4680B_strided_GBs\": %.0f}\n",
           gb / (t_fill8 * 1e-3), gbc / (t_seq16 * 1e-3), gbc / (t_str16 * 1e-3), gbc / (t_seq_w16 * 1e-3),
           nchunks2 * (double)len2 * 8 / 1e9 / (t_seq2 * 1e-3), nchunks2 * (double)len2 * 8 / 1e9 / (t_str2 * 1e-3));
    printf("{\"gpu\": \"%s\", \"fill_v4_GBs\": %.0f, \"rows_1944B_in_order_GBs\": %.0f, \"rows_1944B_strided_order_GBs\": %.0f, "
           "\"rows_1944B_strided_tma_bulk_GBs\": %.0f, \"rows_1944B_in_order_tma_bulk_GBs\": %.0f}\n",
           prop.name, gb / (t_fill * 1e-3), gbc / (t_seq * 1e-3), gbc / (t_str * 1e-3), gbc / (t_tma * 1e-3), gbc / (t_tma_seq * 1e-3));
    return 0;
}
