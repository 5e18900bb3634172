// Write micro-benchmark 3: partial sectors at row boundaries.
//  (a) rows that are 32-byte but not 128-byte aligned (1952 B) in random order: is the penalty of unaligned rows per 32-byte
//      SECTOR or per 128-byte LINE?
//  (b) unaligned rows (1944 B, 8-byte aligned) in random order, written as [full sectors] + [head / tail fragments as whole
//      32-byte slots of a side buffer] + a stitch pass that composes every boundary sector from its two fragments.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/microbench_frag tools/microbench_frag.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(64) k_rows16(double *out, size_t nrows, int len, size_t mult)
{
    const int lane = threadIdx.x & 31;
    const size_t nwarps = (size_t)gridDim.x * 2, w = blockIdx.x * (size_t)2 + (threadIdx.x >> 5);
    for (size_t t = w; t < nrows; t += nwarps) {
        const size_t cc = (t * mult) % nrows;
        double *p = out + cc * (size_t)len;
        const int h = (int)((reinterpret_cast<uintptr_t>(p) >> 3) & 1);
        const int body = (len - h) & ~1;
        if (lane == 0 && h) p[0] = 0.0;
        if (lane == 1 && h + body < len) p[len - 1] = 0.0;
        double2 *q = reinterpret_cast<double2 *>(p + h);
        for (int x = lane; x < body / 2; x += 32) q[x] = make_double2(1.0, 2.0);
    }
}

// full sectors to the row, the partial head / tail sectors as whole 32-byte slots into frag[2 * row], frag[2 * row + 1]
__global__ void __launch_bounds__(64) k_rows_frag(double *out, double4 *frag, size_t nrows, int len, size_t mult)
{
    const int lane = threadIdx.x & 31;
    const size_t nwarps = (size_t)gridDim.x * 2, w = blockIdx.x * (size_t)2 + (threadIdx.x >> 5);
    for (size_t t = w; t < nrows; t += nwarps) {
        const size_t cc = (t * mult) % nrows;
        const size_t s = cc * (size_t)len, e = s + len;           // in doubles (out is 128-byte aligned)
        const size_t s4 = (s + 3) & ~(size_t)3, e4 = e & ~(size_t)3;
        double2 *q = reinterpret_cast<double2 *>(out + s4);
        const int n2 = (int)((e4 - s4) >> 1);
        for (int x = lane; x < n2; x += 32) q[x] = make_double2(1.0, 2.0);
        if (lane == 0 && s4 != s) frag[2 * cc] = make_double4(1.0, 2.0, 3.0, 4.0);
        if (lane == 1 && e4 != e) frag[2 * cc + 1] = make_double4(1.0, 2.0, 3.0, 4.0);
    }
}

// boundary between row c and row c + 1: sector at doubles [e4, e4 + 4) = tail of row c | head of row c + 1
__global__ void __launch_bounds__(256) k_stitch(double *out, const double4 *frag, size_t nrows, int len)
{
    for (size_t c = blockIdx.x * (size_t)blockDim.x + threadIdx.x; c + 1 < nrows; c += (size_t)gridDim.x * blockDim.x) {
        const size_t e = (c + 1) * (size_t)len, e4 = e & ~(size_t)3;
        const int t = (int)(e - e4);
        if (t == 0) continue;
        const double4 a = frag[2 * c + 1], b = frag[2 * (c + 1)];
        double4 o;
        o.x = a.x;
        o.y = t > 1 ? a.y : b.y;
        o.z = t > 2 ? a.z : b.z;
        o.w = b.w;
        *reinterpret_cast<double4 *>(out + e4) = o;
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
    double4 *frag;
    if (cudaMalloc(&out, bytes + 4096) != cudaSuccess) { printf("{\"error\": \"alloc\"}\n"); return 1; }
    printf("{\"gpu\": \"%s\"", prop.name);
    for (int len : {243, 244, 248, 256, 171, 172}) {
        const size_t nrows = bytes / (len * 8);
        const float t = time_ms([&] { k_rows16<<<sms * 8, 64>>>(out, nrows, len, 7919); });
        printf(", \"random_rows_%dB_GBs\": %.0f", len * 8, (double)nrows * len * 8 / 1e9 / (t * 1e-3));
    }
    for (int len : {243, 171, 585}) {
        const size_t nrows = bytes / (len * 8);
        cudaMalloc(&frag, nrows * 64);
        const float t1 = time_ms([&] { k_rows_frag<<<sms * 8, 64>>>(out, frag, nrows, len, 7919); });
        const float t2 = time_ms([&] { k_stitch<<<sms * 8, 256>>>(out, frag, nrows, len); });
        const float t0 = time_ms([&] { k_rows16<<<sms * 8, 64>>>(out, nrows, len, 7919); });
        const double gb = (double)nrows * len * 8 / 1e9;
        printf(", \"rows_%dB\": {\"plain_GBs\": %.0f, \"frag_rows_ms\": %.3f, \"stitch_ms\": %.3f, \"plain_ms\": %.3f, \"frag_total_GBs\": %.0f}", len * 8,
               gb / (t0 * 1e-3), t1, t2, t0, gb / ((t1 + t2) * 1e-3));
        cudaFree(frag);
    }
    printf("}\n");
    return 0;
}
