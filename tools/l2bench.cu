// tools/l2bench.cu -- how large a randomly gathered working set does the B200 L2 really hold?
// Random 16-byte reads (one per 32-byte sector row) inside working sets of 8 MB .. 1 GiB; optionally with a
// concurrent streaming read of a big buffer (what table fills / perm streams do to the cache).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__global__ void __launch_bounds__(1024, 1) k_gather(const uint4* __restrict__ src, size_t nrows, size_t row_vec, uint32_t* out, int iters,
                                                    const uint4* __restrict__ stream, size_t stream_vec, int stream_every) {
    uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u, acc = 0;
    size_t sp = (size_t)(blockIdx.x * blockDim.x + threadIdx.x);
    for (int i = 0; i < iters; i++) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            s = s * 1664525u + 1013904223u;
            const size_t row = ((size_t)(s >> 4) * 2654435761ull) % nrows;
            v[u] = __ldg(src + row * row_vec);
        }
        if (stream_every && (i % stream_every) == 0) {
            uint4 w = __ldg(stream + sp % stream_vec);
            sp += (size_t)gridDim.x * blockDim.x;
            acc += w.y;
        }
#pragma unroll
        for (int u = 0; u < 4; u++) acc += v[u].x ^ v[u].w;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    printf("%s L2 %d MB, persisting max %d MB\n", prop.name, prop.l2CacheSize >> 20, prop.persistingL2CacheMaxSize >> 20);
    const size_t bytes = 1ull << 30;
    uint4 *big, *strm; CK(cudaMalloc(&big, bytes)); CK(cudaMemset(big, 1, bytes));
    CK(cudaMalloc(&strm, bytes)); CK(cudaMemset(strm, 2, bytes));
    uint32_t* out; CK(cudaMalloc(&out, (size_t)sms * 1024 * 4));
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int iters = 256;
    for (int stream_every : {0, 4, 1}) {
        for (int ws_mb : {8, 16, 32, 48, 64, 96, 128, 256, 1024}) {
            const size_t nrows = ((size_t)ws_mb << 20) / 32;
            float ms = 0;
            for (int rep = 0; rep < 3; rep++) {
                cudaEventRecord(a);
                k_gather<<<sms, 1024>>>(big, nrows, 2, out, iters, strm, bytes / 16, stream_every);
                cudaEventRecord(b); cudaEventSynchronize(b);
                cudaEventElapsedTime(&ms, a, b);
            }
            const double rows = (double)sms * 1024 * iters * 4;
            printf("stream 1 x 16 B per %d gathers | working set %4d MB (32-byte rows): %7.2f G sector gathers/s\n", stream_every * 4, ws_mb, rows / ms * 1e-6);
        }
    }
    CK(cudaDeviceSynchronize());
    return 0;
}
