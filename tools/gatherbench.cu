// tools/gatherbench.cu -- random 16/32/64-byte gathers from a 1 GiB array (the cell-major label matrix access pattern).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/gatherbench tools/gatherbench.cu && ./tools/gatherbench
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

// each thread: `per` independent random rows of `row_bytes`, reads `vecs` consecutive uint4 from each
template <int VECS>
__global__ void __launch_bounds__(1024, 1) k_gather(const uint4* __restrict__ src, size_t nrows, size_t row_vec, uint32_t* out, int iters) {
    uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u, acc = 0;
    for (int i = 0; i < iters; i++) {
        uint4 v[4][VECS];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            s = s * 1664525u + 1013904223u;
            const size_t row = ((size_t)(s >> 4) * 2654435761ull) % nrows;
#pragma unroll
            for (int k = 0; k < VECS; k++) v[u][k] = __ldg(src + row * row_vec + k);
        }
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int k = 0; k < VECS; k++) acc += v[u][k].x ^ v[u][k].w;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    const size_t bytes = 1ull << 30;
    uint4* big; CK(cudaMalloc(&big, bytes)); CK(cudaMemset(big, 1, bytes));
    uint32_t* out; CK(cudaMalloc(&out, (size_t)sms * 1024 * 4));
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int iters = 64;
    for (int row_bytes : {16, 32, 64, 1024}) {
        for (int vecs : {1, 2, 4}) {
            if (vecs * 16 > row_bytes) continue;
            const size_t nrows = bytes / row_bytes, row_vec = row_bytes / 16;
            for (int rep = 0; rep < 2; rep++) {
                cudaEventRecord(a);
                if (vecs == 1) k_gather<1><<<sms, 1024>>>(big, nrows, row_vec, out, iters);
                if (vecs == 2) k_gather<2><<<sms, 1024>>>(big, nrows, row_vec, out, iters);
                if (vecs == 4) k_gather<4><<<sms, 1024>>>(big, nrows, row_vec, out, iters);
                cudaEventRecord(b); cudaEventSynchronize(b);
            }
            float ms; cudaEventElapsedTime(&ms, a, b);
            const double rows = (double)sms * 1024 * iters * 4;
            printf("row stride %4d B, read %2d B per row: %7.2f G rows/s, %8.1f GB/s useful\n", row_bytes, vecs * 16, rows / ms * 1e-6, rows * vecs * 16 / ms * 1e-6);
        }
    }
    CK(cudaDeviceSynchronize());
    return 0;
}
