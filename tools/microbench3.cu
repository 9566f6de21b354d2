// tools/microbench3.cu -- shared-memory red.inc (ATOMS.POPC.INC) when ALL warps of a block hit the SAME few words: the histogram
// of a single-row (streaming) item whose cells mostly share one cluster in the partner.  tools/microbench2.cu only measured
// patterns where every warp has words of its own.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/microbench3 tools/microbench3.cu && ./tools/microbench3
// Output: SM-cycles per warp instruction, one block of `warps` warps per SM, nothing else in the loop.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
constexpr int ITERS = 512, WORDS = 8 * 1024;

// pattern 0: every warp its own word (per slot k)            1: all warps, all lanes -> word 64 k (8 hot words, one per slot)
//         2: like 1, 28 lanes hot + 4 lanes random in a 50-word row   3: like 1 but warp w starts at slot (w + k) % 8
//         4: like 1, row replicated 4 x (copy = warp % 4)    5: like 2, replicated 4 x     6: like 1, replicated per warp
__device__ uint32_t make_addr(int pattern, int warp, int lane, int k, uint32_t seed) {
    uint32_t s = seed * 2654435761u + k * 40503u + 12345u;
    s ^= s >> 15; s *= 2246822519u; s ^= s >> 13;
    const uint32_t noisy = 64u * k + 1u + (s >> 8) % 49u;
    switch (pattern) {
        case 0: return (uint32_t)(warp * 64 + k) % WORDS;
        case 1: return 64u * k;
        case 2: return lane < 28 ? 64u * k : noisy;
        case 3: return 64u * ((warp + k) & 7);
        case 4: return 512u * (warp & 3) + 64u * k;
        case 5: return 512u * (warp & 3) + (lane < 28 ? 64u * k : noisy);
        default: return (512u * warp + 64u * k) % WORDS;
    }
}

__global__ void __launch_bounds__(1024, 1) k_probe(uint32_t* out, int pattern) {
    extern __shared__ uint32_t sh[];
    for (int i = threadIdx.x; i < WORDS; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t a[8];
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(sh);
    for (int k = 0; k < 8; k++) a[k] = base + 4u * make_addr(pattern, warp, lane, k, blockIdx.x * 1024 + threadIdx.x);
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(a[k]) : "memory");
    }
    __syncthreads();
    uint32_t acc = 0;
    for (int i = threadIdx.x; i < WORDS; i += blockDim.x) acc += sh[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    const double ghz = prop.clockRate * 1e-6;
    printf("%s  %d SMs  %.3f GHz; red.shared.add.u32 [a], 1 -- SM-cycles per warp instruction\n", prop.name, sms, ghz);
    uint32_t* out; CK(cudaMalloc(&out, (size_t)sms * 1024 * 4));
    CK(cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, WORDS * 4));
    const char* pname[7] = {"own word per warp", "ALL warps -> 8 hot words", "28 lanes hot + 4 in a 50-word row", "8 hot words, warps rotated",
                            "hot words x 4 copies (warp % 4)", "28 hot + 4 noisy, x 4 copies", "hot words, one copy per warp"};
    for (int warps : {8, 19, 27, 32})
        for (int pat = 0; pat < 7; pat++) {
            cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
            k_probe<<<sms, warps * 32, WORDS * 4>>>(out, pat);
            CK(cudaDeviceSynchronize());
            cudaEventRecord(a);
            for (int r = 0; r < 3; r++) k_probe<<<sms, warps * 32, WORDS * 4>>>(out, pat);
            cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            const double wi = (double)warps * ITERS * 8;  // warp instructions per SM
            printf("%2d warps  %-36s : %6.2f\n", warps, pname[pat], ms / 3 * 1e-3 * ghz * 1e9 / wi);
        }
    CK(cudaDeviceSynchronize());
    return 0;
}
