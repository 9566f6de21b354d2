// tools/microbench2.cu -- shared-memory atomic / load throughput on B200 with NOTHING else in the loop.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/microbench2 tools/microbench2.cu && ./tools/microbench2
// (tools/microbench.cu computed its random addresses with a runtime modulo inside the timed loop, which put an
// ALU floor of ~10 cycles under every number.)  Here every thread precomputes 8 addresses, the loop only issues
// the memory instruction.  Output: SM-cycles per warp-instruction at full occupancy (1 block x 1024 threads/SM).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
constexpr int THREADS = 1024, ITERS = 512, WORDS = 32 * 1024;

// pattern: 0 = conflict-free (lane -> bank lane), 1 = random words, 2 = all lanes one word,
//          3 = 8 tables x 4 lanes: lanes of a group share the word, groups in random banks (the tile kernel's hist)
//          4 = 8 tables x 4 lanes, all different random words, 5 = 2 distinct words per warp, 6 = 4 distinct words
__device__ uint32_t make_addr(int pattern, int lane, int k, uint32_t seed) {
    uint32_t s = seed * 2654435761u + k * 40503u + 12345u;
    s ^= s >> 15; s *= 2246822519u; s ^= s >> 13;
    switch (pattern) {
        case 0: return (uint32_t)(lane + 32 * ((s >> 8) % (WORDS / 32)));
        case 1: return (s >> 8) % WORDS;
        case 2: { uint32_t w = (seed / 32) * 2654435761u + k * 40503u; w ^= w >> 15; w *= 2246822519u; w ^= w >> 13; return (w >> 8) % WORDS; }
        case 3: { uint32_t g = (seed / 32) * 8 + (lane & 7); uint32_t w = g * 2654435761u + k * 40503u; w ^= w >> 15; w *= 2246822519u; w ^= w >> 13; return (lane & 7) * (WORDS / 8) + (w >> 8) % (WORDS / 8); }
        case 4: return (lane & 7) * (WORDS / 8) + (s >> 8) % (WORDS / 8);
        case 5: { uint32_t g = (seed / 32) * 2 + (lane & 1); uint32_t w = g * 2654435761u + k * 40503u; w ^= w >> 15; w *= 2246822519u; w ^= w >> 13; return (w >> 8) % WORDS; }
        default: { uint32_t g = (seed / 32) * 4 + (lane & 3); uint32_t w = g * 2654435761u + k * 40503u; w ^= w >> 15; w *= 2246822519u; w ^= w >> 13; return (w >> 8) % WORDS; }
    }
}

template <int OP>  // 0 = red.shared.add 1 (SASS ATOMS.POPC.INC), 1 = ld.shared.u32, 2 = ld.shared.u16, 3 = atom.shared.add (returns),
                   // 4 = red.shared.add of a register value (1 or 0x10000: two 16-bit counters per word)
__global__ void __launch_bounds__(THREADS, 1) k_probe(uint32_t* out, int pattern, int active) {
    extern __shared__ uint32_t sh[];
    for (int i = threadIdx.x; i < WORDS; i += blockDim.x) sh[i] = i & 0xffff;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    uint32_t a[8];
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(sh);
    for (int k = 0; k < 8; k++) a[k] = base + 4u * make_addr(pattern, lane, k, blockIdx.x * THREADS + threadIdx.x);
    uint32_t acc = 0;
    const uint32_t val = (threadIdx.x * 2654435761u >> 13 & 1u) ? 0x10000u : 1u;
    if (lane < active) {
        for (int i = 0; i < ITERS; i++) {
#pragma unroll
            for (int k = 0; k < 8; k++) {
                if (OP == 0) asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a[k]), "r"(1u) : "memory");
                // the loaded value feeds the next address (table values are < 2^31, so the address never changes,
                // but the compiler cannot know): 8 independent chains per thread x 32 warps keep the pipe full
                if (OP == 1) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a[k]) : "memory"); a[k] ^= (v >> 31) << 2; }
                if (OP == 2) { uint16_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a[k]) : "memory"); a[k] ^= ((uint32_t)v >> 16) << 2; }
                if (OP == 4) asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a[k]), "r"(val) : "memory");
                if (OP == 3) { uint32_t v; asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(v) : "r"(a[k]), "r"(1u) : "memory"); acc += v; }
            }
        }
    }
    __syncthreads();
    for (int k = 0; k < 8; k++) acc += a[k];
    for (int i = threadIdx.x; i < WORDS; i += blockDim.x) acc += sh[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <typename F>
static float time_it(F launch) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    launch();
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int r = 0; r < 3; r++) launch();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms / 3;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount, grid = sms;
    const double ghz = prop.clockRate * 1e-6;
    printf("%s  %d SMs  max SM clock %.3f GHz; SM-cycles per warp-instruction, 32 warps/SM\n", prop.name, sms, ghz);
    uint32_t* out; CK(cudaMalloc(&out, (size_t)grid * THREADS * 4));
    CK(cudaFuncSetAttribute(k_probe<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, WORDS * 4));
    CK(cudaFuncSetAttribute(k_probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, WORDS * 4));
    CK(cudaFuncSetAttribute(k_probe<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, WORDS * 4));
    CK(cudaFuncSetAttribute(k_probe<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, WORDS * 4));
    CK(cudaFuncSetAttribute(k_probe<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, WORDS * 4));
    const double warp_instr = (double)grid * (THREADS / 32) * ITERS * 8;
    auto cyc = [&](float ms) { return ms * 1e-3 * ghz * 1e9 * sms / warp_instr; };
    const char* pname[7] = {"conflict-free", "random words", "one word/warp", "8 tables x 4 same", "8 tables x 4 rand", "2 words/warp", "4 words/warp"};
    const char* oname[5] = {"red.inc", "lds.u32", "lds.u16", "atom.add", "red.add"};
    for (int op = 0; op < 5; op++)
        for (int pat = 0; pat < 7; pat++) {
            float ms = 0;
            if (op == 0) ms = time_it([&] { k_probe<0><<<grid, THREADS, WORDS * 4>>>(out, pat, 32); });
            if (op == 1) ms = time_it([&] { k_probe<1><<<grid, THREADS, WORDS * 4>>>(out, pat, 32); });
            if (op == 2) ms = time_it([&] { k_probe<2><<<grid, THREADS, WORDS * 4>>>(out, pat, 32); });
            if (op == 3) ms = time_it([&] { k_probe<3><<<grid, THREADS, WORDS * 4>>>(out, pat, 32); });
            if (op == 4) ms = time_it([&] { k_probe<4><<<grid, THREADS, WORDS * 4>>>(out, pat, 32); });
            printf("%-9s %-18s : %6.2f\n", oname[op], pname[pat], cyc(ms));
        }
    for (int act : {1, 4, 8, 16}) {
        float ms = time_it([&] { k_probe<0><<<grid, THREADS, WORDS * 4>>>(out, 1, act); });
        printf("red.add   random, %2d active lanes : %6.2f\n", act, cyc(ms));
    }
    CK(cudaDeviceSynchronize());
    printf("done\n");
    return 0;
}
