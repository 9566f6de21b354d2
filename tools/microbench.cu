// tools/microbench.cu -- B200 primitive throughput probes that drive the row-block kernel design.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/microbench tools/microbench.cu && ./tools/microbench
// Every probe runs a full grid (148 SMs x 2 blocks x 512 threads, like the row-block kernel) and reports
// SM-cycles per warp-instruction of the probed primitive (lower = faster), derived from elapsed time x SM clock.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int THREADS = 512, ITERS = 2048;

__global__ void k_match(uint32_t* out, int distinct) {
    const int lane = threadIdx.x & 31;
    uint32_t key = (uint32_t)(lane % distinct) * 977u + blockIdx.x, acc = 0;
    for (int i = 0; i < ITERS; i++) {
        unsigned g = __match_any_sync(0xffffffffu, key);
        acc += g;
        key += (g & 1u);  // dependency so the loop cannot be hoisted
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

__global__ void k_redux(uint32_t* out) {
    uint32_t v = threadIdx.x, acc = 0;
    for (int i = 0; i < ITERS; i++) {
        acc += __reduce_add_sync(0xffffffffu, v);
        v += acc & 1u;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

__global__ void k_ballot_shfl(uint32_t* out) {
    uint32_t v = threadIdx.x, acc = 0;
    for (int i = 0; i < ITERS; i++) {
        unsigned b = __ballot_sync(0xffffffffu, (v & 3u) != 0);
        acc += __shfl_sync(0xffffffffu, v, __ffs(b | 1u) - 1);
        v += acc & 1u;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// shared atomics: mode 0 = all lanes distinct random words, 1 = all lanes same word, 2 = `active` lanes distinct,
// 3 = 16-bit packed style (two lanes per word, different halves)
__global__ void k_atoms(uint32_t* out, int mode, int active, int words) {
    extern __shared__ uint32_t sh[];
    for (int i = threadIdx.x; i < words; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    uint32_t s = threadIdx.x * 2654435761u + blockIdx.x;
    for (int i = 0; i < ITERS; i++) {
        s = s * 1664525u + 1013904223u;
        uint32_t w = (s >> 8) % (uint32_t)words;
        if (mode == 1) w = (uint32_t)((i * 7) % words);
        if (mode == 3) w = ((s >> 8) % (uint32_t)words) & ~1u;
        uint32_t v = (mode == 3 && (lane & 1)) ? 0x10000u : 1u;
        if (mode == 3) w |= 0;  // pairs of lanes may share a word by chance only
        if (lane < active) atomicAdd(sh + w, v);
    }
    __syncthreads();
    uint32_t acc = 0;
    for (int i = threadIdx.x; i < words; i += blockDim.x) acc += sh[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// random shared loads (gather from a table), 4-byte and 8-byte
__global__ void k_lds(uint32_t* out, int words, int wide) {
    extern __shared__ uint32_t sh[];
    for (int i = threadIdx.x; i < words; i += blockDim.x) sh[i] = i * 2654435761u;
    __syncthreads();
    uint32_t s = threadIdx.x * 2654435761u + blockIdx.x, acc = 0;
    for (int i = 0; i < ITERS; i++) {
        s = s * 1664525u + 1013904223u;
        uint32_t w = (s >> 8) % (uint32_t)words;
        if (wide) {
            uint2 v = reinterpret_cast<const uint2*>(sh)[w >> 1];
            acc += v.x ^ v.y;
        } else {
            acc += sh[w];
        }
        s += acc & 1u;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// fp64 reciprocal (rcp.approx + 2 Newton) + convert + multiply: the ecs_of() sequence
__global__ void k_ecs(double* out) {
    double acc = 0.0;
    uint32_t a = threadIdx.x + 3, b = blockIdx.x + 5;
    for (int i = 0; i < ITERS; i++) {
        const double mx = (double)max(a, b);
        double r;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(mx));
        r = fma(fma(-mx, r, 1.0), r, r);
        r = fma(fma(-mx, r, 1.0), r, r);
        acc += (double)(a & 0xffffu) * r;
        a += 7; b += (uint32_t)acc & 1u;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
__global__ void k_div(double* out) {
    double acc = 0.0;
    uint32_t a = threadIdx.x + 3, b = blockIdx.x + 5;
    for (int i = 0; i < ITERS; i++) {
        acc += (double)(a & 0xffffu) / (double)max(a, b);
        a += 7; b += (uint32_t)acc & 1u;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// random 16-byte / 32-byte gathers from a big array (the cell-major label matrix access pattern)
__global__ void k_gather(const uint4* __restrict__ src, size_t nvec, uint32_t* out, int two) {
    uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u, acc = 0;
    for (int i = 0; i < 64; i++) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            s = s * 1664525u + 1013904223u;
            size_t idx = ((size_t)s * 2654435761ull) % nvec;
            if (two) idx &= ~(size_t)1;
            v[u] = __ldg(src + idx);
            if (two) { uint4 w = __ldg(src + idx + 1); v[u].x ^= w.x; }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) acc += v[u].x ^ v[u].w;
    }
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
    const int sms = prop.multiProcessorCount, grid = sms * 2;
    const double ghz = prop.clockRate * 1e-6;  // max SM clock
    printf("%s  %d SMs  max SM clock %.3f GHz (cycles below assume this clock)\n", prop.name, sms, ghz);
    uint32_t* out; CK(cudaMalloc(&out, (size_t)grid * THREADS * 8));
    const double warp_instr = (double)grid * (THREADS / 32) * ITERS;
    auto cyc = [&](float ms) { return ms * 1e-3 * ghz * 1e9 * sms / warp_instr; };  // SM-cycles per warp-instruction
    for (int d : {1, 2, 4, 8, 16, 32}) {
        float ms = time_it([&] { k_match<<<grid, THREADS>>>(out, d); });
        printf("match_any  %2d distinct keys/warp : %7.2f SM-cycles per warp-instr  (%.3f per lane)\n", d, cyc(ms), cyc(ms) / 32);
    }
    { float ms = time_it([&] { k_redux<<<grid, THREADS>>>(out); }); printf("redux.sum                       : %7.2f\n", cyc(ms)); }
    { float ms = time_it([&] { k_ballot_shfl<<<grid, THREADS>>>(out); }); printf("ballot+ffs+shfl                 : %7.2f\n", cyc(ms)); }
    const int words = 24 * 1024;  // 96 KB table
    CK(cudaFuncSetAttribute(k_atoms, cudaFuncAttributeMaxDynamicSharedMemorySize, words * 4));
    CK(cudaFuncSetAttribute(k_lds, cudaFuncAttributeMaxDynamicSharedMemorySize, words * 4));
    { float ms = time_it([&] { k_atoms<<<grid, THREADS, words * 4>>>(out, 0, 32, words); }); printf("atoms 32 lanes random words     : %7.2f  (%.3f per lane)\n", cyc(ms), cyc(ms) / 32); }
    { float ms = time_it([&] { k_atoms<<<grid, THREADS, words * 4>>>(out, 1, 32, words); }); printf("atoms 32 lanes same word        : %7.2f  (%.3f per lane)\n", cyc(ms), cyc(ms) / 32); }
    for (int act : {1, 2, 4, 8, 16}) {
        float ms = time_it([&] { k_atoms<<<grid, THREADS, words * 4>>>(out, 2, act, words); });
        printf("atoms %2d lanes random words     : %7.2f  (%.3f per active lane)\n", act, cyc(ms), cyc(ms) / act);
    }
    { float ms = time_it([&] { k_atoms<<<grid, THREADS, words * 4>>>(out, 3, 32, words); }); printf("atoms 32 lanes packed halves    : %7.2f\n", cyc(ms)); }
    { float ms = time_it([&] { k_lds<<<grid, THREADS, words * 4>>>(out, words, 0); }); printf("lds.32 random                   : %7.2f  (%.3f per lane)\n", cyc(ms), cyc(ms) / 32); }
    { float ms = time_it([&] { k_lds<<<grid, THREADS, words * 4>>>(out, words, 1); }); printf("lds.64 random                   : %7.2f  (%.3f per lane)\n", cyc(ms), cyc(ms) / 32); }
    { float ms = time_it([&] { k_ecs<<<grid, THREADS>>>((double*)out); }); printf("ecs_of (rcp+2 newton+cvt+mul)   : %7.2f\n", cyc(ms)); }
    { float ms = time_it([&] { k_div<<<grid, THREADS>>>((double*)out); }); printf("fp64 division + cvt             : %7.2f\n", cyc(ms)); }
    // random sector gathers from 1 GiB
    const size_t bytes = 1ull << 30, nvec = bytes / 16;
    uint4* big; CK(cudaMalloc(&big, bytes)); CK(cudaMemset(big, 1, bytes));
    for (int two : {0, 1}) {
        const int g = sms * 16;
        float ms = time_it([&] { k_gather<<<g, 256>>>(big, nvec, out, two); });
        double loads = (double)g * 256 * 64 * 4;
        printf("random %2d-byte gathers from 1 GiB: %.1f G loads/s, %.1f GB/s useful\n", two ? 32 : 16, loads / ms * 1e-6, loads * (two ? 32 : 16) / ms * 1e-6);
    }
    printf("done\n");
    return 0;
}
