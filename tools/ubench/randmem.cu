// randmem.cu -- what does B200 HBM3e give for random 32/64/128-byte accesses over a working set >> L2?
// (sizing input for the hash-table layout: tools/ubench is measurement scaffolding, not product code)
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>
typedef unsigned long long u64;
typedef unsigned int u32;
__device__ __forceinline__ u64 mix(u64 x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x; }

// MODE 0: 32B read (v4.u64), 1: 32B read + atomicAdd u32 in the same sector, 2: atomicAdd only (RED), 3: 8B read,
// 4: 64B read (two 32B loads), 5: 32B read + atomicMax u64 other array (SoA-like: second sector), 6: prefetch.L2 then read later
template <int MODE>
__global__ void __launch_bounds__(256) k_rand(uint8_t* base, u64 n_sectors, u64 per_thread, u64 seed, u64* sink, uint8_t* base2)
{
    const u64 tid = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 acc = 0;
    u64 r = mix(seed + tid);
    for (u64 i = 0; i < per_thread; i += 4) {
        u64 idx[4];
#pragma unroll
        for (int k = 0; k < 4; k++) { r = r * 6364136223846793005ULL + 1442695040888963407ULL; idx[k] = (mix(r) % n_sectors); }
        if (MODE == 0 || MODE == 1 || MODE == 5) {
            u64 v[4][4];
#pragma unroll
            for (int k = 0; k < 4; k++)
                asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(v[k][0]), "=l"(v[k][1]), "=l"(v[k][2]), "=l"(v[k][3]) : "l"(base + idx[k] * 32));
#pragma unroll
            for (int k = 0; k < 4; k++) {
                acc += v[k][0] ^ v[k][3];
                if (MODE == 1) atomicAdd((u32*)(base + idx[k] * 32 + 24), 1u);
                if (MODE == 5) atomicMax((u64*)(base2 + idx[k] * 8), r);
            }
        } else if (MODE == 7 || MODE == 8 || MODE == 9) {
            u64 pol;
            if (MODE == 7) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
            else if (MODE == 8) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
            else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
            u64 v[4][4];
#pragma unroll
            for (int k = 0; k < 4; k++)
                asm volatile("ld.global.cg.L2::cache_hint.v4.u64 {%0,%1,%2,%3}, [%4], %5;" : "=l"(v[k][0]), "=l"(v[k][1]), "=l"(v[k][2]), "=l"(v[k][3]) : "l"(base + idx[k] * 32), "l"(pol));
#pragma unroll
            for (int k = 0; k < 4; k++) acc += v[k][0] ^ v[k][3];
        } else if (MODE == 2) {
#pragma unroll
            for (int k = 0; k < 4; k++) atomicAdd((u32*)(base + idx[k] * 32 + 24), 1u);
        } else if (MODE == 3) {
            u64 v[4];
#pragma unroll
            for (int k = 0; k < 4; k++) asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v[k]) : "l"(base + idx[k] * 32));
#pragma unroll
            for (int k = 0; k < 4; k++) acc += v[k];
        } else if (MODE == 4) {
            u64 v[4][8];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint8_t* p = base + (idx[k] & ~1ull) * 32;
                asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(v[k][0]), "=l"(v[k][1]), "=l"(v[k][2]), "=l"(v[k][3]) : "l"(p));
                asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(v[k][4]), "=l"(v[k][5]), "=l"(v[k][6]), "=l"(v[k][7]) : "l"(p + 32));
            }
#pragma unroll
            for (int k = 0; k < 4; k++) acc += v[k][0] ^ v[k][7];
        }
    }
    if (acc == 0x1234567ull) *sink = acc;
}

template <int MODE>
static void run(const char* name, uint8_t* buf, u64 bytes, uint8_t* buf2, int ctas_per_sm)
{
    const u64 n_sectors = bytes / 32;
    const int grid = 148 * ctas_per_sm;
    const u64 per_thread = 2048;
    u64* sink; cudaMalloc(&sink, 8);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k_rand<MODE><<<grid, 256>>>(buf, n_sectors, 64, 1, sink, buf2);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    k_rand<MODE><<<grid, 256>>>(buf, n_sectors, per_thread, 7, sink, buf2);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double ops = (double)grid * 256 * per_thread;
    printf("%-28s ws=%6.0f MB ctas/sm=%d  %8.2f G acc/s  (%.3f ms, err=%s)\n", name, bytes / 1048576.0, ctas_per_sm, ops / ms * 1e-6, ms, cudaGetErrorString(cudaGetLastError()));
    cudaFree(sink);
}

int main(int argc, char** argv)
{
    if (argc > 1) {
        size_t g = 0;
        cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(argv[1]));
        cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity);
        printf("cudaLimitMaxL2FetchGranularity := %s -> %s, now %zu\n", argv[1], cudaGetErrorString(e), g);
    } else {
        size_t g = 0;
        cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity);
        printf("cudaLimitMaxL2FetchGranularity default %zu\n", g);
    }
    const u64 big = 4ull << 30;
    uint8_t *buf, *buf2;
    cudaMalloc(&buf, big); cudaMalloc(&buf2, big / 4);
    cudaMemset(buf, 0, big); cudaMemset(buf2, 0, big / 4);
    for (u64 ws : {256ull << 20, 1ull << 30}) {
        run<0>("read32 plain", buf, ws, buf2, 8);
        run<7>("read32 hint evict_last", buf, ws, buf2, 8);
        run<8>("read32 hint evict_first", buf, ws, buf2, 8);
        run<9>("read32 hint evict_normal", buf, ws, buf2, 8);
    }
    {   // streaming copy of 1 GiB (read + write bytes)
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        cudaMemcpy(buf + (2ull << 30), buf, 1ull << 30, cudaMemcpyDeviceToDevice);
        cudaEventRecord(a);
        for (int i = 0; i < 5; i++) cudaMemcpyAsync(buf + (2ull << 30), buf, 1ull << 30, cudaMemcpyDeviceToDevice);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        printf("D2D copy 1 GiB: %.1f GB/s (read+write)\n", 5 * 2.0 * (1ull << 30) / ms * 1e-6);
    }
    for (u64 ws : {48ull << 20, 256ull << 20, 1ull << 30, 4ull << 30}) {
        if (getenv("RANDMEM_SHORT")) break;
        for (int c : {4, 8}) {
            run<0>("read32", buf, ws, buf2, c);
            run<3>("read8", buf, ws, buf2, c);
            run<4>("read64", buf, ws, buf2, c);
            run<2>("red.add32", buf, ws, buf2, c);
            run<1>("read32+red.add same sector", buf, ws, buf2, c);
            run<5>("read32+atomMax 2nd array", buf, ws, buf2, c);
        }
    }
    return 0;
}
