// cost of the forward16 cell core alone (registers only): cycles per column step per SMSP
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include "../multiplexnanopore_b200/csrc/smx_dp16.cuh"
template <int MODE>
__global__ void __launch_bounds__(128, 4) core(uint32_t *out, int iters, uint32_t seed, uint32_t one)
{
    constexpr int K = 8;
    uint32_t plo[K], phi[K], Hl[K], Hlgo[K], El[K];
    for (int k = 0; k < K; ++k) { plo[k] = seed * (k + 3); phi[k] = seed * (k + 11); Hl[k] = seed + k + threadIdx.x; Hlgo[k] = Hl[k] - 3; El[k] = seed ^ k; }
    uint32_t hd = seed, out_h = seed + 1, out_f = seed + 2, sel = 0x4c80 | (threadIdx.x & 3), acc = 0;
    const uint32_t nge2 = 0xffffffffu, ngo2 = 0xfffdfffdu;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        uint32_t hu = out_h, fu = out_f;
        if (MODE == 1) { hu = __shfl_up_sync(0xffffffffu, out_h, 1); fu = __shfl_up_sync(0xffffffffu, out_f, 1); sel = __shfl_up_sync(0xffffffffu, sel, 1) | 0x4c80; }
        Smx16Col c;
        c.hd = hd; c.fu = fu; c.hugo = smx16_add(hu, ngo2); c.h = 0; c.f = 0; c.w_lo = 0; c.w_hi = 0; c.w_lo2 = 0; c.w_hi2 = 0; c.one = one;
        Smx16Cell<K, 0>::run(Hl, Hlgo, El, plo, phi, c, sel, nge2, ngo2);
        hd = hu; out_h = c.h; out_f = c.f;
        acc ^= (c.w_lo | c.w_lo2) + (c.w_hi | c.w_hi2);
    }
    uint32_t s = acc ^ out_h ^ out_f;
    for (int k = 0; k < K; ++k) s ^= Hl[k] ^ El[k];
    if (s == 0x7fffffff) out[0] = s;
}
int main(int argc, char **argv) {
    uint32_t *d; cudaMalloc(&d, 64);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int mode = 0; mode < 2; ++mode)
    for (int wps = 1; wps <= 4; ++wps) {
        const int blocks = 148 * wps;  // 128 threads = 4 warps = one per SMSP per block
        float best = 1e30f;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) core<0><<<blocks, 128>>>(d, iters, 12345u, 1u); else core<1><<<blocks, 128>>>(d, iters, 12345u, 1u);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        const double cyc_total = best * 1e-3 * (double)clk * 1e3;
        printf("mode %d  %d warps/SMSP: %.1f cycles per step per warp (latency view), %.1f cycles per step per SMSP (throughput view) -> %.2f TCUPS\n",
               mode, wps, cyc_total / iters, cyc_total / iters / wps, 148.0 * 4 * wps * 512.0 * iters / (best * 1e-3) / 1e12);
    }
    return 0;
}
