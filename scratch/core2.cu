// cost of the diagonal-frame forward16 cell core alone (registers only): cycles per column step per SMSP
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include "../multiplexnanopore_b200/csrc/smx_dp16h.cuh"
template <int K, int MODE>
__global__ void __launch_bounds__(128, 3) core(uint32_t *out, int iters, uint32_t seed, uint32_t one)
{
    uint32_t plo[K], phi[K], Hc[K], El[K];
    for (int k = 0; k < K; ++k) { plo[k] = seed * (k + 3); phi[k] = seed * (k + 11); Hc[k] = seed + k + threadIdx.x; El[k] = seed ^ k; }
    uint32_t hd = seed, out_h = seed + 1, out_f = seed + 2, sel = 0x4c80 | (threadIdx.x & 3), acc = 0, acc2 = 0;
    const uint32_t negc2 = 0xfffdfffeu;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        uint32_t hu = out_h, fu = out_f;
        if (MODE == 1) { hu = __shfl_up_sync(0xffffffffu, out_h, 1); fu = __shfl_up_sync(0xffffffffu, out_f, 1); sel = __shfl_up_sync(0xffffffffu, sel, 1) | 0x4c80; }
        Smx16hCol c;
        c.hd = hd; c.fu = fu; c.hcu = hu; c.w_lo[0] = 0; c.w_hi[0] = 0; c.w_lo[1] = 0; c.w_hi[1] = 0; c.one = one;
        Smx16hCell<K, 0>::run(Hc, El, plo, phi, c, sel, negc2);
        hd = hu; out_h = c.hcu; out_f = c.fu;
        acc ^= (c.w_lo[0] + 3 * c.w_lo[1]); acc2 += (c.w_hi[0] ^ (5 * c.w_hi[1]));
    }
    uint32_t s = acc ^ acc2 ^ out_h ^ out_f;
    for (int k = 0; k < K; ++k) s ^= Hc[k] ^ El[k];
    if (s == 0x7fffffff) out[0] = s;
}
template <int K, int MODE> void bench(uint32_t *d, int clk) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int wps = 1; wps <= 3; ++wps) {
        const int blocks = 148 * wps;
        float best = 1e30f;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0); core<K, MODE><<<blocks, 128>>>(d, iters, 12345u, 1u); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        const double cyc_total = best * 1e-3 * (double)clk * 1e3;
        printf("K=%d mode %d  %d warps/SMSP: %.1f cycles/step/warp, %.1f cycles/step/SMSP -> %.2f TCUPS\n",
               K, MODE, wps, cyc_total / iters, cyc_total / iters / wps, 148.0 * 4 * wps * 64.0 * K * iters / (best * 1e-3) / 1e12);
    }
}
int main() {
    uint32_t *d; cudaMalloc(&d, 64);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    bench<8, 0>(d, clk); bench<8, 1>(d, clk); bench<16, 0>(d, clk); bench<16, 1>(d, clk);
    return 0;
}
