// which pipes do VIADD / IADD3 / IMAD / VIMNMX.U16x2 use, and which pairs overlap?  (cycles per "unit" per SMSP)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
template <int WHICH>
__global__ void __launch_bounds__(128) k(uint32_t *out, int iters, uint32_t seed, uint32_t one)
{
    uint32_t v[8], w[8], x[8];
    for (int c = 0; c < 8; ++c) { v[c] = seed + threadIdx.x * 8 + c; w[c] = seed * (c + 3); x[c] = seed ^ (c * 77); }
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int d = (c + 3) & 7;
                if (WHICH == 0) { asm volatile("add.u32 %0, %0, 0x1234;" : "+r"(v[c])); }                       // add imm
                else if (WHICH == 1) { asm volatile("add.u32 %0, %0, %1;" : "+r"(v[c]) : "r"(w[d])); }          // add reg
                else if (WHICH == 2) { asm volatile("mad.lo.u32 %0, %1, 0x1234, %0;" : "+r"(v[c]) : "r"(one)); }  // IMAD r*imm+r
                else if (WHICH == 3) { asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(v[c]) : "r"(one), "r"(w[d])); }  // IMAD rrr
                else if (WHICH == 4) { asm volatile("max.u16x2 %0, %0, %1;" : "+r"(v[c]) : "r"(w[d])); }
                else if (WHICH == 5) { asm volatile("max.u16x2 %0, %0, %1;" : "+r"(v[c]) : "r"(w[d])); asm volatile("mad.lo.u32 %0, %1, 0x1234, %0;" : "+r"(x[c]) : "r"(one)); }
                else if (WHICH == 6) { asm volatile("max.u16x2 %0, %0, %1;" : "+r"(v[c]) : "r"(w[d])); asm volatile("add.u32 %0, %0, 0x1234;" : "+r"(x[c])); }
                else if (WHICH == 7) { asm volatile("max.u16x2 %0, %0, %1;" : "+r"(v[c]) : "r"(w[d])); asm volatile("mad.lo.u32 %0, %1, 0x1234, %0;" : "+r"(x[c]) : "r"(one)); asm volatile("add.u32 %0, %0, 0x1234;" : "+r"(w[c])); }
                else if (WHICH == 8) { asm volatile("mad.lo.u32 %0, %1, 0x1234, %0;" : "+r"(x[c]) : "r"(one)); asm volatile("add.u32 %0, %0, 0x1234;" : "+r"(w[c])); }
                else if (WHICH == 9) { asm volatile("max.u16x2 %0, %0, %1;" : "+r"(v[c]) : "r"(w[d])); asm volatile("mad.lo.u32 %0, %1, 0x1234, %0;" : "+r"(x[c]) : "r"(one)); asm volatile("mad.lo.u32 %0, %1, 0x4321, %0;" : "+r"(w[c]) : "r"(one)); }
                else if (WHICH == 10) { asm volatile("max.u16x2 %0, %0, %1;" : "+r"(v[c]) : "r"(w[d])); asm volatile("add.u32 %0, %0, 0x1234;" : "+r"(x[c])); asm volatile("add.u32 %0, %0, 0x4321;" : "+r"(w[c])); }
                else if (WHICH == 11) { asm volatile("lop3.b32 %0, %0, %1, 0x1234, 0x96;" : "+r"(v[c]) : "r"(w[d])); }
                else if (WHICH == 12) { asm volatile("prmt.b32 %0, %0, %1, 0x3210;" : "+r"(v[c]) : "r"(w[d])); }
            }
        }
    }
    uint32_t s = 0;
    for (int c = 0; c < 8; ++c) s ^= v[c] ^ w[c] ^ x[c];
    if (s == 0x7fffffff) out[0] = s;
}
template <int W> void run(uint32_t *d, const char *name, int wps, int clk) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4000; float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0); k<W><<<148 * wps, 128>>>(d, iters, 123u, 1u); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    printf("%-34s %6.2f cycles per unit per SMSP (%d warps/SMSP)\n", name, best * 1e-3 * (double)clk * 1e3 / ((double)wps * iters * 64.0), wps);
}
int main(int argc, char **argv) {
    uint32_t *d; cudaMalloc(&d, 64);
    const int wps = argc > 1 ? atoi(argv[1]) : 4;
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    run<0>(d, "add r,imm", wps, clk); run<1>(d, "add r,r", wps, clk); run<2>(d, "IMAD r*imm+r", wps, clk); run<3>(d, "IMAD r*r+r", wps, clk);
    run<4>(d, "max.u16x2", wps, clk); run<5>(d, "max16 + IMADimm", wps, clk); run<6>(d, "max16 + add imm", wps, clk);
    run<7>(d, "max16 + IMADimm + add imm", wps, clk); run<8>(d, "IMADimm + add imm", wps, clk); run<9>(d, "max16 + 2 IMADimm", wps, clk);
    run<10>(d, "max16 + 2 add imm", wps, clk); run<11>(d, "lop3", wps, clk); run<12>(d, "prmt", wps, clk);
    return 0;
}
