// pipe-balance microbenchmark: cost (cycles per SMSP) of the forward16 building block in several encodings
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#define MAXTR(FLAGOPS) \
    asm("{.reg .pred pu, pv; .reg .s16 r0, r1, r2, r3;\n\t" \
        "max.s16x2 %0, %3, %4;\n\t mov.b32 {r0, r1}, %0;\n\t mov.b32 {r2, r3}, %3;\n\t" \
        "setp.eq.s16 pv, r0, r2;\n\t setp.eq.s16 pu, r1, r3;\n\t" FLAGOPS "}\n\t" \
        : "=r"(v[c]), "+r"(w0[c & 3]), "+r"(w1[c & 3]) : "r"(v[c]), "r"(b + u), "r"(one))
template <int WHICH>
__global__ void __launch_bounds__(128) k(int *out, int iters, int seed, uint32_t one)
{
    uint32_t v[8], w0[4] = {0, 0, 0, 0}, w1[4] = {0, 0, 0, 0};
    for (int c = 0; c < 8; ++c) v[c] = seed + threadIdx.x * 8 + c;
    const uint32_t b = seed | 1, d = seed ^ 0x55;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                if (WHICH == 0) { MAXTR("@!pv or.b32 %1, %1, 4;\n\t @!pu or.b32 %2, %2, 4;\n\t"); }
                else if (WHICH == 1) { MAXTR("@!pv mad.lo.u32 %1, %5, 4, %1;\n\t @!pu mad.lo.u32 %2, %5, 4, %2;\n\t"); }
                else if (WHICH == 2) { MAXTR("@!pv or.b32 %1, %1, 4;\n\t @!pu mad.lo.u32 %2, %5, 4, %2;\n\t"); }
                else if (WHICH == 3) { MAXTR("@!pv or.b32 %1, %1, 4;\n\t @!pu or.b32 %2, %2, 4;\n\t"); asm("add.s16x2 %0, %1, %2;" : "=r"(v[c]) : "r"(v[c]), "r"(d)); }
                else if (WHICH == 4) { MAXTR("@!pv mad.lo.u32 %1, %5, 4, %1;\n\t @!pu mad.lo.u32 %2, %5, 4, %2;\n\t"); asm("add.s16x2 %0, %1, %2;" : "=r"(v[c]) : "r"(v[c]), "r"(d)); }
                else if (WHICH == 5) { MAXTR("@!pv or.b32 %1, %1, 4;\n\t @!pu mad.lo.u32 %2, %5, 4, %2;\n\t"); asm("add.s16x2 %0, %1, %2;" : "=r"(v[c]) : "r"(v[c]), "r"(d)); }
                else if (WHICH == 6) { asm("mad.lo.u32 %0, %1, 4, %0;" : "+r"(v[c]) : "r"(one)); }
                else if (WHICH == 7) { asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(v[c]) : "r"(one), "r"(b + u)); }
                else if (WHICH == 8) { asm("add.u32 %0, %0, %1;" : "+r"(v[c]) : "r"(b + u)); }
                else if (WHICH == 9) { asm("max.s16x2 %0, %0, %1;" : "+r"(v[c]) : "r"(b + u)); }
                else if (WHICH == 10) { asm("max.s16x2 %0, %0, %1;" : "+r"(v[c]) : "r"(b + u)); asm("mad.lo.u32 %0, %1, 4, %0;" : "+r"(w0[c & 3]) : "r"(one)); }
                else if (WHICH == 11) { asm("max.s16x2 %0, %0, %1;" : "+r"(v[c]) : "r"(b + u)); asm("add.u32 %0, %0, %1;" : "+r"(w0[c & 3]) : "r"(b + u)); }
                else if (WHICH == 12) { asm("max.s16x2 %0, %0, %1;" : "+r"(v[c]) : "r"(b + u)); asm("mad.lo.u32 %0, %1, 4, %0;" : "+r"(w0[c & 3]) : "r"(one)); asm("mad.lo.u32 %0, %1, 8, %0;" : "+r"(w1[c & 3]) : "r"(one)); }
                else if (WHICH == 13) { asm("add.s16x2 %0, %0, %1;" : "+r"(v[c]) : "r"(d)); }
                else if (WHICH == 14) { MAXTR("@!pv sub.u32 %1, %1, -4;\n\t @!pu sub.u32 %2, %2, -4;\n\t"); }
                else if (WHICH == 15) { MAXTR(""); }
            }
        }
    }
    uint32_t s = 0;
    for (int c = 0; c < 8; ++c) s ^= v[c];
    for (int c = 0; c < 4; ++c) s ^= w0[c] ^ w1[c];
    if (s == 0x7fffffff) out[0] = s;
}
template <int W> float run(int *d, int blocks, int iters) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0); k<W><<<blocks, 128>>>(d, iters, 123, 1u); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    return best;
}
int main(int argc, char **argv) {
    int *d; cudaMalloc(&d, 64);
    const int wps = argc > 1 ? atoi(argv[1]) : 4;  // warps per SMSP
    const int blocks = 148 * wps, iters = 4000;
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const char *names[] = {"max+2p+2 @P OR(alu)", "max+2p+2 @P IMAD", "max+2p+1 OR+1 IMAD", "0 + add16x2", "1 + add16x2", "2 + add16x2",
                           "IMAD r*imm+r", "IMAD r*r+r", "IADD", "max.s16x2", "max16+IMAD", "max16+IADD", "max16+2 IMAD", "add.s16x2", "max+2p+2 @P SUB", "max+2p only"};
    float ms[16];
    ms[0] = run<0>(d, blocks, iters); ms[1] = run<1>(d, blocks, iters); ms[2] = run<2>(d, blocks, iters); ms[3] = run<3>(d, blocks, iters);
    ms[4] = run<4>(d, blocks, iters); ms[5] = run<5>(d, blocks, iters); ms[6] = run<6>(d, blocks, iters); ms[7] = run<7>(d, blocks, iters);
    ms[8] = run<8>(d, blocks, iters); ms[9] = run<9>(d, blocks, iters); ms[10] = run<10>(d, blocks, iters); ms[11] = run<11>(d, blocks, iters);
    ms[12] = run<12>(d, blocks, iters); ms[13] = run<13>(d, blocks, iters); ms[14] = run<14>(d, blocks, iters); ms[15] = run<15>(d, blocks, iters);
    for (int w = 0; w < 16; ++w) {
        // units per SMSP = wps warps * iters * 64
        const double cyc = ms[w] * 1e-3 * (double)clk * 1e3 / ((double)wps * iters * 64.0);
        printf("%-24s %7.2f ms  %6.2f cycles per unit per SMSP (%d warps/SMSP)\n", names[w], ms[w], cyc, wps);
    }
    return 0;
}
