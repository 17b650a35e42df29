// Shared device/host definitions of the SMX engine (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define SMX_FULL_MASK 0xffffffffu
#define SMX_NEG_INF (INT32_MIN / 2)  // same sentinel as parasail's NEG_INF_32

// One DP problem as seen by the kernels (48 bytes, 16-byte aligned).
struct SmxProblem {
    int64_t s1_off;     // query offset in the pool
    int64_t s2_off;     // reference offset in the pool
    int32_t m;          // query length  (rows)
    int32_t n;          // reference length (columns)
    int32_t mode;       // SMX_MODE_*
    int16_t kshift;     // log2(rows per lane) of its kernel class: K = 1 << kshift
    int16_t cls;        // forward kernel class (host planning only)
    int64_t trace_off;  // byte offset of its packed direction bits in the trace arena
    int64_t cigar_off;  // offset (in uint32 runs) of its scratch run buffer
};

struct SmxResult {
    int32_t score, end_q, end_r, beg_q, beg_r;
    int32_t n_runs;     // runs to deliver (for re-clipped problems: the kept, aligned part only)
    int32_t run_start;  // index of the first delivered run in the problem's scratch (walk order)
    int32_t n_s, n_h;   // re-clipped problems: query bases turned into 'S', reference bases turned into 'H'
    int32_t pad0, pad1, pad2;
};

#define SMX_MODE_MASK 0xff
#define SMX_MODE_CLIP 0x100  // semi-global problem whose CIGAR is re-clipped on the device (MyResult.set_soft_clipping)
#define SMX_MODE_CKPT 0x200  // packed problem whose forward pass keeps checkpoints instead of direction bits (smx_traceback_ck.cuh)

struct SmxScoring {
    int32_t go, ge, match, mismatch;
};

// ASCII -> parasail "ACGT" mapper code (0..3, everything else 4); input is upper-case.
__device__ __forceinline__ int smx_code(unsigned c)
{
    return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : 4;
}

// trace nibble: bits 0-1 H source, bit 2 = E opened from H (DIAG_E), bit 3 = F opened from H (DIAG_F)
#define SMX_SRC_ZERO 0u  // local alignment stop
#define SMX_SRC_DIAG 1u
#define SMX_SRC_DEL 2u   // from F (vertical, consumes a query base, CIGAR 'I')
#define SMX_SRC_INS 3u   // from E (horizontal, consumes a reference base, CIGAR 'D')
#define SMX_BIT_EOPEN 4u
#define SMX_BIT_FOPEN 8u
