// C-ABI of the SMX engine (include/smx.h): handle, device workspaces, planning and launches.
// Host-side C++ only orchestrates; all arithmetic of the path runs in the kernels included below.
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <ctime>
#include <condition_variable>
#include <mutex>
#include <new>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

#include "../../include/smx.h"
#include "smx_common.cuh"
#include "smx_dp.cuh"
#include "smx_dp16.cuh"
#include "smx_dp16h.cuh"
#include "smx_scan.cuh"
#include "smx_traceback.cuh"
#include "smx_traceback_ck.cuh"
#include "smx_outputs.cuh"
#include "smx_msa.cuh"
#include "smx_consensus.cuh"

namespace {

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
};

// page-locked host staging (grow-only): large D2H / H2D copies run at PCIe speed instead of through the
// driver's pageable bounce buffers
struct PinBuf {
    void *p = nullptr;
    size_t cap = 0;
};

struct Chunk {
    int first = 0, count = 0;            // range in the size-sorted order
    int n_big = 0;                       // traceback: the first n_big problems of the chunk's order_all range get a warp each
    int n_ck = 0;                        // the last n_ck problems of that range are checkpointed (smx_traceback_ck)
    int cls_first[14] = {0}, cls_count[14] = {0};  // per kernel class ranges inside order_cls (see problem_class)
    int cls_local[14] = {0};             // packed classes: how many problems at the END of the class range are local (sw_trace)
    int cls_ck[14] = {0};                // packed classes: how many problems before the local ones are checkpointed (SMX_MODE_CKPT)
    size_t trace_bytes = 0;
};

}  // namespace

struct smx_handle {
    int device = 0;
    cudaStream_t stream = nullptr;     // forward DP + traceback (lowest priority: long launches)
    cudaStream_t stream_hi = nullptr;  // everything short: pool upload, scan, select, chain, compaction, copies (highest priority)
    bool owns_stream = false;
    int sm_count = 148;
    std::string err;
    int64_t trace_budget = (int64_t)16 << 30;  // clamped to an eighth of the free device memory at creation
    bool use_s16 = true;  // SMX_S16=0 forces the int32 kernels (cross-check)
    int s16_min_rows = 32;   // the diagonal-frame kernel takes problems of more rows than this (SMX_S16_MINROWS overrides)
    int64_t team_min_cells = 0;  // SMX_TEAM_MIN_MCELLS (in Mi cells): below it one warp per problem (measured: no gain on C2)
    bool persistent_fwd = false;  // SMX_PERSISTENT=1: launch smx_dp16h.cuh as a persistent grid (A/B switch)
    bool use_s16h = true;  // diagonal-frame u16x2 kernel (smx_dp16h.cuh) when the scoring allows it; SMX_S16H=0 keeps smx_dp16.cuh
    int s16_max_wsel = 3;  // upper bound of the team width selector (SMX_S16_MAXW=1|2|4|8; 1 = one warp per problem throughout)
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_fwd = nullptr, ev_sync_lo = nullptr, ev_sync_hi = nullptr;
    // the largest problems of a chunk run as teams of 2 / 4 / 8 warps NEXT TO the one-warp-per-problem launch (own low-priority
    // streams, forked and joined by events) so that their critical path does not become the tail of the forward phase
    cudaStream_t stream_team[4] = {nullptr, nullptr, nullptr, nullptr};  // 2 / 4 / 8-warp teams; [3]: the checkpointed one-warp launch
    cudaEvent_t ev_fork = nullptr, ev_join[4] = {nullptr, nullptr, nullptr, nullptr};
    bool teams_concurrent = false;
    // SMX_GUARD=1 (debug; compute-sanitizer is closed on the pool): 256 poisoned bytes before and after every problem's
    // direction-bit region, verified after the traceback of every chunk
    bool guard = false;
    DevBuf d_guard_err;
    int64_t guard_errors = -1;
    int64_t team_cells[3] = {16ll << 20, 48ll << 20, 128ll << 20};  // cells from which a problem gets 2 / 4 / 8 warps (SMX_TEAM2/4/8_MCELLS)
    float last_ms = 0.f;
    int last_launches = 0;
    // per-kernel timing of the last smx_align_run: [0..13] forward classes (problem_class), [14] traceback, [15] compaction
    std::vector<cudaEvent_t> ev_pool;
    std::vector<int> ev_tag;
    size_t ev_used = 0;
    double kernel_ms[16] = {0};
    int kernel_launches[16] = {0};
    int64_t class_cells[14] = {0};
    int64_t ck_cells = 0, ck_problems = 0;  // of them checkpointed (SMX_MODE_CKPT)

    DevBuf pool, pool2, pool_special;  // ASCII pool, its two bit planes and per-word non-ACGT flags (smx_scan.cuh)
    int64_t pool_len = 0;
    std::vector<int32_t> pool_nonacgt;  // prefix count of bytes other than A/C/G/T (s16 kernel eligibility)

    // align state
    int32_t n = 0;
    SmxScoring sc{};
    int64_t cells = 0;
    std::vector<SmxProblem> probs;
    std::vector<int32_t> order_cls;  // problem ids grouped by chunk then class, largest first
    std::vector<int32_t> order_all;  // problem ids by chunk (for traceback)
    std::vector<Chunk> chunks;
    int32_t nmax = 0;
    size_t max_chunk_trace = 0;
    int64_t runs_scratch = 0;
    int64_t total_runs = -1;
    DevBuf d_probs, d_order_cls, d_order_all, d_results, d_trace, d_boundary, d_runs, d_dense, d_dense_off, d_tickets, d_slots;
    int split_last = 1;  // SMX_S16H_SPLIT=0: lone last bands at K = 8 (cross-check)
    // checkpointed traceback (smx_traceback_ck.cuh) for packed global / semi-global problems of at least this size;
    // SMX_CKPT=0 switches it off, SMX_CKPT_MINROWS / SMX_CKPT_MINCOLS move the thresholds (rows > 512)
    bool ckpt = true;
    bool ck_side = true;  // SMX_CK_SIDE=0: the checkpointed launches queue on the handle's stream instead of running beside the others
    int ck_min_rows = 513, ck_min_cols = 1500;
    DevBuf d_compact, d_expand_items, d_comp;  // pipeline pool: sequences as given, expansion list, complement table
    std::vector<SmxResult> h_results;
    std::vector<int64_t> h_dense_off;

    // scan state
    int32_t scan_n = 0, scan_topology = 0;
    int64_t scan_total = 0, scan_cells = 0;
    int32_t scan_tiles = 0, scan_nq_max = 0;
    DevBuf d_scan_pairs, d_scan_tile_pair, d_counts;

    // pipeline state (smx_pipeline.cuh)
    DevBuf d_sel_pairs, d_sel_out, d_regions;
    DevBuf d_chain_tasks, d_chain_regions, d_chain_traces, d_chain_out, d_chain_scratch;
    std::vector<uint8_t> host_pool;
    PinBuf pin_regions, pin_cigar, pin_pool, pin_special;
    PinBuf pin_stage[16];  // page-locked staging of the small descriptor / result copies (PinSlot)
    struct SmxPipelineState *pipe = nullptr;
    // smx_ops_run results (row N1): final column runs per operation
    std::vector<uint32_t> ops_runs;
    std::vector<int64_t> ops_off;
    std::vector<uint8_t> ops_status;
    bool ops_valid = false;
    // smx_msa_run results (row N3): column matrices on the device until fetched
    DevBuf d_msa_in, d_msa_work, d_msa_out, d_cons;
    int64_t msa_cols = -1;
    int32_t msa_reads = 0;
};

static const int kLineSlots = 4096;  // boundary-line pairs of the non-persistent forward launches (> resident workers of any class)
static thread_local std::string g_create_error;

static int fail(smx_handle *h, int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (h) h->err = buf; else g_create_error = buf;
    return code;
}

#define SMX_CUDA(h, call)                                                                           \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess)                                                                      \
            return fail((h), -2, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

static int ensure(smx_handle *h, DevBuf &b, size_t bytes)
{
    if (bytes <= b.cap) return 0;
    static const bool log_on = getenv("SMX_PHASE_LOG") != nullptr;
    if (log_on) fprintf(stderr, "REALLOC dev %zu -> %zu\n", b.cap, bytes);
    if (b.p) { cudaFree(b.p); b.p = nullptr; b.cap = 0; }
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) {
        want = bytes;
        e = cudaMalloc(&b.p, want);
    }
    if (e != cudaSuccess) { b.p = nullptr; return fail(h, -3, "cudaMalloc(%zu bytes) failed: %s", want, cudaGetErrorString(e)); }
    b.cap = want;
    return 0;
}

static int ensure_pinned(smx_handle *h, PinBuf &b, size_t bytes)
{
    if (bytes <= b.cap) return 0;
    static const bool log_on = getenv("SMX_PHASE_LOG") != nullptr;
    if (log_on) fprintf(stderr, "REALLOC pin %zu -> %zu\n", b.cap, bytes);
    if (b.p) { cudaFreeHost(b.p); b.p = nullptr; b.cap = 0; }
    const size_t want = bytes + bytes / 4 + 4096;
    cudaError_t e = cudaHostAlloc(&b.p, want, cudaHostAllocDefault);
    if (e != cudaSuccess) { b.p = nullptr; return fail(h, -3, "cudaHostAlloc(%zu bytes) failed: %s", want, cudaGetErrorString(e)); }
    b.cap = want;
    return 0;
}

// Every host<->device copy of the path goes through page-locked memory: a cudaMemcpyAsync on pageable memory is staged
// by the driver under locks shared by all streams of the context, and with several handles in flight the small
// descriptor / result copies of a batch's front end were measured to wait 30 ms and more behind the other handles'
// work (the select phase took 41 ms of which the kernel 9 ms).
enum PinSlot { PIN_PROBS, PIN_ORDER_CLS, PIN_ORDER_ALL, PIN_RESULTS, PIN_DENSE_OFF, PIN_SCAN_PAIRS, PIN_SCAN_TILES, PIN_SEL_PAIRS,
               PIN_SEL_OUT, PIN_NREG, PIN_CHAIN_TASKS, PIN_CHAIN_REGS, PIN_CHAIN_OUT, PIN_CHAIN_TR, PIN_EXPAND };

// host -> device through staging slot `slot`; the slot must not be reused before `stream` has been synchronised
static int h2d(smx_handle *h, int slot, void *dst_dev, const void *src, size_t bytes, cudaStream_t stream)
{
    if (bytes == 0) return 0;
    if (int rc = ensure_pinned(h, h->pin_stage[slot], bytes)) return rc;
    memcpy(h->pin_stage[slot].p, src, bytes);
    SMX_CUDA(h, cudaMemcpyAsync(dst_dev, h->pin_stage[slot].p, bytes, cudaMemcpyHostToDevice, stream));
    return 0;
}

// device -> staging slot `slot` (asynchronous); read h->pin_stage[slot].p after synchronising `stream`
static int d2h(smx_handle *h, int slot, const void *src_dev, size_t bytes, cudaStream_t stream)
{
    if (bytes == 0) return 0;
    if (int rc = ensure_pinned(h, h->pin_stage[slot], bytes)) return rc;
    SMX_CUDA(h, cudaMemcpyAsync(h->pin_stage[slot].p, src_dev, bytes, cudaMemcpyDeviceToHost, stream));
    return 0;
}

static void release(DevBuf &b)
{
    if (b.p) cudaFree(b.p);
    b.p = nullptr; b.cap = 0;
}

static void smx_pipeline_free(smx_handle *h);

// records an event after the launch just issued and tags the interval that ends here
static int mark(smx_handle *h, int tag, cudaStream_t stream = nullptr)
{
    if (!stream) stream = h->stream;
    if (h->ev_used == h->ev_pool.size()) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return fail(h, -2, "cudaEventCreate failed");
        h->ev_pool.push_back(e);
        h->ev_tag.push_back(0);
    }
    h->ev_tag[h->ev_used] = tag;
    cudaError_t e = cudaEventRecord(h->ev_pool[h->ev_used], stream);
    if (e != cudaSuccess) return fail(h, -2, "cudaEventRecord failed: %s", cudaGetErrorString(e));
    h->ev_used++;
    return 0;
}

// Host waits sleep instead of spinning: with several handles per GPU and one process per GPU a spinning
// cudaStreamSynchronize per waiting thread eats the host cores the pipeline's own stages need (8 GPUs x 6 handles on 32
// cores).  A blocking-sync event is recorded behind the work and waited for.
static cudaError_t smx_sync(smx_handle *h, cudaStream_t stream)
{
    static const bool spin = [] { const char *e = getenv("SMX_SPIN_SYNC"); return e && e[0] == '1'; }();
    cudaEvent_t ev = stream == h->stream ? h->ev_sync_lo : h->ev_sync_hi;
    if (!ev || spin) return cudaStreamSynchronize(stream);
    cudaError_t e = cudaEventRecord(ev, stream);
    if (e != cudaSuccess) return e;
    return cudaEventSynchronize(ev);
}

extern "C" int smx_version(void) { return 200; }

extern "C" int smx_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" const char *smx_last_error(const smx_handle *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

extern "C" int smx_create(smx_handle **out, int device_ordinal, void *stream)
{
    if (!out) return fail(nullptr, -1, "smx_create: out is NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0)
        return fail(nullptr, -2, "smx_create: no CUDA device available (%s); this engine has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    if (device_ordinal < 0 || device_ordinal >= ndev) return fail(nullptr, -1, "smx_create: bad device ordinal %d", device_ordinal);
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device_ordinal)) != cudaSuccess)
        return fail(nullptr, -2, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major != 10)
        return fail(nullptr, -4, "smx_create: device %d is sm_%d%d; this library is built for sm_100a (B200) only",
                    device_ordinal, prop.major, prop.minor);
    smx_handle *h = new (std::nothrow) smx_handle();
    if (!h) return fail(nullptr, -3, "out of host memory");
    h->device = device_ordinal;
    h->sm_count = prop.multiProcessorCount;
    if ((e = cudaSetDevice(device_ordinal)) != cudaSuccess) { delete h; return fail(nullptr, -2, "cudaSetDevice: %s", cudaGetErrorString(e)); }
    if (stream) { h->stream = (cudaStream_t)stream; h->stream_hi = h->stream; h->owns_stream = false; }
    else {
        // Several handles share a GPU (batches in flight).  The short kernels of a batch's front end must not queue
        // behind the long forward-DP launches of the other handles, or all handles end up in their host phases
        // at the same time (measured: GPU idle 22 % of the wall clock): they run on a high-priority stream, and the
        // forward kernels are launched non-persistent so that SM slots free up continuously.
        int prio_least = 0, prio_greatest = 0;
        cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest);
        if ((e = cudaStreamCreateWithPriority(&h->stream, cudaStreamNonBlocking, prio_least)) != cudaSuccess) { delete h; return fail(nullptr, -2, "cudaStreamCreate: %s", cudaGetErrorString(e)); }
        if ((e = cudaStreamCreateWithPriority(&h->stream_hi, cudaStreamNonBlocking, prio_greatest)) != cudaSuccess) { cudaStreamDestroy(h->stream); delete h; return fail(nullptr, -2, "cudaStreamCreate: %s", cudaGetErrorString(e)); }
        h->owns_stream = true;
        for (int t = 0; t < 4; ++t) {
            if (cudaStreamCreateWithPriority(&h->stream_team[t], cudaStreamNonBlocking, prio_least) != cudaSuccess) h->stream_team[t] = nullptr;
            cudaEventCreateWithFlags(&h->ev_join[t], cudaEventDisableTiming);
        }
        cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming);
    }
    cudaEventCreate(&h->ev0);
    cudaEventCreate(&h->ev1);
    cudaEventCreateWithFlags(&h->ev_fwd, cudaEventDisableTiming | cudaEventBlockingSync);
    cudaEventCreateWithFlags(&h->ev_sync_lo, cudaEventDisableTiming | cudaEventBlockingSync);
    cudaEventCreateWithFlags(&h->ev_sync_hi, cudaEventDisableTiming | cudaEventBlockingSync);
    if (ensure(h, h->d_slots, sizeof(int32_t) * kLineSlots) != 0 || cudaMemset(h->d_slots.p, 0, sizeof(int32_t) * kLineSlots) != cudaSuccess) {
        g_create_error = "smx_create: cannot allocate the boundary-line slot table";
        smx_destroy(h);
        return -3;
    }
    {
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) h->trace_budget = std::max<int64_t>((int64_t)2 << 30, std::min<int64_t>(h->trace_budget, (int64_t)(free_b / 8)));
    }
    if (const char *tb = getenv("SMX_TRACE_BUDGET_GB")) { const long v = atol(tb); if (v >= 1 && v <= 160) h->trace_budget = (int64_t)v << 30; }
    *out = h;
    return 0;
}

extern "C" void smx_destroy(smx_handle *h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    cudaStreamSynchronize(h->stream_hi);
    DevBuf *bufs[] = {&h->pool, &h->pool2, &h->pool_special, &h->d_probs, &h->d_order_cls, &h->d_order_all, &h->d_results, &h->d_trace, &h->d_boundary,
                      &h->d_runs, &h->d_dense, &h->d_dense_off, &h->d_tickets, &h->d_slots, &h->d_guard_err, &h->d_msa_in, &h->d_msa_work, &h->d_msa_out, &h->d_cons, &h->d_scan_pairs, &h->d_scan_tile_pair, &h->d_counts,
                      &h->d_sel_pairs, &h->d_sel_out, &h->d_regions, &h->d_chain_tasks, &h->d_chain_regions, &h->d_chain_traces,
                      &h->d_chain_out, &h->d_chain_scratch, &h->d_compact, &h->d_expand_items, &h->d_comp};
    smx_pipeline_free(h);
    for (PinBuf *b : {&h->pin_regions, &h->pin_cigar, &h->pin_pool, &h->pin_special}) if (b->p) cudaFreeHost(b->p);
    for (PinBuf &b : h->pin_stage) if (b.p) cudaFreeHost(b.p);
    for (DevBuf *b : bufs) release(*b);
    for (cudaEvent_t e : h->ev_pool) cudaEventDestroy(e);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->ev_fwd) cudaEventDestroy(h->ev_fwd);
    if (h->ev_sync_lo) cudaEventDestroy(h->ev_sync_lo);
    if (h->ev_sync_hi) cudaEventDestroy(h->ev_sync_hi);
    if (h->owns_stream) { cudaStreamDestroy(h->stream); cudaStreamDestroy(h->stream_hi); }
    for (int t = 0; t < 4; ++t) { if (h->stream_team[t]) cudaStreamDestroy(h->stream_team[t]); if (h->ev_join[t]) cudaEventDestroy(h->ev_join[t]); }
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    delete h;
}

extern "C" int64_t smx_get_trace_budget(const smx_handle *h) { return h ? h->trace_budget : 0; }

extern "C" int smx_set_trace_budget(smx_handle *h, int64_t bytes)
{
    if (!h) return -1;
    if (bytes < (1 << 20)) return fail(h, -1, "trace budget must be at least 1 MiB");
    h->trace_budget = bytes;
    return 0;
}

static int pool_finish(smx_handle *h, int64_t n);

extern "C" int smx_pool_upload(smx_handle *h, const uint8_t *seqs, int64_t n)
{
    if (!h) return -1;
    if (!seqs || n <= 0) return fail(h, -1, "smx_pool_upload: empty pool");
    SMX_CUDA(h, cudaSetDevice(h->device));
    if (int rc = ensure(h, h->pool, (size_t)n + 64)) return rc;
    SMX_CUDA(h, cudaMemcpyAsync(h->pool.p, seqs, (size_t)n, cudaMemcpyHostToDevice, h->stream_hi));
    return pool_finish(h, n);
}

// The pipeline's pool is built on the device: the sequences cross PCIe once, as given; every pool entry is a sequence
// (or its reverse complement, Biopython's IUPAC table) written twice back to back.
struct SmxExpandItem { int64_t src, dst; int32_t len, rc; };

__global__ void __launch_bounds__(256) smx_expand_pool_kernel(const uint8_t *compact, const SmxExpandItem *items, const uint8_t *comp, uint8_t *pool)
{
    __shared__ uint8_t s_comp[256];
    s_comp[threadIdx.x] = comp[threadIdx.x];
    __syncthreads();
    const SmxExpandItem it = items[blockIdx.x];
    const uint8_t *src = compact + it.src;
    uint8_t *dst = pool + it.dst;
    for (int i = threadIdx.x; i < it.len; i += 256) {
        uint8_t b = src[it.rc ? it.len - 1 - i : i];
        if (b >= 'a' && b <= 'z') b -= 32;   // MySeq upper-cases on construction (my_classes.py:271-272): done here, not in a host pass
        if (it.rc) b = s_comp[b];
        dst[i] = b;
        dst[it.len + i] = b;
    }
}

// bit planes, non-ACGT flags and the host-side block prefix for the pool bytes already in h->pool (on stream_hi)
static int pool_finish(smx_handle *h, int64_t n)
{
    {
        const int64_t n_words = (n + 31) / 32;  // bit planes: one uint2 per 32 bases
        if (int rc = ensure(h, h->pool2, sizeof(uint2) * (size_t)(n_words + 4))) return rc;
        if (int rc = ensure(h, h->pool_special, (size_t)(n_words + 4))) return rc;
        SMX_CUDA(h, cudaMemsetAsync(h->pool2.p, 0, sizeof(uint2) * (size_t)(n_words + 4), h->stream_hi));
        SMX_CUDA(h, cudaMemsetAsync(h->pool_special.p, 0, (size_t)(n_words + 4), h->stream_hi));
        smx_pack_kernel<<<(unsigned)((n_words + 255) / 256), 256, 0, h->stream_hi>>>((const uint8_t *)h->pool.p, n, (uint2 *)h->pool2.p,
                                                                                 (uint8_t *)h->pool_special.p, n_words);
        SMX_CUDA(h, cudaGetLastError());
        // the per-word "holds a byte other than A/C/G/T" flags come back for the host-side kernel choice (1 byte per 32 bases)
        if (int rc = ensure_pinned(h, h->pin_special, (size_t)n_words)) return rc;
        SMX_CUDA(h, cudaMemcpyAsync(h->pin_special.p, h->pool_special.p, (size_t)n_words, cudaMemcpyDeviceToHost, h->stream_hi));
    }
    SMX_CUDA(h, smx_sync(h, h->stream_hi));
    h->pool_len = n;
    // per 64-byte block: does it hold a byte other than A/C/G/T?  prefix over blocks (s16 kernel eligibility); the flags
    // were computed by smx_pack_kernel (the first version scanned the 30 MB pool of a batch again on 16 host threads)
    {
        const int64_t nb = (n + 63) / 64, n_words = (n + 31) / 32;
        const uint8_t *sp = (const uint8_t *)h->pin_special.p;
        h->pool_nonacgt.resize((size_t)nb + 1);
        h->pool_nonacgt[0] = 0;
        int32_t acc = 0;
        for (int64_t b = 0; b < nb; ++b) {
            acc += (sp[2 * b] | (2 * b + 1 < n_words ? sp[2 * b + 1] : 0)) ? 1 : 0;
            h->pool_nonacgt[(size_t)b + 1] = acc;
        }
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// A8: batched DP
// ------------------------------------------------------------------------------------------------
// kernel class of a problem: kshift * 2 + local for the int32 warp-per-problem kernels (K = 1 << kshift rows per
// lane), 8 + local for the int32 CTA-per-problem (team) kernel that takes every problem of 3 or more bands,
// 10..13 for the packed s16x2 kernel with 1, 2, 4, 8 warps per problem
enum { kClsS16 = 10, kNumFwdClasses = 14, kSlotTraceback = 14, kSlotCompact = 15, kNumSlots = 16 };
static const int kTeamMinRows = 513;
static const int kTbBigWalk = [] { const char *e = getenv("SMX_TB_BIGWALK"); const int v = e ? atoi(e) : 768; return v < 64 ? 64 : v; }();  // m + n from which a traceback walk gets its own warp
static int problem_class(int m, int mode);
static int pick_kshift(int m)
{
    if (m > 128) return 3;
    if (m > 64) return 2;
    if (m > 32) return 1;
    return 0;
}

static int problem_class(int m, int mode)
{
    const int local = mode == SMX_MODE_SW ? 1 : 0;
    if (m >= kTeamMinRows) return 8 + local;
    return pick_kshift(m) * 2 + local;
}

// classes 10 (warp per problem) and 11 (CTA per problem): packed s16x2 kernel, smx_dp16.cuh
// scoring-level condition of the diagonal-frame kernel (smx_dp16h.cuh): stored values never fall along a diagonal
// (mismatch + 2 ge >= 0), profile bytes s + go + ge fit 0..127
static bool s16h_scoring_ok(const SmxScoring &sc)
{
    if (sc.go > 100 || sc.ge > 100) return false;
    const int lo = std::min(std::min(sc.match, sc.mismatch), 0), hi = std::max(std::max(sc.match, sc.mismatch), 0);
    return lo + 2 * sc.ge >= 0 && lo + sc.go + sc.ge >= 0 && hi + sc.go + sc.ge <= 127;
}

static bool s16_eligible(const smx_handle *h, int64_t s2_off, int m, int n, int mode)
{
    if ((mode == SMX_MODE_SW && !h->use_s16h) || m <= (h->use_s16h ? h->s16_min_rows : 256)) return false;  // local mode: diagonal-frame kernel only
    const SmxScoring &sc = h->sc;
    if (h->use_s16h) {
        // biased unsigned halves: BIAS + H^ <= BIAS + max(match, 0) * min(rows, cols) + ge * (rows + cols), with the
        // padded rows and the 96 extra steps a band pair runs past / before the matrix, plus their growth
        const int64_t rows = (int64_t)(((m + 255) / 256 + 1) / 2) * 512, cols = (int64_t)n + 96;
        const int64_t bias = 3ll * sc.go + std::abs(sc.go - sc.ge) + 16 + (mode == SMX_MODE_SW ? 32ll * sc.ge : 0);
        const int64_t top = bias + (int64_t)std::max(sc.match, 0) * std::min(rows, cols) + (int64_t)sc.ge * (rows + cols) +
                            96ll * (std::abs(sc.match) + sc.go + 2 * sc.ge) + 64;
        if (top > 65000) return false;
        const int64_t b0 = s2_off / 64, b1 = (s2_off + n - 1) / 64 + 1;
        return h->pool_nonacgt[(size_t)b1] - h->pool_nonacgt[(size_t)b0] == 0;
    }
    if (sc.go > 256 || sc.ge > 256) return false;
    const int64_t m_pad = (int64_t)(((m + 255) / 256 + 1) / 2) * 512;
    if ((int64_t)sc.go * 3 + (int64_t)sc.ge * (m_pad + n) + 256 > 31000) return false;
    if ((int64_t)std::max(sc.match, 0) * std::min(m, n) > 31000) return false;
    // conservative at block granularity: every 64-byte block the slice touches must be pure ACGT
    const int64_t b0 = s2_off / 64, b1 = (s2_off + n - 1) / 64 + 1;
    return h->pool_nonacgt[(size_t)b1] - h->pool_nonacgt[(size_t)b0] == 0;
}

static int align_prepare_impl(smx_handle *h, const int64_t *s1_off, const int32_t *s1_len, const int64_t *s2_off,
                              const int32_t *s2_len, const uint8_t *mode, const uint8_t *clip, int32_t n, int32_t open, int32_t extend,
                              int32_t match, int32_t mismatch);

extern "C" int smx_align_prepare(smx_handle *h, const int64_t *s1_off, const int32_t *s1_len, const int64_t *s2_off,
                                 const int32_t *s2_len, const uint8_t *mode, int32_t n, int32_t open, int32_t extend,
                                 int32_t match, int32_t mismatch)
{
    return align_prepare_impl(h, s1_off, s1_len, s2_off, s2_len, mode, nullptr, n, open, extend, match, mismatch);
}

// clip[p] != 0 (semi-global modes only): re-clip the CIGAR on the device (rows A9) and deliver only the kept part
static int align_prepare_impl(smx_handle *h, const int64_t *s1_off, const int32_t *s1_len, const int64_t *s2_off,
                              const int32_t *s2_len, const uint8_t *mode, const uint8_t *clip, int32_t n, int32_t open, int32_t extend,
                              int32_t match, int32_t mismatch)
{
    if (!h) return -1;
    h->total_runs = -1;
    h->n = 0;
    if (n < 0 || (n > 0 && (!s1_off || !s1_len || !s2_off || !s2_len || !mode))) return fail(h, -1, "smx_align_prepare: NULL argument");
    if (!h->pool.p) return fail(h, -1, "smx_align_prepare: no sequence pool uploaded");
    if (open <= 0 || extend <= 0) return fail(h, -1, "gap penalties must be positive (my_classes.py:537-541)");
    if (match > 127 || match < -128 || mismatch > 127 || mismatch < -128) return fail(h, -1, "match/mismatch must fit in int8");
    SMX_CUDA(h, cudaSetDevice(h->device));
    h->sc = SmxScoring{open, extend, match, mismatch};
    h->probs.resize((size_t)n);
    h->cells = 0;
    for (int c = 0; c < kNumFwdClasses; ++c) h->class_cells[c] = 0;
    {
        const char *env = getenv("SMX_S16");
        h->use_s16 = !(env && env[0] == '0');
        if (const char *mr = getenv("SMX_S16_MINROWS")) { const int v = atoi(mr); if (v >= 32 && v <= 4096) h->s16_min_rows = v; }
        if (const char *tm = getenv("SMX_TEAM_MIN_MCELLS")) h->team_min_cells = (int64_t)atol(tm) << 20;
        { const char *pe = getenv("SMX_PERSISTENT"); h->persistent_fwd = pe && pe[0] == '1'; }
        { const char *se = getenv("SMX_S16H_SPLIT"); h->split_last = !(se && se[0] == '0'); }
        { const char *ce = getenv("SMX_CKPT"); h->ckpt = !(ce && ce[0] == '0'); }
        { const char *ce = getenv("SMX_CK_SIDE"); h->ck_side = !(ce && ce[0] == '0'); }
        if (const char *e = getenv("SMX_CKPT_MINROWS")) { const int v = atoi(e); if (v > 512) h->ck_min_rows = v; }
        if (const char *e = getenv("SMX_CKPT_MINCOLS")) { const int v = atoi(e); if (v >= 1) h->ck_min_cols = v; }
        const char *envh = getenv("SMX_S16H");
        h->use_s16h = !(envh && envh[0] == '0') && s16h_scoring_ok(h->sc);
        const char *mw = getenv("SMX_S16_MAXW");
        h->s16_max_wsel = 3;
        if (mw) { const int v = atoi(mw); h->s16_max_wsel = v >= 8 ? 3 : v >= 4 ? 2 : v >= 2 ? 1 : 0; }
        static const char *const tenv[3] = {"SMX_TEAM2_MCELLS", "SMX_TEAM4_MCELLS", "SMX_TEAM8_MCELLS"};
        static const int64_t tdef[3] = {16, 48, 128};
        for (int t = 0; t < 3; ++t) { const char *e = getenv(tenv[t]); h->team_cells[t] = e && atof(e) > 0 ? (int64_t)(atof(e) * 1048576.0) : tdef[t] << 20; }
        { const char *ge_ = getenv("SMX_GUARD"); h->guard = ge_ && ge_[0] == '1'; if (!h->guard) h->guard_errors = -1; else if (h->guard_errors < 0) h->guard_errors = 0; }
        const char *tc = getenv("SMX_TEAMS_CONCURRENT");
        h->teams_concurrent = h->use_s16h && !h->persistent_fwd && h->stream_team[0] && h->stream_team[1] && h->stream_team[2] && !(tc && tc[0] == '0');
    }
    h->nmax = 1;
    int64_t runs_off = 0;
    for (int p = 0; p < n; ++p) {
        const int m = s1_len[p], nn = s2_len[p];
        if (m <= 0 || nn <= 0) return fail(h, -1, "problem %d: empty sequence (callers never pass empty strings, ref_query_alignment.py:93-106,332-338)", p);
        if (mode[p] > 3) return fail(h, -1, "problem %d: unknown mode %d", p, (int)mode[p]);
        if (s1_off[p] < 0 || s2_off[p] < 0 || s1_off[p] + m > h->pool_len || s2_off[p] + nn > h->pool_len)
            return fail(h, -1, "problem %d: sequence slice outside the uploaded pool", p);
        // int32 score range: |H| <= open + extend*(m+n) + max(|match|,|mismatch|)*min(m,n)
        const int64_t bound = (int64_t)open + (int64_t)extend * ((int64_t)m + nn) + 128LL * std::min(m, nn);
        if (bound > (int64_t)(1 << 29)) return fail(h, -1, "problem %d: score range exceeds int32 headroom", p);
        SmxProblem &q = h->probs[(size_t)p];
        q.s1_off = s1_off[p]; q.s2_off = s2_off[p]; q.m = m; q.n = nn; q.mode = mode[p];
        if (clip && clip[p]) {
            if (mode[p] != SMX_MODE_SG_QE_DE && mode[p] != SMX_MODE_SG_QB_DB) return fail(h, -1, "problem %d: only semi-global problems can be re-clipped", p);
            if (match <= 0 || mismatch >= 0) return fail(h, -1, "device re-clipping needs match > 0 and mismatch < 0 (my_classes.py:547-552)");
            q.mode |= SMX_MODE_CLIP;
        }
        q.kshift = pick_kshift(m);
        q.cls = problem_class(m, mode[p]);
        if (h->use_s16 && s16_eligible(h, s2_off[p], m, nn, mode[p])) {
            // Teams shorten the critical path of the largest problems -- the tail of the forward phase: one warp needs
            // ~30 ms for a 35 M-cell problem, ~0.4 s for the 400 M cells of a 20 kb x 20 kb head of C4 -- but idle a warp
            // whenever the band pairs do not divide evenly and pay a pipeline fill per band pair: only problems of many
            // cells AND at least two band pairs per warp get one (2 / 4 / 8 warps from 16 / 48 / 128 Mi cells).
            const int pairs = ((m + 255) / 256 + 1) / 2;
            const int64_t pc = (int64_t)m * nn;
            int wsel = pc >= h->team_cells[2] && pairs >= 16 ? 3 : pc >= h->team_cells[1] && pairs >= 8 ? 2 : pc >= h->team_cells[0] && pairs >= 4 ? 1 : 0;
            if (wsel > h->s16_max_wsel) wsel = h->s16_max_wsel;
            if (pc < h->team_min_cells) wsel = 0;
            q.cls = (int16_t)(kClsS16 + wsel);
            q.kshift = 3;  // the packed kernels always run 8 rows per lane (trace layout and traceback follow kshift)
        }
        q.trace_off = 0;
        q.cigar_off = runs_off;
        runs_off += (int64_t)m + nn + 2;
        h->cells += (int64_t)m * nn;
        h->nmax = std::max(h->nmax, nn);
    }
    h->runs_scratch = runs_off;
    // Few packed problems (a legacy-dense call holds ~200 problems of 100 M cells; C3 a handful of large pairs) cannot fill
    // the 148 x 16 resident warps one or even four warps each: widen the teams of the largest problems until the batch
    // offers about that many warps (at least one band pair per warp).
    if (h->use_s16 && h->s16_max_wsel > 0) {
        int64_t warps = 0;
        std::vector<int32_t> packed;
        for (int p = 0; p < n; ++p) if (h->probs[(size_t)p].cls >= kClsS16) { warps += 1 << (h->probs[(size_t)p].cls - kClsS16); packed.push_back(p); }
        const int64_t target = (int64_t)h->sm_count * 16 * 3 / 4;
        if (!packed.empty() && warps < target) {
            std::sort(packed.begin(), packed.end(), [&](int32_t x, int32_t y) { return (int64_t)h->probs[x].m * h->probs[x].n > (int64_t)h->probs[y].m * h->probs[y].n; });
            for (int round = 0; round < 3 && warps < target; ++round)
                for (int32_t p : packed) {
                    SmxProblem &q = h->probs[(size_t)p];
                    const int wsel = q.cls - kClsS16, pairs = ((q.m + 255) / 256 + 1) / 2;
                    if (wsel >= h->s16_max_wsel || pairs < (2 << wsel) || (int64_t)q.m * q.n < (1 << 20)) continue;
                    q.cls = (int16_t)(q.cls + 1);
                    warps += 1 << wsel;
                    if (warps >= target) break;
                }
        }
    }
    // checkpointed traceback for the large packed problems: the forward pass runs score-only (1.7x faster) and the walk
    // recomputes the direction bits of the blocks it crosses (~14 % of the cells of a 2 500 x 4 000 problem); pays from
    // about 1 500 columns on (the recomputed share of a band pair is ~560 of its n + 95 steps) and from two band pairs
    // on (C2: 1 025 rows 409 ms per 20 000 reads in the forward phase, 769 rows 405, 513 rows 401)
    h->ck_cells = 0; h->ck_problems = 0;
    if (h->ckpt && h->use_s16 && h->use_s16h)
        for (int p = 0; p < n; ++p) {
            SmxProblem &q = h->probs[(size_t)p];
            if (q.cls >= kClsS16 && (q.mode & SMX_MODE_MASK) != SMX_MODE_SW && q.m >= h->ck_min_rows && q.n >= h->ck_min_cols) {
                q.mode |= SMX_MODE_CKPT;
                h->ck_cells += (int64_t)q.m * q.n; ++h->ck_problems;
            }
        }
    // concurrent teams: the packed forward phase (one-warp launch + team launches side by side) is timed as ONE interval
    for (int p = 0; p < n; ++p) {
        const SmxProblem &q = h->probs[(size_t)p];
        h->class_cells[(h->teams_concurrent && q.cls > kClsS16) ? kClsS16 : q.cls] += (int64_t)q.m * q.n;
    }
    // class-major order (largest kernel class first; counting sort), the large classes size-sorted inside so
    // that their longest problems start first; then greedy chunking under the trace budget.  The small classes
    // (86 % of the problems, 1 % of the cells) are left in input order.
    std::vector<int32_t> order((size_t)n);
    {
        int cnt[kNumFwdClasses + 1] = {0};
        for (int p = 0; p < n; ++p) cnt[kNumFwdClasses - 1 - h->probs[(size_t)p].cls + 1]++;
        for (int c = 0; c < kNumFwdClasses; ++c) cnt[c + 1] += cnt[c];
        int pos_c[kNumFwdClasses];
        for (int c = 0; c < kNumFwdClasses; ++c) pos_c[c] = cnt[c];
        for (int p = 0; p < n; ++p) order[(size_t)pos_c[kNumFwdClasses - 1 - h->probs[(size_t)p].cls]++] = p;
        for (int c = 0; c < kNumFwdClasses; ++c) {
            const int cls = kNumFwdClasses - 1 - c;
            if (cls < 6) continue;  // K <= 4 rows per lane: tiny problems
            std::stable_sort(order.begin() + cnt[c], order.begin() + cnt[c + 1], [&](int32_t x, int32_t y) {
                // packed classes: the local problems form the tail of the class range (they run a kernel of their own)
                // and before them the checkpointed ones: [direction bits][checkpointed][local]
                const int lx = cls < kClsS16 ? 0 : (h->probs[x].mode & SMX_MODE_MASK) == SMX_MODE_SW ? 2 : (h->probs[x].mode & SMX_MODE_CKPT) ? 1 : 0;
                const int ly = cls < kClsS16 ? 0 : (h->probs[y].mode & SMX_MODE_MASK) == SMX_MODE_SW ? 2 : (h->probs[y].mode & SMX_MODE_CKPT) ? 1 : 0;
                if (lx != ly) return lx < ly;
                return (int64_t)h->probs[x].m * h->probs[x].n > (int64_t)h->probs[y].m * h->probs[y].n;
            });
        }
    }
    h->chunks.clear();
    h->order_all = order;
    h->order_cls = order;  // already grouped by class inside every chunk (a chunk is a contiguous range of `order`)
    h->max_chunk_trace = 0;
    int pos = 0;
    while (pos < n) {
        Chunk c;
        c.first = pos;
        size_t bytes = 0;
        for (int cls = 0; cls < kNumFwdClasses; ++cls) { c.cls_first[cls] = pos; c.cls_count[cls] = 0; c.cls_local[cls] = 0; c.cls_ck[cls] = 0; }
        while (pos < n) {
            SmxProblem &q = h->probs[(size_t)order[pos]];
            size_t tb = (size_t)smx_region_bytes(q.m, q.n, 1 << q.kshift, q.mode);
            tb = ((tb + 255) & ~(size_t)255) + (h->guard ? 512 : 0);
            if (c.count > 0 && bytes + tb > (size_t)h->trace_budget) break;
            q.trace_off = (int64_t)bytes + (h->guard ? 256 : 0);
            bytes += tb;
            if (c.cls_count[q.cls] == 0) c.cls_first[q.cls] = pos;
            ++c.cls_count[q.cls];
            if (q.cls >= kClsS16 && (q.mode & SMX_MODE_MASK) == SMX_MODE_SW) ++c.cls_local[q.cls];
            if (q.mode & SMX_MODE_CKPT) ++c.cls_ck[q.cls];
            ++c.count; ++pos;
        }
        c.trace_bytes = bytes;
        h->max_chunk_trace = std::max(h->max_chunk_trace, bytes);
        // traceback order of the chunk: long walks first (one warp each, see smx_traceback), the rest one thread each
        {
            auto big = [&](int32_t p) { return h->probs[(size_t)p].m + h->probs[(size_t)p].n >= kTbBigWalk; };
            auto plain = [&](int32_t p) { return !(h->probs[(size_t)p].mode & SMX_MODE_CKPT); };
            auto ck0 = std::stable_partition(h->order_all.begin() + c.first, h->order_all.begin() + c.first + c.count, plain);
            c.n_ck = (int)((h->order_all.begin() + c.first + c.count) - ck0);
            auto mid = std::stable_partition(h->order_all.begin() + c.first, ck0, big);
            c.n_big = (int)(mid - (h->order_all.begin() + c.first));
        }
        h->chunks.push_back(c);
    }
    if (n == 0) { h->n = 0; return 0; }
    if (int rc = ensure(h, h->d_probs, sizeof(SmxProblem) * (size_t)n)) return rc;
    if (int rc = ensure(h, h->d_order_cls, sizeof(int32_t) * (size_t)n)) return rc;
    if (int rc = ensure(h, h->d_order_all, sizeof(int32_t) * (size_t)n)) return rc;
    if (int rc = ensure(h, h->d_results, sizeof(SmxResult) * (size_t)n)) return rc;
    if (int rc = ensure(h, h->d_trace, h->max_chunk_trace + 256)) return rc;
    if (int rc = ensure(h, h->d_runs, sizeof(uint32_t) * (size_t)h->runs_scratch)) return rc;
    if (int rc = ensure(h, h->d_dense_off, sizeof(int64_t) * ((size_t)n + 1))) return rc;
    if (int rc = ensure(h, h->d_tickets, sizeof(int32_t) * 3 * kNumFwdClasses * h->chunks.size())) return rc;
    if (int rc = h2d(h, PIN_PROBS, h->d_probs.p, h->probs.data(), sizeof(SmxProblem) * (size_t)n, h->stream_hi)) return rc;
    if (int rc = h2d(h, PIN_ORDER_CLS, h->d_order_cls.p, h->order_cls.data(), sizeof(int32_t) * (size_t)n, h->stream_hi)) return rc;
    if (int rc = h2d(h, PIN_ORDER_ALL, h->d_order_all.p, h->order_all.data(), sizeof(int32_t) * (size_t)n, h->stream_hi)) return rc;
    SMX_CUDA(h, smx_sync(h, h->stream_hi));
    h->n = n;
    return 0;
}

// Handles that share a GPU submit their forward-DP phase through this gate: at most `permits` DP phases are on the
// device at a time (FIFO between handles), while the short front-end kernels of the other handles cut in on their
// high-priority streams.  Without it all handles in flight submit their DP phases together, the GPU time-slices
// them, they all finish at the same moment, and every handle is in its host phase at once (measured: GPU idle 22 %).
struct DpGate {
    std::mutex mu;
    std::condition_variable cv;
    int in_flight = 0;
    unsigned long long next_ticket = 0, serving = 0;
    int permits()
    {
        static const int p = [] { const char *e = getenv("SMX_DP_INFLIGHT"); const int v = e ? atoi(e) : 3; return v < 1 ? 1 : v; }();
        return p;
    }
    void enter()
    {
        std::unique_lock<std::mutex> lk(mu);
        const unsigned long long my = next_ticket++;
        cv.wait(lk, [&] { return my == serving && in_flight < permits(); });
        ++serving;
        ++in_flight;
        cv.notify_all();
    }
    void leave()
    {
        { std::lock_guard<std::mutex> lk(mu); --in_flight; }
        cv.notify_all();
    }
};
static DpGate g_dp_gate[16];

template <int K, bool LOCAL, bool TEAM>
static int launch_forward(smx_handle *h, const Chunk &c, int cls, int chunk_idx)
{
    const int count = c.cls_count[cls];
    if (count == 0) return 0;
    const int threads = TEAM ? SMX_TEAM_WARPS * 32 : 128;
    int occ = 0;
    SMX_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, smx_dp_forward<K, LOCAL, TEAM>, threads, 0));
    if (occ < 1) occ = 1;
    const int per_block = TEAM ? 1 : threads / 32;  // problems in flight per CTA
    int blocks = std::min(h->sm_count * occ, (count + per_block - 1) / per_block);
    if (blocks < 1) blocks = 1;
    const size_t workers = TEAM ? (size_t)blocks : (size_t)blocks * (threads / 32);
    if (int rc = ensure(h, h->d_boundary, sizeof(int2) * 2 * (size_t)h->nmax * workers)) return rc;
    SmxFwdArgs a;
    a.pool = (const uint8_t *)h->pool.p;
    a.probs = (const SmxProblem *)h->d_probs.p;
    a.order = (const int32_t *)h->d_order_cls.p + c.cls_first[cls];
    a.n_order = count;
    a.ticket = (int32_t *)h->d_tickets.p + (size_t)chunk_idx * 3 * kNumFwdClasses + cls;
    a.trace = (uint8_t *)h->d_trace.p;
    a.boundary = (int2 *)h->d_boundary.p;
    a.nmax = h->nmax;
    a.results = (SmxResult *)h->d_results.p;
    a.sc = h->sc;
    a.one = 1u;
    a.max_tickets = INT32_MAX; a.n_slots = 0; a.slot_flags = nullptr;
    a.split_last = 0;
    smx_dp_forward<K, LOCAL, TEAM><<<blocks, threads, 0, h->stream>>>(a);
    SMX_CUDA(h, cudaGetLastError());
    h->last_launches++;
    return mark(h, cls);
}

template <int W>
static int launch_forward16(smx_handle *h, const Chunk &c, int cls, int chunk_idx, cudaStream_t stream = nullptr, bool do_mark = true, int sub_mask = 7)
{
    if (c.cls_count[cls] == 0) return 0;
    if (!stream) stream = h->stream;
    const bool TEAM = W > 1;
    const int threads = TEAM ? W * 32 : (h->use_s16h ? 32 * SMX16H_WPC : 128);
    // the global / semi-global problems of the class that keep direction bits, then its checkpointed and its local problems
    // (middle and tail of the class range, kernels of their own)
    for (int sub = 0; sub < 3; ++sub) {
        const bool local = sub == 2, ck = sub == 1;
        const int n_plain = c.cls_count[cls] - c.cls_local[cls] - c.cls_ck[cls];
        const int count = local ? c.cls_local[cls] : ck ? c.cls_ck[cls] : n_plain;
        if (count == 0 || !(sub_mask & (1 << sub))) continue;
        int occ = 0;
        void (*kern)(const SmxFwdArgs) = !h->use_s16h ? smx_dp_forward16<W> : local ? smx_dp_forward16h<W, true, false> : ck ? smx_dp_forward16h<W, false, true> : smx_dp_forward16h<W, false, false>;
        SMX_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, 0));
        if (occ < 1) occ = 1;
        const int per_block = TEAM ? 1 : threads / 32;
        // smx_dp16h.cuh is launched non-persistent (one problem per worker, boundary lines claimed from a slot pool)
        const bool persistent = !h->use_s16h || h->persistent_fwd;
        int blocks = (count + per_block - 1) / per_block;
        if (persistent) blocks = std::min(h->sm_count * occ, blocks);
        if (blocks < 1) blocks = 1;
        const size_t workers = !persistent ? (size_t)kLineSlots : TEAM ? (size_t)blocks : (size_t)blocks * (threads / 32);
        if (int rc = ensure(h, h->d_boundary, sizeof(int2) * 2 * (size_t)h->nmax * workers)) return rc;
        SmxFwdArgs a;
        a.pool = (const uint8_t *)h->pool.p;
        a.probs = (const SmxProblem *)h->d_probs.p;
        a.order = (const int32_t *)h->d_order_cls.p + c.cls_first[cls] + (local ? c.cls_count[cls] - c.cls_local[cls] : ck ? n_plain : 0);
        a.n_order = count;
        a.ticket = (int32_t *)h->d_tickets.p + ((size_t)chunk_idx * 3 + sub) * kNumFwdClasses + cls;
        a.trace = (uint8_t *)h->d_trace.p;
        a.boundary = (int2 *)h->d_boundary.p;
        a.nmax = h->nmax;
        a.results = (SmxResult *)h->d_results.p;
        a.sc = h->sc;
        a.one = 1u;
        a.max_tickets = persistent ? INT32_MAX : 1;
        a.n_slots = kLineSlots;
        a.slot_flags = persistent || ck ? nullptr : (int32_t *)h->d_slots.p;  // checkpointed problems write lines of their own
        a.split_last = h->split_last;
        kern<<<blocks, threads, 0, stream>>>(a);
        SMX_CUDA(h, cudaGetLastError());
        h->last_launches++;
    }
    return do_mark ? mark(h, cls) : 0;
}

// SMX_GUARD: one warp per problem checks the 2 x 256 poisoned bytes around its direction-bit region
__global__ void __launch_bounds__(128) smx_guard_check(const SmxProblem *probs, const int32_t *order, int n, const uint8_t *trace, unsigned long long *err)
{
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= n) return;
    const SmxProblem pr = probs[order[w]];
    const int64_t bytes = (smx_region_bytes(pr.m, pr.n, 1 << pr.kshift, pr.mode) + 255) & ~(int64_t)255;
    const uint32_t *before = reinterpret_cast<const uint32_t *>(trace + pr.trace_off - 256);
    const uint32_t *after = reinterpret_cast<const uint32_t *>(trace + pr.trace_off + bytes);
    int bad = 0;
    for (int x = lane; x < 64; x += 32) bad += (before[x] != 0xA5A5A5A5u) + (after[x] != 0xA5A5A5A5u);
    if (bad) atomicAdd(err, (unsigned long long)bad);
}

extern "C" int64_t smx_debug_guard_errors(const smx_handle *h) { return h ? h->guard_errors : -1; }

static int align_run_impl(smx_handle *h, int64_t *total_runs, DpGate *gate);

extern "C" int smx_align_run(smx_handle *h, int64_t *total_runs) { return align_run_impl(h, total_runs, nullptr); }

// gate != nullptr (pipeline): the launches wait for admission, and the gate is handed on as soon as the last forward
// kernel has finished -- the latency-bound tail of the traceback overlaps the next handle's forward launch
static int align_run_impl(smx_handle *h, int64_t *total_runs, DpGate *gate)
{
    if (!h) return -1;
    if (total_runs) *total_runs = 0;
    h->last_launches = 0;
    h->last_ms = 0.f;
    if (h->n == 0) { h->total_runs = 0; return 0; }
    SMX_CUDA(h, cudaSetDevice(h->device));
    // boundary lines must be sized before the timed region (no allocation between launches)
    {
        const size_t line_bytes = sizeof(int2) * 2 * (size_t)h->nmax * (size_t)std::max(h->sm_count * 16 * 4, kLineSlots);  // >= workers of any class
        if (int rc = ensure(h, h->d_boundary, line_bytes)) return rc;
    }
    struct GateGuard {  // leaves the gate on every exit path
        DpGate *g;
        ~GateGuard() { if (g) g->leave(); }
        void release() { if (g) { g->leave(); g = nullptr; } }
    } guard{gate};
    if (gate) gate->enter();
    SMX_CUDA(h, cudaEventRecord(h->ev0, h->stream));
    SMX_CUDA(h, cudaMemsetAsync(h->d_tickets.p, 0, sizeof(int32_t) * 3 * kNumFwdClasses * h->chunks.size(), h->stream));
    SMX_CUDA(h, cudaMemsetAsync(h->d_slots.p, 0, sizeof(int32_t) * kLineSlots, h->stream));  // an aborted launch must not leave lines claimed
    h->ev_used = 0;
    for (int c = 0; c < kNumSlots; ++c) { h->kernel_ms[c] = 0.0; h->kernel_launches[c] = 0; }
    if (int rc = mark(h, -1)) return rc;
    if (h->guard) {
        if (int rc = ensure(h, h->d_guard_err, 8)) return rc;
        SMX_CUDA(h, cudaMemsetAsync(h->d_guard_err.p, 0, 8, h->stream));
    }
    for (size_t ci = 0; ci < h->chunks.size(); ++ci) {
        const Chunk &c = h->chunks[ci];
        int rc = 0;
        if (h->guard) SMX_CUDA(h, cudaMemsetAsync(h->d_trace.p, 0xA5, c.trace_bytes, h->stream));
        // the checkpointed one-warp problems (the large ones) and the ones that keep direction bits run side by side as well
        const bool side_ck = h->ck_side && h->stream_team[3] && c.cls_ck[kClsS16] > 0 && c.cls_count[kClsS16] > c.cls_ck[kClsS16];
        if (h->teams_concurrent && ((c.cls_count[kClsS16 + 1] | c.cls_count[kClsS16 + 2] | c.cls_count[kClsS16 + 3]) || side_ck)) {
            // fork: the team launches (few CTAs, long) go first on their own streams and take their SM slots, the
            // one-warp-per-problem launch fills the rest; join before the next kernel of this stream
            SMX_CUDA(h, cudaEventRecord(h->ev_fork, h->stream));
            for (int t = 2; t >= 0; --t) {
                if (c.cls_count[kClsS16 + 1 + t] == 0) continue;
                SMX_CUDA(h, cudaStreamWaitEvent(h->stream_team[t], h->ev_fork, 0));
                if (t == 2) rc = launch_forward16<8>(h, c, kClsS16 + 3, (int)ci, h->stream_team[t], false);
                else if (t == 1) rc = launch_forward16<4>(h, c, kClsS16 + 2, (int)ci, h->stream_team[t], false);
                else rc = launch_forward16<2>(h, c, kClsS16 + 1, (int)ci, h->stream_team[t], false);
                if (rc) return rc;
                SMX_CUDA(h, cudaEventRecord(h->ev_join[t], h->stream_team[t]));
            }
            if (side_ck) {
                SMX_CUDA(h, cudaStreamWaitEvent(h->stream_team[3], h->ev_fork, 0));
                if ((rc = launch_forward16<1>(h, c, kClsS16, (int)ci, h->stream_team[3], false, 2))) return rc;
                SMX_CUDA(h, cudaEventRecord(h->ev_join[3], h->stream_team[3]));
            }
            if ((rc = launch_forward16<1>(h, c, kClsS16, (int)ci, h->stream, false, side_ck ? 5 : 7))) return rc;
            for (int t = 0; t < 3; ++t)
                if (c.cls_count[kClsS16 + 1 + t]) SMX_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_join[t], 0));
            if (side_ck) SMX_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_join[3], 0));
            if ((rc = mark(h, kClsS16))) return rc;
        } else {
            if ((rc = launch_forward16<8>(h, c, kClsS16 + 3, (int)ci))) return rc;
            if ((rc = launch_forward16<4>(h, c, kClsS16 + 2, (int)ci))) return rc;
            if ((rc = launch_forward16<2>(h, c, kClsS16 + 1, (int)ci))) return rc;
            if ((rc = launch_forward16<1>(h, c, kClsS16, (int)ci))) return rc;
        }
        if ((rc = launch_forward<8, false, true>(h, c, 8, (int)ci))) return rc;
        if ((rc = launch_forward<8, true, true>(h, c, 9, (int)ci))) return rc;
        if ((rc = launch_forward<8, false, false>(h, c, 6, (int)ci))) return rc;
        if ((rc = launch_forward<8, true, false>(h, c, 7, (int)ci))) return rc;
        if ((rc = launch_forward<4, false, false>(h, c, 4, (int)ci))) return rc;
        if ((rc = launch_forward<4, true, false>(h, c, 5, (int)ci))) return rc;
        if ((rc = launch_forward<2, false, false>(h, c, 2, (int)ci))) return rc;
        if ((rc = launch_forward<2, true, false>(h, c, 3, (int)ci))) return rc;
        if ((rc = launch_forward<1, false, false>(h, c, 0, (int)ci))) return rc;
        if ((rc = launch_forward<1, true, false>(h, c, 1, (int)ci))) return rc;
        if (gate && ci + 1 == h->chunks.size()) SMX_CUDA(h, cudaEventRecord(h->ev_fwd, h->stream));  // last forward kernel is in
        SmxTbArgs t;
        t.pool = (const uint8_t *)h->pool.p;
        t.probs = (const SmxProblem *)h->d_probs.p;
        t.order = (const int32_t *)h->d_order_all.p + c.first;
        t.n_order = c.count - c.n_ck;
        t.n_big = c.n_big;
        t.trace = (const uint8_t *)h->d_trace.p;
        t.results = (SmxResult *)h->d_results.p;
        t.runs = (uint32_t *)h->d_runs.p;
        t.sc = h->sc;
        t.split_last = h->use_s16h ? h->split_last : 0;
        const long long tb_threads = 32ll * c.n_big + (c.count - c.n_ck - c.n_big);
        // the recomputing walks of the checkpointed problems next to the walks that read direction bits (latency-bound)
        const bool side_tb = h->ck_side && h->teams_concurrent && h->stream_team[3] && c.n_ck > 0 && tb_threads > 0;
        cudaStream_t ck_stream = side_tb ? h->stream_team[3] : h->stream;
        if (side_tb) {
            SMX_CUDA(h, cudaEventRecord(h->ev_fork, h->stream));
            SMX_CUDA(h, cudaStreamWaitEvent(ck_stream, h->ev_fork, 0));
        }
        if (c.n_ck > 0) {
            SmxTbCkArgs k;
            k.pool = t.pool; k.probs = t.probs; k.trace = t.trace; k.results = t.results; k.runs = t.runs; k.sc = t.sc; k.split_last = t.split_last;
            k.order = (const int32_t *)h->d_order_all.p + c.first + (c.count - c.n_ck);
            k.n_order = c.n_ck;
            k.one = 1u;
            const size_t smem = sizeof(uint32_t) * SMX_TBCK_WARPS * SMX_TBCK_TILE_WORDS;
            static std::once_flag once[16];
            std::call_once(once[h->device & 15], [&] { cudaFuncSetAttribute(smx_traceback_ck, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); });
            smx_traceback_ck<<<(unsigned)((c.n_ck + SMX_TBCK_WARPS - 1) / SMX_TBCK_WARPS), 32 * SMX_TBCK_WARPS, smem, ck_stream>>>(k);
            SMX_CUDA(h, cudaGetLastError());
            h->last_launches++;
        }
        if (tb_threads > 0) {
            smx_traceback<<<(unsigned)((tb_threads + 127) / 128), 128, 0, h->stream>>>(t);
            SMX_CUDA(h, cudaGetLastError());
            h->last_launches++;
        }
        if (side_tb) {
            SMX_CUDA(h, cudaEventRecord(h->ev_join[3], ck_stream));
            SMX_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_join[3], 0));
        }
        if (int rc2 = mark(h, kSlotTraceback)) return rc2;
        if (h->guard) {
            smx_guard_check<<<(c.count + 3) / 4, 128, 0, h->stream>>>((const SmxProblem *)h->d_probs.p, (const int32_t *)h->d_order_all.p + c.first, c.count,
                                                                      (const uint8_t *)h->d_trace.p, (unsigned long long *)h->d_guard_err.p);
            SMX_CUDA(h, cudaGetLastError());
        }
    }
    SMX_CUDA(h, cudaEventRecord(h->ev1, h->stream));
    // results header to host: run counts -> dense offsets -> compaction
    h->h_results.resize((size_t)h->n);
    if (int rc = d2h(h, PIN_RESULTS, h->d_results.p, sizeof(SmxResult) * (size_t)h->n, h->stream)) return rc;
    if (gate) { SMX_CUDA(h, cudaEventSynchronize(h->ev_fwd)); guard.release(); }
    SMX_CUDA(h, smx_sync(h, h->stream));
    memcpy(h->h_results.data(), h->pin_stage[PIN_RESULTS].p, sizeof(SmxResult) * (size_t)h->n);
    if (h->guard) {
        unsigned long long e = 0;
        SMX_CUDA(h, cudaMemcpy(&e, h->d_guard_err.p, 8, cudaMemcpyDeviceToHost));
        h->guard_errors += (int64_t)e;
    }
    SMX_CUDA(h, cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1));
    h->h_dense_off.resize((size_t)h->n + 1);
    int64_t acc = 0;
    for (int p = 0; p < h->n; ++p) { h->h_dense_off[(size_t)p] = acc; acc += h->h_results[(size_t)p].n_runs; }
    h->h_dense_off[(size_t)h->n] = acc;
    h->total_runs = acc;
    if (int rc = ensure(h, h->d_dense, sizeof(uint32_t) * (size_t)std::max<int64_t>(acc, 1))) return rc;
    if (int rc = mark(h, -1, h->stream_hi)) return rc;  // the compaction interval starts here, not at the end of the traceback (host work lies between)
    if (int rc = h2d(h, PIN_DENSE_OFF, h->d_dense_off.p, h->h_dense_off.data(), sizeof(int64_t) * ((size_t)h->n + 1), h->stream_hi)) return rc;
    const int threads = 256, wpb = threads / 32;
    smx_compact_runs<<<(h->n + wpb - 1) / wpb, threads, 0, h->stream_hi>>>((const SmxProblem *)h->d_probs.p, (const SmxResult *)h->d_results.p,
                                                                        (const int64_t *)h->d_dense_off.p, (const uint32_t *)h->d_runs.p,
                                                                        (uint32_t *)h->d_dense.p, h->n);
    SMX_CUDA(h, cudaGetLastError());
    h->last_launches++;
    if (int rc = mark(h, kSlotCompact, h->stream_hi)) return rc;
    SMX_CUDA(h, smx_sync(h, h->stream_hi));
    for (size_t i = 1; i < h->ev_used; ++i) {
        if (h->ev_tag[i] < 0) continue;
        float ms = 0.f;
        // the compaction interval also spans the small header copy that precedes it
        if (cudaEventElapsedTime(&ms, h->ev_pool[i - 1], h->ev_pool[i]) == cudaSuccess) { h->kernel_ms[h->ev_tag[i]] += ms; h->kernel_launches[h->ev_tag[i]]++; }
    }
    if (total_runs) *total_runs = acc;
    return 0;
}

extern "C" int smx_align_fetch(smx_handle *h, int32_t *score, int32_t *end_query, int32_t *end_ref, int32_t *beg_query,
                               int32_t *beg_ref, int64_t *cigar_off, uint32_t *cigar_rle, int64_t cigar_capacity)
{
    if (!h) return -1;
    if (h->total_runs < 0) return fail(h, -1, "smx_align_fetch: no completed smx_align_run");
    if (cigar_capacity < h->total_runs) return fail(h, -1, "smx_align_fetch: cigar buffer too small (%lld < %lld)", (long long)cigar_capacity, (long long)h->total_runs);
    for (int p = 0; p < h->n; ++p) {
        const SmxResult &r = h->h_results[(size_t)p];
        if (score) score[p] = r.score;
        if (end_query) end_query[p] = r.end_q;
        if (end_ref) end_ref[p] = r.end_r;
        if (beg_query) beg_query[p] = r.beg_q;
        if (beg_ref) beg_ref[p] = r.beg_r;
    }
    if (cigar_off) memcpy(cigar_off, h->h_dense_off.data(), sizeof(int64_t) * ((size_t)h->n + 1));
    if (cigar_rle && h->total_runs > 0) {
        SMX_CUDA(h, cudaSetDevice(h->device));
        SMX_CUDA(h, cudaMemcpyAsync(cigar_rle, h->d_dense.p, sizeof(uint32_t) * (size_t)h->total_runs, cudaMemcpyDeviceToHost, h->stream_hi));
        SMX_CUDA(h, smx_sync(h, h->stream_hi));
    }
    return 0;
}

// diagnostic: raw copy of the packed direction-bit arena of the last chunk (layout: smx_dp.cuh)
extern "C" int smx_debug_copy_trace(smx_handle *h, uint8_t *dst, int64_t nbytes)
{
    if (!h || !dst) return -1;
    if ((size_t)nbytes > h->d_trace.cap) return fail(h, -1, "smx_debug_copy_trace: arena holds %zu bytes", h->d_trace.cap);
    SMX_CUDA(h, cudaSetDevice(h->device));
    SMX_CUDA(h, cudaMemcpy(dst, h->d_trace.p, (size_t)nbytes, cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int smx_align_kernel_ms(const smx_handle *h, double *ms16, int64_t *cells14)
{
    if (!h) return -1;
    if (ms16) for (int c = 0; c < kNumSlots; ++c) ms16[c] = h->kernel_ms[c];
    if (cells14) for (int c = 0; c < kNumFwdClasses; ++c) cells14[c] = h->class_cells[c];
    return 0;
}

extern "C" float smx_last_run_ms(const smx_handle *h) { return h ? h->last_ms : 0.f; }
extern "C" int64_t smx_align_cells(const smx_handle *h) { return h ? h->cells : 0; }
extern "C" int32_t smx_last_run_launches(const smx_handle *h) { return h ? h->last_launches : 0; }

// ------------------------------------------------------------------------------------------------
// A3: diagonal scan
// ------------------------------------------------------------------------------------------------
extern "C" int smx_scan_prepare(smx_handle *h, const int64_t *ref_off, const int32_t *ref_len, const int64_t *qry_off,
                                const int32_t *qry_len, int32_t n, int32_t topology, int64_t *counts_off)
{
    if (!h) return -1;
    h->scan_n = 0;
    if (n < 0 || (n > 0 && (!ref_off || !ref_len || !qry_off || !qry_len)) || !counts_off) return fail(h, -1, "smx_scan_prepare: NULL argument");
    if (topology != 0 && topology != 1) return fail(h, -1, "unknown topology_of_dna: %d (ref_query_alignment.py:368)", topology);
    if (!h->pool.p) return fail(h, -1, "smx_scan_prepare: no sequence pool uploaded");
    SMX_CUDA(h, cudaSetDevice(h->device));
    std::vector<SmxScanPair> pairs((size_t)n);
    std::vector<int32_t> tile_pair;
    int64_t total = 0, cells = 0;
    int32_t tiles = 0, nq_max = 1;
    for (int p = 0; p < n; ++p) {
        const int nr = ref_len[p], nq = qry_len[p];
        if (nr <= 0 || nq <= 0) return fail(h, -1, "scan pair %d: empty sequence", p);
        if (ref_off[p] < 0 || qry_off[p] < 0 || ref_off[p] + nr > h->pool_len || qry_off[p] + nq > h->pool_len)
            return fail(h, -1, "scan pair %d: sequence slice outside the uploaded pool", p);
        SmxScanPair &s = pairs[(size_t)p];
        s.ref_off = ref_off[p]; s.qry_off = qry_off[p]; s.nr = nr; s.nq = nq;
        s.max_k = topology == 0 ? nr : nr + nq - 1;
        s.counts_off = total;
        s.tile0 = tiles;
        counts_off[p] = total;
        total += s.max_k;
        cells += (int64_t)s.max_k * nq;
        const int t = (s.max_k + SMX_SCAN_TILE - 1) / SMX_SCAN_TILE;
        for (int x = 0; x < t; ++x) tile_pair.push_back(p);
        tiles += t;
        nq_max = std::max(nq_max, nq);
    }
    counts_off[n] = total;
    h->scan_total = total; h->scan_cells = cells; h->scan_tiles = tiles; h->scan_nq_max = nq_max; h->scan_topology = topology;
    if (n == 0) return 0;
    if (int rc = ensure(h, h->d_scan_pairs, sizeof(SmxScanPair) * (size_t)n)) return rc;
    if (int rc = ensure(h, h->d_scan_tile_pair, sizeof(int32_t) * (size_t)tiles)) return rc;
    if (int rc = ensure(h, h->d_counts, sizeof(int32_t) * (size_t)total)) return rc;
    if (int rc = h2d(h, PIN_SCAN_PAIRS, h->d_scan_pairs.p, pairs.data(), sizeof(SmxScanPair) * (size_t)n, h->stream_hi)) return rc;
    if (int rc = h2d(h, PIN_SCAN_TILES, h->d_scan_tile_pair.p, tile_pair.data(), sizeof(int32_t) * (size_t)tiles, h->stream_hi)) return rc;
    SMX_CUDA(h, smx_sync(h, h->stream_hi));
    h->scan_n = n;
    return 0;
}

extern "C" int smx_scan_run(smx_handle *h)
{
    if (!h) return -1;
    h->last_launches = 0;
    h->last_ms = 0.f;
    if (h->scan_n == 0) return 0;
    SMX_CUDA(h, cudaSetDevice(h->device));
    const int qwords4 = (((h->scan_nq_max + 31) / 32) + 3) & ~3;
    const int rwords = qwords4 + 34;
    const size_t smem = sizeof(uint2) * (size_t)(qwords4 + rwords) + sizeof(uint32_t) * (size_t)(rwords + 2);
    void (*scan_kern)(const SmxScanArgs) = h->scan_topology == 1 ? smx_scan_kernel<true> : smx_scan_kernel<false>;
    if (smem > 48 * 1024) SMX_CUDA(h, cudaFuncSetAttribute(scan_kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SmxScanArgs a;
    a.pool = (const uint8_t *)h->pool.p;
    a.pairs = (const SmxScanPair *)h->d_scan_pairs.p;
    a.tile_pair = (const int32_t *)h->d_scan_tile_pair.p;
    a.counts = (int32_t *)h->d_counts.p;
    a.topology = h->scan_topology;
    a.pool2 = (const uint2 *)h->pool2.p;
    a.special = (const uint8_t *)h->pool_special.p;
    SMX_CUDA(h, cudaEventRecord(h->ev0, h->stream_hi));
    scan_kern<<<h->scan_tiles, SMX_SCAN_THREADS, smem, h->stream_hi>>>(a);
    SMX_CUDA(h, cudaGetLastError());
    h->last_launches = 1;
    SMX_CUDA(h, cudaEventRecord(h->ev1, h->stream_hi));
    SMX_CUDA(h, smx_sync(h, h->stream_hi));
    SMX_CUDA(h, cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1));
    return 0;
}

extern "C" int smx_scan_fetch(smx_handle *h, int32_t *counts, int64_t counts_capacity)
{
    if (!h) return -1;
    if (!counts || counts_capacity < h->scan_total) return fail(h, -1, "smx_scan_fetch: counts buffer too small");
    if (h->scan_n == 0) return 0;
    SMX_CUDA(h, cudaSetDevice(h->device));
    SMX_CUDA(h, cudaMemcpyAsync(counts, h->d_counts.p, sizeof(int32_t) * (size_t)h->scan_total, cudaMemcpyDeviceToHost, h->stream_hi));
    SMX_CUDA(h, smx_sync(h, h->stream_hi));
    return 0;
}

extern "C" int64_t smx_scan_cells(const smx_handle *h) { return h ? h->scan_cells : 0; }

// ------------------------------------------------------------------------------------------------
// Row N3, first pass: MyMSA.generate_msa (modules/msa.py:882-982) on the device (smx_msa.cuh)
// ------------------------------------------------------------------------------------------------
extern "C" int smx_msa_run(smx_handle *h, const uint8_t *ref, int32_t ref_len, int32_t n_reads, const uint8_t *qry, const int8_t *qscore,
                           const int64_t *qry_off, const int32_t *qry_len, const uint32_t *cigar_rle, const int64_t *cigar_off, int64_t *n_cols)
{
    if (!h) return -1;
    h->msa_cols = -1;
    if (n_cols) *n_cols = 0;
    if (!ref || ref_len < 0 || n_reads < 0 || (n_reads > 0 && (!qry || !qscore || !qry_off || !qry_len || !cigar_rle || !cigar_off)))
        return fail(h, -1, "smx_msa_run: NULL argument");
    SMX_CUDA(h, cudaSetDevice(h->device));
    const int R = ref_len;
    // host pass over the runs: letters per read, and the gaps in which some block holds an 'I' before an 'S' (their column
    // order needs the serial simulation of the reference's loop)
    std::vector<int64_t> let_off((size_t)n_reads + 1, 0);
    std::vector<int32_t> complex_gaps;
    for (int i = 0; i < n_reads; ++i) {
        int64_t letters = 0;
        int refc = 0;
        bool seen_i = false, flagged = false;
        for (int64_t x = cigar_off[i]; x < cigar_off[i + 1]; ++x) {
            const uint32_t op = cigar_rle[x] & 15u;
            const int64_t len = cigar_rle[x] >> 4;
            letters += len;
            if (op == 1u) seen_i = true;
            else if (op == 4u) { if (seen_i && !flagged) { complex_gaps.push_back(refc); flagged = true; } }
            else { refc += (int)std::min<int64_t>(len, (int64_t)R + 1); seen_i = false; flagged = false; }
        }
        let_off[(size_t)i + 1] = let_off[(size_t)i] + letters;
    }
    std::sort(complex_gaps.begin(), complex_gaps.end());
    complex_gaps.erase(std::unique(complex_gaps.begin(), complex_gaps.end()), complex_gaps.end());
    std::vector<int32_t> type_off((size_t)R + 2, 0);
    std::vector<uint8_t> types;
    if (!complex_gaps.empty()) {
        // blocks of every read in the complex gaps, then the reference's loop on them (msa.py:905-969): an 'S' column while any
        // read stands at 'S', else an 'I' column while any read stands at 'I'
        std::vector<std::vector<std::string>> blocks(complex_gaps.size());
        for (int i = 0; i < n_reads; ++i) {
            int refc = 0;
            std::string cur;
            auto close = [&](int gap) {
                if (cur.empty()) return;
                auto it = std::lower_bound(complex_gaps.begin(), complex_gaps.end(), gap);
                if (it != complex_gaps.end() && *it == gap) blocks[(size_t)(it - complex_gaps.begin())].push_back(cur);
                cur.clear();
            };
            for (int64_t x = cigar_off[i]; x < cigar_off[i + 1]; ++x) {
                const uint32_t op = cigar_rle[x] & 15u;
                const int64_t len = cigar_rle[x] >> 4;
                if (op == 1u || op == 4u) cur.append((size_t)len, op == 1u ? 'I' : 'S');
                else { close(refc); refc += (int)std::min<int64_t>(len, (int64_t)R + 1); }
            }
            close(refc);
        }
        std::vector<std::string> seq(complex_gaps.size());
        for (size_t g = 0; g < complex_gaps.size(); ++g) {
            std::vector<size_t> p(blocks[g].size(), 0);
            for (;;) {
                bool any_s = false, any_i = false;
                for (size_t b = 0; b < blocks[g].size(); ++b)
                    if (p[b] < blocks[g][b].size()) { if (blocks[g][b][p[b]] == 'S') any_s = true; else any_i = true; }
                if (!any_s && !any_i) break;
                const char T = any_s ? 'S' : 'I';
                for (size_t b = 0; b < blocks[g].size(); ++b) if (p[b] < blocks[g][b].size() && blocks[g][b][p[b]] == T) ++p[b];
                seq[g].push_back(T);
            }
        }
        size_t g = 0;
        for (int r = 0; r <= R; ++r) {
            type_off[(size_t)r] = (int32_t)types.size();
            if (g < complex_gaps.size() && complex_gaps[g] == r) { types.insert(types.end(), seq[g].begin(), seq[g].end()); ++g; }
        }
        type_off[(size_t)R + 1] = (int32_t)types.size();
    }
    // device input block: ref | qry | qscore | qry_off | qry_len | runs | cigar_off | let_off
    const int64_t n_q = n_reads ? qry_off[n_reads - 1] + qry_len[n_reads - 1] : 0, n_runs = n_reads ? cigar_off[n_reads] : 0;
    auto up8 = [](size_t x) { return (x + 7) & ~(size_t)7; };
    size_t o = 0;
    const size_t o_ref = o; o += up8((size_t)R + 1);
    const size_t o_qry = o; o += up8((size_t)n_q + 1);
    const size_t o_qs = o; o += up8((size_t)n_q + 1);
    const size_t o_qoff = o; o += 8 * ((size_t)n_reads + 1);
    const size_t o_qlen = o; o += up8(4 * ((size_t)n_reads + 1));
    const size_t o_runs = o; o += up8(4 * ((size_t)n_runs + 1));
    const size_t o_coff = o; o += 8 * ((size_t)n_reads + 1);
    const size_t o_loff = o; o += 8 * ((size_t)n_reads + 1);
    if (int rc = ensure(h, h->d_msa_in, o)) return rc;
    char *din = (char *)h->d_msa_in.p;
    cudaStream_t st = h->stream_hi;
    if (R) SMX_CUDA(h, cudaMemcpyAsync(din + o_ref, ref, (size_t)R, cudaMemcpyHostToDevice, st));
    if (n_reads) {
        SMX_CUDA(h, cudaMemcpyAsync(din + o_qry, qry, (size_t)n_q, cudaMemcpyHostToDevice, st));
        SMX_CUDA(h, cudaMemcpyAsync(din + o_qs, qscore, (size_t)n_q, cudaMemcpyHostToDevice, st));
        SMX_CUDA(h, cudaMemcpyAsync(din + o_qoff, qry_off, 8 * (size_t)n_reads, cudaMemcpyHostToDevice, st));
        SMX_CUDA(h, cudaMemcpyAsync(din + o_qlen, qry_len, 4 * (size_t)n_reads, cudaMemcpyHostToDevice, st));
        SMX_CUDA(h, cudaMemcpyAsync(din + o_runs, cigar_rle, 4 * (size_t)n_runs, cudaMemcpyHostToDevice, st));
        SMX_CUDA(h, cudaMemcpyAsync(din + o_coff, cigar_off, 8 * ((size_t)n_reads + 1), cudaMemcpyHostToDevice, st));
    }
    SMX_CUDA(h, cudaMemcpyAsync(din + o_loff, let_off.data(), 8 * ((size_t)n_reads + 1), cudaMemcpyHostToDevice, st));
    // work block: let | pos | qcnt | n_s | n_i | err | col_off | type_off | types
    const size_t per = (size_t)R + 1;
    size_t w = 0;
    const size_t w_let = w; w += up8((size_t)let_off[(size_t)n_reads] + 1);
    const size_t w_pos = w; w += up8(4 * per * (size_t)std::max(n_reads, 1));
    const size_t w_qc = w; w += up8(4 * per * (size_t)std::max(n_reads, 1));
    const size_t w_ns = w; w += up8(4 * per);
    const size_t w_ni = w; w += up8(4 * per);
    const size_t w_err = w; w += up8(4 * ((size_t)n_reads + 1));
    const size_t w_col = w; w += 8 * (per + 1);
    const size_t w_toff = w; w += up8(4 * (per + 1));
    const size_t w_types = w; w += up8(types.size() + 1);
    if (int rc = ensure(h, h->d_msa_work, w)) return rc;
    char *dw = (char *)h->d_msa_work.p;
    SmxMsaArgs a;
    memset(&a, 0, sizeof a);
    a.ref = (const uint8_t *)(din + o_ref); a.R = R; a.n_reads = n_reads;
    a.qry = (const uint8_t *)(din + o_qry); a.qscore = (const int8_t *)(din + o_qs);
    a.qry_off = (const int64_t *)(din + o_qoff); a.qry_len = (const int32_t *)(din + o_qlen);
    a.cigar_rle = (const uint32_t *)(din + o_runs); a.cigar_off = (const int64_t *)(din + o_coff);
    a.let = (uint8_t *)(dw + w_let); a.let_off = (const int64_t *)(din + o_loff);
    a.pos = (int32_t *)(dw + w_pos); a.qcnt = (int32_t *)(dw + w_qc);
    a.n_s = (int32_t *)(dw + w_ns); a.n_i = (int32_t *)(dw + w_ni); a.err = (int32_t *)(dw + w_err);
    a.col_off = (const int64_t *)(dw + w_col); a.type_off = (const int32_t *)(dw + w_toff); a.types = (const uint8_t *)(dw + w_types);
    std::vector<int32_t> ns(per, 0), ni(per, 0), err((size_t)n_reads + 1, 0);
    if (n_reads > 0) {
        smx_msa_expand<<<(n_reads + 3) / 4, 128, 0, st>>>(a);
        SMX_CUDA(h, cudaGetLastError());
        // a CIGAR that does not consume exactly the reference leaves position entries unset: checked before anything reads them
        SMX_CUDA(h, cudaMemcpyAsync(err.data(), a.err, 4 * (size_t)n_reads, cudaMemcpyDeviceToHost, st));
        SMX_CUDA(h, smx_sync(h, st));
        for (int i = 0; i < n_reads; ++i)
            if (err[(size_t)i]) return fail(h, -6, "smx_msa_run: the CIGAR of read %d does not consume exactly the reference (%d) and its query (%d) "
                                                   "(generate_msa's final assertions, msa.py:913-916)", i, R, qry_len[i]);
        smx_msa_gaps<<<(unsigned)((per + 3) / 4), 128, 0, st>>>(a);
        SMX_CUDA(h, cudaGetLastError());
        SMX_CUDA(h, cudaMemcpyAsync(ns.data(), a.n_s, 4 * per, cudaMemcpyDeviceToHost, st));
        SMX_CUDA(h, cudaMemcpyAsync(ni.data(), a.n_i, 4 * per, cudaMemcpyDeviceToHost, st));
        SMX_CUDA(h, smx_sync(h, st));
    }
    // columns: gap r holds its extra columns, then (r < R) the reference column
    std::vector<int64_t> col_off(per + 1, 0);
    for (size_t r = 0; r < per; ++r) {
        const int64_t extra = type_off[r + 1] > type_off[r] ? type_off[r + 1] - type_off[r] : (int64_t)ns[r] + ni[r];
        col_off[r + 1] = col_off[r] + extra + (r < (size_t)R ? 1 : 0);
    }
    const int64_t C = col_off[per];
    SMX_CUDA(h, cudaMemcpyAsync(dw + w_col, col_off.data(), 8 * (per + 1), cudaMemcpyHostToDevice, st));
    SMX_CUDA(h, cudaMemcpyAsync(dw + w_toff, type_off.data(), 4 * (per + 1), cudaMemcpyHostToDevice, st));
    if (!types.empty()) SMX_CUDA(h, cudaMemcpyAsync(dw + w_types, types.data(), types.size(), cudaMemcpyHostToDevice, st));
    const size_t rows = (size_t)n_reads;
    if (int rc = ensure(h, h->d_msa_out, (size_t)C * (1 + 3 * rows) + 64)) return rc;
    char *dout = (char *)h->d_msa_out.p;
    a.n_cols = C;
    a.out_ref = (uint8_t *)dout; a.out_seq = (uint8_t *)dout + (size_t)C; a.out_cig = a.out_seq + (size_t)C * rows; a.out_qs = (int8_t *)(a.out_cig + (size_t)C * rows);
    if (C > 0) {
        dim3 grid((unsigned)((per + 255) / 256), (unsigned)std::min<int64_t>((int64_t)n_reads + 1, 32768));
        smx_msa_fill<<<grid, 256, 0, st>>>(a);
        SMX_CUDA(h, cudaGetLastError());
    }
    SMX_CUDA(h, smx_sync(h, st));
    h->msa_cols = C; h->msa_reads = n_reads;
    if (n_cols) *n_cols = C;
    return 0;
}

extern "C" int smx_msa_fetch(smx_handle *h, uint8_t *ref_aligned, uint8_t *seq_aligned, int8_t *qs_aligned, uint8_t *cigar_aligned)
{
    if (!h) return -1;
    if (h->msa_cols < 0) return fail(h, -1, "smx_msa_fetch: no completed smx_msa_run");
    SMX_CUDA(h, cudaSetDevice(h->device));
    const size_t C = (size_t)h->msa_cols, rows = (size_t)h->msa_reads;
    const char *d = (const char *)h->d_msa_out.p;
    cudaStream_t st = h->stream_hi;
    if (C == 0) return 0;
    if (ref_aligned) SMX_CUDA(h, cudaMemcpyAsync(ref_aligned, d, C, cudaMemcpyDeviceToHost, st));
    if (rows) {
        if (seq_aligned) SMX_CUDA(h, cudaMemcpyAsync(seq_aligned, d + C, C * rows, cudaMemcpyDeviceToHost, st));
        if (cigar_aligned) SMX_CUDA(h, cudaMemcpyAsync(cigar_aligned, d + C + C * rows, C * rows, cudaMemcpyDeviceToHost, st));
        if (qs_aligned) SMX_CUDA(h, cudaMemcpyAsync(qs_aligned, d + C + 2 * C * rows, C * rows, cudaMemcpyDeviceToHost, st));
    }
    SMX_CUDA(h, smx_sync(h, st));
    return 0;
}

// Row N3, second half: the Bayesian consensus of MyMSA.calculate_consensus_core (msa.py:1749-1812) for every column of
// a multiple alignment; see smx_consensus.cuh.  seq / qs / cig == NULL: the matrices of the last smx_msa_run, still on
// the device.  P25[B * 5 + called] = P_base_calling_given_true_refseq_dict["B_called"], pdf[(B * 5 + called) * 93 + k] =
// pdf_core["B_called"][k] (bases in the order "ATCG-"); pn_with / pn_without [n_classes][5] = P_N_dict of the column's
// reference letter (col_class[c]) with and without prior.  p_with / p_without [n_cols][5] = p_list per column.
extern "C" int smx_consensus_run(smx_handle *h, const uint8_t *seq, const int8_t *qs, const uint8_t *cig, int32_t n_rows, int64_t n_cols,
                                 const double *P25, const double *pdf, const int32_t *col_class, const double *pn_with, const double *pn_without,
                                 int32_t n_classes, double *p_with, double *p_without, int32_t *n_events)
{
    if (!h) return -1;
    if (n_rows < 0 || n_cols < 0 || n_classes <= 0 || !P25 || !pdf || !pn_with || !pn_without || (n_cols > 0 && (!col_class || !p_with || !p_without || !n_events)))
        return fail(h, -1, "smx_consensus_run: NULL or empty argument");
    if ((seq == nullptr) != (qs == nullptr) || (seq == nullptr) != (cig == nullptr)) return fail(h, -1, "smx_consensus_run: pass all three matrices or none");
    if (n_cols == 0) return 0;
    for (int64_t c = 0; c < n_cols; ++c) if (col_class[c] < 0 || col_class[c] >= n_classes) return fail(h, -1, "smx_consensus_run: column %lld: class out of range", (long long)c);
    SMX_CUDA(h, cudaSetDevice(h->device));
    cudaStream_t st = h->stream_hi;
    const size_t C = (size_t)n_cols, R = (size_t)n_rows;
    const uint8_t *d_seq, *d_cig;
    const int8_t *d_qs;
    if (!seq) {
        if (h->msa_cols != n_cols || h->msa_reads != n_rows) return fail(h, -1, "smx_consensus_run: no smx_msa_run of %d reads x %lld columns on this handle", n_rows, (long long)n_cols);
        const uint8_t *d = (const uint8_t *)h->d_msa_out.p;   // layout of smx_msa_run: ref, seq, cig, qs
        d_seq = d + C; d_cig = d_seq + C * R; d_qs = (const int8_t *)(d_cig + C * R);
    }
    // factor table and start values in native long double (smx_consensus_host.hpp)
    const int NT = 25 * SMX_CONS_BASES * SMX_CONS_QIDX;
    std::vector<uint64_t> tm, im;
    std::vector<int32_t> te, ie;
    smx_cons_build_table(P25, pdf, tm, te);
    smx_cons_build_init(pn_with, pn_without, n_classes, im, ie);
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t o_tm = 0, o_te = al(o_tm + 8 * (size_t)NT), o_im = al(o_te + 4 * (size_t)NT), o_ie = al(o_im + 8 * im.size()), o_cls = al(o_ie + 4 * ie.size()),
                 o_om = al(o_cls + 4 * C), o_oe = al(o_om + 8 * 50 * C), o_ev = al(o_oe + 4 * 50 * C), o_err = al(o_ev + 4 * C),
                 o_seq = al(o_err + 64), o_cig = al(o_seq + (seq ? C * R : 0)), o_qs = al(o_cig + (seq ? C * R : 0)), total = al(o_qs + (seq ? C * R : 0)) + 256;
    if (int rc = ensure(h, h->d_cons, total)) return rc;
    char *d = (char *)h->d_cons.p;
    SMX_CUDA(h, cudaMemcpyAsync(d + o_tm, tm.data(), 8 * (size_t)NT, cudaMemcpyHostToDevice, st));
    SMX_CUDA(h, cudaMemcpyAsync(d + o_te, te.data(), 4 * (size_t)NT, cudaMemcpyHostToDevice, st));
    SMX_CUDA(h, cudaMemcpyAsync(d + o_im, im.data(), 8 * im.size(), cudaMemcpyHostToDevice, st));
    SMX_CUDA(h, cudaMemcpyAsync(d + o_ie, ie.data(), 4 * ie.size(), cudaMemcpyHostToDevice, st));
    SMX_CUDA(h, cudaMemcpyAsync(d + o_cls, col_class, 4 * C, cudaMemcpyHostToDevice, st));
    SMX_CUDA(h, cudaMemsetAsync(d + o_err, 0, 64, st));
    if (seq) {
        if (R) {
            SMX_CUDA(h, cudaMemcpyAsync(d + o_seq, seq, C * R, cudaMemcpyHostToDevice, st));
            SMX_CUDA(h, cudaMemcpyAsync(d + o_cig, cig, C * R, cudaMemcpyHostToDevice, st));
            SMX_CUDA(h, cudaMemcpyAsync(d + o_qs, qs, C * R, cudaMemcpyHostToDevice, st));
        }
        d_seq = (const uint8_t *)(d + o_seq); d_cig = (const uint8_t *)(d + o_cig); d_qs = (const int8_t *)(d + o_qs);
    }
    SmxConsArgs a;
    a.seq = d_seq; a.qs = d_qs; a.cig = d_cig; a.rows = n_rows; a.cols = n_cols;
    a.tab_m = (const uint64_t *)(d + o_tm); a.tab_e = (const int32_t *)(d + o_te);
    a.col_class = (const int32_t *)(d + o_cls);
    a.init_m = (const uint64_t *)(d + o_im); a.init_e = (const int32_t *)(d + o_ie); a.n_classes = n_classes;
    a.out_m = (uint64_t *)(d + o_om); a.out_e = (int32_t *)(d + o_oe); a.n_events = (int32_t *)(d + o_ev); a.err = (int32_t *)(d + o_err);
    smx_consensus_kernel<<<dim3((unsigned)((C + 127) / 128), 25), 128, 0, st>>>(a);
    SMX_CUDA(h, cudaGetLastError());
    std::vector<uint64_t> om(50 * C);
    std::vector<int32_t> oe(50 * C);
    int32_t err = 0;
    SMX_CUDA(h, cudaMemcpyAsync(om.data(), d + o_om, 8 * 50 * C, cudaMemcpyDeviceToHost, st));
    SMX_CUDA(h, cudaMemcpyAsync(oe.data(), d + o_oe, 4 * 50 * C, cudaMemcpyDeviceToHost, st));
    SMX_CUDA(h, cudaMemcpyAsync(n_events, d + o_ev, 4 * C, cudaMemcpyDeviceToHost, st));
    SMX_CUDA(h, cudaMemcpyAsync(&err, d + o_err, 4, cudaMemcpyDeviceToHost, st));
    SMX_CUDA(h, smx_sync(h, st));
    if (err) return fail(h, -7, "smx_consensus_run: a voting read holds a character outside ATCG- or a q-score above 92 (the reference reads out of bounds there, alignment_functions.pyx:113)");
    smx_cons_finish(om.data(), oe.data(), C, p_with, p_without);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Row N2: assignment on the device, .ir result blocks from the arrays (smx_outputs.cuh)
// ------------------------------------------------------------------------------------------------
extern "C" int smx_assign(smx_handle *h, const int32_t *score, const uint8_t *status, const int32_t *ref_len, int32_t n_reads, int32_t n_refs,
                          double score_threshold, int32_t *assignment, double *normalized, double *summary)
{
    if (!h) return -1;
    if (n_reads < 0 || n_refs <= 0 || (n_reads > 0 && (!score || !status || !assignment || !normalized || !summary)) || !ref_len)
        return fail(h, -1, "smx_assign: NULL or empty argument");
    for (int r = 0; r < n_refs; ++r) if (ref_len[r] <= 0) return fail(h, -1, "smx_assign: reference %d has no length", r);
    if (n_reads == 0) return 0;
    SMX_CUDA(h, cudaSetDevice(h->device));
    const size_t n2 = 2 * (size_t)n_refs, cells = (size_t)n_reads * n2;
    const size_t o_score = 0, o_norm = (cells * 4 + 7) & ~(size_t)7, o_sum = o_norm + cells * 8, o_asg = o_sum + (size_t)n_reads * n_refs * 8,
                 o_len = o_asg + (size_t)n_reads * 12, o_status = o_len + (size_t)n_refs * 4, total = o_status + cells;
    DevBuf buf;
    if (int rc = ensure(h, buf, total)) return rc;
    char *d = (char *)buf.p;
    cudaStream_t s = h->stream_hi;
    SMX_CUDA(h, cudaMemcpyAsync(d + o_score, score, cells * 4, cudaMemcpyHostToDevice, s));
    SMX_CUDA(h, cudaMemcpyAsync(d + o_status, status, cells, cudaMemcpyHostToDevice, s));
    SMX_CUDA(h, cudaMemcpyAsync(d + o_len, ref_len, (size_t)n_refs * 4, cudaMemcpyHostToDevice, s));
    SmxAssignArgs a;
    a.score = (const int32_t *)(d + o_score); a.status = (const uint8_t *)(d + o_status); a.ref_len = (const int32_t *)(d + o_len);
    a.normalized = (double *)(d + o_norm); a.summary = (double *)(d + o_sum); a.assignment = (int32_t *)(d + o_asg);
    a.n_reads = n_reads; a.n_refs = n_refs; a.threshold = score_threshold;
    smx_assign_kernel<<<(n_reads + 127) / 128, 128, 0, s>>>(a);
    SMX_CUDA(h, cudaGetLastError());
    SMX_CUDA(h, cudaMemcpyAsync(normalized, d + o_norm, cells * 8, cudaMemcpyDeviceToHost, s));
    SMX_CUDA(h, cudaMemcpyAsync(summary, d + o_sum, (size_t)n_reads * n_refs * 8, cudaMemcpyDeviceToHost, s));
    SMX_CUDA(h, cudaMemcpyAsync(assignment, d + o_asg, (size_t)n_reads * 12, cudaMemcpyDeviceToHost, s));
    const cudaError_t e = smx_sync(h, s);
    release(buf);
    SMX_CUDA(h, e);
    return 0;
}

// out == NULL: returns the number of bytes the blocks result<first_index> .. result<first_index + n - 1> take;
// otherwise writes them (capacity checked) and returns the bytes written; negative on error.  No handle: pure host code.
extern "C" int64_t smx_ir_format(const int32_t *score, const int32_t *new_query_seq_offset, const uint8_t *status, const int64_t *cigar_off,
                                 const uint32_t *cigar_rle, int64_t n, int64_t first_index, int32_t n_threads, char *out, int64_t capacity)
{
    using namespace smx_out;
    if (n < 0 || (n > 0 && (!score || !new_query_seq_offset || !status || !cigar_off || (!cigar_rle && cigar_off[n] > cigar_off[0])))) return -1;
    if (n == 0) return 0;
    const int nt = (int)std::max<int64_t>(1, std::min<int64_t>(n_threads > 0 ? n_threads : 4, (n + 4095) / 4096));
    std::vector<int64_t> part((size_t)nt + 1, 0);
    auto range = [&](int t, int64_t &lo, int64_t &hi) { lo = n * t / nt; hi = n * (t + 1) / nt; };
    auto each = [&](auto &&fn) {
        if (nt == 1) { fn(0); return; }
        std::vector<std::thread> th;
        for (int t = 0; t < nt; ++t) th.emplace_back(fn, t);
        for (auto &x : th) x.join();
    };
    each([&](int t) {
        int64_t lo, hi, acc = 0;
        range(t, lo, hi);
        for (int64_t i = lo; i < hi; ++i)
            acc += block_len(first_index + i, cigar_rle + cigar_off[i], cigar_off[i + 1] - cigar_off[i], status[i] != 0, score[i], new_query_seq_offset[i]);
        part[(size_t)t + 1] = acc;
    });
    for (int t = 0; t < nt; ++t) part[(size_t)t + 1] += part[(size_t)t];
    const int64_t total = part[(size_t)nt];
    if (!out) return total;
    if (capacity < total) return -2;
    each([&](int t) {
        int64_t lo, hi;
        range(t, lo, hi);
        char *p = out + part[(size_t)t];
        for (int64_t i = lo; i < hi; ++i)
            p = block_put(p, first_index + i, cigar_rle + cigar_off[i], cigar_off[i + 1] - cigar_off[i], status[i] != 0, score[i], new_query_seq_offset[i]);
    });
    return total;
}

// ------------------------------------------------------------------------------------------------
// int32 / DPX issue-rate microbenchmark (roofline denominator)
// ------------------------------------------------------------------------------------------------
template <int WHICH>
__global__ void __launch_bounds__(256) smx_int_peak_kernel(int *out, int iters, int seed)
{
    // 8 independent chains per thread so that the ALU latency (4 cycles) is always covered
    int v[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) v[c] = seed + threadIdx.x * 8 + c;
    const int b = seed | 1, d = seed ^ 0x55;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                if (WHICH == 0) v[c] = __viaddmax_s32(v[c], b, d);
                else if (WHICH == 1) v[c] = __vimax3_s32(v[c], b + u, d);
                else v[c] = v[c] + (v[c] ^ b);  // IADD + LOP3 pair counted as 2 ops
            }
        }
    }
    int s = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) s ^= v[c];
    if (s == 0x7fffffff) out[0] = s;
}

extern "C" int smx_measure_int_peak(smx_handle *h, int32_t iters, double *out_gops)
{
    if (!h || !out_gops) return -1;
    SMX_CUDA(h, cudaSetDevice(h->device));
    if (int rc = ensure(h, h->d_tickets, 64)) return rc;
    const int blocks = h->sm_count * 8, threads = 256;
    for (int which = 0; which < 3; ++which) {
        float best = 1e30f;
        for (int rep = 0; rep < 3; ++rep) {
            SMX_CUDA(h, cudaEventRecord(h->ev0, h->stream_hi));
            if (which == 0) smx_int_peak_kernel<0><<<blocks, threads, 0, h->stream_hi>>>((int *)h->d_tickets.p, iters, 12345 + rep);
            else if (which == 1) smx_int_peak_kernel<1><<<blocks, threads, 0, h->stream_hi>>>((int *)h->d_tickets.p, iters, 12345 + rep);
            else smx_int_peak_kernel<2><<<blocks, threads, 0, h->stream_hi>>>((int *)h->d_tickets.p, iters, 12345 + rep);
            SMX_CUDA(h, cudaGetLastError());
            SMX_CUDA(h, cudaEventRecord(h->ev1, h->stream_hi));
            SMX_CUDA(h, smx_sync(h, h->stream_hi));
            float ms = 0.f;
            SMX_CUDA(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));
            best = std::min(best, ms);
        }
        const double ops = (double)blocks * threads * (double)iters * 16.0 * 8.0 * (which == 2 ? 2.0 : 1.0);
        out_gops[which] = ops / (best * 1e-3) / 1e9;
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// A1: the whole alignment stage (scan -> select -> chain -> DP -> stitch)
// ------------------------------------------------------------------------------------------------
#include "smx_pipeline.cuh"
#include "smx_chain.cuh"
#include <cstdlib>

namespace {

const int kRegionCap = SMX_SEL_REGION_CAP;   // device slots per pair-strand (overflow -> host re-decision)
const int64_t kCountsBudget = 400ll << 20;  // int32 diagonal counts kept on the device per chunk
// chain search: classes with fewer pair-strands than this many CTAs per SM spread the origins of a pair-strand over
// several CTAs (a launch of few long CTAs is bound by its slowest CTA: 3.9 ms for 195 pair-strands of ~100 regions)
const int kChainCtaTarget = []{ const char *e = getenv("SMX_CHAIN_CTAS"); const int v = e ? atoi(e) : 32; return v < 1 ? 1 : v; }();

// complement table, built once (thread-safe magic static: several handles call smx_pipeline_run concurrently) and
// never rewritten after publication
struct CompTable {
    uint8_t t[256];
    CompTable()
    {
        for (int i = 0; i < 256; ++i) t[i] = (uint8_t)i;
        // Bio.Seq.reverse_complement (my_classes.py:306-310): IUPAC-aware
        const char *a = "ACGTMRWSYKVHDBN", *b = "TGCAKYWSRMBDHVN";
        for (int i = 0; a[i]; ++i) t[(uint8_t)a[i]] = (uint8_t)b[i];
    }
};
const uint8_t *comp_table()
{
    static const CompTable tab;
    return tab.t;
}

}  // namespace

static SmxPipelineState *pipeline_state(smx_handle *h);

// SMX_PHASE_LOG=1: one stderr line per GPU phase of a pipeline call ("PH handle label t0 t1", steady-clock seconds)
static double cpu_now_s()
{
    timespec ts;
    clock_gettime(CLOCK_PROCESS_CPUTIME_ID, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

struct PhaseLog {
    bool on = getenv("SMX_PHASE_LOG") != nullptr;
    double cpu_last = cpu_now_s();
    // process CPU seconds since the previous stage boundary (meaningful with one batch in flight)
    void cpu(const void *h, const char *label)
    {
        if (!on) return;
        const double c = cpu_now_s();
        char buf[160];
        snprintf(buf, sizeof buf, "CPU %p %s %.6f\n", h, label, c - cpu_last);
        lines.emplace_back(buf);
        cpu_last = c;
    }
    std::vector<std::string> lines;
    void add(const void *h, const char *label, double t0, double t1)
    {
        if (!on) return;
        char buf[160];
        snprintf(buf, sizeof buf, "PH %p %s %.6f %.6f\n", h, label, t0, t1);
        lines.emplace_back(buf);
    }
    ~PhaseLog() { for (const std::string &l : lines) fputs(l.c_str(), stderr); }
};

extern "C" int smx_pipeline_run(smx_handle *h, const uint8_t *ref_seqs, const int64_t *ref_off, const int32_t *ref_len, int32_t n_refs,
                                const double *ref_sigma, const uint8_t *qry_seqs, const int64_t *qry_off, const int32_t *qry_len,
                                int32_t n_qrys, const int32_t *pair_ref, const int32_t *pair_qry, int32_t n_pairs, int32_t open,
                                int32_t extend, int32_t match, int32_t mismatch, int32_t topology, int32_t omit_too_long,
                                int32_t n_threads, int64_t *total_runs)
{
    using namespace smx;
    if (!h) return -1;
    SmxPipelineState &st = *pipeline_state(h);
    st = SmxPipelineState();
    if (!ref_seqs || !ref_off || !ref_len || !ref_sigma || !qry_seqs || !qry_off || !qry_len || n_refs <= 0 || n_qrys <= 0)
        return fail(h, -1, "smx_pipeline_run: NULL or empty input");
    if (topology != 0 && topology != 1) return fail(h, -1, "unknown topology_of_dna: %d (ref_query_alignment.py:215)", topology);
    if ((pair_ref == nullptr) != (pair_qry == nullptr)) return fail(h, -1, "pair_ref and pair_qry must both be given or both be NULL");
    if (n_threads <= 0) n_threads = (int)std::min(8u, std::max(1u, std::thread::hardware_concurrency()));  // several handles may run per GPU: never every core per handle
    const uint8_t *const g_comp = comp_table();
    PhaseLog plog;
    const double t_begin = now_s();
    SMX_CUDA(h, cudaSetDevice(h->device));
    const Scoring sc{open, extend, match, mismatch};

    // ---- pool: every sequence stored twice back to back so that rotated slices are contiguous ----
    std::vector<int64_t> rpool((size_t)n_refs), qpool((size_t)n_qrys * 2);
    int64_t total = 0;
    for (int r = 0; r < n_refs; ++r) { if (ref_len[r] <= 0) return fail(h, -1, "reference %d is empty", r); rpool[(size_t)r] = total; total += 2ll * ref_len[r]; }
    for (int q = 0; q < n_qrys; ++q) {
        if (qry_len[q] <= 0) return fail(h, -1, "query %d is empty", q);
        // reads of more than 65535 bases: the select kernel hands their pair-strands to the host re-decision (its
        // histogram median holds 16-bit counts); with omit_too_long they are usually dropped before the scan (:207)
        if (qry_len[q] > (1 << 24)) return fail(h, -1, "query %d: %d bases; this engine supports reads up to 16 Mi bases", q, qry_len[q]);
        qpool[(size_t)q * 2] = total; total += 2ll * qry_len[q];
        qpool[(size_t)q * 2 + 1] = total; total += 2ll * qry_len[q];
    }
    {
        // the sequences go up once, as given (page-locked staging); smx_expand_pool_kernel writes the pool
        int64_t compact_total = 0;
        std::vector<int64_t> rsrc((size_t)n_refs), qsrc((size_t)n_qrys);
        for (int r = 0; r < n_refs; ++r) { rsrc[(size_t)r] = compact_total; compact_total += ref_len[r]; }
        for (int q = 0; q < n_qrys; ++q) { qsrc[(size_t)q] = compact_total; compact_total += qry_len[q]; }
        if (int rc = ensure_pinned(h, h->pin_pool, (size_t)compact_total)) return rc;
        uint8_t *stage = (uint8_t *)h->pin_pool.p;
        std::vector<SmxExpandItem> items((size_t)n_refs + 2 * (size_t)n_qrys);
        for (int r = 0; r < n_refs; ++r) {
            memcpy(stage + rsrc[(size_t)r], ref_seqs + ref_off[r], (size_t)ref_len[r]);
            items[(size_t)r] = SmxExpandItem{rsrc[(size_t)r], rpool[(size_t)r], ref_len[r], 0};
        }
        parallel_for(n_qrys, n_threads, [&](int q, int) {
            memcpy(stage + qsrc[(size_t)q], qry_seqs + qry_off[q], (size_t)qry_len[q]);
            items[(size_t)n_refs + 2 * (size_t)q] = SmxExpandItem{qsrc[(size_t)q], qpool[(size_t)q * 2], qry_len[q], 0};
            items[(size_t)n_refs + 2 * (size_t)q + 1] = SmxExpandItem{qsrc[(size_t)q], qpool[(size_t)q * 2 + 1], qry_len[q], 1};
        });
        if (int rc = ensure(h, h->pool, (size_t)total + 64)) return rc;
        if (int rc = ensure(h, h->d_compact, (size_t)compact_total + 64)) return rc;
        if (int rc = ensure(h, h->d_expand_items, sizeof(SmxExpandItem) * items.size())) return rc;
        if (!h->d_comp.p) {
            if (int rc = ensure(h, h->d_comp, 256)) return rc;
            SMX_CUDA(h, cudaMemcpyAsync(h->d_comp.p, g_comp, 256, cudaMemcpyHostToDevice, h->stream_hi));
        }
        SMX_CUDA(h, cudaMemcpyAsync(h->d_compact.p, stage, (size_t)compact_total, cudaMemcpyHostToDevice, h->stream_hi));
        if (int rc = h2d(h, PIN_EXPAND, h->d_expand_items.p, items.data(), sizeof(SmxExpandItem) * items.size(), h->stream_hi)) return rc;
        smx_expand_pool_kernel<<<(unsigned)items.size(), 256, 0, h->stream_hi>>>((const uint8_t *)h->d_compact.p, (const SmxExpandItem *)h->d_expand_items.p,
                                                                                (const uint8_t *)h->d_comp.p, (uint8_t *)h->pool.p);
        SMX_CUDA(h, cudaGetLastError());
        st.launches += 1;
        if (int rc = pool_finish(h, total)) return rc;
    }
    st.t_pool = now_s() - t_begin;
    plog.cpu(h, "pool");

    // ---- pair-strand table ----
    const int64_t np_ = pair_ref ? n_pairs : (int64_t)n_qrys * n_refs;
    if (np_ * 2 > (int64_t)INT32_MAX) return fail(h, -1, "too many pair-strands");
    st.ps.resize((size_t)np_ * 2);
    for (int64_t p = 0; p < np_; ++p) {
        const int r = pair_ref ? pair_ref[p] : (int)(p % n_refs);
        const int q = pair_qry ? pair_qry[p] : (int)(p / n_refs);
        if (r < 0 || r >= n_refs || q < 0 || q >= n_qrys) return fail(h, -1, "pair %lld: index out of range", (long long)p);
        for (int s = 0; s < 2; ++s) {
            PairStrand &x = st.ps[(size_t)p * 2 + s];
            x.ref = r; x.qry = q; x.strand = s; x.nr = ref_len[r]; x.nq = qry_len[q];
            x.ref_pool = rpool[(size_t)r]; x.qry_pool = qpool[(size_t)q * 2 + s];
            // calc_conserved_region (:207): reads of >= 1.75 reference lengths are dropped (after the scan
            // in the reference; the scan has no side effect, so it is skipped here)
            x.skip = omit_too_long && ((double)x.nr * 1.75 <= (double)x.nq);
        }
    }

    // ---- A3 + A4 on the device, chunked by the counts budget ----
    const bool force_host_select = [] { const char *e = getenv("SMX_SELECT"); return e && std::string(e) == "host"; }();  // cross-check switch
    std::vector<int> active;
    for (size_t i = 0; i < st.ps.size(); ++i) if (!st.ps[i].skip) active.push_back((int)i);
    for (int i : active)  // the scan stages the whole query in shared memory as bit planes (20 bytes per 32 bases)
        if (st.ps[(size_t)i].nq > 300000) return fail(h, -1, "query %d: %d bases are aligned against a reference of %d; the scan supports reads up to 300000 bases", st.ps[(size_t)i].qry, st.ps[(size_t)i].nq, st.ps[(size_t)i].nr);
    size_t pos = 0;
    std::vector<int64_t> s_roff, s_qoff, c_off;
    std::vector<int32_t> s_rlen, s_qlen;
    std::vector<SmxSelPair> sel_pairs;
    std::vector<SmxSelOut> sel_out;
    std::vector<SmxRegion> redo;
    std::vector<int32_t> counts_host;
    std::vector<uint8_t> redo_rc;
    while (pos < active.size()) {
        size_t end = pos;
        int64_t budget = 0;
        while (end < active.size()) {
            const PairStrand &x = st.ps[(size_t)active[end]];
            const int64_t mk = topology == 0 ? x.nr : (int64_t)x.nr + x.nq - 1;
            if (end > pos && budget + mk > kCountsBudget) break;
            budget += mk; ++end;
        }
        const int n = (int)(end - pos);
        s_roff.resize(n); s_qoff.resize(n); s_rlen.resize(n); s_qlen.resize(n); c_off.resize((size_t)n + 1);
        for (int i = 0; i < n; ++i) {
            const PairStrand &x = st.ps[(size_t)active[pos + i]];
            s_roff[i] = x.ref_pool; s_qoff[i] = x.qry_pool; s_rlen[i] = x.nr; s_qlen[i] = x.nq;
        }
        double t0 = now_s();
        if (int rc = smx_scan_prepare(h, s_roff.data(), s_rlen.data(), s_qoff.data(), s_qlen.data(), n, topology, c_off.data())) return rc;
        { const double p0 = now_s(); if (int rc = smx_scan_run(h)) return rc; plog.add(h, "scan", p0, now_s()); }
        st.gpu_scan_ms += h->last_ms; st.scan_cells += h->scan_cells; st.launches += 1;
        st.t_scan += now_s() - t0;
        plog.cpu(h, "scan");
        t0 = now_s();
        sel_pairs.resize(n); sel_out.resize(n);
        for (int i = 0; i < n; ++i) {
            const PairStrand &x = st.ps[(size_t)active[pos + i]];
            SmxSelPair &s = sel_pairs[i];
            s.ref_off = x.ref_pool; s.qry_off = x.qry_pool; s.counts_off = c_off[i]; s.region_off = (int64_t)i * kRegionCap;
            s.sigma = ref_sigma[x.ref]; s.nr = x.nr; s.nq = x.nq;
            s.max_k = (int32_t)(c_off[(size_t)i + 1] - c_off[i]);
            if (topology == 0) { s.slice_lo = 0; s.slice_hi = x.nr; }
            else if (x.nr > x.nq) { s.slice_lo = x.nq - 1; s.slice_hi = x.nr; }   // :363-366
            else { s.slice_lo = x.nr - 1; s.slice_hi = x.nq; }
            s.region_cap = kRegionCap;
        }
        if (int rc = ensure(h, h->d_sel_pairs, sizeof(SmxSelPair) * (size_t)n)) return rc;
        if (int rc = ensure(h, h->d_sel_out, sizeof(SmxSelOut) * (size_t)n)) return rc;
        if (int rc = ensure(h, h->d_regions, sizeof(SmxRegion) * (size_t)n * kRegionCap + 64)) return rc;
        int32_t *d_region_counter = (int32_t *)((char *)h->d_regions.p + sizeof(SmxRegion) * (size_t)n * kRegionCap);
        SMX_CUDA(h, cudaMemsetAsync(d_region_counter, 0, sizeof(int32_t), h->stream_hi));
        if (int rc = h2d(h, PIN_SEL_PAIRS, h->d_sel_pairs.p, sel_pairs.data(), sizeof(SmxSelPair) * (size_t)n, h->stream_hi)) return rc;
        SmxSelArgs sa;
        sa.pool = (const uint8_t *)h->pool.p; sa.pairs = (const SmxSelPair *)h->d_sel_pairs.p; sa.counts = (const int32_t *)h->d_counts.p;
        sa.regions = (SmxRegion *)h->d_regions.p; sa.out = (SmxSelOut *)h->d_sel_out.p; sa.region_counter = d_region_counter; sa.topology = topology;
        const double p_sel0 = now_s();
        SMX_CUDA(h, cudaEventRecord(h->ev0, h->stream_hi));
        smx_select_kernel<<<n, SMX_SEL_THREADS, 0, h->stream_hi>>>(sa);
        SMX_CUDA(h, cudaGetLastError());
        SMX_CUDA(h, cudaEventRecord(h->ev1, h->stream_hi));
        st.launches += 1;
        if (plog.on) { cudaEventSynchronize(h->ev1); plog.add(h, "selk", p_sel0, now_s()); }
        int32_t n_regions_total = 0;
        if (int rc = d2h(h, PIN_SEL_OUT, h->d_sel_out.p, sizeof(SmxSelOut) * (size_t)n, h->stream_hi)) return rc;
        if (int rc = d2h(h, PIN_NREG, d_region_counter, sizeof(int32_t), h->stream_hi)) return rc;
        SMX_CUDA(h, smx_sync(h, h->stream_hi));
        memcpy(sel_out.data(), h->pin_stage[PIN_SEL_OUT].p, sizeof(SmxSelOut) * (size_t)n);
        n_regions_total = *(const int32_t *)h->pin_stage[PIN_NREG].p;
        if (int rc = ensure_pinned(h, h->pin_regions, sizeof(SmxRegion) * (size_t)std::max(n_regions_total, 1))) return rc;
        SmxRegion *reg_host = (SmxRegion *)h->pin_regions.p;
        if (n_regions_total > 0) {
            SMX_CUDA(h, cudaMemcpyAsync(reg_host, h->d_regions.p, sizeof(SmxRegion) * (size_t)n_regions_total, cudaMemcpyDeviceToHost, h->stream_hi));
            SMX_CUDA(h, smx_sync(h, h->stream_hi));
        }
        { float ms = 0.f; cudaEventElapsedTime(&ms, h->ev0, h->ev1); st.gpu_select_ms += ms; }
        plog.add(h, "select", p_sel0, now_s());
        for (int i = 0; i < n; ++i) {
            PairStrand &x = st.ps[(size_t)active[pos + i]];
            const SmxSelOut &o = sel_out[i];
            const SmxRegion *rg = &reg_host[(size_t)o.region_base];
            int nreg = o.n_regions;
            if (o.flags || force_host_select) {  // ambiguous threshold or overflow (or SMX_SELECT=host): decide on the host from this pair-strand's counts
                const SmxSelPair &s = sel_pairs[i];
                counts_host.resize((size_t)s.max_k);
                SMX_CUDA(h, cudaMemcpy(counts_host.data(), (const int32_t *)h->d_counts.p + s.counts_off, sizeof(int32_t) * (size_t)s.max_k, cudaMemcpyDeviceToHost));
                // upper-casing and the reverse-complement strand exist only on the device: restate them for this pair-strand
                auto up = [](uint8_t c) { return (uint8_t)(c >= 'a' && c <= 'z' ? c - 32 : c); };
                const uint8_t *qp = qry_seqs + qry_off[x.qry], *rp = ref_seqs + ref_off[x.ref];
                redo_rc.resize((size_t)x.nq + (size_t)x.nr);
                for (int t = 0; t < x.nq; ++t) redo_rc[(size_t)t] = x.strand ? g_comp[up(qp[x.nq - 1 - t])] : up(qp[t]);
                for (int t = 0; t < x.nr; ++t) redo_rc[(size_t)x.nq + (size_t)t] = up(rp[t]);
                host_regions(redo_rc.data() + x.nq, redo_rc.data(), x.nr, x.nq, topology, counts_host.data(),
                             s.max_k, s.slice_lo, s.slice_hi, s.sigma, redo);
                rg = redo.data(); nreg = (int)redo.size();
                ++st.n_host_redo;
            }
            if (nreg == 0) { x.skip = true; continue; }
            std::vector<SmxRegion> tmp(rg, rg + nreg);
            std::sort(tmp.begin(), tmp.end(), [](const SmxRegion &a, const SmxRegion &b) { return a.k != b.k ? a.k < b.k : a.qs < b.qs; });
            x.regions.resize((size_t)nreg);
            for (int t = 0; t < nreg; ++t) {
                int64_t rs = (int64_t)tmp[t].k + tmp[t].qs, re = (int64_t)tmp[t].k + tmp[t].qe;
                int32_t qs = tmp[t].qs, qe = tmp[t].qe;
                if (topology == 0) {  // :411-417
                    if (re - rs > x.nr - 1) { re = rs - 1; qe = qs + x.nr - 1; }
                    rs %= x.nr; re %= x.nr;
                    if (re < 0) re += x.nr;
                } else {  // :418-420
                    rs -= x.nq - 1; re -= x.nq - 1;
                }
                x.regions[(size_t)t] = Region{(int32_t)rs, (int32_t)re, qs, qe};
            }
        }
        st.t_select += now_s() - t0;
        plog.cpu(h, "select");
        pos = end;
    }

    // ---- A6: chain search on the device (one CTA per pair-strand, one warp per origin); pair-strands with more
    //      than SMX_CHAIN_MAXN regions, or SMX_CHAIN=host, use the host implementation (smx_chain.hpp) ----
    double t0 = now_s();
    std::vector<int> todo;
    for (size_t i = 0; i < st.ps.size(); ++i) if (!st.ps[i].skip) todo.push_back((int)i);
    std::vector<std::vector<int>> traces((size_t)todo.size());
    std::vector<uint8_t> on_host((size_t)todo.size(), 1);
    {
        const char *env = getenv("SMX_CHAIN");
        const bool force_host = env && std::string(env) == "host";
        static const int cls_n[4] = {32, 64, 128, 256};
        std::vector<int> members[4];
        if (!force_host)
            for (size_t ti = 0; ti < todo.size(); ++ti) {
                const int n = (int)st.ps[(size_t)todo[ti]].regions.size();
                for (int c = 0; c < 4; ++c) if (n <= cls_n[c]) { members[c].push_back((int)ti); on_host[ti] = 0; break; }
            }
        std::vector<SmxChainTask> tasks;
        std::vector<int4> regs;
        std::vector<int> task_ti;
        int cls_first[5] = {0, 0, 0, 0, 0};
        int64_t trace_total = 0;
        for (int c = 0; c < 4; ++c) {
            cls_first[c] = (int)tasks.size();
            for (int ti : members[c]) {
                const PairStrand &x = st.ps[(size_t)todo[(size_t)ti]];
                SmxChainTask tk;
                tk.region_off = (int64_t)regs.size(); tk.trace_off = trace_total; tk.n = (int)x.regions.size();
                tk.n_ref = x.nr; tk.n_q = x.nq; tk.pad = 0;
                for (const Region &g : x.regions) regs.push_back(make_int4(g.rs, g.re, g.qs, g.qe));
                trace_total += tk.n;
                tasks.push_back(tk); task_ti.push_back(ti);
            }
        }
        cls_first[4] = (int)tasks.size();
        if (!tasks.empty()) {
            if (int rc = ensure(h, h->d_chain_tasks, sizeof(SmxChainTask) * tasks.size())) return rc;
            if (int rc = ensure(h, h->d_chain_regions, sizeof(int4) * regs.size())) return rc;
            if (int rc = ensure(h, h->d_chain_traces, sizeof(int32_t) * (size_t)std::max<int64_t>(trace_total, 1))) return rc;
            if (int rc = ensure(h, h->d_chain_out, sizeof(SmxChainOut) * tasks.size())) return rc;
            if (int rc = ensure(h, h->d_chain_scratch, sizeof(int32_t) * 3 * regs.size())) return rc;
            const double p_ch0 = now_s();
            if (int rc = h2d(h, PIN_CHAIN_TASKS, h->d_chain_tasks.p, tasks.data(), sizeof(SmxChainTask) * tasks.size(), h->stream_hi)) return rc;
            if (int rc = h2d(h, PIN_CHAIN_REGS, h->d_chain_regions.p, regs.data(), sizeof(int4) * regs.size(), h->stream_hi)) return rc;
            SMX_CUDA(h, cudaEventRecord(h->ev0, h->stream_hi));
            for (int c = 0; c < 4; ++c) {
                const int cnt = cls_first[c + 1] - cls_first[c];
                if (cnt == 0) continue;
                const int nmax = cls_n[c];
                const size_t smem = sizeof(int) * (size_t)(4 * nmax + SMX_CHAIN_WARPS * (17 * nmax + nmax * (nmax / 32)));
                if (smem > 48 * 1024) SMX_CUDA(h, cudaFuncSetAttribute(smx_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                SmxChainArgs ca;
                ca.tasks = (const SmxChainTask *)h->d_chain_tasks.p + cls_first[c];
                ca.regions = (const int4 *)h->d_chain_regions.p;
                ca.traces = (int32_t *)h->d_chain_traces.p;
                ca.out = (SmxChainOut *)h->d_chain_out.p + cls_first[c];
                ca.origin_scratch = (int32_t *)h->d_chain_scratch.p;
                ca.topology = topology; ca.nmax = nmax;
                // few pair-strands with many regions: spread the origins of one pair-strand over several CTAs so
                // that the launch fills the GPU, then pick across the slices in a second launch
                int slices = 1;
                if (topology == 0 && cnt < kChainCtaTarget * h->sm_count) slices = std::max(1, std::min(nmax / SMX_CHAIN_WARPS, (kChainCtaTarget * h->sm_count + cnt - 1) / cnt));
                ca.n_slices = slices; ca.phase = 0;
                smx_chain_kernel<<<cnt * slices, SMX_CHAIN_WARPS * 32, smem, h->stream_hi>>>(ca);
                SMX_CUDA(h, cudaGetLastError());
                st.launches += 1;
                if (slices > 1) {
                    ca.n_slices = 1; ca.phase = 1;
                    smx_chain_kernel<<<cnt, SMX_CHAIN_WARPS * 32, smem, h->stream_hi>>>(ca);
                    SMX_CUDA(h, cudaGetLastError());
                    st.launches += 1;
                }
            }
            SMX_CUDA(h, cudaEventRecord(h->ev1, h->stream_hi));
            std::vector<SmxChainOut> couts(tasks.size());
            std::vector<int32_t> tr((size_t)std::max<int64_t>(trace_total, 1));
            if (int rc = d2h(h, PIN_CHAIN_OUT, h->d_chain_out.p, sizeof(SmxChainOut) * tasks.size(), h->stream_hi)) return rc;
            if (int rc = d2h(h, PIN_CHAIN_TR, h->d_chain_traces.p, sizeof(int32_t) * (size_t)std::max<int64_t>(trace_total, 1), h->stream_hi)) return rc;
            SMX_CUDA(h, smx_sync(h, h->stream_hi));
            memcpy(couts.data(), h->pin_stage[PIN_CHAIN_OUT].p, sizeof(SmxChainOut) * tasks.size());
            memcpy(tr.data(), h->pin_stage[PIN_CHAIN_TR].p, sizeof(int32_t) * (size_t)std::max<int64_t>(trace_total, 1));
            { float ms = 0.f; cudaEventElapsedTime(&ms, h->ev0, h->ev1); st.gpu_chain_ms += ms; }
            plog.add(h, "chain", p_ch0, now_s());
            for (size_t k = 0; k < tasks.size(); ++k) {
                std::vector<int> &t = traces[(size_t)task_ti[k]];
                t.assign(tr.begin() + tasks[k].trace_off, tr.begin() + tasks[k].trace_off + couts[k].trace_len);
            }
        }
    }
    // ---- A5/A7: plan on host threads (and A6 for the pair-strands kept on the host) ----
    std::vector<std::vector<PlannedProblem>> local((size_t)todo.size());
    std::vector<ChainSearch> searchers((size_t)std::max(1, n_threads));
    parallel_for((int)todo.size(), n_threads, [&](int ti, int tid) {
        PairStrand &x = st.ps[(size_t)todo[(size_t)ti]];
        std::vector<PlannedProblem> &probs = local[(size_t)ti];
        std::vector<int> &trace = traces[(size_t)ti];
        bool ok = true;
        if (on_host[(size_t)ti])
            ok = topology == 0 ? searchers[(size_t)tid].circular(x.regions, x.nr, x.nq, trace) : searchers[(size_t)tid].linear(x.regions, trace);
        if (!ok || trace.empty()) { x.failed = true; return; }
        std::vector<Region> P(trace.size());
        for (size_t t = 0; t < trace.size(); ++t) P[t] = x.regions[(size_t)trace[t]];
        const int nr = x.nr, nq = x.nq;
        auto add_problem = [&](uint8_t mode, int qs, int qlen, int rs, int rlen) -> int {
            // slices are in rotated coordinates; the doubled pool makes them contiguous
            PlannedProblem pp;
            pp.s1_off = x.qry_pool + ((int64_t)x.q_off + qs) % nq; pp.m = qlen;
            pp.s2_off = x.ref_pool + ((int64_t)x.ref_off + rs) % nr; pp.n = rlen;
            pp.mode = mode;
            probs.push_back(pp);
            return (int)probs.size() - 1;
        };
        auto literal = [&](char c, int len) { if (len > 0) x.pieces.push_back(Piece{PIECE_LITERAL, c, len, {-1, -1}, 0, 0, 0}); };
        auto fill_gap = [&](int rs, int re, int qs, int qe) -> bool {  // execute_alignment_using_conserved_regions_core (:330-345)
            if (rs > re) { if (rs - 1 != re) return false; literal('I', qe - qs + 1); return true; }
            if (qs > qe) { if (qs - 1 != qe) return false; literal('D', re - rs + 1); return true; }
            const int id = add_problem(SMX_MODE_NW, qs, qe - qs + 1, rs, re - rs + 1);
            x.pieces.push_back(Piece{PIECE_NW, 0, 0, {id, -1}, 0, 0, 0});
            return true;
        };
        const size_t np2 = P.size();
        if (topology == 0) {
            // set_tidy_offset (:725-736): start the chain at the region with the smallest query start
            if (!(P.back().qs < P.back().qe)) { x.failed = true; return; }
            size_t amin = 0;
            for (size_t t = 1; t < np2; ++t) if (P[t].qs < P[amin].qs) amin = t;
            std::rotate(P.begin(), P.begin() + (long)amin, P.end());
            x.ref_off = P[0].rs; x.q_off = P[0].qs;
            for (Region &g : P) {
                g.rs -= x.ref_off; g.re -= x.ref_off; g.qs -= x.q_off; g.qe -= x.q_off;
                if (g.rs < 0) g.rs += nr;
                if (g.re < 0) g.re += nr;
                if (g.qs < 0) g.qs += nq;
                if (g.qe < 0) g.qe += nq;
            }
            const int q_end = nq - x.q_off - 1;
            for (size_t t = 0; t < np2; ++t) {  // iter_regions (:763-781)
                const Region &g = P[t];
                const Region &nx = P[(t + 1) % np2];
                literal('=', g.re - g.rs + 1);
                const int rs = g.re + 1, qs = g.qe + 1, re = nx.rs - 1, qe = nx.qs - 1;
                if (re != -1) { if (!fill_gap(rs, re, qs, qe)) { x.failed = true; return; } continue; }
                if (qe != -1) { x.failed = true; return; }  // assert at :285
                if (rs == nr) literal('S', nq - qs);
                else if (qs == nq) literal('H', nr - rs);
                else {  // my_special_DP (:293-302)
                    const int n1 = std::max(0, q_end + 1 - qs), n2 = nq - (q_end + 1), nrf = nr - rs;
                    Piece pc{PIECE_SPECIAL, 0, 0, {-1, -1}, nrf, n1, n2};
                    if (n1 > 0) pc.prob[0] = add_problem(SMX_MODE_SG_QE_DE, qs, n1, rs, nrf);
                    if (n2 > 0) pc.prob[1] = add_problem(SMX_MODE_SG_QB_DB, q_end + 1, n2, rs, nrf);
                    x.pieces.push_back(pc);
                }
            }
        } else {
            x.ref_off = 0; x.q_off = 0;
            // execute_alignment_using_conserved_regions_linear (:226-268)
            const int hre = P[0].rs - 1, hqe = P[0].qs - 1;
            if (hre == -1) literal('S', hqe + 1);
            else if (hqe == -1) literal('H', hre + 1);
            else { const int id = add_problem(SMX_MODE_SG_QB_DB, 0, hqe + 1, 0, hre + 1); x.pieces.push_back(Piece{PIECE_SG_HEAD, 0, 0, {id, -1}, 0, 0, 0}); }
            for (size_t t = 0; t + 1 < np2; ++t) {
                literal('=', P[t].re - P[t].rs + 1);
                if (!fill_gap(P[t].re + 1, P[t + 1].rs - 1, P[t].qe + 1, P[t + 1].qs - 1)) { x.failed = true; return; }
            }
            const Region &g = P.back();
            literal('=', g.re - g.rs + 1);
            const int rs = g.re + 1, qs = g.qe + 1;
            if (rs == nr) literal('S', nq - qs);
            else if (qs == nq) literal('H', nr - rs);
            else { const int id = add_problem(SMX_MODE_SG_QE_DE, qs, nq - qs, rs, nr - rs); x.pieces.push_back(Piece{PIECE_SG_TAIL, 0, 0, {id, -1}, 0, 0, 0}); }
        }
    });
    // global problem ids
    std::vector<int64_t> p_s1, p_s2;
    std::vector<int32_t> p_m, p_n;
    std::vector<uint8_t> p_mode, p_clip;
    for (size_t ti = 0; ti < todo.size(); ++ti) {
        PairStrand &x = st.ps[(size_t)todo[ti]];
        if (x.failed) { ++st.n_failed; continue; }
        const int base = (int)p_m.size();
        for (const PlannedProblem &pp : local[ti]) {
            if (pp.m <= 0 || pp.n <= 0) return fail(h, -5, "internal: empty DP slice planned");
            p_s1.push_back(pp.s1_off); p_s2.push_back(pp.s2_off); p_m.push_back(pp.m); p_n.push_back(pp.n); p_mode.push_back(pp.mode);
            p_clip.push_back(pp.mode == SMX_MODE_NW ? 0 : 1);
        }
        for (Piece &pc : x.pieces) for (int z = 0; z < 2; ++z) if (pc.prob[z] >= 0) pc.prob[z] += base;
    }
    local.clear();
    st.t_chain = now_s() - t0;
    plog.cpu(h, "chain_plan");

    // ---- A8 on the device ----
    t0 = now_s();
    const int n_prob = (int)p_m.size();
    std::vector<int64_t> cg_off((size_t)n_prob + 1, 0);
    uint32_t *cg = nullptr;
    st.n_dp = n_prob;
    if (n_prob > 0) {
        if (int rc = align_prepare_impl(h, p_s1.data(), p_m.data(), p_s2.data(), p_n.data(), p_mode.data(), p_clip.data(), n_prob, open, extend, match, mismatch)) return rc;
        int64_t truns = 0;
        {
            const double p0 = now_s();
            if (int rc = align_run_impl(h, &truns, &g_dp_gate[h->device & 15])) return rc;
            plog.add(h, "dp", p0, now_s());
        }
        st.gpu_dp_ms = h->last_ms; st.dp_cells = h->cells; st.launches += h->last_launches;
        for (int c = 0; c < 16; ++c) st.dp_kernel_ms[c] = h->kernel_ms[c];
        for (int c = 0; c < 14; ++c) st.dp_class_cells[c] = h->class_cells[c];
        st.dp_ck_cells = h->ck_cells; st.dp_ck_problems = h->ck_problems;
        for (int c = 0; c < 16; ++c) st.dp_kernel_launches[c] = h->kernel_launches[c];
        if (int rc = ensure_pinned(h, h->pin_cigar, sizeof(uint32_t) * (size_t)std::max<int64_t>(truns, 1))) return rc;
        cg = (uint32_t *)h->pin_cigar.p;
        { const double p0 = now_s(); if (int rc = smx_align_fetch(h, nullptr, nullptr, nullptr, nullptr, nullptr, cg_off.data(), cg, std::max<int64_t>(truns, 1))) return rc; plog.add(h, "fetch", p0, now_s()); }
    }
    st.t_dp = now_s() - t0;
    plog.cpu(h, "dp");
    plog.add(h, "call", t_begin, now_s());

    // ---- A9-A11 on host threads (run-length domain; re-clipping already happened in the traceback kernel) ----
    t0 = now_s();
    std::atomic<int> n_bad(0);
    const std::vector<SmxResult> &dres = h->h_results;
    const bool merge_chars = [] { const char *e = getenv("SMX_MERGE"); return e && std::string(e) == "chars"; }();  // cross-check switch
    const bool sub_timers = plog.on && n_threads == 1;   // SMX_PHASE_LOG with one host thread: where the stitch stage spends its time
    double t_sub[4] = {0, 0, 0, 0};                      // assemble, my_special_DP merge, rotate, score
    parallel_for((int)todo.size(), n_threads, [&](int ti, int) {
        PairStrand &x = st.ps[(size_t)todo[(size_t)ti]];
        if (x.failed) return;
        Runs &cig = x.runs;
        cig.clear();
        double ts0 = sub_timers ? now_s() : 0.0, t_merge = 0.0;
        auto n_runs_of = [&](int pid) { return pid >= 0 ? cg_off[(size_t)pid + 1] - cg_off[(size_t)pid] : (int64_t)0; };
        {
            int64_t need = 4;
            for (const Piece &pc : x.pieces) need += 4 + n_runs_of(pc.kind == PIECE_LITERAL ? -1 : pc.prob[0]) + (pc.kind == PIECE_SPECIAL ? n_runs_of(pc.prob[1]) : 0);
            cig.reserve((size_t)need);
        }
        auto clipped = [&](int pid) {
            ClippedSg c;
            c.kept = &cg[(size_t)cg_off[(size_t)pid]];
            c.n_kept = cg_off[(size_t)pid + 1] - cg_off[(size_t)pid];
            c.n_s = dres[(size_t)pid].n_s; c.n_h = dres[(size_t)pid].n_h;
            return c;
        };
        for (const Piece &pc : x.pieces) {
            switch (pc.kind) {
            case PIECE_LITERAL: push_run(cig, op_code(pc.letter), pc.len); break;
            case PIECE_NW: append_maximal_runs(cig, &cg[(size_t)cg_off[(size_t)pc.prob[0]]], n_runs_of(pc.prob[0])); break;
            case PIECE_SG_HEAD: {  // my_special_sg_qb_db: H.. S.. aligned
                const ClippedSg c = clipped(pc.prob[0]);
                push_run(cig, 5, c.n_h); push_run(cig, 4, c.n_s); append_maximal_runs(cig, c.kept, c.n_kept);
                break;
            }
            case PIECE_SG_TAIL: {  // my_special_sg_qe_de: aligned S.. H..
                const ClippedSg c = clipped(pc.prob[0]);
                append_maximal_runs(cig, c.kept, c.n_kept); push_run(cig, 4, c.n_s); push_run(cig, 5, c.n_h);
                break;
            }
            case PIECE_SPECIAL: {
                ClippedSg r1, r2;
                if (pc.prob[0] >= 0) r1 = clipped(pc.prob[0]); else r1.n_h = pc.n_ref;   // "H" * len(ref), end_ref = -1
                if (pc.prob[1] >= 0) r2 = clipped(pc.prob[1]); else r2.n_h = pc.n_ref;   // "H" * len(ref), beg_ref = len(ref)
                const double tm0 = sub_timers ? now_s() : 0.0;
                const bool ok = merge_special_runs(r1, r2, pc.n_ref, pc.nq1, pc.nq2, sc, cig, merge_chars);
                if (sub_timers) t_merge += now_s() - tm0;
                if (!ok) { x.failed = true; n_bad++; return; }
                break;
            }
            }
        }
        int64_t new_off = 0;
        const double ts1 = sub_timers ? now_s() : 0.0;
        if (topology == 0) new_off = rotate_to_ref0_runs(cig, x.ref_off, x.q_off);
        const double ts2 = sub_timers ? now_s() : 0.0;
        x.score = (int32_t)cigar_score_runs(cig, sc);
        x.new_q_off = (int32_t)new_off;
        if (sub_timers) { const double ts3 = now_s(); t_sub[0] += ts1 - ts0 - t_merge; t_sub[1] += t_merge; t_sub[2] += ts2 - ts1; t_sub[3] += ts3 - ts2; }
    });
    if (sub_timers) fprintf(stderr, "STITCH %p assemble %.4f merge %.4f rotate %.4f score %.4f\n", (void *)h, t_sub[0], t_sub[1], t_sub[2], t_sub[3]);
    st.n_failed += n_bad.load();
    st.t_stitch = now_s() - t0;
    plog.cpu(h, "stitch");

    st.out_off.assign(st.ps.size() + 1, 0);
    for (size_t i = 0; i < st.ps.size(); ++i) st.out_off[i + 1] = st.out_off[i] + (int64_t)st.ps[i].runs.size();
    st.total_runs = st.out_off.back();
    st.t_total = now_s() - t_begin;
    st.valid = true;
    if (total_runs) *total_runs = st.total_runs;
    return 0;
}

extern "C" int smx_pipeline_fetch(smx_handle *h, int32_t *score, int32_t *new_query_seq_offset, uint8_t *status, int64_t *cigar_off,
                                  uint32_t *cigar_rle, int64_t cigar_capacity)
{
    if (!h) return -1;
    SmxPipelineState &st = *pipeline_state(h);
    if (!st.valid) return fail(h, -1, "smx_pipeline_fetch: no completed smx_pipeline_run");
    if (cigar_capacity < st.total_runs) return fail(h, -1, "smx_pipeline_fetch: cigar buffer too small");
    for (size_t i = 0; i < st.ps.size(); ++i) {
        const smx::PairStrand &x = st.ps[i];
        const bool empty = x.skip || x.failed;
        if (score) score[i] = empty ? 0 : x.score;
        if (new_query_seq_offset) new_query_seq_offset[i] = empty ? INT32_MIN : x.new_q_off;
        if (status) status[i] = x.failed ? 2 : (x.skip ? 1 : 0);
        if (cigar_rle && !x.runs.empty()) memcpy(cigar_rle + st.out_off[i], x.runs.data(), sizeof(uint32_t) * x.runs.size());
    }
    if (cigar_off) memcpy(cigar_off, st.out_off.data(), sizeof(int64_t) * st.out_off.size());
    return 0;
}

extern "C" int smx_pipeline_stats(smx_handle *h, double *out, int32_t n)
{
    if (!h || !out) return -1;
    const SmxPipelineState &st = *pipeline_state(h);
    const double v[] = {st.t_total, st.t_pool, st.t_scan, st.t_select, st.t_chain, st.t_dp, st.t_stitch, st.gpu_scan_ms, st.gpu_select_ms + st.gpu_chain_ms,
                        st.gpu_dp_ms, (double)st.scan_cells, (double)st.dp_cells, (double)st.n_dp, (double)st.n_host_redo, (double)st.n_failed,
                        (double)st.launches, (double)st.ps.size()};
    const int m = (int)(sizeof(v) / sizeof(v[0]));
    for (int i = 0; i < n; ++i) {
        if (i < m) out[i] = v[i];
        else if (i < m + 16) out[i] = st.dp_kernel_ms[i - m];
        else if (i < m + 30) out[i] = (double)st.dp_class_cells[i - m - 16];
        else if (i < m + 46) out[i] = (double)st.dp_kernel_launches[i - m - 30];
        else if (i == m + 46) out[i] = st.gpu_select_ms;
        else if (i == m + 47) out[i] = st.gpu_chain_ms;
        else if (i == m + 48) out[i] = (double)st.dp_ck_cells;
        else if (i == m + 49) out[i] = (double)st.dp_ck_problems;
        else out[i] = 0.0;
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Row N1: the alignment operations of MyAlignerBase (ref_query_alignment.py:21-163) as one batch -- what
// MyMSA.exec_chunk_poa (modules/msa.py:1249-1355) calls per read chunk, collected over a whole polish round
// ------------------------------------------------------------------------------------------------
extern "C" int smx_ops_run(smx_handle *h, const uint8_t *kind, const int64_t *ref_off, const int32_t *ref_len, const int64_t *q1_off,
                           const int32_t *q1_len, const int64_t *q2_off, const int32_t *q2_len, int32_t n, int32_t open, int32_t extend,
                           int32_t match, int32_t mismatch, int32_t n_threads, int64_t *total_runs)
{
    using namespace smx;
    if (!h) return -1;
    h->ops_valid = false;
    if (total_runs) *total_runs = 0;
    if (n < 0 || (n > 0 && (!kind || !ref_off || !ref_len || !q1_off || !q1_len))) return fail(h, -1, "smx_ops_run: NULL argument");
    if (match <= 0 || mismatch >= 0) return fail(h, -1, "smx_ops_run needs match > 0 and mismatch < 0 (my_classes.py:547-552)");
    if (n_threads <= 0) n_threads = (int)std::min(8u, std::max(1u, std::thread::hardware_concurrency()));
    const Scoring sc{open, extend, match, mismatch};
    // plan: one DP problem for kinds 0..3, up to two for my_special_DP (an empty half has no problem)
    std::vector<int64_t> p_s1, p_s2;
    std::vector<int32_t> p_m, p_n;
    std::vector<uint8_t> p_mode, p_clip;
    std::vector<int32_t> first((size_t)n * 2, -1);
    auto add = [&](int64_t qo, int32_t qn, int64_t ro, int32_t rn, uint8_t mode, uint8_t clip) {
        p_s1.push_back(qo); p_m.push_back(qn); p_s2.push_back(ro); p_n.push_back(rn); p_mode.push_back(mode); p_clip.push_back(clip);
        return (int32_t)p_m.size() - 1;
    };
    for (int i = 0; i < n; ++i) {
        if (ref_len[i] <= 0) return fail(h, -1, "operation %d: empty reference", i);
        switch (kind[i]) {
        case SMX_OPKIND_NW: case SMX_OPKIND_SG_QE_DE: case SMX_OPKIND_SG_QB_DB: case SMX_OPKIND_SW:
            if (q1_len[i] <= 0) return fail(h, -1, "operation %d: empty query (callers never pass empty strings)", i);
            first[(size_t)i * 2] = add(q1_off[i], q1_len[i], ref_off[i], ref_len[i],
                                       kind[i] == SMX_OPKIND_NW ? SMX_MODE_NW : kind[i] == SMX_OPKIND_SG_QE_DE ? SMX_MODE_SG_QE_DE : kind[i] == SMX_OPKIND_SG_QB_DB ? SMX_MODE_SG_QB_DB : SMX_MODE_SW,
                                       (kind[i] == SMX_OPKIND_SG_QE_DE || kind[i] == SMX_OPKIND_SG_QB_DB) ? 1 : 0);
            break;
        case SMX_OPKIND_SPECIAL_DP:
            if (!q2_off || !q2_len) return fail(h, -1, "operation %d: my_special_DP needs the second query", i);
            if (q1_len[i] > 0) first[(size_t)i * 2] = add(q1_off[i], q1_len[i], ref_off[i], ref_len[i], SMX_MODE_SG_QE_DE, 1);
            if (q2_len[i] > 0) first[(size_t)i * 2 + 1] = add(q2_off[i], q2_len[i], ref_off[i], ref_len[i], SMX_MODE_SG_QB_DB, 1);
            break;
        default: return fail(h, -1, "operation %d: unknown kind %d", i, (int)kind[i]);
        }
    }
    const int n_prob = (int)p_m.size();
    std::vector<int64_t> cg_off((size_t)n_prob + 1, 0);
    std::vector<uint32_t> cg;
    if (n_prob > 0) {
        if (int rc = align_prepare_impl(h, p_s1.data(), p_m.data(), p_s2.data(), p_n.data(), p_mode.data(), p_clip.data(), n_prob, open, extend, match, mismatch)) return rc;
        int64_t truns = 0;
        if (int rc = align_run_impl(h, &truns, nullptr)) return rc;
        cg.resize((size_t)std::max<int64_t>(truns, 1));
        if (int rc = smx_align_fetch(h, nullptr, nullptr, nullptr, nullptr, nullptr, cg_off.data(), cg.data(), (int64_t)cg.size())) return rc;
    }
    const std::vector<SmxResult> &dres = h->h_results;
    std::vector<Runs> out((size_t)n);
    h->ops_status.assign((size_t)n, 0);
    auto clipped = [&](int pid) {
        ClippedSg c;
        c.kept = &cg[(size_t)cg_off[(size_t)pid]];
        c.n_kept = cg_off[(size_t)pid + 1] - cg_off[(size_t)pid];
        c.n_s = dres[(size_t)pid].n_s; c.n_h = dres[(size_t)pid].n_h;
        return c;
    };
    parallel_for(n, n_threads, [&](int i, int) {
        Runs &v = out[(size_t)i];
        const int p0 = first[(size_t)i * 2], p1 = first[(size_t)i * 2 + 1];
        switch (kind[i]) {
        case SMX_OPKIND_NW:
            append_maximal_runs(v, &cg[(size_t)cg_off[(size_t)p0]], cg_off[(size_t)p0 + 1] - cg_off[(size_t)p0]);
            break;
        case SMX_OPKIND_SG_QE_DE: {  // my_special_sg_qe_de (:60-73): aligned S.. H..
            const ClippedSg c = clipped(p0);
            append_maximal_runs(v, c.kept, c.n_kept); push_run(v, 4, c.n_s); push_run(v, 5, c.n_h);
            break;
        }
        case SMX_OPKIND_SG_QB_DB: {  // my_special_sg_qb_db (:74-87): H.. S.. aligned
            const ClippedSg c = clipped(p0);
            push_run(v, 5, c.n_h); push_run(v, 4, c.n_s); append_maximal_runs(v, c.kept, c.n_kept);
            break;
        }
        case SMX_OPKIND_SW: {  // my_special_sw (:34-56)
            const SmxResult &r = dres[(size_t)p0];
            const uint32_t *runs = &cg[(size_t)cg_off[(size_t)p0]];
            const int64_t nr = cg_off[(size_t)p0 + 1] - cg_off[(size_t)p0];
            if (nr == 0) { h->ops_status[(size_t)i] = 2; break; }  // my_cigar[0] of an empty local alignment raises in the reference
            if ((runs[0] & 15u) != 2u) {
                push_run(v, 5, r.beg_r); push_run(v, 4, r.beg_q);
                append_runs(v, runs, nr);
            } else {  // a leading D run (parasail quirk, :37-54): becomes H; the reference asserts beg_ref == beg_query == 0
                if (r.beg_r != 0 || r.beg_q != 0) { h->ops_status[(size_t)i] = 2; break; }
                push_run(v, 5, runs[0] >> 4);
                append_runs(v, runs + 1, nr - 1);
            }
            push_run(v, 4, (int64_t)q1_len[i] - r.end_q - 1);
            push_run(v, 5, (int64_t)ref_len[i] - r.end_r - 1);
            break;
        }
        default: {  // my_special_DP (:89-163)
            ClippedSg r1, r2;
            if (p0 >= 0) r1 = clipped(p0); else r1.n_h = ref_len[i];   // "H" * len(ref), end_ref = -1
            if (p1 >= 0) r2 = clipped(p1); else r2.n_h = ref_len[i];   // "H" * len(ref), beg_ref = len(ref)
            if (!merge_special_runs(r1, r2, ref_len[i], q1_len[i], q2_len[i], sc, v)) h->ops_status[(size_t)i] = 2;
        }
        }
    });
    h->ops_off.assign((size_t)n + 1, 0);
    for (int i = 0; i < n; ++i) h->ops_off[(size_t)i + 1] = h->ops_off[(size_t)i] + (int64_t)out[(size_t)i].size();
    h->ops_runs.resize((size_t)h->ops_off[(size_t)n]);
    parallel_for(n, n_threads, [&](int i, int) {
        if (!out[(size_t)i].empty()) memcpy(h->ops_runs.data() + h->ops_off[(size_t)i], out[(size_t)i].data(), sizeof(uint32_t) * out[(size_t)i].size());
    });
    h->ops_valid = true;
    if (total_runs) *total_runs = h->ops_off[(size_t)n];
    return 0;
}

extern "C" int smx_ops_fetch(smx_handle *h, int64_t *cigar_off, uint32_t *cigar_rle, int64_t cigar_capacity, uint8_t *status)
{
    if (!h) return -1;
    if (!h->ops_valid) return fail(h, -1, "smx_ops_fetch: no completed smx_ops_run");
    const int64_t total = h->ops_off.empty() ? 0 : h->ops_off.back();
    if (cigar_capacity < total) return fail(h, -1, "smx_ops_fetch: cigar buffer too small");
    if (cigar_off) memcpy(cigar_off, h->ops_off.data(), sizeof(int64_t) * h->ops_off.size());
    if (cigar_rle && total > 0) memcpy(cigar_rle, h->ops_runs.data(), sizeof(uint32_t) * (size_t)total);
    if (status && !h->ops_status.empty()) memcpy(status, h->ops_status.data(), h->ops_status.size());
    return 0;
}

static SmxPipelineState *pipeline_state(smx_handle *h)
{
    if (!h->pipe) h->pipe = new SmxPipelineState();
    return h->pipe;
}

static void smx_pipeline_free(smx_handle *h)
{
    delete h->pipe;
    h->pipe = nullptr;
}
