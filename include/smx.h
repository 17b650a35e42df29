/*
 * smx.h -- C-ABI of the B200 (sm_100a) alignment engine for SAVEMONEY's hot path.
 *
 * Drop-in boundary: these entry points are what the reference's bindings for the path would
 * call instead of
 *   parasail.nw_trace / sg_qe_de_trace / sg_qb_db_trace / sw_trace
 *        (savemoney/modules/ref_query_alignment.py:35,58,68,82,343)          -> smx_align_*
 *   af.k_mer_offset_analysis_2
 *        (savemoney/modules/cython_functions/alignment_functions.pyx:46-64,
 *         called at ref_query_alignment.py:385)                               -> smx_scan_*
 *   multiprocessing.Pool fan-out over reads / plasmid pairs
 *        (post_analysis/post_analysis_core.py:56-65, pre_survey/pre_survey_core.py:237-245)
 *                                                                             -> one handle per GPU
 *
 * Conventions
 *   - plain pointers and sizes only; every HOST buffer is allocated by the caller;
 *     the library owns only the device workspaces inside a handle.
 *   - return value: 0 = success, negative = error (see smx_last_error); never throws, never
 *     aborts, never falls back to a CPU path.
 *   - a handle is bound to one CUDA device and must be used by one host thread at a time.
 *   - sequences are upper-case ASCII bytes in one pool; s1 is the QUERY (read), s2 the
 *     REFERENCE (plasmid), exactly parasail's argument order.  Bytes other than A,C,G,T score 0
 *     against everything (parasail.matrix_create("ACGT", m, mm) wildcard row/column,
 *     ref_query_alignment.py:32) but are printed '=' when the two characters are equal.
 *   - CIGAR runs are uint32 `len << 4 | op` with BAM op codes ('I'=1, 'D'=2, '='=7, 'X'=8),
 *     in forward order; 'I' consumes a query base, 'D' a reference base.
 */
#ifndef SMX_H
#define SMX_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SMX_MODE_NW 0        /* parasail.nw_trace          (global)                              */
#define SMX_MODE_SG_QE_DE 1  /* parasail.sg_qe_de_trace    (free trailing gaps on both)          */
#define SMX_MODE_SG_QB_DB 2  /* parasail.sg_qb_db_trace    (free leading gaps on both)           */
#define SMX_MODE_SW 3        /* parasail.sw_trace          (local)                               */

#define SMX_OP_I 1u
#define SMX_OP_D 2u
#define SMX_OP_S 4u
#define SMX_OP_H 5u
#define SMX_OP_EQ 7u
#define SMX_OP_X 8u

#define SMX_TOPOLOGY_CIRCULAR 0 /* param_dict["topology_of_dna"] == 0 (post_analysis.py:23) */
#define SMX_TOPOLOGY_LINEAR 1

typedef struct smx_handle smx_handle;

int smx_version(void);
/* CUDA devices visible to this process (0 when there is none: the engine has no CPU path).  The host layer shards
 * reads over them, one handle set per device, where the reference fans out over multiprocessing.Pool(n_cpu)
 * (post_analysis_core.py:58-61, pre_survey_core.py:240-244). */
int smx_device_count(void);

/* Creates an engine bound to `device_ordinal`.  `stream` is a cudaStream_t passed as void*
 * (e.g. torch.cuda.current_stream().cuda_stream) or NULL to let the engine create its own. */
int smx_create(smx_handle **out, int device_ordinal, void *stream);
void smx_destroy(smx_handle *h);
/* message of the last failing call on this handle ("" if none); h may be NULL for create errors */
const char *smx_last_error(const smx_handle *h);

/* Upper bound for the device bytes of packed traceback kept in flight at once (default min(16 GiB, an eighth of
 * the free device memory at creation); SMX_TRACE_BUDGET_GB in the environment overrides);
 * larger batches are cut into chunks that are processed back to back.  A problem takes 0.5 byte per cell of
 * direction bits, or -- packed global / semi-global problems of at least 513 rows and 1 500 columns, whose forward
 * pass keeps checkpoints and whose traceback recomputes the blocks it crosses -- 0.19 byte per cell. */
int smx_set_trace_budget(smx_handle *h, int64_t bytes);
int64_t smx_get_trace_budget(const smx_handle *h);

/* ---- sequence pool -------------------------------------------------------------------------
 * Copies `n` bytes of upper-case ASCII sequence data to the device; the offsets handed to the
 * calls below index this pool.  Replaces any previous pool of the handle. */
int smx_pool_upload(smx_handle *h, const uint8_t *seqs, int64_t n);

/* ---- A8: batched affine-gap DP with traceback (replaces the parasail *_trace calls) ----------
 * One call = three phases that can also be driven separately (bench.py times phase 2 alone):
 *   smx_align_prepare : copy the n problem descriptors to the device, plan classes / chunks
 *   smx_align_run     : forward DP (packed direction bits in HBM) + traceback kernels; results
 *                       stay on the device.  *total_runs receives the number of CIGAR runs.
 *   smx_align_fetch   : copy results to caller buffers.
 * Gap of length L costs open + (L-1)*extend (open, extend > 0), as in parasail. */
int smx_align_prepare(smx_handle *h,
                      const int64_t *s1_off, const int32_t *s1_len,
                      const int64_t *s2_off, const int32_t *s2_len,
                      const uint8_t *mode, int32_t n,
                      int32_t open, int32_t extend, int32_t match, int32_t mismatch);
int smx_align_run(smx_handle *h, int64_t *total_runs);
/* score/end_query/end_ref/beg_query/beg_ref: int32[n] (parasail Result.score, .end_query,
 * .end_ref, .cigar.beg_query, .cigar.beg_ref); cigar_off: int64[n+1] (prefix offsets into
 * cigar_rle, filled by the call); cigar_rle: uint32[total_runs] as returned by smx_align_run. */
int smx_align_fetch(smx_handle *h, int32_t *score, int32_t *end_query, int32_t *end_ref,
                    int32_t *beg_query, int32_t *beg_ref, int64_t *cigar_off, uint32_t *cigar_rle,
                    int64_t cigar_capacity);
/* device milliseconds (CUDA events on the engine's stream) of the last smx_align_run / smx_scan_run */
float smx_last_run_ms(const smx_handle *h);
/* DP cells (sum of s1_len*s2_len) of the prepared batch, and the number of kernel launches of the last run */
int64_t smx_align_cells(const smx_handle *h);
int32_t smx_last_run_launches(const smx_handle *h);

/* device milliseconds per kernel of the last smx_align_run, measured with CUDA events on the engine's
 * stream.  Forward DP kernel classes c: kshift*2 + local for the int32 warp-per-problem kernels (K = 1 << kshift
 * rows per lane, c = 0..7), 8 + local for the int32 CTA-per-problem kernel (more than 512 query rows), 10..13 =
 * packed s16x2 kernel with 1 / 2 / 4 / 8 warps per problem (two bands per warp; problems of more than 256 rows
 * whose score range fits 16 bits and whose reference slice is pure ACGT); ms16[14] = traceback, ms16[15] = CIGAR
 * compaction; cells14[c] = DP cells per class.  SMX_S16=0 in the environment keeps everything on int32. */
int smx_align_kernel_ms(const smx_handle *h, double *ms16, int64_t *cells14);

/* diagnostic: raw copy of the first nbytes of the direction-bit arena of the last chunk */
int smx_debug_copy_trace(smx_handle *h, uint8_t *dst, int64_t nbytes);

/* ---- A3: diagonal match-count scan (replaces af.k_mer_offset_analysis_2) ---------------------
 * For pair p with reference r = pool[ref_off[p] : +ref_len[p]] and query q = pool[qry_off[p] : +qry_len[p]]:
 *   circular: counts[k] = #{ i < Nq : r[(k + i) mod Nr] == q[i] },          k in [0, Nr)
 *   linear  : counts[k] = #{ i < Nq : 0 <= k+i-(Nq-1) < Nr and r[k+i-(Nq-1)] == q[i] },  k in [0, Nr+Nq-1)
 * (ref_query_alignment.py:351-366 builds exactly these windows before calling the Cython scan.)
 * Equality is on the raw bytes.  counts_off: int64[n+1] prefix offsets (filled by prepare). */
int smx_scan_prepare(smx_handle *h,
                     const int64_t *ref_off, const int32_t *ref_len,
                     const int64_t *qry_off, const int32_t *qry_len,
                     int32_t n, int32_t topology, int64_t *counts_off);
int smx_scan_run(smx_handle *h);
int smx_scan_fetch(smx_handle *h, int32_t *counts, int64_t counts_capacity);
int64_t smx_scan_cells(const smx_handle *h);

/* ---- A1: the whole alignment stage in one call ------------------------------------------------
 * Replaces execute_alignment_core / execute_alignment_core_loop
 * (savemoney/post_analysis/post_analysis_core.py:56-87) and RecommendedGrouping.set_distance_matrix /
 * calc_distance (savemoney/pre_survey/pre_survey_core.py:237-271).  For every (query, reference)
 * pair and both strands (forward, reverse complement -- Bio.Seq semantics, IUPAC aware):
 *   A3 scan + A4 threshold/conserved runs (GPU) -> A5/A6 chain search (host threads) -> A7 plan ->
 *   A8 batched DP + traceback (GPU) -> A9-A11 re-clipping, my_special_DP merge, rotation, score.
 * Inputs are upper-case ASCII; ref_sigma[r] = scipy.stats.norm.ppf(1 - 0.1/len(ref r))
 * (ref_query_alignment.py:200-201, computed by the Python host with scipy exactly like the reference).
 * pair_ref/pair_qry: explicit pairs (pre_survey) or both NULL for all (query, reference)
 * combinations in query-major order (post_analysis).  Pair-strand index = 2*pair + strand, which is
 * the order of `result_list` in execute_alignment_core_loop (:85).
 * omit_too_long: 1 for post_analysis (:72-73), 0 for pre_survey (pre_survey_core.py:255-256).
 * n_threads <= 0: use every host core. */
int smx_pipeline_run(smx_handle *h,
                     const uint8_t *ref_seqs, const int64_t *ref_off, const int32_t *ref_len, int32_t n_refs,
                     const double *ref_sigma,
                     const uint8_t *qry_seqs, const int64_t *qry_off, const int32_t *qry_len, int32_t n_qrys,
                     const int32_t *pair_ref, const int32_t *pair_qry, int32_t n_pairs,
                     int32_t open, int32_t extend, int32_t match, int32_t mismatch,
                     int32_t topology, int32_t omit_too_long, int32_t n_threads, int64_t *total_runs);
/* Per pair-strand: score (MyResult.score, 0 for an empty result), new_query_seq_offset (INT32_MIN
 * encodes None), status (0 = aligned, 1 = empty MyResult(), 2 = the reference would raise an
 * assertion on this input), CIGAR runs `len << 4 | op` ('I'=1 'D'=2 'S'=4 'H'=5 '='=7 'X'=8) of
 * MyResult.my_cigar with prefix offsets cigar_off[2*n_pairs + 1]. */
int smx_pipeline_fetch(smx_handle *h, int32_t *score, int32_t *new_query_seq_offset, uint8_t *status,
                       int64_t *cigar_off, uint32_t *cigar_rle, int64_t cigar_capacity);
/* out[0..16]: t_total, t_pool, t_scan, t_select, t_chain, t_dp, t_stitch (host seconds), gpu_scan_ms,
 * gpu_select_ms, gpu_dp_ms, scan_cells, dp_cells, n_dp_problems, n_host_redecisions, n_failed,
 * kernel_launches, n_pair_strands, then out[17..32] = smx_align_kernel_ms ms16 of the DP run,
 * out[33..46] = its cells14, out[47..62] = launches per kernel (same order as ms16), out[63] = select kernel ms,
 * out[64] = chain kernel ms, out[65] / out[66] = cells / problems of the packed classes that ran checkpointed
 * (score-only forward pass, smx_traceback_ck) */
int smx_pipeline_stats(smx_handle *h, double *out, int32_t n);

/* ---- row N1: the alignment operations of MyAlignerBase as one batch -------------------------
 * What MyMSA.exec_chunk_poa (savemoney/modules/msa.py:1249-1355) calls once per read chunk -- collected here over all
 * chunks of a polish round.  Sequences live in the uploaded pool (smx_pool_upload); `ref` is the chunk consensus.
 *   SMX_OPKIND_NW          MyAlignerBase.nw_trace             (ref_query_alignment.py:57-59)   -> = X I D
 *   SMX_OPKIND_SG_QE_DE    my_special_sg_qe_de                (:60-73)   aligned prefix, then S.. H..  (set_soft_clipping "beg")
 *   SMX_OPKIND_SG_QB_DB    my_special_sg_qb_db                (:74-87)   H.. S.., then the aligned suffix  ("end")
 *   SMX_OPKIND_SW          my_special_sw                      (:34-56)   H/S flanks around the local alignment, leading-D quirk
 *   SMX_OPKIND_SPECIAL_DP  my_special_DP(q1, q2, ref)         (:89-163)  sg_qe_de(q1) joined with sg_qb_db(q2) at the best cut;
 *                          q1_len or q2_len may be 0 (that half is "H" * len(ref))
 * Result: one column-run CIGAR per operation (ops = X I D S H), fetched with smx_ops_fetch; status 2 = an input on which
 * the reference itself raises (empty local alignment, failed assertion). */
#define SMX_OPKIND_NW 0
#define SMX_OPKIND_SG_QE_DE 1
#define SMX_OPKIND_SG_QB_DB 2
#define SMX_OPKIND_SW 3
#define SMX_OPKIND_SPECIAL_DP 4
int smx_ops_run(smx_handle *h, const uint8_t *kind, const int64_t *ref_off, const int32_t *ref_len, const int64_t *q1_off,
                const int32_t *q1_len, const int64_t *q2_off, const int32_t *q2_len, int32_t n, int32_t open, int32_t extend,
                int32_t match, int32_t mismatch, int32_t n_threads, int64_t *total_runs);
int smx_ops_fetch(smx_handle *h, int64_t *cigar_off, uint32_t *cigar_rle, int64_t cigar_capacity, uint8_t *status);

/* ---- row N3, first pass: MyMSA.generate_msa (savemoney/modules/msa.py:882-982) -----------------
 * The reads assigned to one plasmid, each with its column CIGAR (run-length, ops = X I D S H) against the same reference,
 * laid out as one multiple alignment.  Inputs are host arrays; the query bases and their Q-scores (int8, one per base)
 * share the offsets.  Output: n_cols columns; smx_msa_fetch delivers ref_seq_aligned [n_cols] ('-' in extra columns) and
 * row-major [n_reads, n_cols] matrices: query_seq_list_aligned (bases, '-' for D / N columns, ' ' for H / O columns),
 * q_scores_list_aligned (-1 in every column without a base) and my_cigar_list_aligned (= X I D S H and the fillers N, O).
 * A CIGAR that does not consume exactly the reference and its query is an error (the reference's final assertions). */
int smx_msa_run(smx_handle *h, const uint8_t *ref, int32_t ref_len, int32_t n_reads, const uint8_t *qry, const int8_t *qscore,
                const int64_t *qry_off, const int32_t *qry_len, const uint32_t *cigar_rle, const int64_t *cigar_off, int64_t *n_cols);
int smx_msa_fetch(smx_handle *h, uint8_t *ref_aligned, uint8_t *seq_aligned, int8_t *qs_aligned, uint8_t *cigar_aligned);

/* ---- row N3, second half: the per-column Bayesian consensus --------------------------------------
 * MyMSA.calculate_consensus_core (savemoney/modules/msa.py:1749-1812) calls, for every column of the multiple alignment,
 * SequenceBasecallQscorePDF.calc_consensus_error_rate (cython_functions/alignment_functions.pyx:93-132): long double
 * products over the reads that vote in the column (column letter not H / O).  This entry point returns that p_list for
 * every column, with and without prior, bit-identical to the reference's x87 arithmetic (the product chains run on the
 * device on a restated 80-bit multiplication, the seven remaining operations per value in the host's long double).
 *   seq / qs / cig   [n_rows][n_cols] matrices as smx_msa_fetch delivers them, or all NULL = the result of the last
 *                    smx_msa_run on this handle (still on the device)
 *   P25[B*5 + c]     P_base_calling_given_true_refseq_dict["B_c"], bases in the order "ATCG-" (msa.py:655,699)
 *   pdf[(B*5+c)*93 + k]  pdf_core["B_c"][k] (msa.py:705)
 *   col_class[n_cols], pn_with / pn_without [n_classes][5]: P_N_dict_dict[ref letter] of msa.py:1813-1845 per column
 *   p_with / p_without [n_cols][5], n_events[n_cols] (0: the reference emits "-" with q-score -1, msa.py:1793-1795)
 * Returns -7 if a voting read holds a character outside ATCG- or a q-score above 92 (undefined behaviour in the reference). */
int smx_consensus_run(smx_handle *h, const uint8_t *seq, const int8_t *qs, const uint8_t *cig, int32_t n_rows, int64_t n_cols,
                      const double *P25, const double *pdf, const int32_t *col_class, const double *pn_with, const double *pn_without,
                      int32_t n_classes, double *p_with, double *p_without, int32_t *n_events);

/* ---- row N2: what consumes the score matrix and the CIGARs without per-result objects -------
 * smx_assign: msa.QueryAssignment (savemoney/modules/msa.py:31-64) on the device, one thread per read.
 *   score / status: [n_reads, 2 * n_refs] as smx_pipeline_fetch delivers them (status != 0 = empty MyResult());
 *   normalized[n_reads, 2 * n_refs] = score / len(ref) in float64, NaN where empty; summary[n_reads, n_refs] = best
 *   strand per plasmid, empty / negative -> 0; assignment[n_reads, 3] = (classified_ref_seq_idx, is_reverse_compliment,
 *   assigned): (-1, -1, -1) without any result, (-1, -1, 0) on tied maxima, else (k / 2, k % 2, best > threshold). */
int smx_assign(smx_handle *h, const int32_t *score, const uint8_t *status, const int32_t *ref_len, int32_t n_reads, int32_t n_refs,
               double score_threshold, int32_t *assignment, double *normalized, double *summary);
/* smx_ir_format: the `# result<idx>(dict)` text blocks of <stem>.intermediate_results.txt
 * (post_analysis/post_analysis_core.py:89-128 over MyResult.to_dict, modules/ref_query_alignment.py:512-518; block
 * syntax modules/my_classes.py:43-74) for pair-strands [0, n) numbered from `first_index`, written from the arrays
 * smx_pipeline_fetch delivers (host threads; needs no device).  out == NULL returns the byte count. */
int64_t smx_ir_format(const int32_t *score, const int32_t *new_query_seq_offset, const uint8_t *status, const int64_t *cigar_off,
                      const uint32_t *cigar_rle, int64_t n, int64_t first_index, int32_t n_threads, char *out, int64_t capacity);

/* Debug aid (compute-sanitizer is not available on every pool): with SMX_GUARD=1 in the environment every problem's
 * direction-bit region is surrounded by 2 x 256 poisoned bytes that are verified after the traceback; returns the number
 * of damaged guard words since the handle was created, -1 when the mode is off. */
int64_t smx_debug_guard_errors(const smx_handle *h);

/* ---- measurement helper: dependency-free int32 / DPX issue-rate microbenchmark ---------------
 * Runs `iters` rounds of independent VIADDMNMX / VIMNMX3 / IADD chains on every SM and returns
 * achieved giga-ops/s (one op = one 32-bit lane instruction) in out_gops[3] = {viaddmax, vimax3, iadd}. */
int smx_measure_int_peak(smx_handle *h, int32_t iters, double *out_gops);

#ifdef __cplusplus
}
#endif
#endif /* SMX_H */
