/* CPU oracle (TEST INFRASTRUCTURE ONLY -- never linked or imported by the product path): restatement of
 * SequenceBasecallQscorePDF.calc_consensus_error_rate of the reference
 * (savemoney/modules/cython_functions/alignment_functions.pyx:93-132) in plain C with the same `long double` arithmetic
 * (x86-64: x87 extended precision), operation for operation.
 *
 *   P25[B * 5 + c]          = P_base_calling_given_true_refseq_dict["B_c"]   bases in the order "ATCG-" (msa.py:655)
 *   pdf[(B * 5 + c) * 93 + k] = pdf_core["B_c"][k]                            (msa.py:705: 42 values, 50 zeros, the q = -1 value)
 *   called[n], q[n]         = the column's voting reads in row order (query_list, q_score_list of msa.py:1759-1764)
 *   PN5[b]                  = P_N_dict[b]
 *   p5[B]                   = p_list[B] as the double Cython hands back to Python
 * Pinned by tests/golden/consensus_v1.json.gz (p_lists recorded from the reference's own compiled class).
 * Returns 0, or -1 if a character outside ATCG- or a q-score above 92 occurs (the reference indexes out of bounds there). */
#include <stdint.h>

static int base_index(char c)
{
    switch (c) { case 'A': return 0; case 'T': return 1; case 'C': return 2; case 'G': return 3; case '-': return 4; default: return -1; }
}

int pp_consensus_error_rate(const double *P25, const double *pdf, int n, const char *called, const int32_t *q, const double *PN5, double *p5,
                            long double *scratch /* n values */)
{
    for (int i = 0; i < n; ++i) if (base_index(called[i]) < 0 || q[i] > 92) return -1;
    for (int B = 0; B < 5; ++B) {                                   /* :104  for base_idx in range(self.N_bases) */
        long double *bunshi_list = scratch;
        for (int i = 0; i < n; ++i) {                               /* :110-115 */
            const int key = B * 5 + base_index(called[i]);
            if (q[i] >= 0) bunshi_list[i] = (long double)P25[key] * (long double)pdf[key * 93 + q[i]];
            else bunshi_list[i] = (long double)P25[key];
        }
        const long double bunshi_P_N = (long double)PN5[B];         /* :116 */
        long double bunbo_bunshi_sum = 0;                            /* :119 */
        for (int b = 0; b < 5; ++b) {                               /* :120  for base in self.bases */
            long double val = (long double)PN5[b] / bunshi_P_N;     /* :122 */
            for (int i = 0; i < n; ++i) {                           /* :123-128 */
                const int key = b * 5 + base_index(called[i]);
                if (q[i] >= 0) val *= ((long double)P25[key] * (long double)pdf[key * 93 + q[i]]) / bunshi_list[i];
                else val *= ((long double)P25[key]) / bunshi_list[i];
            }
            bunbo_bunshi_sum += val;                                /* :129 */
        }
        p5[B] = (double)(1 - 1 / bunbo_bunshi_sum);                 /* :130, returned to Python as float */
    }
    return 0;
}
