/* TEST INFRASTRUCTURE — batch driver around the reference's unmodified kpa_ext_glocal() (src/lofreq/kprobaln_ext.c:80-277,
 * compiled where it lies by oracle/Makefile into oracle/_ref/libkparef.so): the banded glocal profile HMM behind BAQ
 * (bam_md_ext.c:407 calls it once per read).  One call per read, exactly as bam_prob_realn_core_ext does, with pd == NULL. */
#include <stdint.h>
#include <stddef.h>
#include "kprobaln_ext.h"

/* reads: ref / query are 0..4 codes, CSR offsets (n + 1); qual may be NULL (the reference then assumes Q30) */
int lfref_kpa_glocal_batch(long long n, const uint8_t *ref, const long long *ref_off, const uint8_t *query, const long long *qry_off,
                           const uint8_t *qual, float d, float e, int bw, int *state, uint8_t *q, int *pr)
{
    kpa_ext_par_t par;
    par.d = d; par.e = e; par.bw = bw;
    for (long long r = 0; r < n; ++r) {
        const int l_ref = (int)(ref_off[r + 1] - ref_off[r]), l_query = (int)(qry_off[r + 1] - qry_off[r]);
        int dummy_bw = 0;
        const int p = kpa_ext_glocal(ref + ref_off[r], l_ref, query + qry_off[r], l_query, qual ? qual + qry_off[r] : NULL, &par,
                                     state + qry_off[r], q + qry_off[r], NULL, &dummy_bw);
        if (pr) pr[r] = p;
    }
    return 0;
}
