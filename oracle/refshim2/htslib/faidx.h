/* TEST INFRASTRUCTURE stub (no htslib in this image): the opaque type and the five prototypes
 * lofreq_call.c:1145-1159,1575 names; oracle/call_harness.c defines them as aborting stubs —
 * the fake pileup never reaches main_call(). */
#ifndef LFB200_STUB2_FAIDX_H
#define LFB200_STUB2_FAIDX_H
typedef struct faidx_t faidx_t;
faidx_t *fai_load(const char *fn);
void fai_destroy(faidx_t *fai);
int faidx_nseq(const faidx_t *fai);
const char *faidx_iseq(const faidx_t *fai, int i);
int faidx_seq_len(const faidx_t *fai, const char *seq);
#endif
