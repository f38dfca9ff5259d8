/* Test-infrastructure stub: opaque handle only. The reference's plp.h
 * (src/lofreq/plp.h:32,66) needs the type name and nothing else on the
 * SNV-test path; htslib itself is not vendored and not present here. */
#ifndef LFB200_STUB_FAIDX_H
#define LFB200_STUB_FAIDX_H
typedef struct faidx_t faidx_t;
#endif
