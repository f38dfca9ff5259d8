/* Test-infrastructure stub: opaque handle only (src/lofreq/vcf.h:34,43). */
#ifndef LFB200_STUB_BGZF_H
#define LFB200_STUB_BGZF_H
#include <stdio.h>
typedef struct BGZF BGZF;
#endif
