/* TEST INFRASTRUCTURE stub: kstring_t as htslib lays it out (vcf.c:288 only zero-initialises one). */
#ifndef LFB200_STUB2_KSTRING_H
#define LFB200_STUB2_KSTRING_H
#include <stddef.h>
typedef struct kstring_t { size_t l, m; char *s; } kstring_t;
#endif
