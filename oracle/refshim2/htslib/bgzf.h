/* TEST INFRASTRUCTURE stub: opaque BGZF and the calls vcf.c:165-289 makes on its .gz paths
 * (never taken by the harness, which writes plain text). */
#ifndef LFB200_STUB2_BGZF_H
#define LFB200_STUB2_BGZF_H
#include <stdio.h>
#include <stdint.h>
#include <sys/types.h>
#include "kstring.h"
typedef struct BGZF BGZF;
BGZF *bgzf_open(const char *path, const char *mode);
int bgzf_close(BGZF *fp);
int bgzf_flush(BGZF *fp);
ssize_t bgzf_write(BGZF *fp, const void *data, size_t length);
int64_t bgzf_seek(BGZF *fp, int64_t pos, int whence);
int bgzf_getline(BGZF *fp, int delim, kstring_t *str);
#endif
