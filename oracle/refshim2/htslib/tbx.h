/* TEST INFRASTRUCTURE stub: vcf.c:265-266 (tabix index of a .gz output; never reached). */
#ifndef LFB200_STUB2_TBX_H
#define LFB200_STUB2_TBX_H
typedef struct tbx_conf_t { int preset, sc, bc, ec, meta_char, line_skip; } tbx_conf_t;
extern const tbx_conf_t tbx_conf_vcf;
int tbx_index_build(const char *fn, int min_shift, const tbx_conf_t *conf);
#endif
