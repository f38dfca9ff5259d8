/* Plain-C host side over the C ABI: what LoFreq's main_call does around its per-column callback
 * (lofreq_call.c:1472-1489, plp.c:1406-1447), with a fake pileup instead of htslib.
 *
 *   mpileup(conf, plp_proc_func, plp_proc_conf, ...)   ->  fake_mpileup(): one column at a time, pointer valid only
 *                                                          during the call (plp.c:1440-1445)
 *   call_vars(const plp_col_t *, void *)               ->  gpu_call_vars(): copies the column into a builder
 *   report_var(...) per significant allele             ->  on_site(): prints CHROM POS REF ALT QUAL DP AF HQA
 *   "Number of substitution tests performed: %lld"     ->  printed from conf.num_snv_tests (lofreq_call.c:1562)
 *
 * The column struct below carries only the fields of plp_col_t (plp.h:73-145) the SNV path reads; inside LoFreq the
 * adapter takes them from the real plp_col_t.
 *
 *   gcc -std=c99 -O2 -Iinclude examples/call_adapter.c -Llofreq_b200/lib -llofreq_b200 -Wl,-rpath,$PWD/lofreq_b200/lib -o call_adapter
 *   ./call_adapter [columns] [depth]
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "lofreq_b200.h"

typedef struct { int *data; int n; } varray;             /* int_varray_t, utils.h:60-66 */

typedef struct {                                          /* the part of plp_col_t the path reads */
    const char *target;
    int pos;
    char ref_base;
    char cons_base0;
    int coverage_plp, num_bases;
    varray base_quals[4], map_quals[4], baq_quals[4];
} fake_col;

typedef struct {
    lfb200_builder *bld;
    lfb200_conf_t conf;
    const char *target;
    long long n_reported;
} gpu_call;

typedef void (*col_func)(const fake_col *, void *);       /* plp_proc_func, plp.h:159-163 */

static unsigned long long rng_state = 20261017ull;
static unsigned rnd(void)
{
    rng_state = rng_state * 6364136223846793005ull + 1442695040888963407ull;
    return (unsigned)(rng_state >> 33);
}

/* the callback of a site: the alt slots are A,C,G,T minus ref (snpcaller.c:383-397) */
static void on_site(const lfb200_site_t *s, long long pos, char ref_base, int coverage_plp, void *user)
{
    gpu_call *g = (gpu_call *)user;
    int i, k = 0;
    for (i = 0; i < 4; i++) {
        const char alt = "ACGT"[i];
        if (alt == ref_base) continue;
        if (s->called[k]) {                                /* pvalue * bonf < sig, lofreq_call.c:832 */
            printf("%s\t%lld\t.\t%c\t%c\t%d\t.\tDP=%d;AF=%f;HQA=%d\n", g->target, pos + 1, ref_base, alt, s->qual[k],
                   coverage_plp, s->alt_raw_count[k] / (float)coverage_plp, s->alt_count[k]);
            g->n_reported++;
        }
        k++;
    }
}

static void gpu_call_vars(const fake_col *p, void *confp)
{
    gpu_call *g = (gpu_call *)confp;
    const int *bq[4], *mq[4], *baq[4];
    int n[4], i;
    if (p->ref_base == 'N') return;                                    /* lofreq_call.c:892 */
    if (p->cons_base0 == '+' || p->cons_base0 == '-') return;          /* lofreq_call.c:928-930 */
    for (i = 0; i < 4; i++) {
        n[i] = p->base_quals[i].n;
        bq[i] = p->base_quals[i].data;
        mq[i] = p->map_quals[i].n ? p->map_quals[i].data : NULL;
        baq[i] = p->baq_quals[i].n ? p->baq_quals[i].data : NULL;
    }
    if (lfb200_builder_add_column(g->bld, p->pos, p->ref_base, p->coverage_plp, p->num_bases, bq, mq, baq, NULL, n)) {
        fprintf(stderr, "FATAL: %s\n", lfb200_last_error());
        exit(1);
    }
}

/* one column at a time; every 97th position carries a variant at 2-30 % */
static int fake_mpileup(col_func func, void *func_conf, int n_cols, int depth)
{
    int *buf = (int *)malloc(sizeof(int) * 3 * (size_t)depth);
    int c, r, b;
    if (!buf) return 1;
    for (c = 0; c < n_cols; c++) {
        fake_col col;
        int cnt[4] = {0, 0, 0, 0}, fill[4], off[4];
        const int ref = c & 3, var = (c % 97 == 5), alt = (ref + 1 + c % 3) & 3;
        const int n_alt = var ? depth * (2 + c % 29) / 100 : 0;
        unsigned char *base = (unsigned char *)malloc((size_t)depth);
        memset(&col, 0, sizeof(col));
        for (r = 0; r < depth; r++) {
            b = r < n_alt ? alt : ref;
            if (rnd() % 1000 == 0) b = (b + 1 + rnd() % 3) & 3;        /* a sequencing error at about Q30 */
            base[r] = (unsigned char)b;
            cnt[b]++;
        }
        off[0] = 0;
        for (b = 1; b < 4; b++) off[b] = off[b - 1] + cnt[b - 1];
        memcpy(fill, off, sizeof(fill));
        for (r = 0; r < depth; r++) {                                  /* compile_plp_col pushes bq / mq / baq together, plp.c:954-975 */
            const int at = fill[base[r]]++;
            buf[at] = 28 + (int)(rnd() % 8);
            buf[depth + at] = 60;
            buf[2 * depth + at] = 40 + (int)(rnd() % 20);
        }
        for (b = 0; b < 4; b++) {
            col.base_quals[b].data = buf + off[b];             col.base_quals[b].n = cnt[b];
            col.map_quals[b].data = buf + depth + off[b];      col.map_quals[b].n = cnt[b];
            col.baq_quals[b].data = buf + 2 * depth + off[b];  col.baq_quals[b].n = cnt[b];
        }
        col.target = "chr1";
        col.pos = c;
        col.ref_base = "ACGT"[ref];
        col.cons_base0 = "ACGT"[ref];
        col.coverage_plp = col.num_bases = depth;
        (*func)(&col, func_conf);                                      /* plp.c:1443; the column is gone after this (plp.c:1445) */
        free(base);
    }
    free(buf);
    return 0;
}

int main(int argc, char **argv)
{
    const int n_cols = argc > 1 ? atoi(argv[1]) : 20000, depth = argc > 2 ? atoi(argv[2]) : 400;
    lfb200_ctx *ctx = NULL;
    gpu_call g;
    memset(&g, 0, sizeof(g));
    g.target = "chr1";
    if (lfb200_create(&ctx, 0)) {
        fprintf(stderr, "FATAL: %s\n", lfb200_last_error());           /* no CUDA device: there is no CPU path */
        return 1;
    }
    lfb200_init_conf(&g.conf);                                         /* init_varcall_conf defaults, dynamic Bonferroni */
    if (lfb200_builder_create(&g.bld, ctx, &g.conf, 8192, on_site, &g)) {
        fprintf(stderr, "FATAL: %s\n", lfb200_last_error());
        return 1;
    }
    if (fake_mpileup(gpu_call_vars, &g, n_cols, depth)) return 1;
    if (lfb200_builder_flush(g.bld)) {                                 /* the single extra call after mpileup() returns */
        fprintf(stderr, "FATAL: %s\n", lfb200_last_error());
        return 1;
    }
    fprintf(stderr, "Number of substitution tests performed: %lld\n", g.conf.num_snv_tests);   /* lofreq_call.c:1562 */
    fprintf(stderr, "bonf_subst = %lld, %lld variants reported\n", g.conf.bonf_subst, g.n_reported);
    lfb200_builder_destroy(g.bld);
    lfb200_destroy(ctx);
    return 0;
}
