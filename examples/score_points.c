/* Plain-C caller of the C-ABI (include/poreseq_b200.h): what a non-Python front end links against.
 *
 *     gcc -std=c99 -Iinclude examples/score_points.c -Lporeseq_b200 -lporeseq_b200 -Wl,-rpath,$PWD/poreseq_b200 -o score_points
 *     ./score_points run.psep            (an event pack written by poreseq_b200/eventpack.py)
 *
 * Opens the pack, builds all its regions straight from the mapping, runs the full single-base scan
 * (FindPointMutations + ScoreMutations, PSAlign.ScorePoints) over them as one batch and prints the edits that
 * improve the likelihood.  Without a B200 the scan fails with PS_E_CUDA ("no CPU fallback") and the program says so;
 * everything before it is host code. */
#include <stdio.h>
#include <stdlib.h>

#include "poreseq_b200.h"

int main(int argc, char** argv)
{
    if (argc < 2) { fprintf(stderr, "usage: %s pack.psep\n", argv[0]); return 2; }
    ps_pack* pack = ps_pack_open(argv[1]);
    if (!pack) { fprintf(stderr, "%s\n", ps_last_error(NULL)); return 1; }
    const int n = ps_pack_num_regions(pack);
    ps_ctx* ctx = ps_create(0);
    ps_region** regs = (ps_region**)calloc((size_t)n, sizeof *regs);
    int rc = ps_pack_regions_create(ctx, pack, 0, n, "point_width", regs);
    if (rc) { fprintf(stderr, "marshalling failed (%d): %s\n", rc, ps_last_error(ctx)); return 1; }
    long long cap = 0;
    for (int k = 0; k < n; k++) cap += 9LL * ps_region_sequence_length(regs[k]);
    int* n_out = (int*)calloc((size_t)n, sizeof *n_out);
    long long* off = (long long*)calloc((size_t)n, sizeof *off);
    int* start = (int*)malloc((size_t)cap * sizeof *start);
    char* orig = (char*)malloc((size_t)cap);
    char* mut = (char*)malloc((size_t)cap);
    double* score = (double*)malloc((size_t)cap * sizeof *score);
    printf("%d regions, room for %lld point mutations\n", n, cap);
    rc = ps_score_points_batch(regs, n, (int)cap, n_out, off, start, orig, mut, score);
    if (rc == PS_E_CUDA) printf("no GPU here: %s\n", ps_last_error(ctx));
    else if (rc) { fprintf(stderr, "scan failed (%d): %s\n", rc, ps_last_error(ctx)); return 1; }
    else
        for (int k = 0; k < n; k++)
            for (long long i = off[k]; i < off[k] + n_out[k]; i++)
                if (score[i] > 0)
                    printf("region %d\t%d\t%c\t%c\t%.6f\n", k, start[i], orig[i] ? orig[i] : '-', mut[i] ? mut[i] : '-', score[i]);
    ps_regions_destroy(regs, n);
    ps_pack_close(pack);
    ps_destroy(ctx);
    free(regs); free(n_out); free(off); free(start); free(orig); free(mut); free(score);
    return 0;
}
