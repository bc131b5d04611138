/* oracle_api.h -- flat C interface shared by the two CPU checkers:
 *
 *   oracle/_ref/libps_ref.so     the UNMODIFIED reference C++ (/root/reference/cpp/*.cpp),
 *                                compiled where it lies, wrapped by oracle/ref_shim.cpp
 *   oracle/_build/libps_oracle.so  an independent CPU restatement (oracle/ps_oracle.cpp)
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in poreseq_b200/ may include, link or load this.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs use it, and only as the checker / the CPU baseline.
 *
 * Variable-length results (mutation lists, sequences, index pairs) come back as text in a
 * caller-supplied buffer, doubles printed with %.17g so they round-trip bit-exactly:
 *   mutation list : one line per mutation  "start\torig\tmut\tscore\n"  ('.' = empty string)
 *   sequence list : one sequence per line
 * Every function returns 0 on success, -1 if the output buffer is too small.
 */
#ifndef ORACLE_API_H_
#define ORACLE_API_H_

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_region
{
    const char*   seq;          /* region sequence, seq_len bases (not NUL-terminated)       */
    int           seq_len;
    int           n_events;
    const int*    lev_off;      /* n_events+1 offsets into the level arrays                  */
    const double* mean;         /* concatenated per-level arrays                             */
    const double* stdv;
    double*       ref_align;    /* in/out (reference updates them in place)                  */
    double*       ref_like;     /* in/out                                                    */
    const double* model;        /* n_events x 4 x 1024: level_mean, level_stdv, sd_mean, sd_stdv */
    const double* trans;        /* n_events x 4: prob_skip, prob_stay, prob_extend, prob_insert  */
    const char* const* ev_seq;  /* n_events per-event "2D" sequences (may be NULL)           */
    double        lik_offset;
    int           scoring_width;
    int           realign_width;
} orc_region;

/* ScoreAlignments (cpp/MakeMutations.cpp:148): scores[n_events]; likes[seq_len] or NULL (accumulated into). */
int orc_score_alignments(orc_region* r, double* scores, double* likes);

/* ScoreMutations (cpp/MakeMutations.cpp:23).  orig/mut: arrays of NUL-terminated strings. */
int orc_score_mutations(orc_region* r, int n, const int* start, const char* const* orig,
                        const char* const* mut, double* scores);

/* FindPointMutations + ScoreMutations (PSAlign.ScorePoints, _poreseqcpp.pyx:278). */
int orc_score_points(orc_region* r, char* out, int cap);

/* MakeMutations (cpp/MakeMutations.cpp:74): applies the scored list, writes the new sequence. */
int orc_make_mutations(orc_region* r, int n, const int* start, const char* const* orig,
                       const char* const* mut, const double* scores,
                       char* seq_out, int cap, int* nbases);

/* PSAlign.Refine (_poreseqcpp.pyx:437): FindPointMutations, ScoreMutations, MakeMutations. */
int orc_refine(orc_region* r, char* seq_out, int cap, int* nbases);

/* FindMutations (cpp/FindMutations.cpp:24) with the given seed sequences. */
int orc_find_mutations(orc_region* r, int n_seeds, const char* const* seeds, char* out, int cap);

/* PSAlign.Mutate loop body (_poreseqcpp.pyx:424-431): reps x (Find, Score, Make). */
int orc_mutate(orc_region* r, int n_seeds, const char* const* seeds, int reps,
               char* seq_out, int cap, int* totbases);

/* ViterbiMutate (cpp/Viterbi.cpp:239).  Caller is responsible for srand() if it wants
 * a reproducible rand() stream (the reference never seeds). */
int orc_viterbi_mutate(orc_region* r, int nkeep, double skip_prob, double stay_prob,
                       double mut_min, double mut_max, char* out, int cap);

/* swfull (cpp/swlib.cpp:211): aligned index pairs (1-based, 0 = gap), score, accuracy. */
int orc_swfull(const char* seq1, const char* seq2, int* inds1, int* inds2, int cap,
               int* n, int* score, double* accuracy);

/* MapAlignments (cpp/EventUtil.cpp:12): remaps r->ref_align onto newseq in place. */
int orc_map_alignments(orc_region* r, const char* newseq);

/* Sequence::populateStates (cpp/Sequence.h:69): states[len-4] (or fewer), returns count. */
int orc_seq_to_states(const char* seq, int len, int* states);

const char* orc_name(void);

#ifdef __cplusplus
}
#endif
#endif
