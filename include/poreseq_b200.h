/* poreseq_b200.h -- C-ABI of the B200-native PoreSeq scoring path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.  Each entry point
 * replaces one native function the reference's Cython module binds through `cdef extern`
 * (poreseq/_poreseqcpp.pyx:63-83); the citation beside each declaration is the reference
 * interface it stands in for.  INTEGRATION.md shows the Cython stub a PoreSeq maintainer would
 * add to call these instead of the sources under cpp/.
 *
 * Model: a `ps_ctx` owns one CUDA device/stream and its scratch memory; a `ps_region` is the
 * native twin of the reference's `AlignData` (cpp/AlignData.h:24-34): one sequence, its events
 * (per-read level arrays + pore model + transition probabilities) and the alignment parameters.
 * Inputs are caller-owned and copied on entry (the reference copies too, cpp/EventData.h:208-215);
 * outputs are written into caller-allocated arrays or fetched with the getters.
 *
 * Error convention: every int-returning function returns 0 on success and a negative PS_E_* code
 * on failure; ps_last_error() gives the message.  There is NO CPU fallback: without a usable
 * sm_100 device every compute entry point fails with PS_E_CUDA.  (The reference itself never
 * reports errors from C++; degenerate inputs follow its silent conventions -- e.g. a mutation with
 * start > len(sequence) keeps the score -1e-6, cpp/MakeMutations.cpp:46.)
 *
 * Threading: a ctx and its regions must be used from one thread at a time (the reference holds
 * the GIL for the whole call).  CUDA is initialised lazily on the first compute call, so a
 * process may fork before that (poreseq train forks workers, poreseq/cmdline.py:258).
 */
#ifndef PORESEQ_B200_H_
#define PORESEQ_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

#define PS_OK            0
#define PS_E_ARG        -1   /* bad argument / unsupported parameter value                      */
#define PS_E_CUDA       -2   /* CUDA runtime failure or no device                               */
#define PS_E_CAPACITY   -3   /* caller-supplied output buffer too small                         */
#define PS_E_INTERNAL   -4

#define PS_N_STATES   1024   /* cpp/AlignUtil.h:19                                              */

typedef struct ps_ctx ps_ctx;
typedef struct ps_region ps_region;

/* cpp/AlignUtil.h:57-66 AlignParams (defaults lik_offset 4.5, scoring_width 150,
 * realign_width 300, verbose 0). */
typedef struct ps_params
{
    double lik_offset;
    int    scoring_width;
    int    realign_width;
    int    verbose;
} ps_params;

/* One timing record per phase of the last compute call, CUDA-event milliseconds on the ctx
 * stream (bench.py reads these for the roofline; see ps_last_timing). */
#define PS_T_H2D        0
#define PS_T_CENTRES    1
#define PS_T_FORWARD    2
#define PS_T_BACKWARD   3
#define PS_T_BACKTRACE  4
#define PS_T_JOIN       5
#define PS_T_MUTSCORE   6
#define PS_T_REDUCE     7
#define PS_T_D2H        8
#define PS_T_TOTAL      9
#define PS_T_COUNT     10

/* ---- context ---------------------------------------------------------------------------- */
ps_ctx*     ps_create(int device);             /* no reference analogue (the reference is CPU-only) */
void        ps_destroy(ps_ctx* ctx);
const char* ps_last_error(ps_ctx* ctx);        /* ctx may be NULL: last error of a failed ps_create */
const char* ps_version(void);
/* Number of kernels this library launched on ctx since creation / algorithmic DP cells of the
 * last compute call (wide fill cells, narrow mutation cells), for bench.py. */
long long   ps_launch_count(ps_ctx* ctx);
/* Arithmetic of the per-(mutation, event) scan.  EXACT (default): IEEE double in the reference's
 * order, every score bit-identical to cpp/Alignment.cpp.  FAST: all pairs in rebased log-space
 * FP32, then every mutation whose total is not clearly negative (> -tau) is re-scored exactly, so
 * accepted mutations and consensus sequences stay bit-identical while clearly negative scores carry
 * ~1e-6 relative error (BASELINE.json tolerance: 1e-4).  The wide fills and backtrace are always FP64. */
#define PS_PRECISION_EXACT 0
#define PS_PRECISION_FAST  1
int         ps_set_precision(ps_ctx* ctx, int mode);
int         ps_last_timing(ps_ctx* ctx, double* ms /*PS_T_COUNT*/);
int         ps_last_cells(ps_ctx* ctx, double* wide_cells, double* narrow_cells);
/* Bytes the last batch copied host->device and device->host (cudaMemcpyAsync payloads). */
int         ps_last_bytes(ps_ctx* ctx, long long* h2d_bytes, long long* d2h_bytes);

/* ---- region = AlignData (cpp/AlignData.h:24-34) ------------------------------------------ */
ps_region*  ps_region_create(ps_ctx* ctx, const char* bases, int len, const ps_params* params);
void        ps_region_destroy(ps_region* r);
/* EventData::setData + ModelData::setData/setParams (cpp/EventData.h:48-73,208-224), as
 * marshalled by PythonToEvents (poreseq/_poreseqcpp.pyx:99-129).  The four level arrays have n0
 * entries, the four model arrays PS_N_STATES.  seq2d may be NULL. */
int         ps_region_add_event(ps_region* r, int n0,
                                const double* mean, const double* stdv,
                                const double* ref_align, const double* ref_like,
                                const double* level_mean, const double* level_stdv,
                                const double* sd_mean, const double* sd_stdv,
                                int complement, double prob_skip, double prob_stay,
                                double prob_extend, double prob_insert, const char* seq2d);
/* The same for all events of a region in one call (the whole PythonToEvents loop,
 * poreseq/_poreseqcpp.pyx:99-129): the level arrays of the events are concatenated (sum of n0[e]
 * entries each), models is a table of n_models x 4 x PS_N_STATES doubles (level_mean, level_stdv,
 * sd_mean, sd_stdv per model), probs n_models x 4 (skip, stay, extend, insert), model_index[e]
 * picks the event's model.  complement and seq2d may be NULL. */
int         ps_region_add_events(ps_region* r, int n_events, const int* n0,
                                 const double* mean, const double* stdv,
                                 const double* ref_align, const double* ref_like,
                                 const int* model_index, int n_models, const double* models,
                                 const double* probs, const int* complement, const char* const* seq2d);
/* Many regions in one call: out[k] = ps_region_create(bases, len, params) + ps_region_add_events(...) for every
 * descriptor, built on the library's host worker threads.  On failure nothing is returned (out[] all NULL). */
typedef struct ps_region_desc
{
    const char*   bases;       int len;
    ps_params     params;
    int           n_events;    const int* n0;
    const double* mean;        const double* stdv;
    const double* ref_align;   const double* ref_like;
    const int*    model_index; int n_models;
    const double* models;      const double* probs;
    const int*    complement;  const char* const* seq2d;
} ps_region_desc;
int         ps_regions_create(ps_ctx* ctx, int n_regions, const ps_region_desc* desc, ps_region** out);
/* Releases a batch of regions in one call (the frees run on the library's worker threads); NULL entries are skipped. */
void        ps_regions_destroy(ps_region* const* regions, int n_regions);
int         ps_region_set_params(ps_region* r, const ps_params* params);
int         ps_region_num_events(ps_region* r);
int         ps_region_sequence_length(ps_region* r);
/* data.sequence.bases / data.events[e].ref_align, ref_like as read back by the boundary
 * (poreseq/_poreseqcpp.pyx:131-137, 374, 433, 470). */
int         ps_region_get_sequence(ps_region* r, char* out, int cap);
int         ps_region_get_event_align(ps_region* r, int e, double* ref_align, double* ref_like);

/* ---- scoring ------------------------------------------------------------------------------ */
/* vector<double> ScoreAlignments(AlignData&, double* likes)   cpp/Mutations.h:24,
 * cpp/MakeMutations.cpp:148-195.  scores[n_events]; likes[len(sequence)] is accumulated into
 * when non-NULL.  Realigns every event in place. */
int         ps_score_alignments(ps_region* r, double* scores, double* likes);
/* PSAlign.ScoreEvents   poreseq/_poreseqcpp.pyx:263-276 (= ScoreAlignments(data, NULL) whose realignment is dropped,
 * pyx:273-276): scores[n_events], the region's events keep their alignments.  PS_PRECISION_FAST: score-only log-space
 * FP32 fill (k_score_f32, nothing stored), scores within 1e-4 relative; PS_PRECISION_EXACT: the FP64 fill, bit-identical.
 * _batch: the events of n DISTINCT regions of ONE context in one launch sequence, scores concatenated in region order
 * (like every batched entry point: PS_E_ARG when a handle appears twice). */
int         ps_score_events(ps_region* r, double* scores);
int         ps_score_events_batch(ps_region* const* regions, int n_regions, double* scores);
/* vector<MutScore> ScoreMutations(AlignData&, const vector<MutInfo>&)   cpp/Mutations.h:23,
 * cpp/MakeMutations.cpp:23-69.  orig[i]/mut[i] are NUL-terminated. */
int         ps_score_mutations(ps_region* r, int n, const int* start, const char* const* orig,
                               const char* const* mut, double* scores);
/* Same, but the per-mutation sums start at 0 instead of -1e-6: the partial sum over THIS region's
 * events, for callers that split one region's events across GPUs and all-reduce the partials
 * (the reference's comment at cpp/MakeMutations.cpp:19-22 anticipates exactly that split). */
int         ps_score_mutations_partial(ps_region* r, int n, const int* start, const char* const* orig,
                                       const char* const* mut, double* partial);

/* ---- the consensus loop ----------------------------------------------------------------------
 * poreseq/Mutate.py:47-99 below the boundary: Mutate('self', reps), then up to `reps` rounds of (Mutate('viterbi'),
 * Refine at point_width) until Refine changes nothing.  The handle is the PSAlign object of the loop: its sequence and
 * its events' alignments are what pa.sequence / pa.events hold afterwards (end_trim, Mutate.py:84-88, is left to the
 * caller: it is a slice of the returned string).  Regions with fewer than 5 events are left untouched (Mutate.py:50-53).
 * ViterbiMutate draws from a rand() stream of the region's own that starts at glibc's default seed -- what the
 * reference's one process per region sees (cpp/Viterbi.cpp:108; the stream is never seeded).
 * *n_stages: stages run; ps_region_get_stage(k) gives the stage's name ("mutate_self", "mutate_viterbi_0", "refine_0",
 * ...), the sequence after it and the bases it changed (returns the sequence length; buffers may be NULL).
 * ps_consensus_batch: n regions, `in_flight` of them side by side on ctx's device (host threads and streams of the
 * library: the reference's one-process-per-region scaling, README.md:48-54, inside one process). */
int         ps_consensus(ps_region* r, int reps, int point_width, int* n_stages);
int         ps_consensus_batch(ps_ctx* ctx, ps_region* const* regions, int n_regions, int reps, int point_width,
                               int in_flight);
int         ps_region_num_stages(ps_region* r);
int         ps_region_get_stage(ps_region* r, int k, char* name, int name_cap, char* seq, int seq_cap, int* nbases);

/* ---- one region's events split across the GPUs of a box ---------------------------------------
 * score[m] = -1e-6 + sum over events of delta(m, e)   (cpp/MakeMutations.cpp:19-22, 38-52; the caller is
 * poreseq/Variant.py:71-76 on deep coverage).  Every rank (one process / context per GPU) holds a contiguous block of
 * the region's events in a handle of its own, rank 0 the first block; ps_score_mutations_sharded scores all mutations
 * against the local block and combines the sums over NCCL (NVLink / NVSwitch) on the context's stream, inside the
 * library: no host round trip, no torch.  Every rank passes the same mutations and receives the complete scores.
 *   ordered != 0  rank r continues the running sums of the ranks before it in event order: bit-identical to the
 *                 single-GPU call and to the reference (n_ranks - 1 small messages in sequence, then a broadcast);
 *   ordered == 0  one ncclAllReduce(sum, float64, n) of partial sums: scores agree to ~1e-16 relative.
 * PS_PRECISION_FAST works as on one GPU (FP32 scan, mutations whose TOTAL is above the threshold re-scored exactly).
 * ps_comm_unique_id: rank 0 makes the 128-byte NCCL id, the caller carries it to the other ranks by its own means
 * (MPI, a file, torch.distributed ...); ps_comm_init is collective over the n_ranks contexts.  NCCL is loaded with
 * dlopen("libnccl.so.2") at that point (PORESEQ_B200_NCCL overrides the path). */
#define PS_COMM_ID_BYTES 128
int         ps_comm_unique_id(void* id, int bytes);
int         ps_comm_init(ps_ctx* ctx, const void* id, int bytes, int rank, int n_ranks, int ordered);
int         ps_comm_destroy(ps_ctx* ctx);
int         ps_comm_rank(ps_ctx* ctx, int* rank, int* n_ranks);
int         ps_score_mutations_sharded(ps_region* r, int n, const int* start, const char* const* orig,
                                       const char* const* mut, double* scores);
/* vector<MutInfo> FindPointMutations(AlignData&)   cpp/Mutations.h:19, cpp/FindMutations.cpp:191-234.
 * Writes up to cap single-base edits (orig/mut as one char, 0 = empty); *n = 8 per state. */
int         ps_find_point_mutations(ps_region* r, int cap, int* n, int* start, char* orig, char* mut);
/* FindPointMutations + ScoreMutations in one call (PSAlign.ScorePoints, _poreseqcpp.pyx:278-308). */
int         ps_score_points(ps_region* r, int cap, int* n, int* start, char* orig, char* mut, double* scores);
/* int MakeMutations(AlignData&, vector<MutScore>)   cpp/Mutations.h:22, cpp/MakeMutations.cpp:74-146. */
int         ps_make_mutations(ps_region* r, int n, const int* start, const char* const* orig,
                              const char* const* mut, const double* scores, int* nbases);
/* PSAlign.Refine body (_poreseqcpp.pyx:463-469): FindPointMutations, ScoreMutations, MakeMutations. */
int         ps_refine(ps_region* r, int* nbases);

/* Batched form of ps_score_points over independent regions (same ctx): one launch sequence for
 * all of them.  n_out[k] edits are written for region k at offset off_out[k] of the flat output
 * arrays (cap entries in total).  No reference analogue: the reference scales by running one
 * process per region (README.md:48-54). */
int         ps_score_points_batch(ps_region* const* regions, int n_regions, int cap,
                                  int* n_out, long long* off_out,
                                  int* start, char* orig, char* mut, double* scores);

/* Asynchronous form: _begin stages the batch and enqueues copies + kernels on the context's stream
 * and returns; _end waits and delivers the scores.  One batch in flight per context; the regions
 * must stay alive and untouched in between.  Two contexts on one device let the host stage batch
 * k+1 while the GPU works on batch k. */
int         ps_score_points_batch_begin(ps_region* const* regions, int n_regions, int cap,
                                        int* n_out, long long* off_out, int* start, char* orig, char* mut);
int         ps_score_points_batch_end(ps_ctx* ctx, double* scores);
/* PSAlign.ScorePoints of n regions straight from the caller's arrays -- poreseq/_poreseqcpp.pyx:278-308: PythonToAlignData,
 * FindPointMutations, ScoreMutations, the scores; the realignment is dropped there (no UpdatePythonEvents), so nothing but
 * the scores comes back here either.  No region handles are made: the level arrays of `desc` are read where they lie (they
 * must stay valid until _end returns), nothing is copied, no alignment travels back from the device.  Outputs as in
 * ps_score_points_batch.  _begin / _end: the asynchronous halves, one batch in flight per context. */
int         ps_score_points_direct(ps_ctx* ctx, int n_regions, const ps_region_desc* desc, int cap, int* n_out,
                                   long long* off_out, int* start, char* orig, char* mut, double* scores);
int         ps_score_points_direct_begin(ps_ctx* ctx, int n_regions, const ps_region_desc* desc, int cap, int* n_out,
                                         long long* off_out, int* start, char* orig, char* mut);
int         ps_score_points_direct_end(ps_ctx* ctx, double* scores);

/* ---- candidate discovery and the consensus iteration ------------------------------------------- */
/* vector<MutInfo> FindMutations(AlignData&, const vector<Sequence>&)   cpp/Mutations.h:18,
 * cpp/FindMutations.cpp:24-186.  The result is held by the region; fetch entry i with
 * ps_get_found_mutation (sizes first with ps_found_mutation_sizes). */
int         ps_find_mutations(ps_region* r, int n_seeds, const char* const* seeds, int* n_found);
/* The host-only second half of FindMutations (cpp/FindMutations.cpp:51-186) for callers that already hold the per-base
 * likelihood profiles ScoreAlignments accumulates (cpp/MakeMutations.cpp:168-189) -- e.g. summed over event shards on
 * several GPUs: swfull of the region's sequence against every seed, CUSUM of the profile differences along the alignments,
 * greedy peak picking.  base_profile has len(sequence) entries, seed_profiles[s] strlen(seeds[s]).  The result is read
 * like that of ps_find_mutations. */
/* What the band planner sees of one event (host only, no device work): ref_index as updaterefs leaves it
 * (cpp/EventData.h:110-169; n0 doubles, may be null) and getrefstate(c) = std::lower_bound(ref_index, c) for the columns
 * c = 0 .. n_cols-1 (cpp/EventData.h:172-183; 1 everywhere when the event carries no alignment, cpp/Alignment.cpp:129-132).
 * *monotone: the centres never go backwards (the wavefront schedule applies).  An event with a single aligned level has
 * a 0/0 slope there and a ref_index of NaNs around that level: the centres are whatever the binary search's probes give. */
int         ps_band_centres(ps_region* r, int event, int n_cols, int* centres, double* ref_index_out, int* ri_empty, int* monotone);
int         ps_pick_candidates(ps_region* r, int n_seeds, const char* const* seeds, const double* base_profile,
                               const double* const* seed_profiles, int* n_found);
int         ps_found_mutation_sizes(ps_region* r, int i, int* n_orig, int* n_mut);
int         ps_get_found_mutation(ps_region* r, int i, int* start, char* orig, int orig_cap, char* mut, int mut_cap);
/* Loop body of PSAlign.Mutate (poreseq/_poreseqcpp.pyx:424-431): reps x (FindMutations,
 * ScoreMutations, MakeMutations), stopping when a round changes nothing. */
int         ps_mutate(ps_region* r, int n_seeds, const char* const* seeds, int reps, int* totbases);
/* The host half of ViterbiMutate alone (no device work; cpp/Viterbi.cpp:262-325): the positions the loop keeps and how
 * many reads sit on each (getrefstates, cpp/EventData.h:187-204: std::find over ref_index, whose extrapolated ends can
 * match positions beyond every read's refend), after the reference's skip / stop rule (nlik <= 0.2 * reads spanning the
 * position: skip, or stop when no read spans it).  PS_E_ARG when an event carries no alignment, like ps_viterbi_mutate;
 * PS_E_CAPACITY (with *n_positions set) when cap is too small. */
int         ps_viterbi_positions(ps_region* r, int cap, int* positions, int* reads_here, int* n_positions);

/* vector<Sequence> ViterbiMutate(vector<EventData>&, int nkeep, double skip, double stay,
 * double mut_min, double mut_max, bool verbose)   cpp/Viterbi.h:67-68, cpp/Viterbi.cpp:239-426.
 * nkeep == 0: the best path; otherwise nkeep forward-weighted samples drawn with libc rand()
 * (never seeded by the reference; call srand() first for a reproducible stream).  The sequences
 * are held by the region; ps_get_viterbi_sequence(r, i, NULL, 0) returns the length of entry i. */
int         ps_viterbi_mutate(ps_region* r, int nkeep, double skip_prob, double stay_prob,
                              double mut_min, double mut_max, int* n_seqs);
int         ps_get_viterbi_sequence(ps_region* r, int i, char* out, int cap);
/* SWAlignment MapAlignments(AlignData&, const Sequence&)   cpp/EventUtil.h:17, cpp/EventUtil.cpp:12-55:
 * swfull + fillinds, then every level's ref_align is carried over to newseq. */
int         ps_map_alignments(ps_region* r, const char* newseq);

/* ---- event-pack files ------------------------------------------------------------------------- */
/* A flat binary file of many regions (layout: poreseq_b200/eventpack.py) that is memory-mapped and marshalled without
 * Python objects.  Stands in for the per-region loading of LoadAlignedEvents (poreseq/LoadData.py:10-65: h5py + pysam
 * into PSEvent objects, then PythonToAlignData poreseq/_poreseqcpp.pyx:139-153 on every call).  Host only. */
typedef struct ps_pack ps_pack;
ps_pack*    ps_pack_open(const char* path);        /* NULL on failure: ps_last_error(NULL) has the reason          */
void        ps_pack_close(ps_pack* pack);
int         ps_pack_num_regions(ps_pack* pack);
/* Region k as a descriptor whose pointers are views into the mapping (valid until ps_pack_close; seq2d is NULL).
 * width_key names the parameter that overrides scoring_width ("point_width", pyx:293,361,465) or is NULL. */
int         ps_pack_region_desc(ps_pack* pack, int k, const char* width_key, ps_region_desc* out);
/* A named parameter of region k (the pack keeps every key of the .conf); PS_E_ARG when absent. */
int         ps_pack_region_param(ps_pack* pack, int k, const char* name, double* value);
/* The 2D read sequence of event e of region k (PSEvent.sequence, the seeds of PSAlign.Mutate('self'), pyx:412-414):
 * a view of *len bytes, not NUL-terminated. */
int         ps_pack_event_sequence(ps_pack* pack, int k, int e, const char** seq, int* len);
/* ps_regions_create over regions [first, first + count) of the pack, straight from the mapping. */
int         ps_pack_regions_create(ps_ctx* ctx, ps_pack* pack, int first, int count, const char* width_key, ps_region** out);

/* ---- helpers that stay on the host ---------------------------------------------------------- */
/* Sequence::populateStates (cpp/Sequence.h:69-100); returns the number of states written. */
int         ps_seq_to_states(const char* seq, int len, int* states);
/* SWAlignment swfull(const string&, const string&)   cpp/swlib.h:36, cpp/swlib.cpp:211-340.
 * Aligned index pairs (1-based, 0 = gap) into inds1/inds2 (cap entries), count in *n. */
int         ps_swfull(const char* seq1, const char* seq2, int* inds1, int* inds2, int cap,
                      int* n, int* score, double* accuracy);

/* The same alignment computed on the GPU (ps_sw.cu: int32 scores, one byte of traceback per cell,
 * anti-diagonal wavefront, the reference's tie order), for sequences of 1..16384 bases.  FindMutations
 * uses the batched form of this internally for its seed realignments (cpp/EventUtil.cpp:16). */
int         ps_swfull_device(ps_ctx* ctx, const char* seq1, const char* seq2, int* inds1, int* inds2, int cap,
                             int* n, int* score, double* accuracy);

#ifdef __cplusplus
}
#endif
#endif
