/*
 * fqd_b200.h -- C ABI of the B200-native clustering path of fastqdedup.
 *
 * Plain C, caller-owned buffers (pointer + size), int status codes and a thread-local
 * error string; no Python, no torch types, no C++ exceptions cross this boundary.  The
 * library is libfqd_b200.so (fastqdedup_b200/csrc); the CPython extension modules
 * fastqdedup_b200._trie/_distance/_fastq are thin shims over it.
 *
 * Every entry point cites the reference interface it replaces (paths relative to
 * /root/reference).  There is no CPU fallback: without a usable CUDA device every compute
 * call returns FQD_ERR_CUDA.
 */
#ifndef FQD_B200_H
#define FQD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* status codes; the shims map them to the exception types the reference raises */
#define FQD_OK 0
#define FQD_ERR_ARG 1          /* ValueError                                             */
#define FQD_ERR_PHRED 2        /* ValueError "Character %c outside of valid phred range"
                                  (src/fastqdedup/_fastqmodule.c:65-70)                  */
#define FQD_ERR_CUDA 3         /* RuntimeError: no device / CUDA failure                 */
#define FQD_ERR_NOMEM 4        /* MemoryError                                            */
#define FQD_ERR_LOOKUP 5       /* LookupError "No sequences left in Trie."
                                  (src/fastqdedup/_triemodule.c:794-797)                 */
#define FQD_ERR_UNSUPPORTED 6  /* key longer / alphabet larger than this build supports  */
#define FQD_ERR_NCCL 7
#define FQD_ERR_FASTQ 8        /* dnaio.FastqFormatError: malformed record, inputs not in sync
                                  (src/fastqdedup/__init__.py:181-185)                     */
#define FQD_ERR_IO 9           /* OSError: a file cannot be opened / read / written         */

#define FQD_METHOD_HIGHEST_COUNT 0 /* cluster_dissection_highest_count, __init__.py:94-102 */
#define FQD_METHOD_ADJACENCY 1     /* cluster_dissection_adjacency,     __init__.py:105-122 */
#define FQD_METHOD_DIRECTIONAL 2   /* cluster_dissection_directional,   __init__.py:60-91  */

#define FQD_MEM_HOST 0
#define FQD_MEM_DEVICE 1

#define FQD_DEFAULT_PHRED_OFFSET 33 /* src/fastqdedup/_fastqmodule.c:22 */

typedef struct fqd_context fqd_context;
typedef struct fqd_trie fqd_trie;

/* Message of the last failing call on this thread ("" when none). */
const char *fqd_last_error(void);

/* Number of usable CUDA devices (0 when there is none or the driver is missing). */
int fqd_device_count(void);

/* One context per GPU: owns a stream, a memory pool and the result of the last job. */
int fqd_context_create(int device_ordinal, fqd_context **out);
void fqd_context_destroy(fqd_context *ctx);
/* Process-wide context created on first use (device = $FQD_DEVICE, else $LOCAL_RANK modulo
 * the device count, else 0); what the CPython shims use.  Never destroyed by callers. */
int fqd_default_context(fqd_context **out);

/* Plain device buffers for callers that keep their inputs resident in HBM. */
int fqd_device_alloc(fqd_context *ctx, size_t bytes, void **dptr);
int fqd_device_free(fqd_context *ctx, void *dptr);
int fqd_device_upload(fqd_context *ctx, void *dptr, const void *host, size_t bytes);
int fqd_device_download(fqd_context *ctx, void *host, const void *dptr, size_t bytes);
int fqd_device_memset(fqd_context *ctx, void *dptr, int value, size_t bytes);
int fqd_context_synchronize(fqd_context *ctx);
/* Host staging memory that H2D copies can stream from (cudaHostAlloc). */
int fqd_host_alloc(size_t bytes, void **hptr);
int fqd_host_free(void *hptr);

/* ------------------------------------------------------------------------------------
 * The batched job.  Replaces the two inner loops of deduplicate_cluster
 * (src/fastqdedup/__init__.py:242-252: per record average_error_rate + Trie.add_sequence;
 * :272-276: pop_cluster + cluster dissection + set insertion) by one call.
 *
 * Record t (0 <= t < n_records) has
 *   key   = keys  + key_offsets[t]  .. key_offsets[t+1]     when key_offsets != NULL, else
 *           keys  + t*key_stride    .. + (key_lengths ? key_lengths[t] : key_length)
 *   qual  likewise (quals == NULL: no quality strings => nothing is filtered).
 * All pointers live in `memory_space` (host or device).
 * ------------------------------------------------------------------------------------ */
typedef struct {
    uint64_t n_records;
    const uint8_t *keys;
    const uint64_t *key_offsets;   /* n_records+1 entries or NULL */
    const uint32_t *key_lengths;   /* n_records entries or NULL (fixed stride only) */
    uint32_t key_stride;
    uint32_t key_length;
    const uint8_t *quals;
    const uint64_t *qual_offsets;
    const uint32_t *qual_lengths;
    uint32_t qual_stride;
    uint32_t qual_length;
    int32_t max_distance;          /* -d, __init__.py:326 */
    int32_t use_edit_distance;     /* --edit, __init__.py:341 */
    int32_t method;                /* FQD_METHOD_*, __init__.py:125-130 */
    int32_t memory_space;          /* FQD_MEM_* */
    double max_average_error_rate; /* -e; >= 1.0 disables the filter (__init__.py:235) */
    uint8_t phred_offset;          /* 33 */
    uint8_t reserved[7];
    const char *alphabet;          /* NULL => "ACGTN" (__init__.py:240); grows on demand */
    const uint32_t *record_counts; /* NULL => every record counts once; else record t stands
                                      for record_counts[t] identical reads (a pre-counted
                                      (count, sequence) list like pop_cluster's, :839-843) */
} fqd_cluster_job;

typedef struct {
    uint64_t total_records;        /* __init__.py:246 */
    uint64_t discarded_records;    /* __init__.py:249 */
    uint64_t number_of_sequences;  /* trie.number_of_sequences after pass 1, __init__.py:258 */
    uint64_t number_of_uniques;
    uint64_t number_of_clusters;   /* __init__.py:274 */
    uint64_t number_selected;      /* len(deduplicated_set), __init__.py:279 */
    uint64_t candidate_pairs;      /* pairs verified by the distance kernels (all passes) */
    uint64_t bad_record;           /* FQD_ERR_PHRED: index of the first offending record */
    uint32_t bad_char;             /* ... and the byte */
    uint32_t key_bits;             /* K: bit planes per symbol actually used */
    uint32_t key_words;            /* 32-bit words per packed key */
    uint32_t n_passes;             /* pigeonhole passes executed */
    float ms_total;                /* device time of the whole job (CUDA events) */
    float ms_ingest;               /* filter + pack + exact dedupe */
    float ms_gather;
    float ms_neighbour;            /* all pigeonhole passes (bucket build + compare) */
    float ms_select;               /* components + dissection + output */
    float ms_h2d;                  /* host -> device staging (memory_space == HOST) */
    float ms_compare;              /* the compare kernels alone (inside ms_neighbour) */
    float ms_ingest_kernel;        /* the ingest kernel launch alone (inside ms_ingest) */
    float ms_table_clear;          /* clearing the dedupe table (inside ms_ingest) */
    float ms_bucket_build;         /* signature count + scan + scatter (inside ms_neighbour) */
    uint32_t launches;             /* kernels launched by this job */
    uint32_t plan_flags;           /* FQD_PLAN_*: which launch plan the stages took */
    float ms_partition_kernel;     /* partitioned dedupe: filter + pack + partition pass */
    float ms_dedupe_kernel;        /* partitioned dedupe: the shared-memory tile kernel (dedupe + fused pass 0) */
    uint64_t own_uniques;          /* tile-sharded job: unique keys this rank owns (what fqd_cluster_fetch returns) */
    uint64_t h2d_bytes;            /* HOST jobs: input bytes copied host -> device (keys packed on the host cross at
                                      3 bits per symbol) */
} fqd_cluster_stats;

#define FQD_PLAN_DEDUPE_PARTITIONED 1u /* exact dedupe: partition by hash, tables in shared-memory tiles */
#define FQD_PLAN_PASSES_PARTITIONED 2u /* Hamming passes: partition by block hash, multimap in shared-memory tiles */
#define FQD_PLAN_PASS0_FUSED 4u        /* pass 0 ran inside the dedupe tiles (records partitioned by block 0) */
#define FQD_PLAN_PASS1_TILES_EMITTED 8u /* the dedupe tiles also filled the tiles of pass 1 */
#define FQD_PLAN_SHARD_TILES 16u       /* sharded job: tiles owned by ranks, fragments fetched from peer memory */

/* Runs the job.  On success the per-unique result stays in the context until the next
 * job.  keep_bitmap (optional, in the job's memory space, (n_records+31)/32 uint32 words,
 * bit t%32 of word t/32) marks the records pass 2 must emit: the first occurrence in
 * file order of every selected key (__init__.py:201-206). */
int fqd_cluster(fqd_context *ctx, const fqd_cluster_job *job, fqd_cluster_stats *stats,
                uint32_t *keep_bitmap);

/* Per-unique view of the last job (host buffers with number_of_uniques entries each, any
 * may be NULL): first = index of the first record carrying the key (over all records,
 * filtered or not); count = kept records with that key; label = `first` of the
 * smallest-first member of the unique's cluster; selected = 1 when dissection kept it.
 * Order is unspecified (sort by `first` for a canonical view).
 * After a tile-sharded job (plan_flags & FQD_PLAN_SHARD_TILES) every rank returns ITS OWN
 * stats.own_uniques keys and `label` is the cluster's root in the job-wide id space -- equal for all
 * members of a cluster on all ranks; the caller holding every rank's view maps it to the smallest
 * `first` (fastqdedup_b200/multigpu.py does). */
int fqd_cluster_fetch(fqd_context *ctx, uint64_t *first, uint32_t *count, uint64_t *label,
                      uint8_t *selected);

/* Ascending record indices of the selected keys' first occurrences
 * (number_selected entries, host buffer). */
int fqd_cluster_fetch_selected(fqd_context *ctx, uint64_t *indices);

/* ------------------------------------------------------------------------------------
 * The same job sharded over the GPUs of one box (BASELINE.json config 5).  Records are split
 * contiguously over `world` ranks (in rank order); rank r passes its own records and the global
 * index of its first record.  Every rank ends with the keep bitmap of ITS OWN records.
 *
 * Two plans.  Tile-sharded (Hamming, keys up to 6 packed words, <= 16 ranks):
 * the shared-memory tiles of the streaming plan get an owner rank each; every rank partitions its
 * records locally and the owners' tile kernels fetch the fragments of a tile straight from the
 * peers' HBM over NVLink (peer memory; CUDA IPC between processes), edges are read from all ranks'
 * lists by the union-find that consumes them; NCCL only carries counters.  Replicated-set (every
 * other job, and the fallback for skew the tiles cannot hold): all-to-all of the locally
 * deduplicated keys to an owner rank, all-gather of the merged unique set, all-gather of
 * spanning-forest pairs / flags.  fqd_cluster_fetch after the replicated-set plan: first, count,
 * label for every key on every rank, `selected` only for the keys whose first record is the rank's
 * own (the union over ranks is the complete set).
 * ------------------------------------------------------------------------------------ */
typedef struct fqd_comm fqd_comm;
/* rank 0 creates the id and hands the 128 bytes to the other ranks by any means */
int fqd_nccl_unique_id(uint8_t id[128]);
int fqd_comm_create(fqd_context *ctx, int rank, int world, const uint8_t id[128], fqd_comm **out);
void fqd_comm_destroy(fqd_comm *comm);
int fqd_cluster_sharded(fqd_context *ctx, fqd_comm *comm, const fqd_cluster_job *local_job,
                        uint64_t index_base, fqd_cluster_stats *stats, uint32_t *keep_bitmap);
/* All `world` ranks driven by one process (one context each; contexts may share a GPU):
 * the exchanges become device-to-device copies.  jobs/index_bases/stats/keep_bitmaps have
 * `world` entries. */
int fqd_cluster_sharded_local(fqd_context **ctxs, int world, const fqd_cluster_job *jobs,
                              const uint64_t *index_bases, fqd_cluster_stats *stats,
                              uint32_t **keep_bitmaps);

/* ------------------------------------------------------------------------------------
 * Function-level entry points mirroring the reference's C extensions one to one.
 * ------------------------------------------------------------------------------------ */

/* _fastq.average_error_rate (src/fastqdedup/_fastqmodule.c:38-76), batched: strings are
 * phred[offsets[i] .. offsets[i+1]); out[i] receives the mean error probability (NaN for
 * an empty string).  FQD_ERR_PHRED names the first bad string/byte via bad_index/bad_char. */
int fqd_average_error_rate(fqd_context *ctx, const uint8_t *phred, const uint64_t *offsets,
                           uint64_t n_strings, uint8_t phred_offset, double *out,
                           uint64_t *bad_index, uint32_t *bad_char);

/* _distance.within_distance (src/fastqdedup/_distancemodule.c:46-93 over distances.h),
 * batched over pairs (a_i, b_i); out[i] in {0,1}. */
int fqd_within_distance(fqd_context *ctx, const uint8_t *a, const uint64_t *a_offsets,
                        const uint8_t *b, const uint64_t *b_offsets, uint64_t n_pairs,
                        int32_t max_distance, int32_t use_edit_distance, uint8_t *out);

/* ------------------------------------------------------------------------------------
 * The two FASTQ passes around the job, native (host threads + zlib; no GPU involved).
 * ------------------------------------------------------------------------------------ */

/* A Python slice object (the --check-lengths of one input file, length_string_to_slices,
 * src/fastqdedup/__init__.py:364-375): has_x == 0 means None. */
typedef struct {
    int64_t start, stop, step;
    uint8_t has_start, has_stop, has_step;
    uint8_t reserved[5];
} fqd_slice;

typedef struct fqd_fastq_scan fqd_fastq_scan;

/* Pass 1 (fastq_files_to_records + joinfunc_from_check_slices, __init__.py:170-186, :160-167, and the key /
 * quality construction of deduplicate_cluster, :243-251): reads the n_files FASTQ files (plain or gzip) in lock
 * step, stops at the shortest, checks that the records of a tuple are mates (FQD_ERR_FASTQ "FASTQ files not in
 * sync: <names> are not mates.") and builds, per record tuple, key = concatenation over the files of
 * sequence[slice of that file] and, with want_quals, the same slices of the quality strings.  slices == NULL: whole
 * sequences.  threads <= 0: one per core (at most 16), besides a reader and a parser thread per file. */
int fqd_fastq_scan_open(const char *const *paths, int n_files, const fqd_slice *slices, int want_quals, int threads,
                        fqd_fastq_scan **out);
uint64_t fqd_fastq_scan_records(const fqd_fastq_scan *scan);
/* The rows, in the form fqd_cluster_job takes (memory_space HOST): *offsets != NULL: ragged rows
 * keys[offsets[t] .. offsets[t+1]); else every row has *stride bytes.  Owned by the scan. */
int fqd_fastq_scan_keys(const fqd_fastq_scan *scan, const uint8_t **keys, const uint64_t **offsets, uint32_t *stride);
int fqd_fastq_scan_quals(const fqd_fastq_scan *scan, const uint8_t **quals, const uint64_t **offsets, uint32_t *stride);
void fqd_fastq_scan_free(fqd_fastq_scan *scan);

/* Pass 2 (filter_fastq_files_on_set, __init__.py:189-206): re-reads the inputs and writes record tuple t of every
 * file to the matching output file iff bit t of keep_bitmap (fqd_cluster) is set, as "@name\nseq\n+\nqual\n"
 * (dnaio's fastq_bytes()).  Outputs ending in .gz are compressed at level 1 (__init__.py:197-198) by the worker
 * threads, one gzip member per block, written in order: the compression is off the serial path and the
 * decompressed bytes are the reference's. */
int fqd_fastq_emit(const char *const *in_paths, const char *const *out_paths, int n_files, const uint32_t *keep_bitmap,
                   uint64_t n_records, int threads, uint64_t *n_written);

/* Host-side key packing (joinfunc_from_check_slices' output, __init__.py:160-167, 3 bits per symbol instead of 8):
 * rows of key_length (<= 64) bytes over ACGTN, key_stride apart -> rows of ceil(3 * key_length / 32) uint32 words,
 * the three code-bit planes of the key back to back (plane p = bit p+1 of every ASCII byte, in bits
 * [p * key_length, (p+1) * key_length)).  Multithreaded (FQD_PACK_THREADS, default: all cores), AVX-512 when the CPU
 * has it.  (fqd_cluster packs HOST jobs with fixed-length keys itself, chunk by chunk ahead of the PCIe copy, in a
 * cheaper internal form: the rows of a chunk as three plane streams.)  FQD_ERR_UNSUPPORTED + *bad_record when a byte
 * is outside ACGTN. */
int fqd_pack_keys(const uint8_t *keys, uint64_t n_records, uint32_t key_length, uint32_t key_stride, uint32_t *packed,
                  uint64_t *bad_record);

/* Measurement aid (not on the product path): the integer-issue peak of the
 * context's GPU in thread-level operations per second, from dependent-free instruction streams --
 * LOP3 alone, POPC alone, and the 4 LOP3 : 1 POPC mix of the XOR+POPC Hamming compare that replaces
 * within_hamming_distance (src/fastqdedup/distances.h:8-31).  bench.py reports the compare phase
 * against it. */
int fqd_int_peak(fqd_context *ctx, double *lop3_ops_per_s, double *popc_ops_per_s,
                 double *mixed_ops_per_s);

/* _trie.Trie (src/fastqdedup/_triemodule.c:596-983).  Sequences are staged on the host;
 * the neighbour search / clustering runs on the GPU when contains_sequence or
 * pop_cluster is called. */
int fqd_trie_new(fqd_context *ctx, const uint8_t *alphabet, size_t alphabet_len,
                 fqd_trie **out);                                   /* :613-642 */
void fqd_trie_free(fqd_trie *trie);                                 /* :606-611 */
int fqd_trie_add_sequence(fqd_trie *trie, const uint8_t *seq, size_t len); /* :677-706 */
int fqd_trie_contains_sequence(fqd_trie *trie, const uint8_t *seq, size_t len,
                               int32_t max_distance, int32_t use_edit_distance,
                               int32_t *found);                     /* :730-758 */
/* pop_cluster (:778-897): removes one cluster; *n_items members can then be read with
 * fqd_trie_cluster_item until the next call on this trie. */
int fqd_trie_pop_cluster(fqd_trie *trie, int32_t max_distance, int32_t use_edit_distance,
                         uint64_t *n_items);
int fqd_trie_cluster_item(fqd_trie *trie, uint64_t i, uint32_t *count,
                          const uint8_t **seq, size_t *len);
uint64_t fqd_trie_number_of_sequences(const fqd_trie *trie);        /* :651-653 */
/* alphabet property (:645-648): writes up to cap bytes, returns the alphabet size */
size_t fqd_trie_alphabet(const fqd_trie *trie, uint8_t *buf, size_t cap);
uint64_t fqd_trie_memory_size(const fqd_trie *trie);                /* :909-913 */
/* raw_stats (:929-964): (max_len+1) rows of (alphabet+1) counters, row-major into buf;
 * returns the number of rows, *row_len the row length.  buf may be NULL to size. */
size_t fqd_trie_raw_stats(const fqd_trie *trie, uint64_t *buf, size_t cap, size_t *row_len);

#ifdef __cplusplus
}
#endif
#endif /* FQD_B200_H */
