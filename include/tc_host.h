/*
 * tc_host.h — host-side native helpers of trueconsense_b200 (libtchost.so).
 *
 * Not part of the GPU drop-in boundary (that is trueconsense_b200.h); this is the
 * step immediately before it: turning a coordinate-sorted BAM into the flat read arrays of
 * tc_reads_t (the job pysam/htslib's reader does for TrueConsense/indexing.py:96),
 * writing such arrays back out as a BAM (test fixtures), and generating the synthetic read
 * sets of BASELINE.json's five configs.  Plain C, zlib + OpenMP only.
 */
#ifndef TC_HOST_H
#define TC_HOST_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* malloc-owned flat read arrays; same layout rules as tc_reads_t (trueconsense_b200.h) */
typedef struct tc_hostreads {
    int64_t   n_reads;
    int64_t   n_seq_words;
    int64_t   n_cigar_ops;
    int32_t*  pos;
    uint16_t* flag;
    uint8_t*  mapq;
    int32_t*  l_seq;
    uint32_t* seq_off;      /* [n+1] */
    uint32_t* cigar_off;    /* [n+1] */
    uint32_t* seq4;
    uint8_t*  qual;         /* [8*n_seq_words] */
    uint32_t* cigar;
    uint64_t* qname_hash;   /* low 32 bits: khash X31 string hash of QNAME (htslib's), high 32: FNV-1a */
    int32_t*  mpos;
    int32_t*  isize;
    int32_t*  tid;          /* [n] reference index of every kept read */
    int32_t*  mtid;         /* [n] */
    /* header */
    int32_t   n_ref;
    int32_t*  ref_len;      /* [n_ref] */
    char*     ref_names;    /* n_ref NUL-terminated names back to back */
    int64_t   ref_names_len;
    /* bookkeeping */
    int64_t   n_records;    /* records in the file */
    int64_t   n_dropped_unplaced; /* records with refID < 0 (never enter a pileup) */
    int64_t   aligned_bases;/* sum over kept reads of M,=,X,D,N lengths */
    int32_t   sorted;       /* 1 if (tid,pos) is non-decreasing in file order */
    int32_t   max_ref_span;
    double    t_inflate_s, t_parse_s;
} tc_hostreads_t;

/* Read a BAM file (BGZF) into flat arrays.  n_threads <= 0: all cores.  Returns 0 or a
 * negative code and a message in err. */
int  tc_bam_read(const char* path, int n_threads, tc_hostreads_t* out, char* err, int errlen);
void tc_hostreads_free(tc_hostreads_t* r);

/* The file's uncompressed payload and where its placed records start: the host half of the decode when the records are
 * parsed on the GPU (trueconsense_b200.h: tc_bam_records_to_reads). */
typedef struct tc_bampayload {
    uint8_t* payload;       /* the BGZF blocks' concatenated payload (malloc-owned) */
    int64_t  n_bytes;
    int64_t* rec_off;       /* [n_reads] offset of every placed record's refID field (its block_size sits 4 bytes in front) */
    int64_t  n_reads;       /* placed records (refID >= 0), file order */
    int64_t  n_records, n_dropped_unplaced;
    int32_t  n_ref;
    int32_t* ref_len;
    char*    ref_names;
    int64_t  ref_names_len;
    double   t_inflate_s, t_index_s;
} tc_bampayload_t;
int  tc_bam_payload(const char* path, int n_threads, tc_bampayload_t* out, char* err, int errlen);
void tc_bampayload_free(tc_bampayload_t* p);

/* One BGZF member: its raw DEFLATE stream is file[coff .. coff + csize) — the member's CRC-32 and ISIZE follow it — and
 * inflates to payload[uoff .. uoff + usize).  (Same layout as trueconsense_b200.h's tc_bgzf_block_t.) */
#ifndef TC_BGZF_BLOCK_T
#define TC_BGZF_BLOCK_T
typedef struct tc_bgzf_block { int64_t coff; int32_t csize; int32_t usize; int64_t uoff; } tc_bgzf_block_t;
#endif
/* The BAM file mapped read-only and the index of its members (one header read per member: the only sequential step left on
 * the host when the GPU inflates — tc_bgzf_inflate).  Replaces htslib's bgzf reader under the reference's
 * indexing.py:6-19 (`Readbam`). */
typedef struct tc_bgzf_map {
    const uint8_t*   file;
    int64_t          file_bytes;
    tc_bgzf_block_t* blocks;
    int64_t          n_blocks;
    int64_t          payload_bytes;     /* sum of the members' ISIZE */
} tc_bgzf_map_t;
int  tc_bgzf_map(const char* path, tc_bgzf_map_t* out, char* err, int errlen);
void tc_bgzf_unmap(tc_bgzf_map_t* m);

/* Write flat arrays as a coordinate-sorted single-contig BAM (names are "q<hash hex>"). */
int  tc_bam_write(const char* path, const tc_hostreads_t* reads, const char* ref_name,
                  int32_t ref_len, int level, char* err, int errlen);

/* ---- compact transport forms (trueconsense_b200.h: tc_reads_t.seq2) ----
 * seq2[w] = the 8 bases of seq4[w] at two bits each (A C G T = 0 1 2 3, base j in bits 2j+1:2j); *exc_idx / *exc_val (malloc'd,
 * release with tc_host_free; NULL when there are none) list, ascending, the words that hold anything else over their valid
 * bases.  Words of seq2 that are on the list are unspecified.  0 on success. */
int  tc_seq2_pack(const uint32_t* seq4, int64_t n_seq_words, const uint32_t* seq_off, const int32_t* l_seq, int64_t n_reads,
                  uint16_t* seq2, uint32_t** exc_idx, uint32_t** exc_val, int64_t* n_exc, int n_threads);
void tc_host_free(void* p);

/* ---- synthetic reads ---- */
enum { TC_VAR_SUB = 0, TC_VAR_INS = 1, TC_VAR_DEL = 2 };

typedef struct tc_synth_variant {
    int32_t pos;        /* 0-based; INS: anchor column (bases inserted after it); DEL: first deleted column */
    int32_t kind;       /* TC_VAR_* */
    int32_t len;        /* INS / DEL length; SUB: 1 */
    int32_t alt;        /* SUB: 4-bit code of the alternative base; INS: seed of the inserted bases */
    double  frac;       /* fraction of covering fragments that carry it */
} tc_synth_variant_t;

typedef struct tc_synth_params {
    uint64_t seed;
    int64_t  n_reads;           /* total reads (paired: rounded down to an even number) */
    int32_t  ref_len;
    int32_t  read_len;          /* nominal reference span of one read */
    int32_t  read_len_jitter;   /* uniform +- */
    int32_t  paired;            /* 1: FR pairs with overlapping mates */
    int32_t  insert_mean;
    int32_t  insert_sd;
    int32_t  n_amplicons;       /* 0: shotgun starts; >0: tiled amplicons with identical starts +- jitter */
    int32_t  amplicon_jitter;
    int32_t  indel_maxlen;      /* random indel length 1..maxlen */
    int32_t  softclip_max;
    int32_t  qual_min, qual_max;
    int32_t  n_variants;
    double   sub_rate;          /* per-base substitution rate */
    double   indel_rate;        /* per-base rate of random indel events (half insertions) */
    double   softclip_rate;     /* per read end */
    double   n_rate;            /* per-base rate of 'N' */
    double   iupac_rate;        /* per-base rate of '=' / IUPAC codes */
    double   refskip_rate;      /* per-read probability of one N (ref-skip) op */
    double   special_flag_rate; /* per-read probability of 0x100/0x200/0x400/0x800/0x4/improper */
    double   lowmapq_rate;
    const tc_synth_variant_t* variants; /* sorted by pos */
} tc_synth_params_t;

/* ref_codes: ref_len 4-bit base codes (1,2,4,8), one per byte. */
int  tc_synth_reads(const tc_synth_params_t* p, const uint8_t* ref_codes, int n_threads,
                    tc_hostreads_t* out, char* err, int errlen);
/* reads [r0, r1) of that start-sorted set only (r1 < 0: to the end) */
int  tc_synth_reads_range(const tc_synth_params_t* p, const uint8_t* ref_codes, int n_threads, int64_t r0, int64_t r1,
                          tc_hostreads_t* out, char* err, int errlen);

#ifdef __cplusplus
}
#endif
#endif /* TC_HOST_H */
