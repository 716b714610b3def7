/*
 * synth.c — seeded synthetic read sets shaped like BASELINE.json's configs, written straight
 * into the flat arrays of tc_reads_t (and from there, optionally, to a BAM by tc_bam_write).
 *
 * Three stages: (1) place fragments (shotgun / tiled amplicons / FR pairs) and sort reads by
 * start with a counting sort, (2) size every read (its CIGAR op count and query length) by
 * running the per-read generator in count mode, (3) prefix-sum the sizes and run the generator
 * again in fill mode.  Stages 2 and 3 are OpenMP-parallel; every read draws from its own
 * splitmix64 stream keyed by (seed, fragment, mate), so the output does not depend on the
 * thread count.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdarg.h>
#include <stdint.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "tc_host.h"

int tc_hostreads_alloc_(tc_hostreads_t* o, int64_t n, int64_t n_words, int64_t n_ops);

static int sfail(char* err, int errlen, int code, const char* fmt, ...) {
    if (err && errlen > 0) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(err, errlen, fmt, ap);
        va_end(ap);
    }
    return code;
}

typedef struct { uint64_t s; } rng_t;
static inline uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
static inline uint64_t rng_next(rng_t* r) { r->s += 0x9e3779b97f4a7c15ULL; return mix64(r->s); }
static inline double rng_u(rng_t* r) { return (double)(rng_next(r) >> 11) * (1.0 / 9007199254740992.0); }
static inline uint32_t rng_below(rng_t* r, uint32_t n) { return (uint32_t)(((rng_next(r) >> 32) * (uint64_t)n) >> 32); }
static inline double rng_normal(rng_t* r) {
    double u1 = rng_u(r), u2 = rng_u(r);
    if (u1 < 1e-300) u1 = 1e-300;
    return sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
}

typedef struct {
    int32_t start, span;    /* reference start and nominal span */
    int32_t mstart, isize;  /* mate start / signed template length (paired) */
    uint32_t frag;          /* fragment id (shared by mates) */
    uint8_t mate;           /* 0 unpaired, 1 first, 2 second */
    uint8_t rev;
} place_t;

static const uint8_t BASE_CODES[4] = {1, 2, 4, 8};
static const uint8_t ODD_CODES[11] = {0, 3, 5, 6, 7, 9, 10, 11, 12, 13, 14};

typedef struct {
    uint32_t* cigar; int n_ops;
    uint8_t* seq; uint8_t* qual; int l_seq;   /* seq: one 4-bit code per byte (scratch) */
    int fill;
} rdout_t;

/* count mode needs the same merging decisions as fill mode: track the last op separately */
typedef struct { rdout_t o; uint32_t last_op; int have_last; } gen_t;

static inline void g_op(gen_t* g, uint32_t op, uint32_t len) {
    if (len == 0) return;
    if (g->have_last && g->last_op == op) {
        if (g->o.fill) g->o.cigar[g->o.n_ops - 1] += len << 4;
        return;
    }
    if (g->o.fill) g->o.cigar[g->o.n_ops] = (len << 4) | op;
    g->o.n_ops++;
    g->last_op = op; g->have_last = 1;
}
static inline void g_base(gen_t* g, uint8_t code, uint8_t q) {
    if (g->o.fill) { g->o.seq[g->o.l_seq] = code; g->o.qual[g->o.l_seq] = q; }
    g->o.l_seq++;
}

static void gen_read(const tc_synth_params_t* P, const uint8_t* ref, const place_t* pl, gen_t* g,
                     uint16_t* flag_out, uint8_t* mapq_out) {
    rng_t r; r.s = mix64(P->seed ^ mix64(((uint64_t)pl->frag << 2) | pl->mate));
    int qspan = P->qual_max - P->qual_min + 1;
    if (qspan < 1) qspan = 1;
#define QUAL() ((uint8_t)(P->qual_min + (int)rng_below(&r, (uint32_t)qspan)))
    g->o.n_ops = 0; g->o.l_seq = 0; g->have_last = 0;
    int start = pl->start, end = pl->start + pl->span;
    if (end > P->ref_len) end = P->ref_len;
    /* leading soft clip */
    if (P->softclip_max > 0 && rng_u(&r) < P->softclip_rate) {
        int n = 1 + (int)rng_below(&r, (uint32_t)P->softclip_max);
        g_op(g, 4, (uint32_t)n);
        for (int i = 0; i < n; ++i) g_base(g, BASE_CODES[rng_below(&r, 4)], QUAL());
    }
    /* optional ref-skip somewhere in the middle */
    int skip_at = -1, skip_len = 0;
    if (P->refskip_rate > 0 && rng_u(&r) < P->refskip_rate && end - start > 40) {
        skip_len = 1 + (int)rng_below(&r, 20);
        skip_at = start + 10 + (int)rng_below(&r, (uint32_t)(end - start - 20 - skip_len > 1 ? end - start - 20 - skip_len : 1));
    }
    /* first variant at or after start */
    int vi = 0;
    {
        int lo = 0, hi = P->n_variants;
        while (lo < hi) { int mid = (lo + hi) / 2; if (P->variants[mid].pos < start) lo = mid + 1; else hi = mid; }
        vi = lo;
    }
    int just_indel = 1;   /* no indel on the first aligned column */
    int x = start;
    while (x < end) {
        int last_col = (x == end - 1);
        if (x == skip_at && !just_indel && x + skip_len < end - 1) {
            g_op(g, 3, (uint32_t)skip_len); x += skip_len; just_indel = 1; continue;
        }
        /* true variants anchored at x */
        uint8_t sub_code = 0; int ins_len = 0; uint32_t ins_seed = 0; int del_len = 0;
        while (vi < P->n_variants && P->variants[vi].pos < x) ++vi;
        for (int v = vi; v < P->n_variants && P->variants[v].pos == x; ++v) {
            const tc_synth_variant_t* V = &P->variants[v];
            uint64_t h = mix64(P->seed ^ mix64(0x5151ULL + ((uint64_t)pl->frag << 20) + (uint64_t)v));
            double u = (double)(h >> 11) * (1.0 / 9007199254740992.0);
            if (u >= V->frac) continue;
            if (V->kind == TC_VAR_SUB) sub_code = (uint8_t)V->alt;
            else if (V->kind == TC_VAR_INS) { ins_len = V->len; ins_seed = (uint32_t)V->alt; }
            else if (V->kind == TC_VAR_DEL) del_len = V->len;
        }
        if (del_len > 0 && !just_indel && x + del_len < end - 1 && x > start) {
            g_op(g, 2, (uint32_t)del_len); x += del_len; just_indel = 1; continue;
        }
        /* random indel error */
        if (P->indel_rate > 0 && !just_indel && !last_col && ins_len == 0 && rng_u(&r) < P->indel_rate) {
            int n = 1 + (int)rng_below(&r, (uint32_t)(P->indel_maxlen > 0 ? P->indel_maxlen : 1));
            if (rng_next(&r) & 1) {
                if (x + n < end - 1) { g_op(g, 2, (uint32_t)n); x += n; just_indel = 1; continue; }
            } else {
                g_op(g, 1, (uint32_t)n);
                for (int i = 0; i < n; ++i) g_base(g, BASE_CODES[rng_below(&r, 4)], QUAL());
                just_indel = 1;
                /* fall through: the column x itself is still emitted as a match below */
            }
        }
        /* the aligned base of column x */
        uint8_t code = sub_code ? sub_code : ref[x];
        double u = rng_u(&r);
        if (u < P->sub_rate) {
            int ci = code == 1 ? 0 : code == 2 ? 1 : code == 4 ? 2 : code == 8 ? 3 : -1;
            uint8_t alt = ci < 0 ? BASE_CODES[rng_below(&r, 4)] : BASE_CODES[(ci + 1 + (int)rng_below(&r, 3)) & 3];
            code = alt;
        } else if (u < P->sub_rate + P->n_rate) code = 15;
        else if (u < P->sub_rate + P->n_rate + P->iupac_rate) code = ODD_CODES[rng_below(&r, 11)];
        g_op(g, 0, 1);
        g_base(g, code, QUAL());
        just_indel = 0;
        ++x;
        if (ins_len > 0 && x < end) {   /* true insertion after anchor column (never after the last column) */
            g_op(g, 1, (uint32_t)ins_len);
            rng_t ir; ir.s = mix64(0xabcdef12ULL + ins_seed);
            for (int i = 0; i < ins_len; ++i) {
                uint8_t b = BASE_CODES[rng_below(&ir, 4)];
                if (rng_u(&r) < P->sub_rate) b = BASE_CODES[rng_below(&r, 4)];
                g_base(g, b, QUAL());
            }
            just_indel = 1;
        }
    }
    /* trailing soft clip */
    if (P->softclip_max > 0 && rng_u(&r) < P->softclip_rate) {
        int n = 1 + (int)rng_below(&r, (uint32_t)P->softclip_max);
        g_op(g, 4, (uint32_t)n);
        for (int i = 0; i < n; ++i) g_base(g, BASE_CODES[rng_below(&r, 4)], QUAL());
    }
    /* flags and mapq */
    uint16_t fl = 0;
    if (pl->mate == 0) fl = pl->rev ? 16 : 0;
    else {
        fl = 1 | 2 | (pl->mate == 1 ? 64 : 128);
        if (pl->rev) fl |= 16; else fl |= 32;
    }
    if (P->special_flag_rate > 0) {
        /* keyed by fragment for the pair-level property (improper), by read for the others */
        double us = rng_u(&r);
        if (us < P->special_flag_rate) {
            switch (rng_below(&r, 6)) {
                case 0: fl |= 0x100; break;
                case 1: fl |= 0x200; break;
                case 2: fl |= 0x400; break;
                case 3: fl |= 0x800; break;
                case 4: fl |= 0x4; break;
                default: if (fl & 1) fl &= (uint16_t)~2u; else fl |= 0x200; break;
            }
        }
    }
    *flag_out = fl;
    *mapq_out = (P->lowmapq_rate > 0 && rng_u(&r) < P->lowmapq_rate) ? (uint8_t)rng_below(&r, 21) : 60;
#undef QUAL
}

int tc_synth_reads(const tc_synth_params_t* P, const uint8_t* ref, int n_threads, tc_hostreads_t* out,
                   char* err, int errlen) {
    return tc_synth_reads_range(P, ref, n_threads, 0, -1, out, err, errlen);
}

/* Reads [r0, r1) of the start-sorted set tc_synth_reads would produce (r1 < 0: to the end): every read draws from its own
 * stream, so a shard is generated without the others — what a rank of a read-range sharded run needs. */
int tc_synth_reads_range(const tc_synth_params_t* P, const uint8_t* ref, int n_threads, int64_t r0, int64_t r1,
                         tc_hostreads_t* out, char* err, int errlen) {
    memset(out, 0, sizeof(*out));
#ifdef _OPENMP
    if (n_threads <= 0) n_threads = omp_get_num_procs();
#else
    n_threads = 1;
#endif
    int64_t n = P->n_reads;
    if (P->paired) n &= ~1LL;
    if (n < 0 || P->ref_len <= 0 || P->read_len <= 0) return sfail(err, errlen, -2, "bad synth parameters");
    if (P->read_len > P->ref_len) return sfail(err, errlen, -2, "read_len exceeds ref_len");
    place_t* pl = malloc(sizeof(place_t) * (size_t)(n > 0 ? n : 1));
    place_t* sorted = malloc(sizeof(place_t) * (size_t)(n > 0 ? n : 1));
    if (!pl || !sorted) { free(pl); free(sorted); return sfail(err, errlen, -5, "out of memory"); }
    int L = P->ref_len;
    /* (1) placement */
    rng_t pr; pr.s = mix64(P->seed ^ 0x706c616365ULL);
    if (P->paired) {
        for (int64_t f = 0; f < n / 2; ++f) {
            int isz = (int)lround(P->insert_mean + P->insert_sd * rng_normal(&pr));
            if (isz < P->read_len) isz = P->read_len;
            if (isz > L) isz = L;
            int fs = (int)rng_below(&pr, (uint32_t)(L - isz + 1));
            int swap = (int)(rng_next(&pr) & 1);   /* which mate is the forward one */
            place_t a, b;
            a.start = fs; a.span = P->read_len; a.mstart = fs + isz - P->read_len; a.isize = isz;
            a.frag = (uint32_t)f; a.mate = swap ? 2 : 1; a.rev = 0;
            b.start = fs + isz - P->read_len; b.span = P->read_len; b.mstart = fs; b.isize = -isz;
            b.frag = (uint32_t)f; b.mate = swap ? 1 : 2; b.rev = 1;
            pl[2 * f] = a; pl[2 * f + 1] = b;
        }
    } else {
        for (int64_t i = 0; i < n; ++i) {
            int span = P->read_len;
            if (P->read_len_jitter > 0) span += (int)rng_below(&pr, (uint32_t)(2 * P->read_len_jitter + 1)) - P->read_len_jitter;
            if (span < 1) span = 1;
            if (span > L) span = L;
            int st;
            if (P->n_amplicons > 0) {
                int a = (int)rng_below(&pr, (uint32_t)P->n_amplicons);
                int64_t room = L - P->read_len - P->read_len_jitter;
                if (room < 0) room = 0;
                st = P->n_amplicons > 1 ? (int)((room * a) / (P->n_amplicons - 1)) : 0;
                if (P->amplicon_jitter > 0) st += (int)rng_below(&pr, (uint32_t)(2 * P->amplicon_jitter + 1)) - P->amplicon_jitter;
                if (st < 0) st = 0;
                if (st + span > L) st = L - span;
            } else {
                st = (int)rng_below(&pr, (uint32_t)(L - span + 1));
            }
            pl[i].start = st; pl[i].span = span; pl[i].mstart = -1; pl[i].isize = 0;
            pl[i].frag = (uint32_t)i; pl[i].mate = 0; pl[i].rev = (uint8_t)(rng_next(&pr) & 1);
        }
    }
    /* counting sort by start (stable) */
    int64_t* bucket = calloc((size_t)L + 1, sizeof(int64_t));
    if (!bucket) { free(pl); free(sorted); return sfail(err, errlen, -5, "out of memory"); }
    for (int64_t i = 0; i < n; ++i) bucket[pl[i].start + 1]++;
    for (int i = 0; i < L; ++i) bucket[i + 1] += bucket[i];
    for (int64_t i = 0; i < n; ++i) sorted[bucket[pl[i].start]++] = pl[i];
    free(bucket); free(pl);
    /* the shard: from here on `sorted` / `n` are its reads only */
    place_t* sorted_all = sorted;
    if (r1 < 0 || r1 > n) r1 = n;
    if (r0 < 0) r0 = 0;
    if (r0 > r1) r0 = r1;
    sorted = sorted_all + r0;
    n = r1 - r0;

    /* (2) sizes */
    uint32_t* nops = malloc(4 * (size_t)(n + 1));
    uint32_t* lseq = malloc(4 * (size_t)(n + 1));
    if (!nops || !lseq) { free(sorted_all); free(nops); free(lseq); return sfail(err, errlen, -5, "out of memory"); }
#pragma omp parallel for schedule(static) num_threads(n_threads)
    for (int64_t i = 0; i < n; ++i) {
        gen_t g; memset(&g, 0, sizeof(g));
        uint16_t fl; uint8_t mq;
        gen_read(P, ref, &sorted[i], &g, &fl, &mq);
        nops[i] = (uint32_t)g.o.n_ops; lseq[i] = (uint32_t)g.o.l_seq;
    }
    uint64_t sw = 0, co = 0;
    for (int64_t i = 0; i < n; ++i) { sw += (lseq[i] + 7) / 8; co += nops[i]; }
    if (sw > 0xffffffffULL || co > 0xffffffffULL) {
        free(sorted_all); free(nops); free(lseq);
        return sfail(err, errlen, -7, "synthetic batch too large for 32-bit offsets; generate per shard");
    }
    if (tc_hostreads_alloc_(out, n, (int64_t)sw, (int64_t)co) != 0) {
        free(sorted_all); free(nops); free(lseq); tc_hostreads_free(out);
        return sfail(err, errlen, -5, "out of memory");
    }
    sw = 0; co = 0;
    for (int64_t i = 0; i < n; ++i) {
        out->seq_off[i] = (uint32_t)sw; out->cigar_off[i] = (uint32_t)co;
        sw += (lseq[i] + 7) / 8; co += nops[i];
    }
    out->seq_off[n] = (uint32_t)sw; out->cigar_off[n] = (uint32_t)co;

    /* (3) fill */
    int64_t aligned = 0; int maxspan = 0;
#pragma omp parallel num_threads(n_threads) reduction(+:aligned) reduction(max:maxspan)
    {
        int cap = P->read_len + P->read_len_jitter + 2 * P->softclip_max + 64;
        for (int v = 0; v < P->n_variants; ++v) if (P->variants[v].kind == TC_VAR_INS) cap += P->variants[v].len;
        cap = cap * 2 + 64 + 2 * (P->indel_maxlen > 0 ? P->indel_maxlen : 1) * (P->read_len + P->read_len_jitter);
        uint8_t* sq = malloc((size_t)cap);
        uint8_t* ql = malloc((size_t)cap);
#pragma omp for schedule(static)
        for (int64_t i = 0; i < n; ++i) {
            gen_t g; memset(&g, 0, sizeof(g));
            g.o.fill = 1; g.o.cigar = out->cigar + out->cigar_off[i]; g.o.seq = sq; g.o.qual = ql;
            uint16_t fl; uint8_t mq;
            gen_read(P, ref, &sorted[i], &g, &fl, &mq);
            out->pos[i] = sorted[i].start;
            out->flag[i] = fl; out->mapq[i] = mq; out->l_seq[i] = g.o.l_seq;
            out->mpos[i] = sorted[i].mstart; out->isize[i] = sorted[i].isize;
            out->tid[i] = 0; out->mtid[i] = sorted[i].mate ? 0 : -1;
            uint64_t qh = mix64(P->seed ^ mix64(0x716e616dULL + sorted[i].frag));
            out->qname_hash[i] = qh;
            uint8_t* dst = (uint8_t*)(out->seq4 + out->seq_off[i]);
            for (int q = 0; q < g.o.l_seq; ++q) {
                if (q & 1) dst[q >> 1] |= sq[q]; else dst[q >> 1] = (uint8_t)(sq[q] << 4);
            }
            memcpy(out->qual + 8ULL * out->seq_off[i], ql, (size_t)g.o.l_seq);
            int span = 0;
            for (int k = 0; k < g.o.n_ops; ++k) {
                uint32_t c = g.o.cigar[k]; uint32_t op = c & 15;
                if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) span += (int)(c >> 4);
            }
            aligned += span;
            if (span > maxspan) maxspan = span;
        }
        free(sq); free(ql);
    }
    free(sorted_all); free(nops); free(lseq);
    out->aligned_bases = aligned; out->max_ref_span = maxspan; out->sorted = 1;
    out->n_records = n; out->n_ref = 1;
    out->ref_len = malloc(sizeof(int32_t)); out->ref_len[0] = L;
    out->ref_names = malloc(4); memcpy(out->ref_names, "ref", 4); out->ref_names_len = 4;
    return 0;
}
