/*
 * pileup_oracle.c — CPU restatement of the pileup the reference obtains from pysam/htslib.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under trueconsense_b200/ may import, link or call this;
 * it is the checker for tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs.
 *
 * PARITY UNPINNED at the pysam/htslib boundary: the reference (TrueConsense 0.5.2) has no tests
 * and its pileup lives in un-vendored third-party code (pysam==0.23.3 wrapping htslib 1.21,
 * pyproject.toml:27) that is not installed in this image.  This file restates, from the
 * published behaviour of those libraries, exactly the part the reference's two call sites use:
 *   TrueConsense/indexing.py:100   pileup(stepper="nofilter", max_depth=10000000, min_base_quality=0)
 *   TrueConsense/indexing.py:139   PileupColumn.get_query_sequences(add_indels=True), .pos
 *   TrueConsense/Events.py:63-67   pileup(rname, pos-1, pos, truncate=True)   [all pysam defaults]
 * and the reference's own classifier, indexing.py:102-132 (parse_query_sequences).
 * Upstream functions restated: htslib sam.c bam_plp_push / bam_plp64_next /
 * resolve_cigar2 / bam_plp_set_maxcnt; pysam libcalignmentfile.pyx __advance_nofilter /
 * __advance_samtools; libcalignedsegment.pyx PileupColumn.get_query_sequences,
 * pileup_base_qual_skip, strand_mark_char.
 * Also restated: the mate-overlap quality rewriting that pysam's default ignore_overlaps=True enables
 * (htslib sam.c overlap_push / overlap_remove / tweak_overlap_quality with cigar_iref2iseq_set / _next,
 * as of htslib >= 1.13: which mate keeps its quality is drawn from a hash of the QNAME, and bases opposite
 * a deletion in the mate are down-weighted too; oparams_t.reserved bits 8-9 select 1 = off
 * (ignore_overlaps=False) or 2 = the rule of htslib <= 1.12).  It cannot change BuildIndex
 * (min_base_quality=0) and is skipped there; it decides which entries of ExtractInserts' columns pass the
 * min_base_quality=13 test wherever both mates of a proper pair cover the column.
 *
 * The engine below is deliberately the streaming, column-by-column one (a list of live reads,
 * one column emitted at a time, per-read CIGAR cursor) — slow and plain, nothing like the GPU
 * kernels it checks.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ctype.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    int64_t n_reads, n_seq_words, n_cigar_ops;
    const int32_t* pos; const uint16_t* flag; const uint8_t* mapq; const int32_t* l_seq;
    const uint32_t* seq_off; const uint32_t* cigar_off; const uint32_t* seq4; const uint8_t* qual;
    const uint32_t* cigar; const uint64_t* qname_hash; const int32_t* mpos; const int32_t* isize;
} oreads_t;   /* same layout as tc_reads_t */

typedef struct {
    uint32_t flag_filter; int32_t min_mapq; int32_t min_base_quality; int32_t ignore_orphans;
    int64_t max_depth; int32_t kernel; int32_t reserved;
} oparams_t;  /* same layout as tc_pileup_params_t */

enum { OP_M = 0, OP_I = 1, OP_D = 2, OP_N = 3, OP_S = 4, OP_H = 5, OP_P = 6, OP_EQ = 7, OP_X = 8 };
static const char NT16[] = "=ACMGRSVTWYHKDBN";

static inline int consumes_ref(int op) { return op == OP_M || op == OP_D || op == OP_N || op == OP_EQ || op == OP_X; }
static inline int is_match(int op) { return op == OP_M || op == OP_EQ || op == OP_X; }

static inline int seq_code(const oreads_t* R, int64_t i, int q) {
    const uint8_t* s = (const uint8_t*)(R->seq4 + R->seq_off[i]);
    return (q & 1) ? (s[q >> 1] & 15) : (s[q >> 1] >> 4);
}

/* one live read (htslib: lbnode_t + cstate_t) */
typedef struct {
    int64_t read;
    int32_t beg, end;       /* [beg, end) on the reference; end = beg + raw rlen */
    int32_t k, x, y;        /* CIGAR cursor: op index (-1 = untouched), ref coord of op start, query coord */
    uint8_t* q;             /* base qualities rewritten by tweak_overlap_quality (NULL: the read's own) */
} node_t;

/* ---------------------------------------------------------------- mate overlaps (htslib sam.c) */
enum { OLAP_HASH = 0, OLAP_OFF = 1, OLAP_FIRST = 2 };      /* oparams_t.reserved bits 8-9 */

/* khash.h __ac_Wang_hash */
static inline uint32_t wang_hash(uint32_t key) {
    key += ~(key << 15); key ^= (key >> 10); key += (key << 3); key ^= (key >> 6); key += ~(key << 11); key ^= (key >> 16);
    return key;
}

/* cigar_iref2iseq_set / cigar_iref2iseq_next: walk to the first / next M,=,X base */
typedef struct { const uint32_t* cig; const uint32_t* end; int64_t icig, iseq, iref; } cwalk_t;

static int iref2iseq_set(cwalk_t* w, int64_t pos) {
    if (pos < 0) return -1;
    w->icig = 0; w->iseq = 0; w->iref = 0;
    while (w->cig < w->end) {
        int cig = (int)(*w->cig & 15); int64_t ncig = (int64_t)(*w->cig >> 4);
        if (cig == OP_S) { w->cig++; w->iseq += ncig; w->icig = 0; continue; }
        if (cig == OP_H || cig == OP_P) { w->cig++; w->icig = 0; continue; }
        if (cig == OP_M || cig == OP_EQ || cig == OP_X) {
            pos -= ncig;
            if (pos < 0) { w->icig = ncig + pos; w->iseq += w->icig; w->iref += w->icig; return OP_M; }
            w->cig++; w->iseq += ncig; w->icig = 0; w->iref += ncig;
            continue;
        }
        if (cig == OP_I) { w->cig++; w->iseq += ncig; w->icig = 0; continue; }
        if (cig == OP_D || cig == OP_N) {
            pos -= ncig;
            if (pos < 0) pos = 0;
            w->cig++; w->icig = 0; w->iref += ncig;
            continue;
        }
        return -2;
    }
    w->iseq = -1;
    return -1;
}

static int iref2iseq_next(cwalk_t* w) {
    while (w->cig < w->end) {
        int cig = (int)(*w->cig & 15); int64_t ncig = (int64_t)(*w->cig >> 4);
        if (cig == OP_M || cig == OP_EQ || cig == OP_X) {
            if (w->icig >= ncig - 1) { w->icig = -1; w->cig++; continue; }
            w->iseq++; w->icig++; w->iref++;
            return OP_M;
        }
        if (cig == OP_D || cig == OP_N) { w->cig++; w->iref += ncig; w->icig = -1; continue; }
        if (cig == OP_I) { w->cig++; w->iseq += ncig; w->icig = -1; continue; }
        if (cig == OP_S) { w->cig++; w->iseq += ncig; w->icig = -1; continue; }
        if (cig == OP_H || cig == OP_P) { w->cig++; w->icig = -1; continue; }
        return -2;
    }
    w->iseq = -1; w->iref = -1;
    return -1;
}

/* tweak_overlap_quality(a, b): a arrived first.  aq / bq are the two reads' modifiable base qualities. */
static void tweak_overlap_quality(const oreads_t* R, int64_t ra, uint8_t* aq, int64_t rb, uint8_t* bq, int mode) {
    const uint32_t* a_first = R->cigar + R->cigar_off[ra]; const uint32_t* b_first = R->cigar + R->cigar_off[rb];
    cwalk_t a = { a_first, R->cigar + R->cigar_off[ra + 1], 0, 0, 0 };
    cwalk_t b = { b_first, R->cigar + R->cigar_off[rb + 1], 0, 0, 0 };
    const int64_t apos = R->pos[ra], bpos = R->pos[rb];
    const int64_t alq = R->l_seq[ra], blq = R->l_seq[rb];
    int64_t iref = bpos;
    a.iref = iref - apos; b.iref = iref - bpos;
    int a_ret = iref2iseq_set(&a, a.iref);
    if (a_ret < 0) return;
    int b_ret = iref2iseq_set(&b, b.iref);
    if (b_ret < 0) return;
    /* which read keeps its qualities: a hash of the first-arrived read's name (htslib >= 1.13); always a before */
    int amul = 1, bmul = 0;
    if (mode == OLAP_HASH) { if (wang_hash((uint32_t)R->qname_hash[ra]) & 1u) { amul = 1; bmul = 0; } else { amul = 0; bmul = 1; } }
    for (;;) {
        while (a_ret >= 0 && a.iref >= 0 && a.iref < iref - apos) a_ret = iref2iseq_next(&a);
        if (a_ret < 0) break;
        if (iref < a.iref + apos) iref = a.iref + apos;
        while (b_ret >= 0 && b.iref >= 0 && b.iref < iref - bpos) b_ret = iref2iseq_next(&b);
        if (b_ret < 0) break;
        if (iref < b.iref + bpos) iref = b.iref + bpos;
        iref++;
        if (a.iref + apos != b.iref + bpos) {
            if (mode != OLAP_HASH) continue;            /* htslib <= 1.12: only positions both reads match */
            /* a deletion in one read: the other catches up, its bases under the deletion are down-weighted */
            if (a.iref + apos < b.iref + bpos && b.cig > b_first && (int)(*(b.cig - 1) & 15) == OP_D) {
                int done = 0;
                do {
                    if (a.iseq < alq) aq[a.iseq] = amul ? (uint8_t)(aq[a.iseq] * 0.8) : 0;
                    a_ret = iref2iseq_next(&a);
                    if (a_ret < 0) { done = 1; break; }
                } while (a.iref + apos < b.iref + bpos);
                if (done) return;
            } else if (a.cig > a_first && (int)(*(a.cig - 1) & 15) == OP_D) {
                int done = 0;
                do {
                    if (b.iseq < blq) bq[b.iseq] = bmul ? (uint8_t)(bq[b.iseq] * 0.8) : 0;
                    b_ret = iref2iseq_next(&b);
                    if (b_ret < 0) { done = 1; break; }
                } while (b.iref + bpos < a.iref + apos);
                if (done) return;
            } else continue;                            /* anything else, e.g. a reference skip */
        }
        if (a.iseq >= alq || b.iseq >= blq) return;     /* fell off the end of SEQ (upstream: > l_qseq is the error, == reads past it) */
        if (seq_code(R, ra, (int)a.iseq) == seq_code(R, rb, (int)b.iseq)) {
            int qual = aq[a.iseq] + bq[b.iseq];
            if (qual > 200) qual = 200;
            aq[a.iseq] = (uint8_t)(amul * qual);
            bq[b.iseq] = (uint8_t)(bmul * qual);
        } else if (mode == OLAP_HASH) {
            if (aq[a.iseq] > bq[b.iseq]) { aq[a.iseq] = (uint8_t)(0.8 * aq[a.iseq]); bq[b.iseq] = 0; }
            else if (aq[a.iseq] < bq[b.iseq]) { bq[b.iseq] = (uint8_t)(0.8 * bq[b.iseq]); aq[a.iseq] = 0; }
            else { aq[a.iseq] = (uint8_t)(amul * 0.8 * aq[a.iseq]); bq[b.iseq] = (uint8_t)(bmul * 0.8 * bq[b.iseq]); }
        } else {
            if (aq[a.iseq] >= bq[b.iseq]) { aq[a.iseq] = (uint8_t)(0.8 * aq[a.iseq]); bq[b.iseq] = 0; }
            else { bq[b.iseq] = (uint8_t)(0.8 * bq[b.iseq]); aq[a.iseq] = 0; }
        }
    }
}

/* khash olap_hash restated as a plain open-addressing map QNAME hash -> read index of the stored node */
typedef struct { uint64_t* key; int64_t* val; uint8_t* used; int64_t cap, n; } omap_t;
static void omap_init(omap_t* m) { m->cap = 1024; m->n = 0; m->key = malloc(m->cap * 8); m->val = malloc(m->cap * 8); m->used = calloc(m->cap, 1); }
static void omap_free(omap_t* m) { free(m->key); free(m->val); free(m->used); }
static int64_t omap_find(const omap_t* m, uint64_t k) {
    int64_t h = (int64_t)((k * 0x9e3779b97f4a7c15ULL) >> 20) & (m->cap - 1);
    while (m->used[h]) { if (m->key[h] == k) return h; h = (h + 1) & (m->cap - 1); }
    return -1;
}
static void omap_put(omap_t* m, uint64_t k, int64_t v);
static void omap_grow(omap_t* m) {
    omap_t o = *m;
    m->cap = o.cap * 2; m->n = 0; m->key = malloc(m->cap * 8); m->val = malloc(m->cap * 8); m->used = calloc(m->cap, 1);
    for (int64_t i = 0; i < o.cap; ++i) if (o.used[i]) omap_put(m, o.key[i], o.val[i]);
    omap_free(&o);
}
static void omap_put(omap_t* m, uint64_t k, int64_t v) {
    if (2 * (m->n + 1) > m->cap) omap_grow(m);
    int64_t h = (int64_t)((k * 0x9e3779b97f4a7c15ULL) >> 20) & (m->cap - 1);
    while (m->used[h]) { if (m->key[h] == k) { m->val[h] = v; return; } h = (h + 1) & (m->cap - 1); }
    m->used[h] = 1; m->key[h] = k; m->val[h] = v; m->n++;
}
static void omap_del(omap_t* m, uint64_t k) {
    int64_t h = omap_find(m, k);
    if (h < 0) return;
    m->used[h] = 0; m->n--;
    /* re-insert the cluster behind the hole */
    int64_t j = (h + 1) & (m->cap - 1);
    while (m->used[j]) {
        uint64_t kk = m->key[j]; int64_t vv = m->val[j];
        m->used[j] = 0; m->n--;
        omap_put(m, kk, vv);
        j = (j + 1) & (m->cap - 1);
    }
}

/* what resolve_cigar2 reports for (read, column) */
typedef struct { int is_del, is_refskip, indel, qpos; } entry_t;

/* htslib sam.c resolve_cigar2 */
static void resolve(const oreads_t* R, node_t* nd, int32_t pos, entry_t* e) {
    const uint32_t* cig = R->cigar + R->cigar_off[nd->read];
    int n = (int)(R->cigar_off[nd->read + 1] - R->cigar_off[nd->read]);
    int k;
    if (nd->k == -1) {
        /* first visit: skip leading I/S (they advance the query), H/P advance nothing */
        nd->x = R->pos[nd->read]; nd->y = 0;
        for (k = 0; k < n; ++k) {
            int op = cig[k] & 15, l = (int)(cig[k] >> 4);
            if (consumes_ref(op)) break;
            if (op == OP_I || op == OP_S) nd->y += l;
        }
        nd->k = k;
    } else {
        int l = (int)(cig[nd->k] >> 4);
        if (pos - nd->x >= l) {     /* move to the next reference-consuming op */
            if (is_match((int)(cig[nd->k] & 15))) nd->y += l;
            nd->x += l;
            for (k = nd->k + 1; k < n; ++k) {
                int op = cig[k] & 15, l2 = (int)(cig[k] >> 4);
                if (consumes_ref(op)) break;
                if (op == OP_I || op == OP_S) nd->y += l2;
            }
            nd->k = k;
        }
    }
    int op = cig[nd->k] & 15, l = (int)(cig[nd->k] >> 4);
    e->is_del = e->indel = e->is_refskip = 0;
    if (nd->x + l - 1 == pos && nd->k + 1 < n) {    /* last column of this op: peek */
        int op2 = cig[nd->k + 1] & 15, l2 = (int)(cig[nd->k + 1] >> 4);
        if (op2 == OP_D && op != OP_D) {
            e->indel = -l2;
            for (k = nd->k + 2; k < n; ++k) {
                if ((cig[k] & 15) == OP_D) e->indel -= (int)(cig[k] >> 4); else break;
            }
        } else if (op2 == OP_I) {
            e->indel = l2;
            for (k = nd->k + 2; k < n; ++k) {
                int o = cig[k] & 15;
                if (o == OP_I) e->indel += (int)(cig[k] >> 4);
                else if (o != OP_P) break;
            }
        } else if (op2 == OP_P && nd->k + 2 < n) {
            int l3 = 0;
            for (k = nd->k + 2; k < n; ++k) {
                int o = cig[k] & 15;
                if (o == OP_I) l3 += (int)(cig[k] >> 4);
                else if (o == OP_D || o == OP_M || o == OP_N || o == OP_EQ || o == OP_X) break;
            }
            if (l3 > 0) e->indel = l3;
        }
    }
    if (is_match(op)) {
        e->qpos = nd->y + (pos - nd->x);
    } else {    /* D or N */
        e->is_del = 1; e->qpos = nd->y; e->is_refskip = (op == OP_N);
    }
}

typedef struct { char* s; size_t l, cap; } sbuf_t;
static void sb_putc(sbuf_t* b, char c) {
    if (b->l + 1 >= b->cap) { b->cap = b->cap ? 2 * b->cap : 4096; b->s = realloc(b->s, b->cap); }
    b->s[b->l++] = c;
}
static void sb_putw(sbuf_t* b, int v) { char t[16]; int n = snprintf(t, sizeof t, "%d", v); for (int i = 0; i < n; ++i) sb_putc(b, t[i]); }

/* pysam strand_mark_char */
static char strand_char(char c, int rev) {
    if (c == '=') return rev ? ',' : '.';
    return rev ? (char)tolower((unsigned char)c) : (char)toupper((unsigned char)c);
}

/* pysam PileupColumn.get_query_sequences(add_indels=True) for one entry; returns 0 if the entry
 * is skipped by the base-quality test (pileup_base_qual_skip) */
static int entry_string(const oreads_t* R, int64_t rd, const uint8_t* node_q, const entry_t* e, int min_bq, sbuf_t* b) {
    int lq = R->l_seq[rd];
    int rev = (R->flag[rd] & 16) != 0;
    int q = (e->qpos < lq) ? (node_q ? node_q[e->qpos] : R->qual[8ULL * R->seq_off[rd] + e->qpos]) : 0;
    if (q < min_bq) return 0;
    if (!e->is_del) {
        char c = (e->qpos < lq) ? NT16[seq_code(R, rd, e->qpos)] : 'N';
        sb_putc(b, strand_char(c, rev));
    } else if (e->is_refskip) sb_putc(b, rev ? '<' : '>');
    else sb_putc(b, '*');
    if (e->indel > 0) {
        sb_putc(b, '+'); sb_putw(b, e->indel);
        for (int j = 1; j <= e->indel; ++j) {
            /* bases qpos+1 .. qpos+indel (upstream does not range-check; past the end reads as 'N' here) */
            int qq = e->qpos + j;
            char c = (qq < lq) ? NT16[seq_code(R, rd, qq)] : 'N';
            sb_putc(b, strand_char(c, rev));
        }
    } else if (e->indel < 0) {
        sb_putc(b, '-'); sb_putw(b, -e->indel);
        for (int j = 0; j < -e->indel; ++j) sb_putc(b, strand_char('N', rev));
    }
    return 1;
}

/* TrueConsense/indexing.py:102-132 parse_query_sequences, one string at a time.
 * row order of out[]: coverage, A, T, C, G, X, I  (indexing.py:134) */
static void classify_string(const char* s, size_t n, int64_t* out) {
    out[0] += 1;                                            /* coverage += 1          :117 */
    if (n == 1 && s[0] == '*') out[5] += 1;                 /* b == "*"  -> x         :118 */
    else if (n > 0) {
        char c = (char)tolower((unsigned char)s[0]);        /* b[0].lower()           :120-127 */
        if (c == 'a') out[1] += 1;
        else if (c == 't') out[2] += 1;
        else if (c == 'c') out[3] += 1;
        else if (c == 'g') out[4] += 1;
    }
    if (memchr(s, '+', n)) out[6] += 1;                     /* "+" in b -> i          :130 */
}

typedef void (*column_cb)(void* ud, int32_t pos, const char* strs, size_t len, int n_entries);

/* The streaming engine: htslib bam_plp_push + bam_plp64_next driven by pysam's stepper.
 * Reads [r0, r1) are offered in order.  If region_start >= 0 only reads overlapping
 * [region_start, region_end) are fetched and only those columns are reported (truncate=True). */
static int run_engine(const oreads_t* R, int64_t r0, int64_t r1, const oparams_t* P,
                      int32_t region_start, int32_t region_end, column_cb cb, void* ud) {
    int64_t cap = 1024, nlive = 0;
    node_t* live = malloc(cap * sizeof(node_t));
    /* pysam's ignore_overlaps=True (its default at both call sites); invisible without a base-quality filter */
    const int olap_mode = (P->reserved >> 8) & 3;
    const int olap = olap_mode != OLAP_OFF && P->min_base_quality > 0 && R->qname_hash && R->mpos && R->isize;
    omap_t omap;
    if (olap) omap_init(&omap);
    sbuf_t sb = {0, 0, 0};
    int32_t it_pos = 0, max_pos = -1;
    int64_t cnt = 1;            /* htslib mempool count: the tail sentinel */
    int rc = 0;
    int64_t i = r0;
    int eof = 0;
    for (;;) {
        /* emit every column that can no longer receive reads (bam_plp64_next) */
        while (eof || max_pos > it_pos) {
            if (eof && nlive == 0) break;
            int n_plp = 0; sb.l = 0;
            int64_t w = 0;
            for (int64_t j = 0; j < nlive; ++j) {
                node_t* nd = &live[j];
                if (nd->end <= it_pos) {                        /* finished: drop (lazy free; overlap_remove by name) */
                    if (olap) omap_del(&omap, R->qname_hash[nd->read]);
                    free(nd->q);
                    --cnt; continue;
                }
                if (nd->beg <= it_pos) {
                    entry_t e;
                    resolve(R, nd, it_pos, &e);
                    size_t before = sb.l;
                    if (entry_string(R, nd->read, nd->q, &e, P->min_base_quality, &sb)) { sb_putc(&sb, ':'); ++n_plp; }
                    else sb.l = before;
                }
                live[w++] = *nd;
            }
            nlive = w;
            int32_t col = it_pos;
            if (nlive > 0 && it_pos < live[0].beg) it_pos = live[0].beg; else ++it_pos;
            /* htslib yields a column when at least one read overlaps it, pysam then drops the
             * entries failing the quality test; a column whose entries all fail still exists
             * (get_query_sequences returns "") */
            if (cb && (region_start < 0 || (col >= region_start && col < region_end))) {
                int any = 0;
                for (int64_t j = 0; j < nlive; ++j) if (live[j].beg <= col && live[j].end > col) { any = 1; break; }
                if (any) cb(ud, col, sb.s, sb.l ? sb.l - 1 : 0, n_plp);
            }
        }
        if (eof) break;
        /* fetch the next read through the stepper */
        int64_t rd = -1;
        while (i < r1) {
            int64_t c = i++;
            uint16_t fl = R->flag[c];
            if (region_start >= 0) {
                /* BAI region query: pos < end && endpos > start (rlen 0 counts as 1) */
                int32_t span = 0;
                for (uint32_t k = R->cigar_off[c]; k < R->cigar_off[c + 1]; ++k)
                    if (consumes_ref((int)(R->cigar[k] & 15))) span += (int32_t)(R->cigar[k] >> 4);
                int32_t endpos = R->pos[c] + (span > 0 ? span : 1);
                if (!(R->pos[c] < region_end && endpos > region_start)) continue;
            }
            if (fl & P->flag_filter) continue;                                  /* stepper flag filter */
            if ((int)R->mapq[c] < P->min_mapq) continue;
            if (P->ignore_orphans && (fl & 1) && !(fl & 2)) continue;
            rd = c; break;
        }
        if (rd < 0) { eof = 1; continue; }
        /* bam_plp_push */
        if (R->flag[rd] & 4) { if (olap) omap_del(&omap, R->qname_hash[rd]); continue; }       /* htslib always drops UNMAP */
        if (it_pos == R->pos[rd] && cnt > P->max_depth) {                       /* bam_plp_set_maxcnt */
            if (olap) omap_del(&omap, R->qname_hash[rd]);                       /* overlap_remove: by name, whichever mate is stored */
            continue;
        }
        int32_t span = 0;
        for (uint32_t k = R->cigar_off[rd]; k < R->cigar_off[rd + 1]; ++k)
            if (consumes_ref((int)(R->cigar[k] & 15))) span += (int32_t)(R->cigar[k] >> 4);
        if (R->pos[rd] < max_pos) { rc = -3; break; }                           /* "Unsorted input. Pileup aborts" */
        max_pos = R->pos[rd];
        if (R->pos[rd] + span > it_pos) {
            if (nlive == cap) { cap *= 2; live = realloc(live, cap * sizeof(node_t)); }
            node_t nd; nd.read = rd; nd.beg = R->pos[rd]; nd.end = R->pos[rd] + span; nd.k = -1; nd.x = nd.y = 0; nd.q = NULL;
            live[nlive++] = nd; ++cnt;
            /* overlap_push: proper pairs with a mapped mate on this reference that can overlap */
            uint16_t fl = R->flag[rd];
            int32_t mp = R->mpos[rd];               /* -1: unavailable, -2: the mate is on another reference */
            int64_t isz = R->isize[rd]; if (isz < 0) isz = -isz;
            if (olap && !(fl & 8) && (fl & 2) && mp != -2 && !(isz >= 2 * (int64_t)R->l_seq[rd] && mp >= nd.end)) {
                int64_t h = omap_find(&omap, R->qname_hash[rd]);
                if (h < 0) {
                    if (mp >= R->pos[rd] || ((fl & 1) && mp == -1)) omap_put(&omap, R->qname_hash[rd], rd);    /* the mate is still to arrive */
                } else {
                    int64_t ra = omap.val[h];
                    node_t* na = NULL;
                    for (int64_t j = 0; j < nlive - 1; ++j) if (live[j].read == ra) { na = &live[j]; break; }
                    if (na) {
                        node_t* nb = &live[nlive - 1];
                        if (!na->q) { int lq = R->l_seq[ra]; na->q = malloc(lq > 0 ? lq : 1); memcpy(na->q, R->qual + 8ULL * R->seq_off[ra], lq); }
                        if (!nb->q) { int lq = R->l_seq[rd]; nb->q = malloc(lq > 0 ? lq : 1); memcpy(nb->q, R->qual + 8ULL * R->seq_off[rd], lq); }
                        tweak_overlap_quality(R, ra, na->q, rd, nb->q, olap_mode);
                    }
                    omap_del(&omap, R->qname_hash[rd]);
                }
            }
        }
    }
    for (int64_t j = 0; j < nlive; ++j) free(live[j].q);
    if (olap) omap_free(&omap);
    free(live); free(sb.s);
    return rc;
}

/* ---------------------------------------------------------------- count tables */
typedef struct { int64_t* t; int32_t L; int err; } count_ud;

static void count_cb(void* ud_, int32_t pos, const char* s, size_t len, int n_entries) {
    count_ud* ud = ud_;
    if (pos < 0 || pos >= ud->L) { ud->err = 1; return; }
    if (n_entries == 0) return;     /* "" -> the reference's classifier sees an empty iterable of chars */
    int64_t row[7] = {0, 0, 0, 0, 0, 0, 0};
    size_t a = 0;
    for (size_t j = 0; j <= len; ++j) {
        if (j == len || s[j] == ':') { classify_string(s + a, j - a, row); a = j + 1; }
    }
    for (int r = 0; r < 7; ++r) ud->t[(size_t)r * ud->L + pos] += row[r];
}

/* BuildIndex's table: counts[8][L] int32 (rows coverage,A,T,C,G,X,I,pad).  n_threads > 1 splits
 * the reads into contiguous ranges (exact while max_depth does not bind). */
int oracle_pileup_counts(const oreads_t* R, int32_t L, const oparams_t* P, int32_t* counts, int n_threads) {
#ifdef _OPENMP
    if (n_threads <= 0) n_threads = omp_get_num_procs();
#else
    n_threads = 1;
#endif
    if (R->n_reads < 4096) n_threads = 1;
    int64_t* acc = calloc((size_t)n_threads * 7 * (size_t)L, sizeof(int64_t));
    int rc = 0;
#pragma omp parallel for num_threads(n_threads) schedule(static, 1)
    for (int t = 0; t < n_threads; ++t) {
        int64_t r0 = R->n_reads * t / n_threads, r1 = R->n_reads * (t + 1) / n_threads;
        count_ud ud; ud.t = acc + (size_t)t * 7 * L; ud.L = L; ud.err = 0;
        int r = run_engine(R, r0, r1, P, -1, -1, count_cb, &ud);
        if (r) rc = r;
        if (ud.err) rc = -7;
    }
    memset(counts, 0, sizeof(int32_t) * 8 * (size_t)L);
    for (int t = 0; t < n_threads; ++t)
        for (int r = 0; r < 7; ++r)
            for (int32_t p = 0; p < L; ++p)
                counts[(size_t)r * L + p] += (int32_t)acc[((size_t)t * 7 + r) * L + p];
    free(acc);
    return rc;
}

/* ---------------------------------------------------------------- column strings (pysam stub) */
typedef struct { sbuf_t out; int32_t* pos; int64_t* off; int32_t* n_ent; int64_t n, cap; } str_ud;

static void str_cb(void* ud_, int32_t pos, const char* s, size_t len, int n_entries) {
    str_ud* ud = ud_;
    if (ud->n == ud->cap) {
        ud->cap = ud->cap ? 2 * ud->cap : 1024;
        ud->pos = realloc(ud->pos, ud->cap * sizeof(int32_t));
        ud->off = realloc(ud->off, (ud->cap + 1) * sizeof(int64_t));
        ud->n_ent = realloc(ud->n_ent, ud->cap * sizeof(int32_t));
    }
    ud->pos[ud->n] = pos; ud->off[ud->n] = (int64_t)ud->out.l; ud->n_ent[ud->n] = n_entries;
    for (size_t j = 0; j < len; ++j) sb_putc(&ud->out, s[j]);
    ud->n++;
}

/* All columns of [region_start, region_end) (or of the whole input when region_start < 0) as
 * ':'-joined strings.  Outputs are malloc'ed; release with oracle_free. */
int oracle_pileup_strings(const oreads_t* R, const oparams_t* P, int32_t region_start, int32_t region_end,
                          char** buf, int64_t* buf_len, int32_t** col_pos, int64_t** col_off, int32_t** col_n,
                          int64_t* n_cols) {
    str_ud ud; memset(&ud, 0, sizeof ud);
    int rc = run_engine(R, 0, R->n_reads, P, region_start, region_end, str_cb, &ud);
    if (ud.cap == 0) { ud.off = malloc(sizeof(int64_t)); ud.pos = malloc(sizeof(int32_t)); ud.n_ent = malloc(sizeof(int32_t)); }
    ud.off[ud.n] = (int64_t)ud.out.l;
    if (!ud.out.s) ud.out.s = malloc(1);
    *buf = ud.out.s; *buf_len = (int64_t)ud.out.l; *col_pos = ud.pos; *col_off = ud.off; *col_n = ud.n_ent; *n_cols = ud.n;
    return rc;
}

void oracle_free(void* p) { free(p); }
