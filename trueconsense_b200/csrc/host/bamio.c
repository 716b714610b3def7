/*
 * bamio.c — BGZF/BAM reader and writer producing / consuming the flat read arrays of
 * tc_reads_t.  Replaces what pysam.AlignmentFile does for TrueConsense/indexing.py:96
 * (the decode half only; the pileup itself is the GPU's job).
 *
 * On-disk format (SAM/BAM specification v1, section 4): a BAM file is a series of BGZF
 * blocks — gzip members of at most 64 KiB with a "BC" extra subfield holding the block
 * size — whose concatenated payload is: magic "BAM\1", l_text, text, n_ref, per reference
 * (l_name, name, l_ref), then records (block_size, refID, pos, l_read_name, mapq, bin,
 * n_cigar_op, flag, l_seq, next_refID, next_pos, tlen, read_name, cigar[u32], seq[4-bit],
 * qual, tags).
 *
 * Reader: (1) hop over block headers to index the file, (2) inflate all blocks in parallel
 * (OpenMP, raw deflate via zlib) into one buffer, (3) hop over records to size the outputs,
 * (4) fill the flat arrays in parallel.
 */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdarg.h>
#include <stdint.h>
#include <fcntl.h>
#include <unistd.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <zlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "tc_host.h"

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

static int fail(char* err, int errlen, int code, const char* fmt, ...) {
    if (err && errlen > 0) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(err, errlen, fmt, ap);
        va_end(ap);
    }
    return code;
}

static inline uint16_t rd16(const uint8_t* p) { return (uint16_t)(p[0] | (p[1] << 8)); }
static inline uint32_t rd32(const uint8_t* p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
static inline void wr16(uint8_t* p, uint16_t v) { p[0] = v & 0xff; p[1] = v >> 8; }
static inline void wr32(uint8_t* p, uint32_t v) {
    p[0] = v & 0xff; p[1] = (v >> 8) & 0xff; p[2] = (v >> 16) & 0xff; p[3] = v >> 24;
}

/* khash's string hash (htslib keys its mate-overlap table by QNAME with it) */
static uint32_t x31_hash(const char* s) {
    uint32_t h = (uint32_t)*s;
    if (h) for (++s; *s; ++s) h = (h << 5) - h + (uint32_t)*s;
    return h;
}
static uint32_t fnv1a_hi(const char* s) {
    uint64_t h = 1469598103934665603ULL;
    for (; *s; ++s) { h ^= (uint8_t)*s; h *= 1099511628211ULL; }
    return (uint32_t)(h >> 32) ^ (uint32_t)h;
}

void tc_hostreads_free(tc_hostreads_t* r) {
    if (!r) return;
    free(r->pos); free(r->flag); free(r->mapq); free(r->l_seq); free(r->seq_off);
    free(r->cigar_off); free(r->seq4); free(r->qual); free(r->cigar); free(r->qname_hash);
    free(r->mpos); free(r->isize); free(r->tid); free(r->mtid); free(r->ref_len);
    free(r->ref_names);
    memset(r, 0, sizeof(*r));
}

static int alloc_reads(tc_hostreads_t* o, int64_t n, int64_t n_words, int64_t n_ops) {
    o->n_reads = n; o->n_seq_words = n_words; o->n_cigar_ops = n_ops;
    size_t n1 = (size_t)(n > 0 ? n : 1);
    o->pos = malloc(n1 * 4); o->flag = malloc(n1 * 2); o->mapq = malloc(n1);
    o->l_seq = malloc(n1 * 4); o->seq_off = malloc((n1 + 1) * 4); o->cigar_off = malloc((n1 + 1) * 4);
    o->seq4 = calloc((size_t)(n_words > 0 ? n_words : 1), 4);
    o->qual = calloc((size_t)(n_words > 0 ? n_words : 1), 8);
    o->cigar = malloc((size_t)(n_ops > 0 ? n_ops : 1) * 4);
    o->qname_hash = malloc(n1 * 8); o->mpos = malloc(n1 * 4); o->isize = malloc(n1 * 4);
    o->tid = malloc(n1 * 4); o->mtid = malloc(n1 * 4);
    if (!o->pos || !o->flag || !o->mapq || !o->l_seq || !o->seq_off || !o->cigar_off || !o->seq4 ||
        !o->qual || !o->cigar || !o->qname_hash || !o->mpos || !o->isize || !o->tid || !o->mtid)
        return -1;
    o->seq_off[0] = 0; o->cigar_off[0] = 0;
    return 0;
}

/* exported for synth.c */
int tc_hostreads_alloc_(tc_hostreads_t* o, int64_t n, int64_t n_words, int64_t n_ops) {
    return alloc_reads(o, n, n_words, n_ops);
}

typedef tc_bgzf_block_t blk_t;

/* header + record hop over the payload while it is still being inflated */
typedef struct {
    const uint8_t* u; int64_t utotal; const blk_t* blk; int64_t nblk; const uint8_t* done; const int* bad;
    int64_t next_b, ready;          /* blocks [0, next_b) are known complete: payload bytes [0, ready) may be read */
    const char* path; char* err; int errlen; tc_hostreads_t* out;
    int64_t* recoff; int64_t nkept; int rc;
} scan_t;

/* payload bytes [0, upto) complete?  Spins behind the inflating threads; 0 when an inflate failed. */
static int scan_wait(scan_t* s, int64_t upto) {
    if (upto > s->utotal) upto = s->utotal;
    while (s->ready < upto) {
        if (s->next_b >= s->nblk) { s->ready = s->utotal; break; }
        if (__atomic_load_n(&s->done[s->next_b], __ATOMIC_ACQUIRE)) {
            s->ready = s->blk[s->next_b].uoff + s->blk[s->next_b].usize;
            s->next_b++;
        } else if (__atomic_load_n(s->bad, __ATOMIC_ACQUIRE)) return 0;
        else {
#if defined(__x86_64__) || defined(__i386__)
            __builtin_ia32_pause();
#endif
        }
    }
    return 1;
}
#define SCAN_NEED(upto) do { if ((upto) > s->ready && !scan_wait(s, (upto))) return -2; } while (0)

static int scan_payload(scan_t* s) {
    const uint8_t* u = s->u; const int64_t utotal = s->utotal; tc_hostreads_t* out = s->out;
    char* err = s->err; const int errlen = s->errlen; const char* path = s->path;
    SCAN_NEED(12);
    if (utotal < 12 || memcmp(u, "BAM\1", 4) != 0) return fail(err, errlen, -2, "%s: missing BAM magic", path);
    int64_t p = 4;
    int32_t l_text = (int32_t)rd32(u + p);
    if (l_text < 0 || p + 4 + (int64_t)l_text + 4 > utotal) return fail(err, errlen, -2, "truncated BAM header");
    p += 4 + l_text;
    SCAN_NEED(p + 4);
    int32_t n_ref = (int32_t)rd32(u + p); p += 4;
    if (n_ref < 0 || (int64_t)n_ref * 8 > utotal - p) return fail(err, errlen, -2, "bad BAM reference count %d", n_ref);
    out->n_ref = n_ref;
    out->ref_len = malloc(sizeof(int32_t) * (n_ref > 0 ? n_ref : 1));
    int64_t names_cap = 64, names_len = 0;
    out->ref_names = malloc(names_cap);
    if (!out->ref_len || !out->ref_names) return fail(err, errlen, -5, "out of memory");
    for (int i = 0; i < n_ref; ++i) {
        if (p + 4 > utotal) return fail(err, errlen, -2, "truncated BAM reference list");
        SCAN_NEED(p + 4);
        int32_t l_name = (int32_t)rd32(u + p); p += 4;
        if (l_name < 0 || p + (int64_t)l_name + 4 > utotal) return fail(err, errlen, -2, "truncated BAM reference list");
        SCAN_NEED(p + l_name + 4);
        while (names_len + l_name + 1 > names_cap) {
            names_cap *= 2;
            char* nn = realloc(out->ref_names, names_cap);
            if (!nn) return fail(err, errlen, -5, "out of memory");
            out->ref_names = nn;
        }
        memcpy(out->ref_names + names_len, u + p, l_name);
        names_len += l_name;
        if (l_name == 0 || out->ref_names[names_len - 1] != 0) out->ref_names[names_len++] = 0;
        p += l_name;
        out->ref_len[i] = (int32_t)rd32(u + p); p += 4;
    }
    out->ref_names_len = names_len;

    int64_t rcap = 1 << 16, nrec = 0, nkept = 0;
    s->recoff = malloc(rcap * sizeof(int64_t));
    if (!s->recoff) return fail(err, errlen, -5, "out of memory");
    int64_t q = p;
    while (q + 4 <= utotal) {
        SCAN_NEED(q + 4);
        int32_t bs = (int32_t)rd32(u + q);
        if (bs < 32 || q + 4 + bs > utotal) return fail(err, errlen, -2, "corrupt BAM record at payload offset %lld", (long long)q);
        ++nrec;
        SCAN_NEED(q + 4 + 36);
        int32_t refid = (int32_t)rd32(u + q + 4);
        {
            /* the record's own sizes against its block_size, before anything is sized from them */
            const uint8_t* r = u + q + 4;
            int32_t l_name = r[8];
            uint32_t n_cig = rd16(r + 12);
            uint32_t l_seq = rd32(r + 16);
            int64_t need = 32 + (int64_t)l_name + 4LL * n_cig + ((int64_t)l_seq + 1) / 2 + (int64_t)l_seq;
            if (l_seq > 0x7fffffffu || need > bs || l_name < 1)
                return fail(err, errlen, -2, "%s: BAM record %lld shorter than its fields", path, (long long)(nrec - 1));
            SCAN_NEED(q + 4 + 32 + l_name + 8);
            if (r[32 + l_name - 1] != 0)
                return fail(err, errlen, -2, "%s: BAM record %lld shorter than its fields (or an unterminated read name)", path, (long long)(nrec - 1));
            /* more than 65535 CIGAR ops: the real CIGAR sits in the CG:B,I tag behind a <l_seq>S<span>N placeholder, which
             * htslib expands transparently.  Not expanded here: refuse, rather than pile the read up as a reference skip */
            if (refid >= 0 && n_cig == 2) {
                const uint8_t* cg = r + 32 + l_name;
                uint32_t c0 = rd32(cg), c1 = rd32(cg + 4);
                if ((c0 & 15) == 4 && (c0 >> 4) == l_seq && (c1 & 15) == 3 && l_seq > 0)
                    return fail(err, errlen, -2, "%s: record %lld keeps its CIGAR in a CG tag (more than 65535 operations): not supported", path, (long long)(nrec - 1));
            }
        }
        if (refid >= 0) {
            if (nkept == rcap) {
                rcap *= 2;
                int64_t* nr2 = realloc(s->recoff, rcap * sizeof(int64_t));
                if (!nr2) return fail(err, errlen, -5, "out of memory");
                s->recoff = nr2;
            }
            s->recoff[nkept++] = q + 4;
        }
        q += 4 + bs;
    }
    out->n_records = nrec;
    out->n_dropped_unplaced = nrec - nkept;
    s->nkept = nkept;
    return 0;
}


/* (1) of every reader: walk the BGZF member headers (each says how long it is) — where every raw DEFLATE stream sits in the
 * file, how many bytes it inflates to, and where those go in the payload.  Every size comes from the file: nothing is
 * used before it has been checked against the file and the format. */
static int bgzf_index_blocks(const uint8_t* f, int64_t fsize, blk_t** blk_out, int64_t* nblk_out, int64_t* utotal_out, char* err, int errlen) {
    int64_t nblk = 0, cap = 1024;
    blk_t* blk = malloc(cap * sizeof(blk_t));
    if (!blk) return fail(err, errlen, -5, "out of memory");
    int64_t off = 0, uoff = 0;
    int rc = 0;
    while (off < fsize) {
        if (off + 18 > fsize) { rc = fail(err, errlen, -2, "truncated BGZF header at %lld", (long long)off); break; }
        const uint8_t* h = f + off;
        if (h[0] != 31 || h[1] != 139 || h[2] != 8 || !(h[3] & 4)) {
            rc = fail(err, errlen, -2, "not a BGZF block at offset %lld", (long long)off); break;
        }
        int xlen = rd16(h + 10);
        if (off + 12 + xlen > fsize) { rc = fail(err, errlen, -2, "truncated BGZF extra field at %lld", (long long)off); break; }
        int bsize = -1;
        for (int x = 0; x + 4 <= xlen;) {
            const uint8_t* e = h + 12 + x;
            int slen = rd16(e + 2);
            if (x + 4 + slen > xlen) break;                      /* a subfield running past XLEN */
            if (e[0] == 'B' && e[1] == 'C' && slen == 2) bsize = rd16(e + 4) + 1;
            x += 4 + slen;
        }
        /* header (12) + extra + at least an empty deflate stream (2) + CRC32 + ISIZE (8) */
        if (bsize < 12 + xlen + 2 + 8 || off + bsize > fsize) { rc = fail(err, errlen, -2, "bad BGZF BSIZE at %lld", (long long)off); break; }
        uint32_t isize = rd32(h + bsize - 4);
        if (isize > 65536u) { rc = fail(err, errlen, -2, "bad BGZF ISIZE %u at %lld (a block inflates to at most 64 KiB)", isize, (long long)off); break; }
        int32_t usize = (int32_t)isize;
        if (nblk == cap) {
            cap *= 2;
            blk_t* nb = realloc(blk, cap * sizeof(blk_t));
            if (!nb) { rc = fail(err, errlen, -5, "out of memory"); break; }
            blk = nb;
        }
        blk[nblk].coff = off + 12 + xlen;
        blk[nblk].csize = bsize - 12 - xlen - 8;
        blk[nblk].uoff = uoff; blk[nblk].usize = usize;
        ++nblk; off += bsize; uoff += usize;
    }
    if (rc) { free(blk); return rc; }
    *blk_out = blk; *nblk_out = nblk; *utotal_out = uoff;
    return 0;
}

/* Stages (1)-(3): the file's uncompressed payload, its header in *out, and the offsets of the placed records (of their
 * refID field, i.e. behind block_size).  The caller owns *u_out and *recoff_out (free). */
static int bam_payload_stage(const char* path, int n_threads, tc_hostreads_t* out, uint8_t** u_out, int64_t* utotal_out,
                             int64_t** recoff_out, int64_t* nkept_out, char* err, int errlen) {
    memset(out, 0, sizeof(*out));
#ifdef _OPENMP
    if (n_threads <= 0) n_threads = omp_get_num_procs();
#else
    n_threads = 1;
#endif
    int fd = open(path, O_RDONLY);
    if (fd < 0) return fail(err, errlen, -1, "cannot open %s", path);
    struct stat st;
    if (fstat(fd, &st) != 0) { close(fd); return fail(err, errlen, -1, "cannot stat %s", path); }
    int64_t fsize = st.st_size;
    if (fsize < 28) { close(fd); return fail(err, errlen, -2, "%s: too short for a BAM file", path); }
    const uint8_t* f = mmap(NULL, (size_t)fsize, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (f == MAP_FAILED) return fail(err, errlen, -1, "mmap failed for %s", path);

    double t0 = now_s();
    /* (1) index BGZF blocks */
    blk_t* blk = NULL;
    int64_t nblk = 0, utotal = 0;
    int rc = bgzf_index_blocks(f, fsize, &blk, &nblk, &utotal, err, errlen);
    if (rc) { munmap((void*)f, fsize); return rc; }
    uint8_t* u = malloc((size_t)(utotal > 0 ? utotal : 1));
    if (!u) { free(blk); munmap((void*)f, fsize); return fail(err, errlen, -5, "out of memory (%lld bytes)", (long long)utotal); }

    /* (2) inflate in parallel and, behind the inflating threads, (3) header + one hop over the records.  The hop is sequential
     * by nature (every record says how long it is) and latency-bound (one cache miss per record): run by one thread as the
     * blocks in front of it complete, it costs no wall time of its own. */
    int bad = 0;
    int64_t next_block = 0;
    uint8_t* done = calloc((size_t)(nblk > 0 ? nblk : 1), 1);
    scan_t sc;
    memset(&sc, 0, sizeof(sc));
    sc.u = u; sc.utotal = utotal; sc.blk = blk; sc.nblk = nblk; sc.done = done; sc.bad = &bad; sc.path = path; sc.err = err; sc.errlen = errlen; sc.out = out;
    if (!done) { free(blk); free(u); munmap((void*)f, fsize); return fail(err, errlen, -5, "out of memory"); }
    double t_inflate_end = t0;
#pragma omp parallel num_threads(n_threads)
    {
        int tid = 0, nth = 1;
#ifdef _OPENMP
        tid = omp_get_thread_num(); nth = omp_get_num_threads();
#endif
        if (tid == 0 && nth > 1) sc.rc = scan_payload(&sc);            /* waits for the blocks it needs */
        for (;;) {
            int64_t b = __atomic_fetch_add(&next_block, 1, __ATOMIC_RELAXED);
            if (b >= nblk) break;
            int ok = 1;
            if (blk[b].usize != 0) {
                z_stream zs;
                memset(&zs, 0, sizeof(zs));
                if (inflateInit2(&zs, -15) != Z_OK) ok = 0;
                else {
                    zs.next_in = (Bytef*)(f + blk[b].coff); zs.avail_in = (uInt)blk[b].csize;
                    zs.next_out = u + blk[b].uoff; zs.avail_out = (uInt)blk[b].usize;       /* inflate never writes more than ISIZE bytes */
                    int r = inflate(&zs, Z_FINISH);
                    ok = (r == Z_STREAM_END && zs.avail_out == 0);
                    inflateEnd(&zs);
                    if (ok) {
                        uint32_t crc = crc32(crc32(0L, Z_NULL, 0), u + blk[b].uoff, blk[b].usize);
                        if (crc != rd32(f + blk[b].coff + blk[b].csize)) ok = 0;
                    }
                }
            }
            if (!ok) __atomic_store_n(&bad, 1, __ATOMIC_RELEASE);
            __atomic_store_n(&done[b], 1, __ATOMIC_RELEASE);
        }
#pragma omp barrier
#pragma omp master
        {
            t_inflate_end = now_s();
            if (nth == 1 && !bad) sc.rc = scan_payload(&sc);           /* a single thread: one after the other */
        }
    }
    free(blk); free(done);
    munmap((void*)f, fsize);
    if (bad) { free(sc.recoff); free(u); tc_hostreads_free(out); return fail(err, errlen, -2, "%s: BGZF inflate / CRC failure", path); }
    if (sc.rc) { free(sc.recoff); free(u); tc_hostreads_free(out); return sc.rc; }
    out->t_inflate_s = t_inflate_end - t0;
    int64_t* recoff = sc.recoff; int64_t nkept = sc.nkept;
    *u_out = u; *utotal_out = utotal; *recoff_out = recoff; *nkept_out = nkept;
    return 0;
}

int tc_bam_read(const char* path, int n_threads, tc_hostreads_t* out, char* err, int errlen) {
#ifdef _OPENMP
    if (n_threads <= 0) n_threads = omp_get_num_procs();
#else
    n_threads = 1;
#endif
    uint8_t* u; int64_t utotal; int64_t* recoff; int64_t nkept;
    int rc = bam_payload_stage(path, n_threads, out, &u, &utotal, &recoff, &nkept, err, errlen);
    if (rc) return rc;
    double t0 = now_s();

    /* sizes -> offsets (sequential prefix sum, cheap) */
    uint32_t* soff = malloc((nkept + 1) * 4);
    uint32_t* coff = malloc((nkept + 1) * 4);
    if (!soff || !coff) { free(soff); free(coff); free(recoff); free(u); tc_hostreads_free(out); return fail(err, errlen, -5, "out of memory"); }
    uint64_t sw = 0, co = 0;
    for (int64_t i = 0; i < nkept; ++i) {
        const uint8_t* r = u + recoff[i];
        soff[i] = (uint32_t)sw; coff[i] = (uint32_t)co;
        uint32_t l_seq = rd32(r + 16);
        sw += (l_seq + 7) / 8;
        co += rd16(r + 12);
    }
    soff[nkept] = (uint32_t)sw; coff[nkept] = (uint32_t)co;
    if (sw > 0xffffffffULL || co > 0xffffffffULL) {
        free(recoff); free(u); free(soff); free(coff); tc_hostreads_free(out);
        return fail(err, errlen, -7, "read batch too large for 32-bit offsets; shard the input");
    }
    {
        int32_t n_ref_keep = out->n_ref; int32_t* rl = out->ref_len; char* rn = out->ref_names; int64_t rnl = out->ref_names_len;
        int64_t nr = out->n_records, nd = out->n_dropped_unplaced; double ti = out->t_inflate_s;
        if (alloc_reads(out, nkept, (int64_t)sw, (int64_t)co) != 0) {
            free(recoff); free(u); free(soff); free(coff);
            out->ref_len = rl; out->ref_names = rn; tc_hostreads_free(out);
            return fail(err, errlen, -5, "out of memory");
        }
        out->n_ref = n_ref_keep; out->ref_len = rl; out->ref_names = rn; out->ref_names_len = rnl;
        out->n_records = nr; out->n_dropped_unplaced = nd; out->t_inflate_s = ti;
    }
    memcpy(out->seq_off, soff, (nkept + 1) * 4);
    memcpy(out->cigar_off, coff, (nkept + 1) * 4);
    free(soff); free(coff);

    /* (4) fill */
    int64_t aligned = 0; int maxspan = 0; int badrec = 0;
#pragma omp parallel for schedule(static) num_threads(n_threads) reduction(+:aligned) reduction(max:maxspan) reduction(|:badrec)
    for (int64_t i = 0; i < nkept; ++i) {
        const uint8_t* r = u + recoff[i];
        int32_t bs = (int32_t)rd32(r - 4);
        int32_t l_name = r[8];
        uint32_t n_cig = rd16(r + 12);
        uint32_t l_seq = rd32(r + 16);
        int64_t need = 32 + (int64_t)l_name + 4LL * n_cig + (l_seq + 1) / 2 + l_seq;
        if (need > bs) { badrec = 1; continue; }
        out->tid[i] = (int32_t)rd32(r);
        out->pos[i] = (int32_t)rd32(r + 4);
        out->mapq[i] = r[9];
        out->flag[i] = rd16(r + 14);
        out->l_seq[i] = (int32_t)l_seq;
        out->mtid[i] = (int32_t)rd32(r + 20);
        out->mpos[i] = (int32_t)rd32(r + 24);
        out->isize[i] = (int32_t)rd32(r + 28);
        const char* name = (const char*)(r + 32);
        out->qname_hash[i] = ((uint64_t)fnv1a_hi(name) << 32) | x31_hash(name);
        const uint8_t* cg = r + 32 + l_name;
        uint32_t* cdst = out->cigar + out->cigar_off[i];
        int span = 0;
        for (uint32_t k = 0; k < n_cig; ++k) {
            uint32_t c = rd32(cg + 4 * k);
            cdst[k] = c;
            uint32_t op = c & 15, len = c >> 4;
            if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) span += (int)len;
        }
        aligned += span;
        if (span > maxspan) maxspan = span;
        const uint8_t* sq = cg + 4 * n_cig;
        memcpy((uint8_t*)(out->seq4 + out->seq_off[i]), sq, (l_seq + 1) / 2);
        memcpy(out->qual + 8ULL * out->seq_off[i], sq + (l_seq + 1) / 2, l_seq);
    }
    free(recoff); free(u);
    if (badrec) { tc_hostreads_free(out); return fail(err, errlen, -2, "%s: BAM record shorter than its fields", path); }
    out->aligned_bases = aligned;
    out->max_ref_span = maxspan;
    int sorted = 1;
    for (int64_t i = 1; i < nkept; ++i) {
        if (out->tid[i] < out->tid[i - 1] || (out->tid[i] == out->tid[i - 1] && out->pos[i] < out->pos[i - 1])) { sorted = 0; break; }
    }
    out->sorted = sorted;
    out->t_parse_s = now_s() - t0;
    return 0;
}

/* The decode split for GPU-side record parsing (trueconsense_b200.h tc_bam_records_to_reads): inflate on the host's cores,
 * hop over the records once (sequential by nature: every record says how long it is), and hand the raw payload plus the
 * record offsets to the device — which parses the fixed fields, hashes the names, walks the CIGARs and repacks SEQ / QUAL /
 * CIGAR into the flat arrays. */
int tc_bam_payload(const char* path, int n_threads, tc_bampayload_t* out, char* err, int errlen) {
    memset(out, 0, sizeof(*out));
#ifdef _OPENMP
    if (n_threads <= 0) n_threads = omp_get_num_procs();
#else
    n_threads = 1;
#endif
    tc_hostreads_t hdr;
    uint8_t* u; int64_t utotal; int64_t* recoff; int64_t nkept;
    double t0 = now_s();
    int rc = bam_payload_stage(path, n_threads, &hdr, &u, &utotal, &recoff, &nkept, err, errlen);
    if (rc) return rc;
    out->payload = u; out->n_bytes = utotal; out->rec_off = recoff; out->n_reads = nkept;
    out->n_records = hdr.n_records; out->n_dropped_unplaced = hdr.n_dropped_unplaced;
    out->n_ref = hdr.n_ref; out->ref_len = hdr.ref_len; out->ref_names = hdr.ref_names; out->ref_names_len = hdr.ref_names_len;
    out->t_inflate_s = hdr.t_inflate_s; out->t_index_s = now_s() - t0 - hdr.t_inflate_s;
    return 0;
}

void tc_bampayload_free(tc_bampayload_t* p) {
    if (!p) return;
    free(p->payload); free(p->rec_off); free(p->ref_len); free(p->ref_names);
    memset(p, 0, sizeof(*p));
}

/* The file as it is, plus the index of its BGZF members: what tc_bgzf_inflate (trueconsense_b200.h) needs to inflate the
 * members on the GPU — the host touches one header per member and nothing else. */
int tc_bgzf_map(const char* path, tc_bgzf_map_t* out, char* err, int errlen) {
    memset(out, 0, sizeof(*out));
    int fd = open(path, O_RDONLY);
    if (fd < 0) return fail(err, errlen, -1, "cannot open %s", path);
    struct stat st;
    if (fstat(fd, &st) != 0) { close(fd); return fail(err, errlen, -1, "cannot stat %s", path); }
    int64_t fsize = st.st_size;
    if (fsize < 28) { close(fd); return fail(err, errlen, -2, "%s: too short for a BAM file", path); }
    const uint8_t* f = mmap(NULL, (size_t)fsize, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (f == MAP_FAILED) return fail(err, errlen, -1, "mmap failed for %s", path);
    int rc = bgzf_index_blocks(f, fsize, &out->blocks, &out->n_blocks, &out->payload_bytes, err, errlen);
    if (rc) { munmap((void*)f, fsize); return rc; }
    out->file = f; out->file_bytes = fsize;
    return 0;
}

void tc_bgzf_unmap(tc_bgzf_map_t* m) {
    if (!m) return;
    if (m->file) munmap((void*)m->file, (size_t)m->file_bytes);
    free(m->blocks);
    memset(m, 0, sizeof(*m));
}

/* ------------------------------------------------------------------ writer */

typedef struct { FILE* fp; uint8_t buf[0xff00]; int n; int level; int err; } bgzf_w;

static void bgzf_flush(bgzf_w* w) {
    if (w->err) return;
    uint8_t outb[0x10000 + 64];
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    if (deflateInit2(&zs, w->level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) { w->err = 1; return; }
    zs.next_in = w->buf; zs.avail_in = w->n;
    zs.next_out = outb + 18; zs.avail_out = sizeof(outb) - 18 - 8;
    if (deflate(&zs, Z_FINISH) != Z_STREAM_END) { w->err = 1; deflateEnd(&zs); return; }
    int clen = (int)zs.total_out;
    deflateEnd(&zs);
    int bsize = 18 + clen + 8;
    if (bsize > 0x10000) { w->err = 1; return; }
    static const uint8_t hdr[16] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 'B', 'C', 2, 0};
    memcpy(outb, hdr, 16);
    wr16(outb + 16, (uint16_t)(bsize - 1));
    wr32(outb + 18 + clen, (uint32_t)crc32(crc32(0L, Z_NULL, 0), w->buf, w->n));
    wr32(outb + 18 + clen + 4, (uint32_t)w->n);
    if (fwrite(outb, 1, bsize, w->fp) != (size_t)bsize) w->err = 1;
    w->n = 0;
}

static void bgzf_write(bgzf_w* w, const void* data, size_t len) {
    const uint8_t* d = data;
    while (len) {
        size_t room = sizeof(w->buf) - w->n;
        size_t k = len < room ? len : room;
        memcpy(w->buf + w->n, d, k);
        w->n += (int)k; d += k; len -= k;
        if (w->n == (int)sizeof(w->buf)) bgzf_flush(w);
    }
}

static int reg2bin(int64_t beg, int64_t end) {
    --end;
    if (beg >> 14 == end >> 14) return (int)(((1 << 15) - 1) / 7 + (beg >> 14));
    if (beg >> 17 == end >> 17) return (int)(((1 << 12) - 1) / 7 + (beg >> 17));
    if (beg >> 20 == end >> 20) return (int)(((1 << 9) - 1) / 7 + (beg >> 20));
    if (beg >> 23 == end >> 23) return (int)(((1 << 6) - 1) / 7 + (beg >> 23));
    if (beg >> 26 == end >> 26) return (int)(((1 << 3) - 1) / 7 + (beg >> 26));
    return 0;
}

int tc_bam_write(const char* path, const tc_hostreads_t* rd, const char* ref_name, int32_t ref_len,
                 int level, char* err, int errlen) {
    bgzf_w* w = calloc(1, sizeof(bgzf_w));
    if (!w) return fail(err, errlen, -5, "out of memory");
    w->fp = fopen(path, "wb");
    if (!w->fp) { free(w); return fail(err, errlen, -1, "cannot create %s", path); }
    w->level = level;
    char text[512];
    int l_text = snprintf(text, sizeof(text), "@HD\tVN:1.6\tSO:coordinate\n@SQ\tSN:%s\tLN:%d\n", ref_name, ref_len);
    uint8_t b4[4];
    bgzf_write(w, "BAM\1", 4);
    wr32(b4, (uint32_t)l_text); bgzf_write(w, b4, 4);
    bgzf_write(w, text, l_text);
    wr32(b4, 1); bgzf_write(w, b4, 4);
    int l_name = (int)strlen(ref_name) + 1;
    wr32(b4, (uint32_t)l_name); bgzf_write(w, b4, 4);
    bgzf_write(w, ref_name, l_name);
    wr32(b4, (uint32_t)ref_len); bgzf_write(w, b4, 4);
    for (int64_t i = 0; i < rd->n_reads; ++i) {
        char name[40];
        int ln = snprintf(name, sizeof(name), "q%016llx", (unsigned long long)rd->qname_hash[i]) + 1;
        uint32_t n_cig = rd->cigar_off[i + 1] - rd->cigar_off[i];
        uint32_t l_seq = (uint32_t)rd->l_seq[i];
        const uint32_t* cg = rd->cigar + rd->cigar_off[i];
        int64_t span = 0;
        for (uint32_t k = 0; k < n_cig; ++k) {
            uint32_t op = cg[k] & 15;
            if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) span += cg[k] >> 4;
        }
        int64_t end = rd->pos[i] + (span > 0 ? span : 1);
        uint8_t core[36];
        int32_t bs = 32 + ln + 4 * (int32_t)n_cig + (int32_t)((l_seq + 1) / 2) + (int32_t)l_seq;
        wr32(core, (uint32_t)bs);
        wr32(core + 4, (uint32_t)(rd->tid ? rd->tid[i] : 0));
        wr32(core + 8, (uint32_t)rd->pos[i]);
        core[12] = (uint8_t)ln; core[13] = rd->mapq[i];
        wr16(core + 14, (uint16_t)reg2bin(rd->pos[i], end));
        wr16(core + 16, (uint16_t)n_cig);
        wr16(core + 18, rd->flag[i]);
        wr32(core + 20, l_seq);
        int32_t mtid = rd->mtid ? rd->mtid[i] : ((rd->flag[i] & 1) ? 0 : -1);
        wr32(core + 24, (uint32_t)mtid);
        wr32(core + 28, (uint32_t)(rd->mpos ? rd->mpos[i] : -1));
        wr32(core + 32, (uint32_t)(rd->isize ? rd->isize[i] : 0));
        bgzf_write(w, core, 36);
        bgzf_write(w, name, ln);
        bgzf_write(w, cg, 4 * (size_t)n_cig);
        bgzf_write(w, (const uint8_t*)(rd->seq4 + rd->seq_off[i]), (l_seq + 1) / 2);
        bgzf_write(w, rd->qual + 8ULL * rd->seq_off[i], l_seq);
    }
    if (w->n) bgzf_flush(w);
    static const uint8_t eof_blk[28] = {0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0, 0xff, 0x06, 0, 0x42, 0x43, 0x02, 0,
                                        0x1b, 0, 0x03, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (fwrite(eof_blk, 1, 28, w->fp) != 28) w->err = 1;
    int e = w->err;
    if (fclose(w->fp) != 0) e = 1;
    free(w);
    if (e) return fail(err, errlen, -1, "write error on %s", path);
    return 0;
}
