/*
 * zstd_oracle.h — Zstandard frame decoder of the CPU ORACLE (TEST INFRASTRUCTURE ONLY; included by horae_oracle.c).
 *
 * The reference decodes ParquetCompression::Zstd pages (config.rs:78-94) through parquet 53.2 -> zstd 0.13.2 (libzstd, Cargo.lock:3893),
 * a dependency that is not under /root/reference.  This is a plain, strictly sequential restatement of the published format (RFC 8878:
 * frames 3.1.1, blocks 3.1.1.2, literals 3.1.1.3.1, sequences 3.1.1.3.2, FSE 4.1, Huffman 4.2), written independently of the device
 * decoder (horaedb_b200/csrc/zstd_core.h): one thread, whole tables built with malloc'ed scratch, byte-by-byte match copies.
 * Pinned in tests/test_oracle_decode.py against libzstd itself (pyarrow's codec) on frames of levels 1, 3 and 9.
 */
#ifndef ZSTD_ORACLE_H
#define ZSTD_ORACLE_H

#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ZO_FAIL(msg) do { snprintf(g_err, sizeof g_err, "zstd: %s", msg); return -1; } while (0)

typedef struct { uint8_t sym, nbits; uint16_t base; } zo_cell;           /* one FSE decoding-table cell */
typedef struct { zo_cell cell[512]; int log; int valid; } zo_fse;        /* accuracy log <= 9 */
typedef struct { uint8_t sym[2048], nbits[2048]; int log; int valid; } zo_huf;

static int zo_highbit(uint32_t v) { int r = -1; while (v) { v >>= 1; r++; } return r; }

/* ---- bit readers ------------------------------------------------------------------------------------------------------------- */
/* forward, LSB first (FSE table descriptions) */
typedef struct { const uint8_t *p; int64_t nbytes; int64_t bit; } zo_fwd;
static uint32_t zo_fwd_bits(zo_fwd *b, int n) {
    uint32_t v = 0;
    for (int i = 0; i < n; i++, b->bit++) {
        int64_t byte = b->bit >> 3;
        uint32_t x = byte < b->nbytes ? (b->p[byte] >> (b->bit & 7)) & 1u : 0u;
        v |= x << i;
    }
    return v;
}
/* backward (entropy-coded streams): `bit` = number of bits not yet consumed; bits below the start of the stream read as 0 */
typedef struct { const uint8_t *p; int64_t bit; } zo_bwd;
static int zo_bwd_init(zo_bwd *b, const uint8_t *p, int64_t n) {
    if (n <= 0 || p[n - 1] == 0) return -1;
    b->p = p;
    b->bit = n * 8 - (8 - zo_highbit(p[n - 1]));
    return 0;
}
static uint32_t zo_bwd_bits(zo_bwd *b, int n) {
    uint32_t v = 0;
    b->bit -= n;
    for (int i = 0; i < n; i++) {
        int64_t at = b->bit + i;
        uint32_t x = at >= 0 ? (b->p[at >> 3] >> (at & 7)) & 1u : 0u;
        v |= x << i;
    }
    return v;
}

/* ---- FSE (RFC 8878 4.1) -------------------------------------------------------------------------------------------------------- */
static int zo_fse_from_counts(zo_fse *t, const int16_t *norm, int nsym, int log) {
    int size = 1 << log, high = size, pos = 0;
    uint8_t spread[512];
    uint16_t next[256];
    if (log > 9 || nsym > 256) return -1;
    for (int s = 0; s < nsym; s++) if (norm[s] == -1) { spread[--high] = (uint8_t)s; next[s] = 1; }
    int step = (size >> 1) + (size >> 3) + 3;
    for (int s = 0; s < nsym; s++) {
        if (norm[s] <= 0) continue;
        next[s] = (uint16_t)norm[s];
        for (int i = 0; i < norm[s]; i++) {
            spread[pos] = (uint8_t)s;
            do { pos = (pos + step) & (size - 1); } while (pos >= high);
        }
    }
    if (pos != 0) return -1;
    for (int i = 0; i < size; i++) {
        int s = spread[i];
        uint32_t x = next[s]++;
        int nb = log - zo_highbit(x);
        t->cell[i].sym = (uint8_t)s;
        t->cell[i].nbits = (uint8_t)nb;
        t->cell[i].base = (uint16_t)((x << nb) - (uint32_t)size);
    }
    t->log = log;
    t->valid = 1;
    return 0;
}
/* table description -> normalised counts; returns bytes consumed or -1 */
static int64_t zo_fse_read_counts(const uint8_t *p, int64_t n, int max_log, int max_sym, int16_t *norm, int *nsym, int *log) {
    zo_fwd b = {p, n, 0};
    int al = (int)zo_fwd_bits(&b, 4) + 5;
    if (al > max_log) return -1;
    int remaining = 1 << al, s = 0;
    while (remaining > 0 && s < max_sym) {
        int bits = zo_highbit((uint32_t)(remaining + 1)) + 1;
        uint32_t v = zo_fwd_bits(&b, bits);
        uint32_t lower = (1u << (bits - 1)) - 1u, thr = (1u << bits) - 1u - (uint32_t)(remaining + 1);
        if ((v & lower) < thr) { b.bit--; v &= lower; }
        else if (v > lower) v -= thr;
        int proba = (int)v - 1;
        remaining -= proba < 0 ? -proba : proba;
        norm[s++] = (int16_t)proba;
        if (proba == 0) {
            uint32_t rep = zo_fwd_bits(&b, 2);
            for (;;) {
                for (uint32_t i = 0; i < rep && s < max_sym; i++) norm[s++] = 0;
                if (rep == 3) rep = zo_fwd_bits(&b, 2); else break;
            }
        }
        if ((b.bit + 7) / 8 > n) return -1;
    }
    if (remaining != 0) return -1;
    *nsym = s; *log = al;
    return (b.bit + 7) / 8;
}

/* ---- Huffman (RFC 8878 4.2) ---------------------------------------------------------------------------------------------------- */
static int zo_huf_from_weights(zo_huf *h, uint8_t *w, int nw) {
    uint32_t sum = 0;
    for (int i = 0; i < nw; i++) { if (w[i] > 11) return -1; if (w[i]) sum += 1u << (w[i] - 1); }
    if (!sum) return -1;
    int maxb = zo_highbit(sum) + 1;
    uint32_t left = (1u << maxb) - sum;
    if (maxb > 11 || (left & (left - 1))) return -1;
    w[nw] = (uint8_t)(zo_highbit(left) + 1);
    int n = nw + 1;
    /* canonical order: by code length descending (weight ascending), symbol ascending inside a length; longest codes lowest */
    uint32_t at = 0;
    for (int wt = 1; wt <= maxb; wt++) {
        int len = maxb + 1 - wt;
        for (int s = 0; s < n; s++) {
            if (w[s] != wt) continue;
            uint32_t span = 1u << (maxb - len);
            for (uint32_t j = 0; j < span; j++) { h->sym[at + j] = (uint8_t)s; h->nbits[at + j] = (uint8_t)len; }
            at += span;
        }
    }
    if (at != (1u << maxb)) return -1;
    h->log = maxb;
    h->valid = 1;
    return 0;
}
static int64_t zo_huf_read_tree(zo_huf *h, const uint8_t *p, int64_t n) {
    if (n < 1) return -1;
    uint8_t w[256];
    int nw = 0, hb = p[0];
    int64_t used;
    if (hb >= 128) {
        nw = hb - 127;
        if (1 + (nw + 1) / 2 > n) return -1;
        for (int i = 0; i < nw; i++) w[i] = (i & 1) ? (p[1 + i / 2] & 15) : (p[1 + i / 2] >> 4);
        used = 1 + (nw + 1) / 2;
    } else {
        if (hb == 0 || 1 + hb > n) return -1;
        int16_t norm[64]; int nsym, log;
        int64_t hdr = zo_fse_read_counts(p + 1, hb, 7, 64, norm, &nsym, &log);
        if (hdr < 0 || hdr > hb) return -1;
        zo_fse *t = malloc(sizeof(zo_fse));
        if (zo_fse_from_counts(t, norm, nsym, log)) { free(t); return -1; }
        zo_bwd b;
        if (zo_bwd_init(&b, p + 1 + hdr, hb - hdr)) { free(t); return -1; }
        uint32_t s1 = zo_bwd_bits(&b, log), s2 = zo_bwd_bits(&b, log);
        for (;;) {                                            /* two interleaved states until the stream runs dry */
            if (nw > 253) { free(t); return -1; }
            w[nw++] = t->cell[s1].sym;
            s1 = t->cell[s1].base + zo_bwd_bits(&b, t->cell[s1].nbits);
            if (b.bit < 0) { w[nw++] = t->cell[s2].sym; break; }
            w[nw++] = t->cell[s2].sym;
            s2 = t->cell[s2].base + zo_bwd_bits(&b, t->cell[s2].nbits);
            if (b.bit < 0) { w[nw++] = t->cell[s1].sym; break; }
        }
        free(t);
        used = 1 + hb;
    }
    if (nw < 1 || nw > 255 || zo_huf_from_weights(h, w, nw)) return -1;
    return used;
}
static int zo_huf_stream(const zo_huf *h, const uint8_t *p, int64_t n, uint8_t *out, int64_t count) {
    zo_bwd b;
    if (zo_bwd_init(&b, p, n)) return -1;
    uint32_t mask = (1u << h->log) - 1u, st = zo_bwd_bits(&b, h->log);
    for (int64_t i = 0; i < count; i++) {
        out[i] = h->sym[st];
        int nb = h->nbits[st];
        st = ((st << nb) + zo_bwd_bits(&b, nb)) & mask;
    }
    return b.bit == -(int64_t)h->log ? 0 : -1;
}

/* ---- sequence code tables (RFC 8878 3.1.1.3.2.1) -------------------------------------------------------------------------------- */
static const uint32_t zo_ll_base[36] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 18, 20, 22, 24, 28, 32, 40, 48, 64, 128, 256, 512,
                                        1024, 2048, 4096, 8192, 16384, 32768, 65536};
static const uint8_t zo_ll_bits[36] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 3, 3, 4, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};
static const uint32_t zo_ml_base[53] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31, 32,
                                        33, 34, 35, 37, 39, 41, 43, 47, 51, 59, 67, 83, 99, 131, 259, 515, 1027, 2051, 4099, 8195, 16387, 32771, 65539};
static const uint8_t zo_ml_bits[53] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
                                       1, 1, 1, 1, 2, 2, 3, 3, 4, 4, 5, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};
static const int16_t zo_ll_default[36] = {4, 3, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 3, 2, 1, 1, 1, 1, 1, -1, -1, -1, -1};
static const int16_t zo_ml_default[53] = {1, 4, 3, 2, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1,
                                          1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1, -1, -1};
static const int16_t zo_of_default[29] = {1, 1, 1, 1, 1, 1, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1};

/* one sequence table per block: mode 0 predefined, 1 RLE, 2 described, 3 repeat */
static int64_t zo_seq_table(zo_fse *t, int mode, const uint8_t *p, int64_t n, const int16_t *def, int def_n, int def_log, int max_log, int max_sym) {
    if (mode == 0) return zo_fse_from_counts(t, def, def_n, def_log) ? -1 : 0;
    if (mode == 1) {
        if (n < 1 || p[0] >= max_sym) return -1;
        t->cell[0].sym = p[0]; t->cell[0].nbits = 0; t->cell[0].base = 0; t->log = 0; t->valid = 1;
        return 1;
    }
    if (mode == 2) {
        int16_t norm[64]; int nsym, log;
        int64_t used = zo_fse_read_counts(p, n, max_log, max_sym, norm, &nsym, &log);
        if (used < 0 || zo_fse_from_counts(t, norm, nsym, log)) return -1;
        return used;
    }
    return t->valid ? 0 : -1;
}

/* ---- frames -------------------------------------------------------------------------------------------------------------------- */
static int zstd_decompress(const uint8_t *src, int64_t n, uint8_t *dst, int64_t cap) {
    const uint8_t *p = src, *end = src + n;
    int64_t o = 0;
    zo_fse *ll = calloc(1, sizeof(zo_fse)), *of = calloc(1, sizeof(zo_fse)), *ml = calloc(1, sizeof(zo_fse));
    zo_huf *huf = calloc(1, sizeof(zo_huf));
    uint8_t *lit = malloc((128u << 10) + 64);
    int rc = -1;
#define ZO_OUT(msg) do { snprintf(g_err, sizeof g_err, "zstd: %s", msg); goto out; } while (0)
    while (p < end) {
        if (end - p < 5) ZO_OUT("truncated frame header");
        uint32_t magic = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
        if ((magic & 0xfffffff0u) == 0x184d2a50u) {            /* skippable frame */
            if (end - p < 8) ZO_OUT("truncated skippable frame");
            uint32_t sz = (uint32_t)p[4] | ((uint32_t)p[5] << 8) | ((uint32_t)p[6] << 16) | ((uint32_t)p[7] << 24);
            if ((int64_t)sz + 8 > end - p) ZO_OUT("truncated skippable frame");
            p += 8 + sz;
            continue;
        }
        if (magic != 0xfd2fb528u) ZO_OUT("bad magic");
        int fhd = p[4];
        p += 5;
        int fcs = fhd >> 6, single = (fhd >> 5) & 1, checksum = (fhd >> 2) & 1, did = fhd & 3;
        if (fhd & 8) ZO_OUT("reserved bit set");
        if (did) ZO_OUT("dictionary frames are not used by Parquet pages");
        if (!single) p += 1;
        p += fcs == 0 ? (single ? 1 : 0) : (fcs == 1 ? 2 : (fcs == 2 ? 4 : 8));
        if (p > end) ZO_OUT("truncated frame header");
        uint32_t rep[3] = {1, 4, 8};
        ll->valid = of->valid = ml->valid = huf->valid = 0;
        for (;;) {
            if (end - p < 3) ZO_OUT("truncated block header");
            uint32_t bh = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16);
            p += 3;
            int last = bh & 1, type = (bh >> 1) & 3;
            int64_t bsize = bh >> 3;
            if (type == 0) {
                if (bsize > end - p || o + bsize > cap) ZO_OUT("raw block overrun");
                memcpy(dst + o, p, (size_t)bsize); p += bsize; o += bsize;
            } else if (type == 1) {
                if (end - p < 1 || o + bsize > cap) ZO_OUT("RLE block overrun");
                memset(dst + o, p[0], (size_t)bsize); p += 1; o += bsize;
            } else if (type == 2) {
                if (bsize > end - p || bsize > (128 << 10) || bsize < 2) ZO_OUT("bad compressed block size");
                const uint8_t *bend = p + bsize;
                /* literals */
                int ltype = p[0] & 3, sf = (p[0] >> 2) & 3, nstreams = 1;
                int64_t regen, csize = 0, hdr;
                uint64_t h5 = 0;
                for (int i = 0; i < 5 && p + i < bend; i++) h5 |= (uint64_t)p[i] << (8 * i);
                if (ltype < 2) {
                    if ((sf & 1) == 0) { regen = p[0] >> 3; hdr = 1; }
                    else if (sf == 1) { regen = (h5 & 0xffff) >> 4; hdr = 2; }
                    else { regen = (h5 & 0xffffff) >> 4; hdr = 3; }
                } else if (sf < 2) { regen = (h5 >> 4) & 0x3ff; csize = (h5 >> 14) & 0x3ff; hdr = 3; nstreams = sf == 0 ? 1 : 4; }
                else if (sf == 2) { regen = (h5 >> 4) & 0x3fff; csize = (h5 >> 18) & 0x3fff; hdr = 4; nstreams = 4; }
                else { regen = (h5 >> 4) & 0x3ffff; csize = (h5 >> 22) & 0x3ffff; hdr = 5; nstreams = 4; }
                if (regen > (128 << 10)) ZO_OUT("literals larger than a block");
                const uint8_t *lp = p + hdr;
                const uint8_t *lits = lit;
                if (ltype == 0) { if (regen > bend - lp) ZO_OUT("raw literals overrun"); lits = lp; p = lp + regen; }
                else if (ltype == 1) { if (bend - lp < 1) ZO_OUT("RLE literals overrun"); memset(lit, lp[0], (size_t)regen); p = lp + 1; }
                else {
                    if (csize > bend - lp) ZO_OUT("compressed literals overrun");
                    const uint8_t *hp = lp;
                    int64_t tree = 0;
                    if (ltype == 2) { tree = zo_huf_read_tree(huf, hp, csize); if (tree < 0) ZO_OUT("bad Huffman tree"); hp += tree; }
                    else if (!huf->valid) ZO_OUT("treeless literals without a tree");
                    int64_t total = csize - tree;
                    if (nstreams == 1) { if (zo_huf_stream(huf, hp, total, lit, regen)) ZO_OUT("bad Huffman stream"); }
                    else {
                        if (total < 6) ZO_OUT("bad jump table");
                        int64_t s1 = hp[0] | (hp[1] << 8), s2 = hp[2] | (hp[3] << 8), s3 = hp[4] | (hp[5] << 8);
                        int64_t s4 = total - 6 - s1 - s2 - s3, per = (regen + 3) / 4;
                        if (s4 < 0 || 3 * per > regen) ZO_OUT("bad jump table");
                        const uint8_t *q = hp + 6;
                        if (zo_huf_stream(huf, q, s1, lit, per) || zo_huf_stream(huf, q + s1, s2, lit + per, per) ||
                            zo_huf_stream(huf, q + s1 + s2, s3, lit + 2 * per, per) || zo_huf_stream(huf, q + s1 + s2 + s3, s4, lit + 3 * per, regen - 3 * per))
                            ZO_OUT("bad Huffman stream");
                    }
                    p = lp + csize;
                }
                /* sequences */
                if (p >= bend) ZO_OUT("missing sequences section");
                int64_t nseq = p[0], lpos = 0;
                if (nseq < 128) p += 1;
                else if (nseq < 255) { if (bend - p < 2) ZO_OUT("truncated sequence count"); nseq = ((nseq - 128) << 8) + p[1]; p += 2; }
                else { if (bend - p < 3) ZO_OUT("truncated sequence count"); nseq = p[1] + (p[2] << 8) + 0x7f00; p += 3; }
                if (nseq) {
                    if (p >= bend) ZO_OUT("missing compression modes");
                    int modes = *p++;
                    if (modes & 3) ZO_OUT("reserved mode bits");
                    int64_t u;
                    if ((u = zo_seq_table(ll, modes >> 6, p, bend - p, zo_ll_default, 36, 6, 9, 36)) < 0) ZO_OUT("bad literal-length table");
                    p += u;
                    if ((u = zo_seq_table(of, (modes >> 4) & 3, p, bend - p, zo_of_default, 29, 5, 8, 32)) < 0) ZO_OUT("bad offset table");
                    p += u;
                    if ((u = zo_seq_table(ml, (modes >> 2) & 3, p, bend - p, zo_ml_default, 53, 6, 9, 53)) < 0) ZO_OUT("bad match-length table");
                    p += u;
                    zo_bwd b;
                    if (p >= bend || zo_bwd_init(&b, p, bend - p)) ZO_OUT("bad sequence bitstream");
                    uint32_t sl = zo_bwd_bits(&b, ll->log), so = zo_bwd_bits(&b, of->log), sm = zo_bwd_bits(&b, ml->log);
                    for (int64_t i = 0; i < nseq; i++) {
                        int oc = of->cell[so].sym, mc = ml->cell[sm].sym, lc = ll->cell[sl].sym;
                        if (oc > 31 || mc > 52 || lc > 35) ZO_OUT("bad sequence code");
                        uint32_t ov = (1u << oc) + zo_bwd_bits(&b, oc);
                        uint32_t mlen = zo_ml_base[mc] + zo_bwd_bits(&b, zo_ml_bits[mc]);
                        uint32_t llen = zo_ll_base[lc] + zo_bwd_bits(&b, zo_ll_bits[lc]);
                        if (i + 1 < nseq) {
                            sl = ll->cell[sl].base + zo_bwd_bits(&b, ll->cell[sl].nbits);
                            sm = ml->cell[sm].base + zo_bwd_bits(&b, ml->cell[sm].nbits);
                            so = of->cell[so].base + zo_bwd_bits(&b, of->cell[so].nbits);
                        }
                        if (b.bit < 0) ZO_OUT("sequence bitstream overrun");
                        uint32_t off;
                        if (ov > 3) { off = ov - 3; rep[2] = rep[1]; rep[1] = rep[0]; rep[0] = off; }
                        else {
                            uint32_t idx = ov - 1 + (llen == 0);
                            if (idx == 0) off = rep[0];
                            else {
                                off = idx < 3 ? rep[idx] : rep[0] - 1;
                                if (idx > 1) rep[2] = rep[1];
                                rep[1] = rep[0];
                                rep[0] = off;
                            }
                        }
                        if (lpos + llen > regen || o + llen + mlen > cap || off == 0 || off > o + llen) ZO_OUT("bad sequence");
                        memcpy(dst + o, lits + lpos, llen); lpos += llen; o += llen;
                        for (uint32_t j = 0; j < mlen; j++) dst[o + j] = dst[o + j - off];
                        o += mlen;
                    }
                    if (b.bit != 0) ZO_OUT("sequence bitstream not consumed");
                }
                if (o + (regen - lpos) > cap) ZO_OUT("trailing literals overrun");
                memcpy(dst + o, lits + lpos, (size_t)(regen - lpos));
                o += regen - lpos;
                p = bend;
            } else ZO_OUT("reserved block type");
            if (last) break;
        }
        if (checksum) p += 4;
    }
    if (o != cap) ZO_OUT("short output");
    rc = 0;
out:
#undef ZO_OUT
    free(ll); free(of); free(ml); free(huf); free(lit);
    return rc;
}

#endif
