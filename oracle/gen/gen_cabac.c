/* gen_cabac.c -- TEST INFRASTRUCTURE ONLY (bitstream fabrication; never linked into the product).  See gen_cabac.h.
 *
 * Bin chooser + CABAC ENCODER behind the reference parser.  The encoder is the arithmetic coder of H.266 clause 9.3.4.4
 * (the counterpart of the decoder in libovvc/cabac_internal.h:45-125): 9-bit range, LPS range from the same
 * range_lps_lut the decoder indexes, renormalisation by lps_renorm_table, carry propagation through buffered 0xff bytes,
 * terminating bin + flush at the end of the slice.  The context state update is the decoder's (two windowed estimators
 * packed in one uint64, cabac_internal.h:76-84), because the parser's later context choices do not depend on it but the
 * decoder's later probabilities do.
 *
 * One slice is in flight at a time (run the decoder with one frame thread and one entry thread).
 *   OVGEN_OUT    file that receives, per slice, [uint32 length][payload bytes] (payload = slice_data() + rbsp trailing bits)
 *   OVGEN_SEED   seed of the chooser
 *   OVGEN_PROBS  "first-last:p,..." probability (0..255, of 256) of bin 1 for context indices first..last (default 128)
 *   OVGEN_MAX_ONES  longest run of 1s among bypass bins (bounds Exp-Golomb / Rice prefixes, i.e. MVDs and levels)
 *   OVGEN_BINLOG optional text log: one line per bin "ctx bin" (ctx -1 = bypass) */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include "vcl_cabac.h"

extern const uint8_t lps_renorm_table[64];
extern const uint8_t range_lps_lut[512];

static struct {
    int active;
    uint32_t low, range;
    int bits_left, num_buffered;
    uint32_t buffered_byte;
    uint8_t *buf;
    size_t n, cap;
    uint64_t rng;
    int ones_run, max_ones;
    uint8_t p1[OVCABAC_NB_CTX];
    int configured;
    FILE *out, *binlog;
    unsigned long nb_bins, nb_slices;
} g;

static void put_byte(uint32_t b)
{
    if (g.n == g.cap) {
        g.cap = g.cap ? 2 * g.cap : (1u << 20);
        g.buf = (uint8_t *)realloc(g.buf, g.cap);
        if (!g.buf) { fprintf(stderr, "ovgen: out of memory\n"); exit(1); }
    }
    g.buf[g.n++] = (uint8_t)b;
}

/* xorshift64*: the chooser's only source of randomness */
static uint32_t rnd(void)
{
    g.rng ^= g.rng >> 12; g.rng ^= g.rng << 25; g.rng ^= g.rng >> 27;
    return (uint32_t)((g.rng * 0x2545F4914F6CDD1Dull) >> 32);
}

static void configure(void)
{
    const char *s;
    g.configured = 1;
    memset(g.p1, 128, sizeof(g.p1));
    g.rng = 0x9E3779B97F4A7C15ull ^ (uint64_t)((s = getenv("OVGEN_SEED")) ? strtoull(s, NULL, 0) : 1);
    g.max_ones = (s = getenv("OVGEN_MAX_ONES")) ? atoi(s) : 5;
    if ((s = getenv("OVGEN_PROBS"))) {
        while (*s) {
            int a, b, p, used = 0;
            if (sscanf(s, "%d-%d:%d%n", &a, &b, &p, &used) == 3 && used > 0) {
                for (int i = a; i <= b && i < OVCABAC_NB_CTX; i++) if (i >= 0) g.p1[i] = (uint8_t)(p < 0 ? 0 : p > 255 ? 255 : p);
                s += used;
            } else break;
            if (*s == ',') s++;
        }
    }
    if ((s = getenv("OVGEN_OUT"))) g.out = fopen(s, "wb");
    if (!g.out) { fprintf(stderr, "ovgen: OVGEN_OUT is not writable\n"); exit(1); }
    if ((s = getenv("OVGEN_BINLOG"))) g.binlog = fopen(s, "w");
}

static void write_out(void)
{
    const uint32_t lead = g.low >> (24 - g.bits_left);
    g.bits_left += 8;
    g.low &= 0xffffffffu >> g.bits_left;
    if (lead == 0xff) {
        g.num_buffered++;
    } else if (g.num_buffered > 0) {
        const uint32_t carry = lead >> 8;
        put_byte(g.buffered_byte + carry);
        g.buffered_byte = lead & 0xff;
        for (const uint32_t fill = (0xff + carry) & 0xff; g.num_buffered > 1; g.num_buffered--) put_byte(fill);
    } else {
        g.num_buffered = 1;
        g.buffered_byte = lead;
    }
}

static void enc_start(void)
{
    g.low = 0; g.range = 510; g.bits_left = 23; g.num_buffered = 0; g.buffered_byte = 0xff;
    g.n = 0; g.ones_run = 0; g.active = 1;
}

static void enc_terminate_and_flush(void)
{
    /* end_of_slice_one_bit = 1 (9.3.4.4.5), then the flush of 9.3.4.4.6 */
    g.range -= 2;
    g.low += g.range;
    g.low <<= 7; g.range = 2 << 7; g.bits_left -= 7;
    if (g.bits_left < 12) write_out();
    if (g.low >> (32 - g.bits_left)) {
        put_byte(g.buffered_byte + 1);
        for (; g.num_buffered > 1; g.num_buffered--) put_byte(0x00);
        g.low -= 1u << (32 - g.bits_left);
    } else {
        if (g.num_buffered > 0) put_byte(g.buffered_byte);
        for (; g.num_buffered > 1; g.num_buffered--) put_byte(0xff);
    }
    /* the remaining 24 - bits_left bits of low >> 8, then rbsp_slice_trailing_bits(): a 1 and zeros up to the byte boundary */
    {
        int nb = 24 - g.bits_left;
        uint64_t v = ((uint64_t)(g.low >> 8) & ((1ull << nb) - 1)) << 1 | 1;
        nb += 1;
        while (nb & 7) { v <<= 1; nb++; }
        for (int s = nb - 8; s >= 0; s -= 8) put_byte((uint32_t)(v >> s) & 0xff);
    }
}

static void finish_slice(void)
{
    if (!g.active) return;
    enc_terminate_and_flush();
    uint32_t len = (uint32_t)g.n;
    fwrite(&len, 4, 1, g.out);
    fwrite(g.buf, 1, g.n, g.out);
    fflush(g.out);
    g.active = 0;
    g.nb_slices++;
}

__attribute__((destructor)) static void ovgen_at_exit(void)
{
    if (!g.configured) return;
    finish_slice();
    if (g.out) fclose(g.out);
    if (g.binlog) fclose(g.binlog);
    fprintf(stderr, "ovgen: %lu slices, %lu bins\n", g.nb_slices, g.nb_bins);
}

/* Takes the place of ovcabac_attach_entry (libovvc/vcl_cabac.c:940; slicedec.c is compiled with the name redirected): a new
 * entry starts, so the previous slice is complete. */
int ovgen_attach_entry(OVCABACCtx *const c, const uint8_t *const entry_point, const uint8_t *const entry_end)
{
    if (!g.configured) configure();
    finish_slice();
    enc_start();
    c->bytestream_start = c->bytestream = entry_point;
    c->bytestream_end = entry_end;
    c->range = 0x1FE;
    c->low_b = 0;
    return 0;
}

uint8_t ovgen_ae_read(OVCABACCtx *c, uint64_t *cabac_state)
{
    const int ctx = (int)(cabac_state - c->ctx_table);
    uint16_t s0 = (uint16_t)(*cabac_state >> 48), s1 = (uint16_t)(*cabac_state >> 32);
    const uint8_t r0 = (uint8_t)(*cabac_state >> 16), r1 = (uint8_t)(*cabac_state);
    int8_t st = (int8_t)((s0 + s1) >> 8);
    const int mps = st < 0;                              /* sign bit of the summed estimators = most probable symbol */
    const int q = (st < 0 ? ~st : st) >> 2;             /* 5-bit probability class of the LPS */
    const int bin = (int)(rnd() & 255) < g.p1[ctx >= 0 && ctx < OVCABAC_NB_CTX ? ctx : 0];
    const uint32_t mask = bin ? 0xffffu : 0;

    /* encode */
    const uint32_t lps = range_lps_lut[(g.range & 0x1E0) | q];
    g.range -= lps;
    if (bin != mps) { g.low += g.range; g.range = lps; }
    {
        const int n = lps_renorm_table[g.range >> 3];
        g.low <<= n; g.range <<= n; g.bits_left -= n;
        if (g.bits_left < 12) write_out();
    }
    /* the decoder's state update with the chosen bin */
    s0 -= (s0 >> r0) & 0x7FE0;
    s1 -= (s1 >> r1) & 0x7FFE;
    s0 += (0x7fffu >> r0) & 0x7FE0 & mask;
    s1 += (0x7fffu >> r1) & 0x7FFE & mask;
    *cabac_state = (*cabac_state & 0xFFFFFFFFull) | ((uint64_t)s0 << 48) | ((uint64_t)s1 << 32);
    g.nb_bins++;
    g.ones_run = 0;
    if (g.binlog) fprintf(g.binlog, "%d %d\n", ctx, bin);
    return (uint8_t)bin;
}

uint8_t ovgen_bypass_read(OVCABACCtx *c)
{
    (void)c;
    int bin = (int)(rnd() & 1);
    if (bin && g.ones_run >= g.max_ones) bin = 0;
    g.ones_run = bin ? g.ones_run + 1 : 0;
    g.low <<= 1;
    if (bin) g.low += g.range;
    g.bits_left--;
    if (g.bits_left < 12) write_out();
    g.nb_bins++;
    if (g.binlog) fprintf(g.binlog, "-1 %d\n", bin);
    return (uint8_t)bin;
}
