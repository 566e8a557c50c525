/* CPU restatement of torchac 0.9.3's range coder -- TEST INFRASTRUCTURE.
 *
 * torchac is a third-party dependency of the reference (README.md:20; call
 * sites entropy_model.py:174,192); its source is NOT under /root/reference.
 * PARITY UNPINNED against real torchac bytes.  The algorithm restated here is
 * the published one: 32-bit low/high interval,
 * 16-bit CDF precision, pending-bit carry resolution, MSB-first bit packing.
 *
 * The CDF is given as T table rows of Lp uint16 entries plus, per symbol, the
 * index of the row it is coded with (torchac takes one row per symbol; the
 * reference tiles the same [C, Lp] table N times, entropy_model.py:173).
 *
 * Build: see oracle/Makefile.  Only tests/, smoke() and bench.py's
 * cpu_baseline leg may load this.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    uint8_t *buf;
    size_t cap, len;
    uint8_t cur;
    int nbits;
} bitout_t;

static void put_bit(bitout_t *o, int bit) {
    o->cur = (uint8_t)((o->cur << 1) | (bit & 1));
    if (++o->nbits == 8) {
        if (o->len < o->cap) o->buf[o->len] = o->cur;
        o->len++;
        o->cur = 0;
        o->nbits = 0;
    }
}

static void put_bit_and_pending(bitout_t *o, int bit, uint64_t *pending) {
    put_bit(o, bit);
    while (*pending > 0) {
        put_bit(o, !bit);
        (*pending)--;
    }
}

/* returns the number of bytes the stream needs (writes at most cap of them) */
size_t rc_ref_encode(const uint16_t *cdf, int32_t Lp, const int32_t *row_of_sym,
                     const int16_t *sym, int64_t n_sym, uint8_t *out, size_t cap) {
    bitout_t o = {out, cap, 0, 0, 0};
    uint32_t low = 0, high = 0xFFFFFFFFu;
    uint64_t pending = 0;
    const int32_t max_symbol = Lp - 2;
    for (int64_t i = 0; i < n_sym; ++i) {
        const uint16_t *row = cdf + (size_t)row_of_sym[i] * Lp;
        const int32_t s = sym[i];
        const uint64_t span = (uint64_t)high - (uint64_t)low + 1;
        const uint32_t c_low = row[s];
        const uint32_t c_high = (s == max_symbol) ? 0x10000u : row[s + 1];
        high = (low - 1) + (uint32_t)((span * (uint64_t)c_high) >> 16);
        low = low + (uint32_t)((span * (uint64_t)c_low) >> 16);
        for (;;) {
            if (high < 0x80000000u) {
                put_bit_and_pending(&o, 0, &pending);
                low <<= 1;
                high = (high << 1) | 1u;
            } else if (low >= 0x80000000u) {
                put_bit_and_pending(&o, 1, &pending);
                low <<= 1;
                high = (high << 1) | 1u;
            } else if (low >= 0x40000000u && high < 0xC0000000u) {
                pending++;
                low = (low << 1) & 0x7FFFFFFFu;
                high = (high << 1) | 0x80000001u;
            } else {
                break;
            }
        }
    }
    pending += 1;
    put_bit_and_pending(&o, low < 0x40000000u ? 0 : 1, &pending);
    if (o.nbits > 0) {                       /* zero-pad the last byte */
        o.cur = (uint8_t)(o.cur << (8 - o.nbits));
        if (o.len < o.cap) o.buf[o.len] = o.cur;
        o.len++;
    }
    return o.len;
}

typedef struct {
    const uint8_t *buf;
    size_t len, pos;
    int nbits;
    uint8_t cur;
} bitin_t;

static uint32_t get_bit(bitin_t *in) {
    if (in->nbits == 0) {
        in->cur = (in->pos < in->len) ? in->buf[in->pos] : 0; /* past the end reads 0 */
        in->pos++;
        in->nbits = 8;
    }
    uint32_t b = (in->cur >> 7) & 1u;
    in->cur = (uint8_t)(in->cur << 1);
    in->nbits--;
    return b;
}

static int32_t find_symbol(const uint16_t *row, uint32_t target, int32_t max_symbol) {
    /* largest s in [0, max_symbol] with row[s] <= target (entry after the max
     * symbol counts as 65536) */
    int32_t left = 0, right = max_symbol + 1;
    while (left + 1 < right) {
        const int32_t m = (left + right) / 2;
        const uint32_t v = row[m];
        if (v < target) left = m;
        else if (v > target) right = m;
        else return m;
    }
    return left;
}

void rc_ref_decode(const uint16_t *cdf, int32_t Lp, const int32_t *row_of_sym,
                   const uint8_t *in, size_t in_len, int16_t *sym_out, int64_t n_sym) {
    bitin_t bi = {in, in_len, 0, 0, 0};
    uint32_t low = 0, high = 0xFFFFFFFFu, value = 0;
    const int32_t max_symbol = Lp - 2;
    for (int i = 0; i < 32; ++i) value = (value << 1) | get_bit(&bi);
    for (int64_t i = 0; i < n_sym; ++i) {
        const uint16_t *row = cdf + (size_t)row_of_sym[i] * Lp;
        const uint64_t span = (uint64_t)high - (uint64_t)low + 1;
        const uint16_t count =
            (uint16_t)((((uint64_t)value - (uint64_t)low + 1) * 0x10000ull - 1) / span);
        const int32_t s = find_symbol(row, count, max_symbol);
        sym_out[i] = (int16_t)s;
        if (i == n_sym - 1) break;
        const uint32_t c_low = row[s];
        const uint32_t c_high = (s == max_symbol) ? 0x10000u : row[s + 1];
        high = (low - 1) + (uint32_t)((span * (uint64_t)c_high) >> 16);
        low = low + (uint32_t)((span * (uint64_t)c_low) >> 16);
        for (;;) {
            if (low >= 0x80000000u || high < 0x80000000u) {
                low <<= 1;
                high = (high << 1) | 1u;
                value = (value << 1) | get_bit(&bi);
            } else if (low >= 0x40000000u && high < 0xC0000000u) {
                low = (low << 1) & 0x7FFFFFFFu;
                high = (high << 1) | 0x80000001u;
                value -= 0x40000000u;
                value = (value << 1) | get_bit(&bi);
            } else {
                break;
            }
        }
    }
}
