// rangecoder.cpp -- host range coder, bit-compatible with torchac 0.9.3's
// encode_float_cdf / decode_float_cdf (reference call sites entropy_model.py:174,192;
// SURVEY section 8 row a15, Appendix B).  The stream is one sequential arithmetic-coded
// sequence (32-bit interval, 16-bit CDF precision, carry via pending bits), so the coding
// loop stays on the host; everything around it (symbols, tables) is produced on the GPU.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../include/pcgc.h"

namespace pcgc {
void set_error(const char *fmt, ...);
}

namespace {

constexpr uint32_t kHalf = 0x80000000u, kQuarter = 0x40000000u, kThreeQuarter = 0xC0000000u;

// MSB-first bit sink with a 64-bit staging word
class BitSink {
public:
    BitSink(uint8_t *out, int64_t cap) : out_(out), cap_(cap) {}
    inline void put(uint32_t bit) {
        acc_ = (acc_ << 1) | bit;
        if (++fill_ == 64) flush_word();
    }
    inline void put_with_pending(uint32_t bit, uint64_t &pending) {
        put(bit);
        const uint32_t inv = bit ^ 1u;
        for (; pending; --pending) put(inv);
    }
    int64_t finish() {                       // zero-pad to a byte boundary
        while (fill_ & 7) { acc_ <<= 1; ++fill_; }
        for (int i = fill_ - 8; i >= 0; i -= 8) emit((uint8_t)(acc_ >> i));
        fill_ = 0;
        return len_;
    }

private:
    inline void emit(uint8_t b) {
        if (len_ < cap_) out_[len_] = b;
        ++len_;
    }
    void flush_word() {
        for (int i = 56; i >= 0; i -= 8) emit((uint8_t)(acc_ >> i));
        acc_ = 0;
        fill_ = 0;
    }
    uint8_t *out_;
    int64_t cap_, len_ = 0;
    uint64_t acc_ = 0;
    int fill_ = 0;
};

class BitSource {
public:
    BitSource(const uint8_t *in, int64_t len) : in_(in), len_(len) {}
    inline uint32_t get() {                  // reads past the end return 0
        if (left_ == 0) {
            cur_ = pos_ < len_ ? in_[pos_] : 0;
            ++pos_;
            left_ = 8;
        }
        --left_;
        return (cur_ >> left_) & 1u;
    }

private:
    const uint8_t *in_;
    int64_t len_, pos_ = 0;
    uint32_t cur_ = 0;
    int left_ = 0;
};

// Appendix B.1: round(cdf * (2^16 - (Lp-1))) as int16 (wraps) + arange(Lp), bits as uint16
void float_table_to_u16(const float *cdf, int64_t n_tables, int32_t lp, std::vector<uint16_t> &out) {
    out.resize((size_t)n_tables * lp);
    const float scale = (float)(65536 - (lp - 1));
    for (int64_t t = 0; t < n_tables; ++t)
        for (int32_t j = 0; j < lp; ++j) {
            const float r = nearbyintf(cdf[t * lp + j] * scale);       // half-to-even, like torch.round
            out[(size_t)t * lp + j] = (uint16_t)((int64_t)r + j);
        }
}

struct Interval {
    uint32_t low = 0, high = 0xFFFFFFFFu;
    inline void narrow(uint32_t c_low, uint32_t c_high) {
        const uint64_t span = (uint64_t)high - (uint64_t)low + 1;
        high = (low - 1) + (uint32_t)((span * c_high) >> 16);
        low = low + (uint32_t)((span * c_low) >> 16);
    }
};

// one coding step of Appendix B.2: narrow to [c_low, c_high) / 2^16, then renormalise
static inline void encode_step(Interval &iv, BitSink &sink, uint64_t &pending, uint32_t c_low, uint32_t c_high) {
        iv.narrow(c_low, c_high);
        for (;;) {
            if (iv.high < kHalf) {
                sink.put_with_pending(0, pending);
            } else if (iv.low >= kHalf) {
                sink.put_with_pending(1, pending);
            } else if (iv.low >= kQuarter && iv.high < kThreeQuarter) {
                ++pending;
                iv.low = (iv.low << 1) & 0x7FFFFFFFu;
                iv.high = (iv.high << 1) | 0x80000001u;
                continue;
            } else {
                break;
            }
            iv.low <<= 1;
            iv.high = (iv.high << 1) | 1u;
        }
}

int64_t encode_u16(const uint16_t *cdf, int64_t n_tables, int32_t lp, const int16_t *sym, int64_t n_sym, uint8_t *out,
                   int64_t cap) {
    BitSink sink(out, cap);
    Interval iv;
    uint64_t pending = 0;
    const int32_t max_symbol = lp - 2;
    int64_t t = 0;
    for (int64_t i = 0; i < n_sym; ++i) {
        const uint16_t *row = cdf + t * lp;
        if (++t == n_tables) t = 0;
        const int32_t s = sym[i];
        if (s < 0 || s > max_symbol) {
            pcgc::set_error("pcgc_rc_encode: symbol %d at %lld outside [0, %d]", s, (long long)i, max_symbol);
            return PCGC_ERR_RANGE;
        }
        encode_step(iv, sink, pending, row[s], s == max_symbol ? 0x10000u : row[s + 1]);
    }
    ++pending;
    sink.put_with_pending(iv.low < kQuarter ? 0u : 1u, pending);
    return sink.finish();
}

int64_t encode_ranges(const uint32_t *ranges, int64_t n_sym, uint8_t *out, int64_t cap) {
    BitSink sink(out, cap);
    Interval iv;
    uint64_t pending = 0;
    for (int64_t i = 0; i < n_sym; ++i) {
        const uint32_t r = ranges[i], c_low = r & 0xFFFFu, c_high = (r >> 16) + 1u;
        if (c_high <= c_low) {
            pcgc::set_error("pcgc_rc_encode_ranges: empty interval at %lld", (long long)i);
            return PCGC_ERR_RANGE;
        }
        encode_step(iv, sink, pending, c_low, c_high);
    }
    ++pending;
    sink.put_with_pending(iv.low < kQuarter ? 0u : 1u, pending);
    return sink.finish();
}

int decode_u16(const uint16_t *cdf, int64_t n_tables, int32_t lp, const uint8_t *in, int64_t in_len, int16_t *sym,
               int64_t n_sym) {
    BitSource src(in, in_len);
    Interval iv;
    uint32_t value = 0;
    for (int i = 0; i < 32; ++i) value = (value << 1) | src.get();
    const int32_t max_symbol = lp - 2;
    int64_t t = 0;
    for (int64_t i = 0; i < n_sym; ++i) {
        const uint16_t *row = cdf + t * lp;
        if (++t == n_tables) t = 0;
        const uint64_t span = (uint64_t)iv.high - (uint64_t)iv.low + 1;
        const uint32_t count = (uint32_t)(((((uint64_t)value - iv.low + 1) << 16) - 1) / span) & 0xFFFFu;
        // largest s in [0, max_symbol] with row[s] <= count (rows are strictly increasing)
        int32_t lo = 0, hi = max_symbol + 1;
        while (lo + 1 < hi) {
            const int32_t mid = (lo + hi) >> 1;
            if (row[mid] <= count) lo = mid; else hi = mid;
        }
        sym[i] = (int16_t)lo;
        if (i == n_sym - 1) break;
        iv.narrow(row[lo], lo == max_symbol ? 0x10000u : row[lo + 1]);
        for (;;) {
            if (iv.low >= kHalf || iv.high < kHalf) {
                // plain shift
            } else if (iv.low >= kQuarter && iv.high < kThreeQuarter) {
                iv.low &= 0x3FFFFFFFu;           // together with the shift: (low << 1) & 0x7FFFFFFF
                iv.high |= kQuarter;             // together with the shift/or: (high << 1) | 0x80000001
                value -= kQuarter;
            } else {
                break;
            }
            iv.low <<= 1;
            iv.high = (iv.high << 1) | 1u;
            value = (value << 1) | src.get();
        }
    }
    return PCGC_OK;
}

bool bad_args(const void *cdf, int64_t n_tables, int32_t lp, const void *sym, int64_t n_sym) {
    return !cdf || n_tables < 1 || lp < 2 || n_sym < 0 || (n_sym > 0 && !sym);
}

}  // namespace

extern "C" {

int64_t pcgc_rc_encode_ranges_host(const uint32_t *ranges_host, int64_t n_sym, uint8_t *out_host, int64_t cap) {
    if (n_sym < 0 || (n_sym > 0 && !ranges_host)) {
        pcgc::set_error("pcgc_rc_encode_ranges: bad arguments");
        return PCGC_ERR_INVALID;
    }
    return encode_ranges(ranges_host, n_sym, out_host, cap);
}

int64_t pcgc_rc_encode_u16_host(const uint16_t *cdf_u16_host, int64_t n_tables, int32_t lp, const int16_t *sym_host,
                                int64_t n_sym, uint8_t *out_host, int64_t cap) {
    if (bad_args(cdf_u16_host, n_tables, lp, sym_host, n_sym)) {
        pcgc::set_error("pcgc_rc_encode: bad arguments");
        return PCGC_ERR_INVALID;
    }
    return encode_u16(cdf_u16_host, n_tables, lp, sym_host, n_sym, out_host, cap);
}

int pcgc_rc_decode_u16_host(const uint16_t *cdf_u16_host, int64_t n_tables, int32_t lp, const uint8_t *in_host,
                            int64_t in_len, int16_t *sym_host, int64_t n_sym) {
    if (bad_args(cdf_u16_host, n_tables, lp, sym_host, n_sym) || in_len < 0) {
        pcgc::set_error("pcgc_rc_decode: bad arguments");
        return PCGC_ERR_INVALID;
    }
    return decode_u16(cdf_u16_host, n_tables, lp, in_host, in_len, sym_host, n_sym);
}

int64_t pcgc_rc_encode_host(const float *cdf_float_host, int64_t n_tables, int32_t lp, const int16_t *sym_host,
                            int64_t n_sym, uint8_t *out_host, int64_t cap) {
    if (bad_args(cdf_float_host, n_tables, lp, sym_host, n_sym)) {
        pcgc::set_error("pcgc_rc_encode: bad arguments");
        return PCGC_ERR_INVALID;
    }
    std::vector<uint16_t> table;
    float_table_to_u16(cdf_float_host, n_tables, lp, table);
    return encode_u16(table.data(), n_tables, lp, sym_host, n_sym, out_host, cap);
}

int pcgc_rc_decode_host(const float *cdf_float_host, int64_t n_tables, int32_t lp, const uint8_t *in_host,
                        int64_t in_len, int16_t *sym_host, int64_t n_sym) {
    if (bad_args(cdf_float_host, n_tables, lp, sym_host, n_sym) || in_len < 0) {
        pcgc::set_error("pcgc_rc_decode: bad arguments");
        return PCGC_ERR_INVALID;
    }
    std::vector<uint16_t> table;
    float_table_to_u16(cdf_float_host, n_tables, lp, table);
    return decode_u16(table.data(), n_tables, lp, in_host, in_len, sym_host, n_sym);
}

}  // extern "C"
