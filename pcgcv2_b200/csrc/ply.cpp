// ply.cpp -- ASCII PLY geometry ingest / egress on the host.
//
// Replaces the per-line Python loops of the reference's read_ply_ascii_geo / write_ply_ascii_geo
// (data_utils.py:19-48; called from coder.py:26,33,128,177 and load_sparse_tensor, data_utils.py:103-110): one pass
// over the text buffer, straight into / out of an int32 [n,3] array that the caller keeps in PINNED memory, so the
// host->device copy of the coordinates is one asynchronous DMA.
//
// Reader semantics follow the reference line by line: a line is split at single spaces, every token (other than the
// bare line terminator) must parse as a float or the WHOLE line is skipped (that is how the reference steps over the
// header, comments and `element` / `property` lines without looking for `end_header`); the first three values of a
// surviving line are truncated toward zero (numpy `.astype('int')`).  Lines with fewer than three values, which make
// the reference's np.array ragged and crash, are skipped.
#include <cerrno>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../include/pcgc.h"

namespace pcgc {
void set_error(const char *fmt, ...);

// fast path: [+-]digits[.digits] ; anything else goes through strtod on a bounded copy (exponents, inf, nan)
static inline bool parse_token(const char *b, const char *e, double &out) {
    const char *p = b;
    bool neg = false;
    if (p < e && (*p == '-' || *p == '+')) neg = *p++ == '-';
    const char *d0 = p;
    int64_t ip = 0;
    while (p < e && *p >= '0' && *p <= '9' && p - d0 < 18) ip = ip * 10 + (*p++ - '0');
    if (p == e && p > d0) { out = neg ? -(double)ip : (double)ip; return true; }
    if (p < e && *p == '.' && p - d0 < 18) {
        const char *f0 = ++p;
        double frac = 0.0, scale = 1.0;
        while (p < e && *p >= '0' && *p <= '9' && p - f0 < 18) { frac = frac * 10.0 + (*p++ - '0'); scale *= 10.0; }
        if (p == e && (p > f0 || f0 - 1 > d0)) { out = (double)ip + frac / scale; if (neg) out = -out; return true; }
    }
    char buf[64];
    const size_t len = (size_t)(e - b);
    if (len == 0 || len >= sizeof(buf)) return false;
    memcpy(buf, b, len);
    buf[len] = 0;
    char *end = nullptr;
    errno = 0;
    out = strtod(buf, &end);
    while (end && (*end == ' ' || *end == '\t' || *end == '\r' || *end == '\n')) ++end;   // float() strips whitespace
    return end == buf + len && end != buf;
}
}  // namespace pcgc

using namespace pcgc;

extern "C" {

int64_t pcgc_ply_count_lines_host(const char *text_host, int64_t len) {
    if (!text_host || len < 0) { set_error("pcgc_ply_count_lines_host: bad buffer"); return PCGC_ERR_INVALID; }
    int64_t n = 0;
    const char *p = text_host, *e = text_host + len;
    while (p < e) {
        const char *nl = (const char *)memchr(p, '\n', (size_t)(e - p));
        ++n;
        if (!nl) break;
        p = nl + 1;
    }
    return n;
}

int64_t pcgc_ply_parse_ascii_host(const char *text_host, int64_t len, int32_t *coords_host, int64_t cap_rows) {
    if (!text_host || len < 0 || (!coords_host && cap_rows > 0)) { set_error("pcgc_ply_parse_ascii_host: bad buffer"); return PCGC_ERR_INVALID; }
    int64_t rows = 0;
    const char *p = text_host, *e = text_host + len;
    while (p < e) {
        const char *nl = (const char *)memchr(p, '\n', (size_t)(e - p));
        const char *le = nl ? nl : e;                       // line = [p, le), terminator excluded
        double v[3];
        int nv = 0;
        bool ok = true;
        const char *t = p;
        while (ok) {                                        // tokens between single spaces, like str.split(' ')
            const char *sp = (const char *)memchr(t, ' ', (size_t)(le - t));
            const char *te = sp ? sp : le;
            const char *tb = t, *tt = te;
            if (!sp) {                                      // last token carries the terminator in the reference: '\n' alone is
                while (tt > tb && (tt[-1] == '\r')) --tt;   // skipped, '3\n' parses as 3 (float() strips whitespace)
                if (tt == tb) break;
            }
            while (tb < tt && (*tb == '\t' || *tb == '\r')) ++tb;
            double x;
            if (!parse_token(tb, tt, x)) { ok = false; break; }
            if (nv < 3) v[nv] = x;
            ++nv;
            if (!sp) break;
            t = sp + 1;
        }
        if (ok && nv >= 3) {
            if (rows >= cap_rows) { set_error("pcgc_ply_parse_ascii_host: more than %lld vertex lines", (long long)cap_rows); return PCGC_ERR_WORKSPACE; }
            for (int i = 0; i < 3; ++i) {
                if (!(std::fabs(v[i]) < 2147483648.0)) { set_error("pcgc_ply_parse_ascii_host: coordinate out of int32 range"); return PCGC_ERR_RANGE; }
                coords_host[3 * rows + i] = (int32_t)v[i];  // truncation toward zero, numpy astype('int')
            }
            ++rows;
        }
        if (!nl) break;
        p = nl + 1;
    }
    return rows;
}

int64_t pcgc_ply_format_ascii_host(const int32_t *coords_host, int64_t n, char *text_host, int64_t cap) {
    if (n < 0 || (!coords_host && n > 0) || !text_host) { set_error("pcgc_ply_format_ascii_host: bad buffer"); return PCGC_ERR_INVALID; }
    char head[160];
    const int hl = snprintf(head, sizeof(head), "ply\nformat ascii 1.0\nelement vertex %lld\nproperty float x\nproperty float y\n"
                                                "property float z\nend_header\n", (long long)n);
    if (cap < hl + n * 36) { set_error("pcgc_ply_format_ascii_host: need %lld bytes", (long long)(hl + n * 36)); return PCGC_ERR_WORKSPACE; }
    char *o = text_host;
    memcpy(o, head, (size_t)hl);
    o += hl;
    for (int64_t r = 0; r < n; ++r) {
        for (int i = 0; i < 3; ++i) {
            int64_t x = coords_host[3 * r + i];
            if (x < 0) { *o++ = '-'; x = -x; }
            char d[12];
            int k = 0;
            do { d[k++] = (char)('0' + x % 10); x /= 10; } while (x);
            while (k) *o++ = d[--k];
            *o++ = i < 2 ? ' ' : '\n';
        }
    }
    return (int64_t)(o - text_host);
}

}  // extern "C"
