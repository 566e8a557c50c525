// octree_coder.cpp -- in-process lossless coder of the bottleneck coordinates.
//
// The reference hands the ~14 k stride-8 coordinates to the external MPEG G-PCC binary through an ASCII PLY file
// and a subprocess (CoordinateCoder, coder.py:17-36 -> gpcc.py:6-36): ~50 ms per call, an order of magnitude more
// than the whole GPU pass.  This is the in-process replacement: a breadth-first octree whose child-occupancy bits are
// range coded with adaptive binary models conditioned on
//   (a) an intra prediction from the 26 neighbours of the node at the PARENT level (all known: the parent level is
//       complete before the child level starts): distance-weighted occupancy score of the child, 8 bins,
//   (b) the three face neighbours of the child at the CHILD level that precede it in Morton order (siblings, or
//       children of the three "lower" neighbour nodes, which are already coded),
//   (c) how many children of this node are already known to be occupied (0, 1, >= 2),
// one model set per tree level.  Own format, NOT a G-PCC stream: on the bottleneck of synthetic_vox10(0) it spends
// 1.5 bits per point (2.6 KB) against tmc3's 1.0 (1.75 KB) -- Tmc3CoordinateCoder (coords_coder.py) keeps the
// reference's bit-exact path for parity runs.  Host code, synchronous; Codec runs it on a side thread.
//
// Stream: "PCO1", depth u8, n u32 LE, then the range coder's bytes.  Points are coded as a SET (duplicates collapse,
// order is not kept: the decoder returns them in Morton order and Codec re-imposes the canonical order, coder.py:97-99).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../include/pcgc.h"

namespace pcgc {
void set_error(const char *fmt, ...);

namespace {

// ---- binary adaptive range coder (carry-propagating, 32-bit range; two-rate probability mix) ----------------------
struct Model {
    uint16_t fast = 32768, slow = 32768;                     // P(bit = 1) * 65536 at adaptation shifts 4 and 7
    inline uint32_t p1() const { uint32_t p = ((uint32_t)fast + slow) >> 1; return p < 64 ? 64 : (p > 65472 ? 65472 : p); }
    inline void update(int bit) {
        if (bit) { fast += (65535 - fast) >> 4; slow += (65535 - slow) >> 7; }
        else { fast -= fast >> 4; slow -= slow >> 7; }
    }
};

struct Encoder {
    std::vector<uint8_t> out;
    uint64_t low = 0;
    uint32_t range = 0xFFFFFFFFu;
    uint8_t cache = 0;
    uint64_t cache_size = 1;
    void shift_low() {
        if ((uint32_t)low < 0xFF000000u || (low >> 32) != 0) {
            uint8_t carry = (uint8_t)(low >> 32), c = cache;
            do { out.push_back((uint8_t)(c + carry)); c = 0xFF; } while (--cache_size != 0);
            cache = (uint8_t)((uint32_t)low >> 24);
        }
        ++cache_size;
        low = (uint32_t)low << 8;
    }
    inline void encode(Model &m, int bit) {
        const uint32_t bound = (range >> 16) * m.p1();
        if (bit) range = bound;
        else { low += bound; range -= bound; }
        m.update(bit);
        while (range < (1u << 24)) { range <<= 8; shift_low(); }
    }
    void finish() { for (int i = 0; i < 5; ++i) shift_low(); }
};

struct Decoder {
    const uint8_t *p, *end;
    uint32_t range = 0xFFFFFFFFu, code = 0;
    Decoder(const uint8_t *b, const uint8_t *e) : p(b), end(e) { for (int i = 0; i < 5; ++i) code = (code << 8) | next(); }
    inline uint8_t next() { return p < end ? *p++ : 0; }
    inline int decode(Model &m) {
        const uint32_t bound = (range >> 16) * m.p1();
        int bit;
        if (code < bound) { range = bound; bit = 1; }
        else { code -= bound; range -= bound; bit = 0; }
        m.update(bit);
        while (range < (1u << 24)) { range <<= 8; code = (code << 8) | next(); }
        return bit;
    }
};

// ---- coordinate keys and the per-level node table ------------------------------------------------------------------
inline uint64_t spread3(uint64_t v) {                        // 21 bits -> every third bit
    v &= 0x1FFFFF;
    v = (v | v << 32) & 0x1F00000000FFFFull;
    v = (v | v << 16) & 0x1F0000FF0000FFull;
    v = (v | v << 8) & 0x100F00F00F00F00Full;
    v = (v | v << 4) & 0x10C30C30C30C30C3ull;
    v = (v | v << 2) & 0x1249249249249249ull;
    return v;
}
inline uint64_t morton(uint32_t x, uint32_t y, uint32_t z) { return spread3(x) | spread3(y) << 1 | spread3(z) << 2; }
inline uint32_t compact3(uint64_t v) {
    v &= 0x1249249249249249ull;
    v = (v | v >> 2) & 0x10C30C30C30C30C3ull;
    v = (v | v >> 4) & 0x100F00F00F00F00Full;
    v = (v | v >> 8) & 0x1F0000FF0000FFull;
    v = (v | v >> 16) & 0x1F00000000FFFFull;
    v = (v | v >> 32) & 0x1FFFFF;
    return (uint32_t)v;
}

struct NodeTable {                                           // Morton key -> node index (open addressing)
    std::vector<uint64_t> keys;
    std::vector<int32_t> vals;
    uint64_t mask = 0;
    void build(const std::vector<uint64_t> &nodes) {
        uint64_t cap = 16;
        while (cap < 2 * nodes.size() + 2) cap <<= 1;
        keys.assign(cap, ~0ull);
        vals.assign(cap, -1);
        mask = cap - 1;
        for (size_t i = 0; i < nodes.size(); ++i) {
            uint64_t h = (nodes[i] * 0x9E3779B97F4A7C15ull) >> 20 & mask;
            while (keys[h] != ~0ull) h = (h + 1) & mask;
            keys[h] = nodes[i];
            vals[h] = (int32_t)i;
        }
    }
    inline int32_t find(uint64_t k) const {
        uint64_t h = (k * 0x9E3779B97F4A7C15ull) >> 20 & mask;
        while (keys[h] != ~0ull) {
            if (keys[h] == k) return vals[h];
            h = (h + 1) & mask;
        }
        return -1;
    }
};

// intra-prediction weights: child j of a node against the node's 3x3x3 neighbourhood (index (dx+1)*9 + (dy+1)*3 + (dz+1),
// the centre weighs 0), w = 1024 / distance^3; tab[g][pattern][j] = summed weight of the occupied neighbours of x-plane g
struct Weights {
    int32_t w[8][27], total[8];
    int32_t tab[3][512][8];
    Weights() {
        for (int j = 0; j < 8; ++j) {
            total[j] = 0;
            // child centres sit at +-1/4 of the node: every squared distance is k/16 with integer k, the weights are rounded
            // from exactly representable inputs with correctly rounded IEEE operations -- identical on every host
            const double cx = (j & 1) ? 0.25 : -0.25, cy = (j & 2) ? 0.25 : -0.25, cz = (j & 4) ? 0.25 : -0.25;
            for (int i = 0; i < 27; ++i) {
                const int dx = i / 9 - 1, dy = (i / 3) % 3 - 1, dz = i % 3 - 1;
                if (!dx && !dy && !dz) { w[j][i] = 0; continue; }
                const double q = (dx - cx) * (dx - cx) + (dy - cy) * (dy - cy) + (dz - cz) * (dz - cz);
                w[j][i] = (int32_t)std::floor(1024.0 / (q * std::sqrt(q)) + 0.5);
                total[j] += w[j][i];
            }
        }
        for (int g = 0; g < 3; ++g)
            for (int p = 0; p < 512; ++p)
                for (int j = 0; j < 8; ++j) {
                    int32_t sum = 0;
                    for (int b = 0; b < 9; ++b)
                        if (p >> b & 1) sum += w[j][9 * g + b];
                    tab[g][p][j] = sum;
                }
    }
};
const Weights &weights() { static const Weights W; return W; }

constexpr int kMaxDepth = 21, kCtxPerLevel = 8 * 8 * 3, kDenseBits = 21;

// occupancy of one tree level: value 0 = empty, 0x100 | child-occupancy byte otherwise (the byte is 0 until the node is coded).
// Levels of up to 2^21 cells live in a dense grid with a one-cell border (no bounds checks, 27 plain loads per node);
// deeper levels fall back to a hash table.
struct LevelMap {
    bool dense = false;
    int dim = 0;                                             // dense: cells per axis + 2
    std::vector<uint16_t> grid;
    NodeTable table;
    std::vector<uint16_t> vals;
    void build(const std::vector<uint64_t> &nodes, int l) {
        dense = 3 * l <= kDenseBits;
        if (dense) {
            dim = (1 << l) + 2;
            grid.assign((size_t)dim * dim * dim, 0);
            for (uint64_t k : nodes) grid[index(compact3(k), compact3(k >> 1), compact3(k >> 2))] = 0x100;
        } else {
            table.build(nodes);
            vals.assign(nodes.size(), 0x100);
        }
    }
    inline size_t index(uint32_t x, uint32_t y, uint32_t z) const { return ((size_t)(x + 1) * dim + (y + 1)) * dim + (z + 1); }
};

// one pass over the tree; CODE(model, bit) either encodes the given bit or returns the decoded one
template <class Coder>
int64_t walk(int depth, std::vector<uint64_t> &level, const std::vector<uint64_t> *leaves, size_t max_nodes, Coder &code) {
    // level: Morton keys of the occupied nodes of the current level, ascending; starts as the root {0}
    const Weights &W = weights();
    std::vector<Model> models((size_t)depth * kCtxPerLevel);
    std::vector<uint64_t> next;
    LevelMap map;
    for (int l = 0; l < depth; ++l) {
        map.build(level, l);
        next.clear();
        Model *M = models.data() + (size_t)l * kCtxPerLevel;
        const int64_t lim = (int64_t)1 << l;                 // nodes of this level have coordinates in [0, lim)
        const int shift = 3 * (depth - l - 1);               // leaves >> shift = child-level keys
        size_t leaf_pos = 0;
        for (size_t ni = 0; ni < level.size(); ++ni) {
            const uint64_t key = level[ni];
            const uint32_t x = compact3(key), y = compact3(key >> 1), z = compact3(key >> 2);
            // (a) occupancy pattern of the 3x3x3 neighbourhood; (b) occupancy bytes of the three lower face neighbours
            uint32_t pat[3] = {0, 0, 0};
            uint32_t lower[3] = {0, 0, 0};
            size_t self = 0;
            if (map.dense) {
                self = map.index(x, y, z);
                const uint16_t *c = map.grid.data() + self;
                const ptrdiff_t sx = (ptrdiff_t)map.dim * map.dim, sy = map.dim;
                for (int g = 0; g < 3; ++g) {
                    const uint16_t *r = c + (g - 1) * sx;
                    uint32_t p = 0;
                    for (int dy = -1; dy <= 1; ++dy)
                        for (int dz = -1; dz <= 1; ++dz) p |= (uint32_t)(r[dy * sy + dz] != 0) << ((dy + 1) * 3 + (dz + 1));
                    pat[g] = p;
                }
                lower[0] = c[-sx] & 0xFF; lower[1] = c[-sy] & 0xFF; lower[2] = c[-1] & 0xFF;
            } else {
                for (int i = 0; i < 27; ++i) {
                    const int dx = i / 9 - 1, dy = (i / 3) % 3 - 1, dz = i % 3 - 1;
                    const int64_t nx = (int64_t)x + dx, ny = (int64_t)y + dy, nz = (int64_t)z + dz;
                    if (nx < 0 || ny < 0 || nz < 0 || nx >= lim || ny >= lim || nz >= lim) continue;
                    const int32_t idx = map.table.find(morton((uint32_t)nx, (uint32_t)ny, (uint32_t)nz));
                    if (idx < 0) continue;
                    pat[i / 9] |= 1u << (i % 9);
                    if (i == 4) lower[0] = map.vals[idx] & 0xFF;       // (-1, 0, 0)
                    if (i == 10) lower[1] = map.vals[idx] & 0xFF;      // (0, -1, 0)
                    if (i == 12) lower[2] = map.vals[idx] & 0xFF;      // (0, 0, -1)
                }
            }
            const int32_t *t0 = W.tab[0][pat[0]], *t1 = W.tab[1][pat[1]], *t2 = W.tab[2][pat[2]];
            uint8_t want = 0;
            if (leaves) {                                    // encoder: this node's true occupancy byte
                const uint64_t lo = key << 3;
                while (leaf_pos < leaves->size() && ((*leaves)[leaf_pos] >> shift) < lo) ++leaf_pos;
                size_t p = leaf_pos;
                while (p < leaves->size() && (((*leaves)[p] >> shift) >> 3) == key) { want |= (uint8_t)(1u << (((*leaves)[p] >> shift) & 7)); ++p; }
                leaf_pos = p;
            }
            uint32_t byte = 0;
            int cnt = 0;
            for (int j = 0; j < 8; ++j) {
                const int ix = j & 1, iy = (j >> 1) & 1, iz = (j >> 2) & 1;
                int sb = (int)(((int64_t)(t0[j] + t1[j] + t2[j]) * 16) / W.total[j]);
                if (sb > 7) sb = 7;
                const int fx = ix ? (byte >> (j - 1)) & 1 : (lower[0] >> (j + 1)) & 1;     // child at x-1: sibling j-1 / neighbour's child j+1
                const int fy = iy ? (byte >> (j - 2)) & 1 : (lower[1] >> (j + 2)) & 1;
                const int fz = iz ? (byte >> (j - 4)) & 1 : (lower[2] >> (j + 4)) & 1;
                Model &m = M[(sb * 8 + (fx | fy << 1 | fz << 2)) * 3 + (cnt > 2 ? 2 : cnt)];
                const int bit = code(m, (want >> j) & 1);
                byte |= (uint32_t)bit << j;
                cnt += bit;
            }
            if (!byte || next.size() > max_nodes) return -1; // corrupt stream: a childless node / more nodes than points
            if (map.dense) map.grid[self] = (uint16_t)(0x100 | byte);
            else map.vals[ni] = (uint16_t)(0x100 | byte);
            for (int j = 0; j < 8; ++j)
                if (byte >> j & 1) next.push_back(key << 3 | (uint64_t)j);
        }
        level.swap(next);
    }
    return (int64_t)level.size();
}

}  // namespace
}  // namespace pcgc

using namespace pcgc;

extern "C" {

int64_t pcgc_octree_encode_host(const int32_t *coords_host, int64_t n, uint8_t *out_host, int64_t cap) {
    if (n < 0 || (n > 0 && !coords_host) || !out_host || cap < 9) { set_error("pcgc_octree_encode_host: bad arguments"); return PCGC_ERR_INVALID; }
    std::vector<uint64_t> leaves((size_t)n);
    uint32_t mx = 0;
    for (int64_t i = 0; i < n; ++i) {
        const int32_t x = coords_host[3 * i], y = coords_host[3 * i + 1], z = coords_host[3 * i + 2];
        if (x < 0 || y < 0 || z < 0 || x >= (1 << kMaxDepth) || y >= (1 << kMaxDepth) || z >= (1 << kMaxDepth)) {
            set_error("pcgc_octree_encode_host: coordinate outside [0, 2^21)");
            return PCGC_ERR_RANGE;
        }
        mx |= (uint32_t)x | (uint32_t)y | (uint32_t)z;
        leaves[(size_t)i] = morton((uint32_t)x, (uint32_t)y, (uint32_t)z);
    }
    std::sort(leaves.begin(), leaves.end());
    leaves.erase(std::unique(leaves.begin(), leaves.end()), leaves.end());
    int depth = 0;
    while (depth < kMaxDepth && (mx >> depth)) ++depth;
    Encoder enc;
    if (!leaves.empty() && depth > 0) {
        std::vector<uint64_t> level{0};
        auto code = [&](Model &m, int bit) { enc.encode(m, bit); return bit; };
        walk(depth, level, &leaves, leaves.size(), code);
        enc.finish();
    }
    const int64_t total = 9 + (int64_t)enc.out.size();
    if (total > cap) { set_error("pcgc_octree_encode_host: need %lld bytes", (long long)total); return PCGC_ERR_WORKSPACE; }
    memcpy(out_host, "PCO1", 4);
    out_host[4] = (uint8_t)depth;
    const uint32_t cnt = (uint32_t)leaves.size();
    memcpy(out_host + 5, &cnt, 4);
    if (!enc.out.empty()) memcpy(out_host + 9, enc.out.data(), enc.out.size());
    return total;
}

int64_t pcgc_octree_decode_host(const uint8_t *in_host, int64_t len, int32_t *coords_host, int64_t cap_rows) {
    if (!in_host || len < 9 || memcmp(in_host, "PCO1", 4) != 0 || in_host[4] > kMaxDepth) {
        set_error("pcgc_octree_decode_host: not a PCO1 stream");
        return PCGC_ERR_INVALID;
    }
    const int depth = in_host[4];
    uint32_t cnt;
    memcpy(&cnt, in_host + 5, 4);
    if (!coords_host) return (int64_t)cnt;                   // size query
    if ((int64_t)cnt > cap_rows) { set_error("pcgc_octree_decode_host: %u points, room for %lld", cnt, (long long)cap_rows); return PCGC_ERR_WORKSPACE; }
    if (cnt == 0) return 0;
    std::vector<uint64_t> level{0};
    if (depth > 0) {
        Decoder dec(in_host + 9, in_host + len);
        auto code = [&](Model &m, int) { return dec.decode(m); };
        const int64_t got = walk(depth, level, nullptr, (size_t)cnt, code);
        if (got != (int64_t)cnt) { set_error("pcgc_octree_decode_host: corrupt stream (%lld of %u points)", (long long)got, cnt); return PCGC_ERR_INVALID; }
    } else if (cnt != 1) {
        set_error("pcgc_octree_decode_host: corrupt stream");
        return PCGC_ERR_INVALID;
    }
    for (uint32_t i = 0; i < cnt; ++i) {
        coords_host[3 * i] = (int32_t)compact3(level[i]);
        coords_host[3 * i + 1] = (int32_t)compact3(level[i] >> 1);
        coords_host[3 * i + 2] = (int32_t)compact3(level[i] >> 2);
    }
    return (int64_t)cnt;
}

}  // extern "C"
