// key.cuh -- packed-key primitives of the B200 clustering path.
//
// A key (the --check-lengths slices of one read tuple, reference
// src/fastqdedup/__init__.py:160-167, :251) is stored bit-sliced: K bit planes of PW
// 32-bit words each, plane-major (w[p*PW + i]); bit (i%32) of word i/32 of plane p is bit
// p of the code of symbol i.  With the default alphabet "ACGTN" (reference
// __init__.py:240) K = 3: planes 0/1 are the classic 2-bit base code and plane 2 is the
// N mask, which is the "2-bit packed with an N mask" layout BASELINE.json asks for,
// generalised to any alphabet of up to 2^K - 1 symbols (one code is reserved as PAD).
//
// Variable-length keys (reads shorter than the check length) are padded to the job's
// maximum length with the PAD code, which no real symbol uses; two keys are the same
// string iff all their plane words are equal, so hashing/equality never look at a length.
//
// Everything here is __host__ __device__ so tests/test_key_primitives.py can exercise the
// exact same code on the CPU (through csrc/host_probe.cu) where no GPU exists; the
// product never calls the host instantiations.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define FQD_HD __host__ __device__ __forceinline__
#else
#define FQD_HD inline
#endif

namespace fqd {

#if defined(__CUDA_ARCH__)
FQD_HD int popc32(uint32_t x) { return __popc(x); }
FQD_HD int ctz32(uint32_t x) { return __ffs((int)x) - 1; }
#else
FQD_HD int popc32(uint32_t x) { return __builtin_popcount(x); }
FQD_HD int ctz32(uint32_t x) { return __builtin_ctz(x); }
#endif

// Symbol coding of one job: byte -> code (0xFF = not in the alphabet), code -> rank in
// ASCII order (PAD ranks below everything so a proper prefix sorts first, like Python's
// str comparison used by sorted() at reference __init__.py:68, :99, :111).
struct Codec {
    uint8_t lut[256];
    uint8_t rank[256];
    uint8_t pad_code;
    uint8_t bits;      // K
    uint8_t n_symbols;
    uint8_t varlen;    // 1: keys of several lengths are present (PAD in use)
    uint8_t swar;      // 1: the ACGTN code of pack_key_acgtn (codes = ASCII bits 1..3) is in use
};

// Codes of the five DNA letters when the code is simply bits 1..3 of the ASCII byte:
// A 0x41 -> 0, C 0x43 -> 1, T 0x54 -> 2, G 0x47 -> 3, N 0x4E -> 7; 4 is free and serves as PAD.
constexpr uint32_t SWAR_PAD_CODE = 4;

template <int K, int PW>
struct Key {
    static constexpr int NW = K * PW;
    uint32_t w[NW];
};

FQD_HD uint64_t mix64(uint64_t h)
{
    h ^= h >> 33; h *= 0xff51afd7ed558ccdULL;
    h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ULL;
    h ^= h >> 33;
    return h;
}

FQD_HD uint32_t fmix32(uint32_t h)
{
    h ^= h >> 16; h *= 0x85EBCA6Bu;
    h ^= h >> 13; h *= 0xC2B2AE35u;
    h ^= h >> 16;
    return h;
}
FQD_HD uint32_t rotl32(uint32_t x, int r) { r &= 31; return r ? (x << r) | (x >> (32 - r)) : x; }

// Two independent 32-bit lanes; every word goes through its own odd multiplier (a bijection of
// the word) and the products are combined by rotate-xor, so the 2*NW multiplies are independent
// of each other (the 64-bit multiply chain this replaces was a quarter of the ingest kernel's
// instructions).  High word: partition / owner / table slot of the HBM table; low word: slot in
// a shared-memory tile.
FQD_HD void hash_lanes(uint32_t v, int i, uint32_t &a, uint32_t &b)
{
    const uint32_t ca = 0x9E3779B1u * (2u * (uint32_t)i + 1u), cb = 0x85EBCA77u * (2u * (uint32_t)i + 3u);
    a ^= rotl32((v ^ (0x7F4A7C15u + 0x01000193u * (uint32_t)i)) * (ca | 1u), 7 * i + 3);
    b ^= rotl32((v + (0x2545F491u ^ (0x9E3779B9u * (uint32_t)i))) * (cb | 1u), 11 * i + 5);
}

template <int K, int PW>
FQD_HD uint64_t hash_key(const Key<K, PW> &k)
{
    uint32_t a = 0x243F6A88u, b = 0x85A308D3u;
#pragma unroll
    for (int i = 0; i < K * PW; i++) hash_lanes(k.w[i], i, a, b);
    return ((uint64_t)fmix32(a ^ rotl32(b, 16)) << 32) | fmix32(b + 0x9E3779B9u * a);
}

// One 32-bit lane of the same construction: the slot of a key inside a shared-memory tile (the
// tile's keys already share the partition bits, any well-mixed function of the key will do).
template <int K, int PW>
FQD_HD uint32_t hash_key32(const Key<K, PW> &k)
{
    uint32_t a = 0x243F6A88u;
#pragma unroll
    for (int i = 0; i < K * PW; i++) a ^= rotl32((k.w[i] ^ (0x7F4A7C15u + 0x01000193u * (uint32_t)i)) * ((0x9E3779B1u * (2u * (uint32_t)i + 1u)) | 1u), 7 * i + 3);
    return fmix32(a);
}

template <int K, int PW>
FQD_HD bool key_equal(const Key<K, PW> &a, const Key<K, PW> &b)
{
    uint32_t d = 0;
#pragma unroll
    for (int i = 0; i < K * PW; i++) d |= a.w[i] ^ b.w[i];
    return d == 0;
}

// Pack `len` bytes (symbols beyond len up to padded_len become PAD).  Returns false when
// a byte is not in the codec's alphabet; *bad receives that byte.
template <int K, int PW>
FQD_HD bool pack_key(const uint8_t *bytes, uint32_t len, uint32_t padded_len,
                     const uint8_t *lut, uint32_t pad_code, Key<K, PW> &out, uint32_t *bad)
{
    bool ok = true;
#pragma unroll
    for (int wi = 0; wi < PW; wi++) {
        uint32_t acc[K];
#pragma unroll
        for (int p = 0; p < K; p++) acc[p] = 0;
        const uint32_t base = wi * 32;
        for (uint32_t b = 0; b < 32 && base + b < padded_len; b++) {
            uint32_t c;
            if (base + b < len) {
                c = lut[bytes[base + b]];
                if (c == 0xFF) { ok = false; *bad = bytes[base + b]; c = 0; }
            } else {
                c = pad_code;
            }
#pragma unroll
            for (int p = 0; p < K; p++) acc[p] |= ((c >> p) & 1u) << b;
        }
#pragma unroll
        for (int p = 0; p < K; p++) out.w[p * PW + wi] = acc[p];
    }
    return ok;
}

#if defined(__CUDA_ARCH__)
FQD_HD uint32_t funnel_l(uint32_t lo, uint32_t hi, uint32_t s) { return __funnelshift_l(lo, hi, s); }
#else
FQD_HD uint32_t funnel_l(uint32_t lo, uint32_t hi, uint32_t s) { return s ? (hi << s) | (lo >> (32u - s)) : hi; }
#endif

// SWAR packing for the ACGTN alphabet: four ASCII bytes per 32-bit word, no table in memory.  The
// code of a letter is bits 1..3 of its byte (A 0, C 1, T 2, G 3, N 7).  `words` must be 4-byte aligned.
//
// Validity: the four codes are gathered into the four selector nibbles of a byte permute (PRMT), which
// looks the ASCII byte each code stands for up in an 8-byte register table; a byte that differs from its
// own code's letter is foreign (codes 4..6 map to 0xFF, whose own code is 7, so they never match).  Five
// instructions and one accumulate per word; a foreign byte makes the caller report the key's bytes one
// by one through the table path.
//
// Per word and plane: mask the four code bits where they sit (bit 8k+1+p), one multiply moves
// them to the top nibble (bits 28..31; the cross terms land on distinct lower bits or overflow,
// so no carry reaches the nibble) and one funnel shift appends that nibble to the plane word --
// three instructions.  Words are visited from the last to the first so nibble j ends at bit 4j.
#if defined(__CUDA_ARCH__)
FQD_HD uint32_t byte_perm(uint32_t a, uint32_t b, uint32_t s) { return __byte_perm(a, b, s); }
#else
FQD_HD uint32_t byte_perm(uint32_t a, uint32_t b, uint32_t s)
{
    const uint64_t v = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int k = 0; k < 4; k++) r |= (uint32_t)((v >> (8u * ((s >> (4 * k)) & 7u))) & 0xFFu) << (8 * k);
    return r;
}
#endif

struct SwarCheck {
    uint32_t bad = 0;
    FQD_HD bool ok() const { return bad == 0; }
};

// four symbols: validity + the three plane nibbles appended to p0/p1/p2
FQD_HD void swar_step(uint32_t w, SwarCheck &c, uint32_t &p0, uint32_t &p1, uint32_t &p2)
{
    const uint32_t codes = (w >> 1) & 0x07070707u;                        // one code per byte
    const uint32_t sel = byte_perm(codes | (codes >> 4), 0u, 0x4420u);    // nibble k = code of byte k
    c.bad |= byte_perm(0x47544341u, 0x4EFFFFFFu, sel) ^ w;                // "ACTG", 0xFF x3, 'N'
    p0 = funnel_l((w & 0x02020202u) * 0x08102040u, p0, 4);
    p1 = funnel_l((w & 0x04040404u) * 0x04081020u, p1, 4);
    p2 = funnel_l((w & 0x08080808u) * 0x02040810u, p2, 4);
}

template <int PW>
FQD_HD bool pack_key_acgtn(const uint32_t *words, uint32_t len, uint32_t padded_len, Key<3, PW> &out)
{
    SwarCheck chk;
#pragma unroll
    for (int wi = 0; wi < PW; wi++) {
        uint32_t p0 = 0, p1 = 0, p2 = 0;
        const uint32_t base = 32u * wi;
#pragma unroll
        for (int j = 7; j >= 0; j--) {
            const uint32_t pos = base + 4u * j;
            if (pos < len) {
                uint32_t w = words[pos >> 2];
                const uint32_t rem = len - pos;                 // symbols of this word that count
                if (rem < 4) {
                    const uint32_t keep = (1u << (8u * rem)) - 1u;
                    w = (w & keep) | (0x41414141u & ~keep);     // the rest reads as 'A' (code 0)
                }
                swar_step(w, chk, p0, p1, p2);
            }
        }
        // PAD (code 4 = plane 2 only) on [len, padded_len)
        if (padded_len > len && padded_len > base && len < base + 32u) {
            const uint32_t lo = len > base ? len - base : 0u;
            const uint32_t hi = padded_len - base >= 32u ? 32u : padded_len - base;
            const uint32_t upto = hi >= 32u ? 0xFFFFFFFFu : ((1u << hi) - 1u);
            const uint32_t from = lo >= 32u ? 0xFFFFFFFFu : ((1u << lo) - 1u);
            p2 |= upto & ~from;
        }
        out.w[0 * PW + wi] = p0;
        out.w[1 * PW + wi] = p1;
        out.w[2 * PW + wi] = p2;
    }
    return chk.ok() || len == 0;
}

// The same for keys of exactly 4*NW symbols (no padding): straight-line code, no length tests.
template <int PW, int NW>
FQD_HD bool pack_key_acgtn_fixed(const uint32_t *words, Key<3, PW> &out)
{
    static_assert(4 * NW <= 32 * PW, "key does not fit");
    SwarCheck chk;
#pragma unroll
    for (int wi = 0; wi < PW; wi++) {
        uint32_t p0 = 0, p1 = 0, p2 = 0;
#pragma unroll
        for (int j = 7; j >= 0; j--)
            if (wi * 8 + j < NW) swar_step(words[wi * 8 + j], chk, p0, p1, p2);
        out.w[0 * PW + wi] = p0;
        out.w[1 * PW + wi] = p1;
        out.w[2 * PW + wi] = p2;
    }
    return chk.ok();
}

// Bit mask (per plane word i) of the positions holding PAD.
template <int K, int PW>
FQD_HD uint32_t pad_mask_word(const Key<K, PW> &a, int i, uint32_t pad_code)
{
    uint32_t m = 0xFFFFFFFFu;
#pragma unroll
    for (int p = 0; p < K; p++) m &= ((pad_code >> p) & 1u) ? a.w[p * PW + i] : ~a.w[p * PW + i];
    return m;
}

// Length of a padded key: index of its first PAD symbol, or max_len.
template <int K, int PW>
FQD_HD uint32_t key_length(const Key<K, PW> &a, uint32_t pad_code, uint32_t max_len)
{
#pragma unroll
    for (int i = 0; i < PW; i++) {
        uint32_t m = pad_mask_word(a, i, pad_code);
        if (m) {
            uint32_t pos = 32u * i + (uint32_t)ctz32(m);
            return pos < max_len ? pos : max_len;
        }
    }
    return max_len;
}

// Hamming predicate of reference distances.h:8-31 on packed keys: XOR the planes, OR them
// into one mismatch bit per symbol, popcount, early exit once the budget is exceeded.
// Keys of different length never match (:16-20): with PAD padding that is "the PAD
// positions differ".
template <int K, int PW>
FQD_HD bool hamming_within(const Key<K, PW> &a, const Key<K, PW> &b, int max_distance,
                           bool varlen, uint32_t pad_code)
{
    int dist = 0;
#pragma unroll
    for (int i = 0; i < PW; i++) {
        uint32_t diff = 0;
#pragma unroll
        for (int p = 0; p < K; p++) diff |= a.w[p * PW + i] ^ b.w[p * PW + i];
        dist += popc32(diff);
        if (dist > max_distance) return false;
        if (varlen && (pad_mask_word(a, i, pad_code) != pad_mask_word(b, i, pad_code))) return false;
    }
    return true;
}

// Code of symbol `pos`.
template <int K, int PW>
FQD_HD uint32_t symbol_at(const Key<K, PW> &a, uint32_t pos)
{
    uint32_t c = 0;
    const uint32_t wi = pos >> 5, sh = pos & 31u;
#pragma unroll
    for (int p = 0; p < K; p++) {
        uint32_t word = 0;
#pragma unroll
        for (int i = 0; i < PW; i++) word = (wi == (uint32_t)i) ? a.w[p * PW + i] : word;
        c |= ((word >> sh) & 1u) << p;
    }
    return c;
}

// Strict "a < b" in Python str order (bytes lexicographic, proper prefix first).
template <int K, int PW>
FQD_HD bool key_less(const Key<K, PW> &a, const Key<K, PW> &b, const uint8_t *rank)
{
#pragma unroll
    for (int i = 0; i < PW; i++) {
        uint32_t diff = 0;
#pragma unroll
        for (int p = 0; p < K; p++) diff |= a.w[p * PW + i] ^ b.w[p * PW + i];
        if (diff) {
            const int sh = ctz32(diff);
            uint32_t ca = 0, cb = 0;
#pragma unroll
            for (int p = 0; p < K; p++) {
                ca |= ((a.w[p * PW + i] >> sh) & 1u) << p;
                cb |= ((b.w[p * PW + i] >> sh) & 1u) << p;
            }
            return rank[ca] < rank[cb];
        }
    }
    return false;
}

// (count, key) tuple order of the reference's sorted() calls.
template <int K, int PW>
FQD_HD bool prio_less(uint32_t ca, const Key<K, PW> &a, uint32_t cb, const Key<K, PW> &b,
                      const uint8_t *rank)
{
    if (ca != cb) return ca < cb;
    return key_less(a, b, rank);
}

// 32 plane bits starting at symbol position `pos` (bits past the end read as 0).
template <int K, int PW>
FQD_HD uint32_t plane_bits32(const Key<K, PW> &a, int p, uint32_t pos)
{
    const uint32_t wi = pos >> 5, sh = pos & 31u;
    uint32_t lo = 0, hi = 0;
#pragma unroll
    for (int i = 0; i < PW; i++) {
        lo = (wi == (uint32_t)i) ? a.w[p * PW + i] : lo;
        hi = (wi + 1 == (uint32_t)i) ? a.w[p * PW + i] : hi;
    }
    return sh ? ((lo >> sh) | (hi << (32u - sh))) : lo;
}

// Hash of the symbols [start, start+len) right-aligned, so that the same substring hashes
// equal wherever it sits in the key (needed by the shifted blocks of the Levenshtein
// pigeonhole).  `salt` separates passes / block ids / lengths.
template <int K, int PW>
FQD_HD uint64_t block_hash(const Key<K, PW> &a, uint32_t start, uint32_t len, uint64_t salt)
{
    uint32_t ha = 0x243F6A88u ^ (uint32_t)salt, hb = 0x85A308D3u + (uint32_t)(salt >> 32) * 0x9E3779B1u;
#pragma unroll
    for (int c = 0; c < PW; c++) {
        if ((uint32_t)(32 * c) < len) {
            const uint32_t rem = len - 32u * c;
            const uint32_t mask = rem >= 32u ? 0xFFFFFFFFu : ((1u << rem) - 1u);
#pragma unroll
            for (int p = 0; p < K; p++) hash_lanes(plane_bits32(a, p, start + 32u * c) & mask, c * K + p, ha, hb);
        }
    }
    return ((uint64_t)fmix32(ha ^ rotl32(hb, 16)) << 32) | fmix32(hb + 0x9E3779B9u * ha);
}

// block_hash of the leading `len` symbols (start = 0: no shifting or word selection needed).
template <int K, int PW>
FQD_HD uint64_t block0_hash(const Key<K, PW> &a, uint32_t len, uint64_t salt)
{
    uint32_t ha = 0x243F6A88u ^ (uint32_t)salt, hb = 0x85A308D3u + (uint32_t)(salt >> 32) * 0x9E3779B1u;
#pragma unroll
    for (int c = 0; c < PW; c++) {
        if ((uint32_t)(32 * c) < len) {
            const uint32_t rem = len - 32u * c;
            const uint32_t mask = rem >= 32u ? 0xFFFFFFFFu : ((1u << rem) - 1u);
#pragma unroll
            for (int p = 0; p < K; p++) hash_lanes(a.w[p * PW + c] & mask, c * K + p, ha, hb);
        }
    }
    return ((uint64_t)fmix32(ha ^ rotl32(hb, 16)) << 32) | fmix32(hb + 0x9E3779B9u * ha);
}

// Pigeonhole block j of d+1 over a key of `len` symbols: [len*j/(d+1), len*(j+1)/(d+1)).
FQD_HD uint32_t block_start(uint32_t len, uint32_t j, uint32_t nblocks)
{
    return (len * j) / nblocks;   // len <= 320 symbols, j <= nblocks <= len: 32-bit arithmetic (a 64-bit divide costs ~100 instructions)
}

// 32 plane bits for the symbol positions pos .. pos+31 where pos may be negative (positions before the key read as 0).
template <int K, int PW>
FQD_HD uint32_t plane_bits32_signed(const Key<K, PW> &a, int p, int pos)
{
    if (pos >= 0) return plane_bits32(a, p, (uint32_t)pos);
    if (pos <= -32) return 0u;
    return a.w[p * PW] << (uint32_t)(-pos);
}

// bits k of a 32-bit word whose symbol position base + k lies in [0, len)
FQD_HD uint32_t range_mask32(int base, int len)
{
    const int lo = base < 0 ? -base : 0, hi = len - base;   // k in [lo, hi)
    if (hi <= lo || lo >= 32) return 0u;
    const uint32_t upto = hi >= 32 ? 0xFFFFFFFFu : ((1u << hi) - 1u);
    return upto & ~((1u << lo) - 1u);
}

// A cheap NECESSARY condition for "Levenshtein distance <= d" (shifted Hamming filter): in an alignment with at most
// d edits every symbol of `a` that is not substituted or deleted matches the symbol of `b` on one of the diagonals
// -d .. d, so the positions of `a` that mismatch on ALL those diagonals number at most d.  Bit-plane XORs under 2d+1
// shifts, AND, POPC: ~60 integer operations for a 24-symbol key at d = 2, against ~400 for the Myers verify it
// guards -- and random candidates of a pigeonhole bucket almost never pass it.  Never rejects a true neighbour.
template <int K, int PW>
FQD_HD bool shifted_hamming_maybe_within(const Key<K, PW> &a, uint32_t la, const Key<K, PW> &b, uint32_t lb, int d)
{
    int count = 0;
#pragma unroll
    for (int c = 0; c < PW; c++) {
        if ((uint32_t)(32 * c) >= la) break;
        uint32_t acc = 0xFFFFFFFFu;
        for (int s = -d; s <= d; s++) {
            uint32_t diff = 0;
#pragma unroll
            for (int p = 0; p < K; p++) diff |= a.w[p * PW + c] ^ plane_bits32_signed(b, p, 32 * c + s);
            acc &= diff | ~range_mask32(32 * c + s, (int)lb);   // no partner on this diagonal = a mismatch
        }
        acc &= range_mask32(32 * c, (int)la);
        count += popc32(acc);
        if (count > d) return false;
    }
    return true;
}

// ---------------------------------------------------------------------------------------
// Levenshtein predicate (reference distances.h:33-87 == "edit distance <= d") by Myers'
// bit-vector algorithm in Hyyro's global-distance form, NW64 64-bit words per column
// (one word for keys up to 64 symbols).  Pattern = a (length la), text = b (length lb).
// ---------------------------------------------------------------------------------------
template <int K, int PW>
FQD_HD bool myers_within(const Key<K, PW> &a, uint32_t la, const Key<K, PW> &b, uint32_t lb,
                         int max_distance)
{
    constexpr int NW64 = (PW + 1) / 2;
    const uint32_t diff = la > lb ? la - lb : lb - la;
    if (max_distance < 0 || diff > (uint32_t)max_distance) return false;
    if (la == 0) return lb <= (uint32_t)max_distance;
    // pattern planes as 64-bit words
    uint64_t P[K][NW64];
#pragma unroll
    for (int p = 0; p < K; p++)
#pragma unroll
        for (int i = 0; i < NW64; i++) {
            uint64_t lo = a.w[p * PW + 2 * i];
            uint64_t hi = (2 * i + 1 < PW) ? a.w[p * PW + 2 * i + 1] : 0u;
            P[p][i] = lo | (hi << 32);
        }
    uint64_t valid[NW64];   // bits < la
#pragma unroll
    for (int i = 0; i < NW64; i++) {
        const uint32_t lo = 64u * i;
        valid[i] = la >= lo + 64u ? ~0ULL : (la > lo ? ((1ULL << (la - lo)) - 1ULL) : 0ULL);
    }
    uint64_t Pv[NW64], Mv[NW64];
#pragma unroll
    for (int i = 0; i < NW64; i++) { Pv[i] = valid[i]; Mv[i] = 0; }
    int score = (int)la;
    const uint32_t top_w = (la - 1) >> 6;
    const uint64_t top_bit = 1ULL << ((la - 1) & 63u);
    for (uint32_t j = 0; j < lb; j++) {
        const uint32_t c = symbol_at(b, j);
        uint64_t carry_add = 0;      // carry of (Eq & Pv) + Pv across words
        uint64_t ph_in = 1;          // global distance: D[0][j] - D[0][j-1] = +1
        uint64_t mh_in = 0;
#pragma unroll
        for (int i = 0; i < NW64; i++) {
            uint64_t Eq = valid[i];
#pragma unroll
            for (int p = 0; p < K; p++) Eq &= ((c >> p) & 1u) ? P[p][i] : ~P[p][i];
            const uint64_t Xv = Eq | Mv[i];
            const uint64_t x = Eq & Pv[i];
            const uint64_t s1 = x + Pv[i];
            const uint64_t c1 = s1 < x;
            const uint64_t sum = s1 + carry_add;
            const uint64_t c2 = sum < s1;
            carry_add = c1 | c2;
            const uint64_t Xh = (sum ^ Pv[i]) | Eq;
            uint64_t Ph = Mv[i] | ~(Xh | Pv[i]);
            uint64_t Mh = Pv[i] & Xh;
            if ((uint32_t)i == top_w) {
                if (Ph & top_bit) score++;
                else if (Mh & top_bit) score--;
            }
            const uint64_t ph_out = Ph >> 63, mh_out = Mh >> 63;
            Ph = (Ph << 1) | ph_in;
            Mh = (Mh << 1) | mh_in;
            ph_in = ph_out; mh_in = mh_out;
            Pv[i] = (Mh | ~(Xv | Ph)) & valid[i];
            Mv[i] = Ph & Xv & valid[i];
        }
        // the final score can still drop by at most one per remaining text symbol
        if (score - (int)(lb - 1 - j) > max_distance) return false;
    }
    return score <= max_distance;
}

// ---------------------------------------------------------------------------------------
// The same predicate for keys of up to 63 symbols and small d by furthest-reaching diagonals (Landau-Vishkin /
// Ukkonen), entirely in registers and without a data-dependent loop: a GPU warp runs it in lockstep, where the
// column loop of Myers' algorithm with its early exit leaves most lanes of a warp idle (config 4, d = 2: a quarter
// of the random candidates of a bucket pass the shifted-Hamming filter, and the verify behind it took 23 of 24 ms).
//
// D_k (bit i) = a[i] differs from b[i + k], or one of the two does not exist.  fr[e][k] = the furthest row i such
// that a[0, i) and b[0, i + k) are within e edits; a diagonal is extended over a run of matches by counting the
// trailing zeros of D_k >> i.  fr[e][k] = extend(max(fr[e-1][k] + 1, fr[e-1][k-1], fr[e-1][k+1] + 1)), and the
// distance is <= d iff some fr[e <= d][lb - la] reaches la.  2d+1 masks and (d+1)^2 extensions: ~200 integer
// operations for d = 2.
// ---------------------------------------------------------------------------------------
template <typename W, int K, int PW>
FQD_HD W plane_word(const Key<K, PW> &a, int p)
{
    static_assert(PW <= 2, "plane_word: keys of up to 64 symbols");
    if constexpr (sizeof(W) == 4) {
        return a.w[p * PW];                      // (callers: every symbol lies in the first word)
    } else {
        uint64_t v = a.w[p * PW];
        if constexpr (PW == 2) v |= (uint64_t)a.w[p * PW + 1] << 32;
        return v;
    }
}

#if defined(__CUDA_ARCH__)
FQD_HD int ctz_word(uint64_t x) { return __ffsll((long long)x) - 1; }   // x != 0
FQD_HD int ctz_word(uint32_t x) { return __ffs((int)x) - 1; }
#else
FQD_HD int ctz_word(uint64_t x) { return __builtin_ctzll(x); }
FQD_HD int ctz_word(uint32_t x) { return __builtin_ctz(x); }
#endif

constexpr int LV_MAX_D = 3;

// W = uint32_t for keys of up to 31 symbols (half the instructions of the 64-bit form), uint64_t up to 63
template <typename W, int K, int PW, int D>
FQD_HD bool lv_within_d(const Key<K, PW> &a, uint32_t la, const Key<K, PW> &b, uint32_t lb)
{
    // la, lb < bits of W; |la - lb| <= D checked by the caller
    constexpr W ONE = 1, TOP = ONE << (sizeof(W) * 8 - 1);
    const W ma = (ONE << la) - ONE, mb = (ONE << lb) - ONE;
    W A[K], B[K];
#pragma unroll
    for (int p = 0; p < K; p++) { A[p] = plane_word<W, K, PW>(a, p); B[p] = plane_word<W, K, PW>(b, p); }
    W Dk[2 * D + 1];   // Dk[k + D]
#pragma unroll
    for (int k = -D; k <= D; k++) {
        W diff = 0;
#pragma unroll
        for (int p = 0; p < K; p++) diff |= A[p] ^ (k >= 0 ? (W)(B[p] >> k) : (W)(B[p] << -k));
        const W have_b = k >= 0 ? (W)(mb >> k) : (W)(mb << -k);   // rows i with 0 <= i + k < lb
        Dk[k + D] = diff | ~have_b | ~ma | TOP;                     // top bit: a run always ends
    }
    const int target = (int)lb - (int)la + D;                      // index of the diagonal the alignment must end on
    int fr[2 * D + 1];
#pragma unroll
    for (int k = 0; k <= 2 * D; k++) fr[k] = -64;                  // unreachable
    fr[D] = ctz_word(Dk[D]);
    bool ok = target == D && fr[D] >= (int)la;
#pragma unroll
    for (int e = 1; e <= D; e++) {
        int nx[2 * D + 1];
#pragma unroll
        for (int k = 0; k <= 2 * D; k++) nx[k] = fr[k];
#pragma unroll
        for (int k = D - e; k <= D + e; k++) {
            int t = fr[k] + 1;                                      // substitution
            if (k > 0 && fr[k - 1] > t) t = fr[k - 1];              // a symbol of b inserted
            if (k < 2 * D && fr[k + 1] + 1 > t) t = fr[k + 1] + 1;  // a symbol of a deleted
            // rows beyond either string do not exist: row <= la, column row + (k - D) <= lb
            const int lim_b = (int)lb - (k - D), lim = (int)la < lim_b ? (int)la : lim_b;
            if (t > lim) t = lim;
            if (t >= 0) t += ctz_word((W)(Dk[k] >> t));             // extend over the run of matches
            else t = -64;
            nx[k] = t;
        }
#pragma unroll
        for (int k = 0; k <= 2 * D; k++) fr[k] = nx[k];
#pragma unroll
        for (int k = D - e; k <= D + e; k++) ok = ok || (k == target && fr[k] >= (int)la);
    }
    return ok;
}

template <typename W, int K, int PW>
FQD_HD bool lv_within(const Key<K, PW> &a, uint32_t la, const Key<K, PW> &b, uint32_t lb, int max_distance)
{
    if (max_distance == 1) return lv_within_d<W, K, PW, 1>(a, la, b, lb);
    if (max_distance == 2) return lv_within_d<W, K, PW, 2>(a, la, b, lb);
    return lv_within_d<W, K, PW, 3>(a, la, b, lb);
}

// edit distance <= max_distance; the fast form when the keys fit (else Myers)
template <int K, int PW>
FQD_HD bool edit_within(const Key<K, PW> &a, uint32_t la, const Key<K, PW> &b, uint32_t lb, int max_distance)
{
    if constexpr (PW <= 2) {
        if (la <= 63u && lb <= 63u && la > 0 && lb > 0 && max_distance >= 1 && max_distance <= LV_MAX_D) {
            const uint32_t diff = la > lb ? la - lb : lb - la;
            if (diff > (uint32_t)max_distance) return false;
            if (la <= 31u && lb <= 31u) return lv_within<uint32_t, K, PW>(a, la, b, lb, max_distance);
            return lv_within<uint64_t, K, PW>(a, la, b, lb, max_distance);
        }
    }
    return myers_within<K, PW>(a, la, b, lb, max_distance);
}

}  // namespace fqd
