// key_probe.cpp -- host build of fastqdedup_b200/csrc/key.cuh for the CPU test-suite.
// The device kernels use exactly these inline functions; compiling them for the host lets
// tests/test_key_primitives.py check packing, Hamming, Myers, ordering and block hashing
// against the oracle where no GPU exists.  Test scaffolding only, never shipped.
#include <string.h>

#include <algorithm>
#include <vector>

#include "key.cuh"

using namespace fqd;

namespace {

struct HostCodec {
    Codec c;
    HostCodec(const uint8_t *alphabet, int n, bool varlen, int bits)
    {
        memset(c.lut, 0xFF, sizeof c.lut);
        memset(c.rank, 0, sizeof c.rank);
        std::vector<uint8_t> sorted(alphabet, alphabet + n);
        std::sort(sorted.begin(), sorted.end());
        for (int i = 0; i < n; i++) {
            c.lut[alphabet[i]] = (uint8_t)i;
            c.rank[i] = (uint8_t)(1 + (std::lower_bound(sorted.begin(), sorted.end(), alphabet[i]) - sorted.begin()));
        }
        c.pad_code = (uint8_t)n;
        bool dna = bits == 3;
        for (int i = 0; i < n; i++) dna = dna && strchr("ACGTN", alphabet[i]) != nullptr;
        if (dna) {
            memset(c.lut, 0xFF, sizeof c.lut);
            memset(c.rank, 0, sizeof c.rank);
            const char *letters = "ACGNT";
            for (int r = 0; r < 5; r++) {
                const uint8_t code = ((uint8_t)letters[r] >> 1) & 7u;
                c.lut[(uint8_t)letters[r]] = code;
                c.rank[code] = (uint8_t)(r + 1);
            }
            c.pad_code = SWAR_PAD_CODE;
        }
        c.bits = (uint8_t)bits;
        c.n_symbols = (uint8_t)n;
        c.varlen = varlen;
    }
};

template <int K, int PW>
int run(int op, const HostCodec &hc, const uint8_t *a, uint32_t la, const uint8_t *b, uint32_t lb,
        uint32_t max_len, int d, uint32_t p0, uint32_t p1, uint32_t p2, uint64_t *out64)
{
    Key<K, PW> ka, kb;
    uint32_t bad = 0;
    const bool varlen = hc.c.varlen;
    if (!pack_key<K, PW>(a, la, varlen ? max_len : la, hc.c.lut, hc.c.pad_code, ka, &bad)) return -2;
    if (!pack_key<K, PW>(b, lb, varlen ? max_len : lb, hc.c.lut, hc.c.pad_code, kb, &bad)) return -2;
    switch (op) {
    case 0: return hamming_within<K, PW>(ka, kb, d, varlen, hc.c.pad_code) ? 1 : 0;
    case 1: return myers_within<K, PW>(ka, la, kb, lb, d) ? 1 : 0;
    case 2: return key_less<K, PW>(ka, kb, hc.c.rank) ? 1 : 0;
    case 3: return key_equal<K, PW>(ka, kb) ? 1 : 0;
    case 4: return (int)key_length<K, PW>(ka, hc.c.pad_code, max_len);
    case 5:   // block hash equality: a[p0:p0+p2] vs b[p1:p1+p2]
        return block_hash<K, PW>(ka, p0, p2, 7) == block_hash<K, PW>(kb, p1, p2, 7) ? 1 : 0;
    case 6: *out64 = hash_key<K, PW>(ka); return 0;
    case 7: return (int)symbol_at<K, PW>(ka, p0);
    case 8:   // table-free ACGTN packing == table packing (or both reject)
        if constexpr (K == 3) {
            alignas(4) uint8_t buf[4 * 8 * PW + 8] = {};
            memcpy(buf, a, la);
            for (uint32_t i = la; i < sizeof buf; i++) buf[i] = (uint8_t)(0x5A + i);   // garbage tail
            Key<K, PW> ks;
            const bool ok = pack_key_acgtn<PW>(reinterpret_cast<const uint32_t *>(buf), la, varlen ? max_len : la, ks);
            return ok ? (key_equal<K, PW>(ks, ka) ? 1 : 0) : 2;
        }
        return -1;
    case 9:   // straight-line packer for keys of exactly 4*NW symbols == table packing (or both reject)
        if constexpr (K == 3) {
            alignas(4) uint8_t buf[4 * 8 * PW + 8] = {};
            memcpy(buf, a, la);
            Key<K, PW> ks;
            bool ok = false, known = true;
            const uint32_t *w = reinterpret_cast<const uint32_t *>(buf);
            if (la == 12) ok = pack_key_acgtn_fixed<PW, 3>(w, ks);
            else if (la == 24) ok = pack_key_acgtn_fixed<PW, 6>(w, ks);
            else if (la == 36) { if constexpr (PW >= 2) ok = pack_key_acgtn_fixed<PW, 9>(w, ks); else known = false; }
            else if (la == 48) { if constexpr (PW >= 2) ok = pack_key_acgtn_fixed<PW, 12>(w, ks); else known = false; }
            else known = false;
            if (!known) return -1;
            return ok ? (key_equal<K, PW>(ks, ka) ? 1 : 0) : 2;
        }
        return -1;
    case 10:  // the partition hash of the leading block is a function of that block (and the salt) only
        return block0_hash<K, PW>(ka, p2, 99) == block0_hash<K, PW>(kb, p2, 99) ? 1 : 0;
    case 11: *out64 = hash_key32<K, PW>(ka); return 0;
    case 12: return shifted_hamming_maybe_within<K, PW>(ka, la, kb, lb, d) ? 1 : 0;
    case 13: return edit_within<K, PW>(ka, la, kb, lb, d) ? 1 : 0;
    }
    return -1;
}

}  // namespace

extern "C" int key_probe(int K, int PW, int op, const uint8_t *alphabet, int n_alpha, int varlen,
                         const uint8_t *a, uint32_t la, const uint8_t *b, uint32_t lb,
                         uint32_t max_len, int d, uint32_t p0, uint32_t p1, uint32_t p2,
                         uint64_t *out64)
{
    HostCodec hc(alphabet, n_alpha, varlen != 0, K);
#define CASE(K_, PW_) if (K == K_ && PW == PW_) return run<K_, PW_>(op, hc, a, la, b, lb, max_len, d, p0, p1, p2, out64);
    CASE(3, 1) CASE(3, 2) CASE(3, 3) CASE(3, 5) CASE(4, 2) CASE(8, 1) CASE(8, 2)
#undef CASE
    return -3;
}
