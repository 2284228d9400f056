// host_pack.cpp -- packing ACGTN keys on the host, so that the host -> device copy of a job carries 3 bits per
// symbol instead of 8 ("pack keys in C straight from the parser's buffers").
//
// With the clustering at a few milliseconds, a job that starts from host buffers is bound by the PCIe copy of its
// keys (100 M x 36 bytes: 65 of 71 ms).  Two forms, both with the code of key.cuh (bit p+1 of every ASCII byte is
// code bit p: A 0, C 1, T 2, G 3, N 7):
//   * rows (fqd_pack_keys, the public utility): per key its three code-bit planes back to back -- plane p in bits
//     [p*L, (p+1)*L) -- in ceil(3L / 32) 32-bit words, 16 bytes for 36 symbols;
//   * plane streams (pack_planes_parallel, what fqd_cluster sends for HOST jobs): the rows of a chunk as one stream of
//     symbols, three bit streams per chunk; partition_planes_kernel (partitioned.cuh) cuts a row out with funnel shifts.
//
// The row packer is one AVX-512 step per row (masked 64-byte load, one VPTESTMB per plane gives the plane as a mask
// register, one VPERMB table look-up validates all bytes at once), the stream packer the same per 64 symbols with no
// per-row work; scalar fallbacks; a small pool of threads over a chunk.  No CUDA here.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include "../../include/fqd_b200.h"

namespace fqd {

void set_error(const char *fmt, ...);

namespace {

inline void put_bits(uint64_t (&out)[4], uint64_t v, uint32_t at)
{
    const uint32_t w = at >> 6, sh = at & 63u;
    out[w] |= v << sh;
    if (sh) out[w + 1] |= v >> (64u - sh);
}

// scalar: returns false at the first byte that is not one of ACGTN
inline bool pack_row_scalar(const uint8_t *row, uint32_t L, uint32_t row_words, uint32_t *dst)
{
    uint64_t plane[3] = {0, 0, 0};
    for (uint32_t i = 0; i < L; i++) {
        const uint8_t c = row[i];
        if (c != 'A' && c != 'C' && c != 'G' && c != 'T' && c != 'N') return false;
        const uint64_t code = (c >> 1) & 7u;
        plane[0] |= (code & 1u) << i;
        plane[1] |= ((code >> 1) & 1u) << i;
        plane[2] |= ((code >> 2) & 1u) << i;
    }
    uint64_t out[4] = {0, 0, 0, 0};
    put_bits(out, plane[0], 0);
    put_bits(out, plane[1], L);
    put_bits(out, plane[2], 2 * L);
    memcpy(dst, out, (size_t)row_words * 4);
    return true;
}

#if defined(__x86_64__)
#define FQD_AVX512 __attribute__((target("avx512f,avx512bw,avx512vbmi,bmi2")))

// one row: planes as mask registers, validity of every byte by one table look-up
template <uint32_t L>
FQD_AVX512 inline __mmask64 row_planes(const uint8_t *row, __m512i table, __m512i b1, __m512i b2, __m512i b3, uint64_t &p0, uint64_t &p1,
                                       uint64_t &p2)
{
    constexpr __mmask64 lanes = L >= 64 ? ~0ull : ((1ull << L) - 1ull);
    const __m512i v = _mm512_maskz_loadu_epi8(lanes, row);
    p0 = _mm512_test_epi8_mask(v, b1);
    p1 = _mm512_test_epi8_mask(v, b2);
    p2 = _mm512_test_epi8_mask(v, b3);
    return _mm512_mask_cmpneq_epi8_mask(lanes, _mm512_permutexvar_epi8(v, table), v);   // bytes that are not ACGTN
}

// the three planes of one row -> its packed words (straight-line code for the key lengths of the configs)
template <uint32_t L>
FQD_AVX512 inline void store_row(uint32_t *dst, uint64_t p0, uint64_t p1, uint64_t p2)
{
    constexpr uint32_t RW = (3u * L + 31u) / 32u;
    uint64_t out[4] = {0, 0, 0, 0};
    put_bits(out, p0, 0);
    put_bits(out, p1, L);
    put_bits(out, p2, 2 * L);
    if constexpr (RW % 2 == 0) {
#pragma GCC unroll 4
        for (uint32_t i = 0; i < RW / 2; i++) memcpy(dst + 2 * i, &out[i], 8);
    } else {
#pragma GCC unroll 4
        for (uint32_t i = 0; i < RW / 2; i++) memcpy(dst + 2 * i, &out[i], 8);
        const uint32_t last = (uint32_t)out[RW / 2];
        memcpy(dst + RW - 1, &last, 4);
    }
}

template <uint32_t L>
FQD_AVX512 size_t pack_rows_avx512_fixed(const uint8_t *src, size_t n, uint32_t stride, uint32_t *dst)
{
    constexpr uint32_t RW = (3u * L + 31u) / 32u;
    // table: index = byte & 63; the five letters map to themselves, every other entry to a value with other low bits
    // than its index (so it never equals a byte that looks it up -- 0xFF as the filler would accept the byte 0xFF)
    alignas(64) uint8_t lut[64];
    for (int i = 0; i < 64; i++) lut[i] = (uint8_t)(i ^ 1);   // never equal to a byte that indexes it (its low 6 bits are i)
    lut['A' & 63] = 'A'; lut['C' & 63] = 'C'; lut['G' & 63] = 'G'; lut['T' & 63] = 'T'; lut['N' & 63] = 'N';
    const __m512i table = _mm512_load_si512(lut);
    const __m512i b1 = _mm512_set1_epi8(0x02), b2 = _mm512_set1_epi8(0x04), b3 = _mm512_set1_epi8(0x08);
    constexpr size_t BLOCK = 64;   // rows validated together (the first bad row is only looked for when there is one)
    for (size_t r0 = 0; r0 < n; r0 += BLOCK) {
        const size_t r1 = r0 + BLOCK < n ? r0 + BLOCK : n;
        __mmask64 bad = 0;
        size_t r = r0;
        for (; r + 2 <= r1; r += 2) {   // two rows per step: their loads and mask moves overlap
            uint64_t a0, a1, a2, c0, c1, c2;
            bad |= row_planes<L>(src + r * stride, table, b1, b2, b3, a0, a1, a2);
            bad |= row_planes<L>(src + (r + 1) * stride, table, b1, b2, b3, c0, c1, c2);
            store_row<L>(dst + r * RW, a0, a1, a2);
            store_row<L>(dst + (r + 1) * RW, c0, c1, c2);
        }
        for (; r < r1; r++) {
            uint64_t a0, a1, a2;
            bad |= row_planes<L>(src + r * stride, table, b1, b2, b3, a0, a1, a2);
            store_row<L>(dst + r * RW, a0, a1, a2);
        }
        if (bad) {
            for (r = r0; r < r1; r++) {
                uint64_t a0, a1, a2;
                if (row_planes<L>(src + r * stride, table, b1, b2, b3, a0, a1, a2)) return r;
            }
        }
    }
    return n;
}

FQD_AVX512 size_t pack_rows_avx512(const uint8_t *src, size_t n, uint32_t L, uint32_t stride, uint32_t *dst, uint32_t row_words)
{
    switch (L) {
    case 12: return pack_rows_avx512_fixed<12>(src, n, stride, dst);
    case 24: return pack_rows_avx512_fixed<24>(src, n, stride, dst);
    case 36: return pack_rows_avx512_fixed<36>(src, n, stride, dst);
    case 48: return pack_rows_avx512_fixed<48>(src, n, stride, dst);
    default: break;
    }
    alignas(64) uint8_t lut[64];
    for (int i = 0; i < 64; i++) lut[i] = (uint8_t)(i ^ 1);   // never equal to a byte that indexes it (its low 6 bits are i)
    lut['A' & 63] = 'A'; lut['C' & 63] = 'C'; lut['G' & 63] = 'G'; lut['T' & 63] = 'T'; lut['N' & 63] = 'N';
    const __m512i table = _mm512_load_si512(lut);
    const __m512i b1 = _mm512_set1_epi8(0x02), b2 = _mm512_set1_epi8(0x04), b3 = _mm512_set1_epi8(0x08);
    const __mmask64 lanes = L >= 64 ? ~0ull : ((1ull << L) - 1ull);
    for (size_t r = 0; r < n; r++) {
        const __m512i v = _mm512_maskz_loadu_epi8(lanes, src + r * stride);
        const __mmask64 ok = _mm512_mask_cmpeq_epi8_mask(lanes, _mm512_permutexvar_epi8(v, table), v);
        if (ok != lanes) return r;
        const uint64_t p0 = _mm512_test_epi8_mask(v, b1), p1 = _mm512_test_epi8_mask(v, b2), p2 = _mm512_test_epi8_mask(v, b3);
        uint64_t out[4] = {0, 0, 0, 0};
        put_bits(out, p0, 0);
        put_bits(out, p1, L);
        put_bits(out, p2, 2 * L);
        memcpy(dst + r * row_words, out, (size_t)row_words * 4);
    }
    return n;
}
#endif

size_t pack_rows(const uint8_t *src, size_t n, uint32_t L, uint32_t stride, uint32_t *dst, uint32_t row_words)
{
#if defined(__x86_64__)
    static const bool fast = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") &&
                             __builtin_cpu_supports("avx512vbmi");
    if (fast) return pack_rows_avx512(src, n, L, stride, dst, row_words);
#endif
    for (size_t r = 0; r < n; r++)
        if (!pack_row_scalar(src + r * stride, L, row_words, dst + r * row_words)) return r;
    return n;
}

// A few persistent worker threads: a parallel-for over row ranges.
class Pool {
public:
    explicit Pool(int threads)
    {
        for (int i = 0; i < threads; i++) workers_.emplace_back([this] { loop(); });
    }
    ~Pool()
    {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
        }
        wake_.notify_all();
        for (auto &t : workers_) t.join();
    }
    int size() const { return (int)workers_.size(); }
    // runs fn(part) for part in [0, parts) on the workers and the caller; returns when all are done
    void run(int parts, const std::function<void(int)> &fn)
    {
        {
            std::lock_guard<std::mutex> lk(mu_);
            fn_ = &fn;
            next_ = 0;
            parts_ = parts;
            pending_ = parts;
            generation_++;
        }
        wake_.notify_all();
        work();
        std::unique_lock<std::mutex> lk(mu_);
        done_.wait(lk, [&] { return pending_ == 0; });
        fn_ = nullptr;
    }

private:
    void work()
    {
        for (;;) {
            int part;
            const std::function<void(int)> *fn;
            {
                std::lock_guard<std::mutex> lk(mu_);
                if (!fn_ || next_ >= parts_) return;
                part = next_++;
                fn = fn_;
            }
            (*fn)(part);
            std::lock_guard<std::mutex> lk(mu_);
            if (--pending_ == 0) done_.notify_all();
        }
    }
    void loop()
    {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(mu_);
                wake_.wait(lk, [&] { return stop_ || generation_ != seen; });
                if (stop_) return;
                seen = generation_;
            }
            work();
        }
    }
    std::mutex mu_;
    std::condition_variable wake_, done_;
    std::vector<std::thread> workers_;
    const std::function<void(int)> *fn_ = nullptr;
    int next_ = 0, parts_ = 0, pending_ = 0;
    uint64_t generation_ = 0;
    bool stop_ = false;
};

Pool &pool()
{
    static Pool p([] {
        const unsigned hc = std::thread::hardware_concurrency();
        const char *e = getenv("FQD_PACK_THREADS");
        const int want = e && *e ? atoi(e) : (int)(hc ? hc : 4u);
        return std::max(1, std::min(want, 64)) - 1;   // the calling thread works too
    }());
    return p;
}

}  // namespace

// ---- plane streams of a whole chunk (what fqd_cluster sends over PCIe for HOST jobs) -------------------------------
//
// The row format above costs ~6 ns per 36-byte row and thread, most of it the per-row bit shuffling; 16 threads pack
// 100 M keys in 54 ms, which made the PACKER the slower leg of a HOST job (the packed rows cross PCIe in 29 ms).  For
// the library's own use the rows of a chunk are therefore treated as one stream of bytes (fixed-length rows back to
// back): bit t of plane p = bit p+1 of byte t of the chunk.  One 64-byte load, three VPTESTMB and three 8-byte stores
// per 64 symbols, validity by one VPERMB look-up -- no per-row work at all on the host, and 3 bits per symbol
// exactly (13.5 bytes per 36-nt key instead of 16).  The GPU cuts a row's L bits out of each stream with a funnel
// shift (partition_planes_kernel).

namespace {

constexpr size_t NO_BAD = ~(size_t)0;

// vectors [v0, v1) of the stream of nbytes bytes; returns the index of the first byte outside ACGTN in them, or NO_BAD
size_t planes_range_scalar(const uint8_t *src, size_t nbytes, size_t v0, size_t v1, uint64_t *p0, uint64_t *p1, uint64_t *p2)
{
    for (size_t v = v0; v < v1; v++) {
        uint64_t m0 = 0, m1 = 0, m2 = 0;
        const size_t lo = v * 64, hi = lo + 64 < nbytes ? lo + 64 : nbytes;
        for (size_t t = lo; t < hi; t++) {
            const uint8_t c = src[t];
            if (c != 'A' && c != 'C' && c != 'G' && c != 'T' && c != 'N') return t;
            m0 |= (uint64_t)((c >> 1) & 1u) << (t - lo);
            m1 |= (uint64_t)((c >> 2) & 1u) << (t - lo);
            m2 |= (uint64_t)((c >> 3) & 1u) << (t - lo);
        }
        p0[v] = m0; p1[v] = m1; p2[v] = m2;
    }
    return NO_BAD;
}

#if defined(__x86_64__)
FQD_AVX512 size_t planes_range_avx512(const uint8_t *src, size_t nbytes, size_t v0, size_t v1, uint64_t *p0, uint64_t *p1, uint64_t *p2)
{
    alignas(64) uint8_t lut[64];
    for (int i = 0; i < 64; i++) lut[i] = (uint8_t)(i ^ 1);   // never equal to a byte that indexes it (its low 6 bits are i)
    lut['A' & 63] = 'A'; lut['C' & 63] = 'C'; lut['G' & 63] = 'G'; lut['T' & 63] = 'T'; lut['N' & 63] = 'N';
    const __m512i table = _mm512_load_si512(lut);
    const __m512i b1 = _mm512_set1_epi8(0x02), b2 = _mm512_set1_epi8(0x04), b3 = _mm512_set1_epi8(0x08);
    const size_t full = nbytes / 64;   // vectors below this index are complete
    size_t v = v0;
    const size_t vfull = v1 < full ? v1 : full;
    for (; v + 4 <= vfull; v += 4) {   // four independent vectors per step
        __m512i x[4];
        __mmask64 bad = 0;
#pragma GCC unroll 4
        for (int k = 0; k < 4; k++) x[k] = _mm512_loadu_si512(src + (v + k) * 64);
#pragma GCC unroll 4
        for (int k = 0; k < 4; k++) {
            p0[v + k] = _mm512_test_epi8_mask(x[k], b1);
            p1[v + k] = _mm512_test_epi8_mask(x[k], b2);
            p2[v + k] = _mm512_test_epi8_mask(x[k], b3);
            bad |= _mm512_cmpneq_epi8_mask(_mm512_permutexvar_epi8(x[k], table), x[k]);
        }
        if (bad) return planes_range_scalar(src, nbytes, v, v + 4, p0, p1, p2);   // (names the byte)
    }
    for (; v < v1; v++) {
        const size_t off = v * 64;
        const __mmask64 lanes = off + 64 <= nbytes ? ~0ull : ((1ull << (nbytes - off)) - 1ull);
        const __m512i xv = _mm512_maskz_loadu_epi8(lanes, src + off);
        p0[v] = _mm512_test_epi8_mask(xv, b1);
        p1[v] = _mm512_test_epi8_mask(xv, b2);
        p2[v] = _mm512_test_epi8_mask(xv, b3);
        if (_mm512_mask_cmpneq_epi8_mask(lanes, _mm512_permutexvar_epi8(xv, table), xv))
            return planes_range_scalar(src, nbytes, v, v + 1, p0, p1, p2);
    }
    return NO_BAD;
}
#endif

size_t planes_range(const uint8_t *src, size_t nbytes, size_t v0, size_t v1, uint64_t *p0, uint64_t *p1, uint64_t *p2)
{
#if defined(__x86_64__)
    static const bool fast = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") &&
                             __builtin_cpu_supports("avx512vbmi");
    if (fast) return planes_range_avx512(src, nbytes, v0, v1, p0, p1, p2);
#endif
    return planes_range_scalar(src, nbytes, v0, v1, p0, p1, p2);
}

}  // namespace

// 64-bit words one plane stream of n rows of L symbols takes, including the zero word behind it that lets the GPU
// read "this word and the next" for the last row
uint64_t plane_stream_words(uint64_t n, uint32_t L) { return (n * L + 63u) / 64u + 1u; }

// rows [0, n) of `src` (L bytes each, back to back) -> three plane streams of plane_stream_words(n, L) words at dst,
// dst + words, dst + 2 * words; returns the index of the first row holding a byte outside ACGTN, or n
uint64_t pack_planes_parallel(const uint8_t *src, uint64_t n, uint32_t L, uint64_t *dst)
{
    const uint64_t words = plane_stream_words(n, L);
    const size_t nbytes = (size_t)(n * L), nv = (nbytes + 63) / 64;
    uint64_t *p0 = dst, *p1 = dst + words, *p2 = dst + 2 * words;
    for (uint64_t w = nv; w < words; w++) p0[w] = p1[w] = p2[w] = 0;
    Pool &p = pool();
    const int parts = (int)std::min<uint64_t>((uint64_t)(p.size() + 1) * 4, std::max<uint64_t>(1, nv / 1024));
    std::atomic<uint64_t> bad{(uint64_t)nbytes};
    const std::function<void(int)> fn = [&](int part) {
        const size_t v0 = nv * (size_t)part / parts, v1 = nv * (size_t)(part + 1) / parts;
        const size_t b = planes_range(src, nbytes, v0, v1, p0, p1, p2);
        if (b != NO_BAD) {
            uint64_t cur = bad.load();
            while (b < cur && !bad.compare_exchange_weak(cur, b)) {}
        }
    };
    p.run(parts, fn);
    const uint64_t b = bad.load();
    return b == nbytes ? n : b / L;
}

uint32_t packed_row_words(uint32_t key_length) { return (3u * key_length + 31u) / 32u; }

// threads a pack_keys_parallel call runs on (the pool's workers + the caller)
int pack_threads() { return pool().size() + 1; }

// rows [0, n) of `src` (L bytes each, `stride` apart) -> packed rows; returns the index of the first row holding a byte
// outside ACGTN, or n
uint64_t pack_keys_parallel(const uint8_t *src, uint64_t n, uint32_t L, uint32_t stride, uint32_t *dst)
{
    const uint32_t rw = packed_row_words(L);
    Pool &p = pool();
    const int parts = (int)std::min<uint64_t>((uint64_t)(p.size() + 1) * 4, std::max<uint64_t>(1, n / 4096));
    std::atomic<uint64_t> bad{n};
    const std::function<void(int)> fn = [&](int part) {
        const uint64_t lo = n * (uint64_t)part / parts, hi = n * (uint64_t)(part + 1) / parts;
        const size_t r = pack_rows(src + lo * stride, (size_t)(hi - lo), L, stride, dst + lo * rw, rw);
        if (r != hi - lo) {
            uint64_t cur = bad.load(), mine = lo + r;
            while (mine < cur && !bad.compare_exchange_weak(cur, mine)) {}
        }
    };
    p.run(parts, fn);
    return bad.load();
}

}  // namespace fqd

extern "C" int fqd_pack_keys(const uint8_t *keys, uint64_t n_records, uint32_t key_length, uint32_t key_stride, uint32_t *packed,
                             uint64_t *bad_record)
{
    if ((!keys || !packed) && n_records) { fqd::set_error("fqd_pack_keys: null buffer"); return FQD_ERR_ARG; }
    if (key_length == 0 || key_length > 64 || key_length > key_stride) {
        fqd::set_error("fqd_pack_keys: key_length must be 1..64 and <= key_stride");
        return FQD_ERR_ARG;
    }
    const uint64_t bad = fqd::pack_keys_parallel(keys, n_records, key_length, key_stride, packed);
    if (bad_record) *bad_record = bad;
    if (bad != n_records) {
        fqd::set_error("fqd_pack_keys: record %llu holds a byte outside ACGTN (packed rows cannot carry it)", (unsigned long long)bad);
        return FQD_ERR_UNSUPPORTED;
    }
    return FQD_OK;
}
