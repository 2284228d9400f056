// fastq_native.cpp -- host side of the two FASTQ passes around the GPU job.
//
// The reference reads its inputs record by record through dnaio/xopen in Python (src/fastqdedup/__init__.py:54-57,
// 170-186), builds every key with Python string slicing + join (:160-167, :243-251), and writes the survivors through
// one single-threaded gzip stream (:189-206).  With the clustering itself down to milliseconds those loops ARE the
// run time, so they are native here:
//
//   pass 1  fqd_fastq_scan_open   one reader (+ inflate) thread and one parser thread per input file produce
//                                 batches of parsed records; a pool of workers takes batch k of every file, checks
//                                 that the records are mates (:181-185, same message) and writes the key -- the
//                                 concatenation of each file's sequence slice (Python slice semantics, negative
//                                 indices and steps included) -- and the quality slice of every record tuple straight
//                                 into the buffers fqd_cluster takes.  Stops at the shortest file, like zip().
//   pass 2  fqd_fastq_emit        the same readers; workers copy the record tuples whose bit is set in the keep
//                                 bitmap to per-file output blocks ("@name\nseq\n+\nqual\n", what dnaio's
//                                 fastq_bytes() writes) and, for .gz outputs, compress each block at level 1 as its own
//                                 gzip member; one writer appends the blocks in order.  The compression is off the
//                                 serial path; the decompressed bytes are what the reference writes.
//
// No CUDA here: plain C++17 + zlib + pthreads, linked into libfqd_b200.so.
#include <errno.h>
#include <fcntl.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/fqd_b200.h"

namespace fqd {
void set_error(const char *fmt, ...);
}

namespace {

constexpr size_t CHUNK_BYTES = 8u << 20;     // bytes a reader hands to its parser at a time
constexpr uint32_t BATCH_RECORDS = 32768;    // records per batch (the unit of work of the pool)
constexpr size_t QUEUE_DEPTH = 6;

// ---- bounded queue ------------------------------------------------------------------------------------

template <typename T>
class Channel {
public:
    explicit Channel(size_t depth) : depth_(depth) {}
    // false: the channel was cancelled
    bool push(T &&v)
    {
        std::unique_lock<std::mutex> lk(mu_);
        not_full_.wait(lk, [&] { return q_.size() < depth_ || cancelled_; });
        if (cancelled_) return false;
        q_.push_back(std::move(v));
        not_empty_.notify_one();
        return true;
    }
    // false: closed and drained (or cancelled)
    bool pop(T &out)
    {
        std::unique_lock<std::mutex> lk(mu_);
        not_empty_.wait(lk, [&] { return !q_.empty() || closed_ || cancelled_; });
        if (cancelled_ || q_.empty()) return false;
        out = std::move(q_.front());
        q_.pop_front();
        not_full_.notify_one();
        return true;
    }
    void close()
    {
        std::lock_guard<std::mutex> lk(mu_);
        closed_ = true;
        not_empty_.notify_all();
    }
    void cancel()
    {
        std::lock_guard<std::mutex> lk(mu_);
        cancelled_ = true;
        not_empty_.notify_all();
        not_full_.notify_all();
    }

private:
    std::mutex mu_;
    std::condition_variable not_full_, not_empty_;
    std::deque<T> q_;
    size_t depth_;
    bool closed_ = false, cancelled_ = false;
};

// ---- byte source: plain file or gzip (concatenated members included) ---------------------------------------

class ByteSource {
public:
    ~ByteSource()
    {
        if (inflating_) inflateEnd(&z_);
        if (fd_ >= 0) ::close(fd_);
    }
    bool open(const std::string &path, std::string &err)
    {
        fd_ = ::open(path.c_str(), O_RDONLY);
        if (fd_ < 0) { err = "cannot open " + path + ": " + strerror(errno); return false; }
#ifdef POSIX_FADV_SEQUENTIAL
        posix_fadvise(fd_, 0, 0, POSIX_FADV_SEQUENTIAL);
#endif
        unsigned char magic[2];
        const ssize_t n = ::pread(fd_, magic, 2, 0);
        gz_ = n == 2 && magic[0] == 0x1f && magic[1] == 0x8b;
        if (gz_) {
            memset(&z_, 0, sizeof z_);
            if (inflateInit2(&z_, 15 + 32) != Z_OK) { err = "zlib: inflateInit2 failed"; return false; }
            inflating_ = true;
            in_.resize(1u << 20);
        }
        path_ = path;
        return true;
    }
    // reads up to `cap` bytes; 0 = end of data, -1 = error
    ssize_t read(char *dst, size_t cap, std::string &err)
    {
        if (!gz_) {
            size_t got = 0;
            while (got < cap) {
                const ssize_t n = ::read(fd_, dst + got, cap - got);
                if (n < 0) { if (errno == EINTR) continue; err = "read error on " + path_ + ": " + strerror(errno); return -1; }
                if (n == 0) break;
                got += (size_t)n;
            }
            return (ssize_t)got;
        }
        z_.next_out = reinterpret_cast<Bytef *>(dst);
        z_.avail_out = (uInt)std::min<size_t>(cap, 1u << 30);
        while (z_.avail_out > 0) {
            if (z_.avail_in == 0 && !file_eof_) {
                const ssize_t n = ::read(fd_, in_.data(), in_.size());
                if (n < 0) { if (errno == EINTR) continue; err = "read error on " + path_ + ": " + strerror(errno); return -1; }
                if (n == 0) file_eof_ = true;
                z_.next_in = reinterpret_cast<Bytef *>(in_.data());
                z_.avail_in = (uInt)n;
            }
            if (z_.avail_in == 0 && file_eof_) {
                if (!member_done_ && started_) { err = "gzip stream of " + path_ + " is truncated"; return -1; }
                break;
            }
            if (member_done_) {   // another gzip member follows
                inflateReset(&z_);
                member_done_ = false;
            }
            started_ = true;
            const int rc = inflate(&z_, Z_NO_FLUSH);
            if (rc == Z_STREAM_END) member_done_ = true;
            else if (rc != Z_OK && rc != Z_BUF_ERROR) { err = "gzip stream of " + path_ + " is corrupt"; return -1; }
            else if (rc == Z_BUF_ERROR && z_.avail_in == 0 && file_eof_) { err = "gzip stream of " + path_ + " is truncated"; return -1; }
        }
        return (ssize_t)(std::min<size_t>(cap, 1u << 30) - z_.avail_out);
    }

private:
    int fd_ = -1;
    bool gz_ = false, inflating_ = false, file_eof_ = false, member_done_ = false, started_ = false;
    z_stream z_;
    std::vector<char> in_;
    std::string path_;
};

// ---- parsed batches ---------------------------------------------------------------------------------

// ---- line ends ----------------------------------------------------------------------------------------
//
// The parser's whole job is to find four line ends per record.  memchr per line costs a library call for ~150 bytes;
// with AVX-512 the buffer is scanned 64 bytes at a time instead (one compare gives the newline positions of a block as
// a mask, the line ends are peeled off it bit by bit), which made the parser ~2x faster.  Without AVX-512BW: memchr.
#if defined(__x86_64__)
#define FQD_NL_AVX512 __attribute__((target("avx512f,avx512bw")))
#endif

class NewlineScanner {
public:
    NewlineScanner(const char *p, size_t n) : p_(p), n_(n)
    {
#if defined(__x86_64__)
        static const bool fast = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw");
        fast_ = fast;
#endif
    }
    // position of the next newline at or after the scan position, or npos; the scan position moves behind it
    static constexpr size_t npos = ~(size_t)0;
    size_t next()
    {
#if defined(__x86_64__)
        if (fast_) return next_avx512();
#endif
        if (at_ >= n_) return npos;
        const char *e = static_cast<const char *>(memchr(p_ + at_, '\n', n_ - at_));
        if (!e) { at_ = n_; return npos; }
        at_ = (size_t)(e - p_) + 1;
        return at_ - 1;
    }

private:
#if defined(__x86_64__)
    FQD_NL_AVX512 size_t next_avx512()
    {
        while (mask_ == 0) {
            if (block_ >= n_) return npos;
            const size_t left = n_ - block_;
            const __m512i nl = _mm512_set1_epi8('\n');
            if (left >= 64) {
                mask_ = _mm512_cmpeq_epi8_mask(_mm512_loadu_si512(p_ + block_), nl);
            } else {
                const __mmask64 lanes = (1ull << left) - 1ull;
                mask_ = _mm512_mask_cmpeq_epi8_mask(lanes, _mm512_maskz_loadu_epi8(lanes, p_ + block_), nl);
            }
            base_ = block_;
            block_ += 64;
        }
        const size_t pos = base_ + (size_t)__builtin_ctzll(mask_);
        mask_ &= mask_ - 1;
        return pos;
    }
    uint64_t mask_ = 0;          // newlines of the block at base_ not handed out yet
    size_t block_ = 0, base_ = 0;
    bool fast_ = false;
#endif
    const char *p_;
    size_t n_;
    size_t at_ = 0;              // (memchr path) scan position
};

// A parsed record points into the reader's chunk buffers (no copy of the text: the parser only finds the line ends);
// a batch keeps the buffers its records point into alive.
struct Rec { const char *name, *seq, *qual; uint32_t name_len, seq_len, qual_len; };

struct Batch {
    std::vector<std::shared_ptr<char[]>> bufs;   // chunk buffers the records point into
    std::vector<Rec> recs;
    std::string error;           // set on the last batch of a failed file
};

// A reader's chunk: CHUNK_BYTES of file data behind CHUNK_HEADROOM free bytes (the parser puts the incomplete record
// left over from the previous chunk there instead of copying the whole chunk behind it) and one byte of tail room (the
// newline a last line may lack).  The buffer is not value-initialised: zero-filling 8 MB per chunk was a memory pass.
constexpr size_t CHUNK_HEADROOM = 64u << 10;

// Chunk buffers are recycled: a fresh 8 MB allocation is fresh pages, i.e. a page fault and a kernel zero-fill per
// 4 KB -- a hidden memory pass over every byte read.
class ChunkPool {
public:
    static constexpr size_t BYTES = CHUNK_HEADROOM + CHUNK_BYTES + 1;
    static std::shared_ptr<char[]> get()
    {
        char *raw = nullptr;
        {
            std::lock_guard<std::mutex> lk(mu());
            if (!free_list().empty()) { raw = free_list().back(); free_list().pop_back(); }
        }
        if (!raw) raw = new char[BYTES];
        return std::shared_ptr<char[]>(raw, [](char *q) {
            std::lock_guard<std::mutex> lk(mu());
            if (free_list().size() < 24) free_list().push_back(q);   // (at most ~200 MB stay with the process)
            else delete[] q;
        });
    }

private:
    struct FreeList {
        std::vector<char *> v;
        ~FreeList() { for (char *q : v) delete[] q; }
    };
    static std::mutex &mu() { static std::mutex m; return m; }
    static std::vector<char *> &free_list() { static FreeList f; return f.v; }
};

struct Chunk {
    std::shared_ptr<char[]> buf;
    size_t n = 0;                // bytes of file data at buf + CHUNK_HEADROOM
    bool last = false;
    std::string error;
    char *data() { return buf.get() + CHUNK_HEADROOM; }
};

// One input file: reader thread (read + inflate) -> parser thread (record boundaries) -> batches of BATCH_RECORDS.
class FileStream {
public:
    FileStream() : chunks_(QUEUE_DEPTH), batches_(QUEUE_DEPTH) {}
    ~FileStream() { stop(); }
    void start(const std::string &path)
    {
        path_ = path;
        reader_ = std::thread([this] { read_loop(); });
        parser_ = std::thread([this] { parse_loop(); });
    }
    // false: no more batches; err non-empty on failure
    bool next(Batch &b, std::string &err)
    {
        if (!batches_.pop(b)) return false;
        if (!b.error.empty()) { err = b.error; return false; }
        return true;
    }
    void stop()
    {
        chunks_.cancel();
        batches_.cancel();
        if (reader_.joinable()) reader_.join();
        if (parser_.joinable()) parser_.join();
    }

private:
    void read_loop()
    {
        ByteSource src;
        std::string err;
        if (!src.open(path_, err)) {
            Chunk c; c.last = true; c.error = err;
            chunks_.push(std::move(c));
            chunks_.close();
            return;
        }
        for (;;) {
            Chunk c;
            c.buf = ChunkPool::get();
            const ssize_t n = src.read(c.data(), CHUNK_BYTES, err);
            if (n < 0) { c.n = 0; c.last = true; c.error = err; chunks_.push(std::move(c)); break; }
            c.n = (size_t)n;
            c.last = n == 0;
            const bool last = c.last;
            if (!chunks_.push(std::move(c)) || last) break;
        }
        chunks_.close();
    }
    void fail(const std::string &msg)
    {
        Batch b;
        b.error = msg;
        batches_.push(std::move(b));
        batches_.close();
        chunks_.cancel();
    }
    void parse_loop()
    {
        std::vector<char> carry;   // an incomplete record at the end of the previous chunk
        Batch cur;
        cur.recs.reserve(BATCH_RECORDS);
        uint64_t lineno = 0;
        Chunk c;
        bool eof = false;
        while (!eof) {
            if (!chunks_.pop(c)) { batches_.close(); return; }
            if (!c.error.empty()) { fail(c.error); return; }
            eof = c.last;
            std::shared_ptr<char[]> owner = c.buf;   // the buffer p points into (one byte of tail room behind n)
            char *p;
            size_t n;
            if (carry.empty()) {
                p = c.data(); n = c.n;
            } else if (carry.size() <= CHUNK_HEADROOM) {
                p = c.data() - carry.size();   // the leftover record goes in front of the chunk's data, in place
                memcpy(p, carry.data(), carry.size());
                n = carry.size() + c.n;
                carry.clear();
            } else {   // (a record longer than the headroom: the slow way)
                n = carry.size() + c.n;
                owner.reset(new char[n + 1]);
                p = owner.get();
                memcpy(p, carry.data(), carry.size());
                memcpy(p + carry.size(), c.data(), c.n);
                carry.clear();
            }
            if (eof && n && p[n - 1] != '\n') p[n++] = '\n';   // a file whose last line lacks the newline
            bool held = false;   // the current batch already keeps `owner` alive
            size_t pos = 0;
            NewlineScanner lines(p, n);
            for (;;) {
                // four lines
                const char *l0 = p + pos;
                const size_t n0 = lines.next(), n1 = lines.next(), n2 = lines.next(), n3 = lines.next();
                if (n3 == NewlineScanner::npos) break;   // (npos is sticky: an incomplete record)
                const char *e0 = p + n0, *e1 = p + n1, *e2 = p + n2, *e3 = p + n3;
                if (l0[0] != '@' || e1[1] != '+') {
                    fail(path_ + ": malformed FASTQ record at line " + std::to_string(lineno + 1));
                    return;
                }
                auto strip = [](const char *b, const char *e) { return (e > b && e[-1] == '\r') ? e - 1 : e; };
                const char *name_b = l0 + 1, *name_e = strip(name_b, e0);
                const char *seq_b = e0 + 1, *seq_e = strip(seq_b, e1);
                const char *qual_b = e2 + 1, *qual_e = strip(qual_b, e3);
                if (seq_e - seq_b != qual_e - qual_b) {
                    fail(path_ + ": sequence and quality lengths differ at line " + std::to_string(lineno + 1));
                    return;
                }
                if (!held) { cur.bufs.push_back(owner); held = true; }
                Rec r;
                r.name = name_b; r.name_len = (uint32_t)(name_e - name_b);
                r.seq = seq_b; r.seq_len = (uint32_t)(seq_e - seq_b);
                r.qual = qual_b; r.qual_len = (uint32_t)(qual_e - qual_b);
                cur.recs.push_back(r);
                lineno += 4;
                pos = (size_t)(e3 + 1 - p);
                if (cur.recs.size() == BATCH_RECORDS) {
                    if (!batches_.push(std::move(cur))) return;
                    cur = Batch();
                    cur.recs.reserve(BATCH_RECORDS);
                    held = false;
                }
                if (pos == n) break;
            }
            if (pos < n) {
                if (eof) {
                    // leftover bytes that are not a whole record: blank lines are tolerated, anything else is an error
                    bool blank = true;
                    for (size_t i = pos; i < n; i++) blank = blank && (p[i] == '\n' || p[i] == '\r');
                    if (!blank) { fail(path_ + ": premature end of file (incomplete FASTQ record)"); return; }
                } else {
                    carry.assign(p + pos, p + n);
                }
            }
        }
        if (!cur.recs.empty()) batches_.push(std::move(cur));
        batches_.close();
    }

    std::string path_;
    Channel<Chunk> chunks_;
    Channel<Batch> batches_;
    std::thread reader_, parser_;
};

// ---- Python slice semantics ------------------------------------------------------------------------------

struct SliceRange { int64_t start, step, count; };

// indices of s[slice] for a string of `len` symbols: PySlice_Unpack + PySlice_AdjustIndices of CPython
SliceRange resolve_slice(const fqd_slice *sl, int64_t len)
{
    if (!sl) return {0, 1, len};
    const int64_t step = sl->has_step ? sl->step : 1;
    constexpr int64_t BIG = INT64_MAX / 4;
    int64_t start = sl->has_start ? sl->start : (step < 0 ? BIG : 0);
    int64_t stop = sl->has_stop ? sl->stop : (step < 0 ? -BIG : BIG);
    if (start < 0) { start += len; if (start < 0) start = step < 0 ? -1 : 0; }
    else if (start >= len) start = step < 0 ? len - 1 : len;
    if (stop < 0) { stop += len; if (stop < 0) stop = step < 0 ? -1 : 0; }
    else if (stop >= len) stop = step < 0 ? len - 1 : len;
    int64_t count = 0;
    if (step < 0) { if (stop < start) count = (start - stop - 1) / (-step) + 1; }
    else if (start < stop) count = (stop - start - 1) / step + 1;
    return {start, step, count};
}

inline void append_slice(std::vector<uint8_t> &out, const char *s, const SliceRange &r)
{
    if (r.count <= 0) return;
    if (r.step == 1) { out.insert(out.end(), s + r.start, s + r.start + r.count); return; }
    for (int64_t k = 0, i = r.start; k < r.count; k++, i += r.step) out.push_back((uint8_t)s[i]);
}

// id of a record for the mates test: the name up to the first whitespace, without a trailing /1 /2 /3
// (what dnaio.records_are_mates compares)
inline void mate_id(const char *name, uint32_t len, const char *&b, uint32_t &n)
{
    uint32_t k = 0;
    while (k < len && name[k] != ' ' && name[k] != '\t' && name[k] != '\n' && name[k] != '\r' && name[k] != '\f' && name[k] != '\v') k++;
    if (k > 2 && name[k - 2] == '/' && (name[k - 1] == '1' || name[k - 1] == '2' || name[k - 1] == '3')) k -= 2;
    b = name;
    n = k;
}

// ---- the pool: batch k of every file -> one task -------------------------------------------------------------

struct Task {
    uint64_t index = 0;
    std::vector<Batch> files;   // batch `index` of every input file, cut to the same number of records
    uint32_t n = 0;
};

// Runs `work(task)` on `threads` workers for every aligned batch of the inputs; returns "" or the first error.
template <typename Work>
std::string for_each_tuple_batch(const std::vector<std::string> &paths, int threads, uint64_t *n_records, Work work)
{
    const int F = (int)paths.size();
    std::vector<std::unique_ptr<FileStream>> streams;
    for (int f = 0; f < F; f++) {
        streams.emplace_back(new FileStream());
        streams.back()->start(paths[f]);
    }
    Channel<Task> tasks((size_t)std::max(2, threads) * 2);
    std::mutex err_mu;
    std::string first_error;
    auto set_err = [&](const std::string &e) {
        std::lock_guard<std::mutex> lk(err_mu);
        if (first_error.empty()) first_error = e;
    };
    std::vector<std::thread> pool;
    for (int w = 0; w < std::max(1, threads); w++)
        pool.emplace_back([&] {
            Task t;
            while (tasks.pop(t)) {
                const std::string e = work(t);
                if (!e.empty()) { set_err(e); tasks.cancel(); return; }
            }
        });
    uint64_t total = 0, index = 0;
    for (;;) {
        Task t;
        t.index = index;
        t.files.resize(F);
        bool any_end = false;
        std::string err;
        uint32_t n = 0xFFFFFFFFu;
        for (int f = 0; f < F; f++) {
            if (!streams[f]->next(t.files[f], err)) {
                if (!err.empty()) { set_err(err); }
                any_end = true;
                t.files[f].recs.clear();
            }
            n = std::min<uint32_t>(n, (uint32_t)t.files[f].recs.size());
        }
        {
            std::lock_guard<std::mutex> lk(err_mu);
            if (!first_error.empty()) break;
        }
        if (n == 0 || n == 0xFFFFFFFFu) break;   // the shortest input is exhausted: zip() stops (__init__.py:180, :201)
        t.n = n;
        total += n;
        index++;
        const bool short_batch = n < BATCH_RECORDS;
        if (!tasks.push(std::move(t))) break;
        if (any_end || short_batch) break;      // only the last batch of a file is short
    }
    tasks.close();
    for (auto &th : pool) th.join();
    for (auto &s : streams) s->stop();
    if (n_records) *n_records = total;
    return first_error;
}

int default_threads()
{
    const unsigned hc = std::thread::hardware_concurrency();
    return (int)std::max(2u, std::min(16u, hc ? hc : 4u));
}

}  // namespace

// ---- pass 1 -------------------------------------------------------------------------------------------------

struct fqd_fastq_scan {
    uint64_t n = 0;
    std::vector<uint8_t> keys, quals;         // rows back to back (fixed stride) or ragged
    std::vector<uint64_t> key_off, qual_off;  // empty when the rows have one length
    uint32_t key_stride = 0, qual_stride = 0;
    double seconds = 0;
};

namespace {

struct ScanPart {
    std::vector<uint8_t> keys, quals;
    std::vector<uint32_t> key_len, qual_len;
};

// part buffers -> one buffer (+ offsets unless every row has the same length)
void assemble(std::map<uint64_t, ScanPart> &parts, bool quals, std::vector<uint8_t> &out, std::vector<uint64_t> &off, uint32_t &stride,
              uint64_t n)
{
    uint64_t bytes = 0;
    bool uniform = true;
    uint32_t len0 = 0;
    bool have0 = false;
    for (auto &kv : parts) {
        const std::vector<uint32_t> &lens = quals ? kv.second.qual_len : kv.second.key_len;
        for (uint32_t l : lens) {
            if (!have0) { len0 = l; have0 = true; }
            uniform = uniform && l == len0;
        }
        bytes += quals ? kv.second.quals.size() : kv.second.keys.size();
    }
    out.resize(bytes ? bytes : 1);
    off.clear();
    stride = 0;
    if (uniform && have0) stride = len0;
    else off.reserve(n + 1);
    uint64_t pos = 0;
    std::vector<std::pair<uint64_t, const std::vector<uint8_t> *>> copies;
    for (auto &kv : parts) {
        const std::vector<uint8_t> &src = quals ? kv.second.quals : kv.second.keys;
        copies.emplace_back(pos, &src);
        if (!(uniform && have0)) {
            uint64_t p = pos;
            for (uint32_t l : (quals ? kv.second.qual_len : kv.second.key_len)) { off.push_back(p); p += l; }
        }
        pos += src.size();
    }
    if (!(uniform && have0)) off.push_back(pos);
    // parallel copy
    std::atomic<size_t> next{0};
    std::vector<std::thread> th;
    for (int w = 0; w < default_threads(); w++)
        th.emplace_back([&] {
            for (size_t i = next++; i < copies.size(); i = next++)
                if (!copies[i].second->empty()) memcpy(out.data() + copies[i].first, copies[i].second->data(), copies[i].second->size());
        });
    for (auto &t : th) t.join();
}

}  // namespace

extern "C" {

int fqd_fastq_scan_open(const char *const *paths, int n_files, const fqd_slice *slices, int want_quals, int threads,
                        fqd_fastq_scan **out)
{
    if (!paths || n_files < 1 || !out) { fqd::set_error("fqd_fastq_scan_open: bad arguments"); return FQD_ERR_ARG; }
    *out = nullptr;
    for (int f = 0; slices && f < n_files; f++)
        if (slices[f].has_step && slices[f].step == 0) { fqd::set_error("slice step cannot be zero"); return FQD_ERR_ARG; }
    std::vector<std::string> p(paths, paths + n_files);
    const auto t0 = std::chrono::steady_clock::now();
    std::mutex mu;
    std::map<uint64_t, ScanPart> parts;
    uint64_t n = 0;
    const std::string err = for_each_tuple_batch(p, threads > 0 ? threads : default_threads(), &n, [&](Task &t) -> std::string {
        ScanPart part;
        part.key_len.reserve(t.n);
        if (want_quals) part.qual_len.reserve(t.n);
        for (uint32_t i = 0; i < t.n; i++) {
            if (n_files > 1) {   // fastq_files_to_records, __init__.py:181-185
                const Rec &r0 = t.files[0].recs[i];
                const char *id0; uint32_t n0;
                mate_id(r0.name, r0.name_len, id0, n0);
                for (int f = 1; f < n_files; f++) {
                    const Rec &r = t.files[f].recs[i];
                    const char *id; uint32_t nn;
                    mate_id(r.name, r.name_len, id, nn);
                    if (nn != n0 || memcmp(id, id0, n0) != 0) {
                        std::string names;
                        for (int g = 0; g < n_files; g++) {
                            const Rec &rg = t.files[g].recs[i];
                            if (g) names += ", ";
                            names.append(rg.name, rg.name_len);
                        }
                        return "FASTQ files not in sync: " + names + " are not mates.";
                    }
                }
            }
            const size_t k0 = part.keys.size(), q0 = part.quals.size();
            for (int f = 0; f < n_files; f++) {
                const Rec &r = t.files[f].recs[i];
                const SliceRange sr = resolve_slice(slices ? &slices[f] : nullptr, r.seq_len);
                append_slice(part.keys, r.seq, sr);
                if (want_quals) append_slice(part.quals, r.qual, sr);
            }
            part.key_len.push_back((uint32_t)(part.keys.size() - k0));
            if (want_quals) part.qual_len.push_back((uint32_t)(part.quals.size() - q0));
        }
        std::lock_guard<std::mutex> lk(mu);
        parts.emplace(t.index, std::move(part));
        return "";
    });
    if (!err.empty()) {
        fqd::set_error("%s", err.c_str());
        return err.rfind("cannot open", 0) == 0 || err.rfind("read error", 0) == 0 ? FQD_ERR_IO : FQD_ERR_FASTQ;
    }
    fqd_fastq_scan *s = new fqd_fastq_scan();
    s->n = n;
    assemble(parts, false, s->keys, s->key_off, s->key_stride, n);
    if (want_quals) assemble(parts, true, s->quals, s->qual_off, s->qual_stride, n);
    s->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    *out = s;
    return FQD_OK;
}

uint64_t fqd_fastq_scan_records(const fqd_fastq_scan *s) { return s ? s->n : 0; }

int fqd_fastq_scan_keys(const fqd_fastq_scan *s, const uint8_t **keys, const uint64_t **offsets, uint32_t *stride)
{
    if (!s) { fqd::set_error("null scan"); return FQD_ERR_ARG; }
    if (keys) *keys = s->keys.data();
    if (offsets) *offsets = s->key_off.empty() ? nullptr : s->key_off.data();
    if (stride) *stride = s->key_stride;
    return FQD_OK;
}

int fqd_fastq_scan_quals(const fqd_fastq_scan *s, const uint8_t **quals, const uint64_t **offsets, uint32_t *stride)
{
    if (!s) { fqd::set_error("null scan"); return FQD_ERR_ARG; }
    if (quals) *quals = s->quals.empty() ? nullptr : s->quals.data();
    if (offsets) *offsets = s->qual_off.empty() ? nullptr : s->qual_off.data();
    if (stride) *stride = s->qual_stride;
    return FQD_OK;
}

void fqd_fastq_scan_free(fqd_fastq_scan *s) { delete s; }

// ---- pass 2 -------------------------------------------------------------------------------------------------

int fqd_fastq_emit(const char *const *in_paths, const char *const *out_paths, int n_files, const uint32_t *keep_bitmap,
                   uint64_t n_records, int threads, uint64_t *n_written)
{
    if (!in_paths || !out_paths || n_files < 1 || (!keep_bitmap && n_records)) { fqd::set_error("fqd_fastq_emit: bad arguments"); return FQD_ERR_ARG; }
    std::vector<std::string> in(in_paths, in_paths + n_files), outp(out_paths, out_paths + n_files);
    std::vector<int> fds(n_files, -1);
    std::vector<bool> gz(n_files);
    for (int f = 0; f < n_files; f++) {
        fds[f] = ::open(outp[f].c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
        if (fds[f] < 0) {
            fqd::set_error("cannot open %s for writing: %s", outp[f].c_str(), strerror(errno));
            for (int g = 0; g < f; g++) ::close(fds[g]);
            return FQD_ERR_IO;
        }
        gz[f] = outp[f].size() >= 3 && outp[f].compare(outp[f].size() - 3, 3, ".gz") == 0;
    }
    // ordered writer: block `index` of every file is written after block index - 1
    std::mutex mu;
    std::condition_variable cv;
    uint64_t next_index = 0;
    std::string write_error;
    std::atomic<uint64_t> written{0};
    auto write_all = [&](int fd, const std::vector<uint8_t> &b) -> bool {
        size_t done = 0;
        while (done < b.size()) {
            const ssize_t n = ::write(fd, b.data() + done, b.size() - done);
            if (n < 0) { if (errno == EINTR) continue; return false; }
            done += (size_t)n;
        }
        return true;
    };
    uint64_t seen = 0;
    const std::string err = for_each_tuple_batch(in, threads > 0 ? threads : default_threads(), &seen, [&](Task &t) -> std::string {
        const uint64_t base = t.index * (uint64_t)BATCH_RECORDS;
        std::vector<std::vector<uint8_t>> blocks(n_files);
        uint64_t kept = 0;
        for (uint32_t i = 0; i < t.n; i++) {
            const uint64_t rec = base + i;
            if (rec >= n_records || !((keep_bitmap[rec >> 5] >> (rec & 31)) & 1u)) continue;
            kept++;
            for (int f = 0; f < n_files; f++) {
                const Rec &r = t.files[f].recs[i];
                std::vector<uint8_t> &b = blocks[f];
                b.push_back('@');
                b.insert(b.end(), r.name, r.name + r.name_len);
                b.push_back('\n');
                b.insert(b.end(), r.seq, r.seq + r.seq_len);
                b.push_back('\n'); b.push_back('+'); b.push_back('\n');
                b.insert(b.end(), r.qual, r.qual + r.qual_len);
                b.push_back('\n');
            }
        }
        for (int f = 0; f < n_files; f++) {
            if (!gz[f] || blocks[f].empty()) continue;
            // one gzip member per block, level 1 like the reference's opener (__init__.py:197-198)
            z_stream z;
            memset(&z, 0, sizeof z);
            if (deflateInit2(&z, 1, Z_DEFLATED, 15 + 16, 8, Z_DEFAULT_STRATEGY) != Z_OK) return "zlib: deflateInit2 failed";
            std::vector<uint8_t> comp(deflateBound(&z, (uLong)blocks[f].size()) + 64);
            z.next_in = blocks[f].data(); z.avail_in = (uInt)blocks[f].size();
            z.next_out = comp.data(); z.avail_out = (uInt)comp.size();
            const int rc = deflate(&z, Z_FINISH);
            const size_t produced = comp.size() - z.avail_out;
            deflateEnd(&z);
            if (rc != Z_STREAM_END) return "zlib: deflate failed";
            comp.resize(produced);
            blocks[f].swap(comp);
        }
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return next_index == t.index || !write_error.empty(); });
        if (write_error.empty())
            for (int f = 0; f < n_files; f++)
                if (!blocks[f].empty() && !write_all(fds[f], blocks[f])) { write_error = "write error on " + outp[f] + ": " + strerror(errno); break; }
        next_index = t.index + 1;
        written += kept;
        cv.notify_all();
        return write_error;
    });
    for (int f = 0; f < n_files; f++) {
        if (gz[f] && written.load() == 0) {   // an empty gzip file is still a gzip file
            z_stream z;
            memset(&z, 0, sizeof z);
            if (deflateInit2(&z, 1, Z_DEFLATED, 15 + 16, 8, Z_DEFAULT_STRATEGY) == Z_OK) {
                std::vector<uint8_t> comp(64);
                z.next_out = comp.data(); z.avail_out = (uInt)comp.size();
                deflate(&z, Z_FINISH);
                comp.resize(comp.size() - z.avail_out);
                deflateEnd(&z);
                write_all(fds[f], comp);
            }
        }
        ::close(fds[f]);
    }
    if (n_written) *n_written = written.load();
    if (!err.empty()) {
        fqd::set_error("%s", err.c_str());
        return err.rfind("cannot open", 0) == 0 || err.rfind("read error", 0) == 0 || err.rfind("write error", 0) == 0 ? FQD_ERR_IO : FQD_ERR_FASTQ;
    }
    return FQD_OK;
}

}  // extern "C"
