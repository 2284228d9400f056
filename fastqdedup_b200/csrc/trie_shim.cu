// trie_shim.cu -- the legacy `_trie.Trie` surface (reference _triemodule.c:596-983) on top
// of the batched GPU path.
//
// The reference keeps its sequences in a pointer-chasing radix trie and answers
// contains_sequence / pop_cluster by a depth-first walk.  Here the container is a host-side
// staging map (sequence -> count); the *arithmetic* -- which staged sequences are within
// the distance of which -- always runs on the GPU: pop_cluster runs the batched
// clustering job over the staged sequences and hands out its connected components,
// contains_sequence runs the batched within_distance kernel of the query against every
// staged sequence.  Only bookkeeping (grouping by label, the lazily grown alphabet that
// the reference's tests pin, ordering of the hand-out) is host code.
#include <string.h>

#include <algorithm>
#include <map>
#include <string>
#include <vector>

#include "common.h"

struct fqd_trie {
    fqd_context *ctx = nullptr;
    std::map<std::string, uint32_t> present;   // byte-lexicographic
    std::vector<uint8_t> alphabet;             // registration order (_triemodule.c:266-273)
    uint8_t to_index[256];
    uint64_t number_of_sequences = 0;
    uint32_t max_sequence_size = 0;
    // clustering cache: valid for (d, edit) until a sequence is added
    bool cache_valid = false;
    int cache_d = -1, cache_edit = -1;
    std::vector<std::vector<std::string>> pending;   // clusters not handed out yet (reverse order)
    // last popped cluster
    std::vector<std::pair<uint32_t, std::string>> last;
};

namespace {

void register_char(fqd_trie *t, uint8_t c)
{
    if (t->to_index[c] != 255 || t->alphabet.size() >= 254) return;
    t->to_index[c] = (uint8_t)t->alphabet.size();
    t->alphabet.push_back(c);
}

size_t lcp(const std::string &a, const std::string &b)
{
    size_t n = std::min(a.size(), b.size()), i = 0;
    while (i < n && a[i] == b[i]) i++;
    return i;
}

// Position of `c` in the trie's child order; characters that were never branched on have
// no index yet -- they only ever occur below a point where two sequences diverge, so they
// never decide a comparison.
int child_index(const fqd_trie *t, uint8_t c) { return t->to_index[c]; }

// DFS order of the reference trie (TrieNode_GetSequence, _triemodule.c:510-551): children
// in alphabet-index order, and a node's own sequence only after all its children -- an
// extension precedes its proper prefix.
bool trie_order_less(const fqd_trie *t, const std::string &a, const std::string &b)
{
    const size_t m = lcp(a, b);
    if (m == a.size() || m == b.size()) return a.size() > b.size();
    return child_index(t, (uint8_t)a[m]) < child_index(t, (uint8_t)b[m]);
}

}  // namespace

using namespace fqd;

extern "C" {

int fqd_trie_new(fqd_context *ctx, const uint8_t *alphabet, size_t alphabet_len, fqd_trie **out)
{
    if (!out) { set_error("null output pointer"); return FQD_ERR_ARG; }
    *out = nullptr;
    if (!ctx) { set_error("a Trie needs a GPU context: fastqdedup_b200 has no CPU fallback"); return FQD_ERR_CUDA; }
    fqd_trie *t = new fqd_trie();
    t->ctx = ctx;
    memset(t->to_index, 255, sizeof t->to_index);
    for (size_t i = 0; i < alphabet_len; i++) {
        const uint8_t c = alphabet[i];
        if (c >= 128) { delete t; set_error("Alphabet should be an ASCII string."); return FQD_ERR_ARG; }
        if (t->to_index[c] != 255) {
            delete t;
            set_error("Alphabet should consist of unique characters.Character %c was repeated. ", c);
            return FQD_ERR_ARG;
        }
        if (t->alphabet.size() >= 254) { delete t; set_error("Maximum alphabet length exceeded"); return FQD_ERR_ARG; }
        t->to_index[c] = (uint8_t)t->alphabet.size();
        t->alphabet.push_back(c);
    }
    *out = t;
    return FQD_OK;
}

void fqd_trie_free(fqd_trie *trie) { delete trie; }

int fqd_trie_add_sequence(fqd_trie *t, const uint8_t *seq, size_t len)
{
    if (!t) { set_error("null trie"); return FQD_ERR_ARG; }
    if (len > 0xFFFFFFFFull) { set_error("Sequences larger than %u can not be stored in the Trie", 0xFFFFFFFFu); return FQD_ERR_ARG; }
    std::string s(reinterpret_cast<const char *>(seq), len);
    auto it = t->present.lower_bound(s);
    if (it != t->present.end() && it->first == s) {
        it->second += 1;
    } else {
        // Lazy alphabet growth exactly as the radix trie does it: a character is registered
        // when the insertion walks through (or splits) a node at its depth.  With partner =
        // the staged sequence sharing the longest prefix m with s, the walk registers
        // s[0..m-1], then the partner's character at depth m (the split re-inserts the
        // leaf's suffix first, _triemodule.c:241-259), then s[m].
        const std::string *partner = nullptr;
        size_t m = 0;
        if (it != t->present.end()) { partner = &it->first; m = lcp(s, it->first); }
        if (it != t->present.begin()) {
            auto pv = std::prev(it);
            const size_t mp = lcp(s, pv->first);
            if (!partner || mp > m) { partner = &pv->first; m = mp; }
        }
        if (partner) {
            for (size_t i = 0; i < m; i++) register_char(t, (uint8_t)s[i]);
            if (partner->size() > m) register_char(t, (uint8_t)(*partner)[m]);
            if (s.size() > m) register_char(t, (uint8_t)s[m]);
        }
        t->present.emplace_hint(it, std::move(s), 1u);
        t->cache_valid = false;
    }
    t->number_of_sequences += 1;
    if (len > t->max_sequence_size) t->max_sequence_size = (uint32_t)len;
    return FQD_OK;
}

int fqd_trie_contains_sequence(fqd_trie *t, const uint8_t *seq, size_t len, int32_t max_distance,
                               int32_t use_edit_distance, int32_t *found)
{
    if (!t || !found) { set_error("null argument"); return FQD_ERR_ARG; }
    *found = 0;
    const size_t n = t->present.size();
    if (!n) return FQD_OK;   // the reference dereferences NULL here (:755 -> :390); an empty trie contains nothing
    // one pair (query, staged sequence) per staged sequence, evaluated on the GPU
    std::vector<uint8_t> a, b;
    std::vector<uint64_t> ao(n + 1), bo(n + 1);
    a.reserve(n * len);
    size_t i = 0;
    for (const auto &kv : t->present) {
        ao[i] = a.size(); bo[i] = b.size();
        a.insert(a.end(), seq, seq + len);
        b.insert(b.end(), kv.first.begin(), kv.first.end());
        i++;
    }
    ao[n] = a.size(); bo[n] = b.size();
    std::vector<uint8_t> out(n);
    FQD_TRY(fqd_within_distance(t->ctx, a.data(), ao.data(), b.data(), bo.data(), n, max_distance,
                                use_edit_distance, out.data()));
    for (uint8_t o : out) if (o) { *found = 1; break; }
    return FQD_OK;
}

int fqd_trie_pop_cluster(fqd_trie *t, int32_t max_distance, int32_t use_edit_distance, uint64_t *n_items)
{
    if (!t || !n_items) { set_error("null argument"); return FQD_ERR_ARG; }
    *n_items = 0;
    if (max_distance < 0) { set_error("max_distance should be non-negative"); return FQD_ERR_ARG; }
    if (t->present.empty()) { set_error("No sequences left in Trie."); return FQD_ERR_LOOKUP; }
    const int edit = use_edit_distance ? 1 : 0;
    if (!t->cache_valid || t->cache_d != max_distance || t->cache_edit != edit) {
        // cluster every staged sequence on the GPU (counts do not influence components)
        const size_t n = t->present.size();
        std::vector<const std::string *> order;
        order.reserve(n);
        std::vector<uint8_t> flat;
        std::vector<uint64_t> off(n + 1);
        size_t i = 0;
        std::string alpha(t->alphabet.begin(), t->alphabet.end());
        for (const auto &kv : t->present) {
            off[i++] = flat.size();
            flat.insert(flat.end(), kv.first.begin(), kv.first.end());
            order.push_back(&kv.first);
        }
        off[n] = flat.size();
        uint8_t dummy = 0;
        fqd_cluster_job job{};
        job.n_records = n;
        job.keys = flat.empty() ? &dummy : flat.data();
        job.key_offsets = off.data();
        job.max_distance = max_distance;
        job.use_edit_distance = edit;
        job.method = FQD_METHOD_HIGHEST_COUNT;
        job.memory_space = FQD_MEM_HOST;
        job.max_average_error_rate = 1.0;
        job.phred_offset = FQD_DEFAULT_PHRED_OFFSET;
        job.alphabet = alpha.empty() ? nullptr : alpha.c_str();
        fqd_cluster_stats st;
        FQD_TRY(fqd_cluster(t->ctx, &job, &st, nullptr));
        if (st.number_of_uniques != n) { set_error("internal: staged sequences were not unique"); return FQD_ERR_CUDA; }
        std::vector<uint64_t> first(n), label(n);
        FQD_TRY(fqd_cluster_fetch(t->ctx, first.data(), nullptr, label.data(), nullptr));
        // group by label (labels are record indices of the staged order)
        std::map<uint64_t, std::vector<std::string>> groups;
        for (size_t u = 0; u < n; u++) groups[label[u]].push_back(*order[first[u]]);
        t->pending.clear();
        for (auto &g : groups) {
            auto &members = g.second;
            // seed = first member in the trie's DFS order; it leads the list like the
            // reference's cluster[0] (:813-843)
            auto seed = std::min_element(members.begin(), members.end(),
                                         [&](const std::string &x, const std::string &y) { return trie_order_less(t, x, y); });
            std::iter_swap(members.begin(), seed);
            t->pending.push_back(std::move(members));
        }
        // hand out in the order the reference would seed them; pop from the back
        std::sort(t->pending.begin(), t->pending.end(),
                  [&](const std::vector<std::string> &x, const std::vector<std::string> &y) {
                      return trie_order_less(t, y[0], x[0]);
                  });
        t->cache_valid = true;
        t->cache_d = max_distance;
        t->cache_edit = edit;
    }
    std::vector<std::string> members = std::move(t->pending.back());
    t->pending.pop_back();
    t->last.clear();
    for (auto &m : members) {
        auto it = t->present.find(m);
        const uint32_t c = it->second;
        t->number_of_sequences -= c;
        t->present.erase(it);
        t->last.emplace_back(c, std::move(m));
    }
    *n_items = t->last.size();
    return FQD_OK;
}

int fqd_trie_cluster_item(fqd_trie *t, uint64_t i, uint32_t *count, const uint8_t **seq, size_t *len)
{
    if (!t || i >= t->last.size()) { set_error("cluster item index out of range"); return FQD_ERR_ARG; }
    if (count) *count = t->last[i].first;
    if (seq) *seq = reinterpret_cast<const uint8_t *>(t->last[i].second.data());
    if (len) *len = t->last[i].second.size();
    return FQD_OK;
}

uint64_t fqd_trie_number_of_sequences(const fqd_trie *t) { return t ? t->number_of_sequences : 0; }

size_t fqd_trie_alphabet(const fqd_trie *t, uint8_t *buf, size_t cap)
{
    if (!t) return 0;
    const size_t n = t->alphabet.size();
    if (buf) memcpy(buf, t->alphabet.data(), std::min(n, cap));
    return n;
}

// The new path has no trie nodes, but the node layout of the reference's trie is a function of the key set alone
// (one level per character, a key that is alone below a node is a leaf holding its suffix, a node has as many
// child slots as the largest alphabet index below it + 1: TrieNode_AddSequence, _triemodule.c:222-288).  raw_stats
// (:929-964) and memory_size (:909-913) therefore report exactly what the reference's trie over the same sequences
// holds when freshly built, so trie_stats (__init__.py:133-157, -v) prints the reference's table.
}  // extern "C"

namespace {

struct TrieShape {
    const fqd_trie *t;
    std::vector<const std::string *> keys;   // byte-lexicographic
    uint64_t *stats;                         // rows of row_len counters, may be null
    size_t row_len, rows;
    uint64_t bytes = 0;

    void node(size_t lo, size_t hi, size_t depth)
    {
        if (hi - lo == 1) {   // a leaf: header + the rest of the key (TrieNode_GetMemorySize, :553-570)
            bytes += 8 + (keys[lo]->size() - depth);
            if (stats && depth < rows) stats[depth * row_len] += 1;
            return;
        }
        // keys equal to the prefix itself (at most one, first in the range) count at this node and have no child
        size_t i = lo;
        if (keys[i]->size() == depth) i++;
        uint32_t slots = 0;
        for (size_t k = i; k < hi; k++) slots = std::max<uint32_t>(slots, (uint32_t)t->to_index[(uint8_t)(*keys[k])[depth]] + 1u);
        bytes += 8 + 8ull * slots;
        if (stats && depth < rows && slots < row_len) stats[depth * row_len + slots] += 1;
        while (i < hi) {
            const char c = (*keys[i])[depth];
            size_t j = i + 1;
            while (j < hi && (*keys[j])[depth] == c) j++;
            node(i, j, depth + 1);
            i = j;
        }
    }
};

uint64_t trie_shape(const fqd_trie *t, uint64_t *stats, size_t row_len, size_t rows)
{
    TrieShape sh{t, {}, stats, row_len, rows};
    sh.keys.reserve(t->present.size());
    for (const auto &kv : t->present) sh.keys.push_back(&kv.first);
    if (!sh.keys.empty()) sh.node(0, sh.keys.size(), 0);
    return sh.bytes;
}

}  // namespace

extern "C" {

uint64_t fqd_trie_memory_size(const fqd_trie *t)
{
    if (!t) return 0;
    return trie_shape(t, nullptr, 0, 0);
}

size_t fqd_trie_raw_stats(const fqd_trie *t, uint64_t *buf, size_t cap, size_t *row_len)
{
    if (!t) return 0;
    const size_t rl = t->alphabet.size() + 1, rows = (size_t)t->max_sequence_size + 1;
    if (row_len) *row_len = rl;
    if (buf && cap >= rl * rows) {
        memset(buf, 0, rl * rows * sizeof(uint64_t));
        trie_shape(t, buf, rl, rows);
    }
    return rows;
}

}  // extern "C"
