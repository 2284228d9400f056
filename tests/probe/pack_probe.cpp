// Host-only probe of the plane-stream packer (csrc/host_pack.cpp), for tests/test_host_pack.py: the library uses it
// internally (HOST jobs), so the C ABI does not export it.
#include <stdarg.h>

#include "../../fastqdedup_b200/csrc/host_pack.cpp"

namespace fqd {
void set_error(const char *, ...) {}
}  // namespace fqd

extern "C" uint64_t pack_planes_probe(const uint8_t *src, uint64_t n, uint32_t L, uint64_t *dst)
{
    return fqd::pack_planes_parallel(src, n, L, dst);
}
extern "C" uint64_t plane_stream_words_probe(uint64_t n, uint32_t L) { return fqd::plane_stream_words(n, L); }
