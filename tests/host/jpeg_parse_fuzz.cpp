// jpeg_parse_fuzz.cpp -- the marker parser and table builder of the capture-side decode (cvs_jpeg_host.hpp) fed with
// damaged camera frames: truncations at every header byte and seeded random byte / length-field mutations of the header.
// The bytes come from a camera (untrusted), so whatever they are parse() must return a status without reading outside
// the buffer.  Built by tests/test_jpeg_parse_fuzz.py with -fsanitize=address,undefined: any out-of-bounds access aborts.
// The frame is copied into an exactly-sized heap block for every call so that one byte too far is caught.
//
//   jpeg_parse_fuzz <in.jpg> <mutations> <seed>      prints: calls, ok, not-jpeg, unsupported
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../cudavideostream_b200/csrc/cvs_jpeg_host.hpp"

using namespace cvs::jpg;

static uint64_t rng_state;
static uint64_t splitmix64()
{
    uint64_t z = (rng_state += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

static unsigned long counts[3];
static Parsed P;

// a status outside the enum or inconsistent accepted geometry is a failure of its own
static bool call(const uint8_t *src, size_t n, uint32_t sub_bits)
{
    uint8_t *blk = (uint8_t *)malloc(n ? n : 1);
    memcpy(blk, src, n);
    const ParseStatus ps = parse(blk, n, sub_bits, &P);
    bool good = ps == kParseOk || ps == kParseNotJpeg || ps == kParseUnsupported;
    if (ps == kParseOk) {
        good = good && P.scan_offset + P.scan_bytes + 2 <= n && P.scan_bytes > 0 && P.g.width > 0 && P.g.height > 0 &&
               (P.g.ncomp == 1 || P.g.ncomp == 3) && P.g.bpm >= 1 && P.g.bpm <= 6 &&
               P.g.nblocks == (uint32_t)P.g.mcux * (uint32_t)P.g.mcuy * (uint32_t)P.g.bpm;
        // the header of an accepted frame can be handed to the fast path of the next frame
        Parsed Q = P;
        good = good && reparse_same_header(blk, n, blk, P.scan_offset, sub_bits, &Q) && Q.scan_bytes == P.scan_bytes;
    }
    free(blk);
    if (good) counts[(int)ps]++;
    return good;
}

int main(int argc, char **argv)
{
    if (argc < 4) return 2;
    FILE *f = fopen(argv[1], "rb");
    if (!f) return 2;
    std::vector<uint8_t> d;
    uint8_t buf[65536];
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) d.insert(d.end(), buf, buf + n);
    fclose(f);
    const long mutations = atol(argv[2]);
    rng_state = strtoull(argv[3], nullptr, 0);
    if (parse(d.data(), d.size(), 1024, &P) != kParseOk) {
        printf("the undamaged frame does not parse\n");
        return 3;
    }
    const size_t header = P.scan_offset; // everything up to the entropy-coded segment
    // 1. every truncation inside the header (+ a little of the scan), and the empty buffer
    for (size_t cut = 0; cut <= header + 8 && cut <= d.size(); cut++)
        if (!call(d.data(), cut, 1024)) { printf("bad result, truncation at %zu\n", cut); return 1; }
    // 2. truncated header followed by an EOI (a parser that trusts a segment length walks past the end here)
    for (size_t cut = 2; cut <= header; cut++) {
        std::vector<uint8_t> t(d.begin(), d.begin() + cut);
        t.push_back(0xFF); t.push_back(0xD9);
        if (!call(t.data(), t.size(), 1024)) { printf("bad result, truncation + EOI at %zu\n", cut); return 1; }
    }
    // 3. random damage: 1..4 header bytes replaced (half of the time by 0x00 / 0xFF / a marker code), on a copy that keeps only
    //    a short piece of the scan so that a call stays cheap
    const size_t keep = header + 256 < d.size() ? header + 256 : d.size();
    std::vector<uint8_t> base(d.begin(), d.begin() + keep);
    base.push_back(0xFF); base.push_back(0xD9);
    static const uint8_t special[] = {0x00, 0xFF, 0xC0, 0xC2, 0xC4, 0xDA, 0xDB, 0xDD, 0xD9, 0xD8, 0x01, 0x10, 0x11, 0x22, 0x7F, 0x80};
    for (long it = 0; it < mutations; it++) {
        std::vector<uint8_t> t = base;
        const int k = 1 + (int)(splitmix64() % 4);
        for (int j = 0; j < k; j++) {
            const size_t at = 2 + (size_t)(splitmix64() % (header - 2));
            const uint64_t r = splitmix64();
            t[at] = (r & 1) ? special[(r >> 8) % sizeof special] : (uint8_t)(r >> 16);
        }
        const uint32_t sub_bits = (it & 1) ? 1024u : 256u;
        if (!call(t.data(), t.size(), sub_bits)) { printf("bad result, mutation %ld\n", it); return 1; }
    }
    printf("calls %lu ok %lu notjpeg %lu unsupported %lu\n", counts[0] + counts[1] + counts[2], counts[0], counts[1], counts[2]);
    return 0;
}
