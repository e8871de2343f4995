// jpeg_sim.cpp -- runs the GPU JPEG decoder's entropy stage on the CPU: the same __host__ __device__ token functions
// (cudavideostream_b200/csrc/cvs_jpeg.cuh) and the same round structure as k_entropy (guess, hand the exit state to the
// successor, decode again whoever received a new entry state, until nothing changes; prefix sums; write pass), executed
// sequentially.  Built with plain g++ by tests/test_jpeg_host_sim.py, which compares the coefficients with the oracle's.
//
//   jpeg_sim <in.jpg> <sub_bits> <out.coef>      prints: rounds, runs per round, blocks
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../cudavideostream_b200/csrc/cvs_jpeg_host.hpp"

using namespace cvs::jpg;

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
    const uint32_t S = (uint32_t)atoi(argv[2]);
    static Parsed P;
    const ParseStatus ps = parse(d.data(), d.size(), S, &P);
    if (ps != kParseOk) {
        printf("parse status %d\n", (int)ps);
        return ps == kParseUnsupported ? 3 : 4;
    }
    // unstuff (k_unstuff_*)
    std::vector<uint8_t> u;
    const uint8_t *raw = d.data() + P.scan_offset;
    for (size_t i = 0; i < P.scan_bytes; i++)
        if (!(raw[i] == 0x00 && i > 0 && raw[i - 1] == 0xFF)) u.push_back(raw[i]);
    const uint32_t T = 8u * (uint32_t)u.size();
    u.resize(u.size() + 64, 0);
    const uint32_t *words = reinterpret_cast<const uint32_t *>(u.data());
    const Geometry g = P.g;
    const uint32_t nsub = (T + S - 1) / S;
    std::vector<uint32_t> entry(nsub + 1, 0), used(nsub, 0), nblk(nsub, 0);
    std::vector<int32_t> dcs(3 * (size_t)nsub, 0);
    int rounds = 0;
    for (uint32_t round = 0;; round++) {
        std::vector<uint32_t> entry_in = entry; // all threads of a round see the states of the previous round
        uint32_t changed = 0, runs = 0;
        for (uint32_t i = 0; i < nsub; i++) {
            uint32_t e = i == 0 ? pack_state(0, 0, 0) : entry_in[i];
            if (round == 0 && i) e = pack_state(0, 0, 0);
            if (round && e == used[i]) continue;
            runs++;
            const RunResult r = run_subsequence<false>(P.t, g, words, T, i, e, kZigzagNatural, nullptr, 0, 0, 0, 0);
            used[i] = e;
            nblk[i] = r.nblocks;
            for (int c = 0; c < 3; c++) dcs[(size_t)c * nsub + i] = r.dcsum[c];
            if (round == 0 || entry[i + 1] != r.exit_state) {
                entry[i + 1] = r.exit_state;
                if (round && i + 1 < nsub) changed++;
            }
        }
        printf("round %u: %u runs, %u exit states changed\n", round, runs, changed);
        rounds = (int)round + 1;
        if (round && !changed) break;
        if (round > 100000) return 5;
    }
    std::vector<int16_t> coef((size_t)g.nblocks * 64, 0);
    uint32_t base = 0;
    int32_t pred[3] = {0, 0, 0};
    uint32_t total = 0;
    for (uint32_t i = 0; i < nsub; i++) {
        const RunResult r = run_subsequence<true>(P.t, g, words, T, i, used[i], kZigzagNatural, coef.data(), base, pred[0], pred[1], pred[2]);
        if (r.nblocks != nblk[i]) return 6;
        base += nblk[i];
        for (int c = 0; c < 3; c++) pred[c] += dcs[(size_t)c * nsub + i];
        total = base;
    }
    printf("rounds %d subsequences %u bits %u blocks %u expected %u\n", rounds, nsub, T, total, g.nblocks);
    FILE *o = fopen(argv[3], "wb");
    fwrite(coef.data(), 2, coef.size(), o);
    fclose(o);
    return total >= g.nblocks ? 0 : 7;
}
