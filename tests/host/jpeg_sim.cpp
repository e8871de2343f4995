// jpeg_sim.cpp -- runs the GPU JPEG decoder's entropy stage on the CPU: the same __host__ __device__ token functions
// (cudavideostream_b200/csrc/cvs_jpeg.cuh) and the same round structure as k_entropy (guess, hand the exit state to the
// successor, decode again whoever received a new entry state, until nothing changes; prefix sums; write pass), executed
// sequentially.  Built with plain g++ by tests/test_jpeg_host_sim.py, which compares the coefficients with the oracle's.
//
//   jpeg_sim <in.jpg> <sub_bits> <out.coef> [hypotheses=1]     prints: rounds, runs per round, blocks
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
    if (P.restart_interval) { // independent intervals: k_entropy_restart, no synchronisation to simulate
        printf("restart interval %u\n", P.restart_interval);
        return 3;
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
    const TableRef tbr = table_ref(&P.t);
    const uint32_t nsub = (T + S - 1) / S;
    const bool hypotheses = argc < 5 || atoi(argv[4]) != 0;
    // the rounds, the prefix sums and the write pass work on HALF subsequences when the hypotheses deliver the states in the
    // middle of the subsequences too (as k_entropy does)
    MidRecords mid;
    mid.nsplit = S % 256 == 0 ? 8u : (S % 128 == 0 ? 4u : 1u);
    mid.G = S / mid.nsplit;
    const bool half = hypotheses && S % 256 == 0 && mid.nsplit % 2 == 0;
    const uint32_t hstep = half ? 2u : 1u;
    std::vector<uint32_t> entry(2 * (size_t)nsub + 2, 0);
    if (hypotheses) {
        // k_entropy's hypothesis phases: X (a block of phase h starts at the boundary), Y (X of the subsequence in front
        // followed through this one), W (Y followed once more) and the maps "candidate of boundary i-1 -> candidate of
        // boundary i"; candidate 0 of boundary 0 is followed through the maps and its states seed entry[]
        const uint32_t B = (uint32_t)g.bpm;
        std::vector<uint32_t> hx((size_t)nsub * B), hy((size_t)nsub * B, kStateUnset), hym((size_t)nsub * B, kStateUnset),
            hwm((size_t)nsub * B, kStateUnset);
        std::vector<uint8_t> hmap((size_t)nsub * 16, (uint8_t)kNoCandidate);
        auto run = [&](uint32_t i, uint32_t e, uint32_t *hs) {
            const RunResult r = run_subsequence<false, false, true>(tbr, g, words, T, i, e, kZigzagNatural, nullptr, 0, 0, 0, 0);
            if (hs) *hs = r.half_state;
            return r.exit_state;
        };
        for (uint32_t i = 0; i < nsub; i++)
            for (uint32_t h = 0; h < B; h++) hx[(size_t)i * B + h] = run(i, pack_state(0, h, 0), (i == 0 && h == 0) ? &hym[0] : nullptr);
        for (uint32_t i = 1; i < nsub; i++)
            for (uint32_t h = 0; h < B; h++) hy[(size_t)i * B + h] = run(i, hx[(size_t)(i - 1) * B + h], &hym[(size_t)i * B + h]);
        for (uint32_t i = 1; i < nsub; i++)
            for (uint32_t h = 0; h < B; h++) {
                uint32_t m0 = B + h, m1 = kNoCandidate;
                for (uint32_t h2 = 0; h2 < B; h2++)
                    if (hx[(size_t)i * B + h2] == hy[(size_t)i * B + h]) {
                        m0 = h2;
                        break;
                    }
                if (i >= 2) {
                    const uint32_t w = run(i, hy[(size_t)(i - 1) * B + h], &hwm[(size_t)i * B + h]);
                    for (uint32_t h2 = 0; h2 < B && m1 == kNoCandidate; h2++)
                        if (hx[(size_t)i * B + h2] == w) m1 = h2;
                    for (uint32_t h2 = 0; h2 < B && m1 == kNoCandidate; h2++)
                        if (hy[(size_t)i * B + h2] == w) m1 = B + h2;
                }
                hmap[(size_t)i * 16 + h] = (uint8_t)m0;
                hmap[(size_t)i * 16 + B + h] = (uint8_t)m1;
            }
        uint32_t t = 0, lost = 0;
        entry[hstep] = hx[0];
        if (half) entry[1] = hym[0] == kStateUnset ? 0u : hym[0];
        for (uint32_t i = 1; i < nsub; i++) {
            if (half) {
                uint32_t hs = kStateUnset;
                if (t < B) hs = hym[(size_t)i * B + t];
                else if (t < 2 * B) hs = hwm[(size_t)i * B + t - B];
                entry[2 * i + 1] = hs == kStateUnset ? 0u : hs;
            }
            if (t < 12u) t = hmap[(size_t)i * 16 + t];
            uint32_t st = hx[(size_t)i * B];
            if (t < B) st = hx[(size_t)i * B + t];
            else if (t < 2 * B) st = hy[(size_t)i * B + t - B];
            else lost++;
            entry[hstep * (i + 1)] = st;
        }
        printf("hypotheses: %u of %u boundaries without a candidate\n", lost, nsub);
    }
    Geometry gr = g;
    uint32_t nsubr = nsub;
    if (half) {
        gr.sub_bits = S / 2;
        mid.nsplit /= 2;
        nsubr = (T + gr.sub_bits - 1) / gr.sub_bits;
    }
    std::vector<uint32_t> used(nsubr, 0), nblk(nsubr, 0);
    std::vector<int32_t> dcs(3 * (size_t)nsubr, 0);
    // inner-boundary records of the counting runs: the write pass runs with one thread per G bits (as k_entropy does)
    mid.stride = mid.nsplit * nsubr;
    std::vector<uint32_t> mid_state((size_t)mid.stride, kStateUnset), mid_nblk((size_t)mid.stride, 0);
    std::vector<int32_t> mid_dc(3 * (size_t)mid.stride, 0);
    mid.state = mid_state.data();
    mid.nblk = mid_nblk.data();
    mid.dc = mid_dc.data();
    int rounds = 0;
    for (uint32_t round = 0;; round++) {
        std::vector<uint32_t> entry_in = entry; // all threads of a round see the states of the previous round
        uint32_t changed = 0, runs = 0;
        for (uint32_t i = 0; i < nsubr; i++) {
            const uint32_t e = i == 0 ? pack_state(0, 0, 0) : entry_in[i];
            if (round && e == used[i]) continue;
            runs++;
            const RunResult r = run_subsequence<false, true>(tbr, gr, words, T, i, e, kZigzagNatural, nullptr, 0, 0, 0, 0, &mid);
            used[i] = e;
            nblk[i] = r.nblocks;
            dcs[i] = r.dc0; dcs[(size_t)nsubr + i] = r.dc1; dcs[2 * (size_t)nsubr + i] = r.dc2;
            if (entry[i + 1] != r.exit_state) {
                entry[i + 1] = r.exit_state;
                if (i + 1 < nsubr) changed++;
            }
        }
        printf("round %u: %u runs, %u exit states changed\n", round, runs, changed);
        rounds = (int)round + 1;
        if (!changed) break;
        if (round > 1000000) return 5;
    }
    std::vector<int16_t> coef((size_t)g.nblocks * 64, 0);
    uint32_t base = 0;
    int32_t pred[3] = {0, 0, 0};
    uint32_t total = 0;
    Geometry gw = gr;
    gw.sub_bits = mid.G;
    for (uint32_t i = 0; i < nsubr; i++) {
        uint32_t seen = 0;
        for (uint32_t part = 0; part < mid.nsplit; part++) {
            const uint32_t j = i * mid.nsplit + part;
            uint32_t e = used[i], blk0 = base;
            int32_t d[3] = {pred[0], pred[1], pred[2]};
            if (part) {
                e = mid_state[j];
                if (e == kStateUnset) continue;
                blk0 += mid_nblk[j];
                for (int c = 0; c < 3; c++) d[c] += mid_dc[(size_t)c * mid.stride + j];
            }
            const RunResult r = run_subsequence<true>(tbr, gw, words, T, j, e, kZigzagNatural, coef.data(), blk0, d[0], d[1], d[2]);
            seen += r.nblocks;
        }
        if (seen != nblk[i]) return 6;
        base += nblk[i];
        for (int c = 0; c < 3; c++) pred[c] += dcs[(size_t)c * nsubr + i];
        total = base;
    }
    printf("rounds %d subsequences %u bits %u blocks %u expected %u\n", rounds, nsub, T, total, g.nblocks);
    FILE *o = fopen(argv[3], "wb");
    fwrite(coef.data(), 2, coef.size(), o);
    fclose(o);
    return total >= g.nblocks ? 0 : 7;
}
