// cvs_jpeg_host.hpp -- host side of the GPU JPEG decode: marker parsing (T.81 Annex B) and the decoder tables.
// Plain C++ (no CUDA): also compiled by tests/host/jpeg_sim.cpp.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#include "cvs_jpeg.cuh"

namespace cvs {
namespace jpg {

struct Parsed {
    Geometry g;
    Tables t;
    size_t scan_offset = 0; // first byte of the entropy-coded segment
    size_t scan_bytes = 0;  // up to (not including) the EOI marker
    uint32_t restart_interval = 0; // MCUs per restart interval (DRI); 0: none
};

enum ParseStatus {
    kParseOk = 0,
    kParseNotJpeg = 1,     // not a JPEG / truncated / inconsistent
    kParseUnsupported = 2, // a JPEG, but not the form this decoder covers (progressive, arithmetic coding, other samplings ...)
};

static const uint8_t kZigzagNatural[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                           41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                           30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

/* ITU-T T.81 Annex K.3 (the tables a camera's MJPG frame implies when it carries no DHT segment; libjpeg-turbo jstdhuff.c) */
static const uint8_t kStd_dc_lum_bits[17] = {
    0x00, 0x00, 0x01, 0x05, 0x01, 0x01, 0x01, 0x01, 0x01, 0x01, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00,
};
static const uint8_t kStd_dc_lum_vals[12] = {
    0x00, 0x01, 0x02, 0x03, 0x04, 0x05, 0x06, 0x07, 0x08, 0x09, 0x0a, 0x0b,
};
static const uint8_t kStd_dc_chrom_bits[17] = {
    0x00, 0x00, 0x03, 0x01, 0x01, 0x01, 0x01, 0x01, 0x01, 0x01, 0x01, 0x01, 0x00, 0x00, 0x00, 0x00, 0x00,
};
static const uint8_t kStd_dc_chrom_vals[12] = {
    0x00, 0x01, 0x02, 0x03, 0x04, 0x05, 0x06, 0x07, 0x08, 0x09, 0x0a, 0x0b,
};
static const uint8_t kStd_ac_lum_bits[17] = {
    0x00, 0x00, 0x02, 0x01, 0x03, 0x03, 0x02, 0x04, 0x03, 0x05, 0x05, 0x04, 0x04, 0x00, 0x00, 0x01, 0x7d,
};
static const uint8_t kStd_ac_lum_vals[162] = {
    0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71, 0x14, 0x32, 0x81, 0x91, 0xa1, 0x08,
    0x23, 0x42, 0xb1, 0xc1, 0x15, 0x52, 0xd1, 0xf0, 0x24, 0x33, 0x62, 0x72, 0x82, 0x09, 0x0a, 0x16, 0x17, 0x18, 0x19, 0x1a, 0x25, 0x26, 0x27, 0x28,
    0x29, 0x2a, 0x34, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59,
    0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89,
    0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6,
    0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe1, 0xe2,
    0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa,
};
static const uint8_t kStd_ac_chrom_bits[17] = {
    0x00, 0x00, 0x02, 0x01, 0x02, 0x04, 0x04, 0x03, 0x04, 0x07, 0x05, 0x04, 0x04, 0x00, 0x01, 0x02, 0x77,
};
static const uint8_t kStd_ac_chrom_vals[162] = {
    0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22, 0x32, 0x81, 0x08, 0x14, 0x42, 0x91,
    0xa1, 0xb1, 0xc1, 0x09, 0x23, 0x33, 0x52, 0xf0, 0x15, 0x62, 0x72, 0xd1, 0x0a, 0x16, 0x24, 0x34, 0xe1, 0x25, 0xf1, 0x17, 0x18, 0x19, 0x1a, 0x26,
    0x27, 0x28, 0x29, 0x2a, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58,
    0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x82, 0x83, 0x84, 0x85, 0x86, 0x87,
    0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4,
    0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda,
    0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa,
};

// bits[1..16] = codes per length, vals = symbols in code order  ->  the decoder's tables (see HuffDev)
inline bool build_huff(const uint8_t *bits, const uint8_t *vals, int nvals, uint32_t is_ac, HuffDev *h)
{
    memset(h, 0, sizeof *h);
    memcpy(h->vals, vals, (size_t)nvals);
    h->is_ac = is_ac;
    const uint32_t none = make_entry(0, 16, is_ac); // a code that does not exist
    for (uint32_t &e : h->lut) e = none;
    for (auto &t : h->lut2)
        for (uint32_t &e : t) e = none;
    int code = 0, k = 0, ntab2 = 0;
    for (int l = 1; l <= 16; l++) {
        h->valoff[l] = k - code;
        if (bits[l]) {
            if (code + bits[l] > (1 << l)) return false; // over-subscribed
            for (int c = 0; c < bits[l]; c++) {
                const uint32_t entry = make_entry(vals[k + c], (uint32_t)l, is_ac);
                if (l <= kLutBits) {
                    const int first = (code + c) << (kLutBits - l);
                    for (int f = 0; f < (1 << (kLutBits - l)); f++) h->lut[first + f] = entry;
                } else {
                    const int prefix = (code + c) >> (l - kLutBits);
                    uint32_t &e1 = h->lut[prefix];
                    if (e1 == none) e1 = ntab2 < kLut2Tables ? (kEntryLevel2 | (uint32_t)ntab2++) : kEntrySlow;
                    if (e1 & kEntryLevel2) {
                        const int rest = ((code + c) << (16 - l)) & ((1 << kLut2Bits) - 1);
                        for (int f = 0; f < (1 << (16 - l)); f++) h->lut2[e1 & 0xff][rest + f] = entry;
                    } // else: no second-level table left for this prefix, the decoder walks maxcode[]
                }
            }
            code += bits[l];
            k += bits[l];
            h->maxcode[l] = code - 1;
        } else {
            h->maxcode[l] = -1;
        }
        code <<= 1;
    }
    h->maxcode[0] = -1;
    return k == nvals;
}

inline ParseStatus parse(const uint8_t *d, size_t n, uint32_t sub_bits, Parsed *out)
{
    if (n < 4 || d[0] != 0xFF || d[1] != 0xD8) return kParseNotJpeg;
    out->restart_interval = 0;
    struct RawHuff {
        uint8_t bits[17];
        uint8_t vals[256];
        int nvals;
        bool present;
    } huff[2][4];
    uint16_t q[4][64];
    bool qpresent[4] = {false, false, false, false};
    memset(huff, 0, sizeof huff);
    // A camera's MJPG frame may carry no DHT segment at all: the standard tables are implied (libjpeg-turbo installs them
    // for a scan whose tables are missing, jdhuff.c / jstdhuff.c; OpenCV decodes such frames).  Table 0 luma, 1 chroma.
    auto set_std = [&](int tc, int th, const uint8_t *bits, const uint8_t *vals, int n) {
        memcpy(huff[tc][th].bits, bits, 17);
        memcpy(huff[tc][th].vals, vals, (size_t)n);
        huff[tc][th].nvals = n;
        huff[tc][th].present = true;
    };
    set_std(0, 0, kStd_dc_lum_bits, kStd_dc_lum_vals, (int)sizeof kStd_dc_lum_vals);
    set_std(0, 1, kStd_dc_chrom_bits, kStd_dc_chrom_vals, (int)sizeof kStd_dc_chrom_vals);
    set_std(1, 0, kStd_ac_lum_bits, kStd_ac_lum_vals, (int)sizeof kStd_ac_lum_vals);
    set_std(1, 1, kStd_ac_chrom_bits, kStd_ac_chrom_vals, (int)sizeof kStd_ac_chrom_vals);
    int width = 0, height = 0, ncomp = 0, hs[3] = {0, 0, 0}, vs[3] = {0, 0, 0}, tq[3] = {0, 0, 0}, td[3] = {0, 0, 0}, ta[3] = {0, 0, 0};
    int cid[3] = {0, 0, 0};
    bool sof = false, sos = false;
    size_t i = 2;
    while (i + 4 <= n) {
        if (d[i] != 0xFF) return kParseNotJpeg;
        while (i < n && d[i] == 0xFF) i++;
        if (i >= n) return kParseNotJpeg;
        const int m = d[i++];
        if (m == 0xD8 || m == 0x01 || (m >= 0xD0 && m <= 0xD7)) continue;
        if (m == 0xD9) return kParseNotJpeg;
        if (i + 2 > n) return kParseNotJpeg;
        const size_t L = ((size_t)d[i] << 8) | d[i + 1];
        if (L < 2 || i + L > n) return kParseNotJpeg;
        const uint8_t *p = d + i + 2;
        const size_t pl = L - 2;
        if (m == 0xDB) {
            size_t o = 0;
            while (o < pl) {
                const int pq = p[o] >> 4, t = p[o] & 15;
                o++;
                if (t > 3 || o + (pq ? 128u : 64u) > pl) return kParseNotJpeg;
                for (int k = 0; k < 64; k++) {
                    q[t][k] = pq ? (uint16_t)((p[o] << 8) | p[o + 1]) : p[o];
                    o += pq ? 2 : 1;
                }
                qpresent[t] = true;
            }
        } else if (m == 0xC4) {
            size_t o = 0;
            while (o < pl) {
                if (o + 17 > pl) return kParseNotJpeg;
                const int tc = p[o] >> 4, th = p[o] & 15;
                if (tc > 1 || th > 3) return kParseNotJpeg;
                RawHuff &h = huff[tc][th];
                int cnt = 0;
                h.bits[0] = 0;
                for (int l = 1; l <= 16; l++) {
                    h.bits[l] = p[o + l];
                    cnt += p[o + l];
                }
                o += 17;
                if (cnt > 256 || o + (size_t)cnt > pl) return kParseNotJpeg;
                memcpy(h.vals, p + o, (size_t)cnt);
                h.nvals = cnt;
                h.present = true;
                o += (size_t)cnt;
            }
        } else if (m == 0xC0 || m == 0xC1) {
            if (pl < 6) return kParseNotJpeg;
            if (p[0] != 8) return kParseUnsupported;
            height = (p[1] << 8) | p[2];
            width = (p[3] << 8) | p[4];
            ncomp = p[5];
            if (ncomp != 3 && ncomp != 1) return kParseUnsupported;
            if (pl < 6 + 3 * (size_t)ncomp) return kParseNotJpeg;
            for (int c = 0; c < ncomp; c++) {
                cid[c] = p[6 + 3 * c];
                hs[c] = p[7 + 3 * c] >> 4;
                vs[c] = p[7 + 3 * c] & 15;
                tq[c] = p[8 + 3 * c];
                if (tq[c] > 3) return kParseNotJpeg;
            }
            sof = true;
        } else if (m >= 0xC2 && m <= 0xCF && m != 0xC4 && m != 0xC8 && m != 0xCC) {
            return kParseUnsupported; // progressive, lossless, arithmetic coding
        } else if (m == 0xDD) {
            if (pl < 2) return kParseNotJpeg;
            out->restart_interval = (uint32_t)((p[0] << 8) | p[1]);
        } else if (m == 0xDA) {
            if (!sof || pl < 1 || p[0] != ncomp || pl < 1 + 2 * (size_t)ncomp + 3) return sof ? kParseUnsupported : kParseNotJpeg;
            for (int c = 0; c < ncomp; c++) {
                if (p[1 + 2 * c] != cid[c]) return kParseUnsupported; // scan components in another order than the frame's
                td[c] = p[2 + 2 * c] >> 4;
                ta[c] = p[2 + 2 * c] & 15;
                if (td[c] > 3 || ta[c] > 3) return kParseNotJpeg;
            }
            out->scan_offset = i + L;
            sos = true;
            break;
        }
        i += L;
    }
    if (!sos || width <= 0 || height <= 0) return kParseNotJpeg;
    if (ncomp == 3) {
        if (hs[1] != 1 || vs[1] != 1 || hs[2] != 1 || vs[2] != 1) return kParseUnsupported;
        if (!((hs[0] == 1 && vs[0] == 1) || (hs[0] == 2 && vs[0] == 1) || (hs[0] == 2 && vs[0] == 2))) return kParseUnsupported;
    } else {
        hs[0] = vs[0] = 1;
    }
    // the entropy-coded segment ends at the EOI marker: the last FF D9 of the buffer (cameras pad behind it at most)
    size_t end = n;
    while (end >= out->scan_offset + 2 && !(d[end - 2] == 0xFF && d[end - 1] == 0xD9)) end--;
    if (end < out->scan_offset + 2) return kParseNotJpeg;
    out->scan_bytes = end - 2 - out->scan_offset;
    if (out->scan_bytes == 0 || out->scan_bytes > 0x1fffffffu) return kParseNotJpeg;

    Geometry &g = out->g;
    g.width = width;
    g.height = height;
    g.H = hs[0];
    g.V = vs[0];
    g.ncomp = ncomp;
    g.bpm = hs[0] * vs[0] + (ncomp == 3 ? 2 : 0);
    g.mcux = (width + 8 * g.H - 1) / (8 * g.H);
    g.mcuy = (height + 8 * g.V - 1) / (8 * g.V);
    g.nblocks = (uint32_t)g.mcux * (uint32_t)g.mcuy * (uint32_t)g.bpm;
    g.sub_bits = sub_bits;
    g.nsub_max = (uint32_t)((out->scan_bytes * 8 + sub_bits - 1) / sub_bits);
    memset(&out->t, 0, sizeof out->t);
    for (int c = 0; c < ncomp; c++) {
        if (!qpresent[tq[c]] || !huff[0][td[c]].present || !huff[1][ta[c]].present) return kParseNotJpeg;
        if (!build_huff(huff[0][td[c]].bits, huff[0][td[c]].vals, huff[0][td[c]].nvals, 0u, &out->t.h[c][0])) return kParseNotJpeg;
        if (!build_huff(huff[1][ta[c]].bits, huff[1][ta[c]].vals, huff[1][ta[c]].nvals, 1u, &out->t.h[c][1])) return kParseNotJpeg;
        for (int k = 0; k < 64; k++) out->t.q[c][kZigzagNatural[k]] = q[tq[c]][k];
    }
    return kParseOk;
}

// A camera repeats the same header (tables, geometry) in front of every frame: when the bytes up to the scan are those of
// the frame `prev` was parsed from, only the length of the entropy-coded segment has to be found again.
inline bool reparse_same_header(const uint8_t *d, size_t n, const uint8_t *prev_header, size_t prev_header_bytes, uint32_t sub_bits,
                                Parsed *prev)
{
    if (!prev_header_bytes || n < prev_header_bytes + 2 || prev->g.sub_bits != sub_bits || memcmp(d, prev_header, prev_header_bytes) != 0)
        return false;
    size_t end = n;
    while (end >= prev->scan_offset + 2 && !(d[end - 2] == 0xFF && d[end - 1] == 0xD9)) end--;
    if (end < prev->scan_offset + 2) return false;
    prev->scan_bytes = end - 2 - prev->scan_offset;
    if (prev->scan_bytes == 0 || prev->scan_bytes > 0x1fffffffu) return false;
    prev->g.nsub_max = (uint32_t)((prev->scan_bytes * 8 + sub_bits - 1) / sub_bits);
    return true;
}

} // namespace jpg
} // namespace cvs
