// cvs_jpeg.cuh -- capture-side decode on the GPU: baseline JPEG (the camera's MJPG frames, server/src/threads.cpp:32-41)
// -> BGR24 frame in device memory, bit for bit what OpenCV's libjpeg-turbo makes of the same bitstream (JDCT_ISLOW,
// fancy chroma upsampling, fixed-point YCbCr -> RGB), i.e. what the reference's hot path is fed.  SURVEY 8(f) row 4.
//
// A camera frame has no restart markers, so its Huffman stream is one sequential dependency chain (432 KB at 1080p).
// It is decoded in parallel through the self-synchronisation of Huffman codes:
//
//   k_unstuff_count / k_unstuff_write   FF 00 -> FF: the entropy-coded segment becomes a plain bit string of T bits
//   k_entropy (one cooperative launch)  the bit string is cut into subsequences of S bits, one thread each.
//       sync     every thread decodes its subsequence from a guessed entry state (bit offset of the first token that
//                starts in it, block-in-MCU phase, zig-zag position) and hands its exit state to its successor; whoever
//                receives an entry state it has not used yet decodes again.  Thread 0's entry state is exact, a wrong
//                guess falls into step with the true token sequence after a few dozen bits, so the states stop
//                changing after a handful of rounds (any number is handled: a grid barrier per round, until no state
//                changed).  Each run also leaves the blocks completed and the DC differences summed per component.
//       scan     exclusive prefix sums of those give every subsequence its first block index and DC predictors
//       write    each thread decodes once more from its now exact entry state and stores the coefficients
//   k_idct      dequantisation + jidctint.c's accurate integer IDCT per 8x8 block -> Y / Cb / Cr planes
//   k_colour    jdsample.c's h2v1 / h2v2 "fancy" upsampling (context rows replicated at the border) + jdcolor.c's
//               fixed-point conversion, stored B, G, R
//
// The arithmetic of every stage is integer; tests compare with the CPU oracle (oracle/jpeg_oracle.c, pinned against cv2)
// and with the digests of cv2's pixels of the reference's own camera frames.
//
// The token-level functions are __host__ __device__ so that tests/host/jpeg_sim.cpp can run the very same
// synchronisation logic on the CPU (g++, no GPU).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define CVS_HD __host__ __device__ __forceinline__
#else
#define CVS_HD inline
#endif

namespace cvs {
namespace jpg {

constexpr int kLutBits = 9;
constexpr int kMaxRounds = 4096; // sync rounds before the decode is declared failed (a real frame needs < 10)

// canonical Huffman table as the decoder wants it: kLutBits of look-ahead resolve every code of that length or shorter
// (entry = length << 8 | symbol; 0: longer code, walk maxcode[] as jdhuff.c's slow path does)
struct HuffDev {
    uint16_t lut[1 << kLutBits];
    int32_t maxcode[17]; // largest code of length l (1..16), -1: none
    int32_t valoff[17];  // valptr[l] - mincode[l]
    uint8_t vals[256];
};

struct Tables {
    HuffDev h[3][2];   // per component: [0] DC table, [1] AC table
    uint16_t q[3][64]; // per component: quantisation table in natural (row-major) order
};

struct Geometry {
    int width, height;   // image
    int H, V;            // luma sampling factors (chroma is 1x1): 1x1, 2x1, 2x2
    int ncomp;           // 3, or 1 (gray: one block per MCU, B = G = R = Y)
    int bpm;             // blocks per MCU: H*V + 2 (or 1)
    int mcux, mcuy;
    uint32_t nblocks;    // mcux * mcuy * bpm
    uint32_t sub_bits;   // S: bits per subsequence (multiple of 32)
    uint32_t nsub_max;   // subsequences the launch covers (from the raw length; the unstuffed string may be shorter)
};

// ---- entry / exit state of a subsequence, packed: bit offset of the first own token (0..31) | phase << 5 | k << 8 -------
CVS_HD uint32_t pack_state(uint32_t off, uint32_t ph, uint32_t k) { return off | (ph << 5) | (k << 8); }
constexpr uint32_t kStateUnset = 0xffffffffu;

// what one run over a subsequence leaves behind
struct RunResult {
    uint32_t exit_state;
    uint32_t nblocks; // blocks completed by tokens that start in this subsequence
    int32_t dcsum[3]; // DC differences decoded in this subsequence, per component
};

CVS_HD uint32_t bswap32(uint32_t x)
{
#if defined(__CUDA_ARCH__)
    return __byte_perm(x, 0u, 0x0123u);
#else
    return __builtin_bswap32(x);
#endif
}

// 32 bits of the unstuffed string starting at bit p (big-endian bit order), through a two-word register window
struct BitWindow {
    const uint32_t *words;
    uint32_t widx, hi, lo;
    CVS_HD void init(const uint32_t *w, uint32_t p)
    {
        words = w;
        widx = p >> 5;
        hi = bswap32(words[widx]);
        lo = bswap32(words[widx + 1]);
    }
    CVS_HD uint32_t peek(uint32_t p)
    {
        const uint32_t wi = p >> 5;
        if (wi != widx) {
            hi = (wi == widx + 1) ? lo : bswap32(words[wi]);
            lo = bswap32(words[wi + 1]);
            widx = wi;
        }
        const uint32_t s = p & 31u;
        return s ? ((hi << s) | (lo >> (32u - s))) : hi;
    }
};

// jdhuff.c HUFF_EXTEND
CVS_HD int32_t huff_extend(uint32_t x, uint32_t s) { return x < (1u << (s - 1)) ? (int32_t)x - (int32_t)((1u << s) - 1u) : (int32_t)x; }

CVS_HD void huff_decode(const HuffDev &h, uint32_t w, uint32_t &len, uint32_t &sym)
{
    const uint32_t e = h.lut[w >> (32 - kLutBits)];
    if (e) {
        len = e >> 8;
        sym = e & 255u;
        return;
    }
    for (uint32_t l = kLutBits + 1; l <= 16; l++) {
        const int32_t code = (int32_t)(w >> (32u - l));
        if (code <= h.maxcode[l]) {
            len = l;
            sym = h.vals[(uint32_t)(code + h.valoff[l]) & 255u];
            return;
        }
    }
    len = 16; // corrupt data: no such code (libjpeg warns and uses a zero symbol)
    sym = 0;
}

// One run over subsequence i: tokens that start in [i*S + off, min((i+1)*S, T)).  WRITE: store the coefficients
// (scan order, natural order inside a block, DC absolute) of blocks first_block.. with the predictors pred[].
template <bool WRITE>
CVS_HD RunResult run_subsequence(const Tables &tb, const Geometry &g, const uint32_t *words, uint32_t total_bits, uint32_t i,
                                 uint32_t entry, const uint8_t *natural, int16_t *coef, uint32_t first_block, int32_t pred0,
                                 int32_t pred1, int32_t pred2)
{
    RunResult r;
    r.nblocks = 0;
    r.dcsum[0] = r.dcsum[1] = r.dcsum[2] = 0;
    const uint32_t S = g.sub_bits;
    const uint32_t begin = i * S, end_nominal = begin + S;
    const uint32_t end = end_nominal < total_bits ? end_nominal : total_bits;
    uint32_t p = begin + (entry & 31u), ph = (entry >> 5) & 7u, k = (entry >> 8) & 63u;
    int32_t pred[3] = {pred0, pred1, pred2};
    uint32_t blk = first_block;
    const uint32_t nluma = (uint32_t)(g.bpm - (g.ncomp == 3 ? 2 : 0));
    if (begin < total_bits) {
        BitWindow bw;
        bw.init(words, p);
        while (p < end) {
            const uint32_t c = ph < nluma ? 0u : ph - nluma + 1u;
            const uint32_t w = bw.peek(p);
            uint32_t len, sym;
            if (k == 0) {
                huff_decode(tb.h[c][0], w, len, sym);
                const uint32_t s = sym & 15u;
                int32_t v = 0;
                if (s) v = huff_extend((w << len) >> (32u - s), s);
                p += len + s;
                if (WRITE) {
                    pred[c] += v;
                    if (blk < g.nblocks) coef[(size_t)blk * 64] = (int16_t)pred[c];
                } else {
                    r.dcsum[c] += v;
                }
                k = 1;
            } else {
                huff_decode(tb.h[c][1], w, len, sym);
                const uint32_t run = sym >> 4, s = sym & 15u;
                if (s) {
                    k += run;
                    if (WRITE) {
                        const int32_t v = huff_extend((w << len) >> (32u - s), s);
                        if (k < 64 && blk < g.nblocks) coef[(size_t)blk * 64 + natural[k]] = (int16_t)v;
                    }
                    p += len + s;
                    k++;
                } else {
                    p += len;
                    k = run == 15 ? k + 16 : 64;
                }
                if (k >= 64) {
                    k = 0;
                    ph = ph + 1 == (uint32_t)g.bpm ? 0u : ph + 1;
                    blk++;
                    r.nblocks++;
                }
            }
        }
    }
    const uint32_t over = p > end_nominal ? p - end_nominal : 0u; // < 32: a token is at most 31 bits long
    r.exit_state = pack_state(over & 31u, ph, k);
    return r;
}

} // namespace jpg
} // namespace cvs

// =====================================================================================================================
#if defined(__CUDACC__)
#include <cooperative_groups.h>

namespace cvs {
namespace jpg {

__constant__ uint8_t c_natural[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                      41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                      30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

constexpr int kUnstuffThreads = 256, kUnstuffBytes = 16; // bytes per thread
constexpr int kEntropyThreads = 256;

// status bits of a decode (OR-ed into the ticket's status word)
constexpr unsigned int kJpegNotConverged = 1u << 8, kJpegBlockCount = 1u << 9;

// ---- byte unstuffing ------------------------------------------------------------------------------------------------
// a byte is dropped when it is the 00 behind an FF (T.81 B.1.1.5); raw_len is the length of the entropy-coded segment
__device__ __forceinline__ uint32_t unstuff_keep_mask(const uint8_t *raw, uint32_t raw_len, uint32_t first, uint8_t (&b)[kUnstuffBytes])
{
    uint32_t keep = 0;
    uint8_t prev = first ? raw[first - 1] : 0;
    if (first + kUnstuffBytes <= raw_len) {
        const uint4 v = *reinterpret_cast<const uint4 *>(raw + first);
        const uint32_t wv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < kUnstuffBytes; j++) b[j] = (uint8_t)(wv[j >> 2] >> (8 * (j & 3)));
    } else {
#pragma unroll
        for (int j = 0; j < kUnstuffBytes; j++) b[j] = first + j < raw_len ? raw[first + j] : 0;
    }
#pragma unroll
    for (int j = 0; j < kUnstuffBytes; j++) {
        if (first + j < raw_len && !(b[j] == 0x00 && prev == 0xFF)) keep |= 1u << j;
        prev = b[j];
    }
    return keep;
}

__global__ void __launch_bounds__(kUnstuffThreads) k_unstuff_count(const uint8_t *__restrict__ raw, uint32_t raw_len, uint32_t *block_kept)
{
    __shared__ uint32_t wsum[kUnstuffThreads / 32];
    const uint32_t first = (blockIdx.x * kUnstuffThreads + threadIdx.x) * kUnstuffBytes;
    uint8_t b[kUnstuffBytes];
    uint32_t n = first < raw_len ? (uint32_t)__popc(unstuff_keep_mask(raw, raw_len, first, b)) : 0u;
    n = __reduce_add_sync(0xffffffffu, n);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = n;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < kUnstuffThreads / 32; w++) t += wsum[w];
        block_kept[blockIdx.x] = t;
    }
}

// out must be zero beyond the string (the decoder looks up to 8 bytes past its end); total_bits <- 8 * kept bytes
__global__ void __launch_bounds__(kUnstuffThreads) k_unstuff_write(const uint8_t *__restrict__ raw, uint32_t raw_len,
                                                                  const uint32_t *__restrict__ block_kept, uint8_t *out,
                                                                  uint32_t *total_bits)
{
    __shared__ uint32_t wsum[kUnstuffThreads / 32];
    __shared__ uint32_t s_base;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // bytes kept by the blocks in front of this one
    uint32_t part = 0;
    for (uint32_t j = tid; j < blockIdx.x; j += kUnstuffThreads) part += block_kept[j];
    part = __reduce_add_sync(0xffffffffu, part);
    if (lane == 0) wsum[warp] = part;
    __syncthreads();
    if (tid == 0) {
        uint32_t t = 0;
        for (int w = 0; w < kUnstuffThreads / 32; w++) t += wsum[w];
        s_base = t;
    }
    __syncthreads();
    const uint32_t base = s_base;
    __syncthreads();
    const uint32_t first = (blockIdx.x * kUnstuffThreads + tid) * kUnstuffBytes;
    uint8_t b[kUnstuffBytes];
    const uint32_t keep = first < raw_len ? unstuff_keep_mask(raw, raw_len, first, b) : 0u;
    const uint32_t n = (uint32_t)__popc(keep);
    uint32_t incl = n;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= (uint32_t)d) incl += o;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    uint32_t woff = 0;
    for (uint32_t w = 0; w < warp; w++) woff += wsum[w];
    uint32_t o = base + woff + incl - n;
    if (keep == 0xffffu && (o & 3u) == 0) { // the common case: nothing dropped, word-aligned destination
        uint32_t *dst = reinterpret_cast<uint32_t *>(out + o);
#pragma unroll
        for (int q = 0; q < 4; q++)
            dst[q] = (uint32_t)b[4 * q] | ((uint32_t)b[4 * q + 1] << 8) | ((uint32_t)b[4 * q + 2] << 16) | ((uint32_t)b[4 * q + 3] << 24);
    } else {
#pragma unroll
        for (int j = 0; j < kUnstuffBytes; j++)
            if (keep >> j & 1u) out[o++] = b[j];
    }
    if (blockIdx.x == gridDim.x - 1 && tid == kUnstuffThreads - 1) {
        uint32_t t = 0;
        for (int w = 0; w < kUnstuffThreads / 32; w++) t += wsum[w];
        *total_bits = 8u * (base + t);
    }
}

// ---- entropy decode -------------------------------------------------------------------------------------------------
struct EntropyParams {
    const Tables *tables;     // device copy
    Geometry g;
    const uint32_t *words;    // unstuffed string
    const uint32_t *total_bits;
    uint32_t *entry;          // [nsub_max + 1] entry state of each subsequence (written by its predecessor)
    uint32_t *used;           // [nsub_max] entry state of the last run
    uint32_t *nblk;           // [nsub_max] blocks completed, then (after the scan) first block index
    int32_t *dcs;             // [3][nsub_max] DC sums, then predictors at entry
    uint32_t *tile_blk;       // [ntiles] per-tile totals
    int32_t *tile_dc;         // [3][ntiles]
    unsigned int *changed;    // [kMaxRounds] states changed per round
    int16_t *coef;            // [nblocks][64], zeroed
    unsigned int *status;
};

__global__ void __launch_bounds__(kEntropyThreads) k_entropy(const EntropyParams p)
{
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    __shared__ Tables tb;
    __shared__ uint32_t s_scan[4][kEntropyThreads / 32];
    __shared__ uint32_t s_tile_base[4];
    __shared__ uint8_t s_nat[64];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 64) s_nat[tid] = c_natural[tid];
    {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(p.tables);
        uint32_t *dst = reinterpret_cast<uint32_t *>(&tb);
        for (uint32_t j = tid; j < sizeof(Tables) / 4; j += kEntropyThreads) dst[j] = src[j];
    }
    __syncthreads();
    const Geometry g = p.g;
    const uint32_t T = *p.total_bits;
    const uint32_t nsub = (T + g.sub_bits - 1) / g.sub_bits; // <= nsub_max
    const uint32_t ntiles = (nsub + kEntropyThreads - 1) / kEntropyThreads;

    // ---- sync rounds
    bool converged = false;
    for (uint32_t round = 0; round < (uint32_t)kMaxRounds; round++) {
        for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const uint32_t i = tile * kEntropyThreads + tid;
            if (i >= nsub) continue;
            uint32_t e = i == 0 ? pack_state(0, 0, 0) : __ldcg(p.entry + i);
            if (round == 0 && i) e = pack_state(0, 0, 0); // first guess: a block starts here
            if (round && e == p.used[i]) continue;
            const RunResult r = run_subsequence<false>(tb, g, p.words, T, i, e, s_nat, nullptr, 0, 0, 0, 0);
            p.used[i] = e;
            p.nblk[i] = r.nblocks;
            p.dcs[i] = r.dcsum[0];
            p.dcs[g.nsub_max + i] = r.dcsum[1];
            p.dcs[2 * g.nsub_max + i] = r.dcsum[2];
            if (round == 0 || __ldcg(p.entry + i + 1) != r.exit_state) {
                __stcg(p.entry + i + 1, r.exit_state);
                if (round && i + 1 < nsub) atomicAdd(p.changed + round, 1u);
            }
        }
        grid.sync();
        if (round && __ldcg(p.changed + round) == 0) {
            converged = true;
            break;
        }
    }
    if (!converged) {
        if (blockIdx.x == 0 && tid == 0) atomicOr(p.status, kJpegNotConverged);
        return;
    }

    // ---- scan: per-tile totals, then every tile sums the totals in front of it
    for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const uint32_t i = tile * kEntropyThreads + tid;
        uint32_t v[4] = {0, 0, 0, 0};
        if (i < nsub) {
            v[0] = p.nblk[i];
            v[1] = (uint32_t)p.dcs[i];
            v[2] = (uint32_t)p.dcs[g.nsub_max + i];
            v[3] = (uint32_t)p.dcs[2 * g.nsub_max + i];
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const uint32_t s = __reduce_add_sync(0xffffffffu, v[q]);
            if (lane == 0) s_scan[q][warp] = s;
        }
        __syncthreads();
        if (tid < 4) {
            uint32_t t = 0;
            for (int w = 0; w < kEntropyThreads / 32; w++) t += s_scan[tid][w];
            if (tid == 0) p.tile_blk[tile] = t;
            else p.tile_dc[(tid - 1) * ntiles + tile] = (int32_t)t;
        }
        __syncthreads();
    }
    grid.sync();
    uint32_t total_blocks_seen = 0;
    for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        // base of this tile
        uint32_t part[4] = {0, 0, 0, 0};
        for (uint32_t j = tid; j < tile; j += kEntropyThreads) {
            part[0] += __ldcg(p.tile_blk + j);
            part[1] += (uint32_t)__ldcg(p.tile_dc + j);
            part[2] += (uint32_t)__ldcg(p.tile_dc + ntiles + j);
            part[3] += (uint32_t)__ldcg(p.tile_dc + 2 * ntiles + j);
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const uint32_t s = __reduce_add_sync(0xffffffffu, part[q]);
            if (lane == 0) s_scan[q][warp] = s;
        }
        __syncthreads();
        if (tid < 4) {
            uint32_t t = 0;
            for (int w = 0; w < kEntropyThreads / 32; w++) t += s_scan[tid][w];
            s_tile_base[tid] = t;
        }
        __syncthreads();
        uint32_t base[4] = {s_tile_base[0], s_tile_base[1], s_tile_base[2], s_tile_base[3]};
        __syncthreads();
        // exclusive scan inside the tile
        const uint32_t i = tile * kEntropyThreads + tid;
        uint32_t v[4] = {0, 0, 0, 0};
        if (i < nsub) {
            v[0] = p.nblk[i];
            v[1] = (uint32_t)p.dcs[i];
            v[2] = (uint32_t)p.dcs[g.nsub_max + i];
            v[3] = (uint32_t)p.dcs[2 * g.nsub_max + i];
        }
        uint32_t incl[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            incl[q] = v[q];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(0xffffffffu, incl[q], d);
                if (lane >= (uint32_t)d) incl[q] += o;
            }
            if (lane == 31) s_scan[q][warp] = incl[q];
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 4; q++) {
            uint32_t woff = 0;
            for (uint32_t w = 0; w < warp; w++) woff += s_scan[q][w];
            base[q] += woff + incl[q] - v[q];
        }
        __syncthreads();
        // ---- write pass
        if (i < nsub) {
            const RunResult r = run_subsequence<true>(tb, g, p.words, T, i, p.used[i], s_nat, p.coef, base[0], (int32_t)base[1],
                                                      (int32_t)base[2], (int32_t)base[3]);
            if (i == nsub - 1) total_blocks_seen = base[0] + r.nblocks;
        }
    }
    // the string must hold exactly the image's blocks (padding bits after the last block decode to no complete block
    // in a well-formed stream; more or fewer blocks means a damaged frame)
    if (total_blocks_seen && total_blocks_seen < g.nblocks) atomicOr(p.status, kJpegBlockCount);
}

// ---- IDCT: jidctint.c jpeg_idct_islow, one thread per 8x8 block ---------------------------------------------------------
struct PlaneParams {
    Geometry g;
    const Tables *tables;
    const int16_t *coef;
    uint8_t *y, *cb, *cr; // planes: luma pitch = mcux*8*H, chroma pitch = mcux*8
};

#define CVS_JPG_FIX_0_298631336 2446
#define CVS_JPG_FIX_0_390180644 3196
#define CVS_JPG_FIX_0_541196100 4433
#define CVS_JPG_FIX_0_765366865 6270
#define CVS_JPG_FIX_0_899976223 7373
#define CVS_JPG_FIX_1_175875602 9633
#define CVS_JPG_FIX_1_501321110 12299
#define CVS_JPG_FIX_1_847759065 15137
#define CVS_JPG_FIX_1_961570560 16069
#define CVS_JPG_FIX_2_053119869 16819
#define CVS_JPG_FIX_2_562915447 20995
#define CVS_JPG_FIX_3_072711026 25172

// the 1-D transform of both passes: eight inputs -> eight outputs descaled by SHIFT (jidctint.c:209-283 / :312-383)
template <int SHIFT>
__device__ __forceinline__ void idct8(const int32_t (&in)[8], int32_t (&out)[8])
{
    int32_t z2 = in[2], z3 = in[6];
    int32_t z1 = (z2 + z3) * CVS_JPG_FIX_0_541196100;
    int32_t tmp2 = z1 + z3 * (-CVS_JPG_FIX_1_847759065);
    int32_t tmp3 = z1 + z2 * CVS_JPG_FIX_0_765366865;
    int32_t tmp0 = (int32_t)((uint32_t)(in[0] + in[4]) << 13);
    int32_t tmp1 = (int32_t)((uint32_t)(in[0] - in[4]) << 13);
    const int32_t tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
    tmp0 = in[7];
    tmp1 = in[5];
    tmp2 = in[3];
    tmp3 = in[1];
    z1 = tmp0 + tmp3;
    z2 = tmp1 + tmp2;
    z3 = tmp0 + tmp2;
    int32_t z4 = tmp1 + tmp3;
    const int32_t z5 = (z3 + z4) * CVS_JPG_FIX_1_175875602;
    tmp0 *= CVS_JPG_FIX_0_298631336;
    tmp1 *= CVS_JPG_FIX_2_053119869;
    tmp2 *= CVS_JPG_FIX_3_072711026;
    tmp3 *= CVS_JPG_FIX_1_501321110;
    z1 *= -CVS_JPG_FIX_0_899976223;
    z2 *= -CVS_JPG_FIX_2_562915447;
    z3 *= -CVS_JPG_FIX_1_961570560;
    z4 *= -CVS_JPG_FIX_0_390180644;
    z3 += z5;
    z4 += z5;
    tmp0 += z1 + z3;
    tmp1 += z2 + z4;
    tmp2 += z2 + z3;
    tmp3 += z1 + z4;
    constexpr int32_t rnd = 1 << (SHIFT - 1);
    out[0] = (tmp10 + tmp3 + rnd) >> SHIFT;
    out[7] = (tmp10 - tmp3 + rnd) >> SHIFT;
    out[1] = (tmp11 + tmp2 + rnd) >> SHIFT;
    out[6] = (tmp11 - tmp2 + rnd) >> SHIFT;
    out[2] = (tmp12 + tmp1 + rnd) >> SHIFT;
    out[5] = (tmp12 - tmp1 + rnd) >> SHIFT;
    out[3] = (tmp13 + tmp0 + rnd) >> SHIFT;
    out[4] = (tmp13 - tmp0 + rnd) >> SHIFT;
}

// sample_range_limit + CENTERJSAMPLE indexed with x & 1023 (jdmaster.c prepare_range_limit_table)
__device__ __forceinline__ uint32_t idct_range_limit(int32_t x)
{
    const uint32_t i = (uint32_t)x & 1023u;
    return i < 128u ? 128u + i : (i < 512u ? 255u : (i < 896u ? 0u : i - 896u));
}

constexpr int kIdctThreads = 128;

__global__ void __launch_bounds__(kIdctThreads) k_idct(const PlaneParams p)
{
    __shared__ uint16_t sq[3][64];
    for (uint32_t j = threadIdx.x; j < 192; j += kIdctThreads) sq[j >> 6][j & 63] = p.tables->q[j >> 6][j & 63];
    __syncthreads();
    const Geometry g = p.g;
    const uint32_t b = blockIdx.x * kIdctThreads + threadIdx.x;
    if (b >= g.nblocks) return;
    const uint32_t m = b / (uint32_t)g.bpm, j = b - m * (uint32_t)g.bpm;
    const uint32_t my = m / (uint32_t)g.mcux, mx = m - my * (uint32_t)g.mcux;
    const uint32_t nluma = (uint32_t)(g.bpm - (g.ncomp == 3 ? 2 : 0));
    uint32_t c;
    uint8_t *dst;
    size_t pitch;
    if (j < nluma) {
        c = 0;
        pitch = (size_t)g.mcux * 8 * g.H;
        dst = p.y + ((size_t)my * g.V + j / (uint32_t)g.H) * 8 * pitch + ((size_t)mx * g.H + j % (uint32_t)g.H) * 8;
    } else {
        c = j - nluma + 1;
        pitch = (size_t)g.mcux * 8;
        dst = (c == 1 ? p.cb : p.cr) + (size_t)my * 8 * pitch + (size_t)mx * 8;
    }
    // coefficients of the block: row r in cf[r][0..7]
    int32_t ws[8][8];
    const uint4 *src = reinterpret_cast<const uint4 *>(p.coef + (size_t)b * 64);
#pragma unroll
    for (int r = 0; r < 8; r++) {
        const uint4 v = __ldg(src + r);
        const uint32_t wv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const int32_t cf = (int32_t)(int16_t)(wv[q >> 1] >> (16 * (q & 1)));
            ws[r][q] = cf * (int32_t)sq[c][8 * r + q];
        }
    }
    // pass 1: columns
#pragma unroll
    for (int col = 0; col < 8; col++) {
        int32_t in[8], out[8];
#pragma unroll
        for (int r = 0; r < 8; r++) in[r] = ws[r][col];
        idct8<13 - 2>(in, out);
#pragma unroll
        for (int r = 0; r < 8; r++) ws[r][col] = out[r];
    }
    // pass 2: rows
#pragma unroll
    for (int r = 0; r < 8; r++) {
        int32_t out[8];
        idct8<13 + 2 + 3>(ws[r], out);
        uint32_t lo = 0, hi = 0;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            lo |= idct_range_limit(out[q]) << (8 * q);
            hi |= idct_range_limit(out[4 + q]) << (8 * q);
        }
        *reinterpret_cast<uint2 *>(dst + (size_t)r * pitch) = make_uint2(lo, hi);
    }
}

// ---- chroma upsampling + colour conversion: one thread per 4 horizontally adjacent pixels -------------------------------
struct ColourParams {
    Geometry g;
    const uint8_t *y, *cb, *cr;
    uint8_t *bgr; // width * height * 3, rows contiguous
};

__device__ __forceinline__ uint32_t clamp255(int32_t v) { return (uint32_t)min(max(v, 0), 255); }

constexpr int kColourThreads = 256;

__global__ void __launch_bounds__(kColourThreads) k_colour(const ColourParams p)
{
    const Geometry g = p.g;
    const int x0 = (blockIdx.x * kColourThreads + threadIdx.x) * 4;
    const int yrow = blockIdx.y;
    if (x0 >= g.width) return;
    const size_t ypitch = (size_t)g.mcux * 8 * g.H, cpitch = (size_t)g.mcux * 8;
    const uint32_t yw = *reinterpret_cast<const uint32_t *>(p.y + (size_t)yrow * ypitch + x0);
    int32_t cbv[4], crv[4];
    if (g.ncomp == 1) {
#pragma unroll
        for (int q = 0; q < 4; q++) cbv[q] = crv[q] = 128;
    } else if (g.H == 1) {
        const uint32_t b4 = *reinterpret_cast<const uint32_t *>(p.cb + (size_t)yrow * cpitch + x0);
        const uint32_t r4 = *reinterpret_cast<const uint32_t *>(p.cr + (size_t)yrow * cpitch + x0);
#pragma unroll
        for (int q = 0; q < 4; q++) {
            cbv[q] = (int32_t)((b4 >> (8 * q)) & 255u);
            crv[q] = (int32_t)((r4 >> (8 * q)) & 255u);
        }
    } else {
        const int dw = (g.width + 1) >> 1;            // downsampled_width
        const int c0 = x0 >> 1;                       // chroma columns c0, c0 + 1 cover the four pixels
        const bool fancy = dw > 2;
#pragma unroll
        for (int pl = 0; pl < 2; pl++) {
            const uint8_t *src = pl ? p.cr : p.cb;
            int32_t *dstv = pl ? crv : cbv;
            if (g.V == 1) {
                const uint8_t *in = src + (size_t)yrow * cpitch;
                if (fancy) { // h2v1_fancy_upsample (jdsample.c)
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                        const int c = min(c0 + e, dw - 1);
                        const int v = in[c];
                        dstv[2 * e] = c == 0 ? v : (3 * v + in[c - 1] + 1) >> 2;
                        dstv[2 * e + 1] = c == dw - 1 ? v : (3 * v + in[c + 1] + 2) >> 2;
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < 4; q++) dstv[q] = in[min((x0 + q) >> 1, dw - 1)];
                }
            } else {
                const int dh = (g.height + 1) >> 1;
                const int r = yrow >> 1;
                if (fancy) { // h2v2_fancy_upsample; the rows above the first / below the last are that row again (jdmainct.c)
                    const int rn = min(max((yrow & 1) ? r + 1 : r - 1, 0), dh - 1);
                    const uint8_t *in0 = src + (size_t)r * cpitch, *in1 = src + (size_t)rn * cpitch;
                    // column sums 3 * nearer + further for columns c0 - 1 .. c0 + 2
                    int32_t cs[4];
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        const int c = min(max(c0 - 1 + e, 0), dw - 1);
                        cs[e] = 3 * (int32_t)in0[c] + (int32_t)in1[c];
                    }
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                        const int c = c0 + e;
                        const int32_t t = cs[1 + e];
                        dstv[2 * e] = c == 0 ? (4 * t + 8) >> 4 : (3 * t + cs[e] + 8) >> 4;
                        dstv[2 * e + 1] = c >= dw - 1 ? (4 * t + 7) >> 4 : (3 * t + cs[2 + e] + 7) >> 4;
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < 4; q++) dstv[q] = src[(size_t)r * cpitch + min((x0 + q) >> 1, dw - 1)];
                }
            }
        }
    }
    // jdcolor.c ycc_rgb_convert with the tables of build_ycc_rgb_table written out
    uint32_t px[12];
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int32_t yy = (int32_t)((yw >> (8 * q)) & 255u), cb = cbv[q] - 128, cr = crv[q] - 128;
        if (g.ncomp == 1) {
            px[3 * q] = px[3 * q + 1] = px[3 * q + 2] = (uint32_t)yy;
        } else {
            px[3 * q + 2] = clamp255(yy + ((91881 * cr + 32768) >> 16));
            px[3 * q + 1] = clamp255(yy + ((-22554 * cb + 32768 - 46802 * cr) >> 16));
            px[3 * q] = clamp255(yy + ((116130 * cb + 32768) >> 16));
        }
    }
    uint8_t *dst = p.bgr + ((size_t)yrow * g.width + x0) * 3;
    if (x0 + 4 <= g.width && (((size_t)yrow * g.width * 3) & 3u) == 0 && (reinterpret_cast<uintptr_t>(p.bgr) & 3u) == 0) {
        uint32_t *d4 = reinterpret_cast<uint32_t *>(dst);
#pragma unroll
        for (int q = 0; q < 3; q++) d4[q] = px[4 * q] | (px[4 * q + 1] << 8) | (px[4 * q + 2] << 16) | (px[4 * q + 3] << 24);
    } else {
        const int npx = min(4, g.width - x0);
        for (int j = 0; j < 3 * npx; j++) dst[j] = (uint8_t)px[j];
    }
}

} // namespace jpg
} // namespace cvs
#endif // __CUDACC__
