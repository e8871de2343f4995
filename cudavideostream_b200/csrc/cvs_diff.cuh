// cvs_diff.cuh -- the fused thresholded-difference + negative-feedback + ordered-compaction kernel.
//
// Replaces kernel2 (server/src/kernels.cu:289-334) and its CPU twin (tests/cuda_streaming/
// test.cu:560-576) with one PERSISTENT launch that walks a whole device-resident sequence of
// frames (T = 1 for the drop-in exec_core path):
//
//   * the frame is cut into rows of 1024 B (32 lanes x 32 B).  Every lane moves its 32 B with one
//     LDG.E.256 (sm_100 256-bit global load, L2 evict-first for the frames), so a warp row is one
//     fully coalesced 1 KiB request;
//   * the rows of a frame are partitioned ONCE over the co-resident grid (balanced contiguous
//     ranges), so a thread owns the same <= PPT pieces in every frame.  When the whole frame fits
//     in one pass (REFREG) the reference frame therefore lives in registers for the entire
//     sequence: HBM sees each frame exactly once (N bytes) plus the payload;
//   * per piece: byte-SIMD |cur-ref| > T (VABSDIFF4 + carry-less add), 0x80 flags, per-lane
//     count; warp shuffle scan + a tiny block scan give every changed byte its rank in ascending
//     byte order;
//   * cross-block offsets: each block publishes its count in a 64-bit descriptor
//     (epoch<<32 | count) and sums its predecessors' descriptors -- a one-round, all-threads
//     decoupled look-back (no chained prefix wait);
//   * the (index, value) entries are staged in shared memory in rank order and flushed with
//     coalesced streaming stores; the reference is updated in place only where it changed.
//
// Ordering, values and the new reference are bit-exact with the CPU loop (see oracle/cvs_oracle.c
// orc_diff_compact); unlike kernel2 the payload order is deterministic (ascending xs).
#pragma once
#include "cvs_device.cuh"

namespace cvs {

constexpr int kPieceBytes = 32;                 // one LDG.E.256 per lane
constexpr int kRowBytes = 32 * kPieceBytes;     // one warp request
constexpr int kMaxWarps = 7;                    // block size <= 224 threads (3 blocks/SM at <= 96 registers)
constexpr int kMaxThreads = kMaxWarps * 32;
constexpr int kMaxPPT = 2;
constexpr unsigned kSpinLimit = 1u << 24;       // watchdog for the descriptor spin (never hit in a healthy run)

struct DiffParams {
    const uint8_t *frames;      // frame t at frames + t*frame_stride (32-byte aligned)
    size_t frame_stride;
    int nframes;
    uint8_t *ref;               // reference frame (padded to a whole number of rows)
    uint32_t nbytes;            // N = 3*W*H
    uint32_t rows;              // ceil(N / 1024)
    uint32_t nseg;              // passes per frame (1 => REFREG possible)
    unsigned int *pos;          // [nframes]
    int *xs;                    // frame t at xs + t*cap
    uint8_t *diff;              // frame t at diff + t*cap
    size_t cap;                 // payload capacity per frame (entries)
    unsigned long long *desc;   // [nframes*nseg*gridDim.x]
    uint32_t epoch;             // tag of this launch
    uint32_t addc;              // threshold constant for changed80<>
    uint32_t stage_cap;         // entries the shared staging area holds (>= nwarps*1024)
    unsigned int *status;       // bit0: capacity overflow, bit1: watchdog
};

__device__ __forceinline__ void ld256(const void *p, uint32_t (&w)[8], bool keep)
{
    if (keep)
        asm volatile("ld.global.L1::no_allocate.L2::evict_last.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
                     : "l"(p));
    else
        asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
                     : "l"(p));
}
__device__ __forceinline__ void st256_keep(void *p, const uint32_t (&w)[8])
{
    asm volatile("st.global.L1::no_allocate.L2::evict_last.v8.b32 [%8], {%0,%1,%2,%3,%4,%5,%6,%7};" ::"r"(w[0]),
                 "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]), "l"(p)
                 : "memory");
}

// rows [lo, hi) of block b in pass s
__device__ __forceinline__ void block_rows(const DiffParams &p, uint32_t s, uint32_t b, uint32_t G, uint32_t &lo,
                                           uint32_t &hi)
{
    uint32_t s0 = (uint32_t)(((uint64_t)p.rows * s) / p.nseg);
    uint32_t s1 = (uint32_t)(((uint64_t)p.rows * (s + 1)) / p.nseg);
    uint32_t n = s1 - s0;
    lo = s0 + (uint32_t)(((uint64_t)n * b) / G);
    hi = s0 + (uint32_t)(((uint64_t)n * (b + 1)) / G);
}

// 0x80 flag per changed byte of word k of a piece with nv valid bytes
template <bool HI>
__device__ __forceinline__ uint32_t piece_flags(uint32_t c, uint32_t r, uint32_t addc, uint32_t nv, int k)
{
    uint32_t m = changed80<HI>(absdiff4(c, r), addc);
    if (nv < (uint32_t)kPieceBytes) { // the last piece of a frame whose size is not a multiple of 32
        int vb = (int)nv - 4 * k;
        uint32_t vm = vb >= 4 ? 0x80808080u : (vb <= 0 ? 0u : (0x80808080u & ((1u << (8 * vb)) - 1u)));
        m &= vm;
    }
    return m;
}

template <int PPT, bool HI, bool REFREG>
__global__ void __maxnreg__(96) k_diff_compact(const DiffParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // layout: int sxs[stage_cap] | u8 sd[stage_cap] | small tables (stage_cap is a multiple of 16)
    int *sxs = reinterpret_cast<int *>(smem_raw);
    uint8_t *sd = smem_raw + (size_t)p.stage_cap * 4;
    uint32_t *tab = reinterpret_cast<uint32_t *>(smem_raw + (size_t)p.stage_cap * 5);
    uint32_t *rowcnt = tab;       // [16] changed bytes per row of this block (row order j*nw + warp)
    uint32_t *rowbase = tab + 16; // [17] exclusive scan of rowcnt, [PPT*nw] = block total
    uint32_t *red = tab + 40;     // [8]  predecessors of this pass
    uint32_t *red2 = tab + 48;    // [8]  grand total of the previous pass

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t nw = blockDim.x >> 5, NT = blockDim.x;
    const uint32_t b = blockIdx.x, G = gridDim.x;
    const uint32_t N = p.nbytes;

    uint32_t c[PPT][8], r[PPT][8], cn[PPT][8];
    uint32_t off[PPT]; // byte offset of the piece in the frame
    uint32_t nv[PPT];  // valid bytes in the piece (0..32)
    uint32_t dirty = 0; // REFREG: pieces whose reference changed during this launch

    // piece geometry of pass s (identical in every frame)
    auto geometry = [&](uint32_t s) {
        uint32_t lo, hi;
        block_rows(p, s, b, G, lo, hi);
#pragma unroll
        for (int j = 0; j < PPT; j++) {
            uint32_t row = lo + j * nw + warp;
            uint32_t o = row * kRowBytes + lane * kPieceBytes;
            bool ok = row < hi && o < N;
            off[j] = o;
            nv[j] = ok ? min(N - o, (uint32_t)kPieceBytes) : 0u;
        }
    };
    auto zero8 = [](uint32_t (&w)[8]) {
#pragma unroll
        for (int k = 0; k < 8; k++) w[k] = 0;
    };
    // prefetch the pieces of pass s of frame t into cn[]
    auto prefetch = [&](int t, uint32_t s) {
        uint32_t lo, hi;
        block_rows(p, s, b, G, lo, hi);
        const uint8_t *f = p.frames + (size_t)t * p.frame_stride;
#pragma unroll
        for (int j = 0; j < PPT; j++) {
            uint32_t row = lo + j * nw + warp;
            uint32_t o = row * kRowBytes + lane * kPieceBytes;
            if (row < hi && o < N) ld256(f + o, cn[j], false);
            else zero8(cn[j]);
        }
    };
    auto load_ref = [&]() {
#pragma unroll
        for (int j = 0; j < PPT; j++) {
            if (nv[j]) ld256(p.ref + off[j], r[j], true);
            else zero8(r[j]);
        }
    };

    geometry(0);
    if (REFREG) load_ref();
    if (p.nframes > 0) prefetch(0, 0);

    uint32_t spin_budget = kSpinLimit;

    for (int t = 0; t < p.nframes; t++) {
        uint32_t carry = 0; // entries of this frame emitted by earlier passes (all blocks)
        for (uint32_t s = 0; s < p.nseg; s++) {
            // ---- 1. take the prefetched pieces, start the next prefetch
#pragma unroll
            for (int j = 0; j < PPT; j++)
#pragma unroll
                for (int k = 0; k < 8; k++) c[j][k] = cn[j][k];
            if (!REFREG) load_ref();
            {
                uint32_t s2 = s + 1;
                int t2 = t;
                if (s2 == p.nseg) { s2 = 0; t2 = t + 1; }
                if (t2 < p.nframes) prefetch(t2, s2);
            }

            // ---- 2. flags and per-lane counts
            uint32_t cnt[PPT];
#pragma unroll
            for (int j = 0; j < PPT; j++) {
                uint32_t acc = 0;
#pragma unroll
                for (int k = 0; k < 8; k++) acc += piece_flags<HI>(c[j][k], r[j][k], p.addc, nv[j], k) >> 7;
                cnt[j] = hsum4(acc);
            }

            // ---- 3. ranks: lane-exclusive inside the row, row totals through shared memory
            uint32_t lex[PPT];
#pragma unroll
            for (int j = 0; j < PPT; j++) {
                uint32_t inc = warp_incl_scan(cnt[j], lane);
                lex[j] = inc - cnt[j];
                if (lane == 31) rowcnt[j * nw + warp] = inc;
            }
            __syncthreads();
            const size_t dbase = ((size_t)t * p.nseg + s) * G;
            if (warp == 0) {
                const uint32_t nrow = PPT * nw; // <= 16
                uint32_t v = lane < nrow ? rowcnt[lane] : 0u;
                uint32_t inc = warp_incl_scan(v, lane);
                if (lane < nrow) rowbase[lane] = inc - v;
                if (lane == 31) {
                    rowbase[nrow] = inc; // block total
                    desc_publish(p.desc + dbase + b, ((unsigned long long)p.epoch << 32) | inc);
                }
            }
            // ---- 4. one-round look-back: every thread fetches a few predecessor descriptors
            uint32_t part = 0, part2 = 0;
            for (uint32_t i = tid; i < b; i += NT) {
                const unsigned long long *d = p.desc + dbase + i;
                unsigned long long v = desc_peek(d);
                while ((uint32_t)(v >> 32) != p.epoch && spin_budget) {
                    --spin_budget;
                    v = desc_peek(d);
                }
                part += (uint32_t)v;
            }
            if (s > 0) { // grand total of the previous pass of this frame
                for (uint32_t i = tid; i < G; i += NT) {
                    const unsigned long long *d = p.desc + dbase - G + i;
                    unsigned long long v = desc_peek(d);
                    while ((uint32_t)(v >> 32) != p.epoch && spin_budget) {
                        --spin_budget;
                        v = desc_peek(d);
                    }
                    part2 += (uint32_t)v;
                }
            }
            part = warp_sum(part);
            part2 = warp_sum(part2);
            if (lane == 0) { red[warp] = part; red2[warp] = part2; }
            __syncthreads();
            uint32_t pred = 0, prevtot = 0;
            for (uint32_t i = 0; i < nw; i++) { pred += red[i]; prevtot += red2[i]; }
            carry += prevtot;
            const uint32_t base = carry + pred;        // rank of this block's first entry in frame t
            const uint32_t total = rowbase[PPT * nw];  // entries of this block in this pass

            // ---- 5. stage (index, value) in rank order, update the reference, flush coalesced
            int *xs_out = p.xs + (size_t)t * p.cap;
            uint8_t *df_out = p.diff + (size_t)t * p.cap;
            const bool single = total <= p.stage_cap;
#pragma unroll
            for (int j = 0; j < PPT; j++) {
                const uint32_t round_lo = rowbase[j * nw];
                const uint32_t round_hi = rowbase[(j + 1) * nw];
                if (cnt[j]) {
                    uint32_t o = rowbase[j * nw + warp] + lex[j] - (single ? 0u : round_lo);
#pragma unroll
                    for (int k = 0; k < 8; k++) {
                        uint32_t m = piece_flags<HI>(c[j][k], r[j][k], p.addc, nv[j], k);
                        uint32_t dv = __vsub4(c[j][k], r[j][k]);
                        uint32_t fm = spread80(m);
                        while (m) {
                            int sh = __ffs((int)m) - 8; // 0, 8, 16 or 24
                            sxs[o] = (int)(off[j] + 4 * k + (sh >> 3));
                            sd[o] = (uint8_t)(dv >> sh);
                            o++;
                            m &= m - 1;
                        }
                        // negative feedback: reference := changed ? current : reference
                        r[j][k] = (c[j][k] & fm) | (r[j][k] & ~fm);
                    }
                    if (REFREG) dirty |= 1u << j;
                    else st256_keep(p.ref + off[j], r[j]);
                }
                if (!single) { // dense block: one flush per round of rows
                    __syncthreads();
                    const uint32_t n = round_hi - round_lo;
                    const size_t g0 = (size_t)base + round_lo;
                    for (uint32_t i = tid; i < n; i += NT) {
                        size_t g = g0 + i;
                        if (g < p.cap) {
                            stg_stream_u32(xs_out + g, (uint32_t)sxs[i]);
                            stg_stream_u8(df_out + g, sd[i]);
                        }
                    }
                    __syncthreads();
                }
            }
            if (single && total) {
                __syncthreads();
                for (uint32_t i = tid; i < total; i += NT) {
                    size_t g = (size_t)base + i;
                    if (g < p.cap) {
                        stg_stream_u32(xs_out + g, (uint32_t)sxs[i]);
                        stg_stream_u8(df_out + g, sd[i]);
                    }
                }
            }
            if (tid == 0) {
                if ((size_t)base + total > p.cap) atomicOr(p.status, 1u);
                if (b == G - 1 && s == p.nseg - 1) p.pos[t] = base + total;
            }
            if (!REFREG) geometry(s + 1 == p.nseg ? 0u : s + 1);
            __syncthreads(); // staging area and tables are reused by the next pass
        }
    }

    // REFREG: the reference lived in registers; write back only the pieces that changed
    if (REFREG) {
#pragma unroll
        for (int j = 0; j < PPT; j++)
            if (dirty & (1u << j)) st256_keep(p.ref + off[j], r[j]);
    }
    if (spin_budget == 0) atomicOr(p.status, 2u);
}

} // namespace cvs
