// cvs_stream_kernel.cuh -- the fused hot path: thresholded difference + negative feedback +
// ordered compaction (+ one display filter), as ONE persistent launch over a sequence of frames.
//
// Replaces kernel2 (server/src/kernels.cu:289-334), its CPU twin (tests/cuda_streaming/
// test.cu:560-576) and the visualiser kernels that read the same frame pair (kernels.cu:31-95,
// 243-281).  nframes = 1 is the drop-in exec_core path; nframes = T walks a device-resident
// sequence (frame t+1 is differenced against the reference frame t left behind).
//
// The path is integer byte work: on B200 it is bound by instruction issue long before HBM, so the
// design minimises instructions per frame byte and keeps HBM traffic at its floor.
//
// Work decomposition
//   * a frame is cut into groups of 48 B = 16 BGR pixels = three 16-byte vectors (cvs_pixel.cuh); a
//     thread owns a CHUNK of two consecutive groups (96 B), which halves the per-byte cost of the
//     scans, the barrier, the look-back and the flush;
//   * the grid is G persistent blocks of 512 threads, one per SM (128 registers per thread fill the register
//     file; measured 5 % faster than two blocks of 256), all co-resident (cooperative launch).
//     A frame is covered in nseg passes ("segments") of G*cps chunks; in segment s block b owns the cps
//     consecutive chunks starting at (s*G + b)*cps and thread i of the block owns chunk i of that
//     slice -- the SAME bytes in every frame.  One (frame, segment) pair is a "step";
//   * ingest: a block streams its slice of the coming steps into a 3-stage shared-memory ring with
//     1-D bulk copies (TMA engine, cp.async.bulk + mbarrier, L2 evict-first): every byte of a frame
//     crosses HBM -> SM exactly once, as 24 KB contiguous requests, and each thread picks its 32
//     whole pixels out of shared memory with six LDS.128;
//   * reference: when a frame fits one segment (1080p on 148 SMs) the thread's 96 reference bytes
//     live in registers for the whole sequence (REFREG): HBM never sees the reference between the
//     first frame and the last.  Otherwise each thread reloads / rewrites its own bytes with L2
//     evict-last accesses (the reference frame stays L2 resident; same thread, same address, so no
//     cross-thread hazard exists);
//   * one pass over the 24 words of a chunk produces, per word: byte-SIMD |cur-ref| > T flags, the
//     4-bit change nibble (one multiply gathers the four flag bits) merged into a 96-bit change mask,
//     the difference bytes cur-ref (parked in the thread's own 96 bytes of the ring stage) and the
//     updated reference (negative feedback).  popc of the mask is the thread's entry count;
//   * compaction: warp shuffle scan + one block scan (the only block-wide barrier of a step);
//     cross-block offsets by a one-round decoupled look-back: each block publishes
//     (epoch<<32 | count) for the step and sums the descriptors of its predecessors, each read by
//     its own thread; the running total of earlier segments of the frame travels in one extra
//     descriptor.  Each WARP then walks the set bits of its lanes' masks, stages (index, value) in
//     its own shared-memory window in rank order and flushes it with 16-byte (xs) / 4-byte (diff)
//     coalesced streaming stores -- warps drift apart freely, the last one to finish refills the stage;
//   * display filter MODE (heat map, red maps, grayscale, binarisation pass 1) is computed from the
//     same registers and written with 16-byte streaming stores;
//   * the step loop is software-pipelined: while a block runs the front half of step q (ingest ...
//     publish) the descriptors it needs for step q-1 are already in flight, and the back half of
//     step q-1 (staging and flush) follows, so the L2 round trip of the look-back stays hidden.
//
// Order, values and the new reference are bit-exact with oracle/cvs_oracle.c orc_diff_compact;
// unlike kernel2 the payload order is deterministic (ascending byte index).
#pragma once
#include "cvs_pixel.cuh"

namespace cvs {

#ifndef CVS_STREAM_THREADS
#define CVS_STREAM_THREADS 512
#endif
constexpr int kThreads = CVS_STREAM_THREADS;           // threads per block
constexpr int kWarps = kThreads / 32;
constexpr int kWarpsPad = (kWarps + 3) / 4 * 4;       // stride of the per-warp total arrays
constexpr int kBlocksPerSM = 512 / kThreads;          // 128 registers per thread fill the register file
constexpr int kLook = (148 * kBlocksPerSM + kThreads - 1) / kThreads + 1; // look-back descriptors a thread may read
constexpr int kGroupsPerThread = 2;
constexpr int kChunkBytes = kGroupsPerThread * kGroupBytes;   // 96
constexpr int kChunkWords = kChunkBytes / 4;                  // 24
constexpr int kMaskWords = kChunkBytes / 32;                  // 3
constexpr int kStageBytes = kThreads * kChunkBytes;   // 24,576 B: one block slice
constexpr int kStages = 3;                            // ring stages: one in process, two slices in flight
constexpr int kWarpEntries = 512;                     // payload entries a warp's staging window holds
constexpr uint32_t kWatchdogPolls = 1u << 24;         // look-back polls (>= 100 ns each) before giving up

enum StatusBits : unsigned { kStatusCapacity = 1u, kStatusWatchdog = 2u };

struct StreamParams {
    const uint8_t *frames;      // frame t at frames + t*frame_stride (16-byte aligned)
    size_t frame_stride;        // multiple of 16, >= nbytes rounded up to 16
    int nframes;
    uint8_t *ref;               // reference frame, padded to a whole number of chunks
    uint32_t nbytes;            // N = 3*W*H
    uint32_t nbytes16;          // N rounded up to 16
    uint32_t nchunks;           // ceil(N / 96)
    uint32_t nseg;              // segments per frame
    uint32_t cps;               // chunks per block per segment (<= kThreads)
    uint32_t nstages;           // ring stages in use (2..kStages); the launch pays SmemLayout::total(nstages)
    unsigned int *pos;          // [nframes]
    int *xs;                    // frame t at xs + t*cap
    uint8_t *diff;              // frame t at diff + t*cap
    size_t cap;                 // payload capacity per frame (entries)
    uint8_t *show;              // display frame t at show + t*show_stride (MODE 1,2,3,4,6)
    size_t show_stride;
    uint8_t *gray1;             // MODE 5/7: one gray byte per pixel, frame t at gray1 + t*gray_stride
    size_t gray_stride;
    unsigned int *hist;         // MODE 5/7: [nframes][256], zeroed by the host before the launch
    const uint32_t *heat_lut;   // MODE 1: 766 entries B | G<<8 | R<<16
    unsigned long long *desc;   // [nframes*nseg][G+1]
    uint32_t epoch;             // tag of this launch
    uint32_t addc;              // threshold constant for changed80<>
    unsigned int *status;       // StatusBits
    uint32_t debug;             // profiling experiments only (CVS_DEBUG_FLAGS): 1 no look-back, 2 no emission, 4 no per-word pass
};

// dynamic shared memory layout (bytes).  The small tables come first and the ring last, so the launch decides
// how many stages it pays for (total(nstages)): shared memory not used stays L1.
struct SmemLayout {
    static constexpr int kXsHalves = kWarpEntries + 8;                     // + alignment shift; 16-bit offsets in the warp's span
    static constexpr int kSdBytes = kWarpEntries + 16;
    static constexpr int lut = 0;                                          // 768 words
    static constexpr int hist = lut + 768 * 4;                             // 256 words
    static constexpr int wtot = hist + 256 * 4;                            // 2 x kWarpsPad words (by step parity)
    static constexpr int red = wtot + 2 * kWarpsPad * 4;                   // 2 x kWarpsPad words
    static constexpr int done = red + 2 * kWarpsPad * 4;                   // kStages words
    static constexpr int bar = done + 8 * 4;                               // kStages mbarriers
    static constexpr int sxs = bar + 8 * 8;                                // kWarps * kXsHalves uint16
    static constexpr int sd = sxs + kWarps * kXsHalves * 2;                // kWarps * kSdBytes bytes
    static constexpr int park = (sd + kWarps * kSdBytes + 127) / 128 * 128; // kStageBytes: reference bytes as they were before the step
    static constexpr int stage = park + kStageBytes;                       // nstages * kStageBytes
    static constexpr int total(int nstages) { return stage + nstages * kStageBytes; }
};
static_assert(kStages <= 8, "done[] / mbarrier slots");
static_assert(SmemLayout::bar % 8 == 0, "mbarrier alignment");
static_assert(SmemLayout::sd % 16 == 0 && SmemLayout::sxs % 16 == 0, "staging alignment");
static_assert((SmemLayout::kXsHalves * 2) % 16 == 0 && SmemLayout::kSdBytes % 16 == 0, "per-warp staging alignment");

// Coalesced flush by one warp of its n staged entries (window element i <-> global rank g0 + i).  The window was
// filled before g0 was known, so a 16-byte vector of xs / a 4-byte word of diff in global memory starts at window
// element `head` = (-g0) mod 4: the four 16-bit offsets are read one by one, the four difference bytes as two aligned
// words and a funnel shift.  A staged index is the 16-bit offset of the byte inside the warp's 3,072-byte span; wbase
// (the frame offset of the span) is added on the way out.
__device__ __forceinline__ void flush_warp(const uint16_t *sxs, const uint8_t *sd, uint32_t wbase, int *xs_out,
                                           uint8_t *df_out, size_t g0, uint32_t n, size_t cap, uint32_t lane)
{
    if (g0 >= cap) return;
    if (g0 + n > cap) n = (uint32_t)(cap - g0);
    int *xg = xs_out + g0;
    uint8_t *dg = df_out + g0;
    const uint32_t head = min(n, (0u - (uint32_t)g0) & 3u); // entries before the first 16-byte boundary of xs
    const uint32_t nq = (n - head) >> 2;                     // whole quads
    const uint32_t *sdw = reinterpret_cast<const uint32_t *>(sd);
    const uint32_t rot = 8u * (head & 3u);
#pragma unroll 1
    for (uint32_t k = lane; k < nq; k += 32) {
        const uint32_t a = head + 4 * k;
        stg_stream(xg + a, make_uint4(wbase + sxs[a], wbase + sxs[a + 1], wbase + sxs[a + 2], wbase + sxs[a + 3]));
        stg_stream_u32(dg + a, __funnelshift_r(sdw[a >> 2], sdw[(a >> 2) + 1], rot));
    }
    // the (at most three + three) entries before the first and after the last whole quad: one lane each
    const uint32_t e1 = lane < 4 ? lane : head + 4 * nq + (lane - 4);
    if (lane < 4 ? lane < head : (lane < 8 && e1 < n)) {
        stg_stream_u32(xg + e1, wbase + sxs[e1]);
        stg_stream_u8(dg + e1, sd[e1]);
    }
}

__device__ __forceinline__ uint32_t lds_u8(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

// walks the set bits of `bits` (bit j <-> byte `jbase + j` of the chunk): the index goes to sxs, the difference
// byte cur - ref (test.cu:566) is formed from the thread's bytes of the ring stage (curaddr) and of the parked
// reference (oldaddr) -- only for the bytes that are entries
__device__ __forceinline__ void emit_bits(uint32_t bits, uint32_t jbase, uint32_t loff, uint32_t curaddr, uint32_t oldaddr,
                                          uint16_t *sxs, uint8_t *sd, uint32_t &o)
{
    while (bits) {
        const uint32_t j = jbase + (uint32_t)__ffs((int)bits) - 1u;
        bits &= bits - 1u;
        sxs[o] = (uint16_t)(loff + j);
        sd[o] = (uint8_t)(lds_u8(curaddr + j) - lds_u8(oldaddr + j));
        o++;
    }
}

// Dense warps (more entries than the staging window holds): the warp walks its 32 chunks one after the other
// and handles each chunk with all lanes -- lane L owns bytes L, L+32 and L+64 of the chunk, finds its rank by a
// popc over the broadcast change mask and stores straight to global memory.  Consecutive changed bytes land on
// consecutive ranks, so every store instruction writes one contiguous run (up to 128 B of indices).
__device__ __forceinline__ void emit_coop(const uint32_t (&m)[kMaskWords], uint32_t coff0, uint32_t dvaddr0,
                                          int *xs_out, uint8_t *df_out, uint32_t g_lane, size_t cap, uint32_t lane)
{
    // 32-bit ranks and one-instruction bit tests: this loop is the bulk of a dense frame's instructions
    const uint32_t cap32 = cap > 0xffffffffull ? 0xffffffffu : (uint32_t)cap;
    const uint32_t lanebit = 1u << lane;
    constexpr int kBatch = 4; // chunks in flight: their shuffles and shared loads are issued before any store
    const uint32_t lt = (1u << lane) - 1u;
    for (uint32_t S0 = 0; S0 < 32; S0 += kBatch) {
        uint32_t sm[kBatch][kMaskWords], r[kBatch], v[kBatch][kMaskWords];
#pragma unroll
        for (int i = 0; i < kBatch; i++) {
#pragma unroll
            for (int w = 0; w < kMaskWords; w++) sm[i][w] = __shfl_sync(0xffffffffu, m[w], S0 + i);
            r[i] = __shfl_sync(0xffffffffu, g_lane, S0 + i); // global rank of the chunk's first entry
        }
#pragma unroll
        for (int i = 0; i < kBatch; i++)
#pragma unroll
            for (int w = 0; w < kMaskWords; w++) v[i][w] = lds_u8(dvaddr0 + (S0 + i) * kChunkBytes + lane + 32 * w);
#pragma unroll
        for (int i = 0; i < kBatch; i++) {
            const uint32_t cb = coff0 + (S0 + i) * kChunkBytes + lane;
            uint32_t rr = r[i];
#pragma unroll
            for (int w = 0; w < kMaskWords; w++) {
                const uint32_t g = rr + (uint32_t)__popc(sm[i][w] & lt);
                if ((sm[i][w] & lanebit) && g < cap32) {
                    stg_stream_u32(xs_out + g, cb + 32 * w);
                    stg_stream_u8(df_out + g, v[i][w]);
                }
                rr += (uint32_t)__popc(sm[i][w]);
            }
        }
    }
}

__device__ __forceinline__ uint32_t warp_add(uint32_t v) { return __reduce_add_sync(0xffffffffu, v); }

// sum of the first n and of all kWarps words at p: lane i reads word i, two warp reductions (REDUX)
__device__ __forceinline__ void sum_warps(const uint32_t *p, uint32_t n, uint32_t lane, uint32_t &first_n, uint32_t &all)
{
    static_assert(kWarps <= 32, "one lane per warp total");
    const uint32_t v = lane < (uint32_t)kWarps ? p[lane] : 0u;
    all = warp_add(v);
    first_n = warp_add(lane < n ? v : 0u);
}
__device__ __forceinline__ uint32_t sum_warps(const uint32_t *p, uint32_t lane)
{
    return warp_add(lane < (uint32_t)kWarps ? p[lane] : 0u);
}

template <int MODE, bool HI, bool REFREG>
__global__ void __launch_bounds__(kThreads, kBlocksPerSM) k_stream(const StreamParams p)
{
    extern __shared__ __align__(128) unsigned char smem[];
    uint32_t *slut = reinterpret_cast<uint32_t *>(smem + SmemLayout::lut);
    uint32_t *shist = reinterpret_cast<uint32_t *>(smem + SmemLayout::hist);
    uint32_t *wtot = reinterpret_cast<uint32_t *>(smem + SmemLayout::wtot);
    uint32_t *red = reinterpret_cast<uint32_t *>(smem + SmemLayout::red);
    uint32_t *done = reinterpret_cast<uint32_t *>(smem + SmemLayout::done);
    const uint32_t stage_addr = smem_u32(smem + SmemLayout::stage);
    const uint32_t park_addr = smem_u32(smem + SmemLayout::park);
    const uint32_t bar_addr = smem_u32(smem + SmemLayout::bar);

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t b = blockIdx.x, G = gridDim.x;
    const uint32_t N = p.nbytes;
    const uint32_t nsteps = (uint32_t)p.nframes * p.nseg;
    // ring depth in use (chosen by the host, <= kStages): all three stages when the reference lives in registers
    // (measured +5 % at 1080p); two when it goes through L2 (a deeper prefetch measured 17 % slower at 3840x2160).
    // A sparse warp releases its part of a stage in the iteration that consumed it, so two / one further slices are
    // in flight while a step is processed.
    const uint32_t nstages = p.nstages;
    constexpr bool kBinarize = (MODE == kModeBinarize || MODE == kModeBinarizeAvg);
    constexpr bool kGrayW = (MODE == kModeGrayWeighted || MODE == kModeBinarize);
    // this warp's staging window
    uint16_t *sxs = reinterpret_cast<uint16_t *>(smem + SmemLayout::sxs) + warp * SmemLayout::kXsHalves;
    uint8_t *sd = smem + SmemLayout::sd + warp * SmemLayout::kSdBytes;

    uint32_t phase = 0;     // bit st: parity the next wait on stage st expects
    bool tripped = false;   // watchdog expired once: stop waiting altogether

    // slice of this block in segment s: byte offset and byte count of the bulk copy
    auto slice = [&](uint32_t s, uint32_t &off, uint32_t &bytes) {
        uint64_t c0 = ((uint64_t)s * G + b) * p.cps;
        uint64_t o = c0 * kChunkBytes;
        if (o >= p.nbytes16) { off = 0; bytes = 0; return; }
        uint64_t e = o + (uint64_t)p.cps * kChunkBytes;
        if (e > p.nbytes16) e = p.nbytes16;
        off = (uint32_t)o;
        bytes = (uint32_t)(e - o);
    };
    // one thread: refill the ring stage of step q
    auto issue = [&](uint32_t q) {
        uint32_t t = q / p.nseg, s = q - t * p.nseg, off, bytes;
        slice(s, off, bytes);
        if (bytes) {
            const uint32_t st = q % nstages;
            const uint64_t pol = l2_policy_evict_first();
            // the stage may have been written through the generic proxy (difference bytes of dense warps)
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(bar_addr + 8 * st, bytes);
            bulk_g2s(stage_addr + st * kStageBytes, p.frames + (size_t)t * p.frame_stride + off, bytes,
                     bar_addr + 8 * st, pol);
        }
    };

    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < kStages; i++) {
            mbar_init(bar_addr + 8 * i, 1);
            done[i] = 0;
        }
        for (int i = 0; i < 2 * kWarpsPad; i++) wtot[i] = red[i] = 0;
        mbar_init_fence();
    }
    if (MODE == kModeHeat)
        for (uint32_t i = tid; i < 766; i += kThreads) slut[i] = p.heat_lut[i];
    __syncthreads();
    if (tid == 0)
        for (uint32_t q = 0; q < nstages && q < nsteps; q++) issue(q);

    uint32_t r[kChunkWords];
    const uint64_t keep = l2_policy_evict_last();
    bool dirty = false;
    uint32_t coff = 0, nv = 0; // byte offset of this thread's chunk in the frame, valid bytes (0..96)
    uint32_t sbytes = 0;       // bytes of the block's slice in the current segment
    auto geometry = [&](uint32_t s) {
        uint32_t soff;
        slice(s, soff, sbytes);
        uint64_t c = ((uint64_t)s * G + b) * p.cps + tid;
        bool ok = tid < p.cps && c < p.nchunks;
        coff = ok ? (uint32_t)(c * kChunkBytes) : 0u;
        nv = ok ? min(N - coff, (uint32_t)kChunkBytes) : 0u;
    };
    auto load_ref = [&]() {
        if (nv) {
#pragma unroll
            for (int v = 0; v < kChunkWords / 4; v++) {
                uint4 a = ldg_keep(p.ref + coff + 16 * v, keep);
                r[4 * v] = a.x; r[4 * v + 1] = a.y; r[4 * v + 2] = a.z; r[4 * v + 3] = a.w;
            }
        } else {
#pragma unroll
            for (int k = 0; k < kChunkWords; k++) r[k] = 0;
        }
    };
    auto store_ref = [&]() {
#pragma unroll
        for (int v = 0; v < kChunkWords / 4; v++)
            stg_keep(p.ref + coff + 16 * v, make_uint4(r[4 * v], r[4 * v + 1], r[4 * v + 2], r[4 * v + 3]), keep);
    };

    geometry(0);
    if (REFREG) load_ref(); // nseg == 1: the geometry never changes and the reference stays in registers

    // The loop is software-pipelined by one step: iteration q runs the FRONT half of step q (ingest, flags,
    // change mask, feedback, counts, publish) and then the BACK half of step q-1 (look-back sum, staging,
    // flush).  The predecessors' descriptors of step q-1 are fetched at the top of the iteration, so their L2
    // round trip hides behind the front half, and the single barrier of an iteration serves both the block
    // scan of step q and the look-back reduction of step q-1.
    uint32_t b_m[kMaskWords] = {0, 0, 0};
    uint32_t b_wrank = 0, b_wexc = 0, b_wtotal = 0, b_total = 0, b_wbase = 0, b_myaddr = 0, b_t = 0, b_s = 0;
    bool pending = false, b_dense = false;
    // this warp is done with ring stage st; the last warp of the block to say so refills it with the slice of step qn
    auto release = [&](uint32_t st, uint32_t qn) {
        __syncwarp();
        if (lane == 0) {
            __threadfence_block();
            if (atomicAdd(&done[st], 1u) == (uint32_t)kWarps - 1u) {
                done[st] = 0;
                if (qn < nsteps) issue(qn);
            }
        }
    };
    uint32_t t = 0, s = 0; // frame and segment of step q

    for (uint32_t q = 0; q <= nsteps; q++) {
        const bool front = q < nsteps;

        // ---- back half, part 1: start fetching the look-back descriptors of step q-1.  Thread i reads
        //      predecessors i, i+256, ...; thread b%256 also reads the running total of the earlier segments
        unsigned long long pv[kLook], pv2 = 0;
        const unsigned long long *prow = p.desc + (size_t)(q ? q - 1 : 0) * (G + 1);
        const bool look = pending && !(p.debug & 1u);
        const bool has2 = look && b_s > 0 && tid == (b % kThreads);
#pragma unroll
        for (int i = 0; i < kLook; i++) {
            pv[i] = 0;
            if (look && tid + i * kThreads < b) pv[i] = desc_peek(prow + tid + i * kThreads);
        }
        if (has2) pv2 = desc_peek(prow - 1); // slot G of the previous step

        uint32_t m[kMaskWords] = {0, 0, 0};
        uint32_t cnt = 0, incl = 0, myaddr = 0;
        if (front) {
            const uint32_t st = q % nstages;
            if (!REFREG) {
                geometry(s);
                load_ref(); // L2 hit; issued before the wait on the frame slice
            }
            if (kBinarize && s == 0) {
                // the thread that zeroes bin i is the one that flushed it at the end of the previous frame;
                // the barrier below orders the zeroing before this frame's atomics
                for (uint32_t i = tid; i < 256; i += kThreads) shist[i] = 0;
                __syncthreads();
            }

            // ---- 1. this thread's 32 pixels out of the ring
            if (sbytes) {
                // steps with an empty slice never touch the barrier, so the parity is tracked per stage
                if (!tripped && !mbar_wait(bar_addr + 8 * st, (phase >> st) & 1u)) {
                    tripped = true;
                    atomicOr(p.status, kStatusWatchdog);
                }
                phase ^= 1u << st;
            }
            myaddr = stage_addr + st * kStageBytes + tid * kChunkBytes;
            const uint32_t myold = park_addr + tid * kChunkBytes;
            uint32_t c[kChunkWords];
            if (nv) {
#pragma unroll
                for (int v = 0; v < kChunkWords / 4; v++) {
                    uint4 x = lds128(myaddr + 16 * v);
                    c[4 * v] = x.x; c[4 * v + 1] = x.y; c[4 * v + 2] = x.z; c[4 * v + 3] = x.w;
                }
                if (nv < (uint32_t)kChunkBytes) { // the chunk that holds the end of the frame: bytes past N never differ
#pragma unroll
                    for (int k = 0; k < kChunkWords; k++) {
                        const int vb = (int)nv - 4 * k;
                        const uint32_t vm = vb >= 4 ? 0xffffffffu : (vb <= 0 ? 0u : ((1u << (8 * vb)) - 1u));
                        c[k] = (c[k] & vm) | (r[k] & ~vm);
                    }
                }
            } else {
#pragma unroll
                for (int k = 0; k < kChunkWords; k++) c[k] = r[k];
            }

            // ---- 2. display filter on the same registers (reference as it was BEFORE this frame)
            if (MODE != kModeNone && nv) {
#pragma unroll
                for (int g = 0; g < kGroupsPerThread; g++) {
                    const uint32_t goff = coff + g * kGroupBytes;
                    const uint32_t gnv = nv > (uint32_t)(g * kGroupBytes) ? min(nv - g * kGroupBytes, (uint32_t)kGroupBytes) : 0u;
                    if (gnv == 0) continue;
                    uint32_t cg[kGroupWords], rg[kGroupWords], o[kGroupWords];
#pragma unroll
                    for (int k = 0; k < kGroupWords; k++) { cg[k] = c[g * kGroupWords + k]; rg[k] = r[g * kGroupWords + k]; }
                    if (MODE == kModeHeat) {
                        uint32_t ad[kGroupWords];
#pragma unroll
                        for (int k = 0; k < kGroupWords; k++) ad[k] = absdiff4(cg[k], rg[k]);
                        group_heat(ad, slut, o);
                        store_group(p.show + (size_t)t * p.show_stride + goff, o, gnv);
                    } else if (MODE == kModeRedBlack || MODE == kModeRedOverlap) {
                        uint32_t mk[kGroupWords];
#pragma unroll
                        for (int k = 0; k < kGroupWords; k++) mk[k] = changed80<HI>(absdiff4(cg[k], rg[k]), p.addc);
                        group_red<MODE == kModeRedOverlap>(mk, rg, o);
                        store_group(p.show + (size_t)t * p.show_stride + goff, o, gnv);
                    } else if (MODE == kModeGrayWeighted || MODE == kModeGrayAverage) {
                        group_gray3<kGrayW>(cg, o);
                        store_group(p.show + (size_t)t * p.show_stride + goff, o, gnv);
                    } else if (kBinarize) {
                        uint32_t g4[4];
                        group_gray1<kGrayW>(cg, g4);
                        const uint32_t npx = gnv / 3u; // whole pixels of this group inside the frame
                        uint8_t *gdst = p.gray1 + (size_t)t * p.gray_stride + goff / 3u;
                        if (npx == (uint32_t)kGroupPixels) stg_keep(gdst, make_uint4(g4[0], g4[1], g4[2], g4[3]), keep);
#pragma unroll
                        for (int px = 0; px < kGroupPixels; px++) {
                            if ((uint32_t)px < npx) {
                                uint32_t gv = (g4[px >> 2] >> (8 * (px & 3))) & 0xffu;
                                if (npx != (uint32_t)kGroupPixels) gdst[px] = (uint8_t)gv;
                                atomicAdd(&shist[gv], 1u); // server.cpp:103-106
                            }
                        }
                    }
                }
            }

            // ---- 3. one pass: flags -> 96-bit change mask, negative feedback
            //         reference := changed ? current : reference                      (test.cu:565-570)
            //         The reference bytes as they are BEFORE the feedback are parked in shared memory: the difference
            //         byte of an entry (current - reference, test.cu:566) is formed later, for the entries only.
            if (!(p.debug & 4u)) { // debug 4: skip the per-word pass (ingest-only experiment)
                if (nv) {
#pragma unroll
                    for (int v = 0; v < kChunkWords / 4; v++)
                        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(myold + 16 * v), "r"(r[4 * v]),
                                     "r"(r[4 * v + 1]), "r"(r[4 * v + 2]), "r"(r[4 * v + 3])
                                     : "memory");
                }
#pragma unroll
                for (int k = 0; k < kChunkWords; k += 2) {
                    const uint32_t f0 = changed80<HI>(absdiff4(c[k], r[k]), p.addc);
                    const uint32_t f1 = changed80<HI>(absdiff4(c[k + 1], r[k + 1]), p.addc);
                    // the eight flag bits of two words (7,15,23,31 and, shifted, 3,11,19,27) -> one byte of the change
                    // mask: every partial product of the multiply lands on its own bit, and bits 32..39 of the product
                    // are the flags in byte order
                    const uint32_t g8 = __umulhi(f1 + (f0 >> 4), 0x20408100u);
                    m[k >> 3] = __byte_perm(m[k >> 3], g8, ((k >> 1) & 3) == 0 ? 0x3214 : ((k >> 1) & 3) == 1 ? 0x3240
                                                          : ((k >> 1) & 3) == 2 ? 0x3410 : 0x4210);
                    const uint32_t fm0 = spread80(f0), fm1 = spread80(f1);
                    r[k] = (c[k] & fm0) | (r[k] & ~fm0);
                    r[k + 1] = (c[k + 1] & fm1) | (r[k + 1] & ~fm1);
                }
                if (nv < (uint32_t)kChunkBytes) { // bytes past the end of the frame are never entries (matters for T < 0)
#pragma unroll
                    for (int w = 0; w < kMaskWords; w++) {
                        const int vb = (int)nv - 32 * w;
                        m[w] &= vb >= 32 ? 0xffffffffu : (vb <= 0 ? 0u : ((1u << vb) - 1u));
                    }
                }
                if (m[0] | m[1] | m[2]) {
                    if (REFREG) dirty = true;
                    else store_ref();
                }
            }
            cnt = (uint32_t)__popc(m[0]) + (uint32_t)__popc(m[1]) + (uint32_t)__popc(m[2]);
            // only the warp total has to cross the barrier; the per-lane ranks are scanned after it
            const uint32_t wsum = warp_add(cnt);
            if (lane == 0) wtot[(q & 1u) * kWarpsPad + warp] = wsum;
        }

        // ---- back half, part 2: the descriptors fetched at the top (retry in the rare case a predecessor
        //      had not published yet)
        if (pending) {
            auto settle = [&](unsigned long long v, const unsigned long long *d) -> uint32_t {
                uint32_t polls = 0;
                while ((uint32_t)(v >> 32) != p.epoch && !tripped) {
                    __nanosleep(64);
                    v = desc_peek(d);
                    if (++polls > kWatchdogPolls) { // each poll costs well over 100 ns: seconds, i.e. a bug
                        tripped = true;
                        atomicOr(p.status, kStatusWatchdog);
                    }
                }
                return (uint32_t)v;
            };
            uint32_t part = 0;
#pragma unroll
            for (int i = 0; i < kLook; i++)
                if (look && tid + i * kThreads < b) part += settle(pv[i], prow + tid + i * kThreads);
            if (has2) part += settle(pv2, prow - 1);
            // G <= kLook * kThreads is enforced by the host, so kLook reads per thread cover every predecessor
            part = warp_add(part);
            if (lane == 0) red[(q & 1u) * kWarpsPad + warp] = part;
        }

        __syncthreads(); // the one barrier of a step: warp totals of step q, look-back partial sums of step q-1

        uint32_t total = 0, wexc = 0, wtotal = 0;
        if (front) {
            sum_warps(wtot + (q & 1u) * kWarpsPad, warp, lane, wexc, total); // entries of the warps before this one / of the block
            if (tid == 0) desc_publish(p.desc + (size_t)q * (G + 1) + b, ((unsigned long long)p.epoch << 32) | total);
            incl = warp_incl_scan(cnt, lane);
            wtotal = __shfl_sync(0xffffffffu, incl, 31); // entries of this warp in step q
        }

        if (pending) {
            const uint32_t base = sum_warps(red + (q & 1u) * kWarpsPad, lane);
            if (tid == 0) {
                if (b == G - 1) {
                    desc_publish(p.desc + (size_t)(q - 1) * (G + 1) + G, ((unsigned long long)p.epoch << 32) | (base + b_total));
                    if (b_s == p.nseg - 1) p.pos[b_t] = base + b_total;
                }
                if ((size_t)base + b_total > p.cap) atomicOr(p.status, kStatusCapacity);
            }

            // ---- back half, part 3: the (index, value) entries of step q-1 go out.  A sparse warp staged them in its
            //      window at the end of the previous iteration and flushes the window coalesced now that the global
            //      rank is known; a dense warp stores them straight from the ring stage (difference bytes in place).
            int *xs_out = p.xs + (size_t)b_t * p.cap;
            uint8_t *df_out = p.diff + (size_t)b_t * p.cap;
            const size_t g0 = (size_t)base + b_wexc; // global rank of this warp's first entry
            if (b_wtotal && !(p.debug & 2u)) {
                if (!b_dense) {
                    flush_warp(sxs, sd, b_wbase, xs_out, df_out, g0, b_wtotal, p.cap, lane);
                } else {
                    // chunk S of the warp starts 96*S bytes after lane 0's chunk (frame and ring stage alike)
                    emit_coop(b_m, b_wbase, b_myaddr - lane * kChunkBytes, xs_out, df_out, (uint32_t)g0 + b_wrank, p.cap, lane);
                }
            }
            if (b_dense) release((q - 1) % nstages, q - 1 + nstages);
        }

        if (front) {
            // ---- staging of step q (needs only ranks inside the warp).  Sparse warp: every lane walks the set bits of
            //      its mask and writes (16-bit offset, current - reference) in rank order into the warp's window; the
            //      warp is then done with the ring stage.  Dense warp (more entries than the window holds): every lane
            //      turns its 96 bytes of the stage into difference bytes in place; the stage stays held until the
            //      entries have gone out in the next iteration.
            const bool dense = wtotal > (uint32_t)kWarpEntries && !(p.debug & 2u);
            const uint32_t myold = park_addr + tid * kChunkBytes;
            if (wtotal && !(p.debug & 2u)) {
                if (!dense) {
                    __syncwarp(); // the flush above has read the window
                    uint32_t o = incl - cnt;
#pragma unroll
                    for (int w = 0; w < kMaskWords; w++) emit_bits(m[w], 32 * w, lane * kChunkBytes, myaddr, myold, sxs, sd, o);
                } else {
#pragma unroll
                    for (int v = 0; v < kChunkWords / 4; v++) {
                        const uint4 x = lds128(myaddr + 16 * v), y = lds128(myold + 16 * v);
                        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(myaddr + 16 * v), "r"(sub4(x.x, y.x)),
                                     "r"(sub4(x.y, y.y)), "r"(sub4(x.z, y.z)), "r"(sub4(x.w, y.w))
                                     : "memory");
                    }
                }
            }
            if (!dense) release(q % nstages, q + nstages);

            if (kBinarize && s == p.nseg - 1) {
                __syncthreads();
                for (uint32_t i = tid; i < 256; i += kThreads)
                    if (shist[i]) atomicAdd(p.hist + (size_t)t * 256 + i, shist[i]);
            }
#pragma unroll
            for (int w = 0; w < kMaskWords; w++) b_m[w] = m[w];
            b_wrank = incl - cnt; b_wexc = wexc; b_wtotal = wtotal; b_total = total;
            b_wbase = __shfl_sync(0xffffffffu, coff, 0); // lane 0 holds a chunk of the frame whenever any lane of the warp does
            b_myaddr = myaddr; b_t = t; b_s = s; b_dense = dense;
            pending = true;
            if (++s == p.nseg) { s = 0; ++t; }
        } else {
            pending = false;
        }
    }

    if (REFREG && dirty) store_ref();
}

} // namespace cvs
