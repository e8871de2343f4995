// cvs_stream_pair.cuh -- k_stream_pair: the fused hot path for frames that fit one pass of the grid (1080p
// and below on a B200), TWO FRAMES PER STEP.
//
// Same contract, decomposition, payload and per-word arithmetic as k_stream<.., REFREG = true>
// (cvs_stream_kernel.cuh): persistent blocks of 512 threads, thread i owns the same 96-byte chunk in every frame,
// reference bytes in registers for the whole sequence, TMA bulk copies into a shared-memory ring, change mask +
// parked difference bytes + negative feedback in one pass, one-round look-back, per-warp staging / flush.
//
// What changes: measured on B200, a step of k_stream costs ~1.5 us even with all per-byte work removed -- the
// latency chain of a step (mbarrier wait, warp reductions, block barrier, descriptor publish / look-back, stage
// hand-back) is per STEP, not per byte.  Because frame t+1 only needs the reference a thread itself has just updated,
// a thread can run the per-word pass over frames 2q and 2q+1 back to back and share everything else: the two
// counts travel packed in one word (16 bits each; a block holds at most 49,152 entries per frame) through one warp
// reduction, one block scan, one descriptor and one look-back, behind two block barriers per PAIR of frames.  The
// look-back is not software-pipelined here (a ring stage holds two slices, 96 KB, and only two stages fit); its L2
// round trip is hidden behind the staging of the sparse warps, which needs warp-local ranks only.
#pragma once
#include "cvs_stream_kernel.cuh"

namespace cvs {

constexpr int kPairStages = 2;
constexpr int kPairStageBytes = 2 * kStageBytes; // two slices

struct PairLayout {
    static constexpr int kXsHalves = kWarpEntries + 8;
    static constexpr int kSdBytes = kWarpEntries + 16;
    static constexpr int lut = 0;                                   // 768 words
    static constexpr int hist = lut + 768 * 4;                      // 256 words
    static constexpr int wtot = hist + 256 * 4;                     // kWarps words (packed pair of counts)
    static constexpr int red = wtot + kWarps * 4;                   // 2 x kWarps words
    static constexpr int done = red + 2 * kWarps * 4;               // kPairStages words (+ pad)
    static constexpr int bar = done + 8 * 4;                        // kPairStages mbarriers
    static constexpr int sxs = bar + 8 * 8;                         // kWarps * kXsHalves uint16
    static constexpr int sd = sxs + kWarps * kXsHalves * 2;         // kWarps * kSdBytes bytes
    static constexpr int stage = (sd + kWarps * kSdBytes + 127) / 128 * 128;
    static constexpr int total = stage + kPairStages * kPairStageBytes;
};
static_assert(kThreads == 512 && kBlocksPerSM == 1, "the pair kernel is written for one 512-thread block per SM");
static_assert(PairLayout::total <= 227 * 1024, "shared memory budget of one block per SM");
static_assert(PairLayout::bar % 8 == 0 && PairLayout::wtot % 16 == 0 && PairLayout::red % 16 == 0, "alignment");

// Flush n entries staged at window elements [0, n) to global ranks [g0, g0 + n); g0 has any alignment.  Aligned
// quads of GLOBAL ranks leave as one 16-byte + one 4-byte store; the ragged ends entry by entry.  A staged index is
// the 16-bit offset of the byte inside the warp's 3,072-byte span (wbase = frame offset of the span).
__device__ __forceinline__ void flush_window(const uint16_t *sxs, const uint8_t *sd, uint32_t wbase, int *xs_out,
                                             uint8_t *df_out, size_t g0, uint32_t n, size_t cap, uint32_t lane)
{
    if (g0 >= cap) return;
    if (g0 + n > cap) n = (uint32_t)(cap - g0);
    const uint32_t head = min(n, (uint32_t)((4 - (g0 & 3)) & 3)); // entries before the first aligned quad
    const uint32_t nq = (n - head) >> 2;
    const uint32_t tail0 = head + 4 * nq;
    for (uint32_t i = lane; i < nq; i += 32) {
        const uint32_t e = head + 4 * i;
        const uint4 x = make_uint4(wbase + sxs[e], wbase + sxs[e + 1], wbase + sxs[e + 2], wbase + sxs[e + 3]);
        const uint32_t v = (uint32_t)sd[e] | ((uint32_t)sd[e + 1] << 8) | ((uint32_t)sd[e + 2] << 16) | ((uint32_t)sd[e + 3] << 24);
        stg_stream(xs_out + g0 + e, x);
        stg_stream_u32(df_out + g0 + e, v);
    }
    if (lane < 8) { // ragged ends: elements [0, head) and [tail0, n), at most 3 + 3
        const uint32_t e = lane < 4 ? lane : tail0 + (lane - 4);
        const bool ok = lane < 4 ? e < head : e < n;
        if (ok) {
            stg_stream_u32(xs_out + g0 + e, wbase + sxs[e]);
            stg_stream_u8(df_out + g0 + e, sd[e]);
        }
    }
}

// sum of the first n and of all kWarps words at p (p holds packed pairs; sums stay inside their 16-bit fields)
__device__ __forceinline__ void sum_packed(const uint32_t *p, uint32_t n, uint32_t &first_n, uint32_t &all)
{
    first_n = 0; all = 0;
#pragma unroll
    for (int q4 = 0; q4 < kWarps / 4; q4++) {
        const uint4 a = *reinterpret_cast<const uint4 *>(p + 4 * q4);
        const uint32_t v[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int i = 0; i < 4; i++) {
            if ((uint32_t)(4 * q4 + i) < n) first_n += v[i];
            all += v[i];
        }
    }
}

template <int MODE, bool HI>
__global__ void __launch_bounds__(kThreads, 1) k_stream_pair(const StreamParams p)
{
    extern __shared__ __align__(128) unsigned char smem[];
    uint32_t *slut = reinterpret_cast<uint32_t *>(smem + PairLayout::lut);
    uint32_t *shist = reinterpret_cast<uint32_t *>(smem + PairLayout::hist);
    uint32_t *wtot = reinterpret_cast<uint32_t *>(smem + PairLayout::wtot);
    uint32_t *red = reinterpret_cast<uint32_t *>(smem + PairLayout::red);
    uint32_t *done = reinterpret_cast<uint32_t *>(smem + PairLayout::done);
    const uint32_t stage_addr = smem_u32(smem + PairLayout::stage);
    const uint32_t bar_addr = smem_u32(smem + PairLayout::bar);

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t b = blockIdx.x, G = gridDim.x;
    const uint32_t N = p.nbytes;
    const uint32_t nframes = (uint32_t)p.nframes;
    const uint32_t nsteps = (nframes + 1) / 2;
    constexpr bool kBinarize = (MODE == kModeBinarize || MODE == kModeBinarizeAvg);
    constexpr bool kGrayW = (MODE == kModeGrayWeighted || MODE == kModeBinarize);
    uint16_t *sxs = reinterpret_cast<uint16_t *>(smem + PairLayout::sxs) + warp * PairLayout::kXsHalves;
    uint8_t *sd = smem + PairLayout::sd + warp * PairLayout::kSdBytes;

    // geometry: the same bytes in every frame
    const uint64_t c0 = (uint64_t)b * p.cps;
    const uint32_t soff = (uint32_t)min((uint64_t)p.nbytes16, c0 * kChunkBytes);
    const uint32_t sbytes = (uint32_t)(min((uint64_t)p.nbytes16, (c0 + p.cps) * kChunkBytes) - soff);
    const bool mine = tid < p.cps && c0 + tid < p.nchunks;
    const uint32_t coff = mine ? (uint32_t)((c0 + tid) * kChunkBytes) : 0u;
    const uint32_t nv = mine ? min(N - coff, (uint32_t)kChunkBytes) : 0u;
    const uint64_t keep = l2_policy_evict_last();

    auto issue = [&](uint32_t q) { // one thread: bulk copies of the block's slices of frames 2q, 2q+1 into stage q & 1
        if (sbytes) {
            const uint32_t st = q & 1u, nf = min(2u, nframes - 2 * q);
            const uint64_t pol = l2_policy_evict_first();
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // the stage held parked bytes (generic proxy)
            mbar_expect_tx(bar_addr + 8 * st, sbytes * nf);
            for (uint32_t f = 0; f < nf; f++)
                bulk_g2s(stage_addr + st * kPairStageBytes + f * kStageBytes,
                         p.frames + (size_t)(2 * q + f) * p.frame_stride + soff, sbytes, bar_addr + 8 * st, pol);
        }
    };

    if (tid == 0) {
        for (int i = 0; i < kPairStages; i++) {
            mbar_init(bar_addr + 8 * i, 1);
            done[i] = 0;
        }
        mbar_init_fence();
    }
    if (MODE == kModeHeat)
        for (uint32_t i = tid; i < 766; i += kThreads) slut[i] = p.heat_lut[i];
    __syncthreads();
    if (tid == 0)
        for (uint32_t q = 0; q < (uint32_t)kPairStages && q < nsteps; q++) issue(q);

    // the reference bytes of this thread's chunk live in registers for the whole launch
    uint32_t r[kChunkWords];
#pragma unroll
    for (int v = 0; v < kChunkWords / 4; v++) {
        uint4 a = make_uint4(0, 0, 0, 0);
        if (nv) a = ldg_keep(p.ref + coff + 16 * v, keep);
        r[4 * v] = a.x; r[4 * v + 1] = a.y; r[4 * v + 2] = a.z; r[4 * v + 3] = a.w;
    }
    uint32_t phase = 0;
    bool tripped = false, dirty = false;

    for (uint32_t q = 0; q < nsteps; q++) {
        const uint32_t st = q & 1u, nf = min(2u, nframes - 2 * q);
        if (sbytes) {
            if (!tripped && !mbar_wait(bar_addr + 8 * st, (phase >> st) & 1u)) {
                tripped = true;
                atomicOr(p.status, kStatusWatchdog);
            }
            phase ^= 1u << st;
        }

        // ---- front half of both frames: flags -> change mask, parked difference bytes, negative feedback
        uint32_t m[2][kMaskWords] = {{0, 0, 0}, {0, 0, 0}};
        uint32_t park[2];
#pragma unroll
        for (int f = 0; f < 2; f++) {
            const uint32_t t = 2 * q + f;
            park[f] = stage_addr + st * kPairStageBytes + f * kStageBytes + tid * kChunkBytes;
            if ((uint32_t)f >= nf) continue; // block-uniform
            if (kBinarize) {
                for (uint32_t i = tid; i < 256; i += kThreads) shist[i] = 0;
                __syncthreads();
            }
            uint32_t c[kChunkWords];
            if (nv) {
#pragma unroll
                for (int v = 0; v < kChunkWords / 4; v++) {
                    const uint4 x = lds128(park[f] + 16 * v);
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

            // display filter on the same registers (reference as it was BEFORE this frame)
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
                        const uint32_t npx = gnv / 3u;
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
            if (kBinarize) {
                __syncthreads();
                for (uint32_t i = tid; i < 256; i += kThreads)
                    if (shist[i]) atomicAdd(p.hist + (size_t)t * 256 + i, shist[i]);
            }

            // one pass per word (test.cu:565-570): reference := changed ? current : reference
            uint32_t dv[kChunkWords];
#pragma unroll
            for (int k = 0; k < kChunkWords; k++) {
                const uint32_t fl = changed80<HI>(absdiff4(c[k], r[k]), p.addc);
                const uint32_t nib = fl * 0x00204081u; // flag bits 7,15,23,31 -> bits 28..31
                m[f][k >> 3] |= (nib >> (28 - 4 * (k & 7))) & (0xFu << (4 * (k & 7)));
                dv[k] = sub4(c[k], r[k]);
                const uint32_t fm = spread80(fl);
                r[k] = (c[k] & fm) | (r[k] & ~fm);
            }
            if (nv < (uint32_t)kChunkBytes) { // bytes past the end of the frame are never entries (matters for T < 0)
#pragma unroll
                for (int w = 0; w < kMaskWords; w++) {
                    const int vb = (int)nv - 32 * w;
                    m[f][w] &= vb >= 32 ? 0xffffffffu : (vb <= 0 ? 0u : ((1u << vb) - 1u));
                }
            }
            if (m[f][0] | m[f][1] | m[f][2]) {
#pragma unroll
                for (int v = 0; v < kChunkWords / 4; v++)
                    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(park[f] + 16 * v), "r"(dv[4 * v]),
                                 "r"(dv[4 * v + 1]), "r"(dv[4 * v + 2]), "r"(dv[4 * v + 3])
                                 : "memory");
                dirty = true;
            }
        }

        // ---- counts of both frames packed in one word
        const uint32_t cnt0 = (uint32_t)__popc(m[0][0]) + (uint32_t)__popc(m[0][1]) + (uint32_t)__popc(m[0][2]);
        const uint32_t cnt1 = (uint32_t)__popc(m[1][0]) + (uint32_t)__popc(m[1][1]) + (uint32_t)__popc(m[1][2]);
        const uint32_t cntp = cnt0 | (cnt1 << 16);
        const uint32_t wtotp = warp_add(cntp); // <= 3,072 per field
        if (lane == 0) wtot[warp] = wtotp;
        __syncthreads(); // barrier 1

        uint32_t wexcp, totalp;
        sum_packed(wtot, warp, wexcp, totalp); // <= 49,152 per field
        unsigned long long *drow = p.desc + (size_t)q * (G + 1);
        if (tid == 0) desc_publish(drow + b, ((unsigned long long)p.epoch << 32) | totalp);
        unsigned long long pv = 0; // G <= 512: one predecessor per thread
        const bool look = !(p.debug & 1u) && tid < b;
        if (look) pv = desc_peek(drow + tid);
        const uint32_t inclp = warp_incl_scan(cntp, lane);
        const uint32_t wrankp = inclp - cntp;

        // ---- sparse warps stage now (warp-local ranks only): this hides the look-back's L2 round trip.  A frame is
        //      sparse for this warp when its entries fit the window; when both frames fit together they share it
        //      (frame 1 behind frame 0), otherwise frame 1 is staged after frame 0 has been flushed.
        const uint32_t cnts[2] = {cnt0, cnt1};
        const uint32_t wts[2] = {wtotp & 0xffffu, wtotp >> 16};
        const bool sparse[2] = {wts[0] <= (uint32_t)kWarpEntries, wts[1] <= (uint32_t)kWarpEntries};
        const bool share = sparse[0] && sparse[1] && wts[0] + wts[1] <= (uint32_t)kWarpEntries;
        const uint32_t woff[2] = {0u, share ? wts[0] : 0u}; // window offset of each frame's entries
        auto stage_frame = [&](int f) {
            if (cnts[f]) {
                uint32_t o = woff[f] + ((wrankp >> (16 * f)) & 0xffffu);
#pragma unroll
                for (int w = 0; w < kMaskWords; w++) emit_bits(m[f][w], 32 * w, lane * kChunkBytes, park[f], sxs, sd, o);
            }
        };
        if (!(p.debug & 2u)) {
            if (sparse[0]) stage_frame(0);
            if (share) stage_frame(1);
        }

        // ---- look-back: sums of the predecessors' two counts
        {
            uint32_t p0 = 0, p1 = 0;
            if (look) {
                uint32_t polls = 0;
                while ((uint32_t)(pv >> 32) != p.epoch && !tripped) {
                    __nanosleep(64);
                    pv = desc_peek(drow + tid);
                    if (++polls > kWatchdogPolls) {
                        tripped = true;
                        atomicOr(p.status, kStatusWatchdog);
                    }
                }
                p0 = (uint32_t)pv & 0xffffu;
                p1 = ((uint32_t)pv >> 16) & 0xffffu;
            }
            p0 = warp_add(p0);
            p1 = warp_add(p1);
            if (lane == 0) { red[warp] = p0; red[kWarps + warp] = p1; }
        }
        __syncthreads(); // barrier 2
        uint32_t bases[2], unused;
        sum_packed(red, 0, unused, bases[0]);
        sum_packed(red + kWarps, 0, unused, bases[1]);

#pragma unroll
        for (int f = 0; f < 2; f++) {
            if ((uint32_t)f >= nf) continue;
            const uint32_t t = 2 * q + f;
            const uint32_t total = (totalp >> (16 * f)) & 0xffffu, wexc = (wexcp >> (16 * f)) & 0xffffu;
            const uint32_t wt = wts[f], wrank = (wrankp >> (16 * f)) & 0xffffu;
            if (tid == 0) {
                if (b == G - 1) p.pos[t] = bases[f] + total;
                if ((size_t)bases[f] + total > p.cap) atomicOr(p.status, kStatusCapacity);
            }
            int *xs_out = p.xs + (size_t)t * p.cap;
            uint8_t *df_out = p.diff + (size_t)t * p.cap;
            const size_t g0 = (size_t)bases[f] + wexc; // global rank of this warp's first entry of frame t
            if (wt && !(p.debug & 2u)) {
                if (sparse[f]) {
                    __syncwarp();
                    if (f == 1 && !share) { // the window was busy with frame 0 (or frame 0 was dense): stage now
                        stage_frame(1);
                        __syncwarp();
                    }
                    flush_window(sxs + woff[f], sd + woff[f], __shfl_sync(0xffffffffu, coff, 0), xs_out, df_out, g0, wt,
                                 p.cap, lane);
                } else {
                    __syncwarp();
                    emit_coop(m[f], __shfl_sync(0xffffffffu, coff, 0), park[f] - lane * kChunkBytes, xs_out, df_out,
                              (uint32_t)g0 + wrank, p.cap, lane);
                }
            }
        }

        // ---- this warp no longer needs the stage: the last warp to get here refills it with the pair q + 2
        __syncwarp();
        if (lane == 0) {
            __threadfence_block();
            if (atomicAdd(&done[st], 1u) == (uint32_t)kWarps - 1u) {
                done[st] = 0;
                if (q + kPairStages < nsteps) issue(q + kPairStages);
            }
        }
    }

    if (dirty && nv) {
#pragma unroll
        for (int v = 0; v < kChunkWords / 4; v++)
            stg_keep(p.ref + coff + 16 * v, make_uint4(r[4 * v], r[4 * v + 1], r[4 * v + 2], r[4 * v + 3]), keep);
    }
}

} // namespace cvs
