// cvs_stream_sref.cuh -- the fused hot path for frames that fit one pass of the grid (1080p and below on a
// B200): same contract, decomposition and payload as k_stream (cvs_stream_kernel.cuh), but the REFERENCE FRAME
// LIVES IN SHARED MEMORY for the whole sequence (148 SMs x 48 KB >= 6.2 MB).
//
// Why: with both the current slice (ring stage) and the reference slice byte-addressable on chip, the per-word
// pass only has to FIND the changed bytes (|cur-ref| > T flags -> 96-bit change mask -> count: 7 instructions
// per 4 bytes).  The difference byte cur-ref and the negative feedback ref := cur are then done per ENTRY while
// the entry is staged (two shared byte loads, a subtract, one shared byte store) instead of per WORD for every
// byte of the frame -- k_stream spends 8 more instructions per 4 bytes on cur-ref and the merge, plus the
// parking of the difference bytes.  Nothing is parked, so the ring stage is released when the step ends; three
// stages keep two slices (2 x 43 KB per SM) in flight, which is what it takes to cover the HBM latency.
//
// Step q of a block (one block of 512 threads per SM, thread i owns chunk i of the block's slice, 96 B):
//   wait for the slice of frame q (TMA bulk copy) -> LDS cur, LDS ref -> [display filter] -> flags / mask / count
//   -> barrier 1 -> block prefix, publish (epoch, count), start loading the predecessors' descriptors
//   -> sparse warps stage their entries (warp-local ranks only) and apply the feedback byte by byte, which hides
//      the L2 round trip of the look-back -> sum the descriptors -> barrier 2 -> global offset
//   -> sparse warps flush their window (16-byte / 4-byte coalesced stores, any alignment), dense warps walk their
//      chunks with all lanes, store directly and apply the feedback word by word
//   -> the last warp to finish hands the stage back to the TMA ring.
#pragma once
#include "cvs_stream_kernel.cuh"

namespace cvs {

constexpr int kSrefThreads = kThreads;                      // same block geometry as k_stream
constexpr int kSrefWarps = kSrefThreads / 32;
constexpr int kSrefStages = 3;                              // two slices in flight while one is processed
constexpr int kSrefStageBytes = kSrefThreads * kChunkBytes; // 49,152
constexpr int kSrefLook = kLook;

struct SrefLayout {
    static constexpr int kXsHalves = kWarpEntries;            // staged index = 16-bit offset inside the warp's 3,072 bytes
    static constexpr int kSdBytes = kWarpEntries;
    static constexpr int stage = 0;
    static constexpr int sref = kSrefStages * kSrefStageBytes;
    static constexpr int sxs = sref + kSrefStageBytes;
    static constexpr int sd = sxs + kSrefWarps * kXsHalves * 2;
    static constexpr int lut = sd + kSrefWarps * kSdBytes;
    static constexpr int hist = lut + 768 * 4;
    static constexpr int wtot = hist + 256 * 4;
    static constexpr int red = wtot + kSrefWarps * 4;
    static constexpr int done = red + kSrefWarps * 4;
    static constexpr int bar = done + 4 * 4;
    static constexpr int total = bar + kSrefStages * 8;
};
static_assert(SrefLayout::bar % 8 == 0 && SrefLayout::wtot % 16 == 0 && SrefLayout::sxs % 16 == 0, "alignment");
static_assert((SrefLayout::total + 1024) * kBlocksPerSM <= 228 * 1024, "shared memory budget of an SM");

__device__ __forceinline__ void sts_u8(uint32_t addr, uint32_t v)
{
    asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// sum of the first n and of all kSrefWarps words at p
__device__ __forceinline__ void sum16(const uint32_t *p, uint32_t n, uint32_t &first_n, uint32_t &all)
{
    first_n = 0; all = 0;
#pragma unroll
    for (int q4 = 0; q4 < kSrefWarps / 4; q4++) {
        const uint4 a = *reinterpret_cast<const uint4 *>(p + 4 * q4);
        const uint32_t v[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int i = 0; i < 4; i++) {
            if ((uint32_t)(4 * q4 + i) < n) first_n += v[i];
            all += v[i];
        }
    }
}

// Flush n entries staged at window elements [0, n) to global ranks [g0, g0 + n); g0 has any alignment.  Aligned
// quads of GLOBAL ranks leave as one 16-byte + one 4-byte store; the ragged ends entry by entry.
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
    // ragged ends: elements [0, head) and [tail0, n), at most 3 + 3
    if (lane < 8) {
        const uint32_t e = lane < 4 ? lane : tail0 + (lane - 4);
        const bool ok = lane < 4 ? e < head : e < n;
        if (ok) {
            stg_stream_u32(xs_out + g0 + e, wbase + sxs[e]);
            stg_stream_u8(df_out + g0 + e, sd[e]);
        }
    }
}

template <int MODE, bool HI>
__global__ void __launch_bounds__(kSrefThreads, kBlocksPerSM) k_stream_sref(const StreamParams p)
{
    extern __shared__ __align__(128) unsigned char smem[];
    uint32_t *slut = reinterpret_cast<uint32_t *>(smem + SrefLayout::lut);
    uint32_t *shist = reinterpret_cast<uint32_t *>(smem + SrefLayout::hist);
    uint32_t *wtot = reinterpret_cast<uint32_t *>(smem + SrefLayout::wtot);
    uint32_t *red = reinterpret_cast<uint32_t *>(smem + SrefLayout::red);
    uint32_t *done = reinterpret_cast<uint32_t *>(smem + SrefLayout::done);
    const uint32_t stage_addr = smem_u32(smem + SrefLayout::stage);
    const uint32_t bar_addr = smem_u32(smem + SrefLayout::bar);

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t b = blockIdx.x, G = gridDim.x;
    const uint32_t N = p.nbytes;
    const uint32_t nsteps = (uint32_t)p.nframes; // one segment per frame
    constexpr bool kBinarize = (MODE == kModeBinarize || MODE == kModeBinarizeAvg);
    constexpr bool kGrayW = (MODE == kModeGrayWeighted || MODE == kModeBinarize);
    uint16_t *sxs = reinterpret_cast<uint16_t *>(smem + SrefLayout::sxs) + warp * SrefLayout::kXsHalves;
    uint8_t *sd = smem + SrefLayout::sd + warp * SrefLayout::kSdBytes;

    // geometry: the same bytes in every frame
    const uint64_t c0 = (uint64_t)b * p.cps;
    const uint32_t soff = (uint32_t)min((uint64_t)p.nbytes16, c0 * kChunkBytes);
    const uint32_t sbytes = (uint32_t)(min((uint64_t)p.nbytes16, (c0 + p.cps) * kChunkBytes) - soff);
    const bool mine = tid < p.cps && c0 + tid < p.nchunks;
    const uint32_t coff = mine ? (uint32_t)((c0 + tid) * kChunkBytes) : 0u;
    const uint32_t nv = mine ? min(N - coff, (uint32_t)kChunkBytes) : 0u;
    const uint32_t myref = smem_u32(smem + SrefLayout::sref) + tid * kChunkBytes; // this thread's reference bytes
    const uint64_t keep = l2_policy_evict_last();

    auto issue = [&](uint32_t q) { // one thread: bulk copy of the block's slice of frame q into stage q % kSrefStages
        if (sbytes) {
            const uint32_t st = q % kSrefStages;
            const uint64_t pol = l2_policy_evict_first();
            mbar_expect_tx(bar_addr + 8 * st, sbytes);
            bulk_g2s(stage_addr + st * kSrefStageBytes, p.frames + (size_t)q * p.frame_stride + soff, sbytes,
                     bar_addr + 8 * st, pol);
        }
    };

    if (tid == 0) {
        for (int i = 0; i < kSrefStages; i++) {
            mbar_init(bar_addr + 8 * i, 1);
            done[i] = 0;
        }
        mbar_init_fence();
    }
    if (MODE == kModeHeat)
        for (uint32_t i = tid; i < 766; i += kSrefThreads) slut[i] = p.heat_lut[i];
    // the reference slice moves on chip for the whole launch
#pragma unroll
    for (int v = 0; v < kChunkWords / 4; v++) {
        uint4 a = make_uint4(0, 0, 0, 0);
        if (nv) a = ldg_keep(p.ref + coff + 16 * v, keep);
        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(myref + 16 * v), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w) : "memory");
    }
    __syncthreads();
    if (tid == 0)
        for (uint32_t q = 0; q < (uint32_t)kSrefStages && q < nsteps; q++) issue(q);

    uint32_t phase = 0;
    bool tripped = false, dirty = false;

    for (uint32_t q = 0; q < nsteps; q++) {
        const uint32_t st = q % kSrefStages;
        const uint32_t mycur = stage_addr + st * kSrefStageBytes + tid * kChunkBytes;
        if (kBinarize) {
            for (uint32_t i = tid; i < 256; i += kSrefThreads) shist[i] = 0;
            __syncthreads();
        }
        if (sbytes) {
            if (!tripped && !mbar_wait(bar_addr + 8 * st, (phase >> st) & 1u)) {
                tripped = true;
                atomicOr(p.status, kStatusWatchdog);
            }
            phase ^= 1u << st;
        }

        // ---- current and reference bytes of this thread's chunk
        uint32_t c[kChunkWords], r[kChunkWords];
#pragma unroll
        for (int v = 0; v < kChunkWords / 4; v++) {
            const uint4 y = lds128(myref + 16 * v);
            r[4 * v] = y.x; r[4 * v + 1] = y.y; r[4 * v + 2] = y.z; r[4 * v + 3] = y.w;
        }
        if (nv) {
#pragma unroll
            for (int v = 0; v < kChunkWords / 4; v++) {
                const uint4 x = lds128(mycur + 16 * v);
                c[4 * v] = x.x; c[4 * v + 1] = x.y; c[4 * v + 2] = x.z; c[4 * v + 3] = x.w;
            }
        } else {
#pragma unroll
            for (int k = 0; k < kChunkWords; k++) c[k] = r[k];
        }

        // ---- display filter on the same registers (reference as it was BEFORE this frame)
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
                    store_group(p.show + (size_t)q * p.show_stride + goff, o, gnv);
                } else if (MODE == kModeRedBlack || MODE == kModeRedOverlap) {
                    uint32_t mk[kGroupWords];
#pragma unroll
                    for (int k = 0; k < kGroupWords; k++) mk[k] = changed80<HI>(absdiff4(cg[k], rg[k]), p.addc);
                    if (gnv < (uint32_t)kGroupBytes) { // bytes past the end of the frame are never changes
#pragma unroll
                        for (int k = 0; k < kGroupWords; k++) {
                            const int vb = (int)gnv - 4 * k;
                            mk[k] &= vb >= 4 ? 0xffffffffu : (vb <= 0 ? 0u : ((1u << (8 * vb)) - 1u));
                        }
                    }
                    group_red<MODE == kModeRedOverlap>(mk, rg, o);
                    store_group(p.show + (size_t)q * p.show_stride + goff, o, gnv);
                } else if (MODE == kModeGrayWeighted || MODE == kModeGrayAverage) {
                    group_gray3<kGrayW>(cg, o);
                    store_group(p.show + (size_t)q * p.show_stride + goff, o, gnv);
                } else if (kBinarize) {
                    uint32_t g4[4];
                    group_gray1<kGrayW>(cg, g4);
                    const uint32_t npx = gnv / 3u;
                    uint8_t *gdst = p.gray1 + (size_t)q * p.gray_stride + goff / 3u;
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

        // ---- flags -> 96-bit change mask -> count
        uint32_t m[kMaskWords] = {0, 0, 0};
#pragma unroll
        for (int k = 0; k < kChunkWords; k++) {
            const uint32_t f = changed80<HI>(absdiff4(c[k], r[k]), p.addc);
            const uint32_t nib = f * 0x00204081u; // flag bits 7,15,23,31 -> bits 28..31
            m[k >> 3] |= (nib >> (28 - 4 * (k & 7))) & (0xFu << (4 * (k & 7)));
        }
        if (nv < (uint32_t)kChunkBytes) { // bytes past the end of the frame are never entries
#pragma unroll
            for (int w = 0; w < kMaskWords; w++) {
                const int vb = (int)nv - 32 * w;
                m[w] &= vb >= 32 ? 0xffffffffu : (vb <= 0 ? 0u : ((1u << vb) - 1u));
            }
        }
        const uint32_t cnt = (uint32_t)__popc(m[0]) + (uint32_t)__popc(m[1]) + (uint32_t)__popc(m[2]);
        const uint32_t wtotal = warp_add(cnt);
        if (lane == 0) wtot[warp] = wtotal;
        __syncthreads(); // barrier 1: warp counts

        uint32_t wexc, total;
        sum16(wtot, warp, wexc, total);
        unsigned long long *drow = p.desc + (size_t)q * (G + 1);
        if (tid == 0) desc_publish(drow + b, ((unsigned long long)p.epoch << 32) | total);
        // the predecessors' descriptors of this step: in flight while the sparse warps stage their entries
        unsigned long long pv[kSrefLook];
        const bool look = !(p.debug & 1u);
#pragma unroll
        for (int i = 0; i < kSrefLook; i++) {
            pv[i] = 0;
            if (look && tid + i * kSrefThreads < b) pv[i] = desc_peek(drow + tid + i * kSrefThreads);
        }
        const uint32_t incl = warp_incl_scan(cnt, lane);
        const uint32_t wrank = incl - cnt; // rank of this lane's first entry inside the warp
        const bool sparse = wtotal <= (uint32_t)kWarpEntries;

        // ---- sparse warps: stage (index, cur - ref) in rank order and apply the negative feedback ref := cur,
        //      entry by entry (test.cu:565-570)
        if (sparse && cnt && !(p.debug & 2u)) {
            uint32_t o = wrank;
#pragma unroll
            for (int w = 0; w < kMaskWords; w++) {
                uint32_t bits = m[w];
                while (bits) {
                    const uint32_t j = 32u * w + (uint32_t)__ffs((int)bits) - 1u;
                    bits &= bits - 1u;
                    const uint32_t cv = lds_u8(mycur + j), rv = lds_u8(myref + j);
                    sxs[o] = (uint16_t)(lane * kChunkBytes + j); // offset inside the warp's span
                    sd[o] = (uint8_t)(cv - rv);
                    sts_u8(myref + j, cv);
                    o++;
                }
            }
            dirty = true;
        }

        // ---- look-back: sum of the predecessors' counts
        {
            uint32_t part = 0;
#pragma unroll
            for (int i = 0; i < kSrefLook; i++) {
                if (look && tid + i * kSrefThreads < b) {
                    unsigned long long v = pv[i];
                    uint32_t polls = 0;
                    while ((uint32_t)(v >> 32) != p.epoch && !tripped) {
                        __nanosleep(64);
                        v = desc_peek(drow + tid + i * kSrefThreads);
                        if (++polls > kWatchdogPolls) {
                            tripped = true;
                            atomicOr(p.status, kStatusWatchdog);
                        }
                    }
                    part += (uint32_t)v;
                }
            }
            part = warp_add(part);
            if (lane == 0) red[warp] = part;
        }
        __syncthreads(); // barrier 2: look-back partial sums
        uint32_t base, unused;
        sum16(red, 0, unused, base);
        if (tid == 0) {
            if (b == G - 1) p.pos[q] = base + total;
            if ((size_t)base + total > p.cap) atomicOr(p.status, kStatusCapacity);
        }

        // ---- payload of this warp
        int *xs_out = p.xs + (size_t)q * p.cap;
        uint8_t *df_out = p.diff + (size_t)q * p.cap;
        const size_t g0 = (size_t)base + wexc; // global rank of this warp's first entry
        if (wtotal && !(p.debug & 2u)) {
            if (sparse) {
                __syncwarp();
                flush_window(sxs, sd, __shfl_sync(0xffffffffu, coff, 0), xs_out, df_out, g0, wtotal, p.cap, lane);
            } else {
                // dense warp.  First the per-word work k_stream does for every chunk: difference bytes cur - ref (parked
                // over the thread's own 96 bytes of the stage, which nobody else needs any more) and the negative
                // feedback merge into the shared-memory reference.
#pragma unroll
                for (int k = 0; k < kChunkWords; k++) {
                    const uint32_t nib = (m[k >> 3] >> (4 * (k & 7))) & 0xFu;
                    const uint32_t fm = ((nib * 0x00204081u) & 0x01010101u) * 0xFFu; // mask bit i -> byte i
                    const uint32_t dvk = sub4(c[k], r[k]);
                    r[k] = (c[k] & fm) | (r[k] & ~fm);
                    c[k] = dvk;
                }
#pragma unroll
                for (int v = 0; v < kChunkWords / 4; v++) {
                    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(mycur + 16 * v), "r"(c[4 * v]), "r"(c[4 * v + 1]),
                                 "r"(c[4 * v + 2]), "r"(c[4 * v + 3]) : "memory");
                    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(myref + 16 * v), "r"(r[4 * v]), "r"(r[4 * v + 1]),
                                 "r"(r[4 * v + 2]), "r"(r[4 * v + 3]) : "memory");
                }
                if (cnt) dirty = true;
                __syncwarp();
                // then all lanes walk the warp's 32 chunks (lane L owns bytes L, L+32, L+64 of a chunk; its rank is a popc
                // over the broadcast mask) and store straight to global memory in contiguous runs
                const uint32_t lt = (1u << lane) - 1u;
                const uint32_t cur0 = mycur - lane * kChunkBytes;
                const uint32_t coff0 = __shfl_sync(0xffffffffu, coff, 0);
                constexpr int kBatch = 4;
                for (uint32_t S0 = 0; S0 < 32; S0 += kBatch) {
                    uint32_t sm[kBatch][kMaskWords], rk[kBatch], dv[kBatch][kMaskWords];
#pragma unroll
                    for (int i = 0; i < kBatch; i++) {
#pragma unroll
                        for (int w = 0; w < kMaskWords; w++) sm[i][w] = __shfl_sync(0xffffffffu, m[w], S0 + i);
                        rk[i] = __shfl_sync(0xffffffffu, wrank, S0 + i);
                    }
#pragma unroll
                    for (int i = 0; i < kBatch; i++)
#pragma unroll
                        for (int w = 0; w < kMaskWords; w++) dv[i][w] = lds_u8(cur0 + (S0 + i) * kChunkBytes + lane + 32 * w);
#pragma unroll
                    for (int i = 0; i < kBatch; i++) {
                        const uint32_t cb = coff0 + (S0 + i) * kChunkBytes + lane;
                        size_t rr = g0 + rk[i];
#pragma unroll
                        for (int w = 0; w < kMaskWords; w++) {
                            const size_t g = rr + (uint32_t)__popc(sm[i][w] & lt);
                            if (((sm[i][w] >> lane) & 1u) && g < p.cap) {
                                stg_stream_u32(xs_out + g, cb + 32 * w);
                                stg_stream_u8(df_out + g, dv[i][w]);
                            }
                            rr += (uint32_t)__popc(sm[i][w]);
                        }
                    }
                }
            }
        } else if (wtotal && (p.debug & 2u)) {
            // profiling experiment "no emission": the feedback still has to happen
            if (cnt) {
#pragma unroll
                for (int k = 0; k < kChunkWords; k++) {
                    const uint32_t nib = (m[k >> 3] >> (4 * (k & 7))) & 0xFu;
                    const uint32_t fm = ((nib * 0x00204081u) & 0x01010101u) * 0xFFu;
                    r[k] = (c[k] & fm) | (r[k] & ~fm);
                }
#pragma unroll
                for (int v = 0; v < kChunkWords / 4; v++)
                    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(myref + 16 * v), "r"(r[4 * v]),
                                 "r"(r[4 * v + 1]), "r"(r[4 * v + 2]), "r"(r[4 * v + 3])
                                 : "memory");
                dirty = true;
            }
        }

        if (kBinarize) {
            __syncthreads();
            for (uint32_t i = tid; i < 256; i += kSrefThreads)
                if (shist[i]) atomicAdd(p.hist + (size_t)q * 256 + i, shist[i]);
        }

        // ---- this warp no longer needs the stage: the last warp to get here refills it with frame q + 2
        __syncwarp();
        if (lane == 0) {
            __threadfence_block();
            if (atomicAdd(&done[st], 1u) == (uint32_t)kSrefWarps - 1u) {
                done[st] = 0;
                if (q + kSrefStages < nsteps) issue(q + kSrefStages);
            }
        }
    }

    // the reference slice goes back to global memory
    if (dirty && nv) {
#pragma unroll
        for (int v = 0; v < kChunkWords / 4; v++) {
            const uint4 y = lds128(myref + 16 * v);
            stg_keep(p.ref + coff + 16 * v, y, keep);
        }
    }
}

} // namespace cvs
