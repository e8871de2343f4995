// cvs_device.cuh -- small device-side helpers shared by the sm_100a kernels of libcvs_b200.
//
// Byte-SIMD-in-a-register helpers (four BGR bytes per 32-bit lane), cache-hinted vector
// loads/stores, the mbarrier / bulk-copy (TMA) primitives of the frame ingest ring and the
// descriptor primitives of the cross-block payload-offset exchange.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cvs {

// ------------------------------------------------------------------------------------------
// 16-byte global accesses with cache hints.
// ------------------------------------------------------------------------------------------
// frames that are read exactly once
__device__ __forceinline__ uint4 ldg_stream(const void *p)
{
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p));
    return v;
}
// reference frame: re-read by the next frame, keep it in L2 (evict-last policy through a cache hint;
// the plain .L2::evict_last qualifier only exists for 256-bit accesses).  NOT .nc: the same thread
// rewrites the line.
__device__ __forceinline__ uint64_t l2_policy_evict_last()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint4 ldg_keep(const void *p, uint64_t pol)
{
    uint4 v;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p), "l"(pol)
                 : "memory");
    return v;
}
__device__ __forceinline__ void stg_keep(void *p, uint4 v, uint64_t pol)
{
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.u32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "r"(v.x),
                 "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol)
                 : "memory");
}
// outputs nobody on the device reads again (payload, display frames)
__device__ __forceinline__ void stg_stream(void *p, uint4 v)
{
    asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
}
__device__ __forceinline__ void stg_stream_u32(void *p, uint32_t v)
{
    asm volatile("st.global.cs.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void stg_stream_u8(void *p, uint32_t v)
{
    asm volatile("st.global.cs.u8 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ------------------------------------------------------------------------------------------
// mbarrier + 1-D bulk copy (TMA engine, SASS UBLKCP): global -> shared, completion on an mbarrier.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint64_t global_ns()
{
    uint64_t t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// watchdog for the two spin waits of the stream kernel: a wait that lasts longer than this is a bug
// (or a lost co-residency guarantee); the kernel then raises a status bit and stops waiting so that
// the launch always terminates.
constexpr uint64_t kWatchdogNs = 2000000000ull;

// returns false when the watchdog expired.
// The wait is try_wait WITH a suspend-time hint: the hardware parks the warp until the phase completes (or the hint
// runs out), so a waiting warp costs no issue slots.  (Round 2: ncu showed 35 % of all executed instructions of the
// warp-specialised kernel inside the former spin loop -- probe, globaltimer read, compare, branch -- competing with
// the working warps for the ALU pipe.)  The clock is only read when a parked wait came back empty-handed.
constexpr uint32_t kMbarSuspendNs = 100000u;
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done)
                 : "r"(bar), "r"(parity), "r"(kMbarSuspendNs)
                 : "memory");
    return done;
}
// sleep_ns: pause between two probes.  A parked try_wait comes back every time the barrier is touched (a bulk copy
// updates the pending byte count of its barrier packet by packet: ~36 wake-ups per 42 KB slice), and with 32 warps
// per SM the probing warps -- ncu: 46 % of all executed instructions -- take issue slots from the working ones; the
// sleep costs at most its own length in latency.  The clock is read every 256th probe only.
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity, uint32_t sleep_ns = 0)
{
    if (mbar_try_wait(bar, parity)) return true;
    uint32_t spins = 0;
    uint64_t t0 = 0;
    while (true) {
        if (sleep_ns) __nanosleep(sleep_ns);
        if (mbar_try_wait(bar, parity)) return true;
        if ((++spins & 255u) == 0u) {
            const uint64_t now = global_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > kWatchdogNs) return false;
        }
    }
}
__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
// bytes: multiple of 16, > 0; src and dst 16-byte aligned
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar, uint64_t pol)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar), "l"(pol)
                 : "memory");
}

__device__ __forceinline__ uint4 lds128(uint32_t addr)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}

// ------------------------------------------------------------------------------------------
// payload-offset descriptors: one 64-bit word per (frame, segment, block) = (epoch << 32) | count.
// The count travels in the same word as the tag, so relaxed accesses are sufficient.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void desc_publish(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long desc_peek(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// ------------------------------------------------------------------------------------------
// per-byte compare on 4 packed bytes.
//   changed <=> |cur - ref| > thr  (server/include/common.h:14 LR_THRESHOLDS; test.cu:565)
// ad   = per-byte absolute difference
// addc = per-byte constant, hi = false: (128 - (thr+1)) for thr <= 127
//                           hi = true : (256 - (thr+1)) for 128 <= thr <= 254, 0x00 never matches for thr = 255
// returns 0x80 in every byte lane that changed.
// ------------------------------------------------------------------------------------------
template <bool HI>
__device__ __forceinline__ uint32_t changed80(uint32_t ad, uint32_t addc)
{
    uint32_t low = (ad & 0x7f7f7f7fu) + addc; // no carry between byte lanes: 127 + 127 < 256
    return (HI ? (low & ad) : (low | ad)) & 0x80808080u;
}

// 0x80 flags -> 0xFF byte masks (PRMT with the sign-replicate bit set in every selector nibble)
__device__ __forceinline__ uint32_t spread80(uint32_t m80)
{
    uint32_t r;
    asm("prmt.b32 %0, %1, %1, 0xba98;" : "=r"(r) : "r"(m80));
    return r;
}

__device__ __forceinline__ uint32_t absdiff4(uint32_t a, uint32_t b) { return __vabsdiffu4(a, b); }

// per-byte a - b (mod 256): the payload value df & 0xFF of test.cu:566.  Bit 7 of every byte lane of the
// minuend is forced on and cleared in the subtrahend so no borrow crosses a lane; the true bit 7 is patched in.
// ONE_LOP: the final t ^ x ^ 0x80808080 forced into one three-input logic op (left to itself the compiler emits two:
// 6 instead of 5 instructions per word).  Costs registers where they are scarce, so the caller chooses.
template <bool ONE_LOP>
__device__ __forceinline__ uint32_t sub4(uint32_t a, uint32_t b)
{
    const uint32_t t = (a | 0x80808080u) - (b & 0x7f7f7f7fu);
    if (ONE_LOP) { // t ^ ((a ^ b) & H) ^ H with the last two xors as one three-input op
        const uint32_t x = (a ^ b) & 0x80808080u;
        uint32_t z;
        asm("lop3.b32 %0, %1, %2, 0x80808080, 0x96;" : "=r"(z) : "r"(t), "r"(x));
        return z;
    }
    return t ^ ((a ^ ~b) & 0x80808080u);
}

// horizontal sum of the four byte lanes (each lane <= 63 so the sum fits)
__device__ __forceinline__ uint32_t hsum4(uint32_t x) { return (x * 0x01010101u) >> 24; }

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t n = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += n;
    }
    return v;
}
__device__ __forceinline__ uint32_t warp_sum(uint32_t v)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

// splitmix64 finaliser (the synthetic camera's counter-based generator; twin in synth.py)
__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

} // namespace cvs
