// ingest.cu -- micro-benchmark: how fast can persistent blocks pull a sequence of 1080p frames through a TMA ring?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ingest ingest.cu ; run on the GPU box
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../cudavideostream_b200/csrc/cvs_device.cuh"
using namespace cvs;

constexpr int NT = 512, CHUNK = 96, SLICE = NT * CHUNK;

// MODE 0: block-level ring (thread 0 issues, __syncthreads per step)   MODE 1: per-warp ring (lane 0 issues 3 KB, no block barrier)
// MODE 2: plain LDG.128 loads, no shared memory
template <int MODE, int STAGES>
__global__ void __launch_bounds__(NT, 1) k_ingest(const uint8_t *frames, size_t stride, int nframes, uint32_t nbytes, uint32_t cps, uint32_t *out)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, b = blockIdx.x;
    const uint32_t soff = b * cps * CHUNK;
    const uint32_t sbytes = soff >= nbytes ? 0 : min(nbytes - soff, cps * CHUNK);
    uint32_t acc = 0;
    if (MODE == 2) {
        const bool mine = tid < cps && soff + tid * CHUNK + CHUNK <= nbytes;
        for (int t = 0; t < nframes; t++) {
            if (mine) {
                const uint4 *p = reinterpret_cast<const uint4 *>(frames + (size_t)t * stride + soff + tid * CHUNK);
#pragma unroll
                for (int v = 0; v < 6; v++) { uint4 x = ldg_stream(p + v); acc ^= x.x ^ x.y ^ x.z ^ x.w; }
            }
        }
        if (acc == 0x12345678) out[0] = acc;
        return;
    }
    const uint32_t stage_addr = smem_u32(smem + 1024), bar_addr = smem_u32(smem);
    const uint64_t pol = l2_policy_evict_first();
    if (MODE == 0) {
        if (tid == 0) { for (int i = 0; i < STAGES; i++) mbar_init(bar_addr + 8 * i, 1); mbar_init_fence(); }
        __syncthreads();
        auto issue = [&](int t) {
            if (sbytes) { const int st = t % STAGES; mbar_expect_tx(bar_addr + 8 * st, sbytes);
                bulk_g2s(stage_addr + st * SLICE, frames + (size_t)t * stride + soff, sbytes, bar_addr + 8 * st, pol); } };
        if (tid == 0) for (int t = 0; t < STAGES && t < nframes; t++) issue(t);
        uint32_t phase = 0;
        for (int t = 0; t < nframes; t++) {
            const int st = t % STAGES;
            if (sbytes) { mbar_wait(bar_addr + 8 * st, (phase >> st) & 1); phase ^= 1u << st; }
            const uint32_t a = stage_addr + st * SLICE + tid * CHUNK;
#pragma unroll
            for (int v = 0; v < 6; v++) { uint4 x = lds128(a + 16 * v); acc ^= x.x ^ x.y ^ x.z ^ x.w; }
            __syncthreads();
            if (tid == 0 && t + STAGES < nframes) issue(t + STAGES);
        }
    } else {
        // per-warp ring: warp w owns 32 chunks = 3072 B of the slice
        const uint32_t woff = warp * 32 * CHUNK;
        const uint32_t wbytes = woff >= sbytes ? 0 : min(sbytes - woff, 32u * CHUNK);
        const uint32_t wbar = bar_addr + 8 * (warp * STAGES);
        const uint32_t wstage = stage_addr + warp * STAGES * 32 * CHUNK;
        if (lane == 0) { for (int i = 0; i < STAGES; i++) mbar_init(wbar + 8 * i, 1); mbar_init_fence(); }
        __syncwarp();
        auto issue = [&](int t) {
            if (wbytes) { const int st = t % STAGES; mbar_expect_tx(wbar + 8 * st, wbytes);
                bulk_g2s(wstage + st * 32 * CHUNK, frames + (size_t)t * stride + soff + woff, wbytes, wbar + 8 * st, pol); } };
        if (lane == 0) for (int t = 0; t < STAGES && t < nframes; t++) issue(t);
        uint32_t phase = 0;
        for (int t = 0; t < nframes; t++) {
            const int st = t % STAGES;
            if (wbytes) { mbar_wait(wbar + 8 * st, (phase >> st) & 1); phase ^= 1u << st; }
            const uint32_t a = wstage + st * 32 * CHUNK + lane * CHUNK;
#pragma unroll
            for (int v = 0; v < 6; v++) { uint4 x = lds128(a + 16 * v); acc ^= x.x ^ x.y ^ x.z ^ x.w; }
            __syncwarp();
            if (lane == 0 && t + STAGES < nframes) issue(t + STAGES);
        }
    }
    if (acc == 0x12345678) out[0] = acc;
}

template <int MODE, int STAGES>
void run(const char *name, const uint8_t *d, size_t stride, int T, uint32_t N, uint32_t *out)
{
    const int G = 148;
    const uint32_t nchunks = (N + CHUNK - 1) / CHUNK, cps = (nchunks + G - 1) / G;
    size_t smem = 1024 + (size_t)STAGES * SLICE;
    if (MODE == 2) smem = 0;
    cudaFuncSetAttribute(k_ingest<MODE, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        k_ingest<MODE, STAGES><<<G, NT, smem>>>(d, stride, T, N, cps, out);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    printf("%-34s %7.3f ms  %6.2f us/frame  %6.0f GB/s  (%s)\n", name, best, best * 1e3 / T, (double)N * T / best / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    const uint32_t N = 3 * 1920 * 1080; const int T = 300;
    uint8_t *d; uint32_t *out; cudaMalloc(&d, (size_t)N * T + 4096); cudaMalloc(&out, 4); cudaMemset(d, 1, (size_t)N * T);
    run<0, 2>("block ring, 2 stages", d, N, T, N, out);
    run<0, 3>("block ring, 3 stages", d, N, T, N, out);
    run<0, 4>("block ring, 4 stages", d, N, T, N, out);
    run<1, 2>("warp rings, 2 stages", d, N, T, N, out);
    run<1, 4>("warp rings, 4 stages", d, N, T, N, out);
    run<2, 1>("plain LDG.128 (no smem)", d, N, T, N, out);
    return 0;
}
