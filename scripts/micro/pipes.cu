// Issue rate of the byte-SIMD building blocks of the stream kernel's per-word pass on sm_100a:
// warp-instructions per cycle per SM sub-partition for VABSDIFF4, LOP3, PRMT, IMAD.HI, IADD, POPC.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int OP>
__global__ void k(uint32_t *out, uint32_t seed, int iters, long long *cycles)
{
    uint32_t a[8];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = seed * (threadIdx.x + 1) + i * 0x01030507u;
    const uint32_t b = seed ^ 0x55aa33ccu;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int i = 0; i < 8; i++) { // eight independent chains per thread
                if (OP == 0) a[i] = __vabsdiffu4(a[i], b);
                if (OP == 1) asm volatile("lop3.b32 %0, %0, %1, 0x0f0f0f0f, 0x96;" : "+r"(a[i]) : "r"(b));
                if (OP == 2) asm volatile("prmt.b32 %0, %0, %1, 0x5140;" : "+r"(a[i]) : "r"(b));
                if (OP == 3) a[i] = __umulhi(a[i], b);
                if (OP == 4) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
                if (OP == 5) a[i] = __popc(a[i]) + b;
                if (OP == 6) asm volatile("shf.r.wrap.b32 %0, %0, %1, 7;" : "+r"(a[i]) : "r"(b));
            }
        }
    }
    const long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int OP>
void run(const char *name, int threads)
{
    uint32_t *out;
    long long *cyc, h;
    cudaMalloc(&out, 148 * 1024 * 4);
    cudaMalloc(&cyc, 8);
    const int iters = 2000;
    k<OP><<<148, threads>>>(out, 12345u, iters, cyc);
    k<OP><<<148, threads>>>(out, 12345u, iters, cyc);
    cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double insts_per_smsp = (double)iters * 32 * (threads / 32) / 4.0; // warp-instructions per sub-partition
    printf("%-10s %4d threads/SM: %.3f warp-inst/cycle/SMSP (%.2f cycles per warp-inst)\n", name, threads, insts_per_smsp / (double)h,
           (double)h / insts_per_smsp);
    cudaFree(out);
    cudaFree(cyc);
}

int main()
{
    for (int threads : {128, 512, 1024}) {
        run<0>("VABSDIFF4", threads);
        run<1>("LOP3", threads);
        run<2>("PRMT", threads);
        run<3>("IMAD.HI", threads);
        run<4>("IADD", threads);
        run<5>("POPC+IADD", threads);
        run<6>("SHF", threads);
    }
    return 0;
}
