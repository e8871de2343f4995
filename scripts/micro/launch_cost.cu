// launch_cost.cu -- what a launch of the stream kernel's shape costs before any work is done: empty kernels timed with
// events (one launch between two events, best / median of 200), plain and cooperative, for several block / shared-memory shapes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o launch_cost launch_cost.cu && ./launch_cost
#include <algorithm>
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>

__global__ void k_empty(int *p) { if (p && threadIdx.x == 9999) *p = 1; }
__global__ void __launch_bounds__(1024, 1) k_touch(const uint4 *in, uint4 *out, int n16)
{
    // every thread reads and writes 96 bytes (the stream kernel's reference traffic for one frame): 6 x 16 B
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    uint4 v[6];
#pragma unroll
    for (int i = 0; i < 6; i++) v[i] = t * 6 + i < n16 ? in[t * 6 + i] : make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int i = 0; i < 6; i++) if (t * 6 + i < n16) out[t * 6 + i] = v[i];
}

static void time_it(const char *name, void (*launch)(cudaStream_t), cudaStream_t st)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    std::vector<float> ms;
    for (int i = 0; i < 220; i++) {
        cudaEventRecord(e0, st);
        launch(st);
        cudaEventRecord(e1, st);
        cudaStreamSynchronize(st);
        float m; cudaEventElapsedTime(&m, e0, e1);
        if (i >= 20) ms.push_back(m);
    }
    std::sort(ms.begin(), ms.end());
    printf("%-58s best %6.2f us  median %6.2f us\n", name, 1e3 * ms[0], 1e3 * ms[ms.size() / 2]);
}

static int g_smem; static dim3 g_grid, g_block; static bool g_coop;
static uint4 *g_in, *g_out; static int g_n16;
static void launch_empty(cudaStream_t st)
{
    int *p = nullptr; void *args[] = {&p};
    if (g_coop) cudaLaunchCooperativeKernel((const void *)k_empty, g_grid, g_block, args, g_smem, st);
    else cudaLaunchKernel((const void *)k_empty, g_grid, g_block, args, g_smem, st);
}
static void launch_touch(cudaStream_t st) { k_touch<<<g_grid, g_block, g_smem, st>>>(g_in, g_out, g_n16); }
static void launch_two(cudaStream_t st)
{
    g_smem = 0; g_block = dim3(256); g_grid = dim3(592); launch_empty(st);
    g_smem = 226 * 1024; g_block = dim3(1024); g_grid = dim3(148); launch_empty(st);
}

int main()
{
    cudaStream_t st; cudaStreamCreate(&st);
    cudaFuncSetAttribute(k_empty, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(k_touch, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    const int N = 1920 * 1080 * 3; g_n16 = N / 16;
    cudaMalloc(&g_in, N + 64); cudaMalloc(&g_out, N + 64); cudaMemset(g_in, 1, N);
    struct { const char *n; int grid, block, smem; bool coop; } cfg[] = {
        {"empty 148 x 1024, 226 KB smem, plain", 148, 1024, 226 * 1024, false},
        {"empty 148 x 1024, 226 KB smem, cooperative", 148, 1024, 226 * 1024, true},
        {"empty 148 x 1024, 0 KB smem, plain", 148, 1024, 0, false},
        {"empty 148 x 512, 226 KB smem, plain", 148, 512, 226 * 1024, false},
        {"empty 592 x 256, 48 KB smem, plain", 592, 256, 48 * 1024, false},
        {"empty 592 x 256, 0 KB smem, plain", 592, 256, 0, false},
        {"empty 1 x 32, 0 KB smem, plain", 1, 32, 0, false},
    };
    for (auto &c : cfg) { g_grid = dim3(c.grid); g_block = dim3(c.block); g_smem = c.smem; g_coop = c.coop; time_it(c.n, launch_empty, st); }
    g_coop = false;
    time_it("small-smem kernel, then 148 x 1024 with 226 KB (carve-out switch)", launch_two, st);
    g_grid = dim3(148); g_block = dim3(1024); g_smem = 226 * 1024;
    time_it("read + write 6.2 MB, 148 x 1024 (96 B per thread), 226 KB smem", launch_touch, st);
    g_grid = dim3(1520); g_block = dim3(256); g_smem = 0;
    k_touch<<<1, 32>>>(g_in, g_out, 0);
    time_it("read + write 6.2 MB, 1520 x 256 (96 B per thread), 0 KB smem", launch_touch, st);
    return 0;
}
