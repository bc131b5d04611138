// synccheck probe: the fill's point-to-point mbarrier pattern in isolation (init by the first threads, __syncthreads,
// per-warp arrive by lane 0, parity wait by the next warp).  nvcc -arch=sm_100a -o mbar_sync mbar_sync.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void mbar_init(unsigned bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive(unsigned bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
{
    asm volatile("{\n.reg .pred p;\nWAIT_LOOP:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra WAIT_DONE;\nbra WAIT_LOOP;\nWAIT_DONE:\n}\n" ::"r"(bar), "r"(parity) : "memory");
}
template <int MODE>
__global__ void k(int steps, int* out)
{
    __shared__ unsigned long long step_bar[64];
    __shared__ int ring[8][160];
    const int tid = threadIdx.x, wrp = tid >> 5, nwarps = blockDim.x >> 5;
    const unsigned bar0 = (unsigned)__cvta_generic_to_shared(&step_bar[0]);
    if (tid < 8 * nwarps) mbar_init(bar0 + 8 * tid, 1);
    if (MODE == 1) asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    if (MODE == 2) asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    const unsigned bar_mine = bar0 + wrp * 64, bar_left = bar0 + (wrp == 0 ? nwarps - 1 : wrp - 1) * 64;
    __syncthreads();
    int acc = 0;
    unsigned slot = 0, par = 0;
    for (int d = 0; d < steps; d++)
    {
        if (d > 0) mbar_wait(bar_left + ((slot + 56u) & 63u), slot == 0 ? par ^ 1u : par);
        const int left = tid == 0 ? blockDim.x - 1 : tid - 1;
        if (d > 0) acc += ring[(d - 1) & 7][left];
        ring[d & 7][tid] = acc + tid;
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(bar_mine + slot);
        slot = (slot + 8u) & 63u; par ^= (slot == 0);
    }
    out[blockIdx.x * blockDim.x + tid] = acc;
}
int main()
{
    int* out; cudaMalloc(&out, 4 * 160 * sizeof(int));
    k<0><<<4, 160>>>(100, out); printf("mode 0: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    k<1><<<4, 160>>>(100, out); printf("mode 1: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    k<2><<<4, 160>>>(100, out); printf("mode 2: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
