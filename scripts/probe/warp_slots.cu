// Which hardware warp slots (%warpid) do the warps of small long-lived blocks get?  (scheduler = slot % 4)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void probe(int* out, long long spin)
{
    extern __shared__ double sm[];
    unsigned smid, warpid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(warpid));
    long long t0 = clock64();
    while (clock64() - t0 < spin) { sm[threadIdx.x] += 1.0; }
    if ((threadIdx.x & 31) == 0) {
        int w = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
        out[3 * w] = smid; out[3 * w + 1] = warpid; out[3 * w + 2] = blockIdx.x;
    }
}
int main(int argc, char** argv)
{
    int threads = argc > 1 ? atoi(argv[1]) : 64, blocks = argc > 2 ? atoi(argv[2]) : 256, smem = argc > 3 ? atoi(argv[3]) : 57344;
    int nw = blocks * threads / 32, *d, *h = new int[3 * nw];
    cudaMalloc(&d, 3 * nw * sizeof(int));
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe<<<blocks, threads, smem>>>(d, 2000000);
    cudaMemcpy(h, d, 3 * nw * sizeof(int), cudaMemcpyDeviceToHost);
    printf("threads %d blocks %d smem %d: %s\n", threads, blocks, smem, cudaGetErrorString(cudaGetLastError()));
    int hist[4] = {0, 0, 0, 0};
    for (int w = 0; w < nw; w++) hist[h[3 * w + 1] % 4]++;
    printf("warps per scheduler index (slot %% 4): %d %d %d %d\n", hist[0], hist[1], hist[2], hist[3]);
    for (int s = 0; s < 3; s++) { printf("SM %d:", s); for (int w = 0; w < nw; w++) if (h[3 * w] == s) printf(" (block %d slot %d)", h[3 * w + 2], h[3 * w + 1]); printf("\n"); }
    return 0;
}
