// probe_device.cu — prints the device facts the cluster-resident kernel is sized with (run on the B200 box).
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dummy_kernel(int* p) { extern __shared__ char sm[]; if (p) p[0] = sm[0]; }

int main() {
    cudaDeviceProp pr;
    cudaGetDeviceProperties(&pr, 0);
    printf("name=%s sms=%d l2=%d MB smem_optin=%zu regs/sm=%d clock=%d kHz mem_clock=%d kHz bus=%d\n", pr.name,
           pr.multiProcessorCount, pr.l2CacheSize >> 20, pr.sharedMemPerBlockOptin, pr.regsPerMultiprocessor, pr.clockRate,
           pr.memoryClockRate, pr.memoryBusWidth);
    const int sizes[] = {1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 14, 16};
    const int smems[] = {160 * 1024, 200 * 1024, 227 * 1024};
    const int threads[] = {512};
    cudaFuncSetAttribute(dummy_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    for (int cs : sizes)
        for (int sm : smems)
            for (int th : threads) {
                cudaFuncSetAttribute(dummy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3(cs * 64);
                cfg.blockDim = dim3(th);
                cfg.dynamicSmemBytes = sm;
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeClusterDimension;
                at[0].val.clusterDim.x = cs;
                at[0].val.clusterDim.y = 1;
                at[0].val.clusterDim.z = 1;
                cfg.attrs = at;
                cfg.numAttrs = 1;
                int n = -1;
                cudaError_t e = cudaOccupancyMaxActiveClusters(&n, dummy_kernel, &cfg);
                printf("cluster=%2d smem=%3dKB threads=%4d -> max active clusters %d (%d CTAs)%s\n", cs, sm >> 10, th, n, n * cs,
                       e == cudaSuccess ? "" : cudaGetErrorString(e));
                cudaGetLastError();
            }
    return 0;
}
