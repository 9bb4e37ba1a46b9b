// kd_core.cu — error plumbing, version and device queries of libkd_b200.so.
#include <stdarg.h>
#include <stdio.h>

#include "kd_common.cuh"

namespace kd {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
    set_error("%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
    return KD_ERR_CUDA;
}

int cluster_capacity(int cluster_size, int threads, size_t smem_bytes);   // logit_kd_cluster.cu

}  // namespace kd

extern "C" {

int kd_abi_version(void) { return KD_B200_ABI_VERSION; }

const char* kd_last_error(void) { return kd::g_error; }

int kd_device_info(int* sm_count, int* max_smem_per_block, int* cluster4_max_active, int* cluster8_max_active) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return kd::cuda_fail(e, "cudaGetDevice");
    int sms = 0, smem = 0;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return kd::cuda_fail(e, "cudaDeviceGetAttribute");
    e = cudaDeviceGetAttribute(&smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (e != cudaSuccess) return kd::cuda_fail(e, "cudaDeviceGetAttribute");
    if (sm_count) *sm_count = sms;
    if (max_smem_per_block) *max_smem_per_block = smem;
    if (cluster4_max_active) *cluster4_max_active = kd::cluster_capacity(4, 512, 160 * 1024);
    if (cluster8_max_active) *cluster8_max_active = kd::cluster_capacity(8, 512, 226 * 1024);
    return KD_OK;
}

}  // extern "C"
