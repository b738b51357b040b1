// Micro-benchmark 4: is tcgen05.mma issue blocking?  Timestamps after every MMA issue, after the commit, after completion.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() { uint32_t pred; asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred)); return pred != 0; }
__device__ __forceinline__ bool try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | ((uint64_t)1 << 46);
}
__global__ void __launch_bounds__(32, 1) k(int n_mma_n, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t done, done2;
    __shared__ uint32_t tmem_base_s;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&done)) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&done2)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 65536 / 4; i += 32) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncwarp();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    const uint32_t idesc = (1u << 4) | ((uint32_t)(n_mma_n >> 3) << 17) | (8u << 24);
    const uint64_t ad = umma_desc(smem_u32(smem), 2048, 128), bd = umma_desc(smem_u32(smem) + 32768, n_mma_n * 16, 128);
    long long t[40];
    for (int rep = 0; rep < 2; rep++) {     // rep 0 warms up
        int n = 0;
        t[n++] = clock64();
#pragma unroll
        for (int i = 0; i < 16; i++) {
            if (elect_one())
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem + 256), "l"(ad + (uint64_t)((i & 7) * 256)), "l"(bd + (uint64_t)((i & 7) * 2 * n_mma_n)), "r"(idesc), "r"(1u) : "memory");
            __syncwarp();
            t[n++] = clock64();
        }
        if (elect_one()) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(rep ? &done2 : &done)) : "memory");
        __syncwarp();
        t[n++] = clock64();
        int polls = 0;
        while (!try_wait(smem_u32(rep ? &done2 : &done), 0)) polls++;
        t[n++] = clock64();
        // an already-complete try_wait
        try_wait(smem_u32(rep ? &done2 : &done), 0);
        t[n++] = clock64();
        if (rep == 1 && threadIdx.x == 0) { for (int i = 0; i < n; i++) out[i] = t[i]; out[39] = polls; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}
int main() {
    long long* d; CK(cudaMalloc(&d, 40 * 8));
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    for (int n : {64, 128, 256}) {
        k<<<1, 32, 100 * 1024>>>(n, d);
        CK(cudaDeviceSynchronize());
        long long h[40]; CK(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
        printf("N=%3d  issue deltas:", n);
        for (int i = 1; i <= 16; i++) printf(" %lld", h[i] - h[i - 1]);
        printf(" | commit %lld | wait-for-completion %lld (polls %lld) | completed try_wait %lld | total %lld\n", h[17] - h[16], h[18] - h[17], h[39], h[19] - h[18], h[18] - h[0]);
    }
    return 0;
}
