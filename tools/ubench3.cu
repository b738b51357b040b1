// Micro-benchmark 3: TMEM -> register load throughput (tcgen05.ld) per SM: shapes, waits per load, 4 vs 8 warps,
// and the fp32 -> fp16x2 convert + tcgen05.st / st.shared write-back cost.  Results: profiles/r01_ubench_v3.txt
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
#define R8(v, o) "=r"(v[o + 0]), "=r"(v[o + 1]), "=r"(v[o + 2]), "=r"(v[o + 3]), "=r"(v[o + 4]), "=r"(v[o + 5]), "=r"(v[o + 6]), "=r"(v[o + 7])
__device__ __forceinline__ void ld32(uint32_t a, uint32_t (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : R8(v, 0), R8(v, 8), R8(v, 16), R8(v, 24) : "r"(a) : "memory");
}
__device__ __forceinline__ void ld16(uint32_t a, uint32_t (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : R8(v, 0), R8(v, 8) : "r"(a) : "memory");
}
__device__ __forceinline__ void ld8(uint32_t a, uint32_t (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : R8(v, 0) : "r"(a) : "memory");
}
// 16 lanes x 256 bit shape: .x8 -> 32 registers per thread, covers 16 lanes... (each warp still only its own quarter)
__device__ __forceinline__ void ld16x256(uint32_t a, uint32_t (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : R8(v, 0), R8(v, 8), R8(v, 16), R8(v, 24) : "r"(a) : "memory");
}
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t pack(float a, float b) { __half2 h = __hmax2(__floats2half2_rn(a, b), __float2half2_rn(0.f)); return *reinterpret_cast<uint32_t*>(&h); }

// mode 0: ld32 + wait each;  1: two ld32 in flight;  2: ld16 + wait each;  3: 16x256b.x8;  4: ld32 + convert + tcgen05.st x16;
// 5: ld32 + convert + st.shared;  6: ld8 + wait each;  7: four ld32 in flight (128 regs)
__global__ void __launch_bounds__(256, 1) k(int mode, int iters, long long* cycles, uint32_t* sink) {
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(16) uint32_t stage[8 * 32 * 64];
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = tmem_base_s + ((uint32_t)((warp & 3) * 32) << 16) + (warp >= 4 ? 256u : 0u);
    uint32_t acc = 0;
    uint32_t v[32], v2[32], v3[32], v4[32];
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        const uint32_t col = (uint32_t)(it & 3) * 32u;
        if (mode == 0) { ld32(base + col, v); wait_ld(); acc += v[0] ^ v[31]; }
        else if (mode == 1) { ld32(base + col, v); ld32(base + ((col + 128) & 255), v2); wait_ld(); acc += v[0] ^ v2[31]; }
        else if (mode == 2) { ld16(base + col, v); wait_ld(); acc += v[0] ^ v[15]; }
        else if (mode == 3) { ld16x256(base + col, v); wait_ld(); acc += v[0] ^ v[31]; }
        else if (mode == 4) {
            ld32(base + col, v); wait_ld();
            uint32_t w[16];
#pragma unroll
            for (int j = 0; j < 16; j++) w[j] = pack(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
            asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(base + (col >> 1)),
                         "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]), "r"(w[8]), "r"(w[9]), "r"(w[10]), "r"(w[11]), "r"(w[12]), "r"(w[13]), "r"(w[14]), "r"(w[15]) : "memory");
        } else if (mode == 5) {
            ld32(base + col, v); wait_ld();
#pragma unroll
            for (int i = 0; i < 4; i++) {
                uint4 q;
                q.x = pack(__uint_as_float(v[8 * i]), __uint_as_float(v[8 * i + 1])); q.y = pack(__uint_as_float(v[8 * i + 2]), __uint_as_float(v[8 * i + 3]));
                q.z = pack(__uint_as_float(v[8 * i + 4]), __uint_as_float(v[8 * i + 5])); q.w = pack(__uint_as_float(v[8 * i + 6]), __uint_as_float(v[8 * i + 7]));
                reinterpret_cast<uint4*>(stage)[(i * 256 + threadIdx.x) & 4095] = q;
            }
        } else if (mode == 6) { ld8(base + col, v); wait_ld(); acc += v[0] ^ v[7]; }
        else { ld32(base, v); ld32(base + 32, v2); ld32(base + 64, v3); ld32(base + 96, v4); wait_ld(); acc += v[0] ^ v2[1] ^ v3[2] ^ v4[3]; }
    }
    wait_st();
    const long long t1 = clock64();
    if (threadIdx.x % 32 == 0) cycles[blockIdx.x * 8 + warp] = t1 - t0;
    if (acc == 0x12345678) sink[0] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base_s), "r"(512) : "memory");
}

int main() {
    long long* d; uint32_t* sink;
    CK(cudaMalloc(&d, 148 * 8 * sizeof(long long))); CK(cudaMalloc(&sink, 4));
    const char* names[] = {"ld.x32 + wait", "2 x ld.x32 + wait", "ld.x16 + wait", "ld.16x256b.x8 + wait", "ld.x32 + relu/cvt + tcgen05.st.x16",
                           "ld.x32 + relu/cvt + st.shared.v4", "ld.x8 + wait", "4 x ld.x32 + wait"};
    const int cols[] = {32, 64, 16, 32, 32, 32, 8, 128};
    for (int warps : {4, 8}) for (int mode = 0; mode < 8; mode++) {
        const int iters = 4096;
        k<<<148, warps * 32>>>(mode, iters, d, sink);
        CK(cudaDeviceSynchronize());
        std::vector<long long> h(148 * 8);
        CK(cudaMemcpy(h.data(), d, h.size() * 8, cudaMemcpyDeviceToHost));
        double avg = 0; for (int b = 0; b < 148; b++) for (int w = 0; w < warps; w++) avg += h[b * 8 + w]; avg /= 148 * warps;
        const double bytes = (double)iters * cols[mode] * 4 * 32 * warps;
        printf("warps=%d %-40s cycles/iter=%7.1f  TMEM read B/cyc/SM=%7.1f  cycles per 128x128 fp32 tile=%7.0f\n", warps, names[mode], avg / iters, bytes / avg,
               65536.0 / (bytes / avg));
    }
    return 0;
}
