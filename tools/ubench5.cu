// Micro-benchmark 5: interference between a running tcgen05.mma stream and the epilogue's memory instructions on ONE SM.
// One warp issues back-to-back M=128 MMAs (A from shared memory or from TMEM, N = 128 / 256) while 8 epilogue warps loop
// over one kind of access (tcgen05.ld, ld + convert + tcgen05.st, st.shared, ld.shared).  Reports cycles per MMA and
// epilogue iterations per 1000 cycles, alone and together.  Results: profiles/r02_ubench_v5_interference.txt
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench5 tools/ubench5.cu && /tmp/ubench5
#include <cuda_fp16.h>
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
#define R8(v, o) "=r"(v[o + 0]), "=r"(v[o + 1]), "=r"(v[o + 2]), "=r"(v[o + 3]), "=r"(v[o + 4]), "=r"(v[o + 5]), "=r"(v[o + 6]), "=r"(v[o + 7])
#define W8(v, o) "r"(v[o + 0]), "r"(v[o + 1]), "r"(v[o + 2]), "r"(v[o + 3]), "r"(v[o + 4]), "r"(v[o + 5]), "r"(v[o + 6]), "r"(v[o + 7])
__device__ __forceinline__ void ld32(uint32_t a, uint32_t (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : R8(v, 0), R8(v, 8), R8(v, 16), R8(v, 24) : "r"(a) : "memory");
}
__device__ __forceinline__ void st16(uint32_t a, const uint32_t (&v)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(a), W8(v, 0), W8(v, 8) : "memory");
}
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t pack_relu(float a, float b) { uint32_t w; asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(w) : "f"(b), "f"(a)); return w; }

// mma_mode: 0 none (the issuer just waits `span` cycles), 1 SS N=128, 2 TS N=128 (A packed fp16 in TMEM), 3 SS N=256
// epi_mode: 0 none, 1 ld.x32 + wait, 2 ld.x32 + wait + relu/cvt + st.x16 + wait, 3 8 x st.shared.v4, 4 8 x ld.shared.v4, 5 2 x ld.x32 + wait
__global__ void __launch_bounds__(288, 1) k(int mma_mode, int epi_mode, int n_groups, long long span, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t done;
    __shared__ uint32_t tmem_base_s;
    __shared__ volatile int stop;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&done)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        stop = 0;
    }
    for (int i = tid; i < 131072 / 4; i += 288) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    if (warp == 8) {
        const int n = mma_mode == 3 ? 256 : 128;
        const uint32_t idesc = (1u << 4) | ((uint32_t)(n >> 3) << 17) | (8u << 24);
        const uint64_t ad = umma_desc(smem_u32(smem), 2048, 128), bd = umma_desc(smem_u32(smem) + 32768, n * 16, 128);
        const long long t0 = clock64();
        if (mma_mode == 0) {
            while (clock64() - t0 < span) {}
        } else {
            for (int g = 0; g < n_groups; g++) {
                if (elect_one()) {
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        if (mma_mode == 2)
                            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem), "r"(tmem + 256 + i * 8), "l"(bd + (uint64_t)(i * 2 * n)), "r"(idesc), "r"(1u) : "memory");
                        else
                            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(ad + (uint64_t)(i * 256)), "l"(bd + (uint64_t)(i * 2 * n)), "r"(idesc), "r"(1u) : "memory");
                    }
                }
                __syncwarp();
            }
            if (elect_one()) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&done)) : "memory");
            __syncwarp();
            while (!try_wait(smem_u32(&done), 0)) {}
        }
        const long long t1 = clock64();
        stop = 1;
        if ((tid & 31) == 0) out[0] = t1 - t0;
    } else {
        const int q = warp & 3, cg = warp >> 2;
        const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
        const uint32_t sa = smem_u32(smem) + 65536 + (uint32_t)tid * 16u;
        long long iters = 0;
        uint32_t v[32], v2[32];
        float sink = 0.f;
        const long long t0 = clock64();
        if (epi_mode)
            while (!stop) {
                if (epi_mode == 1) { ld32(lane_base + 320 + cg * 32, v); wait_ld(); sink += __uint_as_float(v[0]); }
                else if (epi_mode == 5) { ld32(lane_base + 320 + cg * 32, v); ld32(lane_base + 384 + cg * 32, v2); wait_ld(); sink += __uint_as_float(v[0]) + __uint_as_float(v2[0]); }
                else if (epi_mode == 2) {
                    ld32(lane_base + 320 + cg * 32, v); wait_ld();
                    uint32_t w[16];
#pragma unroll
                    for (int j = 0; j < 16; j++) w[j] = pack_relu(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
                    st16(lane_base + 448 + cg * 16, w); wait_st();
                } else if (epi_mode == 3) {
#pragma unroll
                    for (int j = 0; j < 8; j++) asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(sa + j * 4096), "r"((uint32_t)iters) : "memory");
                } else if (epi_mode == 4) {
#pragma unroll
                    for (int j = 0; j < 8; j++) { float4 x; asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w) : "r"(sa + j * 4096)); sink += x.x; }
                }
                iters++;
            }
        const long long t1 = clock64();
        if ((tid & 31) == 0) { out[1 + warp * 2] = iters; out[2 + warp * 2] = t1 - t0; }
        if (sink == 123.456f) out[40] = 1;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}
int main() {
    long long* d; CK(cudaMalloc(&d, 64 * 8));
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    const char* mn[] = {"no MMA", "MMA SS N=128", "MMA TS N=128", "MMA SS N=256"};
    const char* en[] = {"-", "ld.x32+wait", "ld.x32+cvt+st.x16+waits", "8 x st.shared.v4", "8 x ld.shared.v4", "2 x ld.x32+wait"};
    const int n_groups = 512;      // x 8 MMAs
    for (int m = 0; m < 4; m++)
        for (int e = 0; e < 6; e++) {
            if (m == 0 && e == 0) continue;
            for (int rep = 0; rep < 2; rep++) {
                CK(cudaMemset(d, 0, 64 * 8));
                k<<<1, 288, 200 * 1024>>>(m, e, n_groups, 300000, d);
                CK(cudaDeviceSynchronize());
            }
            long long h[64]; CK(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
            long long it = 0; double cyc = 0;
            for (int w = 0; w < 8; w++) { it += h[1 + 2 * w]; cyc += (double)h[2 + 2 * w] / 8; }
            printf("%-14s | %-26s | MMA: %7.1f cycles each (ideal %3d) | epilogue: %7.2f warp-iterations / 1000 cycles (%6.1f cycles per iteration per warp)\n", mn[m], en[e],
                   m ? (double)h[0] / (n_groups * 8) : 0.0, m == 3 ? 128 : 64, e ? 1000.0 * it / cyc : 0.0, e && it ? 8.0 * cyc / it : 0.0);
        }
    return 0;
}
