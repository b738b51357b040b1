// Micro-benchmarks, second set (results: profiles/r01_ubench_v2.txt):
//   A. true L2 -> shared bulk-copy bandwidth per SM (slot released with a plain mbarrier arrive, not tcgen05.commit)
//   B. tcgen05.mma rate per SM vs: N, A from smem/TMEM, 1 or 2 accumulators, no-swizzle vs 128B-swizzle operand
//      layout, cta_group::1 (M=128) vs cta_group::2 (M=256 over a CTA pair)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/ubench2 tools/ubench2.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
    for (int i = 0; i < 2000000; i++) if (mbar_try_wait(bar, parity)) return true;
    return false;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    uint64_t d = (uint64_t)((saddr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout << 61;
    return d;
}
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

struct Params {
    const uint8_t* blob;
    uint32_t blob_bytes, slab_bytes;
    int n_stage, n_copies;
    int n_mma, mma_n, a_tmem, n_acc, swz, n_lanes;
    long long* cycles;
};

template <int CG>
__global__ void __launch_bounds__(64, 1) ubench2_kernel(const Params p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t full[8], empty[8], done;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t rank = CG == 2 ? cluster_rank() : 0;
    const uint32_t ring = smem_u32(smem) + 98304;   // first 96 KB: fixed MMA operands (A 32 KB, B 64 KB)
    if (tid == 0) {
        for (int s = 0; s < 8; s++) { mbar_init(smem_u32(&full[s]), 1); mbar_init(smem_u32(&empty[s]), 1); }
        mbar_init(smem_u32(&done), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < 98304 / 4; i += 64) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (warp == 1) {
        if (CG == 1) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CG == 2) cluster_sync();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    long long t0 = clock64();
    bool ok = true;
    if (warp == 0 && (tid & 31) < p.n_lanes && p.n_copies > 0) {
        // self-consuming ring: slot s is refilled as soon as its previous copy has landed; n_lanes lanes each move
        // 1/n_lanes of every slab (parallel issue)
        uint32_t stage = 0, phase = 0, off = 0;
        const uint32_t lane = tid & 31, part = p.slab_bytes / p.n_lanes;
        for (int i = 0; i < p.n_copies && ok; i++) {
            if (i >= p.n_stage) ok = mbar_wait(smem_u32(&full[stage]), phase ^ 1);
            __syncwarp((1u << p.n_lanes) - 1);
            if (lane == 0) mbar_expect_tx(smem_u32(&full[stage]), p.slab_bytes);
            __syncwarp((1u << p.n_lanes) - 1);
            bulk_g2s(ring + stage * p.slab_bytes + lane * part, p.blob + off + lane * part, part, smem_u32(&full[stage]));
            off += p.slab_bytes;
            if (off + p.slab_bytes > p.blob_bytes) off = 0;
            if (++stage == (uint32_t)p.n_stage) { stage = 0; phase ^= 1; }
        }
        for (int s2 = 0; s2 < p.n_stage && ok; s2++) {   // drain
            ok = mbar_wait(smem_u32(&full[stage]), phase ^ 1);
            if (++stage == (uint32_t)p.n_stage) { stage = 0; phase ^= 1; }
        }
        if (p.n_mma == 0 && lane == 0) p.cycles[blockIdx.x] = ok ? (clock64() - t0) : -1;
    } else if (warp == 1 && (tid & 31) == 0) {
        if (p.n_mma > 0 && rank == 0) {
            const uint32_t m_dim = CG == 2 ? 16u : 8u;
            const uint32_t idesc = (1u << 4) | ((uint32_t)(p.mma_n >> 3) << 17) | (m_dim << 24);
            const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem) + 32768;
            const uint32_t nb = (uint32_t)(p.mma_n / CG);      // B rows held by this CTA
            // descriptors are precomputed; the loop body is 8 back-to-back MMAs (one K=128 pass) with ~2 instructions each
            uint64_t ad[8], bd[8];
            uint32_t at[8];
#pragma unroll
            for (uint32_t kk = 0; kk < 8; kk++) {
                if (p.swz) {
                    const uint32_t kb = kk >> 2, ks = kk & 3u;
                    ad[kk] = umma_desc(a_addr + kb * 16384u + ks * 32u, 16, 1024, 2);
                    bd[kk] = umma_desc(b_addr + kb * nb * 128u + ks * 32u, 16, 1024, 2);
                } else {
                    ad[kk] = umma_desc(a_addr + kk * 4096u, 2048, 128, 0);
                    bd[kk] = umma_desc(b_addr + kk * 2u * nb * 16u, nb * 16u, 128, 0);
                }
                at[kk] = tmem + kk * 8u;
            }
            const uint32_t d0 = tmem + 256, d1 = tmem + 256 + (p.n_acc == 2 ? 128u : 0u);
            for (int i = 0; i < p.n_mma; i += 16) {
#pragma unroll
                for (int kk = 0; kk < 16; kk++) {
                    const uint32_t d = kk < 8 ? d0 : d1;
                    if (CG == 1) {
                        if (p.a_tmem)
                            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(at[kk & 7]), "l"(bd[kk & 7]), "r"(idesc), "r"(1u) : "memory");
                        else
                            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(ad[kk & 7]), "l"(bd[kk & 7]), "r"(idesc), "r"(1u) : "memory");
                    } else {
                        if (p.a_tmem)
                            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(at[kk & 7]), "l"(bd[kk & 7]), "r"(idesc), "r"(1u) : "memory");
                        else
                            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(ad[kk & 7]), "l"(bd[kk & 7]), "r"(idesc), "r"(1u) : "memory");
                    }
                }
            }
            if (CG == 1)
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&done)) : "memory");
            else
                asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&done)), "h"((uint16_t)1) : "memory");
            ok = ok && mbar_wait(smem_u32(&done), 0);
        }
    }
    long long t1 = clock64();
    __syncthreads();
    if (warp == 1 && (tid & 31) == 0 && p.n_mma > 0) p.cycles[blockIdx.x] = ok ? (t1 - t0) : -1;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CG == 2) cluster_sync();
    if (warp == 1) {
        if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    }
}

template <int CG>
static int launch(Params p, int grid, int smem_bytes) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(64); cfg.dynamicSmemBytes = smem_bytes; cfg.stream = 0;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    CK(cudaLaunchKernelEx(&cfg, ubench2_kernel<CG>, p));
    return 0;
}

static int run(const char* name, Params p, int grid, int smem_bytes, int cg) {
    long long* d_cycles;
    CK(cudaMalloc(&d_cycles, grid * sizeof(long long)));
    CK(cudaMemset(d_cycles, 0, grid * sizeof(long long)));
    p.cycles = d_cycles;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    if (cg == 1 ? launch<1>(p, grid, smem_bytes) : launch<2>(p, grid, smem_bytes)) return 1;
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    if (cg == 1 ? launch<1>(p, grid, smem_bytes) : launch<2>(p, grid, smem_bytes)) return 1;
    CK(cudaEventRecord(b));
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, a, b);
    std::vector<long long> c(grid);
    CK(cudaMemcpy(c.data(), d_cycles, grid * sizeof(long long), cudaMemcpyDeviceToHost));
    double avg = 0; int cnt = 0; bool bad = false;
    for (int i = 0; i < grid; i++) { if (cg == 2 && (i & 1) && p.n_copies == 0) continue; if (c[i] < 0) bad = true; avg += c[i]; cnt++; }
    avg /= cnt;
    const double bytes = (double)p.n_copies * p.slab_bytes;
    const double macs_per_sm = (double)p.n_mma * 128.0 * p.mma_n * 16.0;    // per SM (a pair does 256 x N x 16 on two SMs)
    printf("%-58s ms=%7.3f cyc=%9.0f  B/cyc/SM=%7.2f  MAC/cyc/SM=%7.1f  cyc/MMA=%6.1f %s\n", name, ms, avg, bytes / avg,
           macs_per_sm / avg, p.n_mma ? avg / p.n_mma : 0.0, bad ? "TIMEOUT" : "");
    cudaFree(d_cycles);
    return 0;
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    printf("%s, %d SMs\n", prop.name, prop.multiProcessorCount);
    const int grid = prop.multiProcessorCount;
    uint8_t* blob;
    CK(cudaMalloc(&blob, 16u << 20));
    CK(cudaMemset(blob, 0x3c, 16u << 20));
    const int smem_bytes = 98304 + 8 * 16384 - 2048;
    CK(cudaFuncSetAttribute(ubench2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    CK(cudaFuncSetAttribute(ubench2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    char name[160];
    for (uint32_t blob_bytes : {256u << 10, 9u << 20}) {
        for (uint32_t slab : {4096u, 8192u, 16384u}) {
            for (int st : {2, 4, 7}) for (int nl : {1, 4}) {
                Params p{}; p.blob = blob; p.blob_bytes = blob_bytes; p.slab_bytes = slab; p.n_stage = st; p.n_lanes = nl;
                p.n_copies = (int)((32u << 20) / slab); p.mma_n = 256; p.n_acc = 1;
                snprintf(name, sizeof(name), "A stream blob=%4uK slab=%2uK stages=%d lanes=%d", blob_bytes >> 10, slab >> 10, st, nl);
                if (run(name, p, grid, smem_bytes, 1)) return 1;
            }
        }
    }
    for (int cg : {1, 2}) for (int swz : {0, 1}) for (int a_tmem : {0, 1}) for (int n : {64, 128, 256}) for (int n_acc : {1, 2}) {
        if (n_acc == 2 && n > 128) continue;
        if (swz && a_tmem && 0) continue;
        Params p{}; p.blob = blob; p.blob_bytes = 256u << 10; p.slab_bytes = 16384; p.n_stage = 7;
        p.n_copies = 0; p.n_lanes = 1; p.n_mma = 16384; p.mma_n = n; p.a_tmem = a_tmem; p.n_acc = n_acc; p.swz = swz;
        snprintf(name, sizeof(name), "B mma cta_group=%d M=%3d N=%3d A=%s layout=%s accumulators=%d", cg, 128 * cg, n,
                 a_tmem ? "tmem" : "smem", swz ? "sw128" : "none ", n_acc);
        if (run(name, p, grid, smem_bytes, cg)) return 1;
    }
    // C: MMA while the weight ring streams
    for (int cg : {1, 2}) for (int a_tmem : {0, 1}) {
        Params p{}; p.blob = blob; p.blob_bytes = 256u << 10; p.slab_bytes = 16384; p.n_stage = 7;
        p.n_copies = 2048; p.n_lanes = 4; p.n_mma = 16384; p.mma_n = 256; p.a_tmem = a_tmem; p.n_acc = 1; p.swz = 1;
        snprintf(name, sizeof(name), "C mma+stream(32MB) cta_group=%d N=256 A=%s sw128", cg, a_tmem ? "tmem" : "smem");
        if (run(name, p, grid, smem_bytes, cg)) return 1;
    }
    return 0;
}
