// Micro-benchmarks that size the MLP kernel's design (run on the B200 via gpurun; results in profiles/r01_ubench.txt):
//   1. L2 -> shared memory bulk-copy (cp.async.bulk) bandwidth per SM when all 148 CTAs stream the SAME weight blob
//   2. tcgen05.mma issue rate, cta_group::1, M=128, for N in {64,128,256}, A from shared memory (SS) or TMEM (TS)
//   3. both at once (interference between operand fetch and the weight stream)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/ubench tools/ubench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
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
__device__ __forceinline__ void tc_commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = (uint64_t)((saddr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

struct Params {
    const uint8_t* blob;     // weight blob in L2
    uint32_t blob_bytes;
    uint32_t slab_bytes;     // bytes per bulk copy
    int n_stage;
    int n_copies;            // copies per CTA (0: no streaming)
    int n_mma;               // MMAs per CTA (0: no MMA)
    int mma_n;               // N of each MMA
    int a_tmem;              // 1: A operand from TMEM
    int mma_per_slab;        // when both: MMAs issued per consumed slab
    long long* cycles;       // [grid] cycles of the measured region
    int stagger;             // start offset between CTAs in slabs
};

// warp 0 lane 0: producer; warp 1 lane 0: consumer / MMA issuer
__global__ void __launch_bounds__(64, 1) ubench_kernel(const Params p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t full[16], empty[16], done;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t ring = smem_u32(smem) + 65536;   // first 64 KB: fixed MMA operands (A 32 KB, B 32 KB)
    if (tid == 0) {
        for (int s = 0; s < 16; s++) { mbar_init(smem_u32(&full[s]), 1); mbar_init(smem_u32(&empty[s]), 1); }
        mbar_init(smem_u32(&done), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < 65536 / 4; i += 64) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // fp16 1.0
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    long long t0 = clock64();
    bool ok = true;
    if (warp == 0 && (tid & 31) == 0 && p.n_copies > 0) {
        uint32_t stage = 0, phase = 0;
        uint32_t off = (uint32_t)(((long long)blockIdx.x * p.stagger * p.slab_bytes) % p.blob_bytes);
        for (int i = 0; i < p.n_copies && ok; i++) {
            ok = mbar_wait(smem_u32(&empty[stage]), phase ^ 1);
            mbar_expect_tx(smem_u32(&full[stage]), p.slab_bytes);
            bulk_g2s(ring + stage * p.slab_bytes, p.blob + off, p.slab_bytes, smem_u32(&full[stage]));
            off += p.slab_bytes;
            if (off + p.slab_bytes > p.blob_bytes) off = 0;
            if (++stage == (uint32_t)p.n_stage) { stage = 0; phase ^= 1; }
        }
    } else if (warp == 1 && (tid & 31) == 0) {
        uint32_t stage = 0, phase = 0;
        const uint32_t idesc = (1u << 4) | ((uint32_t)(p.mma_n >> 3) << 17) | (8u << 24);
        const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem) + 32768;
        int issued = 0;
        const int n_iter = p.n_copies > 0 ? p.n_copies : 1;
        const int per = p.n_copies > 0 ? p.mma_per_slab : p.n_mma;
        for (int i = 0; i < n_iter && ok; i++) {
            if (p.n_copies > 0) ok = mbar_wait(smem_u32(&full[stage]), phase);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            for (int j = 0; j < per; j++, issued++) {
                const uint32_t kk = (uint32_t)(issued & 7);
                const uint64_t bd = umma_desc(b_addr + kk * 2u * (uint32_t)p.mma_n * 16u, (uint32_t)p.mma_n * 16u, 128);
                if (p.a_tmem) mma_ts(tmem + 256, tmem + kk * 8, bd, idesc, 1);
                else mma_ss(tmem + 256, umma_desc(a_addr + kk * 4096u, 2048, 128), bd, idesc, 1);
            }
            if (p.n_copies > 0) {
                tc_commit(smem_u32(&empty[stage]));
                if (++stage == (uint32_t)p.n_stage) { stage = 0; phase ^= 1; }
            }
        }
        tc_commit(smem_u32(&done));
        ok = ok && mbar_wait(smem_u32(&done), 0);
    }
    long long t1 = clock64();
    __syncthreads();
    if (warp == 1 && (tid & 31) == 0) p.cycles[blockIdx.x] = ok ? (t1 - t0) : -1;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

static int run(const char* name, Params p, int grid, int smem_bytes) {
    long long* d_cycles;
    CK(cudaMalloc(&d_cycles, grid * sizeof(long long)));
    p.cycles = d_cycles;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    ubench_kernel<<<grid, 64, smem_bytes>>>(p);     // warm-up (brings the blob into L2)
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    ubench_kernel<<<grid, 64, smem_bytes>>>(p);
    CK(cudaEventRecord(b));
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, a, b);
    std::vector<long long> c(grid);
    CK(cudaMemcpy(c.data(), d_cycles, grid * sizeof(long long), cudaMemcpyDeviceToHost));
    double avg = 0; long long mx = 0; bool bad = false;
    for (auto v : c) { if (v < 0) bad = true; avg += v; mx = v > mx ? v : mx; }
    avg /= grid;
    const double bytes = (double)p.n_copies * p.slab_bytes;
    const double macs = (double)(p.n_copies > 0 && p.n_mma ? (double)p.n_copies * p.mma_per_slab : p.n_mma) * 128.0 * p.mma_n * 16.0;
    printf("%-46s grid=%3d ms=%8.3f cyc(avg)=%10.0f cyc(max)=%10lld  B/cyc/SM=%7.2f  chip TB/s=%6.2f  MAC/cyc/SM=%7.1f cyc/MMA=%6.1f %s\n",
           name, grid, ms, avg, mx, bytes / avg, bytes * grid / (ms * 1e-3) / 1e12, macs / avg,
           macs > 0 ? avg / (macs / (128.0 * p.mma_n * 16.0)) : 0.0, bad ? "TIMEOUT" : "");
    cudaFree(d_cycles);
    return 0;
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    printf("%s, %d SMs, clock %d kHz\n", prop.name, prop.multiProcessorCount, prop.clockRate);
    const int grid = prop.multiProcessorCount;
    uint8_t* blob;
    const uint32_t blob_max = 16u << 20;
    CK(cudaMalloc(&blob, blob_max));
    CK(cudaMemset(blob, 0x3c, blob_max));
    const int smem_bytes = 65536 + 8 * 16384 + 1024;
    CK(cudaFuncSetAttribute(ubench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    char name[128];
    // 1. streaming only
    for (uint32_t blob_bytes : {256u << 10, 9u << 20}) {
        for (uint32_t slab : {4096u, 8192u, 16384u, 32768u}) {
            for (int st : {2, 4, 8}) {
                if ((uint32_t)st * slab > 8 * 16384) continue;
                for (int stagger : {0, 1}) {
                    Params p{}; p.blob = blob; p.blob_bytes = blob_bytes; p.slab_bytes = slab; p.n_stage = st;
                    p.n_copies = (int)((32u << 20) / slab); p.n_mma = 0; p.mma_n = 256; p.stagger = stagger;
                    snprintf(name, sizeof(name), "stream blob=%4uK slab=%2uK stages=%d stagger=%d", blob_bytes >> 10, slab >> 10, st, stagger);
                    if (run(name, p, grid, smem_bytes)) return 1;
                }
            }
        }
    }
    // fewer CTAs streaming (is it a per-SM or a chip limit?)
    for (int g : {1, 16, 74}) {
        Params p{}; p.blob = blob; p.blob_bytes = 256u << 10; p.slab_bytes = 16384; p.n_stage = 8;
        p.n_copies = 2048; p.mma_n = 256;
        snprintf(name, sizeof(name), "stream blob= 256K slab=16K stages=8 grid=%d", g);
        if (run(name, p, g, smem_bytes)) return 1;
    }
    // 2. MMA only
    for (int a_tmem : {0, 1}) {
        for (int n : {64, 128, 256}) {
            Params p{}; p.blob = blob; p.blob_bytes = 256u << 10; p.slab_bytes = 16384; p.n_stage = 8;
            p.n_copies = 0; p.n_mma = 8192; p.mma_n = n; p.a_tmem = a_tmem;
            snprintf(name, sizeof(name), "mma only N=%3d A=%s", n, a_tmem ? "tmem" : "smem");
            if (run(name, p, grid, smem_bytes)) return 1;
        }
    }
    // 3. both: each 16 KB slab feeds `per` MMAs (per=4 at N=256 is one tile per slab pass; 8 = two tiles; 16 = four)
    for (int a_tmem : {0, 1}) {
        for (int n : {128, 256}) {
            for (int per : {2, 4, 8, 16}) {
                Params p{}; p.blob = blob; p.blob_bytes = 256u << 10; p.slab_bytes = 16384; p.n_stage = 8;
                p.n_copies = 2048; p.n_mma = 1; p.mma_n = n; p.a_tmem = a_tmem; p.mma_per_slab = per * (256 / n);
                snprintf(name, sizeof(name), "stream+mma N=%3d A=%s mma/slab=%2d", n, a_tmem ? "tmem" : "smem", p.mma_per_slab);
                if (run(name, p, grid, smem_bytes)) return 1;
            }
        }
    }
    return 0;
}
