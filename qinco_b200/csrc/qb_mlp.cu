// Fused implicit-codebook MLP for one quantisation step, sm_100a (tcgen05 / TMEM / TMA bulk copy).
//
// Replaces, for a 128-row tile of candidates, the reference's chain of separate launches
// (reference qinco/model/qinco_base.py:262-280 = in_proj, QConcat :60-64, L x QBlockFFN :93-97, out_proj, skip;
// the distance of :343-345 and the x-hat update of :363-369) with ONE persistent kernel:
//
//   warps 0-7  epilogue (two warpgroups): thread (lane quarter q = warp % 4, lane) owns row 32 q + lane of BOTH tiles in
//              flight == TMEM lane; the two warps of a lane quarter split every column range.
//              init    e0 = T_m[code] + u_b  -> fp32 into the TMEM residual accumulator (tcgen05.st.x32)
//                                             -> fp16 into the A_E operand tile in shared memory
//              H-epi   relu(Hacc) -> packed fp16 written back IN PLACE into TMEM (A operand of the down-projection),
//                      handed to the issuer in two K halves
//              E-epi   Eacc -> fp16 -> A_E                               (per residual block)
//              final   score: dist = ||r_b - o||^2 (fp32)   apply: xhat_out = xhat_b + o
//   warp 8     producer: streams the pre-packed fp16 weight slabs with cp.async.bulk (TMA) into an mbarrier ring
//   warp 9-10  MMA issuers, one per tile slot: walk the op list (qb_plan.h, in the kernel-parameter constant bank) and
//              issue tcgen05.mma (M=128, kind::f16, fp32 accumulate in TMEM): up-projection A from shared memory,
//              down-projection A from TMEM; completion is signalled with tcgen05.commit
//   warp 11    idle (completes the service warpgroup for setmaxnreg)
//
// The service warps run converged and only predicate the asynchronous instructions with elect.sync, so descriptors stay
// in uniform registers.  An issuing warp retires ~1 instruction per 6.5 cycles, which makes the issuer's SASS instruction
// count (not the tensor pipe) the first bound: see mma_slab / the issuer loop.  The fp32 residual stream never leaves
// TMEM; activations never touch HBM.  kPair = true is the cta_group::2 variant (one M=256 MMA over a CTA pair, half of
// every weight slab per CTA); DESIGN.md section 5.1 has the measurements behind each of these choices.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <type_traits>
#include <vector>

#include "qb_dev.h"

namespace qb {

namespace {

#ifndef QB_EPI_WARPS
#define QB_EPI_WARPS 8
#endif
constexpr int kEpiWarps = QB_EPI_WARPS;   // 8 (16): two (four) warps per TMEM lane quarter: warp w owns lanes 32*(w%4).., column group w/4
constexpr int kColGroups = kEpiWarps / 4;
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kThreads = kEpiThreads + 128;  // + one warpgroup: producer warp, one MMA warp per tile slot, one idle warp
// Registers are a per-scheduler pool (16 K per SM sub-partition = 3 warps x 168 at launch).  The service warpgroup
// hands most of its share back (setmaxnreg.dec) and the epilogue warpgroups take it (setmaxnreg.inc): 2 x 232 + 40 = 504 <= 512.
constexpr int kEpiRegs = kEpiWarps == 8 ? 232 : 112, kSvcRegs = 40;      // 16 warps: 4 x 112 + 40 = 488 <= 512
// Epilogue -> issuer hand-off: every epilogue THREAD arrives on the barrier (count = kEpiThreads) right after fencing its own
// writes.  The elected form (warp converges, lane 0 arrives) costs an S2R / vote / predicate chain of ~5 dependent
// instructions on each of the ~11 hand-offs per tile, on the warps that bound the kernel.
#ifndef QB_ARRIVE_ALL
#define QB_ARRIVE_ALL 1
#endif
constexpr int kArriveCount = QB_ARRIVE_ALL ? kEpiThreads : kEpiWarps;
constexpr int kProducerWarp = kEpiWarps, kMmaWarp = kEpiWarps + 1;   // MMA warps: kMmaWarp + tile slot
constexpr int kAkcBytes = QB_TILE_M * 16;   // bytes of one 8-element k-chunk of an A operand tile
constexpr uint32_t kDescHi = (128u >> 4) | (1u << 14);   // high word of every UMMA smem descriptor here: SBO = 128 B, version 1

// ------------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ uint64_t globaltimer() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Bounded wait: a protocol bug must trap (and report) instead of hanging the GPU.  Fully inlined: a call here would
// make every live register of the epilogue (prefetched table rows) caller-saved, i.e. spilled around each wait.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, uint32_t* err_flag, uint32_t code) {
    uint32_t spins = 0;
    uint64_t t0 = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++spins & 1023u) == 0) {               // the timer is read only once in a while, off the fast path
            const uint64_t now = globaltimer();
            if (t0 == 0) t0 = now;
            if (now - t0 > 4000000000ull) {          // 4 s
                if (err_flag) atomicExch(err_flag, code);
                __threadfence_system();
                __trap();
            }
        }
    }
}

// Same for a warp whose latency does not matter (the select warp): it sleeps between polls instead of competing with the
// epilogue warps of its scheduler for issue slots.
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity, uint32_t* err_flag, uint32_t code) {
    uint32_t spins = 0;
    uint64_t t0 = 0;
    while (!mbar_try_wait(bar, parity)) {
        __nanosleep(400);
        if ((++spins & 255u) == 0) {
            const uint64_t now = globaltimer();
            if (t0 == 0) t0 = now;
            if (now - t0 > 4000000000ull) {
                if (err_flag) atomicExch(err_flag, code);
                __threadfence_system();
                __trap();
            }
        }
    }
}

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

// Multicast variant: the bytes land at the same CTA-relative offset in every CTA of `mask`, and each destination CTA's
// mbarrier (same offset) gets the complete_tx.
__device__ __forceinline__ void bulk_g2s_mcast(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t mask) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar), "h"(mask)
                 : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void proxy_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// commit of cta_group::1 MMAs whose arrival goes to the barrier at the same offset in BOTH CTAs of a 2-CTA cluster (weight
// multicast variant: a ring slot is recycled cluster-wide)
__device__ __forceinline__ void tc_commit_mcast(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"((uint16_t)3)
                 : "memory");
}

// ---- CTA pair (cta_group::2): one M=256 MMA spans the two CTAs of a cluster; each CTA holds its own 128 rows of A and
// of the accumulator and HALF of the B rows (weights), so weight traffic and weight shared-memory reads per SM halve.
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local_addr` (a shared::cta address) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_rank(uint32_t local_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void proxy_fence_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
// commit of a cta_group::2 MMA group: the arrival is multicast to the barrier at the same offset in both CTAs
__device__ __forceinline__ void tc_commit_pair(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"((uint16_t)3)
                 : "memory");
}
// One K=16 MMA whose operand descriptors advance in place (low descriptor word += step; the address field never carries
// into the LBO field).  Written so that ptxas keeps everything in uniform registers: 5 instructions per MMA when
// unrolled.  kPairI selects cta_group::2.
template <bool kPairI>
__device__ __forceinline__ void mma_ss_step(uint32_t d, uint32_t& alo, uint32_t& blo, uint32_t ahi, uint32_t bhi, uint32_t idesc,
                                            uint32_t acc, uint32_t as, uint32_t bs) {
    if (kPairI)
        asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %6, 0;\n\t"
                     "mov.b64 da, {%0, %3};\n\tmov.b64 db, {%1, %4};\n\t"
                     "tcgen05.mma.cta_group::2.kind::f16 [%2], da, db, %5, p;\n\t"
                     "add.u32 %0, %0, %7;\n\tadd.u32 %1, %1, %8;\n\t}"
                     : "+r"(alo), "+r"(blo) : "r"(d), "r"(ahi), "r"(bhi), "r"(idesc), "r"(acc), "r"(as), "r"(bs) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %6, 0;\n\t"
                     "mov.b64 da, {%0, %3};\n\tmov.b64 db, {%1, %4};\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%2], da, db, %5, p;\n\t"
                     "add.u32 %0, %0, %7;\n\tadd.u32 %1, %1, %8;\n\t}"
                     : "+r"(alo), "+r"(blo) : "r"(d), "r"(ahi), "r"(bhi), "r"(idesc), "r"(acc), "r"(as), "r"(bs) : "memory");
}
// same with the A operand in TMEM (column address += as per K=16: 8 for a contiguous packed operand)
template <bool kPairI>
__device__ __forceinline__ void mma_ts_step(uint32_t d, uint32_t& a_tmem, uint32_t& blo, uint32_t bhi, uint32_t idesc, uint32_t acc,
                                            uint32_t as, uint32_t bs) {
    if (kPairI)
        asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 db;\n\tsetp.ne.b32 p, %5, 0;\n\t"
                     "mov.b64 db, {%1, %3};\n\t"
                     "tcgen05.mma.cta_group::2.kind::f16 [%2], [%0], db, %4, p;\n\t"
                     "add.u32 %0, %0, %7;\n\tadd.u32 %1, %1, %6;\n\t}"
                     : "+r"(a_tmem), "+r"(blo) : "r"(d), "r"(bhi), "r"(idesc), "r"(acc), "r"(bs), "r"(as) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 db;\n\tsetp.ne.b32 p, %5, 0;\n\t"
                     "mov.b64 db, {%1, %3};\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%2], [%0], db, %4, p;\n\t"
                     "add.u32 %0, %0, %7;\n\tadd.u32 %1, %1, %6;\n\t}"
                     : "+r"(a_tmem), "+r"(blo) : "r"(d), "r"(bhi), "r"(idesc), "r"(acc), "r"(bs), "r"(as) : "memory");
}
// nk MMAs of one slab; the common slab depths are fully unrolled.  The A cursor advances by `as` after even and `as2`
// after odd k-steps of the slab (equal except for the blocked TMEM layout of a packed H operand, QbOp::a_blk32).
template <bool kPairI, bool kFromSmem>
__device__ __forceinline__ void mma_slab(int nk, uint32_t d, uint32_t& a, uint32_t& blo, uint32_t ahi, uint32_t bhi, uint32_t idesc,
                                         uint32_t acc, uint32_t as, uint32_t as2, uint32_t bs) {
    auto one = [&](uint32_t ac, int k) {
        if (kFromSmem) mma_ss_step<kPairI>(d, a, blo, ahi, bhi, idesc, ac, as, bs);
        else mma_ts_step<kPairI>(d, a, blo, bhi, idesc, ac, (k & 1) ? as2 : as, bs);
    };
    one(acc, 0);
    if (nk == 8) {
#pragma unroll
        for (int k = 1; k < 8; k++) one(1u, k);
    } else if (nk == 4) {
#pragma unroll
        for (int k = 1; k < 4; k++) one(1u, k);
    } else {
#pragma unroll 1
        for (int k = 1; k < nk; k++) one(1u, k);
    }
}

// kind::f16 instruction descriptor: D=f32, A=B=f16, both K-major, M=128 (cute/arch/mma_sm100_desc.hpp InstrDescriptor)
__device__ __forceinline__ uint32_t umma_idesc(uint32_t n, uint32_t m = QB_TILE_M) {
    return (1u << 4) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

#define QB_R8(v, o) "=r"(v[o + 0]), "=r"(v[o + 1]), "=r"(v[o + 2]), "=r"(v[o + 3]), "=r"(v[o + 4]), "=r"(v[o + 5]), "=r"(v[o + 6]), "=r"(v[o + 7])
#define QB_W8(v, o) "r"(v[o + 0]), "r"(v[o + 1]), "r"(v[o + 2]), "r"(v[o + 3]), "r"(v[o + 4]), "r"(v[o + 5]), "r"(v[o + 6]), "r"(v[o + 7])
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : QB_R8(v, 0), QB_R8(v, 8), QB_R8(v, 16), QB_R8(v, 24)
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : QB_R8(v, 0), QB_R8(v, 8)
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        QB_W8(v, 0), QB_W8(v, 8)
        : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        QB_W8(v, 0), QB_W8(v, 8), QB_W8(v, 16), QB_W8(v, 24)
        : "memory");
}
__device__ __forceinline__ void tmem_st16p(uint32_t taddr, const uint32_t* v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        QB_W8(v, 0), QB_W8(v, 8)
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]),
                 "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
// relu folded into the conversion (F2FP.RELU: one instruction per column pair instead of F2FP + HMNMX2)
__device__ __forceinline__ uint32_t pack_h2_relu(float a, float b) {
    uint32_t w;
    asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(w) : "f"(b), "f"(a));      // low half = a
    return w;
}
// Packed fp32 pairs (FADD2 / FFMA2: two IEEE round-to-nearest operations per issue slot, bit-identical to the scalar
// forms).  The epilogue warps are latency / issue bound, not FLOP bound, so halving the instruction count is the point.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 f2_pack(float lo, float hi) {
    f32x2 v;
    asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(lo), "f"(hi));
    return v;
}
__device__ __forceinline__ void f2_unpack(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 f2_add(f32x2 a, f32x2 b) {
    f32x2 c;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(c) : "l"(a), "l"(b));
    return c;
}
__device__ __forceinline__ f32x2 f2_sub(f32x2 a, f32x2 b) {
    f32x2 c;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(c) : "l"(a), "l"(b));
    return c;
}
__device__ __forceinline__ f32x2 f2_fma(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// explicit shared-space load (32-bit shared address): a C++ pointer to shared memory goes through a generic address,
// which the compiler rebuilds from a slow special-register read (see `opaque` in the kernel)
__device__ __forceinline__ float4 lds4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ float lds1(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts1(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// Loads nc (16 or 32) fp32 accumulator columns; the caller waits.  With nc == 16 the upper half is zeroed so that no
// register of `v` is ever read undefined (undefined reads stretch live ranges over the whole kernel).
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, int nc, uint32_t (&v)[32]) {
    __syncwarp();   // tcgen05.ld is .sync.aligned: reconverge after per-thread mbarrier spins / predicated code
    if (nc >= 32) {
        tmem_ld32(taddr, v);
    } else {
        tmem_ld16(taddr, v);
#pragma unroll
        for (int i = 16; i < 32; i++) v[i] = 0u;
    }
}

__device__ __forceinline__ void named_bar_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }

// Column range [c0, c1) of a width-W block (W a multiple of 16) owned by column group cg of kColGroups: 16-column units
// are dealt out as evenly as possible.
__device__ __forceinline__ void group_range(int W, int cg, int& c0, int& c1) {
    const int units = W >> 4, base = units / kColGroups, rem = units % kColGroups;
    const int u0 = cg * base + (cg < rem ? cg : rem);
    c0 = u0 << 4;
    c1 = (u0 + base + (cg < rem ? 1 : 0)) << 4;
}

// fp32 accumulator columns [c0, c1) at `taddr` -> fp16 -> A_E k-chunks (shared memory, `sdst` already includes the row);
// 64 columns per step with both tcgen05.ld of the step in flight together
__device__ __forceinline__ void acc_to_smem_operand(uint32_t taddr, int c0, int c1, uint32_t sdst) {
#pragma unroll 1
    for (int c = c0; c < c1; c += 64) {
        uint32_t va[32], vb[32];
        const int n = c1 - c;
        tmem_ld_cols(taddr + c, n, va);
        if (n > 32) tmem_ld_cols(taddr + c + 32, n - 32, vb);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 4; i++)
            if (i * 8 < n)
                st_shared_v4(sdst + (uint32_t)((c / 8 + i) * kAkcBytes),
                             pack_h2(__uint_as_float(va[8 * i]), __uint_as_float(va[8 * i + 1])),
                             pack_h2(__uint_as_float(va[8 * i + 2]), __uint_as_float(va[8 * i + 3])),
                             pack_h2(__uint_as_float(va[8 * i + 4]), __uint_as_float(va[8 * i + 5])),
                             pack_h2(__uint_as_float(va[8 * i + 6]), __uint_as_float(va[8 * i + 7])));
        if (n > 32) {
#pragma unroll
            for (int i = 0; i < 4; i++)
                if (32 + i * 8 < n)
                    st_shared_v4(sdst + (uint32_t)((c / 8 + 4 + i) * kAkcBytes),
                                 pack_h2(__uint_as_float(vb[8 * i]), __uint_as_float(vb[8 * i + 1])),
                                 pack_h2(__uint_as_float(vb[8 * i + 2]), __uint_as_float(vb[8 * i + 3])),
                                 pack_h2(__uint_as_float(vb[8 * i + 4]), __uint_as_float(vb[8 * i + 5])),
                                 pack_h2(__uint_as_float(vb[8 * i + 6]), __uint_as_float(vb[8 * i + 7])));
        }
    }
}

// fp32 Hacc columns [c0, c1) (c1 - c0 <= 64) at `taddr` -> relu -> packed fp16 pairs written back IN PLACE at columns
// [c0/2, c1/2).  The packed block of a higher column quarter lands on fp32 columns a lower quarter's thread still has
// to read, so the warps sharing a lane quarter first pull their whole range into registers, meet at a named barrier
// and only then write.
__device__ __forceinline__ void acc_to_tmem_operand(uint32_t taddr, int c0, int c1, int quarter_bar) {
    uint32_t va[32], vb[32];
    const int n = c1 - c0;
    if (n > 0) tmem_ld_cols(taddr + c0, n, va);
    if (n > 32) tmem_ld_cols(taddr + c0 + 32, n - 32, vb);
    tmem_wait_ld();
    named_bar_sync(quarter_bar, kColGroups * 32);
    if (n >= 64) {          // common case: one 32-column store of the 64 packed values
        uint32_t w[32];
#pragma unroll
        for (int j = 0; j < 16; j++) {
            w[j] = pack_h2_relu(__uint_as_float(va[2 * j]), __uint_as_float(va[2 * j + 1]));
            w[16 + j] = pack_h2_relu(__uint_as_float(vb[2 * j]), __uint_as_float(vb[2 * j + 1]));
        }
        __syncwarp();
        tmem_st32(taddr + (c0 >> 1), w);
    } else {
        if (n > 0) {
            uint32_t w[16];
#pragma unroll
            for (int j = 0; j < 16; j++) w[j] = pack_h2_relu(__uint_as_float(va[2 * j]), __uint_as_float(va[2 * j + 1]));
            __syncwarp();
            if (n >= 32) tmem_st16(taddr + (c0 >> 1), w); else tmem_st8(taddr + (c0 >> 1), w);
        }
        if (n > 32) {
            uint32_t w[16];
#pragma unroll
            for (int j = 0; j < 16; j++) w[j] = pack_h2_relu(__uint_as_float(vb[2 * j]), __uint_as_float(vb[2 * j + 1]));
            __syncwarp();
            tmem_st8(taddr + ((c0 + 32) >> 1), w);
        }
    }
    tmem_wait_st();
}

// One K half of a split H chunk (width 2 * kColGroups * 32 = 128... in general cw, a multiple of 64): fp32 columns
// [h * cw/2 + cg * cw/4, + cw/4) of this thread's row -> relu -> packed fp16 written at columns [(h * cw/2 + cg * cw/4) / 2, ...).
// The packed block of the upper column group lands on fp32 columns the lower group reads in the same half, hence the
// barrier between the loads and the stores; across halves the targets of half 1 were all read in half 0.
// `blk32` (n == 32, c0 a multiple of 32; QbStepPlan::h_split == 2): the packed block goes to the START of the thread's own
// 32 fp32 columns, which nobody else reads -- no barrier, and the MMA issuer walks the blocked layout (QbOp::a_blk32).
__device__ __forceinline__ void acc_to_tmem_operand_half(uint32_t taddr, int c0, int n, int quarter_bar, bool blk32) {
    uint32_t va[32];          // n = 16 or 32 columns
    tmem_ld_cols(taddr + c0, n, va);
    tmem_wait_ld();
    if (!blk32) named_bar_sync(quarter_bar, kColGroups * 32);
    uint32_t w[16];
#pragma unroll
    for (int j = 0; j < 16; j++) w[j] = pack_h2_relu(__uint_as_float(va[2 * j]), __uint_as_float(va[2 * j + 1]));
    __syncwarp();
    const uint32_t dst = taddr + (blk32 ? c0 : (c0 >> 1));
    if (n >= 32) tmem_st16(dst, w); else tmem_st8(dst, w);
    tmem_wait_st();
}

// Both K halves of a split H chunk: the two tcgen05.ld travel together (one TMEM round trip of ~130 cycles instead of two, one
// quarter barrier instead of two), then half 0 is packed, stored and handed to the issuer (`ready(0)`) before half 1 is
// packed.  Every packed target column lies in the fp32 columns [0, cw/2) of the chunk, all of which both warps of the lane
// quarter have read before the barrier.
#ifndef QB_HSPLIT_ONE_WAIT
#define QB_HSPLIT_ONE_WAIT 0      // 1 (A-B): both halves stored before the one tcgen05.wait::st, both hand-offs after it
#endif
template <class Ready>
__device__ __forceinline__ void acc_to_tmem_operand_split(uint32_t taddr, int cw, int cg, int quarter_bar, Ready&& ready) {
    const int qw = cw >> 2;            // 16 or 32 columns per (half, column group)
    uint32_t va[32], vb[32];
    tmem_ld_cols(taddr + cg * qw, qw, va);
    tmem_ld_cols(taddr + (cw >> 1) + cg * qw, qw, vb);
    tmem_wait_ld();
    named_bar_sync(quarter_bar, kColGroups * 32);
    uint32_t w[16];
#pragma unroll
    for (int j = 0; j < 16; j++) w[j] = pack_h2_relu(__uint_as_float(va[2 * j]), __uint_as_float(va[2 * j + 1]));
    __syncwarp();
    if (qw >= 32) tmem_st16(taddr + ((cg * qw) >> 1), w); else tmem_st8(taddr + ((cg * qw) >> 1), w);
#if !QB_HSPLIT_ONE_WAIT
    tmem_wait_st();
    ready(0);
#endif
#pragma unroll
    for (int j = 0; j < 16; j++) w[j] = pack_h2_relu(__uint_as_float(vb[2 * j]), __uint_as_float(vb[2 * j + 1]));
    __syncwarp();
    if (qw >= 32) tmem_st16(taddr + (((cw >> 1) + cg * qw) >> 1), w); else tmem_st8(taddr + (((cw >> 1) + cg * qw) >> 1), w);
    tmem_wait_st();
#if QB_HSPLIT_ONE_WAIT
    ready(0);
#endif
    ready(1);
}

// Debug event log (MlpParams::trace, CTA 0 only): role 0 = epilogue thread 0, 1 = MMA issuer, 2 = producer.
#ifdef QB_ENABLE_TRACE
struct Tracer {
    unsigned long long* buf;
    int n;
    bool writer;
    __device__ __forceinline__ void init(const MlpParams& p, int role, bool active) {
        buf = (p.trace && blockIdx.x == 0) ? p.trace + role * QB_TRACE_EVENTS : nullptr;    // warp-uniform: off = one uniform branch
        writer = active;
        n = 0;
    }
    __device__ __forceinline__ void ev(uint32_t id) {
        if (buf) {
            if (writer && n < QB_TRACE_EVENTS) buf[n++] = ((unsigned long long)clock64() << 16) | id;
        }
    }
};
#else
// The event log costs four registers and a branch per event in every role: compiled in only with -DQB_ENABLE_TRACE
// (QB_MLP_TRACE then dumps the timeline; a default build writes an empty one).
struct Tracer {
    __device__ __forceinline__ void init(const MlpParams&, int, bool) {}
    __device__ __forceinline__ void ev(uint32_t) {}
};
#endif

}  // namespace

// kMcast: weight multicast over a 2-CTA cluster.  Both CTAs walk the same slab sequence; each streams HALF of every slab
// and multicasts it into the ring slot of both (cp.async.bulk ... .multicast::cluster), so a slab crosses L2 -> SM once per
// two CTAs while every MMA stays local (cta_group::1, no cross-CTA accumulator traffic as in the pair kernel).  A ring
// slot is recycled when the MMA warps of BOTH CTAs have committed it (multicast tcgen05.commit).
// Plan accessors.  The generic view reads the plan from the kernel-parameter bank.  The S128 view -- QINCo / QINCo2-S at
// d = de = 128, dh = 256, K = 256 (BASELINE configs 1-2), exactly the plan make_step_plan() returns for it -- answers with
// compile-time constants, so the epilogue's addresses become immediates and its shape branches disappear: between an
// accumulator barrier and the first tcgen05.ld the generic code spends ~26 dependent uniform-datapath instructions
// (8 % of the epilogue's time in the H phase alone, ncu).  L and the skip flag stay run-time (QINCo1 vs QINCo2, L = 2 / 16).
#define QB_PLAN_FIELDS(X) X(D) X(De) X(Dh) X(K) X(has_proj) X(n_tiles) X(tmem_alloc_cols) X(n_ops_block) X(n_ops_out) X(hc) X(n_hchunk) \
    X(oc) X(n_ochunk) X(tmem_e_col) X(tmem_h_col) X(tmem_tile_cols) X(smem_tres) X(smem_ring) X(slot_bytes) X(n_stage) X(h_split)   \
    X(n_ops_pre) X(e_split) X(epart)
template <int kShape>
struct PlanView {
    const QbStepPlan& q;
#define X(f) __device__ __forceinline__ int f() const { return q.f; }
    QB_PLAN_FIELDS(X)
    X(L) X(skip)
#undef X
    __device__ __forceinline__ int smem_ae(int t) const { return q.smem_ae[t]; }
    __device__ __forceinline__ int64_t block_w_bytes() const { return q.block_w_bytes; }
};
struct PlanS128 {       // == make_step_plan(128, 128, 256, L, 256, ...) with default options (launch_mlp compares field by field)
    static constexpr int D = 128, De = 128, Dh = 256, K = 256, has_proj = 0, n_tiles = 2, tmem_alloc_cols = 512, n_ops_block = 4,
                         n_ops_out = 0, hc = 128, n_hchunk = 2, oc = 0, n_ochunk = 0, tmem_e_col = 0, tmem_h_col = 128,
                         tmem_tile_cols = 256, smem_tres = 65536, smem_ring = 98304, slot_bytes = 16384, n_stage = 7, h_split = 1,
                         n_ops_pre = 0, e_split = 0, epart = 128, smem_ae1 = 32768, block_w_bytes = 131072;
};
template <>
struct PlanView<1> {
    const QbStepPlan& q;
#define X(f) __device__ __forceinline__ constexpr int f() const { return PlanS128::f; }
    QB_PLAN_FIELDS(X)
#undef X
    __device__ __forceinline__ int L() const { return q.L; }
    __device__ __forceinline__ int skip() const { return q.skip; }
    __device__ __forceinline__ constexpr int smem_ae(int t) const { return t * PlanS128::smem_ae1; }
    __device__ __forceinline__ constexpr int64_t block_w_bytes() const { return PlanS128::block_w_bytes; }
};
bool plan_is_s128(const QbStepPlan& q) {
#define X(f) if (q.f != PlanS128::f) return false;
    QB_PLAN_FIELDS(X)
#undef X
    return q.smem_ae[0] == 0 && q.smem_ae[1] == PlanS128::smem_ae1 && q.block_w_bytes == PlanS128::block_w_bytes && q.pair == 0 && q.mcast == 0;
}
// The L384 view: the QINCo2-L family (de = dh = 384, K = 256: BASELINE configs 3-5, IVF-QINCo2-L) -- one tile per CTA, three H
// chunks, E epilogue in two parts.  D and what follows from it (out_proj chunking) stay run-time: d = 128 / 96 / 768; so do the
// pre-ops of the decode-loop plan, which is otherwise the same plan.
#define QB_PLAN_FIELDS_L384(X) X(De) X(Dh) X(K) X(has_proj) X(n_tiles) X(tmem_alloc_cols) X(n_ops_block) X(hc) X(n_hchunk) X(tmem_e_col) \
    X(tmem_h_col) X(tmem_tile_cols) X(smem_tres) X(smem_ring) X(slot_bytes) X(n_stage) X(h_split) X(e_split) X(epart)
struct PlanL384 {
    static constexpr int De = 384, Dh = 384, K = 256, has_proj = 1, n_tiles = 1, tmem_alloc_cols = 512, n_ops_block = 9, hc = 128,
                         n_hchunk = 3, tmem_e_col = 0, tmem_h_col = 384, tmem_tile_cols = 512, smem_tres = -1, smem_ring = 98304,
                         slot_bytes = 32768, n_stage = 3, h_split = 1, e_split = 1, epart = 192, smem_ae1 = 98304,
                         block_w_bytes = 589824;
};
template <>
struct PlanView<2> {
    const QbStepPlan& q;
#define X(f) __device__ __forceinline__ constexpr int f() const { return PlanL384::f; }
    QB_PLAN_FIELDS_L384(X)
#undef X
#define X(f) __device__ __forceinline__ int f() const { return q.f; }
    X(D) X(oc) X(n_ochunk) X(n_ops_out) X(n_ops_pre) X(L) X(skip)
#undef X
    __device__ __forceinline__ constexpr int smem_ae(int t) const { return t * PlanL384::smem_ae1; }
    __device__ __forceinline__ constexpr int64_t block_w_bytes() const { return PlanL384::block_w_bytes; }
};
bool plan_is_l384(const QbStepPlan& q) {
#define X(f) if (q.f != PlanL384::f) return false;
    QB_PLAN_FIELDS_L384(X)
#undef X
    return q.smem_ae[0] == 0 && q.smem_ae[1] == PlanL384::smem_ae1 && q.block_w_bytes == PlanL384::block_w_bytes && q.pair == 0;
}

template <bool kScore, bool kResident, bool kPair, bool kLoop = false, int kFuse = 0, bool kMcast = false, int kShape = 0>
__global__ void __launch_bounds__(kThreads, 1) qb_mlp_kernel(const __grid_constant__ MlpParams p) {
    static_assert(!kMcast || (!kPair && !kResident && (kFuse == 0 || kFuse == 3)), "weight multicast: non-resident single-CTA-MMA variants");
    static_assert(!kLoop || (!kScore && !kResident && !kPair), "the decode loop is an apply-mode, single-CTA variant");
    static_assert(!kFuse || (kScore && !kPair && !kLoop), "fused selection lives in the single-CTA score variants");
    constexpr bool kFuseB = kFuse == 3;              // in-CTA top-F_out + xhat' / history for one-tile-per-CTA shapes with out_proj
    static_assert(!kFuseB || !kResident, "fused selection B is the non-resident variant");
    // fused selection B: the winners' xhat' rows by stash slot, the running (sorted) candidate list of the vector(s) whose
    // tiles this CTA is walking, the tile-local winners and the slot every tile row was granted (0xff: none)
    __shared__ __align__(16) float selb_stash[kFuseB ? 4096 : 4];
    __shared__ float selb_run_d[kFuseB ? 2 * 128 : 1];
    __shared__ uint16_t selb_run_flat[kFuseB ? 2 * 128 : 1];
    __shared__ uint8_t selb_run_slot[kFuseB ? 2 * 128 : 1], selb_lrow[kFuseB ? 128 : 1], selb_take[kFuseB ? 128 : 1], selb_nrun[kFuseB ? 128 : 1];
    constexpr bool kFuseA = kFuse != 0 && kFuse != 3 && kResident; // cross-CTA arg-min of a beam-1 vector over its four code-quarter CTAs
    // fused selection, resident launches: the local winner's o of [buffer = set parity][tile slot][beam of the tile], its
    // packed (dist, code) key per lane quarter, and the epilogue <-> select-warp hand-off barriers per tile slot
    __shared__ __align__(16) float sel_stash[kFuseA ? (kFuse == 2 ? 2 * 2 * 2 * 128 : 2 * 2 * 256) : 4];
    __shared__ __align__(8) unsigned long long sel_key[kFuseA && kFuse == 2 ? 2 * 2 * 4 : 1];
    __shared__ __align__(8) uint64_t sel_full[2], sel_empty[2];
    extern __shared__ __align__(1024) uint8_t dyn_smem[];
    __shared__ __align__(8) uint64_t bars[2][QB_BAR_COUNT];     // one barrier set per tile slot
    __shared__ __align__(8) uint64_t w_full[QB_MAX_STAGE];
    __shared__ __align__(8) uint64_t w_empty[QB_MAX_STAGE];
    __shared__ __align__(8) uint64_t tres_bar;                  // resident table half has landed
    __shared__ __align__(8) uint64_t rows_full[3], rows_empty[3];   // resident mode: per-beam rows (u_b, r_b) of a set, 3 sets deep
    __shared__ __align__(16) float beam_rows[3][2][2][256];     // [buffer][tile slot][beam of the tile][u_b (De) | r_b (D)]
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(16) float dist_part[2][kColGroups > 1 ? kColGroups - 1 : 1][QB_TILE_M];

    const int tid = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform for the compiler (uniform datapath)
    static_assert(kShape == 0 || (kShape == 1 && kResident && !kPair && !kLoop && !kMcast) || (kShape == 2 && !kResident && !kPair),
                  "fixed-shape views: S128 for the resident score kernels, L384 for the one-tile-per-CTA kernels");
    const PlanView<kShape> pl{p.plan};
    const int n_ops = pl.n_ops_block() + pl.n_ops_out();
    const int NT = pl.n_tiles();
    // One tile set walks `n_ls` steps (1 outside the decode loop); a step is the phases  -1: pre-ops (decode loop only:
    // u = Wx . xhat), 0 .. L-1: residual blocks (one shared op list), L: out_proj ops.
    const int n_ls = kLoop ? p.n_loop_steps : 1;
    constexpr int kFirstPhase = kLoop ? -1 : 0;
    auto phase_ops = [&](int ph, int& i0, int& i1, size_t& w_rel) {
        if (ph < 0) { i0 = n_ops; i1 = n_ops + pl.n_ops_pre(); w_rel = 0; }
        else if (ph < pl.L()) { i0 = 0; i1 = pl.n_ops_block(); w_rel = (size_t)ph * (size_t)pl.block_w_bytes(); }
        else { i0 = pl.n_ops_block(); i1 = n_ops; w_rel = 0; }
    };
    const int64_t n_tiles = (p.n_rows + QB_TILE_M - 1) / QB_TILE_M;
    // Work units ("sets" of NT tiles).  Default: NT consecutive tiles, sets strided over the CTAs.  Resident mode (score,
    // all 256 codes per beam): CTA index % 4 picks a code quarter hq; a tile is that quarter (64 codes) of TWO consecutive
    // beams and a set is 2 such tiles = 4 consecutive beams, so every thread sees the same code for the whole launch and
    // the quarter's table rows never leave the SM.
    const int hq = kResident ? (int)(blockIdx.x & 3) : 0;
    const int64_t n_beams = p.n_rows >> 8;
    const int64_t n_sets = kResident ? (n_beams + 3) / 4 : (n_tiles + NT - 1) / NT;
    const int64_t set_first = kResident ? (int64_t)(blockIdx.x >> 2) : (int64_t)blockIdx.x;
    const int64_t set_stride = kResident ? (int64_t)(gridDim.x >> 2) : (int64_t)gridDim.x;
    // CTA pair: both CTAs walk the same number of sets (they share every MMA).  Resident mode gives both the same set
    // indices (adjacent code quarters); otherwise the peer's set is the leader's + 1 and may be past the end (all rows
    // invalid: it computes on row 0's operands and writes nothing).
    const uint32_t cta_rank = (kPair || kMcast) ? cluster_ctarank() : 0u;
    const bool leader = cta_rank == 0;
    const int64_t set_lead_off = ((kPair || kMcast) && !kResident) ? (int64_t)cta_rank : 0;
    auto more_sets = [&](int64_t set) { return set - set_lead_off < n_sets; };
    // Fused selection B walks the tiles of a vector back to back in ONE CTA (sel_spv tiles per vector, vectors strided over
    // the CTAs) and gives every CTA the same number of iterations (sets past the end have no valid rows).
    const int64_t fb_spv = kFuseB ? (int64_t)p.sel_spv : 1;
    const int64_t fb_groups = kFuseB ? (n_tiles / fb_spv + (n_tiles % fb_spv ? 1 : 0)) : 0;
    const int64_t fb_nit = kFuseB ? ((fb_groups + gridDim.x - 1) / gridDim.x) * fb_spv : 0;
    auto fb_set_at = [&](int64_t it) { return ((it / fb_spv) * (int64_t)gridDim.x + blockIdx.x) * fb_spv + it % fb_spv; };
#define QB_FOR_SETS(IT, SET)                                                                        \
    for (int64_t IT = 0, SET = kFuseB ? fb_set_at(0) : set_first; kFuseB ? (IT < fb_nit) : more_sets(SET); \
         IT++, SET = kFuseB ? fb_set_at(IT) : SET + set_stride)

    // ---- one-time setup ------------------------------------------------------------------------------------------
    if (tid == 0) {
        for (int t = 0; t < 2; t++) {
            mbar_init(smem_u32(&bars[t][QB_BAR_NONE]), 1);
            // epilogue -> MMA issuer: one arrival per epilogue thread (QB_ARRIVE_ALL) or warp, from both CTAs of a pair
            // (the issuer lives in the leader CTA)
            mbar_init(smem_u32(&bars[t][QB_BAR_AE_READY]), (kPair ? 2 : 1) * kArriveCount);
            mbar_init(smem_u32(&bars[t][QB_BAR_AH_READY]), (kPair ? 2 : 1) * kArriveCount);
            mbar_init(smem_u32(&bars[t][QB_BAR_HACC_FREE]), (kPair ? 2 : 1) * kArriveCount);
            mbar_init(smem_u32(&bars[t][QB_BAR_AH2_READY]), (kPair ? 2 : 1) * kArriveCount);
            mbar_init(smem_u32(&bars[t][QB_BAR_HACC_FULL]), 1);
            mbar_init(smem_u32(&bars[t][QB_BAR_EACC_FULL]), 1);
            mbar_init(smem_u32(&bars[t][QB_BAR_EACC_HALF]), 1);
        }
        for (int t = 0; t < 2; t++) { mbar_init(smem_u32(&sel_full[t]), kEpiWarps); mbar_init(smem_u32(&sel_empty[t]), 1); }
        mbar_init(smem_u32(&tres_bar), 1);
        for (int b = 0; b < 3; b++) { mbar_init(smem_u32(&rows_full[b]), 1); mbar_init(smem_u32(&rows_empty[b]), kEpiThreads); }
        for (int s = 0; s < QB_MAX_STAGE; s++) {
            mbar_init(smem_u32(&w_full[s]), (kPair && leader) ? 2 : 1);   // leader: own half landed + the peer's relay
            mbar_init(smem_u32(&w_empty[s]), (uint32_t)pl.n_tiles() * (kMcast ? 2u : 1u));   // released by every tile slot's MMA warp (of both CTAs with multicast)
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kMmaWarp) {
        if (kPair) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                         "r"((uint32_t)pl.tmem_alloc_cols())
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                         "r"((uint32_t)pl.tmem_alloc_cols())
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (kPair || kMcast) cluster_sync_all();      // the peer's barriers are initialised before anything arrives on them remotely
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_base_s, 0);
    // Shared-window addresses, computed ONCE and made opaque to the compiler.  Left alone, it re-derives every
    // __cvta_generic_to_shared() at its use from S2UR SR_CgaCtaId, a special-register read that costs ~300 cycles
    // (device trace) -- per barrier operation, in every role's inner loop.
    auto opaque = [](uint32_t v) { uint32_t r; asm volatile("mov.b32 %0, %1;" : "=r"(r) : "r"(v)); return r; };
    const uint32_t smem_base = opaque(smem_u32(dyn_smem));
    const uint32_t a_bars = opaque(smem_u32(&bars[0][0])), a_wfull = opaque(smem_u32(&w_full[0])), a_wempty = opaque(smem_u32(&w_empty[0]));
    const uint32_t a_tres = opaque(smem_u32(&tres_bar)), a_rfull = opaque(smem_u32(&rows_full[0])), a_rempty = opaque(smem_u32(&rows_empty[0]));
    const uint32_t a_beam = opaque(smem_u32(&beam_rows[0][0][0][0]));
    const uint32_t a_dist = opaque(smem_u32(&dist_part[0][0][0]));
    const uint32_t a_sfull = opaque(smem_u32(&sel_full[0])), a_sempty = opaque(smem_u32(&sel_empty[0]));
    const uint32_t a_stash = opaque(smem_u32(&sel_stash[0])), a_skey = opaque(smem_u32(&sel_key[0]));
    auto bar_addr = [&](int t, int b) { return a_bars + (uint32_t)(t * QB_BAR_COUNT + b) * 8u; };
    const uint32_t tile_cols = (uint32_t)pl.tmem_tile_cols();

    // (unconditional on purpose: ptxas only budgets registers per region when every path executes the setmaxnreg)
    if (warp >= kEpiWarps) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kSvcRegs));
    else asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kEpiRegs));
    if (warp == kProducerWarp) {
        // ======================================================================================= weight producer
        uint32_t stage = 0, phase = 0;
        if (kResident) {    // this CTA's quarter of T_m and C_m: per 4-column block 64 codes x 16 B = 1 KB, contiguous in the tables
            if (elect_one()) {
                mbar_expect_tx(a_tres, (uint32_t)pl.D() * 256u);
                for (int c4 = 0; c4 < (pl.D() >> 2); c4++)
                    bulk_g2s(smem_base + pl.smem_tres() + c4 * 1024, p.cb_blk + ((size_t)c4 * pl.K() + hq * 64) * 4, 1024, a_tres);
            }
            __syncwarp();
        }
        // resident mode: the per-beam rows u_b (init) and r_b (distance) of set number k land in beam_rows[k % 3] one set
        // ahead of their use; buffer k % 3 was last read by set k - 3, long finished
        auto push_rows = [&](int64_t set, int64_t k) {
            if (set >= n_sets) return;
            const int b = (int)(k % 3);
            mbar_wait((a_rempty + (uint32_t)(b) * 8u), (uint32_t)(((k / 3) & 1) ^ 1), p.err_flag, 0x600 + b);
            if (elect_one()) {
                const uint32_t de_b = (uint32_t)pl.De() * 4u, d_b = (uint32_t)pl.D() * 4u;
                mbar_expect_tx((a_rfull + (uint32_t)(b) * 8u), 4u * (de_b + d_b));
                for (int t = 0; t < 4; t++) {       // tile slot t / 2, beam t % 2 of the tile
                    int64_t beam = 4 * set + t;
                    if (beam >= n_beams) beam = 0;
                    bulk_g2s((a_beam + (uint32_t)(((b * 2 + (t >> 1)) * 2 + (t & 1)) * 256) * 4u), p.u + beam * pl.De(), de_b, (a_rfull + (uint32_t)(b) * 8u));
                    bulk_g2s((a_beam + (uint32_t)(((b * 2 + (t >> 1)) * 2 + (t & 1)) * 256 + pl.De()) * 4u), p.r + beam * pl.D(), d_b, (a_rfull + (uint32_t)(b) * 8u));
                }
            }
            __syncwarp();
        };
        int64_t kset = 0;
        if (kResident) push_rows(set_first, 0);
        QB_FOR_SETS(it_p, set) {
            if (kResident) push_rows(set + set_stride, kset + 1);
            for (int ls = 0; ls < n_ls; ls++)
            for (int l = kFirstPhase; l <= pl.L(); l++) {
                int i0, i1;
                size_t w_rel;
                phase_ops(l, i0, i1, w_rel);
                const uint8_t* wbase = (kLoop ? p.loop_steps[ls].w_blob : p.w_blob) + w_rel;
                for (int i = i0; i < i1; i++) {
                    const QbOp& op = p.ops[i];
                    const uint32_t n_slab = op.n_slab, slab_bytes = op.slab_bytes, last_bytes = op.last_bytes;
                    const uint8_t* src = wbase + op.w_off;
                    for (uint32_t s = 0; s < n_slab; s++) {
                        // CTA pair: each CTA streams only its half of the slab's rows (packed half after half)
                        const uint32_t full = (s + 1 == n_slab) ? last_bytes : slab_bytes;
                        const uint32_t bytes = (kPair || kMcast) ? full >> 1 : full;
                        mbar_wait((a_wempty + stage * 8u), phase ^ 1, p.err_flag, 0x100 + stage);
                        if (elect_one()) {
                            if (kMcast) {       // the whole slab lands here: this CTA's half and the peer's, both multicast
                                mbar_expect_tx((a_wfull + stage * 8u), full);
                                bulk_g2s_mcast(smem_base + pl.smem_ring() + stage * pl.slot_bytes() + cta_rank * bytes,
                                               src + (size_t)s * slab_bytes + cta_rank * bytes, bytes, (a_wfull + stage * 8u), (uint16_t)3);
                            } else {
                            mbar_expect_tx((a_wfull + stage * 8u), bytes);
                            bulk_g2s(smem_base + pl.smem_ring() + stage * pl.slot_bytes(),
                                     src + (size_t)s * slab_bytes + (kPair ? cta_rank * bytes : 0u), bytes, (a_wfull + stage * 8u));
                            }
                        }
                        __syncwarp();
                        if (++stage == (uint32_t)pl.n_stage()) { stage = 0; phase ^= 1; }
                    }
                }
            }
            kset++;
        }
    } else if (warp >= kMmaWarp) {
        // ======================================================================================= MMA issuers
        // One warp per tile slot walks the op list for its own tile; both wait on the same "slab landed" barriers and a
        // ring slot is recycled once every slot's warp has committed it (w_empty counts n_tiles arrivals), so the
        // weights are fetched once per two tiles while the two issue streams overlap each other's bookkeeping.
        // Kept lean and warp-uniform on purpose: a lone warp retires a dependent instruction every ~4-6 cycles.
        const int t = warp - kMmaWarp;
        if (kPair && !leader) {
            // Peer CTA of a pair: the leader issues every MMA.  One warp here relays "my half of the slab has landed" to
            // the leader's w_full barrier (a bulk copy can only signal a barrier of the CTA it writes to).
            if (t == 0) {
                uint32_t stage = 0, phase = 0;
                for (int64_t set = set_first; more_sets(set); set += set_stride) {
                    for (int l = 0; l <= pl.L(); l++) {
                        const int i0 = (l < pl.L()) ? 0 : pl.n_ops_block();
                        const int i1 = (l < pl.L()) ? pl.n_ops_block() : n_ops;
                        for (int i = i0; i < i1; i++) {
                            const uint32_t n_slab = p.ops[i].n_slab;
                            for (uint32_t s = 0; s < n_slab; s++) {
                                mbar_wait((a_wfull + stage * 8u), phase, p.err_flag, 0x700 + stage);
                                if (elect_one()) mbar_arrive_cluster(mapa_rank((a_wfull + stage * 8u), 0));
                                __syncwarp();
                                if (++stage == (uint32_t)pl.n_stage()) { stage = 0; phase ^= 1; }
                            }
                        }
                    }
                }
            }
        } else if (kFuseA && t == 2) {
            // ================================================================================= select warp (fused, resident)
            // Per (set, tile slot): publish this CTA's local winners (packed atomicMin + arrival count per vector), then
            // settle the PREVIOUS set: by then the other three code quarters have normally reported too, so the wait is a
            // formality; whoever holds the global winner adds xhat_b to its stashed o and writes xhat' and the history.
            const int lane = tid & 31;
            const int D = pl.D();
            auto ld_acquire_u32 = [](const uint32_t* ptr) { uint32_t v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ptr) : "memory"); return v; };
            auto ld_relaxed_u64 = [](const unsigned long long* ptr) { unsigned long long v; asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(ptr) : "memory"); return v; };
            auto settle = [&](int64_t set, int t, int buf) {
                for (int h = 0; h < 2; h++) {
                    const int64_t v = 4 * set + 2 * t + h;
                    if (v >= n_beams) continue;
                    uint32_t spins = 0;
                    uint64_t t0 = 0;
                    while (ld_acquire_u32(p.sel_cnt + v) < 4u) {          // all four code quarters have reported
                        __nanosleep(200);
                        if ((++spins & 255u) == 0) {
                            const uint64_t now = globaltimer();
                            if (t0 == 0) t0 = now;
                            if (now - t0 > 4000000000ull) { atomicExch(p.err_flag, 0x800u); __threadfence_system(); __trap(); }
                        }
                    }
                    const unsigned long long key = ld_relaxed_u64(p.sel_best + v);
                    const int code = (int)(key & 0xffffffffull);
                    if ((code >> 6) != hq) continue;                      // the winner sits in another quarter's CTA
                    uint8_t* ho = p.hist_out + v * p.hist_M;
                    const uint8_t* hi = p.hist_in + v * p.hist_M;
                    for (int c = lane; c < p.hist_m; c += 32) ho[c] = hi[c];
                    if (lane == 0) ho[p.hist_m] = (uint8_t)code;
                    if (p.xhat_out) {
                        const uint32_t st = a_stash + (uint32_t)(((buf * 2 + t) * 2 + h) * 128) * 4u;
                        for (int d4 = lane; d4 * 4 < D; d4 += 32) {
                            const float4 xi = *reinterpret_cast<const float4*>(p.xhat_in + v * D + d4 * 4);
                            const float4 o = lds4(st + (uint32_t)d4 * 16u);
                            *reinterpret_cast<float4*>(p.xhat_out + v * D + d4 * 4) = make_float4(xi.x + o.x, xi.y + o.y, xi.z + o.z, xi.w + o.w);
                        }
                    }
                }
            };
            int64_t kset = 0, prev_set = -1;
            constexpr bool lite = kFuse == 1;
            for (int64_t set = set_first; more_sets(set); set += set_stride, kset++) {
                const int buf = (int)(kset & 1);
                for (int tt = 0; tt < 2; tt++) {
                    mbar_wait_relaxed(a_sfull + (uint32_t)tt * 8u, (uint32_t)(kset & 1), p.err_flag, 0x810 + tt);
                    if constexpr (lite) {
                        // fuse == 1: the epilogue left the 128 distances of the tile in shared memory (rows 0-63 / 64-127 =
                        // the code quarter of two consecutive vectors); this warp ranks them and reports the two local
                        // minima.  The winner's xhat' is produced by the small update launch that follows the score launch.
                        const uint32_t da = a_stash + (uint32_t)((buf * 2 + tt) * 256) * 4u;
#pragma unroll
                        for (int h = 0; h < 2; h++) {
                            const float d0 = lds1(da + (uint32_t)(64 * h + lane) * 4u), d1 = lds1(da + (uint32_t)(64 * h + 32 + lane) * 4u);
                            const unsigned long long ka = (((unsigned long long)__float_as_uint(d0)) << 32) | (unsigned)(hq * 64 + lane);
                            const unsigned long long kb = (((unsigned long long)__float_as_uint(d1)) << 32) | (unsigned)(hq * 64 + 32 + lane);
                            unsigned long long key = ka < kb ? ka : kb;
#pragma unroll
                            for (int off = 16; off > 0; off >>= 1) {
                                const unsigned long long o = __shfl_xor_sync(0xffffffffu, key, off);
                                key = o < key ? o : key;
                            }
                            const int64_t v = 4 * set + 2 * tt + h;
                            if (lane == 0 && v < n_beams) atomicMin(p.sel_best + v, key);
                        }
                        __syncwarp();
                        if (lane == 0) mbar_arrive(a_sempty + (uint32_t)tt * 8u);
                    } else {
                    if (lane < 2) {
                        const int64_t v = 4 * set + 2 * tt + lane;
                        if (v < n_beams) {
                            unsigned long long k0, k1;
                            const uint32_t ka = a_skey + (uint32_t)(((buf * 2 + tt) * 4 + 2 * lane) * 8);
                            asm volatile("ld.shared.u64 %0, [%1];" : "=l"(k0) : "r"(ka));
                            asm volatile("ld.shared.u64 %0, [%1];" : "=l"(k1) : "r"(ka + 8u));
                            atomicMin(p.sel_best + v, k0 < k1 ? k0 : k1);
                            __threadfence();
                            atomicAdd(p.sel_cnt + v, 1u);
                        }
                    }
                    __syncwarp();
                    if (prev_set >= 0) settle(prev_set, tt, buf ^ 1);
                    // one completion per set on sel_empty: "this set's keys are taken and the set before it is settled" --
                    // the epilogue waits for it one set later, which bounds its lead over this warp to one set (the
                    // barriers carry a single phase bit) and frees the stash buffer it is about to reuse
                    __syncwarp();
                    if (lane == 0) mbar_arrive(a_sempty + (uint32_t)tt * 8u);
                    }   // full
                }
                prev_set = set;
            }
            if (prev_set >= 0 && kFuse == 2) {
                settle(prev_set, 0, (int)((kset - 1) & 1));
                settle(prev_set, 1, (int)((kset - 1) & 1));
            }
        } else if (t < NT) {
            uint32_t stage = 0, phase = 0;
            uint32_t par = 0;
            Tracer tr;
            tr.init(p, 1 + t, (tid & 31) == 0);      // trace roles 1 / 2: MMA warps of tile slots 0 / 1
            const uint32_t ring_lo = ((smem_base + pl.smem_ring()) >> 4) & 0x3FFFu;    // descriptor address fields (14 bits, 16-B units)
            const uint32_t slot_lo = (uint32_t)pl.slot_bytes() >> 4;
            const uint32_t tcol = tmem_base + (uint32_t)t * tile_cols;
            const uint32_t ae_lo = ((smem_base + pl.smem_ae(t)) >> 4) & 0x3FFFu;
            QB_FOR_SETS(it_m, set) {
                for (int ls = 0; ls < n_ls; ls++)
                for (int l = kFirstPhase; l <= pl.L(); l++) {
                    int i0, i1;
                    size_t w_rel_unused;
                    phase_ops(l, i0, i1, w_rel_unused);
                    for (int i = i0; i < i1; i++) {
                        const QbOp& op = p.ops[i];          // kernel-parameter bank -> uniform registers
                        const uint32_t n = op.n;
                        // pair: this CTA's slab half has n / 2 rows per k-chunk; the MMA is M = 256 over both CTAs
                        const uint32_t b_lbo = (kPair ? n >> 1 : n) << 16;         // LBO field (bytes / 16) of the low descriptor word
                        const uint32_t b_step = kPair ? n : 2u * n;                // descriptor address units (16 B) per K=16
                        const uint32_t idesc = umma_idesc(n, kPair ? 2 * QB_TILE_M : QB_TILE_M);
                        const bool from_smem = op.a_src == QB_A_E;
                        const int ks16 = op.ks >> 4;
                        const uint32_t n_slab = op.n_slab;
                        const uint32_t d_tmem = tcol + op.d_col;
                        // A operand cursor: low descriptor word (shared memory; LBO = one k-chunk of the tile) or TMEM column
                        uint32_t a_cur = from_smem ? (((uint32_t)kAkcBytes >> 4) << 16) | (ae_lo + (uint32_t)op.a_off * (kAkcBytes >> 4))
                                                   : tcol + op.a_off;
                        uint32_t acc = op.accumulate;
                        const bool blk32 = op.a_blk32 != 0;
                        uint32_t j0 = 0;                    // k-steps issued so far (blocked TMEM operand layout)
                        int k16_left = op.k_total >> 4;
                        const uint32_t commit_bar = op.commit;
                        const uint32_t wa = op.wait_a, wd = op.wait_d, wa2_slab = op.wait_a2_slab;
                        for (uint32_t s = 0; s < n_slab; s++) {
                            const int nk = k16_left < ks16 ? k16_left : ks16;
                            k16_left -= ks16;
                            // (no tcgen05.fence after this wait: the slab was written and is read by the async proxy)
                            mbar_wait((a_wfull + stage * 8u), phase, p.err_flag, 0x300 + stage);
                            const uint32_t b_lo = b_lbo | (ring_lo + stage * slot_lo);
                            if (s == 0) {
                                // Operand barriers LAST: everything above (op decode, the weight wait) is off the
                                // epilogue -> MMA hand-off path; an issuing warp retires only ~1 instruction per 6 cycles.
                                if (wa) {
                                    mbar_wait(bar_addr(t, wa), (par >> wa) & 1, p.err_flag, 0x200 + wa);
                                    par ^= 1u << wa;
                                }
                                if (wd) {
                                    mbar_wait(bar_addr(t, wd), (par >> wd) & 1, p.err_flag, 0x200 + wd);
                                    par ^= 1u << wd;
                                }
                                if (wa | wd) tc_fence_after();   // orders the epilogue's tcgen05.st / ld before these MMAs
                                tr.ev(0x200 + i);
                            } else if (s == wa2_slab) {     // (wa2_slab != 0) second K half of a split H chunk
                                mbar_wait(bar_addr(t, QB_BAR_AH2_READY), (par >> QB_BAR_AH2_READY) & 1, p.err_flag, 0x200 + QB_BAR_AH2_READY);
                                par ^= 1u << QB_BAR_AH2_READY;
                                tc_fence_after();
                            }
                            if (elect_one()) {
                                uint32_t a_l = a_cur, b_l = b_lo;      // cursors local to the issuing lane (the warp-wide ones advance below)
                                if (from_smem) {
                                    mma_slab<kPair, true>(nk, d_tmem, a_l, b_l, kDescHi, kDescHi, idesc, acc, (2 * kAkcBytes) >> 4, (2 * kAkcBytes) >> 4, b_step);
                                } else if (blk32) {     // k-step j of the op at column 32 * (j / 2) + 8 * (j % 2)
                                    a_l = tcol + op.a_off + 32u * (j0 >> 1) + 8u * (j0 & 1u);
                                    mma_slab<kPair, false>(nk, d_tmem, a_l, b_l, kDescHi, kDescHi, idesc, acc, (j0 & 1u) ? 24u : 8u, (j0 & 1u) ? 8u : 24u, b_step);
                                } else {
                                    mma_slab<kPair, false>(nk, d_tmem, a_l, b_l, kDescHi, kDescHi, idesc, acc, 8u, 8u, b_step);
                                }
                                if (kPair) {        // both CTAs recycle the ring slot / see the accumulator
                                    tc_commit_pair((a_wempty + stage * 8u));
                                    if (s + 1 == n_slab && commit_bar) tc_commit_pair(bar_addr(t, commit_bar));
                                } else {
                                    if (kMcast) tc_commit_mcast((a_wempty + stage * 8u)); else tc_commit((a_wempty + stage * 8u));
                                    if (s + 1 == n_slab && commit_bar) tc_commit(bar_addr(t, commit_bar));
                                }
                            }
                            __syncwarp();
                            a_cur += (uint32_t)nk * (from_smem ? ((2 * kAkcBytes) >> 4) : 8u);
                            j0 += (uint32_t)nk;
                            acc = 1;
                            if (++stage == (uint32_t)pl.n_stage()) { stage = 0; phase ^= 1; }
                        }
                        tr.ev(0x400 + i);
                    }
                }
            }
        }
    } else if (warp < kEpiWarps) {
        // ======================================================================================= epilogue warps
        // Thread (lane quarter q = warp % 4, lane) owns row r = 32 q + lane of BOTH tiles in flight and alternates
        // between them; the kColGroups warps of a lane quarter split every column range (cg = warp / 4).  Table rows
        // (T_m, the skip codeword) are fetched 64 columns at a time with all loads of a batch issued together, the skip
        // codeword before the thread starts waiting for the accumulator.
        const int q = warp & 3, cg = warp >> 2;
        const int r = q * 32 + (tid & 31);
        const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
        uint32_t par0 = 0, par1 = 0;
        Tracer tr;
        tr.init(p, 0, tid == 0);
        // "operand ready / accumulator free" signal to the MMA issuer of tile slot t: every thread fences its own TMEM /
        // shared-memory writes, the warp converges and one lane arrives (barrier counts are per warp).  In a CTA pair
        // the issuer sits in the leader CTA, so the arrival goes to the leader's barrier through the cluster window.
        auto arrive_issuer = [&](int t, int bar, bool wrote_smem) {
            tc_fence_before();
            if (wrote_smem) { if (kPair) proxy_fence_async_all(); else proxy_fence_async(); }
            if (!QB_ARRIVE_ALL) __syncwarp();
            if (QB_ARRIVE_ALL || (tid & 31) == 0) {
                if (kPair) mbar_arrive_cluster(mapa_rank(bar_addr(t, bar), 0));
                else mbar_arrive(bar_addr(t, bar));
            }
        };
        const int De = pl.De(), K = pl.K(), D = pl.D();
        int e0c, e1c, o0c, o1c;             // this thread's columns of e / of the output (no out_proj)
        group_range(De, cg, e0c, e1c);
        group_range(D, cg, o0c, o1c);
        if (kResident) mbar_wait(a_tres, 0, p.err_flag, 0x500);
        // Resident mode: this thread's code never changes, so its slice of the T_m row (<= 64 columns) lives in registers
        // for the whole launch; reading it from shared memory for every tile costs as much shared-memory bandwidth as a
        // quarter of the tile's MMAs.
        constexpr int kTReg = 32 / kColGroups;      // float4 registers: 128 / kColGroups columns
        float4 treg[kTReg];
        if (kResident) {
            const float* tsrc = p.t_blk + ((size_t)(e0c >> 2) * K + (hq * 64 + (r & 63))) * 4;
#pragma unroll
            for (int i = 0; i < kTReg; i++)
                treg[i] = (e0c + 4 * i < e1c) ? ldg4(tsrc + (size_t)i * K * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if constexpr (kLoop) {
            // ============================================================================ decode loop (one launch, all steps)
            // A tile of 128 vectors walks every quantisation step here; its running reconstruction xhat never leaves the
            // rows' owner threads: each thread keeps its column range of xhat in xhat_out (re-read by the same thread one
            // step later, an L1/L2 hit) and hands the NEXT step's MMA operand [xhat_hi | xhat_lo] (fp16 hi/lo split) to the
            // tensor core through the A_E buffer, where the pre-ops add u = Wx . xhat onto Eacc = T_m[code].
            uint32_t parl0 = 0, parl1 = 0;
            auto wait_l = [&](int t, int bar, uint32_t code) {
                uint32_t& par = t ? parl1 : parl0;
                mbar_wait(bar_addr(t, bar), (par >> bar) & 1, p.err_flag, code);
                par ^= 1u << bar;
                tc_fence_after();
            };
            const uint32_t x_lo_chunk = (uint32_t)(D >> 3);        // k-chunk of the first xhat_lo column
            // n (16 or 32) columns of xhat starting at column c: fp16 hi / lo parts -> A operand k-chunks of tile slot t
            auto put_operand = [&](int t, int c, int n, const float (&x)[32]) {
                const uint32_t dst = smem_base + pl.smem_ae(t) + (uint32_t)r * 16u;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    if (8 * j < n) {
                        uint32_t hi[4], lo[4];
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            const float a = x[8 * j + 2 * k], b = x[8 * j + 2 * k + 1];
                            const __half2 h = __floats2half2_rn(a, b);
                            const float2 hf = __half22float2(h);
                            hi[k] = *reinterpret_cast<const uint32_t*>(&h);
                            lo[k] = pack_h2(a - hf.x, b - hf.y);
                        }
                        st_shared_v4(dst + ((uint32_t)(c >> 3) + j) * kAkcBytes, hi[0], hi[1], hi[2], hi[3]);
                        st_shared_v4(dst + (x_lo_chunk + (uint32_t)(c >> 3) + j) * kAkcBytes, lo[0], lo[1], lo[2], lo[3]);
                    }
                }
            };
            // seed: xhat = seed_tab[code0] (C_0[code] or the IVF centroid) over this thread's columns
            auto seed_tile = [&](int t, int64_t row, bool valid) {
                int c0 = 0;
                if (valid) {
                    if (p.seed_codes_i32) {
                        c0 = __ldg(p.seed_codes_i32 + row);
                        if (c0 < 0 || c0 >= p.seed_K) { atomicExch(p.err_flag, 0x20u); c0 = 0; }
                    } else {
                        c0 = (int)__ldg(p.sel_code + row * p.code_stride);
                        if (c0 >= p.seed_K) { atomicExch(p.err_flag, 0x10u); c0 = p.seed_K - 1; }
                    }
                }
                const float* src = p.seed_tab + (size_t)c0 * D;
                float* xrow = p.xhat_out + row * D;
#pragma unroll 1
                for (int c = o0c; c < o1c; c += 32) {
                    const int n = (o1c - c) >= 32 ? 32 : 16;
                    float x[32];
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const float4 v = (4 * i < n) ? ldg4(src + c + 4 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
                        x[4 * i] = v.x; x[4 * i + 1] = v.y; x[4 * i + 2] = v.z; x[4 * i + 3] = v.w;
                        if (valid && 4 * i < n) *reinterpret_cast<float4*>(xrow + c + 4 * i) = v;
                    }
                    put_operand(t, c, n, x);
                }
            };
            auto step_code = [&](int64_t row, bool valid, int ls) {
                int code = 0;
                if (valid) {
                    code = (int)__ldg(p.sel_code + row * p.code_stride + p.code_off + ls);
                    if (code >= K) { atomicExch(p.err_flag, 0x10u); code = K - 1; }
                }
                return code;
            };
            // init of a step: Eacc <- T_m[code] (fp32); the operand [xhat_hi | xhat_lo] is already in shared memory
            auto init_loop = [&](int t, int code, const float* t_blk) {
                const uint32_t tl = lane_base + (uint32_t)t * tile_cols + pl.tmem_e_col();
#pragma unroll 1
                for (int c = e0c; c < e1c; c += 32) {
                    const int n = (e1c - c) >= 32 ? 32 : 16;
                    uint32_t e[32];
                    const float* base = t_blk + ((size_t)(c >> 2) * K + code) * 4;
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const float4 v = (4 * i < n) ? ldg4(base + (size_t)i * K * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                        e[4 * i] = __float_as_uint(v.x); e[4 * i + 1] = __float_as_uint(v.y);
                        e[4 * i + 2] = __float_as_uint(v.z); e[4 * i + 3] = __float_as_uint(v.w);
                    }
                    __syncwarp();
                    if (n >= 32) tmem_st32(tl + c, e); else tmem_st16p(tl + c, e);
                }
                tmem_wait_st();
                arrive_issuer(t, QB_BAR_AE_READY, true);
            };
            // final of a step: xhat' = xhat + o (+ C_m[code]) over this thread's columns; o sits in Eacc (no out_proj) or in
            // the single out_proj chunk in Hacc.  Last step: scale / shift and write the result; otherwise write the running
            // xhat and the next step's operand.
            auto final_loop = [&](int t, int64_t row, bool valid, int code, const float* cb_blk, bool last) {
                const uint32_t taddr = lane_base + (uint32_t)t * tile_cols + (pl.has_proj() ? pl.tmem_h_col() : pl.tmem_e_col());
                float* xrow = p.xhat_out + row * D;
                const bool skip = pl.skip() != 0;
#pragma unroll 1
                for (int c = o0c; c < o1c; c += 32) {
                    const int n = (o1c - c) >= 32 ? 32 : 16;
                    uint32_t v[32];
                    tmem_ld_cols(taddr + c, n, v);
                    float4 xin[8], cb[8];
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const bool on = 4 * i < n;
                        xin[i] = (valid && on) ? *reinterpret_cast<const float4*>(xrow + c + 4 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
                        cb[i] = (skip && on) ? ldg4(cb_blk + ((size_t)((c >> 2) + i) * K + code) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                    tmem_wait_ld();
                    float x[32];
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        x[4 * i] = xin[i].x + (__uint_as_float(v[4 * i]) + cb[i].x);
                        x[4 * i + 1] = xin[i].y + (__uint_as_float(v[4 * i + 1]) + cb[i].y);
                        x[4 * i + 2] = xin[i].z + (__uint_as_float(v[4 * i + 2]) + cb[i].z);
                        x[4 * i + 3] = xin[i].w + (__uint_as_float(v[4 * i + 3]) + cb[i].w);
                    }
                    if (!last) put_operand(t, c, n, x);
                    if (valid) {
#pragma unroll
                        for (int i = 0; i < 8; i++) {
                            if (4 * i < n) {
                                float4 o = make_float4(x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]);
                                if (last) {
                                    if (p.out_shift) {
                                        const float4 sh = ldg4(p.out_shift + c + 4 * i);
                                        o.x = fmaf(o.x, p.out_scale, sh.x); o.y = fmaf(o.y, p.out_scale, sh.y);
                                        o.z = fmaf(o.z, p.out_scale, sh.z); o.w = fmaf(o.w, p.out_scale, sh.w);
                                    } else if (p.out_scale != 1.0f) {
                                        o.x *= p.out_scale; o.y *= p.out_scale; o.z *= p.out_scale; o.w *= p.out_scale;
                                    }
                                }
                                *reinterpret_cast<float4*>(xrow + c + 4 * i) = o;
                            }
                        }
                    }
                }
            };
            for (int64_t set = set_first; more_sets(set); set += set_stride) {
                const int64_t row0 = (set * NT + 0) * QB_TILE_M + r, row1 = (set * NT + 1) * QB_TILE_M + r;
                const bool valid0 = row0 < p.n_rows, valid1 = NT > 1 && row1 < p.n_rows;
                seed_tile(0, valid0 ? row0 : 0, valid0);
                if (NT > 1) seed_tile(1, valid1 ? row1 : 0, valid1);
                int code0 = step_code(row0, valid0, 0), code1 = (NT > 1) ? step_code(row1, valid1, 0) : 0;
                init_loop(0, code0, p.loop_steps[0].t_blk);
                if (NT > 1) init_loop(1, code1, p.loop_steps[0].t_blk);
#pragma unroll 1
                for (int ls = 0; ls < n_ls; ls++) {
                    const bool last = ls + 1 == n_ls;
                    // e0 = T_m[code] + u is complete: fp16 copy for the first up-projection (or the out_proj)
#pragma unroll 1
                    for (int t = 0; t < NT; t++) {
                        wait_l(t, QB_BAR_EACC_FULL, 0x425);
                        if (pl.L() > 0 || pl.has_proj()) {      // (no GEMM follows in a block-less model without out_proj)
                            acc_to_smem_operand(lane_base + (uint32_t)t * tile_cols + pl.tmem_e_col(), e0c, e1c,
                                                smem_base + pl.smem_ae(t) + (uint32_t)r * 16u);
                            arrive_issuer(t, QB_BAR_AE_READY, true);
                        }
                    }
#pragma unroll 1
                    for (int l = 0; l < pl.L(); l++) {
#pragma unroll 1
                        for (int j = 0; j < pl.n_hchunk(); j++) {
                            const int cw = min(pl.hc(), pl.Dh() - j * pl.hc());
                            int c0, c1;
                            group_range(cw, cg, c0, c1);
#pragma unroll 1
                            for (int t = 0; t < NT; t++) {
                                wait_l(t, QB_BAR_HACC_FULL, 0x424);
                                const uint32_t th = lane_base + (uint32_t)t * tile_cols + pl.tmem_h_col();
                                if (kColGroups == 2 && pl.h_split() && cw == pl.hc() && (cw >> 2) <= 32) {
                                    if (pl.h_split() == 2) {
                                        const int qw = cw >> 2;
                                        acc_to_tmem_operand_half(th, cg * qw, qw, 1 + q, true);
                                        arrive_issuer(t, QB_BAR_AH_READY, false);
                                        acc_to_tmem_operand_half(th, (cw >> 1) + cg * qw, qw, 1 + q, true);
                                        arrive_issuer(t, QB_BAR_AH2_READY, false);
                                    } else {
                                        acc_to_tmem_operand_split(th, cw, cg, 1 + q, [&](int h) { arrive_issuer(t, h ? QB_BAR_AH2_READY : QB_BAR_AH_READY, false); });
                                    }
                                } else if (kColGroups == 4 && cw == 128) {     // one 32-column block per warp
                                    acc_to_tmem_operand_half(th, cg * 32, 32, 1 + q, pl.h_split() == 2);
                                    arrive_issuer(t, QB_BAR_AH_READY, false);
                                    if (pl.h_split() && cw == pl.hc()) arrive_issuer(t, QB_BAR_AH2_READY, false);
                                } else {
                                    acc_to_tmem_operand(th, c0, c1, 1 + q);
                                    arrive_issuer(t, QB_BAR_AH_READY, false);
                                    if (kColGroups != 2 && pl.h_split() && cw == pl.hc()) arrive_issuer(t, QB_BAR_AH2_READY, false);
                                }
                            }
                        }
                        if (l + 1 < pl.L() || pl.has_proj()) {
#pragma unroll 1
                            for (int t = 0; t < NT; t++) {
                                const uint32_t te = lane_base + (uint32_t)t * tile_cols + pl.tmem_e_col();
                                const uint32_t sd = smem_base + pl.smem_ae(t) + (uint32_t)r * 16u;
                                const int n_ep = pl.e_split() ? 2 : 1;
#pragma unroll 1
                                for (int ep = 0; ep < n_ep; ep++) {
                                    int h0 = e0c, h1 = e1c;
                                    if (pl.e_split()) {
                                        group_range(ep == 0 ? pl.epart() : De - pl.epart(), cg, h0, h1);
                                        if (ep) { h0 += pl.epart(); h1 += pl.epart(); }
                                    }
                                    wait_l(t, (pl.e_split() && ep == 0) ? QB_BAR_EACC_HALF : QB_BAR_EACC_FULL, 0x425);
                                    acc_to_smem_operand(te, h0, h1, sd);
                                }
                                arrive_issuer(t, QB_BAR_AE_READY, true);
                            }
                        }
                    }
#pragma unroll 1
                    for (int t = 0; t < NT; t++) {
                        if (pl.has_proj()) wait_l(t, QB_BAR_HACC_FULL, 0x434);
                        else if (pl.L() > 0) {
                            if (pl.e_split()) wait_l(t, QB_BAR_EACC_HALF, 0x437);
                            wait_l(t, QB_BAR_EACC_FULL, 0x435);
                        }
                        const int64_t row = t ? row1 : row0;
                        const bool valid = t ? valid1 : valid0;
                        final_loop(t, valid ? row : 0, valid, t ? code1 : code0, p.loop_steps[ls].cb_blk, last);
                        tc_fence_before();
                        if (!last) {            // tile slot t goes straight into its next step
                            const int code = step_code(row, valid, ls + 1);
                            if (t) code1 = code; else code0 = code;
                            init_loop(t, code, p.loop_steps[ls + 1].t_blk);
                        }
                    }
                }
            }
        } else {
        int64_t kset = 0;
        // per-tile row context, kept in scalars (no runtime-indexed arrays).  It belongs to the set whose tiles are
        // currently initialised: without an out_proj the init of the NEXT set's tile t is issued right after the final
        // epilogue of the current tile t, so tile t's MMAs restart while the other tile is still in its final epilogue.
        int code0 = 0, code1 = 0;
        int64_t beam0 = 0, beam1 = 0, row0 = 0, row1 = 0;
        bool valid0 = false, valid1 = false;
        bool primed = false;        // the current set's tiles were initialised by the previous iteration
        QB_FOR_SETS(it_e, set) {
            const int rb = (int)(kset % 3);
            if (kResident && !primed) mbar_wait((a_rfull + (uint32_t)(rb) * 8u), (uint32_t)((kset / 3) & 1), p.err_flag, 0x610 + rb);
            auto row_ctx = [&](int64_t set, int t, int64_t& row, int64_t& beam, int& code, bool& valid) {
                code = 0; beam = 0; row = 0; valid = false;
                if (kResident) {            // tile slot t, rows 0-63 / 64-127 = code quarter hq of beams 4*set + 2t / + 2t + 1
                    beam = 4 * set + 2 * t + (r >> 6);
                    code = hq * 64 + (r & 63);
                    row = beam * 256 + code;
                    valid = beam < n_beams;
                    return;
                }
                row = (set * NT + t) * QB_TILE_M + r;
                valid = t < NT && row < p.n_rows;
                if (valid) {
                    if (kScore) {
                        beam = (int64_t)((uint32_t)row / (uint32_t)p.C);      // rows per launch < 2^31 (launch_mlp checks)
                        const int a = (int)((uint32_t)row - (uint32_t)beam * (uint32_t)p.C);
                        code = p.A > 0 ? (int)__ldg(p.idx + beam * p.A + a) : a;
                    } else {
                        const int64_t v = (int64_t)((uint32_t)row / (uint32_t)p.F_out);
                        const int parent = p.sel_parent ? (int)__ldg(p.sel_parent + row) : 0;
                        beam = v * p.F_in + parent;
                        code = p.sel_best ? (int)(p.sel_best[row] & 0xffull)           // winner of the fused selection (low bits = code)
                                          : (int)__ldg(p.sel_code + row * p.code_stride + p.code_off);
                    }
                    if (code >= K) code = K - 1;   // never read outside the tables (bad codes are rejected on the host)
                    if (!kScore && p.hist_out && cg == 0) {     // update launch after a fused selection: extend the code history
                        uint8_t* ho = p.hist_out + row * p.hist_M;
                        const uint8_t* hi = p.hist_in + beam * p.hist_M;
                        for (int c = 0; c < p.hist_m; c++) ho[c] = hi[c];
                        ho[p.hist_m] = (uint8_t)code;
                    }
                }
            };
            if (!primed && !kResident) {      // (resident launches derive the context from the set index where it is used)
                row_ctx(set, 0, row0, beam0, code0, valid0);
                row_ctx(set, 1, row1, beam1, code1, valid1);
            }
            tr.ev(1);
            // table row block: columns [c, c + 32) of a [cols/4][K][4] table for `code`, `n` (16 or >= 32) of them.  Kept to
            // 32 columns (32 registers): anything that spills is re-read from L2, which costs more than a second batch.
            auto load_row32 = [&](float4 (&tb)[8], const float* tbl, int code, int c, int n) {
                const float* base = tbl + ((size_t)(c >> 2) * K + code) * 4;
#pragma unroll
                for (int i = 0; i < 4; i++) tb[i] = ldg4(base + (size_t)i * K * 4);
                if (n > 16) {
#pragma unroll
                    for (int i = 4; i < 8; i++) tb[i] = ldg4(base + (size_t)i * K * 4);
                }
            };
            // ---- init: e0 = T_m[code] + u_b over this thread's columns ------------------------------------------------
            auto init_tile = [&](int t, int code, int64_t beam, int rb) {
                const uint32_t tl = lane_base + (uint32_t)t * tile_cols;
                const uint32_t ae_dst = smem_base + pl.smem_ae(t) + (uint32_t)r * 16u;
                const float* up = p.u + beam * De;
                if (!kResident && cg == 0 && r * 32 < De) prefetch_l1(up + r * 32);   // the per-beam row is shared by many rows: pull it into L1
                // 32 columns per batch: operand loads first, one wide TMEM store, four shared-memory k-chunk rows
                const uint32_t us_a = a_beam + (uint32_t)((rb * 2 + t) * 2 + (r >> 6)) * 1024u;
                auto batch = [&](int c, int n, auto&& operands) {     // n = 16 or >= 32
                    uint32_t e[32];
                    auto half16 = [&](int h) {
                        float4 tb[4], ub[4];
                        operands(h, tb, ub);
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            float s0, s1, s2, s3;
                            f2_unpack(f2_add(f2_pack(tb[i].x, tb[i].y), f2_pack(ub[i].x, ub[i].y)), s0, s1);
                            f2_unpack(f2_add(f2_pack(tb[i].z, tb[i].w), f2_pack(ub[i].z, ub[i].w)), s2, s3);
                            e[16 * h + 4 * i + 0] = __float_as_uint(s0); e[16 * h + 4 * i + 1] = __float_as_uint(s1);
                            e[16 * h + 4 * i + 2] = __float_as_uint(s2); e[16 * h + 4 * i + 3] = __float_as_uint(s3);
                        }
                    };
                    auto to_smem = [&](int j) {     // 8 columns -> one fp16 k-chunk row of A_E
                        st_shared_v4(ae_dst + (uint32_t)((c >> 3) + j) * kAkcBytes,
                                     pack_h2(__uint_as_float(e[8 * j]), __uint_as_float(e[8 * j + 1])),
                                     pack_h2(__uint_as_float(e[8 * j + 2]), __uint_as_float(e[8 * j + 3])),
                                     pack_h2(__uint_as_float(e[8 * j + 4]), __uint_as_float(e[8 * j + 5])),
                                     pack_h2(__uint_as_float(e[8 * j + 6]), __uint_as_float(e[8 * j + 7])));
                    };
                    half16(0);
                    if (n > 16) {
                        half16(1);
                        __syncwarp();
                        tmem_st32(tl + pl.tmem_e_col() + c, e);
                        to_smem(0); to_smem(1); to_smem(2); to_smem(3);
                    } else {
                        __syncwarp();
                        tmem_st16p(tl + pl.tmem_e_col() + c, e);
                        to_smem(0); to_smem(1);
                    }
                };
                if (kResident) {
#pragma unroll
                    for (int b = 0; b < kTReg / 8; b++) {       // <= 128 / kColGroups columns per thread in resident mode (planner)
                        const int c = e0c + 32 * b;
                        if (c < e1c)
                            batch(c, e1c - c, [&](int h, float4 (&tb)[4], float4 (&ub)[4]) {
#pragma unroll
                                for (int i = 0; i < 4; i++) {
                                    tb[i] = treg[8 * b + 4 * h + i];
                                    ub[i] = lds4(us_a + (uint32_t)((c >> 2) + 4 * h + i) * 16u);
                                }
                            });
                    }
                } else {
#pragma unroll 1
                    for (int c = e0c; c < e1c; c += 32)
                        batch(c, e1c - c, [&](int h, float4 (&tb)[4], float4 (&ub)[4]) {
                            const float* base = p.t_blk + ((size_t)((c >> 2) + 4 * h) * K + code) * 4;
#pragma unroll
                            for (int i = 0; i < 4; i++) { tb[i] = ldg4(base + (size_t)i * K * 4); ub[i] = ldg4(up + c + 16 * h + 4 * i); }
                        });
                }
                tr.ev(14);
                tmem_wait_st();
                arrive_issuer(t, QB_BAR_AE_READY, true);
                tr.ev(2 + 0x80 * t);
            };
            if (!primed) {
#pragma unroll 1
                for (int t = 0; t < NT; t++)
                    init_tile(t, kResident ? hq * 64 + (r & 63) : (t ? code1 : code0), kResident ? 4 * set + 2 * t + (r >> 6) : (t ? beam1 : beam0), rb);
            }
            // ---- residual blocks -----------------------------------------------------------------------------------
            float acc0 = 0.f, acc1 = 0.f;
            auto wait_bar = [&](int t, int bar, uint32_t code) {
                uint32_t& par = t ? par1 : par0;
                mbar_wait(bar_addr(t, bar), (par >> bar) & 1, p.err_flag, code);
                par ^= 1u << bar;
                tc_fence_after();
            };
            // final epilogue of columns [c0, c1) at accumulator address taddr (o[d0 + c0 ..]); cb = skip codeword of the
            // first 32 columns, fetched by the caller BEFORE it waited for the accumulator
            auto final_cols = [&](uint32_t taddr, int c0, int c1, int d0, int code, int64_t beam, int64_t row, bool valid,
                                  float4 (&cb)[8], float& acc, uint32_t rows_a) {
                const float* src = (kScore ? p.r : p.xhat_in) + beam * D + d0;         // not used in resident mode
                const uint32_t rs_a = rows_a + (uint32_t)d0 * 4u;                        // resident: r_b row in shared memory
                const uint32_t cs_a = smem_base + (uint32_t)pl.smem_tres() + (uint32_t)(r & 63) * 16u;   // resident C_m quarter
                const bool skip = pl.skip() != 0;
                f32x2 axy = 0ull, azw = 0ull;                       // four independent accumulation chains, in packed pairs
                // one block of <= 32 accumulator columns already requested into v (`waited`: their tcgen05.ld has been waited for)
                auto proc32 = [&](uint32_t (&v)[32], int cb0, int n, bool waited0) {
                    if (!kResident && cb0 > c0 && skip) load_row32(cb, p.cb_blk, code, d0 + cb0, n);
                    auto half16 = [&](int h, bool waited) {
                        float4 tv[4], cv[4];
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            const int cc = cb0 + 16 * h + 4 * i;
                            if (kResident) {
                                tv[i] = lds4(rs_a + (uint32_t)cc * 4u);
                                cv[i] = skip ? lds4(cs_a + (uint32_t)((d0 + cc) >> 2) * 1024u) : make_float4(0.f, 0.f, 0.f, 0.f);
                            } else {
                                tv[i] = ldg4(src + cc);
                                cv[i] = cb[4 * h + i];          // all zero when the model has no outer skip
                            }
                        }
                        if (!waited) { tmem_wait_ld(); tr.ev(13); }
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            const int vi = 16 * h + 4 * i;
                            if (kScore) {               // e = r - (v + c), acc += e * e: three packed operations per column pair
                                const f32x2 e01 = f2_sub(f2_pack(tv[i].x, tv[i].y),
                                                         f2_add(f2_pack(__uint_as_float(v[vi]), __uint_as_float(v[vi + 1])), f2_pack(cv[i].x, cv[i].y)));
                                const f32x2 e23 = f2_sub(f2_pack(tv[i].z, tv[i].w),
                                                         f2_add(f2_pack(__uint_as_float(v[vi + 2]), __uint_as_float(v[vi + 3])), f2_pack(cv[i].z, cv[i].w)));
                                axy = f2_fma(e01, e01, axy);
                                azw = f2_fma(e23, e23, azw);
                                continue;
                            }
                            const float o0 = __uint_as_float(v[vi]) + cv[i].x, o1 = __uint_as_float(v[vi + 1]) + cv[i].y;
                            const float o2 = __uint_as_float(v[vi + 2]) + cv[i].z, o3 = __uint_as_float(v[vi + 3]) + cv[i].w;
                            if (valid) {
                                const int d = d0 + cb0 + 16 * h + 4 * i;
                                float4 out = make_float4(tv[i].x + o0, tv[i].y + o1, tv[i].z + o2, tv[i].w + o3);
                                if (p.out_shift) {
                                    const float4 sh = ldg4(p.out_shift + d);
                                    out.x = fmaf(out.x, p.out_scale, sh.x); out.y = fmaf(out.y, p.out_scale, sh.y);
                                    out.z = fmaf(out.z, p.out_scale, sh.z); out.w = fmaf(out.w, p.out_scale, sh.w);
                                } else if (p.out_scale != 1.0f) {
                                    out.x *= p.out_scale; out.y *= p.out_scale; out.z *= p.out_scale; out.w *= p.out_scale;
                                }
                                *reinterpret_cast<float4*>(p.xhat_out + row * D + d) = out;
                            }
                        }
                    };
                    half16(0, waited0);
                    if (n > 16) half16(1, true);
                };
                // (both tcgen05.ld of a 64-column thread in flight together would save a TMEM round trip, but with the 64 table
                // registers of the resident variant the live set no longer fits 232 registers: ptxas spills the whole table row)
#pragma unroll 1
                for (int cb0 = c0; cb0 < c1; cb0 += 32) {
                    uint32_t v[32];
                    tr.ev(12);
                    tmem_ld_cols(taddr + cb0, c1 - cb0, v);
                    proc32(v, cb0, c1 - cb0, false);
                }
                if (kScore) {
                    float ax, ay, az, aw;
                    f2_unpack(axy, ax, ay);
                    f2_unpack(azw, az, aw);
                    acc += (ax + ay) + (az + aw);
                }
            };
#pragma unroll 1
            for (int l = 0; l < pl.L(); l++) {
#pragma unroll 1
                for (int j = 0; j < pl.n_hchunk(); j++) {
                    const int cw = min(pl.hc(), pl.Dh() - j * pl.hc());
                    int c0, c1;
                    group_range(cw, cg, c0, c1);
#pragma unroll 1
                    for (int t = 0; t < NT; t++) {
                        wait_bar(t, QB_BAR_HACC_FULL, 0x404);
                        tr.ev(3 + 0x80 * t);
                        const uint32_t th = lane_base + (uint32_t)t * tile_cols + pl.tmem_h_col();
                        if (kColGroups == 2 && pl.h_split() && cw == pl.hc() && (cw >> 2) <= 32) {
                            // two K halves, each split over the two column groups: the down-projection starts on the
                            // first half while the second is converted
                            if (pl.h_split() == 2) {
                                const int qw = cw >> 2;
                                acc_to_tmem_operand_half(th, cg * qw, qw, 1 + q, true);
                                arrive_issuer(t, QB_BAR_AH_READY, false);
                                acc_to_tmem_operand_half(th, (cw >> 1) + cg * qw, qw, 1 + q, true);
                                arrive_issuer(t, QB_BAR_AH2_READY, false);
                            } else {
                                acc_to_tmem_operand_split(th, cw, cg, 1 + q, [&](int h) { arrive_issuer(t, h ? QB_BAR_AH2_READY : QB_BAR_AH_READY, false); });
                            }
                        } else if (kColGroups == 4 && cw == 128) {     // one 32-column block per warp
                            acc_to_tmem_operand_half(th, cg * 32, 32, 1 + q, pl.h_split() == 2);
                            arrive_issuer(t, QB_BAR_AH_READY, false);
                            if (pl.h_split() && cw == pl.hc()) arrive_issuer(t, QB_BAR_AH2_READY, false);   // (a split plan waits for both)
                        } else {
                            acc_to_tmem_operand(th, c0, c1, 1 + q);
                            arrive_issuer(t, QB_BAR_AH_READY, false);
                            if (kColGroups != 2 && pl.h_split() && cw == pl.hc()) arrive_issuer(t, QB_BAR_AH2_READY, false);   // (a split plan waits for both)
                        }
                        tr.ev(4 + 0x80 * t);
                    }
                }
                if (l + 1 < pl.L() || pl.has_proj()) {
#pragma unroll 1
                    for (int t = 0; t < NT; t++) {
                        const uint32_t te = lane_base + (uint32_t)t * tile_cols + pl.tmem_e_col();
                        const uint32_t sd = smem_base + pl.smem_ae(t) + (uint32_t)r * 16u;
                        // with e_split the first column part is final while the second part's MMAs still run (ONE call site
                        // of the conversion on purpose: a second inlined copy made ptxas spill ~900 B per thread)
                        const int n_ep = pl.e_split() ? 2 : 1;
#pragma unroll 1
                        for (int ep = 0; ep < n_ep; ep++) {
                            int h0 = e0c, h1 = e1c;
                            if (pl.e_split()) {
                                group_range(ep == 0 ? pl.epart() : De - pl.epart(), cg, h0, h1);
                                if (ep) { h0 += pl.epart(); h1 += pl.epart(); }
                            }
                            wait_bar(t, (pl.e_split() && ep == 0) ? QB_BAR_EACC_HALF : QB_BAR_EACC_FULL, 0x405);
                            tr.ev(5 + 0x80 * t);
                            acc_to_smem_operand(te, h0, h1, sd);
                        }
                        arrive_issuer(t, QB_BAR_AE_READY, true);
                        tr.ev(6 + 0x80 * t);
                    }
                }
            }
            // ---- final epilogue straight from Eacc (no out_proj); the skip codeword is only alive here ------------------
            constexpr int kParts = kColGroups > 1 ? kColGroups - 1 : 1;       // dist_part[tile][part][row]
            auto publish_dist = [&](int t, float a, int64_t row, bool valid) {    // the column groups of a row meet in shared memory
                if (cg > 0) sts1(a_dist + (uint32_t)((t * kParts + cg - 1) * QB_TILE_M + r) * 4u, a);
                named_bar_sync(5, kEpiThreads);
                if (cg == 0) {
#pragma unroll
                    for (int g = 0; g < kColGroups - 1; g++) a += lds1(a_dist + (uint32_t)((t * kParts + g) * QB_TILE_M + r) * 4u);
                    if (valid) p.dist[row] = a;
                }
            };
            // Fused selection, resident launches (one beam per vector, 64 candidates of it per tile half): CTA-local arg-min,
            // the winner's o = Eacc + C_m[code] re-read from TMEM into the stash, hand-off to the select warp.
            auto fused_select_resident = [&](int t, float a, int code, bool valid) {
                const int buf = (int)(kset & 1);
                if (cg > 0) sts1(a_dist + (uint32_t)((t * kParts + cg - 1) * QB_TILE_M + r) * 4u, a);
                named_bar_sync(5, kEpiThreads);
                if constexpr (kFuse == 1) {  // distances to shared memory for the select warp; nothing else on this path
                    if (cg == 0) {
#pragma unroll
                        for (int g = 0; g < kColGroups - 1; g++) a += lds1(a_dist + (uint32_t)((t * kParts + g) * QB_TILE_M + r) * 4u);
                        if (kset >= 1) mbar_wait(a_sempty + (uint32_t)t * 8u, (uint32_t)((kset - 1) & 1), p.err_flag, 0x820 + t);
                        sts1(a_stash + (uint32_t)((buf * 2 + t) * 256 + r) * 4u, valid ? a : __uint_as_float(0x7f800000u));
                    }
                    __syncwarp();
                    if ((tid & 31) == 0) mbar_arrive(a_sfull + (uint32_t)t * 8u);
                    return;
                } else {
                if (cg == 0) {
#pragma unroll
                    for (int g = 0; g < kColGroups - 1; g++) a += lds1(a_dist + (uint32_t)((t * kParts + g) * QB_TILE_M + r) * 4u);
                    // (dist bits << 32 | code): distances are >= 0, so the integer order is the (dist, code) order
                    unsigned long long key = valid ? (((unsigned long long)__float_as_uint(a)) << 32) | (unsigned)code : ~0ull;
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) {
                        const unsigned long long o = __shfl_xor_sync(0xffffffffu, key, off);
                        key = o < key ? o : key;
                    }
                    // the select warp has taken the previous set's keys and settled the set before it (whose stash / key
                    // buffer is the one reused now)
                    if (kset >= 1) mbar_wait(a_sempty + (uint32_t)t * 8u, (uint32_t)((kset - 1) & 1), p.err_flag, 0x820 + t);
                    if ((tid & 31) == 0)
                        asm volatile("st.shared.u64 [%0], %1;" ::"r"(a_skey + (uint32_t)(((buf * 2 + t) * 4 + q) * 8)), "l"(key) : "memory");
                }
                named_bar_sync(5, kEpiThreads);
                if (p.xhat_out) {
                    const int h = q >> 1;
                    unsigned long long k0, k1;
                    const uint32_t ka = a_skey + (uint32_t)(((buf * 2 + t) * 4 + 2 * h) * 8);
                    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(k0) : "r"(ka));
                    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(k1) : "r"(ka + 8u));
                    const unsigned long long key = k0 < k1 ? k0 : k1;
                    const int wrow = 64 * h + ((int)(key & 0xffffffffull) & 63);        // tile row of the local winner
                    if (key != ~0ull && (wrow >> 5) == q) {                              // warp-uniform: this warp owns that row
                        const uint32_t taddr = lane_base + (uint32_t)t * tile_cols + pl.tmem_e_col();
                        const uint32_t cs_a = smem_base + (uint32_t)pl.smem_tres() + (uint32_t)(wrow & 63) * 16u;
                        const uint32_t st = a_stash + (uint32_t)(((buf * 2 + t) * 2 + h) * 128) * 4u;
                        const bool skip = pl.skip() != 0;
#pragma unroll 1
                        for (int c = o0c; c < o1c; c += 32) {
                            const int n = o1c - c;
                            uint32_t v[32];
                            tmem_ld_cols(taddr + c, n, v);
                            tmem_wait_ld();
                            if ((tid & 31) == (wrow & 31)) {
#pragma unroll
                                for (int i = 0; i < 8; i++) {
                                    if (4 * i < n) {
                                        const float4 cv = skip ? lds4(cs_a + (uint32_t)((c >> 2) + i) * 1024u) : make_float4(0.f, 0.f, 0.f, 0.f);
                                        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(st + (uint32_t)(c + 4 * i) * 4u),
                                                     "f"(__uint_as_float(v[4 * i]) + cv.x), "f"(__uint_as_float(v[4 * i + 1]) + cv.y),
                                                     "f"(__uint_as_float(v[4 * i + 2]) + cv.z), "f"(__uint_as_float(v[4 * i + 3]) + cv.w)
                                                     : "memory");
                                    }
                                }
                            }
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if ((tid & 31) == 0) mbar_arrive(a_sfull + (uint32_t)t * 8u);
                }   // full
            };
            const int64_t set_next = set + set_stride;
            const bool has_next = more_sets(set_next);
            if (!pl.has_proj()) {
                const int rbn = (int)((kset + 1) % 3);
#pragma unroll 1
                for (int t = 0; t < NT; t++) {
                    float4 cb[8];
#pragma unroll
                    for (int i = 0; i < 8; i++) cb[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    // inputs of the final epilogue travel while the last down-projection runs
                    if (!kResident && cg == 0 && r * 32 < D) prefetch_l1((kScore ? p.r : p.xhat_in) + (t ? beam1 : beam0) * D + r * 32);
                    if (!kResident && pl.skip() && o0c < o1c) load_row32(cb, p.cb_blk, t ? code1 : code0, o0c, o1c - o0c);
                    if (pl.L() > 0 && pl.e_split()) wait_bar(t, QB_BAR_EACC_HALF, 0x407);
                    if (pl.L() > 0) wait_bar(t, QB_BAR_EACC_FULL, 0x405);
                    tr.ev(5 + 0x80 * t);
                    float a = 0.f;
                    // resident launches: code / beam / row follow from (set, tile slot, thread); nothing is carried in registers
                    const int64_t beam_t = kResident ? 4 * set + 2 * t + (r >> 6) : (t ? beam1 : beam0);
                    const int code_t = kResident ? hq * 64 + (r & 63) : (t ? code1 : code0);
                    const int64_t row_t = kResident ? beam_t * 256 + code_t : (t ? row1 : row0);
                    const bool valid_t = kResident ? beam_t < n_beams : (t ? valid1 : valid0);
                    final_cols(lane_base + (uint32_t)t * tile_cols + pl.tmem_e_col(), o0c, o1c, 0, code_t, beam_t,
                               row_t, valid_t, cb, a,
                               a_beam + (uint32_t)(((rb * 2 + t) * 2 + (r >> 6)) * 256 + De) * 4u);
                    tc_fence_before();
                    tr.ev(7 + 0x80 * t);
                    if (kFuseA) fused_select_resident(t, a, code_t, valid_t);
                    else if (kScore) publish_dist(t, a, row_t, valid_t);
                    if (has_next) {         // tile slot t is free: start its next set now
                        if (kResident && t == 0)
                            mbar_wait((a_rfull + (uint32_t)(rbn) * 8u), (uint32_t)(((kset + 1) / 3) & 1), p.err_flag, 0x610 + rbn);
                        if (kResident) {
                            init_tile(t, code_t, 4 * set_next + 2 * t + (r >> 6), rbn);
                        } else {
                            if (t) row_ctx(set_next, 1, row1, beam1, code1, valid1); else row_ctx(set_next, 0, row0, beam0, code0, valid0);
                            init_tile(t, t ? code1 : code0, t ? beam1 : beam0, rbn);
                        }
                    }
                }
                primed = has_next;
            }
            // ---- out_proj chunks ---------------------------------------------------------------------------------------
            if (pl.has_proj()) {
                for (int qq = 0; qq < pl.n_ochunk(); qq++) {
                    const int cw = min(pl.oc(), D - qq * pl.oc());
                    int c0, c1;
                    group_range(cw, cg, c0, c1);
#pragma unroll 1
                    for (int t = 0; t < NT; t++) {
                        float4 cb[8];
#pragma unroll
                        for (int i = 0; i < 8; i++) cb[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (pl.skip() && c0 < c1) load_row32(cb, p.cb_blk, t ? code1 : code0, qq * pl.oc() + c0, c1 - c0);
                        if (qq == 0 && cg == 0 && r * 32 < D) prefetch_l1((kScore ? p.r : p.xhat_in) + (t ? beam1 : beam0) * D + r * 32);
                        wait_bar(t, QB_BAR_HACC_FULL, 0x414);
                        const uint32_t ta = lane_base + (uint32_t)t * tile_cols + pl.tmem_h_col();
                        float a = t ? acc1 : acc0;
                        final_cols(ta, c0, c1, qq * pl.oc(), t ? code1 : code0, t ? beam1 : beam0, t ? row1 : row0, t ? valid1 : valid0, cb, a, 0u);
                        if (t) acc1 = a; else acc0 = a;
                        if (qq + 1 < pl.n_ochunk()) arrive_issuer(t, QB_BAR_HACC_FREE, false);
                        else tc_fence_before();
                    }
                }
            }
            if constexpr (kFuseB) {
                // ================================================================= fused selection B (one tile per CTA, out_proj)
                // reference QINCoStep.encode, qinco_base.py:343-372: distances -> topk(F_out) -> gathers of xhat' / history
                const int R = p.F_in * p.C, seg = R < QB_TILE_M ? R : QB_TILE_M, vpt = QB_TILE_M / seg, F_out = p.F_out;
                const int s_idx = r / seg, r_in = r - s_idx * seg;
                const int seg_in_v = (int)(set % fb_spv);                         // which tile of its vector this is
                const int64_t n_vec = p.n_rows / R;
                const uint32_t a_sd = a_dist;                                     // dist totals of the tile: dist_part[0][0][..]
                const uint32_t a_rd = opaque(smem_u32(&selb_run_d[0])), a_rf = opaque(smem_u32(&selb_run_flat[0]));
                float a = acc0;
                if (cg > 0) sts1(a_dist + (uint32_t)((cg - 1) * QB_TILE_M + r) * 4u, a);
                named_bar_sync(5, kEpiThreads);
                if (cg == 0) {
#pragma unroll
                    for (int g = 0; g < kColGroups - 1; g++) a += lds1(a_dist + (uint32_t)(g * QB_TILE_M + r) * 4u);
                    if (!valid0) a = __uint_as_float(0x7f800000u);
                    sts1(a_sd + (uint32_t)r * 4u, a);       // (the slot held this row's partial, which only this thread reads)
                    selb_take[r] = 0xff;
                }
                named_bar_sync(5, kEpiThreads);                                   // totals published
                if (cg == 0) {          // rank inside the segment, ties to the lower row (= lower flat index, like torch.topk)
                    const uint32_t base = a_sd + (uint32_t)(s_idx * seg) * 4u;
                    int cnt = 0;
                    for (int j = 0; j < ((p.dbg & 4) ? 0 : seg); j += 4) {
                        const float4 d4 = lds4(base + (uint32_t)j * 4u);
                        cnt += (d4.x < a || (d4.x == a && j < r_in)) + (d4.y < a || (d4.y == a && j + 1 < r_in)) +
                               (d4.z < a || (d4.z == a && j + 2 < r_in)) + (d4.w < a || (d4.w == a && j + 3 < r_in));
                    }
                    if (p.dbg & 4) cnt = r_in;
                    if (cnt < F_out) {
                        selb_lrow[s_idx * F_out + cnt] = (uint8_t)r;
                        if (fb_spv == 1) {          // the tile holds the whole vector: the local ranking is final, slot = rank
                            selb_take[r] = (uint8_t)cnt;
                            selb_run_d[s_idx * F_out + cnt] = a;
                            selb_run_flat[s_idx * F_out + cnt] = (uint16_t)r_in;
                            selb_run_slot[s_idx * F_out + cnt] = (uint8_t)cnt;
                        }
                    }
                    if (fb_spv == 1 && r_in == 0) selb_nrun[s_idx] = (uint8_t)F_out;
                }
                named_bar_sync(5, kEpiThreads);
                if (fb_spv > 1) {
                    if (warp == 0) {    // one warp merges the tile's sorted winners into the vector's sorted running list (<= 64 entries)
                        const int lane = tid & 31;
                        const int n_run = seg_in_v == 0 ? 0 : (int)selb_nrun[0], k_loc = F_out, total = n_run + k_loc;
                        uint32_t used = 0, n_new_before = 0;
                        float e_d[2]; int e_flat[2], e_slot[2], e_rank[2], e_lr[2]; bool e_new[2], e_surv[2];
#pragma unroll
                        for (int rr = 0; rr < 2; rr++) {
                            const int e = lane + 32 * rr;
                            e_surv[rr] = false; e_new[rr] = false; e_d[rr] = 0.f; e_flat[rr] = 0; e_slot[rr] = 0; e_rank[rr] = 0; e_lr[rr] = 0;
                            if (e < total) {
                                int before = 0;
                                if (e < n_run) {            // running entry: local entries precede it only when strictly smaller
                                    e_d[rr] = selb_run_d[e]; e_flat[rr] = selb_run_flat[e]; e_slot[rr] = selb_run_slot[e];
                                    for (int j = 0; j < k_loc; j++) before += lds1(a_sd + (uint32_t)selb_lrow[j] * 4u) < e_d[rr];
                                    e_rank[rr] = e + before;
                                } else {                    // newcomer: running entries precede it on ties (earlier tile = lower flat index)
                                    const int j = e - n_run;
                                    e_lr[rr] = selb_lrow[j];
                                    e_d[rr] = lds1(a_sd + (uint32_t)e_lr[rr] * 4u);
                                    e_flat[rr] = seg_in_v * QB_TILE_M + e_lr[rr];
                                    e_new[rr] = true;
                                    for (int i = 0; i < n_run; i++) before += selb_run_d[i] <= e_d[rr];
                                    e_rank[rr] = j + before;
                                }
                                e_surv[rr] = e_rank[rr] < F_out;
                            }
                            used |= __reduce_or_sync(0xffffffffu, (e_surv[rr] && !e_new[rr]) ? (1u << e_slot[rr]) : 0u);
                        }
                        const uint32_t free_slots = ~used;
#pragma unroll
                        for (int rr = 0; rr < 2; rr++) {
                            const uint32_t nb = __ballot_sync(0xffffffffu, e_surv[rr] && e_new[rr]);
                            if (e_surv[rr] && e_new[rr]) {
                                const uint32_t t_idx = n_new_before + (uint32_t)__popc(nb & ((1u << lane) - 1u));
                                e_slot[rr] = (int)__fns(free_slots, 0, (int)t_idx + 1);
                                selb_take[e_lr[rr]] = (uint8_t)e_slot[rr];
                            }
                            n_new_before += (uint32_t)__popc(nb);
                        }
                        __syncwarp();       // every lane has read the old list: overwrite it in merged order
#pragma unroll
                        for (int rr = 0; rr < 2; rr++) {
                            if (e_surv[rr]) {
                                selb_run_d[e_rank[rr]] = e_d[rr];
                                selb_run_flat[e_rank[rr]] = (uint16_t)e_flat[rr];
                                selb_run_slot[e_rank[rr]] = (uint8_t)e_slot[rr];
                            }
                        }
                        if (lane == 0) selb_nrun[0] = (uint8_t)(total < F_out ? total : F_out);
                    }
                    named_bar_sync(5, kEpiThreads);
                }
                {                       // the granted rows park their o (still in TMEM) in the stash; xhat_parent and the skip
                                        // codeword are added at emit time, with coalesced loads, once per vector
                    const int slot = (int)selb_take[r];
                    if (!(p.dbg & 1) && __any_sync(0xffffffffu, slot != 0xff)) {
                        int c0, c1;
                        group_range(D, cg, c0, c1);
                        const uint32_t ta = lane_base + (pl.has_proj() ? pl.tmem_h_col() : pl.tmem_e_col());
                        const uint32_t st = opaque(smem_u32(&selb_stash[0])) + (uint32_t)((s_idx * F_out + (slot & 0x3f)) * D) * 4u;
#pragma unroll 1
                        for (int c = c0; c < c1; c += 32) {
                            const int n = c1 - c;
                            uint32_t v[32];
                            tmem_ld_cols(ta + c, n, v);
                            tmem_wait_ld();
                            if (slot != 0xff) {
#pragma unroll
                                for (int i = 0; i < 8; i++)
                                    if (4 * i < n)
                                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(st + (uint32_t)(c + 4 * i) * 4u), "r"(v[4 * i]),
                                                     "r"(v[4 * i + 1]), "r"(v[4 * i + 2]), "r"(v[4 * i + 3])
                                                     : "memory");
                            }
                        }
                    }
                }
                tc_fence_before();
                if (seg_in_v == (int)fb_spv - 1 && !(p.dbg & 2)) {       // the vector(s) are complete: emit xhat' and the histories in rank order
                    named_bar_sync(5, kEpiThreads);
                    const int d4n = D >> 2;
                    const int items = vpt * F_out * d4n;
                    for (int it2 = tid; it2 < items; it2 += kEpiThreads) {
                        const int d4 = it2 % d4n, e = it2 / d4n, sgm = e / F_out, j = e - sgm * F_out;
                        const int64_t vs = fb_spv > 1 ? set / fb_spv : set * vpt + sgm;
                        if (vs < n_vec && j < (int)selb_nrun[sgm]) {
                            // xhat'_j = xhat_parent + (o + C_m[code])     (same association as the update launch: bit-identical)
                            const int slot = selb_run_slot[sgm * F_out + j];
                            const int flat = selb_run_flat[sgm * F_out + j];
                            const int parent = flat / p.C, slot_a = flat - parent * p.C;
                            const float4 o = lds4(opaque(smem_u32(&selb_stash[0])) + (uint32_t)(((sgm * F_out + slot) * D) + 4 * d4) * 4u);
                            const float4 xi = ldg4(p.xhat_in + (vs * p.F_in + parent) * D + 4 * d4);
                            float4 cv = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (pl.skip()) {
                                const int code = p.A > 0 ? (int)__ldg(p.idx + (vs * p.F_in + parent) * p.A + slot_a) : slot_a;
                                cv = ldg4(p.cb_blk + ((size_t)d4 * K + code) * 4);
                            }
                            *reinterpret_cast<float4*>(p.xhat_out + (vs * F_out + j) * D + 4 * d4) =
                                make_float4(xi.x + (o.x + cv.x), xi.y + (o.y + cv.y), xi.z + (o.z + cv.z), xi.w + (o.w + cv.w));
                        }
                    }
                    if (tid < vpt * F_out) {
                        const int sgm = tid / F_out, j = tid - sgm * F_out;
                        const int64_t vs = fb_spv > 1 ? set / fb_spv : set * vpt + sgm;
                        if (vs < n_vec && j < (int)selb_nrun[sgm]) {
                            const int flat = selb_run_flat[sgm * F_out + j];
                            const int parent = flat / p.C, slot_a = flat - parent * p.C;
                            const int code = p.A > 0 ? (int)__ldg(p.idx + (vs * p.F_in + parent) * p.A + slot_a) : slot_a;
                            uint8_t* ho = p.hist_out + (vs * F_out + j) * p.hist_M;
                            const uint8_t* hi = p.hist_in + (vs * p.F_in + parent) * p.hist_M;
                            for (int c = 0; c < p.hist_m; c++) ho[c] = hi[c];
                            ho[p.hist_m] = (uint8_t)code;
                        }
                    }
                    (void)a_rd; (void)a_rf;
                }
            } else if (pl.has_proj() && kScore) {
                publish_dist(0, acc0, row0, valid0);
                if (NT > 1) publish_dist(1, acc1, row1, valid1);
            }
            if (kResident) mbar_arrive((a_rempty + (uint32_t)(rb) * 8u));
            tr.ev(8);
            kset++;
        }
        }   // !kLoop
    }

    // ---- teardown ------------------------------------------------------------------------------------------------
    tc_fence_before();
    __syncthreads();
    if (kPair || kMcast) cluster_sync_all();      // no CTA of a pair leaves while the other may still signal its barriers / write its smem
    tc_fence_after();
    if (warp == kMmaWarp) {
        if (kPair)
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                         "r"((uint32_t)pl.tmem_alloc_cols())
                         : "memory");
        else
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                         "r"((uint32_t)pl.tmem_alloc_cols())
                         : "memory");
    }
}

int mlp_plan_view(const QbStepPlan& plan) { return plan_is_s128(plan) ? 1 : (plan_is_l384(plan) ? 2 : 0); }

cudaError_t mlp_set_smem_attr(int smem_bytes) {
    // the attribute belongs to the function, not to a model: only ever raise it (several models share the process)
    static int current[64] = {0};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    if (smem_bytes <= current[dev]) return cudaSuccess;
    const void* fns[] = {(const void*)qb_mlp_kernel<true, false, false>, (const void*)qb_mlp_kernel<true, true, false>,
                         (const void*)qb_mlp_kernel<false, false, false>, (const void*)qb_mlp_kernel<true, false, true>,
                         (const void*)qb_mlp_kernel<true, true, true>, (const void*)qb_mlp_kernel<false, false, true>,
                         (const void*)qb_mlp_kernel<false, false, false, true>, (const void*)qb_mlp_kernel<true, true, false, false, 1>, (const void*)qb_mlp_kernel<true, true, false, false, 2>,
                         (const void*)qb_mlp_kernel<true, true, false, false, 1, false, 1>,
                         (const void*)qb_mlp_kernel<true, false, false, false, 3, true, 2>, (const void*)qb_mlp_kernel<true, false, false, false, 0, true, 2>,
                         (const void*)qb_mlp_kernel<false, false, false, false, 0, true, 2>, (const void*)qb_mlp_kernel<false, false, false, true, 0, true, 2>,
                         (const void*)qb_mlp_kernel<true, false, false, false, 0, true>, (const void*)qb_mlp_kernel<false, false, false, false, 0, true>,
                         (const void*)qb_mlp_kernel<false, false, false, true, 0, true>,
                         (const void*)qb_mlp_kernel<true, false, false, false, 3, false>, (const void*)qb_mlp_kernel<true, false, false, false, 3, true>};
    for (const void* f : fns) {
        // variants with more static shared memory (fused selection B) can take less dynamic memory; the plans they run
        // (one tile per CTA) stay below that, and launch_mlp refuses anything else
        cudaFuncAttributes fa;
        e = cudaFuncGetAttributes(&fa, f);
        if (e != cudaSuccess) return e;
        const int cap = 232448 - (int)fa.sharedSizeBytes;
        e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes < cap ? smem_bytes : cap);
        if (e != cudaSuccess) return e;
    }
    current[dev] = smem_bytes;
    return e;
}

cudaError_t launch_mlp(const MlpParams& p, int n_sm, cudaStream_t stream) {
    if (p.n_rows <= 0) return cudaSuccess;
    if (p.n_rows >= (1ll << 31) - 256) return cudaErrorInvalidValue;   // 32-bit row arithmetic in the kernel
    const int64_t n_tiles = (p.n_rows + QB_TILE_M - 1) / QB_TILE_M;
    int64_t n_sets = (n_tiles + p.plan.n_tiles - 1) / p.plan.n_tiles;
    // resident tables: score launches over all 256 codes of every beam (2 tiles per beam), shape qualified by the planner
    const bool resident = p.mode == QB_MODE_SCORE && p.A == 0 && p.C == 256 && p.plan.K == 256 && p.plan.smem_tres >= 0 &&
                          p.plan.n_tiles == 2 && (p.n_rows & 255) == 0 && n_sm >= 4;
    const bool pair = p.plan.pair != 0 && n_sm >= 2;     // the weights are packed for the pair kernel: no other choice
    if (p.plan.pair && !pair) return cudaErrorInvalidConfiguration;
    const bool loop = p.n_loop_steps > 0;
    // fused selection needs the resident score variant with one beam per vector, all its CTAs co-resident (they wait for
    // each other's reports) and the selection state; anything else must go through the unfused launches
    if (p.fuse == 3) {      // in-CTA selection: one tile per CTA with a single out_proj chunk, whole vectors per tile or whole tiles per vector
        const int R = p.F_in * p.C;
        const int seg = R < QB_TILE_M ? R : QB_TILE_M, vpt = QB_TILE_M / seg;
        if (p.mode != QB_MODE_SCORE || resident || pair || p.plan.n_tiles != 1 || !p.plan.has_proj || p.plan.n_ochunk != 1 ||
            !((R < QB_TILE_M && QB_TILE_M % R == 0) || R % QB_TILE_M == 0) || p.sel_spv != (R > QB_TILE_M ? R / QB_TILE_M : 1) ||
            p.plan.smem_total > 196608 || p.F_out > 32 || p.F_out > R || R < 4 || vpt * p.F_out * p.plan.D > 4096 || (int64_t)R * QB_TILE_M > 65535 * (int64_t)QB_TILE_M ||
            !p.xhat_in || !p.xhat_out || !p.hist_out || !p.hist_in)
            return cudaErrorInvalidConfiguration;
    } else
    if (p.fuse && (!resident || pair || p.F_in != 1 || p.F_out != 1 || !p.sel_best || p.plan.D > 128)) return cudaErrorInvalidConfiguration;
    if (p.fuse == 2 && (!p.sel_cnt || !p.hist_out || !p.xhat_in)) return cudaErrorInvalidConfiguration;
    if (loop && (p.mode != QB_MODE_APPLY || pair || p.plan.n_ops_pre <= 0 || p.n_loop_steps > QB_MAX_LOOP_STEPS || p.plan.n_ochunk > 1))
        return cudaErrorInvalidConfiguration;
    const bool mcast = p.plan.mcast != 0 && !pair && !resident && n_sm >= 2;     // weight multicast over 2-CTA clusters
    int grid;
    if (resident) {
        const int64_t sets = ((p.n_rows >> 8) + 3) / 4;       // per code quarter
        const int64_t per_quarter = sets < n_sm / 4 ? sets : n_sm / 4;
        grid = (int)(4 * per_quarter);                         // a multiple of 4: CTA pairs are code quarters (0,1) / (2,3)
    } else {
        grid = (int)(n_sets < n_sm ? n_sets : n_sm);
        if (p.fuse == 3) {          // vectors (groups of sel_spv tiles) are strided over the CTAs
            const int64_t groups = (n_sets + p.sel_spv - 1) / p.sel_spv;
            grid = (int)(groups < n_sm ? groups : n_sm);
        }
        if (pair || mcast) grid = (grid + 1) & ~1;             // whole pairs; a peer without work of its own follows its leader
        if ((pair || mcast) && grid > n_sm) grid = n_sm & ~1;
    }
    auto launch = [&](const MlpParams& q0) {
        MlpParams q = q0;
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3((unsigned)grid);
        cfg.blockDim = dim3(kThreads);
        cfg.dynamicSmemBytes = (size_t)q.plan.smem_total;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (pair || mcast) ? 2 : 1;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        static const bool fixed_ok = getenv("QB_NO_FIXED_SHAPE") == nullptr;
        const bool l384 = fixed_ok && mcast && plan_is_l384(q.plan);
        if (loop && l384) return cudaLaunchKernelEx(&cfg, qb_mlp_kernel<false, false, false, true, 0, true, 2>, q);
        if (loop && mcast) return cudaLaunchKernelEx(&cfg, qb_mlp_kernel<false, false, false, true, 0, true>, q);
        if (loop) return cudaLaunchKernelEx(&cfg, qb_mlp_kernel<false, false, false, true>, q);
        if (q.fuse == 3) {
            if (l384) return cudaLaunchKernelEx(&cfg, qb_mlp_kernel<true, false, false, false, 3, true, 2>, q);
            if (mcast) return cudaLaunchKernelEx(&cfg, qb_mlp_kernel<true, false, false, false, 3, true>, q);
            return cudaLaunchKernelEx(&cfg, qb_mlp_kernel<true, false, false, false, 3, false>, q);
        }
        if (l384) {
            if (q.mode == QB_MODE_SCORE) return cudaLaunchKernelEx(&cfg, qb_mlp_kernel<true, false, false, false, 0, true, 2>, q);
            return cudaLaunchKernelEx(&cfg, qb_mlp_kernel<false, false, false, false, 0, true, 2>, q);
        }
        if (mcast) {
            if (q.mode == QB_MODE_SCORE) return cudaLaunchKernelEx(&cfg, qb_mlp_kernel<true, false, false, false, 0, true>, q);
            return cudaLaunchKernelEx(&cfg, qb_mlp_kernel<false, false, false, false, 0, true>, q);
        }
        if (pair) {
            if (resident) return cudaLaunchKernelEx(&cfg, qb_mlp_kernel<true, true, true>, q);
            if (q.mode == QB_MODE_SCORE) return cudaLaunchKernelEx(&cfg, qb_mlp_kernel<true, false, true>, q);
            return cudaLaunchKernelEx(&cfg, qb_mlp_kernel<false, false, true>, q);
        }
        if (resident && q.fuse == 1 && fixed_ok && plan_is_s128(q.plan)) return cudaLaunchKernelEx(&cfg, qb_mlp_kernel<true, true, false, false, 1, false, 1>, q);
        if (resident && q.fuse == 1) return cudaLaunchKernelEx(&cfg, qb_mlp_kernel<true, true, false, false, 1>, q);
        if (resident && q.fuse == 2) return cudaLaunchKernelEx(&cfg, qb_mlp_kernel<true, true, false, false, 2>, q);
        if (resident) return cudaLaunchKernelEx(&cfg, qb_mlp_kernel<true, true, false>, q);
        if (q.mode == QB_MODE_SCORE) return cudaLaunchKernelEx(&cfg, qb_mlp_kernel<true, false, false>, q);
        return cudaLaunchKernelEx(&cfg, qb_mlp_kernel<false, false, false>, q);
    };
    // Debug: QB_MLP_TRACE=<file>[:<launch index>] dumps the event log of CTA 0 for one launch (synchronises).
    static const char* trace_env = getenv("QB_MLP_TRACE");
    static int trace_at = -1, launch_no = 0;
    if (trace_env && trace_at < 0) {
        const char* c = strrchr(trace_env, ':');
        trace_at = c ? atoi(c + 1) : 0;
    }
    if (trace_env && launch_no++ == trace_at) {
        MlpParams q = p;
        const size_t bytes = (3 * QB_TRACE_EVENTS + 512) * sizeof(unsigned long long);
        cudaMalloc((void**)&q.trace, bytes);
        cudaMemsetAsync(q.trace, 0, bytes, stream);
        cudaError_t le = launch(q);
        if (le != cudaSuccess) return le;
        cudaStreamSynchronize(stream);
        std::vector<unsigned long long> h(3 * QB_TRACE_EVENTS + 512);
        cudaMemcpy(h.data(), q.trace, bytes, cudaMemcpyDeviceToHost);
        cudaFree(q.trace);
        std::string path(trace_env);
        const size_t colon = path.rfind(':');
        if (colon != std::string::npos) path = path.substr(0, colon);
        if (FILE* f = fopen(path.c_str(), "w")) {
            fprintf(f, "# grid %d rows %lld mode %d\n", grid, (long long)p.n_rows, p.mode);
            fprintf(f, "# smid of CTA 0..:");
            for (int i = 0; i < 512 && i < grid; i++) fprintf(f, " %llu", h[3 * QB_TRACE_EVENTS + i] - 1);
            fprintf(f, "\n");
            for (int r = 0; r < 3; r++)
                for (int i = 0; i < QB_TRACE_EVENTS && h[r * QB_TRACE_EVENTS + i]; i++)
                    fprintf(f, "%d %llu 0x%llx\n", r, h[r * QB_TRACE_EVENTS + i] >> 16, h[r * QB_TRACE_EVENTS + i] & 0xffff);
            fclose(f);
        }
        return cudaGetLastError();
    }
    cudaError_t le = launch(p);
    return le != cudaSuccess ? le : cudaGetLastError();
}

}  // namespace qb
