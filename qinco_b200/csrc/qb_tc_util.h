// PTX helpers shared by the small tcgen05 kernels (qb_prep_tc.cu, qb_ivf_tc.cu): mbarriers with bounded waits, TMA bulk
// copies, K-major no-swizzle UMMA descriptors (8 x 16 B core matrices, SBO = 128 B), tcgen05.mma / ld / commit.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace qb {
namespace tc {

constexpr uint32_t kDescHi = (128u >> 4) | (1u << 14);     // SBO = 128 B, descriptor version 1

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, uint32_t* err_flag, uint32_t code) {
    uint32_t spins = 0;
    uint64_t t0 = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++spins & 1023u) == 0) {       // bounded: a protocol bug traps and reports instead of hanging the GPU
            uint64_t now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            if (now - t0 > 4000000000ull) { if (err_flag) atomicExch(err_flag, code); __threadfence_system(); __trap(); }
        }
    }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mma_f16(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "mov.b64 da, {%1, %5};\n\tmov.b64 db, {%2, %5};\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}" ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(acc), "r"(kDescHi) : "memory");
}
// the same with operand cursors that advance in place (low descriptor words += step): ~5 SASS instructions per MMA when
// unrolled -- a lone issuing lane retires one instruction per ~6.5 cycles, so an N = 128 MMA (64 cycles) leaves ~10
__device__ __forceinline__ void mma_f16_step(uint32_t d_tmem, uint32_t& a_lo, uint32_t& b_lo, uint32_t idesc, uint32_t acc, uint32_t a_step,
                                             uint32_t b_step) {
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "mov.b64 da, {%0, %5};\n\tmov.b64 db, {%1, %5};\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%2], da, db, %3, p;\n\t"
                 "add.u32 %0, %0, %6;\n\tadd.u32 %1, %1, %7;\n\t}"
                 : "+r"(a_lo), "+r"(b_lo) : "r"(d_tmem), "r"(idesc), "r"(acc), "r"(kDescHi), "r"(a_step), "r"(b_step) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t pack_h2(float a, float b) { __half2 h = __floats2half2_rn(a, b); return *reinterpret_cast<uint32_t*>(&h); }

// 8 consecutive values of an operand row -> fp16 hi / lo k-chunk rows (16 B each)
__device__ __forceinline__ void put_hi_lo(uint32_t dst_hi, uint32_t dst_lo, const float (&x)[8]) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const __half2 h = __floats2half2_rn(x[2 * k], x[2 * k + 1]);
        const float2 hf = __half22float2(h);
        hi[k] = *reinterpret_cast<const uint32_t*>(&h);
        lo[k] = pack_h2(x[2 * k] - hf.x, x[2 * k + 1] - hf.y);
    }
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst_hi), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]) : "memory");
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst_lo), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]) : "memory");
}


}  // namespace tc
}  // namespace qb
