// Fused implicit-codebook MLP for one quantisation step, sm_100a (tcgen05 / TMEM / TMA bulk copy).
//
// Replaces, for a 128-row tile of candidates, the reference's chain of separate launches
// (reference qinco/model/qinco_base.py:262-280 = in_proj, QConcat :60-64, L x QBlockFFN :93-97, out_proj, skip;
// the distance of :343-345 and the x-hat update of :363-369) with ONE persistent kernel:
//
//   warps 0-3  epilogue: thread t owns row t of the tile == TMEM lane t.
//              init    e0 = T_m[code] + u_b  -> fp32 into the TMEM residual accumulator (tcgen05.st)
//                                             -> fp16 into the A_E operand tile in shared memory
//              H-epi   relu(Hacc) -> packed fp16 written back IN PLACE into TMEM (A operand of the down-projection)
//              E-epi   Eacc -> fp16 -> A_E                               (per residual block)
//              final   score: dist = ||r_b - o||^2 (fp32)   apply: xhat_out = xhat_b + o
//              (TMEM reads are latency-bound, ~130 cycles per tcgen05.ld: every phase keeps the next load in flight
//               while it converts the current one; table rows are fetched 64 columns at a time, all loads up front)
//   warp 4     producer: streams the pre-packed fp16 weight slabs with cp.async.bulk (TMA) into an mbarrier ring
//   warp 5     MMA issuer: walks the op list (qb_plan.h, in the kernel-parameter constant bank) and issues
//              tcgen05.mma (M=128, kind::f16, fp32 accumulate in TMEM): up-projection A from shared memory,
//              down-projection A from TMEM; completion is signalled with tcgen05.commit
//
// Warps 4 and 5 run converged and only predicate the asynchronous instructions with elect.sync, so descriptors stay
// in uniform registers (no per-MMA divergence handling).  The fp32 residual stream never leaves TMEM; activations
// never touch HBM.  When a tile needs <= 256 TMEM columns two CTAs share an SM (plan.ctas_per_sm): one CTA's table
// gathers and distance epilogue overlap the other CTA's MMAs.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "qb_dev.h"

namespace qb {

namespace {

constexpr int kEpiThreads = 128;
constexpr int kThreads = 192;
constexpr int kAkcBytes = QB_TILE_M * 16;   // bytes of one 8-element k-chunk of an A operand tile

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
// Bounded wait: a protocol bug must trap (and report) instead of hanging the GPU.
__device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity, uint32_t* err_flag, uint32_t code) {
    const uint64_t t0 = globaltimer();
    while (!mbar_try_wait(bar, parity)) {
        if (globaltimer() - t0 > 4000000000ull) {   // 4 s
            if (err_flag) atomicExch(err_flag, code);
            __threadfence_system();
            __trap();
        }
    }
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, uint32_t* err_flag, uint32_t code) {
#pragma unroll 1
    for (int i = 0; i < 64; i++)
        if (mbar_try_wait(bar, parity)) return;
    mbar_wait_slow(bar, parity, err_flag, code);
}

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
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

// D[tmem] (+)= A[smem] . B[smem]^T, M=128, kind::f16 (fp16 inputs, fp32 accumulate)
__device__ __forceinline__ void tc_mma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        :
        : "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem]^T  (A: 128 lanes x 8 columns of packed fp16 pairs per K=16)
__device__ __forceinline__ void tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        :
        : "r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// UMMA shared-memory descriptor, K-major, SWIZZLE_NONE ("interleave") canonical layout
//   ((8,m),(8,2)) : ((16 B, SBO), (2 B, LBO))      (cute/atom/mma_traits_sm100.hpp, make_umma_desc<Major::K>)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = (uint64_t)((saddr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;   // descriptor version 1 (sm_100)
    return d;                 // base_offset 0, lbo_mode 0, layout_type 0 = SWIZZLE_NONE
}
// kind::f16 instruction descriptor: D=f32, A=B=f16, both K-major, M=128 (cute/arch/mma_sm100_desc.hpp InstrDescriptor)
__device__ __forceinline__ uint32_t umma_idesc(uint32_t n) {
    return (1u << 4) | ((n >> 3) << 17) | ((uint32_t)(QB_TILE_M >> 4) << 24);
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
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]),
                 "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t pack_h2_relu(float a, float b) {
    __half2 h = __hmax2(__floats2half2_rn(a, b), __float2half2_rn(0.f));
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
// same load, but pinned in program order (asm volatile): for operands that should be fetched just in time from L1
// instead of being hoisted by the compiler into a long-lived register block
__device__ __forceinline__ float4 ldg4_jit(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

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

// fp32 accumulator columns [0, cw) at `taddr` -> fp16 -> A_E k-chunks starting at `sdst` (shared memory).
// One tcgen05.ld stays in flight while the previous 32 columns are converted and stored.
__device__ __forceinline__ void acc_to_smem_operand(uint32_t taddr, int cw, uint32_t sdst /* + tid*16 already added */) {
    uint32_t va[32], vb[32];
    auto emit = [&](const uint32_t (&v)[32], int c, int nc) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            if (i * 8 < nc) {
                uint32_t w[4];
#pragma unroll
                for (int j = 0; j < 4; j++)
                    w[j] = pack_h2(__uint_as_float(v[i * 8 + 2 * j]), __uint_as_float(v[i * 8 + 2 * j + 1]));
                st_shared_v4(sdst + (uint32_t)((c / 8 + i) * kAkcBytes), w[0], w[1], w[2], w[3]);
            }
        }
    };
    tmem_ld_cols(taddr, cw, va);
#pragma unroll 1
    for (int c = 0; c < cw; c += 64) {
        tmem_wait_ld();
        if (c + 32 < cw) tmem_ld_cols(taddr + c + 32, cw - c - 32, vb);
        emit(va, c, cw - c);
        if (c + 32 < cw) {
            tmem_wait_ld();
            if (c + 64 < cw) tmem_ld_cols(taddr + c + 64, cw - c - 64, va);
            emit(vb, c + 32, cw - c - 32);
        }
    }
}

// fp32 Hacc columns [0, cw) at `taddr` -> relu -> packed fp16 pairs written back in place at columns [0, cw/2).
// Packed column c/2 + j is written only after fp32 columns [c, c+32) were read, and the load kept in flight reads
// columns >= c + 32 >= c/2 + 16, so the in-place compaction never overwrites unread data.
__device__ __forceinline__ void acc_to_tmem_operand(uint32_t taddr, int cw) {
    uint32_t va[32], vb[32];
    auto emit = [&](const uint32_t (&v)[32], int c, int nc) {
        uint32_t w[16];
#pragma unroll
        for (int j = 0; j < 16; j++) w[j] = pack_h2_relu(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
        __syncwarp();
        if (nc >= 32) tmem_st16(taddr + (c >> 1), w); else tmem_st8(taddr + (c >> 1), w);
    };
    tmem_ld_cols(taddr, cw, va);
#pragma unroll 1
    for (int c = 0; c < cw; c += 64) {
        tmem_wait_ld();
        if (c + 32 < cw) tmem_ld_cols(taddr + c + 32, cw - c - 32, vb);
        emit(va, c, cw - c);
        if (c + 32 < cw) {
            tmem_wait_ld();
            if (c + 64 < cw) tmem_ld_cols(taddr + c + 64, cw - c - 64, va);
            emit(vb, c + 32, cw - c - 32);
        }
    }
    tmem_wait_st();
}

// Debug event log (MlpParams::trace, CTA 0 only): role 0 = epilogue thread 0, 1 = MMA issuer, 2 = producer.
struct Tracer {
    unsigned long long* buf;
    int n;
    __device__ __forceinline__ void init(const MlpParams& p, int role, bool active) {
        buf = (p.trace && blockIdx.x == 0 && active) ? p.trace + role * QB_TRACE_EVENTS : nullptr;
        n = 0;
    }
    __device__ __forceinline__ void ev(uint32_t id) {
        if (buf && n < QB_TRACE_EVENTS) buf[n++] = ((unsigned long long)clock64() << 16) | id;
    }
};

struct RowCtx {
    bool valid;
    int64_t row;    // global row
    int64_t beam;   // (vector, beam) index b
    int code;
};

// 32 columns of the skip codeword C_m[code][d0 .. d0+32) (blocked table [D/8][K][8]); all loads issued together
__device__ __forceinline__ void load_cb64(const MlpParams& p, const RowCtx& rc, int d0, int cols, float4 (&cb)[8]) {
    const int K = p.plan.K;
    const float* base = p.cb_blk + ((size_t)(d0 >> 3) * K + rc.code) * 8 + (d0 & 7);   // d0 is a multiple of 16
#pragma unroll
    for (int i = 0; i < 4; i++) cb[i] = ldg4(base + (size_t)(i >> 1) * K * 8 + (i & 1) * 4);
    if (cols > 16) {
#pragma unroll
        for (int i = 4; i < 8; i++) cb[i] = ldg4(base + (size_t)(i >> 1) * K * 8 + (i & 1) * 4);
    }
}

// Final epilogue over accumulator columns [0, cw) at taddr that hold o[d0 .. d0+cw).  `cb` holds the prefetched skip
// codeword of the first 32 columns (when plan.skip); every 16-column slot of it is refilled with the columns 32 further
// on as soon as it has been used, and the accumulator is read 16 columns at a time with the next read in flight.
template <bool kScore>
__device__ __forceinline__ void consume_out(const MlpParams& p, const RowCtx& rc, uint32_t taddr, int cw, int d0,
                                            float4 (&cb)[8], float& acc) {
    const int D = p.plan.D;
    const float* src = (kScore ? p.r : p.xhat_in) + rc.beam * D + d0;
    uint32_t va[32], vb[32];
    auto emit = [&](const uint32_t (&v)[32], int cc, int h /* 16-column slot (0/1) of the skip-codeword buffer */) {
        float4 t[4];
#pragma unroll
        for (int i = 0; i < 4; i++) t[i] = (p.exp_flags & 8) ? make_float4(0.f, 0.f, 0.f, 0.f) : ldg4_jit(src + cc + i * 4);
        float o[16];
#pragma unroll
        for (int i = 0; i < 16; i++) o[i] = __uint_as_float(v[i]);
        if (p.plan.skip) {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const float4 cv = cb[4 * h + i];
                o[4 * i] += cv.x; o[4 * i + 1] += cv.y; o[4 * i + 2] += cv.z; o[4 * i + 3] += cv.w;
            }
        }
        if (kScore) {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const float e0 = t[i].x - o[4 * i], e1 = t[i].y - o[4 * i + 1], e2 = t[i].z - o[4 * i + 2], e3 = t[i].w - o[4 * i + 3];
                acc = fmaf(e0, e0, acc); acc = fmaf(e1, e1, acc); acc = fmaf(e2, e2, acc); acc = fmaf(e3, e3, acc);
            }
        } else if (rc.valid) {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int d = d0 + cc + i * 4;
                float4 out = make_float4(t[i].x + o[4 * i], t[i].y + o[4 * i + 1], t[i].z + o[4 * i + 2], t[i].w + o[4 * i + 3]);
                if (p.out_shift) {
                    const float4 sh = ldg4(p.out_shift + d);
                    out.x = fmaf(out.x, p.out_scale, sh.x); out.y = fmaf(out.y, p.out_scale, sh.y);
                    out.z = fmaf(out.z, p.out_scale, sh.z); out.w = fmaf(out.w, p.out_scale, sh.w);
                } else if (p.out_scale != 1.0f) {
                    out.x *= p.out_scale; out.y *= p.out_scale; out.z *= p.out_scale; out.w *= p.out_scale;
                }
                *reinterpret_cast<float4*>(p.xhat_out + rc.row * D + d) = out;
            }
        }
        if (p.plan.skip && cc + 32 < cw && !(p.exp_flags & 4)) {   // this slot's registers are free: fetch the same slot of the next 32 columns
            const int K = p.plan.K;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int d = d0 + cc + 32 + i * 4;
                cb[4 * h + i] = ldg4(p.cb_blk + ((size_t)(d >> 3) * K + rc.code) * 8 + (d & 7));
            }
        }
    };
    __syncwarp();
    tmem_ld16(taddr, va);
#pragma unroll 1
    for (int cc = 0; cc < cw; cc += 32) {
        tmem_wait_ld();
        if (cc + 16 < cw) { __syncwarp(); tmem_ld16(taddr + cc + 16, vb); }
        emit(va, cc, 0);
        if (cc + 16 < cw) {
            tmem_wait_ld();
            if (cc + 32 < cw) { __syncwarp(); tmem_ld16(taddr + cc + 32, va); }
            emit(vb, cc + 16, 1);
        }
    }
}

}  // namespace

template <bool kScore>
__global__ void __launch_bounds__(kThreads, 2) qb_mlp_kernel(const __grid_constant__ MlpParams p) {
    extern __shared__ __align__(1024) uint8_t dyn_smem[];
    __shared__ __align__(8) uint64_t bars[QB_BAR_COUNT];
    __shared__ __align__(8) uint64_t w_full[QB_MAX_STAGE];
    __shared__ __align__(8) uint64_t w_empty[QB_MAX_STAGE];
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const QbStepPlan& pl = p.plan;
    const int n_ops = pl.n_ops_block + pl.n_ops_out;
    const int64_t n_tiles = (p.n_rows + QB_TILE_M - 1) / QB_TILE_M;

    // ---- one-time setup ------------------------------------------------------------------------------------------
    if (tid == 0) {
        mbar_init(smem_u32(&bars[QB_BAR_NONE]), 1);
        mbar_init(smem_u32(&bars[QB_BAR_AE_READY]), kEpiThreads);
        mbar_init(smem_u32(&bars[QB_BAR_AH_READY]), kEpiThreads);
        mbar_init(smem_u32(&bars[QB_BAR_HACC_FREE]), kEpiThreads);
        mbar_init(smem_u32(&bars[QB_BAR_HACC_FULL]), 1);
        mbar_init(smem_u32(&bars[QB_BAR_EACC_FULL]), 1);
        for (int s = 0; s < QB_MAX_STAGE; s++) {
            mbar_init(smem_u32(&w_full[s]), 1);
            mbar_init(smem_u32(&w_empty[s]), 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 5) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                     "r"((uint32_t)pl.tmem_alloc_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t smem_base = smem_u32(dyn_smem);
    if (p.trace && tid == 0) {   // debug: which SM does this CTA run on (co-residency map)
        uint32_t smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        if (blockIdx.x < 512) p.trace[3 * QB_TRACE_EVENTS + blockIdx.x] = smid + 1;
    }
    if (p.stagger_cycles > 0) {
        // experiment: delay one of the two co-resident CTAs (stagger_cycles < 2^24: upper half of the grid; else odd CTAs)
        const bool late = (p.stagger_cycles >> 24) ? (blockIdx.x & 1) : (blockIdx.x >= (gridDim.x + 1) / 2);
        if (late) {
            const long long t0 = clock64();
            while (clock64() - t0 < (p.stagger_cycles & 0xffffff)) { }
        }
    }

    if (warp == 4) {
        // ======================================================================================= weight producer
        uint32_t stage = 0, phase = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            for (int l = 0; l <= pl.L; l++) {
                const int i0 = (l < pl.L) ? 0 : pl.n_ops_block;
                const int i1 = (l < pl.L) ? pl.n_ops_block : n_ops;
                const uint8_t* wbase = p.w_blob + ((l < pl.L) ? (size_t)l * (size_t)pl.block_w_bytes : 0);
                for (int i = i0; i < i1; i++) {
                    const uint32_t n_slab = p.ops[i].n_slab, slab_bytes = p.ops[i].slab_bytes, last_bytes = p.ops[i].last_bytes;
                    const uint8_t* src = wbase + p.ops[i].w_off;
                    for (uint32_t s = 0; s < n_slab; s++) {
                        const uint32_t bytes = (s + 1 == n_slab) ? last_bytes : slab_bytes;
                        mbar_wait(smem_u32(&w_empty[stage]), phase ^ 1, p.err_flag, 0x100 + stage);
                        if (elect_one()) {
                            mbar_expect_tx(smem_u32(&w_full[stage]), bytes);
                            bulk_g2s(smem_base + pl.smem_ring + stage * pl.slot_bytes, src + (size_t)s * slab_bytes, bytes,
                                     smem_u32(&w_full[stage]));
                        }
                        __syncwarp();
                        if (++stage == (uint32_t)pl.n_stage) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 5) {
        // ======================================================================================= MMA issuer
        // Kept lean on purpose: a lone warp retires a dependent instruction every ~4-6 cycles, so everything per slab
        // beyond "wait, 4 MMAs, commit" shows up as tensor-pipe idle time.
        uint32_t stage = 0, phase = 0;
        uint32_t par = 0;
        Tracer tr;
        tr.init(p, 1, (tid & 31) == 0);
        const uint32_t ae_lo = (smem_base + pl.smem_ae) >> 4;
        const uint32_t ring_lo = (smem_base + pl.smem_ring) >> 4;
        const uint32_t slot_lo = (uint32_t)pl.slot_bytes >> 4;
        const uint64_t a_hi = umma_desc(0, kAkcBytes, 128);
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            for (int l = 0; l <= pl.L; l++) {
                const int i0 = (l < pl.L) ? 0 : pl.n_ops_block;
                const int i1 = (l < pl.L) ? pl.n_ops_block : n_ops;
                for (int i = i0; i < i1; i++) {
                    const QbOp& op = p.ops[i];
                    if (op.wait_a) {
                        mbar_wait(smem_u32(&bars[op.wait_a]), (par >> op.wait_a) & 1, p.err_flag, 0x200 + op.wait_a);
                        par ^= 1u << op.wait_a;
                    }
                    if (op.wait_d) {
                        mbar_wait(smem_u32(&bars[op.wait_d]), (par >> op.wait_d) & 1, p.err_flag, 0x200 + op.wait_d);
                        par ^= 1u << op.wait_d;
                    }
                    tr.ev(0x200 + i);
                    const uint32_t n = op.n;
                    const uint64_t b_hi = umma_desc(0, n * 16u, 128);
                    const uint32_t b_step = 2u * n;                 // descriptor address units (16 B) per K=16
                    const uint32_t idesc = umma_idesc(n);
                    const uint32_t d_tmem = tmem_base + op.d_col;
                    const bool from_smem = op.a_src == QB_A_E;
                    uint32_t a_cur = from_smem ? ae_lo + (uint32_t)op.a_off * (kAkcBytes >> 4) : tmem_base + op.a_off;
                    uint32_t acc = op.accumulate;
                    int k_left = op.k_total;
                    const int ks = op.ks;
                    const uint32_t n_slab = op.n_slab;
                    for (uint32_t s = 0; s < n_slab; s++) {
                        const int nk = (k_left < ks ? k_left : ks) >> 4;
                        k_left -= ks;
                        mbar_wait(smem_u32(&w_full[stage]), phase, p.err_flag, 0x300 + stage);
                        tc_fence_after();
                        const uint32_t b_lo = ring_lo + stage * slot_lo;
                        if (elect_one()) {
                            if (from_smem) {
#pragma unroll 1
                                for (int t = 0; t < nk; t++) {
                                    tc_mma_ss(d_tmem, a_hi | (uint64_t)(a_cur + (uint32_t)t * ((2 * kAkcBytes) >> 4)),
                                              b_hi | (uint64_t)(b_lo + (uint32_t)t * b_step), idesc, acc);
                                    acc = 1;
                                }
                            } else {
#pragma unroll 1
                                for (int t = 0; t < nk; t++) {
                                    tc_mma_ts(d_tmem, a_cur + (uint32_t)t * 8u, b_hi | (uint64_t)(b_lo + (uint32_t)t * b_step), idesc, acc);
                                    acc = 1;
                                }
                            }
                            tc_commit(smem_u32(&w_empty[stage]));
                            if (s + 1 == n_slab && op.commit) tc_commit(smem_u32(&bars[op.commit]));
                        }
                        __syncwarp();
                        acc = 1;
                        a_cur += (uint32_t)nk * (from_smem ? ((2 * kAkcBytes) >> 4) : 8u);
                        if (++stage == (uint32_t)pl.n_stage) { stage = 0; phase ^= 1; }
                    }
                    tr.ev(0x400 + i);
                }
            }
        }
    } else {
        // ======================================================================================= epilogue warps
        const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
        const uint32_t ae_dst = smem_base + pl.smem_ae + (uint32_t)tid * 16u;
        uint32_t par = 0;
        Tracer tr;
        tr.init(p, 0, tid == 0);
        const int De = pl.De, K = pl.K, nkc = De >> 3;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            RowCtx rc;
            rc.row = tile * QB_TILE_M + tid;
            rc.valid = rc.row < p.n_rows;
            rc.beam = 0;
            rc.code = 0;
            if (rc.valid) {
                if (kScore) {
                    rc.beam = rc.row / p.C;
                    const int a = (int)(rc.row - rc.beam * p.C);
                    rc.code = p.A > 0 ? (int)__ldg(p.idx + rc.beam * p.A + a) : a;
                } else {
                    const int64_t v = rc.row / p.F_out;
                    const int parent = p.sel_parent ? (int)__ldg(p.sel_parent + rc.row) : 0;
                    rc.beam = v * p.F_in + parent;
                    rc.code = (int)__ldg(p.sel_code + rc.row * p.code_stride + p.code_off);
                }
                if (rc.code >= K) rc.code = K - 1;   // never read outside the tables (bad codes are rejected on the host)
            }
            tr.ev(1);
            // ---- init: e0 = T_m[code] + u_b; 64 columns per batch, all table loads of a batch issued up front ----------
            {
                const float* tp = p.t_blk + (size_t)rc.code * 8;
                const size_t tstride = (p.exp_flags & 2) ? 0 : (size_t)K * 8;
                const float* up = p.u + rc.beam * De;
                // the per-beam rows are shared by many rows of the tile: pull them into L1 (one 128 B line per thread)
                if (tid * 32 < De) prefetch_l1(up + tid * 32);
                auto emit_chunk = [&](int kc, const float4 t0, const float4 t1) {
                    float4 u0 = make_float4(0.f, 0.f, 0.f, 0.f), u1 = u0;
                    if (!(p.exp_flags & 1)) { u0 = ldg4_jit(up + kc * 8); u1 = ldg4_jit(up + kc * 8 + 4); }
                    uint32_t e[8];
                    const float f0 = t0.x + u0.x, f1 = t0.y + u0.y, f2 = t0.z + u0.z, f3 = t0.w + u0.w;
                    const float f4 = t1.x + u1.x, f5 = t1.y + u1.y, f6 = t1.z + u1.z, f7 = t1.w + u1.w;
                    e[0] = __float_as_uint(f0); e[1] = __float_as_uint(f1); e[2] = __float_as_uint(f2);
                    e[3] = __float_as_uint(f3); e[4] = __float_as_uint(f4); e[5] = __float_as_uint(f5);
                    e[6] = __float_as_uint(f6); e[7] = __float_as_uint(f7);
                    __syncwarp();
                    if (!(p.exp_flags & 32)) tmem_st8(lane_base + pl.tmem_e_col + kc * 8, e);
                    st_shared_v4(ae_dst + (uint32_t)kc * kAkcBytes, pack_h2(f0, f1), pack_h2(f2, f3), pack_h2(f4, f5),
                                 pack_h2(f6, f7));
                };
                int kc0 = 0;
                if (nkc >= 4) {                          // 32 columns per batch; the next batch's table rows in flight
                    float4 ta[8], tb[8];
#pragma unroll
                    for (int q = 0; q < 4; q++) { ta[2 * q] = ldg4(tp + (size_t)q * tstride); ta[2 * q + 1] = ldg4(tp + (size_t)q * tstride + 4); }
#pragma unroll 1
                    for (; kc0 + 4 <= nkc; kc0 += 8) {
                        const bool more1 = kc0 + 8 <= nkc;
                        if (more1) {
#pragma unroll
                            for (int q = 0; q < 4; q++) { tb[2 * q] = ldg4(tp + (size_t)(kc0 + 4 + q) * tstride); tb[2 * q + 1] = ldg4(tp + (size_t)(kc0 + 4 + q) * tstride + 4); }
                        }
#pragma unroll
                        for (int q = 0; q < 4; q++) emit_chunk(kc0 + q, ta[2 * q], ta[2 * q + 1]);
                        if (more1) {
                            if (kc0 + 12 <= nkc) {
#pragma unroll
                                for (int q = 0; q < 4; q++) { ta[2 * q] = ldg4(tp + (size_t)(kc0 + 8 + q) * tstride); ta[2 * q + 1] = ldg4(tp + (size_t)(kc0 + 8 + q) * tstride + 4); }
                            }
#pragma unroll
                            for (int q = 0; q < 4; q++) emit_chunk(kc0 + 4 + q, tb[2 * q], tb[2 * q + 1]);
                        }
                    }
                    kc0 = nkc & ~3;
                }
#pragma unroll 1
                for (; kc0 + 2 <= nkc; kc0 += 2) {      // tail (de is a multiple of 16): 16 columns at a time
                    const float4 a0 = ldg4(tp + (size_t)kc0 * K * 8), a1 = ldg4(tp + (size_t)kc0 * K * 8 + 4);
                    const float4 b0 = ldg4(tp + (size_t)(kc0 + 1) * K * 8), b1 = ldg4(tp + (size_t)(kc0 + 1) * K * 8 + 4);
                    emit_chunk(kc0, a0, a1);
                    emit_chunk(kc0 + 1, b0, b1);
                }
                tmem_wait_st();
                tc_fence_before();
                if (!(p.exp_flags & 16)) proxy_fence_async();
                mbar_arrive(smem_u32(&bars[QB_BAR_AE_READY]));
                tr.ev(2);
            }
            // ---- residual blocks -----------------------------------------------------------------------------------
            float4 cb[8];
#pragma unroll
            for (int i = 0; i < 8; i++) cb[i] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
            for (int l = 0; l < pl.L; l++) {
#pragma unroll 1
                for (int j = 0; j < pl.n_hchunk; j++) {
                    const int cw = min(pl.hc, pl.Dh - j * pl.hc);
                    mbar_wait(smem_u32(&bars[QB_BAR_HACC_FULL]), (par >> QB_BAR_HACC_FULL) & 1, p.err_flag, 0x404);
                    par ^= 1u << QB_BAR_HACC_FULL;
                    tc_fence_after();
                    tr.ev(3);
                    acc_to_tmem_operand(lane_base + pl.tmem_h_col, cw);
                    tc_fence_before();
                    mbar_arrive(smem_u32(&bars[QB_BAR_AH_READY]));
                    tr.ev(4);
                }
                const bool last = (l + 1 == pl.L);
                if (last) {
                    // inputs of the final epilogue travel while the last down-projection runs
                    const float* src = (kScore ? p.r : p.xhat_in) + rc.beam * pl.D;
                    if (tid * 32 < pl.D) prefetch_l1(src + tid * 32);
                    if (!pl.has_proj && pl.skip && !(p.exp_flags & 4)) load_cb64(p, rc, 0, pl.D, cb);
                }
                mbar_wait(smem_u32(&bars[QB_BAR_EACC_FULL]), (par >> QB_BAR_EACC_FULL) & 1, p.err_flag, 0x405);
                par ^= 1u << QB_BAR_EACC_FULL;
                tc_fence_after();
                tr.ev(5);
                if (!last || pl.has_proj) {
                    acc_to_smem_operand(lane_base + pl.tmem_e_col, De, ae_dst);
                    tc_fence_before();
                    if (!(p.exp_flags & 16)) proxy_fence_async();
                    mbar_arrive(smem_u32(&bars[QB_BAR_AE_READY]));
                    tr.ev(6);
                }
            }
            // ---- final epilogue --------------------------------------------------------------------------------------
            float acc = 0.f;
            if (pl.has_proj) {
                for (int q = 0; q < pl.n_ochunk; q++) {
                    const int cw = min(pl.oc, pl.D - q * pl.oc);
                    if (pl.skip) load_cb64(p, rc, q * pl.oc, cw, cb);
                    mbar_wait(smem_u32(&bars[QB_BAR_HACC_FULL]), (par >> QB_BAR_HACC_FULL) & 1, p.err_flag, 0x414);
                    par ^= 1u << QB_BAR_HACC_FULL;
                    tc_fence_after();
                    consume_out<kScore>(p, rc, lane_base + pl.tmem_h_col, cw, q * pl.oc, cb, acc);
                    tc_fence_before();
                    if (q + 1 < pl.n_ochunk) mbar_arrive(smem_u32(&bars[QB_BAR_HACC_FREE]));
                }
            } else {
                if (pl.L == 0 && pl.skip) load_cb64(p, rc, 0, pl.D, cb);
                consume_out<kScore>(p, rc, lane_base + pl.tmem_e_col, pl.D, 0, cb, acc);
                tc_fence_before();
            }
            if (kScore && rc.valid) p.dist[rc.row] = acc;
            tr.ev(8);
        }
    }

    // ---- teardown ------------------------------------------------------------------------------------------------
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 5) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                     "r"((uint32_t)pl.tmem_alloc_cols)
                     : "memory");
    }
}

cudaError_t mlp_set_smem_attr(int smem_bytes) {
    // the attribute belongs to the function, not to a model: only ever raise it (several models share the process)
    static int current[64] = {0};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    if (smem_bytes <= current[dev]) return cudaSuccess;
    e = cudaFuncSetAttribute(qb_mlp_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(qb_mlp_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e == cudaSuccess) current[dev] = smem_bytes;
    return e;
}

cudaError_t launch_mlp(const MlpParams& p, int n_sm, cudaStream_t stream) {
    if (p.n_rows <= 0) return cudaSuccess;
    const int64_t n_tiles = (p.n_rows + QB_TILE_M - 1) / QB_TILE_M;
    const int64_t slots = (int64_t)n_sm * p.plan.ctas_per_sm;
    const int grid = (int)(n_tiles < slots ? n_tiles : slots);
    // Debug: QB_MLP_TRACE=<file>[:<launch index>] dumps the event log of CTA 0 for one launch (synchronises).
    static const char* trace_env = getenv("QB_MLP_TRACE");
    static const int exp_flags = getenv("QB_EXP") ? atoi(getenv("QB_EXP")) : 0;
    if (exp_flags) const_cast<MlpParams&>(p).exp_flags = exp_flags;
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
        if (p.mode == QB_MODE_SCORE) qb_mlp_kernel<true><<<grid, kThreads, p.plan.smem_total, stream>>>(q);
        else qb_mlp_kernel<false><<<grid, kThreads, p.plan.smem_total, stream>>>(q);
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
    if (p.mode == QB_MODE_SCORE) qb_mlp_kernel<true><<<grid, kThreads, p.plan.smem_total, stream>>>(p);
    else qb_mlp_kernel<false><<<grid, kThreads, p.plan.smem_total, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace qb
