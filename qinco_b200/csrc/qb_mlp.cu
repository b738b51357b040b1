// Fused implicit-codebook MLP for one quantisation step, sm_100a (tcgen05 / TMEM / TMA bulk copy).
//
// Replaces, for a 128-row tile of candidates, the reference's chain of separate launches
// (reference qinco/model/qinco_base.py:262-280 = in_proj, QConcat :60-64, L x QBlockFFN :93-97, out_proj, skip;
// the distance of :343-345 and the x-hat update of :363-369) with ONE persistent kernel:
//
//   warps 0-3  epilogue: thread t owns row t of the tile == TMEM lane t.
//              init    e0 = T_m[code] + u_b  -> fp32 into the TMEM residual accumulator (tcgen05.st)
//                                             -> fp16 into the A_E operand tile in shared memory
//              H-epi   relu(Hacc) -> fp16 -> A_H operand tile            (per hidden chunk)
//              E-epi   Eacc -> fp16 -> A_E                               (per residual block)
//              final   score: dist = ||r_b - o||^2 (fp32)   apply: xhat_out = xhat_b + o
//   warp 4     producer: streams the pre-packed fp16 weight slabs with cp.async.bulk (TMA) into an mbarrier ring
//   warp 5     MMA issuer: one thread walks the op list (qb_plan.h) and issues tcgen05.mma (M=128, kind::f16,
//              fp32 accumulate in TMEM); completion is signalled with tcgen05.commit
//
// The fp32 residual stream never leaves TMEM; activations never touch HBM.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "qb_dev.h"

namespace qb {

namespace {

constexpr int kEpiThreads = 128;
constexpr int kThreads = 192;
constexpr int kMaxStage = 8;
constexpr uint32_t kTmemCols = 512;
constexpr int kAkcBytes = QB_TILE_M * 16;   // bytes of one 8-element k-chunk of an A operand tile

// ------------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

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
    if (mbar_try_wait(bar, parity)) return;
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
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        :
        : "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
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
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
                 "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
                 "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
                 "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
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

// fp32 accumulator columns [0, cw) at `taddr` -> fp16 (optionally relu) -> A operand k-chunks starting at `sdst`
template <bool kRelu>
__device__ __forceinline__ void acc_to_operand(uint32_t taddr, int cw, uint32_t sdst /* + tid*16 already added */) {
    uint32_t v[32];
    for (int c = 0; c < cw; c += 32) {
        const int nc = (cw - c >= 32) ? 32 : 16;
        __syncwarp();   // tcgen05.ld is .sync.aligned: reconverge after the per-thread mbarrier spin
        if (nc == 32) tmem_ld32(taddr + c, v); else tmem_ld16(taddr + c, v);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 4; i++) {
            if (i * 8 < nc) {
                uint32_t w[4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const float a = __uint_as_float(v[i * 8 + 2 * j]), b = __uint_as_float(v[i * 8 + 2 * j + 1]);
                    w[j] = kRelu ? pack_h2_relu(a, b) : pack_h2(a, b);
                }
                st_shared_v4(sdst + (uint32_t)((c / 8 + i) * kAkcBytes), w[0], w[1], w[2], w[3]);
            }
        }
    }
}

struct RowCtx {
    bool valid;
    int64_t row;    // global row
    int64_t beam;   // (vector, beam) index b
    int code;
};

// Final epilogue over accumulator columns [0, cw) that hold o[d0 .. d0+cw)
__device__ __forceinline__ void consume_out(const MlpParams& p, const RowCtx& rc, uint32_t taddr, int cw, int d0,
                                            float& acc) {
    uint32_t v[32];
    const int D = p.plan.D, K = p.plan.K;
    for (int c = 0; c < cw; c += 32) {
        const int nc = (cw - c >= 32) ? 32 : 16;
        __syncwarp();   // tcgen05.ld is .sync.aligned: reconverge first
        if (nc == 32) tmem_ld32(taddr + c, v); else tmem_ld16(taddr + c, v);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (rc.valid && i * 4 < nc) {
                const int d = d0 + c + i * 4;
                float o[4];
#pragma unroll
                for (int j = 0; j < 4; j++) o[j] = __uint_as_float(v[i * 4 + j]);
                if (p.plan.skip) {
                    const float4 cw4 =
                        __ldg(reinterpret_cast<const float4*>(p.cb_blk + ((size_t)(d >> 3) * K + rc.code) * 8 + (d & 7)));
                    o[0] += cw4.x; o[1] += cw4.y; o[2] += cw4.z; o[3] += cw4.w;
                }
                if (p.mode == QB_MODE_SCORE) {
                    const float4 t = __ldg(reinterpret_cast<const float4*>(p.r + rc.beam * D + d));
                    const float e0 = t.x - o[0], e1 = t.y - o[1], e2 = t.z - o[2], e3 = t.w - o[3];
                    acc = fmaf(e0, e0, acc); acc = fmaf(e1, e1, acc); acc = fmaf(e2, e2, acc); acc = fmaf(e3, e3, acc);
                } else {
                    const float4 xh = __ldg(reinterpret_cast<const float4*>(p.xhat_in + rc.beam * D + d));
                    float4 out = make_float4(xh.x + o[0], xh.y + o[1], xh.z + o[2], xh.w + o[3]);
                    if (p.out_shift) {
                        const float4 sh = __ldg(reinterpret_cast<const float4*>(p.out_shift + d));
                        out.x = fmaf(out.x, p.out_scale, sh.x); out.y = fmaf(out.y, p.out_scale, sh.y);
                        out.z = fmaf(out.z, p.out_scale, sh.z); out.w = fmaf(out.w, p.out_scale, sh.w);
                    } else if (p.out_scale != 1.0f) {
                        out.x *= p.out_scale; out.y *= p.out_scale; out.z *= p.out_scale; out.w *= p.out_scale;
                    }
                    *reinterpret_cast<float4*>(p.xhat_out + rc.row * D + d) = out;
                }
            }
        }
    }
}

}  // namespace

__global__ void __launch_bounds__(kThreads, 1) qb_mlp_kernel(const __grid_constant__ MlpParams p) {
    extern __shared__ __align__(1024) uint8_t dyn_smem[];
    __shared__ __align__(16) QbOp ops_s[QB_MAX_OPS];
    __shared__ __align__(8) uint64_t bars[QB_BAR_COUNT];
    __shared__ __align__(8) uint64_t w_full[kMaxStage];
    __shared__ __align__(8) uint64_t w_empty[kMaxStage];
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const QbStepPlan& pl = p.plan;
    const int n_ops = pl.n_ops_block + pl.n_ops_out;
    const int64_t n_tiles = (p.n_rows + QB_TILE_M - 1) / QB_TILE_M;

    // ---- one-time setup ------------------------------------------------------------------------------------------
    for (int i = tid; i < n_ops * (int)(sizeof(QbOp) / 16); i += kThreads)
        reinterpret_cast<uint4*>(ops_s)[i] = __ldg(reinterpret_cast<const uint4*>(p.ops) + i);
    if (tid == 0) {
        mbar_init(smem_u32(&bars[QB_BAR_AE_READY]), kEpiThreads);
        mbar_init(smem_u32(&bars[QB_BAR_AH0_READY]), kEpiThreads);
        mbar_init(smem_u32(&bars[QB_BAR_AH1_READY]), kEpiThreads);
        mbar_init(smem_u32(&bars[QB_BAR_HACC0_FREE]), kEpiThreads);
        mbar_init(smem_u32(&bars[QB_BAR_HACC1_FREE]), kEpiThreads);
        mbar_init(smem_u32(&bars[QB_BAR_HACC0_FULL]), 1);
        mbar_init(smem_u32(&bars[QB_BAR_HACC1_FULL]), 1);
        mbar_init(smem_u32(&bars[QB_BAR_EACC_FULL]), 1);
        mbar_init(smem_u32(&bars[QB_BAR_NONE]), 1);
        for (int s = 0; s < kMaxStage; s++) {
            mbar_init(smem_u32(&w_full[s]), 1);
            mbar_init(smem_u32(&w_empty[s]), 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 5) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                     "r"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t smem_base = smem_u32(dyn_smem);

    if (warp == 4) {
        // ======================================================================================= weight producer
        if ((tid & 31) == 0) {
            uint32_t stage = 0, phase = 0;
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                for (int l = 0; l <= pl.L; l++) {
                    const int i0 = (l < pl.L) ? 0 : pl.n_ops_block;
                    const int i1 = (l < pl.L) ? pl.n_ops_block : n_ops;
                    const uint8_t* wbase = p.w_blob + ((l < pl.L) ? (size_t)l * (size_t)pl.block_w_bytes : 0);
                    for (int i = i0; i < i1; i++) {
                        const uint32_t w_off = ops_s[i].w_off, w_bytes = ops_s[i].w_bytes;
                        mbar_wait(smem_u32(&w_empty[stage]), phase ^ 1, p.err_flag, 0x100 + stage);
                        mbar_expect_tx(smem_u32(&w_full[stage]), w_bytes);
                        bulk_g2s(smem_base + pl.smem_ring + stage * pl.slot_bytes, wbase + w_off, w_bytes,
                                 smem_u32(&w_full[stage]));
                        if (++stage == (uint32_t)pl.n_stage) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 5) {
        // ======================================================================================= MMA issuer
        if ((tid & 31) == 0) {
            uint32_t stage = 0, phase = 0;
            uint32_t par = (1u << QB_BAR_HACC0_FREE) | (1u << QB_BAR_HACC1_FREE);   // "free" barriers pass first time
            const uint32_t a_base[3] = {smem_base + pl.smem_ae, smem_base + pl.smem_ah[0], smem_base + pl.smem_ah[1]};
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                for (int l = 0; l <= pl.L; l++) {
                    const int i0 = (l < pl.L) ? 0 : pl.n_ops_block;
                    const int i1 = (l < pl.L) ? pl.n_ops_block : n_ops;
                    for (int i = i0; i < i1; i++) {
                        const QbOp op = ops_s[i];
                        if (op.wait_a) {
                            mbar_wait(smem_u32(&bars[op.wait_a]), (par >> op.wait_a) & 1, p.err_flag, 0x200 + op.wait_a);
                            par ^= 1u << op.wait_a;
                        }
                        if (op.wait_d) {
                            mbar_wait(smem_u32(&bars[op.wait_d]), (par >> op.wait_d) & 1, p.err_flag, 0x200 + op.wait_d);
                            par ^= 1u << op.wait_d;
                        }
                        mbar_wait(smem_u32(&w_full[stage]), phase, p.err_flag, 0x300 + stage);
                        tc_fence_after();
                        const uint32_t b_addr = smem_base + pl.smem_ring + stage * pl.slot_bytes;
                        const uint32_t b_lbo = (uint32_t)op.n * 16u;
                        const uint32_t idesc = umma_idesc(op.n);
                        const uint32_t d_tmem = tmem_base + op.d_col;
                        const int nk = op.k >> 4;
                        for (int t = 0; t < nk; t++) {
                            const uint64_t ad = umma_desc(a_base[op.a_buf] + (uint32_t)(op.a_kc + 2 * t) * kAkcBytes,
                                                          kAkcBytes, 128);
                            const uint64_t bd = umma_desc(b_addr + (uint32_t)(2 * t) * b_lbo, b_lbo, 128);
                            tc_mma_f16(d_tmem, ad, bd, idesc, (op.accumulate || t > 0) ? 1u : 0u);
                        }
                        tc_commit(smem_u32(&w_empty[stage]));
                        if (op.commit) tc_commit(smem_u32(&bars[op.commit]));
                        if (++stage == (uint32_t)pl.n_stage) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else {
        // ======================================================================================= epilogue warps
        const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
        const uint32_t my16 = (uint32_t)tid * 16u;
        const uint32_t ae_dst = smem_base + pl.smem_ae + my16;
        const uint32_t ah_dst[2] = {smem_base + pl.smem_ah[0] + my16, smem_base + pl.smem_ah[1] + my16};
        uint32_t par = 0;
        const int De = pl.De, K = pl.K;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            RowCtx rc;
            rc.row = tile * QB_TILE_M + tid;
            rc.valid = rc.row < p.n_rows;
            rc.beam = 0;
            rc.code = 0;
            if (rc.valid) {
                if (p.mode == QB_MODE_SCORE) {
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
            // ---- init: e0 = T_m[code] + u_b ----------------------------------------------------------------------
            {
                const float* tp = p.t_blk + (size_t)rc.code * 8;
                const float* up = p.u + rc.beam * De;
                __syncwarp();
                for (int kc = 0; kc < De / 8; kc++) {
                    const float4 t0 = __ldg(reinterpret_cast<const float4*>(tp + (size_t)kc * K * 8));
                    const float4 t1 = __ldg(reinterpret_cast<const float4*>(tp + (size_t)kc * K * 8 + 4));
                    const float4 u0 = __ldg(reinterpret_cast<const float4*>(up + kc * 8));
                    const float4 u1 = __ldg(reinterpret_cast<const float4*>(up + kc * 8 + 4));
                    const float e[8] = {t0.x + u0.x, t0.y + u0.y, t0.z + u0.z, t0.w + u0.w,
                                        t1.x + u1.x, t1.y + u1.y, t1.z + u1.z, t1.w + u1.w};
                    tmem_st8(lane_base + pl.tmem_e_col + kc * 8, e);
                    st_shared_v4(ae_dst + (uint32_t)kc * kAkcBytes, pack_h2(e[0], e[1]), pack_h2(e[2], e[3]),
                                 pack_h2(e[4], e[5]), pack_h2(e[6], e[7]));
                }
                tmem_wait_st();
                tc_fence_before();
                proxy_fence_async();
                mbar_arrive(smem_u32(&bars[QB_BAR_AE_READY]));
            }
            // ---- residual blocks -----------------------------------------------------------------------------------
            for (int l = 0; l < pl.L; l++) {
                for (int j = 0; j < pl.n_hchunk; j++) {
                    const int buf = j % pl.n_hbuf;
                    const int cw = min(pl.hc, pl.Dh - j * pl.hc);
                    const int full = QB_BAR_HACC0_FULL + buf;
                    mbar_wait(smem_u32(&bars[full]), (par >> full) & 1, p.err_flag, 0x400 + full);
                    par ^= 1u << full;
                    tc_fence_after();
                    acc_to_operand<true>(lane_base + pl.tmem_h_col[buf], cw, ah_dst[buf]);
                    tc_fence_before();
                    proxy_fence_async();
                    mbar_arrive(smem_u32(&bars[QB_BAR_AH0_READY + buf]));
                    mbar_arrive(smem_u32(&bars[QB_BAR_HACC0_FREE + buf]));
                }
                mbar_wait(smem_u32(&bars[QB_BAR_EACC_FULL]), (par >> QB_BAR_EACC_FULL) & 1, p.err_flag, 0x408);
                par ^= 1u << QB_BAR_EACC_FULL;
                tc_fence_after();
                if (l + 1 < pl.L || pl.has_proj) {
                    acc_to_operand<false>(lane_base + pl.tmem_e_col, De, ae_dst);
                    tc_fence_before();
                    proxy_fence_async();
                    mbar_arrive(smem_u32(&bars[QB_BAR_AE_READY]));
                }
            }
            // ---- final epilogue --------------------------------------------------------------------------------------
            float acc = 0.f;
            if (pl.has_proj) {
                for (int q = 0; q < pl.n_ochunk; q++) {
                    const int buf = q % pl.n_hbuf;
                    const int cw = min(pl.oc, pl.D - q * pl.oc);
                    const int full = QB_BAR_HACC0_FULL + buf;
                    mbar_wait(smem_u32(&bars[full]), (par >> full) & 1, p.err_flag, 0x400 + full);
                    par ^= 1u << full;
                    tc_fence_after();
                    consume_out(p, rc, lane_base + pl.tmem_h_col[buf], cw, q * pl.oc, acc);
                    tc_fence_before();
                    mbar_arrive(smem_u32(&bars[QB_BAR_HACC0_FREE + buf]));
                }
            } else {
                consume_out(p, rc, lane_base + pl.tmem_e_col, pl.D, 0, acc);
            }
            if (p.mode == QB_MODE_SCORE && rc.valid) p.dist[rc.row] = acc;
        }
    }

    // ---- teardown ------------------------------------------------------------------------------------------------
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 5) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

cudaError_t mlp_set_smem_attr(int smem_bytes) {
    return cudaFuncSetAttribute(qb_mlp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
}

cudaError_t launch_mlp(const MlpParams& p, int n_sm, cudaStream_t stream) {
    if (p.n_rows <= 0) return cudaSuccess;
    const int64_t n_tiles = (p.n_rows + QB_TILE_M - 1) / QB_TILE_M;
    const int grid = (int)(n_tiles < n_sm ? n_tiles : n_sm);
    qb_mlp_kernel<<<grid, kThreads, p.plan.smem_total, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace qb
