// IVF first step on the tensor core, sm_100a (tcgen05 / TMEM / TMA bulk copy).
//
// Replaces IVFBook.quantize / encode (reference qinco/model/qinco_base.py:146-174): per vector the arg-min over ivf_K (up to
// 2^20) centroids of  (|x|^2 + |c|^2) - 2 x.c  (the reference's approx_pairwise_distance, qinco/utils.py:336-346), ties to
// the lower index, then the centroid lookup that seeds the single starting beam.  At 2^20 centroids this is 268 MFLOP per
// vector at d = 128; the fp32 CUDA-core kernel (qb_ivf_assign_kernel, ~16-20 TFLOP/s) would cap an IVF model near 70 k vectors/s
// whatever the MLP does.  Here a CTA holds 128 vectors' operand [x_hi | x_lo] (fp16 hi/lo split of the fp32 value) in shared
// memory for the whole launch and streams the centroids in parts of 128, pre-packed as [c_hi | c_lo | |c|^2]:
//     x . c ~= x_hi . c_hi + x_lo . c_hi + x_hi . c_lo         three tcgen05.mma groups, fp32 accumulate in TMEM
// which is accurate to ~2^-22 relative (fp32 level), so the arg-min only differs from an fp32 evaluation on rounding ties.
//
//   warps 0-7  epilogue: TMEM -> d = (|x|^2 + |c|^2) - 2 g, running (min, arg-min) in registers; thread = (row, half of a part)
//   warp 8     one lane streams the parts (cp.async.bulk: centroids into a 2-slot ring that is recycled as soon as the MMAs
//              reading it have completed, norms into a 4-slot ring) and issues the MMAs into a 4-slot accumulator ring (4 x 128
//              TMEM columns), so loads, MMAs and epilogues of neighbouring parts overlap
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

#include "qb_dev.h"
#include "qb_tc_util.h"

namespace qb {

namespace {

using namespace tc;

constexpr int kEpi = 256, kThreads = kEpi + 64;      // + the MMA-issuer warp and the loader warp
constexpr int kRows = 128, kNp = QB_IVF_NP;      // vectors per CTA, centroids per part
constexpr int kAkc = kRows * 16;

__global__ void __launch_bounds__(kThreads, 1) qb_ivf_tc_kernel(const __grid_constant__ IvfTcParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    constexpr int kAcc = 4;                                                   // accumulator / norm ring depth
    __shared__ __align__(8) uint64_t b_full[2], acc_full[kAcc], acc_free[kAcc];
    // norms of a part: written when the part is streamed (up to kAcc + 1 parts ahead of its epilogue), so twice as deep
    __shared__ __align__(16) float cn_ring[2 * kAcc][kNp];
    __shared__ uint32_t tmem_base_s;
    __shared__ float xn_part[2][kRows];
    __shared__ float best_d[2][kRows];
    __shared__ int best_k[2][kRows];
    const int tid = threadIdx.x, warp = tid >> 5;
    const int D = p.D;
    const int64_t v0 = (int64_t)blockIdx.x * kRows;
    const int nrow = (int)min((int64_t)kRows, p.n - v0);
    const uint32_t sbase = smem_u32(smem);
    const uint32_t a_bytes = (uint32_t)(2 * (D / 8)) * kAkc;                 // [x_hi | x_lo]
    const uint32_t part_bytes = (uint32_t)kNp * (uint32_t)D * 4u + (uint32_t)kNp * 4u;     // c_hi, c_lo, |c|^2
    const uint32_t cent_bytes = (uint32_t)kNp * (uint32_t)D * 4u;             // c_hi + c_lo of a part
    const uint32_t slot_bytes = (cent_bytes + 1023u) & ~1023u;
    const uint32_t b_base = sbase + ((a_bytes + 1023u) & ~1023u);
    const int n_parts = (p.ivf_K + kNp - 1) / kNp;

    if (tid == 0) {
        for (int i = 0; i < 2; i++) mbar_init(smem_u32(&b_full[i]), 1);
        for (int i = 0; i < kAcc; i++) {
            mbar_init(smem_u32(&acc_full[i]), 1);
            mbar_init(smem_u32(&acc_free[i]), kEpi / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // ---- A operand: the tile's normalised vectors as fp16 hi / lo (thread = row, half of the dimensions), and |x|^2
    float xn2 = 0.f;
    if (tid < kEpi) {
        const int row = tid & (kRows - 1), half = tid >> 7;
        const bool live = row < nrow;
        const float* xrow = p.x + (v0 + row) * D;
        const int cw = D >> 1, c_lo = half * cw;                // D is a multiple of 16
        for (int j = 0; j < cw; j += 8) {
            float val[8];
#pragma unroll
            for (int q = 0; q < 8; q++) val[q] = 0.f;
            if (live) {
                const float4 x0 = *reinterpret_cast<const float4*>(xrow + c_lo + j), x1 = *reinterpret_cast<const float4*>(xrow + c_lo + j + 4);
                val[0] = x0.x; val[1] = x0.y; val[2] = x0.z; val[3] = x0.w; val[4] = x1.x; val[5] = x1.y; val[6] = x1.z; val[7] = x1.w;
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    if (p.mean) val[q] -= __ldg(p.mean + c_lo + j + q);
                    val[q] /= p.std_div;                          // the reference divides: (x - mean) / std (qinco_base.py:533)
                    xn2 = fmaf(val[q], val[q], xn2);
                }
            }
            const uint32_t kc = (uint32_t)((c_lo + j) >> 3);
            const uint32_t dst = sbase + kc * kAkc + (uint32_t)row * 16u;
            put_hi_lo(dst, dst + (uint32_t)(D / 8) * kAkc, val);
        }
        xn_part[half][row] = xn2;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == kEpi / 32 + 1) {
        // ============================================================================================ loader (one lane)
        // part i goes into centroid slot i & 1 as soon as the MMAs of part i - 2 (the slot's previous reader) have completed;
        // the norms have their own, deeper ring (a slot is rewritten kAcc + 1 parts after the epilogue consumed it at the earliest)
        if ((tid & 31) == 0) {
            for (int i = 0; i < n_parts; i++) {
                const int buf = i & 1;
                if (i >= 2) mbar_wait(smem_u32(&acc_full[(i - 2) % kAcc]), (uint32_t)(((i - 2) / kAcc) & 1), p.err_flag, 0xa02);
                // (norm slot i % (2 kAcc) was read by the epilogue of part i - 2 kAcc, which is through: the MMAs of part i - 2 could
                //  only be issued after the epilogue of part i - 2 - kAcc had released their accumulator)
                const uint32_t dst = b_base + (uint32_t)buf * slot_bytes, full = smem_u32(&b_full[buf]);
                mbar_expect_tx(full, part_bytes);
                const uint8_t* src = p.cent_pack + (size_t)i * part_bytes;
                for (uint32_t o = 0; o < cent_bytes; o += 32768u) bulk_g2s(dst + o, src + o, min(32768u, cent_bytes - o), full);
                bulk_g2s(smem_u32(&cn_ring[i % (2 * kAcc)][0]), src + cent_bytes, (uint32_t)kNp * 4u, full);
            }
        }
    } else if (warp == kEpi / 32) {
        // ======================================================================================== MMA issuer (one lane)
        if ((tid & 31) == 0) {
            const uint32_t idesc = (1u << 4) | ((uint32_t)(kNp >> 3) << 17) | ((uint32_t)(kRows >> 4) << 24);
            const uint32_t a_hi = (((uint32_t)kAkc >> 4) << 16) | ((sbase >> 4) & 0x3FFFu);
            const uint32_t a_lo = (((uint32_t)kAkc >> 4) << 16) | (((sbase + (uint32_t)(D / 8) * kAkc) >> 4) & 0x3FFFu);
            const uint32_t a_step = (2u * kAkc) >> 4, w_step = 2u * (uint32_t)kNp;
            const int k16 = D >> 4;
            for (int i = 0; i < n_parts; i++) {
                const int buf = i & 1, ab = i % kAcc;
                mbar_wait(smem_u32(&b_full[buf]), (uint32_t)((i >> 1) & 1), p.err_flag, 0xa00);
                // accumulator slot ab was last used by part i - kAcc: its epilogue must be through
                if (i >= kAcc) mbar_wait(smem_u32(&acc_free[ab]), (uint32_t)(((i - kAcc) / kAcc) & 1), p.err_flag, 0xa01);
                tc_fence_after();
                const uint32_t bdst = b_base + (uint32_t)buf * slot_bytes;
                const uint32_t w_hi = ((uint32_t)kNp << 16) | ((bdst >> 4) & 0x3FFFu);
                const uint32_t w_lo = ((uint32_t)kNp << 16) | (((bdst + (uint32_t)kNp * (uint32_t)D * 2u) >> 4) & 0x3FFFu);
                const uint32_t d_tmem = tmem_base + (uint32_t)(ab * kNp);
                uint32_t acc = 0u;
#pragma unroll
                for (int g = 0; g < 3; g++) {
                    uint32_t a = g == 1 ? a_lo : a_hi, w = g == 2 ? w_lo : w_hi;
                    if (k16 == 8) {                      // d = 128: fully unrolled
#pragma unroll
                        for (int k = 0; k < 8; k++) { mma_f16_step(d_tmem, a, w, idesc, acc, a_step, w_step); acc = 1u; }
                    } else {
#pragma unroll 1
                        for (int k = 0; k < k16; k++) { mma_f16_step(d_tmem, a, w, idesc, acc, a_step, w_step); acc = 1u; }
                    }
                }
                tc_commit(smem_u32(&acc_full[ab]));
            }
        }
    } else {
        // ================================================================================================ epilogue warps
        const int row = tid & (kRows - 1), half = tid >> 7;
        const float xn = xn_part[0][row] + xn_part[1][row];
        const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        float bd = FLT_MAX;
        int bk = 0x7fffffff;
        for (int i = 0; i < n_parts; i++) {
            const int ab = i % kAcc;
            mbar_wait(smem_u32(&acc_full[ab]), (uint32_t)((i / kAcc) & 1), p.err_flag, 0xa10);
            tc_fence_after();
            const float* cn = &cn_ring[i % (2 * kAcc)][half * (kNp / 2)];
            const int k0 = i * kNp + half * (kNp / 2);
#pragma unroll
            for (int c = 0; c < kNp / 2; c += 16) {
                uint32_t v[16];
                __syncwarp();
                tmem_ld16(lane_addr + (uint32_t)(ab * kNp + half * (kNp / 2) + c), v);
                tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    const float d = (xn + cn[c + j]) - 2.f * __uint_as_float(v[j]);      // utils.py:346
                    if (d < bd) { bd = d; bk = k0 + c + j; }      // increasing index, strict <: ties keep the lower index
                }
            }
            tc_fence_before();
            __syncwarp();
            if ((tid & 31) == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&acc_free[ab])) : "memory");
        }
        best_d[half][row] = bd;
        best_k[half][row] = bk;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    // ---- the two halves of a row meet; code + centroid lookup (the single starting beam)
    if (tid < kRows) {
        float d = best_d[0][tid];
        int k = best_k[0][tid];
        const float d1 = best_d[1][tid];
        const int k1 = best_k[1][tid];
        if (d1 < d || (d1 == d && k1 < k)) { d = d1; k = k1; }
        if (k >= p.ivf_K || k < 0) k = 0;       // all-NaN row: stay in range
        best_k[0][tid] = k;
        if (tid < nrow) p.codes_out[v0 + tid] = k;
    }
    __syncthreads();
    const int d4n = D >> 2;
    for (int t = tid; t < nrow * d4n; t += kThreads) {
        const int v = t / d4n, d4 = t - v * d4n;
        *reinterpret_cast<float4*>(p.xhat_out + (v0 + v) * D + d4 * 4) = __ldg(reinterpret_cast<const float4*>(p.cent + (size_t)best_k[0][v] * D + d4 * 4));
    }
}

}  // namespace

size_t ivf_tc_smem_bytes(int D) {
    const size_t a = ((size_t)2 * (D / 8) * kAkc + 1023) & ~(size_t)1023;
    const size_t slot = ((size_t)kNp * D * 4 + 1023) & ~(size_t)1023;
    return a + 2 * slot;
}

cudaError_t launch_ivf_tc(const IvfTcParams& p, cudaStream_t stream) {
    if (p.n <= 0) return cudaSuccess;
    if (p.D % 16 || p.D > QB_IVF_TC_MAX_D || p.ivf_K < 1) return cudaErrorInvalidValue;
    const size_t smem = ivf_tc_smem_bytes(p.D);
    static size_t attr_dev[64] = {0};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    if (smem > attr_dev[dev]) {
        e = cudaFuncSetAttribute(qb_ivf_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_dev[dev] = smem;
    }
    const int64_t grid = (p.n + kRows - 1) / kRows;
    qb_ivf_tc_kernel<<<(unsigned)grid, kThreads, smem, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace qb
