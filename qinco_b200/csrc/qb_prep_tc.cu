// Beam preparation on the tensor core, sm_100a (tcgen05 / TMEM / TMA bulk copy).
//
// Per beam row b of step m the encode loop needs (reference QincoSubstep.get_distances_for_codes / select_code_candidates,
// qinco/model/qinco_base.py:114-121; the hoisted half of QConcat, :60-64; distances qinco/utils.py:336-346):
//     r_b  = x_n - xhat_b                                      (fp32, written for the score launch)
//     u_b  = Wcat[:, De:] . xhat_b                             [De]
//     idx_b = the A smallest of  (|r_b|^2 + |S_k|^2) - 2 r_b . S_k   over the K pre-selection codewords (A > 0)
// The two contractions are GEMMs with M = beams, K-dim = D and N = De / K.  The CUDA-core version (qb_prep_kernel) ran them
// at ~20 TFLOP/s and was 41 % of a QINCo2-S A=16 step; here a CTA takes a tile of 128 beams and runs them as tcgen05.mma with
// fp32-level accuracy: both operands are split into fp16 hi + lo parts and three products are accumulated in TMEM
//     a . w  ~=  a_hi . w_hi + a_lo . w_hi + a_hi . w_lo          (error ~2^-22 relative, the dropped lo . lo term)
// so the ranking sees the same numbers as an fp32 evaluation up to rounding-level ties.
//
// Dataflow of one CTA (256 threads, no warp specialisation: the kernel is <2 % of a step, what matters is that it is not 40 %):
//   for pass in {u, distances}:  for every 64-wide chunk of D:
//       all threads   build the A operand chunk [a_hi | a_lo] in shared memory from xhat_b / r_b (K-major core matrices)
//       thread 0      streams the pre-packed weight parts [w_hi | w_lo] (<= 256 rows each) with cp.async.bulk into a
//                     two-deep ring and issues the three MMA groups per part; accumulators: TMEM columns [0, N)
//   epilogue u:       TMEM -> u_b (global, fp32)
//   epilogue dist:    TMEM -> d = (|r|^2 + |S_k|^2) - 2 g -> shared memory -> top-A per row (a warp ranks four rows at a time,
//                     ties to the lower index like torch.topk)
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

#include "qb_dev.h"
#include "qb_tc_util.h"

namespace qb {

namespace {

constexpr int kWorkers = 256, kThreads = kWorkers + 64;      // + the MMA-issuer warp and the loader warp
constexpr int kRows = 128;                  // beams per CTA = MMA M
constexpr int kDc = QB_PREP_DC;             // D chunk
constexpr int kNp = QB_PREP_NP;             // weight rows per part = MMA N (<= 256)
constexpr int kAkc = kRows * 16;            // bytes of one 8-element k-chunk of the A operand
constexpr int kAHalf = (kDc / 8) * kAkc;    // a_hi (or a_lo) of one chunk: 16 KB
constexpr int kBHalf = kNp * kDc * 2;       // w_hi (or w_lo) of one part and chunk: 32 KB
constexpr int kSmemA = 0, kSmemB = 2 * 2 * kAHalf, kSmemTotal = kSmemB + 2 * 2 * kBHalf;      // two A buffers (64 KB) + two ring slots (128 KB)

using namespace tc;

// ---- top-16 of a stream of 16-value batches, exact: keys are (order-preserving bits of the fp32 distance) << 32 | index, so
// the unsigned order is the (distance, index) order torch.topk(largest=False) produces.  Everything is a fixed network on
// registers: no shared memory, no divergence.
__device__ __forceinline__ unsigned long long dist_key(float d, int k) {
    const uint32_t b = __float_as_uint(d);
    const uint32_t ord = b ^ ((uint32_t)((int32_t)b >> 31) | 0x80000000u);      // negative floats (rounding) order below positive ones
    return ((unsigned long long)ord << 32) | (uint32_t)k;
}
__device__ __forceinline__ void cex(unsigned long long& a, unsigned long long& b) {      // a <= b afterwards
    const bool sw = a > b;
    const unsigned long long lo = sw ? b : a, hi = sw ? a : b;
    a = lo; b = hi;
}
// bitonic merge of a bitonic 16-sequence into ascending order (4 stages of 8 compare-exchanges)
__device__ __forceinline__ void bitonic_merge16(unsigned long long (&v)[16]) {
#pragma unroll
    for (int j = 8; j > 0; j >>= 1) {
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const int l = i ^ j;
            if (l > i) cex(v[i], v[l]);
        }
    }
}
// full bitonic sort of 16 keys, ascending (10 stages)
__device__ __forceinline__ void bitonic_sort16(unsigned long long (&v)[16]) {
#pragma unroll
    for (int k = 2; k <= 16; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
            for (int i = 0; i < 16; i++) {
                const int l = i ^ j;
                if (l > i) {
                    if ((i & k) == 0) cex(v[i], v[l]); else cex(v[l], v[i]);
                }
            }
        }
    }
}
// best <- the 16 smallest of best U other (both ascending): min(best[i], other[15 - i]) is bitonic and holds them
__device__ __forceinline__ void keep_smallest16(unsigned long long (&best)[16], const unsigned long long (&other)[16]) {
#pragma unroll
    for (int i = 0; i < 16; i++) best[i] = best[i] < other[15 - i] ? best[i] : other[15 - i];
    bitonic_merge16(best);
}

__device__ __forceinline__ void bar_workers() { asm volatile("bar.sync 1, %0;" ::"n"(kWorkers) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }

__global__ void __launch_bounds__(kThreads, 1) qb_prep_tc_kernel(const __grid_constant__ PrepTcParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    // loader -> issuer (weight part landed) / issuer -> loader (MMAs reading the slot done); workers -> issuer (A chunk built)
    // / issuer -> workers (MMAs reading the A buffer done); issuer -> workers (all MMAs of a pass done); workers -> issuer
    // (u epilogue has read the accumulators the distance pass is about to overwrite)
    __shared__ __align__(8) uint64_t b_full[2], mma_done[2], a_full[2], a_free[2], pass_done[2], u_read;
    __shared__ uint32_t tmem_base_s;
    __shared__ float rn_part[2][kRows];
    __shared__ uint8_t code_s[kRows][16];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int D = p.D, De = p.De, K = p.K;
    const int64_t b0 = (int64_t)blockIdx.x * kRows;
    const int nrow = (int)min((int64_t)kRows, p.n_beams - b0);
    const uint32_t sbase = smem_u32(smem);
    const int n_chunks = (D + kDc - 1) / kDc;
    // accumulators: u in TMEM columns [0, De); the distances next to them when both fit, else in the same columns (then the
    // distance MMAs wait for the u epilogue)
    const int goff = (p.wx_pack && De + p.K16 <= 512) ? De : 0;
    const bool g_waits_u = p.wx_pack && p.sub_pack && goff == 0;

    if (tid == 0) {
        for (int i = 0; i < 2; i++) {
            mbar_init(smem_u32(&b_full[i]), 1); mbar_init(smem_u32(&mma_done[i]), 1);
            mbar_init(smem_u32(&a_full[i]), kWorkers / 32); mbar_init(smem_u32(&a_free[i]), 1); mbar_init(smem_u32(&pass_done[i]), 1);
        }
        mbar_init(smem_u32(&u_read), kWorkers / 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == kWorkers / 32 + 1) {
        // ============================================================================================ loader (one lane)
        if (lane == 0) {
            int it = 0;
            for (int pass = 0; pass < 2; pass++) {
                const uint8_t* pack = pass == 0 ? p.wx_pack : p.sub_pack;
                if (!pack) continue;
                const int N = pass == 0 ? De : p.K16, n_parts = (N + kNp - 1) / kNp;
                size_t off = 0;
                for (int c = 0; c < n_chunks; c++) {
                    const int dc = min(kDc, D - c * kDc);
                    for (int part = 0; part < n_parts; part++, it++) {
                        const int np = min(kNp, N - part * kNp), buf = it & 1;
                        const uint32_t bytes = 2u * (uint32_t)np * (uint32_t)dc * 2u;          // w_hi + w_lo
                        if (it >= 2) mbar_wait(smem_u32(&mma_done[buf]), (uint32_t)(((it - 2) >> 1) & 1), p.err_flag, 0x901);
                        const uint32_t full = smem_u32(&b_full[buf]), dst = sbase + kSmemB + (uint32_t)buf * 2u * kBHalf;
                        mbar_expect_tx(full, bytes);
                        for (uint32_t o = 0; o < bytes; o += 32768u) bulk_g2s(dst + o, pack + off + o, min(32768u, bytes - o), full);
                        off += bytes;
                    }
                }
            }
        }
    } else if (warp == kWorkers / 32) {
        // ======================================================================================== MMA issuer (one lane)
        if (lane == 0) {
            int it = 0, cc = 0;
            for (int pass = 0; pass < 2; pass++) {
                if (!(pass == 0 ? p.wx_pack : p.sub_pack)) continue;
                const int N = pass == 0 ? De : p.K16, n_parts = (N + kNp - 1) / kNp;
                if (pass == 1 && g_waits_u) { mbar_wait(smem_u32(&u_read), 0u, p.err_flag, 0x904); tc_fence_after(); }
                const uint32_t col0 = tmem_base + (uint32_t)(pass == 1 ? goff : 0);
                for (int c = 0; c < n_chunks; c++, cc++) {
                    const int dc = min(kDc, D - c * kDc), ab = cc & 1;
                    mbar_wait(smem_u32(&a_full[ab]), (uint32_t)((cc >> 1) & 1), p.err_flag, 0x900);
                    tc_fence_after();
                    const uint32_t a_base = sbase + kSmemA + (uint32_t)ab * 2u * kAHalf;
                    const uint32_t a_hi = (((uint32_t)kAkc >> 4) << 16) | ((a_base >> 4) & 0x3FFFu);
                    const uint32_t a_lo = (((uint32_t)kAkc >> 4) << 16) | (((a_base + kAHalf) >> 4) & 0x3FFFu);
                    for (int part = 0; part < n_parts; part++, it++) {
                        const int np = min(kNp, N - part * kNp), buf = it & 1;
                        mbar_wait(smem_u32(&b_full[buf]), (uint32_t)((it >> 1) & 1), p.err_flag, 0x902);
                        // D[:, part columns] (+)= a_hi . w_hi^T + a_lo . w_hi^T + a_hi . w_lo^T      (K = dc, in steps of 16)
                        const uint32_t bdst = sbase + kSmemB + (uint32_t)buf * 2u * kBHalf;
                        const uint32_t idesc = (1u << 4) | ((uint32_t)(np >> 3) << 17) | ((uint32_t)(kRows >> 4) << 24);
                        const uint32_t d_tmem = col0 + (uint32_t)(part * kNp);
                        const uint32_t w_hi = ((uint32_t)np << 16) | ((bdst >> 4) & 0x3FFFu);
                        const uint32_t w_lo = ((uint32_t)np << 16) | (((bdst + (uint32_t)np * (uint32_t)dc * 2u) >> 4) & 0x3FFFu);
                        const uint32_t a_step = (2u * kAkc) >> 4, w_step = 2u * (uint32_t)np;
                        uint32_t acc = c > 0 ? 1u : 0u;
#pragma unroll
                        for (int g = 0; g < 3; g++) {
                            uint32_t a = g == 1 ? a_lo : a_hi, w = g == 2 ? w_lo : w_hi;
                            if (dc == 64) {
#pragma unroll
                                for (int k = 0; k < 4; k++) { mma_f16_step(d_tmem, a, w, idesc, acc, a_step, w_step); acc = 1u; }
                            } else {
#pragma unroll 1
                                for (int k = 0; k < dc; k += 16) { mma_f16_step(d_tmem, a, w, idesc, acc, a_step, w_step); acc = 1u; }
                            }
                        }
                        tc_commit(smem_u32(&mma_done[buf]));
                    }
                    tc_commit(smem_u32(&a_free[ab]));
                }
                tc_commit(smem_u32(&pass_done[pass]));
            }
        }
    } else {
        // ====================================================================================================== workers
        const int row = tid & (kRows - 1), half = tid >> 7;            // thread = (row, half of every chunk / of every part)
        const bool live = row < nrow;
        const int64_t b = b0 + row;
        if (p.sel_best && tid < nrow) {       // state of the fused selection in the score launch that follows (one beam per vector)
            p.sel_best[b0 + tid] = ~0ull;
            p.sel_cnt[b0 + tid] = 0u;
        }
        const float* xrow = p.x + (b / p.F) * D;
        const float* hrow = p.xhat ? p.xhat + b * D : nullptr;          // step 0: xhat = 0
        float rn = 0.f;                          // this thread's share of |r_b|^2
        int cc = 0;
        const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);     // TMEM lane = row
        // ---- u_b out of TMEM: the two warps of a lane quarter split the De columns in 16-column units
        auto u_epilogue = [&]() {
            mbar_wait(smem_u32(&pass_done[0]), 0u, p.err_flag, 0x903);
            tc_fence_after();
            const int units = De >> 4, u_lo = half ? units / 2 : 0, u_hi = half ? units : units / 2;
            for (int un = u_lo; un < u_hi; un++) {
                uint32_t v[16];
                __syncwarp();
                tmem_ld16(lane_addr + (uint32_t)(un * 16), v);
                tmem_wait_ld();
                if (live) {
                    float4* dst = reinterpret_cast<float4*>(p.u + b * De + un * 16);
#pragma unroll
                    for (int i = 0; i < 4; i++)
                        dst[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&u_read));
        };
        for (int pass = 0; pass < 2; pass++) {
            if (!(pass == 0 ? p.wx_pack : p.sub_pack)) continue;
            // the distance MMAs reuse u's TMEM columns when both do not fit: they wait for the u epilogue, so it must run
            // BEFORE this thread blocks on operand buffers only those MMAs can release
            if (pass == 1 && g_waits_u) u_epilogue();
            for (int c = 0; c < n_chunks; c++, cc++) {
                const int d0 = c * kDc, dc = min(kDc, D - d0), ab = cc & 1;
                // ---- A operand chunk: this thread's row, columns d0 + [half * dc/2 ..): xhat (pass 0) or r = x_n - xhat (pass 1).
                // The global loads are issued before waiting for the buffer (the MMAs of chunk cc - 2 read it).
                const int cw = dc >> 1;      // dc is a multiple of 16, so cw is a multiple of 8
                const int c_lo = d0 + half * cw;
                const bool want_r = (pass == 1) || (p.sub_pack == nullptr && p.r != nullptr);
                bool waited = cc < 2;
#pragma unroll 1
                for (int j = 0; j < cw; j += 8) {
                    float xh[8], val[8];
#pragma unroll
                    for (int q = 0; q < 8; q++) { xh[q] = 0.f; val[q] = 0.f; }
                    if (live) {
                        if (hrow) {
                            const float4 h0 = *reinterpret_cast<const float4*>(hrow + c_lo + j), h1 = *reinterpret_cast<const float4*>(hrow + c_lo + j + 4);
                            xh[0] = h0.x; xh[1] = h0.y; xh[2] = h0.z; xh[3] = h0.w; xh[4] = h1.x; xh[5] = h1.y; xh[6] = h1.z; xh[7] = h1.w;
                        }
                        if (want_r) {
                            float xn[8];
                            const float4 x0 = *reinterpret_cast<const float4*>(xrow + c_lo + j), x1 = *reinterpret_cast<const float4*>(xrow + c_lo + j + 4);
                            xn[0] = x0.x; xn[1] = x0.y; xn[2] = x0.z; xn[3] = x0.w; xn[4] = x1.x; xn[5] = x1.y; xn[6] = x1.z; xn[7] = x1.w;
#pragma unroll
                            for (int q = 0; q < 8; q++) {
                                if (p.mean) xn[q] -= __ldg(p.mean + c_lo + j + q);
                                xn[q] /= p.std_div;               // the reference divides: (x - mean) / std (qinco_base.py:533)
                                val[q] = xn[q] - xh[q];
                            }
                            if (p.r) {
                                *reinterpret_cast<float4*>(p.r + b * D + c_lo + j) = make_float4(val[0], val[1], val[2], val[3]);
                                *reinterpret_cast<float4*>(p.r + b * D + c_lo + j + 4) = make_float4(val[4], val[5], val[6], val[7]);
                            }
                            if (pass == 1) {
#pragma unroll
                                for (int q = 0; q < 8; q++) rn = fmaf(val[q], val[q], rn);
                            }
                        }
                    }
                    if (!waited) { mbar_wait(smem_u32(&a_free[ab]), (uint32_t)(((cc - 2) >> 1) & 1), p.err_flag, 0x905); waited = true; }
                    const uint32_t kc = (uint32_t)((half * cw + j) >> 3);       // k-chunk inside the A chunk
                    const uint32_t dst = sbase + kSmemA + (uint32_t)ab * 2u * kAHalf + kc * kAkc + (uint32_t)row * 16u;
                    put_hi_lo(dst, dst + kAHalf, pass == 0 ? xh : val);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&a_full[ab]));
            }
        }
        if (p.wx_pack && !g_waits_u) u_epilogue();
        if (p.sub_pack) {
            // ---- d[b][k] = (|r_b|^2 + |S_k|^2) - 2 g -> top-A
            rn_part[half][row] = rn;
            mbar_wait(smem_u32(&pass_done[1]), 0u, p.err_flag, 0x906);
            tc_fence_after();
            bar_workers();                   // rn_part visible; every MMA is done: operands and ring are free
            const float rnorm = rn_part[0][row] + rn_part[1][row];
            const uint32_t g_addr = lane_addr + (uint32_t)goff;
            if (p.A <= 16) {
                // A <= 16 (every preset): each thread ranks its half of the row's candidates straight out of TMEM, 16 at a
                // time (sort the batch, keep the 16 smallest of list + batch); the two halves of a row meet in shared memory
                float* sn = reinterpret_cast<float*>(smem);                       // |S_k|^2, staged once
                unsigned long long* xch = reinterpret_cast<unsigned long long*>(smem + 4096);     // [kRows][2][16]
                for (int k = tid; k < p.K16; k += kWorkers) sn[k] = k < K ? __ldg(p.sub_norm + k) : 0.f;
                bar_workers();
                const int units = p.K16 >> 4, u_lo = half ? units / 2 : 0, u_hi = half ? units : units / 2;
                unsigned long long best[16];
#pragma unroll
                for (int i = 0; i < 16; i++) best[i] = ~0ull;
                for (int un = u_lo; un < u_hi; un++) {
                    uint32_t v[16];
                    __syncwarp();
                    tmem_ld16(g_addr + (uint32_t)(un * 16), v);
                    tmem_wait_ld();
                    unsigned long long key[16];
#pragma unroll
                    for (int i = 0; i < 16; i++) {
                        const int k = un * 16 + i;
                        const float d = (rnorm + sn[k]) - 2.f * __uint_as_float(v[i]);          // utils.py:346
                        key[i] = k < K ? dist_key(d, k) : ~0ull;
                    }
                    bitonic_sort16(key);
                    keep_smallest16(best, key);
                }
#pragma unroll
                for (int i = 0; i < 16; i++) xch[(row * 2 + half) * 16 + i] = best[i];
                tc_fence_before();
                bar_workers();
                if (half == 0) {
                    unsigned long long other[16];
#pragma unroll
                    for (int i = 0; i < 16; i++) other[i] = xch[(row * 2 + 1) * 16 + i];
                    keep_smallest16(best, other);
#pragma unroll
                    for (int a = 0; a < 16; a++) {
                        if (a < p.A) {
                            if (p.step0) code_s[row][a] = (uint8_t)(best[a] & 0xffull);
                            else if (live) p.idx[b * p.A + a] = (uint8_t)(best[a] & 0xffull);
                        }
                    }
                }
                if (p.step0) {
                    // first step (qinco_inference.py:239-246): beam a of vector b starts at codeword code_s[b][a] of C_0
                    bar_workers();
                    const int d4n = D >> 2, F1 = p.A;
                    for (int t = tid; t < nrow * F1 * d4n; t += kWorkers) {
                        const int d4 = t % d4n, e = t / d4n, rr = e / F1, a = e - rr * F1;
                        *reinterpret_cast<float4*>(p.xhat_out + ((b0 + rr) * F1 + a) * D + 4 * d4) =
                            __ldg(reinterpret_cast<const float4*>(p.cb0 + (size_t)code_s[rr][a] * D + 4 * d4));
                    }
                    for (int t = tid; t < nrow * F1; t += kWorkers) {
                        const int rr = t / F1, a = t - rr * F1;
                        p.hist_out[((b0 + rr) * F1 + a) * p.M] = code_s[rr][a];
                    }
                }
            } else {
                // [kRows][K16 + 1] floats from the start of the dynamic region; the odd row stride keeps the per-row stores
                // of a warp (same k, 32 rows) on 32 different banks
                float* dpre = reinterpret_cast<float*>(smem);
                const int K16 = p.K16, ld = K16 + 1;
                const int units = K16 >> 4, u_lo = half ? units / 2 : 0, u_hi = half ? units : units / 2;
                for (int un = u_lo; un < u_hi; un++) {
                    uint32_t v[16];
                    __syncwarp();
                    tmem_ld16(g_addr + (uint32_t)(un * 16), v);
                    tmem_wait_ld();
#pragma unroll
                    for (int i = 0; i < 16; i++) {
                        const int k = un * 16 + i;
                        dpre[row * ld + k] = k < K ? (rnorm + __ldg(p.sub_norm + k)) - 2.f * __uint_as_float(v[i]) : FLT_MAX;      // utils.py:346
                    }
                }
                tc_fence_before();
                bar_workers();
                // the A smallest per row, ascending, ties to the lower index (torch.topk(largest=False)); a warp ranks FOUR rows
                // at a time so that the shuffle latencies of their butterfly reductions overlap
                constexpr int kRG = 4, kWarps = kWorkers / 32;
                for (int i0 = warp; i0 < nrow; i0 += kWarps * kRG) {
                    float vals[kRG][8];
#pragma unroll
                    for (int g = 0; g < kRG; g++) {
                        const int i = i0 + g * kWarps;
#pragma unroll
                        for (int j = 0; j < 8; j++) {
                            const int k = lane + 32 * j;
                            vals[g][j] = (i < nrow && k < K) ? dpre[i * ld + k] : FLT_MAX;
                        }
                    }
                    for (int a = 0; a < p.A; a++) {
                        float bv[kRG];
                        int bi[kRG];
#pragma unroll
                        for (int g = 0; g < kRG; g++) {
                            bv[g] = FLT_MAX;
                            bi[g] = 0x7fffffff;
#pragma unroll
                            for (int j = 0; j < 8; j++) {
                                const int k = lane + 32 * j;
                                if (k < K && (vals[g][j] < bv[g] || (vals[g][j] == bv[g] && k < bi[g]))) { bv[g] = vals[g][j]; bi[g] = k; }
                            }
                        }
#pragma unroll
                        for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
                            for (int g = 0; g < kRG; g++) {
                                const float ov = __shfl_xor_sync(0xffffffffu, bv[g], off);
                                const int oi = __shfl_xor_sync(0xffffffffu, bi[g], off);
                                if (ov < bv[g] || (ov == bv[g] && oi < bi[g])) { bv[g] = ov; bi[g] = oi; }
                            }
                        }
#pragma unroll
                        for (int g = 0; g < kRG; g++) {
                            const int i = i0 + g * kWarps;
                            if (i >= nrow) continue;
                            if (bi[g] >= K) bi[g] = 0;       // all-NaN row: stay in range
#pragma unroll
                            for (int j = 0; j < 8; j++)
                                if (lane + 32 * j == bi[g]) vals[g][j] = FLT_MAX;
                            if (lane == 0) p.idx[(b0 + i) * p.A + a] = (uint8_t)bi[g];
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

}  // namespace

static_assert(kRows * (256 + 1) * 4 <= kSmemTotal, "the distance matrix must fit the dynamic shared memory");

cudaError_t launch_prep_tc(const PrepTcParams& p, cudaStream_t stream) {
    if (p.n_beams <= 0) return cudaSuccess;
    if (p.D % 16 || p.De % 16 || p.De > 512 || p.K > 256 || p.K16 % 16 || p.K16 < p.K) return cudaErrorInvalidValue;
    if (p.step0 && (p.A > 16 || !p.sub_pack || !p.cb0 || !p.xhat_out || !p.hist_out)) return cudaErrorInvalidValue;
    static int attr_dev[64] = {0};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    if (!attr_dev[dev]) {
        e = cudaFuncSetAttribute(qb_prep_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal);
        if (e != cudaSuccess) return e;
        attr_dev[dev] = 1;
    }
    const int64_t grid = (p.n_beams + kRows - 1) / kRows;
    qb_prep_tc_kernel<<<(unsigned)grid, kThreads, kSmemTotal, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace qb
