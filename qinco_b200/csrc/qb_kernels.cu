// CUDA-core (fp32) kernels around the tcgen05 MLP: beam preparation / first step, beam selection, decode start.
// These are <2 % of a step's arithmetic; they stay in exact fp32 so candidate pre-selection and the final ranking
// see the same numbers as the reference's fp32 path.
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

#include "qb_dev.h"

namespace qb {

namespace {

constexpr int kPrepThreads = 256;
constexpr int kPrepRows = 64;                 // (vector, beam) rows per block
constexpr int kTileN = 64, kTileD = 32, kTileLD = kTileD + 4;   // output tile, D chunk, padded shared-memory row
constexpr int kTileF4 = kPrepRows * kTileD / 4 / kPrepThreads;   // float4 loads per thread and operand tile (2)

// Two fp32 FMAs per instruction (Blackwell FFMA2): acc.{x,y} += a.{x,y} * b.{x,y}.  The tile GEMMs below pair the even and
// the odd dimensions of a dot product, which need no operand shuffling (the pairs are the halves of the float4 loads).
__device__ __forceinline__ void ffma2(float2& acc, const float2 a, const float2 b) {
    asm("fma.rn.f32x2 %0, %1, %2, %0;"
        : "+l"(reinterpret_cast<uint64_t&>(acc))
        : "l"(reinterpret_cast<const uint64_t&>(a)), "l"(reinterpret_cast<const uint64_t&>(b)));
}

// lexicographic (value, index) minimum across the warp
__device__ __forceinline__ void warp_argmin(float& v, int& i) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, v, off);
        const int oi = __shfl_xor_sync(0xffffffffu, i, off);
        if (ov < v || (ov == v && oi < i)) { v = ov; i = oi; }
    }
}

// Beam preparation for one step (reference: QincoSubstep.get_distances_for_codes / select_code_candidates,
// qinco_base.py:114-121; the hoisted half of QConcat, :60-64: Wcat[:, De:] . xhat; for step 0 the first lines of
// QINCoInferenceEncoder.forward, qinco_inference.py:239-246).  Two register-tiled fp32 "GEMMs" per block of 64 rows, both
// in the shape of the IVF kernel below (16 x 16 threads, 4 x 4 dot products each, D in chunks of 32 through shared memory):
//   u[b][e]    = xhat_b . Wx[e]                                     (written straight from registers)
//   dpre[b][k] = (|r_b|^2 + |S_k|^2) - 2 r_b . S_k                  (the reference's approx_pairwise_distance form,
//                utils.py:336-346, which is what it uses for > 32 rows) -> shared memory -> top-A per row
__global__ void __launch_bounds__(kPrepThreads) qb_prep_kernel(const PrepParams p) {
    extern __shared__ __align__(16) float dpre[];          // [kPrepRows][K] (only with a codebook to rank against)
    __shared__ __align__(16) float xs[kPrepRows * kTileLD];
    __shared__ __align__(16) float cs[kTileN * kTileLD];
    __shared__ float anorm_s[kPrepRows];
    const int D = p.D, De = p.De, K = p.K;
    const int tid = threadIdx.x, ti = tid >> 4, tj = tid & 15;
    const int64_t b0 = (int64_t)blockIdx.x * kPrepRows;
    const int nrow = (int)min((int64_t)kPrepRows, p.n_beams - b0);
    const int lrow = tid >> 2;                                // row of this thread in the 4-threads-per-row norm pass
    // float4 number i (< kTileF4) of this thread in a 64 x kTileD chunk: row, first column
    auto f4_row = [&](int i) { return (tid + i * kPrepThreads) / (kTileD / 4); };
    auto f4_col = [&](int i) { return ((tid + i * kPrepThreads) % (kTileD / 4)) * 4; };

    // one float4 of a row, columns d0 + c4 ..: the beam's xhat (kind 0) or its residual r = xn - xhat (kind 1);
    // columns past D (D is a multiple of 16, the chunk is 32 wide) read as zero
    auto load_a = [&](int kind, int d0, bool store_r, int row, int c4) {
        float4 xh = make_float4(0.f, 0.f, 0.f, 0.f), r = xh;
        if (row < nrow && d0 + c4 < D) {
            const int64_t b = b0 + row;
            const int d = d0 + c4;
            if (!p.step0) xh = *reinterpret_cast<const float4*>(p.xhat + b * D + d);
            if (kind == 1) {
                float4 xn = *reinterpret_cast<const float4*>(p.x + (b / p.F) * D + d);
                if (p.mean) {
                    const float4 m = *reinterpret_cast<const float4*>(p.mean + d);
                    xn.x -= m.x; xn.y -= m.y; xn.z -= m.z; xn.w -= m.w;
                }
                // the reference divides: (x - mean) / std (qinco_base.py:533); inv_std holds the divisor
                xn.x /= p.inv_std; xn.y /= p.inv_std; xn.z /= p.inv_std; xn.w /= p.inv_std;
                r = make_float4(xn.x - xh.x, xn.y - xh.y, xn.z - xh.z, xn.w - xh.w);
                if (store_r && p.r) *reinterpret_cast<float4*>(p.r + b * D + d) = r;
            }
        }
        return kind == 1 ? r : xh;
    };
    // acc[a][b] = A-row (ti + 16 a) . B-row (n0 + tj + 16 b) over all of D; B is [n_out][D] row-major
    auto tile_gemm = [&](int kind, const float* __restrict__ B, int n_out, int n0, bool store_r, float (&acc)[4][4]) {
        float2 acc2[4][4];      // (sum over even dimensions, sum over odd dimensions)
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
            for (int b = 0; b < 4; b++) acc2[a][b] = make_float2(0.f, 0.f);
        // the next chunk's global loads are issued before the current chunk's FMAs (register double buffering)
        auto load_b = [&](int d0, int row, int c4) {
            float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n0 + row < n_out && d0 + c4 < D) bv = __ldg(reinterpret_cast<const float4*>(B + (size_t)(n0 + row) * D + d0 + c4));
            return bv;
        };
        float4 av[kTileF4], bv[kTileF4];
#pragma unroll
        for (int i = 0; i < kTileF4; i++) { av[i] = load_a(kind, 0, store_r, f4_row(i), f4_col(i)); bv[i] = load_b(0, f4_row(i), f4_col(i)); }
        for (int d0 = 0; d0 < D; d0 += kTileD) {
            __syncthreads();
#pragma unroll
            for (int i = 0; i < kTileF4; i++) {
                *reinterpret_cast<float4*>(xs + f4_row(i) * kTileLD + f4_col(i)) = av[i];
                *reinterpret_cast<float4*>(cs + f4_row(i) * kTileLD + f4_col(i)) = bv[i];
            }
            __syncthreads();
            if (d0 + kTileD < D) {
#pragma unroll
                for (int i = 0; i < kTileF4; i++) {
                    av[i] = load_a(kind, d0 + kTileD, store_r, f4_row(i), f4_col(i));
                    bv[i] = load_b(d0 + kTileD, f4_row(i), f4_col(i));
                }
            }
#pragma unroll
            for (int d = 0; d < kTileD; d += 4) {
                float4 xa[4], cb[4];
#pragma unroll
                for (int a = 0; a < 4; a++) xa[a] = *reinterpret_cast<const float4*>(xs + (ti + 16 * a) * kTileLD + d);
#pragma unroll
                for (int b = 0; b < 4; b++) cb[b] = *reinterpret_cast<const float4*>(cs + (tj + 16 * b) * kTileLD + d);
#pragma unroll
                for (int a = 0; a < 4; a++)
#pragma unroll
                    for (int b = 0; b < 4; b++) {
                        ffma2(acc2[a][b], make_float2(xa[a].x, xa[a].y), make_float2(cb[b].x, cb[b].y));
                        ffma2(acc2[a][b], make_float2(xa[a].z, xa[a].w), make_float2(cb[b].z, cb[b].w));
                    }
            }
        }
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
            for (int b = 0; b < 4; b++) acc[a][b] = acc2[a][b].x + acc2[a][b].y;
    };

    if (p.sel_best && tid < nrow) {      // state of the fused selection in the score launch that follows (one beam per vector)
        p.sel_best[b0 + tid] = ~0ull;
        p.sel_cnt[b0 + tid] = 0u;
    }
    float acc[4][4];
    if (p.wx) {          // u[b][e] = Wx[e] . xhat[b]
        for (int e0 = 0; e0 < De; e0 += kTileN) {
            tile_gemm(0, p.wx, De, e0, false, acc);
#pragma unroll
            for (int a = 0; a < 4; a++) {
                const int row = ti + 16 * a;
                if (row < nrow)
#pragma unroll
                    for (int b = 0; b < 4; b++) {
                        const int e = e0 + tj + 16 * b;
                        if (e < De) p.u[(b0 + row) * De + e] = acc[a][b];
                    }
            }
        }
    }
    if (!p.sub_cb) {     // no ranking in this launch: r still has to be written (A == 0 scoring reads it)
        if (p.r)
            for (int d0 = 0; d0 < D; d0 += kTileD)
#pragma unroll
                for (int i = 0; i < kTileF4; i++) load_a(1, d0, true, f4_row(i), f4_col(i));
        return;
    }
    // |r|^2 of the block's rows: 4 threads per row, same element order for every candidate of the row
    {
        float s = 0.f;
        if (lrow < nrow) {
            const int64_t b = b0 + lrow;
            for (int d = (tid & 3); d < D; d += 4) {
                float xn = p.x[(b / p.F) * D + d];
                if (p.mean) xn -= p.mean[d];
                xn /= p.inv_std;
                const float r = xn - (p.step0 ? 0.f : p.xhat[b * D + d]);
                s = fmaf(r, r, s);
            }
        }
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        if ((tid & 3) == 0) anorm_s[lrow] = s;
    }
    for (int k0 = 0; k0 < K; k0 += kTileN) {
        tile_gemm(1, p.sub_cb, K, k0, k0 == 0, acc);       // (the first pass also writes r; anorm_s is ordered by its barriers)
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const int k = k0 + tj + 16 * b;
            if (k < K) {
                const float bn = __ldg(p.sub_norm + k);
#pragma unroll
                for (int a = 0; a < 4; a++) dpre[(ti + 16 * a) * K + k] = (anorm_s[ti + 16 * a] + bn) - 2.f * acc[a][b];
            }
        }
    }
    __syncthreads();
    // the A smallest per row, ascending, ties to the lower index (torch.topk(largest=False) order).  A warp ranks FOUR of
    // its rows at a time: the butterfly reductions of the rows are independent, so their shuffle latencies overlap.
    const int warp = tid >> 5, lane = tid & 31;
    constexpr int kRG = 4, kWarps = kPrepThreads / 32;
    for (int i0 = warp; i0 < nrow; i0 += kWarps * kRG) {
        float vals[kRG][8];
#pragma unroll
        for (int g = 0; g < kRG; g++) {
            const int i = i0 + g * kWarps;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int k = lane + 32 * j;
                vals[g][j] = (i < nrow && k < K) ? dpre[i * K + k] : FLT_MAX;
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
                if (bi[g] >= K) bi[g] = 0;   // all-NaN row: stay in range
#pragma unroll
                for (int j = 0; j < 8; j++)
                    if (lane + 32 * j == bi[g]) vals[g][j] = FLT_MAX;
                const int64_t b = b0 + i;
                if (!p.step0) {
                    if (lane == 0) p.idx[b * p.A + a] = (uint8_t)bi[g];
                } else {
                    // first step: beam a of vector b starts at codeword bi of C_0 (row-major, = the ranking codebook)
                    if (lane == 0) p.hist_out[(b * p.A + a) * p.M] = (uint8_t)bi[g];
                    for (int d = lane; d < D; d += 32)
                        p.xhat_out[(b * p.A + a) * D + d] = __ldg(p.sub_cb + (size_t)bi[g] * D + d);
                }
            }
        }
    }
}

// Reference: dists.topk(F_out, largest=False) + the three gathers of QINCoStep.encode (qinco_base.py:346-372).
__global__ void __launch_bounds__(256) qb_select_kernel(const SelectParams p) {
    const int lane = threadIdx.x & 31;
    const int64_t v = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (v >= p.n) return;
    const int R = p.F_in * p.C;
    const float* dist = p.dist + v * R;
    float last_v = -FLT_MAX;
    int last_i = -1;
    for (int j = 0; j < p.F_out; j++) {
        float bv = FLT_MAX;
        int bi = 0x7fffffff;
        for (int i = lane; i < R; i += 32) {
            const float d = dist[i];
            const bool after = (d > last_v) || (d == last_v && i > last_i);
            if (after && (d < bv || (d == bv && i < bi))) { bv = d; bi = i; }
        }
        warp_argmin(bv, bi);
        if (bi >= R) bi = (last_i + 1 < R) ? last_i + 1 : 0;   // NaN / exhausted: stay in range
        last_v = bv;
        last_i = bi;
        const int parent = bi / p.C, a = bi - parent * p.C;
        const int code = p.idx ? (int)p.idx[(v * p.F_in + parent) * p.A + a] : a;
        const int64_t o = v * p.F_out + j;
        if (lane == 0) {
            p.sel_parent[o] = (uint8_t)parent;
            p.sel_code[o] = (uint8_t)code;
            p.hist_out[o * p.M + p.m] = (uint8_t)code;
        }
        for (int t = lane; t < p.m; t += 32) p.hist_out[o * p.M + t] = p.hist_in[(v * p.F_in + parent) * p.M + t];
    }
}

// ---- IVF first step: tiled fp32 "GEMM + arg-min".  Block = 64 vectors; centroids in tiles of 64, D in chunks of 32; thread
// (ti, tj) of a 16 x 16 grid owns vectors ti + 16 a and centroids tj + 16 b (a, b < 4): 16 dot products in registers.
constexpr int kIvfVB = kPrepRows, kIvfCT = kTileN, kIvfDC = kTileD, kIvfLD = kTileLD;   // +4 floats: rows 16 B apart in bank space

__global__ void __launch_bounds__(256) qb_ivf_assign_kernel(const IvfParams p) {
    __shared__ __align__(16) float xs[kIvfVB * kIvfLD];
    __shared__ __align__(16) float cs[kIvfCT * kIvfLD];
    __shared__ float anorm_s[kIvfVB];
    __shared__ float best_v[kIvfVB][16];
    __shared__ int best_i[kIvfVB][16];
    const int tid = threadIdx.x, ti = tid >> 4, tj = tid & 15;
    const int64_t v0 = (int64_t)blockIdx.x * kIvfVB;
    const int D = p.D;
    // |x|^2 of the block's (normalised) vectors: 4 threads per vector
    {
        const int v = tid >> 2, part = tid & 3;
        float s = 0.f;
        if (v0 + v < p.n)
            for (int d = part; d < D; d += 4) {
                float xv = p.x[(v0 + v) * D + d];
                if (p.mean) xv -= p.mean[d];
                xv /= p.std_div;
                s = fmaf(xv, xv, s);
            }
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        if (part == 0) anorm_s[v] = s;
    }
    float bv[4];
    int bi[4];
#pragma unroll
    for (int a = 0; a < 4; a++) { bv[a] = FLT_MAX; bi[a] = 0x7fffffff; }
    for (int k0 = 0; k0 < p.ivf_K; k0 += kIvfCT) {
        float2 acc2[4][4];      // (even dimensions, odd dimensions) of each dot product: FFMA2
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
            for (int b = 0; b < 4; b++) acc2[a][b] = make_float2(0.f, 0.f);
        // 64 rows x 32 columns per chunk: two float4 of x and two of the centroids per thread, the next chunk's loads in
        // flight while the current chunk's FMAs run
        auto f4_row = [&](int i) { return (tid + i * 256) / (kIvfDC / 4); };
        auto f4_col = [&](int i) { return ((tid + i * 256) % (kIvfDC / 4)) * 4; };
        auto load_x = [&](int d0, int row, int c4) {
            float4 xv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (v0 + row < p.n && d0 + c4 < D) {
                xv = *reinterpret_cast<const float4*>(p.x + (v0 + row) * D + d0 + c4);
                if (p.mean) {
                    const float4 m = *reinterpret_cast<const float4*>(p.mean + d0 + c4);
                    xv.x -= m.x; xv.y -= m.y; xv.z -= m.z; xv.w -= m.w;
                }
                xv.x /= p.std_div; xv.y /= p.std_div; xv.z /= p.std_div; xv.w /= p.std_div;
            }
            return xv;
        };
        auto load_c = [&](int d0, int row, int c4) {
            float4 cv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (k0 + row < p.ivf_K && d0 + c4 < D) cv = __ldg(reinterpret_cast<const float4*>(p.cent + (size_t)(k0 + row) * D + d0 + c4));
            return cv;
        };
        float4 xv[kTileF4], cv[kTileF4];
#pragma unroll
        for (int i = 0; i < kTileF4; i++) { xv[i] = load_x(0, f4_row(i), f4_col(i)); cv[i] = load_c(0, f4_row(i), f4_col(i)); }
        for (int d0 = 0; d0 < D; d0 += kIvfDC) {
            __syncthreads();
#pragma unroll
            for (int i = 0; i < kTileF4; i++) {
                *reinterpret_cast<float4*>(xs + f4_row(i) * kIvfLD + f4_col(i)) = xv[i];
                *reinterpret_cast<float4*>(cs + f4_row(i) * kIvfLD + f4_col(i)) = cv[i];
            }
            __syncthreads();
            if (d0 + kIvfDC < D) {
#pragma unroll
                for (int i = 0; i < kTileF4; i++) {
                    xv[i] = load_x(d0 + kIvfDC, f4_row(i), f4_col(i));
                    cv[i] = load_c(d0 + kIvfDC, f4_row(i), f4_col(i));
                }
            }
#pragma unroll
            for (int d = 0; d < kIvfDC; d += 4) {
                float4 xa[4], cb[4];
#pragma unroll
                for (int a = 0; a < 4; a++) xa[a] = *reinterpret_cast<const float4*>(xs + (ti + 16 * a) * kIvfLD + d);
#pragma unroll
                for (int b = 0; b < 4; b++) cb[b] = *reinterpret_cast<const float4*>(cs + (tj + 16 * b) * kIvfLD + d);
#pragma unroll
                for (int a = 0; a < 4; a++)
#pragma unroll
                    for (int b = 0; b < 4; b++) {
                        ffma2(acc2[a][b], make_float2(xa[a].x, xa[a].y), make_float2(cb[b].x, cb[b].y));
                        ffma2(acc2[a][b], make_float2(xa[a].z, xa[a].w), make_float2(cb[b].z, cb[b].w));
                    }
            }
        }
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const int k = k0 + tj + 16 * b;
            if (k < p.ivf_K) {
                const float bn = __ldg(p.cnorm + k);
#pragma unroll
                for (int a = 0; a < 4; a++) {
                    const float dist = (anorm_s[ti + 16 * a] + bn) - 2.f * (acc2[a][b].x + acc2[a][b].y);   // utils.py:346
                    if (dist < bv[a] || (dist == bv[a] && k < bi[a])) { bv[a] = dist; bi[a] = k; }
                }
            }
        }
    }
#pragma unroll
    for (int a = 0; a < 4; a++) { best_v[ti + 16 * a][tj] = bv[a]; best_i[ti + 16 * a][tj] = bi[a]; }
    __syncthreads();
    if (tid < kIvfVB) {
        float v = best_v[tid][0];
        int i = best_i[tid][0];
        for (int j = 1; j < 16; j++) {
            const float ov = best_v[tid][j];
            const int oi = best_i[tid][j];
            if (ov < v || (ov == v && oi < i)) { v = ov; i = oi; }
        }
        if (i >= p.ivf_K) i = 0;      // all-NaN row: stay in range
        best_i[tid][0] = i;
        if (v0 + tid < p.n) p.codes_out[v0 + tid] = i;
    }
    __syncthreads();
    for (int t = tid; t < kIvfVB * (D / 4); t += 256) {
        const int v = t / (D / 4), d4 = t - v * (D / 4);
        if (v0 + v < p.n)
            *reinterpret_cast<float4*>(p.xhat_out + (v0 + v) * D + d4 * 4) =
                __ldg(reinterpret_cast<const float4*>(p.cent + (size_t)best_i[v][0] * D + d4 * 4));
    }
}

__global__ void qb_ivf_lookup_kernel(const float* __restrict__ cent, const int32_t* __restrict__ codes, int64_t n, int D,
                                     int ivf_K, float* __restrict__ xhat, uint32_t* err_flag) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t v = t / (D / 4);
    const int d = (int)(t - v * (D / 4)) * 4;
    if (v >= n) return;
    int c = codes[v];
    if (c < 0 || c >= ivf_K) { if (err_flag) atomicExch(err_flag, 0x20u); c = 0; }
    *reinterpret_cast<float4*>(xhat + v * D + d) = __ldg(reinterpret_cast<const float4*>(cent + (size_t)c * D + d));
}

__global__ void qb_decode_init_kernel(const float* __restrict__ cb0, const uint8_t* __restrict__ codes, int64_t n, int M,
                                      int D, int K, float* __restrict__ xhat, uint32_t* err_flag) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t v = t / (D / 4);
    const int d = (int)(t - v * (D / 4)) * 4;
    if (v >= n) return;
    int c = codes[v * M];
    if (c >= K) { if (err_flag) atomicExch(err_flag, 0x10u); c = K - 1; }
    *reinterpret_cast<float4*>(xhat + v * D + d) = __ldg(reinterpret_cast<const float4*>(cb0 + (size_t)c * D + d));
}

__global__ void qb_affine_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t total, int D, float scale,
                                 const float* __restrict__ shift) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    out[t] = fmaf(in[t], scale, shift ? shift[t % D] : 0.f);
}

// [S, n] integer code matrix (any strides, int32 / int64) -> uint8 [n, M] (+ int32 [n] IVF codes), range-checked
template <typename T>
__global__ void qb_codes_pack_kernel(const T* __restrict__ src, int64_t stride_row, int64_t stride_col, int64_t n, int M, int K,
                                     int ivf_K, uint8_t* __restrict__ codes, int32_t* __restrict__ ivf, uint32_t* err_flag) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    const int row0 = ivf_K ? 1 : 0;
    if (ivf_K) {
        long long c = (long long)src[v * stride_col];
        if (c < 0 || c >= ivf_K) { atomicExch(err_flag, 0x20u); c = 0; }
        ivf[v] = (int32_t)c;
    }
    for (int m = 0; m < M; m++) {
        long long c = (long long)src[(int64_t)(m + row0) * stride_row + v * stride_col];
        if (c < 0 || c >= K) { atomicExch(err_flag, 0x10u); c = 0; }
        codes[v * M + m] = (uint8_t)c;
    }
}

// uint8 [n, M] (+ int32 [n]) -> contiguous int64 [S, n]; a thread owns one vector, the stores of a warp are coalesced per row
__global__ void qb_codes_unpack_kernel(const uint8_t* __restrict__ codes, const int32_t* __restrict__ ivf, int64_t n, int M,
                                       int has_ivf, int64_t* __restrict__ dst) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    if (has_ivf) dst[v] = (int64_t)ivf[v];
    for (int m = 0; m < M; m++) dst[(int64_t)(m + has_ivf) * n + v] = (int64_t)codes[v * M + m];
}

__global__ void qb_take_codes_kernel(const unsigned long long* __restrict__ best, int64_t n, const uint8_t* __restrict__ hist_in,
                                     uint8_t* __restrict__ hist_out, int M, int m) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    for (int c = 0; c < m; c++) hist_out[v * M + c] = hist_in[v * M + c];
    hist_out[v * M + m] = (uint8_t)(best[v] & 0xffull);
}

}  // namespace

cudaError_t launch_take_codes(const unsigned long long* sel_best, int64_t n, const uint8_t* hist_in, uint8_t* hist_out, int M,
                              int m, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    qb_take_codes_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(sel_best, n, hist_in, hist_out, M, m);
    return cudaGetLastError();
}

cudaError_t launch_codes_pack(const void* codes_MB, int elem_bytes, int64_t stride_row, int64_t stride_col, int64_t n, int M,
                              int K, int ivf_K, uint8_t* codes_u8, int32_t* ivf, uint32_t* err_flag, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    const unsigned grid = (unsigned)((n + 255) / 256);
    if (elem_bytes == 8)
        qb_codes_pack_kernel<long long><<<grid, 256, 0, stream>>>((const long long*)codes_MB, stride_row, stride_col, n, M, K, ivf_K,
                                                                   codes_u8, ivf, err_flag);
    else if (elem_bytes == 4)
        qb_codes_pack_kernel<int32_t><<<grid, 256, 0, stream>>>((const int32_t*)codes_MB, stride_row, stride_col, n, M, K, ivf_K,
                                                                codes_u8, ivf, err_flag);
    else if (elem_bytes == 1)
        qb_codes_pack_kernel<uint8_t><<<grid, 256, 0, stream>>>((const uint8_t*)codes_MB, stride_row, stride_col, n, M, K, ivf_K,
                                                                codes_u8, ivf, err_flag);
    else
        return cudaErrorInvalidValue;
    return cudaGetLastError();
}

cudaError_t launch_codes_unpack(const uint8_t* codes_u8, const int32_t* ivf, int64_t n, int M, int has_ivf, int64_t* codes_MB,
                                cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    qb_codes_unpack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(codes_u8, ivf, n, M, has_ivf, codes_MB);
    return cudaGetLastError();
}

cudaError_t launch_prep(const PrepParams& p, cudaStream_t stream) {
    if (p.n_beams <= 0) return cudaSuccess;
    if (p.K > 256 || p.D % 16) return cudaErrorInvalidValue;
    const size_t smem = p.sub_cb ? (size_t)kPrepRows * p.K * sizeof(float) : 0;
    if (smem > 48 * 1024) {       // the attribute is per device: remember what each one was raised to
        static size_t attr_set[64] = {0};
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
        if (smem > attr_set[dev]) {
            e = cudaFuncSetAttribute(qb_prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            attr_set[dev] = smem;
        }
    }
    const int64_t grid = (p.n_beams + kPrepRows - 1) / kPrepRows;
    qb_prep_kernel<<<(unsigned)grid, kPrepThreads, smem, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_select(const SelectParams& p, cudaStream_t stream) {
    if (p.n <= 0) return cudaSuccess;
    const int wpb = 8;
    const int64_t grid = (p.n + wpb - 1) / wpb;
    qb_select_kernel<<<(unsigned)grid, wpb * 32, 0, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_ivf_assign(const IvfParams& p, cudaStream_t stream) {
    if (p.n <= 0) return cudaSuccess;
    if (p.D % 16 || p.ivf_K < 1) return cudaErrorInvalidValue;
    qb_ivf_assign_kernel<<<(unsigned)((p.n + kIvfVB - 1) / kIvfVB), 256, 0, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_ivf_lookup(const float* cent, const int32_t* ivf_codes, int64_t n, int D, int ivf_K, float* xhat,
                              uint32_t* err_flag, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    const int64_t total = n * (D / 4);
    qb_ivf_lookup_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(cent, ivf_codes, n, D, ivf_K, xhat, err_flag);
    return cudaGetLastError();
}

cudaError_t launch_decode_init(const float* cb0, const uint8_t* codes, int64_t n, int M, int D, int K, float* xhat,
                               uint32_t* err_flag, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    const int64_t total = n * (D / 4);
    qb_decode_init_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(cb0, codes, n, M, D, K, xhat, err_flag);
    return cudaGetLastError();
}

cudaError_t launch_affine(const float* in, float* out, int64_t n, int D, float scale, const float* shift,
                          cudaStream_t stream) {
    const int64_t total = n * D;
    if (total <= 0) return cudaSuccess;
    qb_affine_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(in, out, total, D, scale, shift);
    return cudaGetLastError();
}

}  // namespace qb
