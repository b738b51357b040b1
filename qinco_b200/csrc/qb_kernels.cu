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
constexpr int kTileN = 64, kTileD = 16, kTileLD = kTileD + 4;   // output tile, D chunk, padded shared-memory row

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
// in the shape of the IVF kernel below (16 x 16 threads, 4 x 4 dot products each, D in chunks of 16 through shared memory):
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
    const int lrow = tid >> 2, lc4 = (tid & 3) * 4;          // this thread's float4 of a 64 x 16 chunk

    // one float4 of row lrow, columns d0 + lc4 ..: the beam's xhat (kind 0) or its residual r = xn - xhat (kind 1)
    auto load_a = [&](int kind, int d0, bool store_r) {
        float4 xh = make_float4(0.f, 0.f, 0.f, 0.f), r = xh;
        if (lrow < nrow) {
            const int64_t b = b0 + lrow;
            const int d = d0 + lc4;
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
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
            for (int b = 0; b < 4; b++) acc[a][b] = 0.f;
        // the next chunk's global loads are issued before the current chunk's FMAs (register double buffering)
        auto load_b = [&](int d0) {
            float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n0 + lrow < n_out) bv = __ldg(reinterpret_cast<const float4*>(B + (size_t)(n0 + lrow) * D + d0 + lc4));
            return bv;
        };
        float4 av = load_a(kind, 0, store_r), bv = load_b(0);
        for (int d0 = 0; d0 < D; d0 += kTileD) {
            __syncthreads();
            *reinterpret_cast<float4*>(xs + lrow * kTileLD + lc4) = av;
            *reinterpret_cast<float4*>(cs + lrow * kTileLD + lc4) = bv;
            __syncthreads();
            if (d0 + kTileD < D) { av = load_a(kind, d0 + kTileD, store_r); bv = load_b(d0 + kTileD); }
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
                        acc[a][b] = fmaf(xa[a].x, cb[b].x, acc[a][b]); acc[a][b] = fmaf(xa[a].y, cb[b].y, acc[a][b]);
                        acc[a][b] = fmaf(xa[a].z, cb[b].z, acc[a][b]); acc[a][b] = fmaf(xa[a].w, cb[b].w, acc[a][b]);
                    }
            }
        }
    };

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
            for (int d0 = 0; d0 < D; d0 += kTileD) load_a(1, d0, true);
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
    // the A smallest per row, ascending, ties to the lower index (torch.topk(largest=False) order)
    const int warp = tid >> 5, lane = tid & 31;
    for (int i = warp; i < nrow; i += kPrepThreads / 32) {
        const int64_t b = b0 + i;
        float vals[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int k = lane + 32 * j;
            vals[j] = k < K ? dpre[i * K + k] : FLT_MAX;
        }
        for (int a = 0; a < p.A; a++) {
            float bv = FLT_MAX;
            int bi = 0x7fffffff;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int k = lane + 32 * j;
                if (k < K && (vals[j] < bv || (vals[j] == bv && k < bi))) { bv = vals[j]; bi = k; }
            }
            warp_argmin(bv, bi);
            if (bi >= K) bi = 0;   // all-NaN row: stay in range
#pragma unroll
            for (int j = 0; j < 8; j++)
                if (lane + 32 * j == bi) vals[j] = FLT_MAX;
            if (!p.step0) {
                if (lane == 0) p.idx[b * p.A + a] = (uint8_t)bi;
            } else {
                // first step: beam a of vector b starts at codeword bi of C_0 (row-major, = the ranking codebook)
                if (lane == 0) p.hist_out[(b * p.A + a) * p.M] = (uint8_t)bi;
                for (int d = lane; d < D; d += 32)
                    p.xhat_out[(b * p.A + a) * D + d] = __ldg(p.sub_cb + (size_t)bi * D + d);
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

// ---- IVF first step: tiled fp32 "GEMM + arg-min".  Block = 64 vectors; centroids in tiles of 64, D in chunks of 16; thread
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
        float acc[4][4];
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
            for (int b = 0; b < 4; b++) acc[a][b] = 0.f;
        // 64 rows x 16 columns per chunk: one float4 of x and one of the centroids per thread, the next chunk's loads in
        // flight while the current chunk's FMAs run
        const int row = tid >> 2, c4 = (tid & 3) * 4;
        auto load_x = [&](int d0) {
            float4 xv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (v0 + row < p.n) {
                xv = *reinterpret_cast<const float4*>(p.x + (v0 + row) * D + d0 + c4);
                if (p.mean) {
                    const float4 m = *reinterpret_cast<const float4*>(p.mean + d0 + c4);
                    xv.x -= m.x; xv.y -= m.y; xv.z -= m.z; xv.w -= m.w;
                }
                xv.x /= p.std_div; xv.y /= p.std_div; xv.z /= p.std_div; xv.w /= p.std_div;
            }
            return xv;
        };
        auto load_c = [&](int d0) {
            float4 cv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (k0 + row < p.ivf_K) cv = __ldg(reinterpret_cast<const float4*>(p.cent + (size_t)(k0 + row) * D + d0 + c4));
            return cv;
        };
        float4 xv = load_x(0), cv = load_c(0);
        for (int d0 = 0; d0 < D; d0 += kIvfDC) {
            __syncthreads();
            *reinterpret_cast<float4*>(xs + row * kIvfLD + c4) = xv;
            *reinterpret_cast<float4*>(cs + row * kIvfLD + c4) = cv;
            __syncthreads();
            if (d0 + kIvfDC < D) { xv = load_x(d0 + kIvfDC); cv = load_c(d0 + kIvfDC); }
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
                        acc[a][b] = fmaf(xa[a].x, cb[b].x, acc[a][b]); acc[a][b] = fmaf(xa[a].y, cb[b].y, acc[a][b]);
                        acc[a][b] = fmaf(xa[a].z, cb[b].z, acc[a][b]); acc[a][b] = fmaf(xa[a].w, cb[b].w, acc[a][b]);
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
                    const float dist = (anorm_s[ti + 16 * a] + bn) - 2.f * acc[a][b];   // utils.py:346
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

}  // namespace

cudaError_t launch_prep(const PrepParams& p, cudaStream_t stream) {
    if (p.n_beams <= 0) return cudaSuccess;
    if (p.K > 256 || p.D % 16) return cudaErrorInvalidValue;
    const size_t smem = p.sub_cb ? (size_t)kPrepRows * p.K * sizeof(float) : 0;
    static size_t attr_set = 0;
    if (smem > 48 * 1024 && smem > attr_set) {
        cudaError_t e = cudaFuncSetAttribute(qb_prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_set = smem;
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
