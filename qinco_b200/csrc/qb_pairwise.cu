// Pairwise additive decoder, forward only.
//
// Replaces PairwiseDecoderIVF.forward + map_codes (reference qinco/search/pairwise_decoder.py:88-93, :126-130):
//     ext[b]   = codes[b, 0 .. M) ++ ivf_code_map[ivf_codes[b], 0 .. 5)                 (M + 5 small codes per vector)
//     comb[j]  = ext[m1_j] * K + ext[m2_j]                    j < Mt                    (index into a K^2-row table)
//     out[b]   = T[0][comb[0]] + T[1][comb[1]] + ... + T[Mt-1][comb[Mt-1]]              (fp32, added in this order)
// A pure gather-accumulate: Mt rows of D floats are read per vector from a table of Mt * K^2 * D floats (3.2 GB at the
// Contriever shape) and D floats are written, so the kernel is HBM-bound; the additions are done in the reference's order,
// which makes the result bit-identical to the PyTorch CPU path.
//
// One thread owns one float4 column slot of one vector: it rebuilds the vector's Mt table indices from the (L1-resident)
// code bytes, issues 8 independent 16-byte streaming loads per batch and then adds them in order.  Consecutive threads own
// consecutive slots of the same row, so every table row is read with fully coalesced 512-byte warp requests.
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/qinco_b200.h"
#include "qb_dev.h"

namespace {

constexpr int kIvfM = 5;          // PairwiseDecoderIVF.IVF_M (pairwise_decoder.py:15)
constexpr int kMaxMt = 64;
constexpr int kMaxExt = 64 + kIvfM;

struct PwParams {
    int32_t D4, M, K, Mt, ivf_K;
    int64_t n;
    const uint8_t* codes;        // [n, M]
    const int32_t* ivf_codes;    // [n]
    const uint8_t* ivf_map;      // [ivf_K, 5]
    const float4* table;         // [Mt][K*K][D/4]
    float4* out;                 // [n][D/4]
    uint32_t* err_flag;
    uint8_t m1[kMaxMt], m2[kMaxMt];
};

__device__ __forceinline__ float4 ld_stream(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream(float4* p, float4 v) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

template <int kBatch, int kBlocks>
__global__ void __launch_bounds__(256, kBlocks) qb_pairwise_kernel(const __grid_constant__ PwParams p) {
    const int64_t total = p.n * p.D4;
    const size_t plane = (size_t)p.K * p.K * p.D4;        // float4 elements of one table
    for (int64_t item = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; item < total; item += (int64_t)gridDim.x * blockDim.x) {
        const int64_t v = item / p.D4;
        const int slot = (int)(item - v * p.D4);
        const uint8_t* cv = p.codes + v * p.M;
        int iv = __ldg(p.ivf_codes + v);
        if (iv < 0 || iv >= p.ivf_K) { atomicExch(p.err_flag, 0x20u); iv = 0; }
        const uint8_t* mv = p.ivf_map + (size_t)iv * kIvfM;
        auto ext = [&](int m) -> uint32_t {
            uint32_t c = m < p.M ? __ldg(cv + m) : __ldg(mv + (m - p.M));
            if (c >= (uint32_t)p.K) { atomicExch(p.err_flag, 0x10u); c = p.K - 1; }
            return c;
        };
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int j0 = 0; j0 < p.Mt; j0 += kBatch) {
            float4 rows[kBatch];
#pragma unroll
            for (int j = 0; j < kBatch; j++) {
                if (j0 + j < p.Mt) {
                    const uint32_t comb = ext(p.m1[j0 + j]) * (uint32_t)p.K + ext(p.m2[j0 + j]);
                    rows[j] = ld_stream(p.table + (size_t)(j0 + j) * plane + (size_t)comb * p.D4 + slot);
                }
            }
#pragma unroll
            for (int j = 0; j < kBatch; j++) {
                if (j0 + j < p.Mt) {
                    if (j0 + j == 0) acc = rows[0];     // xhat = T[0][c0]; then += in table order (pairwise_decoder.py:90-92)
                    else { acc.x += rows[j].x; acc.y += rows[j].y; acc.z += rows[j].z; acc.w += rows[j].w; }
                }
            }
        }
        st_stream(p.out + item, acc);
    }
}

thread_local std::string g_pw_err;
int pw_fail(int code, const std::string& msg) {
    g_pw_err = msg;
    return code;
}

}  // namespace

struct qb_pairwise {
    int D = 0, M = 0, K = 0, Mt = 0, ivf_K = 0, device = 0, n_sm = 0;
    float* table = nullptr;
    uint8_t* ivf_map = nullptr;
    uint8_t m1[kMaxMt], m2[kMaxMt];
    uint32_t* err_host = nullptr;
    uint32_t* err_dev = nullptr;
    int64_t launches = 0;
};

extern "C" {

const char* qb_pairwise_last_error(void) { return g_pw_err.c_str(); }

int qb_pairwise_create(const qb_pairwise_desc* d, qb_pairwise** out) {
    if (!out) return pw_fail(QB_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (!d || !d->codebook || !d->combine || !d->ivf_code_map) return pw_fail(QB_ERR_INVALID, "NULL field in the description");
    if (d->D < 4 || d->D % 4) return pw_fail(QB_ERR_INVALID, "D must be a positive multiple of 4");
    if (d->K < 1 || d->K > 256) return pw_fail(QB_ERR_INVALID, "K must be in [1,256] (codes are uint8)");
    if (d->M < 1 || d->M > 64) return pw_fail(QB_ERR_INVALID, "M must be in [1,64]");
    if (d->Mt < 1 || d->Mt > kMaxMt) return pw_fail(QB_ERR_INVALID, "number of pairwise codebooks must be in [1,64]");
    if (d->ivf_K < 1) return pw_fail(QB_ERR_INVALID, "ivf_K must be >= 1 (map_codes needs IVF codes, pairwise_decoder.py:127)");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || d->device < 0 || d->device >= ndev)
        return pw_fail(QB_ERR_CUDA, "no usable CUDA device (there is no CPU fallback)");
    qb::DeviceGuard guard(d->device);       // the caller's current device is restored on exit
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, d->device) != cudaSuccess || prop.major != 10)
        return pw_fail(QB_ERR_CUDA, "this library is built for sm_100a only");
    qb_pairwise* h = new (std::nothrow) qb_pairwise();
    if (!h) return pw_fail(QB_ERR_NOMEM, "out of host memory");
    h->D = d->D; h->M = d->M; h->K = d->K; h->Mt = d->Mt; h->ivf_K = d->ivf_K; h->device = d->device;
    h->n_sm = prop.multiProcessorCount;
    for (int j = 0; j < d->Mt; j++) {
        const int64_t a = d->combine[j], b = d->combine[d->Mt + j];
        if (a < 0 || a >= d->M + kIvfM || b < 0 || b >= d->M + kIvfM) {
            delete h;
            return pw_fail(QB_ERR_INVALID, "combine_mvals_m entry out of range [0, M + 5)");
        }
        h->m1[j] = (uint8_t)a;
        h->m2[j] = (uint8_t)b;
    }
    std::vector<uint8_t> map8((size_t)d->ivf_K * kIvfM);
    for (size_t i = 0; i < map8.size(); i++) {
        const int64_t c = d->ivf_code_map[i];
        if (c < 0 || c >= d->K) {
            delete h;
            return pw_fail(QB_ERR_INVALID, "ivf_code_map entry out of range [0, K)");
        }
        map8[i] = (uint8_t)c;
    }
    const size_t tbytes = (size_t)d->Mt * d->K * d->K * d->D * sizeof(float);
    auto bail = [&](const char* what, cudaError_t e) {
        std::string msg = std::string(what) + ": " + cudaGetErrorString(e);
        qb_pairwise_destroy(h);
        return pw_fail(QB_ERR_CUDA, msg);
    };
    cudaError_t e;
    if ((e = cudaMalloc((void**)&h->table, tbytes)) != cudaSuccess) return bail("cudaMalloc(table)", e);
    if ((e = cudaMemcpy(h->table, d->codebook, tbytes, cudaMemcpyHostToDevice)) != cudaSuccess) return bail("cudaMemcpy(table)", e);
    if ((e = cudaMalloc((void**)&h->ivf_map, map8.size())) != cudaSuccess) return bail("cudaMalloc(ivf map)", e);
    if ((e = cudaMemcpy(h->ivf_map, map8.data(), map8.size(), cudaMemcpyHostToDevice)) != cudaSuccess) return bail("cudaMemcpy(ivf map)", e);
    if ((e = cudaHostAlloc((void**)&h->err_host, 64, cudaHostAllocMapped)) != cudaSuccess) return bail("cudaHostAlloc", e);
    *h->err_host = 0;
    if ((e = cudaHostGetDevicePointer((void**)&h->err_dev, h->err_host, 0)) != cudaSuccess) return bail("cudaHostGetDevicePointer", e);
    *out = h;
    return QB_OK;
}

int qb_pairwise_destroy(qb_pairwise* h) {
    if (!h) return QB_OK;
    qb::DeviceGuard guard(h->device);
    if (h->table) cudaFree(h->table);
    if (h->ivf_map) cudaFree(h->ivf_map);
    if (h->err_host) cudaFreeHost(h->err_host);
    delete h;
    return QB_OK;
}

int qb_pairwise_decode(qb_pairwise* h, const uint8_t* codes_dev, const int32_t* ivf_codes_dev, int64_t n, float* out_dev,
                       void* stream) {
    if (!h) return pw_fail(QB_ERR_INVALID, "handle is NULL");
    if (n < 0) return pw_fail(QB_ERR_INVALID, "n < 0");
    if (n == 0) return QB_OK;
    if (!codes_dev || !ivf_codes_dev || !out_dev) return pw_fail(QB_ERR_INVALID, "NULL buffer");
    qb::DeviceGuard guard(h->device);
    PwParams p;
    std::memset(&p, 0, sizeof(p));
    p.D4 = h->D / 4; p.M = h->M; p.K = h->K; p.Mt = h->Mt; p.ivf_K = h->ivf_K; p.n = n;
    p.codes = codes_dev; p.ivf_codes = ivf_codes_dev; p.ivf_map = h->ivf_map;
    p.table = reinterpret_cast<const float4*>(h->table);
    p.out = reinterpret_cast<float4*>(out_dev);
    p.err_flag = h->err_dev;
    std::memcpy(p.m1, h->m1, sizeof(p.m1));
    std::memcpy(p.m2, h->m2, sizeof(p.m2));
    const int64_t total = n * p.D4;
    const int64_t want = (total + 255) / 256;
    // 8 loads in flight per thread x 4 resident blocks per SM measured best (0.93 of the HBM copy peak at d = 768; 16 loads
    // x 2 blocks: 0.74, 4 loads x 8 blocks: 0.85 -- profiles/r01_bench_c5pw.json)
    const int64_t cap = (int64_t)h->n_sm * 8;
    const unsigned grid = (unsigned)(want < cap ? want : cap);
    qb_pairwise_kernel<8, 4><<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    h->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return pw_fail(QB_ERR_CUDA, std::string("qb_pairwise_kernel launch: ") + cudaGetErrorString(e));
    return QB_OK;
}

int qb_pairwise_check(qb_pairwise* h) {
    if (!h) return pw_fail(QB_ERR_INVALID, "handle is NULL");
    const uint32_t e = *reinterpret_cast<volatile uint32_t*>(h->err_host);
    if (e) return pw_fail(QB_ERR_KERNEL, e == 0x20u ? "IVF code out of range [0, ivf_K)" : "code out of range [0, K)");
    return QB_OK;
}

int64_t qb_pairwise_launch_count(const qb_pairwise* h) { return h ? h->launches : 0; }

}  // extern "C"
