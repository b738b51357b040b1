// Device-side parameter blocks and launcher prototypes (qb_mlp.cu, qb_kernels.cu), used by qb_api.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "qb_plan.h"

namespace qb {

// Makes `device` current for the lifetime of the guard and restores the caller's device afterwards: no entry point of the
// library changes the process's current device behind the caller's back (multi-GPU callers keep allocating where they were).
struct DeviceGuard {
    int prev = -1;
    cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int device) {
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && prev != device) err = cudaSetDevice(device);
        else if (err == cudaSuccess) prev = -1;       // nothing to restore
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};

enum { QB_MODE_SCORE = 0, QB_MODE_APPLY = 1 };
enum { QB_TRACE_EVENTS = 2048 };

// One launch of the tcgen05 MLP kernel over `n_rows` candidate rows of step m.
//   score: row -> beam b = row / C, slot a = row % C, code = A ? idx[b*A + a] : a;
//          dist[row] = || r[b] - f_m(C_m[code], xhat_b) ||^2                       (qinco_base.py:329-345)
//   apply: row -> vector v = row / F_out, b = v*F_in + parent[row], code = sel_code[row*code_stride + code_off];
//          xhat_out[row] = (xhat_in[b] + f_m(C_m[code], xhat_in[b])) * out_scale + out_shift   (qinco_base.py:363-369,
//          decode :282-290, :447-452)
// per-step tables of the decode loop (one tile walks all steps inside one launch)
struct QbLoopStep {
    const uint8_t* w_blob;
    const float* t_blk;
    const float* cb_blk;
};

struct MlpParams {
    QbStepPlan plan;
    QbOp ops[QB_MAX_OPS];     // the plan's op list; lives in the kernel-parameter constant bank so the MMA / producer
                              // warps read it through the uniform datapath
    int32_t n_ops;
    const uint8_t* w_blob;    // packed fp16 slabs of this step
    const float* t_blk;       // [De/4][K][4]  T_m
    const float* cb_blk;      // [D/4][K][4]   C_m (outer skip), unused in qinco1_mode
    int32_t mode;
    int32_t C, A;             // score
    int32_t F_in, F_out;      // apply
    int64_t n_rows;
    const uint8_t* idx;       // score, A > 0: [n_beams, A]
    const uint8_t* sel_parent;  // apply: [n_rows] or NULL (parent 0)
    const uint8_t* sel_code;    // apply
    int64_t code_stride;
    int32_t code_off;
    const float* u;           // [n_beams, De]
    const float* r;           // score: [n_beams, D]
    const float* xhat_in;     // apply: [n_beams, D]
    float* dist;              // score: [n_rows]
    float* xhat_out;          // apply: [n_rows, D]
    float out_scale;          // apply
    const float* out_shift;   // apply: [D] or NULL
    // Fused beam selection (fuse != 0, score mode; reference QINCoStep.encode, qinco_base.py:343-372: distances -> topk ->
    // gathers).  The score launch itself picks the F_out best candidates of every vector and writes their xhat' = xhat_b + o
    // and code history, so the step needs no `dist` array, no select launch and no second (apply) MLP launch.
    //   resident launches (A == 0, C == 256, F_in == F_out == 1): a vector's 256 candidates are spread over the 4 code-quarter
    //   CTAs; each keeps its local winner's o in shared memory, publishes (dist, code) with a packed 64-bit atomicMin on
    //   sel_best[v] and counts itself on sel_cnt[v]; one set later the select warp of the CTA that owns the global winner adds
    //   xhat_b and writes xhat' / the history.  sel_best / sel_cnt are initialised by the step's prep launch.
    //   fuse == 3 (one-tile-per-CTA shapes with an out_proj, i.e. the QINCo2-L family): the sel_spv tiles of a vector are walked
    //   back to back by one CTA, which keeps the vector's running top-F_out and the winners' xhat' in shared memory and emits
    //   xhat' / history when the vector is complete -- no atomics, no second launch.
    int32_t fuse;
    int32_t dbg;                    // timing experiments only (QB_FUSEB_DEBUG): skip phases of the fused selection
    int32_t sel_spv;                // fuse == 3: tiles per vector (1 when a tile holds whole vectors)
    int32_t hist_M, hist_m;         // history row length, column written by this step
    unsigned long long* sel_best;   // [n vectors]  (dist bits << 32 | code)
    uint32_t* sel_cnt;              // [n vectors]
    const uint8_t* hist_in;         // [n, F_in, M]
    uint8_t* hist_out;              // [n, F_out, M]
    // Decode loop (n_loop_steps > 0, apply mode, F_in = F_out = 1; reference QINCoInferenceDecoder.forward,
    // qinco_inference.py:66-75): row i starts at seed_tab[seed code] (C_0[codes[i][0]], or the IVF centroid of
    // seed_codes_i32[i]) and walks loop step s = 0 .. n_loop_steps-1 with code sel_code[i*code_stride + code_off + s];
    // xhat_out [n_rows, D] holds the running reconstruction between steps (each thread re-reads only what it wrote) and
    // the result (scaled / shifted at the last step).  plan.n_ops_pre > 0 is required.
    int32_t n_loop_steps;
    int32_t seed_K;
    const float* seed_tab;          // [seed_K][D] row-major
    const int32_t* seed_codes_i32;  // IVF codes [n_rows], or NULL: the seed code is sel_code[i*code_stride]
    QbLoopStep loop_steps[QB_MAX_LOOP_STEPS];
    uint32_t* err_flag;       // device word set non-zero when a barrier wait timed out
    unsigned long long* trace;   // debug: per-role event log of CTA 0 ([3][QB_TRACE_EVENTS] of clock<<16 | id), or NULL
};

cudaError_t launch_mlp(const MlpParams& p, int n_sm, cudaStream_t stream);
cudaError_t mlp_set_smem_attr(int smem_bytes);

// Beam preparation for step m (CUDA cores, fp32), one row per (vector, beam) pair b:
//   r[b] = xn[v] - xhat[b];  u[b] = Wx . xhat[b];  idx[b] = A smallest of ||r[b] - S_m[k]||^2   (qinco_base.py:114-121)
// With step0 != 0 it is the first quantisation step instead (qinco_base.py:263, qinco_inference.py:239-246):
//   xhat[v][j] = C_0[code_j], hist[v][j][0] = code_j for the n_sel nearest codewords of xn[v].
struct PrepParams {
    int32_t D, De, K, A;      // A: number of candidates to keep (step0: F_1)
    int32_t F;                // beams per vector (row b -> v = b / F)
    int32_t step0;
    int32_t M;                // hist row length
    int64_t n_beams;
    const float* x;           // [n, D] raw input
    const float* mean;        // [D] or NULL
    float inv_std;            // DIVISOR applied after the mean shift: data_std, or 1 (name kept for the struct layout)
    const float* xhat;        // [n_beams, D] (unused for step0)
    const float* wx;          // [De][D] = Wcat[:, De:]  (NULL: skip u)
    const float* sub_cb;      // [K][D] pre-selection codebook (step0: C_0); NULL: skip selection
    const float* sub_norm;    // [K] squared norms of sub_cb's rows
    float* r;                 // [n_beams, D] or NULL
    float* u;                 // [n_beams, De]
    uint8_t* idx;             // [n_beams, A]
    float* xhat_out;          // step0: [n, A, D]
    uint8_t* hist_out;        // step0: [n, A, M]
    unsigned long long* sel_best;   // fused selection of the following score launch: reset to ~0 / 0 per vector (F == 1), or NULL
    uint32_t* sel_cnt;
};
cudaError_t launch_prep(const PrepParams& p, cudaStream_t stream);

// The same beam preparation on the tensor core (qb_prep_tc.cu): u and the pre-selection distances as tcgen05 GEMMs over
// fp16 hi/lo operand splits (fp32-level accuracy), top-A in the same kernel.  The weights come pre-packed (prep_pack):
// for every chunk of QB_PREP_DC input dimensions, for every part of <= QB_PREP_NP rows (row count padded to a multiple of
// 16 with zero rows): w_hi then w_lo, each as K-major core matrices [dc/8][np][8] fp16.  Not used for step 0.
#define QB_PREP_DC 64
#define QB_PREP_NP 256
struct PrepTcParams {
    int32_t D, De, K, K16, A;   // K16 = K rounded up to a multiple of 16
    int32_t F;                  // beams per vector
    int64_t n_beams;
    const float* x;             // [n, D] raw input
    const float* mean;          // [D] or NULL
    float std_div;              // divisor after the mean shift (data_std or 1)
    const float* xhat;          // [n_beams, D]
    const uint8_t* wx_pack;     // packed Wcat[:, De:] ([De][D]) or NULL: skip u
    const uint8_t* sub_pack;    // packed pre-selection codebook ([K16][D]) or NULL: no ranking
    const float* sub_norm;      // [K] squared row norms
    float* r;                   // [n_beams, D] or NULL
    float* u;                   // [n_beams, De]
    uint8_t* idx;               // [n_beams, A]
    unsigned long long* sel_best;   // fused selection state to reset (F == 1), or NULL
    uint32_t* sel_cnt;
    uint32_t* err_flag;
    int32_t dbg;                // unused (timing experiments)
    // step 0 (qinco_base.py:218,263; qinco_inference.py:239-246): xhat == NULL (zero), sub_pack / sub_norm = C_0, A = F_1 <= 16 beams:
    // xhat_out[v][a] = C_0[code_a], hist_out[v][a][0] = code_a
    int32_t step0, M;
    const float* cb0;           // [K][D] fp32
    float* xhat_out;            // [n, A, D]
    uint8_t* hist_out;          // [n, A, M]
};
cudaError_t launch_prep_tc(const PrepTcParams& p, cudaStream_t stream);

// Beam selection (qinco_base.py:346-372): per vector the F_out smallest of R = F_in*C distances, ascending;
// writes parent beam, code and the extended code history.
struct SelectParams {
    int32_t F_in, F_out, C, A, M, m;   // m = index of this step (history length before it)
    int64_t n;
    const float* dist;        // [n, F_in*C]
    const uint8_t* idx;       // [n*F_in, A] or NULL
    const uint8_t* hist_in;   // [n, F_in, M]
    uint8_t* hist_out;        // [n, F_out, M]
    uint8_t* sel_parent;      // [n, F_out]
    uint8_t* sel_code;        // [n, F_out]
};
cudaError_t launch_select(const SelectParams& p, cudaStream_t stream);

// IVF first step (reference IVFBook.quantize / encode, qinco_base.py:146-174): per vector the arg-min over ivf_K centroids
// of  (|x|^2 + |c|^2) - 2 x.c  (the reference's approx_pairwise_distance, utils.py:336-346), ties to the lower index;
// writes the code and the centroid as the single starting beam.
struct IvfParams {
    int32_t D, ivf_K;
    int64_t n;
    const float* x;           // [n, D] raw input
    const float* mean;        // [D] or NULL
    float std_div;            // divisor after the mean shift (data_std or 1)
    const float* cent;        // [ivf_K, D]
    const float* cnorm;       // [ivf_K]  |c|^2
    int32_t* codes_out;       // [n]
    float* xhat_out;          // [n, D]
};
cudaError_t launch_ivf_assign(const IvfParams& p, cudaStream_t stream);

// The same arg-min on the tensor core (qb_ivf_tc.cu) for D <= QB_IVF_TC_MAX_D: centroids pre-packed (ivf_pack) in parts of
// QB_IVF_NP: [c_hi | c_lo] as K-major fp16 core matrices [D/8][NP][8] each, then the NP squared norms (fp32; +max for the
// padding rows of the last part).
#define QB_IVF_NP 128
#define QB_IVF_TC_MAX_D 128
struct IvfTcParams {
    int32_t D, ivf_K;
    int64_t n;
    const float* x;
    const float* mean;
    float std_div;
    const uint8_t* cent_pack;  // [ceil(ivf_K / NP)] parts of (NP * D * 4 + NP * 4) bytes
    const float* cent;         // [ivf_K, D] fp32 (the lookup)
    int32_t* codes_out;
    float* xhat_out;
    uint32_t* err_flag;
};
cudaError_t launch_ivf_tc(const IvfTcParams& p, cudaStream_t stream);
// xhat[v] = centroids[ivf_codes[v]]  (IVFBook.decode, qinco_base.py:176-183)
cudaError_t launch_ivf_lookup(const float* cent, const int32_t* ivf_codes, int64_t n, int D, int ivf_K, float* xhat,
                              uint32_t* err_flag, cudaStream_t stream);

// xhat[v] = C_0[codes[v*M]]  (decode start, qinco_base.py:447-452 with step 0 = plain codebook lookup)
cudaError_t launch_decode_init(const float* cb0, const uint8_t* codes, int64_t n, int M, int D, int K, float* xhat,
                               uint32_t* err_flag, cudaStream_t stream);
// Code-matrix conversions at the reference's surface (qinco_base.py:447-449, :480-485): the reference moves codes as
// [S, n] integer matrices (S = M, or M + 1 with the IVF code in row 0; int64 from encode, int32 from the search's re-ranking
// loop, search_tasks.py:428-445), the kernels as uint8 [n, M] (+ int32 [n] IVF codes).  `pack` range-checks on the device
// (err word 0x10 / 0x20, the offending code is clamped) so the Python surface needs no reduction kernels and no host sync.
cudaError_t launch_codes_pack(const void* codes_MB, int elem_bytes, int64_t stride_row, int64_t stride_col, int64_t n, int M,
                              int K, int ivf_K, uint8_t* codes_u8, int32_t* ivf, uint32_t* err_flag, cudaStream_t stream);
cudaError_t launch_codes_unpack(const uint8_t* codes_u8, const int32_t* ivf, int64_t n, int M, int has_ivf, int64_t* codes_MB,
                                cudaStream_t stream);

// hist_out[v] = hist_in[v][0 .. m) ++ (code of sel_best[v]): the history update after a fused selection when no update launch
// follows (last step, codes only)
cudaError_t launch_take_codes(const unsigned long long* sel_best, int64_t n, const uint8_t* hist_in, uint8_t* hist_out, int M,
                              int m, cudaStream_t stream);

// out[i] = in[i]*scale + shift[i % D]
cudaError_t launch_affine(const float* in, float* out, int64_t n, int D, float scale, const float* shift,
                          cudaStream_t stream);

}  // namespace qb
