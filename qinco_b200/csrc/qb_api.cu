// C ABI of libqinco_b200.so (include/qinco_b200.h): model packing/upload and the per-chunk launch sequences of the
// encode / decode loops.
//
// Encode of one chunk of vectors (reference qinco/model/qinco_base.py:454-485 loop, :292-374 one beam step):
//   step 0      prep(step0)            F_1 nearest codewords of C_0 -> beam x-hat + code history
//   step m>=1   prep                   r = x - xhat_b, u_b = Wx.xhat_b, top-A pre-selection            (CUDA cores, fp32)
//               mlp(score)             dist[row] = ||r_b - f_m(C_m[code], xhat_b)||^2                  (tcgen05)
//               select                 F_out smallest per vector, parent beam, code history
//               mlp(apply)             xhat'_j = xhat_parent + f_m(C_m[code_j], xhat_parent)           (tcgen05)
// Decode (reference :447-452, :282-290): decode_init, then per step prep(u only) + mlp(apply).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/qinco_b200.h"
#include "qb_dev.h"
#include "qb_host.h"

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}
int fail_cuda(cudaError_t e, const char* what) {
    g_err = std::string(what) + ": " + cudaGetErrorString(e);
    return QB_ERR_CUDA;
}
#define QB_CUDA(call)                                        \
    do {                                                     \
        cudaError_t e__ = (call);                            \
        if (e__ != cudaSuccess) return fail_cuda(e__, #call); \
    } while (0)

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

float row_norm2(const float* v, int n) {       // fp32 like the reference's (b ** 2).sum(-1)
    float s = 0.f;
    for (int i = 0; i < n; i++) s += v[i] * v[i];
    return s;
}

struct StepDev {
    QbStepPlan plan;
    std::vector<QbOp> ops;
    uint8_t* w_blob = nullptr;
    float* t_blk = nullptr;
    float* cb_blk = nullptr;
    float* wx = nullptr;         // [De][D] = Wcat[:, De:]
    float* sub_cb = nullptr;     // [K][D] pre-selection codebook (A > 0)
    float* sub_norm = nullptr;   // [K] squared row norms of sub_cb
    uint8_t* wx_pack = nullptr;  // tensor-core beam preparation: fp16 hi/lo operand parts of wx / sub_cb (qb_prep_tc.cu)
    uint8_t* sub_pack = nullptr;
    // decode loop (one launch walks every step): the same weight blob, extended by the pre-op slabs (u = Wx . xhat)
    QbStepPlan loop_plan;
    std::vector<QbOp> loop_ops;
};

struct HostSlot {
    float* x_pin = nullptr;
    uint8_t* codes_pin = nullptr;
    float* out_pin = nullptr;     // xhat (encode) or decoded vectors (decode)
    float* x_dev = nullptr;
    uint8_t* codes_dev = nullptr;
    float* out_dev = nullptr;
    int32_t* ivf_pin = nullptr;   // IVF models: the int32 IVF codes of the chunk
    int32_t* ivf_dev = nullptr;
    cudaEvent_t ev_h2d = nullptr, ev_comp = nullptr, ev_d2h = nullptr;
    int64_t pending_i0 = -1, pending_n = 0;
};

}  // namespace

struct qb_model {
    int D = 0, De = 0, Dh = 0, L = 0, M = 0, K = 0, A = 0, B = 0, q1 = 0, device = 0, n_sm = 0;
    // M = uint8 codes per vector.  S = quantisation steps: M, or M + 1 with an IVF first step (then step s >= 1 writes
    // code column s - 1 and the IVF code lives in its own int32 array).
    int S = 0, ivf_K = 0;
    uint8_t* ivf_pack = nullptr;  // tensor-core arg-min operand parts (qb_ivf_tc.cu), or NULL: the fp32 CUDA-core kernel
    float* ivf_cent = nullptr;    // [ivf_K][D]
    float* ivf_cnorm = nullptr;   // [ivf_K]
    int col(int step) const { return ivf_K ? step - 1 : step; }
    float data_std = 1.f;
    int stagger = 0;
    bool has_mean = false;
    float* cb0 = nullptr;      // [K][D]
    float* cb0_norm = nullptr; // [K] squared row norms of cb0
    uint8_t* cb0_pack = nullptr;  // C_0 as tensor-core operand parts (qb_prep_tc.cu), step 0
    float* mean = nullptr;     // [D]
    std::vector<StepDev> steps;   // index m, entry 0 unused
    bool fuse_ok = true;          // fused beam selection inside the score launch (QB_NO_FUSE=1 keeps the unfused sequence for A/B runs)
    bool prep_tc = true;          // beam preparation on the tensor core (QB_PREP_CC=1: the fp32 CUDA-core kernel, for A/B runs)
    int fuse_mode = 1;            // 1: arg-min in the score launch, xhat' by a 1/256-size update launch (default);
                                  // 2 (QB_FUSE_FULL=1): the score launch also writes xhat' (measured slower)
    bool loop_ok = false;         // every step has a decode-loop plan (qb_mlp_kernel<.., kLoop>): decode is ONE launch
    uint32_t* err_host = nullptr;   // mapped pinned word written by the kernels before they trap
    uint32_t* err_dev = nullptr;
    int64_t launches = 0;
    std::vector<void*> dev_allocs;
    // optional per-kernel timing (qb_timing_*): CUDA events recorded around every launch on the launch stream
    bool timing = false;
    struct Timed { cudaEvent_t a, b; int kind; int64_t rows; };
    std::vector<Timed> timed;
    std::vector<cudaEvent_t> ev_pool;
    // host-variant staging
    int64_t host_chunk = 0;
    HostSlot slot[2];
    void* host_ws = nullptr;
    size_t host_ws_bytes = 0;
    cudaStream_t s_h2d = nullptr, s_comp = nullptr, s_d2h = nullptr;
};

namespace {

template <typename T>
int dev_upload(qb_model* m, const T* host, size_t count, T** out) {
    void* p = nullptr;
    QB_CUDA(cudaMalloc(&p, std::max<size_t>(count * sizeof(T), 16)));
    m->dev_allocs.push_back(p);
    if (count) QB_CUDA(cudaMemcpy(p, host, count * sizeof(T), cudaMemcpyHostToDevice));
    *out = reinterpret_cast<T*>(p);
    return QB_OK;
}

int check_desc(const qb_model_desc* d) {
    if (!d) return fail(QB_ERR_INVALID, "desc is NULL");
    if (d->M < 1 || d->M > 64) return fail(QB_ERR_INVALID, "M must be in [1,64]");
    if (d->K < 1 || d->K > 256) return fail(QB_ERR_INVALID, "K must be in [1,256] (codes are uint8)");
    if (d->D < 16 || d->D % 16) return fail(QB_ERR_INVALID, "D must be a positive multiple of 16");
    if (d->A < 0 || d->A > d->K) return fail(QB_ERR_INVALID, "A must be in [0,K]");
    if (d->B < 1 || d->B > d->K || d->B > 255) return fail(QB_ERR_INVALID, "B must be in [1,min(K,255)]");
    if (d->L < 0) return fail(QB_ERR_INVALID, "L must be >= 0");
    if (!(d->data_std > 0.f)) return fail(QB_ERR_INVALID, "data_std must be > 0 (qinco_base.py:526)");
    if (!d->codebook) return fail(QB_ERR_INVALID, "codebook pointers missing");
    if (d->ivf_K < 0) return fail(QB_ERR_INVALID, "ivf_K must be >= 0");
    if (d->ivf_K > 0 && !d->ivf_centroids) return fail(QB_ERR_INVALID, "ivf_K > 0 needs ivf_centroids");
    const int S = d->ivf_K > 0 ? d->M + 1 : d->M;
    for (int m = d->ivf_K > 0 ? 1 : 0; m < S; m++)
        if (!d->codebook[m]) return fail(QB_ERR_INVALID, "codebook[m] is NULL");
    if (S > 1) {
        if (!d->concat_w || !d->concat_b) return fail(QB_ERR_INVALID, "concat weights missing");
        if (d->L > 0 && (!d->up_w || !d->down_w)) return fail(QB_ERR_INVALID, "residual block weights missing");
        if (d->A > 0 && !d->substep_codebook) return fail(QB_ERR_INVALID, "A > 0 needs substep codebooks");
        if (d->De != d->D && (!d->in_proj || !d->out_proj)) return fail(QB_ERR_INVALID, "de != D needs in_proj/out_proj");
        for (int m = 1; m < S; m++) {
            if (!d->concat_w[m] || !d->concat_b[m]) return fail(QB_ERR_INVALID, "concat weights of a step are NULL");
            if (d->A > 0 && !d->substep_codebook[m]) return fail(QB_ERR_INVALID, "substep codebook of a step is NULL");
            if (d->De != d->D && (!d->in_proj[m] || !d->out_proj[m])) return fail(QB_ERR_INVALID, "projection of a step is NULL");
            for (int l = 0; l < d->L; l++)
                if (!d->up_w[m * d->L + l] || !d->down_w[m * d->L + l]) return fail(QB_ERR_INVALID, "block weights of a step are NULL");
        }
    }
    return QB_OK;
}

// ---- per-chunk workspace ----------------------------------------------------------------------------------------
struct Workspace {
    float* xhat[2];
    uint8_t* hist[2];
    float* r;
    float* u;
    uint8_t* idx;
    float* dist;
    uint8_t* selp;
    uint8_t* selc;
    unsigned long long* sel_best;   // fused selection state, one entry per vector
    uint32_t* sel_cnt;
};

size_t encode_bytes_per_vector(const qb_model* m) {
    const size_t B = m->B, C = m->A > 0 ? m->A : m->K;
    return 2 * B * m->D * 4 + 2 * B * m->M + B * m->D * 4 + B * m->De * 4 + B * std::max(m->A, 1) + B * C * 4 + 2 * B + 12;
}
constexpr size_t kWsSlack = 20 * 256;   // alignment padding of the carve-up

int64_t default_chunk(const qb_model* m) {
    const int64_t rows_per_vec = (int64_t)m->B * (m->A > 0 ? m->A : m->K);
    static const int64_t rows_env = getenv("QB_CHUNK_ROWS") ? atoll(getenv("QB_CHUNK_ROWS")) : 0;   // tuning knob
    int64_t c = (rows_env > 0 ? rows_env : (int64_t)(16 << 20)) / rows_per_vec;
    c = std::max<int64_t>(c, 1024);
    return c / 128 * 128;
}

void carve(const qb_model* m, void* ws, int64_t nc, Workspace* w) {
    uint8_t* p = reinterpret_cast<uint8_t*>(align_up(reinterpret_cast<size_t>(ws), 256));
    auto take = [&](size_t bytes) {
        uint8_t* q = p;
        p += align_up(bytes, 256);
        return q;
    };
    const size_t B = m->B, C = m->A > 0 ? m->A : m->K, n = (size_t)nc;
    w->xhat[0] = (float*)take(n * B * m->D * 4);
    w->xhat[1] = (float*)take(n * B * m->D * 4);
    w->hist[0] = take(n * B * m->M);
    w->hist[1] = take(n * B * m->M);
    w->r = (float*)take(n * B * m->D * 4);
    w->u = (float*)take(n * B * m->De * 4);
    w->idx = take(n * B * std::max(m->A, 1));
    w->dist = (float*)take(n * B * C * 4);
    w->selp = take(n * B);
    w->selc = take(n * B);
    w->sel_best = (unsigned long long*)take(n * 8);
    w->sel_cnt = (uint32_t*)take(n * 4);
}

enum { KIND_PREP = 0, KIND_SCORE = 1, KIND_SELECT = 2, KIND_APPLY = 3, KIND_OTHER = 4, KIND_IVF = 5, KIND_COUNT = 6 };

cudaEvent_t take_event(qb_model* m) {
    if (!m->ev_pool.empty()) {
        cudaEvent_t e = m->ev_pool.back();
        m->ev_pool.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

// Runs `launch` (returns cudaError_t) and counts it; with timing on, brackets it with events on `st`.
template <typename F>
cudaError_t timed_launch(qb_model* m, int kind, int64_t rows, cudaStream_t st, F&& launch) {
    m->launches++;
    if (!m->timing) return launch();
    qb_model::Timed t{take_event(m), take_event(m), kind, rows};
    cudaEventRecord(t.a, st);
    cudaError_t e = launch();
    cudaEventRecord(t.b, st);
    m->timed.push_back(t);
    return e;
}

qb::MlpParams base_mlp(const qb_model* m, int step) {
    const StepDev& s = m->steps[step];
    qb::MlpParams p;
    std::memset(&p, 0, sizeof(p));
    p.plan = s.plan;
    p.n_ops = (int32_t)s.ops.size();
    std::memcpy(p.ops, s.ops.data(), s.ops.size() * sizeof(QbOp));
    p.w_blob = s.w_blob;
    p.t_blk = s.t_blk;
    p.cb_blk = s.cb_blk;
    p.out_scale = 1.f;
    p.err_flag = m->err_dev;
    return p;
}

int encode_chunk(qb_model* m, const float* x, int64_t n, int normalize, int32_t* ivf_codes, uint8_t* codes, float* xhat_out,
                 void* ws, cudaStream_t st) {
    Workspace w;
    carve(m, ws, n, &w);
    const int D = m->D, M = m->M, K = m->K, A = m->A, B = m->B;
    const int C = A > 0 ? A : K;
    const float* mean = (normalize && m->has_mean) ? m->mean : nullptr;
    const float inv_std_div = normalize ? m->data_std : 1.f;
    int cur = 0;
    const int S = m->S;
    // ---- step 0 (qinco_base.py:218,263; qinco_inference.py:239-246), or the IVF arg-min (IVFBook.encode, :165-174; F = 1)
    const int F1 = m->ivf_K ? 1 : ((M == 1) ? 1 : B);
    if (m->ivf_K && m->ivf_pack) {
        qb::IvfTcParams p;
        std::memset(&p, 0, sizeof(p));
        p.D = D; p.ivf_K = m->ivf_K; p.n = n; p.x = x; p.mean = mean; p.std_div = inv_std_div;
        p.cent_pack = m->ivf_pack; p.cent = m->ivf_cent; p.codes_out = ivf_codes; p.xhat_out = w.xhat[cur]; p.err_flag = m->err_dev;
        QB_CUDA(timed_launch(m, KIND_IVF, n, st, [&] { return qb::launch_ivf_tc(p, st); }));
    } else if (m->ivf_K) {
        qb::IvfParams p;
        std::memset(&p, 0, sizeof(p));
        p.D = D; p.ivf_K = m->ivf_K; p.n = n; p.x = x; p.mean = mean; p.std_div = inv_std_div;
        p.cent = m->ivf_cent; p.cnorm = m->ivf_cnorm; p.codes_out = ivf_codes; p.xhat_out = w.xhat[cur];
        QB_CUDA(timed_launch(m, KIND_IVF, n, st, [&] { return qb::launch_ivf_assign(p, st); }));
    } else if (m->prep_tc && F1 <= 16) {     // step 0 on the tensor core: distances to C_0, the F_1 nearest start the beams
        qb::PrepTcParams p;
        std::memset(&p, 0, sizeof(p));
        p.D = D; p.De = m->De; p.K = K; p.K16 = (K + 15) / 16 * 16; p.A = F1; p.F = 1; p.step0 = 1; p.M = M;
        p.n_beams = n; p.x = x; p.mean = mean; p.std_div = inv_std_div;
        p.sub_pack = m->cb0_pack; p.sub_norm = m->cb0_norm; p.cb0 = m->cb0; p.err_flag = m->err_dev;
        p.xhat_out = (M == 1 && xhat_out) ? xhat_out : w.xhat[cur];
        p.hist_out = (M == 1) ? codes : w.hist[cur];
        QB_CUDA(timed_launch(m, KIND_PREP, n, st, [&] { return qb::launch_prep_tc(p, st); }));
    } else {
        qb::PrepParams p;
        std::memset(&p, 0, sizeof(p));
        p.D = D; p.De = m->De; p.K = K; p.A = F1; p.F = 1; p.step0 = 1; p.M = M;
        p.n_beams = n; p.x = x; p.mean = mean; p.inv_std = inv_std_div;
        p.sub_cb = m->cb0; p.sub_norm = m->cb0_norm;
        p.xhat_out = (M == 1 && xhat_out) ? xhat_out : w.xhat[cur];
        p.hist_out = (M == 1) ? codes : w.hist[cur];
        QB_CUDA(timed_launch(m, KIND_PREP, n, st, [&] { return qb::launch_prep(p, st); }));
    }
    int F_in = F1;
    for (int step = 1; step < S; step++) {
        const StepDev& s = m->steps[step];
        const int F_out = (step < S - 1) ? B : 1;
        const bool last = step == S - 1;
        // the first implicit-codebook step after an IVF step starts from ONE beam and must still supply B of them:
        // it pre-selects max(A, B) candidates (QincoSubstep._n_codes, qinco_base.py:108-112)
        const int A = (m->A > 0 && m->ivf_K && step == 1) ? std::max(m->A, B) : m->A;
        const int C = A > 0 ? A : K;
        // Fused selection (the score launch picks the winners and writes xhat' / the history itself): resident launches
        // with one beam per vector, i.e. A == 0, beam 1 on K = 256 models whose tables fit the SM (QINCo1 / QINCo2-S shapes).
        const bool fuse = m->fuse_ok && A == 0 && K == 256 && F_in == 1 && F_out == 1 && s.plan.smem_tres >= 0 &&
                          s.plan.n_tiles == 2 && !s.plan.pair && m->n_sm >= 4;
        if (m->prep_tc) {       // u and the pre-selection distances as tensor-core GEMMs (fp16 hi/lo splits), top-A in the kernel
            qb::PrepTcParams p;
            std::memset(&p, 0, sizeof(p));
            p.D = D; p.De = m->De; p.K = K; p.K16 = (K + 15) / 16 * 16; p.A = A; p.F = F_in;
            p.n_beams = n * F_in; p.x = x; p.mean = mean; p.std_div = inv_std_div;
            p.xhat = w.xhat[cur]; p.wx_pack = s.wx_pack; p.sub_pack = A > 0 ? s.sub_pack : nullptr; p.sub_norm = s.sub_norm;
            p.r = w.r; p.u = w.u; p.idx = w.idx; p.err_flag = m->err_dev;
            if (fuse) { p.sel_best = w.sel_best; p.sel_cnt = w.sel_cnt; }
            QB_CUDA(timed_launch(m, KIND_PREP, p.n_beams, st, [&] { return qb::launch_prep_tc(p, st); }));
        } else {
            qb::PrepParams p;
            std::memset(&p, 0, sizeof(p));
            p.D = D; p.De = m->De; p.K = K; p.A = A; p.F = F_in; p.step0 = 0; p.M = M;
            p.n_beams = n * F_in; p.x = x; p.mean = mean; p.inv_std = inv_std_div;
            p.xhat = w.xhat[cur]; p.wx = s.wx; p.sub_cb = A > 0 ? s.sub_cb : nullptr; p.sub_norm = s.sub_norm;
            p.r = w.r; p.u = w.u; p.idx = w.idx;
            if (fuse) { p.sel_best = w.sel_best; p.sel_cnt = w.sel_cnt; }
            QB_CUDA(timed_launch(m, KIND_PREP, p.n_beams, st, [&] { return qb::launch_prep(p, st); }));
        }
        // ... and for the one-tile-per-CTA shapes with a single out_proj chunk (QINCo2-L family) the whole selection: running
        // top-F_out per vector and the winners' xhat' inside the score launch (qb_mlp_kernel<.., kFuse = 3>)
        const int R = F_in * C;
        const bool fuse_b = m->fuse_ok && !fuse && s.plan.n_tiles == 1 && s.plan.has_proj && s.plan.n_ochunk == 1 && !s.plan.pair &&
                            s.plan.smem_total <= 196608 &&
                            R >= 4 && ((R < QB_TILE_M && QB_TILE_M % R == 0) || R % QB_TILE_M == 0) && R / QB_TILE_M <= 255 &&
                            F_out <= 32 && F_out <= R && (QB_TILE_M / std::min(R, QB_TILE_M)) * F_out * D <= 4096;
        {
            qb::MlpParams p = base_mlp(m, step);
            p.mode = qb::QB_MODE_SCORE;
            p.C = C; p.A = A;
            p.n_rows = n * F_in * C;
            p.idx = w.idx; p.u = w.u; p.r = w.r; p.dist = w.dist;
            if (fuse_b) {
                p.fuse = 3; p.F_in = F_in; p.F_out = F_out;
                static const int fb_dbg = getenv("QB_FUSEB_DEBUG") ? atoi(getenv("QB_FUSEB_DEBUG")) : 0;
                p.dbg = fb_dbg;
                p.sel_spv = R > QB_TILE_M ? R / QB_TILE_M : 1;
                p.hist_in = w.hist[cur]; p.hist_out = last ? codes : w.hist[cur ^ 1]; p.hist_M = M; p.hist_m = m->col(step);
                p.xhat_in = w.xhat[cur];
                p.xhat_out = (last && xhat_out) ? xhat_out : w.xhat[cur ^ 1];
                p.dist = nullptr;
            }
            if (fuse) {
                p.fuse = m->fuse_mode; p.F_in = F_in; p.F_out = F_out;
                p.sel_best = w.sel_best; p.sel_cnt = w.sel_cnt;
                p.dist = nullptr;
                if (m->fuse_mode == 2) {    // the score launch also writes xhat' and the history
                    p.hist_in = w.hist[cur]; p.hist_out = last ? codes : w.hist[cur ^ 1]; p.hist_M = M; p.hist_m = m->col(step);
                    p.xhat_in = w.xhat[cur];
                    p.xhat_out = last ? xhat_out : w.xhat[cur ^ 1];      // NULL at the last step when the caller wants codes only
                }
            }
            QB_CUDA(timed_launch(m, KIND_SCORE, p.n_rows, st, [&] { return qb::launch_mlp(p, m->n_sm, st); }));
        }
        if (fuse && m->fuse_mode == 1) {
            // the winners are in sel_best: a 1/256-size update launch recomputes their f_m, writes xhat' and the history
            if (!last || xhat_out) {
                qb::MlpParams p = base_mlp(m, step);
                p.mode = qb::QB_MODE_APPLY;
                p.F_in = 1; p.F_out = 1; p.n_rows = n;
                p.sel_best = w.sel_best;
                p.hist_in = w.hist[cur]; p.hist_out = last ? codes : w.hist[cur ^ 1]; p.hist_M = M; p.hist_m = m->col(step);
                p.u = w.u; p.xhat_in = w.xhat[cur];
                p.xhat_out = last ? xhat_out : w.xhat[cur ^ 1];
                QB_CUDA(timed_launch(m, KIND_APPLY, p.n_rows, st, [&] { return qb::launch_mlp(p, m->n_sm, st); }));
            } else {
                QB_CUDA(timed_launch(m, KIND_SELECT, n, st, [&] {
                    return qb::launch_take_codes(w.sel_best, n, w.hist[cur], codes, M, m->col(step), st);
                }));
            }
        }
        if (fuse || fuse_b) {
            cur ^= 1;
            F_in = F_out;
            continue;
        }
        {
            qb::SelectParams p;
            std::memset(&p, 0, sizeof(p));
            p.F_in = F_in; p.F_out = F_out; p.C = C; p.A = A; p.M = M; p.m = m->col(step); p.n = n;
            p.dist = w.dist; p.idx = A > 0 ? w.idx : nullptr;
            p.hist_in = w.hist[cur]; p.hist_out = last ? codes : w.hist[cur ^ 1];
            p.sel_parent = w.selp; p.sel_code = w.selc;
            QB_CUDA(timed_launch(m, KIND_SELECT, n, st, [&] { return qb::launch_select(p, st); }));
        }
        if (!last || xhat_out) {
            qb::MlpParams p = base_mlp(m, step);
            p.mode = qb::QB_MODE_APPLY;
            p.F_in = F_in; p.F_out = F_out;
            p.n_rows = n * F_out;
            p.sel_parent = w.selp; p.sel_code = w.selc; p.code_stride = 1; p.code_off = 0;
            p.u = w.u; p.xhat_in = w.xhat[cur];
            p.xhat_out = last ? xhat_out : w.xhat[cur ^ 1];
            QB_CUDA(timed_launch(m, KIND_APPLY, p.n_rows, st, [&] { return qb::launch_mlp(p, m->n_sm, st); }));
        }
        cur ^= 1;
        F_in = F_out;
    }
    return QB_OK;
}

size_t decode_bytes_per_vector(const qb_model* m) { return (size_t)2 * m->D * 4 + (size_t)m->De * 4; }

int decode_chunk(qb_model* m, const int32_t* ivf_codes, const uint8_t* codes, int64_t n, int denormalize, float* out, void* ws,
                 cudaStream_t st) {
    uint8_t* p0 = reinterpret_cast<uint8_t*>(align_up(reinterpret_cast<size_t>(ws), 256));
    float* xh[2];
    xh[0] = (float*)p0;
    xh[1] = (float*)(p0 + align_up((size_t)n * m->D * 4, 256));
    float* u = (float*)(p0 + 2 * align_up((size_t)n * m->D * 4, 256));
    const int D = m->D, M = m->M;
    const float scale = denormalize ? m->data_std : 1.f;
    const float* shift = (denormalize && m->has_mean) ? m->mean : nullptr;
    const bool affine = denormalize && (scale != 1.f || shift);
    int cur = 0;
    const int S = m->S;
    if (m->loop_ok && S > 1) {
        // ONE launch: every 128-vector tile walks all S - 1 implicit-codebook steps (qb_mlp_kernel<.., kLoop>), codes read
        // in the kernel, u = Wx . xhat on the tensor core, the running xhat kept by the rows' own threads in `out`.
        const StepDev& s1 = m->steps[1];
        qb::MlpParams p;
        std::memset(&p, 0, sizeof(p));
        p.plan = s1.loop_plan;
        p.n_ops = (int32_t)s1.loop_ops.size();
        std::memcpy(p.ops, s1.loop_ops.data(), s1.loop_ops.size() * sizeof(QbOp));
        p.w_blob = s1.w_blob; p.t_blk = s1.t_blk; p.cb_blk = s1.cb_blk;
        p.err_flag = m->err_dev;
        p.mode = qb::QB_MODE_APPLY;
        p.F_in = 1; p.F_out = 1; p.n_rows = n;
        p.sel_code = codes; p.code_stride = M; p.code_off = m->ivf_K ? 0 : 1;
        p.xhat_out = out;
        p.out_scale = scale; p.out_shift = shift;
        p.n_loop_steps = S - 1;
        for (int step = 1; step < S; step++)
            p.loop_steps[step - 1] = qb::QbLoopStep{m->steps[step].w_blob, m->steps[step].t_blk, m->steps[step].cb_blk};
        if (m->ivf_K) { p.seed_tab = m->ivf_cent; p.seed_K = m->ivf_K; p.seed_codes_i32 = ivf_codes; }
        else { p.seed_tab = m->cb0; p.seed_K = m->K; }
        QB_CUDA(timed_launch(m, KIND_APPLY, n * (S - 1), st, [&] { return qb::launch_mlp(p, m->n_sm, st); }));
        return QB_OK;
    }
    if (m->ivf_K)
        QB_CUDA(timed_launch(m, KIND_OTHER, n, st, [&] {
            return qb::launch_ivf_lookup(m->ivf_cent, ivf_codes, n, D, m->ivf_K, xh[cur], m->err_dev, st);
        }));
    else
        QB_CUDA(timed_launch(m, KIND_OTHER, n, st, [&] {
            return qb::launch_decode_init(m->cb0, codes, n, M, D, m->K, (M == 1 && !affine) ? out : xh[cur], m->err_dev, st);
        }));
    if (M == 1 && affine && !m->ivf_K)     // (an IVF model with M == 1 still has one MLP step, which applies the affine)
        QB_CUDA(timed_launch(m, KIND_OTHER, n, st, [&] { return qb::launch_affine(xh[cur], out, n, D, scale, shift, st); }));
    for (int step = 1; step < S; step++) {
        const StepDev& s = m->steps[step];
        const bool last = step == S - 1;
        {
            qb::PrepParams p;
            std::memset(&p, 0, sizeof(p));
            p.D = D; p.De = m->De; p.K = m->K; p.A = 0; p.F = 1; p.step0 = 0; p.M = M;
            p.n_beams = n; p.x = xh[cur]; p.inv_std = 1.f;
            p.xhat = xh[cur]; p.wx = s.wx; p.u = u;
            QB_CUDA(timed_launch(m, KIND_PREP, n, st, [&] { return qb::launch_prep(p, st); }));
        }
        {
            qb::MlpParams p = base_mlp(m, step);
            p.mode = qb::QB_MODE_APPLY;
            p.F_in = 1; p.F_out = 1; p.n_rows = n;
            p.sel_code = codes; p.code_stride = M; p.code_off = m->col(step);
            p.u = u; p.xhat_in = xh[cur];
            p.xhat_out = last ? out : xh[cur ^ 1];
            if (last) { p.out_scale = scale; p.out_shift = shift; }
            QB_CUDA(timed_launch(m, KIND_APPLY, n, st, [&] { return qb::launch_mlp(p, m->n_sm, st); }));
        }
        cur ^= 1;
    }
    return QB_OK;
}

int check_device_flag(qb_model* m) {
    volatile uint32_t* word = reinterpret_cast<volatile uint32_t*>(m->err_host);
    const uint32_t e = *word;
    if (e) {
        char buf[200];
        snprintf(buf, sizeof(buf),
                 "device-side failure, err word 0x%x (0x10: code out of range [0,K); 0x20: IVF code out of range; "
                 "0x1xx/0x2xx/0x3xx/0x4xx: barrier wait time-out in producer/MMA/ring/epilogue)", e);
        // Out-of-range codes are the caller's input error: the kernels clamp them and run to completion, so the word is
        // reported ONCE and cleared -- later calls on the model are good again.  Time-outs precede a trap: the context is
        // gone and the word stays.
        if (e == 0x10u || e == 0x20u) *word = 0;
        return fail(QB_ERR_KERNEL, buf);
    }
    return QB_OK;
}

int ensure_host_staging(qb_model* m, bool decode) {
    (void)decode;
    if (m->host_chunk) return QB_OK;
    const int64_t nc = default_chunk(m);
    QB_CUDA(cudaStreamCreateWithFlags(&m->s_h2d, cudaStreamNonBlocking));
    QB_CUDA(cudaStreamCreateWithFlags(&m->s_comp, cudaStreamNonBlocking));
    QB_CUDA(cudaStreamCreateWithFlags(&m->s_d2h, cudaStreamNonBlocking));
    for (auto& s : m->slot) {
        QB_CUDA(cudaHostAlloc((void**)&s.x_pin, (size_t)nc * m->D * 4, cudaHostAllocDefault));
        QB_CUDA(cudaHostAlloc((void**)&s.out_pin, (size_t)nc * m->D * 4, cudaHostAllocDefault));
        QB_CUDA(cudaHostAlloc((void**)&s.codes_pin, (size_t)nc * m->M, cudaHostAllocDefault));
        QB_CUDA(cudaMalloc((void**)&s.x_dev, (size_t)nc * m->D * 4));
        QB_CUDA(cudaMalloc((void**)&s.out_dev, (size_t)nc * m->D * 4));
        QB_CUDA(cudaMalloc((void**)&s.codes_dev, (size_t)nc * m->M));
        if (m->ivf_K) {
            QB_CUDA(cudaHostAlloc((void**)&s.ivf_pin, (size_t)nc * 4, cudaHostAllocDefault));
            QB_CUDA(cudaMalloc((void**)&s.ivf_dev, (size_t)nc * 4));
        }
        QB_CUDA(cudaEventCreateWithFlags(&s.ev_h2d, cudaEventDisableTiming));
        QB_CUDA(cudaEventCreateWithFlags(&s.ev_comp, cudaEventDisableTiming));
        QB_CUDA(cudaEventCreateWithFlags(&s.ev_d2h, cudaEventDisableTiming));
    }
    m->host_ws_bytes = std::max(qb_encode_workspace_bytes(m, nc), qb_decode_workspace_bytes(m, nc));
    QB_CUDA(cudaMalloc(&m->host_ws, m->host_ws_bytes));
    m->host_chunk = nc;
    return QB_OK;
}

}  // namespace

extern "C" {

int qb_version(void) { return 100; }
const char* qb_last_error(void) { return g_err.c_str(); }

int qb_model_create(const qb_model_desc* d, qb_model** out) {
    if (!out) return fail(QB_ERR_INVALID, "out is NULL");
    *out = nullptr;
    int rc = check_desc(d);
    if (rc) return rc;
    int ndev = 0;
    QB_CUDA(cudaGetDeviceCount(&ndev));
    if (d->device < 0 || d->device >= ndev) return fail(QB_ERR_INVALID, "device ordinal out of range");
    qb::DeviceGuard guard(d->device);       // the caller's current device is restored on every exit
    QB_CUDA(guard.err);
    cudaDeviceProp prop;
    QB_CUDA(cudaGetDeviceProperties(&prop, d->device));
    if (prop.major != 10)
        return fail(QB_ERR_CUDA, std::string("this library is built for sm_100a only; device is ") + prop.name);

    qb_model* m = new (std::nothrow) qb_model();
    if (!m) return fail(QB_ERR_NOMEM, "out of host memory");
    m->D = d->D; m->De = d->De > 0 ? d->De : d->D; m->Dh = d->Dh; m->L = d->L; m->M = d->M; m->K = d->K;
    m->A = d->A; m->B = d->B; m->q1 = d->qinco1_mode; m->device = d->device; m->n_sm = prop.multiProcessorCount;
    m->data_std = d->data_std;
    m->ivf_K = d->ivf_K;
    m->S = d->ivf_K > 0 ? d->M + 1 : d->M;
    m->fuse_ok = getenv("QB_NO_FUSE") == nullptr;
    m->fuse_mode = getenv("QB_FUSE_FULL") ? 2 : 1;
    m->prep_tc = getenv("QB_PREP_CC") == nullptr;
    auto bail = [&](int code) {
        qb_model_destroy(m);
        return code;
    };
    const int D = m->D, De = m->De, K = m->K;

    QB_CUDA(cudaHostAlloc((void**)&m->err_host, 64, cudaHostAllocMapped));
    *m->err_host = 0;
    QB_CUDA(cudaHostGetDevicePointer((void**)&m->err_dev, m->err_host, 0));

    if (m->ivf_K) {     // step 0: frozen IVF centroids and their squared norms (IVFBook, qinco_base.py:128-146)
        std::vector<float> cn((size_t)m->ivf_K);
        for (int k = 0; k < m->ivf_K; k++) cn[k] = row_norm2(d->ivf_centroids + (size_t)k * D, D);
        if ((rc = dev_upload(m, d->ivf_centroids, (size_t)m->ivf_K * D, &m->ivf_cent))) return bail(rc);
        if ((rc = dev_upload(m, cn.data(), cn.size(), &m->ivf_cnorm))) return bail(rc);
        if (D <= QB_IVF_TC_MAX_D && !getenv("QB_IVF_CC")) {       // (QB_IVF_CC=1: keep the fp32 CUDA-core arg-min, for A/B runs)
            std::vector<uint8_t> pk(qb::ivf_pack_bytes(m->ivf_K, D));
            qb::ivf_pack(d->ivf_centroids, m->ivf_K, D, pk.data());
            if ((rc = dev_upload(m, pk.data(), pk.size(), &m->ivf_pack))) return bail(rc);
        }
    } else {            // step 0: plain codebook and its squared row norms
        std::vector<float> nrm((size_t)K);
        for (int k = 0; k < K; k++) nrm[k] = row_norm2(d->codebook[0] + (size_t)k * D, D);
        if ((rc = dev_upload(m, d->codebook[0], (size_t)K * D, &m->cb0))) return bail(rc);
        if ((rc = dev_upload(m, nrm.data(), nrm.size(), &m->cb0_norm))) return bail(rc);
        std::vector<uint16_t> pk(qb::prep_pack_bytes(K, D) / 2);
        qb::prep_pack(d->codebook[0], K, D, pk.data());
        if ((rc = dev_upload(m, (const uint8_t*)pk.data(), pk.size() * 2, &m->cb0_pack))) return bail(rc);
    }
    if (d->data_mean) {
        bool nz = false;
        for (int i = 0; i < D; i++) nz |= d->data_mean[i] != 0.f;
        m->has_mean = nz;
        if ((rc = dev_upload(m, d->data_mean, (size_t)D, &m->mean))) return bail(rc);
    }
    m->steps.resize(m->S);
    {   // default stagger: ~1.5x the ideal MMA time of one tile (128 rows x 2*L*De*Dh MACs at ~3800 MAC/cycle)
        const double mma = 128.0 * 2.0 * m->L * m->De * m->Dh / 3800.0;
        m->stagger = d->opt_stagger > 0 ? d->opt_stagger : 0;
        (void)mma;
    }
    qb::PlanOptions opt;
    opt.hc = d->opt_hc; opt.n_tiles = d->opt_n_tiles & 0xff; opt.pair = (d->opt_n_tiles >> 8) & 0xff; opt.mcast = (d->opt_n_tiles >> 16) & 0xff;
    if (opt.mcast == 0)
        if (const char* e = getenv("QB_MCAST")) opt.mcast = atoi(e);  // debug / A-B runs: 1 = off, 2 = on
    opt.slot_bytes = d->opt_slot_bytes;
    if (opt.pair == 0)
        if (const char* e = getenv("QB_PAIR")) opt.pair = atoi(e);   // debug / A-B runs: 1 = single-CTA kernel, 2 = force pairs
    opt.max_stage = d->opt_max_stage & 0xff; opt.no_resident = (d->opt_max_stage >> 8) & 1; opt.no_hsplit = (d->opt_max_stage >> 9) & 1; opt.max_slab_k = d->opt_max_slab_k;
    opt.no_esplit = ((d->opt_max_stage >> 11) & 1) || getenv("QB_NO_ESPLIT") != nullptr;
    opt.blk32 = getenv("QB_BLK32") != nullptr;
    int max_smem = 0;
    for (int s = 1; s < m->S; s++) {
        StepDev& sd = m->steps[s];
        std::vector<QbOp> ops;
        std::string err;
        if (qb::make_step_plan(D, De, m->Dh, m->L, K, m->q1, opt, &sd.plan, &ops, &err)) return bail(fail(QB_ERR_INVALID, err));
        // Decode-loop plan: same slab geometry (slot size, H chunk), plus the pre-ops; usable when its block / out_proj
        // ops are exactly the plain plan's, so both kernels read ONE weight blob.
        bool loop = !sd.plan.pair && !getenv("QB_NO_DECODE_LOOP");
        std::vector<QbOp> lops;
        if (loop) {
            qb::PlanOptions lo = opt;
            lo.uop = 1; lo.pair = 1; lo.no_esplit = sd.plan.e_split ? 0 : 1; lo.blk32 = opt.blk32; lo.hc = sd.plan.hc; lo.slot_bytes = sd.plan.slot_bytes; lo.mcast = sd.plan.mcast ? 2 : 1;
            std::string lerr;
            loop = qb::make_step_plan(D, De, m->Dh, m->L, K, m->q1, lo, &sd.loop_plan, &lops, &lerr) == 0 &&
                   sd.loop_plan.n_ops_block == sd.plan.n_ops_block && sd.loop_plan.n_ops_out == sd.plan.n_ops_out &&
                   sd.loop_plan.block_w_bytes == sd.plan.block_w_bytes && lops.size() >= ops.size() &&
                   std::memcmp(lops.data(), ops.data(), ops.size() * sizeof(QbOp)) == 0;
        }
        m->loop_ok = (s == 1) ? loop : (m->loop_ok && loop);
        const QbStepPlan& pack_plan = loop ? sd.loop_plan : sd.plan;
        const std::vector<QbOp>& pack_ops = loop ? lops : ops;
        std::vector<uint16_t> blob((size_t)(pack_plan.w_blob_bytes + 1) / 2, 0);
        if (qb::pack_step_weights(pack_plan, pack_ops, d->up_w ? d->up_w + (size_t)s * m->L : nullptr,
                                  d->down_w ? d->down_w + (size_t)s * m->L : nullptr,
                                  De != D ? d->out_proj[s] : nullptr, blob.data(), &err))
            return bail(fail(QB_ERR_INVALID, err));
        std::vector<float> t_blk((size_t)De * K), cb_blk((size_t)D * K), wx_t((size_t)D * De);
        qb::build_tables(D, De, K, d->codebook[s], De != D ? d->in_proj[s] : nullptr, d->concat_w[s], d->concat_b[s],
                         t_blk.data(), cb_blk.data(), wx_t.data());
        sd.ops = ops;
        sd.loop_ops = lops;
        if (loop) max_smem = std::max(max_smem, sd.loop_plan.smem_total);
        if ((rc = dev_upload(m, t_blk.data(), t_blk.size(), &sd.t_blk))) return bail(rc);
        if ((rc = dev_upload(m, cb_blk.data(), cb_blk.size(), &sd.cb_blk))) return bail(rc);
        {   // Wx = Wcat[:, De:] as [De][D] rows (build_tables hands back its transpose)
            std::vector<float> wx((size_t)De * D);
            for (int e = 0; e < De; e++)
                for (int dd = 0; dd < D; dd++) wx[(size_t)e * D + dd] = wx_t[(size_t)dd * De + e];
            if ((rc = dev_upload(m, wx.data(), wx.size(), &sd.wx))) return bail(rc);
            if (loop && qb::pack_pre_weights(sd.loop_plan, lops, wx.data(), blob.data(), &err)) return bail(fail(QB_ERR_INVALID, err));
            std::vector<uint16_t> pk(qb::prep_pack_bytes(De, D) / 2);
            qb::prep_pack(wx.data(), De, D, pk.data());
            if ((rc = dev_upload(m, (const uint8_t*)pk.data(), pk.size() * 2, &sd.wx_pack))) return bail(rc);
        }
        if ((rc = dev_upload(m, (const uint8_t*)blob.data(), blob.size() * 2, &sd.w_blob))) return bail(rc);
        if (m->A > 0) {
            std::vector<float> nrm((size_t)K);
            for (int k = 0; k < K; k++) nrm[k] = row_norm2(d->substep_codebook[s] + (size_t)k * D, D);
            if ((rc = dev_upload(m, d->substep_codebook[s], (size_t)K * D, &sd.sub_cb))) return bail(rc);
            if ((rc = dev_upload(m, nrm.data(), nrm.size(), &sd.sub_norm))) return bail(rc);
            std::vector<uint16_t> pk(qb::prep_pack_bytes(K, D) / 2);
            qb::prep_pack(d->substep_codebook[s], K, D, pk.data());
            if ((rc = dev_upload(m, (const uint8_t*)pk.data(), pk.size() * 2, &sd.sub_pack))) return bail(rc);
        }
        max_smem = std::max(max_smem, sd.plan.smem_total);
    }
    if (max_smem > 0) {
        cudaError_t e = qb::mlp_set_smem_attr(max_smem);
        if (e != cudaSuccess) return bail(fail_cuda(e, "cudaFuncSetAttribute(max dynamic smem)"));
    }
    QB_CUDA(cudaDeviceSynchronize());
    *out = m;
    return QB_OK;
}

int qb_model_destroy(qb_model* m) {
    if (!m) return QB_OK;
    qb::DeviceGuard guard(m->device);
    for (void* p : m->dev_allocs) cudaFree(p);
    for (auto& s : m->slot) {
        if (s.x_pin) cudaFreeHost(s.x_pin);
        if (s.out_pin) cudaFreeHost(s.out_pin);
        if (s.codes_pin) cudaFreeHost(s.codes_pin);
        if (s.x_dev) cudaFree(s.x_dev);
        if (s.out_dev) cudaFree(s.out_dev);
        if (s.codes_dev) cudaFree(s.codes_dev);
        if (s.ivf_pin) cudaFreeHost(s.ivf_pin);
        if (s.ivf_dev) cudaFree(s.ivf_dev);
        if (s.ev_h2d) cudaEventDestroy(s.ev_h2d);
        if (s.ev_comp) cudaEventDestroy(s.ev_comp);
        if (s.ev_d2h) cudaEventDestroy(s.ev_d2h);
    }
    if (m->host_ws) cudaFree(m->host_ws);
    if (m->s_h2d) cudaStreamDestroy(m->s_h2d);
    if (m->s_comp) cudaStreamDestroy(m->s_comp);
    if (m->s_d2h) cudaStreamDestroy(m->s_d2h);
    if (m->err_host) cudaFreeHost(m->err_host);
    for (auto& t : m->timed) { cudaEventDestroy(t.a); cudaEventDestroy(t.b); }
    for (auto e : m->ev_pool) cudaEventDestroy(e);
    delete m;
    return QB_OK;
}

size_t qb_encode_workspace_bytes(const qb_model* m, int64_t n) {
    if (!m || n <= 0) return kWsSlack;
    const int64_t nc = std::min<int64_t>(n, default_chunk(m));
    return (size_t)nc * encode_bytes_per_vector(m) + kWsSlack;
}
size_t qb_decode_workspace_bytes(const qb_model* m, int64_t n) {
    if (!m || n <= 0) return kWsSlack;
    const int64_t nc = std::min<int64_t>(n, default_chunk(m) * 16);
    return (size_t)nc * decode_bytes_per_vector(m) + kWsSlack;
}

static int encode_impl(qb_model* m, const float* x_dev, int64_t n, int normalize, int32_t* ivf_codes_dev, uint8_t* codes_dev,
                       float* xhat_dev, void* workspace_dev, size_t workspace_bytes, void* stream) {
    if (!m) return fail(QB_ERR_INVALID, "model is NULL");
    if (n < 0) return fail(QB_ERR_INVALID, "n < 0");
    if ((m->ivf_K > 0) != (ivf_codes_dev != nullptr))
        return fail(QB_ERR_INVALID, m->ivf_K ? "IVF model: use qb_encode_ivf (needs a buffer for the IVF codes)"
                                             : "not an IVF model: use qb_encode");
    if (n == 0) return QB_OK;
    if (!x_dev || !codes_dev || !workspace_dev) return fail(QB_ERR_INVALID, "NULL buffer");
    if (workspace_bytes <= kWsSlack) return fail(QB_ERR_WORKSPACE, "workspace too small");
    int64_t nc = (int64_t)((workspace_bytes - kWsSlack) / encode_bytes_per_vector(m));
    nc = std::min<int64_t>(nc, default_chunk(m));
    if (nc < 1) return fail(QB_ERR_WORKSPACE, "workspace too small for one vector; see qb_encode_workspace_bytes");
    if (nc >= 128) nc = nc / 128 * 128;
    qb::DeviceGuard guard(m->device);
    QB_CUDA(guard.err);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    for (int64_t i0 = 0; i0 < n; i0 += nc) {
        const int64_t c = std::min(nc, n - i0);
        int rc = encode_chunk(m, x_dev + i0 * m->D, c, normalize, ivf_codes_dev ? ivf_codes_dev + i0 : nullptr,
                              codes_dev + i0 * m->M, xhat_dev ? xhat_dev + i0 * m->D : nullptr, workspace_dev, st);
        if (rc) return rc;
    }
    return QB_OK;
}

int qb_encode(qb_model* m, const float* x_dev, int64_t n, int normalize, uint8_t* codes_dev, float* xhat_dev,
              void* workspace_dev, size_t workspace_bytes, void* stream) {
    return encode_impl(m, x_dev, n, normalize, nullptr, codes_dev, xhat_dev, workspace_dev, workspace_bytes, stream);
}

int qb_encode_ivf(qb_model* m, const float* x_dev, int64_t n, int normalize, int32_t* ivf_codes_dev, uint8_t* codes_dev,
                  float* xhat_dev, void* workspace_dev, size_t workspace_bytes, void* stream) {
    if (!ivf_codes_dev) return fail(QB_ERR_INVALID, "ivf_codes_dev is NULL");
    return encode_impl(m, x_dev, n, normalize, ivf_codes_dev, codes_dev, xhat_dev, workspace_dev, workspace_bytes, stream);
}

static int decode_impl(qb_model* m, const int32_t* ivf_codes_dev, const uint8_t* codes_dev, int64_t n, int denormalize,
                       float* out_dev, void* workspace_dev, size_t workspace_bytes, void* stream) {
    if (!m) return fail(QB_ERR_INVALID, "model is NULL");
    if (n < 0) return fail(QB_ERR_INVALID, "n < 0");
    if ((m->ivf_K > 0) != (ivf_codes_dev != nullptr))
        return fail(QB_ERR_INVALID, m->ivf_K ? "IVF model: use qb_decode_ivf (needs the IVF codes)" : "not an IVF model: use qb_decode");
    if (n == 0) return QB_OK;
    if (!codes_dev || !out_dev || !workspace_dev) return fail(QB_ERR_INVALID, "NULL buffer");
    if (workspace_bytes <= kWsSlack) return fail(QB_ERR_WORKSPACE, "workspace too small");
    int64_t nc = (int64_t)((workspace_bytes - kWsSlack) / decode_bytes_per_vector(m));
    nc = std::min<int64_t>(nc, default_chunk(m) * 16);
    if (nc < 1) return fail(QB_ERR_WORKSPACE, "workspace too small for one vector; see qb_decode_workspace_bytes");
    if (nc >= 128) nc = nc / 128 * 128;
    if (m->loop_ok && m->S > 1) {
        // One-launch decode: the workspace is not used, so a launch may cover any number of vectors.  Models whose step
        // weights together do not fit half the L2 (QINCo2-L: 7 x 9.4 MB) go in waves of ONE tile set per CTA: all CTAs then
        // walk the steps in lockstep and stream the same ~10 MB at any time; with several sets per CTA they drift apart,
        // the working set becomes every step's weights and the stream falls back to HBM (measured: 8.6 -> 6.9 ms per 100 k
        // vectors of BASELINE config 3, and no more 2-5x outliers).
        const QbStepPlan& lp = m->steps[1].loop_plan;
        const int64_t wave = (int64_t)(lp.mcast ? (m->n_sm & ~1) : m->n_sm) * lp.n_tiles * QB_TILE_M;
        const bool big = (int64_t)(m->S - 1) * lp.w_blob_bytes > (int64_t)24 << 20;
        nc = big ? wave : std::max<int64_t>(nc, (int64_t)1 << 22);
    }
    qb::DeviceGuard guard(m->device);
    QB_CUDA(guard.err);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    for (int64_t i0 = 0; i0 < n; i0 += nc) {
        const int64_t c = std::min(nc, n - i0);
        int rc = decode_chunk(m, ivf_codes_dev ? ivf_codes_dev + i0 : nullptr, codes_dev + i0 * m->M, c, denormalize,
                              out_dev + i0 * m->D, workspace_dev, st);
        if (rc) return rc;
    }
    return QB_OK;
}

int qb_decode(qb_model* m, const uint8_t* codes_dev, int64_t n, int denormalize, float* out_dev, void* workspace_dev,
              size_t workspace_bytes, void* stream) {
    return decode_impl(m, nullptr, codes_dev, n, denormalize, out_dev, workspace_dev, workspace_bytes, stream);
}

int qb_decode_ivf(qb_model* m, const int32_t* ivf_codes_dev, const uint8_t* codes_dev, int64_t n, int denormalize,
                  float* out_dev, void* workspace_dev, size_t workspace_bytes, void* stream) {
    if (!ivf_codes_dev) return fail(QB_ERR_INVALID, "ivf_codes_dev is NULL");
    return decode_impl(m, ivf_codes_dev, codes_dev, n, denormalize, out_dev, workspace_dev, workspace_bytes, stream);
}

int qb_check(qb_model* m) {
    if (!m) return fail(QB_ERR_INVALID, "model is NULL");
    return check_device_flag(m);
}

// Pipelined host loops: chunk i is copied in on one stream while chunk i-1 computes and chunk i-2 copies out.
static int host_loop_body(qb_model* m, bool enc, const float* x_host, const int32_t* ivf_in, const uint8_t* codes_in, int64_t n,
                          int flag, int32_t* ivf_out, uint8_t* codes_out, float* out_host);

// Error exits leave copies in flight and staging slots marked pending: drain the three streams and forget the slots, so
// the next call on this model cannot copy a stale chunk into ITS caller's buffers.
static int host_loop(qb_model* m, bool enc, const float* x_host, const int32_t* ivf_in, const uint8_t* codes_in, int64_t n,
                     int flag, int32_t* ivf_out, uint8_t* codes_out, float* out_host) {
    qb::DeviceGuard guard(m->device);
    QB_CUDA(guard.err);
    const int rc = host_loop_body(m, enc, x_host, ivf_in, codes_in, n, flag, ivf_out, codes_out, out_host);
    if (rc != QB_OK) {
        const std::string keep = g_err;
        if (m->s_h2d) cudaStreamSynchronize(m->s_h2d);
        if (m->s_comp) cudaStreamSynchronize(m->s_comp);
        if (m->s_d2h) cudaStreamSynchronize(m->s_d2h);
        cudaGetLastError();
        for (auto& s : m->slot) s.pending_i0 = -1;
        g_err = keep;
    }
    return rc;
}

static int host_loop_body(qb_model* m, bool enc, const float* x_host, const int32_t* ivf_in, const uint8_t* codes_in, int64_t n,
                          int flag, int32_t* ivf_out, uint8_t* codes_out, float* out_host) {
    int rc = ensure_host_staging(m, !enc);
    if (rc) return rc;
    const int64_t nc = m->host_chunk;
    const int D = m->D, M = m->M;
    auto finalize = [&](HostSlot& s) -> int {
        if (s.pending_i0 < 0) return QB_OK;
        QB_CUDA(cudaEventSynchronize(s.ev_d2h));
        if (enc) {
            std::memcpy(codes_out + s.pending_i0 * M, s.codes_pin, (size_t)s.pending_n * M);
            if (ivf_out) std::memcpy(ivf_out + s.pending_i0, s.ivf_pin, (size_t)s.pending_n * 4);
            if (out_host) std::memcpy(out_host + s.pending_i0 * D, s.out_pin, (size_t)s.pending_n * D * 4);
        } else {
            std::memcpy(out_host + s.pending_i0 * D, s.out_pin, (size_t)s.pending_n * D * 4);
        }
        s.pending_i0 = -1;
        return QB_OK;
    };
    int64_t ci = 0;
    for (int64_t i0 = 0; i0 < n; i0 += nc, ci++) {
        HostSlot& s = m->slot[ci & 1];
        const int64_t c = std::min(nc, n - i0);
        if ((rc = finalize(s))) return rc;
        if (enc) {
            std::memcpy(s.x_pin, x_host + i0 * D, (size_t)c * D * 4);
            QB_CUDA(cudaMemcpyAsync(s.x_dev, s.x_pin, (size_t)c * D * 4, cudaMemcpyHostToDevice, m->s_h2d));
        } else {
            std::memcpy(s.codes_pin, codes_in + i0 * M, (size_t)c * M);
            QB_CUDA(cudaMemcpyAsync(s.codes_dev, s.codes_pin, (size_t)c * M, cudaMemcpyHostToDevice, m->s_h2d));
            if (ivf_in) {
                std::memcpy(s.ivf_pin, ivf_in + i0, (size_t)c * 4);
                QB_CUDA(cudaMemcpyAsync(s.ivf_dev, s.ivf_pin, (size_t)c * 4, cudaMemcpyHostToDevice, m->s_h2d));
            }
        }
        QB_CUDA(cudaEventRecord(s.ev_h2d, m->s_h2d));
        QB_CUDA(cudaStreamWaitEvent(m->s_comp, s.ev_h2d, 0));
        if (enc)
            rc = encode_impl(m, s.x_dev, c, flag, m->ivf_K ? s.ivf_dev : nullptr, s.codes_dev, out_host ? s.out_dev : nullptr,
                             m->host_ws, m->host_ws_bytes, m->s_comp);
        else
            rc = decode_impl(m, m->ivf_K ? s.ivf_dev : nullptr, s.codes_dev, c, flag, s.out_dev, m->host_ws, m->host_ws_bytes,
                             m->s_comp);
        if (rc) return rc;
        QB_CUDA(cudaEventRecord(s.ev_comp, m->s_comp));
        QB_CUDA(cudaStreamWaitEvent(m->s_d2h, s.ev_comp, 0));
        if (enc) {
            QB_CUDA(cudaMemcpyAsync(s.codes_pin, s.codes_dev, (size_t)c * M, cudaMemcpyDeviceToHost, m->s_d2h));
            if (ivf_out) QB_CUDA(cudaMemcpyAsync(s.ivf_pin, s.ivf_dev, (size_t)c * 4, cudaMemcpyDeviceToHost, m->s_d2h));
            if (out_host)
                QB_CUDA(cudaMemcpyAsync(s.out_pin, s.out_dev, (size_t)c * D * 4, cudaMemcpyDeviceToHost, m->s_d2h));
        } else {
            QB_CUDA(cudaMemcpyAsync(s.out_pin, s.out_dev, (size_t)c * D * 4, cudaMemcpyDeviceToHost, m->s_d2h));
        }
        QB_CUDA(cudaEventRecord(s.ev_d2h, m->s_d2h));
        s.pending_i0 = i0;
        s.pending_n = c;
    }
    for (int k = 0; k < 2; k++)
        if ((rc = finalize(m->slot[(ci + k) & 1]))) return rc;
    return check_device_flag(m);
}

int qb_encode_host(qb_model* m, const float* x_host, int64_t n, int normalize, uint8_t* codes_host, float* xhat_host) {
    if (!m) return fail(QB_ERR_INVALID, "model is NULL");
    if (n < 0) return fail(QB_ERR_INVALID, "n < 0");
    if (n == 0) return QB_OK;
    if (!x_host || !codes_host) return fail(QB_ERR_INVALID, "NULL buffer");
    if (m->ivf_K) return fail(QB_ERR_INVALID, "IVF model: use qb_encode_ivf_host");
    return host_loop(m, true, x_host, nullptr, nullptr, n, normalize, nullptr, codes_host, xhat_host);
}

int qb_encode_ivf_host(qb_model* m, const float* x_host, int64_t n, int normalize, int32_t* ivf_codes_host, uint8_t* codes_host,
                       float* xhat_host) {
    if (!m) return fail(QB_ERR_INVALID, "model is NULL");
    if (n < 0) return fail(QB_ERR_INVALID, "n < 0");
    if (!m->ivf_K) return fail(QB_ERR_INVALID, "not an IVF model: use qb_encode_host");
    if (n == 0) return QB_OK;
    if (!x_host || !codes_host || !ivf_codes_host) return fail(QB_ERR_INVALID, "NULL buffer");
    return host_loop(m, true, x_host, nullptr, nullptr, n, normalize, ivf_codes_host, codes_host, xhat_host);
}

int qb_decode_host(qb_model* m, const uint8_t* codes_host, int64_t n, int denormalize, float* out_host) {
    if (!m) return fail(QB_ERR_INVALID, "model is NULL");
    if (n < 0) return fail(QB_ERR_INVALID, "n < 0");
    if (n == 0) return QB_OK;
    if (!codes_host || !out_host) return fail(QB_ERR_INVALID, "NULL buffer");
    for (int64_t i = 0; i < n * m->M; i++)
        if (codes_host[i] >= m->K) return fail(QB_ERR_INVALID, "code out of range [0,K)");
    if (m->ivf_K) return fail(QB_ERR_INVALID, "IVF model: use qb_decode_ivf_host");
    return host_loop(m, false, nullptr, nullptr, codes_host, n, denormalize, nullptr, nullptr, out_host);
}

int qb_decode_ivf_host(qb_model* m, const int32_t* ivf_codes_host, const uint8_t* codes_host, int64_t n, int denormalize,
                       float* out_host) {
    if (!m) return fail(QB_ERR_INVALID, "model is NULL");
    if (n < 0) return fail(QB_ERR_INVALID, "n < 0");
    if (!m->ivf_K) return fail(QB_ERR_INVALID, "not an IVF model: use qb_decode_host");
    if (n == 0) return QB_OK;
    if (!ivf_codes_host || !codes_host || !out_host) return fail(QB_ERR_INVALID, "NULL buffer");
    for (int64_t i = 0; i < n * m->M; i++)
        if (codes_host[i] >= m->K) return fail(QB_ERR_INVALID, "code out of range [0,K)");
    for (int64_t i = 0; i < n; i++)
        if (ivf_codes_host[i] < 0 || ivf_codes_host[i] >= m->ivf_K) return fail(QB_ERR_INVALID, "IVF code out of range [0,ivf_K)");
    return host_loop(m, false, nullptr, ivf_codes_host, codes_host, n, denormalize, nullptr, nullptr, out_host);
}

int qb_codes_pack(qb_model* m, const void* codes_MB_dev, int elem_bytes, int64_t stride_row, int64_t stride_col, int64_t n,
                  uint8_t* codes_dev, int32_t* ivf_codes_dev, void* stream) {
    if (!m) return fail(QB_ERR_INVALID, "model is NULL");
    if (n < 0) return fail(QB_ERR_INVALID, "n < 0");
    if (n == 0) return QB_OK;
    if (!codes_MB_dev || !codes_dev) return fail(QB_ERR_INVALID, "NULL buffer");
    if ((m->ivf_K > 0) != (ivf_codes_dev != nullptr)) return fail(QB_ERR_INVALID, "ivf_codes_dev must be given exactly for IVF models");
    if (elem_bytes != 1 && elem_bytes != 4 && elem_bytes != 8) return fail(QB_ERR_INVALID, "codes must be uint8, int32 or int64");
    qb::DeviceGuard guard(m->device);
    QB_CUDA(guard.err);
    m->launches++;
    QB_CUDA(qb::launch_codes_pack(codes_MB_dev, elem_bytes, stride_row, stride_col, n, m->M, m->K, m->ivf_K, codes_dev,
                                  ivf_codes_dev, m->err_dev, reinterpret_cast<cudaStream_t>(stream)));
    return QB_OK;
}

int qb_codes_unpack(qb_model* m, const uint8_t* codes_dev, const int32_t* ivf_codes_dev, int64_t n, int64_t* codes_MB_dev,
                    void* stream) {
    if (!m) return fail(QB_ERR_INVALID, "model is NULL");
    if (n < 0) return fail(QB_ERR_INVALID, "n < 0");
    if (n == 0) return QB_OK;
    if (!codes_MB_dev || !codes_dev) return fail(QB_ERR_INVALID, "NULL buffer");
    if ((m->ivf_K > 0) != (ivf_codes_dev != nullptr)) return fail(QB_ERR_INVALID, "ivf_codes_dev must be given exactly for IVF models");
    qb::DeviceGuard guard(m->device);
    QB_CUDA(guard.err);
    m->launches++;
    QB_CUDA(qb::launch_codes_unpack(codes_dev, ivf_codes_dev, n, m->M, m->ivf_K ? 1 : 0, codes_MB_dev,
                                    reinterpret_cast<cudaStream_t>(stream)));
    return QB_OK;
}

int64_t qb_launch_count(const qb_model* m) { return m ? m->launches : 0; }

int qb_timing_enable(qb_model* m, int on) {
    if (!m) return fail(QB_ERR_INVALID, "model is NULL");
    m->timing = on != 0;
    return QB_OK;
}

int qb_timing_read(qb_model* m, double* ms_out, int64_t* launches_out, int64_t* rows_out, int n_kinds) {
    if (!m || !ms_out || !launches_out || !rows_out) return fail(QB_ERR_INVALID, "NULL argument");
    qb::DeviceGuard guard(m->device);
    QB_CUDA(guard.err);
    for (int k = 0; k < n_kinds; k++) { ms_out[k] = 0; launches_out[k] = 0; rows_out[k] = 0; }
    for (auto& t : m->timed) {
        QB_CUDA(cudaEventSynchronize(t.b));
        float ms = 0.f;
        QB_CUDA(cudaEventElapsedTime(&ms, t.a, t.b));
        if (t.kind < n_kinds) { ms_out[t.kind] += ms; launches_out[t.kind]++; rows_out[t.kind] += t.rows; }
        m->ev_pool.push_back(t.a);
        m->ev_pool.push_back(t.b);
    }
    m->timed.clear();
    return KIND_COUNT;
}

int qb_model_info(const qb_model* m, int step, int32_t* out, int n_out) {
    if (!m || !out) return fail(QB_ERR_INVALID, "NULL argument");
    if (step < 1 || step >= m->S) return fail(QB_ERR_INVALID, "step out of range (MLP steps are 1..S-1)");
    const QbStepPlan& p = m->steps[step].plan;
    const int32_t v[] = {p.D, p.De, p.Dh, p.L, p.K, p.has_proj, p.skip, p.n_ops_block, p.n_ops_out, p.hc, p.n_hchunk,
                         p.n_tiles, p.oc, p.n_ochunk, p.slot_bytes, p.n_stage, p.smem_total, (int32_t)p.block_w_bytes,
                         (int32_t)p.w_blob_bytes, m->n_sm, (int32_t)default_chunk(m), p.pair, m->loop_ok ? 1 : 0,
                         m->loop_ok ? m->steps[step].loop_plan.n_stage : 0, m->loop_ok ? m->steps[step].loop_plan.smem_total : 0};
    const int nv = (int)(sizeof(v) / sizeof(v[0]));
    for (int i = 0; i < n_out && i < nv; i++) out[i] = v[i];
    return nv;
}

int qb_debug_step(qb_model* m, int step, const float* xhat_dev, const uint8_t* codes_dev, int64_t n, float* out_dev,
                  void* workspace_dev, size_t workspace_bytes, void* stream) {
    if (!m) return fail(QB_ERR_INVALID, "model is NULL");
    if (step < 1 || step >= m->S) return fail(QB_ERR_INVALID, "step out of range (MLP steps are 1..S-1)");
    if (n <= 0) return QB_OK;
    if (workspace_bytes < (size_t)n * m->De * 4 + 256) return fail(QB_ERR_WORKSPACE, "workspace too small (n*de*4+256)");
    qb::DeviceGuard guard(m->device);
    QB_CUDA(guard.err);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    float* u = reinterpret_cast<float*>(align_up(reinterpret_cast<size_t>(workspace_dev), 256));
    const StepDev& s = m->steps[step];
    qb::PrepParams pp;
    std::memset(&pp, 0, sizeof(pp));
    pp.D = m->D; pp.De = m->De; pp.K = m->K; pp.F = 1; pp.M = m->M; pp.n_beams = n;
    pp.x = xhat_dev; pp.inv_std = 1.f; pp.xhat = xhat_dev; pp.wx = s.wx; pp.u = u;
    QB_CUDA(qb::launch_prep(pp, st));
    qb::MlpParams p = base_mlp(m, step);
    p.mode = qb::QB_MODE_APPLY;
    p.F_in = 1; p.F_out = 1; p.n_rows = n;
    p.sel_code = codes_dev; p.code_stride = 1; p.code_off = 0;
    p.u = u; p.xhat_in = xhat_dev; p.xhat_out = out_dev;
    QB_CUDA(qb::launch_mlp(p, m->n_sm, st));
    m->launches += 2;
    return QB_OK;
}

}  // extern "C"
