/* qinco_b200 — C ABI of the B200-native QINCo / QINCo2 encode-decode path (libqinco_b200.so).
 *
 * The reference (facebookresearch/Qinco @ 5a324954) has no FFI: its boundary is a duck-typed PyTorch module.
 * Each entry point below names the reference call it stands in for; the Python shim in qinco_b200/model.py and
 * qinco_b200/codec.py rebuilds the reference's signatures on top of these (INTEGRATION.md shows the binding).
 *
 * Conventions
 *   - plain C types only; device buffers are raw CUDA device pointers, the stream is a cudaStream_t passed as void*
 *   - every call returns 0 (QB_OK) or a negative qb_status; qb_last_error() returns a thread-local message
 *   - the caller owns every buffer including the workspace; qb_encode / qb_decode are asynchronous on `stream`,
 *     allocate nothing and never synchronise; the *_host variants stage through pinned memory and return when done
 *   - a model handle is bound to one CUDA device and is not thread-safe
 *   - codes are uint8 [n, M] row-major (vector-major).  The reference's layouts ([M, n] int64 for qinco.model,
 *     [n, M] int64 for qinco_v1/codec_qinco.py) are produced at the Python edge.
 *   - there is no CPU fallback: without a CUDA device every compute call fails with QB_ERR_CUDA
 */
#ifndef QINCO_B200_H_
#define QINCO_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct qb_model qb_model;

typedef enum {
    QB_OK = 0,
    QB_ERR_INVALID = -1,     /* bad argument / unsupported shape */
    QB_ERR_CUDA = -2,        /* CUDA runtime error (message has the cudaError string) */
    QB_ERR_WORKSPACE = -3,   /* workspace too small */
    QB_ERR_KERNEL = -4,      /* device-side protocol time-out or out-of-range code (err word in message) */
    QB_ERR_NOMEM = -5
} qb_status;

/* Model weights, HOST pointers, fp32, row-major, PyTorch [out, in] convention — exactly the tensors of the
 * reference state dict (qinco/model/qinco_base.py:229-260, 432-433):
 *   steps.{m}.codebook.weight [K,D]; steps.{m}.substep.codebook.weight [K,D] (m>=1, A>0);
 *   steps.{m}.concat.mlp.{weight [De,De+D], bias [De]}; steps.{m}.residual_blocks.{l}.{up_proj [Dh,De], down_proj [De,Dh]};
 *   steps.{m}.in_proj [De,D] / out_proj [D,De] (only when De != D); data_mean [D]; data_std [].
 * Arrays are indexed by step m (entry 0 of the MLP arrays is ignored; step 0 is a plain codebook,
 * qinco_base.py:218,263) and up_w / down_w by m*L + l. */
typedef struct {
    int32_t D, De, Dh, L, M, K;
    int32_t A;               /* pre-selected candidates per beam, 0 = all K (cfg.A) */
    int32_t B;               /* beam width (cfg.B) */
    int32_t qinco1_mode;     /* 1: no outer skip connection (cfg.qinco1_mode, qinco_base.py:276-278) */
    int32_t device;          /* CUDA device ordinal */
    const float* const* codebook;          /* [M] */
    const float* const* substep_codebook;  /* [M] or NULL when A == 0 */
    const float* const* concat_w;          /* [M] */
    const float* const* concat_b;          /* [M] */
    const float* const* up_w;              /* [M*L] */
    const float* const* down_w;            /* [M*L] */
    const float* const* in_proj;           /* [M] or NULL when De == D */
    const float* const* out_proj;          /* [M] or NULL when De == D */
    const float* data_mean;                /* [D] or NULL (zeros) */
    float data_std;                        /* > 0 */
    /* kernel planner overrides, 0 = automatic; opt_n_tiles = n_tiles | pair << 8 (pair: 0 auto, 1 off, 2 on) */
    int32_t opt_hc, opt_n_tiles, opt_slot_bytes, opt_max_stage, opt_max_slab_k;
    int32_t opt_stagger;     /* start delay (cycles) of the second CTA per SM; 0 = automatic, < 0 = none */
    /* IVF-QINCo (cfg.ivf_in_use; IVFBook, qinco_base.py:128-196): ivf_K > 0 makes step 0 an arg-min over ivf_K centroids
     * and steps 1..M implicit-codebook steps (cfg._M_ivf = M + 1).  The per-step arrays then have M + 1 entries indexed
     * by step (entry 0 unused, codebook[0] may be NULL); M stays the number of uint8 codes per vector. */
    int32_t ivf_K;
    const float* ivf_centroids;            /* [ivf_K][D], normalised space (steps.0.ivf_centroids.weight) */
} qb_model_desc;

int qb_version(void);
const char* qb_last_error(void);

/* Replaces constructing QINCo(cfg) + load_state_dict (+ QINCoInferenceWrapper.build, qinco_inference.py:290-330):
 * packs the weights for the kernels (fp16 MMA operand slabs, fp32 tables) and uploads them. */
int qb_model_create(const qb_model_desc* desc, qb_model** out);
int qb_model_destroy(qb_model* m);

/* Bytes of device workspace qb_encode / qb_decode want for n vectors (they process in internal chunks, so the value
 * saturates; a smaller workspace still works as long as one 128-vector chunk fits). */
size_t qb_encode_workspace_bytes(const qb_model* m, int64_t n);
size_t qb_decode_workspace_bytes(const qb_model* m, int64_t n);

/* Replaces QINCo.encode / QINCoInferenceWrapper.encode (qinco_base.py:454-485, qinco_inference.py:340-350) and, with
 * normalize != 0, forward(x, step="encode") (qinco_base.py:532-534): x_dev [n, D] fp32 row-major ->
 * codes_dev [n, M] uint8 and, if xhat_dev != NULL, the reconstruction [n, D] fp32 in NORMALISED space. */
int qb_encode(qb_model* m, const float* x_dev, int64_t n, int normalize, uint8_t* codes_dev, float* xhat_dev,
              void* workspace_dev, size_t workspace_bytes, void* stream);

/* Replaces QINCo.decode / QINCoInferenceWrapper.decode (qinco_base.py:447-452, qinco_inference.py:332-338) and, with
 * denormalize != 0, forward(codes, step="decode") (qinco_base.py:536-537): codes_dev [n, M] uint8 -> out_dev [n, D]. */
int qb_decode(qb_model* m, const uint8_t* codes_dev, int64_t n, int denormalize, float* out_dev, void* workspace_dev,
              size_t workspace_bytes, void* stream);

/* IVF-QINCo models (desc.ivf_K > 0): the code matrix of the reference has M + 1 rows, row 0 the IVF code.  Here the IVF
 * codes travel in their own int32 array; codes_dev stays [n, M] uint8.  qb_encode / qb_decode refuse IVF models and these
 * refuse plain ones. */
int qb_encode_ivf(qb_model* m, const float* x_dev, int64_t n, int normalize, int32_t* ivf_codes_dev, uint8_t* codes_dev,
                  float* xhat_dev, void* workspace_dev, size_t workspace_bytes, void* stream);
int qb_decode_ivf(qb_model* m, const int32_t* ivf_codes_dev, const uint8_t* codes_dev, int64_t n, int denormalize,
                  float* out_dev, void* workspace_dev, size_t workspace_bytes, void* stream);
/* ... and their host-buffer variants (the loop of encode_database, qinco/search/search_tasks.py:107-116, for IVF models) */
int qb_encode_ivf_host(qb_model* m, const float* x_host, int64_t n, int normalize, int32_t* ivf_codes_host, uint8_t* codes_host,
                       float* xhat_host);
int qb_decode_ivf_host(qb_model* m, const int32_t* ivf_codes_host, const uint8_t* codes_host, int64_t n, int denormalize,
                       float* out_host);

/* Host-buffer variants = the batch loops of qinco_v1/codec_qinco.py:25-46 and :54-72 (H2D, encode/decode, D2H per
 * chunk, through pinned staging owned by the model).  xhat_host may be NULL. */
int qb_encode_host(qb_model* m, const float* x_host, int64_t n, int normalize, uint8_t* codes_host, float* xhat_host);
int qb_decode_host(qb_model* m, const uint8_t* codes_host, int64_t n, int denormalize, float* out_host);

/* Code-matrix conversions at the reference's tensor surface, on the device and asynchronous on `stream`.
 * qb_codes_pack: the [S, n] integer matrix the reference passes to decode (qinco_base.py:447-449; S = M, or M + 1 with the
 * IVF code in row 0; uint8 / int32 / int64 elements -- the IVF search passes a transposed int32 view, search_tasks.py:428-445,
 * hence the two element strides) -> codes_dev uint8 [n, M] (+ ivf_codes_dev int32 [n] for IVF models, else NULL).  Codes
 * outside [0, K) / [0, ivf_K) set the device error word (0x10 / 0x20, reported once by qb_check) and are clamped: the
 * counterpart of the device-side index assert the reference's codebook lookup would raise.
 * qb_codes_unpack: uint8 [n, M] (+ int32 [n]) -> contiguous int64 [S, n], what encode returns (qinco_base.py:480-485). */
int qb_codes_pack(qb_model* m, const void* codes_MB_dev, int elem_bytes, int64_t stride_row, int64_t stride_col, int64_t n,
                  uint8_t* codes_dev, int32_t* ivf_codes_dev, void* stream);
int qb_codes_unpack(qb_model* m, const uint8_t* codes_dev, const int32_t* ivf_codes_dev, int64_t n, int64_t* codes_MB_dev,
                    void* stream);

/* Reads the device-side error word (set by a kernel before it traps, or by an out-of-range code); 0 when clean.
 * An out-of-range-code report is cleared once returned (the kernels clamp and finish); a time-out is fatal and stays.
 * Cheap (one host read of mapped memory); call it after synchronising the stream. */
int qb_check(qb_model* m);

/* Introspection */
int64_t qb_launch_count(const qb_model* m);      /* kernels launched by this model so far */
/* Per-kernel device timing for bench.py's roofline block: with timing on, every launch is bracketed by CUDA events on
 * its own stream; qb_timing_read waits for them and returns, per kernel kind {0 prep, 1 mlp-score, 2 select,
 * 3 mlp-apply (update), 4 other, 5 IVF arg-min}, the summed milliseconds, launch count and rows processed since the last read. */
int qb_timing_enable(qb_model* m, int on);
int qb_timing_read(qb_model* m, double* ms_out, int64_t* launches_out, int64_t* rows_out, int n_kinds);
int qb_model_info(const qb_model* m, int step, int32_t* out, int n_out); /* plan of step>=1: see qb_api.cu */

/* Test hook: out[i] = xhat[i] + f_m(C_m[codes[i]], xhat[i]) for n independent rows of step `step` (>= 1), i.e. one
 * QINCoStep.decode (qinco_base.py:282-290) through the production kernels.  Device pointers. */
int qb_debug_step(qb_model* m, int step, const float* xhat_dev, const uint8_t* codes_dev, int64_t n, float* out_dev,
                  void* workspace_dev, size_t workspace_bytes, void* stream);

/* Host-only test hooks (no CUDA call): export the tcgen05 op list of one step (32-byte QbOp records, csrc/qb_plan.h),
 * pack weights into the slab blob and build the hoisted tables, so the CPU test-suite can replay the kernel's dataflow
 * in numpy.  opts5 = {hc, n_tiles, slot_bytes, max_stage, max_slab_k} or NULL.  plan_out[29..31] = e_split, mcast and the
 * compile-time plan view the kernels use for this plan (0 generic, 1 S128, 2 L384). */
int qb_plan_export(int D, int De, int Dh, int L, int K, int qinco1_mode, const int32_t* opts5, int32_t* plan_out,
                   int n_plan_out, void* ops_out, int max_ops);
int qb_plan_pack(int D, int De, int Dh, int L, int K, int qinco1_mode, const int32_t* opts5, const float* const* up,
                 const float* const* down, const float* out_proj, uint16_t* blob, int64_t blob_halfs);
/* ... and the slabs of the decode loop's pre-ops (u = Wx . xhat as fp16 hi/lo MMAs; opts5[3] bit 10 selects that plan);
 * wx = Wcat[:, De:] as [De][D] rows. */
int qb_plan_pack_pre(int D, int De, int Dh, int L, int K, int qinco1_mode, const int32_t* opts5, const float* wx, uint16_t* blob,
                     int64_t blob_halfs);
int qb_plan_tables(int D, int De, int K, const float* codebook, const float* in_proj, const float* concat_w,
                   const float* concat_b, float* t_blk, float* cb_blk, float* wx_t);

/* ---- Pairwise additive decoder, forward only ------------------------------------------
 * Replaces PairwiseDecoderIVF.forward + map_codes (qinco/search/pairwise_decoder.py:88-93, :126-130), called by the IVF
 * search's re-ranking stage (qinco/search/search_tasks.py:448-471).  Tables are the module's own tensors:
 * codebook_MKD [Mt, K*K, D] fp32, combine_mvals_m [2, Mt] int64, ivf_code_map [ivf_K, 5] int64 (host pointers, copied). */
typedef struct qb_pairwise qb_pairwise;
typedef struct qb_pairwise_desc {
    int32_t D, M, K;            /* vector dim, codes per vector (without the IVF code), base codebook size (<= 256) */
    int32_t Mt;                 /* number of pairwise codebooks = round(n_pairwise_codebooks * M) */
    int32_t ivf_K;              /* number of IVF centroids */
    int32_t device;
    const float* codebook;      /* [Mt][K*K][D] */
    const int64_t* combine;     /* [2][Mt], entries in [0, M + 5) */
    const int64_t* ivf_code_map;/* [ivf_K][5], entries in [0, K) */
} qb_pairwise_desc;
int qb_pairwise_create(const qb_pairwise_desc* desc, qb_pairwise** out);
int qb_pairwise_destroy(qb_pairwise* h);
/* forward(codes_MB, ivf_codes): codes_dev [n, M] uint8 row-major, ivf_codes_dev [n] int32 -> out_dev [n, D] fp32.
 * Asynchronous on `stream`; bit-identical to the reference's fp32 sum (same order of additions). */
int qb_pairwise_decode(qb_pairwise* h, const uint8_t* codes_dev, const int32_t* ivf_codes_dev, int64_t n, float* out_dev,
                       void* stream);
int qb_pairwise_check(qb_pairwise* h);                 /* device-side error word (out-of-range codes) */
int64_t qb_pairwise_launch_count(const qb_pairwise* h);
const char* qb_pairwise_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* QINCO_B200_H_ */
