// Step plan shared by the host planner/packer (qb_plan.cpp) and the tcgen05 MLP kernel (qb_mlp.cu).
//
// One quantisation step m >= 1 evaluates, for every candidate row, the implicit-codebook MLP
// (reference qinco/model/qinco_base.py:262-280: in_proj -> QConcat -> L x QBlockFFN -> out_proj [+ skip]).
// After hoisting the candidate-invariant terms:
//     e0  = T_m[code] + u_b                      T_m: [K,De] per-step table, u_b = Wcat[:,De:] . xhat_b
//     e  += Wdn_l . relu(Wup_l . e)   l < L      fp16 operands, fp32 accumulate, fp32 residual stream in TMEM
//     o   = Pout . e (+ C_m[code])               Pout absent when De == D; skip absent in qinco1_mode
//
// One CTA per SM works on n_tiles (2 when TMEM allows, else 1) 128-row tiles at a time, in lockstep: every GEMM of
// the op list is issued for tile X and then for tile Y on the SAME weight slabs (fetched once per two tiles), while the
// single group of epilogue warps alternates X, Y, X, Y ... — so the TMEM->relu->fp16 epilogue of one tile always
// overlaps the MMAs of the other.  Per tile, TMEM holds the fp32 residual accumulator Eacc (De columns) and one
// hidden-chunk accumulator Hacc (hc columns); relu(h) is written back IN PLACE as packed fp16 and read by the
// down-projection MMA as a TMEM A operand, so only e (fp16, [De/8][128][16 B] K-major core-matrix layout, one tile per
// slot) lives in shared memory next to the weight ring.
//
// The MMA warp walks a static list of "ops" (kept in the kernel-parameter constant bank).  An op is one GEMM
//     D[128, n] (+)= A[128, k_total] . W[n, k_total]^T
// whose weight matrix W is cut along K into n_slab slabs (n rows x ks cols, fp16, K-major, no-swizzle UMMA core-matrix
// layout [k/8][n][8]); the producer warp streams the slabs in op order through an mbarrier ring with cp.async.bulk.
// All L residual blocks share one op list (only the weight offset moves by block_w_bytes per block), followed by
// the out_proj ops.
#pragma once
#include <stdint.h>

#define QB_TILE_M 128
#define QB_MAX_OPS 64
#define QB_MAX_LOOP_STEPS 64
#define QB_MAX_STAGE 12

enum QbASrc : uint8_t { QB_A_E = 0, QB_A_H = 1 };   // A operand: e (shared memory) or relu(h) (TMEM, in place)
// barrier ids
enum QbBar : uint8_t {
    QB_BAR_NONE = 0,
    QB_BAR_AE_READY = 1,    // epilogue wrote A_E (fp16 e) [+ initialised Eacc]            count 128
    QB_BAR_AH_READY = 2,    // epilogue wrote relu(h) fp16 into TMEM                        count 128
    QB_BAR_HACC_FREE = 3,   // epilogue finished reading an out_proj chunk from Hacc        count 128
    QB_BAR_HACC_FULL = 4,   // MMA -> epilogue (tcgen05.commit)                             count 1
    QB_BAR_EACC_FULL = 5,
    QB_BAR_AH2_READY = 6,   // second K half of a split H chunk is in TMEM (its own barrier: with one barrier the epilogue
                            // could complete two phases before the issuer looks, and a parity wait cannot see that)
    QB_BAR_EACC_HALF = 7,   // first column part of Eacc is final (down-projection in two N parts, de > 256): the E epilogue converts it
                            // while the second part's MMAs run
    QB_BAR_COUNT = 8
};

struct QbOp {
    uint32_t w_off;      // byte offset of the first slab (multiple of 16): block ops relative to the block's weights,
                         // out_proj ops relative to the step blob; slabs are contiguous
    uint32_t slab_bytes; // n * ks * 2 (both halves in pair mode)
    uint32_t last_bytes; // bytes of the last slab (k_total - (n_slab-1)*ks columns)
    uint16_t n;          // MMA N (rows of W): multiple of 16, 16..256
    uint16_t ks;         // K extent of a full slab: multiple of 16
    uint16_t k_total;    // K extent of the GEMM
    uint16_t a_off;      // QB_A_E: first 8-element k-chunk of A_E;  QB_A_H: first TMEM column (k0 / 2) inside Hacc
    uint16_t d_col;      // TMEM column of the accumulator tile
    uint8_t n_slab;      // ring slots this op consumes
    uint8_t a_src;       // QbASrc
    uint8_t accumulate;  // 0: the first MMA overwrites D, 1: accumulates
    uint8_t wait_a;      // QbBar to wait on before issuing (A operand ready), or NONE
    uint8_t wait_d;      // QbBar to wait on before issuing (accumulator free), or NONE
    uint8_t commit;      // QbBar to tcgen05.commit to after the GEMM, or NONE
    uint8_t wait_a2_slab; // != 0: wait on QB_BAR_AH2_READY before this slab (the A operand arrives in two K halves)
    uint8_t a_blk32;     // QB_A_H only: the packed fp16 operand sits at the START of every 32-column fp32 block it was converted
                         // from (k-step j at column 32 * (j / 2) + 8 * (j % 2)) instead of contiguously (h_split == 2)
    uint8_t pad[2];
};
static_assert(sizeof(QbOp) == 32, "QbOp must stay 32 bytes");

struct QbStepPlan {
    int32_t D, De, Dh, L, K;
    int32_t has_proj;        // De != D: out_proj runs on the tensor core
    int32_t skip;            // QINCo2: o += C_m[code]
    int32_t n_tiles;         // tiles in flight per CTA (1 or 2); tile slot t uses TMEM columns [t*tmem_tile_cols, ...)
    int32_t tmem_alloc_cols; // power of two >= n_tiles * tmem_tile_cols
    int32_t n_ops_block;     // ops[0 .. n_ops_block) run once per residual block l (weights at l*block_w_bytes + w_off)
    int32_t n_ops_out;       // ops[n_ops_block .. n_ops_block+n_ops_out) run once per tile set for out_proj
    int32_t hc;              // H chunk width (columns of Hacc per chunk; the last chunk may be narrower)
    int32_t n_hchunk;        // ceil(Dh / hc)
    int32_t oc;              // out-proj chunk width (columns), has_proj only
    int32_t n_ochunk;
    int32_t tmem_e_col;      // per-tile column offsets
    int32_t tmem_h_col;
    int32_t tmem_tile_cols;  // columns per tile slot
    // shared memory carve-up (byte offsets from the 1024-aligned dynamic smem base)
    int32_t smem_ae[2];      // A_E per tile slot: [De/8][128][16B]
    int32_t smem_tres;       // resident quarter of C_m ([D/4][64][4] fp32; the T_m rows live in registers), or -1 when the
                             // shape does not qualify
    int32_t smem_ring;       // ring of n_stage slots of slot_bytes
    int32_t slot_bytes;
    int32_t n_stage;
    int32_t smem_total;
    int32_t h_split;         // != 0: a full-width H chunk is converted and handed to the MMA issuer in two K halves (two
                             // arrivals on AH_READY per chunk; the down-projection's slabs end on the half boundary).
                             // 2 (hc == 128): every epilogue warp writes its packed fp16 block over the start of the 32 fp32
                             // columns it just read, so the warps of a lane quarter need no barrier between loads and stores
    int32_t pair;            // 1: CTA-pair kernel (cta_group::2, M = 256 over two CTAs): every slab is packed as two row
                             // halves, CTA r of a pair streams half r into a ring slot of slot_bytes
    // Decode loop (kLoop kernel: one tile walks every step, xhat stays with its rows): u = Wx . xhat runs on the tensor core
    // as n_ops_pre "pre-ops" placed AFTER the out_proj ops in the op list (so block / out_proj indices and weight offsets
    // are those of the plain plan): Eacc (initialised with T_m[code]) += [xhat_hi | xhat_lo] . [Wx_hi | Wx_hi]^T
    // + xhat_hi . Wx_lo^T, fp16 hi/lo splits of an fp32 value (~22 mantissa bits).  The A operand [xhat_hi | xhat_lo]
    // (2D/8 k-chunks) is written by the previous step's final epilogue into the A_E buffer, which therefore holds
    // ae_chunks = max(De, 2D) / 8 k-chunks.  0 pre-ops: a plain plan.
    int32_t n_ops_pre;
    int32_t ae_chunks;
    int32_t e_split;         // 1: the last down-projection of a block commits its first N part on QB_BAR_EACC_HALF (two parts: de > 256)
    int32_t epart;           // columns of a down-projection N part
    int32_t mcast;           // 1: non-resident launches run as 2-CTA clusters that multicast the weight slabs (each CTA streams
                             // half of every slab into both ring slots); needs slab halves that are multiples of 16 bytes
    int64_t block_w_bytes;   // packed weight bytes of one residual block
    int64_t w_blob_bytes;    // packed weight bytes for this step (L blocks + out_proj)
};
