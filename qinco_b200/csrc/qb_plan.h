// Step plan shared by the host planner/packer (qb_plan.cpp) and the tcgen05 MLP kernel (qb_mlp.cu).
//
// One quantisation step m >= 1 evaluates, for every candidate row, the implicit-codebook MLP
// (reference qinco/model/qinco_base.py:262-280: in_proj -> QConcat -> L x QBlockFFN -> out_proj [+ skip]).
// After hoisting the candidate-invariant terms:
//     e0  = T_m[code] + u_b                      T_m: [K,De] per-step table, u_b = Wcat[:,De:] . xhat_b
//     e  += Wdn_l . relu(Wup_l . e)   l < L      fp16 operands, fp32 accumulate, fp32 residual stream in TMEM
//     o   = Pout . e (+ C_m[code])               Pout absent when De == D; skip absent in qinco1_mode
// The kernel walks a static list of "ops"; each op is one weight slab (n rows x k cols, fp16, K-major,
// no-swizzle UMMA core-matrix layout [k/8][n][8]) streamed by TMA bulk copy into a ring slot and consumed by
// k/16 tcgen05.mma instructions (M=128).  All L residual blocks share one op list (only the weight offset moves by
// block_w_bytes per block), followed by the out_proj ops; the sequence is identical for every 128-row tile.
#pragma once
#include <stdint.h>

#define QB_TILE_M 128
#define QB_MAX_OPS 256

enum QbABuf : uint16_t { QB_A_E = 0, QB_A_H0 = 1, QB_A_H1 = 2 };
// barrier ids used in op wait/commit fields
enum QbBar : uint8_t {
    QB_BAR_NONE = 0,
    QB_BAR_AE_READY = 1,    // epilogue wrote A_E (fp16 e) [+ initialised Eacc]            count 128
    QB_BAR_AH0_READY = 2,   // epilogue wrote A_H[0] (relu(h) chunk)                       count 128
    QB_BAR_AH1_READY = 3,
    QB_BAR_HACC0_FREE = 4,  // epilogue finished reading Hacc[0]                           count 128
    QB_BAR_HACC1_FREE = 5,
    QB_BAR_HACC0_FULL = 6,  // MMA -> epilogue (tcgen05.commit)                            count 1
    QB_BAR_HACC1_FULL = 7,
    QB_BAR_EACC_FULL = 8,
    QB_BAR_COUNT = 9
};

struct QbOp {
    uint32_t w_off;      // byte offset of the slab (multiple of 16): block ops relative to the block's weights,
                         // out_proj ops relative to the step blob
    uint32_t w_bytes;    // n * k * 2
    uint16_t n;          // MMA N (slab rows): multiple of 16, 16..256
    uint16_t k;          // slab K extent: multiple of 16
    uint16_t a_buf;      // QbABuf
    uint16_t a_kc;       // first 8-element k-chunk of the A buffer this slab multiplies
    uint16_t d_col;      // TMEM column (relative to the allocation base) of the accumulator tile
    uint8_t accumulate;  // 0: first MMA of the slab overwrites D, 1: accumulates
    uint8_t wait_a;      // QbBar to wait on before issuing (A operand ready), or NONE
    uint8_t wait_d;      // QbBar to wait on before issuing (accumulator free), or NONE
    uint8_t commit;      // QbBar to tcgen05.commit to after the slab, or NONE
    uint8_t pad[10];
};
static_assert(sizeof(QbOp) == 32, "QbOp must stay 32 bytes");

// Epilogue program of one tile, derived from the same dimensions:
//   init; for l<L { for j<n_hchunk { H-epilogue(buf = hbuf[j]) }  E-epilogue }  [ out-epilogue over n_ochunk ]
struct QbStepPlan {
    int32_t D, De, Dh, L, K;
    int32_t has_proj;        // De != D: out_proj runs on the tensor core
    int32_t skip;            // QINCo2: o += C_m[code]
    int32_t n_ops_block;     // ops[0 .. n_ops_block) run once per residual block l (weights at l*block_w_bytes + w_off)
    int32_t n_ops_out;       // ops[n_ops_block .. n_ops_block+n_ops_out) run once per tile for out_proj
    int32_t hc;              // H chunk width NC (columns of Hacc per chunk; last chunk may be narrower)
    int32_t n_hchunk;        // ceil(Dh / hc)
    int32_t n_hbuf;          // 1 or 2 Hacc/A_H buffers
    int32_t oc;              // out-proj chunk width (columns), has_proj only
    int32_t n_ochunk;
    int32_t tmem_e_col;      // 0
    int32_t tmem_h_col[2];   // Hacc[0], Hacc[1]
    // shared memory carve-up (byte offsets from the 1024-aligned dynamic smem base)
    int32_t smem_ae;         // A_E  [De/8][128][16B]
    int32_t smem_ah[2];      // A_H  [hc/8][128][16B]
    int32_t smem_ring;       // ring of n_stage slots of slot_bytes
    int32_t slot_bytes;
    int32_t n_stage;
    int32_t smem_total;
    int64_t block_w_bytes;   // packed weight bytes of one residual block
    int64_t w_blob_bytes;    // packed weight bytes for this step (L blocks + out_proj)
};
