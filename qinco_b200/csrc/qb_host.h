// Host-side declarations shared by the planner (qb_plan.cpp), the C-ABI (qb_api.cu) and the plan tests.
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

#include "qb_plan.h"

namespace qb {

struct PlanOptions {
    int hc = 0;           // H chunk width (0 = auto)
    int n_tiles = 0;      // 128-row tiles in flight per CTA (0 = auto: 2 when TMEM and shared memory allow)
    int pair = 0;         // CTA-pair (cta_group::2) kernel: 0 = auto (currently off), 1 = off, 2 = on (needs N % 32 == 0 down-projections)
    int no_resident = 0;  // 1: never keep table halves resident (debug / A-B tests)
    int blk32 = 0;        // 1: blocked packed-H layout for hc == 128 (QbStepPlan::h_split == 2; measured 0.5 % slower than the
                          // contiguous layout + quarter barrier with 8 epilogue warps, kept for A-B tests)
    int no_hsplit = 0;    // 1: hand every H chunk to the MMA issuer in one piece (debug / A-B tests)
    int slot_bytes = 0;   // weight ring slot size (0 = 16 KiB)
    int max_stage = 0;    // cap on ring depth (0 = QB_MAX_STAGE)
    int max_slab_k = 0;   // cap on slab K (0 = what fits a slot)
    int smem_budget = 0;  // bytes of dynamic shared memory the kernel may use (0 = 208 KiB)
    int no_esplit = 0;    // 1: the E epilogue waits for the whole residual accumulator (debug / A-B tests)
    int mcast = 0;        // weight multicast over 2-CTA clusters: 0 = auto, 1 = off, 2 = on
    int uop = 0;          // 1: decode-loop plan: pre-ops computing u = Wx . xhat on the tensor core (fp16 hi/lo split), A_E
                          // buffer sized for [xhat_hi | xhat_lo], no resident tables
};

uint16_t f32_to_f16(float f);
float f16_to_f32(uint16_t h);

int make_step_plan(int D, int De, int Dh, int L, int K, int qinco1_mode, const PlanOptions& opt, QbStepPlan* plan,
                   std::vector<QbOp>* ops, std::string* err);

// which compile-time plan view (qb_mlp.cu: PlanView<kShape>) the kernels may use for this plan: 0 generic, 1 S128, 2 L384
int mlp_plan_view(const QbStepPlan& plan);

int pack_step_weights(const QbStepPlan& plan, const std::vector<QbOp>& ops, const float* const* up,
                      const float* const* down, const float* out_proj, uint16_t* blob, std::string* err);

// slabs of the pre-ops (decode-loop plans): wx is Wcat[:, De:] as [De][D] rows
int pack_pre_weights(const QbStepPlan& plan, const std::vector<QbOp>& ops, const float* wx, uint16_t* blob, std::string* err);

// operand blob of the tensor-core beam preparation (layout: qb_dev.h, PrepTcParams): w is [n_rows][D] fp32 row-major
size_t prep_pack_bytes(int n_rows, int D);
void prep_pack(const float* w, int n_rows, int D, uint16_t* out);

// operand blob of the tensor-core IVF arg-min (layout: qb_dev.h, IvfTcParams)
size_t ivf_pack_bytes(int ivf_K, int D);
void ivf_pack(const float* cent, int ivf_K, int D, uint8_t* out);

void build_tables(int D, int De, int K, const float* codebook, const float* in_proj, const float* concat_w,
                  const float* concat_b, float* t_blk, float* cb_blk, float* wx_t);

}  // namespace qb
