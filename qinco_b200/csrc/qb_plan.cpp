// Host-side step planner and weight packer for the tcgen05 MLP kernel (no CUDA dependency).
// See qb_plan.h for the dataflow.  The reference computes the same layers with separate cuBLAS GEMMs
// (reference qinco/model/qinco_base.py:60-64, 93-97, 238-246, 262-280); here they become one op list per step.
#include "qb_plan.h"
#include "qb_host.h"

#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

namespace qb {

uint16_t f32_to_f16(float f) {
    // round-to-nearest-even, IEEE binary16, matches cvt.rn.f16.f32
    uint32_t x;
    std::memcpy(&x, &f, 4);
    uint32_t sign = (x >> 16) & 0x8000u;
    uint32_t exp = (x >> 23) & 0xffu;
    uint32_t man = x & 0x7fffffu;
    if (exp == 0xff) return (uint16_t)(sign | 0x7c00u | (man ? 0x200u : 0));
    int32_t e = (int32_t)exp - 127 + 15;
    if (e >= 31) return (uint16_t)(sign | 0x7c00u);
    if (e <= 0) {
        if (e < -10) return (uint16_t)sign;
        man |= 0x800000u;
        uint32_t shift = (uint32_t)(14 - e);
        uint32_t half = man >> shift;
        uint32_t rem = man & ((1u << shift) - 1);
        uint32_t halfway = 1u << (shift - 1);
        if (rem > halfway || (rem == halfway && (half & 1))) half++;
        return (uint16_t)(sign | half);
    }
    uint32_t half = (uint32_t)(e << 10) | (man >> 13);
    uint32_t rem = man & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (half & 1))) half++;
    return (uint16_t)(sign | half);
}

float f16_to_f32(uint16_t h) {
    uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1fu;
    uint32_t man = h & 0x3ffu;
    uint32_t x;
    if (exp == 0) {
        if (man == 0) x = sign;
        else {
            int e = -1;
            do { man <<= 1; e++; } while (!(man & 0x400u));
            x = sign | (uint32_t)(127 - 15 - e) << 23 | ((man & 0x3ffu) << 13);
        }
    } else if (exp == 31) x = sign | 0x7f800000u | (man << 13);
    else x = sign | ((exp - 15 + 127) << 23) | (man << 13);
    float f;
    std::memcpy(&f, &x, 4);
    return f;
}

static int round_up(int a, int b) { return (a + b - 1) / b * b; }

int make_step_plan(int D, int De, int Dh, int L, int K, int qinco1_mode, const PlanOptions& opt,
                   QbStepPlan* p, std::vector<QbOp>* ops, std::string* err) {
    std::memset(p, 0, sizeof(*p));
    if (D % 16 || De % 16 || Dh % 16 || D < 16 || De < 16 || (L > 0 && Dh < 16)) {
        *err = "D, de and dh must be multiples of 16 (tcgen05 kind::f16 needs K%16==0 and N%16==0 at M=128)";
        return -1;
    }
    if (K < 1 || K > 256) { *err = "K must be in [1,256] (codes are uint8)"; return -1; }
    p->D = D; p->De = De; p->Dh = Dh; p->L = L; p->K = K;
    p->has_proj = De != D;
    p->skip = qinco1_mode ? 0 : 1;

    // ---- TMEM columns: Eacc [0,De) | Hacc0 | Hacc1 -------------------------------------------------------------
    const int e_cols = round_up(De, 32);
    int hc = opt.hc > 0 ? opt.hc : 128;
    int n_hbuf = opt.n_hbuf > 0 ? opt.n_hbuf : 2;
    const int h_need = std::max(L > 0 ? Dh : 0, p->has_proj ? D : 0);
    hc = std::min(hc, round_up(std::max(h_need, 16), 16));
    if (opt.hc <= 0 || opt.n_hbuf <= 0) {            // auto: prefer two 128-wide buffers, then one, then two 64-wide
        if (e_cols + 2 * round_up(hc, 32) > 512) {
            if (e_cols + round_up(hc, 32) <= 512) n_hbuf = 1;
            else { hc = 64; n_hbuf = (e_cols + 128 <= 512) ? 2 : 1; }
        }
    }
    if (hc % 16 || hc > 256) { *err = "hc must be a multiple of 16 and <= 256"; return -1; }
    {   // a second buffer only helps when some phase has more than one chunk
        const int nh = L > 0 ? (Dh + hc - 1) / hc : 0;
        const int no = p->has_proj ? (D + std::min(D, hc) - 1) / std::min(D, hc) : 0;
        if (std::max(nh, no) <= 1) n_hbuf = 1;
    }
    if (e_cols + n_hbuf * round_up(hc, 32) > 512) {
        *err = "de too large for TMEM: round32(de) + n_hbuf*round32(hc) must be <= 512 columns";
        return -1;
    }
    p->hc = hc; p->n_hbuf = n_hbuf;
    p->n_hchunk = L > 0 ? (Dh + hc - 1) / hc : 0;
    p->oc = p->has_proj ? std::min(D, hc) : 0;
    p->n_ochunk = p->has_proj ? (D + p->oc - 1) / p->oc : 0;
    p->tmem_e_col = 0;
    p->tmem_h_col[0] = e_cols;
    p->tmem_h_col[1] = e_cols + round_up(hc, 32);

    // ---- shared memory ------------------------------------------------------------------------------------------
    const int a_kc_bytes = QB_TILE_M * 16;  // one 8-element k-chunk of a 128-row A operand
    int off = 0;
    p->smem_ae = off; off += (De / 8) * a_kc_bytes;
    for (int i = 0; i < 2; i++) { p->smem_ah[i] = off; if (i < n_hbuf && L > 0) off += (hc / 8) * a_kc_bytes; }
    p->slot_bytes = opt.slot_bytes > 0 ? opt.slot_bytes : 16384;
    if (p->slot_bytes % 1024) { *err = "slot_bytes must be a multiple of 1024"; return -1; }
    p->smem_ring = off;
    const int budget = opt.smem_budget > 0 ? opt.smem_budget : 216 * 1024;
    int n_stage = (budget - off) / p->slot_bytes;
    n_stage = std::min(n_stage, opt.max_stage > 0 ? opt.max_stage : 8);
    if (n_stage < 2) { *err = "not enough shared memory for a 2-slot weight ring (de/dh too large)"; return -1; }
    p->n_stage = n_stage;
    p->smem_total = off + n_stage * p->slot_bytes;

    // ---- op list ------------------------------------------------------------------------------------------------
    ops->clear();
    uint32_t w_off = 0;
    auto slab_k = [&](int n, int k_total) {
        int k = (p->slot_bytes / (2 * n)) / 16 * 16;
        k = std::min(k, opt.max_slab_k > 0 ? opt.max_slab_k : 128);
        return std::max(16, std::min(k, k_total));
    };
    // GEMM  D[128, n] (+)= A[128, k_total] . W[rows n][cols k_total]^T, split into K slabs
    auto emit_gemm = [&](int n, int k_total, uint16_t a_buf, int a_kc0, int d_col, bool acc_first,
                         uint8_t wait_a, uint8_t wait_d, uint8_t commit) {
        const int ks = slab_k(n, k_total);
        for (int k0 = 0; k0 < k_total; k0 += ks) {
            QbOp op;
            std::memset(&op, 0, sizeof(op));
            const int k = std::min(ks, k_total - k0);
            op.w_off = w_off; op.w_bytes = (uint32_t)(n * k * 2);
            w_off += op.w_bytes;
            op.n = (uint16_t)n; op.k = (uint16_t)k; op.a_buf = a_buf; op.a_kc = (uint16_t)(a_kc0 + k0 / 8);
            op.d_col = (uint16_t)d_col;
            op.accumulate = (acc_first || k0 > 0) ? 1 : 0;
            op.wait_a = (k0 == 0) ? wait_a : QB_BAR_NONE;
            op.wait_d = (k0 == 0) ? wait_d : QB_BAR_NONE;
            op.commit = (k0 + k >= k_total) ? commit : QB_BAR_NONE;
            ops->push_back(op);
        }
    };
    const int n_eparts = (De + 255) / 256;
    const int epart = round_up((De + n_eparts - 1) / n_eparts, 16);
    if (L > 0) {                   // ops of ONE residual block, offsets relative to the block's weights
        auto m1 = [&](int j) {
            const int cw = std::min(hc, Dh - j * hc), buf = j % n_hbuf;
            emit_gemm(cw, De, QB_A_E, 0, p->tmem_h_col[buf], false, j == 0 ? QB_BAR_AE_READY : QB_BAR_NONE,
                      (uint8_t)(QB_BAR_HACC0_FREE + buf), (uint8_t)(QB_BAR_HACC0_FULL + buf));
        };
        auto m2 = [&](int j) {
            const int cw = std::min(hc, Dh - j * hc), buf = j % n_hbuf;
            for (int n0 = 0; n0 < De; n0 += epart) {
                const int n = std::min(epart, De - n0);
                const bool last = (j == p->n_hchunk - 1) && (n0 + n >= De);
                emit_gemm(n, cw, (uint16_t)(QB_A_H0 + buf), 0, p->tmem_e_col + n0, true,
                          n0 == 0 ? (uint8_t)(QB_BAR_AH0_READY + buf) : QB_BAR_NONE, QB_BAR_NONE,
                          last ? QB_BAR_EACC_FULL : QB_BAR_NONE);
            }
        };
        if (n_hbuf == 2) {          // software pipeline: MMA1(j+1) is issued before MMA2(j)
            m1(0);
            for (int j = 0; j < p->n_hchunk; j++) { if (j + 1 < p->n_hchunk) m1(j + 1); m2(j); }
        } else {
            for (int j = 0; j < p->n_hchunk; j++) { m1(j); m2(j); }
        }
    }
    p->n_ops_block = (int)ops->size();
    p->block_w_bytes = w_off;
    w_off = (uint32_t)(p->block_w_bytes * L);      // out_proj slabs follow the L blocks
    for (int q = 0; q < p->n_ochunk; q++) {
        const int cw = std::min(p->oc, D - q * p->oc), buf = q % n_hbuf;
        emit_gemm(cw, De, QB_A_E, 0, p->tmem_h_col[buf], false, q == 0 ? QB_BAR_AE_READY : QB_BAR_NONE,
                  (uint8_t)(QB_BAR_HACC0_FREE + buf), (uint8_t)(QB_BAR_HACC0_FULL + buf));
    }
    p->n_ops_out = (int)ops->size() - p->n_ops_block;
    if (ops->size() > QB_MAX_OPS) { *err = "op list too long"; return -1; }
    for (auto& op : *ops)
        if ((int)op.w_bytes > p->slot_bytes) { *err = "internal: slab larger than ring slot"; return -1; }
    p->w_blob_bytes = std::max<int64_t>(w_off, 16);
    return 0;
}

int pack_step_weights(const QbStepPlan& p, const std::vector<QbOp>& ops, const float* const* up,
                      const float* const* down, const float* out_proj, uint16_t* blob, std::string* err) {
    // Re-walk the emission order to know which weight sub-matrix every op covers.
    const int D = p.D, De = p.De, Dh = p.Dh, hc = p.hc;
    const int n_eparts = (De + 255) / 256;
    const int epart = round_up((De + n_eparts - 1) / n_eparts, 16);
    auto put = [&](const QbOp& op, size_t base, const float* w, int ld, int row0, int col0) {
        uint16_t* dst = blob + (base + op.w_off) / 2;
        for (int k = 0; k < op.k; k++)
            for (int r = 0; r < op.n; r++)
                dst[((size_t)(k / 8) * op.n + r) * 8 + (k % 8)] = f32_to_f16(w[(size_t)(row0 + r) * ld + col0 + k]);
    };
    for (int l = 0; l <= p.L; l++) {
        const bool out_phase = (l == p.L);
        size_t cursor = out_phase ? (size_t)p.n_ops_block : 0;
        const size_t end = out_phase ? (size_t)(p.n_ops_block + p.n_ops_out) : (size_t)p.n_ops_block;
        const size_t base = out_phase ? 0 : (size_t)l * (size_t)p.block_w_bytes;
        int rc = 0;
        auto take = [&](const float* w, int ld, int row0, int n, int col0, int k_total) {
            int k0 = 0;
            while (k0 < k_total) {
                if (cursor >= end) return -1;
                const QbOp& op = ops[cursor++];
                if (op.n != n) return -1;
                put(op, base, w, ld, row0, col0 + k0);
                k0 += op.k;
            }
            return k0 == k_total ? 0 : -1;
        };
        if (!out_phase) {
            auto m1 = [&](int j) { return take(up[l], De, j * hc, std::min(hc, Dh - j * hc), 0, De); };
            auto m2 = [&](int j) {
                const int cw = std::min(hc, Dh - j * hc);
                for (int n0 = 0; n0 < De; n0 += epart)
                    if (take(down[l], Dh, n0, std::min(epart, De - n0), j * hc, cw)) return -1;
                return 0;
            };
            if (p.n_hbuf == 2) {
                rc |= m1(0);
                for (int j = 0; j < p.n_hchunk; j++) { if (j + 1 < p.n_hchunk) rc |= m1(j + 1); rc |= m2(j); }
            } else {
                for (int j = 0; j < p.n_hchunk; j++) { rc |= m1(j); rc |= m2(j); }
            }
        } else {
            for (int q = 0; q < p.n_ochunk; q++) rc |= take(out_proj, De, q * p.oc, std::min(p.oc, D - q * p.oc), 0, De);
        }
        if (rc || cursor != end) { *err = "internal: pack order does not match the op list"; return -1; }
    }
    return 0;
}

// T_m[k] = e0 + Wcat[:, :De] . e0 + bcat,  e0 = Pin . C_m[k]   (double accumulation, stored fp32, blocked [De/8][K][8])
void build_tables(int D, int De, int K, const float* codebook, const float* in_proj, const float* concat_w,
                  const float* concat_b, float* t_blk, float* cb_blk, float* wx_t) {
    std::vector<double> e0(De), t(De);
    for (int k = 0; k < K; k++) {
        const float* c = codebook + (size_t)k * D;
        for (int e = 0; e < De; e++) {
            if (in_proj) {
                double s = 0;
                for (int d = 0; d < D; d++) s += (double)in_proj[(size_t)e * D + d] * c[d];
                e0[e] = (double)(float)s;      // the reference materialises in_proj(c) in fp32
            } else e0[e] = c[e];
        }
        for (int e = 0; e < De; e++) {
            const float* wrow = concat_w + (size_t)e * (De + D);
            double s = concat_b[e];
            for (int j = 0; j < De; j++) s += (double)wrow[j] * e0[j];
            t[e] = e0[e] + s;
        }
        for (int e = 0; e < De; e++) t_blk[((size_t)(e / 8) * K + k) * 8 + (e % 8)] = (float)t[e];
        for (int d = 0; d < D; d++) cb_blk[((size_t)(d / 8) * K + k) * 8 + (d % 8)] = c[d];
    }
    // Wx^T [D][De]: u = Wcat[:, De:] . xhat
    for (int e = 0; e < De; e++)
        for (int d = 0; d < D; d++) wx_t[(size_t)d * De + e] = concat_w[(size_t)e * (De + D) + De + d];
}

}  // namespace qb

// ---- host-only test hooks (declared in include/qinco_b200.h): let the CPU test-suite check the planner and the packer
// by replaying the op list in numpy, without a GPU.
extern "C" {

int qb_plan_export(int D, int De, int Dh, int L, int K, int qinco1_mode, const int32_t* opts5, int32_t* plan_out,
                   int n_plan_out, void* ops_out, int max_ops) {
    qb::PlanOptions opt;
    if (opts5) {
        opt.hc = opts5[0]; opt.n_hbuf = opts5[1]; opt.slot_bytes = opts5[2]; opt.max_stage = opts5[3];
        opt.max_slab_k = opts5[4];
    }
    QbStepPlan p;
    std::vector<QbOp> ops;
    std::string err;
    if (qb::make_step_plan(D, De, Dh, L, K, qinco1_mode, opt, &p, &ops, &err)) return -1;
    const int32_t v[] = {p.D, p.De, p.Dh, p.L, p.K, p.has_proj, p.skip, p.n_ops_block, p.n_ops_out, p.hc, p.n_hchunk,
                         p.n_hbuf, p.oc, p.n_ochunk, p.tmem_e_col, p.tmem_h_col[0], p.tmem_h_col[1], p.smem_ae,
                         p.smem_ah[0], p.smem_ah[1], p.smem_ring, p.slot_bytes, p.n_stage, p.smem_total,
                         (int32_t)p.block_w_bytes, (int32_t)p.w_blob_bytes};
    const int nv = (int)(sizeof(v) / sizeof(v[0]));
    for (int i = 0; i < nv && i < n_plan_out; i++) plan_out[i] = v[i];
    if ((int)ops.size() > max_ops) return -2;
    if (ops_out) std::memcpy(ops_out, ops.data(), ops.size() * sizeof(QbOp));
    return (int)ops.size();
}

int qb_plan_pack(int D, int De, int Dh, int L, int K, int qinco1_mode, const int32_t* opts5, const float* const* up,
                 const float* const* down, const float* out_proj, uint16_t* blob, int64_t blob_halfs) {
    qb::PlanOptions opt;
    if (opts5) {
        opt.hc = opts5[0]; opt.n_hbuf = opts5[1]; opt.slot_bytes = opts5[2]; opt.max_stage = opts5[3];
        opt.max_slab_k = opts5[4];
    }
    QbStepPlan p;
    std::vector<QbOp> ops;
    std::string err;
    if (qb::make_step_plan(D, De, Dh, L, K, qinco1_mode, opt, &p, &ops, &err)) return -1;
    if (blob_halfs * 2 < p.w_blob_bytes) return -2;
    return qb::pack_step_weights(p, ops, up, down, out_proj, blob, &err);
}

int qb_plan_tables(int D, int De, int K, const float* codebook, const float* in_proj, const float* concat_w,
                   const float* concat_b, float* t_blk, float* cb_blk, float* wx_t) {
    qb::build_tables(D, De, K, codebook, in_proj, concat_w, concat_b, t_blk, cb_blk, wx_t);
    return 0;
}

}  // extern "C"
