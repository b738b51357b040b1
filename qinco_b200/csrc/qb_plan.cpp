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

    // ---- TMEM columns per tile: Eacc [0,De) | Hacc [e_cols, e_cols + hc) ------------------------------------------
    const int e_cols = round_up(De, 32);
    const int h_need = std::max(L > 0 ? Dh : 0, p->has_proj ? D : 0);
    int hc = opt.hc > 0 ? opt.hc : 128;
    hc = std::min(hc, round_up(std::max(h_need, 16), 16));
    if (hc % 16 || hc > 256) { *err = "hc must be a multiple of 16 and <= 256"; return -1; }
    const int h_cols = h_need > 0 ? round_up(hc, 32) : 0;
    if (e_cols + h_cols > 512) { *err = "de too large for TMEM: round32(de) + round32(hc) must be <= 512 columns"; return -1; }
    const int a_kc_bytes = QB_TILE_M * 16;  // one 8-element k-chunk of a 128-row A operand
    p->ae_chunks = (opt.uop ? std::max(De, 2 * D) : De) / 8;
    const int ae_bytes = p->ae_chunks * a_kc_bytes;
    p->slot_bytes = opt.slot_bytes > 0 ? opt.slot_bytes : 32768;
    if (p->slot_bytes % 1024) { *err = "slot_bytes must be a multiple of 1024"; return -1; }
    // two tiles in flight share every weight slab: they need 2x the TMEM columns and 2x the A_E tile
    int n_tiles = opt.n_tiles > 0 ? opt.n_tiles : 2;
    if (n_tiles > 2) { *err = "n_tiles must be 1 or 2"; return -1; }
    int budget = opt.smem_budget > 0 ? opt.smem_budget : 208 * 1024;   // + <= 15 KB static (barriers, beam rows) <= 227 KB
    // shapes the fused-selection-B kernel serves (one tile per CTA, whole out_proj in one chunk) leave it 16 KB for its stash
    if (opt.smem_budget <= 0 && De != D && D <= (opt.hc > 0 ? opt.hc : 128) && 2 * (e_cols + h_cols) > 512) budget = 192 * 1024;
    if (n_tiles == 2 && (2 * (e_cols + h_cols) > 512 || budget - 2 * ae_bytes < 4 * p->slot_bytes)) {
        if (opt.n_tiles == 2) { *err = "two tiles in flight do not fit (TMEM columns or shared memory)"; return -1; }
        n_tiles = 1;
    }
    p->n_tiles = n_tiles;
    {
        int c = 32;
        while (c < n_tiles * (e_cols + h_cols)) c *= 2;
        p->tmem_alloc_cols = c;
    }
    p->hc = hc;
    p->n_hchunk = L > 0 ? (Dh + hc - 1) / hc : 0;
    p->oc = p->has_proj ? std::min(D, hc) : 0;
    p->n_ochunk = p->has_proj ? (D + p->oc - 1) / p->oc : 0;
    p->tmem_e_col = 0;
    p->tmem_h_col = e_cols;
    p->tmem_tile_cols = e_cols + h_cols;

    // ---- shared memory ------------------------------------------------------------------------------------------
    int off = 0;
    for (int t = 0; t < 2; t++) { p->smem_ae[t] = off; if (t < n_tiles) off += ae_bytes; }
    // Resident tables (score launches with all K = 256 candidates per beam): a CTA then only ever works on ONE quarter of
    // the codes (64 codes x 2 beams per tile), whose rows of T_m (registers) and of the skip codebook (shared memory) stay
    // on the SM instead of being gathered from L2 for every tile (the gathers cost as much L2->SM bandwidth as the weights).
    p->smem_tres = -1;
    if (n_tiles == 2 && !p->has_proj && K == 256 && De <= 128 && budget - off - D * 256 >= 4 * 16384 && opt.no_resident == 0 && !opt.uop) {
        p->smem_tres = off;
        off += D * 256;          // the skip-codebook quarter; the T_m row slice of a thread (<= 64 columns) sits in registers
        if (opt.slot_bytes <= 0) p->slot_bytes = 16384;      // a deeper ring of smaller slabs fits next to the tables
    }
    p->smem_ring = off;
    int n_stage = (budget - off) / p->slot_bytes;
    n_stage = std::min(n_stage, opt.max_stage > 0 ? std::min(opt.max_stage, QB_MAX_STAGE) : QB_MAX_STAGE);
    if (n_stage < 2) { *err = "not enough shared memory for a 2-slot weight ring (de too large)"; return -1; }
    p->n_stage = n_stage;
    p->smem_total = off + n_stage * p->slot_bytes;

    // ---- CTA pair: the M = 256 MMA needs N % 16 == 0 (shared-memory A) / N % 32 == 0 (TMEM A); N is the full op width
    {
        const int n_eparts = (De + 255) / 256;
        const int epart = round_up((De + n_eparts - 1) / n_eparts, 16);
        bool ok = true;
        if (L > 0)
            for (int n0 = 0; n0 < De; n0 += epart) ok = ok && (std::min(epart, De - n0) % 32 == 0);
        if (opt.pair == 2 && !ok) { *err = "pair mode needs every down-projection width to be a multiple of 32"; return -1; }
        // Auto = off: the pair kernel is validated (parity tests run it) but the extra cross-CTA signalling latency makes it
        // slower than two independent CTAs while the hand-off chain, not weight traffic, bounds the kernel.
        p->pair = (opt.pair == 2) ? 1 : 0;
    }

    // ---- op list ------------------------------------------------------------------------------------------------
    ops->clear();
    uint32_t w_off = 0;
    // GEMM  D[128, n] (+)= A[128, k_total] . W[rows n][cols k_total]^T, W cut along K into ring-slot-sized slabs
    // H chunks of the full width hc (a multiple of 64) are handed over in two K halves: the down-projection starts on
    // the first half while the epilogue still converts the second (shortens the per-tile dependency chain)
    p->h_split = (opt.no_hsplit == 0 && L > 0 && hc % 64 == 0 && hc >= 64) ? 1 : 0;
    if (p->h_split && hc == 128 && !p->pair && opt.blk32) p->h_split = 2;
    auto emit_gemm = [&](int n, int k_total, uint8_t a_src, int a_unit0, int d_col, bool acc_first,
                         uint8_t wait_a, uint8_t wait_d, uint8_t commit, bool split_k = false) {
        int ks = ((p->pair ? 2 : 1) * p->slot_bytes / (2 * n)) / 16 * 16;   // a pair CTA holds half the rows of a slab
        if (opt.max_slab_k > 0) ks = std::min(ks, opt.max_slab_k);
        ks = std::max(16, std::min(ks, k_total));
        if (split_k) {       // slabs must end on the half boundary: the largest multiple of 16 <= ks dividing k_total / 2
            const int half = k_total / 2;
            while (half % ks) ks -= 16;
        }
        const int n_slab = (k_total + ks - 1) / ks;
        QbOp op;
        std::memset(&op, 0, sizeof(op));
        op.w_off = w_off;
        op.slab_bytes = (uint32_t)(n * ks * 2);
        op.last_bytes = (uint32_t)(n * (k_total - (n_slab - 1) * ks) * 2);
        op.n = (uint16_t)n; op.ks = (uint16_t)ks; op.k_total = (uint16_t)k_total;
        op.a_off = (uint16_t)a_unit0; op.d_col = (uint16_t)d_col;
        op.n_slab = (uint8_t)n_slab; op.a_src = a_src; op.accumulate = acc_first ? 1 : 0;
        op.wait_a = wait_a; op.wait_d = wait_d; op.commit = commit;
        op.wait_a2_slab = (split_k && wait_a) ? (uint8_t)((k_total / 2) / ks) : 0;
        op.a_blk32 = (split_k && a_src == QB_A_H && p->h_split == 2) ? 1 : 0;
        w_off += (uint32_t)(n * k_total * 2);
        ops->push_back(op);
    };
    const int n_eparts = (De + 255) / 256;
    const int epart = round_up((De + n_eparts - 1) / n_eparts, 16);
    p->epart = epart;
    p->e_split = (L > 0 && n_eparts == 2 && !p->pair && !opt.no_esplit) ? 1 : 0;
    if (L > 0) {                   // ops of ONE residual block, offsets relative to the block's weights
        for (int j = 0; j < p->n_hchunk; j++) {
            const int cw = std::min(hc, Dh - j * hc);
            // up-projection chunk: Hacc = A_E . Wup[j*hc .. +cw, :]^T
            emit_gemm(cw, De, QB_A_E, 0, p->tmem_h_col, false, j == 0 ? QB_BAR_AE_READY : QB_BAR_NONE, QB_BAR_NONE,
                      QB_BAR_HACC_FULL);
            // down-projection of the chunk: Eacc[:, n0..] += relu(h)[:, chunk] . Wdn[n0.., j*hc .. +cw]^T
            for (int n0 = 0; n0 < De; n0 += epart) {
                const int n = std::min(epart, De - n0);
                const bool last = (j == p->n_hchunk - 1) && (n0 + n >= De);
                const bool half = p->e_split && (j == p->n_hchunk - 1) && n0 == 0;
                emit_gemm(n, cw, QB_A_H, p->tmem_h_col, p->tmem_e_col + n0, true,
                          n0 == 0 ? QB_BAR_AH_READY : QB_BAR_NONE, QB_BAR_NONE,
                          last ? QB_BAR_EACC_FULL : (half ? QB_BAR_EACC_HALF : QB_BAR_NONE), p->h_split && cw == hc);
            }
        }
    }
    p->n_ops_block = (int)ops->size();
    p->block_w_bytes = w_off;
    w_off = (uint32_t)(p->block_w_bytes * L);      // out_proj slabs follow the L blocks
    for (int q = 0; q < p->n_ochunk; q++) {
        const int cw = std::min(p->oc, D - q * p->oc);
        emit_gemm(cw, De, QB_A_E, 0, p->tmem_h_col, false, q == 0 ? QB_BAR_AE_READY : QB_BAR_NONE,
                  q > 0 ? QB_BAR_HACC_FREE : QB_BAR_NONE, QB_BAR_HACC_FULL);
    }
    p->n_ops_out = (int)ops->size() - p->n_ops_block;
    // pre-ops of the decode loop: per column part of Eacc, [xhat_hi | xhat_lo] . [Wx_hi | Wx_hi]^T then xhat_hi . Wx_lo^T
    if (opt.uop) {
        if (p->pair) { *err = "the decode-loop plan has no CTA-pair variant"; return -1; }
        if (p->n_ochunk > 1) { *err = "the decode-loop plan needs the whole out_proj in one chunk (D <= hc)"; return -1; }
        for (int n0 = 0; n0 < De; n0 += epart) {
            const int n = std::min(epart, De - n0);
            const bool last = n0 + n >= De;
            emit_gemm(n, 2 * D, QB_A_E, 0, p->tmem_e_col + n0, true, n0 == 0 ? QB_BAR_AE_READY : QB_BAR_NONE, QB_BAR_NONE, QB_BAR_NONE);
            emit_gemm(n, D, QB_A_E, 0, p->tmem_e_col + n0, true, QB_BAR_NONE, QB_BAR_NONE, last ? QB_BAR_EACC_FULL : QB_BAR_NONE);
        }
    }
    p->n_ops_pre = (int)ops->size() - p->n_ops_block - p->n_ops_out;
    // Weight multicast (2-CTA clusters, each CTA streams half of every slab into both rings): auto = on for the shapes
    // that run one tile per CTA (de = 384 ...), whose weight stream sits at the L2 -> SM limit; off with CTA pairs.
    p->mcast = (!p->pair && (opt.mcast == 2 || (opt.mcast == 0 && n_tiles == 1))) ? 1 : 0;
    if (ops->size() > QB_MAX_OPS) { *err = "op list too long"; return -1; }
    for (const QbOp& op : *ops)
        if ((int)op.slab_bytes > (p->pair ? 2 : 1) * p->slot_bytes || op.n_slab < 1) { *err = "internal: slab larger than ring slot"; return -1; }
    p->w_blob_bytes = std::max<int64_t>(w_off, 16);
    return 0;
}

int pack_step_weights(const QbStepPlan& p, const std::vector<QbOp>& ops, const float* const* up,
                      const float* const* down, const float* out_proj, uint16_t* blob, std::string* err) {
    // Re-walk the emission order to know which weight sub-matrix every op covers.
    const int D = p.D, De = p.De, Dh = p.Dh, hc = p.hc;
    const int n_eparts = (De + 255) / 256;
    const int epart = round_up((De + n_eparts - 1) / n_eparts, 16);
    // slab s of an op holds columns [s*ks, ...) of W rows [row0, row0+n) as [k/8][n][8]
    auto put = [&](const QbOp& op, size_t base, const float* w, int ld, int row0, int col0) {
        for (int s = 0; s < op.n_slab; s++) {
            uint16_t* dst = blob + (base + op.w_off + (size_t)s * op.slab_bytes) / 2;
            const int k0 = s * op.ks, kn = std::min<int>(op.ks, op.k_total - k0);
            // pair mode: rows [0, n/2) then rows [n/2, n), each half as [k/8][n/2][8]
            const int nh = p.pair ? op.n / 2 : op.n;
            for (int k = 0; k < kn; k++)
                for (int r = 0; r < op.n; r++)
                    dst[(size_t)(r / nh) * nh * kn + ((size_t)(k / 8) * nh + (r % nh)) * 8 + (k % 8)] =
                        f32_to_f16(w[(size_t)(row0 + r) * ld + col0 + k0 + k]);
        }
    };
    for (int l = 0; l <= p.L; l++) {
        const bool out_phase = (l == p.L);
        size_t cursor = out_phase ? (size_t)p.n_ops_block : 0;
        const size_t end = out_phase ? (size_t)(p.n_ops_block + p.n_ops_out) : (size_t)p.n_ops_block;
        const size_t base = out_phase ? 0 : (size_t)l * (size_t)p.block_w_bytes;
        int rc = 0;
        auto take = [&](const float* w, int ld, int row0, int n, int col0, int k_total) {
            if (cursor >= end) return -1;
            const QbOp& op = ops[cursor++];
            if (op.n != n || op.k_total != k_total) return -1;
            put(op, base, w, ld, row0, col0);
            return 0;
        };
        if (!out_phase) {
            for (int j = 0; j < p.n_hchunk; j++) {
                const int cw = std::min(hc, Dh - j * hc);
                rc |= take(up[l], De, j * hc, cw, 0, De);
                for (int n0 = 0; n0 < De; n0 += epart) rc |= take(down[l], Dh, n0, std::min(epart, De - n0), j * hc, cw);
            }
        } else {
            for (int q = 0; q < p.n_ochunk; q++) rc |= take(out_proj, De, q * p.oc, std::min(p.oc, D - q * p.oc), 0, De);
        }
        if (rc || cursor != end) { *err = "internal: pack order does not match the op list"; return -1; }
    }
    return 0;
}

int pack_pre_weights(const QbStepPlan& p, const std::vector<QbOp>& ops, const float* wx, uint16_t* blob, std::string* err) {
    const int D = p.D, De = p.De;
    const int n_eparts = (De + 255) / 256;
    const int epart = round_up((De + n_eparts - 1) / n_eparts, 16);
    size_t cursor = (size_t)(p.n_ops_block + p.n_ops_out);
    const size_t end = cursor + (size_t)p.n_ops_pre;
    // column k of the op's weight matrix -> (input dimension, hi or lo part of Wx)
    auto put = [&](const QbOp& op, int row0, bool lo_only) {
        for (int s = 0; s < op.n_slab; s++) {
            uint16_t* dst = blob + (op.w_off + (size_t)s * op.slab_bytes) / 2;
            const int k0 = s * op.ks, kn = std::min<int>(op.ks, op.k_total - k0);
            for (int k = 0; k < kn; k++) {
                const int d = (k0 + k) % D;
                for (int r = 0; r < op.n; r++) {
                    const float w = wx[(size_t)(row0 + r) * D + d];
                    const uint16_t hi = f32_to_f16(w);
                    const uint16_t v = lo_only ? f32_to_f16(w - f16_to_f32(hi)) : hi;
                    dst[((size_t)(k / 8) * op.n + r) * 8 + (k % 8)] = v;
                }
            }
        }
    };
    for (int n0 = 0; n0 < De && p.n_ops_pre > 0; n0 += epart) {
        const int n = std::min(epart, De - n0);
        if (cursor + 2 > end) { *err = "internal: pre-op pack order does not match the op list"; return -1; }
        const QbOp& a = ops[cursor++];
        const QbOp& b = ops[cursor++];
        if (a.n != n || a.k_total != 2 * D || b.n != n || b.k_total != D) { *err = "internal: pre-op shapes do not match"; return -1; }
        put(a, n0, false);     // [Wx_hi | Wx_hi]
        put(b, n0, true);      // Wx_lo
    }
    if (cursor != end) { *err = "internal: pre-op pack order does not match the op list"; return -1; }
    return 0;
}

size_t prep_pack_bytes(int n_rows, int D) {
    const int n16 = (n_rows + 15) / 16 * 16;
    return (size_t)n16 * D * 2 * 2;
}

void prep_pack(const float* w, int n_rows, int D, uint16_t* out) {
    constexpr int kDc = 64, kNp = 256;          // QB_PREP_DC / QB_PREP_NP (qb_dev.h)
    const int n16 = (n_rows + 15) / 16 * 16;
    size_t off = 0;                             // in fp16 elements
    for (int d0 = 0; d0 < D; d0 += kDc) {
        const int dc = std::min(kDc, D - d0);
        for (int r0 = 0; r0 < n16; r0 += kNp) {
            const int np = std::min(kNp, n16 - r0);
            uint16_t* hi = out + off;
            uint16_t* lo = hi + (size_t)np * dc;
            for (int k = 0; k < dc; k++)
                for (int r = 0; r < np; r++) {
                    const float v = (r0 + r < n_rows) ? w[(size_t)(r0 + r) * D + d0 + k] : 0.f;
                    const uint16_t h = f32_to_f16(v);
                    const size_t at = ((size_t)(k / 8) * np + r) * 8 + (k % 8);
                    hi[at] = h;
                    lo[at] = f32_to_f16(v - f16_to_f32(h));
                }
            off += 2 * (size_t)np * dc;
        }
    }
}

size_t ivf_pack_bytes(int ivf_K, int D) {
    constexpr int kNp = 128;                    // QB_IVF_NP
    return (size_t)((ivf_K + kNp - 1) / kNp) * ((size_t)kNp * D * 4 + kNp * 4);
}

void ivf_pack(const float* cent, int ivf_K, int D, uint8_t* out) {
    constexpr int kNp = 128;
    const size_t part_bytes = (size_t)kNp * D * 4 + kNp * 4;
    for (int k0 = 0, part = 0; k0 < ivf_K; k0 += kNp, part++) {
        uint16_t* hi = reinterpret_cast<uint16_t*>(out + (size_t)part * part_bytes);
        uint16_t* lo = hi + (size_t)kNp * D;
        float* cn = reinterpret_cast<float*>(out + (size_t)part * part_bytes + (size_t)kNp * D * 4);
        for (int r = 0; r < kNp; r++) {
            const bool real = k0 + r < ivf_K;
            const float* c = cent + (size_t)(k0 + r) * D;
            float nrm = 0.f;                     // fp32 like the reference's (b ** 2).sum(-1)
            for (int d = 0; d < D; d++) {
                const float v = real ? c[d] : 0.f;
                nrm += v * v;
                const uint16_t h = f32_to_f16(v);
                const size_t at = ((size_t)(d / 8) * kNp + r) * 8 + (d % 8);
                hi[at] = h;
                lo[at] = f32_to_f16(v - f16_to_f32(h));
            }
            cn[r] = real ? nrm : 3.0e38f;
        }
    }
}

// T_m[k] = e0 + Wcat[:, :De] . e0 + bcat,  e0 = Pin . C_m[k]   (double accumulation, stored fp32, blocked [De/4][K][4]:
// consecutive codes are 16 B apart, so a warp whose lanes hold consecutive codes gathers 512 contiguous bytes)
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
        for (int e = 0; e < De; e++) t_blk[((size_t)(e / 4) * K + k) * 4 + (e % 4)] = (float)t[e];
        for (int d = 0; d < D; d++) cb_blk[((size_t)(d / 4) * K + k) * 4 + (d % 4)] = c[d];
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
        opt.hc = opts5[0]; opt.n_tiles = opts5[1] & 0xff; opt.pair = (opts5[1] >> 8) & 0xff; opt.mcast = (opts5[1] >> 16) & 0xff; opt.slot_bytes = opts5[2];
        opt.max_stage = opts5[3] & 0xff; opt.no_resident = (opts5[3] >> 8) & 1; opt.no_hsplit = (opts5[3] >> 9) & 1; opt.uop = (opts5[3] >> 10) & 1; opt.no_esplit = (opts5[3] >> 11) & 1; opt.max_slab_k = opts5[4];
    }
    QbStepPlan p;
    std::vector<QbOp> ops;
    std::string err;
    if (qb::make_step_plan(D, De, Dh, L, K, qinco1_mode, opt, &p, &ops, &err)) return -1;
    const int32_t v[] = {p.D, p.De, p.Dh, p.L, p.K, p.has_proj, p.skip, p.n_tiles, p.tmem_alloc_cols,
                         p.n_ops_block, p.n_ops_out, p.hc, p.n_hchunk, p.oc, p.n_ochunk, p.tmem_e_col, p.tmem_h_col,
                         p.tmem_tile_cols, p.smem_tres, p.smem_ring, p.slot_bytes, p.n_stage, p.smem_total,
                         (int32_t)p.block_w_bytes, (int32_t)p.w_blob_bytes, p.pair, p.h_split, p.n_ops_pre, p.ae_chunks,
                         p.e_split, p.mcast, qb::mlp_plan_view(p)};
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
        opt.hc = opts5[0]; opt.n_tiles = opts5[1] & 0xff; opt.pair = (opts5[1] >> 8) & 0xff; opt.mcast = (opts5[1] >> 16) & 0xff; opt.slot_bytes = opts5[2];
        opt.max_stage = opts5[3] & 0xff; opt.no_resident = (opts5[3] >> 8) & 1; opt.no_hsplit = (opts5[3] >> 9) & 1; opt.uop = (opts5[3] >> 10) & 1; opt.no_esplit = (opts5[3] >> 11) & 1; opt.max_slab_k = opts5[4];
    }
    QbStepPlan p;
    std::vector<QbOp> ops;
    std::string err;
    if (qb::make_step_plan(D, De, Dh, L, K, qinco1_mode, opt, &p, &ops, &err)) return -1;
    if (blob_halfs * 2 < p.w_blob_bytes) return -2;
    return qb::pack_step_weights(p, ops, up, down, out_proj, blob, &err);
}

int qb_plan_pack_pre(int D, int De, int Dh, int L, int K, int qinco1_mode, const int32_t* opts5, const float* wx, uint16_t* blob,
                     int64_t blob_halfs) {
    qb::PlanOptions opt;
    if (opts5) {
        opt.hc = opts5[0]; opt.n_tiles = opts5[1] & 0xff; opt.pair = (opts5[1] >> 8) & 0xff; opt.mcast = (opts5[1] >> 16) & 0xff; opt.slot_bytes = opts5[2];
        opt.max_stage = opts5[3] & 0xff; opt.no_resident = (opts5[3] >> 8) & 1; opt.no_hsplit = (opts5[3] >> 9) & 1; opt.uop = (opts5[3] >> 10) & 1; opt.no_esplit = (opts5[3] >> 11) & 1; opt.max_slab_k = opts5[4];
    }
    QbStepPlan p;
    std::vector<QbOp> ops;
    std::string err;
    if (qb::make_step_plan(D, De, Dh, L, K, qinco1_mode, opt, &p, &ops, &err)) return -1;
    if (blob_halfs * 2 < p.w_blob_bytes) return -2;
    return qb::pack_pre_weights(p, ops, wx, blob, &err);
}

int qb_plan_tables(int D, int De, int K, const float* codebook, const float* in_proj, const float* concat_w,
                   const float* concat_b, float* t_blk, float* cb_blk, float* wx_t) {
    qb::build_tables(D, De, K, codebook, in_proj, concat_w, concat_b, t_blk, cb_blk, wx_t);
    return 0;
}

}  // extern "C"
