// Dual-tile ("ping-pong") attention kernel for head dims <= 112 with fp16 P / V^T — the d = 40
// self-attention over 4096 tokens that dominates the 512^2 step, and d = 80 over 1024 tokens.
// Included by attention.cu.
//
// One CTA (one per SM, 320 threads) owns TWO 128-query tiles of one (sample, head) and streams the
// keys once for both, 128 keys per block:
//   warp 0        TMA producer: Q0, Q1, then K / V^T blocks through a 4-stage ring
//   warp 1        tcgen05.mma issuer. The score products are numbered k = 2 j + t (key block j, tile t)
//                 and rotate through NBUF TMEM score buffers, NBUF products ahead of the PV products:
//                 ... PV(k), S(k + NBUF), PV(k + 1), S(k + NBUF + 1) ...
//                 With NBUF = 3 (dv <= 64) the scores of a group's next block are already in TMEM when
//                 it finishes the current one. NBUF = 2 (dv = 80) is one score buffer per tile.
//   warps 2..5    softmax group 0 (tile 0), warps 6..9 softmax group 1 (tile 1): one thread per query row
// P never touches shared memory: the fp16 probabilities are written back over the first 64 columns of
// their own score buffer (tcgen05.st) and the PV product reads its A operand from tensor memory. With
// P in shared memory the kernel was bound by shared-memory bandwidth (per block and tile: 32 KB of P
// stores + 32 KB of P reads by the SS-mode MMA, next to 24 KB for Q K^T and 12 KB of V^T), measured as
// ~350 cycles per PV product instead of the 8 x 24-cycle tensor-core floor.
// TMEM: NBUF = 3: S/P [0,384) O0 [384,448) O1 [448,512); NBUF = 2: S/P [0,256) O0 [256,384) O1 [384,512).
// Softmax is single-pass in the steady state: exponentials use the running reference max m_ref
// (exp2 domain); only if a row's new max exceeds m_ref by more than 2^8 is the block redone with the
// new reference and O rescaled. The row sum costs no ALU work: row d of the V^T tile (zero-filled by
// TMA since d < dv) is overwritten with ones in shared memory, so column d of O accumulates sum_j p_ij
// on the tensor core, consistent with the fp16-rounded P the MMA actually sees.
#pragma once

// share of the softmax exponentials evaluated by exp2_poly3 instead of MUFU.EX2:
// 0 = none, 1 = 1/4, 2 = 3/8, 3 = 1/2
#ifndef ATT2_POLY
#define ATT2_POLY 0
#endif
// cycles by which softmax group 1 delays its first block, so that the two groups run out of phase: one
// group's serial part (barrier round trips, P store, PV / next-S issue) then hides under the other group's
// exponentials instead of both idling the MUFU pipe together. The issuer's strict k = 2j + t order sustains
// any lag between one MMA latency and one block. 0 = off.
#ifndef ATT2_STAGGER
#define ATT2_STAGGER 0
#endif
// 1: on the last key block the three-buffer instantiation waits for PV(nkv - 2) before it hands over P(nkv - 1), which
// makes the epilogue's parity wait exact (see the softmax loop). 0 = the kernel as measured in round 2 (A/B only: that
// kernel can read O two PV products early when a warp runs a key block ahead at the very end).
#ifndef ATT2_FINAL_GUARD
#define ATT2_FINAL_GUARD 1
#endif

namespace pp {

static constexpr int ATT2_THREADS = 320;

// dch = 64-channel chunks of the head dim (1: d <= 64, 2: d <= 112), stages = K / V^T ring depth
__host__ __device__ constexpr uint32_t att2_smem_bytes(uint32_t dv, uint32_t dch, uint32_t stages) {
    return 2 * dch * ATT_CHUNK_BYTES                             // Q0, Q1
           + stages * (dch * ATT_CHUNK_BYTES + dv * 256u)        // K + V^T ring
           + 256 + 1024;
}

// NBUF score buffers (3 needs dv <= 64), DCH head-dim chunks, S ring stages — all compile-time: j % S,
// j / S and k % NBUF sit on the MMA issuer's critical path.
// Instantiations: <3,1,4> d <= 48 (SD-1.5 d = 40), <2,1,4> d = 56/64, <2,2,2> d = 72..112 (SD-1.5 d = 80).
template <int NBUF, int DCH, int S>
__global__ void __launch_bounds__(ATT2_THREADS, 1) attn2_kernel(const __grid_constant__ AttnKParams p) {
    constexpr int MAXS = 4;
    static_assert(S <= MAXS, "barrier slots");
    constexpr uint32_t O_BASE = NBUF * 128, O_STRIDE = NBUF == 3 ? 64 : 128;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    const uint32_t v_chunk_bytes = (uint32_t)p.dv * 128u;
    const uint32_t v_stage_bytes = 2u * v_chunk_bytes;
    const uint32_t sQ = base;                                    // [2][DCH][16 KB]
    const uint32_t sK0 = sQ + 2 * DCH * ATT_CHUNK_BYTES;         // [S][DCH][16 KB]
    const uint32_t sV0 = sK0 + S * DCH * ATT_CHUNK_BYTES;        // [S][v_stage_bytes]
    const uint32_t bars = sV0 + S * v_stage_bytes;
    const uint32_t bar_q = bars;
    auto bar_kv_full = [&](int s) { return bars + 8u * (1 + s); };
    auto bar_kv_empty = [&](int s) { return bars + 8u * (1 + MAXS + s); };
    auto bar_s_full = [&](int i) { return bars + 8u * (1 + 2 * MAXS + i); };
    // P hand-over barriers: one per (tile, key-block parity). With three score buffers a softmax warp can be a whole key
    // block ahead of a slower warp of its group (never two: S of block j + 2 needs PV(j + 1), i.e. every warp's P(j + 1));
    // on a single barrier its arrival for block j + 1 could complete the 4-count phase of block j and release PV(j) over
    // raw scores (NaN rows at 16384 tokens, r02). Alternating barriers keep the two blocks' arrivals apart for free.
    auto bar_p_full = [&](int t, int j) { return bars + 8u * (7 + 2 * MAXS + 2 * t + (j & 1)); };
    auto bar_pv_done = [&](int t) { return bars + 8u * (11 + 2 * MAXS + t); };
    const uint32_t tmem_slot = bars + 8u * (13 + 2 * MAXS);
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - raw));

    const int warp = threadIdx.x >> 5;
    const int q0 = blockIdx.x * 2 * ATT_BM;
    const int head = blockIdx.y;
    const int b = blockIdx.z;
    const int nkv = (p.nk + ATT_BN - 1) / ATT_BN;

    if (warp == 0 && elect_one()) {
        prefetch_tmap(&p.tmQ);
        prefetch_tmap(&p.tmK);
        prefetch_tmap(&p.tmV);
        mbar_init(bar_q, 1);
        for (int s = 0; s < S; ++s) {
            mbar_init(bar_kv_full(s), 1);
            mbar_init(bar_kv_empty(s), 1);
        }
        for (int i = 0; i < NBUF; ++i) {
            mbar_init(bar_s_full(i), 1);
        }
        for (int t = 0; t < 2; ++t) {
            mbar_init(bar_p_full(t, 0), 4);
            mbar_init(bar_p_full(t, 1), 4);
            mbar_init(bar_pv_done(t), 1);
        }
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    pdl_wait();  // Q / K / V^T come from the preceding kernels
    pdl_launch_dependents();

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one()) {
            mbar_arrive_expect_tx(bar_q, 2 * DCH * ATT_CHUNK_BYTES);
#pragma unroll
            for (int tc = 0; tc < 2 * DCH; ++tc)  // tile tc / DCH, channel chunk tc % DCH (zero-filled past d)
                tma_load_4d(sQ + tc * ATT_CHUNK_BYTES, &p.tmQ, bar_q, (tc % DCH) * 64, head,
                            q0 + (tc / DCH) * ATT_BM, b);
            for (int j = 0; j < nkv; ++j) {
                const int s = j % S;
                const uint32_t ph = (j / S) & 1;
                mbar_wait(bar_kv_empty(s), ph ^ 1u);
                mbar_arrive_expect_tx(bar_kv_full(s), DCH * ATT_CHUNK_BYTES + v_stage_bytes);
#pragma unroll
                for (int c = 0; c < DCH; ++c)
                    tma_load_4d(sK0 + (s * DCH + c) * ATT_CHUNK_BYTES, &p.tmK, bar_kv_full(s), c * 64, head,
                                j * ATT_BN, b);
                const uint32_t dV = sV0 + s * v_stage_bytes;
                tma_load_3d(dV, &p.tmV, bar_kv_full(s), j * ATT_BN, 0, b * p.heads + head);
                tma_load_3d(dV + v_chunk_bytes, &p.tmV, bar_kv_full(s), j * ATT_BN + 64, 0, b * p.heads + head);
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (elect_one()) {
            const uint32_t idesc_s = umma_idesc_bf16(ATT_BM, ATT_BN);
            const uint32_t idesc_o = umma_idesc_f16(ATT_BM, (uint32_t)p.dv);  // P and V^T are fp16 here
            // score product k = 2 j + t into buffer k % NBUF (free: PV(k - NBUF) has retired, see below)
            auto issue_s = [&](int k) {
                const int j = k >> 1, t = k & 1, buf = k % NBUF;
                if (t == 0) {
                    mbar_wait(bar_kv_full(j % S), (j / S) & 1);  // first use of key block j
                    tc_fence_after();
                }
                const uint32_t qb = sQ + t * DCH * ATT_CHUNK_BYTES;
                const uint32_t kb = sK0 + (j % S) * DCH * ATT_CHUNK_BYTES;
                for (int ks = 0; ks < p.k_steps; ++ks) {
                    const int c = ks >> 2, kk = (ks & 3) * 16;  // 64-channel chunk, offset inside it
                    umma_bf16_ss(tmem_base + buf * 128,
                                 umma_desc_advance_k(umma_desc_kmajor_sw128(qb + c * ATT_CHUNK_BYTES), kk),
                                 umma_desc_advance_k(umma_desc_kmajor_sw128(kb + c * ATT_CHUNK_BYTES), kk), idesc_s,
                                 ks != 0);
                }
                umma_commit(bar_s_full(buf));
            };
            // O(t) += P(k) V(j): P is read straight from tensor memory (the fp16 probabilities overwrite
            // the first 64 columns of their own score buffer), V^T from shared memory
            auto issue_pv = [&](int k) {
                const int j = k >> 1, t = k & 1, buf = k % NBUF;
                const uint32_t vb = sV0 + (j % S) * v_stage_bytes;
                const int pv_steps = (min(ATT_BN, p.nk - j * ATT_BN) + 15) >> 4;  // 16 keys per MMA; 8 unless ragged
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {
                    if (ks >= pv_steps) break;
                    const int c = ks >> 2, kk = (ks & 3) * 16;
                    umma_f16_ts(tmem_base + O_BASE + t * O_STRIDE, tmem_base + buf * 128 + ks * 8,
                                umma_desc_advance_k(umma_desc_kmajor_sw128(vb + c * v_chunk_bytes), kk), idesc_o,
                                (j | ks) != 0);
                }
                umma_commit(bar_pv_done(t));
            };
            const int nprod = 2 * nkv;
            mbar_wait(bar_q, 0);
            for (int k = 0; k < NBUF && k < nprod; ++k) issue_s(k);
            for (int k = 0; k < nprod; ++k) {
                const int j = k >> 1, t = k & 1;
                mbar_wait(bar_p_full(t, j), (j >> 1) & 1);
                tc_fence_after();
                issue_pv(k);
                if (t == 1) umma_commit(bar_kv_empty(j % S));  // both tiles' PV(j) precede this commit
                if (k + NBUF < nprod) {
                    // the next scores for this buffer overwrite P(k): wait until PV(k) has read it
                    mbar_wait(bar_pv_done(t), j & 1);
                    tc_fence_after();
                    issue_s(k + NBUF);
                }
            }
        }
        __syncwarp();
    } else {
        // ===================== softmax groups =====================
        const int t = (warp - 2) >> 2;          // tile / group index
        const int quarter = warp & 3;
        const int r = quarter * 32 + (int)lane_id();
        const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
        const uint32_t tO = tmem_base + O_BASE + t * O_STRIDE + lane_addr;
        float c;  // pinned in a register (otherwise re-fetched from the constant bank per element)
        asm volatile("mov.f32 %0, %1;" : "=f"(c) : "f"(p.scale_log2));
        float m_ref = 0.f;
#if ATT2_STAGGER > 0
        if (t == 1 && nkv > 4) {
            const long long t_start = clock64();
            while (clock64() - t_start < ATT2_STAGGER) {}
        }
#endif
        for (int j = 0; j < nkv; ++j) {
            const int k = 2 * j + t, buf = k % NBUF;
            const uint32_t tS = tmem_base + buf * 128 + lane_addr;
            mbar_wait(bar_s_full(buf), (k / NBUF) & 1);
            tc_fence_after();
            const int nvalid = min(ATT_BN, p.nk - j * ATT_BN);
            if (j == 0) {
                // the first block has no reference yet: one extra pass for the row max
                float mx = -INFINITY;
#pragma unroll 1
                for (int ch = 0; ch * 32 < nvalid; ++ch) {
                    uint32_t sv[32];
                    tmem_ld32(tS + ch * 32, sv);
                    tmem_wait_ld();
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (ch * 32 + i < nvalid) mx = fmaxf(mx, __uint_as_float(sv[i]));
                }
                m_ref = mx * c;
            }
            uint32_t pk[64];
            float alpha = 1.f;
            bool redo;
            do {
                // four independent running maxima break the 128-long fmax dependency chain; the
                // TMEM load of chunk ch+1 is in flight while chunk ch is exponentiated
                float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
                uint32_t sa[32], sb[32];
                tmem_ld32(tS, sa);
                tmem_wait_ld();
                auto process = [&](const uint32_t (&sv)[32], int ch) {
                    if (nvalid == ATT_BN) {
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            const float s0 = __uint_as_float(sv[i]), s1 = __uint_as_float(sv[i + 1]);
                            const float s2 = __uint_as_float(sv[i + 2]), s3 = __uint_as_float(sv[i + 3]);
                            mx0 = fmaxf(mx0, s0); mx1 = fmaxf(mx1, s1); mx2 = fmaxf(mx2, s2); mx3 = fmaxf(mx3, s3);
#if ATT2_POLY == 0
                            pk[ch * 16 + i / 2] = ex2_f16x2(pack_f16x2(fmaf(s0, c, -m_ref), fmaf(s1, c, -m_ref)));
                            pk[ch * 16 + i / 2 + 1] = ex2_f16x2(pack_f16x2(fmaf(s2, c, -m_ref), fmaf(s3, c, -m_ref)));
#else
                            // a fixed share of the keys takes the polynomial path: the block is bound by
                            // the MUFU pipe (2 warps x 128 keys x 8 cycles per scheduler), while the FMA
                            // and integer pipes idle
                            const bool poly1 = ATT2_POLY == 3 || (ATT2_POLY == 2 && (i & 4));
                            const float e0 = fast_exp2(fmaf(s0, c, -m_ref));
                            const float e1 = poly1 ? exp2_poly3(fmaf(s1, c, -m_ref)) : fast_exp2(fmaf(s1, c, -m_ref));
                            const float e2 = fast_exp2(fmaf(s2, c, -m_ref));
                            const float e3 = exp2_poly3(fmaf(s3, c, -m_ref));
                            pk[ch * 16 + i / 2] = pack_f16x2(e0, e1);
                            pk[ch * 16 + i / 2 + 1] = pack_f16x2(e2, e3);
#endif
                        }
                    } else if (ch * 32 >= nvalid) {
                        // chunk entirely past the last key (the 77-key cross-attention ends in chunk 2)
#pragma unroll
                        for (int i = 0; i < 16; ++i) pk[ch * 16 + i] = 0u;
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; i += 2) {
                            const float s0 = __uint_as_float(sv[i]), s1 = __uint_as_float(sv[i + 1]);
                            const bool v0 = ch * 32 + i < nvalid, v1 = ch * 32 + i + 1 < nvalid;
                            if (v0) mx0 = fmaxf(mx0, s0);
                            if (v1) mx1 = fmaxf(mx1, s1);
                            const float p0 = v0 ? fast_exp2(fmaf(s0, c, -m_ref)) : 0.f;
                            const float p1 = v1 ? fast_exp2(fmaf(s1, c, -m_ref)) : 0.f;
                            pk[ch * 16 + i / 2] = pack_f16x2(p0, p1);
                        }
                    }
                };
                tmem_ld32(tS + 32, sb);
                process(sa, 0);
                tmem_wait_ld();
                tmem_ld32(tS + 64, sa);
                process(sb, 1);
                tmem_wait_ld();
                tmem_ld32(tS + 96, sb);
                process(sa, 2);
                tmem_wait_ld();
                process(sb, 3);
                const float mnew = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * c;
                const bool need = mnew > m_ref + 8.0f;
                redo = __any_sync(0xffffffffu, need);
                if (need) {
                    alpha *= fast_exp2(m_ref - mnew);
                    m_ref = mnew;
                }
            } while (redo);
            // The LAST block of the three-buffer instantiation: a warp can finish block nkv - 1 while a slower warp of its
            // group still holds back PV(nkv - 2); bar_pv_done would then be two phases short of the parity the epilogue
            // waits for, and that wait would pass at once (a parity wait cannot tell "two behind" from "done") — O read
            // without the last two key blocks. Here, between the score wait and this warp's own arrival, the barrier is
            // at most one phase short, so the wait is exact; the block's exponentials have just given PV(nkv - 2)
            // ~3000 cycles to retire, so it rarely spins. It is the wait of the rescale path below, which
            // test_attention_large_logits takes on its last block. With two score buffers S(k + 2) needs PV(k), so no
            // warp gets that far ahead. (tests/test_attention_protocol_model.py explores every interleaving.)
            if constexpr (NBUF == 3 && ATT2_FINAL_GUARD != 0) {
                if (j == nkv - 1 && j > 0) {
                    mbar_wait(bar_pv_done(t), (j - 1) & 1);
                    tc_fence_after();
                }
            }
            if (__any_sync(0xffffffffu, alpha != 1.f)) {
                if (j > 0) {
                    mbar_wait(bar_pv_done(t), (j - 1) & 1);  // O holds blocks < j
                    tc_fence_after();
                }
                for (int col = 0; col < p.dv; col += 16) {
                    uint32_t ov[16];
                    tmem_ld16(tO + col, ov);
                    tmem_wait_ld();
#pragma unroll
                    for (int i = 0; i < 16; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * alpha);
                    tmem_st16(tO + col, ov);
                }
            }
            // P(k) (fp16 pairs, one 32-bit column per key pair) replaces the first 64 columns of S(k);
            // every thread owns its TMEM lane, and its own score reads have completed (wait::ld above)
#pragma unroll
            for (int q = 0; q < 4; ++q) tmem_st16_from(tS + q * 16, &pk[q * 16]);
            tmem_wait_st();
            if (t == 0 && warp == 2 && lane_id() < 16) {
                // row d of the V^T tile := 1.0 (fp16 0x3C00) so O[:, d] accumulates the row sums
                const int ch = lane_id() >> 3, piece = lane_id() & 7;
                uint8_t* vrow = smem_raw + (sV0 + (j % S) * v_stage_bytes + ch * v_chunk_bytes - raw) + p.d * 128;
                *reinterpret_cast<uint4*>(vrow + piece * 16) =
                    make_uint4(0x3C003C00u, 0x3C003C00u, 0x3C003C00u, 0x3C003C00u);
                fence_proxy_async_smem();  // the tensor core reads this row through the async proxy
            }
            tc_fence_before();
            __syncwarp();
            if (lane_id() == 0) mbar_arrive(bar_p_full(t, j));
        }
        // epilogue: O[:, :d] / O[:, d]
        mbar_wait(bar_pv_done(t), (nkv - 1) & 1);
        tc_fence_after();
        const int qi = q0 + t * ATT_BM + r;
        __nv_bfloat16* orow = p.out + ((int64_t)b * p.nq + qi) * p.o_ld + head * p.d;
        float inv_l = 1.f;
        {
            uint32_t lv[16];
            tmem_ld16(tO + (p.d & ~15), lv);
            tmem_wait_ld();
            float l = 1.f;
#pragma unroll
            for (int i = 0; i < 16; ++i)
                if (i == (p.d & 15)) l = __uint_as_float(lv[i]);
            inv_l = 1.0f / l;
        }
        for (int col = 0; col < p.d; col += 16) {
            uint32_t ov[16];
            tmem_ld16(tO + col, ov);
            tmem_wait_ld();
            if (qi < p.nq) {
#pragma unroll
                for (int g = 0; g < 16; g += 8) {
                    if (col + g < p.d) {
                        uint4 o;
                        o.x = pack_bf16x2(__uint_as_float(ov[g + 0]) * inv_l, __uint_as_float(ov[g + 1]) * inv_l);
                        o.y = pack_bf16x2(__uint_as_float(ov[g + 2]) * inv_l, __uint_as_float(ov[g + 3]) * inv_l);
                        o.z = pack_bf16x2(__uint_as_float(ov[g + 4]) * inv_l, __uint_as_float(ov[g + 5]) * inv_l);
                        o.w = pack_bf16x2(__uint_as_float(ov[g + 6]) * inv_l, __uint_as_float(ov[g + 7]) * inv_l);
                        *reinterpret_cast<uint4*>(orow + col + g) = o;
                    }
                }
            }
        }
        tc_fence_before();
    }

    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace pp
