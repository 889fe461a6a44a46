// Fused scaled-dot-product attention on tcgen05 / TMEM / TMA (sm_100a), no mask, non-causal.
//
// Reference: diffusers Attention + AttnProcessor2_0 -> F.scaled_dot_product_attention(q, k, v)
// with scale 1/sqrt(d) (processors referenced at powerpaint/models/unet_2d_condition.py:24-31;
// instantiated through Transformer2DModel at unet_2d_blocks.py:807,1289,2514). Self-attention
// over N = h*w latent tokens (4096/1024/256/64 at 512^2) and cross-attention over 77 text
// tokens; 8 heads of d = 40 / 80 / 160.
//
// One CTA = one (128-query tile, head, sample). Flash-style streaming over 128-key blocks:
//   warp 0      TMA producer: Q once, then K / V^T blocks through a ring of kv stages
//   warp 1      tcgen05.mma issuer: S = Q K^T into TMEM, then O += P V into TMEM
//   warps 2..5  softmax: one thread per query row reads its S row with tcgen05.ld (no
//               shuffles), online max / sum in fp32 (exp2 domain), writes P as bf16 into a
//               swizzled K-major smem tile for the second MMA, rescales O in TMEM only when
//               the running max grew by more than 2^8 (lazy rescale), and finally normalises
//               and stores O.
// Head dims that are not multiples of 64/16 need no padding in HBM: the Q/K tensor maps are
// (d, heads, tokens, batch) so TMA zero-fills channels >= d of each head, and V is consumed
// transposed ([batch, heads*d, keys], keys contiguous; produced by the projection GEMM's
// transposed epilogue) so both MMAs use the same K-major SWIZZLE_128B operand layout.
#include <math.h>

#include <algorithm>

#include <cuda_fp16.h>

#include "common.cuh"
#include "ops.h"

namespace pp {

static constexpr int ATT_BM = 128;   // queries per CTA
static constexpr int ATT_BN = 128;   // keys per block
static constexpr int ATT_THREADS = 192;
static constexpr uint32_t ATT_CHUNK_BYTES = 128 * 128;  // [128 rows x 64 bf16]
static constexpr uint32_t TMEM_O_COL = 128;

__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int DCH>
struct AttSmem {
    static constexpr int KV_STAGES = DCH == 3 ? 1 : 2;
    static constexpr uint32_t Q_BYTES = DCH * ATT_CHUNK_BYTES;
    static constexpr uint32_t K_BYTES = DCH * ATT_CHUNK_BYTES;
    // V^T tile: two 64-key chunks of [dv rows x 128 B] (dv = ceil16(d), runtime), so a stage
    // is dv*256 bytes; the attribute maximum assumes dv = 64*DCH.
    static constexpr uint32_t P_BYTES = 2 * ATT_CHUNK_BYTES;
    static constexpr uint32_t total(uint32_t dv) {
        return Q_BYTES + KV_STAGES * (K_BYTES + dv * 256u) + P_BYTES + 256 + 1024;
    }
    static constexpr uint32_t TOTAL_MAX = total(64u * DCH);
    static constexpr uint32_t TMEM_COLS = DCH == 1 ? 256 : 512;
};

template <int DCH>
__global__ void __launch_bounds__(ATT_THREADS) attn_fwd_kernel(const __grid_constant__ AttnKParams p) {
    using L = AttSmem<DCH>;
    constexpr int S = L::KV_STAGES;

    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    const uint32_t sQ = base;
    const uint32_t sK0 = sQ + L::Q_BYTES;
    const uint32_t v_chunk_bytes = (uint32_t)p.dv * 128u;  // bytes one V^T chunk box delivers
    const uint32_t v_stage_bytes = 2u * v_chunk_bytes;
    const uint32_t sV0 = sK0 + S * L::K_BYTES;
    const uint32_t sP = sV0 + S * v_stage_bytes;
    const uint32_t bars = sP + L::P_BYTES;
    const uint32_t bar_q = bars;
    auto bar_kv_full = [&](int s) { return bars + 8u * (1 + s); };
    auto bar_kv_empty = [&](int s) { return bars + 8u * (1 + S + s); };
    const uint32_t bar_s_full = bars + 8u * (1 + 2 * S);
    const uint32_t bar_p_full = bars + 8u * (2 + 2 * S);
    const uint32_t bar_pv_done = bars + 8u * (3 + 2 * S);
    const uint32_t tmem_slot = bars + 8u * (4 + 2 * S);
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - raw));
    uint8_t* sP_ptr = smem_raw + (sP - raw);

    const int warp = threadIdx.x >> 5;
    const int q0 = blockIdx.x * ATT_BM;
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
        mbar_init(bar_s_full, 1);
        mbar_init(bar_p_full, 128);
        mbar_init(bar_pv_done, 1);
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, L::TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    pdl_wait();  // Q / K / V^T come from the preceding kernels
    pdl_launch_dependents();
    const uint32_t tS = tmem_base;
    const uint32_t tO = tmem_base + TMEM_O_COL;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one()) {
            mbar_arrive_expect_tx(bar_q, L::Q_BYTES);
#pragma unroll
            for (int dc = 0; dc < DCH; ++dc)
                tma_load_4d(sQ + dc * ATT_CHUNK_BYTES, &p.tmQ, bar_q, dc * 64, head, q0, b);
            for (int j = 0; j < nkv; ++j) {
                const int s = j % S;
                const uint32_t ph = (j / S) & 1;
                mbar_wait(bar_kv_empty(s), ph ^ 1u);
                mbar_arrive_expect_tx(bar_kv_full(s), L::K_BYTES + 2u * v_chunk_bytes);
                const uint32_t dK = sK0 + s * L::K_BYTES;
                const uint32_t dV = sV0 + s * v_stage_bytes;
#pragma unroll
                for (int dc = 0; dc < DCH; ++dc)
                    tma_load_4d(dK + dc * ATT_CHUNK_BYTES, &p.tmK, bar_kv_full(s), dc * 64, head, j * ATT_BN, b);
                tma_load_3d(dV, &p.tmV, bar_kv_full(s), j * ATT_BN, 0, b * p.heads + head);
                tma_load_3d(dV + v_chunk_bytes, &p.tmV, bar_kv_full(s), j * ATT_BN + 64, 0, b * p.heads + head);
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (elect_one()) {
            const uint32_t idesc_s = umma_idesc_bf16(ATT_BM, ATT_BN);
            const uint32_t idesc_o = p.vt_fp16 ? umma_idesc_f16(ATT_BM, (uint32_t)p.dv) : umma_idesc_bf16(ATT_BM, (uint32_t)p.dv);
            auto issue_s = [&](int j) {
                const int s = j % S;
                mbar_wait(bar_kv_full(s), (j / S) & 1);
                tc_fence_after();
                const uint32_t kb = sK0 + s * L::K_BYTES;
                for (int ks = 0; ks < p.k_steps; ++ks) {
                    const int dc = ks >> 2, kk = (ks & 3) * 16;
                    const uint64_t da = umma_desc_advance_k(umma_desc_kmajor_sw128(sQ + dc * ATT_CHUNK_BYTES), kk);
                    const uint64_t db = umma_desc_advance_k(umma_desc_kmajor_sw128(kb + dc * ATT_CHUNK_BYTES), kk);
                    umma_bf16_ss(tS, da, db, idesc_s, ks != 0);
                }
                umma_commit(bar_s_full);
            };
            auto issue_pv = [&](int j) {
                const int s = j % S;
                const uint32_t vb = sV0 + s * v_stage_bytes;
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {
                    const int c = ks >> 2, kk = (ks & 3) * 16;
                    const uint64_t da = umma_desc_advance_k(umma_desc_kmajor_sw128(sP + c * ATT_CHUNK_BYTES), kk);
                    const uint64_t db = umma_desc_advance_k(umma_desc_kmajor_sw128(vb + c * v_chunk_bytes), kk);
                    umma_bf16_ss(tO, da, db, idesc_o, (j | ks) != 0);
                }
                umma_commit(bar_kv_empty(s));
                umma_commit(bar_pv_done);
            };
            mbar_wait(bar_q, 0);
            issue_s(0);
            for (int j = 0; j < nkv; ++j) {
                mbar_wait(bar_p_full, j & 1);
                tc_fence_after();
                if (S >= 2) {
                    if (j + 1 < nkv) issue_s(j + 1);  // next S first: softmax(j+1) overlaps PV(j)
                    issue_pv(j);
                } else {
                    issue_pv(j);
                    if (j + 1 < nkv) issue_s(j + 1);
                }
            }
        }
        __syncwarp();
    } else {
        // ===================== softmax / correction / epilogue =====================
        const int quarter = warp & 3;
        const int r = quarter * 32 + (int)lane_id();
        const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
        float m_ref = -INFINITY;
        float l = 0.f;
        for (int j = 0; j < nkv; ++j) {
            mbar_wait(bar_s_full, j & 1);
            tc_fence_after();
            const int nvalid = min(ATT_BN, p.nk - j * ATT_BN);
            // pass 1: row max
            float mx = -INFINITY;
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                uint32_t sv[32];
                tmem_ld32(tS + lane_addr + c * 32, sv);
                tmem_wait_ld();
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if (c * 32 + i < nvalid) mx = fmaxf(mx, __uint_as_float(sv[i]));
            }
            mx *= p.scale_log2;
            float alpha = 1.f;
            bool need = false;
            if (j == 0) {
                m_ref = mx;
            } else if (mx > m_ref + 8.0f) {
                alpha = fast_exp2(m_ref - mx);
                m_ref = mx;
                need = true;
            }
            // pass 2: p = exp2(s * scale_log2 - m_ref), bf16, row sum of the rounded values
            uint32_t pk[64];
            float lsum = 0.f;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t sv[32];
                tmem_ld32(tS + lane_addr + c * 32, sv);
                tmem_wait_ld();
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    float p0 = (c * 32 + i < nvalid) ? fast_exp2(fmaf(__uint_as_float(sv[i]), p.scale_log2, -m_ref)) : 0.f;
                    float p1 = (c * 32 + i + 1 < nvalid) ? fast_exp2(fmaf(__uint_as_float(sv[i + 1]), p.scale_log2, -m_ref)) : 0.f;
                    uint32_t q;
                    if (p.vt_fp16) {
                        q = pack_f16x2(p0, p1);
                        const __half2 h = *reinterpret_cast<const __half2*>(&q);
                        lsum += __low2float(h) + __high2float(h);
                    } else {
                        q = pack_bf16x2(p0, p1);
                        lsum += bf16_lo(q) + bf16_hi(q);
                    }
                    pk[c * 16 + i / 2] = q;
                }
            }
            l = l * alpha + lsum;
            if (j > 0) {
                mbar_wait(bar_pv_done, (j - 1) & 1);  // P buffer free, O holds blocks < j
                tc_fence_after();
            }
            if (__any_sync(0xffffffffu, need)) {
                for (int c = 0; c < p.dv; c += 16) {
                    uint32_t ov[16];
                    tmem_ld16(tO + lane_addr + c, ov);
                    tmem_wait_ld();
#pragma unroll
                    for (int i = 0; i < 16; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * alpha);
                    tmem_st16(tO + lane_addr + c, ov);
                }
                tmem_wait_st();
            }
            // P -> smem, K-major SWIZZLE_128B: row r, 16-byte piece i of chunk c at ((i ^ (r & 7)) << 4)
            uint8_t* prow = sP_ptr + r * 128;
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const int c = q >> 3, i = q & 7;
                uint4 v = make_uint4(pk[q * 4], pk[q * 4 + 1], pk[q * 4 + 2], pk[q * 4 + 3]);
                *reinterpret_cast<uint4*>(prow + c * ATT_CHUNK_BYTES + ((i ^ (r & 7)) << 4)) = v;
            }
            fence_proxy_async_smem();
            tc_fence_before();
            mbar_arrive(bar_p_full);
        }
        // epilogue: O / l -> bf16
        mbar_wait(bar_pv_done, (nkv - 1) & 1);
        tc_fence_after();
        const int qi = q0 + r;
        const float inv_l = 1.0f / l;
        __nv_bfloat16* orow = p.out + ((int64_t)b * p.nq + qi) * p.o_ld + head * p.d;
        for (int c = 0; c < p.dv; c += 16) {
            uint32_t ov[16];
            tmem_ld16(tO + lane_addr + c, ov);
            tmem_wait_ld();
            if (qi < p.nq) {
#pragma unroll
                for (int g = 0; g < 16; g += 8) {
                    if (c + g < p.d) {
                        uint4 o;
                        o.x = pack_bf16x2(__uint_as_float(ov[g + 0]) * inv_l, __uint_as_float(ov[g + 1]) * inv_l);
                        o.y = pack_bf16x2(__uint_as_float(ov[g + 2]) * inv_l, __uint_as_float(ov[g + 3]) * inv_l);
                        o.z = pack_bf16x2(__uint_as_float(ov[g + 4]) * inv_l, __uint_as_float(ov[g + 5]) * inv_l);
                        o.w = pack_bf16x2(__uint_as_float(ov[g + 6]) * inv_l, __uint_as_float(ov[g + 7]) * inv_l);
                        *reinterpret_cast<uint4*>(orow + c + g) = o;
                    }
                }
            }
        }
        tc_fence_before();
    }

    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, L::TMEM_COLS);
    }
}

}  // namespace pp

#include "attention2.cuh"

namespace pp {

// ---------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------
template <int NBUF, int DCH, int S>
static int attn2_ensure_attr() {
    static bool done[PP_MAX_DEVICES] = {};  // the attribute is per device
    int dev = 0;
    PP_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= PP_MAX_DEVICES || !done[dev]) {
        // largest V^T tile the variant can meet: dv <= 64 with three score buffers, else <= 128
        PP_CUDA_CHECK(cudaFuncSetAttribute(attn2_kernel<NBUF, DCH, S>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)att2_smem_bytes(NBUF == 3 ? 64 : 128, DCH, S)));
        if (dev >= 0 && dev < PP_MAX_DEVICES) done[dev] = true;
    }
    return PP_OK;
}

template <int DCH>
static int attn_ensure_attr() {
    static bool done[PP_MAX_DEVICES] = {};  // the attribute is per device
    int dev = 0;
    PP_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= PP_MAX_DEVICES || !done[dev]) {
        PP_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd_kernel<DCH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)AttSmem<DCH>::TOTAL_MAX));
        if (dev >= 0 && dev < PP_MAX_DEVICES) done[dev] = true;
    }
    return PP_OK;
}

int attn_prepare(const pp_attn_desc& d, AttnLaunch* out) {
    AttnLaunch l;
    memset(&l, 0, sizeof(l));
    AttnKParams& p = l.p;
    PP_REQUIRE(d.q && d.k && d.vt && d.out, "attention: null pointer");
    PP_REQUIRE(d.batch > 0 && d.heads > 0 && d.nq > 0 && d.nk > 0, "attention: empty problem");
    PP_REQUIRE(d.d >= 8 && d.d % 8 == 0 && d.d <= 192, "attention: head dim %d unsupported (multiple of 8, <= 192)", d.d);
    PP_REQUIRE(d.vt_ld >= d.nk && d.vt_ld % 8 == 0, "attention: vt_ld=%lld invalid", (long long)d.vt_ld);
    PP_REQUIRE(d.q_ld % 8 == 0 && d.k_ld % 8 == 0 && d.o_ld % 8 == 0, "attention: row pitches must be multiples of 8");
    PP_REQUIRE(d.q_batch_stride % 8 == 0 && d.k_batch_stride % 8 == 0, "attention: batch strides must be multiples of 8");
    PP_REQUIRE(d.o_ld >= (int64_t)d.heads * d.d, "attention: o_ld too small");
    PP_REQUIRE((reinterpret_cast<uintptr_t>(d.out) & 15) == 0, "attention: out not 16-byte aligned");
    p.batch = d.batch; p.heads = d.heads; p.d = d.d; p.nq = d.nq; p.nk = d.nk;
    p.d_chunks = (d.d + 63) / 64;
    p.k_steps = (d.d + 15) / 16;
    // dual-tile kernel for d <= 64 when there are at least two query tiles; it keeps the row sums
    // in column d of O, so its V^T tile has one extra (ones) row
    const bool dual = d.d <= 112 && d.nq > ATT_BM && d.vt_fp16;  // dv = ceil16(d + 1) must fit 128 TMEM columns
    p.dv = dual ? ((d.d + 1 + 15) / 16) * 16 : ((d.d + 15) / 16) * 16;
    p.scale_log2 = d.scale * 1.4426950408889634f;
    p.vt_fp16 = d.vt_fp16;
    p.out = reinterpret_cast<__nv_bfloat16*>(d.out);
    p.o_ld = d.o_ld;
    {
        uint64_t dims[4] = {(uint64_t)d.d, (uint64_t)d.heads, (uint64_t)d.nq, (uint64_t)d.batch};
        uint64_t str[3] = {(uint64_t)d.d * 2, (uint64_t)d.q_ld * 2, (uint64_t)d.q_batch_stride * 2};
        uint32_t box[4] = {64, 1, ATT_BM, 1};
        int rc = make_tmap_bf16(&p.tmQ, d.q, 4, dims, str, box, true);
        if (rc) return rc;
    }
    {
        uint64_t dims[4] = {(uint64_t)d.d, (uint64_t)d.heads, (uint64_t)d.nk, (uint64_t)d.batch};
        uint64_t str[3] = {(uint64_t)d.d * 2, (uint64_t)d.k_ld * 2, (uint64_t)d.k_batch_stride * 2};
        uint32_t box[4] = {64, 1, ATT_BN, 1};
        int rc = make_tmap_bf16(&p.tmK, d.k, 4, dims, str, box, true);
        if (rc) return rc;
    }
    {
        uint64_t dims[3] = {(uint64_t)d.nk, (uint64_t)d.d, (uint64_t)d.batch * d.heads};
        uint64_t str[2] = {(uint64_t)d.vt_ld * 2, (uint64_t)d.vt_ld * 2 * d.d};
        uint32_t box[3] = {64, (uint32_t)p.dv, 1};
        int rc = make_tmap_bf16(&p.tmV, d.vt, 3, dims, str, box, true);
        if (rc) return rc;
    }
    if (dual) {
        // three rotating score buffers need dv <= 64 TMEM columns per O; two head-dim chunks leave room
        // for a two-stage K / V^T ring only
        l.variant = p.d_chunks == 2 ? 12 : p.dv <= 64 ? 10 : 11;
        l.grid = dim3((unsigned)((d.nq + 2 * ATT_BM - 1) / (2 * ATT_BM)), (unsigned)d.heads, (unsigned)d.batch);
        l.kv_stages = l.variant == 12 ? 2 : 4;
        p.kv_stages = l.kv_stages;
        l.smem = att2_smem_bytes((uint32_t)p.dv, (uint32_t)p.d_chunks, (uint32_t)l.kv_stages);
        int rc2 = l.variant == 10 ? attn2_ensure_attr<3, 1, 4>()
                  : l.variant == 11 ? attn2_ensure_attr<2, 1, 4>() : attn2_ensure_attr<2, 2, 2>();
        if (rc2) return rc2;
        *out = l;
        return PP_OK;
    }
    l.variant = p.d_chunks;
    l.grid = dim3((unsigned)((d.nq + ATT_BM - 1) / ATT_BM), (unsigned)d.heads, (unsigned)d.batch);
    int rc;
    switch (l.variant) {
        case 1: l.smem = AttSmem<1>::total(p.dv); l.kv_stages = AttSmem<1>::KV_STAGES; rc = attn_ensure_attr<1>(); break;
        case 2: l.smem = AttSmem<2>::total(p.dv); l.kv_stages = AttSmem<2>::KV_STAGES; rc = attn_ensure_attr<2>(); break;
        default: l.smem = AttSmem<3>::total(p.dv); l.kv_stages = AttSmem<3>::KV_STAGES; rc = attn_ensure_attr<3>(); break;
    }
    if (rc) return rc;
    *out = l;
    return PP_OK;
}

int attn_launch(const AttnLaunch& l, cudaStream_t s) {
    switch (l.variant) {
        case 10: PP_CUDA_CHECK(launch(attn2_kernel<3, 1, 4>, l.grid, ATT2_THREADS, l.smem, s, l.p)); break;
        case 11: PP_CUDA_CHECK(launch(attn2_kernel<2, 1, 4>, l.grid, ATT2_THREADS, l.smem, s, l.p)); break;
        case 12: PP_CUDA_CHECK(launch(attn2_kernel<2, 2, 2>, l.grid, ATT2_THREADS, l.smem, s, l.p)); break;
        case 1: PP_CUDA_CHECK(launch(attn_fwd_kernel<1>, l.grid, ATT_THREADS, l.smem, s, l.p)); break;
        case 2: PP_CUDA_CHECK(launch(attn_fwd_kernel<2>, l.grid, ATT_THREADS, l.smem, s, l.p)); break;
        default: PP_CUDA_CHECK(launch(attn_fwd_kernel<3>, l.grid, ATT_THREADS, l.smem, s, l.p)); break;
    }
    return PP_OK;
}

}  // namespace pp

extern "C" pp_status pp_attention(const pp_attn_desc* d, pp_stream stream) {
    if (!d) {
        pp::set_last_error("pp_attention: null descriptor");
        return pp::PP_ERR_INVALID;
    }
    pp::AttnLaunch l;
    int rc = pp::attn_prepare(*d, &l);
    if (rc) return rc;
    return pp::attn_launch(l, reinterpret_cast<cudaStream_t>(stream));
}
