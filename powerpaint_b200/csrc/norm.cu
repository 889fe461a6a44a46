// GroupNorm(+SiLU)(+channel concat) and LayerNorm over channels-last bf16 activations.
//
// Reference semantics: nn.GroupNorm(groups, C, eps, affine=True) followed by SiLU in
// ResnetBlock2D (norm1/norm2), Transformer2DModel.norm (eps 1e-6, no SiLU) and
// conv_norm_out (powerpaint/models/unet_2d_condition.py:466,1351-1353); nn.LayerNorm(C)
// x3 per BasicTransformerBlock. Statistics are biased (divide by count), computed in fp32.
//
// These kernels are HBM-bound. GroupNorm statistics normally come for free: the GEMM / conv that
// PRODUCED the tensor emits per-tile channel sums from its epilogue (pp_gemm_desc.chan_stats) and a tiny
// finalize kernel folds them into (mean, rstd) per (sample, group), so GroupNorm is ONE pass (read x,
// write y). Tensors without producer statistics (odd shapes, ControlNet skip sums) take the standalone
// statistics kernel (a second read, normally an L2 hit). LayerNorm reads once (row held in registers)
// and writes once. Algorithmic bytes: GN 2*|x|*2B, LN 2*|x|*2B.
// Variance: every path sums x - shift (shift = an element of the group / of the tile), so a large mean over a
// small spread does not cancel; per-tile records with different shifts are combined as (count, mean, M2)
// with Chan's parallel formula in fp64.
//
// GroupNorm also performs the up-path `torch.cat([hidden, skip], dim=1)`
// (unet_2d_blocks.py:2589,2732): it normalises over the virtual concat of two sources and
// writes one dense tensor, so the concat never exists un-normalised.
#include "common.cuh"
#include "ops.h"

namespace pp {

// ------------------------------------------------------------------------------------
// GroupNorm pass 1: per-(sample, group) sum and sum of squares — deterministic (no atomics on data).
// grid = (chunks, batch); block = CV * k threads where CV = C/8 channel vectors; each thread
// owns one 8-channel vector and strides over pixels, so loads are 16-byte and coalesced along
// channels. The per-thread per-channel partials go to shared memory, 2*groups threads fold them per
// group in a fixed order and write the block's partial to `partials`; the last block of a sample to
// finish (ticket counter) adds the blocks' partials in chunk order into `stats`. Every float addition
// therefore happens in an order that depends only on the launch geometry: results are bit-identical
// from run to run (the ticket decides WHO does the final sum, not in which order).
// Scratch layout (pp_group_norm_scratch_bytes): stats [batch][groups][2] f32 | tickets [batch] u32
// (zero before the call, left zero by it) | partials [batch][chunks][groups][2] f32.
// ------------------------------------------------------------------------------------
struct GnGeometry {
    int threads, lanes, chunks, ppb;
    size_t smem;           // dynamic shared memory of the statistics kernel
    size_t ticket_offset;  // bytes from the scratch base
    size_t partial_offset;
    size_t scratch_bytes;
};

static GnGeometry gn_geometry(int batch, int hw, int C, int groups) {
    GnGeometry g;
    const int CV = C / 8;
    int k = 256 / CV;
    if (k < 1) k = 1;
    g.threads = CV * k;
    g.lanes = k;
    // ~8 blocks per SM over the machine, at least 8 pixels per pixel lane
    int chunks = (148 * 8 + batch - 1) / batch;
    int ppb = (hw + chunks - 1) / chunks;
    if (ppb < k * 8) ppb = k * 8;
    g.ppb = ppb;
    g.chunks = (hw + ppb - 1) / ppb;
    g.smem = sizeof(float) * 2 * (size_t)k * C + 16;
    const size_t stats_bytes = sizeof(float) * 2 * (size_t)groups * batch;
    g.ticket_offset = (stats_bytes + 15) & ~(size_t)15;
    g.partial_offset = g.ticket_offset + ((sizeof(uint32_t) * (size_t)batch + 15) & ~(size_t)15);
    g.scratch_bytes = g.partial_offset + sizeof(float) * 2 * (size_t)groups * g.chunks * batch;
    return g;
}

// shift of a (sample, group): its first element (pixel 0, first channel of the group) in the virtual concat
__device__ __forceinline__ float gn_group_shift(const __nv_bfloat16* x0, const __nv_bfloat16* x1, int c0, int c1, int hw,
                                                int n, int c) {
    const __nv_bfloat16* p = c < c0 ? x0 + ((int64_t)n * hw) * c0 + c : x1 + ((int64_t)n * hw) * c1 + (c - c0);
    return __bfloat162float(__ldg(p));
}

__global__ void gn_stats_kernel(const __nv_bfloat16* __restrict__ x0, const __nv_bfloat16* __restrict__ x1,
                                int c0, int c1, int hw, int groups, int pix_per_block,
                                float eps, float* __restrict__ stats, uint32_t* __restrict__ tickets,
                                float* __restrict__ partials) {
    pdl_wait();  // inputs come from the preceding kernel
    pdl_launch_dependents();
    extern __shared__ float sh[];  // [lanes][C] sums, then [lanes][C] sums of squares, then the ticket
    const int C = c0 + c1;
    const int CV = C / 8;
    const int cpg = C / groups;
    const int n = blockIdx.y;
    const int lanes = blockDim.x / CV;  // pixel lanes
    const int cv = threadIdx.x % CV;
    const int pl = threadIdx.x / CV;
    float* sh_s = sh;
    float* sh_ss = sh + lanes * C;
    {
        const int c = cv * 8;
        const __nv_bfloat16* src;
        int cs, co;
        if (c < c0) { src = x0; cs = c0; co = c; } else { src = x1; cs = c1; co = c - c0; }
        const int p_begin = blockIdx.x * pix_per_block;
        const int p_end = min(hw, p_begin + pix_per_block);
        float s[8], ss[8], sh8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            s[j] = 0.f; ss[j] = 0.f;
            sh8[j] = gn_group_shift(x0, x1, c0, c1, hw, n, ((c + j) / cpg) * cpg);
        }
        const __nv_bfloat16* sp = src + ((int64_t)n * hw) * cs + co;
        auto accumulate = [&](const uint4& q) {
            const float v[8] = {bf16_lo(q.x), bf16_hi(q.x), bf16_lo(q.y), bf16_hi(q.y),
                                bf16_lo(q.z), bf16_hi(q.z), bf16_lo(q.w), bf16_hi(q.w)};
#pragma unroll
            for (int j = 0; j < 8; ++j) { const float dlt = v[j] - sh8[j]; s[j] += dlt; ss[j] = fmaf(dlt, dlt, ss[j]); }
        };
        // four independent 16-byte loads in flight per thread
        int p = p_begin + pl;
        for (; p + 3 * lanes < p_end; p += 4 * lanes) {
            const uint4 q0 = __ldg(reinterpret_cast<const uint4*>(sp + (int64_t)p * cs));
            const uint4 q1 = __ldg(reinterpret_cast<const uint4*>(sp + (int64_t)(p + lanes) * cs));
            const uint4 q2 = __ldg(reinterpret_cast<const uint4*>(sp + (int64_t)(p + 2 * lanes) * cs));
            const uint4 q3 = __ldg(reinterpret_cast<const uint4*>(sp + (int64_t)(p + 3 * lanes) * cs));
            accumulate(q0); accumulate(q1); accumulate(q2); accumulate(q3);
        }
        for (; p < p_end; p += lanes) accumulate(__ldg(reinterpret_cast<const uint4*>(sp + (int64_t)p * cs)));
        float4* ds = reinterpret_cast<float4*>(sh_s + pl * C + c);
        float4* dss = reinterpret_cast<float4*>(sh_ss + pl * C + c);
        ds[0] = make_float4(s[0], s[1], s[2], s[3]);
        ds[1] = make_float4(s[4], s[5], s[6], s[7]);
        dss[0] = make_float4(ss[0], ss[1], ss[2], ss[3]);
        dss[1] = make_float4(ss[4], ss[5], ss[6], ss[7]);
    }
    __syncthreads();
    // fold: thread i < 2 * groups owns (group i / 2, statistic i % 2); fixed order: lanes outer, channels inner
    const int chunks = gridDim.x;
    for (int i = threadIdx.x; i < groups * 2; i += blockDim.x) {
        const int g = i >> 1;
        const float* srcv = (i & 1) ? sh_ss : sh_s;
        float acc = 0.f;
        for (int l = 0; l < lanes; ++l)
            for (int cc = 0; cc < cpg; ++cc) acc += srcv[l * C + g * cpg + cc];
        partials[(((int64_t)n * chunks + blockIdx.x) * groups) * 2 + i] = acc;
    }
    // ticket: the last block of this sample to arrive sums the partials of all its blocks in chunk order
    __threadfence();
    __syncthreads();
    uint32_t* sh_ticket = reinterpret_cast<uint32_t*>(sh + 2 * lanes * C);
    if (threadIdx.x == 0) *sh_ticket = atomicAdd(&tickets[n], 1u);
    __syncthreads();
    if (*sh_ticket != (uint32_t)(chunks - 1)) return;
    __threadfence();
    // every block summed (x - shift) with the SAME per-(sample, group) shift (the group's first element), so the
    // block partials add up directly; the sums are of spread-sized numbers, the last step runs in fp64
    for (int g = threadIdx.x; g < groups; g += blockDim.x) {
        const float* pp_ = partials + (((int64_t)n * chunks) * groups + g) * 2;
        double s = 0.0, ss = 0.0;
        for (int ch = 0; ch < chunks; ++ch) {
            s += (double)__ldcg(pp_ + (int64_t)ch * groups * 2);
            ss += (double)__ldcg(pp_ + (int64_t)ch * groups * 2 + 1);
        }
        const double cnt = (double)hw * (double)cpg;
        const double md = s / cnt;
        const double var = fmax(ss / cnt - md * md, 0.0);
        stats[((int64_t)n * groups + g) * 2] = (float)((double)gn_group_shift(x0, x1, c0, c1, hw, n, g * cpg) + md);
        stats[((int64_t)n * groups + g) * 2 + 1] = (float)(1.0 / sqrt(var + (double)eps));
    }
    if (threadIdx.x == 0) tickets[n] = 0u;  // ready for the next call
}

// GroupNorm statistics from the producers' epilogue partial sums (pp_gemm_desc.chan_stats): one block per
// sample. Every thread folds the tiles of its channels (fixed order), the channels of a group are then
// folded per group; all in fp64 with Chan's combination. Writes (mean, rstd) like gn_stats_kernel.
struct GnPartSrc {
    const float* part;
    int channels, segs, tiles_per_group, tiles_x, tiles_y, bw, bh, wo, ho;
};

// One WARP per (sample, group): the lanes split the (channel, tile) records of the group, every record is re-referenced
// to ONE shift R (the group's first record) so the shifted sums simply add up, and a fixed-order butterfly folds the
// lanes — bit-identical from run to run. (The first version ran one block per sample with one thread per channel
// walking all tiles serially: 16 blocks, 11.6 us per launch, 3.4 % of the C2 step for 61 launches.)
constexpr int GN_FIN_WARPS = 4;
__global__ void __launch_bounds__(GN_FIN_WARPS * 32) gn_finalize_kernel(GnPartSrc s0, GnPartSrc s1, int groups, int batch,
                                                                       float eps, float* __restrict__ stats) {
    pdl_wait();
    pdl_launch_dependents();
    const int lane = threadIdx.x & 31;
    const int item = blockIdx.x * GN_FIN_WARPS + (threadIdx.x >> 5);  // (sample, group)
    if (item >= batch * groups) return;
    const int n = item / groups, g = item - n * groups;
    const int C = s0.channels + s1.channels;
    const int cpg = C / groups;
    // the group's channels [c_lo, c_hi) may straddle the two sources of a skip concat
    const int c_lo = g * cpg, c_hi = c_lo + cpg;
    auto rec_base = [&](const GnPartSrc& src, int cl) -> const float* {
        const int gidx = n / src.segs, seg = n - gidx * src.segs;
        return src.part + (((int64_t)gidx * src.tiles_per_group * src.segs + seg) * src.channels + cl) * 4;
    };
    const double R = (double)__ldcg((c_lo < s0.channels ? rec_base(s0, c_lo) : rec_base(s1, c_lo - s0.channels)) + 2);
    // sum(x - R) = A + n d,  sum((x - R)^2) = B + 2 d A + n d^2,  d = shift_tile - R
    double S = 0.0, SS = 0.0, cnt = 0.0;
#pragma unroll
    for (int si = 0; si < 2; ++si) {
        const GnPartSrc& src = si == 0 ? s0 : s1;
        const int off = si == 0 ? 0 : s0.channels;
        const int a = max(c_lo, off) - off, b = min(c_hi, off + src.channels) - off;  // local channel range
        if (b <= a) continue;
        const int T = src.tiles_per_group;
        const int items = (b - a) * T;  // (channel, tile) records, split across the lanes
        const float* base = rec_base(src, a);
        const int64_t tstride = (int64_t)src.segs * src.channels * 4;
#pragma unroll 4
        for (int i = lane; i < items; i += 32) {
            const int cc = i / T, t = i - cc * T;
            const float4 v = __ldcg(reinterpret_cast<const float4*>(base + cc * 4 + t * tstride));
            const int ty = t / src.tiles_x, tx = t - ty * src.tiles_x;
            const double cb = (double)(min(src.bw, src.wo - tx * src.bw) * min(src.bh, src.ho - ty * src.bh));
            const double A = (double)v.x, B = (double)v.y, d = (double)v.z - R;
            S += A + cb * d;
            SS += B + d * (2.0 * A + cb * d);
            cnt += cb;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        S += __shfl_xor_sync(0xffffffffu, S, o);
        SS += __shfl_xor_sync(0xffffffffu, SS, o);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    if (lane == 0) {
        const double md = S / cnt;
        const double var = fmax(SS - S * md, 0.0) / cnt;
        stats[((int64_t)n * groups + g) * 2] = (float)(R + md);
        stats[((int64_t)n * groups + g) * 2 + 1] = (float)(1.0 / sqrt(var + (double)eps));
    }
}

// GroupNorm pass 2: y = (x - mean) * rstd * gamma + beta, optional SiLU; writes the concat.
// grid = (pixel chunks, batch); like pass 1 every thread owns one 8-channel vector (so gamma / beta /
// mean / rstd live in registers for the whole kernel) and strides over pixels: 32-bit index math only.
__global__ void gn_apply_kernel(const __nv_bfloat16* __restrict__ x0, const __nv_bfloat16* __restrict__ x1,
                                int c0, int c1, int hw, int groups, const float* __restrict__ gamma,
                                const float* __restrict__ beta, int silu,
                                const float* __restrict__ stats, __nv_bfloat16* __restrict__ y,
                                int pix_per_block) {
    pdl_wait();  // inputs come from the preceding kernel
    pdl_launch_dependents();
    const int C = c0 + c1;
    const int CV = C / 8;
    const int cpg = C / groups;
    const int n = blockIdx.y;
    const int lanes = blockDim.x / CV;
    const int cv = threadIdx.x % CV;
    const int pl = threadIdx.x / CV;
    if (pl >= lanes) return;
    const int c = cv * 8;
    const __nv_bfloat16* src;
    int cs, co;
    if (c < c0) { src = x0; cs = c0; co = c; } else { src = x1; cs = c1; co = c - c0; }
    // per-channel scale / shift: y = x * a + b with a = rstd * gamma, b = beta - mean * a
    float a[8], bsh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int g = (c + j) / cpg;
        const float mean = __ldg(&stats[((int64_t)n * groups + g) * 2]);
        const float rstd = __ldg(&stats[((int64_t)n * groups + g) * 2 + 1]);
        a[j] = rstd * __ldg(gamma + c + j);
        bsh[j] = __ldg(beta + c + j) - mean * a[j];
    }
    const int p_begin = blockIdx.x * pix_per_block;
    const int p_end = min(hw, p_begin + pix_per_block);
    const __nv_bfloat16* sp = src + ((int64_t)n * hw) * cs + co;
    __nv_bfloat16* yp = y + ((int64_t)n * hw) * C + c;
    auto emit = [&](const uint4& q, int p) {
        float v[8] = {bf16_lo(q.x), bf16_hi(q.x), bf16_lo(q.y), bf16_hi(q.y),
                      bf16_lo(q.z), bf16_hi(q.z), bf16_lo(q.w), bf16_hi(q.w)};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float o = fmaf(v[j], a[j], bsh[j]);
            v[j] = silu ? silu_f(o) : o;
        }
        uint4 o;
        o.x = pack_bf16x2(v[0], v[1]);
        o.y = pack_bf16x2(v[2], v[3]);
        o.z = pack_bf16x2(v[4], v[5]);
        o.w = pack_bf16x2(v[6], v[7]);
        *reinterpret_cast<uint4*>(yp + (int64_t)p * C) = o;
    };
    int p = p_begin + pl;
    for (; p + 3 * lanes < p_end; p += 4 * lanes) {  // four loads in flight per thread (see pass 1)
        const uint4 q0 = __ldg(reinterpret_cast<const uint4*>(sp + (int64_t)p * cs));
        const uint4 q1 = __ldg(reinterpret_cast<const uint4*>(sp + (int64_t)(p + lanes) * cs));
        const uint4 q2 = __ldg(reinterpret_cast<const uint4*>(sp + (int64_t)(p + 2 * lanes) * cs));
        const uint4 q3 = __ldg(reinterpret_cast<const uint4*>(sp + (int64_t)(p + 3 * lanes) * cs));
        emit(q0, p); emit(q1, p + lanes); emit(q2, p + 2 * lanes); emit(q3, p + 3 * lanes);
    }
    for (; p < p_end; p += lanes) emit(__ldg(reinterpret_cast<const uint4*>(sp + (int64_t)p * cs)), p);
}

int group_norm_validate(const pp_gn_desc& d) {
    PP_REQUIRE(d.x0 && d.y && d.gamma && d.beta && d.stats, "group_norm: null pointer");
    PP_REQUIRE((d.x1 == nullptr) == (d.c1 == 0), "group_norm: x1/c1 mismatch");
    const int C = d.c0 + d.c1;
    PP_REQUIRE(d.c0 > 0 && d.c0 % 8 == 0 && d.c1 % 8 == 0, "group_norm: channels must be multiples of 8 (c0=%d c1=%d)", d.c0, d.c1);
    PP_REQUIRE(d.groups > 0 && C % d.groups == 0, "group_norm: C=%d not divisible by groups=%d", C, d.groups);
    PP_REQUIRE(C / 8 <= 1024, "group_norm: C=%d too large", C);
    PP_REQUIRE(d.batch > 0 && d.hw > 0, "group_norm: empty input");
    PP_REQUIRE((reinterpret_cast<uintptr_t>(d.stats) & 15) == 0, "group_norm: scratch not 16-byte aligned");
    if (d.from_partials) {
        PP_REQUIRE(d.part0 && d.geom0.supported && d.geom0.channels == d.c0, "group_norm: part0 / geom0 invalid");
        PP_REQUIRE((d.x1 == nullptr) || (d.part1 && d.geom1.supported && d.geom1.channels == d.c1),
                   "group_norm: part1 / geom1 invalid");
        PP_REQUIRE(d.geom0.wo * d.geom0.ho == d.hw && (!d.x1 || d.geom1.wo * d.geom1.ho == d.hw),
                   "group_norm: partial-sum geometry does not cover hw=%d pixels", d.hw);
        PP_REQUIRE((size_t)C * 16 <= 48 * 1024, "group_norm: C=%d too large for the finalize kernel", C);
        return PP_OK;
    }
    PP_REQUIRE(gn_geometry(d.batch, d.hw, C, d.groups).smem <= 48 * 1024,
               "group_norm: C=%d needs more than 48 KB of shared memory for the block reduction", C);
    return PP_OK;
}

static GnPartSrc part_src(const float* part, const pp_stats_geom& g) {
    GnPartSrc s;
    s.part = part;
    s.channels = g.channels; s.segs = g.segs; s.tiles_per_group = g.tiles_per_group;
    s.tiles_x = g.tiles_x; s.tiles_y = g.tiles_y; s.bw = g.bw; s.bh = g.bh; s.wo = g.wo; s.ho = g.ho;
    return s;
}

int64_t group_norm_scratch_bytes(int batch, int hw, int channels, int groups) {
    if (batch <= 0 || hw <= 0 || channels <= 0 || channels % 8 || groups <= 0) return 0;
    return (int64_t)gn_geometry(batch, hw, channels, groups).scratch_bytes;
}

int group_norm_launch(const pp_gn_desc& d, cudaStream_t s) {
    int rc = group_norm_validate(d);
    if (rc) return rc;
    const int C = d.c0 + d.c1;
    const GnGeometry g = gn_geometry(d.batch, d.hw, C, d.groups);
    uint8_t* scratch = reinterpret_cast<uint8_t*>(d.stats);
    uint32_t* tickets = reinterpret_cast<uint32_t*>(scratch + g.ticket_offset);
    float* partials = reinterpret_cast<float*>(scratch + g.partial_offset);
    const int threads = g.threads, chunks = g.chunks, ppb = g.ppb;
    if (d.from_partials) {
        GnPartSrc s0 = part_src(d.part0, d.geom0), s1;
        memset(&s1, 0, sizeof(s1));
        if (d.x1) s1 = part_src(d.part1, d.geom1);
        const int items = d.batch * d.groups;
        PP_CUDA_CHECK(launch(gn_finalize_kernel, dim3((items + GN_FIN_WARPS - 1) / GN_FIN_WARPS), GN_FIN_WARPS * 32, 0, s, s0,
                             s1, d.groups, d.batch, d.eps, d.stats));
    } else {
        if (!d.stats_prezeroed) PP_CUDA_CHECK(cudaMemsetAsync(tickets, 0, sizeof(uint32_t) * d.batch, s));
        PP_CUDA_CHECK(launch(gn_stats_kernel, dim3(chunks, d.batch), threads, g.smem, s,
            reinterpret_cast<const __nv_bfloat16*>(d.x0), reinterpret_cast<const __nv_bfloat16*>(d.x1), d.c0,
            d.c1, d.hw, d.groups, ppb, d.eps, d.stats, tickets, partials));
    }
    PP_CUDA_CHECK(cudaGetLastError());
    PP_CUDA_CHECK(launch(gn_apply_kernel, dim3(chunks, d.batch), threads, 0, s, 
        reinterpret_cast<const __nv_bfloat16*>(d.x0), reinterpret_cast<const __nv_bfloat16*>(d.x1), d.c0,
        d.c1, d.hw, d.groups, d.gamma, d.beta, d.silu, d.stats,
        reinterpret_cast<__nv_bfloat16*>(d.y), ppb));
    PP_CUDA_CHECK(cudaGetLastError());
    return PP_OK;
}

// ------------------------------------------------------------------------------------
// LayerNorm: one warp per row, the row lives in registers (two-pass mean/variance).
// ------------------------------------------------------------------------------------
template <int MAX_VEC>
__global__ void layer_norm_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                  const float* __restrict__ gamma, const float* __restrict__ beta,
                                  int rows, int c, float eps) {
    pdl_wait();  // inputs come from the preceding kernel
    pdl_launch_dependents();
    const int warps_per_block = blockDim.x >> 5;
    const int lane = threadIdx.x & 31;
    const int CV = c / 8;
    for (int64_t row = (int64_t)blockIdx.x * warps_per_block + (threadIdx.x >> 5); row < rows;
         row += (int64_t)gridDim.x * warps_per_block) {
        float v[MAX_VEC][8];
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < MAX_VEC; ++i) {
            const int cv = lane + 32 * i;
            if (cv < CV) {
                uint4 q = __ldg(reinterpret_cast<const uint4*>(x + row * c + cv * 8));
                v[i][0] = bf16_lo(q.x); v[i][1] = bf16_hi(q.x); v[i][2] = bf16_lo(q.y); v[i][3] = bf16_hi(q.y);
                v[i][4] = bf16_lo(q.z); v[i][5] = bf16_hi(q.z); v[i][6] = bf16_lo(q.w); v[i][7] = bf16_hi(q.w);
#pragma unroll
                for (int j = 0; j < 8; ++j) sum += v[i][j];
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float mean = sum / (float)c;
        float sq = 0.f;
#pragma unroll
        for (int i = 0; i < MAX_VEC; ++i) {
            const int cv = lane + 32 * i;
            if (cv < CV) {
#pragma unroll
                for (int j = 0; j < 8; ++j) { const float t = v[i][j] - mean; sq += t * t; }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
        const float rstd = rsqrtf(sq / (float)c + eps);
#pragma unroll
        for (int i = 0; i < MAX_VEC; ++i) {
            const int cv = lane + 32 * i;
            if (cv < CV) {
                const int cc = cv * 8;
                float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + cc));
                float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + cc + 4));
                float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + cc));
                float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + cc + 4));
                float o[8];
                o[0] = (v[i][0] - mean) * rstd * g0.x + b0.x; o[1] = (v[i][1] - mean) * rstd * g0.y + b0.y;
                o[2] = (v[i][2] - mean) * rstd * g0.z + b0.z; o[3] = (v[i][3] - mean) * rstd * g0.w + b0.w;
                o[4] = (v[i][4] - mean) * rstd * g1.x + b1.x; o[5] = (v[i][5] - mean) * rstd * g1.y + b1.y;
                o[6] = (v[i][6] - mean) * rstd * g1.z + b1.z; o[7] = (v[i][7] - mean) * rstd * g1.w + b1.w;
                uint4 q;
                q.x = pack_bf16x2(o[0], o[1]); q.y = pack_bf16x2(o[2], o[3]);
                q.z = pack_bf16x2(o[4], o[5]); q.w = pack_bf16x2(o[6], o[7]);
                *reinterpret_cast<uint4*>(y + row * c + cc) = q;
            }
        }
    }
}

int layer_norm_launch(const void* x, void* y, const float* gamma, const float* beta, int rows, int c,
                      float eps, cudaStream_t s) {
    PP_REQUIRE(x && y && gamma && beta, "layer_norm: null pointer");
    PP_REQUIRE(rows > 0 && c > 0 && c % 8 == 0, "layer_norm: rows=%d c=%d invalid", rows, c);
    const int CV = c / 8;
    PP_REQUIRE(CV <= 32 * 8, "layer_norm: c=%d too large", c);
    const int warps = 8;
    int blocks = std::min((rows + warps - 1) / warps, 148 * 8);
    auto xb = reinterpret_cast<const __nv_bfloat16*>(x);
    auto yb = reinterpret_cast<__nv_bfloat16*>(y);
    if (CV <= 32) PP_CUDA_CHECK(launch(layer_norm_kernel<1>, blocks, warps * 32, 0, s, xb, yb, gamma, beta, rows, c, eps));
    else if (CV <= 64) PP_CUDA_CHECK(launch(layer_norm_kernel<2>, blocks, warps * 32, 0, s, xb, yb, gamma, beta, rows, c, eps));
    else if (CV <= 96) PP_CUDA_CHECK(launch(layer_norm_kernel<3>, blocks, warps * 32, 0, s, xb, yb, gamma, beta, rows, c, eps));
    else if (CV <= 160) PP_CUDA_CHECK(launch(layer_norm_kernel<5>, blocks, warps * 32, 0, s, xb, yb, gamma, beta, rows, c, eps));
    else PP_CUDA_CHECK(launch(layer_norm_kernel<8>, blocks, warps * 32, 0, s, xb, yb, gamma, beta, rows, c, eps));
    PP_CUDA_CHECK(cudaGetLastError());
    return PP_OK;
}

}  // namespace pp

extern "C" int64_t pp_group_norm_scratch_bytes(int32_t batch, int32_t hw, int32_t channels, int32_t groups) {
    return pp::group_norm_scratch_bytes(batch, hw, channels, groups);
}

extern "C" {
pp_status pp_group_norm(const pp_gn_desc* d, pp_stream stream) {
    if (!d) { pp::set_last_error("pp_group_norm: null descriptor"); return pp::PP_ERR_INVALID; }
    return pp::group_norm_launch(*d, reinterpret_cast<cudaStream_t>(stream));
}
pp_status pp_layer_norm(const void* x, void* y, const float* gamma, const float* beta, int32_t rows,
                        int32_t c, float eps, pp_stream stream) {
    return pp::layer_norm_launch(x, y, gamma, beta, rows, c, eps, reinterpret_cast<cudaStream_t>(stream));
}
}
