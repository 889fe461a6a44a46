// Text-encoder kernels: token + position embedding gather with the task-prompt splice,
// and causal self-attention over the 77 prompt tokens.
//
// Reference call sites: `_encode_prompt` -> `self.text_encoder(text_input_ids)` (powerpaint/pipelines/
// pipeline_PowerPaint.py:317-518) with `text_model.embeddings.token_embedding` replaced by
// `EmbeddingLayerWithFixes` (powerpaint/utils/utils.py:256-483: ids >= base vocab are zeroed, then the learned
// [10, 768] task vectors are spliced over the placeholder runs). The splice is resolved on the host into ONE
// gather index per position (base-table row, or row of the concatenated learned vectors), so the device does
// a single pass. CLIP's attention is causal (transformers CLIPTextTransformer builds a causal mask).
#include "common.cuh"
#include "ops.h"

namespace pp {

// out[r, :] = (idx[r] < vocab ? base[idx[r]] : ext[idx[r] - vocab]) + pos[r % seq], bf16 out, fp32 tables
__global__ void embed_gather_kernel(const int32_t* __restrict__ idx, const float* __restrict__ base,
                                    const float* __restrict__ ext, const float* __restrict__ pos,
                                    __nv_bfloat16* __restrict__ out, int vocab, int seq, int dim) {
    pdl_wait();
    pdl_launch_dependents();
    const int r = blockIdx.x;
    const int id = idx[r];
    const float* src = id < vocab ? base + (int64_t)id * dim : ext + (int64_t)(id - vocab) * dim;
    const float* p = pos + (int64_t)(r % seq) * dim;
    for (int c = threadIdx.x * 4; c < dim; c += blockDim.x * 4) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(src + c));
        const float4 b = __ldg(reinterpret_cast<const float4*>(p + c));
        uint2 q = make_uint2(pack_bf16x2(a.x + b.x, a.y + b.y), pack_bf16x2(a.z + b.z, a.w + b.w));
        *reinterpret_cast<uint2*>(out + (int64_t)r * dim + c) = q;
    }
}

int embed_gather_launch(const int32_t* idx, const float* base, const float* ext, const float* pos, void* out, int rows,
                        int vocab, int seq, int dim, cudaStream_t s) {
    PP_REQUIRE(idx && base && pos && out && rows > 0 && vocab > 0 && seq > 0, "embed_gather: invalid arguments");
    PP_REQUIRE(dim > 0 && dim % 4 == 0, "embed_gather: dim=%d must be a multiple of 4", dim);
    PP_CUDA_CHECK(launch(embed_gather_kernel, rows, 128, 0, s, idx, base, ext, pos, reinterpret_cast<__nv_bfloat16*>(out),
                         vocab, seq, dim));
    PP_CUDA_CHECK(cudaGetLastError());
    return PP_OK;
}

// Causal softmax(q k^T * scale) v for short sequences (seq <= 128, head dim <= 64, multiple of 8).
// qkv: [batch * seq, 3 * heads * d] bf16 (q | k | v blocks, head-major inside each), out [batch * seq, heads * d].
// One CTA per (head, sample): K and V of the head live in shared memory as fp32, thread t owns query row t
// (registers: q, the running max / sum and the d output accumulators) and walks keys 0..t.
// 13 GFLOP per prompt for the whole encoder — latency, not throughput, matters here.
template <int D>
__global__ void __launch_bounds__(128) causal_attention_small_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                                      __nv_bfloat16* __restrict__ out, int seq, int heads,
                                                                      float scale) {
    pdl_wait();
    pdl_launch_dependents();
    extern __shared__ float sh[];  // K [seq][D + 1], V [seq][D + 1]
    const int head = blockIdx.x, b = blockIdx.y;
    const int C = heads * D;
    float* sk = sh;
    float* sv = sh + seq * (D + 1);
    const __nv_bfloat16* base = qkv + (int64_t)b * seq * 3 * C + head * D;
    for (int i = threadIdx.x; i < seq * (D / 8); i += blockDim.x) {
        const int t = i / (D / 8), c8 = (i % (D / 8)) * 8;
        const uint4 kq = __ldg(reinterpret_cast<const uint4*>(base + (int64_t)t * 3 * C + C + c8));
        const uint4 vq = __ldg(reinterpret_cast<const uint4*>(base + (int64_t)t * 3 * C + 2 * C + c8));
        float* kd = sk + t * (D + 1) + c8;
        float* vd = sv + t * (D + 1) + c8;
        kd[0] = bf16_lo(kq.x); kd[1] = bf16_hi(kq.x); kd[2] = bf16_lo(kq.y); kd[3] = bf16_hi(kq.y);
        kd[4] = bf16_lo(kq.z); kd[5] = bf16_hi(kq.z); kd[6] = bf16_lo(kq.w); kd[7] = bf16_hi(kq.w);
        vd[0] = bf16_lo(vq.x); vd[1] = bf16_hi(vq.x); vd[2] = bf16_lo(vq.y); vd[3] = bf16_hi(vq.y);
        vd[4] = bf16_lo(vq.z); vd[5] = bf16_hi(vq.z); vd[6] = bf16_lo(vq.w); vd[7] = bf16_hi(vq.w);
    }
    __syncthreads();
    const int t = threadIdx.x;
    if (t >= seq) return;
    float q[D], o[D];
#pragma unroll
    for (int c = 0; c < D; c += 8) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(base + (int64_t)t * 3 * C + c));
        q[c] = bf16_lo(v.x) * scale; q[c + 1] = bf16_hi(v.x) * scale; q[c + 2] = bf16_lo(v.y) * scale;
        q[c + 3] = bf16_hi(v.y) * scale; q[c + 4] = bf16_lo(v.z) * scale; q[c + 5] = bf16_hi(v.z) * scale;
        q[c + 6] = bf16_lo(v.w) * scale; q[c + 7] = bf16_hi(v.w) * scale;
    }
#pragma unroll
    for (int c = 0; c < D; ++c) o[c] = 0.f;
    float m = -INFINITY, l = 0.f;
    for (int j = 0; j <= t; ++j) {
        const float* kr = sk + j * (D + 1);
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < D; ++c) s = fmaf(q[c], kr[c], s);
        const float mn = fmaxf(m, s);
        const float corr = __expf(m - mn), p = __expf(s - mn);
        l = l * corr + p;
        const float* vr = sv + j * (D + 1);
#pragma unroll
        for (int c = 0; c < D; ++c) o[c] = fmaf(o[c], corr, p * vr[c]);
        m = mn;
    }
    const float inv = 1.0f / l;
    __nv_bfloat16* orow = out + ((int64_t)b * seq + t) * C + head * D;
#pragma unroll
    for (int c = 0; c < D; c += 8) {
        uint4 v;
        v.x = pack_bf16x2(o[c] * inv, o[c + 1] * inv);
        v.y = pack_bf16x2(o[c + 2] * inv, o[c + 3] * inv);
        v.z = pack_bf16x2(o[c + 4] * inv, o[c + 5] * inv);
        v.w = pack_bf16x2(o[c + 6] * inv, o[c + 7] * inv);
        *reinterpret_cast<uint4*>(orow + c) = v;
    }
}

int causal_attention_small_launch(const void* qkv, void* out, int batch, int seq, int heads, int d, float scale,
                                  cudaStream_t s) {
    PP_REQUIRE(qkv && out && batch > 0 && heads > 0, "causal_attention: invalid arguments");
    PP_REQUIRE(seq > 0 && seq <= 128, "causal_attention: seq=%d must be in 1..128", seq);
    PP_REQUIRE(d == 8 || d == 16 || d == 32 || d == 64, "causal_attention: head dim %d unsupported (8/16/32/64)", d);
    const size_t smem = sizeof(float) * 2 * (size_t)seq * (d + 1);
    auto q = reinterpret_cast<const __nv_bfloat16*>(qkv);
    auto o = reinterpret_cast<__nv_bfloat16*>(out);
    dim3 grid(heads, batch);
    switch (d) {
        case 8: PP_CUDA_CHECK(launch(causal_attention_small_kernel<8>, grid, 128, smem, s, q, o, seq, heads, scale)); break;
        case 16: PP_CUDA_CHECK(launch(causal_attention_small_kernel<16>, grid, 128, smem, s, q, o, seq, heads, scale)); break;
        case 32: PP_CUDA_CHECK(launch(causal_attention_small_kernel<32>, grid, 128, smem, s, q, o, seq, heads, scale)); break;
        default: {
            static bool done[PP_MAX_DEVICES] = {};
            int dev = 0;
            PP_CUDA_CHECK(cudaGetDevice(&dev));
            if (dev < 0 || dev >= PP_MAX_DEVICES || !done[dev]) {
                PP_CUDA_CHECK(cudaFuncSetAttribute(causal_attention_small_kernel<64>,
                                                   cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 128 * 65 * 4));
                if (dev >= 0 && dev < PP_MAX_DEVICES) done[dev] = true;
            }
            PP_CUDA_CHECK(launch(causal_attention_small_kernel<64>, grid, 128, smem, s, q, o, seq, heads, scale));
        }
    }
    PP_CUDA_CHECK(cudaGetLastError());
    return PP_OK;
}

}  // namespace pp

extern "C" {
pp_status pp_embed_gather(const int32_t* idx, const float* base, const float* ext, const float* pos, void* out,
                          int32_t rows, int32_t vocab, int32_t seq, int32_t dim, pp_stream stream) {
    return pp::embed_gather_launch(idx, base, ext, pos, out, rows, vocab, seq, dim, reinterpret_cast<cudaStream_t>(stream));
}
pp_status pp_causal_attention_small(const void* qkv, void* out, int32_t batch, int32_t seq, int32_t heads, int32_t d,
                                    float scale, pp_stream stream) {
    return pp::causal_attention_small_launch(qkv, out, batch, seq, heads, d, scale, reinterpret_cast<cudaStream_t>(stream));
}
}
