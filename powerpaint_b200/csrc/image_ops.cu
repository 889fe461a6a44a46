// Kernels either side of the denoising loop: the row softmax of the VAE's
// single-head 512-channel attention, uint8 image/mask pre-processing, and the decoded image's
// post-processing to uint8.
//
// Reference call sites: AutoencoderKL mid-block attention (diffusers Attention with one head of 512 channels,
// called through `vae.encode` / `vae.decode`, powerpaint/pipelines/pipeline_PowerPaint.py:657-669,:1051);
// `prepare_mask_and_masked_image` (pipeline_PowerPaint.py:39-153: image / 127.5 - 1, mask binarised at 0.5,
// masked_image = image * (mask < 0.5)); `VaeImageProcessor.postprocess` ((x / 2 + 0.5).clamp(0, 1) -> * 255
// -> round -> uint8, pipeline_PowerPaint.py:1062).
#include "common.cuh"
#include "ops.h"

namespace pp {

// ------------------------------------------------------------------------------------
// P[r, :] = softmax(S[r, :]) for fp32 scores (already scaled by the producing GEMM's alpha), bf16 output.
// One block per row, the row is read once into registers (cols <= 16384 for 256 threads x 64 values).
// HBM-bound: reads 4 B and writes 2 B per element.
// ------------------------------------------------------------------------------------
template <int PER_THREAD>
__global__ void softmax_rows_kernel(const float* __restrict__ s, __nv_bfloat16* __restrict__ p, int cols,
                                    int64_t ld_s, int64_t ld_p) {
    pdl_wait();
    pdl_launch_dependents();
    const int64_t row = blockIdx.x;
    const float* sr = s + row * ld_s;
    __nv_bfloat16* pr = p + row * ld_p;
    float v[PER_THREAD];
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < PER_THREAD; ++i) {
        const int c = threadIdx.x + i * blockDim.x;
        v[i] = c < cols ? __ldg(sr + c) : -INFINITY;
        mx = fmaxf(mx, v[i]);
    }
    __shared__ float red[32];
    __shared__ float bcast;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x < 32) {
        float m = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : -INFINITY;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (threadIdx.x == 0) bcast = m;
    }
    __syncthreads();
    mx = bcast;
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < PER_THREAD; ++i) {
        v[i] = __expf(v[i] - mx);  // exp(-inf) = 0 for the padding columns
        sum += v[i];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x < 32) {
        float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (threadIdx.x == 0) bcast = t;
    }
    __syncthreads();
    const float inv = 1.0f / bcast;
#pragma unroll
    for (int i = 0; i < PER_THREAD; ++i) {
        const int c = threadIdx.x + i * blockDim.x;
        if (c < cols) pr[c] = __float2bfloat16_rn(v[i] * inv);
    }
}

int softmax_rows_launch(const float* s, void* p, int64_t rows, int cols, int64_t ld_s, int64_t ld_p, cudaStream_t st) {
    PP_REQUIRE(s && p && rows > 0 && cols > 0, "softmax_rows: invalid arguments");
    PP_REQUIRE(cols <= 256 * 64, "softmax_rows: at most 16384 columns (got %d)", cols);
    PP_REQUIRE(ld_s >= cols && ld_p >= cols, "softmax_rows: row pitch smaller than the row");
    PP_REQUIRE(rows <= 0x7fffffffLL, "softmax_rows: too many rows");
    auto pb = reinterpret_cast<__nv_bfloat16*>(p);
    const int per = (cols + 255) / 256;
    if (per <= 4) PP_CUDA_CHECK(launch(softmax_rows_kernel<4>, dim3((unsigned)rows), 256, 0, st, s, pb, cols, ld_s, ld_p));
    else if (per <= 16) PP_CUDA_CHECK(launch(softmax_rows_kernel<16>, dim3((unsigned)rows), 256, 0, st, s, pb, cols, ld_s, ld_p));
    else PP_CUDA_CHECK(launch(softmax_rows_kernel<64>, dim3((unsigned)rows), 256, 0, st, s, pb, cols, ld_s, ld_p));
    PP_CUDA_CHECK(cudaGetLastError());
    return PP_OK;
}

// ------------------------------------------------------------------------------------
// uint8 NCHW image (+ optional uint8 / fp32 mask) -> bf16 NHWC [n, h*w, c_pad] in [-1, 1], channels
// beyond 3 zero; with a mask the hole is zeroed: out = image * (mask < 0.5) (masked_image, :147).
// mask_mode: 0 none, 1 uint8 [n,1,h,w] (value / 255 binarised at 0.5), 2 fp32 [n,1,h,w].
// out = px / divisor + shift in fp32 like the reference's `image / 127.5 - 1.0` (control image: / 255, + 0).
// ------------------------------------------------------------------------------------
__global__ void image_preprocess_kernel(const uint8_t* __restrict__ img, const void* __restrict__ mask, int mask_mode,
                                        __nv_bfloat16* __restrict__ out, int hw, int c_pad, float divisor, float shift,
                                        int64_t total) {
    pdl_wait();
    pdl_launch_dependents();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t n = i / hw;
        const int64_t p = i - n * hw;
        float keep = 1.f;
        if (mask_mode == 1) keep = reinterpret_cast<const uint8_t*>(mask)[i] >= 128 ? 0.f : 1.f;  // x/255 >= 0.5
        else if (mask_mode == 2) keep = reinterpret_cast<const float*>(mask)[i] >= 0.5f ? 0.f : 1.f;
        const uint8_t* src = img + n * 3 * hw + p;
        float v[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) v[c] = (__fdiv_rn((float)src[(int64_t)c * hw], divisor) + shift) * keep;
        __nv_bfloat16* o = out + i * c_pad;
        uint2 q = make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], 0.f));
        *reinterpret_cast<uint2*>(o) = q;
        for (int c = 4; c < c_pad; c += 4) *reinterpret_cast<uint2*>(o + c) = make_uint2(0u, 0u);
    }
}

int image_preprocess_launch(const uint8_t* img, const void* mask, int mask_mode, void* out, int nb, int hw, int c_pad,
                            float divisor, float shift, cudaStream_t s) {
    PP_REQUIRE(img && out && nb > 0 && hw > 0 && divisor != 0.f, "image_preprocess: invalid arguments");
    PP_REQUIRE(c_pad >= 4 && c_pad % 4 == 0, "image_preprocess: c_pad=%d must be a multiple of 4 >= 4", c_pad);
    PP_REQUIRE(mask_mode >= 0 && mask_mode <= 2 && (mask_mode == 0 || mask), "image_preprocess: mask / mask_mode mismatch");
    const int64_t total = (int64_t)nb * hw;
    const int blocks = (int)std::min<int64_t>((total + 255) / 256, 148 * 16);
    PP_CUDA_CHECK(launch(image_preprocess_kernel, blocks, 256, 0, s, img, mask, mask_mode,
                         reinterpret_cast<__nv_bfloat16*>(out), hw, c_pad, divisor, shift, total));
    PP_CUDA_CHECK(cudaGetLastError());
    return PP_OK;
}

// ------------------------------------------------------------------------------------
// decoded image NHWC (bf16 or fp32, channel pitch c_ld, channels 0..2) -> either
//   out_u8  != NULL: uint8 NHWC [n, h*w, 3] = round(clamp(x / 2 + 0.5, 0, 1) * 255)
//   out_f32 != NULL: fp32 NCHW [n, 3, h*w] = clamp(x / 2 + 0.5, 0, 1)           (output_type="pt")
// ------------------------------------------------------------------------------------
__global__ void image_postprocess_kernel(const void* __restrict__ x, int x_fp32, int c_ld, uint8_t* __restrict__ out_u8,
                                         float* __restrict__ out_f32, int hw, int64_t total) {
    pdl_wait();
    pdl_launch_dependents();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        float v[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float t = x_fp32 ? reinterpret_cast<const float*>(x)[i * c_ld + c]
                                   : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(x)[i * c_ld + c]);
            v[c] = fminf(fmaxf(t * 0.5f + 0.5f, 0.f), 1.f);
        }
        if (out_u8) {
#pragma unroll
            for (int c = 0; c < 3; ++c) out_u8[i * 3 + c] = (uint8_t)rintf(v[c] * 255.f);
        }
        if (out_f32) {
            const int64_t n = i / hw, p = i - n * hw;
#pragma unroll
            for (int c = 0; c < 3; ++c) out_f32[(n * 3 + c) * hw + p] = v[c];
        }
    }
}

int image_postprocess_launch(const void* x, int x_fp32, int c_ld, uint8_t* out_u8, float* out_f32, int nb, int hw,
                             cudaStream_t s) {
    PP_REQUIRE(x && (out_u8 || out_f32) && nb > 0 && hw > 0 && c_ld >= 3, "image_postprocess: invalid arguments");
    const int64_t total = (int64_t)nb * hw;
    const int blocks = (int)std::min<int64_t>((total + 255) / 256, 148 * 16);
    PP_CUDA_CHECK(launch(image_postprocess_kernel, blocks, 256, 0, s, x, x_fp32, c_ld, out_u8, out_f32, hw, total));
    PP_CUDA_CHECK(cudaGetLastError());
    return PP_OK;
}

}  // namespace pp

extern "C" {
pp_status pp_softmax_rows(const float* s, void* p, int64_t rows, int32_t cols, int64_t ld_s, int64_t ld_p,
                          pp_stream stream) {
    return pp::softmax_rows_launch(s, p, rows, cols, ld_s, ld_p, reinterpret_cast<cudaStream_t>(stream));
}
pp_status pp_image_preprocess_u8(const uint8_t* image, const void* mask, int32_t mask_mode, void* out, int32_t nb,
                                 int32_t hw, int32_t c_pad, float divisor, float shift, pp_stream stream) {
    return pp::image_preprocess_launch(image, mask, mask_mode, out, nb, hw, c_pad, divisor, shift,
                                       reinterpret_cast<cudaStream_t>(stream));
}
pp_status pp_image_postprocess(const void* x, int32_t x_is_fp32, int32_t c_ld, uint8_t* out_u8, float* out_f32,
                               int32_t nb, int32_t hw, pp_stream stream) {
    return pp::image_postprocess_launch(x, x_is_fp32, c_ld, out_u8, out_f32, nb, hw,
                                        reinterpret_cast<cudaStream_t>(stream));
}
}
