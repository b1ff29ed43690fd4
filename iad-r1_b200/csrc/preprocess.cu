// Image preprocessing on the GPU for the Qwen2-VL / Qwen2.5-VL families: uint8 RGB image ->
// antialiased bicubic resize to the smart-resize target -> 1/255 rescale + mean / std normalisation -> patch rows
// [grid_h * grid_w, C * temporal_patch * patch * patch] in merge-block-major order, bf16, written straight into HBM.
//
// Replaces the CPU stage of the reference's processor call (ref: train/stage_rl/trainer/sc_grpo_trainer.py:606-621 ->
// HF Qwen2VLImageProcessor: image_processing_qwen2_vl.py:62-88 smart_resize, :148-232 resize / rescale / normalize /
// patch layout). The resize follows Pillow's ImagingResample (what `resample=BICUBIC` runs): separable, support
// 2 * max(scale, 1), Keys cubic a = -0.5, horizontal pass then vertical pass, each pass rounded to uint8.
#include "runtime.h"
#include <cuda_bf16.h>
#include <stdint.h>

namespace iadr1 {

__device__ __forceinline__ float cubic_w(float x) {
  x = fabsf(x);
  const float a = -0.5f;
  if (x < 1.f) return ((a + 2.f) * x - (a + 3.f)) * x * x + 1.f;
  if (x < 2.f) return (((x - 5.f) * x + 8.f) * x - 4.f) * a;
  return 0.f;
}

// One resampling pass along one axis: out[o] = round(sum_k w_k in[k]) with Pillow's window for output index o.
// in: [n_lines][in_len][3] (line stride / element stride given in bytes), out likewise.
__global__ void resample_u8_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int n_lines, int in_len, int out_len,
                                   long long in_line_stride, long long in_elem_stride, long long out_line_stride,
                                   long long out_elem_stride) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)n_lines * out_len) return;
  const int line = (int)(idx / out_len), o = (int)(idx - (long long)line * out_len);
  const float scale = (float)in_len / (float)out_len;
  const float fscale = fmaxf(scale, 1.f);
  const float support = 2.f * fscale;
  const float center = (o + 0.5f) * scale;
  int xmin = (int)(center - support + 0.5f);
  if (xmin < 0) xmin = 0;
  int xmax = (int)(center + support + 0.5f);
  if (xmax > in_len) xmax = in_len;
  float acc[3] = {0.f, 0.f, 0.f}, wsum = 0.f;
  const uint8_t* src = in + line * in_line_stride;
  for (int x = xmin; x < xmax; ++x) {
    const float w = cubic_w((x - center + 0.5f) / fscale);
    wsum += w;
    const uint8_t* p = src + x * in_elem_stride;
    acc[0] += w * p[0];
    acc[1] += w * p[1];
    acc[2] += w * p[2];
  }
  uint8_t* dst = out + line * out_line_stride + o * out_elem_stride;
  const float inv = wsum != 0.f ? 1.f / wsum : 0.f;
#pragma unroll
  for (int c = 0; c < 3; ++c) dst[c] = (uint8_t)fminf(255.f, fmaxf(0.f, rintf(acc[c] * inv)));
}

// img [H][W][3] uint8 (H = grid_h * patch, W = grid_w * patch) -> out [grid_h * grid_w][3 * tps * patch * patch] bf16,
// rows in merge-block-major order, both temporal slots = the same frame (a still image is tiled `temporal_patch_size` times).
__global__ void patchify_normalize_kernel(const uint8_t* __restrict__ img, __nv_bfloat16* __restrict__ out, int grid_h, int grid_w,
                                          int patch, int merge, int tps, float3 mean, float3 inv_std) {
  const int W = grid_w * patch;
  const long long n = (long long)grid_h * grid_w * patch * patch;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const int px = (int)(idx % patch), py = (int)((idx / patch) % patch);
  const long long row = idx / ((long long)patch * patch);
  // row -> (block row, block col, row in block, col in block)
  const int mw = (int)(row % merge), mh = (int)((row / merge) % merge);
  const long long blk = row / (merge * merge);
  const int bw = (int)(blk % (grid_w / merge)), bh = (int)(blk / (grid_w / merge));
  const int gy = bh * merge + mh, gx = bw * merge + mw;
  const uint8_t* p = img + ((long long)(gy * patch + py) * W + gx * patch + px) * 3;
  const float m[3] = {mean.x, mean.y, mean.z}, is[3] = {inv_std.x, inv_std.y, inv_std.z};
  const int cols = 3 * tps * patch * patch;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const __nv_bfloat16 v = __float2bfloat16((p[c] * (1.f / 255.f) - m[c]) * is[c]);
    for (int t = 0; t < tps; ++t) out[row * cols + ((c * tps + t) * patch + py) * patch + px] = v;
  }
}

}  // namespace iadr1

using namespace iadr1;

extern "C" {

// scratch: uint8 [in_h * out_w * 3 + out_h * out_w * 3] (used only when the size changes)
int iadr1_image_preprocess_qwen(const void* rgb_u8, int in_h, int in_w, int out_h, int out_w, int patch, int merge, int tps,
                                const float* mean3, const float* std3, void* scratch_u8, void* out_bf16, void* stream) {
  if (in_h <= 0 || in_w <= 0 || out_h <= 0 || out_w <= 0) return set_error("image_preprocess: empty image");
  if (out_h % (patch * merge) || out_w % (patch * merge))
    return set_error("image_preprocess: target %dx%d is not a multiple of patch * merge = %d", out_h, out_w, patch * merge);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const uint8_t* src = static_cast<const uint8_t*>(rgb_u8);
  uint8_t* tmp = static_cast<uint8_t*>(scratch_u8);
  if (in_w != out_w) {            // horizontal pass: lines = rows
    if (!tmp) return set_error("image_preprocess: scratch required for a resize");
    const long long n = (long long)in_h * out_w;
    resample_u8_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(src, tmp, in_h, in_w, out_w, (long long)in_w * 3, 3,
                                                                  (long long)out_w * 3, 3);
    IADR1_CHECK_LAUNCH("image_resample_h");
    src = tmp;
    tmp += (size_t)in_h * out_w * 3;
  }
  if (in_h != out_h) {            // vertical pass: lines = columns
    if (!tmp) return set_error("image_preprocess: scratch required for a resize");
    const long long n = (long long)out_w * out_h;
    resample_u8_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(src, tmp, out_w, in_h, out_h, 3, (long long)out_w * 3, 3,
                                                                  (long long)out_w * 3);
    IADR1_CHECK_LAUNCH("image_resample_v");
    src = tmp;
  }
  const long long n = (long long)out_h * out_w;
  patchify_normalize_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(
      src, static_cast<__nv_bfloat16*>(out_bf16), out_h / patch, out_w / patch, patch, merge, tps,
      make_float3(mean3[0], mean3[1], mean3[2]), make_float3(1.f / std3[0], 1.f / std3[1], 1.f / std3[2]));
  IADR1_CHECK_LAUNCH("image_patchify");
  return 0;
}

}  // extern "C"
