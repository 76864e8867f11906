// optimizer.cu -- AdamW over a list of fp32 tensors in one launch (decoupled weight decay, bias correction; the update of
// torch.optim.AdamW / its fused CUDA path, which the reference's get_optimizer builds: src/models/containers.py:100-114).
//
// HBM-bound: per element 16 B read (p, g, m, v) + 12 B written (p, m, v).  A device-resident table describes the tensors, a second
// one maps every chunk of kChunk elements to (tensor, offset); one CTA per chunk, 16-byte accesses, four independent elements in
// flight per thread.  torch's multi-tensor kernel needs 43 launches for the 546 tensors of the fine-tune step (its kernel-argument
// block holds a few dozen pointers per launch) and reaches 4.4 TB/s; this one is a single launch.
#include <math.h>

#include "common.h"

namespace dicow {
namespace {

constexpr int kChunk = 16384;  // elements per CTA
constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads) adamw_kernel(const dicow_adamw_tensor_args_t* __restrict__ tensors,
                                                         const int2* __restrict__ chunks, float beta1, float beta2, float eps) {
  const int2 ch = chunks[blockIdx.x];  // {tensor index, chunk index within the tensor}
  const dicow_adamw_tensor_args_t t = tensors[ch.x];
  const long long base = (long long)ch.y * kChunk;
  const long long n = t.n - base < kChunk ? t.n - base : kChunk;
  float* __restrict__ p = t.p + base;
  const float* __restrict__ g = t.g + base;
  float* __restrict__ m = t.m + base;
  float* __restrict__ v = t.v + base;
  const float decay = 1.0f - t.lr * t.weight_decay;
  const float step_size = t.lr / t.bias_correction1;
  const float inv_bc2_sqrt = 1.0f / t.bias_correction2_sqrt;
  const float omb1 = 1.0f - beta1, omb2 = 1.0f - beta2;
  auto update = [&](float& pe, float ge, float& me, float& ve) {
    pe *= decay;
    me = fmaf(ge - me, omb1, me);  // lerp(m, g, 1 - beta1)
    ve = fmaf(beta2, ve, omb2 * ge * ge);
    const float denom = sqrtf(ve) * inv_bc2_sqrt + eps;
    pe -= step_size * me / denom;
  };
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                     reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  if (vec) {
    const int n4 = (int)(n >> 2);
    for (int i = threadIdx.x; i < n4; i += kThreads) {
      float4 pv = reinterpret_cast<const float4*>(p)[i];
      const float4 gv = __ldg(reinterpret_cast<const float4*>(g) + i);
      float4 mv = reinterpret_cast<const float4*>(m)[i];
      float4 vv = reinterpret_cast<const float4*>(v)[i];
      update(pv.x, gv.x, mv.x, vv.x), update(pv.y, gv.y, mv.y, vv.y), update(pv.z, gv.z, mv.z, vv.z), update(pv.w, gv.w, mv.w, vv.w);
      reinterpret_cast<float4*>(p)[i] = pv;
      reinterpret_cast<float4*>(m)[i] = mv;
      reinterpret_cast<float4*>(v)[i] = vv;
    }
    for (int i = (n4 << 2) + threadIdx.x; i < n; i += kThreads) update(p[i], g[i], m[i], v[i]);
  } else {
    for (int i = threadIdx.x; i < n; i += kThreads) update(p[i], g[i], m[i], v[i]);
  }
}

}  // namespace
}  // namespace dicow

using namespace dicow;

extern "C" int dicow_adamw_chunk_elems(void) { return kChunk; }

extern "C" int dicow_adamw_step(dicow_handle_t h, const dicow_adamw_tensor_args_t* tensors, const int32_t* chunks, int n_chunks,
                                float beta1, float beta2, float eps, void* stream_) {
  if (h == nullptr) return DICOW_ERR_INVALID_ARG;
  dicow_ctx* ctx = h;
  DICOW_REQUIRE(ctx, tensors != nullptr && chunks != nullptr && n_chunks >= 1, "dicow_adamw_step: bad args");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  adamw_kernel<<<n_chunks, kThreads, 0, stream>>>(tensors, reinterpret_cast<const int2*>(chunks), beta1, beta2, eps);
  DICOW_CUDA_OK(ctx, cudaGetLastError());
  return DICOW_OK;
}
