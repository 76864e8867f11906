// common.h -- host-side context, error plumbing and TMA tensor-map encoding shared by all kernels.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/dicow_b200.h"

struct dicow_ctx {
  int device = 0;
  int num_sms = 148;   // SMs the persistent kernels size their grids for (dicow_set_sm_budget can lower it)
  int phys_sms = 148;  // multiProcessorCount
  int max_smem_optin = 0;
  char err[512] = {0};
  // cuTensorMapEncodeTiled fetched through cudaGetDriverEntryPoint (no link-time libcuda dependency)
  void* encode_tiled = nullptr;
  void* attn_prof = nullptr;  // debug: see dicow_debug_set_attention_profile
  float* mel_tables = nullptr;  // Hann window + DFT basis (mel.cu), owned by the handle
};

namespace dicow {

int set_error(dicow_ctx* ctx, int code, const char* fmt, ...);

#define DICOW_CUDA_OK(ctx, expr)                                                                            \
  do {                                                                                                      \
    cudaError_t _e = (expr);                                                                                \
    if (_e != cudaSuccess)                                                                                  \
      return dicow::set_error((ctx), DICOW_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                              __FILE__, __LINE__);                                                          \
  } while (0)

#define DICOW_REQUIRE(ctx, cond, ...)                                              \
  do {                                                                             \
    if (!(cond)) return dicow::set_error((ctx), DICOW_ERR_INVALID_ARG, __VA_ARGS__); \
  } while (0)

// Encode a bf16 tiled tensor map with 128-byte swizzle and zero OOB fill.
//   rank 2 or 3; dims[] innermost first (elements); strides_bytes[] for dims 1..rank-1; box[] elements.
int make_tmap_bf16(dicow_ctx* ctx, CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box);
// same for fp32 tensors (box inner extent <= 32 elements = one 128-byte swizzle row)
int make_tmap_f32(dicow_ctx* ctx, CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                  const uint64_t* strides_bytes, const uint32_t* box);

int mel_tables_create(dicow_ctx* ctx);  // mel.cu

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Per-DEVICE once-flags / high-water marks for cudaFuncSetAttribute (opt-in shared memory): a function attribute belongs to the
// device that was current when it was set, so a process that drives several devices (one handle each) must set it on each.
// Usage:  static DeviceOnce once;  if (once.first(ctx)) cudaFuncSetAttribute(...);
constexpr int kMaxDevices = 64;
struct DeviceOnce {
  bool done[kMaxDevices] = {};
  bool first(const dicow_ctx* ctx) {
    const int d = ctx->device >= 0 && ctx->device < kMaxDevices ? ctx->device : 0;
    if (done[d]) return false;
    done[d] = true;
    return true;
  }
};
struct DeviceHighWater {
  size_t mark[kMaxDevices] = {};
  bool raise(const dicow_ctx* ctx, size_t v) {
    const int d = ctx->device >= 0 && ctx->device < kMaxDevices ? ctx->device : 0;
    if (v <= mark[d]) return false;
    mark[d] = v;
    return true;
  }
};

// Launch of a decode-step kernel: optional cluster dimension, and the programmatic-dependent-launch attribute (ptx.cuh:
// griddep_wait / griddep_launch) unless DICOW_PDL=0.
inline bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("DICOW_PDL");
    return e != nullptr && e[0] == '1';
  }();
  return on;
}

template <typename... KArgs, typename... Args>
cudaError_t launch_step_kernel(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                               unsigned cluster_x, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = stream;
  cudaLaunchAttribute attrs[2];
  unsigned n = 0;
  if (cluster_x > 1) {
    attrs[n].id = cudaLaunchAttributeClusterDimension;
    attrs[n].val.clusterDim.x = cluster_x, attrs[n].val.clusterDim.y = 1, attrs[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl_enabled()) {
    attrs[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attrs[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attrs, cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

}  // namespace dicow
