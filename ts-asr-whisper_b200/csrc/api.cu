// api.cu -- handle lifetime, error plumbing, TMA tensor-map encoding (host side of the C ABI).
#include <stdarg.h>

#include <new>

#include "common.h"

namespace dicow {

int set_error(dicow_ctx* ctx, int code, const char* fmt, ...) {
  if (ctx != nullptr) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(ctx->err, sizeof(ctx->err), fmt, ap);
    va_end(ap);
  }
  return code;
}

typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_tmap(dicow_ctx* ctx, CUtensorMap* out, CUtensorMapDataType dtype, const void* base, int rank,
                     const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box);

int make_tmap_bf16(dicow_ctx* ctx, CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box) {
  return make_tmap(ctx, out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, base, rank, dims, strides_bytes, box);
}

int make_tmap_f32(dicow_ctx* ctx, CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                  const uint64_t* strides_bytes, const uint32_t* box) {
  return make_tmap(ctx, out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, base, rank, dims, strides_bytes, box);
}

static int make_tmap(dicow_ctx* ctx, CUtensorMap* out, CUtensorMapDataType dtype, const void* base, int rank,
                     const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box) {
  if (ctx->encode_tiled == nullptr)
    return set_error(ctx, DICOW_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
  }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  CUresult r = reinterpret_cast<encode_tiled_fn>(ctx->encode_tiled)(
      out, dtype, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es,
      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    return set_error(ctx, DICOW_ERR_CUDA,
                     "cuTensorMapEncodeTiled failed (CUresult %d): rank %d dims [%llu,%llu,%llu] strides [%llu,%llu] "
                     "box [%u,%u,%u] base %p",
                     (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
                     (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 1 ? strides_bytes[0] : 0),
                     (unsigned long long)(rank > 2 ? strides_bytes[1] : 0), box[0], rank > 1 ? box[1] : 0,
                     rank > 2 ? box[2] : 0, base);
  }
  return DICOW_OK;
}

}  // namespace dicow

using namespace dicow;

// 2: round 2 -- dicow_attention_bwd_args_t.workspace_floats, dicow_fddt_ln_args_t.x_out, dicow_ln_bwd_args_t.g_colsum,
//    dicow_ctc_loss_args_t.ld / dicow_ctc_bwd_args_t.ld, dicow_set_sm_budget
extern "C" int dicow_abi_version(void) { return 2; }

extern "C" int dicow_create(int device, dicow_handle_t* out) {
  if (out == nullptr) return DICOW_ERR_INVALID_ARG;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return DICOW_ERR_CUDA;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return DICOW_ERR_CUDA;
  if (prop.major != 10) return DICOW_ERR_ARCH;  // kernels are sm_100a only: fail loudly, no fallback
  if (cudaSetDevice(device) != cudaSuccess) return DICOW_ERR_CUDA;
  dicow_ctx* ctx = new (std::nothrow) dicow_ctx();
  if (ctx == nullptr) return DICOW_ERR_CUDA;
  ctx->device = device;
  ctx->num_sms = ctx->phys_sms = prop.multiProcessorCount;
  ctx->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess || fn == nullptr) {
    delete ctx;
    return DICOW_ERR_CUDA;
  }
  ctx->encode_tiled = fn;
  if (mel_tables_create(ctx) != DICOW_OK) {
    delete ctx;
    return DICOW_ERR_CUDA;
  }
  *out = ctx;
  return DICOW_OK;
}

extern "C" int dicow_destroy(dicow_handle_t h) {
  if (h != nullptr && h->mel_tables != nullptr) cudaFree(h->mel_tables);
  delete h;
  return DICOW_OK;
}

extern "C" int dicow_set_sm_budget(dicow_handle_t h, int sms) {
  if (h == nullptr) return DICOW_ERR_INVALID_ARG;
  if (sms <= 0 || sms >= h->phys_sms) {
    h->num_sms = h->phys_sms;
  } else {
    h->num_sms = sms < 2 ? 2 : (sms & ~1);  // CTA pairs: an even number
  }
  return h->num_sms;
}

extern "C" const char* dicow_last_error(dicow_handle_t h) { return h ? h->err : "null handle"; }

extern "C" int dicow_check(dicow_handle_t h) {
  if (h == nullptr) return DICOW_ERR_INVALID_ARG;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(h, DICOW_ERR_CUDA, "pending CUDA error: %s", cudaGetErrorString(e));
  return DICOW_OK;
}
