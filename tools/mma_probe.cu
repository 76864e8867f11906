// mma_probe.cu -- tcgen05.mma issue-rate microbenchmark for the shapes of the attention kernels (debug aid, not product).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I ts-asr-whisper_b200/csrc tools/mma_probe.cu -o gpurun_out/mma_probe
// One CTA per SM, one issuing thread; each configuration issues REPS batches of MMAs (a batch = one accumulation chain over
// k-steps into one TMEM accumulator, accumulators rotate over NBUF buffers), commits, waits, and reports SM clocks per MMA
// instruction against the tensor-pipe floor M*N/256... (128 x N x 16: N/2 clk).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#include "ptx.cuh"

using namespace dicow;

template <int A_TMEM, int N, int B_MN, int KSTEPS, int NBUF, int MIX, int LD>
__global__ void __launch_bounds__(384, 1) probe(int reps, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
    stop = 0;
  }
  if (warp == 9) {
    tmem_alloc(&slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 9) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(128, N, 0, B_MN);
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);
      const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem + 32768);
      const uint64_t da0 = make_sdesc_sw128(a_addr, 1024, 0);
      const uint64_t db0 = B_MN ? make_sdesc_sw128(b_addr, 1024, 8192) : make_sdesc_sw128(b_addr, 1024, 0);
      const uint64_t ds0 = make_sdesc_sw128(b_addr + 16384, 1024, 0);
      long long t0 = clock64();
      uint32_t buf = 0;
#pragma unroll 1
      for (int r = 0; r < reps; ++r) {
        const uint32_t d = tmem + buf * N;
        buf = (buf + 1 == NBUF) ? 0 : buf + 1;
#pragma unroll
        for (int k = 0; k < KSTEPS; ++k) {
          const uint64_t db = B_MN ? db0 + (uint64_t)((k % 8) * 128) : db0 + (uint64_t)((k % 4) * 2);
          if (A_TMEM) {
            umma_bf16_ts(d, tmem + 384 + (k % 8) * 8, db, idesc, k != 0);
          } else {
            umma_bf16_ss(d, da0 + (uint64_t)((k % 4) * 2), db, idesc, k != 0);
          }
        }
        if (MIX) {  // an S-like batch: SS 128 x 128, 4 k-steps, into the accumulators at columns [128, 384)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_ss(tmem + 128 + (r & 1) * 128, da0 + (uint64_t)(k * 2), ds0 + (uint64_t)(k * 2), idesc_s, k != 0);
        }
      }
      umma_commit(&bar);
      mbar_wait(&bar, 0);
      long long t1 = clock64();
      stop = 1;
      if (blockIdx.x == 0) {
        out[0] = t1 - t0;
        out[1] = (long long)reps * KSTEPS;
      }
    }
    __syncwarp();
  } else if (warp < 8 && LD >= 2) {
    // softmax-like ARITHMETIC on the two warps of every SM sub-partition: exp2 on the SFU + packed FMAs, no memory
    float2 a = make_float2(threadIdx.x * 1e-3f, 0.5f), b = make_float2(0.25f, 0.125f);
    float acc = 0.f;
    long long n = 0;
    while (!stop) {
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        a = fma_f32x2(a, make_float2(0.999f, 0.999f), b);
        b = add_f32x2(b, a);
        acc += fast_exp2(a.x) + fast_exp2(b.y);
      }
      ++n;
    }
    if (acc == 1234.5f) out[3] = n;
    if (blockIdx.x == 0 && threadIdx.x == 0) out[2] = n;
  } else if (warp < 4 && LD == 1) {
    // softmax-like TMEM traffic: read 128 columns, write 64 back, until the MMAs are done
    const uint32_t la = tmem + (uint32_t(warp * 32) << 16) + 256;
    uint32_t r[32];
    long long n = 0;
    while (!stop) {
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        tmem_ld_x32(la + cc * 32, r);
        tmem_ld_wait_regs(r);
      }
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) pk[i] = r[i] ^ r[i + 16];
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) tmem_st_x16(la + cc * 16, pk);
      tmem_st_wait();
      ++n;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) out[2] = n;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

template <int A_TMEM, int N, int B_MN, int KSTEPS, int NBUF, int MIX, int LD>
void run(const char* name, long long* out) {
  auto k = probe<A_TMEM, N, B_MN, KSTEPS, NBUF, MIX, LD>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  for (int it = 0; it < 2; ++it) {
    k<<<148, 384, 100 * 1024>>>(4000, out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("%s: %s\n", name, cudaGetErrorString(e));
      exit(1);
    }
  }
  const double per = double(out[0]) / double(out[1]);
  const double extra = MIX ? 4.0 / KSTEPS : 0.0;  // S-like instructions issued per counted instruction
  const double floor_clk = N / 2.0 + extra * 64.0;
  printf("%-52s: %7.1f clk / counted MMA (tensor floor %.0f) -> %3.0f %% of the floor rate; ld/st loops %lld\n", name, per,
         floor_clk, 100.0 * floor_clk / per, LD ? out[2] : 0LL);
}

int main() {
  long long* out;
  cudaMallocManaged(&out, 64);
  run<0, 256, 0, 4, 2, 0, 0>("SS N=256 K-major, chain 4, 2 acc", out);
  run<0, 256, 0, 16, 2, 0, 0>("SS N=256 K-major, chain 16, 2 acc", out);
  run<0, 128, 0, 4, 3, 0, 0>("SS N=128 K-major, chain 4, 3 acc (S = QK^T)", out);
  run<0, 128, 0, 1, 3, 0, 0>("SS N=128 K-major, chain 1 (independent)", out);
  run<0, 128, 0, 16, 1, 0, 0>("SS N=128 K-major, chain 16, 1 acc", out);
  run<0, 64, 0, 8, 2, 0, 0>("SS N=64 K-major, chain 8", out);
  run<0, 64, 1, 8, 2, 0, 0>("SS N=64 MN-major B, chain 8", out);
  run<1, 64, 1, 8, 2, 0, 0>("TS N=64 MN-major B, chain 8 (O += P V)", out);
  run<1, 64, 0, 8, 2, 0, 0>("TS N=64 K-major B, chain 8", out);
  run<1, 128, 1, 8, 2, 0, 0>("TS N=128 MN-major B, chain 8", out);
  run<1, 128, 0, 4, 2, 0, 0>("TS N=128 K-major B, chain 4 (S with Q in TMEM)", out);
  run<1, 256, 0, 4, 1, 0, 0>("TS N=256 K-major B, chain 4", out);
  run<1, 64, 1, 8, 2, 1, 0>("TS N=64 MN chain 8 + SS N=128 chain 4 (PV, S)", out);
  run<1, 64, 1, 8, 2, 1, 1>("same, with 4 warps doing tcgen05.ld/st", out);
  run<0, 128, 0, 4, 3, 0, 1>("SS N=128 chain 4, with 4 warps ld/st", out);
  run<0, 128, 0, 4, 3, 0, 2>("SS N=128 chain 4, with 8 warps of exp2/FMA arithmetic", out);
  run<1, 64, 1, 8, 2, 1, 2>("PV + S mix, with 8 warps of exp2/FMA arithmetic", out);
  return 0;
}
