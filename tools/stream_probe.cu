// Micro-benchmark: how fast can 148 CTAs stream a weight matrix from HBM into shared memory, as a function of the access
// pattern? (a) the K-major GEMM operand pattern: 128-row x 64-column (128-byte) boxes of a row-major [F][K] matrix through a
// tensor map, k fastest; (b) the same bytes as contiguous 16 KiB blocks (what a pre-tiled weight layout would give);
// (c) pattern (a) plus a second, L2-resident 16 KiB "activation" tile per k-block (what the decode products load).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/stream_probe tools/stream_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                   smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
                   smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

struct Args {
  int F, K, stages, mode;   // mode 0: tensor boxes; 1: contiguous blocks; 2: tensor boxes + activation tile; 3: blocks + activation tile
  int box_rows;             // rows per box (128 or 256)
  const uint8_t* w;
};

__global__ void __launch_bounds__(128, 1) stream_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmX,
                                                        const Args a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int wbytes = a.box_rows * 128;
  const int stage_bytes = wbytes + ((a.mode >= 2) ? 16384 : 0);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)a.stages * stage_bytes);
  uint64_t* empty = full + 16;
  if (threadIdx.x == 0)
    for (int i = 0; i < a.stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  const int nkb = a.K / 64, tiles = a.F / a.box_rows;
  const int units = tiles * nkb;
  const int per = (units + gridDim.x - 1) / gridDim.x;
  const int u0 = min(units, (int)blockIdx.x * per), u1 = min(units, u0 + per);
  if (threadIdx.x == 0) {
    for (int u = u0, i = 0; u < u1; ++u, ++i) {
      const int s = i % a.stages;
      mbar_wait(&empty[s], ((i / a.stages) & 1) ^ 1);
      mbar_arrive_expect_tx(&full[s], stage_bytes);
      const int tile = u / nkb, kb = u % nkb;
      uint8_t* dst = smem + (size_t)s * stage_bytes;
      if (a.mode == 0 || a.mode == 2) tma_load_2d(dst, &tmW, &full[s], kb * 64, tile * a.box_rows);
      else bulk_copy_g2s(dst, a.w + (size_t)u * wbytes, wbytes, &full[s]);
      if (a.mode >= 2) tma_load_2d(dst + wbytes, &tmX, &full[s], kb * 64, 0);
    }
  } else if (threadIdx.x == 32) {
    for (int u = u0, i = 0; u < u1; ++u, ++i) {
      const int s = i % a.stages;
      mbar_wait(&full[s], (i / a.stages) & 1);
      mbar_arrive(&empty[s]);
    }
  }
}

// gate_up emulation: CTA c streams rows [80c, 80c+80) and [I + 80c, ...) of a [2I][K] matrix, all 32 k-blocks, plus the xn tile;
// stagger != 0: CTA c starts at k-block (c * stagger) % nkb and wraps around
struct GuArgs { int I, K, stages, stagger, ft, with_x, rank4, no_w, spin; };
__global__ void __launch_bounds__(256, 1) gu_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmX,
                                                    const GuArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int wbytes = 2 * a.ft * 128;
  const int stage_bytes = ((wbytes + 1023) & ~1023) + 16384;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)a.stages * stage_bytes);
  uint64_t* empty = full + 16;
  if (threadIdx.x == 0)
    for (int i = 0; i < a.stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  const int nkb = a.K / 64, tiles = (a.I + a.ft - 1) / a.ft;
  uint64_t* done = empty + 16;
  if (threadIdx.x == 0) mbar_init(done, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  if ((int)blockIdx.x >= tiles) return;
  if (threadIdx.x >= 128) {
    if (a.spin) mbar_wait(done, 0);      // 128 threads polling try_wait for the whole k-loop
    return;
  }
  const int f0 = blockIdx.x * a.ft, k0 = (blockIdx.x * a.stagger) % nkb;
  if (threadIdx.x == 0) {
    for (int i = 0; i < nkb; ++i) {
      const int s = i % a.stages, kb = (k0 + i) % nkb;
      mbar_wait(&empty[s], ((i / a.stages) & 1) ^ 1);
      mbar_arrive_expect_tx(&full[s], (a.no_w ? 0 : wbytes) + (a.with_x ? 16384 : 0));
      uint8_t* dst = smem + (size_t)s * stage_bytes;
      if (a.no_w) {
        if (a.rank4) tma_load_4d(dst, &tmX, &full[s], kb * 64, 0, 0, 0);
        else tma_load_2d(dst, &tmX, &full[s], kb * 64, 0);
      } else if (a.rank4) {
        tma_load_4d(dst + 16384, &tmW, &full[s], kb * 64, f0, 0, 0);
        tma_load_4d(dst + 16384 + a.ft * 128, &tmW, &full[s], kb * 64, a.I + f0, 0, 0);
        if (a.with_x) tma_load_4d(dst, &tmX, &full[s], kb * 64, 0, 0, 0);
      } else {
        tma_load_2d(dst + 16384, &tmW, &full[s], kb * 64, f0);
        tma_load_2d(dst + 16384 + a.ft * 128, &tmW, &full[s], kb * 64, a.I + f0);
        if (a.with_x) tma_load_2d(dst, &tmX, &full[s], kb * 64, 0);
      }
    }
  } else if (threadIdx.x == 32) {
    for (int i = 0; i < nkb; ++i) {
      const int s = i % a.stages;
      mbar_wait(&full[s], (i / a.stages) & 1);
      mbar_arrive(&empty[s]);
    }
  }
}

__global__ void touch_kernel(uint4* x, int n, unsigned v) {   // every SM rewrites a slice of x (as the producing phase does)
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) x[i] = make_uint4(v, v + i, v, v);
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const int F = 22016 * 8, K = 2048, R = 128;   // 721 MB: long enough that launch overhead does not matter
  uint8_t *w, *x;
  cudaMalloc(&w, (size_t)F * K * 2);
  cudaMalloc(&x, (size_t)R * K * 2);
  cudaMemset(w, 1, (size_t)F * K * 2);
  cudaMemset(x, 1, (size_t)R * K * 2);
  uint8_t* flush;
  cudaMalloc(&flush, 512u << 20);
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  EncodeFn enc = reinterpret_cast<EncodeFn>(fp);
  auto make = [&](CUtensorMap* m, void* p, int rows, int cols, int box_rows, CUtensorMapL2promotion promo) {
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, p, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
  };
  auto make4 = [&](CUtensorMap* m, void* p, int rows, int cols, int box_rows) {
    cuuint64_t dims[4] = {(cuuint64_t)cols, (cuuint64_t)rows, 1, 1};
    cuuint64_t strides[3] = {(cuuint64_t)cols * 2, (cuuint64_t)cols * 2, (cuuint64_t)cols * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)box_rows, 1, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, p, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode4 failed %d\n", (int)r); exit(1); }
  };
  cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const char* names[4] = {"tensor boxes [F][K], 128 B wide", "contiguous blocks", "tensor boxes + L2 activation tile", "blocks + L2 activation tile"};
  for (int promo = 1; promo < 1; ++promo)
    for (int box_rows : {128})
      for (int mode = 0; mode < 4; ++mode)
        for (int stages : {3, 5, 8, 12}) {
          const int stage_bytes = box_rows * 128 + (mode >= 2 ? 16384 : 0);
          if ((size_t)stages * stage_bytes + 2048 > 227 * 1024) continue;
          CUtensorMap tmW, tmX;
          make(&tmW, w, F, K, box_rows, promo ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
          make(&tmX, x, R, K, 128, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
          Args a{F, K, stages, mode, box_rows, w};
          float best = 1e9;
          for (int rep = 0; rep < 5; ++rep) {
            cudaMemsetAsync(flush, rep, 512u << 20);   // evict the weights from L2
            cudaEventRecord(e0);
            stream_kernel<<<148, 128, (size_t)stages * stage_bytes + 2048>>>(tmW, tmX, a);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) best = ms;
          }
          cudaError_t e = cudaGetLastError();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          printf("promo %s box_rows %3d stages %2d  %-36s %7.1f us  %6.2f TB/s (weights only)\n", promo ? "256B" : "128B", box_rows, stages,
                 names[mode], best * 1e3, (double)F * K * 2 / (best * 1e-3) / 1e12);
        }
  // ---- the decode case: 90 MB matrices streamed once each, back to back (20 different matrices = 1.8 GB, cold in L2 and TLB)
  {
    const int F1 = 22016;
    CUtensorMap tmX;
    make(&tmX, x, R, K, 128, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
    for (int mode : {0, 2})
      for (int stages : {3, 5, 8}) {
        std::vector<CUtensorMap> maps(8);
        for (int i = 0; i < 8; ++i) make(&maps[i], w + (size_t)i * F1 * K * 2, F1, K, 128, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
        const int stage_bytes = 16384 + (mode >= 2 ? 16384 : 0);
        float best = 1e9;
        for (int rep = 0; rep < 4; ++rep) {
          cudaMemsetAsync(flush, rep, 512u << 20);
          cudaEventRecord(e0);
          for (int i = 0; i < 8; ++i) {
            Args a{F1, K, stages, mode, 128, w + (size_t)i * F1 * K * 2};
            stream_kernel<<<148, 128, (size_t)stages * stage_bytes + 2048>>>(maps[i], tmX, a);
          }
          cudaEventRecord(e1);
          cudaEventSynchronize(e1);
          float ms;
          cudaEventElapsedTime(&ms, e0, e1);
          if (ms < best) best = ms;
        }
        printf("8 x 90 MB back-to-back launches, mode %d stages %d: %7.1f us per launch, %6.2f TB/s\n", mode, stages, best * 1e3 / 8,
               (double)F1 * K * 2 * 8 / (best * 1e-3) / 1e12);
      }
  }
  {
    const int I = 11008, ft = 80;
    CUtensorMap tmX;
    make(&tmX, x, R, K, 128, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
    cudaFuncSetAttribute(gu_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    for (int rank4 : {1, 3, 9, 11})
      for (int stagger : {0})
        for (int stages : {5}) {
          const int with_x = 1, no_w = (rank4 >> 1) & 1, touch = (rank4 >> 2) & 1;
          std::vector<CUtensorMap> maps(8);
          for (int i = 0; i < 8; ++i) {
            if (rank4) make4(&maps[i], w + (size_t)i * 2 * I * K * 2, 2 * I, K, ft);
            else make(&maps[i], w + (size_t)i * 2 * I * K * 2, 2 * I, K, ft, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
          }
          if (rank4) make4(&tmX, x, R, K, 128);
          const int stage_bytes = ((2 * ft * 128 + 1023) & ~1023) + 16384;
          float best = 1e9;
          for (int rep = 0; rep < 4; ++rep) {
            cudaMemsetAsync(flush, rep, 512u << 20);
            cudaEventRecord(e0);
            for (int i = 0; i < 8; ++i) {
              GuArgs a{I, K, stages, stagger, ft, with_x, rank4 & 1, no_w, (rank4 >> 3) & 1};
              if (touch) touch_kernel<<<148, 256>>>(reinterpret_cast<uint4*>(x), R * K * 2 / 16, (unsigned)(rep * 8 + i));
              gu_kernel<<<148, 256, (size_t)stages * stage_bytes + 2048>>>(maps[i], tmX, a);
            }
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) best = ms;
          }
          cudaError_t e = cudaGetLastError();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          printf("gate_up emulation (138 CTAs x one 80+80-row tile x 32 k-blocks), flags(1 rank4, 2 acts only, 4 acts rewritten first, 8 four warps spin on try_wait) %d stagger %d stages %d: %7.1f us per launch, %6.2f TB/s\n",
                 rank4, stagger, stages, best * 1e3 / 8, (double)2 * I * K * 2 * 8 / (best * 1e-3) / 1e12);
        }
  }
  return 0;
}
