// Micro-benchmark: how fast can one SM gather rows of an fp16 activation matrix into shared memory
//   (a) through TMA tile::gather4 (cp.async.bulk.tensor.2d ... tile::gather4: four rows by index per instruction), and
//   (b) through the 16-byte cp.async (LDGSTS) copies the convolution kernel's producers issue today,
// for rows of 128 / 64 / 32 bytes, with a fraction of the slots absent (-1 -> TMA out-of-bounds zero fill / zero-size copy).
// One 128-row stage per iteration, ring of 4 stages, persistent CTAs.  Prints cycles per stage per CTA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/gather4_probe tools/gather4_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(bar),
               "r"(parity)
               : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void gather4(uint32_t dst, const CUtensorMap* map, int col, int r0, int r1, int r2, int r3, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(dst),
               "l"(reinterpret_cast<uint64_t>(map)), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar)
               : "memory");
}

constexpr int kStages = 4;
constexpr int kRows = 128;

// (a) TMA gather4: `nissue` threads issue the 32 instructions of a stage (each thread 32 / nissue of them); nissue <= 32: lanes of
// warp 0; nissue = 100 + w: lane 0 of w warps; nissue = 200 + w: 32 / w lanes of each of w warps
__global__ void __launch_bounds__(256) k_gather4(const __grid_constant__ CUtensorMap tmap, const int32_t* __restrict__ idx, int iters,
                                                 int row_bytes, int nissue, long long* cycles, uint32_t* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[kStages];
  const int tid = threadIdx.x;
  const uint32_t bar0 = smem_u32(bars), s0 = smem_u32(smem);
  const uint32_t stage_bytes = (uint32_t)kRows * row_bytes;
  const uint32_t q_step = 4 * row_bytes < 128 ? 128 : 4 * row_bytes;   // TMA destinations are 128-byte aligned
  const uint32_t stage_step = 32 * q_step;
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(bar0 + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int32_t* my = idx + (size_t)blockIdx.x * iters * kRows;
  long long t0 = clock64();
  const int warp = tid >> 5, lane = tid & 31;
  int nw = 1, nl = nissue;
  if (nissue >= 200) { nw = nissue - 200; nl = 32 / nw; }
  else if (nissue >= 100) { nw = nissue - 100; nl = 1; }
  const int ni = nw * nl;
  if (warp < nw) {
    for (int it = 0; it < iters; ++it) {
      const int s = it % kStages;
      if (it >= kStages) mbar_wait(bar0 + 8 * s, ((it / kStages) - 1) & 1);
      if (tid == 0) mbar_expect_tx(bar0 + 8 * s, stage_bytes);
      __syncwarp();
      if (lane < nl) {
        for (int q = warp * nl + lane; q < 32; q += ni) {
          const int4 r = __ldg(reinterpret_cast<const int4*>(my + (size_t)it * kRows) + q);
          gather4(s0 + s * stage_step + q * q_step, &tmap, 0, r.x, r.y, r.z, r.w, bar0 + 8 * s);
        }
      }
    }
    for (int it = iters; it < iters + kStages && it - kStages >= 0; ++it) mbar_wait(bar0 + 8 * (it % kStages), ((it / kStages) - 1) & 1);
  }
  __syncthreads();
  long long t1 = clock64();
  if (tid == 0) { cycles[blockIdx.x] = t1 - t0; sink[blockIdx.x] = smem[(t1 & 1023)]; }
}

// (b) the producers of the convolution kernel: 256 threads, each 4 rows x one 16-byte column (or as many as the row has)
__global__ void __launch_bounds__(256) k_ldgsts(const uint8_t* __restrict__ base, int64_t ld_bytes, const int32_t* __restrict__ idx, int iters,
                                                int row_bytes, int sts_zero, long long* cycles, uint32_t* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[kStages];
  const int tid = threadIdx.x;
  const uint32_t bar0 = smem_u32(bars), s0 = smem_u32(smem);
  const int cols = row_bytes / 16;                      // 16-byte columns per row
  const uint32_t stage_bytes = (uint32_t)kRows * 128;   // rows of 128 bytes in the stage (narrow rows: several offsets per stage)
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(bar0 + 8 * s, 256);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int32_t* my = idx + (size_t)blockIdx.x * iters * kRows;
  const int r0 = tid >> 3, cB = tid & 7;
  long long t0 = clock64();
  // a stage = 128 rows x 8 columns; with narrow rows the 8 columns are 8 / cols different offsets: every offset has its own index
  // vector (we reuse the iteration's vector rotated -- what matters is the number of distinct rows per instruction)
  for (int it = 0; it < iters; ++it) {
    const int s = it % kStages;
    if (it >= kStages) mbar_wait(bar0 + 8 * s, ((it / kStages) - 1) & 1);
    const int e = cB / cols;
    const int4 r = __ldg(reinterpret_cast<const int4*>(my + (size_t)it * kRows) + ((r0 + 7 * e) & 31));
    const int rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = 4 * r0 + i;
      const uint32_t dst = s0 + s * stage_bytes + (uint32_t)((row >> 3) * 1024 + (row & 7) * 128) + (((uint32_t)cB ^ (uint32_t)(row & 7)) << 4);
      const bool ok = rr[i] >= 0;
      const uint8_t* src = base + (int64_t)(ok ? rr[i] : 0) * ld_bytes + (cB % cols) * 16;
      if (!sts_zero) asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(ok ? 16 : 0) : "memory");
      else if (ok) asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
      else asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(dst), "r"(0) : "memory");   // absent slot: plain zero store
    }
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar0 + 8 * s) : "memory");
  }
  for (int it = iters; it < iters + kStages && it - kStages >= 0; ++it) mbar_wait(bar0 + 8 * (it % kStages), ((it / kStages) - 1) & 1);
  __syncthreads();
  long long t1 = clock64();
  if (tid == 0) { cycles[blockIdx.x] = t1 - t0; sink[blockIdx.x] = smem[(t1 & 1023)]; }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const int64_t R = argc > 1 ? atoll(argv[1]) : (1 << 20);   // rows of the activation matrix (1 M rows x 128 B = 128 MB > L2)
  const int zero_row = argc > 2 ? atoi(argv[2]) : 0;          // 1: absent slots read one fixed (zero) row instead of going out of bounds
  const int iters = 400;
  void* f = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q));
  EncodeTiledFn enc = (EncodeTiledFn)f;
  uint8_t* act;
  CK(cudaMalloc(&act, R * 128));
  CK(cudaMemset(act, 1, R * 128));
  long long* d_cyc; uint32_t* d_sink;
  CK(cudaMalloc(&d_cyc, 1024 * 8));
  CK(cudaMalloc(&d_sink, 1024 * 4));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  printf("path,row_bytes,absent_frac,locality,ctas_per_sm,issuers,cycles_per_stage_per_cta,us_total,rows_per_us_chip,GBs_chip\n");
  for (int row_bytes : {128, 64, 32}) {
    for (double absent : {0.0, 0.6}) {
      for (int local : {0}) {
        for (int per_sm : {1, 2}) {
          const int grid = 148 * per_sm;
          // indices: random rows (local = 0) or rows near a random centre per stage (local = 1: +-2048 rows, like neighbours in a sorted level)
          std::vector<int32_t> h((size_t)grid * iters * kRows);
          uint64_t sd = 88172645463325252ull;
          auto rnd = [&]() { sd ^= sd << 13; sd ^= sd >> 7; sd ^= sd << 17; return sd; };
          for (size_t st = 0; st < (size_t)grid * iters; ++st) {
            const int64_t centre = rnd() % R;
            for (int i = 0; i < kRows; ++i) {
              int64_t r = local ? (centre + (int64_t)(rnd() % 4096) - 2048 + R) % R : (int64_t)(rnd() % R);
              if ((rnd() % 1000) < absent * 1000) r = zero_row ? R - 1 : -1;
              h[st * kRows + i] = (int32_t)r;
            }
          }
          int32_t* d_idx;
          CK(cudaMalloc(&d_idx, h.size() * 4));
          CK(cudaMemcpy(d_idx, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
          const int64_t real = (int64_t)((double)grid * iters * kRows * (1.0 - absent));
          // ---- TMA gather4 ----
          CUtensorMap tm;
          const cuuint64_t dims[2] = {(cuuint64_t)(row_bytes / 2), (cuuint64_t)R};
          const cuuint64_t strides[1] = {(cuuint64_t)row_bytes};
          const cuuint32_t box[2] = {(cuuint32_t)(row_bytes / 2), 1};
          const cuuint32_t estr[2] = {1, 1};
          const CUtensorMapSwizzle sw = row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                        : row_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
          CUresult cr = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, act, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                            CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
          if (cr != CUDA_SUCCESS) { printf("encode failed %d (row_bytes %d)\n", (int)cr, row_bytes); }
          else {
            for (int nissue : {208}) {
              const size_t smem = (size_t)kStages * 32 * (4 * row_bytes < 128 ? 128 : 4 * row_bytes) + 1024;
              CK(cudaFuncSetAttribute(k_gather4, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
              k_gather4<<<grid, 256, smem>>>(tm, d_idx, iters, row_bytes, nissue, d_cyc, d_sink);   // warm
              CK(cudaDeviceSynchronize());
              cudaEventRecord(e0);
              k_gather4<<<grid, 256, smem>>>(tm, d_idx, iters, row_bytes, nissue, d_cyc, d_sink);
              cudaEventRecord(e1);
              CK(cudaDeviceSynchronize());
              float ms; cudaEventElapsedTime(&ms, e0, e1);
              std::vector<long long> cyc(grid);
              CK(cudaMemcpy(cyc.data(), d_cyc, grid * 8, cudaMemcpyDeviceToHost));
              double avg = 0; for (auto c : cyc) avg += (double)c; avg /= grid;
              printf("gather4,%d,%.1f,%d,%d,%d,%.0f,%.1f,%.0f,%.0f\n", row_bytes, absent, local, per_sm, nissue, avg / iters, ms * 1e3,
                     real / (ms * 1e3), real * (double)row_bytes / (ms * 1e6));
            }
          }
          // ---- LDGSTS ----
          for (int sts_zero : {0, 1}) {
            const size_t smem = (size_t)kStages * kRows * 128 + 1024;
            CK(cudaFuncSetAttribute(k_ldgsts, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_ldgsts<<<grid, 256, smem>>>(act, row_bytes, d_idx, iters, row_bytes, sts_zero, d_cyc, d_sink);
            CK(cudaDeviceSynchronize());
            cudaEventRecord(e0);
            k_ldgsts<<<grid, 256, smem>>>(act, row_bytes, d_idx, iters, row_bytes, sts_zero, d_cyc, d_sink);
            cudaEventRecord(e1);
            CK(cudaDeviceSynchronize());
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            std::vector<long long> cyc(grid);
            CK(cudaMemcpy(cyc.data(), d_cyc, grid * 8, cudaMemcpyDeviceToHost));
            double avg = 0; for (auto c : cyc) avg += (double)c; avg /= grid;
            // a stage of the narrow-row variant carries 128 / row_bytes offsets: rows gathered per stage = 128 * (128 / row_bytes)
            const double rows_stage_scale = 128.0 / row_bytes;
            printf("%s,%d,%.1f,%d,%d,%d,%.0f,%.1f,%.0f,%.0f\n", sts_zero ? "ldgsts+sts0" : "ldgsts", row_bytes, absent, local, per_sm, 256, avg / iters, ms * 1e3,
                   real * rows_stage_scale / (ms * 1e3), real * rows_stage_scale * row_bytes / (ms * 1e6));
          }
          CK(cudaFree(d_idx));
        }
      }
    }
  }
  return 0;
}
