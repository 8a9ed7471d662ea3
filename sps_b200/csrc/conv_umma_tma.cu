// tcgen05 implicit-GEMM sparse convolution, wide-channel variant (Cin >= 24): the gather itself
// runs on the TMA engine.
//
// Same tiling, stage plan and epilogue as conv_umma.cu, but the A operand of a stage (128 gathered
// neighbour rows x 32 channels = 128 B each) is fetched by 32 `cp.async.bulk.tensor.2d ...
// tile::gather4` instructions -- one per lane of a single producer warp, four neighbour rows each,
// row indices straight from the staged kernel-map slice (absent neighbour = index -1 = out of
// bounds = hardware zero fill) -- and lands 128B-swizzled exactly where the UMMA descriptor expects
// it.  The weight stage and the fused 1x1 term are plain 2-D TMA tiles.  The SM's threads no
// longer compute a single gather address: one producer warp issues ~12 instructions per stage
// instead of the ~400 of the cp.async variant, which was instruction-issue bound (profiles/).
//   warp 0: TMA producer (kernel-map slice staging, gather4, weight tiles)
//   warp 1: tcgen05.mma issuer            warps 2-5: epilogue (TMEM quadrant = warp & 3)
#include <cuda.h>
#include "umma_common.cuh"

namespace sps {

constexpr int kTmaThreads = 192;

template <int NPAD>
struct TmaCfg {
#ifndef SPS_TMA_S
#define SPS_TMA_S 5
#endif
  static constexpr int S = SPS_TMA_S;
  static constexpr int kBStage = NPAD * 128;
  static constexpr int kTmemCols = 2 * NPAD < 32 ? 32 : 2 * NPAD;
  static constexpr size_t smem = (size_t)S * (kAStageBytes + kBStage) + 2 * (size_t)kMaxK * kTileM * 4 +
                                 8 * (2 * S + 6) + 2 * 96 + 32;
};

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_gather4(uint32_t dst, const CUtensorMap* tm, int col, int r0, int r1, int r2, int r3,
                                            uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, "
      "%5, %6}], [%7];" ::"r"(dst),
      "l"(tm), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tma_tile2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          dst),
      "l"(tm), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}

template <int NPAD>
__global__ void __launch_bounds__(kTmaThreads, 1)
k_conv_umma_tma(const sps_conv_args a, const UmmaParams p, const __grid_constant__ CUtensorMap tmA,
                const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB) {
  using Cfg = TmaCfg<NPAD>;
  constexpr int S = Cfg::S;
  constexpr int kBStageBytes = Cfg::kBStage;
  constexpr uint32_t kStageTx = kAStageBytes + kBStageBytes;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sB = smem + S * kAStageBytes;
  int32_t* sidx = reinterpret_cast<int32_t*>(sB + S * kBStageBytes);  // [2][K][128] (tile parity)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sidx + 2 * kMaxK * kTileM);
  uint8_t* klist = reinterpret_cast<uint8_t*>(bars + 2 * S + 6);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(klist + 2 * 96);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t sA_u = smem_u32(sA), sB_u = smem_u32(sB), sidx_u = smem_u32(sidx);
  const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 8 * S, bar_idx = bar_empty + 8 * S,
                 bar_accf = bar_idx + 16, bar_acce = bar_accf + 16;
  if (sA_u & 1023) __trap();

  if (tid == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_idx + 8 * i, 32);
      mbar_init(bar_accf + 8 * i, 1);
      mbar_init(bar_acce + 8 * i, 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)Cfg::kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int n_out = *a.n_out;
  const int ntiles = (n_out + kTileM - 1) / kTileM;
  const int K = a.K;
  const int GP = padded_groups(a.cin);   // multiple of 8 here
  const int SPE = GP >> 3;               // stages per offset
  const int gpk2 = a.in2 ? (a.cin2 >> 2) : 0;
  const int st2 = (gpk2 + 7) >> 3;
  const uint32_t* tmask = a.tile_mask;
  auto tile_nact = [&](int tile) {
    return __popc(__ldg(tmask + 4 * tile)) + __popc(__ldg(tmask + 4 * tile + 1)) + __popc(__ldg(tmask + 4 * tile + 2));
  };

  if (warp == 0) {
    // =========================== TMA PRODUCER (one warp) ===========================
    // stage the kernel-map slice of `tile` (present offsets only): lane l carries rows 4l..4l+3
    auto prepare = [&](int tile, int par) {
      int base = 0;
      for (int w = 0; w < 3; ++w) {
        const uint32_t bits = __ldg(tmask + 4 * tile + w);
        if ((bits >> lane) & 1u) klist[par * 96 + base + __popc(bits & ((1u << lane) - 1u))] = (uint8_t)(32 * w + lane);
        base += __popc(bits);
      }
      __syncwarp();
      const int row0 = tile * kTileM + 4 * lane;
      const int rem = n_out - row0;
      const uint32_t bytes = rem >= 4 ? 16u : rem > 0 ? (uint32_t)rem * 4u : 0u;   // rows past the end read as 0
      const int32_t* src = a.map + (bytes ? row0 : 0);
      const uint32_t dst = sidx_u + (uint32_t)(par * kMaxK * kTileM + 4 * lane) * 4u;
      for (int e = 0; e < base; ++e)
        cp_async16(dst + (uint32_t)(e * kTileM) * 4u, src + (int64_t)klist[par * 96 + e] * a.map_ld, bytes);
      cp_async_arrive(bar_idx + 8 * par);
    };

    uint32_t slot = 0, phase = 0;
    int it_tile = 0;
    if ((int)blockIdx.x < ntiles) prepare(blockIdx.x, 0);
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it_tile) {
      const int par = it_tile & 1;
      const int next = tile + gridDim.x;
      if (next < ntiles) prepare(next, par ^ 1);
      const int nact = tile_nact(tile);
      mbar_wait(bar_idx + 8 * par, (it_tile >> 1) & 1);
      const int4* sx = reinterpret_cast<const int4*>(sidx + par * kMaxK * kTileM) + lane;   // [e][32 lanes]
      const uint8_t* kl = klist + par * 96;
      for (int e = 0; e < nact; ++e) {
        const int4 ix = sx[e * (kTileM / 4)];
        const int kbase = (int)kl[e] * GP * 4;
        for (int sub = 0; sub < SPE; ++sub) {
          mbar_wait(bar_empty + 8 * slot, phase ^ 1);
          if (lane == 0) {
            mbar_expect_tx(bar_full + 8 * slot, kStageTx);
            tma_tile2d(sB_u + slot * kBStageBytes, &tmB, kbase + sub * 32, 0, bar_full + 8 * slot);
          }
          __syncwarp();
          tma_gather4(sA_u + slot * kAStageBytes + lane * 512, &tmA, sub * 32, ix.x, ix.y, ix.z, ix.w,
                      bar_full + 8 * slot);
          if (++slot == S) { slot = 0; phase ^= 1; }
        }
      }
      for (int s2 = 0; s2 < st2; ++s2) {   // fused 1x1 term: a plain 128-row tile of in2
        mbar_wait(bar_empty + 8 * slot, phase ^ 1);
        if (lane == 0) {
          mbar_expect_tx(bar_full + 8 * slot, kStageTx);
          tma_tile2d(sB_u + slot * kBStageBytes, &tmB, (K * GP + s2 * 8) * 4, 0, bar_full + 8 * slot);
          tma_tile2d(sA_u + slot * kAStageBytes, &tmA2, s2 * 32, tile * kTileM, bar_full + 8 * slot);
        }
        __syncwarp();
        if (++slot == S) { slot = 0; phase ^= 1; }
      }
    }
    cp_async_wait<0>();
  } else if (warp == 1) {
    // =========================== MMA ISSUER ===========================
    const uint32_t idesc = make_idesc_tf32(NPAD);
    uint32_t slot = 0, phase = 0;
    int n_acc = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int nstages = tile_nact(tile) * SPE + st2;
      if (nstages == 0) continue;
      const int b = n_acc & 1;
      mbar_wait(bar_acce + 8 * b, ((n_acc >> 1) & 1) ^ 1);
      ++n_acc;
      tc_fence_after();
      const uint32_t tacc = tmem_base + (uint32_t)(b * NPAD);
      for (int it = 0; it < nstages; ++it) {
        mbar_wait(bar_full + 8 * slot, phase);
        tc_fence_after();
        if (lane == 0) {
          const uint64_t adesc = make_smem_desc(sA_u + slot * kAStageBytes);
          const uint64_t bdesc = make_smem_desc(sB_u + slot * kBStageBytes);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            umma_tf32(tacc, adesc + (uint64_t)(j * 2), bdesc + (uint64_t)(j * 2), idesc, (it | j) ? 1u : 0u);
          umma_commit(bar_empty + 8 * slot);
          if (it == nstages - 1) umma_commit(bar_accf + 8 * b);
        }
        __syncwarp();
        if (++slot == S) { slot = 0; phase ^= 1; }
      }
    }
  } else {
    // =========================== EPILOGUE (4 warps) ===========================
    const int q = warp & 3;
    const int r = q * 32 + lane;
    int n_acc = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int nstages = tile_nact(tile) * SPE + st2;
      const int b = n_acc & 1;
      const int row = tile * kTileM + r;
      const bool row_ok = row < n_out;
      float acc[NPAD];
      if (nstages > 0) {
        mbar_wait(bar_accf + 8 * b, (n_acc >> 1) & 1);
        ++n_acc;
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * NPAD);
#pragma unroll
        for (int cb = 0; cb < NPAD / 8; ++cb) tmem_ld8(taddr + cb * 8, acc + cb * 8);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        tc_fence_before();
        mbar_arrive(bar_acce + 8 * b);
      } else {
#pragma unroll
        for (int c = 0; c < NPAD; ++c) acc[c] = 0.f;
      }
      if (!row_ok) continue;
      const int cout = a.cout;
#pragma unroll
      for (int c = 0; c < NPAD; ++c)
        if (c < cout) {
          float v = acc[c];
          if (a.shift) v += __ldg(a.shift + c);
          if (a.res) v += __ldg(a.res + (int64_t)row * a.res_ld + c);
          if (a.relu) v = fmaxf(v, 0.f);
          acc[c] = v;
        }
      if (a.head_out) {
        float s = a.head_b;
#pragma unroll
        for (int c = 0; c < 8; ++c) s = fmaf(acc[c], __ldg(a.head_w + c), s);
        a.head_out[row] = s;
      }
      if (a.out) {
        float* o = a.out + (int64_t)row * a.out_ld;
#pragma unroll
        for (int c = 0; c < NPAD; c += 4)
          if (c < cout) {
            float4 v = make_float4(acc[c], acc[c + 1], acc[c + 2], acc[c + 3]);
            if (p.round_out) { v.x = round_tf32(v.x); v.y = round_tf32(v.y); v.z = round_tf32(v.z); v.w = round_tf32(v.w); }
            *reinterpret_cast<float4*>(o + c) = v;
          }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::kTmemCols)
                 : "memory");
}

// ---- host: tensor maps (driver entry point fetched through the runtime, no libcuda link) ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// fp32 row-major [rows][cols] with leading dimension ld (floats); box = 32 columns x box_rows rows, 128B swizzle
static bool make_map(CUtensorMap* tm, const float* base, int64_t cols, int64_t rows, int64_t ld, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int NPAD>
static int launch_tma(const sps_conv_args& a, const UmmaParams& p, cudaStream_t st) {
  CUtensorMap tmA, tmA2, tmB;
  const int64_t rows_cap = (int64_t)1 << 31;   // indices are validated by the kernel map; -1 / beyond = zero fill
  if (!make_map(&tmA, a.in, a.cin, rows_cap - 1, a.in_ld, 1)) return SPS_ERR_UNSUPPORTED;
  if (a.in2) {
    if (!make_map(&tmA2, a.in2, a.cin2, a.n_out_max > 0 ? a.n_out_max : 1, a.in2_ld, kTileM)) return SPS_ERR_UNSUPPORTED;
  } else {
    tmA2 = tmA;
  }
  if (!make_map(&tmB, p.wt, p.ldk, a.cout, p.ldk, NPAD)) return SPS_ERR_UNSUPPORTED;
  const size_t smem = TmaCfg<NPAD>::smem;
  static bool attr_set = false;
  if (!attr_set) {
    SPS_CUDA_CHECK(cudaFuncSetAttribute(k_conv_umma_tma<NPAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  int64_t tiles = (a.n_out_max + kTileM - 1) / kTileM;
  if (tiles < 1) tiles = 1;
  const int grid = (int)(tiles < 148 ? tiles : 148);
  k_conv_umma_tma<NPAD><<<grid, kTmaThreads, smem, st>>>(a, p, tmA, tmA2, tmB);
  SPS_CUDA_CHECK(cudaGetLastError());
  return SPS_OK;
}

bool conv_umma_tma_supports(const sps_conv_args& a) {
  // 16-byte global alignment of rows (TMA stride rule) and the wide-channel stage plan
  return !a.perm && padded_groups(a.cin) >= 8 && (a.in_ld % 4) == 0 && (!a.in2 || (a.in2_ld % 4) == 0) && get_encode() != nullptr;
}

int conv_umma_tma(const sps_conv_args& a, const UmmaParams& p, cudaStream_t st) {
  switch (a.cout) {
    case 8:
    case 16: return launch_tma<16>(a, p, st);
    case 32: return launch_tma<32>(a, p, st);
    case 64: return launch_tma<64>(a, p, st);
    default: return SPS_ERR_UNSUPPORTED;
  }
}

}  // namespace sps
