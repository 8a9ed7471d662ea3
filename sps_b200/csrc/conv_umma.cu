// tcgen05 implicit-GEMM sparse convolution for sm_100a (the one dense contraction of the path).
//
//   out[o] = act( sum_k in[map[k][o]] @ W[k]  (+ in2[o] @ W2)  + shift (+ res[o]) )
//
// One CTA owns a tile of 128 output voxels = the 128 TMEM lanes of one fp32 accumulator
// [128 x N] (N = Cout padded to 16).  The GEMM K dimension is the im2col row
// (kernel offset k, input channel ci), walked in 16-byte groups (4 fp32 channels):
//   * per tile, a prologue ballots which kernel offsets have at least one neighbour present in
//     the tile and only those offsets are walked (sparsity skip at tile granularity);
//   * 8 groups = one pipeline stage = 128 rows x 128 B, written by cp.async (zero-fill for
//     absent neighbours) straight into the UMMA canonical K-major SWIZZLE_128B layout; the
//     matching weight stage (N rows x 128 B of the K-major, TF32-rounded weight matrix) is
//     fetched the same way from L2;
//   * one elected thread issues 4 x tcgen05.mma.kind::tf32 (M=128, N, K=8) per stage into TMEM,
//     tcgen05.commit releases the stage through an mbarrier; a 4-deep ring overlaps gather
//     and MMA;
//   * epilogue: tcgen05.ld the accumulator row of each voxel, add the folded BatchNorm shift,
//     optional residual, ReLU, optional fused 8->1 head, store fp32 (optionally TF32-rounded so
//     that the next layer's operand rounding is round-to-nearest, not truncation).
// Operands are TF32 (fp32 storage), accumulation is fp32: the north star's 2e-3 score budget
// leaves ~10x margin (DESIGN.md "precision"); bf16 storage would halve the gather bytes but
// measured 6e-4..2e-2 score error on the same network.
#include <cstring>
#include "common.cuh"

namespace sps {

constexpr int kTileM = 128;
constexpr int kStages = 4;
constexpr int kAStageBytes = kTileM * 128;  // 16 KB
constexpr int kUmmaThreads = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// SM100 shared-memory matrix descriptor, K-major, SWIZZLE_128B: 8-row x 128-byte atoms, 1024 B
// apart (SBO); LBO unused for swizzled K-major; version 1; layout type 2.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);         // start address   bits [0,14)
  d |= (uint64_t)1 << 16;                          // LBO (ignored)   bits [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;                // SBO = 1024 B    bits [32,46)
  d |= (uint64_t)1 << 46;                          // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                          // SWIZZLE_128B
  return d;
}
// kind::tf32, fp32 accumulate, both operands K-major: c_format F32 (1<<4), a/b format TF32 (2),
// N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc_tf32(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ float round_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

struct UmmaParams {
  const float* wt;   // [cout][ldk] K-major, TF32-rounded; K index = k*cin + ci, then cin2 entries of the 1x1 term
  int64_t ldk;
  int round_out;
};

template <int NPAD>
__global__ void __launch_bounds__(kUmmaThreads, 2) k_conv_umma(const sps_conv_args a, const UmmaParams p) {
  constexpr int kBStageBytes = NPAD * 128;
  constexpr int kTmemCols = NPAD < 32 ? 32 : NPAD;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment is required by the 128B swizzle atoms
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + kStages * kAStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + kStages * kBStageBytes);  // [kStages] empty + [1] accum
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kStages + 1);
  uint32_t* kmask = tmem_slot + 1;                        // [4] bitmask of active offsets
  uint8_t* klist = reinterpret_cast<uint8_t*>(kmask + 4);  // [128] active offsets in order
  int* nact_s = reinterpret_cast<int*>(klist + 128);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t sA_u = smem_u32(sA), sB_u = smem_u32(sB);
  const uint32_t bar_empty = smem_u32(bars), bar_accum = smem_u32(bars + kStages);

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(bar_empty + 8 * s, 1);
    mbar_init(bar_accum, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int n_out = *a.n_out;
  const int ntiles = (n_out + kTileM - 1) / kTileM;
  const int K = a.K, gpk = a.cin >> 2, gpk2 = a.in2 ? (a.cin2 >> 2) : 0;
  const uint32_t idesc = make_idesc_tf32(NPAD);
  const uint32_t swz = (uint32_t)(tid & 7);
  const uint32_t a_row_off = (uint32_t)((tid >> 3) * 1024 + (tid & 7) * 128);

  uint32_t gstage = 0;  // stages issued so far by this CTA (ring position + mbarrier phases)
  uint32_t accum_uses = 0;

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int row = tile * kTileM + tid;
    const bool row_ok = row < n_out;

    // ---- prologue: which kernel offsets are present anywhere in this tile? ----
    if (tid < 4) kmask[tid] = 0;
    __syncthreads();
    for (int k = 0; k < K; ++k) {
      const int idx = row_ok ? __ldg(a.map + (int64_t)k * a.map_ld + row) : -1;
      const bool any = __any_sync(0xffffffffu, idx >= 0);
      if (lane == 0 && any) atomicOr(&kmask[k >> 5], 1u << (k & 31));
    }
    __syncthreads();
    if (tid == 0) {
      int n = 0;
      for (int k = 0; k < K; ++k)
        if (kmask[k >> 5] & (1u << (k & 31))) klist[n++] = (uint8_t)k;
      *nact_s = n;
    }
    __syncthreads();
    const int nact = *nact_s;
    const int G = nact * gpk + gpk2;          // 16-byte groups along the im2col K dimension
    const int nstages = (G + 7) >> 3;

    // ---- main loop: gather (cp.async) -> MMA (tcgen05) through a kStages-deep ring ----
    for (int st = 0; st < nstages + kStages - 1; ++st) {
      if (st < nstages) {
        const uint32_t gs = gstage + st;
        const uint32_t slot = gs % kStages, use = gs / kStages;
        if (use > 0) mbar_wait(bar_empty + 8 * slot, (use - 1) & 1);  // MMAs that read this slot are done
        const uint32_t a_dst = sA_u + slot * kAStageBytes + a_row_off;
        int cur_e = -1, idx = -1;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const int gi = st * 8 + c;
          const float* src = a.in;
          uint32_t bytes = 0;
          if (gi < nact * gpk) {
            const int e = gi / gpk, cg = gi - e * gpk;
            if (e != cur_e) {
              cur_e = e;
              idx = row_ok ? __ldg(a.map + (int64_t)klist[e] * a.map_ld + row) : -1;
            }
            if (idx >= 0) { src = a.in + (int64_t)idx * a.in_ld + cg * 4; bytes = 16; }
          } else if (gi < G && row_ok) {
            src = a.in2 + (int64_t)row * a.in2_ld + (gi - nact * gpk) * 4;
            bytes = 16;
          }
          cp_async16(a_dst + (((uint32_t)c ^ swz) << 4), src, bytes);
        }
        // weights: chunk column c = tid & 7 of rows n = tid/8 + 16*i
        {
          const int c = tid & 7;
          const int gi = st * 8 + c;
          int64_t koff = -1;
          if (gi < nact * gpk) {
            const int e = gi / gpk, cg = gi - e * gpk;
            koff = ((int64_t)klist[e] * gpk + cg) * 4;
          } else if (gi < G) {
            koff = ((int64_t)K * gpk + (gi - nact * gpk)) * 4;
          }
          const uint32_t b_dst = sB_u + slot * kBStageBytes;
#pragma unroll
          for (int i = 0; i < NPAD / 16; ++i) {
            const int n = (tid >> 3) + 16 * i;
            const bool ok = koff >= 0 && n < a.cout;
            const float* src = ok ? p.wt + (int64_t)n * p.ldk + koff : p.wt;
            cp_async16(b_dst + (uint32_t)((n >> 3) * 1024 + (n & 7) * 128) + (((uint32_t)c ^ (uint32_t)(n & 7)) << 4), src,
                       ok ? 16u : 0u);
          }
        }
      }
      cp_async_commit();
      const int cs = st - (kStages - 1);
      if (cs >= 0) {
        cp_async_wait<kStages - 1>();
        fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
        __syncthreads();
        if (tid == 0) {
          tc_fence_after();
          const uint32_t slot = (gstage + cs) % kStages;
          const uint64_t adesc = make_smem_desc(sA_u + slot * kAStageBytes);
          const uint64_t bdesc = make_smem_desc(sB_u + slot * kBStageBytes);
#pragma unroll
          for (int j = 0; j < 4; ++j)  // 4 x (K = 8 tf32 = 32 bytes) inside the 128-byte swizzle atom
            umma_tf32(tmem_base, adesc + (uint64_t)(j * 2), bdesc + (uint64_t)(j * 2), idesc, (cs | j) ? 1u : 0u);
          umma_commit(bar_empty + 8 * slot);
          if (cs == nstages - 1) umma_commit(bar_accum);
        }
      }
    }
    gstage += nstages;

    // ---- epilogue ----
    float acc[NPAD];
    if (nstages > 0) {
      mbar_wait(bar_accum, accum_uses & 1);
      ++accum_uses;
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll
      for (int cb = 0; cb < NPAD / 8; ++cb) tmem_ld8(taddr + cb * 8, acc + cb * 8);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    } else {
#pragma unroll
      for (int c = 0; c < NPAD; ++c) acc[c] = 0.f;
    }
    if (row_ok) {
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
    // the next tile's first MMA overwrites the accumulator: order it after these TMEM reads
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }

  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)kTmemCols)
                 : "memory");
}

template <int NPAD>
static int launch_umma(const sps_conv_args& a, const UmmaParams& p, cudaStream_t st) {
  const size_t smem = 1024 + kStages * (kAStageBytes + NPAD * 128) + 8 * (kStages + 1) + 4 + 16 + 128 + 16;
  static bool attr_set = false;
  if (!attr_set) {
    SPS_CUDA_CHECK(cudaFuncSetAttribute(k_conv_umma<NPAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  int64_t tiles = (a.n_out_max + kTileM - 1) / kTileM;
  if (tiles < 1) tiles = 1;
  const int grid = (int)(tiles < 148 * 2 ? tiles : 148 * 2);
  k_conv_umma<NPAD><<<grid, kUmmaThreads, smem, st>>>(a, p);
  SPS_CUDA_CHECK(cudaGetLastError());
  return SPS_OK;
}

bool conv_umma_supports(const sps_conv_args& a) {
  if (a.mode != SPS_CONV_NBR || !a.map || !a.weight_kmajor) return false;
  if (a.K < 1 || a.K > 125) return false;
  if (a.cin < 4 || (a.cin & 3) || (a.in_ld & 3)) return false;
  if (a.in2 && ((a.cin2 & 3) || (a.in2_ld & 3))) return false;
  if (!(a.cout == 8 || a.cout == 16 || a.cout == 32 || a.cout == 64)) return false;
  if (a.kmajor_ld & 3) return false;
  return true;
}

int conv_umma(const sps_conv_args& a, cudaStream_t st) {
  UmmaParams p;
  p.wt = a.weight_kmajor;
  p.ldk = a.kmajor_ld;
  p.round_out = a.round_out;
  switch (a.cout) {
    case 8:
    case 16: return launch_umma<16>(a, p, st);
    case 32: return launch_umma<32>(a, p, st);
    case 64: return launch_umma<64>(a, p, st);
    default: return SPS_ERR_UNSUPPORTED;
  }
}

}  // namespace sps

// Host helper: ME-layout weights [K][cin][cout] (+ optional 1x1 term [cin2][cout]) -> K-major
// [cout][ldk] with ldk = K*cin + cin2 rounded up to 4, values rounded to TF32 (nearest-even).
extern "C" int64_t sps_conv_kmajor_ld(int K, int cin, int cin2) { return ((int64_t)K * cin + cin2 + 3) & ~int64_t(3); }

extern "C" int sps_conv_pack_kmajor(const float* w, int K, int cin, int cout, const float* w2, int cin2, float* out) {
  if (!w || !out || K < 1 || cin < 1 || cout < 1 || (w2 == nullptr) != (cin2 == 0)) return SPS_ERR_BAD_ARG;
  const int64_t ldk = sps_conv_kmajor_ld(K, cin, cin2);
  auto rnd = [](float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    u += 0xFFFu + ((u >> 13) & 1u);
    u &= ~0x1FFFu;
    float y;
    memcpy(&y, &u, 4);
    return y;
  };
  for (int n = 0; n < cout; ++n) {
    float* row = out + (int64_t)n * ldk;
    for (int64_t i = 0; i < ldk; ++i) row[i] = 0.f;
    for (int k = 0; k < K; ++k)
      for (int ci = 0; ci < cin; ++ci) row[(int64_t)k * cin + ci] = rnd(w[((int64_t)k * cin + ci) * cout + n]);
    for (int ci = 0; ci < cin2; ++ci) row[(int64_t)K * cin + ci] = rnd(w2[(int64_t)ci * cout + n]);
  }
  return SPS_OK;
}
