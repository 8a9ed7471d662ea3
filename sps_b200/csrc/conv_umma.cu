// tcgen05 implicit-GEMM sparse convolution for sm_100a (the one dense contraction of the path).
//
//   out[o] = act( sum_k in[map[k][o]] @ W[k]  (+ in2[o] @ W2)  + shift (+ res[o]) )
//
// A tile is 128 output voxels = the 128 TMEM lanes of one fp32 accumulator [128 x N] (N = Cout
// padded to 16).  One persistent CTA per SM (416 threads) is warp-specialised the Blackwell way:
// 8 producer warps gather, 1 warp issues tcgen05.mma, 4 warps run the epilogue; they only meet
// through mbarrier rings (stage full/empty, accumulator full/empty, kernel-map slice ready), so
// the gathers of tile i+1 overlap the MMAs of tile i and the epilogue of tile i-1 (two TMEM
// accumulators).  The GEMM K dimension is the im2col row
// (kernel offset k, input channel ci), walked in 16-byte groups (4 fp32 channels):
//   * prologue: the tile's slice of the kernel map is staged in shared memory with cp.async
//     (all K loads in flight at once) and a warp ballot finds the offsets that have at least
//     one neighbour inside the tile -- only those are walked (sparsity skip at tile level);
//   * one pipeline stage = 128 rows x 128 B of gathered A (8 groups) written by cp.async
//     (zero-fill for absent neighbours) straight into the UMMA canonical K-major SWIZZLE_128B
//     layout, plus the matching N rows x 128 B of the K-major, TF32-rounded weight matrix.
//     Channel counts are padded per kernel offset so that a stage never straddles offsets in
//     an irregular way: Cin <= 16 packs 8/(Cin/4) offsets per stage, Cin >= 24 uses
//     ceil(Cin/32) stages per offset;
//   * one elected thread issues 4 x tcgen05.mma.kind::tf32 (M=128, N, K=8) per stage into
//     TMEM; tcgen05.commit releases the stage through an mbarrier; a 3-deep ring keeps two
//     stages of gathers in flight behind the MMAs, two CTAs per SM overlap prologue/epilogue;
//   * epilogue: all 8 warps tcgen05.ld half an accumulator row each, add the folded BatchNorm
//     shift, optional residual, ReLU, optional fused 8->1 head, store fp32 (optionally
//     TF32-rounded so the next layer's operand rounding is nearest, not truncation).
// Operands are TF32 (fp32 storage), accumulation fp32: ~1e-4 score error on the reference
// network against the 2e-3 budget; bf16 operands measured 6e-4..2e-2 (DESIGN.md "precision").
#include <cstdlib>
#include "umma_common.cuh"

namespace sps {

#ifndef SPS_PRODUCER_GROUPS
#define SPS_PRODUCER_GROUPS 1
#endif
// Producer groups of 8 warps; group g writes the stages with (stage index % groups) == g, so several stages
// of one tile are being gathered at once without multiplying the per-stage instruction count.
constexpr int kGroups = SPS_PRODUCER_GROUPS;
constexpr int kProducerWarps = 8 * kGroups;
__device__ __forceinline__ void producer_bar() { asm volatile("bar.sync 1, %0;" ::"n"(kProducerWarps * 32) : "memory"); }

constexpr int kProducerThreads = kProducerWarps * 32;  // producer warps: gather A/B, stage kernel-map slices
constexpr int kGroupThreads = 256;                    // threads that write one stage
#ifndef SPS_ARRIVE_LAG
#define SPS_ARRIVE_LAG 0
#endif
// Stage-full signalling.  LAG > 0: every thread commits its copies as a cp.async group, and once the group
// issued LAG stages earlier has landed (cp.async.wait_group) ONE lane per warp arrives on that stage's
// barrier -- 8 arrivals per stage.  LAG == 0: every thread arrives asynchronously
// (cp.async.mbarrier.arrive.noinc) -- 256 arrivals on one mbarrier word per stage, which serialise.
constexpr int kArriveLag = SPS_ARRIVE_LAG;
constexpr int kFullArrivals = kArriveLag > 0 ? kGroupThreads / 32 : kGroupThreads;
static_assert(kGroups == 1 || kArriveLag == 0, "lagged arrivals assume one producer group");
constexpr int kRowsPerThread = kTileM * 8 / kGroupThreads;      // A chunks per thread per stage (4)
constexpr int kRowStep = kGroupThreads / 8;           // row distance between a thread's chunks
constexpr int kMmaWarp = kProducerWarps;              // next warp: tcgen05.mma issue
// the 4 warps after it: epilogue (TMEM -> registers -> global)
constexpr int kCtaThreads = kProducerThreads + 32 + 128;

template <int NPAD>
struct UmmaCfg {
#ifndef SPS_S64
#define SPS_S64 5
#endif
#ifndef SPS_S32
#define SPS_S32 6
#endif
  static constexpr int S = NPAD == 64 ? SPS_S64 : SPS_S32;        // ring depth
  static constexpr int kBStage = NPAD * 128;
  static constexpr int kTmemCols = 2 * NPAD < 32 ? 32 : 2 * NPAD;   // two accumulators (double buffered)
  static constexpr size_t smem = (size_t)S * (kAStageBytes + kBStage) + 2 * (size_t)kMaxK * kTileM * 4 +
                                 8 * (2 * S + 6) + 2 * 96 + 32;
};

// GPC = padded groups per offset class: 2 or 4 (several offsets per stage) or 8 (= "8 or more":
// one offset spans GP/8 stages).  Persistent, warp-specialised: producers, MMA issuer and
// epilogue run decoupled through mbarrier rings and never meet at a block-wide barrier.
template <int NPAD, int GPC>
__global__ void __launch_bounds__(kCtaThreads, 1) k_conv_umma(const sps_conv_args a, const UmmaParams p) {
  using Cfg = UmmaCfg<NPAD>;
  constexpr int S = Cfg::S;
  constexpr int kBStageBytes = Cfg::kBStage;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;                                                     // [S][128 rows x 128 B], 128B-swizzled
  uint8_t* sB = smem + S * kAStageBytes;                                  // [S][NPAD rows x 128 B]
  int32_t* sidx = reinterpret_cast<int32_t*>(sB + S * kBStageBytes);      // [2][K][128] kernel-map slices (tile parity)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sidx + 2 * kMaxK * kTileM);
  // bars: full[S], empty[S], idx_full[2], acc_full[2], acc_empty[2]
  uint8_t* klist = reinterpret_cast<uint8_t*>(bars + 2 * S + 6);           // [2][96] present offsets per tile parity
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(klist + 2 * 96);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t sA_u = smem_u32(sA), sB_u = smem_u32(sB), sidx_u = smem_u32(sidx);
  const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 8 * S, bar_idx = bar_empty + 8 * S,
                 bar_accf = bar_idx + 16, bar_acce = bar_accf + 16;
  if (sA_u & 1023) __trap();  // SWIZZLE_128B atoms need 1024-byte aligned stage bases

  if (tid == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(bar_full + 8 * s, kFullArrivals); mbar_init(bar_empty + 8 * s, 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_idx + 8 * i, kProducerThreads);
      mbar_init(bar_accf + 8 * i, 1);
      mbar_init(bar_acce + 8 * i, 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kMmaWarp) {
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
  const int gpk = a.cin >> 2;                           // real groups per offset
  const int GP = GPC < 8 ? GPC : padded_groups(a.cin);  // padded groups per offset
  const int SPE = GPC < 8 ? 1 : GP >> 3;                // stages per offset (large Cin)
  constexpr int EPS = GPC < 8 ? 8 / GPC : 1;            // offsets per stage (small Cin)
  const int gpk2 = a.in2 ? (a.cin2 >> 2) : 0;
  const int st2 = (gpk2 + 7) >> 3;                      // stages of the fused 1x1 term
  const uint32_t* tmask = a.tile_mask;
  auto tile_nact = [&](int tile) {
    return __popc(__ldg(tmask + 4 * tile)) + __popc(__ldg(tmask + 4 * tile + 1)) + __popc(__ldg(tmask + 4 * tile + 2));
  };
  auto tile_stages = [&](int nact) { return (GPC < 8 ? (nact + EPS - 1) / EPS : nact * SPE) + st2; };

  if (warp < kMmaWarp) {
    // =========================== PRODUCERS (kGroups x 256 threads) ===========================
    const uint32_t in_ld_b = (uint32_t)a.in_ld * 4u, in2_ld_b = (uint32_t)a.in2_ld * 4u;
    const char* in_b = reinterpret_cast<const char*>(a.in);
    const char* in2_b = reinterpret_cast<const char*>(a.in2);
    const int grp = tid / kGroupThreads, tg = tid % kGroupThreads;
    const int r0 = tg >> 3, cB = tg & 7;        // gather: chunk column cB of rows r0 + kRowStep*i
    const int rI = tid & 127, hI = tid >> 7;    // kernel-map staging: row rI, offsets hI, hI + kProducerThreads/128, ...
    const uint32_t a_off = (uint32_t)((r0 >> 3) * 1024 + (r0 & 7) * 128) + (((uint32_t)cB ^ (uint32_t)(r0 & 7)) << 4);

    // stage the kernel-map slice of `tile` (only the offsets present in it) into parity buffer `par`
    auto prepare = [&](int tile, int par) {
      producer_bar();   // everybody is done reading klist/sidx of the tile that used this parity before
      if (warp == 0) {  // ordered list of present offsets from the precomputed tile mask
        int base = 0;
        for (int w = 0; w < 3; ++w) {
          const uint32_t bits = __ldg(tmask + 4 * tile + w);
          if ((bits >> lane) & 1u) klist[par * 96 + base + __popc(bits & ((1u << lane) - 1u))] = (uint8_t)(32 * w + lane);
          base += __popc(bits);
        }
      }
      producer_bar();
      const int nact = tile_nact(tile);
      const int row = tile * kTileM + rI;
      const uint32_t dst = sidx_u + (uint32_t)(par * kMaxK * kTileM + rI) * 4u;
      if (row < n_out) {
        const int32_t* src = a.map + (a.perm ? __ldg(a.perm + row) : row);
        for (int e = hI; e < nact; e += kProducerThreads / 128)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + (uint32_t)(e * kTileM) * 4u),
                       "l"(src + (int64_t)klist[par * 96 + e] * a.map_ld)
                       : "memory");
      } else {
        for (int e = hI; e < nact; e += kProducerThreads / 128) sidx[(par * kMaxK + e) * kTileM + rI] = -1;  // rows past the end
      }
      cp_async_arrive(bar_idx + 8 * par);
    };

    // ---- per-thread invariants of the stage writer (kept out of the stage loop: the producers are
    //      issue-bound, every instruction here is paid once per stage per warp) ----
    constexpr int NB = (NPAD + kRowStep - 1) / kRowStep;   // weight chunks per thread per stage
    const float* wrow[NB];
    uint32_t b_off[NB];
    bool wok[NB];
#pragma unroll
    for (int i = 0; i < NB; ++i) {
      const int n = r0 + kRowStep * i;
      wok[i] = n < a.cout && n < NPAD;
      wrow[i] = p.wt + (int64_t)(wok[i] ? n : 0) * p.ldk;
      b_off[i] = (uint32_t)((n >> 3) * 1024 + (n & 7) * 128) + (((uint32_t)cB ^ (uint32_t)(n & 7)) << 4);
    }
    const bool b_lane = r0 < NPAD;                  // narrow layers: only some threads carry a weight chunk
    uint32_t slot = 0, phase = 0;
    uint32_t a_slot = sA_u + a_off, b_slot = sB_u, bar_e = bar_empty, bar_f = bar_full;
    int turn = 0;   // stage counter modulo kGroups: whose turn it is to write the next stage
    int issued = 0;                 // stages this thread has written so far (lagged arrival bookkeeping)
    uint32_t bar_lag = bar_full;    // barrier of the oldest stage whose arrival is still owed

    // write this thread's share of one stage: 4 A chunks (rows r0+32i, column cB) + its weight chunk(s)
    auto emit = [&](const char* base, uint32_t ld_b, const int (&idx)[kRowsPerThread], uint32_t cg_off, bool cg_ok, int kofB) {
      if (turn == grp) {
        mbar_wait(bar_e, phase ^ 1);                // the MMAs that read this slot have completed
#pragma unroll
        for (int i = 0; i < kRowsPerThread; ++i) {
          const bool ok = cg_ok && idx[i] >= 0;
          cp_async16(a_slot + i * (kRowStep * 128), base + (ok ? (uint32_t)idx[i] * ld_b + cg_off : 0u), ok ? 16u : 0u);
        }
        if (b_lane) {
#pragma unroll
          for (int i = 0; i < NB; ++i) {
            const bool ok = kofB >= 0 && wok[i];
            cp_async16(b_slot + b_off[i], ok ? wrow[i] + kofB : p.wt, ok ? 16u : 0u);
          }
        }
        if (kArriveLag == 0) {
          cp_async_arrive(bar_f);                   // fires when this thread's copies of the stage have landed
        } else {
          cp_async_commit();
          if (++issued > kArriveLag) {              // the stage issued kArriveLag stages ago has landed by now
            cp_async_wait<kArriveLag>();
            fence_proxy_async();                    // generic-proxy smem writes -> visible to the tensor core
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_lag);
            bar_lag = (bar_lag == bar_full + 8 * (S - 1)) ? bar_full : bar_lag + 8;
          }
        }
      }
      if (++turn == kGroups) turn = 0;
      if (++slot == S) {
        slot = 0; phase ^= 1;
        a_slot = sA_u + a_off; b_slot = sB_u; bar_e = bar_empty; bar_f = bar_full;
      } else {
        a_slot += kAStageBytes; b_slot += kBStageBytes; bar_e += 8; bar_f += 8;
      }
    };

    int it_tile = 0;
    if ((int)blockIdx.x < ntiles) prepare(blockIdx.x, 0);
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it_tile) {
      const int par = it_tile & 1;
      const int next = tile + gridDim.x;
      if (next < ntiles) prepare(next, par ^ 1);   // one tile ahead: its latency hides behind this tile's gathers
      const int nact = tile_nact(tile);
      mbar_wait(bar_idx + 8 * par, (it_tile >> 1) & 1);
      const int32_t* sx = sidx + par * kMaxK * kTileM + r0;   // [e][128] compacted by present offset
      const uint8_t* kl = klist + par * 96;
      int idx[kRowsPerThread];
      if (GPC < 8) {
        // small Cin: EPS offsets per stage, GPC chunks each; this thread's chunk column picks offset e_off
        constexpr int GPCc = GPC < 8 ? GPC : 1;
        const int e_off = cB / GPCc, cg = cB % GPCc;
        const bool cg_ok = cg < gpk;
        const int nst_map = (nact + EPS - 1) / EPS;
        for (int st = 0, e = e_off; st < nst_map; ++st, e += EPS) {
          const bool e_ok = e < nact && turn == grp;   // another group's stage: nothing to look up
          const int32_t* sk = sx + (e_ok ? e : 0) * kTileM;
#pragma unroll
          for (int i = 0; i < kRowsPerThread; ++i) idx[i] = e_ok ? sk[kRowStep * i] : -1;
          emit(in_b, in_ld_b, idx, (uint32_t)cg * 16u, cg_ok, e_ok ? ((int)kl[e] * GPCc + cg) * 4 : -1);
        }
      } else {
        // large Cin: one offset per SPE stages, the neighbour rows are looked up once per offset
        for (int e = 0; e < nact; ++e) {
          const int32_t* sk = sx + e * kTileM;
#pragma unroll
          for (int i = 0; i < kRowsPerThread; ++i) idx[i] = sk[kRowStep * i];
          const int kbase = (int)kl[e] * GP;
          for (int sub = 0, cg = cB; sub < SPE; ++sub, cg += 8)
            emit(in_b, in_ld_b, idx, (uint32_t)cg * 16u, cg < gpk, (kbase + cg) * 4);
        }
      }
      if (st2 > 0) {  // fused 1x1 term: identity gather from in2
#pragma unroll
        for (int i = 0; i < kRowsPerThread; ++i) {
          const int rw = tile * kTileM + r0 + kRowStep * i;
          idx[i] = rw < n_out ? (a.perm ? __ldg(a.perm + rw) : rw) : -1;
        }
        const int kbase2 = K * GP;
        for (int s2 = 0, cg = cB; s2 < st2; ++s2, cg += 8)
          emit(in2_b, in2_ld_b, idx, (uint32_t)cg * 16u, cg < gpk2, cg < gpk2 ? (kbase2 + cg) * 4 : -1);
      }
    }
    cp_async_wait<0>();
    if (kArriveLag > 0) {           // flush the arrivals still owed for the last stages
      fence_proxy_async();
      __syncwarp();
      const int owed = issued < kArriveLag ? issued : kArriveLag;
      for (int i = 0; i < owed; ++i) {
        if (lane == 0) mbar_arrive(bar_lag);
        bar_lag = (bar_lag == bar_full + 8 * (S - 1)) ? bar_full : bar_lag + 8;
      }
    }
  } else if (warp == kMmaWarp) {
    // =========================== MMA ISSUER (one lane) ===========================
    const uint32_t idesc = make_idesc_tf32(NPAD);
    uint32_t gs = 0;
    int n_acc = 0;   // tiles that actually accumulate (both sides count the same way)
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int nstages = tile_stages(tile_nact(tile));
      if (nstages == 0) continue;   // nothing to accumulate: the epilogue uses zeros
      const int b = n_acc & 1;
      mbar_wait(bar_acce + 8 * b, ((n_acc >> 1) & 1) ^ 1);   // epilogue has drained this accumulator
      ++n_acc;
      tc_fence_after();
      const uint32_t tacc = tmem_base + (uint32_t)(b * NPAD);
      for (int it = 0; it < nstages; ++it, ++gs) {
        const uint32_t slot = gs % S;
        mbar_wait(bar_full + 8 * slot, (gs / S) & 1);
#ifdef SPS_MMA_PROXY_FENCE
        fence_proxy_async();   // not needed: completion of cp.async through the mbarrier orders the writes (as CUTLASS)
#endif
        tc_fence_after();
        if (lane == 0) {
          const uint64_t adesc = make_smem_desc(sA_u + slot * kAStageBytes);
          const uint64_t bdesc = make_smem_desc(sB_u + slot * kBStageBytes);
#pragma unroll
          for (int j = 0; j < 4; ++j)  // 4 x (K = 8 tf32 = 32 bytes) inside the 128-byte swizzle atom
            umma_tf32(tacc, adesc + (uint64_t)(j * 2), bdesc + (uint64_t)(j * 2), idesc, (it | j) ? 1u : 0u);
          umma_commit(bar_empty + 8 * slot);
          if (it == nstages - 1) umma_commit(bar_accf + 8 * b);
        }
        __syncwarp();
      }
    }
  } else {
    // =========================== EPILOGUE (4 warps = 128 TMEM lanes) ===========================
    const int q = warp & 3;                 // TMEM lane quadrant this warp may access
    const int r = q * 32 + lane;
    int n_acc = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int nstages = tile_stages(tile_nact(tile));
      const int b = n_acc & 1;
      const int prow = tile * kTileM + r;
      const bool row_ok = prow < n_out;
      const int row = row_ok && a.perm ? __ldg(a.perm + prow) : prow;   // output row this tile lane stands for
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
        mbar_arrive(bar_acce + 8 * b);      // accumulator may be overwritten by the tile after next
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
  if (warp == kMmaWarp)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::kTmemCols)
                 : "memory");
}

template <int NPAD, int GPC>
static int launch_umma(const sps_conv_args& a, const UmmaParams& p, cudaStream_t st) {
  const size_t smem = UmmaCfg<NPAD>::smem;
  static bool attr_set = false;
  if (!attr_set) {
    SPS_CUDA_CHECK(
        cudaFuncSetAttribute(k_conv_umma<NPAD, GPC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  int64_t tiles = (a.n_out_max + kTileM - 1) / kTileM;
  if (tiles < 1) tiles = 1;
  const int grid = (int)(tiles < 148 ? tiles : 148);
  k_conv_umma<NPAD, GPC><<<grid, kCtaThreads, smem, st>>>(a, p);
  SPS_CUDA_CHECK(cudaGetLastError());
  return SPS_OK;
}

// Per-tile (128 consecutive output rows) bitmask of the kernel offsets that have at least one
// neighbour in the tile: masks[tile][4] (bits 0..K-1 over the first 3 words).
__global__ void __launch_bounds__(128)
k_tile_masks(const int32_t* __restrict__ map, int64_t ld, int K, const int32_t* __restrict__ n_ptr,
             uint32_t* __restrict__ masks) {
  const int n = *n_ptr;
  const int ntiles = (n + kTileM - 1) / kTileM;
  __shared__ uint32_t m[4];
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    if (threadIdx.x < 4) m[threadIdx.x] = 0;
    __syncthreads();
    const int row = tile * kTileM + threadIdx.x;
    for (int k = 0; k < K; ++k) {
      const bool hit = row < n && __ldg(map + (int64_t)k * ld + row) >= 0;
      if (__any_sync(0xffffffffu, hit) && (threadIdx.x & 31) == 0) atomicOr(&m[k >> 5], 1u << (k & 31));
    }
    __syncthreads();
    if (threadIdx.x < 4) masks[4 * tile + threadIdx.x] = m[threadIdx.x];
    __syncthreads();
  }
}

bool conv_umma_supports(const sps_conv_args& a) {
  if (a.mode != SPS_CONV_NBR || !a.map || !a.weight_kmajor || !a.tile_mask) return false;
  if (a.K < 1 || a.K > kMaxK) return false;
  if (a.cin < 4 || (a.cin & 3) || (a.in_ld & 3)) return false;
  if (a.in2 && ((a.cin2 & 3) || (a.in2_ld & 3))) return false;
  if (!(a.cout == 8 || a.cout == 16 || a.cout == 32 || a.cout == 64)) return false;
  if (a.kmajor_ld & 3) return false;
  return true;
}

template <int NPAD>
static int launch_umma_n(const sps_conv_args& a, const UmmaParams& p, cudaStream_t st) {
  const int gp = padded_groups(a.cin);
  if (gp == 2) return launch_umma<NPAD, 2>(a, p, st);
  if (gp == 4) return launch_umma<NPAD, 4>(a, p, st);
  return launch_umma<NPAD, 8>(a, p, st);
}

bool conv_umma_tma_supports(const sps_conv_args& a);
int conv_umma_tma(const sps_conv_args& a, const UmmaParams& p, cudaStream_t st);
int conv_umma6(const sps_conv_args& a, const UmmaParams& p, cudaStream_t st);
static int g_variant = 6;   // 6: k_conv_umma6 (loader warp, static-slot producers); 5: k_conv_umma
static int g_use_tma = 0;   // 1: wide layers gather through TMA tile::gather4 (measured 3x slower than cp.async producers: 128-byte boxes)

// true when the selected kernel generation gathers from the dense map itself (generation 5, TMA variant): the map
// builder must then keep the dense tables complete (maps.cu writes only present entries otherwise)
bool conv_needs_dense_maps() {
  static const char* env = getenv("SPS_UMMA_VARIANT");
  static const int env_variant = env ? atoi(env) : 0;
  return (env_variant ? env_variant : g_variant) != 6 || g_use_tma != 0;
}

int conv_umma(const sps_conv_args& a, cudaStream_t st) {
  UmmaParams p;
  p.wt = a.weight_kmajor;
  p.ldk = a.kmajor_ld;
  p.round_out = a.round_out;
  if (a.io_dtype == SPS_IO_F16) return conv_umma6(a, p, st);   // fp16 rows: generation 6 only
  if (g_use_tma && conv_umma_tma_supports(a)) return conv_umma_tma(a, p, st);
  static const char* env = getenv("SPS_UMMA_VARIANT");   // A/B runs without touching the host code
  static const int env_variant = env ? atoi(env) : 0;
  if ((env_variant ? env_variant : g_variant) == 6) return conv_umma6(a, p, st);
  switch (a.cout) {
    case 8:
    case 16: return launch_umma_n<16>(a, p, st);
    case 32: return launch_umma_n<32>(a, p, st);
    case 64: return launch_umma_n<64>(a, p, st);
    default: return SPS_ERR_UNSUPPORTED;
  }
}

}  // namespace sps

extern "C" int sps_set_umma_variant(int v) {
  if (v != 5 && v != 6) return SPS_ERR_BAD_ARG;
  sps::g_variant = v;
  return SPS_OK;
}

extern "C" int sps_set_tma_gather(int on) {
  sps::g_use_tma = on != 0;
  return SPS_OK;
}

extern "C" int sps_kernel_map_tile_masks(const int32_t* d_map, int64_t map_ld, int K, const int32_t* d_n_out,
                                         int64_t n_out_max, uint32_t* d_masks, void* stream) {
  if (!d_map || !d_n_out || !d_masks || K < 1 || K > sps::kMaxK || n_out_max < 0) return SPS_ERR_BAD_ARG;
  int64_t tiles = (n_out_max + sps::kTileM - 1) / sps::kTileM;
  if (tiles < 1) tiles = 1;
  const int grid = (int)(tiles < 148 * 16 ? tiles : 148 * 16);
  sps::k_tile_masks<<<grid, 128, 0, (cudaStream_t)stream>>>(d_map, map_ld, K, d_n_out, d_masks);
  SPS_CUDA_CHECK(cudaGetLastError());
  return SPS_OK;
}

// Host helper: ME-layout weights [K][cin][cout] (+ optional 1x1 term [cin2][cout]) -> K-major
// [cout][ld]: row n holds, for every kernel offset k, padded_groups(cin)*4 floats (channels of
// W[k][:, n], zero padded), then the 1x1 term padded to a multiple of 32; values rounded to
// TF32 (nearest even).
extern "C" int64_t sps_conv_kmajor_ld(int K, int cin, int cin2) {
  return (int64_t)K * sps::padded_groups(cin) * 4 + ((cin2 + 31) & ~31);
}

extern "C" int sps_conv_pack_kmajor(const float* w, int K, int cin, int cout, const float* w2, int cin2, float* out) {
  if (!w || !out || K < 1 || cin < 1 || cout < 1 || (w2 == nullptr) != (cin2 == 0)) return SPS_ERR_BAD_ARG;
  const int64_t ldk = sps_conv_kmajor_ld(K, cin, cin2);
  const int cpad = sps::padded_groups(cin) * 4;
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
      for (int ci = 0; ci < cin; ++ci) row[(int64_t)k * cpad + ci] = rnd(w[((int64_t)k * cin + ci) * cout + n]);
    for (int ci = 0; ci < cin2; ++ci) row[(int64_t)K * cpad + ci] = rnd(w2[(int64_t)ci * cout + n]);
  }
  return SPS_OK;
}
