// fp32 CUDA-core sparse convolution for the layers with EIGHT output channels (conv1p1s2, block1, conv2p2s2,
// convtr7p2s2, block8 + the fused `final` head: minkunet.py:64-73,140-158).
//
// Why not the tensor-core kernel: with 8 output channels a gathered input row feeds 64-128 multiply-adds, the
// implicit GEMM spends its time materialising a 128-byte-per-row A tile (zero-filled for the absent neighbours,
// 1-3 % tensor-pipe activity measured in round 1) and the contraction itself is noise.  Here a row is gathered
// straight into registers -- ONE 16-byte load per (row, offset) pair and lane, nothing for absent neighbours --
// and multiplied with fp32 weights held in shared memory (warp-uniform broadcast reads, packed FFMA2).  Weights
// are NOT rounded, accumulation is fp32 in ascending kernel-offset order: these layers are exact fp32, which is
// what the level-0 tail of the network needs for the 2e-3 score bar (tools/precision_study.py).
//
// Work split: a 128-row tile per CTA iteration; a row is served by G = (bytes per input row) / 16 lanes, lane `sub`
// owning the 16-byte chunk `sub` of every gathered row (so a warp-wide load touches one 128-byte line per row,
// whatever the row width) and the partial sums of its channels; the G partial sums meet once per row, after the
// last offset (the contraction is linear).  Offsets are walked per TILE in ascending order (tile masks); a warp
// skips an offset none of its rows has.
#include "umma_common.cuh"

namespace sps {

constexpr int kFmaU = 4;   // offsets in flight per thread: 4 index loads, then 4 row loads, then the arithmetic

template <int CIN, bool IN_F32>
struct FmaCfg {
  static constexpr int G = CIN * (IN_F32 ? 4 : 2) / 16;   // lanes per row = 16-byte chunks per input row
  static constexpr int CPL = CIN / G;                     // input channels per lane: 8 (fp16) or 4 (fp32)
  static constexpr int kThreads = kTileM * G;
  static_assert(G == 1 || G == 2 || G == 4, "one 16-byte chunk per lane");
};

// acc += x * w on two packed fp32 lanes (FFMA2 with a broadcast scalar operand)
__device__ __forceinline__ void ffma2(float2& acc, float x, float2 w) {
  unsigned long long d = *reinterpret_cast<unsigned long long*>(&acc);
  float2 xx = make_float2(x, x);
  asm("fma.rn.f32x2 %0, %1, %2, %0;"
      : "+l"(d)
      : "l"(*reinterpret_cast<unsigned long long*>(&xx)), "l"(*reinterpret_cast<unsigned long long*>(&w)));
  acc = *reinterpret_cast<float2*>(&d);
}

// one 16-byte chunk -> NCH fp32 channels (8 halves or 4 floats)
template <bool F32>
__device__ __forceinline__ void unpack_chunk(const uint4& v, float (&x)[8]) {
  if (F32) {
    x[0] = __uint_as_float(v.x); x[1] = __uint_as_float(v.y); x[2] = __uint_as_float(v.z); x[3] = __uint_as_float(v.w);
  } else {
    const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
    for (int j = 0; j < 4; ++j) { const float2 f = __half22float2(h[j]); x[2 * j] = f.x; x[2 * j + 1] = f.y; }
  }
}

// acc[0..7] += sum_c x[c] * w[c][0..7]
template <int NCH>
__device__ __forceinline__ void fma_rows(float2 (&acc)[4], const float (&x)[8], const float* __restrict__ w) {
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const float4 w0 = *reinterpret_cast<const float4*>(w + c * 8);
    const float4 w1 = *reinterpret_cast<const float4*>(w + c * 8 + 4);
    ffma2(acc[0], x[c], make_float2(w0.x, w0.y));
    ffma2(acc[1], x[c], make_float2(w0.z, w0.w));
    ffma2(acc[2], x[c], make_float2(w1.x, w1.y));
    ffma2(acc[3], x[c], make_float2(w1.z, w1.w));
  }
}

template <int CIN, bool IN_F32>
__global__ void __launch_bounds__(FmaCfg<CIN, IN_F32>::kThreads) k_conv_fma8(const sps_conv_args a) {
  using Cfg = FmaCfg<CIN, IN_F32>;
  constexpr int G = Cfg::G, CPL = Cfg::CPL;
  extern __shared__ __align__(16) float w_s[];   // [K][CIN][8] (+ [cin2][8]): ME layout, BN scale folded by the caller
  __shared__ uint8_t s_klist[96];
  const int tid = threadIdx.x;
  const int K = a.K;
  const int nw = K * CIN * 8, nw2 = a.in2 ? a.cin2 * 8 : 0;
  for (int i = tid * 4; i < nw; i += blockDim.x * 4)
    *reinterpret_cast<float4*>(w_s + i) = __ldg(reinterpret_cast<const float4*>(a.weight + i));
  for (int i = tid * 4; i < nw2; i += blockDim.x * 4)
    *reinterpret_cast<float4*>(w_s + nw + i) = __ldg(reinterpret_cast<const float4*>(a.weight2 + i));

  const int n_out = *a.n_out;
  const int ntiles = (n_out + kTileM - 1) / kTileM;
  const int sub = tid % G, rloc = tid / G;
  const bool in2_f16 = (a.io_dtype & SPS_IO_IN2_F16) != 0, out_f16 = (a.io_dtype & SPS_IO_OUT_F16) != 0;
  const char* in_b = reinterpret_cast<const char*>(a.in) + sub * 16;
  const int64_t in_ld_b = a.in_ld * (IN_F32 ? 4 : 2);
  const float* w_sub = w_s + sub * CPL * 8;      // this lane's channel slice inside every W[k]
  const uint32_t kbits[3] = {K >= 32 ? 0xFFFFFFFFu : (1u << K) - 1u,
                             K >= 64 ? 0xFFFFFFFFu : (K > 32 ? (1u << (K - 32)) - 1u : 0u),
                             K > 64 ? (1u << (K - 64)) - 1u : 0u};

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    __syncthreads();   // the previous tile's offset list is no longer read (first trip: the weights have landed)
    uint32_t m[3];
#pragma unroll
    for (int w = 0; w < 3; ++w) m[w] = a.tile_mask ? (__ldg(a.tile_mask + 4 * tile + w) & kbits[w]) : kbits[w];
    if (tid < 96) {
      const int w = tid >> 5, l = tid & 31;
      if ((m[w] >> l) & 1u)
        s_klist[(w > 0 ? __popc(m[0]) : 0) + (w > 1 ? __popc(m[1]) : 0) + __popc(m[w] & ((1u << l) - 1u))] = (uint8_t)tid;
    }
    const int nact = __popc(m[0]) + __popc(m[1]) + __popc(m[2]);
    __syncthreads();
    const int prow = tile * kTileM + rloc;
    const int32_t* sl = a.tile_slices ? a.tile_slices + (int64_t)tile * (SPS_TILE_SLICE_ENTRIES * kTileM) + rloc : nullptr;
    int own;
    if (sl) own = __ldg(sl + nact * kTileM);
    else own = prow < n_out ? (a.perm ? __ldg(a.perm + prow) : prow) : -1;
    const int32_t* mp = (!sl && a.map && own >= 0) ? a.map + own : nullptr;

    float2 acc[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
    for (int e0 = 0; e0 < nact; e0 += kFmaU) {
      int idx[kFmaU];
      uint4 v[kFmaU];
#pragma unroll
      for (int u = 0; u < kFmaU; ++u) {
        const int e = e0 + u;
        idx[u] = -1;
        if (e < nact) {
          if (sl) idx[u] = __ldg(sl + e * kTileM);
          else if (a.map) { if (mp) idx[u] = __ldg(mp + (int64_t)s_klist[e] * a.map_ld); }
          else idx[u] = own;                    // K == 1 without a map: the identity (1x1 convolution)
        }
      }
#pragma unroll
      for (int u = 0; u < kFmaU; ++u)
        if (idx[u] >= 0) v[u] = __ldg(reinterpret_cast<const uint4*>(in_b + (int64_t)idx[u] * in_ld_b));
#pragma unroll
      for (int u = 0; u < kFmaU; ++u) {
        if (idx[u] >= 0) {                      // a warp whose rows all lack this offset skips it
          float x[8];
          unpack_chunk<IN_F32>(v[u], x);
          fma_rows<CPL>(acc, x, w_sub + (int)s_klist[e0 + u] * (CIN * 8));
        }
      }
    }
    // fused 1x1 term on the tile's own rows (BasicBlock downsample, resnet.py:97-108): chunk j of the row goes to lane j % G
    if (a.in2 && own >= 0) {
      const int cpc = in2_f16 ? 8 : 4;
      const int nch = a.cin2 / cpc;
      const char* row2 = reinterpret_cast<const char*>(a.in2) + (int64_t)own * a.in2_ld * (in2_f16 ? 2 : 4);
      for (int j = sub; j < nch; j += G) {
        const uint4 vv = __ldg(reinterpret_cast<const uint4*>(row2 + j * 16));
        float x[8];
        if (in2_f16) { unpack_chunk<false>(vv, x); fma_rows<8>(acc, x, w_s + nw + j * 64); }
        else { unpack_chunk<true>(vv, x); fma_rows<4>(acc, x, w_s + nw + j * 32); }
      }
    }
    // the G partial sums of a row meet here (every lane of the warp takes part in the shuffles)
#pragma unroll
    for (int d = 1; d < G; d <<= 1) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        acc[q].x += __shfl_xor_sync(0xffffffffu, acc[q].x, d);
        acc[q].y += __shfl_xor_sync(0xffffffffu, acc[q].y, d);
      }
    }
    if (sub == 0 && own >= 0) {
      float o[8] = {acc[0].x, acc[0].y, acc[1].x, acc[1].y, acc[2].x, acc[2].y, acc[3].x, acc[3].y};
      if (a.shift) {
#pragma unroll
        for (int c = 0; c < 8; ++c) o[c] += __ldg(a.shift + c);
      }
      if (a.res) {
        float r[8];
        if (in2_f16) {
          unpack_chunk<false>(__ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __half*>(a.res) + (int64_t)own * a.res_ld)), r);
        } else {
          const float4 r0 = __ldg(reinterpret_cast<const float4*>(a.res + (int64_t)own * a.res_ld));
          const float4 r1 = __ldg(reinterpret_cast<const float4*>(a.res + (int64_t)own * a.res_ld + 4));
          r[0] = r0.x; r[1] = r0.y; r[2] = r0.z; r[3] = r0.w; r[4] = r1.x; r[5] = r1.y; r[6] = r1.z; r[7] = r1.w;
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) o[c] += r[c];
      }
      if (a.relu) {
#pragma unroll
        for (int c = 0; c < 8; ++c) o[c] = fmaxf(o[c], 0.f);
      }
      if (a.head_out) {   // final 1x1 conv + bias (minkunet.py:152-158,219)
        float s = a.head_b;
#pragma unroll
        for (int c = 0; c < 8; ++c) s = fmaf(o[c], __ldg(a.head_w + c), s);
        a.head_out[own] = s;
      }
      if (a.out) store_row8(a.out, a.out_ld, own, o, out_f16 ? kStoreF16 : (a.round_out ? kStoreTF32 : kStoreF32));
    }
  }
}

template <int CIN, bool IN_F32>
static int launch_fma8(const sps_conv_args& a, cudaStream_t st) {
  using Cfg = FmaCfg<CIN, IN_F32>;
  const size_t smem = ((size_t)a.K * CIN * 8 + (a.in2 ? (size_t)a.cin2 * 8 : 0)) * sizeof(float);
  int64_t tiles = (a.n_out_max + kTileM - 1) / kTileM;
  if (tiles < 1) tiles = 1;
  // a CTA re-reads the weights (<= 42 KB, L2-resident) once: enough CTAs to fill every SM's thread slots, not more
  const int per_sm = 2048 / Cfg::kThreads;
  const int64_t cap = (int64_t)148 * per_sm;
  k_conv_fma8<CIN, IN_F32><<<(int)(tiles < cap ? tiles : cap), Cfg::kThreads, smem, st>>>(a);
  SPS_CUDA_CHECK(cudaGetLastError());
  return SPS_OK;
}

// Which calls this kernel takes: 8 output channels, 8 or 16 input channels, 16-byte aligned rows, weights (+ the
// fused 1x1 term) within the default 48 KB of dynamic shared memory.
bool conv_fma8_supports(const sps_conv_args& a) {
  if (a.mode != SPS_CONV_NBR || a.cout != 8 || !(a.cin == 8 || a.cin == 16)) return false;
  if (a.K < 1 || a.K > kMaxK || (!a.map && !a.tile_slices && a.K != 1)) return false;
  if (a.tile_slices && !a.tile_mask) return false;
  const bool in_f16 = (a.io_dtype & SPS_IO_IN_F16) != 0, in2_f16 = (a.io_dtype & SPS_IO_IN2_F16) != 0,
             out_f16 = (a.io_dtype & SPS_IO_OUT_F16) != 0;
  const int ea = in_f16 ? 8 : 4, e2 = in2_f16 ? 8 : 4, eo = out_f16 ? 8 : 4;   // elements per 16 bytes
  if ((a.in_ld % ea) || (reinterpret_cast<uintptr_t>(a.in) & 15)) return false;
  if (a.in2 && (!a.weight2 || a.cin2 < e2 || (a.cin2 % e2) || (a.in2_ld % e2) || (reinterpret_cast<uintptr_t>(a.in2) & 15) ||
                (reinterpret_cast<uintptr_t>(a.weight2) & 15)))
    return false;
  if (a.res && ((a.res_ld % e2) || (reinterpret_cast<uintptr_t>(a.res) & 15))) return false;
  if (a.out && ((a.out_ld % eo) || (reinterpret_cast<uintptr_t>(a.out) & 15))) return false;
  if (!a.out && !a.head_out) return false;
  if (!a.weight || (reinterpret_cast<uintptr_t>(a.weight) & 15)) return false;
  const size_t smem = ((size_t)a.K * a.cin * 8 + (a.in2 ? (size_t)a.cin2 * 8 : 0)) * sizeof(float);
  return smem <= 48 * 1024;
}

int conv_fma8(const sps_conv_args& a, cudaStream_t st) {
  const bool in_f32 = (a.io_dtype & SPS_IO_IN_F16) == 0;
  if (a.cin == 8) return in_f32 ? launch_fma8<8, true>(a, st) : launch_fma8<8, false>(a, st);
  return in_f32 ? launch_fma8<16, true>(a, st) : launch_fma8<16, false>(a, st);
}

}  // namespace sps
