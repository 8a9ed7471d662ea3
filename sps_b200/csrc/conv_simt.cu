// fp32 CUDA-core sparse convolution (output-stationary gather, no atomics, deterministic).
//
// out[o] = act( sum_k in[map[k][o]] @ W[k] + shift (+ in2[o] @ W2) (+ res[o]) )
// One thread owns 8 output channels of one output voxel; the threads of a warp share the
// kernel offset and the channel group, so weight reads from shared memory are broadcasts and
// the map reads are coalesced.  This is the parity/reference GPU path (exact fp32 FMA chain)
// and the fallback for layers the tensor-core kernel does not take.
#include "common.cuh"

namespace sps {

constexpr int kSimtThreads = 256;
constexpr int kSimtSmemBytes = 96 * 1024;

template <int COUT>
__global__ void __launch_bounds__(kSimtThreads)
k_conv_simt(const sps_conv_args a, const int kc) {
  constexpr int TPV = COUT / 8;               // threads per voxel
  constexpr int VB = kSimtThreads / TPV;      // voxels per block tile
  extern __shared__ __align__(16) float w_s[];  // [kc][cin][COUT] (later reused for W2)
  const int tid = threadIdx.x;
  const int cg = tid / VB, vl = tid % VB;
  const int n_out = *a.n_out;
  const int cin = a.cin, K = a.K;
  const int slab = cin * COUT;
  const int ntiles = (n_out + VB - 1) / VB;
  const bool resident = (kc >= K) && a.in2 == nullptr;
  bool loaded = false;

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int v = tile * VB + vl;
    const bool active = v < n_out;
    float acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = 0.f;

    for (int kbase = 0; kbase < K; kbase += kc) {
      const int kn = min(kc, K - kbase);
      if (!(resident && loaded)) {
        __syncthreads();
        const float4* src = reinterpret_cast<const float4*>(a.weight + (int64_t)kbase * slab);
        float4* dst = reinterpret_cast<float4*>(w_s);
        for (int i = tid; i < kn * slab / 4; i += kSimtThreads) dst[i] = __ldg(src + i);
        __syncthreads();
        loaded = true;
      }
      if (!active) continue;
      for (int kk = 0; kk < kn; ++kk) {
        const int k = kbase + kk;
        const int idx = a.map ? __ldg(a.map + (int64_t)k * a.map_ld + v) : v;
        if (idx < 0) continue;
        const float* row = a.in + (int64_t)idx * a.in_ld;
        const float* wk = w_s + kk * slab + cg * 8;
        if (cin == 1) {
          const float x = __ldg(row);
          const float4 w0 = *reinterpret_cast<const float4*>(wk);
          const float4 w1 = *reinterpret_cast<const float4*>(wk + 4);
          acc[0] = fmaf(x, w0.x, acc[0]); acc[1] = fmaf(x, w0.y, acc[1]);
          acc[2] = fmaf(x, w0.z, acc[2]); acc[3] = fmaf(x, w0.w, acc[3]);
          acc[4] = fmaf(x, w1.x, acc[4]); acc[5] = fmaf(x, w1.y, acc[5]);
          acc[6] = fmaf(x, w1.z, acc[6]); acc[7] = fmaf(x, w1.w, acc[7]);
        } else {
          for (int ci = 0; ci < cin; ci += 4) {
            const float4 x4 = __ldg(reinterpret_cast<const float4*>(row + ci));
            const float xs[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 w0 = *reinterpret_cast<const float4*>(wk + (ci + j) * COUT);
              const float4 w1 = *reinterpret_cast<const float4*>(wk + (ci + j) * COUT + 4);
              acc[0] = fmaf(xs[j], w0.x, acc[0]); acc[1] = fmaf(xs[j], w0.y, acc[1]);
              acc[2] = fmaf(xs[j], w0.z, acc[2]); acc[3] = fmaf(xs[j], w0.w, acc[3]);
              acc[4] = fmaf(xs[j], w1.x, acc[4]); acc[5] = fmaf(xs[j], w1.y, acc[5]);
              acc[6] = fmaf(xs[j], w1.z, acc[6]); acc[7] = fmaf(xs[j], w1.w, acc[7]);
            }
          }
        }
      }
    }

    // fused 1x1 term (BasicBlock downsample): + in2[v] @ W2
    if (a.in2) {
      __syncthreads();
      const int n4 = a.cin2 * COUT / 4;
      for (int i = tid; i < n4; i += kSimtThreads)
        reinterpret_cast<float4*>(w_s)[i] = __ldg(reinterpret_cast<const float4*>(a.weight2) + i);
      __syncthreads();
      loaded = false;
      if (active) {
        const float* row = a.in2 + (int64_t)v * a.in2_ld;
        const float* wk = w_s + cg * 8;
        for (int ci = 0; ci < a.cin2; ci += 4) {
          const float4 x4 = __ldg(reinterpret_cast<const float4*>(row + ci));
          const float xs[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 w0 = *reinterpret_cast<const float4*>(wk + (ci + j) * COUT);
            const float4 w1 = *reinterpret_cast<const float4*>(wk + (ci + j) * COUT + 4);
            acc[0] = fmaf(xs[j], w0.x, acc[0]); acc[1] = fmaf(xs[j], w0.y, acc[1]);
            acc[2] = fmaf(xs[j], w0.z, acc[2]); acc[3] = fmaf(xs[j], w0.w, acc[3]);
            acc[4] = fmaf(xs[j], w1.x, acc[4]); acc[5] = fmaf(xs[j], w1.y, acc[5]);
            acc[6] = fmaf(xs[j], w1.z, acc[6]); acc[7] = fmaf(xs[j], w1.w, acc[7]);
          }
        }
      }
    }
    if (!active) continue;

    if (a.shift) {
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[c] += __ldg(a.shift + cg * 8 + c);
    }
    if (a.res) {
      const float* r = a.res + (int64_t)v * a.res_ld + cg * 8;
      const float4 r0 = __ldg(reinterpret_cast<const float4*>(r));
      const float4 r1 = __ldg(reinterpret_cast<const float4*>(r + 4));
      acc[0] += r0.x; acc[1] += r0.y; acc[2] += r0.z; acc[3] += r0.w;
      acc[4] += r1.x; acc[5] += r1.y; acc[6] += r1.z; acc[7] += r1.w;
    }
    if (a.relu) {
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[c] = fmaxf(acc[c], 0.f);
    }
    if (a.round_out) {
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        uint32_t r;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(acc[c]));
        acc[c] = __uint_as_float(r);
      }
    }
    if (a.out) {
      float* o = a.out + (int64_t)v * a.out_ld + cg * 8;
      *reinterpret_cast<float4*>(o) = make_float4(acc[0], acc[1], acc[2], acc[3]);
      *reinterpret_cast<float4*>(o + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
    if (COUT == 8 && a.head_out) {
      float s = a.head_b;
#pragma unroll
      for (int c = 0; c < 8; ++c) s = fmaf(acc[c], __ldg(a.head_w + c), s);
      a.head_out[v] = s;
    }
  }
}

template <int COUT>
static int launch_simt(const sps_conv_args& a, cudaStream_t st) {
  constexpr int VB = kSimtThreads / (COUT / 8);
  const int slab = a.cin * COUT * 4;
  int kc = kSimtSmemBytes / slab;
  if (kc > a.K) kc = a.K;
  if (kc < 1) return SPS_ERR_UNSUPPORTED;
  size_t smem = (size_t)kc * slab;
  if (a.in2 && (size_t)a.cin2 * COUT * 4 > smem) smem = (size_t)a.cin2 * COUT * 4;
  static unsigned long long attr_done = 0;   // bit per device
  SPS_CUDA_CHECK(ensure_dynamic_smem(k_conv_simt<COUT>, kSimtSmemBytes, &attr_done));
  int64_t tiles = (a.n_out_max + VB - 1) / VB;
  if (tiles < 1) tiles = 1;
  const int grid = (int)(tiles < 148 * 8 ? tiles : 148 * 8);
  k_conv_simt<COUT><<<grid, kSimtThreads, smem, st>>>(a, kc);
  SPS_CUDA_CHECK(cudaGetLastError());
  return SPS_OK;
}

// Transposed 2x2x2x1 convolution onto the existing finer map (minkunet.py:107-113): every coarse
// row c scatters in[c] @ W[k] (+shift, ReLU) to its child child[k][c].  Each fine row has exactly
// one parent, so every output row is written once -- no accumulation, no atomics.
template <int COUT>
__global__ void __launch_bounds__(kSimtThreads)
k_conv_up(const sps_conv_args a, const int kc) {
  constexpr int TPV = COUT / 8;
  constexpr int VB = kSimtThreads / TPV;
  extern __shared__ __align__(16) float w_s[];  // [kc][cin][COUT]
  const int tid = threadIdx.x;
  const int cg = tid / VB, vl = tid % VB;
  const int n_in = *a.n_out;   // UP mode: rows iterated = coarse rows
  const int cin = a.cin;
  const int slab = cin * COUT;
  const int ntiles = (n_in + VB - 1) / VB;
  const bool resident = kc >= 8;
  bool loaded = false;
  float sh[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) sh[c] = a.shift ? __ldg(a.shift + cg * 8 + c) : 0.f;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int v = tile * VB + vl;
    const bool active = v < n_in;
    const float* row = a.in + (int64_t)(active ? v : 0) * a.in_ld;
    for (int kbase = 0; kbase < 8; kbase += kc) {
      const int kn = min(kc, 8 - kbase);
      if (!(resident && loaded)) {
        __syncthreads();
        const float4* src = reinterpret_cast<const float4*>(a.weight + (int64_t)kbase * slab);
        float4* dst = reinterpret_cast<float4*>(w_s);
        for (int i = tid; i < kn * slab / 4; i += kSimtThreads) dst[i] = __ldg(src + i);
        __syncthreads();
        loaded = true;
      }
      if (!active) continue;
      for (int kk = 0; kk < kn; ++kk) {
        const int f = __ldg(a.map + (int64_t)(kbase + kk) * a.map_ld + v);
        if (f < 0) continue;
        float acc[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[c] = sh[c];
        const float* wk = w_s + kk * slab + cg * 8;
        for (int ci = 0; ci < cin; ci += 4) {
          const float4 x4 = __ldg(reinterpret_cast<const float4*>(row + ci));
          const float xs[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 w0 = *reinterpret_cast<const float4*>(wk + (ci + j) * COUT);
            const float4 w1 = *reinterpret_cast<const float4*>(wk + (ci + j) * COUT + 4);
            acc[0] = fmaf(xs[j], w0.x, acc[0]); acc[1] = fmaf(xs[j], w0.y, acc[1]);
            acc[2] = fmaf(xs[j], w0.z, acc[2]); acc[3] = fmaf(xs[j], w0.w, acc[3]);
            acc[4] = fmaf(xs[j], w1.x, acc[4]); acc[5] = fmaf(xs[j], w1.y, acc[5]);
            acc[6] = fmaf(xs[j], w1.z, acc[6]); acc[7] = fmaf(xs[j], w1.w, acc[7]);
          }
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          if (a.relu) acc[c] = fmaxf(acc[c], 0.f);
          if (a.round_out) {
            uint32_t r;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(acc[c]));
            acc[c] = __uint_as_float(r);
          }
        }
        float* o = a.out + (int64_t)f * a.out_ld + cg * 8;
        *reinterpret_cast<float4*>(o) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        *reinterpret_cast<float4*>(o + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
      }
    }
  }
}

template <int COUT>
static int launch_up(const sps_conv_args& a, cudaStream_t st) {
  constexpr int VB = kSimtThreads / (COUT / 8);
  const int slab = a.cin * COUT * 4;
  int kc = kSimtSmemBytes / slab;
  if (kc > 8) kc = 8;
  if (kc < 1) return SPS_ERR_UNSUPPORTED;
  static unsigned long long attr_done = 0;   // bit per device
  SPS_CUDA_CHECK(ensure_dynamic_smem(k_conv_up<COUT>, kSimtSmemBytes, &attr_done));
  int64_t tiles = (a.n_out_max + VB - 1) / VB;
  if (tiles < 1) tiles = 1;
  const int grid = (int)(tiles < 148 * 8 ? tiles : 148 * 8);
  k_conv_up<COUT><<<grid, kSimtThreads, (size_t)kc * slab, st>>>(a, kc);
  SPS_CUDA_CHECK(cudaGetLastError());
  return SPS_OK;
}

// 1x1 convolution with a narrow output (the `final` 8 -> 1 layer used as a standalone module,
// minkunet.py:152-158): one thread per row.
__global__ void k_linear_narrow(const sps_conv_args a) {
  const int n = *a.n_out;
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < n; v += gridDim.x * blockDim.x) {
    const float* row = a.in + (int64_t)v * a.in_ld;
    for (int c = 0; c < a.cout; ++c) {
      float acc = a.shift ? __ldg(a.shift + c) : 0.f;
      for (int ci = 0; ci < a.cin; ++ci) acc = fmaf(__ldg(row + ci), __ldg(a.weight + ci * a.cout + c), acc);
      if (a.res) acc += __ldg(a.res + (int64_t)v * a.res_ld + c);
      if (a.relu) acc = fmaxf(acc, 0.f);
      a.out[(int64_t)v * a.out_ld + c] = acc;
    }
  }
}

int conv_dispatch(const sps_conv_args& a, cudaStream_t st);

int conv_simt(const sps_conv_args& a, cudaStream_t st) {
  if (a.mode == SPS_CONV_NBR && a.K == 1 && !a.map && a.cout < 8 && !a.in2 && !a.head_out && a.out) {
    int64_t g = (a.n_out_max + 255) / 256;
    k_linear_narrow<<<(int)(g < 1 ? 1 : g > 148 * 8 ? 148 * 8 : g), 256, 0, st>>>(a);
    SPS_CUDA_CHECK(cudaGetLastError());
    return SPS_OK;
  }
  if (a.mode == SPS_CONV_UP) {
    switch (a.cout) {
      case 8: return launch_up<8>(a, st);
      case 16: return launch_up<16>(a, st);
      case 32: return launch_up<32>(a, st);
      case 64: return launch_up<64>(a, st);
      default: return SPS_ERR_UNSUPPORTED;
    }
  }
  switch (a.cout) {
    case 8: return launch_simt<8>(a, st);
    case 16: return launch_simt<16>(a, st);
    case 32: return launch_simt<32>(a, st);
    case 64: return launch_simt<64>(a, st);
    default: return SPS_ERR_UNSUPPORTED;
  }
}

}  // namespace sps

extern "C" int sps_conv_fwd(const sps_conv_args* a, void* stream) {
  if (!a || !a->in || !a->weight || !a->n_out || (!a->out && !a->head_out)) return SPS_ERR_BAD_ARG;
  if (a->mode != SPS_CONV_NBR && a->mode != SPS_CONV_UP) return SPS_ERR_BAD_ARG;
  if (a->backend < SPS_BACKEND_AUTO || a->backend > SPS_BACKEND_F16) return SPS_ERR_BAD_ARG;
  if ((a->io_dtype != SPS_IO_F32 && a->io_dtype != SPS_IO_F16) || (a->flags & ~(SPS_CONV_FOLD_LO | SPS_CONV_OUT_SPLIT | SPS_CONV_MAP_PARENT)) ||
      a->cin_split < 0 || (a->cin_split && a->cin_split >= a->cin))
    return SPS_ERR_BAD_ARG;
  if (a->mode == SPS_CONV_UP && (!a->map || a->K != 8 || !a->out || a->in2 || a->res || a->head_out || (a->cin & 3)))
    return SPS_ERR_BAD_ARG;
  if (a->mode == SPS_CONV_NBR && !a->map && !a->tile_slices && a->K != 1) return SPS_ERR_BAD_ARG;
  const int ea = a->io_dtype == SPS_IO_F16 ? 8 : 4;   // elements per 16 bytes of a row
  if (a->cin < 1 || (a->cin != 1 && a->cin % ea) || (a->in_ld % ea && a->cin != 1)) return SPS_ERR_BAD_ARG;
  if (a->in2 && (!a->weight2 || a->cin2 % ea || a->in2_ld % ea)) return SPS_ERR_BAD_ARG;
  const bool narrow = a->K == 1 && !a->map && a->cout < 8 && a->io_dtype == SPS_IO_F32;   // scalar kernel, no vector alignment needed
  if (narrow) return sps::conv_dispatch(*a, (cudaStream_t)stream);
  if (a->out && a->out_ld % ea) return SPS_ERR_BAD_ARG;
  if (a->res && a->res_ld % ea) return SPS_ERR_BAD_ARG;
  if (a->head_out && (a->cout != 8 || !a->head_w)) return SPS_ERR_BAD_ARG;
  if (a->n_out_max < 0) return SPS_ERR_BAD_ARG;
  const uintptr_t al = (uintptr_t)a->weight | (uintptr_t)a->out | (uintptr_t)a->in2 | (uintptr_t)a->weight2 |
                       (uintptr_t)a->res | (a->cin != 1 ? (uintptr_t)a->in : 0);
  if (al & 15) return SPS_ERR_BAD_ARG;  // 16-byte vector access everywhere
  return sps::conv_dispatch(*a, (cudaStream_t)stream);
}
