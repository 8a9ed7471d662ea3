// Context management, workspace carve-up, status plumbing and small helpers of the C ABI.
#include <cstdio>
#include <cstring>
#include "ctx.h"

namespace sps {

static thread_local char g_err[512] = "";

int set_cuda_error(cudaError_t e, const char* what) {
  snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
  return SPS_ERR_CUDA;
}

// Bump carve of the workspace; with base == nullptr it only measures.
struct Carver {
  char* base;
  size_t off = 0;
  explicit Carver(void* b) : base((char*)b) {}
  template <class T>
  T* take(size_t count) {
    off = (off + 255) & ~size_t(255);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += count * sizeof(T);
    return p;
  }
};

static size_t carve(sps_ctx* c, void* base, int64_t max_points) {
  Carver cv(base);
  const int64_t N = max_points;
  const int64_t ld = (N + 31) & ~int64_t(31);
  c->max_points = N;
  c->ld = ld;
  c->counts = cv.take<int32_t>(64);  // counts[5], status, ticket, n_dev share one cache line group
  c->status = c->counts + 8;
  c->ticket = reinterpret_cast<uint32_t*>(c->counts + 9);
  c->n_dev = c->counts + 10;
  c->staging = cv.take<float>((size_t)N * 8);
  c->scores = cv.take<float>(N);
  c->table_cap = table_capacity(N);
  c->table = cv.take<Slot>(c->table_cap);
  c->slot_of = cv.take<uint32_t>(N);
  c->rank = cv.take<int32_t>(N);
  c->block_sums = cv.take<int32_t>(N / kScanBlock + 2);
  c->inv = cv.take<int32_t>(N);
  for (int L = 0; L < SPS_NUM_LEVELS; ++L) {
    c->keys[L] = cv.take<unsigned long long>(N);
    c->nbr3[L] = cv.take<int32_t>((size_t)81 * ld);
    c->parent[L] = (L < SPS_NUM_LEVELS - 1) ? cv.take<int32_t>(N) : nullptr;
    c->child[L] = (L > 0) ? cv.take<int32_t>((size_t)8 * ld) : nullptr;
  }
  c->nbr5 = cv.take<int32_t>((size_t)125 * ld);
  for (int b = 0; b < sps_ctx::NBUF; ++b) c->buf[b] = cv.take<float>((size_t)N * kBufWidth[b]);
  return (cv.off + 255) & ~size_t(255);
}

int conv_simt(const sps_conv_args& a, cudaStream_t st);
int conv_umma(const sps_conv_args& a, cudaStream_t st);
bool conv_umma_supports(const sps_conv_args& a);
static int g_backend = 0;  // 0 auto, 1 fp32 CUDA-core, 2 tcgen05
int conv_dispatch(const sps_conv_args& a, cudaStream_t st) {
  if (g_backend == 2) return conv_umma_supports(a) ? conv_umma(a, st) : SPS_ERR_UNSUPPORTED;
  if (g_backend == 0 && conv_umma_supports(a)) return conv_umma(a, st);
  return conv_simt(a, st);
}

}  // namespace sps

using namespace sps;

extern "C" int sps_set_conv_backend(int backend) {
  if (backend < 0 || backend > 2) return SPS_ERR_BAD_ARG;
  g_backend = backend;
  return SPS_OK;
}
extern "C" const char* sps_version(void) { return "sps_b200 0.1 (sm_100a)"; }
extern "C" const char* sps_last_error(void) { return g_err; }

extern "C" size_t sps_workspace_bytes(int64_t max_points) {
  if (max_points < 1) max_points = 1;
  sps_ctx tmp;
  return carve(&tmp, nullptr, max_points);
}

extern "C" int sps_ctx_create(sps_ctx** out, void* d_workspace, size_t workspace_bytes, int64_t max_points) {
  if (!out || !d_workspace || max_points < 1 || max_points > (int64_t(1) << 28)) return SPS_ERR_BAD_ARG;
  if ((uintptr_t)d_workspace & 255) return SPS_ERR_BAD_ARG;
  sps_ctx* c = new sps_ctx();
  const size_t need = carve(c, d_workspace, max_points);
  if (need > workspace_bytes) { delete c; return SPS_ERR_CAPACITY; }
  c->base = (char*)d_workspace;
  c->bytes = workspace_bytes;
  cudaError_t e = cudaMemset(c->counts, 0, 64 * sizeof(int32_t));
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { delete c; return set_cuda_error(e, "cudaMemset(ctx scalars)"); }
  *out = c;
  return SPS_OK;
}

extern "C" int sps_ctx_destroy(sps_ctx* ctx) {
  delete ctx;
  return SPS_OK;
}

extern "C" int sps_ctx_status(sps_ctx* ctx, void* stream) {
  if (!ctx) return SPS_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  int32_t word = 0;
  SPS_CUDA_CHECK(cudaMemcpyAsync(&word, ctx->status, sizeof(word), cudaMemcpyDeviceToHost, st));
  SPS_CUDA_CHECK(cudaStreamSynchronize(st));
  if (word) {
    SPS_CUDA_CHECK(cudaMemsetAsync(ctx->status, 0, sizeof(int32_t), st));
    SPS_CUDA_CHECK(cudaStreamSynchronize(st));
  }
  if (word & kStatusRange) return SPS_ERR_COORD_RANGE;
  if (word & kStatusCapacity) return SPS_ERR_CAPACITY;
  return SPS_OK;
}

extern "C" int sps_ctx_level(sps_ctx* ctx, int level, sps_level_view* v) {
  if (!ctx || !v || level < 0 || level >= SPS_NUM_LEVELS) return SPS_ERR_BAD_ARG;
  v->keys = (const uint64_t*)ctx->keys[level];
  v->count = ctx->counts + level;
  v->nbr3 = ctx->nbr3[level];
  v->nbr5 = level == 0 ? ctx->nbr5 : nullptr;
  v->parent = ctx->parent[level];
  v->child = ctx->child[level];
  v->ld = ctx->ld;
  return SPS_OK;
}

extern "C" const int32_t* sps_ctx_inverse_map(sps_ctx* ctx) { return ctx ? ctx->inv : nullptr; }

extern "C" int sps_memcpy_d2h(void* h_dst, const void* d_src, size_t bytes, void* stream) {
  if (!h_dst || !d_src) return SPS_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  SPS_CUDA_CHECK(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, st));
  SPS_CUDA_CHECK(cudaStreamSynchronize(st));
  return SPS_OK;
}
extern "C" int sps_memcpy_h2d(void* d_dst, const void* h_src, size_t bytes, void* stream) {
  if (!d_dst || !h_src) return SPS_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  SPS_CUDA_CHECK(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, st));
  SPS_CUDA_CHECK(cudaStreamSynchronize(st));
  return SPS_OK;
}
