/*
 * me_cpu.c -- C/OpenMP restatement of MinkowskiEngine's CPU backend for SPSModel.forward.
 *
 * TEST INFRASTRUCTURE / CPU BASELINE ONLY (see oracle/sps_oracle.py header): used by tests/ as a
 * fast second oracle and by bench.py's cpu_baseline / --impl reference legs.  Never imported by
 * the product (sps_b200/).  PARITY UNPINNED: MinkowskiEngine itself is not available offline;
 * this file follows its published CPU algorithm at the reference's call sites:
 *   - voxelise: sequential hash insert, first-occurrence order  (src/sps/models/models.py:21-25)
 *   - stride maps: floor to the new tensor stride, sequential insert (minkunet.py:64-104)
 *   - kernel maps: one hash probe per (output voxel, kernel offset), OpenMP over voxels,
 *     compacted into per-offset (in,out) pair lists, cached per level and shared by layers
 *   - convolution: for every kernel offset, gather -> small GEMM -> scatter-add, fp32
 *     (minkunet.py:55-158; ME BasicBlock mirrored at c_ws/src/mapmos/scripts/minkunet.py:31-82)
 *   - separate BatchNorm (eval) and ReLU passes, concat copies (minkunet.py:161-219)
 *   - slice = F[inverse_mapping], sigmoid (models.py:28-29)
 * It is validated against oracle/sps_oracle.py (tests/test_oracle.py) before being timed.
 *
 * Weight blob = tensors of sps_oracle.layer_shapes() in order, fp32; every "bn" entry expands
 * to weight, bias, running_mean, running_var.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <omp.h>

typedef struct { int32_t c[5]; } coord_t;

/* ---------------------------------------------------------------- coordinate hash ------ */
typedef struct {
  int32_t* slot;   /* -1 = empty, else row index */
  uint32_t mask;
  const coord_t* rows;
} chash_t;

static inline uint32_t hash5(const int32_t* c) {
  uint64_t h = 1469598103934665603ull;
  for (int d = 0; d < 5; ++d) { h ^= (uint32_t)c[d]; h *= 1099511628211ull; }
  h ^= h >> 29; h *= 0xbf58476d1ce4e5b9ull; h ^= h >> 32;
  return (uint32_t)h;
}
static inline int same5(const int32_t* a, const int32_t* b) {
  return a[0] == b[0] && a[1] == b[1] && a[2] == b[2] && a[3] == b[3] && a[4] == b[4];
}
static void chash_init(chash_t* h, int64_t n, const coord_t* rows) {
  uint32_t cap = 1024;
  while ((int64_t)cap < 2 * n) cap <<= 1;
  h->slot = (int32_t*)malloc((size_t)cap * sizeof(int32_t));
  memset(h->slot, 0xff, (size_t)cap * sizeof(int32_t));
  h->mask = cap - 1;
  h->rows = rows;
}
static inline int32_t chash_find(const chash_t* h, const int32_t* c) {
  uint32_t s = hash5(c) & h->mask;
  for (;;) {
    int32_t r = h->slot[s];
    if (r < 0) return -1;
    if (same5(h->rows[r].c, c)) return r;
    s = (s + 1) & h->mask;
  }
}
/* find or insert row `row` (whose coordinates are already stored in rows[row]) */
static inline int32_t chash_insert(chash_t* h, const int32_t* c, int32_t row) {
  uint32_t s = hash5(c) & h->mask;
  for (;;) {
    int32_t r = h->slot[s];
    if (r < 0) { h->slot[s] = row; return row; }
    if (same5(h->rows[r].c, c)) return r;
    s = (s + 1) & h->mask;
  }
}

/* ---------------------------------------------------------------- level structure ------ */
#define NLEV 5
typedef struct {
  int64_t n;            /* voxels */
  coord_t* rows;
  chash_t hash;
  int32_t* parent;      /* [n] row in level+1 */
  int32_t* koff;        /* [n] 2x2x2x1 offset index inside the parent */
  /* 3x3x3x3 kernel map as per-offset pair lists */
  int64_t cnt3[81]; int32_t* in3[81]; int32_t* out3[81];
} level_t;

typedef struct {
  level_t lv[NLEV];
  int64_t cnt5[125]; int32_t* in5[125]; int32_t* out5[125];
  int32_t* inv; int64_t npts;
} maps_t;

static inline int32_t floordiv(int32_t a, int32_t m) {
  int32_t q = a / m;
  if ((a % m) && ((a < 0) != (m < 0))) --q;
  return q;
}

/* kernel map for an odd hyper-cube kernel: probes parallel over voxels, then compaction */
static void build_kmap(const level_t* L, const int ks[4], int s, int64_t* cnt, int32_t** in, int32_t** out) {
  const int K = ks[0] * ks[1] * ks[2] * ks[3];
  const int64_t n = L->n;
  int32_t* nbr = (int32_t*)malloc((size_t)K * (n ? n : 1) * sizeof(int32_t));
#pragma omp parallel for schedule(static)
  for (int64_t o = 0; o < n; ++o) {
    const int32_t* c = L->rows[o].c;
    for (int k = 0; k < K; ++k) {
      int r = k;
      const int i0 = r % ks[0]; r /= ks[0];
      const int i1 = r % ks[1]; r /= ks[1];
      const int i2 = r % ks[2]; r /= ks[2];
      const int i3 = r;
      int32_t q[5] = {c[0], c[1] + (i0 - ks[0] / 2) * s, c[2] + (i1 - ks[1] / 2) * s, c[3] + (i2 - ks[2] / 2) * s,
                      c[4] + (i3 - ks[3] / 2)};
      nbr[(size_t)k * n + o] = chash_find(&L->hash, q);
    }
  }
#pragma omp parallel for schedule(dynamic, 1)
  for (int k = 0; k < K; ++k) {
    int64_t m = 0;
    for (int64_t o = 0; o < n; ++o) m += nbr[(size_t)k * n + o] >= 0;
    cnt[k] = m;
    in[k] = (int32_t*)malloc((size_t)(m ? m : 1) * sizeof(int32_t));
    out[k] = (int32_t*)malloc((size_t)(m ? m : 1) * sizeof(int32_t));
    m = 0;
    for (int64_t o = 0; o < n; ++o) {
      const int32_t i = nbr[(size_t)k * n + o];
      if (i >= 0) { in[k][m] = i; out[k][m] = (int32_t)o; ++m; }
    }
  }
  free(nbr);
}

static void maps_build(maps_t* M, const float* pts, int64_t n, int64_t ld, float vs) {
  memset(M, 0, sizeof(*M));
  M->npts = n;
  M->inv = (int32_t*)malloc((size_t)(n ? n : 1) * sizeof(int32_t));
  /* level 0: quantise (fp32 division, floor) + sequential insert */
  level_t* L0 = &M->lv[0];
  L0->rows = (coord_t*)malloc((size_t)(n ? n : 1) * sizeof(coord_t));
  chash_init(&L0->hash, n, L0->rows);
  const float q[5] = {1.0f, vs, vs, vs, 1.0f};
  int64_t V = 0;
  for (int64_t p = 0; p < n; ++p) {
    coord_t c;
    for (int d = 0; d < 5; ++d) c.c[d] = (int32_t)floorf(pts[p * ld + d] / q[d]);
    L0->rows[V] = c;
    const int32_t r = chash_insert(&L0->hash, c.c, (int32_t)V);
    if (r == V) ++V;
    M->inv[p] = r;
  }
  L0->n = V;
  /* strided levels */
  for (int l = 1; l < NLEV; ++l) {
    level_t* F = &M->lv[l - 1];
    level_t* C = &M->lv[l];
    const int32_t m = 1 << l, s = 1 << (l - 1);
    C->rows = (coord_t*)malloc((size_t)(F->n ? F->n : 1) * sizeof(coord_t));
    chash_init(&C->hash, F->n, C->rows);
    F->parent = (int32_t*)malloc((size_t)(F->n ? F->n : 1) * sizeof(int32_t));
    F->koff = (int32_t*)malloc((size_t)(F->n ? F->n : 1) * sizeof(int32_t));
    int64_t Vc = 0;
    for (int64_t f = 0; f < F->n; ++f) {
      const int32_t* fc = F->rows[f].c;
      coord_t c;
      c.c[0] = fc[0]; c.c[4] = fc[4];
      for (int d = 1; d < 4; ++d) c.c[d] = floordiv(fc[d], m) * m;
      C->rows[Vc] = c;
      const int32_t r = chash_insert(&C->hash, c.c, (int32_t)Vc);
      if (r == Vc) ++Vc;
      F->parent[f] = r;
      F->koff[f] = (fc[1] - c.c[1]) / s + 2 * ((fc[2] - c.c[2]) / s) + 4 * ((fc[3] - c.c[3]) / s);
    }
    C->n = Vc;
  }
  const int k5[4] = {5, 5, 5, 1}, k3[4] = {3, 3, 3, 3};
  build_kmap(L0, k5, 1, M->cnt5, M->in5, M->out5);
  for (int l = 0; l < NLEV; ++l) build_kmap(&M->lv[l], k3, 1 << l, M->lv[l].cnt3, M->lv[l].in3, M->lv[l].out3);
}

static void maps_free(maps_t* M) {
  for (int l = 0; l < NLEV; ++l) {
    level_t* L = &M->lv[l];
    free(L->rows); free(L->hash.slot); free(L->parent); free(L->koff);
    for (int k = 0; k < 81; ++k) { free(L->in3[k]); free(L->out3[k]); }
  }
  for (int k = 0; k < 125; ++k) { free(M->in5[k]); free(M->out5[k]); }
  free(M->inv);
}

/* ---------------------------------------------------------------- dense helpers -------- */
/* out[o] += in[i] @ W  for every pair; out rows are unique within one offset -> race free */
static void gather_gemm_scatter(const float* in, int cin, const float* W, int cout, float* out, const int32_t* ii,
                                const int32_t* oo, int64_t m) {
#pragma omp parallel for schedule(static)
  for (int64_t p = 0; p < m; ++p) {
    const float* x = in + (size_t)ii[p] * cin;
    float* y = out + (size_t)oo[p] * cout;
    for (int ci = 0; ci < cin; ++ci) {
      const float a = x[ci];
      const float* w = W + (size_t)ci * cout;
      for (int co = 0; co < cout; ++co) y[co] += a * w[co];
    }
  }
}

static float* zeros(int64_t n, int c) { return (float*)calloc((size_t)(n ? n : 1) * c, sizeof(float)); }

static float* conv_pairs(const float* in, int cin, const float* W, int cout, int64_t n_out, int K, const int64_t* cnt,
                         int32_t* const* ii, int32_t* const* oo) {
  float* out = zeros(n_out, cout);
  for (int k = 0; k < K; ++k)
    if (cnt[k]) gather_gemm_scatter(in, cin, W + (size_t)k * cin * cout, cout, out, ii[k], oo[k], cnt[k]);
  return out;
}

/* stride-2 conv (kernel 2x2x2x1): per offset k, pairs (f -> parent[f]); children of one parent have
 * distinct k, so every offset is race free */
static float* conv_down(const float* in, int c_in, const float* W, int c_out, const level_t* F, int64_t n_out) {
  float* out = zeros(n_out, c_out);
  for (int k = 0; k < 8; ++k) {
#pragma omp parallel for schedule(static)
    for (int64_t f = 0; f < F->n; ++f) {
      if (F->koff[f] != k) continue;
      const float* x = in + (size_t)f * c_in;
      float* y = out + (size_t)F->parent[f] * c_out;
      const float* Wk = W + (size_t)k * c_in * c_out;
      for (int ci = 0; ci < c_in; ++ci) {
        const float a = x[ci];
        for (int co = 0; co < c_out; ++co) y[co] += a * Wk[(size_t)ci * c_out + co];
      }
    }
  }
  return out;
}

/* transposed conv onto the existing finer map: out[f] = in[parent[f]] @ W[k(f)] */
static float* conv_up(const float* in, int c_in, const float* W, int c_out, const level_t* F) {
  float* out = zeros(F->n, c_out);
  for (int k = 0; k < 8; ++k) {
#pragma omp parallel for schedule(static)
    for (int64_t f = 0; f < F->n; ++f) {
      if (F->koff[f] != k) continue;
      const float* x = in + (size_t)F->parent[f] * c_in;
      float* y = out + (size_t)f * c_out;
      const float* Wk = W + (size_t)k * c_in * c_out;
      for (int ci = 0; ci < c_in; ++ci) {
        const float a = x[ci];
        for (int co = 0; co < c_out; ++co) y[co] += a * Wk[(size_t)ci * c_out + co];
      }
    }
  }
  return out;
}

static float* dense_mm(const float* in, int cin, const float* W, int cout, int64_t n) {
  float* out = zeros(n, cout);
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < n; ++r) {
    const float* x = in + (size_t)r * cin;
    float* y = out + (size_t)r * cout;
    for (int ci = 0; ci < cin; ++ci) {
      const float a = x[ci];
      for (int co = 0; co < cout; ++co) y[co] += a * W[(size_t)ci * cout + co];
    }
  }
  return out;
}

/* MinkowskiBatchNorm (eval) in place; bn = {weight, bias, mean, var} each [c] */
static void batchnorm(float* x, int64_t n, int c, const float* bn, int do_relu) {
  const float *g = bn, *b = bn + c, *mu = bn + 2 * c, *var = bn + 3 * c;
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < n; ++r)
    for (int j = 0; j < c; ++j) {
      float v = (x[(size_t)r * c + j] - mu[j]) / sqrtf(var[j] + 1e-5f) * g[j] + b[j];
      x[(size_t)r * c + j] = (do_relu && v < 0.f) ? 0.f : v;
    }
}

static float* concat(const float* a, int ca, const float* b, int cb, int64_t n) {
  float* out = (float*)malloc((size_t)(n ? n : 1) * (ca + cb) * sizeof(float));
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < n; ++r) {
    memcpy(out + (size_t)r * (ca + cb), a + (size_t)r * ca, (size_t)ca * sizeof(float));
    memcpy(out + (size_t)r * (ca + cb) + ca, b + (size_t)r * cb, (size_t)cb * sizeof(float));
  }
  return out;
}

typedef struct { const float* p; } blob_t;
static const float* take(blob_t* b, int64_t n) { const float* r = b->p; b->p += n; return r; }

/* BasicBlock: relu(bn2(conv2(relu(bn1(conv1(x))))) + (bn_ds(x @ Wds) | x)) */
static float* basic_block(const float* x, int cin, int cout, const level_t* L, blob_t* w) {
  const float* W1 = take(w, (int64_t)81 * cin * cout); const float* bn1 = take(w, 4 * cout);
  const float* W2 = take(w, (int64_t)81 * cout * cout); const float* bn2 = take(w, 4 * cout);
  float* h = conv_pairs(x, cin, W1, cout, L->n, 81, L->cnt3, L->in3, L->out3);
  batchnorm(h, L->n, cout, bn1, 1);
  float* o = conv_pairs(h, cout, W2, cout, L->n, 81, L->cnt3, L->in3, L->out3);
  free(h);
  batchnorm(o, L->n, cout, bn2, 0);
  float* res = NULL;
  const float* r = x;
  if (cin != cout) {
    const float* Wd = take(w, (int64_t)cin * cout); const float* bnd = take(w, 4 * cout);
    res = dense_mm(x, cin, Wd, cout, L->n);
    batchnorm(res, L->n, cout, bnd, 0);
    r = res;
  }
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < L->n * cout; ++i) { float v = o[i] + r[i]; o[i] = v < 0.f ? 0.f : v; }
  free(res);
  return o;
}

static const int PLANES[8] = {8, 16, 32, 64, 64, 32, 16, 8};

/* MinkUNetBase.forward (minkunet.py:161-219); returns logits [V0] */
static float* unet(const maps_t* M, const float* feat0, const float* blob) {
  blob_t w = {blob};
  const level_t* lv = M->lv;
  const float* W0 = take(&w, 125 * 1 * 8); const float* bn0 = take(&w, 4 * 8);
  float* out_p1 = conv_pairs(feat0, 1, W0, 8, lv[0].n, 125, M->cnt5, M->in5, M->out5);
  batchnorm(out_p1, lv[0].n, 8, bn0, 1);
  float* skips[4]; int skipc[4];
  skips[0] = out_p1; skipc[0] = 8;
  float* x = out_p1; int c = 8;
  for (int i = 0; i < 4; ++i) {
    const float* Wd = take(&w, (int64_t)8 * c * c); const float* bnd = take(&w, 4 * c);
    float* e = conv_down(x, c, Wd, c, &lv[i], lv[i + 1].n);
    batchnorm(e, lv[i + 1].n, c, bnd, 1);
    float* b = basic_block(e, c, PLANES[i], &lv[i + 1], &w);
    free(e);
    c = PLANES[i];
    x = b;
    if (i < 3) { skips[i + 1] = b; skipc[i + 1] = c; }
  }
  for (int i = 0; i < 4; ++i) {
    const int L = 3 - i, co = PLANES[4 + i];
    const float* Wu = take(&w, (int64_t)8 * c * co); const float* bnu = take(&w, 4 * co);
    float* u = conv_up(x, c, Wu, co, &lv[L]);
    batchnorm(u, lv[L].n, co, bnu, 1);
    free(x);
    float* cat = concat(u, co, skips[L], skipc[L], lv[L].n);
    free(u); free(skips[L]);
    float* b = basic_block(cat, co + skipc[L], co, &lv[L], &w);
    free(cat);
    x = b; c = co;
  }
  const float* Wf = take(&w, 8); const float* bf = take(&w, 1);
  float* logits = dense_mm(x, 8, Wf, 1, lv[0].n);
  for (int64_t i = 0; i < lv[0].n; ++i) logits[i] += bf[0];
  free(x);
  return logits;
}

/* ---------------------------------------------------------------- exported ------------- */
/* SPSModel.forward (models.py:20-30). counts: [5] voxels per level (may be NULL).
 * timings: [3] seconds for maps / unet / slice (may be NULL). */
int me_cpu_forward(const float* pts, int64_t n, int64_t ld, float vs, const float* blob, float* scores,
                   int nthreads, int64_t* counts, double* timings) {
  if (nthreads > 0) omp_set_num_threads(nthreads);
  maps_t M;
  double t0 = omp_get_wtime();
  maps_build(&M, pts, n, ld, vs);
  double t1 = omp_get_wtime();
  float* feat0 = (float*)malloc((size_t)(M.lv[0].n ? M.lv[0].n : 1) * sizeof(float));
  for (int64_t i = 0; i < M.lv[0].n; ++i) feat0[i] = 0.5f;  /* mean of the constant 0.5 features */
  float* logits = unet(&M, feat0, blob);
  double t2 = omp_get_wtime();
#pragma omp parallel for schedule(static)
  for (int64_t p = 0; p < n; ++p) scores[p] = 1.0f / (1.0f + expf(-logits[M.inv[p]]));
  double t3 = omp_get_wtime();
  if (counts) for (int l = 0; l < NLEV; ++l) counts[l] = M.lv[l].n;
  if (timings) { timings[0] = t1 - t0; timings[1] = t2 - t1; timings[2] = t3 - t2; }
  free(feat0); free(logits);
  maps_free(&M);
  return 0;
}

/* voxelisation only: coords int32 [n,5] (first V rows valid), inv [n]; returns V */
int64_t me_cpu_voxelize(const float* pts, int64_t n, int64_t ld, float vs, int32_t* coords, int32_t* inv) {
  coord_t* rows = (coord_t*)coords;
  chash_t h;
  chash_init(&h, n, rows);
  const float q[5] = {1.0f, vs, vs, vs, 1.0f};
  int64_t V = 0;
  for (int64_t p = 0; p < n; ++p) {
    coord_t c;
    for (int d = 0; d < 5; ++d) c.c[d] = (int32_t)floorf(pts[p * ld + d] / q[d]);
    rows[V] = c;
    const int32_t r = chash_insert(&h, c.c, (int32_t)V);
    if (r == V) ++V;
    inv[p] = r;
  }
  free(h.slot);
  return V;
}

int me_cpu_max_threads(void) { return omp_get_max_threads(); }
