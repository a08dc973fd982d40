/*
 * gs_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C CPU restatement of the forward render path of
 * dcaustin33/intro_to_gaussian_splatting (CPU/torch path, the parity target):
 *   GaussianScene.preprocess      splat/gaussian_scene.py:70-144
 *   GaussianScene.render_image    splat/gaussian_scene.py:200-238
 *   GaussianScene.render_tile     splat/gaussian_scene.py:173-198
 *   GaussianScene.render_pixel    splat/gaussian_scene.py:146-171
 * and of the math they call in splat/utils.py / splat/gaussians.py (cited per function).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library, and only as the checker or the timed CPU baseline.  The product
 * (libgsb_b200.so) never links, loads or calls it.
 *
 * PINNING: the reference has no golden vectors or tests of its own, so
 * this restatement is pinned against outputs of the reference itself, run in the build
 * container by tests/golden/make_golden.py (fixtures committed under tests/golden/) and
 * live by tests/test_oracle_vs_reference.py whenever /root/reference is present:
 * every PreprocessedScene field bit-exact (sigmoid: <= 1 ulp), tile membership identical,
 * pixels within 1e-6 of GaussianScene.render_image.
 *
 * Arithmetic: fp32, every operation individually rounded EXCEPT where torch's CPU kernels
 * were probed to fuse (marked FMA below); build with -ffp-contract=off so the compiler adds
 * no contraction of its own.  The exact op order is SURVEY.md Appendix A.
 *
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/gsb.h" /* struct layouts only (GsbCamera, GsbParams): shared so that tests
                               feed both sides the same bytes */

#define FMA(a, b, c) fmaf((a), (b), (c))

static inline uint32_t f2u(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
}

/* row-vector transform, one output column: [x y z 1] @ M[:, j]
 * (points_homogeneous @ view_matrix: splat/utils.py:305-307, splat/gaussian_scene.py:82-90,
 * splat/utils.py:335).  torch's (N,4)@(4,4) sgemm was probed to be an FMA chain in k order. */
static inline float rowvec_col(float x, float y, float z, const float* M, int j) {
  float t = x * M[0 * 4 + j];
  t = FMA(y, M[1 * 4 + j], t);
  t = FMA(z, M[2 * 4 + j], t);
  t = FMA(1.0f, M[3 * 4 + j], t);
  return t;
}

/* 3-D covariance: splat/gaussians.py:54-69 (F.normalize, build_rotation, R@S, M@M^T),
 * splat/utils.py:132-155 (build_rotation normalises a second time). */
static void covariance_3d(const float* q_in, const float* s, float cov[9]) {
  float q0 = q_in[0], q1 = q_in[1], q2 = q_in[2], q3 = q_in[3];
  /* F.normalize(p=2, dim=1, eps=1e-12): x / max(||x||, eps) */
  float n = sqrtf(((q0 * q0 + q1 * q1) + q2 * q2) + q3 * q3);
  float dn = n > 1e-12f ? n : 1e-12f;
  q0 = q0 / dn; q1 = q1 / dn; q2 = q2 / dn; q3 = q3 / dn;
  /* build_rotation: norm = sqrt(r0*r0 + r1*r1 + r2*r2 + r3*r3); q = r / norm */
  float n2 = sqrtf(((q0 * q0 + q1 * q1) + q2 * q2) + q3 * q3);
  float r = q0 / n2, x = q1 / n2, y = q2 / n2, z = q3 / n2;
  float R[9];
  R[0] = 1.0f - 2.0f * (y * y + z * z);
  R[1] = 2.0f * (x * y - r * z);
  R[2] = 2.0f * (x * z + r * y);
  R[3] = 2.0f * (x * y + r * z);
  R[4] = 1.0f - 2.0f * (x * x + z * z);
  R[5] = 2.0f * (y * z - r * x);
  R[6] = 2.0f * (x * z - r * y);
  R[7] = 2.0f * (y * z + r * x);
  R[8] = 1.0f - 2.0f * (x * x + y * y);
  /* scale_rotation = R @ diag(s): batched bmm, unfused; the zero terms add exactly */
  float M[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) M[i * 3 + j] = R[i * 3 + j] * s[j];
  /* covariance = M @ M^T: batched bmm, unfused, k order */
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      cov[i * 3 + j] = (M[i * 3 + 0] * M[j * 3 + 0] + M[i * 3 + 1] * M[j * 3 + 1]) + M[i * 3 + 2] * M[j * 3 + 2];
}

static inline float clampf(float v, float lo, float hi) {
  /* torch.clamp(x, lo, hi) = min(max(x, lo), hi); NaN propagates */
  if (v != v) return v;
  float t = v < lo ? lo : v;
  return t > hi ? hi : t;
}

/* EWA 2-D covariance: splat/utils.py:320-354 via GaussianScene.get_2d_covariance
 * (splat/gaussian_scene.py:53-68).  NOTE the reference passes tan_fovX/tan_fovY by keyword,
 * so limx uses tan_fovX as written at utils.py:336.
 * (J @ W @ cov3d @ W.T @ J^T)[:2,:2], left to right; broadcast-(3,3) products are FMA chains,
 * batched x batched products are unfused (probed). */
static void covariance_2d(const GsbCamera* cam, const GsbParams* prm, float vx, float vy, float vz,
                          const float cov3[9], float out[4]) {
  const float* V = cam->world2view;
  float limx = prm->fov_clamp * cam->tan_fovx;
  float limy = prm->fov_clamp * cam->tan_fovy;
  float x = vx / vz, y = vy / vz, z = vz;
  x = clampf(x, -limx, limx) * z;
  y = clampf(y, -limy, limy) * z;
  float z2 = z * z; /* z**2 */
  float J[9] = {0};
  J[0] = cam->f_x / z;
  J[2] = -(cam->f_x * x) / z2;
  J[4] = cam->f_y / z;
  J[5] = -(cam->f_y * y) / z2;
  /* W = extrinsic[:3,:3].T  => W[l][k] = V[k][l];  W.T[l][k] = V[l][k] */
  float T1[9], T2[9], T3[9];
  for (int i = 0; i < 3; ++i)
    for (int k = 0; k < 3; ++k) { /* J @ W : FMA chain */
      float t = J[i * 3 + 0] * V[k * 4 + 0];
      t = FMA(J[i * 3 + 1], V[k * 4 + 1], t);
      t = FMA(J[i * 3 + 2], V[k * 4 + 2], t);
      T1[i * 3 + k] = t;
    }
  for (int i = 0; i < 3; ++i)
    for (int k = 0; k < 3; ++k) /* @ cov3d : unfused */
      T2[i * 3 + k] = (T1[i * 3 + 0] * cov3[0 * 3 + k] + T1[i * 3 + 1] * cov3[1 * 3 + k]) + T1[i * 3 + 2] * cov3[2 * 3 + k];
  for (int i = 0; i < 3; ++i)
    for (int k = 0; k < 3; ++k) { /* @ W.T : FMA chain */
      float t = T2[i * 3 + 0] * V[0 * 4 + k];
      t = FMA(T2[i * 3 + 1], V[1 * 4 + k], t);
      t = FMA(T2[i * 3 + 2], V[2 * 4 + k], t);
      T3[i * 3 + k] = t;
    }
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 2; ++j) /* @ J^T : unfused;  J^T[k][j] = J[j][k] */
      out[i * 2 + j] = (T3[i * 3 + 0] * J[j * 3 + 0] + T3[i * 3 + 1] * J[j * 3 + 1]) + T3[i * 3 + 2] * J[j * 3 + 2];
}

/* reference tile grid: len(range(0, W - T, T)) (splat/gaussian_scene.py:208,:214);
 * full_cover: ceil(W/T). */
static int grid_dim(int extent, int T, int full_cover) {
  if (full_cover) return (extent + T - 1) / T;
  int span = extent - T;
  return span <= 0 ? 0 : (span + T - 1) / T;
}

void orc_grid(const GsbCamera* cam, const GsbParams* prm, int32_t* ntx, int32_t* nty) {
  *ntx = grid_dim(cam->width, prm->tile_size, prm->full_cover);
  *nty = grid_dim(cam->height, prm->tile_size, prm->full_cover);
}

/* literal restatement of the tile masks (splat/gaussian_scene.py:209-220):
 *   (min <= t_min + T) & (max >= t_min)   for t_min = 0, T, 2T, ...
 * The mask is an interval in t; return its first/last tile index (lo > hi = empty). */
static void tile_interval(float mn, float mx, int T, int ntiles, int32_t* lo, int32_t* hi) {
  int first = -1, last = -2;
  for (int t = 0; t < ntiles; ++t) {
    float t_min = (float)(t * T);
    float t_hi = (float)(t * T + T);
    if (mn <= t_hi && mx >= t_min) {
      if (first < 0) first = t;
      last = t;
    }
  }
  if (first < 0) { *lo = 0; *hi = -1; } else { *lo = first; *hi = last; }
}

/* GaussianScene.preprocess up to (not including) the depth sort, in Gaussian-index order.
 * All outputs have n rows; rows with in_view==0 are zero-filled.  Returns M. */
int64_t orc_project(const GsbCamera* cam, const GsbParams* prm, int64_t n, const float* xyz,
                    const float* scales, const float* quats, const float* colors,
                    const float* opacity_logit, uint8_t* in_view, float* depth, float* pxy,
                    float* cov2d, float* conic, float* radius, float* bbox, float* sig_op,
                    int32_t* rect, uint32_t* count) {
  int32_t ntx, nty;
  orc_grid(cam, prm, &ntx, &nty);
  const int T = prm->tile_size;
  int64_t m = 0;
  float Wf = (float)cam->width, Hf = (float)cam->height;
#pragma omp parallel for schedule(static) reduction(+ : m)
  for (int64_t i = 0; i < n; ++i) {
    float x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
    /* in_view_frustum: splat/utils.py:293-310 */
    float vz = rowvec_col(x, y, z, cam->world2view, 2);
    int keep = vz >= prm->minimum_z;
    if (in_view) in_view[i] = (uint8_t)keep;
    if (!keep) {
      if (depth) depth[i] = 0;
      if (pxy) pxy[2 * i] = pxy[2 * i + 1] = 0;
      if (cov2d) memset(cov2d + 4 * i, 0, 16);
      if (conic) memset(conic + 4 * i, 0, 16);
      if (radius) radius[i] = 0;
      if (bbox) memset(bbox + 4 * i, 0, 16);
      if (sig_op) sig_op[i] = 0;
      if (rect) { rect[4 * i] = 0; rect[4 * i + 1] = -1; rect[4 * i + 2] = 0; rect[4 * i + 3] = -1; }
      if (count) count[i] = 0;
      continue;
    }
    m += 1;
    float vx = rowvec_col(x, y, z, cam->world2view, 0);
    float vy = rowvec_col(x, y, z, cam->world2view, 1);
    /* NDC + ndc2Pix: splat/gaussian_scene.py:87-97, splat/utils.py:313-317 */
    float cx = rowvec_col(x, y, z, cam->full_proj, 0);
    float cy = rowvec_col(x, y, z, cam->full_proj, 1);
    float cw = rowvec_col(x, y, z, cam->full_proj, 3);
    float ndx = cx / cw, ndy = cy / cw;
    float px = ((ndx + 1.0f) * (Wf - 1.0f)) * 0.5f;
    float py = ((ndy + 1.0f) * (Hf - 1.0f)) * 0.5f;
    float c3[9], c2[4];
    covariance_3d(quats + 4 * i, scales + 3 * i, c3);
    covariance_2d(cam, prm, vx, vy, vz, c3, c2);
    float a = c2[0], b = c2[1], c = c2[2], d = c2[3];
    /* compute_inverted_covariance: splat/utils.py:368-393 */
    float det = a * d - b * c;
    det = det != det ? det : (det < prm->det_min ? prm->det_min : det);
    float i00 = d / det, i11 = a / det, i01 = (-b) / det, i10 = (-c) / det;
    /* compute_extent_and_radius: splat/utils.py:409-423 */
    float mid = 0.5f * (a + d);
    float det2 = a * d - b * b;
    float im = mid * mid - det2;
    /* torch.max over cat([im, 0.1], dim=1): NaN propagates */
    float mv = (im != im) ? im : (im > prm->lambda_floor ? im : prm->lambda_floor);
    float sq = sqrtf(mv);
    float l1 = mid + sq, l2 = mid - sq;
    float lm = (l1 != l1 || l2 != l2) ? NAN : (l1 > l2 ? l1 : l2);
    float rad = ceilf(prm->sigma_extent * sqrtf(lm));
    /* bbox: splat/gaussian_scene.py:108-111 */
    float mnx = floorf(px - rad), mny = floorf(py - rad);
    float mxx = ceilf(px + rad), mxy = ceilf(py + rad);
    if (depth) depth[i] = vz;
    if (pxy) { pxy[2 * i] = px; pxy[2 * i + 1] = py; }
    if (cov2d) { cov2d[4 * i] = a; cov2d[4 * i + 1] = b; cov2d[4 * i + 2] = c; cov2d[4 * i + 3] = d; }
    if (conic) { conic[4 * i] = i00; conic[4 * i + 1] = i01; conic[4 * i + 2] = i10; conic[4 * i + 3] = i11; }
    if (radius) radius[i] = rad;
    if (bbox) { bbox[4 * i] = mnx; bbox[4 * i + 1] = mny; bbox[4 * i + 2] = mxx; bbox[4 * i + 3] = mxy; }
    /* torch.sigmoid(opacity): splat/gaussian_scene.py:143 */
    if (sig_op) sig_op[i] = 1.0f / (1.0f + expf(-opacity_logit[i]));
    int32_t tx0, tx1, ty0, ty1;
    tile_interval(mnx, mxx, T, ntx, &tx0, &tx1);
    tile_interval(mny, mxy, T, nty, &ty0, &ty1);
    uint32_t cnt = (tx1 >= tx0 && ty1 >= ty0) ? (uint32_t)(tx1 - tx0 + 1) * (uint32_t)(ty1 - ty0 + 1) : 0u;
    if (rect) {
      if (cnt) { rect[4 * i] = tx0; rect[4 * i + 1] = tx1; rect[4 * i + 2] = ty0; rect[4 * i + 3] = ty1; }
      else { rect[4 * i] = 0; rect[4 * i + 1] = -1; rect[4 * i + 2] = 0; rect[4 * i + 3] = -1; }
    }
    if (count) count[i] = cnt;
    (void)colors;
  }
  return m;
}

/* stable LSD radix sort of (u64 key, u32 payload); restates "sort by key, ties keep input order". */
void orc_sort_pairs(int64_t n, const uint64_t* keys_in, const uint32_t* vals_in, uint64_t* keys_out,
                    uint32_t* vals_out) {
  if (n <= 0) return;
  uint64_t* ka = (uint64_t*)malloc((size_t)n * 8);
  uint32_t* va = (uint32_t*)malloc((size_t)n * 4);
  uint64_t* kb = (uint64_t*)malloc((size_t)n * 8);
  uint32_t* vb = (uint32_t*)malloc((size_t)n * 4);
  memcpy(ka, keys_in, (size_t)n * 8);
  memcpy(va, vals_in, (size_t)n * 4);
  for (int pass = 0; pass < 8; ++pass) {
    int64_t hist[256] = {0};
    int sh = pass * 8;
    for (int64_t i = 0; i < n; ++i) hist[(ka[i] >> sh) & 255]++;
    int skip = 0;
    for (int b = 0; b < 256; ++b)
      if (hist[b] == n) skip = 1;
    if (skip) continue;
    int64_t off[256], s = 0;
    for (int b = 0; b < 256; ++b) { off[b] = s; s += hist[b]; }
    for (int64_t i = 0; i < n; ++i) {
      int64_t p = off[(ka[i] >> sh) & 255]++;
      kb[p] = ka[i];
      vb[p] = va[i];
    }
    uint64_t* tk = ka; ka = kb; kb = tk;
    uint32_t* tv = va; va = vb; vb = tv;
  }
  memcpy(keys_out, ka, (size_t)n * 8);
  memcpy(vals_out, va, (size_t)n * 4);
  free(ka); free(va); free(kb); free(vb);
}

/* depth order of the in-view Gaussians: torch.argsort(points_view[:,2]) with ties in index
 * order (splat/gaussian_scene.py:117; stable by contract, see header).  Returns M. */
int64_t orc_depth_order(int64_t n, const uint8_t* in_view, const float* depth, int32_t* order) {
  int64_t m = 0;
  for (int64_t i = 0; i < n; ++i) m += in_view[i] ? 1 : 0;
  if (m == 0) return 0;
  uint64_t* k = (uint64_t*)malloc((size_t)m * 8);
  uint32_t* v = (uint32_t*)malloc((size_t)m * 4);
  uint64_t* ko = (uint64_t*)malloc((size_t)m * 8);
  uint32_t* vo = (uint32_t*)malloc((size_t)m * 4);
  int64_t j = 0;
  for (int64_t i = 0; i < n; ++i)
    if (in_view[i]) { k[j] = f2u(depth[i]); v[j] = (uint32_t)i; ++j; } /* z >= 0.2 > 0: uint order == float order */
  orc_sort_pairs(m, k, v, ko, vo);
  for (int64_t t = 0; t < m; ++t) order[t] = (int32_t)vo[t];
  free(k); free(v); free(ko); free(vo);
  return m;
}

/* tile-instance keys in Gaussian-index order, rect scanned row-major:
 * key = (ty*ntx + tx) << 32 | float_as_uint(z_view); payload = Gaussian index.
 * keys == NULL: count only.  Returns K. */
int64_t orc_emit_keys(int64_t n, const uint8_t* in_view, const float* depth, const int32_t* rect,
                      int32_t ntx, uint64_t* keys, uint32_t* payload) {
  int64_t k = 0;
  for (int64_t i = 0; i < n; ++i) {
    if (!in_view[i]) continue;
    int32_t tx0 = rect[4 * i], tx1 = rect[4 * i + 1], ty0 = rect[4 * i + 2], ty1 = rect[4 * i + 3];
    if (tx1 < tx0 || ty1 < ty0) continue;
    uint32_t zb = f2u(depth[i]);
    for (int32_t ty = ty0; ty <= ty1; ++ty)
      for (int32_t tx = tx0; tx <= tx1; ++tx) {
        if (keys) {
          keys[k] = ((uint64_t)(uint32_t)(ty * ntx + tx) << 32) | zb;
          payload[k] = (uint32_t)i;
        }
        ++k;
      }
  }
  return k;
}

/* [start,end) of each tile in the sorted key array; empty tiles = (0,0). */
void orc_tile_ranges(int64_t k, const uint64_t* sorted_keys, int64_t ntiles, uint32_t* ranges) {
  memset(ranges, 0, (size_t)ntiles * 8);
  for (int64_t i = 0; i < k; ++i) {
    uint32_t t = (uint32_t)(sorted_keys[i] >> 32);
    if (i == 0 || (uint32_t)(sorted_keys[i - 1] >> 32) != t) ranges[2 * (int64_t)t] = (uint32_t)i;
    if (i == k - 1 || (uint32_t)(sorted_keys[i + 1] >> 32) != t) ranges[2 * (int64_t)t + 1] = (uint32_t)(i + 1);
  }
}

/* Compositing, REF_CPU semantics: render_pixel (splat/gaussian_scene.py:146-171) over the
 * per-tile depth-ordered list, compute_gaussian_weight (splat/utils.py:357-365).
 *   difference = mean - pixel;  power = ((-0.5*difference) @ inv) @ difference^T
 *   w = exp(power);  alpha = w * sigmoid(sigmoid_opacity)   <- second sigmoid, :164
 *   test = T*(1-alpha);  if test < min_weight: return (Gaussian NOT added)
 *   colour += (T*alpha)*c;  T = test
 * Records are indexed by Gaussian index through `payload`.  Output (H,W,3), image[y][x][c];
 * pixels outside the tile grid stay 0.  *steps receives the executed (pixel,Gaussian) steps;
 * steps_per_pixel (optional, H*W, caller-zeroed) the per-pixel count (work-distribution analysis). */
void orc_composite(const GsbCamera* cam, const GsbParams* prm, const uint32_t* ranges,
                   const uint32_t* payload, const float* pxy, const float* conic, const float* colors,
                   const float* sig_op, float* image, int64_t* steps, int32_t* steps_per_pixel) {
  int32_t ntx, nty;
  orc_grid(cam, prm, &ntx, &nty);
  const int T = prm->tile_size, W = cam->width, H = cam->height;
  memset(image, 0, (size_t)W * H * 3 * sizeof(float));
  int64_t total = 0;
  const float minw = prm->min_weight;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : total)
  for (int tile = 0; tile < ntx * nty; ++tile) {
    uint32_t s = ranges[2 * tile], e = ranges[2 * tile + 1];
    if (e <= s) continue;
    uint32_t L = e - s;
    /* stage the tile list once: (mx,my, i00,i01,i10,i11, op2, r,g,b) */
    float* rec = (float*)malloc((size_t)L * 10 * sizeof(float));
    for (uint32_t j = 0; j < L; ++j) {
      uint32_t g = payload[s + j];
      float* r = rec + 10 * j;
      r[0] = pxy[2 * g]; r[1] = pxy[2 * g + 1];
      r[2] = conic[4 * g]; r[3] = conic[4 * g + 1]; r[4] = conic[4 * g + 2]; r[5] = conic[4 * g + 3];
      r[6] = 1.0f / (1.0f + expf(-sig_op[g])); /* torch.sigmoid(opacities[point_idx]) */
      r[7] = colors[3 * g]; r[8] = colors[3 * g + 1]; r[9] = colors[3 * g + 2];
    }
    int x0 = (tile % ntx) * T, y0 = (tile / ntx) * T;
    for (int py = y0; py < y0 + T && py < H; ++py)
      for (int px = x0; px < x0 + T && px < W; ++px) {
        float Tw = 1.0f, cr = 0.f, cg = 0.f, cb = 0.f;
        float fx = (float)px, fy = (float)py;
        uint32_t j = 0;
        for (; j < L; ++j) {
          const float* r = rec + 10 * j;
          float dx = r[0] - fx, dy = r[1] - fy;
          float hx = -0.5f * dx, hy = -0.5f * dy;
          /* probed against torch CPU on 20 000 ill-conditioned cases (bit-exact): the (1,2)@(2,2)
             product is an FMA chain in k order, the (1,2)@(2,1) product is unfused */
          float u0 = FMA(hy, r[4], hx * r[2]);
          float u1 = FMA(hy, r[5], hx * r[3]);
          float power = u0 * dx + u1 * dy;
          float w = expf(power);
          float alpha = w * r[6];
          float test = Tw * (1.0f - alpha);
          if (test < minw) break;
          float ta = Tw * alpha;
          cr += ta * r[7]; cg += ta * r[8]; cb += ta * r[9];
          Tw = test;
        }
        total += (j < L) ? (int64_t)j + 1 : (int64_t)L;
        if (steps_per_pixel) steps_per_pixel[(size_t)py * W + px] = (int32_t)((j < L) ? j + 1 : L);
        float* o = image + ((size_t)py * W + px) * 3;
        o[0] = cr; o[1] = cg; o[2] = cb;
      }
    free(rec);
  }
  if (steps) *steps = total;
}

/* Compositing, REF_CU semantics: the render_tile kernel of splat/c/render.cu:21-87 restated
 * over per-tile lists (rows are depth-sorted; `payload` indexes rows):
 *   inclusive per-pixel bbox test (:55-60); mean TRUNCATED to int (:8-9 int params);
 *   power = dx*a*dx + 2*dx*dy*b + dy*dy*c with inv[0],inv[1],inv[3] (:17,:66-68);
 *   alpha = min(alpha_max, opacity*w) (:70-71); break when T*(1-alpha) < min_weight (:72-76). */
void orc_composite_cu(int32_t W, int32_t H, const GsbParams* prm, int32_t ntx, int32_t nty,
                      const uint32_t* ranges, const uint32_t* payload, const float* means,
                      const float* conic, const float* colors, const float* opacity,
                      const float* min_x, const float* max_x, const float* min_y, const float* max_y,
                      float* image) {
  const int T = prm->tile_size;
  memset(image, 0, (size_t)W * H * 3 * sizeof(float));
#pragma omp parallel for schedule(dynamic, 1)
  for (int tile = 0; tile < ntx * nty; ++tile) {
    uint32_t s = ranges[2 * tile], e = ranges[2 * tile + 1];
    int x0 = (tile % ntx) * T, y0 = (tile / ntx) * T;
    for (int py = y0; py < y0 + T && py < H; ++py)
      for (int px = x0; px < x0 + T && px < W; ++px) {
        float Tw = 1.0f, cr = 0.f, cg = 0.f, cb = 0.f;
        for (uint32_t j = s; j < e; ++j) {
          uint32_t g = payload[j];
          if (!((float)px >= min_x[g] && (float)px <= max_x[g])) continue;
          if (!((float)py >= min_y[g] && (float)py <= max_y[g])) continue;
          int mx = (int)means[2 * g], my = (int)means[2 * g + 1];
          float dx = (float)(px - mx), dy = (float)(py - my);
          float a = conic[4 * g], b = conic[4 * g + 1], c = conic[4 * g + 3];
          float power = dx * a * dx + 2 * dx * dy * b + dy * dy * c;
          float w = expf(-0.5f * power);
          float al = opacity[g] * w;
          al = al < prm->alpha_max ? al : prm->alpha_max;
          float test = Tw * (1 - al);
          if (test < prm->min_weight) break;
          cr += Tw * al * colors[3 * g];
          cg += Tw * al * colors[3 * g + 1];
          cb += Tw * al * colors[3 * g + 2];
          Tw = test;
        }
        float* o = image + ((size_t)py * W + px) * 3;
        o[0] = cr; o[1] = cg; o[2] = cb;
      }
  }
}

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void orc_set_num_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}
