/* ORACLE (test infrastructure, not product code) -- canonical-arithmetic restatement.
 *
 * Plain C restatement of the reference's geometric guidance path
 *   utils/pano_utils.py:117-161   project_feats_to_equirectangular
 *   utils/pano_utils.py:164-242   equirectangular_to_pointcloud
 *   utils/point_cloud_utils.py:90-183  project_to_feat
 * with the transcendental functions (sin/cos tables, atan2, acos, pow(.,0.5))
 * replaced by a CANONICAL definition that is built only from IEEE-754
 * correctly rounded float32 operations (+ - * / sqrt fma), so that a GPU
 * implementation can reproduce every bit.  Everything else follows the
 * reference op for op (same association order, one rounding per op).
 *
 * Canonical definitions (the bit authority; see DESIGN.md "canonical arithmetic"):
 *   sin/cos tables : float32( libm double sin/cos( double(angle_f32) ) )
 *   pow(x, 0.5)    : sqrtf(x)
 *   atan2(y, x)    : canon_atan2f below (degree-8 odd polynomial, <= ~1.5 ulp)
 *   acos(q)        : canon_acosf below  (degree-5 asin polynomial, <= ~1.5 ulp)
 * The literal libm version lives in oracle/ref_numpy.py; tests cross-check the
 * two (pixel indices may differ for ~1e-5 of the points: 1-ulp boundary cases).
 *
 * Parity status: pinned against the reference's known-answer tests through
 * ref_numpy (tests/test_oracle_kat.py, tests/test_oracle_exact.py); TF itself
 * cannot run in this image, so TF's own bits are unpinned.
 *
 * Build: gcc -O2 -ffp-contract=off -mfma -shared -fPIC  (see oracle/Makefile).
 * -ffp-contract=off is REQUIRED: a*b+c must round twice unless fmaf is written.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define CANON_PI_HI 0x1.921fb6p+1f      /* float32(pi)           */
#define CANON_PI_LO -0x1.777a5cp-24f    /* float32(pi - PI_HI)   */
#define CANON_PIO2_HI 0x1.921fb6p+0f    /* float32(pi/2)         */
#define CANON_PIO2_LO -0x1.777a5cp-25f  /* float32(pi/2 - PIO2_HI) */

/* oracle/gen_coeffs.py output (degree 8 / degree 5 in s). */
static const float kATAN[9] = {
    -0x1.555540p-2f, 0x1.999266p-3f, -0x1.242616p-3f, 0x1.c0c7a4p-4f, -0x1.58b92ep-4f,
    0x1.dc8cccp-5f,  -0x1.0126a6p-5f, 0x1.695cf0p-7f, -0x1.dcc7b0p-10f};
static const float kASIN[6] = {0x1.55554cp-3f, 0x1.333888p-4f, 0x1.6cacd0p-5f,
                               0x1.04a2f6p-5f, 0x1.d816aep-7f, 0x1.33b2a6p-5f};

float se3ds_oracle_atan2f(float y, float x) {
  float ax = fabsf(x), ay = fabsf(y);
  float mx = ax > ay ? ax : ay;
  float mn = ax > ay ? ay : ax;
  float t = (mx == 0.0f) ? 0.0f : mn / mx;
  float s = t * t;
  float p = kATAN[8];
  for (int i = 7; i >= 0; --i) p = fmaf(p, s, kATAN[i]);
  float r = fmaf(p * s, t, t);
  if (ay > ax) r = (CANON_PIO2_HI - r) + CANON_PIO2_LO;
  if (x < 0.0f) r = (CANON_PI_HI - r) + CANON_PI_LO;
  if (y < 0.0f) r = -r;
  return r;
}

float se3ds_oracle_acosf(float q) {
  float a = fabsf(q);
  int small = a <= 0.5f;
  float z = (1.0f - a) * 0.5f;
  float s = small ? q * q : z;
  float xa = small ? q : sqrtf(z);
  float p = kASIN[5];
  for (int i = 4; i >= 0; --i) p = fmaf(p, s, kASIN[i]);
  float r = fmaf(p * s, xa, xa); /* asin(xa) */
  if (small) return (CANON_PIO2_HI - r) + CANON_PIO2_LO;
  float w = r + r;
  return q > 0.0f ? w : (CANON_PI_HI - w) + CANON_PI_LO;
}

/* tf.linspace in float32: exact end points, start + delta*i inside
 * (utils/pano_utils.py:211-219). */
static void linspace_f32(float start, float stop, int n, float* out) {
  if (n == 1) { out[0] = start; return; }
  float delta = (stop - start) / (float)(n - 1);
  out[0] = start;
  for (int i = 1; i < n - 1; ++i) out[i] = start + delta * (float)i;
  out[n - 1] = stop;
}

/* Angle tables of equirectangular_to_pointcloud: elevation (H), heading (W) and
 * their canonical sin/cos. */
void se3ds_oracle_tables(int H, int W, float* elev, float* head, float* sin_e, float* cos_e,
                         float* sin_h, float* cos_h) {
  const double pi = 3.141592653589793;  /* np.pi */
  double hp = 0.5 * pi / (double)H;
  linspace_f32((float)hp, (float)(pi - hp), H, elev);
  linspace_f32((float)(1.5 * pi - hp), (float)(-0.5 * pi + hp), W, head);
  for (int r = 0; r < H; ++r) { sin_e[r] = (float)sin((double)elev[r]); cos_e[r] = (float)cos((double)elev[r]); }
  for (int c = 0; c < W; ++c) { sin_h[c] = (float)sin((double)head[c]); cos_h[c] = (float)cos((double)head[c]); }
}

/* utils/pano_utils.py:220-236.  depth (N,H,W) -> xyz1 (N,4,HW) planar, valid (N,HW). */
void se3ds_oracle_unproject(const float* depth, int N, int H, int W, float depth_scale, float* xyz1,
                            uint8_t* valid) {
  float* tab = (float*)malloc(sizeof(float) * (size_t)(3 * H + 3 * W));
  float *elev = tab, *sin_e = tab + H, *cos_e = tab + 2 * H;
  float *head = tab + 3 * H, *sin_h = head + W, *cos_h = head + 2 * W;
  se3ds_oracle_tables(H, W, elev, head, sin_e, cos_e, sin_h, cos_h);
  size_t HW = (size_t)H * W;
  for (int b = 0; b < N; ++b)
    for (int r = 0; r < H; ++r)
      for (int c = 0; c < W; ++c) {
        size_t i = (size_t)r * W + c;
        float d = depth[b * HW + i];
        float m = (d > 0.0f && d < 1.0f) ? 1.0f : 0.0f;
        float rad = (d * depth_scale) * m;
        float t = rad * sin_e[r];
        float* o = xyz1 + (size_t)b * 4 * HW;
        o[i] = t * cos_h[c];
        o[HW + i] = t * sin_h[c];
        o[2 * HW + i] = rad * cos_e[r];
        o[3 * HW + i] = 1.0f;
        if (valid) valid[b * HW + i] = (uint8_t)(m != 0.0f);
      }
  free(tab);
}

/* tf.cast(float32 -> int32) on x86: trunc toward zero, NaN / out of range -> INT_MIN. */
static int32_t cast_i32(float v) {
  if (!(v > -2147483904.0f && v < 2147483648.0f)) return INT32_MIN;
  return (int32_t)v;
}

static float div_no_nan(float a, float b) { return b == 0.0f ? 0.0f : a / b; }

/* utils/pano_utils.py:139-156: (x,y,z) -> pseudo-perspective (rad*u, rad*v, rad). */
void se3ds_oracle_pseudo_perspective(float x, float y, float z, float* px, float* py, float* pz) {
  const float two_pi = (float)(2 * 3.141592653589793);
  const float pi_f = (float)3.141592653589793;
  const float pi15 = (float)(1.5 * 3.141592653589793);
  float rad = sqrtf((x * x + y * y) + z * z);
  float h = pi15 - se3ds_oracle_atan2f(y, x);
  h = h + two_pi * (h <= 0.0f ? 1.0f : 0.0f);
  h = h - two_pi * (h > two_pi ? 1.0f : 0.0f);
  float e = se3ds_oracle_acosf(div_no_nan(z, rad));
  *px = rad * ((h / two_pi) * 2.0f - 1.0f);
  *py = rad * ((e / pi_f) * 2.0f - 1.0f);
  *pz = rad;
}

/* utils/point_cloud_utils.py:127-153 for one point: returns the in-image flat
 * pixel index (row*W+col) or -1 if the point is invalid. */
static int32_t pixel_of(float px, float py, float pz, int H, int W, int feat_valid) {
  float vx = div_no_nan(px, pz), vy = div_no_nan(py, pz);
  int32_t col = cast_i32((vx + 1.0f) / 2.0f * (float)W);
  int32_t row = cast_i32((vy + 1.0f) / 2.0f * (float)H);
  int ok = col >= 0 && col < W && row >= 0 && row < H && pz > 0.0f && feat_valid;
  return ok ? row * W + col : -1;
}

/* The splat (utils/point_cloud_utils.py:127-183) over an (N,4,M) cloud.
 *   mode 0: coords are cartesian -> project_feats_to_equirectangular (pano_utils.py:117-161)
 *   mode 1: coords are already "transformed_coords" -> project_to_feat itself
 * Outputs (any may be NULL):
 *   depth_out (N,H,W)  clip(zbuf,0,scale)/scale           feat_out (N,H,W,C)
 *   zbuf_out  (N,H,W)  raw scatter-min result             winner_out (N,H,W) int32
 *   flat_out  (N,M)    global flat index before tolerance (0 = rejected, as the reference)
 *   kept_out  (N,M)    global flat index after tolerance  rad_out (N,M) per point depth
 *   valid_out (N,M)    1 if the point passed the validity tests of :137-149
 * winner = lowest point index m among the valid points of the pixel whose depth equals
 * the pixel minimum and is <= depth_scale; -1 if none. */
void se3ds_oracle_splat(const float* coords, const float* feats, int N, long long M, int C, int H,
                        int W, int mode, float void_in, float void_out, float depth_scale,
                        float* depth_out, float* feat_out, float* zbuf_out, int32_t* winner_out,
                        int32_t* flat_out, int32_t* kept_out, float* rad_out, uint8_t* valid_out) {
  size_t HW = (size_t)H * W, P = (size_t)N * HW, K = (size_t)N * (size_t)M;
  float* zbuf = (float*)malloc(sizeof(float) * P);
  float* fbuf = (float*)malloc(sizeof(float) * P * (size_t)C);
  int32_t* flat = (int32_t*)malloc(sizeof(int32_t) * K);
  float* rad = (float*)malloc(sizeof(float) * K);
  uint8_t* isvalid = (uint8_t*)malloc(K);
  for (size_t i = 0; i < P; ++i) zbuf[i] = depth_scale;
  for (size_t i = 0; i < P * (size_t)C; ++i) fbuf[i] = void_out;
  if (winner_out) for (size_t i = 0; i < P; ++i) winner_out[i] = -1;

  for (int b = 0; b < N; ++b) {
    const float* cb = coords + (size_t)b * 4 * (size_t)M;
    for (long long m = 0; m < M; ++m) {
      size_t k = (size_t)b * (size_t)M + (size_t)m;
      float px, py, pz;
      if (mode == 0) se3ds_oracle_pseudo_perspective(cb[m], cb[M + m], cb[2 * M + m], &px, &py, &pz);
      else { px = cb[m]; py = cb[M + m]; pz = cb[2 * M + m]; }
      int fv = 1;
      for (int c = 0; c < C; ++c) fv &= (feats[k * (size_t)C + c] != void_in);
      int32_t pix = pixel_of(px, py, pz, H, W, fv);
      isvalid[k] = pix >= 0;
      flat[k] = pix >= 0 ? (int32_t)((size_t)b * HW + (size_t)pix) : 0;
      rad[k] = pz;
      if (pz < zbuf[flat[k]]) zbuf[flat[k]] = pz; /* scatter_min over ALL points (bin = index 0) */
    }
  }
  if (winner_out) {
    for (size_t k = 0; k < K; ++k)
      if (isvalid[k] && rad[k] == zbuf[flat[k]] && rad[k] <= depth_scale && winner_out[flat[k]] < 0)
        winner_out[flat[k]] = (int32_t)(k % (size_t)M);
  }
  for (size_t k = 0; k < K; ++k) {
    int keep = rad[k] < zbuf[flat[k]] + 0.1f;
    int32_t f2 = keep ? flat[k] : 0;
    if (kept_out) kept_out[k] = f2;
    for (int c = 0; c < C; ++c) {
      float v = feats[k * (size_t)C + c];
      if (v > fbuf[(size_t)f2 * C + c]) fbuf[(size_t)f2 * C + c] = v;
    }
  }
  if (depth_out)
    for (size_t i = 0; i < P; ++i) {
      float z = zbuf[i];
      z = z < 0.0f ? 0.0f : (z > depth_scale ? depth_scale : z);
      depth_out[i] = z / depth_scale;
    }
  if (feat_out) memcpy(feat_out, fbuf, sizeof(float) * P * (size_t)C);
  if (zbuf_out) memcpy(zbuf_out, zbuf, sizeof(float) * P);
  if (flat_out) memcpy(flat_out, flat, sizeof(int32_t) * K);
  if (rad_out) memcpy(rad_out, rad, sizeof(float) * K);
  if (valid_out) memcpy(valid_out, isvalid, K);
  free(zbuf); free(fbuf); free(flat); free(rad); free(isvalid);
}

/* Full SE(3) extension (no reference counterpart; see include/se3ds_geom.h se3ds_reproject_se3):
 * rotates J clouds (J,4,M) by rot (J,3,3) row-major, row-wise fma(r2, z, fma(r1, y, r0 * x)). */
void se3ds_oracle_rotate(const float* coords, const float* rot, int J, long long M, float* out) {
  for (int j = 0; j < J; ++j) {
    const float* c = coords + (size_t)j * 4 * (size_t)M;
    const float* r = rot + (size_t)j * 9;
    float* o = out + (size_t)j * 4 * (size_t)M;
    for (long long m = 0; m < M; ++m) {
      float x = c[m], y = c[M + m], z = c[2 * M + m];
      o[m] = fmaf(r[2], z, fmaf(r[1], y, r[0] * x));
      o[M + m] = fmaf(r[5], z, fmaf(r[4], y, r[3] * x));
      o[2 * M + m] = fmaf(r[8], z, fmaf(r[7], y, r[6] * x));
      o[3 * M + m] = c[3 * M + m];
    }
  }
}

/* Exhaustive-check helper used by tests: evaluates atan2/acos on arrays. */
void se3ds_oracle_atan2f_array(const float* y, const float* x, float* out, long long n) {
  for (long long i = 0; i < n; ++i) out[i] = se3ds_oracle_atan2f(y[i], x[i]);
}
void se3ds_oracle_acosf_array(const float* q, float* out, long long n) {
  for (long long i = 0; i < n; ++i) out[i] = se3ds_oracle_acosf(q[i]);
}
