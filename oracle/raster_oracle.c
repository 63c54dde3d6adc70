/*
 * oracle/raster_oracle.c -- TEST INFRASTRUCTURE ONLY (never imported by the product path).
 *
 * CPU restatement (plain C, fp32 arithmetic) of the reference Gaussian-splatting
 * rasterizer `submodules/diff-gaussian-rasterization` (DGR) of Open-DiffusionGS.
 * The reference has no CPU path (every stage is a __global__ kernel), so this file
 * re-states its algorithm stage by stage; each function cites the reference lines it
 * follows.  Parity pinning: the reference ships no golden vectors; this oracle is
 * pinned against outputs of the reference kernels themselves, compiled from
 * /root/reference into oracle/_ref (see oracle/build_ref.py) and run on a B200
 * (fixtures under tests/golden/, generator tests/golden/make_golden.py).
 *
 * Build: gcc -O2 -fopenmp -shared -fPIC -o liboracle.so raster_oracle.c -lm
 *        (+ -DORC_F64 -> liboracle64.so: every `float` becomes `double`; used only by the
 *         finite-difference self-check of the restated gradients, tests/test_oracle_cpu.py)
 *
 * Conventions: matrices are the 16-float arrays the reference receives, i.e. element
 * m[4*c + r] is row r / column c of the ordinary matrix (auxiliary.h:58-77).
 * Gradient accumulators of the blend backward are fp64 (the reference uses fp32
 * atomics in a non-deterministic order; fp64 is the order-free limit of that sum).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
typedef float f32_t; /* real fp32 even in the ORC_F64 build (sort keys) */
#ifdef ORC_F64 /* fp64 self-check build (finite-difference gradcheck of the restatement only) */
#define float double
#define sqrtf sqrt
#define expf exp
#define fminf fmin
#define fmaxf fmax
#define ceilf ceil
#endif

#define TILE 16 /* config.h:15-17 (BLOCK_X = BLOCK_Y = 16) */
#define NCH 3   /* config.h:15 */

static const float K_SH0 = 0.28209479177387814f;
static const float K_SH1 = 0.4886025119029199f;
static const float K_SH2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                               -1.0925484305920792f, 0.5462742152960396f};
static const float K_SH3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                               0.3731763325901154f,  -0.4570457994644658f, 1.445305721320277f,
                               -0.5900435899266435f};

/* ------------------------------------------------------------------------------------------ */
/* small helpers                                                                              */
/* ------------------------------------------------------------------------------------------ */

/* auxiliary.h:41-44 -- note the double literals: evaluated in fp64, rounded to fp32 */
static float ndc_to_pix(float v, int S) { return (float)(((v + 1.0) * S - 1.0) * 0.5); }

static int imin(int a, int b) { return a < b ? a : b; }
static int imax(int a, int b) { return a > b ? a : b; }

/* auxiliary.h:46-56 */
static void tile_rect(float px, float py, int radius, int gx, int gy, int* x0, int* y0, int* x1,
                      int* y1) {
  *x0 = imin(gx, imax(0, (int)((px - radius) / TILE)));
  *y0 = imin(gy, imax(0, (int)((py - radius) / TILE)));
  *x1 = imin(gx, imax(0, (int)((px + radius + TILE - 1) / TILE)));
  *y1 = imin(gy, imax(0, (int)((py + radius + TILE - 1) / TILE)));
}

/* auxiliary.h:58-77 */
static void xform43(const float* m, const float* p, float* o) {
  o[0] = m[0] * p[0] + m[4] * p[1] + m[8] * p[2] + m[12];
  o[1] = m[1] * p[0] + m[5] * p[1] + m[9] * p[2] + m[13];
  o[2] = m[2] * p[0] + m[6] * p[1] + m[10] * p[2] + m[14];
}
static void xform44(const float* m, const float* p, float* o) {
  o[0] = m[0] * p[0] + m[4] * p[1] + m[8] * p[2] + m[12];
  o[1] = m[1] * p[0] + m[5] * p[1] + m[9] * p[2] + m[13];
  o[2] = m[2] * p[0] + m[6] * p[1] + m[10] * p[2] + m[14];
  o[3] = m[3] * p[0] + m[7] * p[1] + m[11] * p[2] + m[15];
}

/* Real SH basis values b[0..n) and their gradient w.r.t. the unit direction (x,y,z).
 * Same basis/sign convention as forward.cu:20-71 / backward.cu:20-139. */
static int sh_basis(int deg, float x, float y, float z, float* b, float (*db)[3]) {
  int n = (deg + 1) * (deg + 1);
  for (int k = 0; k < n; k++) db[k][0] = db[k][1] = db[k][2] = 0.f;
  b[0] = K_SH0;
  if (deg > 0) {
    b[1] = -K_SH1 * y; db[1][1] = -K_SH1;
    b[2] = K_SH1 * z;  db[2][2] = K_SH1;
    b[3] = -K_SH1 * x; db[3][0] = -K_SH1;
  }
  if (deg > 1) {
    float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    b[4] = K_SH2[0] * xy;                   db[4][0] = K_SH2[0] * y;  db[4][1] = K_SH2[0] * x;
    b[5] = K_SH2[1] * yz;                   db[5][1] = K_SH2[1] * z;  db[5][2] = K_SH2[1] * y;
    b[6] = K_SH2[2] * (2.0f * zz - xx - yy);
    db[6][0] = K_SH2[2] * 2.f * -x; db[6][1] = K_SH2[2] * 2.f * -y; db[6][2] = K_SH2[2] * 2.f * 2.f * z;
    b[7] = K_SH2[3] * xz;                   db[7][0] = K_SH2[3] * z;  db[7][2] = K_SH2[3] * x;
    b[8] = K_SH2[4] * (xx - yy);            db[8][0] = K_SH2[4] * 2.f * x; db[8][1] = K_SH2[4] * 2.f * -y;
    if (deg > 2) {
      b[9] = K_SH3[0] * y * (3.0f * xx - yy);
      db[9][0] = K_SH3[0] * 3.f * 2.f * xy; db[9][1] = K_SH3[0] * 3.f * (xx - yy);
      b[10] = K_SH3[1] * xy * z;
      db[10][0] = K_SH3[1] * yz; db[10][1] = K_SH3[1] * xz; db[10][2] = K_SH3[1] * xy;
      b[11] = K_SH3[2] * y * (4.0f * zz - xx - yy);
      db[11][0] = K_SH3[2] * -2.f * xy; db[11][1] = K_SH3[2] * (-3.f * yy + 4.f * zz - xx);
      db[11][2] = K_SH3[2] * 4.f * 2.f * yz;
      b[12] = K_SH3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy);
      db[12][0] = K_SH3[3] * -3.f * 2.f * xz; db[12][1] = K_SH3[3] * -3.f * 2.f * yz;
      db[12][2] = K_SH3[3] * 3.f * (2.f * zz - xx - yy);
      b[13] = K_SH3[4] * x * (4.0f * zz - xx - yy);
      db[13][0] = K_SH3[4] * (-3.f * xx + 4.f * zz - yy); db[13][1] = K_SH3[4] * -2.f * xy;
      db[13][2] = K_SH3[4] * 4.f * 2.f * xz;
      b[14] = K_SH3[5] * z * (xx - yy);
      db[14][0] = K_SH3[5] * 2.f * xz; db[14][1] = K_SH3[5] * -2.f * yz; db[14][2] = K_SH3[5] * (xx - yy);
      b[15] = K_SH3[6] * x * (xx - 3.0f * yy);
      db[15][0] = K_SH3[6] * 3.f * (xx - yy); db[15][1] = K_SH3[6] * -3.f * 2.f * xy;
    }
  }
  return n;
}

/* Quaternion (r,x,y,z) -> rotation matrix Rq[row][col]; NOT re-normalised
 * (forward.cu:127: the normalisation is commented out in the reference). */
static void quat_to_rot(const float* q, float R[3][3]) {
  float r = q[0], x = q[1], y = q[2], z = q[3];
  R[0][0] = 1.f - 2.f * (y * y + z * z); R[0][1] = 2.f * (x * y - r * z); R[0][2] = 2.f * (x * z + r * y);
  R[1][0] = 2.f * (x * y + r * z); R[1][1] = 1.f - 2.f * (x * x + z * z); R[1][2] = 2.f * (y * z - r * x);
  R[2][0] = 2.f * (x * z - r * y); R[2][1] = 2.f * (y * z + r * x); R[2][2] = 1.f - 2.f * (x * x + y * y);
}

/* forward.cu:118-152: Sigma = (S Rq^T)^T (S Rq^T), six upper-triangular entries. */
static void cov3d_from_scale_rot(const float* s, float mod, const float* q, float* c6) {
  float R[3][3], M[3][3];
  quat_to_rot(q, R);
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) M[i][j] = (mod * s[i]) * R[j][i];
  int k = 0;
  for (int a = 0; a < 3; a++)
    for (int b = a; b < 3; b++) c6[k++] = M[0][a] * M[0][b] + M[1][a] * M[1][b] + M[2][a] * M[2][b];
}

/* A = J * Rw (2x3): the EWA screen-space Jacobian times the world->view rotation
 * (forward.cu:74-106; glm column-major literals decoded). */
typedef struct { float A[2][3]; float t[3]; float txtz, tytz; float limx, limy; } Ewa;

static void ewa_setup(const float* mean, float fx, float fy, float tanx, float tany,
                      const float* view, Ewa* e) {
  xform43(view, mean, e->t);
  e->limx = 1.3f * tanx; e->limy = 1.3f * tany;
  e->txtz = e->t[0] / e->t[2]; e->tytz = e->t[1] / e->t[2];
  e->t[0] = fminf(e->limx, fmaxf(-e->limx, e->txtz)) * e->t[2];
  e->t[1] = fminf(e->limy, fmaxf(-e->limy, e->tytz)) * e->t[2];
  float tz = e->t[2];
  float j00 = fx / tz, j02 = -(fx * e->t[0]) / (tz * tz);
  float j11 = fy / tz, j12 = -(fy * e->t[1]) / (tz * tz);
  for (int k = 0; k < 3; k++) {
    /* Rw[r][k] = view[4k + r] */
    e->A[0][k] = view[4 * k + 0] * j00 + view[4 * k + 2] * j02; /* (+ view[4k+1]*0) */
    e->A[1][k] = view[4 * k + 1] * j11 + view[4 * k + 2] * j12;
  }
}

static void sym6_to_mat(const float* c, float V[3][3]) {
  V[0][0] = c[0]; V[0][1] = V[1][0] = c[1]; V[0][2] = V[2][0] = c[2];
  V[1][1] = c[3]; V[1][2] = V[2][1] = c[4]; V[2][2] = c[5];
}

/* cov2D = A V A^T (+0.3 dilation on the diagonal), forward.cu:99-112 */
static void cov2d(const Ewa* e, const float* c6, float* a, float* b, float* c) {
  float V[3][3], VA0[3], VA1[3];
  sym6_to_mat(c6, V);
  for (int k = 0; k < 3; k++) {
    VA0[k] = V[k][0] * e->A[0][0] + V[k][1] * e->A[0][1] + V[k][2] * e->A[0][2];
    VA1[k] = V[k][0] * e->A[1][0] + V[k][1] * e->A[1][1] + V[k][2] * e->A[1][2];
  }
  *a = e->A[0][0] * VA0[0] + e->A[0][1] * VA0[1] + e->A[0][2] * VA0[2] + 0.3f;
  *b = e->A[0][0] * VA1[0] + e->A[0][1] * VA1[1] + e->A[0][2] * VA1[2];
  *c = e->A[1][0] * VA1[0] + e->A[1][1] * VA1[1] + e->A[1][2] * VA1[2] + 0.3f;
}

/* ------------------------------------------------------------------------------------------ */
/* K1: per-Gaussian projection (forward.cu:155-256; in_frustum auxiliary.h:139-164)           */
/* All outputs are zero-filled for culled Gaussians (the reference leaves them undefined).    */
/* ------------------------------------------------------------------------------------------ */
void orc_preprocess(int P, int deg, int M, const float* means, const float* scales, float mod,
                    const float* rots, const float* opac, const float* shs, const float* cov_pre,
                    const float* col_pre, const float* view, const float* proj, const float* campos,
                    int W, int H, float tanx, float tany, int* radii, float* xy, float* depths,
                    float* cov3d, float* rgb, float* conic_op, uint32_t* tiles, uint8_t* clamped) {
  const float fy = H / (2.0f * tany), fx = W / (2.0f * tanx); /* rasterizer_impl.cu:222-223 */
  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
#pragma omp parallel for schedule(static)
  for (int i = 0; i < P; i++) {
    radii[i] = 0; tiles[i] = 0;
    xy[2 * i] = xy[2 * i + 1] = 0.f; depths[i] = 0.f;
    for (int k = 0; k < 6; k++) cov3d[6 * i + k] = 0.f;
    for (int k = 0; k < 3; k++) { rgb[3 * i + k] = 0.f; clamped[3 * i + k] = 0; }
    for (int k = 0; k < 4; k++) conic_op[4 * i + k] = 0.f;

    const float* p = means + 3 * i;
    float pv[3], ph[4];
    xform43(view, p, pv);
    if (pv[2] <= 0.2f) continue; /* near cull */
    xform44(proj, p, ph);
    float pw = 1.0f / (ph[3] + 0.0000001f);
    float projx = ph[0] * pw, projy = ph[1] * pw;

    const float* c6;
    if (cov_pre) c6 = cov_pre + 6 * i;
    else { cov3d_from_scale_rot(scales + 3 * i, mod, rots + 4 * i, cov3d + 6 * i); c6 = cov3d + 6 * i; }

    Ewa e; float a, b, c;
    ewa_setup(p, fx, fy, tanx, tany, view, &e);
    cov2d(&e, c6, &a, &b, &c);
    float det = a * c - b * b;
    if (det == 0.0f) continue;
    float det_inv = 1.f / det;
    float con0 = c * det_inv, con1 = -b * det_inv, con2 = a * det_inv;

    float mid = 0.5f * (a + c);
    float l1 = mid + sqrtf(fmaxf(0.1f, mid * mid - det));
    float l2 = mid - sqrtf(fmaxf(0.1f, mid * mid - det));
    float radius = ceilf(3.f * sqrtf(fmaxf(l1, l2)));
    float px = ndc_to_pix(projx, W), py = ndc_to_pix(projy, H);
    int x0, y0, x1, y1;
    tile_rect(px, py, (int)radius, gx, gy, &x0, &y0, &x1, &y1);
    if ((x1 - x0) * (y1 - y0) == 0) continue;

    if (!col_pre) { /* forward.cu:20-71 */
      float d[3] = {p[0] - campos[0], p[1] - campos[1], p[2] - campos[2]};
      float len = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
      d[0] /= len; d[1] /= len; d[2] /= len;
      float bs[16], dbs[16][3];
      int n = sh_basis(deg, d[0], d[1], d[2], bs, dbs);
      const float* sh = shs + (size_t)i * M * 3;
      for (int ch = 0; ch < 3; ch++) {
        float r = 0.f;
        for (int k = 0; k < n; k++) r += bs[k] * sh[3 * k + ch];
        r += 0.5f;
        clamped[3 * i + ch] = (r < 0);
        rgb[3 * i + ch] = fmaxf(r, 0.0f);
      }
    }
    depths[i] = pv[2];
    radii[i] = (int)radius;
    xy[2 * i] = px; xy[2 * i + 1] = py;
    conic_op[4 * i + 0] = con0; conic_op[4 * i + 1] = con1; conic_op[4 * i + 2] = con2;
    conic_op[4 * i + 3] = opac[i];
    tiles[i] = (uint32_t)((y1 - y0) * (x1 - x0));
  }
}

/* ------------------------------------------------------------------------------------------ */
/* K2-K5: scan, key emission, STABLE sort by (tile, depth bits), tile ranges                   */
/* (rasterizer_impl.cu:70-138, 277-317).  Returns R; caller sizes point_list >= R via a       */
/* first call with point_list == NULL.                                                         */
/* ------------------------------------------------------------------------------------------ */
static void radix_sort_pairs(uint64_t* k, uint32_t* v, uint64_t* k2, uint32_t* v2, size_t n, int bits) {
  /* LSD radix, 8 bits per pass: stable, like cub::DeviceRadixSort */
  for (int shift = 0; shift < bits; shift += 8) {
    size_t cnt[257]; memset(cnt, 0, sizeof cnt);
    for (size_t i = 0; i < n; i++) cnt[((k[i] >> shift) & 255) + 1]++;
    for (int b = 0; b < 256; b++) cnt[b + 1] += cnt[b];
    for (size_t i = 0; i < n; i++) { size_t d = cnt[(k[i] >> shift) & 255]++; k2[d] = k[i]; v2[d] = v[i]; }
    uint64_t* tk = k; k = k2; k2 = tk; uint32_t* tv = v; v = v2; v2 = tv;
  }
  /* after an odd number of passes the result lives in the scratch arrays: caller checks */
}

long long orc_bin_sort(int P, const int* radii, const float* xy, const float* depths,
                       const uint32_t* tiles, int W, int H, uint32_t* point_list /*[R]*/,
                       uint64_t* keys_out /*[R] or NULL*/, uint32_t* ranges /*[tiles*2]*/) {
  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
  size_t R = 0;
  for (int i = 0; i < P; i++) R += tiles[i];
  if (!point_list) return (long long)R;
  uint64_t* k = (uint64_t*)malloc((R + 1) * sizeof(uint64_t));
  uint64_t* k2 = (uint64_t*)malloc((R + 1) * sizeof(uint64_t));
  uint32_t* v = (uint32_t*)malloc((R + 1) * sizeof(uint32_t));
  uint32_t* v2 = (uint32_t*)malloc((R + 1) * sizeof(uint32_t));
  size_t off = 0;
  for (int i = 0; i < P; i++) { /* duplicateWithKeys: ascending Gaussian index */
    if (radii[i] <= 0) continue;
    int x0, y0, x1, y1;
    tile_rect(xy[2 * i], xy[2 * i + 1], radii[i], gx, gy, &x0, &y0, &x1, &y1);
    f32_t dz32 = (f32_t)depths[i]; /* (fp64 self-check build: keys still use the fp32 bit pattern) */
    uint32_t dbits; memcpy(&dbits, &dz32, 4);
    for (int y = y0; y < y1; y++)
      for (int x = x0; x < x1; x++) {
        uint64_t key = (uint64_t)(y * gx + x);
        key <<= 32; key |= dbits;
        k[off] = key; v[off] = (uint32_t)i; off++;
      }
  }
  /* getHigherMsb (rasterizer_impl.cu:35-50): number of bits needed for the tile count */
  int bit = 0; { uint32_t n = (uint32_t)(gx * gy); while (n >> bit) bit++; }
  int bits = 32 + bit;
  int passes = (bits + 7) / 8;
  radix_sort_pairs(k, v, k2, v2, R, bits);
  uint64_t* ks = (passes & 1) ? k2 : k; uint32_t* vs = (passes & 1) ? v2 : v;
  memcpy(point_list, vs, R * sizeof(uint32_t));
  if (keys_out) memcpy(keys_out, ks, R * sizeof(uint64_t));
  memset(ranges, 0, (size_t)gx * gy * 2 * sizeof(uint32_t));
  for (size_t i = 0; i < R; i++) { /* identifyTileRanges */
    uint32_t cur = (uint32_t)(ks[i] >> 32);
    if (i == 0) ranges[2 * cur] = 0;
    else {
      uint32_t prev = (uint32_t)(ks[i - 1] >> 32);
      if (cur != prev) { ranges[2 * prev + 1] = (uint32_t)i; ranges[2 * cur] = (uint32_t)i; }
    }
    if (i == R - 1) ranges[2 * cur + 1] = (uint32_t)R;
  }
  free(k); free(k2); free(v); free(v2);
  return (long long)R;
}

/* ------------------------------------------------------------------------------------------ */
/* K6: per-pixel front-to-back alpha compositing (forward.cu:261-374)                          */
/* ------------------------------------------------------------------------------------------ */
void orc_render_fwd(int W, int H, const uint32_t* ranges, const uint32_t* point_list, const float* xy,
                    const float* rgb, const float* conic_op, const float* bg, float* out_color,
                    float* final_T, uint32_t* n_contrib, unsigned long long* visited_pairs) {
  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
  unsigned long long visited = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : visited)
  for (int tile = 0; tile < gx * gy; tile++) {
    int tx = tile % gx, ty = tile / gx;
    uint32_t r0 = ranges[2 * tile], r1 = ranges[2 * tile + 1];
    for (int ly = 0; ly < TILE; ly++)
      for (int lx = 0; lx < TILE; lx++) {
        int x = tx * TILE + lx, y = ty * TILE + ly;
        if (x >= W || y >= H) continue;
        float pxf = (float)x, pyf = (float)y; /* integer pixel coords, no +0.5 */
        float T = 1.0f, C[NCH] = {0, 0, 0};
        uint32_t contributor = 0, last = 0;
        for (uint32_t e = r0; e < r1; e++) {
          contributor++;
          uint32_t g = point_list[e];
          float dx = xy[2 * g] - pxf, dy = xy[2 * g + 1] - pyf;
          const float* co = conic_op + 4 * g;
          float power = -0.5f * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
          if (power > 0.0f) continue;
          float alpha = fminf(0.99f, co[3] * expf(power));
          if (alpha < 1.0f / 255.0f) continue;
          float test_T = T * (1 - alpha);
          if (test_T < 0.0001f) break; /* `done`: this entry is counted but not applied */
          for (int ch = 0; ch < NCH; ch++) C[ch] += rgb[3 * g + ch] * alpha * T;
          T = test_T;
          last = contributor;
        }
        visited += contributor;
        size_t pid = (size_t)y * W + x;
        final_T[pid] = T; n_contrib[pid] = last;
        for (int ch = 0; ch < NCH; ch++) out_color[(size_t)ch * H * W + pid] = C[ch] + T * bg[ch];
      }
  }
  if (visited_pairs) *visited_pairs = visited;
}

/* ------------------------------------------------------------------------------------------ */
/* K7: per-pixel back-to-front gradient replay (backward.cu:399-557)                           */
/* Accumulators are fp64 arrays: dmean2D[P*2], dconic[P*3] (xx,xy,yy), dopac[P], dcolor[P*3].   */
/* ------------------------------------------------------------------------------------------ */
static void atomic_add_d(double* p, double v) {
#pragma omp atomic
  *p += v;
}

void orc_render_bwd(int W, int H, const uint32_t* ranges, const uint32_t* point_list, const float* xy,
                    const float* rgb, const float* conic_op, const float* bg, const float* final_T,
                    const uint32_t* n_contrib, const float* dL_dpix, double* dmean2D, double* dconic,
                    double* dopac, double* dcolor) {
  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
  const float ddelx_dx = (float)(0.5 * W), ddely_dy = (float)(0.5 * H);
#pragma omp parallel for schedule(dynamic, 1)
  for (int tile = 0; tile < gx * gy; tile++) {
    int tx = tile % gx, ty = tile / gx;
    uint32_t r0 = ranges[2 * tile], r1 = ranges[2 * tile + 1];
    for (int ly = 0; ly < TILE; ly++)
      for (int lx = 0; lx < TILE; lx++) {
        int x = tx * TILE + lx, y = ty * TILE + ly;
        if (x >= W || y >= H) continue;
        size_t pid = (size_t)y * W + x;
        float pxf = (float)x, pyf = (float)y;
        const float T_final = final_T[pid];
        float T = T_final;
        const uint32_t last = n_contrib[pid];
        float accum[NCH] = {0, 0, 0}, last_color[NCH] = {0, 0, 0}, last_alpha = 0.f, dpix[NCH];
        for (int ch = 0; ch < NCH; ch++) dpix[ch] = dL_dpix[(size_t)ch * H * W + pid];
        /* entries r0 .. r0+last-1 were visited up to the last contributor; walk them backwards */
        for (uint32_t c = last; c-- > 0;) {
          uint32_t g = point_list[r0 + c];
          (void)r1;
          float dx = xy[2 * g] - pxf, dy = xy[2 * g + 1] - pyf;
          const float* co = conic_op + 4 * g;
          float power = -0.5f * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
          if (power > 0.0f) continue;
          float G = expf(power);
          float alpha = fminf(0.99f, co[3] * G);
          if (alpha < 1.0f / 255.0f) continue;
          T = T / (1.f - alpha);
          float dchannel_dcolor = alpha * T;
          float dL_dalpha = 0.0f;
          for (int ch = 0; ch < NCH; ch++) {
            float col = rgb[3 * g + ch];
            accum[ch] = last_alpha * last_color[ch] + (1.f - last_alpha) * accum[ch];
            last_color[ch] = col;
            dL_dalpha += (col - accum[ch]) * dpix[ch];
            atomic_add_d(&dcolor[3 * (size_t)g + ch], (double)(dchannel_dcolor * dpix[ch]));
          }
          dL_dalpha *= T;
          last_alpha = alpha;
          float bg_dot = 0;
          for (int ch = 0; ch < NCH; ch++) bg_dot += bg[ch] * dpix[ch];
          dL_dalpha += (-T_final / (1.f - alpha)) * bg_dot;
          float dL_dG = co[3] * dL_dalpha;
          float gdx = G * dx, gdy = G * dy;
          float dG_ddelx = -gdx * co[0] - gdy * co[1];
          float dG_ddely = -gdy * co[2] - gdx * co[1];
          atomic_add_d(&dmean2D[2 * (size_t)g + 0], (double)(dL_dG * dG_ddelx * ddelx_dx));
          atomic_add_d(&dmean2D[2 * (size_t)g + 1], (double)(dL_dG * dG_ddely * ddely_dy));
          atomic_add_d(&dconic[3 * (size_t)g + 0], (double)(-0.5f * gdx * dx * dL_dG));
          atomic_add_d(&dconic[3 * (size_t)g + 1], (double)(-0.5f * gdx * dy * dL_dG));
          atomic_add_d(&dconic[3 * (size_t)g + 2], (double)(-0.5f * gdy * dy * dL_dG));
          atomic_add_d(&dopac[g], (double)(G * dL_dalpha));
        }
      }
  }
}

/* ------------------------------------------------------------------------------------------ */
/* K8 + K9: conic -> cov2D -> cov3D/mean, mean2D -> mean3D, SH, cov3D -> scale/rotation         */
/* (backward.cu:144-274, 278-341, 346-396, 20-139)                                              */
/* Inputs dmean2D [P,2] (x,y), dconic [P,3] (xx,xy,yy) as fp32.                                 */
/* ------------------------------------------------------------------------------------------ */
void orc_preprocess_bwd(int P, int deg, int M, const float* means, const int* radii, const float* shs,
                        const uint8_t* clamped, const float* scales, const float* rots, float mod,
                        const float* cov3d, const float* view, const float* proj, const float* campos,
                        int W, int H, float tanx, float tany, const float* dmean2D, const float* dconic,
                        const float* dcolor, float* dmeans, float* dcov3d, float* dsh, float* dscale,
                        float* drot) {
  const float fy = H / (2.0f * tany), fx = W / (2.0f * tanx);
#pragma omp parallel for schedule(static)
  for (int i = 0; i < P; i++) {
    if (!(radii[i] > 0)) continue; /* outputs stay at their caller-provided zeros */
    const float* mean = means + 3 * i;
    const float* c6 = cov3d + 6 * i;
    /* ---- K8: computeCov2DCUDA ---- */
    Ewa e; float a, b, c;
    ewa_setup(mean, fx, fy, tanx, tany, view, &e);
    cov2d(&e, c6, &a, &b, &c);
    const float xmul = (e.txtz < -e.limx || e.txtz > e.limx) ? 0.f : 1.f;
    const float ymul = (e.tytz < -e.limy || e.tytz > e.limy) ? 0.f : 1.f;
    float dcx = dconic[3 * i], dcy = dconic[3 * i + 1], dcz = dconic[3 * i + 2];
    float denom = a * c - b * b;
    float da = 0, db = 0, dc = 0;
    float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
    float V[3][3]; sym6_to_mat(c6, V);
    const float(*A)[3] = e.A;
    if (denom2inv != 0) {
      da = denom2inv * (-c * c * dcx + 2 * b * c * dcy + (denom - a * c) * dcz);
      dc = denom2inv * (-a * a * dcz + 2 * a * b * dcy + (denom - a * c) * dcx);
      db = denom2inv * 2 * (b * c * dcx - (denom + 2 * b * b) * dcy + a * b * dcz);
      dcov3d[6 * i + 0] = A[0][0] * A[0][0] * da + A[0][0] * A[1][0] * db + A[1][0] * A[1][0] * dc;
      dcov3d[6 * i + 3] = A[0][1] * A[0][1] * da + A[0][1] * A[1][1] * db + A[1][1] * A[1][1] * dc;
      dcov3d[6 * i + 5] = A[0][2] * A[0][2] * da + A[0][2] * A[1][2] * db + A[1][2] * A[1][2] * dc;
      dcov3d[6 * i + 1] = 2 * A[0][0] * A[0][1] * da + (A[0][0] * A[1][1] + A[0][1] * A[1][0]) * db + 2 * A[1][0] * A[1][1] * dc;
      dcov3d[6 * i + 2] = 2 * A[0][0] * A[0][2] * da + (A[0][0] * A[1][2] + A[0][2] * A[1][0]) * db + 2 * A[1][0] * A[1][2] * dc;
      dcov3d[6 * i + 4] = 2 * A[0][2] * A[0][1] * da + (A[0][1] * A[1][2] + A[0][2] * A[1][1]) * db + 2 * A[1][1] * A[1][2] * dc;
    } else {
      for (int k = 0; k < 6; k++) dcov3d[6 * i + k] = 0;
    }
    float dA0[3], dA1[3];
    for (int k = 0; k < 3; k++) {
      float va0 = A[0][0] * V[k][0] + A[0][1] * V[k][1] + A[0][2] * V[k][2];
      float va1 = A[1][0] * V[k][0] + A[1][1] * V[k][1] + A[1][2] * V[k][2];
      dA0[k] = 2 * va0 * da + va1 * db;
      dA1[k] = 2 * va1 * dc + va0 * db;
    }
    /* Rw[r][k] = view[4k + r] */
    float dJ00 = view[0] * dA0[0] + view[4] * dA0[1] + view[8] * dA0[2];
    float dJ02 = view[2] * dA0[0] + view[6] * dA0[1] + view[10] * dA0[2];
    float dJ11 = view[1] * dA1[0] + view[5] * dA1[1] + view[9] * dA1[2];
    float dJ12 = view[2] * dA1[0] + view[6] * dA1[1] + view[10] * dA1[2];
    float tz = 1.f / e.t[2], tz2 = tz * tz, tz3 = tz2 * tz;
    float dtx = xmul * -fx * tz2 * dJ02;
    float dty = ymul * -fy * tz2 * dJ12;
    float dtz = -fx * tz2 * dJ00 - fy * tz2 * dJ11 + (2 * fx * e.t[0]) * tz3 * dJ02 + (2 * fy * e.t[1]) * tz3 * dJ12;
    float dm[3];
    dm[0] = view[0] * dtx + view[1] * dty + view[2] * dtz;
    dm[1] = view[4] * dtx + view[5] * dty + view[6] * dtz;
    dm[2] = view[8] * dtx + view[9] * dty + view[10] * dtz;

    /* ---- K9: mean2D -> mean3D (backward.cu:366-387) ---- */
    float mh[4]; xform44(proj, mean, mh);
    float mw = 1.0f / (mh[3] + 0.0000001f);
    float mul1 = mh[0] * mw * mw, mul2 = mh[1] * mw * mw;
    float gx2 = dmean2D[2 * i], gy2 = dmean2D[2 * i + 1];
    dm[0] += (proj[0] * mw - proj[3] * mul1) * gx2 + (proj[1] * mw - proj[3] * mul2) * gy2;
    dm[1] += (proj[4] * mw - proj[7] * mul1) * gx2 + (proj[5] * mw - proj[7] * mul2) * gy2;
    dm[2] += (proj[8] * mw - proj[11] * mul1) * gx2 + (proj[9] * mw - proj[11] * mul2) * gy2;

    /* ---- SH backward (backward.cu:20-139) ---- */
    if (shs) {
      float dorig[3] = {mean[0] - campos[0], mean[1] - campos[1], mean[2] - campos[2]};
      float len = sqrtf(dorig[0] * dorig[0] + dorig[1] * dorig[1] + dorig[2] * dorig[2]);
      float d[3] = {dorig[0] / len, dorig[1] / len, dorig[2] / len};
      float bs[16], dbs[16][3];
      int n = sh_basis(deg, d[0], d[1], d[2], bs, dbs);
      const float* sh = shs + (size_t)i * M * 3;
      float g[3];
      for (int ch = 0; ch < 3; ch++) g[ch] = dcolor[3 * i + ch] * (clamped[3 * i + ch] ? 0.f : 1.f);
      float ddir[3] = {0, 0, 0};
      for (int k = 0; k < n; k++) {
        float dot = 0.f;
        for (int ch = 0; ch < 3; ch++) {
          dsh[((size_t)i * M + k) * 3 + ch] = bs[k] * g[ch];
          dot += sh[3 * k + ch] * g[ch];
        }
        for (int ax = 0; ax < 3; ax++) ddir[ax] += dbs[k][ax] * dot;
      }
      /* dnormvdv (auxiliary.h:107-117) */
      float s2 = dorig[0] * dorig[0] + dorig[1] * dorig[1] + dorig[2] * dorig[2];
      float inv32 = 1.0f / sqrtf(s2 * s2 * s2);
      dm[0] += ((+s2 - dorig[0] * dorig[0]) * ddir[0] - dorig[1] * dorig[0] * ddir[1] - dorig[2] * dorig[0] * ddir[2]) * inv32;
      dm[1] += (-dorig[0] * dorig[1] * ddir[0] + (s2 - dorig[1] * dorig[1]) * ddir[1] - dorig[2] * dorig[1] * ddir[2]) * inv32;
      dm[2] += (-dorig[0] * dorig[2] * ddir[0] - dorig[1] * dorig[2] * ddir[1] + (s2 - dorig[2] * dorig[2]) * ddir[2]) * inv32;
    }
    dmeans[3 * i] = dm[0]; dmeans[3 * i + 1] = dm[1]; dmeans[3 * i + 2] = dm[2];

    /* ---- cov3D -> scale / rotation (backward.cu:278-341) ---- */
    if (scales) {
      float R[3][3], Mm[3][3], dS[3][3], dM[3][3], E[3][3];
      const float* q = rots + 4 * i;
      quat_to_rot(q, R);
      float s[3] = {mod * scales[3 * i], mod * scales[3 * i + 1], mod * scales[3 * i + 2]};
      for (int r_ = 0; r_ < 3; r_++)
        for (int k = 0; k < 3; k++) Mm[r_][k] = s[r_] * R[k][r_];
      const float* dc6 = dcov3d + 6 * i;
      dS[0][0] = dc6[0]; dS[0][1] = dS[1][0] = 0.5f * dc6[1]; dS[0][2] = dS[2][0] = 0.5f * dc6[2];
      dS[1][1] = dc6[3]; dS[1][2] = dS[2][1] = 0.5f * dc6[4]; dS[2][2] = dc6[5];
      for (int r_ = 0; r_ < 3; r_++)
        for (int k = 0; k < 3; k++)
          dM[r_][k] = 2.0f * (Mm[r_][0] * dS[0][k] + Mm[r_][1] * dS[1][k] + Mm[r_][2] * dS[2][k]);
      for (int r_ = 0; r_ < 3; r_++)
        dscale[3 * i + r_] = R[0][r_] * dM[r_][0] + R[1][r_] * dM[r_][1] + R[2][r_] * dM[r_][2];
      for (int a_ = 0; a_ < 3; a_++)
        for (int b_ = 0; b_ < 3; b_++) E[a_][b_] = s[b_] * dM[b_][a_]; /* dL/dRq[a][b] */
      float r = q[0], x = q[1], y = q[2], z = q[3];
      drot[4 * i + 0] = 2 * z * (E[1][0] - E[0][1]) + 2 * y * (E[0][2] - E[2][0]) + 2 * x * (E[2][1] - E[1][2]);
      drot[4 * i + 1] = 2 * y * (E[0][1] + E[1][0]) + 2 * z * (E[0][2] + E[2][0]) + 2 * r * (E[2][1] - E[1][2]) - 4 * x * (E[2][2] + E[1][1]);
      drot[4 * i + 2] = 2 * x * (E[0][1] + E[1][0]) + 2 * r * (E[0][2] - E[2][0]) + 2 * z * (E[2][1] + E[1][2]) - 4 * y * (E[2][2] + E[0][0]);
      drot[4 * i + 3] = 2 * r * (E[1][0] - E[0][1]) + 2 * x * (E[0][2] + E[2][0]) + 2 * y * (E[2][1] + E[1][2]) - 4 * z * (E[1][1] + E[0][0]);
    }
  }
}

/* markVisible (rasterizer_impl.cu:54-66) */
void orc_mark_visible(int P, const float* means, const float* view, uint8_t* present) {
  for (int i = 0; i < P; i++) {
    float pv[3]; xform43(view, means + 3 * i, pv);
    present[i] = pv[2] > 0.2f;
  }
}

int orc_num_threads(void) {
#ifdef _OPENMP
  extern int omp_get_max_threads(void);
  return omp_get_max_threads();
#else
  return 1;
#endif
}
