// Per-Gaussian math of the render path, shared by the CUDA kernels (device) and by the host-side
// check harness (tests/hostcheck, g++ -ffp-contract=off), so that the exact arithmetic the kernels
// run can be compared with the oracle without a GPU.
//
// Semantics follow the reference (paths relative to its checkout):
//   gaussian_splatting/gaussian.py:24-127            quaternion (x,y,z,w) -> R, Sigma = R S S R^T
//   gaussian_splatting/spherical_harmonics.py:50-166 SH basis (the reference's own signs), sigmoid
//   gaussian_splatting/utils.py:10-96,152-191        w2c, frustum test, inv2x2
//   gaussian_splatting/render.py:104-258,305-315     S1-S11, S15
// Backward formulas are derived by hand (the reference has no backward code: autograd) and are
// validated against the reference's autograd through tests/golden.
#pragma once
#include <string.h>
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define GS_HD __host__ __device__ __forceinline__
#else
#define GS_HD inline
#endif

namespace gs {

// ------------------------------------------------------------------------------------------------
// Parameters
// ------------------------------------------------------------------------------------------------
// exp(-q/2) = 2^(kBlendExpScale * q)
constexpr float kBlendExpScale = -0.72134752044448170368f;   // -log2(e)/2

struct RenderParams {          // by-value kernel argument; built on the host from b200gs_camera
  float fx, fy, cx, cy;
  float ulo, uhi, vlo, vhi;    // (float)(-guard-cx), (float)(W+guard-cx), ... (utils.py:81-91)
  float near_plane, far_plane;
  float min_conis, alpha_pre;  // alpha_cutoff * 0.5 (render.py:107)
  float chi2, alpha_max, alpha_cutoff;
  float chi2c, cut_e;          // blend gate terms: kBlendExpScale * chi2 and log2(alpha_cutoff) (see SplatRecord)
  int W, H, tiles_x, tiles_y;
  int row_begin, row_end;      // tile rows rendered by this rank [begin,end)
  // depth-sort keys: a survivor has near < z < far and z > 0, so float_bits(z) - key_base fits key_bits bits with the
  // all-ones value left for culled Gaussians (default near / far: 27 bits, three 9-bit radix passes instead of 4 x 8)
  uint32_t key_base;
  int key_bits;
};

// The reference multiplies fp32 tensors by Python floats: each scalar is rounded to fp32 at the op
// (e.g. `z * (-pix_guard - cx)`, utils.py:81: the bracket is evaluated in double first).
inline void fill_render_params(RenderParams& rp, int H, int W, double fx, double fy, double cx, double cy,
                               double near_plane, double far_plane, double pix_guard, double min_conis,
                               double chi_square_clip, double alpha_max, double alpha_cutoff) {
  rp.fx = (float)fx; rp.fy = (float)fy; rp.cx = (float)cx; rp.cy = (float)cy;
  rp.ulo = (float)(-pix_guard - cx);
  rp.uhi = (float)((double)W + pix_guard - cx);
  rp.vlo = (float)(-pix_guard - cy);
  rp.vhi = (float)((double)H + pix_guard - cy);
  rp.near_plane = (float)near_plane; rp.far_plane = (float)far_plane;
  rp.min_conis = (float)min_conis;
  rp.alpha_pre = (float)(alpha_cutoff * 0.5);
  rp.chi2 = (float)chi_square_clip;
  rp.alpha_max = (float)alpha_max;
  rp.alpha_cutoff = (float)alpha_cutoff;
  rp.chi2c = kBlendExpScale * rp.chi2;
  // lower bound of log2(alpha_raw): +inf when nothing can pass (alpha_cutoff > alpha_max), -inf when
  // everything does (alpha_cutoff <= 0)
  if (!(rp.alpha_cutoff <= rp.alpha_max)) rp.cut_e = INFINITY;
  else if (!(rp.alpha_cutoff > 0.f)) rp.cut_e = -INFINITY;
  else rp.cut_e = (float)log2((double)rp.alpha_cutoff);
  rp.W = W; rp.H = H;
  rp.tiles_x = (W + 15) / 16;
  rp.tiles_y = (H + 15) / 16;
  rp.row_begin = 0; rp.row_end = rp.tiles_y;
  {
    auto fbits = [](float x) { uint32_t u; memcpy(&u, &x, 4); return u; };
    const uint32_t lo = (rp.near_plane > 0.f) ? fbits(rp.near_plane) : 0u;          // z > max(near, 0)
    const uint32_t hi = (rp.far_plane > 0.f && rp.far_plane <= 3.0e38f) ? fbits(rp.far_plane) : 0x7F800000u;
    rp.key_base = 0; rp.key_bits = 32;
    if (hi > lo) {
      const uint32_t span = hi - lo;                     // live keys - lo are < span; the culled marker must exceed them
      int b = 1;
      while (b < 32 && ((1u << b) - 1u) < span) ++b;
      if (b < 32) { rp.key_base = lo; rp.key_bits = b; }
    }
  }
}

struct Pose {                  // derived from c2w (utils.py:25-29, render.py:156-157)
  float r[9];                  // world->camera rotation, row-major (= c2w[:3,:3]^T)
  float t[3];                  // w2c translation  = (-R^T) t
  float cam[3];                // camera centre in world space = c2w[:3,3]
};

GS_HD Pose make_pose(const float* c) {
  Pose p;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) p.r[3 * i + j] = c[4 * j + i];
  p.cam[0] = c[3]; p.cam[1] = c[7]; p.cam[2] = c[11];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    float acc = (-p.r[3 * i + 0]) * p.cam[0];
    acc = fmaf(-p.r[3 * i + 1], p.cam[1], acc);
    acc = fmaf(-p.r[3 * i + 2], p.cam[2], acc);
    p.t[i] = acc;
  }
  return p;
}

GS_HD float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// ------------------------------------------------------------------------------------------------
// Sigma from (scale_raw, q_raw): gaussian.py:71-127.  Returns the 6 unique entries of the symmetric
// matrix (xx, xy, xz, yy, yz, zz) and optionally R and s for the backward.
// ------------------------------------------------------------------------------------------------
struct Cov3 { float xx, xy, xz, yy, yz, zz; };

struct QuatScale {
  float R[9];
  float s[3];      // clamped scales
  float e[3];      // exp(scale_raw) before the clamp
  float qn[4];
  float norm;      // |q_raw|
};

GS_HD void quat_scale_forward(const float sr[3], const float q[4], QuatScale& o) {
#pragma unroll
  for (int i = 0; i < 3; ++i) { o.e[i] = expf(sr[i]); o.s[i] = fmaxf(o.e[i], 1e-6f); }
  o.norm = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  const float inv = 1.0f / (o.norm + 1e-9f);
  const float x = q[0] * inv, y = q[1] * inv, z = q[2] * inv, w = q[3] * inv;
  o.qn[0] = x; o.qn[1] = y; o.qn[2] = z; o.qn[3] = w;
  const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, xz = x * z, yz = y * z;
  const float xw = x * w, yw = y * w, zw = z * w;
  o.R[0] = 1.f - 2.f * (yy + zz); o.R[1] = 2.f * (xy - zw);       o.R[2] = 2.f * (xz + yw);
  o.R[3] = 2.f * (xy + zw);       o.R[4] = 1.f - 2.f * (xx + zz); o.R[5] = 2.f * (yz - xw);
  o.R[6] = 2.f * (xz - yw);       o.R[7] = 2.f * (yz + xw);       o.R[8] = 1.f - 2.f * (xx + yy);
}

// Full 3x3 (row-major) the way the reference forms it: ((R S) S) R^T.
GS_HD void sigma_full(const QuatScale& qs, float S[9]) {
  float M[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) M[3 * i + j] = (qs.R[3 * i + j] * qs.s[j]) * qs.s[j];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      float acc = M[3 * i + 0] * qs.R[3 * k + 0];
      acc = fmaf(M[3 * i + 1], qs.R[3 * k + 1], acc);
      acc = fmaf(M[3 * i + 2], qs.R[3 * k + 2], acc);
      S[3 * i + k] = acc;
    }
}

GS_HD Cov3 sym_from_full(const float S[9]) {
  Cov3 c;
  c.xx = S[0]; c.yy = S[4]; c.zz = S[8];
  c.xy = 0.5f * (S[1] + S[3]); c.xz = 0.5f * (S[2] + S[6]); c.yz = 0.5f * (S[5] + S[7]);
  return c;
}

// Backward: G = dL/dSigma as a full symmetric 3x3 (row-major).  Outputs grads of scale_raw, q_raw.
GS_HD void quat_scale_backward(const QuatScale& qs, const float q[4], const float G[9],
                               float g_sr[3], float g_q[4]) {
  // Sigma = R D R^T, D = diag(s^2):  dL/dD_j = (R^T G R)_jj,  dL/dR = 2 G R D (G symmetric).
  float GR[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      GR[3 * i + j] = G[3 * i + 0] * qs.R[0 + j] + G[3 * i + 1] * qs.R[3 + j] + G[3 * i + 2] * qs.R[6 + j];
  float gR[9];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const float dj = qs.R[0 + j] * GR[0 + j] + qs.R[3 + j] * GR[3 + j] + qs.R[6 + j] * GR[6 + j];
    const float gs = 2.f * qs.s[j] * dj;
    g_sr[j] = (qs.e[j] >= 1e-6f) ? gs * qs.e[j] : 0.f;
    const float s2 = 2.f * qs.s[j] * qs.s[j];
#pragma unroll
    for (int i = 0; i < 3; ++i) gR[3 * i + j] = GR[3 * i + j] * s2;
  }
  const float x = qs.qn[0], y = qs.qn[1], z = qs.qn[2], w = qs.qn[3];
  float gn[4];
  gn[0] = 2.f * (y * (gR[1] + gR[3]) + z * (gR[2] + gR[6]) - 2.f * x * (gR[4] + gR[8]) + w * (gR[7] - gR[5]));
  gn[1] = 2.f * (x * (gR[1] + gR[3]) - 2.f * y * (gR[0] + gR[8]) + z * (gR[5] + gR[7]) + w * (gR[2] - gR[6]));
  gn[2] = 2.f * (x * (gR[2] + gR[6]) + y * (gR[5] + gR[7]) - 2.f * z * (gR[0] + gR[4]) + w * (gR[3] - gR[1]));
  gn[3] = 2.f * (z * (gR[3] - gR[1]) + y * (gR[2] - gR[6]) + x * (gR[7] - gR[5]));
  // qn = q / (|q| + eps):  g_q = gn/(n+eps) - q (q.gn) / (n (n+eps)^2)
  const float ne = qs.norm + 1e-9f;
  const float dot = q[0] * gn[0] + q[1] * gn[1] + q[2] * gn[2] + q[3] * gn[3];
  const float k = (qs.norm > 0.f) ? dot / (qs.norm * ne * ne) : 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) g_q[i] = gn[i] / ne - q[i] * k;
}

// ------------------------------------------------------------------------------------------------
// Spherical harmonics: spherical_harmonics.py:129-166.
// coef(k, c): k = 0 -> f_dc[c];  k >= 1 -> f_rest[15 c + k - 1]  (channel-major, :125-127)
// ------------------------------------------------------------------------------------------------
#define GS_C0 0.28209479177387814f
#define GS_C1 0.4886025119029199f
#define GS_C2A 1.0925484305920792f
#define GS_C2B 0.31539156525252005f
#define GS_C2C 0.5462742152960396f
#define GS_C3A 0.5900435899266435f
#define GS_C3B 2.890611442640554f
#define GS_C3C 0.4570457994644658f
#define GS_C3D 0.3731763325901154f
#define GS_C3E 1.445305721320277f

struct ViewDir { float d[3]; float w[3]; float n; };

GS_HD ViewDir view_dir(const float p[3], const float cam[3]) {
  ViewDir v;
  v.w[0] = p[0] - cam[0]; v.w[1] = p[1] - cam[1]; v.w[2] = p[2] - cam[2];
  v.n = sqrtf(v.w[0] * v.w[0] + v.w[1] * v.w[1] + v.w[2] * v.w[2]);
  const float ne = v.n + 1e-8f;
  v.d[0] = v.w[0] / ne; v.d[1] = v.w[1] / ne; v.d[2] = v.w[2] / ne;
  return v;
}

GS_HD void sh_basis(const float d[3], float Y[16]) {
  const float x = d[0], y = d[1], z = d[2];
  const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, xz = x * z, yz = y * z;
  Y[0] = GS_C0;
  Y[1] = -GS_C1 * y;  Y[2] = GS_C1 * z;  Y[3] = -GS_C1 * x;
  Y[4] = GS_C2A * xy; Y[5] = GS_C2A * yz; Y[6] = GS_C2B * (3.f * zz - 1.f);
  Y[7] = GS_C2A * xz; Y[8] = GS_C2C * (xx - yy);
  Y[9] = GS_C3A * y * (3.f * xx - yy);
  Y[10] = GS_C3B * x * y * z;
  Y[11] = GS_C3C * y * (4.f * zz - xx - yy);
  Y[12] = GS_C3D * z * (2.f * zz - 3.f * xx - 3.f * yy);
  Y[13] = GS_C3C * x * (4.f * zz - xx - yy);
  Y[14] = GS_C3E * z * (xx - yy);
  Y[15] = GS_C3A * x * (xx - 3.f * yy);
}

// dL/dd from dL/dY.
GS_HD void sh_basis_backward(const float d[3], const float gY[16], float gd[3]) {
  const float x = d[0], y = d[1], z = d[2];
  const float xx = x * x, yy = y * y, zz = z * z;
  float gx = 0.f, gy = 0.f, gz = 0.f;
  gy += -GS_C1 * gY[1]; gz += GS_C1 * gY[2]; gx += -GS_C1 * gY[3];
  gx += GS_C2A * y * gY[4]; gy += GS_C2A * x * gY[4];
  gy += GS_C2A * z * gY[5]; gz += GS_C2A * y * gY[5];
  gz += GS_C2B * 6.f * z * gY[6];
  gx += GS_C2A * z * gY[7]; gz += GS_C2A * x * gY[7];
  gx += GS_C2C * 2.f * x * gY[8]; gy += -GS_C2C * 2.f * y * gY[8];
  // Y9 = C3A y (3xx - yy)
  gx += GS_C3A * 6.f * x * y * gY[9]; gy += GS_C3A * (3.f * xx - 3.f * yy) * gY[9];
  // Y10 = C3B xyz
  gx += GS_C3B * y * z * gY[10]; gy += GS_C3B * x * z * gY[10]; gz += GS_C3B * x * y * gY[10];
  // Y11 = C3C y (4zz - xx - yy)
  gx += GS_C3C * (-2.f * x * y) * gY[11]; gy += GS_C3C * (4.f * zz - xx - 3.f * yy) * gY[11];
  gz += GS_C3C * 8.f * y * z * gY[11];
  // Y12 = C3D z (2zz - 3xx - 3yy)
  gx += GS_C3D * (-6.f * x * z) * gY[12]; gy += GS_C3D * (-6.f * y * z) * gY[12];
  gz += GS_C3D * (6.f * zz - 3.f * xx - 3.f * yy) * gY[12];
  // Y13 = C3C x (4zz - xx - yy)
  gx += GS_C3C * (4.f * zz - 3.f * xx - yy) * gY[13]; gy += GS_C3C * (-2.f * x * y) * gY[13];
  gz += GS_C3C * 8.f * x * z * gY[13];
  // Y14 = C3E z (xx - yy)
  gx += GS_C3E * 2.f * x * z * gY[14]; gy += -GS_C3E * 2.f * y * z * gY[14]; gz += GS_C3E * (xx - yy) * gY[14];
  // Y15 = C3A x (xx - 3yy)
  gx += GS_C3A * (3.f * xx - 3.f * yy) * gY[15]; gy += GS_C3A * (-6.f * x * y) * gY[15];
  gd[0] = gx; gd[1] = gy; gd[2] = gz;
}

// d = w/(n+eps):  g_w = g_d/(n+eps) - w (w.g_d) / (n (n+eps)^2)
GS_HD void view_dir_backward(const ViewDir& v, const float gd[3], float gp[3]) {
  const float ne = v.n + 1e-8f;
  const float dot = v.w[0] * gd[0] + v.w[1] * gd[1] + v.w[2] * gd[2];
  const float k = (v.n > 0.f) ? dot / (v.n * ne * ne) : 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) gp[i] = gd[i] / ne - v.w[i] * k;
}

// ------------------------------------------------------------------------------------------------
// Projection: render.py S1-S11 + S15.
// ------------------------------------------------------------------------------------------------
struct Projection {
  // camera space
  float x, y, z, invz;
  // 2D
  float u, v, op, sig;           // sig = sigmoid(opacity_raw) before the clamp
  float m0[3], m1[3];            // rows of J * Rwc
  float a, b, d;                 // symmetrised Sigma_2D before the eigen clamp
  float lam1, lam2, l1, l2;      // eigenvalues before / after clamp(1e-6, 1e4)
  float c2, s2, rad;             // eigenvector angle (cos 2t, sin 2t) and half gap
  int clamped;                   // 0: eigen clamp inactive (fast path)
  float a2, b2, d2;              // Sigma_2D after the clamp
  float det, sdet;               // a2*d2-b2*b2, max(det,1e-12)
  float A11r, A22r;              // conic diagonal before the min_conis clamp
  float A11, A12, A22;           // conic used by the blend
  int radius;
  int tu0, tu1, tv0, tv1;        // tile rect (inclusive), already clipped to the screen
  int tiles;
  int offscreen;                 // 1 when the Gaussian passed S1-S7 and was dropped by the on-screen test (S10)
};

// Returns true when the Gaussian survives every cull (S1, S3, S7, S10).
GS_HD bool project_gaussian(const float p[3], const Cov3& S, float opacity_raw, const Pose& ps,
                            const RenderParams& rp, Projection& o) {
  // S1: opacity pre-cull (render.py:106-107)
  o.offscreen = 0;
  o.sig = sigmoidf_(opacity_raw);
  o.op = fminf(fmaxf(o.sig, 0.f), 0.999f);
  if (!(o.op >= rp.alpha_pre)) return false;
  // S2: world -> camera (utils.py:25-34)
  float cam3[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    float acc = ps.r[3 * i] * p[0];
    acc = fmaf(ps.r[3 * i + 1], p[1], acc);
    acc = fmaf(ps.r[3 * i + 2], p[2], acc);
    cam3[i] = acc + ps.t[i];
  }
  const float x = cam3[0], y = cam3[1], z = cam3[2];
  o.x = x; o.y = y; o.z = z;
  // S3: frustum (utils.py:72-96), all strict
  const float fxx = rp.fx * x, fyy = rp.fy * y;
  const bool vis = (z > 0.f) && (z > rp.near_plane) && (z < rp.far_plane) &&
                   (fxx > z * rp.ulo) && (fxx < z * rp.uhi) && (fyy > z * rp.vlo) && (fyy < z * rp.vhi);
  if (!vis) return false;
  // S4 (render.py:146)
  o.u = fxx / z + rp.cx;
  o.v = fyy / z + rp.cy;
  // S5 (render.py:156-175): Sigma_2D = (J Rwc) Sigma (J Rwc)^T, symmetric part
  const float invz = 1.0f / fmaxf(z, 1e-6f);
  const float invz2 = invz * invz;
  o.invz = invz;
  const float j00 = rp.fx * invz, j11 = rp.fy * invz;
  const float j02 = -rp.fx * x * invz2, j12 = -rp.fy * y * invz2;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    o.m0[k] = fmaf(j02, ps.r[6 + k], j00 * ps.r[0 + k]);
    o.m1[k] = fmaf(j12, ps.r[6 + k], j11 * ps.r[3 + k]);
  }
  float v0[3], v1[3];   // Sigma m0, Sigma m1
  v0[0] = fmaf(S.xz, o.m0[2], fmaf(S.xy, o.m0[1], S.xx * o.m0[0]));
  v0[1] = fmaf(S.yz, o.m0[2], fmaf(S.yy, o.m0[1], S.xy * o.m0[0]));
  v0[2] = fmaf(S.zz, o.m0[2], fmaf(S.yz, o.m0[1], S.xz * o.m0[0]));
  v1[0] = fmaf(S.xz, o.m1[2], fmaf(S.xy, o.m1[1], S.xx * o.m1[0]));
  v1[1] = fmaf(S.yz, o.m1[2], fmaf(S.yy, o.m1[1], S.xy * o.m1[0]));
  v1[2] = fmaf(S.zz, o.m1[2], fmaf(S.yz, o.m1[1], S.xz * o.m1[0]));
  o.a = fmaf(o.m0[2], v0[2], fmaf(o.m0[1], v0[1], o.m0[0] * v0[0]));
  o.b = fmaf(o.m1[2], v0[2], fmaf(o.m1[1], v0[1], o.m1[0] * v0[0]));
  o.d = fmaf(o.m1[2], v1[2], fmaf(o.m1[1], v1[1], o.m1[0] * v1[0]));
  // S7: non-finite covariance -> dropped (render.py:187-188)
  if (!(isfinite(o.a) && isfinite(o.b) && isfinite(o.d))) return false;
  // S6: symmetric 2x2 eigen-decomposition in closed form, clamp, rebuild (render.py:177-179)
  const float mid = 0.5f * (o.a + o.d);
  const float h = 0.5f * (o.a - o.d);
  o.rad = sqrtf(fmaf(h, h, o.b * o.b));
  o.lam1 = mid - o.rad;
  o.lam2 = mid + o.rad;
  o.l1 = fminf(fmaxf(o.lam1, 1e-6f), 1e4f);
  o.l2 = fminf(fmaxf(o.lam2, 1e-6f), 1e4f);
  o.clamped = (o.l1 != o.lam1) || (o.l2 != o.lam2);
  if (o.rad > 0.f) { o.c2 = h / o.rad; o.s2 = o.b / o.rad; } else { o.c2 = 1.f; o.s2 = 0.f; }
  if (!o.clamped) {
    o.a2 = o.a; o.b2 = o.b; o.d2 = o.d;
  } else {
    const float sum = 0.5f * (o.l1 + o.l2), dif = 0.5f * (o.l2 - o.l1);
    o.a2 = fmaf(dif, o.c2, sum);
    o.d2 = fmaf(-dif, o.c2, sum);
    o.b2 = dif * o.s2;
  }
  if (!(isfinite(o.a2) && isfinite(o.b2) && isfinite(o.d2))) return false;
  // S9 (render.py:227-233)
  const float lmax = fminf(fmaxf(o.l2, 1e-12f), 1e4f);
  const float rf = ceilf(2.5f * sqrtf(lmax));
  o.radius = (int)rf;
  const float umin_f = floorf(o.u - rf), umax_f = floorf(o.u + rf);
  const float vmin_f = floorf(o.v - rf), vmax_f = floorf(o.v + rf);
  // S10 (render.py:234-247): compare as floats (exact for integers < 2^24), then clamp
  const float Wf = (float)rp.W, Hf = (float)rp.H;
  if (!((umax_f >= 0.f) && (umin_f < Wf) && (vmax_f >= 0.f) && (vmin_f < Hf))) { o.offscreen = 1; return false; }
  const int umin = (int)fminf(fmaxf(umin_f, 0.f), Wf - 1.f), umax = (int)fminf(fmaxf(umax_f, 0.f), Wf - 1.f);
  const int vmin = (int)fminf(fmaxf(vmin_f, 0.f), Hf - 1.f), vmax = (int)fminf(fmaxf(vmax_f, 0.f), Hf - 1.f);
  // S11 (render.py:251-258)
  o.tu0 = umin >> 4; o.tu1 = umax >> 4; o.tv0 = vmin >> 4; o.tv1 = vmax >> 4;
  o.tiles = (o.tu1 - o.tu0 + 1) * (o.tv1 - o.tv0 + 1);
  // S15 (render.py:307-315, utils.py:180-191)
  o.det = o.a2 * o.d2 - o.b2 * o.b2;
  o.sdet = fmaxf(o.det, 1e-12f);
  o.A11r = o.d2 / o.sdet;
  o.A22r = o.a2 / o.sdet;
  o.A12 = -o.b2 / o.sdet;
  o.A11 = fmaxf(o.A11r, rp.min_conis);
  o.A22 = fmaxf(o.A22r, rp.min_conis);
  return true;
}

// Gradients that arrive from the blend backward, per Gaussian.
struct SplatGrad { float u, v, A11, A12, A22, op; };

// The blend backward accumulates, per Gaussian, the moments of dL/dq (q = the Mahalanobis form of
// render.py:362) over the pixels the splat contributed to, with du = px - u, dv = py - v:
//   M0 = sum dL/dq,  Mx = sum dL/dq du,  My = sum dL/dq dv,  Mxx = sum dL/dq du^2,  Mxy = sum dL/dq du dv,  Myy = ...
// q = A11 du^2 + 2 A12 du dv + A22 dv^2 and alpha_raw = op exp(-q/2) (so dL/dop = -2 dL/dq / op pixel by pixel) give
GS_HD SplatGrad splat_grad_from_moments(const Projection& o, float Mx, float My, float Mxx, float Mxy, float Myy,
                                        float M0) {
  SplatGrad g;
  g.u = -(2.f * o.A11 * Mx + 2.f * o.A12 * My);
  g.v = -(2.f * o.A22 * My + 2.f * o.A12 * Mx);
  g.A11 = Mxx;
  g.A12 = 2.f * Mxy;
  g.A22 = Myy;
  g.op = -2.f * M0 / o.op;
  return g;
}

// Backward of project_gaussian.  `o` is the recomputed forward.  Outputs: g_p (adds the projection
// part), G (dL/dSigma, full symmetric row-major 3x3), g_opacity_raw.
GS_HD void project_backward(const float p[3], const Cov3& S, const Pose& ps, const RenderParams& rp,
                            const Projection& o, const SplatGrad& g, float g_p[3], float G[9],
                            float& g_opacity_raw) {
  // opacity: clamp(sigmoid, 0, .999) passes gradient on [0, .999] inclusive
  g_opacity_raw = (o.sig <= 0.999f) ? g.op * o.sig * (1.f - o.sig) : 0.f;
  // conic -> Sigma_2D'
  const float gA11 = (o.A11r >= rp.min_conis) ? g.A11 : 0.f;
  const float gA22 = (o.A22r >= rp.min_conis) ? g.A22 : 0.f;
  const float gA12 = g.A12;
  const float inv = 1.f / o.sdet;
  const float g_sdet = -(gA11 * o.d2 - gA12 * o.b2 + gA22 * o.a2) * inv * inv;
  const float g_det = (o.det >= 1e-12f) ? g_sdet : 0.f;
  float ga = gA22 * inv + g_det * o.d2;
  float gd = gA11 * inv + g_det * o.a2;
  float gb = -gA12 * inv - 2.f * g_det * o.b2;      // total derivative w.r.t. the off-diagonal value
  // eigen clamp (Daleckii-Krein with f = clamp): identity on the fast path
  if (o.clamped) {
    const float f1 = (o.lam1 >= 1e-6f && o.lam1 <= 1e4f) ? 1.f : 0.f;
    const float f2 = (o.lam2 >= 1e-6f && o.lam2 <= 1e4f) ? 1.f : 0.f;
    if (o.rad > 0.f) {
      const float k12 = (o.l2 - o.l1) / (o.lam2 - o.lam1);
      const float c2 = o.c2, s2 = o.s2;
      const float gp1 = 0.5f * (ga * (1.f - c2) - gb * s2 + gd * (1.f + c2));   // <G,P1>
      const float gp2 = 0.5f * (ga * (1.f + c2) + gb * s2 + gd * (1.f - c2));   // <G,P2>
      const float gq = 0.5f * (-ga * s2 + gb * c2 + gd * s2);                   // 1/2 <G,Q>
      const float w1 = f1 * gp1, w2 = f2 * gp2, wq = k12 * gq;
      ga = 0.5f * (w1 * (1.f - c2) + w2 * (1.f + c2)) - wq * s2;
      gd = 0.5f * (w1 * (1.f + c2) + w2 * (1.f - c2)) + wq * s2;
      gb = 2.f * (0.5f * (-w1 * s2 + w2 * s2) + wq * c2);
    } else {
      ga *= f1; gd *= f1; gb *= f1;
    }
  }
  // Sigma_2D = M Sigma M^T
  float v0[3], v1[3];
  v0[0] = S.xx * o.m0[0] + S.xy * o.m0[1] + S.xz * o.m0[2];
  v0[1] = S.xy * o.m0[0] + S.yy * o.m0[1] + S.yz * o.m0[2];
  v0[2] = S.xz * o.m0[0] + S.yz * o.m0[1] + S.zz * o.m0[2];
  v1[0] = S.xx * o.m1[0] + S.xy * o.m1[1] + S.xz * o.m1[2];
  v1[1] = S.xy * o.m1[0] + S.yy * o.m1[1] + S.yz * o.m1[2];
  v1[2] = S.xz * o.m1[0] + S.yz * o.m1[1] + S.zz * o.m1[2];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      G[3 * i + j] = ga * o.m0[i] * o.m0[j] + 0.5f * gb * (o.m0[i] * o.m1[j] + o.m1[i] * o.m0[j]) +
                     gd * o.m1[i] * o.m1[j];
  float gm0[3], gm1[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    gm0[k] = 2.f * ga * v0[k] + gb * v1[k];
    gm1[k] = 2.f * gd * v1[k] + gb * v0[k];
  }
  // M = J Rwc
  const float gj00 = gm0[0] * ps.r[0] + gm0[1] * ps.r[1] + gm0[2] * ps.r[2];
  const float gj02 = gm0[0] * ps.r[6] + gm0[1] * ps.r[7] + gm0[2] * ps.r[8];
  const float gj11 = gm1[0] * ps.r[3] + gm1[1] * ps.r[4] + gm1[2] * ps.r[5];
  const float gj12 = gm1[0] * ps.r[6] + gm1[1] * ps.r[7] + gm1[2] * ps.r[8];
  const float invz = o.invz, invz2 = invz * invz;
  const float g_invz = rp.fx * gj00 + rp.fy * gj11 - 2.f * rp.fx * o.x * invz * gj02 -
                       2.f * rp.fy * o.y * invz * gj12;
  float gx = -rp.fx * invz2 * gj02;
  float gy = -rp.fy * invz2 * gj12;
  float gz = (o.z >= 1e-6f) ? -invz2 * g_invz : 0.f;
  // u = fx x / z + cx
  const float rz = 1.f / o.z;
  gx += g.u * rp.fx * rz;
  gy += g.v * rp.fy * rz;
  gz += -(g.u * rp.fx * o.x + g.v * rp.fy * o.y) * rz * rz;
  // camera = Rwc p + t
#pragma unroll
  for (int k = 0; k < 3; ++k) g_p[k] = ps.r[0 + k] * gx + ps.r[3 + k] * gy + ps.r[6 + k] * gz;
}

// ---- optimizer arithmetic (csrc/optim.cu, csrc/peer.cu; host-checkable like everything in this header) -----------
// The parameter update of Adam in fp32, the arithmetic of torch/optim/adam.py:
//     m <- m + (g - m)(1 - beta1);  v <- v beta2 + (1 - beta2) g g;  p <- p - step_size * m / (sqrt(v) / bc2_sqrt + eps)
// IEEE sqrt and division run a short inline sequence for ordinary operands and branch to a ~100-instruction
// subroutine as soon as ONE lane of the warp holds a zero or denormal operand.  Per-Gaussian gradients do: every culled
// Gaussian has g = m = v = 0 and barely visible ones have g*g in the denormal range, so inside a real training
// iteration nearly every warp took the slow branches (ncu: 2.4x the instructions, 463 us instead of 249 us at
// N = 1M).  Here zero and tiny operands are rescaled by exact powers of two around the operation (zero: replaced by a
// harmless value and the result selected back), which gives the same correctly rounded results for every input whose
// result is a normal number.
GS_HD float sqrt_no_slow_path(float v) {          // v >= 0
  const bool tiny = v < 0x1p-80f;
  const bool zero = v == 0.f;
  const float x = zero ? 1.f : (tiny ? v * 0x1p64f : v);
  const float s = sqrtf(x);
  return zero ? 0.f : (tiny ? s * 0x1p-32f : s);
}
GS_HD float div_no_slow_path(float a, float b) {  // b >= eps > 0, an ordinary number
  const bool tiny = fabsf(a) < 0x1p-60f;
  const bool zero = a == 0.f;
  const float x = zero ? b : (tiny ? a * 0x1p64f : a);
  const float q = x / b;
  return zero ? 0.f : (tiny ? q * 0x1p-64f : q);
}
GS_HD void adam_update_f32(float& p, float g, float& m, float& v, float one_minus_beta1, float beta2,
                                                float one_minus_beta2, float eps, float step_size, float bc2_sqrt) {
  m = m + (g - m) * one_minus_beta1;
  v = fmaf(one_minus_beta2 * g, g, v * beta2);
  const float denom = div_no_slow_path(sqrt_no_slow_path(v), bc2_sqrt) + eps;
  p = p - step_size * div_no_slow_path(m, denom);
}

}  // namespace gs
