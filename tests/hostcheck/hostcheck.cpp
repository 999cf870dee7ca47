// Host-side check harness (TEST INFRASTRUCTURE, never shipped or loaded by the product path).
//
// Compiles csrc/gs_math.cuh - the very same per-Gaussian arithmetic the CUDA kernels run - with g++
// (-ffp-contract=off, explicit fmaf only) so that the projection / SH / backward chain can be compared
// with the oracle and the golden vectors in a container without a GPU.
#include <stdint.h>
#include <string.h>

#include "../../3d-gaussian-splatting-for-novel-view-synthesis_b200/csrc/gs_math.cuh"

using namespace gs;

extern "C" {

// cam: H, W, fx, fy, cx, cy (doubles).  Raw parameters -> per-Gaussian splat values.
// out arrays sized n: vis(int32) u v z op A11 A12 A22 r g b (float) radius(int32) rect[4](int32) tiles(int32)
// lam2 (float) clamped(int32)
void hc_project(int n, const float* pos, const float* scale_raw, const float* q_raw, const float* sigma,
                const float* opacity_raw, const float* f_dc, const float* f_rest, const float* color,
                const float* c2w, const double* cam, int32_t* vis, float* u, float* v, float* z, float* op,
                float* conic, float* rgb, int32_t* radius, int32_t* rect, int32_t* tiles, float* lam2,
                int32_t* clamped) {
  RenderParams rp;
  fill_render_params(rp, (int)cam[0], (int)cam[1], cam[2], cam[3], cam[4], cam[5], 0.01, 100.0, 32.0, 1e-6, 6.25,
                     0.99, 1.0 / 128.0);
  const Pose ps = make_pose(c2w);
  for (int i = 0; i < n; ++i) {
    Cov3 S;
    if (scale_raw) {
      QuatScale qs;
      quat_scale_forward(scale_raw + 3 * i, q_raw + 4 * i, qs);
      float full[9];
      sigma_full(qs, full);
      S = sym_from_full(full);
    } else {
      S = sym_from_full(sigma + 9 * i);
    }
    Projection o;
    memset(&o, 0, sizeof(o));
    const bool ok = project_gaussian(pos + 3 * i, S, opacity_raw[i], ps, rp, o);
    vis[i] = ok ? 1 : 0;
    if (!ok) continue;
    u[i] = o.u; v[i] = o.v; z[i] = o.z; op[i] = o.op;
    conic[3 * i] = o.A11; conic[3 * i + 1] = o.A12; conic[3 * i + 2] = o.A22;
    radius[i] = o.radius;
    rect[4 * i] = o.tu0; rect[4 * i + 1] = o.tu1; rect[4 * i + 2] = o.tv0; rect[4 * i + 3] = o.tv1;
    tiles[i] = o.tiles;
    lam2[i] = o.l2;
    clamped[i] = o.clamped;
    if (f_dc) {
      const ViewDir vd = view_dir(pos + 3 * i, ps.cam);
      float Y[16];
      sh_basis(vd.d, Y);
      for (int c = 0; c < 3; ++c) {
        float acc = f_dc[3 * i + c] * Y[0];
        for (int k = 1; k < 16; ++k) acc = fmaf(f_rest[45 * i + 15 * c + k - 1], Y[k], acc);
        rgb[3 * i + c] = sigmoidf_(acc);
      }
    } else {
      for (int c = 0; c < 3; ++c) rgb[3 * i + c] = color[3 * i + c];
    }
  }
}

// Per-Gaussian backward chain: splat grads sg[n][9] = (u, v, A11, A12, A22, op, r, g, b) -> leaf grads.
void hc_backward(int n, const float* pos, const float* scale_raw, const float* q_raw, const float* opacity_raw,
                 const float* f_dc, const float* f_rest, const float* c2w, const double* cam, const float* sg,
                 float* g_pos, float* g_scale, float* g_q, float* g_op, float* g_dc, float* g_rest) {
  RenderParams rp;
  fill_render_params(rp, (int)cam[0], (int)cam[1], cam[2], cam[3], cam[4], cam[5], 0.01, 100.0, 32.0, 1e-6, 6.25,
                     0.99, 1.0 / 128.0);
  const Pose ps = make_pose(c2w);
  for (int i = 0; i < n; ++i) {
    for (int k = 0; k < 3; ++k) { g_pos[3 * i + k] = 0; g_scale[3 * i + k] = 0; g_dc[3 * i + k] = 0; }
    for (int k = 0; k < 4; ++k) g_q[4 * i + k] = 0;
    for (int k = 0; k < 45; ++k) g_rest[45 * i + k] = 0;
    g_op[i] = 0;
    QuatScale qs;
    quat_scale_forward(scale_raw + 3 * i, q_raw + 4 * i, qs);
    float full[9];
    sigma_full(qs, full);
    const Cov3 S = sym_from_full(full);
    Projection o;
    memset(&o, 0, sizeof(o));
    if (!project_gaussian(pos + 3 * i, S, opacity_raw[i], ps, rp, o)) continue;
    SplatGrad g;
    const float* s = sg + 9 * i;
    g.u = s[0]; g.v = s[1]; g.A11 = s[2]; g.A12 = s[3]; g.A22 = s[4]; g.op = s[5];
    float gp[3], G[9], gop;
    project_backward(pos + 3 * i, S, ps, rp, o, g, gp, G, gop);
    quat_scale_backward(qs, q_raw + 4 * i, G, g_scale + 3 * i, g_q + 4 * i);
    g_op[i] = gop;
    const ViewDir vd = view_dir(pos + 3 * i, ps.cam);
    float Y[16], gY[16];
    sh_basis(vd.d, Y);
    for (int k = 0; k < 16; ++k) gY[k] = 0.f;
    for (int c = 0; c < 3; ++c) {
      float acc = f_dc[3 * i + c] * Y[0];
      for (int k = 1; k < 16; ++k) acc = fmaf(f_rest[45 * i + 15 * c + k - 1], Y[k], acc);
      const float col = sigmoidf_(acc);
      const float ga = s[6 + c] * col * (1.f - col);
      g_dc[3 * i + c] = ga * Y[0];
      for (int k = 1; k < 16; ++k) {
        g_rest[45 * i + 15 * c + k - 1] = ga * Y[k];
        gY[k] = fmaf(ga, f_rest[45 * i + 15 * c + k - 1], gY[k]);
      }
    }
    float gd[3], gpv[3];
    sh_basis_backward(vd.d, gY, gd);
    view_dir_backward(vd, gd, gpv);
    for (int k = 0; k < 3; ++k) g_pos[3 * i + k] = gp[k] + gpv[k];
  }
}

// Optimizer arithmetic of csrc/optim.cu / csrc/peer.cu: the rescaled sqrt / division next to the plain IEEE operations,
// and one Adam update per element.
void hc_sqrt_div(int n, const float* a, const float* b, float* sqrt_resc, float* sqrt_plain, float* div_resc,
                 float* div_plain) {
  for (int i = 0; i < n; ++i) {
    sqrt_resc[i] = gs::sqrt_no_slow_path(a[i] < 0.f ? -a[i] : a[i]);
    sqrt_plain[i] = sqrtf(a[i] < 0.f ? -a[i] : a[i]);
    div_resc[i] = gs::div_no_slow_path(a[i], b[i]);
    div_plain[i] = a[i] / b[i];
  }
}
void hc_adam(int n, float* p, const float* g, float* m, float* v, double beta1, double beta2, double eps, float step_size,
             float bc2_sqrt) {
  // the constants as launch_adam_step / launch_peer_step form them: in double on the host, then rounded to fp32
  const float omb1 = (float)(1.0 - beta1), b2 = (float)beta2, omb2 = (float)(1.0 - beta2), e = (float)eps;
  for (int i = 0; i < n; ++i) gs::adam_update_f32(p[i], g[i], m[i], v[i], omb1, b2, omb2, e, step_size, bc2_sqrt);
}

}  // extern "C"
