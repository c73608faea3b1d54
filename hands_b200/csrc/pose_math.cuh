// Per-joint rotation math of the MANO path, forward and hand-derived backward.
//   log map  : common/rot.py:44-193 (matrix -> best-conditioned quaternion -> axis-angle)
//   Rodrigues: smplx/lbs.py::batch_rodrigues (angle = ||r + 1e-8||)            [smplx-recalled]
// The backward functions reproduce what torch.autograd computes through those exact op
// sequences (zero sub-gradients, branch selection, the 0.1 floor), not the analytic SO(3) maps.
#pragma once
#include <cuda_runtime.h>

namespace hb {

__device__ __forceinline__ void mat3_mul(const float* A, const float* B, float* C) {  // C = A B
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) C[r * 3 + c] = A[r * 3 + 0] * B[0 * 3 + c] + A[r * 3 + 1] * B[1 * 3 + c] + A[r * 3 + 2] * B[2 * 3 + c];
}
__device__ __forceinline__ void mat3_mulT(const float* A, const float* B, float* C) {  // C = A B^T
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) C[r * 3 + c] = A[r * 3 + 0] * B[c * 3 + 0] + A[r * 3 + 1] * B[c * 3 + 1] + A[r * 3 + 2] * B[c * 3 + 2];
}
__device__ __forceinline__ void matT3_mul(const float* A, const float* B, float* C) {  // C = A^T B
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) C[r * 3 + c] = A[0 * 3 + r] * B[0 * 3 + c] + A[1 * 3 + r] * B[1 * 3 + c] + A[2 * 3 + r] * B[2 * 3 + c];
}
__device__ __forceinline__ void mat3_vec(const float* A, const float* x, float* y) {  // y = A x
#pragma unroll
  for (int r = 0; r < 3; ++r) y[r] = A[r * 3 + 0] * x[0] + A[r * 3 + 1] * x[1] + A[r * 3 + 2] * x[2];
}
__device__ __forceinline__ void matT3_vec(const float* A, const float* x, float* y) {  // y = A^T x
#pragma unroll
  for (int r = 0; r < 3; ++r) y[r] = A[0 * 3 + r] * x[0] + A[1 * 3 + r] * x[1] + A[2 * 3 + r] * x[2];
}

// -------------------------------------------------------------------------------------------
// log map.  m: row-major 3x3.  Saves what the backward needs in LogMapCtx.
// -------------------------------------------------------------------------------------------
struct LogMapCtx {
  float q[4];     // selected quaternion (w,x,y,z)
  float cand[4];  // selected un-normalised candidate row
  float qa;       // q_abs[idx]
  float den;      // 2*max(qa, 0.1)
  float n, half, ang, sh;
  int idx;
  bool small;
};

__device__ __forceinline__ void logmap_fwd(const float* m, float* aa, LogMapCtx& c) {
  const float m00 = m[0], m01 = m[1], m02 = m[2], m10 = m[3], m11 = m[4], m12 = m[5], m20 = m[6], m21 = m[7], m22 = m[8];
  float t[4];
  t[0] = __fadd_rn(__fadd_rn(__fadd_rn(1.0f, m00), m11), m22);
  t[1] = __fsub_rn(__fsub_rn(__fadd_rn(1.0f, m00), m11), m22);
  t[2] = __fsub_rn(__fadd_rn(__fsub_rn(1.0f, m00), m11), m22);
  t[3] = __fadd_rn(__fsub_rn(__fsub_rn(1.0f, m00), m11), m22);
  float qa[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) qa[k] = t[k] > 0.0f ? sqrtf(t[k]) : 0.0f;
  int idx = 0;
  float best = qa[0];
#pragma unroll
  for (int k = 1; k < 4; ++k)
    if (qa[k] > best) { best = qa[k]; idx = k; }  // first maximum, like torch.argmax
  const float sq = best * best;
  float cd[4];
  if (idx == 0)      { cd[0] = sq;        cd[1] = m21 - m12; cd[2] = m02 - m20; cd[3] = m10 - m01; }
  else if (idx == 1) { cd[0] = m21 - m12; cd[1] = sq;        cd[2] = m10 + m01; cd[3] = m02 + m20; }
  else if (idx == 2) { cd[0] = m02 - m20; cd[1] = m10 + m01; cd[2] = sq;        cd[3] = m12 + m21; }
  else               { cd[0] = m10 - m01; cd[1] = m20 + m02; cd[2] = m21 + m12; cd[3] = sq; }
  const float den = 2.0f * fmaxf(best, 0.1f);
#pragma unroll
  for (int k = 0; k < 4; ++k) { c.cand[k] = cd[k]; c.q[k] = __fdiv_rn(cd[k], den); }
  c.qa = best; c.den = den; c.idx = idx;
  const float qx = c.q[1], qy = c.q[2], qz = c.q[3];
  const float n = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(qx, qx), __fmul_rn(qy, qy)), __fmul_rn(qz, qz)));
  const float half = atan2f(n, c.q[0]);
  const float ang = 2.0f * half;
  const bool small = fabsf(ang) < 1e-6f;
  const float sh = small ? (0.5f - __fdiv_rn(__fmul_rn(ang, ang), 48.0f)) : __fdiv_rn(sinf(half), ang);
  c.n = n; c.half = half; c.ang = ang; c.sh = sh; c.small = small;
  aa[0] = __fdiv_rn(qx, sh); aa[1] = __fdiv_rn(qy, sh); aa[2] = __fdiv_rn(qz, sh);
}

// g_aa (3) -> g_m (9, row-major)
__device__ __forceinline__ void logmap_bwd(const LogMapCtx& c, const float* g_aa, float* g_m) {
  const float inv_sh = 1.0f / c.sh;
  float gq[4];
  gq[1] = g_aa[0] * inv_sh; gq[2] = g_aa[1] * inv_sh; gq[3] = g_aa[2] * inv_sh;
  const float g_sh = -(g_aa[0] * c.q[1] + g_aa[1] * c.q[2] + g_aa[2] * c.q[3]) * inv_sh * inv_sh;
  float g_half, g_ang;
  if (c.small) { g_ang = -g_sh * c.ang * (1.0f / 24.0f); g_half = 0.0f; }
  else {
    g_half = g_sh * cosf(c.half) / c.ang;
    g_ang = -g_sh * sinf(c.half) / (c.ang * c.ang);
  }
  g_half += 2.0f * g_ang;
  const float w = c.q[0];
  const float d2 = c.n * c.n + w * w;
  const float g_n = g_half * w / d2;
  gq[0] = -g_half * c.n / d2;
  if (c.n > 0.0f) {
    const float s = g_n / c.n;
    gq[1] += s * c.q[1]; gq[2] += s * c.q[2]; gq[3] += s * c.q[3];
  }
  // q = cand / den
  const float inv_den = 1.0f / c.den;
  float gc[4];
  float g_den = 0.0f;
#pragma unroll
  for (int k = 0; k < 4; ++k) { gc[k] = gq[k] * inv_den; g_den -= gq[k] * c.cand[k]; }
  g_den *= inv_den * inv_den;
  // den = 2*max(qa, 0.1); cand[idx] = qa^2; qa = sqrt(t) (zero sub-gradient at t<=0)
  float g_qa = (c.qa > 0.1f ? 2.0f * g_den : (c.qa == 0.1f ? g_den : 0.0f));
  g_qa += 2.0f * c.qa * gc[c.idx];
  const float g_t = c.qa > 0.0f ? g_qa * 0.5f / c.qa : 0.0f;
#pragma unroll
  for (int k = 0; k < 9; ++k) g_m[k] = 0.0f;
  // t[idx] = 1 +- m00 +- m11 +- m22
  const float s00 = (c.idx == 0 || c.idx == 1) ? 1.0f : -1.0f;
  const float s11 = (c.idx == 0 || c.idx == 2) ? 1.0f : -1.0f;
  const float s22 = (c.idx == 0 || c.idx == 3) ? 1.0f : -1.0f;
  g_m[0] = s00 * g_t; g_m[4] = s11 * g_t; g_m[8] = s22 * g_t;
  // off-diagonal combinations; index map: m01=1 m02=2 m10=3 m12=5 m20=6 m21=7
  if (c.idx == 0) {
    g_m[7] += gc[1]; g_m[5] -= gc[1];  // m21 - m12
    g_m[2] += gc[2]; g_m[6] -= gc[2];  // m02 - m20
    g_m[3] += gc[3]; g_m[1] -= gc[3];  // m10 - m01
  } else if (c.idx == 1) {
    g_m[7] += gc[0]; g_m[5] -= gc[0];  // m21 - m12
    g_m[3] += gc[2]; g_m[1] += gc[2];  // m10 + m01
    g_m[2] += gc[3]; g_m[6] += gc[3];  // m02 + m20
  } else if (c.idx == 2) {
    g_m[2] += gc[0]; g_m[6] -= gc[0];  // m02 - m20
    g_m[3] += gc[1]; g_m[1] += gc[1];  // m10 + m01
    g_m[5] += gc[3]; g_m[7] += gc[3];  // m12 + m21
  } else {
    g_m[3] += gc[0]; g_m[1] -= gc[0];  // m10 - m01
    g_m[6] += gc[1]; g_m[2] += gc[1];  // m20 + m02
    g_m[7] += gc[2]; g_m[5] += gc[2];  // m21 + m12
  }
}

// -------------------------------------------------------------------------------------------
// Rodrigues (smplx variant).  r: axis-angle (3) -> R (9).
// -------------------------------------------------------------------------------------------
struct RodCtx {
  float e[3], d[3], ang, s, c;
  float K[9], K2[9];
};

__device__ __forceinline__ void rodrigues_fwd(const float* r, float* R, RodCtx& x) {
#pragma unroll
  for (int k = 0; k < 3; ++k) x.e[k] = __fadd_rn(r[k], 1e-8f);
  x.ang = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(x.e[0], x.e[0]), __fmul_rn(x.e[1], x.e[1])), __fmul_rn(x.e[2], x.e[2])));
#pragma unroll
  for (int k = 0; k < 3; ++k) x.d[k] = __fdiv_rn(r[k], x.ang);
  sincosf(x.ang, &x.s, &x.c);
  const float dx = x.d[0], dy = x.d[1], dz = x.d[2];
  x.K[0] = 0.f; x.K[1] = -dz; x.K[2] = dy;
  x.K[3] = dz;  x.K[4] = 0.f; x.K[5] = -dx;
  x.K[6] = -dy; x.K[7] = dx;  x.K[8] = 0.f;
  mat3_mul(x.K, x.K, x.K2);
  const float omc = 1.0f - x.c;
#pragma unroll
  for (int k = 0; k < 9; ++k) R[k] = ((k % 4 == 0) ? 1.0f : 0.0f) + x.s * x.K[k] + omc * x.K2[k];
}

// gR (9) -> g_r (3)
__device__ __forceinline__ void rodrigues_bwd(const float* r, const RodCtx& x, const float* gR, float* g_r) {
  float g_s = 0.f, g_omc = 0.f;
#pragma unroll
  for (int k = 0; k < 9; ++k) { g_s += gR[k] * x.K[k]; g_omc += gR[k] * x.K2[k]; }
  const float omc = 1.0f - x.c;
  // d(K K): gK = G K^T + K^T G with G = omc * gR
  float t1[9], t2[9], gK[9];
  mat3_mulT(gR, x.K, t1);
  matT3_mul(x.K, gR, t2);
#pragma unroll
  for (int k = 0; k < 9; ++k) gK[k] = x.s * gR[k] + omc * (t1[k] + t2[k]);
  float g_ang = g_s * x.c + g_omc * x.s;
  float gd[3];
  gd[0] = gK[7] - gK[5];
  gd[1] = gK[2] - gK[6];
  gd[2] = gK[3] - gK[1];
  const float inv = 1.0f / x.ang;
  g_ang -= (gd[0] * r[0] + gd[1] * r[1] + gd[2] * r[2]) * inv * inv;
#pragma unroll
  for (int k = 0; k < 3; ++k) g_r[k] = gd[k] * inv + g_ang * x.e[k] * inv;
}

// -------------------------------------------------------------------------------------------
// camera + projection (common/camera.py:456-474, common/transforms.py:316-329, :69-77,
// common/data_utils.py:361-365)
// -------------------------------------------------------------------------------------------
__device__ __forceinline__ float cam_tz(float s, float f, float img_res, float min_s) {
  const float sc = fmaxf(s, min_s);
  return __fdiv_rn(2.0f * f, __fadd_rn(__fmul_rn(img_res, sc), 1e-9f));
}
__device__ __forceinline__ float cam_tz_grad_s(float s, float f, float img_res, float min_s) {
  if (!(s >= min_s)) return 0.0f;  // torch.clamp passes the gradient where s >= min
  const float den = img_res * s + 1e-9f;
  return -2.0f * f * img_res / (den * den);
}

// X (3, camera space) -> uv (2). img_res > 0 additionally applies normalize_kp2d.
__device__ __forceinline__ void project_fwd(const float* K, const float* X, float img_res, float* uv) {
  const float px = K[0] * X[0] + K[1] * X[1] + K[2] * X[2];
  const float py = K[3] * X[0] + K[4] * X[1] + K[5] * X[2];
  const float pz = K[6] * X[0] + K[7] * X[1] + K[8] * X[2];
  float u = __fdiv_rn(px, pz), v = __fdiv_rn(py, pz);
  if (img_res > 0.0f) {
    u = __fsub_rn(__fdiv_rn(2.0f * u, img_res), 1.0f);
    v = __fsub_rn(__fdiv_rn(2.0f * v, img_res), 1.0f);
  }
  uv[0] = u; uv[1] = v;
}
__device__ __forceinline__ void project_bwd(const float* K, const float* X, float img_res, const float* g_uv, float* gX) {
  const float px = K[0] * X[0] + K[1] * X[1] + K[2] * X[2];
  const float py = K[3] * X[0] + K[4] * X[1] + K[5] * X[2];
  const float pz = K[6] * X[0] + K[7] * X[1] + K[8] * X[2];
  const float sc = img_res > 0.0f ? 2.0f / img_res : 1.0f;
  const float gu = g_uv[0] * sc, gv = g_uv[1] * sc;
  const float ipz = 1.0f / pz;
  const float gpx = gu * ipz, gpy = gv * ipz;
  const float gpz = -(gu * px + gv * py) * ipz * ipz;
  gX[0] = K[0] * gpx + K[3] * gpy + K[6] * gpz;
  gX[1] = K[1] * gpx + K[4] * gpy + K[7] * gpz;
  gX[2] = K[2] * gpx + K[5] * gpy + K[8] * gpz;
}

// ---- 6D rotation representation (Zhou et al., CVPR 2019): Gram-Schmidt of two 3-vectors ---------------------------
// The reference has three copies that differ only in memory layout (hb_rot6d_layout in the header):
//   ROWS        pytorch3d rotation_6d_to_matrix as called at src/nets/hand_heads/hand_hmr.py:85-87:
//               a1 = x[0:3], a2 = x[3:6], b1/b2/b3 are the ROWS of the matrix;
//   COLS        src/models/hamer_light/geometry.py:47-62 (reshape(-1,2,3).permute(0,2,1)) and
//               src/models/handoccnet_light/mano_head.py:132-141: a1 = x[0:3], a2 = x[3:6], b's are the COLUMNS;
//   COLS_PAIRED common/rot.py:367-381 (reshape(-1,3,2)): a1 = x[0,2,4], a2 = x[1,3,5], b's are the COLUMNS.
// F.normalize divides by max(||v||, 1e-12).
struct Rot6dCtx {
  float a2[3], b1[3], b2[3], u2[3];
  float n1, n2, d;
};

__device__ __forceinline__ void rot6d_split(const float* x, int layout, float* a1, float* a2) {
  if (layout == 2) { a1[0] = x[0]; a1[1] = x[2]; a1[2] = x[4]; a2[0] = x[1]; a2[1] = x[3]; a2[2] = x[5]; }
  else { a1[0] = x[0]; a1[1] = x[1]; a1[2] = x[2]; a2[0] = x[3]; a2[1] = x[4]; a2[2] = x[5]; }
}

// x (6) -> M (3x3 row-major)
__device__ __forceinline__ void rot6d_fwd(const float* x, int layout, float* M, Rot6dCtx& c) {
  float a1[3];
  rot6d_split(x, layout, a1, c.a2);
  c.n1 = fmaxf(sqrtf(a1[0] * a1[0] + a1[1] * a1[1] + a1[2] * a1[2]), 1e-12f);
#pragma unroll
  for (int k = 0; k < 3; ++k) c.b1[k] = __fdiv_rn(a1[k], c.n1);
  c.d = c.b1[0] * c.a2[0] + c.b1[1] * c.a2[1] + c.b1[2] * c.a2[2];
#pragma unroll
  for (int k = 0; k < 3; ++k) c.u2[k] = __fsub_rn(c.a2[k], __fmul_rn(c.d, c.b1[k]));
  c.n2 = fmaxf(sqrtf(c.u2[0] * c.u2[0] + c.u2[1] * c.u2[1] + c.u2[2] * c.u2[2]), 1e-12f);
#pragma unroll
  for (int k = 0; k < 3; ++k) c.b2[k] = __fdiv_rn(c.u2[k], c.n2);
  float b3[3];
  b3[0] = __fsub_rn(__fmul_rn(c.b1[1], c.b2[2]), __fmul_rn(c.b1[2], c.b2[1]));
  b3[1] = __fsub_rn(__fmul_rn(c.b1[2], c.b2[0]), __fmul_rn(c.b1[0], c.b2[2]));
  b3[2] = __fsub_rn(__fmul_rn(c.b1[0], c.b2[1]), __fmul_rn(c.b1[1], c.b2[0]));
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    if (layout == 0) { M[0 * 3 + k] = c.b1[k]; M[1 * 3 + k] = c.b2[k]; M[2 * 3 + k] = b3[k]; }
    else { M[k * 3 + 0] = c.b1[k]; M[k * 3 + 1] = c.b2[k]; M[k * 3 + 2] = b3[k]; }
  }
}

// gM (3x3 row-major) -> g_x (6)
__device__ __forceinline__ void rot6d_bwd(const Rot6dCtx& c, int layout, const float* gM, float* g_x) {
  float g1[3], g2[3], g3[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    if (layout == 0) { g1[k] = gM[0 * 3 + k]; g2[k] = gM[1 * 3 + k]; g3[k] = gM[2 * 3 + k]; }
    else { g1[k] = gM[k * 3 + 0]; g2[k] = gM[k * 3 + 1]; g3[k] = gM[k * 3 + 2]; }
  }
  // b3 = b1 x b2
  g1[0] += c.b2[1] * g3[2] - c.b2[2] * g3[1]; g1[1] += c.b2[2] * g3[0] - c.b2[0] * g3[2]; g1[2] += c.b2[0] * g3[1] - c.b2[1] * g3[0];
  g2[0] += g3[1] * c.b1[2] - g3[2] * c.b1[1]; g2[1] += g3[2] * c.b1[0] - g3[0] * c.b1[2]; g2[2] += g3[0] * c.b1[1] - g3[1] * c.b1[0];
  // b2 = u2 / max(|u2|, eps)
  float gu[3];
  const bool clamp2 = !(c.n2 > 1e-12f);
  const float p2 = clamp2 ? 0.0f : (c.b2[0] * g2[0] + c.b2[1] * g2[1] + c.b2[2] * g2[2]);
#pragma unroll
  for (int k = 0; k < 3; ++k) gu[k] = (g2[k] - c.b2[k] * p2) / c.n2;
  // u2 = a2 - d b1,  d = b1 . a2
  const float gd = -(c.b1[0] * gu[0] + c.b1[1] * gu[1] + c.b1[2] * gu[2]);
  float ga2[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) { ga2[k] = gu[k] + gd * c.b1[k]; g1[k] += -c.d * gu[k] + gd * c.a2[k]; }
  // b1 = a1 / max(|a1|, eps)
  const bool clamp1 = !(c.n1 > 1e-12f);
  const float p1 = clamp1 ? 0.0f : (c.b1[0] * g1[0] + c.b1[1] * g1[1] + c.b1[2] * g1[2]);
  float ga1[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) ga1[k] = (g1[k] - c.b1[k] * p1) / c.n1;
  if (layout == 2) { g_x[0] = ga1[0]; g_x[2] = ga1[1]; g_x[4] = ga1[2]; g_x[1] = ga2[0]; g_x[3] = ga2[1]; g_x[5] = ga2[2]; }
  else { g_x[0] = ga1[0]; g_x[1] = ga1[1]; g_x[2] = ga1[2]; g_x[3] = ga2[0]; g_x[4] = ga2[1]; g_x[5] = ga2[2]; }
}

}  // namespace hb
