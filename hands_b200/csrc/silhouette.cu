// Soft silhouette of the hand mesh: the consumer of `mano.v3d.cam.{r,l}` in the reference
// (src/models/hands_light/renderer.py:124-199 = pytorch3d MeshRasterizer + SoftSilhouetteShader, blur_radius =
// log(1/1e-6 - 1) * sigma, faces_per_pixel = 10, perspective_correct = False; flip_transpose_canvas :201-209), forward and
// backward.  pytorch3d is not part of the reference tree: the arithmetic restated here is documented in
// oracle/silhouette_oracle.py ("parity unpinned").
//
// Forward:  sil_setup_kernel   thread = (hand, face): project the three corners, build an 80-byte face record
//           sil_raster_kernel  CTA = (hand, band of 16 rows): faces culled to the band by ordered compaction (face order
//                              is the tie-break of the depth selection), then per warp footprint (8x4 pixels, handed
//                              out dynamically) 32 faces per ballot; thread = pixel
//                              keeps the K nearest candidates sorted in registers; writes the mask and, for the
//                              backward, (alpha, depth of the K-th fragment) per pixel.
// Backward: sil_face_bwd_kernel   warp = (hand, face): walks the face's pixel box, re-evaluates the SAME device function
//                                 (no float contraction freedom: explicit _rn intrinsics), keeps the pixels whose
//                                 selection contained the face (pz <= saved threshold), reduces the six corner gradients
//                                 with a fixed butterfly.
//           sil_vertex_bwd_kernel thread = (hand, vertex): sums its incident corners in CSR order, applies the
//                                 projection's Jacobian.  No atomics anywhere: bit-reproducible.
#include <math.h>
#include <vector>
#include "hb_common.cuh"

namespace hb {

constexpr int SIL_K = HB_SIL_FACES_PER_PIXEL;   // 10
constexpr int SIL_REC = 5;                      // float4 per face record
constexpr int SIL_TILE = 16;
constexpr float SIL_EPS = 1e-8f;                // pytorch3d kEpsilon

struct SilFace {
  float v0x, v0y, v1x, v1y, v2x, v2y, z0, z1, z2;
  float inv_area;          // 1 / (edge(v2; v0, v1) + eps)
  float il01, il02, il12;  // 1 / |edge|^2, or -1 for a degenerate edge (|edge|^2 <= eps)
  float xlo, xhi, ylo, yhi;   // NDC box widened by sqrt(blur_radius)
};

__device__ __forceinline__ float pix_to_ndc(int i, float S) {
  return __fadd_rn(-1.0f, __fdiv_rn(__fmaf_rn(2.0f, (float)i, 1.0f), S));
}

__device__ __forceinline__ float edge_fn(float px, float py, float ax, float ay, float bx, float by) {
  return __fsub_rn(__fmul_rn(__fsub_rn(px, ax), __fsub_rn(by, ay)), __fmul_rn(__fsub_rn(py, ay), __fsub_rn(bx, ax)));
}

// squared distance from p to the segment a + t e, t in [0,1]; (dx,dy) = p - a in, p - p_proj out; tt = clamped parameter
// (PointLineDistanceForward).  il = 1/|e|^2, or -1 for a degenerate edge (distance to its end point b = a + e).
__device__ __forceinline__ float seg_dist(float ex, float ey, float il, float& dx, float& dy, float& tt) {
  if (il < 0.0f) {
    tt = 1.0f;
  } else {
    const float t = __fmul_rn(__fmaf_rn(ex, dx, __fmul_rn(ey, dy)), il);
    tt = fminf(fmaxf(t, 0.0f), 1.0f);
  }
  dx = __fmaf_rn(-tt, ex, dx);
  dy = __fmaf_rn(-tt, ey, dy);
  return __fmaf_rn(dx, dx, __fmul_rn(dy, dy));
}

struct SilHit {
  float pz, dist, d01, d02, d12;
  bool inside;
};

// CheckPixelInsideFace without the selection bookkeeping.  false = the face is no candidate at this pixel.
// Every operation is an explicit round-to-nearest intrinsic: forward and backward kernels inline this function separately
// and must agree to the bit on pz (the selection test) and on the distances.
__device__ __forceinline__ bool sil_eval(const SilFace& f, float px, float py, float blur, SilHit& h) {
  if (px > f.xhi || px < f.xlo || py > f.yhi || py < f.ylo) return false;
  const float e01x = __fsub_rn(f.v1x, f.v0x), e01y = __fsub_rn(f.v1y, f.v0y);
  const float e02x = __fsub_rn(f.v2x, f.v0x), e02y = __fsub_rn(f.v2y, f.v0y);
  const float e12x = __fsub_rn(f.v2x, f.v1x), e12y = __fsub_rn(f.v2y, f.v1y);
  const float p0x = __fsub_rn(px, f.v0x), p0y = __fsub_rn(py, f.v0y);
  const float p1x = __fsub_rn(px, f.v1x), p1y = __fsub_rn(py, f.v1y);
  const float p2x = __fsub_rn(px, f.v2x), p2y = __fsub_rn(py, f.v2y);
  // barycentric coordinates: edge(p; v1,v2), edge(p; v2,v0), edge(p; v0,v1) over the face's edge(v2; v0,v1) + eps
  const float w0 = __fmul_rn(__fmaf_rn(p1x, e12y, -__fmul_rn(p1y, e12x)), f.inv_area);
  const float w1 = __fmul_rn(__fmaf_rn(p2y, e02x, -__fmul_rn(p2x, e02y)), f.inv_area);
  const float w2 = __fmul_rn(__fmaf_rn(p0x, e01y, -__fmul_rn(p0y, e01x)), f.inv_area);
  h.inside = w0 > 0.0f && w1 > 0.0f && w2 > 0.0f;
  float dx, dy, tt;
  dx = p0x; dy = p0y;
  h.d01 = seg_dist(e01x, e01y, f.il01, dx, dy, tt);
  dx = p0x; dy = p0y;
  h.d02 = seg_dist(e02x, e02y, f.il02, dx, dy, tt);
  dx = p1x; dy = p1y;
  h.d12 = seg_dist(e12x, e12y, f.il12, dx, dy, tt);
  h.dist = fminf(fminf(h.d01, h.d02), h.d12);
  if (!(h.inside || h.dist < blur)) return false;
  const float c0 = fminf(fmaxf(w0, 0.0f), 1.0f), c1 = fminf(fmaxf(w1, 0.0f), 1.0f), c2 = fminf(fmaxf(w2, 0.0f), 1.0f);
  const float inv = __frcp_rn(fmaxf(__fadd_rn(__fadd_rn(c0, c1), c2), 1e-5f));
  h.pz = __fmul_rn(__fmaf_rn(c2, f.z2, __fmaf_rn(c1, f.z1, __fmul_rn(c0, f.z0))), inv);
  return h.pz >= 0.0f;
}

// sigmoid(-sd / sigma) as torch evaluates it in fp32
__device__ __forceinline__ float sil_prob(float sd, float inv_sigma) {
  return __frcp_rn(__fadd_rn(1.0f, expf(__fmul_rn(sd, inv_sigma))));
}

__device__ __forceinline__ SilFace load_face(const float4* __restrict__ R) {
  const float4 a = __ldg(R), b = __ldg(R + 1), c = __ldg(R + 2), d = __ldg(R + 3), e = __ldg(R + 4);
  SilFace f;
  f.v0x = a.x; f.v0y = a.y; f.v1x = a.z; f.v1y = a.w;
  f.v2x = b.x; f.v2y = b.y; f.z0 = b.z; f.z1 = b.w;
  f.z2 = c.x; f.inv_area = c.y; f.il01 = c.z; f.il02 = c.w;
  f.il12 = d.x; f.xlo = d.y; f.xhi = d.z; f.ylo = d.w;
  f.yhi = e.x;
  return f;
}

// ---------------------------------------------------------------------------------------------------------------------
// NDC intrinsics of renderer.py:171-175,187-190: K' = [[2/S,0,-1],[0,2/S,-1],[0,0,1]] @ K, focal = diag, principal = K'[:2,2]
struct SilCam { float fx, fy, px, py; };
__device__ __forceinline__ SilCam sil_cam(const float* __restrict__ K, float c) {
  SilCam k;
  k.fx = __fsub_rn(__fmul_rn(c, __ldg(K + 0)), __ldg(K + 6));
  k.fy = __fsub_rn(__fmul_rn(c, __ldg(K + 4)), __ldg(K + 7));
  k.px = __fsub_rn(__fmul_rn(c, __ldg(K + 2)), __ldg(K + 8));
  k.py = __fsub_rn(__fmul_rn(c, __ldg(K + 5)), __ldg(K + 8));
  return k;
}

// exact pixel index range [ilo, ihi] of the centres inside the NDC interval [lo, hi]: an estimate from the inverse map,
// then stepped against the very pix_to_ndc values the per-pixel test compares with (empty: ilo > ihi)
__device__ __forceinline__ void pix_range(float lo, float hi, int S, int& ilo, int& ihi) {
  const float Sf = (float)S;
  int a = (int)fminf(fmaxf(ceilf((lo + 1.0f) * 0.5f * Sf - 0.5f), 0.0f), Sf);
  int b = (int)fminf(fmaxf(floorf((hi + 1.0f) * 0.5f * Sf - 0.5f), -1.0f), Sf - 1.0f);
  while (a > 0 && pix_to_ndc(a - 1, Sf) >= lo) --a;
  while (a < S && pix_to_ndc(a, Sf) < lo) ++a;
  while (b < S - 1 && pix_to_ndc(b + 1, Sf) <= hi) ++b;
  while (b >= 0 && pix_to_ndc(b, Sf) > hi) --b;
  ilo = a;
  ihi = b;
}

__global__ void __launch_bounds__(128) sil_setup_kernel(const float* __restrict__ verts, const float* __restrict__ K, const int* __restrict__ faces,
                                                        int F, int V, int S, float sqrt_blur, float4* __restrict__ rec) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (f >= F) return;
  const SilCam cam = sil_cam(K + (size_t)b * 9, 2.0f / (float)S);
  float x[3], y[3], z[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float* v = verts + ((size_t)b * V + __ldg(faces + 3 * f + k)) * 3;
    const float X = __ldg(v), Y = __ldg(v + 1), Z = __ldg(v + 2);
    x[k] = __fdiv_rn(__fadd_rn(__fmul_rn(X, cam.fx), __fmul_rn(Z, cam.px)), Z);
    y[k] = __fdiv_rn(__fadd_rn(__fmul_rn(Y, cam.fy), __fmul_rn(Z, cam.py)), Z);
    z[k] = Z;
  }
  const float zmax = fmaxf(fmaxf(z[0], z[1]), z[2]);
  const float area = edge_fn(x[0], y[0], x[1], y[1], x[2], y[2]);
  const bool finite = isfinite(x[0]) && isfinite(x[1]) && isfinite(x[2]) && isfinite(y[0]) && isfinite(y[1]) && isfinite(y[2]) && isfinite(area);
  const bool valid = finite && !(zmax < 0.0f) && !(area <= SIL_EPS && area >= -SIL_EPS);
  float xlo = fminf(fminf(x[0], x[1]), x[2]) - sqrt_blur, xhi = fmaxf(fmaxf(x[0], x[1]), x[2]) + sqrt_blur;
  float ylo = fminf(fminf(y[0], y[1]), y[2]) - sqrt_blur, yhi = fmaxf(fmaxf(y[0], y[1]), y[2]) + sqrt_blur;
  int clo = S, chi = -1, rlo = S, rhi = -1;
  if (valid) {
    pix_range(xlo, xhi, S, clo, chi);
    pix_range(ylo, yhi, S, rlo, rhi);
  } else {
    xlo = ylo = INFINITY;
    xhi = yhi = -INFINITY;
  }
  auto inv_len2 = [](float ax, float ay, float bx, float by) {
    const float ex = __fsub_rn(bx, ax), ey = __fsub_rn(by, ay);
    const float l2 = __fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey));
    return l2 <= SIL_EPS ? -1.0f : __frcp_rn(l2);
  };
  const float barea = __fadd_rn(edge_fn(x[2], y[2], x[0], y[0], x[1], y[1]), SIL_EPS);
  float4* R = rec + ((size_t)b * F + f) * SIL_REC;
  R[0] = make_float4(x[0], y[0], x[1], y[1]);
  R[1] = make_float4(x[2], y[2], z[0], z[1]);
  R[2] = make_float4(z[2], valid ? __frcp_rn(barea) : 0.0f, inv_len2(x[0], y[0], x[1], y[1]), inv_len2(x[0], y[0], x[2], y[2]));
  R[3] = make_float4(inv_len2(x[1], y[1], x[2], y[2]), xlo, xhi, ylo);
  R[4] = make_float4(yhi, __int_as_float(clo | (chi + 1) << 16), __int_as_float(rlo | (rhi + 1) << 16), 0.0f);   // [lo, hi+1) packed
}

// Ordered compaction: appends to out, in ascending i, value(i) for every i in [0,n) with keep(i).  All 256 threads call it.
template <class Keep>
__device__ __forceinline__ int compact_ordered(int n, Keep keep, unsigned short* out, int* s_wc) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int count = 0;
  for (int base = 0; base < n; base += 256) {
    const int i = base + tid;
    int val = 0;
    const bool k = i < n && keep(i, val);
    const unsigned m = __ballot_sync(0xffffffffu, k);
    if (lane == 0) s_wc[warp] = __popc(m);
    __syncthreads();
    int off = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      const int c = s_wc[w];
      off += w < warp ? c : 0;
      tot += c;
    }
    if (k) out[count + off + __popc(m & ((1u << lane) - 1u))] = (unsigned short)val;
    count += tot;
    __syncthreads();
  }
  return count;
}

__global__ void __launch_bounds__(256) sil_raster_kernel(const float4* __restrict__ rec, int F, int S, float sigma, float blur,
                                                         float* __restrict__ mask, float2* __restrict__ frag) {
  extern __shared__ float s_dyn[];              // pixel-centre table [S + 16], then the band list [F] (unsigned short)
  __shared__ int s_wc[8];
  float* pn = s_dyn;
  unsigned short* blist = reinterpret_cast<unsigned short*>(s_dyn + S + 16);
  const int b = blockIdx.y, r0 = blockIdx.x * SIL_TILE;
  const float4* R = rec + (size_t)b * F * SIL_REC;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float Sf = (float)S, inv_sigma = 1.0f / sigma;
  for (int i = tid; i < S + 16; i += 256) pn[i] = pix_to_ndc(i, Sf);   // one IEEE division per entry instead of per pixel

  __shared__ int s_cmin, s_cmax, s_next;
  if (tid == 0) { s_cmin = S; s_cmax = 0; s_next = 0; }
  __syncthreads();
  int cmin = S, cmax = 0;   // column extent [cmin, cmax) of the band's faces: footprints outside it are empty
  const int nb = compact_ordered(F, [&](int i, int& val) {
    const float4 e = __ldg(&R[i * SIL_REC + 4]);
    const int rows = __float_as_int(e.z), cols = __float_as_int(e.y);
    val = i;
    const bool keep = (rows & 0xffff) < r0 + SIL_TILE && (rows >> 16) > r0 && (cols >> 16) > (cols & 0xffff);
    if (keep) { cmin = min(cmin, cols & 0xffff); cmax = max(cmax, cols >> 16); }
    return keep;
  }, blist, s_wc);
  if (cmin < cmax) { atomicMin(&s_cmin, cmin); atomicMax(&s_cmax, cmax); }
  __syncthreads();
  cmin = s_cmin;
  cmax = s_cmax;

  // Footprints of 8 columns x 4 rows (one warp each), handed out dynamically: the band's warps never wait for each other
  // (which warp renders a footprint does not change its pixels).
  const int fcols = (S + 7) / 8, nfp = fcols * (SIL_TILE / 4);
  for (;;) {
    int fp = 0;
    if (lane == 0) fp = atomicAdd(&s_next, 1);
    fp = __shfl_sync(0xffffffffu, fp, 0);
    if (fp >= nfp) break;
    const int fr = r0 + (fp / fcols) * 4, fc = (fp % fcols) * 8;
    const int row = fr + (lane >> 3), col = fc + (lane & 7);
    const float px = pn[col], py = pn[row];
    const float fx0 = pn[fc], fx1 = pn[fc + 7], fy0 = pn[fr], fy1 = pn[fr + 3];
    float qz[SIL_K], qd[SIL_K];
#pragma unroll
    for (int k = 0; k < SIL_K; ++k) { qz[k] = INFINITY; qd[k] = 0.0f; }
    const int nscan = (fc < cmax && fc + 8 > cmin) ? nb : 0;
    for (int base = 0; base < nscan; base += 32) {
      // 32 faces of the band list against the footprint at once; the survivors are visited in list (= face) order
      const int j = base + lane;
      int fidx = 0;
      bool hit = false;
      if (j < nscan) {
        fidx = blist[j];
        const float4 d = __ldg(R + fidx * SIL_REC + 3);
        const float yhi = __ldg(R + fidx * SIL_REC + 4).x;
        hit = !(d.y > fx1 || d.z < fx0 || d.w > fy1 || yhi < fy0);
      }
      unsigned m = __ballot_sync(0xffffffffu, hit);
      while (m) {
        const int src = __ffs(m) - 1;
        m &= m - 1;
        const SilFace f = load_face(R + __shfl_sync(0xffffffffu, fidx, src) * SIL_REC);
        SilHit h;
        if (!sil_eval(f, px, py, blur, h)) continue;
        float z = h.pz, sd = h.inside ? -h.dist : h.dist;
        if (z < qz[SIL_K - 1]) {   // sorted insertion; a tie stays behind the earlier face
#pragma unroll
          for (int k = 0; k < SIL_K; ++k) {
            const bool sw = z < qz[k];
            const float tz = qz[k], td = qd[k];
            qz[k] = sw ? z : tz;
            qd[k] = sw ? sd : td;
            z = sw ? tz : z;
            sd = sw ? td : sd;
          }
        }
      }
    }
    float alpha = 1.0f;
#pragma unroll
    for (int k = 0; k < SIL_K; ++k)   // in depth order, as sigmoid_alpha_blend multiplies
      if (qz[k] < INFINITY) alpha = __fmul_rn(alpha, __fsub_rn(1.0f, sil_prob(qd[k], inv_sigma)));
    if (row < S && col < S) {
      const size_t o = ((size_t)b * S + row) * S + col;
      mask[o] = __fsub_rn(1.0f, alpha);
      frag[o] = make_float2(alpha, qz[SIL_K - 1]);
    }
  }
}

__global__ void __launch_bounds__(256) sil_face_bwd_kernel(const float4* __restrict__ rec, const float2* __restrict__ frag, const float* __restrict__ g_mask,
                                                           int F, int S, float sigma, float blur, float* __restrict__ gface) {
  extern __shared__ float pn[];   // pixel-centre table [S]
  for (int i = threadIdx.x; i < S; i += 256) pn[i] = pix_to_ndc(i, (float)S);
  __syncthreads();
  const int lane = threadIdx.x & 31, f = blockIdx.x * 8 + (threadIdx.x >> 5), b = blockIdx.y;
  if (f >= F) return;
  const float4* Rf = rec + ((size_t)b * F + f) * SIL_REC;
  const SilFace fc = load_face(Rf);
  const float4 e = __ldg(Rf + 4);
  const int cols = __float_as_int(e.y), rows = __float_as_int(e.z);
  const int clo = cols & 0xffff, w = (cols >> 16) - clo, rlo = rows & 0xffff, hgt = (rows >> 16) - rlo;
  float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const float inv_sigma = 1.0f / sigma;
  if (w > 0 && hgt > 0) {
    const int n = w * hgt;
    const unsigned magic = w > 1 ? 0xFFFFFFFFu / (unsigned)w + 1u : 0u;   // i / w for i * w < 2^32
    for (int i = lane; i < n; i += 32) {
      const int q = w > 1 ? (int)__umulhi((unsigned)i, magic) : i;
      const int r = rlo + q, c = clo + (i - q * w);
      const size_t o = ((size_t)b * S + r) * S + c;
      const float g = __ldg(g_mask + o);
      const float2 fr = __ldg(frag + o);
      const float ga = __fmul_rn(g, fr.x);
      if (ga == 0.0f) continue;
      const float px = pn[c], py = pn[r];
      SilHit h;
      if (!sil_eval(fc, px, py, blur, h)) continue;
      if (!(h.pz <= fr.y)) continue;                      // the face was not among the pixel's K nearest
      const float sd = h.inside ? -h.dist : h.dist;
      const float p = sil_prob(sd, inv_sigma);
      // d mask / d sd = prod_{m != k}(1 - p_m) * (-p (1 - p) / sigma) = -alpha * p / sigma ; then signed -> absolute distance
      float gd = -ga * p * inv_sigma;
      if (h.inside) gd = -gd;
      // PointTriangleDistanceBackward: the nearest edge, ties resolved e01, e02, e12
      int ia, ib;
      float ax, ay, bx, by, il;
      if (h.d01 <= h.d02 && h.d01 <= h.d12) { ia = 0; ib = 1; ax = fc.v0x; ay = fc.v0y; bx = fc.v1x; by = fc.v1y; il = fc.il01; }
      else if (h.d02 <= h.d01 && h.d02 <= h.d12) { ia = 0; ib = 2; ax = fc.v0x; ay = fc.v0y; bx = fc.v2x; by = fc.v2y; il = fc.il02; }
      else { ia = 1; ib = 2; ax = fc.v1x; ay = fc.v1y; bx = fc.v2x; by = fc.v2y; il = fc.il12; }
      float tt, dx = __fsub_rn(px, ax), dy = __fsub_rn(py, ay);
      seg_dist(__fsub_rn(bx, ax), __fsub_rn(by, ay), il, dx, dy, tt);   // (dx,dy) = p - p_proj
      const float sx = -2.0f * gd * dx, sy = -2.0f * gd * dy;   // gd * 2 * (p_proj - p)
      const float wa = il < 0.0f ? 0.0f : 1.0f - tt, wb = il < 0.0f ? 1.0f : tt;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float wk = (k == ia ? wa : 0.0f) + (k == ib ? wb : 0.0f);
        acc[2 * k] = fmaf(wk, sx, acc[2 * k]);
        acc[2 * k + 1] = fmaf(wk, sy, acc[2 * k + 1]);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 6; ++k) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], s);
  }
  if (lane < 6) {
    float v = acc[0];
#pragma unroll
    for (int k = 1; k < 6; ++k) v = lane == k ? acc[k] : v;
    gface[((size_t)b * F + f) * 6 + lane] = v;
  }
}

__global__ void __launch_bounds__(128) sil_vertex_bwd_kernel(const float* __restrict__ verts, const float* __restrict__ K, const float* __restrict__ gface,
                                                             const int* __restrict__ adj_off, const int* __restrict__ adj, int F, int V, int S,
                                                             float* __restrict__ g_verts) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (v >= V) return;
  float gx = 0.0f, gy = 0.0f;
  const float* G = gface + (size_t)b * F * 6;
  for (int i = __ldg(adj_off + v); i < __ldg(adj_off + v + 1); ++i) {
    const int code = __ldg(adj + i);   // face * 3 + corner
    gx += __ldg(G + 2 * code);
    gy += __ldg(G + 2 * code + 1);
  }
  const SilCam cam = sil_cam(K + (size_t)b * 9, 2.0f / (float)S);
  const float* p = verts + ((size_t)b * V + v) * 3;
  const float X = p[0], Y = p[1], Z = p[2];
  const float iz = 1.0f / Z;
  const float ax = gx * cam.fx * iz, ay = gy * cam.fy * iz;
  float* o = g_verts + ((size_t)b * V + v) * 3;
  const bool live = gx != 0.0f || gy != 0.0f;   // an untouched vertex gets an exact zero even when Z is degenerate
  o[0] = live ? ax : 0.0f;
  o[1] = live ? ay : 0.0f;
  o[2] = live ? -(ax * X + ay * Y) * iz : 0.0f;
}

// render_loss (src/utils/loss_modules.py:146-152) with the gate of loss_arctic_sf.py:179-182: mean over (B, n) of
// |pred - gt| * valid[b] * gate[b].  One CTA per sample, fixed-order tree; a second launch sums the B partials in order.
__global__ void __launch_bounds__(256) mask_l1_fwd_kernel(const float* __restrict__ pred, const float* __restrict__ gt, const float* __restrict__ valid,
                                                          const float* __restrict__ gate, int n, float* __restrict__ partial) {
  __shared__ float sh[256];
  const int b = blockIdx.x;
  const float* p = pred + (size_t)b * n;
  const float* t = gt + (size_t)b * n;
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) acc += fabsf(__ldg(p + i) - __ldg(t + i));
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int m = 128; m > 0; m >>= 1) {
    if (threadIdx.x < m) sh[threadIdx.x] += sh[threadIdx.x + m];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[b] = sh[0] * (valid ? __ldg(valid + b) : 1.f) * (gate ? __ldg(gate + b) : 1.f);
}

__global__ void __launch_bounds__(1024) mask_l1_sum_kernel(const float* __restrict__ partial, int B, float scale, float* __restrict__ loss) {
  __shared__ float sh[1024];
  float acc = 0.f;
  for (int b = threadIdx.x; b < B; b += 1024) acc += partial[b];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int m = 512; m > 0; m >>= 1) {
    if (threadIdx.x < m) sh[threadIdx.x] += sh[threadIdx.x + m];
    __syncthreads();
  }
  if (threadIdx.x == 0) loss[0] = sh[0] * scale;
}

__global__ void __launch_bounds__(256) mask_l1_bwd_kernel(const float* __restrict__ pred, const float* __restrict__ gt, const float* __restrict__ valid,
                                                          const float* __restrict__ gate, const float* __restrict__ g_loss, int n, float scale,
                                                          float* __restrict__ g_pred) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const float k = __ldg(g_loss) * scale * (valid ? __ldg(valid + b) : 1.f) * (gate ? __ldg(gate + b) : 1.f);
  const size_t o = (size_t)b * n + i;
  const float d = __ldg(pred + o) - __ldg(gt + o);
  g_pred[o] = d > 0.f ? k : (d < 0.f ? -k : 0.f);   // torch's l1_loss backward: sign(pred - gt), 0 at equality
}

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace hb

using namespace hb;

struct hb_sil {
  int F, V, device;
  int* faces;     // device (F,3)
  int* adj_off;   // device (V+1)
  int* adj;       // device (3F)  face*3 + corner, ascending per vertex
};

extern "C" int hb_sil_create(const int32_t* faces_host, int n_faces, int n_verts, int device, hb_sil** out) {
  if (!faces_host || !out || n_faces <= 0 || n_verts <= 0) { set_error("hb_sil_create: NULL or empty argument"); return HB_E_ARG; }
  if (n_faces > HB_SIL_MAX_FACES) { set_error("hb_sil_create: at most %d faces (got %d)", HB_SIL_MAX_FACES, n_faces); return HB_E_UNSUPPORTED; }
  std::vector<int> off(n_verts + 1, 0), adj(3 * (size_t)n_faces);
  for (int i = 0; i < 3 * n_faces; ++i) {
    const int v = faces_host[i];
    if (v < 0 || v >= n_verts) { set_error("hb_sil_create: faces[%d]=%d outside [0,%d)", i, v, n_verts); return HB_E_ARG; }
    ++off[v + 1];
  }
  for (int v = 0; v < n_verts; ++v) off[v + 1] += off[v];
  std::vector<int> cur(off.begin(), off.end() - 1);
  for (int i = 0; i < 3 * n_faces; ++i) adj[cur[faces_host[i]]++] = i;
  HB_CUDA(cudaSetDevice(device));
  hb_sil* h = new hb_sil();
  h->F = n_faces; h->V = n_verts; h->device = device;
  h->faces = h->adj_off = h->adj = nullptr;
  cudaError_t e;
  if ((e = cudaMalloc(&h->faces, sizeof(int) * 3 * n_faces)) != cudaSuccess || (e = cudaMalloc(&h->adj_off, sizeof(int) * (n_verts + 1))) != cudaSuccess ||
      (e = cudaMalloc(&h->adj, sizeof(int) * 3 * n_faces)) != cudaSuccess) {
    cudaFree(h->faces); cudaFree(h->adj_off); cudaFree(h->adj);
    delete h;
    set_error("hb_sil_create: cudaMalloc failed: %s", cudaGetErrorString(e));
    return (int)e;
  }
  HB_CUDA(cudaMemcpy(h->faces, faces_host, sizeof(int) * 3 * n_faces, cudaMemcpyHostToDevice));
  HB_CUDA(cudaMemcpy(h->adj_off, off.data(), sizeof(int) * (n_verts + 1), cudaMemcpyHostToDevice));
  HB_CUDA(cudaMemcpy(h->adj, adj.data(), sizeof(int) * 3 * n_faces, cudaMemcpyHostToDevice));
  *out = h;
  return 0;
}

extern "C" int hb_sil_destroy(hb_sil* h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  cudaFree(h->faces); cudaFree(h->adj_off); cudaFree(h->adj);
  delete h;
  return 0;
}

namespace {
struct SilWs { float4* rec; float2* frag; float* gface; size_t bytes; };
SilWs sil_ws(const hb_sil* h, int B, int S, void* base) {
  SilWs w;
  char* p = (char*)base;
  const size_t rec = align256((size_t)B * h->F * SIL_REC * sizeof(float4));
  const size_t frag = align256((size_t)B * S * S * sizeof(float2));
  const size_t gf = align256((size_t)B * h->F * 6 * sizeof(float));
  w.rec = (float4*)p;
  w.frag = (float2*)(p + rec);
  w.gface = (float*)(p + rec + frag);
  w.bytes = rec + frag + gf;
  return w;
}
int sil_args(const char* who, const hb_sil* h, const void* verts, const void* K, const void* img, int B, int S, float sigma, float blur,
             const void* ws, size_t ws_bytes) {
  if (!h || !verts || !K || !img || B < 0) { set_error("%s: NULL argument or negative batch", who); return HB_E_ARG; }
  if (S <= 0 || S > 4096 || !(sigma > 0.0f) || !(blur >= 0.0f)) { set_error("%s: img_res in [1,4096], sigma > 0, blur_radius >= 0 required", who); return HB_E_ARG; }
  if (B > 65535) { set_error("%s: at most 65535 meshes per call (got %d)", who, B); return HB_E_UNSUPPORTED; }
  if (B > 0 && (!ws || ws_bytes < sil_ws(h, B, S, nullptr).bytes)) { set_error("%s: workspace too small (%zu < %zu bytes)", who, ws_bytes, sil_ws(h, B, S, nullptr).bytes); return HB_E_WORKSPACE; }
  if ((reinterpret_cast<uintptr_t>(ws) & 15u) != 0) { set_error("%s: workspace must be 16-byte aligned", who); return HB_E_ALIGN; }
  return 0;
}
}  // namespace

extern "C" size_t hb_sil_workspace_bytes(const hb_sil* h, int n_meshes, int img_res) {
  if (!h || n_meshes <= 0 || img_res <= 0) return 0;
  return sil_ws(h, n_meshes, img_res, nullptr).bytes;
}

extern "C" int hb_sil_fwd(const hb_sil* h, const float* verts_cam, const float* K, int n_meshes, int img_res, float sigma, float blur_radius,
                          float* mask, void* workspace, size_t workspace_bytes, void* stream) {
  if (int rc = sil_args("hb_sil_fwd", h, verts_cam, K, mask, n_meshes, img_res, sigma, blur_radius, workspace, workspace_bytes)) return rc;
  if (n_meshes == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const SilWs w = sil_ws(h, n_meshes, img_res, workspace);
  sil_setup_kernel<<<dim3((h->F + 127) / 128, n_meshes), 128, 0, st>>>(verts_cam, K, h->faces, h->F, h->V, img_res, sqrtf(blur_radius), w.rec);
  ++g_launches;
  if (int rc = check_launch("sil_setup_kernel")) return rc;
  sil_raster_kernel<<<dim3((img_res + SIL_TILE - 1) / SIL_TILE, n_meshes), 256, sizeof(float) * (img_res + 16) + sizeof(unsigned short) * h->F, st>>>(w.rec, h->F, img_res, sigma, blur_radius, mask, w.frag);
  ++g_launches;
  return check_launch("sil_raster_kernel");
}

extern "C" int hb_sil_bwd(const hb_sil* h, const float* verts_cam, const float* K, const float* g_mask, int n_meshes, int img_res, float sigma,
                          float blur_radius, void* workspace, size_t workspace_bytes, float* g_verts, void* stream) {
  if (int rc = sil_args("hb_sil_bwd", h, verts_cam, K, g_mask, n_meshes, img_res, sigma, blur_radius, workspace, workspace_bytes)) return rc;
  if (!g_verts) { set_error("hb_sil_bwd: g_verts is NULL"); return HB_E_ARG; }
  if (n_meshes == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const SilWs w = sil_ws(h, n_meshes, img_res, workspace);
  sil_face_bwd_kernel<<<dim3((h->F + 7) / 8, n_meshes), 256, sizeof(float) * img_res, st>>>(w.rec, w.frag, g_mask, h->F, img_res, sigma, blur_radius, w.gface);
  ++g_launches;
  if (int rc = check_launch("sil_face_bwd_kernel")) return rc;
  sil_vertex_bwd_kernel<<<dim3((h->V + 127) / 128, n_meshes), 128, 0, st>>>(verts_cam, K, w.gface, h->adj_off, h->adj, h->F, h->V, img_res, g_verts);
  ++g_launches;
  return check_launch("sil_vertex_bwd_kernel");
}

extern "C" int hb_mask_l1_loss_fwd(const float* pred, const float* gt, const float* valid, const float* gate, int B, int n, float* partial,
                                   float* loss, void* stream) {
  if (B < 0 || n <= 0 || !loss || (B > 0 && (!pred || !gt || !partial))) { set_error("hb_mask_l1_loss_fwd: bad argument"); return HB_E_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  if (B > 0) {
    mask_l1_fwd_kernel<<<B, 256, 0, st>>>(pred, gt, valid, gate, n, partial);
    ++g_launches;
    if (int rc = check_launch("mask_l1_fwd_kernel")) return rc;
  }
  mask_l1_sum_kernel<<<1, 1024, 0, st>>>(partial, B, B > 0 ? 1.0f / ((float)B * (float)n) : 0.0f, loss);
  ++g_launches;
  return check_launch("mask_l1_sum_kernel");
}

extern "C" int hb_mask_l1_loss_bwd(const float* pred, const float* gt, const float* valid, const float* gate, const float* g_loss, int B, int n,
                                   float* g_pred, void* stream) {
  if (B < 0 || n <= 0 || (B > 0 && (!pred || !gt || !g_loss || !g_pred))) { set_error("hb_mask_l1_loss_bwd: bad argument"); return HB_E_ARG; }
  if (B == 0) return 0;
  if (B > 65535) { set_error("hb_mask_l1_loss_bwd: at most 65535 samples per call (got %d)", B); return HB_E_UNSUPPORTED; }
  mask_l1_bwd_kernel<<<dim3((n + 255) / 256, B), 256, 0, (cudaStream_t)stream>>>(pred, gt, valid, gate, g_loss, n, 1.0f / ((float)B * (float)n), g_pred);
  ++g_launches;
  return check_launch("mask_l1_bwd_kernel");
}
