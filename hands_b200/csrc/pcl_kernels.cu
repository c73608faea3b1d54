// Perspective Crop Layer (src/datasets/hands_light_dataset.py:354-467), batched on the GPU, fwd + bwd.
//
// The reference runs, per sample on the CPU:  float64 homography -> fp32 grid (linspace, 3x3 matmul,
// divide, normalise) -> F.grid_sample(bilinear, zeros, align_corners=False) to s x s ->
// F.interpolate(bilinear, align_corners=True) to R x R.  Here:
//   pcl_setup_kernel : the float64 homography, one thread per crop                  (lines 357-386, 425-454)
//   pcl_fwd_kernel   : one CTA per (crop, 16 output rows): the needed rows of the s x s intermediate are
//                      gathered once into shared memory, then resized from there; R x R stores coalesced.
//   pcl_bwd_mid      : transposed resize, gather form  (g_out -> g_mid, the s x s intermediate gradient)
//   pcl_bwd_img      : transposed grid_sample, gather form through the inverse homography: each source
//                      pixel collects from the few intermediate pixels whose bilinear footprint covers it,
//                      so g_img is written exactly once -- no atomics, no memset, deterministic.
// The fp32 operation order (which products are fused) follows torch's CPU kernels exactly; it was pinned
// by bit-comparing a numpy emulation against torch 2.11 single-threaded ("PCL exactness").
#include "hb_common.cuh"

namespace hb {

constexpr int PF = HB_PCL_PARAM_FLOATS;
// record layout (floats): [0..8] P, [9..17] Pinv, [18] s (int bits), [19] resize scale, [20] linspace step,
// [21] g_mid element offset inside the chunk workspace (int bits)

struct Crop {
  float P[9];
  int s;
  float scale, step;
};

__device__ __forceinline__ Crop load_crop(const float* __restrict__ rec) {
  Crop c;
#pragma unroll
  for (int k = 0; k < 9; ++k) c.P[k] = __ldg(rec + k);
  c.s = __float_as_int(__ldg(rec + 18));
  c.scale = __ldg(rec + 19);
  c.step = __ldg(rec + 20);
  return c;
}

// torch.linspace(0, 1, s)[i]  (CPU kernel: start + step*i below the midpoint, fma(-step, s-1-i, end) above)
__device__ __forceinline__ float lin01(const Crop& c, int i) {
  if (c.s == 1) return 0.0f;
  return (i < c.s / 2) ? __fmul_rn(c.step, (float)i) : fmaf(-c.step, (float)(c.s - 1 - i), 1.0f);
}

// source-image sample position (in pixel-index units) of intermediate pixel (row j, col i)
__device__ __forceinline__ void sample_pos(const Crop& c, int j, int i, float R, float& ix, float& iy) {
  const float u = lin01(c, i), v = lin01(c, j);
  const float X = __fadd_rn(fmaf(c.P[1], v, __fmul_rn(c.P[0], u)), c.P[2]);
  const float Y = __fadd_rn(fmaf(c.P[4], v, __fmul_rn(c.P[3], u)), c.P[5]);
  const float Z = __fadd_rn(fmaf(c.P[7], v, __fmul_rn(c.P[6], u)), c.P[8]);
  const float den = __fadd_rn(1e-8f, Z);
  const float gx = __fsub_rn(__fmul_rn(__fdiv_rn(__fdiv_rn(X, den), R), 2.0f), 1.0f);
  const float gy = __fsub_rn(__fmul_rn(__fdiv_rn(__fdiv_rn(Y, den), R), 2.0f), 1.0f);
  const float half = R * 0.5f;
  ix = fmaf(__fadd_rn(gx, 1.0f), half, -0.5f);
  iy = fmaf(__fadd_rn(gy, 1.0f), half, -0.5f);
}

// resize source index / weights for output index d  (align_corners=True)
__device__ __forceinline__ void resize_coef(const Crop& c, int d, int R, int& i0, int& i1, float& l0, float& l1) {
  if (c.s == R) { i0 = i1 = d; l0 = 1.0f; l1 = 0.0f; return; }
  const float src = __fmul_rn(c.scale, (float)d);
  i0 = min((int)src, c.s - 1);
  i1 = i0 + (i0 < c.s - 1 ? 1 : 0);
  l1 = fminf(fmaxf(__fsub_rn(src, (float)i0), 0.0f), 1.0f);
  l0 = __fsub_rn(1.0f, l1);
}

// bilinear gather of one channel plane with zero padding; accumulation order of torch's CPU grid_sample
__device__ __forceinline__ float gather_bilinear(const float* __restrict__ plane, int R, float ix, float iy) {
  if (!(ix > -1.0f && ix < (float)R && iy > -1.0f && iy < (float)R)) return 0.0f;
  const float fx = floorf(ix), fy = floorf(iy);
  const int x0 = (int)fx, y0 = (int)fy;
  const float wx1 = __fsub_rn(ix, fx), wx0 = __fsub_rn(__fadd_rn(fx, 1.0f), ix);
  const float wy1 = __fsub_rn(iy, fy), wy0 = __fsub_rn(__fadd_rn(fy, 1.0f), iy);
  const bool xa = x0 >= 0, xb = x0 + 1 < R, ya = y0 >= 0, yb = y0 + 1 < R;
  const float nw = (xa && ya) ? __ldg(plane + (size_t)y0 * R + x0) : 0.0f;
  const float ne = (xb && ya) ? __ldg(plane + (size_t)y0 * R + x0 + 1) : 0.0f;
  const float sw = (xa && yb) ? __ldg(plane + (size_t)(y0 + 1) * R + x0) : 0.0f;
  const float se = (xb && yb) ? __ldg(plane + (size_t)(y0 + 1) * R + x0 + 1) : 0.0f;
  float acc = __fmul_rn(nw, __fmul_rn(wx0, wy0));
  acc = fmaf(ne, __fmul_rn(wx1, wy0), acc);
  acc = fmaf(sw, __fmul_rn(wx0, wy1), acc);
  acc = fmaf(se, __fmul_rn(wx1, wy1), acc);
  return acc;
}

// ---- forward -------------------------------------------------------------------------------------
constexpr int PCL_TR = 16;       // output rows per CTA
constexpr int PCL_THREADS = 256;

__global__ void __launch_bounds__(PCL_THREADS) pcl_fwd_kernel(const float* __restrict__ img, const float* __restrict__ params,
                                                              int crops_per_img, int C, int R, float* __restrict__ out, int max_rows) {
  extern __shared__ __align__(16) float mid[];  // [C][max_rows][s]
  const int q = blockIdx.y;
  const Crop c = load_crop(params + (size_t)q * PF);
  const int s = c.s;
  const int y0 = blockIdx.x * PCL_TR;
  const int y1 = min(y0 + PCL_TR, R) - 1;
  int jlo, jhi, t0, t1;
  float l0, l1;
  resize_coef(c, y0, R, jlo, t1, l0, l1);
  resize_coef(c, y1, R, t0, jhi, l0, l1);
  const int nrows = jhi - jlo + 1;
  const float* src = img + (size_t)(q / crops_per_img) * C * R * R;
  float* dst = out + (size_t)q * C * R * R;
  const float Rf = (float)R;
  const bool staged = nrows <= max_rows && s <= R;
  if (staged) {
    const int n = nrows * s;
    for (int idx = threadIdx.x; idx < n; idx += PCL_THREADS) {
      const int jr = idx / s, i = idx - jr * s;
      float ix, iy;
      sample_pos(c, jlo + jr, i, Rf, ix, iy);
      for (int ch = 0; ch < C; ++ch) mid[(ch * max_rows + jr) * s + i] = gather_bilinear(src + (size_t)ch * R * R, R, ix, iy);
    }
    __syncthreads();
  }
  const int npix = (y1 - y0 + 1) * R;
  for (int idx = threadIdx.x; idx < npix; idx += PCL_THREADS) {
    const int yy = idx / R, x = idx - yy * R;
    const int y = y0 + yy;
    int a0, a1, b0, b1;
    float ly0, ly1, lx0, lx1;
    resize_coef(c, y, R, a0, a1, ly0, ly1);
    resize_coef(c, x, R, b0, b1, lx0, lx1);
    const float w00 = __fmul_rn(ly0, lx0), w01 = __fmul_rn(ly0, lx1), w10 = __fmul_rn(ly1, lx0), w11 = __fmul_rn(ly1, lx1);
    if (staged) {
      for (int ch = 0; ch < C; ++ch) {
        const float* m0 = mid + (ch * max_rows + (a0 - jlo)) * s;
        const float* m1 = mid + (ch * max_rows + (a1 - jlo)) * s;
        float acc = __fmul_rn(w01, m0[b1]);
        acc = fmaf(w00, m0[b0], acc);
        acc = fmaf(w10, m1[b0], acc);
        acc = fmaf(w11, m1[b1], acc);
        dst[((size_t)ch * R + y) * R + x] = acc;
      }
    } else {
      float px[4], py[4];
      sample_pos(c, a0, b0, Rf, px[0], py[0]);
      sample_pos(c, a0, b1, Rf, px[1], py[1]);
      sample_pos(c, a1, b0, Rf, px[2], py[2]);
      sample_pos(c, a1, b1, Rf, px[3], py[3]);
      for (int ch = 0; ch < C; ++ch) {
        const float* pl = src + (size_t)ch * R * R;
        float acc = __fmul_rn(w01, gather_bilinear(pl, R, px[1], py[1]));
        acc = fmaf(w00, gather_bilinear(pl, R, px[0], py[0]), acc);
        acc = fmaf(w10, gather_bilinear(pl, R, px[2], py[2]), acc);
        acc = fmaf(w11, gather_bilinear(pl, R, px[3], py[3]), acc);
        dst[((size_t)ch * R + y) * R + x] = acc;
      }
    }
  }
}

// ---- backward -------------------------------------------------------------------------------------
// per-chunk exclusive scan of C*s*s -> params[21]; one block per chunk, thread 0 walks its crops.
__global__ void pcl_offsets_kernel(float* __restrict__ params, int n_crops, int chunk_crops, int C) {
  if (threadIdx.x != 0) return;
  const int q0 = blockIdx.x * chunk_crops;
  const int q1 = min(q0 + chunk_crops, n_crops);
  int off = 0;
  for (int q = q0; q < q1; ++q) {
    float* rec = params + (size_t)q * PF;
    const int s = __float_as_int(rec[18]);
    rec[21] = __int_as_float(off);
    off += C * s * s;
  }
}

// contributions of output index range to intermediate index j along one axis
__device__ __forceinline__ void out_range(const Crop& c, int j, int R, int& lo, int& hi) {
  if (c.s == R) { lo = hi = j; return; }
  if (c.scale <= 0.0f) { lo = 0; hi = R - 1; return; }
  const float inv = 1.0f / c.scale;
  lo = max(0, (int)floorf((float)(j - 1) * inv) - 1);
  hi = min(R - 1, (int)ceilf((float)(j + 1) * inv) + 1);
}
__device__ __forceinline__ float axis_weight(const Crop& c, int d, int j, int R) {
  int i0, i1;
  float l0, l1;
  resize_coef(c, d, R, i0, i1, l0, l1);
  return (i0 == j ? l0 : 0.0f) + (i1 == j ? l1 : 0.0f);
}

constexpr int PCL_JR = 8;  // intermediate rows per CTA pass

__global__ void __launch_bounds__(PCL_THREADS) pcl_bwd_mid_kernel(const float* __restrict__ g_out, const float* __restrict__ params,
                                                                  int q_base, int C, int R, float* __restrict__ gmid) {
  const int q = q_base + blockIdx.y;
  const float* rec = params + (size_t)q * PF;
  const Crop c = load_crop(rec);
  const int s = c.s;
  float* gm = gmid + __float_as_int(__ldg(rec + 21));
  const float* go = g_out + (size_t)q * C * R * R;
  for (int jb = blockIdx.x * PCL_JR; jb < s; jb += gridDim.x * PCL_JR) {
    const int n = min(PCL_JR, s - jb) * s;
    for (int idx = threadIdx.x; idx < n; idx += PCL_THREADS) {
      const int jr = idx / s, i = idx - jr * s;
      const int j = jb + jr;
      int ylo, yhi, xlo, xhi;
      out_range(c, j, R, ylo, yhi);
      out_range(c, i, R, xlo, xhi);
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      for (int y = ylo; y <= yhi; ++y) {
        const float wy = axis_weight(c, y, j, R);
        if (wy == 0.0f) continue;
        for (int x = xlo; x <= xhi; ++x) {
          // forward weight of this tap is fl(ly*lx) for each (i0/i1) combination; collapse the <=2x2 combos
          int a0, a1, b0, b1;
          float ly0, ly1, lx0, lx1;
          resize_coef(c, y, R, a0, a1, ly0, ly1);
          resize_coef(c, x, R, b0, b1, lx0, lx1);
          float w = 0.0f;
          if (a0 == j && b0 == i) w += __fmul_rn(ly0, lx0);
          if (a0 == j && b1 == i) w += __fmul_rn(ly0, lx1);
          if (a1 == j && b0 == i) w += __fmul_rn(ly1, lx0);
          if (a1 == j && b1 == i) w += __fmul_rn(ly1, lx1);
          if (w == 0.0f) continue;
          for (int ch = 0; ch < C; ++ch) acc[ch] = fmaf(w, __ldg(go + ((size_t)ch * R + y) * R + x), acc[ch]);
        }
      }
      for (int ch = 0; ch < C; ++ch) gm[((size_t)ch * s + j) * s + i] = acc[ch];
    }
  }
}

constexpr int PCL_TS = 32;              // source tile side
constexpr int PCL_HALO = PCL_TS + 2;

__global__ void __launch_bounds__(PCL_THREADS) pcl_bwd_img_kernel(const float* __restrict__ params, const float* __restrict__ gmid,
                                                                  int img_base, int crops_per_img, int C, int R,
                                                                  float* __restrict__ g_img) {
  __shared__ float pm[PCL_HALO * PCL_HALO][2];
  const int tiles_x = (R + PCL_TS - 1) / PCL_TS;
  const int tx0 = (blockIdx.x % tiles_x) * PCL_TS, ty0 = (blockIdx.x / tiles_x) * PCL_TS;
  const int im = img_base + blockIdx.y;
  const int lx = threadIdx.x & 31, lyb = threadIdx.x >> 5;  // 32 x 8 threads, 4 rows each
  float acc[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) acc[r][ch] = 0.f;
  const float Rf = (float)R;
  for (int k = 0; k < crops_per_img; ++k) {
    const int q = im * crops_per_img + k;
    const float* rec = params + (size_t)q * PF;
    const Crop c = load_crop(rec);
    const int s = c.s;
    float Pi[9];
#pragma unroll
    for (int e = 0; e < 9; ++e) Pi[e] = __ldg(rec + 9 + e);
    const float* gm = gmid + __float_as_int(__ldg(rec + 21));
    const float sm1 = (float)(s - 1);
    __syncthreads();  // previous crop's readers are done with pm
    for (int idx = threadIdx.x; idx < PCL_HALO * PCL_HALO; idx += PCL_THREADS) {
      const int hy = idx / PCL_HALO, hx = idx - hy * PCL_HALO;
      // sample position equal to pixel index (px,py)  <=>  grid-sample pixel coordinate px + 0.5
      const float gx = (float)(tx0 - 1 + hx) + 0.5f, gy = (float)(ty0 - 1 + hy) + 0.5f;
      const float U = Pi[0] * gx + Pi[1] * gy + Pi[2];
      const float V = Pi[3] * gx + Pi[4] * gy + Pi[5];
      const float Wd = Pi[6] * gx + Pi[7] * gy + Pi[8];
      float mi = nanf(""), mj = nanf("");
      if (Wd > 1e-12f) { mi = U / Wd * sm1; mj = V / Wd * sm1; }
      pm[idx][0] = mi; pm[idx][1] = mj;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int ly = lyb + 8 * r;
      const int sx = tx0 + lx, sy = ty0 + ly;
      if (sx >= R || sy >= R) continue;
      // candidate box from the pre-images of the four diagonal neighbours
      float ilo = 3.0e38f, ihi = -3.0e38f, jlo = 3.0e38f, jhi = -3.0e38f;
      bool bad = false;
#pragma unroll
      for (int dy = 0; dy <= 2; dy += 2)
#pragma unroll
        for (int dx = 0; dx <= 2; dx += 2) {
          const float* p = pm[(ly + dy) * PCL_HALO + (lx + dx)];
          bad |= !(p[0] == p[0]);
          ilo = fminf(ilo, p[0]); ihi = fmaxf(ihi, p[0]); jlo = fminf(jlo, p[1]); jhi = fmaxf(jhi, p[1]);
        }
      int i0, i1, j0, j1;
      if (bad) { i0 = 0; i1 = s - 1; j0 = 0; j1 = s - 1; }
      else {
        ilo = fmaxf(ilo - 0.05f, -1.0f); jlo = fmaxf(jlo - 0.05f, -1.0f);
        ihi = fminf(ihi + 0.05f, sm1 + 1.0f); jhi = fminf(jhi + 0.05f, sm1 + 1.0f);
        i0 = max(0, (int)floorf(ilo)); i1 = min(s - 1, (int)ceilf(ihi));
        j0 = max(0, (int)floorf(jlo)); j1 = min(s - 1, (int)ceilf(jhi));
      }
      const float fsx = (float)sx, fsy = (float)sy;
      for (int j = j0; j <= j1; ++j)
        for (int i = i0; i <= i1; ++i) {
          float ix, iy;
          sample_pos(c, j, i, Rf, ix, iy);
          if (!(ix > -1.0f && ix < Rf && iy > -1.0f && iy < Rf)) continue;
          const float fx = floorf(ix), fy = floorf(iy);
          float wx, wy;
          if (fsx == fx) wx = __fsub_rn(__fadd_rn(fx, 1.0f), ix); else if (fsx == fx + 1.0f) wx = __fsub_rn(ix, fx); else continue;
          if (fsy == fy) wy = __fsub_rn(__fadd_rn(fy, 1.0f), iy); else if (fsy == fy + 1.0f) wy = __fsub_rn(iy, fy); else continue;
          const float w = __fmul_rn(wx, wy);
          for (int ch = 0; ch < C; ++ch) acc[r][ch] = fmaf(w, __ldg(gm + ((size_t)ch * s + j) * s + i), acc[r][ch]);
        }
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int sx = tx0 + lx, sy = ty0 + lyb + 8 * r;
    if (sx >= R || sy >= R) continue;
    for (int ch = 0; ch < C; ++ch) g_img[(((size_t)im * C + ch) * R + sy) * R + sx] = acc[r][ch];
  }
}

}  // namespace hb

using namespace hb;

extern "C" int hb_pcl_fwd(const float* img, const float* params, int n_crops, int crops_per_img, int C, int img_res, float* out, void* stream) {
  if (n_crops < 0 || crops_per_img <= 0 || C <= 0 || img_res <= 0 || (n_crops > 0 && (!img || !params || !out)) || n_crops % crops_per_img) {
    set_error("hb_pcl_fwd: bad argument"); return HB_E_ARG;
  }
  if (n_crops == 0) return 0;
  const int max_rows = PCL_TR + 2;
  const size_t smem = sizeof(float) * (size_t)C * max_rows * img_res;
  if (smem > 200 * 1024) { set_error("hb_pcl_fwd: C*img_res too large for the staged kernel"); return HB_E_UNSUPPORTED; }
  HB_CUDA(cudaFuncSetAttribute(pcl_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((img_res + PCL_TR - 1) / PCL_TR, n_crops);
  pcl_fwd_kernel<<<grid, PCL_THREADS, smem, (cudaStream_t)stream>>>(img, params, crops_per_img, C, img_res, out, max_rows);
  g_launches++;
  return check_launch("pcl_fwd_kernel");
}

static const int kPclChunkImgs = 128;

extern "C" size_t hb_pcl_bwd_workspace_bytes(int n_crops, int crops_per_img, int C, int img_res) {
  if (n_crops <= 0 || crops_per_img <= 0) return 0;
  const int n_imgs = n_crops / crops_per_img;
  const int chunk_crops = (n_imgs < kPclChunkImgs ? n_imgs : kPclChunkImgs) * crops_per_img;
  return sizeof(float) * (size_t)chunk_crops * C * img_res * img_res;
}

extern "C" int hb_pcl_bwd(const float* g_out, const float* params, int n_crops, int crops_per_img, int C, int img_res, float* g_img,
                          void* workspace, size_t workspace_bytes, void* stream) {
  if (n_crops < 0 || crops_per_img <= 0 || C <= 0 || C > 4 || img_res <= 0 || (n_crops > 0 && (!g_out || !params || !g_img || !workspace)) ||
      n_crops % crops_per_img) {
    set_error("hb_pcl_bwd: bad argument (C must be <= 4)"); return HB_E_ARG;
  }
  if (n_crops == 0) return 0;
  if (workspace_bytes < hb_pcl_bwd_workspace_bytes(n_crops, crops_per_img, C, img_res)) { set_error("hb_pcl_bwd: workspace too small"); return HB_E_WORKSPACE; }
  cudaStream_t st = (cudaStream_t)stream;
  const int n_imgs = n_crops / crops_per_img;
  const int chunk_imgs = n_imgs < kPclChunkImgs ? n_imgs : kPclChunkImgs;
  const int chunk_crops = chunk_imgs * crops_per_img;
  const int n_chunks = (n_imgs + chunk_imgs - 1) / chunk_imgs;
  pcl_offsets_kernel<<<n_chunks, 32, 0, st>>>(const_cast<float*>(params), n_crops, chunk_crops, C);
  g_launches++;
  int rc = check_launch("pcl_offsets_kernel");
  if (rc) return rc;
  const int tiles = ((img_res + PCL_TS - 1) / PCL_TS) * ((img_res + PCL_TS - 1) / PCL_TS);
  for (int ch = 0; ch < n_chunks; ++ch) {
    const int im0 = ch * chunk_imgs;
    const int nim = (n_imgs - im0) < chunk_imgs ? (n_imgs - im0) : chunk_imgs;
    dim3 g1((img_res + PCL_JR - 1) / PCL_JR, nim * crops_per_img);
    pcl_bwd_mid_kernel<<<g1, PCL_THREADS, 0, st>>>(g_out, params, im0 * crops_per_img, C, img_res, (float*)workspace);
    g_launches++;
    rc = check_launch("pcl_bwd_mid_kernel");
    if (rc) return rc;
    dim3 g2(tiles, nim);
    pcl_bwd_img_kernel<<<g2, PCL_THREADS, 0, st>>>(params, (const float*)workspace, im0, crops_per_img, C, img_res, g_img);
    g_launches++;
    rc = check_launch("pcl_bwd_img_kernel");
    if (rc) return rc;
  }
  return 0;
}
