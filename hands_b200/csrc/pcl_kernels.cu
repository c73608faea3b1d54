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
#include <climits>
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

__device__ __forceinline__ int fast_div(int idx, int s, float inv_s) {
  int q = __float2int_rz(((float)idx + 0.5f) * inv_s);
  int r = idx - q * s;
  if (r < 0) --q; else if (r >= s) ++q;
  return q;
}

// One CTA per (crop, PCL_TR output rows).
//   phase 1: the rows of the s x s intermediate this tile needs are gathered once (4 taps x C) into smem,
//            channel-interleaved so one LDS.128 fetches a pixel;
//   phase 2: thread = output column, walking down the tile's rows; the two intermediate rows in use stay in
//            registers while consecutive output rows share them (up-sampling), stores are coalesced rows.
template <int C>
__global__ void __launch_bounds__(PCL_THREADS) pcl_fwd_kernel(const float* __restrict__ img, const float* __restrict__ params,
                                                              int crops_per_img, int R, float* __restrict__ out, int max_rows) {
  extern __shared__ __align__(16) float4 mid4[];  // [max_rows][s] pixels, channels in .x .y .z .w
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
    const float inv_s = 1.0f / (float)s;
    for (int idx = threadIdx.x; idx < n; idx += PCL_THREADS) {
      const int jr = fast_div(idx, s, inv_s), i = idx - jr * s;
      float ix, iy;
      sample_pos(c, jlo + jr, i, Rf, ix, iy);
      float v[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int ch = 0; ch < C; ++ch) v[ch] = gather_bilinear(src + (size_t)ch * R * R, R, ix, iy);
      mid4[idx] = make_float4(v[0], v[1], v[2], v[3]);
    }
    __syncthreads();
    for (int x = threadIdx.x; x < R; x += PCL_THREADS) {
      int b0, b1;
      float lx0, lx1;
      resize_coef(c, x, R, b0, b1, lx0, lx1);
      int ra = -1, rb = -1;  // intermediate rows currently held in (m00,m01) and (m10,m11)
      float4 m00 = make_float4(0, 0, 0, 0), m01 = m00, m10 = m00, m11 = m00;
      for (int y = y0; y <= y1; ++y) {
        int a0, a1;
        float ly0, ly1;
        resize_coef(c, y, R, a0, a1, ly0, ly1);
        if (a0 != ra) {
          if (a0 == rb) { m00 = m10; m01 = m11; }
          else { const float4* row = mid4 + (a0 - jlo) * s; m00 = row[b0]; m01 = row[b1]; }
          ra = a0;
        }
        if (a1 != rb) {
          if (a1 == ra) { m10 = m00; m11 = m01; }
          else { const float4* row = mid4 + (a1 - jlo) * s; m10 = row[b0]; m11 = row[b1]; }
          rb = a1;
        }
        const float w00 = __fmul_rn(ly0, lx0), w01 = __fmul_rn(ly0, lx1), w10 = __fmul_rn(ly1, lx0), w11 = __fmul_rn(ly1, lx1);
        const float a[4] = {m00.x, m00.y, m00.z, m00.w}, b[4] = {m01.x, m01.y, m01.z, m01.w};
        const float d[4] = {m10.x, m10.y, m10.z, m10.w}, e[4] = {m11.x, m11.y, m11.z, m11.w};
#pragma unroll
        for (int ch = 0; ch < C; ++ch) {
          float acc = __fmul_rn(w01, b[ch]);
          acc = fmaf(w00, a[ch], acc);
          acc = fmaf(w10, d[ch], acc);
          acc = fmaf(w11, e[ch], acc);
          __stcs(dst + ((size_t)ch * R + y) * R + x, acc);
        }
      }
    }
    return;
  }
  // generic path (s > R or the tile needs more intermediate rows than fit): evaluate the four intermediate
  // pixels of every output pixel directly
  const int npix = (y1 - y0 + 1) * R;
  for (int idx = threadIdx.x; idx < npix; idx += PCL_THREADS) {
    const int yy = idx / R, x = idx - yy * R;
    const int y = y0 + yy;
    int a0, a1, b0, b1;
    float ly0, ly1, lx0, lx1;
    resize_coef(c, y, R, a0, a1, ly0, ly1);
    resize_coef(c, x, R, b0, b1, lx0, lx1);
    const float w00 = __fmul_rn(ly0, lx0), w01 = __fmul_rn(ly0, lx1), w10 = __fmul_rn(ly1, lx0), w11 = __fmul_rn(ly1, lx1);
    float px[4], py[4];
    sample_pos(c, a0, b0, Rf, px[0], py[0]);
    sample_pos(c, a0, b1, Rf, px[1], py[1]);
    sample_pos(c, a1, b0, Rf, px[2], py[2]);
    sample_pos(c, a1, b1, Rf, px[3], py[3]);
#pragma unroll
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

// ---- backward -------------------------------------------------------------------------------------
// Chunk workspace, per crop (offset params[21], in floats, a multiple of 4): G float4[s*s] intermediate
// gradient (channels in x,y,z,w), then POS float2[s*s] sample positions.  6*s*s floats per crop.
constexpr int PCL_WS_FLOATS_PER_PX = 6;

__global__ void pcl_offsets_kernel(float* __restrict__ params, int n_crops, int chunk_crops) {
  if (threadIdx.x != 0) return;
  const int q0 = blockIdx.x * chunk_crops;
  const int q1 = min(q0 + chunk_crops, n_crops);
  int off = 0;
  for (int q = q0; q < q1; ++q) {
    float* rec = params + (size_t)q * PF;
    const int s = __float_as_int(rec[18]);
    rec[21] = __int_as_float(off);
    off += (PCL_WS_FLOATS_PER_PX * s * s + 3) & ~3;  // keep every crop's float4 array 16-byte aligned
  }
}

constexpr int PCL_JR = 16;  // intermediate rows per CTA

// Transposed resize (g_out -> intermediate gradient), separable, gather form.
//   tables: for every output index d its source index i0[d] and weight l1[d]; for every intermediate index m
//           the contiguous range [dlo[m], dhi[m]] of output indices that touch it (same table for both axes).
//   pass A: H[y][i] = sum_x wx(x,i) g_out[y][x]   for the output rows this band of intermediate rows needs
//   pass B: G[j][i] = sum_y wy(y,j) H[y][i]       plus the sample position of (j,i) for the next kernel
template <int C>
__global__ void __launch_bounds__(PCL_THREADS) pcl_bwd_mid_kernel(const float* __restrict__ g_out, const float* __restrict__ params,
                                                                  int q_base, int R, float* __restrict__ ws, int cap) {
  extern __shared__ __align__(16) float sm[];
  const int q = q_base + blockIdx.y;
  const float* rec = params + (size_t)q * PF;
  const Crop c = load_crop(rec);
  const int s = c.s;
  const int j0 = blockIdx.x * PCL_JR;
  if (j0 >= s) return;
  const int j1 = min(j0 + PCL_JR, s) - 1;
  int* ti0 = reinterpret_cast<int*>(sm);     // [R]
  float* tl1 = sm + R;                       // [R]
  int* dlo = reinterpret_cast<int*>(sm + 2 * R);  // [R]  (s <= R on this path)
  int* dhi = dlo + R;                        // [R]
  float* H = sm + 4 * R;                     // [C][nrows][s]
  float* base = ws + __float_as_int(__ldg(rec + 21));
  float4* G = reinterpret_cast<float4*>(base);
  float2* POS = reinterpret_cast<float2*>(base + 4 * (size_t)s * s);
  const float* go = g_out + (size_t)q * C * R * R;
  const float Rf = (float)R;
  const float inv_s = 1.0f / (float)s;
  const bool fast = s <= R;
  if (fast) {
    for (int m = threadIdx.x; m < s; m += PCL_THREADS) { dlo[m] = R; dhi[m] = -1; }
    __syncthreads();
    for (int d = threadIdx.x; d < R; d += PCL_THREADS) {
      int i0, i1;
      float l0, l1;
      resize_coef(c, d, R, i0, i1, l0, l1);
      ti0[d] = i0; tl1[d] = l1;
      atomicMin(dlo + i0, d); atomicMax(dhi + i0, d);
      atomicMin(dlo + i1, d); atomicMax(dhi + i1, d);
    }
    __syncthreads();
  }
  int ylo = 0, yhi = -1;
  if (fast) { ylo = dlo[j0]; yhi = dhi[j1]; for (int j = j0; j <= j1; ++j) { ylo = min(ylo, dlo[j]); yhi = max(yhi, dhi[j]); } }
  const int nrows = yhi - ylo + 1;
  if (fast && nrows > 0 && nrows * s <= cap) {
    // pass A
    const int nA = nrows * s;
    for (int idx = threadIdx.x; idx < nA; idx += PCL_THREADS) {
      const int yr = fast_div(idx, s, inv_s), i = idx - yr * s;
      const int y = ylo + yr;
      float acc[C];
#pragma unroll
      for (int ch = 0; ch < C; ++ch) acc[ch] = 0.f;
      const float* row = go + (size_t)y * R;
      const int xl = dlo[i], xh = dhi[i];
      for (int x = xl; x <= xh; ++x) {
        const int i0 = ti0[x];
        const float l1 = tl1[x];
        const int i1 = i0 + (i0 < s - 1 ? 1 : 0);
        const float w = (i0 == i ? __fsub_rn(1.0f, l1) : 0.0f) + (i1 == i ? l1 : 0.0f);
#pragma unroll
        for (int ch = 0; ch < C; ++ch) acc[ch] = fmaf(w, __ldcs(row + (size_t)ch * R * R + x), acc[ch]);
      }
#pragma unroll
      for (int ch = 0; ch < C; ++ch) H[(ch * nrows + yr) * s + i] = acc[ch];
    }
    __syncthreads();
    // pass B
    const int nB = (j1 - j0 + 1) * s;
    for (int idx = threadIdx.x; idx < nB; idx += PCL_THREADS) {
      const int jr = fast_div(idx, s, inv_s), i = idx - jr * s;
      const int j = j0 + jr;
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      const int yl = dlo[j], yh = dhi[j];
      for (int y = yl; y <= yh; ++y) {
        const int i0 = ti0[y];
        const float l1 = tl1[y];
        const int i1 = i0 + (i0 < s - 1 ? 1 : 0);
        const float w = (i0 == j ? __fsub_rn(1.0f, l1) : 0.0f) + (i1 == j ? l1 : 0.0f);
#pragma unroll
        for (int ch = 0; ch < C; ++ch) acc[ch] = fmaf(w, H[(ch * nrows + (y - ylo)) * s + i], acc[ch]);
      }
      float ix, iy;
      sample_pos(c, j, i, Rf, ix, iy);
      POS[(size_t)j * s + i] = make_float2(ix, iy);
      G[(size_t)j * s + i] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    }
    return;
  }
  // generic path: direct 2-D gather per intermediate pixel
  const int nB = (j1 - j0 + 1) * s;
  for (int idx = threadIdx.x; idx < nB; idx += PCL_THREADS) {
    const int jr = idx / s, i = idx - jr * s;
    const int j = j0 + jr;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int y = 0; y < R; ++y) {
      int a0, a1;
      float ly0, ly1;
      resize_coef(c, y, R, a0, a1, ly0, ly1);
      const float wy = (a0 == j ? ly0 : 0.0f) + (a1 == j ? ly1 : 0.0f);
      if (wy == 0.0f) continue;
      for (int x = 0; x < R; ++x) {
        int b0, b1;
        float lx0, lx1;
        resize_coef(c, x, R, b0, b1, lx0, lx1);
        const float wx = (b0 == i ? lx0 : 0.0f) + (b1 == i ? lx1 : 0.0f);
        if (wx == 0.0f) continue;
#pragma unroll
        for (int ch = 0; ch < C; ++ch) acc[ch] = fmaf(wy * wx, __ldg(go + ((size_t)ch * R + y) * R + x), acc[ch]);
      }
    }
    float ix, iy;
    sample_pos(c, j, i, Rf, ix, iy);
    POS[(size_t)j * s + i] = make_float2(ix, iy);
    G[(size_t)j * s + i] = make_float4(acc[0], acc[1], acc[2], acc[3]);
  }
}

constexpr int PCL_TS = 32;              // source tile side
constexpr int PCL_HALO = PCL_TS + 2;
constexpr int PCL_REGCAP = 1600;        // intermediate pixels staged in smem per (tile, crop)

// Transposed grid_sample, gather form.  One CTA per (image, 32x32 source tile); for each crop of the image:
// the inverse homography maps the tile (plus a one-pixel halo) into the intermediate grid; that region's
// sample positions and gradients are staged in shared memory; every source pixel then collects from the
// intermediate pixels inside the pre-image of its +-1 neighbourhood whose bilinear footprint covers it.
template <int C>
__global__ void __launch_bounds__(PCL_THREADS) pcl_bwd_img_kernel(const float* __restrict__ params, const float* __restrict__ ws,
                                                                  int img_base, int crops_per_img, int R, float* __restrict__ g_img) {
  __shared__ float2 pm[PCL_HALO * PCL_HALO];
  __shared__ __align__(16) float4 gS[PCL_REGCAP];
  __shared__ float2 pS[PCL_REGCAP];
  __shared__ int box[4];  // i_min, i_max, j_min, j_max of the tile's pre-image (ints after floor/ceil)
  __shared__ int anybad;
  const int tiles_x = (R + PCL_TS - 1) / PCL_TS;
  const int tx0 = (blockIdx.x % tiles_x) * PCL_TS, ty0 = (blockIdx.x / tiles_x) * PCL_TS;
  const int im = img_base + blockIdx.y;
  const int lx = threadIdx.x & 31, lyb = threadIdx.x >> 5;  // 32 x 8 threads, 4 rows each
  float acc[4][C];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int ch = 0; ch < C; ++ch) acc[r][ch] = 0.f;
  for (int k = 0; k < crops_per_img; ++k) {
    const int q = im * crops_per_img + k;
    const float* rec = params + (size_t)q * PF;
    const int s = __float_as_int(__ldg(rec + 18));
    float Pi[9];
#pragma unroll
    for (int e = 0; e < 9; ++e) Pi[e] = __ldg(rec + 9 + e);
    const float* base = ws + __float_as_int(__ldg(rec + 21));
    const float4* G = reinterpret_cast<const float4*>(base);
    const float2* POS = reinterpret_cast<const float2*>(base + 4 * (size_t)s * s);
    const float sm1 = (float)(s - 1);
    __syncthreads();  // previous crop's readers are done with the shared tiles
    if (threadIdx.x == 0) { box[0] = INT_MAX; box[1] = INT_MIN; box[2] = INT_MAX; box[3] = INT_MIN; anybad = 0; }
    __syncthreads();
    {
      int imin = INT_MAX, imax = INT_MIN, jmin = INT_MAX, jmax = INT_MIN, bad = 0;
      for (int idx = threadIdx.x; idx < PCL_HALO * PCL_HALO; idx += PCL_THREADS) {
        const int hy = idx / PCL_HALO, hx = idx - hy * PCL_HALO;
        // sample position equal to pixel index (px,py)  <=>  grid-sample pixel coordinate px + 0.5
        const float gx = (float)(tx0 - 1 + hx) + 0.5f, gy = (float)(ty0 - 1 + hy) + 0.5f;
        const float U = Pi[0] * gx + Pi[1] * gy + Pi[2];
        const float V = Pi[3] * gx + Pi[4] * gy + Pi[5];
        const float Wd = Pi[6] * gx + Pi[7] * gy + Pi[8];
        float mi, mj;
        if (Wd > 1e-12f) {
          const float iw = 1.0f / Wd;
          mi = fminf(fmaxf(U * iw * sm1, -4.0f), sm1 + 4.0f);
          mj = fminf(fmaxf(V * iw * sm1, -4.0f), sm1 + 4.0f);
          imin = min(imin, (int)floorf(mi - 0.05f)); imax = max(imax, (int)ceilf(mi + 0.05f));
          jmin = min(jmin, (int)floorf(mj - 0.05f)); jmax = max(jmax, (int)ceilf(mj + 0.05f));
        } else { mi = nanf(""); mj = nanf(""); bad = 1; }
        pm[idx] = make_float2(mi, mj);
      }
#pragma unroll
      for (int m = 16; m > 0; m >>= 1) {
        imin = min(imin, __shfl_xor_sync(0xffffffffu, imin, m)); imax = max(imax, __shfl_xor_sync(0xffffffffu, imax, m));
        jmin = min(jmin, __shfl_xor_sync(0xffffffffu, jmin, m)); jmax = max(jmax, __shfl_xor_sync(0xffffffffu, jmax, m));
        bad |= __shfl_xor_sync(0xffffffffu, bad, m);
      }
      if (lx == 0) {
        atomicMin(&box[0], imin); atomicMax(&box[1], imax); atomicMin(&box[2], jmin); atomicMax(&box[3], jmax);
        if (bad) atomicOr(&anybad, 1);
      }
    }
    __syncthreads();
    int ri0, ri1, rj0, rj1;
    if (anybad) { ri0 = 0; ri1 = s - 1; rj0 = 0; rj1 = s - 1; }
    else { ri0 = max(box[0], 0); ri1 = min(box[1], s - 1); rj0 = max(box[2], 0); rj1 = min(box[3], s - 1); }
    if (ri0 > ri1 || rj0 > rj1) continue;  // this crop does not touch the tile (block-uniform)
    const int rw = ri1 - ri0 + 1, rh = rj1 - rj0 + 1;
    const bool in_smem = rw * rh <= PCL_REGCAP;
    if (in_smem) {
      const float inv_rw = 1.0f / (float)rw;
      for (int idx = threadIdx.x; idx < rw * rh; idx += PCL_THREADS) {
        const int rr = fast_div(idx, rw, inv_rw), cc = idx - rr * rw;
        const size_t gidx = (size_t)(rj0 + rr) * s + (ri0 + cc);
        pS[idx] = __ldg(POS + gidx);
        gS[idx] = __ldg(G + gidx);
      }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int ly = lyb + 8 * r;
      const int sx = tx0 + lx, sy = ty0 + ly;
      if (sx >= R || sy >= R) continue;
      float ilo = 3.0e38f, ihi = -3.0e38f, jlo = 3.0e38f, jhi = -3.0e38f;
      bool bad = false;
#pragma unroll
      for (int dy = 0; dy <= 2; dy += 2)
#pragma unroll
        for (int dx = 0; dx <= 2; dx += 2) {
          const float2 p = pm[(ly + dy) * PCL_HALO + (lx + dx)];
          bad |= !(p.x == p.x);
          ilo = fminf(ilo, p.x); ihi = fmaxf(ihi, p.x); jlo = fminf(jlo, p.y); jhi = fmaxf(jhi, p.y);
        }
      int i0, i1, j0, j1;
      if (bad) { i0 = ri0; i1 = ri1; j0 = rj0; j1 = rj1; }
      else {
        i0 = max(ri0, (int)floorf(ilo - 0.05f)); i1 = min(ri1, (int)ceilf(ihi + 0.05f));
        j0 = max(rj0, (int)floorf(jlo - 0.05f)); j1 = min(rj1, (int)ceilf(jhi + 0.05f));
      }
      const float fsx = (float)sx, fsy = (float)sy;
      for (int j = j0; j <= j1; ++j)
        for (int i = i0; i <= i1; ++i) {
          const int lidx = (j - rj0) * rw + (i - ri0);
          const float2 p = in_smem ? pS[lidx] : __ldg(POS + (size_t)j * s + i);
          const float dx = p.x - fsx, dy = p.y - fsy;   // exact (pixel indices are small integers)
          if (!(dx > -1.0f && dx < 1.0f && dy > -1.0f && dy < 1.0f)) continue;
          // forward weights: tap at floor(p): (floor+1) - p ; tap at floor(p)+1: p - floor
          const float wx = dx >= 0.0f ? __fsub_rn(__fadd_rn(fsx, 1.0f), p.x) : __fsub_rn(p.x, __fsub_rn(fsx, 1.0f));
          const float wy = dy >= 0.0f ? __fsub_rn(__fadd_rn(fsy, 1.0f), p.y) : __fsub_rn(p.y, __fsub_rn(fsy, 1.0f));
          const float w = __fmul_rn(wx, wy);
          const float4 g = in_smem ? gS[lidx] : __ldg(G + (size_t)j * s + i);
          const float gv[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
          for (int ch = 0; ch < C; ++ch) acc[r][ch] = fmaf(w, gv[ch], acc[r][ch]);
        }
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int sx = tx0 + lx, sy = ty0 + lyb + 8 * r;
    if (sx >= R || sy >= R) continue;
#pragma unroll
    for (int ch = 0; ch < C; ++ch) __stcs(g_img + (((size_t)im * C + ch) * R + sy) * R + sx, acc[r][ch]);
  }
}

}  // namespace hb

using namespace hb;

template <int C>
static int launch_fwd(const float* img, const float* params, int n_crops, int crops_per_img, int R, float* out, cudaStream_t st) {
  const int max_rows = PCL_TR + 2;
  const size_t smem = sizeof(float4) * (size_t)max_rows * R;
  if (smem > 200 * 1024) { set_error("hb_pcl_fwd: img_res too large for the staged kernel"); return HB_E_UNSUPPORTED; }
  HB_CUDA(cudaFuncSetAttribute(pcl_fwd_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((R + PCL_TR - 1) / PCL_TR, n_crops);
  pcl_fwd_kernel<C><<<grid, PCL_THREADS, smem, st>>>(img, params, crops_per_img, R, out, max_rows);
  g_launches++;
  return check_launch("pcl_fwd_kernel");
}

extern "C" int hb_pcl_fwd(const float* img, const float* params, int n_crops, int crops_per_img, int C, int img_res, float* out, void* stream) {
  if (n_crops < 0 || crops_per_img <= 0 || C <= 0 || C > 4 || img_res <= 0 || (n_crops > 0 && (!img || !params || !out)) || n_crops % crops_per_img) {
    set_error("hb_pcl_fwd: bad argument (1 <= C <= 4)"); return HB_E_ARG;
  }
  if (n_crops == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  switch (C) {
    case 1: return launch_fwd<1>(img, params, n_crops, crops_per_img, img_res, out, st);
    case 2: return launch_fwd<2>(img, params, n_crops, crops_per_img, img_res, out, st);
    case 3: return launch_fwd<3>(img, params, n_crops, crops_per_img, img_res, out, st);
    default: return launch_fwd<4>(img, params, n_crops, crops_per_img, img_res, out, st);
  }
}

static const int kPclChunkImgs = 64;

extern "C" size_t hb_pcl_bwd_workspace_bytes(int n_crops, int crops_per_img, int C, int img_res) {
  (void)C;
  if (n_crops <= 0 || crops_per_img <= 0) return 0;
  const int n_imgs = n_crops / crops_per_img;
  const int chunk_crops = (n_imgs < kPclChunkImgs ? n_imgs : kPclChunkImgs) * crops_per_img;
  return sizeof(float) * (size_t)chunk_crops * (((size_t)PCL_WS_FLOATS_PER_PX * img_res * img_res + 3) & ~(size_t)3);
}

template <int C>
static int launch_bwd(const float* g_out, const float* params, int n_crops, int crops_per_img, int R, float* g_img, float* ws, cudaStream_t st) {
  const int n_imgs = n_crops / crops_per_img;
  const int chunk_imgs = n_imgs < kPclChunkImgs ? n_imgs : kPclChunkImgs;
  const int chunk_crops = chunk_imgs * crops_per_img;
  const int n_chunks = (n_imgs + chunk_imgs - 1) / chunk_imgs;
  pcl_offsets_kernel<<<n_chunks, 32, 0, st>>>(const_cast<float*>(params), n_crops, chunk_crops);
  g_launches++;
  int rc = check_launch("pcl_offsets_kernel");
  if (rc) return rc;
  // pass-A tile capacity: rows*s <= PCL_JR*(R-1)*s/(s-1) + 3s  
  const int cap = PCL_JR * R * 17 / 16 + 3 * R + 64;
  const size_t smem_mid = sizeof(float) * ((size_t)4 * R + (size_t)C * cap);
  HB_CUDA(cudaFuncSetAttribute(pcl_bwd_mid_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_mid));
  const int tiles = ((R + PCL_TS - 1) / PCL_TS) * ((R + PCL_TS - 1) / PCL_TS);
  for (int ch = 0; ch < n_chunks; ++ch) {
    const int im0 = ch * chunk_imgs;
    const int nim = (n_imgs - im0) < chunk_imgs ? (n_imgs - im0) : chunk_imgs;
    dim3 g1((R + PCL_JR - 1) / PCL_JR, nim * crops_per_img);
    pcl_bwd_mid_kernel<C><<<g1, PCL_THREADS, smem_mid, st>>>(g_out, params, im0 * crops_per_img, R, ws, cap);
    g_launches++;
    rc = check_launch("pcl_bwd_mid_kernel");
    if (rc) return rc;
    dim3 g2(tiles, nim);
    pcl_bwd_img_kernel<C><<<g2, PCL_THREADS, 0, st>>>(params, ws, im0, crops_per_img, R, g_img);
    g_launches++;
    rc = check_launch("pcl_bwd_img_kernel");
    if (rc) return rc;
  }
  return 0;
}

extern "C" int hb_pcl_bwd(const float* g_out, const float* params, int n_crops, int crops_per_img, int C, int img_res, float* g_img,
                          void* workspace, size_t workspace_bytes, void* stream) {
  if (n_crops < 0 || crops_per_img <= 0 || C <= 0 || C > 4 || img_res <= 0 || (n_crops > 0 && (!g_out || !params || !g_img || !workspace)) ||
      n_crops % crops_per_img) {
    set_error("hb_pcl_bwd: bad argument (1 <= C <= 4)"); return HB_E_ARG;
  }
  if (n_crops == 0) return 0;
  if (workspace_bytes < hb_pcl_bwd_workspace_bytes(n_crops, crops_per_img, C, img_res)) { set_error("hb_pcl_bwd: workspace too small"); return HB_E_WORKSPACE; }
  if (reinterpret_cast<uintptr_t>(workspace) & 15u) { set_error("hb_pcl_bwd: workspace must be 16-byte aligned"); return HB_E_ALIGN; }
  cudaStream_t st = (cudaStream_t)stream;
  float* ws = (float*)workspace;
  switch (C) {
    case 1: return launch_bwd<1>(g_out, params, n_crops, crops_per_img, img_res, g_img, ws, st);
    case 2: return launch_bwd<2>(g_out, params, n_crops, crops_per_img, img_res, g_img, ws, st);
    case 3: return launch_bwd<3>(g_out, params, n_crops, crops_per_img, img_res, g_img, ws, st);
    default: return launch_bwd<4>(g_out, params, n_crops, crops_per_img, img_res, g_img, ws, st);
  }
}
