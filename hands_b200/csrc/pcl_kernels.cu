// Perspective Crop Layer (src/datasets/hands_light_dataset.py:354-467), batched on the GPU, fwd + bwd.
//
// The reference runs, per sample on the CPU:  float64 homography -> fp32 grid (linspace, 3x3 matmul,
// divide, normalise) -> F.grid_sample(bilinear, zeros, align_corners=False) to s x s ->
// F.interpolate(bilinear, align_corners=True) to R x R.  Here:
//   pcl_setup_kernel : the float64 homography, one thread per crop (pcl_setup.cu)    (lines 357-386, 425-454)
//   pcl_fwd_kernel   : one CTA per (crop, 16 output rows), in sub-blocks that fit a 44 KB shared-memory budget:
//                      the source footprint is staged by TMA bulk copies, the needed rows of the s x s intermediate
//                      are gathered from it once, then resized from shared memory; R x R stores are coalesced rows.
//   pcl_bwd_mid4     : transposed resize, vertical pass first: a thread owns four adjacent output columns (16-byte
//                      loads straight from global, the band's rows split over four two-warp groups with carries), then
//                      the windowed horizontal reduction once per intermediate row, four rows per thread; writes the
//                      intermediate gradient into the chunk workspace.  (pcl_bwd_mid: scalar-column fallback.)
//   pcl_bwd_img      : transposed grid_sample, gather form through the inverse homography: the region of the
//                      intermediate grid that can reach a 32x32 source tile is staged with cp.async, its sample
//                      positions recomputed and binned into per-cell lists; each source pixel collects exactly its
//                      contributors in a fixed order, so g_img is written once -- no float atomics, no memset,
//                      bit-reproducible.  One CTA per ROW of tiles of an image.
//   pcl_bwd_scatter  : the same operator as an atomics-free scatter into a rolling shared-memory window of source rows
//                      (opt-in, HB_PCL_SCATTER=1: measured slower than the gather form; kept with its tests).
// The fp32 operation order (which products are fused) follows torch's CPU kernels exactly; it was pinned
// by bit-comparing a numpy emulation against torch 2.11 single-threaded ("PCL exactness").
#include <climits>
#include <cstdlib>
#include "hb_common.cuh"
#include "tma.cuh"

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

// same from any address space (a record cached in shared memory)
__device__ __forceinline__ Crop load_crop_any(const float* rec) {
  Crop c;
#pragma unroll
  for (int k = 0; k < 9; ++k) c.P[k] = rec[k];
  c.s = __float_as_int(rec[18]);
  c.scale = rec[19];
  c.step = rec[20];
  return c;
}

// torch.linspace(0, 1, s)[i]  (CPU kernel: start + step*i below the midpoint, fma(-step, s-1-i, end) above)
__device__ __forceinline__ float lin01(const Crop& c, int i) {
  if (c.s == 1) return 0.0f;
  return (i < c.s / 2) ? __fmul_rn(c.step, (float)i) : fmaf(-c.step, (float)(c.s - 1 - i), 1.0f);
}

// Correctly rounded a / b for a constant b with rb = RN(1/b) (Markstein's sequence: q = RN(a*rb),
// r = a - b*q exactly by FMA, RN(q + r*rb) == RN(a/b) for normal results) -- 3 instructions instead of an
// IEEE division; bit-identical to `a / b`.
__device__ __forceinline__ float div_by_const(float a, float b, float rb) {
  const float q = __fmul_rn(a, rb);
  const float r = fmaf(-b, q, a);
  return fmaf(r, rb, q);
}

// source-image sample position (in pixel-index units) of intermediate pixel (row j, col i)
__device__ __forceinline__ void sample_pos(const Crop& c, int j, int i, float R, float rcpR, float& ix, float& iy) {
  const float u = lin01(c, i), v = lin01(c, j);
  const float X = __fadd_rn(fmaf(c.P[1], v, __fmul_rn(c.P[0], u)), c.P[2]);
  const float Y = __fadd_rn(fmaf(c.P[4], v, __fmul_rn(c.P[3], u)), c.P[5]);
  const float Z = __fadd_rn(fmaf(c.P[7], v, __fmul_rn(c.P[6], u)), c.P[8]);
  const float den = __fadd_rn(1e-8f, Z);
  const float gx = __fsub_rn(__fmul_rn(div_by_const(__fdiv_rn(X, den), R, rcpR), 2.0f), 1.0f);
  const float gy = __fsub_rn(__fmul_rn(div_by_const(__fdiv_rn(Y, den), R, rcpR), 2.0f), 1.0f);
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
template <int C, int RT>   // RT: image resolution known at compile time (0 = use the runtime argument)
__global__ void __launch_bounds__(PCL_THREADS) pcl_fwd_kernel(const float* __restrict__ img, const float* __restrict__ params,
                                                              int crops_per_img, int R_arg, float* __restrict__ out, int max_rows, int smem_bytes, int tma_ok) {
  const int R = RT ? RT : R_arg;
  extern __shared__ __align__(16) float4 mid4[];  // [nrows][s] pixels, channels in .x .y .z .w; then the staged source tile
  __shared__ int reg[5];                          // staged source region: x0, y0, ncols, nrows, staged?
  __shared__ float4 rowrec[PCL_TR];               // per output row of the tile: ly0, ly1, offsets of its two source rows
  __shared__ uint64_t src_bar;
  const int q = blockIdx.y;   // (launches are split at 65,535 crops by the host)
  const Crop c = load_crop(params + (size_t)q * PF);
  const int s = c.s;
  const int Y0 = blockIdx.x * PCL_TR;
  const int Y1 = min(Y0 + PCL_TR, R) - 1;
  const float* src = img + (size_t)(q / crops_per_img) * C * R * R;
  float* dst = out + (size_t)q * C * R * R;
  const float Rf = (float)R;
  const float rcpR = __fdiv_rn(1.0f, Rf);
  const int plane = R * R;
  // The tile's output rows are processed in sub-blocks: the largest power-of-two fraction of the tile whose
  // intermediate rows (<= rows*scale + 3 of them) fit the shared-memory budget.  Boxes up to ~3/4 of the image take the
  // whole tile at once; s == R (an absent hand: empty box -> s = img_res) takes it in halves.
  int rows_sub = PCL_TR;
  while (rows_sub > 1 && ((int)ceilf((float)rows_sub * c.scale) + 3) * s * 16 > smem_bytes) rows_sub >>= 1;
  const bool staged = s <= R && ((int)ceilf((float)rows_sub * c.scale) + 3) * s * 16 <= smem_bytes;
  (void)max_rows;
  if (staged) {
   if (threadIdx.x == 0) { mbar_init(&src_bar, 1); mbar_fence_init(); }
   int nuse = 0;   // completed uses of src_bar (block-uniform) -> wait parity
   for (int y0 = Y0; y0 <= Y1; y0 += rows_sub) {
    const int y1 = min(y0 + rows_sub - 1, Y1);
    int jlo, jhi, t0, t1;
    float l0, l1;
    resize_coef(c, y0, R, jlo, t1, l0, l1);
    resize_coef(c, y1, R, t0, jhi, l0, l1);
    const int nrows = jhi - jlo + 1;
    const int n = nrows * s;
    const float inv_s = 1.0f / (float)s;
    float* stile = reinterpret_cast<float*>(mid4 + n);   // source tile [C][rows][cols], cols a multiple of 4
    // ---- TMA-staged source tile: the band's sample positions are the image of a rectangle of the intermediate grid
    //      under a homography, a convex quad whose corner box (+ margin) bounds every bilinear tap.  Warp 0 finds the
    //      box and issues one bulk copy per (channel, row); the copies land while all warps do the position math.
    //      (Measured on B200: 4.78 ms vs 4.53 ms for gathering straight through L1 -- the kernel is issue-bound, so the
    //      staging instructions cost more than the L1 misses they remove; HB_PCL_TMA=0 selects the direct gather.)
    if (threadIdx.x < 32) {
      const int lane = threadIdx.x;
      float cx = 0.f, cy = 0.f;
      if (lane < 4) sample_pos(c, (lane >> 1) ? jhi : jlo, (lane & 1) ? s - 1 : 0, Rf, rcpR, cx, cy);
      float xmn = lane < 4 ? cx : 3.0e38f, xmx = lane < 4 ? cx : -3.0e38f, ymn = lane < 4 ? cy : 3.0e38f, ymx = lane < 4 ? cy : -3.0e38f;
#pragma unroll
      for (int m = 1; m < 4; m <<= 1) {
        xmn = fminf(xmn, __shfl_xor_sync(0xffffffffu, xmn, m)); xmx = fmaxf(xmx, __shfl_xor_sync(0xffffffffu, xmx, m));
        ymn = fminf(ymn, __shfl_xor_sync(0xffffffffu, ymn, m)); ymx = fmaxf(ymx, __shfl_xor_sync(0xffffffffu, ymx, m));
      }
      xmn = __shfl_sync(0xffffffffu, xmn, 0); xmx = __shfl_sync(0xffffffffu, xmx, 0);
      ymn = __shfl_sync(0xffffffffu, ymn, 0); ymx = __shfl_sync(0xffffffffu, ymx, 0);
      int use = 0, bx0 = 0, by0 = 0, ncols = 0, nr = 0;
      if (tma_ok && xmn == xmn && xmx == xmx && ymn == ymn && ymx == ymx && xmx > -2.0f && ymx > -2.0f && xmn < Rf + 1.0f && ymn < Rf + 1.0f) {
        const int xl = max(0, (int)floorf(fmaxf(xmn, -4.0f)) - 1), xh = min(R - 1, (int)floorf(fminf(xmx, Rf + 4.0f)) + 2);
        const int yl = max(0, (int)floorf(fmaxf(ymn, -4.0f)) - 1), yh = min(R - 1, (int)floorf(fminf(ymx, Rf + 4.0f)) + 2);
        bx0 = xl & ~3;
        ncols = min(R - bx0, (xh - bx0 + 4) & ~3);
        by0 = yl;
        nr = yh - yl + 1;
        use = nr > 0 && ncols > 0 && (size_t)n * 16 + (size_t)C * nr * ncols * 4 <= (size_t)smem_bytes;
      }
      if (lane == 0) {
        reg[0] = bx0; reg[1] = by0; reg[2] = ncols; reg[3] = nr; reg[4] = use;
        if (use) mbar_arrive_expect_tx(&src_bar, (uint32_t)(C * nr * ncols * 4));
      }
      __syncwarp();
      if (use) {
        for (int k = lane; k < C * nr; k += 32) {
          const int ch = k / nr, r = k - ch * nr;
          bulk_g2s(stile + (size_t)k * ncols, src + (size_t)ch * plane + (size_t)(by0 + r) * R + bx0, (uint32_t)(ncols * 4), &src_bar);
        }
      }
    }
    // pass 1: sample positions (kept in the tile itself) -- the bulk copies land meanwhile
    for (int idx = threadIdx.x; idx < n; idx += PCL_THREADS) {
      const int jr = fast_div(idx, s, inv_s), i = idx - jr * s;
      float ix, iy;
      sample_pos(c, jlo + jr, i, Rf, rcpR, ix, iy);
      mid4[idx] = make_float4(ix, iy, 0.f, 0.f);
    }
    __syncthreads();   // region parameters and the barrier init are visible
    const int rx0 = reg[0], ry0 = reg[1], rnc = reg[2], rnr = reg[3];
    const bool tiled = reg[4] != 0;
    if (tiled) { mbar_wait(&src_bar, nuse & 1); ++nuse; }
    // pass 2: bilinear gather (from the staged tile; a tap outside it -- never expected -- falls back to global)
    for (int idx = threadIdx.x; idx < n; idx += PCL_THREADS) {
      const float4 pq = mid4[idx];
      const float ix = pq.x, iy = pq.y;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (ix > -1.0f && ix < Rf && iy > -1.0f && iy < Rf) {
        const float fx = floorf(ix), fy = floorf(iy);
        const int x0 = (int)fx, yy0 = (int)fy;
        const float wx1 = __fsub_rn(ix, fx), wx0 = __fsub_rn(__fadd_rn(fx, 1.0f), ix);
        const float wy1 = __fsub_rn(iy, fy), wy0 = __fsub_rn(__fadd_rn(fy, 1.0f), iy);
        const float wnw = __fmul_rn(wx0, wy0), wne = __fmul_rn(wx1, wy0), wsw = __fmul_rn(wx0, wy1), wse = __fmul_rn(wx1, wy1);
        const bool xa = x0 >= 0, xb = x0 + 1 < R, ya = yy0 >= 0, yb = yy0 + 1 < R;
        const int o00 = yy0 * R + x0;
        // are the (in-image) taps x0..x0+1, yy0..yy0+1 inside the staged tile?
        const int tx = x0 - rx0, ty = yy0 - ry0;
        const bool in_tile = tiled && tx >= (xa ? 0 : -1) && tx + 1 < rnc + (xb ? 0 : 1) && ty >= (ya ? 0 : -1) && ty + 1 < rnr + (yb ? 0 : 1);
        if (in_tile) {
          const float* tl = stile + ty * rnc + tx;
#pragma unroll
          for (int ch = 0; ch < C; ++ch) {
            const float nw = (xa && ya) ? tl[0] : 0.0f;
            const float ne = (xb && ya) ? tl[1] : 0.0f;
            const float sw = (xa && yb) ? tl[rnc] : 0.0f;
            const float se = (xb && yb) ? tl[rnc + 1] : 0.0f;
            float acc = __fmul_rn(nw, wnw);
            acc = fmaf(ne, wne, acc);
            acc = fmaf(sw, wsw, acc);
            acc = fmaf(se, wse, acc);
            v[ch] = acc;
            tl += rnr * rnc;
          }
        } else {
          const float* pl = src;
#pragma unroll
          for (int ch = 0; ch < C; ++ch) {
            const float nw = (xa && ya) ? __ldg(pl + o00) : 0.0f;
            const float ne = (xb && ya) ? __ldg(pl + o00 + 1) : 0.0f;
            const float sw = (xa && yb) ? __ldg(pl + o00 + R) : 0.0f;
            const float se = (xb && yb) ? __ldg(pl + o00 + R + 1) : 0.0f;
            float acc = __fmul_rn(nw, wnw);
            acc = fmaf(ne, wne, acc);
            acc = fmaf(sw, wsw, acc);
            acc = fmaf(se, wse, acc);
            v[ch] = acc;
            pl += plane;
          }
        }
      }
      mid4[idx] = make_float4(v[0], v[1], v[2], v[3]);
    }
    if (threadIdx.x < y1 - y0 + 1) {   // per-row resize coefficients of this tile, once per CTA
      int a0, a1;
      float ly0, ly1;
      resize_coef(c, y0 + threadIdx.x, R, a0, a1, ly0, ly1);
      rowrec[threadIdx.x] = make_float4(ly0, ly1, __int_as_float((a0 - jlo) * s), __int_as_float((a1 - jlo) * s));
    }
    __syncthreads();
    for (int x = threadIdx.x; x < R; x += PCL_THREADS) {
      int b0, b1;
      float lx0, lx1;
      resize_coef(c, x, R, b0, b1, lx0, lx1);
      // the two live intermediate rows stay in registers and are shifted / reloaded only when the source row index
      // of the next output row changes (block-uniform control flow)
      int cur0 = -1, cur1 = -1;
      float4 m00 = make_float4(0.f, 0.f, 0.f, 0.f), m01 = m00, m10 = m00, m11 = m00;
      float* op = dst + (size_t)y0 * R + x;
      const int nout = y1 - y0 + 1;
#pragma unroll 1
      for (int t = 0; t < nout; ++t, op += R) {
        const float4 rr = rowrec[t];
        const int o0r = __float_as_int(rr.z), o1r = __float_as_int(rr.w);
        if (o0r != cur0) {
          if (o0r == cur1) { m00 = m10; m01 = m11; }
          else { m00 = mid4[o0r + b0]; m01 = mid4[o0r + b1]; }
          cur0 = o0r;
        }
        if (o1r != cur1) { m10 = mid4[o1r + b0]; m11 = mid4[o1r + b1]; cur1 = o1r; }
        const float w00 = __fmul_rn(rr.x, lx0), w01 = __fmul_rn(rr.x, lx1), w10 = __fmul_rn(rr.y, lx0), w11 = __fmul_rn(rr.y, lx1);
        float o0 = __fmul_rn(w01, m01.x), o1 = __fmul_rn(w01, m01.y), o2 = __fmul_rn(w01, m01.z), o3 = __fmul_rn(w01, m01.w);
        o0 = fmaf(w00, m00.x, o0); o1 = fmaf(w00, m00.y, o1); o2 = fmaf(w00, m00.z, o2); o3 = fmaf(w00, m00.w, o3);
        o0 = fmaf(w10, m10.x, o0); o1 = fmaf(w10, m10.y, o1); o2 = fmaf(w10, m10.z, o2); o3 = fmaf(w10, m10.w, o3);
        o0 = fmaf(w11, m11.x, o0); o1 = fmaf(w11, m11.y, o1); o2 = fmaf(w11, m11.z, o2); o3 = fmaf(w11, m11.w, o3);
        __stcs(op, o0);
        if (C > 1) __stcs(op + plane, o1);
        if (C > 2) __stcs(op + 2 * plane, o2);
        if (C > 3) __stcs(op + 3 * plane, o3);
      }
    }
    __syncthreads();   // end of the sub-block: the tile, the row table and the region record are reused
   }
   return;
  }
  // generic path (s > R): evaluate the four intermediate pixels of every output pixel directly
  const int npix = (Y1 - Y0 + 1) * R;
  for (int idx = threadIdx.x; idx < npix; idx += PCL_THREADS) {
    const int yy = idx / R, x = idx - yy * R;
    const int y = Y0 + yy;
    int a0, a1, b0, b1;
    float ly0, ly1, lx0, lx1;
    resize_coef(c, y, R, a0, a1, ly0, ly1);
    resize_coef(c, x, R, b0, b1, lx0, lx1);
    const float w00 = __fmul_rn(ly0, lx0), w01 = __fmul_rn(ly0, lx1), w10 = __fmul_rn(ly1, lx0), w11 = __fmul_rn(ly1, lx1);
    float px[4], py[4];
    sample_pos(c, a0, b0, Rf, rcpR, px[0], py[0]);
    sample_pos(c, a0, b1, Rf, rcpR, px[1], py[1]);
    sample_pos(c, a1, b0, Rf, rcpR, px[2], py[2]);
    sample_pos(c, a1, b1, Rf, rcpR, px[3], py[3]);
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

// ---- forward, default ("fast") form -----------------------------------------------------------------
// Same decomposition and the same sample positions as pcl_fwd_kernel (the reference's rounding sequence
// X/(1e-8+Z) -> /R -> *2-1 -> un-normalise is kept, because its -1/+1 steps quantise the position to ~7e-6 px and a
// white-noise image amplifies any other choice past 1e-5), but with fewer instructions per pixel:
//   * both IEEE divisions share one refined reciprocal (MUFU.RCP + one Newton step, then Markstein's q + r*y correction);
//   * position and gather are one pass (no round trip of the positions through shared memory); taps are clamped into the
//     image and their WEIGHTS masked, so the twelve loads are unconditional;
//   * the resize is evaluated separably: the two live intermediate rows are interpolated horizontally once per row change
//     (3 FMUL + 3 FMA) and every output pixel is 3 FMUL + 3 FMA instead of 4 weight products + 16 FMAs.  This is the only
//     arithmetic difference from torch's kernel (which forms the four 2-D weights first): a few ulp of the output value.
//   * SrcT = uint8_t: the source image is the data loader's 8-bit image; (u/255 - mean)/std (torchvision Normalize as the
//     reference applies it, hands_light_dataset.py:177-184) is a 256-entry table per channel built once per CTA, so the
//     staged tile is a quarter of the bytes and the PCIe copy of the step's images shrinks fourfold.
// HB_PCL_EXACT=1 / hb_pcl_set_exact(1) selects pcl_fwd_kernel (bit-identical to torch on > 99.9 % of the pixels).
struct PclNorm { float mean[4]; float std[4]; };

__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// a / b given y ~ RN(1/b): q = RN(a*y), r = a - b*q (exact by FMA), RN(q + r*y)
__device__ __forceinline__ float div_refined(float a, float b, float y) {
  const float q = __fmul_rn(a, y);
  const float r = fmaf(-b, q, a);
  return fmaf(r, y, q);
}

__device__ __forceinline__ void sample_pos_uv(const Crop& c, float u, float v, float R, float rcpR, float& ix, float& iy) {
  const float X = __fadd_rn(fmaf(c.P[1], v, __fmul_rn(c.P[0], u)), c.P[2]);
  const float Y = __fadd_rn(fmaf(c.P[4], v, __fmul_rn(c.P[3], u)), c.P[5]);
  const float Z = __fadd_rn(fmaf(c.P[7], v, __fmul_rn(c.P[6], u)), c.P[8]);
  const float den = __fadd_rn(1e-8f, Z);
  float y = rcp_approx(den);
  y = fmaf(fmaf(-den, y, 1.0f), y, y);   // one Newton step: y == RN(1/den) except in rare half-way cases
  const float gx = fmaf(div_by_const(div_refined(X, den, y), R, rcpR), 2.0f, -1.0f);
  const float gy = fmaf(div_by_const(div_refined(Y, den, y), R, rcpR), 2.0f, -1.0f);
  const float half = R * 0.5f;
  ix = fmaf(__fadd_rn(gx, 1.0f), half, -0.5f);
  iy = fmaf(__fadd_rn(gy, 1.0f), half, -0.5f);
}

// global-memory tap for the un-staged fallback (fp32 image, or 8-bit image through the normalisation table)
__device__ __forceinline__ float src_ldg(const float* p, const float*) { return __ldg(p); }
__device__ __forceinline__ float src_ldg(const uint8_t* p, const float* lut) { return lut[__ldg(p)]; }

// bilinear gather of one pixel straight from global memory with zero padding (fallback of the fast kernel)
template <int C, typename SrcT>
__device__ __forceinline__ void gather_global(const SrcT* __restrict__ src, const float* lut, int R, int plane, float ix, float iy, float* v) {
  const float Rf = (float)R;
#pragma unroll
  for (int ch = 0; ch < 4; ++ch) v[ch] = 0.0f;
  if (!(ix > -1.0f && ix < Rf && iy > -1.0f && iy < Rf)) return;
  const float fx = floorf(ix), fy = floorf(iy);
  const int x0 = (int)fx, y0 = (int)fy;
  const float wx1 = __fsub_rn(ix, fx), wx0 = __fsub_rn(__fadd_rn(fx, 1.0f), ix);
  const float wy1 = __fsub_rn(iy, fy), wy0 = __fsub_rn(__fadd_rn(fy, 1.0f), iy);
  const float wnw = __fmul_rn(wx0, wy0), wne = __fmul_rn(wx1, wy0), wsw = __fmul_rn(wx0, wy1), wse = __fmul_rn(wx1, wy1);
  const bool xa = x0 >= 0, xb = x0 + 1 < R, ya = y0 >= 0, yb = y0 + 1 < R;
  const SrcT* pl = src + y0 * R + x0;
#pragma unroll
  for (int ch = 0; ch < C; ++ch) {
    const float* lc = lut + (sizeof(SrcT) == 1 ? ch * 256 : 0);
    const float nw = (xa && ya) ? src_ldg(pl, lc) : 0.0f;
    const float ne = (xb && ya) ? src_ldg(pl + 1, lc) : 0.0f;
    const float sw = (xa && yb) ? src_ldg(pl + R, lc) : 0.0f;
    const float se = (xb && yb) ? src_ldg(pl + R + 1, lc) : 0.0f;
    float acc = __fmul_rn(nw, wnw);
    acc = fmaf(ne, wne, acc);
    acc = fmaf(sw, wsw, acc);
    acc = fmaf(se, wse, acc);
    v[ch] = acc;
    pl += plane;
  }
}

constexpr int PCL_TV = 40;   // rows of the per-sub-block linspace table (an intermediate band never has more: 16 * scale + 3, scale <= 1)

constexpr int PCL_XB = 224;   // output columns per CTA of the fast forward at R = 224 (one column block; 112 = two blocks measured slower:
                              // the phases between barriers get too short)
constexpr int PCL_NSEG = 2;   // row segments per crop: a CTA walks the bands of one segment, so the per-crop set-up is paid once per 7 bands

// idx / d for 0 <= idx < 2^16, 1 <= d <= 2^16, without an integer division (~20 instructions): float quotient + fix-up
__device__ __forceinline__ int small_div(int idx, int d) {
  int q = __float2int_rz(__fdividef((float)idx + 0.5f, (float)d));
  const int r = idx - q * d;
  if (r < 0) --q; else if (r >= d) ++q;
  return q;
}

template <int C, int RT, typename SrcT, bool V2>   // V2: out is 8-byte aligned and R even -> 64-bit stores
__global__ void __launch_bounds__(PCL_THREADS, 5) pcl_fwd_fast_kernel(const SrcT* __restrict__ img, const float* __restrict__ params, int crops_per_img, unsigned cpi_mul,
                                                                      int R_arg, float* __restrict__ out, int smem_bytes, int tma_ok, int nxb_arg, PclNorm nrm) {
  const int R = RT ? RT : R_arg;
  constexpr bool U8 = sizeof(SrcT) == 1;
  extern __shared__ __align__(16) float4 mid4[];  // [cap] band pixels, channels in .x .y .z .w; then the staged source tile (fp32)
  __shared__ int reg[2][6];                       // tile record of the current / the next band
  __shared__ __align__(16) float4 rowrec[PCL_TR];
  __shared__ uint64_t src_bar;
  __shared__ float lut[U8 ? C * 256 : 1];
  __shared__ float tu[RT ? PCL_XB + 8 : 1];   // linspace(0,1,s) of the CTA's intermediate columns (compile-time resolution only)
  __shared__ float tv[RT ? PCL_TV : 1];       // ... and of the band's intermediate rows
  __shared__ __align__(16) float2 coltab[RT ? PCL_XB : 1];   // per output column: (weight of the upper tap, lower tap byte offset in a band row)
  const int q = blockIdx.y;   // band index on grid.x (fastest): the CTAs of one crop are adjacent in launch order -- measured 401 vs 420 us
                              // per 2048 crops against crops on grid.x; launches are split at 65,535 crops by the host
  const Crop c = load_crop(params + (size_t)q * PF);
  const int s = c.s;
  // CTA = (crop, row segment, block of output columns).  The column split halves the intermediate band and its source
  // footprint (band + staged tile fit the shared-memory budget of 5 CTAs/SM for every box up to the image size); walking
  // the segment's bands in one CTA pays the per-crop set-up once and lets the next band's tile fly during the resize.
  const int nxb = RT ? RT / PCL_XB : nxb_arg;
  const int yseg = RT ? (int)(blockIdx.x / (RT / PCL_XB)) : (int)blockIdx.x / nxb, xb = blockIdx.x - yseg * nxb;
  const int xbw = RT ? PCL_XB : (((R + nxb - 1) / nxb + 1) & ~1);
  const int segrows = RT ? ((RT / PCL_NSEG + PCL_TR - 1) / PCL_TR) * PCL_TR : (((R + PCL_NSEG - 1) / PCL_NSEG + PCL_TR - 1) / PCL_TR) * PCL_TR;
  const int X0 = xb * xbw, X1 = min(X0 + xbw, R) - 1;
  const int Y0 = yseg * segrows;
  const int Y1 = min(Y0 + segrows, R) - 1;
  const int plane = R * R;
  const SrcT* src = img + (size_t)(cpi_mul ? __umulhi((unsigned)q, cpi_mul) : (unsigned)q) * C * plane;   // q / crops_per_img (0: one crop per image)
  float* dst = out + (size_t)q * C * plane;
  const float Rf = (float)R;
  const float rcpR = __fdiv_rn(1.0f, Rf);
  const int tid = threadIdx.x;
  if (Y0 > Y1 || X0 > X1) return;
  int ilo, ihi;
  {
    int t0, t1;
    float l0, l1;
    resize_coef(c, X0, R, ilo, t1, l0, l1);
    resize_coef(c, X1, R, t0, ihi, l0, l1);
  }
  const int ncol = ihi - ilo + 1;   // intermediate columns this CTA needs
  if (U8) {
    for (int e = tid; e < C * 256; e += PCL_THREADS) {
      const int ch = e >> 8;
      lut[e] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)(e & 255), 255.0f), nrm.mean[ch]), nrm.std[ch]);
    }
  }
  if (RT) {
    if (tid < ncol && tid < PCL_XB + 8) tu[tid] = lin01(c, ilo + tid);
    for (int x = tid; x <= X1 - X0; x += PCL_THREADS) {
      int b0, b1;
      float lx0, lx1;
      resize_coef(c, X0 + x, R, b0, b1, lx0, lx1);
      coltab[x] = make_float2(lx1, __int_as_float((b0 - ilo) * 16));   // lx0 = 1 - lx1, b1 = b0 + (b0 < s-1): recomputed exactly
    }
  }
  // rows per band: the band (16 B per intermediate pixel) must fit
  int rows_sub = PCL_TR;
  auto band_rows = [&](int rs) { return (int)ceilf((float)rs * c.scale) + 3; };
  // (the band alone decides: a tile that does not fit beside it is gathered through L1/L2 instead -- shrinking the band until
  //  band + tile fit measured slower, 491 vs 401 us per 2048 crops: shorter phases between barriers, more halo rows)
  while (rows_sub > 1 && band_rows(rows_sub) * ncol * 16 > smem_bytes) rows_sub >>= 1;
  const int cap = band_rows(rows_sub) * ncol;          // band capacity in pixels; the tile lives behind it
  const bool fits = cap * 16 <= smem_bytes && ncol <= (RT ? PCL_XB + 8 : 1 << 30);
  if (s <= R && fits) {
    float* stile = reinterpret_cast<float*>(mid4 + cap);   // source tile [C][rows][cols] fp32, cols a multiple of 4
    const int tile_bytes = smem_bytes - cap * 16;
    // warp 0: tile record of the band starting at output row y0 (+ the TMA bulk copies for an fp32 image)
    auto issue_tile = [&](int y0, int slot, int jdone) {   // jdone: last intermediate row the previous band left in the ring (-1: none)
      // The band's sample positions are the image of a rectangle of the intermediate grid under a homography: a convex
      // quad whose corner box (+ the bilinear margin) bounds every tap.  The tile spans that box in UNCLAMPED image
      // coordinates, at most [-1, R]: the parts outside the image are the reference's zero padding, so the gather needs
      // neither bounds predicates nor clamps.  fp32 images: one TMA bulk copy per (channel, in-image row).
      const int lane = tid;
      const int y1 = min(y0 + rows_sub - 1, Y1);
      int jlo, jhi, t0, t1;
      float l0, l1;
      resize_coef(c, y0, R, jlo, t1, l0, l1);
      resize_coef(c, y1, R, t0, jhi, l0, l1);
      jlo = max(jlo, jdone + 1);   // the band's rows up to jdone are already in the ring: the tile covers the new rows only
      float cx = 0.f, cy = 0.f;
      if (lane < 4) sample_pos_uv(c, lin01(c, (lane & 1) ? ihi : ilo), lin01(c, (lane >> 1) ? jhi : jlo), Rf, rcpR, cx, cy);
      float xmn = lane < 4 ? cx : 3.0e38f, xmx = lane < 4 ? cx : -3.0e38f, ymn = lane < 4 ? cy : 3.0e38f, ymx = lane < 4 ? cy : -3.0e38f;
#pragma unroll
      for (int m = 1; m < 4; m <<= 1) {
        xmn = fminf(xmn, __shfl_xor_sync(0xffffffffu, xmn, m)); xmx = fmaxf(xmx, __shfl_xor_sync(0xffffffffu, xmx, m));
        ymn = fminf(ymn, __shfl_xor_sync(0xffffffffu, ymn, m)); ymx = fmaxf(ymx, __shfl_xor_sync(0xffffffffu, ymx, m));
      }
      xmn = __shfl_sync(0xffffffffu, xmn, 0); xmx = __shfl_sync(0xffffffffu, xmx, 0);
      ymn = __shfl_sync(0xffffffffu, ymn, 0); ymx = __shfl_sync(0xffffffffu, ymx, 0);
      int use = 0, bx0 = 0, by0 = 0, ncols = 0, nr = 0;
      if (tma_ok && jlo <= jhi && xmn == xmn && xmx == xmx && ymn == ymn && ymx == ymx && xmx > -2.0f && ymx > -2.0f && xmn < Rf + 1.0f && ymn < Rf + 1.0f) {
        const int xl = max(-1, (int)floorf(fmaxf(xmn, -4.0f)) - 1), xh = min(R, (int)floorf(fminf(xmx, Rf + 4.0f)) + 2);
        const int yl = max(-1, (int)floorf(fmaxf(ymn, -4.0f)) - 1), yh = min(R, (int)floorf(fminf(ymx, Rf + 4.0f)) + 2);
        bx0 = xl & ~3;                      // two's complement: -1 -> -4
        ncols = (xh - bx0 + 4) & ~3;
        by0 = yl;
        nr = yh - yl + 1;
        use = nr > 0 && ncols > 0 && C * nr * ncols * 4 <= tile_bytes;
      }
      const int cx0 = max(bx0, 0), cx1 = min(bx0 + ncols, R);        // in-image column range (multiples of 4)
      const int ry0 = max(by0, 0), ry1 = min(by0 + nr, R);           // in-image row range
      const int pad = use && (bx0 < 0 || bx0 + ncols > R || by0 < 0 || by0 + nr > R);
      if (lane == 0) {
        reg[slot][0] = bx0; reg[slot][1] = by0; reg[slot][2] = ncols; reg[slot][3] = nr; reg[slot][4] = use; reg[slot][5] = pad;
        if (use && !U8) mbar_arrive_expect_tx(&src_bar, (uint32_t)(C * (ry1 - ry0) * (cx1 - cx0) * 4));
      }
      __syncwarp();
      if (use && !U8) {
        fence_proxy_async();   // the previous band's generic-proxy accesses to the tile area are ordered before these copies
        const int nrin = ry1 - ry0;
        for (int k = lane; k < C * nrin; k += 32) {
          const int ch = small_div(k, nrin), r = ry0 + (k - ch * nrin);
          bulk_g2s(stile + ((size_t)(ch * nr + (r - by0)) * ncols + (cx0 - bx0)),
                   reinterpret_cast<const float*>(src) + (size_t)ch * plane + (size_t)r * R + cx0, (uint32_t)((cx1 - cx0) * 4), &src_bar);
        }
      }
    };
    if (tid == 0) { mbar_init(&src_bar, 1); mbar_fence_init(); }
    __syncwarp();
    if (tid < 32) issue_tile(Y0, 0, -1);
    int nuse = 0, slot = 0;
    // The band buffer is a ring of RG intermediate rows: consecutive bands share their last 2-3 rows (the resize's halo), which
    // stay where they are -- a band gathers only the rows above the previous band's last one, and its tile covers only those.
    const int RG = band_rows(rows_sub);
    int jdone = -1, ring_lo = 0, jlo_prev = 0;   // ring_lo: ring slot of row jlo_prev
    const int q256 = small_div(PCL_THREADS, ncol), r256 = PCL_THREADS - q256 * ncol;   // idx += 256  <=>  (row += q256, col += r256) with one carry
    for (int y0 = Y0; y0 <= Y1; y0 += rows_sub, slot ^= 1) {
      const int y1 = min(y0 + rows_sub - 1, Y1);
      int jlo, jhi, t0, t1;
      float l0, l1;
      resize_coef(c, y0, R, jlo, t1, l0, l1);
      resize_coef(c, y1, R, t0, jhi, l0, l1);
      // ring slot of row jlo (rows advance by less than RG per band)
      if (jdone >= 0) { ring_lo += jlo - jlo_prev; if (ring_lo >= RG) ring_lo -= RG; }
      jlo_prev = jlo;
      const int jnew = max(jlo, jdone + 1);            // first row this band has to gather
      const int nrows = jhi - jnew + 1;                // new rows (>= 0)
      const int n = nrows * ncol;                      // <= cap
      int snew = ring_lo + (jnew - jlo);               // ring slot of row jnew
      if (snew >= RG) snew -= RG;
      if (tid >= 32 && tid - 32 < y1 - y0 + 1) {   // per-row resize coefficients of this band (ring-row byte offsets)
        int a0, a1;
        float ly0, ly1;
        resize_coef(c, y0 + tid - 32, R, a0, a1, ly0, ly1);
        int s0 = ring_lo + (a0 - jlo), s1 = ring_lo + (a1 - jlo);
        if (s0 >= RG) s0 -= RG;
        if (s1 >= RG) s1 -= RG;
        rowrec[tid - 32] = make_float4(ly0, ly1, __int_as_float(s0 * ncol * 16), __int_as_float(s1 * ncol * 16));
      }
      if (RT && tid >= 64 && tid - 64 < nrows && tid - 64 < PCL_TV) tv[tid - 64] = lin01(c, jnew + tid - 64);
      jdone = jhi;
      __syncthreads();   // tile record, row table, linspace tables, barrier init (and the LUT) are visible
      const int rx0 = reg[slot][0], ry0 = reg[slot][1], rnc = reg[slot][2], rnr = reg[slot][3];
      const bool tiled = reg[slot][4] != 0;
      const bool padded = reg[slot][5] != 0;
      if (tiled && (U8 || padded)) {
        // 8-bit image: the threads stage the tile themselves, 4 pixels per 32-bit load, normalised through the table,
        // out-of-image groups zero.  fp32 image: only the out-of-image groups are written (the bulk copies fill the rest).
        const int nc4 = rnc >> 2;
        const int tot = C * rnr * nc4;
        const float inv_nc4 = 1.0f / (float)nc4;
        for (int e = tid; e < tot; e += PCL_THREADS) {
          const int rowi = fast_div(e, nc4, inv_nc4), c4 = e - rowi * nc4;   // rowi = ch * rnr + r
          const int ch = small_div(rowi, rnr), r = rowi - ch * rnr;
          const int y = ry0 + r, x = rx0 + 4 * c4;
          const bool inside = y >= 0 && y < R && x >= 0 && x < R;            // x, R multiples of 4: a group is all in or all out
          float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
          if (U8) {
            if (inside) {
              const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(src) + (size_t)ch * plane + (size_t)y * R + x));
              const float* lc = lut + ch * 256;
              val = make_float4(lc[w & 255u], lc[(w >> 8) & 255u], lc[(w >> 16) & 255u], lc[w >> 24]);
            }
            reinterpret_cast<float4*>(stile)[e] = val;
          } else if (!inside) {
            reinterpret_cast<float4*>(stile)[e] = val;
          }
        }
      }
      if (tiled) {
        if (U8) __syncthreads();
        else { mbar_wait(&src_bar, nuse & 1); ++nuse; if (padded) __syncthreads(); }
      }
      // gather: one pass, thread = intermediate pixel
      {
        int jr = small_div(tid, ncol), i = tid - jr * ncol;
        const float xlo = (float)rx0, xhi = (float)(rx0 + rnc - 1), ylo = (float)ry0, yhi = (float)(ry0 + rnr - 1);
        const int toff = ry0 * rnc + rx0;
        const int chs = rnr * rnc;
        for (int idx = tid; idx < n; idx += PCL_THREADS) {
          float ix, iy;
          if (RT) sample_pos_uv(c, tu[i], tv[min(jr, PCL_TV - 1)], Rf, rcpR, ix, iy);
          else sample_pos_uv(c, lin01(c, ilo + i), lin01(c, jnew + jr), Rf, rcpR, ix, iy);
          float v[4];
          // both taps of each axis inside the tile?  (float compares: a NaN position fails them and takes the fallback)
          if (tiled && ix >= xlo && ix < xhi && iy >= ylo && iy < yhi) {
            const float fx = floorf(ix), fy = floorf(iy);
            const float wx1 = __fsub_rn(ix, fx), wx0 = __fsub_rn(__fadd_rn(fx, 1.0f), ix);
            const float wy1 = __fsub_rn(iy, fy), wy0 = __fsub_rn(__fadd_rn(fy, 1.0f), iy);
            const float wnw = __fmul_rn(wx0, wy0), wne = __fmul_rn(wx1, wy0), wsw = __fmul_rn(wx0, wy1), wse = __fmul_rn(wx1, wy1);
            const float* tl = stile + ((int)fy * rnc + (int)fx - toff);
            const float* tb = tl + rnc;
            v[3] = 0.0f;
#pragma unroll
            for (int ch = 0; ch < C; ++ch) {
              float acc = __fmul_rn(tl[0], wnw);
              acc = fmaf(tl[1], wne, acc);
              acc = fmaf(tb[0], wsw, acc);
              acc = fmaf(tb[1], wse, acc);
              v[ch] = acc;
              tl += chs; tb += chs;
            }
#pragma unroll
            for (int ch = C; ch < 3; ++ch) v[ch] = 0.0f;
          } else {   // tile not staged (does not fit / unaligned image), position outside the image, or a tap outside the tile
            gather_global<C, SrcT>(src, lut, R, plane, ix, iy, v);
          }
          {
            int sl = snew + jr;   // ring slot of this row
            if (sl >= RG) sl -= RG;
            mid4[sl * ncol + i] = make_float4(v[0], v[1], v[2], v[3]);
          }
          i += r256; jr += q256;
          if (i >= ncol) { i -= ncol; ++jr; }
        }
      }
      __syncthreads();   // the band is complete; the tile is free
      // the next band's tile flies while this band is resized: warp 0 is the producer, warps 1..7 resize
      if (tid < 32 && y0 + rows_sub <= Y1) issue_tile(y0 + rows_sub, slot ^ 1, jhi);
      // separable resize: a thread owns two adjacent output columns of one group of the band's rows; the two live
      // intermediate rows, interpolated horizontally, stay in registers while consecutive output rows share them
      if (tid >= 32) {
        const int nout = y1 - y0 + 1;
        const int rt = tid - 32;                                   // 224 resize threads
        const int gthreads = (PCL_THREADS - 32) / 2;               // two row groups of 112 threads (112 column pairs at R = 224)
        const int NG = 2;
        const int grows = (nout + NG - 1) / NG;
        const int g = rt >= gthreads ? 1 : 0, tg = rt - g * gthreads;
        const int tA = g * grows, tB = min(tA + grows, nout);
        const char* midb = reinterpret_cast<const char*>(mid4);
        for (int x = X0 + 2 * tg; x <= X1; x += 2 * gthreads) {
          int ob0[2], ob1[2];   // byte offsets of the two horizontal taps inside a band row
          float lx0[2], lx1[2];
          if (RT) {
            const float4 ct = *reinterpret_cast<const float4*>(coltab + (x - X0));   // x - X0 even: both columns in one 16-byte load
            lx1[0] = ct.x; ob0[0] = __float_as_int(ct.y); lx1[1] = ct.z; ob0[1] = __float_as_int(ct.w);
#pragma unroll
            for (int k = 0; k < 2; ++k) { lx0[k] = __fsub_rn(1.0f, lx1[k]); ob1[k] = ob0[k] + (ob0[k] < (s - 1 - ilo) * 16 ? 16 : 0); }
          } else {
            int b0, b1;
            resize_coef(c, x, R, b0, b1, lx0[0], lx1[0]);
            ob0[0] = (b0 - ilo) * 16; ob1[0] = (b1 - ilo) * 16;
            resize_coef(c, min(x + 1, R - 1), R, b0, b1, lx0[1], lx1[1]);
            ob0[1] = (b0 - ilo) * 16; ob1[1] = (b1 - ilo) * 16;
          }
          int cur0 = -1, cur1 = -1;
          float h0[2][4], h1[2][4];
#pragma unroll
          for (int k = 0; k < 2; ++k)
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) { h0[k][ch] = 0.f; h1[k][ch] = 0.f; }
          auto hrow = [&](int ob, float (*h)[4]) {
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              const float4 a = *reinterpret_cast<const float4*>(midb + ob + ob0[k]), b = *reinterpret_cast<const float4*>(midb + ob + ob1[k]);
              h[k][0] = fmaf(lx1[k], b.x, __fmul_rn(lx0[k], a.x));
              if (C > 1) h[k][1] = fmaf(lx1[k], b.y, __fmul_rn(lx0[k], a.y));
              if (C > 2) h[k][2] = fmaf(lx1[k], b.z, __fmul_rn(lx0[k], a.z));
              if (C > 3) h[k][3] = fmaf(lx1[k], b.w, __fmul_rn(lx0[k], a.w));
            }
          };
          float* op = dst + (size_t)(y0 + tA) * R + x;
          const bool two = x + 1 <= X1;
#pragma unroll 1
          for (int t = tA; t < tB; ++t, op += R) {
            const float4 rr = rowrec[t];
            const int o0r = __float_as_int(rr.z), o1r = __float_as_int(rr.w);
            if (o0r != cur0) {   // uniform across the row group
              if (o0r == cur1) {
#pragma unroll
                for (int k = 0; k < 2; ++k)
#pragma unroll
                  for (int ch = 0; ch < C; ++ch) h0[k][ch] = h1[k][ch];
              } else hrow(o0r, h0);
              cur0 = o0r;
            }
            if (o1r != cur1) { hrow(o1r, h1); cur1 = o1r; }
#pragma unroll
            for (int ch = 0; ch < C; ++ch) {
              // both columns in one packed multiply + one packed FMA (bit-identical lanes)
              const float2 o2 = __ffma2_rn(make_float2(rr.y, rr.y), make_float2(h1[0][ch], h1[1][ch]),
                                           __fmul2_rn(make_float2(rr.x, rr.x), make_float2(h0[0][ch], h0[1][ch])));
              if (V2) __stcs(reinterpret_cast<float2*>(op + ch * plane), o2);
              else { __stcs(op + ch * plane, o2.x); if (two) __stcs(op + ch * plane + 1, o2.y); }
            }
          }
        }
      }
      __syncthreads();   // end of the band: the band buffer, the row tables and this band's tile record are reused
    }
    return;
  }
  // generic path (s > R, or an intermediate row that does not fit): the four intermediate pixels of every output pixel directly
  if (U8) __syncthreads();
  const int wcols = X1 - X0 + 1;
  const int npix = (Y1 - Y0 + 1) * wcols;
  for (int idx = tid; idx < npix; idx += PCL_THREADS) {
    const int yy = idx / wcols, x = X0 + idx - yy * wcols;
    const int y = Y0 + yy;
    int a0, a1, b0, b1;
    float ly0, ly1, lx0, lx1;
    resize_coef(c, y, R, a0, a1, ly0, ly1);
    resize_coef(c, x, R, b0, b1, lx0, lx1);
    const float w00 = __fmul_rn(ly0, lx0), w01 = __fmul_rn(ly0, lx1), w10 = __fmul_rn(ly1, lx0), w11 = __fmul_rn(ly1, lx1);
    float px[4], py[4], t[4][4];
    sample_pos(c, a0, b0, Rf, rcpR, px[0], py[0]);
    sample_pos(c, a0, b1, Rf, rcpR, px[1], py[1]);
    sample_pos(c, a1, b0, Rf, rcpR, px[2], py[2]);
    sample_pos(c, a1, b1, Rf, rcpR, px[3], py[3]);
#pragma unroll
    for (int k = 0; k < 4; ++k) gather_global<C, SrcT>(src, lut, R, plane, px[k], py[k], t[k]);
#pragma unroll
    for (int ch = 0; ch < C; ++ch) {
      float acc = __fmul_rn(w01, t[1][ch]);
      acc = fmaf(w00, t[0][ch], acc);
      acc = fmaf(w10, t[2][ch], acc);
      acc = fmaf(w11, t[3][ch], acc);
      dst[((size_t)ch * R + y) * R + x] = acc;
    }
  }
}

// ---- backward -------------------------------------------------------------------------------------
// Chunk workspace, per crop (offset params[21], in floats, a multiple of 4): G float4[s*s] intermediate
// gradient (channels in x,y,z,w); 4*s*s floats per crop.  (Sample positions are recomputed by the consumer.)
constexpr int PCL_WS_FLOATS_PER_PX = 4;

// exclusive scan of the per-crop workspace sizes inside each chunk (one 1024-thread block per chunk)
__global__ void __launch_bounds__(1024) pcl_offsets_kernel(float* __restrict__ params, int n_crops, int chunk_crops, int img_res) {
  __shared__ int warp_tot[32];
  __shared__ int carry;
  const int q0 = blockIdx.x * chunk_crops;
  const int q1 = min(q0 + chunk_crops, n_crops);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = q0; base < q1; base += 1024) {
    const int q = base + threadIdx.x;
    int sz = 0;
    if (q < q1) {
      const int s = __float_as_int(params[(size_t)q * PF + 18]);
      sz = s > img_res ? 0 : ((PCL_WS_FLOATS_PER_PX * s * s + 3) & ~3);  // keep every crop's float4 array 16-byte aligned
    }
    int inc = sz;  // inclusive scan within the warp
#pragma unroll
    for (int m = 1; m < 32; m <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, m); if (lane >= m) inc += t; }
    if (lane == 31) warp_tot[wid] = inc;
    __syncthreads();
    if (wid == 0) {
      int t = warp_tot[lane];
#pragma unroll
      for (int m = 1; m < 32; m <<= 1) { const int u = __shfl_up_sync(0xffffffffu, t, m); if (lane >= m) t += u; }
      warp_tot[lane] = t;  // inclusive totals of the warps
    }
    __syncthreads();
    const int before = carry + (wid > 0 ? warp_tot[wid - 1] : 0) + inc - sz;
    if (q < q1) params[(size_t)q * PF + 21] = __int_as_float(before);
    __syncthreads();
    if (threadIdx.x == 1023) carry = before + sz;
    __syncthreads();
  }
}

constexpr int PCL_JR = 16;  // intermediate rows per CTA


// approximate sample position for the backward pass (the gradient is continuous in the position, so the
// correctly-rounded divisions of the forward are not needed here)
__device__ __forceinline__ void sample_pos_fast(const Crop& c, int j, int i, float R, float& ix, float& iy) {
  const float u = lin01(c, i), v = lin01(c, j);
  const float X = fmaf(c.P[1], v, c.P[0] * u) + c.P[2];
  const float Y = fmaf(c.P[4], v, c.P[3] * u) + c.P[5];
  const float Z = fmaf(c.P[7], v, c.P[6] * u) + c.P[8];
  const float iz = __frcp_rn(1e-8f + Z);
  ix = X * iz - 0.5f;   // ((x/R*2-1)+1)*R/2 - 0.5 == x - 0.5 up to rounding
  iy = Y * iz - 0.5f;
  (void)R;
}

// 1-D transposed linear interpolation tables (same for both axes of a crop):
//   l1[d]     weight of output index d on its upper source index i0(d)+1  (1-l1[d] on i0(d))
//   start[m]  first output index d with i0(d) >= m; run(m) = [start[m], start[m+1]) are the d with i0(d) == m
// mid[m] = sum_{d in run(m)} (1-l1[d]) g[d] + sum_{d in run(m-1)} l1[d] g[d]  (+ run(s-1)'s l1 part when m == s-1,
// because the upper index is clamped there).
__device__ __forceinline__ void build_tables(const Crop& c, int R, float* tl1, int* start) {
  const int s = c.s;
  for (int m = threadIdx.x; m <= s; m += blockDim.x) start[m] = R;
  __syncthreads();
  for (int d = threadIdx.x; d < R; d += blockDim.x) {
    int i0, i1, p0 = -1, p1;
    float l0, l1, q0, q1;
    resize_coef(c, d, R, i0, i1, l0, l1);
    if (d > 0) resize_coef(c, d - 1, R, p0, p1, q0, q1);
    tl1[d] = l1;
    for (int m = p0 + 1; m <= i0; ++m) start[m] = d;
  }
  __syncthreads();
}

// Transposed resize (g_out -> intermediate gradient), vertical-first form.
// One CTA (256 threads) per (crop, band of PCL_JR intermediate rows).
//   vertical pass  : thread = OUTPUT column x walks down the output rows the band needs, reading g_out straight from
//                    global memory (a warp reads 128 contiguous bytes per row and channel; four rows are in flight per
//                    thread).  Row y adds (1-l1[y]) g to the accumulator of intermediate row i0(y) and l1[y] g to the one
//                    of i0(y)+1 -- six FMAs per element, no barrier; a finished row is parked in this thread's column of
//                    the band buffer in shared memory.
//   horizontal pass: after one barrier, thread = (intermediate row, intermediate column) reduces its contiguous window
//                    run(i-1) U run(i) of the band buffer and writes the gradient with its sample position.  The windowed
//                    reduction runs once per intermediate row (s of them) instead of once per output row (R of them).
// (A variant that streamed the rows through a TMA-fed shared-memory ring measured slower, 4.10 ms vs this one: with only
//  six FMAs per element the ring's per-stage barriers dominated.)
constexpr int PCL_MT = 256;  // threads per CTA

template <int C, int RT>   // RT: image resolution known at compile time (0 = use the runtime argument)
__global__ void __launch_bounds__(PCL_MT) pcl_bwd_mid_kernel(const float* __restrict__ g_out, const float* __restrict__ params,
                                                             int q_base, int R_arg, float* __restrict__ ws, int use_tma) {
  const int R = RT ? RT : R_arg;
  (void)use_tma;
  extern __shared__ __align__(16) float sm[];
  const int q = q_base + blockIdx.y;
  const float* rec = params + (size_t)q * PF;
  const int j0 = blockIdx.x * PCL_JR;
  {
    const int s_only = __float_as_int(__ldg(rec + 18));   // one load decides whether this band exists
    if (j0 >= s_only || s_only > (RT ? RT : R_arg)) return;
  }
  const Crop c = load_crop(rec);
  const int s = c.s;
  if (j0 >= s || s > R) return;   // s > R is outside the supported domain of the backward (workspace sized for s <= R)
  const int j1 = min(j0 + PCL_JR, s) - 1;
  float* tl1 = sm;                                   // [R]
  int* start = reinterpret_cast<int*>(sm + R);       // [R+1]
  float* Vb = sm + 2 * R + 4;                        // [PCL_JR][C][R] finished intermediate rows at output-column resolution
  float* base = ws + __float_as_int(__ldg(rec + 21));
  float4* G = reinterpret_cast<float4*>(base);
  const float* go = g_out + (size_t)q * C * R * R;
  const int tid = threadIdx.x;
  if (s <= R) {
    build_tables(c, R, tl1, start);
    const int jfirst = max(j0 - 1, 0);
    const int ylo = start[jfirst], yhi = start[j1 + 1];
    const int plane = R * R;
    for (int x = tid; x < R; x += PCL_MT) {
      float cur[C], nxt[C];
#pragma unroll
      for (int ch = 0; ch < C; ++ch) { cur[ch] = 0.f; nxt[ch] = 0.f; }
      int jc = jfirst;   // intermediate row the accumulators `cur` belong to (`nxt` belongs to jc+1)
      auto finish_row = [&](int j) {
        if (j >= j0 && j <= j1) {
#pragma unroll
          for (int ch = 0; ch < C; ++ch) Vb[((j - j0) * C + ch) * R + x] = cur[ch];
        }
#pragma unroll
        for (int ch = 0; ch < C; ++ch) { cur[ch] = nxt[ch]; nxt[ch] = 0.f; }
      };
      const float* gp = go + (size_t)ylo * R + x;
      for (int y = ylo; y < yhi; y += 4, gp += 4 * R) {
        float gv[4][C];
#pragma unroll
        for (int u = 0; u < 4; ++u)   // four rows in flight
#pragma unroll
          for (int ch = 0; ch < C; ++ch) gv[u][ch] = (y + u < yhi) ? __ldcs(gp + u * R + ch * plane) : 0.f;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int yy = y + u;
          if (yy < yhi) {
            while (yy >= start[jc + 1]) { finish_row(jc); ++jc; }   // run of jc finished (block-uniform)
            const float ly1 = tl1[yy];
            const bool last_row = jc >= s - 1;
            const float ly0 = last_row ? 1.0f : 1.0f - ly1;
#pragma unroll
            for (int ch = 0; ch < C; ++ch) {
              cur[ch] = fmaf(ly0, gv[u][ch], cur[ch]);
              if (!last_row) nxt[ch] = fmaf(ly1, gv[u][ch], nxt[ch]);
            }
          }
        }
      }
      while (jc <= j1) { finish_row(jc); ++jc; }   // flush (also covers a clamped last row whose own run is empty)
    }
    __syncthreads();
    const int nB = (j1 - j0 + 1) * s;
    const float inv_s = 1.0f / (float)s;
    for (int idx = tid; idx < nB; idx += PCL_MT) {
      const int jr = fast_div(idx, s, inv_s), i = idx - jr * s;
      const int j = j0 + jr;
      // window [wb, we) split at wa: d < wa weighs l1[d], else 1-l1[d]; the last column takes weight 1 on its own run
      const int wa = start[i], we = start[i + 1], wb = i > 0 ? start[i - 1] : wa;
      const bool last_col = i == s - 1;
      const float* vrow = Vb + (size_t)jr * C * R;
      float h[C];
#pragma unroll
      for (int ch = 0; ch < C; ++ch) h[ch] = 0.f;
      for (int d = wb; d < wa; ++d) {
        const float w = tl1[d];
#pragma unroll
        for (int ch = 0; ch < C; ++ch) h[ch] = fmaf(w, vrow[ch * R + d], h[ch]);
      }
      for (int d = wa; d < we; ++d) {
        const float w = last_col ? 1.0f : 1.0f - tl1[d];
#pragma unroll
        for (int ch = 0; ch < C; ++ch) h[ch] = fmaf(w, vrow[ch * R + d], h[ch]);
      }
      float v[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int ch = 0; ch < C; ++ch) v[ch] = h[ch];
      G[(size_t)j * s + i] = make_float4(v[0], v[1], v[2], v[3]);
    }
    return;
  }
  // generic path (s > R, very large s, or unaligned rows): direct 2-D gather per intermediate pixel
  const int nB = (j1 - j0 + 1) * s;
  for (int idx = tid; idx < nB; idx += PCL_MT) {
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
    G[(size_t)j * s + i] = make_float4(acc[0], acc[1], acc[2], acc[3]);
  }
}


// Transposed resize, vectorised form of the kernel above (needs R % 4 == 0 and a 16-byte aligned g_out; s <= R).
// Same band decomposition (one CTA per crop and PCL_JR intermediate rows), fewer instructions per element:
//   tables         : rowtab[d] = (weight on i0(d), weight on i0(d)+1, i0(d)) serves both axes;
//   vertical pass  : the band's rows are split over PCL_MG groups of two warps; a thread owns FOUR adjacent output
//                    columns (one 16-byte load per row and channel, two rows in flight) and walks down its group's
//                    output rows.  The partial sum a group leaves for the first row of the next group (the l1 taps of
//                    its last run) goes to a carry buffer and is folded in after the barrier, so no row is read twice
//                    inside a CTA (the band's own halo, 1/16 of the rows, still is);
//   horizontal pass: a thread owns one intermediate column of FOUR intermediate rows, so window bounds, weights and loop
//                    control are paid once per 4*C FMAs; it then writes the gradient.
constexpr int PCL_MG = 4;               // row groups (two warps each)
constexpr int PCL_M4T = 64 * PCL_MG;     // threads per CTA
constexpr int PCL_MU = 2;               // output rows in flight per thread in the vertical pass (3: 427 vs 377 us, spills)
constexpr int PCL_MB = 2;               // bands per CTA (measured per 1024 / 4096 images: 1 -> 380.9 / 1473.5 us, 2 -> 376.8 / 1442.5, 4 -> 388.8 / 1445.7)

template <int C, int RT>
__global__ void __launch_bounds__(PCL_M4T, 3) pcl_bwd_mid4_kernel(const float* __restrict__ g_out, const float* __restrict__ params,
                                                                 int q_base, int R_arg, float* __restrict__ ws) {
  const int R = RT ? RT : R_arg;
  extern __shared__ __align__(16) float sm[];
  const int q = q_base + blockIdx.y;
  const float* rec = params + (size_t)q * PF;
  const int jfirst = blockIdx.x * PCL_MB * PCL_JR;   // a CTA walks PCL_MB consecutive bands: the tables are built once, half as many CTAs exit empty
  {
    const int s_only = __float_as_int(__ldg(rec + 18));   // one load decides whether this CTA's first band exists
    if (jfirst >= s_only || s_only > (RT ? RT : R_arg)) return;
  }
  const Crop c = load_crop(rec);
  const int s = c.s;
  if (jfirst >= s || s > R) return;
  const int nx4 = R >> 2;
  float4* rowtab = reinterpret_cast<float4*>(sm);            // [R]
  int* start = reinterpret_cast<int*>(sm + 4 * R);           // [R+1] (padded to R+4)
  float* Vb = sm + 5 * R + 4;                                // [PCL_JR][C][R]
  float4* Vb4 = reinterpret_cast<float4*>(Vb);
  float4* Cb4 = reinterpret_cast<float4*>(Vb + PCL_JR * C * R);   // [PCL_MG-1][C][R/4]
  float* base = ws + __float_as_int(__ldg(rec + 21));
  float4* G = reinterpret_cast<float4*>(base);
  const float4* go4 = reinterpret_cast<const float4*>(g_out + (size_t)q * C * R * R);
  const int tid = threadIdx.x;
  // tables
  for (int m = tid; m <= s; m += PCL_M4T) start[m] = R;
  __syncthreads();
  for (int d = tid; d < R; d += PCL_M4T) {
    int i0, i1, p0 = -1, p1;
    float l0, l1, q0, q1;
    resize_coef(c, d, R, i0, i1, l0, l1);
    if (d > 0) resize_coef(c, d - 1, R, p0, p1, q0, q1);
    const bool last = i0 >= s - 1;   // the upper tap is clamped onto the same row/column: one tap of weight l0 + l1 = 1
    rowtab[d] = make_float4(last ? 1.0f : l0, last ? 0.0f : l1, __int_as_float(i0), 0.0f);
    for (int m = p0 + 1; m <= i0; ++m) start[m] = d;
  }
  __syncthreads();
  for (int j0 = jfirst; j0 < jfirst + PCL_MB * PCL_JR && j0 < s; j0 += PCL_JR) {
  const int j1 = min(j0 + PCL_JR, s) - 1;
  // vertical pass
  const int nrows = j1 - j0 + 1;
  const int rpg = (nrows + PCL_MG - 1) / PCL_MG;
  {
    const int g = tid >> 6, tx = tid & 63;   // PCL_M4T / PCL_MG == 64 threads per group
    const int jA = j0 + g * rpg, jB = min(jA + rpg, j1 + 1);
    const int plane4 = R * nx4;
    if (jA < jB) {
      for (int x4 = tx; x4 < nx4; x4 += 64) {
        float4 cur[C], nxt[C];
#pragma unroll
        for (int ch = 0; ch < C; ++ch) { cur[ch] = make_float4(0.f, 0.f, 0.f, 0.f); nxt[ch] = make_float4(0.f, 0.f, 0.f, 0.f); }
        int jc = (g == 0 && j0 > 0) ? j0 - 1 : jA;   // the band's halo: the run of row j0-1 feeds row j0 through its l1 taps
        auto finish_row = [&](int j) {
          if (j >= j0) {
#pragma unroll
            for (int ch = 0; ch < C; ++ch) Vb4[((j - j0) * C + ch) * nx4 + x4] = cur[ch];
          }
#pragma unroll
          for (int ch = 0; ch < C; ++ch) { cur[ch] = nxt[ch]; nxt[ch] = make_float4(0.f, 0.f, 0.f, 0.f); }
        };
        const int ylo = start[jc], yhi = start[jB];
        const float4* gp = go4 + (size_t)ylo * nx4 + x4;
        for (int y = ylo; y < yhi; y += PCL_MU, gp += PCL_MU * nx4) {
          float4 gv[PCL_MU][C];
#pragma unroll
          for (int u = 0; u < PCL_MU; ++u)
#pragma unroll
            for (int ch = 0; ch < C; ++ch) gv[u][ch] = (y + u < yhi) ? __ldcs(gp + u * nx4 + ch * plane4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int u = 0; u < PCL_MU; ++u) {
            if (y + u < yhi) {
              const float4 T = rowtab[y + u];
              const int i0 = __float_as_int(T.z);
              while (i0 > jc) { finish_row(jc); ++jc; }   // group-uniform (groups are whole warps)
#pragma unroll
              for (int ch = 0; ch < C; ++ch) {
                // packed f32x2 FMAs (SASS FFMA2): each lane is the same IEEE fmaf as the scalar form, half the issue slots
                const float2 w0 = make_float2(T.x, T.x), w1 = make_float2(T.y, T.y);
                const float2 glo = make_float2(gv[u][ch].x, gv[u][ch].y), ghi = make_float2(gv[u][ch].z, gv[u][ch].w);
                const float2 c0 = __ffma2_rn(w0, glo, make_float2(cur[ch].x, cur[ch].y)), c1 = __ffma2_rn(w0, ghi, make_float2(cur[ch].z, cur[ch].w));
                const float2 n0 = __ffma2_rn(w1, glo, make_float2(nxt[ch].x, nxt[ch].y)), n1 = __ffma2_rn(w1, ghi, make_float2(nxt[ch].z, nxt[ch].w));
                cur[ch] = make_float4(c0.x, c0.y, c1.x, c1.y);
                nxt[ch] = make_float4(n0.x, n0.y, n1.x, n1.y);
              }
            }
          }
        }
        while (jc < jB) { finish_row(jc); ++jc; }   // flush (also rows whose own run is empty)
        if (g < PCL_MG - 1 && jB <= j1) {           // what is left in `cur` belongs to the next group's first row
#pragma unroll
          for (int ch = 0; ch < C; ++ch) Cb4[(g * C + ch) * nx4 + x4] = cur[ch];
        }
      }
    }
  }
  __syncthreads();
  for (int idx = tid; idx < (PCL_MG - 1) * C * nx4; idx += PCL_M4T) {
    const int gg = idx / (C * nx4), rem = idx - gg * (C * nx4);
    const int row = (gg + 1) * rpg;
    if (row < nrows) {
      float4 a = Vb4[row * C * nx4 + rem];
      const float4 b = Cb4[idx];
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
      Vb4[row * C * nx4 + rem] = a;
    }
  }
  __syncthreads();
  // horizontal pass
  const int nq = (nrows + 3) >> 2;
  const unsigned magic_s = s > 1 ? 0xFFFFFFFFu / (unsigned)s + 1u : 0u;   // idx / s by multiply-high (idx < 4 s: exact)
  for (int idx = tid; idx < nq * s; idx += PCL_M4T) {
    const int rq = s > 1 ? (int)__umulhi((unsigned)idx, magic_s) : idx, i = idx - rq * s;
    const int wa = start[i], we = start[i + 1], wb = i > 0 ? start[i - 1] : wa;
    const float* v0 = Vb + (size_t)(rq * 4) * C * R;
    float h[4][C];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int ch = 0; ch < C; ++ch) h[r][ch] = 0.f;
    for (int d = wb; d < we; ++d) {
      const float2 T = *reinterpret_cast<const float2*>(rowtab + d);
      const float w = d < wa ? T.y : T.x;   // run(i-1) reaches column i through its l1 taps, run(i) through its l0 taps
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int ch = 0; ch < C; ++ch) h[r][ch] = fmaf(w, v0[(r * C + ch) * R + d], h[r][ch]);
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int j = j0 + rq * 4 + r;
      if (j <= j1) {
        float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int ch = 0; ch < C; ++ch) o[ch] = h[r][ch];
        G[(size_t)j * s + i] = make_float4(o[0], o[1], o[2], o[3]);
      }
    }
  }
  __syncthreads();   // the next band's vertical pass overwrites the band buffer
  }
}


constexpr int PCL_TS = 32;              // source tile side
constexpr int PCL_CELLS = PCL_TS + 1;   // cells = floor(sample position) in [tile-1, tile+31]
constexpr int PCL_K = 4;                // list capacity per cell (longer lists take the scan fallback)
constexpr int PCL_CNT_BYTES = (PCL_CELLS * PCL_CELLS * 4 + 15) & ~15;            // counters, padded: the lists stay 8-byte aligned
constexpr int PCL_LST_BYTES = (PCL_CELLS * PCL_CELLS * 2 * PCL_K + 15) & ~15;
constexpr int PCL_REG = 1536;           // region pixels staged in shared memory per (tile, crop) (39 x 39; 4 CTAs/SM)

// Transposed grid_sample, gather form.  One CTA per (image, 32x32 source tile) -- or per row of such tiles (WALK) --; for each
// tile and each crop of the image:
//   1. the tile's pre-image under the inverse homography bounds a region of the intermediate grid;
//   2. every intermediate pixel of the region is binned by floor(sample position) into per-cell lists in
//      shared memory (integer atomics claim the slots);
//   3. every source pixel reads the <= 4 cells whose bilinear footprint covers it and accumulates exactly its
//      contributors, in index order (so the result does not depend on the order the atomics ran in).
// g_img is written once per pixel: no float atomics, no memset.
constexpr int PCL_RECS = 4;            // crop records cached in shared memory per image

// WALK: blockIdx.x is a ROW of tiles and the CTA walks its tiles left to right (the image's records are fetched once per row
// instead of once per tile, tiles no crop touches cost no launch and no load latency of their own)
template <int C, int RT, bool LIST = false, int WALK = 0>   // WALK: tile rows per CTA (0: one CTA per tile)   // LIST: walk the images of a list (fall-back of the scatter form) instead of image blockIdx.y
__global__ void __launch_bounds__(PCL_THREADS, 4) pcl_bwd_img_kernel(const float* __restrict__ params, const float* __restrict__ ws,
                                                                  int img_base, int crops_per_img, int R_arg, float* __restrict__ g_img,
                                                                  const int* __restrict__ list) {
  // list == nullptr: blockIdx.y is the image inside the chunk; otherwise list[0] images list[1..] (the ones the scatter
  // kernel left) are walked by gridDim.y CTA rows
  const int R = RT ? RT : R_arg;
  extern __shared__ __align__(16) uint8_t img_sm[];
  float4* ent_g = reinterpret_cast<float4*>(img_sm);                                   // [PCL_REG] staged region: gradient ...
  float2* ent_p = reinterpret_cast<float2*>(img_sm + PCL_REG * 16);                    // [PCL_REG] ... and sample position
  int* cnt = reinterpret_cast<int*>(img_sm + PCL_REG * 24);                            // [cells]
  unsigned short* lst = reinterpret_cast<unsigned short*>(img_sm + PCL_REG * 24 + PCL_CNT_BYTES);  // [cells][K] region-local indices
  float* tu = reinterpret_cast<float*>(img_sm + PCL_REG * 24 + PCL_CNT_BYTES + PCL_LST_BYTES);             // [64] linspace of the region's columns
  float* tv = tu + 64;                                                                                              // [64] ... and rows
  __shared__ int overflow;
  __shared__ int box[4];
  const int tiles_x = (R + PCL_TS - 1) / PCL_TS;
  const int lx = threadIdx.x & 31, lyb = threadIdx.x >> 5;  // 32 x 8 threads, 4 adjacent rows each
  __shared__ float recs[PCL_RECS * PF];
  const int n_list = LIST ? __ldg(list) : 1;
  for (int le = LIST ? (int)blockIdx.y : 0; le < n_list; le += LIST ? (int)gridDim.y : 1) {
  const int im = LIST ? __ldg(list + 1 + le) : img_base + (int)blockIdx.y;
  if (LIST) __syncthreads();   // the previous image's readers are done with recs
  // the records of the image's (first PCL_RECS) crops are fetched once, all fields in flight together, instead of a chain
  // of dependent global loads per crop
  for (int e = threadIdx.x; e < min(crops_per_img, PCL_RECS) * PF; e += PCL_THREADS) recs[e] = __ldg(params + (size_t)im * crops_per_img * PF + e);
  __syncthreads();
  for (int tw = 0; tw < (WALK ? tiles_x * WALK : 1); ++tw) {
  const int trow = WALK ? (int)blockIdx.x * WALK + tw / tiles_x : (int)(blockIdx.x / tiles_x);
  if (WALK > 1 && trow * PCL_TS >= R) break;
  const int tx0 = WALK ? (tw % tiles_x) * PCL_TS : (int)(blockIdx.x % tiles_x) * PCL_TS, ty0 = trow * PCL_TS;
  float acc[4][C];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int ch = 0; ch < C; ++ch) acc[r][ch] = 0.f;
  for (int k = 0; k < crops_per_img; ++k) {
    if (k >= PCL_RECS) {   // more crops per image than cached records: fetch this one into slot 0 (block-uniform)
      __syncthreads();
      if (threadIdx.x < PF) recs[threadIdx.x] = __ldg(params + (size_t)(im * crops_per_img + k) * PF + threadIdx.x);
      __syncthreads();
    }
    const float* rec = recs + (k < PCL_RECS ? k : 0) * PF;   // always shared memory
    // cheap cull: the crop's footprint box (from the setup kernel) against this tile
    if (__float_as_int(*(rec + 22)) > tx0 + PCL_TS || __float_as_int(*(rec + 23)) < tx0 - 1 ||
        __float_as_int(*(rec + 24)) > ty0 + PCL_TS || __float_as_int(*(rec + 25)) < ty0 - 1) continue;
    const int s = __float_as_int(*(rec + 18));
    if (s > R) continue;   // outside the supported domain of the backward (see hb_pcl_bwd in the header)
    const float* base = ws + __float_as_int(*(rec + 21));
    const float4* G = reinterpret_cast<const float4*>(base);
    __syncthreads();  // previous crop's readers are done with cnt/lst/ent and the region box
    // 1. region of the intermediate grid whose samples can land in cells [tx0-1, tx0+31] x [ty0-1, ty0+31]:
    //    sample positions in [tx0-1, tx0+32) <=> grid-sample pixel coordinates in [tx0-0.5, tx0+32.5); the pre-image
    //    of that box under the inverse homography is a convex quad, bounded by the box of its four corners.  EVERY warp
    //    computes it (lanes 0..3 one corner each, then a shuffle reduction): the same values in all warps, no single-warp
    //    section, and the counters / linspace tables are set up in the same barrier interval.
    int ri0, ri1, rj0, rj1;
    {
      const int ln = threadIdx.x & 31;
      const float sm1 = (float)(s - 1);
      const float gx = (float)tx0 - 0.5f + 33.0f * (float)(ln & 1), gy = (float)ty0 - 0.5f + 33.0f * (float)((ln >> 1) & 1);
      const float U = rec[9] * gx + rec[10] * gy + rec[11];
      const float V = rec[12] * gx + rec[13] * gy + rec[14];
      const float Wd = rec[15] * gx + rec[16] * gy + rec[17];
      const bool okc = Wd > 1e-12f;
      const float iw = __frcp_rn(okc ? Wd : 1.0f);
      const float mi = fminf(fmaxf(U * iw * sm1, -8.0f), sm1 + 8.0f), mj = fminf(fmaxf(V * iw * sm1, -8.0f), sm1 + 8.0f);
      float ilo = mi, ihi = mi, jlo = mj, jhi = mj;
#pragma unroll
      for (int m = 1; m < 4; m <<= 1) {
        ilo = fminf(ilo, __shfl_xor_sync(0xffffffffu, ilo, m)); ihi = fmaxf(ihi, __shfl_xor_sync(0xffffffffu, ihi, m));
        jlo = fminf(jlo, __shfl_xor_sync(0xffffffffu, jlo, m)); jhi = fmaxf(jhi, __shfl_xor_sync(0xffffffffu, jhi, m));
      }
      const bool bad = (__ballot_sync(0xffffffffu, !okc) & 0xfu) != 0;
      ilo = __shfl_sync(0xffffffffu, ilo, 0); ihi = __shfl_sync(0xffffffffu, ihi, 0);
      jlo = __shfl_sync(0xffffffffu, jlo, 0); jhi = __shfl_sync(0xffffffffu, jhi, 0);
      ri0 = bad ? 0 : max(0, (int)floorf(ilo - 0.25f)); ri1 = bad ? s - 1 : min(s - 1, (int)ceilf(ihi + 0.25f));
      rj0 = bad ? 0 : max(0, (int)floorf(jlo - 0.25f)); rj1 = bad ? s - 1 : min(s - 1, (int)ceilf(jhi + 0.25f));
    }
    if (ri0 > ri1 || rj0 > rj1) continue;  // this crop does not touch the tile (block-uniform: every warp computed the same box)
    for (int idx = threadIdx.x; idx < (PCL_CELLS * PCL_CELLS + 3) / 4; idx += PCL_THREADS) reinterpret_cast<int4*>(cnt)[idx] = make_int4(0, 0, 0, 0);
    if (threadIdx.x == 0) overflow = 0;
    const int rw = ri1 - ri0 + 1, rh = rj1 - rj0 + 1;
    const Crop c = load_crop_any(rec);
    float P[9];
#pragma unroll
    for (int e = 0; e < 9; ++e) P[e] = c.P[e];
    if (rw <= 64 && rh <= 64) {
      if (threadIdx.x < rw) tu[threadIdx.x] = lin01(c, ri0 + (int)threadIdx.x);
      else if (threadIdx.x >= 64 && threadIdx.x < 64 + rh) tv[threadIdx.x - 64] = lin01(c, rj0 + (int)threadIdx.x - 64);
    }
    __syncthreads();
    // 2. binning
    {
      // idx / rw for idx < 4096, rw <= 64 by multiply-high: exact while idx * ((2^32 / rw + 1) * rw - 2^32) < 2^32
      const unsigned magic = rw > 1 ? 0xFFFFFFFFu / (unsigned)rw + 1u : 0u;   // (rw == 1: the quotient is idx itself)
      const float cx_lo = (float)(tx0 - 1), cy_lo = (float)(ty0 - 1);
      if (rw * rh > PCL_REG || rw > 64 || rh > 64) {
        if (threadIdx.x == 0) overflow = 1;   // region too large to stage (extreme foreshortening): scan fallback
      } else {
        // stage the region's gradients with cp.async (all of a thread's loads in flight at once, no register staging); they
        // are only waited for after the position / binning work below, which does not need them.  (One TMA bulk copy per
        // region row completing on an mbarrier measured slower, 530 vs 505 us: ~37 copies of ~600 B per tile and crop.)
#pragma unroll 2
        for (int idx = threadIdx.x; idx < rw * rh; idx += PCL_THREADS) {
          const int rr = rw > 1 ? (int)__umulhi((unsigned)idx, magic) : idx, cc = idx - rr * rw;
          const int gidx = (rj0 + rr) * s + (ri0 + cc);
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(ent_g + idx)), "l"(G + gidx) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        // sample positions are recomputed (they are not stored in the workspace) and binned by the thread that made them
#pragma unroll 2
        for (int idx = threadIdx.x; idx < rw * rh; idx += PCL_THREADS) {
          const int rr = rw > 1 ? (int)__umulhi((unsigned)idx, magic) : idx, cc = idx - rr * rw;
          const float u = tu[cc], v = tv[rr];
          const float X = fmaf(P[1], v, P[0] * u) + P[2];
          const float Y = fmaf(P[4], v, P[3] * u) + P[5];
          const float Z = fmaf(P[7], v, P[6] * u) + P[8];
          const float iz = __fdividef(1.0f, 1e-8f + Z);   // 1-ulp reciprocal: the gradient is continuous in the position
          const float2 p = make_float2(X * iz - 0.5f, Y * iz - 0.5f);
          ent_p[idx] = p;
          const float fx = floorf(p.x) - cx_lo, fy = floorf(p.y) - cy_lo;
          if (fx >= 0.0f && fx < (float)PCL_CELLS && fy >= 0.0f && fy < (float)PCL_CELLS) {
            const int cell = (int)fy * PCL_CELLS + (int)fx;
            const int slot = atomicAdd(&cnt[cell], 1);
            if (slot < PCL_K) lst[cell * PCL_K + slot] = (unsigned short)idx; else overflow = 1;
          }
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
      }
    }
    __syncthreads();
    const bool slow = overflow != 0;
    // 3. gather.  A thread owns four vertically adjacent pixels (rows 4*lyb .. 4*lyb+3 of column lx): they touch five
    //    rows of two cells; each cell's entries are read once and feed the pixel below (as its upper taps) and the
    //    pixel above (as its lower taps).  Fixed order: cell row, cell column, list position.
    {
      const int sx = tx0 + lx;
      const float fsx = (float)sx;
      const int ly0 = 4 * lyb;
      if (!slow) {
        if (sx < R) {
#pragma unroll
          for (int cr = 0; cr < 5; ++cr) {
            // counts and lists of the row's two cells are fetched together, before any of them is used (independent
            // shared-memory loads in flight instead of a dependent chain per cell)
            const int cell0 = (ly0 + cr) * PCL_CELLS + lx;   // floor(pos) == (sx-1+dx, ty0+ly0+cr-1)
            const int n2[2] = {cnt[cell0], cnt[cell0 + 1]};
            const uint2 pk2[2] = {*reinterpret_cast<const uint2*>(lst + cell0 * PCL_K), *reinterpret_cast<const uint2*>(lst + (cell0 + 1) * PCL_K)};
            // ... and so are the first entries of both lists (index clamped: an empty cell's slot holds stale bits)
            const int c0 = min((int)(pk2[0].x & 0xffffu), PCL_REG - 1), c1 = min((int)(pk2[1].x & 0xffffu), PCL_REG - 1);
            const float2 p2[2] = {ent_p[c0], ent_p[c1]};
            const float4 g2[2] = {ent_g[c0], ent_g[c1]};
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
              const int n = n2[dx];
              if (n == 0) continue;
              auto add = [&](const float2 p, const float4 g) {
                const float wx = 1.0f - fabsf(p.x - fsx);
                const float gv[4] = {g.x, g.y, g.z, g.w};
                if (cr < 4) {   // pixel row cr: floor(pos.y) == sy - 1
                  const float w = wx * (1.0f - fabsf(p.y - (float)(ty0 + ly0 + cr)));
#pragma unroll
                  for (int ch = 0; ch < C; ++ch) acc[cr][ch] = fmaf(w, gv[ch], acc[cr][ch]);
                }
                if (cr > 0) {   // pixel row cr-1: floor(pos.y) == sy
                  const float w = wx * (1.0f - fabsf(p.y - (float)(ty0 + ly0 + cr - 1)));
#pragma unroll
                  for (int ch = 0; ch < C; ++ch) acc[cr - 1][ch] = fmaf(w, gv[ch], acc[cr - 1][ch]);
                }
              };
              if (n == 1) { add(p2[dx], g2[dx]); continue; }   // the usual case
              // several entries: in index order (the slots were claimed by atomics in arbitrary order); never more than
              // PCL_K = 4 here
              const uint2 pk = pk2[dx];
              int ord[PCL_K] = {(int)(pk.x & 0xffffu), (int)(pk.x >> 16), (int)(pk.y & 0xffffu), (int)(pk.y >> 16)};
#pragma unroll
              for (int e = 1; e < PCL_K; ++e) ord[e] = e < n ? ord[e] : 0x7fffffff;
#define HB_CSWAP(a, b) { const int lo_ = min(ord[a], ord[b]), hi_ = max(ord[a], ord[b]); ord[a] = lo_; ord[b] = hi_; }
              HB_CSWAP(0, 1) HB_CSWAP(2, 3) HB_CSWAP(0, 2) HB_CSWAP(1, 3) HB_CSWAP(1, 2)
#undef HB_CSWAP
#pragma unroll
              for (int e = 0; e < PCL_K; ++e) {
                if (e >= n) break;
                add(ent_p[ord[e]], ent_g[ord[e]]);
              }
            }
          }
        }
      } else {
        // a cell list overflowed or the region did not fit (extreme foreshortening): scan the whole region
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int sy = ty0 + ly0 + r;
          if (sx >= R || sy >= R) continue;
          const float fsy = (float)sy;
          for (int j = rj0; j <= rj1; ++j)
            for (int i = ri0; i <= ri1; ++i) {
              float2 p;
              sample_pos_fast(c, j, i, (float)R, p.x, p.y);
              const float ax = fabsf(p.x - fsx), ay = fabsf(p.y - fsy);
              if (!(ax < 1.0f && ay < 1.0f)) continue;
              const float4 g = __ldg(G + (size_t)j * s + i);
              const float w = (1.0f - ax) * (1.0f - ay);
              const float gv[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
              for (int ch = 0; ch < C; ++ch) acc[r][ch] = fmaf(w, gv[ch], acc[r][ch]);
            }
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int sx = tx0 + lx, sy = ty0 + 4 * lyb + r;
    if (sx >= R || sy >= R) continue;
    float* op = g_img + (size_t)im * C * R * R + sy * R + sx;
#pragma unroll
    for (int ch = 0; ch < C; ++ch) __stcs(op + ch * R * R, acc[r][ch]);
  }
  }
  }
}


// ---- transposed grid_sample, scatter form ----------------------------------------------------------
// The crop's samples are about one source pixel apart (s = the box side), so binning them into per-pixel lists (the
// gather kernel above) spends ten warp instructions per sample on bookkeeping.  Here the samples are scattered instead,
// WITHOUT atomics and in a fixed order, into a rolling window of source rows in shared memory:
//   * one CTA per image; the window [base, base + 32) of source rows moves down the image once, every row of g_img is
//     written exactly once when no remaining intermediate row of any crop can reach it (rows no crop touches: zeros);
//   * per window position each crop of the image takes a band of its next (warps x NP) intermediate rows that fit;
//     a warp owns a whole intermediate row; lanes are consecutive samples of it, 31 new ones per strip (lane 0 repeats the
//     previous strip's last sample as a provider).  x grows by more than half a pixel per sample (checked by the setup
//     kernel, slot 26 of the record), so the left pixel columns of a strip's samples are distinct unless two neighbours
//     share one (ballot -> four sub-phases for that strip).  A sample's right-column contribution is handed to its
//     neighbour through shuffles when that neighbour's left column is the same pixel column in the same pixel row;
//     otherwise (pixel-row crossing, skipped column, row end) the lane adds it itself after a __syncwarp;
//   * rows NP apart touch disjoint pixels wherever they come within two columns of each other (setup kernel's bound), so
//     the warps run rows j, j+NP, j+2NP, ... concurrently and a block barrier separates the NP phases of a band;
//   * the gradients of a warp's next row (or next 124 samples of a long row) are loaded while it works on the current ones.
// Every pixel's sum has a fixed order (window position, crop, phase, row, strip, sub-phase): bit-reproducible and
// independent of the sharding.  Images with a crop the bounds reject are listed by pcl_fallback_list_kernel and taken by
// the gather kernel.
constexpr int PCL_SC_WARPS = 8;
constexpr int PCL_SC_THREADS = 32 * PCL_SC_WARPS;
constexpr int PCL_SC_PAD = 4;            // window columns left and right of the image (16-byte aligned image rows)
constexpr int PCL_SC_SPT = 4;            // strips (31 samples) per load group
constexpr int PCL_SC_MAXC = 8;           // most crops per image (more: gather kernel)

struct ScRow { float lo, hi; };

// y range of intermediate row j: along a row y is a Moebius function of u, so its extremes are at the two ends
__device__ __forceinline__ ScRow sc_row_y(const Crop& c, int j) {
  const float v = lin01(c, j);
  const float Y0 = fmaf(c.P[4], v, c.P[5]), Z0 = fmaf(c.P[7], v, c.P[8]) + 1e-8f;
  const float ya = __fdividef(Y0, Z0) - 0.5f, yb = __fdividef(Y0 + c.P[3], Z0 + c.P[6]) - 0.5f;
  ScRow r;
  r.lo = fminf(ya, yb); r.hi = fmaxf(ya, yb);
  return r;
}

template <int RT>
__global__ void __launch_bounds__(PCL_SC_THREADS, 2) pcl_bwd_scatter_kernel(const float* __restrict__ params, const float* __restrict__ ws,
                                                                          int img_base, int crops_per_img, float* __restrict__ g_img) {
  constexpr int R = RT, W = R + 2 * PCL_SC_PAD, PLANE = PCL_SC_H * W, HM = PCL_SC_H - 1, R4 = R / 4, RW = 3 * R4;
  static_assert(PCL_SC_THREADS > RW && PCL_SC_THREADS < 2 * RW, "incremental index split of the row loops");
  extern __shared__ __align__(16) float win[];   // [3][PCL_SC_H][W]
  __shared__ int s_j[PCL_SC_MAXC];               // next intermediate row of each crop
  const int im = img_base + blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const float* rec0 = params + (size_t)im * crops_per_img * PF;
  for (int k = 0; k < crops_per_img; ++k)
    if (__float_as_int(__ldg(rec0 + k * PF + 26)) == 0) return;   // the gather kernel takes this image
  for (int idx = tid; idx < 3 * PLANE / 4; idx += PCL_SC_THREADS) reinterpret_cast<float4*>(win)[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (tid < PCL_SC_MAXC) s_j[tid] = 0;
  __syncthreads();
  float* gout = g_img + (size_t)im * 3 * R * R;
  // image rows [r0, r1) are zero
  auto zero_rows = [&](int r0, int r1) {
    const int total = (r1 - r0) * RW;
    int rr = tid / RW, rem = tid - rr * RW;
    for (int idx = tid; idx < total; idx += PCL_SC_THREADS) {
      const int ch = rem >= 2 * R4 ? 2 : (rem >= R4 ? 1 : 0), x4 = rem - ch * R4;
      __stcs(reinterpret_cast<float4*>(gout + ((size_t)ch * R + r0 + rr) * R) + x4, make_float4(0.f, 0.f, 0.f, 0.f));
      rem += PCL_SC_THREADS - RW; rr += 1;
      if (rem >= RW) { rem -= RW; rr += 1; }
    }
  };
  // smallest pixel row an unfinished crop can still touch (R + 1: all crops done)
  // (lane k evaluates crop k; one warp reduction instead of a loop every thread repeats)
  auto next_base = [&]() {
    int nb = R + 1;
    if (lane < crops_per_img) {
      const Crop c = load_crop(rec0 + lane * PF);
      const int j = s_j[lane];
      if (j < c.s) nb = min(max((int)floorf(sc_row_y(c, j).lo - 0.05f), -1), R);
    }
    return __reduce_min_sync(0xffffffffu, nb);
  };
  int base = next_base();
  zero_rows(0, min(base, R));
  for (int step = 0; step < 8 * R && base <= R; ++step) {   // (the cap only guards against a hang; every step retires rows or takes a band)
    for (int k = 0; k < crops_per_img; ++k) {
      const float* rec = rec0 + k * PF;
      const Crop c = load_crop(rec);
      const int s = c.s, j = s_j[k];
      if (j >= s) continue;
      const int NP = __float_as_int(__ldg(rec + 26));
      // band: as many rows (one group of NP per warp) as the window holds
      // (lane l tries PCL_SC_WARPS - l warps; the first lane that fits wins)
      int nw = 0, nb = 0;
      {
        const int tw = PCL_SC_WARPS - (lane & (PCL_SC_WARPS - 1));
        const int n = min(tw * NP, s - j);
        const int top = min((int)floorf(sc_row_y(c, j + n - 1).hi + 0.05f) + 1, R);
        const unsigned fits = __ballot_sync(0xffffffffu, top - base <= HM) & ((1u << PCL_SC_WARPS) - 1u);
        if (fits) { nw = PCL_SC_WARPS - (__ffs(fits) - 1); nb = min(nw * NP, s - j); }
      }
      if (nb == 0) continue;   // block-uniform
      const float4* G = reinterpret_cast<const float4*>(ws + __float_as_int(__ldg(rec + 21)));
      const int row0 = j + wid * NP;
      const bool wact = wid < nw;
      // next band of this crop (after the other crops' bands): its gradients into L2 now
      {
        const char* nx = reinterpret_cast<const char*>(G + (size_t)(j + nb) * s);
        const int nbytes = min(PCL_SC_WARPS * NP, s - j - nb) * s * 16;
        for (int o = tid * 128; o < nbytes; o += PCL_SC_THREADS * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + o));
      }
      const int sm1 = s - 1, half = s >> 1;
      const float fsm1 = (float)sm1;
      // (indices are clamped instead of predicated: a lane outside the row reads a neighbour's value and never uses it)
      auto load_task = [&](int row, int t0, float4 (&g)[PCL_SC_SPT]) {
        const float4* Grow = G + (size_t)min(row, sm1) * s;
#pragma unroll
        for (int q = 0; q < PCL_SC_SPT; ++q) g[q] = __ldcs(Grow + min(max(t0 + 31 * q - 1 + lane, 0), sm1));
      };
      float4 gq[PCL_SC_SPT], gn[PCL_SC_SPT];
      load_task(row0, 0, gn);
      for (int p = 0; p < NP; ++p) {
        const int row = row0 + p;
        const bool ract = wact && row < j + nb;
        const float v = lin01(c, min(row, sm1));
        for (int t0 = 0; t0 < s; t0 += 31 * PCL_SC_SPT) {
#pragma unroll
          for (int q = 0; q < PCL_SC_SPT; ++q) gq[q] = gn[q];
          if (t0 + 31 * PCL_SC_SPT < s) load_task(row, t0 + 31 * PCL_SC_SPT, gn);
          else if (p + 1 < NP) load_task(row + 1, 0, gn);
          if (!ract) continue;   // warp-uniform
#pragma unroll
          for (int qp = 0; qp < PCL_SC_SPT; qp += 2) {
            if (t0 + 31 * qp >= s) break;   // warp-uniform
            // two strips at once (independent instruction chains): their samples' left columns are all distinct unless
            // two neighbours share one (then: one strip after the other, four sub-phases each)
            float ax[2], ay[2], gv[2][3];
            float* at[2];
            float* ab[2];
            int key[2];
            bool ok[2], dup[2], fwd_in[2], defer[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int i = t0 + 31 * (qp + h) - 1 + lane;
              const float fi = (float)i;
              const float u = (i < half) ? __fmul_rn(c.step, fi) : fmaf(-c.step, fsm1 - fi, 1.0f);   // lin01 (s >= 2 here)
              // same expression as the gather kernel's binning pass (the two forms then differ by summation order only)
              const float X = fmaf(c.P[1], v, c.P[0] * u) + c.P[2];
              const float Y = fmaf(c.P[4], v, c.P[3] * u) + c.P[5];
              const float Z = fmaf(c.P[7], v, c.P[6] * u) + c.P[8];
              const float iz = __fdividef(1.0f, 1e-8f + Z);
              const float x = X * iz - 0.5f, y = Y * iz - 0.5f;
              const float fxf = floorf(x), fyf = floorf(y);
              ax[h] = x - fxf; ay[h] = y - fyf;
              // floor in [-1, R-1] on both axes  <=>  |floor - (R/2 - 1)| <= R/2   (NaN: false)
              ok[h] = (unsigned)i < (unsigned)s && (h == 0 || t0 + 31 * (qp + 1) < s) &&
                      fabsf(fxf - (float)(R / 2 - 1)) <= (float)(R / 2) && fabsf(fyf - (float)(R / 2 - 1)) <= (float)(R / 2);
              const int fx = (int)fxf, fy = (int)fyf;
              key[h] = ok[h] ? ((fy + 1) << 10) + fx + 1 : -5;
              defer[h] = lane == 31 && i < sm1;   // the next strip's lane 0 handles this sample's right column
              at[h] = win + ((fy & HM) * W + fx + PCL_SC_PAD);
              ab[h] = win + (((fy + 1) & HM) * W + fx + PCL_SC_PAD);
              const float4 g = gq[qp + h];
              gv[h][0] = g.x; gv[h][1] = g.y; gv[h][2] = g.z;
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              int nkey = __shfl_up_sync(0xffffffffu, key[h], 1);
              if (lane == 0) nkey = -9;
              fwd_in[h] = ok[h] && nkey + 1 == key[h];                            // neighbour's right column is my left column, same pixel row
              dup[h] = ok[h] && nkey >= 0 && ((nkey ^ key[h]) & 1023) == 0;        // neighbour shares my left column
            }
            const unsigned m_dup = __ballot_sync(0xffffffffu, dup[0] || dup[1]);
            if (m_dup == 0) {
              bool needB[2];
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const unsigned m_fwd = __ballot_sync(0xffffffffu, fwd_in[h]);
                const float n_ax = __shfl_up_sync(0xffffffffu, ax[h], 1), n_ay = __shfl_up_sync(0xffffffffu, ay[h], 1);
                float ng[3];
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) ng[ch] = __shfl_up_sync(0xffffffffu, gv[h][ch], 1);
                const float nwx = fwd_in[h] ? n_ax : 0.0f;
                const float bx = 1.0f - ax[h], by = 1.0f - ay[h];
                const float wlt = bx * by, wlb = bx * ay[h], nt = nwx * (1.0f - n_ay), nbm = nwx * n_ay;
                if (ok[h] && lane > 0) {   // lane 0 of a later strip is a provider: its sample's left column was the previous strip's
#pragma unroll
                  for (int ch = 0; ch < 3; ++ch) {
                    at[h][ch * PLANE] = fmaf(wlt, gv[h][ch], fmaf(nt, ng[ch], at[h][ch * PLANE]));
                    ab[h][ch * PLANE] = fmaf(wlb, gv[h][ch], fmaf(nbm, ng[ch], ab[h][ch * PLANE]));
                  }
                }
                const bool sent = lane < 31 && ((m_fwd >> (lane + 1)) & 1u);   // my right column went to lane + 1
                needB[h] = ok[h] && !sent && !defer[h];
              }
              __syncwarp();
              if (__ballot_sync(0xffffffffu, needB[0] || needB[1])) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                  if (needB[h]) {
                    const float wrt = ax[h] * (1.0f - ay[h]), wrb = ax[h] * ay[h];
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch) {
                      at[h][ch * PLANE + 1] = fmaf(wrt, gv[h][ch], at[h][ch * PLANE + 1]);
                      ab[h][ch * PLANE + 1] = fmaf(wrb, gv[h][ch], ab[h][ch * PLANE + 1]);
                    }
                  }
                }
                __syncwarp();
              }
            } else {
              // two neighbours share a pixel column somewhere in these strips: strip by strip, first-of-a-column lanes,
              // then the others; left columns, then right columns (no hand-over)
              for (int h = 0; h < 2; ++h) {
                const float bx = 1.0f - ax[h], by = 1.0f - ay[h];
                for (int sub = 0; sub < 4; ++sub) {
                  const bool mine = ok[h] && (dup[h] == (sub >= 2));
                  if ((sub & 1) == 0) {
                    if (mine && lane > 0) {
                      const float wlt = bx * by, wlb = bx * ay[h];
#pragma unroll
                      for (int ch = 0; ch < 3; ++ch) {
                        at[h][ch * PLANE] = fmaf(wlt, gv[h][ch], at[h][ch * PLANE]);
                        ab[h][ch * PLANE] = fmaf(wlb, gv[h][ch], ab[h][ch * PLANE]);
                      }
                    }
                  } else if (mine && !defer[h]) {
                    const float wrt = ax[h] * by, wrb = ax[h] * ay[h];
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch) {
                      at[h][ch * PLANE + 1] = fmaf(wrt, gv[h][ch], at[h][ch * PLANE + 1]);
                      ab[h][ch * PLANE + 1] = fmaf(wrb, gv[h][ch], ab[h][ch * PLANE + 1]);
                    }
                  }
                  __syncwarp();
                }
              }
            }
          }
        }
        __syncthreads();
        if (p == 0 && tid == 0) s_j[k] = j + nb;   // every thread read s_j[k] before the barrier above; seen after the next one (NP >= 3)
      }
    }
    // rows below nbs are final: no remaining intermediate row of any crop reaches them (y grows from row to row)
    const int nbs = max(next_base(), base);
    {
      const int nrow = min(nbs, base + PCL_SC_H) - base;
      if (tid < RW) {   // thread = (channel, four columns), walking down the retired rows
        const int ch = tid / R4, x4 = tid - ch * R4;
        float* wcol = win + ch * PLANE + PCL_SC_PAD + 4 * x4;
        float* gcol = gout + (size_t)ch * R * R + 4 * x4;
#pragma unroll 4
        for (int rr = 0; rr < nrow; ++rr) {
          const int Y = base + rr;
          float4* wp = reinterpret_cast<float4*>(wcol + (Y & HM) * W);
          const float4 val = *wp;
          *wp = make_float4(0.f, 0.f, 0.f, 0.f);
          if (Y >= 0 && Y < R) __stcs(reinterpret_cast<float4*>(gcol + Y * R), val);
        }
      }
      if (nbs > base + PCL_SC_H) zero_rows(min(base + PCL_SC_H, R), min(nbs, R));
    }
    __syncthreads();
    base = nbs;
  }
}

// images of a chunk the scatter kernel leaves to the gather kernel (a crop with record slot 26 == 0), in image order
__global__ void __launch_bounds__(1024) pcl_fallback_list_kernel(const float* __restrict__ params, int img_base, int n_imgs, int crops_per_img,
                                                                int* __restrict__ list /* [0] = count, [1..] = image indices */) {
  __shared__ int warp_tot[32];
  __shared__ int carry;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int b0 = 0; b0 < n_imgs; b0 += 1024) {
    const int e = b0 + threadIdx.x;
    int flag = 0;
    if (e < n_imgs)
      for (int k = 0; k < crops_per_img; ++k) flag |= __float_as_int(__ldg(params + (size_t)((img_base + e) * crops_per_img + k) * PF + 26)) == 0;
    const unsigned m = __ballot_sync(0xffffffffu, flag);
    if (lane == 0) warp_tot[wid] = __popc(m);
    __syncthreads();
    int before = carry;
    for (int w = 0; w < wid; ++w) before += warp_tot[w];
    if (flag) list[1 + before + __popc(m & ((1u << lane) - 1u))] = img_base + e;
    __syncthreads();
    if (threadIdx.x == 1023) carry = before + __popc(m);
    __syncthreads();
  }
  if (threadIdx.x == 0) list[0] = carry;
}

}  // namespace hb

using namespace hb;

static int g_pcl_scatter = -1;   // -1: not decided yet (env HB_PCL_SCATTER, default 0: measured ~530 vs 361 us per 1024 images)
static int pcl_scatter() {
  if (g_pcl_scatter < 0) { const char* e = getenv("HB_PCL_SCATTER"); g_pcl_scatter = (e && e[0] == '1') ? 1 : 0; }
  return g_pcl_scatter;
}
static int g_pcl_exact = -1;   // -1: not decided yet (env HB_PCL_EXACT, default 0 = fast forward)
static int pcl_exact() {
  if (g_pcl_exact < 0) { const char* e = getenv("HB_PCL_EXACT"); g_pcl_exact = (e && e[0] == '1') ? 1 : 0; }
  return g_pcl_exact;
}
extern "C" int hb_pcl_set_scatter(int on) { const int prev = pcl_scatter(); g_pcl_scatter = on ? 1 : 0; return prev; }
extern "C" int hb_pcl_set_exact(int on) { const int prev = pcl_exact(); g_pcl_exact = on ? 1 : 0; return prev; }

template <int C, typename SrcT>
static int launch_fwd(const SrcT* img, const float* params, int n_crops, int crops_per_img, int R, float* out, const PclNorm& nrm, cudaStream_t st) {
  constexpr bool U8 = sizeof(SrcT) == 1;
  // shared-memory budget per CTA: 44 KB -> 5 CTAs/SM (measured on B200: 3.31 ms vs 4.24 ms at 72 KB / 3 CTAs per SM;
  // the kernel is issue/latency bound, occupancy pays).  One intermediate row must fit: R*16 bytes * 4 rows.  The 8-bit
  // variant keeps its normalisation table (C KB) in static shared memory on top.
  size_t smem = (pcl_exact() && !U8 ? 44 : (U8 ? 41 - C : 41)) * 1024;   // the fast kernel keeps ~3 KB of tables in static shared memory
  { const char* e = getenv("HB_PCL_FWD_SMEM_KB"); if (e) smem = (size_t)atoi(e) * 1024; }   // experiment knob
  if (smem < (size_t)R * 16 * 4) smem = (size_t)R * 16 * 4;
  if (smem > 200 * 1024) { set_error("hb_pcl_fwd: img_res too large for the staged kernel"); return HB_E_UNSUPPORTED; }
  // bulk copies need 16-byte aligned row segments: R % 4 == 0 (R % 16 for 8-bit images) and a 16-byte aligned image
  static int want_tma = -1;
  if (want_tma < 0) { const char* e = getenv("HB_PCL_TMA"); want_tma = (e && e[0] == '0') ? 0 : 1; }
  const int tma_ok = want_tma && (R % (U8 ? 16 : 4) == 0) && ((reinterpret_cast<uintptr_t>(img) & 15u) == 0);
  const int nxb = (R == 224) ? 224 / PCL_XB : (R > 128 ? 2 : 1);   // column blocks of the fast kernel
  const unsigned cpi_mul = (unsigned)((0x100000000ull + (unsigned)crops_per_img - 1) / (unsigned)crops_per_img);   // q / cpi == umulhi(q, cpi_mul) for q * cpi < 2^32
  const bool v2 = (R % 2 == 0) && ((reinterpret_cast<uintptr_t>(out) & 7u) == 0);
  const size_t plane = (size_t)C * R * R;
  // crops sit on grid.y (limit 65,535): larger batches go out as several launches, split at a multiple of crops_per_img
  const int max_chunk = (65535 / crops_per_img) * crops_per_img;
  if (max_chunk <= 0) { set_error("hb_pcl_fwd: crops_per_img too large"); return HB_E_ARG; }
  for (int base = 0; base < n_crops; base += max_chunk) {
    const int nc = n_crops - base < max_chunk ? n_crops - base : max_chunk;
    const SrcT* img_c = img + (size_t)(base / crops_per_img) * plane;
    const float* par_c = params + (size_t)base * PF;
    float* out_c = out + (size_t)base * plane;
    if constexpr (!U8) {
      if (pcl_exact()) {
        auto fwd_kernel = (R == 224) ? pcl_fwd_kernel<C, 224> : pcl_fwd_kernel<C, 0>;
        HB_CUDA(cudaFuncSetAttribute(fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        fwd_kernel<<<dim3((R + PCL_TR - 1) / PCL_TR, nc), PCL_THREADS, smem, st>>>(img_c, par_c, crops_per_img, R, out_c, PCL_TR + 2, (int)smem, tma_ok);
        g_launches++;
        const int rc = check_launch("pcl_fwd_kernel");
        if (rc) return rc;
        continue;
      }
    }
    auto fast_kernel = (R == 224 && v2) ? pcl_fwd_fast_kernel<C, 224, SrcT, true> : (v2 ? pcl_fwd_fast_kernel<C, 0, SrcT, true> : pcl_fwd_fast_kernel<C, 0, SrcT, false>);
    HB_CUDA(cudaFuncSetAttribute(fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    fast_kernel<<<dim3(PCL_NSEG * nxb, nc), PCL_THREADS, smem, st>>>(img_c, par_c, crops_per_img, crops_per_img == 1 ? 0u : cpi_mul, R, out_c, (int)smem, tma_ok, nxb, nrm);
    g_launches++;
    const int rc = check_launch("pcl_fwd_fast_kernel");
    if (rc) return rc;
  }
  return 0;
}

template <typename SrcT>
static int pcl_fwd_dispatch(const SrcT* img, const float* params, int n_crops, int crops_per_img, int C, int img_res, float* out, const PclNorm& nrm, void* stream) {
  if (n_crops < 0 || crops_per_img <= 0 || C <= 0 || C > 4 || img_res <= 0 || (n_crops > 0 && (!img || !params || !out)) || n_crops % crops_per_img) {
    set_error("hb_pcl_fwd: bad argument (1 <= C <= 4)"); return HB_E_ARG;
  }
  if (n_crops == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  switch (C) {
    case 1: return launch_fwd<1, SrcT>(img, params, n_crops, crops_per_img, img_res, out, nrm, st);
    case 2: return launch_fwd<2, SrcT>(img, params, n_crops, crops_per_img, img_res, out, nrm, st);
    case 3: return launch_fwd<3, SrcT>(img, params, n_crops, crops_per_img, img_res, out, nrm, st);
    default: return launch_fwd<4, SrcT>(img, params, n_crops, crops_per_img, img_res, out, nrm, st);
  }
}

extern "C" int hb_pcl_fwd(const float* img, const float* params, int n_crops, int crops_per_img, int C, int img_res, float* out, void* stream) {
  PclNorm nrm{};
  return pcl_fwd_dispatch<float>(img, params, n_crops, crops_per_img, C, img_res, out, nrm, stream);
}

extern "C" int hb_pcl_fwd_u8(const uint8_t* img, const float* mean_host, const float* std_host, const float* params, int n_crops, int crops_per_img,
                             int C, int img_res, float* out, void* stream) {
  if (!mean_host || !std_host || C <= 0 || C > 4) { set_error("hb_pcl_fwd_u8: bad argument (mean/std are C host floats, 1 <= C <= 4)"); return HB_E_ARG; }
  PclNorm nrm{};
  for (int ch = 0; ch < C; ++ch) {
    nrm.mean[ch] = mean_host[ch];
    nrm.std[ch] = std_host[ch];
    if (!(std_host[ch] != 0.0f)) { set_error("hb_pcl_fwd_u8: std[%d] must be non-zero", ch); return HB_E_ARG; }
  }
  return pcl_fwd_dispatch<uint8_t>(img, params, n_crops, crops_per_img, C, img_res, out, nrm, stream);
}

// The backward runs in chunks of images: per chunk, pcl_bwd_mid fills the workspace (intermediate gradient +
// sample positions of the chunk's crops, packed) and pcl_bwd_img consumes it.  The chunk size follows from the
// workspace the caller provides: any size >= one image's worth works; hb_pcl_bwd_workspace_bytes() recommends
// kPclChunkImgs images per chunk (large launches: fewer kernel tails; the packed region actually touched is ~27% of the capacity
// for boxes of side U{56..168}).
static const int kPclChunkImgs = 4096;   // measured per 8192-image step: 512 -> 6.28 ms, 1024 -> 6.11, 2048 -> 6.05, 4096 -> 6.01, 8192 -> 6.00 (1.6 MB of workspace per image)

static size_t pcl_ws_g_bytes_per_img(int crops_per_img, int img_res) {
  return sizeof(float) * (size_t)crops_per_img * (((size_t)PCL_WS_FLOATS_PER_PX * img_res * img_res + 3) & ~(size_t)3);
}
// + 16 bytes per image behind the chunk's gradients: the list of images the scatter kernel leaves to the gather kernel
static size_t pcl_ws_bytes_per_img(int crops_per_img, int img_res) { return pcl_ws_g_bytes_per_img(crops_per_img, img_res) + 16; }

extern "C" size_t hb_pcl_bwd_workspace_bytes(int n_crops, int crops_per_img, int C, int img_res) {
  (void)C;
  if (n_crops <= 0 || crops_per_img <= 0) return 0;
  const int n_imgs = n_crops / crops_per_img;
  int chunk = kPclChunkImgs;
  { const char* e = getenv("HB_PCL_CHUNK_IMGS"); if (e && atoi(e) > 0) chunk = atoi(e); }   // experiment knob
  return (size_t)(n_imgs < chunk ? n_imgs : chunk) * pcl_ws_bytes_per_img(crops_per_img, img_res);
}

template <int C>
static int launch_bwd(const float* g_out, const float* params, int n_crops, int crops_per_img, int R, float* g_img, float* ws, size_t ws_bytes, int stages, cudaStream_t st) {
  const int n_imgs = n_crops / crops_per_img;
  const size_t fit = ws_bytes / pcl_ws_bytes_per_img(crops_per_img, R);
  const int chunk_imgs = (size_t)n_imgs < fit ? n_imgs : (int)fit;
  const int chunk_crops = chunk_imgs * crops_per_img;
  const int n_chunks = (n_imgs + chunk_imgs - 1) / chunk_imgs;
  int rc = 0;
  if (stages & 1) {
    pcl_offsets_kernel<<<n_chunks, 1024, 0, st>>>(const_cast<float*>(params), n_crops, chunk_crops, R);
    g_launches++;
    rc = check_launch("pcl_offsets_kernel");
    if (rc) return rc;
  }
  const size_t smem_mid = sizeof(float) * ((size_t)2 * R + 4 + (size_t)PCL_JR * C * R);
  // bulk copies need 16-byte aligned rows: R % 4 == 0 and a 16-byte aligned g_out
  const int use_tma = (R % 4 == 0) && ((reinterpret_cast<uintptr_t>(g_out) & 15u) == 0);
  auto mid_kernel = (R == 224) ? pcl_bwd_mid_kernel<C, 224> : pcl_bwd_mid_kernel<C, 0>;
  HB_CUDA(cudaFuncSetAttribute(mid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_mid));
  // vectorised variant: 16-byte row segments
  const size_t smem_mid4 = sizeof(float) * ((size_t)5 * R + 4 + (size_t)(PCL_JR + PCL_MG - 1) * C * R);
  static int want_mid4 = -1;
  if (want_mid4 < 0) { const char* e = getenv("HB_PCL_MID4"); want_mid4 = (e && e[0] == '0') ? 0 : 1; }
  const bool mid4 = want_mid4 && use_tma && smem_mid4 <= 200 * 1024;
  auto mid4_kernel = (R == 224) ? pcl_bwd_mid4_kernel<C, 224> : pcl_bwd_mid4_kernel<C, 0>;
  if (mid4) HB_CUDA(cudaFuncSetAttribute(mid4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_mid4));
  // scatter form of the transposed grid_sample (C == 3, R == 224); HB_PCL_SCATTER=0 keeps the gather kernel for every image
  const bool scatter = pcl_scatter() && C == 3 && R == 224 && crops_per_img <= PCL_SC_MAXC;
  const size_t smem_sc = sizeof(float) * 3 * PCL_SC_H * (size_t)(R + 2 * PCL_SC_PAD);
  if (scatter) HB_CUDA(cudaFuncSetAttribute(pcl_bwd_scatter_kernel<224>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_sc));
  if (scatter) HB_CUDA(cudaFuncSetAttribute(pcl_bwd_img_kernel<C, 224, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(PCL_REG * 24 + PCL_CNT_BYTES + PCL_LST_BYTES + 128 * sizeof(float))));
  int* fb_list = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(ws) + (size_t)chunk_imgs * pcl_ws_g_bytes_per_img(crops_per_img, R));
  const int tiles = ((R + PCL_TS - 1) / PCL_TS) * ((R + PCL_TS - 1) / PCL_TS);
  const size_t smem_img = (size_t)PCL_REG * 24 + PCL_CNT_BYTES + PCL_LST_BYTES + 128 * sizeof(float);
  auto img_kernel = (R == 224) ? pcl_bwd_img_kernel<C, 224> : pcl_bwd_img_kernel<C, 0>;
  HB_CUDA(cudaFuncSetAttribute(img_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_img));
  for (int ch = 0; ch < n_chunks; ++ch) {
    const int im0 = ch * chunk_imgs;
    const int nim = (n_imgs - im0) < chunk_imgs ? (n_imgs - im0) : chunk_imgs;
    dim3 g1((R + PCL_JR - 1) / PCL_JR, nim * crops_per_img);
    if (stages & 1) {
      if (mid4) mid4_kernel<<<dim3((g1.x + PCL_MB - 1) / PCL_MB, g1.y), PCL_M4T, smem_mid4, st>>>(g_out, params, im0 * crops_per_img, R, ws);
      else mid_kernel<<<g1, PCL_MT, smem_mid, st>>>(g_out, params, im0 * crops_per_img, R, ws, use_tma);
      g_launches++;
      rc = check_launch("pcl_bwd_mid_kernel");
      if (rc) return rc;
    }
    dim3 g2(tiles, nim);
    if ((stages & 2) && scatter) {
      pcl_fallback_list_kernel<<<1, 1024, 0, st>>>(params, im0, nim, crops_per_img, fb_list);
      pcl_bwd_scatter_kernel<224><<<nim, PCL_SC_THREADS, smem_sc, st>>>(params, ws, im0, crops_per_img, g_img);
      dim3 g3(tiles, nim < 8 ? nim : 8);   // the listed images (usually none) walked by 8 CTA rows
      pcl_bwd_img_kernel<C, 224, true><<<g3, PCL_THREADS, smem_img, st>>>(params, ws, im0, crops_per_img, R, g_img, fb_list);
      g_launches += 3;
      rc = check_launch("pcl_bwd_scatter_kernel");
      if (rc) return rc;
    } else if (stages & 2) {
      static int walk = -1;   // HB_PCL_IMG_WALK=0: one CTA per tile instead of one per row of tiles
      if (walk < 0) { const char* e = getenv("HB_PCL_IMG_WALK"); walk = (e && e[0] == '0') ? 0 : 1; }
      if (walk && R == 224 && crops_per_img <= PCL_RECS) {   // measured per 1024 / 4096 images: one CTA per tile 388 / 1513 us, per tile row 361 / 1352, per two rows 406 / 1456, per image 489 / 1491
        HB_CUDA(cudaFuncSetAttribute(pcl_bwd_img_kernel<C, 224, false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_img));
        pcl_bwd_img_kernel<C, 224, false, 1><<<dim3((R + PCL_TS - 1) / PCL_TS, nim), PCL_THREADS, smem_img, st>>>(params, ws, im0, crops_per_img, R, g_img, nullptr);
      } else {
        img_kernel<<<g2, PCL_THREADS, smem_img, st>>>(params, ws, im0, crops_per_img, R, g_img, nullptr);
      }
      g_launches++;
      rc = check_launch("pcl_bwd_img_kernel");
      if (rc) return rc;
    }
  }
  return 0;
}

extern "C" int hb_pcl_bwd_stages(const float* g_out, const float* params, int n_crops, int crops_per_img, int C, int img_res, float* g_img,
                                 void* workspace, size_t workspace_bytes, int stages, void* stream) {
  if (n_crops < 0 || crops_per_img <= 0 || C <= 0 || C > 4 || img_res <= 0 || (n_crops > 0 && (!g_out || !params || !g_img || !workspace)) ||
      n_crops % crops_per_img || (stages & 3) == 0) {
    set_error("hb_pcl_bwd: bad argument (1 <= C <= 4)"); return HB_E_ARG;
  }
  if (n_crops == 0) return 0;
  if (workspace_bytes < pcl_ws_bytes_per_img(crops_per_img, img_res)) { set_error("hb_pcl_bwd: workspace too small (need at least one image's worth: %zu bytes)", pcl_ws_bytes_per_img(crops_per_img, img_res)); return HB_E_WORKSPACE; }
  if (reinterpret_cast<uintptr_t>(workspace) & 15u) { set_error("hb_pcl_bwd: workspace must be 16-byte aligned"); return HB_E_ALIGN; }
  cudaStream_t st = (cudaStream_t)stream;
  float* ws = (float*)workspace;
  switch (C) {
    case 1: return launch_bwd<1>(g_out, params, n_crops, crops_per_img, img_res, g_img, ws, workspace_bytes, stages, st);
    case 2: return launch_bwd<2>(g_out, params, n_crops, crops_per_img, img_res, g_img, ws, workspace_bytes, stages, st);
    case 3: return launch_bwd<3>(g_out, params, n_crops, crops_per_img, img_res, g_img, ws, workspace_bytes, stages, st);
    default: return launch_bwd<4>(g_out, params, n_crops, crops_per_img, img_res, g_img, ws, workspace_bytes, stages, st);
  }
}

extern "C" int hb_pcl_bwd(const float* g_out, const float* params, int n_crops, int crops_per_img, int C, int img_res, float* g_img,
                          void* workspace, size_t workspace_bytes, void* stream) {
  return hb_pcl_bwd_stages(g_out, params, n_crops, crops_per_img, C, img_res, g_img, workspace, workspace_bytes, 3, stream);
}
