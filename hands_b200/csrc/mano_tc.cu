// Blendshape contraction on the 5th-generation tensor cores (tcgen05 + TMEM), operands fed by TMA bulk copies.
//
//   v_posed[h][c] = v_template[c] + sum_p Pext[p][c] * F[h][p]        Pext = [posedirs; shapedirs^T]  (145 x 2334)
//
// GEMM view per CTA: D[M = 128 hands][N = 240 vertex-coordinate columns] += A[128 x 8] * B[240 x 8]^T over 19 k-steps,
// `tcgen05.mma.cta_group::1.kind::tf32`, accumulators in TMEM (two 240-column buffers so the epilogue of tile t
// overlaps the MMAs of tile t+1).  Single-pass TF32 (2^-11) would break the 1e-5 vertex tolerance, so every product is
// error-compensated: A = Ah + Al, B = Bh + Bl (each part exactly representable in TF32),
//   A*B ~= Ah*Bh + Ah*Bl + Al*Bh     (three MMAs per k-step; the dropped Al*Bl term is 2^-22 relative).
//
// Operand layout in shared memory: the canonical K-major, no-swizzle UMMA layout -- 8-row x 16-byte core matrices,
// core matrices of one k-step ("slab") stored [k-half][row-group][row][4 floats]; descriptor LBO = bytes between the
// two k-halves, SBO = 128 B between row groups.  The constants are pre-arranged in that order in global memory at
// hb_mano_create() and the pose kernel writes the feature rows in it, so every operand transfer is one contiguous
// `cp.async.bulk` (SASS UBLKCP) completing on an mbarrier.
//
// Warp roles (192 threads): warp 0 = TMA producer + TMEM allocator, warp 1 = MMA issuer (one elected lane),
// warps 2-5 = epilogue (tcgen05.ld -> + v_template -> 128-bit stores; thread = hand, i.e. TMEM lane).
#include "hb_common.cuh"
#include "tma.cuh"

namespace hb {

constexpr int TC_M = 128;          // hands per CTA
constexpr int TC_N = 240;          // columns per tile (10 tiles cover 3 x 800)
constexpr int TC_TILES = 10;
constexpr int TC_KSTEPS = 19;      // 152 / 8
constexpr int TC_A_SLAB = TC_M * 8;        // floats per k-step slab of A (4096 B)
constexpr int TC_B_SLAB = TC_N * 8;        // floats per k-step slab of B (7680 B)
constexpr int TC_KC = 2;           // k-steps per B stage
constexpr int TC_NSTG = 2;
constexpr int TC_THREADS = 192;
constexpr int TC_VP = 3 * VP;      // 2400 floats of v_posed per hand, coordinate-major [k][800]

__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  // cute::UMMA::SmemDescriptor: start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) | layout NONE [61,64)
  return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | ((uint64_t)1 << 46);
}
// cute::UMMA::InstrDescriptor: c=F32 (1<<4) | a=TF32 (2<<7) | b=TF32 (2<<10) | K-major both | N>>3 at [17,23) | M>>4 at [24,29)
constexpr uint32_t TC_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);

__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(TC_IDESC), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int k = 0; k < 16; ++k) v[k] = __uint_as_float(r[k]);
}

struct TcSmem {
  // offsets in bytes inside dynamic shared memory
  static constexpr int A_HI = 0;
  static constexpr int A_LO = A_HI + TC_KSTEPS * TC_A_SLAB * 4;
  static constexpr int B0 = A_LO + TC_KSTEPS * TC_A_SLAB * 4;                 // [NSTG][hi KC slabs | lo KC slabs]
  static constexpr int B_STAGE = 2 * TC_KC * TC_B_SLAB * 4;
  static constexpr int BARS = B0 + TC_NSTG * B_STAGE;                         // a_full, b_full[2], b_empty[2], acc_full[2], acc_empty[2]
  static constexpr int TMEM_PTR = BARS + 16 * 8;
  static constexpr int TOTAL = TMEM_PTR + 16;
};

// Fhi/Flo: [groups][19][4096 B] feature slabs (written by the pose kernel); Bhi/Blo: [10][19][7680 B] constant slabs.
__global__ void __launch_bounds__(TC_THREADS, 1) mano_blend_tc_kernel(const float* __restrict__ Fhi, const float* __restrict__ Flo,
                                                                       const float* __restrict__ Bhi, const float* __restrict__ Blo,
                                                                       const float* __restrict__ vt, int B, int nsplit,
                                                                       float* __restrict__ vp) {
  extern __shared__ __align__(128) uint8_t smem[];
  float* a_hi = reinterpret_cast<float*>(smem + TcSmem::A_HI);
  float* a_lo = reinterpret_cast<float*>(smem + TcSmem::A_LO);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TcSmem::BARS);
  uint64_t* a_full = bars;
  uint64_t* b_full = bars + 1;
  uint64_t* b_empty = bars + 3;
  uint64_t* acc_full = bars + 5;
  uint64_t* acc_empty = bars + 7;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + TcSmem::TMEM_PTR);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = blockIdx.x, split = blockIdx.y;
  const int ntiles = (TC_TILES - split + nsplit - 1) / nsplit;   // tiles split, split+nsplit, ...

  if (threadIdx.x == 0) {
    mbar_init(a_full, 1);
    for (int s = 0; s < TC_NSTG; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 4); }
    mbar_fence_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {
      // ---- TMA producer ----
      const uint32_t a_bytes = TC_KSTEPS * TC_A_SLAB * 4;
      mbar_arrive_expect_tx(a_full, 2 * a_bytes);
      bulk_g2s(a_hi, Fhi + (size_t)g * TC_KSTEPS * TC_A_SLAB, a_bytes, a_full);
      bulk_g2s(a_lo, Flo + (size_t)g * TC_KSTEPS * TC_A_SLAB, a_bytes, a_full);
      int fill = 0;
      for (int it = 0; it < ntiles; ++it) {
        const int tile = split + it * nsplit;
        for (int ks = 0; ks < TC_KSTEPS; ks += TC_KC, ++fill) {
          const int s = fill % TC_NSTG, n = fill / TC_NSTG;
          mbar_wait(&b_empty[s], (n & 1) ^ 1);
          const int nk = min(TC_KC, TC_KSTEPS - ks);
          const uint32_t bytes = nk * TC_B_SLAB * 4;
          float* dst = reinterpret_cast<float*>(smem + TcSmem::B0 + s * TcSmem::B_STAGE);
          mbar_arrive_expect_tx(&b_full[s], 2 * bytes);
          const size_t src = ((size_t)tile * TC_KSTEPS + ks) * TC_B_SLAB;
          bulk_g2s(dst, Bhi + src, bytes, &b_full[s]);
          bulk_g2s(dst + TC_KC * TC_B_SLAB, Blo + src, bytes, &b_full[s]);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---- MMA issuer ----
      mbar_wait(a_full, 0);
      tc_fence_after();
      const uint32_t a_hi_addr = smem_u32(a_hi), a_lo_addr = smem_u32(a_lo);
      int fill = 0;
      for (int it = 0; it < ntiles; ++it) {
        const int buf = it & 1, m = it >> 1;
        mbar_wait(&acc_empty[buf], (m & 1) ^ 1);
        tc_fence_after();
        const uint32_t d = tmem_base + buf * 256;
        uint32_t acc = 0;
        for (int ks = 0; ks < TC_KSTEPS; ks += TC_KC, ++fill) {
          const int s = fill % TC_NSTG, n = fill / TC_NSTG;
          mbar_wait(&b_full[s], n & 1);
          tc_fence_after();
          const uint32_t b_hi_addr = smem_u32(smem + TcSmem::B0 + s * TcSmem::B_STAGE);
          const uint32_t b_lo_addr = b_hi_addr + TC_KC * TC_B_SLAB * 4;
          const int nk = min(TC_KC, TC_KSTEPS - ks);
          for (int k = 0; k < nk; ++k) {
            const uint64_t ah = umma_desc(a_hi_addr + (ks + k) * TC_A_SLAB * 4, TC_M * 16, 128);
            const uint64_t al = umma_desc(a_lo_addr + (ks + k) * TC_A_SLAB * 4, TC_M * 16, 128);
            const uint64_t bh = umma_desc(b_hi_addr + k * TC_B_SLAB * 4, TC_N * 16, 128);
            const uint64_t bl = umma_desc(b_lo_addr + k * TC_B_SLAB * 4, TC_N * 16, 128);
            umma_tf32(d, al, bh, acc);   // small terms first
            umma_tf32(d, ah, bl, 1);
            umma_tf32(d, ah, bh, 1);
            acc = 1;
          }
          umma_commit(&b_empty[s]);      // frees the stage when these MMAs have read it
        }
        umma_commit(&acc_full[buf]);     // accumulator complete
      }
    }
  } else {
    // ---- epilogue: warps 2..5, TMEM lane quarter = warp % 4 ----
    const int q = warp & 3;
    const int h = q * 32 + lane;
    const int b = g * TC_M + h;
    for (int it = 0; it < ntiles; ++it) {
      const int tile = split + it * nsplit;
      const int buf = it & 1, m = it >> 1;
      mbar_wait(&acc_full[buf], m & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * 256;
      float* out = vp + (size_t)b * TC_VP + tile * TC_N;
      const float* vtt = vt + tile * TC_N;
#pragma unroll 1
      for (int cc = 0; cc < TC_N; cc += 16) {
        float v[16];
        tmem_ld16(taddr + cc, v);
        if (b < B) {
#pragma unroll
          for (int k = 0; k < 16; k += 4) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(vtt + cc + k));
            float4 o;
            o.x = v[k] + t.x; o.y = v[k + 1] + t.y; o.z = v[k + 2] + t.z; o.w = v[k + 3] + t.w;
            *reinterpret_cast<float4*>(out + cc + k) = o;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// -------------------------------------------------------------------------------------------------------
// Backward contraction:  gF[h][p] = sum_{c'} g_vposed[h][c'] * Pext[p][c']      (K = 2400 vertex coordinates)
// D[M = 128 hands][N = 160 features] accumulated over 300 k-steps, again 3xTF32.  Both operands stream through a
// 3-stage ring of 4 k-steps (A = the dL/dv_posed slabs the skinning backward wrote, B = the constant slabs).
// -------------------------------------------------------------------------------------------------------
constexpr int G_N = 160;
constexpr int G_KSTEPS = 300;
constexpr int G_KC = 4;
constexpr int G_NSTG = 3;
constexpr int G_A_SLAB = TC_M * 8;   // floats
constexpr int G_B_SLAB = G_N * 8;
constexpr uint32_t G_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(G_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);

struct GSmem {
  static constexpr int STAGE = 2 * G_KC * (G_A_SLAB + G_B_SLAB) * 4;   // A hi | A lo | B hi | B lo
  static constexpr int BARS = G_NSTG * STAGE;                          // full[3], empty[3], acc_full
  static constexpr int TMEM_PTR = BARS + 8 * 8;
  static constexpr int TOTAL = TMEM_PTR + 16;
};

__device__ __forceinline__ void umma_tf32_g(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(G_IDESC), "r"(accumulate)
      : "memory");
}

__global__ void __launch_bounds__(TC_THREADS, 1) mano_gfeat_tc_kernel(const float* __restrict__ gvh, const float* __restrict__ gvl,
                                                                       const float* __restrict__ Ph, const float* __restrict__ Pl,
                                                                       int B, int nsplit, float* __restrict__ gF) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + GSmem::BARS);
  uint64_t* full = bars;
  uint64_t* empty = bars + G_NSTG;
  uint64_t* acc_full = bars + 2 * G_NSTG;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + GSmem::TMEM_PTR);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = blockIdx.x;
  if (threadIdx.x == 0) {
    for (int s = 0; s < G_NSTG; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(acc_full, 1);
    mbar_fence_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  // split-K: part blockIdx.y of nsplit takes a contiguous range of the 75 ring fills (>= 3 fills each, so every one of the
  // three accumulators is initialised) and writes its own partial gF[part][B][160]; the pose backward adds them in order
  constexpr int NFILL_ALL = G_KSTEPS / G_KC;
  const int part = blockIdx.y;
  const int f0 = part * NFILL_ALL / nsplit, f1 = (part + 1) * NFILL_ALL / nsplit;
  const int NFILL = f1 - f0;
  gF += (size_t)part * (size_t)((B + TC_M - 1) / TC_M) * TC_M * G_N;
  constexpr uint32_t A_BYTES = G_KC * G_A_SLAB * 4, B_BYTES = G_KC * G_B_SLAB * 4;

  if (warp == 0) {
    if (lane == 0) {
      const float* ah = gvh + (size_t)g * G_KSTEPS * G_A_SLAB;
      const float* al = gvl + (size_t)g * G_KSTEPS * G_A_SLAB;
      for (int f = 0; f < NFILL; ++f) {
        const int s = f % G_NSTG, n = f / G_NSTG;
        mbar_wait(&empty[s], (n & 1) ^ 1);
        uint8_t* st = smem + s * GSmem::STAGE;
        mbar_arrive_expect_tx(&full[s], 2 * (A_BYTES + B_BYTES));
        bulk_g2s(st, ah + (size_t)(f0 + f) * G_KC * G_A_SLAB, A_BYTES, &full[s]);
        bulk_g2s(st + A_BYTES, al + (size_t)(f0 + f) * G_KC * G_A_SLAB, A_BYTES, &full[s]);
        bulk_g2s(st + 2 * A_BYTES, Ph + (size_t)(f0 + f) * G_KC * G_B_SLAB, B_BYTES, &full[s]);
        bulk_g2s(st + 2 * A_BYTES + B_BYTES, Pl + (size_t)(f0 + f) * G_KC * G_B_SLAB, B_BYTES, &full[s]);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // The tensor core's fp32 accumulation rounds toward zero on every add, so the error of one TMEM accumulator
      // grows with the length of its chain: 900 k-step MMAs are spread over three accumulators (one per ring
      // stage, 160 columns apart) that the epilogue adds in fp32.
      for (int f = 0; f < NFILL; ++f) {
        const int s = f % G_NSTG, n = f / G_NSTG;
        uint32_t acc = n > 0 ? 1u : 0u;
        const uint32_t dcol = tmem_base + s * G_N;
        mbar_wait(&full[s], n & 1);
        tc_fence_after();
        const uint32_t a_hi = smem_u32(smem + s * GSmem::STAGE), a_lo = a_hi + A_BYTES;
        const uint32_t b_hi = a_hi + 2 * A_BYTES, b_lo = b_hi + B_BYTES;
#pragma unroll
        for (int k = 0; k < G_KC; ++k) {
          const uint64_t ah = umma_desc(a_hi + k * G_A_SLAB * 4, TC_M * 16, 128);
          const uint64_t al = umma_desc(a_lo + k * G_A_SLAB * 4, TC_M * 16, 128);
          const uint64_t bh = umma_desc(b_hi + k * G_B_SLAB * 4, G_N * 16, 128);
          const uint64_t bl = umma_desc(b_lo + k * G_B_SLAB * 4, G_N * 16, 128);
          umma_tf32_g(dcol, al, bh, acc);
          umma_tf32_g(dcol, ah, bl, 1);
          umma_tf32_g(dcol, ah, bh, 1);
          acc = 1;
        }
        umma_commit(&empty[s]);
      }
      umma_commit(acc_full);
    }
  } else {
    const int q = warp & 3;
    const int b = g * TC_M + q * 32 + lane;
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    float* out = gF + (size_t)b * G_N;
#pragma unroll 1
    for (int cc = 0; cc < G_N; cc += 16) {
      float v[16], v1[16], v2[16];
      tmem_ld16(taddr + cc, v);
      tmem_ld16(taddr + G_N + cc, v1);
      tmem_ld16(taddr + 2 * G_N + cc, v2);
#pragma unroll
      for (int k = 0; k < 16; ++k) v[k] = (v[k] + v1[k]) + v2[k];
      if (b < B) {
#pragma unroll
        for (int k = 0; k < 16; k += 4) *reinterpret_cast<float4*>(out + cc + k) = make_float4(v[k], v[k + 1], v[k + 2], v[k + 3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

int gfeat_nsplit(int B) {
  // K (2400 vertex coordinates) is always split over G_MAXSPLIT CTAs: one CTA per 128 hands left SMs idle below 19k hands
  // (64 of 148 at B = 8192, 8 at B = 1024).  The count does not depend on B, so a hand's gradient is the same sum in the same
  // order however the batch is sharded.
  (void)B;
  return G_MAXSPLIT;
}

int launch_gfeat_tc(const float* gvh, const float* gvl, const float* Ph, const float* Pl, int B, float* gF, cudaStream_t st) {
  const int groups = (B + TC_M - 1) / TC_M;
  const int nsplit = gfeat_nsplit(B);
  HB_CUDA(cudaFuncSetAttribute(mano_gfeat_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GSmem::TOTAL));
  mano_gfeat_tc_kernel<<<dim3(groups, nsplit), TC_THREADS, GSmem::TOTAL, st>>>(gvh, gvl, Ph, Pl, B, nsplit, gF);
  g_launches++;
  return check_launch("mano_gfeat_tc_kernel");
}

size_t tc_smem_bytes() { return TcSmem::TOTAL; }

int launch_blend_tc(const float* Fhi, const float* Flo, const float* Bhi, const float* Blo, const float* vt, int B, float* vp,
                    cudaStream_t st) {
  const int groups = (B + TC_M - 1) / TC_M;
  int nsplit = 1;
  nsplit = 148 / groups;   // the most CTAs that still run as ONE wave (one CTA per SM): 229 vs 235 us at 8192 hands against
                           // the next power of two (two waves); small batches split all ten tiles
  if (nsplit < 1) nsplit = 1;
  if (nsplit > TC_TILES) nsplit = TC_TILES;
  HB_CUDA(cudaFuncSetAttribute(mano_blend_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TcSmem::TOTAL));
  dim3 grid(groups, nsplit);
  mano_blend_tc_kernel<<<grid, TC_THREADS, TcSmem::TOTAL, st>>>(Fhi, Flo, Bhi, Blo, vt, B, nsplit, vp);
  g_launches++;
  return check_launch("mano_blend_tc_kernel");
}

}  // namespace hb
