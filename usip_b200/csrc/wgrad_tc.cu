// wgrad_tc.cu -- tcgen05 3xTF32 weight-gradient GEMM:  gW[Cout,Cin] += GY[P,Cout]^T * act(X)[P,Cin].
//
// The reduction runs over the ROW dimension, so both operands are MN-major for the tensor core
// (element (m,k) = GY[k][m] has m contiguous).  For MN-major TF32 the only UMMA shared-memory layout is
// SWIZZLE_128B_BASE32B (cutlass sm100_common.inl:92; cute::UMMA::LayoutType 1, Swizzle<2,5,2> on byte addresses):
//   atom = 4 k-rows x 128 B (32 fp32 of the M/N dimension); the 32-byte chunk q of k-row r sits at chunk (q ^ (r & 3)),
//   M/N blocks of 32 elements LBO = 512 B apart, 4-row k-groups SBO = (blocks * 512) B apart,
// instruction descriptor with a_major = b_major = MN; one K=8 MMA consumes two k-groups.  One CTA owns one [128 x BN] tile of gW for one slice of
// rows (split-K over CTAs), accumulates it in TMEM over all its k-stages and adds it to gW with red.global once.
// Both operands come through registers (GY raw, X through the folded BN + ReLU prologue), are split into hi/lo
// TF32 halves and written with swizzled st.shared; 8 producer warps in two groups alternate over 32-row stages.
#include "common.cuh"

namespace usip {

constexpr int WT_THREADS = 13 * 32;        // warps 0-3 epilogue, 4 MMA, 5-12 producers
constexpr int WT_BK = 32;                  // rows per stage (= 4 MMA k-steps of 8)

__device__ __forceinline__ uint32_t wt_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void wt_mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void wt_mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void wt_mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nWT_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra WT_DONE;\nbra WT_WAIT;\nWT_DONE:\n}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ uint32_t wt_tf32(float x) { uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x)); return r; }

// MN-major SWIZZLE_128B_BASE32B descriptor: LBO = stride between 32-element M/N blocks, SBO = stride between 4-row k groups
__device__ __forceinline__ uint64_t wt_desc_mn(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;                                   // LayoutType::SWIZZLE_128B_BASE32B
  return d;
}
// byte offset of the 16-byte chunk c16 (0..7 inside a 32-element block) of stage-local row rl for block mb
__device__ __forceinline__ uint32_t wt_off(int rl, int mb, int nblocks, int c16) {
  const int g4 = rl >> 2, rr = rl & 3;
  return (uint32_t)((g4 * nblocks + mb) * 512 + rr * 128 + ((((c16 >> 1) ^ rr) << 5) | ((c16 & 1) << 4)));
}
__host__ __device__ constexpr uint32_t wt_idesc(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void wt_umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d), "l"(adesc),
               "l"(bdesc), "r"(idesc), "r"(accum)
               : "memory");
}

template <int BN, int STAGES>
struct WtSmem {
  static constexpr int A_HALF = WT_BK * 128 * 4;          // 32 rows x 128 fp32
  static constexpr int B_HALF = WT_BK * BN * 4;
  static constexpr int STAGE = 2 * A_HALF + 2 * B_HALF;
  static constexpr int BYTES = STAGES * STAGE + 256 + 1024;
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(WT_THREADS, 1)
wgrad_tc_kernel(const float* __restrict__ GY, int ldg, const float* __restrict__ X, int ldx,
                const float* __restrict__ in_scale, const float* __restrict__ in_shift, int in_relu,
                float* __restrict__ gW, int ldw, int P, int Cout, int Cin, int rows_per_cta, int single) {
  using SM = WtSmem<BN, STAGES>;
  constexpr int MBA = 4, MBB = BN / 32;
  extern __shared__ uint8_t wt_smem_raw[];
  const uint32_t base = (wt_smem_u32(wt_smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = wt_smem_raw + (base - wt_smem_u32(wt_smem_raw));
  const uint32_t bar_base = base + STAGES * SM::STAGE;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + STAGES * SM::STAGE + 192);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  const uint32_t done_bar = bar_base + 8u * (2 * STAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * 128, n0 = blockIdx.z * BN;
  const int r_begin = blockIdx.x * rows_per_cta, r_end = min(P, r_begin + rows_per_cta);
  const int nst = (r_end - r_begin + WT_BK - 1) / WT_BK;
  constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;

  if (warp == 4) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) { wt_mbar_init(full_bar(s), 128); wt_mbar_init(empty_bar(s), 1); }
      wt_mbar_init(done_bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(wt_smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = *tmem_slot;

  if (warp >= 5) {
    // ================================ producers ================================================
    const int pw = warp - 5, grp = pw >> 2, pt = (pw & 3) * 32 + lane;       // 0..127 inside the group
    // A operand (GY^T): 32 chunks of 16 B per row; thread -> chunk ca, rows ra + 4j
    const int ca = pt & 31, ra = pt >> 5;
    // B operand (act(X)^T): BN/4 chunks per row
    constexpr int CPRB = BN / 4, RPPB = 128 / CPRB, NJB = WT_BK / RPPB;      // chunks/row, rows/pass, passes
    const int cb = pt % CPRB, rb = pt / CPRB;
    float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
    if (in_scale) {
      sc = __ldg(reinterpret_cast<const float4*>(in_scale + n0 + cb * 4));
      sh = __ldg(reinterpret_cast<const float4*>(in_shift + n0 + cb * 4));
    }
    // The group's work is a stream of batches of <= 8 float4 per thread: batch 0 of an iteration is the A operand, the rest
    // are the B operand.  Two register sets alternate: the loads of batch g+1 are in flight while batch g is converted and
    // stored (round 1 waited for every batch of 8 loads in turn -- three exposed memory latencies per 32-row stage).
    constexpr int NBB = (NJB + 7) / 8, NB = 1 + NBB;
    const int niter = nst > grp ? (nst - grp + 1) / 2 : 0;
    const int total = niter * NB;
    const bool a_cols = m0 + ca * 4 < Cout;
    auto load = [&](float4 (&x)[8], int g) {
      const int b = g % NB, r0 = r_begin + (grp + 2 * (g / NB)) * WT_BK;
      if (b == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int r = r0 + ra + 4 * j;
          // output channels past Cout (a 64-wide layer in the 128-row UMMA tile) are zero rows of the operand
          x[j] = (r < r_end && a_cols) ? __ldg(reinterpret_cast<const float4*>(GY + (size_t)r * ldg + m0 + ca * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      } else {
        const int jb = (b - 1) * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int r = r0 + rb + RPPB * (jb + j);
          x[j] = (jb + j < NJB && r < r_end) ? __ldg(reinterpret_cast<const float4*>(X + (size_t)r * ldx + n0 + cb * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    };
    auto process = [&](const float4 (&x)[8], int g) {
      const int b = g % NB, it = grp + 2 * (g / NB), s = it % STAGES;
      const uint32_t a_hi = base + s * SM::STAGE, a_lo = a_hi + SM::A_HALF;
      const uint32_t b_hi = a_lo + SM::A_HALF, b_lo = b_hi + SM::B_HALF;
      if (b == 0) {
        wt_mbar_wait(empty_bar(s), ((it / STAGES) & 1) ^ 1);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int rl = ra + 4 * j;
          const float v[4] = {x[j].x, x[j].y, x[j].z, x[j].w};
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) { hi[q] = wt_tf32(v[q]); lo[q] = wt_tf32(v[q] - __uint_as_float(hi[q])); }
          const uint32_t off = wt_off(rl, ca >> 3, MBA, ca & 7);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a_hi + off), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]) : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a_lo + off), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]) : "memory");
        }
      } else {
        const int jb = (b - 1) * 8, r0 = r_begin + it * WT_BK;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (jb + j < NJB) {
            const int rl = rb + RPPB * (jb + j);
            float v[4] = {x[j].x, x[j].y, x[j].z, x[j].w};
            if (in_scale) { v[0] = fmaf(v[0], sc.x, sh.x); v[1] = fmaf(v[1], sc.y, sh.y); v[2] = fmaf(v[2], sc.z, sh.z); v[3] = fmaf(v[3], sc.w, sh.w); }
            if (in_relu) { v[0] = fmaxf(v[0], 0.f); v[1] = fmaxf(v[1], 0.f); v[2] = fmaxf(v[2], 0.f); v[3] = fmaxf(v[3], 0.f); }
            if (r0 + rl >= r_end) { v[0] = v[1] = v[2] = v[3] = 0.f; }
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) { hi[q] = wt_tf32(v[q]); lo[q] = wt_tf32(v[q] - __uint_as_float(hi[q])); }
            const uint32_t off = wt_off(rl, cb >> 3, MBB, cb & 7);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(b_hi + off), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]) : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(b_lo + off), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]) : "memory");
          }
        }
      }
      if (b == NB - 1) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        wt_mbar_arrive(full_bar(s));
      }
    };
    float4 x0[8], x1[8];
    if (total > 0) load(x0, 0);
#pragma unroll 1
    for (int g = 0; g < total; g += 2) {
      if (g + 1 < total) load(x1, g + 1);
      process(x0, g);
      if (g + 2 < total) load(x0, g + 2);
      if (g + 1 < total) process(x1, g + 1);
    }
  } else if (warp == 4) {
    // ================================ MMA issuer ===============================================
    constexpr uint32_t idesc = wt_idesc(128, BN);
    for (int it = 0; it < nst; ++it) {
      const int s = it % STAGES;
      wt_mbar_wait(full_bar(s), (it / STAGES) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (lane == 0) {
        const uint32_t a_hi = base + s * SM::STAGE, a_lo = a_hi + SM::A_HALF;
        const uint32_t b_hi = a_lo + SM::A_HALF, b_lo = b_hi + SM::B_HALF;
#pragma unroll
        for (int kg = 0; kg < 4; ++kg) {
          // k-step kg covers stage rows 8kg..8kg+7 = two 4-row groups, each (blocks * 512) B
          const uint64_t dah = wt_desc_mn(a_hi + kg * MBA * 1024, 512, MBA * 512), dal = wt_desc_mn(a_lo + kg * MBA * 1024, 512, MBA * 512);
          const uint64_t dbh = wt_desc_mn(b_hi + kg * MBB * 1024, 512, MBB * 512), dbl = wt_desc_mn(b_lo + kg * MBB * 1024, 512, MBB * 512);
          if (single) {                                  // plain TF32: hi x hi only (backward-precision option)
            wt_umma(tmem_d, dah, dbh, idesc, (it | kg) != 0);
          } else {
            wt_umma(tmem_d, dal, dbh, idesc, (it | kg) != 0);
            wt_umma(tmem_d, dah, dbl, idesc, 1u);
            wt_umma(tmem_d, dah, dbh, idesc, 1u);
          }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(empty_bar(s)) : "memory");
        if (it == nst - 1) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(done_bar) : "memory");
      }
      __syncwarp();
    }
  } else if (nst > 0) {
    // ================================ epilogue: TMEM -> red.global.add =========================
    wt_mbar_wait(done_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int m = m0 + warp * 32 + lane;                  // gW row (output channel) of this thread
    const bool vec = (ldw % 4 == 0) && ((reinterpret_cast<uintptr_t>(gW) & 15) == 0);
    const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
    for (int ch = 0; ch < BN / 32; ++ch) {
      uint32_t r[32];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
            "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
            "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
            "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
          : "r"(taddr + ch * 32)
          : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;"
                   : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                     "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]),
                     "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),
                     "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                   :
                   : "memory");
      if (m < Cout) {
        float* dst = gW + (size_t)m * ldw + n0 + ch * 32;
        if (vec) {                                         // 16-byte L2 reductions (REDG.ADD.F32x4): a quarter of the atomic traffic
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            atomicAdd(reinterpret_cast<float4*>(dst + j), make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                                      __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3])));
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) atomicAdd(dst + j, __uint_as_float(r[j]));
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 4) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(TMEM_COLS) : "memory");
  }
}

template <int BN, int STAGES>
static int launch_wgrad_tc(const float* GY, int ldg, const float* X, int ldx, const float* sc, const float* sh, int relu,
                           float* gW, int ldw, int P, int Cout, int Cin, int single, cudaStream_t st) {
  using SM = WtSmem<BN, STAGES>;
  static_assert(SM::BYTES <= 232448, "shared memory budget");
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::BYTES);
    if (e != cudaSuccess) { set_last_error("wgrad_tc smem attr"); return (int)e; }
    attr = true;
  }
  const int tiles = cdiv(Cout, 128) * (Cin / BN);
  int splits = max(1, min(cdiv(P, 4 * WT_BK), 148 / tiles));
  int rows = cdiv(cdiv(P, splits), WT_BK) * WT_BK;
  splits = cdiv(P, rows);
  dim3 grid(splits, cdiv(Cout, 128), Cin / BN);
  wgrad_tc_kernel<BN, STAGES><<<grid, WT_THREADS, SM::BYTES, st>>>(GY, ldg, X, ldx, sc, sh, relu, gW, ldw, P, Cout, Cin, rows, single);
  return check_launch("wgrad_tc_kernel");
}

// returns -2 when the shape is not eligible (caller falls back to the SIMT kernel)
int wgrad_tc(const float* GY, int ldg, const float* X, int ldx, const float* sc, const float* sh, int relu, float* gW,
             int ldw, int P, int Cout, int Cin, int single, cudaStream_t st) {
  const bool ok = (Cout % 4 == 0) && (Cout >= 64) && (Cin % 64 == 0) && P >= 4096 && (ldg % 4 == 0) && (ldx % 4 == 0) &&
                  (reinterpret_cast<uintptr_t>(GY) % 16 == 0) && (reinterpret_cast<uintptr_t>(X) % 16 == 0) &&
                  (!sc || (reinterpret_cast<uintptr_t>(sc) % 16 == 0 && reinterpret_cast<uintptr_t>(sh) % 16 == 0));
  if (!ok) return -2;
  if (Cin % 256 == 0) return launch_wgrad_tc<256, 2>(GY, ldg, X, ldx, sc, sh, relu, gW, ldw, P, Cout, Cin, single, st);
  if (Cin % 128 == 0) return launch_wgrad_tc<128, 3>(GY, ldg, X, ldx, sc, sh, relu, gW, ldw, P, Cout, Cin, single, st);
  return launch_wgrad_tc<64, 4>(GY, ldg, X, ldx, sc, sh, relu, gW, ldw, P, Cout, Cin, single, st);
}

}  // namespace usip
