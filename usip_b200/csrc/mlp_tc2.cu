// mlp_tc2.cu -- CTA-pair (cta_group::2) variant of the 3xTF32 layer kernel for the wide layers (Cout % 256 == 0).
//
// A cluster of two CTAs (one TPC) owns a 256-row x 256-column output tile: each CTA produces the A operand of its own
// 128 rows and HALF of the weight tile (128 of the 256 columns); `tcgen05.mma.cta_group::2` (M = 256), issued by the
// leader CTA only, reads A from both CTAs and the two B halves, and leaves rows 0-127 of the accumulator in the
// leader's TMEM and rows 128-255 in the peer's.  Compared with the single-CTA kernel this halves the per-CTA weight
// traffic (shared-memory writes and tensor-core operand reads), which makes room for a third pipeline stage -- the
// single-CTA kernel is limited by exactly those two things.
//
// Synchronisation (all mbarriers, no __syncthreads in the steady state):
//   bfull[s]  (local)  : the CTA's own two bulk-TMA weight copies of stage s (expect_tx)
//   full[s]   (leader) : 2 x 4 producer warps (the peer arrives remotely through its cluster-mapped address)
//   empty[s]  (local)  : tcgen05.commit.cta_group::2 ... multicast::cluster from the leader's MMA thread
//   tfull[b]  (local)  : same multicast commit after the last K chunk of a tile
//   tempty[b] (leader) : 2 x 8 epilogue warps
#include "tc_common.cuh"

namespace usip {

constexpr int T2_EPI_WARPS = 8;
constexpr int T2_MMA_WARP = 8;
constexpr int T2_PROD_WARP0 = 9;
constexpr int T2_THREADS = 17 * 32;
constexpr int T2_BN = 256;
constexpr int T2_STAGES = 3;

struct T2Smem {
  static constexpr int A_STAGE = 2 * TC_BM * 128;               // A hi + lo (this CTA's 128 rows)
  static constexpr int B_STAGE = 2 * (T2_BN / 2) * 128;         // B hi + lo, this CTA's 128 of the 256 columns
  static constexpr int STAGE = A_STAGE + B_STAGE;               // 64 KB
  static constexpr int TRANS = T2_EPI_WARPS * 32 * 32 * 4;
  static constexpr int BARS = 256;
  static constexpr int BYTES = T2_STAGES * STAGE + TRANS + BARS + 1024;
};

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void umma_tf32_2cta(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2cta(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((uint16_t)3)
               : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(T2_THREADS, 1)
layer_fwd_tc2_kernel(const usip_layer_desc d, const uint32_t* __restrict__ wpack) {
  using SM = T2Smem;
  constexpr int BN = T2_BN, STAGES = T2_STAGES, HB = BN / 2;
  extern __shared__ uint8_t smem_raw2[];
  const uint32_t smem_base = (smem_u32(smem_raw2) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw2 + (smem_base - smem_u32(smem_raw2));
  const uint32_t stage_base = smem_base;
  float* trans = reinterpret_cast<float*>(smem + STAGES * SM::STAGE);
  const uint32_t bar_base = smem_base + STAGES * SM::STAGE + SM::TRANS;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + STAGES * SM::STAGE + SM::TRANS + 192);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };                       // used in the leader
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto bfull_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
  auto tfull_bar = [&](int b) { return bar_base + 8u * (3 * STAGES + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8u * (3 * STAGES + 2 + b); };  // used in the leader

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int P = d.P, Cin = d.Cin, Cout = d.Cout;
  const int KC = Cin / TC_BK;
  const int m_tiles = (P + 2 * TC_BM - 1) / (2 * TC_BM), n_tiles = Cout / BN;
  const int num_tiles = m_tiles * n_tiles;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  constexpr uint32_t TMEM_COLS = 512;

  if (warp == T2_MMA_WARP) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 8); mbar_init(empty_bar(s), 1); mbar_init(bfull_bar(s), 1); }
      for (int b = 0; b < 2; ++b) { mbar_init(tfull_bar(b), 1); mbar_init(tempty_bar(b), 2 * T2_EPI_WARPS); }
      fence_barrier_init();
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                               // both CTAs' barriers are initialised before any remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= T2_PROD_WARP0) {
    // =============================== A producers (+ this CTA's half of the weight tile) =========
    const int pw = warp - T2_PROD_WARP0;
    const int grp = pw >> 2;
    const int pt = (pw & 3) * 32 + lane;
    const int c = pt & 7, r0 = pt >> 3;
    const bool has_aff = d.in_scale != nullptr;
    // flat iteration space over (tile, K chunk); this group handles every other iteration; the X tile of the NEXT
    // iteration is fetched into registers before waiting for the stage to be freed (hides the global-load latency)
    const int my_tiles = (num_tiles - cluster_id + num_clusters - 1) / num_clusters;
    const uint32_t total_it = (uint32_t)max(my_tiles, 0) * (uint32_t)KC;
    auto fetch = [&](uint32_t it2, float4 (&x)[8]) {
      const int tile2 = cluster_id + (int)(it2 / KC) * num_clusters;
      const int kc2 = (int)(it2 % KC);
      const int row02 = (tile2 / n_tiles) * 2 * TC_BM + (int)rank * TC_BM;
      const int k2 = kc2 * TC_BK + c * 4;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int row = row02 + r0 + 16 * j;
        x[j] = row < P ? __ldg(reinterpret_cast<const float4*>(d.X + (size_t)row * d.ldx + k2)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    float4 x[8];
    if ((uint32_t)grp < total_it) fetch((uint32_t)grp, x);
    for (uint32_t it = (uint32_t)grp; it < total_it; it += 2) {
      const int tile = cluster_id + (int)(it / KC) * num_clusters;
      const int kc = (int)(it % KC);
      const int mt = tile / n_tiles, nt = tile - mt * n_tiles;
      const int row0 = mt * 2 * TC_BM + (int)rank * TC_BM;
      {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(empty_bar(s), ph ^ 1);
        const uint32_t a_hi = stage_base + s * SM::STAGE;
        const uint32_t a_lo = a_hi + TC_BM * 128;
        const uint32_t b_hi = a_hi + SM::A_STAGE, b_lo = b_hi + HB * 128;
        if (pt == 0) {
          mbar_arrive_expect_tx(bfull_bar(s), SM::B_STAGE);
          // packed layout (BN = 256): [hi: 256 rows x 128 B | lo: 256 rows x 128 B]; this CTA takes rows rank*128..+127
          const uint32_t* src = wpack + ((size_t)nt * KC + kc) * 2 * (size_t)BN * TC_BK;
          bulk_g2s(b_hi, src + (size_t)rank * HB * TC_BK, HB * 128, bfull_bar(s));
          bulk_g2s(b_lo, src + (size_t)BN * TC_BK + (size_t)rank * HB * TC_BK, HB * 128, bfull_bar(s));
        }
        const int k = kc * TC_BK + c * 4;
        float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
        if (has_aff) {
          sc = __ldg(reinterpret_cast<const float4*>(d.in_scale + k));
          sh = __ldg(reinterpret_cast<const float4*>(d.in_shift + k));
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int r = r0 + 16 * j;
          float v[4] = {x[j].x, x[j].y, x[j].z, x[j].w};
          if (has_aff) {
            v[0] = fmaf(v[0], sc.x, sh.x); v[1] = fmaf(v[1], sc.y, sh.y);
            v[2] = fmaf(v[2], sc.z, sh.z); v[3] = fmaf(v[3], sc.w, sh.w);
          }
          if (d.in_relu) { v[0] = fmaxf(v[0], 0.f); v[1] = fmaxf(v[1], 0.f); v[2] = fmaxf(v[2], 0.f); v[3] = fmaxf(v[3], 0.f); }
          if (row0 + r >= P) { v[0] = v[1] = v[2] = v[3] = 0.f; }
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) split_tf32(v[q], hi[q], lo[q]);
          const uint32_t off = (uint32_t)r * 128u + (uint32_t)((c ^ (r & 7)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a_hi + off), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]) : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a_lo + off), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]) : "memory");
        }
        if (it + 2 < total_it) fetch(it + 2, x);
        if (pt == 0) mbar_wait(bfull_bar(s), ph);  // this CTA's weight half has landed (async proxy writes)
        asm volatile("fence.proxy.async;" ::: "memory");
        __syncwarp();                              // one (remote) arrive per warp: 8 per stage instead of 256
        if (lane == 0) mbar_arrive_cluster(mapa_u32(full_bar(s), 0));    // arrive on the LEADER's full barrier
      }
    }
  } else if (warp == T2_MMA_WARP) {
    if (leader) {
      // =============================== MMA issuer (leader CTA, one thread) =======================
      constexpr uint32_t idesc = make_idesc_tf32(2 * TC_BM, BN);
      uint32_t it = 0, tcount = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++tcount) {
        const uint32_t buf = tcount & 1;
        mbar_wait(tempty_bar(buf), ((tcount >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + buf * BN;
        for (int kc = 0; kc < KC; ++kc, ++it) {
          const int s = it % STAGES;
          mbar_wait(full_bar(s), (it / STAGES) & 1);
          tc_fence_after();
          if (lane == 0) {
            const uint32_t a_hi = stage_base + s * SM::STAGE;
            const uint32_t a_lo = a_hi + TC_BM * 128;
            const uint32_t b_hi = a_hi + SM::A_STAGE, b_lo = b_hi + HB * 128;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint64_t dah = make_kmajor_sw128_desc(a_hi + ks * 32), dal = make_kmajor_sw128_desc(a_lo + ks * 32);
              const uint64_t dbh = make_kmajor_sw128_desc(b_hi + ks * 32), dbl = make_kmajor_sw128_desc(b_lo + ks * 32);
              umma_tf32_2cta(tmem_d, dal, dbh, idesc, (kc | ks) != 0);
              umma_tf32_2cta(tmem_d, dah, dbl, idesc, 1u);
              umma_tf32_2cta(tmem_d, dah, dbh, idesc, 1u);
            }
            umma_commit_2cta(empty_bar(s));                       // frees stage s in BOTH CTAs
            if (kc == KC - 1) umma_commit_2cta(tfull_bar(buf));   // accumulator ready in BOTH CTAs
          }
          __syncwarp();
        }
      }
    }
  } else {
    // =============================== epilogue warps 0..7 (own 128 rows) =========================
    const int q = warp & 3, half = warp >> 2;
    float* tw = trans + warp * (32 * 32);
    const int g = d.group;
    const bool want_stats = d.stat_partial != nullptr;
    const bool want_grp = (d.gmax != nullptr) || (d.gmin != nullptr);
    uint32_t tcount = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++tcount) {
      const int mt = tile / n_tiles, nt = tile - mt * n_tiles;
      const int row0 = mt * 2 * TC_BM + (int)rank * TC_BM, n0 = nt * BN;
      const uint32_t buf = tcount & 1;
      const int wrow0 = row0 + q * 32;
      const int row = wrow0 + lane;
      const bool rok = row < P;
      const int nvalid = min(32, max(0, P - wrow0));
      const float* addp = nullptr;
      if (d.addend && rok) {
        const int gi = d.add_index ? __ldg(d.add_index + row) : row / d.add_group;
        addp = d.addend + (size_t)gi * d.ld_add + n0;
      }
      mbar_wait(tfull_bar(buf), (tcount >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + buf * BN + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int ch = half; ch < BN / 32; ch += 2) {
        const int cb = n0 + ch * 32;
        uint32_t raw[32];
        tmem_ld_32x32_issue(taddr + ch * 32, raw);
        float4 a4[8];
        if (addp) {
#pragma unroll
          for (int j = 0; j < 8; ++j) a4[j] = __ldg(reinterpret_cast<const float4*>(addp + ch * 32) + j);
        }
        tmem_ld_wait(raw);
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
        if (d.bias) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(d.bias + cb) + j);
            v[4 * j] += b4.x; v[4 * j + 1] += b4.y; v[4 * j + 2] += b4.z; v[4 * j + 3] += b4.w;
          }
        }
        if (addp) {
#pragma unroll
          for (int j = 0; j < 8; ++j) { v[4 * j] += a4[j].x; v[4 * j + 1] += a4[j].y; v[4 * j + 2] += a4[j].z; v[4 * j + 3] += a4[j].w; }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<float4*>(tw + lane * 32 + ((j ^ (lane & 7)) << 2)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        __syncwarp();
        if (d.Y) {
          const int l8 = lane & 7, rsub = lane >> 3;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = 4 * i + rsub;
            if (r < nvalid) {
              const float4 o = *reinterpret_cast<const float4*>(tw + r * 32 + ((l8 ^ (r & 7)) << 2));
              *reinterpret_cast<float4*>(d.Y + (size_t)(wrow0 + r) * d.ldy + cb + l8 * 4) = o;
            }
          }
        }
        const int csub = lane & 3, cchunk = lane >> 2;
        const size_t st = (size_t)(wrow0 / 32);
        if (want_stats && !want_grp) {
          float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
          if (nvalid == 32) {
#pragma unroll
            for (int r = 0; r < 32; r += 2) {
              const float x0 = tw[r * 32 + ((cchunk ^ (r & 7)) << 2) + csub];
              const float x1 = tw[(r + 1) * 32 + ((cchunk ^ ((r + 1) & 7)) << 2) + csub];
              s0 += x0; q0 = fmaf(x0, x0, q0); s1 += x1; q1 = fmaf(x1, x1, q1);
            }
          } else {
            for (int r = 0; r < nvalid; ++r) { const float x0 = tw[r * 32 + ((cchunk ^ (r & 7)) << 2) + csub]; s0 += x0; q0 = fmaf(x0, x0, q0); }
          }
          if (nvalid > 0) {
            d.stat_partial[(st * 2 + 0) * Cout + cb + lane] = s0 + s1;
            d.stat_partial[(st * 2 + 1) * Cout + cb + lane] = q0 + q1;
          }
        } else if (want_grp) {
          float s = 0.f, ss = 0.f;
          float mx0 = -INFINITY, mn0 = INFINITY, mx1 = -INFINITY, mn1 = INFINITY;
          int ax0 = 0, an0 = 0, ax1 = 0, an1 = 0;
          const int hrows = (g == 16) ? 16 : 32;
#pragma unroll 8
          for (int r = 0; r < 32; ++r) {
            const float x = tw[r * 32 + ((cchunk ^ (r & 7)) << 2) + csub];
            const bool ok = r < nvalid;
            if (ok) { s += x; ss = fmaf(x, x, ss); }
            if (r < hrows) {
              if (ok && x > mx0) { mx0 = x; ax0 = r; }
              if (ok && x < mn0) { mn0 = x; an0 = r; }
            } else {
              if (ok && x > mx1) { mx1 = x; ax1 = r; }
              if (ok && x < mn1) { mn1 = x; an1 = r; }
            }
          }
          if (want_stats && nvalid > 0) {
            d.stat_partial[(st * 2 + 0) * Cout + cb + lane] = s;
            d.stat_partial[(st * 2 + 1) * Cout + cb + lane] = ss;
          }
          if (g == 16) {
            const int grow = wrow0 / 16;
            if (nvalid > 0) {
              if (d.gmax) d.gmax[(size_t)grow * Cout + cb + lane] = mx0;
              if (d.gmin) d.gmin[(size_t)grow * Cout + cb + lane] = mn0;
              if (d.garg_max) d.garg_max[(size_t)grow * Cout + cb + lane] = ax0;
              if (d.garg_min) d.garg_min[(size_t)grow * Cout + cb + lane] = an0;
            }
            if (nvalid > 16) {
              if (d.gmax) d.gmax[(size_t)(grow + 1) * Cout + cb + lane] = mx1;
              if (d.gmin) d.gmin[(size_t)(grow + 1) * Cout + cb + lane] = mn1;
              if (d.garg_max) d.garg_max[(size_t)(grow + 1) * Cout + cb + lane] = ax1 - 16;
              if (d.garg_min) d.garg_min[(size_t)(grow + 1) * Cout + cb + lane] = an1 - 16;
            }
          } else if (nvalid > 0) {                   // g == 32 (enforced by the launcher)
            const int grow = wrow0 / 32;
            if (d.gmax) d.gmax[(size_t)grow * Cout + cb + lane] = mx0;
            if (d.gmin) d.gmin[(size_t)grow * Cout + cb + lane] = mn0;
            if (d.garg_max) d.garg_max[(size_t)grow * Cout + cb + lane] = ax0;
            if (d.garg_min) d.garg_min[(size_t)grow * Cout + cb + lane] = an0;
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(tempty_bar(buf), 0));   // leader's tempty collects both CTAs
    }
  }

  // ---------------------------------------------------------------------------------- teardown
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                               // neither CTA may exit / free TMEM while the peer still uses it
  if (warp == T2_MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

bool tc2_eligible(const usip_layer_desc& d) {
  if (d.Cout % T2_BN != 0) return false;
  if ((d.gmax || d.gmin) && !(d.group == 16 || d.group == 32)) return false;
  const long long tiles = (long long)cdiv(d.P, 2 * TC_BM) * (d.Cout / T2_BN);
  return tiles >= 64;                               // enough 256x256 tiles for the 74 CTA pairs
}

int launch_tc2(const usip_layer_desc& d, const uint32_t* wpack, cudaStream_t st) {
  static_assert(T2Smem::BYTES <= 232448, "shared memory budget");
  static int sm_count = 0;
  if (sm_count == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
    cudaError_t e = cudaFuncSetAttribute(layer_fwd_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, T2Smem::BYTES);
    if (e != cudaSuccess) { sm_count = 0; set_last_error("layer_fwd_tc2 smem attr"); return (int)e; }
  }
  const int tiles = cdiv(d.P, 2 * TC_BM) * (d.Cout / T2_BN);
  const int grid = 2 * min(sm_count / 2, tiles);
  layer_fwd_tc2_kernel<<<grid, T2_THREADS, T2Smem::BYTES, st>>>(d, wpack);
  return check_launch("layer_fwd_tc2_kernel");
}

}  // namespace usip
