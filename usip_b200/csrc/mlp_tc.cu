// mlp_tc.cu -- tcgen05 / TMEM path of usip_layer_fwd (precision == 1): 3xTF32 error-compensated GEMM
//
//     Y[P,Cout] = act(X)[P,Cin] * W[Cout,Cin]^T (+bias, +addend),     act = folded BN affine + ReLU
//
// with  a = a_hi + a_lo (a_hi = rna_tf32(a), a_lo = rna_tf32(a - a_hi)) and likewise for W:
//     D += a_lo*w_hi + a_hi*w_lo + a_hi*w_hi        (dropped term a_lo*w_lo ~ 2^-22 relative)
// which keeps fp32-level accuracy (the 1e-4 parity bar) at 1/3 of the TF32 tensor rate.
//
// Persistent, warp-specialised CTA (one per SM), 17 warps:
//   warps 0-7  : epilogue  (TMEM lanes 32*(w%4)..+31, two warps per quarter on alternating 32-column
//                chunks): tcgen05.ld -> bias/addend -> Y, BN statistic partials and per-group max/min
//                (+arg) of the raw output
//   warp  8    : TMEM allocation, mbarrier init, single-thread tcgen05.mma issue
//   warps 9-16 : A-operand producers, two groups of four warps on alternating K chunks: coalesced
//                float4 loads of X, BN/ReLU prologue, hi/lo split, 128B-swizzled st.shared;
//                one elected thread per chunk also issues the bulk-TMA (cp.async.bulk) copies of the
//                pre-split, pre-swizzled weight tiles.
// Pipelines: smem full/empty ring (STAGES deep) between producers and MMA, TMEM full/empty (2
// accumulators) between MMA and epilogue, static round-robin tile scheduler.
//
// Operand tiles are K-major, SWIZZLE_128B: row r (128 B = 32 tf32) at r*128, 16-byte chunk c stored at
// chunk (c ^ (r & 7)); 8-row groups 1024 B apart (SBO).  One K chunk = 32 floats = 4 MMA k-steps of 8.
#include "tc_common.cuh"

namespace usip {

constexpr int TC_EPI_WARPS = 8;            // two warps per TMEM lane quarter, on alternating 32-column chunks
constexpr int TC_MMA_WARP = 8;
constexpr int TC_PROD_WARP0 = 9;
constexpr int TC_PROD_WARPS = 8;
constexpr int TC_THREADS = (TC_PROD_WARP0 + TC_PROD_WARPS) * 32;   // 544
constexpr int TC_STAT_ROWS = 32;           // BN-statistic partials are emitted per 32-row warp slice

// ------------------------------------------------------------------------------------------------
// weight pre-pack: W[Cout,Cin] (row stride ldw) -> per (n_tile, k_chunk): [hi | lo] blocks of BN x 128 B,
// already 128B-swizzled, so a tile is one contiguous bulk copy.
// ------------------------------------------------------------------------------------------------
__global__ void tc_pack_weights_kernel(const float* __restrict__ W, int ldw, int transposed, int Cout, int Cin, int BN,
                                       uint32_t* __restrict__ out) {
  const int KC = Cin / TC_BK;
  const int total = Cout * (Cin / 4);                 // one thread per 16-byte chunk
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int n = t / (Cin / 4), k4 = t - n * (Cin / 4);
  const int kc = k4 / 8, c = k4 & 7;                  // chunk c of K chunk kc
  const int nt = n / BN, r = n - nt * BN;
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float w = transposed ? W[(size_t)(k4 * 4 + j) * ldw + n] : W[(size_t)n * ldw + k4 * 4 + j];
    split_tf32(w, hi[j], lo[j]);
  }
  const size_t blk = ((size_t)nt * KC + kc) * 2 * (size_t)BN * TC_BK;     // in 32-bit words
  const size_t off = (size_t)r * TC_BK + (size_t)((c ^ (r & 7)) * 4);
  *reinterpret_cast<uint4*>(out + blk + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  *reinterpret_cast<uint4*>(out + blk + (size_t)BN * TC_BK + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

// ------------------------------------------------------------------------------------------------
template <int BN, int STAGES, bool COMBINE>
struct TcSmem {
  static constexpr int A_STAGE = 2 * TC_BM * 128;                 // hi + lo
  static constexpr int B_STAGE = 2 * BN * 128;
  static constexpr int STAGE = A_STAGE + B_STAGE;
  static constexpr int TRANS = TC_EPI_WARPS * 32 * 32 * 4;       // per-warp 32x32 staging tile, XOR-swizzled 16B chunks
  static constexpr int COMB = COMBINE ? 4 * 4 * BN * 4 : 0;      // max, min, argmax, argmin per lane quarter (group > 32 only)
  static constexpr int BARS = 256;
  static constexpr int BYTES = STAGES * STAGE + TRANS + COMB + BARS + 1024;   // +1024 alignment slack
};

template <int BN, int STAGES, bool COMBINE>
__global__ void __launch_bounds__(TC_THREADS, 1)
layer_fwd_tc_kernel(const usip_layer_desc d, const uint32_t* __restrict__ wpack) {
  using SM = TcSmem<BN, STAGES, COMBINE>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;       // SWIZZLE_128B needs 1024-B alignment
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
  // carve
  const uint32_t stage_base = smem_base;
  float* trans = reinterpret_cast<float*>(smem + STAGES * SM::STAGE);
  float* comb = reinterpret_cast<float*>(smem + STAGES * SM::STAGE + SM::TRANS);
  const uint32_t bar_base = smem_base + STAGES * SM::STAGE + SM::TRANS + SM::COMB;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + STAGES * SM::STAGE + SM::TRANS + SM::COMB + 192);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int b) { return bar_base + 8u * (2 * STAGES + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8u * (2 * STAGES + 2 + b); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int P = d.P, Cin = d.Cin, Cout = d.Cout;
  const int KC = Cin / TC_BK;
  const int m_tiles = (P + TC_BM - 1) / TC_BM, n_tiles = Cout / BN;
  const int num_tiles = m_tiles * n_tiles;
  constexpr uint32_t TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64 ? 64 : (2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512)));

  if (warp == TC_MMA_WARP) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 128 + 1); mbar_init(empty_bar(s), 1); }
      for (int b = 0; b < 2; ++b) { mbar_init(tfull_bar(b), 1); mbar_init(tempty_bar(b), TC_EPI_WARPS); }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= TC_PROD_WARP0) {
    // =============================== A producers (+ weight bulk copies) ===========================
    const int pw = warp - TC_PROD_WARP0;            // 0..7
    const int grp = pw >> 2;                        // K chunks with (it & 1) == grp
    const int pt = (pw & 3) * 32 + lane;            // 0..127 inside the group
    const int c = pt & 7;                           // 16-byte chunk (4 floats) of the 128-byte row
    const int r0 = pt >> 3;                         // rows r0 + 16*j
    const bool has_aff = d.in_scale != nullptr;
    // flat iteration space over (tile, K chunk); this group handles every other iteration.  The X tile of the NEXT
    // iteration is fetched into registers before waiting for the stage to be freed, so the global-load latency
    // (1-2 us under load) overlaps the tensor core working on the other stage(s).
    const int my_tiles = (num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const uint32_t total_it = (uint32_t)max(my_tiles, 0) * (uint32_t)KC;
    auto fetch = [&](uint32_t it2, float4 (&x)[8]) {
      const int tile2 = (int)blockIdx.x + (int)(it2 / KC) * (int)gridDim.x;
      const int kc2 = (int)(it2 % KC);
      const int row02 = (tile2 / n_tiles) * TC_BM;
      const int k2 = kc2 * TC_BK + c * 4;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int row = row02 + r0 + 16 * j;
        x[j] = row < P ? __ldg(reinterpret_cast<const float4*>(d.X + (size_t)row * d.ldx + k2)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    float4 x[8];
    const bool dbg_noload = (d.debug_flags & 2) != 0, dbg_notma = (d.debug_flags & 4) != 0;
    if (dbg_noload) {
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = make_float4(1.f, 2.f, 3.f, 4.f);
    }
    if ((uint32_t)grp < total_it && !dbg_noload) fetch((uint32_t)grp, x);
    for (uint32_t it = (uint32_t)grp; it < total_it; it += 2) {
      const int tile = (int)blockIdx.x + (int)(it / KC) * (int)gridDim.x;
      const int kc = (int)(it % KC);
      const int mt = tile / n_tiles, nt = tile - mt * n_tiles;
      const int row0 = mt * TC_BM;
      {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(empty_bar(s), ph ^ 1);
        const uint32_t a_hi = stage_base + s * SM::STAGE;
        const uint32_t a_lo = a_hi + TC_BM * 128;
        const uint32_t b_hi = a_hi + SM::A_STAGE;
        if (pt == 0) {
          if (dbg_notma) {
            mbar_arrive(full_bar(s));
          } else {
            mbar_arrive_expect_tx(full_bar(s), SM::B_STAGE);
            const uint32_t* src = wpack + ((size_t)nt * KC + kc) * 2 * (size_t)BN * TC_BK;
            bulk_g2s(b_hi, src, SM::B_STAGE, full_bar(s));          // [hi | lo] contiguous
          }
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
        if (it + 2 < total_it && !dbg_noload) fetch(it + 2, x);  // next tile of this group: in flight while we fence / arrive / wait
        fence_proxy_async_smem();                  // generic-proxy writes -> visible to the tensor core (async proxy)
        mbar_arrive(full_bar(s));                  // every producer thread arrives after fencing its own stores
      }
    }
  } else if (warp == TC_MMA_WARP) {
    // =============================== MMA issuer (one thread) =====================================
    constexpr uint32_t idesc = make_idesc_tf32(TC_BM, BN);
    uint32_t it = 0, tcount = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tcount) {
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
          const uint32_t b_hi = a_hi + SM::A_STAGE;
          const uint32_t b_lo = b_hi + BN * 128;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t dah = make_kmajor_sw128_desc(a_hi + ks * 32), dal = make_kmajor_sw128_desc(a_lo + ks * 32);
            const uint64_t dbh = make_kmajor_sw128_desc(b_hi + ks * 32), dbl = make_kmajor_sw128_desc(b_lo + ks * 32);
            umma_tf32(tmem_d, dal, dbh, idesc, (kc | ks) != 0);
            if (!(d.debug_flags & 8)) {
              umma_tf32(tmem_d, dah, dbl, idesc, 1u);
              umma_tf32(tmem_d, dah, dbh, idesc, 1u);
            }
          }
          umma_commit(empty_bar(s));                       // frees the smem stage when these MMAs retire
          if (kc == KC - 1) umma_commit(tfull_bar(buf));   // accumulator ready for the epilogue
        }
        __syncwarp();
      }
    }
  } else {
    // =============================== epilogue warps 0..7 ========================================
    const int q = warp & 3;                           // TMEM lane quarter
    const int half = warp >> 2;                       // handles chunks with (ch & 1) == half
    float* tw = trans + warp * (32 * 32);
    const int g = d.group;
    const bool want_stats = d.stat_partial != nullptr;
    const bool want_grp = (d.gmax != nullptr) || (d.gmin != nullptr);
    // staging tile addressing: element (r, c) lives at r*32 + (((c>>2) ^ (r&7))<<2) + (c&3)
    uint32_t tcount = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tcount) {
      const int mt = tile / n_tiles, nt = tile - mt * n_tiles;
      const int row0 = mt * TC_BM, n0 = nt * BN;
      const uint32_t buf = tcount & 1;
      const int wrow0 = row0 + q * 32;                // first row of this warp
      const int row = wrow0 + lane;
      const bool rok = row < P;
      const int nvalid = min(32, max(0, P - wrow0));  // valid rows of this warp
      const float* addp = nullptr;
      if (d.addend && rok) {
        const int gi = d.add_index ? __ldg(d.add_index + row) : row / d.add_group;
        addp = d.addend + (size_t)gi * d.ld_add + n0;
      }
      mbar_wait(tfull_bar(buf), (tcount >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + buf * BN + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int ch = half; ch < ((d.debug_flags & 1) ? 0 : BN / 32); ch += 2) {
        const int cb = n0 + ch * 32;
        // TMEM read and the addend gather are issued back to back so their latencies overlap
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
        // stage the 32x32 chunk in shared memory: row = lane, 16-byte chunk j stored at chunk (j ^ (lane & 7))
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<float4*>(tw + lane * 32 + ((j ^ (lane & 7)) << 2)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        __syncwarp();
        if (d.Y) {
          // coalesced stores: 8 lanes cover one 128-byte row segment, a warp instruction writes 4 full lines
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
        // column view: lane owns column cb + lane; element (r, lane) at r*32 + (((lane>>2) ^ (r&7))<<2) + (lane&3)
        const int cl = ch * 32 + lane;                // column inside the tile
        const int csub = lane & 3, cchunk = lane >> 2;
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
          if (nvalid > 0) {                          // slices entirely past P have no partial row
            const size_t st = (size_t)(mt * 4 + q);
            d.stat_partial[(st * 2 + 0) * Cout + cb + lane] = s0 + s1;
            d.stat_partial[(st * 2 + 1) * Cout + cb + lane] = q0 + q1;
          }
        } else if (want_grp) {
          float s = 0.f, ss = 0.f;
          float mx0 = -INFINITY, mn0 = INFINITY, mx1 = -INFINITY, mn1 = INFINITY;
          int ax0 = 0, an0 = 0, ax1 = 0, an1 = 0;
          const int hrows = (g == 16) ? 16 : 32;      // rows per in-warp group segment
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
            const size_t st = (size_t)(mt * 4 + q);
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
          } else if (g == 32) {
            if (nvalid > 0) {
              const int grow = wrow0 / 32;
              if (d.gmax) d.gmax[(size_t)grow * Cout + cb + lane] = mx0;
              if (d.gmin) d.gmin[(size_t)grow * Cout + cb + lane] = mn0;
              if (d.garg_max) d.garg_max[(size_t)grow * Cout + cb + lane] = ax0;
              if (d.garg_min) d.garg_min[(size_t)grow * Cout + cb + lane] = an0;
            }
          } else if (COMBINE) {                       // 64 / 128: combine the lane quarters below
            comb[(0 * 4 + q) * BN + cl] = mx0; comb[(1 * 4 + q) * BN + cl] = mn0;
            reinterpret_cast<int*>(comb)[(2 * 4 + q) * BN + cl] = q * 32 + ax0;
            reinterpret_cast<int*>(comb)[(3 * 4 + q) * BN + cl] = q * 32 + an0;
          }
        }
        __syncwarp();                                   // tw is rewritten by the next chunk
      }
      // accumulator drained: hand the TMEM buffer back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(buf));
      if (COMBINE && want_grp && g > 32) {
        named_bar_sync(1, TC_EPI_WARPS * 32);
        const int wpg = g / 32;                        // lane quarters per group: 2 or 4
        for (int cl = threadIdx.x; cl < BN; cl += TC_EPI_WARPS * 32) {
          for (int gi = 0; gi < 4 / wpg; ++gi) {
            const int grow = row0 / g + gi;
            if ((size_t)grow * g >= (size_t)P) break;
            float mx = -INFINITY, mn = INFINITY; int ax = 0, an = 0;
            for (int w = gi * wpg; w < (gi + 1) * wpg; ++w) {
              const float a = comb[(0 * 4 + w) * BN + cl], b = comb[(1 * 4 + w) * BN + cl];
              if (a > mx) { mx = a; ax = reinterpret_cast<int*>(comb)[(2 * 4 + w) * BN + cl]; }
              if (b < mn) { mn = b; an = reinterpret_cast<int*>(comb)[(3 * 4 + w) * BN + cl]; }
            }
            if (d.gmax) d.gmax[(size_t)grow * Cout + n0 + cl] = mx;
            if (d.gmin) d.gmin[(size_t)grow * Cout + n0 + cl] = mn;
            if (d.garg_max) d.garg_max[(size_t)grow * Cout + n0 + cl] = ax - gi * g;
            if (d.garg_min) d.garg_min[(size_t)grow * Cout + n0 + cl] = an - gi * g;
          }
        }
        named_bar_sync(1, TC_EPI_WARPS * 32);          // comb is reused by the next tile
      }
    }
  }

  // ---------------------------------------------------------------------------------- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == TC_MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <int BN, int STAGES, bool COMBINE>
static int launch_tc(const usip_layer_desc& d, const uint32_t* wpack, cudaStream_t st) {
  using SM = TcSmem<BN, STAGES, COMBINE>;
  static_assert(SM::BYTES <= 232448, "shared memory budget");
  static int sm_count = 0;
  if (sm_count == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
    cudaError_t e = cudaFuncSetAttribute(layer_fwd_tc_kernel<BN, STAGES, COMBINE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::BYTES);
    if (e != cudaSuccess) { sm_count = 0; set_last_error("layer_fwd_tc smem attr"); return (int)e; }
  }
  const int m_tiles = cdiv(d.P, TC_BM), n_tiles = d.Cout / BN;
  const int grid = min(sm_count, m_tiles * n_tiles);
  layer_fwd_tc_kernel<BN, STAGES, COMBINE><<<grid, TC_THREADS, SM::BYTES, st>>>(d, wpack);
  return check_launch("layer_fwd_tc_kernel");
}

bool tc2_eligible(const usip_layer_desc& d);                                   // mlp_tc2.cu
int launch_tc2(const usip_layer_desc& d, const uint32_t* wpack, cudaStream_t st);

int tc_tile_n(int Cout) { return Cout % 256 == 0 ? 256 : (Cout % 128 == 0 ? 128 : 64); }
int tc_stat_rows() { return TC_STAT_ROWS; }

int layer_fwd_tc(const usip_layer_desc& d, cudaStream_t st) {
  USIP_REQUIRE(d.Cin % TC_BK == 0 && d.Cout % 64 == 0, "layer_fwd_tc: needs Cin%32==0 and Cout%64==0");
  USIP_REQUIRE((d.ldx % 4) == 0 && (reinterpret_cast<uintptr_t>(d.X) % 16) == 0, "layer_fwd_tc: X must be 16B aligned");
  USIP_REQUIRE(!d.Y || ((d.ldy % 4) == 0 && (reinterpret_cast<uintptr_t>(d.Y) % 16) == 0), "layer_fwd_tc: Y must be 16B aligned");
  USIP_REQUIRE(!d.addend || ((d.ld_add % 4) == 0 && (reinterpret_cast<uintptr_t>(d.addend) % 16) == 0), "layer_fwd_tc: addend alignment");
  USIP_REQUIRE(!d.bias || (reinterpret_cast<uintptr_t>(d.bias) % 16) == 0, "layer_fwd_tc: bias alignment");
  USIP_REQUIRE(!d.in_scale || ((reinterpret_cast<uintptr_t>(d.in_scale) % 16) == 0 && (reinterpret_cast<uintptr_t>(d.in_shift) % 16) == 0),
               "layer_fwd_tc: scale/shift alignment");
  USIP_REQUIRE(d.tc_workspace && d.tc_workspace_bytes >= (int64_t)2 * d.Cout * d.Cin * 4, "layer_fwd_tc: workspace too small");
  if (d.gmax || d.gmin) USIP_REQUIRE(d.group == 16 || d.group == 32 || d.group == 64 || d.group == 128, "layer_fwd_tc: group must be 16/32/64/128");
  int BN = tc_tile_n(d.Cout);
  if ((d.gmax || d.gmin) && d.group > 32 && BN > 128) BN = 128;       // cross-warp group combine needs the small tile
  // few row tiles (node-level GEMMs): narrower column tiles fill more SMs
  while (BN > 64 && (long long)cdiv(d.P, TC_BM) * (d.Cout / BN) < 120 && d.Cout % (BN / 2) == 0) BN /= 2;
  // precision 2 opts into the CTA-pair (cta_group::2) kernel for wide layers with enough tiles.  Measured on B200 it
  // is ~10-25% SLOWER than the single-CTA kernel with register prefetch, so it is not the default.
  const bool pair = d.precision == 2 && d.Cout % 256 == 0 && tc2_eligible(d);
  if (pair) BN = 256;
  uint32_t* wpack = reinterpret_cast<uint32_t*>(d.tc_workspace);
  if (!d.tc_weights_packed) {
    const int total = d.Cout * (d.Cin / 4);
    tc_pack_weights_kernel<<<cdiv(total, 256), 256, 0, st>>>(d.W, d.ldw, d.w_transposed, d.Cout, d.Cin, BN, wpack);
    int e = check_launch("tc_pack_weights_kernel");
    if (e) return e;
  }
  if (pair) return launch_tc2(d, wpack, st);
  const bool combine = (d.gmax || d.gmin) && d.group > 32;
  if (combine) return BN == 128 ? launch_tc<128, 2, true>(d, wpack, st) : launch_tc<64, 3, true>(d, wpack, st);
  if (BN == 256) return launch_tc<256, 2, false>(d, wpack, st);
  if (BN == 128) return launch_tc<128, 3, false>(d, wpack, st);
  return launch_tc<64, 4, false>(d, wpack, st);
}

}  // namespace usip
