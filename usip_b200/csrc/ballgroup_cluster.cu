// ballgroup_cluster.cu -- usip_ball_group_f32 as ONE kernel: a thread-block cluster of 8 CTAs owns one cloud.
//
// Reference: models/networks.py:355-373 (DescriptorLiteOld: ball query -> gather -> decenter) and
// models/ball_query_ext/ball_query_cuda.cu:22-46 (first K hits in index order, `<=`, cyclic pad).
//
// The cloud (N <= 16384 points, S <= 4 feature channels) is read ONCE: binned into a uniform grid with cell size
// h >= 1.001 r by a counting sort whose histogram / cursors and cell-sorted 32-byte records (x, y, z, index, features) are
// distributed over the shared memories of the 8 CTAs (remote atomics and stores over DSMEM); then every CTA answers 1/8 of
// the keypoints, one warp per keypoint: 9 x-contiguous cell ranges -> candidate records -> warp-ballot radius test -> hits
// ranked by point index (first K in index order, bit-exact with the in-order scan) -> cyclic pad + gather of the
// records + decentred group in one pass, all out of distributed shared memory.  Global memory sees only the compulsory
// traffic.  Balls with more than BC_CAP hits and degenerate grids (non-finite points) take the early-exit in-order scan
// over global memory instead.
#include "common.cuh"
#include <cstdio>
#include <cstdlib>

namespace usip {

constexpr int BC_CTAS = 8;
constexpr int BC_THREADS = 512;
constexpr int BC_WARPS = BC_THREADS / 32;
constexpr int BC_PTS = 2048;                       // points per CTA
constexpr int BC_PTS_LOG = 11;
constexpr int BC_PPT = BC_PTS / BC_THREADS;        // points per thread
constexpr int BC_CPC_MAX = 4096;                   // cells per CTA slice: the power of two >= cells / 8 (owner = cell >> log2)
constexpr int BC_CELLS = BC_CTAS * BC_CPC_MAX;     // grid cells per cloud
constexpr int BC_CAP = 64;                         // hits kept per keypoint before the in-order fallback; also max K
// 64 KB records + 16 KB cell table + 12 KB hit lists: two CTAs fit on an SM.  That matters because a B200 only hosts 15
// clusters of 8 single-occupancy CTAs at a time (cudaOccupancyMaxActiveClusters) -- a 16-cloud batch then ran as two waves.
constexpr int BC_SMEM = BC_PTS * 32 + BC_CPC_MAX * 4 + 3 * BC_WARPS * BC_CAP * 4;

struct BcGrid { float ox, oy, oz, inv_h; int nx, ny, nz, ok; };

// phase time stamps (clock64) of CTA 0 / thread 0 of the last launch: load, grid, histogram, scan, scatter, query, exit
__device__ unsigned long long bc_phase_clock[8];
#define BC_STAMP(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) bc_phase_clock[i] = (unsigned long long)clock64(); } while (0)

__device__ __forceinline__ uint32_t bc_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void bc_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// cluster-window address of `p` (a pointer into THIS CTA's shared memory) in CTA `rank`
__device__ __forceinline__ uint32_t bc_map(const void* p, uint32_t rank) {
  uint32_t a = (uint32_t)__cvta_generic_to_shared(p), r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
  return r;
}
__device__ __forceinline__ float4 bc_ld4(uint32_t a) {
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t bc_ld1(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void bc_st4(uint32_t a, float4 v) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ uint32_t bc_atom_add(uint32_t a, uint32_t v) {
  uint32_t old;
  asm volatile("atom.shared::cluster.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(a), "r"(v) : "memory");
  return old;
}
__device__ __forceinline__ void bc_red_add(uint32_t a, uint32_t v) {
  asm volatile("red.shared::cluster.add.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ int bc_cell1(float v, float o, float inv_h, int n) {
  const int c = (int)floorf((v - o) * inv_h);
  return min(max(c, 0), n - 1);
}

__global__ void __cluster_dims__(BC_CTAS, 1, 1) __launch_bounds__(BC_THREADS, 2)
bg_cluster_kernel(const float* __restrict__ xyz, const float* __restrict__ feat, const float* __restrict__ centers,
                  float radius, float t_max, int32_t* __restrict__ out_idx, float* __restrict__ out_group,
                  float* __restrict__ out_rows, int ld_rows, int S, int N, int M, int K) {
  extern __shared__ __align__(16) uint8_t bc_smem[];
  float4* srec = reinterpret_cast<float4*>(bc_smem);                // [BC_PTS][2]: slots [rank*2048, +2048) of the CELL-SORTED
                                                                    // cloud: (x, y, z, point index) (f0, f1, f2, f3)
  int* table = reinterpret_cast<int*>(srec + 2 * BC_PTS);           // [cpc] cells [rank*cpc, +cpc): count -> cursor -> end,
                                                                    // relative to this slice's first slot (sbase[rank])
  int* hits = table + BC_CPC_MAX;                                   // [warps][3][BC_CAP]: slots and point indices in discovery order, slots in index order
  __shared__ float sbb[8];                                          // this CTA's min xyz, max xyz, non-finite flag
  __shared__ float wmin[3][BC_WARPS], wmax[3][BC_WARPS];
  __shared__ int wbad[BC_WARPS], wsum[BC_WARPS];
  __shared__ int s_tot;                                             // points binned into this CTA's cell slice
  __shared__ int sbase[BC_CTAS + 1];                                // first slot of every CTA's cell slice
  __shared__ BcGrid sg;

  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const uint32_t rank = bc_rank();
  const int b = blockIdx.x / BC_CTAS;
  const float* p = xyz + (size_t)b * 3 * N;

  BC_STAMP(0);
  // ---------------------------------------------------------------- phase 1: my 2048 points -> registers, bounding box
  float px[BC_PPT], py[BC_PPT], pz[BC_PPT];
  bool pok[BC_PPT];
  float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  bool bad = false;
#pragma unroll
  for (int j = 0; j < BC_PPT; ++j) {
    const int n = (int)rank * BC_PTS + tid + j * BC_THREADS;
    pok[j] = n < N;
    px[j] = py[j] = pz[j] = 0.f;
    if (pok[j]) { px[j] = p[n]; py[j] = p[N + n]; pz[j] = p[2 * N + n]; }
  }
  for (int i = tid; i < BC_CPC_MAX; i += BC_THREADS) table[i] = 0;
#pragma unroll
  for (int j = 0; j < BC_PPT; ++j)
    if (pok[j]) {
      const float x = px[j], y = py[j], z = pz[j];
      bad |= !(fabsf(x) <= 1e30f) || !(fabsf(y) <= 1e30f) || !(fabsf(z) <= 1e30f);
      mn[0] = fminf(mn[0], x); mx[0] = fmaxf(mx[0], x); mn[1] = fminf(mn[1], y); mx[1] = fmaxf(mx[1], y);
      mn[2] = fminf(mn[2], z); mx[2] = fmaxf(mx[2], z);
    }
  {
    const unsigned anybad = __ballot_sync(0xffffffffu, bad);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        mn[c] = fminf(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], o));
        mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o));
      }
      if (lane == 0) { wmin[c][w] = mn[c]; wmax[c][w] = mx[c]; }
    }
    if (lane == 0) wbad[w] = anybad != 0;
  }
  __syncthreads();
  if (tid < 3) {
    float lo = INFINITY, hi = -INFINITY;
    for (int i = 0; i < BC_WARPS; ++i) { lo = fminf(lo, wmin[tid][i]); hi = fmaxf(hi, wmax[tid][i]); }
    sbb[tid] = lo; sbb[3 + tid] = hi;
  }
  if (tid == 3) { int any = 0; for (int i = 0; i < BC_WARPS; ++i) any |= wbad[i]; sbb[6] = any ? 1.f : 0.f; }
  bc_sync();                                                        // every CTA's box is visible, every table is zero
  BC_STAMP(1);
  if (tid == 0) {
    BcGrid g; g.ok = 1;
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (uint32_t r = 0; r < BC_CTAS; ++r) {
      for (int c = 0; c < 3; ++c) {
        lo[c] = fminf(lo[c], __uint_as_float(bc_ld1(bc_map(&sbb[c], r))));
        hi[c] = fmaxf(hi[c], __uint_as_float(bc_ld1(bc_map(&sbb[3 + c], r))));
      }
      if (__uint_as_float(bc_ld1(bc_map(&sbb[6], r))) != 0.f) g.ok = 0;
    }
    if (!(radius >= 0.f) || !(radius <= 1e30f)) g.ok = 0;
    const float ext = fmaxf(fmaxf(hi[0] - lo[0], hi[1] - lo[1]), hi[2] - lo[2]);
    float h = fmaxf(fmaxf(radius * 1.001f, ext * (1.0f / 126.0f)), 1e-6f);
    int nx = 1, ny = 1, nz = 1;
    if (g.ok) {
      for (int it = 0; it < 64; ++it) {
        nx = (int)floorf((hi[0] - lo[0]) / h) + 1; ny = (int)floorf((hi[1] - lo[1]) / h) + 1; nz = (int)floorf((hi[2] - lo[2]) / h) + 1;
        if ((long long)nx * ny * nz <= BC_CELLS) break;
        h *= 1.26f;
      }
      if ((long long)nx * ny * nz > BC_CELLS) g.ok = 0;
    }
    g.ox = lo[0]; g.oy = lo[1]; g.oz = lo[2]; g.inv_h = 1.0f / h; g.nx = nx; g.ny = ny; g.nz = nz;
    sg = g;
  }
  __syncthreads();
  const BcGrid g = sg;                                              // identical in all 8 CTAs
  BC_STAMP(2);
  const int cells = g.ok ? g.nx * g.ny * g.nz : 0;
  int lg = 0;                                                       // cells per CTA slice = 1 << lg  (8 << lg >= cells)
  while ((BC_CTAS << lg) < cells) ++lg;
  const int cpc = 1 << lg;

  if (g.ok) {
    // -------------------------------------------------------------- phase 2: distributed histogram
    uint32_t caddr[BC_PPT];
    int owner[BC_PPT];
#pragma unroll
    for (int j = 0; j < BC_PPT; ++j) {
      caddr[j] = 0; owner[j] = 0;
      if (pok[j]) {
        const int cell = (bc_cell1(pz[j], g.oz, g.inv_h, g.nz) * g.ny + bc_cell1(py[j], g.oy, g.inv_h, g.ny)) * g.nx + bc_cell1(px[j], g.ox, g.inv_h, g.nx);
        owner[j] = cell >> lg;
        caddr[j] = bc_map(&table[cell & (cpc - 1)], (uint32_t)owner[j]);
        bc_red_add(caddr[j], 1u);
      }
    }
    // the feature channels are only needed for the scatter: fetch them now, they arrive during the scan
    float4 fj[BC_PPT];
#pragma unroll
    for (int j = 0; j < BC_PPT; ++j) {
      float f[4] = {0.f, 0.f, 0.f, 0.f};
      if (pok[j]) {
        const int n = (int)rank * BC_PTS + tid + j * BC_THREADS;
        for (int c = 0; c < S && c < 4; ++c) f[c] = __ldg(feat + ((size_t)b * S + c) * N + n);
      }
      fj[j] = make_float4(f[0], f[1], f[2], f[3]);
    }
    bc_sync();
    BC_STAMP(3);
    // -------------------------------------------------------------- phase 3: exclusive scan of my slice
    const int per = (cpc + BC_THREADS - 1) / BC_THREADS, lo = min(tid * per, cpc), hi = min(lo + per, cpc);
    int s = 0;
    for (int i = lo; i < hi; ++i) s += table[i];
    int incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    if (lane == 31) wsum[w] = incl;
    __syncthreads();
    if (w == 0) {
      const int v = lane < BC_WARPS ? wsum[lane] : 0;
      int iv = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, iv, o); if (lane >= o) iv += u; }
      if (lane < BC_WARPS) wsum[lane] = iv - v;                     // exclusive warp offsets
      if (lane == 31) s_tot = iv;
    }
    __syncthreads();
    int run = wsum[w] + incl - s;
    for (int i = lo; i < hi; ++i) { const int v = table[i]; table[i] = run; run += v; }
    bc_sync();                                                      // slice-relative cursors and slice totals are final
    if (w == 0) {
      const int v = lane < BC_CTAS ? (int)bc_ld1(bc_map(&s_tot, (uint32_t)lane)) : 0;
      int iv = v;
#pragma unroll
      for (int o = 1; o < 16; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, iv, o); if (lane >= o) iv += u; }
      if (lane <= BC_CTAS) sbase[lane] = iv - v;                    // sbase[8] = N
    }
    __syncthreads();
    BC_STAMP(4);
    // -------------------------------------------------------------- phase 4: scatter into the cell-sorted record array
#pragma unroll
    for (int j = 0; j < BC_PPT; ++j)
      if (pok[j]) {
        const uint32_t pos = (uint32_t)sbase[owner[j]] + bc_atom_add(caddr[j], 1u);
        const int n = (int)rank * BC_PTS + tid + j * BC_THREADS;
        const uint32_t a = bc_map(&srec[2 * (pos & (BC_PTS - 1))], pos >> BC_PTS_LOG);
        bc_st4(a, make_float4(px[j], py[j], pz[j], __int_as_float(n)));
        bc_st4(a + 16, fj[j]);
      }
  }
  bc_sync();                                                        // table[c] is now the (slice-relative) END of cell c
  BC_STAMP(5);

  // ---------------------------------------------------------------- phase 5: one warp per keypoint
  const int Mc = (M + BC_CTAS - 1) / BC_CTAS;
  const int m_lo = (int)rank * Mc, m_end = min(M, m_lo + Mc);
  int* hl = hits + w * (3 * BC_CAP);                                // hit slots in discovery order
  int* hn = hl + BC_CAP;                                            // their point indices
  int* sl = hn + BC_CAP;                                            // hit slots in ascending point index (the reference's order)
  const unsigned lt = (1u << lane) - 1u;
  const int C = 3 + S;
  auto cell_end = [&](int c) -> int {                               // end of cell c == start of cell c+1 (global slot)
    const int ow = c >> lg;
    return sbase[ow] + (int)bc_ld1(bc_map(&table[c & (cpc - 1)], (uint32_t)ow));
  };
  auto slot_addr = [&](int slot) -> uint32_t { return bc_map(&srec[2 * (slot & (BC_PTS - 1))], (uint32_t)(slot >> BC_PTS_LOG)); };
  const float* cp = centers + (size_t)b * 3 * M;
  const float* pf = feat + (size_t)b * S * N;
  for (int mb = m_lo + w; mb < m_end; mb += BC_WARPS * 32) {
    // lane l holds the centre of this warp's l-th keypoint of the block: one global round trip for up to 32 keypoints
    const int ml = mb + lane * BC_WARPS;
    float lcx = 0.f, lcy = 0.f, lcz = 0.f;
    if (ml < m_end) { lcx = cp[ml]; lcy = cp[M + ml]; lcz = cp[2 * M + ml]; }
    for (int t = 0; t < 32; ++t) {
      const int m = mb + t * BC_WARPS;
      if (m >= m_end) break;
      const float cx = __shfl_sync(0xffffffffu, lcx, t), cy = __shfl_sync(0xffffffffu, lcy, t), cz = __shfl_sync(0xffffffffu, lcz, t);
      const size_t row = (size_t)b * M + m;
      int cnt = 0;
      bool brute = !g.ok;
      if (!brute) {
        const int kx = (int)floorf((cx - g.ox) * g.inv_h), ky = (int)floorf((cy - g.oy) * g.inv_h), kz = (int)floorf((cz - g.oz) * g.inv_h);
        const bool cfin = fabsf(cx) <= 1e30f && fabsf(cy) <= 1e30f && fabsf(cz) <= 1e30f;
        const int x0 = max(kx - 1, 0), x1 = min(kx + 1, g.nx - 1);
        int rs = 0, rn = 0;                                         // lanes 0..8: the 9 x-contiguous ranges
        if (lane < 9 && cfin && x0 <= x1) {
          const int y = ky + (lane % 3) - 1, z = kz + (lane / 3) - 1;
          if (y >= 0 && y < g.ny && z >= 0 && z < g.nz) {
            const int base = (z * g.ny + y) * g.nx;
            rs = (base + x0) > 0 ? cell_end(base + x0 - 1) : 0;
            rn = cell_end(base + x1) - rs;
          }
        }
        int incl = rn;
#pragma unroll
        for (int o2 = 1; o2 < 16; o2 <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o2); if (lane >= o2) incl += v; }
        const int total = __shfl_sync(0xffffffffu, incl, 8);
        int pre[9], st[9];
#pragma unroll
        for (int r = 0; r < 9; ++r) { pre[r] = __shfl_sync(0xffffffffu, incl - rn, r); st[r] = __shfl_sync(0xffffffffu, rs, r); }
        for (int base = 0; base < total; base += 32) {
          const int j = base + lane;
          bool hit = false; int src = 0, n = 0;
          if (j < total) {
            src = st[0] + j;
#pragma unroll
            for (int r = 1; r < 9; ++r) if (j >= pre[r]) src = st[r] + (j - pre[r]);
            const float4 q = bc_ld4(slot_addr(src));
            n = __float_as_int(q.w);
            hit = sqdist_rn(cx, cy, cz, q.x, q.y, q.z) <= t_max;
          }
          const unsigned bal = __ballot_sync(0xffffffffu, hit);
          if (bal) {
            const int pos = cnt + __popc(bal & lt);
            if (hit && pos < BC_CAP) { hl[pos] = src; hn[pos] = n; }
            cnt += __popc(bal);
            if (cnt > BC_CAP) { brute = true; break; }
          }
        }
      }
      __syncwarp();
      const int u = min(cnt, K);
      int32_t* o = out_idx + row * K;
      if (brute) {
        // early-exit in-order scan (exactly the reference loop) straight from global memory; used for balls with more
        // than BC_CAP hits and for degenerate grids
        cnt = 0;
        for (int base = 0; base < N && cnt < K; base += 32) {
          const int n = base + lane;
          const bool hit = (n < N) && (sqdist_rn(cx, cy, cz, __ldg(p + n), __ldg(p + N + n), __ldg(p + 2 * N + n)) <= t_max);
          const unsigned bal = __ballot_sync(0xffffffffu, hit);
          if (bal) {
            const int pos = cnt + __popc(bal & lt);
            if (hit && pos < K) sl[pos] = n;                        // point indices, not slots
            cnt += __popc(bal);
          }
        }
      } else {
        // order by point index: rank = number of hits with a smaller index (indices are distinct).  cnt is ~3 on
        // LiDAR-density clouds, so this is a handful of broadcast shared loads where a bitonic network costs ~300 instructions
        for (int i = lane; i < cnt; i += 32) {
          const int nmine = hn[i];
          int rk = 0;
          for (int j = 0; j < cnt; ++j) rk += hn[j] < nmine ? 1 : 0;
          sl[rk] = hl[i];
        }
      }
      __syncwarp();
      // out[k] = hits[k % u] (first u in index order, then the cyclic pad of ball_query_cuda.cu:40-46); no hit -> point 0.
      // The gather and the decentred group are written in the same pass.
      const int uu = brute ? min(cnt, K) : u;
      const uint32_t um = uu > 1 ? (0xffffffffu / (uint32_t)uu + 1u) : 0u;   // uu >= 2: k % uu == k - uu * umulhi(k, ceil(2^32 / uu))
      for (int k = lane; k < K; k += 32) {
        const int e = uu > 1 ? sl[k - uu * (int)__umulhi((uint32_t)k, um)] : (uu == 1 ? sl[0] : -1);
        int n;
        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (!brute && e >= 0) {                                     // e is a slot of the cell-sorted record array
          const uint32_t ra = slot_addr(e);
          const float4 r0 = bc_ld4(ra), r1 = bc_ld4(ra + 16);
          n = __float_as_int(r0.w);
          v[0] = r0.x - cx; v[1] = r0.y - cy; v[2] = r0.z - cz; v[3] = r1.x; v[4] = r1.y; v[5] = r1.z; v[6] = r1.w;   // networks.py:373
        } else {                                                    // e is a point index (or "none": the reference gathers point 0)
          n = e >= 0 ? e : 0;
          v[0] = __ldg(p + n) - cx; v[1] = __ldg(p + N + n) - cy; v[2] = __ldg(p + 2 * N + n) - cz;
          for (int c = 0; c < S && c < 4; ++c) v[3 + c] = __ldg(pf + (size_t)c * N + n);
        }
        o[k] = n;
        if (out_group) {
          float* gp = out_group + ((size_t)b * C * M + m) * K + k;
          for (int c = 0; c < C; ++c) gp[(size_t)c * M * K] = v[c];
        }
        if (out_rows) {
          float* rowp = out_rows + (row * K + k) * ld_rows;
          if (ld_rows == 8) {
            reinterpret_cast<float4*>(rowp)[0] = make_float4(v[0], v[1], v[2], v[3]);
            reinterpret_cast<float4*>(rowp)[1] = make_float4(v[4], v[5], v[6], v[7]);
          } else {
            for (int c = 0; c < ld_rows; ++c) rowp[c] = c < C ? v[c] : 0.f;
          }
        }
      }
      __syncwarp();                                                 // hl / sl are reused by the next keypoint
    }
  }
  BC_STAMP(6);
  bc_sync();                                                        // nobody exits while its shared memory is still being read
  BC_STAMP(7);
}

bool bg_cluster_eligible(int S, int N, int K) { return S <= 4 && N <= BC_CTAS * BC_PTS && K <= BC_CAP; }

int launch_bg_cluster(const float* xyz, const float* feat, const float* centers, float radius, float t_max,
                      int32_t* out_idx, float* out_group, float* out_rows, int ld_rows, int B, int S, int N, int M, int K,
                      cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(bg_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BC_SMEM);
    if (e != cudaSuccess) { set_last_error("bg_cluster smem attr"); return (int)e; }
    attr = true;
    if (getenv("USIP_BG_DEBUG")) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(B * BC_CTAS); cfg.blockDim = dim3(BC_THREADS); cfg.dynamicSmemBytes = BC_SMEM;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = BC_CTAS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int nc = -1;
      cudaError_t e2 = cudaOccupancyMaxActiveClusters(&nc, bg_cluster_kernel, &cfg);
      fprintf(stderr, "[usip] bg_cluster: max active clusters = %d (err %d)\n", nc, (int)e2);
    }
  }
  bg_cluster_kernel<<<B * BC_CTAS, BC_THREADS, BC_SMEM, st>>>(xyz, feat, centers, radius, t_max, out_idx, out_group, out_rows,
                                                              ld_rows, S, N, M, K);
  return check_launch("bg_cluster_kernel");
}

int bg_cluster_phase_clocks(unsigned long long* host8) {
  return (int)cudaMemcpyFromSymbol(host8, bc_phase_clock, sizeof(unsigned long long) * 8);
}

}  // namespace usip
