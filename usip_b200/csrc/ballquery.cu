// ballquery.cu -- ball query operators.
//   usip_ball_query_dist_f32 : drop-in for ball_query.forward_cuda_shared_mem on a pre-computed
//                              distance matrix (models/ball_query_ext/ball_query_cuda.cu:10-49)
//   usip_ball_group_f32      : fused distance + ball query + gather + decentre from xyz
//                              (models/networks.py:355-373), never materialising (B,M,N).
// Both keep the reference's order-dependent semantics: FIRST K hits in ascending point index,
// `<=` on the sqrt distance, 0 hits -> zeros, u<K hits -> cyclic repeat out[u+i] = out[i % u].
#include "common.cuh"
#include <cstdlib>

namespace usip {

__device__ __forceinline__ void ball_pad(int32_t* o, int cnt, int K, int lane) {
  __syncwarp();
  const int u = min(cnt, K);
  if (u == 0) {
    for (int i = lane; i < K; i += 32) o[i] = 0;
  } else if (u < K) {
    // out[u+i] = out[i % u]: the source is always one of the first u (already final) entries
    for (int i = lane; i < K - u; i += 32) o[u + i] = o[i % u];
  }
  __syncwarp();
}

// one warp per (b,m) row of the distance matrix; 4 x 32 coalesced elements per iteration
__global__ void __launch_bounds__(256)
ball_query_dist_kernel(const float* __restrict__ dist, float radius, int32_t* __restrict__ out,
                       int rows, int N, int K) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= rows) return;
  const float* row = dist + (size_t)w * N;
  int32_t* o = out + (size_t)w * K;
  const unsigned lt = (1u << lane) - 1u;
  int cnt = 0;
  for (int base = 0; base < N && cnt < K; base += 128) {
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { int n = base + j * 32 + lane; v[j] = n < N ? __ldg(row + n) : INFINITY; }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = base + j * 32 + lane;
      bool hit = (n < N) && (v[j] <= radius);
      unsigned bal = __ballot_sync(0xffffffffu, hit);
      if (bal) {
        int pos = cnt + __popc(bal & lt);
        if (hit && pos < K) o[pos] = n;
        cnt += __popc(bal);
      }
    }
  }
  ball_pad(o, cnt, K, lane);
}

// fused: brute-force scan of the cloud in index order.  d^2 is formed with the reference's fp32 op
// order and compared against t_max = max{t : sqrtf(t) <= radius} (exactly equivalent to
// sqrtf(d2) <= radius because correctly-rounded sqrt is monotone), so no sqrt per pair.
__global__ void __launch_bounds__(256)
ball_group_brute_kernel(const float* __restrict__ xyz, const float* __restrict__ feat,
                        const float* __restrict__ centers, float t_max, int32_t* __restrict__ out_idx,
                        float* __restrict__ out_group, float* __restrict__ out_rows, int ld_rows, int B, int S, int N,
                        int M, int K) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= B * M) return;
  const int b = w / M, m = w - b * M;
  const float* px = xyz + (size_t)b * 3 * N;
  const float cx = centers[(size_t)b * 3 * M + m], cy = centers[(size_t)b * 3 * M + M + m],
              cz = centers[(size_t)b * 3 * M + 2 * M + m];
  int32_t* o = out_idx + (size_t)w * K;
  const unsigned lt = (1u << lane) - 1u;
  int cnt = 0;
  for (int base = 0; base < N && cnt < K; base += 128) {
    float d[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = base + j * 32 + lane;
      d[j] = n < N ? sqdist_rn(cx, cy, cz, __ldg(px + n), __ldg(px + N + n), __ldg(px + 2 * N + n)) : INFINITY;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = base + j * 32 + lane;
      bool hit = (n < N) && (d[j] <= t_max);
      unsigned bal = __ballot_sync(0xffffffffu, hit);
      if (bal) {
        int pos = cnt + __popc(bal & lt);
        if (hit && pos < K) o[pos] = n;
        cnt += __popc(bal);
      }
    }
  }
  ball_pad(o, cnt, K, lane);
  if (!out_group && !out_rows) return;
  const int C = 3 + S;
  const float ctr[3] = {cx, cy, cz};
  for (int k = lane; k < K; k += 32) {
    int n = o[k];
    float* rowp = out_rows ? out_rows + ((size_t)w * K + k) * ld_rows : nullptr;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float v = px[(size_t)c * N + n] - ctr[c];                                            // networks.py:373
      if (out_group) out_group[(((size_t)b * C + c) * M + m) * K + k] = v;
      if (rowp) rowp[c] = v;
    }
    for (int c = 0; c < S; ++c) {
      float v = feat[((size_t)b * S + c) * N + n];
      if (out_group) out_group[(((size_t)b * C + 3 + c) * M + m) * K + k] = v;
      if (rowp) rowp[3 + c] = v;
    }
    if (rowp) for (int c = C; c < ld_rows; ++c) rowp[c] = 0.f;
  }
}

// ------------------------------------------------------------------------------------------------
// Cell-binned variant: points are binned once per call into a uniform grid with cell size h >= 1.001*r by a counting
// sort on ALL SMs (global-memory atomics, L2-resident scratch); a keypoint then only tests the <= 27 neighbouring cells
// (9 x-contiguous ranges of the cell-sorted 32-byte records), ranks the hits by point index and keeps the first K --
// bit-identical to the in-order scan, without the O(N) pass per keypoint.  Balls with more than BG_CAP hits fall back
// to the early-exit in-order scan, which is cheap exactly there.
//   bg_bbox_kernel    per 1024-point chunk: partial bounding box (+ zeroes its share of the cell table)
//   bg_hist_kernel    reduces the partial boxes -> grid; cell histogram (atomicAdd); the last CTA of a cloud to arrive
//                     scans it: table = cell start (the scatter cursor)
//   bg_scatter_kernel cursor atomicAdd -> cell-sorted records (x,y,z,index | f0..f3); afterwards table[c] = END of cell c
//   bg_query_kernel   one warp per keypoint
// ------------------------------------------------------------------------------------------------
constexpr int BG_MAX_CELLS = 49152;
constexpr int BG_CAP = 64;                        // hits kept per keypoint before the in-order fallback
constexpr int BG_CHUNK = 1024;                    // points per CTA in the per-point kernels (256 threads x 4)

struct BgGrid { float ox, oy, oz, inv_h; int nx, ny, nz, ok; };

__device__ __forceinline__ int bg_cell1(float v, float o, float inv_h, int n) {
  int c = (int)floorf((v - o) * inv_h);
  return min(max(c, 0), n - 1);
}
__device__ __forceinline__ int bg_cell(const BgGrid& g, float x, float y, float z) {
  return (bg_cell1(z, g.oz, g.inv_h, g.nz) * g.ny + bg_cell1(y, g.oy, g.inv_h, g.ny)) * g.nx + bg_cell1(x, g.ox, g.inv_h, g.nx);
}

__global__ void __launch_bounds__(256)
bg_bbox_kernel(const float* __restrict__ xyz, float* __restrict__ part, int32_t* __restrict__ table, int32_t* __restrict__ done,
               int N, int nch) {
  __shared__ float smin[3][8], smax[3][8];
  __shared__ int sbad[8];
  const int b = blockIdx.y, ch = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const float* p = xyz + (size_t)b * 3 * N;
  float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  bool bad = false;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int n = ch * BG_CHUNK + j * 256 + tid;
    if (n < N) {
      const float x = p[n], y = p[N + n], z = p[2 * N + n];
      bad |= !(fabsf(x) <= 1e30f) || !(fabsf(y) <= 1e30f) || !(fabsf(z) <= 1e30f);
      mn[0] = fminf(mn[0], x); mx[0] = fmaxf(mx[0], x); mn[1] = fminf(mn[1], y); mx[1] = fmaxf(mx[1], y);
      mn[2] = fminf(mn[2], z); mx[2] = fmaxf(mx[2], z);
    }
  }
  if (ch == 0 && tid == 0) done[b] = 0;                 // arrival counter of the histogram CTAs (last one scans)
  // this CTA's share of the cloud's cell table
  {
    int32_t* t = table + (size_t)b * (BG_MAX_CELLS + 1);
    const int per = (BG_MAX_CELLS + 1 + nch - 1) / nch, lo = ch * per, hi = min(lo + per, BG_MAX_CELLS + 1);
    for (int i = lo + tid; i < hi; i += 256) t[i] = 0;
  }
  const unsigned anybad = __ballot_sync(0xffffffffu, bad);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn[c] = fminf(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], o));
      mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o));
    }
    if (lane == 0) { smin[c][w] = mn[c]; smax[c][w] = mx[c]; }
  }
  if (lane == 0) sbad[w] = anybad != 0;
  __syncthreads();
  if (tid < 8) {
    float* o = part + ((size_t)b * nch + ch) * 8;
    if (tid < 3) { float v = INFINITY; for (int i = 0; i < 8; ++i) v = fminf(v, smin[tid][i]); o[tid] = v; }
    else if (tid < 6) { float v = -INFINITY; for (int i = 0; i < 8; ++i) v = fmaxf(v, smax[tid - 3][i]); o[tid] = v; }
    else if (tid == 6) { int any = 0; for (int i = 0; i < 8; ++i) any |= sbad[i]; o[6] = any ? 1.f : 0.f; }
    else o[7] = 0.f;
  }
}

// every CTA of a cloud derives the same grid from the partial boxes (deterministic, no atomics, no extra launch)
__device__ __forceinline__ BgGrid bg_make_grid(const float* __restrict__ part, int nch, float radius, int lane) {
  float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
  float bad = 0.f;
  for (int i = lane; i < nch; i += 32) {
    const float4 a = *reinterpret_cast<const float4*>(part + (size_t)i * 8), c = *reinterpret_cast<const float4*>(part + (size_t)i * 8 + 4);
    lo[0] = fminf(lo[0], a.x); lo[1] = fminf(lo[1], a.y); lo[2] = fminf(lo[2], a.z);
    hi[0] = fmaxf(hi[0], a.w); hi[1] = fmaxf(hi[1], c.x); hi[2] = fmaxf(hi[2], c.y);
    bad = fmaxf(bad, c.z);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      lo[c] = fminf(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], o));
      hi[c] = fmaxf(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], o));
    }
    bad = fmaxf(bad, __shfl_xor_sync(0xffffffffu, bad, o));
  }
  BgGrid g; g.ok = bad == 0.f;
  if (!(radius >= 0.f) || !(radius <= 1e30f)) g.ok = 0;
  const float ext = fmaxf(fmaxf(hi[0] - lo[0], hi[1] - lo[1]), hi[2] - lo[2]);
  float h = fmaxf(fmaxf(radius * 1.001f, ext * (1.0f / 126.0f)), 1e-6f);
  int nx = 1, ny = 1, nz = 1;
  if (g.ok) {
    for (int it = 0; it < 64; ++it) {
      nx = (int)floorf((hi[0] - lo[0]) / h) + 1; ny = (int)floorf((hi[1] - lo[1]) / h) + 1; nz = (int)floorf((hi[2] - lo[2]) / h) + 1;
      if ((long long)nx * ny * nz <= BG_MAX_CELLS) break;
      h *= 1.26f;
    }
    if ((long long)nx * ny * nz > BG_MAX_CELLS) g.ok = 0;
  }
  g.ox = lo[0]; g.oy = lo[1]; g.oz = lo[2]; g.inv_h = 1.0f / h; g.nx = nx; g.ny = ny; g.nz = nz;
  return g;
}

// exclusive scan of one cloud's cell histogram by one 1024-thread CTA: all tiles (4096 cells each, coalesced) are loaded
// up front, so the CTA pays one memory round trip and two barriers
constexpr int BG_SCAN_TILES = BG_MAX_CELLS / 4096;      // 12
__device__ void bg_scan_cta(int32_t* __restrict__ t, int cells, int (*wtot)[32]) {
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  int v[BG_SCAN_TILES][4], incl[BG_SCAN_TILES];
#pragma unroll
  for (int k = 0; k < BG_SCAN_TILES; ++k) {
    const int i0 = k * 4096 + tid * 4;
#pragma unroll
    for (int j = 0; j < 4; ++j) v[k][j] = (i0 + j) < cells ? __ldcg(t + i0 + j) : 0;      // L2: written by other CTAs' atomics
  }
#pragma unroll
  for (int k = 0; k < BG_SCAN_TILES; ++k) {
    int x = v[k][0] + v[k][1] + v[k][2] + v[k][3];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += u; }
    incl[k] = x;
    if (lane == 31) wtot[k][w] = x;
  }
  __syncthreads();
  if (w == 0) {                                         // 12 x 32 warp totals in (tile, warp) order -> exclusive prefix
    int carry = 0;
#pragma unroll
    for (int k = 0; k < BG_SCAN_TILES; ++k) {
      const int x = wtot[k][lane];
      int ix = x;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, ix, o); if (lane >= o) ix += u; }
      wtot[k][lane] = carry + ix - x;
      carry += __shfl_sync(0xffffffffu, ix, 31);
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < BG_SCAN_TILES; ++k) {
    const int i0 = k * 4096 + tid * 4;
    int run = wtot[k][w] + incl[k] - (v[k][0] + v[k][1] + v[k][2] + v[k][3]);
#pragma unroll
    for (int j = 0; j < 4; ++j) { if ((i0 + j) < cells) t[i0 + j] = run; run += v[k][j]; }
  }
}

// cell histogram, one point per thread; the LAST CTA of a cloud to finish (arrival counter, no spinning) turns the
// histogram into cell starts -- the scatter cursor -- which saves a launch (~7 us of fixed cost at this problem size)
__global__ void __launch_bounds__(1024)
bg_hist_kernel(const float* __restrict__ xyz, const float* __restrict__ part, float radius, BgGrid* __restrict__ grids,
               int32_t* __restrict__ table, int32_t* __restrict__ done, int N, int nch) {
  __shared__ BgGrid sg;
  __shared__ int wtot[BG_SCAN_TILES][32];
  __shared__ int s_last;
  const int b = blockIdx.y, tid = threadIdx.x;
  if (tid < 32) {
    const BgGrid g = bg_make_grid(part + (size_t)b * nch * 8, nch, radius, tid);
    if (tid == 0) { sg = g; if (blockIdx.x == 0) grids[b] = g; }
  }
  __syncthreads();
  const BgGrid g = sg;
  if (!g.ok) return;
  const float* p = xyz + (size_t)b * 3 * N;
  int32_t* t = table + (size_t)b * (BG_MAX_CELLS + 1);
  const int n = blockIdx.x * 1024 + tid;
  if (n < N) atomicAdd(&t[bg_cell(g, p[n], p[N + n], p[2 * N + n])], 1);
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = atomicAdd(&done[b], 1) == (int)gridDim.x - 1;
  __syncthreads();
  if (s_last) {
    __threadfence();
    bg_scan_cta(t, g.nx * g.ny * g.nz, wtot);
  }
}

__global__ void __launch_bounds__(256)
bg_scatter_kernel(const float* __restrict__ xyz, const float* __restrict__ feat, const BgGrid* __restrict__ grids,
                  int32_t* __restrict__ table, float4* __restrict__ srec, int S, int N) {
  const int b = blockIdx.y, n = blockIdx.x * 256 + threadIdx.x;
  const BgGrid g = grids[b];
  if (!g.ok || n >= N) return;
  const float* p = xyz + (size_t)b * 3 * N;
  int32_t* t = table + (size_t)b * (BG_MAX_CELLS + 1);
  float4* sp = srec + (size_t)b * N * 2;
  const float x = p[n], y = p[N + n], z = p[2 * N + n];
  float f[4] = {0.f, 0.f, 0.f, 0.f};
  for (int c = 0; c < S && c < 4; ++c) f[c] = feat[((size_t)b * S + c) * N + n];
  const int pos = atomicAdd(&t[bg_cell(g, x, y, z)], 1);
  sp[2 * pos] = make_float4(x, y, z, __int_as_float(n));
  sp[2 * pos + 1] = make_float4(f[0], f[1], f[2], f[3]);
}

__global__ void __launch_bounds__(256, 5)
bg_query_kernel(const float* __restrict__ xyz, const float* __restrict__ feat, const float* __restrict__ centers,
                const BgGrid* __restrict__ grids, const int32_t* __restrict__ table, const float4* __restrict__ srec,
                float t_max, int32_t* __restrict__ out_idx, float* __restrict__ out_group, float* __restrict__ out_rows,
                int ld_rows, int B, int S, int N, int M, int K) {
  __shared__ int hits[8][2][BG_CAP];                       // point indices in discovery order; discovery positions in index order
  __shared__ float4 hrec[8][2][BG_CAP];                    // the hits' records (x,y,z,n | f0..f3): no second trip to memory
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int w = blockIdx.x * 8 + wib;
  if (w >= B * M) return;
  const int b = w / M, m = w - b * M;
  const BgGrid g = grids[b];
  const float* cp = centers + (size_t)b * 3 * M;
  const float cx = cp[m], cy = cp[M + m], cz = cp[2 * M + m];
  const float* p = xyz + (size_t)b * 3 * N;
  const float4* sp = srec + (size_t)b * N * 2;
  int* hn = hits[wib][0]; int* sl = hits[wib][1];
  float4* h0 = hrec[wib][0]; float4* h1 = hrec[wib][1];
  const unsigned lt = (1u << lane) - 1u;
  const int C = 3 + S;
  int cnt = 0;
  bool brute = !g.ok || K > BG_CAP;
  if (!brute) {
    // un-clamped cell coordinates of the centre; neighbour ranges are clipped to the grid
    const int kx = (int)floorf((cx - g.ox) * g.inv_h), ky = (int)floorf((cy - g.oy) * g.inv_h), kz = (int)floorf((cz - g.oz) * g.inv_h);
    const bool cfin = fabsf(cx) <= 1e30f && fabsf(cy) <= 1e30f && fabsf(cz) <= 1e30f;
    const int x0 = max(kx - 1, 0), x1 = min(kx + 1, g.nx - 1);
    const int32_t* cs = table + (size_t)b * (BG_MAX_CELLS + 1);   // after the scatter: cs[c] = END of cell c
    int rs = 0, rn = 0;                                           // lanes 0..8: the 9 x-contiguous ranges
    if (lane < 9 && cfin && x0 <= x1) {
      const int y = ky + (lane % 3) - 1, z = kz + (lane / 3) - 1;
      if (y >= 0 && y < g.ny && z >= 0 && z < g.nz) {
        const int base = (z * g.ny + y) * g.nx;
        rs = (base + x0) > 0 ? __ldg(cs + base + x0 - 1) : 0;
        rn = __ldg(cs + base + x1) - rs;
      }
    }
    int incl = rn;
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    const int total = __shfl_sync(0xffffffffu, incl, 8);
    int pre[9], st[9];
#pragma unroll
    for (int r = 0; r < 9; ++r) { pre[r] = __shfl_sync(0xffffffffu, incl - rn, r); st[r] = __shfl_sync(0xffffffffu, rs, r); }
    for (int base = 0; base < total; base += 32) {
      const int j = base + lane;
      bool hit = false;
      float4 q = make_float4(0.f, 0.f, 0.f, 0.f), qf = q;
      if (j < total) {
        int src = st[0] + j;
#pragma unroll
        for (int r = 1; r < 9; ++r) if (j >= pre[r]) src = st[r] + (j - pre[r]);
        q = __ldg(sp + 2 * src); qf = __ldg(sp + 2 * src + 1);        // both halves: ~17 candidates x 32 B beat a dependent round trip
        hit = sqdist_rn(cx, cy, cz, q.x, q.y, q.z) <= t_max;
      }
      const unsigned bal = __ballot_sync(0xffffffffu, hit);
      if (bal) {
        const int pos = cnt + __popc(bal & lt);
        if (hit && pos < BG_CAP) { hn[pos] = __float_as_int(q.w); h0[pos] = q; h1[pos] = qf; }
        cnt += __popc(bal);
        if (cnt > BG_CAP) { brute = true; break; }
      }
    }
  }
  __syncwarp();
  int32_t* o = out_idx + (size_t)w * K;
  if (brute) {
    // early-exit in-order scan (exactly the reference loop); dense balls, degenerate grids, K > BG_CAP
    cnt = 0;
    for (int base = 0; base < N && cnt < K; base += 32) {
      const int n = base + lane;
      const bool hit = (n < N) && (sqdist_rn(cx, cy, cz, __ldg(p + n), __ldg(p + N + n), __ldg(p + 2 * N + n)) <= t_max);
      const unsigned bal = __ballot_sync(0xffffffffu, hit);
      if (bal) {
        const int pos = cnt + __popc(bal & lt);
        if (hit && pos < K) o[pos] = n;
        cnt += __popc(bal);
      }
    }
    __syncwarp();
  } else {
    // order by point index: rank = number of hits with a smaller index (indices are distinct).  cnt is ~3 on
    // LiDAR-density clouds: a handful of broadcast shared loads where a bitonic network costs ~300 instructions
    for (int i = lane; i < cnt; i += 32) {
      const int nmine = hn[i];
      int rk = 0;
      for (int j = 0; j < cnt; ++j) rk += hn[j] < nmine ? 1 : 0;
      sl[rk] = i;
    }
    __syncwarp();
  }
  // out[k] = hits[k % u] (first u in index order, then the cyclic pad of ball_query_cuda.cu:40-46); no hit -> point 0.
  // Index, gather and the decentred group are written in the same pass.
  const int u = min(cnt, K);
  const uint32_t um = u > 1 ? (0xffffffffu / (uint32_t)u + 1u) : 0u;      // u >= 2: k % u == k - u * umulhi(k, ceil(2^32 / u))
  const float* pf = feat + (size_t)b * S * N;
  for (int k = lane; k < K; k += 32) {
    const int e = u > 1 ? (k - u * (int)__umulhi((uint32_t)k, um)) : 0;   // k % u
    int n = 0;
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (!brute && u > 0) {
      const float4 r0 = h0[sl[e]], r1 = h1[sl[e]];
      n = __float_as_int(r0.w);
      v[0] = r0.x - cx; v[1] = r0.y - cy; v[2] = r0.z - cz; v[3] = r1.x; v[4] = r1.y; v[5] = r1.z; v[6] = r1.w;   // networks.py:373
    } else {
      n = u > 0 ? o[e] : 0;                                               // brute: the first u entries of o are final
      if (out_group || out_rows) {
        v[0] = __ldg(p + n) - cx; v[1] = __ldg(p + N + n) - cy; v[2] = __ldg(p + 2 * N + n) - cz;
        for (int c = 0; c < S && c < 4; ++c) v[3 + c] = __ldg(pf + (size_t)c * N + n);
      }
    }
    if (!brute || k >= u) o[k] = n;
    if (out_group) {
      float* gp = out_group + ((size_t)b * C * M + m) * K + k;
      for (int c = 0; c < C; ++c) gp[(size_t)c * M * K] = v[c];
    }
    if (out_rows) {
      float* rowp = out_rows + ((size_t)w * K + k) * ld_rows;
      if (ld_rows == 8) {
        reinterpret_cast<float4*>(rowp)[0] = make_float4(v[0], v[1], v[2], v[3]);
        reinterpret_cast<float4*>(rowp)[1] = make_float4(v[4], v[5], v[6], v[7]);
      } else {
        for (int c = 0; c < ld_rows; ++c) rowp[c] = c < C ? v[c] : 0.f;
      }
    }
  }
}

// largest float t with sqrtf(t) <= radius  (host, exact)
static float radius_to_tmax(float radius) {
  if (!(radius >= 0.0f)) return -1.0f;           // negative or NaN radius: nothing is ever inside
  if (isinf(radius)) return INFINITY;
  float t = radius * radius;
  while (sqrtf(t) > radius) t = nextafterf(t, -INFINITY);
  for (;;) {
    float up = nextafterf(t, INFINITY);
    if (isinf(up) || sqrtf(up) > radius) break;
    t = up;
  }
  return t;
}

}  // namespace usip

using namespace usip;


extern "C" int usip_ball_query_dist_f32(const float* dist, float radius, int32_t* out_idx,
                                        int B, int M, int N, int K, void* stream) {
  USIP_REQUIRE(dist && out_idx && B > 0 && M > 0 && N > 0 && K > 0, "ball_query_dist: bad args");
  int rows = B * M;
  ball_query_dist_kernel<<<cdiv(rows, 8), 256, 0, (cudaStream_t)stream>>>(dist, radius, out_idx, rows, N, K);
  return check_launch("ball_query_dist_kernel");
}

static size_t bg_align(size_t x) { return (x + 255) & ~(size_t)255; }
struct BgScratch {
  BgGrid* grids; float* part; int32_t* table; int32_t* done; float4* srec; size_t total; int nch;
  BgScratch(void* base, int B, int N) {
    nch = cdiv(N, BG_CHUNK);
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += bg_align(bytes); return o; };
    const size_t o_g = take(sizeof(BgGrid) * B), o_p = take(sizeof(float) * 8 * (size_t)B * nch);
    const size_t o_t = take(sizeof(int32_t) * (size_t)B * (BG_MAX_CELLS + 1)), o_d = take(sizeof(int32_t) * B);
    const size_t o_s = take(sizeof(float4) * 2 * (size_t)B * N);
    total = off;
    char* p = (char*)base;
    grids = (BgGrid*)(p + o_g); part = (float*)(p + o_p); table = (int32_t*)(p + o_t); done = (int32_t*)(p + o_d); srec = (float4*)(p + o_s);
  }
};

extern "C" int64_t usip_ball_group_scratch_bytes(int B, int S, int N, int M, int K) {
  (void)S; (void)M; (void)K;
  BgScratch sc(nullptr, B, N);
  return (int64_t)sc.total + 256;
}

extern "C" int usip_ball_group_f32(const float* xyz, const float* feat, const float* centers, float radius,
                                   int32_t* out_idx, float* out_group, float* out_rows, int ld_rows, void* scratch,
                                   int64_t scratch_bytes, int B, int S, int N, int M, int K, void* stream) {
  USIP_REQUIRE(xyz && centers && out_idx && (S == 0 || feat) && B > 0 && N > 0 && M > 0 && K > 0,
               "ball_group: bad args");
  USIP_REQUIRE(!out_rows || ld_rows >= 3 + S, "ball_group: ld_rows too small");
  cudaStream_t st = (cudaStream_t)stream;
  const float t_max = radius_to_tmax(radius);
  const int rows = B * M;
  const bool grid_ok = scratch && S <= 4 && (reinterpret_cast<uintptr_t>(scratch) % 256) == 0 &&
                       scratch_bytes >= usip_ball_group_scratch_bytes(B, S, N, M, K) - 256 &&
                       (!out_rows || (reinterpret_cast<uintptr_t>(out_rows) % 16) == 0);
  if (!grid_ok) {
    ball_group_brute_kernel<<<cdiv(rows, 8), 256, 0, st>>>(xyz, feat, centers, t_max, out_idx, out_group, out_rows,
                                                           ld_rows, B, S, N, M, K);
    return check_launch("ball_group_brute_kernel");
  }
  BgScratch sc(scratch, B, N);
  const dim3 pgrid(sc.nch, B);
  bg_bbox_kernel<<<pgrid, 256, 0, st>>>(xyz, sc.part, sc.table, sc.done, N, sc.nch);
  bg_hist_kernel<<<dim3(cdiv(N, 1024), B), 1024, 0, st>>>(xyz, sc.part, radius, sc.grids, sc.table, sc.done, N, sc.nch);
  bg_scatter_kernel<<<dim3(cdiv(N, 256), B), 256, 0, st>>>(xyz, feat, sc.grids, sc.table, sc.srec, S, N);
  bg_query_kernel<<<cdiv(rows, 8), 256, 0, st>>>(xyz, feat, centers, sc.grids, sc.table, sc.srec, t_max, out_idx, out_group,
                                                 out_rows, ld_rows, B, S, N, M, K);
  return check_launch("ball_group_grid");
}
