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
// Cell-binned variant: points are binned once per cloud into a uniform grid with cell size h >= 1.001*r, a
// keypoint only tests the <= 27 neighbouring cells (9 x-contiguous ranges), collects the hits (unordered),
// sorts their indices ascending and keeps the first K -- bit-identical to the in-order scan, without the O(N)
// pass per keypoint.  Dense balls (more than BG_CAP hits) fall back to the early-exit in-order scan, which is
// cheap exactly there.
// ------------------------------------------------------------------------------------------------
constexpr int BG_MAX_CELLS = 65536;
constexpr int BG_CAP = 256;

struct BgGrid { float ox, oy, oz, inv_h; int nx, ny, nz, ok; };

__device__ __forceinline__ int bg_cell1(float v, float o, float inv_h, int n) {
  int c = (int)floorf((v - o) * inv_h);
  return min(max(c, 0), n - 1);
}

// Fused per-cloud preparation (one CTA per cloud): bounding box -> grid -> shared-memory cell histogram -> exclusive
// scan -> cell-sorted point array + 32-byte AoS records.  Used for clouds too large for the single-kernel cluster path (ballgroup_cluster.cu).
constexpr int BG_PREP_CELLS = 49152;      // int32 counters in dynamic shared memory (192 KB)

__global__ void __launch_bounds__(1024)
bg_prep_kernel(const float* __restrict__ xyz, const float* __restrict__ feat, float radius, BgGrid* __restrict__ grids,
               int32_t* __restrict__ cell_start, float4* __restrict__ sorted, float* __restrict__ rec, int S, int N) {
  extern __shared__ int hist[];                          // [cells]
  __shared__ float smin[3][32], smax[3][32];
  __shared__ int wsum[32];
  __shared__ BgGrid sg;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const float* p = xyz + (size_t)b * 3 * N;
  // ---- phase 1: bounding box (+ AoS records)
  float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  bool bad = false;
  for (int n = tid; n < N; n += 1024) {
    const float x = p[n], y = p[N + n], z = p[2 * N + n];
    bad |= !(fabsf(x) <= 1e30f) || !(fabsf(y) <= 1e30f) || !(fabsf(z) <= 1e30f);
    mn[0] = fminf(mn[0], x); mx[0] = fmaxf(mx[0], x); mn[1] = fminf(mn[1], y); mx[1] = fmaxf(mx[1], y);
    mn[2] = fminf(mn[2], z); mx[2] = fmaxf(mx[2], z);
    float r[8] = {x, y, z, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int c = 0; c < S && c < 5; ++c) r[3 + c] = feat[((size_t)b * S + c) * N + n];
    float4* rp = reinterpret_cast<float4*>(rec + ((size_t)b * N + n) * 8);
    rp[0] = make_float4(r[0], r[1], r[2], r[3]); rp[1] = make_float4(r[4], r[5], r[6], r[7]);
  }
  const unsigned anybad = __ballot_sync(0xffffffffu, bad);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn[c] = fminf(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], o));
      mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o));
    }
    if (lane == 0) { smin[c][w] = anybad ? NAN : mn[c]; smax[c][w] = mx[c]; }
  }
  __syncthreads();
  if (tid == 0) {
    BgGrid g; g.ok = 1;
    float lo[3], hi[3];
    for (int c = 0; c < 3; ++c) {
      lo[c] = INFINITY; hi[c] = -INFINITY;
      for (int i = 0; i < 32; ++i) { if (smin[c][i] != smin[c][i]) g.ok = 0; lo[c] = fminf(lo[c], smin[c][i]); hi[c] = fmaxf(hi[c], smax[c][i]); }
    }
    if (!(radius >= 0.f) || !(radius <= 1e30f)) g.ok = 0;
    const float ext = fmaxf(fmaxf(hi[0] - lo[0], hi[1] - lo[1]), hi[2] - lo[2]);
    float h = fmaxf(fmaxf(radius * 1.001f, ext * (1.0f / 126.0f)), 1e-6f);
    int nx = 1, ny = 1, nz = 1;
    if (g.ok) {
      for (int it = 0; it < 64; ++it) {
        nx = (int)floorf((hi[0] - lo[0]) / h) + 1; ny = (int)floorf((hi[1] - lo[1]) / h) + 1; nz = (int)floorf((hi[2] - lo[2]) / h) + 1;
        if ((long long)nx * ny * nz <= BG_PREP_CELLS) break;
        h *= 1.26f;
      }
      if ((long long)nx * ny * nz > BG_PREP_CELLS) g.ok = 0;
    }
    g.ox = lo[0]; g.oy = lo[1]; g.oz = lo[2]; g.inv_h = 1.0f / h; g.nx = nx; g.ny = ny; g.nz = nz;
    sg = g; grids[b] = g;
  }
  __syncthreads();
  const BgGrid g = sg;
  if (!g.ok) return;
  const int cells = g.nx * g.ny * g.nz;
  // ---- phase 2: histogram in shared memory
  for (int i = tid; i < cells; i += 1024) hist[i] = 0;
  __syncthreads();
  for (int n = tid; n < N; n += 1024) {
    const int cell = (bg_cell1(p[2 * N + n], g.oz, g.inv_h, g.nz) * g.ny + bg_cell1(p[N + n], g.oy, g.inv_h, g.ny)) * g.nx +
                     bg_cell1(p[n], g.ox, g.inv_h, g.nx);
    atomicAdd(&hist[cell], 1);
  }
  __syncthreads();
  // ---- phase 3: exclusive scan (thread slice -> warp scan -> block)
  const int per = (cells + 1023) / 1024, lo = min(tid * per, cells), hi = min(lo + per, cells);
  int s = 0;
  for (int i = lo; i < hi; ++i) s += hist[i];
  int incl = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
  if (lane == 31) wsum[w] = incl;
  __syncthreads();
  if (w == 0) {
    int v = wsum[lane], iv = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int u = __shfl_up_sync(0xffffffffu, iv, o); if (lane >= o) iv += u; }
    wsum[lane] = iv - v;                                 // exclusive warp offsets
  }
  __syncthreads();
  int run = wsum[w] + incl - s;
  int32_t* cs = cell_start + (size_t)b * (BG_MAX_CELLS + 1);
  for (int i = lo; i < hi; ++i) { const int v = hist[i]; hist[i] = run; cs[i] = run; run += v; }
  if (tid == 0) cs[cells] = N;
  __syncthreads();
  // ---- phase 4: fill the cell-sorted array (hist is now the running cursor)
  float4* sp = sorted + (size_t)b * N;
  for (int n = tid; n < N; n += 1024) {
    const float x = p[n], y = p[N + n], z = p[2 * N + n];
    const int cell = (bg_cell1(z, g.oz, g.inv_h, g.nz) * g.ny + bg_cell1(y, g.oy, g.inv_h, g.ny)) * g.nx + bg_cell1(x, g.ox, g.inv_h, g.nx);
    const int pos = atomicAdd(&hist[cell], 1);
    sp[pos] = make_float4(x, y, z, __int_as_float(n));
  }
}

__global__ void __launch_bounds__(256)
bg_query_kernel(const float* __restrict__ xyz, const float* __restrict__ centers, const BgGrid* __restrict__ grids,
                const int32_t* __restrict__ cell_start, const float4* __restrict__ sorted, const float* __restrict__ rec,
                float t_max, int32_t* __restrict__ out_idx, float* __restrict__ out_group, float* __restrict__ out_rows,
                int ld_rows, int B, int S, int N, int M, int K) {
  __shared__ int hits[8][BG_CAP];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int w = blockIdx.x * 8 + wib;
  if (w >= B * M) return;
  const int b = w / M, m = w - b * M;
  const BgGrid g = grids[b];
  const float cx = centers[(size_t)b * 3 * M + m], cy = centers[(size_t)b * 3 * M + M + m], cz = centers[(size_t)b * 3 * M + 2 * M + m];
  int32_t* o = out_idx + (size_t)w * K;
  int* hl = hits[wib];
  const unsigned lt = (1u << lane) - 1u;
  int cnt = 0;
  bool brute = !g.ok;
  if (!brute) {
    // un-clamped cell coordinates of the centre; neighbour ranges are clipped to the grid
    const int kx = (int)floorf((cx - g.ox) * g.inv_h), ky = (int)floorf((cy - g.oy) * g.inv_h), kz = (int)floorf((cz - g.oz) * g.inv_h);
    const bool cfin = fabsf(cx) <= 1e30f && fabsf(cy) <= 1e30f && fabsf(cz) <= 1e30f;
    const int x0 = max(kx - 1, 0), x1 = min(kx + 1, g.nx - 1);
    const int32_t* cs = cell_start + (size_t)b * (BG_MAX_CELLS + 1);
    const float4* sp = sorted + (size_t)b * N;
    // lanes 0..8 fetch the bounds of the 9 x-contiguous ranges in one round trip
    int rs = 0, rn = 0;
    if (lane < 9 && cfin && x0 <= x1) {
      const int y = ky + (lane % 3) - 1, z = kz + (lane / 3) - 1;
      if (y >= 0 && y < g.ny && z >= 0 && z < g.nz) {
        const int base = (z * g.ny + y) * g.nx;
        rs = cs[base + x0]; rn = cs[base + x1 + 1] - rs;
      }
    }
    int incl = rn;                                       // inclusive prefix of the range lengths over lanes 0..8
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    const int total = __shfl_sync(0xffffffffu, incl, 8);
    int pre[9], st[9];
#pragma unroll
    for (int r = 0; r < 9; ++r) { pre[r] = __shfl_sync(0xffffffffu, incl - rn, r); st[r] = __shfl_sync(0xffffffffu, rs, r); }
    for (int base = 0; base < total; base += 32) {
      const int j = base + lane;
      bool hit = false; int n = 0;
      if (j < total) {
        int src = st[0] + j;
#pragma unroll
        for (int r = 1; r < 9; ++r) if (j >= pre[r]) src = st[r] + (j - pre[r]);
        const float4 q = sp[src];
        n = __float_as_int(q.w);
        hit = sqdist_rn(cx, cy, cz, q.x, q.y, q.z) <= t_max;
      }
      const unsigned bal = __ballot_sync(0xffffffffu, hit);
      if (bal) {
        const int pos = cnt + __popc(bal & lt);
        if (hit && pos < BG_CAP) hl[pos] = n;
        cnt += __popc(bal);
        if (cnt > BG_CAP) { brute = true; break; }
      }
    }
  }
  if (brute) {
    // early-exit in-order scan (exactly the reference loop); used for dense balls / degenerate grids
    const float* px = xyz + (size_t)b * 3 * N;
    cnt = 0;
    for (int base = 0; base < N && cnt < K; base += 32) {
      const int n = base + lane;
      const bool hit = (n < N) && (sqdist_rn(cx, cy, cz, __ldg(px + n), __ldg(px + N + n), __ldg(px + 2 * N + n)) <= t_max);
      const unsigned bal = __ballot_sync(0xffffffffu, hit);
      if (bal) {
        const int pos = cnt + __popc(bal & lt);
        if (hit && pos < K) o[pos] = n;
        cnt += __popc(bal);
      }
    }
  } else {
    // bitonic sort of the hit list (ascending point index), padded with INT_MAX to a power of two
    __syncwarp();
    int sz = 32;
    while (sz < cnt) sz <<= 1;
    for (int i = cnt + lane; i < sz; i += 32) hl[i] = 0x7fffffff;
    __syncwarp();
    for (int k2 = 2; k2 <= sz; k2 <<= 1)
      for (int j2 = k2 >> 1; j2 > 0; j2 >>= 1) {
        for (int i = lane; i < sz; i += 32) {
          const int ixj = i ^ j2;
          if (ixj > i) {
            const int a = hl[i], c2 = hl[ixj];
            const bool up = (i & k2) == 0;
            if ((a > c2) == up) { hl[i] = c2; hl[ixj] = a; }
          }
        }
        __syncwarp();
      }
    for (int i = lane; i < min(cnt, K); i += 32) o[i] = hl[i];
  }
  ball_pad(o, cnt, K, lane);
  if (!out_group && !out_rows) return;
  const int C = 3 + S;
  for (int k = lane; k < K; k += 32) {
    const int n = o[k];
    const float4* rp = reinterpret_cast<const float4*>(rec + ((size_t)b * N + n) * 8);
    const float4 r0 = rp[0], r1 = rp[1];
    float v[8] = {r0.x - cx, r0.y - cy, r0.z - cz, r0.w, r1.x, r1.y, r1.z, r1.w};          // networks.py:373
    if (out_group)
      for (int c = 0; c < C; ++c) out_group[(((size_t)b * C + c) * M + m) * K + k] = v[c];
    if (out_rows) {
      float* rowp = out_rows + ((size_t)w * K + k) * ld_rows;
      if (ld_rows == 8) {
        for (int c = C; c < 8; ++c) v[c] = 0.f;
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

bool bg_cluster_eligible(int S, int N, int K);                                         // ballgroup_cluster.cu
int launch_bg_cluster(const float* xyz, const float* feat, const float* centers, float radius, float t_max,
                      int32_t* out_idx, float* out_group, float* out_rows, int ld_rows, int B, int S, int N, int M, int K,
                      cudaStream_t st);

int bg_cluster_phase_clocks(unsigned long long* host8);

}  // namespace usip

using namespace usip;

extern "C" int usip_ball_group_phase_clocks(unsigned long long* out8) {
  USIP_REQUIRE(out8, "ball_group_phase_clocks: null");
  return bg_cluster_phase_clocks(out8);
}

extern "C" int usip_ball_query_dist_f32(const float* dist, float radius, int32_t* out_idx,
                                        int B, int M, int N, int K, void* stream) {
  USIP_REQUIRE(dist && out_idx && B > 0 && M > 0 && N > 0 && K > 0, "ball_query_dist: bad args");
  int rows = B * M;
  ball_query_dist_kernel<<<cdiv(rows, 8), 256, 0, (cudaStream_t)stream>>>(dist, radius, out_idx, rows, N, K);
  return check_launch("ball_query_dist_kernel");
}

static size_t bg_align(size_t x) { return (x + 255) & ~(size_t)255; }
struct BgScratch {
  BgGrid* grids; int32_t* cell_cnt; float4* sorted; float* rec; size_t total;
  BgScratch(void* base, int B, int N) {
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += bg_align(bytes); return o; };
    size_t o_g = take(sizeof(BgGrid) * B), o_c = take(sizeof(int32_t) * (size_t)B * (BG_MAX_CELLS + 1));
    size_t o_s = take(sizeof(float4) * (size_t)B * N), o_r = take(sizeof(float) * 8 * (size_t)B * N);
    total = off;
    char* p = (char*)base;
    grids = (BgGrid*)(p + o_g); cell_cnt = (int32_t*)(p + o_c); sorted = (float4*)(p + o_s); rec = (float*)(p + o_r);
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
  // single-kernel path: one 8-CTA cluster per cloud, everything in distributed shared memory (no scratch needed).
  // USIP_BG_NO_CLUSTER=1 forces the older two-kernel grid path (A/B measurements only).
  static const bool no_cluster = getenv("USIP_BG_NO_CLUSTER") != nullptr;
  if (!no_cluster && bg_cluster_eligible(S, N, K) && (!out_rows || (reinterpret_cast<uintptr_t>(out_rows) % 16) == 0))
    return launch_bg_cluster(xyz, feat, centers, radius, t_max, out_idx, out_group, out_rows, ld_rows, B, S, N, M, K, st);
  const bool grid_ok = scratch && S <= 5 && (reinterpret_cast<uintptr_t>(scratch) % 256) == 0 &&
                       scratch_bytes >= usip_ball_group_scratch_bytes(B, S, N, M, K) - 256 &&
                       (!out_rows || (reinterpret_cast<uintptr_t>(out_rows) % 16) == 0);
  if (!grid_ok) {
    ball_group_brute_kernel<<<cdiv(rows, 8), 256, 0, st>>>(xyz, feat, centers, t_max, out_idx, out_group, out_rows,
                                                           ld_rows, B, S, N, M, K);
    return check_launch("ball_group_brute_kernel");
  }
  BgScratch sc(scratch, B, N);
  static bool prep_attr = false;
  if (!prep_attr) {
    cudaError_t e = cudaFuncSetAttribute(bg_prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BG_PREP_CELLS * 4);
    if (e != cudaSuccess) { set_last_error("bg_prep smem attr"); return (int)e; }
    prep_attr = true;
  }
  bg_prep_kernel<<<B, 1024, BG_PREP_CELLS * 4, st>>>(xyz, feat, radius, sc.grids, sc.cell_cnt, sc.sorted, sc.rec, S, N);
  bg_query_kernel<<<cdiv(rows, 8), 256, 0, st>>>(xyz, centers, sc.grids, sc.cell_cnt, sc.sorted, sc.rec, t_max, out_idx,
                                                 out_group, out_rows, ld_rows, B, S, N, M, K);
  return check_launch("ball_group_grid");
}
