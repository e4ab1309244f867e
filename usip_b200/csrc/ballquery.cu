// ballquery.cu -- ball query operators.
//   usip_ball_query_dist_f32 : drop-in for ball_query.forward_cuda_shared_mem on a pre-computed
//                              distance matrix (models/ball_query_ext/ball_query_cuda.cu:10-49)
//   ball_group_brute         : fused distance + ball query + gather + decentre from xyz by the reference's own in-order
//                              scan (the bucket-grid operator usip_ball_group_f32 lives in ballgroup.cu)
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

// largest float t with sqrtf(t) <= radius  (host, exact)
float radius_to_tmax_host(float radius) {
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

namespace usip {
// the reference loop itself (one warp per keypoint, in-order scan with early exit); also the path of ballgroup.cu when no
// scratch is given
int ball_group_brute(const float* xyz, const float* feat, const float* centers, float t_max, int32_t* out_idx, float* out_group,
                     float* out_rows, int ld_rows, int B, int S, int N, int M, int K, cudaStream_t st) {
  ball_group_brute_kernel<<<cdiv(B * M, 8), 256, 0, st>>>(xyz, feat, centers, t_max, out_idx, out_group, out_rows, ld_rows, B, S,
                                                          N, M, K);
  return check_launch("ball_group_brute_kernel");
}
}  // namespace usip
