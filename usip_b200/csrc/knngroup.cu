// knngroup.cu -- k-nearest-POINT grouping of the ablation detector RPN_Detector_KNN (models/networks.py:556-565):
//   node_x_dist = torch.norm(node - x), topk(k=64, largest=False, sorted=False), gather x_aug, subtract the node
// without the (B,M,N) distance matrix: one warp per node selects its K nearest points EXACTLY by a 4-pass radix select on
// the ordered bits of d^2 (d^2 formed with the reference's fp32 op order; sqrt is monotone, so the K smallest d^2 are the K
// smallest norms), then emits them in ascending point index together with the gathered, decentred group.  The reference
// asks for unsorted top-k and only ever max-pools / batch-normalises over the K neighbours, so the order is free; points
// tied with the K-th distance are taken lowest index first (torch.topk leaves that choice unspecified).
#include "common.cuh"

namespace usip {

constexpr int KG_TILE = 2048;                     // points staged per shared-memory tile (24 KB)

__device__ __forceinline__ uint32_t kg_key(float cx, float cy, float cz, float x, float y, float z) {
  const float d2 = sqdist_rn(cx, cy, cz, x, y, z);
  return d2 <= 3.4028234e38f ? __float_as_uint(d2) : 0x7f800000u;     // d2 >= 0: the bit pattern is already ordered; NaN/inf last
}

__global__ void __launch_bounds__(256)
knn_group_kernel(const float* __restrict__ xyz, const float* __restrict__ feat, const float* __restrict__ centers,
                 int32_t* __restrict__ out_idx, float* __restrict__ out_group, float* __restrict__ out_rows, int ld_rows,
                 int S, int N, int M, int K) {
  __shared__ float tx[KG_TILE], ty[KG_TILE], tz[KG_TILE];
  __shared__ int hist[8][256];
  const int b = blockIdx.y, wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m = blockIdx.x * 8 + wib;
  const bool active = m < M;
  const float* p = xyz + (size_t)b * 3 * N;
  const float* cp = centers + (size_t)b * 3 * M;
  const float cx = active ? cp[m] : 0.f, cy = active ? cp[M + m] : 0.f, cz = active ? cp[2 * M + m] : 0.f;
  const unsigned lt = (1u << lane) - 1u;
  uint32_t prefix = 0, pmask = 0;
  int need = K;                                   // the need-th smallest among the keys that match (prefix, pmask)
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    for (int i = lane; i < 256; i += 32) hist[wib][i] = 0;
    for (int t0 = 0; t0 < N; t0 += KG_TILE) {
      const int tn = min(KG_TILE, N - t0);
      __syncthreads();
      for (int i = threadIdx.x; i < tn; i += 256) { tx[i] = p[t0 + i]; ty[i] = p[N + t0 + i]; tz[i] = p[2 * N + t0 + i]; }
      __syncthreads();
      if (active) {
        for (int j0 = 0; j0 < tn; j0 += 32) {
          const int j = j0 + lane;
          int bin = -1;
          if (j < tn) {
            const uint32_t key = kg_key(cx, cy, cz, tx[j], ty[j], tz[j]);
            if ((key & pmask) == prefix) bin = (int)((key >> shift) & 255u);
          }
          const unsigned same = __match_any_sync(0xffffffffu, bin);           // one shared-memory add per distinct bin
          if (bin >= 0 && (same & lt) == 0) hist[wib][bin] += __popc(same);
        }
      }
    }
    __syncwarp();
    // the bin holding the need-th smallest: prefix sums of 8 bins per lane
    int loc[8], tot = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { loc[i] = hist[wib][lane * 8 + i]; tot += loc[i]; }
    int incl = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    int run = incl - tot, sel = -1, before = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (sel < 0 && run + loc[i] >= need && run < need) { sel = lane * 8 + i; before = run; }
      run += loc[i];
    }
    const unsigned who = __ballot_sync(0xffffffffu, sel >= 0);
    const int src = who ? __ffs(who) - 1 : 0;
    sel = __shfl_sync(0xffffffffu, sel, src); before = __shfl_sync(0xffffffffu, before, src);
    if (!who) { sel = 255; before = 0; }                                       // K > N (rejected by the host)
    need -= before;
    prefix |= (uint32_t)sel << shift; pmask |= 255u << shift;
    __syncwarp();
  }
  // prefix is now the key of the K-th nearest point; `need` of the points carrying exactly that key are taken
  const uint32_t T = prefix;
  int32_t* o = out_idx + ((size_t)b * M + m) * K;
  int cnt = 0, eq_seen = 0;
  for (int t0 = 0; t0 < N; t0 += KG_TILE) {
    const int tn = min(KG_TILE, N - t0);
    __syncthreads();
    for (int i = threadIdx.x; i < tn; i += 256) { tx[i] = p[t0 + i]; ty[i] = p[N + t0 + i]; tz[i] = p[2 * N + t0 + i]; }
    __syncthreads();
    if (active) {
      for (int j0 = 0; j0 < tn && cnt < K; j0 += 32) {
        const int j = j0 + lane;
        bool less = false, eq = false;
        if (j < tn) { const uint32_t key = kg_key(cx, cy, cz, tx[j], ty[j], tz[j]); less = key < T; eq = key == T; }
        const unsigned beq = __ballot_sync(0xffffffffu, eq);
        const bool take = less || (eq && eq_seen + __popc(beq & lt) < need);
        const unsigned bt = __ballot_sync(0xffffffffu, take);
        if (take) { const int pos = cnt + __popc(bt & lt); if (pos < K) o[pos] = t0 + j; }
        cnt += __popc(bt); eq_seen += __popc(beq);
      }
    }
  }
  if (!active) return;
  __syncwarp();
  const int C = 3 + S;
  const float ctr[3] = {cx, cy, cz};
  const size_t w = (size_t)b * M + m;
  for (int k = lane; k < K; k += 32) {
    const int n = o[k];
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int c = 0; c < 3; ++c) v[c] = p[(size_t)c * N + n] - ctr[c];                       // networks.py:565
    for (int c = 0; c < S && c < 5; ++c) v[3 + c] = feat[((size_t)b * S + c) * N + n];
    if (out_group) for (int c = 0; c < C; ++c) out_group[(((size_t)b * C + c) * M + m) * K + k] = v[c];
    if (out_rows) {
      float* rowp = out_rows + (w * K + k) * ld_rows;
      for (int c = 0; c < ld_rows; ++c) rowp[c] = c < C ? v[c] : 0.f;
    }
  }
}

}  // namespace usip

extern "C" int usip_knn_group_f32(const float* xyz, const float* feat, const float* centers, int32_t* out_idx,
                                  float* out_group, float* out_rows, int ld_rows, int B, int S, int N, int M, int K,
                                  void* stream) {
  using namespace usip;
  USIP_REQUIRE(xyz && centers && out_idx && (S == 0 || feat) && B > 0 && N > 0 && M > 0, "knn_group: bad args");
  USIP_REQUIRE(K > 0 && K <= N, "knn_group: need 0 < K <= N");
  USIP_REQUIRE(S <= 5 && (!out_rows || ld_rows >= 3 + S), "knn_group: S <= 5 and ld_rows >= 3+S");
  knn_group_kernel<<<dim3(cdiv(M, 8), B), 256, 0, (cudaStream_t)stream>>>(xyz, feat, centers, out_idx, out_group, out_rows, ld_rows,
                                                                         S, N, M, K);
  return check_launch("knn_group_kernel");
}
