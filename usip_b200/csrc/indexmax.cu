// indexmax.cu -- stand-alone index_max operator (models/index_max_ext/index_max_cuda.cu:9-61) at HBM speed.
//
// out[b,c,k] = smallest n attaining max{ data[b,c,n] : index[b,n] = k, data > -1000 }, else 0.
//
// The reference walks the N points of every (b,c) row serially; round 1 of this repo used one shared-memory 64-bit
// atomicMax per element (1.2 TB/s).  Here the atomics are paid ONCE per cloud instead of once per element: a persistent
// CTA (one per SM) buckets the N point indices of its cloud by cluster in shared memory (counting sort: 2N shared-memory
// atomics, shared by all C channel rows of that cloud), then streams its [N]-float rows through a double-buffered
// cp.async.bulk (TMA) ring and lets thread k take the maximum of cluster k by gathering the row through the bucket order
// -- no atomics, no global-memory round trip per element.  Ties: (value, n) lexicographic, i.e. the reference's first
// maximum, independent of the (unstable) bucket order.  134 MB of data are read exactly once.
#include "tc_common.cuh"

namespace usip {

constexpr int IMS_THREADS = 1024;          // two threads per cluster at K = 512: half a segment each, merged by one shuffle
constexpr int IMS_MAXN = 16384;            // one row buffer = 64 KB; bucket order as uint16
constexpr int IMS_MAXK = 4096;
constexpr int IMS_SPLIT = 16;              // concurrent bulk copies per row
constexpr int IMS_REG = 24;                // point indices a thread keeps in registers (clusters up to 48 points stay out of smem)

struct ImsLayout {
  int npad; size_t row_bytes, perm_off, seg_off, cur_off, wtot_off, bar_off, total;
  __host__ __device__ ImsLayout(int N, int K) {
    npad = (N + 3) & ~3;
    row_bytes = (size_t)npad * 4;
    perm_off = 2 * row_bytes;
    seg_off = perm_off + (((size_t)N * 2 + 15) & ~(size_t)15);
    cur_off = seg_off + (((size_t)(K + 1) * 4 + 15) & ~(size_t)15);
    wtot_off = cur_off + (((size_t)K * 4 + 15) & ~(size_t)15);
    bar_off = wtot_off + 32 * 4;
    total = bar_off + 16;
  }
};

__global__ void __launch_bounds__(IMS_THREADS, 1)
index_max_bucket_kernel(const float* __restrict__ data, const int32_t* __restrict__ index, int32_t* __restrict__ out,
                        int B, int C, int N, int K, int rows_per_cta) {
  extern __shared__ __align__(128) uint8_t ims_smem[];
  const ImsLayout L(N, K);
  float* rowbuf = reinterpret_cast<float*>(ims_smem);
  uint16_t* perm = reinterpret_cast<uint16_t*>(ims_smem + L.perm_off);
  int* seg = reinterpret_cast<int*>(ims_smem + L.seg_off);
  int* cur = reinterpret_cast<int*>(ims_smem + L.cur_off);
  int* wtot = reinterpret_cast<int*>(ims_smem + L.wtot_off);
  const uint32_t bar0 = smem_u32(ims_smem + L.bar_off);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int row0 = blockIdx.x * rows_per_cta, row1 = min(B * C, row0 + rows_per_cta);
  if (row0 >= row1) return;
  if (tid == 0) { mbar_init(bar0, 1); mbar_init(bar0 + 8, 1); fence_barrier_init(); }
  __syncthreads();
  const uint32_t nbytes = (uint32_t)N * 4u;
  // a row is fetched as IMS_SPLIT concurrent bulk copies (one per lane of warp 0) on one mbarrier: a single 64 KB
  // cp.async.bulk per SM leaves the TMA unit latency-bound (measured: ~4 us per row = 16 GB/s per SM)
  const int nsplit = (N % (4 * IMS_SPLIT) == 0) ? IMS_SPLIT : 1;
  const uint32_t cbytes = nbytes / (uint32_t)nsplit;
  auto fetch_row = [&](int r, int buf) {
    const uint32_t bar = bar0 + 8 * buf;
    if (tid == 0) mbar_arrive_expect_tx(bar, nbytes);
    __syncwarp();
    if (tid < nsplit)
      bulk_g2s(smem_u32(rowbuf + (size_t)buf * L.npad) + tid * cbytes,
               reinterpret_cast<const char*>(data + (size_t)r * N) + (size_t)tid * cbytes, cbytes, bar);
  };
  if (warp == 0) fetch_row(row0, 0);
  int cur_b = -1, reg_b = -1, my_j0 = 0, my_j1 = 0;
  const int sub = tid & 1;
  uint32_t preg[IMS_REG / 2];
  const int ipt = (K + IMS_THREADS - 1) / IMS_THREADS;            // clusters per thread in the scan
  for (int r = row0, it = 0; r < row1; ++r, ++it) {
    const int s = it & 1;
    if (warp == 0 && r + 1 < row1) fetch_row(r + 1, s ^ 1);        // the other buffer was released by the barrier that ended it-1
    const int b = r / C;
    if (b != cur_b) {
      // ---- bucket order of cloud b: counting sort of n by index[b,n] in shared memory
      const int32_t* idx = index + (size_t)b * N;
      for (int k = tid; k < K; k += IMS_THREADS) cur[k] = 0;
      __syncthreads();
      for (int n = tid; n < N; n += IMS_THREADS) { const int k = idx[n]; if ((unsigned)k < (unsigned)K) atomicAdd(&cur[k], 1); }
      __syncthreads();
      int loc = 0;
      for (int i = 0; i < ipt; ++i) { const int k = tid * ipt + i; if (k < K) loc += cur[k]; }
      int incl = loc;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
      if (lane == 31) wtot[warp] = incl;
      __syncthreads();
      if (warp == 0) {
        const int x = lane < IMS_THREADS / 32 ? wtot[lane] : 0;
        int ix = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, ix, o); if (lane >= o) ix += v; }
        if (lane < IMS_THREADS / 32) wtot[lane] = ix - x;
        if (lane == 31) seg[K] = ix;
      }
      __syncthreads();
      int run = wtot[warp] + incl - loc;
      for (int i = 0; i < ipt; ++i) {
        const int k = tid * ipt + i;
        if (k < K) { const int c = cur[k]; seg[k] = run; cur[k] = run; run += c; }
      }
      __syncthreads();
      for (int n = tid; n < N; n += IMS_THREADS) {
        const int k = idx[n];
        if ((unsigned)k < (unsigned)K) perm[atomicAdd(&cur[k], 1)] = (uint16_t)n;
      }
      __syncthreads();
      cur_b = b;
    }
    if (b != reg_b) {
      // this thread's share of its cluster (every other element) is the same for all C rows of the cloud: keep up to
      // IMS_REG point indices in registers (two 16-bit indices per register) -- per row only the row gather remains
      reg_b = b;
#pragma unroll
      for (int t = 0; t < IMS_REG / 2; ++t) preg[t] = 0xffffffffu;
      my_j0 = 0; my_j1 = 0;
      const int k = tid >> 1;
      if (k < K) {
        my_j0 = seg[k] + sub; my_j1 = seg[k + 1];
#pragma unroll
        for (int t = 0; t < IMS_REG; ++t) {
          const int j = my_j0 + 2 * t;
          const uint32_t n = j < my_j1 ? perm[j] : 0xffffu;
          if (t & 1) preg[t >> 1] = (preg[t >> 1] & 0xffffu) | (n << 16); else preg[t >> 1] = (preg[t >> 1] & 0xffff0000u) | n;
        }
      }
    }
    mbar_wait(bar0 + 8 * s, (it >> 1) & 1);
    const float* row = rowbuf + (size_t)s * L.npad;
    // lanes 2i and 2i+1 share cluster k: each walks every other element, then the pair is merged.
    // Order: (value, n) lexicographic = the reference's first maximum, independent of the (unstable) bucket order.
    for (int k0 = 0; k0 < K; k0 += IMS_THREADS / 2) {
      const int k = k0 + (tid >> 1);
      float best = -1000.0f;                                      // index_max_cuda.cu:38 / :71
      int bn = 0x7fffffff;                                        // "nothing above the floor yet"
      if (k0 == 0) {
#pragma unroll
        for (int t = 0; t < IMS_REG; ++t) {
          const int n = (int)((preg[t >> 1] >> ((t & 1) * 16)) & 0xffffu);
          if (n != 0xffff) {                                       // N <= 16384: 0xffff is never a point
            const float v = row[n];
            if (v > best || (v == best && n < bn && bn != 0x7fffffff)) { best = v; bn = n; }
          }
        }
        for (int j = my_j0 + 2 * IMS_REG; j < my_j1; j += 2) {      // the tail of an unusually large cluster
          const int n = perm[j];
          const float v = row[n];
          if (v > best || (v == best && n < bn && bn != 0x7fffffff)) { best = v; bn = n; }
        }
      } else if (k < K) {                                          // K > IMS_THREADS / 2: further clusters through shared memory
        const int j1 = seg[k + 1];
        for (int j = seg[k] + sub; j < j1; j += 2) {
          const int n = perm[j];
          const float v = row[n];
          if (v > best || (v == best && n < bn && bn != 0x7fffffff)) { best = v; bn = n; }
        }
      }
      const float ob = __shfl_xor_sync(0xffffffffu, best, 1);
      const int on = __shfl_xor_sync(0xffffffffu, bn, 1);
      if (on != 0x7fffffff && (bn == 0x7fffffff || ob > best || (ob == best && on < bn))) { best = ob; bn = on; }
      if (k < K && sub == 0) out[(size_t)r * K + k] = bn == 0x7fffffff ? 0 : bn;
    }
    __syncthreads();                                              // rowbuf[s] may be refilled from the next iteration on
  }
}

// true when the bucket kernel applies (otherwise the shared-memory atomic kernel of group.cu runs)
bool index_max_bucket_ok(const float* data, int N, int K) {
  return N <= IMS_MAXN && (N % 4) == 0 && K <= IMS_MAXK && (reinterpret_cast<uintptr_t>(data) % 16) == 0;
}

int launch_index_max_bucket(const float* data, const int32_t* index, int32_t* out, int B, int C, int N, int K, cudaStream_t st) {
  const ImsLayout L(N, K);
  static size_t attr_bytes = 0;
  if (L.total > attr_bytes) {
    cudaError_t e = cudaFuncSetAttribute(index_max_bucket_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total);
    if (e != cudaSuccess) { set_last_error("index_max_bucket smem attr"); return (int)e; }
    attr_bytes = L.total;
  }
  int sm_count = 148, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
  const int rows = B * C;
  const int rpc = cdiv(rows, min(rows, sm_count));
  index_max_bucket_kernel<<<cdiv(rows, rpc), IMS_THREADS, L.total, st>>>(data, index, out, B, C, N, K, rpc);
  return check_launch("index_max_bucket_kernel");
}

}  // namespace usip
