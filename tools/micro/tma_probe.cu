// Micro-benchmark (developer tool): gather / scatter of statevector tiles with TMA tensor copies.
//
// A tile = 2^m complex128 amplitudes whose index has the low l bits (one contiguous 16*2^l-byte row in HBM) and
// m - l further "resident" bits chosen by a mask; the other n - m bits number the tiles.  One
// cp.async.bulk.tensor.2d per row (tensor = [N/8 rows of 128 B][16 doubles], box = {16, 2^l / 8}, SWIZZLE_128B)
// lands the row in shared memory in the executor's XOR-swizzled slot order (slot = i ^ ((i >> 3) & 7)).
//   mode 0: verify the shared-memory layout against qsb_slot
//   mode 1: in-place streaming copy (load tile -> store tile) through a 3-deep ring, one issuing warp per CTA;
//           reports GB/s for read + write
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_probe.bin tma_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>

typedef double2 c128;
extern __shared__ __align__(1024) unsigned char smem[];

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %2, %2, %2}], [%4];"
               ::"r"(dst), "l"(map), "r"(0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_store_5d(const CUtensorMap* map, int c1, uint32_t src) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%1, %2, %1, %1, %1}], [%3];"
               ::"l"(map), "r"(0), "r"(c1), "r"(src) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, int c0, int c1, uint32_t src) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
               ::"l"(map), "r"(c0), "r"(c1), "r"(src) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ uint32_t deposit(uint32_t v, uint32_t mask) {
  uint32_t r = 0;
  for (int b = 0; mask; ++b, mask &= mask - 1) {
    const int q = __ffs(mask) - 1;
    r |= ((v >> b) & 1u) << q;
  }
  return r;
}

struct Args {
  int n, m, l;
  int e;                           // resident bits above the row that ride in the box as extra dimensions (0..3)
  uint32_t res_mask, non_mask;     // resident bits above l that number the ops / non-resident bits (both within [l, n))
  int ebit[3];                     // positions of the extra box dimensions
  int mode;
  unsigned long long* mismatches;
  const c128* state;
};

constexpr int NBUF = 3;

__global__ void __launch_bounds__(128, 1) tma_probe_kernel(const __grid_constant__ CUtensorMap map, Args a) {
  const int tile_amps = 1 << a.m, tile_bytes = tile_amps * 16;
  // one TMA op moves 2^e rows; "rows" below counts ops, "row" = the 2^(l+e) amplitudes of one op (contiguous in the tile)
  const int rows = 1 << (a.m - a.l - a.e), row_amps = 1 << (a.l + a.e), row_bytes = row_amps * 16;
  unsigned char* bufs = smem;                                          // NBUF tiles, 1024-byte aligned
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + NBUF * tile_bytes);
  uint32_t* row_off = reinterpret_cast<uint32_t*>(full + NBUF);        // amplitude offset of row r inside the state
  for (int r = threadIdx.x; r < rows; r += blockDim.x) row_off[r] = deposit((uint32_t)r, a.res_mask);
  if (threadIdx.x == 0) {
    for (int b = 0; b < NBUF; ++b) mbar_init(smem_u32(&full[b]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int ntiles = 1 << (a.n - a.m);
  if (a.mode == 0) {
    // one tile per CTA, then check the layout
    const int t = blockIdx.x;
    if (t >= ntiles) return;
    const uint32_t tbase = deposit((uint32_t)t, a.non_mask);
    if (threadIdx.x == 0) mbar_expect(smem_u32(&full[0]), tile_bytes);
    __syncthreads();
    for (int r = threadIdx.x; r < rows; r += blockDim.x)
      tma_load_5d(smem_u32(bufs + (size_t)r * row_bytes), &map, (int)((tbase | row_off[r]) >> 3), smem_u32(&full[0]));
    mbar_wait(smem_u32(&full[0]), 0);
    const c128* tile = reinterpret_cast<const c128*>(bufs);
    unsigned long long bad = 0;
    for (int i = threadIdx.x; i < tile_amps; i += blockDim.x) {
      uint32_t g = tbase | row_off[i >> (a.l + a.e)] | (uint32_t)(i & ((1 << a.l) - 1));
      for (int j = 0; j < a.e; ++j) g |= ((uint32_t)(i >> (a.l + j)) & 1u) << a.ebit[j];
      const c128 want = a.state[g], got = tile[i ^ ((i >> 3) & 7)];
      if (want.x != got.x || want.y != got.y) ++bad;
    }
    if (bad) atomicAdd(a.mismatches, bad);
    return;
  }
  if (threadIdx.x >= 32) return;
  const int lane = threadIdx.x;
  // tiles of this CTA: t = blockIdx.x, + gridDim.x, ...
  int k = 0;
  uint32_t prev_base = 0;
  for (int t = blockIdx.x;; t += gridDim.x, ++k) {
    const bool have = t < ntiles;
    const int b = k % NBUF;
    if (have) {
      // buffer b was last stored from in iteration k - NBUF + ... : at most NBUF - 2 newer store groups may still read
      bulk_wait_read<NBUF - 2>();
      __syncwarp();
      const uint32_t tbase = deposit((uint32_t)t, a.non_mask);
      if (lane == 0) mbar_expect(smem_u32(&full[b]), tile_bytes);
      __syncwarp();
      for (int r = lane; r < rows; r += 32)
        tma_load_5d(smem_u32(bufs + (size_t)b * tile_bytes + (size_t)r * row_bytes), &map,
                    (int)((tbase | row_off[r]) >> 3), smem_u32(&full[b]));
      if (k > 0) {
        const int pb = (k - 1) % NBUF;
        mbar_wait(smem_u32(&full[pb]), ((k - 1) / NBUF) & 1);
        for (int r = lane; r < rows; r += 32)
          tma_store_5d(&map, (int)((prev_base | row_off[r]) >> 3), smem_u32(bufs + (size_t)pb * tile_bytes + (size_t)r * row_bytes));
        bulk_commit();
      }
      prev_base = tbase;
    } else {
      if (k > 0) {
        const int pb = (k - 1) % NBUF;
        mbar_wait(smem_u32(&full[pb]), ((k - 1) / NBUF) & 1);
        for (int r = lane; r < rows; r += 32)
          tma_store_5d(&map, (int)((prev_base | row_off[r]) >> 3), smem_u32(bufs + (size_t)pb * tile_bytes + (size_t)r * row_bytes));
        bulk_commit();
      }
      break;
    }
  }
  bulk_wait_read<0>();
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// plain copy kernel for reference: every thread moves 16-byte elements in place (read + write), grid-stride
__global__ void plain_copy_kernel(c128* s, size_t total) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    c128 v = s[i];
    v.x += 0.0;
    s[i] = v;
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 26;
  cudaSetDevice(0);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) {
    printf("cuTensorMapEncodeTiled not available\n");
    return 1;
  }
  EncodeFn encode = (EncodeFn)fn;
  const size_t total = (size_t)1 << n;
  c128* d = nullptr;
  cudaMalloc(&d, total * 16);
  std::vector<c128> h(1 << 20);
  for (size_t i = 0; i < h.size(); ++i) h[i] = make_double2((double)i, -(double)i);
  // fill: amplitude i = (i, -i) (built on the device from a small host pattern is not needed: use a kernel-free memcpy loop)
  {
    std::vector<c128> chunk(1 << 20);
    for (size_t base = 0; base < total; base += chunk.size()) {
      for (size_t i = 0; i < chunk.size(); ++i) chunk[i] = make_double2((double)(base + i), -(double)(base + i));
      cudaMemcpy(d + base, chunk.data(), chunk.size() * 16, cudaMemcpyHostToDevice);
    }
  }
  unsigned long long* d_bad;
  cudaMalloc(&d_bad, 8);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  {
    plain_copy_kernel<<<sms * 8, 256>>>(d, total);
    cudaEventRecord(e0);
    for (int it = 0; it < 5; ++it) plain_copy_kernel<<<sms * 8, 256>>>(d, total);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("plain in-place copy kernel: %.3f ms/pass  %.0f GB/s (read + write)\n", ms / 5, 5 * 2.0 * total * 16 / ms / 1e6);
  }
  for (int m = 12; m <= 12; ++m)
    for (int l = 4; l <= 7; ++l)
      for (int e = 0; e <= 3; ++e)
        for (int pattern = 1; pattern < 3; ++pattern) {
          if (l + e > m) continue;
          // resident bits above the row: pattern 1 = the top bits, 2 = every other bit starting at l + 1; the e lowest of them
          // ride in the box as extra dimensions of size 2
          uint32_t res = 0;
          const int need = m - l;
          if (pattern == 1) for (int j = 0; j < need; ++j) res |= 1u << (n - 1 - j);
          if (pattern == 2) for (int j = 0, b = l + 1; j < need && b < n; ++j, b += 2) res |= 1u << b;
          if (__builtin_popcount(res) != need) continue;
          const uint32_t all = (uint32_t)(((uint64_t)1 << n) - 1) & ~((1u << l) - 1);
          Args a;
          a.n = n; a.m = m; a.l = l; a.e = e; a.non_mask = all & ~res; a.mismatches = d_bad; a.state = d;
          uint32_t rest = res;
          for (int j = 0; j < 3; ++j) a.ebit[j] = 0;
          for (int j = 0; j < e; ++j) { a.ebit[j] = __builtin_ctz(rest); rest &= rest - 1; }
          a.res_mask = rest;
          CUtensorMap map;
          cuuint64_t gdim[5] = {16, (cuuint64_t)(total / 8), 1, 1, 1};
          cuuint64_t gstride[4] = {128, 128, 128, 128};
          cuuint32_t box[5] = {16, (cuuint32_t)((1 << l) / 8), 1, 1, 1};
          cuuint32_t estr[5] = {1, 1, 1, 1, 1};
          for (int j = 0; j < e; ++j) { gdim[2 + j] = 2; gstride[1 + j] = (cuuint64_t)16 << a.ebit[j]; box[2 + j] = 2; }
          CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 5, d, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
          if (r != CUDA_SUCCESS) { printf("encode failed %d (l=%d e=%d)\n", (int)r, l, e); continue; }
          const size_t smem_bytes = (size_t)NBUF * (16 << m) + 64 + 4 * (1 << (m - l)) + 1024;
          cudaFuncSetAttribute(tma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
          cudaMemset(d_bad, 0, 8);
          a.mode = 0;
          tma_probe_kernel<<<sms, 128, smem_bytes>>>(map, a);
          unsigned long long bad = 0;
          cudaMemcpy(&bad, d_bad, 8, cudaMemcpyDeviceToHost);
          a.mode = 1;
          tma_probe_kernel<<<sms, 128, smem_bytes>>>(map, a);
          cudaEventRecord(e0);
          for (int it = 0; it < 5; ++it) tma_probe_kernel<<<sms, 128, smem_bytes>>>(map, a);
          cudaEventRecord(e1);
          cudaError_t err = cudaEventSynchronize(e1);
          float ms = 0;
          cudaEventElapsedTime(&ms, e0, e1);
          printf("m=%d l=%d e=%d (%4d B rows, %5d B per op, %3d ops/tile) pattern %d: layout mismatches %llu; %.3f ms/pass  %.0f GB/s (read + write) %s\n",
                 m, l, e, 16 << l, 16 << (l + e), 1 << (m - l - e), pattern, bad, ms / 5, 5 * 2.0 * total * 16 / ms / 1e6,
                 err == cudaSuccess ? "" : cudaGetErrorString(err));
        }
  // integrity: amplitude i still (i, -i)
  {
    std::vector<c128> chk(1 << 16);
    cudaMemcpy(chk.data(), d + (total - chk.size()), chk.size() * 16, cudaMemcpyDeviceToHost);
    size_t bad = 0;
    for (size_t i = 0; i < chk.size(); ++i) if (chk[i].x != (double)(total - chk.size() + i)) ++bad;
    printf("integrity after in-place passes: %zu bad of %zu\n", bad, chk.size());
  }
  return 0;
}
