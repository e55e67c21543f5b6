// Micro-benchmark (developer tool): cost of exchanging half a 128 KB tile between two CTAs of a cluster of 8.
//   V0  pull: ld.shared::cluster into registers, cluster barrier, store locally (the executor's round-1 scheme)
//   V1  bulk push: cp.async.bulk.shared::cluster from my tile into the partner's staging buffer, mbarrier
//       complete_tx on the partner's side, cluster barrier, local copy staging -> tile
//   V2  register push: st.shared::cluster into the partner's staging buffer, cluster barrier, local copy
//   V3  bulk push only (no copy-back): raw DSMEM bulk bandwidth
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o xchg_bench.bin xchg_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef double2 c128;
extern __shared__ __align__(128) unsigned char smem[];

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t a, int rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
  return r;
}
__device__ __forceinline__ void cl_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ c128 ld_cluster(uint32_t a) {
  c128 v;
  asm volatile("ld.shared::cluster.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ void st_cluster(uint32_t a, c128 v) {
  asm volatile("st.shared::cluster.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(v.x), "d"(v.y) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}

// the executor's workers-only cluster barrier: bar.sync, CS parallel remote release-arrives, one poller, bar.sync
struct MSync {
  uint32_t local; int tid; uint32_t phase;
  __device__ __forceinline__ void sync() {
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (tid < 8) {
      uint32_t remote = mapa(local, tid);
      asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
    }
    if (tid == 0) {
      uint32_t done = 0;
      while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(local), "r"(phase) : "memory");
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    phase ^= 1;
  }
};
constexpr int M = 13, TILE = 1 << M, HALF = TILE / 2;   // amplitudes
constexpr int REGS = 16;

template <int V>
__global__ void __cluster_dims__(8, 1, 1) __launch_bounds__(256, 1) xchg(int reps, int gb, int chunk_bytes, long long* out) {
  c128* tile = reinterpret_cast<c128*>(smem);
  c128* stage = tile + TILE;                                   // 64 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(stage + HALF);
  const int tid = threadIdx.x, T = blockDim.x;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int peer = rank ^ (1 << gb), mybit = (rank >> gb) & 1;
  for (int i = tid; i < TILE; i += T) tile[i] = make_double2(rank * 100000.0 + i, 0.0);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 8;" ::"r"(smem_u32(bar + 1)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  MSync ms{smem_u32(bar + 1), tid, 0};
  __syncthreads();
  cl_sync();
  const uint32_t my_tile = smem_u32(tile), my_stage = smem_u32(stage), my_bar = smem_u32(bar);
  const uint32_t peer_tile = mapa(my_tile, peer), peer_stage = mapa(my_stage, peer), peer_bar = mapa(my_bar, peer);
  // my outgoing half: [out_off, out_off + HALF); contiguous (local bit m-1 swapped with the rank bit)
  const int out_off = (1 - mybit) * HALF, in_from = mybit * HALF;
  uint32_t parity = 0;
  long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    if (V == 0) {
      c128 val[REGS];
      cl_sync();
      for (int base = 0; base < HALF; base += REGS * T) {
#pragma unroll
        for (int e = 0; e < REGS; ++e) val[e] = ld_cluster(peer_tile + 16u * (in_from + base + e * T + tid));
        cl_sync();
#pragma unroll
        for (int e = 0; e < REGS; ++e) tile[out_off + base + e * T + tid] = val[e];
      }
    } else if (V == 1 || V == 3) {
      cl_sync();                                                        // partner's staging buffer is free
      if (tid == 0)
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(my_bar), "r"(HALF * 16) : "memory");
      const int nchunk = HALF * 16 / chunk_bytes;
      if (tid < 32) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy tile writes -> async proxy
        for (int c = tid; c < nchunk; c += 32)
          asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(peer_stage + c * chunk_bytes), "r"(my_tile + out_off * 16 + c * chunk_bytes), "r"(chunk_bytes), "r"(peer_bar) : "memory");
      }
      mbar_wait(my_bar, parity);
      parity ^= 1;
      if (V == 1) {
        cl_sync();                                                      // my own push has been read completely
        for (int i = tid; i < HALF; i += T) tile[out_off + i] = stage[i];
      }
    } else if (V == 4 || V == 5) {
      // executor-style remap: swizzled slots, local bit lb = chunk_bytes (reused argument), 16 registers per round
      const int lb = chunk_bytes;
      c128 val[REGS];
      if (V == 5) ms.sync(); else cl_sync();
      for (int base = 0; base < HALF; base += REGS * T) {
#pragma unroll
        for (int e = 0; e < REGS; ++e) {
          int g = base + e * T + tid;
          int i = (((g >> lb) << (lb + 1)) | (g & ((1 << lb) - 1))) | (mybit << lb);
          val[e] = ld_cluster(peer_tile + 16u * (i ^ ((i >> 3) & 7)));
        }
        if (V == 5) ms.sync(); else cl_sync();
#pragma unroll
        for (int e = 0; e < REGS; ++e) {
          int g = base + e * T + tid;
          int i = (((g >> lb) << (lb + 1)) | (g & ((1 << lb) - 1))) | ((1 - mybit) << lb);
          tile[i ^ ((i >> 3) & 7)] = val[e];
        }
      }
    } else if (V == 8) {
      for (int k = 0; k < 4; ++k) ms.sync();          // cost of four workers-only barriers
    } else if (V == 9) {
      for (int k = 0; k < 4; ++k) cl_sync();          // cost of four hardware cluster barriers
    } else if (V == 6) {
      // executor-style gflush: full tile, same slot on both sides, 2 rounds
      c128 val[REGS];
      const c128 pm = make_double2(0.6, 0.1), po = make_double2(0.3, -0.2);
      cl_sync();
      for (int base = 0; base < TILE; base += REGS * T) {
#pragma unroll
        for (int e = 0; e < REGS; ++e) {
          int i = base + e * T + tid;
          c128 a = tile[i], b = ld_cluster(peer_tile + 16u * i);
          val[e] = make_double2(pm.x * a.x - pm.y * a.y + po.x * b.x - po.y * b.y, pm.x * a.y + pm.y * a.x + po.x * b.y + po.y * b.x);
        }
        cl_sync();
#pragma unroll
        for (int e = 0; e < REGS; ++e) tile[base + e * T + tid] = val[e];
      }
    } else if (V == 10) {
      // swap by halves: each CTA of a pair swaps HALF of the groups in both directions (remote load + remote store),
      // so 32 KB flow in and 32 KB flow out per CTA instead of 64 KB in; no barrier between load and store
      const int lb = chunk_bytes;
      const int mine = (int)(rank > (uint32_t)peer);              // which half of the groups this CTA handles
      c128 vr[8], vl[8];
      cl_sync();
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        int g = mine * (HALF / 2) + e * T + tid;                   // contiguous split of the 4096 groups
        int b = (((g >> lb) << (lb + 1)) | (g & ((1 << lb) - 1)));
        int isrc = b | (mybit << lb), idst = b | ((1 - mybit) << lb);
        vr[e] = ld_cluster(peer_tile + 16u * (isrc ^ ((isrc >> 3) & 7)));
        vl[e] = tile[idst ^ ((idst >> 3) & 7)];
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        int g = mine * (HALF / 2) + e * T + tid;
        int b = (((g >> lb) << (lb + 1)) | (g & ((1 << lb) - 1)));
        int isrc = b | (mybit << lb), idst = b | ((1 - mybit) << lb);
        tile[idst ^ ((idst >> 3) & 7)] = vr[e];
        st_cluster(peer_tile + 16u * (isrc ^ ((isrc >> 3) & 7)), vl[e]);
      }
      cl_sync();
    } else if (V == 11) {
      // rank-bit flush by halves: a' and b' for half of the indices, one remote load + one remote store each
      const int mine = (int)(rank > (uint32_t)peer);
      const c128 p00 = make_double2(0.6, 0.1), p01 = make_double2(0.3, -0.2), p10 = make_double2(-0.3, -0.2), p11 = make_double2(0.6, -0.1);
      cl_sync();
      for (int base = 0; base < HALF; base += 8 * T) {
        c128 va[8], vb[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          int i = mine * HALF + base + e * T + tid;
          va[e] = tile[i];
          vb[e] = ld_cluster(peer_tile + 16u * i);
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          int i = mine * HALF + base + e * T + tid;
          c128 a = va[e], b = vb[e];
          tile[i] = make_double2(p00.x * a.x - p00.y * a.y + p01.x * b.x - p01.y * b.y, p00.x * a.y + p00.y * a.x + p01.x * b.y + p01.y * b.x);
          st_cluster(peer_tile + 16u * i, make_double2(p10.x * a.x - p10.y * a.y + p11.x * b.x - p11.y * b.y, p10.x * a.y + p10.y * a.x + p11.x * b.y + p11.y * b.x));
        }
      }
      cl_sync();
    } else if (V == 2) {
      cl_sync();
      c128 val[REGS];
      for (int base = 0; base < HALF; base += REGS * T) {
#pragma unroll
        for (int e = 0; e < REGS; ++e) val[e] = tile[out_off + base + e * T + tid];
#pragma unroll
        for (int e = 0; e < REGS; ++e) st_cluster(peer_stage + 16u * (base + e * T + tid), val[e]);
      }
      cl_sync();
      for (int i = tid; i < HALF; i += T) tile[out_off + i] = stage[i];
    }
  }
  long long t1 = clock64();
  cl_sync();
  if (tid == 0) out[blockIdx.x] = (t1 - t0) / reps;
  if (tid == 0 && blockIdx.x == 0 && reps == 1) out[200] = (long long)tile[out_off].x;   // sanity: partner's value
}

template <int V>
void run(const char* name, int gb, int chunk, int grid) {
  long long* d;
  cudaMalloc(&d, 256 * sizeof(long long));
  cudaMemset(d, 0, 256 * sizeof(long long));
  size_t sm = (size_t)TILE * 16 + HALF * 16 + 64;
  cudaFuncSetAttribute(xchg<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  xchg<V><<<grid, 256, sm>>>(1, gb, chunk, d);
  long long h[256];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  long long sanity = h[200];
  xchg<V><<<grid, 256, sm>>>(50, gb, chunk, d);
  cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0, mn = 1LL << 60;
  for (int i = 0; i < grid; ++i) { mx = h[i] > mx ? h[i] : mx; mn = h[i] < mn ? h[i] : mn; }
  printf("%-10s gb=%d chunk=%6d grid=%3d  cycles/exchange min %lld max %lld  (%.1f B/clk/SM)  sanity %lld  %s\n", name, gb, chunk, grid,
         mn, mx, 65536.0 / mx, sanity, cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  for (int grid : {120}) {
    for (int gb : {0, 2}) {
      run<0>("pull", gb, 0, grid);
      run<2>("regpush", gb, 0, grid);
      for (int chunk : {1024, 8192, 65536}) run<1>("bulk", gb, chunk, grid);
      run<3>("bulk-raw", gb, 8192, grid);
      for (int lb : {0, 1, 2, 3, 5, 8, 12}) run<4>("exec-remap", gb, lb, grid);
      run<6>("exec-gflush", gb, 0, grid);
      for (int lb : {0, 3, 12}) run<5>("remap-mbar", gb, lb, grid);
      for (int lb : {0, 3, 12}) run<10>("swap-halves", gb, lb, grid);
      run<11>("gflush-halves", gb, 0, grid);
      run<8>("4x mbar sync", gb, 0, grid);
      run<9>("4x hw sync", gb, 0, grid);
    }
  }
  return 0;
}
