// sweep_real.cu -- the executor's own sweep code (qsb_do_sweep from qsb_exec.cuh) in isolation: 256 workers, one
// 128 KiB complex128 tile, descriptors prepared by thread 0 (no control warp, no ring).  Separates the cost of the sweep
// loop from the cost of the descriptor hand-off in the trajectory kernel.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../include -I../../quantum-simulator_b200/csrc -o sweep_real.bin sweep_real.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "qsb_exec.cuh"

extern __shared__ __align__(16) unsigned char smem[];

struct MiniEnv {
  typedef c128 amp;
  static constexpr int C = 1;
  static constexpr bool PROF = false;
  static constexpr int CL = 1;
  int clane = 0;
  int wid, W, wbits, lane, warp, nwarps;
  __device__ c128* tile() { return reinterpret_cast<c128*>(smem); }
  __device__ bool prof_on() { return false; }
  __device__ unsigned long long clock() { return 0; }
  __device__ void prof_add(int, unsigned long long) {}
};

// mode: 0 = bits >= 6 only, 1 = any bits; ndense = dense pending matrices per sweep (0..2); gate = QSB_G_*
__global__ void __launch_bounds__(256, 1) k(int sweeps, int mode, int ndense, int gate, int same_desc, long long* cyc, int fix) {
  MiniEnv env;
  env.wid = threadIdx.x; env.W = 256; env.wbits = 8; env.lane = threadIdx.x & 31; env.warp = threadIdx.x >> 5; env.nwarps = 8;
  c128* tile = env.tile();
  qsb_desc* ring = reinterpret_cast<qsb_desc*>(smem + 131072);
  for (int i = threadIdx.x; i < 8192; i += 256) tile[i] = qsb_c(1.0 / (1 + i), 0.5);
  const int m = 13;
  unsigned int rng = 12345u + blockIdx.x;
  __syncthreads();
  const long long t0 = clock64();
  for (int s = 0; s < sweeps; ++s) {
    qsb_desc* d = &ring[s & 1];
    if (threadIdx.x == 0 && (!same_desc || s < 2)) {
      rng = rng * 1664525u + 1013904223u;
      int b0 = mode ? (rng >> 8) % 13 : 6 + (rng >> 8) % 7;
      rng = rng * 1664525u + 1013904223u;
      int b1 = mode ? (rng >> 8) % 13 : 6 + (rng >> 8) % 7;
      if (b1 == b0) b1 = mode ? (b0 + 1) % 13 : 6 + (b0 - 6 + 1) % 7;
      d->kind = QSB_D_SWEEP; d->gate = gate; d->k = 2; d->flags = (ndense > 0 ? 1 : 0) | (ndense > 1 ? 2 : 0);
      d->b[0] = b0; d->b[1] = b1; d->b[2] = 0;
      d->cls[0] = ndense > 0 ? QSB_CLS_DENSE : QSB_CLS_NONE; d->cls[1] = ndense > 1 ? QSB_CLS_DENSE : QSB_CLS_NONE; d->cls[2] = 0;
      int hm;
      d->pos = qsb_group_order(m, (1u << b0) | (1u << b1), 8, m - 2, 3, &hm);
      d->hmask = hm;
      for (int e = 0; e < 32; ++e) { d->tabl[e] = qsb_deposit(e, d->pos, 5); if (e < 8) d->tabw[e] = qsb_deposit(e << 5, d->pos, 8); }
      for (int kk = 0; kk < 2; ++kk) { d->P[kk][0] = qsb_c(0.8, 0); d->P[kk][1] = qsb_c(0, -0.6); d->P[kk][2] = qsb_c(0, -0.6); d->P[kk][3] = qsb_c(0.8, 0); }
      uint64_t w = qsb_cls_set(qsb_cls_set(0, b0, d->cls[0]), b1, d->cls[1]);
      qsb_sweep_tables(env, d, w, gate, 2, b0, b1, 0);
    }
    __syncthreads();
    qsb_sweep_pro pro;
    qsb_sweep_prologue(env, d, pro);
    qsb_do_sweep(env, m, d, *reinterpret_cast<const qsb_desc_hdr*>(d), pro);
  }
  __syncthreads();
  if (threadIdx.x == 0) cyc[blockIdx.x] = clock64() - t0;
}

int main() {
  const int sweeps = 400, grid = 148;
  long long* cyc;
  cudaMallocManaged(&cyc, grid * sizeof(long long));
  const size_t sm = 131072 + 2 * sizeof(qsb_desc);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  for (int fix = 0; fix < 1; ++fix)
    for (int mode = 1; mode < 2; ++mode)
      for (int nd = 0; nd < 3; ++nd)
        for (int gate : {1, 2}) {
          const int same = 1;
          for (int rep = 0; rep < 2; ++rep) { k<<<grid, 256, sm>>>(sweeps, mode, nd, gate, same, cyc, fix); cudaDeviceSynchronize(); }
          double avg = 0;
          for (int i = 0; i < grid; ++i) avg += (double)cyc[i];
          printf("%s desc %s  bits %-5s dense pending %d  gate %s: %7.0f cycles / sweep\n", "qsb_do_sweep", same ? "fixed   " : "per sweep", mode ? "any" : ">= 6", nd,
                 gate == 1 ? "cx" : "cz", avg / grid / sweeps);
        }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
