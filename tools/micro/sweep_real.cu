// Micro-benchmark of the executor's real sweep code (qsb_exec.cuh qsb_sweep<K,G>) in isolation (developer tool).
#include <cstdio>
#include <cuda_runtime.h>
#include "qsb_exec.cuh"

extern __shared__ __align__(16) unsigned char smem[];
struct MiniEnv {
  int wid, W, wbits;
  __device__ __forceinline__ c128* tile() { return reinterpret_cast<c128*>(smem); }
};

template <int K, int G>
__global__ void __launch_bounds__(544, 1) bench(int ctlmode, int reps, int b0, int b1, int b2, int c0, int c1, int c2, long long* out, double* sink) {
  c128* tile = reinterpret_cast<c128*>(smem);
  qsb_desc* d = reinterpret_cast<qsb_desc*>(smem + (16 << 13));
  const int m = 13, T = blockDim.x, tid = threadIdx.x;
  for (int i = tid; i < (1 << m); i += T) tile[i] = make_double2(i * 1e-4, -i * 2e-4);
  if (tid == 0) {
    d->kind = QSB_D_SWEEP; d->gate = G; d->k = K;
    d->b[0] = b0; d->b[1] = b1; d->b[2] = b2; d->cls[0] = c0; d->cls[1] = c1; d->cls[2] = c2;
    for (int k = 0; k < 3; ++k) { d->P[k][0] = make_double2(0.6, 0.1 * k); d->P[k][1] = make_double2(-0.3, 0.2); d->P[k][2] = make_double2(0.3, 0.2); d->P[k][3] = make_double2(0.6, -0.1 * k); }
    int bb[3] = {b0, b1, b2};
    qsb_group_order(m, K, bb, d->pos);
    int wb = 31 - __clz(ctlmode ? T - 32 : T), hm = 0;
    for (int t = wb; t < m - K; ++t) hm |= 1 << d->pos[t];
    d->hmask = hm;
    for (int e = 0; e < 64; ++e) d->mat[e] = make_double2(e == 9 * (e / 9) ? 1.0 : 0.0, 0.0);
  }
  __syncthreads();
  MiniEnv env{tid, T, 31 - __clz(T)};
  long long t0 = clock64();
  if (ctlmode == 0) {
    for (int r = 0; r < reps; ++r) {
      qsb_sweep<K, G>(env, m, d);
      __syncthreads();
    }
  } else {
    // executor-like shape: the last warp is a control warp that hands out one descriptor per sweep
    const int W = T - 32;
    env.W = W; env.wbits = 31 - __clz(W);
    if (tid >= W) {
      for (int r = 0; r < reps; ++r) {
        if (r >= 4) asm volatile("bar.sync %0, %1;" ::"r"(5 + (r & 3)), "r"(T) : "memory");
        if (ctlmode == 2 && (tid & 31) == 0) { volatile double* p = (volatile double*)(d + 1); double x = p[0]; for (int q = 0; q < 40; ++q) x = x * 1.0000001 + 0.5; p[0] = x; }
        __syncwarp();
        asm volatile("bar.arrive %0, %1;" ::"r"(1 + (r & 3)), "r"(T) : "memory");
      }
    } else {
      for (int r = 0; r < reps; ++r) {
        asm volatile("bar.sync %0, %1;" ::"r"(1 + (r & 3)), "r"(T) : "memory");
        qsb_sweep<K, G>(env, m, d);
        asm volatile("bar.arrive %0, %1;" ::"r"(5 + (r & 3)), "r"(T) : "memory");
      }
    }
  }
  long long t1 = clock64();
  if (tid == 0) out[blockIdx.x] = (t1 - t0) / reps;
  if (tid == 0) sink[blockIdx.x] = tile[5].x;
}

template <int K, int G>
void run(const char* name, int threads, int ctlmode, int b0, int b1, int b2, int c0, int c1, int c2) {
  long long* out; double* sink;
  cudaMalloc(&out, 148 * 8); cudaMalloc(&sink, 148 * 8);
  size_t sm = (16 << 13) + sizeof(qsb_desc) + 64;
  cudaFuncSetAttribute(bench<K, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  bench<K, G><<<148, threads + (ctlmode ? 32 : 0), sm>>>(ctlmode, 200, b0, b1, b2, c0, c1, c2, out, sink);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, out, sizeof h, cudaMemcpyDeviceToHost);
  long long mx = 0, mn = 1LL << 60; for (int i = 0; i < 148; ++i) { if (h[i] > mx) mx = h[i]; if (h[i] < mn) mn = h[i]; }
  printf("%-30s ctl=%d thr=%4d bits=(%2d,%2d,%2d) cls=(%d,%d,%d) cycles/sweep min %6lld max %6lld  %s\n", name, ctlmode, threads, b0, b1, b2, c0, c1, c2, mn, mx, cudaGetErrorString(e));
  cudaFree(out); cudaFree(sink);
}

int main() {
  for (int ctl : {0, 1, 2}) {
    int thr = 256;
    run<2, QSB_G_CX>("CX none 12,11", thr, ctl, 12, 11, 0, 0, 0, 0);
    run<2, QSB_G_CX>("CX dense,dense 7,1", thr, ctl, 7, 1, 0, 3, 3, 0);
    run<3, QSB_G_CCX>("CCX none 12,11,10", thr, ctl, 12, 11, 10, 0, 0, 0);
    run<1, QSB_G_NONE>("flush1 dense b5", thr, ctl, 5, 0, 0, 3, 0, 0);
  }
  return 0;
}
