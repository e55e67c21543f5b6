// Micro-benchmark: cost of one sweep over a 2^13-amplitude c128 tile in shared memory (developer tool).
#include <cstdio>
#include <cuda_runtime.h>
typedef double2 c128;
__device__ __forceinline__ int slot(int i) { return i ^ ((i >> 3) & 7); }
__device__ __forceinline__ int ins0(int g, int b) { return ((g >> b) << (b + 1)) | (g & ((1 << b) - 1)); }
__device__ __forceinline__ c128 cmul(c128 a, c128 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ c128 cfma(c128 a, c128 b, c128 c) { return make_double2(fma(a.x, b.x, fma(-a.y, b.y, c.x)), fma(a.x, b.y, fma(a.y, b.x, c.y))); }

extern __shared__ __align__(16) unsigned char smem[];

// MODE 0: copy sweep K=2 (CX-like swap), MODE 1: K=2 with two pending 2x2, MODE 2: K=3 with three pending,
// MODE 3: K=1 dense, MODE 4: plain linear copy (every element read+written), MODE 5: K=2 swap, touch only swapped pair
template <int MODE, bool SWZ>
__global__ void __launch_bounds__(1024, 1) bench(int reps, int b0, int b1, int b2, long long* out, double* sink) {
  c128* tile = reinterpret_cast<c128*>(smem);
  const int m = 13, T = blockDim.x, tid = threadIdx.x;
  for (int i = tid; i < (1 << m); i += T) tile[i] = make_double2(i * 1e-4, -i * 2e-4);
  c128 P[3][4];
  for (int k = 0; k < 3; ++k) { P[k][0] = make_double2(0.6, 0.1 * k); P[k][1] = make_double2(-0.3, 0.2); P[k][2] = make_double2(0.3, 0.2); P[k][3] = make_double2(0.6, -0.1 * k); }
  __syncthreads();
  long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    if (MODE == 4) {
      for (int i = tid; i < (1 << m); i += T) { c128 v = tile[i]; v.x += 1e-9; tile[i] = v; }
    } else if (MODE == 3) {
      const int cnt = 1 << (m - 1);
      for (int g = tid; g < cnt; g += T) {
        int i0 = ins0(g, b0), i1 = i0 | (1 << b0);
        int s0 = SWZ ? slot(i0) : i0, s1 = SWZ ? slot(i1) : i1;
        c128 a0 = tile[s0], a1 = tile[s1];
        tile[s0] = cfma(P[0][1], a1, cmul(P[0][0], a0));
        tile[s1] = cfma(P[0][3], a1, cmul(P[0][2], a0));
      }
    } else if (MODE == 0 || MODE == 1 || MODE == 5) {
      const int lo = b0 < b1 ? b0 : b1, hi = b0 < b1 ? b1 : b0;
      const int cnt = 1 << (m - 2);
#pragma unroll 2
      for (int g = tid; g < cnt; g += T) {
        int base = ins0(ins0(g, lo), hi);
        int i[4] = {base, base | (1 << b1), base | (1 << b0), base | (1 << b0) | (1 << b1)};
        int s[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) s[k] = SWZ ? slot(i[k]) : i[k];
        if (MODE == 5) { c128 x = tile[s[2]], y = tile[s[3]]; tile[s[2]] = y; tile[s[3]] = x; continue; }
        c128 a[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) a[k] = tile[s[k]];
        if (MODE == 1) {
#pragma unroll
          for (int k = 0; k < 2; ++k) {   // axis b1 (bit 0 of r)
            c128 l = a[2 * k], h = a[2 * k + 1];
            a[2 * k] = cfma(P[1][1], h, cmul(P[1][0], l)); a[2 * k + 1] = cfma(P[1][3], h, cmul(P[1][2], l));
          }
#pragma unroll
          for (int k = 0; k < 2; ++k) {   // axis b0 (bit 1 of r)
            c128 l = a[k], h = a[k + 2];
            a[k] = cfma(P[0][1], h, cmul(P[0][0], l)); a[k + 2] = cfma(P[0][3], h, cmul(P[0][2], l));
          }
        }
        c128 t = a[2]; a[2] = a[3]; a[3] = t;
#pragma unroll
        for (int k = 0; k < 4; ++k) tile[s[k]] = a[k];
      }
    } else if (MODE == 2) {
      int sb[3] = {b0, b1, b2};
      if (sb[0] > sb[1]) { int t = sb[0]; sb[0] = sb[1]; sb[1] = t; }
      if (sb[1] > sb[2]) { int t = sb[1]; sb[1] = sb[2]; sb[2] = t; }
      if (sb[0] > sb[1]) { int t = sb[0]; sb[0] = sb[1]; sb[1] = t; }
      const int cnt = 1 << (m - 3);
      for (int g = tid; g < cnt; g += T) {
        int base = ins0(ins0(ins0(g, sb[0]), sb[1]), sb[2]);
        int s[8]; c128 a[8];
#pragma unroll
        for (int r8 = 0; r8 < 8; ++r8) {
          int i = base | ((r8 & 4) ? 1 << b0 : 0) | ((r8 & 2) ? 1 << b1 : 0) | ((r8 & 1) ? 1 << b2 : 0);
          s[r8] = SWZ ? slot(i) : i;
          a[r8] = tile[s[r8]];
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const int bit = 4 >> k;
#pragma unroll
          for (int r8 = 0; r8 < 8; ++r8) {
            if (r8 & bit) continue;
            c128 l = a[r8], h = a[r8 | bit];
            a[r8] = cfma(P[k][1], h, cmul(P[k][0], l)); a[r8 | bit] = cfma(P[k][3], h, cmul(P[k][2], l));
          }
        }
        c128 t = a[6]; a[6] = a[7]; a[7] = t;
#pragma unroll
        for (int r8 = 0; r8 < 8; ++r8) tile[s[r8]] = a[r8];
      }
    }
    __syncthreads();
  }
  long long t1 = clock64();
  if (tid == 0) out[blockIdx.x] = (t1 - t0) / reps;
  if (tid == 0) sink[blockIdx.x] = tile[5].x;
}

template <int MODE, bool SWZ>
void run(const char* name, int threads, int b0, int b1, int b2) {
  long long* out; double* sink;
  cudaMalloc(&out, 148 * 8); cudaMalloc(&sink, 148 * 8);
  size_t sm = (16 << 13) + 1024;
  cudaFuncSetAttribute(bench<MODE, SWZ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  bench<MODE, SWZ><<<148, threads, sm>>>(200, b0, b1, b2, out, sink);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, out, sizeof h, cudaMemcpyDeviceToHost);
  long long mx = 0, mn = 1LL << 60; for (int i = 0; i < 148; ++i) { if (h[i] > mx) mx = h[i]; if (h[i] < mn) mn = h[i]; }
  printf("%-34s thr=%4d bits=(%2d,%2d,%2d) swz=%d  cycles/sweep min %6lld max %6lld  %s\n", name, threads, b0, b1, b2, (int)SWZ, mn, mx, cudaGetErrorString(e));
  cudaFree(out); cudaFree(sink);
}

int main() {
  for (int thr : {256, 512, 1024}) {
    run<4, true>("linear copy", thr, 0, 0, 0);
    run<3, true>("K=1 dense", thr, 5, 0, 0);
    run<3, true>("K=1 dense bit0", thr, 0, 0, 0);
    run<3, false>("K=1 dense bit0 noswz", thr, 0, 0, 0);
    run<0, true>("K=2 CX copy all 4", thr, 7, 3, 0);
    run<0, true>("K=2 CX copy all 4 low bits", thr, 1, 0, 0);
    run<0, false>("K=2 CX copy all 4 low bits noswz", thr, 1, 0, 0);
    run<5, true>("K=2 CX swap pair only", thr, 7, 3, 0);
    run<1, true>("K=2 CX + 2 pending", thr, 7, 3, 0);
    run<1, true>("K=2 CX + 2 pending low bits", thr, 2, 0, 0);
    run<2, true>("K=3 CCX + 3 pending", thr, 9, 4, 1);
    run<2, true>("K=3 CCX + 3 pending low", thr, 2, 1, 0);
    run<2, false>("K=3 CCX + 3 pending low noswz", thr, 2, 1, 0);
  }
  return 0;
}
