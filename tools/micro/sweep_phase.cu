// sweep_phase.cu -- does a shared-memory sweep run faster when two worker groups are OUT OF PHASE?
// 256 threads sweep a 128 KiB tile of complex128 (2 steps of 16 amplitudes per thread per sweep, LDS.128 / STS.128,
// the executor's XOR swizzle), with an optional dense 2x2 on one of the two target bits:
//   mode 0: one group of 256 threads, bar.sync 256 after every sweep (the resident executor's shape)
//   mode 1: two groups of 128 threads on the two halves of the tile (split on bit 12), group barriers, started together
//   mode 2: as mode 1, group 1 starts `delay` cycles late
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o sweep_phase.bin sweep_phase.cu && ./sweep_phase.bin
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

extern __shared__ __align__(16) unsigned char smem[];
__device__ __forceinline__ int slot(int i) { return i ^ ((i >> 3) & 7); }
__device__ __forceinline__ int ins0(int g, int b) { return ((g >> b) << (b + 1)) | (g & ((1 << b) - 1)); }
__device__ __forceinline__ void bar(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

template <int MODE, int DENSE>
__global__ void __launch_bounds__(256, 1) k(int sweeps, int delay, double2* out, long long* cyc, double2 p0, double2 p1, double2 p2, double2 p3) {
  double2* tile = reinterpret_cast<double2*>(smem);
  const int tid = threadIdx.x;
  for (int i = tid; i < 8192; i += 256) tile[i] = make_double2(1.0 / (1 + i), 0.5);
  __syncthreads();
  const int grp = MODE ? (tid >> 7) : 0, W = MODE ? 128 : 256, wid = MODE ? (tid & 127) : tid;
  if (MODE == 2 && grp == 1) { long long t0 = clock64(); while (clock64() - t0 < delay) {} }
  const long long t0 = clock64();
  for (int s = 0; s < sweeps; ++s) {
    // target bits wander over 6..11 (the low bits stay with the lanes: conflict-free)
    const int b1 = 6 + (s % 5), b0 = b1 + 1 == 12 ? 6 : b1 + 1;
    const int lo = b0 < b1 ? b0 : b1, hi = b0 < b1 ? b1 : b0;
    int off[4];
    for (int r = 0; r < 4; ++r) off[r] = slot((((r >> 1) & 1) << b0) | ((r & 1) << b1));
    const int cnt = MODE ? 1024 : 2048;                 // groups of 4 amplitudes per (half) tile
    for (int g0 = wid; g0 < cnt; g0 += 4 * W) {
      double2 a[4][4];
      int base[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int g = g0 + j * W;
        int bs = ins0(ins0(g, lo), hi);
        if (MODE) bs = ins0(bs, 12) | (grp << 12);      // hi < 12: inserting bit 12 last keeps the order
        base[j] = slot(bs);
#pragma unroll
        for (int r = 0; r < 4; ++r) a[j][r] = tile[base[j] ^ off[r]];
      }
      if (DENSE) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            const double2 x = a[j][r], y = a[j][r + 2];
            a[j][r] = make_double2(p0.x * x.x - p0.y * x.y + p1.x * y.x - p1.y * y.y, p0.x * x.y + p0.y * x.x + p1.x * y.y + p1.y * y.x);
            a[j][r + 2] = make_double2(p2.x * x.x - p2.y * x.y + p3.x * y.x - p3.y * y.y, p2.x * x.y + p2.y * x.x + p3.x * y.y + p3.y * y.x);
          }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const double2 t = a[j][2]; a[j][2] = a[j][3]; a[j][3] = t;      // CX
#pragma unroll
        for (int r = 0; r < 4; ++r) tile[base[j] ^ off[r]] = a[j][r];
      }
    }
    if (MODE) bar(1 + grp, 128); else bar(1, 256);
  }
  const long long t1 = clock64();
  if (tid == 0 || tid == 128) cyc[blockIdx.x * 2 + (tid >> 7)] = t1 - t0;
  __syncthreads();
  if (out) for (int i = tid; i < 8192; i += 256) out[blockIdx.x * 8192 + i] = tile[i];
}

template <int MODE, int DENSE>
void run(const char* name, int delay) {
  const int sweeps = 400, grid = 148;
  long long* cyc;
  cudaMallocManaged(&cyc, grid * 2 * sizeof(long long));
  cudaFuncSetAttribute(k<MODE, DENSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072);
  const double c = 0.8, sn = 0.6;
  for (int rep = 0; rep < 2; ++rep) {
    k<MODE, DENSE><<<grid, 256, 131072>>>(sweeps, delay, nullptr, cyc, make_double2(c, 0), make_double2(0, -sn), make_double2(0, -sn), make_double2(c, 0));
    cudaDeviceSynchronize();
  }
  double avg = 0;
  for (int i = 0; i < grid; ++i) avg += (double)(cyc[2 * i] > cyc[2 * i + 1] ? cyc[2 * i] : cyc[2 * i + 1]);
  printf("%-46s delay %5d: %7.0f cycles / sweep  (%5.1f B/clk/SM of 128)\n", name, delay, avg / grid / sweeps, 262144.0 / (avg / grid / sweeps));
  cudaFree(cyc);
}

int main() {
  run<0, 0>("one group of 256, CX only", 0);
  run<1, 0>("two groups of 128 in phase, CX only", 0);
  for (int d : {500, 1000, 1500, 2000}) run<2, 0>("two groups of 128 out of phase, CX only", d);
  run<0, 1>("one group of 256, dense 2x2 + CX", 0);
  run<1, 1>("two groups of 128 in phase, dense 2x2 + CX", 0);
  for (int d : {500, 1000, 1500, 2000, 3000}) run<2, 1>("two groups of 128 out of phase, dense 2x2 + CX", d);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
