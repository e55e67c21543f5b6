// qsb_kernels.cuh -- CUDA side of the executor (DeviceEnv + trajectory kernel) and the
// reduction kernels over stored states.  sm_100a only.
#pragma once

#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include "qsb_exec.cuh"

namespace cg = cooperative_groups;

#ifndef QSB_MAX_WORKERS
#define QSB_MAX_WORKERS 256
#endif
#define QSB_CTL_THREADS 32
#define QSB_DEC_THREADS 32       // decode warp: stages the op list one chunk ahead of the control warp
#define QSB_SMEM_EXTRA (sizeof(qsb_ctl))

// the one dynamic shared-memory block of the executor kernel: [tile | qsb_ctl]
extern __shared__ __align__(16) unsigned char qsb_smem[];

// named barriers (id 0 is left to __syncthreads, which the executor never uses)
#define QSB_BAR_FULL 1                    // + ring slot: control arrives, workers sync
#define QSB_BAR_EMPTY (1 + QSB_RING)      // + ring slot: workers arrive, control syncs
#define QSB_BAR_WORKERS (1 + 2 * QSB_RING)
#define QSB_BAR_ALL (2 + 2 * QSB_RING)

__device__ __forceinline__ void qsb_bar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void qsb_bar_arrive(int id, int count) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void qsb_cluster_arrive() {
  asm volatile("barrier.cluster.arrive.release;" ::: "memory");
}
__device__ __forceinline__ void qsb_cluster_wait() {
  asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
}

// threads [0, W) are workers, [W, W + 32) is the control warp, [W + 32, W + 64) the decode warp
// PF: the cycle counters of qsb_debug_profile are compiled in (a separate instantiation, launched only while
// profiling is enabled: the counters cost registers, local memory and ~2 % of the run time)
template <int CS, class A = c128, bool PF = false>
struct DeviceEnv {
  typedef A amp;
  static constexpr int C = CS;
  static constexpr bool PROF = PF;
  static constexpr int CL = QSB_CTL_THREADS;
  int wid, W, wbits, rank, m_;
  int lane, warp, nwarps;      // worker warp geometry
  int clane;                   // lane inside the control warp
  bool lead;                   // the control lane that writes shared state
  int role;                    // 0 worker, 1 control warp, 2 decode warp

  __device__ DeviceEnv(int m) {
    W = (int)blockDim.x - QSB_CTL_THREADS - QSB_DEC_THREADS;      // a power of two >= 32
    wbits = 31 - __clz(W);
    const int tid = threadIdx.x;
    wid = tid < W ? tid : -1;
    lane = tid & 31;
    warp = tid >> 5;
    nwarps = W >> 5;
    clane = (tid - W) & 31;                      // lane inside the control warp / the decode warp
    lead = clane == 0;
    role = tid < W ? 0 : (tid < W + QSB_CTL_THREADS ? 1 : 2);
    rank = (CS > 1) ? (int)cg::this_cluster().block_rank() : 0;
    m_ = m;
  }
  // pointers are re-derived from the shared symbol at every use so that loads/stores stay LDS/STS
  __device__ __forceinline__ A* tile() { return reinterpret_cast<A*>(qsb_smem); }
  __device__ __forceinline__ qsb_ctl* ctl() { return reinterpret_cast<qsb_ctl*>(qsb_smem + (sizeof(A) << m_)); }
  __device__ __forceinline__ const A* peer_tile(int r) {
    if (CS > 1) return cg::this_cluster().map_shared_rank(tile(), r);
    return tile();
  }
  __device__ __forceinline__ void fence_cluster() { asm volatile("fence.acq_rel.cluster;" ::: "memory"); }
  __device__ __forceinline__ A* peer_tile_w(int r) {
    if (CS > 1) return cg::this_cluster().map_shared_rank(tile(), r);
    return tile();
  }
  __device__ __forceinline__ const qsb_ctl* peer_ctl(int r) {
    if (CS > 1) return cg::this_cluster().map_shared_rank(ctl(), r);
    return ctl();
  }
  __device__ __forceinline__ qsb_ctl* peer_ctl_w(int r) {
    if (CS > 1) return cg::this_cluster().map_shared_rank(ctl(), r);
    return ctl();
  }
  __device__ __forceinline__ void atomic_add(double* p, double v) { atomicAdd(p, v); }
  __device__ __forceinline__ double warp_sum(double x) {
    for (int off = 16; off > 0; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
    return x;
  }
  __device__ __forceinline__ unsigned long long clock() { return (unsigned long long)clock64(); }
  unsigned long long* prof_;
  __device__ __forceinline__ bool prof_on() { return PF && prof_ != nullptr && threadIdx.x == 0; }
  __device__ __forceinline__ void prof_add(int slot, unsigned long long v) { prof_[(size_t)blockIdx.x * QSB_PROF_WORDS + slot] += v; }
  __device__ __forceinline__ int cta_id() { return (int)blockIdx.x; }
  // control-warp collectives: lanes with the same key; bit `slot` of the result = this lane's predicate (lane == slot)
  __device__ __forceinline__ uint32_t match_any(int key) { return __match_any_sync(0xffffffffu, key); }
  __device__ __forceinline__ uint32_t ballot_slot(int pred, int slot) { (void)slot; return __ballot_sync(0xffffffffu, pred); }
  __device__ __forceinline__ uint32_t or_reduce(uint32_t x) { return __reduce_or_sync(0xffffffffu, x); }
  // lane 0 of the control warp -> every lane
  __device__ __forceinline__ int bcast_i(int x) { return __shfl_sync(0xffffffffu, x, 0); }
  __device__ __forceinline__ uint64_t bcast_u64(uint64_t x) {
    const uint32_t lo = __shfl_sync(0xffffffffu, (uint32_t)x, 0), hi = __shfl_sync(0xffffffffu, (uint32_t)(x >> 32), 0);
    return ((uint64_t)hi << 32) | lo;
  }
  // ---- decode warp <-> control warp: one mbarrier per chunk buffer and direction (count 1: the lead lane arrives
  // after a __syncwarp, every lane of the other warp polls the phase)
  __device__ __forceinline__ uint32_t dbar_addr(int k) { return (uint32_t)__cvta_generic_to_shared(&ctl()->dbar[k]); }
  __device__ __forceinline__ void dbar_arrive(int k) {
    __syncwarp();
    if (lead) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(dbar_addr(k)) : "memory");
  }
  __device__ __forceinline__ void dbar_wait(int k, uint32_t parity) {
    const uint32_t addr = dbar_addr(k);
    uint32_t done = 0;
    while (!done) {
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    }
    __syncwarp();
  }
  __device__ __forceinline__ void dec_publish(int b) { dbar_arrive(b); }
  __device__ __forceinline__ void dec_wait_full(int b, uint32_t parity) { dbar_wait(b, parity); }
  __device__ __forceinline__ void dec_release(int b) { dbar_arrive(2 + b); }
  __device__ __forceinline__ void dec_wait_empty(int b, uint32_t parity) { dbar_wait(2 + b, parity); }
  // ---- descriptor ring
  __device__ __forceinline__ void ring_wait_empty(int s) { __syncwarp(); qsb_bar_sync(QSB_BAR_EMPTY + s, W + QSB_CTL_THREADS); }
  __device__ __forceinline__ void ring_publish(int s) {
    __syncwarp();                              // bar.arrive orders this warp's prior shared-memory writes for the
    qsb_bar_arrive(QSB_BAR_FULL + s, W + QSB_CTL_THREADS);   // threads that complete the barrier: no extra fence
  }
  __device__ __forceinline__ void ring_wait_full(int s) { qsb_bar_sync(QSB_BAR_FULL + s, W + QSB_CTL_THREADS); }
  __device__ __forceinline__ void ring_release(int s) { qsb_bar_arrive(QSB_BAR_EMPTY + s, W + QSB_CTL_THREADS); }
  // ---- barriers
  __device__ __forceinline__ void sync_workers() { qsb_bar_sync(QSB_BAR_WORKERS, W); }
  __device__ __forceinline__ void sync_control() { __syncwarp(); }
  // Cluster barrier among the WORKERS of all CTAs (the control warps run ahead and never take part): the CTA's
  // workers meet on a named barrier, one of them arrives on every peer's mbarrier (expected count = cluster
  // size), then every worker waits for its own CTA's mbarrier phase.
  int xphase;
  __device__ __forceinline__ uint32_t xbar_addr() { return (uint32_t)__cvta_generic_to_shared(&ctl()->xbar); }
  __device__ __forceinline__ void cluster_sync_w() {
    if (CS == 1) { qsb_bar_sync(QSB_BAR_WORKERS, W); return; }
    qsb_bar_sync(QSB_BAR_WORKERS, W);
    const uint32_t local = xbar_addr();
    if (wid < CS) {                            // worker r signals peer r (release arrives issued one after the other
      uint32_t remote;                         // by a single thread cost ~500 cycles each)
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(wid));
      asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
    }
    if (wid == 0) {
      // one thread waits for the phase (256 pollers would fight the arrivals for the barrier unit) ...
      uint32_t done = 0;
      while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(local), "r"((uint32_t)xphase) : "memory");
      }
    }
    qsb_bar_sync(QSB_BAR_WORKERS, W);          // ... and releases the others
    xphase ^= 1;
  }
  // every thread of every CTA (kernel start: the mbarriers are initialised; kernel end: nobody reads a peer's
  // shared memory any more)
  __device__ __forceinline__ void cluster_exit() { __syncwarp(); qsb_cluster_arrive(); qsb_cluster_wait(); }
  // workers -> local control warp hand-off of a reduction result
  __device__ __forceinline__ void handoff_w() { qsb_bar_sync(QSB_BAR_ALL, W + QSB_CTL_THREADS); }
  __device__ __forceinline__ void handoff_c() { __syncwarp(); qsb_bar_sync(QSB_BAR_ALL, W + QSB_CTL_THREADS); }
};

template <int CS, class A, bool PF>
__global__ void __launch_bounds__(QSB_MAX_WORKERS + QSB_CTL_THREADS + QSB_DEC_THREADS, 1)
qsb_traj_kernel(const __grid_constant__ qsb_exec_args a) {
  DeviceEnv<CS, A, PF> env(a.m);
  env.prof_ = a.prof;
  env.xphase = 0;
  if (threadIdx.x == 0) {
    for (int k = 0; k < 4; ++k) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(env.dbar_addr(k)) : "memory");
    if (CS > 1) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(env.xbar_addr()), "r"(CS) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (CS > 1) env.cluster_exit();             // every mbarrier of the cluster is initialised before its first use
  const int64_t first = a.tile_bits ? (int64_t)blockIdx.x : (int64_t)(blockIdx.x / CS);
  const int64_t stride = a.tile_bits ? (int64_t)gridDim.x : (int64_t)(gridDim.x / CS);
  if (env.role == 0) qsb_worker_loop(env, a);
  else if (env.role == 1) qsb_control_loop(env, a, first, stride);
  else qsb_decode_loop(env, a, first, stride);
}

// ---------------------------------------------------------------------------------------
// block-wide sum of NV doubles (result valid in every thread); scratch >= 32 * NV doubles
template <int NV>
__device__ __forceinline__ void qsb_block_sum(double* v, double* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double x = v[k];
    for (int off = 16; off > 0; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
    if (lane == 0) scratch[warp * NV + k] = x;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double s = 0.0;
    for (int w = 0; w < nw; ++w) s += scratch[w * NV + k];
    v[k] = s;
  }
}

// The reductions read states in the context's amplitude type (complex128, or complex64 in c64 mode) and
// accumulate in double.
// |a|^2 elementwise (state_vector.py:36-39)
template <class A>
__global__ void qsb_probs_kernel(const A* __restrict__ psi, double* __restrict__ out, int64_t total) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    c128 a = qsb_wide(psi[i]);
    out[i] = a.x * a.x + a.y * a.y;
  }
}

// out[i] += sum_t |psi_t[i]|^2
template <class A>
__global__ void qsb_probs_sum_kernel(const A* __restrict__ psi, double* __restrict__ out, int64_t dim, int64_t count) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < dim; i += (int64_t)gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int64_t t = 0; t < count; ++t) {
      c128 a = qsb_wide(psi[t * dim + i]);
      s += a.x * a.x + a.y * a.y;
    }
    out[i] += s;
  }
}

// StateVector.measure_all (state_vector.py:107-113): first index whose running probability mass exceeds
// u * total  (== searchsorted(cumsum(p / p.sum()) / cdf[-1], u, side='right'))
template <class A>
__global__ void qsb_sample_kernel(const A* __restrict__ psi, const double* __restrict__ u, int64_t* __restrict__ out,
                                  int64_t dim) {
  __shared__ double chunk_sum[256];
  __shared__ double prefix[257];
  const A* s = psi + blockIdx.x * dim;
  const int T = blockDim.x, tid = threadIdx.x;
  const int64_t per = (dim + T - 1) / T, lo = tid * per, hi = (lo + per < dim) ? lo + per : dim;
  double acc = 0.0;
  for (int64_t i = lo; i < hi; ++i) { c128 a = qsb_wide(s[i]); acc += a.x * a.x + a.y * a.y; }
  chunk_sum[tid] = acc;
  __syncthreads();
  if (tid == 0) {
    double run = 0.0;
    for (int k = 0; k < T; ++k) { prefix[k] = run; run += chunk_sum[k]; }
    prefix[T] = run;
    out[blockIdx.x] = dim - 1;     // fallback when rounding leaves the target at the very end
  }
  __syncthreads();
  const double target = u[blockIdx.x] * prefix[T];
  if (lo < hi && prefix[tid] <= target && (target < prefix[tid + 1] || tid == T - 1)) {
    // the in-chunk running sum can differ from chunk_sum by rounding: fall back to the chunk's last
    // amplitude with weight instead of leaving the sample unassigned
    double run = prefix[tid];
    int64_t pick = -1, last_nz = hi - 1;
    for (int64_t i = lo; i < hi; ++i) {
      c128 a = qsb_wide(s[i]);
      double p = a.x * a.x + a.y * a.y;
      run += p;
      if (p > 0.0) last_nz = i;
      if (run > target) { pick = i; break; }
    }
    out[blockIdx.x] = pick >= 0 ? pick : last_nz;
  }
}

// out[t] = sum_i conj(a_t[i]) * b_t[i]   (np.vdot, analysis.py:40)
template <class A>
__global__ void qsb_overlap_kernel(const A* __restrict__ a, const A* __restrict__ b, int64_t b_stride,
                                   c128* __restrict__ out, int64_t dim) {
  __shared__ double scratch[64];
  const A* x = a + blockIdx.x * dim;
  const A* y = b + blockIdx.x * b_stride;
  double v[2] = {0.0, 0.0};
  for (int64_t i = threadIdx.x; i < dim; i += blockDim.x) {
    c128 p = qsb_wide(x[i]), q = qsb_wide(y[i]);
    v[0] += p.x * q.x + p.y * q.y;
    v[1] += p.x * q.y - p.y * q.x;
  }
  qsb_block_sum<2>(v, scratch);
  if (threadIdx.x == 0) out[blockIdx.x] = make_double2(v[0], v[1]);
}

// large states: stage 1, CTA c sums the slice [c * per, (c + 1) * per) of one state pair into part[c];
// stage 2 (one CTA) adds the partials in index order, so the result does not depend on scheduling
template <class A>
__global__ void qsb_overlap_partial_kernel(const A* __restrict__ x, const A* __restrict__ y, int64_t dim,
                                           int64_t per, c128* __restrict__ part) {
  __shared__ double scratch[64];
  const int64_t lo = blockIdx.x * per, hi = (lo + per < dim) ? lo + per : dim;
  double v[2] = {0.0, 0.0};
  for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    c128 p = qsb_wide(x[i]), q = qsb_wide(y[i]);
    v[0] += p.x * q.x + p.y * q.y;
    v[1] += p.x * q.y - p.y * q.x;
  }
  qsb_block_sum<2>(v, scratch);
  if (threadIdx.x == 0) part[blockIdx.x] = make_double2(v[0], v[1]);
}
__global__ void qsb_overlap_final_kernel(const c128* __restrict__ part, int n_part, c128* __restrict__ out) {
  __shared__ double scratch[64];
  double v[2] = {0.0, 0.0};
  for (int i = threadIdx.x; i < n_part; i += blockDim.x) { v[0] += part[i].x; v[1] += part[i].y; }
  qsb_block_sum<2>(v, scratch);
  if (threadIdx.x == 0) out[0] = make_double2(v[0], v[1]);
}

// (p_even, p_odd) per mask (qec.py:466-484); up to 8 masks per launch
template <class A>
__global__ void qsb_parity_kernel(const A* __restrict__ psi, int64_t dim, const uint64_t* __restrict__ masks,
                                  int n_masks, double* __restrict__ out) {
  __shared__ double scratch[32 * 16];
  const A* s = psi + blockIdx.x * dim;
  double v[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) v[k] = 0.0;
  uint64_t mk[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) mk[k] = k < n_masks ? masks[k] : 0;
  for (int64_t i = threadIdx.x; i < dim; i += blockDim.x) {
    c128 a = qsb_wide(s[i]);
    double p = a.x * a.x + a.y * a.y;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      int odd = __popcll((uint64_t)i & mk[k]) & 1;
      v[2 * k] += odd ? 0.0 : p;
      v[2 * k + 1] += odd ? p : 0.0;
    }
  }
  qsb_block_sum<16>(v, scratch);
  if (threadIdx.x == 0)
    for (int k = 0; k < n_masks; ++k) {
      out[(blockIdx.x * (int64_t)n_masks + k) * 2] = v[2 * k];
      out[(blockIdx.x * (int64_t)n_masks + k) * 2 + 1] = v[2 * k + 1];
    }
}

// 1-qubit RDMs: block (t, q) -> rdm1[t][q][2][2]  (state_vector.py:121-140)
template <class A>
__global__ void qsb_rdm1_kernel(const A* __restrict__ psi, int n, c128* __restrict__ out) {
  __shared__ double scratch[32 * 4];
  const int64_t dim = (int64_t)1 << n;
  const int q = blockIdx.x % n;
  const int64_t t = blockIdx.x / n;
  const int b = n - 1 - q;
  const A* s = psi + t * dim;
  double v[4] = {0, 0, 0, 0};
  for (int64_t g = threadIdx.x; g < dim / 2; g += blockDim.x) {
    int64_t i0 = ((g >> b) << (b + 1)) | (g & (((int64_t)1 << b) - 1));
    c128 a0 = qsb_wide(s[i0]), a1 = qsb_wide(s[i0 | ((int64_t)1 << b)]);
    v[0] += a0.x * a0.x + a0.y * a0.y;
    v[1] += a1.x * a1.x + a1.y * a1.y;
    v[2] += a0.x * a1.x + a0.y * a1.y;      // a0 conj(a1)
    v[3] += a0.y * a1.x - a0.x * a1.y;
  }
  qsb_block_sum<4>(v, scratch);
  if (threadIdx.x == 0) {
    c128* o = out + (t * n + q) * 4;
    o[0] = make_double2(v[0], 0.0);
    o[1] = make_double2(v[2], v[3]);
    o[2] = make_double2(v[2], -v[3]);
    o[3] = make_double2(v[1], 0.0);
  }
}

// 2-qubit RDMs: block (t, pair) -> rdm2[t][pair][4][4], row index = (bit_i << 1 | bit_j), i < j
// rho[r][c] = sum_env psi[r,env] conj(psi[c,env])   (analysis.py:159-166)
template <class A>
__global__ void qsb_rdm2_kernel(const A* __restrict__ psi, int n, int npairs, c128* __restrict__ out) {
  __shared__ double scratch[32 * 16];
  const int64_t dim = (int64_t)1 << n;
  int pair = blockIdx.x % npairs;
  const int64_t t = blockIdx.x / npairs;
  int qi = 0, rem = pair;
  while (rem >= n - 1 - qi) { rem -= n - 1 - qi; ++qi; }
  const int qj = qi + 1 + rem;
  const int bh = n - 1 - qi, bl = n - 1 - qj;    // bh > bl
  const A* s = psi + t * dim;
  double v[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) v[k] = 0.0;
  for (int64_t g = threadIdx.x; g < dim / 4; g += blockDim.x) {
    int64_t i = ((g >> bl) << (bl + 1)) | (g & (((int64_t)1 << bl) - 1));
    i = ((i >> bh) << (bh + 1)) | (i & (((int64_t)1 << bh) - 1));
    c128 a[4];
    a[0] = qsb_wide(s[i]);
    a[1] = qsb_wide(s[i | ((int64_t)1 << bl)]);
    a[2] = qsb_wide(s[i | ((int64_t)1 << bh)]);
    a[3] = qsb_wide(s[i | ((int64_t)1 << bh) | ((int64_t)1 << bl)]);
    // diag (4 reals) + upper triangle (6 complex)
    int k = 0;
#pragma unroll
    for (int r = 0; r < 4; ++r) v[k++] += a[r].x * a[r].x + a[r].y * a[r].y;
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = r + 1; c < 4; ++c) {
        v[k++] += a[r].x * a[c].x + a[r].y * a[c].y;
        v[k++] += a[r].y * a[c].x - a[r].x * a[c].y;
      }
  }
  qsb_block_sum<16>(v, scratch);
  if (threadIdx.x == 0) {
    c128* o = out + (t * npairs + pair) * 16;
    int k = 4;
    for (int r = 0; r < 4; ++r) o[r * 4 + r] = make_double2(v[r], 0.0);
    for (int r = 0; r < 4; ++r)
      for (int c = r + 1; c < 4; ++c) {
        o[r * 4 + c] = make_double2(v[k], v[k + 1]);
        o[c * 4 + r] = make_double2(v[k], -v[k + 1]);
        k += 2;
      }
  }
}

__device__ __forceinline__ void qsb_dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// ---- all 1- and 2-qubit RDMs of a state from ONE read, on the FP64 tensor cores -------------------------------------
// (analysis.py:120-166 builds the 2^n x 2^n outer product per call; state_vector.py:121-140; the event detector asks
// for all n(n-1)/2 pairs of every state, analysis.py:315-333.)
// rho_ij[r][c] = sum_env M[r][env] conj(M[c][env]) is the Gram matrix of the 4 x 2^(n-2) complex matrix M of the state
// viewed with bits (i, j) as the row index.  Stacked as the real 8 x 2^(n-2) matrix X = [Re M; Im M] it is one real 8 x 8
// Gram X X^T -- exactly the shape of mma.sync.m8n8k4.f64 with B = A^T, so ONE register per lane feeds both operands:
// lane (g, t) holds X[g][k0 + t].  rho = (G_rr + G_ii) + i (G_ir - G_ri) over the four 4 x 4 blocks of G.
// One CTA keeps one state in shared memory (n <= 13; XOR-folded so that the four rows of a lane group hit different
// banks) and its eight warps share the pairs; 1-qubit RDMs fall out of the pair (q, q+1) [(n-2, n-1) for the last
// qubit] by a partial trace.  HBM sees the state once instead of n(n-1)/2 + n times.
// Shared-memory slot of amplitude i: the parity of the index bits above bit 2 is XORed into bit 2.  The eight amplitudes a
// half-warp reads differ in three index bits -- the row bit of the pair and the two lowest other bits -- so whichever bit
// the pair uses, they land in eight different 16-byte bank groups.  (Linear over XOR: swz(a ^ b) = swz(a) ^ swz(b).)
__device__ __forceinline__ int qsb_rdm_swz(int i) { return i ^ ((__popc(i >> 3) & 1) << 2); }

template <class A>
__global__ void __launch_bounds__(256) qsb_rdm_gram_kernel(const A* __restrict__ psi, int n, int npairs, int64_t count,
                                                            c128* __restrict__ rdm1, c128* __restrict__ rdm2) {
  extern __shared__ __align__(16) unsigned char qsb_rdm_smem[];
  c128* st = reinterpret_cast<c128*>(qsb_rdm_smem);
  const int dim = 1 << n;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  // rows of X interleave the parts: row 2r = Re M[r], row 2r + 1 = Im M[r].  A half-warp (g = 0..3) then reads BOTH
  // 8-byte halves of eight amplitudes -- 128 contiguous-per-amplitude bytes, all 32 banks -- instead of the real halves of
  // sixteen (which can use only half of the banks: ncu showed 58 % of the wavefronts as conflict replays)
  const int r = g >> 1, part = g & 1;
  for (int64_t s = blockIdx.x; s < count; s += gridDim.x) {
    __syncthreads();                                   // the previous state's readers are done
    const A* src = psi + s * dim;
    for (int i = threadIdx.x; i < dim; i += 256) st[qsb_rdm_swz(i)] = qsb_wide(src[i]);
    __syncthreads();
    for (int p = warp; p < npairs; p += 8) {
      int qi = 0, rem = p;
      while (rem >= n - 1 - qi) { rem -= n - 1 - qi; ++qi; }
      const int qj = qi + 1 + rem;
      if (!rdm2 && rem != 0 && p != npairs - 1) continue;    // 1-qubit RDMs only: n - 1 pairs carry them all
      const int bh = n - 1 - qi, bl = n - 1 - qj;      // index bits of qubits i (row MSB) and j; bh > bl
      // the two lowest index bits outside {bh, bl} carry t (the k index inside one DMMA); the other n - 4 free bits are
      // walked by a masked increment
      int p0 = 0;
      while (p0 == bl || p0 == bh) ++p0;
      int p1 = p0 + 1;
      while (p1 == bl || p1 == bh) ++p1;
      const int fixed = (1 << bh) | (1 << bl) | (1 << p0) | (1 << p1);
      const int freemask = (dim - 1) & ~fixed;
      const int lane_off = ((r >> 1) << bh) | ((r & 1) << bl) | ((t & 1) << p0) | ((t >> 1) << p1);
      const int lane_swz = qsb_rdm_swz(lane_off);      // the fold is linear over XOR and the bit sets are disjoint
      // two independent accumulator pairs (even / odd k steps): one dependent DMMA chain per warp left the tensor pipe
      // half idle (ncu: dmma sub-pipe 50 % active, math_pipe_throttle the top stall)
      double d0 = 0.0, d1 = 0.0, f0 = 0.0, f1 = 0.0;
      int base = 0;
      const unsigned char* lane_ptr = qsb_rdm_smem + 8 * part;
      const int steps = dim >> 4;                       // >= 1; even for n >= 5
      int k = 0;
#pragma unroll 4
      for (; k + 1 < steps; k += 2) {
        const int slot_a = qsb_rdm_swz(base) ^ lane_swz;
        base = ((base | fixed) + 1) & freemask;
        const int slot_b = qsb_rdm_swz(base) ^ lane_swz;
        base = ((base | fixed) + 1) & freemask;
        const double va = *reinterpret_cast<const double*>(lane_ptr + (slot_a << 4));
        const double vb = *reinterpret_cast<const double*>(lane_ptr + (slot_b << 4));
        qsb_dmma(d0, d1, va, va);
        qsb_dmma(f0, f1, vb, vb);
      }
      if (k < steps) {
        const double va = *reinterpret_cast<const double*>(lane_ptr + ((qsb_rdm_swz(base) ^ lane_swz) << 4));
        qsb_dmma(d0, d1, va, va);
      }
      d0 += f0;
      d1 += f1;
      // lane (g, t): d0 = G[g][2t], d1 = G[g][2t + 1].  With g = 2r: d0 = Re_r.Re_c, d1 = Re_r.Im_c (c = t); the lane
      // four places on (g + 1) holds Im_r.Re_c, Im_r.Im_c:  rho[r][c] = (ReRe + ImIm) + i (ImRe - ReIm)
      const double o0 = __shfl_down_sync(0xffffffffu, d0, 4);
      const double o1 = __shfl_down_sync(0xffffffffu, d1, 4);
      c128 e = make_double2(d0 + o1, o0 - d1);
      if (!(g & 1)) {
        if (r == t) e.y = 0.0;                         // the diagonal of a Gram matrix is real
        if (rdm2) rdm2[(s * npairs + p) * 16 + r * 4 + t] = e;
      }
      if (rdm1) {
        // partial traces: rho_i[a][b] = sum_c rho_ij[(a, c)][(b, c)] from pair (qi, qi + 1); the last qubit takes
        // rho_j[a][b] = sum_c rho_ij[(c, a)][(c, b)] from pair (n - 2, n - 1)
        const bool first_of_qi = rem == 0;
        const bool last_pair = p == npairs - 1;
        if (first_of_qi || last_pair) {
          // gather the 16 entries: entry (row, col) sits in lane (g = 2 row, t = col)
          c128 m[4][4];
#pragma unroll
          for (int row = 0; row < 4; ++row)
#pragma unroll
            for (int col = 0; col < 4; ++col)
              m[row][col] = make_double2(__shfl_sync(0xffffffffu, e.x, 8 * row + col), __shfl_sync(0xffffffffu, e.y, 8 * row + col));
          if (lane == 0) {
            if (first_of_qi) {
              c128* o = rdm1 + (s * n + qi) * 4;
#pragma unroll
              for (int a2 = 0; a2 < 2; ++a2)
#pragma unroll
                for (int b2 = 0; b2 < 2; ++b2)
                  o[a2 * 2 + b2] = make_double2(m[2 * a2][2 * b2].x + m[2 * a2 + 1][2 * b2 + 1].x,
                                                a2 == b2 ? 0.0 : m[2 * a2][2 * b2].y + m[2 * a2 + 1][2 * b2 + 1].y);
            }
            if (last_pair) {
              c128* o = rdm1 + (s * n + qj) * 4;
#pragma unroll
              for (int a2 = 0; a2 < 2; ++a2)
#pragma unroll
                for (int b2 = 0; b2 < 2; ++b2)
                  o[a2 * 2 + b2] = make_double2(m[a2][b2].x + m[2 + a2][2 + b2].x,
                                                a2 == b2 ? 0.0 : m[a2][b2].y + m[2 + a2][2 + b2].y);
            }
          }
        }
      }
    }
  }
}

// Reduced density matrix of an arbitrary set of k <= 6 kept qubits (analysis.py:120-166):
// rho[r][c] = sum_env psi[r, env] conj(psi[c, env]); kept_bits[0] = index bit of the first kept qubit (the MSB of
// r), env_bits = the other n - k index bits.  One CTA per state; a thread owns whole (r, c) entries, so every
// entry is one ordered sum (deterministic).
template <class A>
__global__ void qsb_rdm_general_kernel(const A* __restrict__ psi, int n, int k, const int* __restrict__ kept_bits,
                                       const int* __restrict__ env_bits, c128* __restrict__ out) {
  const int64_t dim = (int64_t)1 << n;
  const A* s = psi + blockIdx.x * dim;
  const int D = 1 << k, ne = n - k;
  __shared__ int kb[8], eb[32];
  if (threadIdx.x < k) kb[threadIdx.x] = kept_bits[threadIdx.x];
  if (threadIdx.x < ne) eb[threadIdx.x] = env_bits[threadIdx.x];
  __syncthreads();
  for (int e = threadIdx.x; e < D * D; e += blockDim.x) {
    const int r = e / D, c = e % D;
    if (c < r) continue;                               // Hermitian: the lower triangle is mirrored below
    int64_t ir = 0, ic = 0;
    for (int j = 0; j < k; ++j) {
      if ((r >> (k - 1 - j)) & 1) ir |= (int64_t)1 << kb[j];
      if ((c >> (k - 1 - j)) & 1) ic |= (int64_t)1 << kb[j];
    }
    double re = 0.0, im = 0.0;
    for (int64_t env = 0; env < ((int64_t)1 << ne); ++env) {
      int64_t base = 0;
      for (int j = 0; j < ne; ++j) base |= ((env >> j) & 1) << eb[j];
      const c128 a = qsb_wide(s[base | ir]), b = qsb_wide(s[base | ic]);
      re += a.x * b.x + a.y * b.y;                     // a conj(b)
      im += a.y * b.x - a.x * b.y;
    }
    c128* o = out + (int64_t)blockIdx.x * D * D;
    o[r * D + c] = make_double2(re, r == c ? 0.0 : im);
    if (r != c) o[c * D + r] = make_double2(re, -im);
  }
}

// ---- entropies and mutual information on the device -----------------------------------------------------
// Eigenvalues of a D x D Hermitian matrix (D = 2 or 4, row-major complex) by cyclic complex Jacobi rotations;
// the reference calls np.linalg.eigvalsh (analysis.py:102).  Off-diagonal mass below 1e-32 of the norm ends it.
template <int D>
__device__ void qsb_herm_eigvals(const c128* a_in, double* lam) {
  c128 A[D][D];
#pragma unroll
  for (int r = 0; r < D; ++r)
#pragma unroll
    for (int c = 0; c < D; ++c) A[r][c] = a_in[r * D + c];
  for (int sweep = 0; sweep < 24; ++sweep) {
    double off = 0.0, dia = 0.0;
#pragma unroll
    for (int r = 0; r < D; ++r) {
      dia += A[r][r].x * A[r][r].x;
#pragma unroll
      for (int c = r + 1; c < D; ++c) off += A[r][c].x * A[r][c].x + A[r][c].y * A[r][c].y;
    }
    if (off <= 1e-34 * dia || off == 0.0) break;
#pragma unroll
    for (int p = 0; p < D - 1; ++p)
#pragma unroll
      for (int q = p + 1; q < D; ++q) {
        const double ax = A[p][q].x, ay = A[p][q].y;
        const double mag = sqrt(ax * ax + ay * ay);
        if (mag == 0.0) continue;
        // phase e = a_pq / |a_pq|; real rotation angle from the 2x2 block [[app, mag], [mag, aqq]]
        const double ex = ax / mag, ey = ay / mag;
        const double tau = (A[q][q].x - A[p][p].x) / (2.0 * mag);
        const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
        const double c = 1.0 / sqrt(1.0 + t * t), sn = t * c;
        // columns: A <- A G with G[p][p] = c, G[q][p] = -sn conj(e)... implemented as the unitary
        // U = [[c, sn e], [-sn conj(e), c]] acting on rows/cols (p, q):  A <- U^H A U
#pragma unroll
        for (int k = 0; k < D; ++k) {                      // A <- A U  (columns p, q)
          const c128 akp = A[k][p], akq = A[k][q];
          // new_p = c akp - sn conj(e) akq ; new_q = sn e akp + c akq
          A[k][p] = make_double2(c * akp.x - sn * (ex * akq.x + ey * akq.y), c * akp.y - sn * (ex * akq.y - ey * akq.x));
          A[k][q] = make_double2(sn * (ex * akp.x - ey * akp.y) + c * akq.x, sn * (ex * akp.y + ey * akp.x) + c * akq.y);
        }
#pragma unroll
        for (int k = 0; k < D; ++k) {                      // A <- U^H A  (rows p, q)
          const c128 apk = A[p][k], aqk = A[q][k];
          // new_p = c apk - sn e aqk ; new_q = sn conj(e) apk + c aqk
          A[p][k] = make_double2(c * apk.x - sn * (ex * aqk.x - ey * aqk.y), c * apk.y - sn * (ex * aqk.y + ey * aqk.x));
          A[q][k] = make_double2(sn * (ex * apk.x + ey * apk.y) + c * aqk.x, sn * (ex * apk.y - ey * apk.x) + c * aqk.y);
        }
      }
  }
#pragma unroll
  for (int r = 0; r < D; ++r) lam[r] = A[r][r].x;
}

// -sum lambda log2 lambda over eigenvalues > 1e-15 (analysis.py:102-104)
template <int D>
__device__ double qsb_entropy_bits(const c128* rho) {
  double lam[D], s = 0.0;
  qsb_herm_eigvals<D>(rho, lam);
#pragma unroll
  for (int r = 0; r < D; ++r) if (lam[r] > 1e-15) s -= lam[r] * log2(lam[r]);
  return s;
}

// I(i:j) = max(0, S_i + S_j - S_ij) for all pairs i < j (analysis.py:183-191, :315-333); one thread per (state, pair)
__global__ void qsb_mi_kernel(const c128* __restrict__ rdm1, const c128* __restrict__ rdm2, int n, int npairs,
                              int64_t total, double* __restrict__ s1_out, double* __restrict__ mi) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int64_t t = e / npairs;
  const int pair = (int)(e % npairs);
  int qi = 0, rem = pair;
  while (rem >= n - 1 - qi) { rem -= n - 1 - qi; ++qi; }
  const int qj = qi + 1 + rem;
  const double si = qsb_entropy_bits<2>(rdm1 + (t * n + qi) * 4);
  const double sj = qsb_entropy_bits<2>(rdm1 + (t * n + qj) * 4);
  const double sij = qsb_entropy_bits<4>(rdm2 + (t * npairs + pair) * 16);
  const double v = si + sj - sij;
  mi[e] = v > 0.0 ? v : 0.0;
  if (s1_out) {                                   // single-qubit entropies, written once per qubit
    if (rem == 0 && qi < n - 1) s1_out[t * n + qi] = si;
    if (pair == npairs - 1) s1_out[t * n + qj] = sj;
  }
}

// rho[i][j] += scale * sum_t psi_t[i] conj(psi_t[j])   (simulator.py:195-198)
// The one dense contraction of the path: with A = Re Psi, B = Im Psi (dim x N),
//   Re rho = A A^T + B B^T,   Im rho = B A^T - A B^T
// i.e. real GEMMs with K = N trajectories, run on the FP64 tensor-core path (DMMA, mma.sync m8n8k4).
// One CTA = one 64x64 tile of the upper triangle (rho is Hermitian; the mirrored tile is written from the
// same accumulators); 8 warps, each a 32x16 patch = 4x2 m8n8 tiles x (Re, Im); 16 trajectories per
// shared-memory stage.
#define QSB_RHO_TILE 64
#define QSB_RHO_KC 16
#define QSB_RHO_PAD 2          // doubles of row padding: the 4 k-rows a fragment load touches hit different banks


template <class A>
__global__ void __launch_bounds__(256) qsb_rho_kernel(const A* __restrict__ psi, int64_t dim, int64_t count,
                                                       double scale, c128* __restrict__ rho, int tiles_per_side) {
  // re / im planes of the two 64-wide panels, [k][x]
  __shared__ double sa_re[QSB_RHO_KC][QSB_RHO_TILE + QSB_RHO_PAD], sa_im[QSB_RHO_KC][QSB_RHO_TILE + QSB_RHO_PAD];
  __shared__ double sb_re[QSB_RHO_KC][QSB_RHO_TILE + QSB_RHO_PAD], sb_im[QSB_RHO_KC][QSB_RHO_TILE + QSB_RHO_PAD];
  // linear tile id -> (ti <= tj) in the upper triangle
  int ti = 0, rem = blockIdx.x;
  while (rem >= tiles_per_side - ti) { rem -= tiles_per_side - ti; ++ti; }
  const int tj = ti + rem;
  const int64_t i0 = (int64_t)ti * QSB_RHO_TILE, j0 = (int64_t)tj * QSB_RHO_TILE;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wr = (warp >> 2) * 32, wc = (warp & 3) * 16;         // warp patch origin inside the tile
  const int grp = lane >> 2, tig = lane & 3;                     // fragment coordinates
  double re[4][2][2], im[4][2][2];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) { re[a][b][0] = re[a][b][1] = im[a][b][0] = im[a][b][1] = 0.0; }
  for (int64_t t0 = 0; t0 < count; t0 += QSB_RHO_KC) {
    for (int e = threadIdx.x; e < QSB_RHO_KC * QSB_RHO_TILE; e += 256) {
      const int k = e / QSB_RHO_TILE, x = e % QSB_RHO_TILE;
      c128 za = make_double2(0.0, 0.0), zb = za;
      if (t0 + k < count) {
        if (i0 + x < dim) za = qsb_wide(psi[(t0 + k) * dim + i0 + x]);
        if (j0 + x < dim) zb = qsb_wide(psi[(t0 + k) * dim + j0 + x]);
      }
      sa_re[k][x] = za.x; sa_im[k][x] = za.y;
      sb_re[k][x] = zb.x; sb_im[k][x] = zb.y;
    }
    __syncthreads();
#pragma unroll
    for (int k0 = 0; k0 < QSB_RHO_KC; k0 += 4) {
      double ar[4], ai[4], br[2], bi[2];
#pragma unroll
      for (int a = 0; a < 4; ++a) { ar[a] = sa_re[k0 + tig][wr + a * 8 + grp]; ai[a] = sa_im[k0 + tig][wr + a * 8 + grp]; }
#pragma unroll
      for (int b = 0; b < 2; ++b) { br[b] = sb_re[k0 + tig][wc + b * 8 + grp]; bi[b] = sb_im[k0 + tig][wc + b * 8 + grp]; }
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          // psi_i conj(psi_j) = (ar br + ai bi) + i (ai br - ar bi)
          qsb_dmma(re[a][b][0], re[a][b][1], ar[a], br[b]);
          qsb_dmma(re[a][b][0], re[a][b][1], ai[a], bi[b]);
          qsb_dmma(im[a][b][0], im[a][b][1], ai[a], br[b]);
          qsb_dmma(im[a][b][0], im[a][b][1], -ar[a], bi[b]);
        }
    }
    __syncthreads();
  }
  // accumulator (row = grp, cols 2 tig, 2 tig + 1) -> rho and, off the diagonal tiles, its mirror image
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int64_t i = i0 + wr + a * 8 + grp, j = j0 + wc + b * 8 + 2 * tig + c;
        // every pair (i, j), i <= j, is written once together with its mirror image, so rho is exactly
        // Hermitian (as the reference's sum of outer products is); a diagonal tile skips its lower half
        if (i < dim && j < dim && i <= j) {
          const double vim = (i == j) ? 0.0 : im[a][b][c];
          c128 o = rho[i * dim + j];
          o.x += scale * re[a][b][c];
          o.y += scale * vim;
          rho[i * dim + j] = o;
          if (i != j) {
            c128 m = rho[j * dim + i];
            m.x += scale * re[a][b][c];
            m.y -= scale * vim;
            rho[j * dim + i] = m;
          }
        }
      }
}

// one axis of ReadoutError.apply_to_distribution (noise.py:163-169): out[m] = C[m][0] p0 + C[m][1] p1
__global__ void qsb_readout_axis_kernel(double* __restrict__ p, int64_t dim, int64_t count, int b, double c00, double c01,
                                        double c10, double c11) {
  const int64_t half = dim / 2, total = half * count;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    int64_t t = e / half, g = e % half;
    int64_t i0 = ((g >> b) << (b + 1)) | (g & (((int64_t)1 << b) - 1));
    double* q = p + t * dim;
    double p0 = q[i0], p1 = q[i0 | ((int64_t)1 << b)];
    q[i0] = c00 * p0 + c01 * p1;
    q[i0 | ((int64_t)1 << b)] = c10 * p0 + c11 * p1;
  }
}

// ReadoutError.apply_to_distribution (noise.py:141-175) for n <= 14 in ONE launch: a CTA keeps one distribution
// (8 * 2^n bytes <= 128 KiB) in shared memory, applies the 2x2 confusion butterfly of every axis in the reference's
// order (axis q = index bit n-1-q, q ascending), sums, renormalises (total > 1e-15) and writes it back -- one read
// and one write of HBM instead of n + 1 of each.
__global__ void __launch_bounds__(512) qsb_readout_fused_kernel(double* __restrict__ p, int n, double c00, double c01,
                                                                 double c10, double c11) {
  extern __shared__ __align__(16) unsigned char qsb_ro_smem[];
  double* sp = reinterpret_cast<double*>(qsb_ro_smem);
  __shared__ double scratch[32];
  const int dim = 1 << n, half = dim >> 1;
  double* q = p + (int64_t)blockIdx.x * dim;
  for (int i = threadIdx.x; i < dim; i += blockDim.x) sp[i] = q[i];
  __syncthreads();
  for (int b = n - 1; b >= 0; --b) {
    for (int g = threadIdx.x; g < half; g += blockDim.x) {
      const int i0 = ((g >> b) << (b + 1)) | (g & ((1 << b) - 1));
      const double p0 = sp[i0], p1 = sp[i0 | (1 << b)];
      sp[i0] = c00 * p0 + c01 * p1;
      sp[i0 | (1 << b)] = c10 * p0 + c11 * p1;
    }
    __syncthreads();
  }
  double v[1] = {0.0};
  for (int i = threadIdx.x; i < dim; i += blockDim.x) v[0] += sp[i];
  qsb_block_sum<1>(v, scratch);
  const bool norm = v[0] > 1e-15;
  for (int i = threadIdx.x; i < dim; i += blockDim.x) q[i] = norm ? sp[i] / v[0] : sp[i];
}

// divide each distribution by its total if total > 1e-15 (noise.py:172-174); one CTA per distribution
__global__ void qsb_normalize_dist_kernel(double* __restrict__ p, int64_t dim) {
  __shared__ double scratch[32];
  double* q = p + blockIdx.x * dim;
  double v[1] = {0.0};
  for (int64_t i = threadIdx.x; i < dim; i += blockDim.x) v[0] += q[i];
  qsb_block_sum<1>(v, scratch);
  if (v[0] > 1e-15)
    for (int64_t i = threadIdx.x; i < dim; i += blockDim.x) q[i] = q[i] / v[0];
}


// Generic dense k-qubit operator on states at rest in HBM (state_vector.py:41-74 for k > 3, where the tile executor
// has no sweep): out[perm(x)] = sum_c M[r(x)][c] * in[x with the target bits set to c].  tb[j] = index bit of
// targets[j] (targets[0] is the most significant bit of the matrix index); perm[b] = destination bit of index bit b
// (the reference's axis scramble, or the identity).  One thread per output amplitude; 2^k loads each come from L2.
template <class A>
__global__ void qsb_dense_kernel(const A* __restrict__ in, A* __restrict__ out, int n, int k, int64_t count,
                                 const c128* __restrict__ M, const int* __restrict__ tb, const int* __restrict__ perm) {
  __shared__ int s_tb[16], s_perm[32];
  if (threadIdx.x < k) s_tb[threadIdx.x] = tb[threadIdx.x];
  if (threadIdx.x < n) s_perm[threadIdx.x] = perm[threadIdx.x];
  __syncthreads();
  const int64_t dim = (int64_t)1 << n, total = count * dim;
  const int D = 1 << k;
  for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < total; g += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = g >> n;
    const uint32_t x = (uint32_t)(g & (dim - 1));
    uint32_t base = x, r = 0;
    for (int j = 0; j < k; ++j) {
      const int b = s_tb[j];
      r = (r << 1) | ((x >> b) & 1u);
      base &= ~(1u << b);
    }
    const A* src = in + t * dim;
    const c128* row = M + (size_t)r * D;
    double ax = 0.0, ay = 0.0;
    for (int c = 0; c < D; ++c) {
      uint32_t idx = base;
      for (int j = 0; j < k; ++j) idx |= ((uint32_t)(c >> (k - 1 - j)) & 1u) << s_tb[j];
      const c128 v = qsb_wide(src[idx]);
      const c128 mv = row[c];
      ax = fma(mv.x, v.x, fma(-mv.y, v.y, ax));
      ay = fma(mv.x, v.y, fma(mv.y, v.x, ay));
    }
    uint32_t y = 0;
    for (int b = 0; b < n; ++b) y |= ((x >> b) & 1u) << s_perm[b];
    c128 res; res.x = ax; res.y = ay;
    out[t * dim + y] = qsb_cvt<A>(res);
  }
}
