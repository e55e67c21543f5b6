// qsb_stream.cuh -- streamed passes over a statevector that lives in HBM (n > 16 qubits, BASELINE config 5):
// a TMA-fed, triple-buffered tile pipeline.  sm_100a only.
//
// Reference semantics: StateVector.apply_gate (state_vector.py:41-74) applied gate by gate to one 2^n state whose
// constructor cap was bypassed (state_vector.py:156-158).  Here a PASS = one launch that reads every amplitude once
// and writes it once; between the read and the write a tile of 2^m amplitudes sits in shared memory and takes all
// the BLOCK SWEEPS of the pass.  A block sweep is one shared-memory round trip of the tile in which every worker holds
// the 16 amplitudes of four index bits in registers and applies a whole list of ops to them: the host compiler
// (qsb/stream.py) has turned every 1-qubit gate into a 2x2 that rides in front of the next multi-qubit gate on its
// qubit, and has grouped consecutive gates whose qubits fit four bits into one block (124 -> 73 round trips for the
// 26-qubit layered circuit).  ncu on the one-gate-per-sweep version showed the shared-memory pipe at 79 % of its peak
// with HBM at 38 %: the round trips, not the arithmetic (FP64 pipe 10 %), were the bound.
//
// Tile geometry.  Slot bits 0..l-1 are the low index bits (one contiguous 16 * 2^l-byte row in HBM); slots l..l+e-1
// are resident bits that ride in the TMA box as extra dimensions of extent 2; slots l+e..m-1 number the TMA ops of a
// tile; slots m..n-1 number the tiles.  One cp.async.bulk.tensor.5d moves 2^(l+e) amplitudes (>= 2 KiB keeps the TMA
// unit off its ~50-cycle per-op floor: measured 1.4 / 2.7 / 4.9 / 5.9 TB/s at 256 B / 512 B / 1 KiB / 2 KiB per op,
// tools/micro/tma_probe.cu) and lands them in the executor's XOR-swizzled slot order: CU_TENSOR_MAP_SWIZZLE_128B is
// slot = i ^ ((i >> 3) & 7) on 16-byte elements.
//
// Pipeline (one CTA per SM, persistent over its tiles t = blockIdx.x + j * gridDim.x):
//   * producer warp: issues the loads of tiles j, j+1, j+2 into the three tile buffers, waits for "tile j swept",
//     issues its store, waits until the store has READ shared memory, re-loads that buffer with tile j+3;
//   * two worker groups of 256 threads, group g sweeps tiles j = g, g+2, ...: while one group sits in a barrier or
//     waits for LDS data the other one keeps the shared-memory pipe busy; between sweeps a group meets on its own
//     named barrier, after the last sweep it fences its writes for the async proxy and signals the producer.
#pragma once

#include <cuda.h>

#include "qsb_exec.cuh"

#define QSB_ST_BUFS 3
#define QSB_ST_GROUPS 2
#define QSB_ST_MAX_SWEEPS 16           // block sweeps per pass
#define QSB_ST_BLOCK_OPS 12            // ops per block of the C ABI (qsb_stream_block)
#define QSB_ST_MAX_TILE_BITS 12          // 3 x 64 KiB tiles + the sweep list fit 227 KiB

extern __shared__ __align__(1024) unsigned char qsb_stream_smem[];

#define QSB_ST_MAX_PEERS 8
// tensor maps of one launch: the shard(s) tiles are loaded from (one per peer when the qubit exchange is folded into
// the pass, else only in[0]) and the shard they are stored to
struct qsb_stream_maps {
  CUtensorMap in[QSB_ST_MAX_PEERS];
  CUtensorMap out[QSB_ST_MAX_PEERS];
};

struct qsb_blk;
struct qsb_stream_kargs {
  int32_t n, m, l, e;
  int32_t peer_shift, pad0;              // source element x comes from in[x >> peer_shift] (peer_shift = n: one source)
  uint32_t peer_or, pad1;                // ... at offset (x & (2^peer_shift - 1)) | peer_or
  int32_t out_shift, pad2;               // destination element x goes to out[x >> out_shift] (out_shift = 32: one destination)
  uint32_t out_or;                       // ... at offset (x & (2^out_shift - 1)) | out_or
  uint32_t tile_xor;                     // XORed into the tile number: with peers, rank r walks its tiles in an order
                                         // in which the peer-selecting tile bits are flipped by r, so that at any
                                         // moment the ranks talk to DIFFERENT peers (no incast on one GPU's links)
  int32_t n_sweeps, n_ops;               // sweeps per tile; TMA ops per tile = 2^(m - l - e)
  int32_t op_pos[16];                    // slot l+e+j -> bit position in the (local) amplitude index, load side
  int32_t tile_pos[32];                  // slot m+j   -> bit position, load side
  int32_t op_pos_out[16];                // the same for the store side (identical for an in-place pass; a reorder
  int32_t tile_pos_out[32];              //   pass stores every slot to another position, out of place)
  const qsb_blk* sweeps;                 // device memory, n_sweeps block sweeps
};

__device__ __forceinline__ uint32_t qsb_st_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void qsb_st_mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void qsb_st_mbar_expect(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void qsb_st_mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void qsb_st_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void qsb_st_tma_load(uint32_t dst, const CUtensorMap* map, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %2, %2, %2}], [%4];"
               ::"r"(dst), "l"(map), "r"(0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void qsb_st_tma_store(const CUtensorMap* map, int c1, uint32_t src) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%1, %2, %1, %1, %1}], [%3];"
               ::"l"(map), "r"(0), "r"(c1), "r"(src) : "memory");
}

// what qsb_sweep needs from its environment: the tile, the worker's place in its group, no profiling
struct StreamEnv {
  typedef c128 amp;
  int wid, W, wbits;
  int cur;                               // tile buffer this group works on
  int tile_bytes;
  __device__ __forceinline__ c128* tile() { return reinterpret_cast<c128*>(qsb_stream_smem + (size_t)cur * tile_bytes); }
  __device__ __forceinline__ bool prof_on() { return false; }
  __device__ __forceinline__ unsigned long long clock() { return 0; }
  __device__ __forceinline__ void prof_add(int, unsigned long long) {}
};

// ---- block sweeps ------------------------------------------------------------------------------------------------
// A block in the form the workers run it (built by qsb_stream_create from the op list of include/qsb.h's
// qsb_stream_block): straight-line code, no per-op dispatch --
//   1. load the 16 amplitudes of the block's four index bits (local bit i = tile slot bit b[i]);
//   2. a 2x2 on each local bit that has one (class cls[t]);
//   3. at most one dense 4x4 on local bits (3, 2) or 8x8 on (3, 2, 1);
//   4. the block's permutation-type gates (CX, SWAP, Toffoli, Fredkin) and sign gates (CZ) are never executed on
//      registers: composed on the host they say "register r goes to the place of pi(r), negated if neg_mask has bit r";
//   5. store.
// Offsets are byte offsets inside the tile, already swizzled: slot() is linear over XOR, so the address of amplitude
// (base | off[r]) is (slot(base) * 16) ^ ld_off[r].
struct alignas(16) qsb_blk {
  uint32_t ld_off[16];
  uint32_t st_off[16];
  c128 U[4][4];              // 2x2 of local bit t, row-major
  int32_t cls[4];            // QSB_CLS_* of U[t] (NONE: nothing to do)
  uint64_t pos;              // group order (qsb_group_order) of the 2^(m-4) register blocks of a tile
  int32_t hmask;
  uint32_t neg_mask;
  const c128* mat;           // dense matrix in device memory
  int32_t dense;             // 0 | 2 | 3
  int32_t b[4];
  int32_t variant;           // dense == 0: mask of the local bits with a 2x2 (picks qsb_block_sweep_s<MASK>)
  int32_t tabl[32];          // first register block of worker w: tabl[w & 31] | tabw[w >> 5] (qsb_deposit of w onto `pos`, per worker count)
  int32_t tabw[8];
};

// 2x2 on local bit T of the register block
template <int T>
__device__ __forceinline__ void qsb_blk_mat1(c128 (&a)[16], const c128* U, int cls) {
  if (cls >= QSB_CLS_DIAG) {
    const c128 u0 = U[0], u1 = U[1], u2 = U[2], u3 = U[3];
    if (cls == QSB_CLS_DIAG) {
#pragma unroll
      for (int r = 0; r < 16; ++r) a[r] = qsb_mul((r >> T) & 1 ? u3 : u0, a[r]);
    } else {
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        if ((r >> T) & 1) continue;
        const c128 lo = a[r], hi = a[r | (1 << T)];
        a[r] = qsb_fma(u1, hi, qsb_mul(u0, lo));
        a[r | (1 << T)] = qsb_fma(u3, hi, qsb_mul(u2, lo));
      }
    }
  } else if (cls == QSB_CLS_RDIAG) {
    const double sc = U[3].x;
#pragma unroll
    for (int r = 0; r < 16; ++r) if ((r >> T) & 1) { a[r].x *= sc; a[r].y *= sc; }
  }
}
// dense 4x4 on local bits (3, 2): matrix index = (bit 3, bit 2); the four values of bits (1, 0) are four independent columns
__device__ __forceinline__ void qsb_blk_dense2(c128 (&a)[16], const c128* __restrict__ M) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    c128 v[4], o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = a[(j << 2) | g];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      o[i] = qsb_mul(M[i * 4], v[0]);
#pragma unroll
      for (int j = 1; j < 4; ++j) o[i] = qsb_fma(M[i * 4 + j], v[j], o[i]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) a[(i << 2) | g] = o[i];
  }
}
// dense 8x8 on local bits (3, 2, 1)
__device__ __forceinline__ void qsb_blk_dense3(c128 (&a)[16], const c128* __restrict__ M) {
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    c128 o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      o[i] = qsb_mul(M[i * 8], a[g]);
#pragma unroll
      for (int j = 1; j < 8; ++j) o[i] = qsb_fma(M[i * 8 + j], a[(j << 1) | g], o[i]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) a[(i << 1) | g] = o[i];
  }
}

// one shared-memory round trip of the tile: every worker owns whole 16-amplitude register blocks
__device__ __forceinline__ void qsb_block_sweep(StreamEnv& env, int m, const qsb_blk* d) {
  unsigned char* tile = reinterpret_cast<unsigned char*>(env.tile());
  const int free_bits = m - 4;
  const int lo_base = qsb_deposit(env.wid, d->pos, env.wbits < free_bits ? env.wbits : free_bits);
  const int hmask = d->hmask;
  const int c0 = d->cls[0], c1 = d->cls[1], c2 = d->cls[2], c3 = d->cls[3];
  const int dense = d->dense;
  const uint32_t neg = d->neg_mask;
  // the offset tables are re-read from shared memory in every step (broadcast LDS.128): holding them in registers for
  // the whole sweep was measured 27 % slower (15.5 ms against 12.2 ms on the 26-qubit circuit; registers, not the two
  // extra shared-memory round trips, are what this loop is short of)
  const uint4* ldo = reinterpret_cast<const uint4*>(d->ld_off);
  const uint4* sto = reinterpret_cast<const uint4*>(d->st_off);
  int hi = 0;
  for (int g0 = env.wid; g0 < (1 << free_bits); g0 += env.W) {
    const int bs = lo_base | hi;
    hi = ((hi | ~hmask) + 1) & hmask;
    const uint32_t sbs = (uint32_t)qsb_slot(bs) << 4;
    c128 a[16];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint4 o = ldo[q];
      a[4 * q + 0] = *reinterpret_cast<const c128*>(tile + (sbs ^ o.x));
      a[4 * q + 1] = *reinterpret_cast<const c128*>(tile + (sbs ^ o.y));
      a[4 * q + 2] = *reinterpret_cast<const c128*>(tile + (sbs ^ o.z));
      a[4 * q + 3] = *reinterpret_cast<const c128*>(tile + (sbs ^ o.w));
    }
    if (c0 != QSB_CLS_NONE) qsb_blk_mat1<0>(a, d->U[0], c0);
    if (c1 != QSB_CLS_NONE) qsb_blk_mat1<1>(a, d->U[1], c1);
    if (c2 != QSB_CLS_NONE) qsb_blk_mat1<2>(a, d->U[2], c2);
    if (c3 != QSB_CLS_NONE) qsb_blk_mat1<3>(a, d->U[3], c3);
    if (dense == 2) qsb_blk_dense2(a, d->mat);
    else if (dense == 3) qsb_blk_dense3(a, d->mat);
    if (neg) {
#pragma unroll
      for (int r = 0; r < 16; ++r) if ((neg >> r) & 1u) a[r] = qsb_neg(a[r]);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint4 o = sto[q];
      *reinterpret_cast<c128*>(tile + (sbs ^ o.x)) = a[4 * q + 0];
      *reinterpret_cast<c128*>(tile + (sbs ^ o.y)) = a[4 * q + 1];
      *reinterpret_cast<c128*>(tile + (sbs ^ o.z)) = a[4 * q + 2];
      *reinterpret_cast<c128*>(tile + (sbs ^ o.w)) = a[4 * q + 3];
    }
  }
}

// The same round trip for blocks without a dense 4x4 / 8x8, with the run-time choices taken out of the loop: MASK (compile
// time) says which local bits carry a 2x2 (whatever its class: a real or complex diagonal is applied as a full matrix),
// the signs of CZ are XORed into the sign bits.  With the class branches inside the loop ptxas cannot move a step's
// arithmetic under its loads and stores (the same finding as in the resident executor, tools/micro/sweep_real.cu).
template <int MASK>
__device__ __forceinline__ void qsb_block_sweep_s(StreamEnv& env, int m, const qsb_blk* d) {
  unsigned char* tile = reinterpret_cast<unsigned char*>(env.tile());
  const int free_bits = m - 4;
  const int lo_base = d->tabl[env.wid & 31] | d->tabw[env.wid >> 5];
  const int hmask = d->hmask;
  const uint32_t neg = d->neg_mask;
  const uint4* ldo = reinterpret_cast<const uint4*>(d->ld_off);
  const uint4* sto = reinterpret_cast<const uint4*>(d->st_off);
  int hi = 0;
  for (int g0 = env.wid; g0 < (1 << free_bits); g0 += env.W) {
    const int bs = lo_base | hi;
    hi = ((hi | ~hmask) + 1) & hmask;
    const uint32_t sbs = (uint32_t)qsb_slot(bs) << 4;
    c128 a[16];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint4 o = ldo[q];
      a[4 * q + 0] = *reinterpret_cast<const c128*>(tile + (sbs ^ o.x));
      a[4 * q + 1] = *reinterpret_cast<const c128*>(tile + (sbs ^ o.y));
      a[4 * q + 2] = *reinterpret_cast<const c128*>(tile + (sbs ^ o.z));
      a[4 * q + 3] = *reinterpret_cast<const c128*>(tile + (sbs ^ o.w));
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      if (!((MASK >> t) & 1)) continue;
      const c128 u0 = d->U[t][0], u1 = d->U[t][1], u2 = d->U[t][2], u3 = d->U[t][3];
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        if ((r >> t) & 1) continue;
        const c128 lo = a[r], up = a[r | (1 << t)];
        a[r] = qsb_fma(u1, up, qsb_mul(u0, lo));
        a[r | (1 << t)] = qsb_fma(u3, up, qsb_mul(u2, lo));
      }
    }
#pragma unroll
    for (int r = 0; r < 16; ++r) {                     // sign gates: flip the sign bits (integer pipe, no branch)
      const long long flip = (long long)((neg >> r) & 1u) << 63;
      a[r].x = __longlong_as_double(__double_as_longlong(a[r].x) ^ flip);
      a[r].y = __longlong_as_double(__double_as_longlong(a[r].y) ^ flip);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint4 o = sto[q];
      *reinterpret_cast<c128*>(tile + (sbs ^ o.x)) = a[4 * q + 0];
      *reinterpret_cast<c128*>(tile + (sbs ^ o.y)) = a[4 * q + 1];
      *reinterpret_cast<c128*>(tile + (sbs ^ o.z)) = a[4 * q + 2];
      *reinterpret_cast<c128*>(tile + (sbs ^ o.w)) = a[4 * q + 3];
    }
  }
}
__device__ __forceinline__ void qsb_block_dispatch(StreamEnv& env, int m, const qsb_blk* d) {
  if (d->dense != 0) { qsb_block_sweep(env, m, d); return; }
  switch (d->variant) {
    case 0: qsb_block_sweep_s<0>(env, m, d); break;
    case 1: qsb_block_sweep_s<1>(env, m, d); break;
    case 2: qsb_block_sweep_s<2>(env, m, d); break;
    case 3: qsb_block_sweep_s<3>(env, m, d); break;
    case 4: qsb_block_sweep_s<4>(env, m, d); break;
    case 5: qsb_block_sweep_s<5>(env, m, d); break;
    case 6: qsb_block_sweep_s<6>(env, m, d); break;
    case 7: qsb_block_sweep_s<7>(env, m, d); break;
    case 8: qsb_block_sweep_s<8>(env, m, d); break;
    case 9: qsb_block_sweep_s<9>(env, m, d); break;
    case 10: qsb_block_sweep_s<10>(env, m, d); break;
    case 11: qsb_block_sweep_s<11>(env, m, d); break;
    case 12: qsb_block_sweep_s<12>(env, m, d); break;
    case 13: qsb_block_sweep_s<13>(env, m, d); break;
    case 14: qsb_block_sweep_s<14>(env, m, d); break;
    default: qsb_block_sweep_s<15>(env, m, d); break;
  }
}

// GT worker threads per group (two groups) + one TMA warp.  (A build without the dedicated TMA warp -- warp 0 of each group
// issuing its group's stores and refills between sweeps -- was measured 18 % slower and is gone.)
template <int GT>
__global__ void __launch_bounds__(QSB_ST_GROUPS * GT + 32, 1)
qsb_stream_kernel(const __grid_constant__ qsb_stream_maps maps, const __grid_constant__ qsb_stream_kargs a) {
  const int tile_bytes = 16 << a.m;
  const int op_bytes = 16 << (a.l + a.e);
  unsigned char* base = qsb_stream_smem;
  qsb_blk* descs = reinterpret_cast<qsb_blk*>(base + (size_t)QSB_ST_BUFS * tile_bytes);
  uint32_t* row_off = reinterpret_cast<uint32_t*>(descs + QSB_ST_MAX_SWEEPS);        // [32] load side, [32] store side
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(row_off + 64);    // full[3], done[3]
  const int tid = threadIdx.x;
  // ---- prologue: the sweep list, the op -> row offset table, the barriers
  {
    const int words = a.n_sweeps * (int)(sizeof(qsb_blk) / 16);
    const uint4* src = reinterpret_cast<const uint4*>(a.sweeps);
    uint4* dst = reinterpret_cast<uint4*>(descs);
    for (int i = tid; i < words; i += blockDim.x) dst[i] = src[i];
    if (tid < a.n_ops) {
      uint32_t off = 0, off_out = 0;
      for (int j = 0; j < a.m - a.l - a.e; ++j) {
        off |= ((uint32_t)(tid >> j) & 1u) << a.op_pos[j];
        off_out |= ((uint32_t)(tid >> j) & 1u) << a.op_pos_out[j];
      }
      row_off[tid] = off;
      row_off[32 + tid] = off_out;
    }
    if (tid == 0) {
      for (int b = 0; b < 2 * QSB_ST_BUFS; ++b) qsb_st_mbar_init(qsb_st_smem_u32(&bars[b]), 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
  }
  __syncthreads();
  const int64_t ntiles = (int64_t)1 << (a.n - a.m);
  const int64_t mine = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;   // tiles of this CTA
  auto tile_base = [&](int64_t j, const int32_t* pos) {
    const uint32_t t = (uint32_t)(blockIdx.x + j * gridDim.x) ^ a.tile_xor;
    uint32_t off = 0;
    for (int q = 0; q < a.n - a.m; ++q) off |= ((t >> q) & 1u) << pos[q];
    return off;
  };

  if (tid >= QSB_ST_GROUPS * GT) {
    // ================= producer warp: TMA loads and stores, buffer turnover =================
    const int lane = tid & 31;
    // gather tile j into buffer j % 3 (completes on full[j % 3])
    auto issue_load = [&](int64_t j) {
      const int b = (int)(j % QSB_ST_BUFS);
      const uint32_t full = qsb_st_smem_u32(&bars[b]);
      const uint32_t tb = tile_base(j, a.tile_pos);
      if (lane == 0) qsb_st_mbar_expect(full, (uint32_t)tile_bytes);
      __syncwarp();
      const uint32_t pmask = a.peer_shift >= 32 ? 0xffffffffu : ((1u << a.peer_shift) - 1u);
      for (int r = lane; r < a.n_ops; r += 32) {
        const uint32_t x = tb | row_off[r];
        const uint32_t src = a.peer_shift >= 32 ? 0u : (x >> a.peer_shift);
        qsb_st_tma_load(qsb_st_smem_u32(base + (size_t)b * tile_bytes + (size_t)r * op_bytes), &maps.in[src],
                        (int)(((x & pmask) | a.peer_or) >> 3), full);
      }
    };
    for (int64_t j = 0; j < mine && j < QSB_ST_BUFS; ++j) issue_load(j);
    for (int64_t j = 0; j < mine; ++j) {
      const int b = (int)(j % QSB_ST_BUFS);
      qsb_st_mbar_wait(qsb_st_smem_u32(&bars[QSB_ST_BUFS + b]), (uint32_t)((j / QSB_ST_BUFS) & 1));   // tile j is swept
      // scatter buffer b to the (store-side) positions of tile j: one bulk group per lane
      const uint32_t tb = tile_base(j, a.tile_pos_out);
      const uint32_t omask = a.out_shift >= 32 ? 0xffffffffu : ((1u << a.out_shift) - 1u);
      for (int r = lane; r < a.n_ops; r += 32) {
        // with out_shift < 32 the store IS the qubit exchange: the top index bits of the destination pick the peer
        // whose shard receives the box (a posted write over NVLink), my rank fills the bits it vacates
        const uint32_t x = tb | row_off[32 + r];
        const uint32_t dst = a.out_shift >= 32 ? 0u : (x >> a.out_shift);
        qsb_st_tma_store(&maps.out[dst], (int)(((x & omask) | a.out_or) >> 3),
                         qsb_st_smem_u32(base + (size_t)b * tile_bytes + (size_t)r * op_bytes));
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      if (j + QSB_ST_BUFS < mine) {
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");    // the store has read the buffer: refill it
        __syncwarp();
        issue_load(j + QSB_ST_BUFS);
      }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");             // every store is complete before the CTA exits
    return;
  }

  // ================= worker groups =================
  const int g = tid / GT;
  StreamEnv env;
  env.wid = tid % GT;
  env.W = GT;
  env.wbits = GT == 256 ? 8 : 7;
  env.tile_bytes = tile_bytes;
  const int bar_id = 1 + g;
  for (int64_t j = g; j < mine; j += QSB_ST_GROUPS) {
    const int b = (int)(j % QSB_ST_BUFS);
    env.cur = b;
    // A parity wait only tells "the phase before the current one is complete".  Tile j - 3 used this buffer and was
    // swept by the OTHER group: unless that is known to be over, full[b] may still be in the phase of tile j - 3 (its
    // load not even landed) and the wait for tile j's phase would fall straight through.  With no sweeps in a pass
    // (a reorder) a group does get three tiles ahead of a slow load, so: first "tile j - 3 is swept", then "tile j is in".
    if (j >= QSB_ST_BUFS)
      qsb_st_mbar_wait(qsb_st_smem_u32(&bars[QSB_ST_BUFS + b]), (uint32_t)(((j - QSB_ST_BUFS) / QSB_ST_BUFS) & 1));
    qsb_st_mbar_wait(qsb_st_smem_u32(&bars[b]), (uint32_t)((j / QSB_ST_BUFS) & 1));                 // tile j has landed
    for (int s = 0; s < a.n_sweeps; ++s) {
      qsb_block_dispatch(env, a.m, &descs[s]);
      if (s + 1 < a.n_sweeps) asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(GT) : "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          // generic-proxy writes -> visible to the TMA store
    asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(GT) : "memory");
    if (env.wid == 0) qsb_st_mbar_arrive(qsb_st_smem_u32(&bars[QSB_ST_BUFS + b]));
  }
}

// dynamic shared memory of the kernel for m tile bits
static inline size_t qsb_stream_smem_bytes(int m) {
  return (size_t)QSB_ST_BUFS * ((size_t)16 << m) + sizeof(qsb_blk) * QSB_ST_MAX_SWEEPS + 64 * sizeof(uint32_t) +
         2 * QSB_ST_BUFS * sizeof(unsigned long long);
}
