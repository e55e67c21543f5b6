// qsb_exec.cuh -- the resident-tile executor (trajectories for n <= 16, streamed tiles above).
//
// One CTA cluster (1, 2, 4 or 8 CTAs) keeps one 2^n statevector in shared memory
// (2^m amplitudes per CTA, n - m cluster-rank bits) for a whole trajectory: gates,
// Kraus steps, snapshots and the final store, so a 16-qubit trajectory touches HBM
// only for its op list, its uniforms and its final state.  For n > 16 the same code
// runs one CTA per 2^m-amplitude tile of a state that lives in HBM ("streaming"
// mode: the high index bits select the tile instead of the cluster rank).
//
// Warp specialisation.  Each CTA has W worker threads and one control warp:
//   * the CONTROL warp walks the op list.  Every 1-qubit operation (gate, Pauli
//     branch, amplitude-damping K0/K1, generic Kraus operator) is only multiplied
//     into a pending 2x2 matrix of its slot bit -- no amplitude is touched.  When a
//     multi-qubit gate, a state-dependent Kraus draw, a cluster remap or a store needs
//     the data, the control warp publishes a DESCRIPTOR (what to sweep, the pending
//     matrices to apply on the way, their structure class) into a small ring in
//     shared memory and runs ahead;
//   * the WORKERS consume descriptors: one pass over the tile applies the pending
//     matrices of the touched bits plus the gate.
// Every CTA of a cluster runs an identical control warp on the same ops and the same
// uniforms, so all copies take the same decisions without communication.
//
// The code is written once against an `Env` (thread identity, barriers, ring hand-off,
// peer tiles):
//   * DeviceEnv  (qsb_kernels.cuh) -- CUDA: named barriers, barrier.cluster, DSMEM
//   * HostEnv    (tests/emu)       -- test-only: OS threads + std::barrier, used to
//                                     check protocol, index math and host compiler on CPU.
//
// Reference semantics implemented here (file:line in the reference tree):
//   StateVector.apply_gate      state_vector.py:41-74   (textbook action; the axis
//                               scramble of :66-73 is bookkeeping done by the host compiler)
//   NoiseModel._apply_channel   noise.py:224-260        (Kraus selection + renormalise)
//   gate formulas               gates.py:37-125
#pragma once

#include <math.h>
#include <stdint.h>
#include <string.h>

#include "qsb.h"

#if defined(__CUDACC__)
#define QSB_HD __host__ __device__ __forceinline__
#define QSB_PASS __device__ __forceinline__   // whole-tile sweeps: inlined into the worker loop (an ABI call would leave them ~48 registers)
#define QSB_CTL __device__ __forceinline__    // control-warp helpers: inlined so that the control state stays in registers (an ABI call spills it to local memory)
typedef double2 c128;
typedef float2 c64;
#else
#define QSB_HD inline
#define QSB_PASS inline
#define QSB_CTL inline
struct alignas(16) c128 { double x, y; };
struct alignas(8) c64 { float x, y; };
#endif

// count trailing zeros / index of the highest set bit of a non-zero 32-bit word
#if defined(__CUDA_ARCH__)
#define QSB_CTZ(x) (__ffs((int)(x)) - 1)
#define QSB_MSB(x) (31 - __clz((int)(x)))
#define QSB_POPC(x) __popc((unsigned)(x))
#else
#define QSB_POPC(x) __builtin_popcount((unsigned)(x))
#define QSB_CTZ(x) __builtin_ctz((unsigned)(x))
#define QSB_MSB(x) (31 - __builtin_clz((unsigned)(x)))
#endif

#define QSB_MAX_QUBITS 16          // resident (cluster) mode
#define QSB_MAX_STREAM_QUBITS 32   // streaming mode (index arithmetic is 32-bit)
#define QSB_MAX_LOCAL_BITS 13      // complex128 tile of 128 KiB; complex64 tiles hold one bit more (QSB_MAX_LOCAL_BITS_C64)
#define QSB_MAX_LOCAL_BITS_C64 14
#define QSB_AD_MARGIN 1e-10
#ifndef QSB_AD_BOUNDS
#define QSB_AD_BOUNDS 1         // amplitude-damping draws outside the certain-K0 range are decided from bounds first (no flush)
#endif
#define QSB_CHUNK 128          // ops staged in shared memory per refill
#ifndef QSB_REMAP_REGS
#define QSB_REMAP_REGS 16      // amplitudes a thread stages per remap / rank-bit flush round
#endif
#ifndef QSB_GROUP_POS
#define QSB_GROUP_POS 1        // 1: bank-conflict-free group order (qsb_group_order), 0: ascending free bits
#endif
#ifndef QSB_REMAP_PUSH
#define QSB_REMAP_PUSH 0      // measured: push (local read + DSMEM store, 3 barriers) 13.5k cycles vs pull 12.8k
#endif
#ifndef QSB_REMAP_HALVES
#define QSB_REMAP_HALVES 1     // exchanges as swaps split between the two CTAs of every pair (remote load + remote store)
#endif
#ifndef QSB_AMPS
#define QSB_AMPS 16           // amplitudes a worker keeps in registers per sweep step (K <= 2)
#endif
#define QSB_PROF_WORDS 128
#define QSB_RING 6             // descriptors in flight between control warp and workers

struct qsb_exec_args {
  const qsb_op* ops;
  int64_t n_ops;
  int64_t ops_stride;     // 0: all units share ops[0..n_ops)
  const double* cdata;
  int64_t n_cdata;
  const int32_t* idata;
  int32_t n, m;
  int32_t load_perm, store_perm, n_snapshots;
  int32_t flags;
  int32_t tile_bits;      // streaming mode: n - m index bits select the tile (0 in resident mode)
  int32_t pad0;
  void* states;           // already offset to `first`; complex128 or complex64 elements (amp_bytes)
  void* states_out;       // STORE destination, already offset (== states when in place)
  int64_t count;          // resident: trajectories; streaming: states (each 2^tile_bits tiles)
  const double* params;   int64_t params_stride;
  const double* uniforms; int64_t uniforms_stride;
  uint64_t seed;          int64_t traj_offset;
  const int64_t* init_basis; int64_t default_basis;
  int32_t* branches;      int64_t branches_stride;
  void* snapshots;
  int64_t amp_bytes;      // 16 (complex128) or 8 (complex64): element size of states / snapshots
  double* probs_accum;
  unsigned long long* prof;   // optional cycle counters, QSB_PROF_WORDS per CTA (qsb_debug_profile), or NULL
  const void* const* peer_ptrs;   // streamed LOAD from the peers' shards (exchange folded into the pass), or NULL
  int32_t peer_shift;  int32_t pad1;
  int64_t peer_rank_or;
};

// ---- descriptors: control warp -> workers ---------------------------------------------
enum { QSB_D_EXIT = 0, QSB_D_INIT, QSB_D_SWEEP, QSB_D_REMAP, QSB_D_GFLUSH, QSB_D_RDM1, QSB_D_STORE };
// QSB_CLS_* (structure class of a pending 2x2) and QSB_G_* (gate applied inside a sweep after the pending matrices of
// its bits) are part of the C ABI since the streamed passes take host-fused sweeps: see include/qsb.h

struct alignas(16) qsb_desc {
  int32_t kind, gate, k, flags;
  int32_t b[4];            // slot bits, b[0] = MSB of the gate's matrix index
  int32_t cls[4];          // class of the pending matrix of b[k]
  int64_t unit;            // INIT: unit index (trajectory; streaming: state * tiles + tile)
  int64_t basis;           // INIT: reference-order basis index
  const int32_t* perm;     // INIT / STORE: n-entry bit permutation (slot bit -> reference-order bit)
  void* gptr;              // INIT: source state (LOAD) ; STORE: destination (or NULL); elements of the tile's type
  double* probs;           // STORE: |psi|^2 accumulation target (or NULL)
  int64_t tile;            // INIT / STORE: value of the non-resident index bits (cluster rank or tile id)
  uint64_t pos;            // SWEEP: 16 nibbles, nibble t = index bit that bit t of the group number lands on
  int32_t hmask;           // SWEEP: index bits the group-number bits above log2(W) land on (ascending)
  int32_t pad1;
  c128 P[3][4];            // pending matrices (row-major), valid where cls != NONE
  // SWEEP without a dense gate (filled by the control warp, qsb_sweep_tables): swizzled tile offsets local index r is
  // loaded from / stored to (the permutation gates only move the store), and its real factor (damping scales of the
  // bits without a full 2x2, sign of CZ)
  // SWEEP: tile-index bits of a worker's first group = tabl[worker & 31] | tabw[worker >> 5] (the group-number bits below
  // log2(W) laid onto `pos`; one entry per control lane, so the workers replace a dependent bit loop by two loads)
  alignas(16) int32_t tabl[32];
  alignas(16) int32_t tabw[8];
  alignas(16) int32_t off[8];
  alignas(16) int32_t ost[8];
  alignas(16) double f[8];
  c128 mat[64];            // dense 4x4 / 8x8 gate of this sweep
};

// decoded op: what the serial part of the control warp does with it
enum { QSB_DEC_SKIP = 0, QSB_DEC_MUL = 1, QSB_DEC_SLOW = 2 };
struct alignas(16) qsb_dec {
  int32_t type;            // SKIP | MUL: pend[b] <- U pend[b] | SLOW: needs the descriptor ring / the state
  int32_t b;               // slot bit
  int32_t ucls;            // structure class of U
  int32_t next;            // MUL: index (in the chunk) of the next MUL op on the same slot, or QSB_CHUNK
  c128 U[4];               // MUL: the matrix.  Multi-qubit gate ops (SLOW): U[0] holds the sweep's group order
};                         //   (bits of pos in .x, hmask in .y), worked out during the lane-parallel decode

// one staged chunk of the op list: filled by the decode warp (records, uniforms, decoded form, per-slot lists,
// SLOW-op mask), consumed by the control warp; two of them so that decoding runs one chunk ahead
struct alignas(16) qsb_chunk {
  qsb_op ops[QSB_CHUNK];       // staged op records ...
  double u[QSB_CHUNK];         // ... the uniform each Kraus op consumes
  qsb_dec dec[QSB_CHUNK];      // ... and their decoded form
  int32_t head[32];            // first MUL op of each slot in the chunk (QSB_CHUNK = none)
  uint32_t slowmask[QSB_CHUNK / 32];   // bit i: op i needs the SLOW path
};

// per-CTA control block that lives behind the tile in shared memory
struct qsb_ctl {
  uint32_t perm[1024];         // bit-permutation byte tables (workers: INIT / STORE)
  c128 pend[32][4];            // pending 2x2 per slot bit (row-major); written by the control warp only
  qsb_desc ring[QSB_RING];
  qsb_chunk chunk[2];
  double wpart[32 * 4];        // per-warp partial sums (workers)
  double red[2][4];            // this CTA's contribution to a cluster reduction, double-buffered
  double red_total[4];         // cluster-wide result handed to the control warp (RDM1)
  double redx[2][8][4];        // RDM1: the partial sums every CTA of the cluster pushed here (by source rank), double-buffered
  double wtab[256];            // weighted marginal: products of diagonal-pending weights over index bits 0..6 | 7..13
  double slotw[32][4];         // amplitude-damping draw: diag(P^H P) of each slot's pending matrix and its off-diagonal ratio (control warp)
  unsigned long long xbar;     // mbarrier of the workers-only cluster barrier (device)
  unsigned long long dbar[4];  // decode warp <-> control warp: chunk[b] full (b), chunk[b] empty (2 + b) (device)
};

// ---- small helpers ---------------------------------------------------------------
// Swizzled slot of amplitude i inside the tile: XOR bits 3..5 into bits 0..2 so that the
// 8 lanes of a quarter-warp hit 8 different 16-byte bank groups for every target-bit choice.
QSB_HD int qsb_slot(int i) { return i ^ ((i >> 3) & 7); }
// insert a zero bit at position b
QSB_HD int qsb_ins0(int g, int b) { return ((g >> b) << (b + 1)) | (g & ((1 << b) - 1)); }

QSB_HD c128 qsb_c(double x, double y) { c128 r; r.x = x; r.y = y; return r; }
// bit casts (a 64-bit group order rides in a double field of the decode record)
#if defined(__CUDA_ARCH__)
QSB_HD double qsb_u64_as_double(uint64_t v) { return __longlong_as_double((long long)v); }
QSB_HD uint64_t qsb_double_as_u64(double v) { return (uint64_t)__double_as_longlong(v); }
#else
QSB_HD double qsb_u64_as_double(uint64_t v) { double d; memcpy(&d, &v, 8); return d; }
QSB_HD uint64_t qsb_double_as_u64(double v) { uint64_t u; memcpy(&u, &v, 8); return u; }
#endif
QSB_HD c128 qsb_mul(c128 a, c128 b) { return qsb_c(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
QSB_HD c128 qsb_fma(c128 a, c128 b, c128 c) {   // a*b + c
  return qsb_c(fma(a.x, b.x, fma(-a.y, b.y, c.x)), fma(a.x, b.y, fma(a.y, b.x, c.y)));
}
QSB_HD c128 qsb_neg(c128 a) { return qsb_c(-a.x, -a.y); }
QSB_HD double qsb_norm2(c128 a) { return a.x * a.x + a.y * a.y; }

// ---- amplitude types: complex128 (default) and complex64 (BASELINE's separately reported 1e-5 mode) ----------
// SW = log2(elements per 128-byte shared-memory row): the XOR swizzle folds index bits SW..2SW-1 onto bits 0..SW-1.
template <class A> struct qsb_amp;
template <> struct qsb_amp<c128> { typedef double real; enum { SW = 3 }; };
template <> struct qsb_amp<c64> { typedef float real; enum { SW = 4 }; };
QSB_HD c64 qsb_cf(float x, float y) { c64 r; r.x = x; r.y = y; return r; }
QSB_HD c64 qsb_mul(c64 a, c64 b) { return qsb_cf(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
QSB_HD c64 qsb_fma(c64 a, c64 b, c64 c) {
  return qsb_cf(fmaf(a.x, b.x, fmaf(-a.y, b.y, c.x)), fmaf(a.x, b.y, fmaf(a.y, b.x, c.y)));
}
QSB_HD c64 qsb_neg(c64 a) { return qsb_cf(-a.x, -a.y); }
QSB_HD double qsb_norm2(c64 a) { return (double)a.x * a.x + (double)a.y * a.y; }
template <class A> QSB_HD A qsb_cvt(c128 z);
template <> QSB_HD c128 qsb_cvt<c128>(c128 z) { return z; }
template <> QSB_HD c64 qsb_cvt<c64>(c128 z) { return qsb_cf((float)z.x, (float)z.y); }
QSB_HD c128 qsb_wide(c128 z) { return z; }
QSB_HD c128 qsb_wide(c64 z) { return qsb_c(z.x, z.y); }
template <class A> QSB_HD A qsb_scale(A a, double s) {
  typedef typename qsb_amp<A>::real R;
  A r; r.x = a.x * (R)s; r.y = a.y * (R)s; return r;
}
template <int SW> QSB_HD int qsb_slot_sw(int i) { return i ^ ((i >> SW) & ((1 << SW) - 1)); }

QSB_HD void qsb_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                       uint32_t out[4]) {
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// uniform in [0,1) with 53 bits; draw d of trajectory t: block d/2, words 2*(d%2), 2*(d%2)+1
QSB_HD double qsb_philox_uniform(uint64_t seed, uint64_t traj, uint32_t d) {
  uint32_t r[4];
  qsb_philox((uint32_t)traj, (uint32_t)(traj >> 32), d >> 1, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
  uint32_t lo = r[2 * (d & 1)], hi = r[2 * (d & 1) + 1];
  return ((double)(hi >> 5) * 67108864.0 + (double)(lo >> 6)) * (1.0 / 9007199254740992.0);
}

// ascending sort of K <= 3 slot bits (branch-free: keeps the array in registers)
template <int K>
QSB_HD void qsb_sort_bits(int* b) {
  if (K >= 2) { int lo = b[0] < b[1] ? b[0] : b[1], hi = b[0] < b[1] ? b[1] : b[0]; b[0] = lo; b[1] = hi; }
  if (K >= 3) {
    int lo = b[1] < b[2] ? b[1] : b[2], hi = b[1] < b[2] ? b[2] : b[1]; b[1] = lo; b[2] = hi;
    lo = b[0] < b[1] ? b[0] : b[1]; hi = b[0] < b[1] ? b[1] : b[0]; b[0] = lo; b[1] = hi;
  }
}

// numpy Generator.choice(k, p) from its one uniform: cdf = cumsum(p / p.sum()); cdf /= cdf[-1];
// searchsorted(cdf, u, side='right')  (noise.py:248-254)
QSB_HD int qsb_choice(const double* p, int k, double u) {
  double cdf[8], tot = 0.0;
  for (int i = 0; i < k; ++i) tot += p[i];
  double s = 0.0;
  for (int i = 0; i < k; ++i) { s += (tot > 1e-15 ? p[i] / tot : p[i]); cdf[i] = s; }
  int idx = 0;
  for (int i = 0; i < k; ++i) if (cdf[i] / cdf[k - 1] <= u) ++idx;
  return idx < k ? idx : k - 1;
}

// Order in which the bits of a group number are laid onto the non-target index bits of the tile.
// The swizzled slot of index i has 16-byte bank group (i0^i3, i1^i4, i2^i5); the 8 lanes of a
// quarter-warp differ in group bits 0..2, so these go to one free position out of each pair
// {0,3}, {1,4}, {2,5}: every LDS.128 / STS.128 of a sweep is then conflict-free whatever the targets
// (unless the targets cover both members of a pair).  The remaining positions follow in ascending order.
QSB_HD uint64_t qsb_group_order(int m, uint32_t used, int wbits, int nfree, int sw, int* hmask) {
  const uint32_t all = m >= 32 ? 0xffffffffu : ((1u << m) - 1u);
  uint64_t pos = 0;
  int cnt = 0, hm = 0;
  for (int j = 0; j < sw; ++j) {
    int q = -1;
    if (j < m && !((used >> j) & 1u)) q = j;
    else if (j + sw < m && !((used >> (j + sw)) & 1u)) q = j + sw;
    if (q >= 0) {
      pos |= (uint64_t)q << (4 * cnt);
      if (cnt >= wbits && cnt < nfree) hm |= 1 << q;
      ++cnt;
      used |= 1u << q;
    }
  }
  uint32_t rest = ~used & all;
  while (rest && cnt < 16) {
    const int q = QSB_CTZ(rest);
    pos |= (uint64_t)q << (4 * cnt);
    if (cnt >= wbits && cnt < nfree) hm |= 1 << q;
    ++cnt;
    rest &= rest - 1;
  }
  *hmask = hm;
  return pos;
}
// the same result computed by the lanes of the control warp together: the (at most `sw`) conflict-avoiding
// positions serially, then index bit q finds its own place among the remaining ones with a population count and
// the lanes OR their nibbles together
template <class Env>
QSB_CTL uint64_t qsb_group_order_lanes(Env& env, int m, uint32_t used, int wbits, int nfree, int sw, int* hmask) {
  const uint32_t all = m >= 32 ? 0xffffffffu : ((1u << m) - 1u);
  uint32_t plo = 0, phi = 0, hm = 0;
  int cnt = 0;
  for (int j = 0; j < sw; ++j) {
    int q = -1;
    if (j < m && !((used >> j) & 1u)) q = j;
    else if (j + sw < m && !((used >> (j + sw)) & 1u)) q = j + sw;
    if (q >= 0) {
      if (env.lead) {
        if (cnt < 8) plo |= (uint32_t)q << (4 * cnt); else phi |= (uint32_t)q << (4 * (cnt - 8));
        if (cnt >= wbits && cnt < nfree) hm |= 1u << q;
      }
      ++cnt;
      used |= 1u << q;
    }
  }
  const uint32_t rest = ~used & all;
  for (int q = env.clane; q < 32; q += env.CL) {
    if (!((rest >> q) & 1u)) continue;
    const int c = cnt + QSB_POPC(rest & ((1u << q) - 1u));
    if (c >= 16) continue;
    if (c < 8) plo |= (uint32_t)q << (4 * c); else phi |= (uint32_t)q << (4 * (c - 8));
    if (c >= wbits && c < nfree) hm |= 1u << q;
  }
  plo = env.or_reduce(plo); phi = env.or_reduce(phi); hm = env.or_reduce(hm);
  *hmask = (int)hm;
  return ((uint64_t)phi << 32) | plo;
}
// deposit the low `nbits` bits of g onto the positions packed in `pos`
QSB_HD int qsb_deposit(int g, uint64_t pos, int nbits) {
  int r = 0;
  for (int t = 0; t < nbits; ++t) r |= ((g >> t) & 1) << (int)((pos >> (4 * t)) & 15u);
  return r;
}

// bit-permutation byte tables: index x (<= 32 bits) -> OR_j bit_j(x) << perm[j]
QSB_HD uint32_t qsb_permute(const uint32_t* tab, uint32_t x) {
  return tab[x & 255] | tab[256 + ((x >> 8) & 255)] | tab[512 + ((x >> 16) & 255)] | tab[768 + (x >> 24)];
}

// =========================================================================================
//                                      WORKER SIDE
// =========================================================================================

// swizzled slot for the tile's element type (Env in scope)
#define QSB_SLOT(i) qsb_slot_sw<qsb_amp<typename Env::amp>::SW>(i)

// One sweep over the tile for K slot bits (d->b[0] = MSB of the local index r): apply the pending 2x2 of
// every bit whose class is not NONE, then gate G.  Each worker owns whole 2^K groups and keeps 16
// amplitudes (NG groups) in registers per step: all loads are issued before the arithmetic, and the
// (CTA-uniform) class branches are taken once per step instead of once per group.
// amplitudes a worker keeps in registers per step: 16 (all loads in flight before the arithmetic); 8 for K = 3,
// where 16 amplitudes plus three dense pending matrices would spill (measured: profiles/README.md)
template <int K, bool DG, class Env, int AMPS = QSB_AMPS>
QSB_PASS void qsb_sweep(Env& env, int m, const qsb_desc* d) {
  typedef typename Env::amp A;
  A* tile = env.tile();             // re-derived here so device code keeps the shared address space (LDS/STS)
  const unsigned long long sq0 = env.prof_on() ? env.clock() : 0;
  constexpr int D = 1 << K;
  constexpr int NG = DG ? (K == 2 && AMPS >= 16 ? 2 : 1) : (K == 3 ? 1 : (AMPS / D > 0 ? AMPS / D : 1));
  const int G = d->gate;
  int bits[K], cls[K];
#pragma unroll
  for (int k = 0; k < K; ++k) { bits[k] = d->b[k]; cls[k] = d->cls[k]; }
  // group g = wid + j W (W = 2^wbits): the bits of wid are laid onto the tile index once per sweep; the
  // bits of j land on `hmask` in ascending order, so consecutive j are a masked increment
#if QSB_GROUP_POS
  const int lo_base = d->tabl[env.wid & 31] | d->tabw[env.wid >> 5];
  const int hmask = d->hmask;
  int hi = 0;
#else
  int sb[K];
#pragma unroll
  for (int k = 0; k < K; ++k) sb[k] = bits[k];
  qsb_sort_bits<K>(sb);
#endif
  // off[r] = SWIZZLED slot offset of local index r: the swizzle is linear over XOR and a group's base has none of
  // the target bits, so slot(base | o) = slot(base) ^ slot(o) -- one XOR per access instead of a shift / mask / XOR chain
  int off[D];
#pragma unroll
  for (int r = 0; r < D; ++r) {
    int o = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) if ((r >> (K - 1 - k)) & 1) o |= 1 << bits[k];
    off[r] = QSB_SLOT(o);
  }
  A P[K][4];                        // always a valid matrix (identity where nothing is pending)
#pragma unroll
  for (int k = 0; k < K; ++k) {
#pragma unroll
    for (int e = 0; e < 4; ++e) P[k][e] = qsb_cvt<A>(d->P[k][e]);
  }
  const int cnt = 1 << (m - K);
  if (env.prof_on()) env.prof_add(124, env.clock() - sq0);      // sweep set-up (descriptor fields, offsets, matrices)
  for (int g0 = env.wid; g0 < cnt; g0 += NG * env.W) {
    A a[NG][D];
    int base[NG];
#pragma unroll
    for (int j = 0; j < NG; ++j) {
#if QSB_GROUP_POS
      const int bs = lo_base | hi;        // out-of-range groups of a short tile wrap onto valid ones (never stored)
      hi = ((hi | ~hmask) + 1) & hmask;
#else
      const int g = g0 + j * env.W;
      int bs = g < cnt ? g : g0;          // out-of-range groups of the last step alias a valid one (never stored)
#pragma unroll
      for (int k = 0; k < K; ++k) bs = qsb_ins0(bs, sb[k]);
#endif
      base[j] = QSB_SLOT(bs);
#pragma unroll
      for (int r = 0; r < D; ++r) a[j][r] = tile[base[j] ^ off[r]];
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const int bit = 1 << (K - 1 - k);
      if (cls[k] >= QSB_CLS_DIAG) {          // complex-diagonal matrices take the dense path (keeps the code small)
#pragma unroll
        for (int j = 0; j < NG; ++j)
#pragma unroll
          for (int r = 0; r < D; ++r) {
            if (r & bit) continue;
            A lo = a[j][r], hi = a[j][r | bit];
            a[j][r] = qsb_fma(P[k][1], hi, qsb_mul(P[k][0], lo));
            a[j][r | bit] = qsb_fma(P[k][3], hi, qsb_mul(P[k][2], lo));
          }
      } else if (cls[k] == QSB_CLS_RDIAG) {
        const typename qsb_amp<A>::real sc = P[k][3].x;
#pragma unroll
        for (int j = 0; j < NG; ++j)
#pragma unroll
          for (int r = 0; r < D; ++r) if (r & bit) { a[j][r].x *= sc; a[j][r].y *= sc; }
      }
    }
    if (DG) {
      // every output row is written straight to the tile: the group is owned by this thread and all of
      // its inputs are already in registers
#pragma unroll
      for (int r = 0; r < D; ++r) {
        A acc[NG];
#pragma unroll
        for (int c = 0; c < D; ++c) {
          const A mv = qsb_cvt<A>(d->mat[r * D + c]);
#pragma unroll
          for (int j = 0; j < NG; ++j) acc[j] = c == 0 ? qsb_mul(mv, a[j][0]) : qsb_fma(mv, a[j][c], acc[j]);
        }
#pragma unroll
        for (int j = 0; j < NG; ++j)
          if (g0 + j * env.W < cnt) tile[base[j] ^ off[r]] = acc[j];
      }
      continue;
    }
#pragma unroll
    for (int j = 0; j < NG; ++j) {
      A* x = a[j];
      // the gate is a CTA-uniform run-time choice: one sweep body per K keeps the instruction footprint small
      if (K == 2) {
        if (G == QSB_G_CX) { A t = x[2]; x[2] = x[3]; x[3] = t; }                // b[0] control, b[1] target
        else if (G == QSB_G_CZ) { x[3] = qsb_neg(x[3]); }
        else if (G == QSB_G_SWAP) { A t = x[1]; x[1] = x[2]; x[2] = t; }
      } else if (K == 3) {
        if (G == QSB_G_CCX) { A t = x[D - 2]; x[D - 2] = x[D - 1]; x[D - 1] = t; }   // 110 <-> 111
        else if (G == QSB_G_CSWAP) { A t = x[(D >> 1) | 1]; x[(D >> 1) | 1] = x[(D >> 1) | 2]; x[(D >> 1) | 2] = t; }  // 101 <-> 110
      }
      if (g0 + j * env.W < cnt) {
#pragma unroll
        for (int r = 0; r < D; ++r) tile[base[j] ^ off[r]] = x[r];
      }
    }
  }
}

// The same sweep with the run-time choices taken out of the loop.  DM (compile time): bit k set = the pending matrix of
// b[k] is a full 2x2.  Everything else is data: the other bits carry a real scale of their |1> half (the damping
// K0's; 1.0 when nothing is pending), folded with the sign of CZ into one factor per local index, and the permutation
// gates (CX, SWAP, Toffoli, Fredkin) only change where a register is STORED.  With the class and gate branches inside
// the loop (qsb_sweep above) ptxas cannot move a group's arithmetic under the loads and stores of its neighbours and
// the workers of a CTA, which enter a sweep together, alternate between the shared-memory pipe and the FP64 pipe:
// tools/micro/sweep_real.cu, one dense pending matrix: 3 720 -> 2 620 cycles per sweep.
// what every specialised sweep reads from its descriptor before the first tile load; fetched by the worker loop TOGETHER
// with the descriptor header, so the loads are in flight while the kind / variant jumps resolve
struct qsb_sweep_pro {
  int lo_base, hmask;
  int off[8], ost[8];
  double f[8];
};
template <class Env>
QSB_HD void qsb_sweep_prologue(Env& env, const qsb_desc* d, qsb_sweep_pro& p) {
  p.lo_base = d->tabl[env.wid & 31] | d->tabw[env.wid >> 5];
  p.hmask = d->hmask;
#pragma unroll
  for (int r = 0; r < 8; ++r) { p.off[r] = d->off[r]; p.ost[r] = d->ost[r]; p.f[r] = d->f[r]; }
}

template <int K, int DM, class Env>
QSB_PASS void qsb_sweep_s(Env& env, int m, const qsb_desc* d, const qsb_sweep_pro& pro) {
  typedef typename Env::amp A;
  typedef typename qsb_amp<A>::real R;
  A* tile = env.tile();
  const unsigned long long sq0 = env.prof_on() ? env.clock() : 0;
  constexpr int D = 1 << K;
  constexpr int ND = (DM & 1) + ((DM >> 1) & 1) + ((DM >> 2) & 1);
  constexpr int DIDX = ((DM & 1) ? (1 << (K - 1)) : 0) | ((K > 1 && (DM & 2)) ? (1 << (K - 2)) : 0) | ((K > 2 && (DM & 4)) ? 1 : 0);   // local-index bits with a full 2x2
  constexpr int NG = (K == 3 && ND >= 2) ? 1 : (QSB_AMPS / D > 0 ? QSB_AMPS / D : 1);   // 16 amplitudes + 3 matrices would spill
  const int lo_base = pro.lo_base;
  const int hmask = pro.hmask;
  int hi = 0;
  int off[D], ost[D];               // swizzled offsets: where local index r is loaded from / stored to
  R f[D];                           // real factor of local index r (scales of the non-dense bits, sign of CZ)
#pragma unroll
  for (int r = 0; r < D; ++r) { off[r] = pro.off[r]; ost[r] = pro.ost[r]; f[r] = (R)pro.f[r]; }
  A P[K][4];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    if (!((DM >> k) & 1)) continue;
#pragma unroll
    for (int e = 0; e < 4; ++e) P[k][e] = qsb_cvt<A>(d->P[k][e]);
  }
  const int cnt = 1 << (m - K);
  if (env.prof_on()) env.prof_add(124, env.clock() - sq0);
  for (int g0 = env.wid; g0 < cnt; g0 += NG * env.W) {
    A a[NG][D];
    int base[NG];
#pragma unroll
    for (int j = 0; j < NG; ++j) {
      const int bs = lo_base | hi;        // out-of-range groups of a short tile wrap onto valid ones (never stored)
      hi = ((hi | ~hmask) + 1) & hmask;
      base[j] = QSB_SLOT(bs);
#pragma unroll
      for (int r = 0; r < D; ++r) a[j][r] = tile[base[j] ^ off[r]];
    }
#pragma unroll
    for (int j = 0; j < NG; ++j) {
#pragma unroll
      for (int k = 0; k < K; ++k) {
        if (!((DM >> k) & 1)) continue;
        const int bit = 1 << (K - 1 - k);
#pragma unroll
        for (int r = 0; r < D; ++r) {
          if (r & bit) continue;
          const A lo = a[j][r], up = a[j][r | bit];
          a[j][r] = qsb_fma(P[k][1], up, qsb_mul(P[k][0], lo));
          a[j][r | bit] = qsb_fma(P[k][3], up, qsb_mul(P[k][2], lo));
        }
      }
      if (g0 + j * env.W < cnt) {
#pragma unroll
        for (int r = 0; r < D; ++r) {
          A v = a[j][r];
          // a factor other than 1 needs a |1> on a bit without a full 2x2, or is the sign of CZ on |11>
          if ((r & ~DIDX) != 0 || (K == 2 && r == D - 1)) { v.x *= f[r]; v.y *= f[r]; }
          tile[base[j] ^ ost[r]] = v;
        }
      }
    }
  }
}

// first 16 bytes of a descriptor, read with one load
struct alignas(16) qsb_desc_hdr { int32_t kind, gate, k, flags; };

template <class Env>
QSB_HD void qsb_do_sweep(Env& env, int m, const qsb_desc* d, const qsb_desc_hdr& h, const qsb_sweep_pro& pro) {
  const bool dg = h.gate == QSB_G_DENSE;
#if QSB_GROUP_POS
  if (!dg) {
    switch (h.k * 8 + h.flags) {            // flags = dense mask (qsb_emit_sweep)
      case 8: qsb_sweep_s<1, 0>(env, m, d, pro); break;
      case 9: qsb_sweep_s<1, 1>(env, m, d, pro); break;
      case 16: qsb_sweep_s<2, 0>(env, m, d, pro); break;
      case 17: qsb_sweep_s<2, 1>(env, m, d, pro); break;
      case 18: qsb_sweep_s<2, 2>(env, m, d, pro); break;
      case 19: qsb_sweep_s<2, 3>(env, m, d, pro); break;
      case 24: qsb_sweep_s<3, 0>(env, m, d, pro); break;
      case 25: qsb_sweep_s<3, 1>(env, m, d, pro); break;
      case 26: qsb_sweep_s<3, 2>(env, m, d, pro); break;
      case 27: qsb_sweep_s<3, 3>(env, m, d, pro); break;
      case 28: qsb_sweep_s<3, 4>(env, m, d, pro); break;
      case 29: qsb_sweep_s<3, 5>(env, m, d, pro); break;
      case 30: qsb_sweep_s<3, 6>(env, m, d, pro); break;
      case 31: qsb_sweep_s<3, 7>(env, m, d, pro); break;
      default: break;
    }
    return;
  }
  // dense 4x4 / 8x8 gates: the generic sweep (run-time pending-matrix classes)
  if (h.k == 2) qsb_sweep<2, true>(env, m, d);
  else if (h.k == 3) qsb_sweep<3, true>(env, m, d);
#else
  (void)pro;
  if (h.k == 1) qsb_sweep<1, false>(env, m, d);
  else if (h.k == 2) { if (dg) qsb_sweep<2, true>(env, m, d); else qsb_sweep<2, false>(env, m, d); }
  else if (h.k == 3) { if (dg) qsb_sweep<3, true>(env, m, d); else qsb_sweep<3, false>(env, m, d); }
#endif
}

// sum v[0..nv) over the workers of this CTA; every worker gets the bit-identical result
template <class Env>
QSB_HD void qsb_block_reduce(Env& env, double* v, int nv) {
  double* wp = env.ctl()->wpart;
  env.sync_workers();                         // previous users of wpart are done
  for (int k = 0; k < nv; ++k) {
    double x = env.warp_sum(v[k]);
    if (env.lane == 0) wp[env.warp * 4 + k] = x;
  }
  env.sync_workers();
  for (int k = 0; k < nv; ++k) {
    double s = 0.0;
    for (int w = 0; w < env.nwarps; ++w) s += wp[w * 4 + k];
    v[k] = s;
  }
}

// partial sums for the 1-qubit reduced density matrix of slot bit b (unnormalised), this CTA's share:
// v[0] = sum |a0|^2, v[1] = sum |a1|^2, v[2] + i v[3] = sum a0 conj(a1).  b >= m is a cluster-rank bit:
// the whole tile belongs to one side (off-diagonal terms are not available there and are not requested).
template <class Env>
QSB_PASS void qsb_partial_rdm1(Env& env, int m, int b, double v[4]) {
  typedef typename Env::amp A;
  const A* tile = env.tile();
  v[0] = v[1] = v[2] = v[3] = 0.0;
  if (b >= m) {
    double s = 0.0;
    for (int i = env.wid; i < (1 << m); i += env.W) s += qsb_norm2(tile[i]);
    v[(env.rank >> (b - m)) & 1] = s;
    return;
  }
  const int cnt = 1 << (m - 1);
  for (int g = env.wid; g < cnt; g += env.W) {
    int i0 = qsb_ins0(g, b);
    const c128 a0 = qsb_wide(tile[QSB_SLOT(i0)]), a1 = qsb_wide(tile[QSB_SLOT(i0 | (1 << b))]);
    v[0] += qsb_norm2(a0);
    v[1] += qsb_norm2(a1);
    v[2] += a0.x * a1.x + a0.y * a1.y;
    v[3] += a0.y * a1.x - a0.x * a1.y;
  }
}

// 1-qubit reduced density matrix of slot bit b under per-slot diagonal weights: wt[j][v] = weight of value v of slot j
// (|P_j[v][v]|^2 of a diagonal pending matrix, which needs no flush: a diagonal only reweights |amplitude|^2; the
// diagonal of P_j^H P_j in the bounded amplitude-damping draw).  W(x) = prod_j wt[j][x_j]; rank bits contribute this
// CTA's constant factor.  v[0], v[1] = sum over x with x_b = 0 / 1 of W(x) |phi_x|^2; for a local bit b whose own
// weights are (1, 1) also v[2] + i v[3] = sum W(x) phi_x0 conj(phi_x1).
template <class Env>
QSB_PASS void qsb_partial_marginal_w(Env& env, int m, int b, const double* wt, double v[4]) {
  typedef typename Env::amp A;
  const A* tile = env.tile();
  double* tab = env.ctl()->wtab;
  const int nlo = m < 7 ? m : 7, nhi = m - nlo;
  env.sync_workers();                                    // previous users of wtab are done
  for (int e = env.wid; e < 256; e += env.W) {
    double w = 1.0;
    if (e < 128) { for (int j = 0; j < nlo; ++j) w *= wt[2 * j + ((e >> j) & 1)]; }
    else { for (int j = 0; j < nhi; ++j) w *= wt[2 * (nlo + j) + (((e - 128) >> j) & 1)]; }
    tab[e] = w;
  }
  env.sync_workers();
  double crank = 1.0;
  for (int g = 0; (1 << g) < env.C; ++g) crank *= wt[2 * (m + g) + ((env.rank >> g) & 1)];
  v[0] = v[1] = v[2] = v[3] = 0.0;
  if (b >= m) {
    double s = 0.0;
    for (int i = env.wid; i < (1 << m); i += env.W) s += tab[i & 127] * tab[128 + (i >> 7)] * qsb_norm2(tile[QSB_SLOT(i)]);
    v[(env.rank >> (b - m)) & 1] = s * crank;
    return;
  }
  const int cnt = 1 << (m - 1);
  // four pairs per step: all the loads (two amplitudes and four table entries per pair) are in flight together
  for (int g = env.wid; g < cnt; g += 4 * env.W) {
    c128 a0[4], a1[4];
    double w0[4], w1[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gg = g + j * env.W;
      const bool ok = gg < cnt;
      const int i0 = qsb_ins0(ok ? gg : g, b), i1 = i0 | (1 << b);
      a0[j] = qsb_wide(tile[QSB_SLOT(i0)]); a1[j] = qsb_wide(tile[QSB_SLOT(i1)]);
      w0[j] = ok ? tab[i0 & 127] * tab[128 + (i0 >> 7)] : 0.0;
      w1[j] = ok ? tab[i1 & 127] * tab[128 + (i1 >> 7)] : 0.0;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[0] += w0[j] * qsb_norm2(a0[j]);
      v[1] += w1[j] * qsb_norm2(a1[j]);
      v[2] += w0[j] * (a0[j].x * a1[j].x + a0[j].y * a1[j].y);   // meaningful when wt[b] = (1, 1): w0 is the pair's weight
      v[3] += w0[j] * (a0[j].y * a1[j].x - a0[j].x * a1[j].y);
    }
  }
  v[0] *= crank; v[1] *= crank; v[2] *= crank; v[3] *= crank;
}

template <class Env>
QSB_HD void qsb_build_perm(Env& env, const int32_t* perm, int n) {
  uint32_t* tab = env.ctl()->perm;
  for (int v = env.wid; v < 1024; v += env.W) {
    const int lo = v & 255, byte = v >> 8;
    uint32_t r = 0;
    for (int j = 0; j < 8; ++j) {
      const int bit = byte * 8 + j;
      if (bit < n && ((lo >> j) & 1)) r |= 1u << perm[bit];
    }
    tab[v] = r;
  }
  env.sync_workers();
}

template <class Env>
QSB_PASS void qsb_do_init(Env& env, const qsb_exec_args& a, const qsb_desc* d) {
  typedef typename Env::amp A;
  A* tile = env.tile();
  const uint32_t* tab = env.ctl()->perm;
  const int m = a.m;
  qsb_build_perm(env, d->perm, a.n);
  const uint32_t hi = (uint32_t)d->tile << m;
  const bool hoist = (env.W & 31) == 0 && m >= 5;      // the low 5 index bits are then local and equal to the lane
  const uint32_t lo = hoist ? qsb_permute(tab, (uint32_t)env.wid & 31u) : 0u;
  if (d->flags & QSB_RUN_LOAD) {
    const A* src = static_cast<const A*>(d->gptr);
    // 8 loads in flight per worker: addresses first (the table look-ups and the tile stores are both shared
    // memory, so the compiler will not reorder them itself), then the global loads, then the stores
    for (int i0 = env.wid; i0 < (1 << m); i0 += 8 * env.W) {
      uint32_t s[8];
      A v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int i = i0 + e * env.W;
        const uint32_t x = hi | (uint32_t)(i < (1 << m) ? i : i0);
        s[e] = hoist ? (qsb_permute(tab, x & ~31u) | lo) : qsb_permute(tab, x);
      }
      if (a.peer_ptrs) {
        // the shard this tile comes from sits on another GPU: element s of the post-exchange shard = peer
        // (s >> shift), offset (s & mask) | my rank field -- plain loads over NVLink peer mappings
        const uint32_t pmask = (1u << a.peer_shift) - 1u;
#pragma unroll
        for (int e = 0; e < 8; ++e)
          v[e] = static_cast<const A*>(a.peer_ptrs[s[e] >> a.peer_shift])[(int64_t)(s[e] & pmask) | a.peer_rank_or];
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = src[s[e]];
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int i = i0 + e * env.W;
        if (i < (1 << m)) tile[QSB_SLOT(i)] = v[e];
      }
    }
    // LOAD + STORE run in place and the two bit permutations differ, so a CTA's stores land on addresses
    // another CTA of the cluster loads from: nobody may go on before every CTA has its tile
    if (env.C > 1) env.cluster_sync_w();
  } else {
    const uint32_t basis = (uint32_t)d->basis;
    for (int i = env.wid; i < (1 << m); i += env.W) {
      const uint32_t x = hi | (uint32_t)i;
      const uint32_t s = hoist ? (qsb_permute(tab, x & ~31u) | lo) : qsb_permute(tab, x);
      tile[QSB_SLOT(i)] = qsb_cvt<A>(qsb_c(s == basis ? 1.0 : 0.0, 0.0));
    }
  }
}

// write the tile (normalised when asked) to gptr[perm(x)], x = tile << m | i
template <class Env>
QSB_PASS void qsb_do_store(Env& env, const qsb_exec_args& a, const qsb_desc* d, int& parity) {
  typedef typename Env::amp A;
  const A* tile = env.tile();
  const uint32_t* tab = env.ctl()->perm;
  const int m = a.m;
  double scale = 1.0;
  if (d->flags & QSB_RUN_NORMALIZE) {
    double v[4] = {0.0, 0.0, 0.0, 0.0};
    for (int i = env.wid; i < (1 << m); i += env.W) v[0] += qsb_norm2(tile[i]);   // slot order irrelevant
    qsb_block_reduce(env, v, 1);
    if (env.C > 1) {
      if (env.wid == 0) env.ctl()->red[parity][0] = v[0];
      env.cluster_sync_w();
      double s = 0.0;
      for (int r = 0; r < env.C; ++r) s += env.peer_ctl(r)->red[parity][0];
      v[0] = s;
    }
    parity ^= 1;
    if (v[0] > 1e-30) scale = 1.0 / sqrt(v[0]);
  }
  qsb_build_perm(env, d->perm, a.n);
  const uint32_t hi = (uint32_t)d->tile << m;
  const bool hoist = (env.W & 31) == 0 && m >= 5;      // the low 5 index bits are then local and equal to the lane
  const uint32_t lo = hoist ? qsb_permute(tab, (uint32_t)env.wid & 31u) : 0u;
  A* out = static_cast<A*>(d->gptr);
  double* probs = d->probs;
  for (int i0 = env.wid; i0 < (1 << m); i0 += 8 * env.W) {
    uint32_t dst[8];
    A v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int i = i0 + e * env.W;
      const int ii = i < (1 << m) ? i : i0;
      const uint32_t x = hi | (uint32_t)ii;
      dst[e] = hoist ? (qsb_permute(tab, x & ~31u) | lo) : qsb_permute(tab, x);
      v[e] = qsb_scale(tile[QSB_SLOT(ii)], scale);
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int i = i0 + e * env.W;
      if (i < (1 << m)) {
        if (out) out[dst[e]] = v[e];
        if (probs) env.atomic_add(probs + dst[e], qsb_norm2(v[e]));
      }
    }
  }
}

// Exchange k <= 3 cluster-rank bits g_j with local slot bits l_j in ONE pass (d->b[j] = g_j, d->cls[j] = l_j, d->k = k).
// With r_j = bit g_j of my rank and x_j = bit l_j of a local index, new[base | x] = old of CTA (rank with g_j := x_j)
// at [base | r]: the 2^-k of the tile with x == r stays, the rest is pulled from the 2^k - 1 peers, so k bits cost
// (1 - 2^-k) tile volumes instead of k/2.  Element q = (xv - 1) * 2^(m-k) + base number, xv = x ^ r != 0; every CTA
// enumerates q the same way, so the positions a CTA overwrites in a round are the ones its peers pulled in that
// round and one cluster barrier between the two halves of a round is enough.
// Pending matrices travel with their qubits (control-warp bookkeeping), so nothing is flushed here.
template <class Env>
QSB_PASS void qsb_do_remap(Env& env, int m, const qsb_desc* d) {
  typedef typename Env::amp A;
  A* tile = env.tile();
  const int k = d->k;
  int gb[3], lb[3], sl[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) { gb[j] = j < k ? d->b[j] : 0; lb[j] = j < k ? d->cls[j] : 0; sl[j] = lb[j]; }
  // ascending local bits for the zero insertion
  if (k > 1 && sl[0] > sl[1]) { int t = sl[0]; sl[0] = sl[1]; sl[1] = t; }
  if (k > 2 && sl[1] > sl[2]) { int t = sl[1]; sl[1] = sl[2]; sl[2] = t; }
  if (k > 2 && sl[0] > sl[1]) { int t = sl[0]; sl[0] = sl[1]; sl[1] = t; }
  int rmask = 0;                                  // my rank bits laid onto the local bits
#pragma unroll
  for (int j = 0; j < 3; ++j) if (j < k) rmask |= ((env.rank >> gb[j]) & 1) << lb[j];
  const int sh = m - k, cnt = 1 << sh;
  const int total = ((1 << k) - 1) << sh;
  const unsigned long long pt0 = env.prof_on() ? env.clock() : 0;
  env.cluster_sync_w();                       // every CTA finished the sweeps before the exchange
  if (env.prof_on()) env.prof_add(120, env.clock() - pt0);
#if QSB_REMAP_HALVES
  // Swap by halves: the element (xv, g) of CTA r and the element (xv, g) of its peer r ^ d(xv) are the two ends of
  // one swap, so ONE of the two CTAs does it -- remote load + remote store -- the lower rank for the lower half of
  // the group numbers g, the higher rank for the upper half (a tile with a single group: the lower rank does all).
  // Per CTA half the volume flows in and half flows out, which the DSMEM fabric moves concurrently, and there is no
  // barrier between the loads and the stores (measured in isolation: 4.5-5.0 k cycles per pair against 6.1-6.4 k).
  {
    const int shh = sh >= 1 ? sh - 1 : 0;
    const int E = sh >= 1 ? total >> 1 : total;
    constexpr int R = QSB_REMAP_REGS / 2;
    for (int base = 0; base < E; base += R * env.W) {
      const unsigned long long q0 = env.prof_on() ? env.clock() : 0;
      const bool tiny = E - base < env.W;
      A vl[R], vr[R];
      int mine[R], theirs[R], peer_rank[R];
#pragma unroll
      for (int e = 0; e < R; ++e) {
        int q = base + e * env.W + env.wid;
        const bool in_range = q < E;
        // out of range: the arrays are still written unconditionally (see above) -- with this worker's own first
        // element of the round (nobody else stores there), or the last element when the round is shorter than W
        q = in_range ? q : (tiny ? E - 1 : base + env.wid);
        const int xv = 1 + (q >> shh);
        int dr = 0, dl = 0;                           // xv spread onto the rank bits / the local bits
#pragma unroll
        for (int j = 0; j < 3; ++j) if (j < k) { dr |= ((xv >> j) & 1) << gb[j]; dl |= ((xv >> j) & 1) << lb[j]; }
        const int pr = env.rank ^ dr;
        const int upper = env.rank > pr ? 1 : 0;
        int bs = sh >= 1 ? ((q & ((1 << shh) - 1)) | (upper << shh)) : 0;
        bs = qsb_ins0(bs, sl[0]);
        if (k > 1) bs = qsb_ins0(bs, sl[1]);
        if (k > 2) bs = qsb_ins0(bs, sl[2]);
        mine[e] = QSB_SLOT(bs | (rmask ^ dl));
        theirs[e] = QSB_SLOT(bs | rmask);
        peer_rank[e] = (in_range && (sh >= 1 || !upper)) ? pr : -1;
        vl[e] = tile[mine[e]];
        vr[e] = env.peer_tile(pr)[theirs[e]];
      }
      const unsigned long long q1 = env.prof_on() ? env.clock() : 0;
      if (tiny) env.sync_workers();   // the discarded loads read a slot another worker is about to store
      const unsigned long long q2 = env.prof_on() ? env.clock() : 0;
#pragma unroll
      for (int e = 0; e < R; ++e) {
        if (peer_rank[e] >= 0) {
          tile[mine[e]] = vr[e];
          env.peer_tile_w(peer_rank[e])[theirs[e]] = vl[e];
        }
      }
      if (env.prof_on()) { env.prof_add(121, q1 - q0); env.prof_add(122, q2 - q1); env.prof_add(123, env.clock() - q2); env.prof_add(124 + k, 1); }
    }
    env.fence_cluster();                              // this worker's remote stores are performed at cluster scope ...
    env.cluster_sync_w();                             // ... and nobody sweeps before the peers' stores have landed
  }
#else
  A val[QSB_REMAP_REGS];
  for (int base = 0; base < total; base += QSB_REMAP_REGS * env.W) {
    const unsigned long long q0 = env.prof_on() ? env.clock() : 0;
    int mine[QSB_REMAP_REGS];
#pragma unroll
    for (int e = 0; e < QSB_REMAP_REGS; ++e) {
      // unconditional (index clamped): a conditionally written val[] is demoted to local memory and the
      // remote loads then run one at a time
      int q = base + e * env.W + env.wid;
      q = q < total ? q : total - 1;
      const int xv = 1 + (q >> sh);
      int bs = q & (cnt - 1);
      bs = qsb_ins0(bs, sl[0]);
      if (k > 1) bs = qsb_ins0(bs, sl[1]);
      if (k > 2) bs = qsb_ins0(bs, sl[2]);
      int dr = 0, dl = 0;                        // xv spread onto the rank bits / the local bits
#pragma unroll
      for (int j = 0; j < 3; ++j) if (j < k) { dr |= ((xv >> j) & 1) << gb[j]; dl |= ((xv >> j) & 1) << lb[j]; }
      mine[e] = bs | (rmask ^ dl);
      val[e] = env.peer_tile(env.rank ^ dr)[QSB_SLOT(bs | rmask)];
    }
    const unsigned long long q1 = env.prof_on() ? env.clock() : 0;
    env.cluster_sync_w();
    const unsigned long long q2 = env.prof_on() ? env.clock() : 0;
#pragma unroll
    for (int e = 0; e < QSB_REMAP_REGS; ++e) {
      const int q = base + e * env.W + env.wid;
      if (q < total) tile[QSB_SLOT(mine[e])] = val[e];
    }
    if (env.prof_on()) { env.prof_add(121, q1 - q0); env.prof_add(122, q2 - q1); env.prof_add(123, env.clock() - q2); env.prof_add(124 + k, 1); }
  }
#endif
}

// apply the pending 2x2 of cluster-rank bit gb to the pair (a in the CTA with the bit clear, b in its partner):
// a' = P00 a + P01 b, b' = P10 a + P11 b.
#ifndef QSB_GFLUSH_HALVES
#define QSB_GFLUSH_HALVES 1
#endif
template <class Env>
QSB_PASS void qsb_do_gflush(Env& env, int m, const qsb_desc* d) {
  typedef typename Env::amp A;
  A* tile = env.tile();
  const int gb = d->b[0];
  const int mybit = (env.rank >> gb) & 1;
  const int cnt = 1 << m;
#if QSB_GFLUSH_HALVES
  // Each CTA of the pair updates BOTH sides of half of the slots: one remote load and one remote store per slot
  // (64 KiB in + 64 KiB out per CTA, which the DSMEM fabric moves concurrently) instead of 128 KiB of remote loads,
  // and no barrier between the loads and the stores -- only before (the sweeps are done) and after (the remote
  // stores have landed) the pass.  Measured in isolation (tools/micro/xchg_bench.cu): 8.2-9.4 k cycles vs 10.6-11.2 k.
  A* peer = env.peer_tile_w(env.rank ^ (1 << gb));
  const A p00 = qsb_cvt<A>(d->P[0][0]), p01 = qsb_cvt<A>(d->P[0][1]), p10 = qsb_cvt<A>(d->P[0][2]), p11 = qsb_cvt<A>(d->P[0][3]);
  const int half = cnt >> 1, lo = mybit * half;                 // m >= 1
  constexpr int R = QSB_REMAP_REGS / 2;
  env.cluster_sync_w();
  for (int base = 0; base < half; base += R * env.W) {
    const bool tiny = half - base < env.W;
    A va[R], vb[R];
#pragma unroll
    for (int e = 0; e < R; ++e) {
      int i = base + e * env.W + env.wid;
      i = lo + (i < half ? i : (tiny ? half - 1 : base + env.wid));   // out of range: see qsb_do_remap
      const A mine = tile[i], theirs = peer[i];                 // same slot on both sides
      va[e] = mybit ? theirs : mine;                            // a lives where the rank bit is clear
      vb[e] = mybit ? mine : theirs;
    }
    if (tiny) env.sync_workers();     // the discarded loads read a slot another worker is about to store
#pragma unroll
    for (int e = 0; e < R; ++e) {
      const int j = base + e * env.W + env.wid;
      if (j < half) {
        const A na = qsb_fma(p01, vb[e], qsb_mul(p00, va[e])), nb = qsb_fma(p11, vb[e], qsb_mul(p10, va[e]));
        tile[lo + j] = mybit ? nb : na;
        peer[lo + j] = mybit ? na : nb;
      }
    }
  }
  env.fence_cluster();
  env.cluster_sync_w();                                          // nobody sweeps before the partner's stores landed
#else
  const A* peer = env.peer_tile(env.rank ^ (1 << gb));
  const A pm = qsb_cvt<A>(d->P[0][mybit * 2 + mybit]), po = qsb_cvt<A>(d->P[0][mybit * 2 + (1 - mybit)]);
  A val[QSB_REMAP_REGS];
  env.cluster_sync_w();
  for (int base = 0; base < cnt; base += QSB_REMAP_REGS * env.W) {
#pragma unroll
    for (int e = 0; e < QSB_REMAP_REGS; ++e) {
      int i = base + e * env.W + env.wid;
      i = i < cnt ? i : cnt - 1;                                            // clamped, see qsb_do_remap
      val[e] = qsb_fma(po, peer[i], qsb_mul(pm, tile[i]));                  // same slot on both sides
    }
    env.cluster_sync_w();
#pragma unroll
    for (int e = 0; e < QSB_REMAP_REGS; ++e) {
      int i = base + e * env.W + env.wid;
      if (i < cnt) tile[i] = val[e];
    }
  }
#endif
}

// worker main loop: consume descriptors until EXIT
template <class Env>
QSB_HD void qsb_worker_loop(Env& env, const qsb_exec_args& a) {
  const int m = a.m;
  int parity = 0, xpar = 0;
  const bool prof = Env::PROF && a.prof != nullptr && env.wid == 0;
  const bool wprof = Env::PROF && a.prof != nullptr && env.lane == 0;      // per-warp busy / wait cycles
  unsigned long long wb = 0, ww = 0, w0 = 0, w1 = 0;
  unsigned long long pw = 0, pb[8] = {0, 0, 0, 0, 0, 0, 0, 0}, pn[8] = {0, 0, 0, 0, 0, 0, 0, 0}, t0 = 0, t1 = 0;
  for (int slot = 0;; slot = slot + 1 == QSB_RING ? 0 : slot + 1) {
    if (prof) t0 = env.clock();
    if (wprof) w0 = env.clock();
    env.ring_wait_full(slot);               // also: every worker finished the previous descriptor
    if (wprof) { w1 = env.clock(); ww += w1 - w0; }
    if (prof) { t1 = env.clock(); pw += t1 - t0; }
    const qsb_desc* d = &env.ctl()->ring[slot];
    const qsb_desc_hdr hdr = *reinterpret_cast<const qsb_desc_hdr*>(d);      // kind, gate, k, flags: one 16-byte load
    qsb_sweep_pro pro;
    qsb_sweep_prologue(env, d, pro);          // only sweeps use it; for the other kinds these are loads of unused fields
    const int kind = hdr.kind;
    switch (kind) {
      case QSB_D_INIT: qsb_do_init(env, a, d); break;
      case QSB_D_SWEEP: qsb_do_sweep(env, m, d, hdr, pro); break;
      case QSB_D_REMAP: qsb_do_remap(env, m, d); break;
      case QSB_D_GFLUSH: qsb_do_gflush(env, m, d); break;
      case QSB_D_RDM1: {
        double v[4];
        if (d->flags & 1) qsb_partial_marginal_w(env, m, d->b[0], reinterpret_cast<const double*>(d->mat), v);
        else qsb_partial_rdm1(env, m, d->b[0], v);
        // CTA sums: one partial per warp; cluster sums: entry (r, k) of this CTA's sums is PUSHED into the buffer of
        // CTA r (remote store, then the workers' cluster barrier), so that after the barrier every CTA adds eight
        // local values in the same order -- the same bits everywhere, and no remote load on the critical path
        double* wp = env.ctl()->wpart;
        env.sync_workers();                 // previous users of wpart are done
        for (int k = 0; k < 4; ++k) {
          const double x = env.warp_sum(v[k]);
          if (env.lane == 0) wp[env.warp * 4 + k] = x;
        }
        env.sync_workers();
        if (env.C > 1) {
          for (int e = env.wid; e < 4 * env.C; e += env.W) {
            const int r = e >> 2, k = e & 3;
            double sum = 0.0;
            for (int w = 0; w < env.nwarps; ++w) sum += wp[w * 4 + k];
            env.peer_ctl_w(r)->redx[xpar][env.rank][k] = sum;
            env.fence_cluster();            // the remote store is ordered before the barrier's release-arrive of another thread
          }
          env.cluster_sync_w();
          if (env.wid < 4) {
            double sum = 0.0;
            for (int r = 0; r < env.C; ++r) sum += env.ctl()->redx[xpar][r][env.wid];
            env.ctl()->red_total[env.wid] = sum;
          }
          xpar ^= 1;
        } else if (env.wid < 4) {
          double sum = 0.0;
          for (int w = 0; w < env.nwarps; ++w) sum += wp[w * 4 + env.wid];
          env.ctl()->red_total[env.wid] = sum;
        }
        env.handoff_w();                    // the local control warp reads red_total after this barrier
        break;
      }
      case QSB_D_STORE: qsb_do_store(env, a, d, parity); break;
      default: break;
    }
    env.ring_release(slot);
    if (wprof) wb += env.clock() - w1;
    if (prof) {
      const unsigned long long dt = env.clock() - t1;
      pb[kind & 7] += dt; pn[kind & 7] += 1;
      if (kind == QSB_D_SWEEP) {          // per sweep variant (k * 8 + gate) and per number of dense pending matrices
        unsigned long long* o = a.prof + (size_t)env.cta_id() * QSB_PROF_WORDS;
        const int key = (d->k * 8 + d->gate) & 31;
        int nd = 0;
        for (int k = 0; k < d->k; ++k) nd += d->cls[k] == QSB_CLS_DENSE;
        o[32 + key] += dt; o[64 + key] += 1;
        o[96 + nd] += dt; o[100 + nd] += 1;
      }
    }
    if (kind == QSB_D_EXIT) break;
  }
  if (wprof && env.warp < 8) {
    unsigned long long* o = a.prof + (size_t)env.cta_id() * QSB_PROF_WORDS;
    o[104 + env.warp] = wb; o[112 + env.warp] = ww;
  }
  if (prof) {
    unsigned long long* o = a.prof + (size_t)env.cta_id() * QSB_PROF_WORDS;
    o[0] = pw;
    for (int k = 0; k < 8; ++k) { o[1 + k] = pb[k]; o[9 + k] = pn[k]; }
  }
  if (env.C > 1) env.cluster_exit();        // no CTA may exit while a peer can still read its shared memory
}

// =========================================================================================
//                                      CONTROL SIDE
// =========================================================================================
// Pending matrices are ALWAYS valid 2x2 matrices: a slot with nothing pending holds the exact identity, so
// folding an op is an unconditional P <- U P.  The class word (2 bits per slot) only tells the workers how
// much arithmetic a pending matrix needs.
struct qsb_cstate {
  uint32_t seq;                    // descriptors published so far
  uint32_t gchunk;                 // chunks consumed (control) / produced (decode warp) so far, over all units
  uint64_t clsword;                // 2-bit structure class per slot bit
  int parity;
  // per-unit constants
  int64_t t, tile, dim;
  bool record;
  bool prof;
  unsigned long long ring_wait;    // cycles blocked on a full ring
  unsigned long long t_decode, t_fold, t_slow, n_slow;   // control-warp cycles per phase (profiling only)
  unsigned long long t_emit_body, t_emit_pub, n_emit;
  unsigned long long t_e[4];
};

// class word: two bit planes, bit b of the low word = class bit 0 of slot b, bit b of the high word = class bit 1
QSB_HD int qsb_cls_of(uint64_t w, int b) { return (int)(((w >> b) & 1u) | (((w >> (32 + b)) & 1u) << 1)); }
QSB_HD uint64_t qsb_cls_set(uint64_t w, int b, int cls) {
  const uint64_t keep = ~(((uint64_t)1 << b) | ((uint64_t)1 << (32 + b)));
  return (w & keep) | ((uint64_t)(cls & 1) << b) | ((uint64_t)(cls >> 1) << (32 + b));
}
// slot bits (as a bit mask) whose class is not NONE / is DENSE
QSB_HD uint32_t qsb_cls_mask(uint64_t w) { return (uint32_t)w | (uint32_t)(w >> 32); }
QSB_HD uint32_t qsb_cls_dense_mask(uint64_t w) { return (uint32_t)w & (uint32_t)(w >> 32); }
QSB_HD void qsb_pend_identity(c128* P) { P[0] = qsb_c(1, 0); P[1] = qsb_c(0, 0); P[2] = qsb_c(0, 0); P[3] = qsb_c(1, 0); }
// P <- U P
QSB_HD void qsb_pend_apply(c128* P, const c128* U) {
  const c128 p00 = P[0], p01 = P[1], p10 = P[2], p11 = P[3];
  const c128 u00 = U[0], u01 = U[1], u10 = U[2], u11 = U[3];
  P[0] = qsb_fma(u01, p10, qsb_mul(u00, p00));
  P[1] = qsb_fma(u01, p11, qsb_mul(u00, p01));
  P[2] = qsb_fma(u11, p10, qsb_mul(u10, p00));
  P[3] = qsb_fma(u11, p11, qsb_mul(u10, p01));
}

template <class Env>
QSB_HD qsb_desc* qsb_desc_begin(Env& env, qsb_cstate& st) {
  const int slot = (int)(st.seq % QSB_RING);
  if (st.seq >= QSB_RING) {
    const unsigned long long t0 = st.prof ? env.clock() : 0;
    env.ring_wait_empty(slot);
    if (st.prof) st.ring_wait += env.clock() - t0;
  }
  return &env.ctl()->ring[slot];
}
template <class Env>
QSB_HD void qsb_desc_end(Env& env, qsb_cstate& st) {
  env.ring_publish((int)(st.seq % QSB_RING));
  ++st.seq;
}
// load / store offsets and real factors of a sweep's local indices (qsb_desc::off, ost, f): lane r works out entry r.
// d->P and the class word `w` (before this sweep's reset) describe the pending matrices of the sweep's bits.
template <class Env>
QSB_CTL void qsb_sweep_tables(Env& env, qsb_desc* d, uint64_t w, int gate, int nb, int b0, int b1, int b2) {
  const int D = 1 << nb;
  for (int r = env.clane; r < D; r += env.CL) {
    int pr = r;                     // the gate moves the amplitude of local index r to local index pr
    if (nb == 2) {
      if (gate == QSB_G_CX) pr = (r & 2) ? (r ^ 1) : r;                          // b[0] control, b[1] target
      else if (gate == QSB_G_SWAP) pr = ((r & 1) << 1) | (r >> 1);
    } else if (nb == 3) {
      if (gate == QSB_G_CCX) pr = ((r & 6) == 6) ? (r ^ 1) : r;                  // 110 <-> 111
      else if (gate == QSB_G_CSWAP) pr = (r & 4) ? (4 | ((r & 1) << 1) | ((r >> 1) & 1)) : r;   // 101 <-> 110
    }
    int o = 0, os = 0;
    double fr = 1.0;
    for (int k = 0; k < nb; ++k) {
      const int b = k == 0 ? b0 : (k == 1 ? b1 : b2);
      if ((r >> (nb - 1 - k)) & 1) {
        o |= 1 << b;
        if (qsb_cls_of(w, b) == QSB_CLS_RDIAG) fr *= d->P[k][3].x;
      }
      if ((pr >> (nb - 1 - k)) & 1) os |= 1 << b;
    }
    if (nb == 2 && r == 3 && gate == QSB_G_CZ) fr = -fr;
    d->off[r] = QSB_SLOT(o); d->ost[r] = QSB_SLOT(os); d->f[r] = fr;
  }
}

// publish one sweep over `nb` local bits (bits[0] = MSB of the gate index) and reset their pending matrices.
// The lanes of the control warp share the copies: entry e of bit k goes through lane 4k + e.
template <class Env>
QSB_CTL void qsb_emit_sweep(Env& env, qsb_cstate& st, int m, int gate, int nb, int b0, int b1, int b2, const c128* mat_src,
                            const qsb_dec* order = nullptr) {
  qsb_desc* d = qsb_desc_begin(env, st);
  qsb_ctl* ctl = env.ctl();
  const unsigned long long pe0 = st.prof ? env.clock() : 0;
  const uint64_t w = st.clsword;
  env.sync_control();                          // the fold's pending matrices are visible to every lane
  const unsigned long long pq0 = st.prof ? env.clock() : 0;
  for (int e = env.clane; e < 4 * nb; e += env.CL) {
    const int k = e >> 2, j = e & 3;
    const int b = k == 0 ? b0 : (k == 1 ? b1 : b2);
    d->P[k][j] = ctl->pend[b][j];
    ctl->pend[b][j] = qsb_c((j == 0 || j == 3) ? 1.0 : 0.0, 0.0);
  }
  if (gate != QSB_G_DENSE) {
    env.sync_control();                        // d->P is complete
    qsb_sweep_tables(env, d, w, gate, nb, b0, b1, b2);
  }
  const unsigned long long pq1 = st.prof ? env.clock() : 0;
  uint32_t used = 1u << b0;
  if (nb > 1) used |= 1u << b1;
  if (nb > 2) used |= 1u << b2;
  int hm;
  uint64_t pos;
  if (order) { pos = qsb_double_as_u64(order->U[0].x); hm = (int)order->U[0].y; }
  else pos = qsb_group_order_lanes(env, m, used, env.wbits, m - nb, (int)qsb_amp<typename Env::amp>::SW, &hm);
  const unsigned long long pq2 = st.prof ? env.clock() : 0;
  if (st.prof) { st.t_e[0] += pq0 - pe0; st.t_e[1] += pq1 - pq0; st.t_e[2] += pq2 - pq1; }
  if (env.lead) {
    d->b[0] = b0; d->b[1] = b1; d->b[2] = b2;
    d->cls[0] = qsb_cls_of(w, b0); d->cls[1] = nb > 1 ? qsb_cls_of(w, b1) : 0; d->cls[2] = nb > 2 ? qsb_cls_of(w, b2) : 0;
    // flags: which bits carry a full 2x2 -- the workers pick the sweep variant from the descriptor's first 16 bytes
    d->kind = QSB_D_SWEEP; d->gate = gate; d->k = nb;
    d->flags = (d->cls[0] >= QSB_CLS_DIAG ? 1 : 0) | (d->cls[1] >= QSB_CLS_DIAG ? 2 : 0) | (d->cls[2] >= QSB_CLS_DIAG ? 4 : 0);
    d->pos = pos;
    d->hmask = hm;
  }
  {
    const int nbits = env.wbits < m - nb ? env.wbits : m - nb;
    for (int e = env.clane; e < 32; e += env.CL) {
      d->tabl[e] = qsb_deposit(e, pos, nbits < 5 ? nbits : 5);
      if (e < 8) d->tabw[e] = qsb_deposit(e << 5, pos, nbits);        // bits 0..4 of e << 5 are clear
    }
  }
  if (mat_src) {
    const int cnt = 1 << (2 * nb);
    for (int e = env.clane; e < cnt; e += env.CL) d->mat[e] = mat_src[e];
  }
  uint64_t nw = qsb_cls_set(w, b0, QSB_CLS_NONE);
  if (nb > 1) nw = qsb_cls_set(nw, b1, QSB_CLS_NONE);
  if (nb > 2) nw = qsb_cls_set(nw, b2, QSB_CLS_NONE);
  st.clsword = nw;
  const unsigned long long pe1 = st.prof ? env.clock() : 0;
  qsb_desc_end(env, st);
  if (st.prof) { st.t_emit_body += pe1 - pe0; st.t_emit_pub += env.clock() - pe1; st.n_emit += 1; }
}

// apply and reset every pending matrix in `which` (slot-bit mask; local bits three per sweep, rank bits
// one cluster exchange each)
template <class Env>
QSB_CTL void qsb_flush(Env& env, qsb_cstate& st, int m, uint32_t which) {
  const uint32_t pending = qsb_cls_mask(st.clsword) & which;
  uint32_t todo = pending & ((1u << m) - 1u);
  while (todo) {
    int bits[3], nb = 0;
    while (todo && nb < 3) {
      const int b = QSB_MSB(todo);
      bits[nb++] = b;
      todo &= ~(1u << b);
    }
    qsb_emit_sweep(env, st, m, QSB_G_NONE, nb, bits[0], nb > 1 ? bits[1] : 0, nb > 2 ? bits[2] : 0, (const c128*)nullptr);
  }
  uint32_t gtodo = m >= 32 ? 0u : (pending >> m);
  for (int gb = 0; gtodo; ++gb, gtodo >>= 1) {
    if (!(gtodo & 1u)) continue;
    qsb_desc* d = qsb_desc_begin(env, st);
    if (env.lead) {
      d->kind = QSB_D_GFLUSH; d->k = 1; d->b[0] = gb; d->cls[0] = QSB_CLS_DENSE;
      for (int e = 0; e < 4; ++e) d->P[0][e] = env.ctl()->pend[m + gb][e];
      qsb_pend_identity(env.ctl()->pend[m + gb]);
    }
    st.clsword = qsb_cls_set(st.clsword, m + gb, QSB_CLS_NONE);
    qsb_desc_end(env, st);
  }
}

// (unnormalised) 1-qubit reduced density matrix sums of slot bit b of the CURRENT state, in every control warp
template <class Env>
QSB_CTL void qsb_ctl_rdm1(Env& env, qsb_cstate& st, int b, double v[4], int weighted = 0, int n = 0, double* clu = nullptr) {
  qsb_desc* d = qsb_desc_begin(env, st);
  if (env.lead) { d->kind = QSB_D_RDM1; d->k = 1; d->b[0] = b; d->flags = weighted ? 1 : 0; }
  if (weighted == 2) {
    // weights = diag(P^H P) of every other slot (ctl->slotw, filled by the caller), (1, 1) for b itself
    env.sync_control();
    double* wt = reinterpret_cast<double*>(d->mat);
    for (int e = env.clane; e < 64; e += env.CL) wt[e] = (e >> 1) == b ? 1.0 : env.ctl()->slotw[e >> 1][e & 1];
  } else if (weighted) {
    // weights |P[v][v]|^2 of the (diagonal) pending matrices that stay pending; identity where nothing is pending
    env.sync_control();
    double* wt = reinterpret_cast<double*>(d->mat);
    for (int e = env.clane; e < 64; e += env.CL) {
      const c128 z = env.ctl()->pend[e >> 1][(e & 1) ? 3 : 0];
      wt[e] = z.x * z.x + z.y * z.y;
    }
  }
  qsb_desc_end(env, st);
  if (clu) {             // while the workers read the tile: prod(1 -+ rho_s) of the bounded amplitude-damping draw
    double cl = 1.0, cu = 1.0;
    for (int slot = 0; slot < n; ++slot) { const double rho = env.ctl()->slotw[slot][2]; cl *= 1.0 - rho; cu *= 1.0 + rho; }
    clu[0] = cl; clu[1] = cu;
  }
  env.handoff_c();
  for (int k = 0; k < 4; ++k) v[k] = env.ctl()->red_total[k];
}

template <class Env>
QSB_CTL void qsb_emit_store(Env& env, qsb_cstate& st, const qsb_exec_args& a, const int32_t* perm,
                            void* out, double* probs) {
  qsb_desc* d = qsb_desc_begin(env, st);
  if (env.lead) {
    d->kind = QSB_D_STORE; d->flags = a.flags & QSB_RUN_NORMALIZE;
    d->perm = perm; d->gptr = out; d->probs = probs; d->tile = st.tile;
  }
  qsb_desc_end(env, st);
}

// ---- lane-parallel decode of one op (any control lane) ------------------------------------------
// 1-qubit gates, Pauli draws and "certain K0" amplitude-damping draws become MUL records; the rest is SLOW.
template <class Env>
QSB_HD void qsb_decode_op(Env& env, const qsb_exec_args& a, const qsb_cstate& st, const qsb_op& op, double u,
                          const double* prm, qsb_dec* d) {
  const double* cd = a.cdata + (op.data >= 0 ? op.data : 0);
  c128 U[4];
  int type = QSB_DEC_MUL, ucls = QSB_CLS_DENSE;
  U[0] = qsb_c(1, 0); U[1] = qsb_c(0, 0); U[2] = qsb_c(0, 0); U[3] = qsb_c(1, 0);
  switch (op.kind) {
    case QSB_OP_NOP: type = QSB_DEC_SKIP; break;
    case QSB_OP_U1:
      U[0] = qsb_c(cd[0], cd[1]); U[1] = qsb_c(cd[2], cd[3]); U[2] = qsb_c(cd[4], cd[5]); U[3] = qsb_c(cd[6], cd[7]);
      break;
    case QSB_OP_D1:
      U[0] = qsb_c(cd[0], cd[1]); U[3] = qsb_c(cd[2], cd[3]);
      ucls = (U[0].x == 1.0 && U[0].y == 0.0 && U[3].y == 0.0) ? QSB_CLS_RDIAG : QSB_CLS_DIAG;
      break;
    case QSB_OP_X: U[0] = qsb_c(0, 0); U[1] = qsb_c(1, 0); U[2] = qsb_c(1, 0); U[3] = qsb_c(0, 0); break;
    case QSB_OP_Y: U[0] = qsb_c(0, 0); U[1] = qsb_c(0, -1); U[2] = qsb_c(0, 1); U[3] = qsb_c(0, 0); break;
    case QSB_OP_Z: U[3] = qsb_c(-1, 0); ucls = QSB_CLS_RDIAG; break;
    case QSB_OP_RX: case QSB_OP_RY: {     // gates.py:66-75
      double s, c;
      sincos(prm[op.param] * 0.5, &s, &c);
      U[0] = qsb_c(c, 0); U[3] = qsb_c(c, 0);
      if (op.kind == QSB_OP_RX) { U[1] = qsb_c(0, -s); U[2] = qsb_c(0, -s); }
      else { U[1] = qsb_c(-s, 0); U[2] = qsb_c(s, 0); }
      break;
    }
    case QSB_OP_RZ: {                     // gates.py:78-80
      double s, c;
      sincos(prm[op.param] * 0.5, &s, &c);
      U[0] = qsb_c(c, -s); U[3] = qsb_c(c, s); ucls = QSB_CLS_DIAG;
      break;
    }
    case QSB_OP_PHASE: {                  // gates.py:83-85
      double s, c;
      sincos(prm[op.param], &s, &c);
      U[3] = qsb_c(c, s); ucls = QSB_CLS_DIAG;
      break;
    }
    case QSB_OP_U3: {                     // gates.py:88-94
      double s, c, sp, cp, sl, cl, spl, cpl;
      sincos(prm[op.param] * 0.5, &s, &c);
      sincos(prm[op.param + 1], &sp, &cp);
      sincos(prm[op.param + 2], &sl, &cl);
      sincos(prm[op.param + 1] + prm[op.param + 2], &spl, &cpl);
      U[0] = qsb_c(c, 0); U[1] = qsb_c(-cl * s, -sl * s); U[2] = qsb_c(cp * s, sp * s); U[3] = qsb_c(cpl * c, spl * c);
      break;
    }
    case QSB_OP_KRAUS_PAULI: {
      // K_i = sqrt(w_i) P_i: ||K_i psi||^2 / sum = w_i / sum(w) for any psi, so the branch needs no
      // reduction and (with the norm deferred to the store) the update is the bare Pauli.
      int idx = 0;
      for (int k = 0; k < 3; ++k) if (cd[k] <= u) ++idx;
      const int code = (int)cd[3 + idx];
      if (st.record) a.branches[st.t * a.branches_stride + op.draw] = idx;
      if (code == 0) type = QSB_DEC_SKIP;
      else if (code == 1) { U[0] = qsb_c(0, 0); U[1] = qsb_c(1, 0); U[2] = qsb_c(1, 0); U[3] = qsb_c(0, 0); }
      else if (code == 2) { U[0] = qsb_c(0, 0); U[1] = qsb_c(0, -1); U[2] = qsb_c(0, 1); U[3] = qsb_c(0, 0); }
      else { U[3] = qsb_c(-1, 0); ucls = QSB_CLS_RDIAG; }
      break;
    }
    case QSB_OP_KRAUS_AD: {
      // K0 = diag(1, sqrt(1-g)), K1 = sqrt(g)|0><1| (noise.py:98-103).  cdf[0] = p0/(p0+p1) >= 1-g, so a
      // draw below 1-g is K0 whatever the state; only the rest needs P(q=1) of the actual state.
      if (u < 1.0 - cd[0] - QSB_AD_MARGIN) {
        if (st.record) a.branches[st.t * a.branches_stride + op.draw] = 0;
        U[3] = qsb_c(cd[1], 0); ucls = QSB_CLS_RDIAG;
      } else type = QSB_DEC_SLOW;
      break;
    }
    case QSB_OP_CX: case QSB_OP_CZ: case QSB_OP_SWAP: case QSB_OP_U2:
    case QSB_OP_CCX: case QSB_OP_CSWAP: case QSB_OP_U3Q: {
      // group order of the sweep (depends on the target bits only): ~100 dependent instructions that the serial part
      // of the control warp would otherwise run once per sweep
      type = QSB_DEC_SLOW;
      const int nb = (op.kind == QSB_OP_CCX || op.kind == QSB_OP_CSWAP || op.kind == QSB_OP_U3Q) ? 3 : 2;
      uint32_t used = (1u << op.b0) | (1u << op.b1);
      if (nb > 2) used |= 1u << op.b2;
      int hm;
      const uint64_t pos = qsb_group_order(a.m, used, env.wbits, a.m - nb, (int)qsb_amp<typename Env::amp>::SW, &hm);
      U[0].x = qsb_u64_as_double(pos);
      U[0].y = (double)hm;
      d->U[0] = U[0];
      break;
    }
    default: type = QSB_DEC_SLOW; break;
  }
  d->type = type; d->b = op.b0; d->ucls = ucls;
  if (type == QSB_DEC_MUL)
    for (int e = 0; e < 4; ++e) d->U[e] = U[e];
}

// ---- ops that need the descriptor ring or the state (whole control warp) --------------------------
template <class Env>
QSB_CTL void qsb_control_slow(Env& env, const qsb_exec_args& a, qsb_cstate& st, qsb_chunk* ck, int i) {
  qsb_ctl* ctl = env.ctl();
  const int m = a.m, n = a.n;
  const qsb_op op = ck->ops[i];
  const uint32_t all_bits = n >= 32 ? 0xffffffffu : ((1u << n) - 1u);
  const int b = op.b0;
  switch (op.kind) {
    case QSB_OP_KRAUS_AD: {               // the draw fell in [1-g, 1): the branch depends on P(q=1)
      const double* cd = a.cdata + op.data;
      const double u = ck->u[i], gam = cd[0];
      double v[4];
      int idx = -1;
#if QSB_AD_BOUNDS
      // Bounded draw: nothing is flushed.  With M_s = P_s^H P_s of the pending matrix of every other slot s,
      // (1 - rho_s) D_s <= M_s <= (1 + rho_s) D_s  (D_s = diag M_s, rho_s = |M_s01| / sqrt(M_s00 M_s11), Loewner order),
      // so the weighted marginals N_v taken with the DIAGONAL weights D_s bracket the true ones between
      // prod(1 - rho_s) N_v and prod(1 + rho_s) N_v.  The pending matrices are unitaries times a few damping K0's:
      // rho_s ~ gamma / 2, the bracket on P(q = 1) is a few per cent wide and decides all but ~2 % of the draws with
      // ONE read of the tile; the rest take the exact path below.  Slots that are far from diagonal (rho > 0.1: a K1
      // branch that no sweep has absorbed yet) are flushed on their own first, and so is b when it is a cluster-rank
      // slot whose pending matrix mixes the two sides (the off-diagonal of its marginal lives in two CTAs).
      {
        uint32_t far = 0;
        env.sync_control();
        for (int slot = env.clane; slot < 32; slot += env.CL) {
          double sa = 1.0, sd = 1.0, rho = 0.0;
          int isfar = 0;
          if (slot < n) {
            const c128* P = ctl->pend[slot];
            if (slot == b) {
              isfar = slot >= m && (qsb_norm2(P[0]) * qsb_norm2(P[1]) + qsb_norm2(P[2]) * qsb_norm2(P[3]) > 0.0);
            } else {
              sa = qsb_norm2(P[0]) + qsb_norm2(P[2]); sd = qsb_norm2(P[1]) + qsb_norm2(P[3]);
              const double ex = P[0].x * P[1].x + P[0].y * P[1].y + P[2].x * P[3].x + P[2].y * P[3].y;
              const double ey = P[0].x * P[1].y - P[0].y * P[1].x + P[2].x * P[3].y - P[2].y * P[3].x;
              const double e2 = ex * ex + ey * ey;
              if (e2 > 0.0) {
                rho = sa * sd > 0.0 ? sqrt(e2 / (sa * sd)) : 1.0;
                if (rho > 1.0) rho = 1.0;
                isfar = rho > 0.1;
              }
            }
          }
          ctl->slotw[slot][0] = sa; ctl->slotw[slot][1] = sd; ctl->slotw[slot][2] = rho;
          far |= env.ballot_slot(isfar, slot);
        }
        env.sync_control();
        if (far) {
          qsb_flush(env, st, m, far);                 // their pending matrices are the identity now
          for (int slot = env.clane; slot < 32; slot += env.CL)
            if ((far >> slot) & 1u) { ctl->slotw[slot][0] = 1.0; ctl->slotw[slot][1] = 1.0; ctl->slotw[slot][2] = 0.0; }
          env.sync_control();
        }
        double clu[2];
        qsb_ctl_rdm1(env, st, b, v, 2, n, clu);
        const double cl = clu[0], cu = clu[1];
        const c128* P = ctl->pend[b];
        double nv[2];
        for (int r = 0; r < 2; ++r) {
          const c128 p0 = P[2 * r], p1 = P[2 * r + 1];      // N_r = sum_ij P[r][i] conj(P[r][j]) R_ij
          nv[r] = qsb_norm2(p0) * v[0] + qsb_norm2(p1) * v[1] +
                  2.0 * ((p0.x * p1.x + p0.y * p1.y) * v[2] - (p0.y * p1.x - p0.x * p1.y) * v[3]);
          if (nv[r] < 0.0) nv[r] = 0.0;
        }
        const double dlo = cl * nv[1] + cu * nv[0], dhi = cu * nv[1] + cl * nv[0];
        if (dlo > 0.0 && dhi > 0.0) {
          const double lo = cl * nv[1] / dlo, hi = cu * nv[1] / dhi;      // lo <= P(q = 1) <= hi
          if (u < 1.0 - gam * hi - QSB_AD_MARGIN) idx = 0;                 // cdf[0] = 1 - gamma P(q = 1)
          else if (u > 1.0 - gam * lo + QSB_AD_MARGIN) idx = 1;
        }
      }
#endif
      if (idx < 0) {
        // exact draw.  Only DENSE pending matrices have to be applied before the marginal is taken: diagonal ones (the
        // K0's of earlier draws, Rz / phase gates) just reweight |amplitude|^2 and ride along as weights
        qsb_flush(env, st, m, qsb_cls_dense_mask(st.clsword));
        qsb_ctl_rdm1(env, st, b, v, 1);
        double p[2] = {v[0] + (1.0 - gam) * v[1], gam * v[1]};
        idx = qsb_choice(p, 2, u);
      }
      if (st.record) a.branches[st.t * a.branches_stride + op.draw] = idx;
      const int cls = qsb_cls_of(st.clsword, b);
      if (env.lead) {
        c128 K[4];
        if (idx == 0) { K[0] = qsb_c(1, 0); K[1] = qsb_c(0, 0); K[2] = qsb_c(0, 0); K[3] = qsb_c(cd[1], 0); }
        else { K[0] = qsb_c(0, 0); K[1] = qsb_c(1, 0); K[2] = qsb_c(0, 0); K[3] = qsb_c(0, 0); }   // a0 <- a1, a1 <- 0
        qsb_pend_apply(ctl->pend[b], K);
      }
      const int kcls = idx == 0 ? QSB_CLS_RDIAG : QSB_CLS_DENSE;
      st.clsword = qsb_cls_set(st.clsword, b, cls > kcls ? cls : kcls);
      break;
    }
    case QSB_OP_KRAUS_GEN: {              // target is a local bit (host compiler remaps it in)
      const double u = ck->u[i];
      const double* g = a.cdata + op.data;
      const int nk = (int)g[0];
      double v[4], p[8];
      qsb_flush(env, st, m, all_bits);
      qsb_ctl_rdm1(env, st, b, v);
      for (int k = 0; k < nk; ++k) {
        const double* e = g + 1 + k * 12 + 8;       // e00, e11, re e01, im e01 ; rho10 = conj(v2 + i v3)
        p[k] = e[0] * v[0] + e[1] * v[1] + 2.0 * (e[2] * v[2] + e[3] * v[3]);
      }
      const int idx = qsb_choice(p, nk, u);
      if (st.record) a.branches[st.t * a.branches_stride + op.draw] = idx;
      const double* kk = g + 1 + idx * 12;
      if (env.lead) {
        c128* P = ctl->pend[b];
        P[0] = qsb_c(kk[0], kk[1]); P[1] = qsb_c(kk[2], kk[3]); P[2] = qsb_c(kk[4], kk[5]); P[3] = qsb_c(kk[6], kk[7]);
      }
      st.clsword = qsb_cls_set(st.clsword, b, QSB_CLS_DENSE);
      break;
    }
    // ---------------- multi-qubit gates: one sweep = pending matrices of the bits + the gate
    case QSB_OP_CX: case QSB_OP_CZ: case QSB_OP_SWAP: case QSB_OP_U2: {
      const int g = op.kind == QSB_OP_CX ? QSB_G_CX : op.kind == QSB_OP_CZ ? QSB_G_CZ :
                    op.kind == QSB_OP_SWAP ? QSB_G_SWAP : QSB_G_DENSE;
      qsb_emit_sweep(env, st, m, g, 2, op.b0, op.b1, 0, op.kind == QSB_OP_U2 ? (const c128*)(a.cdata + op.data) : (const c128*)nullptr,
                     &ck->dec[i]);
      break;
    }
    case QSB_OP_CCX: case QSB_OP_CSWAP: case QSB_OP_U3Q: {
      const int g = op.kind == QSB_OP_CCX ? QSB_G_CCX : op.kind == QSB_OP_CSWAP ? QSB_G_CSWAP : QSB_G_DENSE;
      qsb_emit_sweep(env, st, m, g, 3, op.b0, op.b1, op.b2, op.kind == QSB_OP_U3Q ? (const c128*)(a.cdata + op.data) : (const c128*)nullptr,
                     &ck->dec[i]);
      break;
    }
    case QSB_OP_REMAP: {
      // swap cluster-rank bit b0 with local slot bit b1 (and, in the same pass, the b2 further pairs packed in aux:
      // g1 | l1 << 8 | g2 << 16 | l2 << 24); the pending matrices move with their qubits
      const int k = 1 + op.b2;
      int gbs[3] = {op.b0, op.aux & 255, (op.aux >> 16) & 255}, lbs[3] = {op.b1, (op.aux >> 8) & 255, (op.aux >> 24) & 255};
      qsb_desc* d = qsb_desc_begin(env, st);
      if (env.lead) {
        d->kind = QSB_D_REMAP; d->k = k;
        for (int j = 0; j < k; ++j) { d->b[j] = gbs[j]; d->cls[j] = lbs[j]; }
      }
      qsb_desc_end(env, st);
      for (int j = 0; j < k; ++j) {
        const int gb = gbs[j], lb = lbs[j];
        const int ca = qsb_cls_of(st.clsword, lb), cb = qsb_cls_of(st.clsword, m + gb);
        if (env.lead)
          for (int e = 0; e < 4; ++e) { c128 x = ctl->pend[lb][e]; ctl->pend[lb][e] = ctl->pend[m + gb][e]; ctl->pend[m + gb][e] = x; }
        st.clsword = qsb_cls_set(qsb_cls_set(st.clsword, lb, cb), m + gb, ca);
      }
      break;
    }
    case QSB_OP_SNAPSHOT: {
      qsb_flush(env, st, m, all_bits);
      qsb_emit_store(env, st, a, a.idata + op.aux,
                     a.snapshots ? static_cast<char*>(a.snapshots) + (st.t * a.n_snapshots + op.b0) * st.dim * a.amp_bytes : nullptr, nullptr);
      break;
    }
    default:
      break;
  }
}

// Per-slot lists over the decoded chunk: head[slot] = first MUL op of the slot, dec[i].next = next one (QSB_CHUNK =
// none).  Device: 32 ops per step, __match_any_sync groups the lanes by slot; the windows are walked from the
// last to the first so that each op can point at the first op of its slot in the later windows.
template <class Env>
QSB_HD void qsb_link_chunk(Env& env, qsb_chunk* ctl, int len) {
  env.sync_control();
  for (int slot = env.clane; slot < 32; slot += env.CL) ctl->head[slot] = QSB_CHUNK;
  env.sync_control();
  if (Env::CL == 1) {                                   // one control thread (host emulation): plain backward scan
    for (int j = len - 1; j >= 0; --j) {
      qsb_dec* d = &ctl->dec[j];
      if (d->type != QSB_DEC_MUL) continue;
      d->next = ctl->head[d->b & 31];
      ctl->head[d->b & 31] = j;
    }
    return;
  }
  for (int base = ((len - 1) / 32) * 32; base >= 0; base -= 32) {
    const int j = base + env.clane;
    const bool mul = j < len && ctl->dec[j].type == QSB_DEC_MUL;
    const int slot = mul ? (ctl->dec[j].b & 31) : 0;
    // lanes of the same slot; ops that are not MUL get a key of their own
    const uint32_t grp = env.match_any(mul ? slot : 32 + env.clane);
    const uint32_t above = grp & ~((2u << env.clane) - 1u);
    if (mul) ctl->dec[j].next = above ? base + QSB_CTZ(above) : ctl->head[slot];
    env.sync_control();                                 // every lane read head[] before the group leaders update it
    if (mul && QSB_CTZ(grp) == env.clane) ctl->head[slot] = j;
    env.sync_control();
  }
}

// ---- one unit (trajectory, or tile of a streamed state) on the control warp ----------------
template <class Env>
QSB_HD void qsb_control_unit(Env& env, qsb_cstate& st, const qsb_exec_args& a, int64_t unit) {
  const int n = a.n;
  qsb_ctl* ctl = env.ctl();
  const int64_t t = unit >> a.tile_bits;                          // state / trajectory index
  st.t = t;
  st.tile = a.tile_bits ? (unit & (((int64_t)1 << a.tile_bits) - 1)) : (int64_t)env.rank;
  st.dim = (int64_t)1 << n;
  st.record = a.branches != nullptr && env.rank == 0;
  st.clsword = 0;
  const uint32_t all_bits = n >= 32 ? 0xffffffffu : ((1u << n) - 1u);
  const bool lead = env.lead;

  {  // ---- initial state; every pending matrix starts as the identity
    for (int e = env.clane; e < 32; e += env.CL) qsb_pend_identity(ctl->pend[e]);
    qsb_desc* d = qsb_desc_begin(env, st);
    if (lead) {
      d->kind = QSB_D_INIT; d->flags = a.flags & QSB_RUN_LOAD;
      d->unit = unit; d->tile = st.tile;
      d->basis = a.init_basis ? a.init_basis[t] : a.default_basis;
      d->perm = a.idata + a.load_perm;
      d->gptr = a.states ? static_cast<char*>(a.states) + ((a.flags & QSB_RUN_LOAD_BROADCAST) ? 0 : t * st.dim) * a.amp_bytes : nullptr;
    }
    qsb_desc_end(env, st);
  }

  for (int64_t pc0 = 0; pc0 < a.n_ops; pc0 += QSB_CHUNK) {
    const int len = (int)((a.n_ops - pc0) < QSB_CHUNK ? (a.n_ops - pc0) : QSB_CHUNK);
    // ---- the decode warp staged this chunk (records, uniforms, decoded ops, per-slot MUL lists, SLOW mask)
    const int cb = (int)(st.gchunk & 1u);
    qsb_chunk* ck = &ctl->chunk[cb];
    unsigned long long pc_t0 = st.prof ? env.clock() : 0;
    env.dec_wait_full(cb, (st.gchunk >> 1) & 1u);
    if (st.prof) st.t_decode += env.clock() - pc_t0;     // cycles the control warp waited for the decode warp
    uint32_t slowmask[QSB_CHUNK / 32];
#pragma unroll
    for (int r = 0; r < QSB_CHUNK / 32; ++r) slowmask[r] = ck->slowmask[r];
    int cur[(32 + Env::CL - 1) / Env::CL];              // next unapplied op of the slot(s) this lane owns
    {
      int q = 0;
      for (int slot = env.clane; slot < 32; slot += env.CL) cur[q++] = ck->head[slot];
    }

    // ---- phase B: fold the MUL records up to the next SLOW op.  Control lane L owns slot L and walks that slot's
    // list; the lanes run their (independent) chains side by side, so a segment costs its longest chain.
    int i = 0;
    while (i < len) {
      pc_t0 = st.prof ? env.clock() : 0;
      env.sync_control();
      int stop = len;                                      // next SLOW op at or after i
#pragma unroll
      for (int r = QSB_CHUNK / 32 - 1; r >= 0; --r) {
        const uint32_t w32 = slowmask[r] & (r == (i >> 5) ? (0xffffffffu << (i & 31)) : (r > (i >> 5) ? 0xffffffffu : 0u));
        if (w32) stop = 32 * r + QSB_CTZ(w32);
      }
      uint64_t w = st.clsword;
      uint32_t lo = 0, hi = 0;                            // class bit planes rebuilt from the lanes
      {
        int q = 0;
        for (int slot = env.clane; slot < 32; slot += env.CL, ++q) {
          int c = qsb_cls_of(w, slot);
          if (cur[q] < stop) {
            c128 p0 = ctl->pend[slot][0], p1 = ctl->pend[slot][1], p2 = ctl->pend[slot][2], p3 = ctl->pend[slot][3];
            int j = cur[q];
            while (j < stop) {
              const qsb_dec* d = &ck->dec[j];
              const c128 u0 = d->U[0], u1 = d->U[1], u2 = d->U[2], u3 = d->U[3];
              const c128 n0 = qsb_fma(u1, p2, qsb_mul(u0, p0)), n1 = qsb_fma(u1, p3, qsb_mul(u0, p1));
              const c128 n2 = qsb_fma(u3, p2, qsb_mul(u2, p0)), n3 = qsb_fma(u3, p3, qsb_mul(u2, p1));
              p0 = n0; p1 = n1; p2 = n2; p3 = n3;
              c = c > d->ucls ? c : d->ucls;
              j = d->next;
            }
            cur[q] = j;
            ctl->pend[slot][0] = p0; ctl->pend[slot][1] = p1; ctl->pend[slot][2] = p2; ctl->pend[slot][3] = p3;
          }
          lo |= env.ballot_slot(c & 1, slot);
          hi |= env.ballot_slot(c >> 1, slot);
        }
      }
      i = stop;
      st.clsword = ((uint64_t)hi << 32) | lo;
      env.sync_control();
      if (st.prof) st.t_fold += env.clock() - pc_t0;
      if (i < len) {
        pc_t0 = st.prof ? env.clock() : 0;
        const unsigned long long rw0 = st.ring_wait;
        qsb_control_slow(env, a, st, ck, i);
        if (st.prof) { st.t_slow += env.clock() - pc_t0 - (st.ring_wait - rw0); st.n_slow += 1; }
        ++i;
      }
    }
    env.dec_release(cb);                                 // the decode warp may refill this buffer
    ++st.gchunk;
  }

  // ---- epilogue
  env.sync_control();
  qsb_flush(env, st, a.m, all_bits);
  if (a.flags & (QSB_RUN_STORE | QSB_RUN_ACCUM_PROBS))
    qsb_emit_store(env, st, a, a.idata + a.store_perm,
                   (a.flags & QSB_RUN_STORE) ? static_cast<char*>(a.states_out) + t * st.dim * a.amp_bytes : nullptr,
                   (a.flags & QSB_RUN_ACCUM_PROBS) ? a.probs_accum : nullptr);
}

// ---- decode warp: stages the op list one chunk ahead of the control warp -------------------------------------
// Lane-parallel and stateless with respect to the pending matrices: op records, the uniform of each draw, rotation
// matrices, "certain K0" tests, sweep group orders, the per-slot MUL lists and the SLOW-op mask.
template <class Env>
QSB_HD void qsb_decode_loop(Env& env, const qsb_exec_args& a, int64_t first, int64_t stride) {
  qsb_ctl* ctl = env.ctl();
  qsb_cstate st;
  st.seq = 0; st.gchunk = 0; st.prof = Env::PROF && a.prof != nullptr;
  unsigned long long t_dec = 0, t_wait = 0;
  const int64_t total = a.count << a.tile_bits;
  for (int64_t unit = first; unit < total; unit += stride) {
    const int64_t t = unit >> a.tile_bits;
    st.t = t;
    st.record = a.branches != nullptr && env.rank == 0;
    const qsb_op* ops = a.ops + t * a.ops_stride;
    const double* prm = a.params ? a.params + t * a.params_stride : nullptr;
    const double* uni = a.uniforms ? a.uniforms + t * a.uniforms_stride : nullptr;
    const uint64_t tglob = (uint64_t)(a.traj_offset + t);
    for (int64_t pc0 = 0; pc0 < a.n_ops; pc0 += QSB_CHUNK) {
      const int len = (int)((a.n_ops - pc0) < QSB_CHUNK ? (a.n_ops - pc0) : QSB_CHUNK);
      const int cb = (int)(st.gchunk & 1u);
      qsb_chunk* ck = &ctl->chunk[cb];
      unsigned long long c0 = st.prof ? env.clock() : 0;
      if (st.gchunk >= 2) env.dec_wait_empty(cb, ((st.gchunk >> 1) - 1u) & 1u);
      unsigned long long c1 = st.prof ? env.clock() : 0;
      uint32_t slowmask[QSB_CHUNK / 32];
#pragma unroll
      for (int r = 0; r < QSB_CHUNK / 32; ++r) slowmask[r] = 0;
      for (int base = 0; base < len; base += env.CL) {
        const int i = base + env.clane;
        int slow = 0;
        if (i < len) {
          const qsb_op op = ops[pc0 + i];
          double u = 0.0;
          if (op.draw >= 0) u = uni ? uni[op.draw] : qsb_philox_uniform(a.seed, tglob, (uint32_t)op.draw);
          ck->ops[i] = op;
          ck->u[i] = u;
          qsb_decode_op(env, a, st, op, u, prm, &ck->dec[i]);
          slow = ck->dec[i].type == QSB_DEC_SLOW;
        }
        const uint32_t bal = env.ballot_slot(slow, i & 31);
#pragma unroll
        for (int r = 0; r < QSB_CHUNK / 32; ++r) if (r == (i >> 5)) slowmask[r] |= bal;
      }
      if (env.lead)
        for (int r = 0; r < QSB_CHUNK / 32; ++r) ck->slowmask[r] = slowmask[r];
      qsb_link_chunk(env, ck, len);
      env.dec_publish(cb);
      ++st.gchunk;
      if (st.prof) { t_wait += c1 - c0; t_dec += env.clock() - c1; }
    }
  }
  if (st.prof && env.lead) {
    unsigned long long* o = a.prof + (size_t)env.cta_id() * QSB_PROF_WORDS;
    o[30] = t_dec; o[31] = t_wait;
  }
  if (env.C > 1) env.cluster_exit();
}

// control main loop over the units [first, total) of this CTA (stride = number of resident clusters / CTAs)
template <class Env>
QSB_HD void qsb_control_loop(Env& env, const qsb_exec_args& a, int64_t first, int64_t stride) {
  qsb_cstate st;
  st.seq = 0; st.gchunk = 0; st.clsword = 0; st.parity = 0;
  st.prof = Env::PROF && a.prof != nullptr; st.ring_wait = 0;
  st.t_decode = st.t_fold = st.t_slow = st.n_slow = 0;
  st.t_emit_body = st.t_emit_pub = st.n_emit = 0;
  st.t_e[0] = st.t_e[1] = st.t_e[2] = st.t_e[3] = 0;
  const unsigned long long c0 = st.prof ? env.clock() : 0;
  const int64_t total = a.count << a.tile_bits;
  for (int64_t u = first; u < total; u += stride) qsb_control_unit(env, st, a, u);
  if (st.prof && env.lead) {
    unsigned long long* o = a.prof + (size_t)env.cta_id() * QSB_PROF_WORDS;
    o[17] = st.ring_wait;
    o[18] = env.clock() - c0;
    o[19] = st.seq;
    o[20] = st.t_decode; o[21] = st.t_fold; o[22] = st.t_slow; o[23] = st.n_slow;
    o[24] = st.t_emit_body; o[25] = st.t_emit_pub; o[26] = st.n_emit;
    o[27] = st.t_e[0]; o[28] = st.t_e[1]; o[29] = st.t_e[2];
  }
  qsb_desc* d = qsb_desc_begin(env, st);
  if (env.lead) d->kind = QSB_D_EXIT;
  qsb_desc_end(env, st);
  if (env.C > 1) env.cluster_exit();
}
