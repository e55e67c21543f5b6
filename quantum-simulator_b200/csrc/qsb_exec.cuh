// qsb_exec.cuh -- the resident-trajectory executor.
//
// One CTA cluster (1, 2, 4 or 8 CTAs) keeps one 2^n statevector in shared memory
// (2^m amplitudes per CTA, n - m cluster-rank bits) for a whole trajectory: gates,
// Kraus steps, snapshots and the final store, so a 16-qubit trajectory touches HBM
// only for its op list, its uniforms and its final state.
//
// Runtime 1-qubit fusion: every 1-qubit operation (gate, Pauli branch, amplitude-damping
// K0/K1, generic Kraus operator) is multiplied into a pending 2x2 matrix of its slot bit
// by one thread; the state is only swept when a multi-qubit gate, a cluster remap, a
// state-dependent Kraus draw or a store needs the affected bits, and that sweep applies the
// pending matrices of its bits and the gate in one pass over shared memory.
//
// The op loop is written once against an `Env` (thread id, barriers, all-reduce, peer tile):
//   * DeviceEnv  (qsb_kernels.cuh) -- CUDA: __syncthreads / cluster.sync / DSMEM
//   * HostEnv    (tests/emu)       -- test-only: a few OS threads + std::barrier, used
//                                     to check index math and the host compiler on CPU.
//
// Reference semantics implemented here (file:line in the reference tree):
//   StateVector.apply_gate      state_vector.py:41-74   (textbook action; the axis
//                               scramble of :66-73 is bookkeeping done by the host compiler)
//   NoiseModel._apply_channel   noise.py:224-260        (Kraus selection + renormalise)
//   gate formulas               gates.py:37-125
#pragma once

#include <math.h>
#include <stdint.h>

#include "qsb.h"

#if defined(__CUDACC__)
#define QSB_HD __host__ __device__ __forceinline__
#define QSB_PASS __device__ __noinline__      // whole-tile sweeps: own register allocation, called from the op loop
typedef double2 c128;
#else
#define QSB_HD inline
#define QSB_PASS inline
struct alignas(16) c128 { double x, y; };
#endif

#define QSB_MAX_QUBITS 16
#define QSB_MAX_LOCAL_BITS 13
#define QSB_AD_MARGIN 1e-10
#define QSB_CHUNK 128          // ops staged in shared memory per refill
#define QSB_REMAP_REGS 8       // amplitudes a thread stages per remap round

struct qsb_exec_args {
  const qsb_op* ops;
  int64_t n_ops;
  int64_t ops_stride;     // 0: all trajectories share ops[0..n_ops)
  const double* cdata;
  int64_t n_cdata;
  const int32_t* idata;
  int32_t n, m;
  int32_t load_perm, store_perm, n_snapshots;
  int32_t flags;
  c128* states;           // already offset to `first`
  int64_t count;
  const double* params;   int64_t params_stride;
  const double* uniforms; int64_t uniforms_stride;
  uint64_t seed;          int64_t traj_offset;
  const int64_t* init_basis; int64_t default_basis;
  int32_t* branches;      int64_t branches_stride;
  c128* snapshots;
  double* probs_accum;
};

// per-CTA control block that lives behind the tile in shared memory
struct qsb_ctl {
  uint32_t perm[512];          // bit-permutation tables (load / store / snapshot)
  c128 pend[16][4];            // pending 2x2 per slot bit (row-major), valid where the mask bit is set
  c128 mat[64];                // dense 4x4 / 8x8 gate staged for the current pass
  qsb_op ops[QSB_CHUNK];       // staged op records ...
  double u[QSB_CHUNK];         // ... the uniform each Kraus op consumes
  double prm[QSB_CHUNK][3];    // ... the angles each parameterised gate reads
  double cd[QSB_CHUNK][8];     // ... and the first 8 doubles of each op's cdata
};

// ---- small helpers ---------------------------------------------------------------
// Swizzled slot of amplitude i inside the tile: XOR bits 3..5 into bits 0..2 so that the
// 8 lanes of a quarter-warp hit 8 different 16-byte bank groups for every target-bit choice.
QSB_HD int qsb_slot(int i) { return i ^ ((i >> 3) & 7); }
// insert a zero bit at position b
QSB_HD int qsb_ins0(int g, int b) { return ((g >> b) << (b + 1)) | (g & ((1 << b) - 1)); }

QSB_HD c128 qsb_c(double x, double y) { c128 r; r.x = x; r.y = y; return r; }
QSB_HD c128 qsb_mul(c128 a, c128 b) { return qsb_c(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
QSB_HD c128 qsb_fma(c128 a, c128 b, c128 c) {   // a*b + c
  return qsb_c(fma(a.x, b.x, fma(-a.y, b.y, c.x)), fma(a.x, b.y, fma(a.y, b.x, c.y)));
}
QSB_HD c128 qsb_neg(c128 a) { return qsb_c(-a.x, -a.y); }
QSB_HD double qsb_norm2(c128 a) { return a.x * a.x + a.y * a.y; }

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

// ---- pending 2x2 matrices (one writer: thread 0 of each CTA; every CTA of the cluster keeps an
//      identical copy because all of them walk the same ops with the same uniforms) -----------
// P <- U P, or P <- U when the bit had nothing pending
QSB_HD void qsb_pend_mul(c128* P, bool has, c128 u00, c128 u01, c128 u10, c128 u11) {
  if (!has) { P[0] = u00; P[1] = u01; P[2] = u10; P[3] = u11; return; }
  c128 p00 = P[0], p01 = P[1], p10 = P[2], p11 = P[3];
  P[0] = qsb_fma(u01, p10, qsb_mul(u00, p00));
  P[1] = qsb_fma(u01, p11, qsb_mul(u00, p01));
  P[2] = qsb_fma(u11, p10, qsb_mul(u10, p00));
  P[3] = qsb_fma(u11, p11, qsb_mul(u10, p01));
}
QSB_HD void qsb_pend_diag(c128* P, bool has, c128 d0, c128 d1) {
  if (!has) { P[0] = d0; P[1] = qsb_c(0, 0); P[2] = qsb_c(0, 0); P[3] = d1; return; }
  P[0] = qsb_mul(d0, P[0]); P[1] = qsb_mul(d0, P[1]);
  P[2] = qsb_mul(d1, P[2]); P[3] = qsb_mul(d1, P[3]);
}
// Pauli code 1 = X, 2 = Y, 3 = Z (gates.py:39-46)
QSB_HD void qsb_pend_pauli(c128* P, bool has, int code) {
  if (!has) { P[0] = qsb_c(1, 0); P[1] = qsb_c(0, 0); P[2] = qsb_c(0, 0); P[3] = qsb_c(1, 0); }
  c128 p00 = P[0], p01 = P[1], p10 = P[2], p11 = P[3];
  if (code == 1) { P[0] = p10; P[1] = p11; P[2] = p00; P[3] = p01; }
  else if (code == 2) {           // Y = [[0,-i],[i,0]]: row0' = -i row1, row1' = i row0
    P[0] = qsb_c(p10.y, -p10.x); P[1] = qsb_c(p11.y, -p11.x);
    P[2] = qsb_c(-p00.y, p00.x); P[3] = qsb_c(-p01.y, p01.x);
  } else { P[2] = qsb_neg(p10); P[3] = qsb_neg(p11); }
}

// gate applied inside a fused pass after the pending matrices of its bits
enum { QSB_G_NONE = 0, QSB_G_CX, QSB_G_CZ, QSB_G_SWAP, QSB_G_CCX, QSB_G_CSWAP, QSB_G_DENSE };

// One sweep over the tile for K slot bits (bits[0] = MSB of the local index r): apply the pending 2x2 of
// every bit in `pmask` (bit k of pmask <-> bits[k]), then gate G.  Each thread owns whole 2^K groups.
template <int K, int G, class Env>
QSB_PASS void qsb_pass_fused(Env& env, int m, const int* bits, int pmask) {
  c128* tile = env.tile();          // re-derived here so device code keeps the shared address space (LDS/STS)
  const qsb_ctl* ctl = env.ctl();
  constexpr int D = 1 << K;
  int sb[K];
#pragma unroll
  for (int k = 0; k < K; ++k) sb[k] = bits[k];
  qsb_sort_bits<K>(sb);
  int off[D];
#pragma unroll
  for (int r = 0; r < D; ++r) {
    int o = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) if ((r >> (K - 1 - k)) & 1) o |= 1 << bits[k];
    off[r] = o;
  }
  c128 P[K][4];
#pragma unroll
  for (int k = 0; k < K; ++k)
    if ((pmask >> k) & 1) {
#pragma unroll
      for (int e = 0; e < 4; ++e) P[k][e] = ctl->pend[bits[k]][e];
    }
  const int cnt = 1 << (m - K);
  for (int g = env.tid; g < cnt; g += env.T) {
    int base = g;
#pragma unroll
    for (int k = 0; k < K; ++k) base = qsb_ins0(base, sb[k]);
    c128 a[D];
#pragma unroll
    for (int r = 0; r < D; ++r) a[r] = tile[qsb_slot(base | off[r])];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      if (!((pmask >> k) & 1)) continue;
      const int bit = 1 << (K - 1 - k);
#pragma unroll
      for (int r = 0; r < D; ++r) {
        if (r & bit) continue;
        c128 lo = a[r], hi = a[r | bit];
        a[r] = qsb_fma(P[k][1], hi, qsb_mul(P[k][0], lo));
        a[r | bit] = qsb_fma(P[k][3], hi, qsb_mul(P[k][2], lo));
      }
    }
    if (G == QSB_G_DENSE) {
      // every output row is written straight to the tile: the group is owned by this thread and all of
      // its inputs are already in registers
#pragma unroll
      for (int r = 0; r < D; ++r) {
        c128 acc = qsb_mul(ctl->mat[r * D], a[0]);
#pragma unroll
        for (int c = 1; c < D; ++c) acc = qsb_fma(ctl->mat[r * D + c], a[c], acc);
        tile[qsb_slot(base | off[r])] = acc;
      }
      continue;
    }
    if (G == QSB_G_CX) { c128 t = a[2]; a[2] = a[3]; a[3] = t; }                  // bits[0] control, bits[1] target
    else if (G == QSB_G_CZ) { a[3] = qsb_neg(a[3]); }
    else if (G == QSB_G_SWAP) { c128 t = a[1]; a[1] = a[2]; a[2] = t; }
    else if (G == QSB_G_CCX) { c128 t = a[D - 2]; a[D - 2] = a[D - 1]; a[D - 1] = t; }   // 110 <-> 111
    else if (G == QSB_G_CSWAP) { c128 t = a[(D >> 1) | 1]; a[(D >> 1) | 1] = a[(D >> 1) | 2]; a[(D >> 1) | 2] = t; }  // 101 <-> 110
#pragma unroll
    for (int r = 0; r < D; ++r) tile[qsb_slot(base | off[r])] = a[r];
  }
}

// apply and clear every pending matrix in `which` (slot-bit mask), three bits per sweep.  Collective.
template <class Env>
QSB_PASS void qsb_flush(Env& env, int m, uint32_t& mask, uint32_t which) {
  uint32_t todo = mask & which;
  if (!todo) return;
  env.sync_block();                       // thread 0's pending updates are visible, previous sweep done
  while (todo) {
    int bits[3], nb = 0;
    while (todo && nb < 3) {
      int b = 15;
      while (!((todo >> b) & 1u)) --b;
      bits[nb++] = b;
      todo &= ~(1u << b);
    }
    if (nb == 3) qsb_pass_fused<3, QSB_G_NONE>(env, m, bits, 7);
    else if (nb == 2) qsb_pass_fused<2, QSB_G_NONE>(env, m, bits, 3);
    else qsb_pass_fused<1, QSB_G_NONE>(env, m, bits, 1);
    env.sync_block();
  }
  mask &= ~which;
}

// partial sums for the 1-qubit reduced density matrix of bit b (unnormalised):
// v[0] = sum |a0|^2, v[1] = sum |a1|^2, v[2] + i v[3] = sum a0 conj(a1)
template <class Env>
QSB_PASS void qsb_partial_rdm1(Env& env, int m, int b, double v[4]) {
  const c128* tile = env.tile();
  v[0] = v[1] = v[2] = v[3] = 0.0;
  const int cnt = 1 << (m - 1);
  for (int g = env.tid; g < cnt; g += env.T) {
    int i0 = qsb_ins0(g, b);
    c128 a0 = tile[qsb_slot(i0)], a1 = tile[qsb_slot(i0 | (1 << b))];
    v[0] += qsb_norm2(a0);
    v[1] += qsb_norm2(a1);
    v[2] += a0.x * a1.x + a0.y * a1.y;
    v[3] += a0.y * a1.x - a0.x * a1.y;
  }
}

template <class Env>
QSB_PASS double qsb_norm2_all(Env& env, int m) {
  const c128* tile = env.tile();
  double v[4] = {0.0, 0.0, 0.0, 0.0};
  for (int i = env.tid; i < (1 << m); i += env.T) v[0] += qsb_norm2(tile[i]);   // slot order irrelevant
  env.allreduce(v, 1);
  return v[0];
}

// bit-permutation tables: index x (n <= 16 bits) -> sum_j bit_j(x) << perm[j]
template <class Env>
QSB_HD void qsb_build_perm(Env& env, qsb_ctl* ctl, const int32_t* perm, int n) {
  for (int v = env.tid; v < 512; v += env.T) {
    int lo = v & 255, hi = v >> 8;       // hi = 0: table for bits 0..7, hi = 1: bits 8..15
    uint32_t r = 0;
    for (int j = 0; j < 8; ++j) {
      int bit = hi * 8 + j;
      if (bit < n && ((lo >> j) & 1)) r |= 1u << perm[bit];
    }
    ctl->perm[v] = r;
  }
  env.sync_block();
}
QSB_HD uint32_t qsb_permute(const uint32_t* tab, uint32_t x) { return tab[x & 255] | tab[256 + (x >> 8)]; }

// write the tile (scaled) to out[perm(x)], x = rank << m | i
template <class Env>
QSB_PASS void qsb_store_state(Env& env, const qsb_exec_args& a, const int32_t* perm,
                            double scale, c128* out, double* probs_accum) {
  const c128* tile = env.tile();
  qsb_ctl* ctl = env.ctl();
  qsb_build_perm(env, ctl, perm, a.n);
  const int m = a.m;
  for (int i = env.tid; i < (1 << m); i += env.T) {
    uint32_t dst = qsb_permute(ctl->perm, ((uint32_t)env.rank << m) | (uint32_t)i);
    c128 v = tile[qsb_slot(i)];
    v.x *= scale; v.y *= scale;
    if (out) out[dst] = v;
    if (probs_accum) env.atomic_add(probs_accum + dst, qsb_norm2(v));
  }
  env.sync_block();
}

// ---- one trajectory ------------------------------------------------------------------
template <class Env>
QSB_HD void qsb_exec_trajectory(Env& env, const qsb_exec_args& a, int64_t t) {
  const int m = a.m, n = a.n;
  c128* tile = env.tile();
  qsb_ctl* ctl = env.ctl();
  const qsb_op* ops = a.ops + t * a.ops_stride;
  const double* prm = a.params ? a.params + t * a.params_stride : nullptr;
  const double* uni = a.uniforms ? a.uniforms + t * a.uniforms_stride : nullptr;
  const uint64_t tglob = (uint64_t)(a.traj_offset + t);
  const int64_t dim = (int64_t)1 << n;
  const uint32_t all_local = (1u << m) - 1u;
  const bool lead = env.tid == 0;
  const bool record = a.branches && lead && env.rank == 0;
  uint32_t mask = 0;                 // slot bits with a pending matrix; every thread keeps the same value

  // ---- initial state
  {
    qsb_build_perm(env, ctl, a.idata + a.load_perm, n);
    if (a.flags & QSB_RUN_LOAD) {
      const c128* src = a.states + t * dim;
      for (int i = env.tid; i < (1 << m); i += env.T)
        tile[qsb_slot(i)] = src[qsb_permute(ctl->perm, ((uint32_t)env.rank << m) | (uint32_t)i)];
    } else {
      const uint32_t basis = (uint32_t)(a.init_basis ? a.init_basis[t] : a.default_basis);
      for (int i = env.tid; i < (1 << m); i += env.T)
        tile[qsb_slot(i)] = qsb_c(qsb_permute(ctl->perm, ((uint32_t)env.rank << m) | (uint32_t)i) == basis ? 1.0 : 0.0, 0.0);
    }
    env.sync_block();
  }

  for (int64_t pc0 = 0; pc0 < a.n_ops; pc0 += QSB_CHUNK) {
    const int len = (int)((a.n_ops - pc0) < QSB_CHUNK ? (a.n_ops - pc0) : QSB_CHUNK);
    // ---- stage a chunk: op records, their uniforms / angles / leading cdata (one latency per chunk)
    env.sync_block();
    for (int i = env.tid; i < len; i += env.T) {
      const qsb_op op = ops[pc0 + i];
      ctl->ops[i] = op;
      if (op.draw >= 0) ctl->u[i] = uni ? uni[op.draw] : qsb_philox_uniform(a.seed, tglob, (uint32_t)op.draw);
      if (op.param >= 0 && prm) {
        const int np = op.kind == QSB_OP_U3 ? 3 : 1;
        for (int k = 0; k < np; ++k) ctl->prm[i][k] = prm[op.param + k];
      }
      if (op.data >= 0)
        for (int k = 0; k < 8; ++k) ctl->cd[i][k] = (op.data + k < a.n_cdata) ? a.cdata[op.data + k] : 0.0;
    }
    env.sync_block();

    for (int i = 0; i < len; ++i) {
      const qsb_op op = ctl->ops[i];
      const double* cd = ctl->cd[i];
      const int b = op.b0;
      const bool has = (mask >> b) & 1u;
      switch (op.kind) {
        case QSB_OP_NOP:
          break;
        // ---------------- 1-qubit operations: fold into the pending matrix of bit b
        case QSB_OP_U1:
          if (lead) qsb_pend_mul(ctl->pend[b], has, qsb_c(cd[0], cd[1]), qsb_c(cd[2], cd[3]), qsb_c(cd[4], cd[5]), qsb_c(cd[6], cd[7]));
          mask |= 1u << b;
          break;
        case QSB_OP_D1:
          if (lead) qsb_pend_diag(ctl->pend[b], has, qsb_c(cd[0], cd[1]), qsb_c(cd[2], cd[3]));
          mask |= 1u << b;
          break;
        case QSB_OP_X: case QSB_OP_Y: case QSB_OP_Z:
          if (lead) qsb_pend_pauli(ctl->pend[b], has, op.kind - QSB_OP_X + 1);
          mask |= 1u << b;
          break;
        case QSB_OP_RX: case QSB_OP_RY: {     // gates.py:66-75
          if (lead) {
            double s, c;
            sincos(ctl->prm[i][0] * 0.5, &s, &c);
            if (op.kind == QSB_OP_RX) qsb_pend_mul(ctl->pend[b], has, qsb_c(c, 0), qsb_c(0, -s), qsb_c(0, -s), qsb_c(c, 0));
            else qsb_pend_mul(ctl->pend[b], has, qsb_c(c, 0), qsb_c(-s, 0), qsb_c(s, 0), qsb_c(c, 0));
          }
          mask |= 1u << b;
          break;
        }
        case QSB_OP_RZ: {                     // gates.py:78-80
          if (lead) {
            double s, c;
            sincos(ctl->prm[i][0] * 0.5, &s, &c);
            qsb_pend_diag(ctl->pend[b], has, qsb_c(c, -s), qsb_c(c, s));
          }
          mask |= 1u << b;
          break;
        }
        case QSB_OP_PHASE: {                  // gates.py:83-85
          if (lead) {
            double s, c;
            sincos(ctl->prm[i][0], &s, &c);
            qsb_pend_diag(ctl->pend[b], has, qsb_c(1, 0), qsb_c(c, s));
          }
          mask |= 1u << b;
          break;
        }
        case QSB_OP_U3: {                     // gates.py:88-94
          if (lead) {
            double s, c, sp, cp, sl, cl, spl, cpl;
            sincos(ctl->prm[i][0] * 0.5, &s, &c);
            sincos(ctl->prm[i][1], &sp, &cp);
            sincos(ctl->prm[i][2], &sl, &cl);
            sincos(ctl->prm[i][1] + ctl->prm[i][2], &spl, &cpl);
            qsb_pend_mul(ctl->pend[b], has, qsb_c(c, 0), qsb_c(-cl * s, -sl * s), qsb_c(cp * s, sp * s), qsb_c(cpl * c, spl * c));
          }
          mask |= 1u << b;
          break;
        }
        case QSB_OP_KRAUS_PAULI: {
          // K_i = sqrt(w_i) P_i: ||K_i psi||^2 / sum = w_i / sum(w) for any psi, so the branch needs no
          // reduction and (with the norm deferred to the store) the update is the bare Pauli.
          const double u = ctl->u[i];
          int idx = 0;
          for (int k = 0; k < 3; ++k) if (cd[k] <= u) ++idx;
          const int code = (int)cd[3 + idx];
          if (record) a.branches[t * a.branches_stride + op.draw] = idx;
          if (code != 0) {
            if (lead) qsb_pend_pauli(ctl->pend[b], has, code);
            mask |= 1u << b;
          }
          break;
        }
        case QSB_OP_KRAUS_AD: {
          // K0 = diag(1, sqrt(1-g)), K1 = sqrt(g)|0><1| (noise.py:98-103).  cdf[0] = p0/(p0+p1) >= 1-g, so a
          // draw below 1-g is K0 whatever the state; only the rest needs P(q=1) of the actual state.
          const double u = ctl->u[i];
          const double gam = cd[0];
          int idx = 0;
          bool had = has;
          if (!(u < 1.0 - gam - QSB_AD_MARGIN)) {
            qsb_flush(env, m, mask, all_local);
            had = false;
            double v[4];
            qsb_partial_rdm1(env, m, b, v);
            env.allreduce(v, 2);
            double p[2] = {v[0] + (1.0 - gam) * v[1], gam * v[1]};
            idx = qsb_choice(p, 2, u);
          }
          if (record) a.branches[t * a.branches_stride + op.draw] = idx;
          if (lead) {
            if (idx == 0) qsb_pend_diag(ctl->pend[b], had, qsb_c(1, 0), qsb_c(cd[1], 0));
            else qsb_pend_mul(ctl->pend[b], had, qsb_c(0, 0), qsb_c(1, 0), qsb_c(0, 0), qsb_c(0, 0));   // a0 <- a1, a1 <- 0
          }
          mask |= 1u << b;
          break;
        }
        case QSB_OP_KRAUS_GEN: {
          const double u = ctl->u[i];
          const double* g = a.cdata + op.data;
          const int nk = (int)g[0];
          double v[4], p[8];
          qsb_flush(env, m, mask, all_local);
          qsb_partial_rdm1(env, m, b, v);
          env.allreduce(v, 4);
          for (int k = 0; k < nk; ++k) {
            const double* e = g + 1 + k * 12 + 8;       // e00, e11, re e01, im e01 ; rho10 = conj(v2 + i v3)
            p[k] = e[0] * v[0] + e[1] * v[1] + 2.0 * (e[2] * v[2] + e[3] * v[3]);
          }
          const int idx = qsb_choice(p, nk, u);
          if (record) a.branches[t * a.branches_stride + op.draw] = idx;
          const double* kk = g + 1 + idx * 12;
          if (lead) qsb_pend_mul(ctl->pend[b], false, qsb_c(kk[0], kk[1]), qsb_c(kk[2], kk[3]), qsb_c(kk[4], kk[5]), qsb_c(kk[6], kk[7]));
          mask |= 1u << b;
          break;
        }
        // ---------------- multi-qubit gates: one sweep = pending matrices of the bits + the gate
        case QSB_OP_CX: case QSB_OP_CZ: case QSB_OP_SWAP: case QSB_OP_U2: {
          int bits[2] = {op.b0, op.b1};
          const int pm = (int)((mask >> op.b0) & 1u) | (int)(((mask >> op.b1) & 1u) << 1);
          env.sync_block();
          if (op.kind == QSB_OP_U2) {
            for (int k = env.tid; k < 16; k += env.T) ctl->mat[k] = ((const c128*)(a.cdata + op.data))[k];
            env.sync_block();
            qsb_pass_fused<2, QSB_G_DENSE>(env, m, bits, pm);
          } else if (op.kind == QSB_OP_CX) qsb_pass_fused<2, QSB_G_CX>(env, m, bits, pm);
          else if (op.kind == QSB_OP_CZ) qsb_pass_fused<2, QSB_G_CZ>(env, m, bits, pm);
          else qsb_pass_fused<2, QSB_G_SWAP>(env, m, bits, pm);
          env.sync_block();
          mask &= ~((1u << op.b0) | (1u << op.b1));
          break;
        }
        case QSB_OP_CCX: case QSB_OP_CSWAP: case QSB_OP_U3Q: {
          int bits[3] = {op.b0, op.b1, op.b2};
          const int pm = (int)((mask >> op.b0) & 1u) | (int)(((mask >> op.b1) & 1u) << 1) | (int)(((mask >> op.b2) & 1u) << 2);
          env.sync_block();
          if (op.kind == QSB_OP_U3Q) {
            for (int k = env.tid; k < 64; k += env.T) ctl->mat[k] = ((const c128*)(a.cdata + op.data))[k];
            env.sync_block();
            qsb_pass_fused<3, QSB_G_DENSE>(env, m, bits, pm);
          } else if (op.kind == QSB_OP_CCX) qsb_pass_fused<3, QSB_G_CCX>(env, m, bits, pm);
          else qsb_pass_fused<3, QSB_G_CSWAP>(env, m, bits, pm);
          env.sync_block();
          mask &= ~((1u << op.b0) | (1u << op.b1) | (1u << op.b2));
          break;
        }
        case QSB_OP_REMAP: {
          // swap cluster-rank bit b0 with local slot bit b1: pull the partner's half, then overwrite ours
          const int gb = op.b0, lb = op.b1;
          qsb_flush(env, m, mask, 1u << lb);      // the outgoing qubit takes no pending matrix along
          const int mybit = (env.rank >> gb) & 1;
          const c128* peer = env.peer_tile(env.rank ^ (1 << gb));
          c128 val[QSB_REMAP_REGS];
          const int cnt = 1 << (m - 1);
          env.sync_cluster();                       // everyone finished the ops before the exchange
          // Round r pulls the partner's groups g and then overwrites OUR groups g (the ones the partner pulls
          // in the same round), so one cluster barrier between the two halves of a round is enough.
          for (int base = 0; base < cnt; base += QSB_REMAP_REGS * env.T) {
            for (int e = 0; e < QSB_REMAP_REGS; ++e) {
              int g = base + e * env.T + env.tid;
              if (g < cnt) val[e] = peer[qsb_slot(qsb_ins0(g, lb) | (mybit << lb))];
            }
            env.sync_cluster();
            for (int e = 0; e < QSB_REMAP_REGS; ++e) {
              int g = base + e * env.T + env.tid;
              if (g < cnt) tile[qsb_slot(qsb_ins0(g, lb) | ((1 - mybit) << lb))] = val[e];
            }
          }
          env.sync_block();
          break;
        }
        case QSB_OP_SNAPSHOT: {
          qsb_flush(env, m, mask, all_local);
          double scale = 1.0;
          if (a.flags & QSB_RUN_NORMALIZE) {
            double nn = qsb_norm2_all(env, m);
            if (nn > 1e-30) scale = 1.0 / sqrt(nn);
          }
          if (a.snapshots)
            qsb_store_state(env, a, a.idata + op.aux, scale,
                            a.snapshots + (t * a.n_snapshots + op.b0) * dim, nullptr);
          break;
        }
        default:
          break;
      }
    }
  }

  // ---- epilogue
  qsb_flush(env, m, mask, all_local);
  env.sync_block();
  if (a.flags & (QSB_RUN_STORE | QSB_RUN_ACCUM_PROBS)) {
    double scale = 1.0;
    if (a.flags & QSB_RUN_NORMALIZE) {
      double nn = qsb_norm2_all(env, m);
      if (nn > 1e-30) scale = 1.0 / sqrt(nn);
    }
    qsb_store_state(env, a, a.idata + a.store_perm, scale,
                    (a.flags & QSB_RUN_STORE) ? a.states + t * dim : nullptr,
                    (a.flags & QSB_RUN_ACCUM_PROBS) ? a.probs_accum : nullptr);
  }
  env.sync_block();
}
