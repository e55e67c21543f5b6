// qsb_exec.cuh -- the resident-trajectory executor.
//
// One CTA cluster (1, 2, 4 or 8 CTAs) keeps one 2^n statevector in shared memory
// (2^m amplitudes per CTA, n - m cluster-rank bits) for a whole trajectory: gates,
// Kraus steps, snapshots and the final store, so a 16-qubit trajectory touches HBM
// only for its uniforms and its final state.  The op loop is written once against
// an `Env` (thread id, barriers, all-reduce, peer tile access):
//   * DeviceEnv  (qsb_kernels.cu)  -- CUDA: __syncthreads / cluster.sync / DSMEM
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
typedef double2 c128;
#else
#define QSB_HD inline
struct alignas(16) c128 { double x, y; };
#endif

#define QSB_MAX_QUBITS 16
#define QSB_MAX_LOCAL_BITS 13
#define QSB_AD_MARGIN 1e-10

struct qsb_exec_args {
  const qsb_op* ops;
  int64_t n_ops;
  int64_t ops_stride;     // 0: all trajectories share ops[0..n_ops)
  const double* cdata;
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

// ---- element passes (each thread owns whole groups, so no intra-pass hazards) -------
template <class Env>
QSB_HD void qsb_pass_u1(Env& env, c128* tile, int m, int b, c128 m00, c128 m01, c128 m10, c128 m11) {
  const int cnt = 1 << (m - 1);
  for (int g = env.tid; g < cnt; g += env.T) {
    int i0 = qsb_ins0(g, b), i1 = i0 | (1 << b);
    c128 a0 = tile[qsb_slot(i0)], a1 = tile[qsb_slot(i1)];
    tile[qsb_slot(i0)] = qsb_fma(m01, a1, qsb_mul(m00, a0));
    tile[qsb_slot(i1)] = qsb_fma(m11, a1, qsb_mul(m10, a0));
  }
}

template <class Env>
QSB_HD void qsb_pass_d1(Env& env, c128* tile, int m, int b, c128 d0, c128 d1) {
  const bool skip0 = (d0.x == 1.0 && d0.y == 0.0);
  const int cnt = 1 << (m - 1);
  for (int g = env.tid; g < cnt; g += env.T) {
    int i0 = qsb_ins0(g, b), i1 = i0 | (1 << b);
    if (!skip0) tile[qsb_slot(i0)] = qsb_mul(d0, tile[qsb_slot(i0)]);
    tile[qsb_slot(i1)] = qsb_mul(d1, tile[qsb_slot(i1)]);
  }
}

// Pauli on bit b: code 1 = X, 2 = Y, 3 = Z (gates.py:39-46)
template <class Env>
QSB_HD void qsb_pass_pauli(Env& env, c128* tile, int m, int b, int code) {
  const int cnt = 1 << (m - 1);
  for (int g = env.tid; g < cnt; g += env.T) {
    int s0 = qsb_slot(qsb_ins0(g, b)), s1 = qsb_slot(qsb_ins0(g, b) | (1 << b));
    c128 a0 = tile[s0], a1 = tile[s1];
    if (code == 1) { tile[s0] = a1; tile[s1] = a0; }
    else if (code == 2) { tile[s0] = qsb_c(a1.y, -a1.x); tile[s1] = qsb_c(-a0.y, a0.x); }
    else { tile[s1] = qsb_c(-a1.x, -a1.y); }
  }
}

// swap amplitudes (base|set|1<<x) <-> (base|set|1<<y) over all bases with the bits in `fixed` cleared
template <class Env>
QSB_HD void qsb_pass_swap(Env& env, c128* tile, int m, const int* sorted_bits, int nb, int set_mask,
                          int xa, int xb) {
  const int cnt = 1 << (m - nb);
  for (int g = env.tid; g < cnt; g += env.T) {
    int i = g;
    for (int k = 0; k < nb; ++k) i = qsb_ins0(i, sorted_bits[k]);
    i |= set_mask;
    int sa = qsb_slot(i | xa), sb = qsb_slot(i | xb);
    c128 t = tile[sa]; tile[sa] = tile[sb]; tile[sb] = t;
  }
}

QSB_HD void qsb_sort_bits(int* b, int nb) {
  for (int i = 1; i < nb; ++i) { int v = b[i], j = i - 1; while (j >= 0 && b[j] > v) { b[j + 1] = b[j]; --j; } b[j + 1] = v; }
}

// dense 2^K x 2^K (K = 2, 3); bits[0] = MSB of the matrix index; mat in (shared/global) memory
template <int K, class Env>
QSB_HD void qsb_pass_dense(Env& env, c128* tile, int m, const int* bits, const c128* mat) {
  constexpr int D = 1 << K;
  int sb[K];
  for (int k = 0; k < K; ++k) sb[k] = bits[k];
  qsb_sort_bits(sb, K);
  int off[D];
  for (int r = 0; r < D; ++r) {
    int o = 0;
    for (int k = 0; k < K; ++k) if ((r >> (K - 1 - k)) & 1) o |= 1 << bits[k];
    off[r] = o;
  }
  const int cnt = 1 << (m - K);
  for (int g = env.tid; g < cnt; g += env.T) {
    int base = g;
    for (int k = 0; k < K; ++k) base = qsb_ins0(base, sb[k]);
    c128 a[D], o[D];
    for (int r = 0; r < D; ++r) a[r] = tile[qsb_slot(base | off[r])];
    for (int r = 0; r < D; ++r) {
      c128 acc = qsb_mul(mat[r * D], a[0]);
      for (int c = 1; c < D; ++c) acc = qsb_fma(mat[r * D + c], a[c], acc);
      o[r] = acc;
    }
    for (int r = 0; r < D; ++r) tile[qsb_slot(base | off[r])] = o[r];
  }
}

// partial sums for the 1-qubit reduced density matrix of bit b (unnormalised):
// v[0] = sum |a0|^2, v[1] = sum |a1|^2, v[2] + i v[3] = sum a0 conj(a1)
template <class Env>
QSB_HD void qsb_partial_rdm1(Env& env, const c128* tile, int m, int b, double v[4]) {
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
QSB_HD double qsb_norm2_all(Env& env, const c128* tile, int m) {
  double v[4] = {0.0, 0.0, 0.0, 0.0};
  for (int i = env.tid; i < (1 << m); i += env.T) v[0] += qsb_norm2(tile[i]);   // slot order irrelevant
  env.allreduce(v, 1);
  return v[0];
}

// numpy Generator.choice(k, p) from its one uniform: cdf = cumsum(p); cdf /= cdf[-1];
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

// bit-permutation tables: index x (n <= 16 bits) -> sum_j bit_j(x) << perm[j]
template <class Env>
QSB_HD void qsb_build_perm(Env& env, const int32_t* perm, int n) {
  uint32_t* tab = env.perm_table();      // [512]
  for (int v = env.tid; v < 512; v += env.T) {
    int lo = v & 255, hi = v >> 8;       // hi = 0: table for bits 0..7, hi = 1: bits 8..15
    uint32_t r = 0;
    for (int j = 0; j < 8; ++j) {
      int bit = hi * 8 + j;
      if (bit < n && ((lo >> j) & 1)) r |= 1u << perm[bit];
    }
    tab[v] = r;
  }
  env.sync_block();
}
QSB_HD uint32_t qsb_permute(const uint32_t* tab, uint32_t x) { return tab[x & 255] | tab[256 + (x >> 8)]; }

// write the tile (scaled) to out[perm(x)], x = rank << m | i
template <class Env>
QSB_HD void qsb_store_state(Env& env, const c128* tile, const qsb_exec_args& a, const int32_t* perm,
                            double scale, c128* out, double* probs_accum) {
  qsb_build_perm(env, perm, a.n);
  const uint32_t* tab = env.perm_table();
  const int m = a.m;
  for (int i = env.tid; i < (1 << m); i += env.T) {
    uint32_t dst = qsb_permute(tab, ((uint32_t)env.rank << m) | (uint32_t)i);
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
  const qsb_op* ops = a.ops + t * a.ops_stride;
  const double* prm = a.params ? a.params + t * a.params_stride : nullptr;
  const double* uni = a.uniforms ? a.uniforms + t * a.uniforms_stride : nullptr;
  const uint64_t tglob = (uint64_t)(a.traj_offset + t);
  const int64_t dim = (int64_t)1 << n;

  // ---- initial state
  {
    qsb_build_perm(env, a.idata + a.load_perm, n);
    const uint32_t* tab = env.perm_table();
    if (a.flags & QSB_RUN_LOAD) {
      const c128* src = a.states + t * dim;
      for (int i = env.tid; i < (1 << m); i += env.T)
        tile[qsb_slot(i)] = src[qsb_permute(tab, ((uint32_t)env.rank << m) | (uint32_t)i)];
    } else {
      const uint32_t basis = (uint32_t)(a.init_basis ? a.init_basis[t] : a.default_basis);
      for (int i = env.tid; i < (1 << m); i += env.T)
        tile[qsb_slot(i)] = qsb_c(qsb_permute(tab, ((uint32_t)env.rank << m) | (uint32_t)i) == basis ? 1.0 : 0.0, 0.0);
    }
    env.sync_block();
  }

  for (int64_t pc = 0; pc < a.n_ops; ++pc) {
    const qsb_op op = ops[pc];
    const double* cd = a.cdata + (op.data >= 0 ? op.data : 0);
    switch (op.kind) {
      case QSB_OP_NOP:
        break;
      case QSB_OP_U1:
        qsb_pass_u1(env, tile, m, op.b0, qsb_c(cd[0], cd[1]), qsb_c(cd[2], cd[3]), qsb_c(cd[4], cd[5]),
                    qsb_c(cd[6], cd[7]));
        env.sync_block();
        break;
      case QSB_OP_D1:
        qsb_pass_d1(env, tile, m, op.b0, qsb_c(cd[0], cd[1]), qsb_c(cd[2], cd[3]));
        env.sync_block();
        break;
      case QSB_OP_U2: {
        int bits[2] = {op.b0, op.b1};
        qsb_pass_dense<2>(env, tile, m, bits, (const c128*)cd);
        env.sync_block();
        break;
      }
      case QSB_OP_U3Q: {
        int bits[3] = {op.b0, op.b1, op.b2};
        qsb_pass_dense<3>(env, tile, m, bits, (const c128*)cd);
        env.sync_block();
        break;
      }
      case QSB_OP_X: case QSB_OP_Y: case QSB_OP_Z:
        qsb_pass_pauli(env, tile, m, op.b0, op.kind - QSB_OP_X + 1);
        env.sync_block();
        break;
      case QSB_OP_CX: {       // |c=1>: swap t=0 <-> t=1
        int sb[2] = {op.b0, op.b1};
        qsb_sort_bits(sb, 2);
        qsb_pass_swap(env, tile, m, sb, 2, 1 << op.b0, 0, 1 << op.b1);
        env.sync_block();
        break;
      }
      case QSB_OP_CZ: {
        int sb[2] = {op.b0, op.b1};
        qsb_sort_bits(sb, 2);
        const int cnt = 1 << (m - 2), set = (1 << op.b0) | (1 << op.b1);
        for (int g = env.tid; g < cnt; g += env.T) {
          int s = qsb_slot(qsb_ins0(qsb_ins0(g, sb[0]), sb[1]) | set);
          c128 v = tile[s];
          tile[s] = qsb_c(-v.x, -v.y);
        }
        env.sync_block();
        break;
      }
      case QSB_OP_SWAP: {
        int sb[2] = {op.b0, op.b1};
        qsb_sort_bits(sb, 2);
        qsb_pass_swap(env, tile, m, sb, 2, 0, 1 << op.b0, 1 << op.b1);
        env.sync_block();
        break;
      }
      case QSB_OP_CCX: {
        int sb[3] = {op.b0, op.b1, op.b2};
        qsb_sort_bits(sb, 3);
        qsb_pass_swap(env, tile, m, sb, 3, (1 << op.b0) | (1 << op.b1), 0, 1 << op.b2);
        env.sync_block();
        break;
      }
      case QSB_OP_CSWAP: {
        int sb[3] = {op.b0, op.b1, op.b2};
        qsb_sort_bits(sb, 3);
        qsb_pass_swap(env, tile, m, sb, 3, 1 << op.b0, 1 << op.b1, 1 << op.b2);
        env.sync_block();
        break;
      }
      case QSB_OP_RX: case QSB_OP_RY: {     // gates.py:66-75
        double s, c;
        sincos(prm[op.param] * 0.5, &s, &c);
        if (op.kind == QSB_OP_RX)
          qsb_pass_u1(env, tile, m, op.b0, qsb_c(c, 0), qsb_c(0, -s), qsb_c(0, -s), qsb_c(c, 0));
        else
          qsb_pass_u1(env, tile, m, op.b0, qsb_c(c, 0), qsb_c(-s, 0), qsb_c(s, 0), qsb_c(c, 0));
        env.sync_block();
        break;
      }
      case QSB_OP_RZ: {                     // gates.py:78-80
        double s, c;
        sincos(prm[op.param] * 0.5, &s, &c);
        qsb_pass_d1(env, tile, m, op.b0, qsb_c(c, -s), qsb_c(c, s));
        env.sync_block();
        break;
      }
      case QSB_OP_PHASE: {                  // gates.py:83-85
        double s, c;
        sincos(prm[op.param], &s, &c);
        qsb_pass_d1(env, tile, m, op.b0, qsb_c(1, 0), qsb_c(c, s));
        env.sync_block();
        break;
      }
      case QSB_OP_U3: {                     // gates.py:88-94
        double s, c, sp, cp, sl, cl, spl, cpl;
        sincos(prm[op.param] * 0.5, &s, &c);
        sincos(prm[op.param + 1], &sp, &cp);
        sincos(prm[op.param + 2], &sl, &cl);
        sincos(prm[op.param + 1] + prm[op.param + 2], &spl, &cpl);
        qsb_pass_u1(env, tile, m, op.b0, qsb_c(c, 0), qsb_c(-cl * s, -sl * s), qsb_c(cp * s, sp * s),
                    qsb_c(cpl * c, spl * c));
        env.sync_block();
        break;
      }
      case QSB_OP_KRAUS_PAULI: {
        // K_i = sqrt(w_i) P_i: ||K_i psi||^2 / sum = w_i / sum(w) for any psi, so the branch needs no
        // reduction and (with the norm deferred to the store) the update is the bare Pauli.
        const double u = uni ? uni[op.draw] : qsb_philox_uniform(a.seed, tglob, (uint32_t)op.draw);
        int idx = 0;
        for (int i = 0; i < 3; ++i) if (cd[i] <= u) ++idx;
        const int code = (int)cd[3 + idx];
        if (a.branches && env.tid == 0 && env.rank == 0) a.branches[t * a.branches_stride + op.draw] = idx;
        if (code != 0) {
          qsb_pass_pauli(env, tile, m, op.b0, code);
          env.sync_block();
        }
        break;
      }
      case QSB_OP_KRAUS_AD: {
        // K0 = diag(1, sqrt(1-g)), K1 = sqrt(g)|0><1| (noise.py:98-103).  cdf[0] = p0/(p0+p1) >= 1-g, so a
        // draw below 1-g is K0 whatever the state; only the rest needs P(q=1).
        const double u = uni ? uni[op.draw] : qsb_philox_uniform(a.seed, tglob, (uint32_t)op.draw);
        const double gam = cd[0];
        int idx = 0;
        if (!(u < 1.0 - gam - QSB_AD_MARGIN)) {
          double v[4];
          qsb_partial_rdm1(env, tile, m, op.b0, v);
          env.allreduce(v, 2);
          double p[2] = {v[0] + (1.0 - gam) * v[1], gam * v[1]};
          idx = qsb_choice(p, 2, u);
        }
        if (a.branches && env.tid == 0 && env.rank == 0) a.branches[t * a.branches_stride + op.draw] = idx;
        if (idx == 0) {
          qsb_pass_d1(env, tile, m, op.b0, qsb_c(1, 0), qsb_c(cd[1], 0));
        } else {
          const int cnt = 1 << (m - 1);
          for (int g = env.tid; g < cnt; g += env.T) {
            int i0 = qsb_ins0(g, op.b0);
            int s0 = qsb_slot(i0), s1 = qsb_slot(i0 | (1 << op.b0));
            tile[s0] = tile[s1];
            tile[s1] = qsb_c(0, 0);
          }
        }
        env.sync_block();
        break;
      }
      case QSB_OP_KRAUS_GEN: {
        const double u = uni ? uni[op.draw] : qsb_philox_uniform(a.seed, tglob, (uint32_t)op.draw);
        const int nk = (int)cd[0];
        double v[4], p[8];
        qsb_partial_rdm1(env, tile, m, op.b0, v);
        env.allreduce(v, 4);
        for (int i = 0; i < nk; ++i) {
          const double* e = cd + 1 + i * 12 + 8;      // e00, e11, re e01, im e01 ; rho10 = conj(v2 + i v3)
          p[i] = e[0] * v[0] + e[1] * v[1] + 2.0 * (e[2] * v[2] + e[3] * v[3]);
        }
        const int idx = qsb_choice(p, nk, u);
        if (a.branches && env.tid == 0 && env.rank == 0) a.branches[t * a.branches_stride + op.draw] = idx;
        const double* k = cd + 1 + idx * 12;
        qsb_pass_u1(env, tile, m, op.b0, qsb_c(k[0], k[1]), qsb_c(k[2], k[3]), qsb_c(k[4], k[5]), qsb_c(k[6], k[7]));
        env.sync_block();
        break;
      }
      case QSB_OP_REMAP: {
        // swap cluster-rank bit b0 with local slot bit b1: pull the partner's half, then overwrite ours
        const int gb = op.b0, lb = op.b1;
        const int mybit = (env.rank >> gb) & 1;
        const c128* peer = env.peer_tile(env.rank ^ (1 << gb));
        constexpr int MAXE = 16;
        c128 val[MAXE];
        const int cnt = 1 << (m - 1);
        env.sync_cluster();                       // everyone finished the ops before the exchange
        // Round r pulls the partner's groups g and then overwrites OUR groups g (the ones the partner pulls
        // in the same round), so one cluster barrier between the two halves of a round is enough.
        for (int base = 0; base < cnt; base += MAXE * env.T) {
          for (int e = 0; e < MAXE; ++e) {
            int g = base + e * env.T + env.tid;
            if (g < cnt) val[e] = peer[qsb_slot(qsb_ins0(g, lb) | (mybit << lb))];
          }
          env.sync_cluster();
          for (int e = 0; e < MAXE; ++e) {
            int g = base + e * env.T + env.tid;
            if (g < cnt) tile[qsb_slot(qsb_ins0(g, lb) | ((1 - mybit) << lb))] = val[e];
          }
        }
        env.sync_block();
        break;
      }
      case QSB_OP_SNAPSHOT: {
        double scale = 1.0;
        if (a.flags & QSB_RUN_NORMALIZE) {
          double nn = qsb_norm2_all(env, tile, m);
          if (nn > 1e-30) scale = 1.0 / sqrt(nn);
        }
        if (a.snapshots)
          qsb_store_state(env, tile, a, a.idata + op.aux, scale,
                          a.snapshots + (t * a.n_snapshots + op.b0) * dim, nullptr);
        break;
      }
      default:
        break;
    }
  }

  // ---- epilogue
  if (a.flags & (QSB_RUN_STORE | QSB_RUN_ACCUM_PROBS)) {
    double scale = 1.0;
    if (a.flags & QSB_RUN_NORMALIZE) {
      double nn = qsb_norm2_all(env, tile, m);
      if (nn > 1e-30) scale = 1.0 / sqrt(nn);
    }
    qsb_store_state(env, tile, a, a.idata + a.store_perm, scale,
                    (a.flags & QSB_RUN_STORE) ? a.states + t * dim : nullptr,
                    (a.flags & QSB_RUN_ACCUM_PROBS) ? a.probs_accum : nullptr);
  }
  env.sync_block();
}
