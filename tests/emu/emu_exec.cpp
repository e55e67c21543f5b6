// emu_exec.cpp -- TEST-ONLY host emulation of the resident-tile executor.
//
// Compiles quantum-simulator_b200/csrc/qsb_exec.cuh with g++ against a HostEnv made of OS
// threads and std::barrier, so the control/worker protocol, the op loop, the index math and the
// host compiler (qsb/compiler.py) can be checked against the oracle on a machine without a GPU.
// It exports one symbol (emu_run) that libqsb.so does not have; the product package never loads
// this library and has no switch that could route to it.
#include <barrier>
#include <cstring>
#include <memory>
#include <mutex>
#include <optional>
#include <thread>
#include <vector>

#include "qsb_exec.cuh"

// One emulated CTA = W worker threads + 1 control thread; the barriers mirror the CUDA named barriers
// (std::barrier::arrive() = bar.arrive, arrive_and_wait() = bar.sync).
struct Cta {
  std::vector<c128> tile;
  qsb_ctl ctl;
  std::unique_ptr<std::barrier<>> full[QSB_RING], empty[QSB_RING], workers, all;
  std::unique_ptr<std::barrier<>> dfull[2], dempty[2];   // decode thread <-> control thread, per chunk buffer
};

struct Shared {
  int C, W, m;
  std::vector<std::unique_ptr<Cta>> cta;
  std::unique_ptr<std::barrier<>> cluster;    // the workers of every CTA of the cluster (mbarrier-based on the device)
  std::unique_ptr<std::barrier<>> cluster_all; // every thread of every CTA, like barrier.cluster (kernel exit)
  std::mutex mu;
};

struct HostEnv {
  typedef c128 amp;
  static constexpr int CL = 1;
  static constexpr bool PROF = false;
  int wid, W, wbits, rank, C;
  int cta;                      // index of this CTA's storage (== rank in cluster mode)
  int lane, warp, nwarps, clane;
  bool lead;
  Shared* sh;
  std::optional<std::barrier<>::arrival_token> tok;

  Cta& me() { return *sh->cta[cta]; }
  c128* tile() { return me().tile.data(); }
  qsb_ctl* ctl() { return &me().ctl; }
  const c128* peer_tile(int r) { return sh->cta[r]->tile.data(); }
  c128* peer_tile_w(int r) { return sh->cta[r]->tile.data(); }
  void fence_cluster() {}
  const qsb_ctl* peer_ctl(int r) { return C > 1 ? &sh->cta[r]->ctl : &me().ctl; }
  qsb_ctl* peer_ctl_w(int r) { return C > 1 ? &sh->cta[r]->ctl : &me().ctl; }
  void atomic_add(double* p, double v) { std::lock_guard<std::mutex> g(sh->mu); *p += v; }
  double warp_sum(double x) { return x; }     // one-lane "warps"
  unsigned long long clock() { return 0; }
  bool prof_on() { return false; }
  void prof_add(int, unsigned long long) {}
  int cta_id() { return cta; }
  uint32_t match_any(int) { return 1u; }
  uint32_t ballot_slot(int pred, int slot) { return pred ? (1u << slot) : 0u; }
  uint32_t or_reduce(uint32_t x) { return x; }
  int bcast_i(int x) { return x; }
  uint64_t bcast_u64(uint64_t x) { return x; }
  void dec_publish(int b) { (void)me().dfull[b]->arrive(); }
  void dec_wait_full(int b, uint32_t) { me().dfull[b]->arrive_and_wait(); }
  void dec_release(int b) { (void)me().dempty[b]->arrive(); }
  void dec_wait_empty(int b, uint32_t) { me().dempty[b]->arrive_and_wait(); }
  void ring_wait_empty(int s) { me().empty[s]->arrive_and_wait(); }
  void ring_publish(int s) { (void)me().full[s]->arrive(); }
  void ring_wait_full(int s) { me().full[s]->arrive_and_wait(); }
  void ring_release(int s) { (void)me().empty[s]->arrive(); }
  void sync_workers() { me().workers->arrive_and_wait(); }
  void sync_control() {}
  void cluster_sync_w() { if (C > 1) sh->cluster->arrive_and_wait(); else me().workers->arrive_and_wait(); }
  void cluster_exit() { sh->cluster_all->arrive_and_wait(); }
  void handoff_w() { me().all->arrive_and_wait(); }
  void handoff_c() { me().all->arrive_and_wait(); }
};

extern "C" int emu_run(int n, int m, int T, const qsb_op* ops, int64_t n_ops, int64_t ops_stride, const double* cdata, int64_t n_cdata,
                       const int32_t* idata, int load_perm, int store_perm, int n_snapshots, int flags, void* states,
                       int64_t count, const double* params, int64_t params_stride, const double* uniforms,
                       int64_t uniforms_stride, uint64_t seed, int64_t traj_offset, const int64_t* init_basis,
                       int64_t default_basis, int32_t* branches, int64_t branches_stride, void* snapshots,
                       double* probs_accum, void* states_out) {
  if (n < 1 || n > 30 || m < 1 || m > n || m > QSB_MAX_LOCAL_BITS || T < 8 || T > 32 || (T & (T - 1))) return -1;   // W: power of two >= 8
  const bool streaming = (n - m > 3) || n > QSB_MAX_QUBITS;
  Shared sh;
  sh.C = streaming ? 1 : 1 << (n - m);
  sh.W = T;
  sh.m = m;
  // streaming mode: two independent "CTAs" stride over the tiles (no cluster)
  const int n_cta = streaming ? 2 : sh.C;
  for (int r = 0; r < n_cta; ++r) {
    sh.cta.emplace_back(new Cta());
    Cta& c = *sh.cta.back();
    c.tile.assign((size_t)1 << m, c128{0.0, 0.0});
    memset(&c.ctl, 0, sizeof c.ctl);
    for (int s = 0; s < QSB_RING; ++s) {
      c.full[s].reset(new std::barrier<>(T + 1));
      c.empty[s].reset(new std::barrier<>(T + 1));
      if (s < 2) { c.dfull[s].reset(new std::barrier<>(2)); c.dempty[s].reset(new std::barrier<>(2)); }
    }
    c.workers.reset(new std::barrier<>(T));
    c.all.reset(new std::barrier<>(T + 1));
  }
  sh.cluster.reset(new std::barrier<>(sh.C * T));
  sh.cluster_all.reset(new std::barrier<>(sh.C * (T + 2)));

  qsb_exec_args a;
  memset(&a, 0, sizeof a);
  a.ops = ops; a.n_ops = n_ops; a.ops_stride = ops_stride; a.cdata = cdata; a.n_cdata = n_cdata; a.idata = idata;
  a.n = n; a.m = m; a.load_perm = load_perm; a.store_perm = store_perm; a.n_snapshots = n_snapshots;
  a.flags = flags; a.states = (c128*)states; a.states_out = states_out ? (c128*)states_out : (c128*)states; a.count = count;
  a.tile_bits = streaming ? n - m : 0;
  a.amp_bytes = 16;
  a.params = params; a.params_stride = params_stride;
  a.uniforms = uniforms; a.uniforms_stride = uniforms_stride;
  a.seed = seed; a.traj_offset = traj_offset;
  a.init_basis = init_basis; a.default_basis = default_basis;
  a.branches = branches; a.branches_stride = branches_stride;
  a.snapshots = (c128*)snapshots; a.probs_accum = probs_accum;

  std::vector<std::thread> th;
  for (int r = 0; r < n_cta; ++r)
    for (int t = 0; t <= T + 1; ++t)          // T workers, the control thread, the decode thread
      th.emplace_back([&sh, &a, r, t, T, streaming, n_cta]() {
        HostEnv env;
        env.sh = &sh; env.cta = r; env.rank = streaming ? 0 : r; env.C = sh.C; env.W = T; env.wbits = T == 8 ? 3 : T == 16 ? 4 : 5;
        env.wid = t < T ? t : -1;
        env.lane = 0; env.warp = t; env.nwarps = T; env.clane = 0; env.lead = true;
        if (env.wid >= 0) qsb_worker_loop(env, a);
        else if (t == T) { if (streaming) qsb_control_loop(env, a, r, n_cta); else qsb_control_loop(env, a, 0, 1); }
        else { if (streaming) qsb_decode_loop(env, a, r, n_cta); else qsb_decode_loop(env, a, 0, 1); }
      });
  for (auto& x : th) x.join();
  return 0;
}
