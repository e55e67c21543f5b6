// emu_exec.cpp -- TEST-ONLY host emulation of the resident-trajectory executor.
//
// Compiles quantum-simulator_b200/csrc/qsb_exec.cuh with g++ against a HostEnv made of a few
// OS threads and std::barrier, so the op loop, the index math and the host compiler
// (qsb/compiler.py) can be checked against the oracle on a machine without a GPU.  It exports one
// symbol (emu_run) that libqsb.so does not have; the product package never loads this library and
// has no switch that could route to it.
#include <barrier>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>
#include <cstring>

#include "qsb_exec.cuh"

struct Shared {
  int C, T, m;
  std::vector<std::vector<c128>> tiles;
  std::vector<qsb_ctl> ctls;
  std::vector<std::unique_ptr<std::barrier<>>> block_bar;
  std::unique_ptr<std::barrier<>> cluster_bar;
  std::vector<double> red;   // [C*T][4]
  std::mutex mu;
};

struct HostEnv {
  int tid, T, rank;
  Shared* sh;
  c128* tile() { return sh->tiles[rank].data(); }
  qsb_ctl* ctl() { return &sh->ctls[rank]; }
  void sync_block() { sh->block_bar[rank]->arrive_and_wait(); }
  void sync_cluster() { sh->cluster_bar->arrive_and_wait(); }
  const c128* peer_tile(int r) { return sh->tiles[r].data(); }
  void atomic_add(double* p, double v) { std::lock_guard<std::mutex> g(sh->mu); *p += v; }
  void allreduce(double* v, int nv) {
    double* mine = &sh->red[(rank * T + tid) * 4];
    for (int k = 0; k < nv; ++k) mine[k] = v[k];
    sync_cluster();
    for (int k = 0; k < nv; ++k) {
      double s = 0.0;
      for (int i = 0; i < sh->C * T; ++i) s += sh->red[i * 4 + k];
      v[k] = s;
    }
    sync_cluster();
  }
};

extern "C" int emu_run(int n, int m, int T, const qsb_op* ops, int64_t n_ops, int64_t ops_stride, const double* cdata, int64_t n_cdata,
                       const int32_t* idata, int load_perm, int store_perm, int n_snapshots, int flags, void* states,
                       int64_t count, const double* params, int64_t params_stride, const double* uniforms,
                       int64_t uniforms_stride, uint64_t seed, int64_t traj_offset, const int64_t* init_basis,
                       int64_t default_basis, int32_t* branches, int64_t branches_stride, void* snapshots,
                       double* probs_accum) {
  if (n < 1 || n > QSB_MAX_QUBITS || m < 1 || m > n || n - m > 3 || T < 1) return -1;
  Shared sh;
  sh.C = 1 << (n - m);
  sh.T = T;
  sh.m = m;
  sh.tiles.assign(sh.C, std::vector<c128>((size_t)1 << m));
  sh.ctls.resize(sh.C);
  for (int r = 0; r < sh.C; ++r) sh.block_bar.emplace_back(new std::barrier<>(T));
  sh.cluster_bar.reset(new std::barrier<>(sh.C * T));
  sh.red.assign((size_t)sh.C * T * 4, 0.0);

  qsb_exec_args a;
  memset(&a, 0, sizeof a);
  a.ops = ops; a.n_ops = n_ops; a.ops_stride = ops_stride; a.cdata = cdata; a.n_cdata = n_cdata; a.idata = idata;
  a.n = n; a.m = m; a.load_perm = load_perm; a.store_perm = store_perm; a.n_snapshots = n_snapshots;
  a.flags = flags; a.states = (c128*)states; a.count = count;
  a.params = params; a.params_stride = params_stride;
  a.uniforms = uniforms; a.uniforms_stride = uniforms_stride;
  a.seed = seed; a.traj_offset = traj_offset;
  a.init_basis = init_basis; a.default_basis = default_basis;
  a.branches = branches; a.branches_stride = branches_stride;
  a.snapshots = (c128*)snapshots; a.probs_accum = probs_accum;

  std::vector<std::thread> th;
  for (int r = 0; r < sh.C; ++r)
    for (int t = 0; t < T; ++t)
      th.emplace_back([&sh, &a, r, t, T]() {
        HostEnv env{t, T, r, &sh};
        for (int64_t j = 0; j < a.count; ++j) qsb_exec_trajectory(env, a, j);
      });
  for (auto& x : th) x.join();
  return 0;
}
