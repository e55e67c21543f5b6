// emu_tsan_main.cpp -- TEST-ONLY: runs dumped executor programs through the host emulator under
// ThreadSanitizer (g++ -fsanitize=thread).  The CUDA executor and this emulator share qsb_exec.cuh, so a
// data race in the control-warp / worker / cluster protocol shows up here on a CPU-only machine.
//   usage: emu_tsan <dump file>...     (format written by tests/test_emu_tsan.py)
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "emu_exec.cpp"

template <class T>
static std::vector<T> rd(FILE* f) {
  int64_t n = 0;
  if (fread(&n, 8, 1, f) != 1) { fprintf(stderr, "short read\n"); exit(2); }
  std::vector<T> v((size_t)n);
  if (n && fread(v.data(), sizeof(T), (size_t)n, f) != (size_t)n) { fprintf(stderr, "short read\n"); exit(2); }
  return v;
}

int main(int argc, char** argv) {
  for (int k = 1; k < argc; ++k) {
    FILE* f = fopen(argv[k], "rb");
    if (!f) { perror(argv[k]); return 2; }
    std::vector<int64_t> h = rd<int64_t>(f);   // n, m, W, load_perm, store_perm, n_snapshots, flags, count, ops_stride, out_of_place
    std::vector<qsb_op> ops = rd<qsb_op>(f);
    std::vector<double> cdata = rd<double>(f);
    std::vector<int32_t> idata = rd<int32_t>(f);
    std::vector<double> uniforms = rd<double>(f);
    std::vector<double> params = rd<double>(f);
    std::vector<c128> states = rd<c128>(f);
    fclose(f);
    const int64_t count = h[7], dim = (int64_t)1 << h[0];
    std::vector<c128> snaps((size_t)(h[5] ? count * h[5] * dim : 1));
    std::vector<c128> out(states.size());
    std::vector<int32_t> branches((size_t)(uniforms.size() ? uniforms.size() : 1));
    const int64_t us = uniforms.size() / (count ? count : 1), ps = params.size() / (count ? count : 1);
    int rc = emu_run((int)h[0], (int)h[1], (int)h[2], ops.data(), h[8] ? h[8] : (int64_t)ops.size(), h[8], cdata.data(),
                     (int64_t)cdata.size(), idata.data(), (int)h[3], (int)h[4], (int)h[5], (int)h[6], states.data(), count,
                     ps ? params.data() : nullptr, ps, us ? uniforms.data() : nullptr, us, 7, 0, nullptr, 0,
                     us ? branches.data() : nullptr, us, h[5] ? snaps.data() : nullptr, nullptr, h.size() > 9 && h[9] ? out.data() : nullptr);
    printf("%s rc=%d\n", argv[k], rc);
    if (rc) return 1;
  }
  return 0;
}
