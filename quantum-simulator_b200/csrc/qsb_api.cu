// qsb_api.cu -- extern "C" entry points of libqsb.so (see include/qsb.h).
// Thin host glue: handles, argument checks, launches.  No CPU compute path exists here:
// every function that produces numbers launches a kernel from qsb_kernels.cuh.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "qsb_kernels.cuh"
#include "qsb_stream.cuh"

struct qsb_ctx {
  int device;
  cudaStream_t stream;
  bool own_stream;
  cudaEvent_t ev0, ev1;
  int sm_count, cc_major, cc_minor;
  size_t total_mem;
  int64_t launches;
  std::string err;
  uint64_t* d_masks;      // scratch for qsb_masked_parity
  c128* d_part;                 // partial sums of the large-state reductions
  int* d_bits;                  // bit lists of qsb_rdm_general
  int amp_bytes;                // 16: complex128 states (default); 8: complex64 mode (qsb_ctx_set_precision)
  unsigned long long* d_prof;   // cycle counters of the last qsb_run (qsb_debug_profile), or NULL
  int prof_ctas;
  cudaMemPool_t pool = nullptr; // device buffers come from here (stream-ordered, freed blocks stay cached)
};

struct qsb_buffer {
  qsb_ctx* ctx;
  void* ptr;
  int64_t bytes;
  bool owned;
};

struct qsb_program {
  qsb_ctx* ctx;
  int32_t n, m;
  int64_t n_ops, ops_stride, n_programs;
  qsb_op* d_ops;
  double* d_cdata;
  int32_t* d_idata;
  int32_t load_perm, store_perm, n_snapshots;
  int64_t n_idata, n_cdata;
  int32_t max_param, max_draw;   // highest parameter / draw index any op touches (argument checks)
  bool has_param;
  int32_t amp_bytes;             // element size of the states this program runs on (the ctx precision at creation)
  int32_t tile_bits;             // > 0: streaming mode (one CTA per 2^m-amplitude tile of a state in HBM)
};

static thread_local std::string g_err;

static int fail(qsb_ctx* ctx, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  if (ctx) ctx->err = buf;
  return code;
}

#define CU(ctx, call)                                                                          \
  do {                                                                                         \
    cudaError_t e_ = (call);                                                                   \
    if (e_ != cudaSuccess)                                                                     \
      return fail(ctx, e_ == cudaErrorMemoryAllocation ? QSB_E_OOM : QSB_E_CUDA, "%s: %s", #call, \
                  cudaGetErrorString(e_));                                                     \
  } while (0)

extern "C" {

int qsb_version(void) { return QSB_VERSION; }

int qsb_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    fail(nullptr, QSB_E_NODEV, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return QSB_E_NODEV;
  }
  return n;
}

const char* qsb_last_error(qsb_ctx* ctx) { return ctx ? ctx->err.c_str() : g_err.c_str(); }

int qsb_ctx_create(int device, qsb_ctx** out) {
  if (!out) return fail(nullptr, QSB_E_INVAL, "qsb_ctx_create: out is NULL");
  *out = nullptr;
  int n = qsb_device_count();
  if (n <= 0) return fail(nullptr, QSB_E_NODEV, "no CUDA device visible (libqsb has no CPU fallback)");
  if (device < 0 || device >= n) return fail(nullptr, QSB_E_NODEV, "device %d out of range [0, %d)", device, n);
  CU(nullptr, cudaSetDevice(device));
  cudaDeviceProp prop;
  CU(nullptr, cudaGetDeviceProperties(&prop, device));
  qsb_ctx* c = new qsb_ctx();
  c->device = device;
  c->own_stream = true;
  c->sm_count = prop.multiProcessorCount;
  c->cc_major = prop.major;
  c->cc_minor = prop.minor;
  c->total_mem = prop.totalGlobalMem;
  c->launches = 0;
  c->d_masks = nullptr;
  c->d_prof = nullptr;
  c->d_part = nullptr;
  c->d_bits = nullptr;
  c->amp_bytes = 16;
  c->prof_ctas = 0;
  cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreate(&c->ev0);
  if (e == cudaSuccess) e = cudaEventCreate(&c->ev1);
  if (e == cudaSuccess) e = cudaMalloc(&c->d_masks, 8 * sizeof(uint64_t));
  if (e == cudaSuccess) {                      // private stream-ordered pool that keeps freed blocks mapped
    cudaMemPoolProps pp = {};
    pp.allocType = cudaMemAllocationTypePinned;
    pp.handleTypes = cudaMemHandleTypeNone;
    pp.location.type = cudaMemLocationTypeDevice;
    pp.location.id = device;
    e = cudaMemPoolCreate(&c->pool, &pp);
    if (e == cudaSuccess) {
      uint64_t keep = UINT64_MAX;
      e = cudaMemPoolSetAttribute(c->pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
  }
  if (e != cudaSuccess) {
    delete c;
    return fail(nullptr, QSB_E_CUDA, "context setup: %s", cudaGetErrorString(e));
  }
  *out = c;
  return QSB_OK;
}

int qsb_ctx_destroy(qsb_ctx* ctx) {
  if (!ctx) return QSB_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  cudaEventDestroy(ctx->ev0);
  cudaEventDestroy(ctx->ev1);
  cudaFree(ctx->d_masks);
  cudaFree(ctx->d_prof);
  cudaFree(ctx->d_part);
  cudaFree(ctx->d_bits);
  if (ctx->pool) cudaMemPoolDestroy(ctx->pool);
  delete ctx;
  return QSB_OK;
}

int qsb_ctx_set_stream(qsb_ctx* ctx, void* cuda_stream) {
  if (!ctx) return fail(nullptr, QSB_E_INVAL, "ctx is NULL");
  CU(ctx, cudaSetDevice(ctx->device));
  CU(ctx, cudaStreamSynchronize(ctx->stream));
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  ctx->stream = (cudaStream_t)cuda_stream;
  ctx->own_stream = false;
  return QSB_OK;
}

int qsb_ctx_sync(qsb_ctx* ctx) {
  if (!ctx) return fail(nullptr, QSB_E_INVAL, "ctx is NULL");
  CU(ctx, cudaSetDevice(ctx->device));
  CU(ctx, cudaStreamSynchronize(ctx->stream));
  return QSB_OK;
}

int qsb_ctx_set_precision(qsb_ctx* ctx, int precision) {
  if (!ctx) return fail(nullptr, QSB_E_INVAL, "ctx is NULL");
  if (precision != QSB_C128 && precision != QSB_C64) return fail(ctx, QSB_E_INVAL, "precision must be QSB_C128 or QSB_C64");
  ctx->amp_bytes = precision == QSB_C64 ? 8 : 16;
  return QSB_OK;
}

int qsb_ctx_info(qsb_ctx* ctx, int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor, int64_t* total_mem) {
  if (!ctx) return fail(nullptr, QSB_E_INVAL, "ctx is NULL");
  if (sm_count) *sm_count = ctx->sm_count;
  if (cc_major) *cc_major = ctx->cc_major;
  if (cc_minor) *cc_minor = ctx->cc_minor;
  if (total_mem) *total_mem = (int64_t)ctx->total_mem;
  return QSB_OK;
}

int qsb_timer_start(qsb_ctx* ctx) {
  if (!ctx) return fail(nullptr, QSB_E_INVAL, "ctx is NULL");
  CU(ctx, cudaSetDevice(ctx->device));
  CU(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
  return QSB_OK;
}

int qsb_timer_stop(qsb_ctx* ctx, float* ms_out) {
  if (!ctx || !ms_out) return fail(ctx, QSB_E_INVAL, "qsb_timer_stop: NULL argument");
  CU(ctx, cudaSetDevice(ctx->device));
  CU(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
  CU(ctx, cudaEventSynchronize(ctx->ev1));
  CU(ctx, cudaEventElapsedTime(ms_out, ctx->ev0, ctx->ev1));
  return QSB_OK;
}

int64_t qsb_launch_count(qsb_ctx* ctx) { return ctx ? ctx->launches : 0; }

// ---- buffers ---------------------------------------------------------------------------
int qsb_buffer_alloc(qsb_ctx* ctx, int64_t bytes, qsb_buffer** out) {
  if (!ctx || !out || bytes < 0) return fail(ctx, QSB_E_INVAL, "qsb_buffer_alloc: bad argument");
  CU(ctx, cudaSetDevice(ctx->device));
  // stream-ordered allocation from the context's pool, which keeps freed blocks (release threshold
  // raised in qsb_ctx_create): per-call state batches of a few GB neither page-fault nor stall in cudaFree
  void* p = nullptr;
  size_t want = bytes > 0 ? (size_t)bytes : 16;
  cudaError_t e = cudaMallocFromPoolAsync(&p, want, ctx->pool, ctx->stream);
  if (e != cudaSuccess) {                    // give cached blocks back to the driver and try once more
    cudaGetLastError();
    cudaStreamSynchronize(ctx->stream);
    cudaMemPoolTrimTo(ctx->pool, 0);
    e = cudaMallocFromPoolAsync(&p, want, ctx->pool, ctx->stream);
  }
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(ctx, QSB_E_OOM, "cudaMallocAsync(%lld bytes): %s", (long long)bytes, cudaGetErrorString(e));
  }
  *out = new qsb_buffer{ctx, p, bytes, true};
  return QSB_OK;
}

int qsb_ctx_trim(qsb_ctx* ctx) {
  if (!ctx) return fail(nullptr, QSB_E_INVAL, "ctx is NULL");
  CU(ctx, cudaSetDevice(ctx->device));
  CU(ctx, cudaStreamSynchronize(ctx->stream));
  CU(ctx, cudaMemPoolTrimTo(ctx->pool, 0));
  return QSB_OK;
}

int qsb_buffer_wrap(qsb_ctx* ctx, void* device_ptr, int64_t bytes, qsb_buffer** out) {
  if (!ctx || !out || !device_ptr || bytes < 0) return fail(ctx, QSB_E_INVAL, "qsb_buffer_wrap: bad argument");
  *out = new qsb_buffer{ctx, device_ptr, bytes, false};
  return QSB_OK;
}

int qsb_buffer_free(qsb_buffer* buf) {
  if (!buf) return QSB_OK;
  if (buf->owned) {
    cudaSetDevice(buf->ctx->device);
    cudaFreeAsync(buf->ptr, buf->ctx->stream);      // ordered after every kernel that used it on the ctx stream
  }
  delete buf;
  return QSB_OK;
}

static int check_range(qsb_buffer* buf, int64_t off, int64_t bytes, const char* what) {
  if (!buf) return fail(nullptr, QSB_E_INVAL, "%s: buffer is NULL", what);
  if (off < 0 || bytes < 0 || off + bytes > buf->bytes)
    return fail(buf->ctx, QSB_E_INVAL, "%s: range [%lld, %lld) outside buffer of %lld bytes", what, (long long)off,
                (long long)(off + bytes), (long long)buf->bytes);
  return QSB_OK;
}

int qsb_buffer_upload(qsb_buffer* buf, int64_t offset, const void* host, int64_t bytes) {
  int rc = check_range(buf, offset, bytes, "qsb_buffer_upload");
  if (rc) return rc;
  if (!host && bytes) return fail(buf->ctx, QSB_E_INVAL, "qsb_buffer_upload: host is NULL");
  qsb_ctx* ctx = buf->ctx;
  CU(ctx, cudaSetDevice(ctx->device));
  CU(ctx, cudaMemcpyAsync((char*)buf->ptr + offset, host, (size_t)bytes, cudaMemcpyHostToDevice, ctx->stream));
  CU(ctx, cudaStreamSynchronize(ctx->stream));
  return QSB_OK;
}

int qsb_buffer_upload_async(qsb_buffer* buf, int64_t offset, const void* host, int64_t bytes) {
  int rc = check_range(buf, offset, bytes, "qsb_buffer_upload_async");
  if (rc) return rc;
  if (!host && bytes) return fail(buf->ctx, QSB_E_INVAL, "qsb_buffer_upload_async: host is NULL");
  qsb_ctx* ctx = buf->ctx;
  CU(ctx, cudaSetDevice(ctx->device));
  CU(ctx, cudaMemcpyAsync((char*)buf->ptr + offset, host, (size_t)bytes, cudaMemcpyHostToDevice, ctx->stream));
  return QSB_OK;
}

struct qsb_event {
  qsb_ctx* ctx;
  cudaEvent_t ev;
};

int qsb_event_create(qsb_ctx* ctx, qsb_event** out) {
  if (!ctx || !out) return fail(ctx, QSB_E_INVAL, "qsb_event_create: bad argument");
  CU(ctx, cudaSetDevice(ctx->device));
  cudaEvent_t e;
  CU(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  *out = new qsb_event{ctx, e};
  return QSB_OK;
}

int qsb_event_record(qsb_event* ev) {
  if (!ev) return fail(nullptr, QSB_E_INVAL, "event is NULL");
  CU(ev->ctx, cudaSetDevice(ev->ctx->device));
  CU(ev->ctx, cudaEventRecord(ev->ev, ev->ctx->stream));
  return QSB_OK;
}

int qsb_event_wait(qsb_event* ev) {
  if (!ev) return fail(nullptr, QSB_E_INVAL, "event is NULL");
  CU(ev->ctx, cudaEventSynchronize(ev->ev));
  return QSB_OK;
}

int qsb_event_free(qsb_event* ev) {
  if (!ev) return QSB_OK;
  cudaEventDestroy(ev->ev);
  delete ev;
  return QSB_OK;
}

int qsb_buffer_download(qsb_buffer* buf, int64_t offset, void* host, int64_t bytes) {
  int rc = check_range(buf, offset, bytes, "qsb_buffer_download");
  if (rc) return rc;
  if (!host && bytes) return fail(buf->ctx, QSB_E_INVAL, "qsb_buffer_download: host is NULL");
  qsb_ctx* ctx = buf->ctx;
  CU(ctx, cudaSetDevice(ctx->device));
  CU(ctx, cudaMemcpyAsync(host, (char*)buf->ptr + offset, (size_t)bytes, cudaMemcpyDeviceToHost, ctx->stream));
  CU(ctx, cudaStreamSynchronize(ctx->stream));
  return QSB_OK;
}

int qsb_buffer_zero(qsb_buffer* buf, int64_t offset, int64_t bytes) {
  int rc = check_range(buf, offset, bytes, "qsb_buffer_zero");
  if (rc) return rc;
  qsb_ctx* ctx = buf->ctx;
  CU(ctx, cudaSetDevice(ctx->device));
  CU(ctx, cudaMemsetAsync((char*)buf->ptr + offset, 0, (size_t)bytes, ctx->stream));
  return QSB_OK;
}

int qsb_buffer_copy(qsb_buffer* dst, int64_t dst_off, qsb_buffer* src, int64_t src_off, int64_t bytes) {
  int rc = check_range(dst, dst_off, bytes, "qsb_buffer_copy(dst)");
  if (rc) return rc;
  rc = check_range(src, src_off, bytes, "qsb_buffer_copy(src)");
  if (rc) return rc;
  qsb_ctx* ctx = dst->ctx;
  CU(ctx, cudaSetDevice(ctx->device));
  CU(ctx, cudaMemcpyAsync((char*)dst->ptr + dst_off, (char*)src->ptr + src_off, (size_t)bytes, cudaMemcpyDeviceToDevice,
                          ctx->stream));
  return QSB_OK;
}

void* qsb_buffer_ptr(qsb_buffer* buf) { return buf ? buf->ptr : nullptr; }
int64_t qsb_buffer_bytes(qsb_buffer* buf) { return buf ? buf->bytes : 0; }

int qsb_host_alloc(int64_t bytes, void** out) {
  if (!out || bytes < 0) return fail(nullptr, QSB_E_INVAL, "qsb_host_alloc: bad argument");
  CU(nullptr, cudaHostAlloc(out, bytes > 0 ? (size_t)bytes : 16, cudaHostAllocDefault));
  return QSB_OK;
}

int qsb_host_free(void* p) {
  if (p) CU(nullptr, cudaFreeHost(p));
  return QSB_OK;
}

// ---- programs --------------------------------------------------------------------------
static bool is_perm(const int32_t* p, int n) {
  unsigned seen = 0;
  for (int i = 0; i < n; ++i) {
    if (p[i] < 0 || p[i] >= n || (seen >> p[i]) & 1) return false;
    seen |= 1u << p[i];
  }
  return true;
}

int qsb_program_create(qsb_ctx* ctx, int32_t n, int32_t m, const qsb_op* ops, int64_t n_ops, int64_t ops_stride,
                       int64_t n_programs, const double* cdata, int64_t n_cdata, const int32_t* idata, int64_t n_idata,
                       int32_t load_perm, int32_t store_perm, int32_t n_snapshots, qsb_program** out) {
  if (!ctx || !out) return fail(ctx, QSB_E_INVAL, "qsb_program_create: NULL argument");
  *out = nullptr;
  if (n < 1) return fail(ctx, QSB_E_INVAL, "num_qubits must be >= 1, got %d", n);
  if (n > 30)
    return fail(ctx, QSB_E_UNSUPPORTED, "executor holds n <= 30 qubits per device, got %d", n);
  const int max_m = ctx->amp_bytes == 8 ? QSB_MAX_LOCAL_BITS_C64 : QSB_MAX_LOCAL_BITS;
  if (m < 1 || m > n || m > max_m)
    return fail(ctx, QSB_E_INVAL, "local_bits %d invalid for n = %d (need 1 <= m <= min(n, %d))", m, n, max_m);
  // n - m <= 3 and n <= 16: the state is resident in a cluster of 2^(n-m) CTAs; otherwise it is streamed
  // through shared memory tile by tile (one pass over HBM per program)
  const bool streaming = (n - m > 3) || n > QSB_MAX_QUBITS;
  if (streaming && n_snapshots > 0)
    return fail(ctx, QSB_E_UNSUPPORTED, "snapshots are not available in streaming mode (n - local_bits > 3)");
  if (n_ops < 0 || (n_ops > 0 && !ops) || n_programs < 1 || (ops_stride != 0 && ops_stride < n_ops))
    return fail(ctx, QSB_E_INVAL, "bad op list");
  if (n_idata < 0 || load_perm < 0 || store_perm < 0 || load_perm + n > n_idata || store_perm + n > n_idata)
    return fail(ctx, QSB_E_INVAL, "load/store permutation outside idata");
  if (!is_perm(idata + load_perm, n) || !is_perm(idata + store_perm, n))
    return fail(ctx, QSB_E_INVAL, "load/store entries are not permutations of 0..n-1");
  const int64_t total_ops = ops_stride ? ops_stride * n_programs : n_ops;
  int32_t max_param = -1, max_draw = -1;
  bool has_param = false;
  const int gbits = streaming ? 0 : n - m;
  for (int64_t i = 0; i < total_ops; ++i) {
    const qsb_op& o = ops[i];
    int nb = 0, need = 0;
    if (streaming && (o.kind == QSB_OP_KRAUS_AD || o.kind == QSB_OP_KRAUS_GEN || o.kind == QSB_OP_REMAP ||
                      o.kind == QSB_OP_SNAPSHOT))
      return fail(ctx, QSB_E_UNSUPPORTED, "op %lld: kind %d needs the whole state resident (streaming mode)",
                  (long long)i, o.kind);
    switch (o.kind) {
      case QSB_OP_NOP: break;
      case QSB_OP_U1: nb = 1; need = 8; break;
      case QSB_OP_D1: nb = 1; need = 4; break;
      case QSB_OP_U2: nb = 2; need = 32; break;
      case QSB_OP_U3Q: nb = 3; need = 128; break;
      case QSB_OP_X: case QSB_OP_Y: case QSB_OP_Z: nb = 1; break;
      case QSB_OP_CX: case QSB_OP_CZ: case QSB_OP_SWAP: nb = 2; break;
      case QSB_OP_CCX: case QSB_OP_CSWAP: nb = 3; break;
      case QSB_OP_RX: case QSB_OP_RY: case QSB_OP_RZ: case QSB_OP_PHASE: case QSB_OP_U3:
        nb = 1;
        has_param = true;
        if (o.param < 0) return fail(ctx, QSB_E_INVAL, "op %lld: parameterised gate without param index", (long long)i);
        if (o.param + (o.kind == QSB_OP_U3 ? 2 : 0) > max_param) max_param = o.param + (o.kind == QSB_OP_U3 ? 2 : 0);
        break;
      case QSB_OP_KRAUS_PAULI: nb = 1; need = 7; break;
      case QSB_OP_KRAUS_AD: nb = 1; need = 3; break;
      case QSB_OP_KRAUS_GEN: nb = 1; need = 1; break;
      case QSB_OP_REMAP:
        if (o.b0 < 0 || o.b0 >= gbits || o.b1 < 0 || o.b1 >= m || o.b2 < 0 || o.b2 > 2)
          return fail(ctx, QSB_E_INVAL, "op %lld: remap bits (%d, %d) invalid", (long long)i, o.b0, o.b1);
        {
          // further pairs of the same exchange: g1 | l1 << 8 | g2 << 16 | l2 << 24; all rank bits and all local
          // bits of one op must be distinct
          int gs[3] = {o.b0, o.aux & 255, (o.aux >> 16) & 255}, ls[3] = {o.b1, (o.aux >> 8) & 255, (o.aux >> 24) & 255};
          for (int j = 1; j <= o.b2; ++j) {
            if (gs[j] >= gbits || ls[j] >= m)
              return fail(ctx, QSB_E_INVAL, "op %lld: remap pair %d (%d, %d) invalid", (long long)i, j, gs[j], ls[j]);
            for (int q = 0; q < j; ++q)
              if (gs[q] == gs[j] || ls[q] == ls[j])
                return fail(ctx, QSB_E_INVAL, "op %lld: remap pairs %d and %d overlap", (long long)i, q, j);
          }
        }
        break;
      case QSB_OP_SNAPSHOT:
        if (o.b0 < 0 || o.b0 >= n_snapshots || o.aux < 0 || o.aux + n > n_idata || !is_perm(idata + o.aux, n))
          return fail(ctx, QSB_E_INVAL, "op %lld: bad snapshot slot / permutation", (long long)i);
        break;
      default:
        return fail(ctx, QSB_E_INVAL, "op %lld: unknown kind %d", (long long)i, o.kind);
    }
    const int bits[3] = {o.b0, o.b1, o.b2};
    if (o.kind != QSB_OP_REMAP && o.kind != QSB_OP_SNAPSHOT) {
      // 1-qubit gates and Pauli / amplitude-damping draws only touch the pending matrix of their slot, which
      // may be a cluster-rank slot; everything that sweeps the tile needs local slots
      const bool deferred = nb == 1 && o.kind != QSB_OP_KRAUS_GEN;
      const int limit = deferred ? m + gbits : m;
      for (int k = 0; k < nb; ++k) {
        if (bits[k] < 0 || bits[k] >= limit)
          return fail(ctx, QSB_E_INVAL, "op %lld: target bit %d not resident (local_bits = %d)", (long long)i, bits[k], m);
        for (int l = 0; l < k; ++l)
          if (bits[l] == bits[k]) return fail(ctx, QSB_E_INVAL, "op %lld: repeated target bit", (long long)i);
      }
    }
    if (need > 0) {
      if (o.data < 0 || o.data + need > n_cdata) return fail(ctx, QSB_E_INVAL, "op %lld: cdata out of range", (long long)i);
      if (o.kind == QSB_OP_KRAUS_GEN) {
        int nk = (int)cdata[o.data];
        if (nk < 1 || nk > 8 || o.data + 1 + 12 * nk > n_cdata)
          return fail(ctx, QSB_E_INVAL, "op %lld: bad Kraus set", (long long)i);
      }
      if ((o.kind == QSB_OP_U2 || o.kind == QSB_OP_U3Q) && (o.data & 1))
        return fail(ctx, QSB_E_INVAL, "op %lld: dense matrix must be 16-byte aligned in cdata", (long long)i);
    }
    if (o.kind >= QSB_OP_KRAUS_PAULI && o.kind <= QSB_OP_KRAUS_GEN) {
      if (o.draw < 0) return fail(ctx, QSB_E_INVAL, "op %lld: Kraus op without draw index", (long long)i);
      if (o.draw > max_draw) max_draw = o.draw;
    }
  }
  CU(ctx, cudaSetDevice(ctx->device));
  qsb_program* p = new qsb_program();
  p->ctx = ctx;
  p->n = n;
  p->m = m;
  p->n_ops = n_ops;
  p->ops_stride = ops_stride;
  p->n_programs = n_programs;
  p->load_perm = load_perm;
  p->store_perm = store_perm;
  p->n_snapshots = n_snapshots;
  p->n_idata = n_idata;
  p->n_cdata = n_cdata;
  p->max_param = max_param;
  p->max_draw = max_draw;
  p->has_param = has_param;
  p->tile_bits = streaming ? n - m : 0;
  p->amp_bytes = ctx->amp_bytes;
  p->d_ops = nullptr;
  p->d_cdata = nullptr;
  p->d_idata = nullptr;
  // Stream-ordered allocations from the context's pool and no synchronisation: the batch drivers (QEC sweeps) build
  // two programs of a few MB per batch, and a cudaMalloc / cudaFree / stream-sync per program serialised the host
  // with the GPU (5 ms per program freed).  The host arrays are pageable: cudaMemcpyAsync returns once they are staged.
  cudaError_t e = cudaMallocFromPoolAsync((void**)&p->d_ops, sizeof(qsb_op) * (size_t)(total_ops > 0 ? total_ops : 1), ctx->pool, ctx->stream);
  if (e == cudaSuccess) e = cudaMallocFromPoolAsync((void**)&p->d_cdata, sizeof(double) * (size_t)(n_cdata > 0 ? n_cdata : 2), ctx->pool, ctx->stream);
  if (e == cudaSuccess) e = cudaMallocFromPoolAsync((void**)&p->d_idata, sizeof(int32_t) * (size_t)(n_idata > 0 ? n_idata : 1), ctx->pool, ctx->stream);
  if (e == cudaSuccess && total_ops)
    e = cudaMemcpyAsync(p->d_ops, ops, sizeof(qsb_op) * (size_t)total_ops, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess && n_cdata)
    e = cudaMemcpyAsync(p->d_cdata, cdata, sizeof(double) * (size_t)n_cdata, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess && n_idata)
    e = cudaMemcpyAsync(p->d_idata, idata, sizeof(int32_t) * (size_t)n_idata, cudaMemcpyHostToDevice, ctx->stream);
  if (e != cudaSuccess) {
    cudaGetLastError();
    cudaStreamSynchronize(ctx->stream);
    if (p->d_ops) cudaFreeAsync(p->d_ops, ctx->stream);
    if (p->d_cdata) cudaFreeAsync(p->d_cdata, ctx->stream);
    if (p->d_idata) cudaFreeAsync(p->d_idata, ctx->stream);
    delete p;
    cudaGetLastError();
    return fail(ctx, e == cudaErrorMemoryAllocation ? QSB_E_OOM : QSB_E_CUDA, "program upload: %s", cudaGetErrorString(e));
  }
  *out = p;
  return QSB_OK;
}

int qsb_program_free(qsb_program* p) {
  if (!p) return QSB_OK;
  cudaSetDevice(p->ctx->device);
  cudaFreeAsync(p->d_ops, p->ctx->stream);          // ordered after every launch that read them on the ctx stream
  cudaFreeAsync(p->d_cdata, p->ctx->stream);
  cudaFreeAsync(p->d_idata, p->ctx->stream);
  delete p;
  return QSB_OK;
}

}  // extern "C"

template <int C, class A, bool PF>
static cudaError_t launch_traj_pf(qsb_ctx* ctx, const qsb_exec_args& a, int threads, size_t smem, int* grid_out) {
  auto kern = qsb_traj_kernel<C, A, PF>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.blockDim = dim3(threads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = ctx->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int64_t max_units;
  if (C > 1) {
    int nclusters = 0;
    cfg.gridDim = dim3(C * ctx->sm_count, 1, 1);
    e = cudaOccupancyMaxActiveClusters(&nclusters, kern, &cfg);
    if (e != cudaSuccess) return e;
    if (nclusters < 1) return cudaErrorLaunchOutOfResources;
    max_units = nclusters;
  } else {
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    max_units = (int64_t)per_sm * ctx->sm_count;
  }
  const int64_t total_units = a.count << a.tile_bits;
  int64_t units = total_units < max_units ? total_units : max_units;
  cfg.gridDim = dim3((unsigned)(units * C), 1, 1);
  *grid_out = (int)(units * C);
  return cudaLaunchKernelEx(&cfg, kern, a);
}

// the profiling instantiation (cycle counters compiled in) is launched only while qsb_debug_profile is enabled
template <int C, class A>
static cudaError_t launch_traj(qsb_ctx* ctx, const qsb_exec_args& a, int threads, size_t smem, int* grid_out) {
  return a.prof ? launch_traj_pf<C, A, true>(ctx, a, threads, smem, grid_out)
                : launch_traj_pf<C, A, false>(ctx, a, threads, smem, grid_out);
}

extern "C" {

static int need(qsb_ctx* ctx, qsb_buffer* b, int64_t bytes, const char* what) {
  if (!b) return fail(ctx, QSB_E_INVAL, "%s buffer is NULL", what);
  if (b->bytes < bytes)
    return fail(ctx, QSB_E_INVAL, "%s buffer too small: %lld < %lld bytes", what, (long long)b->bytes, (long long)bytes);
  return QSB_OK;
}

int qsb_run(qsb_program* p, const qsb_run_args* r) {
  if (!p || !r) return fail(nullptr, QSB_E_INVAL, "qsb_run: NULL argument");
  qsb_ctx* ctx = p->ctx;
  if (r->count < 0 || r->first < 0) return fail(ctx, QSB_E_INVAL, "qsb_run: negative range");
  if (r->count == 0) return QSB_OK;
  const int64_t dim = (int64_t)1 << p->n;
  const int64_t AB = p->amp_bytes;
  int rc;
  if (ctx->amp_bytes != p->amp_bytes)
    return fail(ctx, QSB_E_INVAL, "qsb_run: the program was created in another precision mode than the context is in now");
  if ((r->flags & QSB_RUN_LOAD_BROADCAST) && (r->flags & QSB_RUN_STORE) && !r->states_out)
    return fail(ctx, QSB_E_INVAL, "qsb_run: LOAD_BROADCAST cannot store in place");
  if ((r->flags & QSB_RUN_LOAD) || ((r->flags & QSB_RUN_STORE) && !r->states_out))
    if ((rc = need(ctx, r->states, (r->first + ((r->flags & QSB_RUN_LOAD_BROADCAST) ? 1 : r->count)) * dim * AB, "states")))
      return rc;
  if (p->ops_stride && r->count > p->n_programs)
    return fail(ctx, QSB_E_INVAL, "qsb_run: %lld trajectories but only %lld per-trajectory programs",
                (long long)r->count, (long long)p->n_programs);
  if (p->has_param) {
    if (r->params_stride <= p->max_param) return fail(ctx, QSB_E_INVAL, "qsb_run: params_stride too small");
    if ((rc = need(ctx, r->params, r->count * r->params_stride * 8, "params"))) return rc;
  }
  if (r->uniforms && p->max_draw >= 0) {
    if (r->uniforms_stride <= p->max_draw) return fail(ctx, QSB_E_INVAL, "qsb_run: uniforms_stride too small");
    if ((rc = need(ctx, r->uniforms, r->count * r->uniforms_stride * 8, "uniforms"))) return rc;
  }
  if (r->branches && p->max_draw >= 0) {
    if (r->branches_stride <= p->max_draw) return fail(ctx, QSB_E_INVAL, "qsb_run: branches_stride too small");
    if ((rc = need(ctx, r->branches, r->count * r->branches_stride * 4, "branches"))) return rc;
  }
  if (r->init_basis && (rc = need(ctx, r->init_basis, r->count * 8, "init_basis"))) return rc;
  if (!r->init_basis && (r->default_basis < 0 || r->default_basis >= dim) && !(r->flags & QSB_RUN_LOAD))
    return fail(ctx, QSB_E_INVAL, "qsb_run: default_basis out of range");
  if (p->n_snapshots > 0 && r->snapshots &&
      (rc = need(ctx, r->snapshots, r->count * p->n_snapshots * dim * AB, "snapshots")))
    return rc;
  if ((r->flags & QSB_RUN_ACCUM_PROBS) && (rc = need(ctx, r->probs_accum, dim * 8, "probs_accum"))) return rc;

  qsb_exec_args a;
  memset(&a, 0, sizeof a);
  a.ops = p->d_ops;
  a.n_ops = p->n_ops;
  a.ops_stride = p->ops_stride;
  a.cdata = p->d_cdata;
  a.n_cdata = p->n_cdata;
  a.idata = p->d_idata;
  a.n = p->n;
  a.m = p->m;
  a.load_perm = p->load_perm;
  a.store_perm = p->store_perm;
  a.n_snapshots = p->n_snapshots;
  a.flags = r->flags;
  a.amp_bytes = AB;
  a.states = r->states ? (char*)r->states->ptr + r->first * dim * AB : nullptr;
  a.states_out = a.states;
  if (r->states_out) {
    if (r->out_first < 0) return fail(ctx, QSB_E_INVAL, "qsb_run: negative out_first");
    if ((rc = need(ctx, r->states_out, (r->out_first + r->count) * dim * AB, "states_out"))) return rc;
    a.states_out = (char*)r->states_out->ptr + r->out_first * dim * AB;
  }
  a.count = r->count;
  a.params = r->params ? (const double*)r->params->ptr : nullptr;
  a.params_stride = r->params_stride;
  a.uniforms = r->uniforms ? (const double*)r->uniforms->ptr : nullptr;
  a.uniforms_stride = r->uniforms_stride;
  a.seed = r->philox_seed;
  a.traj_offset = r->traj_offset;
  a.init_basis = r->init_basis ? (const int64_t*)r->init_basis->ptr : nullptr;
  a.default_basis = r->default_basis;
  a.branches = r->branches ? (int32_t*)r->branches->ptr : nullptr;
  a.branches_stride = r->branches_stride;
  a.snapshots = (p->n_snapshots > 0 && r->snapshots) ? r->snapshots->ptr : nullptr;
  a.probs_accum = (r->flags & QSB_RUN_ACCUM_PROBS) ? (double*)r->probs_accum->ptr : nullptr;

  a.peer_ptrs = nullptr; a.peer_shift = 0; a.peer_rank_or = 0;
  if (r->peer_table) {
    if (!p->tile_bits || !(r->flags & QSB_RUN_LOAD) || !r->states_out || r->count != 1)
      return fail(ctx, QSB_E_INVAL, "qsb_run: peer_table needs a streamed pass with LOAD, states_out and count = 1");
    if (r->peer_shift < 1 || r->peer_shift > p->n || r->peer_table->bytes < (int64_t)(sizeof(void*) << (p->n - r->peer_shift)))
      return fail(ctx, QSB_E_INVAL, "qsb_run: bad peer_shift / peer_table size");
    a.peer_ptrs = (const void* const*)r->peer_table->ptr;
    a.peer_shift = r->peer_shift;
    a.peer_rank_or = r->peer_rank_or;
  }
  a.prof = ctx->d_prof;
  if (ctx->d_prof) CU(ctx, cudaMemsetAsync(ctx->d_prof, 0, (size_t)8 * ctx->sm_count * 4 * QSB_PROF_WORDS * sizeof(unsigned long long), ctx->stream));
  a.tile_bits = p->tile_bits;
  if (p->tile_bits) {
    if (r->flags & QSB_RUN_NORMALIZE)
      return fail(ctx, QSB_E_UNSUPPORTED, "qsb_run: normalisation needs the whole state resident (streaming mode)");
    if ((r->flags & QSB_RUN_ACCUM_PROBS) || a.branches)
      return fail(ctx, QSB_E_UNSUPPORTED, "qsb_run: probs_accum / branches are not available in streaming mode");
  }

  CU(ctx, cudaSetDevice(ctx->device));
  const int C = p->tile_bits ? 1 : 1 << (p->n - p->m);
  int workers = 1 << (p->m > 4 ? p->m - 4 : 1);       // at least 16 amplitudes per worker per pass
  if (workers < 32) workers = 32;
  if (workers > QSB_MAX_WORKERS) workers = QSB_MAX_WORKERS;
  const int threads = workers + QSB_CTL_THREADS + QSB_DEC_THREADS;
  const bool c64m = p->amp_bytes == 8;
  const size_t smem = ((size_t)p->amp_bytes << p->m) + QSB_SMEM_EXTRA;
  int grid = 0;
  cudaError_t e;
  switch (C) {
    case 1: e = c64m ? launch_traj<1, c64>(ctx, a, threads, smem, &grid) : launch_traj<1, c128>(ctx, a, threads, smem, &grid); break;
    case 2: e = c64m ? launch_traj<2, c64>(ctx, a, threads, smem, &grid) : launch_traj<2, c128>(ctx, a, threads, smem, &grid); break;
    case 4: e = c64m ? launch_traj<4, c64>(ctx, a, threads, smem, &grid) : launch_traj<4, c128>(ctx, a, threads, smem, &grid); break;
    case 8: e = c64m ? launch_traj<8, c64>(ctx, a, threads, smem, &grid) : launch_traj<8, c128>(ctx, a, threads, smem, &grid); break;
    default: return fail(ctx, QSB_E_UNSUPPORTED, "cluster size %d", C);
  }
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(ctx, QSB_E_CUDA, "trajectory kernel launch (C=%d, threads=%d, smem=%zu): %s", C, threads, smem,
                cudaGetErrorString(e));
  }
  ctx->launches += 1;
  ctx->prof_ctas = grid;
  if (!(r->flags & QSB_RUN_ASYNC)) CU(ctx, cudaStreamSynchronize(ctx->stream));
  return QSB_OK;
}

// Developer aid: enable (enable != 0) per-CTA cycle counters for subsequent qsb_run calls, and/or read the
// counters of the last run: out[cta][32] = {worker wait, busy per descriptor kind [8], count per kind [8],
// control ring-full wait, control total, descriptors}.  Returns the number of CTAs of the last run.
int qsb_debug_profile(qsb_ctx* ctx, int enable, unsigned long long* out, int64_t max_ctas) {
  if (!ctx) return fail(nullptr, QSB_E_INVAL, "ctx is NULL");
  CU(ctx, cudaSetDevice(ctx->device));
  const int64_t cap = (int64_t)8 * ctx->sm_count * 4;   // every grid qsb_run can launch (<= 8 CTAs x resident units per SM) fits
  if (enable && !ctx->d_prof) {
    CU(ctx, cudaMalloc(&ctx->d_prof, cap * QSB_PROF_WORDS * sizeof(unsigned long long)));
    CU(ctx, cudaMemset(ctx->d_prof, 0, cap * QSB_PROF_WORDS * sizeof(unsigned long long)));
  }
  int n = 0;
  if (out && ctx->d_prof) {
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    n = ctx->prof_ctas < max_ctas ? ctx->prof_ctas : (int)max_ctas;
    CU(ctx, cudaMemcpy(out, ctx->d_prof, (size_t)n * QSB_PROF_WORDS * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  }
  if (!enable && ctx->d_prof) {
    cudaFree(ctx->d_prof);
    ctx->d_prof = nullptr;
  }
  return n;
}

// ---- reductions ------------------------------------------------------------------------
#define CHECK_N(ctx, n)                                                                       \
  if (!(ctx)) return fail(nullptr, QSB_E_INVAL, "ctx is NULL");                               \
  if ((n) < 1 || (n) > 30) return fail(ctx, QSB_E_INVAL, "num_qubits %d out of range", (int)(n))

static int after_launch(qsb_ctx* ctx, const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(ctx, QSB_E_CUDA, "%s launch: %s", what, cudaGetErrorString(e));
  ctx->launches += 1;
  e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) return fail(ctx, QSB_E_CUDA, "%s: %s", what, cudaGetErrorString(e));
  return QSB_OK;
}

// run STMT with `A` = the context's amplitude type (complex128, or complex64 in c64 mode)
#define QSB_BY_AMP(ctx, STMT)            \
  do {                                   \
    if ((ctx)->amp_bytes == 8) {         \
      typedef c64 A;                     \
      STMT;                              \
    } else {                             \
      typedef c128 A;                    \
      STMT;                              \
    }                                    \
  } while (0)

static int grid_for(qsb_ctx* ctx, int64_t items, int threads) {
  int64_t g = (items + threads - 1) / threads;
  int64_t cap = (int64_t)ctx->sm_count * 16;
  if (g > cap) g = cap;
  return g < 1 ? 1 : (int)g;
}

int qsb_probabilities(qsb_ctx* ctx, int32_t n, qsb_buffer* states, int64_t first, int64_t count, qsb_buffer* out,
                      int64_t out_first) {
  CHECK_N(ctx, n);
  const int64_t dim = (int64_t)1 << n;
  int rc;
  if (count <= 0) return count == 0 ? QSB_OK : fail(ctx, QSB_E_INVAL, "negative count");
  if (first < 0) return fail(ctx, QSB_E_INVAL, "negative first");
  if ((rc = need(ctx, states, (first + count) * dim * ctx->amp_bytes, "states"))) return rc;
  if (out_first < 0) return fail(ctx, QSB_E_INVAL, "negative out_first");
  if ((rc = need(ctx, out, (out_first + count) * dim * 8, "probabilities"))) return rc;
  CU(ctx, cudaSetDevice(ctx->device));
  QSB_BY_AMP(ctx, (qsb_probs_kernel<A><<<grid_for(ctx, count * dim, 256), 256, 0, ctx->stream>>>(
                      (const A*)states->ptr + first * dim, (double*)out->ptr + out_first * dim, count * dim)));
  return after_launch(ctx, "probabilities");
}

int qsb_probabilities_sum(qsb_ctx* ctx, int32_t n, qsb_buffer* states, int64_t first, int64_t count, qsb_buffer* out) {
  CHECK_N(ctx, n);
  const int64_t dim = (int64_t)1 << n;
  int rc;
  if (count <= 0) return count == 0 ? QSB_OK : fail(ctx, QSB_E_INVAL, "negative count");
  if (first < 0) return fail(ctx, QSB_E_INVAL, "negative first");
  if ((rc = need(ctx, states, (first + count) * dim * ctx->amp_bytes, "states"))) return rc;
  if ((rc = need(ctx, out, dim * 8, "probability sum"))) return rc;
  CU(ctx, cudaSetDevice(ctx->device));
  QSB_BY_AMP(ctx, (qsb_probs_sum_kernel<A><<<grid_for(ctx, dim, 128), 128, 0, ctx->stream>>>(
                      (const A*)states->ptr + first * dim, (double*)out->ptr, dim, count)));
  return after_launch(ctx, "probabilities_sum");
}

int qsb_sample_index(qsb_ctx* ctx, int32_t n, qsb_buffer* states, int64_t first, int64_t count, qsb_buffer* uniforms,
                     qsb_buffer* out) {
  CHECK_N(ctx, n);
  const int64_t dim = (int64_t)1 << n;
  int rc;
  if (count <= 0) return count == 0 ? QSB_OK : fail(ctx, QSB_E_INVAL, "negative count");
  if (first < 0) return fail(ctx, QSB_E_INVAL, "negative first");
  if ((rc = need(ctx, states, (first + count) * dim * ctx->amp_bytes, "states"))) return rc;
  if ((rc = need(ctx, uniforms, count * 8, "uniforms"))) return rc;
  if ((rc = need(ctx, out, count * 8, "sample output"))) return rc;
  CU(ctx, cudaSetDevice(ctx->device));
  int threads = dim >= 256 ? 256 : 32;
  QSB_BY_AMP(ctx, (qsb_sample_kernel<A><<<(unsigned)count, threads, 0, ctx->stream>>>(
                      (const A*)states->ptr + first * dim, (const double*)uniforms->ptr, (int64_t*)out->ptr, dim)));
  return after_launch(ctx, "sample_index");
}

int qsb_overlap(qsb_ctx* ctx, int32_t n, qsb_buffer* a, int64_t a_first, qsb_buffer* b, int64_t b_first,
                int64_t b_stride_states, int64_t count, qsb_buffer* out) {
  CHECK_N(ctx, n);
  const int64_t dim = (int64_t)1 << n;
  int rc;
  if (count <= 0) return count == 0 ? QSB_OK : fail(ctx, QSB_E_INVAL, "negative count");
  if (b_stride_states != 0 && b_stride_states != 1) return fail(ctx, QSB_E_INVAL, "b_stride_states must be 0 or 1");
  if (a_first < 0 || b_first < 0) return fail(ctx, QSB_E_INVAL, "negative first");
  if ((rc = need(ctx, a, (a_first + count) * dim * ctx->amp_bytes, "states a"))) return rc;
  if ((rc = need(ctx, b, (b_first + (b_stride_states ? count : 1)) * dim * ctx->amp_bytes, "states b"))) return rc;
  if ((rc = need(ctx, out, count * 16, "overlap output"))) return rc;
  CU(ctx, cudaSetDevice(ctx->device));
  if (dim >= ((int64_t)1 << 18)) {
    // big states: two-stage reduction over many CTAs, one state pair at a time
    const int n_part = ctx->sm_count * 8;
    const int64_t per = (dim + n_part - 1) / n_part;
    if (!ctx->d_part) CU(ctx, cudaMalloc(&ctx->d_part, sizeof(c128) * (size_t)n_part));
    for (int64_t t = 0; t < count; ++t) {
      QSB_BY_AMP(ctx, (qsb_overlap_partial_kernel<A><<<n_part, 256, 0, ctx->stream>>>(
                          (const A*)a->ptr + (a_first + t) * dim, (const A*)b->ptr + (b_first + t * b_stride_states) * dim,
                          dim, per, ctx->d_part)));
      qsb_overlap_final_kernel<<<1, 256, 0, ctx->stream>>>(ctx->d_part, n_part, (c128*)out->ptr + t);
      ctx->launches += 1;
    }
    return after_launch(ctx, "overlap");
  }
  QSB_BY_AMP(ctx, (qsb_overlap_kernel<A><<<(unsigned)count, 256, 0, ctx->stream>>>(
                      (const A*)a->ptr + a_first * dim, (const A*)b->ptr + b_first * dim, b_stride_states * dim,
                      (c128*)out->ptr, dim)));
  return after_launch(ctx, "overlap");
}

int qsb_masked_parity(qsb_ctx* ctx, int32_t n, qsb_buffer* states, int64_t first, int64_t count, const uint64_t* masks,
                      int32_t n_masks, qsb_buffer* out) {
  CHECK_N(ctx, n);
  const int64_t dim = (int64_t)1 << n;
  int rc;
  if (count <= 0) return count == 0 ? QSB_OK : fail(ctx, QSB_E_INVAL, "negative count");
  if (n_masks < 1 || n_masks > 8 || !masks) return fail(ctx, QSB_E_INVAL, "n_masks must be 1..8");
  if ((rc = need(ctx, states, (first + count) * dim * ctx->amp_bytes, "states"))) return rc;
  if ((rc = need(ctx, out, count * n_masks * 16, "parity output"))) return rc;
  CU(ctx, cudaSetDevice(ctx->device));
  CU(ctx, cudaMemcpyAsync(ctx->d_masks, masks, sizeof(uint64_t) * n_masks, cudaMemcpyHostToDevice, ctx->stream));
  QSB_BY_AMP(ctx, (qsb_parity_kernel<A><<<(unsigned)count, 256, 0, ctx->stream>>>(
                      (const A*)states->ptr + first * dim, dim, ctx->d_masks, n_masks, (double*)out->ptr)));
  return after_launch(ctx, "masked_parity");
}

int qsb_rdm_all(qsb_ctx* ctx, int32_t n, qsb_buffer* states, int64_t first, int64_t count, qsb_buffer* rdm1,
                qsb_buffer* rdm2) {
  CHECK_N(ctx, n);
  const int64_t dim = (int64_t)1 << n;
  const int npairs = n * (n - 1) / 2;
  int rc;
  if (count <= 0) return count == 0 ? QSB_OK : fail(ctx, QSB_E_INVAL, "negative count");
  if (first < 0) return fail(ctx, QSB_E_INVAL, "negative first");
  if ((rc = need(ctx, states, (first + count) * dim * ctx->amp_bytes, "states"))) return rc;
  CU(ctx, cudaSetDevice(ctx->device));
  const char* s = (const char*)states->ptr + first * dim * ctx->amp_bytes;
  if (rdm1 && (rc = need(ctx, rdm1, count * n * 4 * 16, "rdm1"))) return rc;
  if (rdm2 && npairs > 0 && (rc = need(ctx, rdm2, count * npairs * 16 * 16, "rdm2"))) return rc;
  if (n >= 4 && n <= 13 && !getenv("QSB_RDM_SCALAR")) {
    // one read of every state: all pairs as 8x8 real Grams on the FP64 tensor cores (qsb_rdm_gram_kernel)
    const size_t smem = (size_t)16 << n;
    static bool attr[2][64] = {{false}};
    const int which = ctx->amp_bytes == 8 ? 1 : 0;
    if (!attr[which][ctx->device & 63]) {
      if (which) CU(ctx, cudaFuncSetAttribute(qsb_rdm_gram_kernel<c64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 << 13));
      else CU(ctx, cudaFuncSetAttribute(qsb_rdm_gram_kernel<c128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 << 13));
      attr[which][ctx->device & 63] = true;
    }
    const int per_sm = (int)((200u << 10) / (smem + 1024)) < 1 ? 1 : (int)((200u << 10) / (smem + 1024));
    int64_t grid = (int64_t)ctx->sm_count * (per_sm > 4 ? 4 : per_sm);
    if (grid > count) grid = count;
    QSB_BY_AMP(ctx, (qsb_rdm_gram_kernel<A><<<(unsigned)grid, 256, smem, ctx->stream>>>(
                        (const A*)s, n, npairs, count, rdm1 ? (c128*)rdm1->ptr : nullptr,
                        (rdm2 && npairs > 0) ? (c128*)rdm2->ptr : nullptr)));
    return after_launch(ctx, "rdm_gram");
  }
  if (rdm1) {
    QSB_BY_AMP(ctx, (qsb_rdm1_kernel<A><<<(unsigned)(count * n), 256, 0, ctx->stream>>>((const A*)s, n, (c128*)rdm1->ptr)));
    if ((rc = after_launch(ctx, "rdm1"))) return rc;
  }
  if (rdm2 && npairs > 0) {
    QSB_BY_AMP(ctx, (qsb_rdm2_kernel<A><<<(unsigned)(count * npairs), 256, 0, ctx->stream>>>((const A*)s, n, npairs,
                                                                                          (c128*)rdm2->ptr)));
    if ((rc = after_launch(ctx, "rdm2"))) return rc;
  }
  return QSB_OK;
}

int qsb_rdm_general(qsb_ctx* ctx, int32_t n, qsb_buffer* states, int64_t first, int64_t count, const int32_t* keep_qubits,
                    int32_t k, qsb_buffer* out) {
  CHECK_N(ctx, n);
  if (k < 1 || k > 6 || k > n || !keep_qubits) return fail(ctx, QSB_E_INVAL, "keep 1..6 qubits, got %d", (int)k);
  const int64_t dim = (int64_t)1 << n;
  int rc;
  if (count <= 0) return count == 0 ? QSB_OK : fail(ctx, QSB_E_INVAL, "negative count");
  if (first < 0) return fail(ctx, QSB_E_INVAL, "negative first");
  if ((rc = need(ctx, states, (first + count) * dim * ctx->amp_bytes, "states"))) return rc;
  if ((rc = need(ctx, out, count * ((int64_t)1 << (2 * k)) * 16, "rdm"))) return rc;
  int host[40];                                     // kept index bits (qubit q = bit n-1-q), then the environment bits
  unsigned used = 0;
  for (int j = 0; j < k; ++j) {
    const int q = keep_qubits[j];
    if (q < 0 || q >= n || ((used >> q) & 1u)) return fail(ctx, QSB_E_INVAL, "bad keep_qubits entry %d", q);
    if (j > 0 && q <= keep_qubits[j - 1]) return fail(ctx, QSB_E_INVAL, "keep_qubits must be ascending");
    used |= 1u << q;
    host[j] = n - 1 - q;
  }
  int ne = 0;
  for (int b = 0; b < n; ++b)
    if (!((used >> (n - 1 - b)) & 1u)) host[8 + ne++] = b;
  CU(ctx, cudaSetDevice(ctx->device));
  if (!ctx->d_bits) CU(ctx, cudaMalloc(&ctx->d_bits, 40 * sizeof(int)));
  CU(ctx, cudaMemcpyAsync(ctx->d_bits, host, 40 * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  QSB_BY_AMP(ctx, (qsb_rdm_general_kernel<A><<<(unsigned)count, 256, 0, ctx->stream>>>(
                      (const A*)states->ptr + first * dim, n, k, ctx->d_bits, ctx->d_bits + 8, (c128*)out->ptr)));
  return after_launch(ctx, "rdm_general");
}

int qsb_mi_all_pairs(qsb_ctx* ctx, int32_t n, qsb_buffer* states, int64_t first, int64_t count, qsb_buffer* mi,
                     qsb_buffer* entropy1) {
  CHECK_N(ctx, n);
  if (n < 2) return fail(ctx, QSB_E_INVAL, "mutual information needs at least 2 qubits");
  const int npairs = n * (n - 1) / 2;
  int rc;
  if (count <= 0) return count == 0 ? QSB_OK : fail(ctx, QSB_E_INVAL, "negative count");
  if ((rc = need(ctx, mi, count * npairs * 8, "mutual information"))) return rc;
  if (entropy1 && (rc = need(ctx, entropy1, count * n * 8, "entropy1"))) return rc;
  qsb_buffer *r1 = nullptr, *r2 = nullptr;
  if ((rc = qsb_buffer_alloc(ctx, count * n * 4 * 16, &r1))) return rc;
  if ((rc = qsb_buffer_alloc(ctx, count * npairs * 16 * 16, &r2))) { qsb_buffer_free(r1); return rc; }
  rc = qsb_rdm_all(ctx, n, states, first, count, r1, r2);
  if (rc == QSB_OK) {
    const int64_t total = count * npairs;
    qsb_mi_kernel<<<(unsigned)((total + 127) / 128), 128, 0, ctx->stream>>>((const c128*)r1->ptr, (const c128*)r2->ptr, n,
                                                                           npairs, total,
                                                                           entropy1 ? (double*)entropy1->ptr : nullptr,
                                                                           (double*)mi->ptr);
    rc = after_launch(ctx, "mutual_information");
  }
  qsb_buffer_free(r1);
  qsb_buffer_free(r2);
  return rc;
}

int qsb_rho_accumulate(qsb_ctx* ctx, int32_t n, qsb_buffer* states, int64_t first, int64_t count, double scale,
                       qsb_buffer* rho) {
  CHECK_N(ctx, n);
  if (n > 14) return fail(ctx, QSB_E_UNSUPPORTED, "rho of %d qubits does not fit (4^n complex128)", n);
  const int64_t dim = (int64_t)1 << n;
  int rc;
  if (count <= 0) return count == 0 ? QSB_OK : fail(ctx, QSB_E_INVAL, "negative count");
  if (first < 0) return fail(ctx, QSB_E_INVAL, "negative first");
  if ((rc = need(ctx, states, (first + count) * dim * ctx->amp_bytes, "states"))) return rc;
  if ((rc = need(ctx, rho, dim * dim * 16, "rho"))) return rc;
  CU(ctx, cudaSetDevice(ctx->device));
  const unsigned g = (unsigned)((dim + QSB_RHO_TILE - 1) / QSB_RHO_TILE);
  QSB_BY_AMP(ctx, (qsb_rho_kernel<A><<<g * (g + 1) / 2, 256, 0, ctx->stream>>>((const A*)states->ptr + first * dim, dim, count,
                                                                              scale, (c128*)rho->ptr, (int)g)));
  return after_launch(ctx, "rho_accumulate");
}

int qsb_readout_transform(qsb_ctx* ctx, int32_t n, qsb_buffer* probs, int64_t count, double p01, double p10) {
  CHECK_N(ctx, n);
  const int64_t dim = (int64_t)1 << n;
  int rc;
  if (count <= 0) return count == 0 ? QSB_OK : fail(ctx, QSB_E_INVAL, "negative count");
  if (!(p01 >= 0 && p01 <= 1 && p10 >= 0 && p10 <= 1))
    return fail(ctx, QSB_E_INVAL, "Readout error probabilities must be in [0, 1]");
  if ((rc = need(ctx, probs, count * dim * 8, "probabilities"))) return rc;
  CU(ctx, cudaSetDevice(ctx->device));
  if (n <= 14 && !getenv("QSB_READOUT_AXIS")) {        // the whole transform in shared memory, one launch
    static bool attr[64] = {false};
    if (!attr[ctx->device & 63]) {
      CU(ctx, cudaFuncSetAttribute(qsb_readout_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 << 14));
      attr[ctx->device & 63] = true;
    }
    qsb_readout_fused_kernel<<<(unsigned)count, 512, (size_t)8 << n, ctx->stream>>>((double*)probs->ptr, n, 1.0 - p01, p10,
                                                                                    p01, 1.0 - p10);
    return after_launch(ctx, "readout_transform");
  }
  // axis q of the reference's [2]*n tensor is bit n-1-q; axes are transformed in order q = 0..n-1 (noise.py:163)
  for (int q = 0; q < n; ++q) {
    qsb_readout_axis_kernel<<<grid_for(ctx, count * dim / 2, 256), 256, 0, ctx->stream>>>(
        (double*)probs->ptr, dim, count, n - 1 - q, 1.0 - p01, p10, p01, 1.0 - p10);
    ctx->launches += 1;
  }
  qsb_normalize_dist_kernel<<<(unsigned)count, 256, 0, ctx->stream>>>((double*)probs->ptr, dim);
  return after_launch(ctx, "readout_transform");
}

int qsb_apply_dense(qsb_ctx* ctx, int32_t n, qsb_buffer* in, int64_t first, int64_t count, qsb_buffer* out, int64_t out_first,
                    int32_t k, const int32_t* target_bits, const double* matrix, const int32_t* out_perm) {
  CHECK_N(ctx, n);
  const int64_t dim = (int64_t)1 << n;
  int rc;
  if (count <= 0) return count == 0 ? QSB_OK : fail(ctx, QSB_E_INVAL, "negative count");
  if (k < 1 || k > n || k > 8) return fail(ctx, QSB_E_INVAL, "qsb_apply_dense: k = %d outside [1, min(n, 8)]", k);
  if (!target_bits || !matrix) return fail(ctx, QSB_E_INVAL, "qsb_apply_dense: NULL argument");
  if (first < 0 || out_first < 0) return fail(ctx, QSB_E_INVAL, "negative offset");
  if (in == out) return fail(ctx, QSB_E_INVAL, "qsb_apply_dense works out of place");
  uint32_t seen = 0;
  for (int j = 0; j < k; ++j) {
    if (target_bits[j] < 0 || target_bits[j] >= n || ((seen >> target_bits[j]) & 1u))
      return fail(ctx, QSB_E_INVAL, "qsb_apply_dense: bad target bit list");
    seen |= 1u << target_bits[j];
  }
  int perm[32];
  for (int b = 0; b < n; ++b) perm[b] = out_perm ? out_perm[b] : b;
  if (!is_perm(perm, n)) return fail(ctx, QSB_E_INVAL, "qsb_apply_dense: out_perm is not a permutation");
  if ((rc = need(ctx, in, (first + count) * dim * ctx->amp_bytes, "input states"))) return rc;
  if ((rc = need(ctx, out, (out_first + count) * dim * ctx->amp_bytes, "output states"))) return rc;
  CU(ctx, cudaSetDevice(ctx->device));
  const size_t mbytes = (size_t)16 << (2 * k);
  void* scratch = nullptr;
  CU(ctx, cudaMallocFromPoolAsync(&scratch, mbytes + 64 * sizeof(int), ctx->pool, ctx->stream));
  int meta[64];
  for (int j = 0; j < 16; ++j) meta[j] = j < k ? target_bits[j] : 0;
  for (int b = 0; b < 32; ++b) meta[16 + b] = b < n ? perm[b] : 0;
  cudaError_t e = cudaMemcpyAsync(scratch, matrix, mbytes, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync((char*)scratch + mbytes, meta, sizeof(int) * 48, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);        // `matrix` and `meta` are borrowed host memory
  if (e != cudaSuccess) { cudaFreeAsync(scratch, ctx->stream); return fail(ctx, QSB_E_CUDA, "qsb_apply_dense: %s", cudaGetErrorString(e)); }
  const c128* dM = (const c128*)scratch;
  const int* dtb = (const int*)((char*)scratch + mbytes);
  const int* dperm = dtb + 16;
  const unsigned grid = grid_for(ctx, count * dim, 256);
  QSB_BY_AMP(ctx, (qsb_dense_kernel<A><<<grid, 256, 0, ctx->stream>>>(
      (const A*)in->ptr + first * dim, (A*)out->ptr + out_first * dim, n, k, count, dM, dtb, dperm)));
  ctx->launches += 1;
  rc = after_launch(ctx, "apply_dense");
  cudaFreeAsync(scratch, ctx->stream);
  return rc;
}

}  // extern "C"

// ---- streamed passes (TMA tile pipeline, qsb_stream.cuh) ---------------------------------------------------------
struct qsb_stream {
  qsb_ctx* ctx;
  qsb_stream_kargs ka;
  int32_t ebit[3];              // positions of the box's extra dimensions, load side
  int32_t ebit_out[3];          // ... and store side
  bool in_place;
  qsb_blk* d_sweeps;            // block list with group orders for 256 workers per group ...
  qsb_blk* d_sweeps128;         // ... and for 128
  c128* d_mats;                 // dense 2- / 3-qubit matrices of the pass
};

typedef CUresult (*qsb_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// cuTensorMapEncodeTiled through the runtime's driver entry point (libcuda is not linked)
static qsb_encode_fn tensor_map_encoder() {
  static qsb_encode_fn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (qsb_encode_fn)p;
    cudaGetLastError();
  }
  return fn;
}

// rank-5 map of a 2^n-amplitude complex128 shard: [16 doubles = 128 B][2^(n-3) chunks][2][2][2]; the box holds one row
// of 2^l amplitudes times the e extra bit dimensions, swizzled like the executor's tile (SWIZZLE_128B)
static int stream_map(qsb_ctx* ctx, CUtensorMap* map, void* base, const qsb_stream* s, const int32_t* ebit) {
  qsb_encode_fn enc = tensor_map_encoder();
  if (!enc) return fail(ctx, QSB_E_UNSUPPORTED, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t gdim[5] = {16, (cuuint64_t)1 << (s->ka.n - 3), 1, 1, 1};
  cuuint64_t gstride[4] = {128, 128, 128, 128};
  cuuint32_t box[5] = {16, (cuuint32_t)1 << (s->ka.l - 3), 1, 1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  for (int j = 0; j < s->ka.e; ++j) {
    gdim[2 + j] = 2;
    gstride[1 + j] = (cuuint64_t)16 << ebit[j];
    box[2 + j] = 2;
  }
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 5, base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(ctx, QSB_E_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return QSB_OK;
}

extern "C" {

int qsb_stream_create(qsb_ctx* ctx, int32_t n, int32_t m, int32_t l, int32_t e, const int32_t* positions,
                      const int32_t* positions_out, const qsb_stream_block* blocks, int32_t n_blocks, const double* cdata,
                      int64_t n_cdata, qsb_stream** out) {
  if (!ctx || !out || !positions) return fail(ctx, QSB_E_INVAL, "qsb_stream_create: NULL argument");
  *out = nullptr;
  if (ctx->amp_bytes != 16) return fail(ctx, QSB_E_UNSUPPORTED, "streamed passes are complex128 only");
  if (n < 6 || n > 30) return fail(ctx, QSB_E_INVAL, "streamed pass: n = %d outside [6, 30]", n);
  if (m < 6 || m > QSB_ST_MAX_TILE_BITS || m > n || l < 3 || e < 0 || e > 3 || l + e > m || m - l - e > 5)
    return fail(ctx, QSB_E_INVAL, "streamed pass: bad tile geometry n=%d m=%d l=%d e=%d", n, m, l, e);
  if (n_blocks < 0 || n_blocks > QSB_ST_MAX_SWEEPS || (n_blocks > 0 && !blocks))
    return fail(ctx, QSB_E_INVAL, "streamed pass: %d block sweeps (at most %d per pass)", n_blocks, QSB_ST_MAX_SWEEPS);
  if (!is_perm(positions, n)) return fail(ctx, QSB_E_INVAL, "streamed pass: positions is not a permutation of 0..n-1");
  if (!positions_out) positions_out = positions;
  if (!is_perm(positions_out, n)) return fail(ctx, QSB_E_INVAL, "streamed pass: positions_out is not a permutation of 0..n-1");
  for (int j = 0; j < l; ++j)
    if (positions[j] != j || positions_out[j] != j)
      return fail(ctx, QSB_E_INVAL, "streamed pass: the low %d slots must be the low index bits", l);
  // dense matrices of the pass, packed behind one another (device copy below)
  std::vector<c128> mats;
  std::vector<qsb_blk> descs((size_t)(n_blocks > 0 ? n_blocks : 1));
  std::vector<int64_t> mat_at((size_t)(n_blocks > 0 ? n_blocks : 1), -1);
  memset((void*)descs.data(), 0, descs.size() * sizeof(qsb_blk));
  for (int i = 0; i < n_blocks; ++i) {
    const qsb_stream_block& bk = blocks[i];
    qsb_blk& d = descs[i];
    if (bk.n_ops < 0 || bk.n_ops > QSB_ST_BLOCK_OPS) return fail(ctx, QSB_E_INVAL, "block %d: %d ops", i, bk.n_ops);
    uint32_t used = 0;
    for (int j = 0; j < 4; ++j) {
      if (bk.b[j] < 0 || bk.b[j] >= m || ((used >> bk.b[j]) & 1u)) return fail(ctx, QSB_E_INVAL, "block %d: bad slot bit", i);
      used |= 1u << bk.b[j];
      d.b[j] = bk.b[j];
    }
    // The op list must be CANONICAL: a 2x2 on a local bit comes before every gate on that bit (at most one per bit), a
    // dense gate comes after the 2x2s and before everything else.  Then the 2x2s run in bit order and every
    // permutation-type gate (CX, SWAP, Toffoli, Fredkin) and sign gate (CZ) is composed here into "register r is stored
    // to the place of x[r], negated when sign[r] < 0".
    int x[16], sign[16];
    for (int r = 0; r < 16; ++r) { x[r] = r; sign[r] = 1; }
    uint32_t gate_bits = 0, mat_bits = 0;
    bool after_dense = false;
    for (int t = 0; t < 4; ++t) {
      d.cls[t] = QSB_CLS_NONE;
      d.U[t][0] = d.U[t][3] = qsb_c(1.0, 0.0);
      d.U[t][1] = d.U[t][2] = qsb_c(0.0, 0.0);
    }
    for (int q = 0; q < bk.n_ops; ++q) {
      const qsb_stream_op& o = bk.ops[q];
      const int t0 = o.t[0], t1 = o.t[1], t2 = o.t[2];
      const bool in0 = t0 >= 0 && t0 < 4, in1 = t1 >= 0 && t1 < 4, in2 = t2 >= 0 && t2 < 4;
      bool ok = false;
      switch (o.kind) {
        case QSB_B_MAT1:
          ok = in0 && o.cls >= QSB_CLS_RDIAG && o.cls <= QSB_CLS_DENSE && !((gate_bits | mat_bits) >> t0 & 1u) && !after_dense;
          if (ok) {
            mat_bits |= 1u << t0;
            d.cls[t0] = o.cls;
            for (int z = 0; z < 4; ++z) d.U[t0][z] = qsb_c(o.U[2 * z], o.U[2 * z + 1]);
          }
          break;
        case QSB_B_CZ:
          ok = in0 && in1 && t0 != t1;
          if (ok) for (int r = 0; r < 16; ++r) if (((x[r] >> t0) & 1) && ((x[r] >> t1) & 1)) sign[r] = -sign[r];
          if (ok) gate_bits |= (1u << t0) | (1u << t1);
          break;
        case QSB_B_CX:
          ok = in0 && in1 && t0 != t1;
          if (ok) for (int r = 0; r < 16; ++r) x[r] ^= ((x[r] >> t0) & 1) << t1;
          if (ok) gate_bits |= (1u << t0) | (1u << t1);
          break;
        case QSB_B_SWAP:
          ok = in0 && in1 && t0 != t1;
          if (ok) for (int r = 0; r < 16; ++r) {
            const int ba = (x[r] >> t0) & 1, bb = (x[r] >> t1) & 1;
            x[r] = (x[r] & ~((1 << t0) | (1 << t1))) | (bb << t0) | (ba << t1);
          }
          if (ok) gate_bits |= (1u << t0) | (1u << t1);
          break;
        case QSB_B_CCX:
          ok = in0 && in1 && in2 && t0 < t1 && t2 != t0 && t2 != t1;
          if (ok) for (int r = 0; r < 16; ++r) x[r] ^= (((x[r] >> t0) & (x[r] >> t1)) & 1) << t2;
          if (ok) gate_bits |= (1u << t0) | (1u << t1) | (1u << t2);
          break;
        case QSB_B_CSWAP:
          ok = in0 && in1 && in2 && t1 < t2 && t0 != t1 && t0 != t2;
          if (ok) for (int r = 0; r < 16; ++r) if ((x[r] >> t0) & 1) {
            const int ba = (x[r] >> t1) & 1, bb = (x[r] >> t2) & 1;
            x[r] = (x[r] & ~((1 << t1) | (1 << t2))) | (bb << t1) | (ba << t2);
          }
          if (ok) gate_bits |= (1u << t0) | (1u << t1) | (1u << t2);
          break;
        case QSB_B_DENSE2: case QSB_B_DENSE3: {
          const int cnt = o.kind == QSB_B_DENSE2 ? 16 : 64;
          ok = !after_dense && gate_bits == 0;
          if (ok && (!cdata || bk.mat < 0 || (bk.mat & 1) || bk.mat + 2 * cnt > n_cdata))
            return fail(ctx, QSB_E_INVAL, "block %d: dense matrix outside cdata", i);
          if (ok) {
            after_dense = true;
            d.dense = o.kind == QSB_B_DENSE2 ? 2 : 3;
            gate_bits |= o.kind == QSB_B_DENSE2 ? 0xcu : 0xeu;
            mat_at[i] = (int64_t)mats.size();
            for (int z = 0; z < cnt; ++z) mats.push_back(qsb_c(cdata[bk.mat + 2 * z], cdata[bk.mat + 2 * z + 1]));
          }
          break;
        }
        default: break;
      }
      if (!ok)
        return fail(ctx, QSB_E_INVAL, "block %d op %d: kind %d with bits (%d, %d, %d) is invalid or not in canonical order", i, q,
                    o.kind, t0, t1, t2);
    }
    auto tile_off = [&](int v) {
      int idx = 0;
      for (int j = 0; j < 4; ++j) idx |= ((v >> j) & 1) << bk.b[j];
      return (uint32_t)qsb_slot(idx) << 4;
    };
    d.neg_mask = 0;
    for (int r = 0; r < 16; ++r) {
      d.ld_off[r] = tile_off(r);
      d.st_off[r] = tile_off(x[r]);
      if (sign[r] < 0) d.neg_mask |= 1u << r;
    }
  }
  CU(ctx, cudaSetDevice(ctx->device));
  qsb_stream* s = new qsb_stream();
  memset(&s->ka, 0, sizeof s->ka);
  s->ctx = ctx;
  s->ka.n = n; s->ka.m = m; s->ka.l = l; s->ka.e = e;
  s->ka.n_sweeps = n_blocks;
  s->ka.n_ops = 1 << (m - l - e);
  s->in_place = true;
  for (int j = 0; j < n; ++j) s->in_place = s->in_place && positions[j] == positions_out[j];
  for (int j = 0; j < 3; ++j) { s->ebit[j] = j < e ? positions[l + j] : 0; s->ebit_out[j] = j < e ? positions_out[l + j] : 0; }
  for (int j = 0; j < m - l - e; ++j) { s->ka.op_pos[j] = positions[l + e + j]; s->ka.op_pos_out[j] = positions_out[l + e + j]; }
  for (int j = 0; j < n - m; ++j) { s->ka.tile_pos[j] = positions[m + j]; s->ka.tile_pos_out[j] = positions_out[m + j]; }
  s->d_sweeps = s->d_sweeps128 = nullptr;
  s->d_mats = nullptr;
  // device copies: dense matrices, then the block list twice (group orders for 256 and for 128 workers per group)
  const size_t nb = descs.size();
  cudaError_t er = cudaMalloc(&s->d_sweeps, 2 * nb * sizeof(qsb_blk));
  if (er == cudaSuccess && !mats.empty()) er = cudaMalloc(&s->d_mats, mats.size() * sizeof(c128));
  if (er == cudaSuccess && !mats.empty())
    er = cudaMemcpyAsync(s->d_mats, mats.data(), mats.size() * sizeof(c128), cudaMemcpyHostToDevice, ctx->stream);
  for (int pass = 0; pass < 2 && er == cudaSuccess; ++pass) {
    for (int i = 0; i < n_blocks; ++i) {
      uint32_t used = 0;
      for (int j = 0; j < 4; ++j) used |= 1u << descs[i].b[j];
      int hm = 0;
      const int wb = pass == 0 ? 8 : 7;
      descs[i].pos = qsb_group_order(m, used, wb, m - 4, 3, &hm);
      descs[i].hmask = hm;
      const int nbits = wb < m - 4 ? wb : m - 4;
      for (int e = 0; e < 32; ++e) descs[i].tabl[e] = qsb_deposit(e, descs[i].pos, nbits < 5 ? nbits : 5);
      for (int e = 0; e < 8; ++e) descs[i].tabw[e] = qsb_deposit(e << 5, descs[i].pos, nbits);
      descs[i].variant = (descs[i].cls[0] != QSB_CLS_NONE ? 1 : 0) | (descs[i].cls[1] != QSB_CLS_NONE ? 2 : 0) |
                         (descs[i].cls[2] != QSB_CLS_NONE ? 4 : 0) | (descs[i].cls[3] != QSB_CLS_NONE ? 8 : 0);
      descs[i].mat = mat_at[i] >= 0 ? s->d_mats + mat_at[i] : nullptr;
    }
    er = cudaMemcpyAsync(s->d_sweeps + pass * nb, descs.data(), nb * sizeof(qsb_blk), cudaMemcpyHostToDevice, ctx->stream);
    if (er == cudaSuccess) er = cudaStreamSynchronize(ctx->stream);        // `descs` is re-used for the second copy
  }
  s->d_sweeps128 = s->d_sweeps + nb;
  if (er != cudaSuccess) {
    cudaFree(s->d_sweeps);
    cudaFree(s->d_mats);
    delete s;
    cudaGetLastError();
    return fail(ctx, er == cudaErrorMemoryAllocation ? QSB_E_OOM : QSB_E_CUDA, "streamed pass upload: %s", cudaGetErrorString(er));
  }
  s->ka.sweeps = s->d_sweeps;
  *out = s;
  return QSB_OK;
}

// tile-number bits that select the peer (positions >= shift on the given side), each set to the matching bit of `rank`
static uint32_t peer_tile_xor(const qsb_stream* s, const int32_t* tile_pos, int shift, uint32_t rank) {
  uint32_t x = 0;
  for (int q = 0; q < s->ka.n - s->ka.m; ++q)
    if (tile_pos[q] >= shift && ((rank >> (tile_pos[q] - shift)) & 1u)) x |= 1u << q;
  return x;
}

static int stream_launch(qsb_stream* s, qsb_stream_maps& maps, int32_t flags) {
  qsb_ctx* ctx = s->ctx;
  const size_t smem = qsb_stream_smem_bytes(s->ka.m);
  static bool attr_set[64] = {false};
  if (!attr_set[ctx->device & 63]) {          // the opt-in to > 48 KiB of dynamic shared memory is per device
    CU(ctx, cudaFuncSetAttribute(qsb_stream_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)qsb_stream_smem_bytes(QSB_ST_MAX_TILE_BITS)));
    attr_set[ctx->device & 63] = true;
  }
  const int64_t ntiles = (int64_t)1 << (s->ka.n - s->ka.m);
  const int grid = (int)(ntiles < ctx->sm_count ? ntiles : ctx->sm_count);
  qsb_stream_kargs ka = s->ka;
  ka.sweeps = s->d_sweeps128;
  qsb_stream_kernel<128><<<grid, QSB_ST_GROUPS * 128 + 32, smem, ctx->stream>>>(maps, ka);
  cudaError_t er = cudaGetLastError();
  if (er != cudaSuccess) return fail(ctx, QSB_E_CUDA, "streamed pass launch: %s", cudaGetErrorString(er));
  ctx->launches += 1;
  if (!(flags & QSB_RUN_ASYNC)) CU(ctx, cudaStreamSynchronize(ctx->stream));
  return QSB_OK;
}

int qsb_stream_run(qsb_stream* s, qsb_buffer* in, int64_t in_offset, qsb_buffer* out, int64_t out_offset, int32_t flags) {
  if (!s || !in) return fail(nullptr, QSB_E_INVAL, "qsb_stream_run: NULL argument");
  qsb_ctx* ctx = s->ctx;
  if (!out) { out = in; out_offset = in_offset; }
  const int64_t dim = (int64_t)1 << s->ka.n;
  int rc;
  if (!s->in_place && (char*)in->ptr + in_offset * 16 == (char*)out->ptr + out_offset * 16)
    return fail(ctx, QSB_E_INVAL, "qsb_stream_run: a pass that stores to other positions than it loads from cannot run in place");
  if (in_offset < 0 || out_offset < 0 || (in_offset & 7) || (out_offset & 7))
    return fail(ctx, QSB_E_INVAL, "qsb_stream_run: offsets must be non-negative multiples of 8 amplitudes");
  if ((rc = need(ctx, in, (in_offset + dim) * 16, "input shard"))) return rc;
  if ((rc = need(ctx, out, (out_offset + dim) * 16, "output shard"))) return rc;
  CU(ctx, cudaSetDevice(ctx->device));
  qsb_stream_maps maps;
  memset(&maps, 0, sizeof maps);
  if ((rc = stream_map(ctx, &maps.in[0], (char*)in->ptr + in_offset * 16, s, s->ebit))) return rc;
  if ((rc = stream_map(ctx, &maps.out[0], (char*)out->ptr + out_offset * 16, s, s->ebit_out))) return rc;
  s->ka.peer_shift = 32;
  s->ka.peer_or = 0;
  s->ka.out_shift = 32;
  s->ka.out_or = 0;
  s->ka.tile_xor = 0;
  return stream_launch(s, maps, flags);
}

int qsb_stream_run_scatter(qsb_stream* s, qsb_buffer* in, int64_t in_offset, const void* const* peers, int32_t n_peers,
                           int32_t peer_shift, int64_t peer_rank_or, int32_t flags) {
  if (!s || !in || !peers) return fail(nullptr, QSB_E_INVAL, "qsb_stream_run_scatter: NULL argument");
  qsb_ctx* ctx = s->ctx;
  const int n = s->ka.n;
  const int64_t dim = (int64_t)1 << n;
  int rc;
  if (n_peers < 2 || n_peers > QSB_ST_MAX_PEERS || peer_shift < s->ka.l || peer_shift >= n || (1 << (n - peer_shift)) != n_peers)
    return fail(ctx, QSB_E_INVAL, "qsb_stream_run_scatter: %d peers do not match peer_shift %d of %d bits", n_peers, peer_shift, n);
  if (peer_rank_or < 0 || peer_rank_or >= dim || (peer_rank_or & (((int64_t)1 << peer_shift) - 1)))
    return fail(ctx, QSB_E_INVAL, "qsb_stream_run_scatter: bad peer_rank_or");
  for (int j = 0; j < s->ka.e; ++j)
    if (s->ebit_out[j] >= peer_shift)
      return fail(ctx, QSB_E_UNSUPPORTED, "qsb_stream_run_scatter: a TMA box dimension of the store lies on a peer-selecting bit");
  if (in_offset < 0 || (in_offset & 7)) return fail(ctx, QSB_E_INVAL, "qsb_stream_run_scatter: bad in_offset");
  if ((rc = need(ctx, in, (in_offset + dim) * 16, "input shard"))) return rc;
  CU(ctx, cudaSetDevice(ctx->device));
  qsb_stream_maps maps;
  memset(&maps, 0, sizeof maps);
  if ((rc = stream_map(ctx, &maps.in[0], (char*)in->ptr + in_offset * 16, s, s->ebit))) return rc;
  for (int p = 0; p < n_peers; ++p) {
    if (!peers[p]) return fail(ctx, QSB_E_INVAL, "qsb_stream_run_scatter: peer %d is NULL", p);
    if (peers[p] == (const void*)((char*)in->ptr + in_offset * 16))
      return fail(ctx, QSB_E_INVAL, "qsb_stream_run_scatter: the pass would overwrite the shard it reads");
    if ((rc = stream_map(ctx, &maps.out[p], const_cast<void*>(peers[p]), s, s->ebit_out))) return rc;
  }
  s->ka.peer_shift = 32;
  s->ka.peer_or = 0;
  s->ka.out_shift = peer_shift;
  s->ka.out_or = (uint32_t)peer_rank_or;
  s->ka.tile_xor = peer_tile_xor(s, s->ka.tile_pos_out, peer_shift, (uint32_t)(peer_rank_or >> peer_shift));
  return stream_launch(s, maps, flags);
}

int qsb_stream_run_peers(qsb_stream* s, const void* const* peers, int32_t n_peers, int32_t peer_shift, int64_t peer_rank_or,
                         qsb_buffer* out, int64_t out_offset, int32_t flags) {
  if (!s || !peers || !out) return fail(nullptr, QSB_E_INVAL, "qsb_stream_run_peers: NULL argument");
  qsb_ctx* ctx = s->ctx;
  const int n = s->ka.n;
  const int64_t dim = (int64_t)1 << n;
  int rc;
  if (n_peers < 2 || n_peers > QSB_ST_MAX_PEERS || peer_shift < s->ka.l || peer_shift >= n || (1 << (n - peer_shift)) != n_peers)
    return fail(ctx, QSB_E_INVAL, "qsb_stream_run_peers: %d peers do not match peer_shift %d of %d bits", n_peers, peer_shift, n);
  if (peer_rank_or < 0 || peer_rank_or >= dim || (peer_rank_or & (((int64_t)1 << s->ka.l) - 1)))
    return fail(ctx, QSB_E_INVAL, "qsb_stream_run_peers: bad peer_rank_or");
  for (int j = 0; j < s->ka.e; ++j)
    if (s->ebit[j] >= peer_shift)
      return fail(ctx, QSB_E_UNSUPPORTED, "qsb_stream_run_peers: a TMA box dimension lies on a peer-selecting bit");
  if (out_offset < 0 || (out_offset & 7)) return fail(ctx, QSB_E_INVAL, "qsb_stream_run_peers: bad out_offset");
  if ((rc = need(ctx, out, (out_offset + dim) * 16, "output shard"))) return rc;
  CU(ctx, cudaSetDevice(ctx->device));
  qsb_stream_maps maps;
  memset(&maps, 0, sizeof maps);
  for (int p = 0; p < n_peers; ++p) {
    if (!peers[p]) return fail(ctx, QSB_E_INVAL, "qsb_stream_run_peers: peer %d is NULL", p);
    if ((rc = stream_map(ctx, &maps.in[p], const_cast<void*>(peers[p]), s, s->ebit))) return rc;
  }
  if ((rc = stream_map(ctx, &maps.out[0], (char*)out->ptr + out_offset * 16, s, s->ebit_out))) return rc;
  s->ka.peer_shift = peer_shift;
  s->ka.peer_or = (uint32_t)peer_rank_or;
  s->ka.out_shift = 32;
  s->ka.out_or = 0;
  s->ka.tile_xor = peer_tile_xor(s, s->ka.tile_pos, peer_shift, (uint32_t)(peer_rank_or >> peer_shift));
  return stream_launch(s, maps, flags);
}

int qsb_stream_free(qsb_stream* s) {
  if (!s) return QSB_OK;
  cudaSetDevice(s->ctx->device);
  cudaStreamSynchronize(s->ctx->stream);
  cudaFree(s->d_sweeps);
  cudaFree(s->d_mats);
  delete s;
  return QSB_OK;
}

}  // extern "C"
