/* qsb.h -- C ABI of libqsb.so, the B200 (sm_100a) statevector backend that sits
 * behind the Python entry points of justinbrianhwang/Quantum-Simulator's NumPy
 * engine (quantum_sim/engine).
 *
 * The reference has no FFI layer at all (pure Python + NumPy); the boundary is
 * its Python API.  Each entry point below names the reference code whose array
 * work it replaces (file:line relative to the reference tree).  The Python side
 * (quantum-simulator_b200/qsb/capi.py) binds these with ctypes; INTEGRATION.md
 * shows the stub a maintainer adds to the reference.
 *
 * Conventions
 *   - every function returns int: 0 = QSB_OK, negative = error; text through
 *     qsb_last_error().  There is NO CPU fallback: without a CUDA device
 *     qsb_ctx_create fails with QSB_E_NODEV.
 *   - host pointers are caller-owned, borrowed for the duration of the call.
 *   - device memory is owned by opaque handles (or wrapped, never freed, when
 *     created with *_wrap from a caller's device pointer, e.g. a torch tensor).
 *   - a ctx is single-threaded by contract; calls are synchronous on return
 *     unless QSB_RUN_ASYNC is given (then order on the ctx stream; qsb_ctx_sync).
 *   - state layout: complex128[batch][2^n] (or complex64 in c64 mode, later),
 *     qubit 0 = most significant bit of the amplitude index (state_vector.py:87-88).
 */
#ifndef QSB_H
#define QSB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QSB_VERSION 100

enum {
  QSB_OK = 0,
  QSB_E_INVAL = -1,       /* bad argument (Python raises ValueError)            */
  QSB_E_NODEV = -2,       /* no CUDA device / device index out of range         */
  QSB_E_CUDA = -3,        /* CUDA runtime error                                 */
  QSB_E_OOM = -4,         /* device allocation failed                           */
  QSB_E_UNSUPPORTED = -5  /* valid request this build cannot run (e.g. n > 16)  */
};

/* amplitude type of the state buffers a context works on (qsb_ctx_set_precision) */
enum {
  QSB_C128 = 0,           /* complex128 states: the reference's precision, tolerance 1e-12 (default)       */
  QSB_C64 = 1             /* complex64 states (BASELINE's separately reported 1e-5 mode): tiles hold twice the
                             amplitudes (local_bits <= 14), arithmetic in fp32, reductions accumulate in fp64;
                             every *state* buffer (states, snapshots, states_out) is complex64, every result
                             buffer (probabilities, overlaps, RDMs, rho) keeps its double / complex128 type     */
};

typedef struct qsb_ctx qsb_ctx;
typedef struct qsb_buffer qsb_buffer;     /* raw device bytes                      */
typedef struct qsb_program qsb_program;   /* lowered circuit (+noise) on device    */

/* ---- lowered operation ------------------------------------------------------
 * One entry of a program.  Bit fields are SLOT bit positions of the amplitude
 * index inside the resident tile (0 = least significant); the host compiler
 * (qsb/compiler.py) maps reference qubits -> physical bits -> slots, tracks the
 * reference's axis scramble (state_vector.py:66-73) and inserts QSB_OP_REMAP
 * when a target lives in a cluster-rank bit.
 *   data  : offset (in doubles) into the program's cdata
 *   param : offset into the per-state parameter row (QSB_OP_R*), else -1
 *   draw  : index of the uniform this Kraus op consumes, else -1
 *   aux   : op specific (snapshot: offset of its bit permutation in idata)
 */
typedef struct qsb_op {
  int32_t kind;
  int32_t b0, b1, b2;
  int32_t data;
  int32_t param;
  int32_t draw;
  int32_t aux;
} qsb_op;

enum {
  QSB_OP_NOP = 0,
  /* dense unitaries: cdata[data ..] = row-major (re,im) matrix; b0 = MSB of the
   * matrix index = target_qubits[0] (state_vector.py:57-63)                      */
  QSB_OP_U1 = 1,      /* 2x2,  8 doubles   */
  QSB_OP_U2 = 2,      /* 4x4,  32 doubles  */
  QSB_OP_U3Q = 3,     /* 8x8,  128 doubles */
  QSB_OP_D1 = 4,      /* diagonal 2x2: cdata = d0.re d0.im d1.re d1.im (Z,S,T,Rz,Phase) */
  /* structured gates (gates.py:99-125) */
  QSB_OP_X = 10, QSB_OP_Y = 11, QSB_OP_Z = 12,
  QSB_OP_CX = 13,     /* b0 control, b1 target                    */
  QSB_OP_CZ = 14,
  QSB_OP_SWAP = 15,
  QSB_OP_CCX = 16,    /* b0,b1 controls, b2 target (Toffoli)      */
  QSB_OP_CSWAP = 17,  /* b0 control, swaps b1,b2 (Fredkin)        */
  /* per-state parameterised 1-qubit gates (gates.py:66-94): angle(s) read from
   * the state's parameter row at [param], [param+1], [param+2]               */
  QSB_OP_RX = 20, QSB_OP_RY = 21, QSB_OP_RZ = 22, QSB_OP_PHASE = 23, QSB_OP_U3 = 24,
  /* stochastic Kraus steps (noise.py:224-260), one uniform each              */
  QSB_OP_KRAUS_PAULI = 30,  /* cdata: c0 c1 c2 (cdf of choice()), then 4 Pauli codes 0..3 as doubles */
  QSB_OP_KRAUS_AD = 31,     /* cdata: gamma, sqrt(1-gamma), sqrt(gamma)                   */
  QSB_OP_KRAUS_GEN = 32,    /* cdata: nK, then per K: 8 doubles K, 4 doubles K^dag K (e00,e11,re e01,im e01) */
  /* cluster data movement: swap rank bit b0 (0..log2 C-1) with local slot bit b1; b2 = number of FURTHER
   * (rank bit, local bit) pairs exchanged in the same pass, packed in aux as g1 | l1 << 8 | g2 << 16 | l2 << 24 */
  QSB_OP_REMAP = 40,
  /* copy the (normalised) state to snapshot slot b0 with bit permutation idata[aux..aux+n) */
  QSB_OP_SNAPSHOT = 50
};

/* flags of qsb_run */
enum {
  QSB_RUN_LOAD = 1,        /* start from states[first+t] instead of a basis state        */
  QSB_RUN_STORE = 2,       /* write the final state to states[first+t] in reference order */
  QSB_RUN_NORMALIZE = 4,   /* divide by ||psi|| on store/snapshot (set when Kraus ops ran) */
  QSB_RUN_ASYNC = 8,       /* do not synchronise the ctx stream before returning          */
  QSB_RUN_ACCUM_PROBS = 16,/* atomically add |psi|^2 (reference order) into probs_accum   */
  QSB_RUN_LOAD_BROADCAST = 32 /* with LOAD: every unit starts from states[first] (one codeword, many trials) */
};

typedef struct qsb_run_args {
  qsb_buffer* states;        /* complex128[batch][2^n] or NULL (no LOAD/STORE)             */
  int64_t first, count;      /* states / trajectories [first, first+count)                  */
  qsb_buffer* params;        /* double[>=count][params_stride] or NULL                      */
  int64_t params_stride;
  qsb_buffer* uniforms;      /* double[>=count][uniforms_stride] ("reference draws" mode)   */
  int64_t uniforms_stride;   /*   NULL => counter-based Philox4x32-10 in the kernel         */
  uint64_t philox_seed;
  int64_t traj_offset;       /* global index of trajectory `first` (Philox counter, shards) */
  qsb_buffer* init_basis;    /* int64[>=count] reference-order basis index per state, or NULL */
  int64_t default_basis;     /* used when init_basis is NULL                                */
  qsb_buffer* branches;      /* int32[>=count][branches_stride] chosen Kraus index per draw, or NULL */
  int64_t branches_stride;
  qsb_buffer* snapshots;     /* complex128[>=count][n_snapshots][2^n] or NULL               */
  qsb_buffer* probs_accum;   /* double[2^n] or NULL                                         */
  int32_t flags;
  int32_t reserved;
  qsb_buffer* states_out;    /* STORE destination (same layout as states); NULL = in place.  Needed when a
                                streamed pass (n - local_bits > 3) stores with a bit permutation other than
                                the one it loaded with: tiles then read and write different addresses      */
  int64_t out_first;
  /* Streamed passes on a state sharded over GPUs (config 5): LOAD straight from the peers' shards, which folds the
   * qubit exchange (rank bits <-> top local bits, the all-to-all of bigstate.py) into the pass that follows it.
   * peer_table = device array of 2^g device pointers (the peers' shard bases, e.g. from CUDA IPC / torch symmetric
   * memory), or NULL.  Source element s of the (post-exchange) local shard is read from
   * peer_table[s >> peer_shift] at offset (s & (2^peer_shift - 1)) | peer_rank_or.  Needs states_out. */
  qsb_buffer* peer_table;
  int32_t peer_shift;
  int32_t reserved2;
  int64_t peer_rank_or;
} qsb_run_args;

/* ---- lifecycle --------------------------------------------------------------- */
int qsb_version(void);
int qsb_device_count(void);                       /* <0 on error                              */
int qsb_ctx_create(int device, qsb_ctx** out);
int qsb_ctx_destroy(qsb_ctx* ctx);
int qsb_ctx_set_stream(qsb_ctx* ctx, void* cuda_stream);  /* run on a caller's stream (torch) */
int qsb_ctx_sync(qsb_ctx* ctx);
int qsb_ctx_set_precision(qsb_ctx* ctx, int precision);   /* QSB_C128 | QSB_C64; programs keep the mode they were created in */
const char* qsb_last_error(qsb_ctx* ctx);         /* ctx may be NULL: last error of the thread */
int qsb_ctx_info(qsb_ctx* ctx, int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor,
                 int64_t* total_mem);
/* device timer on the ctx stream (CUDA events) */
int qsb_timer_start(qsb_ctx* ctx);
int qsb_timer_stop(qsb_ctx* ctx, float* ms_out);  /* synchronises */
/* number of kernels this ctx has launched so far (bench.py's gpu_launches) */
int64_t qsb_launch_count(qsb_ctx* ctx);

/* ---- device buffers / host staging ------------------------------------------- */
/* Buffers come from a per-context stream-ordered pool that keeps freed blocks cached, so the per-call batch
 * allocations of Simulator.run_with_noise (simulator.py:140-193 builds a fresh StateVector per shot) cost
 * microseconds; qsb_ctx_trim hands the cached blocks back to the driver. */
int qsb_buffer_alloc(qsb_ctx* ctx, int64_t bytes, qsb_buffer** out);
int qsb_ctx_trim(qsb_ctx* ctx);
int qsb_buffer_wrap(qsb_ctx* ctx, void* device_ptr, int64_t bytes, qsb_buffer** out);
int qsb_buffer_free(qsb_buffer* buf);
int qsb_buffer_upload(qsb_buffer* buf, int64_t offset, const void* host, int64_t bytes);
/* no host wait: `host` must stay untouched until an event recorded after the call has been waited for
 * (truly asynchronous only from qsb_host_alloc memory) */
int qsb_buffer_upload_async(qsb_buffer* buf, int64_t offset, const void* host, int64_t bytes);
int qsb_buffer_download(qsb_buffer* buf, int64_t offset, void* host, int64_t bytes);
int qsb_buffer_zero(qsb_buffer* buf, int64_t offset, int64_t bytes);
int qsb_buffer_copy(qsb_buffer* dst, int64_t dst_off, qsb_buffer* src, int64_t src_off, int64_t bytes);
void* qsb_buffer_ptr(qsb_buffer* buf);
int64_t qsb_buffer_bytes(qsb_buffer* buf);
int qsb_host_alloc(int64_t bytes, void** out);    /* pinned host memory for NumPy views */
int qsb_host_free(void* p);
/* markers on the ctx stream, for pipelining host-side draw generation against running trajectories */
typedef struct qsb_event qsb_event;
int qsb_event_create(qsb_ctx* ctx, qsb_event** out);
int qsb_event_record(qsb_event* ev);
int qsb_event_wait(qsb_event* ev);                 /* blocks the host until the recorded point has executed */
int qsb_event_free(qsb_event* ev);

/* ---- programs ------------------------------------------------------------------
 * Replaces the per-gate loop of Simulator.run (simulator.py:57-71), NoiseModel.apply
 * (noise.py:212-260) and StateVector.apply_gate (state_vector.py:41-74).
 *   local_bits = index bits resident per CTA (<= 13).
 *   RESIDENT mode (n_qubits <= 16 and n_qubits - local_bits <= 3): a cluster of 2^(n-local_bits)
 *     CTAs holds the whole state in shared memory for the whole program.
 *   STREAMING mode (otherwise, n_qubits <= 30): the state lives in HBM; one launch = one pass that
 *     loads every 2^local_bits-amplitude tile, applies the program (all op bits < local_bits) and stores
 *     it.  No state-dependent Kraus draws, snapshots or normalisation in this mode.
 *   load_perm/store_perm = idata offsets of n-entry bit permutations (slot bit j -> bit of the
 *   amplitude index in memory).                                                        */
int qsb_program_create(qsb_ctx* ctx, int32_t n_qubits, int32_t local_bits,
                       const qsb_op* ops, int64_t n_ops, int64_t ops_stride /*0 = shared*/,
                       int64_t n_programs /*1 if shared*/,
                       const double* cdata, int64_t n_cdata,
                       const int32_t* idata, int64_t n_idata,
                       int32_t load_perm, int32_t store_perm, int32_t n_snapshots,
                       qsb_program** out);
int qsb_program_free(qsb_program* prog);
int qsb_run(qsb_program* prog, const qsb_run_args* args);
/* developer aid: per-CTA cycle counters of the executor kernel (worker wait / busy per descriptor kind,
 * control-warp stalls); enable != 0 switches them on for later qsb_run calls; out (may be NULL) receives
 * unsigned long long[ctas][128] of the last run; returns the number of CTAs copied or a negative error */
int qsb_debug_profile(qsb_ctx* ctx, int enable, unsigned long long* out, int64_t max_ctas);

/* ---- reductions over stored (reference-order) states ---------------------------- */
/* StateVector.probabilities (state_vector.py:36-39): out double[count][2^n] */
int qsb_probabilities(qsb_ctx* ctx, int32_t n, qsb_buffer* states, int64_t first, int64_t count,
                      qsb_buffer* out, int64_t out_first);
/* sum_t |psi_t|^2 -> double[2^n] (added into out) */
int qsb_probabilities_sum(qsb_ctx* ctx, int32_t n, qsb_buffer* states, int64_t first, int64_t count,
                          qsb_buffer* out);
/* StateVector.measure_all (state_vector.py:107-113): idx = choice(2^n, p) from one uniform each */
int qsb_sample_index(qsb_ctx* ctx, int32_t n, qsb_buffer* states, int64_t first, int64_t count,
                     qsb_buffer* uniforms /*double[count]*/, qsb_buffer* out /*int64[count]*/);
/* <a_t|b_t> (np.vdot, analysis.py:40): out complex128[count]; stride_b = 0 broadcasts one b */
int qsb_overlap(qsb_ctx* ctx, int32_t n, qsb_buffer* a, int64_t a_first, qsb_buffer* b,
                int64_t b_first, int64_t b_stride_states, int64_t count, qsb_buffer* out);
/* masked parity weights (qec.py:466-484, :131-151): out double[count][n_masks][2] = (p_even, p_odd) */
int qsb_masked_parity(qsb_ctx* ctx, int32_t n, qsb_buffer* states, int64_t first, int64_t count,
                      const uint64_t* masks, int32_t n_masks, qsb_buffer* out);
/* all 1- and 2-qubit reduced density matrices (analysis.py:120-166, state_vector.py:121-140):
 * rdm1 complex128[count][n][2][2], rdm2 complex128[count][n(n-1)/2][4][4], pair order (i<j) row-major */
int qsb_rdm_all(qsb_ctx* ctx, int32_t n, qsb_buffer* states, int64_t first, int64_t count,
                qsb_buffer* rdm1, qsb_buffer* rdm2);
/* reduced density matrix of k <= 6 kept qubits, ascending (StateAnalysis.partial_trace, analysis.py:120-166):
 * out complex128[count][2^k][2^k], first kept qubit = most significant bit of the row / column index */
int qsb_rdm_general(qsb_ctx* ctx, int32_t n, qsb_buffer* states, int64_t first, int64_t count,
                    const int32_t* keep_qubits, int32_t k, qsb_buffer* out);
/* all-pairs mutual information I(i:j) = max(0, S_i + S_j - S_ij) in bits (analysis.py:99-104, :183-191, :315-333):
 * reduced density matrices as in qsb_rdm_all, eigenvalues by Jacobi rotations on the device, eigenvalues <= 1e-15
 * dropped.  mi double[count][n(n-1)/2] (pairs i<j row-major); entropy1 double[count][n] or NULL */
int qsb_mi_all_pairs(qsb_ctx* ctx, int32_t n, qsb_buffer* states, int64_t first, int64_t count,
                     qsb_buffer* mi, qsb_buffer* entropy1);
/* ensemble rho (simulator.py:195-198): rho[i][j] += scale * sum_t psi_t[i] conj(psi_t[j]) */
int qsb_rho_accumulate(qsb_ctx* ctx, int32_t n, qsb_buffer* states, int64_t first, int64_t count,
                       double scale, qsb_buffer* rho);
/* ReadoutError.apply_to_distribution (noise.py:141-175) in place on double[count][2^n] */
int qsb_readout_transform(qsb_ctx* ctx, int32_t n, qsb_buffer* probs, int64_t count,
                          double p01, double p10);

/* StateVector.apply_gate for a dense k-qubit operator with k > 3 that is not a Kronecker product
 * (state_vector.py:41-74 accepts any k; every registered gate has k <= 3 and Pauli strings factorise, so this is the
 * rare path): out-of-place on states at rest, target_bits[j] = index bit (n-1-qubit) of targets[j], matrix =
 * complex128[2^k][2^k] on the host (targets[0] = most significant bit of its index), out_perm[b] = destination bit
 * of index bit b (the reference's axis scramble) or NULL for none.  1 <= k <= min(n, 8). */
int qsb_apply_dense(qsb_ctx* ctx, int32_t n, qsb_buffer* in, int64_t first, int64_t count, qsb_buffer* out,
                    int64_t out_first, int32_t k, const int32_t* target_bits, const double* matrix,
                    const int32_t* out_perm);

/* ---- streamed passes over a state in HBM (n > 16; BASELINE config 5) ---------------------------------------------
 * StateVector.apply_gate (state_vector.py:41-74) gate by gate on one 2^n state beyond the constructor's 16-qubit cap
 * (bypassed the way state_vector.py:156-158 does).  A pass = one launch that reads every amplitude once and writes it
 * once: tiles of 2^m amplitudes are gathered by TMA tensor copies into shared memory (triple buffered: load of tile
 * j+1 / sweeps of tile j / store of tile j-1 overlap), take every sweep of the pass, and are scattered back.
 * The host compiler (qsb/stream.py) packs gates into passes, turns every 1-qubit gate into a 2x2 that rides in front
 * of the next multi-qubit gate on its qubit, and groups consecutive gates whose qubits fit FOUR index bits into one
 * BLOCK SWEEP: one shared-memory round trip of the tile in which a worker holds the 16 amplitudes of those four bits in
 * registers and applies the block's whole op list (the round trips, not the arithmetic, bound a pass: ncu, profiles/).
 *
 * Slot bits: 0..l-1 = the low index bits (identity: one contiguous 16 * 2^l-byte row); l..l+e-1 ride in the TMA box
 * as extra dimensions (e <= 3; 16 * 2^(l+e) bytes per TMA op, >= 2 KiB for full bandwidth); l+e..m-1 number the TMA
 * ops of a tile (m - l - e <= 5); m..n-1 number the tiles.  positions[j] = bit of the amplitude index (of this
 * device's shard) that slot j stands for.  6 <= m <= 12, 3 <= l, n <= 30.  complex128 only. */
enum { QSB_CLS_NONE = 0, QSB_CLS_RDIAG = 1 /* diag(1, real) */, QSB_CLS_DIAG = 2, QSB_CLS_DENSE = 3 };
/* gate applied by a resident-executor sweep after the pending matrices of its bits (internal descriptor field) */
enum { QSB_G_NONE = 0, QSB_G_CX = 1, QSB_G_CZ = 2, QSB_G_SWAP = 3, QSB_G_CCX = 4, QSB_G_CSWAP = 5, QSB_G_DENSE = 6 };
/* ops of a block sweep; t[] are LOCAL bits 0..3 of the register block (local bit i = tile slot bit b[i]) */
enum {
  QSB_B_MAT1 = 1,     /* 2x2 U (class cls) on local bit t[0]                                           */
  QSB_B_CX = 2,       /* control t[0], target t[1]                                                    */
  QSB_B_CZ = 3,       /* t[0], t[1]                                                                   */
  QSB_B_SWAP = 4,     /* t[0], t[1]                                                                   */
  QSB_B_CCX = 5,      /* controls t[0] < t[1], target t[2]                                            */
  QSB_B_CSWAP = 6,    /* control t[0], swapped t[1] < t[2]                                            */
  QSB_B_DENSE2 = 7,   /* 4x4 at cdata[mat] on local bits (3, 2): local bit 3 = targets[0]              */
  QSB_B_DENSE3 = 8    /* 8x8 at cdata[mat] on local bits (3, 2, 1)                                     */
};
#define QSB_STREAM_MAX_BLOCKS 16
#define QSB_STREAM_BLOCK_OPS 12
typedef struct qsb_stream_op {
  int32_t kind;           /* QSB_B_*                                                                   */
  int32_t t[3];
  int32_t cls;            /* QSB_B_MAT1: QSB_CLS_RDIAG | QSB_CLS_DIAG | QSB_CLS_DENSE                  */
  int32_t pad[3];
  double U[8];            /* QSB_B_MAT1: row-major (re, im) 2x2                                        */
} qsb_stream_op;
typedef struct qsb_stream_block {
  int32_t n_ops;
  int32_t b[4];           /* four distinct tile slot bits < m (bits no op uses are filler)             */
  int32_t mat;            /* offset (doubles, even) of the dense matrix in cdata for DENSE2 / DENSE3 (at most one per block), else -1 */
  int32_t pad[2];
  qsb_stream_op ops[QSB_STREAM_BLOCK_OPS];
} qsb_stream_block;
typedef struct qsb_stream qsb_stream;
/* positions_out (NULL = positions): where the store puts slot j -- a pass whose store permutes index positions
 * (used to bring the qubits that leave in the next exchange to the top local positions) must run out of place. */
int qsb_stream_create(qsb_ctx* ctx, int32_t n, int32_t m, int32_t l, int32_t e, const int32_t* positions /*[n]*/,
                      const int32_t* positions_out /*[n] or NULL*/, const qsb_stream_block* blocks,
                      int32_t n_blocks, const double* cdata, int64_t n_cdata,
                      qsb_stream** out);
/* in / out: complex128[>= offset + 2^n] shards (out NULL or == in: in place); offsets in amplitudes, multiples of 8.
 * flags: QSB_RUN_ASYNC.  With peer_table (see qsb_run_args): tiles are LOADED from the peers' shards -- element s of
 * this device's post-exchange shard = peer (s >> peer_shift), offset (s & (2^peer_shift - 1)) | peer_rank_or -- which
 * folds the all-to-all qubit exchange into the pass; peers = host array of 2^(n - peer_shift) device pointers. */
int qsb_stream_run(qsb_stream* pass, qsb_buffer* in, int64_t in_offset, qsb_buffer* out, int64_t out_offset,
                   int32_t flags);
int qsb_stream_run_peers(qsb_stream* pass, const void* const* peers, int32_t n_peers, int32_t peer_shift,
                         int64_t peer_rank_or, qsb_buffer* out, int64_t out_offset, int32_t flags);
/* The exchange folded into the STORE of a pass (the compute + collective kernel of config 5): tiles are loaded from
 * `in`, swept, and every TMA box is stored straight into the shard of the peer its destination index names -- element x
 * of the pass's output goes to peers[x >> peer_shift] at offset (x & (2^peer_shift - 1)) | peer_rank_or -- posted
 * writes over NVLink that overlap the sweeps of the following tiles.  With positions_out != positions the same pass
 * also does the position reorder that brings the leaving qubits to the top.  peers = host array of the 2^(n -
 * peer_shift) ranks' destination shard pointers (mine included); none may be the shard being read. */
int qsb_stream_run_scatter(qsb_stream* pass, qsb_buffer* in, int64_t in_offset, const void* const* peers, int32_t n_peers,
                           int32_t peer_shift, int64_t peer_rank_or, int32_t flags);
int qsb_stream_free(qsb_stream* pass);

#ifdef __cplusplus
}
#endif
#endif /* QSB_H */
