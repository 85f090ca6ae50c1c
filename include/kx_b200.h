/*
 * kx_b200.h -- C ABI of the B200 (sm_100a) stencil engine behind kernex's kmap/kscan.
 *
 * This is the drop-in boundary for ONE path of ASEM000/kernex: the four engine
 * factories  kernel_map / offset_kernel_map (kernex/_src/map.py:61-124, :127-156)
 * and        kernel_scan / offset_kernel_scan (kernex/_src/scan.py:65-115, :118-147),
 * i.e. "apply a per-window function over the sliding windows of an N-D array",
 * which the reference executes as jnp.pad + int32 index views + gather + roll +
 * user function under jax.vmap / lax.scan (kernex/_src/utils.py:25-30, :74-90).
 *
 * Shape of the ABI
 *   - plain C: pointers, sizes, PODs; no torch / jax / C++ types.
 *   - every launcher only ENQUEUES work on the caller's stream: it never
 *     synchronises, never allocates device memory that outlives the call and never
 *     touches global mutable state other than the thread-local error string, so it
 *     is re-entrant and fits the XLA FFI handler contract
 *     (XLA_FFI_Error* handler(XLA_FFI_CallFrame*): buffers owned by XLA, results
 *     pre-allocated, work queued on ffi::PlatformStream<cudaStream_t>).
 *     INTEGRATION.md shows the jax.ffi shim and the ctypes stub.
 *   - device pointers are raw CUdeviceptr-compatible addresses; `stream` is a
 *     cudaStream_t passed as void*.
 *   - return value: 0 = KX_OK, otherwise a KxStatus; kx_last_error() returns the
 *     message of the last failure on the calling thread.
 *
 * A window function never crosses this boundary.  The host side (Python tracer,
 * kernex_b200/tracer.py) classifies it into one of three templates and describes
 * it with a KxProgram:
 *   AFFINE  out = bias + sum_t coef[t] * x[origin + tap[t]]      (Laplacian, conv-as-kmap, FDM updates, constants)
 *   REDUCE  out = max_t / min_t x[origin + tap[t]]                (window max/min; sum/mean are AFFINE)
 *   GATHER  out[e] = x[origin + tap[e]], e < n_taps              (patch extraction, kmap(lambda x: x))
 * Reads outside [0, n) return 0 (the reference's zero padding,
 * kernex/_src/utils.py:43-53 + jnp.pad in map.py:85); negative borders crop.
 */
#ifndef KX_B200_H_
#define KX_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KX_ABI_VERSION 2

#define KX_MAX_DIMS 4      /* leading batch axes are folded by the host            */
#define KX_MAX_TAPS 1024   /* total taps over all functions of one program (a (64,3,3) conv window has 576) */
#define KX_MAX_FUNCS 8     /* functions of one F[slice] = fn mesh                 */
#define KX_MAX_KEYS 32     /* total region keys over all functions                */

typedef enum KxStatus {
  KX_OK = 0,
  KX_ERR_INVALID = 1,      /* malformed geometry / program                         */
  KX_ERR_UNSUPPORTED = 2,  /* valid, but no kernel template covers it              */
  KX_ERR_CUDA = 3,         /* a CUDA runtime/driver call failed                    */
  KX_ERR_NO_DEVICE = 4
} KxStatus;

typedef enum KxDType { KX_F32 = 0, KX_I32 = 1 } KxDType;

typedef enum KxKind { KX_AFFINE = 0, KX_REDUCE_MAX = 1, KX_REDUCE_MIN = 2, KX_GATHER = 3 } KxKind;

typedef enum KxKeyKind { KX_KEY_ANY = 0, KX_KEY_EQ = 1, KX_KEY_RANGE = 2, KX_KEY_RANGE_STEP = 3 } KxKeyKind;

/* Window geometry; replaces the int32 "views" of kernex/_src/utils.py:74-90,205-280.
 * Window j along axis d covers input coordinates [-lo[d] + j*stride[d], ... + ksize[d]);
 * out_shape[d] = (in_shape[d] + lo[d] + hi[d] - ksize[d]) / stride[d] + 1
 * (kernex/_src/utils.py:93-112).  `centre[d]` is (ksize[d]-1)/2 (utils.py:56-71). */
typedef struct KxGeom {
  int32_t ndim;
  int32_t in_dtype;                 /* KxDType of the input array                   */
  int32_t out_dtype;                /* KxDType of the result (f32 when an i32 input meets float taps) */
  int32_t reserved;
  int64_t in_shape[KX_MAX_DIMS];
  int64_t out_shape[KX_MAX_DIMS];
  int32_t ksize[KX_MAX_DIMS];
  int32_t stride[KX_MAX_DIMS];
  int32_t lo[KX_MAX_DIMS];          /* left border (may be negative = crop)          */
  int32_t hi[KX_MAX_DIMS];
} KxGeom;

/* One dimension of one region key (kernex/_src/utils.py:312-333): int -> EQ,
 * (a,b) -> RANGE, (a,b,step) -> RANGE_STEP (tests centre % step == 0), inf -> ANY. */
typedef struct KxKeyItem {
  int32_t kind;
  int32_t step;
  int64_t a, b;
} KxKeyItem;

typedef struct KxKey {
  int32_t n_items;                  /* dims present in the key; trailing dims are free (zip truncation) */
  int32_t func;                     /* owning function (insertion index)             */
  KxKeyItem item[KX_MAX_DIMS];
} KxKey;

typedef struct KxFunc {
  int32_t kind;                     /* KxKind                                        */
  int32_t tap_begin, tap_end;       /* slice of KxProgram.tap_off / coef             */
  int32_t coef_is_int;              /* coefficients & bias are exact integers        */
  double bias;
  double post_div;                  /* 0 = none; else result = (bias + sum) / post_div (window mean) */
} KxFunc;

/* The classified window functions of one kmap/kscan call.  Function selection per
 * window follows kernex/_src/utils.py:338-382 + lax.switch's clamp: functions are
 * tried from the LAST inserted to the first, a function matches if any of its keys
 * matches the window centre, no match selects function 0. */
typedef struct KxProgram {
  int32_t n_funcs;
  int32_t n_keys;
  int32_t n_taps;
  int32_t reserved;
  KxFunc func[KX_MAX_FUNCS];
  KxKey key[KX_MAX_KEYS];
  int16_t tap_off[KX_MAX_TAPS][KX_MAX_DIMS]; /* offset from the window ORIGIN, 0 <= off < ksize */
  double coef[KX_MAX_TAPS];                  /* host-known coefficients (cast to the compute type) */
} KxProgram;

int kx_abi_version(void);
const char* kx_last_error(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
uint64_t kx_launch_count(void);
/* name of the kernel template the last launcher in this process chose (diagnostics) */
const char* kx_last_kernel(void);

/* ---- kmap -------------------------------------------------------------------
 * Replaces the closure returned by kernel_map (kernex/_src/map.py:84-122):
 * jnp.pad -> vmap(gather -> roll_view -> fn) -> reshape.
 *   in        device, geom.in_shape, C-contiguous, geom.in_dtype
 *   coef_dev  optional device array [prog.n_taps] of the COMPUTE type; when non-null
 *             it replaces prog.coef (runtime weights: conv-as-kmap `sum(x*w)`)
 *   out       device, geom.out_shape (+ [n_taps] trailing axis for KX_GATHER)
 * `batch` independent arrays laid out back to back are processed by one launch
 * (jax.vmap over a leading axis, README.md:310-336); per-batch coefficient rows of
 * length prog.n_taps are used when coef_batch_stride != 0.
 */
int kx_map(const KxGeom* geom, const KxProgram* prog, const void* in, const void* coef_dev,
           int64_t coef_batch_stride, void* out, int64_t batch, void* stream);

/* offset_kernel_map (kernex/_src/map.py:127-156): the result of a scalar kmap written into a copy of
 * the input at the strided interior [off0[d] : off0[d] + out_shape[d]*stride[d] : stride[d]].
 * `out` has geom.in_shape and geom.in_dtype (== out_dtype) and is WRITTEN COMPLETELY by the call: the
 * launcher either runs one fused pass (stride 1, single AFFINE function: every cell is produced by
 * the stencil kernel, cells outside the interior copy their input value) or enqueues the
 * device-to-device copy itself followed by the strided write-back kernel. */
/* int32 -> float32 conversion of a whole device array (16-byte aligned), round to nearest even.  JAX converts an int32
 * window before it applies a float-valued function (weak-type promotion, SURVEY 9.8 R: jnp.mean of integers, Python
 * float coefficients; kernex/_src/map.py:46 gathers the patch, the user function promotes it), so the host side
 * converts once and runs the float32 templates. */
int kx_cast_i32_f32(const void* in, void* out, int64_t n, void* stream);

int kx_map_offset(const KxGeom* geom, const KxProgram* prog, const void* in, const void* coef_dev,
                  const int32_t* off0, void* out, void* stream);

/* ---- VJP of the AFFINE template --------------------------------------------
 * JAX derives these by transposing gather/pad (tests/test_interface.py:345-348).
 * dx[i]  = sum_{j,t : origin(j)+tap[t] = i} coef[t] * g[j]     (flipped-tap stencil)
 * dcoef[t] = sum_j g[j] * x[origin(j) + tap[t]]                (window correlation)
 * `scratch` (device, kx_vjp_coef_scratch_bytes) holds per-block partials; the
 * two-stage reduction is deterministic. */
int kx_map_vjp_input(const KxGeom* geom, const KxProgram* prog, const void* g, const void* coef_dev,
                     void* dx, void* stream);
int64_t kx_vjp_coef_scratch_bytes(const KxGeom* geom, const KxProgram* prog);
int kx_map_vjp_coef(const KxGeom* geom, const KxProgram* prog, const void* x, const void* g,
                    void* scratch, void* dcoef, void* stream);

/* ---- kscan ------------------------------------------------------------------
 * Replaces the closure returned by kernel_scan (kernex/_src/scan.py:84-113): a
 * lax.scan over all windows in row-major order carrying the zero-padded array;
 * each value is written at the window centre before the next window is read and
 * the written values are returned.
 *   state  device scratch: the padded array, kx_scan_state_bytes(); the launcher
 *          initialises it from `in` itself
 *   out    device, geom.out_shape, geom.in_dtype
 * The launcher analyses the tap offsets: if no function reads the row
 * being written (all taps at axis-0 offset < centre, or never-written cells) the
 * rows are independent given the previous ones and a row-marching kernel
 * parallelises each row; otherwise the wavefront kernels run every hyperplane of
 * mutually independent windows in parallel (warp-skewed for 2-D in-row dependencies),
 * which reproduces the exact sweep order's values.  3-D scans whose functions read only the PLANE above the cell they
 * write (time marching with two space axes, tests/test_interface.py:28-97) run as one 2-D kmap launch per plane.
 * `reverse` = lax.scan(reverse=True).
 */
int64_t kx_scan_state_bytes(const KxGeom* geom);
int kx_scan(const KxGeom* geom, const KxProgram* prog, const void* in, const void* coef_dev,
            void* state, void* out, int32_t reverse, void* stream);

/* Offset scan (sscan) in ONE call: replaces offset_kernel_scan (kernex/_src/scan.py:118-147) -- kernel_scan over the
 * offset interior followed by `.at[set_indices].set(result)` into a copy of the input.  `out` has geom.in_shape; off0[d]
 * is the low offset of axis d (geom.lo / geom.hi carry the borders _offset_to_padding derives from the offsets).
 * Implemented for 3-D time-marching scans (one AFFINE function reading only the plane above the cell it writes,
 * tests/test_interface.py:28-97): one fused 2-D pass per plane.  Anything else returns KX_ERR_UNSUPPORTED and the
 * caller composes kx_scan with a strided write-back. */
int kx_scan_offset(const KxGeom* geom, const KxProgram* prog, const void* in, const void* coef_dev, const int32_t* off0,
                   void* out, void* stream);

/* Axis-1 shard of a row-marching kscan (time-marching stencils, README.md:231-266): produces
 * columns [col_base, col_base + geom.in_shape[1]) of the scan kernel_scan would run on a
 * (geom.in_shape[0], global_nx) array; region keys and the zero pad are in GLOBAL columns.
 * Axis 0 is the sequential axis, so a multi-GPU kscan shards columns; a shard that does not
 * touch the global edge must be over-allocated by nt*R columns on that side (R = the largest
 * column offset the functions read) and the caller drops them: inside that cone the values
 * are computed from cells the shard does not hold.  Returns KX_ERR_UNSUPPORTED unless the
 * dependency analysis of kx_scan's fast path applies (every tap reads the row above). */
int kx_scan_rowmarch_shard(const KxGeom* geom, const KxProgram* prog, int64_t col_base, int64_t global_nx,
                           void* out, void* stream);

/* The same shard advanced over rows [row_begin, row_end) only: row row_begin-1 of `out` (all held
 * columns, ghost columns included) must already be final.  This is the building block of the
 * halo-EXCHANGE form of the multi-GPU kscan: a shard over-allocated by only
 * T*R ghost columns per side, T = kx_scan_rowmarch_rows_per_launch(), receives row n0-1 of its ghost
 * columns from the neighbouring rank (T*R*4 bytes, NCCL send/recv) before each T-row block; inside the
 * block the ghost columns decay with the dependency cone exactly like the per-CTA halo does. */
int kx_scan_rowmarch_rows_per_launch(void);
int kx_scan_rowmarch_shard_rows(const KxGeom* geom, const KxProgram* prog, int64_t col_base, int64_t global_nx,
                                int32_t row_begin, int32_t row_end, void* out, void* stream);

/* ---- multi-GPU halo transport (axis-0 slab decomposition, one process per GPU) -----------------------------
 * The reference has no multi-device data path to mirror: kernex/_src/utils.py:25-30 only offers jax.pmap with one
 * WINDOW per device.  The slab decomposition of kernex_b200/dist.py keeps every rank's rows in a buffer allocated
 * here (cudaMalloc), exports it with CUDA IPC and lets the neighbouring ranks map it, so the halo rows travel over
 * NVLink by ONE one-sided kernel per step (kx_halo_pull) instead of a two-sided NCCL send/recv.
 * A slab buffer starts with KX_PEER_FLAG_BYTES of zero-initialised 64-bit flags (monotonic epochs). */
#define KX_PEER_HANDLE_BYTES 64
#define KX_PEER_FLAG_BYTES 256
int kx_peer_alloc(int64_t bytes, void** ptr);             /* zero-filled device buffer on the current device       */
int kx_peer_free(void* ptr);
int kx_peer_export(const void* ptr, void* handle64);      /* KX_PEER_HANDLE_BYTES opaque bytes for another PROCESS  */
int kx_peer_open(const void* handle64, void** ptr);       /* map a neighbour's buffer (peer access enabled lazily)  */
int kx_peer_close(void* ptr);

typedef struct KxHaloSide {
  void* dst;             /* my halo rows (local device memory)                                              */
  const void* src;       /* the neighbour's edge rows (peer-mapped); bytes == 0: nothing to pull from it    */
  int64_t bytes;         /* multiple of 16, both pointers 16-byte aligned                                   */
  uint64_t* peer_ready;  /* neighbour's flag: set to `epoch` = "my own rows are final"                      */
  const uint64_t* my_ready; /* my flag the neighbour sets the same way; waited for when bytes > 0           */
  uint64_t* peer_done;   /* neighbour's flag: set to `epoch` = "I have read your rows, overwrite them"       */
} KxHaloSide;

typedef struct KxHaloPull {
  int32_t n_sides;       /* 0..2 neighbours                                                                  */
  int32_t accumulate;    /* 0: dst = src (halo exchange); 1: dst += src as fp32 (its adjoint: cotangent rows that landed in
                            the neighbour's halo are added to my edge rows)                                  */
  uint64_t epoch;        /* strictly increasing per step, the same sequence on every rank                    */
  uint32_t* counter;     /* local scratch word, zero between launches                                        */
  KxHaloSide side[2];
} KxHaloPull;

/* signal READY to the neighbours, wait for theirs, pull their edge rows, signal DONE -- one launch on `stream`.
 * A neighbour that never signals traps the kernel after 20 s (the error surfaces at the next synchronisation). */
int kx_halo_pull(const KxHaloPull* p, void* stream);
/* enqueue a wait until *flag >= value (before own rows are overwritten: the neighbours' DONE flags) */
int kx_flag_wait(const uint64_t* flag, uint64_t value, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* KX_B200_H_ */
