/*
 * icp_b200.h — C ABI of the B200-native ICP iteration loop (libicp_b200.so).
 *
 * Drop-in boundary for tier4/icp_rust's estimation API.  Each entry point names the
 * reference interface it replaces (file:line relative to the icp_rust repository); the Rust
 * `extern "C"` block a maintainer adds to src/lib.rs is shown in INTEGRATION.md.
 *
 * Conventions
 *  - Points are array-of-structs f64 exactly as `&[Vector2]` / `&[Vector3]` are laid out
 *    (src/types.rs:4-12): x0 y0 x1 y1 ... or x0 y0 z0 x1 ...
 *  - A `Transform` (src/transform.rs:6-10, no #[repr(C)]) crosses the boundary as 6 doubles
 *    [r00 r01 r10 r11 tx ty]: the full 2x2 rotation matrix, row-major, then the translation.
 *  - Batches are ragged: pair b owns points [off[b], off[b+1]) of the concatenated array.
 *  - Every function returns an icp_b200_status.  Nothing aborts; where the reference would
 *    panic (empty target: src/lib.rs:122,165; NaN: src/stats.rs:12) the pair's status says so
 *    and its output transform equals its initial transform.
 *  - There is no CPU fallback: without a CUDA device every compute entry returns
 *    ICP_B200_ERR_NO_DEVICE / ICP_B200_ERR_CUDA.
 *  - Threading: a context owns one device, its streams and scratch memory (work counters, the
 *    re-run list of the guard, per-CTA slabs).  Calls on ONE context must be serialised by the
 *    caller, and that includes work still in flight: after an asynchronous device-pointer call
 *    (synchronize == 0) the next call on the same context must be issued on the SAME stream, or
 *    after that stream has been synchronised.  Different contexts (also on the same device) are
 *    independent and may be used from different threads: one context per thread.
 *  - Host-pointer entries copy with cudaMemcpyAsync on two streams so that transfers overlap
 *    compute.  That overlap needs PINNED (page-locked) caller buffers (cudaHostAlloc /
 *    cudaHostRegister); pageable memory works but every copy then serialises.
 *  - The reference's outer loop has no exit (src/lib.rs:112-128).  An inner solve that accepts no
 *    update returns the identity, so the outer transform, the correspondences and the next solve
 *    stay what they were: every later iteration is an exact repeat.  The kernels write those
 *    iterations' outputs (same correspondences and iteration word per iteration, checksum advanced
 *    per iteration) instead of recomputing them; all outputs keep the bits of the full computation.
 *    ICP_B200_FIXED_POINT=0 in the environment computes every iteration.
 *  - Parity: results are checked against oracle/, a C restatement of the crate.  Its median is the
 *    mathematically exact order statistic ((b + c) / 2 of the true middle pair for even n);
 *    src/stats.rs:23-26 reads input[n/2 - 1] after a second select_nth_unstable_by, which the
 *    Rust standard library does not promise to leave in place — rust/tests/dump_golden.rs prints
 *    what a real toolchain does so that a maintainer can pin it (DESIGN.md "Oracle").
 */
#ifndef ICP_B200_H
#define ICP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ICP_B200_ABI_VERSION 4

typedef enum {
    ICP_B200_OK = 0,
    ICP_B200_ERR_INVALID_ARGUMENT = 1,
    ICP_B200_ERR_CUDA = 2,
    ICP_B200_ERR_NO_DEVICE = 3,
    ICP_B200_ERR_EMPTY_TARGET = 4, /* a pair had no target points but source points (reference panics) */
    ICP_B200_ERR_NON_FINITE = 5,   /* a pair held NaN/Inf coordinates or transform (reference panics on NaN) */
    ICP_B200_ERR_TOO_LARGE = 6,    /* a scan exceeds what the device kernels of this build hold */
    ICP_B200_ERR_OUT_OF_MEMORY = 7
} icp_b200_status;

/* per-pair status values written to icp_b200_batch.out_status */
#define ICP_B200_PAIR_OK 0
#define ICP_B200_PAIR_EMPTY_TARGET 1
#define ICP_B200_PAIR_TOO_LARGE 2
#define ICP_B200_PAIR_NON_FINITE 3

/* break reasons of the inner IRLS loop, src/lib.rs:66-82 (upper half of out_inner words) */
#define ICP_B200_BREAK_CAP 0   /* 200 iterations exhausted */
#define ICP_B200_BREAK_NONE 1  /* weighted_gauss_newton_update returned None */
#define ICP_B200_BREAK_DELTA 2 /* delta.dot(delta) < 1e-6 */
#define ICP_B200_BREAK_ERROR 3 /* huber_error increased */

#define ICP_B200_CORR_NONE 0
#define ICP_B200_CORR_LAST 1   /* out_corr[off_src[b] + i]: index chosen in the last outer iteration */
#define ICP_B200_CORR_ALL 2    /* out_corr[off_src[b]*max_iter + it*n_b + i] */

/* icp_b200_batch.flags
 * Default (0): the normal-equation sums J^T W J, J^T W r and the Huber cost are reduced with warp
 * shuffles in a fixed order; they differ from the reference's sequential fold by summation rounding
 * (~1e-16 relative), so transforms agree to ~1e-13, far inside the 1e-9 contract.  The inner loop's
 * decisions (det == 0 src/linalg.rs:18, delta.dot(delta) < 1e-6 src/lib.rs:71, error > prev_error
 * src/lib.rs:76) are discontinuous, so they are GUARDED: a pair with a decision quantity within a
 * relative band of its threshold (1e-10 for det and the error test, 1e-6 for the delta test — bands
 * at least a thousand times wider than the perturbation a different summation order can cause) is
 * re-run by the library, inside the same call, with ORDERED_SUMS.  Iteration counts and
 * correspondences therefore equal the ordered mode's: a pair either took every decision with a
 * margin the summation order cannot bridge, or was recomputed in the reference's order.
 * icp_b200_call_info.redo_pairs counts the re-runs (about 1 pair in 40000 on BASELINE config 2; a lone
 * 1080-point pair takes ~5 ms in reference order, which host-pointer calls hide beside their later chunks).
 * Results are deterministic (independent of scheduling) but, for unflagged pairs, depend in their
 * last bits on the kernel instance chosen for the scan size.
 * ORDERED_SUMS: every sum is accumulated in the reference's own order (src/lib.rs:242-255,45-50):
 * transforms are bit-identical to the reference arithmetic by construction, at roughly a third of
 * the throughput.  Any scan size (large scans stream their row terms through shared-memory tiles).
 * NO_GUARD: default mode without the guard and the re-run (measurement / diagnosis only).
 * SHARED_SOURCE: every pair aligns the SAME source scan src[0 .. src_off[1]) (src_off then holds
 * two entries) — the multi-hypothesis form: one scan, `batch` initial transforms, no copies.
 * out_corr (if any) is laid out as if the scan were repeated: pair b starts at b * n.
 * INDEPENDENT (device-pointer calls with synchronize == 0; ignored elsewhere): the caller promises
 * that this call neither reads nor overwrites anything the PREVIOUS icp_b200 call on the stream
 * writes (its out_* arrays).  The library may then let this call's thread blocks start on the SMs
 * the previous call's last pairs leave idle (programmatic dependent launch) instead of waiting for
 * its last pair — about 8 % more pairs/s for back-to-back 4096-pair batches, where the last of ten
 * rounds of 444 resident pairs is a quarter full.  Everything else on the stream (copies, other
 * kernels, events) keeps plain stream order with respect to both calls.  ICP_B200_PDL=0 in the
 * environment makes the flag a no-op. */
#define ICP_B200_FLAG_ORDERED_SUMS 1u
#define ICP_B200_FLAG_SHARED_SOURCE 2u
#define ICP_B200_FLAG_NO_GUARD 4u
#define ICP_B200_FLAG_INDEPENDENT 8u

typedef struct icp_b200_context icp_b200_context;
typedef struct icp_b200_target icp_b200_target;

/* One batched call = `batch` independent Icp2d::estimate / Icp3d::estimate calls
 * (src/lib.rs:105-110, 148-153).  batch == 1 reproduces the reference call exactly. */
typedef struct {
    uint32_t dim;            /* 2 -> Icp2d, 3 -> Icp3d (NN in xyz, solve in xy: src/lib.rs:148-173) */
    uint32_t batch;          /* number of scan pairs B */
    uint32_t max_iter;       /* `max_iter` argument of estimate (outer iterations, no early exit) */
    int32_t corr_mode;       /* ICP_B200_CORR_* */
    uint32_t max_src_points; /* largest src scan of the batch; 0 = let the library find out (the */
    uint32_t max_dst_points; /* device-pointer entries then read the offset tables back once)    */
    const double *src;       /* `src: &[VectorD]` of every pair, concatenated */
    const uint32_t *src_off; /* [B+1] */
    const double *dst;       /* `dst` given to Icp2d::new / Icp3d::new, concatenated; ignored by
                                icp_b200_target_estimate_batch */
    const uint32_t *dst_off; /* [B+1] */
    const double *init_T;    /* [B*6] `initial_transform` */
    double *out_T;           /* [B*6] returned `Transform` */
    uint32_t *out_inner;     /* nullable [B*max_iter]: accepted IRLS updates | break reason << 16 */
    int32_t *out_corr;       /* nullable, see ICP_B200_CORR_* */
    uint64_t *out_corr_hash; /* nullable [B]: order-independent checksum of every outer
                                iteration's correspondences (definition in DESIGN.md) */
    int32_t *out_status;     /* nullable [B]: ICP_B200_PAIR_* */
    uint32_t flags;          /* ICP_B200_FLAG_* */
    double *out_cost;        /* nullable [B]: huber_error (src/lib.rs:45-50) of the last inner solve's
                                transform over the last outer iteration's transformed source points and
                                their matched target points — the cost the loop itself minimises; 0 for
                                rejected pairs and max_iter == 0 */
} icp_b200_batch;

/* counters of the most recent compute call on a context */
typedef struct {
    uint32_t kernel_launches;  /* kernels of this library launched by the call */
    float kernel_ms;           /* device time of those kernels, CUDA events on the launching stream */
    uint64_t h2d_bytes;        /* bytes the call copied host -> device */
    uint64_t d2h_bytes;        /* bytes the call copied device -> host */
    uint32_t grid_ctas, block_threads, smem_bytes; /* launch shape of the ICP kernel */
    uint32_t redo_pairs;       /* pairs the knife-edge guard re-ran with ordered sums (synchronous calls) */
} icp_b200_call_info;

int icp_b200_abi_version(void);
const char *icp_b200_status_string(int status);
int icp_b200_device_count(int *count);

/* A context binds one CUDA device.  `device` is the CUDA ordinal. */
int icp_b200_create(int device, icp_b200_context **out);
void icp_b200_destroy(icp_b200_context *ctx);
int icp_b200_last_call_info(const icp_b200_context *ctx, icp_b200_call_info *info);
const char *icp_b200_last_error(const icp_b200_context *ctx);

/* Icp2d::new(dst).estimate(src, init, max_iter) for every pair; HOST pointers.  Inputs are
 * copied to the device in chunks overlapped with compute; results are copied back before
 * the call returns.  Replaces src/lib.rs:97-130 and src/lib.rs:139-173. */
int icp_b200_estimate_batch(icp_b200_context *ctx, const icp_b200_batch *batch);

/* The same call in two halves, for ONE host thread that streams batches (a scan front end calling
 * Icp2d::estimate frame after frame, examples/scan2d.rs:84-90, batched): `submit` validates, enqueues every
 * copy and kernel of the batch and returns without waiting; `wait` completes it (the guard's re-run, the
 * status word, the last copies) and returns what icp_b200_estimate_batch would have returned.
 * submit + wait IS icp_b200_estimate_batch.  Between the two, on that context: no other compute call
 * (INVALID_ARGUMENT), and the batch's arrays — inputs, outputs and the offset tables, PINNED — stay valid and
 * untouched; the icp_b200_batch struct itself is copied.  With two contexts a thread keeps the device busy
 * across batch boundaries: submit(A, k) ... submit(B, k+1); wait(A); submit(A, k+2); wait(B) ... — batch
 * k+1's copies and first pairs run under batch k's last pairs, as two threads calling
 * icp_b200_estimate_batch on a context each would achieve.  A batch that holds a pair the knife-edge guard
 * re-runs is waited for about a batch time longer (the ordered re-run shares its SM with the next batch); with
 * THREE contexts in rotation (submit k+2 before wait k) the device stays busy through that as well — the form
 * bench.py's e2e figure uses.  wait without a submitted batch returns OK.  icp_b200_last_call_info is complete
 * after wait. */
int icp_b200_estimate_batch_submit(icp_b200_context *ctx, const icp_b200_batch *batch);
int icp_b200_estimate_batch_wait(icp_b200_context *ctx);

/* Same with DEVICE pointers (inputs resident in HBM, outputs written to HBM).  Enqueued on
 * `stream` (a cudaStream_t, NULL = the context's stream) and synchronised before returning
 * only if `synchronize` is non-zero.  The call's status word (worst pair status) is read back by
 * synchronous calls only and is sticky: a synchronous call also reports rejected pairs of the
 * asynchronous calls before it (out_status always says which pairs).  No host-side memset or
 * event sits between two asynchronous calls, so they run back to back (see
 * ICP_B200_FLAG_INDEPENDENT). */
int icp_b200_estimate_batch_device(icp_b200_context *ctx, const icp_b200_batch *batch,
                                   void *stream, int synchronize);

/* Single-pair wrappers with the reference's argument order (default reduction mode: guarded warp-tree sums,
 * transform within 1e-9 of the reference arithmetic; environment ICP_B200_SINGLE_ORDERED=1 when the context is
 * created: ordered sums, bit-identical transform, about twice the latency):
 *   Icp2d::new(dst) (src/lib.rs:97) + Icp2d::estimate(src, initial_transform, max_iter)
 *   (src/lib.rs:105-110); Icp3d twins (src/lib.rs:139, 148-153). */
int icp_b200_icp2d_estimate(icp_b200_context *ctx, const double *dst_xy, uint32_t n_dst,
                            const double *src_xy, uint32_t n_src, const double initial_transform[6],
                            uint32_t max_iter, double out_transform[6]);
int icp_b200_icp3d_estimate(icp_b200_context *ctx, const double *dst_xyz, uint32_t n_dst,
                            const double *src_xyz, uint32_t n_src, const double initial_transform[6],
                            uint32_t max_iter, double out_transform[6]);

/* Reusable target: the `Icp2d<'a>` / `Icp3d<'a>` value itself (src/lib.rs:91-102, 133-144).
 * Builds the search structure once on the device; any number of estimates and searches may
 * follow.  `dst` is a HOST pointer and is copied (the reference borrows it for 'a). */
int icp_b200_target_create(icp_b200_context *ctx, uint32_t dim, const double *dst, uint32_t n_dst,
                           icp_b200_target **out);
void icp_b200_target_destroy(icp_b200_target *target);

/* The same target with a changed point set, without building a new one — `Icp2d::new` on the next frame's
 * map (the reference rebuilds its kd-tree per frame: src/lib.rs:97-102, examples/scan2d.rs:87).  New point
 * set = the old one without the points `remove_idx` names (the survivors keep their order, so their indices
 * close up), then the `n_add` points of `add` (HOST pointers) appended: searches and estimates return
 * indices into that new array.  While the added points stay inside the bounding box of the last full build
 * and the point count within a factor two of it, the grid geometry, the cell tables and the seeds are kept
 * and only re-sorted / carried over (see icp_b200.cu); otherwise the structure is rebuilt.  Either way every
 * later search is exact.  Not to be called while work using the target is in flight. */
int icp_b200_target_update(icp_b200_target *target, const double *add, uint32_t n_add,
                           const uint32_t *remove_idx, uint32_t n_remove);

/* estimate() of every `src` scan against the one shared target; HOST pointers
 * (batch->dst / dst_off are ignored). */
int icp_b200_target_estimate_batch(icp_b200_target *target, const icp_b200_batch *batch);
/* submit half for a shared target; completed by icp_b200_estimate_batch_wait on the target's context */
int icp_b200_target_estimate_batch_submit(icp_b200_target *target, const icp_b200_batch *batch);
int icp_b200_target_estimate_batch_device(icp_b200_target *target, const icp_b200_batch *batch,
                                          void *stream, int synchronize);

/* The correspondence search alone: KdTree::search(&q).0 for q in queries
 * (src/lib.rs:121,164).  out_index[i] = index into the target's `dst`; -1 for a non-finite query
 * (the reference would panic).  Device form: 2-D queries must be 16-byte aligned. */
int icp_b200_target_search(icp_b200_target *target, const double *queries, uint64_t n_queries,
                           int32_t *out_index);
int icp_b200_target_search_device(icp_b200_target *target, const double *queries, uint64_t n_queries,
                                  int32_t *out_index, void *stream, int synchronize);

/* calc_stddevs over ragged residual sets (src/stats.rs:49-60): out_sigma[b*2 + axis].
 * HOST pointers; residuals are AoS (x y).  Exposes the device selection for parity tests. */
int icp_b200_calc_stddevs_batch(icp_b200_context *ctx, const double *residuals_xy,
                                const uint32_t *off, uint32_t batch, double *out_sigma,
                                double *out_median);

#ifdef __cplusplus
}
#endif
#endif /* ICP_B200_H */
