/*
 * icp_oracle.h — CPU ORACLE for the ICP iteration-loop hot path of tier4/icp_rust.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (icp_rust_b200/, include/) never includes, links or calls anything in oracle/.
 *
 * It is a plain-C restatement of the reference's algorithm, function by function,
 * each citing the reference file:line it follows (paths relative to /root/reference).
 * All arithmetic is IEEE-754 binary64, evaluated left to right exactly as the Rust
 * source writes it, compiled with -ffp-contract=off (Rust never fuses mul+add).
 *
 * PARITY STATUS
 *  - Pinned: every known-answer test the reference holds for this path
 *    (stats.rs, huber.rs, linalg.rs, se2.rs, so2.rs, transform.rs, lib.rs tests) is
 *    replayed against this file by tests/test_oracle_golden.py.
 *  - Unpinned (the reference cannot be built here: no cargo/rustc, and the kd-tree
 *    crate `nearest_neighbor` (git branch main, no rev, no Cargo.lock) is absent):
 *      * nearest-neighbour tie-break.  Contract adopted: exact Euclidean 1-NN on
 *        d2 = dx*dx + dy*dy (+ dz*dz), lowest index wins on equal d2.
 *      * last-ulp behaviour of sin/cos.  Restated from the published fdlibm/musl
 *        algorithm that Rust's `libm` crate (num-traits 0.2 feature "libm",
 *        Cargo.toml:17-20) ports; checked against glibc to <= 1 ulp.
 *
 * Layouts: points are AoS f64 (x0 y0 x1 y1 ... / x0 y0 z0 ...), as &[Vector2]/&[Vector3]
 * are in memory (src/types.rs:4-12).  A Transform is 6 doubles
 * [r00 r01 r10 r11 tx ty] (src/transform.rs:6-10, rotation kept as a full 2x2).
 * 3x3 matrices are row-major double[9].
 */
#ifndef ICP_ORACLE_H
#define ICP_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_HUBER_K 1.345 /* src/lib.rs:32 */

/* break reasons of the inner IRLS loop (src/lib.rs:66-82) */
enum {
    ORC_BREAK_CAP = 0,      /* 200 iterations exhausted (src/lib.rs:61,66) */
    ORC_BREAK_NONE = 1,     /* weighted_gauss_newton_update returned None (src/lib.rs:67-69) */
    ORC_BREAK_DELTA = 2,    /* delta.dot(delta) < 1e-6 (src/lib.rs:71-73) */
    ORC_BREAK_ERROR = 3     /* error > prev_error (src/lib.rs:76-78) */
};

/* ---- libm restatement (num-traits "libm" -> Rust libm crate = musl/fdlibm port) ---- */
double orc_sin(double x);
double orc_cos(double x);

/* ---- src/huber.rs ---- */
double orc_huber_rho(double e, double k);   /* huber.rs:6-15 */
double orc_huber_drho(double e, double k);  /* huber.rs:17-26 */

/* ---- src/stats.rs ----  return 1 = Some(*out), 0 = None */
int orc_median(double *buf, size_t n, double *out);            /* stats.rs:11-28 (mutates buf) */
int orc_mad(double *buf, size_t n, double *out);               /* stats.rs:30-37 */
int orc_standard_deviation(double *buf, size_t n, double *out);/* stats.rs:39-47 */
int orc_calc_stddevs(const double *residuals, size_t n, int dim, double *out); /* stats.rs:49-60 */

/* ---- src/linalg.rs ---- return 1 = Some, 0 = None */
int orc_inverse3x3(const double m[9], double out[9]);          /* linalg.rs:3-29 */

/* ---- src/so2.rs / src/se2.rs ---- */
void orc_so2_new_rotation2(double theta, double rot[4]);       /* so2.rs:8-17 */
double orc_so2_log(const double rot[4]);                       /* so2.rs:19-21 */
void orc_so2_exp(double theta, double rot[4]);                 /* so2.rs:23-31 */
void orc_se2_get_rt(const double m[9], double rot[4], double t[2]); /* se2.rs:11-19 */
void orc_se2_calc_rt(const double param[3], double rot[4], double t[2]); /* se2.rs:21-41 */
void orc_se2_exp(const double param[3], double out[9]);        /* se2.rs:43-52 */
void orc_se2_log(const double m[9], double out[3]);            /* se2.rs:54-77 */

/* ---- src/transform.rs ---- T = [r00 r01 r10 r11 tx ty] */
void orc_transform_new(const double param[3], double T[6]);    /* transform.rs:13-16 */
void orc_transform_identity(double T[6]);                      /* transform.rs:34-39 */
void orc_transform_apply(const double T[6], const double p[2], double out[2]); /* transform.rs:22-24 */
void orc_transform_inverse(const double T[6], double out[6]);  /* transform.rs:26-32 */
void orc_transform_mul(const double A[6], const double B[6], double out[6]);   /* transform.rs:42-51 */

/* ---- src/norm.rs ---- */
double orc_norm(const double *v, size_t n);                    /* norm.rs:8-21 */

/* ---- src/lib.rs solver ---- */
void orc_residual(const double T[6], const double s[2], const double d[2], double out[2]); /* lib.rs:34-36 */
double orc_error(const double T[6], const double *src, const double *dst, size_t n);       /* lib.rs:38-43 */
double orc_huber_error(const double T[6], const double *src, const double *dst, size_t n); /* lib.rs:45-50 */
int orc_gauss_newton_update(const double T[6], const double *src, const double *dst, size_t n,
                            double delta[3]);                  /* lib.rs:191-216 */
int orc_weighted_gauss_newton_update(const double T[6], const double *src, const double *dst,
                                     size_t n, double delta[3]);/* lib.rs:218-261 */
/* accepted (nullable) = number of applied updates; reason (nullable) = ORC_BREAK_* */
void orc_estimate_transform(const double *src, const double *dst, size_t n, double T[6],
                            uint32_t *accepted, uint32_t *reason); /* lib.rs:59-84 */

/* ---- nearest neighbour (external crate `nearest_neighbor`; contract in header comment) ---- */
int64_t orc_nn_brute(const double *dst, size_t m, int dim, const double *q);
typedef struct orc_kdtree orc_kdtree;
orc_kdtree *orc_kdtree_new(const double *dst, size_t m, int dim);  /* KdTree::new(dst, 1), lib.rs:99,141 */
void orc_kdtree_free(orc_kdtree *t);
int64_t orc_kdtree_search(const orc_kdtree *t, const double *q);   /* .search(&sp).0, lib.rs:121,164 */
/* batched query helper: out_idx[i] = NN of queries[i] */
void orc_nn_search_batch(const double *dst, size_t m, int dim, const double *queries, size_t q,
                         int mode, int32_t *out_idx);

/* ---- ICP drivers (lib.rs:105-130 Icp2d::estimate, lib.rs:148-173 Icp3d::estimate) ---- */
enum { ORC_NN_BRUTE = 0, ORC_NN_KDTREE = 1 };
enum { ORC_CORR_NONE = 0, ORC_CORR_LAST = 1, ORC_CORR_ALL = 2 };

typedef struct {
    int nn_mode;            /* ORC_NN_* */
    int corr_mode;          /* ORC_CORR_* */
    int32_t *corr;          /* [n] (LAST) or [max_iter*n] (ALL); nullable */
    uint32_t *inner;        /* [max_iter]: accepted | reason<<16; nullable */
    uint64_t *corr_hash;    /* one value; nullable.  Definition: see orc_corr_hash_step */
    const void *shared_tree;/* nullable: an orc_kdtree built once over dst (Icp2d::new once, many estimates) */
    double *final_cost;     /* nullable: huber_error (lib.rs:45-50) of the LAST inner solve's transform over the
                               last outer iteration's (transformed source, matched target) points; 0 when no
                               outer iteration ran */
    double *margins;        /* nullable [2], caller-initialised to +inf: smallest relative distance of a
                               delta-norm / error break test (lib.rs:71,76) from its threshold */
} orc_trace;

/* status: 0 ok, 1 = empty target with non-empty source (reference panics, lib.rs:122,165) */
int orc_icp2d_estimate(const double *src, size_t n, const double *dst, size_t m,
                       const double init[6], uint32_t max_iter, double out[6], orc_trace *tr);
int orc_icp3d_estimate(const double *src, size_t n, const double *dst, size_t m,
                       const double init[6], uint32_t max_iter, double out[6], orc_trace *tr);

/* Order-independent per-iteration checksum of the correspondences, chained over iterations:
 *   h_it = sum_i mix64((i << 32) | idx_i)   (mod 2^64, mix64 = splitmix64 finaliser)
 *   H    = H * 0x100000001B3 + h_it          (H starts at 0)                             */
uint64_t orc_mix64(uint64_t x);

/* Ragged batch over pthreads (one pair per task).  off arrays have B+1 entries (points).
 * dim = 2 or 3.  out_T = B*6.  inner (nullable) = B*max_iter.  corr_hash (nullable) = B.
 * status (nullable) = B.  Returns number of threads used. */
int orc_icp_estimate_batch(int dim, const double *src, const uint32_t *src_off,
                           const double *dst, const uint32_t *dst_off, const double *init_T,
                           uint32_t B, uint32_t max_iter, int nn_mode, int nthreads,
                           double *out_T, uint32_t *inner, uint64_t *corr_hash, int32_t *status);

/* Same, plus: dst_off == NULL -> every pair is estimated against the ONE target dst[0..shared_m)
 * (kd-tree built once, as `Icp2d::new(&dst)` followed by many `estimate` calls);
 * final_cost (nullable) = B, margins (nullable) = B*2: see orc_trace. */
int orc_icp_estimate_batch_ex(int dim, const double *src, const uint32_t *src_off,
                              const double *dst, const uint32_t *dst_off, size_t shared_m,
                              const double *init_T, uint32_t B, uint32_t max_iter, int nn_mode,
                              int nthreads, double *out_T, uint32_t *inner, uint64_t *corr_hash,
                              int32_t *status, double *final_cost, double *margins);

#ifdef __cplusplus
}
#endif
#endif
