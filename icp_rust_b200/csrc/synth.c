/*
 * synth.c — synthetic 2-D LiDAR scan pairs for the benchmark configurations of BASELINE.json
 * (configs 2, 3, 5).  Workload tooling: not part of the ICP path and not part of the oracle.
 *
 * Scene per pair: rectangular room W,H ~ U(4,12) m with two circular
 * pillars; the sensor sweeps `beams` bearings evenly over 360 degrees in order; closed-form ray
 * casting; Gaussian range noise (default 3 mm — without noise the reference's MAD scale is 0 and
 * the solver returns identity); a fraction of beams is replaced by the invalid
 * return (0,0) exactly as the bundled scans contain them.  Units: millimetres.
 * dst is observed from pose A, src from pose A o D with D ~ U(+-50 mm, +-50 mm, +-0.02 rad);
 * true_T (optional) is D as [r00 r01 r10 r11 tx ty], the transform taking src onto dst.
 * Every pair has its own counter-based stream: pair p of seed s is the same numbers whatever
 * the batch split or thread count.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>

typedef struct { uint64_t s; } rng_t;

static uint64_t splitmix(uint64_t *s)
{
    uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static double uni(rng_t *r) { return (double)(splitmix(&r->s) >> 11) * (1.0 / 9007199254740992.0); }
static double urange(rng_t *r, double a, double b) { return a + (b - a) * uni(r); }
static double gauss(rng_t *r)
{
    double u1 = uni(r), u2 = uni(r);
    if (u1 < 1e-300) u1 = 1e-300;
    return sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
}

typedef struct { double W, H, cx[2], cy[2], cr[2]; } scene_t;

static double cast(const scene_t *sc, double px, double py, double c, double s)
{
    double t = 1e30, q;
    int k;
    if (c > 1e-12) { q = (sc->W - px) / c; if (q < t) t = q; }
    if (c < -1e-12) { q = (0.0 - px) / c; if (q < t) t = q; }
    if (s > 1e-12) { q = (sc->H - py) / s; if (q < t) t = q; }
    if (s < -1e-12) { q = (0.0 - py) / s; if (q < t) t = q; }
    for (k = 0; k < 2; k++) {
        double ox = px - sc->cx[k], oy = py - sc->cy[k];
        double b = ox * c + oy * s, cc = ox * ox + oy * oy - sc->cr[k] * sc->cr[k];
        double disc = b * b - cc;
        if (disc > 0.0) {
            q = -b - sqrt(disc);
            if (q > 1e-6 && q < t) t = q;
        }
    }
    return t;
}

static void scan(const scene_t *sc, double px, double py, double th, uint32_t beams, double noise,
                 double dropout, rng_t *r, double *out)
{
    uint32_t k;
    for (k = 0; k < beams; k++) {
        double a = 6.283185307179586 * (double)k / (double)beams;
        double c = cos(th + a), s = sin(th + a);
        double rg = cast(sc, px, py, c, s) + noise * gauss(r);
        if (uni(r) < dropout) {
            out[2 * k] = 0.0;
            out[2 * k + 1] = 0.0;
        } else {
            out[2 * k] = rg * cos(a);
            out[2 * k + 1] = rg * sin(a);
        }
    }
}

static int inside_free(const scene_t *sc, double x, double y)
{
    int k;
    if (x < 500 || y < 500 || x > sc->W - 500 || y > sc->H - 500) return 0;
    for (k = 0; k < 2; k++) {
        double dx = x - sc->cx[k], dy = y - sc->cy[k];
        if (dx * dx + dy * dy < (sc->cr[k] + 300) * (sc->cr[k] + 300)) return 0;
    }
    return 1;
}

static void make_pair(uint64_t seed, uint64_t pair, uint32_t beams, double noise, double dropout,
                      double *src, double *dst, double *true_T)
{
    rng_t r;
    scene_t sc;
    double ax, ay, ath, dx, dy, dth, bx, by, bth, ca, sa;
    int k, tries = 0;
    {   /* independent stream per pair: hash (seed, pair) into the state */
        uint64_t h = seed ^ 0x2545F4914F6CDD1Dull, g = pair * 0xD6E8FEB86659FD93ull + 0x632BE59BD9B4E019ull;
        r.s = splitmix(&h) ^ splitmix(&g);
        r.s = splitmix(&r.s);
    }
    sc.W = urange(&r, 4000, 12000);
    sc.H = urange(&r, 4000, 12000);
    for (k = 0; k < 2; k++) {
        sc.cr[k] = urange(&r, 150, 400);
        sc.cx[k] = urange(&r, 800, sc.W - 800);
        sc.cy[k] = urange(&r, 800, sc.H - 800);
    }
    do {
        ax = urange(&r, 500, sc.W - 500);
        ay = urange(&r, 500, sc.H - 500);
    } while (!inside_free(&sc, ax, ay) && ++tries < 1000);
    ath = urange(&r, -3.141592653589793, 3.141592653589793);
    dx = urange(&r, -50, 50);
    dy = urange(&r, -50, 50);
    dth = urange(&r, -0.02, 0.02);
    /* pose B = A o D (D expressed in A's frame) */
    ca = cos(ath); sa = sin(ath);
    bx = ax + ca * dx - sa * dy;
    by = ay + sa * dx + ca * dy;
    bth = ath + dth;
    scan(&sc, ax, ay, ath, beams, noise, dropout, &r, dst);
    scan(&sc, bx, by, bth, beams, noise, dropout, &r, src);
    if (true_T) {
        true_T[0] = cos(dth); true_T[1] = -sin(dth);
        true_T[2] = sin(dth); true_T[3] = cos(dth);
        true_T[4] = dx; true_T[5] = dy;
    }
}

typedef struct {
    uint64_t seed, first; uint32_t count, beams; double noise, dropout;
    double *src, *dst, *true_T; volatile uint32_t next;
} job_t;

static void *worker(void *arg)
{
    job_t *j = (job_t *)arg;
    for (;;) {
        uint32_t p = __sync_fetch_and_add(&j->next, 1u);
        if (p >= j->count) break;
        make_pair(j->seed, j->first + p, j->beams, j->noise, j->dropout,
                  j->src + (size_t)p * j->beams * 2, j->dst + (size_t)p * j->beams * 2,
                  j->true_T ? j->true_T + (size_t)p * 6 : NULL);
    }
    return NULL;
}

int icp_synth_scan_pairs(uint64_t seed, uint64_t first_pair, uint32_t count, uint32_t beams,
                         double noise_mm, double dropout, double *src, double *dst, double *true_T,
                         int nthreads)
{
    job_t j;
    pthread_t th[256];
    int i, started = 0;
    j.seed = seed; j.first = first_pair; j.count = count; j.beams = beams; j.noise = noise_mm;
    j.dropout = dropout; j.src = src; j.dst = dst; j.true_T = true_T; j.next = 0;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    if (nthreads == 1) { worker(&j); return 1; }
    for (i = 0; i < nthreads; i++)
        if (pthread_create(&th[started], NULL, worker, &j) == 0) started++;
    if (!started) worker(&j);
    for (i = 0; i < started; i++) pthread_join(th[i], NULL);
    return started ? started : 1;
}

/* config 3: `count` scans observed from random poses inside ONE fixed scene (seeded by `seed`),
 * expressed in the WORLD frame perturbed by a small pose error, plus the scene's map points. */
int icp_synth_map_scene(uint64_t seed, uint32_t map_points, double *map_xy)
{
    /* map = dense noise-free-ish sampling of the scene from a trajectory of poses */
    rng_t r;
    scene_t sc;
    uint32_t got = 0, beams = 1080, k;
    int q;
    double *tmp;
    r.s = seed * 0xD1342543DE82EF95ull + 0x12345;
    sc.W = 40000; sc.H = 40000;
    for (q = 0; q < 2; q++) {
        sc.cr[q] = urange(&r, 1500, 4000);
        sc.cx[q] = urange(&r, 8000, 32000);
        sc.cy[q] = urange(&r, 8000, 32000);
    }
    tmp = (double *)malloc((size_t)beams * 2 * sizeof(double));
    while (got < map_points) {
        double px, py, th, c, s;
        do { px = urange(&r, 500, sc.W - 500); py = urange(&r, 500, sc.H - 500); } while (!inside_free(&sc, px, py));
        th = urange(&r, -3.141592653589793, 3.141592653589793);
        scan(&sc, px, py, th, beams, 3.0, 0.0, &r, tmp);
        c = cos(th); s = sin(th);
        for (k = 0; k < beams && got < map_points; k++) {
            map_xy[2 * got] = px + c * tmp[2 * k] - s * tmp[2 * k + 1];
            map_xy[2 * got + 1] = py + s * tmp[2 * k] + c * tmp[2 * k + 1];
            got++;
        }
    }
    free(tmp);
    return 0;
}

typedef struct { uint64_t seed; uint32_t count, beams; double *out, *init_T; volatile uint32_t next; } mjob_t;

static void *mworker(void *arg)
{
    mjob_t *j = (mjob_t *)arg;
    for (;;) {
        uint32_t p = __sync_fetch_and_add(&j->next, 1u), k;
        rng_t r, rs;
        scene_t sc;
        int q;
        double px, py, th, ex, ey, eth, c, s;
        double *o;
        if (p >= j->count) break;
        rs.s = j->seed * 0xD1342543DE82EF95ull + 0x12345;
        sc.W = 40000; sc.H = 40000;
        for (q = 0; q < 2; q++) {
            sc.cr[q] = urange(&rs, 1500, 4000);
            sc.cx[q] = urange(&rs, 8000, 32000);
            sc.cy[q] = urange(&rs, 8000, 32000);
        }
        {
            uint64_t h = j->seed ^ 0xA0761D6478BD642Full, g = (uint64_t)p * 0xD6E8FEB86659FD93ull + 77;
            r.s = splitmix(&h) ^ splitmix(&g);
            r.s = splitmix(&r.s);
        }
        do { px = urange(&r, 500, sc.W - 500); py = urange(&r, 500, sc.H - 500); } while (!inside_free(&sc, px, py));
        th = urange(&r, -3.141592653589793, 3.141592653589793);
        o = j->out + (size_t)p * j->beams * 2;
        scan(&sc, px, py, th, j->beams, 3.0, 0.0, &r, o);   /* sensor frame */
        /* initial transform = true pose composed with a small error */
        ex = urange(&r, -50, 50); ey = urange(&r, -50, 50); eth = urange(&r, -0.02, 0.02);
        c = cos(th + eth); s = sin(th + eth);
        j->init_T[(size_t)p * 6 + 0] = c; j->init_T[(size_t)p * 6 + 1] = -s;
        j->init_T[(size_t)p * 6 + 2] = s; j->init_T[(size_t)p * 6 + 3] = c;
        j->init_T[(size_t)p * 6 + 4] = px + ex; j->init_T[(size_t)p * 6 + 5] = py + ey;
        (void)k;
    }
    return NULL;
}

int icp_synth_map_scans(uint64_t seed, uint32_t count, uint32_t beams, double *scans_xy, double *init_T,
                        int nthreads)
{
    mjob_t j;
    pthread_t th[256];
    int i, started = 0;
    j.seed = seed; j.count = count; j.beams = beams; j.out = scans_xy; j.init_T = init_T; j.next = 0;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    if (nthreads == 1) { mworker(&j); return 1; }
    for (i = 0; i < nthreads; i++)
        if (pthread_create(&th[started], NULL, mworker, &j) == 0) started++;
    if (!started) mworker(&j);
    for (i = 0; i < started; i++) pthread_join(th[i], NULL);
    return started ? started : 1;
}
