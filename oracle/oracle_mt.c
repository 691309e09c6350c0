/* TEST INFRASTRUCTURE -- CPU restatement of the reference's random-number stream and of the particle load a deck
 * builds from it.  Only tests/ (and bench.py's cpu_baseline leg) may call this; the product never does.
 *
 * What it follows:
 *   src/util/mtrand/mtrand.c:16-37,54-62   MT19937: 624-word state, twist, temper; seed_mt_rng's seeding
 *                                          (state[0] = seed ^ 0x900df00c, the Knuth recurrence for the rest)
 *   src/util/mtrand/mtrand.c:218-240       mt_drand  = drand53_o of two words (mtrand_conv.h:58)
 *   src/util/mtrand/mtrand.c:253-438       mt_drandn = 256-layer ziggurat over word PAIRS: one pair for the 1-bit
 *                                          sign, 8-bit layer and 53-bit trapezoid deviate; on rejection one more pair
 *                                          for y (two more in the tail layer), repeated until accepted
 *   src/util/mtrand/make_zig.c             the layer table itself: bisection on the tail start r in long double, with
 *                                          the inverse density taken through a DOUBLE sqrt (which is why the table is
 *                                          not the mathematically exact one: 1e-14 relative)
 *   src/vpic/vpic.hxx:491-505              seed_rand, uniform_rand = low*(1-d) + high*d, maxwellian_rand = dev*drandn
 *   src/vpic/misc.cxx:16-105               inject_particle (age = 0, update_rhob = 0)
 * Pinned: tests/test_oracle_mt.py compares the word stream, mt_drand, mt_drandn (5e6 draws, which visits the
 * rejection and tail branches) and the layer table with the reference compiled from source
 * (oracle/_ref/libvpic_ref_scalar.so), and the loaded particle arrays with what the reference's own
 * initialize() leaves after running oracle/decks/thermal_c1.cxx -- all bit for bit.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>
#include "vpic_oracle.h"

#define ORC_MT_N 624
#define ORC_MT_M 397

typedef struct orc_mt {
  uint32_t next;
  uint32_t state[ORC_MT_N];
} orc_mt_t;

static double g_zx[257], g_zy[257];
static int g_zig_ready = 0;
static const double kZigScale = 1. / 1.8446744073709551616e+19;   /* 2^-64 */

/* make_zig.c:9-62 */
static long double zig_pdf(long double x) { return expl(-0.5l * x * x); }
static long double zig_pdf_inv(long double y) { return (y <= 0 || y >= 1) ? 0 : (long double)sqrt((double)(-2.0l * logl(y))); }
static long double zig_build(long double *x, long double *y, int N, long double r) {
  const long double v = zig_pdf(r) / r + r * zig_pdf(r);
  x[N] = v / zig_pdf(r);
  y[N] = zig_pdf(x[N]);
  x[N - 1] = r;
  y[N - 1] = zig_pdf(x[N - 1]);
  for (int n = N - 2; n > 0; n--) {
    x[n] = zig_pdf_inv(y[n + 1] + v / x[n + 1]);
    y[n] = zig_pdf(x[n]);
  }
  x[0] = 0;
  y[0] = zig_pdf(x[0]);
  return v - (x[1] - x[0]) * (y[0] - y[1]);
}

static double g_zig_r;

void orc_mt_zig_table(double *x_out, double *y_out, double *r_out) {
  if (!g_zig_ready) {
    long double x[257], y[257], a = 0, b = 10, r;
    for (;;) {
      r = 0.5 * (a + b);
      if (r == a || r == b) break;
      const long double dv = zig_build(x, y, 256, r);
      if (dv == 0) break;
      if (dv > 0) a = r; else b = r;
    }
    for (int n = 0; n <= 256; n++) { g_zx[n] = (double)x[n]; g_zy[n] = (double)y[n]; }
    g_zig_r = (double)r;
    g_zig_ready = 1;
  }
  if (x_out) memcpy(x_out, g_zx, sizeof(g_zx));
  if (y_out) memcpy(y_out, g_zy, sizeof(g_zy));
  if (r_out) *r_out = g_zig_r;
}

int orc_mt_sizeof(void) { return (int)sizeof(orc_mt_t); }

/* mtrand.c:54-62 */
void orc_mt_seed(orc_mt_t *rng, unsigned int seed) {
  rng->next = ORC_MT_N;
  rng->state[0] = seed ^ 0x900df00cu;
  for (int j = 1; j < ORC_MT_N; j++) rng->state[j] = 1812433253u * (rng->state[j - 1] ^ (rng->state[j - 1] >> 30)) + (uint32_t)j;
}

static uint32_t twist(uint32_t u, uint32_t v) {
  return (((u & 0x80000000u) | (v & 0x7fffffffu)) >> 1) ^ ((0u - (v & 1u)) & 0x9908b0dfu);
}

/* mtrand.c:22-37 */
uint32_t orc_mt_u32(orc_mt_t *rng) {
  if (rng->next == ORC_MT_N) {
    uint32_t *p = rng->state;
    int j;
    rng->next = 0;
    for (j = 0; j < ORC_MT_N - ORC_MT_M; j++) p[j] = p[j + ORC_MT_M] ^ twist(p[j], p[j + 1]);
    for (; j < ORC_MT_N - 1; j++) p[j] = p[j + ORC_MT_M - ORC_MT_N] ^ twist(p[j], p[j + 1]);
    p[ORC_MT_N - 1] = p[ORC_MT_M - 1] ^ twist(p[ORC_MT_N - 1], p[0]);
  }
  uint32_t y = rng->state[rng->next++];
  y ^= y >> 11;
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= y >> 18;
  return y;
}

static double d53_o(uint32_t a, uint32_t b) { return ((a >> 5) * 67108864. + (b >> 6) + 1.5) * (1. / 9007199254740994.); }
static double d53_c(uint32_t a, uint32_t b) { return ((a >> 5) * 67108864. + ((b >> 6) + (b & 1))) * (1. / 9007199254740992.); }
static double d53_c1(uint32_t a, uint32_t b) { return ((a >> 5) * 67108864. + ((b >> 6) + 1)) * (1. / 9007199254740992.); }

/* mtrand.c:240 */
double orc_mt_drand(orc_mt_t *rng) {
  const uint32_t a = orc_mt_u32(rng), b = orc_mt_u32(rng);
  return d53_o(a, b);
}

/* mtrand.c:395-438 */
double orc_mt_drandn(orc_mt_t *rng) {
  orc_mt_zig_table(0, 0, 0);
  const double R = g_zig_r;
  uint32_t a, b, i, s;
  double x, y, j;
  for (;;) {
    a = orc_mt_u32(rng);
    b = orc_mt_u32(rng);
    s = a & 1u;
    i = (a & 0x1feu) >> 1;
    j = 4294967296. * b + ((a & 0xfffff800u) + ((a & 0x400u) << 1));
    x = j * (kZigScale * g_zx[i + 1]);
    if (x < g_zx[i]) break;
    a = orc_mt_u32(rng);
    b = orc_mt_u32(rng);
    y = d53_c(a, b);
    if (i != 255) y = g_zy[i] + (g_zy[i + 1] - g_zy[i]) * y;
    else {
      a = orc_mt_u32(rng);
      b = orc_mt_u32(rng);
      x = R - (1. / R) * log(d53_c1(a, b));
      y *= exp(-R * (x - 0.5 * R));
    }
    if (y < exp(-0.5 * x * x)) break;
  }
  return s ? -x : x;
}

void orc_mt_fill_u32(orc_mt_t *rng, uint32_t *out, long n) { for (long k = 0; k < n; k++) out[k] = orc_mt_u32(rng); }
void orc_mt_fill_drand(orc_mt_t *rng, double *out, long n) { for (long k = 0; k < n; k++) out[k] = orc_mt_drand(rng); }
void orc_mt_fill_drandn(orc_mt_t *rng, double *out, long n) { for (long k = 0; k < n; k++) out[k] = orc_mt_drandn(rng); }

/* `prog` is a string of 'U' (mt_drand) and 'N' (mt_drandn); n records of it, in stream order */
void orc_mt_draw(orc_mt_t *rng, const char *prog, long n, double *out) {
  const int len = (int)strlen(prog);
  for (long r = 0; r < n; r++)
    for (int t = 0; t < len; t++) out[r * len + t] = prog[t] == 'N' ? orc_mt_drandn(rng) : orc_mt_drand(rng);
}

/* misc.cxx:16-105 with age = 0 and update_rhob = 0.  Returns 1 when the particle was appended. */
int orc_inject_particle(vpb_particle_t *p0, int *np, int max_np, double x, double y, double z, double ux, double uy, double uz,
                        double q, int64_t tag, const vpb_grid_t *g) {
  const double x0 = (double)g->x0, y0 = (double)g->y0, z0 = (double)g->z0;
  const double x1 = (double)g->x1, y1 = (double)g->y1, z1 = (double)g->z1;
  const int nx = g->nx, ny = g->ny, nz = g->nz;
  int ix, iy, iz;
  if ((x < x0) | (x > x1) | ((x == x1) & (g->bc[VPB_BOUNDARY(1, 0, 0)] >= 0))) return 0;
  if ((y < y0) | (y > y1) | ((y == y1) & (g->bc[VPB_BOUNDARY(0, 1, 0)] >= 0))) return 0;
  if ((z < z0) | (z > z1) | ((z == z1) & (g->bc[VPB_BOUNDARY(0, 0, 1)] >= 0))) return 0;
  if (*np >= max_np) return -1;
  x = ((double)nx) * ((x - x0) / (x1 - x0)); ix = (int)x; x -= (double)ix; x = (x + x) - 1;
  if (ix == nx) x = 1;
  if (ix == nx) ix = nx - 1;
  ix++;
  y = ((double)ny) * ((y - y0) / (y1 - y0)); iy = (int)y; y -= (double)iy; y = (y + y) - 1;
  if (iy == ny) y = 1;
  if (iy == ny) iy = ny - 1;
  iy++;
  z = ((double)nz) * ((z - z0) / (z1 - z0)); iz = (int)z; z -= (double)iz; z = (z + z) - 1;
  if (iz == nz) z = 1;
  if (iz == nz) iz = nz - 1;
  iz++;
  vpb_particle_t *p = p0 + ((*np)++);
  p->dx = (float)x; p->dy = (float)y; p->dz = (float)z;
  p->i = ix + (nx + 2) * (iy + (ny + 2) * iz);
  p->ux = (float)ux; p->uy = (float)uy; p->uz = (float)uz;
  p->q = (float)q;
  p->tag = tag;
  return 1;
}

/* The load loop of oracle/decks/thermal_c1.cxx (SURVEY.md 8d, C1/C4 recipe): per iteration one position from three
 * uniform_rand(lo, hi), then an electron and a co-located ion with three maxwellian_rand(vth) each.  Returns the
 * number of iterations done (stops when an array is full). */
long orc_load_thermal_pairs_tagged(orc_mt_t *rng, long n, const double lo[3], const double hi[3], double vth_e, double vth_i, double q,
                                   vpb_particle_t *pe, int *npe, int max_e, vpb_particle_t *pi, int *npi, int max_i, const vpb_grid_t *g,
                                   long tag0, long tag_step);
long orc_load_thermal_pairs(orc_mt_t *rng, long n, const double lo[3], const double hi[3], double vth_e, double vth_i, double q,
                            vpb_particle_t *pe, int *npe, int max_e, vpb_particle_t *pi, int *npi, int max_i, const vpb_grid_t *g) {
  return orc_load_thermal_pairs_tagged(rng, n, lo, hi, vth_e, vth_i, q, pe, npe, max_e, pi, npi, max_i, g, 0, 0);
}

/* the same with inject_particle's tag argument = tag0 + k*tag_step (oracle/decks/thermal_small.cxx passes the loop counter) */
long orc_load_thermal_pairs_tagged(orc_mt_t *rng, long n, const double lo[3], const double hi[3], double vth_e, double vth_i, double q,
                                   vpb_particle_t *pe, int *npe, int max_e, vpb_particle_t *pi, int *npi, int max_i, const vpb_grid_t *g,
                                   long tag0, long tag_step) {
  for (long k = 0; k < n; k++) {
    double d = orc_mt_drand(rng);
    const double x = lo[0] * (1 - d) + hi[0] * d;
    d = orc_mt_drand(rng);
    const double y = lo[1] * (1 - d) + hi[1] * d;
    d = orc_mt_drand(rng);
    const double z = lo[2] * (1 - d) + hi[2] * d;
    /* the deck writes inject_particle( sp, x, y, z, maxwellian_rand(v), maxwellian_rand(v), maxwellian_rand(v), ... ): the
     * order in which a call's arguments are evaluated is the compiler's choice, and g++ on x86-64 (what builds the
     * reference here and on the Cray the deck was written for) goes right to left -- the FIRST deviate is uz */
    const double ez = vth_e * orc_mt_drandn(rng), ey = vth_e * orc_mt_drandn(rng), ex = vth_e * orc_mt_drandn(rng);
    if (orc_inject_particle(pe, npe, max_e, x, y, z, ex, ey, ez, -q, tag0 + k * tag_step, g) < 0) return k;
    const double jz = vth_i * orc_mt_drandn(rng), jy = vth_i * orc_mt_drandn(rng), jx = vth_i * orc_mt_drandn(rng);
    if (orc_inject_particle(pi, npi, max_i, x, y, z, jx, jy, jz, q, tag0 + k * tag_step, g) < 0) return k;
  }
  return n;
}
