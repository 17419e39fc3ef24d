/* TEST INFRASTRUCTURE ONLY (oracle/). Not part of the product path: only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library.
 *
 * Plain-C restatement of the hot path of Milias/ParallelNbody (force evaluation + integration), each
 * function citing the reference lines it follows. Parity status: PINNED - tests/test_oracle.py checks this
 * file against the unmodified reference compiled from /root/reference (oracle/_ref/liboracle_ref.so) and
 * against the golden vectors under tests/golden/ that were generated from that build
 * (tests/golden/make_golden.py). The reference itself ships no tests or golden vectors.
 *
 * Body layout at this boundary: SoA-of-float4, posm[i] = (x, y, z, mass), vel[i] = (vx, vy, vz, 0),
 * acc[i] = (ax, ay, az, 0) - the same layout the CUDA path keeps in HBM.
 *
 * Build: see oracle/Makefile (-O2 -ffp-contract=off, no -ffast-math, so fp32 results are those of the
 * statement order written here).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

int oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ---------------------------------------------------------------------------------------------------
 * Direct sum, fp64 accumulate: the accuracy yardstick for the "<= 1e-5 relative L2" criterion.
 * Force law of OctreeSearch.h:101-104 with d^2 -> d^2 + eps^2 and the d == 0 skip of OctreeSearch.h:102;
 * every source is a one-body leaf, i.e. the Theta = 0 limit of the walk. Targets [i0, i1), all n sources.
 * acc3 is double[3 * (i1 - i0)].
 * ------------------------------------------------------------------------------------------------- */
double oracle_direct_f64(int n, const float* posm, double G, double eps, int i0, int i1, double* acc3,
                         int nthreads) {
  const double eps2 = eps * eps;
  double t0 = now_s();
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : 1)
  for (int i = i0; i < i1; i++) {
    const double xi = posm[4 * (size_t)i], yi = posm[4 * (size_t)i + 1], zi = posm[4 * (size_t)i + 2];
    double ax = 0, ay = 0, az = 0;
    for (int j = 0; j < n; j++) {
      const double dx = (double)posm[4 * (size_t)j] - xi, dy = (double)posm[4 * (size_t)j + 1] - yi,
                   dz = (double)posm[4 * (size_t)j + 2] - zi;
      const double r2 = dx * dx + dy * dy + dz * dz;
      if (r2 == 0.0) continue; /* OctreeSearch.h:102 */
      const double d2 = r2 + eps2;
      const double s = G * (double)posm[4 * (size_t)j + 3] / (d2 * sqrt(d2));
      ax += s * dx; ay += s * dy; az += s * dz;
    }
    acc3[3 * (size_t)(i - i0)] = ax; acc3[3 * (size_t)(i - i0) + 1] = ay; acc3[3 * (size_t)(i - i0) + 2] = az;
  }
  return now_s() - t0;
}

/* Direct sum in the reference's own arithmetic (OctreeSearch.h:101-104): d by fp32 sqrtf, the scalar
 * G*M/pow(d,3) in double then rounded to fp32, fp32 vector multiply and fp32 accumulation, sources in
 * index order. acc4 is float[4 * (i1 - i0)] (w = 0). */
double oracle_direct_f32(int n, const float* posm, double G, float eps, int i0, int i1, float* acc4,
                         int nthreads) {
  const float eps2 = eps * eps;
  double t0 = now_s();
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : 1)
  for (int i = i0; i < i1; i++) {
    const float xi = posm[4 * (size_t)i], yi = posm[4 * (size_t)i + 1], zi = posm[4 * (size_t)i + 2];
    float ax = 0, ay = 0, az = 0;
    for (int j = 0; j < n; j++) {
      const float dx = posm[4 * (size_t)j] - xi, dy = posm[4 * (size_t)j + 1] - yi,
                  dz = posm[4 * (size_t)j + 2] - zi;
      const float r2 = dx * dx + dy * dy + dz * dz;
      if (r2 == 0.f) continue;
      const float d = sqrtf(r2 + eps2);
      const float s = (float)(G * (double)posm[4 * (size_t)j + 3] / pow((double)d, 3.0));
      ax += s * dx; ay += s * dy; az += s * dz;
    }
    float* o = acc4 + 4 * (size_t)(i - i0);
    o[0] = ax; o[1] = ay; o[2] = az; o[3] = 0.f;
  }
  return now_s() - t0;
}

/* Kick-drift integrator, OctreeSearch.cpp:28-31: v += dt*a ; x += dt*v (new v), fp32, product then add. */
void oracle_kick_drift(int n, float* posm, float* vel, const float* acc, float dt) {
  for (int i = 0; i < n; i++) {
    for (int k = 0; k < 3; k++) {
      float t = dt * acc[4 * (size_t)i + k];
      vel[4 * (size_t)i + k] += t;
      float u = dt * vel[4 * (size_t)i + k];
      posm[4 * (size_t)i + k] += u;
    }
  }
}

/* Root-cube half-width, OctreeSearch.cpp:47-56: max_i max(|x|,|y|,|z|) about the world origin. */
float oracle_cube_size(int n, const float* posm) {
  float size = 0.f;
  for (int i = 0; i < n; i++) {
    const float* p = posm + 4 * (size_t)i;
    float t = fmaxf(fmaxf(fabsf(p[0]), fabsf(p[1])), fabsf(p[2]));
    if (i == 0 || t > size) size = t;
  }
  return size;
}

/* Total energy in fp64: KE = 1/2 sum m v^2 ; PE = -G sum_{i<j} m_i m_j / sqrt(r^2 + eps^2).
 * (Not in the reference; the potential that the force law of OctreeSearch.h:104 derives from.) */
void oracle_energy(int n, const float* posm, const float* vel, double G, double eps, double* ke, double* pe,
                   int nthreads) {
  double k = 0, p = 0;
  const double eps2 = eps * eps;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : k, p) num_threads(nthreads > 0 ? nthreads : 1)
  for (int i = 0; i < n; i++) {
    const float* a = posm + 4 * (size_t)i;
    const float* v = vel + 4 * (size_t)i;
    k += 0.5 * (double)a[3] * ((double)v[0] * v[0] + (double)v[1] * v[1] + (double)v[2] * v[2]);
    double pi = 0;
    for (int j = i + 1; j < n; j++) {
      const float* b = posm + 4 * (size_t)j;
      const double dx = (double)b[0] - a[0], dy = (double)b[1] - a[1], dz = (double)b[2] - a[2];
      pi += (double)b[3] / sqrt(dx * dx + dy * dy + dz * dz + eps2);
    }
    p -= G * (double)a[3] * pi;
  }
  *ke = k; *pe = p;
}

/* ---------------------------------------------------------------------------------------------------
 * Barnes-Hut octree, restating class Octree (OctreeSearch.h:21-109) with an index-linked node pool
 * instead of new/delete. Same statement order, same fp32/double mix, so results are bit-identical to the
 * reference build (checked in tests/test_oracle.py).
 * ------------------------------------------------------------------------------------------------- */
typedef struct {
  int particle;  /* -1 = none (Octree::Particle == NULL) */
  int child;     /* index of Children[0]; the 8 children are contiguous; -1 = leaf (h:58) */
  float ox, oy, oz, size; /* Origin, Size (HALF-width, h:70-74) */
  float mass, cx, cy, cz; /* TotalMass, CenterOfMass */
} onode;

typedef struct {
  onode* nodes;
  size_t count, cap;
  const float* posm;
  double G;
  float eps2;
} otree;

static int otree_new(otree* t, float ox, float oy, float oz, float size) {
  if (t->count == t->cap) {
    t->cap = t->cap ? 2 * t->cap : 1024;
    t->nodes = (onode*)realloc(t->nodes, t->cap * sizeof(onode));
  }
  onode* nd = &t->nodes[t->count];
  nd->particle = -1; nd->child = -1;
  nd->ox = ox; nd->oy = oy; nd->oz = oz; nd->size = size;
  nd->mass = 0.f; nd->cx = nd->cy = nd->cz = 0.f;
  return (int)(t->count++);
}

/* Octree::GetOctant, h:50-56: X is the most significant bit. */
static int otree_octant(const onode* nd, const float* p) {
  int o = 0;
  if (p[0] >= nd->ox) o |= 4;
  if (p[1] >= nd->oy) o |= 2;
  if (p[2] >= nd->oz) o |= 1;
  return o;
}

/* Octree::Add, h:60-81. Depth is capped (the reference recurses forever on coincident bodies); returns
 * -1 when the cap is hit so the caller can report it. */
static int otree_add(otree* t, int node, int p, int depth) {
  if (depth > 200) return -1;
  if (t->nodes[node].child < 0) {
    if (t->nodes[node].particle < 0) { t->nodes[node].particle = p; return 0; }
    int old = t->nodes[node].particle;
    t->nodes[node].particle = -1;
    int first = -1;
    for (int i = 0; i < 8; i++) {
      onode* nd = &t->nodes[node];
      float cx = nd->ox, cy = nd->oy, cz = nd->oz;
      cx = (float)((double)cx + (double)nd->size * ((i & 4) ? 0.5 : -0.5)); /* h:71 */
      cy = (float)((double)cy + (double)nd->size * ((i & 2) ? 0.5 : -0.5)); /* h:72 */
      cz = (float)((double)cz + (double)nd->size * ((i & 1) ? 0.5 : -0.5)); /* h:73 */
      float half = (float)(0.5 * (double)nd->size);                          /* h:74 */
      int c = otree_new(t, cx, cy, cz, half); /* may realloc: nd is re-read each iteration */
      if (i == 0) first = c;
    }
    t->nodes[node].child = first;
    int r = otree_add(t, first + otree_octant(&t->nodes[node], t->posm + 4 * (size_t)old), old, depth + 1);
    if (r) return r;
    return otree_add(t, first + otree_octant(&t->nodes[node], t->posm + 4 * (size_t)p), p, depth + 1);
  }
  return otree_add(t, t->nodes[node].child + otree_octant(&t->nodes[node], t->posm + 4 * (size_t)p), p, depth + 1);
}

/* Octree::ComputeMass, h:83-97. */
static void otree_mass(otree* t, int node) {
  onode* nd = &t->nodes[node];
  if (nd->child < 0) {
    if (nd->particle >= 0) {
      const float* p = t->posm + 4 * (size_t)nd->particle;
      nd->cx = p[0]; nd->cy = p[1]; nd->cz = p[2];
      nd->mass = p[3];
    }
  } else {
    for (int i = 0; i < 8; i++) {
      otree_mass(t, nd->child + i);
      const onode* c = &t->nodes[nd->child + i];
      nd->mass += c->mass;
      nd->cx += c->cx * c->mass; nd->cy += c->cy * c->mass; nd->cz += c->cz * c->mass;
    }
    if (nd->mass) {
      const float rv = 1.f / nd->mass; /* UE FVector::operator/= multiplies by the fp32 reciprocal */
      nd->cx *= rv; nd->cy *= rv; nd->cz *= rv;
    } else { nd->cx = nd->ox; nd->cy = nd->oy; nd->cz = nd->oz; }
  }
}

/* Octree::ComputeForces, h:99-108 (eps2 = 0 reproduces it exactly; eps2 > 0 softens d). */
static void otree_force(const otree* t, int node, const float* p, float theta, float* a, long long* visits) {
  const onode* nd = &t->nodes[node];
  if (nd->child < 0 && nd->particle < 0) return;                       /* h:100 */
  const float dx = p[0] - nd->cx, dy = p[1] - nd->cy, dz = p[2] - nd->cz;
  float d = sqrtf(dx * dx + dy * dy + dz * dz);                        /* h:101 FVector::Dist */
  if (d == 0) return;                                                  /* h:102 */
  if (nd->size / d < theta || nd->particle >= 0) {                     /* h:103 */
    if (t->eps2 > 0.f) d = sqrtf(dx * dx + dy * dy + dz * dz + t->eps2);
    const float s = (float)(t->G * (double)nd->mass / pow((double)d, 3.0)); /* h:104 */
    a[0] += (nd->cx - p[0]) * s; a[1] += (nd->cy - p[1]) * s; a[2] += (nd->cz - p[2]) * s;
    if (visits) (*visits)++;
  } else if (nd->child >= 0) {
    for (int i = 0; i < 8; i++) otree_force(t, nd->child + i, p, theta, a, visits); /* h:105-107 */
  }
}

void* oracle_bh_build(int n, const float* posm, const float* origin3, float half, double G, float eps, int* status) {
  otree* t = (otree*)calloc(1, sizeof(otree));
  t->posm = posm; t->G = G; t->eps2 = eps * eps;
  otree_new(t, origin3[0], origin3[1], origin3[2], half); /* OctreeSearch.cpp:79 */
  int st = 0;
  for (int i = 0; i < n && !st; i++) st = otree_add(t, 0, i, 0); /* OctreeSearch.cpp:80 */
  if (!st) otree_mass(t, 0);                                     /* OctreeSearch.cpp:81 */
  if (status) *status = st;
  return t;
}
void oracle_bh_free(void* h) {
  otree* t = (otree*)h;
  free(t->nodes);
  free(t);
}
long long oracle_bh_num_nodes(void* h) { return (long long)((otree*)h)->count; }
void oracle_bh_root(void* h, float* mass, float* com3) {
  otree* t = (otree*)h;
  *mass = t->nodes[0].mass;
  com3[0] = t->nodes[0].cx; com3[1] = t->nodes[0].cy; com3[2] = t->nodes[0].cz;
}
/* Walk for targets [i0, i1) (OctreeSearch.cpp:83-86 with a free Theta). acc4 = float[4*(i1-i0)].
 * Returns the number of accepted node interactions. */
long long oracle_bh_forces(void* h, float theta, int i0, int i1, float* acc4, int nthreads) {
  otree* t = (otree*)h;
  long long total = 0;
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : total) num_threads(nthreads > 0 ? nthreads : 1)
  for (int i = i0; i < i1; i++) {
    float a[3] = {0.f, 0.f, 0.f};
    long long v = 0;
    otree_force(t, 0, t->posm + 4 * (size_t)i, theta, a, &v);
    float* o = acc4 + 4 * (size_t)(i - i0);
    o[0] = a[0]; o[1] = a[1]; o[2] = a[2]; o[3] = 0.f;
    total += v;
  }
  return total;
}
/* Occupied-leaf boxes in DFS order, as DrawOctreeBoxes visits them (OctreeSearch.cpp:36-45):
 * out8 = (cx, cy, cz, half, px, py, pz, particle index). Returns the count. */
static long long otree_leaves(const otree* t, int node, float* out8, long long cap, long long k) {
  const onode* nd = &t->nodes[node];
  if (nd->child < 0) {
    if (nd->particle >= 0) {
      if (k < cap) {
        float* o = out8 + 8 * (size_t)k;
        const float* p = t->posm + 4 * (size_t)nd->particle;
        o[0] = nd->ox; o[1] = nd->oy; o[2] = nd->oz; o[3] = nd->size; o[4] = p[0]; o[5] = p[1]; o[6] = p[2];
        o[7] = (float)nd->particle;
      }
      k++;
    }
    return k;
  }
  for (int i = 0; i < 8; i++) k = otree_leaves(t, nd->child + i, out8, cap, k);
  return k;
}
long long oracle_bh_leaf_boxes(void* h, float* out8, long long cap) { return otree_leaves((otree*)h, 0, out8, cap, 0); }

/* One full reference step, AOctreeSearch::Tick (OctreeSearch.cpp:21-34): cube size -> build at
 * (origin = previous root COM) -> monopoles -> walk(theta) -> kick-drift. prev_com3 is updated in place
 * (pass zeros on the first step, OctreeSearch.cpp:77). method: 1 = tree walk, 0 = direct sum
 * (oracle_direct_f32). Returns 0, or -1 if the tree hit the depth cap. */
int oracle_tick(int n, float* posm, float* vel, float* acc, float dt, float theta, double G, float eps,
                int method, float* prev_com3, int nthreads) {
  if (dt <= 0.f) return 0; /* OctreeSearch.cpp:25: PhDeltaTime <= 0 pauses */
  if (method == 1) {
    float size = oracle_cube_size(n, posm);
    int st = 0;
    void* t = oracle_bh_build(n, posm, prev_com3, size, G, eps, &st);
    if (st) { oracle_bh_free(t); return -1; }
    float m;
    oracle_bh_root(t, &m, prev_com3);
    oracle_bh_forces(t, theta, 0, n, acc, nthreads);
    oracle_bh_free(t);
  } else {
    oracle_direct_f32(n, posm, G, eps, 0, n, acc, nthreads);
  }
  oracle_kick_drift(n, posm, vel, acc, dt);
  return 0;
}
