// TEST INFRASTRUCTURE ONLY (oracle/). Not part of the product path.
//
// C wrapper around the UNMODIFIED reference simulation actor. The reference sources are compiled where
// they lie (/root/reference/Source/NBody/OctreeSearch.{h,cpp}, NBody.h) against oracle/shim; nothing is
// copied. Output goes to oracle/_ref/liboracle_ref.so (see oracle/Makefile). Every entry point is a thin
// forward to a public member of AOctreeSearch (OctreeSearch.h:111-149).
//
// The same wrapper also drives the GPU adapter actor (integration/ue4/OctreeSearch.{h,cpp}: AOctreeSearch forwarding to
// libnbody_b200.so) when compiled with -DNBODY_B200_ADAPTER -I integration/ue4, so a test can put the CPU actor and the
// GPU actor side by side through identical calls.
#include "NBody.h"
#include "OctreeSearch.h"
#include <chrono>
#include <cstring>
#ifdef _OPENMP
#include <omp.h>
#endif

const FVector FVector::ZeroVector(0.f, 0.f, 0.f);
const FColor FColor::Red = {255, 0, 0, 255};
const FColor FColor::Black = {0, 0, 0, 255};
const FColor FColor::White = {255, 255, 255, 255};

static_assert(sizeof(FParticle) == 40, "FParticle must be the 40-byte AoS record (OctreeSearch.h:9-18)");

extern "C" {

void* ref_create() { return new AOctreeSearch(); }
void ref_destroy(void* h) {
  AOctreeSearch* s = (AOctreeSearch*)h;
#ifndef NBODY_B200_ADAPTER
  if (s->ParticleOctree) s->CleanParticles();
#endif
  delete s;
}
int ref_sizeof_particle() { return (int)sizeof(FParticle); }

// Inject initial conditions into the public Particles array (OctreeSearch.h:118) - 40-byte AoS records.
void ref_set_particles(void* h, const void* aos, int n) {
  AOctreeSearch* s = (AOctreeSearch*)h;
  s->Particles.SetNum(n);
  if (n > 0) memcpy(&s->Particles[0], aos, (size_t)n * sizeof(FParticle));
#ifdef NBODY_B200_ADAPTER
  s->PushParticles();   // the device copy follows the host array
#else
  s->Initialized = true;
#endif
}
int ref_num(void* h) { return ((AOctreeSearch*)h)->Particles.Num(); }
void ref_get_particles(void* h, void* aos) {
  AOctreeSearch* s = (AOctreeSearch*)h;
  if (s->Particles.Num() > 0) memcpy(aos, &s->Particles[0], (size_t)s->Particles.Num() * sizeof(FParticle));
}
void ref_set_dt(void* h, float dt) { ((AOctreeSearch*)h)->PhDeltaTime = dt; }
float ref_get_dt(void* h) { return ((AOctreeSearch*)h)->PhDeltaTime; }
void ref_set_show_octree(void* h, int on) { ((AOctreeSearch*)h)->ShowOctree = on != 0; }
float ref_get_size(void* h) { return ((AOctreeSearch*)h)->Size; }

// OctreeSearch.cpp:58-72 with the C library RNG seeded first (the reference never seeds it).
void ref_create_space_points(void* h, int n, float size, unsigned seed) {
  srand(seed);
  ((AOctreeSearch*)h)->CreateSpacePoints(n, size);
}
void ref_clean_particles(void* h) { ((AOctreeSearch*)h)->CleanParticles(); }
void ref_compute_cube_size(void* h) { ((AOctreeSearch*)h)->ComputeCubeSize(); }
// OctreeSearch.cpp:74-89: build + monopole + Theta=1.0 walk, exactly as shipped.
void ref_create_octree(void* h) { ((AOctreeSearch*)h)->CreateOctree(); }
// OctreeSearch.cpp:21-34, nsteps times. Returns wall seconds.
double ref_tick(void* h, int nsteps) {
  AOctreeSearch* s = (AOctreeSearch*)h;
  auto t0 = std::chrono::steady_clock::now();
  for (int k = 0; k < nsteps; k++) s->Tick(0.016f);
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

#ifdef NBODY_B200_ADAPTER
// Adapter-only knobs: the constants the reference bakes in.
void ref_adapter_config(void* h, float theta, float eps, int direct, int parity) {
  AOctreeSearch* s = (AOctreeSearch*)h;
  s->Theta = theta; s->Softening = eps; s->bDirectSum = direct != 0; s->bReferenceParity = parity != 0;
}
int ref_is_adapter() { return 1; }
#else
int ref_is_adapter() { return 0; }
// Force walk with a caller-chosen Theta on the CURRENT tree (ParticleOctree is public, OctreeSearch.h:119,
// and Theta is a parameter of Octree::ComputeForces, OctreeSearch.h:99). Theta = 0 is the reference's only
// "direct sum". Targets [i0, i1). Calls for different targets are independent (each writes only its own
// particle's Acceleration and reads the tree), so they may run on several host threads; nthreads <= 1
// keeps the reference's single-threaded order. Returns wall seconds.
double ref_compute_forces(void* h, float theta, int i0, int i1, int nthreads) {
  AOctreeSearch* s = (AOctreeSearch*)h;
  if (!s->ParticleOctree) return -1.0;
  Octree* root = s->ParticleOctree;
  auto t0 = std::chrono::steady_clock::now();
#ifdef _OPENMP
  if (nthreads > 1) {
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
    for (int i = i0; i < i1; i++) {
      s->Particles[i].Acceleration = FVector::ZeroVector;
      root->ComputeForces(&s->Particles[i], theta);
    }
  } else
#endif
  {
    (void)nthreads;
    for (int i = i0; i < i1; i++) {
      s->Particles[i].Acceleration = FVector::ZeroVector;
      root->ComputeForces(&s->Particles[i], theta);
    }
  }
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// Kick-drift with the reference's own statement order (OctreeSearch.cpp:28-31) but without rebuilding
// the tree: used after ref_compute_forces to step with a Theta other than the shipped 1.0.
void ref_integrate(void* h) {
  AOctreeSearch* s = (AOctreeSearch*)h;
  for (int32 i = 0; i < s->Particles.Num(); i++) {
    s->Particles[i].Velocity += s->PhDeltaTime * s->Particles[i].Acceleration;
    s->Particles[i].Position += s->PhDeltaTime * s->Particles[i].Velocity;
  }
}

// Root cell of the current tree (OctreeSearch.cpp:77-79): origin = previous tree's COM, half-width = Size.
int ref_root(void* h, float* origin3, float* half, float* mass, float* com3) {
  AOctreeSearch* s = (AOctreeSearch*)h;
  if (!s->ParticleOctree) return 0;
  FVector o = s->ParticleOctree->GetOrigin(), c = s->ParticleOctree->GetCenterOfMass();
  origin3[0] = o.X; origin3[1] = o.Y; origin3[2] = o.Z;
  com3[0] = c.X; com3[1] = c.Y; com3[2] = c.Z;
  *half = s->ParticleOctree->GetSize();
  *mass = s->ParticleOctree->GetTotalMass();
  return 1;
}

#endif

// What DrawOctreeBoxes emitted on the last Tick (OctreeSearch.cpp:36-45): kind 0 = box (centre, half
// extent), kind 1 = point. Returns the number of records; fills up to cap records of 7 floats.
int ref_debug_draws(void* h, float* out7, int cap) {
  AOctreeSearch* s = (AOctreeSearch*)h;
  int n = (int)s->World.draws.size();
  for (int i = 0; i < n && i < cap; i++) {
    const FDebugDraw& d = s->World.draws[i];
    float* o = out7 + 7 * (size_t)i;
    o[0] = (float)d.kind; o[1] = d.a.X; o[2] = d.a.Y; o[3] = d.a.Z; o[4] = d.b.X; o[5] = d.b.Y; o[6] = d.b.Z;
  }
  return n;
}

// Full Ticks (OctreeSearch.cpp:21-34) are single threaded by construction - that is how the actor runs inside UE.
int ref_hw_threads() {
#ifdef _OPENMP
  return omp_get_num_procs();
#else
  return 1;
#endif
}

int ref_max_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

}  // extern "C"
