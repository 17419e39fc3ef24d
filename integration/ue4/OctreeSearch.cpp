// AOctreeSearch on the B200: every verb of the reference actor (/root/reference/Source/NBody/OctreeSearch.cpp) forwarded to
// the C ABI of libnbody_b200.so. No physics happens in this file.
#include "NBody.h"
#include "OctreeSearch.h"

#include <cstdio>

#include "nbody.h"

static_assert(sizeof(FParticle) == sizeof(nbody_particle), "FParticle must stay the 40-byte record of the C ABI");

namespace {
void Report(const char* What) { std::fprintf(stderr, "AOctreeSearch: %s failed: %s\n", What, nbody_last_error()); }
}  // namespace

AOctreeSearch::AOctreeSearch()
    : Size(0), ParticleOctree(NULL), Initialized(false), ShowOctree(false), PhDeltaTime(0.01f), Theta(1.0f), GravityG(1e4f),
      Softening(0.f), bDirectSum(false), bReferenceParity(false), SimMethod(-1), SimParity(false) {
  PrimaryActorTick.bCanEverTick = true;   // as OctreeSearch.cpp:11
}

AOctreeSearch::~AOctreeSearch() {
  nbody_destroy(ParticleOctree);
  ParticleOctree = NULL;
}

void AOctreeSearch::BeginPlay() { Super::BeginPlay(); }

// The handle is created on first use; method and parity mode are fixed per handle, so changing them re-creates it.
bool AOctreeSearch::EnsureSim() {
  const int32 Method = bDirectSum ? NBODY_DIRECT : NBODY_BARNES_HUT;
  if (ParticleOctree && (Method != SimMethod || bReferenceParity != SimParity)) {
    nbody_destroy(ParticleOctree);
    ParticleOctree = NULL;
  }
  if (!ParticleOctree) {
    nbody_config Cfg;
    nbody_config_default(&Cfg);
    Cfg.method = Method;
    if (bReferenceParity) { Cfg.leaf_size = 1; Cfg.reference_root = 1; Cfg.mac = 1; }
    if (nbody_create(&ParticleOctree, &Cfg) != NBODY_OK) { Report("nbody_create"); ParticleOctree = NULL; return false; }
    SimMethod = Method;
    SimParity = bReferenceParity;
  }
  ApplyParams();
  return true;
}

void AOctreeSearch::ApplyParams() {
  nbody_set_param(ParticleOctree, NBODY_PARAM_PH_DELTA_TIME, PhDeltaTime);   // Blueprints write it (pause = 0)
  nbody_set_param(ParticleOctree, NBODY_PARAM_THETA, Theta);
  nbody_set_param(ParticleOctree, NBODY_PARAM_G, GravityG);
  nbody_set_param(ParticleOctree, NBODY_PARAM_EPS, Softening);
  nbody_set_param(ParticleOctree, NBODY_PARAM_SHOW_OCTREE, ShowOctree ? 1.0 : 0.0);
}

void AOctreeSearch::PushParticles() {
  if (Particles.Num() < 1 || !EnsureSim()) return;
  if (nbody_set_particles_aos(ParticleOctree, &Particles[0], Particles.Num(), sizeof(FParticle)) != NBODY_OK) { Report("nbody_set_particles_aos"); return; }
  Initialized = true;
}

void AOctreeSearch::PullParticles() {
  if (!ParticleOctree || Particles.Num() < 1) return;
  if (nbody_get_particles_aos(ParticleOctree, &Particles[0], Particles.Num(), sizeof(FParticle)) != NBODY_OK) Report("nbody_get_particles_aos");
}

// Frame driver (reference: OctreeSearch.cpp:21-34). The frame time is ignored there too; physics advances by PhDeltaTime.
void AOctreeSearch::Tick(float DeltaSeconds) {
  Super::Tick(DeltaSeconds);
  FlushPersistentDebugLines(GetWorld());
  if (ParticleOctree && Initialized) {
    ApplyParams();
    if (nbody_tick(ParticleOctree) != NBODY_OK) Report("nbody_tick");   // cube size, tree, forces, kick-drift: all on the GPU
    if (PhDeltaTime > 0) {
      PullParticles();
      nbody_stats St;
      if (nbody_stats_get(ParticleOctree, &St) == NBODY_OK) Size = St.cube_size;
    }
  }
  DrawOctreeBoxes(ParticleOctree);
}

// One point per body and - when ShowOctree - one box per occupied leaf cell, as OctreeSearch.cpp:36-45 draws them.
void AOctreeSearch::DrawOctreeBoxes(nbody_sim* Sim) {
  if (Sim == NULL || !Initialized || Particles.Num() < 1) return;
  if (ShowOctree && !bDirectSum) {
    TArray<float> Boxes;
    Boxes.SetNum(7 * Particles.Num());
    int64_t Count = 0;
    if (nbody_octree_boxes(Sim, &Boxes[0], Particles.Num(), &Count) == NBODY_OK) {
      for (int64_t k = 0; k < Count && k < Particles.Num(); k++) {
        const float* B = &Boxes[(int32)(7 * k)];
        DrawDebugBox(GetWorld(), FVector(B[0], B[1], B[2]), FVector(B[3], B[4], B[5]), FColor::Red, true);
      }
    }
  }
  for (int32 i = 0; i < Particles.Num(); i++) DrawDebugPoint(GetWorld(), Particles[i].Position, 10.0, FColor::Black, true);
}

void AOctreeSearch::ComputeCubeSize() {
  if (!Initialized || !ParticleOctree) return;
  if (nbody_compute_cube_size(ParticleOctree, &Size) != NBODY_OK) Report("nbody_compute_cube_size");
}

// Reference: OctreeSearch.cpp:58-72. The bodies are drawn on the device (seeded counter RNG; the reference uses the
// engine's unseeded global generator), then mirrored to the host array.
void AOctreeSearch::CreateSpacePoints(int32 N, float SizeArg) {
  if (N < 1 || !EnsureSim()) return;
  Size = SizeArg;
  if (nbody_create_space_points(ParticleOctree, N, SizeArg, (uint64_t)FMath::RandHelper(0x7fffffff)) != NBODY_OK) { Report("nbody_create_space_points"); return; }
  Particles.SetNum(N);
  PullParticles();
  Initialized = true;
}

// Reference: OctreeSearch.cpp:74-89 - rebuild the tree and evaluate every body's acceleration at the current positions.
void AOctreeSearch::CreateOctree() {
  if (!Initialized || !ParticleOctree) return;
  ApplyParams();
  if (nbody_create_octree(ParticleOctree) != NBODY_OK) { Report("nbody_create_octree"); return; }
  PullParticles();
}

// Reference: OctreeSearch.cpp:91-97. The handle (device buffers, stream) is kept for the next CreateSpacePoints.
void AOctreeSearch::CleanParticles() {
  Initialized = false;
  if (ParticleOctree) nbody_clean_particles(ParticleOctree);
  Particles.Empty();
}
