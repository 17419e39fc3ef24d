// nbody.hpp - header-only C++ mirror of the reference's simulation actor over the C ABI in nbody.h.
//
// `nbody::OctreeSearch` keeps the public surface of `AOctreeSearch`
// (/root/reference/Source/NBody/OctreeSearch.h:111-149): the same verbs (CreateSpacePoints, ComputeCubeSize, CreateOctree,
// Tick, CleanParticles), the same public members (Particles, Size, Initialized, ShowOctree, PhDeltaTime) and the same
// behaviour when nothing is loaded (silent no-ops, OctreeSearch.cpp:49,76). It is what a host application - or the Unreal
// adapter in INTEGRATION.md - uses instead of the CPU actor; no Unreal types, no CUDA types.
//
// Differences forced by device residency: `Particles` is a host mirror that Tick() refreshes after the step (set
// `MirrorParticles = false` to skip the read-back and call Download() when needed); bodies written into `Particles` by
// the caller are sent with Upload() (the reference's callers write the TArray in place, OctreeSearch.h:118).
// Errors never throw across the C ABI; this wrapper turns a failing status into std::runtime_error.
#ifndef NBODY_B200_HPP
#define NBODY_B200_HPP

#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "nbody.h"

namespace nbody {

// FParticle, OctreeSearch.h:9-18 (40 bytes: Mass, Position, Velocity, Acceleration).
struct FVector { float X = 0.f, Y = 0.f, Z = 0.f; };
struct FParticle {
  float Mass = 0.f;
  FVector Position, Velocity, Acceleration;
};
static_assert(sizeof(FParticle) == sizeof(nbody_particle) && sizeof(FParticle) == 40, "FParticle must stay the 40-byte record");

class OctreeSearch {
 public:
  // ---- the reference's public members (OctreeSearch.h:116-127)
  float Size = 0.f;
  std::vector<FParticle> Particles;
  bool Initialized = false;
  bool ShowOctree = false;
  float PhDeltaTime = 0.01f;      // OctreeSearch.cpp:8
  // ---- additions
  bool MirrorParticles = true;    // refresh `Particles` from the GPU after every Tick (the reference's consumers read it)

  // The reference's constructor takes nothing and ships Barnes-Hut with G = 1e4, Theta = 1.0, no softening
  // (OctreeSearch.cpp:8-12,85; OctreeSearch.h:104); pass a config to choose otherwise.
  OctreeSearch() {
    nbody_config cfg;
    nbody_config_default(&cfg);
    Open(cfg);
  }
  explicit OctreeSearch(const nbody_config& cfg) { Open(cfg); }
  ~OctreeSearch() { nbody_destroy(sim_); }
  OctreeSearch(const OctreeSearch&) = delete;
  OctreeSearch& operator=(const OctreeSearch&) = delete;

  // OctreeSearch.cpp:58-72 (default Size = 200 as in OctreeSearch.h:142). Seeded, unlike the reference's rand().
  void CreateSpacePoints(int32_t N, float SizeArg = 200.f, uint64_t seed = 1234) {
    Check(nbody_create_space_points(sim_, N, SizeArg, seed));
    Size = SizeArg;
    Initialized = true;
    Particles.resize((size_t)N);
    Download();
  }
  // OctreeSearch.cpp:47-56; silently nothing when not Initialized (cpp:49).
  void ComputeCubeSize() {
    if (Initialized) Check(nbody_compute_cube_size(sim_, &Size));
  }
  // OctreeSearch.cpp:74-89: accelerations at the current positions; silently nothing when not Initialized (cpp:76).
  void CreateOctree() {
    if (!Initialized) return;
    Check(nbody_create_octree(sim_));
    if (MirrorParticles) Download();
  }
  // OctreeSearch.cpp:21-34. DeltaTime is ignored, as in the reference: physics advances by PhDeltaTime (<= 0 pauses).
  void Tick(float /*DeltaTime*/ = 0.f) {
    Check(nbody_set_param(sim_, NBODY_PARAM_PH_DELTA_TIME, PhDeltaTime));
    Check(nbody_set_param(sim_, NBODY_PARAM_SHOW_OCTREE, ShowOctree ? 1.0 : 0.0));
    if (!Initialized || !(PhDeltaTime > 0.f)) return;
    Check(nbody_tick(sim_));
    Size = Stats().cube_size;   // ComputeCubeSize ran at the start of the step (OctreeSearch.cpp:26)
    if (MirrorParticles) Download();
  }
  // OctreeSearch.cpp:91-97.
  void CleanParticles() {
    Initialized = false;
    Check(nbody_clean_particles(sim_));
    Particles.clear();
  }
  // What DrawOctreeBoxes draws (OctreeSearch.cpp:36-45), as data: 7 floats per occupied leaf (centre, half extents, count).
  std::vector<float> OctreeBoxes() {
    std::vector<float> boxes(7 * Particles.size() + 7);
    int64_t k = 0;
    Check(nbody_octree_boxes(sim_, boxes.data(), (int64_t)Particles.size() + 1, &k));
    boxes.resize((size_t)(7 * k));
    return boxes;
  }

  // ---- host <-> device
  void Upload() {     // after the caller wrote `Particles` in place
    if (Particles.empty()) return;
    Check(nbody_set_particles_aos(sim_, Particles.data(), (int64_t)Particles.size(), sizeof(FParticle)));
    Initialized = true;
  }
  void Download() {
    if (!Particles.empty()) Check(nbody_get_particles_aos(sim_, Particles.data(), (int64_t)Particles.size(), sizeof(FParticle)));
  }
  void Step(float dt, int32_t nsteps) {   // nsteps Ticks without per-step host synchronisation or read-back
    Check(nbody_step(sim_, dt, nsteps));
    if (MirrorParticles) Download();
  }
  nbody_stats Stats() {
    nbody_stats st;
    Check(nbody_stats_get(sim_, &st));
    return st;
  }
  nbody_sim* handle() { return sim_; }

 private:
  nbody_sim* sim_ = nullptr;
  void Open(const nbody_config& cfg) {
    PhDeltaTime = cfg.ph_delta_time;
    const int rc = nbody_create(&sim_, &cfg);
    if (rc != NBODY_OK) throw std::runtime_error(std::string("nbody_create failed (") + std::to_string(rc) + "): " + nbody_last_error());
  }
  static void Check(int rc) {
    if (rc != NBODY_OK) throw std::runtime_error(std::string("nbody error ") + std::to_string(rc) + ": " + nbody_last_error());
  }
};

}  // namespace nbody
#endif  // NBODY_B200_HPP
