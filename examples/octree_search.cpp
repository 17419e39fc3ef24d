// Host program written against the reference's actor interface (AOctreeSearch, OctreeSearch.h:111-149), running on
// libnbody_b200.so through include/nbody.hpp. What BP_NBodyHUD / BP_ScreenUI do in the reference: create the actor, call
// CreateSpacePoints(2000, 1000), tick with PhDeltaTime = 0.01, pause, clean, restart.
//   g++ -std=c++14 -I include examples/octree_search.cpp -o examples/octree_search -L parallelnbody_b200 -lnbody_b200 -Wl,-rpath,$PWD/parallelnbody_b200
#include <cmath>
#include <cstdio>
#include <exception>

#include "nbody.hpp"

int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 2000;       // UI defaults: Particles 2000, Box size 1000, DeltaTime 0.01
  try {
    nbody::OctreeSearch actor;                          // Barnes-Hut, G = 1e4, Theta = 1.0, eps = 0: as shipped
    actor.Tick();                                       // not Initialized: silently nothing (OctreeSearch.cpp:49,76)
    actor.CreateSpacePoints(n, 1000.f);
    actor.ShowOctree = true;
    double e0 = 0;
    for (const auto& p : actor.Particles) e0 += p.Mass;
    for (int frame = 0; frame < 10; frame++) actor.Tick(0.016f);
    const nbody::FParticle before = actor.Particles[1];
    actor.PhDeltaTime = 0.f;                            // pause (OctreeSearch.cpp:25)
    actor.Tick();
    const bool paused_ok = actor.Particles[1].Position.X == before.Position.X;
    actor.PhDeltaTime = 0.01f;
    actor.Tick();
    const bool moved = actor.Particles[1].Position.X != before.Position.X;
    const auto boxes = actor.OctreeBoxes();
    const nbody_stats st = actor.Stats();
    double e1 = 0;
    bool finite = true;
    for (const auto& p : actor.Particles) { e1 += p.Mass; finite = finite && std::isfinite(p.Position.X + p.Velocity.X + p.Acceleration.X); }
    std::printf("octree_search: N=%d steps=%lld tree nodes=%d depth=%d leaf boxes=%zu cube size=%.1f central body mass=%.0f at (%.3f, %.3f, %.3f)\n",
                n, (long long)st.steps, st.tree_nodes, st.tree_depth, boxes.size() / 7, st.cube_size, actor.Particles[0].Mass,
                actor.Particles[0].Position.X, actor.Particles[0].Position.Y, actor.Particles[0].Position.Z);
    actor.CleanParticles();                             // restart button: CleanParticles -> CreateSpacePoints
    actor.CreateSpacePoints(n / 2, 500.f, 7);
    actor.Tick();
    const bool ok = paused_ok && moved && finite && e0 == e1 && st.steps == 11 && boxes.size() > 0 && actor.Particles.size() == (size_t)(n / 2) &&
                    actor.Stats().steps == 1;
    std::printf(ok ? "OK\n" : "FAILED\n");
    return ok ? 0 : 1;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "octree_search: %s\n", e.what());
    return 2;
  }
}
