// AOctreeSearch forwarding to libnbody_b200.so - the Unreal-side half of the drop-in.
//
// Replaces /root/reference/Source/NBody/OctreeSearch.h + OctreeSearch.cpp in the game module. The Blueprint-visible
// surface is the reference's (OctreeSearch.h:111-149): UPROPERTYs ShowOctree / PhDeltaTime, UFUNCTIONs CreateSpacePoints /
// CreateOctree / CleanParticles, the public members Size / Particles / Initialized, Tick, ComputeCubeSize. What changed:
//   * class Octree is gone - the tree lives in HBM behind the C ABI (include/nbody.h); `ParticleOctree` survives as an
//     opaque handle so that `ParticleOctree != NULL` keeps meaning "a tree exists";
//   * `Particles` is the host mirror of the device-resident bodies: Tick refreshes it after every step (the renderer and
//     Blueprints read it), and code that WRITES it calls PushParticles() afterwards;
//   * the literals the reference bakes in (G, Theta, no softening, one-body leaves) are members with the same defaults.
// Builds inside UE 4.9 (add include/ and the library to NBody.Build.cs) and - for the tests of this
// repo - against the stand-in engine header oracle/shim/Engine.h.
#pragma once

#include "GameFramework/Actor.h"
#include "OctreeSearch.generated.h"

struct nbody_sim;   // include/nbody.h

// Body record, bit-compatible with nbody_particle (include/nbody.h) and with the reference's FParticle
// (OctreeSearch.h:9-18): 40 bytes, Mass first.
USTRUCT()
struct FParticle {
  GENERATED_USTRUCT_BODY()
  float Mass;
  FVector Position, Velocity, Acceleration;
  FParticle() : Mass(0.f), Position(FVector::ZeroVector), Velocity(FVector::ZeroVector), Acceleration(FVector::ZeroVector) {}
};

UCLASS()
class NBODY_API AOctreeSearch : public AActor {
  GENERATED_BODY()

 public:
  float Size;
  TArray<FParticle> Particles;
  nbody_sim* ParticleOctree;     // device-side simulation state (tree, bodies); NULL until bodies exist
  bool Initialized;

  UPROPERTY(BlueprintReadWrite)
  bool ShowOctree;

  UPROPERTY(BlueprintReadWrite)
  float PhDeltaTime;

  // --- what the reference hard-codes, now data (defaults = the reference's values) -----------------------------
  UPROPERTY(BlueprintReadWrite)
  float Theta;                   // 1.0 (OctreeSearch.cpp:85), half-width / distance convention
  UPROPERTY(BlueprintReadWrite)
  float GravityG;                // 1e4 (OctreeSearch.h:104)
  UPROPERTY(BlueprintReadWrite)
  float Softening;               // 0 (OctreeSearch.h:102)
  UPROPERTY(BlueprintReadWrite)
  bool bDirectSum;               // all-pairs kernel instead of the tree
  UPROPERTY(BlueprintReadWrite)
  bool bReferenceParity;         // one-body leaves, root cube and per-body walk exactly as the CPU actor (slower)

  AOctreeSearch();
  virtual ~AOctreeSearch();

  virtual void BeginPlay() override;
  virtual void Tick(float DeltaSeconds) override;

  void DrawOctreeBoxes(nbody_sim* Sim);
  void ComputeCubeSize();
  // Upload the host mirror after writing Particles by hand (the reference's callers simply wrote the array).
  void PushParticles();

  UFUNCTION(BlueprintCallable, Category = "Octree")
  void CreateSpacePoints(int32 N, float Size = 200);

  UFUNCTION(BlueprintCallable, Category = "Octree")
  void CreateOctree();

  UFUNCTION(BlueprintCallable, Category = "Octree")
  void CleanParticles();

 private:
  bool EnsureSim();
  void PullParticles();
  void ApplyParams();
  int32 SimMethod;               // method the handle was created with (-1 = none)
  bool SimParity;
};
