// TEST INFRASTRUCTURE ONLY (oracle/). Not part of the product path.
//
// Minimal stand-in for the Unreal Engine 4.9 "Engine.h" umbrella header so that the reference's
// Source/NBody/OctreeSearch.{h,cpp} compile UNMODIFIED, by path, with plain g++ (see oracle/Makefile).
// UE 4.9 is pinned by /root/reference/NBody.uproject:3 and is not vendored; this file restates only the
// identifiers the reference's hot path uses (SURVEY.md §8c lists them with their call sites):
//   FVector arithmetic   OctreeSearch.h:92-95,101,104   OctreeSearch.cpp:28-31,51-53
//   FMath random helpers OctreeSearch.cpp:64-66
//   TArray               OctreeSearch.h:118             OctreeSearch.cpp:62,28,91-97
//   AActor               OctreeSearch.h:112             OctreeSearch.cpp:8-19
//   debug drawing        OctreeSearch.cpp:24,40-41 (captured into a per-world list so the
//                        DrawOctreeBoxes output can be compared; UE draws them on screen)
// Semantics follow UE's documented component-wise fp32 behaviour. The one non-obvious choice is
// operator/= which UE implements as a multiply by the fp32 reciprocal.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <vector>

typedef int32_t int32;
typedef uint32_t uint32;
typedef uint8_t uint8;

#define USTRUCT(...)
#define UCLASS(...)
#define UPROPERTY(...)
#define UFUNCTION(...)
#define GENERATED_USTRUCT_BODY(...)
#define GENERATED_BODY(...) public: typedef AActor Super; private:
#define NBODY_API

struct FVector {
  float X, Y, Z;
  static const FVector ZeroVector;
  FVector() {}
  explicit FVector(float f) : X(f), Y(f), Z(f) {}
  FVector(float x, float y, float z) : X(x), Y(y), Z(z) {}
  FVector operator+(const FVector& v) const { return FVector(X + v.X, Y + v.Y, Z + v.Z); }
  FVector operator-(const FVector& v) const { return FVector(X - v.X, Y - v.Y, Z - v.Z); }
  FVector operator*(float s) const { return FVector(X * s, Y * s, Z * s); }
  FVector operator+=(const FVector& v) { X += v.X; Y += v.Y; Z += v.Z; return *this; }
  FVector operator/=(float v) { const float rv = 1.f / v; X *= rv; Y *= rv; Z *= rv; return *this; }
  float GetAbsMax() const { return fmaxf(fmaxf(fabsf(X), fabsf(Y)), fabsf(Z)); }
  static float DistSquared(const FVector& a, const FVector& b) {
    const float dx = b.X - a.X, dy = b.Y - a.Y, dz = b.Z - a.Z;
    return dx * dx + dy * dy + dz * dz;
  }
  static float Dist(const FVector& a, const FVector& b) { return sqrtf(DistSquared(a, b)); }
};
inline FVector operator*(float s, const FVector& v) { return v * s; }

struct FBox {
  FVector Min, Max;
  FBox(const FVector& mn, const FVector& mx) : Min(mn), Max(mx) {}
};

struct FColor {
  uint8 R, G, B, A;
  static const FColor Red, Black, White;
};

// UE's FMath random helpers sit on the C library rand(); seeding is srand() by the caller.
struct FMath {
  static float FRand() { return rand() / (float)RAND_MAX; }
  static int32 RandHelper(int32 a) {
    if (a <= 0) return 0;
    int32 r = (int32)(FRand() * a);
    return r < a ? r : a - 1;
  }
  // The reference calls RandRange(25.0, 50.0) / RandRange(1.0, 5000.0) with double literals; UE 4.9 only
  // has the int32 overload, so the values are integers.
  static int32 RandRange(int32 mn, int32 mx) { return mn + RandHelper(mx - mn + 1); }
  static float FRandRange(float mn, float mx) { return mn + (mx - mn) * FRand(); }
  static FVector RandPointInBox(const FBox& b) {
    return FVector(FRandRange(b.Min.X, b.Max.X), FRandRange(b.Min.Y, b.Max.Y), FRandRange(b.Min.Z, b.Max.Z));
  }
  static FVector VRand() {
    FVector r; float l;
    do {
      r.X = FRand() * 2.f - 1.f; r.Y = FRand() * 2.f - 1.f; r.Z = FRand() * 2.f - 1.f;
      l = r.X * r.X + r.Y * r.Y + r.Z * r.Z;
    } while (l > 1.f || l < 1e-8f);
    return r * (1.f / sqrtf(l));
  }
};

template <class T> class TArray {
  std::vector<T> v;
 public:
  void SetNum(int32 n) { v.resize((size_t)n); }
  int32 Num() const { return (int32)v.size(); }
  void Empty() { v.clear(); v.shrink_to_fit(); }
  T& operator[](int32 i) { return v[(size_t)i]; }
  const T& operator[](int32 i) const { return v[(size_t)i]; }
};

// One record per DrawDebugBox / DrawDebugPoint call (OctreeSearch.cpp:40-41), kept so tests can
// compare the octree read-back with the boxes the reference would have drawn.
struct FDebugDraw { int kind; FVector a, b; };
struct UWorld { std::vector<FDebugDraw> draws; };

inline void FlushPersistentDebugLines(UWorld* w) { if (w) w->draws.clear(); }
inline void DrawDebugBox(UWorld* w, const FVector& c, const FVector& e, const FColor&, bool = false) {
  if (w) w->draws.push_back(FDebugDraw{0, c, e});
}
inline void DrawDebugPoint(UWorld* w, const FVector& p, float, const FColor&, bool = false) {
  if (w) w->draws.push_back(FDebugDraw{1, p, FVector(0, 0, 0)});
}

class AActor {
 public:
  struct { bool bCanEverTick; } PrimaryActorTick;
  UWorld World;
  AActor() { PrimaryActorTick.bCanEverTick = false; }
  virtual ~AActor() {}
  virtual void BeginPlay() {}
  virtual void Tick(float) {}
  UWorld* GetWorld() { return &World; }
  FVector GetActorLocation() const { return FVector(0, 0, 0); }
};
