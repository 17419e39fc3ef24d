/* nbody.h - C ABI of the B200-native N-body hot path (libnbody_b200.so).
 *
 * This is the drop-in boundary for the one hot path of Milias/ParallelNbody: per-step gravitational force
 * evaluation (all-pairs direct sum, or Barnes-Hut) + kick-drift integration. Each entry point names the
 * member of the reference's simulation actor it replaces (class AOctreeSearch,
 * /root/reference/Source/NBody/OctreeSearch.h:111-149, OctreeSearch.cpp:1-97). The reference has no FFI layer:
 * its "operator API" is that class's public surface, so the C ABI mirrors its verbs one to one and adds the
 * explicit copies a device-resident implementation needs (set / get), error codes, and multi-GPU plumbing.
 *
 * Conventions
 *   - plain pointers and sizes only; all host buffers are caller-owned and copied during the call;
 *   - every function returns NBODY_OK (0) or a negative nbody_status; nbody_last_error() gives the message
 *     for the calling thread. Nothing throws or aborts across this boundary. (The reference has no error
 *     convention at all: void everywhere, silent return when !Initialized, OctreeSearch.cpp:49,76.)
 *   - one host thread drives a handle; calls are synchronous unless stated otherwise;
 *   - there is NO CPU fallback: every compute entry point runs CUDA kernels built for sm_100a and fails with
 *     NBODY_ERR_CUDA when no such device is present.
 *   - theta uses the REFERENCE's convention: a cell is accepted when (cell HALF-width) / (distance to its
 *     centre of mass) < theta (OctreeSearch.h:103 with Size = half-width, h:70-74). The conventional
 *     full-width opening angle is 2 * theta. The reference ships theta = 1.0 (OctreeSearch.cpp:85).
 */
#ifndef NBODY_B200_H
#define NBODY_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NBODY_ABI_VERSION 1

typedef enum nbody_status {
  NBODY_OK = 0,
  NBODY_ERR_INVALID = -1,  /* bad argument / bad state */
  NBODY_ERR_CUDA = -2,     /* CUDA runtime error, or no sm_100 device */
  NBODY_ERR_NCCL = -3,     /* NCCL error, or libnccl.so.2 not loadable */
  NBODY_ERR_OOM = -4,      /* device allocation failed */
  NBODY_ERR_STATE = -5     /* not initialised (reference: Initialized == false) */
} nbody_status;

typedef enum nbody_method {
  NBODY_DIRECT = 0,     /* all-pairs; equals the reference walk at Theta = 0 (OctreeSearch.h:99-108) */
  NBODY_BARNES_HUT = 1  /* Morton sort + LBVH + monopoles + warp-coherent walk; replaces OctreeSearch.h:60-108 */
} nbody_method;

/* 40-byte array-of-structs body record, binary compatible with the reference's FParticle
 * (OctreeSearch.h:9-18): Mass @0, Position @4, Velocity @16, Acceleration @28. */
typedef struct nbody_particle {
  float mass;
  float position[3];
  float velocity[3];
  float acceleration[3];
} nbody_particle;

typedef struct nbody_config {
  uint32_t struct_size;   /* = sizeof(nbody_config); set by nbody_config_default */
  int32_t method;         /* nbody_method */
  float G;                /* reference: 1e4 (OctreeSearch.h:104) */
  float eps;              /* Plummer softening length; reference: 0 (OctreeSearch.h:102 only skips d == 0) */
  float theta;            /* reference convention (see above); reference: 1.0 (OctreeSearch.cpp:85) */
  float ph_delta_time;    /* AOctreeSearch::PhDeltaTime, reference: 0.01 (OctreeSearch.cpp:8); <= 0 pauses */
  int32_t device;         /* CUDA device ordinal for this handle */
  int32_t rank;           /* multi-GPU: this process' rank in [0, world) */
  int32_t world;          /* multi-GPU: number of ranks (1 = single GPU) */
  int32_t leaf_size;      /* Barnes-Hut: max bodies per leaf bucket (default 16; 1 = reference's one-body leaves) */
  int32_t reference_root; /* Barnes-Hut: 1 = root cube as the reference (origin = previous root COM, half-width =
                             max |coord|, OctreeSearch.cpp:47-56,77-79); 0 = tight cube around the bodies */
  int32_t mac;            /* Barnes-Hut acceptance test: 0 = per walk group of <= group_size neighbouring bodies (production: a cell
                             is accepted when half-width / distance(group box, cell COM) < theta - never accepts what the
                             reference's per-body test would open); 1 = per body, exactly OctreeSearch.h:100-107 incl. the
                             visiting order (parity mode, slower) */
  int32_t group_size;     /* Barnes-Hut (mac = 0): bodies per walk group, 32 / 64 / 128 (default 32) */
  int32_t group_pack;     /* Barnes-Hut (mac = 0): tree cells of <= group_pack * group_size bodies are cut into equal walk
                             groups (default 2; larger = fuller lanes, looser group boxes) */
  int32_t bh_exchange;    /* multi-GPU Barnes-Hut: 0 = Morton domain split, body migration and locally-essential-tree exchange
                             (each rank holds only its domain); 1 = replicated tree (every rank holds all bodies, walks its
                             slice of the Morton order, all-gathers positions and velocities); -1 (default) = by size:
                             replicated up to 2^23 bodies, domain split above */
  int32_t reserved[1];
  uint8_t nccl_unique_id[128]; /* multi-GPU: the ncclUniqueId from nbody_comm_unique_id on rank 0. All zeros with world > 1 =
                                  an EMULATED rank: no communicator; the handle evaluates its slice of the bodies given by
                                  nbody_set_bodies and never exchanges (several ranks can then be checked on one GPU) */
  void* stream;           /* optional cudaStream_t to run on (NULL = the handle creates its own) */
} nbody_config;

typedef struct nbody_stats {
  uint32_t struct_size;
  int32_t method;
  int64_t n_global;        /* bodies in the whole simulation */
  int64_t n_local;         /* bodies this rank integrates */
  int64_t steps;           /* steps taken since the bodies were set */
  double interactions;     /* pair (direct) or body-node + body-body (BH) interactions of the LAST force evaluation,
                              summed over this rank's targets */
  double kernel_launches;  /* kernels launched by this handle since creation */
  float ms_last_call;      /* device time of the last nbody_step / nbody_create_octree call (CUDA events on the
                              handle's stream), all steps of that call */
  float ms_force;          /* same call: force kernel(s) only (direct: K1; BH: walk; domain split: local walk beside the
                              LET exchange + the walk of the received points) */
  float ms_build;          /* same call: BH build (bbox, keys, sort, tree, monopoles; domain split: + body migration);
                              0 for direct */
  float ms_integrate;      /* same call: fused reduce + kick-drift */
  float ms_comm;           /* same call: collectives */
  float cube_size;         /* AOctreeSearch::Size after the last ComputeCubeSize */
  int32_t jsplit, i_per_thread, tree_nodes, tree_depth;
  float root_com[3];       /* BH: root centre of mass of the last build (the next reference-mode root origin) */
  float root_mass;
  int32_t walk_groups;     /* BH: groups of the last build */
  int32_t let_points;      /* BH domain split: locally-essential points received from the peers in the last step */
  int32_t equal_mass;      /* direct sum: 1 = all sources carry one mass, the 11-lane-op kernel runs (mass applied in K2) */
  int32_t sort_passes;     /* BH: 8-bit radix passes of the last build (only the key levels the tree needs are sorted) */
  int32_t migrated;        /* BH domain split: bodies this rank received from other ranks in the last step */
  /* BH domain split, last step of the last synchronous call, device time (ms): sending bodies to their domains; boundary
   * tree + export descent + count exchange; local walk; exchange of the export lists + tree over the received points;
   * walk through the received points */
  float ms_let_migrate, ms_let_plan, ms_let_walk_local, ms_let_import, ms_let_walk_let;
} nbody_stats;

typedef struct nbody_sim nbody_sim; /* opaque handle: owns device buffers, stream, events, NCCL communicator */

/* ---- lifecycle ------------------------------------------------------------------------------------ */
int nbody_abi_version(void);
const char* nbody_last_error(void);
/* Fill cfg with the reference's shipped PHYSICS: method = BARNES_HUT, G = 1e4, eps = 0, theta = 1.0, dt = 0.01
 * (OctreeSearch.cpp:8,85; OctreeSearch.h:104), device 0, world 1 - and this library's production TREE SHAPE: leaf_size 16,
 * group walk (mac 0), tight root cube (reference_root 0). That walk never accepts a cell the reference would open, so its
 * forces are at least as accurate as the reference's at the same theta, but they are not bit-for-bit the reference's:
 * the reference's own tree and visiting order are leaf_size 1, mac 1, reference_root 1 (the parity configuration). */
int nbody_config_default(nbody_config* cfg);
/* Replaces the actor's construction, AOctreeSearch::AOctreeSearch (OctreeSearch.cpp:8-12). */
int nbody_create(nbody_sim** out, const nbody_config* cfg);
/* Replaces actor destruction (+ CleanParticles, OctreeSearch.cpp:91-97). NULL is allowed. */
void nbody_destroy(nbody_sim* sim);

/* ---- bodies in ------------------------------------------------------------------------------------ */
/* Replaces AOctreeSearch::CreateSpacePoints(int32 N, float Size) (OctreeSearch.cpp:58-72): N bodies uniform in the
 * slab +-(Size, Size, Size/10), speed 10*randint[25,50] in a random direction, mass randint[1,5000], body 0 = mass
 * 5000 at rest at the origin; generated ON DEVICE from a counter-based RNG with the given seed (the reference
 * draws from the unseeded C rand()). Sets Size = Size and Initialized = true. N >= 1. */
int nbody_create_space_points(nbody_sim* sim, int64_t n, float size, uint64_t seed);
/* Replaces writing the public member TArray<FParticle> Particles (OctreeSearch.h:118): n records, `stride` bytes apart
 * (40 for a packed FParticle array). All ranks pass the same GLOBAL array; each keeps its own share. */
int nbody_set_particles_aos(nbody_sim* sim, const void* particles, int64_t n, size_t stride);
/* Same, from the device-native layout: posm4 = n x (x, y, z, mass), vel4 = n x (vx, vy, vz, unused); vel4 may be NULL
 * (zero velocities). */
int nbody_set_bodies(nbody_sim* sim, const float* posm4, const float* vel4, int64_t n);
/* Replaces AOctreeSearch::CleanParticles (OctreeSearch.cpp:91-97): Initialized = false, bodies and tree dropped
 * (device buffers are kept for reuse). */
int nbody_clean_particles(nbody_sim* sim);

/* ---- the hot path --------------------------------------------------------------------------------- */
/* Replaces AOctreeSearch::ComputeCubeSize (OctreeSearch.cpp:47-56): Size = max_i max(|x|,|y|,|z|), over ALL ranks. */
int nbody_compute_cube_size(nbody_sim* sim, float* size_out);
/* Replaces AOctreeSearch::CreateOctree (OctreeSearch.cpp:74-89): (build the tree, monopoles,) zero and evaluate the
 * accelerations of all bodies at the current positions with the handle's method / theta / eps / G. */
int nbody_create_octree(nbody_sim* sim);
/* Replaces AOctreeSearch::Tick (OctreeSearch.cpp:21-34): if PhDeltaTime > 0: cube size, forces, then
 * v += dt*a ; x += dt*v (kick-drift, OctreeSearch.cpp:28-31). The frame DeltaSeconds the reference ignores is
 * not a parameter. */
int nbody_tick(nbody_sim* sim);
/* nsteps Ticks with an explicit dt (PhDeltaTime is left unchanged); all steps are enqueued without host
 * synchronisation in between and the call returns when the stream is idle. */
int nbody_step(nbody_sim* sim, float dt, int32_t nsteps);
/* Same but returns right after enqueueing; pair with nbody_synchronize. */
int nbody_step_async(nbody_sim* sim, float dt, int32_t nsteps);
int nbody_synchronize(nbody_sim* sim);

/* ---- bodies out (replaces reading Particles[i].Position/Velocity/Acceleration, OctreeSearch.h:118) ---- */
/* Each writes this rank's share into the GLOBAL-size output at the bodies' original indices; with world > 1
 * the caller combines ranks (shares are disjoint). `n` is the capacity of the output in bodies (>= n_global).
 * Barnes-Hut reorders the bodies along the Morton curve every step; the original indices travel with them.
 * Domain-split Barnes-Hut (bh_exchange = 0, world > 1): bodies migrate between ranks, so the read-backs (and the
 * set_* calls, nbody_energy) are COLLECTIVES - every rank calls them together; each rank gets back the rows of its own
 * slice of the caller's order, [rank * ceil(n / world), ...), returned to it by the ranks that hold those bodies now, in
 * one contiguous copy. */
int nbody_get_particles_aos(nbody_sim* sim, void* particles, int64_t n, size_t stride);
int nbody_get_positions(nbody_sim* sim, float* posm4, int64_t n);
int nbody_get_velocities(nbody_sim* sim, float* vel4, int64_t n);
int nbody_get_accelerations(nbody_sim* sim, float* acc4, int64_t n);
/* The rows this rank's read-backs fill: ids[n_local] (capacity cap). Direct sum and domain-split Barnes-Hut: a fixed
 * contiguous slice of the caller's order (the bodies a domain-split rank INTEGRATES change as they migrate: see
 * nbody_stats.n_local / migrated); replicated-tree Barnes-Hut: the bodies of the rank's slice of the Morton order. */
int nbody_get_local_ids(nbody_sim* sim, int64_t* ids, int64_t cap, int64_t* n_local);

/* ---- parameters (replaces the Blueprint-exposed members PhDeltaTime / ShowOctree, OctreeSearch.h:123-127, and the
 *      literals G = 1e4, Theta = 1.0) ---- */
typedef enum nbody_param {
  NBODY_PARAM_G = 0, NBODY_PARAM_EPS = 1, NBODY_PARAM_THETA = 2, NBODY_PARAM_PH_DELTA_TIME = 3,
  NBODY_PARAM_METHOD = 4, NBODY_PARAM_LEAF_SIZE = 5, NBODY_PARAM_REFERENCE_ROOT = 6, NBODY_PARAM_SHOW_OCTREE = 7,
  NBODY_PARAM_INITIALIZED = 8 /* read-only */, NBODY_PARAM_MAC = 9, NBODY_PARAM_GROUP_SIZE = 10, NBODY_PARAM_GROUP_PACK = 11
} nbody_param;
int nbody_set_param(nbody_sim* sim, int32_t which, double value);
int nbody_get_param(nbody_sim* sim, int32_t which, double* value);

/* ---- diagnostics ---------------------------------------------------------------------------------- */
/* Kinetic and potential energy of the whole system (sum over ranks), potential with the same softening as the
 * force law; fp64 accumulation on device. */
int nbody_energy(nbody_sim* sim, double* kinetic, double* potential);
int nbody_stats_get(nbody_sim* sim, nbody_stats* out);
/* Replaces AOctreeSearch::DrawOctreeBoxes (OctreeSearch.cpp:36-45) as a read-back: one record per occupied leaf of the
 * last Barnes-Hut build, boxes7 = (cx, cy, cz, hx, hy, hz, body count); returns the count in *n_boxes (cap = capacity). */
int nbody_octree_boxes(nbody_sim* sim, float* boxes7, int64_t cap, int64_t* n_boxes);
/* Device pointers of this rank's state (posm float4[n_local] inside the gathered array, vel, acc), for
 * zero-copy consumers such as a renderer. Valid until the next set/clean/destroy. */
int nbody_device_ptrs(nbody_sim* sim, void** posm4, void** vel4, void** acc4);

/* Inspection of the last Barnes-Hut build (replaces walking the public Octree* ParticleOctree, OctreeSearch.h:119, through
 * its getters h:43-48). Node k: com4[4k..] = (centre of mass xyz, total mass) as Octree::CenterOfMass / TotalMass;
 * meta4[4k..] = (first child | first body, #children | #bodies, level + 256 * is_leaf, parent or -1); range2[2k..] = body
 * range [begin, end) in Morton order; cell half-width (Octree::Size) = root half-width / 2^level. keys = the sorted 63-bit
 * Morton keys (3 bits per level, 4*X + 2*Y + Z as Octree::GetOctant, h:50-56). Any output pointer may be NULL. */
int nbody_octree_nodes(nbody_sim* sim, float* com4, int32_t* meta4, int32_t* range2, uint64_t* keys, int64_t cap_nodes,
                       int64_t cap_keys, int64_t* n_nodes);

/* Checkpoint / resume (no counterpart in the reference): a flat little-endian file - 80-byte header
 * (magic "NBODYB2", version, n, steps, G, eps, theta, PhDeltaTime, method) + posm float4[n] + vel float4[n] in the caller's
 * original body order. Loading sets the bodies (all ranks read the same file and keep their share), restores the
 * parameters and the step count; the handle's method is kept. Saving needs a single-rank handle. */
int nbody_save_snapshot(nbody_sim* sim, const char* path);
int nbody_load_snapshot(nbody_sim* sim, const char* path);

/* ---- multi-GPU plumbing --------------------------------------------------------------------------- */
/* 128-byte ncclUniqueId; rank 0 creates it and the launcher broadcasts it to all ranks before nbody_create. */
int nbody_comm_unique_id(uint8_t out128[128]);
/* Id of a fresh LOOP-BACK group: `world` handles created with it IN ONE PROCESS (one host thread per handle, usually all on
 * one GPU) form a communicator whose collectives are device-to-device copies - the domain-split Barnes-Hut path
 * (migration + locally-essential-tree exchange) then runs its real code on a single-GPU box. Not a performance path. */
int nbody_comm_loopback_id(uint8_t out128[128]);

/* ---- measurement helpers -------------------------------------------------------------------------- */
/* FP32 FMA-chain microbenchmark on `device`: sustained FFMA throughput in TFLOP/s (2 flops per FMA) and the
 * SM clock (MHz) implied by the in-kernel cycle counter. Used as the measured roofline denominator. */
int nbody_measure_fp32_peak(int32_t device, double* tflops, double* sm_mhz);
/* The Morton-key radix sort on its own (K5): host keys in, sorted keys and the STABLE permutation (idx_out[i] = input
 * position of output i) out; the low key_bits bits are sorted. *ms (may be NULL) = best-of-3 device time of the sort. */
int nbody_sort_pairs_u64(int32_t device, const uint64_t* keys_in, int64_t n, int32_t key_bits, uint64_t* keys_out,
                         uint32_t* idx_out, float* ms);

#ifdef __cplusplus
}
#endif
#endif /* NBODY_B200_H */
