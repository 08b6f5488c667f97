/*
 * rtb_internal.h -- structures shared by the translation units of librtb200.so.
 * Not part of the C ABI (that is include/rtb200.h).
 */
#ifndef RTB_INTERNAL_H
#define RTB_INTERNAL_H

#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <string>

#include "rtb200.h"

/* ---- byte layouts of the reference records (raytracer.h:60-131) ------------ */

struct RefVec3 { double x, y, z; };
struct RefVertex { RefVec3 pos; double tex[2]; };                                     /* 40 B */
struct RefObject { uint32_t flags; uint32_t pad; double radius; RefVec3 center, color, emission; }; /* 88 B */
struct RefMaterial { uint32_t flags; uint32_t pad; RefVec3 color, emission; double ka, ks, kd; };   /* 80 B */
struct RefSphere { RefVec3 center; double radius; };                                  /* 32 B */
struct RefMesh { size_t num_triangles; const RefVertex *vertices; };                  /* 16 B */
struct RefSceneObject { int32_t type; int32_t pad; RefMaterial material; const void *geometry; }; /* 96 B */

static_assert(sizeof(RefVertex) == 40, "Vertex");
static_assert(sizeof(RefObject) == 88, "Object");
static_assert(sizeof(RefMaterial) == 80, "Material");
static_assert(sizeof(RefSphere) == 32, "Sphere");
static_assert(sizeof(RefMesh) == 16, "TriangleMesh");
static_assert(sizeof(RefSceneObject) == 96, "SceneObject");

#define RT_M_REFLECTION (1u << 2)
#define RT_M_REFRACTION (1u << 3)
#define RT_M_CHECKERED (1u << 4)

/* ---- device geometry --------------------------------------------------------
 * One 48-byte record per primitive, three 16-byte loads (SoA of float4 triples):
 *   sphere:   a = {cx, cy} as two doubles, b = {cz, r} as two doubles, c = {-, gid, obj|SPHERE, -}
 *   triangle: a = {v0x, v0y, v0z, v1x}, b = {v1y, v1z, v2x, v2y}, c = {v2z, gid, obj, -}
 * gid = index of the primitive in the reference's loop order (object order, then
 * triangle order inside a mesh): the tie-break key of the nearest-hit query.
 * Sphere centres and radii stay IEEE double (the exact test needs them); triangle
 * vertices are float (OBJ positions are parsed as float upstream). */
struct PrimRec { float4 a, b, c; };
#define RTB_PRIM_SPHERE_BIT 0x80000000u

/* BVH2 node, 64 bytes: both children's boxes + child references.
 *   n0 = {c0.lo.x, c0.hi.x, c0.lo.y, c0.hi.y}
 *   n1 = {c1.lo.x, c1.hi.x, c1.lo.y, c1.hi.y}
 *   n2 = {c0.lo.z, c0.hi.z, c1.lo.z, c1.hi.z}
 *   n3 = {ref0, ref1, -, -} as int bits
 * ref >= 0: inner node index; ref < 0: leaf, ~ref = (first << 3) | (count - 1), `first`
 * indexing the BVH-ordered primitive array. */
struct BvhNode { float4 n0, n1, n2, n3; };
#define RTB_LEAF_MAX 4
#define RTB_STACK_SIZE 64
#define RTB_REF_NONE 0x7FFFFFFF

struct SceneView
{
  const float4 *nodes; /* 4 float4 per node */
  const float4 *prims; /* 3 float4 per BVH primitive, BVH order */
  const float4 *big;   /* 3 float4 per oversized primitive (tested for every ray) */
  const float4 *mats;  /* 2 float4 per object */
  const float2 *tex;   /* 3 float2 per BVH primitive (BVH order) or NULL */
  int n_prims, n_big, root_ref, n_objects;
  float guard_lo[3], guard_hi[3]; /* box rays are re-based into before FP32 traversal */
};

struct rtb_scene
{
  int device = 0;
  SceneView view{};
  /* owned device allocations */
  float4 *d_nodes = nullptr, *d_prims = nullptr, *d_big = nullptr, *d_mats = nullptr;
  float2 *d_tex = nullptr;
  float *d_scratch = nullptr; /* split planes */
  size_t scratch_bytes = 0;
  void *d_wf = nullptr;       /* wavefront kernels: ray queues + accumulation planes (rtb_wavefront.cu) */
  size_t wf_bytes = 0;
  unsigned long long *d_counters = nullptr;
  rtb_scene_info info{};
};

/* error plumbing */
void rtb_set_error(const std::string &msg);
#define RTB_CUDA(call)                                                                         \
  do                                                                                           \
  {                                                                                            \
    cudaError_t e_ = (call);                                                                   \
    if (e_ != cudaSuccess)                                                                     \
    {                                                                                          \
      rtb_set_error(std::string(#call) + ": " + cudaGetErrorString(e_));                       \
      return RTB_ECUDA;                                                                        \
    }                                                                                          \
  } while (0)

#endif
