// Deformable reuse (first step): new vertex positions for registered triangle-mesh assets whose
// surface deforms from step to step -- the surface mesh of a deformable body in deformable-rigid contact, which the
// reference keeps in a DeformableSurfaceMeshWithBvh and intersects with the rigid body's pressure-field volume mesh
// through the same SurfaceVolumeIntersector as rigid-compliant contact (deformable_mesh_intersection.cc:83-110). One
// update = DeformableMeshWithBvh::UpdateVertexPositions (deformable_mesh_with_bvh.h:86-93):
//   k_refit_tris    TriangleSurfaceMesh::SetAllPositions: per-triangle vertices and unit normal
//                   (triangle_surface_mesh.h:452-487), the fp32 cull copy, and the leaf's box the way the updater sets
//                   it (Aabb::set_bounds, aabb.h:188-191: centre (lo + up) / 2, half width (up - lo) / 2);
//   k_refit_nodes   BvhUpdater::Update (bvh_updater.h:45-103): topology kept, boxes refitted bottom-up. One thread per
//                   leaf walks towards the root; the second child to arrive at a node (atomic counter) merges the two
//                   child boxes and continues. Also the outward-rounded fp32 node the descent reads;
//   k_refit_blocks  the leaf-block boxes of the half-space pairs' tree cut.
// Deforming VOLUME meshes (both bodies of a deformable x deformable pair carry their signed-distance field on one,
// deformable_field_intersection.cc) take the same path with k_refit_tets in place of k_refit_tris: the vertex field
// values stay, everything that depends on the positions -- gradient, value at the origin, inward normals, face half
// spaces, cull copies, leaf box -- is derived again by the function registration uses (hydro_derive.cuh).
// All on the device, for any number of assets per launch (their pool ranges are contiguous when they were registered
// one after the other; the host groups an update into such runs).
#pragma once
#include <cuda_runtime.h>

#include "hydro_derive.cuh"
#include "hydro_math.cuh"
#include "hydro_types.h"

namespace dbh {

struct RefitArgs {
  const double* verts;      // device pool of the triangle assets' vertices [*][3]
  const int* tri_idx;       // [pool triangle][3]: pool vertex indices
  TriRec* tris;
  TriCull* cull_tri;
  ElemBox* ebox_tri;
  BvhNode* nodes;
  Node32* nodes32;
  const int* node_parent;   // pool index of the parent, -1 for a root and for the unused node 0 of every asset
  int* node_visit;          // zeroed before k_refit_nodes
  int tri_begin, tri_end;   // pool ranges of this run of assets
  int node_begin, node_end;
  const int* order_tri;     // leaf order (asset-local element per pool position)
  // Tetrahedron assets (k_refit_tets): pool tetrahedra [tet_begin, tet_end).
  const int* tet_idx;       // [pool tetrahedron][4]: pool vertex indices
  const double* tet_vp;     // [pool tetrahedron][4]: the vertex field values
  TetRec* tets;
  TetAux* aux;
  TetCullA* cull_a;
  TetCullB* cull_b;
  ElemBox* ebox_tet;
  int tet_begin, tet_end;
  int* error;               // set to 1 + pool index of a tetrahedron that became degenerate
  // What the node / block passes read: the element boxes and leaf order of this run's kind.
  const ElemBox* ebox_leaf;
  const int* order_leaf;
  ElemBox* blocks;
  // Per block of this run: pool position of its first leaf, number of leaves, pool index of element 0 of its asset.
  const int4* block_desc;
  int block_begin, block_end;
};

__device__ __forceinline__ void store_node32(Node32* out, const double c[3], const double h[3], int child) {
  Node32 m;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    m.c[d] = (float)c[d];
    const double need = h[d] + fabs(c[d] - (double)m.c[d]);
    float hf = (float)need;
    if ((double)hf < need) hf = nextafterf(hf, INFINITY);
    m.h[d] = nextafterf(hf, INFINITY);
  }
  m.child = child;
  m.l1 = ((fabsf(m.c[0]) + fabsf(m.c[1])) + fabsf(m.c[2])) + ((m.h[0] + m.h[1]) + m.h[2]);
  *out = m;
}

__global__ void __launch_bounds__(256) k_refit_tris(RefitArgs a) {
  const int t = a.tri_begin + blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= a.tri_end) return;
  D3 v[3];
  TriRec r;
  TriCull c;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    v[i] = load3(a.verts + 3ll * a.tri_idx[3ll * t + i]);
    r.v[i][0] = v[i].x;
    r.v[i][1] = v[i].y;
    r.v[i][2] = v[i].z;
    c.v[i][0] = (float)v[i].x;
    c.v[i][1] = (float)v[i].y;
    c.v[i][2] = (float)v[i].z;
  }
  // TriangleSurfaceMesh::ComputePositionDependentQuantities (triangle_surface_mesh.h:452-487).
  const D3 cr = cross(v[1] - v[0], v[2] - v[0]);
  const double nm = norm(cr);
  const D3 n = (nm != 0.0) ? cr / nm : cr;
  r.n[0] = n.x;
  r.n[1] = n.y;
  r.n[2] = n.z;
  c.n[0] = (float)n.x;
  c.n[1] = (float)n.y;
  c.n[2] = (float)n.z;
  a.tris[t] = r;
  a.cull_tri[t] = c;
  // The leaf's box as BvhUpdater sets it (Aabb::set_bounds).
  ElemBox b;
  const double x[3][3] = {{v[0].x, v[0].y, v[0].z}, {v[1].x, v[1].y, v[1].z}, {v[2].x, v[2].y, v[2].z}};
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const double lo = fmin(fmin(x[0][d], x[1][d]), x[2][d]), up = fmax(fmax(x[0][d], x[1][d]), x[2][d]);
    b.c[d] = (lo + up) / 2;
    b.h[d] = (up - lo) / 2;
  }
  a.ebox_tri[t] = b;
}

__global__ void __launch_bounds__(128) k_refit_tets(RefitArgs a) {
  const int t = a.tet_begin + blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= a.tet_end) return;
  D3 v[4];
  double u[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[i] = load3(a.verts + 3ll * a.tet_idx[4ll * t + i]);
    u[i] = a.tet_vp[4ll * t + i];
  }
  TetRec r;
  TetAux x;
  if (!derive_tet(v, u, &r, &x)) {
    atomicCAS(a.error, 0, t + 1);
    return;
  }
  TetCullA ca;
  TetCullB cb;
  derive_tet_cull(r, x, &ca, &cb);
  a.tets[t] = r;
  a.aux[t] = x;
  a.cull_a[t] = ca;
  a.cull_b[t] = cb;
  // The leaf's box as BvhUpdater sets it (Aabb::set_bounds).
  ElemBox b;
  const double p[4][3] = {{v[0].x, v[0].y, v[0].z}, {v[1].x, v[1].y, v[1].z}, {v[2].x, v[2].y, v[2].z}, {v[3].x, v[3].y, v[3].z}};
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const double lo = fmin(fmin(p[0][d], p[1][d]), fmin(p[2][d], p[3][d])), up = fmax(fmax(p[0][d], p[1][d]), fmax(p[2][d], p[3][d]));
    b.c[d] = (lo + up) / 2;
    b.h[d] = (up - lo) / 2;
  }
  a.ebox_tet[t] = b;
}

__global__ void __launch_bounds__(256) k_refit_nodes(RefitArgs a) {
  int n = a.node_begin + blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= a.node_end) return;
  BvhNode node = a.nodes[n];
  if (!(node.left < 0 && node.right == -1) || a.node_parent[n] == -2) return;   // leaves only start a walk (-2: unused node 0)
  {
    const ElemBox b = a.ebox_leaf[~node.left];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      node.c[d] = b.c[d];
      node.h[d] = b.h[d];
    }
    a.nodes[n] = node;
    store_node32(&a.nodes32[n], node.c, node.h, node.left);
  }
  for (;;) {
    const int p = a.node_parent[n];
    if (p < 0) return;                           // n was the root
    __threadfence();                             // my box is visible before the counter says so
    if (atomicAdd(&a.node_visit[p], 1) == 0) return;   // the sibling's walk will do the parent
    __threadfence();
    BvhNode parent = a.nodes[p];
    const volatile BvhNode* l = &a.nodes[parent.left];
    const volatile BvhNode* r = &a.nodes[parent.right];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      // lower / upper of the children (Aabb::lower / upper: centre -+ half width), then set_bounds.
      const double lo = fmin(l->c[d] - l->h[d], r->c[d] - r->h[d]), up = fmax(l->c[d] + l->h[d], r->c[d] + r->h[d]);
      parent.c[d] = (lo + up) / 2;
      // Half widths only ever bound (the descent is conservative): keep the box closed under the rounding of c.
      parent.h[d] = fmax(up - parent.c[d], parent.c[d] - lo);
    }
    a.nodes[p] = parent;
    store_node32(&a.nodes32[p], parent.c, parent.h, parent.left);
    n = p;
  }
}

__global__ void __launch_bounds__(128) k_refit_blocks(RefitArgs a) {
  const int b = a.block_begin + blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= a.block_end) return;
  const int4 d = a.block_desc[b - a.block_begin];
  double lo[3] = {1e300, 1e300, 1e300}, up[3] = {-1e300, -1e300, -1e300};
  for (int k = 0; k < d.y; ++k) {
    const ElemBox e = a.ebox_leaf[d.z + a.order_leaf[d.x + k]];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      lo[c] = fmin(lo[c], e.c[c] - e.h[c]);
      up[c] = fmax(up[c], e.c[c] + e.h[c]);
    }
  }
  ElemBox out;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    out.c[c] = 0.5 * (lo[c] + up[c]);
    out.h[c] = fmax(up[c] - out.c[c], out.c[c] - lo[c]) * (1 + 1e-15);
  }
  a.blocks[b] = out;
}

// The root boxes of the re-posed assets, gathered for one device-to-host copy (the geometry-level broad phase of the
// no-surface pairs reads them on the host).
__global__ void __launch_bounds__(256) k_gather_roots(const BvhNode* nodes, const int* root_index, int n, double* out /* [n][6] */) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const BvhNode& r = nodes[root_index[i]];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    out[6 * i + d] = r.c[d];
    out[6 * i + 3 + d] = r.h[d];
  }
}

}  // namespace dbh
