// Host side of libdrake_b200_hydro.so: the C-ABI declared in
// include/drake_b200_hydro.h, asset preparation (per-element records, BVH),
// the static geometry-pair list, kernel orchestration on one CUDA stream and
// result export. No CPU implementation of the query exists in this library:
// without a CUDA device every entry point fails with DBH_ERR_NO_DEVICE.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <memory>
#include <string>
#include <vector>

#include "../../include/drake_b200_hydro.h"
#include "hydro_kernels.cuh"
#include "hydro_small.cuh"
#include "hydro_bfs_kernels.cuh"
#include "hydro_derive.cuh"
#include "hydro_refit.cuh"
#include "hydro_lbvh.h"

using namespace dbh;

namespace {

thread_local std::string g_error;

dbh_status Fail(dbh_status s, const std::string& msg) {
  g_error = msg;
  return s;
}

// No C++ exception crosses the C boundary: host allocations (std::bad_alloc, std::length_error) and anything else a
// body throws become a status.
template <class F>
dbh_status Guarded(F&& body) {
  try {
    return body();
  } catch (const std::bad_alloc&) {
    return Fail(DBH_ERR_HOST, "out of host memory");
  } catch (const std::exception& e) {
    return Fail(DBH_ERR_HOST, std::string("host-side failure: ") + e.what());
  } catch (...) {
    return Fail(DBH_ERR_HOST, "host-side failure");
  }
}

#define CUDA_OK(expr)                                                                              \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess)                                                                         \
      return Fail(DBH_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));               \
  } while (0)

template <class T>
struct DevBuf {
  T* p{nullptr};
  size_t n{0};
  ~DevBuf() { reset(); }
  void reset() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  cudaError_t alloc(size_t count) {
    reset();
    if (count == 0) count = 1;
    cudaError_t e = cudaMalloc(&p, count * sizeof(T));
    if (e == cudaSuccess) n = count;
    return e;
  }
  cudaError_t ensure(size_t count) { return count <= n ? cudaSuccess : alloc(count); }
};

// ---------------------------------------------------------------------------
// Asset preparation (host, set-up time only).
struct HostAsset {
  int type{0};
  int num_vertices{0};
  int num_elements{0};
  int height{0};
  std::vector<TetRec> tets;
  std::vector<TetAux> aux;
  std::vector<TriRec> tris;
  std::vector<double> tet_vp;   // compliant meshes: the four vertex pressures of every tetrahedron
  std::vector<BvhNode> nodes;   // links local to this asset; rebased into the pool on upload
  // HalfSpace shape (dbh_register_half_space): no elements, no nodes.
  bool half_space{false};
  double pressure_scale{0}, margin{0};
  // First pool index of this asset's elements (tet pool or tri pool) and nodes.
  int elem_base{0};
  int node_base{0};
  // Bvh::Collide(plane) for the half-space pairs (bvh.h:285-309), flattened to one cut of the tree: the elements in
  // the tree's leaf order (depth first), and the box of every run of kPlaneBlock consecutive leaves.
  // Mesh assets keep their vertices and connectivity: dbh_context_update_vertices re-poses them on the device.
  std::vector<double> rverts;        // [num_vertices * 3], the current positions
  std::vector<int> tri_idx;          // [num_elements * 3 (triangles) or 4 (tetrahedra)], asset-local vertex indices
  int vert_base{0};                  // first pool index of this asset's vertices
  bool deformed{false};              // re-posed since registration: the per-element host records are stale
  bool host_verts_stale{false};      // re-posed from device memory: rverts is refreshed from the pool before it is regrown
  std::vector<int> leaf_order;       // asset-local element indices
  // The surface-seeded soft-soft intersector (hydro_bfs.cuh), compliant meshes registered while the context's
  // intersector is DBH_INTERSECTOR_SURFACE_BFS: the face neighbours of every tetrahedron (VolumeMeshTopology) and a
  // second tree over the boundary triangles, appended to `nodes` (its node k lives at vol_nodes + k - 1, its leaves
  // name the triangle's tetrahedron: tri_to_tet folded in).
  int vol_nodes{0};                  // nodes of the element tree (= nodes.size() without a boundary tree)
  int bfs_nodes{0};                  // nodes of the boundary tree incl. its unused node 0 (0: none)
  int bfs_height{0};
  std::vector<int> neighbors;        // [num_elements * 4], asset-local, -1 on the boundary
  std::vector<ElemBox> block_boxes;  // ceil(num_elements / kPlaneBlock)
  int block_base{0};                 // first pool index of this asset's block boxes
};

struct Box {
  double mn[3], mx[3];
};

// Leaf order + block boxes of a built tree (nodes[1] is the root, the two children of a node are adjacent).
void BuildPlaneBlocks(HostAsset* a) {
  a->leaf_order.clear();
  a->block_boxes.clear();
  if (a->half_space || a->num_elements == 0 || a->nodes.size() < 2) return;
  a->leaf_order.reserve(a->num_elements);
  std::vector<const BvhNode*> leaves;
  leaves.reserve(a->num_elements);
  std::vector<int> stack{1};
  while (!stack.empty()) {
    const int n = stack.back();
    stack.pop_back();
    const BvhNode& node = a->nodes[n];
    if (node.left < 0) {
      a->leaf_order.push_back(~node.left);
      leaves.push_back(&node);
    } else {
      stack.push_back(node.right);
      stack.push_back(node.left);
    }
  }
  const int blocks = (a->num_elements + kPlaneBlock - 1) / kPlaneBlock;
  a->block_boxes.assign(blocks, ElemBox{});
  for (int b = 0; b < blocks; ++b) {
    double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
    for (int k = b * kPlaneBlock; k < std::min(a->num_elements, (b + 1) * kPlaneBlock); ++k)
      for (int d = 0; d < 3; ++d) {
        mn[d] = std::min(mn[d], leaves[k]->c[d] - leaves[k]->h[d]);
        mx[d] = std::max(mx[d], leaves[k]->c[d] + leaves[k]->h[d]);
      }
    for (int d = 0; d < 3; ++d) {
      a->block_boxes[b].c[d] = 0.5 * (mn[d] + mx[d]);
      // Outward: the box must contain every leaf box whatever the rounding of the centre.
      a->block_boxes[b].h[d] = std::max(mx[d] - a->block_boxes[b].c[d], a->block_boxes[b].c[d] - mn[d]) * (1 + 1e-15);
    }
  }
}

// Aabb from min/max exactly as AabbMaker::Compute (aabb.cc:102-117).
void BoxToNode(const Box& b, BvhNode* n) {
  for (int k = 0; k < 3; ++k) {
    n->c[k] = (b.mn[k] + b.mx[k]) / 2;
    n->h[k] = b.mx[k] - n->c[k];
  }
}

struct BuildItem {
  int elem;
  double centroid[3];
};

// Top-down median split on the longest box axis (the reference's strategy,
// bvh.cc:32-104), one element per leaf. Node 0 is unused, node 1 the root, the
// two children of a node are adjacent (so a descent step fetches both with one
// 64 B line of the fp32 copy).
void BuildBvh(const std::vector<Box>& elem_boxes, std::vector<BuildItem>& items, int start, int end,
              std::vector<BvhNode>* nodes, int me, int depth, int* height) {
  *height = std::max(*height, depth);
  Box b = elem_boxes[items[start].elem];
  for (int i = start + 1; i < end; ++i) {
    const Box& e = elem_boxes[items[i].elem];
    for (int k = 0; k < 3; ++k) {
      b.mn[k] = std::min(b.mn[k], e.mn[k]);
      b.mx[k] = std::max(b.mx[k], e.mx[k]);
    }
  }
  BvhNode node{};
  BoxToNode(b, &node);
  if (end - start == 1) {
    node.left = ~items[start].elem;
    node.right = -1;
    (*nodes)[me] = node;
    return;
  }
  int axis = 0;
  if (node.h[1] > node.h[axis]) axis = 1;
  if (node.h[2] > node.h[axis]) axis = 2;
  const int mid = start + (end - start) / 2;
  std::nth_element(items.begin() + start, items.begin() + mid, items.begin() + end,
                   [axis](const BuildItem& x, const BuildItem& y) {
                     if (x.centroid[axis] != y.centroid[axis]) return x.centroid[axis] < y.centroid[axis];
                     return x.elem < y.elem;
                   });
  const int base = static_cast<int>(nodes->size());
  nodes->resize(base + 2);
  node.left = base;
  node.right = base + 1;
  (*nodes)[me] = node;
  BuildBvh(elem_boxes, items, start, mid, nodes, base, depth + 1, height);
  BuildBvh(elem_boxes, items, mid, end, nodes, base + 1, depth + 1, height);
}

template <int kVerts>
void BuildAssetBvh(const double* verts, const int32_t* elems, int num_elems, HostAsset* a) {
  // Host builder (default): top-down median split, the reference's strategy.
  std::vector<Box> boxes(num_elems);
  std::vector<BuildItem> items(num_elems);
  for (int e = 0; e < num_elems; ++e) {
    Box b;
    double c[3] = {0, 0, 0};
    for (int k = 0; k < 3; ++k) {
      b.mn[k] = std::numeric_limits<double>::infinity();
      b.mx[k] = -std::numeric_limits<double>::infinity();
    }
    for (int v = 0; v < kVerts; ++v) {
      const double* p = verts + 3 * elems[kVerts * e + v];
      for (int k = 0; k < 3; ++k) {
        b.mn[k] = std::min(b.mn[k], p[k]);
        b.mx[k] = std::max(b.mx[k], p[k]);
        c[k] += p[k];
      }
    }
    boxes[e] = b;
    items[e].elem = e;
    for (int k = 0; k < 3; ++k) items[e].centroid[k] = c[k] / kVerts;
  }
  a->nodes.clear();
  a->nodes.reserve(2 * static_cast<size_t>(num_elems));
  a->nodes.resize(2);  // [0] unused, [1] root
  a->height = 0;
  BuildBvh(boxes, items, 0, num_elems, &a->nodes, 1, 0, &a->height);
}

}  // namespace

struct dbh_context {
  int device{0};
  int bvh_builder{DBH_BVH_GPU_LBVH};   // measured: 24-31 % fewer box tests than the median split (profiles/README.md)
  int soft_soft_intersector{DBH_INTERSECTOR_VOLUME_BVH};
  DevBuf<int4> d_tet_neighbors;        // per pool tetrahedron (surface intersector only; -2: the asset has no topology)
  cudaDeviceProp prop{};
  std::vector<std::unique_ptr<HostAsset>> assets;
  // Pools: every asset's records concatenated in registration order, so pool
  // indices handed out at registration stay valid when later assets are added.
  long long num_tets{0}, num_tris{0}, num_nodes{0};
  DevBuf<TetRec> d_tets;
  DevBuf<TetAux> d_aux;
  DevBuf<TriRec> d_tris;
  DevBuf<BvhNode> d_nodes;
  DevBuf<Node32> d_nodes32;
  DevBuf<TetCullA> d_cull_a;
  DevBuf<TetCullB> d_cull_b;
  DevBuf<TriCull> d_cull_tri;
  DevBuf<ElemBox> d_ebox_tet, d_ebox_tri;
  DevBuf<double> d_tet_vp;   // [4 * pool tet] vertex pressures (the half-space slicer interpolates them)
  DevBuf<int> d_order_tet, d_order_tri;   // per pool element position: asset-local element index in BVH leaf order
  DevBuf<ElemBox> d_plane_blocks;         // box of every run of kPlaneBlock consecutive leaves, all assets
  long long num_blocks{0};
  // Deformable reuse: the triangle assets' vertex pool and connectivity, parent links of every node, visit counters.
  DevBuf<double> d_rverts;
  DevBuf<int> d_tri_idx, d_tet_idx, d_node_parent, d_node_visit, d_refit_error, d_root_index;
  DevBuf<double> d_root_box;
  DevBuf<int4> d_block_desc;
  long long num_rverts{0};
  int bounds_version{0};   // bumped when an update moves root boxes: batches refresh their geometry bounds
  size_t uploaded_assets{0};
  long long up_tets{0}, up_tris{0}, up_nodes{0};

  Pools pools() const {
    return Pools{d_tets.p,   d_aux.p,      d_tris.p,     d_nodes.p,    d_nodes32.p,
                 d_cull_a.p, d_cull_b.p,   d_cull_tri.p, d_ebox_tet.p, d_ebox_tri.p};
  }

  void Place(HostAsset* a) {
    if (a->vol_nodes == 0) a->vol_nodes = static_cast<int>(a->nodes.size());
    a->node_base = static_cast<int>(num_nodes);
    num_nodes += static_cast<long long>(a->nodes.size());
    if (a->type == DBH_COMPLIANT) {
      a->elem_base = static_cast<int>(num_tets);
      num_tets += a->num_elements;
    } else {
      a->elem_base = static_cast<int>(num_tris);
      num_tris += a->num_elements;
    }
    if (!a->half_space) {
      a->vert_base = static_cast<int>(num_rverts);
      num_rverts += a->num_vertices;
    }
    BuildPlaneBlocks(a);
    a->block_base = static_cast<int>(num_blocks);
    num_blocks += static_cast<long long>(a->block_boxes.size());
  }

  // Uploads assets registered since the last call. Pools grow geometrically;
  // growing re-uploads everything (set-up path).
  dbh_status UploadPools() {
    if (uploaded_assets == assets.size()) return DBH_OK;
    // Assets re-posed from device memory (dbh_context_update_vertices_device): the only current copy of their vertices
    // is in the pool about to be regrown.
    for (size_t i = 0; i < uploaded_assets; ++i) {
      HostAsset& a = *assets[i];
      if (!a.host_verts_stale) continue;
      CUDA_OK(cudaMemcpy(a.rverts.data(), d_rverts.p + 3ll * a.vert_base, a.rverts.size() * sizeof(double), cudaMemcpyDeviceToHost));
      a.host_verts_stale = false;
    }
    auto grow = [](auto& buf, long long need, bool* regrown) -> cudaError_t {
      if (static_cast<long long>(buf.n) >= need && buf.p) return cudaSuccess;
      *regrown = true;
      return buf.alloc(static_cast<size_t>(std::max<long long>(need + need / 2, 1024)));
    };
    bool regrown = false;
    CUDA_OK(grow(d_tets, num_tets, &regrown));
    CUDA_OK(grow(d_aux, num_tets, &regrown));
    CUDA_OK(grow(d_tet_vp, 4 * num_tets, &regrown));
    CUDA_OK(grow(d_tris, num_tris, &regrown));
    CUDA_OK(grow(d_cull_a, num_tets, &regrown));
    CUDA_OK(grow(d_cull_b, num_tets, &regrown));
    CUDA_OK(grow(d_cull_tri, num_tris, &regrown));
    CUDA_OK(grow(d_ebox_tet, num_tets, &regrown));
    CUDA_OK(grow(d_ebox_tri, num_tris, &regrown));
    CUDA_OK(grow(d_nodes, num_nodes, &regrown));
    CUDA_OK(grow(d_nodes32, num_nodes, &regrown));
    CUDA_OK(grow(d_order_tet, num_tets, &regrown));
    CUDA_OK(grow(d_order_tri, num_tris, &regrown));
    CUDA_OK(grow(d_plane_blocks, num_blocks, &regrown));
    CUDA_OK(grow(d_rverts, 3 * num_rverts, &regrown));
    CUDA_OK(grow(d_tri_idx, 3 * num_tris, &regrown));
    CUDA_OK(grow(d_tet_idx, 4 * num_tets, &regrown));
    CUDA_OK(grow(d_node_parent, num_nodes, &regrown));
    CUDA_OK(grow(d_node_visit, num_nodes, &regrown));
    bool any_topology = false;
    for (const auto& a : assets) any_topology = any_topology || !a->neighbors.empty();
    if (any_topology) CUDA_OK(grow(d_tet_neighbors, num_tets, &regrown));
    const size_t first = regrown ? 0 : uploaded_assets;
    std::vector<BvhNode> rebased;
    std::vector<Node32> rebased32;
    for (size_t i = first; i < assets.size(); ++i) {
      const HostAsset& a = *assets[i];
      if (a.half_space) continue;
      if (a.type == DBH_COMPLIANT) {
        CUDA_OK(cudaMemcpy(d_tets.p + a.elem_base, a.tets.data(), a.tets.size() * sizeof(TetRec),
                           cudaMemcpyHostToDevice));
        CUDA_OK(cudaMemcpy(d_tet_vp.p + 4ll * a.elem_base, a.tet_vp.data(), a.tet_vp.size() * sizeof(double),
                           cudaMemcpyHostToDevice));
        CUDA_OK(cudaMemcpy(d_aux.p + a.elem_base, a.aux.data(), a.aux.size() * sizeof(TetAux), cudaMemcpyHostToDevice));
        // fp32 copies for the cull stage.
        std::vector<TetCullA> ca(a.tets.size());
        std::vector<TetCullB> cb(a.tets.size());
        for (size_t e = 0; e < a.tets.size(); ++e) derive_tet_cull(a.tets[e], a.aux[e], &ca[e], &cb[e]);
        CUDA_OK(cudaMemcpy(d_cull_a.p + a.elem_base, ca.data(), ca.size() * sizeof(TetCullA), cudaMemcpyHostToDevice));
        CUDA_OK(cudaMemcpy(d_cull_b.p + a.elem_base, cb.data(), cb.size() * sizeof(TetCullB), cudaMemcpyHostToDevice));
        CUDA_OK(cudaMemcpy(d_rverts.p + 3ll * a.vert_base, a.rverts.data(), a.rverts.size() * sizeof(double),
                           cudaMemcpyHostToDevice));
        std::vector<int> idx(a.tri_idx);
        for (int& i : idx) i += a.vert_base;
        CUDA_OK(cudaMemcpy(d_tet_idx.p + 4ll * a.elem_base, idx.data(), idx.size() * sizeof(int), cudaMemcpyHostToDevice));
        if (d_tet_neighbors.p) {
          std::vector<int> nb(a.neighbors);
          if (nb.empty()) nb.assign(4 * static_cast<size_t>(a.num_elements), -2);
          CUDA_OK(cudaMemcpy(d_tet_neighbors.p + a.elem_base, nb.data(), nb.size() * sizeof(int), cudaMemcpyHostToDevice));
        }
      } else {
        CUDA_OK(cudaMemcpy(d_tris.p + a.elem_base, a.tris.data(), a.tris.size() * sizeof(TriRec),
                           cudaMemcpyHostToDevice));
        std::vector<TriCull> ct(a.tris.size());
        for (size_t e = 0; e < a.tris.size(); ++e)
          for (int d = 0; d < 3; ++d) {
            for (int k = 0; k < 3; ++k) ct[e].v[k][d] = static_cast<float>(a.tris[e].v[k][d]);
            ct[e].n[d] = static_cast<float>(a.tris[e].n[d]);
          }
        CUDA_OK(cudaMemcpy(d_cull_tri.p + a.elem_base, ct.data(), ct.size() * sizeof(TriCull), cudaMemcpyHostToDevice));
        CUDA_OK(cudaMemcpy(d_rverts.p + 3ll * a.vert_base, a.rverts.data(), a.rverts.size() * sizeof(double),
                           cudaMemcpyHostToDevice));
        std::vector<int> idx(a.tri_idx);
        for (int& i : idx) i += a.vert_base;
        CUDA_OK(cudaMemcpy(d_tri_idx.p + 3ll * a.elem_base, idx.data(), idx.size() * sizeof(int), cudaMemcpyHostToDevice));
      }
      {
        // Parent links (pool indices): -1 for the root, -2 for the unused node 0.
        std::vector<int> parent(a.nodes.size(), -1);
        parent[0] = -2;
        for (size_t k = 1; k < a.nodes.size(); ++k)
          if (a.nodes[k].left >= 0) {
            parent[a.nodes[k].left] = static_cast<int>(k) + a.node_base;
            parent[a.nodes[k].right] = static_cast<int>(k) + a.node_base;
          }
        CUDA_OK(cudaMemcpy(d_node_parent.p + a.node_base, parent.data(), parent.size() * sizeof(int), cudaMemcpyHostToDevice));
      }
      CUDA_OK(cudaMemcpy((a.type == DBH_COMPLIANT ? d_order_tet.p : d_order_tri.p) + a.elem_base, a.leaf_order.data(),
                         a.leaf_order.size() * sizeof(int), cudaMemcpyHostToDevice));
      if (!a.block_boxes.empty())
        CUDA_OK(cudaMemcpy(d_plane_blocks.p + a.block_base, a.block_boxes.data(), a.block_boxes.size() * sizeof(ElemBox),
                           cudaMemcpyHostToDevice));
      // Per-element exact boxes = the leaves' boxes.
      {
        std::vector<ElemBox> eb(a.num_elements);
        for (int k = 1; k < a.vol_nodes; ++k) {   // the element tree's leaves (not the boundary tree's)
          const BvhNode& n = a.nodes[k];
          if (n.left < 0 && n.right == -1) {
            ElemBox& x = eb[~n.left];
            for (int d = 0; d < 3; ++d) {
              x.c[d] = n.c[d];
              x.h[d] = n.h[d];
            }
          }
        }
        ElemBox* dst = (a.type == DBH_COMPLIANT ? d_ebox_tet.p : d_ebox_tri.p) + a.elem_base;
        CUDA_OK(cudaMemcpy(dst, eb.data(), eb.size() * sizeof(ElemBox), cudaMemcpyHostToDevice));
      }
      rebased = a.nodes;
      for (BvhNode& n : rebased) {
        if (n.left < 0) {
          n.left = ~(~n.left + a.elem_base);
        } else {
          n.left += a.node_base;
          n.right += a.node_base;
        }
      }
      CUDA_OK(cudaMemcpy(d_nodes.p + a.node_base, rebased.data(), rebased.size() * sizeof(BvhNode),
                         cudaMemcpyHostToDevice));
      // Outward-rounded fp32 copy: |c - c32| is added to the half width, then both roundings go up.
      rebased32.resize(rebased.size());
      for (size_t k = 0; k < rebased.size(); ++k) {
        const BvhNode& n = rebased[k];
        Node32& m = rebased32[k];
        for (int d = 0; d < 3; ++d) {
          m.c[d] = static_cast<float>(n.c[d]);
          const double need = n.h[d] + std::fabs(n.c[d] - static_cast<double>(m.c[d]));
          float h = static_cast<float>(need);
          if (static_cast<double>(h) < need) h = std::nextafterf(h, std::numeric_limits<float>::infinity());
          m.h[d] = std::nextafterf(h, std::numeric_limits<float>::infinity());
        }
        m.child = n.left;
        m.l1 = ((std::fabs(m.c[0]) + std::fabs(m.c[1])) + std::fabs(m.c[2])) + ((m.h[0] + m.h[1]) + m.h[2]);
      }
      CUDA_OK(cudaMemcpy(d_nodes32.p + a.node_base, rebased32.data(), rebased32.size() * sizeof(Node32),
                         cudaMemcpyHostToDevice));
    }
    uploaded_assets = assets.size();
    // The per-element records and boxes just uploaded are the registration-time ones: assets re-posed since then are
    // derived again from their current vertices (set-up path: only after a pool regrowth).
    if (regrown) {
      for (size_t i = 0; i < assets.size(); ++i)
        if (assets[i]->deformed) regrown_deformed.push_back(static_cast<int>(i));
    }
    return DBH_OK;
  }
  std::vector<int> regrown_deformed;   // consumed by Upload()
};

// Timing marks per chunk of scenes: pose + root | per-level descent | cull tet-tri (k_filter + k_small_pairs) | cull
// tet-tet | exact clip tet-tri | exact clip tet-tet (+ exceptional) | half-space setup + count | scan | pack | surface.
constexpr int kStageEvents = 11;

struct dbh_batch {
  dbh_context* ctx{nullptr};
  cudaStream_t stream{nullptr};
  // The surface-seeded soft-soft intersector (hydro_bfs.cuh): chosen when the batch is created, from the context.
  bool surface_bfs{false};
  DevBuf<unsigned long long> d_bfs_table;   // visited set: open addressing, a power of two >= 2 x cap_frontier entries
  unsigned long long last_bfs_pairs{0}, last_bfs_waves{0};
  int num_scenes{0}, num_geoms{0}, num_pairs{0};
  std::vector<int32_t> scene_offsets;
  std::vector<int32_t> geom_asset;
  std::vector<int64_t> geom_id;
  std::vector<PairRec> pairs;
  dbh_batch_info info{};

  // Work-list capacities (entries) of ONE chunk of scenes; result-array capacities of the whole query.
  long long cap_frontier{0}, cap_tri{0}, cap_tet{0}, cap_polys{0};
  long long cap_faces{0}, cap_verts{0};
  // A query processes the scenes in num_chunks contiguous chunks that reuse the work lists, so their
  // size is bounded by work_budget_bytes whatever the batch size (chunk c = scenes
  // [c S / C, (c + 1) S / C)). scene_first_pair[s] = index of the first pair of scene s.
  int num_chunks{1};
  size_t work_budget_bytes{0};
  std::vector<int> scene_first_pair;
  std::vector<int> scene_small_before;   // [num_scenes + 1]: small pairs before each scene
  std::vector<int> scene_mesh_before;    // [num_scenes + 1]: mesh pairs (soft-rigid + soft-soft) before each scene

  // Half-space pairs (kSoftMeshRigidPlane / kSoftPlaneRigidMesh): their indices in the pair list (ascending) and
  // the exclusive prefix of the element counts of their one mesh; per geometry, the parameters of a compliant half space.
  std::vector<int> plane_pairs;
  std::vector<long long> plane_elem_off;
  std::vector<int> plane_nelem, plane_block_base;   // per half-space pair: elements of its mesh, first block box
  DevBuf<int> d_plane_pairs, d_plane_nelem, d_plane_block_base;
  DevBuf<long long> d_plane_elem_off;
  // Contact-surface consumer (k_wrench): per-geometry body origins / spatial velocities / contact properties and the
  // per-surface spatial forces.
  DevBuf<double> d_body_origin, d_body_vel, d_props, d_forces, d_discrete, d_deformable;
  bool discrete_done{false};
  bool deformable_done{false};
  long long discrete_pairs{0};
  bool have_body_origin{false}, have_body_vel{false}, have_props{false}, forces_done{false};
  // Small pairs (both trees at most kSmallMaxNodes nodes): k_small_pairs descends and culls them on chip, one warp
  // per pair. small_mode: 0 never, 1 whenever a pair is eligible, 2 (default) when a chunk holds enough eligible
  // pairs to fill the machine (a handful of pairs is faster through the generic pipeline, which spreads ONE pair
  // over the whole grid). record_candidates: keep the candidate lists of the last query (dbh_batch_copy_candidates).
  int small_mode{2};
  int small_stack_cap{1024};
  int num_large_small_pairs{0};   // small pairs whose descent is split into work items (PairRec::small == 2)
  DevBuf<int2> d_small_items;     // [2][cap_small_items]
  long long cap_small_items{1 << 20};
  int num_small_pairs{0};
  bool record_candidates{false};
  bool small_used{false};           // the last query ran k_small_pairs
  DevBuf<int> d_small_list;
  // Geometry-level broad phase of the no-surface pairs (rejected combinations, rigid-rigid, undefined): per-geometry
  // bounds in the geometry frame, and the list of such pairs whose world boxes overlapped in the last WithFallback
  // query (the pairs the reference hands to MaybeMakePointPair, proximity_engine.cc:958-970).
  std::vector<GeomBound> geom_bounds;
  DevBuf<GeomBound> d_geom_bounds;
  DevBuf<int> d_fallback;
  int num_no_surface_pairs{0};
  int last_query_kind{DBH_QUERY_STRICT};
  int bounds_version{0};
  DevBuf<unsigned int> d_plane_emit;  // the current chunk's emitting half-space elements (chunk-local indices)
  int plane_emit_sized_for{0};        // the chunk count d_plane_emit was last sized for
  DevBuf<double> d_geom_hs;

  DevBuf<PairRec> d_pairs;
  DevBuf<Xf> d_poses_own;
  const Xf* d_poses{nullptr};
  DevBuf<Xf> d_pair_X;
  DevBuf<PairF32> d_pair_f32;
  DevBuf<int4> d_frontier[2];
  DevBuf<int4> d_cand_tri, d_cand_tet, d_slow, d_undecided;
  long long cap_slow{1 << 16};
  long long cap_undecided{1 << 19};   // leaf pairs left to the fp64 leaf predicate, per chunk
  DevBuf<Counters> d_counters;
  DevBuf<int> d_pair_polys, d_pair_verts, d_surf_index, d_surf_pair, d_big_list;
  DevBuf<unsigned long long> d_pair_cursor;
  DevBuf<long long> d_poly_off, d_vert_off, d_block_sums, d_surf_face_off, d_surf_vert_off;
  DevBuf<int4> d_soup_hdr;
  DevBuf<double> d_soup_v, d_soup_nhat;
  // Final results in the caller's layout.
  DevBuf<double> r_vertices, r_pressures, r_areas, r_normals, r_centroids, r_e_centroid, r_grad_eM, r_grad_eN;
  DevBuf<int> r_face_data, r_elem_a, r_elem_b, r_face_pair, r_vertex_origin;
  DevBuf<double> r_surf_total_area, r_surf_centroid;
  DevBuf<int> r_surf_scene;
  DevBuf<long long> r_surf_id_m, r_surf_id_n;
  DevBuf<unsigned char> r_surf_flags;
  long long cap_face_data{0};
  Counters* h_counters{nullptr};  // pinned

  int representation{DBH_POLYGON};
  bool query_done{false};
  bool query_enqueued{false};
  bool timing{false};
  bool verify{false};
  bool normal_filter{true};
  int last_launches{0};   // kernels launched by the last query
  // Timing (optional): two events around the whole query, eight per chunk around its stages.
  cudaEvent_t ev_begin{nullptr}, ev_end{nullptr};
  std::vector<cudaEvent_t> ev_chunk;   // kStageEvents per chunk that ran in the last query
  int timed_chunks{0};
  float times_ms[8]{};
  float stage_ms[kStageEvents]{};      // kStageEvents - 1 stages, then the whole query

  ~dbh_batch() {
    if (h_counters) cudaFreeHost(h_counters);
    if (ev_begin) cudaEventDestroy(ev_begin);
    if (ev_end) cudaEventDestroy(ev_end);
    for (auto& e : ev_chunk)
      if (e) cudaEventDestroy(e);
    if (stream) cudaStreamDestroy(stream);
  }

  size_t WorkBytes(long long fr, long long tri, long long tet, long long polys) const {
    return static_cast<size_t>(fr) * 32 + static_cast<size_t>(tri + tet) * 16 +
           static_cast<size_t>(polys) * (16 + kMaxPolyVerts * 32 + 24);
  }
  int MaxPairsPerChunk() const {
    int m = 0;
    for (int c = 0; c < num_chunks; ++c) {
      const int s0 = static_cast<int>(static_cast<long long>(num_scenes) * c / num_chunks);
      const int s1 = static_cast<int>(static_cast<long long>(num_scenes) * (c + 1) / num_chunks);
      m = std::max(m, scene_first_pair[s1] - scene_first_pair[s0]);
    }
    return m;
  }

  // The visited set of the surface intersector: a power of two, at least four times the queue (= tet-tet survivor list).
  size_t BfsTableEntries() const {
    size_t n = 1024;
    while (n < 4 * static_cast<size_t>(cap_frontier)) n <<= 1;
    return n;
  }

  dbh_status AllocWorkBuffers() {
    CUDA_OK(d_frontier[0].ensure(cap_frontier));
    CUDA_OK(d_frontier[1].ensure(cap_frontier));
    if (surface_bfs && info.num_soft_soft_pairs) CUDA_OK(d_bfs_table.ensure(BfsTableEntries()));
    CUDA_OK(d_cand_tri.ensure(info.num_soft_rigid_pairs ? cap_tri : 1));
    CUDA_OK(d_cand_tet.ensure(info.num_soft_soft_pairs ? cap_tet : 1));
    CUDA_OK(d_slow.ensure(cap_slow));
    if (num_small_pairs) CUDA_OK(d_small_list.ensure(2 * static_cast<size_t>(MaxPairsPerChunk())));
    CUDA_OK(d_undecided.ensure(cap_undecided));
    if (num_large_small_pairs) CUDA_OK(d_small_items.ensure(2 * static_cast<size_t>(cap_small_items)));
    if (!plane_pairs.empty() && plane_emit_sized_for != num_chunks) {  // depends on the chunking only
      long long most = 1;
      for (int c = 0; c < num_chunks; ++c) {
        const int s0 = static_cast<int>(static_cast<long long>(num_scenes) * c / num_chunks);
        const int s1 = static_cast<int>(static_cast<long long>(num_scenes) * (c + 1) / num_chunks);
        const auto q0 = std::lower_bound(plane_pairs.begin(), plane_pairs.end(), scene_first_pair[s0]) - plane_pairs.begin();
        const auto q1 = std::lower_bound(plane_pairs.begin(), plane_pairs.end(), scene_first_pair[s1]) - plane_pairs.begin();
        most = std::max(most, plane_elem_off[q1] - plane_elem_off[q0]);
      }
      if (most >= (1ll << 32))
        return Fail(DBH_ERR_CAPACITY, "more than 2^32 half-space elements in one chunk of scenes: raise dbh_batch_set_chunks");
      CUDA_OK(d_plane_emit.ensure(static_cast<size_t>(most)));
      plane_emit_sized_for = num_chunks;
    }
    CUDA_OK(d_soup_hdr.ensure(cap_polys));
    CUDA_OK(d_soup_v.ensure(static_cast<size_t>(cap_polys) * kMaxPolyVerts * 4));
    CUDA_OK(d_soup_nhat.ensure(static_cast<size_t>(cap_polys) * 3));
    cap_face_data = std::max(cap_faces + cap_verts, 3 * cap_faces);
    CUDA_OK(r_vertices.ensure(static_cast<size_t>(cap_verts) * 3));
    CUDA_OK(r_pressures.ensure(cap_verts));
    CUDA_OK(r_face_data.ensure(cap_face_data));
    CUDA_OK(r_areas.ensure(cap_faces));
    CUDA_OK(r_normals.ensure(static_cast<size_t>(cap_faces) * 3));
    CUDA_OK(r_centroids.ensure(static_cast<size_t>(cap_faces) * 3));
    CUDA_OK(r_e_centroid.ensure(cap_faces));
    CUDA_OK(r_grad_eM.ensure(static_cast<size_t>(cap_faces) * 3));
    CUDA_OK(r_grad_eN.ensure(static_cast<size_t>(cap_faces) * 3));
    CUDA_OK(r_elem_a.ensure(cap_faces));
    CUDA_OK(r_elem_b.ensure(cap_faces));
    CUDA_OK(r_face_pair.ensure(cap_faces));
    if (!plane_pairs.empty()) CUDA_OK(r_vertex_origin.ensure(cap_verts));
    return DBH_OK;
  }
};

namespace {
// Builds the asset's BVH with the context's builder (the GPU LBVH needs at least two elements).
template <int kVerts>
dbh_status BuildBvhFor(dbh_context* ctx, const double* verts, int nv, const int32_t* elems, int ne, HostAsset* a) {
  if (ctx->bvh_builder == DBH_BVH_GPU_LBVH && ne >= 2) {
    std::string err;
    if (!BuildLbvhOnDevice(verts, nv, elems, ne, kVerts, &a->nodes, &a->height, &err)) return Fail(DBH_ERR_CUDA, err);
    return DBH_OK;
  }
  BuildAssetBvh<kVerts>(verts, elems, ne, a);
  return DBH_OK;
}
}  // namespace

namespace {
// The surface side of a compliant mesh for the surface-seeded soft-soft intersector (hydro_bfs.cuh): the face
// neighbours of every tetrahedron (VolumeMeshTopology::Initialize, volume_mesh_topology.cc:13-68; face f is the one
// opposite vertex f) by sorting the faces instead of a std::map, the boundary triangles with their tetrahedron
// (ConvertVolumeToSurfaceMeshWithBoundaryVertices' tri_to_tet, volume_to_surface_mesh.cc:152-187), and a tree over
// the boundary triangles built by the context's builder, appended to the asset's nodes.
dbh_status BuildSurfaceSide(dbh_context* ctx, const double* verts, int nv, const int32_t* tets, int nt, HostAsset* a) {
  struct Face {
    int v[3];
    int tet, f;
  };
  std::vector<Face> faces(4 * static_cast<size_t>(nt));
  for (int e = 0; e < nt; ++e)
    for (int f = 0; f < 4; ++f) {
      Face& x = faces[4 * static_cast<size_t>(e) + f];
      x.v[0] = tets[4 * e + (f + 1) % 4];
      x.v[1] = tets[4 * e + (f + 2) % 4];
      x.v[2] = tets[4 * e + (f + 3) % 4];
      std::sort(x.v, x.v + 3);
      x.tet = e;
      x.f = f;
    }
  std::sort(faces.begin(), faces.end(), [](const Face& p, const Face& q) {
    if (p.v[0] != q.v[0]) return p.v[0] < q.v[0];
    if (p.v[1] != q.v[1]) return p.v[1] < q.v[1];
    if (p.v[2] != q.v[2]) return p.v[2] < q.v[2];
    return p.tet < q.tet;
  });
  a->neighbors.assign(4 * static_cast<size_t>(nt), -1);
  std::vector<int32_t> boundary;   // 3 vertex indices per boundary triangle
  std::vector<int> tri_to_tet;
  auto same = [](const Face& p, const Face& q) { return p.v[0] == q.v[0] && p.v[1] == q.v[1] && p.v[2] == q.v[2]; };
  for (size_t i = 0; i < faces.size();) {
    size_t j = i + 1;
    while (j < faces.size() && same(faces[i], faces[j])) ++j;
    if (j - i == 1) {
      boundary.insert(boundary.end(), faces[i].v, faces[i].v + 3);
      tri_to_tet.push_back(faces[i].tet);
    } else if (j - i == 2) {
      a->neighbors[4 * static_cast<size_t>(faces[i].tet) + faces[i].f] = faces[i + 1].tet;
      a->neighbors[4 * static_cast<size_t>(faces[i + 1].tet) + faces[i + 1].f] = faces[i].tet;
    } else {
      return Fail(DBH_ERR_INVALID_ARGUMENT, "a face is shared by more than two tetrahedra: not a manifold volume mesh");
    }
    i = j;
  }
  const int ntri = static_cast<int>(tri_to_tet.size());
  if (ntri < 2) return Fail(DBH_ERR_INVALID_ARGUMENT, "the volume mesh has no boundary");
  HostAsset tree;
  if (dbh_status st = BuildBvhFor<3>(ctx, verts, nv, boundary.data(), ntri, &tree)) return st;
  a->vol_nodes = static_cast<int>(a->nodes.size());
  a->bfs_nodes = static_cast<int>(tree.nodes.size());
  a->bfs_height = tree.height;
  const int shift = a->vol_nodes - 1;   // the boundary tree's node k lives at vol_nodes + k - 1 (its node 0 is unused)
  for (size_t k = 1; k < tree.nodes.size(); ++k) {
    BvhNode n = tree.nodes[k];
    if (n.left < 0) {
      n.left = ~tri_to_tet[~n.left];
    } else {
      n.left += shift;
      n.right += shift;
    }
    a->nodes.push_back(n);
  }
  return DBH_OK;
}
}  // namespace

namespace {
// k_clip's launch shapes: threads per block x resident blocks per SM, bounded by the polygon scratch in shared memory
// (2 x 8 x 24 B per tet-tet thread, 2 x 7 x 24 B per tet-triangle thread).
struct ClipLaunch {
  void (*fn)(ClipArgs);
  int threads, blocks_per_sm, smem;
};
ClipLaunch ClipConfig(bool tet) {
  // Measured (profiles/r2v_clip_launch_shapes.txt): tet-triangle 64 x 10 (20 warps, 94 registers, no spills) is 5 % faster
  // than 128 x 4; tet-tet needs 120 registers and loses more to spills under a 96-register bound than 18-20 warps give.
  return tet ? ClipLaunch{k_clip<true, 128, 4>, 128, 4, clip_smem_bytes(128, 8)} : ClipLaunch{k_clip<false, 64, 10>, 64, 10, clip_smem_bytes(64, 7)};
}

// k_small_pairs by element kind and pass (hydro_small.cuh). Resident 128-thread blocks per SM, by measurement
// (profiles/r2s_small_pairs_blocks_per_sm.txt): 6 for tet-triangle pairs (80 registers, no spills), 5 for tet-tet pairs
// (96 registers; 6 spills and is 8 % slower).
constexpr int kSmallBlocksTri = 6, kSmallBlocksTet = 5;
using SmallKernelFn = void (*)(SmallArgs);
SmallKernelFn SmallKernel(bool tet, int pass) {
  static const SmallKernelFn kTri[3] = {k_small_pairs<false, kSmallBlocksTri, 0>, k_small_pairs<false, kSmallBlocksTri, 1>,
                                        k_small_pairs<false, kSmallBlocksTri, 2>};
  static const SmallKernelFn kTetK[3] = {k_small_pairs<true, kSmallBlocksTet, 0>, k_small_pairs<true, kSmallBlocksTet, 1>,
                                         k_small_pairs<true, kSmallBlocksTet, 2>};
  return (tet ? kTetK : kTri)[pass];
}
constexpr int kSmallSplitNodes = 2048;

}  // namespace

extern "C" {

const char* dbh_last_error(void) { return g_error.c_str(); }

dbh_status dbh_context_set_bvh_builder(dbh_context* ctx, int32_t builder) {
  return Guarded([&]() -> dbh_status {
  if (!ctx || (builder != DBH_BVH_HOST_MEDIAN_SPLIT && builder != DBH_BVH_GPU_LBVH))
    return Fail(DBH_ERR_INVALID_ARGUMENT, "set_bvh_builder: bad argument");
  ctx->bvh_builder = builder;
  return DBH_OK;
  });
}

dbh_status dbh_context_set_soft_soft_intersector(dbh_context* ctx, int32_t intersector) {
  return Guarded([&]() -> dbh_status {
  if (!ctx || (intersector != DBH_INTERSECTOR_VOLUME_BVH && intersector != DBH_INTERSECTOR_SURFACE_BFS))
    return Fail(DBH_ERR_INVALID_ARGUMENT, "set_soft_soft_intersector: bad argument");
  ctx->soft_soft_intersector = intersector;
  return DBH_OK;
  });
}

dbh_status dbh_asset_copy_bvh(dbh_context* ctx, int32_t asset, double* boxes, int32_t* left) {
  return Guarded([&]() -> dbh_status {
  if (!ctx || asset < 0 || asset >= static_cast<int>(ctx->assets.size()))
    return Fail(DBH_ERR_INVALID_ARGUMENT, "asset_copy_bvh: bad argument");
  const HostAsset& a = *ctx->assets[asset];
  for (size_t k = 0; k < static_cast<size_t>(a.vol_nodes); ++k) {
    if (boxes) {
      for (int d = 0; d < 3; ++d) {
        boxes[6 * k + d] = a.nodes[k].c[d];
        boxes[6 * k + 3 + d] = a.nodes[k].h[d];
      }
    }
    if (left) left[k] = a.nodes[k].left;
  }
  return DBH_OK;
  });
}
dbh_status dbh_asset_copy_topology(dbh_context* ctx, int32_t asset, int32_t* neighbors, int32_t* num_boundary_triangles) {
  return Guarded([&]() -> dbh_status {
  if (!ctx || asset < 0 || asset >= static_cast<int>(ctx->assets.size()))
    return Fail(DBH_ERR_INVALID_ARGUMENT, "asset_copy_topology: bad argument");
  const HostAsset& a = *ctx->assets[asset];
  if (a.neighbors.empty())
    return Fail(DBH_ERR_INVALID_ARGUMENT, "asset_copy_topology: the asset was not registered for the surface intersector");
  if (neighbors) std::memcpy(neighbors, a.neighbors.data(), a.neighbors.size() * sizeof(int32_t));
  if (num_boundary_triangles) *num_boundary_triangles = a.bfs_nodes / 2;   // 2 n nodes incl. the unused node 0
  return DBH_OK;
  });
}
int32_t dbh_abi_version(void) { return 1000; }

dbh_status dbh_context_create(int32_t device_ordinal, dbh_context** out) {
  return Guarded([&]() -> dbh_status {
  if (!out) return Fail(DBH_ERR_INVALID_ARGUMENT, "out is null");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return Fail(DBH_ERR_NO_DEVICE, std::string("no CUDA device (") + cudaGetErrorString(e) +
                                       "); this back end has no CPU fallback");
  if (device_ordinal < 0 || device_ordinal >= count) return Fail(DBH_ERR_INVALID_ARGUMENT, "bad device ordinal");
  auto ctx = std::make_unique<dbh_context>();
  ctx->device = device_ordinal;
  CUDA_OK(cudaSetDevice(device_ordinal));
  CUDA_OK(cudaGetDeviceProperties(&ctx->prop, device_ordinal));
  if (ctx->prop.major < 10)
    return Fail(DBH_ERR_NO_DEVICE, "device is not sm_100 class; kernels are built for sm_100a only");
  for (int tet = 0; tet < 2; ++tet) {
    const ClipLaunch cl = ClipConfig(tet != 0);
    CUDA_OK(cudaFuncSetAttribute(cl.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, cl.smem));
  }
  *out = ctx.release();
  return DBH_OK;
  });
}

void dbh_context_destroy(dbh_context* ctx) { delete ctx; }

dbh_status dbh_register_compliant_mesh(dbh_context* ctx, int32_t nv, const double* verts, int32_t nt,
                                       const int32_t* tets, const double* pressure, int32_t* asset_out) {
  return Guarded([&]() -> dbh_status {
  if (!ctx || !verts || !tets || !pressure || !asset_out || nv <= 0 || nt <= 0)
    return Fail(DBH_ERR_INVALID_ARGUMENT, "register_compliant_mesh: bad argument");
  CUDA_OK(cudaSetDevice(ctx->device));
  for (int i = 0; i < 4 * nt; ++i)
    if (tets[i] < 0 || tets[i] >= nv) return Fail(DBH_ERR_INVALID_ARGUMENT, "tet vertex index out of range");
  auto a = std::make_unique<HostAsset>();
  a->type = DBH_COMPLIANT;
  a->num_vertices = nv;
  a->num_elements = nt;
  a->tets.resize(nt);
  a->aux.resize(nt);
  a->tet_vp.resize(4 * static_cast<size_t>(nt));
  a->rverts.assign(verts, verts + 3ll * nv);     // kept: dbh_context_update_vertices re-poses deforming volume meshes
  a->tri_idx.assign(tets, tets + 4ll * nt);
  for (int e = 0; e < nt; ++e) {
    D3 v[4];
    double u[4];
    for (int i = 0; i < 4; ++i) {
      const int vi = tets[4 * e + i];
      v[i] = load3(verts + 3 * vi);
      u[i] = pressure[vi];
      a->tet_vp[4 * static_cast<size_t>(e) + i] = u[i];
    }
    if (!derive_tet(v, u, &a->tets[e], &a->aux[e]))
      return Fail(DBH_ERR_DEGENERATE_FIELD, "tetrahedron " + std::to_string(e) + " is degenerate");
  }
  if (dbh_status st = BuildBvhFor<4>(ctx, verts, nv, tets, nt, a.get())) return st;
  if (ctx->soft_soft_intersector == DBH_INTERSECTOR_SURFACE_BFS)
    if (dbh_status st = BuildSurfaceSide(ctx, verts, nv, tets, nt, a.get())) return st;
  ctx->Place(a.get());
  ctx->assets.push_back(std::move(a));
  *asset_out = static_cast<int32_t>(ctx->assets.size()) - 1;
  return DBH_OK;
  });
}

dbh_status dbh_register_rigid_mesh(dbh_context* ctx, int32_t nv, const double* verts, int32_t nt,
                                   const int32_t* tris, int32_t* asset_out) {
  return Guarded([&]() -> dbh_status {
  if (!ctx || !verts || !tris || !asset_out || nv <= 0 || nt <= 0)
    return Fail(DBH_ERR_INVALID_ARGUMENT, "register_rigid_mesh: bad argument");
  CUDA_OK(cudaSetDevice(ctx->device));
  for (int i = 0; i < 3 * nt; ++i)
    if (tris[i] < 0 || tris[i] >= nv) return Fail(DBH_ERR_INVALID_ARGUMENT, "triangle vertex index out of range");
  auto a = std::make_unique<HostAsset>();
  a->type = DBH_RIGID;
  a->num_vertices = nv;
  a->num_elements = nt;
  a->rverts.assign(verts, verts + 3ll * nv);
  a->tri_idx.assign(tris, tris + 3ll * nt);
  a->tris.resize(nt);
  for (int e = 0; e < nt; ++e) {
    TriRec& r = a->tris[e];
    D3 v[3];
    for (int i = 0; i < 3; ++i) {
      v[i] = load3(verts + 3 * tris[3 * e + i]);
      r.v[i][0] = v[i].x;
      r.v[i][1] = v[i].y;
      r.v[i][2] = v[i].z;
    }
    // TriangleSurfaceMesh::ComputePositionDependentQuantities (triangle_surface_mesh.h:452-487).
    const D3 c = cross(v[1] - v[0], v[2] - v[0]);
    const double nm = norm(c);
    const D3 n = (nm != 0.0) ? c / nm : c;
    r.n[0] = n.x;
    r.n[1] = n.y;
    r.n[2] = n.z;
  }
  if (dbh_status st = BuildBvhFor<3>(ctx, verts, nv, tris, nt, a.get())) return st;
  ctx->Place(a.get());
  ctx->assets.push_back(std::move(a));
  *asset_out = static_cast<int32_t>(ctx->assets.size()) - 1;
  return DBH_OK;
  });
}

// Re-derives everything that depends on the vertex positions for the listed mesh assets, whose current positions are in
// the device vertex pool (and in HostAsset::rverts): per-element records, leaf boxes, the trees' boxes bottom-up, the
// half-space pairs' leaf-block boxes. The list is grouped into runs of assets of one kind whose pool ranges follow one
// another (assets registered consecutively): three launches per run.
static dbh_status RefitAssets(dbh_context* ctx, const std::vector<int>& ids) {
  CUDA_OK(ctx->d_refit_error.ensure(1));
  CUDA_OK(cudaMemsetAsync(ctx->d_refit_error.p, 0, sizeof(int), nullptr));
  size_t k = 0;
  while (k < ids.size()) {
    const HostAsset& first = *ctx->assets[ids[k]];
    size_t k1 = k + 1;
    while (k1 < ids.size()) {
      const HostAsset& prev = *ctx->assets[ids[k1 - 1]];
      const HostAsset& next = *ctx->assets[ids[k1]];
      if (next.type != prev.type || next.elem_base != prev.elem_base + prev.num_elements ||
          next.node_base != prev.node_base + static_cast<int>(prev.nodes.size()) ||
          next.block_base != prev.block_base + static_cast<int>(prev.block_boxes.size()))
        break;
      ++k1;
    }
    const HostAsset& last = *ctx->assets[ids[k1 - 1]];
    const bool tets = first.type == DBH_COMPLIANT;
    RefitArgs ra{};
    ra.verts = ctx->d_rverts.p;
    ra.tri_idx = ctx->d_tri_idx.p;
    ra.tris = ctx->d_tris.p;
    ra.cull_tri = ctx->d_cull_tri.p;
    ra.ebox_tri = ctx->d_ebox_tri.p;
    ra.nodes = ctx->d_nodes.p;
    ra.nodes32 = ctx->d_nodes32.p;
    ra.node_parent = ctx->d_node_parent.p;
    ra.node_visit = ctx->d_node_visit.p;
    ra.order_tri = ctx->d_order_tri.p;
    ra.tet_idx = ctx->d_tet_idx.p;
    ra.tet_vp = ctx->d_tet_vp.p;
    ra.tets = ctx->d_tets.p;
    ra.aux = ctx->d_aux.p;
    ra.cull_a = ctx->d_cull_a.p;
    ra.cull_b = ctx->d_cull_b.p;
    ra.ebox_tet = ctx->d_ebox_tet.p;
    ra.error = ctx->d_refit_error.p;
    ra.ebox_leaf = tets ? ctx->d_ebox_tet.p : ctx->d_ebox_tri.p;
    ra.order_leaf = tets ? ctx->d_order_tet.p : ctx->d_order_tri.p;
    (tets ? ra.tet_begin : ra.tri_begin) = first.elem_base;
    (tets ? ra.tet_end : ra.tri_end) = last.elem_base + last.num_elements;
    ra.node_begin = first.node_base;
    ra.node_end = last.node_base + static_cast<int>(last.nodes.size());
    ra.blocks = ctx->d_plane_blocks.p;
    std::vector<int4> desc;
    for (size_t j = k; j < k1; ++j) {
      const HostAsset& a = *ctx->assets[ids[j]];
      for (int b = 0; b < static_cast<int>(a.block_boxes.size()); ++b)
        desc.push_back(make_int4(a.elem_base + b * kPlaneBlock, std::min(kPlaneBlock, a.num_elements - b * kPlaneBlock),
                                 a.elem_base, 0));
    }
    CUDA_OK(ctx->d_block_desc.ensure(std::max<size_t>(desc.size(), 1)));
    CUDA_OK(cudaMemcpyAsync(ctx->d_block_desc.p, desc.data(), desc.size() * sizeof(int4), cudaMemcpyHostToDevice, nullptr));
    ra.block_desc = ctx->d_block_desc.p;
    ra.block_begin = first.block_base;
    ra.block_end = first.block_base + static_cast<int>(desc.size());
    CUDA_OK(cudaMemsetAsync(ctx->d_node_visit.p + ra.node_begin, 0, sizeof(int) * (ra.node_end - ra.node_begin), nullptr));
    const int nelem = last.elem_base + last.num_elements - first.elem_base, nnode = ra.node_end - ra.node_begin;
    if (tets) k_refit_tets<<<(nelem + 127) / 128, 128>>>(ra);
    else k_refit_tris<<<(nelem + 255) / 256, 256>>>(ra);
    k_refit_nodes<<<(nnode + 255) / 256, 256>>>(ra);
    if (!desc.empty()) k_refit_blocks<<<(static_cast<int>(desc.size()) + 127) / 128, 128>>>(ra);
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaStreamSynchronize(nullptr));   // d_block_desc and `desc` are reused by the next run
    k = k1;
  }
  int bad = 0;
  CUDA_OK(cudaMemcpy(&bad, ctx->d_refit_error.p, sizeof(int), cudaMemcpyDeviceToHost));
  if (bad)
    return Fail(DBH_ERR_DEGENERATE_FIELD, "update_vertices: tetrahedron " + std::to_string(bad - 1) +
                                              " (pool index) became degenerate: its field has no gradient");
  // Root boxes feed the geometry-level broad phase of the no-surface pairs (batches refresh them lazily): gathered on
  // the device, one copy for all assets.
  {
    const int n = static_cast<int>(ids.size());
    std::vector<int> roots(n);
    for (int i = 0; i < n; ++i) roots[i] = ctx->assets[ids[i]]->node_base + 1;
    CUDA_OK(ctx->d_root_index.ensure(n));
    CUDA_OK(ctx->d_root_box.ensure(6 * static_cast<size_t>(n)));
    CUDA_OK(cudaMemcpyAsync(ctx->d_root_index.p, roots.data(), n * sizeof(int), cudaMemcpyHostToDevice, nullptr));
    k_gather_roots<<<(n + 255) / 256, 256>>>(ctx->d_nodes.p, ctx->d_root_index.p, n, ctx->d_root_box.p);
    CUDA_OK(cudaGetLastError());
    std::vector<double> boxes(6 * static_cast<size_t>(n));
    CUDA_OK(cudaMemcpy(boxes.data(), ctx->d_root_box.p, boxes.size() * sizeof(double), cudaMemcpyDeviceToHost));
    for (int i = 0; i < n; ++i) {
      HostAsset& a = *ctx->assets[ids[i]];
      for (int d = 0; d < 3; ++d) {
        a.nodes[1].c[d] = boxes[6 * i + d];
        a.nodes[1].h[d] = boxes[6 * i + 3 + d];
      }
    }
  }
  ++ctx->bounds_version;
  return DBH_OK;
}

// UploadPools + the re-derivation a pool regrowth owes to the assets re-posed since their registration.
static dbh_status Upload(dbh_context* ctx) {
  if (dbh_status s = ctx->UploadPools()) return s;
  if (ctx->regrown_deformed.empty()) return DBH_OK;
  const std::vector<int> ids = std::move(ctx->regrown_deformed);
  ctx->regrown_deformed.clear();
  return RefitAssets(ctx, ids);
}

// The two entry points differ in where the new positions are: host memory (the host copies follow) or this device (the
// host copies are refreshed lazily, before a pool regrowth).
static dbh_status UpdateVertices(dbh_context* ctx, int32_t num_assets, const int32_t* assets, const double* vertices,
                                 bool on_device) {
  if (!ctx || num_assets <= 0 || !assets || !vertices)
    return Fail(DBH_ERR_INVALID_ARGUMENT, "update_vertices: bad argument");
  CUDA_OK(cudaSetDevice(ctx->device));
  for (int k = 0; k < num_assets; ++k) {
    if (assets[k] < 0 || assets[k] >= static_cast<int>(ctx->assets.size()))
      return Fail(DBH_ERR_INVALID_ARGUMENT, "update_vertices: bad asset handle");
    if (ctx->assets[assets[k]]->half_space)
      return Fail(DBH_ERR_INVALID_ARGUMENT, "update_vertices: a half space has no vertices");
  }
  if (dbh_status s = Upload(ctx)) return s;
  // Runs of assets whose vertex ranges follow one another go into the pool in one copy.
  const double* src = vertices;
  int k = 0;
  while (k < num_assets) {
    const HostAsset& first = *ctx->assets[assets[k]];
    const double* run_src = src;
    long long nverts = 0;
    int k1 = k;
    while (k1 < num_assets) {
      HostAsset& a = *ctx->assets[assets[k1]];
      if (k1 > k && a.vert_base != first.vert_base + nverts) break;
      if (!on_device) std::copy(src, src + 3ll * a.num_vertices, a.rverts.begin());
      a.deformed = true;
      a.host_verts_stale = on_device;
      src += 3ll * a.num_vertices;
      nverts += a.num_vertices;
      ++k1;
    }
    CUDA_OK(cudaMemcpyAsync(ctx->d_rverts.p + 3ll * first.vert_base, run_src, 3ll * nverts * sizeof(double),
                            on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, nullptr));
    k = k1;
  }
  return RefitAssets(ctx, std::vector<int>(assets, assets + num_assets));
}

dbh_status dbh_context_update_vertices(dbh_context* ctx, int32_t num_assets, const int32_t* assets, const double* vertices) {
  return Guarded([&]() -> dbh_status { return UpdateVertices(ctx, num_assets, assets, vertices, false); });
}

dbh_status dbh_context_update_vertices_device(dbh_context* ctx, int32_t num_assets, const int32_t* assets,
                                              const double* device_vertices) {
  return Guarded([&]() -> dbh_status { return UpdateVertices(ctx, num_assets, assets, device_vertices, true); });
}

dbh_status dbh_register_half_space(dbh_context* ctx, int32_t type, double pressure_scale, double margin,
                                   int32_t* asset_out) {
  return Guarded([&]() -> dbh_status {
  if (!ctx || !asset_out || (type != DBH_RIGID && type != DBH_COMPLIANT))
    return Fail(DBH_ERR_INVALID_ARGUMENT, "register_half_space: bad argument");
  // PositiveDouble / NonNegativeDouble validators of MakeSoftRepresentation(HalfSpace) (hydroelastic_internal.cc:539-553).
  if (type == DBH_COMPLIANT && (!(pressure_scale > 0) || !(margin >= 0) || std::isinf(pressure_scale) || std::isinf(margin)))
    return Fail(DBH_ERR_INVALID_ARGUMENT,
                "register_half_space: a compliant half space needs a finite pressure_scale > 0 and a finite margin >= 0");
  auto a = std::make_unique<HostAsset>();
  a->type = type;
  a->half_space = true;
  a->pressure_scale = type == DBH_COMPLIANT ? pressure_scale : 0.0;
  a->margin = type == DBH_COMPLIANT ? margin : 0.0;
  ctx->Place(a.get());
  ctx->assets.push_back(std::move(a));
  *asset_out = static_cast<int32_t>(ctx->assets.size()) - 1;
  return DBH_OK;
  });
}

dbh_status dbh_asset_get_info(dbh_context* ctx, int32_t asset, dbh_asset_info* out) {
  return Guarded([&]() -> dbh_status {
  if (!ctx || !out || asset < 0 || asset >= static_cast<int>(ctx->assets.size()))
    return Fail(DBH_ERR_INVALID_ARGUMENT, "asset_get_info: bad argument");
  const HostAsset& a = *ctx->assets[asset];
  std::memset(out, 0, sizeof(*out));
  out->type = a.type;
  out->num_vertices = a.num_vertices;
  out->num_elements = a.num_elements;
  if (a.half_space) {
    out->is_half_space = 1;
    return DBH_OK;
  }
  out->num_bvh_nodes = a.vol_nodes - 1;
  out->bvh_height = a.height;
  for (int k = 0; k < 3; ++k) {
    out->root_center[k] = a.nodes[1].c[k];
    out->root_half_width[k] = a.nodes[1].h[k];
  }
  return DBH_OK;
  });
}

dbh_status dbh_asset_get_derived(dbh_context* ctx, int32_t asset, double* g_or_n, double* at_Mo, double* inward) {
  return Guarded([&]() -> dbh_status {
  if (!ctx || asset < 0 || asset >= static_cast<int>(ctx->assets.size()))
    return Fail(DBH_ERR_INVALID_ARGUMENT, "asset_get_derived: bad argument");
  CUDA_OK(cudaSetDevice(ctx->device));
  HostAsset& a = *ctx->assets[asset];
  if (dbh_status st = Upload(ctx)) return st;
  // Read back from the device pools so tests see what the kernels see.
  if (a.type == DBH_COMPLIANT) {
    std::vector<TetRec> t(a.num_elements);
    std::vector<TetAux> n(a.num_elements);
    CUDA_OK(cudaMemcpy(t.data(), ctx->d_tets.p + a.elem_base, t.size() * sizeof(TetRec), cudaMemcpyDeviceToHost));
    CUDA_OK(cudaMemcpy(n.data(), ctx->d_aux.p + a.elem_base, n.size() * sizeof(TetAux), cudaMemcpyDeviceToHost));
    for (int e = 0; e < a.num_elements; ++e) {
      if (g_or_n) std::memcpy(g_or_n + 3 * e, t[e].grad, 24);
      if (at_Mo) at_Mo[e] = t[e].p_mo;
      if (inward) std::memcpy(inward + 12 * e, n[e].inward, 96);
    }
  } else {
    std::vector<TriRec> t(a.num_elements);
    CUDA_OK(cudaMemcpy(t.data(), ctx->d_tris.p + a.elem_base, t.size() * sizeof(TriRec), cudaMemcpyDeviceToHost));
    for (int e = 0; e < a.num_elements; ++e)
      if (g_or_n) std::memcpy(g_or_n + 3 * e, t[e].n, 24);
  }
  return DBH_OK;
  });
}

dbh_status dbh_batch_create(dbh_context* ctx, int32_t num_scenes, const int32_t* scene_offsets,
                            const int32_t* geom_asset, const int64_t* geom_id, const uint8_t* geom_anchored,
                            int32_t num_excluded, const int32_t* excluded_pairs, dbh_batch** out) {
  return Guarded([&]() -> dbh_status {
  if (!ctx || !out || num_scenes <= 0 || !scene_offsets || !geom_asset || !geom_id)
    return Fail(DBH_ERR_INVALID_ARGUMENT, "batch_create: bad argument");
  *out = nullptr;
  if (scene_offsets[0] != 0) return Fail(DBH_ERR_INVALID_ARGUMENT, "batch_create: scene_offsets[0] must be 0");
  for (int s = 0; s < num_scenes; ++s)
    if (scene_offsets[s + 1] < scene_offsets[s])
      return Fail(DBH_ERR_INVALID_ARGUMENT, "batch_create: scene_offsets must be non-decreasing");
  if (num_excluded < 0 || (num_excluded > 0 && !excluded_pairs))
    return Fail(DBH_ERR_INVALID_ARGUMENT, "batch_create: excluded_pairs is null but num_excluded > 0");
  for (int i = 0; i < 2 * num_excluded; ++i)
    if (excluded_pairs[i] < 0 || excluded_pairs[i] >= scene_offsets[num_scenes])
      return Fail(DBH_ERR_INVALID_ARGUMENT, "batch_create: excluded pair index out of range");
  CUDA_OK(cudaSetDevice(ctx->device));
  auto b = std::make_unique<dbh_batch>();
  b->ctx = ctx;
  b->surface_bfs = ctx->soft_soft_intersector == DBH_INTERSECTOR_SURFACE_BFS;
  b->num_scenes = num_scenes;
  b->num_geoms = scene_offsets[num_scenes];
  b->scene_offsets.assign(scene_offsets, scene_offsets + num_scenes + 1);
  b->geom_asset.assign(geom_asset, geom_asset + b->num_geoms);
  b->geom_id.assign(geom_id, geom_id + b->num_geoms);
  const int num_assets = static_cast<int>(ctx->assets.size());
  for (int g = 0; g < b->num_geoms; ++g)
    if (geom_asset[g] < -1 || geom_asset[g] >= num_assets)
      return Fail(DBH_ERR_INVALID_ARGUMENT, "batch_create: bad asset handle");
  // Excluded pairs as a sorted list of (min, max) flat indices.
  std::vector<std::pair<int, int>> excl;
  for (int i = 0; i < num_excluded; ++i) {
    int x = excluded_pairs[2 * i], y = excluded_pairs[2 * i + 1];
    if (x > y) std::swap(x, y);
    excl.emplace_back(x, y);
  }
  std::sort(excl.begin(), excl.end());
  // Static geometry-pair list in (scene, id_M, id_N) order.
  int max_levels = 1;
  long long est_tri = 0, est_tet = 0, est_plane = 0;
  b->scene_first_pair.assign(num_scenes + 1, 0);
  b->scene_small_before.assign(num_scenes + 1, 0);
  b->scene_mesh_before.assign(num_scenes + 1, 0);
  for (int s = 0; s < num_scenes; ++s) {
    b->scene_first_pair[s] = static_cast<int>(b->pairs.size());
    b->scene_small_before[s] = b->num_small_pairs;
    b->scene_mesh_before[s] = b->info.num_soft_rigid_pairs + b->info.num_soft_soft_pairs;
    const int g0 = scene_offsets[s], g1 = scene_offsets[s + 1];
    std::vector<int> order(g1 - g0);
    for (int i = 0; i < g1 - g0; ++i) order[i] = g0 + i;
    std::sort(order.begin(), order.end(), [&](int x, int y) { return geom_id[x] < geom_id[y]; });
    for (size_t i = 0; i < order.size(); ++i) {
      for (size_t j = i + 1; j < order.size(); ++j) {
        const int ga = order[i], gb = order[j];  // id(ga) < id(gb)
        if (geom_id[ga] == geom_id[gb]) return Fail(DBH_ERR_INVALID_ARGUMENT, "duplicate geometry id in a scene");
        if (geom_anchored && geom_anchored[ga] && geom_anchored[gb]) continue;
        if (std::binary_search(excl.begin(), excl.end(), std::make_pair(std::min(ga, gb), std::max(ga, gb)))) continue;
        PairRec pr{};
        pr.scene = s;
        pr.id_m = geom_id[ga];
        pr.id_n = geom_id[gb];
        const int ta = geom_asset[ga] < 0 ? DBH_UNDEFINED : ctx->assets[geom_asset[ga]]->type;
        const int tb = geom_asset[gb] < 0 ? DBH_UNDEFINED : ctx->assets[geom_asset[gb]]->type;
        pr.geom_a = ga;
        pr.geom_b = gb;
        const bool ha = ta != DBH_UNDEFINED && ctx->assets[geom_asset[ga]]->half_space;
        const bool hb = tb != DBH_UNDEFINED && ctx->assets[geom_asset[gb]]->half_space;
        if (ta == DBH_UNDEFINED || tb == DBH_UNDEFINED) {
          pr.kind = kUnsupported;
          ++b->info.num_unsupported_pairs;
        } else if (ta == DBH_RIGID && tb == DBH_RIGID) {
          pr.kind = (ha || hb) ? kRigidRigidPlane : kRigidRigid;
          ++b->info.num_rigid_rigid_pairs;
        } else if (ha && hb) {
          pr.kind = kHalfSpaceHalfSpace;  // hydroelastic_calculator.cc:113-118,145-147
          ++b->info.num_rejected_half_space_pairs;
        } else if (ta == DBH_COMPLIANT && tb == DBH_COMPLIANT) {
          if (ha || hb) {
            pr.kind = kSoftPlaneSoftMesh;  // kCompliantHalfSpaceCompliantMesh (:120-123)
            ++b->info.num_rejected_half_space_pairs;
          } else {
            pr.kind = kSoftSoft;  // smaller id is M / field0 (hydroelastic_calculator.cc:109-111)
            ++b->info.num_soft_soft_pairs;
          }
        } else {
          // Rigid-compliant, built as (S, R), swapped when id_R < id_S; CalcRigidCompliant picks the
          // half-space special cases by shape (hydroelastic_calculator.cc:25-41).
          if (ta == DBH_RIGID) {
            pr.geom_a = gb;
            pr.geom_b = ga;
            pr.swapped = 1;
          }
          const bool soft_is_half_space = ta == DBH_RIGID ? hb : ha;
          if (ha || hb) {
            pr.kind = soft_is_half_space ? kSoftPlaneRigidMesh : kSoftMeshRigidPlane;
            ++b->info.num_half_space_pairs;
          } else {
            pr.kind = kSoftRigid;
            ++b->info.num_soft_rigid_pairs;
          }
        }
        if (pr.kind == kUnsupported || pr.kind == kRigidRigid || pr.kind >= kHalfSpaceHalfSpace) ++b->num_no_surface_pairs;
        pr.root_a = pr.root_b = -1;
        pr.elem_base_a = pr.elem_base_b = -1;
        if (pr.kind <= kRigidRigid) {
          const HostAsset& A = *ctx->assets[geom_asset[pr.geom_a]];
          const HostAsset& B = *ctx->assets[geom_asset[pr.geom_b]];
          pr.root_a = A.node_base + 1;
          pr.root_b = B.node_base + 1;
          pr.elem_base_a = A.elem_base;
          pr.elem_base_b = B.elem_base;
          pr.root_leaf = (A.nodes[1].left < 0 ? 1 : 0) | (B.nodes[1].left < 0 ? 2 : 0);
          int nodes_a = A.vol_nodes, nodes_b = B.vol_nodes, height = std::max(A.height, B.height);
          if (pr.kind == kSoftSoft && b->surface_bfs) {
            // The surface-seeded intersector: the mid phase descends the two BOUNDARY trees (hydro_bfs.cuh); node 0 of
            // a boundary tree is the slot before its root.
            if (!A.bfs_nodes || !B.bfs_nodes)
              return Fail(DBH_ERR_INVALID_ARGUMENT,
                          "the surface intersector needs compliant meshes registered after "
                          "dbh_context_set_soft_soft_intersector(DBH_INTERSECTOR_SURFACE_BFS)");
            if (A.num_elements > (1 << kBfsTetBits) || B.num_elements > (1 << kBfsTetBits))
              return Fail(DBH_ERR_INVALID_ARGUMENT, "the surface intersector handles at most 2^21 tetrahedra per mesh");
            pr.root_a = A.node_base + A.vol_nodes;
            pr.root_b = B.node_base + B.vol_nodes;
            pr.root_leaf = (A.nodes[A.vol_nodes].left < 0 ? 1 : 0) | (B.nodes[B.vol_nodes].left < 0 ? 2 : 0);
            nodes_a = A.bfs_nodes;
            nodes_b = B.bfs_nodes;
            height = std::max(A.bfs_height, B.bfs_height);
          }
          if (pr.kind <= kSoftSoft && nodes_a <= kSmallMaxNodes && nodes_b <= kSmallMaxNodes) {
            pr.small = 1;
            ++b->num_small_pairs;
            if (nodes_a + nodes_b > kSmallSplitNodes) {
              pr.small = 2;
              ++b->num_large_small_pairs;
            }
          }
          if (pr.kind <= kSoftSoft) {
            max_levels = std::max(max_levels, height + 1);
            // First guess of the list lengths (they grow on demand): proportional to the meshes' volume for the
            // hierarchy descent, to their surface for the surface-seeded intersector.
            const bool surface = pr.kind == kSoftSoft && b->surface_bfs;
            (pr.kind == kSoftRigid ? est_tri : est_tet) +=
                surface ? 2ll * (nodes_a + nodes_b) : 2ll * (A.num_elements + B.num_elements);
          }
        } else if (pr.kind == kSoftMeshRigidPlane || pr.kind == kSoftPlaneRigidMesh) {
          // The one mesh of the pair: the compliant volume (a) or the rigid surface (b).
          const HostAsset& M = *ctx->assets[geom_asset[pr.kind == kSoftMeshRigidPlane ? pr.geom_a : pr.geom_b]];
          (pr.kind == kSoftMeshRigidPlane ? pr.root_a : pr.root_b) = M.node_base + 1;
          (pr.kind == kSoftMeshRigidPlane ? pr.elem_base_a : pr.elem_base_b) = M.elem_base;
          // Work items are padded to whole blocks of kPlaneBlock leaves, so that one 128-thread block of the count
          // pass is one (pair, leaf block) and can cull it as a whole.
          b->plane_pairs.push_back(static_cast<int>(b->pairs.size()));
          b->plane_elem_off.push_back(est_plane);
          b->plane_nelem.push_back(M.num_elements);
          b->plane_block_base.push_back(M.block_base);
          est_plane += static_cast<long long>((M.num_elements + kPlaneBlock - 1) / kPlaneBlock) * kPlaneBlock;
        }
        b->pairs.push_back(pr);
      }
    }
  }
  b->num_pairs = static_cast<int>(b->pairs.size());
  if (b->surface_bfs && b->num_pairs > (1 << kBfsPairBits) - 2)
    return Fail(DBH_ERR_INVALID_ARGUMENT, "the surface intersector handles at most 2^22 - 2 geometry pairs per batch");
  b->scene_first_pair[num_scenes] = b->num_pairs;
  b->scene_small_before[num_scenes] = b->num_small_pairs;
  b->scene_mesh_before[num_scenes] = b->info.num_soft_rigid_pairs + b->info.num_soft_soft_pairs;
  b->plane_elem_off.push_back(est_plane);
  b->info.num_scenes = num_scenes;
  b->info.num_geometries = b->num_geoms;
  b->info.num_geometry_pairs = b->num_pairs;
  b->info.max_traversal_levels = max_levels;
  if (max_levels > 62) return Fail(DBH_ERR_INVALID_ARGUMENT, "BVH too deep");

  // Default capacities: a few entries per element of every pair, bounded (a
  // scene with many geometries has mostly non-overlapping pairs); the first
  // queries of a batch grow exactly the lists that overflow, or split the batch
  // into more chunks of scenes when the lists would outgrow the memory budget
  // (dbh_batch_grow_capacity); steady state never reallocates.
  {
    size_t free_b = 0, total_b = 0;
    CUDA_OK(cudaMemGetInfo(&free_b, &total_b));
    b->work_budget_bytes = static_cast<size_t>(0.35 * static_cast<double>(total_b));
    if (const char* e = std::getenv("DBH_WORK_BUDGET_MB")) b->work_budget_bytes = static_cast<size_t>(std::atoll(e)) << 20;
  }
  auto clampest = [](long long e) { return std::min(std::max(e, 1ll << 16), 1ll << 23); };
  for (;;) {
    const long long est = clampest((est_tri + est_tet) / b->num_chunks);
    const long long min_list = b->MaxPairsPerChunk();
    b->cap_frontier = std::max<long long>(est, min_list);
    b->cap_tri = std::max<long long>(clampest(est_tri / b->num_chunks), min_list);
    b->cap_tet = std::max<long long>(clampest(est_tet / b->num_chunks), min_list);
    b->cap_polys = est / 4;
    if (b->WorkBytes(b->cap_frontier, b->cap_tri, b->cap_tet, b->cap_polys) <= b->work_budget_bytes ||
        b->num_chunks >= num_scenes)
      break;
    b->num_chunks = std::min(num_scenes, b->num_chunks * 2);
  }
  {
    const long long est = clampest(est_tri + est_tet + est_plane);
    b->cap_faces = est / 2;
    b->cap_verts = est;
  }

  CUDA_OK(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking));
  CUDA_OK(cudaMallocHost(&b->h_counters, sizeof(Counters)));
  std::memset(b->h_counters, 0, sizeof(Counters));
  CUDA_OK(cudaEventCreate(&b->ev_begin));
  CUDA_OK(cudaEventCreate(&b->ev_end));
  const size_t np = std::max(1, b->num_pairs);
  CUDA_OK(b->d_pairs.alloc(np));
  if (b->num_pairs)
    CUDA_OK(cudaMemcpy(b->d_pairs.p, b->pairs.data(), b->num_pairs * sizeof(PairRec), cudaMemcpyHostToDevice));
  // Bounds: a mesh's root box, a half space, or "everywhere" for a geometry with no hydroelastic representation
  // (until dbh_batch_set_geometry_bounds says otherwise).
  b->geom_bounds.assign(std::max(1, b->num_geoms), GeomBound{});
  for (int g = 0; g < b->num_geoms; ++g) {
    GeomBound& gb = b->geom_bounds[g];
    gb.kind = kBoundEverywhere;
    if (geom_asset[g] < 0) continue;
    const HostAsset& A = *ctx->assets[geom_asset[g]];
    if (A.half_space) {
      gb.kind = kBoundHalfSpace;
    } else {
      gb.kind = kBoundBox;
      for (int k = 0; k < 3; ++k) {
        gb.c[k] = A.nodes[1].c[k];
        gb.h[k] = A.nodes[1].h[k];
      }
    }
  }
  CUDA_OK(b->d_geom_bounds.alloc(b->geom_bounds.size()));
  CUDA_OK(cudaMemcpy(b->d_geom_bounds.p, b->geom_bounds.data(), b->geom_bounds.size() * sizeof(GeomBound),
                     cudaMemcpyHostToDevice));
  CUDA_OK(b->d_fallback.alloc(std::max(1, b->num_no_surface_pairs)));
  CUDA_OK(b->d_poses_own.alloc(std::max(1, b->num_geoms)));
  b->d_poses = b->d_poses_own.p;
  CUDA_OK(b->d_pair_X.alloc(np));
  CUDA_OK(b->d_pair_f32.alloc(np));
  CUDA_OK(b->d_counters.alloc(1));
  CUDA_OK(b->d_pair_polys.alloc(np));
  CUDA_OK(b->d_pair_verts.alloc(np));
  CUDA_OK(b->d_pair_cursor.alloc(np));
  CUDA_OK(b->d_surf_index.alloc(np + 1));
  CUDA_OK(b->d_surf_pair.alloc(np));
  CUDA_OK(b->d_big_list.alloc(np));
  CUDA_OK(b->d_poly_off.alloc(np + 1));
  CUDA_OK(b->d_vert_off.alloc(np + 1));
  CUDA_OK(b->d_surf_face_off.alloc(np + 1));
  CUDA_OK(b->d_surf_vert_off.alloc(np + 1));
  CUDA_OK(b->r_surf_total_area.alloc(np));
  CUDA_OK(b->r_surf_centroid.alloc(np * 3));
  CUDA_OK(b->r_surf_scene.alloc(np));
  CUDA_OK(b->r_surf_id_m.alloc(np));
  CUDA_OK(b->r_surf_id_n.alloc(np));
  CUDA_OK(b->r_surf_flags.alloc(np));
  const int scan_blocks = (b->num_pairs + kScanTile - 1) / kScanTile;
  CUDA_OK(b->d_block_sums.alloc(3 * std::max(1, scan_blocks)));
  if (!b->plane_pairs.empty()) {
    CUDA_OK(b->d_plane_pairs.alloc(b->plane_pairs.size()));
    CUDA_OK(cudaMemcpy(b->d_plane_pairs.p, b->plane_pairs.data(), b->plane_pairs.size() * sizeof(int),
                       cudaMemcpyHostToDevice));
    CUDA_OK(b->d_plane_nelem.alloc(b->plane_nelem.size()));
    CUDA_OK(cudaMemcpy(b->d_plane_nelem.p, b->plane_nelem.data(), b->plane_nelem.size() * sizeof(int), cudaMemcpyHostToDevice));
    CUDA_OK(b->d_plane_block_base.alloc(b->plane_block_base.size()));
    CUDA_OK(cudaMemcpy(b->d_plane_block_base.p, b->plane_block_base.data(), b->plane_block_base.size() * sizeof(int),
                       cudaMemcpyHostToDevice));
    CUDA_OK(b->d_plane_elem_off.alloc(b->plane_elem_off.size()));
    CUDA_OK(cudaMemcpy(b->d_plane_elem_off.p, b->plane_elem_off.data(), b->plane_elem_off.size() * sizeof(long long),
                       cudaMemcpyHostToDevice));
    std::vector<double> hs(2 * static_cast<size_t>(b->num_geoms), 0.0);
    for (int g = 0; g < b->num_geoms; ++g)
      if (geom_asset[g] >= 0) {
        hs[2 * g] = ctx->assets[geom_asset[g]]->pressure_scale;
        hs[2 * g + 1] = ctx->assets[geom_asset[g]]->margin;
      }
    CUDA_OK(b->d_geom_hs.alloc(hs.size()));
    CUDA_OK(cudaMemcpy(b->d_geom_hs.p, hs.data(), hs.size() * sizeof(double), cudaMemcpyHostToDevice));
  }
  *out = b.release();
  return DBH_OK;
  });
}

void dbh_batch_destroy(dbh_batch* batch) { delete batch; }

dbh_status dbh_batch_set_capacity(dbh_batch* b, int64_t max_bv_pairs, int64_t max_element_pairs, int64_t max_faces) {
  return Guarded([&]() -> dbh_status {
  if (!b) return Fail(DBH_ERR_INVALID_ARGUMENT, "batch is null");
  if (max_bv_pairs > 0) b->cap_frontier = max_bv_pairs;
  if (max_element_pairs > 0) b->cap_tri = b->cap_tet = max_element_pairs;
  if (max_faces > 0) {
    b->cap_polys = max_faces;
    b->cap_faces = max_faces * 4;   // triangle rep: up to ~5 triangles per polygon on average
    b->cap_verts = max_faces * 6;
  }
  return DBH_OK;
  });
}

dbh_status dbh_batch_set_chunks(dbh_batch* b, int32_t num_chunks) {
  return Guarded([&]() -> dbh_status {
  if (!b || num_chunks < 1) return Fail(DBH_ERR_INVALID_ARGUMENT, "set_chunks: bad argument");
  b->num_chunks = std::min<int>(num_chunks, b->num_scenes);
  // The lists must hold at least one entry per geometry pair of a chunk (k_pairs writes them unchecked).
  const long long min_list = b->MaxPairsPerChunk();
  b->cap_frontier = std::max(b->cap_frontier, min_list);
  b->cap_tri = std::max(b->cap_tri, min_list);
  b->cap_tet = std::max(b->cap_tet, min_list);
  return DBH_OK;
  });
}

int32_t dbh_batch_get_chunks(dbh_batch* b) { return b ? b->num_chunks : 0; }
int32_t dbh_batch_last_launch_count(dbh_batch* b) { return b ? b->last_launches : 0; }

dbh_status dbh_batch_grow_capacity(dbh_batch* b) {
  return Guarded([&]() -> dbh_status {
  if (!b) return Fail(DBH_ERR_INVALID_ARGUMENT, "batch is null");
  // The device counters keep counting past a full list, so after a failed
  // query they say how many entries the failing stage wanted (the maximum over
  // the chunks that ran). Later stages did not run; the caller re-runs until
  // the query fits (a handful of set-up iterations, each at least doubling
  // what overflowed -- or halving the chunks when the lists would outgrow the
  // memory budget).
  const Counters& c = *b->h_counters;
  if (c.error == DBH_ERR_CAPACITY && c.error_detail == 9) {
    // A small pair's node-pair stack outgrew its shared-memory slot: double it while four warps still fit a
    // block's shared memory, else hand the batch's pairs to the generic pipeline.
    if (b->small_stack_cap < 8192) b->small_stack_cap *= 2;
    else b->small_mode = 0;
  }
  auto want = [](unsigned long long needed, long long cap) {
    return static_cast<long long>(needed) > cap ? static_cast<long long>(needed) * 2 : cap;
  };
  const long long fr = want(c.frontier_max, b->cap_frontier);
  const long long tri = want(c.leaf_tri_max, b->cap_tri);
  const long long tet = want(c.leaf_tet_max, b->cap_tet);
  const long long polys = want(c.polys_max, b->cap_polys);
  const bool lists_grow = fr != b->cap_frontier || tri != b->cap_tri || tet != b->cap_tet || polys != b->cap_polys;
  if (lists_grow && b->WorkBytes(fr, tri, tet, polys) > b->work_budget_bytes && b->num_chunks < b->num_scenes) {
    b->num_chunks = std::min(b->num_scenes, b->num_chunks * 2);
  } else {
    b->cap_frontier = fr;
    b->cap_tri = tri;
    b->cap_tet = tet;
    b->cap_polys = polys;
  }
  b->cap_slow = want(c.slow_max, b->cap_slow);
  b->cap_undecided = want(c.undecided_max, b->cap_undecided);
  b->cap_small_items = want(c.small_items_max, b->cap_small_items);
  // Faces / vertices of either representation (triangle: one face per polygon vertex, one extra
  // vertex per polygon), so switching representation does not overflow again.
  b->cap_faces = want(std::max(c.faces, c.poly_verts), b->cap_faces);
  b->cap_verts = want(std::max(c.vertices, c.poly_verts + c.polygons), b->cap_verts);
  return DBH_OK;
  });
}

dbh_status dbh_batch_get_info(dbh_batch* b, dbh_batch_info* out) {
  return Guarded([&]() -> dbh_status {
  if (!b || !out) return Fail(DBH_ERR_INVALID_ARGUMENT, "batch_get_info: bad argument");
  *out = b->info;
  return DBH_OK;
  });
}

dbh_status dbh_batch_set_poses_host(dbh_batch* b, const double* X_WGs) {
  return Guarded([&]() -> dbh_status {
  if (!b || !X_WGs) return Fail(DBH_ERR_INVALID_ARGUMENT, "set_poses_host: bad argument");
  CUDA_OK(cudaSetDevice(b->ctx->device));
  b->d_poses = b->d_poses_own.p;
  CUDA_OK(cudaMemcpyAsync(b->d_poses_own.p, X_WGs, static_cast<size_t>(b->num_geoms) * sizeof(Xf),
                          cudaMemcpyHostToDevice, b->stream));
  return DBH_OK;
  });
}

dbh_status dbh_batch_set_poses_device(dbh_batch* b, const double* X_WGs_device) {
  return Guarded([&]() -> dbh_status {
  if (!b || !X_WGs_device) return Fail(DBH_ERR_INVALID_ARGUMENT, "set_poses_device: bad argument");
  b->d_poses = reinterpret_cast<const Xf*>(X_WGs_device);
  return DBH_OK;
  });
}

void* dbh_batch_stream(dbh_batch* b) { return b ? b->stream : nullptr; }

dbh_status dbh_batch_enable_timing(dbh_batch* b, int32_t enabled) {
  return Guarded([&]() -> dbh_status {
  if (!b) return Fail(DBH_ERR_INVALID_ARGUMENT, "batch is null");
  b->timing = enabled != 0;
  return DBH_OK;
  });
}

dbh_status dbh_batch_set_small_pair_path(dbh_batch* b, int32_t mode) {
  return Guarded([&]() -> dbh_status {
  if (!b || mode < 0 || mode > 2) return Fail(DBH_ERR_INVALID_ARGUMENT, "set_small_pair_path: bad argument");
  b->small_mode = mode;
  return DBH_OK;
  });
}

dbh_status dbh_batch_set_record_candidates(dbh_batch* b, int32_t enabled) {
  return Guarded([&]() -> dbh_status {
  if (!b) return Fail(DBH_ERR_INVALID_ARGUMENT, "batch is null");
  b->record_candidates = enabled != 0;
  return DBH_OK;
  });
}

dbh_status dbh_batch_set_normal_filter(dbh_batch* b, int32_t enabled) {
  if (!b) return Fail(DBH_ERR_INVALID_ARGUMENT, "batch is null");
  b->normal_filter = enabled != 0;
  return DBH_OK;
}

dbh_status dbh_batch_set_verify(dbh_batch* b, int32_t enabled) {
  return Guarded([&]() -> dbh_status {
  if (!b) return Fail(DBH_ERR_INVALID_ARGUMENT, "batch is null");
  b->verify = enabled != 0;
  return DBH_OK;
  });
}

dbh_status dbh_batch_query(dbh_batch* b, int32_t representation, int32_t query_kind, int64_t* scene_stats_dev) {
  return Guarded([&]() -> dbh_status {
  if (!b || (representation != DBH_TRIANGLE && representation != DBH_POLYGON))
    return Fail(DBH_ERR_INVALID_ARGUMENT, "batch_query: bad argument");
  dbh_context* ctx = b->ctx;
  CUDA_OK(cudaSetDevice(ctx->device));
  if (dbh_status s = Upload(ctx)) return s;
  if (dbh_status s = b->AllocWorkBuffers()) return s;
  if (b->bounds_version != ctx->bounds_version) {   // dbh_context_update_vertices moved root boxes
    for (int g = 0; g < b->num_geoms; ++g) {
      if (b->geom_asset[g] < 0 || b->geom_bounds[g].kind != kBoundBox) continue;
      const HostAsset& A = *ctx->assets[b->geom_asset[g]];
      for (int k = 0; k < 3; ++k) {
        b->geom_bounds[g].c[k] = A.nodes[1].c[k];
        b->geom_bounds[g].h[k] = A.nodes[1].h[k];
      }
    }
    CUDA_OK(cudaMemcpyAsync(b->d_geom_bounds.p, b->geom_bounds.data(), b->geom_bounds.size() * sizeof(GeomBound),
                            cudaMemcpyHostToDevice, b->stream));
    b->bounds_version = ctx->bounds_version;
  }
  b->representation = representation;
  b->query_done = false;
  b->query_enqueued = true;
  b->forces_done = false;
  b->discrete_done = false;
  b->deformable_done = false;
  b->last_launches = 0;
  b->last_query_kind = query_kind;
  cudaStream_t st = b->stream;
  const int sms = ctx->prop.multiProcessorCount;
  if (b->timing) {
    while (b->ev_chunk.size() < static_cast<size_t>(kStageEvents) * static_cast<size_t>(b->num_chunks)) {
      cudaEvent_t e = nullptr;
      CUDA_OK(cudaEventCreate(&e));
      b->ev_chunk.push_back(e);
    }
    cudaEventRecord(b->ev_begin, st);
  }
  b->timed_chunks = 0;
  auto mark = [&](int i) {
    if (b->timing) cudaEventRecord(b->ev_chunk[kStageEvents * b->timed_chunks + i], st);
  };
  CUDA_OK(cudaMemsetAsync(b->d_counters.p, 0, sizeof(Counters), st));
  if (scene_stats_dev)
    CUDA_OK(cudaMemsetAsync(scene_stats_dev, 0, static_cast<size_t>(b->num_scenes) * 6 * sizeof(int64_t), st));
  if (b->num_pairs > 0) {
    if (b->MaxPairsPerChunk() > std::min(b->cap_frontier, std::min(b->cap_tri, b->cap_tet)))
      return Fail(DBH_ERR_CAPACITY, "work-list capacity below the geometry-pair count of a chunk");
    PairsArgs pa{};
    pa.pairs = b->d_pairs.p;
    pa.pools = ctx->pools();
    pa.poses = b->d_poses;
    pa.pair_X = b->d_pair_X.p;
    pa.pair_f32 = b->d_pair_f32.p;
    pa.frontier0 = b->d_frontier[0].p;
    pa.leaf_tri = b->d_cand_tri.p;
    pa.leaf_tet = b->d_cand_tet.p;
    pa.counters = b->d_counters.p;
    pa.strict = query_kind == DBH_QUERY_STRICT;
    pa.pair_polys = b->d_pair_polys.p;
    pa.pair_verts = b->d_pair_verts.p;
    pa.pair_cursor = b->d_pair_cursor.p;
    pa.small_list = nullptr;
    pa.geom_bounds = b->d_geom_bounds.p;
    pa.fallback_list = b->d_fallback.p;
    pa.fallback_capacity = static_cast<unsigned long long>(b->num_no_surface_pairs);
    TraverseArgs ta{};
    ta.pair_f32 = b->d_pair_f32.p;
    ta.pools = ctx->pools();
    ta.leaf_tri = b->d_cand_tri.p;
    ta.leaf_tet = b->d_cand_tet.p;
    ta.leaf_capacity_tri = static_cast<unsigned long long>(b->cap_tri);
    ta.leaf_capacity_tet = static_cast<unsigned long long>(b->cap_tet);
    ta.out_capacity = static_cast<unsigned long long>(b->cap_frontier);
    ta.counters = b->d_counters.p;
    ClipArgs ca{};
    ca.pair_X = b->d_pair_X.p;
    ca.pair_f32 = b->d_pair_f32.p;
    ca.pools = ctx->pools();
    ca.soup.hdr = b->d_soup_hdr.p;
    ca.soup.v = b->d_soup_v.p;
    ca.soup.nhat = b->d_soup_nhat.p;
    ca.soup.capacity = static_cast<unsigned long long>(b->cap_polys);
    ca.pair_polys = b->d_pair_polys.p;
    ca.pair_verts = b->d_pair_verts.p;
    ca.counters = b->d_counters.p;
    ca.slow = b->d_slow.p;
    ca.slow_capacity = static_cast<unsigned int>(b->cap_slow);
    ca.undecided = b->d_undecided.p;
    ca.undecided_capacity = static_cast<unsigned long long>(b->cap_undecided);
    ca.verify = b->verify ? 1 : 0;
    ca.no_normal_filter = b->normal_filter ? 0 : 1;
    ca.trace = make_int4(-1, -1, -1, 0);
#ifdef DBH_TRACE
    if (const char* t = std::getenv("DBH_TRACE_PAIR")) {
      int p = -1, e0 = -1, e1 = -1;
      if (std::sscanf(t, "%d,%d,%d", &p, &e0, &e1) == 3 && p >= 0 && p < b->num_pairs)
        ca.trace = make_int4(p, e0 + b->pairs[p].elem_base_a, e1 + b->pairs[p].elem_base_b, 0);
    }
#endif
    SmallArgs sm{};
    sm.c = ca;
    sm.pairs = b->d_pairs.p;
    sm.surv_capacity = static_cast<unsigned long long>(b->cap_frontier);
    sm.record = b->record_candidates ? 1 : 0;
    sm.stack_cap = b->small_stack_cap;
    UndecidedArgs ua{};
    ua.c = ca;
    ua.pairs = b->d_pairs.p;
    ua.surv[0] = b->d_frontier[0].p;
    ua.surv[1] = b->d_frontier[1].p;
    ua.surv_capacity = static_cast<unsigned long long>(b->cap_frontier);
    ua.cand[0] = b->d_cand_tri.p;
    ua.cand[1] = b->d_cand_tet.p;
    b->small_used = false;
    ScanArgs sa{};
    sa.pair_polys = b->d_pair_polys.p;
    sa.pair_verts = b->d_pair_verts.p;
    sa.poly_off = b->d_poly_off.p;
    sa.vert_off = b->d_vert_off.p;
    sa.surf_index = b->d_surf_index.p;
    sa.block_sums = b->d_block_sums.p;
    sa.surf_pair = b->d_surf_pair.p;
    sa.counters = b->d_counters.p;
    PackArgs pk{};
    pk.soup = ca.soup;
    pk.pairs = b->d_pairs.p;
    pk.pools = ctx->pools();
    pk.poses = b->d_poses;
    pk.poly_off = b->d_poly_off.p;
    pk.vert_off = b->d_vert_off.p;
    pk.pair_cursor = b->d_pair_cursor.p;
    pk.representation = representation;
    pk.out.vertices = b->r_vertices.p;
    pk.out.pressures = b->r_pressures.p;
    pk.out.face_data = b->r_face_data.p;
    pk.out.areas = b->r_areas.p;
    pk.out.normals = b->r_normals.p;
    pk.out.centroids = b->r_centroids.p;
    pk.out.e_centroid = b->r_e_centroid.p;
    pk.out.grad_eM = b->r_grad_eM.p;
    pk.out.grad_eN = b->r_grad_eN.p;
    pk.out.elem_a = b->r_elem_a.p;
    pk.out.elem_b = b->r_elem_b.p;
    pk.out.face_pair = b->r_face_pair.p;
    pk.out.vertex_origin = b->plane_pairs.empty() ? nullptr : b->r_vertex_origin.p;
    pk.out.face_capacity = b->cap_faces;
    pk.out.vert_capacity = b->cap_verts;
    pk.out.face_data_capacity = b->cap_face_data;
    pk.counters = b->d_counters.p;
    SurfaceArgs su{};
    su.surf_index = b->d_surf_index.p;
    su.surf_pair = b->d_surf_pair.p;
    su.pairs = b->d_pairs.p;
    su.poses = b->d_poses;
    su.poly_off = b->d_poly_off.p;
    su.vert_off = b->d_vert_off.p;
    su.representation = representation;
    su.areas = b->r_areas.p;
    su.centroids = b->r_centroids.p;
    su.surf_face_off = b->d_surf_face_off.p;
    su.surf_vert_off = b->d_surf_vert_off.p;
    su.surf_total_area = b->r_surf_total_area.p;
    su.surf_centroid = b->r_surf_centroid.p;
    su.surf_scene = b->r_surf_scene.p;
    su.surf_id_m = b->r_surf_id_m.p;
    su.surf_id_n = b->r_surf_id_n.p;
    su.surf_flags = b->r_surf_flags.p;
    su.scene_stats = reinterpret_cast<long long*>(scene_stats_dev);
    su.big_list = b->d_big_list.p;
    su.counters = b->d_counters.p;
    PlaneArgs pl{};
    pl.pairs = b->d_pairs.p;
    pl.poses = b->d_poses;
    pl.pair_X = b->d_pair_X.p;
    pl.pools = ctx->pools();
    pl.tet_vp = ctx->d_tet_vp.p;
    pl.plane_pairs = b->d_plane_pairs.p;
    pl.plane_elem_off = b->d_plane_elem_off.p;
    pl.plane_nelem = b->d_plane_nelem.p;
    pl.plane_block_base = b->d_plane_block_base.p;
    pl.order_tet = ctx->d_order_tet.p;
    pl.order_tri = ctx->d_order_tri.p;
    pl.blocks = ctx->d_plane_blocks.p;
    pl.geom_hs = b->d_geom_hs.p;
    pl.emit_list = b->d_plane_emit.p;
    pl.pair_polys = b->d_pair_polys.p;
    pl.pair_verts = b->d_pair_verts.p;
    pl.pair_cursor = b->d_pair_cursor.p;
    pl.poly_off = b->d_poly_off.p;
    pl.vert_off = b->d_vert_off.p;
    pl.representation = representation;
    pl.out = pk.out;
    pl.counters = b->d_counters.p;
    // Launch shapes are fixed by measurement (profiles/README.md): block sizes and resident blocks per SM are template
    // arguments of the kernels. What stays adjustable for experiments (DBH_TUNE="tgrid=2,small=0,..."): grid sizes in units
    // of the resident blocks (persistent grids beat over-subscribed ones by 5-15 %: no tail wave), the mid-phase path,
    // the stage-A ablation bits and the work-item split of the large small pairs.
    struct Tune { int tgrid = 3, fgrid = 1, pgrid = 1, small = -1, stagea = 0, split = 1, split_at = 32; } tune;
    if (const char* t = std::getenv("DBH_TUNE")) {
      const std::string str(t);
      auto get = [&](const char* key, int* v) {
        const size_t p = str.find(std::string(key) + "=");
        if (p != std::string::npos) *v = std::atoi(str.c_str() + p + std::strlen(key) + 1);
      };
      get("tgrid", &tune.tgrid);
      get("fgrid", &tune.fgrid);
      get("pgrid", &tune.pgrid);
      get("small", &tune.small);  // overrides dbh_batch_set_small_pair_path: 0 never, 1 always, 2 auto
      get("stagea", &tune.stagea);    // experiments on the conservative culls (ClipArgs::tune)
      get("split_at", &tune.split_at);
      get("split", &tune.split);      // 0: large small pairs are not split into work items
    }
    ca.tune = tune.stagea;
    sm.c.tune = tune.stagea;
    ua.c.tune = tune.stagea;
    const int small_mode = tune.small >= 0 ? tune.small : b->small_mode;
    const int small_smem = kSmallWarps * small_smem_bytes_per_warp(b->small_stack_cap);
    if (b->num_small_pairs && small_mode != 0 && small_smem > 48 * 1024) {
      for (int tet = 0; tet < 2; ++tet)
        for (int pass = 0; pass < 3; ++pass)
          CUDA_OK(cudaFuncSetAttribute(SmallKernel(tet != 0, pass), cudaFuncAttributeMaxDynamicSharedMemorySize, small_smem));
    }
    for (int chunk = 0; chunk < b->num_chunks; ++chunk) {
      const int s0 = static_cast<int>(static_cast<long long>(b->num_scenes) * chunk / b->num_chunks);
      const int s1 = static_cast<int>(static_cast<long long>(b->num_scenes) * (chunk + 1) / b->num_chunks);
      const int p0 = b->scene_first_pair[s0], p1 = b->scene_first_pair[s1];
      if (p1 == p0) continue;
      pa.pair_begin = p0;
      pa.pair_end = p1;
      // Small pairs take the on-chip path when the chunk has enough of them to fill the machine (one warp per pair).
      const int n_small = b->scene_small_before[s1] - b->scene_small_before[s0];
      const int n_mesh = b->scene_mesh_before[s1] - b->scene_mesh_before[s0];
      const bool use_small =
          n_small > 0 && (small_mode == 1 || (small_mode == 2 && n_small >= 2 * sms * 4 * kSmallWarps));
      const bool generic = n_mesh - (use_small ? n_small : 0) > 0;   // pairs left to the per-level pipeline
      pa.small_list = use_small ? b->d_small_list.p : nullptr;
      pa.small_list_tet = use_small ? b->d_small_list.p + b->MaxPairsPerChunk() : nullptr;
      mark(0);
      ++b->last_launches;
      k_pairs<<<(p1 - p0 + 127) / 128, 128, 0, st>>>(pa);
      mark(1);
      // Dual-BVH traversal, one launch per level; counts stay on the device.
      for (int level = 0; generic && level < b->info.max_traversal_levels; ++level) {
        ta.in = b->d_frontier[level & 1].p;
        ta.out = b->d_frontier[(level + 1) & 1].p;
        ta.level = level;
        ++b->last_launches;
        k_traverse<6, 128><<<sms * tune.tgrid * 2, 128, 0, st>>>(ta);
      }
      mark(2);
      // Filter (leaf predicate + conservative culls) then the exact stage, per element kind. The
      // survivor lists reuse the two frontier buffers.
      ca.surv_capacity = static_cast<unsigned long long>(b->cap_frontier);
      // The generic pairs' leaf lists go through the filter first (the survivor lists reuse the frontier buffers,
      // which the descent has released); then the small pairs are descended and culled on chip and append their
      // survivors to the same lists; then the exact stage runs over both.
      b->small_used = b->small_used || use_small;
      // After both kinds' small pairs: the leaf pairs their fp32 leaf predicate left undecided (one short launch).
      auto settle_undecided = [&]() {
        ++b->last_launches;
        k_leaf_undecided<<<sms * 8, 128, 0, st>>>(ua);
      };
      for (int tet = 0; tet < 2; ++tet) {
        if (!(tet ? b->info.num_soft_soft_pairs : b->info.num_soft_rigid_pairs)) {
          if (tet) settle_undecided();
          mark(3 + tet);
          continue;
        }
        ca.cand = tet ? b->d_cand_tet.p : b->d_cand_tri.p;
        ca.cand_capacity = static_cast<unsigned long long>(tet ? b->cap_tet : b->cap_tri);
        ca.surv = b->d_frontier[tet].p;
        if (generic) {
          ++b->last_launches;
          // 20 resident warps per SM in 128-thread blocks (95 registers, no spills, now that the fp64 leaf predicate lives
          // in k_leaf_undecided; 16 warps at 127 registers: 6 % slower, 24 at 80: spills, profiles/r2u_split_and_filter.txt).
          if (tet) k_filter<true, 128, 5><<<sms * 5 * tune.fgrid, 128, 0, st>>>(ca);
          else k_filter<false, 128, 5><<<sms * 5 * tune.fgrid, 128, 0, st>>>(ca);
        }
        if (use_small) {
          sm.small_list = tet ? pa.small_list_tet : pa.small_list;
          sm.surv = ca.surv;
          sm.cand = ca.cand;
          sm.cand_capacity = ca.cand_capacity;
          ++b->last_launches;
          // Resident 128-thread blocks per SM (measured: 6 for tet-triangle pairs, 5 for tet-tet, whose stage A needs
          // 96 registers; profiles/README.md).
          const int per_sm = tet ? kSmallBlocksTet : kSmallBlocksTri;
          const int small_blocks = std::min(sms * per_sm, (n_small + kSmallWarps - 1) / kSmallWarps);
          sm.items = b->d_small_items.p + (tet ? b->cap_small_items : 0);
          sm.items_capacity = static_cast<unsigned long long>(b->cap_small_items);
          sm.split_at = std::min(tune.split_at, b->small_stack_cap / 2);
          const bool split = b->num_large_small_pairs > 0 && tune.split != 0;
          SmallKernel(tet != 0, split ? 1 : 0)<<<small_blocks, kSmallWarps * 32, small_smem, st>>>(sm);
          if (split) {   // the work items of the large pairs
            ++b->last_launches;
            SmallKernel(tet != 0, 2)<<<sms * per_sm, kSmallWarps * 32, small_smem, st>>>(sm);
          }
        }
        if (tet) settle_undecided();
        mark(3 + tet);
      }
      if (b->info.num_soft_rigid_pairs) {
        ca.surv = b->d_frontier[0].p;
        ++b->last_launches;
        const ClipLaunch cl = ClipConfig(false);
        cl.fn<<<sms * cl.blocks_per_sm, cl.threads, cl.smem, st>>>(ca);
      }
      mark(5);
      if (b->info.num_soft_soft_pairs && b->surface_bfs) {
        // The surface-seeded intersector (hydro_bfs_kernels.cuh): the survivors of the descent over the boundary trees
        // are the seeds; breadth-first waves until one appends nothing. The wave bounds live on the device; the host
        // looks at them once per kBfsWavesPerLook waves (a wave over an empty range returns at once).
        constexpr int kBfsWavesPerLook = 8;
        BfsArgs ba{};
        ba.c = ca;
        ba.pairs = b->d_pairs.p;
        ba.queue = b->d_frontier[1].p;
        ba.queue_capacity = static_cast<unsigned long long>(b->cap_frontier);
        ba.table = b->d_bfs_table.p;
        ba.table_mask = static_cast<unsigned long long>(b->BfsTableEntries()) - 1ull;
        ba.neighbors = ctx->d_tet_neighbors.p;
        CUDA_OK(cudaMemsetAsync(ba.table, 0xFF, b->BfsTableEntries() * sizeof(unsigned long long), st));
        b->last_launches += 2;
        k_bfs_seed<<<sms * 8, 128, 0, st>>>(ba);
        ba.first = 1;
        k_bfs_advance<<<1, 1, 0, st>>>(ba);
        ba.first = 0;
        for (int look = 0;; ++look) {
          for (int w = 0; w < kBfsWavesPerLook; ++w) {
            b->last_launches += 2;
            // 7 resident 64-thread blocks per SM at 142 registers; bounding the kernel to 96 registers for 10 blocks
            // changed nothing (profiles/r4d_*: 77.2 / 77.0 ms on 1024 scenes of config 5).
            k_bfs_wave<64, 1><<<sms * 7, 64, clip_smem_bytes(64, 8), st>>>(ba);
            k_bfs_advance<<<1, 1, 0, st>>>(ba);
          }
          unsigned long long bounds[2] = {0, 0};
          int failed = 0;
          CUDA_OK(cudaMemcpyAsync(bounds, &b->d_counters.p->bfs_begin, sizeof(bounds), cudaMemcpyDeviceToHost, st));
          CUDA_OK(cudaMemcpyAsync(&failed, &b->d_counters.p->error, sizeof(failed), cudaMemcpyDeviceToHost, st));
          CUDA_OK(cudaStreamSynchronize(st));
          if (bounds[0] >= bounds[1] || failed != 0) break;
          if (look > 100000) return Fail(DBH_ERR_CUDA, "surface intersector: the breadth-first search does not end");
        }
      } else if (b->info.num_soft_soft_pairs) {
        ca.surv = b->d_frontier[1].p;
        ++b->last_launches;
        const ClipLaunch cl = ClipConfig(true);
        cl.fn<<<sms * cl.blocks_per_sm, cl.threads, cl.smem, st>>>(ca);
      }
      // Exceptional candidates (normally none): a tiny launch that returns at once when the list is empty.
      ++b->last_launches;
      k_clip_slow<<<sms, 64, 0, st>>>(ca);
      mark(6);
      // Half-space pairs of this chunk: pose + root test per pair, then one thread per element (count pass).
      pl.q_begin = static_cast<int>(std::lower_bound(b->plane_pairs.begin(), b->plane_pairs.end(), p0) - b->plane_pairs.begin());
      pl.q_end = static_cast<int>(std::lower_bound(b->plane_pairs.begin(), b->plane_pairs.end(), p1) - b->plane_pairs.begin());
      int plane_blocks = 0;
      if (pl.q_end > pl.q_begin) {
        const long long elems = b->plane_elem_off[pl.q_end] - b->plane_elem_off[pl.q_begin];
        plane_blocks = static_cast<int>(std::min<long long>((elems + 127) / 128, static_cast<long long>(sms) * 16));
        b->last_launches += 2;
        k_plane_setup<<<(pl.q_end - pl.q_begin + 127) / 128, 128, 0, st>>>(pl);
        k_plane<false><<<plane_blocks, 128, 0, st>>>(pl);
      }
      mark(7);
      sa.pair_begin = p0;
      sa.num_pairs = p1;
      sa.first_chunk = p0 == 0;
      sa.num_blocks = (p1 - p0 + kScanTile - 1) / kScanTile;
      ++b->last_launches;
      k_scan_partials<<<sa.num_blocks, kScanThreads, 0, st>>>(sa);
      ++b->last_launches;
      k_scan_blocks<<<1, kScanThreads, 0, st>>>(sa);
      ++b->last_launches;
      k_scan_apply<<<sa.num_blocks, kScanThreads, 0, st>>>(sa);
      mark(8);
      ++b->last_launches;
      if (pk.representation == DBH_POLYGON) {
        k_pack<4, 1><<<sms * 4 * tune.pgrid, 128, 0, st>>>(pk);
      } else {
        k_pack<4, 0><<<sms * 4 * tune.pgrid, 128, 0, st>>>(pk);
      }
      if (plane_blocks) {  // pack pass of the half-space pairs: straight into the final layout
        ++b->last_launches;
        k_plane<true><<<plane_blocks, 128, 0, st>>>(pl);
      }
      mark(9);
      su.pair_begin = p0;
      su.pair_end = p1;
      b->last_launches += 3;
      k_surface<<<sms * 4, 128, 0, st>>>(su);
      k_surface_big_sum<<<sms * 4, 128, 0, st>>>(su);
      k_surface_big_finish<<<1, 128, 0, st>>>(su);
      ++b->last_launches;
      k_next_chunk<<<1, 64, 0, st>>>(b->d_counters.p);
      mark(10);
      if (b->timing) ++b->timed_chunks;
    }
    ++b->last_launches;
    k_finalize<<<1, 32, 0, st>>>(su, b->num_pairs);
    if (scene_stats_dev) {
      ++b->last_launches;
      k_scene_stats<<<(b->num_pairs + 255) / 256, 256, 0, st>>>(b->d_pairs.p, b->num_pairs,
                                                               reinterpret_cast<long long*>(scene_stats_dev));
    }
  }
  CUDA_OK(cudaMemcpyAsync(b->h_counters, b->d_counters.p, sizeof(Counters), cudaMemcpyDeviceToHost, st));
  if (b->timing) cudaEventRecord(b->ev_end, st);
  CUDA_OK(cudaGetLastError());
  return DBH_OK;
  });
}

dbh_status dbh_batch_sync(dbh_batch* b) {
  return Guarded([&]() -> dbh_status {
  if (!b) return Fail(DBH_ERR_INVALID_ARGUMENT, "batch is null");
  CUDA_OK(cudaSetDevice(b->ctx->device));
  CUDA_OK(cudaStreamSynchronize(b->stream));
  if (b->timing) {
    float st[kStageEvents] = {};
    for (int c = 0; c < b->timed_chunks; ++c)
      for (int i = 0; i + 1 < kStageEvents; ++i) {
        float ms = 0;
        cudaEventElapsedTime(&ms, b->ev_chunk[kStageEvents * c + i], b->ev_chunk[kStageEvents * c + i + 1]);
        st[i] += ms;
      }
    cudaEventElapsedTime(&st[kStageEvents - 1], b->ev_begin, b->ev_end);
    std::memcpy(b->stage_ms, st, sizeof(st));
    // The eight figures of dbh_batch_kernel_times_ms: {pose+root, per-level descent, narrow phase tet-tri (cull +
    // exact clip), narrow phase tet-tet (cull + exact clip + half-space count), scan, pack, surface, total}.
    const float t[8] = {st[0], st[1], st[2] + st[4], st[3] + st[5] + st[6], st[7], st[8], st[9], st[kStageEvents - 1]};
    std::memcpy(b->times_ms, t, sizeof(t));
  }
  const Counters& c = *b->h_counters;
  if (c.error != 0) {
    static const char* kWhere[] = {"",          "BV-pair frontier", "element-pair list", "polygon staging",
                                   "(unused)", "(unused)", "face arrays",
                                   "vertex arrays", "exceptional-candidate list", "node-pair stack of a small pair"};
    if (c.error == DBH_ERR_CAPACITY) {
      const int d = (c.error_detail >= 1 && c.error_detail <= 9) ? c.error_detail : 0;
      return Fail(DBH_ERR_CAPACITY, std::string("device work list overflow: ") + kWhere[d]);
    }
    if (c.error == DBH_ERR_POLYGON_SIZE)
      return Fail(DBH_ERR_POLYGON_SIZE, "geometry pair " + std::to_string(c.error_detail) +
                                            ": a contact polygon outgrew what the reference defines (tet-triangle: more "
                                            "than 7 vertices, mesh_intersection.cc:261-262; tet-tet: more than 16 "
                                            "intermediate vertices)");
    if (c.error == DBH_ERR_RIGID_RIGID)
      return Fail(DBH_ERR_RIGID_RIGID, "rigid-rigid contact between geometries " +
                                           std::to_string(b->pairs[c.error_detail].id_m) + " and " +
                                           std::to_string(b->pairs[c.error_detail].id_n) +
                                           " is not supported by the hydroelastic contact model");
    if (c.error == DBH_ERR_HALF_SPACE_HALF_SPACE)
      return Fail(DBH_ERR_HALF_SPACE_HALF_SPACE, "contact between two half spaces with ids " +
                                                     std::to_string(b->pairs[c.error_detail].id_m) + " and " +
                                                     std::to_string(b->pairs[c.error_detail].id_n) + " is not allowed");
    if (c.error == DBH_ERR_COMPLIANT_HALF_SPACE_COMPLIANT_MESH)
      return Fail(DBH_ERR_COMPLIANT_HALF_SPACE_COMPLIANT_MESH,
                  "hydroelastic contact between two compliant geometries, one of which is a half space (ids " +
                      std::to_string(b->pairs[c.error_detail].id_m) + " and " +
                      std::to_string(b->pairs[c.error_detail].id_n) + "), is not allowed");
    if (c.error == DBH_ERR_PLANE_NORMAL || c.error == DBH_ERR_REFERENCE_DEMAND)
      return Fail(static_cast<dbh_status>(c.error),
                  "geometry pair " + std::to_string(c.error_detail) +
                      (c.error == DBH_ERR_PLANE_NORMAL
                           ? ": half-space normal too small to normalise (plane.cc:16-29)"
                           : ": a triangle edge does not cross the half-space boundary it was asked to be split by "
                             "(DRAKE_DEMAND, mesh_half_space_intersection.cc:43,66)"));
    if (c.error == DBH_ERR_UNSUPPORTED)
      return Fail(DBH_ERR_UNSUPPORTED, "geometry pair " + std::to_string(c.error_detail) +
                                           " has a geometry without a hydroelastic representation");
    return Fail(static_cast<dbh_status>(c.error), "device-side failure " + std::to_string(c.error));
  }
  b->query_done = true;
  return DBH_OK;
  });
}

dbh_status dbh_batch_set_body_kinematics(dbh_batch* b, const double* p_WBo, const double* V_WB) {
  return Guarded([&]() -> dbh_status {
  if (!b) return Fail(DBH_ERR_INVALID_ARGUMENT, "batch is null");
  CUDA_OK(cudaSetDevice(b->ctx->device));
  const size_t G = static_cast<size_t>(std::max(1, b->num_geoms));
  b->have_body_origin = p_WBo != nullptr;
  b->have_body_vel = V_WB != nullptr;
  if (p_WBo) {
    CUDA_OK(b->d_body_origin.ensure(3 * G));
    CUDA_OK(cudaMemcpyAsync(b->d_body_origin.p, p_WBo, 3 * static_cast<size_t>(b->num_geoms) * sizeof(double),
                            cudaMemcpyHostToDevice, b->stream));
  }
  if (V_WB) {
    CUDA_OK(b->d_body_vel.ensure(6 * G));
    CUDA_OK(cudaMemcpyAsync(b->d_body_vel.p, V_WB, 6 * static_cast<size_t>(b->num_geoms) * sizeof(double),
                            cudaMemcpyHostToDevice, b->stream));
  }
  return DBH_OK;
  });
}

dbh_status dbh_batch_set_contact_properties(dbh_batch* b, const double* hydroelastic_modulus,
                                            const double* hunt_crossley_dissipation, const double* dynamic_friction) {
  return Guarded([&]() -> dbh_status {
  if (!b || !hydroelastic_modulus || !hunt_crossley_dissipation || !dynamic_friction)
    return Fail(DBH_ERR_INVALID_ARGUMENT, "set_contact_properties: bad argument");
  CUDA_OK(cudaSetDevice(b->ctx->device));
  std::vector<double> props(3 * static_cast<size_t>(std::max(1, b->num_geoms)), 0.0);
  for (int g = 0; g < b->num_geoms; ++g) {
    // DRAKE_DEMAND(modulus > 0), DRAKE_DEMAND(dissipation >= 0) (hydroelastic_engine.cc:52-53).
    if (!(hydroelastic_modulus[g] > 0.0) || !(hunt_crossley_dissipation[g] >= 0.0) || !(dynamic_friction[g] >= 0.0))
      return Fail(DBH_ERR_INVALID_ARGUMENT, "set_contact_properties: modulus must be > 0, dissipation and friction >= 0");
    props[3 * g] = hydroelastic_modulus[g];
    props[3 * g + 1] = hunt_crossley_dissipation[g];
    props[3 * g + 2] = dynamic_friction[g];
  }
  CUDA_OK(b->d_props.ensure(props.size()));
  CUDA_OK(cudaMemcpy(b->d_props.p, props.data(), props.size() * sizeof(double), cudaMemcpyHostToDevice));
  b->have_props = true;
  return DBH_OK;
  });
}

dbh_status dbh_batch_compute_spatial_forces(dbh_batch* b, double vslip_regularizer) {
  return Guarded([&]() -> dbh_status {
  if (!b || !(vslip_regularizer > 0.0)) return Fail(DBH_ERR_INVALID_ARGUMENT, "compute_spatial_forces: bad argument");
  if (!b->have_props) return Fail(DBH_ERR_NOT_READY, "compute_spatial_forces: call dbh_batch_set_contact_properties first");
  if (!b->query_enqueued) return Fail(DBH_ERR_NOT_READY, "compute_spatial_forces: no query on this batch yet");
  CUDA_OK(cudaSetDevice(b->ctx->device));
  CUDA_OK(b->d_forces.ensure(18 * static_cast<size_t>(std::max(1, b->num_pairs))));
  if (b->num_pairs > 0) {
    WrenchArgs w{};
    w.surf_pair = b->d_surf_pair.p;
    w.pairs = b->d_pairs.p;
    w.poses = b->d_poses;
    w.body_origin = b->have_body_origin ? b->d_body_origin.p : nullptr;
    w.body_vel = b->have_body_vel ? b->d_body_vel.p : nullptr;
    w.props = b->d_props.p;
    w.surf_face_off = b->d_surf_face_off.p;
    w.surf_vert_off = b->d_surf_vert_off.p;
    w.surf_centroid = b->r_surf_centroid.p;
    w.representation = b->representation;
    w.areas = b->r_areas.p;
    w.normals = b->r_normals.p;
    w.centroids = b->r_centroids.p;
    w.e_centroid = b->r_e_centroid.p;
    w.vertices = b->r_vertices.p;
    w.pressures = b->r_pressures.p;
    w.face_data = b->r_face_data.p;
    w.vslip_regularizer = vslip_regularizer;
    w.forces = b->d_forces.p;
    w.counters = b->d_counters.p;
    k_wrench<<<b->ctx->prop.multiProcessorCount * 8, 128, 0, b->stream>>>(w);
    CUDA_OK(cudaGetLastError());
  }
  b->forces_done = true;
  return DBH_OK;
  });
}

dbh_status dbh_batch_copy_spatial_forces(dbh_batch* b, double* F_Ac_W, double* F_Ao_W, double* F_Bo_W) {
  return Guarded([&]() -> dbh_status {
  if (!b) return Fail(DBH_ERR_INVALID_ARGUMENT, "batch is null");
  if (!b->forces_done) return Fail(DBH_ERR_NOT_READY, "copy_spatial_forces: dbh_batch_compute_spatial_forces has not run");
  CUDA_OK(cudaSetDevice(b->ctx->device));
  if (dbh_status st = dbh_batch_sync(b)) return st;
  const size_t ns = static_cast<size_t>(b->h_counters->surfaces);
  if (ns == 0) return DBH_OK;
  std::vector<double> all(18 * ns);
  CUDA_OK(cudaMemcpy(all.data(), b->d_forces.p, all.size() * sizeof(double), cudaMemcpyDeviceToHost));
  for (size_t s = 0; s < ns; ++s) {
    if (F_Ac_W) std::memcpy(F_Ac_W + 6 * s, all.data() + 18 * s, 6 * sizeof(double));
    if (F_Ao_W) std::memcpy(F_Ao_W + 6 * s, all.data() + 18 * s + 6, 6 * sizeof(double));
    if (F_Bo_W) std::memcpy(F_Bo_W + 6 * s, all.data() + 18 * s + 12, 6 * sizeof(double));
  }
  return DBH_OK;
  });
}

dbh_status dbh_batch_copy_spatial_forces_async(dbh_batch* b, double* out18, int64_t* num_surfaces) {
  return Guarded([&]() -> dbh_status {
  if (!b || !num_surfaces) return Fail(DBH_ERR_INVALID_ARGUMENT, "copy_spatial_forces_async: bad argument");
  if (!b->query_done || !b->forces_done)
    return Fail(DBH_ERR_NOT_READY, "dbh_batch_compute_spatial_forces + dbh_batch_sync first");
  CUDA_OK(cudaSetDevice(b->ctx->device));
  const size_t ns = static_cast<size_t>(b->h_counters->surfaces);
  *num_surfaces = static_cast<int64_t>(ns);
  if (out18 && ns)
    CUDA_OK(cudaMemcpyAsync(out18, b->d_forces.p, ns * 18 * sizeof(double), cudaMemcpyDeviceToHost, b->stream));
  return DBH_OK;
  });
}

dbh_status dbh_batch_compute_discrete_pairs(dbh_batch* b) {
  return Guarded([&]() -> dbh_status {
  if (!b) return Fail(DBH_ERR_INVALID_ARGUMENT, "batch is null");
  if (!b->query_enqueued) return Fail(DBH_ERR_NOT_READY, "no query enqueued");
  CUDA_OK(cudaSetDevice(b->ctx->device));
  // Sized for the face capacity: the face count of the query is still on the device.
  CUDA_OK(b->d_discrete.ensure(static_cast<size_t>(std::max<long long>(b->cap_faces, 1)) * 9));
  DiscreteArgs da{};
  da.pairs = b->d_pairs.p;
  da.face_pair = b->r_face_pair.p;
  da.areas = b->r_areas.p;
  da.normals = b->r_normals.p;
  da.centroids = b->r_centroids.p;
  da.e_centroid = b->r_e_centroid.p;
  da.grad_eM = b->r_grad_eM.p;
  da.grad_eN = b->r_grad_eN.p;
  da.out = b->d_discrete.p;
  da.counters = b->d_counters.p;
  CUDA_OK(cudaMemsetAsync(&b->d_counters.p->discrete_pairs, 0, sizeof(unsigned long long), b->stream));
  k_discrete<<<b->ctx->prop.multiProcessorCount * 8, 256, 0, b->stream>>>(da);
  CUDA_OK(cudaMemcpyAsync(&b->h_counters->discrete_pairs, &b->d_counters.p->discrete_pairs, sizeof(unsigned long long),
                          cudaMemcpyDeviceToHost, b->stream));
  CUDA_OK(cudaGetLastError());
  b->discrete_done = true;
  return DBH_OK;
  });
}

dbh_status dbh_batch_copy_discrete_pairs_async(dbh_batch* b, double* out, int64_t* num_faces) {
  return Guarded([&]() -> dbh_status {
  if (!b || !num_faces) return Fail(DBH_ERR_INVALID_ARGUMENT, "copy_discrete_pairs: bad argument");
  if (!b->query_done || !b->discrete_done)
    return Fail(DBH_ERR_NOT_READY, "dbh_batch_compute_discrete_pairs + dbh_batch_sync first");
  CUDA_OK(cudaSetDevice(b->ctx->device));
  const size_t nf = b->h_counters->faces;
  *num_faces = static_cast<int64_t>(nf);
  if (out && nf) CUDA_OK(cudaMemcpyAsync(out, b->d_discrete.p, nf * 9 * sizeof(double), cudaMemcpyDeviceToHost, b->stream));
  return DBH_OK;
  });
}

dbh_status dbh_batch_compute_deformable_contact(dbh_batch* b) {
  return Guarded([&]() -> dbh_status {
  if (!b) return Fail(DBH_ERR_INVALID_ARGUMENT, "batch is null");
  if (!b->query_enqueued) return Fail(DBH_ERR_NOT_READY, "no query enqueued");
  if (b->representation != DBH_POLYGON)
    return Fail(DBH_ERR_INVALID_ARGUMENT, "deformable contact data is defined on polygon surfaces (DeformableContactSurface holds "
                                          "a PolygonSurfaceMesh): query with DBH_POLYGON");
  CUDA_OK(cudaSetDevice(b->ctx->device));
  CUDA_OK(b->d_deformable.ensure(static_cast<size_t>(std::max<long long>(b->cap_faces, 1)) * kDeformableRecord));
  DeformableArgs da{};
  da.pairs = b->d_pairs.p;
  da.poses = b->d_poses;
  da.tris = b->ctx->d_tris.p;
  da.tets = b->ctx->d_tets.p;
  da.face_pair = b->r_face_pair.p;
  da.elem_a = b->r_elem_a.p;
  da.elem_b = b->r_elem_b.p;
  da.normals = b->r_normals.p;
  da.centroids = b->r_centroids.p;
  da.e_centroid = b->r_e_centroid.p;
  da.grad_eM = b->r_grad_eM.p;
  da.grad_eN = b->r_grad_eN.p;
  da.out = b->d_deformable.p;
  da.counters = b->d_counters.p;
  k_deformable<<<b->ctx->prop.multiProcessorCount * 8, 256, 0, b->stream>>>(da);
  CUDA_OK(cudaGetLastError());
  b->deformable_done = true;
  return DBH_OK;
  });
}

dbh_status dbh_batch_copy_deformable_contact_async(dbh_batch* b, double* out, int32_t* element_a, int32_t* element_b,
                                                   int64_t* num_faces) {
  return Guarded([&]() -> dbh_status {
  if (!b || !num_faces) return Fail(DBH_ERR_INVALID_ARGUMENT, "copy_deformable_contact: bad argument");
  if (!b->query_done || !b->deformable_done)
    return Fail(DBH_ERR_NOT_READY, "dbh_batch_compute_deformable_contact + dbh_batch_sync first");
  CUDA_OK(cudaSetDevice(b->ctx->device));
  const size_t nf = b->h_counters->faces;
  *num_faces = static_cast<int64_t>(nf);
  if (out && nf)
    CUDA_OK(cudaMemcpyAsync(out, b->d_deformable.p, nf * kDeformableRecord * sizeof(double), cudaMemcpyDeviceToHost, b->stream));
  if (element_a && nf)
    CUDA_OK(cudaMemcpyAsync(element_a, b->r_elem_a.p, nf * sizeof(int32_t), cudaMemcpyDeviceToHost, b->stream));
  if (element_b && nf)
    CUDA_OK(cudaMemcpyAsync(element_b, b->r_elem_b.p, nf * sizeof(int32_t), cudaMemcpyDeviceToHost, b->stream));
  return DBH_OK;
  });
}

dbh_status dbh_batch_stage_times_ms(dbh_batch* b, float* out, int32_t n) {
  if (!b || !out || n < 0) return Fail(DBH_ERR_INVALID_ARGUMENT, "stage_times: bad argument");
  for (int i = 0; i < n; ++i) out[i] = i < kStageEvents ? b->stage_ms[i] : 0.f;
  return DBH_OK;
}

dbh_status dbh_batch_kernel_times_ms(dbh_batch* b, float* out8) {
  return Guarded([&]() -> dbh_status {
  if (!b || !out8) return Fail(DBH_ERR_INVALID_ARGUMENT, "kernel_times: bad argument");
  std::memcpy(out8, b->times_ms, sizeof(b->times_ms));
  return DBH_OK;
  });
}

dbh_status dbh_batch_result_sizes(dbh_batch* b, dbh_result_sizes* out) {
  return Guarded([&]() -> dbh_status {
  if (!b || !out) return Fail(DBH_ERR_INVALID_ARGUMENT, "result_sizes: bad argument");
  if (!b->query_done) return Fail(DBH_ERR_NOT_READY, "no completed query");
  const Counters& c = *b->h_counters;
  std::memset(out, 0, sizeof(*out));
  out->num_surfaces = static_cast<int64_t>(c.surfaces);
  out->num_faces = static_cast<int64_t>(c.faces);
  out->num_vertices = static_cast<int64_t>(c.vertices);
  out->num_polygons = static_cast<int64_t>(c.polygons);
  out->representation = b->representation;
  out->face_data_ints = b->representation == DBH_POLYGON ? static_cast<int64_t>(c.polygons + c.poly_verts)
                                                         : 3 * static_cast<int64_t>(c.faces);
  return DBH_OK;
  });
}

dbh_status dbh_batch_stats(dbh_batch* b, dbh_stats* out) {
  return Guarded([&]() -> dbh_status {
  if (!b || !out) return Fail(DBH_ERR_INVALID_ARGUMENT, "stats: bad argument");
  if (!b->query_done) return Fail(DBH_ERR_NOT_READY, "no completed query");
  const Counters& c = *b->h_counters;
  std::memset(out, 0, sizeof(*out));
  out->geometry_pairs_tested =
      static_cast<int64_t>(b->info.num_soft_rigid_pairs + b->info.num_soft_soft_pairs + b->info.num_half_space_pairs);
  out->half_space_elements = static_cast<int64_t>(c.plane_tets + c.plane_tris);
  out->geometry_pairs_overlapping = static_cast<int64_t>(c.root_overlaps);
  out->bv_pair_tests = static_cast<int64_t>(c.bv_tests);
  out->bv_pair_tests_tet_tet = static_cast<int64_t>(c.bv_tests_tet);
  out->leaf_pairs_tet_tet = static_cast<int64_t>(c.leaf_tet_total + c.small_leaf_tet);
  out->leaf_pairs = static_cast<int64_t>(c.leaf_tri_total + c.leaf_tet_total + c.small_leaf_tri + c.small_leaf_tet);
  out->cull_violations = static_cast<int64_t>(c.cull_violations);
  out->exceptional_clips = static_cast<int64_t>(c.slow_total);
  out->leaf_exact_tests = static_cast<int64_t>(c.leaf_exact);
  out->tet_tri_pairs = static_cast<int64_t>(c.cand_tri);
  out->tet_tet_pairs = static_cast<int64_t>(c.cand_tet);
  out->polygons = static_cast<int64_t>(c.polygons);
  out->faces = static_cast<int64_t>(c.faces);
  out->vertices = static_cast<int64_t>(c.vertices);
  out->surfaces = static_cast<int64_t>(c.surfaces);
  out->total_area = c.total_area;
  out->tet_tri_full_clips = static_cast<int64_t>(c.full_clips_tri);
  out->tet_tet_full_clips = static_cast<int64_t>(c.full_clips_tet);
  out->discrete_pairs = b->discrete_done ? static_cast<int64_t>(c.discrete_pairs) : 0;
  out->fallback_pairs = b->last_query_kind == DBH_QUERY_WITH_FALLBACK ? static_cast<int64_t>(c.fallback) : 0;
  out->bfs_pairs = static_cast<int64_t>(c.bfs_pairs);
  out->bfs_waves = static_cast<int64_t>(c.bfs_waves);
  // BV test 64 B (two 32 B fp32 nodes; SURVEY.md §8(d) counts 128 B for the
  // reference-shaped fp64 node); tet-tri pair 260 B in; tet-tet pair 392 B in;
  // emitted n-gon 32 n + 96 (soft-rigid) / + 120 (soft-soft) out; pose 96 B per geometry.
  // Triangle representation adds 32 B centroid vertex + n (12 + 24 + 8) B per polygon.
  long long bytes = out->bv_pair_tests * 64 + out->tet_tri_pairs * 260 + out->tet_tet_pairs * 392 +
                    static_cast<long long>(c.poly_verts) * 32 + out->polygons * 108 + 96ll * b->num_geoms;
  // Half-space pairs: one record per element (tetrahedron: 16 B indices + 96 B vertices + 32 B vertex pressures +
  // 24 B gradient = 168 B; triangle: 12 + 72 + 24 = 108 B).
  bytes += static_cast<long long>(c.plane_tets) * 168 + static_cast<long long>(c.plane_tris) * 108;
  if (b->representation == DBH_TRIANGLE) bytes += out->polygons * 32 + static_cast<long long>(c.poly_verts) * 44;
  out->algorithmic_bytes = bytes;
  return DBH_OK;
  });
}

dbh_status dbh_batch_copy_candidates(dbh_batch* b, int32_t kind, int32_t* out_triples, int64_t* count) {
  return Guarded([&]() -> dbh_status {
  if (!b || !count || kind < 0 || kind > 3) return Fail(DBH_ERR_INVALID_ARGUMENT, "copy_candidates: bad argument");
  if (!b->query_done) return Fail(DBH_ERR_NOT_READY, "no completed query");
  CUDA_OK(cudaSetDevice(b->ctx->device));
  const bool tet = (kind & 1) != 0, raw = (kind & 2) != 0;
  // The narrow phase left the leaf list as {pair, pool elem a, pool elem b, flag}; the candidates
  // are the flagged entries (kind | 2: every leaf pair the conservative descent passed).
  if (b->num_chunks != 1)
    return Fail(DBH_ERR_NOT_READY, "candidate lists are only kept for single-chunk queries (dbh_batch_set_chunks(batch, 1))");
  if (b->small_used && !b->record_candidates)
    return Fail(DBH_ERR_NOT_READY,
                "the last query culled its small pairs on chip without keeping their candidate lists: call "
                "dbh_batch_set_record_candidates(batch, 1) before the query");
  const int64_t listed = static_cast<int64_t>(tet ? b->h_counters->leaf_tet_total : b->h_counters->leaf_tri_total);
  const int64_t n = raw ? listed : static_cast<int64_t>(tet ? b->h_counters->cand_tet : b->h_counters->cand_tri);
  *count = n;
  if (!out_triples || n == 0) return DBH_OK;
  std::vector<int4> tmp(listed);
  CUDA_OK(cudaMemcpy(tmp.data(), tet ? b->d_cand_tet.p : b->d_cand_tri.p, listed * sizeof(int4),
                     cudaMemcpyDeviceToHost));
  int64_t k = 0;
  for (int64_t i = 0; i < listed && k < n; ++i) {
    if (!raw && tmp[i].w != 1) continue;
    const PairRec& pr = b->pairs[tmp[i].x];
    out_triples[3 * k] = tmp[i].x;
    out_triples[3 * k + 1] = tmp[i].y - pr.elem_base_a;
    out_triples[3 * k + 2] = tmp[i].z - pr.elem_base_b;
    ++k;
  }
  if (k != n) return Fail(DBH_ERR_CUDA, "candidate flags disagree with the candidate counter");
  return DBH_OK;
  });
}

dbh_status dbh_batch_set_geometry_bounds(dbh_batch* b, const double* boxes) {
  return Guarded([&]() -> dbh_status {
  if (!b || !boxes) return Fail(DBH_ERR_INVALID_ARGUMENT, "set_geometry_bounds: bad argument");
  CUDA_OK(cudaSetDevice(b->ctx->device));
  for (int g = 0; g < b->num_geoms; ++g) {
    if (b->geom_asset[g] >= 0) continue;   // meshes and half spaces carry their own bound
    GeomBound& gb = b->geom_bounds[g];
    bool finite = true;
    for (int k = 0; k < 6; ++k) finite = finite && std::isfinite(boxes[6 * g + k]);
    gb.kind = finite ? kBoundBox : kBoundEverywhere;
    for (int k = 0; k < 3; ++k) {
      gb.c[k] = finite ? boxes[6 * g + k] : 0.0;
      gb.h[k] = finite ? std::fabs(boxes[6 * g + 3 + k]) : 0.0;
    }
  }
  CUDA_OK(cudaMemcpy(b->d_geom_bounds.p, b->geom_bounds.data(), b->geom_bounds.size() * sizeof(GeomBound),
                     cudaMemcpyHostToDevice));
  return DBH_OK;
  });
}

dbh_status dbh_batch_copy_fallback_pairs(dbh_batch* b, int32_t* out_quads, int64_t* count) {
  return Guarded([&]() -> dbh_status {
  if (!b || !count) return Fail(DBH_ERR_INVALID_ARGUMENT, "copy_fallback_pairs: bad argument");
  if (!b->query_done) return Fail(DBH_ERR_NOT_READY, "no completed query");
  CUDA_OK(cudaSetDevice(b->ctx->device));
  const int64_t n = b->last_query_kind == DBH_QUERY_WITH_FALLBACK ? static_cast<int64_t>(b->h_counters->fallback) : 0;
  *count = n;
  if (!out_quads || n == 0) return DBH_OK;
  std::vector<int> idx(n);
  CUDA_OK(cudaMemcpy(idx.data(), b->d_fallback.p, n * sizeof(int), cudaMemcpyDeviceToHost));
  std::sort(idx.begin(), idx.end());   // the static pair list is in (scene, id_M, id_N) order
  for (int64_t k = 0; k < n; ++k) {
    const PairRec& pr = b->pairs[idx[k]];
    out_quads[4 * k] = pr.scene;
    out_quads[4 * k + 1] = pr.geom_a;
    out_quads[4 * k + 2] = pr.geom_b;
    out_quads[4 * k + 3] = pr.kind;
  }
  return DBH_OK;
  });
}

dbh_status dbh_batch_copy_geometry_pairs(dbh_batch* b, int32_t* out) {
  return Guarded([&]() -> dbh_status {
  if (!b || !out) return Fail(DBH_ERR_INVALID_ARGUMENT, "copy_geometry_pairs: bad argument");
  for (int i = 0; i < b->num_pairs; ++i) {
    out[4 * i] = b->pairs[i].scene;
    out[4 * i + 1] = b->pairs[i].geom_a;
    out[4 * i + 2] = b->pairs[i].geom_b;
    out[4 * i + 3] = b->pairs[i].kind;
  }
  return DBH_OK;
  });
}

dbh_status dbh_batch_device_results(dbh_batch* b, dbh_device_results* out) {
  return Guarded([&]() -> dbh_status {
  if (!b || !out) return Fail(DBH_ERR_INVALID_ARGUMENT, "device_results: bad argument");
  if (!b->query_done) return Fail(DBH_ERR_NOT_READY, "no completed query");
  std::memset(out, 0, sizeof(*out));
  out->surf_pair = b->d_surf_pair.p;
  out->surf_scene = b->r_surf_scene.p;
  out->surf_id_M = reinterpret_cast<const int64_t*>(b->r_surf_id_m.p);
  out->surf_id_N = reinterpret_cast<const int64_t*>(b->r_surf_id_n.p);
  out->surf_face_offset = reinterpret_cast<const int64_t*>(b->d_surf_face_off.p);
  out->surf_vertex_offset = reinterpret_cast<const int64_t*>(b->d_surf_vert_off.p);
  out->surf_total_area = b->r_surf_total_area.p;
  out->surf_centroid = b->r_surf_centroid.p;
  out->surf_flags = b->r_surf_flags.p;
  out->vertices = b->r_vertices.p;
  out->pressures = b->r_pressures.p;
  out->face_data = b->r_face_data.p;
  out->areas = b->r_areas.p;
  out->normals = b->r_normals.p;
  out->centroids = b->r_centroids.p;
  out->pressure_at_centroid = b->r_e_centroid.p;
  out->grad_eM = b->r_grad_eM.p;
  out->grad_eN = b->r_grad_eN.p;
  out->elem_M = b->r_elem_a.p;
  out->elem_N = b->r_elem_b.p;
  out->face_pair = b->r_face_pair.p;
  out->vertex_origin = b->plane_pairs.empty() ? nullptr : b->r_vertex_origin.p;
  return DBH_OK;
  });
}

dbh_status dbh_batch_copy_results(dbh_batch* b, const dbh_result_buffers* buf) {
  return Guarded([&]() -> dbh_status {
  if (dbh_status s = dbh_batch_copy_results_async(b, buf)) return s;
  CUDA_OK(cudaStreamSynchronize(b->stream));
  return DBH_OK;
  });
}

dbh_status dbh_batch_copy_results_async(dbh_batch* b, const dbh_result_buffers* buf) {
  return Guarded([&]() -> dbh_status {
  if (!b || !buf) return Fail(DBH_ERR_INVALID_ARGUMENT, "copy_results: bad argument");
  if (!b->query_done) return Fail(DBH_ERR_NOT_READY, "no completed query");
  CUDA_OK(cudaSetDevice(b->ctx->device));
  const Counters& c = *b->h_counters;
  const size_t ns = c.surfaces, nf = c.faces, nv = c.vertices;
  const size_t nfd = b->representation == DBH_POLYGON ? static_cast<size_t>(c.polygons + c.poly_verts) : 3 * nf;
  cudaStream_t st = b->stream;
  // The device arrays already have the caller's layout: straight D2H copies
  // (asynchronous when the destination is pinned), one synchronisation.
#define DBH_COPY(dst, src, count, T)                                                                  \
  if ((dst) && (count))                                                                               \
    CUDA_OK(cudaMemcpyAsync((dst), (src), (count) * sizeof(T), cudaMemcpyDeviceToHost, st))
  DBH_COPY(buf->surf_scene, b->r_surf_scene.p, ns, int32_t);
  DBH_COPY(buf->surf_id_M, b->r_surf_id_m.p, ns, int64_t);
  DBH_COPY(buf->surf_id_N, b->r_surf_id_n.p, ns, int64_t);
  DBH_COPY(buf->surf_face_offset, b->d_surf_face_off.p, ns + 1, int64_t);
  DBH_COPY(buf->surf_vertex_offset, b->d_surf_vert_off.p, ns + 1, int64_t);
  DBH_COPY(buf->surf_total_area, b->r_surf_total_area.p, ns, double);
  DBH_COPY(buf->surf_centroid, b->r_surf_centroid.p, 3 * ns, double);
  DBH_COPY(buf->surf_flags, b->r_surf_flags.p, ns, uint8_t);
  DBH_COPY(buf->vertices, b->r_vertices.p, 3 * nv, double);
  DBH_COPY(buf->pressures, b->r_pressures.p, nv, double);
  DBH_COPY(buf->face_data, b->r_face_data.p, nfd, int32_t);
  DBH_COPY(buf->areas, b->r_areas.p, nf, double);
  DBH_COPY(buf->normals, b->r_normals.p, 3 * nf, double);
  DBH_COPY(buf->centroids, b->r_centroids.p, 3 * nf, double);
  DBH_COPY(buf->pressure_at_centroid, b->r_e_centroid.p, nf, double);
  DBH_COPY(buf->grad_eM, b->r_grad_eM.p, 3 * nf, double);
  DBH_COPY(buf->grad_eN, b->r_grad_eN.p, 3 * nf, double);
  DBH_COPY(buf->elem_M, b->r_elem_a.p, nf, int32_t);
  DBH_COPY(buf->elem_N, b->r_elem_b.p, nf, int32_t);
  if (buf->vertex_origin) {
    if (b->plane_pairs.empty()) return Fail(DBH_ERR_INVALID_ARGUMENT, "vertex_origin: the batch has no half-space pair");
    DBH_COPY(buf->vertex_origin, b->r_vertex_origin.p, nv, int32_t);
  }
#undef DBH_COPY
  return DBH_OK;
  });
}

}  // extern "C"
