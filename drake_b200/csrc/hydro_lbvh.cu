// GPU construction of a mesh's bounding-volume hierarchy (set-up path; replaces the top-down
// median-split build of geometry/proximity/bvh.cc:18-122 for large assets, and is the build a
// deforming mesh would re-run every step).
//
// Linear BVH (Karras 2012): Morton codes of the element-box centres, one radix sort (cub), every
// internal node found independently from the sorted keys, boxes fitted bottom-up with one atomic
// flag per internal node. The result is written in the layout the traversal kernels expect
// (hydro_types.h BvhNode): node 0 unused, node 1 the root, the two children of a node adjacent,
// one element per leaf (left = ~element). Leaf boxes are the elements' own Aabb computed with the
// reference's expressions (aabb.cc:102-117) in fp64, so they are bit-identical to the host
// builder's; internal boxes are exact min/max unions. The candidate element-pair set does not
// depend on the tree (DESIGN.md "Candidate sets"), so both builders give identical query results.
#include <cub/cub.cuh>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <string>
#include <vector>

#include "hydro_lbvh.h"

namespace dbh {
namespace {

struct MinMax {
  double mn[3], mx[3];
};

__device__ __forceinline__ unsigned int expand_bits(unsigned int v) {
  v = (v * 0x00010001u) & 0xFF0000FFu;
  v = (v * 0x00000101u) & 0x0F00F00Fu;
  v = (v * 0x00000011u) & 0xC30C30C3u;
  v = (v * 0x00000005u) & 0x49249249u;
  return v;
}

// Per element: exact box (min / max over its vertices) and the 30-bit Morton code of the box centre.
template <int kVerts>
__global__ void k_lbvh_boxes(const double* verts, const int32_t* elems, int n, double ox, double oy, double oz,
                             double sx, double sy, double sz, MinMax* boxes, unsigned int* codes, int* ids) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  MinMax b;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    b.mn[d] = 1e300;
    b.mx[d] = -1e300;
  }
#pragma unroll
  for (int v = 0; v < kVerts; ++v) {
    const double* p = verts + 3ll * elems[kVerts * e + v];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      b.mn[d] = fmin(b.mn[d], p[d]);
      b.mx[d] = fmax(b.mx[d], p[d]);
    }
  }
  boxes[e] = b;
  const double cx = ((b.mn[0] + b.mx[0]) / 2 - ox) * sx, cy = ((b.mn[1] + b.mx[1]) / 2 - oy) * sy,
               cz = ((b.mn[2] + b.mx[2]) / 2 - oz) * sz;
  const unsigned int qx = (unsigned int)fmin(fmax(cx * 1024.0, 0.0), 1023.0);
  const unsigned int qy = (unsigned int)fmin(fmax(cy * 1024.0, 0.0), 1023.0);
  const unsigned int qz = (unsigned int)fmin(fmax(cz * 1024.0, 0.0), 1023.0);
  codes[e] = (expand_bits(qx) << 2) | (expand_bits(qy) << 1) | expand_bits(qz);
  ids[e] = e;
}

// Length of the common prefix of the keys of sorted positions i and j; the position breaks ties so
// that all keys are distinct. -1 outside the array.
__device__ __forceinline__ int delta(const unsigned int* codes, int n, int i, int j) {
  if (j < 0 || j >= n) return -1;
  const unsigned long long a = ((unsigned long long)codes[i] << 32) | (unsigned int)i;
  const unsigned long long b = ((unsigned long long)codes[j] << 32) | (unsigned int)j;
  return __clzll((long long)(a ^ b));
}

// Children are encoded as: >= 0 internal node index, < 0 ~(sorted leaf position).
__global__ void k_lbvh_internal(const unsigned int* codes, int n, int* child_l, int* child_r, int* parent_internal,
                                int* parent_leaf) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n - 1) return;
  const int d = delta(codes, n, i, i + 1) - delta(codes, n, i, i - 1) >= 0 ? 1 : -1;
  const int dmin = delta(codes, n, i, i - d);
  int lmax = 2;
  while (delta(codes, n, i, i + lmax * d) > dmin) lmax <<= 1;
  int l = 0;
  for (int t = lmax >> 1; t >= 1; t >>= 1)
    if (delta(codes, n, i, i + (l + t) * d) > dmin) l += t;
  const int j = i + l * d;
  const int dnode = delta(codes, n, i, j);
  int s = 0;
  int t = l;
  do {
    t = (t + 1) >> 1;
    if (delta(codes, n, i, i + (s + t) * d) > dnode) s += t;
  } while (t > 1);
  const int gamma = i + s * d + min(d, 0);
  const int lo = min(i, j), hi = max(i, j);
  const int cl = lo == gamma ? ~gamma : gamma;
  const int cr = hi == gamma + 1 ? ~(gamma + 1) : gamma + 1;
  child_l[i] = cl;
  child_r[i] = cr;
  if (cl >= 0) parent_internal[cl] = i; else parent_leaf[~cl] = i;
  if (cr >= 0) parent_internal[cr] = i; else parent_leaf[~cr] = i;
}

__device__ __forceinline__ MinMax load_box_cg(const MinMax* p) {
  MinMax b;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    b.mn[d] = __ldcg(&p->mn[d]);
    b.mx[d] = __ldcg(&p->mx[d]);
  }
  return b;
}

// Bottom-up fit: the second thread to reach an internal node merges its children's boxes.
__global__ void k_lbvh_fit(int n, const int* ids, const MinMax* elem_boxes, const int* child_l, const int* child_r,
                           const int* parent_internal, const int* parent_leaf, MinMax* node_boxes, int* flags,
                           int* height) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  int node = parent_leaf[k];
  int depth = 1;
  // Depth of this leaf (for the level count of the traversal).
  for (int p = node; p > 0; p = parent_internal[p]) ++depth;
  atomicMax(height, depth);
  while (true) {
    __threadfence();
    if (atomicAdd(&flags[node], 1) == 0) return;  // the sibling subtree is not done yet
    __threadfence();
    const int cl = child_l[node], cr = child_r[node];
    // Boxes of internal children were written by other SMs: read them past L1 (a neighbouring box in
    // the same line may have been cached here before this one was written).
    const MinMax a = cl >= 0 ? load_box_cg(node_boxes + cl) : elem_boxes[ids[~cl]];
    const MinMax b = cr >= 0 ? load_box_cg(node_boxes + cr) : elem_boxes[ids[~cr]];
    MinMax u;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      u.mn[d] = fmin(a.mn[d], b.mn[d]);
      u.mx[d] = fmax(a.mx[d], b.mx[d]);
    }
    node_boxes[node] = u;
    if (node == 0) return;
    node = parent_internal[node];
  }
}

__device__ __forceinline__ void write_node(BvhNode* out, const MinMax& b, int left, int right) {
  BvhNode nd;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    nd.c[d] = (b.mn[d] + b.mx[d]) / 2;  // AabbMaker::Compute (aabb.cc:102-117)
    nd.h[d] = b.mx[d] - nd.c[d];
  }
  nd.left = left;
  nd.right = right;
  nd.pad[0] = nd.pad[1] = 0;
  *out = nd;
}

// Final layout: root at 1; the children of internal node k at 2 + 2k and 3 + 2k.
__global__ void k_lbvh_emit(int n, const int* ids, const MinMax* elem_boxes, const MinMax* node_boxes,
                            const int* child_l, const int* child_r, BvhNode* out) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n - 1) return;
  if (k == 0) write_node(out + 1, node_boxes[0], 2, 3);
  const int c[2] = {child_l[k], child_r[k]};
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    BvhNode* dst = out + 2 + 2 * k + s;
    if (c[s] >= 0) {
      write_node(dst, node_boxes[c[s]], 2 + 2 * c[s], 3 + 2 * c[s]);
    } else {
      const int elem = ids[~c[s]];
      write_node(dst, elem_boxes[elem], ~elem, -1);
    }
  }
}

template <class T>
struct Dev {
  T* p{nullptr};
  ~Dev() {
    if (p) cudaFree(p);
  }
  cudaError_t alloc(size_t n) { return cudaMalloc(&p, (n ? n : 1) * sizeof(T)); }
};

}  // namespace

#define LBVH_OK(expr)                                                         \
  do {                                                                        \
    cudaError_t _e = (expr);                                                  \
    if (_e != cudaSuccess) {                                                  \
      *err = std::string(#expr) + ": " + cudaGetErrorString(_e);              \
      return false;                                                           \
    }                                                                         \
  } while (0)

bool BuildLbvhOnDevice(const double* verts, int num_verts, const int32_t* elems, int num_elems, int verts_per_elem,
                       std::vector<BvhNode>* nodes, int* height, std::string* err) {
  const int n = num_elems;
  if (n < 2 || (verts_per_elem != 3 && verts_per_elem != 4)) {
    *err = "BuildLbvhOnDevice: needs at least two elements of 3 or 4 vertices";
    return false;
  }
  // Bounds of the mesh (host: one pass over the vertices).
  double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
  for (int v = 0; v < num_verts; ++v)
    for (int d = 0; d < 3; ++d) {
      lo[d] = std::min(lo[d], verts[3 * v + d]);
      hi[d] = std::max(hi[d], verts[3 * v + d]);
    }
  double sc[3];
  for (int d = 0; d < 3; ++d) sc[d] = hi[d] > lo[d] ? 1.0 / (hi[d] - lo[d]) : 0.0;

  Dev<double> d_verts;
  Dev<int32_t> d_elems;
  Dev<MinMax> d_boxes, d_node_boxes;
  Dev<unsigned int> d_codes, d_codes_sorted;
  Dev<int> d_ids, d_ids_sorted, d_cl, d_cr, d_pi, d_pl, d_flags, d_height;
  Dev<BvhNode> d_out;
  Dev<unsigned char> d_temp;
  LBVH_OK(d_verts.alloc(3 * static_cast<size_t>(num_verts)));
  LBVH_OK(d_elems.alloc(static_cast<size_t>(verts_per_elem) * n));
  LBVH_OK(d_boxes.alloc(n));
  LBVH_OK(d_node_boxes.alloc(n));
  LBVH_OK(d_codes.alloc(n));
  LBVH_OK(d_codes_sorted.alloc(n));
  LBVH_OK(d_ids.alloc(n));
  LBVH_OK(d_ids_sorted.alloc(n));
  LBVH_OK(d_cl.alloc(n));
  LBVH_OK(d_cr.alloc(n));
  LBVH_OK(d_pi.alloc(n));
  LBVH_OK(d_pl.alloc(n));
  LBVH_OK(d_flags.alloc(n));
  LBVH_OK(d_height.alloc(1));
  LBVH_OK(d_out.alloc(2 * static_cast<size_t>(n)));
  LBVH_OK(cudaMemcpy(d_verts.p, verts, 3 * sizeof(double) * num_verts, cudaMemcpyHostToDevice));
  LBVH_OK(cudaMemcpy(d_elems.p, elems, sizeof(int32_t) * verts_per_elem * n, cudaMemcpyHostToDevice));
  LBVH_OK(cudaMemset(d_flags.p, 0, sizeof(int) * n));
  LBVH_OK(cudaMemset(d_height.p, 0, sizeof(int)));
  LBVH_OK(cudaMemset(d_pi.p, 0, sizeof(int) * n));
  LBVH_OK(cudaMemset(d_out.p, 0, sizeof(BvhNode) * 2 * n));

  const int threads = 256, blocks = (n + threads - 1) / threads;
  if (verts_per_elem == 4)
    k_lbvh_boxes<4><<<blocks, threads>>>(d_verts.p, d_elems.p, n, lo[0], lo[1], lo[2], sc[0], sc[1], sc[2], d_boxes.p,
                                        d_codes.p, d_ids.p);
  else
    k_lbvh_boxes<3><<<blocks, threads>>>(d_verts.p, d_elems.p, n, lo[0], lo[1], lo[2], sc[0], sc[1], sc[2], d_boxes.p,
                                        d_codes.p, d_ids.p);
  size_t temp_bytes = 0;
  LBVH_OK(cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, d_codes.p, d_codes_sorted.p, d_ids.p, d_ids_sorted.p, n, 0,
                                          30));
  LBVH_OK(d_temp.alloc(temp_bytes));
  LBVH_OK(cub::DeviceRadixSort::SortPairs(d_temp.p, temp_bytes, d_codes.p, d_codes_sorted.p, d_ids.p, d_ids_sorted.p, n,
                                          0, 30));
  k_lbvh_internal<<<blocks, threads>>>(d_codes_sorted.p, n, d_cl.p, d_cr.p, d_pi.p, d_pl.p);
  k_lbvh_fit<<<blocks, threads>>>(n, d_ids_sorted.p, d_boxes.p, d_cl.p, d_cr.p, d_pi.p, d_pl.p, d_node_boxes.p,
                                  d_flags.p, d_height.p);
  k_lbvh_emit<<<blocks, threads>>>(n, d_ids_sorted.p, d_boxes.p, d_node_boxes.p, d_cl.p, d_cr.p, d_out.p);
  LBVH_OK(cudaGetLastError());
  nodes->assign(2 * static_cast<size_t>(n), BvhNode{});
  LBVH_OK(cudaMemcpy(nodes->data(), d_out.p, sizeof(BvhNode) * 2 * n, cudaMemcpyDeviceToHost));
  LBVH_OK(cudaMemcpy(height, d_height.p, sizeof(int), cudaMemcpyDeviceToHost));
  return true;
}

}  // namespace dbh
