// GPU (linear) BVH construction for a mesh asset; see hydro_lbvh.cu.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "hydro_types.h"

namespace dbh {

// Builds the BVH of `num_elems` elements (3 or 4 vertex indices each) on the current CUDA device and
// returns it in the BvhNode layout (node 0 unused, node 1 root, siblings adjacent, one element per
// leaf). `height` = largest leaf depth. Returns false and sets `err` on failure.
bool BuildLbvhOnDevice(const double* verts, int num_verts, const int32_t* elems, int num_elems, int verts_per_elem,
                       std::vector<BvhNode>* nodes, int* height, std::string* err);

}  // namespace dbh
