// Host-only translation unit (built by nvcc with -Xcompiler -ffp-contract=off):
// mesh and pressure-field generators declared in hydro_shapes.h, plus their
// C-ABI wrappers (dbh_mesh_*). Set-up time code: nothing here runs per query.
#include "hydro_shapes.h"

#include <algorithm>
#include <array>
#include <cmath>
#include <limits>
#include <map>
#include <stdexcept>
#include <string>
#include <unordered_map>

#include "../../include/drake_b200_hydro.h"

namespace dbh {
namespace {

struct P3 {
  double x, y, z;
};
inline P3 sub(P3 a, P3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline P3 crs(P3 a, P3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline double dt(P3 a, P3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline double len(P3 a) { return std::sqrt(dt(a, a)); }

// Growing unit-sphere tessellation: vertices, tets and the on-surface flag.
struct SphereTess {
  std::vector<P3> v;
  std::vector<std::array<int, 4>> t;
  std::vector<uint8_t> on_surface;
};

inline uint64_t EdgeKey(int a, int b) {
  const uint32_t lo = static_cast<uint32_t>(std::min(a, b)), hi = static_cast<uint32_t>(std::max(a, b));
  return (static_cast<uint64_t>(lo) << 32) | hi;
}

// Midpoint vertex of edge (a, b), created on first use; surface edges are
// pushed back onto the unit sphere (make_sphere_mesh.h:308-328,384-404).
struct EdgeSplitter {
  SphereTess* m;
  const std::vector<uint8_t>* parent_flag;  // null: every split vertex is a surface vertex
  std::unordered_map<uint64_t, int> mid;
  int Get(int a, int b) {
    const uint64_t key = EdgeKey(a, b);
    auto it = mid.find(key);
    if (it != mid.end()) return it->second;
    const int lo = std::min(a, b), hi = std::max(a, b);
    const P3 pa = m->v[lo], pb = m->v[hi];
    P3 p{(pa.x + pb.x) / 2.0, (pa.y + pb.y) / 2.0, (pa.z + pb.z) / 2.0};
    const bool surf = parent_flag ? ((*parent_flag)[lo] && (*parent_flag)[hi]) : true;
    if (surf) {
      const double n = len(p);
      p = {p.x / n, p.y / n, p.z / n};
    }
    const int idx = static_cast<int>(m->v.size());
    m->v.push_back(p);
    m->on_surface.push_back(surf ? 1 : 0);
    mid.emplace(key, idx);
    return idx;
  }
};

// Quality-driven choice of the inner octahedron's diagonal (make_sphere_mesh.h:101-207):
// maximise sum over the four tets of volume^2 / (sum of squared edge lengths)^3.
void AppendBestOctahedronSplit(const int o[6], const std::vector<P3>& v, std::vector<std::array<int, 4>>* out) {
  static const int kSplits[3][4][4] = {
      {{2, 3, 4, 0}, {2, 1, 3, 0}, {2, 4, 3, 5}, {2, 3, 1, 5}},   // diagonal g-h
      {{1, 4, 0, 3}, {1, 4, 3, 5}, {1, 4, 5, 2}, {1, 0, 4, 2}},   // diagonal f-i
      {{0, 3, 5, 4}, {0, 4, 5, 2}, {0, 2, 5, 1}, {0, 1, 5, 3}}};  // diagonal e-j
  auto sq = [&](int a, int b) {
    const P3 d = sub(v[a], v[b]);
    return dt(d, d);
  };
  int best = -1;
  double best_q = -1.0;
  for (int s = 0; s < 3; ++s) {
    double q = 0;
    for (int k = 0; k < 4; ++k) {
      const int a = o[kSplits[s][k][0]], b = o[kSplits[s][k][1]], c = o[kSplits[s][k][2]], d = o[kSplits[s][k][3]];
      const double vol = dt(crs(sub(v[b], v[a]), sub(v[c], v[a])), sub(v[d], v[a])) / 6.0;
      double denom = 0.0;
      denom += sq(a, b);
      denom += sq(a, c);
      denom += sq(a, d);
      denom += sq(b, c);
      denom += sq(b, d);
      denom += sq(c, d);
      q += vol * vol / (denom * denom * denom);
    }
    if (q > best_q) {
      best_q = q;
      best = s;
    }
  }
  if (!(best_q > 0.0)) throw std::logic_error("sphere refinement produced a degenerate octahedron");
  for (int k = 0; k < 4; ++k)
    out->push_back({o[kSplits[best][k][0]], o[kSplits[best][k][1]], o[kSplits[best][k][2]], o[kSplits[best][k][3]]});
}

SphereTess Octahedron() {
  SphereTess m;
  m.v = {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {-1, 0, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};
  m.t = {{1, 5, 2, 0}, {2, 5, 3, 0}, {3, 5, 4, 0}, {4, 5, 1, 0}, {2, 6, 1, 0}, {3, 6, 2, 0}, {4, 6, 3, 0}, {1, 6, 4, 0}};
  m.on_surface = {0, 1, 1, 1, 1, 1, 1};
  return m;
}

// One x8 refinement with interior vertices (make_sphere_mesh.h:294-341).
SphereTess RefineDense(const SphereTess& in) {
  SphereTess out;
  out.v = in.v;
  out.on_surface = in.on_surface;
  EdgeSplitter split{&out, &in.on_surface, {}};
  // First pass creates the split vertices in the reference's order (tets in
  // order, edges 01 02 03 12 13 23), second pass emits the tets.
  for (const auto& t : in.t) {
    split.Get(t[0], t[1]);
    split.Get(t[0], t[2]);
    split.Get(t[0], t[3]);
    split.Get(t[1], t[2]);
    split.Get(t[1], t[3]);
    split.Get(t[2], t[3]);
  }
  out.t.reserve(in.t.size() * 8);
  for (const auto& t : in.t) {
    const int a = t[0], b = t[1], c = t[2], d = t[3];
    const int e = split.Get(a, b), f = split.Get(a, c), g = split.Get(a, d);
    const int h = split.Get(b, c), i = split.Get(b, d), j = split.Get(c, d);
    out.t.push_back({a, e, f, g});
    out.t.push_back({b, h, e, i});
    out.t.push_back({f, h, c, j});
    out.t.push_back({j, g, i, d});
    const int o[6] = {e, f, g, h, i, j};
    AppendBestOctahedronSplit(o, out.v, &out.t);
  }
  return out;
}

// One x4 refinement of the surface fan around the single centre vertex 0
// (make_sphere_mesh.h:370-445).
SphereTess RefineSparse(const SphereTess& in) {
  SphereTess out;
  out.v = in.v;
  out.on_surface = in.on_surface;
  EdgeSplitter split{&out, nullptr, {}};
  for (const auto& t : in.t) {
    if (t[3] != 0) throw std::logic_error("sparse sphere: centre vertex must be the 4th vertex");
    split.Get(t[0], t[1]);
    split.Get(t[0], t[2]);
    split.Get(t[1], t[2]);
  }
  out.t.reserve(in.t.size() * 4);
  for (const auto& t : in.t) {
    const int a = t[0], b = t[1], c = t[2];
    const int d = split.Get(b, c), e = split.Get(a, b), f = split.Get(a, c);
    out.t.push_back({a, e, f, 0});
    out.t.push_back({b, d, e, 0});
    out.t.push_back({c, f, d, 0});
    out.t.push_back({e, d, f, 0});
  }
  return out;
}

HostMesh FromTables(const std::vector<std::array<int, 4>>& t, const std::vector<P3>& v) {
  HostMesh m;
  m.verts_per_elem = 4;
  m.verts.reserve(v.size() * 3);
  for (const P3& p : v) {
    m.verts.push_back(p.x);
    m.verts.push_back(p.y);
    m.verts.push_back(p.z);
  }
  m.elems.reserve(t.size() * 4);
  for (const auto& e : t)
    for (int k = 0; k < 4; ++k) m.elems.push_back(e[k]);
  return m;
}

double RelTol(double size) { return 1e-14 * std::max(1.0, size); }  // proximity_utilities.h:29-31

// Medial-axis box tessellations (make_box_mesh.cc:69-471): connectivity tables.
const std::vector<std::array<int, 4>> kPointMaTets = {
    {8, 0, 4, 11}, {3, 8, 1, 9},  {0, 8, 4, 13}, {8, 7, 3, 14}, {8, 5, 4, 10}, {8, 0, 1, 9},  {5, 8, 1, 14}, {0, 8, 2, 9},
    {2, 8, 3, 9},  {8, 4, 6, 10}, {7, 8, 6, 10}, {8, 5, 1, 11}, {4, 8, 6, 13}, {8, 5, 7, 14}, {8, 7, 6, 12}, {8, 6, 2, 12},
    {5, 8, 7, 10}, {5, 8, 4, 11}, {8, 3, 1, 14}, {8, 2, 3, 12}, {7, 8, 3, 12}, {8, 0, 2, 13}, {6, 8, 2, 13}, {0, 8, 1, 11}};
const std::vector<std::array<int, 4>> kLineMaTets = {
    {5, 10, 4, 13}, {3, 1, 8, 16},  {7, 8, 5, 16},  {6, 9, 4, 12},  {8, 1, 5, 16},  {2, 10, 3, 11}, {6, 7, 10, 12},
    {10, 5, 4, 12}, {10, 0, 1, 11}, {7, 6, 10, 14}, {8, 7, 5, 12},  {3, 8, 7, 16},  {6, 9, 2, 15},  {9, 0, 2, 15},
    {1, 8, 10, 11}, {1, 3, 8, 11},  {8, 3, 10, 11}, {2, 9, 10, 11}, {0, 9, 2, 11},  {10, 9, 0, 11}, {8, 3, 7, 14},
    {8, 7, 10, 14}, {9, 6, 2, 14},  {10, 6, 9, 14}, {10, 2, 3, 14}, {3, 8, 10, 14}, {9, 2, 10, 14}, {9, 4, 0, 15},
    {9, 6, 4, 15},  {10, 9, 4, 13}, {4, 9, 0, 13},  {8, 10, 5, 13}, {1, 8, 5, 13},  {0, 10, 1, 13}, {9, 10, 0, 13},
    {8, 1, 10, 13}, {10, 8, 5, 12}, {7, 8, 10, 12}, {9, 10, 4, 12}, {6, 10, 9, 12}};
const std::vector<std::array<int, 4>> kRectMaTets = {
    {4, 11, 6, 16},  {8, 5, 1, 14},   {5, 9, 8, 17},   {5, 4, 8, 13},   {5, 8, 1, 17},   {0, 11, 2, 12}, {7, 9, 6, 13},
    {3, 7, 9, 15},   {3, 2, 9, 12},   {9, 7, 6, 15},   {1, 9, 3, 17},   {6, 11, 2, 16},  {2, 11, 9, 12}, {9, 8, 1, 12},
    {9, 5, 8, 13},   {5, 9, 7, 13},   {8, 9, 11, 12},  {8, 0, 1, 12},   {11, 4, 6, 13},  {9, 1, 3, 12},  {4, 5, 8, 14},
    {7, 3, 9, 17},   {9, 5, 7, 17},   {8, 11, 10, 12}, {0, 8, 10, 12},  {11, 0, 10, 12}, {8, 9, 1, 17},  {11, 6, 2, 15},
    {11, 9, 6, 15},  {2, 3, 9, 15},   {11, 2, 9, 15},  {4, 0, 10, 16},  {11, 4, 10, 16}, {11, 0, 2, 16}, {0, 11, 10, 16},
    {4, 8, 10, 14},  {0, 4, 10, 14},  {0, 8, 1, 14},   {8, 0, 10, 14},  {9, 11, 6, 13},  {9, 8, 11, 13}, {8, 4, 10, 13},
    {11, 8, 10, 13}, {4, 11, 10, 13}};

}  // namespace

int SphereRefinementLevel(double radius, double resolution_hint) {
  // make_sphere_mesh.h:576-590: equatorial chords no longer than the hint, capped at level 8.
  const double edge = std::min(resolution_hint, 2.0 * radius);
  const int L = std::max(0, static_cast<int>(std::ceil(std::log2(M_PI / std::asin(edge / (2.0 * radius)))) - 2));
  return std::min(L, 8);
}

HostMesh MakeSphereVolumeMesh(double radius, double resolution_hint, int strategy) {
  if (!(resolution_hint > 0.0) || !(radius > 0.0)) throw std::invalid_argument("sphere: radius and hint must be > 0");
  if (strategy != kSingleInteriorVertex && strategy != kDenseInteriorVertices)
    throw std::invalid_argument("sphere: unknown tessellation strategy");  // tessellation_strategy.h has these two only
  SphereTess m = Octahedron();
  const int L = SphereRefinementLevel(radius, resolution_hint);
  for (int l = 0; l < L; ++l) m = (strategy == kDenseInteriorVertices) ? RefineDense(m) : RefineSparse(m);
  for (P3& p : m.v) p = {p.x * radius, p.y * radius, p.z * radius};
  return FromTables(m.t, m.v);
}

HostMesh MakeEllipsoidVolumeMesh(double a, double b, double c, double resolution_hint, int strategy) {
  if (!(resolution_hint > 0.0)) throw std::invalid_argument("ellipsoid: hint must be > 0");
  if (!(a > 0.0) || !(b > 0.0) || !(c > 0.0)) throw std::invalid_argument("ellipsoid: the semi-axes must be > 0");  // Ellipsoid's constructor
  const double r = std::max({a, b, c});
  HostMesh m = MakeSphereVolumeMesh(1.0, resolution_hint / r, strategy);
  for (int i = 0; i < m.num_vertices(); ++i) {
    m.verts[3 * i] = a * m.verts[3 * i];
    m.verts[3 * i + 1] = b * m.verts[3 * i + 1];
    m.verts[3 * i + 2] = c * m.verts[3 * i + 2];
  }
  return m;
}

HostMesh MakeBoxVolumeMeshWithMa(double sx, double sy, double sz) {
  if (!(sx > 0.0) || !(sy > 0.0) || !(sz > 0.0)) throw std::invalid_argument("box: the sizes must be > 0");  // Box's constructor
  const double h[3] = {sx / 2., sy / 2., sz / 2.};
  const double hmin = std::min({h[0], h[1], h[2]});
  int zero_axes = 0;
  for (int k = 0; k < 3; ++k) zero_axes += !((h[k] - hmin) > RelTol(hmin));
  std::vector<P3> v;
  if (zero_axes == 3) {  // the medial axis is a point
    const double L = sx / 2.0;
    v = {{-L, -L, -L}, {-L, -L, L}, {-L, L, -L}, {-L, L, L}, {L, -L, -L}, {L, -L, L}, {L, L, -L}, {L, L, L},
         {0., 0., 0.}, {-L, 0., 0.}, {L, 0., 0.}, {0., -L, 0.}, {0., L, 0.}, {0., 0., -L}, {0., 0., L}};
    return FromTables(kPointMaTets, v);
  }
  if (zero_axes == 2) {  // a line: canonical frame has the long side along z
    int zd = 0;
    for (int k = 1; k < 3; ++k)
      if (h[k] > h[zd]) zd = k;
    const double L2 = h[zd], L1 = h[(zd + 1) % 3], Lm = L2 - L1;
    v = {{-L1, -L1, -L2}, {-L1, -L1, L2}, {-L1, L1, -L2}, {-L1, L1, L2}, {L1, -L1, -L2}, {L1, -L1, L2},
         {L1, L1, -L2},   {L1, L1, L2},   {0, 0, Lm},     {0, 0, -Lm},   {0, 0, 0},      {-L1, 0, 0},
         {L1, 0, 0},      {0, -L1, 0},    {0, L1, 0},     {0, 0, -L2},   {0, 0, L2}};
    if (zd == 0)
      for (P3& p : v) p = {-p.z, p.y, p.x};
    if (zd == 1)
      for (P3& p : v) p = {p.x, -p.z, p.y};
    return FromTables(kLineMaTets, v);
  }
  if (zero_axes != 1) throw std::logic_error("box medial-axis classification failed");
  int xd = 0;  // a rectangle: canonical frame has the short side along x
  for (int k = 1; k < 3; ++k)
    if (h[k] < h[xd]) xd = k;
  const double L1 = h[xd];
  double L2 = h[1], L3 = h[2];
  if (xd == 2) L3 = h[0];
  if (xd == 1) L2 = h[0];
  const double m2 = L2 - L1, m3 = L3 - L1;
  v = {{-L1, -L2, -L3}, {-L1, -L2, L3}, {-L1, L2, -L3}, {-L1, L2, L3}, {L1, -L2, -L3}, {L1, -L2, L3},
       {L1, L2, -L3},   {L1, L2, L3},   {0, -m2, m3},   {0, m2, m3},   {0, -m2, -m3},  {0, m2, -m3},
       {-L1, 0, 0},     {L1, 0, 0},     {0, -L2, 0},    {0, L2, 0},    {0, 0, -L3},    {0, 0, L3}};
  if (xd == 2)
    for (P3& p : v) p = {-p.z, p.y, p.x};
  if (xd == 1)
    for (P3& p : v) p = {-p.y, p.x, p.z};
  return FromTables(kRectMaTets, v);
}

HostMesh MakeBoxVolumeMesh(double sx, double sy, double sz, double hint) {
  if (!(hint > 0.)) throw std::invalid_argument("box: hint must be > 0");
  if (!(sx > 0.0) || !(sy > 0.0) || !(sz > 0.0)) throw std::invalid_argument("box: the sizes must be > 0");
  const double size[3] = {sx, sy, sz};
  int n[3];
  std::vector<double> coord[3];
  for (int k = 0; k < 3; ++k) {
    n[k] = 1 + static_cast<int>(std::ceil(size[k] / hint));
    // Eigen::VectorXd::LinSpaced(n, -h, h).
    const double lo = -size[k] / 2, hi = size[k] / 2;
    coord[k].resize(n[k]);
    const double step = n[k] > 1 ? (hi - lo) / (n[k] - 1) : 0.0;
    for (int i = 0; i < n[k]; ++i) coord[k][i] = lo + i * step;
    coord[k][n[k] - 1] = hi;
  }
  HostMesh m;
  m.verts_per_elem = 4;
  m.verts.reserve(static_cast<size_t>(3) * n[0] * n[1] * n[2]);
  for (int i = 0; i < n[0]; ++i)
    for (int j = 0; j < n[1]; ++j)
      for (int k = 0; k < n[2]; ++k) {
        m.verts.push_back(coord[0][i]);
        m.verts.push_back(coord[1][j]);
        m.verts.push_back(coord[2][k]);
      }
  auto id = [&](int i, int j, int k) { return (i * n[1] + j) * n[2] + k; };
  // Six tets around the cell diagonal v0-v7 (make_box_mesh.cc:805-860).
  static const int kCycle[7] = {4, 6, 2, 3, 1, 5, 4};
  for (int i = 0; i + 1 < n[0]; ++i)
    for (int j = 0; j + 1 < n[1]; ++j)
      for (int k = 0; k + 1 < n[2]; ++k) {
        int c[8];
        for (int s = 0; s < 8; ++s) c[s] = id(i + ((s >> 2) & 1), j + ((s >> 1) & 1), k + (s & 1));
        for (int t = 0; t < 6; ++t) {
          m.elems.push_back(c[0]);
          m.elems.push_back(c[7]);
          m.elems.push_back(c[kCycle[t]]);
          m.elems.push_back(c[kCycle[t + 1]]);
        }
      }
  return m;
}

HostMesh ConvertVolumeToSurfaceMesh(const HostMesh& vol) {
  // Boundary faces are those seen exactly once; emitted in the order of their
  // sorted vertex triple, outward winding (volume_to_surface_mesh.cc:31-134).
  static const int kFace[4][3] = {{1, 2, 3}, {3, 2, 0}, {1, 3, 0}, {2, 1, 0}};
  std::map<std::array<int, 3>, std::array<int, 3>> open;
  for (int t = 0; t < vol.num_elements(); ++t) {
    const int32_t* e = &vol.elems[4 * t];
    for (const auto& f : kFace) {
      const std::array<int, 3> wound{e[f[0]], e[f[1]], e[f[2]]};
      std::array<int, 3> key = wound;
      std::sort(key.begin(), key.end());
      auto it = open.find(key);
      if (it == open.end()) {
        open.emplace(key, wound);
      } else {
        open.erase(it);
      }
    }
  }
  std::vector<int> used;
  for (const auto& kv : open)
    for (int v : kv.second) used.push_back(v);
  std::sort(used.begin(), used.end());
  used.erase(std::unique(used.begin(), used.end()), used.end());
  HostMesh s;
  s.verts_per_elem = 3;
  for (int v : used)
    for (int k = 0; k < 3; ++k) s.verts.push_back(vol.verts[3 * v + k]);
  for (const auto& kv : open)
    for (int v : kv.second)
      s.elems.push_back(static_cast<int>(std::lower_bound(used.begin(), used.end(), v) - used.begin()));
  return s;
}

void SetSpherePressureField(HostMesh* m, double radius, double modulus, double margin) {
  if (!(modulus > 0) || !(radius > margin)) throw std::invalid_argument("sphere field: bad parameters");
  const double min_extent = -margin / (radius - margin);
  m->pressure.resize(m->num_vertices());
  for (int i = 0; i < m->num_vertices(); ++i) {
    const P3 r{m->verts[3 * i], m->verts[3 * i + 1], m->verts[3 * i + 2]};
    const double distance = radius - len(r);
    double extent = (distance - margin) / (radius - margin);
    if (std::abs(distance / radius) < 1e-14) extent = min_extent;
    m->pressure[i] = modulus * extent;
  }
}

void SetEllipsoidPressureField(HostMesh* m, double a, double b, double c, double modulus, double margin) {
  if (!(modulus > 0) || !(a > margin) || !(b > margin) || !(c > margin))
    throw std::invalid_argument("ellipsoid field: bad parameters");
  const double sx = 1.0 / (a - margin), sy = 1.0 / (b - margin), sz = 1.0 / (c - margin);
  m->pressure.resize(m->num_vertices());
  for (int i = 0; i < m->num_vertices(); ++i) {
    const P3 u{sx * m->verts[3 * i], sy * m->verts[3 * i + 1], sz * m->verts[3 * i + 2]};
    double extent = 1.0 - len(u);
    if (margin == 0 && std::abs(extent) < 1e-14) extent = 0.0;
    m->pressure[i] = modulus * extent;
  }
}

void SetBoxPressureField(HostMesh* m, double sx, double sy, double sz, double modulus, double margin) {
  const double h[3] = {sx / 2.0, sy / 2.0, sz / 2.0};
  const double hmin = std::min({h[0], h[1], h[2]});
  if (!(modulus > 0) || !(hmin > margin)) throw std::invalid_argument("box field: bad parameters");
  m->pressure.resize(m->num_vertices());
  for (int i = 0; i < m->num_vertices(); ++i) {
    const double q[3] = {m->verts[3 * i], m->verts[3 * i + 1], m->verts[3 * i + 2]};
    // Signed distance to the box = grad . (q - nearest) with the classification of
    // DistanceToPoint::ComputeDistanceToBox (distance_to_point_callback.cc:405-494).
    double nearest[3], grad[3] = {0, 0, 0};
    int cls[3];  // 0 inside, 1 boundary, 2 outside
    for (int k = 0; k < 3; ++k) {
      const double tol = RelTol(h[k]);
      const double sgn = q[k] < 0.0 ? -1.0 : 1.0;
      if (std::abs(q[k]) > h[k] + tol) {
        cls[k] = 2;
        nearest[k] = sgn * h[k];
      } else if (std::abs(q[k]) >= h[k] - tol) {
        cls[k] = 1;
        nearest[k] = sgn * h[k];
      } else {
        cls[k] = 0;
        nearest[k] = q[k];
      }
    }
    if (cls[0] == 2 || cls[1] == 2 || cls[2] == 2) {
      const P3 d{q[0] - nearest[0], q[1] - nearest[1], q[2] - nearest[2]};
      const double n = len(d);
      grad[0] = d.x / n;
      grad[1] = d.y / n;
      grad[2] = d.z / n;
    } else if (cls[0] == 1 || cls[1] == 1 || cls[2] == 1) {
      for (int k = 0; k < 3; ++k)
        if (cls[k] == 1) grad[k] = nearest[k] < 0.0 ? -1.0 : 1.0;
      const double n = len(P3{grad[0], grad[1], grad[2]});
      for (int k = 0; k < 3; ++k) grad[k] = grad[k] / n;
    } else {
      double best = std::numeric_limits<double>::infinity();
      int axis = -1;
      for (int k = 0; k < 3; ++k)
        for (const double bound : {h[k], -h[k]}) {
          const double dist = std::abs(bound - q[k]);
          if (dist < best) {
            best = dist;
            axis = k;
          }
        }
      const double sgn = q[axis] < 0.0 ? -1.0 : 1.0;
      nearest[axis] = sgn * h[axis];
      grad[axis] = sgn;
    }
    const double sd = dt(P3{grad[0], grad[1], grad[2]}, P3{q[0] - nearest[0], q[1] - nearest[1], q[2] - nearest[2]});
    m->pressure[i] = modulus * ((-sd - margin) / (hmin - margin));
  }
}


// ---- Cylinder and capsule (make_cylinder_mesh.cc, make_capsule_mesh.cc, meshing_utilities.cc) ------------------------
namespace {

using Tet = std::array<int, 4>;

// SplitTriangularPrismToTetrahedra (meshing_utilities.cc:21-33): v0 v1 v2 the bottom triangle, v3 v4 v5 above them; the
// three tetrahedra walk the path v3 -> v4 -> v1 -> v2 and all hold the diagonal v0-v5.
void AppendPrism(int v0, int v1, int v2, int v3, int v4, int v5, std::vector<Tet>* out) {
  const int path[4] = {v3, v4, v1, v2};
  for (int k = 0; k < 3; ++k) out->push_back({path[k], path[k + 1], v0, v5});
}
// SplitPyramidToTetrahedra (meshing_utilities.cc:35-46): base v0 v1 v2 v3, apex v4; both tetrahedra hold v0-v2.
void AppendPyramid(int v0, int v1, int v2, int v3, int v4, std::vector<Tet>* out) {
  out->push_back({v3, v4, v0, v2});
  out->push_back({v4, v1, v0, v2});
}

// The signed distance of DistanceToPoint::operator()(Cylinder) at the identity pose (distance_to_point_callback.cc:
// 194-288 over ComputeDistanceToBox<2>, :406-482) for a point inside or on the cylinder; mesh vertices are nowhere else.
double CylinderSignedDistance(double radius, double length, P3 q) {
  const double r_q = std::sqrt(q.x * q.x + q.y * q.y);
  const bool on_axis = r_q < RelTol(radius);
  const double vx = on_axis ? 1.0 : q.x / r_q, vy = on_axis ? 0.0 : q.y / r_q;
  const double h[2] = {radius, length / 2.0}, p[2] = {r_q, q.z};
  double nearest[2], grad[2] = {0.0, 0.0};
  int cls[2];  // 0 inside, 1 boundary, 2 outside
  for (int k = 0; k < 2; ++k) {
    const double tol = RelTol(h[k]);
    const double sgn = p[k] < 0.0 ? -1.0 : 1.0;
    cls[k] = std::abs(p[k]) > h[k] + tol ? 2 : (std::abs(p[k]) >= h[k] - tol ? 1 : 0);
    nearest[k] = cls[k] ? sgn * h[k] : p[k];
  }
  if (cls[0] == 2 || cls[1] == 2) {
    const double dx = p[0] - nearest[0], dy = p[1] - nearest[1];
    const double n = std::sqrt(dx * dx + dy * dy);
    grad[0] = dx / n;
    grad[1] = dy / n;
  } else if (cls[0] == 1 || cls[1] == 1) {
    for (int k = 0; k < 2; ++k)
      if (cls[k] == 1) grad[k] = nearest[k] < 0.0 ? -1.0 : 1.0;
    const double n = std::sqrt(grad[0] * grad[0] + grad[1] * grad[1]);
    grad[0] = grad[0] / n;
    grad[1] = grad[1] / n;
  } else {
    double best = std::numeric_limits<double>::infinity();
    int axis = -1;
    for (int k = 0; k < 2; ++k)
      for (const double bound : {h[k], -h[k]}) {
        const double dist = std::abs(bound - p[k]);
        if (dist < best) {
          best = dist;
          axis = k;
        }
      }
    const double sgn = p[axis] < 0.0 ? -1.0 : 1.0;
    nearest[axis] = sgn * h[axis];
    grad[axis] = sgn;
  }
  const P3 n3{nearest[0] * vx, nearest[0] * vy, nearest[1]};
  const P3 g3{grad[0] * vx, grad[0] * vy, grad[1]};
  return dt(g3, sub(q, n3));
}

// CollectUniqueVertices(IdentifyBoundaryFaces(tetrahedra)) (volume_to_surface_mesh.cc:31-150) as a per-vertex flag.
std::vector<char> BoundaryVertexFlags(const HostMesh& vol) {
  static const int kFace[4][3] = {{1, 2, 3}, {3, 2, 0}, {1, 3, 0}, {2, 1, 0}};
  std::map<std::array<int, 3>, int> seen;
  for (int t = 0; t < vol.num_elements(); ++t) {
    const int32_t* e = &vol.elems[4 * t];
    for (const auto& f : kFace) {
      std::array<int, 3> key{e[f[0]], e[f[1]], e[f[2]]};
      std::sort(key.begin(), key.end());
      ++seen[key];
    }
  }
  std::vector<char> flag(vol.num_vertices(), 0);
  for (const auto& kv : seen)
    if (kv.second == 1)
      for (int v : kv.first) flag[v] = 1;
  return flag;
}

}  // namespace

// MakeCylinderVolumeMeshWithMa (make_cylinder_mesh.cc:325-425): the two rim circles (bottom / top vertex interleaved)
// around the two cap centres, then the medial axis -- two points (long), one (medium) or a disc (short) -- and the
// fan of tetrahedra of CalcLong / Medium / ShortCylinderVolumeMeshWithMa (:47-95, :127-165, :198-260).
HostMesh MakeCylinderVolumeMeshWithMa(double radius, double length, double hint) {
  if (!(radius > 0) || !(length > 0) || !(hint > 0)) throw std::invalid_argument("cylinder mesh: bad parameters");
  const double top_z = length / 2.0;
  const double tol = RelTol(std::min(top_z, radius));
  const int kind = top_z - radius > tol ? 0 : (radius - top_z > tol ? 2 : 1);  // long, medium, short
  const int n = std::max(3, static_cast<int>(std::ceil(2.0 * M_PI * radius / hint)));
  std::vector<P3> v;
  v.push_back({0, 0, -top_z});
  v.push_back({0, 0, top_z});
  const int bottom_center = 0, top_center = 1;
  std::vector<int> bottom(n), top(n);
  const double step = 2.0 * M_PI / n;
  for (int i = 0; i < n; ++i) {
    const double x = radius * std::cos(step * i), y = radius * std::sin(step * i);
    bottom[i] = static_cast<int>(v.size());
    v.push_back({x, y, -top_z});
    top[i] = static_cast<int>(v.size());
    v.push_back({x, y, top_z});
  }
  std::vector<Tet> t;
  if (kind == 0) {
    const int m0 = static_cast<int>(v.size());
    v.push_back({0, 0, -(top_z - radius)});
    const int m1 = static_cast<int>(v.size());
    v.push_back({0, 0, top_z - radius});
    for (int j = 0, i = n - 1; j < n; i = j++) {
      t.push_back({bottom_center, bottom[i], bottom[j], m0});
      t.push_back({top_center, top[j], top[i], m1});
      AppendPrism(m0, bottom[i], bottom[j], m1, top[i], top[j], &t);
    }
  } else if (kind == 1) {
    const int m = static_cast<int>(v.size());
    v.push_back({0, 0, 0});
    for (int j = 0, i = n - 1; j < n; i = j++) {
      t.push_back({bottom_center, bottom[i], bottom[j], m});
      t.push_back({top_center, top[j], top[i], m});
      AppendPyramid(top[i], top[j], bottom[j], bottom[i], m, &t);
    }
  } else {
    const int center = static_cast<int>(v.size());
    v.push_back({0, 0, 0});
    std::vector<int> medial(n);
    const double scale = (radius - length / 2.0) / radius;
    for (int i = 0; i < n; ++i) {
      medial[i] = static_cast<int>(v.size());
      v.push_back({v[bottom[i]].x * scale, v[bottom[i]].y * scale, 0});
    }
    for (int j = 0, i = n - 1; j < n; i = j++) {
      AppendPrism(bottom_center, bottom[i], bottom[j], center, medial[i], medial[j], &t);
      AppendPrism(center, medial[i], medial[j], top_center, top[i], top[j], &t);
      AppendPrism(bottom[i], medial[i], top[i], bottom[j], medial[j], top[j], &t);
    }
  }
  return FromTables(t, v);
}

// MakeCapsuleVolumeMesh (make_capsule_mesh.cc:13-180): the two medial end points, the two poles, then per latitude
// circle (equator first) the top- and bottom-cap vertex interleaved; pyramids under the cap quads, one tetrahedron per
// pole wedge, a prism per barrel quad.
HostMesh MakeCapsuleVolumeMesh(double radius, double length, double hint) {
  if (!(radius > 0) || !(length > 0) || !(hint > 0)) throw std::invalid_argument("capsule mesh: bad parameters");
  const double medial_top_z = length / 2.0;
  const double top_z = medial_top_z + radius;
  const int n = static_cast<int>(std::min(std::max(2.0 * M_PI * radius / hint, 3.0), 706.0));
  const int circles = n / 2;
  std::vector<P3> v;
  const int medial_top = 0, medial_bottom = 1, top = 2, bottom = 3;
  v.push_back({0, 0, medial_top_z});
  v.push_back({0, 0, -medial_top_z});
  v.push_back({0, 0, top_z});
  v.push_back({0, 0, -top_z});
  std::vector<int> top_cap(circles * n), bottom_cap(circles * n);
  const double theta_step = M_PI_2 / circles, phi_step = 2 * M_PI / n;
  for (int i = 0; i < circles; ++i) {
    const double theta = M_PI_2 - i * theta_step;
    const double s = std::sin(theta);
    const double top_circle_z = radius * std::cos(theta) + medial_top_z;
    for (int j = 0; j < n; ++j) {
      const double phi = j * phi_step;
      const double x = radius * s * std::cos(phi), y = radius * s * std::sin(phi);
      top_cap[n * i + j] = static_cast<int>(v.size());
      v.push_back({x, y, top_circle_z});
      bottom_cap[n * i + j] = static_cast<int>(v.size());
      v.push_back({x, y, -top_circle_z});
    }
  }
  std::vector<Tet> t;
  for (int i = 0; i < circles - 1; ++i)
    for (int j = 0; j < n; ++j) {
      const int j1 = (j + 1) % n;
      AppendPyramid(top_cap[(i + 1) * n + j], top_cap[(i + 1) * n + j1], top_cap[i * n + j1], top_cap[i * n + j], medial_top, &t);
      AppendPyramid(bottom_cap[i * n + j], bottom_cap[i * n + j1], bottom_cap[(i + 1) * n + j1], bottom_cap[(i + 1) * n + j],
                    medial_bottom, &t);
    }
  const int last = (circles - 1) * n;
  for (int j = 0; j < n; ++j) {
    const int j1 = (j + 1) % n;
    t.push_back({top, top_cap[last + j1], top_cap[last + j], medial_top});
    t.push_back({bottom, bottom_cap[last + j], bottom_cap[last + j1], medial_bottom});
    AppendPrism(medial_bottom, bottom_cap[j], bottom_cap[j1], medial_top, top_cap[j], top_cap[j1], &t);
  }
  return FromTables(t, v);
}

// MakeCylinderPressureField (make_cylinder_field.cc:18-75): extent = (-signed distance - margin) / (min(r, L/2) - margin)
// capped at the modulus; every boundary vertex gets the (negative, for margin > 0) surface pressure exactly.
void SetCylinderPressureField(HostMesh* m, double radius, double length, double modulus, double margin) {
  if (!(modulus > 0) || !(radius > margin) || !(length > 2 * margin)) throw std::invalid_argument("cylinder field: bad parameters");
  const double min_half_size = std::min(radius, length / 2.0);
  m->pressure.resize(m->num_vertices());
  for (int i = 0; i < m->num_vertices(); ++i) {
    const double sd = CylinderSignedDistance(radius, length, {m->verts[3 * i], m->verts[3 * i + 1], m->verts[3 * i + 2]});
    const double extent = (-sd - margin) / (min_half_size - margin);
    m->pressure[i] = std::min(modulus * extent, modulus);
  }
  const double min_pressure = -modulus * margin / (min_half_size - margin);
  const std::vector<char> boundary = BoundaryVertexFlags(*m);
  for (int i = 0; i < m->num_vertices(); ++i)
    if (boundary[i]) m->pressure[i] = min_pressure;
}

// MakeCapsulePressureField (make_capsule_field.h:55-90): the modulus on the two medial end points (vertices 0 and 1 of
// MakeCapsuleVolumeMesh), the surface pressure -margin * E / (r - margin) everywhere else.
void SetCapsulePressureField(HostMesh* m, double radius, double length, double modulus, double margin) {
  if (!(modulus > 0) || !(radius > margin)) throw std::invalid_argument("capsule field: bad parameters");
  if (m->num_vertices() < 2 || m->verts[0] != 0 || m->verts[1] != 0 || m->verts[2] != length / 2 || m->verts[3] != 0 ||
      m->verts[4] != 0 || m->verts[5] != -length / 2)
    throw std::invalid_argument("capsule field: the first two vertices must be the ends of the medial axis");
  const double pressure_gradient = modulus / (radius - margin);
  const double surface_pressure = -margin * pressure_gradient;
  m->pressure.assign(m->num_vertices(), surface_pressure);
  m->pressure[0] = m->pressure[1] = modulus;
}


// ---- Convex / Mesh: MakeConvexVolumeMesh (make_convex_mesh.cc:17-60), MakeVolumeMeshPressureField (make_mesh_field.cc:
//      45-96) over CalcDistanceToSurfaceMesh (calc_distance_to_surface_mesh.cc:30-120; its BVH only prunes: the result is
//      the minimum over the boundary triangles, taken here directly) ----------------------------------------------------
namespace {

double SquaredDistanceToSegment(P3 q, P3 a, P3 b) {
  const P3 ab = sub(b, a), aq = sub(q, a);
  if (dt(aq, ab) <= 0.0) return dt(aq, aq);
  const P3 bq = sub(q, b);
  if (dt(bq, ab) >= 0.0) return dt(bq, bq);
  const P3 c = crs(ab, aq);
  return dt(c, c) / dt(ab, ab);
}

// In the triangle's frame T (origin A, x along AB, z the normal): inside the triangle's prism the distance is |z|, else
// it is that to the nearest edge.
double SquaredDistanceToTriangle(P3 q, P3 a, P3 b, P3 c) {
  const P3 ab = sub(b, a), ac = sub(c, a);
  const double lab = len(ab);
  const P3 tx{ab.x / lab, ab.y / lab, ab.z / lab};
  const P3 n = crs(ab, ac);
  const double ln = len(n);
  const P3 tz{n.x / ln, n.y / ln, n.z / ln};
  const P3 ty = crs(tz, tx);
  const P3 aq = sub(q, a);
  const P3 tb{dt(tx, ab), dt(ty, ab), dt(tz, ab)}, tc{dt(tx, ac), dt(ty, ac), dt(tz, ac)}, tq{dt(tx, aq), dt(ty, aq), dt(tz, aq)};
  const double det = tb.x * tc.y - tc.x * tb.y;
  const double b0 = (tq.x * tc.y - tc.x * tq.y) / det, b1 = (tb.x * tq.y - tq.x * tb.y) / det;
  if (b0 >= 0.0 && b1 >= 0.0 && b0 + b1 <= 1.0) return tq.z * tq.z;
  const P3 o{0, 0, 0};
  return std::min({SquaredDistanceToSegment(tq, o, tb), SquaredDistanceToSegment(tq, o, tc), SquaredDistanceToSegment(tq, tb, tc)});
}

}  // namespace

// The surface's vertices, then the centroid of the enclosed volume; one tetrahedron {centroid, v0, v1, v2} per triangle.
HostMesh MakeConvexVolumeMesh(const HostMesh& surface) {
  if (surface.verts_per_elem != 3 || surface.num_elements() < 4) throw std::invalid_argument("convex mesh: not a closed surface");
  double six_total_volume = 0;
  P3 centroid{0, 0, 0};
  auto vertex = [&](int v) { return P3{surface.verts[3 * v], surface.verts[3 * v + 1], surface.verts[3 * v + 2]}; };
  for (int t = 0; t < surface.num_elements(); ++t) {
    const P3 p = vertex(surface.elems[3 * t]), q = vertex(surface.elems[3 * t + 1]), r = vertex(surface.elems[3 * t + 2]);
    const double six_volume = dt(crs(p, q), r);
    six_total_volume += six_volume;
    centroid = {centroid.x + six_volume * ((p.x + q.x) + r.x), centroid.y + six_volume * ((p.y + q.y) + r.y),
                centroid.z + six_volume * ((p.z + q.z) + r.z)};
  }
  if (!(six_total_volume > 0)) throw std::invalid_argument("convex mesh: the surface does not enclose a positive volume");
  const double k = 4 * six_total_volume;
  HostMesh m;
  m.verts_per_elem = 4;
  m.verts = surface.verts;
  m.verts.push_back(centroid.x / k);
  m.verts.push_back(centroid.y / k);
  m.verts.push_back(centroid.z / k);
  const int c = surface.num_vertices();
  for (int t = 0; t < surface.num_elements(); ++t)
    for (int v : {c, surface.elems[3 * t], surface.elems[3 * t + 1], surface.elems[3 * t + 2]}) m.elems.push_back(v);
  return m;
}

// Distance of every interior vertex to the boundary surface (+ margin), scaled so that the deepest vertex carries the
// modulus; boundary vertices sit at -E margin / (max - margin) (0 without margin).
void SetVolumeMeshPressureField(HostMesh* m, double modulus, double margin) {
  if (m->verts_per_elem != 4 || !(modulus > 0) || !(margin >= 0)) throw std::invalid_argument("mesh field: bad parameters");
  const HostMesh surface = ConvertVolumeToSurfaceMesh(*m);
  const std::vector<char> boundary = BoundaryVertexFlags(*m);
  auto sv = [&](int v) { return P3{surface.verts[3 * v], surface.verts[3 * v + 1], surface.verts[3 * v + 2]}; };
  std::vector<double> values(m->num_vertices(), 0.0);
  double max_value = std::numeric_limits<double>::lowest();
  for (int v = 0; v < m->num_vertices(); ++v) {
    if (boundary[v]) continue;
    const P3 q{m->verts[3 * v], m->verts[3 * v + 1], m->verts[3 * v + 2]};
    double min_squared = std::numeric_limits<double>::infinity();
    for (int t = 0; t < surface.num_elements(); ++t)
      min_squared = std::min(min_squared, SquaredDistanceToTriangle(q, sv(surface.elems[3 * t]), sv(surface.elems[3 * t + 1]),
                                                                    sv(surface.elems[3 * t + 2])));
    values[v] = std::sqrt(min_squared) + margin;
    max_value = std::max(values[v], max_value);
  }
  if (!(max_value > 0))
    throw std::invalid_argument("MakeVolumeMeshPressureField(): the computed max distance to boundary among all mesh vertices "
                                "is non-positive. Perhaps the mesh lacks interior vertices.");
  m->pressure.resize(values.size());
  for (size_t v = 0; v < values.size(); ++v) m->pressure[v] = modulus * (values[v] - margin) / (max_value - margin);
}

}  // namespace dbh

// ---------------------------------------------------------------------------
// C-ABI wrappers.
struct dbh_mesh {
  dbh::HostMesh m;
};

namespace {
thread_local std::string g_shape_error;
template <class F>
dbh_status Guard(dbh_mesh** out, F&& make) {
  if (!out) return DBH_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  try {
    auto p = new dbh_mesh{make()};
    *out = p;
    return DBH_OK;
  } catch (const std::exception& e) {
    g_shape_error = e.what();
    return DBH_ERR_INVALID_ARGUMENT;
  }
}
}  // namespace

extern "C" {

const char* dbh_mesh_last_error(void) { return g_shape_error.c_str(); }

dbh_status dbh_mesh_make_sphere_volume(double radius, double resolution_hint, int32_t strategy, dbh_mesh** out) {
  return Guard(out, [&] { return dbh::MakeSphereVolumeMesh(radius, resolution_hint, strategy); });
}
dbh_status dbh_mesh_make_ellipsoid_volume(double a, double b, double c, double resolution_hint, int32_t strategy,
                                          dbh_mesh** out) {
  return Guard(out, [&] { return dbh::MakeEllipsoidVolumeMesh(a, b, c, resolution_hint, strategy); });
}
dbh_status dbh_mesh_make_box_volume_with_ma(double x, double y, double z, dbh_mesh** out) {
  return Guard(out, [&] { return dbh::MakeBoxVolumeMeshWithMa(x, y, z); });
}
dbh_status dbh_mesh_make_box_volume(double x, double y, double z, double resolution_hint, dbh_mesh** out) {
  return Guard(out, [&] { return dbh::MakeBoxVolumeMesh(x, y, z, resolution_hint); });
}
dbh_status dbh_mesh_convert_volume_to_surface(const dbh_mesh* volume, dbh_mesh** out) {
  if (!volume || volume->m.verts_per_elem != 4) return DBH_ERR_INVALID_ARGUMENT;
  return Guard(out, [&] { return dbh::ConvertVolumeToSurfaceMesh(volume->m); });
}
dbh_status dbh_mesh_set_sphere_pressure(dbh_mesh* mesh, double radius, double modulus, double margin) {
  if (!mesh) return DBH_ERR_INVALID_ARGUMENT;
  try {
    dbh::SetSpherePressureField(&mesh->m, radius, modulus, margin);
  } catch (const std::exception& e) {
    g_shape_error = e.what();
    return DBH_ERR_INVALID_ARGUMENT;
  }
  return DBH_OK;
}
dbh_status dbh_mesh_set_ellipsoid_pressure(dbh_mesh* mesh, double a, double b, double c, double modulus,
                                           double margin) {
  if (!mesh) return DBH_ERR_INVALID_ARGUMENT;
  try {
    dbh::SetEllipsoidPressureField(&mesh->m, a, b, c, modulus, margin);
  } catch (const std::exception& e) {
    g_shape_error = e.what();
    return DBH_ERR_INVALID_ARGUMENT;
  }
  return DBH_OK;
}
dbh_status dbh_mesh_set_box_pressure(dbh_mesh* mesh, double x, double y, double z, double modulus, double margin) {
  if (!mesh) return DBH_ERR_INVALID_ARGUMENT;
  try {
    dbh::SetBoxPressureField(&mesh->m, x, y, z, modulus, margin);
  } catch (const std::exception& e) {
    g_shape_error = e.what();
    return DBH_ERR_INVALID_ARGUMENT;
  }
  return DBH_OK;
}
dbh_status dbh_mesh_make_cylinder_volume_with_ma(double radius, double length, double resolution_hint, dbh_mesh** out) {
  return Guard(out, [&] { return dbh::MakeCylinderVolumeMeshWithMa(radius, length, resolution_hint); });
}
dbh_status dbh_mesh_make_capsule_volume(double radius, double length, double resolution_hint, dbh_mesh** out) {
  return Guard(out, [&] { return dbh::MakeCapsuleVolumeMesh(radius, length, resolution_hint); });
}
dbh_status dbh_mesh_set_cylinder_pressure(dbh_mesh* mesh, double radius, double length, double modulus, double margin) {
  if (!mesh) return DBH_ERR_INVALID_ARGUMENT;
  try {
    dbh::SetCylinderPressureField(&mesh->m, radius, length, modulus, margin);
  } catch (const std::exception& e) {
    g_shape_error = e.what();
    return DBH_ERR_INVALID_ARGUMENT;
  }
  return DBH_OK;
}
dbh_status dbh_mesh_set_capsule_pressure(dbh_mesh* mesh, double radius, double length, double modulus, double margin) {
  if (!mesh) return DBH_ERR_INVALID_ARGUMENT;
  try {
    dbh::SetCapsulePressureField(&mesh->m, radius, length, modulus, margin);
  } catch (const std::exception& e) {
    g_shape_error = e.what();
    return DBH_ERR_INVALID_ARGUMENT;
  }
  return DBH_OK;
}
dbh_status dbh_mesh_from_arrays(int32_t num_vertices, const double* vertices, int32_t num_elements, const int32_t* elements,
                                int32_t vertices_per_element, dbh_mesh** out) {
  return Guard(out, [&] {
    if (num_vertices <= 0 || num_elements <= 0 || !vertices || !elements || (vertices_per_element != 3 && vertices_per_element != 4))
      throw std::invalid_argument("mesh_from_arrays: bad argument");
    dbh::HostMesh m;
    m.verts_per_elem = vertices_per_element;
    m.verts.assign(vertices, vertices + 3 * static_cast<size_t>(num_vertices));
    m.elems.assign(elements, elements + static_cast<size_t>(vertices_per_element) * num_elements);
    for (int32_t v : m.elems)
      if (v < 0 || v >= num_vertices) throw std::invalid_argument("mesh_from_arrays: vertex index out of range");
    return m;
  });
}
dbh_status dbh_mesh_make_convex_volume(const dbh_mesh* surface, dbh_mesh** out) {
  if (!surface) return DBH_ERR_INVALID_ARGUMENT;
  return Guard(out, [&] { return dbh::MakeConvexVolumeMesh(surface->m); });
}
dbh_status dbh_mesh_set_volume_mesh_pressure(dbh_mesh* mesh, double modulus, double margin) {
  if (!mesh) return DBH_ERR_INVALID_ARGUMENT;
  try {
    dbh::SetVolumeMeshPressureField(&mesh->m, modulus, margin);
  } catch (const std::exception& e) {
    g_shape_error = e.what();
    return DBH_ERR_INVALID_ARGUMENT;
  }
  return DBH_OK;
}
dbh_status dbh_mesh_get_sizes(const dbh_mesh* mesh, int32_t* num_vertices, int32_t* num_elements,
                              int32_t* vertices_per_element, int32_t* has_pressure) {
  if (!mesh) return DBH_ERR_INVALID_ARGUMENT;
  if (num_vertices) *num_vertices = mesh->m.num_vertices();
  if (num_elements) *num_elements = mesh->m.num_elements();
  if (vertices_per_element) *vertices_per_element = mesh->m.verts_per_elem;
  if (has_pressure) *has_pressure = !mesh->m.pressure.empty();
  return DBH_OK;
}
dbh_status dbh_mesh_copy(const dbh_mesh* mesh, double* vertices, int32_t* elements, double* pressure) {
  if (!mesh) return DBH_ERR_INVALID_ARGUMENT;
  if (vertices) std::copy(mesh->m.verts.begin(), mesh->m.verts.end(), vertices);
  if (elements) std::copy(mesh->m.elems.begin(), mesh->m.elems.end(), elements);
  if (pressure) {
    if (mesh->m.pressure.empty()) return DBH_ERR_INVALID_ARGUMENT;
    std::copy(mesh->m.pressure.begin(), mesh->m.pressure.end(), pressure);
  }
  return DBH_OK;
}
void dbh_mesh_destroy(dbh_mesh* mesh) { delete mesh; }

dbh_status dbh_register_mesh(dbh_context* ctx, const dbh_mesh* mesh, int32_t* asset_out) {
  if (!mesh) return DBH_ERR_INVALID_ARGUMENT;
  const dbh::HostMesh& m = mesh->m;
  if (m.verts_per_elem == 4) {
    if (m.pressure.empty()) return DBH_ERR_INVALID_ARGUMENT;
    return dbh_register_compliant_mesh(ctx, m.num_vertices(), m.verts.data(), m.num_elements(), m.elems.data(),
                                       m.pressure.data(), asset_out);
  }
  return dbh_register_rigid_mesh(ctx, m.num_vertices(), m.verts.data(), m.num_elements(), m.elems.data(), asset_out);
}

}  // extern "C"
