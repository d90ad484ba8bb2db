"""BASELINE.json configs as concrete synthetic inputs.

Meshes come from the library's own generators (drake_b200.Mesh, the product
side of MakeSphereVolumeMesh / MakeEllipsoidVolumeMesh / MakeBoxVolumeMesh*) --
or, when a config is given another mesh factory with the same constructors
(`Mesh=`; the reference arm of bench.py passes the CPU checker's, so that it never
maps the product library), from that one --
poses from a counter-based RNG (SceneDraws: draw k of scene i is a pure
function of (config, i, k)) so any rank can generate any scene range.

  C1  mesh_intersection_benchmark: soft ellipsoid x rigid sphere, 13 argument triples
  C2  compliant_mesh_intersection_benchmark: soft ellipsoid x soft sphere (+ two dense spheres)
  C3  ball_plate batched: soft ball + soft floor box + rigid plate stand-in, S scenes
  C4  dense clutter: 32 soft boxes/ellipsoids in a rigid bin, S scenes
  C5  high-resolution stress: two dense L=6 spheres (2 097 152 tets each), S scenes
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from .engine import DENSE_INTERIOR_VERTICES, POLYGON, SINGLE_INTERIOR_VERTEX, TRIANGLE, Mesh

SEED = 0xD2A4E


def xform(R: Optional[np.ndarray] = None, p: Sequence[float] = (0, 0, 0)) -> np.ndarray:
    """RigidTransform as 12 doubles: R column-major then p."""
    R = np.eye(3) if R is None else np.asarray(R, dtype=np.float64)
    return np.concatenate([R.T.reshape(-1), np.asarray(p, dtype=np.float64)])


def xforms(R: np.ndarray, p: np.ndarray) -> np.ndarray:
    """Batched: R [N,3,3], p [N,3] -> [N,12]."""
    return np.concatenate([np.transpose(R, (0, 2, 1)).reshape(len(R), 9), p], axis=1)


def quat_to_matrix(q: np.ndarray) -> np.ndarray:
    """Unit quaternions [N,4] (w, x, y, z) -> rotation matrices [N,3,3]."""
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = np.empty((len(q), 3, 3))
    R[:, 0, 0] = 1 - 2 * (y * y + z * z)
    R[:, 0, 1] = 2 * (x * y - z * w)
    R[:, 0, 2] = 2 * (x * z + y * w)
    R[:, 1, 0] = 2 * (x * y + z * w)
    R[:, 1, 1] = 1 - 2 * (x * x + z * z)
    R[:, 1, 2] = 2 * (y * z - x * w)
    R[:, 2, 0] = 2 * (x * z - y * w)
    R[:, 2, 1] = 2 * (y * z + x * w)
    R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def angle_axis_matrix(angle: float, axis: Sequence[float]) -> np.ndarray:
    a = np.asarray(axis, dtype=np.float64)
    a = a / np.linalg.norm(a)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return np.eye(3) + math.sin(angle) * K + (1 - math.cos(angle)) * (K @ K)


class SceneDraws:
    """Counter-based random draws: draw k of scene i of config c is a pure function of (c, i, k)
    (SplitMix64 of the packed counter), so any rank generates any scene range and gets exactly the
    poses a single process would (tests/test_sharding_cpu.py)."""

    def __init__(self, config_id: int, scene0: int, num_scenes: int):
        self.scenes = np.arange(scene0, scene0 + num_scenes, dtype=np.uint64)
        self.key = np.uint64((SEED ^ (config_id * 0x9E3779B1)) & 0xFFFFFFFF)
        self.next = 0

    def _bits(self, n: int) -> np.ndarray:
        """[S, n] uint64, draws next .. next + n of every scene."""
        k = np.arange(self.next, self.next + n, dtype=np.uint64)
        self.next += n
        with np.errstate(over="ignore"):
            z = (self.scenes[:, None] << np.uint64(20)) + k[None, :] + (self.key << np.uint64(44))
            z = z + np.uint64(0x9E3779B97F4A7C15)
            z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            z = z ^ (z >> np.uint64(31))
        return z

    def uniform(self, n: int) -> np.ndarray:
        """[S, n] doubles in [0, 1)."""
        return (self._bits(n) >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)

    def normal(self, n: int) -> np.ndarray:
        """[S, n] standard normals (Box-Muller)."""
        m = (n + 1) // 2
        u1 = 1.0 - self.uniform(m)
        u2 = self.uniform(m)
        r = np.sqrt(-2.0 * np.log(u1))
        out = np.concatenate([r * np.cos(2 * np.pi * u2), r * np.sin(2 * np.pi * u2)], axis=1)
        return out[:, :n]

    def rotations(self, n: int = 1) -> np.ndarray:
        """[S, n, 3, 3] uniform random rotations (normalised 4-vectors of normals)."""
        q = self.normal(4 * n).reshape(len(self.scenes) * n, 4)
        q /= np.linalg.norm(q, axis=1, keepdims=True)
        return quat_to_matrix(q).reshape(len(self.scenes), n, 3, 3)

    def directions(self) -> np.ndarray:
        d = self.normal(3)
        return d / np.linalg.norm(d, axis=1, keepdims=True)


@dataclass
class HalfSpaceShape:
    """HalfSpace (z <= 0 of its frame) in a workload's asset list; compliant: pressure_scale = E / slab thickness."""
    compliant: bool = False
    pressure_scale: float = 0.0
    margin: float = 0.0
    vertices_per_element: int = 0


@dataclass
class Workload:
    """A batch of scenes instancing one scene template."""
    name: str
    description: str
    meshes: List[Mesh]                 # distinct mesh assets
    geom_mesh: List[int]               # per template geometry: index into meshes
    geom_id: List[int]                 # GeometryId per template geometry
    geom_anchored: List[int]
    poses: np.ndarray                  # [S, G, 12]
    representation: int = POLYGON
    excluded_local: Optional[np.ndarray] = None
    per_scene_mesh: Optional[np.ndarray] = None  # [S, G] mesh index when scenes differ in assets
    normal_filter: bool = True         # False: the reference's mesh_intersection_benchmark runs its intersector with
                                       # filter_face_normal_along_field_gradient = false

    @property
    def num_scenes(self) -> int:
        return len(self.poses)

    @property
    def geoms_per_scene(self) -> int:
        return len(self.geom_id)

    def register(self, ctx, assets=None):
        """Registers the assets (unless the handles of an earlier registration in the same context are
        passed) and returns a SceneBatch for all scenes. The handles used are kept in self.assets."""
        from .engine import SceneBatch
        if assets is None:
            assets = [ctx.register_half_space(m.compliant, m.pressure_scale, m.margin) if isinstance(m, HalfSpaceShape)
                      else ctx.register_mesh(m) for m in self.meshes]
        self.assets = list(assets)
        S, G = self.num_scenes, self.geoms_per_scene
        if self.per_scene_mesh is None:
            batch = SceneBatch.replicated(ctx, S, [assets[i] for i in self.geom_mesh], self.geom_id,
                                          self.geom_anchored, self.excluded_local)
            if not self.normal_filter:
                batch.set_normal_filter(False)
            return batch
        so = np.arange(S + 1, dtype=np.int32) * G
        ga = np.asarray(assets, np.int32)[self.per_scene_mesh.reshape(-1)]
        gi = np.tile(np.asarray(self.geom_id, np.int64), S)
        an = np.tile(np.asarray(self.geom_anchored, np.uint8), S)
        batch = SceneBatch(ctx, so, ga, gi, an, None)
        if not self.normal_filter:
            batch.set_normal_filter(False)
        return batch


# ------------------------------------------------------------------ C1 / C2 ---

C1_ARGS = [(0, 4, 0), (1, 4, 0), (2, 4, 0), (3, 4, 0), (2, 0, 0), (2, 1, 0), (2, 2, 0), (2, 3, 0), (2, 4, 1), (2, 4, 2),
           (2, 4, 3), (2, 3, 1), (2, 2, 2)]  # mesh_intersection_benchmark.cc:259-271
_HINTS = [4.0, 3.0, 2.0, 1.0]
_TRANS = [(7, 7, 7), (4, 4, 4), (3.5, 3.5, 3.5), (1.2, 1.2, 1.2), (0, 0, 0)]


def config1(args=C1_ARGS, Mesh=Mesh) -> Workload:
    """geometry/benchmarking/mesh_intersection_benchmark.cc:153-271, one scene per argument triple."""
    meshes, index = [], {}
    for res in sorted({a[0] for a in args}):
        s = Mesh.ellipsoid_volume(3.01, 3.5, 4.0, _HINTS[res], DENSE_INTERIOR_VERTICES)
        s.set_ellipsoid_pressure(3.01, 3.5, 4.0, 1e5)
        r = Mesh.sphere_surface(3.0, _HINTS[res])
        index[res] = (len(meshes), len(meshes) + 1)
        meshes += [s, r]
    per_scene = np.array([index[a[0]] for a in args], np.int32)
    poses = np.stack([np.stack([xform(), xform(angle_axis_matrix(a[2] / 3.0 * math.pi / 4, [1, 1, 1]), _TRANS[a[1]])])
                      for a in args])
    return Workload("c1_mesh_intersection_benchmark", "soft ellipsoid (3.01,3.5,4) dense vs rigid sphere r=3",
                    meshes, [0, 1], [1, 2], [0, 0], poses, TRIANGLE, None, per_scene)


def config2(resolutions=(0, 1, 2, 3), overlaps=(0, 1, 2, 3), Mesh=Mesh) -> Workload:
    """geometry/benchmarking/compliant_mesh_intersection_benchmark.cc:112-202."""
    meshes, index = [], {}
    for res in resolutions:
        s = Mesh.ellipsoid_volume(3.01, 3.5, 4.0, _HINTS[res], DENSE_INTERIOR_VERTICES)
        s.set_ellipsoid_pressure(3.01, 3.5, 4.0, 1e5)
        r = Mesh.sphere_volume(3.0, _HINTS[res], DENSE_INTERIOR_VERTICES).set_sphere_pressure(3.0, 1e5)
        index[res] = (len(meshes), len(meshes) + 1)
        meshes += [s, r]
    scenes = [(r, o) for r in resolutions for o in overlaps]
    per_scene = np.array([index[r] for r, _ in scenes], np.int32)
    poses = np.stack([np.stack([xform(), xform(None, _TRANS[o])]) for _, o in scenes])
    return Workload("c2_compliant_mesh_intersection_benchmark", "soft ellipsoid vs soft sphere r=3, dense", meshes,
                    [0, 1], [1, 2], [0, 0], poses, POLYGON, None, per_scene)


def config2_spheres(num_scenes: int, scene0: int = 0, hint: float = 0.5, Mesh=Mesh) -> Workload:
    """Two soft spheres r=3 (dense, fine resolution_hint) at random relative poses."""
    m = Mesh.sphere_volume(3.0, hint, DENSE_INTERIOR_VERTICES).set_sphere_pressure(3.0, 1e5)
    rng = SceneDraws(2, scene0, num_scenes)
    R = rng.rotations(2)
    d = rng.directions()
    dist = 3.6 + (5.85 - 3.6) * rng.uniform(1)
    poses = np.stack([xforms(R[:, 0], np.zeros((num_scenes, 3))), xforms(R[:, 1], dist * d)], axis=1)
    return Workload(f"c2_two_dense_spheres_hint{hint}", f"two soft spheres r=3 dense ({m.num_elements} tets each)",
                    [m], [0, 0], [1, 2], [0, 0], poses, POLYGON)


# ----------------------------------------------------------------------- C3 ---

def config3(num_scenes: int = 16384, scene0: int = 0, Mesh=Mesh) -> Workload:
    """examples/hydroelastic/ball_plate geometry, batched. Per scene: soft floor box 0.3x0.3x0.05
    (floor.sdf, E = 3e4, medial-axis mesh, 44 tets, anchored), rigid plate stand-in = box
    0.2x0.2x0.01 surface (24 triangles; plate_8in.sdf is external to the reference tree), soft ball
    r = 0.05, hint 0.015, E = 3e4 (512 tets, ball_plate_run_dynamics.cc:27-30,85). The plate rests
    0.5 mm into the floor; the ball is dropped at x~U(-0.05,0.15), y~U(-0.1,0.1) with penetration
    into the plate ~U(0, r/2) and a uniform random orientation."""
    floor = Mesh.box_volume_with_ma(0.3, 0.3, 0.05).set_box_pressure(0.3, 0.3, 0.05, 3e4)
    plate = Mesh.box_surface(0.2, 0.2, 0.01)
    r = 0.05
    ball = Mesh.sphere_volume(r, 0.015, SINGLE_INTERIOR_VERTEX).set_sphere_pressure(r, 3e4)
    rng = SceneDraws(3, scene0, num_scenes)
    S = num_scenes
    draws = rng.uniform(3)
    R = rng.rotations(1)[:, 0]
    plate_top = 0.005 - 0.0005 + 0.005
    x = -0.05 + 0.2 * draws[:, 0]
    y = -0.1 + 0.2 * draws[:, 1]
    pen = 0.5 * r * draws[:, 2]
    pb = np.stack([x, y, plate_top + r - pen], axis=1)
    I = np.tile(np.eye(3), (S, 1, 1))
    poses = np.stack([xforms(I, np.tile([0, 0, -0.025], (S, 1))), xforms(I, np.tile([0, 0, 0.0045], (S, 1))),
                      xforms(R, pb)], axis=1)
    return Workload("c3_ball_plate_batched", "soft ball (512 tets) + soft floor (44 tets) + rigid plate (24 tris)",
                    [floor, plate, ball], [0, 1, 2], [1, 2, 3], [1, 0, 0], poses, POLYGON)


def config3_half_space(num_scenes: int = 16384, scene0: int = 0, Mesh=Mesh) -> Workload:
    """The ball_plate scene on a rigid HalfSpace ground (the most common rigid geometry):
    rigid half space z <= 0 (anchored), the same soft ball (512 tets, E = 3e4) resting on it with penetration
    ~U(0, r/2) under a uniform random orientation, and the rigid plate (24 triangles) pressed ~U(0, r/4) into
    the top of the ball. Per scene: one compliant-mesh x rigid-half-space pair (k_plane) and one soft-rigid
    mesh pair; plate x ground is rigid-rigid and filtered, as MultibodyPlant's default filters would."""
    r = 0.05
    ball = Mesh.sphere_volume(r, 0.015, SINGLE_INTERIOR_VERTEX).set_sphere_pressure(r, 3e4)
    plate = Mesh.box_surface(0.2, 0.2, 0.01)
    rng = SceneDraws(6, scene0, num_scenes)
    S = num_scenes
    draws = rng.uniform(4)
    R = rng.rotations(1)[:, 0]
    x = -0.05 + 0.1 * draws[:, 0]
    y = -0.05 + 0.1 * draws[:, 1]
    zb = r - 0.5 * r * draws[:, 2]
    zp = zb + r - 0.25 * r * draws[:, 3] + 0.005
    I = np.tile(np.eye(3), (S, 1, 1))
    poses = np.stack([xforms(I, np.zeros((S, 3))), xforms(I, np.stack([np.zeros(S), np.zeros(S), zp], axis=1)),
                      xforms(R, np.stack([x, y, zb], axis=1))], axis=1)
    return Workload("c3h_ball_plate_half_space_ground",
                    "soft ball (512 tets) on a rigid half space, rigid plate (24 tris) on the ball",
                    [HalfSpaceShape(), plate, ball], [0, 1, 2], [1, 2, 3], [1, 0, 0], poses, POLYGON,
                    excluded_local=np.array([[0, 1]], np.int32))


# ----------------------------------------------------------------------- C4 ---

def config4(num_scenes: int = 4096, scene0: int = 0, bodies: int = 32, Mesh=Mesh) -> Workload:
    """Dense clutter: `bodies` soft bodies (half boxes via MakeBoxVolumeMesh hint 0.02, sizes
    ~U(0.04,0.10)^3; half dense ellipsoids, semi-axes ~U(0.02,0.05), hint 0.01) in a rigid bin of
    5 anchored box walls (interior 0.4x0.4x0.2). The body palette is shared by all scenes; poses
    differ per scene (centres uniform in the bin interior, uniform orientations)."""
    prng = np.random.Generator(np.random.Philox(key=SEED ^ 0x40))
    meshes: List[Mesh] = []
    nb = bodies // 2
    for _ in range(nb):
        sx, sy, sz = prng.uniform(0.04, 0.10, size=3)
        meshes.append(Mesh.box_volume(sx, sy, sz, 0.02).set_box_pressure(sx, sy, sz, 1e5))
    for _ in range(bodies - nb):
        a, b, c = prng.uniform(0.02, 0.05, size=3)
        meshes.append(Mesh.ellipsoid_volume(a, b, c, 0.01, DENSE_INTERIOR_VERTICES).set_ellipsoid_pressure(a, b, c, 1e5))
    w = 0.01  # wall thickness
    walls = [((0.42, 0.42, w), (0, 0, -0.1 - w / 2)), ((w, 0.42, 0.2), (-0.2 - w / 2, 0, 0)),
             ((w, 0.42, 0.2), (0.2 + w / 2, 0, 0)), ((0.42, w, 0.2), (0, -0.2 - w / 2, 0)),
             ((0.42, w, 0.2), (0, 0.2 + w / 2, 0))]
    wall_mesh = []
    for size, _ in walls:
        wall_mesh.append(len(meshes))
        meshes.append(Mesh.box_surface(*size))
    S = num_scenes
    rng = SceneDraws(4, scene0, num_scenes)
    c = rng.uniform(3 * bodies).reshape(S, bodies, 3) * np.array([0.36, 0.36, 0.16]) - np.array([0.18, 0.18, 0.08])
    R = rng.rotations(bodies).reshape(S * bodies, 3, 3)
    body_poses = xforms(R, c.reshape(-1, 3)).reshape(S, bodies, 12)
    wall_poses = np.tile(np.stack([xform(None, p) for _, p in walls])[None], (S, 1, 1))
    poses = np.concatenate([wall_poses, body_poses], axis=1)
    geom_mesh = wall_mesh + list(range(bodies))
    ids = list(range(1, 6)) + list(range(100, 100 + bodies))
    anchored = [1] * 5 + [0] * bodies
    return Workload("c4_dense_clutter", f"{bodies} soft boxes/ellipsoids in a rigid 5-wall bin", meshes, geom_mesh,
                    ids, anchored, poses, POLYGON)


# ----------------------------------------------------------------------- C5 ---

def config5(num_scenes: int = 1024, scene0: int = 0, hint: float = 0.03, Mesh=Mesh) -> Workload:
    """High-resolution stress: two soft spheres r = 1, dense, hint 0.03 -> level 6, 2 097 152 tets
    each ("~1M tets" is not reachable: level 5 gives 262 144); one shared asset instanced twice per
    scene; centre distance ~U(1.2, 1.95), uniform relative orientation."""
    m = Mesh.sphere_volume(1.0, hint, DENSE_INTERIOR_VERTICES).set_sphere_pressure(1.0, 1e5)
    rng = SceneDraws(5, scene0, num_scenes)
    S = num_scenes
    R = rng.rotations(2)
    d = rng.directions()
    dist = 1.2 + (1.95 - 1.2) * rng.uniform(1)
    poses = np.stack([xforms(R[:, 0], np.zeros((S, 3))), xforms(R[:, 1], dist * d)], axis=1)
    return Workload("c5_high_resolution_stress", f"two soft spheres r=1 dense ({m.num_elements} tets each)", [m],
                    [0, 0], [1, 2], [0, 0], poses, POLYGON)


CONFIGS = {"c1": config1, "c2": config2, "c2s": config2_spheres, "c3": config3, "c3h": config3_half_space, "c4": config4,
           "c5": config5}
