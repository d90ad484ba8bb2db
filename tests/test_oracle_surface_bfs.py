"""The oracle's restatement of the surface-BVH + BFS soft-soft intersector (the second VolumeIntersector::IntersectFields
overload, field_intersection.cc:254-375, with VolumeMeshTopology, volume_mesh_topology.cc:13-68), pinned by the
reference's own fixtures: volume_mesh_topology_test.cc:17-122 (two tetrahedra, octahedron) and
field_intersection_test.cc:305-344 (FullyPenetrated), :346-476 (the box pad x octahedron fixture: the two overloads emit
the same multiset of tetrahedron pairs, both representations), :532-560 (no intersection). CPU only."""
import numpy as np

from oracle import ho


def _multiset(surface):
    return sorted(zip(surface.elem_first.tolist(), surface.elem_second.tolist()))


def test_topology_of_two_tetrahedra():
    # volume_mesh_topology_test.cc:17-64.
    v = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [0, 0, -1]], float)
    m = ho.VolumeMesh.from_arrays(v, np.array([[0, 1, 2, 3], [0, 2, 1, 4]], np.int32))
    np.testing.assert_array_equal(m.neighbors(), [[-1, -1, -1, 1], [-1, -1, -1, 0]])


def test_topology_of_the_octahedron_fixture():
    # volume_mesh_topology_test.cc:66-122.
    v = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, -1], [0, 0, 1]], float)
    t = np.array([[0, 1, 3, 5], [1, 2, 3, 5], [0, 3, 1, 4], [1, 3, 2, 4]], np.int32)
    nb = ho.VolumeMesh.from_arrays(v, t).neighbors()
    want = -np.ones((4, 4), int)
    want[0, 0], want[1, 1] = 1, 0
    want[1, 3], want[3, 3] = 3, 1
    want[2, 0], want[3, 2] = 3, 2
    want[0, 3], want[2, 3] = 2, 0
    np.testing.assert_array_equal(nb, want)


def test_topology_is_symmetric_on_a_generated_mesh():
    m = ho.VolumeMesh.sphere(1.0, 0.3, dense=True)
    nb = m.neighbors()
    tets = m.tets
    for e in range(len(tets)):
        for f in range(4):
            o = nb[e, f]
            if o < 0:
                continue
            assert e in nb[o]
            face = {tets[e][(f + 1) % 4], tets[e][(f + 2) % 4], tets[e][(f + 3) % 4]}
            assert face <= set(tets[o].tolist())
    # Boundary faces = the triangles of the surface mesh (volume_to_surface_mesh.cc).
    assert int((nb < 0).sum()) == len(m.to_surface().tris)


def _pad_and_octahedron():
    # field_intersection_test.cc:346-391: Box(0.06, 0.10, 0.14) with medial axis, E = 1e5; Sphere(0.03) meshed with a
    # hint of 10 radii (an octahedron of 8 tetrahedra), E = 1e5.
    box = ho.VolumeMesh.box_ma(0.06, 0.10, 0.14)
    box_field = ho.VolumeField.box(box, 0.06, 0.10, 0.14, 1e5)
    octa = ho.VolumeMesh.sphere(0.03, 0.3)
    assert len(octa.tets) == 8
    octa_field = ho.VolumeField.sphere(octa, 0.03, 1e5)
    return box, box_field, octa, octa_field


def test_both_overloads_emit_the_same_tetrahedron_pairs():
    # field_intersection_test.cc:423-476 (X_MN = 0.03 x) and a few more poses of the same fixture.
    box, bf, octa, of = _pad_and_octahedron()
    rng = np.random.default_rng(5)
    poses = [ho.xform(None, [0.03, 0, 0])]
    for _ in range(6):
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        w, x, y, z = q
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                      [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                      [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        poses.append(ho.xform(R, rng.uniform(-1, 1, 3) * [0.04, 0.06, 0.08]))
    X_WM = ho.xform()
    produced = 0
    for X_MN in poses:
        pairs = ho.surface_bfs_pairs(box, bf, octa, of, X_MN)
        for rep in (ho.KTRIANGLE, ho.KPOLYGON):
            by_volume = ho.soft_soft(box, bf, octa, of, X_WM, X_MN, rep=rep, mode=ho.MODE_BRUTE)
            by_surface = ho.soft_soft(box, bf, octa, of, X_WM, X_MN, rep=rep, pairs=pairs) if len(pairs) else None
            assert (by_volume is None) == (by_surface is None)
            if by_volume is None:
                continue
            produced += 1
            assert by_volume.num_faces == by_surface.num_faces
            assert _multiset(by_volume) == _multiset(by_surface)
            np.testing.assert_allclose(by_volume.total_area, by_surface.total_area, rtol=1e-13)
    assert produced >= 2   # the reference's own pose produces a surface in both representations


def test_no_intersection():
    # field_intersection_test.cc:532-560: X_MN = 0.2 x.
    box, bf, octa, of = _pad_and_octahedron()
    assert len(ho.surface_bfs_pairs(box, bf, octa, of, ho.xform(None, [0.2, 0, 0]))) == 0


def test_fully_penetrated_small_sphere_is_missed_by_the_surface_overload():
    # field_intersection_test.cc:305-344: a sphere of radius 1/8 well inside a unit sphere, surfaces not touching: the
    # volume overload reports a surface, the surface overload none ("not a desirable behavior", as the reference says).
    big = ho.VolumeMesh.sphere(1.0, 1.0)
    big_field = ho.VolumeField.sphere(big, 1.0, 1e5)
    small = ho.VolumeMesh.sphere(0.125, 0.125)
    small_field = ho.VolumeField.sphere(small, 0.125, 1e5)
    X_MN = ho.xform(None, [0.5, 0, 0])
    assert ho.soft_soft(big, big_field, small, small_field, ho.xform(), X_MN, mode=ho.MODE_BRUTE) is not None
    assert len(ho.surface_bfs_pairs(big, big_field, small, small_field, X_MN)) == 0


def test_dense_spheres_same_surface_through_both_overloads():
    # Mesh scale: two dense spheres (4096 tetrahedra each), partially overlapping under a rotation.
    s = ho.VolumeMesh.sphere(1.0, 0.25, dense=True)
    f = ho.VolumeField.sphere(s, 1.0, 1e5)
    X_MN = ho.xform(ho.rpy_matrix(0.3, -0.2, 0.5), [1.5, 0.1, -0.2])
    pairs = ho.surface_bfs_pairs(s, f, s, f, X_MN)
    by_volume = ho.soft_soft(s, f, s, f, ho.xform(), X_MN, mode=ho.MODE_TREE)
    by_surface = ho.soft_soft(s, f, s, f, ho.xform(), X_MN, pairs=pairs)
    assert _multiset(by_volume) == _multiset(by_surface)
    # The BFS visits a small multiple of the emitting pairs, far fewer than the volume hierarchy's candidates.
    cand, _ = ho.candidates(s, s, X_MN, mode=ho.MODE_TREE)
    assert by_volume.num_faces <= len(pairs) < len(cand)


def test_the_surface_overload_loses_polygons_in_dense_clutter():
    """What the reference's equivalence test (field_intersection_test.cc:397-476, one pose of one fixture) does not show:
    the breadth-first walk stops at a tetrahedron pair that produces no polygon (a gradient filter rejects it, or its
    sliver dedupes below three vertices), so polygons that are only reachable through such a pair are lost. On a scene of
    BASELINE config 4 (axis-aligned grid boxes and ellipsoids in a bin) the faithful restatement emits a strict subset
    of the volume overload's tetrahedron pairs for some body pairs -- never a pair the volume overload does not emit --
    and about 0.3 % less area. The device intersector shares the property
    (tests/test_zz_gpu_surface_bfs.py::test_dense_clutter_is_a_subset_of_the_volume_result)."""
    from drake_b200 import workloads as W
    from tests import parity
    wl = W.config4(1)
    osc, order = parity.oracle_scene_for(wl)
    poses = wl.poses[0]
    want, _ = osc.compute(poses[order], ho.KPOLYGON)
    ids = list(wl.geom_id)
    meshidx = wl.per_scene_mesh[0] if wl.per_scene_mesh is not None else wl.geom_mesh
    cache = {}

    def oracle_mesh(mi):
        if mi not in cache:
            v, e, p = wl.meshes[mi].arrays()
            m = ho.VolumeMesh.from_arrays(v, e)
            cache[mi] = (m, ho.VolumeField.from_values(m, p))
        return cache[mi]

    soft_soft = incomplete = 0
    area_volume = area_surface = 0.0
    for s in want:
        if s.grad_eM is None or s.grad_eN is None:
            continue
        gM, gN = ids.index(s.id_M), ids.index(s.id_N)
        (mM, fM), (mN, fN) = oracle_mesh(meshidx[gM]), oracle_mesh(meshidx[gN])
        pairs = ho.surface_bfs_pairs(mM, fM, mN, fN, ho.invert_and_compose(poses[gM], poses[gN]))
        b = ho.soft_soft(mM, fM, mN, fN, poses[gM], poses[gN], s.id_M, s.id_N, ho.KPOLYGON, pairs=pairs) if len(pairs) else None
        soft_soft += 1
        area_volume += s.total_area
        if b is None:
            incomplete += 1
            continue
        area_surface += b.total_area
        full, part = set(_multiset(s)), _multiset(b)
        assert len(set(part)) == len(part) and set(part) <= full       # nothing invented
        incomplete += len(part) < len(full)
    assert soft_soft >= 20 and incomplete >= 1
    assert 0.99 * area_volume < area_surface < area_volume
