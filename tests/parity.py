"""Shared helpers for the GPU-vs-oracle parity tests (test infrastructure)."""
from __future__ import annotations

import numpy as np

from oracle import ho

REL_TOL = 1e-10  # north_star: centroids, areas, normals, pressures within 1e-10 relative (fp64)

_MEMO = {}


def memo(key, fn):
    """Oracle results shared by the runs of one test through both mid-phase implementations."""
    if key not in _MEMO:
        _MEMO[key] = fn()
    return _MEMO[key]


def register_compliant(ctx, mesh: "ho.VolumeMesh", fld: "ho.VolumeField") -> int:
    return ctx.register_compliant_mesh(mesh.verts, mesh.tets, fld.values)


def register_rigid(ctx, mesh: "ho.TriMesh") -> int:
    return ctx.register_rigid_mesh(mesh.verts, mesh.tris)


def face_keys(elem_a, elem_b):
    return np.asarray(elem_a, np.int64) * (1 << 32) + np.asarray(elem_b, np.int64)


def assert_close(name, got, want, scale=None, tol=REL_TOL):
    got, want = np.asarray(got, float), np.asarray(want, float)
    assert got.shape == want.shape, f"{name}: shape {got.shape} vs {want.shape}"
    if got.size == 0:
        return
    s = np.max(np.abs(want)) if scale is None else scale
    s = max(s, 1e-300)
    err = np.max(np.abs(got - want)) / s
    assert err <= tol, f"{name}: relative error {err:.3e} > {tol:.1e}"


def polygon_vertex_lists(surface):
    """Per-face vertex arrays in winding order."""
    return [surface.vertices[idx] for idx in surface.polygons()]


def compare_surface(gpu, ora, check_bitwise_keys=True, welded=False):
    """gpu: drake_b200.ContactSurface; ora: oracle ho.Surface. Faces are matched by their emitting
    element pair (the reference's own notion of surface equality, field_intersection_test.cc:397-421);
    within a polygon's triangle fan the order is the builder's. welded: a half-space surface, whose
    vertices the oracle welds like the reference (cut_edges, mesh_plane_intersection.cc:205-231) while
    the kernels keep them per polygon: vertex counts differ, every face's vertices must still agree."""
    assert (gpu.id_M, gpu.id_N) == (ora.id_M, ora.id_N)
    assert gpu.representation == ora.rep
    assert gpu.num_faces == ora.num_faces, f"face count {gpu.num_faces} vs {ora.num_faces}"
    if welded:
        assert len(gpu.vertices) >= len(ora.vertices)
    else:
        assert len(gpu.vertices) == len(ora.vertices)
    gk, ok = face_keys(gpu.elem_M, gpu.elem_N), face_keys(ora.elem_first, ora.elem_second)
    gorder, oorder = np.argsort(gk, kind="stable"), np.argsort(ok, kind="stable")
    np.testing.assert_array_equal(gk[gorder], ok[oorder])  # bit-exact emitted element-pair multiset
    L = max(np.max(np.abs(ora.vertices)), 1e-300)
    assert_close("areas", gpu.areas[gorder], ora.areas[oorder], scale=max(ora.areas.max(), 1e-300))
    # Degenerate polygons of a half-space surface (a vertex or an edge exactly on the boundary: area ~ 0) have no
    # meaningful normal or centroid: quotients of rounding noise (polygon_surface_mesh.cc:189-207), which differs
    # between welded and per-polygon vertices. Mesh-mesh surfaces are compared in full.
    solid = (ora.areas[oorder] > 1e-12 * max(ora.areas.max(), 1e-300)) if welded else np.ones(len(oorder), bool)
    assert_close("normals", gpu.normals[gorder][solid], ora.normals[oorder][solid], scale=1.0)
    assert_close("centroids", gpu.centroids[gorder][solid], ora.centroids[oorder][solid], scale=L)
    pscale = max(np.max(np.abs(ora.pressures)), 1e-300)
    assert_close("pressure_at_centroid", gpu.pressure_at_centroid[gorder][solid], ora.e_centroid[oorder][solid], scale=pscale)
    assert (gpu.grad_eM is None) == (ora.grad_eM is None)
    assert (gpu.grad_eN is None) == (ora.grad_eN is None)
    if ora.grad_eM is not None:
        assert_close("grad_eM", gpu.grad_eM[gorder], ora.grad_eM[oorder])
    if ora.grad_eN is not None:
        assert_close("grad_eN", gpu.grad_eN[gorder], ora.grad_eN[oorder])
    # Vertices and per-vertex pressures, polygon by polygon in winding order.
    gp, op = gpu.polygons(), ora.polygons()
    for a, b, is_solid in zip(gorder, oorder, solid):
        ia, ib = np.asarray(gp[a]), np.asarray(op[b])
        assert len(ia) == len(ib)
        if not is_solid and ora.rep == ho.KTRIANGLE:
            ia, ib = ia[:2], ib[:2]   # the fan's third vertex is the degenerate polygon's (meaningless) centroid
        assert_close("vertices", gpu.vertices[ia], ora.vertices[ib], scale=L)
        assert_close("pressures", gpu.pressures[ia], ora.pressures[ib], scale=pscale)
    assert_close("total_area", gpu.total_area, ora.total_area)
    assert_close("surface centroid", gpu.centroid, ora.centroid, scale=L)


def random_pose(rng, translation_scale=1.0, center=(0, 0, 0)):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    return ho.xform(R, np.asarray(center, float) + translation_scale * rng.uniform(-1, 1, size=3))


def oracle_scene_for(workload):
    """Oracle scene built from the SAME arrays the GPU registers for a drake_b200.workloads.Workload
    (one template scene; poses differ per scene). Returns (scene, order) where `order` sorts the
    template geometries by ascending id, the order in which the oracle wants poses."""
    sc = ho.Scene()
    cache = {}
    for g, mi in enumerate(workload.geom_mesh):
        m = workload.meshes[mi]
        if m.vertices_per_element == 0:  # drake_b200.workloads.HalfSpaceShape
            sc.add_half_space(workload.geom_id[g], m.compliant, m.pressure_scale, m.margin, bool(workload.geom_anchored[g]))
            continue
        if mi not in cache:
            v, e, p = m.arrays()
            if m.vertices_per_element == 4:
                om = ho.VolumeMesh.from_arrays(v, e)
                cache[mi] = (om, ho.VolumeField.from_values(om, p))
            else:
                cache[mi] = (ho.TriMesh.from_arrays(v, e), None)
        om, of = cache[mi]
        if of is None:
            sc.add_rigid(workload.geom_id[g], om, bool(workload.geom_anchored[g]))
        else:
            sc.add_compliant(workload.geom_id[g], om, of, bool(workload.geom_anchored[g]))
    if workload.excluded_local is not None:
        for a, b in np.asarray(workload.excluded_local).reshape(-1, 2):
            sc.exclude(int(workload.geom_id[a]), int(workload.geom_id[b]))
    return sc, np.argsort(np.asarray(workload.geom_id))
