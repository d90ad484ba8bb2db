"""The point-pair half of ComputeContactSurfacesWithFallback: the oracle's numpy restatement of
Drake's closed forms for pairs with a sphere (oracle/ho_point_pair.py) is pinned by the reference's own test scenes
(penetration_as_point_pair_callback_test.cc), and the adapter's C++ (include/drake_b200_point_pair.h), compiled for
the host, must agree with it to the last bit on those scenes and on random ones."""
import math
import os
import subprocess

import numpy as np
import pytest

from oracle import ho
from oracle import ho_point_pair as pp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
K_RADIUS, K_EPS = 0.5, np.finfo(float).eps   # test fixture constants (:342-365, kEps = epsilon)
BOX = (0.1, 0.2, 0.3)
CYLINDER, CAPSULE = (1.2, 2.1), (1.3, 2.3)   # radius, length (:342-347 of the fixture)


def rot(axis, angle):
    c, s = math.cos(angle), math.sin(angle)
    return {"x": np.array([[1, 0, 0], [0, c, -s], [0, s, c]]), "y": np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]]),
            "z": np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])}[axis]


def _check_pair(res, target_depth, tol=16 * K_EPS):
    assert res is not None
    assert abs(res["depth"] - target_depth) <= tol
    assert abs(float((res["p_WCb"] - res["p_WCa"]) @ res["nhat_BA_W"]) - target_depth) <= tol
    assert abs(np.linalg.norm(res["nhat_BA_W"]) - 1) <= tol


def test_sphere_sphere_no_collision_and_target_depth():
    """TestNoCollision (:98-117) and TestCollisionFilterRespected's geometry (:119-176): depth 0.1 at an arbitrary
    relative pose, identical whichever sphere comes first once ids are ordered."""
    I = ho.xform(None, [0, 0, 0])
    assert pp.point_pair(("sphere", K_RADIUS), I, ("sphere", K_RADIUS), ho.xform(None, [K_RADIUS * 3, 0, 0])) is None
    target = 0.1
    d = np.array([1.0, -2.0, 3.0])
    p_WBo = d / np.linalg.norm(d) * (K_RADIUS * 2 - target)
    X_WB = ho.xform(rot("y", math.pi / 3) @ rot("z", -math.pi / 7), p_WBo)
    first = pp.point_pair(("sphere", K_RADIUS), I, ("sphere", K_RADIUS), X_WB)
    _check_pair(first, target)
    assert abs(np.linalg.norm(first["p_WCa"] - first["p_WCb"]) - first["depth"]) <= 4 * K_EPS
    np.testing.assert_allclose(first["p_WCb"] - first["p_WCa"], first["depth"] * first["nhat_BA_W"], rtol=0, atol=4 * K_EPS)


def test_sphere_box_cylinderless_and_half_space_target_depths():
    """SphereBoxDouble (:258-271,430-435) and SphereHalfspaceDouble (:292-310,458-461), both operand orders."""
    I = ho.xform(None, [0, 0, 0])
    target = 0.1
    R_WB = rot("z", 0.2)
    X_WB = ho.xform(R_WB, R_WB @ np.array([1.0, 0, 0]) * (K_RADIUS + BOX[0] / 2 - target))
    a = pp.point_pair(("sphere", K_RADIUS), I, ("box", BOX), X_WB)
    _check_pair(a, target)
    b = pp.point_pair(("box", BOX), X_WB, ("sphere", K_RADIUS), I)   # reversed order: A and B swapped
    assert b["depth"] == a["depth"]
    np.testing.assert_array_equal(b["p_WCa"], a["p_WCb"])
    np.testing.assert_array_equal(b["p_WCb"], a["p_WCa"])
    np.testing.assert_array_equal(b["nhat_BA_W"], -a["nhat_BA_W"])
    R_WH = rot("z", 0.1) @ rot("x", 0.5)
    p_WHo = np.array([0.5, -0.2, 0.9])
    target_h = 0.0 + float(np.array([0, 0, 1.0]) @ (R_WH.T @ p_WHo)) + K_RADIUS   # :302-305
    h = pp.point_pair(("sphere", K_RADIUS), I, ("half_space",), ho.xform(R_WH, p_WHo))
    _check_pair(h, target_h)


def test_box_distance_regions():
    """ComputeDistanceToBox's three regions (:406-482): outside (vertex region), on the boundary (edge: normalised sum of
    the face normals), inside (nearest face; ties go to +x first)."""
    I = ho.xform(None, [0, 0, 0])
    p_GN, dist, grad = pp.distance_to_box((2, 2, 2), I, [2, 2, 2])
    np.testing.assert_array_equal(p_GN, [1, 1, 1])
    assert dist == pytest.approx(math.sqrt(3)) and np.allclose(grad, np.ones(3) / math.sqrt(3))
    p_GN, dist, grad = pp.distance_to_box((2, 2, 2), I, [1, 1, 0.25])
    assert dist == pytest.approx(0, abs=1e-15) and np.allclose(grad, [1 / math.sqrt(2), 1 / math.sqrt(2), 0])
    p_GN, dist, grad = pp.distance_to_box((2, 2, 2), I, [0, 0, 0])
    np.testing.assert_array_equal(grad, [1, 0, 0])
    assert dist == -1.0
    p_GN, dist, grad = pp.distance_to_box((2, 4, 6), I, [0.5, -1.9, 0.2])
    np.testing.assert_array_equal(grad, [0, -1, 0])
    assert dist == pytest.approx(-0.1)
    # Sphere centre exactly at the other sphere's centre: the documented arbitrary gradient +x (:24-37).
    _, dist, grad = pp.distance_to_sphere(0.5, I, [0, 0, 0])
    np.testing.assert_array_equal(grad, [1, 0, 0])
    assert dist == -0.5


def _angle_axis(angle, axis):
    axis = np.asarray(axis, float) / np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + math.sin(angle) * K + (1 - math.cos(angle)) * (K @ K)


def _distance_case(fn, X_WG, p_GN, p_NQ, inside=False, gradient_defined=True, tol=64 * K_EPS):
    """PointShapeAutoDiffSignedDistanceTester::Test (distance_to_point_callback_test.cc:100-200), double part: the query
    point Q = N + p_NQ must give back the witness N, the signed distance -/+ |p_NQ| and the gradient along +/- p_NQ."""
    R, t = np.asarray(X_WG[:9]).reshape(3, 3).T, np.asarray(X_WG[9:])
    p_GN, p_NQ = np.asarray(p_GN, float), np.asarray(p_NQ, float)
    got_N, got_d, got_g = fn(X_WG, R @ (p_GN + p_NQ) + t)
    d = float(np.linalg.norm(p_NQ))
    np.testing.assert_allclose(got_N, p_GN, rtol=0, atol=tol)
    assert abs(got_d - (-d if inside else d)) <= tol
    if gradient_defined and d > 0:
        np.testing.assert_allclose(got_g, R @ (p_NQ / d) * (-1 if inside else 1), rtol=0, atol=tol)
    assert abs(np.linalg.norm(got_g) - 1) <= tol


def test_capsule_distance_reference_cases():
    """GTEST_TEST(DistanceToPoint, Capsule) (distance_to_point_callback_test.cc:276-333): Capsule(0.7, 1.3) at an arbitrary
    pose; nearest point on the top cap, beside the spine, on the bottom cap; query outside / on / inside / on the spine."""
    r, L = 0.7, 1.3
    X_WG = ho.xform(_angle_axis(math.pi / 5, [1, 2, 3]), [0.5, 1.25, -2])
    fn = lambda X, q: pp.distance_to_capsule(r, L, X, q)
    dirs = [np.array(v, float) / np.linalg.norm(v) for v in ([2, -3, 6], [2, -3, 0], [2, -3, -6])]
    offs = [np.array([0, 0, L / 2]), np.array([0, 0, L / 4]), np.array([0, 0, -L / 2])]
    for v, o in zip(dirs, offs):
        N = r * v + o
        _distance_case(fn, X_WG, N, 1.5 * v)
        _distance_case(fn, X_WG, N, np.zeros(3))
        _distance_case(fn, X_WG, N, -(0.5 * r) * v, inside=True)
    # On the spine the gradient is the documented arbitrary +x of the sphere at its own centre; distance = -radius.
    _, d, g = pp.distance_to_capsule(r, L, ho.xform(None, [0, 0, 0]), [0, 0, 0.2])
    assert d == -r and g.tolist() == [1, 0, 0]


def test_cylinder_distance_reference_cases():
    """GTEST_TEST(DistanceToPoint, Cylinder) (distance_to_point_callback_test.cc:579-663): Cylinder(0.75, 2.5) at an
    arbitrary pose; nearest point on a cap (also on the axis), on a rim, on the barrel."""
    r, L = 0.75, 2.5
    X_WG = ho.xform(_angle_axis(math.pi / 5, [1, 2, 3]), [0.5, 1.25, -2])
    fn = lambda X, q: pp.distance_to_cylinder(r, L, X, q)
    for sgn in (-1.0, 1.0):
        N = [r * 0.25, r * 0.25, sgn * L / 2]
        _distance_case(fn, X_WG, N, [0, 0, sgn * 0.75])
        _distance_case(fn, X_WG, N, [0, 0, 0])
        _distance_case(fn, X_WG, N, [0, 0, -0.1 * sgn], inside=True)
        _distance_case(fn, X_WG, [0, 0, L / 2], [0, 0, 1.5])
        _distance_case(fn, X_WG, [0, 0, L / 2], [0, 0, 0])
        _distance_case(fn, X_WG, [0, 0, L / 2], [0, 0, -0.1], inside=True)
    theta = math.pi / 7
    for sgn in (-1.0, 1.0):
        p_NQ = np.array([math.cos(theta), math.sin(theta), sgn])
        _distance_case(fn, X_WG, p_NQ * [r, r, L / 2], p_NQ)
        _distance_case(fn, X_WG, p_NQ * [r, r, L / 2], np.zeros(3))
        p_NQ = np.array([math.cos(theta), math.sin(theta), 0.0])
        N = np.array([r, r, L / 4]) * [math.cos(theta), math.sin(theta), sgn]
        _distance_case(fn, X_WG, N, p_NQ)
        _distance_case(fn, X_WG, N, np.zeros(3))
        _distance_case(fn, X_WG, N, -0.1 * p_NQ, inside=True)


def test_sphere_cylinder_and_sphere_capsule_target_depths():
    """SphereCylinderDouble (:276-290,444-449) and SphereCapsuleDouble (:311-325,468-473): the other shape placed along
    its rotated x axis so that the sphere penetrates its barrel by 0.1; both operand orders."""
    I = ho.xform(None, [0, 0, 0])
    target = 0.1
    R_WB = rot("z", 0.2)
    for kind, size in (("cylinder", CYLINDER), ("capsule", CAPSULE)):
        X_WB = ho.xform(R_WB, R_WB @ np.array([1.0, 0, 0]) * (K_RADIUS + size[0] - target))
        a = pp.point_pair(("sphere", K_RADIUS), I, (kind,) + size, X_WB)
        _check_pair(a, target)
        b = pp.point_pair((kind,) + size, X_WB, ("sphere", K_RADIUS), I)
        assert b["depth"] == a["depth"]
        np.testing.assert_array_equal(b["p_WCa"], a["p_WCb"])
        np.testing.assert_array_equal(b["nhat_BA_W"], -a["nhat_BA_W"])


def _exe():
    exe = os.path.join(ROOT, "tests", "cpp", "_build", "point_pair_host_check")
    src = os.path.join(ROOT, "tests", "cpp", "point_pair_host_check.cc")
    hdr = os.path.join(ROOT, "include", "drake_b200_point_pair.h")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-Wall", "-Wextra", "-Werror", "-I",
                        os.path.join(ROOT, "include"), src, "-o", exe], check=True)
    return exe


def test_adapter_arithmetic_matches_the_oracle():
    """include/drake_b200_point_pair.h (what ProximityEngine::ComputeContactSurfacesWithFallback of the C++ mirror runs)
    against the numpy restatement on 5000 random pairs -- sphere x {sphere, box, half space, cylinder, capsule}, both operand orders,
    centres inside / on / outside the other shape -- to 4 ulp of the scale (numpy associates its products differently)."""
    rng = np.random.default_rng(42)
    from tests import parity
    lines, cases = [], []
    for k in range(5000):
        kind = ("sphere", "box", "half_space", "cylinder", "capsule")[k % 5]
        r = float(rng.uniform(0.05, 0.6))
        other = {"sphere": lambda: ("sphere", float(rng.uniform(0.05, 0.6))),
                 "box": lambda: ("box", tuple(float(x) for x in rng.uniform(0.1, 0.8, 3))),
                 "half_space": lambda: ("half_space",),
                 "cylinder": lambda: ("cylinder", float(rng.uniform(0.05, 0.5)), float(rng.uniform(0.1, 1.0))),
                 "capsule": lambda: ("capsule", float(rng.uniform(0.05, 0.5)), float(rng.uniform(0.1, 1.0)))}[kind]()
        X_S = parity.random_pose(rng, 0.5)
        X_O = parity.random_pose(rng, 0.5)
        if k % 7 == 0:   # sphere centre exactly on a face / at the centre of the other shape
            X_O = ho.xform(None, [0, 0, 0])
            X_S = ho.xform(None, [0, 0, other[1][2] / 2 if kind == "box" else (other[2] / 2 if kind in ("cylinder", "capsule")
                                                                              and k % 14 == 0 else 0.0)])
        pair = ((("sphere", r), X_S), (other, X_O))
        if k % 2:
            pair = pair[::-1]
        cases.append(pair)

        def fmt(shape, X):
            head = {"sphere": lambda: f"s {shape[1]!r}", "box": lambda: "b " + " ".join(repr(v) for v in shape[1]),
                    "half_space": lambda: "h", "cylinder": lambda: f"c {shape[1]!r} {shape[2]!r}",
                    "capsule": lambda: f"k {shape[1]!r} {shape[2]!r}"}[shape[0]]()
            return head + " " + " ".join(repr(float(v)) for v in X)
        lines.append(fmt(*pair[0]) + " " + fmt(*pair[1]))
    out = subprocess.run([_exe()], input="\n".join(lines) + "\n", capture_output=True, text=True, check=True).stdout.split("\n")
    hits = 0
    for (a, b), line in zip(cases, out):
        want = pp.point_pair(a[0], a[1], b[0], b[1])
        if want is None:
            assert line.strip() == "none"
            continue
        got = np.array([float(x) for x in line.split()])
        ref = np.concatenate([[want["depth"]], want["p_WCa"], want["p_WCb"], want["nhat_BA_W"]])
        scale = 1.0 + np.max(np.abs(ref))
        assert np.max(np.abs(got - ref)) <= 4 * K_EPS * scale, (a, b)
        hits += 1
    assert hits > 1500


# ---- FCL's half-space primitives behind CalcDistanceFallback (the pairs (HalfSpace, X), X not a sphere) ----------------

def _box_penetration_poses():
    """BoxPenetrationTest (penetration_as_point_pair_callback_test.cc:575-910): a unit box whose corner (-.5,-.5,-.5) is
    its lowest point, kDepth = 1e-3 under the z = 0 plane; X_WB_1 (:822-836) and X_WB_2 (:806-820)."""
    k_depth = 1e-3
    p_BoC_B, p_WC = np.array([-0.5, -0.5, -0.5]), np.array([0.0, 0.0, -k_depth])
    R_1 = _angle_axis(-math.pi / 4, [0, 1, 0]) @ _angle_axis(math.pi / 4, [1, 0, 0])
    R_2 = _angle_axis(math.atan(math.sqrt(2)), [math.sqrt(0.5), -math.sqrt(0.5), 0])
    return k_depth, [ho.xform(R, p_WC - R @ p_BoC_B) for R in (R_1, R_2)]


def test_box_on_tangent_plane_reference_scenes():
    """TangentPlane1 / TangentPlane2 (:906-912): A = the plane, B = the box; expected p_WCa = 0, p_WCb = (0, 0, -depth),
    nhat_BA_W = (0, 0, -1), depth = kDepth, all to 1e-12 -- and the reversed encoding (:741-746)."""
    k_depth, poses = _box_penetration_poses()
    I = ho.xform(None, [0, 0, 0])
    for X_WB in poses:
        R_WB, p_WB = X_WB[:9].reshape(3, 3).T, X_WB[9:]   # the fixture's own check (:661-677): every other corner has z > 0
        corners = np.array([[x, y, z] for x in (-0.5, 0.5) for y in (-0.5, 0.5) for z in (-0.5, 0.5)])[1:]
        assert np.all((corners @ R_WB.T + p_WB)[:, 2] > 0)
        res = pp.point_pair(("half_space",), I, ("box", (1.0, 1.0, 1.0)), X_WB)
        np.testing.assert_allclose(res["nhat_BA_W"], [0, 0, -1], rtol=0, atol=1e-12)
        np.testing.assert_allclose(res["p_WCa"], [0, 0, 0], rtol=0, atol=1e-12)
        np.testing.assert_allclose(res["p_WCb"], [0, 0, -k_depth], rtol=0, atol=1e-12)
        assert abs(res["depth"] - k_depth) <= 1e-12
        rev = pp.point_pair(("box", (1.0, 1.0, 1.0)), X_WB, ("half_space",), I)
        np.testing.assert_array_equal(rev["nhat_BA_W"], -res["nhat_BA_W"])
        np.testing.assert_array_equal(rev["p_WCa"], res["p_WCb"])
        np.testing.assert_array_equal(rev["p_WCb"], res["p_WCa"])
        assert rev["depth"] == res["depth"]


def _surface_samples(shape, rng, n=200000):
    """Points on the boundary of a shape in its own frame (dense enough for a 1e-3 brute-force check of the depth)."""
    kind = shape[0]
    if kind == "box":
        h = np.asarray(shape[1]) / 2
        P = rng.uniform(-1, 1, (n, 3)) * h
        ax = rng.integers(0, 3, n)
        P[np.arange(n), ax] = np.where(rng.random(n) < 0.5, -1, 1) * h[ax]
        corners = np.array([[sx, sy, sz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)]) * h
        return np.vstack([P, corners])
    if kind == "ellipsoid":
        U = rng.normal(size=(n, 3))
        return U / np.linalg.norm(U, axis=1)[:, None] * np.asarray(shape[1])
    r, L = shape[1], shape[2]
    th = rng.uniform(0, 2 * math.pi, n)
    if kind == "cylinder":
        z = np.where(rng.random(n) < 0.5, -L / 2, L / 2)   # the extreme is always on a rim (or the barrel's edge lines)
        return np.stack([r * np.cos(th), r * np.sin(th), z], axis=1)
    U = rng.normal(size=(n, 3))
    U = U / np.linalg.norm(U, axis=1)[:, None] * r
    U[:, 2] += np.where(U[:, 2] >= 0, L / 2, -L / 2)
    return U


def _support_depth(shape, X_WS, X_WH):
    """Depth from the shape's analytic support function h(-n) - n . c (independent of both restatements)."""
    R_WS, p_WS = X_WS[:9].reshape(3, 3).T, X_WS[9:]
    n_W = X_WH[:9].reshape(3, 3).T[:, 2]
    n = R_WS.T @ n_W
    height = float(n_W @ (p_WS - X_WH[9:]))
    kind = shape[0]
    if kind == "box":
        sup = float(np.abs(n) @ (np.asarray(shape[1]) / 2))
    elif kind == "ellipsoid":
        sup = math.sqrt(float((n * np.asarray(shape[1])) @ (n * np.asarray(shape[1]))))
    elif kind == "capsule":
        sup = shape[1] + shape[2] / 2 * abs(n[2])
    else:
        sup = shape[1] * math.sqrt(max(0.0, 1 - n[2] * n[2])) + shape[2] / 2 * abs(n[2])
    return sup - height


@pytest.mark.parametrize("kind", ["box", "cylinder", "capsule", "ellipsoid"])
def test_half_space_primitives_against_support_functions_and_sampling(kind):
    """What the geometry fixes regardless of FCL's coding: the depth is the support function's, the normal is the plane's,
    the two witness points are `depth` apart along it, p on the plane side lies ON the plane, the other is a deepest
    point of the shape (brute force over 2e5 boundary samples)."""
    from tests import parity
    rng = np.random.default_rng(7)
    hits = 0
    for k in range(40):
        shape = {"box": ("box", tuple(rng.uniform(0.2, 0.9, 3))), "cylinder": ("cylinder", rng.uniform(0.1, 0.5), rng.uniform(0.2, 1.2)),
                 "capsule": ("capsule", rng.uniform(0.1, 0.5), rng.uniform(0.2, 1.2)),
                 "ellipsoid": ("ellipsoid", tuple(rng.uniform(0.1, 0.6, 3)))}[kind]
        X_WH = parity.random_pose(rng, 0.3)
        X_WS = parity.random_pose(rng, 0.3)
        n_W = X_WH[:9].reshape(3, 3).T[:, 2]
        want_depth = _support_depth(shape, X_WS, X_WH)
        res = pp.point_pair(("half_space",), X_WH, shape, X_WS)
        if want_depth < 0:
            assert res is None
            continue
        hits += 1
        assert abs(res["depth"] - want_depth) <= 1e-13
        np.testing.assert_allclose(res["nhat_BA_W"], -n_W, rtol=0, atol=1e-15)
        np.testing.assert_allclose(res["p_WCb"] - res["p_WCa"], res["depth"] * res["nhat_BA_W"], rtol=0, atol=1e-13)
        assert abs(float(n_W @ (res["p_WCa"] - X_WH[9:]))) <= 1e-13          # A's witness is on the plane
        S = _surface_samples(shape, rng) @ X_WS[:9].reshape(3, 3) + X_WS[9:]   # rows: R p + t
        sd = (S - X_WH[9:]) @ n_W
        assert abs(-sd.min() - res["depth"]) <= 2e-3
        if kind != "ellipsoid":   # FCL's ellipsoid position is on the centre-to-deepest-point ray, not under the deepest point
            assert np.linalg.norm(S[np.argmin(sd)] - res["p_WCb"]) <= 0.08
        else:
            assert abs(float(n_W @ (res["p_WCb"] - X_WH[9:])) + res["depth"]) <= 1e-13
    assert hits >= 10


def test_half_space_primitives_degenerate_orientations():
    """The branches FCL takes at its 1e-7 tolerance: a box lying flat reports the face centre, a prone cylinder / capsule
    the point under its centre, a standing cylinder the disc centre; osculation (depth <= epsilon) is no contact
    (penetration_as_point_pair_callback.cc:179-185); a vertex cloud reports its first lowest vertex."""
    I = ho.xform(None, [0, 0, 0])
    d = 0.01
    flat = pp.point_pair(("half_space",), I, ("box", (0.4, 0.6, 0.2)), ho.xform(None, [0.3, -0.2, 0.1 - d]))
    np.testing.assert_allclose(flat["p_WCb"], [0.3, -0.2, -d], rtol=0, atol=1e-15)
    np.testing.assert_allclose(flat["p_WCa"], [0.3, -0.2, 0.0], rtol=0, atol=1e-15)
    assert abs(flat["depth"] - d) <= 1e-15
    prone = ho.xform(rot("x", math.pi / 2), [0.1, 0.2, 0.25 - d])
    for kind in ("cylinder", "capsule"):
        res = pp.point_pair((kind, 0.25, 0.8), prone, ("half_space",), I)
        np.testing.assert_allclose(res["p_WCa"], [0.1, 0.2, -d], rtol=0, atol=1e-15)
        np.testing.assert_allclose(res["nhat_BA_W"], [0, 0, 1], rtol=0, atol=0)
    standing = pp.point_pair(("half_space",), I, ("cylinder", 0.25, 0.8), ho.xform(None, [0.1, 0.2, 0.4 - d]))
    np.testing.assert_allclose(standing["p_WCb"], [0.1, 0.2, -d], rtol=0, atol=1e-15)
    upside_down = pp.point_pair(("half_space",), I, ("capsule", 0.25, 0.8), ho.xform(rot("x", math.pi), [0.1, 0.2, 0.65 - d]))
    np.testing.assert_allclose(upside_down["p_WCb"], [0.1, 0.2, -d], rtol=0, atol=1e-15)
    assert pp.point_pair(("half_space",), I, ("box", (1.0, 1.0, 1.0)), ho.xform(None, [0, 0, 0.5])) is None   # osculation
    assert pp.point_pair(("half_space",), I, ("box", (1.0, 1.0, 1.0)), ho.xform(None, [0, 0, 0.5 + 1e-9])) is None
    V = [[0, 0, 0.5], [1, 0, -0.2], [0, 1, -0.2], [1, 1, 0.1]]
    cloud = pp.point_pair(("half_space",), I, ("vertices", V), I)
    np.testing.assert_allclose(cloud["p_WCb"], [1, 0, -0.2], rtol=0, atol=1e-16)
    assert cloud["depth"] == 0.2
    with pytest.raises(NotImplementedError):
        pp.point_pair(("half_space",), I, ("half_space",), I)


def test_adapter_half_space_arithmetic_matches_the_oracle():
    """include/drake_b200_point_pair.h (World-frame formulation) against the oracle (half-space-frame formulation) on 4000
    random pairs of a half space with a box, cylinder, capsule, ellipsoid or vertex cloud, both operand orders, incl.
    the axis-aligned orientations that take FCL's tolerance branches, plus the reference's TangentPlane scenes."""
    from tests import parity
    rng = np.random.default_rng(43)
    lines, cases = [], []

    def fmt(shape, X):
        head = {"box": lambda: "b " + " ".join(repr(float(v)) for v in shape[1]), "half_space": lambda: "h",
                "cylinder": lambda: f"c {float(shape[1])!r} {float(shape[2])!r}", "capsule": lambda: f"k {float(shape[1])!r} {float(shape[2])!r}",
                "ellipsoid": lambda: "e " + " ".join(repr(float(v)) for v in shape[1]),
                "vertices": lambda: f"v {len(shape[1])} " + " ".join(repr(float(v)) for v in np.asarray(shape[1]).reshape(-1))}[shape[0]]()
        return head + " " + " ".join(repr(float(v)) for v in X)

    k_depth, tangent = _box_penetration_poses()
    I = ho.xform(None, [0, 0, 0])
    for X_WB in tangent:
        cases.append(((("half_space",), I), (("box", (1.0, 1.0, 1.0)), X_WB)))
    for k in range(4000):
        kind = ("box", "cylinder", "capsule", "ellipsoid", "vertices")[k % 5]
        shape = {"box": lambda: ("box", tuple(rng.uniform(0.1, 0.8, 3))), "cylinder": lambda: ("cylinder", rng.uniform(0.05, 0.5), rng.uniform(0.1, 1.0)),
                 "capsule": lambda: ("capsule", rng.uniform(0.05, 0.5), rng.uniform(0.1, 1.0)),
                 "ellipsoid": lambda: ("ellipsoid", tuple(rng.uniform(0.1, 0.6, 3))),
                 "vertices": lambda: ("vertices", rng.uniform(-0.4, 0.4, (int(rng.integers(4, 40)), 3)))}[kind]()
        X_H = parity.random_pose(rng, 0.4)
        X_S = parity.random_pose(rng, 0.4)
        if k % 6 == 0:    # axis-aligned with the plane: quarter turns about x / y relative to the half space
            Rq = rot("xy"[k % 2], (k // 6 % 4) * math.pi / 2)
            R_WH = X_H[:9].reshape(3, 3).T
            X_S = ho.xform(R_WH @ Rq, X_S[9:])
        pair = ((("half_space",), X_H), (shape, X_S))
        cases.append(pair[::-1] if k % 2 else pair)
    for a, b in cases:
        lines.append(fmt(*a) + " " + fmt(*b))
    out = subprocess.run([_exe()], input="\n".join(lines) + "\n", capture_output=True, text=True, check=True).stdout.split("\n")
    hits = 0
    for (a, b), line in zip(cases, out):
        want = pp.point_pair(a[0], a[1], b[0], b[1])
        if want is None:
            assert line.strip() == "none", (a, b, line)
            continue
        assert line.strip() != "none", (a, b)
        got = np.array([float(x) for x in line.split()])
        ref = np.concatenate([[want["depth"]], want["p_WCa"], want["p_WCb"], want["nhat_BA_W"]])
        scale = 1.0 + np.max(np.abs(ref))
        assert np.max(np.abs(got - ref)) <= 16 * K_EPS * scale, (a, b, got - ref)
        hits += 1
    assert hits > 1200
    first = np.array([float(x) for x in out[0].split()])   # TangentPlane1 through the C++: the reference's expectation
    np.testing.assert_allclose(first, [k_depth, 0, 0, 0, 0, 0, -k_depth, 0, 0, -1], rtol=0, atol=1e-12)


# ---- FCL's box x box behind CalcDistanceFallback ----------------------------------------------------------------------

def test_box_on_tangent_box_reference_scenes():
    """TangentBox1 / TangentBox2 (penetration_as_point_pair_callback_test.cc:914-920): A = a box of side 10 whose top face
    is the z = 0 plane (:872-876), B = the unit box on its corner; same expectation as the plane scenes, 1e-12."""
    k_depth, poses = _box_penetration_poses()
    X_WA = ho.xform(None, [0, 0, -5.0])
    for X_WB in poses:
        res = pp.point_pair(("box", (10.0, 10.0, 10.0)), X_WA, ("box", (1.0, 1.0, 1.0)), X_WB)
        np.testing.assert_allclose(res["nhat_BA_W"], [0, 0, -1], rtol=0, atol=1e-12)
        np.testing.assert_allclose(res["p_WCa"], [0, 0, 0], rtol=0, atol=1e-12)
        np.testing.assert_allclose(res["p_WCb"], [0, 0, -k_depth], rtol=0, atol=1e-12)
        assert abs(res["depth"] - k_depth) <= 1e-12
        rev = pp.point_pair(("box", (1.0, 1.0, 1.0)), X_WB, ("box", (10.0, 10.0, 10.0)), X_WA)
        np.testing.assert_allclose(rev["nhat_BA_W"], [0, 0, 1], rtol=0, atol=1e-12)
        np.testing.assert_allclose(rev["p_WCa"], [0, 0, -k_depth], rtol=0, atol=1e-12)
        np.testing.assert_allclose(rev["p_WCb"], [0, 0, 0], rtol=0, atol=1e-12)


def _box_vertices(size, X):
    h = np.asarray(size) / 2
    C = np.array([[sx, sy, sz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)]) * h
    return C @ X[:9].reshape(3, 3) + X[9:]


def _axis_overlaps(size_A, X_A, size_B, X_B):
    """Overlap of the two boxes' projections on each of the 15 candidate axes, from their 8 + 8 vertices (no shared code
    with either restatement). Returns (face overlaps [6], edge overlaps [<= 9])."""
    RA, RB = X_A[:9].reshape(3, 3).T, X_B[:9].reshape(3, 3).T
    VA, VB = _box_vertices(size_A, X_A), _box_vertices(size_B, X_B)

    def overlap(L):
        a, b = VA @ L, VB @ L
        return min(a.max() - b.min(), b.max() - a.min())
    faces = [overlap(RA[:, i]) for i in range(3)] + [overlap(RB[:, i]) for i in range(3)]
    edges = []
    for i in range(3):
        for j in range(3):
            L = np.cross(RA[:, i], RB[:, j])
            if np.linalg.norm(L) > 1e-6:
                edges.append(overlap(L / np.linalg.norm(L)))
    return np.array(faces), np.array(edges)


def test_box_box_is_the_minimum_overlap_axis_and_its_witnesses_lie_on_the_boxes():
    """What the geometry fixes whatever FCL's coding. The contact normal is the axis of smallest overlap over the 15 (an
    edge axis only when more than 5 % smaller than the best so far). The depth is that overlap whenever the deepest
    vertex of the incident face lies over the reference face (most poses) -- then moving B by it along the normal
    separates the boxes -- and otherwise the depth of the deepest point of the incident face that does (ODE reports
    per-point depths), never more than the overlap. For a face contact both witness points lie on their box (Drake's
    own DistanceToPoint(Box)), for an edge contact within the 1e-6 fudge."""
    from tests import parity
    rng = np.random.default_rng(11)
    hits = edge_hits = full_depth = 0
    for k in range(600):
        sa, sb = tuple(rng.uniform(0.2, 0.9, 3)), tuple(rng.uniform(0.2, 0.9, 3))
        X_A, X_B = parity.random_pose(rng, 0.25), parity.random_pose(rng, 0.25)
        faces, edges = _axis_overlaps(sa, X_A, sb, X_B)
        res = pp.point_pair(("box", sa), X_A, ("box", sb), X_B)
        if min(faces.min(), edges.min() if len(edges) else np.inf) < -1e-9:
            assert res is None
            continue
        if res is None:
            continue
        hits += 1
        best_face = faces.min()
        # ODE walks the axes in order; an edge axis replaces the current best only when its overlap is more than 5 %
        # smaller (so a later, slightly smaller edge overlap does not displace an earlier edge axis).
        sat, edge_wins, borderline = best_face, False, False
        for e in edges:
            borderline = borderline or abs(e * 1.05 - sat) < 1e-4   # within the 1e-6 fudge: either is a faithful answer
            if e * 1.05 < sat:
                sat, edge_wins = e, True
        if borderline:
            continue
        tol = 2e-5 if edge_wins else 1e-12
        assert res["depth"] <= sat + tol, (k, res["depth"], sat)
        n_AB = -res["nhat_BA_W"]
        assert abs(np.linalg.norm(n_AB) - 1) <= 1e-12
        # The normal is the minimum-overlap axis: the overlap along it, measured from the vertices, is `sat`.
        a, b = _box_vertices(sa, X_A) @ n_AB, _box_vertices(sb, X_B) @ n_AB
        assert abs((a.max() - b.min()) - sat) <= tol, (k, a.max() - b.min(), sat)
        if abs(res["depth"] - sat) <= tol:
            full_depth += 1
            shifted = X_B.copy()
            shifted[9:] += (res["depth"] + 1e-5) * n_AB
            f2, e2 = _axis_overlaps(sa, X_A, sb, shifted)
            assert min(f2.min(), e2.min() if len(e2) else np.inf) < 0          # separated after the push
        np.testing.assert_allclose(res["p_WCb"] - res["p_WCa"], res["depth"] * res["nhat_BA_W"], rtol=0, atol=1e-12)
        # Witnesses on the boxes: always for a face contact; for an edge contact when the axis is the true minimum over
        # all 15 (ODE takes the closest points of the two edges' LINES: on a non-minimal edge axis, which the 5 % rule can
        # keep, they may fall outside the edges).
        if not edge_wins or sat <= min(faces.min(), edges.min()) + 1e-12:
            assert abs(pp.distance_to_box(sa, X_A, res["p_WCa"])[1]) <= tol
            assert abs(pp.distance_to_box(sb, X_B, res["p_WCb"])[1]) <= tol
        edge_hits += edge_wins
    assert hits > 150 and edge_hits > 5 and full_depth > hits // 2


def test_box_box_resting_flat_and_separated():
    """Two faces flat on each other: depth and normal are the stacking axis', the reported point is a point of the overlap
    of the two faces (which one FCL picks among the equally deep ones is not pinned); osculation and separation give none;
    """
    I = ho.xform(None, [0, 0, 0])
    top = ho.xform(rot("z", 0.3), [0.1, -0.05, 0.5 + 0.25 - 0.01])
    res = pp.point_pair(("box", (1.0, 1.0, 1.0)), I, ("box", (0.5, 0.5, 0.5)), top)
    assert abs(res["depth"] - 0.01) <= 1e-15
    np.testing.assert_allclose(res["nhat_BA_W"], [0, 0, -1], rtol=0, atol=1e-15)
    assert abs(res["p_WCa"][2] - 0.5) <= 1e-15 and abs(res["p_WCb"][2] - 0.49) <= 1e-15
    assert abs(pp.distance_to_box((0.5, 0.5, 0.5), top, res["p_WCb"])[1]) <= 1e-15
    assert pp.point_pair(("box", (1.0, 1.0, 1.0)), I, ("box", (0.5, 0.5, 0.5)), ho.xform(None, [0, 0, 0.75])) is None
    assert pp.point_pair(("box", (1.0, 1.0, 1.0)), I, ("box", (0.5, 0.5, 0.5)), ho.xform(None, [0, 0, 0.8])) is None


def test_adapter_box_box_arithmetic_matches_the_oracle():
    """include/drake_b200_point_pair.h::BoxBox (ODE's 2-D incident-face route) against the oracle (3-D clipping) on 3000
    random box pairs, both operand orders, plus the reference's TangentBox scenes through the C++."""
    from tests import parity
    rng = np.random.default_rng(44)
    k_depth, tangent = _box_penetration_poses()
    cases = [((("box", (10.0, 10.0, 10.0)), ho.xform(None, [0, 0, -5.0])), (("box", (1.0, 1.0, 1.0)), X)) for X in tangent]
    for k in range(3000):
        a = (("box", tuple(float(x) for x in rng.uniform(0.15, 0.9, 3))), parity.random_pose(rng, 0.3))
        b = (("box", tuple(float(x) for x in rng.uniform(0.15, 0.9, 3))), parity.random_pose(rng, 0.3))
        cases.append((b, a) if k % 2 else (a, b))
    lines = []
    for a, b in cases:
        lines.append(" ".join("b " + " ".join(repr(float(v)) for v in g[0][1]) + " " + " ".join(repr(float(v)) for v in g[1]) for g in (a, b)))
    out = subprocess.run([_exe()], input="\n".join(lines) + "\n", capture_output=True, text=True, check=True).stdout.split("\n")
    hits = 0
    for (a, b), line in zip(cases, out):
        want = pp.point_pair(a[0], a[1], b[0], b[1])
        if want is None:
            assert line.strip() == "none", (a, b, line)
            continue
        assert line.strip() != "none", (a, b)
        got = np.array([float(x) for x in line.split()])
        ref = np.concatenate([[want["depth"]], want["p_WCa"], want["p_WCb"], want["nhat_BA_W"]])
        assert np.max(np.abs(got - ref)) <= 64 * K_EPS * (1.0 + np.max(np.abs(ref))), (a, b, got - ref)
        hits += 1
    assert hits > 800
    first = np.array([float(x) for x in out[0].split()])
    np.testing.assert_allclose(first, [k_depth, 0, 0, 0, 0, 0, -k_depth, 0, 0, -1], rtol=0, atol=1e-12)


# ---- every other convex pair: libccd's Minkowski Portal Refinement behind fcl::collide ------------------------------------

def _fmt_shape(shape, X):
    head = {"sphere": lambda: f"s {float(shape[1])!r}", "box": lambda: "b " + " ".join(repr(float(v)) for v in shape[1]),
            "half_space": lambda: "h", "cylinder": lambda: f"c {float(shape[1])!r} {float(shape[2])!r}",
            "capsule": lambda: f"k {float(shape[1])!r} {float(shape[2])!r}",
            "ellipsoid": lambda: "e " + " ".join(repr(float(v)) for v in shape[1]),
            "vertices": lambda: f"v {len(shape[1])} " + " ".join(repr(float(v)) for v in np.asarray(shape[1]).reshape(-1))}[shape[0]]()
    return head + " " + " ".join(repr(float(v)) for v in X)


def _through_the_header(cases):
    lines = [_fmt_shape(*a) + " " + _fmt_shape(*b) for a, b in cases]
    out = subprocess.run([_exe()], input="\n".join(lines) + "\n", capture_output=True, text=True, check=True).stdout.split("\n")
    res = []
    for line in out[:len(cases)]:
        if line.strip() == "none":
            res.append(None)
        else:
            g = np.array([float(x) for x in line.split()])
            res.append({"depth": g[0], "p_WCa": g[1:4], "p_WCb": g[4:7], "nhat_BA_W": g[7:10]})
    return res


def _tangent_scenes():
    """BoxPenetrationTest's tangent shapes other than plane / box / sphere (penetration_as_point_pair_callback_test.cc:
    846-905) with the tolerance the reference states for box pose 1 and box pose 2 (:933-984)."""
    k_radius, k_length = 1.0, 10.0
    prone = ho.xform(_angle_axis(math.pi / 2, [1, 0, 0]), [0, 0, -k_radius])
    cube = [[x, y, z] for x in (-5.0, 5.0) for y in (-5.0, 5.0) for z in (-5.0, 5.0)]   # quad_cube.obj scaled by 5
    return [("standing cylinder", ("cylinder", k_radius, k_length), ho.xform(None, [0, 0, -k_length / 2]), (1e-3, 1e-12)),
            ("prone cylinder", ("cylinder", k_radius, k_length), prone, (1e-1, 1e-4)),
            ("convex", ("vertices", cube), ho.xform(None, [0, 0, -k_length / 2]), (1e-3, 1e-3)),
            ("standing capsule", ("capsule", k_radius, k_length), ho.xform(None, [0, 0, -k_length / 2 - k_radius]), (1e-1, 1e-12)),
            ("prone capsule", ("capsule", k_radius, k_length), prone, (1e-1, 1e-4))]


def test_mpr_reference_tangent_scenes():
    """TangentStandingCylinder1/2, TangentProneCylinder1/2, TangentConvex1/2, TangentStandingCapsule1/2,
    TangentProneCapsule1/2 (:933-984): normal, both witness points and the depth within the tolerance the reference itself
    states for each, through the oracle's restatement AND the product header."""
    k_depth, poses = _box_penetration_poses()
    cases, tols, names = [], [], []
    for name, shape, X_WA, tol in _tangent_scenes():
        for i, X_WB in enumerate(poses):
            cases.append(((shape, X_WA), (("box", (1.0, 1.0, 1.0)), X_WB)))
            tols.append(tol[i])
            names.append(f"{name} {i + 1}")
    want = {"nhat_BA_W": [0, 0, -1], "p_WCa": [0, 0, 0], "p_WCb": [0, 0, -k_depth]}
    for results in ([pp.point_pair(a[0], a[1], b[0], b[1]) for a, b in cases], _through_the_header(cases)):
        for res, tol, name in zip(results, tols, names):
            assert res is not None, name
            for key, value in want.items():
                np.testing.assert_allclose(res[key], value, rtol=0, atol=tol, err_msg=name)
            assert abs(res["depth"] - k_depth) <= tol, name


def _support_height(shape, X, n_W):
    """max over the shape of n . x, from the analytic support FUNCTIONS (independent of the support POINTS MPR uses)."""
    R, t = X[:9].reshape(3, 3).T, X[9:]
    n = R.T @ n_W
    kind = shape[0]
    h = {"sphere": lambda: shape[1], "box": lambda: float(np.abs(n) @ (np.asarray(shape[1]) / 2)),
         "cylinder": lambda: shape[1] * math.hypot(n[0], n[1]) + shape[2] / 2 * abs(n[2]),
         "capsule": lambda: shape[1] + shape[2] / 2 * abs(n[2]),
         "ellipsoid": lambda: float(np.linalg.norm(n * np.asarray(shape[1]))),
         "vertices": lambda: float(np.max(np.asarray(shape[1]) @ n))}[kind]()
    return float(n_W @ t) + h


def _random_convex(rng, kind):
    return {"sphere": lambda: ("sphere", rng.uniform(0.1, 0.4)), "box": lambda: ("box", tuple(rng.uniform(0.2, 0.7, 3))),
            "cylinder": lambda: ("cylinder", rng.uniform(0.1, 0.35), rng.uniform(0.2, 0.8)),
            "capsule": lambda: ("capsule", rng.uniform(0.1, 0.3), rng.uniform(0.2, 0.7)),
            "ellipsoid": lambda: ("ellipsoid", tuple(rng.uniform(0.15, 0.45, 3))),
            "vertices": lambda: ("vertices", rng.uniform(-0.35, 0.35, (int(rng.integers(8, 30)), 3)))}[kind]()


MPR_PAIRS = [("box", "cylinder"), ("box", "capsule"), ("cylinder", "cylinder"), ("capsule", "capsule"), ("cylinder", "capsule"),
             ("sphere", "ellipsoid"), ("ellipsoid", "ellipsoid"), ("box", "ellipsoid"), ("vertices", "box"),
             ("vertices", "vertices"), ("sphere", "vertices"), ("ellipsoid", "capsule")]


def test_mpr_answers_are_consistent_with_the_shapes_support_functions():
    """What MPR promises, checked against analytic support functions on 1200 random (mostly DEEP) overlaps of twelve kinds
    of pairs through the product header: a reported contact has a unit normal along which the two shapes really overlap by
    at least the reported depth (the portal lies inside A - B); the witness points are `depth` apart along the normal;
    shapes reported apart have a direction with (numerically) no overlap. How far the overlap along the normal exceeds the
    depth is MPR's known looseness (it measures along the ray from the centres' difference, Drake issue 7656): the median
    excess is recorded as a regression guard, not as an accuracy claim."""
    from tests import parity
    rng = np.random.default_rng(21)
    cases = []
    for k in range(1200):
        ka, kb = MPR_PAIRS[k % len(MPR_PAIRS)]
        a, b = (_random_convex(rng, ka), parity.random_pose(rng, 0.3)), (_random_convex(rng, kb), parity.random_pose(rng, 0.3))
        cases.append((b, a) if k % 2 else (a, b))
    dirs = rng.normal(size=(4000, 3))
    dirs /= np.linalg.norm(dirs, axis=1)[:, None]
    hits, excess = 0, []
    for (a, b), res in zip(cases, _through_the_header(cases)):
        overlap = lambda n: _support_height(a[0], a[1], n) + _support_height(b[0], b[1], -n)   # max_A n.x - min_B n.x
        if res is None:
            best = min(dirs, key=overlap)
            assert overlap(best) < 2e-2, (a[0][0], b[0][0], overlap(best))   # (a coarse direction sample: nearly apart)
            continue
        hits += 1
        n_AB = -res["nhat_BA_W"]
        assert abs(np.linalg.norm(n_AB) - 1) <= 1e-12
        np.testing.assert_allclose(res["p_WCb"] - res["p_WCa"], res["depth"] * res["nhat_BA_W"], rtol=0, atol=1e-12)
        w = overlap(n_AB)
        assert w >= res["depth"] - 1e-9, (a[0][0], b[0][0], w, res["depth"])
        excess.append(w - res["depth"])
    assert hits > 300 and np.median(excess) < 1e-2 and np.max(excess) < 0.5, (hits, np.median(excess), np.max(excess))


def test_mpr_header_agrees_with_the_oracle_restatement():
    """The numpy restatement of the same published algorithm (oracle/ho_point_pair.py::mpr_contact) against the header on
    600 random pairs: the same overlap decision, and depth / normal / witnesses equal to 1e-6 for at least 97 % (the two sum
    in different orders; where a support point is tied -- box faces, cylinder rims, vertex clouds -- a refinement step
    decided differently leads to another portal, and for these deep overlaps to a visibly different ray-cast depth)."""
    from tests import parity
    rng = np.random.default_rng(22)
    cases = []
    for k in range(600):
        ka, kb = MPR_PAIRS[k % len(MPR_PAIRS)]
        cases.append(((_random_convex(rng, ka), parity.random_pose(rng, 0.3)), (_random_convex(rng, kb), parity.random_pose(rng, 0.3))))
    same = hits = 0
    for (a, b), got in zip(cases, _through_the_header(cases)):
        want = pp.point_pair(a[0], a[1], b[0], b[1])
        if want is None or got is None:
            assert want is None and got is None or (want or got)["depth"] < 1e-5, (a[0][0], b[0][0])
            continue
        hits += 1
        err = max(abs(got["depth"] - want["depth"]), np.max(np.abs(got["nhat_BA_W"] - want["nhat_BA_W"])),
                  np.max(np.abs(got["p_WCa"] - want["p_WCa"])), np.max(np.abs(got["p_WCb"] - want["p_WCb"])))
        same += err <= 1e-6
        assert abs(got["depth"] - want["depth"]) <= 2e-2 and float(got["nhat_BA_W"] @ want["nhat_BA_W"]) > 0.95
    assert hits > 150 and same >= 0.97 * hits, (hits, same)
