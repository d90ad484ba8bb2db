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
