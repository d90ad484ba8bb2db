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
    against the numpy restatement on 3000 random pairs -- sphere x {sphere, box, half space}, both operand orders,
    centres inside / on / outside the other shape -- to 4 ulp of the scale (numpy associates its products differently)."""
    rng = np.random.default_rng(42)
    from tests import parity
    lines, cases = [], []
    for k in range(3000):
        kind = ("sphere", "box", "half_space")[k % 3]
        r = float(rng.uniform(0.05, 0.6))
        other = ("sphere", float(rng.uniform(0.05, 0.6))) if kind == "sphere" else \
            (("box", tuple(float(x) for x in rng.uniform(0.1, 0.8, 3))) if kind == "box" else ("half_space",))
        X_S = parity.random_pose(rng, 0.5)
        X_O = parity.random_pose(rng, 0.5)
        if k % 7 == 0:   # sphere centre exactly on a face / at the centre of the other shape
            X_O = ho.xform(None, [0, 0, 0])
            X_S = ho.xform(None, [0, 0, other[1][2] / 2 if kind == "box" else 0.0])
        pair = ((("sphere", r), X_S), (other, X_O))
        if k % 2:
            pair = pair[::-1]
        cases.append(pair)

        def fmt(shape, X):
            head = {"sphere": lambda: f"s {shape[1]!r}", "box": lambda: "b " + " ".join(repr(v) for v in shape[1]),
                    "half_space": lambda: "h"}[shape[0]]()
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
    assert hits > 1000
