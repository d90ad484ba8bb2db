"""Oracle (test infrastructure): Drake's closed-form point-pair narrow phase for pairs with a sphere, restated in
numpy -- penetration_as_point_pair_callback.cc:92-118,203-300 (SphereShapePenetration, ComputeNarrowPhasePenetration)
over point_distance::DistanceToPoint for Sphere, Box, HalfSpace, Cylinder and Capsule
(distance_to_point_callback.cc:15-160,160-288,384-482). Pinned by the reference's own test scenes (penetration_as_point_pair_callback_test.cc:119-176,196-230,
258-310,430-462), restated in tests/test_point_pair_cpu.py. The general (non-sphere) fallback is FCL's GJK / EPA, an
external dependency of the reference: not restated.

Poses are 12 doubles: rotation column-major, then translation (math/rigid_transform.h:1004-1007)."""
from __future__ import annotations

import math

import numpy as np


def _R(X):
    return np.asarray(X[:9], float).reshape(3, 3).T


def _p(X):
    return np.asarray(X[9:12], float)


def relative_tolerance(size: float) -> float:
    """DistanceToPointRelativeTolerance (proximity_utilities.h:29-31)."""
    return 1e-14 * max(1.0, size)


def sign(x: float) -> float:
    """DistanceToPoint::Sign: Sign(0) = 1 (distance_to_point_callback.h:185-191)."""
    return -1.0 if x < 0.0 else 1.0


def distance_to_sphere(radius, X_WG, p_WQ):
    """(p_GN, distance, grad_W): distance_to_point_callback.cc:15-58."""
    p_GQ = _R(X_WG).T @ (np.asarray(p_WQ, float) - _p(X_WG))
    dist = math.sqrt(float(p_GQ @ p_GQ))
    grad_G = p_GQ / dist if dist > relative_tolerance(radius) else np.array([1.0, 0.0, 0.0])
    p_GN = radius * grad_G
    return p_GN, float((p_GQ - p_GN) @ grad_G), _R(X_WG) @ grad_G


def distance_to_half_space(X_WG, p_WQ):
    """distance_to_point_callback.cc:60-77 (the half space is z <= 0 of its frame)."""
    p_GQ = _R(X_WG).T @ (np.asarray(p_WQ, float) - _p(X_WG))
    return np.array([p_GQ[0], p_GQ[1], 0.0]), float(p_GQ[2]), _R(X_WG) @ np.array([0.0, 0.0, 1.0])


def distance_to_box(size, X_WG, p_WQ):
    """distance_to_point_callback.cc:160-178 + ComputeDistanceToBox (:406-482) + ExtremalAxis (:384-400)."""
    p_WQ = np.asarray(p_WQ, float)
    p_GQ = _R(X_WG).T @ (p_WQ - _p(X_WG))
    h = np.asarray(size, float) / 2.0
    loc, p_GC = [], np.zeros(3)
    for i in range(3):
        tol = relative_tolerance(h[i])
        if abs(p_GQ[i]) > h[i] + tol:
            loc.append("out")
            p_GC[i] = sign(p_GQ[i]) * h[i]
        elif abs(p_GQ[i]) >= h[i] - tol:
            loc.append("bnd")
            p_GC[i] = sign(p_GQ[i]) * h[i]
        else:
            loc.append("in")
            p_GC[i] = p_GQ[i]
    p_GN, grad_G = p_GC.copy(), np.zeros(3)
    if "out" in loc:
        d = p_GQ - p_GN
        grad_G = d / math.sqrt(float(d @ d))
    elif "bnd" in loc:
        for i in range(3):
            if loc[i] == "bnd":
                grad_G[i] = sign(p_GC[i])
        grad_G = grad_G / math.sqrt(float(grad_G @ grad_G))
    else:
        best, axis = math.inf, -1
        for i in range(3):
            for bound in (h[i], -h[i]):
                dist = abs(bound - p_GQ[i])
                if dist < best:
                    best, axis = dist, i
        s = sign(p_GQ[axis])
        p_GN[axis] = s * h[axis]
        grad_G[axis] = s
    grad_W = _R(X_WG) @ grad_G
    p_WN = _R(X_WG) @ p_GN + _p(X_WG)
    return p_GN, float(grad_W @ (p_WQ - p_WN)), grad_W


def _distance_to_box_nd(h, p):
    """ComputeDistanceToBox<dim> (:406-482) + ExtremalAxis (:387-403) in the box frame: (p_GN, grad_G)."""
    h, p = np.asarray(h, float), np.asarray(p, float)
    n = len(h)
    loc, c = [], np.zeros(n)
    for i in range(n):
        tol = relative_tolerance(h[i])
        if abs(p[i]) > h[i] + tol:
            loc.append("out")
            c[i] = sign(p[i]) * h[i]
        elif abs(p[i]) >= h[i] - tol:
            loc.append("bnd")
            c[i] = sign(p[i]) * h[i]
        else:
            loc.append("in")
            c[i] = p[i]
    p_GN, grad = c.copy(), np.zeros(n)
    if "out" in loc:
        d = p - c
        grad = d / math.sqrt(float(d @ d))
    elif "bnd" in loc:
        for i in range(n):
            if loc[i] == "bnd":
                grad[i] = sign(c[i])
        grad = grad / math.sqrt(float(grad @ grad))
    else:
        best, axis = math.inf, -1
        for i in range(n):
            for bound in (h[i], -h[i]):
                dist = abs(bound - p[i])
                if dist < best:
                    best, axis = dist, i
        s = sign(p[axis])
        p_GN[axis] = s * h[axis]
        grad[axis] = s
    return p_GN, grad


def distance_to_cylinder(radius, length, X_WG, p_WQ):
    """DistanceToPoint::operator()(Cylinder) (:194-288): the 2-D box [-r, r] x [-L/2, L/2] of the cross section through
    the query point and the axis (+x for a point on the axis)."""
    p_WQ = np.asarray(p_WQ, float)
    p_GQ = _R(X_WG).T @ (p_WQ - _p(X_WG))
    r_Q = math.sqrt(p_GQ[0] * p_GQ[0] + p_GQ[1] * p_GQ[1])
    v = np.array([1.0, 0.0]) if r_Q < relative_tolerance(radius) else np.array([p_GQ[0], p_GQ[1]]) / r_Q
    p_BN, grad_B = _distance_to_box_nd([radius, length / 2.0], [r_Q, p_GQ[2]])
    p_GN = np.array([p_BN[0] * v[0], p_BN[0] * v[1], p_BN[1]])
    grad_G = np.array([grad_B[0] * v[0], grad_B[0] * v[1], grad_B[1]])
    grad_W = _R(X_WG) @ grad_G
    p_WN = _R(X_WG) @ p_GN + _p(X_WG)
    return p_GN, float(grad_W @ (p_WQ - p_WN)), grad_W


def _sphere_in_its_frame(radius, p_SQ):
    """SphereDistanceInSphereFrame (:15-58)."""
    dist = math.sqrt(float(p_SQ @ p_SQ))
    grad = p_SQ / dist if dist > relative_tolerance(radius) else np.array([1.0, 0.0, 0.0])
    p_SN = radius * grad
    return p_SN, float((p_SQ - p_SN) @ grad), grad


def distance_to_capsule(radius, length, X_WG, p_WQ):
    """ComputeDistanceToPrimitive(Capsule) (:79-160): a cap sphere beyond the spine's ends, else the projection onto the
    x-y plane against a sphere at the origin, shifted back along z."""
    half = length / 2
    p_GQ = _R(X_WG).T @ (np.asarray(p_WQ, float) - _p(X_WG))
    if p_GQ[2] >= half or p_GQ[2] <= -half:
        p_GS = np.array([0.0, 0.0, half if p_GQ[2] >= half else -half])
        p_SN, distance, grad_G = _sphere_in_its_frame(radius, p_GQ - p_GS)
        p_GN = p_GS + p_SN
    else:
        p_GM, distance, grad_G = _sphere_in_its_frame(radius, np.array([p_GQ[0], p_GQ[1], 0.0]))
        p_GN = np.array([p_GM[0], p_GM[1], p_GQ[2]])
    return p_GN, distance, _R(X_WG) @ grad_G


def point_pair(shape_A, X_WA, shape_B, X_WB):
    """ComputeNarrowPhasePenetration for shapes given as ('sphere', r) | ('box', (x, y, z)) | ('half_space',) |
    ('cylinder', r, L) | ('capsule', r, L); A is the
    geometry with the smaller id. Returns None (no penetration) or dict(p_WCa, p_WCb, nhat_BA_W, depth). Raises for pairs
    without a sphere (the reference calls FCL there)."""
    a_is_sphere, b_is_sphere = shape_A[0] == "sphere", shape_B[0] == "sphere"
    if not (a_is_sphere or b_is_sphere):
        raise NotImplementedError("no sphere in the pair: FCL's general narrow phase")
    (S, X_WS), (O, X_WO) = ((shape_A, X_WA), (shape_B, X_WB)) if a_is_sphere else ((shape_B, X_WB), (shape_A, X_WA))
    p_WSo = _p(X_WS)
    if O[0] == "sphere":
        p_GN, distance, grad_W = distance_to_sphere(O[1], X_WO, p_WSo)
    elif O[0] == "box":
        p_GN, distance, grad_W = distance_to_box(O[1], X_WO, p_WSo)
    elif O[0] == "half_space":
        p_GN, distance, grad_W = distance_to_half_space(X_WO, p_WSo)
    elif O[0] == "cylinder":
        p_GN, distance, grad_W = distance_to_cylinder(O[1], O[2], X_WO, p_WSo)
    elif O[0] == "capsule":
        p_GN, distance, grad_W = distance_to_capsule(O[1], O[2], X_WO, p_WSo)
    else:
        raise NotImplementedError(O[0])
    r = S[1]
    depth = r - distance
    if not depth >= 0:
        return None
    out = {"depth": depth, "p_WCb": _R(X_WO) @ p_GN + _p(X_WO), "nhat_BA_W": grad_W, "p_WCa": p_WSo - r * grad_W}
    if not a_is_sphere:   # SwapAAndB
        out["p_WCa"], out["p_WCb"] = out["p_WCb"], out["p_WCa"]
        out["nhat_BA_W"] = -out["nhat_BA_W"]
    return out
