"""Oracle (test infrastructure): the point-pair narrow phase of ComputeContactSurfacesWithFallback restated in numpy.

(1) Drake's own closed forms for pairs with a sphere -- penetration_as_point_pair_callback.cc:92-118,203-300
(SphereShapePenetration, ComputeNarrowPhasePenetration) over point_distance::DistanceToPoint for Sphere, Box, HalfSpace,
Cylinder and Capsule (distance_to_point_callback.cc:15-160,160-288,384-482). Pinned by the reference's own test scenes
(penetration_as_point_pair_callback_test.cc:119-176,196-230,258-310,430-462), restated in tests/test_point_pair_cpu.py.

(2) Of the FCL half (CalcDistanceFallback<double>, :145-207; FCL is an external dependency of the reference, pinned in
tools/workspace/fcl_internal/repository.bzl at RobotLocomotion/fcl commit e5efcc41b57b and absent from its tree): FCL's
published closed-form half-space primitives (fcl/narrowphase/detail/primitive_shape_algorithm/halfspace-inl.h) for a
HalfSpace against Box, Capsule, Cylinder, Ellipsoid and a vertex cloud (Convex), worked in the HALF SPACE's frame --
the product header works in World, so the two are independent derivations. PARITY UNPINNED against FCL itself: pinned by
the reference's BoxPenetrationTest TangentPlane scenes (penetration_as_point_pair_callback_test.cc:575-910) and by the
shapes' analytic support functions / brute-force surface sampling in tests/test_point_pair_cpu.py.

(3) FCL's box x box (boxBox2, the ODE dBoxBox algorithm): box_box_contact below, pinned by the reference's TangentBox
scenes and by exact geometry (the minimum-overlap axis over all 15, vertices of one box inside the other).

(4) Every other convex pair: libccd's Minkowski Portal Refinement (ccdMPRPenetration, what FCL's GJKCollide calls for a
contact) over FCL's support functions: mpr_contact below, pinned by the reference's BoxPenetrationTest scenes for
cylinder / capsule / convex tangent shapes at the tolerances the reference states.

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


HALFSPACE_TOLERANCE = 1e-7   # fcl::detail::halfspaceIntersectTolerance<double>()


def _fcl_sign(a: float) -> float:
    """FCL's `(A > 0) ? -1 : 1`: which end of an axis is the lower one; zero goes to +1."""
    return -1.0 if a > 0 else 1.0


def half_space_contact(shape, X_HS):
    """FCL's shape x Halfspace primitives in the half space's own frame H (solid z <= 0): shape = ('box', (x, y, z)) |
    ('capsule', r, L) | ('cylinder', r, L) | ('ellipsoid', (a, b, c)) | ('vertices', [[x, y, z], ...]) posed at X_HS.
    Returns None or (p_H, depth): FCL's contact position (half the depth above the deepest point) and the depth."""
    R, t = _R(X_HS), _p(X_HS)
    ez = np.array([0.0, 0.0, 1.0])
    kind = shape[0]
    if kind == "box":
        h = np.asarray(shape[1], float) / 2.0
        nz = R[2, :]                                   # the plane normal in the box frame
        depth = 0.5 * float(abs(nz[0] * 2 * h[0]) + abs(nz[1] * 2 * h[1]) + abs(nz[2] * 2 * h[2])) - t[2]
        if depth < 0:
            return None
        axes = [i for i in range(3) if abs(abs(nz[i]) - 1) < HALFSPACE_TOLERANCE][:1] or [0, 1, 2]
        p = t.copy()
        for i in axes:
            p = p + R[:, i] * (h[i] * _fcl_sign(nz[i] * 2 * h[i]))
        return p + ez * (0.5 * depth), depth
    if kind in ("capsule", "cylinder"):
        r, L = float(shape[1]), float(shape[2])
        axis = R[:, 2]
        cosa = float(axis[2])
        if abs(cosa) < HALFSPACE_TOLERANCE:            # lying on its side: the point under the centre
            depth = r - t[2]
            return None if depth < 0 else (t + ez * (0.5 * depth - r), depth)
        end = t + axis * (0.5 * L * _fcl_sign(cosa))   # the lower end of the axis
        if kind == "capsule":
            depth = r - end[2]
            return None if depth < 0 else ((end - ez * r) + ez * (0.5 * depth), depth)
        if abs(abs(cosa) - 1) < HALFSPACE_TOLERANCE:
            rim = np.zeros(3)
        else:
            down = axis * cosa - ez                    # -z projected into the end disc
            rim = down * (r / math.sqrt(float(down @ down)))
        p = end + rim
        depth = -float(p[2])
        return None if depth < 0 else (p + ez * (0.5 * depth), depth)
    if kind == "ellipsoid":
        radii2 = np.asarray(shape[1], float) ** 2
        n = R.T @ ez                                   # plane n . x <= d in the ellipsoid frame
        d = -float(t[2])
        c = math.sqrt(float((n * n) @ radii2))         # height of the centre above the tangent plane
        depth = c + d
        if not depth >= 0:
            return None
        support = (radii2 * n) / c
        p_E = support * (0.5 * depth / float(n @ support) - 1.0)
        return R @ p_E + t, depth
    if kind == "vertices":
        V = np.asarray(shape[1], float).reshape(-1, 3)
        n, d = R.T @ ez, -float(t[2])
        sd = V @ n - d
        k = int(np.argmin(sd))                         # the first minimum
        if not sd[k] <= 0:
            return None
        depth = -float(sd[k])
        return (R @ V[k] + t) + ez * (0.5 * depth), depth
    raise NotImplementedError(kind)


def half_space_point_pair(shape_A, X_WA, shape_B, X_WB):
    """CalcDistanceFallback<double> (:145-207) for a pair exactly one of whose shapes is ('half_space',)."""
    a_is_hs = shape_A[0] == "half_space"
    X_WH, (shape, X_WS) = (X_WA, (shape_B, X_WB)) if a_is_hs else (X_WB, (shape_A, X_WA))
    R_WH, p_WH = _R(X_WH), _p(X_WH)
    R_HS = R_WH.T @ _R(X_WS)
    X_HS = np.concatenate([R_HS.T.reshape(-1), R_WH.T @ (_p(X_WS) - p_WH)])
    hit = half_space_contact(shape, X_HS)
    if hit is None:
        return None
    p_H, depth = hit
    if depth <= np.finfo(float).eps:                   # osculation is no contact (:179-185)
        return None
    n_W = R_WH @ np.array([0.0, 0.0, 1.0])
    pos = R_WH @ p_H + p_WH
    # FCL's normal points from object A into object B; nhat_BA_W is its negative (:188-190).
    nhat_BA_W = -n_W if a_is_hs else n_W
    return {"depth": depth, "nhat_BA_W": nhat_BA_W, "p_WCa": pos - 0.5 * depth * nhat_BA_W,
            "p_WCb": pos + 0.5 * depth * nhat_BA_W}


def box_box_contact(size_A, X_WA, size_B, X_WB):
    """FCL's box-box (boxBox2, the ODE dBoxBox algorithm; fcl/narrowphase/detail/primitive_shape_algorithm/box_box-inl.h)
    as fcl::collide() reports its deepest contact: None or (normal from A into B, position, depth). The 15 separating
    axes follow the published algorithm (face axes first; an edge x edge axis wins only by more than 5 %; parallel edges
    skipped; 1e-6 added to |R| for the edge tests). The contact POINT is derived here in 3-D World coordinates --
    Sutherland-Hodgman of the incident face against the four side planes of the reference face, deepest vertex, moved
    half its depth back; edge x edge: midpoint of the closest points of the two lines from the 2 x 2 normal equations --
    where the product header follows ODE's 2-D route. PARITY UNPINNED against FCL beyond the reference's TangentBox
    scenes (penetration_as_point_pair_callback_test.cc:914-920)."""
    R1, T1, R2, T2 = _R(X_WA), _p(X_WA), _R(X_WB), _p(X_WB)
    A, B = np.asarray(size_A, float) * 0.5, np.asarray(size_B, float) * 0.5
    p = T2 - T1
    pp = R1.T @ p
    R = R1.T @ R2
    Q = np.abs(R)
    s, code, normal, invert = -np.finfo(float).max, 0, None, False
    for i in range(3):
        tmp = pp[i]
        s2 = abs(tmp) - (((Q[i, 0] * B[0] + Q[i, 1] * B[1]) + Q[i, 2] * B[2]) + A[i])
        if s2 > 0:
            return None
        if s2 > s:
            s, code, normal, invert = s2, 1 + i, R1[:, i], tmp < 0
    for i in range(3):
        tmp = float(R2[:, i] @ p)
        s2 = abs(tmp) - (((Q[0, i] * A[0] + Q[1, i] * A[1]) + Q[2, i] * A[2]) + B[i])
        if s2 > 0:
            return None
        if s2 > s:
            s, code, normal, invert = s2, 4 + i, R2[:, i], tmp < 0
    Qf = Q + 1.0e-6
    eps = np.finfo(float).eps
    for i in range(3):
        i1, i2 = (i + 1) % 3, (i + 2) % 3
        for j in range(3):
            j1, j2 = (j + 1) % 3, (j + 2) % 3
            tmp = pp[i2] * R[i1, j] - pp[i1] * R[i2, j]
            s2 = abs(tmp) - (((A[i1] * Qf[i2, j] + A[i2] * Qf[i1, j]) + B[j1] * Qf[i, j2]) + B[j2] * Qf[i, j1])
            if s2 > 0:
                return None
            n = np.zeros(3)
            n[i1], n[i2] = -R[i2, j], R[i1, j]
            l = math.sqrt(n[i1] * n[i1] + n[i2] * n[i2])
            if l > eps:
                s2 /= l
                if s2 * 1.05 > s:
                    s, code, normal, invert = s2, 7 + 3 * i + j, R1 @ (n / l), tmp < 0
    if code == 0:
        return None
    normal = -normal if invert else normal.copy()
    depth = -s
    if code > 6:
        pa = T1 + sum(R1[:, j] * (A[j] * (1.0 if R1[:, j] @ normal > 0 else -1.0)) for j in range(3))
        pb = T2 + sum(R2[:, j] * (B[j] * (-1.0 if R2[:, j] @ normal > 0 else 1.0)) for j in range(3))
        ua, ub = R1[:, (code - 7) // 3], R2[:, (code - 7) % 3]
        c = float(ua @ ub)
        if 1 - c * c <= 0.0001:
            alpha = beta = 0.0
        else:   # minimise |pa + alpha ua - pb - beta ub|^2
            alpha, beta = np.linalg.solve(np.array([[1.0, -c], [-c, 1.0]]), np.array([ua @ (pb - pa), -(ub @ (pb - pa))]))
        return normal, 0.5 * ((pa + alpha * ua) + (pb + beta * ub)), depth
    first = code <= 3
    (Ra, Ta, Sa), (Rb, Tb, Sb) = ((R1, T1, A), (R2, T2, B)) if first else ((R2, T2, B), (R1, T1, A))
    normal2 = normal if first else -normal
    nr = Rb.T @ normal2
    anr = np.abs(nr)
    if anr[1] > anr[0]:
        lanr, a1, a2 = (1, 0, 2) if anr[1] > anr[2] else (2, 0, 1)
    else:
        lanr, a1, a2 = (0, 1, 2) if anr[0] > anr[2] else (2, 0, 1)
    centre = Tb + Rb[:, lanr] * (Sb[lanr] if nr[lanr] < 0 else -Sb[lanr])
    e1, e2 = Rb[:, a1] * Sb[a1], Rb[:, a2] * Sb[a2]
    poly = [centre - e1 - e2, centre - e1 + e2, centre + e1 + e2, centre + e1 - e2]
    codeN = code - 1 if first else code - 4
    code1, code2 = (1 if codeN == 0 else 0), (1 if codeN == 2 else 2)
    for axis, h in ((Ra[:, code1], Sa[code1]), (Ra[:, code2], Sa[code2])):
        for sign in (-1.0, 1.0):
            out = []
            for k, P in enumerate(poly):
                Pn = poly[(k + 1) % len(poly)]
                dP, dn = sign * float(axis @ (P - Ta)), sign * float(axis @ (Pn - Ta))
                if dP < h:
                    out.append(P)
                if (dP < h) != (dn < h):
                    out.append(P + (Pn - P) * ((h - dP) / (dn - dP)))
            poly = out
    best = None
    for P in poly:
        dep = Sa[codeN] - float(normal2 @ (P - Ta))
        if dep >= 0 and (best is None or dep > best[1]):
            best = (P, dep)
    if best is None:
        return None
    return normal, best[0] + normal2 * (0.5 * best[1]), best[1]


def box_box_point_pair(size_A, X_WA, size_B, X_WB):
    """CalcDistanceFallback<double> (:164-207) on FCL's box-box contact."""
    hit = box_box_contact(size_A, X_WA, size_B, X_WB)
    if hit is None or hit[2] <= np.finfo(float).eps:
        return None
    normal, pos, depth = hit
    nhat_BA_W = -normal
    return {"depth": depth, "nhat_BA_W": nhat_BA_W, "p_WCa": pos - 0.5 * depth * nhat_BA_W, "p_WCb": pos + 0.5 * depth * nhat_BA_W}


# ---- Minkowski Portal Refinement (libccd's ccdMPRPenetration behind fcl::collide for the remaining convex pairs) -------

def support(shape, X_WS, dir_W):
    """Farthest point of the shape along a World direction, in World (FCL's per-shape support functions)."""
    R, t = _R(X_WS), _p(X_WS)
    d = R.T @ np.asarray(dir_W, float)
    eps = np.finfo(float).eps
    sgn = lambda x: 0.0 if abs(x) < eps else (-1.0 if x < 0 else 1.0)
    kind = shape[0]
    if kind == "sphere":
        v = d * (shape[1] / math.sqrt(float(d @ d)))
    elif kind == "box":
        v = np.array([sgn(d[k]) * shape[1][k] * 0.5 for k in range(3)])
    elif kind == "cylinder":
        zdist = math.sqrt(d[0] * d[0] + d[1] * d[1])
        if abs(zdist) < eps:
            v = np.array([0.0, 0.0, sgn(d[2]) * shape[2] * 0.5])
        else:
            v = np.array([shape[1] / zdist * d[0], shape[1] / zdist * d[1], sgn(d[2]) * shape[2] * 0.5])
    elif kind == "capsule":
        v = d * (shape[1] / math.sqrt(float(d @ d)))
        v[2] += shape[2] * 0.5 if d[2] > 0 else -shape[2] * 0.5
    elif kind == "ellipsoid":
        abc2 = np.asarray(shape[1], float) ** 2 * d
        v = abc2 / math.sqrt(float(d @ abc2))
    elif kind == "vertices":
        V = np.asarray(shape[1], float).reshape(-1, 3)
        v = V[int(np.argmax(V @ d))]
    else:
        raise NotImplementedError(kind)
    return R @ v + t


def _center(shape, X_WS):
    if shape[0] == "vertices":
        return _R(X_WS) @ np.mean(np.asarray(shape[1], float).reshape(-1, 3), axis=0) + _p(X_WS)
    return _p(X_WS)


def mpr_contact(shape_A, X_WA, shape_B, X_WB, tolerance=1e-7, max_iterations=500):
    """G. Snethen's Minkowski Portal Refinement as libccd's ccdMPRPenetration runs it (src/mpr.c; FCL reaches it through
    GJKCollide for every convex pair without a primitive algorithm): None or (normal from A into B, position, depth).
    The product header's default tolerance (1e-7, see there) is used; libccd's path through the refinement is not
    reproduced bit for bit, so the reference anchors are its BoxPenetrationTest scenes at the tolerances they state."""
    eps = np.finfo(float).eps
    zero = lambda x: abs(x) < eps

    def sup(d):
        v1, v2 = support(shape_A, X_WA, d), support(shape_B, X_WB, -d)
        return [v1 - v2, v1, v2]

    def unit(v):
        n = math.sqrt(float(v @ v))
        return v / n if n else v
    c1, c2 = _center(shape_A, X_WA), _center(shape_B, X_WB)
    P = [[c1 - c2, c1, c2], None, None, None]
    if all(zero(x) for x in P[0][0]):
        P[0][0] = np.array([eps * 10, 0.0, 0.0])
    d = unit(-P[0][0])
    P[1] = sup(d)
    dot = float(P[1][0] @ d)
    if zero(dot) or dot < 0:
        return None
    d = np.cross(P[0][0], P[1][0])
    mid = lambda s: 0.5 * (s[1] + s[2])
    if zero(math.sqrt(float(d @ d))):
        if all(zero(x) for x in P[1][0]):
            return np.zeros(3), mid(P[1]), 0.0
        n = P[1][0]
        return unit(n), mid(P[1]), math.sqrt(float(n @ n))
    d = unit(d)
    P[2] = sup(d)
    dot = float(P[2][0] @ d)
    if zero(dot) or dot < 0:
        return None
    d = unit(np.cross(P[1][0] - P[0][0], P[2][0] - P[0][0]))
    if float(d @ P[0][0]) > 0:
        P[1], P[2] = P[2], P[1]
        d = -d
    for guard in range(max_iterations + 2):
        if guard > max_iterations:
            return None
        P[3] = sup(d)
        dot = float(P[3][0] @ d)
        if zero(dot) or dot < 0:
            return None
        if float(np.cross(P[1][0], P[3][0]) @ P[0][0]) < -eps:
            P[2] = P[3]
        elif float(np.cross(P[3][0], P[2][0]) @ P[0][0]) < -eps:
            P[1] = P[3]
        else:
            break
        d = unit(np.cross(P[1][0] - P[0][0], P[2][0] - P[0][0]))
    portal_dir = lambda: unit(np.cross(P[2][0] - P[1][0], P[3][0] - P[1][0]))

    def expand(v4):
        v4v0 = np.cross(v4[0], P[0][0])
        if float(P[1][0] @ v4v0) > 0:
            if float(P[2][0] @ v4v0) > 0:
                P[1] = v4
            else:
                P[3] = v4
        elif float(P[3][0] @ v4v0) > 0:
            P[2] = v4
        else:
            P[1] = v4

    def reach(v4, d):
        d4 = float(v4[0] @ d)
        return min(d4 - float(P[1][0] @ d), d4 - float(P[2][0] @ d), d4 - float(P[3][0] @ d))
    guard = 0
    while True:
        d = portal_dir()
        dot = float(P[1][0] @ d)
        if zero(dot) or dot > 0:
            break
        v4 = sup(d)
        d4 = float(v4[0] @ d)
        if not (zero(d4) or d4 > 0):
            return None
        r = reach(v4, d)
        if r < tolerance or zero(r - tolerance) or guard > max_iterations:
            return None
        expand(v4)
        guard += 1
    it = 0
    while True:
        d = portal_dir()
        v4 = sup(d)
        r = reach(v4, d)
        if r < tolerance or zero(r - tolerance) or it > max_iterations:
            a, b, c = P[1][0], P[2][0], P[3][0]
            d1, d2 = b - a, c - a
            v, w, p, q, rr = float(d1 @ d1), float(d2 @ d2), float(a @ d1), float(a @ d2), float(d1 @ d2)
            div = w * v - rr * rr
            wit = None
            if not zero(div):
                s = (q * rr - w * p) / div
                t = (-s * rr - q) / w
                ok = lambda x: (zero(x) or x > 0) and (x < 1 or zero(x - 1))
                if ok(s) and ok(t) and (t + s < 1 or zero(t + s - 1)):
                    wit = a + s * d1 + t * d2
            if wit is None:
                best = None
                for x0, x1 in ((a, b), (a, c), (b, c)):
                    e = x1 - x0
                    k = min(1.0, max(0.0, -float(x0 @ e) / float(e @ e)))
                    cand = x0 + k * e
                    if best is None or float(cand @ cand) < float(best @ best):
                        best = cand
                wit = best
            depth = math.sqrt(float(wit @ wit))
            pdir = d if all(zero(x) for x in wit) else wit
            pdir = unit(pdir)
            bc = [float(np.cross(P[1][0], P[2][0]) @ P[3][0]), float(np.cross(P[3][0], P[2][0]) @ P[0][0]),
                  float(np.cross(P[0][0], P[1][0]) @ P[3][0]), float(np.cross(P[2][0], P[1][0]) @ P[0][0])]
            total = ((bc[0] + bc[1]) + bc[2]) + bc[3]
            if zero(total) or total < 0:
                bc = [0.0, float(np.cross(P[2][0], P[3][0]) @ d), float(np.cross(P[3][0], P[1][0]) @ d),
                      float(np.cross(P[1][0], P[2][0]) @ d)]
                total = (bc[1] + bc[2]) + bc[3]
            p1 = sum(bc[k] * P[k][1] for k in range(4))
            p2 = sum(bc[k] * P[k][2] for k in range(4))
            return pdir, 0.5 * (p1 + p2) / total, depth
        expand(v4)
        it += 1


def mpr_point_pair(shape_A, X_WA, shape_B, X_WB):
    """CalcDistanceFallback<double> (:164-207) on the MPR contact."""
    hit = mpr_contact(shape_A, X_WA, shape_B, X_WB)
    if hit is None or hit[2] <= np.finfo(float).eps:
        return None
    normal, pos, depth = hit
    nhat_BA_W = -normal
    return {"depth": depth, "nhat_BA_W": nhat_BA_W, "p_WCa": pos - 0.5 * depth * nhat_BA_W, "p_WCb": pos + 0.5 * depth * nhat_BA_W}


def point_pair(shape_A, X_WA, shape_B, X_WB):
    """ComputeNarrowPhasePenetration for shapes given as ('sphere', r) | ('box', (x, y, z)) | ('half_space',) |
    ('cylinder', r, L) | ('capsule', r, L) | ('ellipsoid', (a, b, c)) | ('vertices', V); A is the
    geometry with the smaller id. Returns None (no penetration) or dict(p_WCa, p_WCb, nhat_BA_W, depth). Raises for pairs
    without a sphere or a half space (the reference calls FCL's GJK / EPA or box-box there) and for two half spaces (the
    reference throws, :329-338)."""
    a_is_sphere, b_is_sphere = shape_A[0] == "sphere", shape_B[0] == "sphere"
    if not (a_is_sphere or b_is_sphere):
        if (shape_A[0] == "half_space") != (shape_B[0] == "half_space"):
            return half_space_point_pair(shape_A, X_WA, shape_B, X_WB)
        if shape_A[0] == "half_space":
            raise NotImplementedError("two half spaces: the reference throws (:329-338)")
        if shape_A[0] == "box" and shape_B[0] == "box":
            return box_box_point_pair(shape_A[1], X_WA, shape_B[1], X_WB)
        return mpr_point_pair(shape_A, X_WA, shape_B, X_WB)
    if (shape_B if a_is_sphere else shape_A)[0] in ("ellipsoid", "vertices"):   # the reference sends these to FCL (:273-278)
        return mpr_point_pair(shape_A, X_WA, shape_B, X_WB)
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
