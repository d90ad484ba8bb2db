"""Known-answer tests pinning the oracle's traction integration to the reference's own:
multibody/plant/test/hydroelastic_traction_calculator_test.cc. CPU only."""
from __future__ import annotations

import math

import numpy as np
import pytest

from oracle import ho

EPS = np.finfo(float).eps
TOL = 10 * EPS          # the reference fixture's tol_ (:373)
VS = 1e-6               # default vslip_regularizer (hydroelastic_traction_calculator.h:91)

POSES = [ho.IDENTITY,   # frame Y in World (the reference instantiates identity, translations and rotations)
         ho.xform(None, [1.0, 2.0, 3.0]),
         ho.xform(ho.rpy_matrix(math.pi / 2, 0, 0), [0, 0, 0]),
         ho.xform(ho.rpy_matrix(0.3, -0.7, 1.1), [-0.5, 0.25, 2.0])]


def make_surface(rep, X_WH):
    """CreateContactSurface (hydroelastic_traction_calculator_test.cc:62-187): the bottom face of a unit box at
    z = -0.5 in the half-space frame H, e = -z, two triangles or one quad, normals along -Hz; M = half space."""
    R, p = ho.xform_parts(X_WH)
    v_H = np.array([[0.5, 0.5, -0.5], [-0.5, 0.5, -0.5], [-0.5, -0.5, -0.5], [0.5, -0.5, -0.5]])
    e = -v_H[:, 2]
    v_W = v_H @ R.T + p
    if rep == ho.KTRIANGLE:
        return ho.surface_from_arrays(rep, v_W, e, [[0, 2, 1], [2, 0, 3]])
    return ho.surface_from_arrays(rep, v_W, e, [4, 3, 2, 1, 0], grads=[R @ np.array([0, 0, -1.0])])


def tractions_at_body_origins(surface, X_WY, v_box_Y, dissipation, mu):
    """ComputeSpatialTractionsAtBodyOriginsFromHydroelasticModel (:243-281) needs the traction at the centroid of
    face 0 only: a one-face copy of the surface with area 1 gives exactly that through the oracle's integrator."""
    R, p = ho.xform_parts(X_WY)
    V_WB = np.concatenate([np.zeros(3), R @ np.asarray(v_box_Y, float)])
    return ho.traction(surface, p, p, np.zeros(6), V_WB, dissipation, mu, VS)


@pytest.mark.parametrize("rep", [ho.KTRIANGLE, ho.KPOLYGON])
@pytest.mark.parametrize("X_WY", POSES)
def test_forces_over_the_patch(rep, X_WY):
    # VanillaTractionOverPatch / FrictionalTractionOverPatch / TractionOverPatchWithDissipation (:536-650).
    s = make_surface(rep, X_WY)
    R, _ = ho.xform_parts(X_WY)
    in_Y = lambda F: (R.T @ F[3:6], R.T @ F[0:3])  # (f, tau) of a spatial force, re-expressed in Y
    tol = 40 * EPS  # the sums here run over rotated coordinates of magnitude ~3
    _, F_Ao, F_Bo = tractions_at_body_origins(s, X_WY, [0, 0, 0], 0.0, 0.0)
    f, tau = in_Y(F_Bo)
    np.testing.assert_allclose(f, [0, 0, 0.5], atol=tol)
    np.testing.assert_allclose(tau, [0, 0, 0], atol=tol)
    np.testing.assert_allclose(F_Ao[3:6] + F_Bo[3:6], 0, atol=tol)
    _, _, F_Bo = tractions_at_body_origins(s, X_WY, [1, 0, 0], 0.0, 1.0)
    f, tau = in_Y(F_Bo)
    assert abs(f[0] + 0.5) <= VS and abs(f[1]) <= tol and abs(f[2] - 0.5) <= tol
    assert abs(tau[0]) <= tol and abs(tau[1] - 0.25) <= VS and abs(tau[2]) <= tol
    _, _, F_Bo = tractions_at_body_origins(s, X_WY, [0, 0, -1.0], 1.0, 0.0)
    f, tau = in_Y(F_Bo)
    np.testing.assert_allclose(f, [0, 0, 0.5 + 0.5 * 1.0 * 1.0], atol=tol)
    np.testing.assert_allclose(tau, [0, 0, 0], atol=tol)


def test_traction_at_a_point_polygon():
    # VanillaTraction / TractionWithFriction / TractionWithDissipation (:382-533) for the polygon representation,
    # where the single face's centroid is the quadrature point: per-area tractions equal the patch forces (area 1)
    # and the moment about Bo is p_BoCo x f with p_BoCo = (0, 0, -0.5).
    for X_WY in POSES:
        s = make_surface(ho.KPOLYGON, X_WY)
        R, _ = ho.xform_parts(X_WY)
        for v_box, d, mu, f_expected, slack in [([0, 0, 0], 0.0, 0.0, [0, 0, 0.5], 0.0),
                                                ([1, 0, 0], 0.0, 1.0, [-0.5, 0, 0.5], VS),
                                                ([0, 0, -1], 1.0, 0.0, [0, 0, 1.0], 0.0)]:
            _, _, F_Bo = tractions_at_body_origins(s, X_WY, v_box, d, mu)
            f, tau = R.T @ F_Bo[3:6], R.T @ F_Bo[0:3]
            np.testing.assert_allclose(f, f_expected, atol=40 * EPS + slack)
            np.testing.assert_allclose(tau, np.cross([0, 0, -0.5], f_expected), atol=40 * EPS + slack)


def test_calc_atan_x_over_x_from_x_squared():
    # hydroelastic_traction_calculator_test.cc:676-690 (f_ref)
    xs = [0.0, 1e-6, 1e-4, 0.119999, 0.12, 0.120001, 1.0]
    ref = [1.0, 9.9999999999966671e-01, 9.9999999666666672e-01, 9.9524112879111848e-01, 9.9524105015282038e-01,
           9.9524097151388935e-01, 7.8539816339744828e-01]
    for x, f in zip(xs, ref):
        assert abs(ho.atan_x_over_x_from_x_squared(x * x) - f) <= 4 * EPS


def test_combination_rules():
    # multibody/hydroelastics/test/hydroelastic_engine_test.cc:124-178 and coulomb_friction.h:122-139.
    inf = float("inf")
    assert ho.combined_dissipation(10.0, inf, 1.0, 0.0) == 1.0      # CombineSoftAndRigidMaterialProperties
    assert ho.combined_dissipation(inf, 10.0, 0.0, 1.0) == 1.0
    assert ho.combined_dissipation(inf, inf, 0.0, 0.0) == 0.0       # CombineRigidAndRigidMaterialProperties
    assert ho.combined_dissipation(inf, inf, 1.0, 3.0) == 2.0
    for args in ((2.0, 8.0, 1.0, 4.0), (8.0, 2.0, 4.0, 1.0)):       # CombineSoftAndSoftMaterialProperties: d* = 1.6
        assert abs(ho.combined_dissipation(*args) - 1.6) <= np.finfo(float).eps
    assert ho.combined_friction(0.0, 0.0) == 0.0
    assert ho.combined_friction(0.5, 0.5) == 0.5
    assert ho.combined_friction(1.0, 0.25) == pytest.approx(0.4, rel=1e-15)
