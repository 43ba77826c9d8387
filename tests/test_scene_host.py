"""Host helpers of the per-frame segment table (CPU)."""
import numpy as np

from street_gaussians_ns_b200 import scene


def _rot(yaw, pitch=0.0, roll=0.0):
    cy, sy, cp, sp, cr, sr = np.cos(yaw), np.sin(yaw), np.cos(pitch), np.sin(pitch), np.cos(roll), np.sin(roll)
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rx = np.array([[1, 0, 0], [0, cp, -sp], [0, sp, cp]])
    Rz = np.array([[cr, -sr, 0], [sr, cr, 0], [0, 0, 1]])
    return Ry @ Rx @ Rz


def test_batched_quaternions_equal_the_one_by_one_form_bit_for_bit():
    rng = np.random.RandomState(0)
    mats = np.stack([_rot(*rng.uniform(-3.1, 3.1, 3)) for _ in range(200)] + [np.eye(3), _rot(np.pi), _rot(0.0, np.pi)])
    mats[5] *= 1.0 + 1e-7  # slightly non-orthonormal, as float64 box rotations after a dataparser transform are
    batched = scene.quaternions_from_matrices(mats)
    single = np.stack([scene.quaternion_from_matrix(m) for m in mats])
    np.testing.assert_array_equal(batched, single)
    assert np.all(batched[:, 0] >= 0.0)
    np.testing.assert_allclose(np.linalg.norm(batched, axis=1), 1.0, atol=1e-12)
