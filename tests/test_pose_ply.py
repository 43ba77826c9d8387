"""Host-side rows either side of the path: per-timestamp actor poses with the reference's
interpolation rules (street_gaussians_ns/data/utils/dynamic_annotation.py) and the Inria-layout PLY the exporter
writes (street_gaussians_ns/scripts/exporter.py:59-137).  CPU only."""
import numpy as np
import pytest
import torch

import street_gaussians_ns_b200.synthetic as syn
from street_gaussians_ns_b200 import ply_io, pose_table as pt
from street_gaussians_ns_b200.scene import quaternion_from_matrix


def yaw_quat(a):  # rotation about y, (w, x, y, z)
    return [np.cos(a / 2), 0.0, np.sin(a / 2), 0.0]


def frames():
    def obj(gid, x, yaw, moving=True, typ="car"):
        return {"type": typ, "is_moving": moving, "gid": gid, "translation": [x, 0.0, -10.0], "rotation": yaw_quat(yaw),
                "size": [4.0, 2.0, 1.5]}
    return [
        {"timestamp": 1000.0, "objects": [obj("a", 0.0, 0.0), obj("b", 5.0, 0.1), obj("s", 9.0, 0.0, moving=False)]},
        {"timestamp": 1002.0, "objects": [obj("a", 2.0, 0.4), obj("p", 1.0, 0.0, typ="pedestrian"), obj("c", 7.0, 0.0, typ="PoliceCar")]},
        {"timestamp": 1001.0, "objects": [obj("a", 1.0, 0.2), obj("b", 5.5, 0.1)]},
    ]


def test_parse_timestamp_and_quaternion_helpers():
    assert pt.parse_timestamp(1000.0) == "1000000000000000" and len(pt.parse_timestamp("1700000000.25")) == 16
    for a in (0.0, 0.3, -1.2, 3.0):
        R = pt.quaternion_matrix(yaw_quat(a))[:3, :3]
        assert np.allclose(R @ R.T, np.eye(3), atol=1e-12) and np.isclose(np.linalg.det(R), 1.0)
        assert np.allclose(R, [[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]], atol=1e-12)
        q = quaternion_from_matrix(R)
        assert np.allclose(np.abs(np.dot(q, yaw_quat(a))), 1.0, atol=1e-9)  # same rotation up to sign
    q0, q1 = np.array(yaw_quat(0.2)), np.array(yaw_quat(1.0))
    assert np.allclose(pt.quaternion_slerp(q0, q1, 0.0), q0) and np.allclose(pt.quaternion_slerp(q0, q1, 1.0), q1)
    assert np.allclose(pt.quaternion_slerp(q0, q1, 0.25), yaw_quat(0.4), atol=1e-12)      # constant angular velocity
    assert np.allclose(pt.quaternion_slerp(q0, -q1, 0.25), yaw_quat(0.4), atol=1e-12)     # shortest path
    assert np.allclose(pt.quaternion_slerp(q0, q0, 0.5), q0)


def test_pose_table_filters_lookup_and_interpolation():
    tab = pt.PoseTable(frames(), has_points=lambda gid: gid != "nopoints")
    assert tab.all_names == [pt.parse_timestamp(t) for t in (1000.0, 1001.0, 1002.0)]      # sorted by timestamp
    at0 = tab[1000.0]
    assert [b.track_id for b in at0] == ["a", "b"]                                          # static box dropped
    assert [b.track_id for b in tab[1002.0]] == ["a", "c"]                                  # label filter: car / *Car
    assert tab.objects_frames == {"a": [0, 1, 2], "b": [0, 1], "c": [2]}
    assert np.allclose(at0[0].size, pt.EXP_RATE * np.array([4.0, 2.0, 1.5])) and at0[0].frame == 0
    # between two annotated timestamps: only tracks present in both, centre lerp + rotation slerp
    mid = tab[1001.5]
    assert [b.track_id for b in mid] == ["a"] and mid[0].frame == -1
    assert np.allclose(mid[0].center, [1.5, 0.0, -10.0])
    assert np.allclose(mid[0].rot, pt.quaternion_matrix(yaw_quat(0.3))[:3, :3], atol=1e-12)
    assert tab[999.0] == [] and tab[1003.0] == []                                           # out of range
    assert [b.track_id for b in tab[0.5]] == [b.track_id for b in tab[1002.0]]             # fraction of the sequence: round(1.5) = 2
    poses = tab.poses_at(1001.0)
    assert [p.track_id for p in poses] == ["a", "b"] and poses[0].frame == 1 and list(poses[0].frame_list) == [0, 1, 2]
    # dataparser transform + scale are applied to centre / rotation / size (dynamic_annotation.py:194-204, 330-333)
    T = np.eye(4); T[:3, :3] = pt.quaternion_matrix(yaw_quat(np.pi / 2))[:3, :3]; T[:3, 3] = [1.0, 2.0, 3.0]
    tab2 = pt.PoseTable(frames(), transform_matrix=T, scale_factor=0.5, self_car_label="b")
    b = tab2[1000.0]
    assert [x.track_id for x in b] == ["a"]
    assert np.allclose(b[0].center, 0.5 * (T[:3, :3] @ np.array([0.0, 0.0, -10.0]) + T[:3, 3]))
    assert np.allclose(b[0].rot, T[:3, :3] @ pt.quaternion_matrix(yaw_quat(0.0))[:3, :3])
    assert np.allclose(b[0].size, 0.5 * pt.EXP_RATE * np.array([4.0, 2.0, 1.5]))


def test_ply_layout_and_round_trip(tmp_path):
    fr = syn.make_frame(n_background=300, n_actors=1, n_per_actor=257, width=64, height=48, seed=5)
    ps = fr.segments[1].params      # an actor: features_dc [n, 5, 3]
    with torch.no_grad():
        ps.means[3, 1] = float("nan")  # non-finite rows are not exported (exporter.py:103-116)
        ps.scales[10, 0] = float("inf")
    path = tmp_path / "point_cloud_object_x.ply"
    n = ply_io.write_ply(path, ps)
    assert n == 255
    raw = path.read_bytes()
    head, body = raw.split(b"end_header\n", 1)
    lines = head.decode().split("\n")
    assert lines[0] == "ply" and lines[1] == "format binary_little_endian 1.0" and lines[2] == "element vertex 255"
    props = [l.split()[2] for l in lines if l.startswith("property")]
    assert all(l.split()[1] == "float" for l in lines if l.startswith("property"))
    assert props == (["x", "y", "z", "nx", "ny", "nz", "f_dc_0", "f_dc_1", "f_dc_2"] + [f"f_rest_{i}" for i in range(45)]
                     + ["opacity", "scale_0", "scale_1", "scale_2", "rot_0", "rot_1", "rot_2", "rot_3"])
    assert len(body) == 255 * len(props) * 4
    cols = ply_io.read_ply_columns(path)
    keep = np.ones(257, bool); keep[[3, 10]] = False
    assert np.array_equal(cols["nx"], np.zeros(255, np.float32))
    assert np.array_equal(cols["f_dc_1"], ps.features_dc[:, 0, 1].numpy()[keep])            # first Fourier coefficient only
    # Inria order: channel-major rest coefficients, f_rest_{c*15 + k} = features_rest[:, k, c]
    assert np.array_equal(cols["f_rest_16"], ps.features_rest[:, 1, 1].numpy()[keep])
    back = ply_io.read_ply(path)
    assert back.features_dc.shape == (255, 1, 3) and back.features_rest.shape == (255, 15, 3)
    for name in ("means", "scales", "quats", "features_rest", "opacities"):
        assert torch.equal(getattr(back, name), getattr(ps, name)[torch.from_numpy(keep)]), name
    assert torch.equal(back.features_dc[:, 0], ps.features_dc[torch.from_numpy(keep), 0])
    with pytest.raises(ValueError):
        bad = tmp_path / "bad.ply"
        bad.write_bytes(b"ply\nformat ascii 1.0\nelement vertex 0\nend_header\n")
        ply_io.read_ply_columns(bad)
