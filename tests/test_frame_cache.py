"""Host logic of SceneGraphRasterModel._frame (CPU): a timestamp's segments are kept per timestamp and validated by the CONTENT
of the boxes -- not by object identity, which says nothing when a box is mutated in place or when CPython recycles ids."""
import numpy as np
import torch

import street_gaussians_ns_b200.synthetic as syn
from street_gaussians_ns_b200.model import ActorPose, SceneGraphConfig, SceneGraphRasterModel


def _model(poses_at):
    bg = syn.make_background(64, seed=0)
    actors = {str(a): syn.make_actor(16, seed=1 + a) for a in range(3)}
    return SceneGraphRasterModel(bg, actors, SceneGraphConfig(use_sky_sphere=False), poses_at=poses_at)


def test_same_content_reuses_segments_and_new_content_rebuilds():
    rot, c = np.eye(3), np.array([1.0, -1.2, -8.0])
    persistent = [ActorPose("0", rot.copy(), c.copy(), 3, list(range(10))), ActorPose("2", rot.copy(), c + 2.0, 3, list(range(10)))]
    m = _model(lambda t: persistent)
    cam = syn.make_camera(64, 48, time=3.0)
    f1 = m._frame(cam)
    f2 = m._frame(syn.make_camera(64, 48, time=3.0))  # another Camera object at the same timestamp
    assert f2.segments is f1.segments and m.visible_model_names == ["background", "object_0", "object_2"]
    persistent[1].center[2] -= 1.5  # mutated in place: same ids
    f3 = m._frame(cam)
    assert f3.segments is not f1.segments
    np.testing.assert_array_equal(f3.segments[2].center, persistent[1].center)
    # fresh objects with equal content (ids may be recycled or not: irrelevant) -> reuse
    m2 = _model(lambda t: [ActorPose("1", rot.copy(), c.copy(), 3, list(range(10)))])
    g1, g2 = m2._frame(cam), m2._frame(cam)
    assert g2.segments is g1.segments
    # the Fourier time of a box depends on its frame and frame list: part of the content
    m3_poses = [ActorPose("1", rot.copy(), c.copy(), 3, list(range(10)))]
    m3 = _model(lambda t: m3_poses)
    h1 = m3._frame(cam)
    m3_poses[0].frame = 7
    h2 = m3._frame(cam)
    assert h2.segments is not h1.segments and not np.array_equal(h1.segments[1].idft, h2.segments[1].idft)


def test_replaced_parameters_invalidate():
    rot, c = np.eye(3), np.array([1.0, -1.2, -8.0])
    m = _model(lambda t: [ActorPose("0", rot, c, 0, [0])])
    cam = syn.make_camera(64, 48, time=0.0)
    f1 = m._frame(cam)
    g = m.all_models["background"].gauss_params
    g["means"] = torch.nn.Parameter(torch.cat([g["means"].data, g["means"].data[:4]]))  # what a refinement does
    m.invalidate_frames()
    f2 = m._frame(cam)
    assert f2.segments is not f1.segments and f2.segments[0].params.means.shape[0] == 68
