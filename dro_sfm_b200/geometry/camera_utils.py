"""Drop-in for dro_sfm.geometry.camera_utils (reference: dro_sfm/geometry/camera_utils.py:7-56)."""
import torch

from .. import ops


def construct_K(fx, fy, cx, cy, dtype=torch.float, device=None):
    """[3,3] pinhole intrinsics (camera_utils.py:7-11)."""
    return torch.tensor([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], dtype=dtype, device=device)


def scale_intrinsics(K, x_scale, y_scale):
    """In-place rescale of intrinsics (camera_utils.py:13-19)."""
    K[..., 0, 0] *= x_scale
    K[..., 1, 1] *= y_scale
    K[..., 0, 2] = (K[..., 0, 2] + 0.5) * x_scale - 0.5
    K[..., 1, 2] = (K[..., 1, 2] + 0.5) * y_scale - 0.5
    return K


def view_synthesis(ref_image, depth, ref_cam, cam, mode='bilinear', padding_mode='zeros'):
    """Synthesize ref_image in the frame of `cam` (camera_utils.py:23-56).

    One fused kernel (reconstruct -> project -> bilinear gather) when the target camera sits at the
    identity, which is how every caller in the reference builds it; otherwise the three operators run
    back to back."""
    assert depth.size(1) == 1
    if mode != 'bilinear':
        raise NotImplementedError("dro_sfm_b200: interpolation mode {!r} is not supported".format(mode))
    # `_is_identity` is set by Pose.identity() and cleared by the mat setter; an in-place edit of the matrix storage
    # (pose.mat[:, :3, 3] = t) bumps the tensor's version counter, which is checked here
    if cam.Tcw._is_identity and cam.Tcw.mat._version == 0:
        return ops.view_synthesis(ref_image, depth, ref_cam.Tcw.mat, cam.K, ref_cam.K, 1.0, padding_mode)
    world_points = cam.reconstruct(depth, frame='w')
    ref_coords = ref_cam.project(world_points, frame='w')
    return ops.grid_gather(ref_image, ref_coords, padding_mode)
