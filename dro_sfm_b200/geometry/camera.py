"""Drop-in for dro_sfm.geometry.camera.Camera (reference: dro_sfm/geometry/camera.py:12-194)."""
import torch
import torch.nn as nn

from .. import ops
from .pose import Pose
from .camera_utils import scale_intrinsics


class Camera(nn.Module):
    """Differentiable pinhole camera: reconstruct (lift) and project, backed by the CUDA kernels."""

    def __init__(self, K, Tcw=None):
        super().__init__()
        self.K = K
        self.Tcw = Pose.identity(len(K), device=K.device) if Tcw is None else Tcw
        self._Twc = None
        self._Kinv = None

    def __len__(self):
        return len(self.K)

    def to(self, *args, **kwargs):
        self.K = self.K.to(*args, **kwargs)
        self.Tcw = self.Tcw.to(*args, **kwargs)
        self._Twc = self._Kinv = None
        return self

    @property
    def fx(self):
        return self.K[:, 0, 0]

    @property
    def fy(self):
        return self.K[:, 1, 1]

    @property
    def cx(self):
        return self.K[:, 0, 2]

    @property
    def cy(self):
        return self.K[:, 1, 2]

    @property
    def Twc(self):
        """World -> camera (inverse of Tcw), cached like the reference's lru_cache (camera.py:64-68)."""
        if self._Twc is None:
            self._Twc = self.Tcw.inverse()
        return self._Twc

    @property
    def Kinv(self):
        """Closed-form inverse intrinsics (camera.py:70-79)."""
        if self._Kinv is None:
            Kinv = self.K.clone()
            Kinv[:, 0, 0] = 1. / self.fx
            Kinv[:, 1, 1] = 1. / self.fy
            Kinv[:, 0, 2] = -1. * self.cx / self.fx
            Kinv[:, 1, 2] = -1. * self.cy / self.fy
            self._Kinv = Kinv
        return self._Kinv

    def scaled(self, x_scale, y_scale=None):
        """Camera with rescaled intrinsics; returns self when no scaling is needed (camera.py:83-107)."""
        if y_scale is None:
            y_scale = x_scale
        if x_scale == 1. and y_scale == 1.:
            return self
        K = scale_intrinsics(self.K.clone(), x_scale, y_scale)
        return Camera(K, Tcw=self.Tcw)

    def reconstruct(self, depth, frame='w'):
        """depth [B,1,H,W] -> points [B,3,H,W] in the camera ('c') or world ('w') frame (camera.py:111-147)."""
        B, C, H, W = depth.shape
        assert C == 1
        if frame == 'c':
            return ops.reconstruct(depth, self.K, None)
        if frame == 'w':
            return ops.reconstruct(depth, self.K, self.Twc.mat)
        raise ValueError('Unknown reference frame {}'.format(frame))

    def project(self, X, frame='w', normalize=True):
        """points [B,3,H,W] -> coordinates [B,H,W,2] (camera.py:149-194)."""
        B, C, H, W = X.shape
        assert C == 3
        if frame == 'c':
            return ops.project(X, self.K, None, normalize)
        if frame == 'w':
            return ops.project(X, self.K, self.Tcw.mat, normalize)
        raise ValueError('Unknown reference frame {}'.format(frame))
