"""Drop-in for dro_sfm.geometry.pose.Pose (reference: dro_sfm/geometry/pose.py:7-98).

Same constructor, class methods and operators.  ``from_vec(vec, 'euler')`` is one kernel launch
(drosfm_pose_vec2mat) instead of ~40 ATen ops; the 4x4 bookkeeping (inverse, composition) stays in
PyTorch -- it is B x 16 numbers and provides autograd for free.
"""
import torch

from .. import ops


class Pose:
    """Encapsulates a [B,4,4] rigid transformation."""

    def __init__(self, mat):
        assert tuple(mat.shape[-2:]) == (4, 4)
        if mat.dim() == 2:
            mat = mat.unsqueeze(0)
        assert mat.dim() == 3
        self._mat = mat
        self._vec = None            # euler vector this pose was built from (from_vec), if any
        self._is_identity = False

    @property
    def mat(self):
        """[B,4,4] matrix; built on first use when the pose came from ``from_vec`` (one kernel launch)."""
        if self._mat is None:
            self._mat = ops.pose_vec2mat(self._vec)
        return self._mat

    @mat.setter
    def mat(self, value):
        # a new matrix is no longer known to be the identity (view_synthesis picks its fused kernel on that flag);
        # to() / repeat() restore the flag themselves because they keep the values
        self._mat, self._vec, self._is_identity = value, None, False

    @staticmethod
    def materialize(poses):
        """[B,4,4] matrices of a list of poses (Pose objects or tensors).  Poses that still are un-converted euler vectors
        (from_vec) of one shape are converted by ONE launch for the whole list instead of one each -- the supervised loss
        needs the matrices of all V x n predicted poses at once."""
        lazy = [k for k, p in enumerate(poses) if isinstance(p, Pose) and p._mat is None and p._vec is not None]
        if len(lazy) > 1 and len({tuple(poses[k]._vec.shape) for k in lazy}) == 1 \
                and len({(poses[k]._vec.dtype, poses[k]._vec.device) for k in lazy}) == 1:
            vecs = torch.stack([poses[k]._vec for k in lazy], dim=0)                  # [m, B, 6]
            mats = ops.pose_vec2mat(vecs.reshape(-1, 6)).reshape(len(lazy), -1, 4, 4)
            for k, m in zip(lazy, mats.unbind(0)):
                poses[k]._mat = m
        return [p.mat if hasattr(p, "mat") else p for p in poses]

    def kernel_arg(self):
        """What the fused kernels consume: the [B,6] euler vector when known (the conversion then runs in
        the kernel prologue and the gradient lands on the vector directly), else the [B,4,4] matrix."""
        return self._vec if (self._vec is not None and self._mat is None) else self.mat

    def __len__(self):
        return len(self._vec) if self._mat is None else len(self._mat)

    @classmethod
    def identity(cls, N=1, device=None, dtype=torch.float):
        pose = cls(torch.eye(4, device=device, dtype=dtype).repeat([N, 1, 1]))
        pose._is_identity = True
        return pose

    @classmethod
    def from_vec(cls, vec, mode):
        """[B,6] = (tx,ty,tz,rx,ry,rz) -> Pose; reference pose.py:38-45 + pose_utils.py:73-85."""
        if mode is None:
            return cls(vec)
        if mode != "euler":
            raise NotImplementedError("dro_sfm_b200: rotation mode {!r} is not supported (configs use 'euler', "
                                      "configs/default_config.py:93)".format(mode))
        if not vec.is_cuda or vec.dim() != 2 or vec.shape[-1] != 6:
            return cls(ops.pose_vec2mat(vec))
        pose = cls.__new__(cls)
        pose._mat, pose._vec, pose._is_identity = None, vec, False
        return pose

    @property
    def shape(self):
        return self.mat.shape

    def item(self):
        return self.mat

    def repeat(self, *args, **kwargs):
        ident = self._is_identity
        self.mat = self.mat.repeat(*args, **kwargs)
        self._is_identity = ident
        return self

    def inverse(self):
        """[R|t]^-1 = [R^T | -R^T t] (reference pose_utils.py:89-94)."""
        T = self.mat
        Tinv = torch.eye(4, device=T.device, dtype=T.dtype).repeat([len(T), 1, 1])
        Tinv[:, :3, :3] = torch.transpose(T[:, :3, :3], -2, -1)
        Tinv[:, :3, -1] = torch.bmm(-1. * Tinv[:, :3, :3], T[:, :3, -1].unsqueeze(-1)).squeeze(-1)
        out = Pose(Tinv)
        out._is_identity = self._is_identity
        return out

    def to(self, *args, **kwargs):
        ident = self._is_identity
        self.mat = self.mat.to(*args, **kwargs)
        self._is_identity = ident
        return self

    def transform_pose(self, pose):
        assert tuple(pose.shape[-2:]) == (4, 4)
        return Pose(self.mat.bmm(pose.item()))

    def transform_points(self, points):
        """R X + t for [B,3,H,W] (or [B,3,N]) points.  Inside view synthesis / cost / loss this
        step is fused into the kernels; the stand-alone operator is kept for API parity."""
        assert points.shape[1] == 3
        shape = points.shape
        out = self.mat[:, :3, :3].bmm(points.reshape(shape[0], 3, -1)) + self.mat[:, :3, -1].unsqueeze(-1)
        return out.view(shape)

    def __matmul__(self, other):
        if isinstance(other, Pose):
            return self.transform_pose(other)
        if isinstance(other, torch.Tensor):
            if other.shape[1] == 3 and other.dim() > 2:
                assert other.dim() == 3 or other.dim() == 4
                return self.transform_points(other)
            raise ValueError("Unknown tensor dimensions {}".format(other.shape))
        raise NotImplementedError()
