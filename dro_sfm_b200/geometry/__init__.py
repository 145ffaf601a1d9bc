from .pose import Pose
from .camera import Camera
from .camera_utils import view_synthesis, scale_intrinsics, construct_K

__all__ = ["Pose", "Camera", "view_synthesis", "scale_intrinsics", "construct_K"]
