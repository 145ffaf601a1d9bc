from .depth import post_process_inv_depth, compute_depth_metrics

__all__ = ["post_process_inv_depth", "compute_depth_metrics"]
