from .cost import get_cost_each, depth_cost_calc, cost_batch, upsample_depth, FeatureMetricCost

__all__ = ["get_cost_each", "depth_cost_calc", "cost_batch", "upsample_depth", "FeatureMetricCost"]
