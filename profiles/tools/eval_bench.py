import sys, time, torch
sys.path.insert(0, '.')
from oracle import eval_oracle as eo
from dro_sfm_b200 import ops, synthetic as syn
g = syn.gen(5)
for name, (crop, B, H, W, h, w, lo, hi) in {"kitti": ("garg", 4, 375, 1242, 320, 960, 0.1, 80.0), "scannet": ("", 12, 480, 640, 240, 320, 0.1, 10.0)}.items():
    gt = ((lo + torch.rand(B, 1, H, W, generator=g) * hi) * (torch.rand(B, 1, H, W, generator=g) < 0.3)).cuda()
    pred = (lo + torch.rand(B, 1, h, w, generator=g) * hi * 0.6).cuda()
    inv = 1.0 / pred
    def timed(f, n=20):
        for _ in range(3): f()
        torch.cuda.synchronize(); t = time.perf_counter()
        for _ in range(n): f()
        torch.cuda.synchronize(); return (time.perf_counter() - t) / n * 1e3
    for scale in (True, False):
        a = timed(lambda: ops.depth_metrics(gt, pred, lo, hi, crop, scale))
        b = timed(lambda: eo.compute_depth_metrics_torch(crop, lo, hi, gt, pred, scale))
        print(name, "metrics scale=%d  b200 %.3f ms   aten %.3f ms  x%.1f" % (scale, a, b, b / a))
    a = timed(lambda: ops.post_process_inv_depth(inv, inv, "mean"))
    b = timed(lambda: eo.post_process_inv_depth_torch(inv, inv, "mean"))
    print(name, "post_process  b200 %.3f ms   aten %.3f ms  x%.1f" % (a, b, b / a))
