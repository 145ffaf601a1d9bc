"""What section 8f-1 could save: the 1x1 convolution + ReLU that consumes the cost map (update.py:81,90), as cuDNN runs it."""
import torch, json
torch.backends.cudnn.benchmark = True
res = {}
for name, (B, C, h, w) in {"kitti_320x960": (2, 128, 40, 120), "scannet_240x320": (12, 128, 30, 40)}.items():
    for tf32 in (True, False):
        torch.backends.cudnn.allow_tf32 = tf32
        conv = torch.nn.Conv2d(C, 128, 1).cuda().to(memory_format=torch.channels_last)
        x = torch.randn(B, C, h, w, device="cuda").contiguous(memory_format=torch.channels_last).requires_grad_(True)
        def fwd():
            return torch.relu(conv(x))
        for _ in range(5): y = fwd(); y.sum().backward()
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            with torch.no_grad():
                for _ in range(3): fwd()
        torch.cuda.current_stream().wait_stream(s)
        with torch.no_grad(), torch.cuda.graph(g):
            for _ in range(24): y = fwd()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20): g.replay()
        b.record(); torch.cuda.synchronize()
        res["%s tf32=%d" % (name, tf32)] = round(a.elapsed_time(b) / 20 / 24 * 1e3, 2)
print(json.dumps({"conv1x1+relu us per call (graph replay, 24 calls back to back)": res}))
