"""Probe (not part of the product): cuDNN time of the PWC decoder's dense-block convolutions, forward + backward,
for NCHW-contiguous activations vs channels_last activations with 8-aligned channel counts."""
import sys
import torch
import torch.nn.functional as F

torch.backends.cudnn.benchmark = True
dev = "cuda"


def run(fmt, B, H, W, c0, aligned):
    widths = [128, 128, 96, 64, 32]
    cin = c0
    layers = []
    for c in widths:
        w = torch.randn(c, cin, 3, 3, device=dev) * 0.01
        x = torch.randn(B, cin, H, W, device=dev)
        if fmt == "nhwc":
            w = w.contiguous(memory_format=torch.channels_last)
            x = x.contiguous(memory_format=torch.channels_last)
        layers.append((x.requires_grad_(True), w.requires_grad_(True)))
        cin += c

    def step():
        for x, w in layers:
            y = F.conv2d(x, w, None, 1, 1)
            gy = torch.ones_like(y)
            torch.autograd.grad(y, [x, w], gy)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        step()
    ts = []
    for _ in range(5):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); g.replay(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return sorted(ts)[2]


for (B, H, W) in [(16, 96, 128), (16, 48, 64), (16, 24, 32)]:
    for c0, label in [(147, "c0=147"), (152, "c0=152 (8-aligned)")]:
        for fmt in ("nchw", "nhwc"):
            print("B%d %dx%d %-20s %s  dense block fwd+bwd %.3f ms" % (B, H, W, label, fmt, run(fmt, B, H, W, c0, True)), flush=True)
# refinement (dilated) and pyramid shapes
def conv_time(fmt, B, cin, cout, H, W, stride=1, dil=1):
    w = torch.randn(cout, cin, 3, 3, device=dev) * 0.01
    x = torch.randn(B, cin, H, W, device=dev)
    if fmt == "nhwc":
        w = w.contiguous(memory_format=torch.channels_last); x = x.contiguous(memory_format=torch.channels_last)
    x.requires_grad_(True); w.requires_grad_(True)
    def step():
        y = F.conv2d(x, w, None, stride, dil, dil)
        torch.autograd.grad(y, [x, w], torch.ones_like(y))
    for _ in range(3): step()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g): step()
    ts = []
    for _ in range(5):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); g.replay(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    return sorted(ts)[2]
for name, args in [("refine 40->128 d1", (16, 40, 128, 96, 128, 1, 1)), ("refine 128->128 d2", (16, 128, 128, 96, 128, 1, 2)),
                   ("refine 128->128 d4", (16, 128, 128, 96, 128, 1, 4)), ("refine 128->96 d8", (16, 128, 96, 96, 128, 1, 8)),
                   ("pyr 8->32 s2 384x512", (16, 8, 32, 384, 512, 2, 1)), ("pyr 32->32 192x256", (16, 32, 32, 192, 256, 1, 1)),
                   ("pyr 32->32 s2 192x256", (16, 32, 32, 192, 256, 2, 1)), ("pyr 32->32 96x128", (16, 32, 32, 96, 128, 1, 1))]:
    print("%-24s nchw %.3f ms   nhwc %.3f ms" % (name, conv_time("nchw", *args), conv_time("nhwc", *args)), flush=True)
