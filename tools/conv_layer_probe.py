"""Probe (not part of the product): cuDNN forward / backward time of single NHWC 3x3 convolutions around the
408 -> 96 decoder layer that falls back to an sm80 kernel."""
import torch
import torch.nn.functional as F

torch.backends.cudnn.benchmark = True
CL = torch.channels_last


def t(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    ts = []
    for _ in range(5):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); g.replay(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    return sorted(ts)[2] * 1e3


for cin, cout in [(408, 96), (408, 128), (408, 64), (408, 32), (416, 96), (384, 96), (448, 96), (408, 192), (504, 64), (280, 128)]:
    x = torch.randn(16, cin, 96, 128, device="cuda").contiguous(memory_format=CL).requires_grad_(True)
    w = (torch.randn(cout, cin, 3, 3, device="cuda") * 0.01).contiguous(memory_format=CL).requires_grad_(True)
    y = F.conv2d(x, w, None, 1, 1)
    gy = torch.randn_like(y)
    fwd = t(lambda: F.conv2d(x, w, None, 1, 1))
    bwd = t(lambda: torch.ops.aten.convolution_backward(gy, x, w, None, [1, 1], [1, 1], [1, 1], False, [0, 0], 1, [True, True, False]))
    gf = 2 * cin * cout * 9 * 16 * 96 * 128 / 1e9
    print("%4d -> %3d   fwd %7.1f us (%5.0f TF/s)   bwd %7.1f us (%5.0f TF/s)" % (cin, cout, fwd, gf / fwd * 1e-3 * 1e3, bwd, 2 * gf / bwd * 1e-3 * 1e3), flush=True)
