"""Would one cuDNN call for the two stride-2 convolutions of a down ResidualUnit (unit0 + residual branch, same input) beat two?"""
import sys, torch
import torch.nn.functional as F
torch.backends.cudnn.benchmark = True
N = int(sys.argv[1]) if len(sys.argv) > 1 else 400
def timeit(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 10 * 1000
with torch.no_grad():
    for cin, cout, hw in ((16, 32, 128), (32, 64, 64), (64, 128, 32)):
        x = torch.randn(N, cin, hw, hw, device="cuda").contiguous(memory_format=torch.channels_last)
        w1 = (torch.randn(cout, cin, 3, 3, device="cuda") * 0.05).contiguous(memory_format=torch.channels_last)
        w2 = (torch.randn(2 * cout, cin, 3, 3, device="cuda") * 0.05).contiguous(memory_format=torch.channels_last)
        t1 = timeit(lambda: F.conv2d(x, w1, None, 2, 1))
        t2 = timeit(lambda: F.conv2d(x, w2, None, 2, 1))
        print(f"s2 {cin}->{cout} @{hw}: single {t1:.1f} us, two calls {2*t1:.1f} us, merged ({2*cout} out) {t2:.1f} us")
    # bottom: 128->256 3x3 and the 1x1 residual 128->256 cannot merge (different kernels); listed for reference
    for name, cin, cout, k, hw in (("d0.u1 16->16", 16, 16, 3, 128), ("u1.c 16->16", 16, 16, 3, 128), ("d1.u1 32->32", 32, 32, 3, 64), ("b.u1 256->256", 256, 256, 3, 16)):
        x = torch.randn(N, cin, hw, hw, device="cuda").contiguous(memory_format=torch.channels_last)
        w = (torch.randn(cout, cin, k, k, device="cuda") * 0.05).contiguous(memory_format=torch.channels_last)
        mb = (cin + cout) * hw * hw * 4 * N / 1e6
        print(f"{name} @{hw}: {timeit(lambda: F.conv2d(x, w, None, 1, k // 2)):.1f} us (traffic floor {mb / 6459.3 * 1000:.1f} us)")
    for name, cin, cout, hw in (("u1.T 64->16", 64, 16, 64), ("u2.T 128->32", 128, 32, 32), ("u3.T 384->64", 384, 64, 16)):
        x = torch.randn(N, cin, hw, hw, device="cuda").contiguous(memory_format=torch.channels_last)
        w = (torch.randn(cin, cout, 3, 3, device="cuda") * 0.05).contiguous(memory_format=torch.channels_last)
        mb = (cin * hw * hw + cout * 4 * hw * hw) * 4 * N / 1e6
        print(f"{name} @{hw}: {timeit(lambda: F.conv_transpose2d(x, w, None, 2, 1, 1)):.1f} us (traffic floor {mb / 6459.3 * 1000:.1f} us)")
    # ConvTranspose2d(k3, s2, p1, op1) as ONE regular 2x2 convolution with 4*Cout outputs (one group per output parity) on the
    # input padded by one (symmetric padding 1 → (H+1)x(W+1) outputs, of which [1:,1:] are used), then a depth-to-space shuffle
    for name, cin, cout, hw in (("u1.T 64->16", 64, 16, 64), ("u2.T 128->32", 128, 32, 32), ("u3.T 384->64", 384, 64, 16)):
        x = torch.randn(N, cin, hw, hw, device="cuda").contiguous(memory_format=torch.channels_last)
        w = (torch.randn(4 * cout, cin, 2, 2, device="cuda") * 0.05).contiguous(memory_format=torch.channels_last)
        print(f"{name} @{hw} as 2x2 conv -> {4*cout} ch: {timeit(lambda: F.conv2d(x, w, None, 1, 1)):.1f} us")
