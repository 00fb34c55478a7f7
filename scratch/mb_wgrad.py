import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from building_gan_b200 import lib
dev = "cuda"
N = 15145
def timeit(fn, it=20):
    """GPU time per call with the CPU launch cost removed: `it` calls captured in one CUDA graph, replayed."""
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn()
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            for _ in range(it): fn()
    torch.cuda.synchronize()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (5 * it) * 1e3
for cout, k, ones in [(2, 32, False), (32, 1, True), (32, 64, False), (64, 36, True), (64, 64, True), (1, 8, True), (16, 8, False),
                      (128, 524, True), (128, 128, True), (64, 128, True), (7, 16, True)]:
    gz = torch.randn(N, cout, device=dev)
    x = torch.randn(N, k, device=dev)
    segs = [None] if (ones and k == 1) else ([x, None] if ones else [x])
    dW = torch.empty(cout, sum(1 if s is None else s.shape[1] for s in segs), device=dev)
    t = timeit(lambda: lib.dense_wgrad(gz, segs, dW=dW))
    print(f"wgrad Cout={cout:4d} K={k:4d} ones={ones}: {t:8.1f} us")
for cin, cout in [(128, 128), (268, 128), (524, 128), (64, 32), (36, 64), (8, 1), (128, 64)]:
    x = torch.randn(N, cin, device=dev); W = torch.randn(cout, cin, device=dev); b = torch.randn(cout, device=dev)
    t = timeit(lambda: lib.dense_fwd([x], W, b, None, 1))
    print(f"dense_fwd {cin}->{cout}: {t:8.1f} us  ({2*N*cin*cout/t/1e6:.2f} TFLOP/s)")
for C in [1, 8, 32, 64, 128]:
    o = torch.randn(N, C, device=dev); one = torch.ones(C, device=dev); zero = torch.zeros(C, device=dev)
    t = timeit(lambda: lib.graphnorm_fwd(o, one, zero, one, None, 0.8, 1, 2))
    print(f"graphnorm_fwd C={C}: {t:8.1f} us")
