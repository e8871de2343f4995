"""Times the stand-alone display filters / noise filter over a ring of L2-cold 1080p frames."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cudavideostream_b200 as cvs
W, H, T = 1920, 1080, 64
N, P = 3 * W * H, W * H
st = torch.cuda.current_stream().cuda_stream
fr = torch.empty((T + 1) * N, dtype=torch.uint8, device="cuda")
cvs.synth.base_frame_device(fr.data_ptr(), W, H, 1, st)
for t in range(T):
    cvs.synth.next_frame_device(fr.data_ptr() + t * N, fr.data_ptr() + (t + 1) * N, W, H, 1, t, 100000, st)
out = torch.empty(T * N, dtype=torch.uint8, device="cuda")
gray = torch.empty(P + 64, dtype=torch.uint8, device="cuda")
ht = torch.zeros(257, dtype=torch.int32, device="cuda")
K3 = np.full(9, 1 / 9, dtype=np.float32)
def timeit(name, fn, bytes_per_frame):
    for t in range(3): fn(t)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(T): fn(t)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1000 / T
    print(f"{name:28s} {us:7.2f} us/frame  {bytes_per_frame / us / 1e3:7.0f} GB/s algorithmic")
a = lambda t: fr.data_ptr() + t * N
b = lambda t: fr.data_ptr() + (t + 1) * N
o = lambda t: out.data_ptr() + t * N
timeit("heat map (3N)", lambda t: cvs.filters.heat_map(a(t), b(t), o(t), W, H, st), 3 * N)
timeit("red map (3N)", lambda t: cvs.filters.red_map(a(t), b(t), o(t), W, H, 20, st), 3 * N)
timeit("gray avg 3ch (2N)", lambda t: cvs.filters.grayscale(b(t), o(t), W, H, False, 3, st), 2 * N)
timeit("gray weighted 3ch (2N)", lambda t: cvs.filters.grayscale(b(t), o(t), W, H, True, 3, st), 2 * N)
timeit("gray weighted 1ch (N+P)", lambda t: cvs.filters.grayscale(b(t), o(t), W, H, True, 1, st), N + P)
timeit("binarize weighted (N+2P+N)", lambda t: cvs.filters.binarize(b(t), o(t), gray.data_ptr(), ht.data_ptr(), W, H, True, 50, 200, st), 2 * N + 2 * P)
timeit("noise filter K=3 (2N)", lambda t: cvs.filters.noise_filter(b(t), o(t), W, H, 3, K3, st), 2 * N)
timeit("noise filter K=5 (2N)", lambda t: cvs.filters.noise_filter(b(t), o(t), W, H, 5, np.full(25, 1 / 25, dtype=np.float32), st), 2 * N)
