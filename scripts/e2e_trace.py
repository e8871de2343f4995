"""Device timeline of the interleaved three-stream end-to-end loop (CVS_TRACE=1): where does a ticket wait?
Run on the GPU box: CVS_TRACE=1 python scripts/e2e_trace.py 2> gpurun_out/e2e_trace.log"""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cudavideostream_b200 as cvs
W, H = 1920, 1080
N = 3 * W * H
R, F = 16, int(os.environ.get("FRAMES", "40"))
WIRE = os.environ.get("WIRE", "0") == "1"
st = torch.cuda.current_stream().cuda_stream
rings = []
for d in (10000, 100000, 500000):
    fr = torch.empty((R + 1) * N, dtype=torch.uint8, device="cuda")
    cvs.synth.base_frame_device(fr.data_ptr(), W, H, 1, st)
    for t in range(R):
        cvs.synth.next_frame_device(fr.data_ptr() + t * N, fr.data_ptr() + (t + 1) * N, W, H, 1, t, d, st)
    torch.cuda.synchronize()
    hb = cvs.alloc_host(R * N)
    hb.array()[:] = fr[N:].cpu().numpy()
    s = cvs.Stream(W, H, fr[:N].cpu().numpy())
    out = [(cvs.alloc_host(N + 32), cvs.alloc_host(4 * N + 32), (C.c_uint * 1)()) for _ in range(4)]
    rings.append((s, hb, out, []))
    print("stream", d, hex(s._h.value), file=sys.stderr)
order = list(range(R)) + list(range(R - 2, 0, -1))
def run(n):
    for i in range(n):
        for s, hb, out, pend in rings:
            fb, xb, pb = out[i % 4]
            if len(pend) == 4:
                s.wait(pend.pop(0))
            src = hb.ptr + order[i % len(order)] * N
            pend.append(s.submit_wire_raw(src, xb.ptr, None, "") if WIRE else
                        s.submit_io_raw(src, fb.ptr, None, "", C.addressof(pb), xb.ptr))
    for s, hb, out, pend in rings:
        for tk in pend:
            s.wait(tk)
        pend.clear()
run(8)
print("---- timed", file=sys.stderr)
t0 = time.perf_counter(); run(F); dt = time.perf_counter() - t0
print(f"wire={WIRE} {3*F/dt:8.0f} fps total  {dt/F*1e6:7.1f} us per frame-triple")
