"""nvJPEG backends for cvs_decode_jpeg_device: decode rate and distance from OpenCV's (libjpeg-turbo) pixels on the
reference's fixture frames.  Run on the GPU box: python scripts/jpeg_probe.py"""
import os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) == 1:
    for be in ("hw", "gpu", "default"):
        for interp in ("0", "1"):
            env = dict(os.environ, CVS_JPEG_BACKEND=be, CVS_JPEG_INTERP=interp)
            r = subprocess.run([sys.executable, __file__, "run"], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=300)
            print(f"backend={be} interp={interp}:", r.stdout.decode().strip().splitlines()[-1] if r.stdout.strip() else "no output", flush=True)
    sys.exit(0)
sys.path.insert(0, ROOT)
import numpy as np, torch, cv2
import cudavideostream_b200 as cvs
w, h = 1920, 1080
n = 3 * w * h
gold = os.path.join(ROOT, "tests", "golden")
j = [open(os.path.join(gold, f), "rb").read() for f in ("k1_f1.jpg", "k1_f2.jpg")]
c = [cv2.imread(os.path.join(gold, f)).reshape(-1) for f in ("k1_f1.jpg", "k1_f2.jpg")]
s = cvs.Stream(w, h, np.zeros(n, dtype=np.uint8))
d = torch.zeros(n + 64, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream
g = []
for k in range(2):
    s.decode_jpeg_device(j[k], d.data_ptr(), st)
    torch.cuda.synchronize()
    g.append(d[:n].cpu().numpy().copy())
dd = np.abs(g[0].astype(np.int16) - c[0].astype(np.int16))
changed = int((np.abs(g[0].astype(np.int16) - g[1].astype(np.int16)) > 20).sum())
t0 = time.perf_counter()
K = 100
for i in range(K):
    s.decode_jpeg_device(j[i & 1], d.data_ptr(), st)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"{K/dt:7.0f} decodes/s  vs OpenCV: max |d| {dd.max()} mean {dd.mean():.4f} differing {100*(dd>0).mean():.2f} %  changed bytes f1->f2 {changed} (OpenCV 369350)")
