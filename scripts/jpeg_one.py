"""A few decodes of the reference's fixture frame through cvs_decode_jpeg_device (for ncu launch lists)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import cudavideostream_b200 as cvs
w, h = 1920, 1080
n = 3 * w * h
j = open(os.path.join(ROOT, "tests", "golden", "k1_f1.jpg"), "rb").read()
s = cvs.Stream(w, h, np.zeros(n, dtype=np.uint8))
d = torch.zeros(n + 64, dtype=torch.uint8, device="cuda")
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    s.decode_jpeg_device(j, d.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
s.close()
