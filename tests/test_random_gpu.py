"""Randomised parity: frame size, threshold, display mode, density, noise filter and text drawn from a seeded
generator; the CUDA path (drop-in call and device sequence alternately) must match the oracle bit for bit."""
import numpy as np
import pytest

from util import CHARS_STR, glyph_atlas, random_sequence

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", range(24))
def test_random_configuration(cvs, oracle, seed):
    import torch
    rng = np.random.default_rng(1000 + seed)
    w, h = int(rng.integers(1, 300)), int(rng.integers(1, 120))
    thr = int(rng.choice([0, 5, 20, 20, 20, 60, 127, 128, 200]))
    mode = int(rng.integers(0, 8))
    density = float(rng.choice([0.0, 0.002, 0.02, 0.1, 0.3, 0.7, 1.0]))
    nframes = int(rng.integers(1, 5))
    use_noise = bool(rng.random() < 0.3) and w >= 3 and h >= 3
    K = int(rng.choice([3, 3, 5]))
    k = oracle.gaussian_kernel(K, K * K / 6.0) if use_noise else None
    gw, gh = 5, 4
    use_text = bool(rng.random() < 0.3) and h >= gh
    atlas = glyph_atlas(gw, gh, seed) if use_text else None
    text = "".join(rng.choice(list(CHARS_STR + "xy"), size=int(rng.integers(1, 12)))) if use_text else ""
    base, frames = random_sequence(w, h, nframes, density, seed=seed)
    okw = dict(thr=thr, mode=mode, noise_filter=int(use_noise), K=K, k=k, glyphs=atlas, gw=gw if use_text else 0,
               gh=gh if use_text else 0, chars=CHARS_STR if use_text else "")
    ckw = dict(threshold=thr, mode=mode, noise_filter=use_noise, ksize=K, kweights=k, glyphs=atlas,
               glyph_w=gw if use_text else 0, glyph_h=gh if use_text else 0)
    oc = oracle.OracleCore(w, h, base, **okw)
    want = [oc.exec_core(f, text) for f in frames]
    s = cvs.Stream(w, h, base, **ckw)
    n = 3 * w * h
    if seed % 2 == 0:   # drop-in call, frame by frame
        for t, f in enumerate(frames):
            pos, xs, diff, show = s.exec(f, text)
            opos, oxs, odiff, oshow, _ = want[t]
            assert pos == opos and np.array_equal(xs, oxs) and np.array_equal(diff, odiff), (seed, t)
            if mode:
                assert np.array_equal(show, oshow), (seed, t)
    else:               # one device-resident sequence launch
        stride = (n + 15) // 16 * 16
        cap = (n + 3) // 4 * 4
        d_frames = torch.zeros(nframes * stride + 64, dtype=torch.uint8, device="cuda")
        for t in range(nframes):
            d_frames[t * stride: t * stride + n] = torch.from_numpy(frames[t]).cuda()
        d_pos = torch.zeros(nframes, dtype=torch.int32, device="cuda")
        d_xs = torch.empty(nframes * cap, dtype=torch.int32, device="cuda")
        d_diff = torch.empty(nframes * cap, dtype=torch.uint8, device="cuda")
        d_show = torch.zeros(nframes * stride, dtype=torch.uint8, device="cuda")
        s.run_sequence_device(d_frames.data_ptr(), stride, nframes, d_pos.data_ptr(), d_xs.data_ptr(), d_diff.data_ptr(),
                              cap, d_show.data_ptr() if mode else 0, stride, text=text,
                              cuda_stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        s.sequence_status()
        pos = d_pos.cpu().numpy()
        for t in range(nframes):
            opos, oxs, odiff, oshow, _ = want[t]
            assert pos[t] == opos, (seed, t)
            assert np.array_equal(d_xs[t * cap: t * cap + opos].cpu().numpy(), oxs), (seed, t)
            assert np.array_equal(d_diff[t * cap: t * cap + opos].cpu().numpy(), odiff), (seed, t)
            if mode:
                assert np.array_equal(d_show[t * stride: t * stride + n].cpu().numpy(), oshow), (seed, t)
    assert np.array_equal(s.reference(), oc.reference()), seed
    s.close()
    oc.close()
