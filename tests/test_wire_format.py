"""Opt-in compact wire format CVW1 (include/cvs_b200.h): the GPU encoder and decoder
against the host-side parser, the oracle's payload and the client round trip `reconstructed == reference`.
Default format stays the reference's (threads.cpp:229-231 / client/opencv.cpp:52-66)."""
import ctypes as C
import struct

import numpy as np
import pytest

from util import random_sequence


def _encode_numpy(xs, diff, n):
    """Specification-level encoder (numpy), the twin of cudavideostream_b200.wire.parse."""
    from cudavideostream_b200 import wire
    nt = (n + wire.TILE - 1) // wire.TILE
    pos = int(xs.size)
    counts = np.bincount(xs // wire.TILE, minlength=nt).astype(np.uint8)
    out = np.zeros(wire.HEADER + wire._pad16(nt) + wire._pad16(pos) + pos, dtype=np.uint8)
    out[:16] = np.frombuffer(struct.pack("<4I", wire.MAGIC, pos, nt, wire.TILE), dtype=np.uint8)
    out[16:16 + nt] = counts
    o0 = 16 + wire._pad16(nt)
    out[o0:o0 + pos] = (xs % wire.TILE).astype(np.uint8)
    d0 = o0 + wire._pad16(pos)
    out[d0:d0 + pos] = diff
    return out


def test_parser_round_trip_on_the_oracle_payload(oracle):
    from cudavideostream_b200 import wire
    for w, h, dens in [(7, 5, 0.5), (64, 48, 0.1), (250, 130, 0.01), (64, 48, 1.0), (64, 48, 0.0)]:
        n = 3 * w * h
        base, frames = random_sequence(w, h, 1, dens, seed=w)
        pos, xs, diff, _, _ = oracle.diff_compact(frames[0], base, 20)
        enc = _encode_numpy(xs, diff, n)
        assert wire.size_of(enc) == enc.size == 16 + wire._pad16(wire.ntiles(w, h)) + wire._pad16(pos) + pos
        p2, xs2, diff2 = wire.parse(enc)
        assert p2 == pos and np.array_equal(xs2, xs) and np.array_equal(diff2, diff)
    with pytest.raises(ValueError):
        wire.parse(np.zeros(64, dtype=np.uint8))


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,density", [(7, 5, 0.3), (64, 48, 0.0), (64, 48, 1.0), (250, 130, 0.02), (641, 359, 0.3),
                                         (1920, 1080, 0.1)])
def test_submit_wire_matches_the_oracle_and_round_trips(cvs, oracle, w, h, density):
    import torch
    from cudavideostream_b200 import wire
    n = 3 * w * h
    T = 5 if w < 1000 else 3
    base, frames = random_sequence(w, h, T, density, seed=w + 1)
    s = cvs.Stream(w, h, base)
    oc = oracle.OracleCore(w, h, base)
    fin = cvs.alloc_host(n + 32)
    wout = cvs.alloc_host(wire.bound(w, h) + 64)
    st = torch.cuda.current_stream().cuda_stream
    client = torch.from_numpy(base).cuda()
    client = torch.cat([client, torch.zeros(64, dtype=torch.uint8, device="cuda")])
    d_scr = torch.zeros(wire.scratch_words(w, h), dtype=torch.int32, device="cuda")
    d_xs = torch.zeros(n + 16, dtype=torch.int32, device="cuda")
    d_df = torch.zeros(n + 16, dtype=torch.uint8, device="cuda")
    d_pos = torch.zeros(1, dtype=torch.int32, device="cuda")
    for t, f in enumerate(frames):
        fin.array()[:n] = f
        wout.array()[:] = 0xA5
        tk = s.submit_wire_raw(fin.ptr, wout.ptr, None, "")
        s.wait(tk)
        opos, oxs, odiff, _, _ = oc.exec_core(f)
        enc = wout.array()
        size = wire.size_of(enc)
        assert size == 16 + wire._pad16(wire.ntiles(w, h)) + wire._pad16(opos) + opos
        assert bool((enc[wire.bound(w, h):] == 0xA5).all()), "write past cvs_wire_bound"
        pos, xs, diff = wire.parse(enc[:size])
        assert pos == opos and np.array_equal(xs, oxs) and np.array_equal(diff, odiff), f"frame {t}"
        assert np.array_equal(enc[:size], _encode_numpy(oxs, odiff, n)), f"frame {t}: encoded bytes"
        # client side on the GPU: decode into the reference format and apply to the client's frame
        d_wire = torch.from_numpy(enc[:(size + 15) // 16 * 16].copy()).cuda()
        wire.decode_device(d_wire.data_ptr(), d_scr.data_ptr(), w, h, d_frame=client.data_ptr(), d_xs=d_xs.data_ptr(),
                           d_diff=d_df.data_ptr(), d_pos=d_pos.data_ptr(), stream=st)
        wire.decode_status(d_scr.data_ptr(), w, h, st)
        assert int(d_pos[0]) == opos
        assert np.array_equal(d_xs[:opos].cpu().numpy(), oxs) and np.array_equal(d_df[:opos].cpu().numpy(), odiff)
        assert np.array_equal(client[:n].cpu().numpy(), oc.reference()), f"frame {t}: reconstructed != reference"
    assert np.array_equal(s.reference(), oc.reference())
    s.close()


@pytest.mark.gpu
def test_wire_pageable_buffer_and_device_encoder(cvs, oracle):
    import torch
    from cudavideostream_b200 import wire
    w, h = 320, 180
    n = 3 * w * h
    base, frames = random_sequence(w, h, 3, 0.15, seed=8)
    s = cvs.Stream(w, h, base)
    oc = oracle.OracleCore(w, h, base)
    out = np.zeros(wire.bound(w, h), dtype=np.uint8)      # pageable: delivered by a copy in cvs_wait
    for f in frames:
        fr = np.ascontiguousarray(f)
        tk = s.submit_wire_raw(fr.ctypes.data, out.ctypes.data, None, "")
        s.wait(tk)
        opos, oxs, odiff, _, _ = oc.exec_core(f)
        pos, xs, diff = wire.parse(out[:wire.size_of(out)])
        assert pos == opos and np.array_equal(xs, oxs) and np.array_equal(diff, odiff)
    s.close()
    # the stand-alone device encoder on a device-resident payload (what a sequence launch leaves behind)
    st = torch.cuda.current_stream().cuda_stream
    opos, oxs, odiff, _, _ = oracle.diff_compact(frames[0], base, 20)[:5]
    d_xs = torch.zeros(n + 16, dtype=torch.int32, device="cuda")
    d_df = torch.zeros(n + 16, dtype=torch.uint8, device="cuda")
    d_xs[:opos] = torch.from_numpy(oxs).cuda()
    d_df[:opos] = torch.from_numpy(odiff).cuda()
    d_pos = torch.tensor([opos], dtype=torch.int32, device="cuda")
    d_scr = torch.zeros(wire.scratch_words(w, h), dtype=torch.int32, device="cuda")
    d_wire = torch.zeros(wire.bound(w, h) + 16, dtype=torch.uint8, device="cuda")
    wire.encode_device(d_xs.data_ptr(), d_df.data_ptr(), d_pos.data_ptr(), n, w, h, d_scr.data_ptr(), d_wire.data_ptr(), st)
    torch.cuda.synchronize()
    enc = d_wire.cpu().numpy()
    assert np.array_equal(enc[:wire.size_of(enc)], _encode_numpy(oxs, odiff, n))
    # a corrupted frame is rejected as a whole
    bad = d_wire.clone()
    bad[16] = (int(bad[16]) + 1) % 193
    frame = torch.zeros(n + 64, dtype=torch.uint8, device="cuda")
    wire.decode_device(bad.data_ptr(), d_scr.data_ptr(), w, h, d_frame=frame.data_ptr(), stream=st)
    with pytest.raises(cvs.CVSError):
        wire.decode_status(d_scr.data_ptr(), w, h, st)
    assert int(frame.sum()) == 0
