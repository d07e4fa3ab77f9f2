"""CPU: the oracle's restatement of the visual-prompting frame transform and chunked encode (oracle/restate.py
frame_transform / vision_encode) against (a) torchvision itself — the transform the reference builds
(vision_encoder.py:91-113) — and (b) golden outputs of the reference's own PerceptionEncoder class (tests/golden/vision.pt,
third-party CLIP tower replaced by a stand-in)."""
import os

import torch

from oracle import restate
from oracle.make_golden import FakeClip
from _util import rel_l2

torch.set_grad_enabled(False)


def _levels(x):
    return ((x * 0.5 + 0.5) * 255).round()


def test_frame_transform_matches_torchvision():
    import torchvision
    T = torchvision.transforms
    tf = T.Compose([T.Resize((336, 336), interpolation=T.InterpolationMode.BICUBIC),
                    T.Lambda(lambda x: x.float() / 255.0), T.Normalize([0.5] * 3, [0.5] * 3, inplace=True)])
    g = torch.Generator().manual_seed(1)
    for shape in [(3, 3, 36, 64), (1, 3, 360, 640), (1, 3, 500, 300), (2, 3, 336, 336)]:     # up, down, mixed, identity
        v = torch.randint(0, 256, shape, generator=g, dtype=torch.uint8)
        a, b = restate.frame_transform(v), tf(v)
        la, lb = _levels(a), _levels(b)
        # fp32 summation order differs in the last bits: a few pixels in 10^5 land on the other side of a uint8
        # rounding boundary (by one level); everything else is bit-equal
        assert float((la != lb).float().mean()) < 1e-4 and float((la - lb).abs().max()) <= 1
        assert torch.equal(a[la == lb], b[la == lb])


def test_chunked_encode_and_levels_match_reference_golden(golden_dir):
    g = torch.load(os.path.join(golden_dir, "vision.pt"))
    gen = torch.Generator().manual_seed(g["seed"])
    vids = [torch.randint(0, 256, (n, 3, 20, 28), generator=gen, dtype=torch.uint8) for n in g["video_lens"]]
    big = torch.randint(0, 256, (2, 3, 360, 640), generator=gen, dtype=torch.uint8)
    fake = FakeClip()
    assert torch.equal(fake.proj, g["proj"])
    out = restate.vision_encode(vids, lambda x: fake.encode_image(x, normalize=True), 336, 300)
    assert out.shape == g["feats"].shape and rel_l2(out, g["feats"]) < 1e-4
    assert float(out[1, 7:].abs().max()) == 0.0 and float(out[2, 1:].abs().max()) == 0.0     # pad_sequence zeros
    lv = _levels(restate.frame_transform(big)).to(torch.uint8)
    assert float((lv != g["big_levels"]).float().mean()) < 1e-4


def test_native_host_tap_tables_equal_the_restatement():
    """The C++ host code that builds the resize kernels' tap windows and weights (engine.cu aa_taps, exported through
    the host-only seam sab_test_aa_taps; no GPU involved) against the oracle's float32 restatement: bit-equal."""
    import ctypes
    import numpy as np
    import __graft_entry__ as g
    g.build()
    from sam_audio_b200 import _capi
    lib = _capi.lib()
    for n_in, n_out in [(640, 336), (360, 336), (36, 336), (336, 336), (1920, 336), (5, 336), (300, 7)]:
        cap = 256
        taps = ctypes.c_int(0)
        lo = np.zeros(n_out, np.int32)
        cnt = np.zeros(n_out, np.int32)
        w = np.zeros((n_out, cap), np.float32)
        _capi.check(lib.sab_test_aa_taps(n_in, n_out, cap, ctypes.addressof(taps), lo.ctypes.data, cnt.ctypes.data, w.ctypes.data))
        ref = restate._aa_cubic_taps(n_in, n_out)
        for i, (rlo, rws) in enumerate(ref):
            assert lo[i] == rlo and cnt[i] == len(rws) <= taps.value
            assert np.array_equal(w[i, : len(rws)], np.array(rws, np.float32)), (n_in, n_out, i)
