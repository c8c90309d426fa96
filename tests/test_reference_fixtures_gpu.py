"""GPU side of the reference's own fixtures:
  * example.png, the reference's only shipped file and fpng_test's default input (src/fpng_test.cpp:1118, 1237-1327):
    decoded on the GPU to 3 and 4 channels, re-encoded (687 x 1012 RGB: 2061-byte scanlines, no alignment) 1-pass and
    2-pass and byte-compared with the reference;
  * `fpng_test -e`: the six mutation families with the reference's seeds (trial number), 457 of the 1000 trials 1-pass (all 276 trials of the five rare families + every fourth bit-flip trial) and
    every fifth of those 2-pass, byte-compared with the reference encoder and round-tripped through the GPU decoder;
  * `fpng_test -E`: the first 200 trials of the default-seeded session (w, h in [1, 8194], 3/4 channels, uniform random).
"""
import numpy as np
import pytest

import fuzzgen
from test_reference_fixtures_cpu import example_bytes, fuzz_golden

pytestmark = pytest.mark.gpu


def test_example_png_decode_gpu(gpu, ref):
    data = example_bytes()
    assert gpu.fpng_get_info(data) == (0, 687, 1012, 3)
    err, lode3, w, h = ref.lodepng_decode(data, 3)
    assert err == 0
    for desired in (3, 4):
        st, px, ww, hh, cc = gpu.fpng_decode_memory(data, desired)
        rst, rpx, *_ = ref.decode(data, desired)
        assert st == rst == 0 and (ww, hh, cc) == (687, 1012, 3)
        assert np.array_equal(px, rpx)
    st, px3, *_ = gpu.fpng_decode_memory(data, 3)
    assert np.array_equal(px3, lode3)


@pytest.mark.parametrize("flags", [0, 1, 2])
def test_example_png_reencode_gpu(gpu, ref, flags):
    err, px, w, h = ref.lodepng_decode(example_bytes(), 3)
    ok, png = gpu.fpng_encode_image_to_memory(px, w, h, 3, flags)
    assert ok and png == ref.encode(px, w, h, 3, flags)
    err, back, *_ = ref.lodepng_decode(png, 3)                     # lodepng checks the IDAT CRC and the Adler-32
    assert err == 0 and np.array_equal(back, px)
    comp, stb, *_ = ref.stb_decode(png, 3)
    assert comp and np.array_equal(stb, px)
    # fpng_test's channel-conversion checks (src/fpng_test.cpp:1276-1327)
    st, px4, *_ = gpu.fpng_decode_memory(png, 4)
    q = px4.reshape(-1, 4)
    assert st == 0 and np.array_equal(q[:, :3].reshape(-1), px) and (q[:, 3] == 255).all()
    # with the alpha = green swizzle of `fpng_test -a` (src/fpng_test.cpp:1147-1152): a 32bpp file
    rgba = np.concatenate([px.reshape(-1, 3), px.reshape(-1, 3)[:, 1:2]], axis=1).reshape(-1)
    ok, png4 = gpu.fpng_encode_image_to_memory(rgba, w, h, 4, flags)
    assert ok and png4 == ref.encode(rgba, w, h, 4, flags)
    st, back3, *_ = gpu.fpng_decode_memory(png4, 3)
    assert st == 0 and np.array_equal(back3, px)


def _encode_many(gpu, bufs, w, h, c, flags):
    """One device batch per call: the inputs of a fuzz chunk share their dimensions."""
    import torch
    dev = torch.from_numpy(np.stack(bufs).reshape(len(bufs), h, w, c)).cuda()
    out, sizes = gpu.encode_batch_device(dev, flags)
    torch.cuda.synchronize()
    sizes = sizes.cpu().numpy().astype(np.int64)
    out = out.cpu().numpy()
    return [out[i, : sizes[i]].tobytes() for i in range(len(bufs))]


def test_fuzz_e_reference_seeds(gpu, ref):
    g = fuzz_golden()
    err, src, w, h = ref.lodepng_decode(example_bytes(), 3)
    fams = set()
    chunk = 25
    # every trial of the five rare families (276 of the 1000) and every fourth bit-flip trial (181 more): the bit-flip
    # generator alone costs 16.7 M rand() calls per trial
    chosen = [t for t in range(1000) if g["e_family"][t] != 5 or t % 4 == 0]
    assert len(chosen) >= 400
    for c0 in range(0, len(chosen), chunk):
        trials = chosen[c0:c0 + chunk]
        t0 = trials[0]
        bufs, fam = zip(*[fuzzgen.mutate(t, src, 3) for t in trials])
        fams.update(fam)
        pngs = _encode_many(gpu, list(bufs), w, h, 3, 0)
        for t, buf, f, png in zip(trials, bufs, fam, pngs):
            assert f == g["e_family"][t] and len(png) == g["e_sizes"][t], (t, f, len(png))
            if t % 8 == 0 or f != 5:                                  # byte-compare every rare-family trial and half of the bit-flip ones
                assert png == ref.encode(buf, w, h, 3, 0), (t, f)
        # decode side of the reference's loop: fpng decode to 4 channels must give the mutated pixels + 0xFF (src/fpng_test.cpp:560-606)
        for t, buf, png in list(zip(trials, bufs, pngs))[::5]:
            st, px, ww, hh, cc = gpu.fpng_decode_memory(png, 4)
            q = px.reshape(-1, 4)
            assert st == 0 and (ww, hh, cc) == (w, h, 3) and np.array_equal(q[:, :3].reshape(-1), buf) and (q[:, 3] == 255).all(), t
        # 2-pass (fpng_test -s -e) on every fifth trial
        sub = [i for i in range(len(trials)) if trials[i] % 5 == 0]
        pngs2 = _encode_many(gpu, [bufs[i] for i in sub], w, h, 3, 1) if sub else []
        for i, png in zip(sub, pngs2):
            assert png == ref.encode(bufs[i], w, h, 3, 1), (trials[i], "2-pass")
    assert fams == {0, 1, 2, 3, 4, 5}


def test_fuzz_E_reference_seeds(gpu, ref):
    g = fuzz_golden()
    s = fuzzgen.DimSession()
    n_big = 0
    try:
        for t in range(200):
            w, h, c, buf = s.next()
            if t < len(g["E_trials"]):
                assert [w, h, c] == g["E_trials"][t][:3], t
            ok, png = gpu.fpng_encode_image_to_memory(buf, w, h, c, 0)
            assert ok, (t, w, h, c)
            if t < len(g["E_trials"]):
                assert len(png) == g["E_trials"][t][3], (t, w, h, c)
            if w * h <= 12_000_000 or t % 8 == 0:
                assert png == ref.encode(buf, w, h, c, 0), (t, w, h, c)
            else:
                n_big += 1
            st, px, ww, hh, cc = gpu.fpng_decode_memory(png, c)       # what the reference's -E loop checks
            assert st == 0 and (ww, hh, cc) == (w, h, c) and np.array_equal(px, buf), (t, w, h, c)
    finally:
        s.close()
