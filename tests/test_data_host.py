"""CPU: the data path's oracle against the reference loader's own outputs (fixtures made by
tests/golden/make_golden_loader.py), the random-draw order, and the sharded sampler."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))

from make_golden_loader import loader_cases  # noqa: E402
from oracle import data as odata  # noqa: E402
from uncltmo_b200.data import ShardedSampler, draw_augment, draw_video_crop  # noqa: E402


@pytest.fixture(scope="module")
def loader_golden():
    return np.load(os.path.join(HERE, "golden", "reference_loader_outputs.npz"))


def oracle_case(img, kw):
    """Replays npy_loader for one file: seeded draws in the reference's order, two crops."""
    np.random.seed(kw["seed"])
    h, w = img.shape[:2]
    outs = []
    for _ in range(2):
        rh, rw, xx, yy = draw_video_crop(h, w) if kw["video"] else draw_augment(h, w, always_resize=kw["ldr_neg"])
        f = kw["lam"] * 255 * 0.1 if kw["hdr_mode"] else None
        outs.append(odata.crop_sample(img, rh, rw, xx, yy, kw["hdr_mode"], f, kw["normalization"], 1.05, 0.1))
    return outs


@pytest.mark.parametrize("name", ["hdr", "ldr", "neg", "vid"])
def test_oracle_matches_reference_loader(loader_golden, name):
    img, kw = loader_cases()[name]
    outs = oracle_case(img, kw)
    for k, (inp, color, gnorm, gray) in enumerate(outs):
        np.testing.assert_array_equal(color.numpy()[..., ::4, ::4], loader_golden[name + "_color"][k])
        np.testing.assert_allclose(inp.numpy()[..., ::4, ::4], loader_golden[name + "_input"][k], rtol=0, atol=1e-7)
        if kw["hdr_mode"]:
            np.testing.assert_allclose(gnorm.numpy()[..., ::4, ::4], loader_golden[name + "_gray_norm"][k], rtol=0, atol=1e-7)
            np.testing.assert_allclose(gray.numpy()[..., ::4, ::4], loader_golden[name + "_gray"][k], rtol=1e-6, atol=0)
    if kw["hdr_mode"]:
        assert float(loader_golden[name + "_factor"]) == pytest.approx(kw["lam"] * 255 * 0.1, rel=1e-6)


def test_draw_order_follows_reference():
    np.random.seed(5)
    a = draw_augment(300, 400, always_resize=False)
    np.random.seed(5)
    mode = np.random.randint(0, 2)
    rh = 256 if mode == 0 else int(np.random.uniform(256, 512))
    want = (rh, rh, 0, 0) if rh == 256 else (rh, rh, np.random.randint(0, rh - 256), np.random.randint(0, rh - 256))
    assert a == want
    state = np.random.get_state()[1].copy()
    assert draw_augment(256, 256, always_resize=False) == (256, 256, 0, 0)     # no draw at all for a 256^2 file
    assert (np.random.get_state()[1] == state).all()
    with pytest.raises(ValueError):
        draw_video_crop(300, 400)


def test_sharded_sampler_partitions_each_epoch():
    n, world = 37, 4
    shards = [ShardedSampler(n, world, r, seed=3) for r in range(world)]
    for epoch in (0, 1):
        seen = []
        for s in shards:
            s.set_epoch(epoch)
            idx = s.indices()
            assert len(idx) == len(s) == 10
            seen += idx
        assert set(seen) == set(range(n)) and len(seen) == 40          # padded by wrapping, as DistributedSampler does
    a0, a1 = ShardedSampler(n, 1, 0, seed=3), ShardedSampler(n, 1, 0, seed=3)
    a1.set_epoch(1)
    assert a0.indices() != a1.indices() and sorted(a0.indices()) == list(range(n))
    assert ShardedSampler(5, 1, 0, shuffle=False).indices() == [0, 1, 2, 3, 4]
    with pytest.raises(ValueError):
        ShardedSampler(5, 2, 2)
