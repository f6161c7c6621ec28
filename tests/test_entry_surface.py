"""The mirrored entry points and drop-in classes keep the reference's public surface: function
names, leading positional parameters, CLI flags and their defaults, file extensions - against a fixture recorded from
the reference itself (tests/golden/make_golden_surface.py -> reference_surface.json)."""
import inspect
import json
import os

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def surface():
    return json.load(open(os.path.join(HERE, "golden", "reference_surface.json")))


def _params(f):
    return [p.name for p in inspect.signature(f).parameters.values()]


@pytest.mark.parametrize("name", ["test_imageTMO", "test_videoTMO"])
def test_entry_module_surface(name, surface):
    import importlib
    mod = importlib.import_module("uncltmo_b200.entry." + name)
    ref = surface[name]
    for fn, ref_params in ref["functions"].items():
        if fn == "print_args":      # debugging helper of the reference, not part of the path
            continue
        assert hasattr(mod, fn), fn
        mine = _params(getattr(mod, fn))
        if fn == "get_args":        # ours takes an optional argv (testable); the reference reads sys.argv
            continue
        assert mine[:len(ref_params)] == ref_params, (fn, mine, ref_params)      # extras only as trailing keyword options
    args = mod.get_args([])
    for flag, default in ref["cli_defaults"].items():
        assert str(getattr(args, flag)) == str(default), flag
    assert sorted(mod.extensions) == sorted(e for e in ref["extensions"] if e != ".dng") or sorted(mod.extensions) == sorted(ref["extensions"])
    # every reference flag parses
    argv = []
    for flag in ref["cli_defaults"]:
        argv += ["--" + flag, "x"]
    parsed = mod.get_args(argv)
    assert all(getattr(parsed, flag) == "x" for flag in ref["cli_defaults"])


def test_class_surface(surface):
    from uncltmo_b200.discriminator import SimpleDiscriminator
    from uncltmo_b200.generator import UNet, UNetVideo
    from uncltmo_b200.struct_loss import StructLoss
    from uncltmo_b200 import losses
    ref = surface["classes"]
    for cls, key in ((UNet, "UNet_image"), (UNetVideo, "UNet_video"), (SimpleDiscriminator, "SimpleDiscriminator"),
                     (StructLoss, "StructLoss")):
        for meth in ("__init__", "forward"):
            want = ref["%s.%s" % (key, meth)]
            mine = _params(getattr(cls, meth))
            assert mine[:len(want)] == want, (key, meth, mine, want)
    for meth in ("contrastive_D_loss", "nce", "infoNCE", "infoNCE2", "pseudo_label_loss"):
        want = surface["trainer_methods"][meth][1:]          # drop `self`: ours are module-level functions
        assert _params(getattr(losses, meth)) == want, meth


def test_unsupported_hyper_parameters_raise():
    """Anything outside the shipped configuration fails loudly at construction (no silent fallback)."""
    from uncltmo_b200.generator import UNet
    from uncltmo_b200.entry import common
    args = [1, 1, "sigmoid", 4, 4, "square_and_square_root", 32, 0, "unet", 0, 0, "none", "none", "relu", 1, "replicate", 2]
    UNet(*args, up_mode=0)
    for i, bad in ((5, "original_unet"), (6, 64), (11, "batch_norm"), (13, "leakyrelu"), (16, 3)):
        a = list(args)
        a[i] = bad
        with pytest.raises(NotImplementedError):
            UNet(*a, up_mode=0)
    with pytest.raises(NotImplementedError):
        UNet(*args, up_mode=1)
    with pytest.raises(AssertionError):
        common.get_layer_factor("no_such_operator")
    assert common.get_layer_factor("square_and_square_root") == 4
