"""Record the reference's public surface for the hot path as a JSON fixture (tests/golden/reference_surface.json):
function names + positional parameter names of activate_trained_model/test_imageTMO.py and test_videoTMO.py, their CLI
flags and defaults, and the constructor / forward signatures of the generator, discriminator and StructLoss classes.
Run in the build container only (imports /root/reference in place):   python tests/golden/make_golden_surface.py
"""
import inspect
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shims  # noqa: E402

R = ref_shims.reference_modules()
sys.argv = sys.argv[:1]
sys.modules.setdefault("tranforms", type(sys)("tranforms"))
sys.path.insert(0, os.path.join(ref_shims.REF, "activate_trained_model"))
import importlib  # noqa: E402


def params(f):
    return [p.name for p in inspect.signature(f).parameters.values()]


def cli(mod):
    sys.argv = ["x"]
    a = mod.get_args()
    return {k: v for k, v in vars(a).items()}


out = {}
for name in ("test_imageTMO", "test_videoTMO"):
    m = importlib.import_module(name)
    fns = {k: params(v) for k, v in vars(m).items() if inspect.isfunction(v) and v.__module__ == m.__name__}
    out[name] = {"functions": fns, "cli_defaults": cli(m), "extensions": list(m.extensions)}
out["classes"] = {
    "UNet_image.__init__": params(R.gen_img.UNet.__init__), "UNet_image.forward": params(R.gen_img.UNet.forward),
    "UNet_video.__init__": params(R.gen_vid.UNet.__init__), "UNet_video.forward": params(R.gen_vid.UNet.forward),
    "SimpleDiscriminator.__init__": params(R.disc.SimpleDiscriminator.__init__),
    "SimpleDiscriminator.forward": params(R.disc.SimpleDiscriminator.forward),
    "StructLoss.__init__": params(R.struct_loss.StructLoss.__init__), "StructLoss.forward": params(R.struct_loss.StructLoss.forward),
}
import GanTrainerImg  # noqa: E402
out["trainer_methods"] = {k: params(getattr(GanTrainerImg.GanTrainer, k)) for k in
                          ("contrastive_D_loss", "nce", "infoNCE", "infoNCE2", "pseudo_label_loss", "train_D", "train_G")}
path = os.path.join(HERE, "reference_surface.json")
json.dump(out, open(path, "w"), indent=1, sort_keys=True)
print("wrote", path)
