"""pytest plugin (`-p ref_dropin_plugin`) that runs the REFERENCE's own test files against the B200 drop-in: `src.models`
(and the negative sampler of `src.data`) are re-bound to this package before the test modules import them, and - because
those tests build their inputs with bare `torch.tensor(...)` and pass `device=torch.device("cpu")` - the default device is
CUDA and an explicit CPU device request is redirected to it (the package has no CPU path).  TEST INFRASTRUCTURE."""
import functools
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
for cand in (ROOT / "baseline" / "_ref", Path("/root/reference")):
    if (cand / "src" / "models").exists():
        sys.path.insert(0, str(cand))
        break
sys.path.insert(0, str(ROOT))

import two_tower_augmented_with_adaptive_mimic_mechanism_b200 as tt  # noqa: E402
from two_tower_augmented_with_adaptive_mimic_mechanism_b200 import sampler as tt_sampler  # noqa: E402

torch.set_default_device("cuda")


def _on_cuda(fn):
    @functools.wraps(fn)
    def wrapper(*a, **k):
        if "device" in k and (k["device"] is None or torch.device(k["device"]).type == "cpu"):
            k["device"] = torch.device("cuda")
        return fn(*a, **k)
    return wrapper


import src.data  # noqa: E402
import src.data.samplers  # noqa: E402
import src.models  # noqa: E402
import src.models.adaptive_mimic  # noqa: E402
import src.models.encoders  # noqa: E402
import src.models.two_tower  # noqa: E402

for mod in (src.models, src.models.encoders):
    for name in ("build_tower_encoder", "build_id_embedding"):
        if hasattr(mod, name):
            setattr(mod, name, _on_cuda(getattr(tt, name)))
    for name in ("TowerEncoder", "FeatureFusionGate"):
        if hasattr(mod, name):
            setattr(mod, name, getattr(tt, name))
for mod in (src.models, src.models.adaptive_mimic):
    mod.AdaptiveMimicMechanism = tt.AdaptiveMimicMechanism
for mod in (src.models, src.models.two_tower):
    mod.TwoTowerModel = tt.TwoTowerModel
for mod in (src.data, src.data.samplers):
    mod.sample_negative_items = _on_cuda(tt_sampler.sample_negative_items)
