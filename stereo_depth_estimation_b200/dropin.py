"""Level-B1 installer: make an installed copy of the reference bind the B200 model.

    import stereo_depth_estimation_b200.dropin as dropin; dropin.install()
    from foundation_stereo_depth import train; train.main()      # unmodified reference CLI

Patches ``foundation_stereo_depth.model`` (and any already-imported module that did
``from .model import StereoUNet``: train.py:21, live_camera/depth_live_dl.py:18) in
place; the reference package itself is not edited."""
from __future__ import annotations

import importlib
import sys

from .model import StereoUNet, load_state_dict_compat


def install(package: str = "foundation_stereo_depth") -> list:
    """Returns the names of the modules that were patched."""
    patched = []
    model_mod = importlib.import_module(f"{package}.model")
    model_mod.StereoUNet = StereoUNet
    model_mod.load_state_dict_compat = load_state_dict_compat
    patched.append(model_mod.__name__)
    for name, mod in list(sys.modules.items()):
        if mod is None or mod is model_mod:
            continue
        if name.startswith(package + ".") or name.startswith("live_camera."):
            for attr, new in (("StereoUNet", StereoUNet), ("load_state_dict_compat", load_state_dict_compat)):
                if hasattr(mod, attr):
                    setattr(mod, attr, new)
                    patched.append(f"{name}.{attr}")
    return patched
