"""Test helper: import the UNMODIFIED reference from baseline/_ref (the copy `__graft_entry__.build()`
makes of /root/reference/{src,tests}; it travels to the GPU box, /root/reference does not) with a
recording stand-in for the one dependency that is not installed offline, `mlflow` (train.py:12).

Test infrastructure only - nothing in the product imports this."""
from __future__ import annotations

import contextlib
import importlib
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SRC = os.path.join(ROOT, "baseline", "_ref", "src")
REF_TESTS = os.path.join(ROOT, "baseline", "_ref", "tests")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_SRC, "foundation_stereo_depth", "train.py"))


class MlflowStub(types.ModuleType):
    """Records what train.py:365-403,580-599,638-681 logs."""

    def __init__(self) -> None:
        super().__init__("mlflow")
        self.metrics, self.params, self.artifacts, self.tags = [], {}, [], {}
        self._run = None

    def set_tracking_uri(self, uri):
        self.tracking_uri = uri

    def set_experiment(self, name):
        self.experiment = name

    @contextlib.contextmanager
    def start_run(self, run_name=None):
        self._run = types.SimpleNamespace(info=types.SimpleNamespace(run_id="stubrun0001"))
        try:
            yield self._run
        finally:
            self._run = None

    def active_run(self):
        return self._run

    def log_params(self, params):
        self.params.update(params)

    def log_metrics(self, metrics, step=None):
        self.metrics.append((step, dict(metrics)))

    def log_artifact(self, path, artifact_path=None):
        self.artifacts.append((str(path), artifact_path))

    def log_artifacts(self, path, artifact_path=None):
        self.artifacts.append((str(path), artifact_path))

    def set_tag(self, key, value):
        self.tags[key] = value


def load(fresh: bool = True):
    """Returns (train_module, model_module, mlflow_stub) of the vendored reference.  fresh=True drops
    previously imported copies first, so a test that patched the modules (dropin.install) cannot leak."""
    if not available():
        raise RuntimeError("baseline/_ref is missing: run __graft_entry__.build() where /root/reference exists")
    if fresh:
        for name in list(sys.modules):
            if name == "foundation_stereo_depth" or name.startswith("foundation_stereo_depth.") or \
                    name == "live_camera" or name.startswith("live_camera."):
                del sys.modules[name]
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    stub = MlflowStub()
    sys.modules["mlflow"] = stub
    train = importlib.import_module("foundation_stereo_depth.train")
    model = importlib.import_module("foundation_stereo_depth.model")
    train.mlflow = stub
    return train, model, stub


def load_live():
    """live_camera.depth_live_dl of the vendored reference (needs cv2, which the image has)."""
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    return importlib.import_module("live_camera.depth_live_dl")
