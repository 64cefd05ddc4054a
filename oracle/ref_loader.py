"""TEST INFRASTRUCTURE -- imports the UNMODIFIED reference modules (from /root/reference in the build container, else from
the staged copy oracle/_ref made by oracle/build_ref.py) with the framework-only imports stubbed: pytorch_lightning,
ruamel.yaml, torchmetrics, matplotlib, seaborn (and wandb when absent) are not in this image and carry no arithmetic of
the path; every op that runs is the reference's own (src/models_multimodal.py, src/transformer_utils.py, src/loss.py).
Only tests/, tests/golden/*.py and bench.py's CPU legs may import this module; the product never does."""
import os
import sys
import types

import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))


def reference_root():
    """Directory that contains the reference's `src` package, or None."""
    for cand in (os.environ.get("MAVEN_REFERENCE"), "/root/reference", os.path.join(HERE, "_ref")):
        if cand and os.path.isfile(os.path.join(cand, "src", "models_multimodal.py")):
            return cand
    return None


def stub_framework_modules():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    class _LM(nn.Module):                      # LightningModule's only use on this path: nn.Module + self.log
        def log(self, *a, **k):
            pass

    if "pytorch_lightning" not in sys.modules:
        mod("pytorch_lightning", LightningModule=_LM, Callback=object, Trainer=object)
        mod("pytorch_lightning.callbacks", Callback=object)
    if "ruamel.yaml" not in sys.modules:
        mod("ruamel"); mod("ruamel.yaml", YAML=object)
    if "torchmetrics" not in sys.modules:
        mod("torchmetrics"); mod("torchmetrics.classification", MulticlassFBetaScore=object)
    if "matplotlib" not in sys.modules:
        mp = mod("matplotlib"); mod("matplotlib.pyplot"); mod("matplotlib.ticker", MaxNLocator=object)
        mp.pyplot = sys.modules["matplotlib.pyplot"]
        mp.ticker = sys.modules["matplotlib.ticker"]
    if "seaborn" not in sys.modules:
        mod("seaborn")
    if "wandb" not in sys.modules:
        try:
            import wandb  # noqa: F401
        except Exception:
            mod("wandb")


def import_reference():
    """-> (src.transformer_utils, src.loss, src.models_multimodal) of the unmodified reference."""
    root = reference_root()
    if root is None:
        raise ImportError("no reference tree: neither /root/reference nor oracle/_ref (run `python oracle/build_ref.py` in the build container)")
    stub_framework_modules()
    if root not in sys.path:
        sys.path.insert(0, root)
    from src import loss as rloss
    from src import models_multimodal as rmm
    from src import transformer_utils as rtu
    return rtu, rloss, rmm
