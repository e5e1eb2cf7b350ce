"""Oracle (test infrastructure, never imported by the product): import the UNMODIFIED reference package ``turbdiff``.

The reference is pure Python but imports packages that are absent from this image (pytorch_lightning, h5py, omegaconf,
more_itertools, lightning_utilities, torchmetrics, ot, deadpool).  None of them is on the denoising path, so they are
stubbed in ``sys.modules`` before the import.  The package itself is taken from

1. ``/root/reference`` when it exists (the build container), else
2. ``oracle/_ref/`` - the in-tree, git-ignored install made by ``oracle/install_ref.py`` (what
   ``pip install --target oracle/_ref /root/reference`` would produce; it travels to the GPU box with the snapshot).

Users: tests/golden/make_golden.py (fixture generation), bench.py's reference arms (``--impl reference``,
``--impl reference-gpu`` and the ``gpu_reference`` block), the drop-in tests.
"""

from __future__ import annotations

import importlib
import importlib.machinery
import sys
import types
import warnings
from pathlib import Path

HERE = Path(__file__).resolve().parent
CANDIDATES = [Path("/root/reference"), HERE / "_ref"]


def reference_root() -> Path | None:
    for c in CANDIDATES:
        if (c / "turbdiff" / "models" / "ddpm.py").is_file():
            return c
    return None


def available() -> bool:
    return reference_root() is not None


def _stub(name, pkg=False, **attrs):
    if name in sys.modules and not getattr(sys.modules[name], "__tdb_stub__", False):
        return sys.modules[name]  # the real package exists in this environment: use it
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__tdb_stub__ = True
    m.__spec__ = importlib.machinery.ModuleSpec(name, None, is_package=pkg)
    if pkg:
        m.__path__ = []
    sys.modules[name] = m
    return m


def _have(name) -> bool:
    try:
        importlib.import_module(name)
        return True
    except Exception:
        return False


def install_stubs():
    import torch

    class LightningModule(torch.nn.Module):
        """Stand-in for pl.LightningModule: an nn.Module with no-op logging and a `device` property."""

        trainer = None
        current_epoch = 0

        def log(self, *a, **k):
            self.__dict__.setdefault("logged", []).append((a, k))

        def log_dict(self, *a, **k):
            pass

        @property
        def device(self):
            return next(self.parameters()).device

    if not _have("h5py"):
        _stub("h5py", File=object, Group=object)
    if not _have("pytorch_lightning"):
        pl = _stub("pytorch_lightning", pkg=True, LightningModule=LightningModule, LightningDataModule=object, Callback=object, Trainer=object)
        _stub("pytorch_lightning.callbacks", ModelCheckpoint=object)
        _stub("pytorch_lightning.utilities", rank_zero_only=lambda f: f, move_data_to_device=lambda b, d: b)
        pl.loggers = _stub("pytorch_lightning.loggers", Logger=object)
    if not _have("lightning_utilities"):
        _stub("lightning_utilities", pkg=True)
        _stub("lightning_utilities.core", pkg=True)
        _stub("lightning_utilities.core.apply_func", apply_to_collection=lambda *a, **k: None)
    if not _have("more_itertools"):
        _stub("more_itertools", chunked=lambda it, n: [it[i : i + n] for i in range(0, len(it), n)])
    if not _have("omegaconf"):
        _stub("omegaconf", DictConfig=dict, OmegaConf=object)
    # only needed by turbdiff.models.metrics / .diffusion (the TKE-spectrum oracle and the drop-in test)
    if not _have("torchmetrics"):
        _stub("torchmetrics", Metric=torch.nn.Module, MetricCollection=torch.nn.ModuleDict)
    if not _have("ot"):
        _stub("ot", emd2=None)
    if not _have("deadpool"):
        _stub("deadpool", Deadpool=object)


_loaded = {}


def load(with_task: bool = False):
    """Import the reference.  Returns a namespace with ddpm, ofles, utils, Conditioning, CellTypeLearnedEmbedding,
    Normalization (+ metrics and DiffusionTraining with with_task=True)."""
    key = bool(with_task)
    if key in _loaded:
        return _loaded[key]
    root = reference_root()
    if root is None:
        raise RuntimeError("the reference package is neither at /root/reference nor installed under oracle/_ref "
                           "(run `python oracle/install_ref.py` in the build container)")
    install_stubs()
    if str(root) not in sys.path:
        sys.path.insert(0, str(root))
    warnings.filterwarnings("ignore", category=FutureWarning)
    ns = types.SimpleNamespace(root=root)
    ns.ddpm = importlib.import_module("turbdiff.models.ddpm")
    ns.ofles = importlib.import_module("turbdiff.data.ofles")
    ns.utils = importlib.import_module("turbdiff.models.utils")
    ns.Conditioning = importlib.import_module("turbdiff.models.conditioning").Conditioning
    ns.CellTypeLearnedEmbedding = importlib.import_module("turbdiff.models.cell_type_embeddings").CellTypeLearnedEmbedding
    ns.Normalization = importlib.import_module("turbdiff.models.normalization").Normalization
    if with_task:
        ns.metrics = importlib.import_module("turbdiff.models.metrics")
        ns.diffusion = importlib.import_module("turbdiff.models.diffusion")
    _loaded[key] = ns
    return ns
