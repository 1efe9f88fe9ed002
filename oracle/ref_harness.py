"""TEST INFRASTRUCTURE — loads the *unmodified* reference from /root/reference.

Only usable in the authoring container (``/root/reference`` does not exist on the
GPU box).  It is used by ``oracle/make_golden.py`` to run the reference's own
``compute_moco_contrastive`` (vilt/modules/objectives.py:217-447) and
``PGDAttack_moco.pgd_attack`` (attack/pgd_attack_vilt.py:130-175) on CPU and dump
golden vectors into ``tests/golden/``, and by ``tests/test_oracle_vs_reference.py``
(skipped when the reference tree is absent) to pin ``oracle/rmcl_oracle.py``.

The reference depends on packages this image does not have (pytorch_lightning,
timm, sacred, nltk, sentence_transformers, matplotlib).  None of them does
arithmetic on the hot path, so they are replaced by inert stand-ins in
``sys.modules`` *after* ``import transformers`` (transformers probes for timm
with ``find_spec`` and chokes on spec-less stand-ins).
"""
import importlib.machinery
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("RMCL_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "vilt", "modules", "objectives.py"))


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, None)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _install_stand_ins():
    import torch
    import torch.nn as nn
    import transformers  # noqa: F401  (must come first, see module docstring)
    import transformers.optimization as t_opt

    if not hasattr(t_opt, "AdamW"):
        t_opt.AdamW = torch.optim.AdamW

    class _HParams(dict):
        __getattr__ = dict.__getitem__

    class LightningModule(nn.Module):
        """nn.Module + the three PL conveniences vilt_module.py touches."""

        def save_hyperparameters(self):
            import inspect
            frame = inspect.currentframe().f_back
            self.hparams = _HParams(config=frame.f_locals["config"])

        def log(self, name, value, *a, **k):
            self.__dict__.setdefault("_logged", {})[name] = value

        @property
        def device(self):
            return next(self.parameters()).device

    class LightningDataModule:
        def __init__(self, *a, **k):
            pass

    class Metric(nn.Module):
        def __init__(self, dist_sync_on_step=False):
            super().__init__()
            self._defaults = {}

        def add_state(self, name, default, dist_reduce_fx=None):
            self._defaults[name] = default
            setattr(self, name, default.clone())

        def forward(self, *a, **k):
            self.update(*a, **k)
            return self.compute()

        def reset(self):
            for k, v in self._defaults.items():
                setattr(self, k, v.clone())

    def seed_everything(seed):
        torch.manual_seed(seed)

    pl = _mod("pytorch_lightning", LightningModule=LightningModule,
              LightningDataModule=LightningDataModule, seed_everything=seed_everything)
    pl.metrics = _mod("pytorch_lightning.metrics", Metric=Metric)

    class DropPath(nn.Module):
        def __init__(self, p=0.0):
            super().__init__()
            self.p = p

        def forward(self, x):
            assert self.p == 0.0 or not self.training, "parity runs use drop_path 0"
            return x

    _mod("timm")
    _mod("timm.data", IMAGENET_DEFAULT_MEAN=(0.485, 0.456, 0.406),
         IMAGENET_DEFAULT_STD=(0.229, 0.224, 0.225))
    _mod("timm.models")
    _mod("timm.models.helpers", load_pretrained=lambda *a, **k: None)
    _mod("timm.models.layers", DropPath=DropPath, StdConv2dSame=nn.Conv2d,
         to_2tuple=lambda x: x if isinstance(x, tuple) else (x, x),
         trunc_normal_=nn.init.trunc_normal_)
    _mod("timm.models.resnet", resnet26d=None, resnet50d=None)
    _mod("timm.models.resnetv2", ResNetV2=None)
    _mod("timm.models.registry", register_model=lambda f: f)

    class Experiment:
        def __init__(self, *a, **k):
            pass

        def _passthrough(self, f=None, *a, **k):
            return f

        config = named_config = automain = main = capture = _passthrough

    _mod("sacred", Experiment=Experiment)
    nltk = _mod("nltk")
    nltk.corpus = _mod("nltk.corpus", stopwords=None, wordnet=None)
    _mod("sentence_transformers", SentenceTransformer=None, util=None)
    plt = _mod("matplotlib.pyplot", rc=lambda *a, **k: None)
    cm = _mod("matplotlib.cm")
    _mod("matplotlib", pyplot=plt, cm=cm, rc=lambda *a, **k: None)
    _mod("ipdb")
    _mod("gradio")


_loaded = {}


def load_reference():
    """Returns a namespace with the reference's hot-path modules (unmodified)."""
    if _loaded:
        return types.SimpleNamespace(**_loaded)
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    _install_stand_ins()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import attack.pgd_attack_vilt as pgd_mod
    import vilt.modules.objectives as objectives
    import vilt.modules.heads as heads
    _loaded.update(pgd=pgd_mod, objectives=objectives, heads=heads)
    return types.SimpleNamespace(**_loaded)


def load_reference_model_module():
    """The full LightningModule (ViLTransformerSS); heavier import."""
    load_reference()
    import vilt.modules.vilt_module as vilt_module
    import vilt.modules.vilt_utils as vilt_utils
    return vilt_module, vilt_utils


def ensure_process_group():
    """objectives.py:231 calls all_gather unconditionally → 1-rank gloo group."""
    import tempfile
    import torch.distributed as dist
    if not dist.is_initialized():
        f = tempfile.NamedTemporaryFile(prefix="rmcl_pg_", delete=False)
        f.close()
        os.unlink(f.name)
        dist.init_process_group("gloo", init_method=f"file://{f.name}", rank=0, world_size=1)
