"""TEST INFRASTRUCTURE ONLY -- makes the *unmodified* reference tree importable in this container.

Used by oracle/make_golden.py (and nothing else) to run the reference's own Python code on CPU so
that (1) the restatement in oracle/restate.py can be validated against it and (2) golden vectors
can be written to tests/golden/.  /root/reference does not exist on the GPU box, so nothing under
tests/ -m gpu, bench.py or __graft_entry__.smoke() may import this module.

Recipe follows SURVEY.md section 8c:
  * the package directory is ``streammind/`` but every file imports ``videollama2.*``
    (e.g. /root/reference/streammind/model/multimodal_projector/builder.py:29), so the tree is
    registered under that name;
  * timm / lightning / pytorch_lightning / torchmetrics / decord / imageio / moviepy are needed at
    import time only (builder.py:22-28, ssm.py:7-9, mm_utils.py:7-12) and are stubbed;
  * the vendored ``streammind/model/mamba_ssm`` is exposed as top-level ``mamba_ssm``; its external
    CUDA extension ``selective_scan_cuda`` (ops/selective_scan_interface.py:16) is stubbed and
    ``selective_scan_fn`` is rebound to the in-tree pure-torch ``selective_scan_ref`` (:91-157);
  * ``mamba_ssm.models.mixer_seq_simple`` (mamba-ssm==2.2.2, requirements.txt:156; NOT vendored) is
    restated from its consumers: Block.__init__/forward (modules/block.py:11-88) and Mamba.__init__
    (modules/mamba_simple.py:31-117).
"""
import importlib
import importlib.util
import sys
import types
from functools import partial

REFERENCE_ROOT = "/root/reference"
_installed = False


def _stub(name, **attrs):
    mod = types.ModuleType(name)
    mod.__spec__ = importlib.machinery.ModuleSpec(name, loader=None)
    mod.__path__ = []
    for k, v in attrs.items():
        setattr(mod, k, v)
    sys.modules[name] = mod
    return mod


def install(reference_root: str = REFERENCE_ROOT):
    """Idempotently install all shims; returns the ``videollama2`` package module."""
    global _installed
    if _installed:
        return sys.modules["videollama2"]
    import importlib.machinery  # noqa: F401
    import torch
    import torch.nn as nn

    # 1. transformers first (its lazy loader replaces sys.modules['transformers'] on deep import).
    import transformers
    from transformers import CLIPVisionModel, MistralForCausalLM  # noqa: F401
    sys.modules["transformers"].__dict__["TRANSFORMERS_CACHE"] = "/tmp/hf_cache_unused"

    # 2. import-time-only third parties.
    class _LightningModule(nn.Module):
        def save_hyperparameters(self, *a, **k):
            pass

    _stub("timm"); _stub("timm.models")
    _stub("timm.models.regnet", RegStage=type("RegStage", (nn.Module,), {}))
    _stub("timm.models.layers", LayerNorm=nn.LayerNorm, LayerNorm2d=nn.LayerNorm)
    _stub("pytorch_lightning", LightningModule=_LightningModule)
    _stub("lightning", LightningModule=_LightningModule)
    _stub("lightning.pytorch")
    _stub("lightning.pytorch.callbacks", LearningRateMonitor=object)
    _stub("torchmetrics")
    _stub("torchmetrics.functional", accuracy=lambda *a, **k: None)
    _stub("decord", VideoReader=object, cpu=lambda *a, **k: None)
    _stub("imageio")
    _stub("moviepy")
    _stub("moviepy.editor", VideoFileClip=object)
    for name in ("cv2",):
        try:
            importlib.import_module(name)
        except Exception:
            _stub(name)

    # 3. vendored mamba_ssm as a top-level package + stubbed CUDA ext.
    _stub("selective_scan_cuda")
    base = f"{reference_root}/streammind/model/mamba_ssm"
    spec = importlib.util.spec_from_file_location(
        "mamba_ssm", f"{base}/__init__.py", submodule_search_locations=[base])
    pkg = types.ModuleType("mamba_ssm")       # do NOT exec its __init__ (it imports Mamba2/triton)
    pkg.__spec__ = spec
    pkg.__path__ = [base]
    sys.modules["mamba_ssm"] = pkg
    ssi = importlib.import_module("mamba_ssm.ops.selective_scan_interface")
    ssi.selective_scan_fn = ssi.selective_scan_ref
    ms = importlib.import_module("mamba_ssm.modules.mamba_simple")
    ms.selective_scan_fn = ssi.selective_scan_ref
    ms.selective_state_update = None
    blk = importlib.import_module("mamba_ssm.modules.block")

    # 4. restated mamba_ssm.models.mixer_seq_simple (2.2.2 defaults: Mamba-1, LayerNorm eps 1e-5,
    #    fused_add_norm=False, residual_in_fp32=False).
    def create_block(d_model, d_intermediate=0, ssm_cfg=None, attn_layer_idx=None, attn_cfg=None,
                     norm_epsilon=1e-5, rms_norm=False, residual_in_fp32=False,
                     fused_add_norm=False, layer_idx=None, device=None, dtype=None):
        assert d_intermediate == 0 and not rms_norm and not fused_add_norm
        mixer_cls = partial(ms.Mamba, layer_idx=layer_idx)
        norm_cls = partial(nn.LayerNorm, eps=norm_epsilon)
        block = blk.Block(d_model, mixer_cls, nn.Identity, norm_cls=norm_cls,
                          fused_add_norm=False, residual_in_fp32=residual_in_fp32)
        block.layer_idx = layer_idx
        return block

    def _init_weights(module, n_layer, **kw):   # initial values are overridden by the harness
        return None

    _stub("mamba_ssm.models")
    _stub("mamba_ssm.models.mixer_seq_simple", create_block=create_block, _init_weights=_init_weights)

    # 5. the reference tree under the name its own files use.
    root = f"{reference_root}/streammind"
    spec = importlib.util.spec_from_file_location(
        "videollama2", f"{root}/__init__.py", submodule_search_locations=[root])
    v = types.ModuleType("videollama2")       # skip the package __init__ (pulls in decord IO paths)
    v.__spec__ = spec
    v.__path__ = [root]
    sys.modules["videollama2"] = v
    _installed = True
    return v
