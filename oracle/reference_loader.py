"""Import the UNMODIFIED reference model (test / baseline infrastructure only; see oracle/__init__.py).

Looks for the reference's `core/models/ff-raft` folder under /root/reference (the build container) and then under
`baseline/_ref` (the verbatim copy `oracle/install_reference.py` makes, which travels to the GPU box).  Nothing here
restates reference code: it only puts the folder on sys.path the way the reference's own `train.py:16-21` expects and
instantiates `FF_RAFT_FUSION` the way `train.py` does.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CANDIDATES = [os.environ.get("FFCORR_REFERENCE", "/root/reference"), os.path.join(ROOT, "baseline", "_ref")]


def reference_root(model: str = "ff-raft") -> str | None:
    for base in CANDIDATES:
        p = os.path.join(base, "core", "models", model)
        if os.path.isfile(os.path.join(p, "common.py")):
            return p
    return None


def reference_kind() -> str:
    """'reference' when the unmodified code is importable here, else 'port'."""
    return "reference" if reference_root() else "port"


def load_ff_raft(config: str = "ffraft_chairs_orb.yaml", quiet: bool = True):
    """-> (FF_RAFT_FUSION instance with the reference's default init, cfg, module namespace dict)."""
    ref = reference_root("ff-raft")
    if ref is None:
        raise FileNotFoundError("the reference is neither at /root/reference nor installed under baseline/_ref "
                                "(run `python oracle/install_reference.py` where /root/reference exists)")
    if ref not in sys.path:
        sys.path.insert(0, ref)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from common import yaml_parser  # reference: common.py:32-42
        from FF_RAFT_Core.ff_raft import FF_RAFT_FUSION  # reference: ff_raft.py:75
        import FF_RAFT_Core.raft as ref_raft

        cfg = yaml_parser(os.path.join(ref, "config", "experiment", config))
        sink = io.StringIO() if quiet else sys.stdout
        with contextlib.redirect_stdout(sink):
            model = FF_RAFT_FUSION(use_fusion="parallel", fusion_channels=cfg.MODEL.FUSION_CHANNEL, raft_small=False,
                                   dropout=0.0, alternate_corr=False, abandon_fnet=False, fuse_cnet=True, cfg=cfg)
    return model, cfg, {"raft": ref_raft, "root": ref}


def load_ff_pwc(correlation_fn, config: str = "ffpwc_chairs.yaml", quiet: bool = True):
    """-> (unmodified FF_PWCNET instance, cfg).  The reference's model file does `from correlation import correlation`
    (ff_pwcnet.py:15-19) and that module needs CuPy (absent): a stand-in module named `correlation` is registered whose
    `correlation.FunctionCorrelation(tenOne=, tenTwo=)` is `correlation_fn` -- the caller passes either the reference's
    own kernels (oracle/pwc_ref_cuda.py) or the closed-form torch restatement of correlation.py:46-98.  Everything else
    (extractor, decoders, backwarp, refiner, preprocess) is the reference's code, untouched."""
    import types

    ref = reference_root("ff-pwcnet")
    if ref is None:
        raise FileNotFoundError("the FF-PWC reference is neither at /root/reference nor under baseline/_ref")
    # the ff-raft folder has modules of the same names (common, losses): the PWC folder must come first
    sys.path[:] = [p for p in sys.path if not p.rstrip("/").endswith("ff-raft")]
    for name in [m for m in sys.modules if m == "common" or m.startswith("losses")]:
        del sys.modules[name]
    sys.path.insert(0, ref)
    stub = types.ModuleType("correlation")
    stub.correlation = types.SimpleNamespace(FunctionCorrelation=correlation_fn)
    stub.FunctionCorrelation = correlation_fn
    sys.modules["correlation"] = stub
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from common import yaml_parser
        from PWCNet_Core.ff_pwcnet import FF_PWCNET

        cfg = yaml_parser(os.path.join(ref, "config", config))
        with contextlib.redirect_stdout(io.StringIO() if quiet else sys.stdout):
            model = FF_PWCNET(cfg)
    return model, cfg
