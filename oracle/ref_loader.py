"""Import the staged, unmodified reference (``oracle/_ref``, see build_ref.py) - test infrastructure / baselines only.

Must run in a process that has NOT imported the product's ``deva`` package (both are named ``deva``): bench.py's
reference legs are their own processes.  The shims are the ones of SURVEY.md section 8(c): a stub ``pulp`` module
(absent here; only the automatic-consensus ILP needs it) and ``pretrained=False`` ResNets (no network for the
torchvision checkpoints - weights come from the product's seeded synthetic checkpoint instead).
"""
import importlib.util
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, '_ref')
ROOT = os.path.dirname(HERE)


def available() -> bool:
    return os.path.isfile(os.path.join(REF, 'deva', 'inference', 'inference_core.py'))


def load():
    """-> (DEVA class, DEVAInferenceCore class, synthetic_state_dict fn) of the reference."""
    if 'deva' in sys.modules and not getattr(sys.modules['deva'], '__file__', '').startswith(REF):
        raise RuntimeError('the product deva package is already imported in this process')
    if not available():
        raise RuntimeError('oracle/_ref is not staged: run python oracle/build_ref.py in the build container')
    sys.modules.setdefault('pulp', types.ModuleType('pulp'))
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import deva.model.resnet as R
    r18, r50 = R.resnet18, R.resnet50
    R.resnet18 = lambda pretrained=True, extra_dim=0: r18(pretrained=False, extra_dim=extra_dim)
    R.resnet50 = lambda pretrained=True, extra_dim=0: r50(pretrained=False, extra_dim=extra_dim)
    from deva.inference.inference_core import DEVAInferenceCore
    from deva.model.network import DEVA
    # the product's checkpoint synthesiser (pure python), loaded by path because the package names collide
    spec = importlib.util.spec_from_file_location(
        'b200_param_spec', os.path.join(ROOT, 'tracking-anything-with-deva_b200', 'deva', 'model', 'param_spec.py'))
    ps = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ps)
    return DEVA, DEVAInferenceCore, ps.synthetic_state_dict
