"""Network front end, engines and packing (API of the reference's deva/model/network.py)."""
# Modules this package does not provide (dataset readers, result savers, detectors, training code, ...) resolve to the
# reference checkout when one is on sys.path *after* this package: same-named package directories are chained,
# ours first (pkgutil.extend_path), so `evaluation/eval_vos.py` and `deva/ext/*` import unchanged.
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)
