"""Drop-in for the reference's top-level `models` module (models.py:10-176): put this directory in front of the reference root on
sys.path and `import models as model` in steps/train_pa.py:5, steps/train_dpd.py:7 and steps/run_dpd.py:8 resolves to the native
CoreModel / CascadedModel — same constructor signature, attributes, parameter names and counts (so model ids and checkpoints are the
reference's), arithmetic in libodpd.so.  See shim/README.md."""
from opendpd_b200.models import CoreModel, CascadedModel, NATIVE_BACKBONES  # noqa: F401
