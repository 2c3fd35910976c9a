"""B200-native DCLL hot path behind the Python API of ohjay/snn-modulation-classification.

Sub-packages mirror the reference's module layout:
  dcll.pytorch_libdcll  -- Conv2dDCLLlayer, DenseDCLLlayer, DCLLClassification ... (ref: dcll/pytorch_libdcll.py)
  networks              -- ConvNetwork, load_network_spec (ref: networks/__init__.py)
  data.utils            -- iq2spiketrain, to_one_hot (ref: data/utils.py)
All arithmetic runs in libdcll_b200.so (csrc/, C ABI in include/dcll_b200.h); importing this
package fails if that library has not been built -- there is no CPU fallback.
"""
from . import _lib  # noqa: F401  (fails loudly when libdcll_b200.so is missing)

__all__ = ["_lib"]
