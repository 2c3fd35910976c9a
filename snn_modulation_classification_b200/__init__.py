"""B200-native DCLL hot path behind the Python API of ohjay/snn-modulation-classification.

Sub-packages mirror the reference's module layout:
  dcll.pytorch_libdcll  -- Conv2dDCLLlayer, DenseDCLLlayer, DCLLClassification ... (ref: dcll/pytorch_libdcll.py)
  networks              -- ConvNetwork, load_network_spec (ref: networks/__init__.py)
  data.utils            -- iq2spiketrain, to_one_hot (ref: data/utils.py)
All arithmetic runs in libdcll_b200.so (csrc/, C ABI in include/dcll_b200.h).  Every sub-module imports
``_lib``, which raises ImportError when that library has not been built
(``python -m snn_modulation_classification_b200.build``) -- there is no CPU fallback.  The package root
itself stays import-light so that the build module can run before the library exists.
"""
