"""pai_b200 -- host-side (Python/PyTorch) binding of the B200-native hot path.

``lib``      ctypes loader of the C-ABI library (csrc/ -> libpai_b200.so, include/pai_b200.h)
``ops``      thin tensor-level wrappers + weight packing
``metrics``  SSIM / PSNR / RMSE (models/utils.py:38-47, report.py:72-104,188-217)
"""
