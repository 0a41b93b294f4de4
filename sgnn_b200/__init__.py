"""sgnn_b200 -- B200-native sparse 3D convolution + generative upsampling engine behind SG-NN's
`sparseconvnet` operator surface (reference torch/model.py:7).

  sgnn_b200.scn      drop-in operator surface (InputLayer, SubmanifoldConvolution, Convolution, ...)
  sgnn_b200.GenModel host-side mirror of the reference generator (torch/model.py:276) with a fused forward
  sgnn_b200.engine   functional layer over the C ABI of libsgnn_b200.so (include/sgnn_b200.h)

Importing this package loads libsgnn_b200.so and fails loudly if it is missing: the product path has no
CPU / PyTorch fallback.
"""
import torch

from . import _lib            # noqa: F401  (raises ImportError when the CUDA library is absent)
from . import engine, scn     # noqa: F401
from .model import GenModel   # noqa: F401

# fp32 parity of the dense 8^3 U-Net: cuDNN / cuBLAS TF32 would perturb the first mask.
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
# run-to-run reproducibility of the library part (cuDNN may otherwise pick atomics-based algorithms)
torch.backends.cudnn.deterministic = True
torch.backends.cudnn.benchmark = False

__version__ = '0.1.0'
