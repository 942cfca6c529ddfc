"""sphy_b200 -- B200-native (sm_100a) implementation of SPHY's per-timestep raster hot path.

Host side is Python and mirrors the reference's `initial()/dynamic()` contract and module call
signatures; all raster arithmetic runs in hand-written CUDA kernels behind the C-ABI declared in
`include/sphy_b200.h` (ctypes).  PyTorch is used for device buffers, streams and torch.distributed only.
There is no CPU fallback: using the model without the CUDA library or without a GPU raises.
"""
__version__ = "0.1.0"
