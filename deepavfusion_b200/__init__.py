"""deepavfusion_b200 -- B200-native (sm_100a) implementation of the DeepAVFusion pre-training hot
path behind the reference's own class API (``DeepAVFusion`` / ``AVMAE`` constructors, forward
signatures and ``state_dict`` layout).  All arithmetic runs in hand-written CUDA
kernels (``csrc/``, C ABI in ``include/davf.h``); there is no CPU or library fallback."""
__version__ = "0.1.0"
