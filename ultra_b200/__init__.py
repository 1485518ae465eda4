"""ultra_b200 - Blackwell-native (sm_100a) implementation of ULTRA's relational message-passing hot path.

Public surface (mirrors /root/reference/ultra/rspmm):
    from ultra_b200.rspmm import generalized_rspmm, RSPMM{Add,Min,Max}{Mul,Add}Function
The CUDA library behind it is ultra_b200/lib/libultra_rspmm.so (C ABI: include/ultra_rspmm.h), built in-tree
by `python -m ultra_b200.build`.  There is no CPU / PyTorch fallback.
"""
__version__ = "0.1.0"
