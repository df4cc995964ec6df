"""vllm-kvcompress_b200: B200-native (sm_100a) KV-Compress eviction / compaction /
paged-attention hot path behind the reference's operator surface.

  native.py       locating + loading the in-tree libraries (fails loudly, no fallback)
  capi.py         ctypes view of the C ABI (include/kvc_b200.h) on torch device tensors
  _custom_ops.py  Python op shim with the reference's names and argument order
                  (vllm/_custom_ops.py:135-203, 641-658, 934-1256)
"""
__version__ = "0.1.0"
