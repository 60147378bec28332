"""B200-native MPEG Layer II (MP2) DAB encode path -- Python binding of the C ABI (include/toolame_b200.h).

The product is the CUDA library `libtoolame_b200.so` built in-tree by `make` (or `__graft_entry__.build()`);
this module only loads it through ctypes.  There is no CPU path: a missing library raises ImportError-like
RuntimeError at first use, and a missing GPU makes `BatchEncoder(...)` raise.
"""
from .binding import (BatchEncoder, ToolameStream, encode_services, config_check, selftest_log10, TlbError, SIDE_DTYPE, lib, lib_path,  # noqa: F401
                      TAP_SB_SAMPLE, TAP_SCALAR_PRE, TAP_J_SCALE, TAP_SMR, TAP_SIDE)
