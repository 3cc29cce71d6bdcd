"""libparanumal_b200: B200-native (sm_100a CUDA + NCCL) implementation of libParanumal's elliptic hot
path - hex Ax, ogs gather-scatter, PCG and the preconditioner applies - behind a C ABI
(include/libp_b200.h).  This package is the thin host-side mirror used by tests and bench.py."""
from . import _lib  # noqa: F401
from ._lib import (ADD, DOUBLE, FLOAT, HALO, INT32, INT64, MAX, MIN, MUL, NOTRANS, SIGNED, SYM, TRANS,  # noqa: F401
                   UNSIGNED, LibpError)
