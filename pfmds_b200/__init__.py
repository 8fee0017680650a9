"""pfmds_b200 — B200-native MD inner loop of PFMDS behind a C ABI (see include/pfmds_b200.h)."""
from .engine import Engine, PfmdsError, configure, load_library, LIB_PATH, NVE, NVT, NVMS  # noqa: F401
