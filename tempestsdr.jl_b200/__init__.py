"""tempestsdr.jl_b200 -- B200-native raw-IQ -> image DSP chain (host side).

Import it as `tempestsdr_b200` (see ../tempestsdr_b200/__init__.py).  The
compute lives in libtempest_b200.so (csrc/, hand-written sm_100a CUDA behind
the C ABI of include/tempest_b200.h); this package is the Python mirror of the
reference's exported Julia functions, bound with ctypes.
"""
