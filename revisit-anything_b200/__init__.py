"""segvlad-b200: B200-native SegVLAD retrieval engine (aggregate -> match -> vote hot path of
AnyLoc/Revisit-Anything) behind the reference's `func_vpr` / `place_rec_main` call surface.

Layout: `csrc/` = hand-written sm_100a CUDA kernels + the C-ABI (`include/segvlad.h`);
`_lib.py` = ctypes binding (fails loudly when the shared library is missing); `engine.py` = device
resident API; `func_vpr.py` / `place_rec_main.py` = drop-in mirrors of the reference functions on the
path; `distributed.py` = row-sharded bank + single all-gather; `synth.py` = seeded synthetic inputs.
"""
__version__ = "0.1.0"
