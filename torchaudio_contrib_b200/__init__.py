"""torchaudio_contrib_b200 -- the STFT -> ComplexNorm -> ApplyFilterbank(mel) -> AmplitudeToDb and
mu-law path of keunwoochoi/torchaudio-contrib as hand-written sm_100a CUDA kernels behind a C ABI
(include/tac_b200.h), wrapped as drop-in functions and `nn.Module`s:

    import torchaudio_contrib_b200 as torchaudio_contrib

The package surface is the union of `.functional` and `.layers`, like the reference's
`torchaudio_contrib/__init__.py:1-2`.  Importing needs the built library
(`python build_native.py`); running needs a CUDA device.  No CPU fallback.
"""
from . import _cabi

_cabi.load()                       # fail loudly at import time if the CUDA library is missing

from .functional import *          # noqa: F401,F403,E402
from .layers import *              # noqa: F401,F403,E402
from .host import HostPipeline     # noqa: F401,E402

__version__ = "0.1.0"
