"""B200-native wideband-TOA engine behind PulsePortraiture's fit API.

Only the extended-FFTFIT hot path is implemented: the batched
C-ABI library ``libppb200.so`` (hand-written sm_100a CUDA) plus a Python facade
that keeps the reference's call signatures:

    pulseportraiture_b200.pplib.fit_portrait        (pplib.py:2102)
    pulseportraiture_b200.pplib.fit_phase_shift     (pplib.py:2054)
    pulseportraiture_b200.pptoaslib.fit_portrait_full (pptoaslib.py:928)
    pulseportraiture_b200.pptoas.GetTOAs.get_TOAs   (pptoas.py:150)

There is no CPU fallback: importing works anywhere, computing needs the built
library and a CUDA device.
"""
__version__ = "0.1.0"
