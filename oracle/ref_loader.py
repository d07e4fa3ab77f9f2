"""TEST INFRASTRUCTURE — import the UNMODIFIED reference package from
/root/reference with ``sys.modules`` stubs for the third-party packages that are
absent from this container.

Only works where /root/reference exists (the build container).  Used by
``oracle/make_golden.py`` to pin ``oracle/restate.py`` and to generate the
golden fixtures under ``tests/golden/``.  Never imported at run time on the GPU
box and never by the product package.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("SAM_AUDIO_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "sam_audio"))


def _stub(name: str, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _missing(what):
    class _Missing:  # placeholder type: instantiation means a test reached third-party code
        def __init__(self, *a, **k):
            raise RuntimeError(f"{what} is a stub (third-party package absent)")

        @classmethod
        def from_config(cls, *a, **k):
            raise RuntimeError(f"{what} is a stub (third-party package absent)")
    _Missing.__name__ = what
    return _Missing


def load():
    """Returns the imported ``sam_audio`` reference package."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    from oracle.restate import odeint_midpoint

    def odeint(func, y0, t, method="midpoint", options=None):
        assert method == "midpoint"
        n = round(float(t[-1] - t[0]) / options["step_size"])
        y1 = odeint_midpoint(func, y0, n)
        return [y0, y1]

    _stub("core")
    _stub("core.audio_visual_encoder", PEAudioFrame=_missing("PEAudioFrame"),
          PEAudioFrameTransform=_missing("PEAudioFrameTransform"))
    _stub("core.audio_visual_encoder.config", TransformerConfig=_missing("PEAVTransformerConfig"))
    _stub("core.audio_visual_encoder.transformer",
          BaseModelOutputWithPooling=_missing("BaseModelOutputWithPooling"),
          Transformer=_missing("Transformer"))
    _stub("core.vision_encoder")
    _stub("core.vision_encoder.pe", CLIP=_missing("CLIP"))
    _stub("torchdiffeq", odeint=odeint)
    _stub("dacvae", DACVAE=_missing("DACVAE"))
    _stub("torchcodec")
    _stub("torchcodec.decoders", AudioDecoder=_missing("AudioDecoder"), VideoDecoder=_missing("VideoDecoder"))
    _stub("torchcodec.encoders", AudioEncoder=_missing("AudioEncoder"))
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import sam_audio  # noqa: F401  (the reference)
    return sam_audio
