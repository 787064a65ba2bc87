"""gapless_lossy_codec_b200 -- B200-native (sm_100a) encode/decode hot path of
ajcm474/gapless-lossy-codec behind the reference's own `codec` / `flac` API.

Layout (only what the hot path needs):
    csrc/          CUDA kernels + the C ABI (libglc_b200.so, declared in include/glc.h)
    _ffi.py        ctypes binding of that ABI
    codec.py       mirror of the reference's codec module (Encoder, Decoder, EncodedAudio, ...)
    flac.py        mirror of the reference's flac module (encode_flac_with_level, ...)

Like `pub use codec::*` in the reference's src/lib.rs, the codec items are re-exported here.
There is no CPU fallback: without libglc_b200.so and a CUDA device every call raises GlcError.
"""
from ._ffi import GlcError, LIB_PATH  # noqa: F401
from .codec import (  # noqa: F401
    FRAME_SIZE, HOP_SIZE, FRAMES_PER_CHUNK, AudioChunk, AudioHeader, Context, Decoder, EncodedAudio,
    EncodedFrame, Encoder, GaplessInfo, Progress, default_context, encoded_from_bytes,
    encoded_to_bytes, load_encoded, save_encoded,
)
from . import codec, flac  # noqa: F401
