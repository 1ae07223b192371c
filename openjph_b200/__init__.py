"""openjph_b200 -- a Blackwell-native HTJ2K encode/decode hot path behind OpenJPH's codestream
interface: HT block coder, lifting DWT (5/3, 9/7) and RCT/ICT as hand-written sm_100a CUDA
kernels in libojph_b200.so (C-ABI: include/ojph_b200.h), with this package as the thin
Python-side mirror of the reference's interface used by tests and bench.py."""
from .codestream import (Encoder, Decoder, OjphError, make_params, encode_blocks, decode_blocks,
                         U8, U16, I32)
