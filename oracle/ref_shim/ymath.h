/* Stand-in for MSVC's <ymath.h>: CudaThomas/Thomas.cu uses only `_Nan._Double` (in its disabled SaveThomas dump).
   Written for this repo; part of the oracle/_ref build recipe (oracle/Makefile), not reference code. */
#pragma once
static const struct { double _Double; } _Nan = { __builtin_nan("") };
