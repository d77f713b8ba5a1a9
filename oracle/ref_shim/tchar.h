/* Stand-in for the MSVC header the reference's CudaWrapper/stdafx.h includes (nothing from it is used).
   Written for this repo; part of the oracle/_ref build recipe (oracle/Makefile), not reference code. */
#pragma once
