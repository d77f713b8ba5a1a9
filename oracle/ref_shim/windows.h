/* Stand-in for <windows.h>: CudaWrapper/Timer.h declares LARGE_INTEGER members, MatrixTranspose.cu uses LOWORD/HIWORD.
   Written for this repo; part of the oracle/_ref build recipe (oracle/Makefile), not reference code. */
#pragma once
typedef union { struct { unsigned int LowPart; int HighPart; }; long long QuadPart; } LARGE_INTEGER;
#define LOWORD(l) ((unsigned short)(((unsigned long)(l)) & 0xffff))
#define HIWORD(l) ((unsigned short)((((unsigned long)(l)) >> 16) & 0xffff))
