// Prelude that lets the UNMODIFIED reference device sources
// (/root/reference/src/cudaraster/cuda/*.inl) compile with nvcc 12.9 for sm_100a.
//
// TEST INFRASTRUCTURE ONLY. Nothing in the product path includes this file.
//
// The reference (PixelPipe.inl:32-37) uses CUDA texture<>/surface<> *references*
// (removed in CUDA 12) and legacy warp votes (__ballot/__any/__all, rejected for
// sm_70+).  This shim maps them onto plain pointers and *_sync votes:
//   texture<T,1>      -> struct { const T* ptr }        tex1Dfetch -> __ldg
//   surface<void,2>   -> struct { u8* ptr; u32 pitch }  surf2Dread/write -> pitched pointer
//   __ballot/__any/__all(p) -> *_sync(__activemask(), p)
// __activemask() is weaker than Fermi's lock-step guarantee;
// the harness therefore treats this oracle as best effort.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

template <class T, int D> struct cr_texture { const T* ptr; };
template <class T, int D> struct cr_surface { unsigned char* ptr; unsigned int pitch; };

#define texture __device__ cr_texture
#define surface __device__ cr_surface

template <class T>
__device__ __forceinline__ T tex1Dfetch(const cr_texture<T, 1>& t, int i) { return __ldg(t.ptr + i); }

template <class T>
__device__ __forceinline__ T surf2Dread(const cr_surface<void, 2>& s, int xBytes, int y)
{ return *(const volatile T*)(s.ptr + (size_t)y * s.pitch + xBytes); }

template <class T>
__device__ __forceinline__ void surf2Dwrite(T v, const cr_surface<void, 2>& s, int xBytes, int y)
{ *(volatile T*)(s.ptr + (size_t)y * s.pitch + xBytes) = v; }

#define __ballot(p) __ballot_sync(__activemask(), (p))
#define __any(p)    __any_sync(__activemask(), (p))
#define __all(p)    __all_sync(__activemask(), (p))
