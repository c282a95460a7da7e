// Minimal replacements for the reference's base/Defs.hpp typedefs and fail()
// (reference: src/framework/base/Defs.hpp:33-117).  Only what the kept API surface needs.
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

namespace FW {

typedef uint8_t U8;
typedef uint16_t U16;
typedef uint32_t U32;
typedef int8_t S8;
typedef int16_t S16;
typedef int32_t S32;
typedef float F32;
typedef double F64;
typedef unsigned long long U64;
typedef signed long long S64;

#define FW_U32_MAX (0xFFFFFFFFu)
#define FW_S32_MAX (0x7FFFFFFF)
#define FW_ARRAY_SIZE(X) (sizeof(X) / sizeof((X)[0]))

#ifdef __CUDACC__
#define FW_CUDA 1
#define FW_CUDA_FUNC __host__ __device__ __forceinline__
#else
#define FW_CUDA 0
#define FW_CUDA_FUNC inline
#endif

// Same convention as the reference: print the message and exit(EXIT_FAILURE).
inline void fail(const char* fmt, ...) {
    va_list args;
    va_start(args, fmt);
    vprintf(fmt, args);
    va_end(args);
    putchar('\n');
    exit(EXIT_FAILURE);
}

}  // namespace FW
