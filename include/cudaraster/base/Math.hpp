// Small POD vector types that appear in the kept host and device API (Vec2i, Vec3i, Vec2f..Vec4f,
// Mat4f).  Stand-in for the reference's 1000-line template library
// (src/framework/base/Math.hpp); only the operations the pixel-pipe API and the stock shaders use.
#pragma once
#include "Defs.hpp"
#include <cmath>

namespace FW {

struct Vec2i {
    S32 x, y;
    FW_CUDA_FUNC Vec2i() : x(0), y(0) {}
    FW_CUDA_FUNC Vec2i(S32 a) : x(a), y(a) {}
    FW_CUDA_FUNC Vec2i(S32 xx, S32 yy) : x(xx), y(yy) {}
    FW_CUDA_FUNC bool operator==(const Vec2i& o) const { return x == o.x && y == o.y; }
    FW_CUDA_FUNC bool operator!=(const Vec2i& o) const { return !(*this == o); }
    FW_CUDA_FUNC Vec2i operator+(const Vec2i& o) const { return Vec2i(x + o.x, y + o.y); }
    FW_CUDA_FUNC Vec2i operator-(const Vec2i& o) const { return Vec2i(x - o.x, y - o.y); }
    FW_CUDA_FUNC Vec2i operator*(const Vec2i& o) const { return Vec2i(x * o.x, y * o.y); }
    FW_CUDA_FUNC Vec2i operator&(S32 m) const { return Vec2i(x & m, y & m); }
    FW_CUDA_FUNC Vec2i operator>>(int s) const { return Vec2i(x >> s, y >> s); }
    FW_CUDA_FUNC S32 min() const { return x < y ? x : y; }
    FW_CUDA_FUNC S32 max() const { return x > y ? x : y; }
};

struct Vec3i {
    S32 x, y, z;
    FW_CUDA_FUNC Vec3i() : x(0), y(0), z(0) {}
    FW_CUDA_FUNC Vec3i(S32 xx, S32 yy, S32 zz) : x(xx), y(yy), z(zz) {}
};

struct Vec2f {
    F32 x, y;
    FW_CUDA_FUNC Vec2f() : x(0), y(0) {}
    FW_CUDA_FUNC Vec2f(F32 a) : x(a), y(a) {}
    FW_CUDA_FUNC Vec2f(F32 xx, F32 yy) : x(xx), y(yy) {}
    FW_CUDA_FUNC Vec2f operator+(const Vec2f& o) const { return Vec2f(x + o.x, y + o.y); }
    FW_CUDA_FUNC Vec2f operator-(const Vec2f& o) const { return Vec2f(x - o.x, y - o.y); }
    FW_CUDA_FUNC Vec2f operator*(F32 s) const { return Vec2f(x * s, y * s); }
};

struct Vec3f {
    F32 x, y, z;
    FW_CUDA_FUNC Vec3f() : x(0), y(0), z(0) {}
    FW_CUDA_FUNC Vec3f(F32 a) : x(a), y(a), z(a) {}
    FW_CUDA_FUNC Vec3f(F32 xx, F32 yy, F32 zz) : x(xx), y(yy), z(zz) {}
    FW_CUDA_FUNC Vec3f operator+(const Vec3f& o) const { return Vec3f(x + o.x, y + o.y, z + o.z); }
    FW_CUDA_FUNC Vec3f operator-(const Vec3f& o) const { return Vec3f(x - o.x, y - o.y, z - o.z); }
    FW_CUDA_FUNC Vec3f operator*(F32 s) const { return Vec3f(x * s, y * s, z * s); }
    FW_CUDA_FUNC Vec3f operator-() const { return Vec3f(-x, -y, -z); }
};
FW_CUDA_FUNC Vec3f operator*(F32 s, const Vec3f& v) { return v * s; }

struct Vec4f {
    F32 x, y, z, w;
    FW_CUDA_FUNC Vec4f() : x(0), y(0), z(0), w(0) {}
    FW_CUDA_FUNC Vec4f(F32 a) : x(a), y(a), z(a), w(a) {}
    FW_CUDA_FUNC Vec4f(F32 xx, F32 yy, F32 zz, F32 ww) : x(xx), y(yy), z(zz), w(ww) {}
    FW_CUDA_FUNC Vec4f(const Vec3f& v, F32 ww) : x(v.x), y(v.y), z(v.z), w(ww) {}
    FW_CUDA_FUNC Vec4f(const Vec2f& v, F32 zz, F32 ww) : x(v.x), y(v.y), z(zz), w(ww) {}
    FW_CUDA_FUNC Vec2f getXY() const { return Vec2f(x, y); }
    FW_CUDA_FUNC Vec3f getXYZ() const { return Vec3f(x, y, z); }
    FW_CUDA_FUNC Vec4f operator+(const Vec4f& o) const { return Vec4f(x + o.x, y + o.y, z + o.z, w + o.w); }
    FW_CUDA_FUNC Vec4f operator-(const Vec4f& o) const { return Vec4f(x - o.x, y - o.y, z - o.z, w - o.w); }
    FW_CUDA_FUNC Vec4f operator*(F32 s) const { return Vec4f(x * s, y * s, z * s, w * s); }
    // Host-side packing with the reference's rounding (base/Math.cpp:41-48): round-half-up of
    // clamp(c,0,1)*255 per channel, R in the low byte.
    inline U32 toABGR() const {
        const F32 c[4] = {x, y, z, w};
        U32 r = 0;
        for (int i = 0; i < 4; i++) {
            F32 v = c[i] < 0.0f ? 0.0f : (c[i] > 1.0f ? 1.0f : c[i]);
            if (!(v == v)) v = 0.0f;
            U64 q = (U64)((F64)v * 72057594037927936.0) * 255u;
            r |= ((U32)((q >> 55) + 1) >> 1) << (8 * i);
        }
        return r;
    }
    static inline Vec4f fromABGR(U32 abgr) {
        return Vec4f((F32)(abgr & 0xFF) * (1.0f / 255.0f), (F32)((abgr >> 8) & 0xFF) * (1.0f / 255.0f),
                     (F32)((abgr >> 16) & 0xFF) * (1.0f / 255.0f), (F32)(abgr >> 24) * (1.0f / 255.0f));
    }
};

struct Mat4f {  // column-major like the reference: m[col][row]
    F32 m[4][4];
    FW_CUDA_FUNC Vec4f operator*(const Vec4f& v) const {
        return Vec4f(m[0][0] * v.x + m[1][0] * v.y + m[2][0] * v.z + m[3][0] * v.w,
                     m[0][1] * v.x + m[1][1] * v.y + m[2][1] * v.z + m[3][1] * v.w,
                     m[0][2] * v.x + m[1][2] * v.y + m[2][2] * v.z + m[3][2] * v.w,
                     m[0][3] * v.x + m[1][3] * v.y + m[2][3] * v.z + m[3][3] * v.w);
    }
};

FW_CUDA_FUNC F32 dot(const Vec3f& a, const Vec3f& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
FW_CUDA_FUNC Vec3f normalize(const Vec3f& v) { return v * (1.0f / sqrtf(dot(v, v))); }
FW_CUDA_FUNC F32 sqr(F32 a) { return a * a; }
FW_CUDA_FUNC Vec4f lerp(const Vec4f& a, const Vec4f& b, F32 t) { return a + (b - a) * t; }

}  // namespace FW
