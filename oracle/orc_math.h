// ORACLE — test infrastructure only (CPU restatement of the reference algorithms).
// Never linked, imported or executed by the product path (libvpt.so); only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg use it.
//
// Math layer. Follows /root/reference/renderer/shaders/LinearMath.h:
//   FMA/dop/TwoProd/TwoSum/InnerProduct   :80-146
//   normalize (1e-8 guard -> (0,0,1))      :962-973
//   cross via dop                          :980
//   dot (compensated, 3 terms)             :1017
//   Mat3 (column storage), Mat3*v          :1040-1108
//   Quat, rotate, rotationBetween          :1311-1366
//   luminance                              :1582-1586
//   ConcentricSampleDisk                   :1663-1690
//   YawPitchToDir / DirToYawPitch          :1692-1740
//   alignVector                            :1806-1814
//   EqualAreaSphereMap / EqualAreaMapCone  :1858-1913
//   LocalizeSample                         :1449-1461, refract :1483-1512
// All arithmetic is fp32, compiled with -ffp-contract=off; FMA only where the
// reference writes FMA explicitly. M_PI in the reference is a float literal
// (LinearMath.h:17, MSVC has no M_PI without _USE_MATH_DEFINES) so it is float here.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <cfloat>

namespace orc {

constexpr float kPi       = 3.1415926535897932384626422832795028841971f;
constexpr float kTwoPi    = 6.2831853071795864769252867665590057683943f;
constexpr float kPiOver2  = 1.5707963267948966192313216916397514420985f;
constexpr float kPiOver4  = 0.7853981633974483096156608458198757210492f;
constexpr float kInvTwoPi = 0.15915494309f;
constexpr float kPiOver180 = 0.01745329251f;
constexpr float kSafeCosEps = 1e-5f;
constexpr float kRayMax = 1.0e27f;

struct f2 { float x, y; };
struct f3 { float x, y, z; };
struct f4 { float x, y, z, w; };
struct i2 { int x, y; };

inline f3 F3(float a) { return {a, a, a}; }
inline f3 F3(float x, float y, float z) { return {x, y, z}; }
inline f4 F4(float a) { return {a, a, a, a}; }
inline f4 F4(f3 v, float w) { return {v.x, v.y, v.z, w}; }
inline f3 xyz(f4 v) { return {v.x, v.y, v.z}; }

inline f3 operator+(f3 a, f3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline f3 operator-(f3 a, f3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline f3 operator*(f3 a, f3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline f3 operator/(f3 a, f3 b) { return {a.x / b.x, a.y / b.y, a.z / b.z}; }
inline f3 operator*(f3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline f3 operator*(float s, f3 a) { return {a.x * s, a.y * s, a.z * s}; }
inline f3 operator/(f3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
inline f3 operator-(f3 a) { return {-a.x, -a.y, -a.z}; }
inline f3 &operator+=(f3 &a, f3 b) { a = a + b; return a; }
inline f3 &operator*=(f3 &a, f3 b) { a = a * b; return a; }
inline f3 &operator*=(f3 &a, float s) { a = a * s; return a; }
inline f3 &operator/=(f3 &a, float s) { a = a / s; return a; }

inline f4 operator+(f4 a, f4 b) { return {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
inline f4 operator-(f4 a, f4 b) { return {a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w}; }
inline f4 operator*(f4 a, f4 b) { return {a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w}; }
inline f4 operator/(f4 a, f4 b) { return {a.x / b.x, a.y / b.y, a.z / b.z, a.w / b.w}; }
inline f4 operator*(f4 a, float s) { return {a.x * s, a.y * s, a.z * s, a.w * s}; }
inline f4 operator*(float s, f4 a) { return {a.x * s, a.y * s, a.z * s, a.w * s}; }
inline f4 operator/(f4 a, float s) { return {a.x / s, a.y / s, a.z / s, a.w / s}; }
inline f4 &operator+=(f4 &a, f4 b) { a = a + b; return a; }

inline f2 operator+(f2 a, f2 b) { return {a.x + b.x, a.y + b.y}; }
inline f2 operator-(f2 a, f2 b) { return {a.x - b.x, a.y - b.y}; }
inline f2 operator*(f2 a, f2 b) { return {a.x * b.x, a.y * b.y}; }
inline f2 operator*(f2 a, float s) { return {a.x * s, a.y * s}; }
inline f2 operator*(float s, f2 a) { return {a.x * s, a.y * s}; }

inline float fminr(float a, float b) { return a < b ? a : b; }   // template min   (LinearMath.h:72)
inline float fmaxr(float a, float b) { return a > b ? a : b; }   // template max   (LinearMath.h:69)
inline float max1f(float a, float b) { return (a < b) ? b : a; } // LinearMath.h:423
inline float clampf(float a, float lo = 0.0f, float hi = 1.0f) { return a < lo ? lo : a > hi ? hi : a; }
inline int clampi(int a, int lo, int hi) { return a < lo ? lo : (a > hi ? hi : a); }
inline float saturate(float x) { return fminf(fmaxf(x, 0.0f), 1.0f); }
inline float lerpf(float a, float b, float w) { return a + w * (b - a); }
inline f3 lerp3(f3 a, f3 b, float w) { return a + w * (b - a); }
inline f4 lerp4(f4 a, f4 b, float w) { return a + w * (b - a); }
inline f3 max3f(f3 a, f3 b) { return {fmaxr(a.x, b.x), fmaxr(a.y, b.y), fmaxr(a.z, b.z)}; }
inline f4 max4f(f4 a, f4 b) { return {fmaxr(a.x, b.x), fmaxr(a.y, b.y), fmaxr(a.z, b.z), fmaxr(a.w, b.w)}; }
inline f3 abs3(f3 v) { return {fabsf(v.x), fabsf(v.y), fabsf(v.z)}; }
inline f3 clamp3(f3 a, f3 lo, f3 hi) { return {clampf(a.x, lo.x, hi.x), clampf(a.y, lo.y, hi.y), clampf(a.z, lo.z, hi.z)}; }
inline f3 sqrt3(f3 v) { return {sqrtf(v.x), sqrtf(v.y), sqrtf(v.z)}; }
inline float pow5(float e) { float e2 = e * e; return e2 * e2 * e; }
inline bool isNull(f3 v) { return v.x == 0.0f && v.y == 0.0f && v.z == 0.0f; }

// ---- compensated arithmetic (LinearMath.h:80-146) ----
inline float dop(float a, float b, float c, float d)
{
    float cd = c * d;
    float err = fmaf(-c, d, cd);
    float r = fmaf(a, b, -cd);
    return r + err;
}
struct cfloat { float v, err; };
inline cfloat twoProd(float a, float b) { float ab = a * b; return {ab, fmaf(a, b, -ab)}; }
inline cfloat twoSum(float a, float b)
{
    float s = a + b, delta = s - a;
    return {s, (a - (s - delta)) + (b - delta)};
}
inline float inner3(float a0, float b0, float a1, float b1, float a2, float b2)
{
    cfloat p0 = twoProd(a0, b0);
    cfloat p1 = twoProd(a1, b1);
    cfloat p2 = twoProd(a2, b2);
    cfloat s12 = twoSum(p1.v, p2.v);
    cfloat tp = {s12.v, p1.err + (p2.err + s12.err)};
    cfloat s = twoSum(p0.v, tp.v);
    cfloat r = {s.v, p0.err + (tp.err + s.err)};
    return r.v + r.err;
}
inline float dot(f3 a, f3 b) { return inner3(a.x, b.x, a.y, b.y, a.z, b.z); }
inline float dot4(f4 a, f4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
inline f3 cross(f3 a, f3 b)
{
    return {dop(a.y, b.z, a.z, b.y), dop(a.z, b.x, a.x, b.z), dop(a.x, b.y, a.y, b.x)};
}
inline float length(f3 v) { return sqrtf(dot(v, v)); }
inline float length2(f3 v) { return dot(v, v); } // Float3::length2() is compensated (LinearMath.h:536)
inline float distance(f3 a, f3 b)
{
    return sqrtf((a.x - b.x) * (a.x - b.x) + (a.y - b.y) * (a.y - b.y) + (a.z - b.z) * (a.z - b.z));
}
inline f3 normalize(f3 v)
{
    float norm = sqrtf(v.x * v.x + v.y * v.y + v.z * v.z);
    if (norm < 1e-8f || std::isnan(norm)) return {0.0f, 0.0f, 1.0f};
    return {v.x / norm, v.y / norm, v.z / norm};
}
inline float luminance(f3 c) { return dot(c, F3(0.2126f, 0.7152f, 0.0722f)); }
inline f3 reflect3(f3 i, f3 n) { return i - 2.0f * n * dot(n, i); }

// ---- Mat3: column storage m00,m10,m20 | m01,m11,m21 | m02,m12,m22 (LinearMath.h:1040-1108) ----
struct mat3 { float m00, m10, m20, m01, m11, m21, m02, m12, m22; };
inline mat3 mat3Zero() { mat3 m; std::memset(&m, 0, sizeof m); return m; }
inline mat3 mat3Cols(f3 c0, f3 c1, f3 c2) { return {c0.x, c0.y, c0.z, c1.x, c1.y, c1.z, c2.x, c2.y, c2.z}; }
inline mat3 transpose(mat3 m)
{
    mat3 r = m;
    r.m01 = m.m10; r.m10 = m.m01; r.m02 = m.m20; r.m20 = m.m02; r.m12 = m.m21; r.m21 = m.m12;
    return r;
}
inline mat3 mul(const mat3 &A, const mat3 &B)
{
    mat3 C;
    C.m00 = A.m00 * B.m00 + A.m01 * B.m10 + A.m02 * B.m20;
    C.m01 = A.m00 * B.m01 + A.m01 * B.m11 + A.m02 * B.m21;
    C.m02 = A.m00 * B.m02 + A.m01 * B.m12 + A.m02 * B.m22;
    C.m10 = A.m10 * B.m00 + A.m11 * B.m10 + A.m12 * B.m20;
    C.m11 = A.m10 * B.m01 + A.m11 * B.m11 + A.m12 * B.m21;
    C.m12 = A.m10 * B.m02 + A.m11 * B.m12 + A.m12 * B.m22;
    C.m20 = A.m20 * B.m00 + A.m21 * B.m10 + A.m22 * B.m20;
    C.m21 = A.m20 * B.m01 + A.m21 * B.m11 + A.m22 * B.m21;
    C.m22 = A.m20 * B.m02 + A.m21 * B.m12 + A.m22 * B.m22;
    return C;
}
inline f3 mul(const mat3 &m, f3 v)
{
    return {inner3(m.m00, v.x, m.m01, v.y, m.m02, v.z),
            inner3(m.m10, v.x, m.m11, v.y, m.m12, v.z),
            inner3(m.m20, v.x, m.m21, v.y, m.m22, v.z)};
}

// ---- Quat (LinearMath.h:1311-1366) ----
struct quat { f3 v; float w; };
inline quat qmul(quat p, quat q)
{
    return {p.w * q.v + q.w * p.v + cross(p.v, q.v), p.w * q.w - dot(p.v, q.v)};
}
inline quat qconj(quat q) { return {-q.v, q.w}; }
inline quat qnormalized(quat q)
{
    float n = sqrtf(q.v.x * q.v.x + q.v.y * q.v.y + q.v.z * q.v.z + q.w * q.w);
    return {q.v / n, q.w / n};
}
inline quat rotationBetween(f3 p, f3 q)
{
    return qnormalized({cross(p, q), sqrtf(length2(p) * length2(q)) + dot(p, q)});
}
inline f3 qrotate(quat q, f3 v) { return qmul(qmul(q, quat{v, 0.0f}), qconj(q)).v; }

// ---- sampling helpers ----
inline void alignVector(f3 axis, f3 &w)
{
    const float s = copysignf(1.0f, axis.z);
    w.z *= s;
    const f3 h = {axis.x, axis.y, axis.z + s};
    const float k = dot(w, h) / (1.0f + fabsf(axis.z));
    w = k * h - w;
}
inline void localizeSample(f3 n, f3 &u, f3 &v)
{
    f3 w = {1, 0, 0};
    if (fabsf(n.x) > 0.707f) w = {0, 1, 0};
    u = cross(n, w);
    v = cross(n, u);
}
inline f3 equalAreaSphereMap(float u, float v)
{
    float y = 2.0f * v - 1.0f;
    float r = sqrtf(1.0f - y * y);
    float phi = kTwoPi * u;
    return {r * cosf(phi), y, r * sinf(phi)};
}
inline f2 equalAreaSphereMapInv(f3 dir)
{
    float u = atan2f(-dir.z, -dir.x) / kTwoPi + 0.5f;
    float v = (dir.y + 1.0f) * 0.5f;
    return {u, v};
}
inline f3 equalAreaMapCone(f3 sunDir, float u, float v, float cosThetaMax)
{
    float cosTheta = (1.0f - u) + u * cosThetaMax;
    float sinTheta = sqrtf(1.0f - cosTheta * cosTheta);
    float phi = v * kTwoPi;
    f3 t, b;
    localizeSample(sunDir, t, b);
    mat3 trans = mat3Cols(t, sunDir, b);
    f3 coords = {cosf(phi) * sinTheta, cosTheta, sinf(phi) * sinTheta};
    return mul(trans, coords);
}
inline bool equalAreaMapConeInv(f2 &uv, f3 sunDir, f3 rayDir, float cosThetaMax)
{
    f3 t, b;
    localizeSample(sunDir, t, b);
    mat3 trans = transpose(mat3Cols(t, sunDir, b));
    f3 coords = mul(trans, rayDir);
    float cosTheta = coords.y;
    if (cosTheta < cosThetaMax) return false;
    float u = (1.0f - cosTheta) / (1.0f - cosThetaMax);
    float sinTheta = sqrtf(1.0f - cosTheta * cosTheta);
    if (sinTheta < 1e-5f || (coords.x / sinTheta) < -1.0f || (coords.x / sinTheta) > 1.0f) return false;
    float v = acosf(coords.x / sinTheta) * kInvTwoPi;
    uv = {u, v};
    return true;
}
inline f2 concentricSampleDisk(f2 u)
{
    f2 o = {2.0f * u.x - 1.0f, 2.0f * u.y - 1.0f};
    if (fabsf(o.x) < 1e-10f && fabsf(o.y) < 1e-10f) return {0, 0};
    float theta, r;
    if (fabsf(o.x) > fabsf(o.y)) { r = o.x; theta = kPiOver4 * (o.y / o.x); }
    else { r = o.y; theta = kPiOver2 - kPiOver4 * (o.x / o.y); }
    return {r * cosf(theta), r * sinf(theta)};
}
inline bool refract(f3 &r, f3 i, f3 n, float ior)
{
    f3 nn = n;
    float negNdotV = dot(i, nn);
    float eta;
    if (negNdotV > 0.0f) { eta = ior; nn = -n; negNdotV = -negNdotV; }
    else eta = 1.f / ior;
    const float k = 1.f - eta * eta * (1.f - negNdotV * negNdotV);
    if (k < 0.0f) { r = F3(0.f); return false; }
    r = normalize(eta * i - (eta * negNdotV + sqrtf(k)) * nn);
    return true;
}
inline f3 yawPitchToDir(float yaw, float pitch)
{
    if (std::isnan(yaw) || std::isnan(pitch)) return {0, 0, 1};
    pitch = clampf(pitch, -kPiOver2 + 0.01f, kPiOver2 - 0.01f);
    float sy = sinf(yaw), cy = cosf(yaw), sp = sinf(pitch), cp = cosf(pitch);
    return normalize(F3(sy * cp, sp, cy * cp));
}
inline f2 dirToYawPitch(f3 dir)
{
    // Float3::normalize() member: n = sqrt(compensated length2) (LinearMath.h:536-548)
    float n = length(dir);
    dir = dir / n;
    return {atan2f(dir.x, dir.z), asinf(dir.y)};
}

} // namespace orc
