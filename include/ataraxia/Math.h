// Math.h — the small slice of vector/matrix math the hot path's host side needs,
// written so that every function reproduces, operation for operation, the
// evaluation order of the glm 1.0.2 routines the reference calls (file:line of
// /root/reference/thirdparty/glm/glm cited per function). That order is what makes
// camera matrices, primary rays and flattened spheres bit-identical to the
// reference's. Compile host code WITHOUT -ffast-math and
// without FMA contraction (-ffp-contract=off): the reference's host TUs are built by
// plain g++ -O3 for baseline x86-64.
//
// The types are layout-compatible with glm::vec2/3/4, glm::quat (x,y,z,w storage)
// and glm::mat4 (column-major), so reference user code can keep glm by defining
// ATX_WITH_GLM (then the public API uses the real glm types); without glm the
// names glm::vec3 etc. alias the types below so the same source compiles.
#pragma once

#include <cmath>
#include <cstdint>

namespace atx
{
    struct vec2
    {
        float x, y;
        constexpr vec2() : x(0), y(0) {}
        constexpr explicit vec2(float s) : x(s), y(s) {}
        constexpr vec2(float x_, float y_) : x(x_), y(y_) {}
    };

    struct vec3
    {
        union { float x; float r; };
        union { float y; float g; };
        union { float z; float b; };
        constexpr vec3() : x(0), y(0), z(0) {}
        constexpr explicit vec3(float s) : x(s), y(s), z(s) {}
        constexpr vec3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
        float& operator[](int i) { return (&x)[i]; }
        const float& operator[](int i) const { return (&x)[i]; }
        vec3& operator+=(const vec3& o) { x += o.x; y += o.y; z += o.z; return *this; }
        vec3& operator-=(const vec3& o) { x -= o.x; y -= o.y; z -= o.z; return *this; }
        vec3& operator*=(float s) { x *= s; y *= s; z *= s; return *this; }
        vec3& operator/=(float s) { x /= s; y /= s; z /= s; return *this; } // true divisions, type_vec3.inl:582-585
    };

    struct vec4
    {
        float x, y, z, w;
        constexpr vec4() : x(0), y(0), z(0), w(0) {}
        constexpr explicit vec4(float s) : x(s), y(s), z(s), w(s) {}
        constexpr vec4(float x_, float y_, float z_, float w_) : x(x_), y(y_), z(z_), w(w_) {}
        constexpr vec4(const vec3& v, float w_) : x(v.x), y(v.y), z(v.z), w(w_) {}
        float& operator[](int i) { return (&x)[i]; }
        const float& operator[](int i) const { return (&x)[i]; }
    };

    // glm::quat stores x,y,z,w but its value constructor takes (w, x, y, z)
    // (detail/type_quat.inl); a value-initialised glm::quat() is all zeros, which
    // mat3_cast maps to the identity (gtc/quaternion.inl:47-72).
    struct quat
    {
        float x, y, z, w;
        constexpr quat() : x(0), y(0), z(0), w(0) {}
        constexpr quat(float w_, float x_, float y_, float z_) : x(x_), y(y_), z(z_), w(w_) {}
    };

    struct mat4
    {
        vec4 c[4]; // columns
        constexpr mat4() : c{ vec4(), vec4(), vec4(), vec4() } {}
        constexpr explicit mat4(float d) : c{ vec4(d, 0, 0, 0), vec4(0, d, 0, 0), vec4(0, 0, d, 0), vec4(0, 0, 0, d) } {}
        vec4& operator[](int i) { return c[i]; }
        const vec4& operator[](int i) const { return c[i]; }
    };

    inline bool operator==(const vec3& a, const vec3& b) { return a.x == b.x && a.y == b.y && a.z == b.z; }

    inline vec2 operator*(const vec2& v, float s) { return { v.x * s, v.y * s }; }
    inline vec2 operator-(const vec2& v, float s) { return { v.x - s, v.y - s }; }

    inline vec3 operator+(const vec3& a, const vec3& b) { return { a.x + b.x, a.y + b.y, a.z + b.z }; }
    inline vec3 operator-(const vec3& a, const vec3& b) { return { a.x - b.x, a.y - b.y, a.z - b.z }; }
    inline vec3 operator-(const vec3& a) { return { -a.x, -a.y, -a.z }; }
    inline vec3 operator*(const vec3& a, const vec3& b) { return { a.x * b.x, a.y * b.y, a.z * b.z }; }
    inline vec3 operator*(const vec3& a, float s) { return { a.x * s, a.y * s, a.z * s }; }
    inline vec3 operator*(float s, const vec3& a) { return { s * a.x, s * a.y, s * a.z }; }
    inline vec3 operator/(const vec3& a, float s) { return { a.x / s, a.y / s, a.z / s }; }

    inline vec4 operator+(const vec4& a, const vec4& b) { return { a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w }; }
    inline vec4 operator-(const vec4& a, const vec4& b) { return { a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w }; }
    inline vec4 operator*(const vec4& a, const vec4& b) { return { a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w }; }
    inline vec4 operator*(const vec4& a, float s) { return { a.x * s, a.y * s, a.z * s, a.w * s }; }

    // detail/func_geometric.inl:48-54 — tmp = a*b; (tmp.x + tmp.y) + tmp.z
    inline float dot(const vec3& a, const vec3& b)
    {
        const vec3 tmp = a * b;
        return tmp.x + tmp.y + tmp.z;
    }

    // detail/func_geometric.inl:8-14
    inline float length(const vec3& v) { return std::sqrt(dot(v, v)); }

    // detail/func_exponential.inl:134-139 — 1 / sqrt(x), both IEEE
    inline float inversesqrt(float x) { return 1.0f / std::sqrt(x); }

    // detail/func_geometric.inl:98-105 — v * inversesqrt(dot(v, v))
    inline vec3 normalize(const vec3& v) { return v * inversesqrt(dot(v, v)); }

    // detail/func_geometric.inl:73-83
    inline vec3 cross(const vec3& x, const vec3& y)
    {
        return { x.y * y.z - y.y * x.z,
                 x.z * y.x - y.z * x.x,
                 x.x * y.y - y.x * x.y };
    }

    // ---- quaternion slice used by Camera::onUpdate (Camera.cpp:88-96) ----
    inline vec2 operator-(const vec2& a, const vec2& b) { return { a.x - b.x, a.y - b.y }; }

    // ext/quaternion_trigonometric.inl:30-36 — s = sin(a/2) first, then (cos(a/2), v*s)
    inline quat angleAxis(float angle, const vec3& v)
    {
        const float s = std::sin(angle * 0.5f);
        const vec3 vs = v * s;
        return quat(std::cos(angle * 0.5f), vs.x, vs.y, vs.z);
    }

    // ext/quaternion_geometric.inl:27-34
    inline quat cross(const quat& q1, const quat& q2)
    {
        return quat(q1.w * q2.w - q1.x * q2.x - q1.y * q2.y - q1.z * q2.z,
                    q1.w * q2.x + q1.x * q2.w + q1.y * q2.z - q1.z * q2.y,
                    q1.w * q2.y + q1.y * q2.w + q1.z * q2.x - q1.x * q2.z,
                    q1.w * q2.z + q1.z * q2.w + q1.x * q2.y - q1.y * q2.x);
    }

    // detail/type_quat.inl:17-24 — (w*w + x*x) + (y*y + z*z)
    inline float dot(const quat& a, const quat& b) { return (a.w * b.w + a.x * b.x) + (a.y * b.y + a.z * b.z); }

    // ext/quaternion_geometric.inl:11-24
    inline quat normalize(const quat& q)
    {
        const float len = std::sqrt(dot(q, q));
        if (len <= 0.0f)
            return quat(1.0f, 0.0f, 0.0f, 0.0f);
        const float oneOverLen = 1.0f / len;
        return quat(q.w * oneOverLen, q.x * oneOverLen, q.y * oneOverLen, q.z * oneOverLen);
    }

    // detail/type_quat.inl:359-366 (what gtx/quaternion.inl:51-54 rotate() returns): v + ((uv*w) + uuv) * 2
    inline vec3 rotate(const quat& q, const vec3& v)
    {
        const vec3 qv(q.x, q.y, q.z);
        const vec3 uv = cross(qv, v);
        const vec3 uuv = cross(qv, uv);
        return v + ((uv * q.w) + uuv) * 2.0f;
    }

    // detail/func_trigonometric.inl:9-14
    inline float radians(float degrees) { return degrees * static_cast<float>(0.01745329251994329576923690768489); }

    // detail/type_mat4x4.inl:562-573 — (m0*v0 + m1*v1) + (m2*v2 + m3*v3)
    inline vec4 operator*(const mat4& m, const vec4& v)
    {
        const vec4 Mul0 = m[0] * vec4(v.x);
        const vec4 Mul1 = m[1] * vec4(v.y);
        const vec4 Add0 = Mul0 + Mul1;
        const vec4 Mul2 = m[2] * vec4(v.z);
        const vec4 Mul3 = m[3] * vec4(v.w);
        const vec4 Add1 = Mul2 + Mul3;
        return Add0 + Add1;
    }

    // detail/type_mat4x4.inl mul4x4<T,Q,false>::call — tmp = A0*b.x; tmp += A1*b.y; tmp += A2*b.z; tmp += A3*b.w
    inline mat4 operator*(const mat4& m1, const mat4& m2)
    {
        mat4 r;
        for (int j = 0; j < 4; j++)
        {
            vec4 tmp = m1[0] * m2[j].x;
            tmp = tmp + m1[1] * m2[j].y;
            tmp = tmp + m1[2] * m2[j].z;
            tmp = tmp + m1[3] * m2[j].w;
            r[j] = tmp;
        }
        return r;
    }

    inline mat4 operator*(const mat4& m, float s)
    {
        mat4 r;
        for (int j = 0; j < 4; j++) r[j] = m[j] * s;
        return r;
    }

    // ext/matrix_transform.inl:10-15
    inline mat4 translate(const mat4& m, const vec3& v)
    {
        mat4 r = m;
        r[3] = m[0] * v.x + m[1] * v.y + m[2] * v.z + m[3];
        return r;
    }

    // ext/matrix_transform.inl:78-86
    inline mat4 scale(const mat4& m, const vec3& v)
    {
        mat4 r;
        r[0] = m[0] * v.x;
        r[1] = m[1] * v.y;
        r[2] = m[2] * v.z;
        r[3] = m[3];
        return r;
    }

    // gtc/quaternion.inl:47-72 (mat3_cast) + :75-78 (mat4_cast) + type_mat4x4.inl:190-201
    inline mat4 mat4_cast(const quat& q)
    {
        const float qxx(q.x * q.x), qyy(q.y * q.y), qzz(q.z * q.z);
        const float qxz(q.x * q.z), qxy(q.x * q.y), qyz(q.y * q.z);
        const float qwx(q.w * q.x), qwy(q.w * q.y), qwz(q.w * q.z);
        mat4 r(1.0f);
        r[0][0] = 1.0f - 2.0f * (qyy + qzz);
        r[0][1] = 2.0f * (qxy + qwz);
        r[0][2] = 2.0f * (qxz - qwy);
        r[1][0] = 2.0f * (qxy - qwz);
        r[1][1] = 1.0f - 2.0f * (qxx + qzz);
        r[1][2] = 2.0f * (qyz + qwx);
        r[2][0] = 2.0f * (qxz + qwy);
        r[2][1] = 2.0f * (qyz - qwx);
        r[2][2] = 1.0f - 2.0f * (qxx + qyy);
        return r;
    }

    // ext/matrix_clip_space.inl:249-262 (perspectiveRH_NO — glm's default clip control)
    inline mat4 perspective(float fovy, float aspect, float zNear, float zFar)
    {
        const float tanHalfFovy = std::tan(fovy / 2.0f);
        mat4 r(0.0f);
        r[0][0] = 1.0f / (aspect * tanHalfFovy);
        r[1][1] = 1.0f / (tanHalfFovy);
        r[2][2] = -(zFar + zNear) / (zFar - zNear);
        r[2][3] = -1.0f;
        r[3][2] = -(2.0f * zFar * zNear) / (zFar - zNear);
        return r;
    }

    // ext/matrix_transform.inl:153-173 (lookAtRH)
    inline mat4 lookAt(const vec3& eye, const vec3& center, const vec3& up)
    {
        const vec3 f(normalize(center - eye));
        const vec3 s(normalize(cross(f, up)));
        const vec3 u(cross(s, f));
        mat4 r(1.0f);
        r[0][0] = s.x; r[1][0] = s.y; r[2][0] = s.z;
        r[0][1] = u.x; r[1][1] = u.y; r[2][1] = u.z;
        r[0][2] = -f.x; r[1][2] = -f.y; r[2][2] = -f.z;
        r[3][0] = -dot(s, eye);
        r[3][1] = -dot(u, eye);
        r[3][2] = dot(f, eye);
        return r;
    }

    // detail/func_matrix.inl:388-446 (compute_inverse<4,4>)
    inline mat4 inverse(const mat4& m)
    {
        const float Coef00 = m[2][2] * m[3][3] - m[3][2] * m[2][3];
        const float Coef02 = m[1][2] * m[3][3] - m[3][2] * m[1][3];
        const float Coef03 = m[1][2] * m[2][3] - m[2][2] * m[1][3];
        const float Coef04 = m[2][1] * m[3][3] - m[3][1] * m[2][3];
        const float Coef06 = m[1][1] * m[3][3] - m[3][1] * m[1][3];
        const float Coef07 = m[1][1] * m[2][3] - m[2][1] * m[1][3];
        const float Coef08 = m[2][1] * m[3][2] - m[3][1] * m[2][2];
        const float Coef10 = m[1][1] * m[3][2] - m[3][1] * m[1][2];
        const float Coef11 = m[1][1] * m[2][2] - m[2][1] * m[1][2];
        const float Coef12 = m[2][0] * m[3][3] - m[3][0] * m[2][3];
        const float Coef14 = m[1][0] * m[3][3] - m[3][0] * m[1][3];
        const float Coef15 = m[1][0] * m[2][3] - m[2][0] * m[1][3];
        const float Coef16 = m[2][0] * m[3][2] - m[3][0] * m[2][2];
        const float Coef18 = m[1][0] * m[3][2] - m[3][0] * m[1][2];
        const float Coef19 = m[1][0] * m[2][2] - m[2][0] * m[1][2];
        const float Coef20 = m[2][0] * m[3][1] - m[3][0] * m[2][1];
        const float Coef22 = m[1][0] * m[3][1] - m[3][0] * m[1][1];
        const float Coef23 = m[1][0] * m[2][1] - m[2][0] * m[1][1];

        const vec4 Fac0(Coef00, Coef00, Coef02, Coef03);
        const vec4 Fac1(Coef04, Coef04, Coef06, Coef07);
        const vec4 Fac2(Coef08, Coef08, Coef10, Coef11);
        const vec4 Fac3(Coef12, Coef12, Coef14, Coef15);
        const vec4 Fac4(Coef16, Coef16, Coef18, Coef19);
        const vec4 Fac5(Coef20, Coef20, Coef22, Coef23);

        const vec4 Vec0(m[1][0], m[0][0], m[0][0], m[0][0]);
        const vec4 Vec1(m[1][1], m[0][1], m[0][1], m[0][1]);
        const vec4 Vec2(m[1][2], m[0][2], m[0][2], m[0][2]);
        const vec4 Vec3(m[1][3], m[0][3], m[0][3], m[0][3]);

        const vec4 Inv0(Vec1 * Fac0 - Vec2 * Fac1 + Vec3 * Fac2);
        const vec4 Inv1(Vec0 * Fac0 - Vec2 * Fac3 + Vec3 * Fac4);
        const vec4 Inv2(Vec0 * Fac1 - Vec1 * Fac3 + Vec3 * Fac5);
        const vec4 Inv3(Vec0 * Fac2 - Vec1 * Fac4 + Vec2 * Fac5);

        const vec4 SignA(+1, -1, +1, -1);
        const vec4 SignB(-1, +1, -1, +1);
        mat4 Inverse;
        Inverse[0] = Inv0 * SignA;
        Inverse[1] = Inv1 * SignB;
        Inverse[2] = Inv2 * SignA;
        Inverse[3] = Inv3 * SignB;

        const vec4 Row0(Inverse[0][0], Inverse[1][0], Inverse[2][0], Inverse[3][0]);
        const vec4 Dot0(m[0] * Row0);
        const float Dot1 = (Dot0.x + Dot0.y) + (Dot0.z + Dot0.w);
        const float OneOverDeterminant = 1.0f / Dot1;
        return Inverse * OneOverDeterminant;
    }
} // namespace atx

#ifndef ATX_WITH_GLM
// Reference user code spells these glm::vec3 etc.; keep that source compiling.
namespace glm
{
    using vec2 = atx::vec2;
    using vec3 = atx::vec3;
    using vec4 = atx::vec4;
    using quat = atx::quat;
    using mat4 = atx::mat4;
    using atx::cross;
    using atx::dot;
    using atx::inverse;
    using atx::length;
    using atx::lookAt;
    using atx::mat4_cast;
    using atx::normalize;
    using atx::perspective;
    using atx::radians;
    using atx::angleAxis;
    using atx::rotate;
    using atx::scale;
    using atx::translate;
}
#endif
