// Pose.h -- the little of Eigen::Isometry3d the front end uses (reference src/GraphicEnd.h:13-14):
// identity, product, inverse, exact comparison with identity (the failure convention, src/GraphicEnd.cpp:173),
// and conversion to the translation + unit quaternion that g2o's text format stores.
#pragma once
#include <cmath>
#include <cstring>

struct Isometry3d {
    double m[16];   // row-major 4x4
    Isometry3d() { setIdentity(); }
    static Isometry3d Identity() { return Isometry3d(); }
    void setIdentity() { std::memset(m, 0, sizeof(m)); m[0] = m[5] = m[10] = m[15] = 1.0; }
    // T.matrix() == Identity, compared by VALUE like Eigen does: the inverse of the identity carries -0.0 in its translation
    bool isIdentity() const { for (int i = 0; i < 16; ++i) if (m[i] != ((i % 5 == 0) ? 1.0 : 0.0)) return false; return true; }
    double &operator()(int r, int c) { return m[4 * r + c]; }
    double operator()(int r, int c) const { return m[4 * r + c]; }
    Isometry3d operator*(const Isometry3d &o) const
    {
        Isometry3d r;
        for (int i = 0; i < 3; ++i) {
            for (int j = 0; j < 4; ++j) {
                double s = 0;
                for (int k = 0; k < 3; ++k) s += m[4 * i + k] * o.m[4 * k + j];
                r.m[4 * i + j] = s + (j == 3 ? m[4 * i + 3] : 0.0);
            }
        }
        return r;
    }
    Isometry3d inverse() const
    {
        Isometry3d r;
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[4 * i + j] = m[4 * j + i];
        for (int i = 0; i < 3; ++i) r.m[4 * i + 3] = -(r.m[4 * i] * m[3] + r.m[4 * i + 1] * m[7] + r.m[4 * i + 2] * m[11]);
        return r;
    }
    // from a (not necessarily normalised) quaternion (qx,qy,qz,qw) and a translation
    static Isometry3d fromQuaternion(const double q[4], const double t[3])
    {
        double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
        Isometry3d r;
        if (!(n > 0)) { r.m[3] = t[0]; r.m[7] = t[1]; r.m[11] = t[2]; return r; }
        const double x = q[0] / n, y = q[1] / n, z = q[2] / n, w = q[3] / n;
        r.m[0] = 1 - 2 * (y * y + z * z); r.m[1] = 2 * (x * y - z * w);     r.m[2] = 2 * (x * z + y * w);      r.m[3] = t[0];
        r.m[4] = 2 * (x * y + z * w);     r.m[5] = 1 - 2 * (x * x + z * z); r.m[6] = 2 * (y * z - x * w);      r.m[7] = t[1];
        r.m[8] = 2 * (x * z - y * w);     r.m[9] = 2 * (y * z + x * w);     r.m[10] = 1 - 2 * (x * x + y * y); r.m[11] = t[2];
        return r;
    }
    // unit quaternion (qx,qy,qz,qw), qw >= 0
    void quaternion(double q[4]) const
    {
        double tr = m[0] + m[5] + m[10];
        double qw, qx, qy, qz;
        if (tr > 0) { double s = std::sqrt(tr + 1.0) * 2; qw = 0.25 * s; qx = (m[9] - m[6]) / s; qy = (m[2] - m[8]) / s; qz = (m[4] - m[1]) / s; }
        else if (m[0] > m[5] && m[0] > m[10]) { double s = std::sqrt(1.0 + m[0] - m[5] - m[10]) * 2; qw = (m[9] - m[6]) / s; qx = 0.25 * s; qy = (m[1] + m[4]) / s; qz = (m[2] + m[8]) / s; }
        else if (m[5] > m[10]) { double s = std::sqrt(1.0 + m[5] - m[0] - m[10]) * 2; qw = (m[2] - m[8]) / s; qx = (m[1] + m[4]) / s; qy = 0.25 * s; qz = (m[6] + m[9]) / s; }
        else { double s = std::sqrt(1.0 + m[10] - m[0] - m[5]) * 2; qw = (m[4] - m[1]) / s; qx = (m[2] + m[8]) / s; qy = (m[6] + m[9]) / s; qz = 0.25 * s; }
        if (qw < 0) { qw = -qw; qx = -qx; qy = -qy; qz = -qz; }
        q[0] = qx; q[1] = qy; q[2] = qz; q[3] = qw;
    }
};
