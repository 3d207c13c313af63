// Drop-in for the part of the reference's quaternion.hpp that its pose reporting uses (SLAM.cpp:164-169,284-293):
// public fields x, y, z, w; construction from a 3x3 CV_32F rotation matrix (quaternion.cpp:23-79), from four
// scalars (w first, quaternion.cpp:87-93) or an array; product (:184-192), conjugate, squared norm (:294-297),
// scale and inverse (:325-328).  Scalar host arithmetic, as in the reference; the rest of the 2000-era generic
// class (SHOEMAKE Euler code, stream operators, slerp helpers) is outside the hot path and not provided.
#ifndef QUATERNION_HPP
#define QUATERNION_HPP

#include <math.h>

#include "cv_min.hpp"

inline float SIGN(float x) { return (x >= 0.0f) ? +1.0f : -1.0f; }
inline float NORM(float a, float b, float c, float d) { return sqrtf(a * a + b * b + c * c + d * d); }

class Quaternion {
public:
    Quaternion(void);
    Quaternion(cv::Mat rotationMatrix);
    Quaternion(float wi, float xi, float yi, float zi);
    Quaternion(float v[4]);
    Quaternion operator*(const Quaternion &q);
    bool operator==(const Quaternion &q);
    float norm();
    float magnitude();
    Quaternion scale(float s);
    Quaternion inverse();
    Quaternion conjugate();

    float x, y, z, w;
};

// SLAM.hpp:44-45 / SLAM.cpp:613-648: Euler angles in degrees
void toEulerianAngle(Quaternion q, float &x, float &y, float &z);
void transformationMatToEulerianAngle(cv::Mat t, float &x, float &y, float &z);

#endif
