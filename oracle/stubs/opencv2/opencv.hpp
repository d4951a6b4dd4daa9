// Stand-in for OpenCV (see README.md): only the names the compiled reference sources mention.
#pragma once
#include <cmath>
#include <iostream>   // the real header brings it in; track.cpp relies on that
namespace cv {
struct Point { int x, y; Point(int x_ = 0, int y_ = 0) : x(x_), y(y_) {} };
struct Vec3b { unsigned char v[3]; unsigned char& operator[](int i) { return v[i]; } unsigned char operator[](int i) const { return v[i]; } };
struct Mat {
    int rows = 0, cols = 0;
    template <typename T> T at(Point const&) const { return T(); }
};
}  // namespace cv
