// TEST INFRASTRUCTURE.  Minimal stand-in for <opencv2/opencv.hpp> so that the reference's own
// examples/YOLO-Master-Cross-Platform-Edge-Deployment/cpp/src/common.cpp compiles in this image (no OpenCV headers here)
// into oracle/_ref/libcwnms_ref.so.  Only the VALUE types that the NMS path (common.cpp:56-207) computes with are real:
// cv::Rect_<T> {x, y, width, height, area(), empty(), operator&} with the published semantics of OpenCV 4.x
// (modules/core/include/opencv2/core/types.hpp, Rect_ and `operator&=`).  Everything image-related (Mat, resize, drawing)
// is a declaration-level stub that aborts if ever called: the checker calls decode_candidates / nms_and_cap only.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>

#define CV_8UC4 24
static inline int cvRound(double v) { return static_cast<int>(std::lrint(v)); }

namespace cv {

[[noreturn]] inline void shim_abort_(const char* what) {
    std::fprintf(stderr, "oracle/cv_shim: %s is not available (NMS-only shim)\n", what);
    std::abort();
}

template <typename T> struct Point_ { T x{}, y{}; Point_() = default; Point_(T x_, T y_) : x(x_), y(y_) {} };
using Point = Point_<int>;
template <typename T> struct Size_ { T width{}, height{}; Size_() = default; Size_(T w, T h) : width(w), height(h) {} };
using Size = Size_<int>;

template <typename T> struct Rect_ {
    T x{}, y{}, width{}, height{};
    Rect_() = default;
    Rect_(T x_, T y_, T w_, T h_) : x(x_), y(y_), width(w_), height(h_) {}
    T area() const { return width * height; }
    bool empty() const { return width <= 0 || height <= 0; }
};
using Rect = Rect_<int>;
using Rect2f = Rect_<float>;
using Rect2d = Rect_<double>;

// OpenCV 4.x intersection: the rectangle starting later on each axis gives the origin; the extent is what is left of the earlier one
template <typename T> inline Rect_<T>& operator&=(Rect_<T>& a, const Rect_<T>& b) {
    if (a.empty() || b.empty()) { a = Rect_<T>(); return a; }
    const Rect_<T>& Rx_min = (a.x < b.x) ? a : b;
    const Rect_<T>& Rx_max = (a.x < b.x) ? b : a;
    const Rect_<T>& Ry_min = (a.y < b.y) ? a : b;
    const Rect_<T>& Ry_max = (a.y < b.y) ? b : a;
    if ((Rx_min.x < 0 && Rx_min.x + Rx_min.width < Rx_max.x) || (Ry_min.y < 0 && Ry_min.y + Ry_min.height < Ry_max.y)) {
        a = Rect_<T>();
        return a;
    }
    const T w = std::min(Rx_min.width - (Rx_max.x - Rx_min.x), Rx_max.width);
    const T h = std::min(Ry_min.height - (Ry_max.y - Ry_min.y), Ry_max.height);
    a = Rect_<T>(Rx_max.x, Ry_max.y, w, h);
    if (a.empty()) a = Rect_<T>();
    return a;
}
template <typename T> inline Rect_<T> operator&(const Rect_<T>& a, const Rect_<T>& b) { Rect_<T> c = a; return c &= b; }

struct Scalar { double v[4]; Scalar(double a = 0, double b = 0, double c = 0, double d = 0) : v{a, b, c, d} {} };

struct Mat {
    int rows = 0, cols = 0;
    Mat() = default;
    Mat(int, int, int, const Scalar&) { shim_abort_("cv::Mat"); }
    int type() const { return 0; }
    Mat operator()(const Rect&) const { shim_abort_("cv::Mat::operator()"); }
    void copyTo(const Mat&) const { shim_abort_("cv::Mat::copyTo"); }
    template <typename T> T* ptr(int) { shim_abort_("cv::Mat::ptr"); }
};

enum { FILLED = -1, FONT_HERSHEY_SIMPLEX = 0 };
inline void resize(const Mat&, Mat&, Size) { shim_abort_("cv::resize"); }
inline void rectangle(Mat&, Rect, const Scalar&, int = 1) { shim_abort_("cv::rectangle"); }
inline Size getTextSize(const std::string&, int, double, int, int*) { shim_abort_("cv::getTextSize"); }
inline void putText(Mat&, const std::string&, Point, int, double, Scalar, int = 1) { shim_abort_("cv::putText"); }

}  // namespace cv
