// ORACLE (test infrastructure).  Minimal stand-in for the slice of the OpenCV 2.4/3.x C++ API that
// /root/reference/src/ORBextractor.cc uses, so that the reference translation unit can be compiled
// VERBATIM (oracle/Makefile -> oracle/_ref/libfbe_ref.so) in an image that has no C++ OpenCV.
// Only 8-bit single-channel matrices exist here.  The numeric primitives behind it live in
// oracle/prim.hpp and are pinned bit-exact against cv2 4.13.0 (tests/test_oracle_prims.py).
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <iterator>
#include <memory>
#include <vector>

typedef unsigned char uchar;

#define CV_8U 0
#define CV_8UC1 0
#define CV_PI 3.1415926535897932384626433832795

int cvRound(double v);
int cvRound(float v);
int cvRound(int v);
int cvFloor(double v);
int cvCeil(double v);

namespace cv {

enum { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_WRAP = 3,
       BORDER_REFLECT_101 = 4, BORDER_DEFAULT = 4, BORDER_ISOLATED = 16 };
enum { INTER_NEAREST = 0, INTER_LINEAR = 1 };

template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T _x, T _y) : x(_x), y(_y) {}
    template <typename U> Point_(const Point_<U>& o) : x((T)o.x), y((T)o.y) {}
    Point_& operator*=(float s) { x = (T)(x * s); y = (T)(y * s); return *this; }
};
typedef Point_<int> Point2i;
typedef Point_<int> Point;
typedef Point_<float> Point2f;

template <typename T> struct Size_ {
    T width, height;
    Size_() : width(0), height(0) {}
    Size_(T w, T h) : width(w), height(h) {}
};
typedef Size_<int> Size;

struct Rect {
    int x, y, width, height;
    Rect() : x(0), y(0), width(0), height(0) {}
    Rect(int _x, int _y, int w, int h) : x(_x), y(_y), width(w), height(h) {}
};

struct KeyPoint {               // 28 bytes, the cv::KeyPoint layout
    Point2f pt;
    float size;
    float angle;
    float response;
    int octave;
    int class_id;
    KeyPoint() : pt(0, 0), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(float x, float y, float _size, float _angle = -1, float _response = 0, int _octave = 0, int _class_id = -1)
        : pt(x, y), size(_size), angle(_angle), response(_response), octave(_octave), class_id(_class_id) {}
};

struct MatZeros { int rows, cols, type; };   // plays the role of the MatExpr returned by Mat::zeros

class Mat {
public:
    int rows, cols;
    size_t step;
    uchar* data;

    Mat();
    Mat(Size sz, int type);
    Mat(int rows, int cols, int type);
    Mat(int rows, int cols, int type, void* data, size_t step);   // header over external memory (not owned)
    Mat(const Mat& m);
    Mat(const MatZeros& z);
    ~Mat();
    Mat& operator=(const Mat& m);
    Mat& operator=(const MatZeros& z);   // like MatExpr assignment: fills IN PLACE when shape matches

    void create(int rows, int cols, int type);
    void release();
    Mat operator()(const Rect& r) const;
    Mat rowRange(int a, int b) const;
    Mat colRange(int a, int b) const;
    Mat clone() const;
    int type() const;
    size_t step1() const;
    bool empty() const;
    uchar* ptr(int y = 0);
    const uchar* ptr(int y = 0) const;
    template <typename T> T& at(int y, int x);
    template <typename T> const T& at(int y, int x) const;
    static MatZeros zeros(int rows, int cols, int type);

private:
    std::shared_ptr<std::vector<uchar> > buf_;
};

template <> uchar& Mat::at<uchar>(int y, int x);
template <> const uchar& Mat::at<uchar>(int y, int x) const;

class _InputArray {
public:
    _InputArray(const Mat& m);
    Mat getMat() const;
    bool empty() const;
protected:
    Mat* obj_;
};

class _OutputArray : public _InputArray {
public:
    _OutputArray(Mat& m);
    void create(int rows, int cols, int type) const;
    void create(Size sz, int type) const;
    void release() const;
};

typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;

float fastAtan2(float y, float x);

void copyMakeBorder(InputArray src, OutputArray dst, int top, int bottom, int left, int right, int borderType);
void resize(InputArray src, OutputArray dst, Size dsize, double fx = 0, double fy = 0, int interpolation = INTER_LINEAR);
void GaussianBlur(InputArray src, OutputArray dst, Size ksize, double sigmaX, double sigmaY = 0, int borderType = BORDER_DEFAULT);
void FAST(InputArray image, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression = true);

class KeyPointsFilter {
public:
    static void retainBest(std::vector<KeyPoint>& keypoints, int npoints);
};

}  // namespace cv
