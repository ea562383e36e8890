// ORACLE (test infrastructure).  Header-only stand-in for the slice of the OpenCV C++ API that the reference's matcher
// translation unit (src/ORBmatcher.cc) and the class declarations it pulls in (Frame.h, KeyFrame.h, MapPoint*.h, Map.h,
// KeyFrameDatabase.h) need, so that ORBmatcher.cc can be compiled VERBATIM in an image without C++ OpenCV
// (oracle/Makefile -> oracle/_ref/libfbe_refmatch.so).  cv::Mat here is a small dense matrix of 8-bit or 32-bit float
// elements with view semantics for row/col/rowRange/colRange and value semantics for arithmetic (float products are
// accumulated in double like cv::gemm does for CV_32F).  Nothing in the product uses this.
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <list>
#include <map>
#include <memory>
#include <set>
#include <string>
#include <vector>

typedef unsigned char uchar;
#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5
#define CV_32FC1 5
#define CV_PI 3.1415926535897932384626433832795

inline int cvRound(double v) { return (int)std::nearbyint(v); }
inline int cvFloor(double v) { return (int)std::floor(v); }
inline int cvCeil(double v) { return (int)std::ceil(v); }

namespace cv {

enum { NORM_L2 = 4 };

template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T _x, T _y) : x(_x), y(_y) {}
    template <typename U> Point_(const Point_<U>& o) : x((T)o.x), y((T)o.y) {}
};
typedef Point_<int> Point2i;
typedef Point_<int> Point;
typedef Point_<float> Point2f;
template <typename T> struct Point3_ {
    T x, y, z;
    Point3_() : x(0), y(0), z(0) {}
    Point3_(T _x, T _y, T _z) : x(_x), y(_y), z(_z) {}
};
typedef Point3_<float> Point3f;
template <typename T> struct Size_ { T width, height; Size_() : width(0), height(0) {} Size_(T w, T h) : width(w), height(h) {} };
typedef Size_<int> Size;
struct Rect { int x, y, width, height; Rect() : x(0), y(0), width(0), height(0) {} Rect(int a, int b, int c, int d) : x(a), y(b), width(c), height(d) {} };
template <typename T, int N> struct Vec { T val[N]; Vec() { for (int i = 0; i < N; ++i) val[i] = 0; } T& operator[](int i) { return val[i]; } const T& operator[](int i) const { return val[i]; } };
typedef Vec<double, 3> Vec3d;

struct KeyPoint {               // 28 bytes, the cv::KeyPoint layout
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
    KeyPoint() : pt(0, 0), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
};

struct DMatch {
    int queryIdx, trainIdx, imgIdx;
    float distance;
    DMatch() : queryIdx(-1), trainIdx(-1), imgIdx(-1), distance(3.4e38f) {}
    DMatch(int q, int t, float d) : queryIdx(q), trainIdx(t), imgIdx(-1), distance(d) {}
};

class Mat {
public:
    int rows, cols;
    Mat() : rows(0), cols(0), type_(CV_32F), step_(0), data_(nullptr) {}
    Mat(int r, int c, int type) : rows(0), cols(0), type_(type), step_(0), data_(nullptr) { create(r, c, type); }
    explicit Mat(const Point3f& p) : rows(0), cols(0), type_(CV_32F), step_(0), data_(nullptr) {     // 3x1 CV_32F, like cv::Mat(Point3_<T>)
        create(3, 1, CV_32F);
        at<float>(0, 0) = p.x; at<float>(1, 0) = p.y; at<float>(2, 0) = p.z;
    }
    void create(int r, int c, int type) {
        type_ = type; rows = r; cols = c; step_ = (size_t)c * esz();
        buf_ = std::make_shared<std::vector<uchar> >((size_t)r * step_ + 16, 0);
        data_ = buf_->data();
    }
    void release() { buf_.reset(); data_ = nullptr; rows = cols = 0; step_ = 0; }
    int type() const { return type_; }
    bool empty() const { return data_ == nullptr || rows == 0 || cols == 0; }
    size_t esz() const { return type_ == CV_32F ? 4 : 1; }
    Mat clone() const {
        Mat m(rows, cols, type_);
        for (int y = 0; y < rows; ++y)
            for (int x = 0; x < cols; ++x) std::memcpy(m.elem(y, x), elem(y, x), esz());
        return m;
    }
    // like cv::Mat::copyTo: writes THROUGH an existing destination of the same shape (views stay views), else reallocates
    void copyTo(Mat& m) const {
        if (m.data_ && m.rows == rows && m.cols == cols && m.type_ == type_) {
            for (int y = 0; y < rows; ++y)
                for (int x = 0; x < cols; ++x) std::memcpy(m.elem(y, x), elem(y, x), esz());
        } else m = clone();
    }
    void copyTo(const Mat& m) const { copyTo(const_cast<Mat&>(m)); }       // destination given as a temporary view
    // views
    Mat view(int y0, int x0, int r, int c) const {
        Mat m; m.rows = r; m.cols = c; m.type_ = type_; m.step_ = step_; m.buf_ = buf_;
        m.data_ = data_ + (size_t)y0 * step_ + (size_t)x0 * esz();
        return m;
    }
    Mat row(int y) const { return view(y, 0, 1, cols); }
    Mat col(int x) const { return view(0, x, rows, 1); }
    Mat rowRange(int a, int b) const { return view(a, 0, b - a, cols); }
    Mat colRange(int a, int b) const { return view(0, a, rows, b - a); }
    uchar* elem(int y, int x) { return data_ + (size_t)y * step_ + (size_t)x * esz(); }
    const uchar* elem(int y, int x) const { return data_ + (size_t)y * step_ + (size_t)x * esz(); }
    template <typename T> T& at(int y, int x) { return *reinterpret_cast<T*>(elem(y, x)); }
    template <typename T> const T& at(int y, int x) const { return *reinterpret_cast<const T*>(elem(y, x)); }
    template <typename T> T& at(int i) { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }
    template <typename T> const T& at(int i) const { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }
    template <typename T> T* ptr(int y = 0) { return reinterpret_cast<T*>(elem(y, 0)); }
    template <typename T> const T* ptr(int y = 0) const { return reinterpret_cast<const T*>(elem(y, 0)); }
    uchar* ptr(int y = 0) { return elem(y, 0); }
    const uchar* ptr(int y = 0) const { return elem(y, 0); }
    float f(int y, int x) const { return at<float>(y, x); }
    Mat t() const {
        Mat m(cols, rows, type_);
        for (int y = 0; y < rows; ++y)
            for (int x = 0; x < cols; ++x) m.at<float>(x, y) = f(y, x);
        return m;
    }
    double dot(const Mat& o) const {
        double s = 0;
        for (int y = 0; y < rows; ++y)
            for (int x = 0; x < cols; ++x) s += (double)f(y, x) * (double)o.f(y, x);
        return s;
    }
    static Mat zeros(int r, int c, int type) { return Mat(r, c, type); }
    static Mat eye(int r, int c, int type) { Mat m(r, c, type); for (int i = 0; i < std::min(r, c); ++i) m.at<float>(i, i) = 1.f; return m; }
private:
    int type_;
    size_t step_;
    uchar* data_;
    std::shared_ptr<std::vector<uchar> > buf_;
};

inline Mat operator*(const Mat& a, const Mat& b) {
    assert(a.cols == b.rows);
    Mat m(a.rows, b.cols, CV_32F);
    for (int y = 0; y < a.rows; ++y)
        for (int x = 0; x < b.cols; ++x) {
            double s = 0;
            for (int k = 0; k < a.cols; ++k) s += (double)a.f(y, k) * (double)b.f(k, x);
            m.at<float>(y, x) = (float)s;
        }
    return m;
}
inline Mat binop(const Mat& a, const Mat& b, float sb) {
    Mat m(a.rows, a.cols, CV_32F);
    for (int y = 0; y < a.rows; ++y)
        for (int x = 0; x < a.cols; ++x) m.at<float>(y, x) = a.f(y, x) + sb * b.f(y, x);
    return m;
}
inline Mat operator+(const Mat& a, const Mat& b) { return binop(a, b, 1.f); }
inline Mat operator-(const Mat& a, const Mat& b) { return binop(a, b, -1.f); }
inline Mat scale(const Mat& a, double s) {
    Mat m(a.rows, a.cols, CV_32F);
    for (int y = 0; y < a.rows; ++y)
        for (int x = 0; x < a.cols; ++x) m.at<float>(y, x) = (float)((double)a.f(y, x) * s);
    return m;
}
inline Mat operator-(const Mat& a) { return scale(a, -1.0); }
inline Mat operator*(double s, const Mat& a) { return scale(a, s); }
inline Mat operator*(const Mat& a, double s) { return scale(a, s); }
inline Mat operator/(const Mat& a, double s) { return scale(a, 1.0 / s); }
inline double norm(const Mat& a, int = NORM_L2) { return std::sqrt(a.dot(a)); }

// cv::FileStorage / FileNode: DBoW2's TemplatedVocabulary has virtual YAML save / load members that must compile; the
// harness only ever uses loadFromTextFile, so these are inert.
class FileNode {
public:
    FileNode operator[](const std::string&) const { return FileNode(); }
    FileNode operator[](const char*) const { return FileNode(); }
    FileNode operator[](int) const { return FileNode(); }
    size_t size() const { return 0; }
    operator int() const { return 0; }
    operator double() const { return 0.0; }
    operator std::string() const { return std::string(); }
};
class FileStorage {
public:
    enum { READ = 0, WRITE = 1 };
    FileStorage(const char*, int) {}
    FileStorage(const std::string&, int) {}
    bool isOpened() const { return false; }
    void release() {}
    FileNode operator[](const std::string&) const { return FileNode(); }
    FileNode operator[](const char*) const { return FileNode(); }
    template <class T> FileStorage& operator<<(const T&) { return *this; }
};

// declarations only (ORBextractor.h mentions them; the matcher build never calls the extractor)
class _InputArray; class _OutputArray;
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;

}  // namespace cv
