// ORACLE (test infrastructure).  Definitions for oracle/cvshim/opencv2/core/core.hpp on top of the
// cv2-pinned scalar primitives in oracle/prim.hpp.  Semantics that matter to the reference TU:
//   * Mat headers are views onto a shared buffer; operator()/rowRange/colRange never copy.
//   * OutputArray::create on a Mat that already has the requested shape keeps the existing view,
//     so resize()/copyMakeBorder() write THROUGH ROI views (src/ORBextractor.cc:1113-1123).
//   * `m = Mat::zeros(...)` fills in place when shapes match (src/ORBextractor.cc:1037 assigns into a
//     rowRange view of the caller's descriptor matrix).
#include "opencv2/core/core.hpp"
#include "../prim.hpp"

int cvRound(double v) { return fbe_oracle::cv_round(v); }
int cvRound(float v) { return fbe_oracle::cv_round(v); }
int cvRound(int v) { return v; }
int cvFloor(double v) { return fbe_oracle::cv_floor(v); }
int cvCeil(double v) { return fbe_oracle::cv_ceil(v); }

namespace cv {

Mat::Mat() : rows(0), cols(0), step(0), data(nullptr) {}
Mat::Mat(Size sz, int type) : rows(0), cols(0), step(0), data(nullptr) { create(sz.height, sz.width, type); }
Mat::Mat(int r, int c, int type) : rows(0), cols(0), step(0), data(nullptr) { create(r, c, type); }
Mat::Mat(int r, int c, int, void* d, size_t st) : rows(r), cols(c), step(st), data(static_cast<uchar*>(d)) {}
Mat::Mat(const Mat& m) : rows(m.rows), cols(m.cols), step(m.step), data(m.data), buf_(m.buf_) {}
Mat::Mat(const MatZeros& z) : rows(0), cols(0), step(0), data(nullptr) { *this = z; }
Mat::~Mat() {}
Mat& Mat::operator=(const Mat& m) {
    rows = m.rows; cols = m.cols; step = m.step; data = m.data; buf_ = m.buf_;
    return *this;
}
Mat& Mat::operator=(const MatZeros& z) {
    create(z.rows, z.cols, z.type);
    for (int y = 0; y < rows; ++y) std::memset(data + (size_t)y * step, 0, cols);
    return *this;
}
void Mat::create(int r, int c, int type) {
    assert(type == CV_8UC1);
    if (data && rows == r && cols == c) return;
    buf_ = std::make_shared<std::vector<uchar> >((size_t)r * c + 64);
    rows = r; cols = c; step = (size_t)c; data = buf_->data();
}
void Mat::release() { buf_.reset(); rows = cols = 0; step = 0; data = nullptr; }
Mat Mat::operator()(const Rect& r) const {
    assert(r.x >= 0 && r.y >= 0 && r.x + r.width <= cols && r.y + r.height <= rows);
    Mat m(*this);
    m.data = data + (size_t)r.y * step + r.x;
    m.rows = r.height; m.cols = r.width;
    return m;
}
Mat Mat::rowRange(int a, int b) const { return (*this)(Rect(0, a, cols, b - a)); }
Mat Mat::colRange(int a, int b) const { return (*this)(Rect(a, 0, b - a, rows)); }
Mat Mat::clone() const {
    Mat m;
    if (!data) return m;
    m.create(rows, cols, CV_8UC1);
    for (int y = 0; y < rows; ++y) std::memcpy(m.data + (size_t)y * m.step, data + (size_t)y * step, cols);
    return m;
}
int Mat::type() const { return CV_8UC1; }
size_t Mat::step1() const { return step; }
bool Mat::empty() const { return data == nullptr || rows == 0 || cols == 0; }
uchar* Mat::ptr(int y) { return data + (size_t)y * step; }
const uchar* Mat::ptr(int y) const { return data + (size_t)y * step; }
template <> uchar& Mat::at<uchar>(int y, int x) { return data[(size_t)y * step + x]; }
template <> const uchar& Mat::at<uchar>(int y, int x) const { return data[(size_t)y * step + x]; }
MatZeros Mat::zeros(int r, int c, int type) { return MatZeros{r, c, type}; }

_InputArray::_InputArray(const Mat& m) : obj_(const_cast<Mat*>(&m)) {}
Mat _InputArray::getMat() const { return *obj_; }
bool _InputArray::empty() const { return obj_->empty(); }
_OutputArray::_OutputArray(Mat& m) : _InputArray(m) {}
void _OutputArray::create(int r, int c, int type) const { obj_->create(r, c, type); }
void _OutputArray::create(Size sz, int type) const { obj_->create(sz.height, sz.width, type); }
void _OutputArray::release() const { obj_->release(); }

float fastAtan2(float y, float x) { return fbe_oracle::fast_atan2_deg(y, x); }

void copyMakeBorder(InputArray _src, OutputArray _dst, int top, int bottom, int left, int right, int borderType) {
    Mat src = _src.getMat();
    assert((borderType & ~BORDER_ISOLATED) == BORDER_REFLECT_101);
    assert(top == bottom && left == right && top == left);
    // Non-isolated mode would read the parent matrix around a sub-view; every reference caller
    // hands an owning matrix at level 0, so the isolated behaviour is the observable one.
    _dst.create(src.rows + top + bottom, src.cols + left + right, CV_8UC1);
    Mat dst = _dst.getMat();
    fbe_oracle::border_reflect101_u8(src.data, src.cols, src.rows, src.step, dst.data, dst.step, top);
}

void resize(InputArray _src, OutputArray _dst, Size dsize, double, double, int interpolation) {
    assert(interpolation == INTER_LINEAR);
    Mat src = _src.getMat();
    _dst.create(dsize, CV_8UC1);
    Mat dst = _dst.getMat();
    fbe_oracle::resize_linear_u8(src.data, src.cols, src.rows, src.step, dst.data, dst.cols, dst.rows, dst.step);
}

void GaussianBlur(InputArray _src, OutputArray _dst, Size ksize, double sigmaX, double sigmaY, int borderType) {
    assert(ksize.width == 7 && ksize.height == 7 && sigmaX == 2 && sigmaY == 2 && borderType == BORDER_REFLECT_101);
    Mat src = _src.getMat();
    _dst.create(src.rows, src.cols, CV_8UC1);
    Mat dst = _dst.getMat();
    if (dst.data == src.data) {
        Mat tmp = src.clone();
        fbe_oracle::gauss7_u8(tmp.data, tmp.cols, tmp.rows, tmp.step, dst.data, dst.step);
    } else {
        fbe_oracle::gauss7_u8(src.data, src.cols, src.rows, src.step, dst.data, dst.step);
    }
}

void FAST(InputArray _img, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression) {
    assert(nonmaxSuppression);
    Mat img = _img.getMat();
    std::vector<fbe_oracle::FastKp> v;
    fbe_oracle::fast9_nms(img.data, img.cols, img.rows, img.step, threshold, v);
    keypoints.clear();
    for (size_t i = 0; i < v.size(); ++i)
        keypoints.push_back(KeyPoint((float)v[i].x, (float)v[i].y, 7.f, -1.f, (float)v[i].score));
}

// Only referenced from the reference's dead ComputeKeyPointsOld (src/ORBextractor.cc:855-1032).
void KeyPointsFilter::retainBest(std::vector<KeyPoint>& kps, int n) {
    if (n >= 0 && (int)kps.size() > n) {
        std::stable_sort(kps.begin(), kps.end(), [](const KeyPoint& a, const KeyPoint& b) { return a.response > b.response; });
        kps.resize(n);
    }
}

}  // namespace cv
