// Minimal stand-in for the slice of OpenCV 2.4 <opencv2/core/core.hpp> that the ORBextractor / ORBmatcher
// boundary touches (cv::Mat, cv::KeyPoint, cv::Point2f, cv::InputArray / cv::OutputArray).
// TEST STUB ONLY: it exists so that the facades and the call expressions of Frame.cc / Tracking.cc can be
// compile-checked in an image without OpenCV.  A real integration builds against real OpenCV headers.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5

namespace cv {

typedef unsigned char uchar;

template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
};
typedef Point_<int> Point;
typedef Point_<float> Point2f;

struct KeyPoint {  // 28 bytes, same layout as OpenCV 2.4
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
    KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(float x, float y, float s, float a = -1, float r = 0, int o = 0, int c = -1)
        : pt(x, y), size(s), angle(a), response(r), octave(o), class_id(c) {}
};

class Mat {
public:
    int rows, cols;
    size_t step;
    uchar *data;
    Mat() : rows(0), cols(0), step(0), data(nullptr), type_(0) {}
    Mat(int r, int c, int type) : rows(0), cols(0), step(0), data(nullptr), type_(0) { create(r, c, type); }
    Mat(int r, int c, int type, void *ext, size_t step_ = 0) : rows(r), cols(c), step(step_ ? step_ : c * esz(type)), data((uchar *)ext), type_(type) {}
    void create(int r, int c, int type) {
        type_ = type; rows = r; cols = c; step = (size_t)c * esz(type);
        buf_.reset(new std::vector<uchar>((size_t)r * step));
        data = buf_->data();
    }
    void release() { buf_.reset(); data = nullptr; rows = cols = 0; step = 0; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    int type() const { return type_; }
    bool isContinuous() const { return step == (size_t)cols * esz(type_); }
    template <typename T> T *ptr(int r = 0) { return (T *)(data + (size_t)r * step); }
    template <typename T> const T *ptr(int r = 0) const { return (const T *)(data + (size_t)r * step); }
    uchar *ptr(int r = 0) { return data + (size_t)r * step; }
    const uchar *ptr(int r = 0) const { return data + (size_t)r * step; }
    template <typename T> T &at(int r, int c) { return ((T *)(data + (size_t)r * step))[c]; }
    template <typename T> const T &at(int r, int c) const { return ((const T *)(data + (size_t)r * step))[c]; }
    Mat row(int r) const { Mat m; m.rows = 1; m.cols = cols; m.step = step; m.data = data + (size_t)r * step; m.type_ = type_; m.buf_ = buf_; return m; }
    Mat clone() const { Mat m; if (!empty()) { m.create(rows, cols, type_); for (int r = 0; r < rows; r++) memcpy(m.ptr(r), ptr(r), (size_t)cols * esz(type_)); } return m; }
private:
    static size_t esz(int t) { return t == CV_32F ? 4 : 1; }
    int type_;
    std::shared_ptr<std::vector<uchar> > buf_;
};

// InputArray / OutputArray: thin proxies, like OpenCV's _InputArray / _OutputArray
class _InputArray {
public:
    _InputArray(const Mat &m) : m_(&m) {}
    Mat getMat() const { return *m_; }
    bool empty() const { return m_->empty(); }
protected:
    const Mat *m_;
};
class _OutputArray : public _InputArray {
public:
    _OutputArray(Mat &m) : _InputArray(m), w_(&m) {}
    void create(int r, int c, int type) const { w_->create(r, c, type); }
    void release() const { w_->release(); }
    Mat getMat() const { return *w_; }
private:
    Mat *w_;
};
typedef const _InputArray &InputArray;
typedef const _OutputArray &OutputArray;

}  // namespace cv
