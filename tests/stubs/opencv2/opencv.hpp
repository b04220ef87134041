// TEST INFRASTRUCTURE: the sliver of OpenCV that hnsw_sifts_retrieval/{siftsIndex.hpp,siftsIndex.cpp,makeSearch.cpp} touch, so
// that those reference sources compile without OpenCV (absent from this image; xfeatures2d is a non-free contrib module).
// Not an OpenCV re-implementation: float matrices only, and the "SIFT detector" reads keypoints + descriptors that the real
// cv2.SIFT_create(128) produced for the image (a `<image>.sift` file next to it: int32 n, n x {pt.x, pt.y, angle, size,
// response f32, class_id, octave i32}, n x 128 f32; written by tests/gen_makesearch_fixture.py).
#pragma once
// (the real opencv2/opencv.hpp pulls these standard headers in; makeSearch.cpp relies on that for std::function / std::set)
#include <algorithm>
#include <cmath>
#include <functional>
#include <map>
#include <set>
#include <stdexcept>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <memory>
#include <string>
#include <vector>

#define CV_32FC1 5
#define CV_32F 5
#define CV_REDUCE_SUM 0

namespace cv {

enum { NORM_L2 = 4 };

struct Point2f {
    float x = 0, y = 0;
};

struct KeyPoint {
    Point2f pt;
    float size = 0, angle = -1, response = 0;
    int octave = 0, class_id = -1;
};

class Mat {
public:
    int rows = 0, cols = 0;
    unsigned char* data = nullptr;
    std::string stub_path;  // set by imread: the stub detector finds `<stub_path>.sift`

    Mat() {}
    void create(int r, int c, int /*type*/) {
        rows = r; cols = c;
        buf_.reset(new std::vector<float>((size_t)r * c, 0.0f));
        data = (unsigned char*)buf_->data();
    }
    void release() { buf_.reset(); data = nullptr; rows = cols = 0; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    size_t elemSize() const { return sizeof(float); }
    size_t total() const { return (size_t)rows * cols; }
    template <typename T> T& at(int i, int j) { return ((T*)data)[(size_t)i * cols + j]; }
    template <typename T> T& at(int i) { return ((T*)data)[i]; }
    Mat row(int r) const {  // a header over the same storage, as in OpenCV
        Mat m;
        m.rows = 1; m.cols = cols; m.buf_ = buf_;
        m.data = data + (size_t)r * cols * sizeof(float);
        return m;
    }

private:
    std::shared_ptr<std::vector<float> > buf_;
};

inline Mat abs(const Mat& a) {
    Mat r;
    r.create(a.rows, a.cols, CV_32FC1);
    for (size_t i = 0; i < a.total(); i++) ((float*)r.data)[i] = std::fabs(((const float*)a.data)[i]);
    return r;
}

// cv::reduce(src, dst, 1, CV_REDUCE_SUM, CV_32FC1): row sums (accumulated in double, as OpenCV does for 32F -> 32F sums)
inline void reduce(const Mat& src, Mat& dst, int dim, int /*rtype*/, int /*dtype*/) {
    if (dim != 1) throw std::runtime_error("cv stub: reduce supports dim = 1 only");
    dst.create(src.rows, 1, CV_32FC1);
    for (int r = 0; r < src.rows; r++) {
        double s = 0;
        for (int c = 0; c < src.cols; c++) s += ((const float*)src.data)[(size_t)r * src.cols + c];
        ((float*)dst.data)[r] = (float)s;
    }
}

// cv::normalize(src, dst, alpha, 0, NORM_L2) with dst aliasing src: x * (alpha / ||x||), norm accumulated in double
inline void normalize(const Mat& src, Mat dst, double alpha, double /*beta*/, int norm_type) {
    if (norm_type != NORM_L2) throw std::runtime_error("cv stub: normalize supports NORM_L2 only");
    double n2 = 0;
    for (size_t i = 0; i < src.total(); i++) n2 += (double)((const float*)src.data)[i] * ((const float*)src.data)[i];
    const double nrm = std::sqrt(n2);
    const double scale = nrm > 2.220446049250313e-16 ? alpha / nrm : 0.0;  // DBL_EPSILON guard of cv::normalize
    for (size_t i = 0; i < src.total(); i++) ((float*)dst.data)[i] = (float)(((const float*)src.data)[i] * scale);
}

inline Mat imread(const std::string& path, int /*flags*/ = 1) {
    Mat m;
    std::ifstream f((path + ".sift").c_str(), std::ios::binary);
    if (f.good()) {
        m.create(1, 1, CV_32FC1);  // non-empty placeholder; the pixels are never looked at
        m.stub_path = path;
    }
    return m;
}

template <typename T>
class Ptr {
    std::shared_ptr<T> p_;

public:
    Ptr() {}
    explicit Ptr(T* p) : p_(p) {}
    template <typename U> Ptr(const Ptr<U>& o) : p_(o.shared()) {}
    T* operator->() const { return p_.get(); }
    std::shared_ptr<T> shared() const { return p_; }
};

class Feature2D {
public:
    virtual ~Feature2D() {}
    virtual void detect(const Mat& image, std::vector<KeyPoint>& keypoints) {
        load(image);
        keypoints = kps_;
    }
    virtual void compute(const Mat& image, std::vector<KeyPoint>& keypoints, Mat& descriptors) {
        load(image);
        keypoints = kps_;
        descriptors.create((int)kps_.size(), 128, CV_32FC1);
        if (!desc_.empty()) memcpy(descriptors.data, desc_.data(), desc_.size() * sizeof(float));
    }

private:
    void load(const Mat& image) {
        if (!kps_.empty()) return;
        std::ifstream f((image.stub_path + ".sift").c_str(), std::ios::binary);
        int n = 0;
        f.read((char*)&n, 4);
        kps_.resize(n);
        for (int i = 0; i < n; i++) {
            f.read((char*)&kps_[i].pt.x, 4); f.read((char*)&kps_[i].pt.y, 4); f.read((char*)&kps_[i].angle, 4);
            f.read((char*)&kps_[i].size, 4); f.read((char*)&kps_[i].response, 4); f.read((char*)&kps_[i].class_id, 4);
            f.read((char*)&kps_[i].octave, 4);
        }
        desc_.resize((size_t)n * 128);
        f.read((char*)desc_.data(), desc_.size() * sizeof(float));
    }
    std::vector<KeyPoint> kps_;
    std::vector<float> desc_;
};

}  // namespace cv
