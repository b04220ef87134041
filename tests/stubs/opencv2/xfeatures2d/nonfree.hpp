// TEST INFRASTRUCTURE: see ../opencv.hpp.  cv::xfeatures2d::SIFT::create(nfeatures) hands out the stub detector that reads
// the descriptors the real SIFT produced for the image.
#pragma once
#include "../opencv.hpp"

namespace cv {
namespace xfeatures2d {
class SIFT : public Feature2D {
public:
    static Ptr<SIFT> create(int /*nfeatures*/ = 0) { return Ptr<SIFT>(new SIFT()); }
};
}  // namespace xfeatures2d
}  // namespace cv
